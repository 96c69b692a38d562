--  Patch to GID.Decoding_JPG.Load (gid-decoding_jpg.adb, GID 013) that routes the
--  pixel reconstruction through GID.CUDA.  These are the NEW nested declarations and
--  the replacement bodies; everything not shown (segment readers, bit buffer,
--  Get_VLC, Check_Restart, the progressive scan decoder, Read_SOS) is unchanged.
--  Line numbers refer to the unpatched file.
--
--  Reviewed, not compiled (no Ada toolchain in this repository's build image); the
--  C++ host mirror gid_b200/host/gid_jpeg_host.cpp performs the identical call
--  sequence and is what the tests exercise.
--
--  with GID.CUDA;  with System.Storage_Elements;   --  added to the context clause

    ------------------------------------------------------------------
    --  New state, declared next to info_A / info_B (around :550)
    ------------------------------------------------------------------

    type Coefficient_Plane is array (Natural_32 range <>) of Interfaces.Integer_16;
    pragma Convention (C, Coefficient_Plane);

    type Plane_Info is record
      base     : System.Address := System.Null_Address;  --  pinned, owned by the library
      blocks_x : Natural := 0;                            --  MCU-padded block grid
      blocks_y : Natural := 0;
    end record;

    cuda_ctx   : GID.CUDA.Context := GID.CUDA.Null_Context;
    cuda_job   : GID.CUDA.Job     := GID.CUDA.Null_Job;
    plane      : array (Component) of Plane_Info;
    comp_index : array (Component) of Natural := (others => 0);  --  Y -> 0, Cb -> 1, Cr -> 2
    last_row   : array (Component) of Natural := (others => 0);  --  last block row a coefficient was stored in
    --  largest magnitude stored per component (the DC term with all the bits of dc_predictor, :766-771),
    --  held against GID.CUDA.Fast_Coef_Limit before Submit (see Reconstruct_On_Device_And_Output)
    max_mag    : array (Component) of Natural := (others => 0);
    coef_limit : array (Component) of Natural := (others => 0);

    --  Called once from Read_SOS, after the MCU geometry is known (:1974) and before
    --  the first scan is decoded.  Replaces Create_Image_Array (:1857-1883) too: the
    --  progressive accumulators now ARE the int16 planes.
    procedure Begin_Device_Frame is
      desc : GID.CUDA.Frame_Desc;
      idx  : Natural := 0;
    begin
      desc.width      := Interfaces.Integer_32 (image.width);
      desc.height     := Interfaces.Integer_32 (image.height);
      desc.flags      := GID.CUDA.Out_RGB8 + GID.CUDA.Flag_Packed_Rows;
      desc.out_stride := 0;
      for c in Component loop
        if image.JPEG_stuff.compo_set (c) then
          comp_index (c) := idx;
          desc.samp_h (idx) := Interfaces.Integer_32 (info_A (c).ups.samples_hor);
          desc.samp_v (idx) := Interfaces.Integer_32 (info_A (c).ups.samples_ver);
          --  Natural-order table, as Finalize_Progressive_DCT_Decoding builds qt_zz (:1771-1775)
          for i in 0 .. 63 loop
            desc.qt (idx)(i) :=
              Interfaces.Unsigned_16
                (image.JPEG_stuff.qt_list (info_A (c).qt_assoc)(undo_zig_zag (i)));
          end loop;
          idx := idx + 1;
        end if;
      end loop;
      desc.ncomp := Interfaces.Integer_32 (idx);
      GID.CUDA.Create (device => 0, ctx => cuda_ctx);
      GID.CUDA.Job_Begin (cuda_ctx, desc, cuda_job);
      for c in Component loop
        if image.JPEG_stuff.compo_set (c) then
          coef_limit (c) := GID.CUDA.Fast_Coef_Limit (desc, comp_index (c));
          plane (c).base :=
            GID.CUDA.Plane (cuda_job, comp_index (c), plane (c).blocks_x, plane (c).blocks_y);
        end if;
      end loop;
    end Begin_Device_Frame;

    --  Address of coefficient `nat` (natural order 0 .. 63) of block (bx, by) of c.
    function Coefficient_Address (c : Component; bx, by, nat : Natural) return System.Address is
      use System.Storage_Elements;
    begin
      return plane (c).base +
        Storage_Offset (2 * (64 * (by * plane (c).blocks_x + bx) + nat));
    end Coefficient_Address;

    ------------------------------------------------------------------
    --  Decode_8x8_Block (:758-788): same entropy decoding, but the value is
    --  stored UN-dequantised at zig_zag (coef); the multiplication by
    --  qt_local (coef) (:771, :785) moves to the device.
    ------------------------------------------------------------------
    procedure Decode_8x8_Block (c : Component; bx, by : Natural) is
      value, coef : Integer;
      code : U8;
      procedure Store (nat : Natural; v : Integer) is
        cell : Interfaces.Integer_16;
        for cell'Address use Coefficient_Address (c, bx, by, nat);
        pragma Import (Ada, cell);
      begin
        max_mag (c) := Natural'Max (max_mag (c), abs v);
        if abs v > 32767 then
          --  (a DC predictor driven out of 16 bits by a crafted stream: the int16 planes cannot carry it)
          raise unsupported_image_subformat with "JPEG: a coefficient of the stream exceeds 16 bits";
        end if;
        cell := Interfaces.Integer_16 (v);  --  planes were zeroed by Job_Begin
        last_row (c) := Natural'Max (last_row (c), nat / 8);
      end Store;
    begin
      Get_VLC (image.JPEG_stuff.vlc_defs (DC, info_B (c).ht_idx_DC).all, code, value);
      info_B (c).dc_predictor := info_B (c).dc_predictor + value;
      Store (0, info_B (c).dc_predictor);
      coef := 0;
      loop
        Get_VLC (image.JPEG_stuff.vlc_defs (AC, info_B (c).ht_idx_AC).all, code, value);
        exit when code = 0;
        if (code and 16#0F#) = 0 and code /= 16#F0# then
          raise error_in_image_data with "JPEG: error in VLC AC code for de-quantization";
        end if;
        coef := coef + Integer (Shift_Right (code, 4)) + 1;
        if coef > 63 then
          raise error_in_image_data with "JPEG: coefficient for de-quantization is > 63";
        end if;
        Store (zig_zag (coef), value);
        exit when coef = 63;
      end loop;
    end Decode_8x8_Block;

    ------------------------------------------------------------------
    --  Sparse variant of the above (the one the C++ host mirror uses by default):
    --  instead of storing into the dense plane, append one record per non-zero
    --  value; 128 bytes per block on the host link become 4 bytes per non-zero
    --  coefficient + 4 per block.  State per component, set up in
    --  Begin_Device_Frame with GID.CUDA.Records:
    --     rec_first (c), rec_data (c) : System.Address;  rec_cap (c), rec_n (c), rec_seq (c) : Natural
    --  If rec_n (c) + 64 > rec_cap (c) at the start of a block the component switches
    --  to its dense plane for good: the records emitted so far are replayed into the
    --  plane (block of sequence number s = the s-th block this loop visited) and
    --  Decode_8x8_Block above takes over.  After the scan:
    --     first (nblocks) := rec_n (c);  GID.CUDA.Records_Commit (cuda_job, comp_index (c), rec_n (c));
    --  Progressive files keep accumulating in (ordinary, not pinned) int16 planes and
    --  compact them the same way in Finalize_Progressive_DCT_Decoding.
    ------------------------------------------------------------------
    procedure Decode_8x8_Block_Sparse (c : Component) is
      value, coef : Integer;
      code : U8;
      procedure Emit (nat : Natural; v : Integer) is
        type Record_Array is array (Natural) of Interfaces.Unsigned_32;
        data : Record_Array;
        for data'Address use rec_data (c);
        pragma Import (Ada, data);
        use Interfaces;
      begin
        if v /= 0 then
          data (rec_n (c)) :=
            Shift_Left (Unsigned_32 (Unsigned_16'Mod (v)), 16) or Unsigned_32 (nat);
          rec_n (c) := rec_n (c) + 1;
        end if;
      end Emit;
    begin
      Set_First (c, rec_seq (c), rec_n (c));   --  first (seq) := n
      rec_seq (c) := rec_seq (c) + 1;
      Get_VLC (image.JPEG_stuff.vlc_defs (DC, info_B (c).ht_idx_DC).all, code, value);
      info_B (c).dc_predictor := info_B (c).dc_predictor + value;
      Emit (0, info_B (c).dc_predictor);
      coef := 0;
      loop
        Get_VLC (image.JPEG_stuff.vlc_defs (AC, info_B (c).ht_idx_AC).all, code, value);
        exit when code = 0;
        if (code and 16#0F#) = 0 and code /= 16#F0# then
          raise error_in_image_data with "JPEG: error in VLC AC code for de-quantization";
        end if;
        coef := coef + Integer (Shift_Right (code, 4)) + 1;
        if coef > 63 then
          raise error_in_image_data with "JPEG: coefficient for de-quantization is > 63";
        end if;
        Emit (zig_zag (coef), value);
        exit when coef = 63;
      end loop;
    end Decode_8x8_Block_Sparse;

    ------------------------------------------------------------------
    --  Steps 3-6 for the whole frame: replaces Inverse_DCT_8x8_Block (:1199),
    --  Upsampling_and_Output (:1206) and, for progressive files, the body of
    --  Finalize_Progressive_DCT_Decoding (:1789-1850).
    ------------------------------------------------------------------
    procedure Reconstruct_On_Device_And_Output (feedback_base, feedback_span : Natural) is
      pixels : System.Address;
      stride : Natural;
    begin
      for c in Component loop
        if image.JPEG_stuff.compo_set (c) then
          GID.CUDA.Hint_Rows (cuda_job, comp_index (c), last_row (c) + 1);
          if max_mag (c) > coef_limit (c) then
            --  outside the range of the fast kernel instances: the exact ones take the frame
            GID.CUDA.Wide_Coefficients (cuda_job);
          end if;
        end if;
      end loop;
      GID.CUDA.Submit (cuda_job);
      GID.CUDA.Wait (cuda_job, pixels, stride);
      declare
        type Byte_Array is array (0 .. stride * Natural (image.height) - 1) of Interfaces.Unsigned_8;
        rgb : Byte_Array;
        for rgb'Address use pixels;
        pragma Import (Ada, rgb);
        idx : Natural;
      begin
        --  Top-down rows; GID's y axis counts from the bottom (:1024).
        for row in 0 .. Natural (image.height) - 1 loop
          Set_X_Y (0, Natural (image.height) - 1 - row);
          idx := row * stride;
          for x in 0 .. Natural (image.width) - 1 loop
            Out_Pixel_8 (rgb (idx), rgb (idx + 1), rgb (idx + 2));  --  :909-938 (8 / 16-bit ranges)
            idx := idx + 3;
          end loop;
          Feedback (feedback_base + (feedback_span * (row + 1)) / Natural (image.height));
        end loop;
      end;
      GID.CUDA.Release (cuda_job);
      GID.CUDA.Destroy (cuda_ctx);
      --  (Better, and what the C++ mirror does: keep the context in the Image_Descriptor -- a
      --  controlled type, gid.ads:347-370, so Finalize can call GID.CUDA.Destroy -- and reuse it
      --  for the next image loaded through the same descriptor; creating a context and its pinned
      --  buffers costs more than decoding a small image, measured 3.5 ms against 0.5 ms for 512 x 512.)
    end Reconstruct_On_Device_And_Output;

    ------------------------------------------------------------------
    --  Baseline_DCT_Decoding_Scan (:1179-1220): the MCU loop keeps only the
    --  entropy decoding; Feedback runs 0 .. 50 here and 50 .. 100 during output.
    ------------------------------------------------------------------
    procedure Baseline_DCT_Decoding_Scan is
      mb_x, mb_y : Natural := 0;
    begin
      rst_count := image.JPEG_stuff.restart_interval;
      next_rst := 0;
      macro_blocks_loop :
      loop
        for c in Component loop
          if image.JPEG_stuff.compo_set (c) then
            for sby in 1 .. info_A (c).ups.samples_ver loop
              for sbx in 1 .. info_A (c).ups.samples_hor loop
                Decode_8x8_Block
                  (c,
                   bx => mb_x * info_A (c).ups.samples_hor + sbx - 1,
                   by => mb_y * info_A (c).ups.samples_ver + sby - 1);
              end loop;
            end loop;
          end if;
        end loop;
        mb_x := mb_x + 1;
        if mb_x >= mcu_count_image_h then
          mb_x := 0;
          mb_y := mb_y + 1;
          Feedback ((50 * mb_y) / mcu_count_image_v);
          exit macro_blocks_loop when mb_y >= mcu_count_image_v;
        end if;
        Check_Restart (do_reset_predictors => True);
      end loop macro_blocks_loop;
      Reconstruct_On_Device_And_Output (feedback_base => 50, feedback_span => 50);
    end Baseline_DCT_Decoding_Scan;

    ------------------------------------------------------------------
    --  OPTIONAL: Huffman decoding on the device (GID.CUDA.Scan = gidcuda_job_scan).
    --  Called first by Baseline_DCT_Decoding_Scan; when it returns False, or when the
    --  later Wait raises GID.CUDA.Device_Declined, the MCU loop above runs as before
    --  (the stream is rewound to scan_start; GID.Buffering needs a Set_Index for that,
    --  or the scan's bytes are kept in scan_bytes and re-read from there).
    --    * the scan's bytes are read up to the marker that ends it (Get_Byte; FF 00 is
    --      data, FF D0 .. D7 an RSTn, any other FF xx the end);
    --    * with restart intervals: interval k runs from the byte after RSTk-1 to the FF
    --      of RSTk; the bytes go over as they are;
    --    * without: the stuffing is removed (FF 00 -> FF) and the device's
    --      self-synchronising decoder takes the whole scan.
    ------------------------------------------------------------------
    function Baseline_Scan_On_Device return Boolean is
      scan_bytes : Byte_Vector;          --  (an unbounded array of U8 filled below)
      first, last : Offset_Vector;       --  restart intervals, offsets into scan_bytes
      desc : GID.CUDA.Scan_Desc;
      b : U8;
      ri : constant Natural := image.JPEG_stuff.restart_interval;
    begin
      if ri > 0 then
        first.Append (0);
      end if;
      Read_Scan :
      loop
        Get_Byte (image.buffer, b);
        if b = 16#FF# then
          Get_Byte (image.buffer, b);
          if b = 0 then
            scan_bytes.Append (16#FF#);
            if ri > 0 then
              scan_bytes.Append (0);     --  intervals travel with their stuffing
            end if;
          elsif b in 16#D0# .. 16#D7# and then ri > 0 then
            last.Append (scan_bytes.Length);
            scan_bytes.Append (16#FF#);
            scan_bytes.Append (b);
            first.Append (scan_bytes.Length);
          else
            memo_marker := b;            --  the segment loop carries on from this marker (:2043-2118)
            exit Read_Scan;
          end if;
        else
          scan_bytes.Append (b);
        end if;
      end loop Read_Scan;
      if ri > 0 then
        last.Append (scan_bytes.Length);
        if Natural (first.Length) /= (mcu_count_image_h * mcu_count_image_v + ri - 1) / ri
          or else Natural (first.Length) < 64
        then
          return False;                  --  not what the geometry needs, or too few streams for a GPU
        end if;
      end if;
      desc.restart_interval := Interfaces.Integer_32 (ri);
      desc.nintervals := Interfaces.Integer_32 (first.Length);
      for c in Component loop
        if image.JPEG_stuff.compo_set (c) then
          --  counts and symbols of the tables the scan header selected (Read_DHT keeps them
          --  beside the 65 536-entry VLC tables it builds, :320-391)
          desc.dc (Component'Pos (c)) := Raw_Table (DC, info_B (c).ht_idx_DC);
          desc.ac (Component'Pos (c)) := Raw_Table (AC, info_B (c).ht_idx_AC);
        end if;
      end loop;
      GID.CUDA.Scan
        (cuda_job, desc, scan_bytes.Address, Natural (scan_bytes.Length),
         To_Offset_Array (first), To_Offset_Array (last));
      return True;
    end Baseline_Scan_On_Device;

    ------------------------------------------------------------------
    --  Progressive files: every image_array (x, y, c_idx) access in
    --  Progressive_DCT_Decoding_Scan (:1263-1690) becomes an access to the int16
    --  plane cell  Coefficient_Address (c, x / 8, y / 8, 8 * (y mod 8) + x mod 8)
    --  (same values: the scans accumulate un-dequantised coefficients at
    --  pixel-positioned natural order, :580-588), and
    ------------------------------------------------------------------
    procedure Finalize_Progressive_DCT_Decoding is
    begin
      Reconstruct_On_Device_And_Output (feedback_base => 50, feedback_span => 50);
    end Finalize_Progressive_DCT_Decoding;
