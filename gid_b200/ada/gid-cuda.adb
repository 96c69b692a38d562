--  Body of GID.CUDA: status checking and type conversions only.
--  (Reviewed, not compiled: no Ada toolchain in this repository's image.)

package body GID.CUDA is

  use type C.int;

  --  Status codes of include/gidcuda.h
  ERR_BAD_ARGUMENT  : constant := -1;
  ERR_UNSUPPORTED   : constant := -2;
  ERR_DATA          : constant := -7;
  ERR_NO_DEVICE     : constant := -3;
  ERR_CUDA          : constant := -4;
  ERR_OUT_OF_MEMORY : constant := -5;
  ERR_STATE         : constant := -6;

  procedure Check (status : C.int; ctx : Context) is
    function Message return String is
      p : constant C.Strings.chars_ptr := gidcuda_last_error (ctx);
      use type C.Strings.chars_ptr;
    begin
      if p = C.Strings.Null_Ptr then
        return "";
      end if;
      return C.Strings.Value (p);
    end Message;
  begin
    case status is
      when 0 =>
        null;
      when ERR_UNSUPPORTED =>
        raise unsupported_image_subformat with "JPEG (CUDA): " & Message;
      when ERR_BAD_ARGUMENT =>
        raise error_in_image_data with "JPEG (CUDA): " & Message;
      when ERR_DATA =>
        --  gidcuda_job_scan: the device left the scan to the serial loop of GID.Decoding_JPG
        raise Device_Declined with "JPEG (CUDA): " & Message;
      when others =>
        --  ERR_NO_DEVICE, ERR_CUDA, ERR_OUT_OF_MEMORY, ERR_STATE:
        --  there is deliberately no CPU fallback.
        raise Device_Error with "JPEG (CUDA):" & C.int'Image (status) & " " & Message;
    end case;
  end Check;

  function Device_Count return Natural is
  begin
    return Natural (gidcuda_device_count);
  end Device_Count;

  procedure Create (device : Natural; ctx : out Context) is
    c_ctx : aliased Context := Null_Context;
  begin
    Check (gidcuda_create (C.int (device), c_ctx'Access), Null_Context);
    ctx := c_ctx;
  end Create;

  procedure Destroy (ctx : in out Context) is
  begin
    if ctx /= Null_Context then
      Check (gidcuda_destroy (ctx), ctx);
      ctx := Null_Context;
    end if;
  end Destroy;

  procedure Job_Begin (ctx : Context; desc : Frame_Desc; j : out Job) is
    c_desc : aliased constant Frame_Desc := desc;
    c_job  : aliased Job := Null_Job;
  begin
    Check (gidcuda_job_begin (ctx, c_desc'Access, c_job'Access), ctx);
    j := c_job;
  end Job_Begin;

  function Plane
    (j : Job; comp : Natural; blocks_x, blocks_y : out Natural) return System.Address
  is
    bx, by : aliased Interfaces.Integer_32 := 0;
    use type System.Address;
    a : constant System.Address :=
      gidcuda_job_plane (j, C.int (comp), bx'Access, by'Access);
  begin
    if a = System.Null_Address then
      raise Device_Error with "JPEG (CUDA): no plane for component" & comp'Image;
    end if;
    blocks_x := Natural (bx);
    blocks_y := Natural (by);
    return a;
  end Plane;

  procedure Hint_Rows (j : Job; comp : Natural; rows_used : Positive) is
  begin
    Check (gidcuda_job_hint_rows (j, C.int (comp), C.int (rows_used)), Null_Context);
  end Hint_Rows;

  function Fast_Coef_Limit (desc : Frame_Desc; comp : Natural) return Natural is
    d : aliased constant Frame_Desc := desc;
  begin
    return Natural (gidcuda_fast_coef_limit (d'Access, C.int (comp)));
  end Fast_Coef_Limit;

  procedure Wide_Coefficients (j : Job) is
  begin
    Check (gidcuda_job_wide_coefficients (j), Null_Context);
  end Wide_Coefficients;

  procedure Records
    (j : Job; comp : Natural;
     first, records : out System.Address; capacity, nblocks : out Natural)
  is
    f, r : aliased System.Address := System.Null_Address;
    cap  : aliased C.size_t := 0;
    nb   : aliased Interfaces.Integer_32 := 0;
  begin
    Check (gidcuda_job_records (j, C.int (comp), f'Access, r'Access, cap'Access, nb'Access),
           Null_Context);
    first    := f;
    records  := r;
    capacity := Natural (cap);
    nblocks  := Natural (nb);
  end Records;

  procedure Records_Commit (j : Job; comp : Natural; count : Natural) is
  begin
    Check (gidcuda_job_records_commit (j, C.int (comp), C.size_t (count)), Null_Context);
  end Records_Commit;

  procedure Records_Commit_Runs (j : Job; comp : Natural; runs : Run_Array) is
    b, n : Size_Array (0 .. runs'Length - 1);
  begin
    for i in runs'Range loop
      b (i - runs'First) := C.size_t (runs (i).first_record);
      n (i - runs'First) := C.size_t (runs (i).count);
    end loop;
    Check (gidcuda_job_records_commit_runs
             (j, C.int (comp), Interfaces.Integer_32 (runs'Length), b (0)'Access, n (0)'Access),
           Null_Context);
  end Records_Commit_Runs;

  procedure Scan
    (j : Job; desc : Scan_Desc; data : System.Address; length : Natural;
     interval_begin, interval_end : Offset_Array)
  is
    d : aliased constant Scan_Desc := desc;
  begin
    if interval_begin'Length = 0 then
      Check (gidcuda_job_scan (j, d'Access, data, C.size_t (length), null, null), Null_Context);
    else
      Check (gidcuda_job_scan
               (j, d'Access, data, C.size_t (length),
                interval_begin (interval_begin'First)'Access, interval_end (interval_end'First)'Access),
             Null_Context);
    end if;
  end Scan;

  procedure Output (j : Job; pixels : System.Address; capacity : Natural) is
  begin
    Check (gidcuda_job_output (j, pixels, C.size_t (capacity)), Null_Context);
  end Output;

  function Host_Alloc (bytes : Natural) return System.Address is
    p : aliased System.Address := System.Null_Address;
  begin
    Check (gidcuda_host_alloc (C.size_t (bytes), p'Access), Null_Context);
    return p;
  end Host_Alloc;

  procedure Host_Free (p : System.Address) is
  begin
    Check (gidcuda_host_free (p), Null_Context);
  end Host_Free;

  function Is_Pinned (p : System.Address) return Boolean is
  begin
    return gidcuda_host_is_pinned (p) /= 0;
  end Is_Pinned;

  procedure Plane_Geometry (desc : Frame_Desc; comp : Natural; blocks_x, blocks_y : out Natural) is
    d : aliased constant Frame_Desc := desc;
    bx, by : aliased Interfaces.Integer_32 := 0;
  begin
    Check (gidcuda_plane_geometry (d'Access, C.int (comp), bx'Access, by'Access), Null_Context);
    blocks_x := Natural (bx);
    blocks_y := Natural (by);
  end Plane_Geometry;

  procedure Output_Geometry (desc : Frame_Desc; stride, bytes : out Natural) is
    d : aliased constant Frame_Desc := desc;
    s, b : aliased C.size_t := 0;
  begin
    Check (gidcuda_output_geometry (d'Access, s'Access, b'Access), Null_Context);
    stride := Natural (s);
    bytes  := Natural (b);
  end Output_Geometry;

  procedure Submit (j : Job) is
  begin
    Check (gidcuda_job_submit (j), Null_Context);
  end Submit;

  procedure Wait (j : Job; pixels : out System.Address; stride : out Natural) is
    p : aliased System.Address := System.Null_Address;
    s : aliased C.size_t := 0;
  begin
    Check (gidcuda_job_wait (j, p'Access, s'Access), Null_Context);
    pixels := p;
    stride := Natural (s);
  end Wait;

  procedure Release (j : in out Job) is
  begin
    if j /= Null_Job then
      Check (gidcuda_job_release (j), Null_Context);
      j := Null_Job;
    end if;
  end Release;

end GID.CUDA;
