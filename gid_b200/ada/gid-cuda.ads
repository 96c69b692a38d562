--  GID.CUDA -- thin binding of GID to the B200 JPEG pixel-reconstruction
--  library (libgidcuda.so, C ABI declared in include/gidcuda.h).
--
--  This package replaces, for JPEG images only, the work that
--  GID.Decoding_JPG did inline per MCU (gid-decoding_jpg.adb):
--     de-quantisation            :771, :785, :1820-1822
--     Inverse_DCT_8x8_Block      :790-907
--     Upsampling_and_Output*     :957-1124
--     Color_Transformation_...   :1018-1054
--  The entropy decoder keeps running on the host and stores its
--  coefficients, NOT de-quantised, into the planes lent by Plane.
--
--  One Context per Ada task (GID is "task safe", doc/gid.txt:20).
--  Every imported function returns an int status; Check turns a
--  failure into an exception, so nothing unwinds through C frames.
--
--  NOTE: written for GNAT, but no Ada compiler exists in the build
--  image of this repository: this unit has been reviewed, not compiled.
--  The tested stand-in is the C++ host mirror (gid_b200/host).

with Interfaces.C;
with Interfaces.C.Strings;
with System;

private package GID.CUDA is

  package C renames Interfaces.C;

  Device_Error : exception;
  --  GIDCUDA_ERR_DATA-like classes map to GID.error_in_image_data and
  --  GIDCUDA_ERR_UNSUPPORTED to GID.unsupported_image_subformat (see Check).

  type Context is private;
  type Job is private;
  Null_Context : constant Context;
  Null_Job     : constant Job;

  type Byte_Array is array (Natural range <>) of aliased Interfaces.Unsigned_8;
  pragma Convention (C, Byte_Array);
  type Sampling_Array is array (0 .. 2) of aliased Interfaces.Integer_32;
  pragma Convention (C, Sampling_Array);
  type Natural_Order_Table is array (0 .. 63) of aliased Interfaces.Unsigned_16;
  pragma Convention (C, Natural_Order_Table);
  type Table_Array is array (0 .. 2) of aliased Natural_Order_Table;
  pragma Convention (C, Table_Array);

  Out_RGB8           : constant := 0;
  Out_BGR8_Bottom_Up : constant := 1;         --  test/to_bmp.adb's buffer
  Out_RGBA8          : constant := 2;
  Flag_Packed_Rows   : constant := 16#100#;
  --  GID.Display_Orientation applied while the pixels are written (Set_X_Y of test/to_bmp.adb:104-122)
  Flag_Rotate_90     : constant := 16#200#;
  Flag_Rotate_180    : constant := 16#400#;
  Flag_Rotate_270    : constant := 16#600#;
  Flag_Wide_Coef     : constant := 16#800#;   --  see Fast_Coef_Limit

  --  = struct gidcuda_frame_desc
  type Frame_Desc is record
    width, height : Interfaces.Integer_32;
    ncomp         : Interfaces.Integer_32;   --  1 = Y, 3 = Y, Cb, Cr
    samp_h        : Sampling_Array;          --  samples_hor, :246
    samp_v        : Sampling_Array;          --  samples_ver, :247
    qt            : Table_Array;             --  natural order, :1771-1775
    flags         : Interfaces.Unsigned_32;
    out_stride    : Interfaces.Integer_32;   --  0 = library default
  end record;
  pragma Convention (C, Frame_Desc);

  function Device_Count return Natural;

  procedure Create  (device : Natural; ctx : out Context);
  procedure Destroy (ctx : in out Context);

  --  Begin a frame: pinned, zeroed coefficient planes are reserved.
  procedure Job_Begin (ctx : Context; desc : Frame_Desc; j : out Job);

  --  Address of the int16 plane of component comp (0 = Y, 1 = Cb, 2 = Cr):
  --  block-major (block_row, block_col, 0 .. 63), natural order.
  function Plane
    (j : Job; comp : Natural; blocks_x, blocks_y : out Natural) return System.Address;

  --  Promise that rows >= rows_used of every block of the plane are zero (1 .. 8).
  procedure Hint_Rows (j : Job; comp : Natural; rows_used : Positive);

  --  Range of the fast kernel instances (include/gidcuda.h, GIDCUDA_FLAG_WIDE_COEF): the largest
  --  coefficient magnitude of component comp for which the decisions the kernels take on raw
  --  coefficients are those Inverse_DCT_8x8_Block takes on de-quantised ones (:803-812, :858-862).
  --  An entropy decoder that has stored a larger one -- Get_VLC accepts up to 15 bits, :727-733; no
  --  JPEG of 8-bit samples holds one -- calls Wide_Coefficients before Submit: the frame then runs the
  --  instances that reproduce the reference for every 16-bit coefficient.
  function  Fast_Coef_Limit (desc : Frame_Desc; comp : Natural) return Natural;
  procedure Wide_Coefficients (j : Job);

  --  Sparse alternative to Plane (include/gidcuda.h, "sparse coefficient input"):
  --  one 32-bit record per NON-ZERO coefficient,
  --     record = Shift_Left (Unsigned_32 (Unsigned_16'Mod (value)), 16) or natural_index,
  --  and first (seq) = index of the first record of block seq, blocks counted in the
  --  order Baseline_DCT_Decoding_Scan visits them (:1190-1206); first (nblocks) = count.
  --  capacity = 32 records per block; a decoder that runs out does not commit and
  --  fills the dense plane instead.
  procedure Records
    (j : Job; comp : Natural;
     first, records : out System.Address; capacity, nblocks : out Natural);
  procedure Records_Commit (j : Job; comp : Natural; count : Natural);
  --  The same for records produced by several tasks (one run per task; runs ascending, disjoint)
  type Run is record
    first_record, count : Natural;
  end record;
  type Run_Array is array (Positive range <>) of Run;
  procedure Records_Commit_Runs (j : Job; comp : Natural; runs : Run_Array);

  --  Entropy decoding on the device (include/gidcuda.h, gidcuda_job_scan): instead of running
  --  Decode_8x8_Block (:758-788) over a baseline scan, hand over its compressed bytes.
  --    restart_interval > 0 : data = the scan's bytes as they are; interval_begin / interval_end =
  --        where each restart interval lies (from the byte after RSTn-1 to the FF of RSTn);
  --    restart_interval = 0 : data = the scan's bytes with the stuffing removed (FF 00 -> FF);
  --        the two arrays are empty.
  --  Wait then raises Device_Declined when the device met something it leaves to the serial
  --  loop (a damaged or unusual stream): decode the scan the old way and Submit the job again.
  Device_Declined : exception;
  type Huffman_Table is record
    counts  : Byte_Array (0 .. 15);    --  Read_DHT :320-391: codes of length 1 .. 16
    symbols : Byte_Array (0 .. 255);   --  in code order
  end record;
  pragma Convention (C, Huffman_Table);
  type Huffman_Table_Array is array (0 .. 2) of aliased Huffman_Table;
  pragma Convention (C, Huffman_Table_Array);
  type Scan_Desc is record
    restart_interval : Interfaces.Integer_32;
    nintervals       : Interfaces.Integer_32;
    dc, ac           : Huffman_Table_Array;   --  per component of the frame
  end record;
  pragma Convention (C, Scan_Desc);
  type Offset_Array is array (Natural range <>) of aliased Interfaces.Unsigned_32;
  pragma Convention (C, Offset_Array);
  procedure Scan
    (j : Job; desc : Scan_Desc; data : System.Address; length : Natural;
     interval_begin, interval_end : Offset_Array);

  --  Let the pixels land in the client's own image buffer (pinned host memory: Host_Alloc):
  --  no copy after the device-to-host transfer.  Between Job_Begin and Submit.
  procedure Output (j : Job; pixels : System.Address; capacity : Natural);
  function  Host_Alloc (bytes : Natural) return System.Address;
  procedure Host_Free (p : System.Address);
  function  Is_Pinned (p : System.Address) return Boolean;

  --  MCU-padded block grid of a component / row stride and size of the pixel buffer
  procedure Plane_Geometry (desc : Frame_Desc; comp : Natural; blocks_x, blocks_y : out Natural);
  procedure Output_Geometry (desc : Frame_Desc; stride, bytes : out Natural);

  procedure Submit (j : Job);   --  asynchronous: H2D, kernel(s), D2H
  procedure Wait   (j : Job; pixels : out System.Address; stride : out Natural);
  procedure Release (j : in out Job);

private

  type Context is new System.Address;
  type Job     is new System.Address;
  Null_Context : constant Context := Context (System.Null_Address);
  Null_Job     : constant Job     := Job (System.Null_Address);

  --  Raw imports (names = include/gidcuda.h)

  function gidcuda_device_count return C.int;
  pragma Import (C, gidcuda_device_count, "gidcuda_device_count");

  function gidcuda_last_error (ctx : Context) return C.Strings.chars_ptr;
  pragma Import (C, gidcuda_last_error, "gidcuda_last_error");

  function gidcuda_create (device : C.int; ctx : access Context) return C.int;
  pragma Import (C, gidcuda_create, "gidcuda_create");

  function gidcuda_destroy (ctx : Context) return C.int;
  pragma Import (C, gidcuda_destroy, "gidcuda_destroy");

  function gidcuda_job_begin
    (ctx : Context; desc : access constant Frame_Desc; j : access Job) return C.int;
  pragma Import (C, gidcuda_job_begin, "gidcuda_job_begin");

  function gidcuda_job_plane
    (j : Job; comp : C.int;
     blocks_x, blocks_y : access Interfaces.Integer_32) return System.Address;
  pragma Import (C, gidcuda_job_plane, "gidcuda_job_plane");

  function gidcuda_job_hint_rows (j : Job; comp, rows_used : C.int) return C.int;
  pragma Import (C, gidcuda_job_hint_rows, "gidcuda_job_hint_rows");

  function gidcuda_fast_coef_limit
    (desc : access constant Frame_Desc; comp : C.int) return Interfaces.Integer_32;
  pragma Import (C, gidcuda_fast_coef_limit, "gidcuda_fast_coef_limit");
  function gidcuda_job_wide_coefficients (j : Job) return C.int;
  pragma Import (C, gidcuda_job_wide_coefficients, "gidcuda_job_wide_coefficients");

  function gidcuda_job_records
    (j : Job; comp : C.int; first, records : access System.Address;
     capacity : access C.size_t; nblocks : access Interfaces.Integer_32) return C.int;
  pragma Import (C, gidcuda_job_records, "gidcuda_job_records");

  function gidcuda_job_records_commit (j : Job; comp : C.int; count : C.size_t) return C.int;
  pragma Import (C, gidcuda_job_records_commit, "gidcuda_job_records_commit");

  --  Records produced by several tasks, each decoding its own group of restart
  --  intervals into its own slice of the lent array (include/gidcuda.h)
  type Size_Array is array (Natural range <>) of aliased C.size_t;
  pragma Convention (C, Size_Array);
  function gidcuda_job_records_commit_runs
    (j : Job; comp : C.int; nruns : Interfaces.Integer_32;
     run_begin, run_count : access constant C.size_t) return C.int;
  pragma Import (C, gidcuda_job_records_commit_runs, "gidcuda_job_records_commit_runs");

  function gidcuda_job_submit (j : Job) return C.int;
  pragma Import (C, gidcuda_job_submit, "gidcuda_job_submit");

  function gidcuda_job_wait
    (j : Job; pixels : access System.Address; stride : access C.size_t) return C.int;
  pragma Import (C, gidcuda_job_wait, "gidcuda_job_wait");

  function gidcuda_job_release (j : Job) return C.int;
  pragma Import (C, gidcuda_job_release, "gidcuda_job_release");

  function gidcuda_job_scan
    (j : Job; scan : access constant Scan_Desc; data : System.Address; len : C.size_t;
     interval_begin, interval_end : access constant Interfaces.Unsigned_32) return C.int;
  pragma Import (C, gidcuda_job_scan, "gidcuda_job_scan");

  --  Progressive frames on the device (include/gidcuda.h, gidcuda_job_pscan*): first scans and DC refinement
  --  scans as compressed bytes, AC refinement decoded by the host from the non-zero history.  Raw imports; the
  --  C++ host mirror (gid_b200/host/jpeg_entropy.hpp, device_progressive_scan) is the worked example of the
  --  call sequence a patched Progressive_DCT_Decoding_Scan (:1243-1747) makes.
  type Pscan_Desc is record
    ncomp          : Interfaces.Integer_32;
    comp           : Sampling_Array;
    ss, se, ah, al : Interfaces.Integer_32;
    dc, ac         : Huffman_Table_Array;
  end record;
  pragma Convention (C, Pscan_Desc);
  function gidcuda_job_pscan_reserve (j : Job; total_bytes : C.size_t) return C.int;
  pragma Import (C, gidcuda_job_pscan_reserve, "gidcuda_job_pscan_reserve");
  function gidcuda_job_pscan
    (j : Job; scan : access constant Pscan_Desc; data : System.Address; len : C.size_t) return C.int;
  pragma Import (C, gidcuda_job_pscan, "gidcuda_job_pscan");
  function gidcuda_job_pscan_masks
    (j : Job; comp : C.int; masks : access System.Address; nblocks : access Interfaces.Integer_32) return C.int;
  pragma Import (C, gidcuda_job_pscan_masks, "gidcuda_job_pscan_masks");
  function gidcuda_job_pscan_refine
    (j : Job; comp : C.int; corr, newpos, newneg : access System.Address) return C.int;
  pragma Import (C, gidcuda_job_pscan_refine, "gidcuda_job_pscan_refine");
  function gidcuda_job_pscan_refine_commit (j : Job; comp, al : C.int) return C.int;
  pragma Import (C, gidcuda_job_pscan_refine_commit, "gidcuda_job_pscan_refine_commit");

  function gidcuda_job_output (j : Job; host_pixels : System.Address; capacity : C.size_t) return C.int;
  pragma Import (C, gidcuda_job_output, "gidcuda_job_output");

  function gidcuda_host_alloc (bytes : C.size_t; ptr : access System.Address) return C.int;
  pragma Import (C, gidcuda_host_alloc, "gidcuda_host_alloc");
  function gidcuda_host_free (ptr : System.Address) return C.int;
  pragma Import (C, gidcuda_host_free, "gidcuda_host_free");
  function gidcuda_host_is_pinned (ptr : System.Address) return C.int;
  pragma Import (C, gidcuda_host_is_pinned, "gidcuda_host_is_pinned");

  function gidcuda_plane_geometry
    (desc : access constant Frame_Desc; comp : C.int;
     blocks_x, blocks_y : access Interfaces.Integer_32) return C.int;
  pragma Import (C, gidcuda_plane_geometry, "gidcuda_plane_geometry");
  function gidcuda_output_geometry
    (desc : access constant Frame_Desc; stride, bytes : access C.size_t) return C.int;
  pragma Import (C, gidcuda_output_geometry, "gidcuda_output_geometry");

  function gidcuda_abi_version return C.int;
  pragma Import (C, gidcuda_abi_version, "gidcuda_abi_version");
  function gidcuda_set_option (name : C.Strings.chars_ptr; value : C.int) return C.int;
  pragma Import (C, gidcuda_set_option, "gidcuda_set_option");
  function gidcuda_counter (name : C.Strings.chars_ptr) return C.long;
  pragma Import (C, gidcuda_counter, "gidcuda_counter");
  function gidcuda_synchronize (ctx : Context) return C.int;
  pragma Import (C, gidcuda_synchronize, "gidcuda_synchronize");

  --  gidcuda_plan_* (planes already in device memory) and gidcuda_batch_* (many frames, all GPUs of
  --  a box) serve callers other than GID.Decoding_JPG and are not bound here.

end GID.CUDA;
