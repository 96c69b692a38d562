--  gid_pin.adb -- pins the oracle to the reference itself (TEST INFRASTRUCTURE, not shipped).
--
--  Decodes every file named on the command line with the UNMODIFIED GID
--  (GID.Load_Image_Header + the generic GID.Load_Image_Contents, gid.ads:62-143) into a
--  top-down packed RGB8 buffer -- the client of test/benchmark.adb:67-81 and
--  test/mini.adb:55-69: idx = 3 * (x + W * (H - 1 - y)), one Put_Pixel after the other --
--  and prints one line per file:
--
--      <sha-256 of the buffer> <width> <height> <file name>
--
--  tests/test_reference_pin.py compares these lines with the hashes the oracle
--  (oracle/gidref.c) gives for the same files: tests/golden/index.json and the five
--  KAT files of SURVEY.md 8c (test/img/*.jpg of the GID tree).
--
--  Build (needs GNAT; there is none in the image this repository was written in):
--      oracle/ref_pin/build_pin.sh /path/to/gid          ->  oracle/_ref/gid_pin
--  Run:
--      oracle/_ref/gid_pin tests/golden/*.jpg /path/to/gid/test/img/*.jpg > tests/golden/gid_pin_hashes.txt

with GID;

with Ada.Calendar;
with Ada.Command_Line;
with Ada.Exceptions;
with Ada.Streams.Stream_IO;
with Ada.Text_IO;
with Ada.Unchecked_Deallocation;

with GNAT.SHA256;

with Interfaces;

procedure GID_Pin is

  use Ada.Streams, Ada.Text_IO, Interfaces;

  type p_Bytes is access Stream_Element_Array;
  procedure Dispose is new Ada.Unchecked_Deallocation (Stream_Element_Array, p_Bytes);

  procedure Process (file_name : String) is
    f      : Ada.Streams.Stream_IO.File_Type;
    image  : GID.Image_Descriptor;
    buffer : p_Bytes := null;
  begin
    Ada.Streams.Stream_IO.Open (f, Ada.Streams.Stream_IO.In_File, file_name);
    GID.Load_Image_Header (image, Ada.Streams.Stream_IO.Stream (f).all, try_tga => False);
    declare
      subtype Primary_Color_Range is Unsigned_8;
      image_width  : constant Positive := GID.Pixel_Width (image);
      image_height : constant Positive := GID.Pixel_Height (image);
      idx : Stream_Element_Offset;
      --
      procedure Set_X_Y (x, y : Natural) is
      begin
        --  y = 0 is the BOTTOM row (gid-decoding_jpg.adb:1024): top-down buffer
        idx := Stream_Element_Offset (3 * (x + image_width * (image_height - 1 - y)));
      end Set_X_Y;
      --
      procedure Put_Pixel
        (red, green, blue : Primary_Color_Range;
         alpha            : Primary_Color_Range)
      is
      pragma Warnings (off, alpha);
      begin
        buffer (idx)     := Stream_Element (red);
        buffer (idx + 1) := Stream_Element (green);
        buffer (idx + 2) := Stream_Element (blue);
        idx := idx + 3;
      end Put_Pixel;
      --
      procedure Feedback (percents : Natural) is null;
      --
      procedure Load_Image is
        new GID.Load_Image_Contents
          (Primary_Color_Range, Set_X_Y, Put_Pixel, Feedback, GID.fast);
      --
      next_frame : Ada.Calendar.Day_Duration;
    begin
      buffer := new Stream_Element_Array
        (0 .. Stream_Element_Offset (3 * image_width * image_height) - 1);
      buffer.all := (others => 0);
      Load_Image (image, next_frame);
      Put_Line
        (GNAT.SHA256.Digest (buffer.all) &
         Integer'Image (image_width) & Integer'Image (image_height) & ' ' & file_name);
    end;
    Dispose (buffer);
    Ada.Streams.Stream_IO.Close (f);
  exception
    when e : others =>
      --  a file GID refuses is reported, not fatal: the oracle must refuse it too
      Put_Line
        ("error:" & Ada.Exceptions.Exception_Name (e) & " 0 0 " & file_name);
      if Ada.Streams.Stream_IO.Is_Open (f) then
        Ada.Streams.Stream_IO.Close (f);
      end if;
      Dispose (buffer);
  end Process;

begin
  if Ada.Command_Line.Argument_Count = 0 then
    Put_Line (Current_Error, "gid_pin <image_1> [<image_2> ...]   (GID " & GID.version & ")");
    return;
  end if;
  for i in 1 .. Ada.Command_Line.Argument_Count loop
    Process (Ada.Command_Line.Argument (i));
  end loop;
end GID_Pin;
