--  One 16 x 16 grey frame, every block DC-only with a de-quantised DC term of -17: the reference gives
--  127 everywhere (SURVEY.md 8c KAT-1: Row_IDCT's shortcut blk (0) * 8 = -136, Col_IDCT's
--  Clip ((-136 + 32) / 64 + 128) with Ada's truncating "/" = 127; an arithmetic shift would give 126).
with GID.CUDA;

with Ada.Text_IO;
with Interfaces;
with System;

procedure GID.CUDA_Check is
  use Ada.Text_IO, Interfaces;
  ctx  : GID.CUDA.Context;
  job  : GID.CUDA.Job;
  desc : GID.CUDA.Frame_Desc;
  bx, by, stride : Natural;
  base, pixels : System.Address;
begin
  if GID.CUDA.Device_Count = 0 then
    Put_Line ("GID.CUDA_Check: binding linked; no CUDA device on this box, nothing run");
    return;
  end if;
  desc.width      := 16;
  desc.height     := 16;
  desc.ncomp      := 1;
  desc.samp_h     := (others => 1);
  desc.samp_v     := (others => 1);
  desc.qt         := (others => (others => 1));
  desc.flags      := GID.CUDA.Out_RGB8 + GID.CUDA.Flag_Packed_Rows;
  desc.out_stride := 0;
  GID.CUDA.Create (device => 0, ctx => ctx);
  GID.CUDA.Job_Begin (ctx, desc, job);
  base := GID.CUDA.Plane (job, 0, bx, by);
  declare
    type Plane_Array is array (0 .. 64 * bx * by - 1) of Integer_16;
    plane : Plane_Array;
    for plane'Address use base;
    pragma Import (Ada, plane);
  begin
    for b in 0 .. bx * by - 1 loop
      plane (64 * b) := -17;
    end loop;
  end;
  GID.CUDA.Hint_Rows (job, 0, 1);
  GID.CUDA.Submit (job);
  GID.CUDA.Wait (job, pixels, stride);
  declare
    type Pixel_Array is array (0 .. stride * 16 - 1) of Unsigned_8;
    rgb : Pixel_Array;
    for rgb'Address use pixels;
    pragma Import (Ada, rgb);
    bad : Natural := 0;
  begin
    for row in 0 .. 15 loop
      for i in 0 .. 3 * 16 - 1 loop
        if rgb (row * stride + i) /= 127 then
          bad := bad + 1;
        end if;
      end loop;
    end loop;
    if bad = 0 then
      Put_Line ("GID.CUDA_Check: ok (16 x 16 grey frame, DC -17 -> 127 everywhere, as GID paints it)");
    else
      Put_Line ("GID.CUDA_Check: FAILED," & Natural'Image (bad) & " bytes differ from 127");
    end if;
  end;
  GID.CUDA.Release (job);
  GID.CUDA.Destroy (ctx);
end GID.CUDA_Check;
