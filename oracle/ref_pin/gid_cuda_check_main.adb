--  Main of the binding check (see gid-cuda_check.ads).
with GID.CUDA_Check;
procedure GID_CUDA_Check_Main is
begin
  GID.CUDA_Check;
end GID_CUDA_Check_Main;
