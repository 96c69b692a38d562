--  GID.CUDA_Check -- compile-and-link check of the binding GID.CUDA (gid_b200/ada) against libgidcuda.so,
--  for the day a GNAT box exists (SURVEY.md 8f rank 4).  TEST INFRASTRUCTURE.  A public child of GID, so
--  that its body may "with" the private child GID.CUDA.
procedure GID.CUDA_Check;
