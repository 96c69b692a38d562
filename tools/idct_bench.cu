// idct_bench.cu -- how fast does the register-resident IDCT issue on a B200 SM,
// as a function of resident warps?  Each thread runs the product's idct_8x8
// (gid_b200/csrc/recon_math.cuh) repeatedly on registers; occupancy is set
// through the dynamic shared-memory size.  Prints blocks/s and the implied
// warp-instructions per clock per SM (instruction count per IDCT taken from
// the SASS: see the constant below).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gid_b200/csrc -o tools/idct_bench tools/idct_bench.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "recon_math.cuh"

using namespace gidk;

#ifndef INSTR_PER_IDCT
#define INSTR_PER_IDCT 1230  // SASS count of one idct_8x8 + 32 I2IP (tools: cuobjdump)
#endif

template <int VARIANT>
__global__ void __launch_bounds__(128) idct_loop(const int *__restrict__ in, unsigned *__restrict__ out, int iters) {
  extern __shared__ unsigned char dummy[];
  int b[64];
#pragma unroll
  for (int i = 0; i < 64; i++) b[i] = in[(threadIdx.x * 64 + i) & 4095];
  unsigned acc = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    idct_8x8(b);
    // fold the result back into small coefficients so that the next round does real work
#pragma unroll
    for (int i = 0; i < 16; i++) acc += pack_sat4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]);
#pragma unroll
    for (int i = 0; i < 64; i++) b[i] = (b[i] - 128) >> (VARIANT == 0 ? 1 : 2);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + dummy[0];
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  int *d_in;
  unsigned *d_out;
  cudaMalloc(&d_in, 4096 * sizeof(int));
  int h[4096];
  for (int i = 0; i < 4096; i++) h[i] = (i * 37 % 201) - 100;
  cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, sizeof(unsigned) * 128 * p.multiProcessorCount * 16);
  cudaFuncSetAttribute(idct_loop<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 200;
  printf("device %s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, clock_khz / 1000);
  for (int ctas = 1; ctas <= 8; ctas++) {
    // dynamic smem so that exactly `ctas` CTAs of 128 threads fit on an SM
    int smem = (227 * 1024) / ctas - 2048;
    if (smem > 200 * 1024) smem = 200 * 1024;
    int maxb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, idct_loop<0>, 128, smem);
    const int blocks = p.multiProcessorCount * maxb;
    idct_loop<0><<<blocks, 128, smem>>>(d_in, d_out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0);
      idct_loop<0><<<blocks, 128, smem>>>(d_in, d_out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double idcts = (double)blocks * 128 * iters;
    const double per_s = idcts / (best * 1e-3);
    const double ipc = per_s / 32.0 * (INSTR_PER_IDCT + 64) / ((double)clock_khz * 1e3) / p.multiProcessorCount;
    printf("CTAs/SM %d (warps/SM %2d): %.3f ms, %.2f G blocks/s, ~%.2f warp-instr/clk/SM, err=%s\n", maxb, maxb * 4, best,
           per_s * 1e-9, ipc, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
