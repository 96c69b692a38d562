// icache_bench.cu -- does the SM issue straight-line integer code as fast as a loop that fits the
// instruction caches?  The fused reconstruction kernels are ~40-46 KB of mostly straight-line SASS (the
// IDCT is unrolled over register-resident rows and columns) run by 16 warps per SM that drift apart; if
// instruction fetch bounds them, extra instructions cost time even on an idle pipe (measured: they do).
// Each thread runs `iters` passes over a body of REP x 32 independent integer instructions (IMAD on the FMA
// pipe and SHF/LOP3/IADD3 on the ALU pipe, about the kernels' 40:60 mix); body size = REP x 32 x 16 bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/icache_bench tools/icache_bench.cu
#include <cuda_runtime.h>

#include <cstdio>

template <int REP>
__global__ void __launch_bounds__(128, 4) body_kernel(const int *__restrict__ in, int *__restrict__ out, int iters) {
  int a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = in[(threadIdx.x + i) & 255];
  const int k = in[0] | 1;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < REP; r++) {
      // 32 instructions: 12 IMAD (FMA pipe), 20 on the ALU pipe; 16 independent chains
#pragma unroll
      for (int i = 0; i < 12; i++) a[i] = a[i] * k + (r + i);                 // IMAD
#pragma unroll
      for (int i = 0; i < 8; i++) a[i + 8] = (a[i + 8] >> 1) ^ a[(i + 3) & 15];  // SHF feeding LOP3 (fused: LOP3 only) ...
#pragma unroll
      for (int i = 0; i < 6; i++) a[i] = __funnelshift_r(a[i], a[i + 1], 3);  // SHF
#pragma unroll
      for (int i = 0; i < 6; i++) a[i + 6] = min(a[i + 6], a[i] + r);         // IADD + VIMNMX
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int REP>
void run(const int *d_in, int *d_out, int sms, double clock_ghz) {
  cudaFuncAttributes at;
  cudaFuncGetAttributes(&at, body_kernel<REP>);
  const int iters = 4096 / REP > 4 ? 4096 / REP : 4;
  for (int ctas = 1; ctas <= 4; ctas *= 2) {
    body_kernel<REP><<<sms * ctas, 128>>>(d_in, d_out, 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    body_kernel<REP><<<sms * ctas, 128>>>(d_in, d_out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("body %6.1f KB of SASS  %2d warps/SM  %8.3f ms  for %d passes", REP * 38 * 16 / 1024.0, ctas * 4, ms, iters);
    // 38 SASS instructions per unit (cuobjdump: 208 for REP = 4 ... 2488 for REP = 64, loop and prologue included)
    const double ops = (double)REP * 38.0 * iters * (double)(ctas * 4);
    printf("   -> %.2f warp-instructions / clk / SM\n", ops / (ms * 1e-3 * clock_ghz * 1e9));
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  int *d_in, *d_out;
  cudaMalloc(&d_in, 256 * sizeof(int));
  int h[256];
  for (int i = 0; i < 256; i++) h[i] = i * 2654435 + 17;
  cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, sizeof(int) * 128 * 4 * p.multiProcessorCount);
  printf("device %s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<4>(d_in, d_out, p.multiProcessorCount, ghz);
  run<16>(d_in, d_out, p.multiProcessorCount, ghz);
  run<32>(d_in, d_out, p.multiProcessorCount, ghz);
  run<64>(d_in, d_out, p.multiProcessorCount, ghz);
  run<96>(d_in, d_out, p.multiProcessorCount, ghz);
  run<128>(d_in, d_out, p.multiProcessorCount, ghz);
  run<192>(d_in, d_out, p.multiProcessorCount, ghz);
  run<256>(d_in, d_out, p.multiProcessorCount, ghz);
  return 0;
}
