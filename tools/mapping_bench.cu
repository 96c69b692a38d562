// mapping_bench.cu -- lane-per-block against lanes-per-row/column for the exact 8x8 IDCT, measured.
//
// The fused kernels keep a whole block in the registers of ONE lane (64 values, unrolled row and column
// passes: ~1 200 instructions = 19 KB of straight-line code per IDCT, 117-128 registers, 16 warps/SM).
// The alternative the north star sketches: 8 lanes per block, a lane runs ONE row transform, the 64
// row outputs cross shared memory, the lane runs ONE column transform -- ~170 instructions of code, ~40
// registers, up to 48 warps/SM.  Same arithmetic (recon_math.cuh), same work per block: dense blocks,
// de-quantise + row pass + column pass + clip/pack + 64 sample bytes to shared memory, no pixel stage.
//   A  lane = block:   raw rows from shared memory (8 x LDS.128), idct in registers, 8 x STS.64
//   B  lane = row / column: LDS.128 (its row), row pass, 8 x STS.32 (transposed), syncwarp, 2 x LDS.128
//      (its column), column pass, 1 x STS.64 (column-packed bytes: the byte transpose a pixel stage would
//      need on top is NOT included -- this is B's best case)
// Both loop `iters` times over blocks resident in shared memory; reported: blocks/s per SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gid_b200/csrc -o tools/mapping_bench tools/mapping_bench.cu
#include <cuda_runtime.h>

#include <cstdio>

#include "recon_body.cuh"

using namespace gidk;

constexpr int kWarpsPerCta = 4;

// A: one lane owns one block.  Shared memory per warp: 32 blocks x 128 B raw + 32 x 64 B samples.
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) lane_per_block(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                                    int iters) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4 *raw = reinterpret_cast<uint4 *>(smem) + warp * 32 * 8;                 // [block][8 rows] x 16 B
  u32x2 *samples = reinterpret_cast<u32x2 *>(smem + kWarpsPerCta * 4096) + warp * 8 * 32;  // [row][lane]
  for (int i = lane; i < 32 * 8; i += 32) raw[i] = reinterpret_cast<const uint4 *>(in)[(i * 7 + warp) & 1023];
  uint32_t qw[32];
#pragma unroll
  for (int i = 0; i < 32; i++) qw[i] = 0x03000002u + (uint32_t)i;              // 8-bit tables, two entries per word
  __syncwarp();
  uint32_t acc = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    u32x4 r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = raw[lane * 8 + (j ^ (lane & 7))];       // (rotated: conflict-free like the swizzle)
    int b[64];
    dequant_idct<true>(r, qw, b);
    store_block_packed(b, samples + lane, 32);
    acc += samples[lane].x;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// B: eight lanes own one block; a warp works on 4 blocks at a time, 8 groups of 4 per "32 blocks".
__global__ void __launch_bounds__(kWarpsPerCta * 32) lanes_per_row(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                                   int iters) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = lane >> 3, k = lane & 7;                                     // block of the group, row then column
  uint4 *raw = reinterpret_cast<uint4 *>(smem) + warp * 32 * 8;
  // row outputs of one group: [block][column][row + pad] int32 (33 words per column set apart: conflict-free both ways)
  int *mid = reinterpret_cast<int *>(smem + kWarpsPerCta * 4096) + warp * (4 * 8 * 9);
  u32x2 *samples = reinterpret_cast<u32x2 *>(smem + kWarpsPerCta * (4096 + 4 * 8 * 9 * 4)) + warp * 32 * 8;  // [block][column]
  for (int i = lane; i < 32 * 8; i += 32) raw[i] = reinterpret_cast<const uint4 *>(in)[(i * 7 + warp) & 1023];
  uint32_t qrow[4];
#pragma unroll
  for (int i = 0; i < 4; i++) qrow[i] = 0x03000002u + (uint32_t)(4 * k + i);  // the row's table words (a kernel would LDS them)
  __syncwarp();
  uint32_t acc = 0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll 1
    for (int g = 0; g < 8; g++) {                                              // 8 groups of 4 blocks = 32 blocks
      const u32x4 w = raw[(g * 4 + blk) * 8 + k];
      int c0, c1, c2, c3, c4, c5, c6, c7;
      dequant_row<true>(w, qrow, c0, c1, c2, c3, c4, c5, c6, c7);
      row_idct_rnd(128, c0, c1, c2, c3, c4, c5, c6, c7);
      int *m = mid + blk * 72 + k;                                             // [block][column c][row k]
      m[0] = c0; m[9] = c1; m[18] = c2; m[27] = c3; m[36] = c4; m[45] = c5; m[54] = c6; m[63] = c7;
      __syncwarp();
      const int *col = mid + blk * 72 + k * 9;                                 // column k, rows 0 .. 7
      int b0 = col[0], b1 = col[1], b2 = col[2], b3 = col[3], b4 = col[4], b5 = col[5], b6 = col[6], b7 = col[7];
      col_idct(b0, b1, b2, b3, b4, b5, b6, b7);
      u32x2 v;
      v.x = pack_sat4(b0, b1, b2, b3);
      v.y = pack_sat4(b4, b5, b6, b7);
      samples[(g * 4 + blk) * 8 + k] = v;
      __syncwarp();
    }
    acc += samples[lane].x;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
void run(const char *name, K kernel, int smem_per_cta, const uint32_t *d_in, uint32_t *d_out, int sms, double ghz) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncAttributes at;
  cudaFuncGetAttributes(&at, kernel);
  printf("%s: %d registers/thread, %d B of code\n", name, at.numRegs, 0);
  for (int ctas = 1; ctas <= 12; ctas++) {
    int smem = (227 * 1024) / ctas - 2048;                                     // exactly `ctas` CTAs per SM fit
    if (smem > 200 * 1024) smem = 200 * 1024;
    if (smem < smem_per_cta) break;
    int maxb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kernel, kWarpsPerCta * 32, smem);
    if (maxb != ctas) continue;                                                // registers cap it earlier
    const int iters = 400;
    kernel<<<sms * maxb, kWarpsPerCta * 32, smem>>>(d_in, d_out, 4);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0);
      kernel<<<sms * maxb, kWarpsPerCta * 32, smem>>>(d_in, d_out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double blocks = (double)sms * maxb * kWarpsPerCta * 32.0 * iters;    // 32 blocks per warp per iteration, both mappings
    printf("  %2d warps/SM: %8.3f ms  %7.2f G blocks/s  = %.3f blocks / clk / SM   (%s)\n", maxb * kWarpsPerCta, best,
           blocks / (best * 1e-3) * 1e-9, blocks / (best * 1e-3) / (ghz * 1e9) / sms, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  uint32_t *d_in, *d_out;
  cudaMalloc(&d_in, 1024 * 16);
  uint32_t h[4096];
  for (int i = 0; i < 4096; i++) {  // small dense coefficients of both signs, two per word
    const int a = (i * 37 % 61) - 30, b = (i * 53 % 41) - 20;
    h[i] = ((uint32_t)(uint16_t)(int16_t)a) | ((uint32_t)(uint16_t)(int16_t)b << 16);
  }
  cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, sizeof(uint32_t) * 128 * 16 * p.multiProcessorCount);
  printf("device %s, %d SMs, %.3f GHz; dense blocks, de-quantise + IDCT + clip/pack per block\n", p.name,
         p.multiProcessorCount, khz / 1e6);
  run("A  lane = block (64 values in registers, unrolled)", lane_per_block, kWarpsPerCta * (4096 + 2048), d_in, d_out,
      p.multiProcessorCount, khz / 1e6);
  run("B  lane = row, then column (8 lanes per block, row outputs through shared memory)", lanes_per_row,
      kWarpsPerCta * (4096 + 4 * 8 * 9 * 4 + 2048), d_in, d_out, p.multiProcessorCount, khz / 1e6);
  return 0;
}
