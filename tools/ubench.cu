// ubench.cu -- integer-pipe microbenchmarks for the B200 reconstruction kernel.
//
// The reconstruction path is integer-issue bound, so the design depends on how
// many IMAD / IADD3 / SHF / LOP3 / IDP / VIMNMX / PRMT / I2IP ... instructions an
// SM retires per clock, alone and mixed.  Each kernel runs ILP independent
// dependency chains per thread; result = warp-instructions per clock per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
//   tools/ubench            (prints one line per case)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define ILP 8
#define REP 8

#define CHECK(x)                                                                        \
  do {                                                                                  \
    cudaError_t e = (x);                                                                \
    if (e != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);    \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

enum Op {
  OP_IMAD, OP_IADD, OP_IADD3, OP_SHF, OP_LOP3, OP_LEAHI, OP_TDIV, OP_DP2A, OP_VIMNMX, OP_VIMNMX_RELU,
  OP_PRMT, OP_I2IP, OP_VIADDMNMX_S16X2, OP_IMAD_IADD, OP_IMAD_SHF, OP_IMAD_LOP_SHF, OP_FFMA,
  OP_FFMA_IMAD, OP_FFMA_IADD, OP_MULHI, OP_SEL, OP_I2F_F2I, OP_SHFL, OP_LDS128, OP_IABS, OP_DFMA,
  OP_IMAD_IADD_2_1, OP_VIADD3_RELU, OP_COUNT
};

static const char *op_name[OP_COUNT] = {
  "IMAD (mad.lo.s32)", "IADD (add.s32)", "IADD3 (3-input add)", "SHF (shr.s32)", "LOP3 (a&b^c)",
  "LEA.HI pattern (x + (s>>>24))", "trunc div x/256 (3 ops)", "IDP.2A (dp2a.lo.s32.u32)",
  "VIMNMX (min.s32)", "VIMNMX.RELU (min.relu)", "PRMT", "I2IP (cvt.pack.sat.u8.s32)",
  "VIADDMNMX.S16x2.RELU (DPX)", "mix IMAD+IADD 1:1", "mix IMAD+SHF 1:1", "mix IMAD+LOP3+SHF 1:1:1",
  "FFMA", "mix FFMA+IMAD 1:1", "mix FFMA+IADD 1:1", "IMAD.HI (mul.hi.s32)", "SEL (selp)",
  "I2F+F2I pair", "SHFL.BFLY", "LDS.128", "IABS+ISUB pair", "DFMA", "mix IMAD+IADD 2:1",
  "VIADDMNMX s32 relu (viaddmin_s32_relu)"};

// instructions issued per inner statement (for the normalisation)
static const int op_instr[OP_COUNT] = {1, 1, 1, 1, 1, 1, 3, 1, 1, 1, 1, 1, 1, 2, 2, 3, 1, 2, 2, 1, 1, 2, 1, 1, 2, 1, 3, 1};

template <int OP>
__global__ void __launch_bounds__(256) bench_kernel(int *out, int iters, int a0, int b0, float f0) {
  __shared__ int4 sm[256];
  int x[ILP];
  float f[ILP];
  double d[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) {
    x[j] = threadIdx.x * 7 + j + a0;
    f[j] = (float)(threadIdx.x + j) * f0;
    d[j] = (double)f[j];
  }
  sm[threadIdx.x] = make_int4(x[0], x[1], x[2], x[3]);
  __syncthreads();
  const int b = b0, c = a0 | 5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < REP; r++) {
#pragma unroll
      for (int j = 0; j < ILP; j++) {
        if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
        const int y = x[(j + 1) % ILP], z = x[(j + 2) % ILP];
        if (OP == OP_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));
        if (OP == OP_IADD3) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(x[j]) : "r"(y), "r"(z));
        if (OP == OP_SHF) asm volatile("shr.s32 %0, %0, %1; " : "+r"(x[j]) : "r"(b));
        if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6A;" : "+r"(x[j]) : "r"(b), "r"(c));
        if (OP == OP_LEAHI) asm volatile("{.reg .u32 t; shr.u32 t, %1, 24; add.s32 %0, %0, t;}" : "+r"(x[j]) : "r"(b));
        if (OP == OP_TDIV) x[j] = x[j] / 256 + y;
        if (OP == OP_DP2A) asm volatile("dp2a.lo.s32.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
        if (OP == OP_VIMNMX) asm volatile("min.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));
        if (OP == OP_VIMNMX_RELU) asm volatile("min.relu.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));
        if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x[j]) : "r"(b));
        if (OP == OP_I2IP) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
        if (OP == OP_VIADDMNMX_S16X2) x[j] = (int)__viaddmin_s16x2_relu((unsigned)x[j], (unsigned)b, (unsigned)c);
        if (OP == OP_VIADD3_RELU) x[j] = __viaddmin_s32_relu(x[j], b, c);
        if (OP == OP_IMAD_IADD) {
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[(j + 1) % ILP]) : "r"(z));
        }
        if (OP == OP_IMAD_IADD_2_1) {
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[(j + 3) % ILP]) : "r"(c), "r"(b));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[(j + 1) % ILP]) : "r"(z));
        }
        if (OP == OP_IMAD_SHF) {
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
          asm volatile("shr.s32 %0, %0, %1;" : "+r"(x[(j + 1) % ILP]) : "r"(b));
        }
        if (OP == OP_IMAD_LOP_SHF) {
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x6A;" : "+r"(x[(j + 1) % ILP]) : "r"(b), "r"(c));
          asm volatile("shr.s32 %0, %0, %1;" : "+r"(x[(j + 2) % ILP]) : "r"(b));
        }
        if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(f0), "f"(1.0f));
        if (OP == OP_FFMA_IMAD) {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(f0), "f"(1.0f));
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b), "r"(c));
        }
        if (OP == OP_FFMA_IADD) {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(f0), "f"(1.0f));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(y));
        }
        if (OP == OP_MULHI) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(b));
        if (OP == OP_SEL) asm volatile("{.reg .pred p; setp.lt.s32 p, %1, 0; selp.s32 %0, %0, %2, p;}" : "+r"(x[j]) : "r"(b), "r"(c));
        if (OP == OP_I2F_F2I) {
          float t;
          asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(x[j]));
          asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(x[j]) : "f"(t));
        }
        if (OP == OP_SHFL) x[j] = __shfl_xor_sync(0xffffffffu, x[j], 1);
        if (OP == OP_LDS128) {
          int4 v = sm[(x[j] & 255)];
          x[j] = v.x ^ v.w;
        }
        if (OP == OP_IABS) asm volatile("{.reg .s32 t; abs.s32 t, %0; sub.s32 %0, %1, t;}" : "+r"(x[j]) : "r"(y));
        if (OP == OP_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"((double)f0), "d"(1.0));
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += x[j] + (int)f[j] + (int)d[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(int *d_out, int sms, int clock_khz) {
  const int threads = 256, blocks = sms * 4;  // 32 warps per SM
  const int iters = (OP == OP_I2F_F2I || OP == OP_SHFL || OP == OP_LDS128 || OP == OP_DFMA || OP == OP_MULHI) ? 400 : 2000;
  bench_kernel<OP><<<blocks, threads>>>(d_out, 10, 3, 7, 1.0001f);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int t = 0; t < 3; t++) {
    CHECK(cudaEventRecord(e0));
    bench_kernel<OP><<<blocks, threads>>>(d_out, iters, 3, 7, 1.0001f);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double warp_instr = (double)blocks * (threads / 32) * (double)iters * REP * ILP * op_instr[OP];
  const double per_s = warp_instr / (best * 1e-3);
  const double per_clk_sm = per_s / ((double)clock_khz * 1e3) / sms;
  printf("%-42s %8.3f ms  %7.2f Gwarp-instr/s  %5.2f warp-instr/clk/SM (at %d MHz nominal)\n", op_name[OP],
         best, per_s * 1e-9, per_clk_sm, clock_khz / 1000);
}

// ---- semantic probes: what do the packed / saturating instructions compute? ----
__global__ void probe_kernel(int *out) {
  int a = 300, b = -5, c = 0x11223344;
  int r0, r1, r2;
  asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r0) : "r"(a), "r"(b), "r"(c));
  asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r1) : "r"(17), "r"(200), "r"(0));
  asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r2) : "r"((int)0xFFFE0003), "r"(0x0A000007), "r"(0));
  out[0] = r0;
  out[1] = r1;
  out[2] = r2;  // expect 3*7 + (-2)*0 = 21
  int r3;
  asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(r3) : "r"((int)0xFFFE0003), "r"(0x0A000007), "r"(0));
  out[3] = r3;  // expect 3*0 + (-2)*10 = -20
  out[4] = (int)__viaddmin_s16x2_relu(0x00100020u, 0xFFF000F0u, 0x00FF00FFu);  // lanes: 0x20+0xF0=0x110->0xFF ; 0x10+(-16)=0
  out[5] = __viaddmin_s32_relu(100, 300, 255);
  out[6] = __viaddmin_s32_relu(100, -300, 255);
  int r7;
  asm volatile("min.relu.s32 %0, %1, %2;" : "=r"(r7) : "r"(-7), "r"(255));
  out[7] = r7;
}

int main() {
  int dev = 0;
  CHECK(cudaSetDevice(dev));
  cudaDeviceProp p;
  CHECK(cudaGetDeviceProperties(&p, dev));
  int clock_khz = 0;
  CHECK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
  printf("device %s, %d SMs, nominal clock %d MHz, L2 %d MB\n", p.name, p.multiProcessorCount, clock_khz / 1000,
         p.l2CacheSize >> 20);
  int *d_out;
  CHECK(cudaMalloc(&d_out, sizeof(int) * 256 * p.multiProcessorCount * 4));
  probe_kernel<<<1, 1>>>(d_out);
  int h[8];
  CHECK(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
  printf("probe cvt.pack.sat.u8.s32(a=300,b=-5,c=0x11223344) = 0x%08x\n", h[0]);
  printf("probe cvt.pack.sat.u8.s32(a=17,b=200,c=0)          = 0x%08x\n", h[1]);
  printf("probe dp2a.lo.s32.u32(0xFFFE0003, 0x0A000007)      = %d (expect 21)\n", h[2]);
  printf("probe dp2a.hi.s32.u32(0xFFFE0003, 0x0A000007)      = %d (expect -20)\n", h[3]);
  printf("probe viaddmin_s16x2_relu                          = 0x%08x (expect 0x000000ff)\n", h[4]);
  printf("probe viaddmin_s32_relu(100,300,255)=%d (100,-300,255)=%d ; min.relu(-7,255)=%d\n", h[5], h[6], h[7]);
  const int sms = p.multiProcessorCount;
#define RUN(op) run<op>(d_out, sms, clock_khz);
  RUN(OP_IMAD) RUN(OP_IADD) RUN(OP_IADD3) RUN(OP_SHF) RUN(OP_LOP3) RUN(OP_LEAHI) RUN(OP_TDIV) RUN(OP_DP2A)
  RUN(OP_VIMNMX) RUN(OP_VIMNMX_RELU) RUN(OP_PRMT) RUN(OP_I2IP) RUN(OP_VIADDMNMX_S16X2) RUN(OP_VIADD3_RELU)
  RUN(OP_IMAD_IADD) RUN(OP_IMAD_IADD_2_1) RUN(OP_IMAD_SHF) RUN(OP_IMAD_LOP_SHF) RUN(OP_FFMA) RUN(OP_FFMA_IMAD)
  RUN(OP_FFMA_IADD) RUN(OP_MULHI) RUN(OP_SEL) RUN(OP_I2F_F2I) RUN(OP_SHFL) RUN(OP_LDS128) RUN(OP_IABS) RUN(OP_DFMA)
  return 0;
}
