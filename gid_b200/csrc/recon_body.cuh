// recon_body.cuh -- per-thread body of the reconstruction kernels: one thread
// reconstructs one MCU (all its Y blocks + Cb + Cr) and writes its pixels.
//
// Reference semantics reproduced (file:line in the GID tree):
//   geometry / MCU raster            gid-decoding_jpg.adb:1187-1219, :1945-1951
//   dequant (natural-order table)    :771, :785, :1771-1775, :1820-1822
//   IDCT                             :790-907           (recon_math.cuh)
//   upsampling = replication         :974-1008: full-res pixel (px, py) of the MCU
//                                    reads component sample (px / ux, py / uy)
//   colour + clipping to W x H       :1018-1054 (rows/cols beyond the image are
//                                    skipped, :1023, :1026)
//   output layout                    test/benchmark.adb:67-81: top-down RGB8
#pragma once

#include "recon_math.cuh"

namespace gidk {

// Sampling layouts with a dedicated fused kernel.  Anything else GID accepts
// (integer up-factors, e.g. 4:4:0, 4:1:1) goes through the generic two-pass
// kernels in recon_kernels.cu.
enum Layout : int { LAYOUT_GREY = 0, LAYOUT_444 = 1, LAYOUT_422 = 2, LAYOUT_420 = 3, LAYOUT_GENERIC = 4 };

// Output pixel formats, = GIDCUDA_OUT_* of include/gidcuda.h
enum OutFormat : int { OUT_RGB8 = 0, OUT_BGR8_BOTTOM_UP = 1 };

// Per-frame descriptor in device memory (one per image of a launch).
struct FrameDev {
  const int16_t *plane[3];  // block-major [by][bx][64] int16, natural order, not de-quantised
  uint8_t *pixels;          // output, rows `stride` bytes apart
  uint32_t stride;
  int32_t width, height;
  int32_t mcux, mcuy;       // MCU grid (ceil(W / mcu_w), ceil(H / mcu_h)), :1948-1949
  int32_t bx[3], by[3];     // block grid per plane = mcu grid * sampling
  int32_t samp_h[3], samp_v[3];
  int32_t out_format;
  uint32_t tile_base;       // index of this frame's first tile in the launch
  // Quantisation tables, natural order, two per 32-bit word:
  //   q8 == 1 (all entries <= 255): word k = q[2k] | q[2k+1] << 24   (dp2a operand)
  //   q8 == 0:                      word k = q[2k] | q[2k+1] << 16
  uint32_t qw[3][32];
  int32_t q8;
  // generic layouts only: 8-bit sample planes written by the IDCT pass
  uint8_t *samples[3];
};

#if defined(__CUDACC__)
typedef ::uint4 u32x4;
typedef ::uint2 u32x2;
#else
struct u32x4 { uint32_t x, y, z, w; };
struct alignas(8) u32x2 { uint32_t x, y; };
#endif

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ u32x4 ld_block_row(const int16_t *p) {
  return __ldg(reinterpret_cast<const uint4 *>(p));
}
// a = two signed 16-bit lanes, b = four unsigned bytes: a.lo*b0 + a.hi*b1 + c
__device__ __forceinline__ int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// a.lo*b2 + a.hi*b3 + c
__device__ __forceinline__ int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
#else
inline u32x4 ld_block_row(const int16_t *p) {
  u32x4 v;
  const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
  v.x = q[0]; v.y = q[1]; v.z = q[2]; v.w = q[3];
  return v;
}
inline int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
  return (int)(int16_t)(a & 0xFFFF) * (int)(b & 0xFF) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xFF) + c;
}
inline int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
  return (int)(int16_t)(a & 0xFFFF) * (int)((b >> 16) & 0xFF) + (int)(int16_t)(a >> 16) * (int)(b >> 24) + c;
}
#endif

// De-quantise one coefficient pair (two int16 in w) with table word q.
template <bool Q8>
GID_HD void dequant_pair(uint32_t w, uint32_t q, int &lo, int &hi) {
  if (Q8) {
    lo = dp2a_lo_su(w, q, 0);  // w.lo * q[7:0]   (q[15:8] = 0)
    hi = dp2a_hi_su(w, q, 0);  // w.hi * q[31:24] (q[23:16] = 0)
  } else {
    lo = wmul((int)(int16_t)(w & 0xFFFFu), (int)(q & 0xFFFFu));
    hi = wmul((int)w >> 16, (int)(q >> 16));
  }
}

// Load + de-quantise + inverse DCT of one block.  qw: the component's 32
// table words (in shared memory on the device).  Result: un-clipped samples.
template <bool Q8>
GID_HD void load_idct_block(const int16_t *src, const uint32_t *qw, int (&b)[64]) {
  u32x4 raw[8];
#pragma unroll
  for (int r = 0; r < 8; r++) raw[r] = ld_block_row(src + 8 * r);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    dequant_pair<Q8>(raw[r].x, qw[4 * r + 0], b[8 * r + 0], b[8 * r + 1]);
    dequant_pair<Q8>(raw[r].y, qw[4 * r + 1], b[8 * r + 2], b[8 * r + 3]);
    dequant_pair<Q8>(raw[r].z, qw[4 * r + 2], b[8 * r + 4], b[8 * r + 5]);
    dequant_pair<Q8>(raw[r].w, qw[4 * r + 3], b[8 * r + 6], b[8 * r + 7]);
  }
  idct_8x8(b);
}

GID_HD uint32_t pack4(int a, int b, int c, int d) {
  return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}

// Clip + pack the 64 samples of a block into 16 words (row r -> words 2r, 2r+1).
GID_HD void pack_block(const int (&b)[64], uint32_t (&p)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++)
    p[i] = pack4(clip255(b[4 * i]), clip255(b[4 * i + 1]), clip255(b[4 * i + 2]), clip255(b[4 * i + 3]));
}

GID_HD int byte_of(uint32_t w, int k) { return (int)((w >> (8 * k)) & 0xFFu); }

// Write one row of 8 pixels (r/g/b un-clipped ints) of a block whose left edge
// is pixel column x0, on image row y.  Fast path: the 24 bytes are stored as
// three 8-byte words when the row lies fully inside the image and is 8-byte
// aligned; otherwise byte by byte with clipping to the image (:1023, :1026).
GID_HD void store_row8(const FrameDev &f, int x0, int y, const int (&r)[8], const int (&g)[8],
                       const int (&bl)[8]) {
  if (y >= f.height || x0 >= f.width) return;
  int c0[8], c2[8];
  int yy = y;
  if (f.out_format == OUT_BGR8_BOTTOM_UP) {
    // test/to_bmp.adb:94-146: bottom-up rows, B G R byte order
    yy = f.height - 1 - y;
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = bl[i]; c2[i] = r[i]; }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = r[i]; c2[i] = bl[i]; }
  }
  uint8_t *dst = f.pixels + (size_t)yy * f.stride + (size_t)x0 * 3;
  if (x0 + 8 <= f.width && (((uintptr_t)dst) & 7u) == 0) {
    uint32_t w[6];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = 4 * h;
      w[3 * h + 0] = pack4(clip255(c0[i]), clip255(g[i]), clip255(c2[i]), clip255(c0[i + 1]));
      w[3 * h + 1] = pack4(clip255(g[i + 1]), clip255(c2[i + 1]), clip255(c0[i + 2]), clip255(g[i + 2]));
      w[3 * h + 2] = pack4(clip255(c2[i + 2]), clip255(c0[i + 3]), clip255(g[i + 3]), clip255(c2[i + 3]));
    }
    u32x2 *d2 = reinterpret_cast<u32x2 *>(dst);
    u32x2 v;
    v.x = w[0]; v.y = w[1]; d2[0] = v;
    v.x = w[2]; v.y = w[3]; d2[1] = v;
    v.x = w[4]; v.y = w[5]; d2[2] = v;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (x0 + i < f.width) {
        dst[3 * i + 0] = (uint8_t)clip255(c0[i]);
        dst[3 * i + 1] = (uint8_t)clip255(g[i]);
        dst[3 * i + 2] = (uint8_t)clip255(c2[i]);
      }
    }
  }
}

// One MCU of a grey frame: a single block (single-component scans use 8x8
// MCUs, :1952-1971).  R = G = B = Y (:1036-1038).
template <bool Q8>
GID_HD void mcu_grey(const FrameDev &f, const uint32_t *qw, int mx, int my) {
  int b[64];
  load_idct_block<Q8>(f.plane[0] + ((size_t)my * f.bx[0] + mx) * 64, qw, b);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    int v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = b[8 * r + c];
    store_row8(f, mx * 8, my * 8 + r, v, v, v);
  }
}

// One MCU of a YCbCr frame with luma sampling H x V in {1x1, 2x1, 2x2} and
// 1x1 chroma: 4:4:4, 4:2:2 (h2v1), 4:2:0.  Chroma blocks first (kept as packed
// bytes), then each luma block is transformed and turned into pixels.
template <int H, int V, bool Q8>
GID_HD void mcu_ycc(const FrameDev &f, const uint32_t *qw /* [3][32] */, int mx, int my) {
  uint32_t cbp[16], crp[16];
  {
    int b[64];
    load_idct_block<Q8>(f.plane[1] + ((size_t)my * f.bx[1] + mx) * 64, qw + 32, b);
    pack_block(b, cbp);
    load_idct_block<Q8>(f.plane[2] + ((size_t)my * f.bx[2] + mx) * 64, qw + 64, b);
    pack_block(b, crp);
  }
#pragma unroll
  for (int sy = 0; sy < V; sy++) {
#pragma unroll
    for (int sx = 0; sx < H; sx++) {
      const int bxi = mx * H + sx, byi = my * V + sy;
      int b[64];
      load_idct_block<Q8>(f.plane[0] + ((size_t)byi * f.bx[0] + bxi) * 64, qw, b);
      // chroma rows covering this luma block: V == 1 -> 8 rows, V == 2 -> 4 rows
#pragma unroll
      for (int cr = 0; cr < 8 / V; cr++) {
        const int crow = sy * (8 / V) + cr;  // row inside the chroma block
        // chroma samples under the 8 pixel columns: H == 1 -> 8 samples, H == 2 -> 4
        ChromaTerms t[8 / H];
#pragma unroll
        for (int k = 0; k < 8 / H; k++) {
          const int ccol = sx * (8 / H) + k;  // column inside the chroma block
          const uint32_t wb = cbp[crow * 2 + (ccol >> 2)], wr = crp[crow * 2 + (ccol >> 2)];
          t[k] = chroma_terms(byte_of(wb, ccol & 3), byte_of(wr, ccol & 3));
        }
#pragma unroll
        for (int ry = 0; ry < V; ry++) {
          const int row = cr * V + ry;  // row inside the luma block
          int r[8], g[8], bl[8];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const int yv = clip255(b[8 * row + c]);
            const ChromaTerms &tt = t[c / H];
            r[c] = yv + tt.r;
            g[c] = yv + tt.g;
            bl[c] = yv + tt.b;
          }
          store_row8(f, bxi * 8, byi * 8 + row, r, g, bl);
        }
      }
    }
  }
}

}  // namespace gidk
