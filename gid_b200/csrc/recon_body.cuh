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
//
// The body is parameterised by a "block source": where the 8 rows of 8 int16
// coefficients of a block come from -- global memory (LdgSource, also the host
// build used by the CPU tests) or the TMA-filled shared-memory ring of
// recon_kernels.cu (TmaSource).
#pragma once

#include "recon_math.cuh"

namespace gidk {

// Sampling layouts with a dedicated fused kernel: grey, and Y with 1x1 chroma for luma samplings
// 1x1 (4:4:4), 2x1 (4:2:2), 2x2 (4:2:0), 1x2 (4:4:0) and 4x1 (4:1:1).  Anything else GID accepts
// (integer up-factors, e.g. 4:1:0, or chroma that is not 1x1) goes through the generic two-pass
// kernels in recon_kernels.cu.
enum Layout : int {
  LAYOUT_GREY = 0, LAYOUT_444 = 1, LAYOUT_422 = 2, LAYOUT_420 = 3, LAYOUT_440 = 4, LAYOUT_411 = 5,
  LAYOUT_GENERIC = 6
};
constexpr int kFusedLayouts = 6;
// luma sampling and items per tile (Cb, Cr, then H x V luma items; grey: one) of a fused layout
GID_HD int layout_h(int layout) { return layout == LAYOUT_422 || layout == LAYOUT_420 ? 2 : (layout == LAYOUT_411 ? 4 : 1); }
GID_HD int layout_v(int layout) { return layout == LAYOUT_420 || layout == LAYOUT_440 ? 2 : 1; }
GID_HD int layout_items(int layout) { return layout == LAYOUT_GREY ? 1 : 2 + layout_h(layout) * layout_v(layout); }
// grey and 4:4:4: one block per component per MCU, tiles are linear runs of MCUs that may wrap over MCU rows
GID_HD bool layout_linear(int layout) { return layout == LAYOUT_GREY || layout == LAYOUT_444; }

// Output pixel formats, = GIDCUDA_OUT_* of include/gidcuda.h
enum OutFormat : int { OUT_RGB8 = 0, OUT_BGR8_BOTTOM_UP = 1, OUT_RGBA8 = 2 };

// FrameDev::flags
enum FrameFlags : uint32_t {
  FRAME_ALIGNED8 = 1u,   // pixel base and stride are multiples of 8: rows of a block go out as 3 x 8 bytes
  FRAME_ALIGNED16 = 2u,  // ... multiples of 16: RGBA rows of a block go out as 2 x 16 bytes
  // GIDCUDA_FLAG_ROTATE_180 in the fused kernels: pixel (x, r) lands at column W - 1 - x, row H - 1 - r
  // (test/to_bmp.adb:104-122).  The alignment flags above then also promise W % 8 == 0 (W % 4 for RGBA), so
  // that the mirrored runs of whole blocks start on the same boundaries.
  FRAME_ROT180 = 4u,
  // GIDCUDA_FLAG_ROTATE_90 / _270 in the fused kernels: pixel (x, r) lands at column r, row W - 1 - x / at
  // column H - 1 - r, row x of an image H wide and W tall.  FRAME_ROT270 mirrors the runs: the alignment flags
  // then also promise H % 8 == 0 (H % 4 for RGBA).
  FRAME_ROT90 = 8u,
  FRAME_ROT270 = 16u,
};

// Display orientation a kernel instance applies (= Geometry::rotate, GIDCUDA_FLAG_ROTATE_* >> 9)
enum Turn : int { TURN_NONE = 0, TURN_90 = 1, TURN_180 = 2, TURN_270 = 3 };

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
  uint32_t tile_rows;       // frames turned by a quarter: MCU rows per tile (32, 16, 8 or 4; 32 / tile_rows MCU columns)
  // Quantisation tables, natural order, two per 32-bit word:
  //   q8 == 1 (all entries <= 255): word k = q[2k] | q[2k+1] << 24   (dp2a operand)
  //   q8 == 0:                      word k = q[2k] | q[2k+1] << 16
  uint32_t qw[3][32];
  int32_t q8;
  uint32_t flags;
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
inline int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
  return (int)(int16_t)(a & 0xFFFF) * (int)(b & 0xFF) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xFF) + c;
}
inline int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
  return (int)(int16_t)(a & 0xFFFF) * (int)((b >> 16) & 0xFF) + (int)(int16_t)(a >> 16) * (int)(b >> 24) + c;
}
#endif

// Block source reading global memory directly (fallback kernel, generic pass 1,
// and the host build).
struct LdgSource {
  const FrameDev *f;
  static constexpr bool kItemsCanBeSkipped = true;  // no ring of fetches in flight: an item nobody needs is not read
  // The block of the NEXT item, asked for while this item is computed (the plain loads have no ring that runs
  // ahead: without this every item starts with the latency of a trip to DRAM).
  GID_HD void prefetch(bool active, int comp, int bx, int by) const {
#if defined(__CUDA_ARCH__)
    if (active) {
      const int16_t *p = f->plane[comp] + ((size_t)by * (size_t)f->bx[comp] + (size_t)bx) * 64;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }
#else
    (void)active; (void)comp; (void)bx; (void)by;
#endif
  }
  // where the packed samples of the luma block being emitted go; nothing to hand back
  GID_HD u32x2 *luma_scratch(u32x2 *own) const { return own; }
  GID_HD void release() const {}
  GID_HD void load(bool active, int comp, int bx, int by, u32x4 (&raw)[8]) const {
    if (!active) {  // lane without an MCU: nothing to read, contributes zeros
#pragma unroll
      for (int r = 0; r < 8; r++) { raw[r].x = 0; raw[r].y = 0; raw[r].z = 0; raw[r].w = 0; }
      return;
    }
    const int16_t *p = f->plane[comp] + ((size_t)by * (size_t)f->bx[comp] + (size_t)bx) * 64;
#pragma unroll
    for (int r = 0; r < 8; r++) {
#if defined(__CUDA_ARCH__)
      raw[r] = __ldg(reinterpret_cast<const uint4 *>(p + 8 * r));
#else
      const uint32_t *q = reinterpret_cast<const uint32_t *>(p + 8 * r);
      raw[r].x = q[0]; raw[r].y = q[1]; raw[r].z = q[2]; raw[r].w = q[3];
#endif
    }
  }
};

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

// De-quantise + inverse DCT of one block held as 8 raw rows.  qw: the
// component's 32 table words.  Result: un-clipped samples.
template <bool Q8>
GID_HD void dequant_idct(const u32x4 (&raw)[8], const uint32_t *qw, int (&b)[64]) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    dequant_pair<Q8>(raw[r].x, qw[4 * r + 0], b[8 * r + 0], b[8 * r + 1]);
    dequant_pair<Q8>(raw[r].y, qw[4 * r + 1], b[8 * r + 2], b[8 * r + 3]);
    dequant_pair<Q8>(raw[r].z, qw[4 * r + 2], b[8 * r + 4], b[8 * r + 5]);
    dequant_pair<Q8>(raw[r].w, qw[4 * r + 3], b[8 * r + 6], b[8 * r + 7]);
  }
  idct_8x8(b);
}

// Clip (:755-756) + pack the 64 samples of a block to bytes and store them as 8
// rows of 8 bytes, `stride` u32x2 apart (the per-lane scratch plane).
//   TR   the block is stored transposed (scratch row c = column c of the block): what the instances that
//        turn the image by a quarter emit from.  The samples sit in registers: the transposition is the
//        order in which they are packed, no instruction more.
template <bool TR = false>
GID_HD void store_block_packed(const int (&b)[64], u32x2 *dst, int stride) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    u32x2 v;
    if (TR) {
      v.x = pack_sat4(b[r], b[8 + r], b[16 + r], b[24 + r]);
      v.y = pack_sat4(b[32 + r], b[40 + r], b[48 + r], b[56 + r]);
    } else {
      v.x = pack_sat4(b[8 * r + 0], b[8 * r + 1], b[8 * r + 2], b[8 * r + 3]);
      v.y = pack_sat4(b[8 * r + 4], b[8 * r + 5], b[8 * r + 6], b[8 * r + 7]);
    }
    dst[r * stride] = v;
  }
}

// "Does any lane of the warp see p?"  Every lane of the warp votes; lanes that
// own no MCU in this tile vote false.  The host build runs one thread at a time,
// so there the answer is p itself and every block picks its own variant.
GID_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
  return __any_sync(0xFFFFFFFFu, p) != 0;
#else
  return p;
#endif
}

// De-quantise one row (8 coefficients in 4 words) with table words q[0..3].
template <bool Q8>
GID_HD void dequant_row(const u32x4 &w, const uint32_t *q, int &c0, int &c1, int &c2, int &c3, int &c4, int &c5,
                        int &c6, int &c7) {
  dequant_pair<Q8>(w.x, q[0], c0, c1);
  dequant_pair<Q8>(w.y, q[1], c2, c3);
  dequant_pair<Q8>(w.z, q[2], c4, c5);
  dequant_pair<Q8>(w.w, q[3], c6, c7);
}

// De-quantise + inverse DCT with the work that real streams do not need left
// out.  All decisions are WARP-UNIFORM (a vote over the 32 blocks the warp holds)
// so that no lane idles in a branch it does not need, and all of them are exact:
//   * a row with no AC term in any lane: every output is 8 * dc (GID's own
//     shortcut, :810-812) -- one multiply instead of the 65-instruction row pass;
//   * rows 3 .. 7 zero in every lane (chroma at ordinary qualities, flat luma):
//     those rows are not touched and the column pass runs its 3-input form
//     (col_idct_lo3), otherwise the general column pass.
// Dense data takes the general path everywhere; the result is the same either way.
//
// Q8 is also the kernels' FAST DOMAIN: every table entry in 1 .. 255 and every |coefficient * entry| <= 2**16
// (frame_setup.hpp, fast_coef_limit; anything else runs dequant_idct_exact below).  Inside it the decisions
// taken here on RAW coefficients are the reference's, which takes them on de-quantised values (:803-812):
//   * raw != 0  <=>  raw * q != 0, and blk(4) * 2**11 is zero only for blk(4) = 0;
//   * no sum of the row pass wraps, so the general path with rounding constant 0 IS blk(0) * 8;
//   * column inputs stay below 2**23 - 32, where the column shortcut (:858-862) is the general path.

// The exact instance (Q8 = false: 16-bit tables, tables with an entry of 0, wide coefficients -- rare):
// every row de-quantised, the reference's own tests per block on the de-quantised values, both shortcuts
// carried literally.  Reproduces the reference's 32-bit wrap-around for every 16-bit input.
GID_HD void dequant_idct_exact(bool active, const u32x4 (&raw)[8], const uint32_t *qw, u32x2 *dst, int dst_stride) {
  constexpr bool Q8 = false;
  int b[64];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    int *row = &b[8 * r];
    dequant_row<Q8>(raw[r], qw + 4 * r, row[0], row[1], row[2], row[3], row[4], row[5], row[6], row[7]);
    const bool ac = row_has_ac(row[1], row[2], row[3], row[4], row[5], row[6], row[7]);
    const int flat = wmul(row[0], 8);
    if (warp_any(active && ac)) row_idct_rnd(128, row[0], row[1], row[2], row[3], row[4], row[5], row[6], row[7]);
    if (!ac) {
#pragma unroll
      for (int c = 0; c < 8; c++) row[c] = flat;
    }
  }
#pragma unroll
  for (int c = 0; c < 8; c++)
    col_idct_exact(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
  store_block_packed(b, dst, dst_stride);
}

// The fast instance.
template <bool Q8, bool LO3, bool TR = false>
GID_HD void dequant_idct_sparse(bool active, const u32x4 (&raw)[8], const uint32_t *qw, u32x2 *dst, int dst_stride) {
  int b[64];
  uint32_t ac[8];
#pragma unroll
  for (int r = 0; r < 8; r++) ac[r] = (raw[r].x & 0xFFFF0000u) | raw[r].y | raw[r].z | raw[r].w;
  const uint32_t low = (raw[3].x | raw[4].x) | (raw[5].x | raw[6].x) | raw[7].x;
  const uint32_t tail = (ac[3] | ac[4]) | (ac[5] | ac[6]) | (ac[7] | low);
  const bool tall = warp_any(active && tail != 0);
#pragma unroll
  for (int r = 0; r < 8; r++) {
    if (r >= 3 && !tall) break;
    int *row = &b[8 * r];
    if (warp_any(active && ac[r] != 0)) {
      dequant_row<Q8>(raw[r], qw + 4 * r, row[0], row[1], row[2], row[3], row[4], row[5], row[6], row[7]);
      row_idct_rnd(ac[r] != 0 ? 128 : 0, row[0], row[1], row[2], row[3], row[4], row[5], row[6], row[7]);
    } else {
      int dc, unused;
      dequant_pair<Q8>(raw[r].x, qw[4 * r], dc, unused);
      dc = wmul(dc, 8);
#pragma unroll
      for (int c = 0; c < 8; c++) row[c] = dc;
    }
  }
  // Clip + pack + store inside each variant: joining the two register sets first
  // would cost a move per sample.  LO3 = false (4:2:x kernels, where chroma is a third
  // of the blocks or less): the 3-input column pass is left out -- there its code
  // costs more instruction-cache misses than it saves instructions (measured).
  if (tall || !LO3) {
    if (!LO3 && !tall) {
#pragma unroll
      for (int i = 24; i < 64; i++) b[i] = 0;
    }
#pragma unroll
    for (int c = 0; c < 8; c++)
      col_idct(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
    store_block_packed<TR>(b, dst, dst_stride);
  } else {
#pragma unroll
    for (int c = 0; c < 8; c++)
      col_idct_lo3(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
    store_block_packed<TR>(b, dst, dst_stride);
  }
}

// Clip + pack the 64 samples of a block into 16 words (row r -> words 2r, 2r+1).
GID_HD void pack_block(const int (&b)[64], uint32_t (&p)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) p[i] = pack_sat4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]);
}

// Destination of pixel (x0, y): rows are flipped for the bottom-up format.
GID_HD uint8_t *row_ptr(const FrameDev &f, int x0, int y) {
  const int yy = f.out_format == OUT_BGR8_BOTTOM_UP ? f.height - 1 - y : y;
  return f.pixels + (size_t)yy * f.stride + (size_t)x0 * (f.out_format == OUT_RGBA8 ? 4 : 3);
}
// ... of an output image `rows` high (quarter turns: the frame's width)
GID_HD uint8_t *row_ptr(const FrameDev &f, int rows, int x0, int y) {
  const int yy = f.out_format == OUT_BGR8_BOTTOM_UP ? rows - 1 - y : y;
  return f.pixels + (size_t)yy * f.stride + (size_t)x0 * (f.out_format == OUT_RGBA8 ? 4 : 3);
}

// Slow path: byte stores with clipping to the image (:1023, :1026); used for
// partial blocks at the right edge and for frames whose rows are not 8-byte
// aligned.  w[0..5] = the 24 output bytes of the row, already packed; n = how
// many of them lie inside the image.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
void store_row8_bytes(uint8_t *dst, int n, int skip, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                      uint32_t w5) {
  // bytes skip .. skip + n - 1 of the 24 (skip > 0: a mirrored row of a block cut by the image's edge)
  const uint32_t w[6] = {w0, w1, w2, w3, w4, w5};
  for (int i = 0; i < n; i++) dst[i] = (uint8_t)(w[(i + skip) >> 2] >> (8 * ((i + skip) & 3)));
}

// One row of 8 pixels (24 packed bytes in w0..w5) at dst.  nbytes == 24 and an
// 8-byte aligned dst (`fast`): the row leaves as three 8-byte stores.
GID_HD void store_row8(uint8_t *dst, bool fast, int nbytes, int skip, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                       uint32_t w4, uint32_t w5) {
  if (fast) {
    u32x2 *d2 = reinterpret_cast<u32x2 *>(dst);
    u32x2 v;
    v.x = w0; v.y = w1; d2[0] = v;
    v.x = w2; v.y = w3; d2[1] = v;
    v.x = w4; v.y = w5; d2[2] = v;
  } else {
    store_row8_bytes(dst, nbytes, skip, w0, w1, w2, w3, w4, w5);
  }
}

// RGBA: one row of 8 pixels = 8 words [R, G, B, 255].  Partial blocks write whole pixels.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
void store_row8_rgba_words(uint8_t *dst, int npix, int skip, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                           uint32_t w5, uint32_t w6, uint32_t w7) {
  const uint32_t w[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  for (int i = 0; i < npix; i++)
    for (int k = 0; k < 4; k++) dst[4 * i + k] = (uint8_t)(w[i + skip] >> (8 * k));
}
GID_HD void store_row8_rgba(uint8_t *dst, bool fast, int npix, int skip, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                            uint32_t w4, uint32_t w5, uint32_t w6, uint32_t w7) {
  if (fast) {
    u32x4 *d4 = reinterpret_cast<u32x4 *>(dst);
    u32x4 v;
    v.x = w0; v.y = w1; v.z = w2; v.w = w3; d4[0] = v;
    v.x = w4; v.y = w5; v.z = w6; v.w = w7; d4[1] = v;
  } else {
    store_row8_rgba_words(dst, npix, skip, w0, w1, w2, w3, w4, w5, w6, w7);
  }
}

// Per-thread scratch for the packed 8-bit samples of the current MCU: 8 rows of
// 8 bytes for Cb, Cr and the luma block being emitted.  On the device it lives
// in shared memory as [row][lane] (row r of lane t at base[r * stride + t],
// conflict-free 8-byte accesses); the host build uses plain arrays (stride 1).
// Keeping these bytes out of registers lets the kernel LOOP over the blocks of
// an MCU and over pixel rows, so that one copy of the IDCT (~20 KiB of code)
// and one copy of the row emitter serve every block: a fully unrolled MCU body
// (~100 KiB) thrashes the instruction cache (measured: no_instructions was the
// top stall reason, icc hit rate 76 %).
struct Scratch {
  u32x2 *cb, *cr, *y;
  int stride;
};

// byte k (0..3) of w, zero-extended: one PRMT
template <int K>
GID_HD int byte_of(uint32_t w) { return (int)prmt(w, 0u, 0x4440u + K); }

// Emit the pixel rows of one luma block from the scratch (Step 4-6 of GID:
// upsampling by replication, colour transformation, output).
//   H, V   luma sampling (1 or 2); GREY = no chroma
//   sx, sy position of the block inside the MCU
// The destination pointer walks down the image one row per step (up for the
// bottom-up format); rows below the image (:1023) are skipped by count.
//   RGBA   4 bytes per pixel (a separate instance, so that the 3-byte loop stays lean)
//   FAST   the whole block lies inside the image (8 pixels x 8 rows) and its rows are aligned:
//          no per-row bounds checks, no call to the byte-wise store in the loop
//   MIRROR the run of a block row is written mirrored (pixel x0 + k at column W - 1 - x0 - k).  The sample
//          words are byte-reversed as they are read, so that the arithmetic and the packing below produce
//          the mirrored run as it is.
//   FLIP   the rows go out bottom to top (row y0 + k at row H - 1 - y0 - k)
//   TURN   the scratch holds the blocks TRANSPOSED (store_block_packed<true>) and the caller passes the block's
//          place in the transposed image -- (y0, x0), (sy, sx), with H and V swapped: the output image is
//          f.height wide and f.width high.
//          FRAME_ROT180 = MIRROR + FLIP, FRAME_ROT90 = TURN + FLIP, FRAME_ROT270 = TURN + MIRROR
//          (test/to_bmp.adb:104-122); separate kernel instances, so that frames that are not rotated run the
//          code they always ran.
GID_HD uint32_t brev4(uint32_t w) { return prmt(w, 0u, 0x0123u); }
template <int H, int V, bool GREY, bool RGBA, bool FAST, bool MIRROR = false, bool FLIP = false, bool TURN = false>
GID_HD void emit_rows(const FrameDev &f, const Scratch &sc, int x0, int y0, int sx, int sy) {
  constexpr bool REV = MIRROR;
  const int fw = TURN ? f.height : f.width, fh = TURN ? f.width : f.height;
  constexpr bool rgba = RGBA;
#if defined(GIDK_EXTRACT_ADD_ALL)
  constexpr bool kExtractAdd = true;
#elif defined(GIDK_EXTRACT_ADD_NONE)
  constexpr bool kExtractAdd = false;
#else
  constexpr bool kExtractAdd = H * V > 1;  // a chroma term serves several pixels
#endif
  int npix = fw - x0;
  if (npix > 8) npix = 8;
  const int nbytes = npix * 3;
  // !FAST: rows that are whole and aligned still leave as 8-byte stores
  const bool fast = FAST || ((f.flags & (rgba ? FRAME_ALIGNED16 : FRAME_ALIGNED8)) != 0 && npix == 8);
  const bool bottom_up = f.out_format == OUT_BGR8_BOTTOM_UP;
  const ColourOrder co = colour_order(bottom_up);
  // REV: the run starts where its LAST pixel lands; of a mirrored run cut by the edge the leading
  // 8 - npix pixels of the 8 computed are not written
  const int skip_pix = REV ? 8 - npix : 0;
  uint8_t *dst = row_ptr(f, fh, MIRROR ? fw - x0 - npix : x0, FLIP ? fh - 1 - y0 : y0);
  const long long step = (bottom_up != FLIP) ? -(long long)f.stride : (long long)f.stride;
  const int rows_left = FAST ? 8 : fh - y0;  // >= 1
  const int crow0 = sy * (8 / V);                   // first row inside the chroma block
  // (The byte selectors of the extract-add are literals: the compiler re-materialises them into
  // uniform registers on every trip, 9 UMOV, but holding them in registers instead was measured
  // slower on 4:2:0, -1.6 %.)
  const uint32_t pick[4] = {0x1u, 0x100u, 0x10000u, 0x1000000u};
  // colour kernels: rolled (their hot loop is at the edge of the 32 KiB instruction cache; measured +3 %);
  // grey: two rows per trip
#pragma unroll(GREY ? 2 : 1)
  for (int i = 0; i < 8 / V; i++) {
    if (!FAST && i * V >= rows_left) break;
    u32x2 cbc, crc;
    ChromaTerms t[8 / H];
    if (!GREY) {
      cbc = sc.cb[(crow0 + i) * sc.stride];
      crc = sc.cr[(crow0 + i) * sc.stride];
      if (REV) {  // columns 7 .. 0: the block's word sx (H = 2) / half-word (H = 4) is found mirrored below
        const u32x2 b = cbc, r = crc;
        cbc.x = brev4(b.y); cbc.y = brev4(b.x);
        crc.x = brev4(r.y); crc.y = brev4(r.x);
      }
      const int sxm = REV ? H - 1 - sx : sx;
      if (H == 1) {
        t[0] = chroma_terms_packed<0>(co, cbc.x, crc.x); t[1] = chroma_terms_packed<1>(co, cbc.x, crc.x);
        t[2] = chroma_terms_packed<2>(co, cbc.x, crc.x); t[3] = chroma_terms_packed<3>(co, cbc.x, crc.x);
        t[4 / H] = chroma_terms_packed<0>(co, cbc.y, crc.y); t[5 / H] = chroma_terms_packed<1>(co, cbc.y, crc.y);
        t[6 / H] = chroma_terms_packed<2>(co, cbc.y, crc.y); t[7 / H] = chroma_terms_packed<3>(co, cbc.y, crc.y);
      } else {
        if (H == 2) {
          // block sx covers chroma columns 4*sx .. 4*sx+3, i.e. word sx of the row
          const uint32_t wb = sxm ? cbc.y : cbc.x, wr = sxm ? crc.y : crc.x;
          t[0] = chroma_terms_packed<0>(co, wb, wr); t[1] = chroma_terms_packed<1>(co, wb, wr);
          constexpr int k2 = H == 2 ? 2 : 0, k3 = H == 2 ? 3 : 0;  // (only H = 2 runs this; keeps the index in range elsewhere)
          t[k2] = chroma_terms_packed<2>(co, wb, wr); t[k3] = chroma_terms_packed<3>(co, wb, wr);
        } else {
          // H = 4: block sx covers chroma columns 2*sx, 2*sx+1: half-word (sx & 1) of word (sx >> 1)
          const uint32_t sh = (uint32_t)(sxm & 1) * 16u;
          const uint32_t wb = ((sxm & 2) ? cbc.y : cbc.x) >> sh, wr = ((sxm & 2) ? crc.y : crc.x) >> sh;
          t[0] = chroma_terms_packed<0>(co, wb, wr); t[1] = chroma_terms_packed<1>(co, wb, wr);
        }
      }
    }
#pragma unroll
    for (int ry = 0; ry < V; ry++) {
      if (!FAST && V > 1 && i * V + ry >= rows_left) break;
      u32x2 yv = sc.y[(i * V + ry) * sc.stride];
      if (REV) {
        const u32x2 y = yv;
        yv.x = brev4(y.y); yv.y = brev4(y.x);
      }
      if (GREY && rgba) {
        store_row8_rgba(dst, fast, npix, skip_pix, prmt(yv.x, 255u, 0x4000u), prmt(yv.x, 255u, 0x4111u), prmt(yv.x, 255u, 0x4222u),
                        prmt(yv.x, 255u, 0x4333u), prmt(yv.y, 255u, 0x4000u), prmt(yv.y, 255u, 0x4111u),
                        prmt(yv.y, 255u, 0x4222u), prmt(yv.y, 255u, 0x4333u));
      } else if (GREY) {
        // R = G = B = Y (:1036-1038)
        store_row8(dst, fast, nbytes, 3 * skip_pix, prmt(yv.x, 0u, 0x1000u), prmt(yv.x, 0u, 0x2211u), prmt(yv.x, 0u, 0x3332u),
                   prmt(yv.y, 0u, 0x1000u), prmt(yv.y, 0u, 0x2211u), prmt(yv.y, 0u, 0x3332u));
      } else {
        int c0[8], c1[8], c2[8];
        if (kExtractAdd) {
          // Y + term as a byte dot product: dp4a(yword, 1 << 8k, term) picks luma byte k and adds
          // it on the FMA pipe (IDP.4A) -- no PRMT per pixel, and the term's shift is done once
          // per chroma sample instead of being fused into every add (LEA.HI on the ALU pipe).
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const ChromaTerms &tt = t[c / H];
            const uint32_t yw = c < 4 ? yv.x : yv.y;
            c0[c] = dp4a_uu(yw, pick[c & 3], tt.r);
            c1[c] = dp4a_uu(yw, pick[c & 3], tt.g);
            c2[c] = dp4a_uu(yw, pick[c & 3], tt.b);
          }
        } else {
          const int ys[8] = {byte_of<0>(yv.x), byte_of<1>(yv.x), byte_of<2>(yv.x), byte_of<3>(yv.x),
                             byte_of<0>(yv.y), byte_of<1>(yv.y), byte_of<2>(yv.y), byte_of<3>(yv.y)};
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const ChromaTerms &tt = t[c / H];
            c0[c] = ys[c] + tt.r;
            c1[c] = ys[c] + tt.g;
            c2[c] = ys[c] + tt.b;
          }
        }
        if (rgba)
          store_row8_rgba(dst, fast, npix, skip_pix, pack_sat4(c0[0], c1[0], c2[0], 255), pack_sat4(c0[1], c1[1], c2[1], 255),
                          pack_sat4(c0[2], c1[2], c2[2], 255), pack_sat4(c0[3], c1[3], c2[3], 255),
                          pack_sat4(c0[4], c1[4], c2[4], 255), pack_sat4(c0[5], c1[5], c2[5], 255),
                          pack_sat4(c0[6], c1[6], c2[6], 255), pack_sat4(c0[7], c1[7], c2[7], 255));
        else
          store_row8(dst, fast, nbytes, 3 * skip_pix, pack_sat4(c0[0], c1[0], c2[0], c0[1]), pack_sat4(c1[1], c2[1], c0[2], c1[2]),
                     pack_sat4(c2[2], c0[3], c1[3], c2[3]), pack_sat4(c0[4], c1[4], c2[4], c0[5]),
                     pack_sat4(c1[5], c2[5], c0[6], c1[6]), pack_sat4(c2[6], c0[7], c1[7], c2[7]));
      }
      dst += step;
    }
  }
}

// Blocks cut by the right or bottom edge of the image (:1023, :1026) and frames whose rows are
// not aligned: the same loop with its bounds checks and byte-wise stores, kept out of line so
// that the hot loop above carries none of it.
template <int H, int V, bool GREY, bool RGBA>
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
void emit_rows_edge(const FrameDev *f, u32x2 *cb, u32x2 *cr, u32x2 *y, int stride, int x0, int y0, int sx, int sy) {
  // (scalars by value: a reference to the caller's Scratch would force it into local memory
  // on the hot path too)
  Scratch sc;
  sc.cb = cb; sc.cr = cr; sc.y = y; sc.stride = stride;
  emit_rows<H, V, GREY, RGBA, false>(*f, sc, x0, y0, sx, sy);
}

template <int H, int V, bool GREY, bool RGBA, int ROT>
GID_HD void emit_block_rows(const FrameDev &f, const Scratch &sc, int x0, int y0, int sx, int sy) {
  if (x0 >= f.width) return;
  if (ROT != TURN_NONE) {
    // Rotated frames run kernel instances of their own: a second copy of the row emitter inside the
    // ordinary instances cost frames that are not rotated 6 % (measured on 4:2:0; the kernels are bound by
    // instruction fetch, tools/icache_bench.cu)
    if (ROT == TURN_180) {
      emit_rows<H, V, GREY, RGBA, false, true, true>(f, sc, x0, y0, sx, sy);
    } else {
      // quarter turns: the block's rows are the image's columns -- the samples were stored transposed
      if (y0 >= f.height) return;
      emit_rows<V, H, GREY, RGBA, false, ROT == TURN_270, ROT == TURN_90, true>(f, sc, y0, x0, sy, sx);
    }
    return;
  }
  // The split pays where registers are plentiful (grey: +1.7 %).  The colour kernels sit at the
  // 128-register limit: there the call made two values of the tile loop spill and the second
  // copy of the loop cost instruction-cache misses (-5 % on 4:2:0, profiles/r01s), so they keep
  // the one loop with its checks.
  if (GREY) {
    const bool whole = f.width - x0 >= 8 && f.height - y0 >= 8 &&
                       (f.flags & (RGBA ? FRAME_ALIGNED16 : FRAME_ALIGNED8)) != 0;
    if (whole) emit_rows<H, V, GREY, RGBA, true>(f, sc, x0, y0, sx, sy);
    else emit_rows_edge<H, V, GREY, RGBA>(&f, sc.cb, sc.cr, sc.y, sc.stride, x0, y0, sx, sy);
  } else {
    emit_rows<H, V, GREY, RGBA, false>(f, sc, x0, y0, sx, sy);
  }
}

GID_HD void warp_sync() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}

// One item of a tile for one lane.  A tile is up to 32 consecutive MCUs; its items are
// Cb, Cr, then the luma items (a single luma item for grey).  Each item: fetch 8 raw
// rows from the block source, de-quantise, inverse DCT, clip + pack to bytes into the
// scratch; luma blocks are then turned into pixels.
//
// Which block a lane owns:
//   grey, 4:4:4 (one block per component per MCU): lane t <-> MCU t of the tile, all items.
//   4:4:0 (H = 1, V = 2): the same for every item; luma item 2 + sy is block row sy of the MCUs.
//   4:2:2 / 4:2:0 / 4:1:1 (H = 2 or 4): chroma items as above (lane t <-> MCU t = (mx, my)); the luma items
//   walk the tile's luma block rows 32 CONSECUTIVE blocks at a time -- item 2 + H*sy + part holds blocks
//   32*part .. 32*part + 31 of block row sy -- so lane t owns luma block j = 32*part + t, which
//   belongs to MCU j / H and is its block number j % H from the left.  Neighbouring lanes therefore
//   write neighbouring 24-byte runs of the same pixel rows in the same store instructions, and
//   every 32-byte sector of the output is completed at once.  (With lane <-> MCU for luma the two
//   halves of each 48-byte run were written a whole block of arithmetic apart: partially written
//   sectors were evicted and re-fetched, +38 % DRAM reads and +24 % writes on 4:2:0, ncu.)
//   The chroma samples of MCU j / 2 were left in the scratch by lane j / 2.
//   Quarter turns (ROT = TURN_90 / TURN_270): a row of the turned image is a COLUMN of the frame, so the tile
//   runs DOWN the frame -- th MCU rows x 32 / th MCU columns (FrameDev::tile_rows: 32 x 1 where the frame's
//   height fills such tiles, flatter ones where it does not) -- and the luma items walk its block columns th
//   consecutive blocks at a time (item 2 + V*sx + part: blocks th*part .. th*part + th - 1 of block column sx
//   of every MCU column): the lanes of one MCU column then write neighbouring 24-byte runs of the same rows
//   of the turned image, as above.  (With the tiles along the
//   frame's rows every lane wrote its run into a row of its own, 8 rows apart: sectors written a third at a
//   time by different warps were evicted from L2 before they were complete -- twice the DRAM traffic, 0.19 of
//   the HBM roofline on 1080p 4:2:0 batches against 0.68 unrotated; ncu, profiles/r02_ncu_summary.md.)
// Which block item `it` is for this lane (the mapping of mcu_item below, without its side effects): for the
// sources that fetch ahead by hand
template <int H, int V, bool GREY, bool VERT>
GID_HD bool item_block(int it, bool active, const FrameDev &f, int lane, int mx, int my, int &comp, int &bxi, int &byi) {
  comp = 0;
  bxi = mx;
  byi = my;
  if (GREY) return active;
  if (it < 2) {
    comp = it + 1;
  } else if (VERT && V == 1) {
    bxi = mx * H + (it - 2);
  } else if (VERT) {
    const int th = (int)f.tile_rows, r = lane & (th - 1);
    bxi = mx * H + (it - 2) / V;
    byi = V * (my - r) + th * ((it - 2) % V) + r;
    active = mx < f.mcux && byi < f.by[0];
  } else if (H == 1) {
    byi = my * V + (it - 2);
  } else {
    bxi = H * (mx - lane) + 32 * ((it - 2) % H) + lane;
    byi = my * V + (it - 2) / H;
    active = bxi < f.bx[0];
  }
  return active;
}

template <int H, int V, bool GREY, bool Q8, bool RGBA, int ROT, class Src>
GID_HD void mcu_item(int it, bool active, const FrameDev &f, const uint32_t *qw, Src &src, const Scratch &sc,
                     int lane, int mx, int my) {
  int comp = 0, sx = 0, sy = 0;
  int bxi = mx, byi = my;
  Scratch se = sc;
  const bool lane_has_mcu = active;
  // quarter turns: the tile is 32 MCUs of one COLUMN of MCUs (lane t <-> MCU (mx, my0 + t)), the mirror image of
  // everything above -- see the end of this comment block
  constexpr bool VERT = ROT == TURN_90 || ROT == TURN_270;
  if (!GREY) {
    if (it < 2) {
      comp = it + 1;
    } else if (VERT && V == 1) {
      sx = it - 2;
      bxi = mx * H + sx;
    } else if (VERT) {
      // the tile: th MCU rows x 32 / th MCU columns; lane = (column dx, row r) = (lane / th, lane % th)
      const int th = (int)f.tile_rows, r = lane & (th - 1);
      const int part = (it - 2) % V;        // which th of a block column's th * V luma blocks
      sx = (it - 2) / V;
      const int j = th * part + r;          // luma block of this lane inside its block column
      bxi = mx * H + sx;
      byi = V * (my - r) + j;               // my - r = first MCU row of the tile
      active = mx < f.mcux && byi < f.by[0];
      sy = r & (V - 1);                     // = j % V (th is a multiple of V)
      se.cb = sc.cb + ((j / V) - r);        // scratch column of MCU j / V of the same MCU column
      se.cr = sc.cr + ((j / V) - r);
      if (it == 2) warp_sync();
    } else if (H == 1) {
      sy = it - 2;
      byi = my * V + sy;
    } else {
      const int part = (it - 2) % H;        // which 32 of the block row's 32 * H luma blocks
      sy = (it - 2) / H;
      const int j = 32 * part + lane;       // luma block of this lane inside the tile's block row
      bxi = H * (mx - lane) + j;            // mx - lane = first MCU of the tile (tiles stay in one MCU row)
      byi = my * V + sy;
      active = bxi < f.bx[0];
      sx = lane & (H - 1);                  // = j % H (32 is a multiple of H)
      se.cb = sc.cb + ((j / H) - lane);     // scratch column of MCU j / H
      se.cr = sc.cr + ((j / H) - lane);
      if (it == 2) warp_sync();             // every lane's chroma samples are in the scratch
    }
  }
  // the last tile of a column of MCUs is rarely full (1080p 4:2:0: 68 MCU rows = 2 tiles and one of 4 lanes,
  // whose second half of every block column holds no block at all)
  if (VERT && Src::kItemsCanBeSkipped && !warp_any(active)) return;
  u32x4 raw[8];
  src.load(active, comp, bxi, byi, raw);
  if (Src::kItemsCanBeSkipped && it + 1 < (GREY ? 1 : 2 + H * V)) {
    int c2, bx2, by2;
    const bool a2 = item_block<H, V, GREY, VERT>(it + 1, lane_has_mcu, f, lane, mx, my, c2, bx2, by2);
    src.prefetch(a2, c2, bx2, by2);
  }
  // Chroma samples stay in the scratch for the whole tile.  The luma samples only live from
  // here to the end of the emit: the TMA source lends the ring slot the raw block just came
  // out of for them (no luma scratch of its own: 2 KiB less shared memory per warp, which is
  // what lets a 16th warp onto the SM) and gets it back -- re-armed for the item two steps
  // ahead -- when the pixels are out.  Chroma items hand their slot back at once.
  u32x2 *dst;
  if (comp == 0) {
    dst = src.luma_scratch(sc.y);
    se.y = dst;
  } else {
    dst = comp == 1 ? sc.cb : sc.cr;
    src.release();
  }
#if defined(GIDK_NO_LO3)
  constexpr bool kLo3 = false;   // experiment: 8 KB less code in the 4:4:4 / grey kernels, 27 % more column-pass work on sparse blocks
#else
  constexpr bool kLo3 = H * V == 1;
#endif
  constexpr bool kTransposed = ROT == TURN_90 || ROT == TURN_270;  // (fast instances only: frame_setup.hpp)
  if (Q8) dequant_idct_sparse<true, kLo3, kTransposed>(active, raw, qw + 32 * comp, dst, sc.stride);
  else dequant_idct_exact(active, raw, qw + 32 * comp, dst, sc.stride);
  if (comp == 0) {
    if (active) emit_block_rows<H, V, GREY, RGBA, ROT>(f, se, bxi * 8, byi * 8, sx, sy);
    src.release();
  }
}

template <int H, int V, bool GREY, bool Q8, bool RGBA, int ROT, class Src>
GID_HD void mcu_run(bool active, const FrameDev &f, const uint32_t *qw /* [3][32] */, Src &src,
                    const Scratch &sc, int lane, int mx, int my) {
  constexpr int NI = GREY ? 1 : 2 + H * V;
#pragma unroll 1
  for (int it = 0; it < NI; it++) mcu_item<H, V, GREY, Q8, RGBA, ROT>(it, active, f, qw, src, sc, lane, mx, my);
}

}  // namespace gidk
