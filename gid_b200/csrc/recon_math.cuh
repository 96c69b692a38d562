// recon_math.cuh -- integer arithmetic of GID's JPEG reconstruction, written for
// one thread owning a whole 8x8 block in registers.
//
// Everything here must produce, bit for bit, what the reference computes in
//   Decode_8x8_Block (dequant)          gid-decoding_jpg.adb:771, :785, :1820-1822
//   Inverse_DCT_8x8_Block               gid-decoding_jpg.adb:790-907
//   Color_Transformation_and_Output     gid-decoding_jpg.adb:1018-1054
// The arithmetic is 32-bit two's-complement with wrap-around (GNAT Integer in
// the Fast_unchecked build) and Ada's "/" truncates toward zero, so every
// division below is a C "/" on a signed int, never a shift.
//
// The functions are __host__ __device__ so that tests/ can compile the same
// code for the CPU and compare it with the oracle without a GPU (test-only
// build; the product library contains device code only).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GID_HD __host__ __device__ __forceinline__
#else
#define GID_HD inline
#endif

namespace gidk {

// Chen-Wang constants, gid-decoding_jpg.adb:792-797
constexpr int W1 = 2841, W2 = 2676, W3 = 2408, W5 = 1609, W6 = 1108, W7 = 565;

// Wrap-around helpers (signed overflow is UB in C++; the reference wraps).
GID_HD int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
GID_HD int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
GID_HD int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
// a * b + c
GID_HD int wmad(int a, int b, int c) { return (int)((unsigned)a * (unsigned)b + (unsigned)c); }

GID_HD int clip255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

// Row_IDCT, gid-decoding_jpg.adb:799-845.
// The "all AC zero" shortcut (:810-812) returns blk(0)*8 for the 8 outputs.
// The general path with zero AC terms gives (blk(0)*2048 + 128) / 256, which
// differs for negative blk(0) only through the rounding constant; so the
// shortcut is reproduced by selecting the constant: 0 instead of 128.
// (Exact while |blk(0)| < 2**20, which every 8-bit JPEG stream satisfies:
// |DC| <= 2047 * 255.)
GID_HD void row_idct(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  int x0, x1, x2, x3, x4, x5, x6, x7, x8;
  x1 = wmul(b4, 2048);
  x2 = b6;
  x3 = b2;
  x4 = b1;
  x5 = b7;
  x6 = b5;
  x7 = b3;
  const int any_ac = x1 | x2 | x3 | x4 | x5 | x6 | x7;
  x0 = wmad(b0, 2048, any_ac != 0 ? 128 : 0);
  // odd part: x4' = x8 + (W1-W7)*x4 with x8 = W7*(x4+x5)  ==  W1*x4 + W7*x5 (mod 2**32), etc.
  const int o4 = wmad(W1, x4, wmul(W7, x5));
  const int o5 = wsub(wmul(W7, x4), wmul(W1, x5));
  const int o6 = wmad(W5, x6, wmul(W3, x7));
  const int o7 = wsub(wmul(W3, x6), wmul(W5, x7));
  x8 = wadd(x0, x1);
  x0 = wsub(x0, x1);
  // even part: x2' = W6*(x3+x2) - (W2+W6)*x2 == W6*x3 - W2*x2 ; x3' = W6*x2 + W2*x3
  const int e2 = wsub(wmul(W6, x3), wmul(W2, x2));
  const int e3 = wmad(W2, x3, wmul(W6, x2));
  x1 = wadd(o4, o6);
  x4 = wsub(o4, o6);
  x6 = wadd(o5, o7);
  x5 = wsub(o5, o7);
  x7 = wadd(x8, e3);
  x8 = wsub(x8, e3);
  x3 = wadd(x0, e2);
  x0 = wsub(x0, e2);
  x2 = wmad(181, wadd(x4, x5), 128) / 256;
  x4 = wmad(181, wsub(x4, x5), 128) / 256;
  b0 = wadd(x7, x1) / 256;
  b1 = wadd(x3, x2) / 256;
  b2 = wadd(x0, x4) / 256;
  b3 = wadd(x8, x6) / 256;
  b4 = wsub(x8, x6) / 256;
  b5 = wsub(x0, x4) / 256;
  b6 = wsub(x3, x2) / 256;
  b7 = wsub(x7, x1) / 256;
}

// Col_IDCT, gid-decoding_jpg.adb:847-895, general path only: the shortcut at
// :858-862, Clip((b0+32)/64+128), equals the general path when x1..x7 = 0
// ((b0*256+8192)/16384 is the same rational number).  Outputs are NOT clipped
// here: v = s / 2**14 + 128; the caller saturates to 0..255 (Clip, :755-756).
GID_HD void col_idct(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  int x0, x1, x2, x3, x4, x5, x6, x7, x8;
  x1 = wmul(b4, 256);
  x2 = b6;
  x3 = b2;
  x4 = b1;
  x5 = b7;
  x6 = b5;
  x7 = b3;
  x0 = wmad(b0, 256, 8192);
  // x4' = (W7*(x4+x5) + 4 + (W1-W7)*x4) / 8 == (W1*x4 + W7*x5 + 4) / 8, etc.
  const int o4 = wmad(W1, x4, wmad(W7, x5, 4)) / 8;
  const int o5 = wsub(wmad(W7, x4, 4), wmul(W1, x5)) / 8;
  const int o6 = wmad(W5, x6, wmad(W3, x7, 4)) / 8;
  const int o7 = wsub(wmad(W3, x6, 4), wmul(W5, x7)) / 8;
  x8 = wadd(x0, x1);
  x0 = wsub(x0, x1);
  const int e2 = wsub(wmad(W6, x3, 4), wmul(W2, x2)) / 8;
  const int e3 = wmad(W2, x3, wmad(W6, x2, 4)) / 8;
  x1 = wadd(o4, o6);
  x4 = wsub(o4, o6);
  x6 = wadd(o5, o7);
  x5 = wsub(o5, o7);
  x7 = wadd(x8, e3);
  x8 = wsub(x8, e3);
  x3 = wadd(x0, e2);
  x0 = wsub(x0, e2);
  x2 = wmad(181, wadd(x4, x5), 128) / 256;
  x4 = wmad(181, wsub(x4, x5), 128) / 256;
  b0 = wadd(wadd(x7, x1) / 16384, 128);
  b1 = wadd(wadd(x3, x2) / 16384, 128);
  b2 = wadd(wadd(x0, x4) / 16384, 128);
  b3 = wadd(wadd(x8, x6) / 16384, 128);
  b4 = wadd(wsub(x8, x6) / 16384, 128);
  b5 = wadd(wsub(x0, x4) / 16384, 128);
  b6 = wadd(wsub(x3, x2) / 16384, 128);
  b7 = wadd(wsub(x7, x1) / 16384, 128);
}

// Inverse_DCT_8x8_Block on 64 registers (natural order, index = 8*row + col).
// On return b[] holds the un-clipped samples.
GID_HD void idct_8x8(int (&b)[64]) {
#pragma unroll
  for (int r = 0; r < 8; r++)
    row_idct(b[8 * r + 0], b[8 * r + 1], b[8 * r + 2], b[8 * r + 3], b[8 * r + 4], b[8 * r + 5],
             b[8 * r + 6], b[8 * r + 7]);
#pragma unroll
  for (int c = 0; c < 8; c++)
    col_idct(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
}

// Chroma terms of the YCbCr arm, gid-decoding_jpg.adb:1028-1035.
//   R = Clip((Y*256 + 359*cr' + 128) / 256),  cr' = Cr - 128, cb' = Cb - 128
//   G = Clip((Y*256 -  88*cb' - 183*cr' + 128) / 256)
//   B = Clip((Y*256 + 454*cb' + 128) / 256)
// Y*256 is a multiple of 256, so floor((Y*256 + T) / 256) = Y + floor(T / 256),
// and truncation differs from floor only where the result is negative, which
// Clip maps to 0 either way.  Hence R = Clip(Y + (T_r >> 8)) etc. with an
// arithmetic shift.  (Checked exhaustively over all 2**24 inputs in tests.)
struct ChromaTerms { int r, g, b; };
GID_HD ChromaTerms chroma_terms(int cb, int cr) {
  ChromaTerms t;
  t.r = (359 * cr - 45824) >> 8;              // 359*(cr-128) + 128
  t.g = (-88 * cb - 183 * cr + 34816) >> 8;   // -88*(cb-128) - 183*(cr-128) + 128
  t.b = (454 * cb - 57984) >> 8;              // 454*(cb-128) + 128
  return t;
}

}  // namespace gidk
