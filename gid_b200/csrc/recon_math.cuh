// recon_math.cuh -- integer arithmetic of GID's JPEG reconstruction, written for
// one thread owning a whole 8x8 block in registers.
//
// Everything here must produce, bit for bit, what the reference computes in
//   Decode_8x8_Block (dequant)          gid-decoding_jpg.adb:771, :785, :1820-1822
//   Inverse_DCT_8x8_Block               gid-decoding_jpg.adb:790-907
//   Color_Transformation_and_Output     gid-decoding_jpg.adb:1018-1054
// The arithmetic is 32-bit two's-complement with wrap-around (GNAT Integer in
// the Fast_unchecked build) and Ada's "/" truncates toward zero.
//
// Instruction budget (B200, measured with tools/ubench.cu): every integer
// instruction class issues at 2 warp-instructions / clock / SM; IMAD (FMA pipe)
// and ALU-pipe instructions (SHF, LOP3, IADD3, LEA, VIMNMX, PRMT, I2IP) overlap
// up to the issue limit of 4 / clock / SM.  The code below is therefore written
// to (a) minimise the instruction count and (b) split it evenly over the two
// pipes: multiplications are re-associated into 2-term IMAD chains, sums feeding
// a division are 3-input adds, and the truncating division puts its sign
// correction on the FMA pipe (tdiv).
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
GID_HD int wmad(int a, int b, int c) { return (int)((unsigned)a * (unsigned)b + (unsigned)c); }
GID_HD int wadd3(int a, int b, int c) { return (int)((unsigned)a + (unsigned)b + (unsigned)c); }
GID_HD int wneg(int a) { return (int)(0u - (unsigned)a); }

GID_HD int clip255(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

// Truncating division by 2**K (Ada "/"):  x / 2**K = (x + (x < 0 ? 2**K - 1 : 0)) >> K.
// Device: sign bit (SHF, ALU pipe), bias by IMAD sign * (2**K - 1) + x (FMA pipe),
// arithmetic shift (ALU pipe).
template <int K>
GID_HD int tdiv(int x) {
#if defined(__CUDA_ARCH__)
  const unsigned s = (unsigned)x >> 31;
  int t;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(s), "r"((1 << K) - 1), "r"(x));
  return t >> K;
#else
  return x / (1 << K);
#endif
}

// ---- byte permute and byte / half-word dot products (PRMT, IDP.4A, IDP.2A) -----------
// The host versions restate the PTX semantics so that the CPU suite exercises
// the same formulas.
GID_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
#else
  // selector nibble: bits 2..0 = source byte of {b, a}; bit 3 = replicate that byte's sign bit
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t d = 0;
  for (int i = 0; i < 4; i++) {
    const uint32_t n = (sel >> (4 * i)) & 0xF;
    uint32_t byte = (uint32_t)((src >> (8 * (n & 7))) & 0xFF);
    if (n & 8) byte = (byte & 0x80) ? 0xFFu : 0u;
    d |= byte << (8 * i);
  }
  return d;
#endif
}
GID_HD int dp4a_uu(uint32_t a, uint32_t b, int c) {  // unsigned bytes x unsigned bytes
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  unsigned acc = (unsigned)c;
  for (int i = 0; i < 4; i++) acc += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
  return (int)acc;
#endif
}
GID_HD int dp4a_us(uint32_t a, uint32_t b, int c) {  // unsigned bytes x signed bytes
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  unsigned acc = (unsigned)c;
  for (int i = 0; i < 4; i++)
    acc += (unsigned)((int)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF));
  return (int)acc;
#endif
}
GID_HD int dp4a_ss(uint32_t a, uint32_t b, int c) {  // signed bytes x signed bytes
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  unsigned acc = (unsigned)c;
  for (int i = 0; i < 4; i++)
    acc += (unsigned)((int)(int8_t)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF));
  return (int)acc;
#endif
}
// a = two unsigned 16-bit lanes, b = four unsigned bytes: .lo = a.lo*b0 + a.hi*b1 + c, .hi = a.lo*b2 + a.hi*b3 + c
GID_HD int dp2a_lo_uu(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return (int)((a & 0xFFFFu) * (b & 0xFFu) + (a >> 16) * ((b >> 8) & 0xFFu) + (unsigned)c);
#endif
}
GID_HD int dp2a_hi_uu(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return (int)((a & 0xFFFFu) * ((b >> 16) & 0xFFu) + (a >> 16) * (b >> 24) + (unsigned)c);
#endif
}

// ---- truncating division, two values at a time ------------------------------------------
// The fused kernels are bound by the ALU pipe (shifts, permutes, packs: 2 warp-
// instructions / clock / SM), and tdiv spends two ALU instructions per value
// (sign, final shift) -- 208 values per block.  Pairing halves the sign work:
//   sp = PRMT(x1, x2, 0xFFBB): bytes 0,1 = 0xFF if x1 < 0 else 0, bytes 2,3 the same for x2
//        (selector bit 3 replicates the sign bit of the selected byte) -- ONE ALU instruction;
//   x + bias = a byte / half-word dot product of sp with a constant, accumulating
//        into x (IDP.4A / IDP.2A, FMA pipe):
//          K = 8 : bias 255   = u8 0xFF * 1
//          K = 3 : bias 7     = s8 (-1) * (-7)
//          K = 16: bias 65535 = u16 0xFFFF * 1
//   and the final arithmetic shift (SHF, or LEA.HI.SX32 when an add follows).
// Exact for every 32-bit x (the same wrap-around sum as tdiv).
GID_HD uint32_t sign_pair(int x1, int x2) { return prmt((uint32_t)x1, (uint32_t)x2, 0xFFBBu); }
template <int K> GID_HD int bias_first(uint32_t sp, int x) {
  static_assert(K == 3 || K == 8 || K == 16, "supported divisors");
  if (K == 8) return dp4a_uu(sp, 0x00000001u, x);
  if (K == 3) return dp4a_ss(sp, 0x000000F9u, x);
  return dp2a_lo_uu(sp, 0x00000001u, x);
}
template <int K> GID_HD int bias_second(uint32_t sp, int x) {
  if (K == 8) return dp4a_uu(sp, 0x00010000u, x);
  if (K == 3) return dp4a_ss(sp, 0x00F90000u, x);
  return dp2a_hi_uu(sp, 0x01000000u, x);
}
// final shift: SHF on the ALU pipe, or (MH) mul.hi by 2**(32-K) = floor(t / 2**K) on the
// FMA pipe (IMAD.HI, half rate) -- used where the ALU pipe is the busier one and the
// result feeds multiplications, so that there is no add for the shift to fuse with.
template <int K, bool MH> GID_HD int floor_shift(int t) {
#if defined(__CUDA_ARCH__)
  if (MH) {
    int r;
    asm("mul.hi.s32 %0, %1, %2;" : "=r"(r) : "r"(t), "r"((int)(1u << (32 - K))));
    return r;
  }
#endif
  return t >> K;
}
template <int K, bool MH = false>
GID_HD void tdiv2(int x1, int x2, int &r1, int &r2) {
  const uint32_t sp = sign_pair(x1, x2);
  r1 = floor_shift<K, MH>(bias_first<K>(sp, x1));
  r2 = floor_shift<K, MH>(bias_second<K>(sp, x2));
}

#if defined(GIDK_MULHI_ROWS)
constexpr bool kRowOutMulHi = true;
#else
constexpr bool kRowOutMulHi = false;
#endif

// Row_IDCT, gid-decoding_jpg.adb:799-845, with the rounding constant supplied by
// the caller: `rnd` = 128 when the row has a non-zero AC term, 0 when it has
// none.  The "all AC zero" shortcut (:810-812) returns blk(0)*8 for the 8
// outputs; the general path with zero AC terms gives (blk(0)*2048 + 128) / 256,
// which differs for negative blk(0) only through the rounding constant, so the
// shortcut is reproduced by selecting the constant.  (Exact while
// |blk(0)| < 2**20, which every 8-bit JPEG stream satisfies: |DC| <= 2047 * 255.)
//
// Re-association (all exact modulo 2**32):
//   x8 = W7*(x4+x5); x4' = x8 + (W1-W7)*x4 = W1*x4 + W7*x5;  x5' = x8 - (W1+W7)*x5 = W7*x4 - W1*x5
//   x8 = W3*(x6+x7); x6' = x8 - (W3-W5)*x6 = W5*x6 + W3*x7;  x7' = x8 - (W3+W5)*x7 = W3*x6 - W5*x7
//   x1 = W6*(x3+x2); x2' = x1 - (W2+W6)*x2 = W6*x3 - W2*x2;  x3' = x1 + (W2-W6)*x3 = W2*x3 + W6*x2
//   181*(x4+x5)+128 = 181*x4 + (181*x5 + 128)
GID_HD void row_idct_rnd(int rnd, int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  const int x1 = wmul(b4, 2048);
  const int x0 = wmad(b0, 2048, rnd);
  const int o4 = wmad(W1, b1, wmul(W7, b7)), o5 = wsub(wmul(W7, b1), wmul(W1, b7));
  const int o6 = wmad(W5, b5, wmul(W3, b3)), o7 = wsub(wmul(W3, b5), wmul(W5, b3));
  const int e2 = wsub(wmul(W6, b2), wmul(W2, b6)), e3 = wmad(W2, b2, wmul(W6, b6));
  const int A = wadd(x0, x1), B = wsub(x0, x1);
  const int p1 = wadd(o4, o6), q4 = wsub(o4, o6), p6 = wadd(o5, o7), q5 = wsub(o5, o7);
  const int h = wmad(181, q4, 128);
  int x2, x4;
  tdiv2<8>(wmad(181, q5, h), wmad(-181, q5, h), x2, x4);
  const int P = wadd(A, e3), M = wsub(A, e3), Q = wadd(B, e2), N = wsub(B, e2);
  tdiv2<8, kRowOutMulHi>(wadd(P, p1), wsub(P, p1), b0, b7);
  tdiv2<8, kRowOutMulHi>(wadd(M, p6), wsub(M, p6), b3, b4);
  tdiv2<8, kRowOutMulHi>(wadd(Q, x2), wsub(Q, x2), b1, b6);
  tdiv2<8, kRowOutMulHi>(wadd(N, x4), wsub(N, x4), b2, b5);
}
// Row_IDCT exactly as the reference decides it (:803-812): the shortcut is taken when x1 .. x7 are zero,
// where x1 = blk(4) * 2**11 -- a blk(4) that is a non-zero multiple of 2**21 counts as zero -- and its
// result is blk(0) * 8 (which the general path with rnd = 0 only equals while blk(0) * 2**11 does not wrap).
// One block per thread with its own branch: the generic kernels and the exact variant of the fused ones.
GID_HD bool row_has_ac(int b1, int b2, int b3, int b4, int b5, int b6, int b7) {
  return (b1 | b2 | b3 | wmul(b4, 2048) | b5 | b6 | b7) != 0;
}
GID_HD void row_idct(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  if (!row_has_ac(b1, b2, b3, b4, b5, b6, b7)) {
    b0 = b1 = b2 = b3 = b4 = b5 = b6 = b7 = wmul(b0, 8);
    return;
  }
  row_idct_rnd(128, b0, b1, b2, b3, b4, b5, b6, b7);
}

// Col_IDCT, gid-decoding_jpg.adb:847-895, general path only: the shortcut at
// :858-862, Clip((b0+32)/64+128), equals the general path when x1..x7 = 0
// ((b0*256+8192)/16384 is the same rational number).  Outputs are NOT clipped
// here: v = s / 2**14 + 128; the caller saturates to 0..255 (Clip, :755-756).
// K = 14 (the output stage) keeps the one-value form: its bias 16383 is not a byte or
// half-word multiple, and scaling the sums by 4 to use the K = 16 form would change
// the wrap-around behaviour of out-of-range streams.
// GIDK_TDIV14_PAIR (experiment): the two outputs of a butterfly share one sign extraction (PRMT) and
// take their bias 16383 = 255 * 64 + 63 from two byte dot products on the FMA pipe (u8 0xFF * 64,
// then s8 (-1) * (-63)): 1.5 ALU + 2 FMA instructions per output instead of 2 + 1.
GID_HD void tdiv14_pair_128(int x1, int x2, int &r1, int &r2) {
  const uint32_t sp = sign_pair(x1, x2);
  const int t1 = dp4a_ss(sp, 0x000000C1u, dp4a_uu(sp, 0x00000040u, x1));
  const int t2 = dp4a_ss(sp, 0x00C10000u, dp4a_uu(sp, 0x00400000u, x2));
  r1 = wadd(t1 >> 14, 128);
  r2 = wadd(t2 >> 14, 128);
}
GID_HD void col_outputs(int A, int B, int e2, int e3, int p1, int p6, int x2, int x4, int &b0, int &b1, int &b2,
                        int &b3, int &b4, int &b5, int &b6, int &b7) {
#if defined(GIDK_TDIV14_PAIR)
  const int P = wadd(A, e3), M = wsub(A, e3), Q = wadd(B, e2), N = wsub(B, e2);
  tdiv14_pair_128(wadd(P, p1), wsub(P, p1), b0, b7);
  tdiv14_pair_128(wadd(M, p6), wsub(M, p6), b3, b4);
  tdiv14_pair_128(wadd(Q, x2), wsub(Q, x2), b1, b6);
  tdiv14_pair_128(wadd(N, x4), wsub(N, x4), b2, b5);
  return;
#endif
  b0 = wadd(tdiv<14>(wadd3(A, e3, p1)), 128);
  b7 = wadd(tdiv<14>(wadd3(A, e3, wneg(p1))), 128);
  b3 = wadd(tdiv<14>(wadd3(A, wneg(e3), p6)), 128);
  b4 = wadd(tdiv<14>(wadd3(A, wneg(e3), wneg(p6))), 128);
  b1 = wadd(tdiv<14>(wadd3(B, e2, x2)), 128);
  b6 = wadd(tdiv<14>(wadd3(B, e2, wneg(x2))), 128);
  b2 = wadd(tdiv<14>(wadd3(B, wneg(e2), x4)), 128);
  b5 = wadd(tdiv<14>(wadd3(B, wneg(e2), wneg(x4))), 128);
}
GID_HD void col_idct(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  const int x1 = wmul(b4, 256);
  const int x0 = wmad(b0, 256, 8192);
  int o4, o5, o6, o7, e2, e3, x2, x4;
  tdiv2<3>(wmad(W1, b1, wmad(W7, b7, 4)), wsub(wmad(W7, b1, 4), wmul(W1, b7)), o4, o5);
  tdiv2<3>(wmad(W5, b5, wmad(W3, b3, 4)), wsub(wmad(W3, b5, 4), wmul(W5, b3)), o6, o7);
  tdiv2<3>(wsub(wmad(W6, b2, 4), wmul(W2, b6)), wmad(W2, b2, wmad(W6, b6, 4)), e2, e3);
  const int A = wadd(x0, x1), B = wsub(x0, x1);
  const int p1 = wadd(o4, o6), q4 = wsub(o4, o6), p6 = wadd(o5, o7), q5 = wsub(o5, o7);
  const int h = wmad(181, q4, 128);
  tdiv2<8>(wmad(181, q5, h), wmad(-181, q5, h), x2, x4);
  col_outputs(A, B, e2, e3, p1, p6, x2, x4, b0, b1, b2, b3, b4, b5, b6, b7);
}

// Col_IDCT with the reference's shortcut (:858-862) where it is NOT the general path: x1 .. x7 zero with
// x1 = blk(4 rows down) * 2**8, result Clip((b0 + 32) / 64 + 128).  It only differs from the general path
// when b0 * 256 + 8192 wraps (|b0| within 32 of 2**23: streams far outside any JPEG), so the fast kernels
// leave it out; the exact variant and the generic kernels carry it.
GID_HD void col_idct_exact(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  const bool shortcut = (b1 | b2 | b3 | wmul(b4, 256) | b5 | b6 | b7) == 0;
  const int v = wadd(tdiv<6>(wadd(b0, 32)), 128);
  col_idct(b0, b1, b2, b3, b4, b5, b6, b7);
  if (shortcut) b0 = b1 = b2 = b3 = b4 = b5 = b6 = b7 = v;
}

// Col_IDCT for a column whose inputs b3 .. b7 are zero (rows 3 .. 7 of the
// block are zero rows after the row pass).  col_idct with those zeros folded:
//   x1 = 0;  o6 = tdiv3(4) = 0;  o7 = tdiv3(4) = 0;  e2 = tdiv3(W6*b2 + 4);  e3 = tdiv3(W2*b2 + 4)
//   p1 = q4 = o4 = tdiv3(W1*b1 + 4);  p6 = q5 = o5 = tdiv3(W7*b1 + 4);  A = B = x0
// Writes all 8 outputs (un-clipped samples).
GID_HD void col_idct_lo3(int &b0, int &b1, int &b2, int &b3, int &b4, int &b5, int &b6, int &b7) {
  const int x0 = wmad(b0, 256, 8192);
  int o4, o5, e2, e3, x2, x4;
  tdiv2<3>(wmad(W1, b1, 4), wmad(W7, b1, 4), o4, o5);
  tdiv2<3>(wmad(W6, b2, 4), wmad(W2, b2, 4), e2, e3);
  const int h = wmad(181, o4, 128);
  tdiv2<8>(wmad(181, o5, h), wmad(-181, o5, h), x2, x4);
  col_outputs(x0, x0, e2, e3, o4, o5, x2, x4, b0, b1, b2, b3, b4, b5, b6, b7);
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
    col_idct_exact(b[c], b[8 + c], b[16 + c], b[24 + c], b[32 + c], b[40 + c], b[48 + c], b[56 + c]);
}

// Four samples -> four saturated bytes (Clip, :755-756, then packing), v0 in the
// low byte.  Device: two I2IP (cvt.pack.sat.u8.s32), which saturate and merge.
GID_HD uint32_t pack_sat4(int v0, int v1, int v2, int v3) {
#if defined(__CUDA_ARCH__)
  uint32_t t, w;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(v3), "r"(v2), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(v1), "r"(v0), "r"(t));
  return w;
#else
  return (uint32_t)clip255(v0) | ((uint32_t)clip255(v1) << 8) | ((uint32_t)clip255(v2) << 16) |
         ((uint32_t)clip255(v3) << 24);
#endif
}

// Clip to 0..255 in one instruction on the device (VIMNMX.RELU).
GID_HD int clip255_fast(int x) {
#if defined(__CUDA_ARCH__)
  int r;
  asm("min.relu.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(255));
  return r;
#else
  return clip255(x);
#endif
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

// The same terms from packed sample words: sample K (0..3) of cbw / crw.
//   v = [cb, cr, cb, cr];  359 = 255 + 104,  454 = 255 + 199,  -183 = -128 - 55
// The constants of the first and third output byte come from the caller
// (ColourOrder), so that B, G, R output costs nothing extra: .r is the term of
// the FIRST byte of the pixel and .b the term of the THIRD.
struct ColourOrder { uint32_t k_first; int bias_first; uint32_t k_third; int bias_third; };
GID_HD ColourOrder colour_order(bool bgr) {
  ColourOrder o;
  o.k_first = bgr ? 0x00C700FFu : 0x6800FF00u;   // (255, 0, 199, 0) : (0, 255, 0, 104)
  o.bias_first = bgr ? -57984 : -45824;
  o.k_third = bgr ? 0x6800FF00u : 0x00C700FFu;
  o.bias_third = bgr ? -45824 : -57984;
  return o;
}
template <int K>
GID_HD ChromaTerms chroma_terms_packed(const ColourOrder &o, uint32_t cbw, uint32_t crw) {
  const uint32_t v = prmt(cbw, crw, (uint32_t)(((4 + K) << 12) | (K << 8) | ((4 + K) << 4) | K));
  ChromaTerms t;
  t.r = dp4a_uu(v, o.k_first, o.bias_first) >> 8;
  t.g = dp4a_us(v, 0xC90080A8u, 34816) >> 8;    // (-88, -128, 0, -55)
  t.b = dp4a_uu(v, o.k_third, o.bias_third) >> 8;
  return t;
}

}  // namespace gidk
