// entropy_kernel.cu -- baseline Huffman decoding on the device, one thread per restart interval.
//
// SURVEY.md section 8f, rank 1 (the step BEFORE the reconstruction path).  After RSTn the DC
// predictors are reset and the bit buffer is byte-aligned (Check_Restart,
// gid-decoding_jpg.adb:1133-1177), so the restart intervals of a baseline scan are independent
// bit streams: an 8192 x 8192 frame with DRI = 64 MCUs is 16 384 of them.  The host only finds
// the RSTn markers (a memchr pass over the scan's bytes) and sends the COMPRESSED bytes -- about
// 0.1 byte per pixel instead of the 6 bytes per pixel of dense planes -- and this kernel runs
// Decode_8x8_Block (:758-788, without the de-quantisation, which the reconstruction kernel
// does) for every block of its interval, storing the values at their natural position in the
// int16 planes that the reconstruction kernel reads.
//
// Bit-level behaviour follows the reference's bit buffer (:594-753): FF 00 is a stuffed FF
// (the host has cut the intervals at every other FF xx, so nothing else can follow an FF here),
// codes are canonical Huffman codes looked at 16 bits at a time (Next_Huffval, :738-746), values
// are extended by Bin_Twos_Complement (:748-753).  Whatever the reference would not decode
// silently -- a code that is in no table, a run past coefficient 63, an interval that needs more
// bits than it has or leaves whole bytes unused -- is NOT handled here: the thread raises a flag,
// gidcuda_job_wait reports GIDCUDA_ERR_DATA, and the caller decodes the scan on the host, where
// the error behaviour of the reference is reproduced exactly.
#include "recon_launch.h"

namespace gidk {

namespace {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Bits {
  const uint8_t *p, *e;
  uint64_t buf;  // left-aligned: the next bit is bit 63
  int len, pad;  // valid bits; how many of them (at the tail) are padding past the interval's end
  __device__ __forceinline__ void refill() {
    while (len <= 56) {
      uint32_t b = 0xFFu;
      if (p < e) {
        b = __ldg(p++);
        if (b == 0xFFu && p < e) p++;  // the stuffed 00
      } else {
        pad += 8;
      }
      buf |= (uint64_t)b << (56 - len);
      len += 8;
    }
  }
  __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(buf >> (64 - n)); }
  __device__ __forceinline__ void drop(int n) {
    buf <<= n;
    len -= n;
  }
};

// Next_Huffval: 10-bit direct table, then the canonical search for longer codes; -1 = no such code
__device__ __forceinline__ int symbol(Bits &b, const HuffDev &t) {
  const uint32_t look = b.peek(16);
  const uint32_t e = t.lut[look >> 6];
  if (e != 0) {
    b.drop((int)(e >> 8));
    return (int)(e & 0xFFu);
  }
  for (int l = 11; l <= 16; l++) {
    const int code = (int)(look >> (16 - l));
    if (code <= t.maxcode[l]) {
      b.drop(l);
      return t.vals[(t.valoff[l] + code) & 0xFF];
    }
  }
  return -1;
}

__device__ __forceinline__ int extend(uint32_t v, int n) {  // Bin_Twos_Complement
  return v < (1u << (n - 1)) ? (int)v + 1 - (1 << n) : (int)v;
}

// One block; false = something the host decoder has to look at
__device__ __forceinline__ bool decode_block(Bits &b, const HuffDev &dc, const HuffDev &ac, int &pred, int16_t *blk,
                                             const uint8_t *zz) {
  if (b.len <= 32) b.refill();
  const int s = symbol(b, dc);
  if (s < 0) return false;
  const int n = s & 15;
  if (n) {
    pred += extend(b.peek(n), n);
    b.drop(n);
  }
  blk[0] = (int16_t)pred;
  int coef = 0;
  for (;;) {
    if (b.len <= 32) b.refill();
    const int code = symbol(b, ac);
    if (code < 0) return false;
    if (code == 0) break;  // EOB
    const int m = code & 15;
    if (m == 0 && code != 0xF0) return false;
    coef += (code >> 4) + 1;
    if (coef > 63) return false;
    if (m) {
      blk[zz[coef]] = (int16_t)extend(b.peek(m), m);
      b.drop(m);
    }
    if (coef == 63) break;
  }
  return b.pad <= b.len;  // false: bits past the end of the interval were used
}

__global__ void __launch_bounds__(128) huffman_intervals_kernel(ScanArgs a) {
  __shared__ HuffDev tabs[6];
  __shared__ uint8_t zz[64];
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.tables);
    uint32_t *dst = reinterpret_cast<uint32_t *>(tabs);
    for (uint32_t i = threadIdx.x; i < sizeof(tabs) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  }
  __syncthreads();
  const uint32_t it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= a.nintervals) return;
  Bits b;
  b.p = a.data + __ldg(a.begin + it);
  b.e = a.data + __ldg(a.end + it);
  b.buf = 0;
  b.len = 0;
  b.pad = 0;
  int pred[3] = {0, 0, 0};
  const uint32_t m0 = it * a.ri;
  const uint32_t m1 = min(m0 + a.ri, a.nmcu);
  bool ok = true;
  for (uint32_t m = m0; m < m1 && ok; m++) {
    const uint32_t my = m / a.mcux, mx = m - my * a.mcux;
    for (int c = 0; c < a.ncomp && ok; c++)
      for (int sy = 0; sy < a.v[c] && ok; sy++)
        for (int sx = 0; sx < a.h[c] && ok; sx++) {
          int16_t *blk = a.plane[c] + ((size_t)(my * a.v[c] + sy) * a.bx[c] + (mx * a.h[c] + sx)) * 64;
          ok = decode_block(b, tabs[c], tabs[3 + c], pred[c], blk, zz);
        }
  }
  // every interval but the last ends where its RSTn begins: all bytes read, less than a byte
  // of real bits left after byte alignment (the serial decoder tolerates more: let it decide)
  if (ok && it != a.nintervals - 1) {
    b.drop(b.len & 7);
    const int real = b.len - min(b.pad, b.len);
    if (b.p < b.e || real >= 8) ok = false;
  }
  if (!ok) {
    atomicOr(a.status, 1u);
    atomicAdd(a.status + 1, 1u);
  }
}

}  // namespace

cudaError_t launch_huffman_intervals(const ScanArgs &a, cudaStream_t stream) {
  if (a.nintervals == 0) return cudaSuccess;
  huffman_intervals_kernel<<<(a.nintervals + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace gidk
