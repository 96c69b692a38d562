// entropy_kernel.cu -- baseline Huffman decoding on the device (gidcuda_job_scan):
//   part 1  scans with restart intervals: one thread per interval      (huffman_intervals_kernel)
//   part 2  scans without restart markers: self-synchronising parallel decoding
//           (huffman_sync_kernel x 2, huffman_scan_kernel, huffman_output_kernel)
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

#include <cstdlib>

// -DGIDK_CHECKED: device-side assertions on every coefficient address and table index (a debug
// build for the GPU test suite; compute-sanitizer is not available on the test pool).
#if defined(GIDK_CHECKED)
#include <cassert>
#define GIDK_ASSERT(c) assert(c)
#else
#define GIDK_ASSERT(c) ((void)0)
#endif

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
  // The four bytes at p, most significant first, fetched ONE REFILL AHEAD (valid when `have`): a refill
  // consumes bytes whose loads were issued 32 bits of decoding earlier, so their latency is not on the
  // decoder's dependent chain (it was 29 % of the kernel's stall samples, profiles/r02_ncu_summary.md).
  // (kept as the four bytes the loads return: combined where they are used, so that nothing waits for them here)
  uint32_t a0, a1, a2, a3;
  bool have;
  __device__ __forceinline__ void fetch() {
    have = p + 4 <= e;
    if (have) {
      a0 = __ldg(p);
      a1 = __ldg(p + 1);
      a2 = __ldg(p + 2);
      a3 = __ldg(p + 3);
    }
  }
  __device__ __forceinline__ void start(const uint8_t *begin, const uint8_t *end) {
    p = begin;
    e = end;
    buf = 0;
    len = 0;
    pad = 0;
    fetch();
  }
  __device__ __forceinline__ void refill() {
    // four bytes at once where none of them is FF (byte stuffing, :640-662, makes the byte-wise loop below
    // a chain of dependent loads: one load latency per byte)
    if (len <= 32 && have) {
      const uint32_t w = (a0 << 24) | (a1 << 16) | (a2 << 8) | a3, n = ~w;
      if (((n - 0x01010101u) & w & 0x80808080u) == 0u) {  // no byte of n is zero: no FF among the four
        buf |= (uint64_t)w << (32 - len);
        len += 32;
        p += 4;
        fetch();
        return;
      }
    }
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
    fetch();
  }
  __device__ __forceinline__ void drop(int n) {
    buf <<= n;
    len -= n;
  }
};

__device__ __forceinline__ int extend(uint32_t v, int n) {  // Bin_Twos_Complement
  return v < (1u << (n - 1)) ? (int)v + 1 - (1 << n) : (int)v;
}

// One block; false = something the host decoder has to look at
// |v| < 2**bits  (v = -2**bits counts as outside)
__device__ __forceinline__ bool within_bits(int v, int bits) { return ((uint32_t)(v ^ (v >> 31)) >> bits) == 0u; }

// The interval kernel reads its tables through addresses in the shared window (a reference to a table in
// shared memory made the compiler re-derive the window's base -- S2R + LEA -- at every symbol).
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_s16(uint32_t a, int v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((short)v) : "memory");
}
constexpr uint32_t kOffMaxcode = 2048, kOffValoff = 2048 + 68, kOffVals = 2048 + 136;
static_assert(sizeof(HuffDev) == kOffVals + 256, "HuffDev layout");

// Next_Huffval on the top 32 bits of the bit buffer (`hi`): the symbol and the length of its code, -1 = no such code
__device__ __forceinline__ int symbol_in(uint32_t hi, uint32_t tab, int &codelen) {
  const uint32_t look = hi >> 16;
  const uint32_t e = lds_u16(tab + 2u * (look >> 6));
  if (e != 0) {
    codelen = (int)(e >> 8);
    return (int)(e & 0xFFu);
  }
  for (int l = 11; l <= 16; l++) {
    const int code = (int)(look >> (16 - l));
    if (code <= lds_s32(tab + kOffMaxcode + 4u * l)) {
      codelen = l;
      return (int)lds_u8(tab + kOffVals + (uint32_t)((lds_s32(tab + kOffValoff + 4u * l) + code) & 0xFF));
    }
  }
  codelen = 0;
  return -1;
}

// A code and the value bits behind it are at most 16 + 15 bits: both come out of one 32-bit window of the
// buffer, which is then advanced once.  (dc, ac, zz, blk: shared-window addresses; blk = the thread's staging block)
__device__ __forceinline__ bool decode_block(Bits &b, uint32_t dc, uint32_t ac, int &pred, uint32_t blk, uint32_t zz,
                                             int limbits) {
  if (b.len <= 32) b.refill();
  uint32_t hi = (uint32_t)(b.buf >> 32);
  int cl;
  const int s = symbol_in(hi, dc, cl);
  if (s < 0) return false;
  const int n = s & 15;
  if (n) {
    pred += extend((hi << cl) >> (32 - n), n);
    // the reference keeps a 32-bit predictor (:766-771): one that leaves the range of the planes (or of
    // the fast reconstruction instances) is the host decoder's case
    if (!within_bits(pred, limbits)) return false;
  }
  b.drop(cl + n);
  sts_s16(blk, pred);
  int coef = 0;
  for (;;) {
    if (b.len <= 32) b.refill();
    hi = (uint32_t)(b.buf >> 32);
    const int code = symbol_in(hi, ac, cl);
    if (code < 0) return false;
    if (code == 0) {  // EOB
      b.drop(cl);
      break;
    }
    const int m = code & 15;
    if (m == 0 && code != 0xF0) return false;
    coef += (code >> 4) + 1;
    if (coef > 63) return false;
    if (m > limbits) return false;  // a value of category m is below 2**m
    if (m) {
      GIDK_ASSERT(coef >= 1 && coef <= 63);
      sts_s16(blk + 2u * lds_u8(zz + (uint32_t)coef), extend((hi << cl) >> (32 - m), m));
    }
    b.drop(cl + m);
    if (coef == 63) break;
  }
  return b.pad <= b.len;  // false: bits past the end of the interval were used
}

// Every thread assembles its block in shared memory and writes it out whole, eight 16-byte stores: the
// planes need no zeroing beforehand (every block of the frame belongs to exactly one interval) and no
// sector is written piecemeal (scattered int16 stores into zeroed planes cost a 134 MB memset and 85 MB of
// read-modify-write DRAM traffic per 8192 x 8192 frame: profiles/r01f_ncu_summary.md).  A thread's buffer is
// 144 bytes from the next one's: its 16-byte pieces fall on different banks for the 8 threads of a
// quarter warp.
constexpr int kStageStride = 72;  // int16 per thread (64 used)

// Intervals per warp (a.lanes, a power of two): a warp runs the union of its lanes' paths through the decoder
// for as long as its longest interval, and a frame has far fewer intervals than the GPU has thread slots -- so
// the intervals are spread over as many warps as the SMs hold (8 lanes per warp for the 16 384 intervals of an
// 8192 x 8192 grey frame, one lane for a frame with an interval per MCU row) instead of filling few warps.
__global__ void __launch_bounds__(128) huffman_intervals_kernel(ScanArgs a) {
  __shared__ HuffDev tabs[6];
  __shared__ uint8_t zz[64];
  __shared__ __align__(16) int16_t stage[128 * kStageStride];
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a.tables);
    uint32_t *dst = reinterpret_cast<uint32_t *>(tabs);
    for (uint32_t i = threadIdx.x; i < sizeof(tabs) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t it = warp * a.lanes + lane;
  if (lane >= a.lanes || it >= a.nintervals) return;
  uint4 *mine = reinterpret_cast<uint4 *>(stage + threadIdx.x * kStageStride);
  uint32_t tabs_s = (uint32_t)__cvta_generic_to_shared(tabs), zz_s = (uint32_t)__cvta_generic_to_shared(zz);
  uint32_t mine_s = (uint32_t)__cvta_generic_to_shared(mine);
  // (made opaque: otherwise the compiler re-derives the shared window's base -- S2UR SR_CgaCtaId + LEA -- at every
  // symbol instead of keeping three registers)
  asm volatile("" : "+r"(tabs_s), "+r"(zz_s), "+r"(mine_s));
  Bits b;
  b.start(a.data + __ldg(a.begin + it), a.data + __ldg(a.end + it));
  int pred[3] = {0, 0, 0};  // (registers: only indexed by the unrolled component loop)
  const uint32_t m0 = it * a.ri;
  const uint32_t m1 = min(m0 + a.ri, a.nmcu);
  bool ok = true;
  for (uint32_t m = m0; m < m1 && ok; m++) {
    const uint32_t my = m / a.mcux, mx = m - my * a.mcux;
#pragma unroll
    for (int c = 0; c < 3; c++)  // (unrolled: a run-time index into the kernel parameters would put them in local memory)
      for (int sy = 0; c < a.ncomp && sy < a.v[c] && ok; sy++)
        for (int sx = 0; sx < a.h[c] && ok; sx++) {
          const size_t blockno = (size_t)(my * a.v[c] + sy) * a.bx[c] + (mx * a.h[c] + sx);
          GIDK_ASSERT(blockno < (size_t)a.bx[c] * a.v[c] * ((a.nmcu + a.mcux - 1) / a.mcux));
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int j = 0; j < 8; j++) mine[j] = z;
          ok = decode_block(b, tabs_s + (uint32_t)(c * sizeof(HuffDev)), tabs_s + (uint32_t)((3 + c) * sizeof(HuffDev)),
                            pred[c], mine_s, zz_s, a.limbits[c]);
          uint4 *dst = reinterpret_cast<uint4 *>(a.plane[c] + blockno * 64);
#pragma unroll
          for (int j = 0; j < 8; j++) dst[j] = mine[j];
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

cudaError_t launch_huffman_intervals(const ScanArgs &args, int num_sms, cudaStream_t stream) {
  if (args.nintervals == 0) return cudaSuccess;
  ScanArgs a = args;
  // as few lanes per warp as keep the warps within what the SMs hold at once (16 per SM)
  a.lanes = 1;
  while (a.lanes < 32u && (a.nintervals + a.lanes - 1) / a.lanes > (uint32_t)num_sms * 16u) a.lanes *= 2;
  static const int forced = [] { const char *e = getenv("GIDCUDA_INTERVAL_LANES"); return e ? atoi(e) : 0; }();  // (tuning)
  if (forced >= 1 && forced <= 32) a.lanes = (uint32_t)forced;
  const uint32_t nwarps = (a.nintervals + a.lanes - 1) / a.lanes;
  huffman_intervals_kernel<<<(nwarps + 3) / 4, 128, 0, stream>>>(a);
  return cudaGetLastError();
}


// =================================================================================================
// Baseline scans WITHOUT restart markers: self-synchronising parallel Huffman decoding.
//
// A scan without RSTn is one long bit stream, but Huffman codes re-synchronise: a decoder started
// at a wrong bit position or in a wrong state falls into step with the true code-word sequence
// after a few code words.  (Klein & Wiseman 2003; Weissenberger & Schmidt, "Massively Parallel
// Huffman Decoding on GPUs", ICPP 2018, and their JPEG follow-up.)  The stream, with the byte
// stuffing removed by the host, is cut into SUBSEQUENCES of kSubBits bits, one thread each.
//   state      = (bit position of the next code word, block index inside the MCU, position inside
//                the block: 0 = the DC code comes next) -- everything Decode_8x8_Block (:758-788)
//                and the MCU loop (:1187-1219) need to know besides the predictors;
//   exit[i]    = the state in which the decoder leaves subsequence i (first code word that starts at
//                or after bit (i + 1) * kSubBits), given the state entry[i] it entered with;
//   1. sync    every thread decodes its subsequence from a guessed state, then, window by window
//              (256 subsequences in shared memory), re-decodes from its predecessor's exit state
//              until nothing changes; a second launch with the windows shifted by half repairs the
//              window heads.  Subsequence 0 starts from the true state, so a chain that has
//              stopped changing is the true one;
//   2. scan    exclusive prefix sums of (blocks completed, sum of DC differences per component)
//              give every subsequence its first block number and its DC predictors;
//   3. output  every thread decodes once more, storing coefficients; it VERIFIES that it entered in
//              its predecessor's exit state and leaves in its own recorded exit state.  By
//              induction from subsequence 0 a verified chain is exactly the serial decode.  Any
//              mismatch, any code the serial decoder would raise on, or a block count that does not
//              reach the frame's: flag -> GIDCUDA_ERR_DATA -> the host decodes the scan.

namespace {

constexpr int kWindow = 256;
static_assert(kWindow == (int)kPscanWindow, "the shim sizes the sync launches of progressive scans with kPscanWindow");
// Re-decoding rounds per window.  A chain normally settles in 2 - 5 rounds.  Where it does not -- a
// flat region (periodic bit stream: a decoder in the wrong phase stays there), or full blocks without
// EOB at quality 100 (a wrong position inside the block survives until an EOB turns up) -- truth only
// advances one subsequence per round, which is a serial walk, and a host core does that an order of
// magnitude faster than a GPU thread: after this many rounds the window gives up and leaves the
// stretch to the chain repair in gidcuda_job_wait (measured: 256 rounds cost 3 ms per launch on a
// black frame, for nothing).
constexpr int kMaxRounds = 32;

struct RawBits {  // no stuffing, no end: the buffer is padded with FF bytes past the data
  const uint32_t *w;  // next 32-bit word of the (4-byte aligned) stream
  uint64_t buf;       // left-aligned: the next bit is bit 63
  int len;
  static __device__ __forceinline__ uint32_t be(uint32_t x) { return __byte_perm(x, 0u, 0x0123); }
  __device__ __forceinline__ void start(const uint8_t *data, uint32_t bitpos) {
    w = reinterpret_cast<const uint32_t *>(data) + (bitpos >> 5);
    const uint64_t two = ((uint64_t)be(__ldg(w)) << 32) | be(__ldg(w + 1));
    w += 2;
    buf = two << (bitpos & 31u);
    len = 64 - (int)(bitpos & 31u);
  }
  __device__ __forceinline__ void refill() {  // call when len <= 32
    GIDK_ASSERT(len >= 0 && len <= 32);
    buf |= (uint64_t)be(__ldg(w++)) << (32 - len);
    len += 32;
  }
  __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(buf >> (64 - n)); }
  __device__ __forceinline__ void drop(int n) {
    buf <<= n;
    len -= n;
  }
};

__device__ __forceinline__ uint64_t pack_state(uint32_t pos, uint32_t c, uint32_t zpos) {
  return (uint64_t)pos | ((uint64_t)c << 32) | ((uint64_t)zpos << 40);
}

// Where the blocks of one MCU go (shared memory: indexing the kernel parameters with a run-time
// index would make the compiler copy them to local memory).
struct BlockSlot {
  int16_t *plane;
  uint32_t row_blocks;  // blocks per block row of the component's plane
  uint32_t h, v, sx, sy, comp;
  int32_t limbits;  // WholeArgs::limbits of the component
};

// Decodes the code words of subsequence `i` that start before bit `limit`, from state `entry`.
// FINAL = false: returns the exit state, counts completed blocks and sums DC differences (tup).
// FINAL = true : also stores the coefficients; `b` = number of the first block (in scan order),
//                pred = predictors on entry; stops at block `total`; *bad is set for anything
//                the serial decoder would not decode silently.
template <bool FINAL>
__device__ __forceinline__ uint64_t run_subsequence(const WholeArgs &a, const HuffDev *tabs, const uint8_t *zz,
                                                    const BlockSlot *slots,
                                                    uint64_t entry, uint32_t limit, int4 &tup, uint32_t b, int pred0,
                                                    int pred1, int pred2, bool *bad, uint32_t *done_blocks) {
  uint32_t pos = (uint32_t)entry, c = (uint32_t)(entry >> 32) & 0xFFu, zpos = (uint32_t)(entry >> 40) & 0xFFu;
  RawBits bits;
  bits.start(a.data, pos);
  int nb = 0, dcs0 = 0, dcs1 = 0, dcs2 = 0;  // (scalars: an indexed array would live in local memory)
  int16_t *blk = nullptr;
  auto block_ptr = [&](uint32_t blockno, uint32_t cc) -> int16_t * {
    const uint32_t mcu = (blockno - cc) / a.bpm;
    const uint32_t my = mcu / a.mcux, mx = mcu - my * a.mcux;
    GIDK_ASSERT(cc < a.bpm && blockno >= cc && blockno < a.total_blocks);
    const BlockSlot &q = slots[cc];
    const size_t at = (size_t)(my * q.v + q.sy) * q.row_blocks + (mx * q.h + q.sx);
    GIDK_ASSERT(mx < a.mcux && at < (size_t)q.row_blocks * q.v * (a.total_blocks / a.bpm / a.mcux));
    return q.plane + at * 64;
  };
  if (FINAL) blk = block_ptr(b, c);
  while (pos < limit) {
    if (FINAL && b >= a.total_blocks) break;
    if (bits.len <= 32) bits.refill();
    GIDK_ASSERT(c < a.bpm && zpos < 64 && pos < a.nbits);
    const int k = (int)slots[c].comp;
    const HuffDev &t = zpos == 0 ? tabs[k] : tabs[3 + k];
    const uint32_t look = bits.peek(16);
    const uint32_t e = t.lut[look >> 6];
    int l = (int)(e >> 8), sym = (int)(e & 0xFFu);
    if (e == 0) {
      sym = -1;
      for (l = 11; l <= 16; l++) {
        const int code = (int)(look >> (16 - l));
        if (code <= t.maxcode[l]) {
          sym = t.vals[(t.valoff[l] + code) & 0xFF];
          break;
        }
      }
      if (sym < 0) {  // no such code: the serial decoder raises; a guessing decoder moves on one bit
        if (FINAL) { *bad = true; break; }
        bits.drop(1);
        pos += 1;
        continue;
      }
    }
    bits.drop(l);
    pos += (uint32_t)l;
    const int n = sym & 15;
    bool block_done = false;
    if (zpos == 0) {
      int v = 0;
      if (n) {
        v = extend(bits.peek(n), n);
        bits.drop(n);
        pos += (uint32_t)n;
      }
      if (FINAL) {
        pred0 += k == 0 ? v : 0;
        pred1 += k == 1 ? v : 0;
        pred2 += k == 2 ? v : 0;
        const int pk = k == 0 ? pred0 : (k == 1 ? pred1 : pred2);
        // (the reference's predictor has 32 bits, :766-771; out of range: the host decoder's case)
        if (!within_bits(pk, slots[c].limbits)) { *bad = true; break; }
        blk[0] = (int16_t)pk;
      } else {
        dcs0 += k == 0 ? v : 0;
        dcs1 += k == 1 ? v : 0;
        dcs2 += k == 2 ? v : 0;
      }
      zpos = 1;
    } else if (sym == 0) {
      block_done = true;  // EOB
    } else {
      const uint32_t idx = zpos + (uint32_t)(sym >> 4);
      if (FINAL && ((n == 0 && sym != 0xF0) || idx > 63 || n > (int)slots[c].limbits)) { *bad = true; break; }
      if (n) {
        const int v = extend(bits.peek(n), n);
        bits.drop(n);
        pos += (uint32_t)n;
        GIDK_ASSERT(!FINAL || (idx >= 1 && idx <= 63));
        if (FINAL) blk[zz[idx & 63u]] = (int16_t)v;
      }
      if (idx >= 63) block_done = true;
      else zpos = idx + 1;
    }
    if (block_done) {
      zpos = 0;
      c = c + 1 == a.bpm ? 0 : c + 1;
      nb++;
      if (FINAL) {
        b++;
        if (b < a.total_blocks) blk = block_ptr(b, c);
      }
    }
  }
  tup = make_int4(nb, dcs0, dcs1, dcs2);
  if (FINAL) *done_blocks = (uint32_t)nb;
  return pack_state(pos, c, zpos);
}

__device__ __forceinline__ void load_tables(const WholeArgs &a, HuffDev *tabs, uint8_t *zz, BlockSlot *slots) {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(a.tables);
  uint32_t *dst = reinterpret_cast<uint32_t *>(tabs);
  for (uint32_t i = threadIdx.x; i < 6 * sizeof(HuffDev) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  if (threadIdx.x == 0) {
    uint32_t cc = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)  // static component index; the MCU's blocks in scan order (:1190-1206)
      for (int sy = 0; sy < a.v[k]; sy++)
        for (int sx = 0; sx < a.h[k]; sx++, cc++) {
          BlockSlot q;
          q.plane = a.plane[k];
          q.row_blocks = a.bx[k];
          q.h = (uint32_t)a.h[k];
          q.v = (uint32_t)a.v[k];
          q.sx = (uint32_t)sx;
          q.sy = (uint32_t)sy;
          q.comp = (uint32_t)k;
          q.limbits = a.limbits[k];
          if (cc < 12) slots[cc] = q;
        }
  }
}

// Step 1.  Launch A: shift = 0, first = 1 (guessed entries).  Launch B: shift = kWindow / 2, first = 0.
__global__ void __launch_bounds__(kWindow) huffman_sync_kernel(WholeArgs a, uint32_t shift, int first) {
  __shared__ HuffDev tabs[6];
  __shared__ uint8_t zz[64];
  __shared__ uint64_t X[kWindow];
  __shared__ BlockSlot slots[12];
  load_tables(a, tabs, zz, slots);
  const uint32_t t = threadIdx.x;
  const uint32_t i = blockIdx.x * kWindow + t + shift;
  const bool active = i < a.nsub;
  const uint64_t true_start = pack_state(0u, 0u, 0u);
  uint64_t entry = 0, exit_ = 0, head = 0;
  int4 tup = make_int4(0, 0, 0, 0);
  bool dirty = false;
  if (first) {
    // the planes the output pass will store into, and the status words (both used only by later launches)
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t k = blockIdx.x * kWindow + t; k < a.zero_count; k += gridDim.x * kWindow) a.zero_base[k] = z;
    if (blockIdx.x == 0 && t < 4) a.status[t] = 0u;
  }
  __syncthreads();
  const uint32_t limit = active ? min((i + 1) * a.sub_bits, a.nbits) : 0u;
  if (active) {
    if (first) {
      entry = i == 0 ? true_start : pack_state(i * a.sub_bits, 0u, 0u);
      exit_ = run_subsequence<false>(a, tabs, zz, slots, entry, limit, tup, 0u, 0, 0, 0, nullptr, nullptr);
      dirty = true;
    } else {
      entry = a.entry[i];
      exit_ = a.exit_[i];
    }
    if (t == 0) head = i == 0 ? true_start : (first ? entry : a.exit_[i - 1]);
  }
  X[t] = exit_;
  for (int iter = 0; iter < kMaxRounds; iter++) {
    __syncthreads();
    const uint64_t want = t == 0 ? head : X[t - 1];
    const bool redo = active && want != entry;
    uint64_t nx = exit_;
    if (redo) {
      entry = want;
      nx = run_subsequence<false>(a, tabs, zz, slots, entry, limit, tup, 0u, 0, 0, 0, nullptr, nullptr);
      dirty = true;
    }
    __syncthreads();  // every X[t - 1] has been read
    const bool changed = redo && nx != exit_;
    if (redo) {
      exit_ = nx;
      X[t] = nx;
    }
    if (!__syncthreads_or(changed ? 1 : 0)) break;
  }
  if (active && dirty) {
    a.entry[i] = entry;
    a.exit_[i] = exit_;
    a.tuple[i] = tup;
  }
}

// Step 2.  One CTA: exclusive prefix sums of the tuples, 4096 at a time (four per thread).
__device__ __forceinline__ int4 add4(int4 a, int4 b) { return make_int4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ int4 sub4(int4 a, int4 b) { return make_int4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ int4 warp_inclusive(int4 s, uint32_t lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int x = __shfl_up_sync(0xFFFFFFFFu, s.x, d), y = __shfl_up_sync(0xFFFFFFFFu, s.y, d);
    const int z = __shfl_up_sync(0xFFFFFFFFu, s.z, d), q = __shfl_up_sync(0xFFFFFFFFu, s.w, d);
    if (lane >= (uint32_t)d) { s.x += x; s.y += y; s.z += z; s.w += q; }
  }
  return s;
}
__global__ void __launch_bounds__(1024) huffman_scan_kernel(WholeArgs a) {
  __shared__ int4 warp_tot[32];
  __shared__ int4 carry_s;
  const uint32_t t = threadIdx.x, lane = t & 31u, w = t >> 5;
  const int4 zero = make_int4(0, 0, 0, 0);
  if (t == 0) carry_s = zero;
  __syncthreads();
  for (uint32_t base = 0; base < a.nsub; base += 4096) {
    const uint32_t i0 = base + 4 * t;
    int4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = i0 + k < a.nsub ? a.tuple[i0 + k] : zero;
    const int4 mine = add4(add4(v[0], v[1]), add4(v[2], v[3]));
    const int4 s = warp_inclusive(mine, lane);
    if (lane == 31) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
      const int4 u = warp_tot[lane];
      warp_tot[lane] = sub4(warp_inclusive(u, lane), u);  // exclusive over the warps
    }
    __syncthreads();
    int4 ex = add4(add4(carry_s, warp_tot[w]), sub4(s, mine));  // everything before element i0
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < a.nsub) a.prefix[i0 + k] = ex;
      ex = add4(ex, v[k]);
    }
    __syncthreads();
    if (t == 1023) carry_s = ex;
    __syncthreads();
  }
}

// Step 3.
__global__ void __launch_bounds__(128) huffman_output_kernel(WholeArgs a) {
  __shared__ HuffDev tabs[6];
  __shared__ uint8_t zz[64];
  __shared__ BlockSlot slots[12];
  load_tables(a, tabs, zz, slots);
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nsub) return;
  const int4 pre = a.prefix[i];
  const uint32_t b = (uint32_t)pre.x;
  if (b >= a.total_blocks) return;  // padding after the last block
  const uint64_t entry = i == 0 ? pack_state(0u, 0u, 0u) : a.exit_[i - 1];
  bool chain_bad = entry != a.entry[i];  // the tuples (and so the prefixes) belong to another chain
  const uint32_t c = (uint32_t)(entry >> 32) & 0xFFu;
  if (c >= a.bpm || b % a.bpm != c) chain_bad = true;
  bool code_bad = false;
  uint32_t done = 0;
  if (!chain_bad) {
    int4 tup;
    const uint32_t limit = min((i + 1) * a.sub_bits, a.nbits);
    const uint64_t out = run_subsequence<true>(a, tabs, zz, slots, entry, limit, tup, b, pre.y, pre.z, pre.w, &code_bad, &done);
    // a subsequence that ends with the frame's last block stops early: nothing follows it
    if (!code_bad && b + done < a.total_blocks && out != a.exit_[i]) chain_bad = true;
    // ... and tells where the scan's last code word ended (1 + bit position): past the scan's last bit it was
    // fed from the padding, which is not what follows the scan in the file
    if (!code_bad && done != 0 && b + done == a.total_blocks) a.status[3] = (uint32_t)out + 1u;
  }
  // 2: the chain of states does not hold here (not settled: the host can repair it, see
  // gidcuda_job_wait); 8: something the serial decoder would raise on
  if (chain_bad) atomicOr(a.status, 2u);
  if (code_bad) atomicOr(a.status, 8u);
  if (done) atomicAdd(a.status + 2, done);
}


// =================================================================================================
// Progressive scans (ITU T.81 annex G; GID: gid-decoding_jpg.adb:1243-1747).
//
// FIRST scans are ordinary Huffman streams and take the self-synchronising decoder above with a
// smaller state:
//   DC first (Ss = 0, Ah = 0; :1263-1340): per block one DC code and its difference bits; state =
//       (bit position, block of the MCU); value stored = predictor << Al; prefix sums of the
//       differences give the predictors, as for baseline scans;
//   AC first (Ss > 0, Ah = 0; :1381-1470): one component, band Ss .. Se; state = (bit position,
//       position in the band).  EOBn (run < 15, size 0) ends the block and the next 2**n - 1 + bits
//       blocks read nothing, so a run only advances the block count -- it is not part of the state
//       at a code-word boundary.
// DC REFINEMENT scans (:1341-1380) are one bit per block: no code words at all.
// AC REFINEMENT scans (:1560-1690) are NOT decodable this way: inside an EOB run a block still reads one
// correction bit per coefficient that is already non-zero, so the decoder's state includes the
// absolute block number, which no guessed start can recover from the bits.  They stay on the host,
// which only needs the non-zero history (pscan_masks_kernel) and sends back, per block, which
// positions took a correction bit and which became +-(1 << Al) (pscan_apply_kernel).
// Block order: interleaved scans walk the frame's MCUs (:1187-1219); a single-component scan walks
// the component's OWN block grid, ceil(W_c / 8) x ceil(H_c / 8), row by row (:1952-1971).

struct PSlot {  // one block of the MCU of an interleaved scan
  int16_t *plane;
  uint32_t row_blocks, h, v, sx, sy, comp;
  int32_t limbits;
};

__device__ __forceinline__ void pscan_slots(const PscanArgs &a, PSlot *slots) {  // one thread
  uint32_t cc = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {  // static component index (see load_tables)
    bool in_scan = false;
    for (uint32_t j = 0; j < a.ncomp; j++) in_scan = in_scan || a.comps[j] == (uint32_t)k;
    if (!in_scan) continue;
    const int hh = a.ncomp > 1 ? a.h[k] : 1, vv = a.ncomp > 1 ? a.v[k] : 1;
    for (int sy = 0; sy < vv; sy++)
      for (int sx = 0; sx < hh; sx++, cc++) {
        PSlot q;
        q.plane = a.plane[k];
        q.row_blocks = a.row_blocks[k];
        q.h = (uint32_t)hh;
        q.v = (uint32_t)vv;
        q.sx = (uint32_t)sx;
        q.sy = (uint32_t)sy;
        q.comp = (uint32_t)k;
        q.limbits = a.limbits[k];
        if (cc < 12) slots[cc] = q;
      }
  }
}
__device__ __forceinline__ void pscan_load(const PscanArgs &a, HuffDev *tabs, uint8_t *zz, PSlot *slots) {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(a.tables);
  uint32_t *dst = reinterpret_cast<uint32_t *>(tabs);
  for (uint32_t i = threadIdx.x; i < 6 * sizeof(HuffDev) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  if (threadIdx.x == 0) pscan_slots(a, slots);
}

// block number `b` of the scan (its block `c` of the MCU) -> the block in its plane
__device__ __forceinline__ int16_t *pscan_block(const PscanArgs &a, const PSlot *slots, uint32_t b, uint32_t c) {
  const PSlot &q = slots[c];
  size_t at;
  if (a.ncomp > 1) {
    const uint32_t mcu = (b - c) / a.bpm;
    const uint32_t my = mcu / a.mcux, mx = mcu - my * a.mcux;
    at = (size_t)(my * q.v + q.sy) * q.row_blocks + (mx * q.h + q.sx);
  } else {
    const uint32_t by = b / a.nbx, bx = b - by * a.nbx;
    at = (size_t)by * q.row_blocks + bx;
  }
  return q.plane + at * 64;
}

// One Huffman symbol at the head of `bits`; -1 = no such code.  Advances pos.
__device__ __forceinline__ int pscan_symbol(RawBits &bits, const HuffDev &t, uint32_t &pos) {
  const uint32_t look = bits.peek(16);
  const uint32_t e = t.lut[look >> 6];
  int l = (int)(e >> 8), sym = (int)(e & 0xFFu);
  if (e == 0) {
    sym = -1;
    for (l = 11; l <= 16; l++) {
      const int code = (int)(look >> (16 - l));
      if (code <= t.maxcode[l]) {
        sym = t.vals[(t.valoff[l] + code) & 0xFF];
        break;
      }
    }
    if (sym < 0) return -1;
  }
  bits.drop(l);
  pos += (uint32_t)l;
  return sym;
}

// The code words of a subsequence that start before bit `limit`, from state `entry` = (pos, c, k).
// MODE 1: DC first, MODE 2: AC first.  FINAL as in run_subsequence.
template <bool FINAL, int MODE>
__device__ __forceinline__ uint64_t run_pscan(const PscanArgs &a, const HuffDev *tabs, const uint8_t *zz, const PSlot *slots,
                                              uint64_t entry, uint32_t limit, int4 &tup, uint32_t b, int pred0, int pred1,
                                              int pred2, bool *bad, uint32_t *done_blocks) {
  uint32_t pos = (uint32_t)entry, c = (uint32_t)(entry >> 32) & 0xFFu, k = (uint32_t)(entry >> 40) & 0xFFu;
  RawBits bits;
  bits.start(a.data, pos);
  int nb = 0, dcs0 = 0, dcs1 = 0, dcs2 = 0;
  const uint32_t b0 = b;
  while (pos < limit) {
    if (FINAL && b >= a.total_blocks) break;
    if (bits.len <= 32) bits.refill();
    if (MODE == 1) {
      const int kc = (int)slots[c].comp;
      const int sym = pscan_symbol(bits, tabs[kc], pos);
      if (sym < 0) {
        if (FINAL) { *bad = true; break; }
        bits.drop(1);
        pos += 1;
        continue;
      }
      const int n = sym & 15;
      int v = 0;
      if (n) {
        v = extend(bits.peek(n), n);
        bits.drop(n);
        pos += (uint32_t)n;
      }
      if (FINAL) {
        pred0 += kc == 0 ? v : 0;
        pred1 += kc == 1 ? v : 0;
        pred2 += kc == 2 ? v : 0;
        const int pk = kc == 0 ? pred0 : (kc == 1 ? pred1 : pred2);
        const int room = slots[c].limbits - (int)a.al;  // |pk << al| must stay inside the planes' range
        if (room <= 0 || !within_bits(pk, room)) { *bad = true; break; }
        pscan_block(a, slots, b, c)[0] = (int16_t)(pk * (1 << a.al));
        b++;
      } else {
        dcs0 += kc == 0 ? v : 0;
        dcs1 += kc == 1 ? v : 0;
        dcs2 += kc == 2 ? v : 0;
      }
      c = c + 1 == a.bpm ? 0 : c + 1;
      nb++;
    } else {
      const int kc = (int)a.comps[0];
      const int sym = pscan_symbol(bits, tabs[3 + kc], pos);
      if (sym < 0) {
        if (FINAL) { *bad = true; break; }
        bits.drop(1);
        pos += 1;
        continue;
      }
      const uint32_t r = (uint32_t)sym >> 4, sz = (uint32_t)sym & 15u;
      uint32_t adv = 0;  // blocks completed by this code word
      if (sz == 0) {
        if (r < 15) {  // EOBn: this block and the next 2**r - 1 + bits blocks are done
          uint32_t run = (1u << r) - 1u;
          if (r) {
            run += bits.peek((int)r);
            bits.drop((int)r);
            pos += r;
          }
          adv = 1u + run;
        } else {  // ZRL
          k += 16;
          if (k > a.se) adv = 1;
        }
      } else {
        k += r;
        const bool outside = k > a.se;
        if (FINAL && (outside || (int)(sz + a.al) > slots[0].limbits)) { *bad = true; break; }
        const int v = extend(bits.peek((int)sz), (int)sz);
        bits.drop((int)sz);
        pos += sz;
        if (FINAL) pscan_block(a, slots, b, 0)[zz[k & 63u]] = (int16_t)(v * (1 << a.al));
        k++;
        if (outside || k > a.se) adv = 1;
      }
      if (adv) {
        k = a.ss;
        nb += (int)adv;
        if (FINAL) b += adv;
      }
    }
  }
  tup = make_int4(nb, dcs0, dcs1, dcs2);
  if (FINAL) *done_blocks = min(b, a.total_blocks) - b0;
  return pack_state(pos, c, k);
}

template <bool FINAL>
__device__ __forceinline__ uint64_t run_pscan_mode(const PscanArgs &a, const HuffDev *tabs, const uint8_t *zz, const PSlot *slots,
                                                   uint64_t entry, uint32_t limit, int4 &tup, uint32_t b, int p0, int p1, int p2,
                                                   bool *bad, uint32_t *done) {
  return a.mode == 1 ? run_pscan<FINAL, 1>(a, tabs, zz, slots, entry, limit, tup, b, p0, p1, p2, bad, done)
                     : run_pscan<FINAL, 2>(a, tabs, zz, slots, entry, limit, tup, b, p0, p1, p2, bad, done);
}

// Step 1, as huffman_sync_kernel.  A block starts in state (pos, 0, k0): k0 = 0 for DC scans, Ss for AC scans.
// Which scan of the group a CTA works for, and which CTA of that scan it is: `first_cta[s]` = the launch's first
// CTA of scan s (nscans + 1 entries).  The scan's arguments are copied to shared memory.
__device__ __forceinline__ uint32_t pscan_pick(const PscanArgs *__restrict__ args, const uint32_t *__restrict__ first_cta,
                                               uint32_t nscans, PscanArgs *sa) {
  uint32_t sc = 0;
  while (sc + 1 < nscans && blockIdx.x >= __ldg(first_cta + sc + 1)) sc++;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(args + sc);
  uint32_t *dst = reinterpret_cast<uint32_t *>(sa);
  for (uint32_t i = threadIdx.x; i < sizeof(PscanArgs) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
  return blockIdx.x - __ldg(first_cta + sc);
}

__global__ void __launch_bounds__(kWindow) pscan_sync_kernel(const PscanArgs *__restrict__ args,
                                                             const uint32_t *__restrict__ first_cta, uint32_t nscans,
                                                             uint32_t shift, int first) {
  __shared__ HuffDev tabs[6];
  __shared__ uint8_t zz[64];
  __shared__ uint64_t X[kWindow];
  __shared__ PSlot slots[12];
  __shared__ PscanArgs sa;
  const uint32_t cta = pscan_pick(args, first_cta, nscans, &sa);
  const PscanArgs &a = sa;
  pscan_load(a, tabs, zz, slots);
  const uint32_t t = threadIdx.x;
  const uint32_t i = cta * kWindow + t + shift;
  const bool active = i < a.nsub;
  const uint32_t k0 = a.mode == 2 ? a.ss : 0u;
  const uint64_t true_start = pack_state(0u, 0u, k0);
  uint64_t entry = 0, exit_ = 0, head = 0;
  int4 tup = make_int4(0, 0, 0, 0);
  bool dirty = false;
  __syncthreads();
  const uint32_t limit = active ? min((i + 1) * a.sub_bits, a.nbits) : 0u;
  if (active) {
    if (first) {
      entry = i == 0 ? true_start : pack_state(i * a.sub_bits, 0u, k0);
      exit_ = run_pscan_mode<false>(a, tabs, zz, slots, entry, limit, tup, 0u, 0, 0, 0, nullptr, nullptr);
      dirty = true;
    } else {
      entry = a.entry[i];
      exit_ = a.exit_[i];
    }
    if (t == 0) head = i == 0 ? true_start : (first ? entry : a.exit_[i - 1]);
  }
  X[t] = exit_;
  for (int iter = 0; iter < kMaxRounds; iter++) {
    __syncthreads();
    const uint64_t want = t == 0 ? head : X[t - 1];
    const bool redo = active && want != entry;
    uint64_t nx = exit_;
    if (redo) {
      entry = want;
      nx = run_pscan_mode<false>(a, tabs, zz, slots, entry, limit, tup, 0u, 0, 0, 0, nullptr, nullptr);
      dirty = true;
    }
    __syncthreads();
    const bool changed = redo && nx != exit_;
    if (redo) {
      exit_ = nx;
      X[t] = nx;
    }
    if (!__syncthreads_or(changed ? 1 : 0)) break;
  }
  if (active && dirty) {
    a.entry[i] = entry;
    a.exit_[i] = exit_;
    a.tuple[i] = tup;
  }
}

// Step 3, as huffman_output_kernel.
__global__ void __launch_bounds__(128) pscan_output_kernel(const PscanArgs *__restrict__ args,
                                                           const uint32_t *__restrict__ first_cta, uint32_t nscans) {
  __shared__ HuffDev tabs[6];
  __shared__ uint8_t zz[64];
  __shared__ PSlot slots[12];
  __shared__ PscanArgs sa;
  const uint32_t cta = pscan_pick(args, first_cta, nscans, &sa);
  const PscanArgs &a = sa;
  pscan_load(a, tabs, zz, slots);
  __syncthreads();
  const uint32_t i = cta * blockDim.x + threadIdx.x;
  if (i >= a.nsub) return;
  const int4 pre = a.prefix[i];
  const uint32_t b = (uint32_t)pre.x;
  if (b >= a.total_blocks) return;  // padding after the last block
  const uint32_t k0 = a.mode == 2 ? a.ss : 0u;
  const uint64_t entry = i == 0 ? pack_state(0u, 0u, k0) : a.exit_[i - 1];
  bool chain_bad = entry != a.entry[i];
  const uint32_t c = (uint32_t)(entry >> 32) & 0xFFu;
  if (c >= a.bpm || b % a.bpm != c) chain_bad = true;
  bool code_bad = false;
  uint32_t done = 0;
  if (!chain_bad) {
    int4 tup;
    const uint32_t limit = min((i + 1) * a.sub_bits, a.nbits);
    const uint64_t out = run_pscan_mode<true>(a, tabs, zz, slots, entry, limit, tup, b, pre.y, pre.z, pre.w, &code_bad, &done);
    if (!code_bad && b + (uint32_t)tup.x < a.total_blocks && out != a.exit_[i]) chain_bad = true;
    // the thread that stored the scan's last block says where the scan's code words ended: the host holds
    // that against the scan's length (bytes left over, or bits taken from the padding, are a damaged
    // stream: the serial decoder's case)
    if (!code_bad && done != 0 && b + done == a.total_blocks) a.status[3] = (uint32_t)out + 1u;
  }
  if (chain_bad) atomicOr(a.status, 2u);
  if (code_bad) atomicOr(a.status, 8u);
  if (done) atomicAdd(a.status + 2, done);
}

// The prefix sums of step 2 over PscanArgs (same arrays as WholeArgs)
__global__ void __launch_bounds__(1024) pscan_scan_kernel(const PscanArgs *__restrict__ args) {
  __shared__ int4 warp_tot[32];
  __shared__ int4 carry_s;
  __shared__ PscanArgs sa;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(args + blockIdx.x);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sa);
    for (uint32_t i = threadIdx.x; i < sizeof(PscanArgs) / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const PscanArgs &a = sa;
  const uint32_t t = threadIdx.x, lane = t & 31u, w = t >> 5;
  const int4 zero = make_int4(0, 0, 0, 0);
  if (t == 0) carry_s = zero;
  __syncthreads();
  for (uint32_t base = 0; base < a.nsub; base += 4096) {
    const uint32_t i0 = base + 4 * t;
    int4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = i0 + k < a.nsub ? a.tuple[i0 + k] : zero;
    const int4 mine = add4(add4(v[0], v[1]), add4(v[2], v[3]));
    const int4 s = warp_inclusive(mine, lane);
    if (lane == 31) warp_tot[w] = s;
    __syncthreads();
    if (w == 0) {
      const int4 u = warp_tot[lane];
      warp_tot[lane] = sub4(warp_inclusive(u, lane), u);
    }
    __syncthreads();
    int4 ex = add4(add4(carry_s, warp_tot[w]), sub4(s, mine));
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < a.nsub) a.prefix[i0 + k] = ex;
      ex = add4(ex, v[k]);
    }
    __syncthreads();
    if (t == 1023) carry_s = ex;
    __syncthreads();
  }
}

// DC refinement (:1341-1380): bit b of the scan belongs to block b; a 1 sets bit Al of the DC term.
__global__ void __launch_bounds__(256) pscan_dc_refine_kernel(PscanArgs a) {
  __shared__ PSlot slots[12];
  if (threadIdx.x == 0) pscan_slots(a, slots);
  __syncthreads();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.total_blocks) return;
  const uint32_t bit = (__ldg(a.data + (b >> 3)) >> (7u - (b & 7u))) & 1u;
  if (!bit) return;
  int16_t *blk = pscan_block(a, slots, b, b % a.bpm);
  blk[0] = (int16_t)(blk[0] | (int16_t)(1 << a.al));
}

// Non-zero history of a plane for the host's AC refinement decoder: bit k of masks[b] = the coefficient at
// zig-zag position k of block b is non-zero.
__global__ void __launch_bounds__(128) pscan_masks_kernel(const int16_t *__restrict__ plane, uint32_t nblocks,
                                                          uint64_t *__restrict__ masks) {
  __shared__ uint8_t zz[64];
  if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  __syncthreads();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint4 *src = reinterpret_cast<const uint4 *>(plane + (size_t)b * 64);
  uint64_t nat = 0;  // bit n = coefficient at natural position n is non-zero
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint4 w = __ldg(src + j);
    const uint32_t q[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (q[i] & 0xFFFFu) nat |= 1ull << (8 * j + 2 * i);
      if (q[i] >> 16) nat |= 1ull << (8 * j + 2 * i + 1);
    }
  }
  uint64_t m = 0;
  for (int k = 0; k < 64; k++) m |= ((nat >> zz[k]) & 1ull) << k;
  masks[b] = m;
}

// The host's AC refinement verdict on the planes (:1600-1690 as the host mirror's ac_refine applies it):
//   corr   the coefficient, non-zero before the scan, read a correction bit of 1: add 1 << al away from zero
//          unless bit al of it is set already;
//   newpos the coefficient was zero and becomes +(1 << al), or -(1 << al) where newneg is set as well.
__global__ void __launch_bounds__(128) pscan_apply_kernel(int16_t *__restrict__ plane, uint32_t nblocks,
                                                          const uint64_t *__restrict__ corr, const uint64_t *__restrict__ newpos,
                                                          const uint64_t *__restrict__ newneg, uint32_t al) {
  __shared__ uint8_t zz[64];
  if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  __syncthreads();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  int16_t *blk = plane + (size_t)b * 64;
  const int p1 = 1 << al, m1 = -(1 << al);
  uint64_t m = __ldg(corr + b);
  while (m) {
    const int k = __ffsll((long long)m) - 1;
    m &= m - 1;
    const int z = blk[zz[k]];
    if ((z & p1) == 0) blk[zz[k]] = (int16_t)(z >= 0 ? z + p1 : z + m1);
  }
  m = __ldg(newpos + b);
  const uint64_t neg = __ldg(newneg + b);
  while (m) {
    const int k = __ffsll((long long)m) - 1;
    m &= m - 1;
    blk[zz[k]] = (int16_t)(((neg >> k) & 1ull) ? m1 : p1);
  }
}

}  // namespace

cudaError_t launch_huffman_whole(const WholeArgs &a, cudaStream_t stream) {
  if (a.nsub == 0) return cudaSuccess;
  huffman_sync_kernel<<<(a.nsub + kWindow - 1) / kWindow, kWindow, 0, stream>>>(a, 0u, 1);
  if (a.nsub > kWindow / 2)
    huffman_sync_kernel<<<(a.nsub - kWindow / 2 + kWindow - 1) / kWindow, kWindow, 0, stream>>>(a, kWindow / 2, 0);
  return launch_huffman_finish(a, stream);
}

// Steps 2 and 3 alone: prefix sums and the verified output pass over the chain as it stands
// (after the sync launches, or after the host has repaired it).
cudaError_t launch_huffman_finish(const WholeArgs &a, cudaStream_t stream) {
  if (a.nsub == 0) return cudaSuccess;
  huffman_scan_kernel<<<1, 1024, 0, stream>>>(a);
  huffman_output_kernel<<<(a.nsub + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_pscan_group(const PscanArgs *d_args, const uint32_t *d_tables, uint32_t nscans, uint32_t ctas_a,
                               uint32_t ctas_b, uint32_t ctas_out, cudaStream_t stream) {
  if (nscans == 0 || ctas_a == 0) return cudaSuccess;
  pscan_sync_kernel<<<ctas_a, kWindow, 0, stream>>>(d_args, d_tables, nscans, 0u, 1);
  if (ctas_b) pscan_sync_kernel<<<ctas_b, kWindow, 0, stream>>>(d_args, d_tables + (nscans + 1), nscans, kWindow / 2, 0);
  pscan_scan_kernel<<<nscans, 1024, 0, stream>>>(d_args);
  pscan_output_kernel<<<ctas_out, 128, 0, stream>>>(d_args, d_tables + 2 * (nscans + 1), nscans);
  return cudaGetLastError();
}

cudaError_t launch_pscan_dc_refine(const PscanArgs &a, cudaStream_t stream) {
  if (a.total_blocks == 0) return cudaSuccess;
  pscan_dc_refine_kernel<<<(a.total_blocks + 255) / 256, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_pscan_masks(const int16_t *plane, uint32_t nblocks, uint64_t *masks, cudaStream_t stream) {
  if (nblocks == 0) return cudaSuccess;
  pscan_masks_kernel<<<(nblocks + 127) / 128, 128, 0, stream>>>(plane, nblocks, masks);
  return cudaGetLastError();
}

cudaError_t launch_pscan_apply(int16_t *plane, uint32_t nblocks, const uint64_t *corr, const uint64_t *newpos,
                               const uint64_t *newneg, uint32_t al, cudaStream_t stream) {
  if (nblocks == 0) return cudaSuccess;
  pscan_apply_kernel<<<(nblocks + 127) / 128, 128, 0, stream>>>(plane, nblocks, corr, newpos, newneg, al);
  return cudaGetLastError();
}

}  // namespace gidk
