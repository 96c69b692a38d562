// recon_kernels.cu -- sm_100a kernels of the JPEG reconstruction path.
//
// Mapping: one thread = one MCU, 128 MCUs per tile (recon_launch.h).  A thread
// does, entirely in registers, dequant + IDCT of its chroma blocks (kept as
// packed bytes), then dequant + IDCT + upsample + colour of each luma block and
// writes 24-byte pixel runs; neighbouring threads write neighbouring runs of
// the same image rows.
//
// recon_tma_kernel (main path): persistent CTAs (3 per SM), each walking a
//   contiguous chunk of the tile list.  Four consumer warps (128 MCUs) plus one
//   producer warp.  Coefficients reach the consumers through a 3-slot ring of
//   16 KiB shared-memory slots filled by TMA (cp.async.bulk.tensor with a
//   128-byte swizzle, so that thread t reading row t of a slot in 16-byte
//   pieces is bank-conflict free); full/empty mbarriers per slot; a slot is
//   released as soon as the four warps have pulled their blocks into registers,
//   so the producer keeps the ring full while the arithmetic runs.
//   The kernel is bound by integer issue (see recon_math.cuh), not by HBM, for
//   4:4:4; the ring keeps the issue slots busy while HBM streams.
// recon_ldg_kernel (fallback): same body, 128-bit global loads; used when the
//   coefficient planes cannot be described by one tensor map (alignment).
// generic_*: two-pass path for sampling layouts without a fused kernel.
#include "recon_launch.h"

namespace gidk {

namespace {

constexpr int kFrameWords = (int)(sizeof(FrameDev) / sizeof(uint32_t));
static_assert(sizeof(FrameDev) % sizeof(uint32_t) == 0, "FrameDev must be word sized");

constexpr int kNumSlots = 3;
constexpr int kSlotBytes = kTileMcus * 128;  // one block (64 x int16) per thread
constexpr int kCtasPerSm = 3;
constexpr int kTmaThreads = kTileMcus + 32;  // consumers + the producer warp

template <int LAYOUT> struct LayoutInfo;
template <> struct LayoutInfo<LAYOUT_GREY> { static constexpr int H = 1, V = 1, NI = 1; static constexpr bool linear = true; };
template <> struct LayoutInfo<LAYOUT_444> { static constexpr int H = 1, V = 1, NI = 3; static constexpr bool linear = true; };
template <> struct LayoutInfo<LAYOUT_422> { static constexpr int H = 2, V = 1, NI = 4; static constexpr bool linear = false; };
template <> struct LayoutInfo<LAYOUT_420> { static constexpr int H = 2, V = 2, NI = 6; static constexpr bool linear = false; };

// ---- PTX wrappers -------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ---- block source: the TMA ring -------------------------------------------------

// Items flow through the ring in a fixed order: for every tile of the CTA's
// chunk, items 0 .. NI-1.  Item k lives in slot k % 3; the n-th use of a slot
// completes phase n & 1 of its full barrier (producer: arrive.expect_tx + TMA
// complete_tx) and, once the four consumer warps have pulled their blocks into
// registers, of its empty barrier.
struct RingCursor {
  uint32_t slot = 0, phase = 0;
  __device__ __forceinline__ void advance() {
    if (++slot == (uint32_t)kNumSlots) { slot = 0; phase ^= 1u; }
  }
};

template <int LAYOUT>
struct TmaSource {
  uint32_t slots, full, empty;  // shared-window addresses
  uint32_t lane_off;            // t * 128 + ((t & 7) << 4): row t, chunk index pre-xored
  RingCursor cur;

  // consume the next item: this thread's block (row t of the slot) -> raw[8]
  __device__ __forceinline__ void load(int, int, int, u32x4 (&raw)[8]) {
    mbar_wait(full + 8 * cur.slot, cur.phase);
    const uint32_t base = slots + cur.slot * kSlotBytes + lane_off;
#pragma unroll
    for (int j = 0; j < 8; j++) raw[j] = lds128(base ^ (uint32_t)(j << 4));
    release();
  }
  // threads without an MCU in this tile still take part in the protocol
  __device__ __forceinline__ void skip() {
    mbar_wait(full + 8 * cur.slot, cur.phase);
    release();
  }
  __device__ __forceinline__ void release() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(empty + 8 * cur.slot);
    cur.advance();
  }
};

// The producer: one thread streams every item of the CTA's tile chunk.
template <int LAYOUT>
__device__ __forceinline__ void produce_all(uint32_t slots, uint32_t full, uint32_t empty, const TileDesc *tiles,
                                            const CUtensorMap *map2, const CUtensorMap *map3, uint32_t t_begin,
                                            uint32_t t_end) {
  using LI = LayoutInfo<LAYOUT>;
  RingCursor cur;
  uint32_t k = 0;
  for (uint32_t tile = t_begin; tile < t_end; tile++) {
#pragma unroll
    for (int item = 0; item < LI::NI; item++, k++) {
      const uint32_t coord = __ldg(&tiles[tile].coord[item]);
      if (k >= (uint32_t)kNumSlots) mbar_wait(empty + 8 * cur.slot, cur.phase ^ 1u);
      const uint32_t bar = full + 8 * cur.slot;
      mbar_arrive_expect_tx(bar, kSlotBytes);
      const uint32_t dst = slots + cur.slot * kSlotBytes;
      if (LI::H == 2 && item >= 2) tma_load_3d(dst, map3, 0, (item - 2) & 1, (int)coord, bar);
      else tma_load_2d(dst, map2, 0, (int)coord, bar);
      cur.advance();
    }
  }
}

// scratch for the packed samples: [3 planes][8 rows][128 threads] x 8 bytes
constexpr int kScratchBytes = 3 * 8 * kTileMcus * 8;

__device__ __forceinline__ Scratch make_scratch(uint8_t *base, uint32_t t) {
  Scratch sc;
  u32x2 *p = reinterpret_cast<u32x2 *>(base) + t;
  sc.cb = p;
  sc.cr = p + 8 * kTileMcus;
  sc.y = p + 16 * kTileMcus;
  sc.stride = kTileMcus;
  return sc;
}

template <int LAYOUT, bool Q8, class Src>
__device__ __forceinline__ void run_mcu(uint32_t lanes, const FrameDev &f, Src &src, const Scratch &sc, int mx,
                                        int my) {
  using LI = LayoutInfo<LAYOUT>;
  mcu_run<LI::H, LI::V, LAYOUT == LAYOUT_GREY, Q8>(lanes, f, &f.qw[0][0], src, sc, mx, my);
}

// MCU (mx, my) of thread t in a tile starting at m0, or false when the thread
// has no MCU (tail of the frame / of the MCU row).
template <int LAYOUT>
__device__ __forceinline__ bool tile_mcu(const FrameDev &f, uint32_t m0, uint32_t t, int &mx, int &my) {
  if (LayoutInfo<LAYOUT>::linear) {
    const uint32_t m = m0 + t;
    if (m >= (uint32_t)(f.mcux * f.mcuy)) return false;
    my = (int)(m / (uint32_t)f.mcux);
    mx = (int)(m - (uint32_t)my * (uint32_t)f.mcux);
    return true;
  }
  my = (int)(m0 / (uint32_t)f.mcux);
  mx = (int)(m0 - (uint32_t)my * (uint32_t)f.mcux + t);
  return mx < f.mcux;
}

constexpr int kDynSmem = kNumSlots * kSlotBytes + kScratchBytes + 1024;  // + slack for 1024-byte alignment

template <int LAYOUT, bool Q8>
__global__ void __launch_bounds__(kTmaThreads, kCtasPerSm)
recon_tma_kernel(const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3,
                 const FrameDev *__restrict__ frames, const TileDesc *__restrict__ tiles, uint32_t ntiles) {
  using LI = LayoutInfo<LAYOUT>;
  extern __shared__ uint8_t dyn_smem[];
  __shared__ __align__(16) uint32_t s_frame_words[kFrameWords];
  __shared__ __align__(8) uint64_t s_full[kNumSlots], s_empty[kNumSlots];

  // contiguous chunk of the tile list per CTA
  const uint32_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const uint32_t t_begin = blockIdx.x * per;
  const uint32_t t_end = min(ntiles, t_begin + per);
  if (t_begin >= t_end) return;

  const uint32_t t = threadIdx.x;
  const uint32_t slots = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t full = smem_u32(s_full), empty = smem_u32(s_empty);
  if (t == 0) {
    for (int s = 0; s < kNumSlots; s++) {
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, kTileMcus / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (t >= (uint32_t)kTileMcus) {  // producer warp
    if (t == (uint32_t)kTileMcus) produce_all<LAYOUT>(slots, full, empty, tiles, &map2, &map3, t_begin, t_end);
    return;
  }

  TmaSource<LAYOUT> src;
  src.slots = slots;
  src.full = full;
  src.empty = empty;
  src.lane_off = t * 128u + ((t & 7u) << 4);
  const Scratch scratch = make_scratch(dyn_smem + (slots - smem_u32(dyn_smem)) + kNumSlots * kSlotBytes, t);
  uint32_t cur_frame = 0xFFFFFFFFu;
#pragma unroll 1
  for (uint32_t tile = t_begin; tile < t_end; tile++) {
    const uint32_t fi = __ldg(&tiles[tile].frame);
    const uint32_t m0 = __ldg(&tiles[tile].m0);
    if (fi != cur_frame) {  // uniform over the consumer warps
      asm volatile("bar.sync 1, %0;" ::"n"(kTileMcus) : "memory");
      const uint32_t *g = reinterpret_cast<const uint32_t *>(frames + fi);
      for (int i = t; i < kFrameWords; i += kTileMcus) s_frame_words[i] = __ldg(g + i);
      asm volatile("bar.sync 1, %0;" ::"n"(kTileMcus) : "memory");
      cur_frame = fi;
    }
    const FrameDev &f = *reinterpret_cast<const FrameDev *>(s_frame_words);
    int mx, my;
    const bool has_mcu = tile_mcu<LAYOUT>(f, m0, t, mx, my);
    const uint32_t lanes = __ballot_sync(0xFFFFFFFFu, has_mcu);  // the lanes that vote in the sparse IDCT
    if (has_mcu) {
      run_mcu<LAYOUT, Q8>(lanes, f, src, scratch, mx, my);
    } else {
#pragma unroll
      for (int i = 0; i < LI::NI; i++) src.skip();
    }
  }
}

template <int LAYOUT, bool Q8>
__global__ void __launch_bounds__(kTileMcus)
recon_ldg_kernel(const FrameDev *__restrict__ frames, const TileDesc *__restrict__ tiles, uint32_t ntiles) {
  __shared__ __align__(16) uint32_t s_frame_words[kFrameWords];
  __shared__ __align__(16) uint8_t s_scratch[kScratchBytes];
  const uint32_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  const uint32_t fi = __ldg(&tiles[tile].frame);
  const uint32_t m0 = __ldg(&tiles[tile].m0);
  {
    const uint32_t *g = reinterpret_cast<const uint32_t *>(frames + fi);
    for (int i = threadIdx.x; i < kFrameWords; i += kTileMcus) s_frame_words[i] = __ldg(g + i);
  }
  __syncthreads();
  const FrameDev &f = *reinterpret_cast<const FrameDev *>(s_frame_words);
  int mx, my;
  const bool has_mcu = tile_mcu<LAYOUT>(f, m0, threadIdx.x, mx, my);
  const uint32_t lanes = __ballot_sync(0xFFFFFFFFu, has_mcu);
  if (!has_mcu) return;
  LdgSource src{&f};
  run_mcu<LAYOUT, Q8>(lanes, f, src, make_scratch(s_scratch, threadIdx.x), mx, my);
}

// ---- generic path -----------------------------------------------------------

// Pass 1: one thread per block of any component; writes clipped samples.
template <bool Q8>
__global__ void __launch_bounds__(128)
generic_idct_kernel(const FrameDev *__restrict__ frame, int ncomp) {
  __shared__ __align__(16) uint32_t s_frame_words[kFrameWords];
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frame);
    for (int i = threadIdx.x; i < kFrameWords; i += 128) s_frame_words[i] = src[i];
  }
  __syncthreads();
  const FrameDev &f = *reinterpret_cast<const FrameDev *>(s_frame_words);
  uint32_t idx = blockIdx.x * 128u + threadIdx.x;
  int c = 0;
  for (; c < ncomp; c++) {
    const uint32_t n = (uint32_t)(f.bx[c] * f.by[c]);
    if (idx < n) break;
    idx -= n;
  }
  if (c >= ncomp) return;
  const int by = (int)(idx / (uint32_t)f.bx[c]);
  const int bx = (int)(idx - (uint32_t)by * (uint32_t)f.bx[c]);
  int b[64];
  {
    LdgSource src{&f};
    u32x4 raw[8];
    src.load(c, bx, by, raw);
    dequant_idct<Q8>(raw, &f.qw[c][0], b);
  }
  uint32_t p[16];
  pack_block(b, p);
  const size_t pitch = (size_t)f.bx[c] * 8;
  uint8_t *dst = f.samples[c] + (size_t)by * 8 * pitch + (size_t)bx * 8;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    uint2 v;
    v.x = p[2 * r];
    v.y = p[2 * r + 1];
    *reinterpret_cast<uint2 *>(dst + (size_t)r * pitch) = v;
  }
}

// Pass 2: one thread per pixel.  Pixel (x, y) reads sample (x / ux_c, y / uy_c)
// of component c (replication, gid-decoding_jpg.adb:974-1008).
__global__ void __launch_bounds__(256)
generic_color_kernel(const FrameDev *__restrict__ frame, int ncomp, int hmax, int vmax) {
  const FrameDev &f = *frame;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= f.width || y >= f.height) return;
  int s[3] = {0, 0, 0};
  for (int c = 0; c < ncomp; c++) {
    const int ux = hmax / f.samp_h[c], uy = vmax / f.samp_v[c];
    s[c] = f.samples[c][(size_t)(y / uy) * ((size_t)f.bx[c] * 8) + (size_t)(x / ux)];
  }
  int r, g, b;
  if (ncomp == 3) {
    const ChromaTerms t = chroma_terms(s[1], s[2]);
    r = clip255(s[0] + t.r);
    g = clip255(s[0] + t.g);
    b = clip255(s[0] + t.b);
  } else {
    r = g = b = s[0];
  }
  int yy = y;
  if (f.out_format == OUT_BGR8_BOTTOM_UP) {
    yy = f.height - 1 - y;
    const int t = r; r = b; b = t;
  }
  uint8_t *dst = f.pixels + (size_t)yy * f.stride + (size_t)x * 3;
  dst[0] = (uint8_t)r;
  dst[1] = (uint8_t)g;
  dst[2] = (uint8_t)b;
}

template <int LAYOUT, bool Q8>
cudaError_t launch_one(bool use_tma, const CUtensorMap *map2, const CUtensorMap *map3, const FrameDev *d_frames,
                       const TileDesc *d_tiles, uint32_t ntiles, int num_sms, cudaStream_t stream) {
  if (use_tma) {
    uint32_t grid = (uint32_t)(num_sms * kCtasPerSm);
    if (grid > ntiles) grid = ntiles;
    recon_tma_kernel<LAYOUT, Q8><<<grid, kTmaThreads, kDynSmem, stream>>>(*map2, *map3, d_frames, d_tiles, ntiles);
  } else {
    recon_ldg_kernel<LAYOUT, Q8><<<ntiles, kTileMcus, 0, stream>>>(d_frames, d_tiles, ntiles);
  }
  return cudaGetLastError();
}

template <int LAYOUT>
cudaError_t launch_q(bool q8, bool use_tma, const CUtensorMap *map2, const CUtensorMap *map3,
                     const FrameDev *d_frames, const TileDesc *d_tiles, uint32_t ntiles, int num_sms,
                     cudaStream_t stream) {
  return q8 ? launch_one<LAYOUT, true>(use_tma, map2, map3, d_frames, d_tiles, ntiles, num_sms, stream)
            : launch_one<LAYOUT, false>(use_tma, map2, map3, d_frames, d_tiles, ntiles, num_sms, stream);
}

template <int LAYOUT, bool Q8>
cudaError_t set_attr() {
  return cudaFuncSetAttribute(recon_tma_kernel<LAYOUT, Q8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem);
}

}  // namespace

cudaError_t init_kernels() {
  cudaError_t e;
  if ((e = set_attr<LAYOUT_GREY, true>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_GREY, false>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_444, true>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_444, false>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_422, true>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_422, false>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_420, true>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT_420, false>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_fused(Layout layout, bool q8, bool use_tma, const CUtensorMap *map2, const CUtensorMap *map3,
                         const FrameDev *d_frames, const TileDesc *d_tiles, uint32_t ntiles, int num_sms,
                         cudaStream_t stream) {
  if (ntiles == 0) return cudaSuccess;
  switch (layout) {
    case LAYOUT_GREY: return launch_q<LAYOUT_GREY>(q8, use_tma, map2, map3, d_frames, d_tiles, ntiles, num_sms, stream);
    case LAYOUT_444: return launch_q<LAYOUT_444>(q8, use_tma, map2, map3, d_frames, d_tiles, ntiles, num_sms, stream);
    case LAYOUT_422: return launch_q<LAYOUT_422>(q8, use_tma, map2, map3, d_frames, d_tiles, ntiles, num_sms, stream);
    case LAYOUT_420: return launch_q<LAYOUT_420>(q8, use_tma, map2, map3, d_frames, d_tiles, ntiles, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_generic(const FrameDev &h, const FrameDev *d_frame, cudaStream_t stream) {
  int ncomp = h.plane[1] ? 3 : 1;
  uint32_t nblocks = 0;
  int hmax = 1, vmax = 1;
  for (int c = 0; c < ncomp; c++) {
    nblocks += (uint32_t)(h.bx[c] * h.by[c]);
    if (h.samp_h[c] > hmax) hmax = h.samp_h[c];
    if (h.samp_v[c] > vmax) vmax = h.samp_v[c];
  }
  const uint32_t grid1 = (nblocks + 127u) / 128u;
  if (h.q8) generic_idct_kernel<true><<<grid1, 128, 0, stream>>>(d_frame, ncomp);
  else generic_idct_kernel<false><<<grid1, 128, 0, stream>>>(d_frame, ncomp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 grid2((unsigned)((h.width + 255) / 256), (unsigned)h.height);
  generic_color_kernel<<<grid2, 256, 0, stream>>>(d_frame, ncomp, hmax, vmax);
  return cudaGetLastError();
}

}  // namespace gidk
