// recon_kernels.cu -- sm_100a kernels of the JPEG reconstruction path.
//
// Mapping: one thread = one MCU, one WARP = one tile of 32 consecutive MCUs
// (recon_launch.h).  A thread does, entirely in registers, dequant + IDCT of its
// chroma blocks (kept as packed bytes in a shared-memory scratch), then dequant +
// IDCT + upsample + colour of each luma block and writes 24-byte pixel runs;
// neighbouring lanes write neighbouring runs of the same image rows.
//
// recon_tma_kernel (main path): every warp is an autonomous pipeline -- no
//   CTA-wide barrier after set-up.  Warp w of the grid owns tiles w, w + W,
//   w + 2W, ... (W = warps in the grid), so that the warps resident at any moment
//   stream neighbouring 4 KiB pieces of the planes.  Coefficients reach the lanes
//   through a private 2-slot ring of 4 KiB shared-memory slots filled by TMA
//   (cp.async.bulk.tensor, box = 32 blocks x 64 coefficients, 128-byte swizzle so
//   that lane t reading row t in 16-byte pieces is bank-conflict free), one
//   mbarrier per slot.  As soon as the warp has pulled an item into registers,
//   lane 0 re-arms the slot with the item two steps ahead; the arithmetic of one
//   item (about a thousand instructions per lane) covers the fetch of the next.
//   Small tiles + interleaved ownership keep the tail of single-frame launches
//   short; 15 warps per SM (3 CTAs x 5) are resident, bounded by shared memory.
// recon_ldg_kernel (fallback): same body, 128-bit global loads; used when the
//   coefficient planes cannot be described by one tensor map (alignment).
// generic_*: two-pass path for sampling layouts without a fused kernel.
#include "recon_launch.h"

namespace gidk {

namespace {

constexpr int kFrameWords = (int)(sizeof(FrameDev) / sizeof(uint32_t));
static_assert(sizeof(FrameDev) % 16 == 0, "FrameDev is copied in 16-byte pieces");

// CTA shapes.  Warps are assigned to the four SM sub-partitions round-robin inside
// a CTA, so the warp count per CTA must spread evenly over them (5 warps per CTA
// put 6 of 15 warps on one scheduler and cost 20 %: measured, profiles/r01c).
//   Wide  : 1 CTA x 15 warps per SM -- the most warps shared memory allows; best for
//           launches with many tiles per warp (batches).
//   Narrow: 3 CTAs x 4 warps per SM -- SMs are handed to the next launch a CTA at a
//           time, which shortens tail + ramp of short launches (one frame): +3 %.
// 16 warps per SM = 4 per scheduler, the most 128 registers per thread allow.  Shared memory
// per warp: the 8 KiB ring, 4 KiB of chroma scratch, the frame copy and the barriers (the luma
// samples borrow the ring slot of their own item, see mcu_item).  (Dropping the per-warp copy of
// the frame descriptor instead and reading it from global memory was measured: -6 % on 4:2:0
// batches, -13 % on grey.)
constexpr int kWideWarps = 16;
template <int WPC> struct Shape {
  static_assert(WPC == kWideWarps || WPC == 4, "supported CTA shapes");
  static constexpr int kCtasPerSm = WPC == kWideWarps ? 1 : 4;
  static constexpr int kThreads = WPC * 32;
};
constexpr int kNumSlots = 2;
constexpr int kSlotBytes = kTileMcus * 128;            // one block (64 x int16) per lane
constexpr int kRingBytes = kNumSlots * kSlotBytes;
constexpr int kScratchBytes = 2 * 8 * kTileMcus * 8;   // per warp: [Cb, Cr][8 rows][32 lanes] x 8 bytes
constexpr int kLdgScratchBytes = 3 * 8 * kTileMcus * 8;  // plain-load kernel: + a luma plane
constexpr int kFrameBytes = (int)((sizeof(FrameDev) + 127) / 128 * 128);
constexpr int kBarBytes = kNumSlots * 8 + 64;  // per warp: the slot barriers + the ring of claimed tiles (4 x uint4)
// dynamic shared memory of one CTA: rings (1024-byte aligned), scratches, frame caches, barriers
template <int WPC> struct SmemLayout {
  static constexpr int kOffScratch = WPC * kRingBytes;
  static constexpr int kOffFrame = kOffScratch + WPC * kScratchBytes;
  static constexpr int kOffBars = kOffFrame + WPC * kFrameBytes;
  static constexpr int kBytes = kOffBars + WPC * kBarBytes + 1024;  // + slack for the alignment
};
// launches with fewer tiles than this per resident warp use the narrow shape
constexpr uint32_t kNarrowBelowTilesPerWarp = 6;
static_assert(kTileMcus == 32, "one tile per warp");

template <int LAYOUT> struct LayoutInfo;
template <> struct LayoutInfo<LAYOUT_GREY> { static constexpr int H = 1, V = 1, NI = 1; static constexpr bool linear = true; };
template <> struct LayoutInfo<LAYOUT_444> { static constexpr int H = 1, V = 1, NI = 3; static constexpr bool linear = true; };
template <> struct LayoutInfo<LAYOUT_422> { static constexpr int H = 2, V = 1, NI = 4; static constexpr bool linear = false; };
template <> struct LayoutInfo<LAYOUT_420> { static constexpr int H = 2, V = 2, NI = 6; static constexpr bool linear = false; };
template <> struct LayoutInfo<LAYOUT_440> { static constexpr int H = 1, V = 2, NI = 4; static constexpr bool linear = false; };
template <> struct LayoutInfo<LAYOUT_411> { static constexpr int H = 4, V = 1, NI = 6; static constexpr bool linear = false; };

// ---- PTX wrappers -------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ---- block source: the warp's private TMA ring -------------------------------------

// Tiles are handed out dynamically: a warp claims its next tile with an atomic
// on a launch-wide counter, so warps on a lightly loaded scheduler (or a faster
// SM) simply take more of them and nobody waits at the end of the launch (static
// ownership left the SMs idle 14-22 % of the time: profiles/r01e).  The claim for
// the tile AFTER the next one is issued early, so the atomic's latency hides
// behind a whole tile of arithmetic.  The last warp to finish resets the
// counters for the next launch of the same plan.
struct TileClaims {
  uint32_t *counters;  // [0] next dynamic tile - 2 * nwarps, [1] warps finished
  uint32_t pending;    // lane 0: result of the claim in flight
  uint32_t gw, nwarps; // this warp's index in the grid, warps in the grid
  uint32_t taken;
  // The first two tiles of a warp are fixed (gw, gw + nwarps): 2 220 warps hitting
  // one counter at the same instant at kernel start serialise in L2 for microseconds.
  // Later claims are spread in time and hidden behind a tile of work.
  // Plain atom (not atomicAdd: the compiler's warp-aggregated form reads the result
  // back at once, which would expose the latency this scheme exists to hide).
  __device__ __forceinline__ void issue() {
    if ((threadIdx.x & 31) == 0)
      asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(pending) : "l"(counters) : "memory");
  }
  __device__ __forceinline__ uint32_t take() {  // the next tile, to every lane
    const uint32_t n = taken++;
    if (n < 2u) return gw + n * nwarps;
    const uint32_t t = 2u * nwarps + __shfl_sync(0xFFFFFFFFu, pending, 0);
    issue();
    return t;
  }
  // The last warp to finish leaves the counters zero for the next launch.  Reading
  // `pending` first makes sure this warp's last claim has been performed (its value
  // has come back) before the warp counts itself as finished, so no claim can land
  // after the reset; no fence needed.
  __device__ __forceinline__ void finish() {
    if ((threadIdx.x & 31) == 0) {
      uint32_t old;
      asm volatile(
          "{ .reg .u32 t;\n\tand.b32 t, %2, 0;\n\tadd.u32 t, t, 1;\n\tatom.global.add.u32 %0, [%1], t; }"
          : "=r"(old)
          : "l"(counters + 1), "r"(pending)
          : "memory");
      if (old == nwarps - 1u) {
        counters[0] = 0u;
        counters[1] = 0u;
      }
    }
  }
};

// Items flow through the ring in a fixed order: for every tile the warp claims,
// items 0 .. NI-1.  Item k lives in slot k & 1; the n-th use of a slot completes
// phase n & 1 of its barrier.  The fetch cursor runs kNumSlots items ahead of the
// consumer; the coordinate of the next box is loaded one step early so that its
// latency hides behind the arithmetic.  Claimed tiles travel from the fetch
// cursor to the consumer through a 4-entry ring in shared memory.
template <int LAYOUT>
struct TmaSource {
  using LI = LayoutInfo<LAYOUT>;
  static constexpr bool kItemsCanBeSkipped = false;  // every item has a slot of the ring armed for it
  __device__ __forceinline__ void prefetch(bool, int, int, int) const {}  // (the ring runs two items ahead)
  uint32_t slots, full;   // shared-window addresses (this warp)
  uint8_t *slots_gen;     // the ring again, as a generic pointer
  uint32_t lane_off;      // lane * 128 + ((lane & 7) << 4): row `lane`, chunk index pre-xored
  uint32_t slot, phase;   // consume cursor
  const TileDesc *tiles;
  const CUtensorMap *map2;
  uint32_t ntiles;
  TileClaims claims;
  volatile uint4 *claimed;  // shared: ring of claimed tiles {tile, frame, m0, -}
  uint32_t n_claimed, n_consumed;
  uint32_t f_tile, f_item, f_coord;  // fetch cursor
  uint32_t h_frame, h_m0;            // header of the tile claimed last (lane 0), stored one step later
  bool h_pending;

  __device__ __forceinline__ void next_fetch_tile() {
    f_tile = claims.take();
    n_claimed++;
    h_pending = true;
    h_frame = h_m0 = 0;
    if (f_tile < ntiles) {
      f_coord = __ldg(&tiles[f_tile].coord[0]);
      h_frame = __ldg(&tiles[f_tile].frame);
      h_m0 = __ldg(&tiles[f_tile].m0);
    }
  }
  // the header loads were issued a whole item ago: publishing them costs no wait
  __device__ __forceinline__ void publish_header() {
    if (h_pending) {
      if ((threadIdx.x & 31) == 0) {
        uint4 e;
        e.x = f_tile; e.y = h_frame; e.z = h_m0; e.w = 0;
        const_cast<uint4 *>(claimed)[(n_claimed - 1u) & 3u] = e;
      }
      h_pending = false;
    }
  }
  __device__ __forceinline__ void start() {
    slot = 0;
    phase = 0;
    n_claimed = n_consumed = 0;
    f_item = 0;
    h_pending = false;
    claims.issue();
    next_fetch_tile();
#pragma unroll
    for (int s = 0; s < kNumSlots; s++) refill((uint32_t)s);
    publish_header();
  }
  // the next tile of this warp (tile >= ntiles: none left).  Its ring entry was
  // published at least one refill ago (the fetch cursor is kNumSlots items ahead).
  __device__ __forceinline__ uint4 next_tile() {
    __syncwarp();
    const volatile uint4 *e = &claimed[(n_consumed++) & 3u];
    uint4 v;
    v.x = e->x; v.y = e->y; v.z = e->z; v.w = 0;
    return v;
  }
  // lane 0 arms slot s with the item under the fetch cursor; every lane advances the cursor
  __device__ __forceinline__ void refill(uint32_t s) {
    publish_header();
    if (f_tile >= ntiles) return;
    if ((threadIdx.x & 31) == 0) {
      const uint32_t bar = full + 8 * s, dst = slots + s * kSlotBytes;
      mbar_arrive_expect_tx(bar, kSlotBytes);
      tma_load_2d(dst, map2, 0, (int)f_coord, bar);
    }
    if (++f_item == (uint32_t)LI::NI) {
      f_item = 0;
      next_fetch_tile();
    } else {
      f_coord = __ldg(&tiles[f_tile].coord[f_item]);
    }
  }
  // consume the next item: this lane's block (row `lane` of the slot) -> raw[8]
  __device__ __forceinline__ void load(bool, int, int, int, u32x4 (&raw)[8]) {
    mbar_wait(full + 8 * slot, phase);
    const uint32_t base = slots + slot * kSlotBytes + lane_off;
#pragma unroll
    for (int j = 0; j < 8; j++) raw[j] = lds128(base ^ (uint32_t)(j << 4));
    __syncwarp();  // every lane has its block in registers: the slot's bytes are free
  }
  // The slot just read, as [8 rows][32 lanes] x 8 bytes for the packed luma samples (same layout
  // as the chroma scratch; the first 2 KiB of the 4 KiB slot).
  __device__ __forceinline__ u32x2 *luma_scratch(u32x2 *) const {
    return reinterpret_cast<u32x2 *>(slots_gen + slot * kSlotBytes) + (threadIdx.x & 31);
  }
  // hand the slot back to the async proxy, re-armed with the item under the fetch cursor
  __device__ __forceinline__ void release() {
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    refill(slot);
    slot ^= 1u;
    if (slot == 0) phase ^= 1u;
  }
};

__device__ __forceinline__ Scratch make_scratch(uint8_t *base, uint32_t lane, bool with_luma) {
  Scratch sc;
  u32x2 *p = reinterpret_cast<u32x2 *>(base) + lane;
  sc.cb = p;
  sc.cr = p + 8 * kTileMcus;
  sc.y = with_luma ? p + 16 * kTileMcus : nullptr;
  sc.stride = kTileMcus;
  return sc;
}

template <int LAYOUT, bool Q8, bool RGBA, int ROT, class Src>
__device__ __forceinline__ void run_mcu(bool active, const FrameDev &f, Src &src, const Scratch &sc, int lane, int mx,
                                        int my) {
  using LI = LayoutInfo<LAYOUT>;
  mcu_run<LI::H, LI::V, LAYOUT == LAYOUT_GREY, Q8, RGBA, ROT>(active, f, &f.qw[0][0], src, sc, lane, mx, my);
}

// MCU (mx, my) of lane t in a tile starting at m0, or false when the lane has
// no MCU (tail of the frame / of the MCU row).
template <int LAYOUT, bool VERT = false>
__device__ __forceinline__ bool tile_mcu(const FrameDev &f, uint32_t m0, uint32_t t, int &mx, int &my) {
  if (VERT) {  // quarter turns: tile_rows MCUs down x 32 / tile_rows MCUs across, from MCU m0
    const uint32_t my0 = m0 / (uint32_t)f.mcux, th = f.tile_rows;
    mx = (int)(m0 - my0 * (uint32_t)f.mcux + t / th);
    my = (int)(my0 + (t & (th - 1u)));
    return mx < f.mcux && my < f.mcuy;
  }
  if (LayoutInfo<LAYOUT>::linear) {
    const uint32_t m = m0 + t;
    my = (int)(m / (uint32_t)f.mcux);
    mx = (int)(m - (uint32_t)my * (uint32_t)f.mcux);
    return m < (uint32_t)(f.mcux * f.mcuy);
  }
  my = (int)(m0 / (uint32_t)f.mcux);
  mx = (int)(m0 - (uint32_t)my * (uint32_t)f.mcux + t);
  return mx < f.mcux;
}

// warp-level copy of a frame descriptor into the warp's shared-memory cache
__device__ __forceinline__ void cache_frame(uint8_t *dst, const FrameDev *src, uint32_t lane) {
  const uint4 *g = reinterpret_cast<const uint4 *>(src);
  uint4 *d = reinterpret_cast<uint4 *>(dst);
  __syncwarp();
  for (uint32_t i = lane; i < sizeof(FrameDev) / 16; i += 32) d[i] = __ldg(g + i);
  __syncwarp();
}

template <int LAYOUT, bool Q8, bool RGBA, int WPC, int ROT = TURN_NONE>
__global__ void __launch_bounds__(Shape<WPC>::kThreads, Shape<WPC>::kCtasPerSm)
recon_tma_kernel(const __grid_constant__ CUtensorMap map2, const FrameDev *__restrict__ frames, const TileDesc *__restrict__ tiles, uint32_t ntiles,
                 uint32_t *__restrict__ counters) {
  extern __shared__ uint8_t dyn_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kOffScratch = SmemLayout<WPC>::kOffScratch, kOffFrame = SmemLayout<WPC>::kOffFrame,
                kOffBars = SmemLayout<WPC>::kOffBars;
  const uint32_t nwarps = gridDim.x * WPC;
  // Programmatic dependent launch (plan launches, LaunchSet::launch(pdl = true)): frames of consecutive
  // launches are independent, so the next launch of the stream may start filling SMs as soon as this
  // one's CTAs leave them -- the tail of one launch runs under the ramp of the next, in ONE stream.
  // A no-op when the launch does not carry the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t base = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  uint8_t *gen = dyn_smem + (base - smem_u32(dyn_smem));
  TmaSource<LAYOUT> src;
  src.slots = base + warp * kRingBytes;
  src.slots_gen = gen + warp * kRingBytes;
  src.full = base + kOffBars + warp * kBarBytes;
  src.claimed = reinterpret_cast<volatile uint4 *>(gen + kOffBars + warp * kBarBytes + kNumSlots * 8);
  src.lane_off = lane * 128u + ((lane & 7u) << 4);
  src.tiles = tiles;
  src.map2 = &map2;
  src.ntiles = ntiles;
  src.claims.counters = counters;
  src.claims.pending = 0;
  src.claims.gw = blockIdx.x * WPC + warp;
  src.claims.nwarps = nwarps;
  src.claims.taken = 0;
  if (lane == 0) {
    for (int s = 0; s < kNumSlots; s++) mbar_init(src.full + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  src.start();

  const Scratch scratch = make_scratch(gen + kOffScratch + warp * kScratchBytes, lane, false);
  uint8_t *frame_cache = gen + kOffFrame + warp * kFrameBytes;
  const FrameDev &f = *reinterpret_cast<const FrameDev *>(frame_cache);
  uint32_t cur_frame = 0xFFFFFFFFu;
#pragma unroll 1
  for (uint4 t = src.next_tile(); t.x < ntiles; t = src.next_tile()) {
    const uint32_t this_frame = t.y, this_m0 = t.z;
    if (this_frame != cur_frame) {
      cache_frame(frame_cache, frames + this_frame, lane);
      cur_frame = this_frame;
    }
    int mx, my;
    const bool active = tile_mcu<LAYOUT>(f, this_m0, lane, mx, my);
    run_mcu<LAYOUT, Q8, RGBA, ROT>(active, f, src, scratch, (int)lane, mx, my);
  }
  src.claims.finish();
  // ... and this launch does not complete before its predecessor has (stream order as seen from
  // outside: whatever follows in the stream finds both launches' pixels written)
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

constexpr int kLdgWarps = 4;

template <int LAYOUT, bool Q8, bool RGBA, int ROT = TURN_NONE>
__global__ void __launch_bounds__(kLdgWarps * 32)
recon_ldg_kernel(const FrameDev *__restrict__ frames, const TileDesc *__restrict__ tiles, uint32_t ntiles) {
  __shared__ __align__(16) uint8_t s_frame[kLdgWarps][kFrameBytes];
  __shared__ __align__(16) uint8_t s_scratch[kLdgWarps][kLdgScratchBytes];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * kLdgWarps + warp;
  if (tile >= ntiles) return;
  const uint32_t fi = __ldg(&tiles[tile].frame);
  const uint32_t m0 = __ldg(&tiles[tile].m0);
  cache_frame(s_frame[warp], frames + fi, lane);
  const FrameDev &f = *reinterpret_cast<const FrameDev *>(s_frame[warp]);
  int mx, my;
  const bool active = tile_mcu<LAYOUT, ROT == TURN_90 || ROT == TURN_270>(f, m0, lane, mx, my);
  LdgSource src{&f};
  run_mcu<LAYOUT, Q8, RGBA, ROT>(active, f, src, make_scratch(s_scratch[warp], lane, true), (int)lane, mx, my);
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
    src.load(true, c, bx, by, raw);
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

// Pass 2: one thread per pixel, 32 x 32 pixels per CTA.  Pixel (x, y) reads sample (x / ux_c, y / uy_c)
// of component c (replication, gid-decoding_jpg.adb:974-1008).  Quarter turns: the tile is transposed
// through shared memory, so that the threads of a warp write neighbouring pixels of a row of the ROTATED
// image (a row of the frame is a column there: written as it is read, every pixel would be a store of
// its own sector).
__global__ void __launch_bounds__(1024)
generic_color_kernel(const FrameDev *__restrict__ frame, int ncomp, int hmax, int vmax) {
  __shared__ uint32_t tile[32][33];
  const FrameDev &f = *frame;
  const int fmt = f.out_format & 0xFF, rot = (f.out_format >> 8) & 3;
  const bool quarter = (rot & 1) != 0;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  int x = x0 + (int)threadIdx.x, y = y0 + (int)threadIdx.y;
  uint32_t px = 0;
  if (x < f.width && y < f.height) {
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
    if (fmt == OUT_BGR8_BOTTOM_UP) { const int t = r; r = b; b = t; }
    px = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | 0xFF000000u;
  }
  if (quarter) {  // this thread now writes pixel (x0 + threadIdx.y, y0 + threadIdx.x)
    tile[threadIdx.y][threadIdx.x] = px;
    __syncthreads();
    px = tile[threadIdx.x][threadIdx.y];
    x = x0 + (int)threadIdx.y;
    y = y0 + (int)threadIdx.x;
  }
  if (x >= f.width || y >= f.height) return;
  // display orientation (GIDCUDA_FLAG_ROTATE_*, = Set_X_Y of test/to_bmp.adb:104-122): column and
  // top-down row of the pixel in the output image, whose height is out_h
  int ox = x, oy = y, out_h = f.height;
  if (rot == 1) { ox = y; oy = f.width - 1 - x; out_h = f.width; }
  else if (rot == 2) { ox = f.width - 1 - x; oy = f.height - 1 - y; }
  else if (rot == 3) { ox = f.height - 1 - y; oy = x; out_h = f.width; }
  if (fmt == OUT_BGR8_BOTTOM_UP) oy = out_h - 1 - oy;
  if (fmt == OUT_RGBA8) {
    uint8_t *dst = f.pixels + (size_t)oy * f.stride + (size_t)ox * 4;
    if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
      *reinterpret_cast<uint32_t *>(dst) = px;
    } else {  // a caller-chosen stride that is not a multiple of 4
      dst[0] = (uint8_t)px;
      dst[1] = (uint8_t)(px >> 8);
      dst[2] = (uint8_t)(px >> 16);
      dst[3] = 255;
    }
  } else {
    uint8_t *dst = f.pixels + (size_t)oy * f.stride + (size_t)ox * 3;
    dst[0] = (uint8_t)px;
    dst[1] = (uint8_t)(px >> 8);
    dst[2] = (uint8_t)(px >> 16);
  }
}

template <int LAYOUT, bool Q8, bool RGBA, int WPC, int ROT = TURN_NONE>
cudaError_t launch_shape(const CUtensorMap *map2, const FrameDev *d_frames,
                         const TileDesc *d_tiles, uint32_t ntiles, uint32_t *d_counters, int num_sms,
                         bool pdl, cudaStream_t stream) {
  uint32_t grid = (uint32_t)(num_sms * Shape<WPC>::kCtasPerSm);
  const uint32_t need = (ntiles + WPC - 1) / WPC;
  if (grid > need) grid = need;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Shape<WPC>::kThreads);
  cfg.dynamicSmemBytes = SmemLayout<WPC>::kBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, recon_tma_kernel<LAYOUT, Q8, RGBA, WPC, ROT>, *map2, d_frames, d_tiles, ntiles, d_counters);
}

template <int LAYOUT, bool Q8, bool RGBA>
cudaError_t launch_one(bool use_tma, const CUtensorMap *map2, const FrameDev *d_frames,
                       const TileDesc *d_tiles, uint32_t ntiles, uint32_t *d_counters, int num_sms,
                       bool pdl, cudaStream_t stream) {
  if (use_tma) {
    const bool narrow = ntiles < (uint32_t)num_sms * (uint32_t)kWideWarps * kNarrowBelowTilesPerWarp;
    return narrow ? launch_shape<LAYOUT, Q8, RGBA, 4>(map2, d_frames, d_tiles, ntiles, d_counters, num_sms, pdl, stream)
                  : launch_shape<LAYOUT, Q8, RGBA, kWideWarps>(map2, d_frames, d_tiles, ntiles, d_counters, num_sms, pdl, stream);
  }
  recon_ldg_kernel<LAYOUT, Q8, RGBA><<<(ntiles + kLdgWarps - 1) / kLdgWarps, kLdgWarps * 32, 0, stream>>>(
      d_frames, d_tiles, ntiles);
  return cudaGetLastError();
}

// Rotated frames (FRAME_ROT90 / _ROT180 / _ROT270): instances of their own (fast domain, wide shape) -- see
// emit_block_rows
template <int LAYOUT, bool RGBA, int ROT>
cudaError_t launch_rot(bool use_tma, const CUtensorMap *map2, const FrameDev *d_frames, const TileDesc *d_tiles,
                       uint32_t ntiles, uint32_t *d_counters, int num_sms, bool pdl, cudaStream_t stream) {
  // Quarter turns: tiles of 32 MCUs down a column of MCUs (mcu_item).  Their blocks are bx * 128 bytes apart,
  // not one box of the arena's tensor map: they run the plain-load kernel.
  constexpr bool kVert = ROT == TURN_90 || ROT == TURN_270;
  if (use_tma && !kVert)
    return launch_shape<LAYOUT, true, RGBA, kWideWarps, kVert ? TURN_180 : ROT>(map2, d_frames, d_tiles, ntiles, d_counters, num_sms, pdl, stream);
  recon_ldg_kernel<LAYOUT, true, RGBA, ROT><<<(ntiles + kLdgWarps - 1) / kLdgWarps, kLdgWarps * 32, 0, stream>>>(
      d_frames, d_tiles, ntiles);
  return cudaGetLastError();
}

template <int LAYOUT>
cudaError_t launch_q(bool q8, bool rgba, int rot, bool use_tma, const CUtensorMap *map2,
                     const FrameDev *d_frames, const TileDesc *d_tiles, uint32_t ntiles, uint32_t *d_counters,
                     int num_sms, bool pdl, cudaStream_t stream) {
#define GIDK_ARGS use_tma, map2, d_frames, d_tiles, ntiles, d_counters, num_sms, pdl, stream
  if (rot != TURN_NONE) {
    if (!q8) return cudaErrorInvalidValue;  // (frame_setup.hpp sends those frames to the generic kernels)
    switch (rot) {
      case TURN_90: return rgba ? launch_rot<LAYOUT, true, TURN_90>(GIDK_ARGS) : launch_rot<LAYOUT, false, TURN_90>(GIDK_ARGS);
      case TURN_180: return rgba ? launch_rot<LAYOUT, true, TURN_180>(GIDK_ARGS) : launch_rot<LAYOUT, false, TURN_180>(GIDK_ARGS);
      case TURN_270: return rgba ? launch_rot<LAYOUT, true, TURN_270>(GIDK_ARGS) : launch_rot<LAYOUT, false, TURN_270>(GIDK_ARGS);
      default: return cudaErrorInvalidValue;
    }
  }
  if (q8) return rgba ? launch_one<LAYOUT, true, true>(GIDK_ARGS) : launch_one<LAYOUT, true, false>(GIDK_ARGS);
  return rgba ? launch_one<LAYOUT, false, true>(GIDK_ARGS) : launch_one<LAYOUT, false, false>(GIDK_ARGS);
#undef GIDK_ARGS
}

template <int LAYOUT, bool Q8, bool RGBA>
cudaError_t set_attr() {
  cudaError_t e = cudaFuncSetAttribute(recon_tma_kernel<LAYOUT, Q8, RGBA, kWideWarps>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<kWideWarps>::kBytes);
  if (e != cudaSuccess) return e;
  if (Q8) {
    e = cudaFuncSetAttribute(recon_tma_kernel<LAYOUT, true, RGBA, kWideWarps, TURN_180>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<kWideWarps>::kBytes);
    if (e != cudaSuccess) return e;
  }
  return cudaFuncSetAttribute(recon_tma_kernel<LAYOUT, Q8, RGBA, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              SmemLayout<4>::kBytes);
}
template <int LAYOUT>
cudaError_t set_attr_all() {
  cudaError_t e;
  if ((e = set_attr<LAYOUT, true, false>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT, false, false>()) != cudaSuccess) return e;
  if ((e = set_attr<LAYOUT, true, true>()) != cudaSuccess) return e;
  return set_attr<LAYOUT, false, true>();
}

}  // namespace

cudaError_t init_kernels() {
  cudaError_t e;
  if ((e = set_attr_all<LAYOUT_GREY>()) != cudaSuccess) return e;
  if ((e = set_attr_all<LAYOUT_444>()) != cudaSuccess) return e;
  if ((e = set_attr_all<LAYOUT_422>()) != cudaSuccess) return e;
  if ((e = set_attr_all<LAYOUT_420>()) != cudaSuccess) return e;
  if ((e = set_attr_all<LAYOUT_440>()) != cudaSuccess) return e;
  return set_attr_all<LAYOUT_411>();
}

cudaError_t launch_fused(Layout layout, bool q8, bool rgba, int rotate, bool use_tma, const CUtensorMap *map2,
                         const FrameDev *d_frames, const TileDesc *d_tiles, uint32_t ntiles,
                         uint32_t *d_counters, int num_sms, bool pdl, cudaStream_t stream) {
  if (ntiles == 0) return cudaSuccess;
#define GIDK_ARGS q8, rgba, rotate, use_tma, map2, d_frames, d_tiles, ntiles, d_counters, num_sms, pdl, stream
  switch (layout) {
    case LAYOUT_GREY: return launch_q<LAYOUT_GREY>(GIDK_ARGS);
    case LAYOUT_444: return launch_q<LAYOUT_444>(GIDK_ARGS);
    case LAYOUT_422: return launch_q<LAYOUT_422>(GIDK_ARGS);
    case LAYOUT_420: return launch_q<LAYOUT_420>(GIDK_ARGS);
    case LAYOUT_440: return launch_q<LAYOUT_440>(GIDK_ARGS);
    case LAYOUT_411: return launch_q<LAYOUT_411>(GIDK_ARGS);
    default: return cudaErrorInvalidValue;
  }
#undef GIDK_ARGS
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
  dim3 grid2((unsigned)((h.width + 31) / 32), (unsigned)((h.height + 31) / 32));
  generic_color_kernel<<<grid2, dim3(32, 32), 0, stream>>>(d_frame, ncomp, hmax, vmax);
  return cudaGetLastError();
}

}  // namespace gidk
