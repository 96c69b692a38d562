// launch_set.hpp -- turns a list of frames into kernel launches: groups frames
// by (sampling layout, table width), builds the tile table with the TMA
// coordinates, encodes the coefficient tensor maps, uploads and launches.
// Used by jobs (one frame), plans (device-resident batches) and the batch
// driver's chunks.  Internal to the shim.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "frame_setup.hpp"
#include "recon_launch.h"

namespace gidk {

struct FrameItem {
  const gidcuda_frame_desc *desc;
  Geometry geo;
  const int16_t *planes[3];  // device pointers
  uint8_t *pixels;           // device pointer
  uint8_t *samples;          // device scratch (generic layouts only)
};

struct LaunchGroup {
  Layout layout;
  bool q8;
  bool rgba;
  int rot;    // Geometry::rotate of the group's frames (rotated frames: kernel instances of their own)
  bool tma;
  uint32_t first_frame, nframes, first_tile, ntiles;
  CUtensorMap map2;
};

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library
// links cudart statically and does not link libcuda).
inline TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}

// TMA staging can be switched off (environment GIDCUDA_DISABLE_TMA=1 or
// gidcuda_set_option("tma", 0)); the plain-load kernel is then used.
inline int &tma_option() {
  static int enabled = [] {
    const char *e = getenv("GIDCUDA_DISABLE_TMA");
    return (e && *e && *e != '0') ? 0 : 1;
  }();
  return enabled;
}
inline bool tma_disabled_by_env() { return tma_option() == 0; }

class LaunchSet {
 public:
  std::vector<FrameDev> frames;  // fused frames (group by group), then generic frames
  std::vector<TileDesc> tiles;
  std::vector<LaunchGroup> groups;
  uint32_t n_fused = 0;
  int tma_groups = 0;

  ~LaunchSet() { destroy(); }

  void destroy() {
    if (d_frames_) cudaFree(d_frames_);
    if (d_tiles_) cudaFree(d_tiles_);
    if (d_counters_) cudaFree(d_counters_);
    if (h_frames_) cudaFreeHost(h_frames_);
    if (h_tiles_) cudaFreeHost(h_tiles_);
    d_frames_ = nullptr;
    d_tiles_ = nullptr;
    d_counters_ = nullptr;
    h_frames_ = nullptr;
    h_tiles_ = nullptr;
    cap_frames_ = cap_tiles_ = 0;
    device_valid_ = false;
  }

  // Host-side preparation.  Every item must have passed make_geometry.
  void build(const std::vector<FrameItem> &items) {
    frames.clear();
    tiles.clear();
    groups.clear();
    tma_groups = 0;
    for (int lay = 0; lay < kFusedLayouts; lay++)
      for (int qa = 15; qa >= 0; qa--) {
        const int q = qa & 1;
        LaunchGroup g;
        memset(&g, 0, sizeof g);
        g.layout = (Layout)lay;
        g.q8 = q == 1;
        g.rgba = (qa & 2) == 0;
        g.rot = 3 - (qa >> 2);   // frames that are not rotated first
        g.first_frame = (uint32_t)frames.size();
        g.first_tile = (uint32_t)tiles.size();
        uintptr_t lo = ~(uintptr_t)0, hi = 0;
        for (const FrameItem &it : items) {
          if ((int)it.geo.layout != lay || it.geo.q8 != g.q8 || (it.geo.out_format == (int)OUT_RGBA8) != g.rgba || it.geo.rotate != g.rot) continue;
          for (int c = 0; c < it.geo.ncomp; c++) {
            const uintptr_t p = reinterpret_cast<uintptr_t>(it.planes[c]);
            if (p < lo) lo = p;
            if (p + it.geo.plane_bytes[c] > hi) hi = p + it.geo.plane_bytes[c];
          }
        }
        if (hi == 0) continue;
        const uintptr_t base = lo & ~(uintptr_t)255;
        bool aligned = true;
        for (const FrameItem &it : items) {
          if ((int)it.geo.layout != lay || it.geo.q8 != g.q8 || (it.geo.out_format == (int)OUT_RGBA8) != g.rgba || it.geo.rotate != g.rot) continue;
          for (int c = 0; c < it.geo.ncomp; c++)
            if ((reinterpret_cast<uintptr_t>(it.planes[c]) - base) % 256 != 0) aligned = false;
        }
        g.tma = aligned && !(g.rot & 1) && !tma_disabled_by_env() && encode_maps(&g, base, hi);  // (quarter turns: plain loads)
        for (const FrameItem &it : items) {
          if ((int)it.geo.layout != lay || it.geo.q8 != g.q8 || (it.geo.out_format == (int)OUT_RGBA8) != g.rgba || it.geo.rotate != g.rot) continue;
          FrameDev f;
          fill_frame_dev(it.desc, it.geo, it.planes, it.pixels, nullptr, 0, &f);
          const uint32_t fi = (uint32_t)frames.size() - g.first_frame;
          frames.push_back(f);
          append_tiles(it, fi, base);
        }
        g.nframes = (uint32_t)frames.size() - g.first_frame;
        g.ntiles = (uint32_t)tiles.size() - g.first_tile;
        if (g.tma) tma_groups++;
        groups.push_back(g);
      }
    n_fused = (uint32_t)frames.size();
    for (const FrameItem &it : items) {
      if (it.geo.layout != LAYOUT_GENERIC) continue;
      FrameDev f;
      fill_frame_dev(it.desc, it.geo, it.planes, it.pixels, it.samples, 0, &f);
      frames.push_back(f);
    }
  }

  int launch_count() const { return (int)groups.size() * launches_fused() + (int)(frames.size() - n_fused) * launches_generic(); }

  // Copies the tables to the device (asynchronously, from pinned staging).
  cudaError_t upload(cudaStream_t s) {
    cudaError_t e;
    if (!d_counters_) {  // tile-claim counters: kCounterSlots pairs per possible launch group; kernels leave them zero
      if ((e = cudaMalloc((void **)&d_counters_, kMaxGroups * kCounterSlots * 2 * sizeof(uint32_t))) != cudaSuccess) return e;
      if ((e = cudaMemsetAsync(d_counters_, 0, kMaxGroups * kCounterSlots * 2 * sizeof(uint32_t), s)) != cudaSuccess) return e;
    }
    if (cap_frames_ < frames.size()) {
      if (d_frames_) cudaFree(d_frames_);
      if (h_frames_) cudaFreeHost(h_frames_);
      d_frames_ = nullptr;
      h_frames_ = nullptr;
      cap_frames_ = 0;
      device_valid_ = false;
      const size_t n = frames.size() + frames.size() / 2 + 1;
      if ((e = cudaMalloc((void **)&d_frames_, n * sizeof(FrameDev))) != cudaSuccess) return e;
      if ((e = cudaHostAlloc((void **)&h_frames_, n * sizeof(FrameDev), cudaHostAllocDefault)) != cudaSuccess) return e;
      cap_frames_ = n;
    }
    if (cap_tiles_ < tiles.size() + 1) {
      if (d_tiles_) cudaFree(d_tiles_);
      if (h_tiles_) cudaFreeHost(h_tiles_);
      d_tiles_ = nullptr;
      h_tiles_ = nullptr;
      cap_tiles_ = 0;
      device_valid_ = false;
      const size_t n = tiles.size() + tiles.size() / 2 + 1;
      if ((e = cudaMalloc((void **)&d_tiles_, n * sizeof(TileDesc))) != cudaSuccess) return e;
      if ((e = cudaHostAlloc((void **)&h_tiles_, n * sizeof(TileDesc), cudaHostAllocDefault)) != cudaSuccess) return e;
      cap_tiles_ = n;
    }
    // A recycled job for a frame like the last one (same geometry, tables, buffers) builds the very
    // same tables: what the device holds is still right, two stream operations saved per frame.
    if (device_valid_ && uploaded_frames_ == frames.size() && uploaded_tiles_ == tiles.size() &&
        memcmp(h_frames_, frames.data(), frames.size() * sizeof(FrameDev)) == 0 &&
        (tiles.empty() || memcmp(h_tiles_, tiles.data(), tiles.size() * sizeof(TileDesc)) == 0)) {
      upload_bytes = 0;
      return cudaSuccess;
    }
    device_valid_ = false;
    memcpy(h_frames_, frames.data(), frames.size() * sizeof(FrameDev));
    if (!tiles.empty()) memcpy(h_tiles_, tiles.data(), tiles.size() * sizeof(TileDesc));
    if ((e = cudaMemcpyAsync(d_frames_, h_frames_, frames.size() * sizeof(FrameDev), cudaMemcpyHostToDevice, s)) != cudaSuccess)
      return e;
    if (!tiles.empty())
      if ((e = cudaMemcpyAsync(d_tiles_, h_tiles_, tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, s)) != cudaSuccess)
        return e;
    upload_bytes = frames.size() * sizeof(FrameDev) + tiles.size() * sizeof(TileDesc);
    uploaded_frames_ = frames.size();
    uploaded_tiles_ = tiles.size();
    device_valid_ = true;
    return cudaSuccess;
  }

  // pdl: programmatic dependent launches (recon_launch.h).  Two launches of the same set may then be
  // on the device at once (the tail of one, the ramp of the next), so consecutive launches take turns
  // on kCounterSlots pairs of claim counters; a third launch cannot start before the first has left the
  // device (it needs every CTA of the second resident or gone, and a launch with dynamic claims fills
  // the device).
  cudaError_t launch(int num_sms, cudaStream_t s, bool pdl = false) {
    cudaError_t e;
    uint32_t gi = 0;
    const uint32_t slot = launch_seq_++ % kCounterSlots;
    for (const LaunchGroup &g : groups) {
      if ((e = launch_fused(g.layout, g.q8, g.rgba, g.rot, g.tma, &g.map2, d_frames_ + g.first_frame,
                            d_tiles_ + g.first_tile, g.ntiles, d_counters_ + 2 * (gi * kCounterSlots + slot), num_sms,
                            pdl, s)) != cudaSuccess)
        return e;
      gi++;
    }
    for (size_t i = n_fused; i < frames.size(); i++)
      if ((e = launch_generic(frames[i], d_frames_ + i, s)) != cudaSuccess) return e;
    return cudaSuccess;
  }

  size_t upload_bytes = 0;

 private:
  static constexpr int kMaxGroups = 16 * kFusedLayouts;  // layouts x 2 table widths x 2 pixel sizes x 4 orientations
  static constexpr uint32_t kCounterSlots = 4;
  uint32_t launch_seq_ = 0;
  uint32_t *d_counters_ = nullptr;
  FrameDev *d_frames_ = nullptr, *h_frames_ = nullptr;
  TileDesc *d_tiles_ = nullptr, *h_tiles_ = nullptr;
  size_t cap_frames_ = 0, cap_tiles_ = 0;
  size_t uploaded_frames_ = 0, uploaded_tiles_ = 0;  // what d_frames_ / d_tiles_ hold (= the pinned staging)
  bool device_valid_ = false;

  static bool encode_maps(LaunchGroup *g, uintptr_t base, uintptr_t hi) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    uint64_t rows = (uint64_t)((hi - base + 127) / 128);
    rows = (rows + 1) & ~(uint64_t)1;
    if (rows >= ((uint64_t)1 << 32)) return false;
    {
      const cuuint64_t dims[2] = {64, rows};
      const cuuint64_t strides[1] = {128};
      const cuuint32_t box[2] = {64, (cuuint32_t)kTileMcus};
      const cuuint32_t estr[2] = {1, 1};
      if (enc(&g->map2, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, reinterpret_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    }
    return true;
  }

  // Tiles of one frame.  coord[] = box row (block index relative to `base`) of each item.
  void append_tiles(const FrameItem &it, uint32_t frame_index, uintptr_t base) {
    const Geometry &g = it.geo;
    uint64_t row0[3] = {0, 0, 0};
    for (int c = 0; c < g.ncomp; c++) row0[c] = (reinterpret_cast<uintptr_t>(it.planes[c]) - base) / 128;
    const uint32_t nmcu = (uint32_t)g.mcux * (uint32_t)g.mcuy;
    TileDesc t;
    memset(&t, 0, sizeof t);
    t.frame = frame_index;
    if (g.rotate & 1) {
      // quarter turns: turn_tile_rows MCUs down x 32 / turn_tile_rows across (mcu_item); the plain-load kernel
      // finds its blocks from the frame descriptor, no coordinates.  Tiles of one column of tiles after each
      // other: they fill the same rows of the turned image side by side.
      const int rows = g.turn_tile_rows, cols = kTileMcus / rows;
      for (int mx0 = 0; mx0 < g.mcux; mx0 += cols)
        for (int my0 = 0; my0 < g.mcuy; my0 += rows) {
          t.m0 = (uint32_t)my0 * (uint32_t)g.mcux + (uint32_t)mx0;
          tiles.push_back(t);
        }
      return;
    }
    if (g.layout == LAYOUT_GREY || g.layout == LAYOUT_444) {
      for (uint32_t m0 = 0; m0 < nmcu; m0 += kTileMcus) {
        t.m0 = m0;
        if (g.layout == LAYOUT_GREY) {
          t.coord[0] = (uint32_t)(row0[0] + m0);
        } else {
          t.coord[0] = (uint32_t)(row0[1] + m0);
          t.coord[1] = (uint32_t)(row0[2] + m0);
          t.coord[2] = (uint32_t)(row0[0] + m0);
        }
        tiles.push_back(t);
      }
      return;
    }
    const int H = layout_h(g.layout), V = layout_v(g.layout);
    for (int my = 0; my < g.mcuy; my++)
      for (int mx0 = 0; mx0 < g.mcux; mx0 += kTileMcus) {
        const uint32_t m0 = (uint32_t)my * (uint32_t)g.mcux + (uint32_t)mx0;
        t.m0 = m0;
        t.coord[0] = (uint32_t)(row0[1] + m0);
        t.coord[1] = (uint32_t)(row0[2] + m0);
        for (int sy = 0; sy < V; sy++)
          for (int part = 0; part < H; part++) {  // 32 consecutive luma blocks per item (mcu_item)
            const uint64_t blk = row0[0] + (uint64_t)(my * V + sy) * (uint64_t)g.bx[0] + (uint64_t)H * (uint64_t)mx0;
            t.coord[2 + sy * H + part] = (uint32_t)(blk + 32u * (uint64_t)part);
          }
        tiles.push_back(t);
      }
  }
};

}  // namespace gidk
