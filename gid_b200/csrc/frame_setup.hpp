// frame_setup.hpp -- frame geometry and device descriptor set-up, plain C++
// (no CUDA runtime dependency) so that tests can drive the host-compiled
// kernel body with exactly the descriptors the shim builds.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/gidcuda.h"
#include "recon_body.cuh"

namespace gidk {

constexpr int kTileMcusSetup = 32;  // = kTileMcus of recon_launch.h

inline int geo_fail(std::string *err, int code, const char *fmt, ...) {
  if (err) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    *err = buf;
  }
  return code;
}

struct Geometry {
  int ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  int bx[3] = {0, 0, 0}, by[3] = {0, 0, 0};
  size_t plane_bytes[3] = {0, 0, 0};
  size_t plane_off[3] = {0, 0, 0};  // offsets inside one contiguous coefficient buffer
  size_t coef_bytes = 0;
  size_t stride = 0, pixel_bytes = 0;
  size_t sample_off[3] = {0, 0, 0}, sample_bytes = 0;  // generic layout scratch
  Layout layout = LAYOUT_GENERIC;
  bool q8 = true;
  int out_format = 0;
  int rotate = 0;  // 0, 1 = Rotation_90, 2 = Rotation_180, 3 = Rotation_270 (GIDCUDA_FLAG_ROTATE_*)
  uint32_t ntiles = 0;
  int turn_tile_rows = 0;  // fused frames turned by a quarter: MCU rows per tile (32, 16, 8 or 4), see make_geometry
};

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Largest coefficient magnitude of component c for which the fast kernel instances are the reference:
// every |coefficient * table entry| <= 2**16.  Then no sum of Row_IDCT (:813-843: at most
// 18 377 * max|blk| + 256 in magnitude) wraps, blk(4) * 2**11 is zero only for blk(4) = 0, blk(0) * 8 is the
// general path with rounding constant 0, and the row outputs (< 2**23 - 32) keep Col_IDCT's shortcut
// (:858-862) equal to its general path.  JPEGs of 8-bit samples stay far inside (|coefficient * entry| <=
// 1024 + entry / 2 before the encoder's rounding).
inline int32_t fast_coef_limit(const gidcuda_frame_desc *d, int c);
// ... as a bit count: magnitudes below 2**bits are inside the limit (what the Huffman decoders test: a
// value of category n is below 2**n)
inline int32_t fast_coef_bits(const gidcuda_frame_desc *d, int c) {
  const uint32_t lim = (uint32_t)fast_coef_limit(d, c) + 1u;
  int32_t b = 0;
  while ((2u << b) <= lim) b++;
  return b;
}
inline int32_t fast_coef_limit(const gidcuda_frame_desc *d, int c) {
  uint32_t most = 1;
  for (int i = 0; i < 64; i++) most = d->qt[c][i] > most ? d->qt[c][i] : most;
  const uint32_t lim = 65536u / most;
  return (int32_t)(lim > 32767u ? 32767u : lim);
}

// Validates a frame descriptor and derives everything the launch needs.
// Mirrors Read_SOF / Read_SOS geometry (gid-decoding_jpg.adb:244-276, :1945-1951).
// tuning / tests: 0 = choose (default), else force the rows per tile of frames turned by a quarter (32, 16, 8, 4)
inline int &turn_tile_rows_option() {
  static int rows = 0;
  return rows;
}

inline int make_geometry(std::string *err, const gidcuda_frame_desc *d, Geometry *g) {
  if (!d) return geo_fail(err, GIDCUDA_ERR_BAD_ARGUMENT, "null frame descriptor");
  if (d->width < 1 || d->width > 65535 || d->height < 1 || d->height > 65535)
    return geo_fail(err, GIDCUDA_ERR_BAD_ARGUMENT, "frame size %d x %d out of range 1..65535",
                d->width, d->height);
  if (d->ncomp != 1 && d->ncomp != 3)
    return geo_fail(err, GIDCUDA_ERR_UNSUPPORTED,
                "%d components: only Y and YCbCr frames are decoded (gid-decoding_jpg.adb:298-317)",
                d->ncomp);
  g->ncomp = d->ncomp;
  g->hmax = g->vmax = 1;
  for (int c = 0; c < d->ncomp; c++) {
    if (d->samp_h[c] < 1 || d->samp_h[c] > 4 || d->samp_v[c] < 1 || d->samp_v[c] > 4)
      return geo_fail(err, GIDCUDA_ERR_BAD_ARGUMENT, "component %d: sampling factors %d x %d", c,
                  d->samp_h[c], d->samp_v[c]);
    if (d->samp_h[c] > g->hmax) g->hmax = d->samp_h[c];
    if (d->samp_v[c] > g->vmax) g->vmax = d->samp_v[c];
  }
  for (int c = 0; c < d->ncomp; c++)
    if (g->hmax % d->samp_h[c] != 0 || g->vmax % d->samp_v[c] != 0)
      return geo_fail(err, GIDCUDA_ERR_UNSUPPORTED,
                  "component %d: non-integer upsampling factor (GID needs max_samples / samples "
                  "to be exact, gid-decoding_jpg.adb:274-275)", c);
  if (d->ncomp == 1 && (d->samp_h[0] != 1 || d->samp_v[0] != 1))
    return geo_fail(err, GIDCUDA_ERR_UNSUPPORTED,
                "grey frame with sampling factors other than 1x1 (single-component scans are "
                "8x8 MCUs, gid-decoding_jpg.adb:1952-1971)");
  g->mcux = (d->width + 8 * g->hmax - 1) / (8 * g->hmax);
  g->mcuy = (d->height + 8 * g->vmax - 1) / (8 * g->vmax);
  size_t off = 0, soff = 0;
  g->q8 = true;
  for (int c = 0; c < d->ncomp; c++) {
    g->bx[c] = g->mcux * d->samp_h[c];
    g->by[c] = g->mcuy * d->samp_v[c];
    g->plane_bytes[c] = (size_t)g->bx[c] * g->by[c] * 128;
    g->plane_off[c] = off;
    off += round_up(g->plane_bytes[c], 256);
    g->sample_off[c] = soff;
    soff += round_up((size_t)g->bx[c] * g->by[c] * 64, 256);
    // the fast domain of the fused kernels (recon_body.cuh, dequant_idct_sparse): entries 1 .. 255
    for (int i = 0; i < 64; i++)
      if (d->qt[c][i] > 255 || d->qt[c][i] == 0) g->q8 = false;
  }
  if (d->flags & GIDCUDA_FLAG_WIDE_COEF) g->q8 = false;  // ... and coefficients within +-4095
  g->coef_bytes = off;
  g->sample_bytes = soff;
  g->out_format = (int)(d->flags & GIDCUDA_OUT_MASK);
  if (g->out_format != (int)GIDCUDA_OUT_RGB8 && g->out_format != (int)GIDCUDA_OUT_BGR8_BOTTOM_UP &&
      g->out_format != (int)GIDCUDA_OUT_RGBA8)
    return geo_fail(err, GIDCUDA_ERR_BAD_ARGUMENT, "unknown output format %d", g->out_format);
  g->rotate = (int)((d->flags & GIDCUDA_FLAG_ROTATE_MASK) >> 9);
  const int out_w = (g->rotate & 1) ? d->height : d->width, out_h = (g->rotate & 1) ? d->width : d->height;
  const size_t packed = (size_t)(g->out_format == (int)GIDCUDA_OUT_RGBA8 ? 4 : 3) * out_w;
  if (d->out_stride != 0) {
    if ((size_t)d->out_stride < packed)
      return geo_fail(err, GIDCUDA_ERR_BAD_ARGUMENT, "out_stride %d is smaller than a packed row", d->out_stride);
    g->stride = (size_t)d->out_stride;
  } else if (g->out_format == (int)GIDCUDA_OUT_BGR8_BOTTOM_UP) {
    g->stride = round_up(packed, 4);  // BMP rows, test/to_bmp.adb:94-102
  } else if (d->flags & GIDCUDA_FLAG_PACKED_ROWS) {
    g->stride = packed;
  } else {
    g->stride = round_up(packed, 16);
  }
  g->pixel_bytes = g->stride * (size_t)out_h;
  if (g->rotate != 0 && !g->q8) {
    // rotated frames outside the fast domain: the per-pixel pass applies the orientation (the rotating
    // instances of the fused kernels exist for the fast kernels only)
    g->layout = LAYOUT_GENERIC;
  } else if (d->ncomp == 1) {
    g->layout = LAYOUT_GREY;
  } else if (d->samp_h[1] == 1 && d->samp_v[1] == 1 && d->samp_h[2] == 1 && d->samp_v[2] == 1) {
    if (d->samp_h[0] == 1 && d->samp_v[0] == 1) g->layout = LAYOUT_444;
    else if (d->samp_h[0] == 2 && d->samp_v[0] == 1) g->layout = LAYOUT_422;
    else if (d->samp_h[0] == 2 && d->samp_v[0] == 2) g->layout = LAYOUT_420;
    else if (d->samp_h[0] == 1 && d->samp_v[0] == 2) g->layout = LAYOUT_440;
    else if (d->samp_h[0] == 4 && d->samp_v[0] == 1) g->layout = LAYOUT_411;
    else g->layout = LAYOUT_GENERIC;
  } else {
    g->layout = LAYOUT_GENERIC;
  }
  g->turn_tile_rows = 0;
  if (g->layout != LAYOUT_GENERIC && (g->rotate & 1)) {
    // Quarter turns: a tile is `rows` MCUs down times 32 / rows MCUs across (mcu_item).  The shape that costs
    // least: the number of tiles, weighted by what a tile of that shape was measured to cost (its runs along a
    // row of the turned image are rows x 24 bytes: on 1080p 4:2:0 batches a tile of 16 / 8 / 4 rows takes 1.06 /
    // 1.12 / 1.29 times as long as one of 32; profiles/r02_rotations.jsonl).  A frame 32 k MCUs high keeps 32
    // rows; a 1080p 4:2:0 frame is 68 MCUs high -- 32 + 32 + 4 would leave a third of its tiles an eighth full --
    // and takes 8.
    size_t best = 0, best_tiles = 0;
    const int opt = turn_tile_rows_option();
    const int forced = opt == 32 || opt == 16 || opt == 8 || opt == 4 ? opt : 0;
    for (int rows = kTileMcusSetup; rows >= 4; rows /= 2) {
      const int cols = kTileMcusSetup / rows;
      const size_t n = (size_t)((g->mcux + cols - 1) / cols) * (size_t)((g->mcuy + rows - 1) / rows);
      const size_t cost = n * (size_t)(rows == 32 ? 100 : (rows == 16 ? 106 : (rows == 8 ? 112 : 129)));
      if (forced == rows || (forced == 0 && (best == 0 || cost < best))) {
        best = cost;
        best_tiles = n;
        g->turn_tile_rows = rows;
      }
    }
    g->ntiles = (uint32_t)best_tiles;
  } else
    g->ntiles = layout_linear(g->layout) || g->layout == LAYOUT_GENERIC
                    ? (uint32_t)(((size_t)g->mcux * g->mcuy + kTileMcusSetup - 1) / kTileMcusSetup)
                    : (uint32_t)((size_t)g->mcuy * ((g->mcux + kTileMcusSetup - 1) / kTileMcusSetup));
  return GIDCUDA_OK;
}

inline void fill_frame_dev(const gidcuda_frame_desc *d, const Geometry &g, const int16_t *const planes[3],
                    uint8_t *pixels, uint8_t *samples, uint32_t tile_base, FrameDev *f) {
  memset(f, 0, sizeof *f);
  for (int c = 0; c < g.ncomp; c++) {
    f->plane[c] = planes[c];
    f->bx[c] = g.bx[c];
    f->by[c] = g.by[c];
    f->samp_h[c] = d->samp_h[c];
    f->samp_v[c] = d->samp_v[c];
    f->samples[c] = samples ? samples + g.sample_off[c] : nullptr;
    for (int k = 0; k < 32; k++) {
      const uint32_t lo = d->qt[c][2 * k], hi = d->qt[c][2 * k + 1];
      f->qw[c][k] = g.q8 ? (lo | (hi << 24)) : (lo | (hi << 16));
    }
  }
  f->pixels = pixels;
  f->stride = (uint32_t)g.stride;
  f->width = d->width;
  f->height = d->height;
  f->mcux = g.mcux;
  f->mcuy = g.mcuy;
  const bool fused = g.layout != LAYOUT_GENERIC;
  // the generic kernels read the rotation here; the fused kernels are instances per orientation
  f->out_format = fused ? g.out_format : (g.out_format | (g.rotate << 8));
  (void)tile_base;
  f->tile_rows = (uint32_t)g.turn_tile_rows;
  f->q8 = g.q8 ? 1 : 0;
  f->flags = 0;
  const int turn = fused ? g.rotate : 0;
  if (turn == 1) f->flags |= FRAME_ROT90;
  if (turn == 2) f->flags |= FRAME_ROT180;
  if (turn == 3) f->flags |= FRAME_ROT270;
  // (mirrored runs of whole blocks start at column W' - x0 - 8, W' = the frame's width for the half turn,
  // its height for FRAME_ROT270: the same boundaries only when W' % 8 == 0)
  const int mirrored = turn == 2 ? d->width : (turn == 3 ? d->height : 0);
  if (g.stride % 8 == 0 && (reinterpret_cast<uintptr_t>(pixels) & 7u) == 0 && mirrored % 8 == 0)
    f->flags |= FRAME_ALIGNED8;
  if (g.stride % 16 == 0 && (reinterpret_cast<uintptr_t>(pixels) & 15u) == 0 && mirrored % 4 == 0)
    f->flags |= FRAME_ALIGNED16;
}

}  // namespace gidk
