// emu_recon.cpp -- TEST HARNESS.  Compiles the per-thread kernel body
// (gid_b200/csrc/recon_body.cuh) for the host and runs it for every
// (tile, thread) of a launch, with the frame descriptors the shim builds
// (frame_setup.hpp).  Lets the CPU test-suite check the kernel's arithmetic
// and indexing against the oracle without a GPU.  Never shipped: the product
// library contains device code only and has no CPU path.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../gid_b200/csrc/frame_setup.hpp"

using namespace gidk;

// the layout make_geometry chooses for a frame (LAYOUT_GENERIC = 6: the two-pass kernels), < 0: refused
extern "C" int emu_layout(const gidcuda_frame_desc *desc) {
  Geometry g;
  std::string e;
  const int rc = make_geometry(&e, desc, &g);
  return rc != 0 ? rc : (int)g.layout;
}

// tests: force the rows per tile of frames turned by a quarter (0 = let make_geometry choose)
extern "C" void emu_set_turn_tile_rows(int rows) { turn_tile_rows_option() = rows; }

extern "C" int emu_reconstruct(const gidcuda_frame_desc *desc, const int16_t *const planes[3],
                               uint8_t *pixels, char *err, size_t errlen) {
  Geometry g;
  std::string e;
  int rc = make_geometry(&e, desc, &g);
  if (rc != 0) {
    if (err && errlen) snprintf(err, errlen, "%s", e.c_str());
    return rc;
  }
  std::vector<uint8_t> samples(g.sample_bytes + 16);
  FrameDev f;
  fill_frame_dev(desc, g, planes, pixels, samples.data(), 0, &f);
  const uint32_t *qw = &f.qw[0][0];
  if (g.layout == LAYOUT_GENERIC) {
    // pass 1 (generic_idct_kernel)
    for (int c = 0; c < g.ncomp; c++)
      for (int by = 0; by < f.by[c]; by++)
        for (int bx = 0; bx < f.bx[c]; bx++) {
          int b[64];
          LdgSource src{&f};
          u32x4 raw[8];
          src.load(true, c, bx, by, raw);
          if (f.q8) dequant_idct<true>(raw, &f.qw[c][0], b);
          else dequant_idct<false>(raw, &f.qw[c][0], b);
          uint32_t p[16];
          pack_block(b, p);
          const size_t pitch = (size_t)f.bx[c] * 8;
          uint8_t *dst = f.samples[c] + (size_t)by * 8 * pitch + (size_t)bx * 8;
          for (int r = 0; r < 8; r++) memcpy(dst + r * pitch, &p[2 * r], 8);
        }
    // pass 2 (generic_color_kernel)
    for (int y = 0; y < f.height; y++)
      for (int x = 0; x < f.width; x++) {
        int s[3] = {0, 0, 0};
        for (int c = 0; c < g.ncomp; c++) {
          const int ux = g.hmax / f.samp_h[c], uy = g.vmax / f.samp_v[c];
          s[c] = f.samples[c][(size_t)(y / uy) * ((size_t)f.bx[c] * 8) + (size_t)(x / ux)];
        }
        int r, gg, b;
        if (g.ncomp == 3) {
          const ChromaTerms t = chroma_terms(s[1], s[2]);
          r = clip255(s[0] + t.r); gg = clip255(s[0] + t.g); b = clip255(s[0] + t.b);
        } else {
          r = gg = b = s[0];
        }
        int yy = y;
        if (f.out_format == OUT_BGR8_BOTTOM_UP) { yy = f.height - 1 - y; int t = r; r = b; b = t; }
        const bool rgba = f.out_format == OUT_RGBA8;
        uint8_t *dst = f.pixels + (size_t)yy * f.stride + (size_t)x * (rgba ? 4 : 3);
        dst[0] = (uint8_t)r; dst[1] = (uint8_t)gg; dst[2] = (uint8_t)b;
        if (rgba) dst[3] = 255;
      }
    return 0;
  }
  // Fused layouts: the tiles, items and lanes of recon_tma_kernel / recon_ldg_kernel, run
  // item by item (the lanes of a warp run an item together; in 4:2:x the luma items read the
  // chroma samples another lane left in the warp's scratch).
  LdgSource src{&f};
  u32x2 s_plane[3][8][kTileMcusSetup];
  const bool linear = layout_linear(g.layout);
  const int ni = layout_items(g.layout);
  const bool rgba = f.out_format == OUT_RGBA8;
  const uint32_t nmcu = (uint32_t)(f.mcux * f.mcuy);
  std::vector<uint32_t> tile_m0;
  const bool vert = (f.flags & (FRAME_ROT90 | FRAME_ROT270)) != 0;  // quarter turns: tiles down the columns of MCUs
  if (vert) {
    const int rows = g.turn_tile_rows, cols = kTileMcusSetup / rows;
    for (int mx0 = 0; mx0 < f.mcux; mx0 += cols)
      for (int my0 = 0; my0 < f.mcuy; my0 += rows) tile_m0.push_back((uint32_t)(my0 * f.mcux + mx0));
  } else if (linear) {
    for (uint32_t m0 = 0; m0 < nmcu; m0 += kTileMcusSetup) tile_m0.push_back(m0);
  } else {
    for (int my = 0; my < f.mcuy; my++)
      for (int mx0 = 0; mx0 < f.mcux; mx0 += kTileMcusSetup) tile_m0.push_back((uint32_t)(my * f.mcux + mx0));
  }
  if (tile_m0.size() != g.ntiles) return GIDCUDA_ERR_BAD_ARGUMENT;
  for (uint32_t m0 : tile_m0)
    for (int it = 0; it < ni; it++)
      for (int t = 0; t < kTileMcusSetup; t++) {
        int mx, my;
        bool active;
        if (vert) {
          const int th = (int)f.tile_rows;
          mx = (int)(m0 % (uint32_t)f.mcux) + t / th;
          my = (int)(m0 / (uint32_t)f.mcux) + (t & (th - 1));
          active = mx < f.mcux && my < f.mcuy;
        } else if (linear) {  // tile_mcu of recon_kernels.cu
          const uint32_t m = m0 + (uint32_t)t;
          my = (int)(m / (uint32_t)f.mcux);
          mx = (int)(m - (uint32_t)my * (uint32_t)f.mcux);
          active = m < nmcu;
        } else {
          my = (int)(m0 / (uint32_t)f.mcux);
          mx = (int)(m0 - (uint32_t)my * (uint32_t)f.mcux) + t;
          active = mx < f.mcux;
        }
        Scratch sc{&s_plane[0][0][t], &s_plane[1][0][t], &s_plane[2][0][t], kTileMcusSetup};
#define EMU_RUN2(H, V, GREY, ROT)                                                                          \
  do {                                                                                                     \
    if (f.q8) { if (rgba) mcu_item<H, V, GREY, true, true, ROT>(it, active, f, qw, src, sc, t, mx, my);    \
                else mcu_item<H, V, GREY, true, false, ROT>(it, active, f, qw, src, sc, t, mx, my); }      \
    else { if (rgba) mcu_item<H, V, GREY, false, true, ROT>(it, active, f, qw, src, sc, t, mx, my);        \
           else mcu_item<H, V, GREY, false, false, ROT>(it, active, f, qw, src, sc, t, mx, my); }          \
  } while (0)
#define EMU_RUN(H, V, GREY)                                                                          \
  do {                                                                                               \
    if (f.flags & FRAME_ROT180) EMU_RUN2(H, V, GREY, TURN_180);                                      \
    else if (f.flags & FRAME_ROT90) EMU_RUN2(H, V, GREY, TURN_90);                                   \
    else if (f.flags & FRAME_ROT270) EMU_RUN2(H, V, GREY, TURN_270);                                 \
    else EMU_RUN2(H, V, GREY, TURN_NONE);                                                            \
  } while (0)
        switch (g.layout) {
          case LAYOUT_GREY: EMU_RUN(1, 1, true); break;
          case LAYOUT_444: EMU_RUN(1, 1, false); break;
          case LAYOUT_422: EMU_RUN(2, 1, false); break;
          case LAYOUT_440: EMU_RUN(1, 2, false); break;
          case LAYOUT_411: EMU_RUN(4, 1, false); break;
          default: EMU_RUN(2, 2, false); break;
        }
#undef EMU_RUN
#undef EMU_RUN2
      }
  return 0;
}

// idct on one de-quantised block with the kernel arithmetic; clipped samples out
extern "C" void emu_idct_block(int32_t blk[64]) {
  int b[64];
  for (int i = 0; i < 64; i++) b[i] = blk[i];
  idct_8x8(b);
  for (int i = 0; i < 64; i++) blk[i] = clip255(b[i]);
}

extern "C" void emu_chroma_rgb(int y, int cb, int cr, uint8_t out[3]) {
  const ChromaTerms t = chroma_terms(cb, cr);
  out[0] = (uint8_t)clip255(y + t.r);
  out[1] = (uint8_t)clip255(y + t.g);
  out[2] = (uint8_t)clip255(y + t.b);
}

// exhaustive check of the colour identity against the reference formula
// (gid-decoding_jpg.adb:1028-1035), written out here with truncating "/"
extern "C" int emu_chroma_exhaustive(void) {
  int bad = 0;
  for (int y = 0; y < 256; y++)
    for (int cb = 0; cb < 256; cb++)
      for (int cr = 0; cr < 256; cr++) {
        const int yv = y * 256, cbv = cb - 128, crv = cr - 128;
        const int r = clip255((yv + 359 * crv + 128) / 256);
        const int g = clip255((yv - 88 * cbv - 183 * crv + 128) / 256);
        const int b = clip255((yv + 454 * cbv + 128) / 256);
        const ChromaTerms t = chroma_terms(cb, cr);
        if (clip255(y + t.r) != r || clip255(y + t.g) != g || clip255(y + t.b) != b) bad++;
      }
  return bad;
}
