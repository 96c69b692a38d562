// recon_kernels.cu -- sm_100a kernels of the JPEG reconstruction path.
//
// v0 layout: one thread = one MCU, 128 MCUs per CTA ("tile"); a tile is a run
// of consecutive MCUs of one frame in raster order, so neighbouring threads
// read neighbouring coefficient blocks and write neighbouring 24-byte pixel
// runs of the same image rows.
#include "recon_launch.h"

namespace gidk {

namespace {

constexpr int kFrameWords = (int)(sizeof(FrameDev) / sizeof(uint32_t));
static_assert(sizeof(FrameDev) % sizeof(uint32_t) == 0, "FrameDev must be word sized");

template <int LAYOUT, bool Q8>
__global__ void __launch_bounds__(kTileMcus)
recon_fused_kernel(const FrameDev *__restrict__ frames, const uint32_t *__restrict__ tile_frame,
                   uint32_t ntiles) {
  __shared__ __align__(16) uint32_t s_frame_words[kFrameWords];
  const uint32_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  const uint32_t fi = tile_frame ? tile_frame[tile] : 0u;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frames + fi);
    for (int i = threadIdx.x; i < kFrameWords; i += kTileMcus) s_frame_words[i] = src[i];
  }
  __syncthreads();
  const FrameDev &f = *reinterpret_cast<const FrameDev *>(s_frame_words);
  const uint32_t m = (tile - f.tile_base) * kTileMcus + threadIdx.x;
  if (m >= (uint32_t)(f.mcux * f.mcuy)) return;
  const int my = (int)(m / (uint32_t)f.mcux);
  const int mx = (int)(m - (uint32_t)my * (uint32_t)f.mcux);
  const uint32_t *qw = &f.qw[0][0];
  if (LAYOUT == LAYOUT_GREY) mcu_grey<Q8>(f, qw, mx, my);
  else if (LAYOUT == LAYOUT_444) mcu_ycc<1, 1, Q8>(f, qw, mx, my);
  else if (LAYOUT == LAYOUT_422) mcu_ycc<2, 1, Q8>(f, qw, mx, my);
  else mcu_ycc<2, 2, Q8>(f, qw, mx, my);
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
  load_idct_block<Q8>(f.plane[c] + (size_t)idx * 64, &f.qw[c][0], b);
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

template <int LAYOUT>
cudaError_t launch_fused_q(bool q8, const FrameDev *d_frames, const uint32_t *d_tile_frame,
                           uint32_t ntiles, cudaStream_t stream) {
  if (q8)
    recon_fused_kernel<LAYOUT, true><<<ntiles, kTileMcus, 0, stream>>>(d_frames, d_tile_frame, ntiles);
  else
    recon_fused_kernel<LAYOUT, false><<<ntiles, kTileMcus, 0, stream>>>(d_frames, d_tile_frame, ntiles);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fused(Layout layout, bool q8, const FrameDev *d_frames,
                         const uint32_t *d_tile_frame, uint32_t ntiles, cudaStream_t stream) {
  if (ntiles == 0) return cudaSuccess;
  switch (layout) {
    case LAYOUT_GREY: return launch_fused_q<LAYOUT_GREY>(q8, d_frames, d_tile_frame, ntiles, stream);
    case LAYOUT_444: return launch_fused_q<LAYOUT_444>(q8, d_frames, d_tile_frame, ntiles, stream);
    case LAYOUT_422: return launch_fused_q<LAYOUT_422>(q8, d_frames, d_tile_frame, ntiles, stream);
    case LAYOUT_420: return launch_fused_q<LAYOUT_420>(q8, d_frames, d_tile_frame, ntiles, stream);
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
