// recon_launch.h -- internal interface between the C-ABI shim (gidcuda.cu) and
// the kernels (recon_kernels.cu).  Not installed.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "recon_body.cuh"

namespace gidk {

constexpr int kTileMcus = 128;  // MCUs (= threads) per tile of the fused kernels

// Fused dequant + IDCT + upsample + colour kernel over `ntiles` tiles.
//   d_frames     array of frame descriptors (device)
//   d_tile_frame frame index of every tile (device), or nullptr when the launch
//                covers a single frame (index 0)
// Tile t of frame f covers MCUs (t - f.tile_base) * kTileMcus .. + kTileMcus - 1.
cudaError_t launch_fused(Layout layout, bool q8, const FrameDev *d_frames,
                         const uint32_t *d_tile_frame, uint32_t ntiles, cudaStream_t stream);

// Generic path for one frame (any integer up-factors): pass 1 writes 8-bit
// sample planes (f.samples), pass 2 upsamples + colour-converts.  `h_frame` is
// the host copy of d_frame[0] (for grid sizing).
cudaError_t launch_generic(const FrameDev &h_frame, const FrameDev *d_frame, cudaStream_t stream);

// How many kernel launches each of the above performs.
inline int launches_fused() { return 1; }
inline int launches_generic() { return 2; }

}  // namespace gidk
