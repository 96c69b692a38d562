// expand_kernel.cu -- sparse coefficient records -> dense int16 planes in HBM.
//
// An entropy decoder produces a handful of (position, value) pairs per block
// (Decode_8x8_Block, gid-decoding_jpg.adb:774-786: run / value pairs until EOB).
// Moving them over the host link as records instead of dense 128-byte blocks cuts
// the host-to-device bytes of a frame several-fold (include/gidcuda.h, "sparse
// coefficient input").  This kernel rebuilds the planes the reconstruction kernels
// read: one thread per block zeroes its 128 bytes and stores the block's records.
//
// Block sequence numbers follow the order in which an interleaved scan visits the
// blocks of a component (:1190-1206): MCUs in raster order, inside an MCU the
// component's blocks row by row.
#include "recon_launch.h"

namespace gidk {

namespace {

__global__ void __launch_bounds__(256) expand_records_kernel(ExpandArgs a) {
  const int c = blockIdx.y;
  const uint32_t seq = blockIdx.x * blockDim.x + threadIdx.x;
  if (seq >= a.nblocks[c]) return;
  const uint32_t hv = (uint32_t)(a.h[c] * a.v[c]);
  const uint32_t mcu = seq / hv, r = seq - mcu * hv;
  const uint32_t sy = r / (uint32_t)a.h[c], sx = r - sy * (uint32_t)a.h[c];
  const uint32_t my = mcu / a.mcux, mx = mcu - my * a.mcux;
  const size_t block = (size_t)(my * a.v[c] + sy) * a.bx[c] + (mx * a.h[c] + sx);
  int16_t *dst = a.plane[c] + block * 64;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < 8; i++) reinterpret_cast<uint4 *>(dst)[i] = z;
  // same thread, same addresses: the stores below are ordered after the zeroes
  const uint32_t *rec = a.records[c];
  // first[] is the caller's: whatever it holds, no read leaves the committed records
  const uint32_t end = min(__ldg(a.first[c] + seq + 1), a.count[c]);
  for (uint32_t i = __ldg(a.first[c] + seq); i < end; i++) {
    const uint32_t w = __ldg(rec + i);
    dst[w & 63u] = (int16_t)(w >> 16);
  }
}

}  // namespace

cudaError_t launch_expand(const ExpandArgs &a, int ncomp, cudaStream_t stream) {
  uint32_t most = 0;
  for (int c = 0; c < ncomp; c++) most = a.nblocks[c] > most ? a.nblocks[c] : most;
  if (most == 0) return cudaSuccess;
  const dim3 grid((most + 255) / 256, (unsigned)ncomp);
  expand_records_kernel<<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace gidk
