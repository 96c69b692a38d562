// recon_launch.h -- internal interface between the C-ABI shim (gidcuda.cu) and
// the kernels (recon_kernels.cu).  Not installed.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "recon_body.cuh"

namespace gidk {

constexpr int kTileMcus = 32;  // MCUs (= lanes of one warp) per tile of the fused kernels

// One tile = up to kTileMcus consecutive MCUs of one frame, starting at linear
// MCU index m0.  Layouts whose components all have one block per MCU (grey,
// 4:4:4) use linear runs that may wrap over MCU rows; 4:2:2 / 4:2:0 tiles stay
// inside one MCU row, so that every item of the tile is ONE box of the
// coefficient tensor map.  coord[i] = row coordinate of item i for the TMA
// kernel (items: grey {Y}; 4:4:4 {Cb, Cr, Y}; 4:2:x {Cb, Cr, then per luma block
// row sy the two halves of 32 consecutive luma blocks}, see mcu_item).
struct TileDesc {
  uint32_t frame;
  uint32_t m0;
  uint32_t coord[6];
};
static_assert(sizeof(TileDesc) == 32, "TileDesc is 32 bytes");

// Fused dequant + IDCT + upsample + colour over `ntiles` tiles of one layout.
//   use_tma: coefficients are staged through a shared-memory ring by TMA
//            (cp.async.bulk.tensor, 128-byte swizzle); needs map2.
//            Otherwise every thread reads its blocks with 128-bit global loads.
//   map2: [rows][64] int16 view of the coefficient arena, box 128 x 64
//   d_counters: two zero-initialised words of device memory private to this launch
//            group (dynamic tile claims; the kernel leaves them zero again), so one
//            group must not run concurrently with itself.
//   rgba:    4-byte pixels (a separate kernel instance: the 3-byte kernels do not carry its code)
//   rotate:  Geometry::rotate of the frames; rotated frames run instances of their own (fast domain only:
//            q8 must be set)
//   pdl:     launch with programmatic stream serialization: the launch may start while the kernel
//            before it in the stream is still draining.  Only for launches whose inputs do not come
//            from that kernel (plans: frames resident and independent); completion order is kept.
cudaError_t launch_fused(Layout layout, bool q8, bool rgba, int rotate, bool use_tma, const CUtensorMap *map2,
                         const FrameDev *d_frames, const TileDesc *d_tiles,
                         uint32_t ntiles, uint32_t *d_counters, int num_sms, bool pdl, cudaStream_t stream);

// Generic path for one frame (any integer up-factors): pass 1 writes 8-bit
// sample planes (f.samples), pass 2 upsamples + colour-converts.  `h_frame` is
// the host copy of d_frame[0] (for grid sizing).
cudaError_t launch_generic(const FrameDev &h_frame, const FrameDev *d_frame, cudaStream_t stream);

// Sparse coefficient input (expand_kernel.cu): per component, records[first[seq] .. first[seq + 1])
// are the non-zero coefficients of the block with sequence number seq (scan order, see the kernel),
// record = (uint16)value << 16 | natural index.  Writes every block of the dense planes.
struct ExpandArgs {
  const uint32_t *first[3];
  const uint32_t *records[3];
  int16_t *plane[3];
  uint32_t nblocks[3];
  uint32_t count[3];  // records committed: a first[] entry beyond it is clamped (the arrays come from the caller)
  uint32_t bx[3];
  uint32_t mcux;
  int32_t h[3], v[3];
};
cudaError_t launch_expand(const ExpandArgs &a, int ncomp, cudaStream_t stream);
inline int launches_expand() { return 1; }

// Baseline Huffman decoding on the device (entropy_kernel.cu): one thread per restart interval.
struct HuffDev {        // one Huffman table as the device decoder reads it
  uint16_t lut[1024];   // 10-bit prefix -> length << 8 | symbol; 0 = the code is longer than 10 bits
  int32_t maxcode[17];  // [1 .. 16]: largest code of that length, -1 = none
  int32_t valoff[17];   // index of the first symbol of that length - its smallest code
  uint8_t vals[256];
};
static_assert(sizeof(HuffDev) == 2440, "HuffDev layout");
struct ScanArgs {
  const uint8_t *data;          // the scan's entropy-coded bytes
  const uint32_t *begin, *end;  // [nintervals] byte offsets into data: interval i = [begin[i], end[i])
  const HuffDev *tables;        // [6]: DC tables of components 0 .. 2, then their AC tables
  int16_t *plane[3];            // dense planes (block-major, natural order); every block is written whole
  uint32_t bx[3];
  int32_t h[3], v[3];
  int32_t ncomp;
  uint32_t mcux, nmcu, ri, nintervals;
  uint32_t *status;             // [0] flags (1 = an interval was left to the host), [1] how many
  // Magnitudes of component c must stay below 2**limbits[c] (the range of the fast reconstruction
  // instances, frame_setup.hpp fast_coef_limit, rounded down to a power of two; at most 15 so that the
  // DC predictor fits the int16 planes): anything wider is left to the host decoder, which marks the frame.
  int32_t limbits[3];
  uint32_t lanes;               // intervals per warp (set by launch_huffman_intervals)
};
cudaError_t launch_huffman_intervals(const ScanArgs &a, int num_sms, cudaStream_t stream);

// ... and for scans without restart markers: self-synchronising parallel decoding (entropy_kernel.cu)
struct WholeArgs {
  const uint8_t *data;     // the scan's bytes with the stuffing removed, padded with FF
  const HuffDev *tables;   // [6]
  uint64_t *entry, *exit_; // [nsub] decoder state on entering / leaving each subsequence
  int4 *tuple, *prefix;    // [nsub] (blocks completed, DC difference sums) and their exclusive prefix sums
  int16_t *plane[3];
  uint32_t bx[3];
  int32_t h[3], v[3];
  uint32_t bpm;            // blocks per MCU (<= 12)
  uint32_t mcux, total_blocks;
  uint32_t nbits, sub_bits, nsub;
  uint32_t *status;        // [0] flags (2 = chain of states not settled, 8 = undecodable), [2] blocks stored
  uint4 *zero_base;        // the job's coefficient planes, zeroed by the first sync launch together with
  uint32_t zero_count;     // status[] (one stream operation each less per frame); count in 16-byte units
  int32_t limbits[3];      // as in ScanArgs
};
cudaError_t launch_huffman_whole(const WholeArgs &a, cudaStream_t stream);
cudaError_t launch_huffman_finish(const WholeArgs &a, cudaStream_t stream);  // scan + output over the chain as it stands
inline int launches_huffman_whole() { return 4; }

// ... and the scans of PROGRESSIVE frames (entropy_kernel.cu, "progressive scans"): first scans (DC first,
// AC first) through the same self-synchronising decoder, DC refinement (one bit per block), and -- for AC
// refinement, which only the host can decode (see the kernel file) -- the non-zero history of a component for
// the host and the host's verdict applied to the planes.
struct PscanArgs {
  const uint8_t *data;      // the scan's bytes, stuffing removed, padded with FF
  const HuffDev *tables;    // [6]: DC tables of frame components 0 .. 2, then their AC tables
  uint64_t *entry, *exit_;  // [nsub]
  int4 *tuple, *prefix;     // [nsub]
  int16_t *plane[3];        // the frame's planes (accumulating from scan to scan)
  uint32_t row_blocks[3];   // blocks per block row of each plane (MCU-padded)
  int32_t h[3], v[3];       // sampling of the frame's components
  int32_t limbits[3];       // as in ScanArgs
  uint32_t mode;            // 1 = DC first, 2 = AC first, 3 = DC refinement
  uint32_t ncomp, comps[3]; // components of the scan (frame indices)
  uint32_t bpm;             // blocks per MCU of an interleaved scan; 1 for a single-component scan
  uint32_t mcux;            // interleaved: MCUs per row
  uint32_t nbx;             // single component: blocks per row of its OWN grid (8x8 MCUs, :1952-1971)
  uint32_t total_blocks;
  uint32_t ss, se, al;
  uint32_t nbits, sub_bits, nsub;
  uint32_t *status;         // 4 words of this scan: [0] flags (2 chain, 8 undecodable), [2] blocks stored,
                            // [3] 1 + the bit position after the last code word of the scan
};
constexpr uint32_t kPscanWindow = 256, kPscanOutputThreads = 128;  // CTA sizes of the sync and output launches
// A GROUP of first scans (modes 1, 2) in four launches -- sync, sync (windows shifted), prefix sums, output --
// however many scans it holds: d_args[nscans] in device memory, d_tables = three tables of nscans + 1 entries
// each (first CTA of every scan in the first sync launch, in the second, in the output launch).  The scans of a
// group store into different components or bands, so they are independent of each other.
cudaError_t launch_pscan_group(const PscanArgs *d_args, const uint32_t *d_tables, uint32_t nscans, uint32_t ctas_a,
                               uint32_t ctas_b, uint32_t ctas_out, cudaStream_t stream);
cudaError_t launch_pscan_dc_refine(const PscanArgs &a, cudaStream_t stream);   // mode 3
// masks[b], b = block index in the (padded) plane: bit k = coefficient at zig-zag position k is non-zero
cudaError_t launch_pscan_masks(const int16_t *plane, uint32_t nblocks, uint64_t *masks, cudaStream_t stream);
// AC refinement decoded by the host (Ah > 0, :1560-1690): corr = positions that received a correction bit of 1,
// newpos / newneg = positions that became +-(1 << al); all three indexed like masks
cudaError_t launch_pscan_apply(int16_t *plane, uint32_t nblocks, const uint64_t *corr, const uint64_t *newpos,
                               const uint64_t *newneg, uint32_t al, cudaStream_t stream);

// One-time kernel attribute set-up (dynamic shared memory opt-in).
cudaError_t init_kernels();

inline int launches_fused() { return 1; }
inline int launches_generic() { return 2; }

}  // namespace gidk
