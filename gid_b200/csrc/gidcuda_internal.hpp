// gidcuda_internal.hpp -- what the translation units of the C-ABI shim share (gidcuda.cu: library,
// jobs, plans; gidcuda_scan.cu: Huffman decoding on the device, chain repair; gidcuda_batch.cu: the
// multi-GPU batch driver).  Not installed.
#pragma once

#include "../../include/gidcuda.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "launch_set.hpp"

using gidk::FrameDev;
using gidk::Geometry;
using gidk::Layout;
using gidk::fill_frame_dev;
using gidk::round_up;

struct gidcuda_ctx {
  int device = -1;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::string last_error;
  std::vector<gidcuda_job *> free_jobs;  // recycled jobs (buffers kept)
  int live_jobs = 0;
};

// Records the message (per thread and per context) and returns `code`.
int fail(gidcuda_ctx *ctx, int code, const char *fmt, ...);

#define CU_TRY(ctx, call)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail((ctx), e_ == cudaErrorMemoryAllocation ? GIDCUDA_ERR_OUT_OF_MEMORY       \
                                                         : GIDCUDA_ERR_CUDA,               \
                  "%s failed: %s", #call, cudaGetErrorString(e_));                         \
  } while (0)

int geometry_or_fail(gidcuda_ctx *ctx, const gidcuda_frame_desc *d, Geometry *g);

// Options and counters (gidcuda_set_option / gidcuda_counter)
extern int g_blocking_wait, g_chain_repair;
extern std::atomic<long> g_chain_repairs, g_chain_repair_subsequences;

struct gidcuda_job {
  gidcuda_ctx *ctx = nullptr;
  gidcuda_frame_desc desc;
  Geometry geo;
  // capacities of the recycled buffers
  size_t cap_coef = 0, cap_pixels = 0, cap_samples = 0, cap_hcoef = 0, cap_sparse = 0;
  uint8_t *h_coef = nullptr;    // pinned: planes, contiguous; allocated at the first gidcuda_job_plane
  bool h_coef_clean = false;    // zeroed since gidcuda_job_begin
  // sparse input: per component [first: nblocks + 1 words][records: capacity words], pinned + device twin
  uint8_t *h_sparse = nullptr, *d_sparse = nullptr;
  size_t sparse_off[3] = {0, 0, 0}, sparse_cap[3] = {0, 0, 0}, sparse_count[3] = {0, 0, 0};
  bool sparse[3] = {false, false, false};  // component committed as records
  // gidcuda_job_records_commit_runs: (begin, count) of each run inside the lent record array;
  // empty = one run starting at 0 (gidcuda_job_records_commit)
  std::vector<std::pair<size_t, size_t>> sparse_runs[3];
  uint8_t *h_pixels = nullptr;  // pinned
  // gidcuda_job_scan: [6 HuffDev][begin][end][scan bytes] pinned + device twin, status words
  uint8_t *h_scan = nullptr, *d_scan = nullptr;
  size_t cap_scan = 0, scan_bytes = 0, scan_off_begin = 0, scan_off_end = 0, scan_off_data = 0;
  uint32_t *h_status = nullptr, *d_status = nullptr;
  bool scan_active = false;
  int32_t scan_ri = 0, scan_nintervals = 0;
  // ... scans without restart markers (scan_ri == 0): the self-synchronising decoder's arrays
  uint8_t *d_sync = nullptr;
  size_t cap_sync = 0, scan_len = 0;
  uint32_t scan_nsub = 0, scan_sub_bits = 0;
  size_t repaired_subsequences = 0;  // subsequences the host walked in the last chain repair
  // gidcuda_job_pscan*: the entropy-coded bytes + code tables of every scan of a progressive frame (pinned +
  // device twin, one slice per scan), four status words per scan, and per component the non-zero masks and
  // the three arrays of an AC refinement (pinned + device twins)
  bool pscan_active = false;
  uint8_t *h_pscan = nullptr, *d_pscan = nullptr;
  size_t cap_pscan = 0, pscan_used = 0;
  uint32_t *h_pstatus = nullptr, *d_pstatus = nullptr;
  uint32_t pscan_count = 0;
  uint32_t pscan_nbits[64] = {0};      // bits of each decoded scan (0: not checked)
  uint32_t pscan_expected[64] = {0};   // blocks a decoded scan must store (0: a DC refinement scan, not counted)
  uint64_t *h_masks = nullptr, *d_masks = nullptr, *h_refine = nullptr, *d_refine = nullptr;
  size_t cap_masks = 0;                // blocks the four arrays above hold (masks: x1, refine: x3)
  cudaEvent_t refine_copied = nullptr; // the host may overwrite h_refine again
  // scans staged by gidcuda_job_pscan and not yet on the stream: ONE host-to-device copy and their launches go
  // out together (pscan_flush), at the next point the host needs the device's answer.  A copy per scan,
  // interleaved with that scan's kernels, would queue behind the copies of other frames' streams in the copy
  // engine's FIFO while those wait for their own kernels: the streams of concurrent frames serialise each other.
  std::vector<gidk::PscanArgs> pscan_pending;
  uint8_t *h_pargs = nullptr, *d_pargs = nullptr;  // the arguments + CTA tables of the groups of a frame (pinned + device)
  size_t pargs_used = 0;
  size_t pscan_flushed = 0;            // bytes of h_pscan already copied
  cudaEvent_t masks_ready = nullptr;   // blocking: the thread sleeps until the device has the history ready
  bool refine_pending = false;
  uint8_t *user_pixels = nullptr;  // gidcuda_job_output: the caller's own pinned buffer receives the pixels
  uint8_t *d_coef = nullptr, *d_pixels = nullptr, *d_samples = nullptr;
  gidk::LaunchSet set;
  cudaEvent_t done = nullptr;
  cudaStream_t stream = nullptr;  // one stream per job: two jobs in flight overlap H2D and D2H
  enum { IDLE, BEGUN, SUBMITTED } state = IDLE;
  int rows_hint[3] = {8, 8, 8};          // gidcuda_job_hint_rows
  bool d_tail_zero[3] = {false, false, false};  // rows 4..7 of every block of the device plane are zero
};

// gidcuda_scan.cu: the scan a job holds (gidcuda_job_scan) -- its stream operations at submit, and at
// wait the verdict of the device decoder (0 = pixels valid; GIDCUDA_ERR_DATA = left to the host decoder),
// the host's chain repair included.
int job_submit_scan(gidcuda_job *job, cudaStream_t s);
int job_submit_scan_status(gidcuda_job *job, cudaStream_t s);  // ... its status words, once the reconstruction is queued
int job_finish_scan(gidcuda_job *job);
// ... and the scans of a progressive frame (gidcuda_job_pscan*): status words back at submit, verdict at wait
int job_submit_pscan(gidcuda_job *job, cudaStream_t s);
int job_submit_pscan_status(gidcuda_job *job, cudaStream_t s);
int job_finish_pscan(gidcuda_job *job);
void job_free_pscan(gidcuda_job *job);
