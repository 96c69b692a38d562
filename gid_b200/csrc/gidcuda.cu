// gidcuda.cu -- C-ABI shim of the B200 reconstruction path (include/gidcuda.h).
//
// Owns pinned host buffers, device buffers, streams and launches; the Ada
// binding GID.CUDA and the C++ host mirror only see plain pointers and sizes.
// No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/gidcuda.h"

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "launch_set.hpp"

using gidk::FrameDev;
using gidk::Layout;

// ---------------------------------------------------------------------------
// errors

static thread_local std::string t_last_error;

struct gidcuda_ctx {
  int device = -1;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::string last_error;
  std::vector<gidcuda_job *> free_jobs;  // recycled jobs (buffers kept)
  int live_jobs = 0;
};

static int fail(gidcuda_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_last_error = buf;
  if (ctx) ctx->last_error = buf;
  return code;
}

#define CU_TRY(ctx, call)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail((ctx), e_ == cudaErrorMemoryAllocation ? GIDCUDA_ERR_OUT_OF_MEMORY       \
                                                         : GIDCUDA_ERR_CUDA,               \
                  "%s failed: %s", #call, cudaGetErrorString(e_));                         \
  } while (0)

// ---------------------------------------------------------------------------
// geometry (frame_setup.hpp)

using gidk::Geometry;
using gidk::fill_frame_dev;
using gidk::round_up;

static int geometry_or_fail(gidcuda_ctx *ctx, const gidcuda_frame_desc *d, Geometry *g) {
  std::string err;
  const int rc = gidk::make_geometry(&err, d, g);
  if (rc != GIDCUDA_OK) return fail(ctx, rc, "%s", err.c_str());
  return rc;
}

// ---------------------------------------------------------------------------
// library / device

// Frames in flight live on their own streams (one per job), and the device runs as many of them side
// by side as it has hardware work queues: 8 by default, which is what bounded small-frame throughput
// (512 x 512 files: 22 000 frames/s whatever the number of host threads, 39 000 with 32 queues; a CUDA
// graph per frame instead of nine stream operations was tried as well and changed nothing).  The
// variable is read when the CUDA context is created, so it is set -- unless the user has set it --
// when this library is loaded.
__attribute__((constructor)) static void gidcuda_more_work_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" int gidcuda_abi_version(void) { return GIDCUDA_ABI_VERSION; }

extern "C" int gidcuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char *gidcuda_last_error(const gidcuda_ctx *ctx) {
  return ctx ? ctx->last_error.c_str() : t_last_error.c_str();
}

extern "C" int gidcuda_create(int device, gidcuda_ctx **out) {
  if (!out) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null ctx pointer");
  *out = nullptr;
  const int n = gidcuda_device_count();
  if (n == 0)
    return fail(nullptr, GIDCUDA_ERR_NO_DEVICE,
                "no CUDA device: the reconstruction path has no CPU fallback");
  if (device < 0 || device >= n)
    return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "device %d out of range 0..%d", device, n - 1);
  gidcuda_ctx *ctx = new (std::nothrow) gidcuda_ctx();
  if (!ctx) return fail(nullptr, GIDCUDA_ERR_OUT_OF_MEMORY, "out of host memory");
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = gidk::init_kernels();
  if (e != cudaSuccess) {
    delete ctx;
    return fail(nullptr, GIDCUDA_ERR_CUDA, "cannot initialise device %d: %s", device,
                cudaGetErrorString(e));
  }
  *out = ctx;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_plane_geometry(const gidcuda_frame_desc *desc, int comp, int32_t *blocks_x,
                                      int32_t *blocks_y) {
  Geometry g;
  int rc = geometry_or_fail(nullptr, desc, &g);
  if (rc != GIDCUDA_OK) return rc;
  if (comp < 0 || comp >= g.ncomp) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "component %d", comp);
  if (blocks_x) *blocks_x = g.bx[comp];
  if (blocks_y) *blocks_y = g.by[comp];
  return GIDCUDA_OK;
}

extern "C" int gidcuda_output_geometry(const gidcuda_frame_desc *desc, size_t *stride, size_t *bytes) {
  Geometry g;
  int rc = geometry_or_fail(nullptr, desc, &g);
  if (rc != GIDCUDA_OK) return rc;
  if (stride) *stride = g.stride;
  if (bytes) *bytes = g.pixel_bytes;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_host_alloc(size_t bytes, void **ptr) {
  if (!ptr) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null pointer");
  *ptr = nullptr;
  if (gidcuda_device_count() == 0) return fail(nullptr, GIDCUDA_ERR_NO_DEVICE, "no CUDA device");
  CU_TRY(nullptr, cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  return GIDCUDA_OK;
}

extern "C" int gidcuda_host_free(void *ptr) {
  if (ptr) CU_TRY(nullptr, cudaFreeHost(ptr));
  return GIDCUDA_OK;
}

// ---------------------------------------------------------------------------
// jobs

// gidcuda_set_option("blocking_wait"): jobs created from now on sleep in gidcuda_job_wait instead
// of spinning -- for callers that run about as many host threads as there are cores.
static int g_blocking_wait = 0;
// gidcuda_set_option("chain_repair"): 1 (default) = an unsettled chain of the self-synchronising Huffman
// decoder is repaired by the host (job_repair_chain); 0 = reported as GIDCUDA_ERR_DATA at once.
static int g_chain_repair = 1;
static std::atomic<long> g_chain_repairs{0}, g_chain_repair_subsequences{0};

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
  uint8_t *user_pixels = nullptr;  // gidcuda_job_output: the caller's own pinned buffer receives the pixels
  uint8_t *d_coef = nullptr, *d_pixels = nullptr, *d_samples = nullptr;
  gidk::LaunchSet set;
  cudaEvent_t done = nullptr;
  cudaStream_t stream = nullptr;  // one stream per job: two jobs in flight overlap H2D and D2H
  enum { IDLE, BEGUN, SUBMITTED } state = IDLE;
  int rows_hint[3] = {8, 8, 8};          // gidcuda_job_hint_rows
  bool d_tail_zero[3] = {false, false, false};  // rows 4..7 of every block of the device plane are zero
};

static void job_free_buffers(gidcuda_job *j) {
  if (j->h_coef) cudaFreeHost(j->h_coef);
  if (j->h_sparse) cudaFreeHost(j->h_sparse);
  if (j->d_sparse) cudaFree(j->d_sparse);
  if (j->h_pixels) cudaFreeHost(j->h_pixels);
  if (j->h_scan) cudaFreeHost(j->h_scan);
  if (j->d_scan) cudaFree(j->d_scan);
  if (j->h_status) cudaFreeHost(j->h_status);
  if (j->d_status) cudaFree(j->d_status);
  if (j->d_sync) cudaFree(j->d_sync);
  j->d_sync = nullptr;
  j->cap_sync = 0;
  j->h_scan = j->d_scan = nullptr;
  j->h_status = j->d_status = nullptr;
  j->cap_scan = 0;
  if (j->d_coef) cudaFree(j->d_coef);
  if (j->d_pixels) cudaFree(j->d_pixels);
  if (j->d_samples) cudaFree(j->d_samples);
  if (j->done) cudaEventDestroy(j->done);
  if (j->stream) cudaStreamDestroy(j->stream);
  j->set.destroy();
  j->stream = nullptr;
  j->h_coef = j->h_pixels = j->d_coef = j->d_pixels = j->d_samples = j->h_sparse = j->d_sparse = nullptr;
  j->done = nullptr;
  j->cap_coef = j->cap_pixels = j->cap_samples = j->cap_hcoef = j->cap_sparse = 0;
}

// The pinned dense planes and the pinned + device record buffers are allocated on first use:
// a decoder uses one input form or the other.
static int job_reserve_planes(gidcuda_job *j) {
  gidcuda_ctx *ctx = j->ctx;
  const Geometry &g = j->geo;
  if (j->cap_hcoef < g.coef_bytes) {
    if (j->h_coef) cudaFreeHost(j->h_coef);
    j->h_coef = nullptr;
    j->cap_hcoef = 0;
    CU_TRY(ctx, cudaHostAlloc((void **)&j->h_coef, g.coef_bytes, cudaHostAllocDefault));
    j->cap_hcoef = g.coef_bytes;
  }
  if (!j->h_coef_clean) {
    // planes must be zero before entropy decode (:771, :1876-1882)
    memset(j->h_coef, 0, g.coef_bytes);
    j->h_coef_clean = true;
  }
  return GIDCUDA_OK;
}

constexpr size_t kSparseRecordsPerBlock = 32;  // capacity: as many bytes as the dense plane

static int job_reserve_sparse(gidcuda_job *j) {
  gidcuda_ctx *ctx = j->ctx;
  const Geometry &g = j->geo;
  size_t off = 0;
  for (int c = 0; c < g.ncomp; c++) {
    const size_t nb = (size_t)g.bx[c] * g.by[c];
    j->sparse_off[c] = off;
    j->sparse_cap[c] = nb * kSparseRecordsPerBlock;
    off += round_up((nb + 1 + j->sparse_cap[c]) * 4, 256);
  }
  if (j->cap_sparse < off) {
    if (j->h_sparse) cudaFreeHost(j->h_sparse);
    if (j->d_sparse) cudaFree(j->d_sparse);
    j->h_sparse = j->d_sparse = nullptr;
    j->cap_sparse = 0;
    CU_TRY(ctx, cudaHostAlloc((void **)&j->h_sparse, off, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&j->d_sparse, off));
    j->cap_sparse = off;
  }
  return GIDCUDA_OK;
}

static int job_reserve(gidcuda_job *j) {
  gidcuda_ctx *ctx = j->ctx;
  const Geometry &g = j->geo;
  if (!j->done)
    CU_TRY(ctx, cudaEventCreateWithFlags(&j->done, cudaEventDisableTiming | (g_blocking_wait ? cudaEventBlockingSync : 0)));
  if (!j->stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking));
  if (j->cap_coef < g.coef_bytes) {
    if (j->d_coef) cudaFree(j->d_coef);
    j->d_coef = nullptr;
    j->cap_coef = 0;
    CU_TRY(ctx, cudaMalloc((void **)&j->d_coef, g.coef_bytes));
    j->cap_coef = g.coef_bytes;
    for (int c = 0; c < 3; c++) j->d_tail_zero[c] = false;
  }
  if (j->cap_pixels < g.pixel_bytes) {
    if (j->h_pixels) cudaFreeHost(j->h_pixels);
    if (j->d_pixels) cudaFree(j->d_pixels);
    j->h_pixels = j->d_pixels = nullptr;
    j->cap_pixels = 0;
    // (the pinned host twin is allocated at submit, unless the caller supplies its own: gidcuda_job_output)
    CU_TRY(ctx, cudaMalloc((void **)&j->d_pixels, g.pixel_bytes));
    j->cap_pixels = g.pixel_bytes;
  }
  if (g.layout == gidk::LAYOUT_GENERIC && j->cap_samples < g.sample_bytes) {
    if (j->d_samples) cudaFree(j->d_samples);
    j->d_samples = nullptr;
    j->cap_samples = 0;
    CU_TRY(ctx, cudaMalloc((void **)&j->d_samples, g.sample_bytes));
    j->cap_samples = g.sample_bytes;
  }
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_begin(gidcuda_ctx *ctx, const gidcuda_frame_desc *desc, gidcuda_job **out) {
  if (!ctx || !out) return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "null argument");
  *out = nullptr;
  Geometry g;
  int rc = geometry_or_fail(ctx, desc, &g);
  if (rc != GIDCUDA_OK) return rc;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  gidcuda_job *j = nullptr;
  // recycle the free job whose buffers fit
  for (size_t i = 0; i < ctx->free_jobs.size(); i++) {
    gidcuda_job *c = ctx->free_jobs[i];
    if (c->cap_coef >= g.coef_bytes && c->cap_pixels >= g.pixel_bytes) {
      j = c;
      ctx->free_jobs.erase(ctx->free_jobs.begin() + (long)i);
      break;
    }
  }
  if (!j && !ctx->free_jobs.empty()) {
    j = ctx->free_jobs.back();
    ctx->free_jobs.pop_back();
  }
  if (!j) {
    j = new (std::nothrow) gidcuda_job();
    if (!j) return fail(ctx, GIDCUDA_ERR_OUT_OF_MEMORY, "out of host memory");
  }
  j->ctx = ctx;
  j->desc = *desc;
  j->geo = g;
  rc = job_reserve(j);
  if (rc != GIDCUDA_OK) {
    job_free_buffers(j);
    delete j;
    return rc;
  }
  // launch tables for this frame (device addresses are fixed from here on)
  gidk::FrameItem it;
  it.desc = &j->desc;
  it.geo = g;
  for (int c = 0; c < 3; c++)
    it.planes[c] = c < g.ncomp ? reinterpret_cast<const int16_t *>(j->d_coef + g.plane_off[c]) : nullptr;
  it.pixels = j->d_pixels;
  it.samples = j->d_samples;
  j->set.build(std::vector<gidk::FrameItem>(1, it));
  j->h_coef_clean = false;  // zeroed when the decoder asks for the planes
  j->user_pixels = nullptr;
  j->scan_active = false;
  for (int c = 0; c < 3; c++) {
    j->rows_hint[c] = 8;
    j->sparse[c] = false;
    j->sparse_count[c] = 0;
    j->sparse_runs[c].clear();
  }
  j->state = gidcuda_job::BEGUN;
  ctx->live_jobs++;
  *out = j;
  return GIDCUDA_OK;
}

extern "C" int16_t *gidcuda_job_plane(gidcuda_job *job, int comp, int32_t *blocks_x, int32_t *blocks_y) {
  if (!job || comp < 0 || comp >= job->geo.ncomp || job->state == gidcuda_job::IDLE) {
    fail(job ? job->ctx : nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_plane: bad job or component");
    return nullptr;
  }
  if (job_reserve_planes(job) != GIDCUDA_OK) return nullptr;
  if (blocks_x) *blocks_x = job->geo.bx[comp];
  if (blocks_y) *blocks_y = job->geo.by[comp];
  return reinterpret_cast<int16_t *>(job->h_coef + job->geo.plane_off[comp]);
}

extern "C" int gidcuda_job_records(gidcuda_job *job, int comp, uint32_t **first, uint32_t **records,
                                   size_t *capacity, int32_t *nblocks) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  if (job->state == gidcuda_job::IDLE) return fail(job->ctx, GIDCUDA_ERR_STATE, "gidcuda_job_records: job was released");
  if (comp < 0 || comp >= job->geo.ncomp || !first || !records)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_records: bad component or null argument");
  if (job->geo.layout == gidk::LAYOUT_GENERIC && job->geo.ncomp == 3 &&
      (job->desc.samp_h[1] != 1 || job->desc.samp_v[1] != 1 || job->desc.samp_h[2] != 1 || job->desc.samp_v[2] != 1)) {
    // (block sequence numbers are defined for every layout; nothing to reject -- kept for clarity)
  }
  const int rc = job_reserve_sparse(job);
  if (rc != GIDCUDA_OK) return rc;
  const size_t nb = (size_t)job->geo.bx[comp] * job->geo.by[comp];
  uint32_t *base = reinterpret_cast<uint32_t *>(job->h_sparse + job->sparse_off[comp]);
  *first = base;
  *records = base + nb + 1;
  if (capacity) *capacity = job->sparse_cap[comp];
  if (nblocks) *nblocks = (int32_t)nb;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_records_commit(gidcuda_job *job, int comp, size_t count) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  if (job->state == gidcuda_job::IDLE)
    return fail(job->ctx, GIDCUDA_ERR_STATE, "gidcuda_job_records_commit: job was released");
  if (comp < 0 || comp >= job->geo.ncomp || !job->h_sparse)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_records_commit: bad component, or gidcuda_job_records was not called");
  if (count > job->sparse_cap[comp])
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_records_commit: %zu records exceed the capacity %zu", count,
                job->sparse_cap[comp]);
  const size_t nb = (size_t)job->geo.bx[comp] * job->geo.by[comp];
  const uint32_t *first = reinterpret_cast<const uint32_t *>(job->h_sparse + job->sparse_off[comp]);
  if (first[0] != 0 || first[nb] != count)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT,
                "gidcuda_job_records_commit: first[0] = %u, first[nblocks] = %u do not frame %zu records", first[0], first[nb],
                count);
  job->sparse[comp] = true;
  job->sparse_count[comp] = count;
  job->sparse_runs[comp].clear();
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_records_commit_runs(gidcuda_job *job, int comp, int32_t nruns, const size_t *run_begin,
                                               const size_t *run_count) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  if (job->state == gidcuda_job::IDLE)
    return fail(job->ctx, GIDCUDA_ERR_STATE, "gidcuda_job_records_commit_runs: job was released");
  if (comp < 0 || comp >= job->geo.ncomp || !job->h_sparse || nruns < 1 || !run_begin || !run_count)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT,
                "gidcuda_job_records_commit_runs: bad component or runs, or gidcuda_job_records was not called");
  const size_t cap = job->sparse_cap[comp];
  size_t total = 0, floor_ = 0;
  for (int32_t r = 0; r < nruns; r++) {
    // runs are ascending and disjoint, so that the compacted copy never outgrows the device twin
    if (run_begin[r] < floor_ || run_count[r] > cap || run_begin[r] > cap - run_count[r])
      return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT,
                  "gidcuda_job_records_commit_runs: run %d (begin %zu, count %zu) overlaps its predecessor or exceeds the capacity %zu",
                  (int)r, run_begin[r], run_count[r], cap);
    floor_ = run_begin[r] + run_count[r];
    total += run_count[r];
  }
  const size_t nb = (size_t)job->geo.bx[comp] * job->geo.by[comp];
  const uint32_t *first = reinterpret_cast<const uint32_t *>(job->h_sparse + job->sparse_off[comp]);
  if (first[0] != 0 || first[nb] != total)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT,
                "gidcuda_job_records_commit_runs: first[0] = %u, first[nblocks] = %u do not frame %zu records", first[0],
                first[nb], total);
  job->sparse[comp] = true;
  job->sparse_count[comp] = total;
  job->sparse_runs[comp].clear();
  for (int32_t r = 0; r < nruns; r++) job->sparse_runs[comp].emplace_back(run_begin[r], run_count[r]);
  return GIDCUDA_OK;
}

// Canonical code of a DHT table (Read_DHT, :320-391) in the form the device decoder reads.
static bool build_huff_dev(const gidcuda_huffman_table &t, gidk::HuffDev *d) {
  memset(d, 0, sizeof *d);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; len++) {
    d->valoff[len] = k - code;
    for (int i = 0; i < t.counts[len - 1]; i++, k++, code++) {
      if (k > 255 || code >= (1 << len)) return false;  // more symbols than a table holds / over-subscribed
      d->vals[k] = t.symbols[k];
      if (len <= 10) {
        const int first = code << (10 - len), n = 1 << (10 - len);
        for (int j = 0; j < n; j++) d->lut[first + j] = (uint16_t)((len << 8) | t.symbols[k]);
      }
    }
    d->maxcode[len] = t.counts[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  return true;
}

// One device thread per this many bits of a scan without restart markers.  Short subsequences mean
// more threads and shorter rounds; but a decoder only re-synchronises at block boundaries, so where
// blocks are long (high qualities: hundreds of bits each) short subsequences would need many rounds
// to carry the true state across a window -- those scans get longer subsequences.
constexpr uint32_t kSubsequenceBits = 256, kSubsequenceBitsDense = 1024;
constexpr uint32_t kDenseBitsPerBlock = 96;

// gidcuda_job_scan with restart_interval == 0: the whole scan, byte stuffing removed by the caller.
static int job_scan_whole(gidcuda_job *job, const gidcuda_scan_desc *scan, const uint8_t *data, size_t len) {
  gidcuda_ctx *ctx = job->ctx;
  const Geometry &g = job->geo;
  int bpm = 0;
  for (int c = 0; c < g.ncomp; c++) bpm += job->desc.samp_h[c] * job->desc.samp_v[c];
  if (bpm > 12) return fail(ctx, GIDCUDA_ERR_UNSUPPORTED, "gidcuda_job_scan: %d blocks per MCU", bpm);
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t off_data = round_up(6 * sizeof(gidk::HuffDev), 256);
  const size_t total = off_data + round_up(len + 64, 256);  // FF padding: the decoder looks a few bytes ahead
  if (job->cap_scan < total) {
    if (job->h_scan) cudaFreeHost(job->h_scan);
    if (job->d_scan) cudaFree(job->d_scan);
    job->h_scan = job->d_scan = nullptr;
    job->cap_scan = 0;
    const size_t cap = total + total / 4;
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_scan, cap, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_scan, cap));
    job->cap_scan = cap;
  }
  if (!job->h_status) {
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_status, 16, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_status, 16));
  }
  uint64_t frame_blocks = 0;
  for (int c = 0; c < g.ncomp; c++) frame_blocks += (uint64_t)g.bx[c] * g.by[c];
  const uint32_t sub_bits = (uint64_t)len * 8 > frame_blocks * kDenseBitsPerBlock ? kSubsequenceBitsDense : kSubsequenceBits;
  const uint32_t nsub = (uint32_t)((len * 8 + sub_bits - 1) / sub_bits);
  const size_t sync_bytes = (size_t)nsub * (8 + 8 + 16 + 16) + 1024;
  if (job->cap_sync < sync_bytes) {
    if (job->d_sync) cudaFree(job->d_sync);
    job->d_sync = nullptr;
    job->cap_sync = 0;
    const size_t cap = sync_bytes + sync_bytes / 4;
    CU_TRY(ctx, cudaMalloc((void **)&job->d_sync, cap));
    job->cap_sync = cap;
  }
  gidk::HuffDev *tabs = reinterpret_cast<gidk::HuffDev *>(job->h_scan);
  for (int c = 0; c < 3; c++) {
    const bool used = c < g.ncomp;
    if (!build_huff_dev(scan->dc[used ? c : 0], &tabs[c]) || !build_huff_dev(scan->ac[used ? c : 0], &tabs[3 + c]))
      return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_scan: Huffman table of component %d is not a prefix code", c);
  }
  memcpy(job->h_scan + off_data, data, len);
  memset(job->h_scan + off_data + len, 0xFF, total - off_data - len);
  job->scan_off_data = off_data;
  job->scan_bytes = total;
  job->scan_len = len;
  job->scan_nsub = nsub;
  job->scan_sub_bits = sub_bits;
  job->scan_ri = 0;
  job->scan_nintervals = 0;
  job->scan_active = true;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_scan(gidcuda_job *job, const gidcuda_scan_desc *scan, const uint8_t *data, size_t len,
                                const uint32_t *begin, const uint32_t *end) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::BEGUN)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_scan: only between gidcuda_job_begin and gidcuda_job_submit");
  if (!scan || !data || len == 0 || len >= ((size_t)1 << 28) || (scan->restart_interval != 0 && (!begin || !end)))
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_scan: null argument, no data or more than 256 MiB of it");
  const Geometry &g = job->geo;
  const int64_t nmcu = (int64_t)g.mcux * g.mcuy;
  if (scan->restart_interval == 0) return job_scan_whole(job, scan, data, len);
  if (scan->restart_interval <= 0 || scan->nintervals <= 0 ||
      (nmcu + scan->restart_interval - 1) / scan->restart_interval != scan->nintervals)
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_scan: %d intervals of %d MCUs do not cover %lld MCUs",
                (int)scan->nintervals, (int)scan->restart_interval, (long long)nmcu);
  const size_t nint = (size_t)scan->nintervals;
  for (size_t i = 0; i < nint; i++)
    if (begin[i] > end[i] || end[i] > len)
      return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_scan: interval %zu = [%u, %u) outside the %zu bytes", i,
                  begin[i], end[i], len);
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t off_begin = round_up(6 * sizeof(gidk::HuffDev), 256);
  const size_t off_end = off_begin + round_up(nint * 4, 256);
  const size_t off_data = off_end + round_up(nint * 4, 256);
  const size_t total = off_data + round_up(len, 256);
  if (job->cap_scan < total) {
    if (job->h_scan) cudaFreeHost(job->h_scan);
    if (job->d_scan) cudaFree(job->d_scan);
    job->h_scan = job->d_scan = nullptr;
    job->cap_scan = 0;
    const size_t cap = total + total / 4;  // the next frame of a batch is rarely the same size
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_scan, cap, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_scan, cap));
    job->cap_scan = cap;
  }
  if (!job->h_status) {
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_status, 16, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_status, 16));
  }
  gidk::HuffDev *tabs = reinterpret_cast<gidk::HuffDev *>(job->h_scan);
  for (int c = 0; c < 3; c++) {
    const bool used = c < g.ncomp;
    if (!build_huff_dev(scan->dc[used ? c : 0], &tabs[c]) || !build_huff_dev(scan->ac[used ? c : 0], &tabs[3 + c]))
      return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_scan: Huffman table of component %d is not a prefix code", c);
  }
  memcpy(job->h_scan + off_begin, begin, nint * 4);
  memcpy(job->h_scan + off_end, end, nint * 4);
  memcpy(job->h_scan + off_data, data, len);
  job->scan_off_begin = off_begin;
  job->scan_off_end = off_end;
  job->scan_off_data = off_data;
  job->scan_bytes = total;
  job->scan_ri = scan->restart_interval;
  job->scan_nintervals = scan->nintervals;
  job->scan_active = true;
  return GIDCUDA_OK;
}

// The self-synchronising decoder's arguments for the scan a job holds (gidcuda_job_scan, restart_interval = 0).
static void make_whole_args(gidcuda_job *job, gidk::WholeArgs *out) {
  const Geometry &g = job->geo;
  gidk::WholeArgs &wa = *out;
  memset(&wa, 0, sizeof wa);
  wa.data = job->d_scan + job->scan_off_data;
  wa.tables = reinterpret_cast<const gidk::HuffDev *>(job->d_scan);
  const size_t n = job->scan_nsub;
  wa.entry = reinterpret_cast<uint64_t *>(job->d_sync);
  wa.exit_ = wa.entry + n;
  wa.tuple = reinterpret_cast<int4 *>(job->d_sync + round_up(16 * n, 256));
  wa.prefix = wa.tuple + n;
  uint32_t k = 0, blocks = 0;
  for (int c = 0; c < g.ncomp; c++) {
    wa.plane[c] = reinterpret_cast<int16_t *>(job->d_coef + g.plane_off[c]);
    wa.bx[c] = (uint32_t)g.bx[c];
    wa.h[c] = job->desc.samp_h[c];
    wa.v[c] = job->desc.samp_v[c];
    blocks += (uint32_t)g.bx[c] * (uint32_t)g.by[c];
    k += (uint32_t)(wa.h[c] * wa.v[c]);
    job->d_tail_zero[c] = false;
  }
  wa.bpm = k;
  wa.mcux = (uint32_t)g.mcux;
  wa.total_blocks = blocks;
  wa.nbits = (uint32_t)(job->scan_len * 8);
  wa.sub_bits = job->scan_sub_bits;
  wa.nsub = job->scan_nsub;
  wa.status = job->d_status;
  wa.zero_base = reinterpret_cast<uint4 *>(job->d_coef);
  wa.zero_count = (uint32_t)(g.coef_bytes / 16);
}

// ---- chain repair on the host -------------------------------------------------------------------
// A stretch of the scan that does not re-synchronise (a flat region: the same two code words over and
// over, so a decoder in the wrong phase stays there) leaves the device's chain of states broken: truth
// only advances one subsequence per round there.  Walking such a stretch is a SERIAL job, which one host
// core does an order of magnitude faster than one GPU thread.  So when the output pass reports an
// unsettled chain (and nothing undecodable), gidcuda_job_wait fetches the chain, walks it from the true
// start state -- where a subsequence's recorded entry state is the true one its recorded exit and counts
// are kept, where not the host decodes that subsequence (the mirror of run_subsequence<false> in
// entropy_kernel.cu) -- sends it back and runs prefix sums, output pass and reconstruction again.  The
// output pass verifies the repaired chain like any other.
namespace {

struct HostWalk {
  const uint8_t *data;  // unstuffed scan bytes, FF-padded (the job's pinned staging copy)
  const gidk::HuffDev *tabs;
  uint32_t bpm;
  uint8_t comp[12];
};

inline uint32_t host_peek(const uint8_t *d, uint32_t pos, int n) {  // n <= 16
  const uint8_t *p = d + (pos >> 3);
  const uint32_t w = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  return (w << (pos & 7u)) >> (32 - n);
}

inline int host_extend(uint32_t v, int n) { return v < (1u << (n - 1)) ? (int)v + 1 - (1 << n) : (int)v; }

uint64_t host_run_subsequence(const HostWalk &w, uint64_t entry, uint32_t limit, int32_t tup[4]) {
  uint32_t pos = (uint32_t)entry, c = (uint32_t)(entry >> 32) & 0xFFu, zpos = (uint32_t)(entry >> 40) & 0xFFu;
  int nb = 0, dcs[3] = {0, 0, 0};
  while (pos < limit) {
    const int k = w.comp[c];
    const gidk::HuffDev &t = zpos == 0 ? w.tabs[k] : w.tabs[3 + k];
    const uint32_t look = host_peek(w.data, pos, 16);
    const uint32_t e = t.lut[look >> 6];
    int l = (int)(e >> 8), sym = (int)(e & 0xFFu);
    if (e == 0) {
      sym = -1;
      for (l = 11; l <= 16; l++) {
        const int code = (int)(look >> (16 - l));
        if (code <= t.maxcode[l]) {
          sym = t.vals[(t.valoff[l] + code) & 0xFF];
          break;
        }
      }
      if (sym < 0) {
        pos += 1;
        continue;
      }
    }
    pos += (uint32_t)l;
    const int n = sym & 15;
    bool block_done = false;
    if (zpos == 0) {
      if (n) {
        dcs[k] += host_extend(host_peek(w.data, pos, n), n);
        pos += (uint32_t)n;
      }
      zpos = 1;
    } else if (sym == 0) {
      block_done = true;
    } else {
      const uint32_t idx = zpos + (uint32_t)(sym >> 4);
      pos += (uint32_t)n;
      if (idx >= 63) block_done = true;
      else zpos = idx + 1;
    }
    if (block_done) {
      zpos = 0;
      c = c + 1 == w.bpm ? 0 : c + 1;
      nb++;
    }
  }
  tup[0] = nb; tup[1] = dcs[0]; tup[2] = dcs[1]; tup[3] = dcs[2];
  return (uint64_t)pos | ((uint64_t)c << 32) | ((uint64_t)zpos << 40);
}

}  // namespace

// Returns GIDCUDA_OK when the repaired chain has been re-submitted (the caller waits for job->done again).
static int job_repair_chain(gidcuda_job *job) {
  gidcuda_ctx *ctx = job->ctx;
  const Geometry &g = job->geo;
  cudaStream_t s = job->stream;
  gidk::WholeArgs wa;
  make_whole_args(job, &wa);
  const size_t n = wa.nsub;
  std::vector<uint64_t> st(2 * n);   // entry[], exit[]
  std::vector<int32_t> tup(4 * n);
  CU_TRY(ctx, cudaMemcpyAsync(st.data(), wa.entry, 16 * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, cudaMemcpyAsync(tup.data(), wa.tuple, 16 * n, cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, cudaStreamSynchronize(s));
  HostWalk w;
  w.data = job->h_scan + job->scan_off_data;
  w.tabs = reinterpret_cast<const gidk::HuffDev *>(job->h_scan);
  w.bpm = wa.bpm;
  uint32_t k = 0;
  for (int c = 0; c < g.ncomp; c++)
    for (int i = 0; i < job->desc.samp_h[c] * job->desc.samp_v[c]; i++) w.comp[k++] = (uint8_t)c;
  uint64_t state = 0;  // the true start: bit 0, first block of the MCU, DC next
  uint64_t blocks = 0;
  size_t walked = 0;
  for (size_t i = 0; i < n && blocks < wa.total_blocks; i++) {
    if (st[i] != state) {
      const uint32_t limit = std::min<uint64_t>((uint64_t)(i + 1) * wa.sub_bits, wa.nbits);
      st[i] = state;
      st[n + i] = host_run_subsequence(w, state, limit, &tup[4 * i]);
      walked++;
    }
    state = st[n + i];
    blocks += (uint64_t)tup[4 * i];
  }
  job->repaired_subsequences = walked;
  CU_TRY(ctx, cudaMemcpyAsync(wa.entry, st.data(), 16 * n, cudaMemcpyHostToDevice, s));
  CU_TRY(ctx, cudaMemcpyAsync(wa.tuple, tup.data(), 16 * n, cudaMemcpyHostToDevice, s));
  CU_TRY(ctx, cudaMemsetAsync(job->d_coef, 0, g.coef_bytes, s));
  CU_TRY(ctx, cudaMemsetAsync(job->d_status, 0, 16, s));
  CU_TRY(ctx, gidk::launch_huffman_finish(wa, s));
  CU_TRY(ctx, cudaMemcpyAsync(job->h_status, job->d_status, 16, cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, job->set.launch(ctx->num_sms, s));
  CU_TRY(ctx, cudaMemcpyAsync(job->user_pixels ? job->user_pixels : job->h_pixels, job->d_pixels, g.pixel_bytes,
                              cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, cudaStreamSynchronize(s));  // (st / tup are this function's: the copies must be done before they go)
  return GIDCUDA_OK;
}

static bool is_pinned_host(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

extern "C" int gidcuda_host_is_pinned(const void *ptr) { return ptr && is_pinned_host(ptr) ? 1 : 0; }

extern "C" int gidcuda_job_output(gidcuda_job *job, uint8_t *host_pixels, size_t capacity) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  if (job->state != gidcuda_job::BEGUN)
    return fail(job->ctx, GIDCUDA_ERR_STATE, "gidcuda_job_output: only between gidcuda_job_begin and gidcuda_job_submit");
  if (!host_pixels) {
    job->user_pixels = nullptr;
    return GIDCUDA_OK;
  }
  if (capacity < job->geo.pixel_bytes)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_output: %zu bytes, the frame needs %zu", capacity,
                job->geo.pixel_bytes);
  if (!is_pinned_host(host_pixels))
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT,
                "gidcuda_job_output: the buffer is not pinned host memory (gidcuda_host_alloc / cudaHostRegister)");
  job->user_pixels = host_pixels;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_hint_rows(gidcuda_job *job, int comp, int rows_used) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  if (job->state == gidcuda_job::IDLE) return fail(job->ctx, GIDCUDA_ERR_STATE, "gidcuda_job_hint_rows: job was released");
  if (comp < 0 || comp >= job->geo.ncomp || rows_used < 1 || rows_used > 8)
    return fail(job->ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_hint_rows: component %d, rows %d", comp, rows_used);
  job->rows_hint[comp] = rows_used;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_submit(gidcuda_job *job) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  // A job may be submitted again after a wait (same planes, or planes the caller
  // has updated in place): the work is stream-ordered behind the previous run.
  if (job->state == gidcuda_job::IDLE)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_submit: job was released");
  const Geometry &g = job->geo;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = job->stream;
  if (job->state == gidcuda_job::BEGUN) CU_TRY(ctx, job->set.upload(s));  // tables: once per begin
  // Host -> device, per component:
  //   records (gidcuda_job_records_commit): first[] + the records used, then the expand kernel;
  //   dense plane: one copy -- or, when the producer promised empty rows 4 .. 7 and the plane is
  //   large, a pitched copy of the upper 64 bytes of every block (measured 1.5x faster than the
  //   full plane on this link; narrower or wider partial rows are not faster than a full copy,
  //   and below a few MiB the extra calls cost more than the bytes saved: measured on 1080p).
  if (job->scan_active) {
    // compressed bytes -> device, planes zeroed, one thread per restart interval decodes
    CU_TRY(ctx, cudaMemcpyAsync(job->d_scan, job->h_scan, job->scan_bytes, cudaMemcpyHostToDevice, s));
    if (job->scan_ri != 0) {  // (the self-synchronising decoder's first launch zeroes both itself)
      CU_TRY(ctx, cudaMemsetAsync(job->d_coef, 0, g.coef_bytes, s));
      CU_TRY(ctx, cudaMemsetAsync(job->d_status, 0, 16, s));
    }
    if (job->scan_ri == 0) {
      gidk::WholeArgs wa;
      make_whole_args(job, &wa);
      CU_TRY(ctx, gidk::launch_huffman_whole(wa, s));
    } else {
      gidk::ScanArgs sa;
      memset(&sa, 0, sizeof sa);
      sa.data = job->d_scan + job->scan_off_data;
      sa.begin = reinterpret_cast<const uint32_t *>(job->d_scan + job->scan_off_begin);
      sa.end = reinterpret_cast<const uint32_t *>(job->d_scan + job->scan_off_end);
      sa.tables = reinterpret_cast<const gidk::HuffDev *>(job->d_scan);
      for (int c = 0; c < g.ncomp; c++) {
        sa.plane[c] = reinterpret_cast<int16_t *>(job->d_coef + g.plane_off[c]);
        sa.bx[c] = (uint32_t)g.bx[c];
        sa.h[c] = job->desc.samp_h[c];
        sa.v[c] = job->desc.samp_v[c];
        job->d_tail_zero[c] = false;
      }
      sa.ncomp = g.ncomp;
      sa.mcux = (uint32_t)g.mcux;
      sa.nmcu = (uint32_t)g.mcux * (uint32_t)g.mcuy;
      sa.ri = (uint32_t)job->scan_ri;
      sa.nintervals = (uint32_t)job->scan_nintervals;
      sa.status = job->d_status;
      CU_TRY(ctx, gidk::launch_huffman_intervals(sa, s));
    }
    CU_TRY(ctx, cudaMemcpyAsync(job->h_status, job->d_status, 16, cudaMemcpyDeviceToHost, s));
  }
  constexpr size_t kMinPitchedPlane = (size_t)8 << 20;
  bool any_sparse = false, all_dense_plain = true;
  for (int c = 0; c < g.ncomp; c++) {
    any_sparse = any_sparse || job->sparse[c];
    if (job->sparse[c] || (job->rows_hint[c] <= 4 && g.plane_bytes[c] >= kMinPitchedPlane)) all_dense_plain = false;
  }
  if (job->scan_active) {
    any_sparse = false;  // every plane comes from the device decoder
  } else if (!all_dense_plain || any_sparse) {
    for (int c = 0; c < g.ncomp; c++) {
      uint8_t *dp = job->d_coef + g.plane_off[c];
      if (job->sparse[c]) {
        const size_t nb = (size_t)g.bx[c] * g.by[c];
        uint8_t *dsp = job->d_sparse + job->sparse_off[c];
        const uint8_t *hsp = job->h_sparse + job->sparse_off[c];
        if (job->sparse_runs[c].empty()) {
          CU_TRY(ctx, cudaMemcpyAsync(dsp, hsp, (nb + 1 + job->sparse_count[c]) * 4, cudaMemcpyHostToDevice, s));
        } else {
          // first[] as it is, then every run to its place in the concatenation: the copy engine
          // does the compaction that several producer threads could not do in place
          CU_TRY(ctx, cudaMemcpyAsync(dsp, hsp, (nb + 1) * 4, cudaMemcpyHostToDevice, s));
          size_t at = nb + 1;
          for (const auto &run : job->sparse_runs[c]) {
            if (run.second)
              CU_TRY(ctx, cudaMemcpyAsync(dsp + at * 4, hsp + (nb + 1 + run.first) * 4, run.second * 4,
                                          cudaMemcpyHostToDevice, s));
            at += run.second;
          }
        }
        job->d_tail_zero[c] = false;
        continue;
      }
      if (!job->h_coef || !job->h_coef_clean)
        return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_submit: component %d has neither a plane nor records", c);
      const uint8_t *hp = job->h_coef + g.plane_off[c];
      if (job->rows_hint[c] <= 4 && g.plane_bytes[c] >= kMinPitchedPlane) {
        if (!job->d_tail_zero[c]) {
          CU_TRY(ctx, cudaMemsetAsync(dp, 0, g.plane_bytes[c], s));
          job->d_tail_zero[c] = true;
        }
        CU_TRY(ctx, cudaMemcpy2DAsync(dp, 128, hp, 128, 64, g.plane_bytes[c] / 128, cudaMemcpyHostToDevice, s));
      } else {
        CU_TRY(ctx, cudaMemcpyAsync(dp, hp, g.plane_bytes[c], cudaMemcpyHostToDevice, s));
        job->d_tail_zero[c] = false;
      }
    }
  } else {
    if (!job->h_coef || !job->h_coef_clean)
      return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_submit: the frame has neither planes nor records");
    CU_TRY(ctx, cudaMemcpyAsync(job->d_coef, job->h_coef, g.coef_bytes, cudaMemcpyHostToDevice, s));
    for (int c = 0; c < 3; c++) job->d_tail_zero[c] = false;
  }
  if (any_sparse) {
    gidk::ExpandArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.mcux = (uint32_t)g.mcux;
    for (int c = 0; c < g.ncomp; c++) {
      const size_t nb = (size_t)g.bx[c] * g.by[c];
      ea.h[c] = job->desc.samp_h[c];
      ea.v[c] = job->desc.samp_v[c];
      ea.bx[c] = (uint32_t)g.bx[c];
      if (!job->sparse[c]) continue;  // nblocks stays 0: the component's plane came as a dense copy
      ea.nblocks[c] = (uint32_t)nb;
      ea.first[c] = reinterpret_cast<const uint32_t *>(job->d_sparse + job->sparse_off[c]);
      ea.records[c] = ea.first[c] + nb + 1;
      ea.plane[c] = reinterpret_cast<int16_t *>(job->d_coef + g.plane_off[c]);
    }
    CU_TRY(ctx, gidk::launch_expand(ea, g.ncomp, s));
  }
  CU_TRY(ctx, job->set.launch(ctx->num_sms, s));
  if (!job->user_pixels && !job->h_pixels)
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_pixels, job->cap_pixels, cudaHostAllocDefault));
  CU_TRY(ctx, cudaMemcpyAsync(job->user_pixels ? job->user_pixels : job->h_pixels, job->d_pixels, g.pixel_bytes,
                              cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, cudaEventRecord(job->done, s));
  job->state = gidcuda_job::SUBMITTED;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_wait(gidcuda_job *job, const uint8_t **pixels, size_t *stride) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::SUBMITTED)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_wait: job was not submitted");
  CU_TRY(ctx, cudaEventSynchronize(job->done));
  if (job->scan_active) {
    job->scan_active = false;  // a re-submission after a host decode takes planes / records
    uint32_t blocks = 0;
    for (int c = 0; c < job->geo.ncomp; c++) blocks += (uint32_t)job->geo.bx[c] * (uint32_t)job->geo.by[c];
    if (job->scan_ri == 0 && (job->h_status[0] & 2u) && g_chain_repair) {
      // an unsettled chain: the host walks the stretches that did not re-synchronise.  (Threads that
      // decoded from a consistent but untrue state may have met undecodable codes, flag 8, as well:
      // the output pass over the repaired chain tells whether the stream itself is damaged.)
      const int rc = job_repair_chain(job);
      if (rc != GIDCUDA_OK) return rc;
      g_chain_repairs.fetch_add(1, std::memory_order_relaxed);
      g_chain_repair_subsequences.fetch_add((long)job->repaired_subsequences, std::memory_order_relaxed);
    }
    if (job->scan_ri == 0 && job->h_status[0] == 0 && job->h_status[2] != blocks) job->h_status[0] = 4;
    if (job->scan_ri == 0 && job->h_status[0] != 0) {
      job->state = gidcuda_job::BEGUN;
      return fail(ctx, GIDCUDA_ERR_DATA,
                  "device entropy decoder: scan left to the host decoder (flags %u, %u of %u blocks)", job->h_status[0],
                  job->h_status[2], blocks);
    }
    if (job->h_status[0] != 0) {
      job->state = gidcuda_job::BEGUN;
      return fail(ctx, GIDCUDA_ERR_DATA,
                  "device entropy decoder: %u restart interval(s) left to the host decoder (damaged or unusual stream)",
                  job->h_status[1]);
    }
  }
  if (pixels) *pixels = job->user_pixels ? job->user_pixels : job->h_pixels;
  if (stride) *stride = job->geo.stride;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_release(gidcuda_job *job) {
  if (!job) return GIDCUDA_OK;
  gidcuda_ctx *ctx = job->ctx;
  if (job->state == gidcuda_job::IDLE) return fail(ctx, GIDCUDA_ERR_STATE, "job released twice");
  if (job->state == gidcuda_job::SUBMITTED) (void)cudaEventSynchronize(job->done);
  job->state = gidcuda_job::IDLE;
  ctx->live_jobs--;
  ctx->free_jobs.push_back(job);
  return GIDCUDA_OK;
}

extern "C" int gidcuda_synchronize(gidcuda_ctx *ctx) {
  if (!ctx) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null ctx");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GIDCUDA_OK;
}

extern "C" int gidcuda_destroy(gidcuda_ctx *ctx) {
  if (!ctx) return GIDCUDA_OK;
  if (ctx->live_jobs != 0)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_destroy: %d job(s) not released", ctx->live_jobs);
  (void)cudaSetDevice(ctx->device);
  (void)cudaStreamSynchronize(ctx->stream);
  for (gidcuda_job *j : ctx->free_jobs) {
    job_free_buffers(j);
    delete j;
  }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return GIDCUDA_OK;
}

// ---------------------------------------------------------------------------
// plans: prepared launches over device-resident frames

struct gidcuda_plan {
  gidcuda_ctx *ctx = nullptr;
  std::vector<gidcuda_frame_desc> descs;
  gidk::LaunchSet set;
  std::vector<uint8_t *> d_samples;  // scratch of generic frames
};

extern "C" int gidcuda_plan_destroy(gidcuda_plan *plan) {
  if (!plan) return GIDCUDA_OK;
  (void)cudaSetDevice(plan->ctx->device);
  for (uint8_t *p : plan->d_samples) cudaFree(p);
  delete plan;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_plan_create(gidcuda_ctx *ctx, int32_t n, const gidcuda_frame_desc *descs,
                                   const int16_t *const *d_planes, uint8_t *const *d_pixels,
                                   gidcuda_plan **out) {
  if (!ctx || !out || n < 0 || (n > 0 && (!descs || !d_planes || !d_pixels)))
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_plan_create: bad argument");
  *out = nullptr;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  gidcuda_plan *plan = new (std::nothrow) gidcuda_plan();
  if (!plan) return fail(ctx, GIDCUDA_ERR_OUT_OF_MEMORY, "out of host memory");
  plan->ctx = ctx;
  plan->descs.assign(descs, descs + n);
  std::vector<gidk::FrameItem> items;
  int rc = GIDCUDA_OK;
  for (int32_t i = 0; i < n && rc == GIDCUDA_OK; i++) {
    gidk::FrameItem it;
    it.desc = &plan->descs[(size_t)i];
    rc = geometry_or_fail(ctx, it.desc, &it.geo);
    if (rc != GIDCUDA_OK) break;
    for (int c = 0; c < 3; c++) {
      it.planes[c] = c < it.geo.ncomp ? d_planes[3 * i + c] : nullptr;
      if (c < it.geo.ncomp && (!it.planes[c] || (reinterpret_cast<uintptr_t>(it.planes[c]) & 15u))) {
        rc = fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "frame %d plane %d: null or not 16-byte aligned", i, c);
        break;
      }
    }
    if (rc != GIDCUDA_OK) break;
    if (!d_pixels[i]) { rc = fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "frame %d: null pixel buffer", i); break; }
    it.pixels = d_pixels[i];
    it.samples = nullptr;
    if (it.geo.layout == gidk::LAYOUT_GENERIC) {
      cudaError_t e = cudaMalloc((void **)&it.samples, it.geo.sample_bytes);
      if (e != cudaSuccess) { rc = fail(ctx, GIDCUDA_ERR_OUT_OF_MEMORY, "cudaMalloc: %s", cudaGetErrorString(e)); break; }
      plan->d_samples.push_back(it.samples);
    }
    items.push_back(it);
  }
  if (rc == GIDCUDA_OK) {
    plan->set.build(items);
    cudaError_t e = plan->set.upload(ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, GIDCUDA_ERR_CUDA, "plan upload: %s", cudaGetErrorString(e));
  }
  if (rc != GIDCUDA_OK) {
    gidcuda_plan_destroy(plan);
    return rc;
  }
  *out = plan;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_plan_launch(gidcuda_plan *plan, void *cuda_stream) {
  if (!plan) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null plan");
  gidcuda_ctx *ctx = plan->ctx;
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  CU_TRY(ctx, plan->set.launch(ctx->num_sms, s));
  return GIDCUDA_OK;
}

extern "C" int gidcuda_plan_launch_count(const gidcuda_plan *plan) { return plan ? plan->set.launch_count() : 0; }

extern "C" int gidcuda_set_option(const char *name, int value) {
  if (!name) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null option name");
  if (strcmp(name, "chain_repair") == 0) {
    g_chain_repair = value ? 1 : 0;
    return GIDCUDA_OK;
  }
  if (strcmp(name, "blocking_wait") == 0) {
    g_blocking_wait = value ? 1 : 0;
    return GIDCUDA_OK;
  }
  if (strcmp(name, "tma") == 0) {
    gidk::tma_option() = value ? 1 : 0;
    return GIDCUDA_OK;
  }
  return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "unknown option '%s'", name);
}

extern "C" long gidcuda_counter(const char *name) {
  if (name && strcmp(name, "chain_repairs") == 0) return g_chain_repairs.load();
  if (name && strcmp(name, "chain_repair_subsequences") == 0) return g_chain_repair_subsequences.load();
  return -1;
}

extern "C" int gidcuda_plan_uses_tma(const gidcuda_plan *plan) { return plan ? plan->set.tma_groups : 0; }

// ---------------------------------------------------------------------------
// batch driver: per-GPU worker thread + work queue + ring of in-flight chunks

namespace {

struct BatchRun {
  int32_t n = 0;
  const gidcuda_frame_desc *descs = nullptr;
  const int16_t *const *planes = nullptr;
  uint8_t *const *pixels = nullptr;
  int32_t *status = nullptr;
  std::atomic<int32_t> next{0};  // shared work queue: next image index (dynamic sharding)
  std::atomic<int64_t> h2d{0}, d2h{0};
  std::atomic<int32_t> launches{0}, failed{0};
  int32_t per_device[16] = {0};
};

// A chunk in flight on one device: staging buffers + launch tables.
struct ChunkSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  uint8_t *h_in = nullptr, *d_in = nullptr, *h_out = nullptr, *d_out = nullptr;
  size_t cap_in = 0, cap_out = 0;
  uint8_t *d_samples = nullptr;
  size_t cap_samples = 0;
  gidk::LaunchSet set;
  // contents
  std::vector<int32_t> images;
  std::vector<Geometry> geos;
  std::vector<size_t> out_off;
  bool busy = false;
};

struct Worker {
  int device = 0;
  int index = 0;
  int num_sms = 148;
  std::thread thread;
  std::vector<ChunkSlot> slots;
};

}  // namespace

struct gidcuda_batch {
  std::vector<Worker> workers;
  int chunk_images = 8;
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  BatchRun *run = nullptr;
  uint64_t generation = 0;
  int pending = 0;
  bool quit = false;
  std::string error;
};

namespace {

bool cu_ok(gidcuda_batch *b, cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  std::lock_guard<std::mutex> lk(b->mu);
  if (b->error.empty()) b->error = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

bool grow_pair(gidcuda_batch *b, uint8_t **h, uint8_t **d, size_t *cap, size_t need, bool host_too) {
  if (*cap >= need) return true;
  if (host_too && *h) cudaFreeHost(*h);
  if (*d) cudaFree(*d);
  if (host_too) *h = nullptr;
  *d = nullptr;
  *cap = 0;
  need += need / 4;
  if (host_too && !cu_ok(b, cudaHostAlloc((void **)h, need, cudaHostAllocDefault), "cudaHostAlloc")) return false;
  if (!cu_ok(b, cudaMalloc((void **)d, need), "cudaMalloc")) return false;
  *cap = need;
  return true;
}

// Finish a slot: wait for its stream, copy pixels to the caller's buffers.
void slot_finish(gidcuda_batch *b, BatchRun *run, ChunkSlot &s) {
  if (!s.busy) return;
  const bool ok = cu_ok(b, cudaEventSynchronize(s.done), "cudaEventSynchronize");
  for (size_t k = 0; k < s.images.size(); k++) {
    const int32_t i = s.images[k];
    if (ok) {
      memcpy(run->pixels[i], s.h_out + s.out_off[k], s.geos[k].pixel_bytes);
      if (run->status) run->status[i] = GIDCUDA_OK;
    } else {
      if (run->status) run->status[i] = GIDCUDA_ERR_CUDA;
      run->failed++;
    }
  }
  s.busy = false;
}

// Fill a slot with up to chunk_images images from the shared queue and submit.
// Returns false when the queue is empty.
bool slot_submit(gidcuda_batch *b, BatchRun *run, Worker &w, ChunkSlot &s) {
  s.images.clear();
  s.geos.clear();
  s.out_off.clear();
  size_t in_bytes = 0, out_bytes = 0, samples = 0;
  std::vector<size_t> in_off, sample_off;
  bool drained = false;
  while ((int)s.images.size() < b->chunk_images) {
    const int32_t i = run->next.fetch_add(1);
    if (i >= run->n) { drained = true; break; }
    Geometry g;
    const int rc = geometry_or_fail(nullptr, &run->descs[i], &g);
    bool good = rc == GIDCUDA_OK && run->pixels[i] != nullptr;
    for (int c = 0; good && c < g.ncomp; c++) good = run->planes[3 * i + c] != nullptr;
    if (!good) {
      // a failed image is reported and skipped; the others are unaffected
      if (run->status) run->status[i] = rc != GIDCUDA_OK ? rc : GIDCUDA_ERR_BAD_ARGUMENT;
      run->failed++;
      continue;
    }
    s.images.push_back(i);
    s.geos.push_back(g);
    in_off.push_back(in_bytes);
    s.out_off.push_back(out_bytes);
    sample_off.push_back(samples);
    in_bytes += g.coef_bytes;
    out_bytes += round_up(g.pixel_bytes, 256);
    if (g.layout == gidk::LAYOUT_GENERIC) samples += round_up(g.sample_bytes, 256);
  }
  if (s.images.empty()) return !drained;
  const size_t nf = s.images.size();
  uint8_t *none = nullptr;
  if (!grow_pair(b, &s.h_in, &s.d_in, &s.cap_in, in_bytes, true)) return true;
  if (!grow_pair(b, &s.h_out, &s.d_out, &s.cap_out, out_bytes, true)) return true;
  if (samples && !grow_pair(b, &none, &s.d_samples, &s.cap_samples, samples, false)) return true;
  // stage coefficients into pinned memory, then one H2D copy per chunk
  std::vector<gidk::FrameItem> items(nf);
  for (size_t k = 0; k < nf; k++) {
    const int32_t i = s.images[k];
    const Geometry &g = s.geos[k];
    gidk::FrameItem &it = items[k];
    it.desc = &run->descs[i];
    it.geo = g;
    for (int c = 0; c < 3; c++) {
      it.planes[c] = nullptr;
      if (c >= g.ncomp) continue;
      memcpy(s.h_in + in_off[k] + g.plane_off[c], run->planes[3 * i + c], g.plane_bytes[c]);
      it.planes[c] = reinterpret_cast<const int16_t *>(s.d_in + in_off[k] + g.plane_off[c]);
    }
    it.pixels = s.d_out + s.out_off[k];
    it.samples = g.layout == gidk::LAYOUT_GENERIC ? s.d_samples + sample_off[k] : nullptr;
  }
  s.set.build(items);
  cudaStream_t st = s.stream;
  bool ok = cu_ok(b, cudaMemcpyAsync(s.d_in, s.h_in, in_bytes, cudaMemcpyHostToDevice, st), "H2D coefficients");
  ok = ok && cu_ok(b, s.set.upload(st), "H2D launch tables");
  ok = ok && cu_ok(b, s.set.launch(w.num_sms, st), "kernel launch");
  ok = ok && cu_ok(b, cudaMemcpyAsync(s.h_out, s.d_out, out_bytes, cudaMemcpyDeviceToHost, st), "D2H pixels");
  ok = ok && cu_ok(b, cudaEventRecord(s.done, st), "cudaEventRecord");
  if (!ok) {
    for (int32_t i : s.images) {
      if (run->status) run->status[i] = GIDCUDA_ERR_CUDA;
      run->failed++;
    }
    return !drained;  // keep draining the queue so that every image gets a status
  }
  run->h2d += (int64_t)(in_bytes + s.set.upload_bytes);
  run->d2h += (int64_t)out_bytes;
  run->launches += s.set.launch_count();
  run->per_device[w.index] += (int32_t)nf;
  s.busy = true;
  return !drained;
}

void worker_main(gidcuda_batch *b, Worker *w) {
  uint64_t seen = 0;
  cu_ok(b, cudaSetDevice(w->device), "cudaSetDevice");
  cu_ok(b, cudaDeviceGetAttribute(&w->num_sms, cudaDevAttrMultiProcessorCount, w->device), "device attribute");
  cu_ok(b, gidk::init_kernels(), "kernel set-up");
  for (ChunkSlot &s : w->slots) {
    cu_ok(b, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    cu_ok(b, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate");
  }
  for (;;) {
    BatchRun *run;
    {
      std::unique_lock<std::mutex> lk(b->mu);
      b->cv_start.wait(lk, [&] { return b->quit || b->generation != seen; });
      if (b->quit) break;
      seen = b->generation;
      run = b->run;
    }
    // ring of slots: finish the oldest slot, refill it, move on
    size_t cur = 0;
    bool more = true;
    while (more) {
      ChunkSlot &s = w->slots[cur];
      slot_finish(b, run, s);
      more = slot_submit(b, run, *w, s);
      cur = (cur + 1) % w->slots.size();
    }
    for (size_t k = 0; k < w->slots.size(); k++) slot_finish(b, run, w->slots[(cur + k) % w->slots.size()]);
    {
      std::lock_guard<std::mutex> lk(b->mu);
      if (--b->pending == 0) b->cv_done.notify_all();
    }
  }
  for (ChunkSlot &s : w->slots) {
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_samples) cudaFree(s.d_samples);
    s.set.destroy();
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
}

}  // namespace

extern "C" int gidcuda_batch_create(const int32_t *devices, int32_t ndevices, int32_t chunk_images,
                                    gidcuda_batch **out) {
  if (!out || ndevices < 1 || ndevices > 16 || !devices)
    return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_batch_create: bad argument");
  *out = nullptr;
  const int n = gidcuda_device_count();
  if (n == 0) return fail(nullptr, GIDCUDA_ERR_NO_DEVICE, "no CUDA device: no CPU fallback");
  for (int i = 0; i < ndevices; i++)
    if (devices[i] < 0 || devices[i] >= n)
      return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "device %d out of range", devices[i]);
  gidcuda_batch *b = new (std::nothrow) gidcuda_batch();
  if (!b) return fail(nullptr, GIDCUDA_ERR_OUT_OF_MEMORY, "out of host memory");
  b->chunk_images = chunk_images > 0 ? chunk_images : 8;
  b->workers.resize((size_t)ndevices);
  for (int i = 0; i < ndevices; i++) {
    b->workers[(size_t)i].device = devices[i];
    b->workers[(size_t)i].index = i;
    b->workers[(size_t)i].slots.resize(3);
  }
  for (Worker &w : b->workers) w.thread = std::thread(worker_main, b, &w);
  *out = b;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_batch_run(gidcuda_batch *b, int32_t n, const gidcuda_frame_desc *descs,
                                 const int16_t *const *planes, uint8_t *const *pixels, int32_t *status,
                                 gidcuda_batch_stats *stats) {
  if (!b || n < 0 || (n > 0 && (!descs || !planes || !pixels)))
    return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_batch_run: bad argument");
  BatchRun run;
  run.n = n;
  run.descs = descs;
  run.planes = planes;
  run.pixels = pixels;
  run.status = status;
  const auto t0 = std::chrono::steady_clock::now();
  {
    std::unique_lock<std::mutex> lk(b->mu);
    b->error.clear();
    b->run = &run;
    b->pending = (int)b->workers.size();
    b->generation++;
    b->cv_start.notify_all();
    b->cv_done.wait(lk, [&] { return b->pending == 0; });
    b->run = nullptr;
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->seconds = std::chrono::duration<double>(t1 - t0).count();
    for (int32_t i = 0; i < n; i++) stats->megapixels += (double)descs[i].width * descs[i].height / 1e6;
    stats->h2d_bytes = run.h2d;
    stats->d2h_bytes = run.d2h;
    stats->kernel_launches = run.launches;
    stats->images_failed = run.failed;
    for (int i = 0; i < 16; i++) stats->images_per_device[i] = run.per_device[i];
  }
  if (!b->error.empty()) return fail(nullptr, GIDCUDA_ERR_CUDA, "batch: %s", b->error.c_str());
  return GIDCUDA_OK;
}

extern "C" int gidcuda_batch_destroy(gidcuda_batch *b) {
  if (!b) return GIDCUDA_OK;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->quit = true;
    b->cv_start.notify_all();
  }
  for (Worker &w : b->workers)
    if (w.thread.joinable()) w.thread.join();
  delete b;
  return GIDCUDA_OK;
}
