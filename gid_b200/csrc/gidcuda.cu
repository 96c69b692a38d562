// gidcuda.cu -- C-ABI shim of the B200 reconstruction path (include/gidcuda.h): library and device,
// jobs (one frame through pinned buffers), plans (frames already in device memory), options.
// gidcuda_scan.cu holds the part for Huffman decoding on the device, gidcuda_batch.cu the multi-GPU
// batch driver.
//
// Owns pinned host buffers, device buffers, streams and launches; the Ada
// binding GID.CUDA and the C++ host mirror only see plain pointers and sizes.
// No CPU fallback: every compute entry point needs a CUDA device.
#include "gidcuda_internal.hpp"

// ---------------------------------------------------------------------------
// errors

static thread_local std::string t_last_error;

int fail(gidcuda_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_last_error = buf;
  if (ctx) ctx->last_error = buf;
  return code;
}

// ---------------------------------------------------------------------------
// geometry (frame_setup.hpp)

int geometry_or_fail(gidcuda_ctx *ctx, const gidcuda_frame_desc *d, Geometry *g) {
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
int g_blocking_wait = 0;
// gidcuda_set_option("chain_repair"): 1 (default) = an unsettled chain of the self-synchronising Huffman
// decoder is repaired by the host (job_repair_chain); 0 = reported as GIDCUDA_ERR_DATA at once.
int g_chain_repair = 1;
// gidcuda_set_option("plan_pdl"): 1 (default) = plan launches are programmatic dependent launches
int g_plan_pdl = 1;
std::atomic<long> g_chain_repairs{0}, g_chain_repair_subsequences{0};


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
  job_free_pscan(j);
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

// launch tables for this frame (device addresses are fixed once job_reserve has run)
static void job_build_set(gidcuda_job *j) {
  gidk::FrameItem it;
  it.desc = &j->desc;
  it.geo = j->geo;
  for (int c = 0; c < 3; c++)
    it.planes[c] = c < j->geo.ncomp ? reinterpret_cast<const int16_t *>(j->d_coef + j->geo.plane_off[c]) : nullptr;
  it.pixels = j->d_pixels;
  it.samples = j->d_samples;
  j->set.build(std::vector<gidk::FrameItem>(1, it));
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
  job_build_set(j);
  j->h_coef_clean = false;  // zeroed when the decoder asks for the planes
  j->user_pixels = nullptr;
  j->scan_active = false;
  j->pscan_active = false;
  for (int c = 0; c < 3; c++) {
    j->rows_hint[c] = 8;
    // a recycled job's device planes hold another frame's geometry: whatever was zeroed there, the
    // rows 4 .. 7 of THIS frame's blocks are not known to be
    j->d_tail_zero[c] = false;
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

extern "C" int32_t gidcuda_fast_coef_limit(const gidcuda_frame_desc *desc, int comp) {
  if (!desc || comp < 0 || comp >= 3) return 0;
  return gidk::fast_coef_limit(desc, comp);
}

extern "C" int gidcuda_job_wide_coefficients(gidcuda_job *job) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  if (job->state != gidcuda_job::BEGUN)
    return fail(job->ctx, GIDCUDA_ERR_STATE, "gidcuda_job_wide_coefficients: only between gidcuda_job_begin and gidcuda_job_submit");
  if (job->desc.flags & GIDCUDA_FLAG_WIDE_COEF) return GIDCUDA_OK;
  job->desc.flags |= GIDCUDA_FLAG_WIDE_COEF;
  Geometry g;
  const int rc = geometry_or_fail(job->ctx, &job->desc, &g);  // same geometry, exact kernel instances
  if (rc != GIDCUDA_OK) return rc;
  job->geo = g;
  job_build_set(job);
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
  if (job->scan_active) {  // compressed bytes -> device, Huffman decoding there (gidcuda_scan.cu)
    const int rc = job_submit_scan(job, s);
    if (rc != GIDCUDA_OK) return rc;
  }
  if (job->pscan_active) {  // a progressive frame whose scans were queued one by one: the planes are on the device
    const int rc = job_submit_pscan(job, s);
    if (rc != GIDCUDA_OK) return rc;
  }
  constexpr size_t kMinPitchedPlane = (size_t)8 << 20;
  bool any_sparse = false, all_dense_plain = true;
  for (int c = 0; c < g.ncomp; c++) {
    any_sparse = any_sparse || job->sparse[c];
    if (job->sparse[c] || (job->rows_hint[c] <= 4 && g.plane_bytes[c] >= kMinPitchedPlane)) all_dense_plain = false;
  }
  if (job->scan_active || job->pscan_active) {
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
      ea.count[c] = (uint32_t)job->sparse_count[c];
      ea.first[c] = reinterpret_cast<const uint32_t *>(job->d_sparse + job->sparse_off[c]);
      ea.records[c] = ea.first[c] + nb + 1;
      ea.plane[c] = reinterpret_cast<int16_t *>(job->d_coef + g.plane_off[c]);
    }
    CU_TRY(ctx, gidk::launch_expand(ea, g.ncomp, s));
  }
  CU_TRY(ctx, job->set.launch(ctx->num_sms, s));
  if (job->scan_active) {
    const int rc = job_submit_scan_status(job, s);
    if (rc != GIDCUDA_OK) return rc;
  }
  if (job->pscan_active) {
    const int rc = job_submit_pscan_status(job, s);
    if (rc != GIDCUDA_OK) return rc;
  }
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
  if (job->scan_active) {  // what the device decoder made of the scan (the host's chain repair included)
    const int rc = job_finish_scan(job);
    if (rc != GIDCUDA_OK) return rc;
  }
  if (job->pscan_active) {
    const int rc = job_finish_pscan(job);
    if (rc != GIDCUDA_OK) return rc;
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
  CU_TRY(ctx, plan->set.launch(ctx->num_sms, s, g_plan_pdl != 0));
  return GIDCUDA_OK;
}

extern "C" int gidcuda_plan_launch_count(const gidcuda_plan *plan) { return plan ? plan->set.launch_count() : 0; }

extern "C" int gidcuda_set_option(const char *name, int value) {
  if (!name) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null option name");
  if (strcmp(name, "chain_repair") == 0) {
    g_chain_repair = value ? 1 : 0;
    return GIDCUDA_OK;
  }
  if (strcmp(name, "turn_tile_rows") == 0) {  // tuning: MCU rows per tile of frames turned by a quarter (0 = choose)
    gidk::turn_tile_rows_option() = value;
    return GIDCUDA_OK;
  }
  if (strcmp(name, "blocking_wait") == 0) {
    g_blocking_wait = value ? 1 : 0;
    return GIDCUDA_OK;
  }
  if (strcmp(name, "plan_pdl") == 0) {
    g_plan_pdl = value ? 1 : 0;
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

