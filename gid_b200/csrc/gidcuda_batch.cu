// gidcuda_batch.cu -- gidcuda_batch_*: independent images sharded over the GPUs of one box
// (include/gidcuda.h): one worker thread, one work queue and a ring of in-flight chunks per device; no
// inter-GPU communication.
#include "gidcuda_internal.hpp"

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
