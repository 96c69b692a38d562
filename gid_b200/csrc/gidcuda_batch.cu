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
  const int16_t *const *planes = nullptr;        // dense input (gidcuda_batch_run) ...
  const uint32_t *const *first = nullptr;         // ... or records (gidcuda_batch_run_records): [n * 3]
  const uint32_t *const *records = nullptr;
  const size_t *counts = nullptr;
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
  size_t cap_hin = 0, cap_hout = 0;  // pinned staging, allocated only for images whose buffers are not pinned
  uint8_t *d_samples = nullptr;
  size_t cap_samples = 0;
  uint8_t *d_sparse = nullptr;       // records input: per image and component [first: nblocks + 1][records]
  size_t cap_sparse = 0;
  gidk::LaunchSet set;
  // contents
  std::vector<int32_t> images;
  std::vector<Geometry> geos;
  std::vector<size_t> out_off;
  std::vector<uint8_t> staged_out;   // per image: 1 = pixels travel through h_out and are copied at the end
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

bool grow_dev(gidcuda_batch *b, uint8_t **d, size_t *cap, size_t need) {
  if (*cap >= need) return true;
  if (*d) cudaFree(*d);
  *d = nullptr;
  *cap = 0;
  need += need / 4;
  if (!cu_ok(b, cudaMalloc((void **)d, need), "cudaMalloc")) return false;
  *cap = need;
  return true;
}
bool grow_host(gidcuda_batch *b, uint8_t **h, size_t *cap, size_t need) {
  if (*cap >= need) return true;
  if (*h) cudaFreeHost(*h);
  *h = nullptr;
  *cap = 0;
  need += need / 4;
  if (!cu_ok(b, cudaHostAlloc((void **)h, need, cudaHostAllocDefault), "cudaHostAlloc")) return false;
  *cap = need;
  return true;
}

bool pinned(const void *p) { return gidcuda_host_is_pinned(p) != 0; }

// Every image a slot has claimed gets a status, whatever happens to the slot.
void slot_fail(BatchRun *run, ChunkSlot &s, int code) {
  for (int32_t i : s.images) {
    if (run->status) run->status[i] = code;
    run->failed++;
  }
  s.images.clear();
  s.busy = false;
}

// Finish a slot: wait for its stream; pixels of images whose buffers are not pinned are copied out of
// the staging buffer (pinned buffers were written by the copy engine itself).
void slot_finish(gidcuda_batch *b, BatchRun *run, ChunkSlot &s) {
  if (!s.busy) return;
  if (!cu_ok(b, cudaEventSynchronize(s.done), "cudaEventSynchronize")) {
    slot_fail(run, s, GIDCUDA_ERR_CUDA);
    return;
  }
  for (size_t k = 0; k < s.images.size(); k++) {
    const int32_t i = s.images[k];
    if (s.staged_out[k]) memcpy(run->pixels[i], s.h_out + s.out_off[k], s.geos[k].pixel_bytes);
    if (run->status) run->status[i] = GIDCUDA_OK;
  }
  s.busy = false;
}

// Fill a slot with up to chunk_images images from the shared queue and submit.
// Returns false when the queue is empty.
bool slot_submit(gidcuda_batch *b, BatchRun *run, Worker &w, ChunkSlot &s) {
  s.images.clear();
  s.geos.clear();
  s.out_off.clear();
  s.staged_out.clear();
  const bool sparse = run->first != nullptr;
  size_t in_bytes = 0, out_bytes = 0, samples = 0, sparse_bytes = 0, stage_in = 0, stage_out = 0;
  std::vector<size_t> in_off, sample_off, sparse_off, hin_off;
  std::vector<uint8_t> staged_in;
  bool drained = false;
  while ((int)s.images.size() < b->chunk_images) {
    const int32_t i = run->next.fetch_add(1);
    if (i >= run->n) { drained = true; break; }
    Geometry g;
    const int rc = geometry_or_fail(nullptr, &run->descs[i], &g);
    bool good = rc == GIDCUDA_OK && run->pixels[i] != nullptr;
    bool in_pinned = true;
    for (int c = 0; good && c < g.ncomp; c++) {
      if (sparse) {
        const size_t nb = (size_t)g.bx[c] * g.by[c];
        const uint32_t *f = run->first[3 * i + c];
        good = f != nullptr && run->records[3 * i + c] != nullptr && f[0] == 0 && f[nb] == run->counts[3 * i + c];
        in_pinned = in_pinned && good && pinned(f) && pinned(run->records[3 * i + c]);
      } else {
        good = run->planes[3 * i + c] != nullptr;
        in_pinned = in_pinned && good && pinned(run->planes[3 * i + c]);
      }
    }
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
    sparse_off.push_back(sparse_bytes);
    hin_off.push_back(stage_in);
    staged_in.push_back(in_pinned ? 0 : 1);
    s.staged_out.push_back(pinned(run->pixels[i]) ? 0 : 1);
    in_bytes += g.coef_bytes;
    out_bytes += round_up(g.pixel_bytes, 256);
    if (s.staged_out.back()) stage_out = out_bytes;  // (h_out mirrors d_out's layout: simplest)
    if (g.layout == gidk::LAYOUT_GENERIC) samples += round_up(g.sample_bytes, 256);
    if (sparse) {
      size_t bytes = 0;
      for (int c = 0; c < g.ncomp; c++)
        bytes += round_up(((size_t)g.bx[c] * g.by[c] + 1 + run->counts[3 * i + c]) * 4, 256);
      sparse_bytes += bytes;
      if (!in_pinned) stage_in += bytes;
    } else if (!in_pinned) {
      stage_in += g.coef_bytes;
    }
  }
  if (s.images.empty()) return !drained;
  const size_t nf = s.images.size();
  if (!grow_dev(b, &s.d_in, &s.cap_in, in_bytes) || !grow_dev(b, &s.d_out, &s.cap_out, out_bytes) ||
      (samples && !grow_dev(b, &s.d_samples, &s.cap_samples, samples)) ||
      (sparse_bytes && !grow_dev(b, &s.d_sparse, &s.cap_sparse, sparse_bytes)) ||
      (stage_in && !grow_host(b, &s.h_in, &s.cap_hin, stage_in)) ||
      (stage_out && !grow_host(b, &s.h_out, &s.cap_hout, stage_out))) {
    slot_fail(run, s, GIDCUDA_ERR_OUT_OF_MEMORY);
    return !drained;
  }
  cudaStream_t st = s.stream;
  bool ok = true;
  int64_t moved_in = 0;
  std::vector<gidk::FrameItem> items(nf);
  std::vector<gidk::ExpandArgs> expands;
  for (size_t k = 0; k < nf && ok; k++) {
    const int32_t i = s.images[k];
    const Geometry &g = s.geos[k];
    gidk::FrameItem &it = items[k];
    it.desc = &run->descs[i];
    it.geo = g;
    gidk::ExpandArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.mcux = (uint32_t)g.mcux;
    size_t sp = sparse_off[k], hin = hin_off[k];
    for (int c = 0; c < 3; c++) {
      it.planes[c] = nullptr;
      if (c >= g.ncomp) continue;
      uint8_t *dp = s.d_in + in_off[k] + g.plane_off[c];
      it.planes[c] = reinterpret_cast<const int16_t *>(dp);
      if (sparse) {
        // records: first[] and the records used cross the link; the expand kernel rebuilds the plane
        const size_t nb = (size_t)g.bx[c] * g.by[c], cnt = run->counts[3 * i + c];
        const uint32_t *f = run->first[3 * i + c], *r = run->records[3 * i + c];
        uint8_t *ds = s.d_sparse + sp;
        if (staged_in[k]) {
          memcpy(s.h_in + hin, f, (nb + 1) * 4);
          memcpy(s.h_in + hin + (nb + 1) * 4, r, cnt * 4);
          ok = ok && cu_ok(b, cudaMemcpyAsync(ds, s.h_in + hin, (nb + 1 + cnt) * 4, cudaMemcpyHostToDevice, st), "H2D records");
          hin += round_up((nb + 1 + cnt) * 4, 256);
        } else {
          ok = ok && cu_ok(b, cudaMemcpyAsync(ds, f, (nb + 1) * 4, cudaMemcpyHostToDevice, st), "H2D first[]");
          if (cnt) ok = ok && cu_ok(b, cudaMemcpyAsync(ds + (nb + 1) * 4, r, cnt * 4, cudaMemcpyHostToDevice, st), "H2D records");
        }
        moved_in += (int64_t)((nb + 1 + cnt) * 4);
        ea.h[c] = run->descs[i].samp_h[c];
        ea.v[c] = run->descs[i].samp_v[c];
        ea.bx[c] = (uint32_t)g.bx[c];
        ea.nblocks[c] = (uint32_t)nb;
        ea.count[c] = (uint32_t)cnt;
        ea.first[c] = reinterpret_cast<const uint32_t *>(ds);
        ea.records[c] = ea.first[c] + nb + 1;
        ea.plane[c] = reinterpret_cast<int16_t *>(dp);
        sp += round_up((nb + 1 + cnt) * 4, 256);
      } else {
        const void *src = run->planes[3 * i + c];
        if (staged_in[k]) {  // pageable memory: through the slot's pinned staging buffer
          memcpy(s.h_in + hin + g.plane_off[c], src, g.plane_bytes[c]);
          src = s.h_in + hin + g.plane_off[c];
        }
        ok = ok && cu_ok(b, cudaMemcpyAsync(dp, src, g.plane_bytes[c], cudaMemcpyHostToDevice, st), "H2D coefficients");
        moved_in += (int64_t)g.plane_bytes[c];
      }
    }
    if (sparse) expands.push_back(ea);
    it.pixels = s.d_out + s.out_off[k];
    it.samples = g.layout == gidk::LAYOUT_GENERIC ? s.d_samples + sample_off[k] : nullptr;
  }
  s.set.build(items);
  ok = ok && cu_ok(b, s.set.upload(st), "H2D launch tables");
  for (size_t k = 0; k < expands.size() && ok; k++)
    ok = cu_ok(b, gidk::launch_expand(expands[k], s.geos[k].ncomp, st), "expand kernel");
  ok = ok && cu_ok(b, s.set.launch(w.num_sms, st), "kernel launch");
  // pixels: straight into the caller's buffer when it is pinned, else through the staging buffer
  for (size_t k = 0; k < nf && ok; k++) {
    uint8_t *dst = s.staged_out[k] ? s.h_out + s.out_off[k] : run->pixels[s.images[k]];
    ok = cu_ok(b, cudaMemcpyAsync(dst, s.d_out + s.out_off[k], s.geos[k].pixel_bytes, cudaMemcpyDeviceToHost, st), "D2H pixels");
  }
  ok = ok && cu_ok(b, cudaEventRecord(s.done, st), "cudaEventRecord");
  if (!ok) {
    (void)cudaStreamSynchronize(st);  // nothing of this chunk may still be reading the staging buffers
    slot_fail(run, s, GIDCUDA_ERR_CUDA);
    return !drained;  // keep draining the queue so that every image gets a status
  }
  run->h2d += moved_in + (int64_t)s.set.upload_bytes;
  for (size_t k = 0; k < nf; k++) run->d2h += (int64_t)s.geos[k].pixel_bytes;
  run->launches += s.set.launch_count() + (int32_t)expands.size();
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
    if (s.d_sparse) cudaFree(s.d_sparse);
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

static int batch_run(gidcuda_batch *b, BatchRun &run, int32_t n, const gidcuda_frame_desc *descs,
                     gidcuda_batch_stats *stats) {
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
  return batch_run(b, run, n, descs, stats);
}

extern "C" int gidcuda_batch_run_records(gidcuda_batch *b, int32_t n, const gidcuda_frame_desc *descs,
                                         const uint32_t *const *first, const uint32_t *const *records,
                                         const size_t *counts, uint8_t *const *pixels, int32_t *status,
                                         gidcuda_batch_stats *stats) {
  if (!b || n < 0 || (n > 0 && (!descs || !first || !records || !counts || !pixels)))
    return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_batch_run_records: bad argument");
  BatchRun run;
  run.n = n;
  run.descs = descs;
  run.first = first;
  run.records = records;
  run.counts = counts;
  run.pixels = pixels;
  run.status = status;
  return batch_run(b, run, n, descs, stats);
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
