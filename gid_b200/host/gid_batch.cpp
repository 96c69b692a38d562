// gid_batch.cpp -- many JPEG files on many host threads: the caller side of the
// reconstruction path when the entropy decode, not the GPU, sets the pace.
//
// GID is "task safe" (doc/gid.txt:20): no package-level state, one Image_Descriptor per
// task, so an application decodes N files with N tasks each calling Load_Image_Header +
// Load_Image_Contents.  This driver is that application for the host mirror: `threads`
// workers, each with its own gidcuda context (its own streams and pinned buffers) on
// device devices[w mod ndevices], take the next file from a shared counter, entropy-decode
// it on their core (gid-decoding_jpg.adb:1179-1220 / :1243-1747) straight into the pinned
// record arrays of a job, submit it, and -- while that frame crosses the link and is
// reconstructed on the B200 -- start on their next file.  Three frames are in flight per
// worker.  Finished pixels go to the caller's top-down RGB8 buffers (the buffer
// test/benchmark.adb:67-81 fills through Put_Pixel): straight from the device when a buffer
// is pinned host memory (gidcuda_host_alloc), through a copy otherwise.  No inter-GPU
// communication: images carry no cross-GPU data.
//
// A file that fails (any of GID's exceptions) is reported in status[] and skipped; the
// others are unaffected.  When there are fewer files than threads (one big frame), the
// spare threads go to the restart-interval-parallel scan of each file
// (Image_Descriptor::entropy_threads).
#include <chrono>
#include <deque>
#include <memory>
#include <mutex>

#pragma GCC diagnostic ignored "-Wunused-function"  // the header's readers that only gid_jpeg_host.cpp uses
#include "jpeg_entropy.hpp"

extern "C" {

typedef struct gidhost_batch_stats {
  double seconds;      // wall time of the whole call
  double megapixels;   // width * height / 1e6 over the images that decoded
  int64_t jpeg_bytes;  // compressed bytes of those images
  int32_t images_failed;
  int32_t threads;     // host threads used
  int32_t devices;
  int32_t entropy_threads;  // threads inside one image (restart-interval-parallel scan)
} gidhost_batch_stats;

int gidhost_status_of_current_exception(char *err, size_t errlen);  // gid_jpeg_host.cpp
int gidhost_option_device_entropy(void);
long gidhost_option_device_progressive_min_bytes(void);
int gidhost_option_sparse_input(void);
int gidhost_option_apply_orientation(void);

// Frames a worker keeps in flight: while it prepares frame k (header, entropy decode or -- when the
// device decodes the scan -- just finding the scan's end), frames k - 1 and k - 2 cross the link and
// are decoded / reconstructed.  Small frames are latency-bound (a dozen stream operations each).
constexpr int kFramesInFlight = 3;
// ... and of small frames (under a megapixel) twice as many: their time on the device is latency, not
// work, and only frames in flight hide it (512 x 512 files on 4 workers: 7.0 GP/s with 3 in flight).
constexpr int kSmallFramesInFlight = 6;
constexpr long kSmallFramePixels = 1L << 20;

// The workers' device contexts outlive a call: their pinned and device buffers are recycled
// from one batch to the next (gidcuda_job_release returns them to the context's pool).
struct gidhost_batch {
  std::vector<int32_t> devices;
  int threads = 1;
  std::vector<gidcuda_ctx *> ctx;  // one per worker, created by the worker on first use
};

int gidhost_batch_create(const int32_t *devices, int32_t ndevices, int32_t threads, gidhost_batch **batch) {
  if (!batch || ndevices < 0 || (ndevices > 0 && !devices) || threads < 0) return -8;
  gidhost_batch *b = new gidhost_batch;
  if (ndevices == 0) b->devices.push_back(0);
  for (int32_t i = 0; i < ndevices; i++) b->devices.push_back(devices[i]);
  b->threads = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
  b->ctx.assign((size_t)b->threads, nullptr);
  // GIDHOST_BLOCKING_WAIT=1: a waiting worker sleeps instead of spinning.  (Measured on 16 cores with 16
  // workers: spinning 16.2 GP/s, sleeping 14.4 GP/s on 1080p files -- so spinning stays the default.)
  if (const char *env = getenv("GIDHOST_BLOCKING_WAIT")) gidcuda_set_option("blocking_wait", atoi(env));
  *batch = b;
  return 0;
}

int gidhost_batch_destroy(gidhost_batch *b) {
  if (!b) return 0;
  for (gidcuda_ctx *c : b->ctx)
    if (c) gidcuda_destroy(c);
  delete b;
  return 0;
}

int gidhost_batch_decode(gidhost_batch *b, int32_t n, const uint8_t *const *data, const size_t *len,
                         uint8_t *const *rgb, const size_t *rgb_capacity, int32_t *status,
                         gidhost_batch_stats *stats, char *err, size_t errlen) {
  if (!b || n < 0 || (n > 0 && (!data || !len || !rgb || !rgb_capacity))) {
    if (err && errlen) snprintf(err, errlen, "gidhost_batch_decode: bad argument");
    return -8;
  }
  const auto t0 = std::chrono::steady_clock::now();
  int T = b->threads;
  const int per_image = n > 0 && n < T ? std::max(1, T / n) : 1;
  if (n < T) T = std::max(1, (int)n);
  const int nd = (int)b->devices.size();

  if (status)
    for (int32_t i = 0; i < n; i++) status[i] = -9;  // until a worker with a device context gets to it
  std::atomic<int32_t> next{0}, failed{0};
  std::atomic<int64_t> bytes{0};
  std::mutex mu;
  double megapixels = 0.0;
  std::string first_error;  // a worker that cannot even get a context
  int first_error_rc = 0;

  auto work = [&](int w) {
    const int dev = b->devices[(size_t)(w % nd)];
    const int rc = b->ctx[(size_t)w] ? GIDCUDA_OK : gidcuda_create(dev, &b->ctx[(size_t)w]);
    gidcuda_ctx *ctx = b->ctx[(size_t)w];
    if (rc != GIDCUDA_OK) {
      std::lock_guard<std::mutex> g(mu);
      if (first_error_rc == 0) {
        first_error_rc = -9;
        const char *m = gidcuda_last_error(nullptr);
        first_error = std::string("gidcuda_create: ") + (m ? m : "");
      }
      return;
    }
    // a frame in flight keeps its descriptor: should the device entropy decoder hand the scan
    // back, the image is decoded again on this thread (wait_pixels)
    struct InFlight {
      void *job = nullptr;
      int32_t index = -1;
      std::unique_ptr<gid::Stream> st;
      std::unique_ptr<gid::Image_Descriptor> im;
    };
    std::deque<InFlight> inflight;  // oldest first; kFramesInFlight - 1 of them while the next one is decoded
    double mp = 0.0;
    auto fail = [&](int32_t i) {
      char msg[256];
      const int code = gidhost_status_of_current_exception(msg, sizeof msg);
      if (status) status[i] = code;
      failed++;
    };
    auto finish = [&](InFlight &f) {
      if (!f.job) return;
      try {
        size_t stride = 0;
        uint8_t *out = rgb[f.index];
        const uint8_t *px = gid::detail::wait_pixels(*f.im, ctx, &f.job, &stride, out, rgb_capacity[f.index]);
        const bool turned = f.im->apply_orientation && (f.im->display_orientation == gid::Orientation::Rotation_90 ||
                                                        f.im->display_orientation == gid::Orientation::Rotation_270);
        const int fw = turned ? f.im->height : f.im->width, fh = turned ? f.im->width : f.im->height;
        const size_t row = (size_t)3 * (size_t)fw;
        if (px == out) {
          // the caller's buffer is pinned: the pixels arrived there straight from the device
        } else if (stride == row) memcpy(out, px, row * (size_t)fh);
        else
          for (int y = 0; y < fh; y++) memcpy(out + (size_t)y * row, px + (size_t)y * stride, row);
        mp += (double)fw * fh / 1e6;
        bytes += (int64_t)len[f.index];
        if (status) status[f.index] = 0;
      } catch (...) {
        fail(f.index);
      }
      if (f.job) gidcuda_job_release(static_cast<gidcuda_job *>(f.job));
      f.job = nullptr;
    };
    const std::function<void(int)> no_feedback = [](int) {};
    for (;;) {
      const int32_t i = next.fetch_add(1);
      if (i >= n) break;
      InFlight cur;
      try {
        cur.st.reset(new gid::Stream(data[i], len[i]));
        cur.im.reset(new gid::Image_Descriptor);
        gid::Image_Descriptor &im = *cur.im;
        gid::Load_Image_Header(im, *cur.st);
        im.device = dev;
        im.entropy_threads = per_image;
        im.device_entropy = gidhost_option_device_entropy() != 0;
        im.device_progressive_min_bytes = gidhost_option_device_progressive_min_bytes();
        im.sparse_input = gidhost_option_sparse_input() != 0;
        im.apply_orientation = gidhost_option_apply_orientation() != 0;
        if ((size_t)3 * (size_t)im.width * (size_t)im.height > rgb_capacity[i])
          throw gid::gid_error("output buffer too small");
        cur.index = i;
        cur.job = gid::detail::decode_and_submit(im, ctx, no_feedback, rgb[i], rgb_capacity[i]);
      } catch (...) {
        fail(i);
        continue;
      }
      const bool small = (long)cur.im->width * cur.im->height <= kSmallFramePixels;
      inflight.push_back(std::move(cur));
      while ((int)inflight.size() >= (small ? kSmallFramesInFlight : kFramesInFlight)) {  // the oldest frame: done by now, or nearly
        finish(inflight.front());
        inflight.pop_front();
      }
    }
    for (InFlight &f : inflight) finish(f);
    std::lock_guard<std::mutex> g(mu);
    megapixels += mp;
  };

  std::vector<std::thread> pool;
  for (int w = 1; w < T; w++) pool.emplace_back(work, w);
  if (n > 0) work(0);
  for (std::thread &t : pool) t.join();

  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    stats->megapixels = megapixels;
    stats->jpeg_bytes = bytes.load();
    stats->images_failed = failed.load();
    stats->threads = T;
    stats->devices = nd;
    stats->entropy_threads = per_image;
  }
  if (first_error_rc != 0) {
    if (err && errlen) snprintf(err, errlen, "%s", first_error.c_str());
    return first_error_rc;
  }
  return 0;
}

// One call: create, decode, destroy.
int gidhost_decode_batch(int32_t n, const uint8_t *const *data, const size_t *len, const int32_t *devices,
                         int32_t ndevices, int32_t threads, uint8_t *const *rgb, const size_t *rgb_capacity,
                         int32_t *status, gidhost_batch_stats *stats, char *err, size_t errlen) {
  gidhost_batch *b = nullptr;
  if (gidhost_batch_create(devices, ndevices, threads, &b) != 0) {
    if (err && errlen) snprintf(err, errlen, "gidhost_decode_batch: bad argument");
    return -8;
  }
  const int rc = gidhost_batch_decode(b, n, data, len, rgb, rgb_capacity, status, stats, err, errlen);
  gidhost_batch_destroy(b);
  return rc;
}

}  // extern "C"
