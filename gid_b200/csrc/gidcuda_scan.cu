// gidcuda_scan.cu -- the C-ABI shim's part for Huffman decoding on the device (gidcuda_job_scan,
// include/gidcuda.h): staging of the scan's bytes and code tables, the launches of entropy_kernel.cu,
// and -- for the self-synchronising decoder -- the host's repair of a chain of decoder states that has
// not settled.
#include "gidcuda_internal.hpp"

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
    wa.limbits[c] = gidk::fast_coef_bits(&job->desc, c);
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

// The scan's stream operations: compressed bytes to the device, planes zeroed, Huffman decoding (one
// thread per restart interval, or the self-synchronising decoder), status words back.
int job_submit_scan(gidcuda_job *job, cudaStream_t s) {
  gidcuda_ctx *ctx = job->ctx;
  const Geometry &g = job->geo;
  CU_TRY(ctx, cudaMemcpyAsync(job->d_scan, job->h_scan, job->scan_bytes, cudaMemcpyHostToDevice, s));
  // (status words: the self-synchronising decoder's first launch zeroes them itself, together with the
  // planes; the interval decoder writes every block whole and needs no zeroed planes)
  if (job->scan_ri != 0) CU_TRY(ctx, cudaMemsetAsync(job->d_status, 0, 16, s));
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
      sa.limbits[c] = gidk::fast_coef_bits(&job->desc, c);
      job->d_tail_zero[c] = false;
    }
    sa.ncomp = g.ncomp;
    sa.mcux = (uint32_t)g.mcux;
    sa.nmcu = (uint32_t)g.mcux * (uint32_t)g.mcuy;
    sa.ri = (uint32_t)job->scan_ri;
    sa.nintervals = (uint32_t)job->scan_nintervals;
    sa.status = job->d_status;
    CU_TRY(ctx, gidk::launch_huffman_intervals(sa, ctx->num_sms, s));
  }
  return GIDCUDA_OK;
}

// The decoder's status words travel AFTER the reconstruction kernel is in the stream (job_submit calls this once
// it is).  Ahead of it the 16-byte copy would wait its turn in the device-to-host copy engine behind the pixels
// of the other frames in flight, the kernel would wait for the copy, and the engine would then sit idle for as
// long as the kernel runs before this frame's pixels can follow: ~0.1 ms per 4096 x 4096 frame, 9 % of the link.
int job_submit_scan_status(gidcuda_job *job, cudaStream_t s) {
  gidcuda_ctx *ctx = job->ctx;
  CU_TRY(ctx, cudaMemcpyAsync(job->h_status, job->d_status, 16, cudaMemcpyDeviceToHost, s));
  return GIDCUDA_OK;
}

// After the job's work has finished: the device decoder's verdict on the scan.
int job_finish_scan(gidcuda_job *job) {
  gidcuda_ctx *ctx = job->ctx;
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
  if (job->scan_ri == 0 && job->h_status[0] == 0) {
    // a valid scan ends within a byte of its last code word (1-bits pad it to the marker, :640-662): whole bytes
    // left over, or code words fed from the padding (the reference reads the marker's own bytes there), are
    // what the serial decoder decides about
    const uint32_t nbits = (uint32_t)(job->scan_len * 8), end = job->h_status[3];
    if (end == 0 || end - 1u > nbits || nbits - (end - 1u) >= 8u) job->h_status[0] = 16;
  }
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
  return GIDCUDA_OK;
}


// ---------------------------------------------------------------------------
// progressive frames: gidcuda_job_pscan* (include/gidcuda.h)

constexpr uint32_t kMaxPscans = 64;
constexpr size_t kPscanTables = (6 * sizeof(gidk::HuffDev) + 255) / 256 * 256;
constexpr size_t kPargsBytes = kMaxPscans * (sizeof(gidk::PscanArgs) + 64) + 4096;

void job_free_pscan(gidcuda_job *job) {
  if (job->h_pscan) cudaFreeHost(job->h_pscan);
  if (job->d_pscan) cudaFree(job->d_pscan);
  if (job->h_pstatus) cudaFreeHost(job->h_pstatus);
  if (job->d_pstatus) cudaFree(job->d_pstatus);
  if (job->h_pargs) cudaFreeHost(job->h_pargs);
  if (job->d_pargs) cudaFree(job->d_pargs);
  job->h_pargs = job->d_pargs = nullptr;
  if (job->h_masks) cudaFreeHost(job->h_masks);
  if (job->d_masks) cudaFree(job->d_masks);
  if (job->h_refine) cudaFreeHost(job->h_refine);
  if (job->d_refine) cudaFree(job->d_refine);
  if (job->refine_copied) cudaEventDestroy(job->refine_copied);
  if (job->masks_ready) cudaEventDestroy(job->masks_ready);
  job->masks_ready = nullptr;
  job->h_pscan = job->d_pscan = nullptr;
  job->h_pstatus = job->d_pstatus = nullptr;
  job->h_masks = job->d_masks = job->h_refine = job->d_refine = nullptr;
  job->refine_copied = nullptr;
  job->cap_pscan = job->cap_masks = 0;
}

extern "C" int gidcuda_job_pscan_reserve(gidcuda_job *job, size_t total_bytes) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::BEGUN || job->pscan_active || job->scan_active)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_pscan_reserve: once, between gidcuda_job_begin and the first scan");
  if (total_bytes == 0 || total_bytes >= ((size_t)1 << 30))
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan_reserve: %zu bytes", total_bytes);
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  // every scan: its six code tables, its bytes, FF padding (the decoder looks a few bytes ahead)
  const size_t need = round_up(total_bytes, 256) + kMaxPscans * (kPscanTables + 512);
  if (job->cap_pscan < need) {
    if (job->h_pscan) cudaFreeHost(job->h_pscan);
    if (job->d_pscan) cudaFree(job->d_pscan);
    job->h_pscan = job->d_pscan = nullptr;
    job->cap_pscan = 0;
    const size_t cap = need + need / 4;
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_pscan, cap, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_pscan, cap));
    job->cap_pscan = cap;
  }
  if (!job->h_pstatus) {
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_pstatus, kMaxPscans * 16, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_pstatus, kMaxPscans * 16));
  }
  if (!job->h_pargs) {
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_pargs, kPargsBytes, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_pargs, kPargsBytes));
  }
  job->pargs_used = 0;
  size_t most = 0;
  const Geometry &g = job->geo;
  for (int c = 0; c < g.ncomp; c++) most = std::max(most, (size_t)g.bx[c] * g.by[c]);
  if (job->cap_masks < most) {
    if (job->h_masks) cudaFreeHost(job->h_masks);
    if (job->d_masks) cudaFree(job->d_masks);
    if (job->h_refine) cudaFreeHost(job->h_refine);
    if (job->d_refine) cudaFree(job->d_refine);
    job->h_masks = job->d_masks = job->h_refine = job->d_refine = nullptr;
    job->cap_masks = 0;
    const size_t cap = most + most / 4;
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_masks, cap * 8, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_masks, cap * 8));
    CU_TRY(ctx, cudaHostAlloc((void **)&job->h_refine, cap * 24, cudaHostAllocDefault));
    CU_TRY(ctx, cudaMalloc((void **)&job->d_refine, cap * 24));
    job->cap_masks = cap;
  }
  if (!job->refine_copied) CU_TRY(ctx, cudaEventCreateWithFlags(&job->refine_copied, cudaEventDisableTiming));
  if (!job->masks_ready) CU_TRY(ctx, cudaEventCreateWithFlags(&job->masks_ready, cudaEventDisableTiming | cudaEventBlockingSync));
  job->pscan_used = 0;
  job->pscan_flushed = 0;
  job->pscan_pending.clear();
  job->pscan_count = 0;
  job->refine_pending = false;
  job->pscan_active = true;
  // the planes accumulate from scan to scan: zero them once (progressive zero-fills, :1876-1882)
  cudaStream_t s = job->stream;
  CU_TRY(ctx, cudaMemsetAsync(job->d_coef, 0, g.coef_bytes, s));
  CU_TRY(ctx, cudaMemsetAsync(job->d_pstatus, 0, kMaxPscans * 16, s));
  for (int c = 0; c < 3; c++) job->d_tail_zero[c] = false;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_pscan(gidcuda_job *job, const gidcuda_pscan_desc *scan, const uint8_t *data, size_t len) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::BEGUN || !job->pscan_active)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_pscan: after gidcuda_job_pscan_reserve, before gidcuda_job_submit");
  const Geometry &g = job->geo;
  if (!scan || !data || len == 0 || scan->ss < 0 || scan->se > 63 || scan->ss > scan->se || scan->al < 0 || scan->al > 13 ||
      scan->ah < 0 || scan->ah > 13)
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan: null argument, no data, or a bad band / approximation");
  const bool dc = scan->ss == 0;
  if ((dc && scan->se != 0) || (!dc && scan->ncomp != 1) || (!dc && scan->ah != 0))
    return fail(ctx, GIDCUDA_ERR_UNSUPPORTED,
                "gidcuda_job_pscan: DC and AC terms in one scan, an AC scan with several components, or an AC refinement scan");
  if (scan->ncomp != 1 && scan->ncomp != g.ncomp)
    return fail(ctx, GIDCUDA_ERR_UNSUPPORTED, "gidcuda_job_pscan: a scan of %d of the frame's %d components", scan->ncomp, g.ncomp);
  for (int i = 0; i < scan->ncomp; i++)
    if (scan->comp[i] < 0 || scan->comp[i] >= g.ncomp || (scan->ncomp > 1 && scan->comp[i] != i))
      return fail(ctx, GIDCUDA_ERR_UNSUPPORTED, "gidcuda_job_pscan: components out of order or out of range");
  if (job->pscan_count >= kMaxPscans) return fail(ctx, GIDCUDA_ERR_UNSUPPORTED, "gidcuda_job_pscan: more than %u scans", kMaxPscans);
  const size_t slice = kPscanTables + round_up(len + 64, 256);
  if (job->pscan_used + slice > job->cap_pscan)
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan: the scans exceed what gidcuda_job_pscan_reserve was told");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  uint8_t *h = job->h_pscan + job->pscan_used, *d = job->d_pscan + job->pscan_used;
  gidk::PscanArgs a;
  memset(&a, 0, sizeof a);
  a.mode = dc ? (scan->ah == 0 ? 1u : 3u) : 2u;
  gidk::HuffDev *tabs = reinterpret_cast<gidk::HuffDev *>(h);
  memset(tabs, 0, 6 * sizeof(gidk::HuffDev));
  if (a.mode != 3u)
    for (int i = 0; i < scan->ncomp; i++) {
      const int c = scan->comp[i];
      if (!build_huff_dev(dc ? scan->dc[c] : scan->ac[c], &tabs[dc ? c : 3 + c]))
        return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan: Huffman table of component %d is not a prefix code", c);
    }
  memcpy(h + kPscanTables, data, len);
  memset(h + kPscanTables + len, 0xFF, slice - kPscanTables - len);
  a.data = d + kPscanTables;
  a.tables = reinterpret_cast<const gidk::HuffDev *>(d);
  a.ncomp = (uint32_t)scan->ncomp;
  uint32_t bpm = 0;
  for (int c = 0; c < g.ncomp; c++) {
    a.plane[c] = reinterpret_cast<int16_t *>(job->d_coef + g.plane_off[c]);
    a.row_blocks[c] = (uint32_t)g.bx[c];
    a.h[c] = job->desc.samp_h[c];
    a.v[c] = job->desc.samp_v[c];
    a.limbits[c] = gidk::fast_coef_bits(&job->desc, c);
    bpm += (uint32_t)(a.h[c] * a.v[c]);
  }
  for (int i = 0; i < scan->ncomp; i++) a.comps[i] = (uint32_t)scan->comp[i];
  if (scan->ncomp > 1) {
    if (bpm > 12) return fail(ctx, GIDCUDA_ERR_UNSUPPORTED, "gidcuda_job_pscan: %u blocks per MCU", bpm);
    a.bpm = bpm;
    a.mcux = (uint32_t)g.mcux;
    a.total_blocks = (uint32_t)g.mcux * (uint32_t)g.mcuy * bpm;
  } else {
    // a single-component scan walks the component's own block grid (8 x 8 MCUs, :1952-1971)
    const int c = scan->comp[0];
    const int wc = (job->desc.width * job->desc.samp_h[c] + g.hmax - 1) / g.hmax;
    const int hc = (job->desc.height * job->desc.samp_v[c] + g.vmax - 1) / g.vmax;
    a.bpm = 1;
    a.nbx = (uint32_t)((wc + 7) / 8);
    a.total_blocks = a.nbx * (uint32_t)((hc + 7) / 8);
  }
  a.ss = (uint32_t)scan->ss;
  a.se = (uint32_t)scan->se;
  a.al = (uint32_t)scan->al;
  a.nbits = (uint32_t)(len * 8);
  a.status = job->d_pstatus + 4 * job->pscan_count;
  if (a.mode == 3u) {
    if ((uint64_t)len * 8 < a.total_blocks)
      return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan: a DC refinement scan of %zu bytes for %u blocks", len, a.total_blocks);
    if ((uint64_t)len * 8 - a.total_blocks >= 8)  // whole bytes left over: the serial decoder's case
      return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan: a DC refinement scan of %zu bytes for %u blocks", len, a.total_blocks);
    job->pscan_expected[job->pscan_count] = 0;
    job->pscan_nbits[job->pscan_count] = 0;
  } else {
    a.sub_bits = (uint64_t)len * 8 > (uint64_t)a.total_blocks * kDenseBitsPerBlock ? kSubsequenceBitsDense : kSubsequenceBits;
    a.nsub = (uint32_t)((len * 8 + a.sub_bits - 1) / a.sub_bits);
    job->pscan_expected[job->pscan_count] = a.total_blocks;
    job->pscan_nbits[job->pscan_count] = a.nbits;
  }
  job->pscan_pending.push_back(a);  // on the stream at the next pscan_flush
  job->pscan_used += slice;
  job->pscan_count++;
  return GIDCUDA_OK;
}

// The staged scans onto the stream: their bytes in one copy, the first scans (DC first, AC first) of the group
// in four launches however many they are, then its DC refinement scans (which follow the DC first scan they
// refine; the first scans of a group store into different components or bands and are independent).
static int pscan_flush(gidcuda_job *job) {
  gidcuda_ctx *ctx = job->ctx;
  if (job->pscan_pending.empty()) return GIDCUDA_OK;
  cudaStream_t s = job->stream;
  std::vector<gidk::PscanArgs *> decoded;
  size_t sync_bytes = 0;
  for (gidk::PscanArgs &a : job->pscan_pending)
    if (a.mode != 3u) {
      decoded.push_back(&a);
      sync_bytes += round_up((size_t)a.nsub * (8 + 8 + 16 + 16), 256) + 512;
    }
  if (job->cap_sync < sync_bytes) {
    CU_TRY(ctx, cudaStreamSynchronize(s));  // (the arrays of an earlier group may still be in use on the stream)
    if (job->d_sync) cudaFree(job->d_sync);
    job->d_sync = nullptr;
    job->cap_sync = 0;
    const size_t cap = sync_bytes + sync_bytes / 2;
    CU_TRY(ctx, cudaMalloc((void **)&job->d_sync, cap));
    job->cap_sync = cap;
  }
  CU_TRY(ctx, cudaMemcpyAsync(job->d_pscan + job->pscan_flushed, job->h_pscan + job->pscan_flushed,
                              job->pscan_used - job->pscan_flushed, cudaMemcpyHostToDevice, s));
  job->pscan_flushed = job->pscan_used;
  const uint32_t ns = (uint32_t)decoded.size();
  if (ns) {
    const size_t need = round_up(ns * sizeof(gidk::PscanArgs), 256) + round_up(3 * (ns + 1) * 4, 256);
    if (job->pargs_used + need > kPargsBytes) return fail(ctx, GIDCUDA_ERR_UNSUPPORTED, "gidcuda_job_pscan: too many scans");
    uint8_t *h = job->h_pargs + job->pargs_used, *d = job->d_pargs + job->pargs_used;
    gidk::PscanArgs *ha = reinterpret_cast<gidk::PscanArgs *>(h);
    uint32_t *ht = reinterpret_cast<uint32_t *>(h + round_up(ns * sizeof(gidk::PscanArgs), 256));
    uint32_t ca = 0, cb = 0, co = 0;
    size_t off = 0;
    for (uint32_t k = 0; k < ns; k++) {
      gidk::PscanArgs &a = *decoded[k];
      const size_t n = a.nsub;
      a.entry = reinterpret_cast<uint64_t *>(job->d_sync + off);
      a.exit_ = a.entry + n;
      a.tuple = reinterpret_cast<int4 *>(job->d_sync + off + round_up(16 * n, 256));
      a.prefix = a.tuple + n;
      off += round_up(n * (8 + 8 + 16 + 16), 256) + 512;
      ha[k] = a;
      ht[k] = ca;
      ht[(ns + 1) + k] = cb;
      ht[2 * (ns + 1) + k] = co;
      ca += (a.nsub + gidk::kPscanWindow - 1) / gidk::kPscanWindow;
      cb += a.nsub > gidk::kPscanWindow / 2 ? (a.nsub - gidk::kPscanWindow / 2 + gidk::kPscanWindow - 1) / gidk::kPscanWindow : 0u;
      co += (a.nsub + gidk::kPscanOutputThreads - 1) / gidk::kPscanOutputThreads;
    }
    ht[ns] = ca;
    ht[(ns + 1) + ns] = cb;
    ht[2 * (ns + 1) + ns] = co;
    CU_TRY(ctx, cudaMemcpyAsync(d, h, need, cudaMemcpyHostToDevice, s));
    job->pargs_used += need;
    CU_TRY(ctx, gidk::launch_pscan_group(reinterpret_cast<const gidk::PscanArgs *>(d),
                                         reinterpret_cast<const uint32_t *>(d + round_up(ns * sizeof(gidk::PscanArgs), 256)), ns, ca,
                                         cb, co, s));
  }
  for (const gidk::PscanArgs &a : job->pscan_pending)
    if (a.mode == 3u) CU_TRY(ctx, gidk::launch_pscan_dc_refine(a, s));
  job->pscan_pending.clear();
  return GIDCUDA_OK;
}

// flags and block counts of the scans whose status words are in h_pstatus
static int pscan_verdict(gidcuda_job *job) {
  for (uint32_t i = 0; i < job->pscan_count; i++) {
    const uint32_t *st = job->h_pstatus + 4 * i;
    // a valid scan ends within a byte of its last code word (1-bits pad it to the marker, :640-662): whole
    // bytes left over, or code words fed from the padding, are what the serial decoder decides about
    const bool ends_badly = job->pscan_nbits[i] != 0 &&
                            (st[3] == 0 || st[3] - 1u > job->pscan_nbits[i] || job->pscan_nbits[i] - (st[3] - 1u) >= 8u);
    if (st[0] != 0 || (job->pscan_expected[i] != 0 && st[2] != job->pscan_expected[i]) || ends_badly) {
      job->state = gidcuda_job::BEGUN;
      job->pscan_active = false;
      return fail(job->ctx, GIDCUDA_ERR_DATA,
                  "device entropy decoder: scan %u of the progressive frame left to the host decoder (flags %u, %u of %u blocks)",
                  i, st[0], st[2], job->pscan_expected[i]);
    }
  }
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_pscan_masks(gidcuda_job *job, int comp, const uint64_t **masks, int32_t *nblocks) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::BEGUN || !job->pscan_active)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_pscan_masks: after gidcuda_job_pscan_reserve, before gidcuda_job_submit");
  if (comp < 0 || comp >= job->geo.ncomp || !masks)
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan_masks: bad component or null argument");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const Geometry &g = job->geo;
  cudaStream_t s = job->stream;
  const uint32_t nb = (uint32_t)g.bx[comp] * (uint32_t)g.by[comp];
  {
    const int rcf = pscan_flush(job);
    if (rcf != GIDCUDA_OK) return rcf;
  }
  CU_TRY(ctx, gidk::launch_pscan_masks(reinterpret_cast<const int16_t *>(job->d_coef + g.plane_off[comp]), nb, job->d_masks, s));
  // The host thread has nothing to do for this frame until the device has caught up with its first scans
  // (a quarter of a millisecond of kernels).  It spins, like gidcuda_job_wait -- measured: sleeping on a blocking
  // event costs 2 - 6 ms per wake-up on these boxes, cfg5 from files 7.5 against 10.2 GP/s on 16 threads --
  // unless the option "blocking_wait" says that the caller runs more threads than cores.
  // The copies back are only issued once the kernels are done: queued earlier they would sit at the head of the
  // device-to-host engine's FIFO meanwhile and hold up the pixels of every other frame behind them.
  if (g_blocking_wait) {
    CU_TRY(ctx, cudaEventRecord(job->masks_ready, s));
    CU_TRY(ctx, cudaEventSynchronize(job->masks_ready));
  } else {
    CU_TRY(ctx, cudaStreamSynchronize(s));
  }
  CU_TRY(ctx, cudaMemcpyAsync(job->h_masks, job->d_masks, (size_t)nb * 8, cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, cudaMemcpyAsync(job->h_pstatus, job->d_pstatus, (size_t)job->pscan_count * 16, cudaMemcpyDeviceToHost, s));
  CU_TRY(ctx, cudaStreamSynchronize(s));
  const int rc = pscan_verdict(job);   // a history built on a scan the device gave up on is worthless
  if (rc != GIDCUDA_OK) return rc;
  *masks = job->h_masks;
  if (nblocks) *nblocks = (int32_t)nb;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_pscan_refine(gidcuda_job *job, int comp, uint64_t **corr, uint64_t **newpos, uint64_t **newneg) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::BEGUN || !job->pscan_active)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_pscan_refine: after gidcuda_job_pscan_reserve, before gidcuda_job_submit");
  if (comp < 0 || comp >= job->geo.ncomp || !corr || !newpos || !newneg)
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan_refine: bad component or null argument");
  if (job->refine_pending) {  // the arrays of the previous refinement scan are still being copied
    CU_TRY(ctx, cudaEventSynchronize(job->refine_copied));
    job->refine_pending = false;
  }
  const size_t nb = (size_t)job->geo.bx[comp] * job->geo.by[comp];
  memset(job->h_refine, 0, nb * 24);
  *corr = job->h_refine;
  *newpos = job->h_refine + nb;
  *newneg = job->h_refine + 2 * nb;
  return GIDCUDA_OK;
}

extern "C" int gidcuda_job_pscan_refine_commit(gidcuda_job *job, int comp, int al) {
  if (!job) return fail(nullptr, GIDCUDA_ERR_BAD_ARGUMENT, "null job");
  gidcuda_ctx *ctx = job->ctx;
  if (job->state != gidcuda_job::BEGUN || !job->pscan_active)
    return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_pscan_refine_commit: after gidcuda_job_pscan_refine");
  if (comp < 0 || comp >= job->geo.ncomp || al < 0 || al > 13)
    return fail(ctx, GIDCUDA_ERR_BAD_ARGUMENT, "gidcuda_job_pscan_refine_commit: bad component or approximation");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const Geometry &g = job->geo;
  cudaStream_t s = job->stream;
  const size_t nb = (size_t)g.bx[comp] * g.by[comp];
  {
    const int rcf = pscan_flush(job);  // scans staged before this refinement run before it
    if (rcf != GIDCUDA_OK) return rcf;
  }
  CU_TRY(ctx, cudaMemcpyAsync(job->d_refine, job->h_refine, nb * 24, cudaMemcpyHostToDevice, s));
  CU_TRY(ctx, cudaEventRecord(job->refine_copied, s));
  job->refine_pending = true;
  CU_TRY(ctx, gidk::launch_pscan_apply(reinterpret_cast<int16_t *>(job->d_coef + g.plane_off[comp]), (uint32_t)nb, job->d_refine,
                                       job->d_refine + nb, job->d_refine + 2 * nb, (uint32_t)al, s));
  return GIDCUDA_OK;
}

int job_submit_pscan(gidcuda_job *job, cudaStream_t s) {
  gidcuda_ctx *ctx = job->ctx;
  if (job->pscan_count == 0) return fail(ctx, GIDCUDA_ERR_STATE, "gidcuda_job_submit: a progressive frame without scans");
  {
    const int rcf = pscan_flush(job);
    if (rcf != GIDCUDA_OK) return rcf;
  }
  return GIDCUDA_OK;
}

int job_submit_pscan_status(gidcuda_job *job, cudaStream_t s) {  // (after the reconstruction kernel: see job_submit_scan_status)
  gidcuda_ctx *ctx = job->ctx;
  CU_TRY(ctx, cudaMemcpyAsync(job->h_pstatus, job->d_pstatus, (size_t)job->pscan_count * 16, cudaMemcpyDeviceToHost, s));
  return GIDCUDA_OK;
}

int job_finish_pscan(gidcuda_job *job) {
  const int rc = pscan_verdict(job);
  job->pscan_active = false;  // a re-submission after a host decode takes planes / records
  return rc;
}
