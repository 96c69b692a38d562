// gid_jpeg_host.cpp -- host side of the JPEG path above the C ABI: segment
// readers, the entropy decoder (baseline and progressive) that emits frame-wide
// int16 coefficient planes, and the call sequence into libgidcuda.so.
//
// This is subsystem (1) of the design: "the entropy-decode loop now emits a
// frame-wide int16 coefficient plane per component plus quant tables instead of
// calling the IDCT inline".  It mirrors what the patched GID.Decoding_JPG does
// (gid_b200/ada/gid-decoding_jpg-cuda_patch.ada); references are to the GID 013
// tree.  Huffman decoding stays serial on the host, as upstream input.
//
//   segment reader            gid-decoding_jpg.adb:122-181
//   Read_SOF                  :184-318      Read_DHT :320-391     Read_DQT :393-422
//   Read_DRI                  :424-434      Read_EXIF :436-543
//   bit buffer / Get_VLC      :594-753
//   Check_Restart             :1133-1177
//   baseline scan             :1179-1220    progressive scans :1243-1747
//   Read_SOS + geometry       :1855-2034    segment loop :2043-2118
//   header loop               gid-headers.adb:429-474, signature :92-99
//
// Product code: independent of the test-only CPU restatement.
#include "jpeg_entropy.hpp"

namespace gid {

void Load_Image_Header(Image_Descriptor &image, Stream &from, bool /*try_tga*/) {
  image = Image_Descriptor();
  image.stream = &from;
  Reader r(from);
  // Load_Signature (gid-headers.adb:25-138): JPEG = FF D8 (:92-99).  The other
  // formats GID recognises are outside this path.
  if (from.size - from.pos < 2) throw unknown_image_format("unknown image format (too short)");
  const uint8_t s0 = r.byte(), s1 = r.byte();
  if (!(s0 == 0xFF && s1 == M_SOI)) {
    static const char *const others[] = {"BM", "GIF", "\x89PNG", "P1", "P2", "P3", "P4", "P5", "P6", "qoif"};
    for (const char *sig : others)
      if (from.size >= strlen(sig) && memcmp(from.data, sig, strlen(sig)) == 0)
        throw known_but_unsupported_image_format("image format is outside the JPEG path of this build");
    throw unknown_image_format("unknown image format");
  }
  image.format = Image_Format_Type::JPEG;
  // Load_JPEG_Header (gid-headers.adb:429-474): segments until the frame header
  image.restart_interval = 0;
  for (;;) {
    const Segment sg = read_segment(r, false, 0);
    if (sg.marker == M_DHT) read_dht(image, r, sg.length);
    else if (sg.marker == M_DQT) read_dqt(image, r, sg.length);
    else if (sg.marker == M_DRI) image.restart_interval = (int)r.be16();
    else if (sg.marker >= 0xC0 && sg.marker <= 0xCF && sg.marker != M_DHT) { read_sof(image, r, sg); break; }
    else if (sg.marker == M_APP1) read_exif(image, r, sg.length);
    else r.skip(sg.length);
  }
  image.stream_pos_after_header = from.pos;
}

namespace detail {

std::atomic<long> parallel_scans{0}, parallel_fallbacks{0}, device_scans{0}, device_scan_fallbacks{0};
std::atomic<long> ns_masks_wait{0}, ns_refine_host{0}, ns_pscan_queue{0}, selftest_refine_scans{0};
int g_selftest_refine = 0;  // gidhost_set_option("selftest_refine"): 1 = on, 2 = on with the portable decoder

// Frame descriptor from the header: geometry now, tables when the first scan starts.
static void describe_frame(const Image_Descriptor &image, gidcuda_frame_desc *desc, int comp_of[3]) {
  memset(desc, 0, sizeof *desc);
  desc->width = image.width;
  desc->height = image.height;
  int idx = 0;
  for (int c = 0; c < 3; c++)
    if (image.info[c].present) {
      desc->samp_h[idx] = image.info[c].samples_hor;
      desc->samp_v[idx] = image.info[c].samples_ver;
      comp_of[c] = idx++;
    }
  desc->ncomp = idx;
  desc->flags = GIDCUDA_OUT_RGB8 | GIDCUDA_FLAG_PACKED_ROWS;
  if (image.apply_orientation) switch (image.display_orientation) {
      case Orientation::Rotation_90: desc->flags |= GIDCUDA_FLAG_ROTATE_90; break;
      case Orientation::Rotation_180: desc->flags |= GIDCUDA_FLAG_ROTATE_180; break;
      case Orientation::Rotation_270: desc->flags |= GIDCUDA_FLAG_ROTATE_270; break;
      default: break;
    }
}

// natural-order tables, as Finalize_Progressive_DCT_Decoding builds qt_zz (:1771-1775)
static void fill_tables(const Image_Descriptor &image, gidcuda_frame_desc *desc, const int comp_of[3]) {
  for (int c = 0; c < 3; c++)
    if (image.info[c].present) {
      const int t = image.info[c].qt_assoc;
      if (t < 0 || t > 7) throw constraint_error("JPEG: quantization table index out of range");
      for (int k = 0; k < 64; k++) desc->qt[comp_of[c]][kZigZag[k]] = image.qt_list[t].q[k];
    }
}

void decode_planes(Image_Descriptor &image, gidcuda_frame_desc *desc, int16_t *planes[3], int32_t blocks_x[3],
                   int32_t blocks_y[3]) {
  if (!image.stream) throw constraint_error("decode_planes: no header was loaded");
  int comp_of[3] = {0, 0, 0};
  describe_frame(image, desc, comp_of);
  const std::function<void(int)> no_feedback = [](int) {};
  Decoder dec(image, no_feedback);
  dec.parallel_threads = image.entropy_threads > 0 ? image.entropy_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  dec.wide_progressive = image.wide_bit_buffer;
  dec.selftest_refine = g_selftest_refine != 0;
  dec.force_portable_refine = g_selftest_refine == 2;
  dec.begin_frame = [&]() {
    fill_tables(image, desc, comp_of);
    for (int c = 0; c < 3; c++)
      if (image.info[c].present) {
        int32_t bx = 0, by = 0;
        const int st = gidcuda_plane_geometry(desc, comp_of[c], &bx, &by);
        if (st != GIDCUDA_OK) raise_from_status(st, nullptr);
        int16_t *p = static_cast<int16_t *>(malloc((size_t)bx * (size_t)by * 64 * sizeof(int16_t)));
        if (!p) throw cuda_device_error("out of host memory");
        memset(p, 0, (size_t)bx * (size_t)by * 64 * sizeof(int16_t));
        planes[comp_of[c]] = p;
        if (blocks_x) blocks_x[comp_of[c]] = bx;
        if (blocks_y) blocks_y[comp_of[c]] = by;
        dec.plane[c].base = p;
        dec.plane[c].blocks_x = bx;
        dec.plane[c].blocks_y = by;
      }
  };
  dec.run();
  if (!dec.plane[0].base) throw error_in_image_data("JPEG: no scan in the image");
  selftest_refine_scans += dec.selftest_scans;
}

// Block (bx, by) with sequence number seq: MCUs in raster order, inside an MCU the
// component's h x v blocks row by row (the visiting order of :1190-1206).
static void seq_to_block(uint32_t seq, int h, int v, int mcux, int *bx, int *by) {
  const uint32_t hv = (uint32_t)(h * v), mcu = seq / hv, r = seq - mcu * hv;
  const uint32_t sy = r / (uint32_t)h, sx = r - sy * (uint32_t)h;
  *bx = (int)(mcu % (uint32_t)mcux) * h + (int)sx;
  *by = (int)(mcu / (uint32_t)mcux) * v + (int)sy;
}

// Progressive frames accumulate in dense planes over the scans; when the last scan is done the
// non-zero coefficients are compacted into records (the loop of Finalize_Progressive_DCT_Decoding,
// :1764-1775, visits every coefficient anyway).  False when they do not fit.
static bool compact_plane(const PlaneRef &pl, int h, int v, uint32_t *first, uint32_t *rec, size_t cap,
                          size_t *count) {
  const int mcux = pl.blocks_x / h;
  const uint32_t nb = (uint32_t)pl.blocks_x * (uint32_t)pl.blocks_y;
  size_t n = 0;
  for (uint32_t seq = 0; seq < nb; seq++) {
    int bx, by;
    seq_to_block(seq, h, v, mcux, &bx, &by);
    const int16_t *blk = pl.block(bx, by);
    first[seq] = (uint32_t)n;
    if (n + 64 > cap) return false;
    for (int i = 0; i < 64; i += 4) {  // four coefficients at a time: most groups are empty
      uint64_t four;
      memcpy(&four, blk + i, 8);
      if (four == 0) continue;
      for (int k = i; k < i + 4; k++)
        if (blk[k] != 0) rec[n++] = ((uint32_t)(uint16_t)blk[k] << 16) | (uint32_t)k;
    }
  }
  first[nb] = (uint32_t)n;
  *count = n;
  return true;
}

static int resolve_threads(int n) {
  if (n > 0) return n;
  const unsigned hw = std::thread::hardware_concurrency();
  return hw ? (int)hw : 1;
}

void *decode_and_submit(Image_Descriptor &image, void *ctx_, const std::function<void(int)> &feedback, uint8_t *out,
                        size_t out_capacity) {
  if (!image.stream) throw constraint_error("Load_Image_Contents: no header was loaded");
  gidcuda_ctx *ctx = static_cast<gidcuda_ctx *>(ctx_);
  gidcuda_frame_desc desc;
  int comp_of[3] = {0, 0, 0};
  describe_frame(image, &desc, comp_of);
  int rc;
  gidcuda_job *job = nullptr;
  // progressive + sparse input: the scans accumulate in ordinary host memory -- this thread's, kept from
  // one image to the next (a fresh 16 MB vector per image is thousands of page faults); it is only in
  // use until the records have been made, i.e. inside this call
  static thread_local std::vector<int16_t> accum[3];
  const bool sparse = image.sparse_input;
  // The job is begun at the first SOS, like Begin_Device_Frame of the Ada patch:
  // every DQT has been read by then (GID accepts none after entropy-coded data).
  Decoder dec(image, feedback);
  dec.parallel_threads = resolve_threads(image.entropy_threads);
  dec.wide_progressive = image.wide_bit_buffer;
  auto use_job_plane = [&](int c) {
    int32_t bx = 0, by = 0;
    dec.plane[c].base = gidcuda_job_plane(job, comp_of[c], &bx, &by);
    dec.plane[c].blocks_x = bx;
    dec.plane[c].blocks_y = by;
    if (!dec.plane[c].base) raise_from_status(GIDCUDA_ERR_STATE, ctx);
  };
  try {
    dec.begin_frame = [&]() {
      fill_tables(image, &desc, comp_of);
      const int st = gidcuda_job_begin(ctx, &desc, &job);
      if (st != GIDCUDA_OK) raise_from_status(st, ctx);
    };
    dec.prepare_host_output = [&]() {
      int st;
      for (int c = 0; c < 3; c++) {
        if (!image.info[c].present) continue;
        if (!sparse) {
          use_job_plane(c);
        } else if (image.progressive) {
          int32_t bx = 0, by = 0;
          st = gidcuda_plane_geometry(&desc, comp_of[c], &bx, &by);
          if (st != GIDCUDA_OK) raise_from_status(st, ctx);
          const size_t need = (size_t)bx * (size_t)by * 64;
          if (accum[c].size() < need) accum[c].resize(need);
          memset(accum[c].data(), 0, need * sizeof(int16_t));
          dec.plane[c].base = accum[c].data();
          dec.plane[c].blocks_x = bx;
          dec.plane[c].blocks_y = by;
        } else {
          RecordSink &k = dec.sink[c];
          int32_t nb = 0;
          st = gidcuda_job_records(job, comp_of[c], &k.first, &k.rec, &k.cap, &nb);
          if (st != GIDCUDA_OK) raise_from_status(st, ctx);
          k.nblocks = (uint32_t)nb;
          k.n = 0;
          k.seq = 0;
          k.on = true;
        }
      }
    };
    if (image.device_entropy && !image.progressive)
      dec.device_scan = [&](const Decoder::DeviceScan &ds) {
        gidcuda_scan_desc sd;
        memset(&sd, 0, sizeof sd);
        sd.restart_interval = ds.restart_interval;
        sd.nintervals = (int32_t)ds.begin.size();
        for (int c = 0; c < 3; c++) {
          if (!image.info[c].present) continue;
          gidcuda_huffman_table *t[2] = {&sd.dc[comp_of[c]], &sd.ac[comp_of[c]]};
          const std::vector<uint8_t> *bits[2] = {ds.dc_bits[c], ds.ac_bits[c]}, *vals[2] = {ds.dc_vals[c], ds.ac_vals[c]};
          for (int k = 0; k < 2; k++) {
            if (bits[k]->size() != 16 || vals[k]->size() > 256) return false;
            memcpy(t[k]->counts, bits[k]->data(), 16);
            memcpy(t[k]->symbols, vals[k]->data(), vals[k]->size());
          }
        }
        // (a table the device cannot take, e.g. an over-subscribed one, is a case for the host loop)
        if (gidcuda_job_scan(job, &sd, ds.data, ds.len, ds.restart_interval ? ds.begin.data() : nullptr,
                             ds.restart_interval ? ds.end.data() : nullptr) != GIDCUDA_OK)
          return false;
        device_scans++;
        return true;
      };
    if (image.device_entropy && image.progressive && image.device_progressive_min_bytes >= 0) {
      dec.device_pscan_min_bytes = (size_t)image.device_progressive_min_bytes;
      dec.device_pscan_begin = [&](size_t bytes) { return gidcuda_job_pscan_reserve(job, bytes) == GIDCUDA_OK; };
      dec.device_pscan = [&](const Decoder::DevicePscan &ds) {
        gidcuda_pscan_desc pd;
        memset(&pd, 0, sizeof pd);
        pd.ncomp = ds.ncomps;
        for (int k = 0; k < ds.ncomps; k++) pd.comp[k] = comp_of[ds.comps[k]];
        pd.ss = ds.ss; pd.se = ds.se; pd.ah = ds.ah; pd.al = ds.al;
        for (int c = 0; c < 3; c++) {
          if (!image.info[c].present) continue;
          gidcuda_huffman_table *t[2] = {&pd.dc[comp_of[c]], &pd.ac[comp_of[c]]};
          const std::vector<uint8_t> *bits[2] = {ds.dc_bits[c], ds.ac_bits[c]}, *vals[2] = {ds.dc_vals[c], ds.ac_vals[c]};
          for (int k = 0; k < 2; k++) {
            if (!bits[k]) continue;
            if (bits[k]->size() != 16 || vals[k]->size() > 256) return false;
            memcpy(t[k]->counts, bits[k]->data(), 16);
            memcpy(t[k]->symbols, vals[k]->data(), vals[k]->size());
          }
        }
        // (a table the device cannot take, a scan it does not cover: the host decodes the frame)
        return gidcuda_job_pscan(job, &pd, ds.data, ds.len) == GIDCUDA_OK;
      };
      dec.device_masks = [&](int c, int32_t *blocks_x, int32_t *nblocks) {
        const uint64_t *m = nullptr;
        int32_t by = 0;
        if (gidcuda_plane_geometry(&desc, comp_of[c], blocks_x, &by) != GIDCUDA_OK ||
            gidcuda_job_pscan_masks(job, comp_of[c], &m, nblocks) != GIDCUDA_OK)
          throw Decoder::device_declined();
        return m;
      };
      dec.device_refine_begin = [&](int c, uint64_t **corr, uint64_t **newpos, uint64_t **newneg) {
        if (gidcuda_job_pscan_refine(job, comp_of[c], corr, newpos, newneg) != GIDCUDA_OK) throw Decoder::device_declined();
      };
      dec.device_refine_commit = [&](int c, int al) {
        if (gidcuda_job_pscan_refine_commit(job, comp_of[c], al) != GIDCUDA_OK) throw Decoder::device_declined();
      };
    }
    // A baseline component outgrew its record buffer (more than 32 non-zero coefficients per
    // block on average): replay its records into the dense plane and carry on there.
    dec.spill = [&](int c) {
      RecordSink &k = dec.sink[c];
      use_job_plane(c);
      const Component_Info &ci = image.info[c];
      const int mcux = dec.plane[c].blocks_x / ci.samples_hor;
      for (uint32_t seq = 0; seq < k.seq; seq++) {
        int bx, by;
        seq_to_block(seq, ci.samples_hor, ci.samples_ver, mcux, &bx, &by);
        int16_t *blk = dec.plane[c].block(bx, by);
        const uint32_t end = seq + 1 < k.seq ? k.first[seq + 1] : (uint32_t)k.n;
        for (uint32_t i = k.first[seq]; i < end; i++) {
          const int nat = (int)(k.rec[i] & 63u);
          blk[nat] = (int16_t)(k.rec[i] >> 16);
          if ((nat >> 3) > dec.last_row[c]) dec.last_row[c] = nat >> 3;
        }
      }
      k.on = false;
    };
    try {
      dec.run();
    } catch (const Decoder::device_declined &) {
      // a scan of the progressive frame does not fit the device path: the host decoder takes the whole frame
      device_scan_fallbacks++;
      if (job) gidcuda_job_release(job);
      job = nullptr;
      image.stream->pos = image.stream_pos_after_header;
      const long keep = image.device_progressive_min_bytes;
      image.device_progressive_min_bytes = -1;
      const std::function<void(int)> no_feedback = [](int) {};
      void *again = nullptr;
      try {
        again = decode_and_submit(image, ctx_, no_feedback, out, out_capacity);
      } catch (...) {
        image.device_progressive_min_bytes = keep;
        throw;
      }
      image.device_progressive_min_bytes = keep;
      return again;
    }
    if (!job) throw error_in_image_data("JPEG: no scan in the image");
    if (dec.device_progressive) device_scans++;
    for (int c = 0; c < 3; c++) {
      if (!image.info[c].present || dec.device_scanned) continue;
      RecordSink &k = dec.sink[c];
      if (sparse && image.progressive) {
        uint32_t *first = nullptr, *rec = nullptr;
        size_t cap = 0, count = 0;
        rc = gidcuda_job_records(job, comp_of[c], &first, &rec, &cap, nullptr);
        if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
        const Component_Info &ci = image.info[c];
        if (compact_plane(dec.plane[c], ci.samples_hor, ci.samples_ver, first, rec, cap, &count)) {
          rc = gidcuda_job_records_commit(job, comp_of[c], count);
          if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
          continue;
        }
        const PlaneRef acc = dec.plane[c];
        use_job_plane(c);
        memcpy(dec.plane[c].base, acc.base, (size_t)acc.blocks_x * (size_t)acc.blocks_y * 64 * sizeof(int16_t));
      } else if (k.on) {
        if (k.seq != k.nblocks) throw error_in_image_data("JPEG: scan ended before the last block");
        k.first[k.nblocks] = (uint32_t)k.n;
        if (dec.runs.empty()) {
          rc = gidcuda_job_records_commit(job, comp_of[c], k.n);
        } else {  // one run per worker of the restart-parallel scan
          std::vector<size_t> rb, rn;
          for (const Decoder::Run &run : dec.runs) { rb.push_back(run.begin[c]); rn.push_back(run.count[c]); }
          rc = gidcuda_job_records_commit_runs(job, comp_of[c], (int32_t)rb.size(), rb.data(), rn.data());
        }
        if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
        continue;
      }
      rc = gidcuda_job_hint_rows(job, comp_of[c], dec.last_row[c] + 1);
      if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
    }
    // What the decoder stored, against what the planes and the fast kernel instances hold
    // (include/gidcuda.h, GIDCUDA_FLAG_WIDE_COEF).  The reference keeps 32-bit values all the way
    // (dc_predictor * qt_local (0), gid-decoding_jpg.adb:766-771, with differences of up to 15 bits, :727-733):
    // a term that has left the 16 bits of a plane cannot be handed over -- reported, not painted wrongly.
    if (!dec.device_scanned) {
      bool wide = false;
      for (int c = 0; c < 3; c++) {
        if (!image.info[c].present) continue;
        if (dec.mag[c] > 32767u)
          throw unsupported_image_subformat("JPEG: a coefficient of the stream exceeds 16 bits (DC predictor out of range)");
        if (dec.mag[c] > (uint32_t)gidcuda_fast_coef_limit(&desc, comp_of[c])) wide = true;
      }
      if (wide) {
        rc = gidcuda_job_wide_coefficients(job);
        if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
      }
    }
    if (out && gidcuda_host_is_pinned(out)) {
      rc = gidcuda_job_output(job, out, out_capacity);
      if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
    }
    rc = gidcuda_job_submit(job);
    if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
    return job;
  } catch (...) {
    if (job) gidcuda_job_release(job);
    throw;
  }
}

const uint8_t *wait_pixels(Image_Descriptor &image, void *ctx_, void **job, size_t *stride, uint8_t *out,
                           size_t out_capacity) {
  gidcuda_ctx *ctx = static_cast<gidcuda_ctx *>(ctx_);
  const uint8_t *px = nullptr;
  int rc = gidcuda_job_wait(static_cast<gidcuda_job *>(*job), &px, stride);
  if (rc == GIDCUDA_ERR_DATA) {
    // the device entropy decoder met something unusual: the host decoder takes the image, and
    // with it the reference's behaviour on damaged streams
    device_scan_fallbacks++;
    gidcuda_job_release(static_cast<gidcuda_job *>(*job));
    *job = nullptr;
    image.stream->pos = image.stream_pos_after_header;
    const bool keep = image.device_entropy;
    image.device_entropy = false;
    const std::function<void(int)> no_feedback = [](int) {};
    try {
      *job = decode_and_submit(image, ctx, no_feedback, out, out_capacity);
    } catch (...) {
      image.device_entropy = keep;
      throw;
    }
    image.device_entropy = keep;
    rc = gidcuda_job_wait(static_cast<gidcuda_job *>(*job), &px, stride);
  }
  if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
  return px;
}

// One device context per host thread, kept from one image to the next: its stream and the pinned
// and device buffers of its recycled jobs are what makes the second Load_Image_Contents of a
// thread cost a tenth of the first (GID is "task safe" with one descriptor per task; so is this:
// nothing is shared between threads).  Destroyed when the thread ends.
namespace {
struct ThreadContext {
  gidcuda_ctx *ctx = nullptr;
  int device = -1;
  ~ThreadContext() {
    if (ctx) gidcuda_destroy(ctx);
  }
  gidcuda_ctx *get(int dev) {
    if (ctx && device != dev) {
      if (gidcuda_destroy(ctx) != GIDCUDA_OK) return nullptr;  // (a job of another image is still out)
      ctx = nullptr;
    }
    if (!ctx) {
      const int rc = gidcuda_create(dev, &ctx);
      if (rc != GIDCUDA_OK) raise_from_status(rc, nullptr);
      device = dev;
    }
    return ctx;
  }
};
thread_local ThreadContext t_context;
}  // namespace

Reconstruction decode_and_reconstruct(Image_Descriptor &image, const std::function<void(int)> &feedback) {
  gidcuda_ctx *ctx = t_context.get(image.device);
  bool own = false;  // a context of its own, when the thread's is busy on another device
  int rc;
  if (!ctx) {
    rc = gidcuda_create(image.device, &ctx);
    if (rc != GIDCUDA_OK) raise_from_status(rc, nullptr);
    own = true;
  }
  gidcuda_job *job = nullptr;
  try {
    job = static_cast<gidcuda_job *>(decode_and_submit(image, ctx, feedback));
    Reconstruction rec;
    void *j = job;
    try {
      rec.pixels = wait_pixels(image, ctx, &j, &rec.stride);
    } catch (...) {
      job = static_cast<gidcuda_job *>(j);
      throw;
    }
    job = static_cast<gidcuda_job *>(j);
    rec.job = job;
    rec.ctx = own ? ctx : nullptr;  // release() destroys only a context of its own
    return rec;
  } catch (...) {
    if (job) gidcuda_job_release(job);
    if (own) gidcuda_destroy(ctx);
    throw;
  }
}

void release(Reconstruction &r) {
  if (r.job) gidcuda_job_release(static_cast<gidcuda_job *>(r.job));
  if (r.ctx) gidcuda_destroy(static_cast<gidcuda_ctx *>(r.ctx));
  r = Reconstruction();
}

}  // namespace detail

}  // namespace gid

// ---- flat C entry points (ctypes / other FFIs) ------------------------------------------

extern "C" {

enum {
  GIDHOST_OK = 0,
  GIDHOST_UNKNOWN_IMAGE_FORMAT = -1,
  GIDHOST_KNOWN_BUT_UNSUPPORTED_IMAGE_FORMAT = -2,
  GIDHOST_UNSUPPORTED_IMAGE_SUBFORMAT = -3,
  GIDHOST_ERROR_IN_IMAGE_DATA = -4,
  GIDHOST_INVALID_PRIMARY_COLOR_RANGE = -5,
  GIDHOST_CONSTRAINT_ERROR = -7,
  GIDHOST_CUDA_DEVICE_ERROR = -9,
  GIDHOST_BAD_ARGUMENT = -8
};

static int status_of_exception(char *err, size_t errlen) {
  int rc = GIDHOST_BAD_ARGUMENT;
  const char *what = "unknown exception";
  try {
    throw;
  } catch (const gid::unknown_image_format &e) { rc = GIDHOST_UNKNOWN_IMAGE_FORMAT; what = e.what();
  } catch (const gid::known_but_unsupported_image_format &e) { rc = GIDHOST_KNOWN_BUT_UNSUPPORTED_IMAGE_FORMAT; what = e.what();
  } catch (const gid::unsupported_image_subformat &e) { rc = GIDHOST_UNSUPPORTED_IMAGE_SUBFORMAT; what = e.what();
  } catch (const gid::error_in_image_data &e) { rc = GIDHOST_ERROR_IN_IMAGE_DATA; what = e.what();
  } catch (const gid::invalid_primary_color_range &e) { rc = GIDHOST_INVALID_PRIMARY_COLOR_RANGE; what = e.what();
  } catch (const gid::constraint_error &e) { rc = GIDHOST_CONSTRAINT_ERROR; what = e.what();
  } catch (const gid::cuda_device_error &e) { rc = GIDHOST_CUDA_DEVICE_ERROR; what = e.what();
  } catch (const std::exception &e) { what = e.what();
  } catch (...) {
  }
  if (err && errlen) snprintf(err, errlen, "%s", what);
  return rc;
}

// Load_Image_Header only: width, height, components, progressive, orientation.
int gidhost_probe(const uint8_t *data, size_t len, int32_t *width, int32_t *height, int32_t *ncomp,
                  int32_t *progressive, int32_t *orientation, char *err, size_t errlen) {
  try {
    gid::Stream st(data, len);
    gid::Image_Descriptor im;
    gid::Load_Image_Header(im, st);
    if (width) *width = gid::Pixel_Width(im);
    if (height) *height = gid::Pixel_Height(im);
    if (ncomp) *ncomp = gid::Subformat(im);
    if (progressive) *progressive = im.progressive ? 1 : 0;
    if (orientation) *orientation = (int)gid::Display_Orientation(im);
    return GIDHOST_OK;
  } catch (...) {
    return status_of_exception(err, errlen);
  }
}

// Load_Image_Header + Load_Image_Contents with the client of test/benchmark.adb:57-108
// (top-down packed RGB8, idx = 3 * (x + W * (H - 1 - y))).  modulus = 256 or 65536
// selects the Primary_Color_Range instance (16-bit values are stored as their high
// byte, = value / 257); anything else raises invalid_primary_color_range.
// Process-wide options of the flat entry points: "sparse_input" 1 (default) = coefficients reach
// the device as records, 0 = as dense planes (Image_Descriptor::sparse_input).
//   "entropy_threads" host threads for the entropy decode of one image (Image_Descriptor::entropy_threads;
//   default 1, 0 = one per hardware thread).
//   "device_entropy" 1 (default) = baseline scans with restart intervals are entropy-decoded on the device
static int g_sparse_input = 1;
static int g_entropy_threads = 1;
static int g_device_entropy = 1;
static int g_wide_bit_buffer = 1;  // "wide_bit_buffer": Image_Descriptor::wide_bit_buffer
static long g_device_progressive_min_bytes = 96 * 1024;  // "device_progressive_min_bytes": Image_Descriptor::device_progressive_min_bytes
static int g_apply_orientation = 0;  // "apply_orientation": gidhost_batch_decode writes the pixels turned as the EXIF tag says
int gidhost_set_option(const char *name, int value) {
  if (name && strcmp(name, "sparse_input") == 0) {
    g_sparse_input = value != 0;
    return GIDHOST_OK;
  }
  if (name && strcmp(name, "apply_orientation") == 0) {
    g_apply_orientation = value != 0;
    return GIDHOST_OK;
  }
  if (name && strcmp(name, "wide_bit_buffer") == 0) {
    g_wide_bit_buffer = value != 0;
    return GIDHOST_OK;
  }
  if (name && strcmp(name, "device_entropy") == 0) {
    g_device_entropy = value != 0;
    return GIDHOST_OK;
  }
  if (name && strcmp(name, "selftest_refine") == 0) {
    // decode_planes checks every AC refinement scan against its decode from the non-zero history alone (the
    // host half of the device path for progressive files; 2: with the portable form of that decoder)
    gid::detail::g_selftest_refine = value;
    return GIDHOST_OK;
  }
  if (name && strcmp(name, "device_progressive_min_bytes") == 0) {
    g_device_progressive_min_bytes = value;
    return GIDHOST_OK;
  }
  if (name && strcmp(name, "entropy_threads") == 0 && value >= 0) {
    g_entropy_threads = value;
    return GIDHOST_OK;
  }
  return GIDHOST_BAD_ARGUMENT;
}

int gidhost_decode_rgb(const uint8_t *data, size_t len, int device, uint32_t modulus, uint8_t *rgb,
                       size_t rgb_capacity, int32_t *feedback_log, int32_t feedback_cap, int32_t *feedback_n,
                       char *err, size_t errlen) {
  try {
    gid::Stream st(data, len);
    gid::Image_Descriptor im;
    gid::Load_Image_Header(im, st);
    im.device = device;
    im.sparse_input = g_sparse_input != 0;
    im.entropy_threads = g_entropy_threads;
    im.device_entropy = g_device_entropy != 0;
    im.device_progressive_min_bytes = g_device_progressive_min_bytes;
    im.wide_bit_buffer = g_wide_bit_buffer != 0;
    const int W = gid::Pixel_Width(im), H = gid::Pixel_Height(im);
    if ((size_t)3 * (size_t)W * (size_t)H > rgb_capacity) {
      if (err && errlen) snprintf(err, errlen, "output buffer too small");
      return GIDHOST_BAD_ARGUMENT;
    }
    size_t idx = 0;
    int nfb = 0;
    double next_frame = -1.0;
    auto set_x_y = [&](int x, int y) { idx = (size_t)3 * ((size_t)x + (size_t)W * (size_t)(H - 1 - y)); };
    auto fb = [&](int p) {
      if (feedback_log && nfb < feedback_cap) feedback_log[nfb] = p;
      nfb++;
    };
    if (modulus == 256) {
      gid::Load_Image_Contents<uint8_t>(im, next_frame, set_x_y,
          [&](uint8_t r8, uint8_t g8, uint8_t b8, uint8_t) { rgb[idx] = r8; rgb[idx + 1] = g8; rgb[idx + 2] = b8; idx += 3; }, fb);
    } else if (modulus == 65536) {
      gid::Load_Image_Contents<uint16_t>(im, next_frame, set_x_y,
          [&](uint16_t r16, uint16_t g16, uint16_t b16, uint16_t) {
            rgb[idx] = (uint8_t)(r16 / 257); rgb[idx + 1] = (uint8_t)(g16 / 257); rgb[idx + 2] = (uint8_t)(b16 / 257); idx += 3;
          }, fb);
    } else {
      throw gid::invalid_primary_color_range("JPEG: color range not supported");
    }
    if (feedback_n) *feedback_n = nfb;
    return next_frame == 0.0 ? GIDHOST_OK : GIDHOST_BAD_ARGUMENT;
  } catch (...) {
    return status_of_exception(err, errlen);
  }
}

// Entropy decode only: Load_Image_Header + every scan, into int16 coefficient
// planes in ordinary host memory (block-major, natural order, un-dequantised) plus
// the frame descriptor the reconstruction needs -- the input of gidcuda_batch_run /
// gidcuda_plan_create when many images are decoded on the host cores and
// reconstructed together.  planes[c] are calloc'd; release them with gidhost_free.
int gidhost_decode_planes(const uint8_t *data, size_t len, gidcuda_frame_desc *desc, int16_t *planes[3],
                          int32_t blocks_x[3], int32_t blocks_y[3], char *err, size_t errlen) {
  if (!desc || !planes) return GIDHOST_BAD_ARGUMENT;
  planes[0] = planes[1] = planes[2] = nullptr;
  try {
    gid::Stream st(data, len);
    gid::Image_Descriptor im;
    gid::Load_Image_Header(im, st);
    im.entropy_threads = g_entropy_threads;
    im.wide_bit_buffer = g_wide_bit_buffer != 0;
    gid::detail::decode_planes(im, desc, planes, blocks_x, blocks_y);
    return GIDHOST_OK;
  } catch (...) {
    for (int c = 0; c < 3; c++) { free(planes[c]); planes[c] = nullptr; }
    return status_of_exception(err, errlen);
  }
}

void gidhost_free(void *p) { free(p); }

// (for gid_batch.cpp) status code + message of the exception being handled
int gidhost_status_of_current_exception(char *err, size_t errlen) { return status_of_exception(err, errlen); }
int gidhost_option_device_entropy(void) { return g_device_entropy; }
long gidhost_option_device_progressive_min_bytes(void) { return g_device_progressive_min_bytes; }
int gidhost_option_sparse_input(void) { return g_sparse_input; }
int gidhost_option_apply_orientation(void) { return g_apply_orientation; }

// Process-wide counters: "parallel_scans" (baseline scans decoded restart-interval-parallel),
// "parallel_fallbacks" (parallel attempts that handed the scan back to the serial loop).
long gidhost_counter(const char *name) {
  if (name && strcmp(name, "parallel_scans") == 0) return gid::detail::parallel_scans.load();
  if (name && strcmp(name, "parallel_fallbacks") == 0) return gid::detail::parallel_fallbacks.load();
  if (name && strcmp(name, "device_scans") == 0) return gid::detail::device_scans.load();
  if (name && strcmp(name, "device_scan_fallbacks") == 0) return gid::detail::device_scan_fallbacks.load();
  if (name && strcmp(name, "selftest_refine_scans") == 0) return gid::detail::selftest_refine_scans.load();
  if (name && strcmp(name, "ns_masks_wait") == 0) return gid::detail::ns_masks_wait.load();
  if (name && strcmp(name, "ns_refine_host") == 0) return gid::detail::ns_refine_host.load();
  if (name && strcmp(name, "ns_pscan_queue") == 0) return gid::detail::ns_pscan_queue.load();
  return -1;
}

}  // extern "C"
