/*
 * gidcuda.h -- C ABI of the B200 JPEG pixel-reconstruction path for GID.
 *
 * This is the drop-in boundary.  The Ada package
 * GID.CUDA (gid_b200/ada/gid-cuda.ads) binds exactly these entry points with
 * pragma Import (C, ...); the patched GID.Decoding_JPG calls them where the
 * reference ran the reconstruction inline:
 *
 *   reference (file:line, relative to the GID tree)        replaced by
 *   ------------------------------------------------------  ------------------
 *   Decode_8x8_Block dequant, gid-decoding_jpg.adb:771,785  kernel (dequant)
 *   Inverse_DCT_8x8_Block, :790-907                         kernel (IDCT)
 *   Upsampling_and_Output*, :957-1124                       kernel (upsample)
 *   Color_Transformation_and_Output, :1018-1054             kernel (colour)
 *   call sites Baseline_DCT_Decoding_Scan :1199, :1206      gidcuda_job_*
 *   Finalize_Progressive_DCT_Decoding :1789-1850            gidcuda_job_*
 *
 * Conventions: plain C types only; every function returns an int status
 * (0 = ok, < 0 = error class below) and never throws; the message of the last
 * error is available from gidcuda_last_error.  A context belongs to one host
 * thread (GID is "task safe": one descriptor per task, doc/gid.txt:20).
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with GIDCUDA_ERR_NO_DEVICE.
 */
#ifndef GIDCUDA_H
#define GIDCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIDCUDA_ABI_VERSION 5

/* Status codes.  The Ada binding maps DATA -> GID.error_in_image_data,
 * UNSUPPORTED -> GID.unsupported_image_subformat, the others ->
 * GID.CUDA.Device_Error. */
enum {
  GIDCUDA_OK = 0,
  GIDCUDA_ERR_BAD_ARGUMENT = -1,
  GIDCUDA_ERR_UNSUPPORTED = -2, /* geometry the path does not cover */
  GIDCUDA_ERR_NO_DEVICE = -3,
  GIDCUDA_ERR_CUDA = -4,        /* a CUDA runtime call failed */
  GIDCUDA_ERR_OUT_OF_MEMORY = -5,
  GIDCUDA_ERR_STATE = -6,       /* call sequence violated (e.g. wait before submit) */
  GIDCUDA_ERR_DATA = -7         /* gidcuda_job_scan: the device entropy decoder met something it leaves to the
                                   host decoder (which then reproduces the reference's behaviour, error or not) */
};

/* Output pixel formats (flags, low byte).  RGB8 top-down is what
 * test/benchmark.adb:67-81 and test/mini.adb:55-69 build with Put_Pixel. */
#define GIDCUDA_OUT_RGB8 0u      /* top-down, 3 bytes per pixel: R, G, B */
#define GIDCUDA_OUT_BGR8_BOTTOM_UP 1u /* test/to_bmp.adb:94-146: bottom-up B, G, R, rows padded to 4 bytes */
#define GIDCUDA_OUT_RGBA8 2u     /* top-down R, G, B, A; A = 255 = the alpha GID passes to Put_Pixel
                                    (Primary_Color_Range'Last, gid-decoding_jpg.adb:909-938) */
#define GIDCUDA_OUT_MASK 0xFFu
/* Force stride = 3 * width (no row padding) for RGB8 even when that makes rows
 * unaligned; the default stride is the packed row (3 or 4 bytes per pixel) rounded up to 16 bytes. */
#define GIDCUDA_FLAG_PACKED_ROWS 0x100u
/* Apply the display orientation (EXIF tag 16#112#, Read_EXIF gid-decoding_jpg.adb:436-543, what
 * GID.Display_Orientation reports) while writing the pixels, as the client of test/to_bmp.adb does in
 * its Set_X_Y / Put_Pixel (:104-146).  With (x, r) the column and top-down row of a pixel of the
 * W x H frame, the output image is
 *     ROTATE_90  : H x W,  pixel -> column r,         row W - 1 - x   (counted from the top)
 *     ROTATE_180 : W x H,  pixel -> column W - 1 - x, row H - 1 - r
 *     ROTATE_270 : H x W,  pixel -> column H - 1 - r, row x
 * in the row order and byte order of the chosen format (for GIDCUDA_OUT_BGR8_BOTTOM_UP that is exactly
 * to_bmp's buffer).  Stride and size (gidcuda_output_geometry) are those of the rotated image.
 * All three are done by the fused kernels themselves, in instances of their own: the half turn writes mirrored
 * runs, rows bottom to top (the price of an unrotated frame); quarter turns store the samples of a block
 * transposed, emit its columns as rows of the turned image and work through the frame in tiles of 32 MCUs down
 * a column (0.75 - 0.9 of the unrotated rate on the device, the same rate end to end).  Frames outside the fast
 * domain (see GIDCUDA_FLAG_WIDE_COEF) and samplings without a fused kernel rotate in the two-pass kernels. */
#define GIDCUDA_FLAG_ROTATE_90 0x200u
#define GIDCUDA_FLAG_ROTATE_180 0x400u
#define GIDCUDA_FLAG_ROTATE_270 0x600u
#define GIDCUDA_FLAG_ROTATE_MASK 0x600u
/* Set by the entropy decoder when it has stored, in some component c, a coefficient whose magnitude
 * exceeds gidcuda_fast_coef_limit(desc, c) = 65536 / (largest entry of the component's table).  No JPEG of
 * 8-bit samples holds one (|coefficient * entry| stays near 1024); Decode_8x8_Block accepts magnitudes up
 * to 15 bits (gid-decoding_jpg.adb:727-733).  The fast kernel instances decide the row shortcut of
 * Inverse_DCT_8x8_Block (:803-812) on raw coefficients and leave the column shortcut (:858-862) out, which
 * is the reference's arithmetic while every |coefficient * entry| <= 2**16.  Frames that carry this flag --
 * like frames whose tables hold an entry of 0 or above 255 -- run the instances that decide on
 * de-quantised values per block and carry both shortcuts literally: they reproduce the reference's
 * 32-bit wrap-around for EVERY 16-bit coefficient and every table (a few percent slower).  The host mirror
 * and the device Huffman decoders track the range themselves (gidcuda_job_wide_coefficients); a caller
 * that fills planes or records by other means sets the flag unless it knows its coefficients are in range. */
#define GIDCUDA_FLAG_WIDE_COEF 0x800u

/* One JPEG frame, as the entropy decoder knows it after SOF/DQT/SOS
 * (Read_SOF gid-decoding_jpg.adb:184-318, Read_DQT :393-422). */
typedef struct gidcuda_frame_desc {
  int32_t width;        /* 1 .. 65535 (:206-215) */
  int32_t height;       /* 1 .. 65535 */
  int32_t ncomp;        /* 1 = Y (grey), 3 = Y, Cb, Cr (:298-317) */
  int32_t samp_h[3];    /* samples_hor per component, 1 .. 4 (:246) */
  int32_t samp_v[3];    /* samples_ver per component, 1 .. 4 (:247) */
  uint16_t qt[3][64];   /* quantisation table per component in NATURAL order:
                           qt[c][i] = qt_list(qt_assoc(c))(undo_zig_zag(i)), :1771-1775 */
  uint32_t flags;       /* GIDCUDA_OUT_* | GIDCUDA_FLAG_* */
  int32_t out_stride;   /* bytes per output row; 0 = library default */
} gidcuda_frame_desc;

typedef struct gidcuda_ctx gidcuda_ctx;
typedef struct gidcuda_job gidcuda_job;
typedef struct gidcuda_plan gidcuda_plan;
typedef struct gidcuda_batch gidcuda_batch;

/* ---- library / device ---------------------------------------------------- */

int gidcuda_abi_version(void);
/* Number of usable CUDA devices (0 when there is none; never fails). */
int gidcuda_device_count(void);
/* Message of the last error raised on this thread (ctx may be NULL). */
const char *gidcuda_last_error(const gidcuda_ctx *ctx);

/* Loading the library sets CUDA_DEVICE_MAX_CONNECTIONS=32 in the environment unless it is set already:
 * frames in flight live on one stream each, and the device runs as many streams side by side as it has
 * hardware work queues (8 by default).  The variable is read when the CUDA context is created; a
 * process that creates its context before loading the library sets it itself.
 *
 * Library options (process-wide; set before creating jobs / plans):
 *   "tma"  1 (default) = stage coefficients through TMA, 0 = plain 128-bit loads.
 *   "chain_repair"  1 (default) = when the self-synchronising Huffman decoder (gidcuda_job_scan without restart
 *          markers) is left with an unsettled chain of states -- a stretch of the scan that does not
 *          re-synchronise, e.g. a large flat region -- gidcuda_job_wait lets the host walk those stretches
 *          serially and runs the output pass again; 0 = report GIDCUDA_ERR_DATA at once.
 *   "blocking_wait"  0 (default) = gidcuda_job_wait spins (lowest latency); 1 = it sleeps until the
 *          frame is done (for callers with about as many host threads as cores). */
int gidcuda_set_option(const char *name, int value);
/* Process-wide statistics: "chain_repairs" (scans whose chain of states the host repaired),
 * "chain_repair_subsequences" (subsequences it walked for that); -1 for an unknown name. */
long gidcuda_counter(const char *name);

/* One context per host thread / Ada task.  Owns a device, streams and a pool
 * of pinned host + device buffers that jobs recycle. */
int gidcuda_create(int device, gidcuda_ctx **ctx);
int gidcuda_destroy(gidcuda_ctx *ctx);

/* Geometry helper: MCU-padded block grid of a component, = what GID computes
 * in Read_SOS (:1945-1951) times the sampling factors. */
int gidcuda_plane_geometry(const gidcuda_frame_desc *desc, int comp, int32_t *blocks_x,
                           int32_t *blocks_y);
/* Output geometry: stride in bytes and total size of the pixel buffer. */
int gidcuda_output_geometry(const gidcuda_frame_desc *desc, size_t *stride, size_t *bytes);

/* ---- one frame: the path GID.Decoding_JPG calls --------------------------- */

/* Begin a frame.  Allocates (or recycles) pinned, ZEROED coefficient planes
 * (baseline zeroes each block, :771; progressive zero-fills, :1876-1882). */
int gidcuda_job_begin(gidcuda_ctx *ctx, const gidcuda_frame_desc *desc, gidcuda_job **job);
/* Pinned host plane of component comp: int16, block-major
 * [block_row][block_col][64], natural order (index = 8 * row + col), NOT
 * de-quantised.  The entropy decoder stores dc_predictor / value at
 * zig_zag(k) (:771, :785) or the progressive accumulators (:1339, :1615). */
int16_t *gidcuda_job_plane(gidcuda_job *job, int comp, int32_t *blocks_x, int32_t *blocks_y);
/* Optional promise from the entropy decoder: in EVERY block of plane comp the
 * coefficients of rows >= rows_used (natural index >= 8 * rows_used) are zero.
 * The library then does not move those rows over the host link (a frame's planes
 * are 6 bytes per pixel as int16, twice its pixels; at ordinary qualities the lower
 * half of every chroma block is empty).  rows_used = 1 .. 8; 8 (the state after
 * gidcuda_job_begin) promises nothing.  Holds until changed or the job is released.
 * A decoder tracks it for free: rows_used = 1 + (largest zig_zag(k) it stored) / 8. */
int gidcuda_job_hint_rows(gidcuda_job *job, int comp, int rows_used);
/* Range of the fast kernel instances for component comp (see GIDCUDA_FLAG_WIDE_COEF), and the call by which
 * an entropy decoder that only learns the range while decoding -- after gidcuda_job_begin -- marks the
 * frame: same effect as the flag in the descriptor.  Between gidcuda_job_begin and gidcuda_job_submit. */
int32_t gidcuda_fast_coef_limit(const gidcuda_frame_desc *desc, int comp);
int gidcuda_job_wide_coefficients(gidcuda_job *job);
/* ---- sparse coefficient input (optional, per component) --------------------
 * What an entropy decoder produces is sparse: Decode_8x8_Block reads (run, value)
 * pairs until EOB (:774-786), a handful per block at ordinary qualities, while a
 * dense plane costs 128 bytes per block on the host link (the link, not the GPU,
 * bounds a frame's end-to-end time).  Instead of storing into the plane of
 * gidcuda_job_plane, the decoder may hand a component over as RECORDS:
 *     records[n] = (uint32_t)(uint16_t)value << 16 | natural_index     (index = zig_zag(k), 0 .. 63)
 *     first[seq] = n of the first record of block `seq`, first[nblocks] = number of records
 * with `seq` counting the component's blocks in the order an interleaved scan visits
 * them (:1190-1206): MCUs in raster order, inside an MCU the component's blocks row by
 * row -- the order Baseline_DCT_Decoding_Scan decodes them in.  A progressive decoder
 * compacts its accumulators in the same order when the last scan is done (the loop of
 * Finalize_Progressive_DCT_Decoding, :1764-1775, visits every coefficient anyway).
 * Zero-valued records are allowed; a natural index may appear once per block.
 * gidcuda_job_records lends the two pinned arrays of component comp (first: nblocks + 1
 * entries; records: *capacity entries = 32 per block, the size of the dense plane);
 * gidcuda_job_records_commit(job, comp, count) declares them complete.  A decoder that
 * runs out of capacity does not commit and fills the dense plane instead.  Components may
 * be mixed (some committed, some dense).  On submit the library moves first[] and the
 * records used, and a kernel rebuilds the dense plane in device memory. */
int gidcuda_job_records(gidcuda_job *job, int comp, uint32_t **first, uint32_t **records, size_t *capacity,
                        int32_t *nblocks);
int gidcuda_job_records_commit(gidcuda_job *job, int comp, size_t count);
/* The same, for a component whose records were produced by SEVERAL host threads, each decoding
 * its own group of restart intervals (Check_Restart, gid-decoding_jpg.adb:1133-1177: after RSTn the
 * predictors are reset and the bit buffer is byte-aligned, so the intervals decode independently).
 * Thread r stores into its own slice of the lent record array, records[run_begin[r] ..
 * run_begin[r] + run_count[r]), the runs ascending and disjoint; first[] numbers the records of the
 * CONCATENATION of the runs (first[0] = 0, first[nblocks] = sum of run_count).  On submit the
 * library copies every run to its place in that concatenation: no host-side compaction pass. */
int gidcuda_job_records_commit_runs(gidcuda_job *job, int comp, int32_t nruns, const size_t *run_begin,
                                    const size_t *run_count);
/* ---- entropy decoding on the device (optional; baseline scans with restart intervals) ------
 * The step BEFORE the path.  After RSTn the predictors are reset and the bit buffer is
 * byte-aligned (Check_Restart, gid-decoding_jpg.adb:1133-1177): the restart intervals of a baseline
 * scan are independent bit streams.  Instead of decoding them (Decode_8x8_Block, :758-788) and
 * handing over planes or records, the host may hand over the scan's COMPRESSED bytes and where
 * each interval lies; one device thread per interval decodes them into the job's planes, and the
 * reconstruction follows in the same submit.  Only a few hundredths of a byte per pixel cross the
 * host link.
 *   data, len          the entropy-coded bytes of the scan (copied by the call)
 *   begin[i], end[i]   interval i = data[begin[i] .. end[i]): from the byte after RSTn-1 (or the
 *                      first byte of the scan) to the FF of the next marker; inside an interval every
 *                      FF is followed by 00 (byte stuffing, :640-662)
 *   scan               restart interval in MCUs, and per component the Huffman tables of the scan
 *                      (Read_DHT :320-391: 16 counts, then the symbols in code order)
 * A scan WITHOUT restart markers (restart_interval = 0, nintervals = 0, begin = end = NULL) is taken
 * too: data = the scan's bytes up to the marker that ends it, with the byte stuffing REMOVED
 * (FF 00 -> FF).  The device cuts the bit stream into subsequences of 256 bits (1024 where blocks are long), one thread each,
 * and relies on Huffman codes re-synchronising: every thread decodes from a guessed state, then from
 * its predecessor's exit state until the chain of states stops changing; prefix sums give every
 * subsequence its first block and its DC predictors; a last pass stores the coefficients and
 * VERIFIES the chain (by induction from the first subsequence a verified chain is the serial
 * decode).  A chain that has not settled (a stretch that does not re-synchronise: flat regions make the
 * bit stream periodic) is repaired by the host, which walks those stretches serially, and verified
 * again (option "chain_repair"); one that still does not verify is reported like a damaged stream.
 * The MCU order, block order and predictor handling are those of Baseline_DCT_Decoding_Scan
 * (:1179-1220).  The device decoder does not reproduce the reference's handling of damaged streams:
 * when an interval holds a code that is in no table, a run past coefficient 63, too few or too many
 * bytes, gidcuda_job_wait returns GIDCUDA_ERR_DATA and the pixels are undefined; the caller then
 * decodes the scan on the host (planes or records) and submits the same job again.  The scan is consumed
 * by the submit that follows it: a job that is submitted again needs its input handed over again. */
typedef struct gidcuda_huffman_table {
  uint8_t counts[16];   /* codes of length 1 .. 16 */
  uint8_t symbols[256]; /* in code order */
} gidcuda_huffman_table;
typedef struct gidcuda_scan_desc {
  int32_t restart_interval;        /* MCUs per interval (DRI, :424-434); 0 = the scan has no restart markers */
  int32_t nintervals;              /* ceil(MCUs of the frame / restart_interval); 0 with restart_interval = 0 */
  gidcuda_huffman_table dc[3], ac[3]; /* per component of the frame */
} gidcuda_scan_desc;
int gidcuda_job_scan(gidcuda_job *job, const gidcuda_scan_desc *scan, const uint8_t *data, size_t len,
                     const uint32_t *begin, const uint32_t *end);

/* ---- the scans of a PROGRESSIVE frame on the device (optional) -------------------------------------
 * Progressive_DCT_Decoding_Scan (gid-decoding_jpg.adb:1243-1747) is, like the baseline scan, the step BEFORE the
 * path.  Its FIRST scans (Ah = 0: DC first :1263-1340, AC first :1381-1470) are ordinary Huffman streams and are
 * decoded by the device's self-synchronising decoder straight into the job's planes, which accumulate from
 * scan to scan in device memory; DC refinement scans (:1341-1380) are one bit per block and go there too.
 * AC REFINEMENT scans (:1560-1690) cannot be decoded in parallel (inside an EOB run every block still reads one
 * correction bit per coefficient that is already non-zero: the decoder's state includes the absolute block
 * number) and stay on the host -- which needs no coefficient values for that, only which coefficients are
 * non-zero: gidcuda_job_pscan_masks hands that history over (8 bytes per block), and the host answers with,
 * per block, the positions that read a correction bit of 1 and the positions that became +-(1 << Al).
 *   gidcuda_job_pscan_reserve   once per frame, before the first scan: an upper bound of the entropy-coded
 *                               bytes of all its scans (the file size will do)
 *   gidcuda_job_pscan           a first scan or a DC refinement scan: data = the scan's bytes up to the marker
 *                               that ends it, stuffing removed (FF 00 -> FF); copied by the call, staged: the scans staged so
 *                               far go to the device together (one copy, then their kernels) with the next
 *                               _masks, _refine_commit or gidcuda_job_submit.
 *                               Scans with restart markers are the host decoder's case.
 *   gidcuda_job_pscan_masks     waits for the scans queued so far; masks[b], b = by * blocks_x + bx in the
 *                               component's (MCU-padded) plane: bit k = the coefficient at zig-zag position k
 *                               is non-zero.  GIDCUDA_ERR_DATA: see below.
 *   gidcuda_job_pscan_refine    lends three zeroed arrays indexed like masks; _refine_commit(al) applies them:
 *                               corr = read a correction bit of 1 (adds 1 << al away from zero unless bit al is
 *                               set already, :1600-1612), newpos / newneg = became +(1 << al) / -(1 << al).
 * After the last scan gidcuda_job_submit reconstructs from the device planes; no coefficient crosses the link.
 * As with gidcuda_job_scan the device does not reproduce the reference's handling of damaged streams: a code
 * in no table, a coefficient outside the band, a value outside the planes' range, a chain of states that does
 * not settle, or a scan that does not fill its blocks make gidcuda_job_pscan_masks / gidcuda_job_wait return
 * GIDCUDA_ERR_DATA; the caller then decodes the frame on the host and submits the job again. */
typedef struct gidcuda_pscan_desc {
  int32_t ncomp;                       /* components of the scan: 1, or all components of the frame */
  int32_t comp[3];                     /* their indices in the frame */
  int32_t ss, se, ah, al;              /* spectral selection, successive approximation (Read_SOS :1855-1944) */
  gidcuda_huffman_table dc[3], ac[3];  /* per component of the FRAME; only those the scan uses are read */
} gidcuda_pscan_desc;
int gidcuda_job_pscan_reserve(gidcuda_job *job, size_t total_bytes);
int gidcuda_job_pscan(gidcuda_job *job, const gidcuda_pscan_desc *scan, const uint8_t *data, size_t len);
int gidcuda_job_pscan_masks(gidcuda_job *job, int comp, const uint64_t **masks, int32_t *nblocks);
int gidcuda_job_pscan_refine(gidcuda_job *job, int comp, uint64_t **corr, uint64_t **newpos, uint64_t **newneg);
int gidcuda_job_pscan_refine_commit(gidcuda_job *job, int comp, int al);

/* Optional: let the pixels land in the CALLER's buffer instead of one owned by the job -- the
 * image buffer a client fills through Put_Pixel (test/benchmark.adb:67-81) handed over whole, so
 * that no copy follows the device-to-host transfer.  host_pixels must be pinned host memory
 * (gidcuda_host_alloc, or registered with the CUDA runtime) of at least gidcuda_output_geometry
 * bytes, and stay valid until gidcuda_job_wait returns; the layout is the job's (format, stride).
 * Between gidcuda_job_begin and gidcuda_job_submit; NULL returns to the job's own buffer. */
int gidcuda_job_output(gidcuda_job *job, uint8_t *host_pixels, size_t capacity);
/* Asynchronous: host-to-device copies, reconstruction kernel, device-to-host
 * copy of the pixels, all on the job's own stream (two jobs in flight overlap
 * their transfers).  A job may be submitted again after gidcuda_job_wait. */
int gidcuda_job_submit(gidcuda_job *job);
/* Block until the frame is done.  *pixels points to pinned host memory owned
 * by the job (valid until release) -- or to the caller's buffer, after gidcuda_job_output --:
 * rows in the order of the output format, *stride bytes apart.  GIDCUDA_ERR_DATA: see gidcuda_job_scan. */
int gidcuda_job_wait(gidcuda_job *job, const uint8_t **pixels, size_t *stride);
/* Return the job's buffers to the context pool. */
int gidcuda_job_release(gidcuda_job *job);

/* ---- device-resident path (planes already in HBM) ------------------------- */

/* A plan is a prepared launch over n frames whose coefficient planes and
 * output buffers live in device memory (one kernel launch per sampling
 * layout present in the batch, all frames of a layout in the same launch). */
int gidcuda_plan_create(gidcuda_ctx *ctx, int32_t n, const gidcuda_frame_desc *descs,
                        const int16_t *const *d_planes /* [n][3] device pointers */,
                        uint8_t *const *d_pixels /* [n] device pointers */,
                        gidcuda_plan **plan);
/* Launch on the given CUDA stream (cudaStream_t as void*; NULL = context stream).
 * Asynchronous.  Launches of one plan are ordered one after the other: do not
 * run a plan concurrently with itself on two streams (its kernels share the
 * plan's tile counters). */
int gidcuda_plan_launch(gidcuda_plan *plan, void *cuda_stream);
/* Number of kernel launches one gidcuda_plan_launch performs. */
int gidcuda_plan_launch_count(const gidcuda_plan *plan);
/* Number of those launches that stage coefficients through TMA (0 when the
 * planes' alignment forced the plain-load kernel). */
int gidcuda_plan_uses_tma(const gidcuda_plan *plan);
int gidcuda_plan_destroy(gidcuda_plan *plan);
/* Wait for everything queued on the context stream. */
int gidcuda_synchronize(gidcuda_ctx *ctx);

/* ---- batch driver: independent images sharded over the GPUs of one box ---- */

typedef struct gidcuda_batch_stats {
  double seconds;            /* wall time of the whole run */
  double megapixels;         /* sum of width * height / 1e6 */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int32_t kernel_launches;
  int32_t images_failed;
  int32_t images_per_device[16];
} gidcuda_batch_stats;

/* One worker thread, one work queue and a ring of in-flight chunks per device;
 * no inter-GPU communication (images carry no cross-GPU data). */
int gidcuda_batch_create(const int32_t *devices, int32_t ndevices, int32_t chunk_images,
                         gidcuda_batch **batch);
/* planes[i*3 + c]: HOST pointer to plane c of image i; pixels[i]: HOST output buffer of
 * gidcuda_output_geometry bytes.  Buffers in pinned memory (gidcuda_host_alloc, or registered with the
 * CUDA runtime) are read and written by the copy engines directly -- no host-side copy at all; pageable
 * buffers go through the worker's pinned staging buffers (one host copy each way).  Three chunks of
 * chunk_images images are in flight per device.  status[i] (optional) receives the per-image status: a
 * failed image is reported and skipped, the others are unaffected; every image gets a status. */
int gidcuda_batch_run(gidcuda_batch *batch, int32_t n, const gidcuda_frame_desc *descs,
                      const int16_t *const *planes, uint8_t *const *pixels, int32_t *status,
                      gidcuda_batch_stats *stats);
/* The same with the components handed over as RECORDS (see gidcuda_job_records: what an entropy decoder
 * emits, a few bytes per block instead of 128): first[i*3 + c] = the component's nblocks + 1 offsets,
 * records[i*3 + c] = its records, counts[i*3 + c] = how many (= first[nblocks]).  The link carries the
 * records; a kernel rebuilds the planes on the device. */
int gidcuda_batch_run_records(gidcuda_batch *batch, int32_t n, const gidcuda_frame_desc *descs,
                              const uint32_t *const *first, const uint32_t *const *records, const size_t *counts,
                              uint8_t *const *pixels, int32_t *status, gidcuda_batch_stats *stats);
int gidcuda_batch_destroy(gidcuda_batch *batch);

/* Pinned host memory for callers that want the fast copy path. */
int gidcuda_host_alloc(size_t bytes, void **ptr);
int gidcuda_host_free(void *ptr);
/* 1 when ptr points into pinned (page-locked) host memory known to the CUDA runtime, else 0. */
int gidcuda_host_is_pinned(const void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* GIDCUDA_H */
