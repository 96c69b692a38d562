/*
 * gidref.h -- ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement in plain C of the JPEG path of the Ada "Generic Image
 * Decoder" (zertovitch/gid, version 013).  Every function in gidref.c cites
 * the reference file:line it follows.  GID itself is Ada and cannot be built
 * in this environment (no GNAT), so this restatement is the checker the CUDA
 * path is compared with.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing under gid_b200/ may
 * include, link or dlopen it.
 *
 * Parity pinning: the reference ships no golden outputs for this path
 * (SURVEY.md section 8c).  The oracle is pinned against the survey-time
 * known-answer vectors KAT-1..KAT-10 (an independent restatement) and
 * cross-checked against libjpeg (via Pillow) within the expected IDCT /
 * upsampling tolerance; see tests/test_oracle_kat.py.
 */
#ifndef GIDREF_H
#define GIDREF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error classes = the Ada exceptions of gid.ads:73-77 (+ the two predefined
 * ones the JPEG path can propagate). */
enum {
  GIDREF_OK = 0,
  GIDREF_UNKNOWN_IMAGE_FORMAT = -1,
  GIDREF_KNOWN_BUT_UNSUPPORTED_IMAGE_FORMAT = -2,
  GIDREF_UNSUPPORTED_IMAGE_SUBFORMAT = -3,
  GIDREF_ERROR_IN_IMAGE_DATA = -4,
  GIDREF_INVALID_PRIMARY_COLOR_RANGE = -5,
  GIDREF_END_ERROR = -6,        /* Ada.IO_Exceptions.End_Error */
  GIDREF_CONSTRAINT_ERROR = -7, /* Constraint_Error */
  GIDREF_BAD_ARGUMENT = -8
};

/* Geometry of the frame as the patched decoder would hand it to the GPU
 * (same fields as gidcuda_frame_desc in include/gidcuda.h, kept separate on
 * purpose: the oracle must not depend on product headers). */
typedef struct gidref_frame {
  int32_t width, height;
  int32_t ncomp;          /* 1 (Y) or 3 (Y, Cb, Cr) */
  int32_t samp_h[3];      /* samples_hor per component (gid-decoding_jpg.adb:246) */
  int32_t samp_v[3];      /* samples_ver per component (:247) */
  int32_t blocks_x[3];    /* MCU-padded block grid per component */
  int32_t blocks_y[3];
  uint16_t qt[3][64];     /* natural order: qt_zz(c)(i) = qt(undo_zig_zag(i)), :1771-1775 */
  int32_t progressive;
  int32_t restart_interval;
  int32_t orientation;    /* 0 Unchanged, 1 Rotation_90, 2 Rotation_180, 3 Rotation_270 */
  int32_t scan_count;
  int32_t coef_overflow;  /* 1 if a coefficient did not fit int16 (outside the stated domain) */
} gidref_frame;

/* Client callbacks = generic formals of GID.Load_Image_Contents (gid.ads:117-143). */
typedef struct gidref_client {
  void *user;
  void (*set_x_y)(void *user, int x, int y);
  void (*put_pixel)(void *user, unsigned red, unsigned green, unsigned blue, unsigned alpha);
  void (*feedback)(void *user, int percents);
  unsigned modulus; /* Primary_Color_Range'Modulus: 256 or 65536; anything else raises */
} gidref_client;

/* Full GID decode (Load_Image_Header + Load_Image_Contents) with the client of
 * test/benchmark.adb:57-108: top-down packed RGB8, idx = 3*(x + W*(H-1-y)).
 * rgb must hold 3*W*H bytes; call gidref_probe first for W/H.
 * planes (optional, may be NULL): receives malloc'd int16 coefficient planes,
 * block-major [by][bx][64] natural order, un-dequantised, per component; free
 * with gidref_free.  feedback_log (optional): receives up to feedback_cap
 * Feedback values; *feedback_n the count actually issued. */
int gidref_probe(const uint8_t *data, size_t len, gidref_frame *frame, char *err, size_t errlen);
int gidref_decode_rgb(const uint8_t *data, size_t len, uint8_t *rgb, gidref_frame *frame,
                      int16_t **planes /* [3] */, int *feedback_log, int feedback_cap,
                      int *feedback_n, char *err, size_t errlen);
/* Same decode, arbitrary client (Set_X_Y / Put_Pixel / Feedback as function pointers). */
int gidref_decode_client(const uint8_t *data, size_t len, const gidref_client *client,
                         gidref_frame *frame, char *err, size_t errlen);

/* Reconstruction only: dequant + IDCT + upsampling + colour + output from int16
 * planes, following Finalize_Progressive_DCT_Decoding (gid-decoding_jpg.adb:1749-1851)
 * and the procedures it calls.  Output: top-down RGB8 with the given row stride. */
int gidref_reconstruct(const gidref_frame *frame, const int16_t *const planes[3],
                       uint8_t *rgb, size_t stride);

/* Building blocks, exported for known-answer tests. */
void gidref_idct_8x8(int32_t block[64]);            /* Inverse_DCT_8x8_Block, :790-907 */
void gidref_idct_8x8_variant(int32_t block[64], int no_row_shortcut, int no_col_shortcut,
                             int use_shifts);       /* for the survey's statistics only */
void gidref_ycbcr_to_rgb(int y, int cb, int cr, uint8_t out[3]); /* :1028-1035 */
const int32_t *gidref_zig_zag(void);                /* :570-578 */
const int32_t *gidref_undo_zig_zag(void);           /* :556-564 */

void gidref_free(void *p);
const char *gidref_version(void);

#ifdef __cplusplus
}
#endif
#endif
