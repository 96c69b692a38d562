/*
 * gidref.c -- ORACLE (test infrastructure, NOT product code).  See gidref.h.
 *
 * Restatement in C of GID's JPEG decoder (reference: /root/reference,
 * GID version 013).  File:line citations are relative to the reference root.
 * Semantics that matter and are kept on purpose:
 *   - Ada "/" on Integer truncates toward zero; C "/" does too.  Never ">>".
 *   - Ada Integer is 32 bits on GNAT x86-64; in the benchmark build mode
 *     (Fast_unchecked, gid.gpr:72-74) overflow wraps.  Build with -fwrapv.
 *   - The row-pass "all AC zero" shortcut (gid-decoding_jpg.adb:810-812) is not
 *     equivalent to the general path; the column-pass one (:858-862) is.
 *
 * Build: see oracle/Makefile (gcc -O2 -fwrapv -fPIC -shared).
 */
#include "gidref.h"

#include <math.h>
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GIDREF_VERSION "gidref 1 (restates GID 013, 02-Mar-2024)"

/* ------------------------------------------------------------------------ */
/* Types mirroring gid.ads:262-333 (JPEG_Defs, JPEG_Stuff_Type)              */
/* ------------------------------------------------------------------------ */

enum { C_Y = 0, C_Cb = 1, C_Cr = 2, C_I = 3, C_Q = 4, N_COMPONENT = 5 };
enum { KIND_AC = 0, KIND_DC = 1 };
enum { CS_YCbCr, CS_Y_Grey, CS_CMYK };
enum { /* Upsampling_Profile_Type, gid.ads:277-283 */
  component_not_covered, hor_1_1_ver_1_1, hor_1_2_ver_1_2, hor_2_1_ver_2_1, other_profile
};

typedef struct { uint8_t bits, code; } VLC_Code;

typedef struct {
  int samples_hor, up_factor_x, samples_ver, up_factor_y;
} Upsampling_Parameters;

typedef struct {
  int qt_assoc, repeat, shape_x, shape_y;
  Upsampling_Parameters ups;
  int upsampling_profile;
} Info_per_Component_A;

typedef struct {
  /* input buffer (GID.Buffering; a memory stream here) */
  const uint8_t *data;
  size_t len, pos;
  /* descriptor fields used by the JPEG path */
  int32_t width, height;
  int subformat_id; /* number of components */
  int progressive, greyscale;
  int display_orientation;
  int compo_set[N_COMPONENT];
  int color_space;
  Info_per_Component_A info[N_COMPONENT];
  int max_samples_hor, max_samples_ver;
  int32_t qt_list[8][64]; /* stream (zig-zag) order, as read: :393-422 */
  VLC_Code *vlc_defs[2][8];
  int restart_interval;
  /* error handling (Ada exceptions) */
  jmp_buf jb;
  int err_code;
  char err_msg[256];
} Image;

__attribute__((noreturn)) static void raise_exc(Image *im, int code, const char *msg) {
  im->err_code = code;
  snprintf(im->err_msg, sizeof im->err_msg, "%s", msg);
  longjmp(im->jb, 1);
}

/* gid-buffering.adb:95-105.  Returns 0 at end of input instead of raising. */
static inline int try_get_byte(Image *im, uint8_t *b) {
  if (im->pos >= im->len) return 0;
  *b = im->data[im->pos++];
  return 1;
}
static inline uint8_t get_byte(Image *im) {
  uint8_t b;
  if (!try_get_byte(im, &b)) raise_exc(im, GIDREF_END_ERROR, "End_Error");
  return b;
}
/* gid-decoding_jpg.adb:37-56 */
static uint16_t big_endian_u16(Image *im) {
  uint16_t n = 0;
  for (int i = 0; i < 2; i++) n = (uint16_t)(n * 256 + get_byte(im));
  return n;
}

/* ------------------------------------------------------------------------ */
/* Constant tables, gid-decoding_jpg.adb:556-588                             */
/* ------------------------------------------------------------------------ */

static const int32_t undo_zig_zag[64] = {
   0,  1,  5,  6, 14, 15, 27, 28,  2,  4,  7, 13, 16, 26, 29, 42,
   3,  8, 12, 17, 25, 30, 41, 43,  9, 11, 18, 24, 31, 40, 44, 53,
  10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
  21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

static int32_t zig_zag[64];      /* :570-578, derived as the inverse permutation */
static int32_t zag_zig[64][2];   /* :580-588, (x, y) = (col, row) of zig-zag index */
static int tables_ready = 0;

static void init_tables(void) {
  if (tables_ready) return;
  for (int n = 0; n < 64; n++) zig_zag[undo_zig_zag[n]] = n;
  for (int k = 0; k < 64; k++) {
    zag_zig[k][0] = zig_zag[k] % 8;
    zag_zig[k][1] = zig_zag[k] / 8;
  }
  tables_ready = 1;
}

const int32_t *gidref_zig_zag(void) { init_tables(); return zig_zag; }
const int32_t *gidref_undo_zig_zag(void) { return undo_zig_zag; }

/* ------------------------------------------------------------------------ */
/* Segment reading, gid-decoding_jpg.adb:58-181                              */
/* ------------------------------------------------------------------------ */

typedef enum {
  M_SOI, M_SOF_0, M_SOF_1, M_SOF_2, M_SOF_3, M_SOF_5, M_SOF_6, M_SOF_7, M_SOF_8, M_SOF_9,
  M_SOF_10, M_SOF_11, M_SOF_13, M_SOF_14, M_SOF_15, M_DHT, M_DAC, M_DQT, M_DRI,
  M_RST, M_APP_1, M_APP_OTHER, M_COM, M_SOS, M_EOI, M_UNKNOWN
} Marker_Kind;

#define COM_code 0xFE
#define DHT_code 0xC4
#define DRI_code 0xDD
#define EOI_code 0xD9
#define SOS_code 0xDA

/* marker_id table, :73-107.  APP_15 (EF) and JPG0..13 are not in GID's list. */
static Marker_Kind classify_marker(uint8_t b) {
  switch (b) {
    case 0xD8: return M_SOI;
    case 0xC0: return M_SOF_0;  case 0xC1: return M_SOF_1;  case 0xC2: return M_SOF_2;
    case 0xC3: return M_SOF_3;  case 0xC5: return M_SOF_5;  case 0xC6: return M_SOF_6;
    case 0xC7: return M_SOF_7;  case 0xC8: return M_SOF_8;  case 0xC9: return M_SOF_9;
    case 0xCA: return M_SOF_10; case 0xCB: return M_SOF_11; case 0xCD: return M_SOF_13;
    case 0xCE: return M_SOF_14; case 0xCF: return M_SOF_15;
    case 0xCC: return M_DAC;    case DHT_code: return M_DHT; case 0xDB: return M_DQT;
    case DRI_code: return M_DRI;
    case 0xE1: return M_APP_1;
    case COM_code: return M_COM; case SOS_code: return M_SOS; case EOI_code: return M_EOI;
    default: break;
  }
  if (b >= 0xD0 && b <= 0xD7) return M_RST;
  if (b >= 0xE0 && b <= 0xEE) return M_APP_OTHER;
  return M_UNKNOWN;
}

typedef struct { uint16_t length; Marker_Kind kind; } Segment_Head;

/* Read, :122-167 */
static void read_segment_head(Image *im, int known_marker, uint8_t buffered_marker,
                              Segment_Head *head) {
  uint8_t b;
  if (known_marker) {
    b = buffered_marker;
  } else {
    b = get_byte(im);
    if (b != 0xFF) raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA,
                             "JPEG: expected marker prefix (16#FF#) here");
    b = get_byte(im);
  }
  Marker_Kind m = classify_marker(b);
  if (m == M_UNKNOWN) {
    char msg[96];
    snprintf(msg, sizeof msg, "JPEG: unknown marker here: 16#FF#, then 16#%02X#", b);
    raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, msg);
  }
  head->kind = m;
  if (m == M_EOI) {
    head->length = 0;
  } else {
    head->length = big_endian_u16(im);
    head->length = (uint16_t)(head->length - 2); /* U16 arithmetic */
  }
}

/* Skip_Segment_Data, :169-181 */
static void skip_segment_data(Image *im, const Segment_Head *head) {
  for (unsigned i = 1; i <= head->length; i++) (void)get_byte(im);
}

/* JPEG_Defs.Identify, gid.adb:225-242 */
static int identify_profile(const Upsampling_Parameters *p) {
  if (p->samples_hor == 1 && p->up_factor_x == 1 && p->samples_ver == 1 && p->up_factor_y == 1)
    return hor_1_1_ver_1_1;
  if (p->samples_hor == 1 && p->up_factor_x == 2 && p->samples_ver == 1 && p->up_factor_y == 2)
    return hor_1_2_ver_1_2;
  if (p->samples_hor == 2 && p->up_factor_x == 1 && p->samples_ver == 2 && p->up_factor_y == 1)
    return hor_2_1_ver_2_1;
  return other_profile;
}

/* Read_SOF, :184-318 */
static void read_sof(Image *im, const Segment_Head *sh) {
  switch (sh->kind) {
    case M_SOF_0: break;
    case M_SOF_2: im->progressive = 1; break;
    default:
      raise_exc(im, GIDREF_UNSUPPORTED_IMAGE_SUBFORMAT, "JPEG: image type not yet supported");
  }
  uint8_t bits_pp_primary = get_byte(im);
  if (bits_pp_primary != 8)
    raise_exc(im, GIDREF_UNSUPPORTED_IMAGE_SUBFORMAT,
              "JPEG: bits per primary color (not supported)");
  uint16_t h = big_endian_u16(im);
  uint16_t w = big_endian_u16(im);
  if (w == 0) raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: zero image width");
  if (h == 0) raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: zero image height");
  im->width = w;
  im->height = h;
  uint8_t b = get_byte(im);
  im->subformat_id = b;
  im->max_samples_hor = 0;
  im->max_samples_ver = 0;
  uint8_t id_base = 1;
  for (int i = 1; i <= im->subformat_id; i++) {
    b = get_byte(im);
    if (b == 0) id_base = 0; /* encoder off-by-one workaround, :227-234 */
    if ((uint8_t)(b - id_base) > C_Q)
      raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: SOF: invalid component ID");
    int compo = (uint8_t)(b - id_base);
    im->compo_set[compo] = 1;
    Info_per_Component_A *info = &im->info[compo];
    b = get_byte(im);
    info->ups.samples_hor = b / 16;
    info->ups.samples_ver = b % 16;
    info->repeat = info->ups.samples_hor * info->ups.samples_ver;
    info->shape_x = info->ups.samples_hor * 8;
    info->shape_y = info->ups.samples_ver * 8;
    if (info->ups.samples_hor > im->max_samples_hor) im->max_samples_hor = info->ups.samples_hor;
    if (info->ups.samples_ver > im->max_samples_ver) im->max_samples_ver = info->ups.samples_ver;
    info->qt_assoc = get_byte(im);
  }
  for (int c = 0; c < N_COMPONENT; c++) {
    if (im->compo_set[c]) {
      Info_per_Component_A *info = &im->info[c];
      if (info->ups.samples_hor == 0 || info->ups.samples_ver == 0)
        /* Ada: division by zero -> Constraint_Error */
        raise_exc(im, GIDREF_CONSTRAINT_ERROR, "JPEG: SOF: zero sampling factor");
      info->ups.up_factor_x = im->max_samples_hor / info->ups.samples_hor;
      info->ups.up_factor_y = im->max_samples_ver / info->ups.samples_ver;
      info->upsampling_profile = identify_profile(&info->ups);
    }
  }
  if ((int)sh->length < 6 + 3 * im->subformat_id)
    raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: SOF segment too short");
  const int *s = im->compo_set;
  if (s[C_Y] && s[C_Cb] && s[C_Cr] && !s[C_I] && !s[C_Q]) {
    im->color_space = CS_YCbCr;
  } else if (s[C_Y] && !s[C_Cb] && !s[C_Cr] && !s[C_I] && !s[C_Q]) {
    im->color_space = CS_Y_Grey;
    im->greyscale = 1;
  } else if (s[C_Y] && s[C_Cb] && s[C_Cr] && s[C_I] && !s[C_Q]) {
    im->color_space = CS_CMYK;
  } else {
    raise_exc(im, GIDREF_UNSUPPORTED_IMAGE_SUBFORMAT,
              "JPEG: only YCbCr, Y_Grey and CMYK color spaces are currently defined");
  }
  if (im->color_space == CS_CMYK)
    raise_exc(im, GIDREF_UNSUPPORTED_IMAGE_SUBFORMAT,
              "JPEG: CMYK color space is currently not properly decoded");
}

/* Read_DHT, :320-391 -- builds the 65 536-entry direct lookup table. */
static void read_dht(Image *im, int data_length) {
  int32_t remaining = data_length;
  while (remaining > 0) {
    uint8_t b = get_byte(im);
    remaining -= 1;
    int kind = (b >= 8) ? KIND_AC : KIND_DC;
    int ht_idx = b & 7;
    if (im->vlc_defs[kind][ht_idx] == NULL) {
      im->vlc_defs[kind][ht_idx] = (VLC_Code *)calloc(65536, sizeof(VLC_Code));
      if (!im->vlc_defs[kind][ht_idx]) raise_exc(im, GIDREF_CONSTRAINT_ERROR, "Storage_Error");
    }
    VLC_Code *vlc = im->vlc_defs[kind][ht_idx];
    int32_t counts[17];
    for (int i = 1; i <= 16; i++) {
      counts[i] = get_byte(im);
      remaining -= 1;
    }
    int32_t remain_vlc = 65536, spread = 65536;
    int32_t idx = 0;
    for (int code_len = 1; code_len <= 16; code_len++) {
      spread /= 2;
      int32_t current_count = counts[code_len];
      if (current_count > 0) {
        if (remaining < current_count)
          raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: DHT data too short [1]");
        remain_vlc -= current_count * spread;
        if (remain_vlc < 0)
          raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: DHT data too short [2]");
        for (int32_t i = current_count; i >= 1; i--) {
          b = get_byte(im);
          for (int32_t j = spread; j >= 1; j--) {
            vlc[idx].bits = (uint8_t)code_len;
            vlc[idx].code = b;
            idx++;
          }
        }
        remaining -= current_count;
      }
    }
    while (remain_vlc > 0) {
      remain_vlc--;
      vlc[idx].bits = 0;
      idx++;
    }
  }
}

/* Read_DQT, :393-422 -- entries stay in stream (zig-zag) order. */
static void read_dqt(Image *im, int data_length) {
  int remaining = data_length;
  while (remaining > 0) {
    uint8_t b = get_byte(im);
    remaining -= 1;
    int high_prec = b >= 8;
    int qt_idx = b & 7;
    for (int i = 0; i < 64; i++) {
      if (high_prec) {
        im->qt_list[qt_idx][i] = big_endian_u16(im);
        remaining -= 2;
      } else {
        im->qt_list[qt_idx][i] = get_byte(im);
        remaining -= 1;
      }
    }
  }
}

/* Read_DRI, :424-434 */
static void read_dri(Image *im) { im->restart_interval = big_endian_u16(im); }

/* Read_EXIF, :436-543 -- only the orientation tag is used. */
static void read_exif(Image *im, int data_length) {
  uint8_t b, orientation_value = 0;
  if (data_length < 6) {
    for (int i = 1; i <= data_length; i++) (void)get_byte(im);
    return;
  }
  static const char sig_ref[6] = {'E', 'x', 'i', 'f', 0, 0};
  char signature[6];
  for (int i = 0; i < 6; i++) signature[i] = (char)get_byte(im);
  if (memcmp(signature, sig_ref, 6) != 0) {
    for (int i = 7; i <= data_length; i++) (void)get_byte(im);
    return;
  }
  char endianness = (char)get_byte(im);
  for (int i = 8; i <= 14; i++) (void)get_byte(im);
  (void)get_byte(im); /* number of IFD0 entries: read, unused */
  (void)get_byte(im);
  int x = 17;
  while (x <= data_length - 12) {
    uint16_t tag = get_byte(im);
    b = get_byte(im);
    if (endianness == 'I') tag = (uint16_t)(tag + 0x100 * b);
    else tag = (uint16_t)(b + 0x100 * tag);
    for (int i = 3; i <= 8; i++) (void)get_byte(im);
    if (endianness == 'I') {
      orientation_value = get_byte(im);
      for (int i = 10; i <= 12; i++) (void)get_byte(im);
    } else {
      (void)get_byte(im);
      orientation_value = get_byte(im);
      (void)get_byte(im);
      (void)get_byte(im);
    }
    x += 12;
    if (tag == 0x112) {
      switch (orientation_value) {
        case 8: im->display_orientation = 1; break;  /* Rotation_90 */
        case 3: im->display_orientation = 2; break;  /* Rotation_180 */
        case 6: im->display_orientation = 3; break;  /* Rotation_270 */
        default: im->display_orientation = 0; break; /* Unchanged */
      }
      break;
    }
  }
  for (int i = x; i <= data_length; i++) (void)get_byte(im);
}

/* Load_Signature (JPEG arm only), gid-headers.adb:92-99, then
 * Load_JPEG_Header, gid-headers.adb:429-474. */
static void load_image_header(Image *im) {
  if (im->len < 1) raise_exc(im, GIDREF_END_ERROR, "End_Error");
  if (get_byte(im) != 0xFF) raise_exc(im, GIDREF_UNKNOWN_IMAGE_FORMAT, "unknown image format");
  if (get_byte(im) != 0xD8)
    raise_exc(im, GIDREF_UNKNOWN_IMAGE_FORMAT, "unknown image format (only JPEG is restated)");
  im->restart_interval = 0;
  for (;;) {
    Segment_Head head;
    read_segment_head(im, 0, 0, &head);
    switch (head.kind) {
      case M_DHT: read_dht(im, head.length); break;
      case M_DQT: read_dqt(im, head.length); break;
      case M_DRI: read_dri(im); break;
      case M_SOF_0: case M_SOF_1: case M_SOF_2: case M_SOF_3: case M_SOF_5: case M_SOF_6:
      case M_SOF_7: case M_SOF_8: case M_SOF_9: case M_SOF_10: case M_SOF_11: case M_SOF_13:
      case M_SOF_14: case M_SOF_15:
        read_sof(im, &head);
        return;
      case M_APP_1: read_exif(im, head.length); break;
      default: skip_segment_data(im, &head); break;
    }
  }
}

static void image_init(Image *im, const uint8_t *data, size_t len) {
  memset(im, 0, sizeof *im);
  im->data = data;
  im->len = len;
  init_tables();
}

static void image_dispose(Image *im) {
  for (int k = 0; k < 2; k++)
    for (int i = 0; i < 8; i++) {
      free(im->vlc_defs[k][i]);
      im->vlc_defs[k][i] = NULL;
    }
}

/* ------------------------------------------------------------------------ */
/* Clip, IDCT, colour: the reconstruction arithmetic                         */
/* ------------------------------------------------------------------------ */

/* Clip, :755-756 */
static inline int32_t clip(int32_t x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

#define W1 2841
#define W2 2676
#define W3 2408
#define W5 1609
#define W6 1108
#define W7 565

/* Row_IDCT, :799-845 */
static inline void row_idct(int32_t *blk) {
  int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8;
  x1 = blk[4] * 2048;
  x2 = blk[6];
  x3 = blk[2];
  x4 = blk[1];
  x5 = blk[7];
  x6 = blk[5];
  x7 = blk[3];
  if (x1 == 0 && x2 == 0 && x3 == 0 && x4 == 0 && x5 == 0 && x6 == 0 && x7 == 0) {
    int32_t val = blk[0] * 8;
    for (int i = 0; i < 8; i++) blk[i] = val;
    return;
  }
  x0 = (blk[0] * 2048) + 128;
  x8 = W7 * (x4 + x5);
  x4 = x8 + (W1 - W7) * x4;
  x5 = x8 - (W1 + W7) * x5;
  x8 = W3 * (x6 + x7);
  x6 = x8 - (W3 - W5) * x6;
  x7 = x8 - (W3 + W5) * x7;
  x8 = x0 + x1;
  x0 = x0 - x1;
  x1 = W6 * (x3 + x2);
  x2 = x1 - (W2 + W6) * x2;
  x3 = x1 + (W2 - W6) * x3;
  x1 = x4 + x6;
  x4 = x4 - x6;
  x6 = x5 + x7;
  x5 = x5 - x7;
  x7 = x8 + x3;
  x8 = x8 - x3;
  x3 = x0 + x2;
  x0 = x0 - x2;
  x2 = (181 * (x4 + x5) + 128) / 256;
  x4 = (181 * (x4 - x5) + 128) / 256;
  blk[0] = (x7 + x1) / 256;
  blk[1] = (x3 + x2) / 256;
  blk[2] = (x0 + x4) / 256;
  blk[3] = (x8 + x6) / 256;
  blk[4] = (x8 - x6) / 256;
  blk[5] = (x0 - x4) / 256;
  blk[6] = (x3 - x2) / 256;
  blk[7] = (x7 - x1) / 256;
}

/* Col_IDCT, :847-895 */
static inline void col_idct(int32_t *blk) {
  int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8;
  x1 = blk[8 * 4] * 256;
  x2 = blk[8 * 6];
  x3 = blk[8 * 2];
  x4 = blk[8 * 1];
  x5 = blk[8 * 7];
  x6 = blk[8 * 5];
  x7 = blk[8 * 3];
  if (x1 == 0 && x2 == 0 && x3 == 0 && x4 == 0 && x5 == 0 && x6 == 0 && x7 == 0) {
    int32_t val = clip(((blk[0] + 32) / 64) + 128);
    for (int row = 7; row >= 0; row--) blk[row * 8] = val;
    return;
  }
  x0 = (blk[0] * 256) + 8192;
  x8 = W7 * (x4 + x5) + 4;
  x4 = (x8 + (W1 - W7) * x4) / 8;
  x5 = (x8 - (W1 + W7) * x5) / 8;
  x8 = W3 * (x6 + x7) + 4;
  x6 = (x8 - (W3 - W5) * x6) / 8;
  x7 = (x8 - (W3 + W5) * x7) / 8;
  x8 = x0 + x1;
  x0 = x0 - x1;
  x1 = W6 * (x3 + x2) + 4;
  x2 = (x1 - (W2 + W6) * x2) / 8;
  x3 = (x1 + (W2 - W6) * x3) / 8;
  x1 = x4 + x6;
  x4 = x4 - x6;
  x6 = x5 + x7;
  x5 = x5 - x7;
  x7 = x8 + x3;
  x8 = x8 - x3;
  x3 = x0 + x2;
  x0 = x0 - x2;
  x2 = (181 * (x4 + x5) + 128) / 256;
  x4 = (181 * (x4 - x5) + 128) / 256;
  blk[8 * 0] = clip(((x7 + x1) / 16384) + 128);
  blk[8 * 1] = clip(((x3 + x2) / 16384) + 128);
  blk[8 * 2] = clip(((x0 + x4) / 16384) + 128);
  blk[8 * 3] = clip(((x8 + x6) / 16384) + 128);
  blk[8 * 4] = clip(((x8 - x6) / 16384) + 128);
  blk[8 * 5] = clip(((x0 - x4) / 16384) + 128);
  blk[8 * 6] = clip(((x3 - x2) / 16384) + 128);
  blk[8 * 7] = clip(((x7 - x1) / 16384) + 128);
}

/* Inverse_DCT_8x8_Block, :790-907 */
static inline void inverse_dct_8x8_block(int32_t *block) {
  for (int row = 0; row < 8; row++) row_idct(block + row * 8);
  for (int column = 0; column < 8; column++) col_idct(block + column);
}

void gidref_idct_8x8(int32_t block[64]) { inverse_dct_8x8_block(block); }

/* Variant used only to reproduce the survey's statistics (how many samples
 * change if a shortcut is dropped or if ">>" replaces "/").  Not GID. */
static inline int32_t dv(int32_t x, int k, int use_shifts) {
  return use_shifts ? (x >> k) : (x / (1 << k));
}
void gidref_idct_8x8_variant(int32_t block[64], int no_row_shortcut, int no_col_shortcut,
                             int use_shifts) {
  for (int row = 0; row < 8; row++) {
    int32_t *blk = block + row * 8;
    int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8;
    x1 = blk[4] * 2048; x2 = blk[6]; x3 = blk[2]; x4 = blk[1]; x5 = blk[7]; x6 = blk[5]; x7 = blk[3];
    if (!no_row_shortcut && !(x1 | x2 | x3 | x4 | x5 | x6 | x7)) {
      int32_t val = blk[0] * 8;
      for (int i = 0; i < 8; i++) blk[i] = val;
      continue;
    }
    x0 = (blk[0] * 2048) + 128;
    x8 = W7 * (x4 + x5); x4 = x8 + (W1 - W7) * x4; x5 = x8 - (W1 + W7) * x5;
    x8 = W3 * (x6 + x7); x6 = x8 - (W3 - W5) * x6; x7 = x8 - (W3 + W5) * x7;
    x8 = x0 + x1; x0 = x0 - x1;
    x1 = W6 * (x3 + x2); x2 = x1 - (W2 + W6) * x2; x3 = x1 + (W2 - W6) * x3;
    x1 = x4 + x6; x4 = x4 - x6; x6 = x5 + x7; x5 = x5 - x7;
    x7 = x8 + x3; x8 = x8 - x3; x3 = x0 + x2; x0 = x0 - x2;
    x2 = dv(181 * (x4 + x5) + 128, 8, use_shifts);
    x4 = dv(181 * (x4 - x5) + 128, 8, use_shifts);
    blk[0] = dv(x7 + x1, 8, use_shifts); blk[1] = dv(x3 + x2, 8, use_shifts);
    blk[2] = dv(x0 + x4, 8, use_shifts); blk[3] = dv(x8 + x6, 8, use_shifts);
    blk[4] = dv(x8 - x6, 8, use_shifts); blk[5] = dv(x0 - x4, 8, use_shifts);
    blk[6] = dv(x3 - x2, 8, use_shifts); blk[7] = dv(x7 - x1, 8, use_shifts);
  }
  for (int col = 0; col < 8; col++) {
    int32_t *blk = block + col;
    int32_t x0, x1, x2, x3, x4, x5, x6, x7, x8;
    x1 = blk[32] * 256; x2 = blk[48]; x3 = blk[16]; x4 = blk[8]; x5 = blk[56]; x6 = blk[40]; x7 = blk[24];
    if (!no_col_shortcut && !(x1 | x2 | x3 | x4 | x5 | x6 | x7)) {
      int32_t val = clip(dv(blk[0] + 32, 6, use_shifts) + 128);
      for (int r = 0; r < 8; r++) blk[r * 8] = val;
      continue;
    }
    x0 = (blk[0] * 256) + 8192;
    x8 = W7 * (x4 + x5) + 4; x4 = dv(x8 + (W1 - W7) * x4, 3, use_shifts); x5 = dv(x8 - (W1 + W7) * x5, 3, use_shifts);
    x8 = W3 * (x6 + x7) + 4; x6 = dv(x8 - (W3 - W5) * x6, 3, use_shifts); x7 = dv(x8 - (W3 + W5) * x7, 3, use_shifts);
    x8 = x0 + x1; x0 = x0 - x1;
    x1 = W6 * (x3 + x2) + 4; x2 = dv(x1 - (W2 + W6) * x2, 3, use_shifts); x3 = dv(x1 + (W2 - W6) * x3, 3, use_shifts);
    x1 = x4 + x6; x4 = x4 - x6; x6 = x5 + x7; x5 = x5 - x7;
    x7 = x8 + x3; x8 = x8 - x3; x3 = x0 + x2; x0 = x0 - x2;
    x2 = dv(181 * (x4 + x5) + 128, 8, use_shifts);
    x4 = dv(181 * (x4 - x5) + 128, 8, use_shifts);
    blk[0]  = clip(dv(x7 + x1, 14, use_shifts) + 128); blk[8]  = clip(dv(x3 + x2, 14, use_shifts) + 128);
    blk[16] = clip(dv(x0 + x4, 14, use_shifts) + 128); blk[24] = clip(dv(x8 + x6, 14, use_shifts) + 128);
    blk[32] = clip(dv(x8 - x6, 14, use_shifts) + 128); blk[40] = clip(dv(x0 - x4, 14, use_shifts) + 128);
    blk[48] = clip(dv(x3 - x2, 14, use_shifts) + 128); blk[56] = clip(dv(x7 - x1, 14, use_shifts) + 128);
  }
}

/* YCbCr arm of Color_Transformation_and_Output, :1028-1035 */
static inline void ycbcr_to_rgb(int32_t yv, int32_t cbv, int32_t crv, uint8_t *r, uint8_t *g,
                                uint8_t *b) {
  int32_t y_val = yv * 256;
  int32_t cb_val = cbv - 128;
  int32_t cr_val = crv - 128;
  *r = (uint8_t)clip((y_val + 359 * cr_val + 128) / 256);
  *g = (uint8_t)clip((y_val - 88 * cb_val - 183 * cr_val + 128) / 256);
  *b = (uint8_t)clip((y_val + 454 * cb_val + 128) / 256);
}

void gidref_ycbcr_to_rgb(int y, int cb, int cr, uint8_t out[3]) {
  ycbcr_to_rgb(y, cb, cr, &out[0], &out[1], &out[2]);
}

/* ------------------------------------------------------------------------ */
/* Load (gid-decoding_jpg.adb:590-2118): state shared by both instantiations */
/* ------------------------------------------------------------------------ */

typedef struct {
  int ht_idx_AC, ht_idx_DC;
  int width, height, stride;
  int32_t dc_predictor;
} Info_per_component_B; /* :700-705 */

#define MAX_REFINE_INDEX 2000 /* :1397 */

typedef struct {
  Image *im;
  /* bit buffer, :594-607 */
  uint32_t bit_buffer;
  int bit_buffer_length;
  uint8_t memo_marker;
  Info_per_component_B info_B[N_COMPONENT];
  int ssxmax, ssymax, sample_shape_max_x, sample_shape_max_y;
  int scan_compo_set[N_COMPONENT];
  int mcu_count, mcu_count_h, mcu_count_v, mcu_count_image_h, mcu_count_image_v;
  int rst_count;
  uint16_t next_rst;
  /* progressive, :1225-1241 */
  int32_t *image_array; /* (x, y, c) Ada order: ((x * ah) + y) * ncomp + (c_idx - 1) */
  int32_t array_width, array_height;
  int components_amount;
  int compo;
  int scan_count;
  /* macro block and flat scratch */
  int32_t *mb;   /* [(c * ssxmax + (sbx-1)) * ssymax + (sby-1)][64] */
  int32_t *flat; /* [(x * my + y) * N_COMPONENT + c], :963-964 */
  /* client */
  const gidref_client *client;    /* callback instantiation */
  uint8_t *bm_rgb; size_t bm_idx; /* benchmark instantiation, test/benchmark.adb:67-81 */
  size_t bm_stride;               /* = 3*W for the benchmark client */
  int *feedback_log; int feedback_cap, feedback_n;
  /* coefficient tap (what the patched front end would emit) */
  int16_t *planes[3];
  int plane_bx[3], plane_by[3];
  int coef_overflow;
  /* refine list, :1397-1400 */
  int32_t refine_point_x[MAX_REFINE_INDEX + 1], refine_point_y[MAX_REFINE_INDEX + 1];
  int refine_index_last;
} Load_State;

static inline int32_t *mb_block(Load_State *L, int c, int sbx, int sby) {
  return L->mb + (size_t)(((c * L->ssxmax) + (sbx - 1)) * L->ssymax + (sby - 1)) * 64;
}

static inline int32_t *ia(Load_State *L, int32_t x, int32_t y, int c_idx) {
  Image *im = L->im;
  if (x < 0 || x >= L->array_width || y < 0 || y >= L->array_height || c_idx < 1 ||
      c_idx > im->subformat_id)
    raise_exc(im, GIDREF_CONSTRAINT_ERROR, "image_array index out of range");
  return &L->image_array[((size_t)x * (size_t)L->array_height + (size_t)y) *
                             (size_t)im->subformat_id + (size_t)(c_idx - 1)];
}

static void tap_alloc(Load_State *L) {
  /* Plane geometry: MCU-padded grid of the frame (interleaved MCU counts),
   * blocks_x(c) = mcu_count_image_h * samples_hor(c) etc.  Not part of GID:
   * this is the layout the patched entropy loop writes. */
  Image *im = L->im;
  int mcu_w = 8 * im->max_samples_hor, mcu_h = 8 * im->max_samples_ver;
  int mh = (im->width + mcu_w - 1) / mcu_w, mv = (im->height + mcu_h - 1) / mcu_h;
  for (int c = 0; c < 3; c++) {
    if (!im->compo_set[c] || L->planes[c]) continue;
    L->plane_bx[c] = mh * im->info[c].ups.samples_hor;
    L->plane_by[c] = mv * im->info[c].ups.samples_ver;
    L->planes[c] = (int16_t *)calloc((size_t)L->plane_bx[c] * L->plane_by[c] * 64, sizeof(int16_t));
    if (!L->planes[c]) raise_exc(im, GIDREF_CONSTRAINT_ERROR, "Storage_Error");
  }
}

static inline void tap_store(Load_State *L, int c, int bx, int by, int n, int32_t v) {
  if (bx >= L->plane_bx[c] || by >= L->plane_by[c]) return;
  if (v < -32768 || v > 32767) L->coef_overflow = 1;
  L->planes[c][((size_t)by * L->plane_bx[c] + bx) * 64 + n] = (int16_t)v;
}

/* Initialize_Bit_Buffer, :602-607 */
static void initialize_bit_buffer(Load_State *L) {
  L->bit_buffer = 0;
  L->bit_buffer_length = 0;
  L->memo_marker = 0;
}

/* Show_Bits, :609-677 */
static inline int32_t show_bits(Load_State *L, int bits) {
  Image *im = L->im;
  if (bits == 0) return 0;
  while (L->bit_buffer_length < bits) {
    uint8_t newbyte, possible_marker;
    if (!try_get_byte(im, &newbyte)) goto end_error;
    L->bit_buffer_length += 8;
    L->bit_buffer = (L->bit_buffer << 8) + newbyte;
    if (newbyte == 0xFF) {
      if (!try_get_byte(im, &possible_marker)) goto end_error;
      if (possible_marker != 0) {
        L->bit_buffer_length += 8;
        L->bit_buffer = (L->bit_buffer << 8) + possible_marker;
        L->memo_marker = possible_marker;
        if (possible_marker == EOI_code || possible_marker == COM_code ||
            possible_marker == DHT_code || possible_marker == DRI_code ||
            (possible_marker >= 0xD0 && possible_marker <= 0xD7) ||
            possible_marker == SOS_code) {
          /* accepted */
        } else {
          raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA,
                    "JPEG: Invalid marker within filling of bit buffer");
        }
      }
    }
    continue;
  end_error: /* exception handler, :665-670 */
    L->bit_buffer_length += 8;
    L->bit_buffer = (L->bit_buffer << 8) + 0xFF;
  }
  return (int32_t)((L->bit_buffer >> (L->bit_buffer_length - bits)) & ((1u << bits) - 1u));
}

/* Skip_Bits, :679-688 */
static inline void skip_bits(Load_State *L, int bits) {
  if (L->bit_buffer_length < bits) (void)show_bits(L, bits);
  L->bit_buffer_length -= bits;
}

/* Get_Bits, :690-696 */
static inline int32_t get_bits(Load_State *L, int bits) {
  int32_t res = show_bits(L, bits);
  skip_bits(L, bits);
  return res;
}

static VLC_Code *vlc_table(Load_State *L, int kind, int idx) {
  if (idx < 0 || idx > 7 || L->im->vlc_defs[kind][idx] == NULL)
    raise_exc(L->im, GIDREF_CONSTRAINT_ERROR, "JPEG: access to an undefined Huffman table");
  return L->im->vlc_defs[kind][idx];
}

/* Get_VLC, :710-736 */
static inline void get_vlc(Load_State *L, const VLC_Code *vlc, uint8_t *code, int32_t *value_ret) {
  int32_t value = show_bits(L, 16);
  int bits = vlc[value].bits;
  if (bits == 0) raise_exc(L->im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: VLC table: bits = 0");
  skip_bits(L, bits);
  value = vlc[value].code;
  *code = (uint8_t)value;
  bits = value & 15;
  *value_ret = 0;
  if (bits != 0) {
    value = get_bits(L, bits);
    if (value < (int32_t)(1u << (bits - 1))) value = value + 1 - (int32_t)(1u << bits);
    *value_ret = value;
  }
}

/* Next_Huffval, :738-747 */
static inline int32_t next_huffval(Load_State *L, const VLC_Code *vlc) {
  int32_t value = show_bits(L, 16);
  int bits = vlc[value].bits;
  if (bits == 0) raise_exc(L->im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: Huffman value: bits = 0");
  skip_bits(L, bits);
  return vlc[value].code;
}

/* Bin_Twos_Complement, :749-753 */
static inline int32_t bin_twos_complement(int32_t value, int bit_length) {
  if (value < (int32_t)(1u << (bit_length - 1))) return value + 1 - (int32_t)(1u << bit_length);
  return value;
}

/* Decode_8x8_Block, :758-788.  (bx, by) only feed the coefficient tap. */
static void decode_8x8_block(Load_State *L, int c, int32_t *block, int bx, int by) {
  Image *im = L->im;
  int32_t value;
  int coef;
  uint8_t code;
  const int32_t *qt_local = im->qt_list[im->info[c].qt_assoc & 7];
  if (im->info[c].qt_assoc > 7) raise_exc(im, GIDREF_CONSTRAINT_ERROR, "qt_assoc out of range");
  get_vlc(L, vlc_table(L, KIND_DC, L->info_B[c].ht_idx_DC), &code, &value);
  L->info_B[c].dc_predictor += value;
  memset(block, 0, 64 * sizeof(int32_t));
  block[0] = L->info_B[c].dc_predictor * qt_local[0];
  if (L->planes[c]) {
    for (int n = 0; n < 64; n++) tap_store(L, c, bx, by, n, 0);
    tap_store(L, c, bx, by, 0, L->info_B[c].dc_predictor);
  }
  coef = 0;
  const VLC_Code *ac = vlc_table(L, KIND_AC, L->info_B[c].ht_idx_AC);
  for (;;) {
    get_vlc(L, ac, &code, &value);
    if (code == 0) break; /* EOB */
    if ((code & 0x0F) == 0 && code != 0xF0)
      raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: error in VLC AC code for de-quantization");
    coef = coef + (code >> 4) + 1;
    if (coef > 63)
      raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, "JPEG: coefficient for de-quantization is > 63");
    block[zig_zag[coef]] = value * qt_local[coef];
    if (L->planes[c]) tap_store(L, c, bx, by, zig_zag[coef], value);
    if (coef == 63) break;
  }
}

/* Check_Restart, :1133-1177 */
static void check_restart(Load_State *L, int do_reset_predictors) {
  Image *im = L->im;
  if (im->restart_interval > 0) {
    L->rst_count -= 1;
    if (L->rst_count == 0) {
      L->bit_buffer_length = (int)((uint32_t)L->bit_buffer_length & 0xF8u);
      uint16_t w = (uint16_t)get_bits(L, 16);
      uint16_t expected = (uint16_t)(0xFFD0 + L->next_rst);
      if (w != expected) {
        char msg[128];
        snprintf(msg, sizeof msg,
                 "JPEG: expected RST (restart) marker Nb %d; code found 16#%04X#; expected 16#%04X#",
                 L->next_rst, w, expected);
        raise_exc(im, GIDREF_ERROR_IN_IMAGE_DATA, msg);
      }
      L->next_rst = (uint16_t)((L->next_rst + 1) & 7);
      L->rst_count = im->restart_interval;
      if (do_reset_predictors)
        for (int c = 0; c < N_COMPONENT; c++) L->info_B[c].dc_predictor = 0;
    }
  }
}

/* Two instantiations of the generic part of Load: the inlined client of
 * test/benchmark.adb, and an arbitrary client through function pointers. */

#define INST(name) name##_bm
#define CL_SET_X_Y(L, x, y) \
  ((L)->bm_idx = (size_t)(L)->bm_stride * (size_t)((L)->im->height - 1 - (y)) + 3u * (size_t)(x))
#define CL_PUT_PIXEL8(L, r, g, b)         \
  do {                                    \
    (L)->bm_rgb[(L)->bm_idx] = (r);       \
    (L)->bm_rgb[(L)->bm_idx + 1] = (g);   \
    (L)->bm_rgb[(L)->bm_idx + 2] = (b);   \
    (L)->bm_idx += 3;                     \
  } while (0)
#define CL_FEEDBACK(L, p)                                                          \
  do {                                                                             \
    if ((L)->feedback_log && (L)->feedback_n < (L)->feedback_cap)                  \
      (L)->feedback_log[(L)->feedback_n] = (p);                                    \
    (L)->feedback_n++;                                                             \
  } while (0)
#include "gidref_load.inc"
#undef INST
#undef CL_SET_X_Y
#undef CL_PUT_PIXEL8
#undef CL_FEEDBACK

/* Out_Pixel_8, :909-938, for an arbitrary Primary_Color_Range modulus */
static inline void out_pixel_8_client(Load_State *L, uint8_t br, uint8_t bg, uint8_t bb) {
  const gidref_client *cl = L->client;
  if (cl->modulus == 256) {
    cl->put_pixel(cl->user, br, bg, bb, 255u);
  } else if (cl->modulus == 65536) {
    cl->put_pixel(cl->user, 257u * br, 257u * bg, 257u * bb, 65535u);
  } else {
    raise_exc(L->im, GIDREF_INVALID_PRIMARY_COLOR_RANGE, "JPEG: color range not supported");
  }
}
#define INST(name) name##_cb
#define CL_SET_X_Y(L, x, y) (L)->client->set_x_y((L)->client->user, (int)(x), (int)(y))
#define CL_PUT_PIXEL8(L, r, g, b) out_pixel_8_client((L), (r), (g), (b))
#define CL_FEEDBACK(L, p) (L)->client->feedback((L)->client->user, (int)(p))
#include "gidref_load.inc"
#undef INST
#undef CL_SET_X_Y
#undef CL_PUT_PIXEL8
#undef CL_FEEDBACK

/* ------------------------------------------------------------------------ */
/* Public entry points                                                       */
/* ------------------------------------------------------------------------ */

static void fill_frame(const Image *im, const Load_State *L, gidref_frame *f) {
  memset(f, 0, sizeof *f);
  f->width = im->width;
  f->height = im->height;
  f->ncomp = im->subformat_id;
  f->progressive = im->progressive;
  f->restart_interval = im->restart_interval;
  f->orientation = im->display_orientation;
  int mcu_w = 8 * im->max_samples_hor, mcu_h = 8 * im->max_samples_ver;
  int mh = (im->width + mcu_w - 1) / mcu_w, mv = (im->height + mcu_h - 1) / mcu_h;
  for (int c = 0; c < 3 && c < im->subformat_id; c++) {
    f->samp_h[c] = im->info[c].ups.samples_hor;
    f->samp_v[c] = im->info[c].ups.samples_ver;
    f->blocks_x[c] = mh * f->samp_h[c];
    f->blocks_y[c] = mv * f->samp_v[c];
    /* natural-order table, :1771-1775 */
    for (int i = 0; i < 64; i++)
      f->qt[c][i] = (uint16_t)im->qt_list[im->info[c].qt_assoc & 7][undo_zig_zag[i]];
  }
  if (L) {
    f->scan_count = L->scan_count;
    f->coef_overflow = L->coef_overflow;
  }
}

static void copy_err(const Image *im, char *err, size_t errlen) {
  if (err && errlen) snprintf(err, errlen, "%s", im->err_msg);
}

int gidref_probe(const uint8_t *data, size_t len, gidref_frame *frame, char *err, size_t errlen) {
  Image *im = (Image *)malloc(sizeof(Image));
  if (!im) return GIDREF_BAD_ARGUMENT;
  image_init(im, data, len);
  int rc = GIDREF_OK;
  if (setjmp(im->jb) == 0) {
    load_image_header(im);
    if (frame) fill_frame(im, NULL, frame);
  } else {
    rc = im->err_code;
    copy_err(im, err, errlen);
  }
  image_dispose(im);
  free(im);
  return rc;
}

static void load_state_dispose(Load_State *L, int keep_planes) {
  free(L->image_array);
  free(L->mb);
  free(L->flat);
  if (!keep_planes)
    for (int c = 0; c < 3; c++) free(L->planes[c]);
}

int gidref_decode_rgb(const uint8_t *data, size_t len, uint8_t *rgb, gidref_frame *frame,
                      int16_t **planes, int *feedback_log, int feedback_cap, int *feedback_n,
                      char *err, size_t errlen) {
  Image *im = (Image *)malloc(sizeof(Image));
  Load_State *L = (Load_State *)calloc(1, sizeof(Load_State));
  if (!im || !L) { free(im); free(L); return GIDREF_BAD_ARGUMENT; }
  image_init(im, data, len);
  L->im = im;
  L->bm_rgb = rgb;
  L->feedback_log = feedback_log;
  L->feedback_cap = feedback_cap;
  volatile int want_planes = planes != NULL;
  int rc = GIDREF_OK;
  if (setjmp(im->jb) == 0) {
    load_image_header(im);
    L->bm_stride = 3u * (size_t)im->width;
    if (want_planes) tap_alloc(L);
    load_bm(L);
    if (frame) fill_frame(im, L, frame);
  } else {
    rc = im->err_code;
    copy_err(im, err, errlen);
    if (frame) fill_frame(im, L, frame);
  }
  if (feedback_n) *feedback_n = L->feedback_n;
  if (planes) {
    for (int c = 0; c < 3; c++) planes[c] = (rc == GIDREF_OK) ? L->planes[c] : NULL;
  }
  load_state_dispose(L, rc == GIDREF_OK && planes != NULL);
  image_dispose(im);
  free(L);
  free(im);
  return rc;
}

int gidref_decode_client(const uint8_t *data, size_t len, const gidref_client *client,
                         gidref_frame *frame, char *err, size_t errlen) {
  if (!client || !client->set_x_y || !client->put_pixel || !client->feedback)
    return GIDREF_BAD_ARGUMENT;
  Image *im = (Image *)malloc(sizeof(Image));
  Load_State *L = (Load_State *)calloc(1, sizeof(Load_State));
  if (!im || !L) { free(im); free(L); return GIDREF_BAD_ARGUMENT; }
  image_init(im, data, len);
  L->im = im;
  L->client = client;
  int rc = GIDREF_OK;
  if (setjmp(im->jb) == 0) {
    load_image_header(im);
    load_cb(L);
    if (frame) fill_frame(im, L, frame);
  } else {
    rc = im->err_code;
    copy_err(im, err, errlen);
  }
  load_state_dispose(L, 0);
  image_dispose(im);
  free(L);
  free(im);
  return rc;
}

/* Reconstruction only.  Builds the state Finalize_Progressive_DCT_Decoding
 * (:1749-1851) runs on, but reads the blocks from int16 block-major planes
 * instead of the x-major image_array (the :1810-1815 gather); everything
 * from :1819 on is the same code as in the full decode. */
int gidref_reconstruct(const gidref_frame *f, const int16_t *const planes[3], uint8_t *rgb,
                       size_t stride) {
  if (!f || !planes || !rgb || (f->ncomp != 1 && f->ncomp != 3)) return GIDREF_BAD_ARGUMENT;
  if (f->width <= 0 || f->height <= 0 || stride < 3u * (size_t)f->width) return GIDREF_BAD_ARGUMENT;
  Image *im = (Image *)malloc(sizeof(Image));
  Load_State *L = (Load_State *)calloc(1, sizeof(Load_State));
  if (!im || !L) { free(im); free(L); return GIDREF_BAD_ARGUMENT; }
  image_init(im, NULL, 0);
  im->width = f->width;
  im->height = f->height;
  im->subformat_id = f->ncomp;
  im->color_space = f->ncomp == 3 ? CS_YCbCr : CS_Y_Grey;
  for (int c = 0; c < f->ncomp; c++) {
    if (f->samp_h[c] <= 0 || f->samp_v[c] <= 0 || !planes[c]) { free(im); free(L); return GIDREF_BAD_ARGUMENT; }
    im->compo_set[c] = 1;
    im->info[c].ups.samples_hor = f->samp_h[c];
    im->info[c].ups.samples_ver = f->samp_v[c];
    im->info[c].qt_assoc = c;
    if (f->samp_h[c] > im->max_samples_hor) im->max_samples_hor = f->samp_h[c];
    if (f->samp_v[c] > im->max_samples_ver) im->max_samples_ver = f->samp_v[c];
    /* store the natural-order table back in stream order so that the
     * qt_zz derivation of :1771-1775 reproduces it */
    for (int i = 0; i < 64; i++) im->qt_list[c][undo_zig_zag[i]] = f->qt[c][i];
  }
  for (int c = 0; c < f->ncomp; c++) {
    im->info[c].ups.up_factor_x = im->max_samples_hor / im->info[c].ups.samples_hor;
    im->info[c].ups.up_factor_y = im->max_samples_ver / im->info[c].ups.samples_ver;
    im->info[c].upsampling_profile = identify_profile(&im->info[c].ups);
  }
  L->im = im;
  L->bm_rgb = rgb;
  L->bm_stride = stride;
  int rc = GIDREF_OK;
  if (setjmp(im->jb) == 0) {
    load_setup_bm(L);
    int mcu_w = L->sample_shape_max_x, mcu_h = L->sample_shape_max_y;
    L->mcu_count_image_h = (im->width + mcu_w - 1) / mcu_w;
    L->mcu_count_image_v = (im->height + mcu_h - 1) / mcu_h;
    for (int c = 0; c < f->ncomp; c++) {
      L->planes[c] = (int16_t *)planes[c]; /* read-only use */
      L->plane_bx[c] = f->blocks_x[c];
      L->plane_by[c] = f->blocks_y[c];
    }
    finalize_from_planes_bm(L);
  } else {
    rc = im->err_code;
  }
  for (int c = 0; c < 3; c++) L->planes[c] = NULL; /* not ours */
  load_state_dispose(L, 0);
  image_dispose(im);
  free(L);
  free(im);
  return rc;
}

void gidref_free(void *p) { free(p); }
const char *gidref_version(void) { return GIDREF_VERSION; }
