/*
 * synth_jpeg.c -- TEST HARNESS (not product code, not the oracle).
 *
 * A small JPEG encoder producing the synthetic inputs BASELINE.json names
 * (no image corpus): baseline (SOF0) and progressive (SOF2), 8-bit, grey or
 * YCbCr with arbitrary integer sampling factors, optional restart interval
 * (baseline only), 8- or 16-bit DQT.  It returns the .jpg byte stream AND the
 * quantised coefficient planes it encoded (int16, block-major
 * [block_row][block_col][64], natural order, MCU-padded), so that tests can
 * check  front_end(jpg) == planes  and feed the same planes to the GPU.
 *
 * The streams honour the constraints GID's decoder imposes:
 * one interleaved scan for baseline; progressive = interleaved DC scans and
 * single-component AC scans, successive approximation one bit at a time, no
 * restart markers, EOB runs flushed before 937 pending correction bits; all
 * DQT before the first SOS; component ids 1, 2, 3.
 *
 * Pixel source: per channel a sum of three low-frequency 2-D
 * sinusoids with seeded phases plus uniform noise in [-8, 8],
 * seed = 0x61D0000 + i through splitmix64.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct synth_scan {
  int32_t ncomp;      /* components in the scan */
  int32_t comp[3];    /* component indices 0..2 */
  int32_t ss, se, ah, al;
} synth_scan;

typedef struct synth_params {
  int32_t width, height;
  int32_t ncomp;             /* 1 or 3 */
  int32_t samp_h[3], samp_v[3];
  int32_t quality;           /* 1..100, libjpeg scaling of the Annex K tables */
  int32_t progressive;       /* 0 = SOF0 baseline, 1 = SOF2 */
  int32_t restart_interval;  /* MCUs; baseline only */
  int32_t optimize_huffman;  /* baseline: 0 = Annex K tables, 1 = per-image optimal */
  int32_t dqt16;             /* write 16-bit DQT entries */
  int32_t nscans;            /* progressive: 0 = default script */
  synth_scan scans[32];
  uint16_t qt_override[2][64]; /* natural order; used when use_qt_override != 0 */
  int32_t use_qt_override;
} synth_params;

typedef struct synth_result {
  uint8_t *jpg;
  size_t jpg_len;
  int16_t *planes[3];
  int32_t blocks_x[3], blocks_y[3];
  uint16_t qt[3][64]; /* natural order, per component */
} synth_result;

/* ------------------------------------------------------------------------ */

static const uint8_t zigzag_nat[64] = { /* zig-zag index -> natural index */
   0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5,
  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
  58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static const uint8_t std_lum_q[64] = { /* Annex K.1, natural order */
  16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
  14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
  18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
  49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t std_chr_q[64] = { /* Annex K.2, natural order */
  17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
  24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
  99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* Annex K.3 typical Huffman tables */
static const uint8_t bits_dc_lum[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t val_dc[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t bits_dc_chr[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t bits_ac_lum[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t val_ac_lum[162] = {
  0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
  0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
  0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
  0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
  0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
  0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
  0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
  0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
  0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
  0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
  0xf9, 0xfa};
static const uint8_t bits_ac_chr[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t val_ac_chr[162] = {
  0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
  0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
  0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
  0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
  0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
  0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
  0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
  0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
  0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
  0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
  0xf9, 0xfa};

/* ---- output byte buffer -------------------------------------------------- */

typedef struct {
  uint8_t *p;
  size_t n, cap;
} Buf;

static void buf_put(Buf *b, uint8_t v) {
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : (1u << 16);
    b->p = (uint8_t *)realloc(b->p, b->cap);
    if (!b->p) abort();
  }
  b->p[b->n++] = v;
}
static void buf_put16(Buf *b, unsigned v) { buf_put(b, (uint8_t)(v >> 8)); buf_put(b, (uint8_t)v); }

/* ---- Huffman tables ------------------------------------------------------ */

typedef struct {
  uint8_t bits[17];
  uint8_t vals[256];
  int nvals;
  uint16_t ehufco[256];
  uint8_t ehufsi[256];
} HuffTab;

static void huff_derive(HuffTab *h) {
  memset(h->ehufco, 0, sizeof h->ehufco);
  memset(h->ehufsi, 0, sizeof h->ehufsi);
  unsigned code = 0;
  int k = 0;
  for (int l = 1; l <= 16; l++) {
    for (int i = 0; i < h->bits[l]; i++) {
      h->ehufco[h->vals[k]] = (uint16_t)code;
      h->ehufsi[h->vals[k]] = (uint8_t)l;
      code++;
      k++;
    }
    code <<= 1;
  }
  h->nvals = k;
}

static void huff_set(HuffTab *h, const uint8_t *bits, const uint8_t *vals, int n) {
  memcpy(h->bits, bits, 17);
  memcpy(h->vals, vals, (size_t)n);
  huff_derive(h);
}

/* Optimal table from symbol frequencies (ISO 10918-1 Annex K.2 procedure). */
static void huff_optimal(HuffTab *h, const long *freq_in) {
  long freq[257];
  int codesize[257], others[257];
  uint8_t bits[33];
  memset(bits, 0, sizeof bits);
  memset(codesize, 0, sizeof codesize);
  for (int i = 0; i < 257; i++) others[i] = -1;
  for (int i = 0; i < 256; i++) freq[i] = freq_in[i];
  freq[256] = 1; /* reserves the all-ones code */
  for (;;) {
    int c1 = -1, c2 = -1;
    long v = 1000000000L;
    for (int i = 0; i <= 256; i++)
      if (freq[i] && freq[i] <= v) { v = freq[i]; c1 = i; }
    v = 1000000000L;
    for (int i = 0; i <= 256; i++)
      if (freq[i] && freq[i] <= v && i != c1) { v = freq[i]; c2 = i; }
    if (c2 < 0) break;
    freq[c1] += freq[c2];
    freq[c2] = 0;
    codesize[c1]++;
    while (others[c1] >= 0) { c1 = others[c1]; codesize[c1]++; }
    others[c1] = c2;
    codesize[c2]++;
    while (others[c2] >= 0) { c2 = others[c2]; codesize[c2]++; }
  }
  for (int i = 0; i <= 256; i++)
    if (codesize[i]) bits[codesize[i]]++;
  for (int i = 32; i > 16; i--) {
    while (bits[i] > 0) {
      int j = i - 2;
      while (bits[j] == 0) j--;
      bits[i] -= 2;
      bits[i - 1]++;
      bits[j + 1] += 2;
      bits[j]--;
    }
  }
  int i = 16;
  while (bits[i] == 0) i--;
  bits[i]--; /* remove the reserved symbol */
  memcpy(h->bits, bits, 17);
  h->bits[0] = 0;
  int p = 0;
  for (int l = 1; l <= 32; l++)
    for (int s = 0; s < 256; s++)
      if (codesize[s] == l) h->vals[p++] = (uint8_t)s;
  huff_derive(h);
}

/* ---- entropy encoder state ------------------------------------------------ */

typedef struct {
  Buf *out;
  uint32_t acc;
  int nacc;
  int gather;          /* 1: only count symbols */
  long *freq;          /* current statistics table when gathering */
  const HuffTab *tab;  /* current table when emitting */
  /* progressive */
  unsigned eobrun;
  unsigned be;
  uint8_t corr[1000];
} Ent;

static void emit_bits(Ent *e, unsigned code, int size) {
  if (e->gather || size == 0) return;
  e->acc = (e->acc << size) | (code & ((1u << size) - 1u));
  e->nacc += size;
  while (e->nacc >= 8) {
    uint8_t c = (uint8_t)(e->acc >> (e->nacc - 8));
    buf_put(e->out, c);
    if (c == 0xFF) buf_put(e->out, 0);
    e->nacc -= 8;
  }
}
static void flush_bits(Ent *e) {
  if (e->gather) return;
  if (e->nacc > 0) emit_bits(e, 0x7F, 8 - e->nacc); /* pad with ones */
  e->acc = 0;
  e->nacc = 0;
}
static void emit_symbol(Ent *e, int sym) {
  if (e->gather) { e->freq[sym]++; return; }
  if (e->tab->ehufsi[sym] == 0) { fprintf(stderr, "synth_jpeg: symbol %02x has no code\n", sym); abort(); }
  emit_bits(e, e->tab->ehufco[sym], e->tab->ehufsi[sym]);
}
static inline int bitlen(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

/* ---- geometry ------------------------------------------------------------- */

typedef struct {
  int w, h, ncomp, hmax, vmax, mcux, mcuy;
  int sh[3], sv[3], bx[3], by[3]; /* MCU-padded block grid */
  int cbx[3], cby[3];             /* component's own block grid: ceil(ceil(W*h/hmax)/8) */
} Geo;

static void geo_init(Geo *g, const synth_params *p) {
  g->w = p->width; g->h = p->height; g->ncomp = p->ncomp;
  g->hmax = g->vmax = 1;
  for (int c = 0; c < p->ncomp; c++) {
    g->sh[c] = p->samp_h[c]; g->sv[c] = p->samp_v[c];
    if (g->sh[c] > g->hmax) g->hmax = g->sh[c];
    if (g->sv[c] > g->vmax) g->vmax = g->sv[c];
  }
  g->mcux = (g->w + 8 * g->hmax - 1) / (8 * g->hmax);
  g->mcuy = (g->h + 8 * g->vmax - 1) / (8 * g->vmax);
  for (int c = 0; c < p->ncomp; c++) {
    g->bx[c] = g->mcux * g->sh[c];
    g->by[c] = g->mcuy * g->sv[c];
    int cw = (g->w * g->sh[c] + g->hmax - 1) / g->hmax;
    int ch = (g->h * g->sv[c] + g->vmax - 1) / g->vmax;
    g->cbx[c] = (cw + 7) / 8;
    g->cby[c] = (ch + 7) / 8;
  }
}

/* ---- pixel source --------------------------------------------------------- */

static uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* rgb: packed top-down 3*W*H bytes */
void synth_image(uint64_t index, int w, int h, uint8_t *rgb) {
  uint64_t st = 0x61D0000ull + index;
  int n = w > h ? w : h;
  float *tab = (float *)malloc(sizeof(float) * (size_t)n * 4 * 9);
  float amp[9];
  for (int k = 0; k < 9; k++) {
    /* low spatial frequencies: 0.5 .. 6 periods across the frame */
    double fx = (0.5 + 5.5 * (double)(splitmix64(&st) >> 11) / 9007199254740992.0) * 2.0 * M_PI / (double)w;
    double fy = (0.5 + 5.5 * (double)(splitmix64(&st) >> 11) / 9007199254740992.0) * 2.0 * M_PI / (double)h;
    double ph = 2.0 * M_PI * (double)(splitmix64(&st) >> 11) / 9007199254740992.0;
    amp[k] = 36.0f;
    float *sx = tab + (size_t)k * 4 * n, *cx = sx + n, *sy = cx + n, *cy = sy + n;
    for (int x = 0; x < w; x++) { sx[x] = (float)sin(fx * x + ph); cx[x] = (float)cos(fx * x + ph); }
    for (int y = 0; y < h; y++) { sy[y] = (float)sin(fy * y); cy[y] = (float)cos(fy * y); }
  }
  uint64_t ns = splitmix64(&st);
  for (int y = 0; y < h; y++) {
    uint8_t *row = rgb + (size_t)3 * w * y;
    for (int x = 0; x < w; x++) {
      for (int ch = 0; ch < 3; ch++) {
        float v = 128.0f;
        for (int j = 0; j < 3; j++) {
          int k = ch * 3 + j;
          const float *sx = tab + (size_t)k * 4 * n, *cx = sx + n, *sy = cx + n, *cy = sy + n;
          v += amp[k] * (sx[x] * cy[y] + cx[x] * sy[y]); /* sin(fx x + ph + fy y) */
        }
        /* xorshift noise in [-8, 8] */
        ns ^= ns << 13; ns ^= ns >> 7; ns ^= ns << 17;
        v += (float)((int)(ns % 17u) - 8);
        int iv = (int)lrintf(v);
        row[3 * x + ch] = (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
      }
    }
  }
  free(tab);
}

/* ---- pixels -> quantised planes ------------------------------------------- */

static void make_qtables(const synth_params *p, uint16_t qt[2][64]) {
  if (p->use_qt_override) { memcpy(qt, p->qt_override, sizeof(uint16_t) * 128); return; }
  int q = p->quality < 1 ? 1 : (p->quality > 100 ? 100 : p->quality);
  int scale = q < 50 ? 5000 / q : 200 - 2 * q;
  for (int t = 0; t < 2; t++)
    for (int i = 0; i < 64; i++) {
      long v = ((long)(t ? std_chr_q[i] : std_lum_q[i]) * scale + 50) / 100;
      if (v < 1) v = 1;
      if (v > 255) v = 255;
      qt[t][i] = (uint16_t)v;
    }
}

static void fdct_quant(const float *blk, const uint16_t *q, int16_t *out) {
  static float cm[8][8];
  static int ready = 0;
  if (!ready) {
    for (int u = 0; u < 8; u++)
      for (int x = 0; x < 8; x++)
        cm[u][x] = (float)((u == 0 ? sqrt(0.125) : 0.5) * cos((2 * x + 1) * u * M_PI / 16.0));
    ready = 1;
  }
  float tmp[64];
  for (int y = 0; y < 8; y++)
    for (int u = 0; u < 8; u++) {
      float s = 0;
      for (int x = 0; x < 8; x++) s += cm[u][x] * blk[y * 8 + x];
      tmp[y * 8 + u] = s;
    }
  for (int v = 0; v < 8; v++)
    for (int u = 0; u < 8; u++) {
      float s = 0;
      for (int y = 0; y < 8; y++) s += cm[v][y] * tmp[y * 8 + u];
      float qv = s / (float)q[v * 8 + u];
      out[v * 8 + u] = (int16_t)lrintf(qv);
    }
}

static int planes_from_pixels(const synth_params *p, const Geo *g, const uint8_t *pix,
                              const uint16_t qt[2][64], int16_t *planes[3]) {
  /* component sample planes at MCU-padded size, edge replicated */
  for (int c = 0; c < g->ncomp; c++) {
    int pw = g->bx[c] * 8, ph = g->by[c] * 8;
    int ux = g->hmax / g->sh[c], uy = g->vmax / g->sv[c];
    float *sp = (float *)malloc(sizeof(float) * (size_t)pw * ph);
    if (!sp) return -1;
    for (int y = 0; y < ph; y++)
      for (int x = 0; x < pw; x++) {
        float acc = 0;
        for (int dy = 0; dy < uy; dy++)
          for (int dx = 0; dx < ux; dx++) {
            int sx = x * ux + dx, sy = y * uy + dy;
            if (sx >= g->w) sx = g->w - 1;
            if (sy >= g->h) sy = g->h - 1;
            float v;
            if (p->ncomp == 1) {
              /* grey input: one byte per pixel */
              v = pix[(size_t)sy * g->w + sx];
            } else {
              const uint8_t *q3 = pix + ((size_t)sy * g->w + sx) * 3;
              float r = q3[0], gg = q3[1], b = q3[2];
              if (c == 0) v = 0.299f * r + 0.587f * gg + 0.114f * b;
              else if (c == 1) v = -0.168736f * r - 0.331264f * gg + 0.5f * b + 128.0f;
              else v = 0.5f * r - 0.418688f * gg - 0.081312f * b + 128.0f;
            }
            acc += v;
          }
        sp[(size_t)y * pw + x] = acc / (float)(ux * uy) - 128.0f;
      }
    const uint16_t *q = qt[c == 0 ? 0 : 1];
    for (int by = 0; by < g->by[c]; by++)
      for (int bx = 0; bx < g->bx[c]; bx++) {
        float blk[64];
        for (int y = 0; y < 8; y++)
          for (int x = 0; x < 8; x++) blk[y * 8 + x] = sp[(size_t)(by * 8 + y) * pw + bx * 8 + x];
        fdct_quant(blk, q, planes[c] + ((size_t)by * g->bx[c] + bx) * 64);
      }
    free(sp);
  }
  return 0;
}

/* ---- baseline scan --------------------------------------------------------- */

typedef struct {
  HuffTab dc[2], ac[2];
  long fdc[2][256], fac[2][256];
} Tables;

static void encode_block_baseline(Ent *e, Tables *t, int tsel, const int16_t *blk, int *last_dc) {
  int diff = blk[0] - *last_dc;
  *last_dc = blk[0];
  int temp = diff, temp2 = diff;
  if (temp < 0) { temp = -temp; temp2--; }
  int nbits = bitlen((unsigned)temp);
  e->tab = &t->dc[tsel]; e->freq = t->fdc[tsel];
  emit_symbol(e, nbits);
  emit_bits(e, (unsigned)temp2, nbits);
  e->tab = &t->ac[tsel]; e->freq = t->fac[tsel];
  int r = 0;
  for (int k = 1; k < 64; k++) {
    int v = blk[zigzag_nat[k]];
    if (v == 0) { r++; continue; }
    while (r > 15) { emit_symbol(e, 0xF0); r -= 16; }
    temp = v; temp2 = v;
    if (temp < 0) { temp = -temp; temp2--; }
    nbits = bitlen((unsigned)temp);
    emit_symbol(e, (r << 4) + nbits);
    emit_bits(e, (unsigned)temp2, nbits);
    r = 0;
  }
  if (r > 0) emit_symbol(e, 0);
}

static void scan_baseline(Ent *e, Tables *t, const synth_params *p, const Geo *g,
                          int16_t *const planes[3]) {
  int last_dc[3] = {0, 0, 0};
  int rst = 0, count = 0;
  for (int my = 0; my < g->mcuy; my++)
    for (int mx = 0; mx < g->mcux; mx++) {
      if (p->restart_interval > 0 && count == p->restart_interval) {
        flush_bits(e);
        if (!e->gather) { buf_put(e->out, 0xFF); buf_put(e->out, (uint8_t)(0xD0 + rst)); }
        rst = (rst + 1) & 7;
        count = 0;
        last_dc[0] = last_dc[1] = last_dc[2] = 0;
      }
      for (int c = 0; c < g->ncomp; c++)
        for (int sy = 0; sy < g->sv[c]; sy++)
          for (int sx = 0; sx < g->sh[c]; sx++) {
            int bx = mx * g->sh[c] + sx, by = my * g->sv[c] + sy;
            encode_block_baseline(e, t, c == 0 ? 0 : 1,
                                  planes[c] + ((size_t)by * g->bx[c] + bx) * 64, &last_dc[c]);
          }
      count++;
    }
  flush_bits(e);
}

/* ---- progressive scans ----------------------------------------------------- */

static void emit_buffered_bits(Ent *e, const uint8_t *bits, unsigned n) {
  if (e->gather) return;
  for (unsigned i = 0; i < n; i++) emit_bits(e, bits[i], 1);
}

static void emit_eobrun(Ent *e) {
  if (e->eobrun > 0) {
    int nbits = bitlen(e->eobrun) - 1;
    emit_symbol(e, nbits << 4);
    if (nbits) emit_bits(e, e->eobrun, nbits);
    e->eobrun = 0;
    emit_buffered_bits(e, e->corr, e->be);
    e->be = 0;
  }
}

static void scan_dc(Ent *e, Tables *t, const Geo *g, const synth_scan *s, int16_t *const planes[3]) {
  int last_dc[3] = {0, 0, 0};
  if (s->ncomp > 1) {
    for (int my = 0; my < g->mcuy; my++)
      for (int mx = 0; mx < g->mcux; mx++)
        for (int ci = 0; ci < s->ncomp; ci++) {
          int c = s->comp[ci];
          for (int sy = 0; sy < g->sv[c]; sy++)
            for (int sx = 0; sx < g->sh[c]; sx++) {
              int bx = mx * g->sh[c] + sx, by = my * g->sv[c] + sy;
              int coef = planes[c][((size_t)by * g->bx[c] + bx) * 64];
              if (s->ah == 0) {
                int v = coef >> s->al; /* point transform: arithmetic shift */
                int diff = v - last_dc[c];
                last_dc[c] = v;
                int temp = diff, temp2 = diff;
                if (temp < 0) { temp = -temp; temp2--; }
                int nbits = bitlen((unsigned)temp);
                e->tab = &t->dc[c == 0 ? 0 : 1]; e->freq = t->fdc[c == 0 ? 0 : 1];
                emit_symbol(e, nbits);
                emit_bits(e, (unsigned)temp2, nbits);
              } else {
                emit_bits(e, (unsigned)(coef >> s->al) & 1u, 1);
              }
            }
        }
  } else {
    /* single-component DC scan: the component's own block grid, raster order */
    int c = s->comp[0];
    for (int by = 0; by < g->cby[c]; by++)
      for (int bx = 0; bx < g->cbx[c]; bx++) {
        int coef = planes[c][((size_t)by * g->bx[c] + bx) * 64];
        if (s->ah == 0) {
          int v = coef >> s->al;
          int diff = v - last_dc[c];
          last_dc[c] = v;
          int temp = diff, temp2 = diff;
          if (temp < 0) { temp = -temp; temp2--; }
          int nbits = bitlen((unsigned)temp);
          e->tab = &t->dc[c == 0 ? 0 : 1]; e->freq = t->fdc[c == 0 ? 0 : 1];
          emit_symbol(e, nbits);
          emit_bits(e, (unsigned)temp2, nbits);
        } else {
          emit_bits(e, (unsigned)(coef >> s->al) & 1u, 1);
        }
      }
  }
  flush_bits(e);
}

static void scan_ac_first(Ent *e, Tables *t, const Geo *g, const synth_scan *s,
                          int16_t *const planes[3]) {
  int c = s->comp[0];
  e->tab = &t->ac[c == 0 ? 0 : 1]; e->freq = t->fac[c == 0 ? 0 : 1];
  e->eobrun = 0; e->be = 0;
  for (int by = 0; by < g->cby[c]; by++)
    for (int bx = 0; bx < g->cbx[c]; bx++) {
      const int16_t *blk = planes[c] + ((size_t)by * g->bx[c] + bx) * 64;
      int r = 0;
      for (int k = s->ss; k <= s->se; k++) {
        int temp = blk[zigzag_nat[k]], temp2;
        if (temp == 0) { r++; continue; }
        if (temp < 0) { temp = -temp; temp >>= s->al; temp2 = ~temp; }
        else { temp >>= s->al; temp2 = temp; }
        if (temp == 0) { r++; continue; }
        if (e->eobrun > 0) emit_eobrun(e);
        while (r > 15) { emit_symbol(e, 0xF0); r -= 16; }
        int nbits = bitlen((unsigned)temp);
        emit_symbol(e, (r << 4) + nbits);
        emit_bits(e, (unsigned)temp2, nbits);
        r = 0;
      }
      if (r > 0) {
        e->eobrun++;
        if (e->eobrun == 0x7FFF) emit_eobrun(e);
      }
    }
  emit_eobrun(e);
  flush_bits(e);
}

static void scan_ac_refine(Ent *e, Tables *t, const Geo *g, const synth_scan *s,
                           int16_t *const planes[3]) {
  int c = s->comp[0];
  e->tab = &t->ac[c == 0 ? 0 : 1]; e->freq = t->fac[c == 0 ? 0 : 1];
  e->eobrun = 0; e->be = 0;
  for (int by = 0; by < g->cby[c]; by++)
    for (int bx = 0; bx < g->cbx[c]; bx++) {
      const int16_t *blk = planes[c] + ((size_t)by * g->bx[c] + bx) * 64;
      int absval[64];
      int eob = 0;
      for (int k = s->ss; k <= s->se; k++) {
        int temp = blk[zigzag_nat[k]];
        if (temp < 0) temp = -temp;
        temp >>= s->al;
        absval[k] = temp;
        if (temp == 1) eob = k;
      }
      int r = 0;
      unsigned br = 0;
      uint8_t *br_buf = e->corr + e->be;
      for (int k = s->ss; k <= s->se; k++) {
        int temp = absval[k];
        if (temp == 0) { r++; continue; }
        while (r > 15 && k <= eob) {
          emit_eobrun(e);
          emit_symbol(e, 0xF0);
          r -= 16;
          emit_buffered_bits(e, br_buf, br);
          br_buf = e->corr;
          br = 0;
        }
        if (temp > 1) { br_buf[br++] = (uint8_t)(temp & 1); continue; }
        emit_eobrun(e);
        emit_symbol(e, (r << 4) + 1);
        emit_bits(e, blk[zigzag_nat[k]] < 0 ? 0u : 1u, 1);
        emit_buffered_bits(e, br_buf, br);
        br_buf = e->corr;
        br = 0;
        r = 0;
      }
      if (r > 0 || br > 0) {
        e->eobrun++;
        e->be += br;
        if (e->eobrun == 0x7FFF || e->be > (1000 - 64 + 1)) emit_eobrun(e);
      }
    }
  emit_eobrun(e);
  flush_bits(e);
}

/* ---- markers ---------------------------------------------------------------- */

static void write_dqt(Buf *b, int id, const uint16_t *q_nat, int prec16) {
  buf_put(b, 0xFF); buf_put(b, 0xDB);
  buf_put16(b, (unsigned)(2 + 1 + (prec16 ? 128 : 64)));
  buf_put(b, (uint8_t)((prec16 ? 0x10 : 0) | id));
  for (int k = 0; k < 64; k++) {
    unsigned v = q_nat[zigzag_nat[k]];
    if (prec16) buf_put16(b, v); else buf_put(b, (uint8_t)v);
  }
}
static void write_dht(Buf *b, int cls, int id, const HuffTab *h) {
  buf_put(b, 0xFF); buf_put(b, 0xC4);
  buf_put16(b, (unsigned)(2 + 1 + 16 + h->nvals));
  buf_put(b, (uint8_t)((cls << 4) | id));
  for (int l = 1; l <= 16; l++) buf_put(b, h->bits[l]);
  for (int i = 0; i < h->nvals; i++) buf_put(b, h->vals[i]);
}
static void write_sos(Buf *b, const synth_scan *s) {
  buf_put(b, 0xFF); buf_put(b, 0xDA);
  buf_put16(b, (unsigned)(6 + 2 * s->ncomp));
  buf_put(b, (uint8_t)s->ncomp);
  for (int i = 0; i < s->ncomp; i++) {
    int c = s->comp[i];
    buf_put(b, (uint8_t)(c + 1));
    buf_put(b, (uint8_t)(c == 0 ? 0x00 : 0x11));
  }
  buf_put(b, (uint8_t)s->ss); buf_put(b, (uint8_t)s->se);
  buf_put(b, (uint8_t)((s->ah << 4) | s->al));
}

static int default_script(const synth_params *p, synth_scan *sc) {
  /* DC (Al=1) interleaved, DC refine, per-component AC 1-5,
   * AC 6-63 (Al=1), AC 6-63 refine. */
  int n = 0;
  synth_scan dc = {p->ncomp, {0, 1, 2}, 0, 0, 0, 1};
  sc[n++] = dc;
  for (int c = 0; c < p->ncomp; c++) { synth_scan a = {1, {c, 0, 0}, 1, 5, 0, 0}; sc[n++] = a; }
  synth_scan dcr = {p->ncomp, {0, 1, 2}, 0, 0, 1, 0};
  sc[n++] = dcr;
  for (int c = 0; c < p->ncomp; c++) { synth_scan a = {1, {c, 0, 0}, 6, 63, 0, 1}; sc[n++] = a; }
  for (int c = 0; c < p->ncomp; c++) { synth_scan a = {1, {c, 0, 0}, 6, 63, 1, 0}; sc[n++] = a; }
  return n;
}

/* ---- planes -> jpg ------------------------------------------------------------ */

static void run_scan(Ent *e, Tables *t, const synth_params *p, const Geo *g, const synth_scan *s,
                     int16_t *const planes[3]) {
  if (!p->progressive) scan_baseline(e, t, p, g, planes);
  else if (s->ss == 0) scan_dc(e, t, g, s, planes);
  else if (s->ah == 0) scan_ac_first(e, t, g, s, planes);
  else scan_ac_refine(e, t, g, s, planes);
}

int synth_encode_planes(const synth_params *p, int16_t *const planes_in[3], synth_result *res) {
  Geo g;
  geo_init(&g, p);
  uint16_t qt[2][64];
  make_qtables(p, qt);
  Buf out = {0, 0, 0};
  /* progressive: blocks outside a component's own grid never get AC data */
  int16_t *planes[3] = {0, 0, 0};
  for (int c = 0; c < g.ncomp; c++) {
    size_t n = (size_t)g.bx[c] * g.by[c] * 64;
    planes[c] = (int16_t *)malloc(n * sizeof(int16_t));
    if (!planes[c]) return -1;
    memcpy(planes[c], planes_in[c], n * sizeof(int16_t));
    if (p->progressive)
      for (int by = 0; by < g.by[c]; by++)
        for (int bx = 0; bx < g.bx[c]; bx++)
          if (bx >= g.cbx[c] || by >= g.cby[c])
            memset(planes[c] + ((size_t)by * g.bx[c] + bx) * 64 + 1, 0, 63 * sizeof(int16_t));
  }

  buf_put(&out, 0xFF); buf_put(&out, 0xD8);
  /* JFIF APP0 */
  static const uint8_t app0[] = {0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  for (size_t i = 0; i < sizeof app0; i++) buf_put(&out, app0[i]);
  write_dqt(&out, 0, qt[0], p->dqt16);
  if (g.ncomp == 3) write_dqt(&out, 1, qt[1], p->dqt16);
  /* SOF */
  buf_put(&out, 0xFF); buf_put(&out, p->progressive ? 0xC2 : 0xC0);
  buf_put16(&out, (unsigned)(8 + 3 * g.ncomp));
  buf_put(&out, 8);
  buf_put16(&out, (unsigned)g.h); buf_put16(&out, (unsigned)g.w);
  buf_put(&out, (uint8_t)g.ncomp);
  for (int c = 0; c < g.ncomp; c++) {
    buf_put(&out, (uint8_t)(c + 1));
    buf_put(&out, (uint8_t)((g.sh[c] << 4) | g.sv[c]));
    buf_put(&out, (uint8_t)(c == 0 ? 0 : 1));
  }
  if (p->restart_interval > 0 && !p->progressive) {
    buf_put(&out, 0xFF); buf_put(&out, 0xDD); buf_put16(&out, 4);
    buf_put16(&out, (unsigned)p->restart_interval);
  }

  Tables *t = (Tables *)calloc(1, sizeof(Tables));
  Ent e;
  memset(&e, 0, sizeof e);
  e.out = &out;
  synth_scan scans[32];
  int nscans;
  if (!p->progressive) {
    synth_scan s = {g.ncomp, {0, 1, 2}, 0, 63, 0, 0};
    scans[0] = s;
    nscans = 1;
  } else if (p->nscans > 0) {
    nscans = p->nscans;
    memcpy(scans, p->scans, sizeof(synth_scan) * (size_t)nscans);
  } else {
    nscans = default_script(p, scans);
  }
  for (int si = 0; si < nscans; si++) {
    const synth_scan *s = &scans[si];
    int optimise = p->progressive || p->optimize_huffman;
    if (optimise) {
      memset(t->fdc, 0, sizeof t->fdc);
      memset(t->fac, 0, sizeof t->fac);
      e.gather = 1;
      run_scan(&e, t, p, &g, s, planes);
      e.gather = 0;
      for (int k = 0; k < 2; k++) {
        long sdc = 0, sac = 0;
        for (int i = 0; i < 256; i++) { sdc += t->fdc[k][i]; sac += t->fac[k][i]; }
        if (sdc) { huff_optimal(&t->dc[k], t->fdc[k]); write_dht(&out, 0, k, &t->dc[k]); }
        if (sac) { huff_optimal(&t->ac[k], t->fac[k]); write_dht(&out, 1, k, &t->ac[k]); }
      }
    } else {
      huff_set(&t->dc[0], bits_dc_lum, val_dc, 12);
      huff_set(&t->ac[0], bits_ac_lum, val_ac_lum, 162);
      write_dht(&out, 0, 0, &t->dc[0]);
      write_dht(&out, 1, 0, &t->ac[0]);
      if (g.ncomp == 3) {
        huff_set(&t->dc[1], bits_dc_chr, val_dc, 12);
        huff_set(&t->ac[1], bits_ac_chr, val_ac_chr, 162);
        write_dht(&out, 0, 1, &t->dc[1]);
        write_dht(&out, 1, 1, &t->ac[1]);
      }
    }
    write_sos(&out, s);
    e.acc = 0; e.nacc = 0;
    run_scan(&e, t, p, &g, s, planes);
  }
  buf_put(&out, 0xFF); buf_put(&out, 0xD9);
  free(t);

  memset(res, 0, sizeof *res);
  res->jpg = out.p;
  res->jpg_len = out.n;
  for (int c = 0; c < g.ncomp; c++) {
    res->planes[c] = planes[c];
    res->blocks_x[c] = g.bx[c];
    res->blocks_y[c] = g.by[c];
    memcpy(res->qt[c], qt[c == 0 ? 0 : 1], sizeof(uint16_t) * 64);
  }
  return 0;
}

/* pixels: packed RGB (ncomp = 3) or one byte per pixel (ncomp = 1) */
int synth_encode_pixels(const synth_params *p, const uint8_t *pixels, synth_result *res) {
  Geo g;
  geo_init(&g, p);
  uint16_t qt[2][64];
  make_qtables(p, qt);
  int16_t *planes[3] = {0, 0, 0};
  for (int c = 0; c < g.ncomp; c++) {
    planes[c] = (int16_t *)calloc((size_t)g.bx[c] * g.by[c] * 64, sizeof(int16_t));
    if (!planes[c]) return -1;
  }
  int rc = planes_from_pixels(p, &g, pixels, qt, planes);
  if (rc == 0) rc = synth_encode_planes(p, planes, res);
  for (int c = 0; c < 3; c++) free(planes[c]);
  return rc;
}

/* quantised planes only (no entropy coding): the bench uses this to fill
 * device-resident inputs quickly */
int synth_planes_only(const synth_params *p, const uint8_t *pixels, synth_result *res) {
  Geo g;
  geo_init(&g, p);
  uint16_t qt[2][64];
  make_qtables(p, qt);
  memset(res, 0, sizeof *res);
  for (int c = 0; c < g.ncomp; c++) {
    res->planes[c] = (int16_t *)calloc((size_t)g.bx[c] * g.by[c] * 64, sizeof(int16_t));
    if (!res->planes[c]) return -1;
    res->blocks_x[c] = g.bx[c];
    res->blocks_y[c] = g.by[c];
    memcpy(res->qt[c], qt[c == 0 ? 0 : 1], sizeof(uint16_t) * 64);
  }
  return planes_from_pixels(p, &g, pixels, qt, res->planes);
}

void synth_result_free(synth_result *res) {
  free(res->jpg);
  for (int c = 0; c < 3; c++) free(res->planes[c]);
  memset(res, 0, sizeof *res);
}
