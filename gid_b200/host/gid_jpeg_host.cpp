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
#include "gid.hpp"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/gidcuda.h"

namespace gid {

namespace {

std::string fmt(const char *f, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

// zig-zag index -> natural index (gid-decoding_jpg.adb:570-578)
const uint8_t kZigZag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

enum Marker : uint8_t {
  M_SOF0 = 0xC0, M_SOF2 = 0xC2, M_DHT = 0xC4, M_RST0 = 0xD0, M_RST7 = 0xD7, M_SOI = 0xD8, M_EOI = 0xD9,
  M_SOS = 0xDA, M_DQT = 0xDB, M_DRI = 0xDD, M_APP1 = 0xE1, M_COM = 0xFE
};

// Markers GID knows (marker_id table, gid-decoding_jpg.ads:41-87): SOF0-15
// (C0..CF incl. DHT C4, JPG C8, DAC CC), RST0-7, SOI, EOI, SOS, DQT, DNL, DRI,
// DHP, EXP, APP0-15, JPG0-13, COM.  Anything else raises "unknown marker".
bool known_marker(uint8_t b) { return b >= 0xC0 && b <= 0xFE; }

struct Reader {
  Stream &s;
  explicit Reader(Stream &st) : s(st) {}
  uint8_t byte() {  // GID.Buffering.Get_Byte: End_Error at EOF (gid-buffering.adb:95-105)
    if (s.pos >= s.size) throw error_in_image_data("JPEG: unexpected end of data (End_Error)");
    return s.data[s.pos++];
  }
  unsigned be16() {
    const unsigned h = byte();
    return (h << 8) | byte();
  }
  void skip(int n) {
    for (int i = 0; i < n; i++) (void)byte();
  }
};

struct Segment {
  uint8_t marker;
  int length;  // of the contents, without the marker and the length field
};

Segment read_segment(Reader &r, bool have_marker, uint8_t buffered) {  // :122-167
  uint8_t b = buffered;
  if (!have_marker) {
    if (r.byte() != 0xFF) throw error_in_image_data("JPEG: expected marker prefix (16#FF#) here");
    b = r.byte();
  }
  if (!known_marker(b)) throw error_in_image_data(fmt("JPEG: unknown marker here: 16#FF#, then 16#%02X#", b));
  Segment sg{b, 0};
  if (b != M_EOI) sg.length = (int)r.be16() - 2;
  return sg;
}

void read_dht(Image_Descriptor &im, Reader &r, int length) {  // :320-391
  int remaining = length;
  while (remaining > 0) {
    const uint8_t b = r.byte();
    remaining--;
    const int kind = b >= 8 ? 1 : 0;  // GID: ">= 8" means AC
    const int idx = b & 7;
    std::vector<uint8_t> counts(16);
    int total = 0;
    for (int i = 0; i < 16; i++) {
      counts[i] = r.byte();
      remaining--;
      total += counts[i];
    }
    // the code space must not be over-subscribed ("DHT data too short [2]")
    long space = 65536, spread = 65536;
    for (int len = 0; len < 16; len++) {
      spread /= 2;
      if (counts[len] > 0) {
        if (remaining < counts[len]) throw error_in_image_data("JPEG: DHT data too short [1]");
        space -= (long)counts[len] * spread;
        if (space < 0) throw error_in_image_data("JPEG: DHT data too short [2]");
      }
    }
    std::vector<uint8_t> vals((size_t)total);
    for (int i = 0; i < total; i++) vals[(size_t)i] = r.byte();
    remaining -= total;
    im.huff_bits[kind][idx] = counts;
    im.huff_vals[kind][idx] = vals;
  }
}

void read_dqt(Image_Descriptor &im, Reader &r, int length) {  // :393-422
  int remaining = length;
  while (remaining > 0) {
    const uint8_t b = r.byte();
    remaining--;
    const bool high = b >= 8;
    Quantization_Table &t = im.qt_list[b & 7];
    for (int i = 0; i < 64; i++) {
      if (high) { t.q[i] = (uint16_t)r.be16(); remaining -= 2; }
      else { t.q[i] = r.byte(); remaining -= 1; }
    }
    t.defined = true;
  }
}

void read_exif(Image_Descriptor &im, Reader &r, int length) {  // :436-543
  if (length < 6) { r.skip(length); return; }
  char sig[6];
  for (char &c : sig) c = (char)r.byte();
  if (memcmp(sig, "Exif\0\0", 6) != 0) { r.skip(length - 6); return; }
  const char endian = (char)r.byte();
  r.skip(7);
  (void)r.byte();  // IFD0 entry count (unused by the scan below, as in the reference)
  (void)r.byte();
  int x = 17;
  while (x <= length - 12) {
    const unsigned b0 = r.byte(), b1 = r.byte();
    const unsigned tag = endian == 'I' ? (b0 | (b1 << 8)) : (b1 | (b0 << 8));
    r.skip(6);
    uint8_t value;
    if (endian == 'I') { value = r.byte(); r.skip(3); }
    else { (void)r.byte(); value = r.byte(); r.skip(2); }
    x += 12;
    if (tag == 0x112) {
      switch (value) {
        case 8: im.display_orientation = Orientation::Rotation_90; break;
        case 3: im.display_orientation = Orientation::Rotation_180; break;
        case 6: im.display_orientation = Orientation::Rotation_270; break;
        default: im.display_orientation = Orientation::Unchanged; break;
      }
      break;
    }
  }
  for (int i = x; i <= length; i++) (void)r.byte();
}

void read_sof(Image_Descriptor &im, Reader &r, const Segment &sg) {  // :184-318
  if (sg.marker == M_SOF0) im.detailed_format = "JPEG, Baseline DCT (SOF_0)";
  else if (sg.marker == M_SOF2) { im.detailed_format = "JPEG, Progressive DCT (SOF_2)"; im.progressive = true; }
  else throw unsupported_image_subformat(fmt("JPEG: image type not yet supported: SOF_%d", sg.marker - 0xC0));
  const int bits = r.byte();
  if (bits != 8) throw unsupported_image_subformat(fmt("JPEG: bits per primary color= %d (not supported)", bits));
  im.bits_per_pixel = 3 * bits;
  const int h = (int)r.be16(), w = (int)r.be16();
  if (w == 0) throw error_in_image_data("JPEG: zero image width");
  if (h == 0) throw error_in_image_data("JPEG: zero image height");
  im.width = w;
  im.height = h;
  im.subformat_id = r.byte();
  im.max_samples_hor = im.max_samples_ver = 0;
  int id_base = 1;
  bool seen[5] = {false, false, false, false, false};  // Y, Cb, Cr, I, Q
  for (int i = 0; i < im.subformat_id; i++) {
    int b = r.byte();
    if (b == 0) id_base = 0;  // encoders that number the components 0, 1, 2 (:226-233)
    if (b - id_base > 4 || b - id_base < 0) throw error_in_image_data(fmt("JPEG: SOF: invalid component ID: %d", b));
    const int c = b - id_base;
    seen[c] = true;
    const int sampling = r.byte(), tq = r.byte();
    if (c < 3) {
      Component_Info &ci = im.info[c];
      ci.present = true;
      ci.samples_hor = sampling / 16;
      ci.samples_ver = sampling % 16;
      ci.qt_assoc = tq;
      if (ci.samples_hor > im.max_samples_hor) im.max_samples_hor = ci.samples_hor;
      if (ci.samples_ver > im.max_samples_ver) im.max_samples_ver = ci.samples_ver;
    }
  }
  if (sg.length < 6 + 3 * im.subformat_id) throw error_in_image_data("JPEG: SOF segment too short");
  const bool ycbcr = seen[0] && seen[1] && seen[2] && !seen[3] && !seen[4];
  const bool grey = seen[0] && !seen[1] && !seen[2] && !seen[3] && !seen[4];
  if (!ycbcr && !grey) {
    if (seen[0] && seen[1] && seen[2] && seen[3] && !seen[4])
      throw unsupported_image_subformat("JPEG: CMYK color space is currently not properly decoded");
    throw unsupported_image_subformat("JPEG: only YCbCr, Y_Grey and CMYK color spaces are currently defined");
  }
  im.greyscale = grey;
  im.detailed_format += grey ? ", Y_GREY" : ", YCBCR";
  for (Component_Info &ci : im.info)
    if (ci.present) {
      if (ci.samples_hor == 0 || ci.samples_ver == 0)
        throw constraint_error("JPEG: zero sampling factor (division by zero in Read_SOF)");
      ci.up_factor_x = im.max_samples_hor / ci.samples_hor;
      ci.up_factor_y = im.max_samples_ver / ci.samples_ver;
    }
}

// ---- entropy decoding --------------------------------------------------------------

// Canonical Huffman table with an 8-bit fast path (the reference uses one 65 536
// entry direct table per Huffman table, gid.ads:316; the decoded symbols are the same).
struct HuffTable {
  bool defined = false;
  uint8_t fast_len[256], fast_sym[256];
  int maxcode[18], valptr[17], mincode[17];
  std::vector<uint8_t> vals;
  void build(const std::vector<uint8_t> &counts, const std::vector<uint8_t> &v) {
    defined = !counts.empty();
    if (!defined) return;
    vals = v;
    memset(fast_len, 0, sizeof fast_len);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
      valptr[len] = k;
      mincode[len] = code;
      for (int i = 0; i < counts[(size_t)len - 1]; i++, k++, code++) {
        if (len <= 8) {
          const int first = code << (8 - len), n = 1 << (8 - len);
          for (int j = 0; j < n; j++) { fast_len[first + j] = (uint8_t)len; fast_sym[first + j] = vals[(size_t)k]; }
        }
      }
      maxcode[len] = counts[(size_t)len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7FFFFFFF;
  }
};

struct BitReader {  // :594-690
  Reader &r;
  uint32_t buf = 0;
  int len = 0;
  uint8_t memo_marker = 0;  // a marker met while filling (it WILL appear at the end of the scan)
  explicit BitReader(Reader &rd) : r(rd) {}
  void reset() { buf = 0; len = 0; memo_marker = 0; }
  void fill(int bits) {
    while (len < bits) {
      uint8_t b;
      if (memo_marker != 0 || r.s.pos >= r.s.size) {
        b = 0xFF;  // past a marker / end of data: pad (the reference pads with FF at End_Error, :665-670)
      } else {
        b = r.s.data[r.s.pos++];
        if (b == 0xFF) {
          const uint8_t m = r.s.pos < r.s.size ? r.s.data[r.s.pos++] : 0xD9;
          if (m != 0) {  // byte stuffing FF 00 -> FF; anything else is a marker
            const bool ok = m == M_EOI || m == M_COM || m == M_DHT || m == M_DRI || m == M_SOS ||
                            (m >= M_RST0 && m <= M_RST7);
            if (!ok) throw error_in_image_data(fmt("JPEG: Invalid marker within filling of bit buffer: 16#%02X#", m));
            memo_marker = m;
          }
        }
      }
      buf = (buf << 8) | b;
      len += 8;
    }
  }
  int show(int bits) {
    if (bits == 0) return 0;
    fill(bits);
    return (int)((buf >> (len - bits)) & ((1u << bits) - 1u));
  }
  void skip(int bits) {
    if (len < bits) fill(bits);
    len -= bits;
  }
  int get(int bits) {
    const int v = show(bits);
    skip(bits);
    return v;
  }
  int symbol(const HuffTable &t) {  // Next_Huffval, :738-746
    const int look = show(16);
    int l = t.fast_len[look >> 8];
    if (l != 0) {
      skip(l);
      return t.fast_sym[look >> 8];
    }
    for (l = 9; l <= 16; l++) {
      const int code = look >> (16 - l);
      if (code <= t.maxcode[l]) {
        skip(l);
        return t.vals[(size_t)(t.valptr[l] + code - t.mincode[l])];
      }
    }
    throw error_in_image_data("JPEG: VLC table: bits = 0");
  }
  static int extend(int v, int bits) {  // Bin_Twos_Complement, :748-753
    return v < (1 << (bits - 1)) ? v + 1 - (1 << bits) : v;
  }
  int receive_extend(int bits) { return bits ? extend(get(bits), bits) : 0; }
};

struct PlaneRef {
  int16_t *base = nullptr;
  int blocks_x = 0, blocks_y = 0;
  int16_t *block(int bx, int by) const { return base + ((size_t)by * (size_t)blocks_x + (size_t)bx) * 64; }
};

// Sparse output of the entropy decoder for one component (include/gidcuda.h, "sparse
// coefficient input"): first[seq] = index of the first record of block seq (blocks in the
// order the interleaved scan visits them), record = (uint16)value << 16 | natural index.
struct RecordSink {
  uint32_t *first = nullptr, *rec = nullptr;
  size_t cap = 0, n = 0;
  uint32_t seq = 0, nblocks = 0;
  bool on = false;
  void put(int nat, int value) { rec[n++] = ((uint32_t)(uint16_t)(int16_t)value << 16) | (uint32_t)nat; }
};

struct Decoder {
  Image_Descriptor &im;
  Reader r;
  BitReader bits;
  const std::function<void(int)> &feedback;
  std::function<void()> begin_frame;  // called at the first SOS: tables and geometry are final (:1857)
  HuffTable huff[2][8];
  PlaneRef plane[3];
  RecordSink sink[3];                 // baseline: components emitted as records while sink[c].on
  std::function<void(int)> spill;     // sink[c] is nearly full: switch component c to its dense plane
  int dc_pred[3] = {0, 0, 0};
  int last_row[3] = {0, 0, 0};  // last block row (natural order) a coefficient was stored in, per component
  int ht_dc[3] = {0, 0, 0}, ht_ac[3] = {0, 0, 0};
  int mcu_count_image_v = 0, mcu_count_image_h = 0;
  int rst_count = 0, next_rst = 0;
  int eobrun = 0;
  int scans = 0;
  int last_feedback = 0;

  Decoder(Image_Descriptor &image, const std::function<void(int)> &fb) : im(image), r(*image.stream), bits(r), feedback(fb) {}

  void emit_feedback(int p) {
    if (p < last_feedback) p = last_feedback;  // clients draw progress bars: never go back
    last_feedback = p;
    feedback(p);
  }

  void rebuild_tables() {
    for (int k = 0; k < 2; k++)
      for (int i = 0; i < 8; i++) huff[k][i].build(im.huff_bits[k][i], im.huff_vals[k][i]);
  }

  void check_restart(bool reset_predictors) {  // :1133-1177
    if (im.restart_interval <= 0) return;
    if (--rst_count != 0) return;
    bits.len &= 0xF8;  // byte alignment
    const int w = bits.get(16);
    // the marker bytes themselves were not pushed into the buffer by this reader:
    // it remembers the marker and pads; accept either view of the 16 bits
    const int expected = 0xFFD0 + next_rst;
    const bool seen = bits.memo_marker == (uint8_t)(0xD0 + next_rst);
    if (!seen && w != expected)
      throw error_in_image_data(fmt("JPEG: expected RST (restart) marker Nb %d; code found %04X; expected %04X",
                                    next_rst, w, expected));
    bits.reset();
    next_rst = (next_rst + 1) & 7;
    rst_count = im.restart_interval;
    eobrun = 0;
    if (reset_predictors) dc_pred[0] = dc_pred[1] = dc_pred[2] = 0;
  }

  const HuffTable &table(int kind, int idx) {
    if (idx < 0 || idx > 7 || !huff[kind][idx].defined) throw constraint_error("JPEG: Huffman table not defined");
    return huff[kind][idx];
  }

  // Decode_8x8_Block (:758-788) without the de-quantisation: values are stored as
  // read, at their natural position -- as records (one per non-zero value) or into the
  // dense plane; the multiplication by qt_local moves to the GPU.
  void baseline_block(int c, int bx, int by) {
    RecordSink &k = sink[c];
    if (k.on && k.n + 64 > k.cap) spill(c);
    int16_t *blk = k.on ? nullptr : plane[c].block(bx, by);
    if (k.on) k.first[k.seq++] = (uint32_t)k.n;
    const int s = bits.symbol(table(0, ht_dc[c]));
    dc_pred[c] += bits.receive_extend(s & 15);
    if (blk) blk[0] = (int16_t)dc_pred[c];
    else if ((int16_t)dc_pred[c] != 0) k.put(0, dc_pred[c]);
    const HuffTable &ac = table(1, ht_ac[c]);
    int coef = 0;
    for (;;) {
      const int code = bits.symbol(ac);
      if (code == 0) break;  // EOB
      if ((code & 0x0F) == 0 && code != 0xF0)
        throw error_in_image_data("JPEG: error in VLC AC code for de-quantization");
      coef += (code >> 4) + 1;
      if (coef > 63) throw error_in_image_data("JPEG: coefficient for de-quantization is > 63");
      const int nat = kZigZag[coef];
      const int value = bits.receive_extend(code & 15);
      if (blk) {
        blk[nat] = (int16_t)value;
        if ((nat >> 3) > last_row[c]) last_row[c] = nat >> 3;
      } else if ((int16_t)value != 0) {
        k.put(nat, value);
      }
      if (coef == 63) break;
    }
  }

  void baseline_scan() {  // :1179-1220; loops over FRAME components like the reference (:1190-1191)
    const int mh = mcu_count_image_h, mv = mcu_count_image_v;
    rst_count = im.restart_interval;
    next_rst = 0;
    for (int my = 0; my < mv; my++) {
      for (int mx = 0; mx < mh; mx++) {
        for (int c = 0; c < 3; c++) {
          if (!im.info[c].present) continue;
          const Component_Info &ci = im.info[c];
          for (int sy = 0; sy < ci.samples_ver; sy++)
            for (int sx = 0; sx < ci.samples_hor; sx++)
              baseline_block(c, mx * ci.samples_hor + sx, my * ci.samples_ver + sy);
        }
        if (!(my == mv - 1 && mx == mh - 1)) check_restart(true);
      }
      emit_feedback((50 * (my + 1)) / mv);
    }
  }

  // ---- progressive (:1243-1747; ITU T.81 annex G) ----

  void dc_first(int c, int16_t *blk, int al) {
    const int s = bits.symbol(table(0, ht_dc[c]));
    dc_pred[c] += bits.receive_extend(s & 15);
    blk[0] = (int16_t)(dc_pred[c] * (1 << al));
  }
  void dc_refine(int16_t *blk, int al) {
    if (bits.get(1)) blk[0] = (int16_t)(blk[0] | (1 << al));
  }
  void ac_first(int c, int16_t *blk, int ss, int se, int al) {
    if (eobrun > 0) { eobrun--; return; }
    const HuffTable &ac = table(1, ht_ac[c]);
    for (int k = ss; k <= se;) {
      const int rs = bits.symbol(ac), rr = rs >> 4, s = rs & 15;
      if (s == 0) {
        if (rr < 15) {
          eobrun = (1 << rr) - 1;
          if (rr) eobrun += bits.get(rr);
          break;
        }
        k += 16;
      } else {
        k += rr;
        if (k > 63) throw error_in_image_data("JPEG: progressive: coefficient index > 63");
        const int nat = kZigZag[k];
        blk[nat] = (int16_t)(bits.receive_extend(s) * (1 << al));
        if ((nat >> 3) > last_row[c]) last_row[c] = nat >> 3;
        k++;
      }
    }
  }
  void ac_refine(int c, int16_t *blk, int ss, int se, int al) {
    const int p1 = 1 << al, m1 = -(1 << al);
    int k = ss;
    if (eobrun <= 0) {
      const HuffTable &ac = table(1, ht_ac[c]);
      for (; k <= se; k++) {
        const int rs = bits.symbol(ac);
        int rr = rs >> 4;
        const int s = rs & 15;
        int value = 0;
        if (s == 0) {
          if (rr < 15) {
            eobrun = 1 << rr;
            if (rr) eobrun += bits.get(rr);
            break;
          }
        } else {
          if (s != 1) throw error_in_image_data("JPEG: progressive: AC refinement with a size other than 1");
          value = bits.get(1) ? p1 : m1;
        }
        for (; k <= se; k++) {
          int16_t &z = blk[kZigZag[k]];
          if (z != 0) {
            if (bits.get(1) && (z & p1) == 0) z = (int16_t)(z >= 0 ? z + p1 : z + m1);
          } else {
            if (rr == 0) {
              if (value) {
                z = (int16_t)value;
                if ((kZigZag[k] >> 3) > last_row[c]) last_row[c] = kZigZag[k] >> 3;
              }
              break;
            }
            rr--;
          }
        }
      }
    }
    if (eobrun > 0) {
      for (; k <= se; k++) {
        int16_t &z = blk[kZigZag[k]];
        if (z != 0 && bits.get(1) && (z & p1) == 0) z = (int16_t)(z >= 0 ? z + p1 : z + m1);
      }
      eobrun--;
    }
  }

  void progressive_scan(const int *comps, int ncomps, int ss, int se, int ah, int al) {
    if (ss > se || se > 63) throw error_in_image_data("JPEG: progressive: bad spectral selection");
    if (ss == 0 && se != 0) throw error_in_image_data("JPEG: progressive: DC and AC in the same scan");
    if (ss > 0 && ncomps != 1) throw error_in_image_data("JPEG: progressive: AC scan with several components");
    rst_count = im.restart_interval;
    next_rst = 0;
    eobrun = 0;
    const bool first_scan = scans == 0;
    if (ncomps > 1) {  // interleaved (DC only): MCU structure
      for (int my = 0; my < mcu_count_image_v; my++) {
        for (int mx = 0; mx < mcu_count_image_h; mx++) {
          for (int i = 0; i < ncomps; i++) {
            const int c = comps[i];
            const Component_Info &ci = im.info[c];
            for (int sy = 0; sy < ci.samples_ver; sy++)
              for (int sx = 0; sx < ci.samples_hor; sx++) {
                int16_t *blk = plane[c].block(mx * ci.samples_hor + sx, my * ci.samples_ver + sy);
                if (ah == 0) dc_first(c, blk, al);
                else dc_refine(blk, al);
              }
          }
          if (!(my == mcu_count_image_v - 1 && mx == mcu_count_image_h - 1)) check_restart(true);
        }
        if (first_scan) emit_feedback((50 * (my + 1)) / mcu_count_image_v);
      }
      return;
    }
    // single component: 8x8 "MCUs" over the component's own size (:1952-1971)
    const int c = comps[0];
    const Component_Info &ci = im.info[c];
    const int wc = (im.width * ci.samples_hor + im.max_samples_hor - 1) / im.max_samples_hor;
    const int hc = (im.height * ci.samples_ver + im.max_samples_ver - 1) / im.max_samples_ver;
    const int bw = (wc + 7) / 8, bh = (hc + 7) / 8;
    for (int by = 0; by < bh; by++) {
      for (int bx = 0; bx < bw; bx++) {
        int16_t *blk = plane[c].block(bx, by);
        if (ss == 0) {
          if (ah == 0) dc_first(c, blk, al);
          else dc_refine(blk, al);
        } else if (ah == 0) {
          ac_first(c, blk, ss, se, al);
        } else {
          ac_refine(c, blk, ss, se, al);
        }
        if (!(by == bh - 1 && bx == bw - 1)) check_restart(true);
      }
      if (first_scan) emit_feedback((50 * (by + 1)) / bh);
    }
  }

  // Read_SOS (:1855-2034): scan header, then the scan's entropy-coded data.
  void read_sos(const Segment &sg) {
    const int n = r.byte();
    int ncomp_image = 0;
    for (const Component_Info &ci : im.info) ncomp_image += ci.present ? 1 : 0;
    if (n > ncomp_image) throw error_in_image_data("JPEG: Scan segment has more color components than the image");
    if (n < 1) throw error_in_image_data("JPEG: Scan segment without components");
    if (sg.length != 4 + 2 * n) throw error_in_image_data("JPEG: Scan segment to be decoded: bad length");
    int comps[3];
    int id_base = 1;
    for (int i = 0; i < n; i++) {
      const int b = r.byte();
      if (b == 0) id_base = 0;
      const int c = b - id_base;
      if (c < 0 || c > 2 || !im.info[c].present) throw error_in_image_data(fmt("JPEG: Scan: invalid ID: %d", b));
      comps[i] = c;
      const int t = r.byte();
      ht_dc[c] = t >> 4;
      ht_ac[c] = t & 15;
    }
    const int ss = r.byte(), se = r.byte(), approx = r.byte();
    const int mw = 8 * im.max_samples_hor, mh = 8 * im.max_samples_ver;
    mcu_count_image_h = (im.width + mw - 1) / mw;
    mcu_count_image_v = (im.height + mh - 1) / mh;
    // a subsampled component must be at least 3 samples wide / high (:2014-2019)
    for (int i = 0; i < n; i++) {
      const Component_Info &ci = im.info[comps[i]];
      const int wc = (im.width * ci.samples_hor + im.max_samples_hor - 1) / im.max_samples_hor;
      const int hc = (im.height * ci.samples_ver + im.max_samples_ver - 1) / im.max_samples_ver;
      if ((wc < 3 && ci.samples_hor != im.max_samples_hor) || (hc < 3 && ci.samples_ver != im.max_samples_ver))
        throw error_in_image_data("JPEG: component: sample dimension mismatch");
    }
    if (!plane[0].base) begin_frame();
    rebuild_tables();
    bits.reset();
    dc_pred[0] = dc_pred[1] = dc_pred[2] = 0;
    if (im.progressive) progressive_scan(comps, n, ss, se, approx >> 4, approx & 15);
    else baseline_scan();
    scans++;
  }

  // the segment loop of Load (:2043-2118)
  void run() {
    bool have_marker = false;
    uint8_t buffered = 0;
    for (;;) {
      Segment sg = read_segment(r, have_marker, buffered);
      have_marker = false;
      switch (sg.marker) {
        case M_DQT: read_dqt(im, r, sg.length); break;
        case M_DHT: read_dht(im, r, sg.length); break;
        case M_DRI: im.restart_interval = (int)r.be16(); break;
        case M_EOI:
          if (im.progressive && scans == 0) throw error_in_image_data("JPEG: progressive image without scans");
          return;
        case M_SOS:
          read_sos(sg);
          if (!im.progressive) return;  // baseline: the first scan is the image (:2077)
          if (bits.memo_marker != 0) {  // the marker that ended the scan was swallowed by the bit reader
            have_marker = true;
            buffered = bits.memo_marker;
          }
          break;
        default: r.skip(sg.length); break;
      }
    }
  }
};

[[noreturn]] void raise_from_status(int rc, gidcuda_ctx *ctx) {
  const char *m = gidcuda_last_error(ctx);
  const std::string msg = std::string("JPEG (CUDA): ") + (m ? m : "");
  switch (rc) {
    case GIDCUDA_ERR_UNSUPPORTED: throw unsupported_image_subformat(msg);
    case GIDCUDA_ERR_BAD_ARGUMENT: throw error_in_image_data(msg);
    default: throw cuda_device_error(msg + fmt(" (status %d)", rc));
  }
}

}  // namespace

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
}

namespace detail {

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
  dec.begin_frame = [&]() {
    fill_tables(image, desc, comp_of);
    for (int c = 0; c < 3; c++)
      if (image.info[c].present) {
        int32_t bx = 0, by = 0;
        const int st = gidcuda_plane_geometry(desc, comp_of[c], &bx, &by);
        if (st != GIDCUDA_OK) raise_from_status(st, nullptr);
        int16_t *p = static_cast<int16_t *>(calloc((size_t)bx * (size_t)by * 64, sizeof(int16_t)));
        if (!p) throw cuda_device_error("out of host memory");
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
    for (int i = 0; i < 64; i++)
      if (blk[i] != 0) rec[n++] = ((uint32_t)(uint16_t)blk[i] << 16) | (uint32_t)i;
  }
  first[nb] = (uint32_t)n;
  *count = n;
  return true;
}

Reconstruction decode_and_reconstruct(Image_Descriptor &image, const std::function<void(int)> &feedback) {
  if (!image.stream) throw constraint_error("Load_Image_Contents: no header was loaded");
  gidcuda_frame_desc desc;
  int comp_of[3] = {0, 0, 0};
  describe_frame(image, &desc, comp_of);

  gidcuda_ctx *ctx = nullptr;
  int rc = gidcuda_create(image.device, &ctx);
  if (rc != GIDCUDA_OK) raise_from_status(rc, nullptr);
  gidcuda_job *job = nullptr;
  // progressive + sparse input: the scans accumulate in ordinary host memory
  std::vector<std::vector<int16_t>> accum(3);
  const bool sparse = image.sparse_input;
  // The job is begun at the first SOS, like Begin_Device_Frame of the Ada patch:
  // every DQT has been read by then (GID accepts none after entropy-coded data).
  Decoder dec(image, feedback);
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
      int st = gidcuda_job_begin(ctx, &desc, &job);
      if (st != GIDCUDA_OK) raise_from_status(st, ctx);
      for (int c = 0; c < 3; c++) {
        if (!image.info[c].present) continue;
        if (!sparse) {
          use_job_plane(c);
        } else if (image.progressive) {
          int32_t bx = 0, by = 0;
          st = gidcuda_plane_geometry(&desc, comp_of[c], &bx, &by);
          if (st != GIDCUDA_OK) raise_from_status(st, ctx);
          accum[c].assign((size_t)bx * (size_t)by * 64, 0);
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
    dec.run();
    if (!job) throw error_in_image_data("JPEG: no scan in the image");
    for (int c = 0; c < 3; c++) {
      if (!image.info[c].present) continue;
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
        memcpy(dec.plane[c].base, acc.base, accum[c].size() * sizeof(int16_t));
      } else if (k.on) {
        if (k.seq != k.nblocks) throw error_in_image_data("JPEG: scan ended before the last block");
        k.first[k.nblocks] = (uint32_t)k.n;
        rc = gidcuda_job_records_commit(job, comp_of[c], k.n);
        if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
        continue;
      }
      rc = gidcuda_job_hint_rows(job, comp_of[c], dec.last_row[c] + 1);
      if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
    }
    rc = gidcuda_job_submit(job);
    if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
    Reconstruction rec;
    rc = gidcuda_job_wait(job, &rec.pixels, &rec.stride);
    if (rc != GIDCUDA_OK) raise_from_status(rc, ctx);
    rec.job = job;
    rec.ctx = ctx;
    return rec;
  } catch (...) {
    if (job) gidcuda_job_release(job);
    gidcuda_destroy(ctx);
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
static int g_sparse_input = 1;
int gidhost_set_option(const char *name, int value) {
  if (name && strcmp(name, "sparse_input") == 0) {
    g_sparse_input = value != 0;
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
    gid::detail::decode_planes(im, desc, planes, blocks_x, blocks_y);
    return GIDHOST_OK;
  } catch (...) {
    for (int c = 0; c < 3; c++) { free(planes[c]); planes[c] = nullptr; }
    return status_of_exception(err, errlen);
  }
}

void gidhost_free(void *p) { free(p); }

}  // extern "C"
