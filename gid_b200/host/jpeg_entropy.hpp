// jpeg_entropy.hpp -- internal header of the C++ host mirror: segment readers and the
// entropy decoder (baseline and progressive) that emits frame-wide int16 coefficient
// planes / coefficient records.  Included by gid_jpeg_host.cpp (GID's API above the C
// ABI) and gid_batch.cpp (many files on many host threads); everything lives in an
// anonymous namespace, so each translation unit gets its own copy.
//
// References are to the GID 013 tree:
//   segment reader            gid-decoding_jpg.adb:122-181
//   Read_SOF                  :184-318      Read_DHT :320-391     Read_DQT :393-422
//   Read_DRI                  :424-434      Read_EXIF :436-543
//   bit buffer / Get_VLC      :594-753
//   Check_Restart             :1133-1177
//   baseline scan             :1179-1220    progressive scans :1243-1747
//   Read_SOS + geometry       :1855-2034    segment loop :2043-2118
//
// Product code: independent of the test-only CPU restatement.
#pragma once

#include "gid.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

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
    // As the reference: per code length, check the bytes left in the segment ("[1]") and the code
    // space ("[2]": over-subscribed), then read that length's symbols.
    long space = 65536, spread = 65536;
    std::vector<uint8_t> vals;
    vals.reserve((size_t)total);
    for (int len = 0; len < 16; len++) {
      spread /= 2;
      if (counts[len] > 0) {
        if (remaining < counts[len]) throw error_in_image_data("JPEG: DHT data too short [1]");
        space -= (long)counts[len] * spread;
        if (space < 0) throw error_in_image_data("JPEG: DHT data too short [2]");
        for (int i = 0; i < counts[len]; i++) vals.push_back(r.byte());
        remaining -= counts[len];
      }
    }
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
  // 10-bit direct table of the wide reader: len << 8 | symbol, 0 = code longer than 10 bits
  static constexpr int kLutBits = 10;
  uint16_t lut[1 << kLutBits];
  int maxcode[18], valptr[17], mincode[17];
  std::vector<uint8_t> vals;
  void build(const std::vector<uint8_t> &counts, const std::vector<uint8_t> &v) {
    defined = !counts.empty();
    if (!defined) return;
    vals = v;
    memset(fast_len, 0, sizeof fast_len);
    memset(lut, 0, sizeof lut);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
      valptr[len] = k;
      mincode[len] = code;
      for (int i = 0; i < counts[(size_t)len - 1]; i++, k++, code++) {
        if (len <= 8) {
          const int first = code << (8 - len), n = 1 << (8 - len);
          for (int j = 0; j < n; j++) { fast_len[first + j] = (uint8_t)len; fast_sym[first + j] = vals[(size_t)k]; }
        }
        if (len <= kLutBits && code < (1 << len)) {
          const int first = code << (kLutBits - len), n = 1 << (kLutBits - len);
          for (int j = 0; j < n; j++) lut[first + j] = (uint16_t)((len << 8) | vals[(size_t)k]);
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
  // Show_Bits' loop (:621-671) as it is: a marker's two bytes go into the buffer like any others, the marker is
  // remembered, and the loading carries on behind it; past the end of the data every byte asked for is FF.
  void fill(int bits) {
    while (len < bits) {
      if (r.s.pos >= r.s.size) {  // End_Error
        buf = (buf << 8) | 0xFFu;
        len += 8;
        continue;
      }
      const uint8_t b = r.s.data[r.s.pos++];
      buf = (buf << 8) | b;
      len += 8;
      if (b != 0xFF) continue;
      if (r.s.pos >= r.s.size) {  // End_Error at the byte behind the FF: the handler puts one more FF in
        buf = (buf << 8) | 0xFFu;
        len += 8;
        continue;
      }
      const uint8_t m = r.s.data[r.s.pos++];
      if (m == 0) continue;       // byte stuffing: FF 00 is a data byte FF
      buf = (buf << 8) | m;
      len += 8;
      memo_marker = m;
      const bool ok = m == M_EOI || m == M_COM || m == M_DHT || m == M_DRI || m == M_SOS || (m >= M_RST0 && m <= M_RST7);
      if (!ok) throw error_in_image_data(fmt("JPEG: Invalid marker within filling of bit buffer: 16#%02X#", m));
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

// The bit buffer again, 64 bits wide.  Same observable behaviour as BitReader / the reference's byte-wise buffer
// (Show_Bits, :609-677):
//   * FF 00 is a stuffed FF; entropy-coded bytes are loaded ahead of the requests, eight at a time where none
//     of them is FF -- which nobody can tell, they are the bytes the byte-wise buffer loads a little later;
//   * at the first marker (FF followed by anything but 00) the loading ahead stops BEFORE it (`lazy`).  From
//     there on bytes are loaded exactly when the byte-wise buffer loads them -- when a request finds fewer bits
//     than it asks for (lazy_fill): the marker's two bytes go into the buffer, the marker is remembered
//     (memo_marker) or raises if it may not appear inside entropy-coded data, and the loading carries on behind
//     it, as the reference's does.  Both buffers hold the same bits at that point (the wide one never went
//     beyond the marker), so they load the same bytes at the same requests from then on, and `pos` is the
//     reference's stream position (`exact`);
//   * past the end of the data every byte asked for is FF (End_Error, :665-670).
// The buffer is left-aligned: the next bit is bit 63; `len` bits are valid, the last `pad` of them are not
// entropy-coded data of the scan (marker bytes, what follows them, FF past the end).
struct WideBits {
  const uint8_t *data;
  size_t pos, size;
  uint64_t buf = 0;
  int len = 0, pad = 0;
  uint8_t memo_marker = 0;
  bool stopped = false;  // past the end of the data
  bool lazy = false;     // the next byte is the FF of a marker (or an FF that ends the data): no loading ahead
  bool exact = false;    // bytes were loaded on request: pos and memo_marker are the reference's
  bool overrun = false;  // bits that are not the scan's entropy-coded data have been consumed
  // Progressive scans: the segment loop carries on where the reference's byte-wise buffer stopped
  // reading.  With `track` set the wide buffer follows that buffer's fill level (it loads a byte
  // whenever it holds fewer bits than are asked for, :622-690) and counts the bytes it would
  // have loaded; reference_position() turns the count into a stream position.
  bool track = false;
  int ref_len = 0;
  size_t ref_loaded = 0;
  WideBits(const uint8_t *d, size_t p, size_t n) : data(d), pos(p), size(n) {}
  void reset() {
    buf = 0; len = 0; pad = 0; memo_marker = 0; stopped = false; lazy = false; exact = false; overrun = false;
    ref_len = 0; ref_loaded = 0;
  }
  int real() const { return len - pad; }
  void refill() {
    if (lazy) return;
    if (!stopped && len <= 32 && pos + 8 <= size) {  // eight bytes at once when none of them is FF
      uint64_t v;
      memcpy(&v, data + pos, 8);
      v = __builtin_bswap64(v);
      const uint64_t inv = ~v;  // a byte of v is FF <=> the same byte of inv is zero
      if (((inv - 0x0101010101010101ull) & ~inv & 0x8080808080808080ull) == 0) {
        const int nb = (64 - len) >> 3, spare = 64 - len - 8 * nb;
        buf |= (v >> len) & ~((1ull << spare) - 1ull);
        len += 8 * nb;
        pos += (size_t)nb;
        return;
      }
    }
    while (len <= 56) {
      uint64_t b = 0xFF;
      if (stopped || pos >= size) {
        stopped = true;  // the end of the data: FF padding (:665-670)
        pad += 8;
      } else {
        b = data[pos];
        if (b == 0xFF) {
          if (pos + 1 >= size || data[pos + 1] != 0) {  // a marker, or an FF that ends the data
            lazy = true;
            return;
          }
          pos += 2;  // byte stuffing
        } else {
          pos++;
        }
      }
      buf |= b << (56 - len);
      len += 8;
    }
  }
  // Show_Bits' loop as it is (:621-671), from the marker on
  void push_beyond(uint64_t b) {
    buf |= b << (56 - len);
    len += 8;
    pad += 8;
  }
  void lazy_fill(int n) {
    exact = true;
    while (len < n) {
      if (pos >= size) { push_beyond(0xFF); continue; }  // End_Error
      const uint8_t b = data[pos++];
      push_beyond(b);
      if (b != 0xFF) continue;
      if (pos >= size) { push_beyond(0xFF); continue; }  // End_Error at the byte behind the FF
      const uint8_t m = data[pos++];
      if (m == 0) continue;  // byte stuffing
      push_beyond(m);
      memo_marker = m;
      const bool ok = m == M_EOI || m == M_COM || m == M_DHT || m == M_DRI || m == M_SOS || (m >= M_RST0 && m <= M_RST7);
      if (!ok) throw error_in_image_data(fmt("JPEG: Invalid marker within filling of bit buffer: 16#%02X#", m));
    }
    ref_len = len;
  }
  // the decoder is about to look at n bits (n <= 32)
  void need(int n) {
    if (track && !exact)
      while (ref_len < n) { ref_len += 8; ref_loaded++; }
    if (len < n) {
      refill();
      if (len < n) lazy_fill(n);
    }
  }
  int peek(int n) const { return (int)(buf >> (64 - n)); }  // 1 <= n <= 32
  void drop(int n) {
    buf <<= n;
    len -= n;
    ref_len -= n;
    if (pad > len) {
      pad = len;
      overrun = true;
    }
  }
  // Where the byte-wise buffer would stand after the same requests, the scan having begun at
  // `start`: the stream position after the ref_loaded-th data byte (FF 00 counts as one), or just
  // past the first marker when it would have run into it (*marker = that marker, else 0).
  size_t reference_position(size_t start, uint8_t *marker) const {
    if (exact) {  // bytes were loaded on request from the marker on: the position is the reference's own
      *marker = memo_marker;
      return pos;
    }
    size_t i = start, count = 0;
    *marker = 0;
    while (count < ref_loaded && i < size) {
      const uint8_t b = data[i++];
      if (b == 0xFF) {
        const uint8_t m = i < size ? data[i++] : (uint8_t)M_EOI;
        if (m != 0) { *marker = m; break; }
      }
      count++;
    }
    return i;
  }
  int get(int n) {  // n = 0 .. 16
    if (n == 0) return 0;
    need(n);
    const int v = peek(n);
    drop(n);
    return v;
  }
  // n (1 .. 16) bits that the caller would otherwise ask for ONE AT A TIME (the correction bits of an AC
  // refinement scan): the same bits, and the same bookkeeping of what the byte-wise buffer would have loaded
  // by then (it loads a byte whenever it is empty), in one step
  int get_serial(int n) {
    if (track && !exact && n > ref_len) {
      const int k = (n - ref_len + 7) >> 3;
      ref_len += 8 * k;
      ref_loaded += (size_t)k;
    }
    if (len < n) {
      refill();
      if (len < n) lazy_fill(n);
    }
    const int v = peek(n);
    drop(n);
    return v;
  }
  int symbol(const HuffTable &t) {  // Next_Huffval, :738-746: always looks at 16 bits
    need(16);
    const int look = peek(16);
    const uint16_t e = t.lut[look >> (16 - HuffTable::kLutBits)];
    if (e != 0) {
      drop(e >> 8);
      return e & 0xFF;
    }
    for (int l = HuffTable::kLutBits + 1; l <= 16; l++) {
      const int code = look >> (16 - l);
      if (code <= t.maxcode[l]) {
        drop(l);
        return t.vals[(size_t)(t.valptr[l] + code - t.mincode[l])];
      }
    }
    throw error_in_image_data("JPEG: VLC table: bits = 0");
  }
  int receive_extend(int bits) { return bits ? BitReader::extend(get(bits), bits) : 0; }
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
  // called once before the first coefficient is stored on the HOST (planes / record sinks are set
  // up here rather than in begin_frame, so that a scan decoded on the device never reserves them)
  std::function<void()> prepare_host_output;
  bool host_output_ready = false;
  bool frame_begun = false;
  void ensure_host_output() {
    if (host_output_ready) return;
    host_output_ready = true;
    if (prepare_host_output) prepare_host_output();
  }
  HuffTable huff[2][8];
  PlaneRef plane[3];
  RecordSink sink[3];                 // baseline: components emitted as records while sink[c].on
  std::function<void(int)> spill;     // sink[c] is nearly full: switch component c to its dense plane
  int dc_pred[3] = {0, 0, 0};
  int last_row[3] = {0, 0, 0};  // last block row (natural order) a coefficient was stored in, per component
  // OR of the magnitudes stored per component (v >= 0 ? v : ~v): below 2**n exactly when every magnitude is.
  // The DC predictor enters with all its 32 bits (:766-771), before it is cut to the planes' 16.
  // Read after the last scan: gidcuda_job_wide_coefficients / the range of the planes (gid_jpeg_host.cpp).
  uint32_t mag[3] = {0, 0, 0};
  static uint32_t magnitude(int v) { return (uint32_t)(v ^ (v >> 31)); }
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
    const int expected = 0xFFD0 + next_rst;
    if (w != expected)
      throw error_in_image_data(fmt("JPEG: expected RST (restart) marker Nb %d; code found %04X; expected %04X",
                                    next_rst, w, expected));
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
  // k = record sink (blk == nullptr) or nullptr (dense block blk).  Static: the restart-parallel
  // scan runs it on several threads, each with its own bit buffer, predictor and sink.
  static void baseline_block(WideBits &bits, const HuffTable &dc, const HuffTable &ac, int &dc_pred, RecordSink *k,
                             int16_t *blk, int &last_row, uint32_t &mag) {
    if (k) k->first[k->seq++] = (uint32_t)k->n;
    const int s = bits.symbol(dc);
    dc_pred += bits.receive_extend(s & 15);
    mag |= magnitude(dc_pred);
    if (blk) blk[0] = (int16_t)dc_pred;
    else if ((int16_t)dc_pred != 0) k->put(0, dc_pred);
    int coef = 0;
    for (;;) {
      int code, value;
      if (bits.len <= 32) bits.refill();
      if (bits.real() >= 32) {
        // symbol (<= 16 bits) and its value bits (<= 15) are all real bits: no checks needed
        const int look = bits.peek(16);
        const uint16_t e = ac.lut[look >> (16 - HuffTable::kLutBits)];
        if (e != 0) {
          bits.drop(e >> 8);
          code = e & 0xFF;
        } else {
          code = bits.symbol(ac);
        }
        if (code == 0) break;  // EOB
        const int n = code & 15;
        if (n == 0 && code != 0xF0) throw error_in_image_data("JPEG: error in VLC AC code for de-quantization");
        coef += (code >> 4) + 1;
        if (coef > 63) throw error_in_image_data("JPEG: coefficient for de-quantization is > 63");
        value = 0;
        if (n) {
          value = BitReader::extend(bits.peek(n), n);
          bits.drop(n);
        }
      } else {
        code = bits.symbol(ac);
        if (code == 0) break;  // EOB
        if ((code & 0x0F) == 0 && code != 0xF0)
          throw error_in_image_data("JPEG: error in VLC AC code for de-quantization");
        coef += (code >> 4) + 1;
        if (coef > 63) throw error_in_image_data("JPEG: coefficient for de-quantization is > 63");
        value = bits.receive_extend(code & 15);
      }
      const int nat = kZigZag[coef];
      mag |= magnitude(value);
      if (blk) {
        blk[nat] = (int16_t)value;
        if ((nat >> 3) > last_row) last_row = nat >> 3;
      } else if ((int16_t)value != 0) {
        k->put(nat, value);
      }
      if (coef == 63) break;
    }
  }

  // Check_Restart (:1133-1177) on the wide buffer: byte-align, the next 16 bits must be the
  // expected RSTn -- the marker's own two bytes, loaded by this request (WideBits::lazy_fill).
  void check_restart_wide(WideBits &wb) {
    if (im.restart_interval <= 0) return;
    if (--rst_count != 0) return;
    wb.drop(wb.len & 7);
    wb.need(16);
    if (wb.peek(16) != 0xFFD0 + next_rst)
      throw error_in_image_data(fmt("JPEG: expected RST (restart) marker Nb %d; code found %04X; expected %04X",
                                    next_rst, wb.peek(16), 0xFFD0 + next_rst));
    // The 16 bits are consumed and nothing else changes (:1141-1166): what the buffer holds behind them stays --
    // 16 bits that merely LOOK like the marker (a stuffed FF followed by Dn) are followed by data already loaded.
    wb.drop(16);
    wb.lazy = false;  // past the marker: load ahead again
    next_rst = (next_rst + 1) & 7;
    rst_count = im.restart_interval;
    dc_pred[0] = dc_pred[1] = dc_pred[2] = 0;
  }

  // ---- restart-interval-parallel baseline scan ----------------
  // After RSTn the predictors are reset and the bit buffer is byte-aligned (:1133-1177), so the
  // restart intervals of a scan decode independently.  The scan's bytes are searched for the
  // markers first; `threads` workers then decode contiguous groups of intervals, each with its
  // own bit buffer, into disjoint blocks of the dense planes or into its own slice of the record
  // arrays (one RUN per worker, handed over with gidcuda_job_records_commit_runs).
  // Anything unusual -- markers out of sequence or not as many as the geometry needs, an interval
  // with bytes left over, a record slice that fills up, any exception in a worker -- makes the
  // function return false with the output state rewound, and the serial loop decodes the scan:
  // error behaviour and messages stay those of the serial decoder.
  struct Run { size_t begin[3] = {0, 0, 0}, count[3] = {0, 0, 0}; };
  std::vector<Run> runs;  // one per worker after a successful parallel scan, else empty
  int parallel_threads = 1;
  bool wide_progressive = true;  // false: progressive scans through the byte-wise buffer (tests compare the two)

  // Where the restart intervals of the scan at r.s.pos begin and end: interval k runs from the
  // byte after RSTk-1 to the FF of RSTk (the last one to the marker that ends the scan, or to the
  // end of the data).  False unless the markers are exactly the nseg - 1 the geometry needs,
  // numbered in sequence.
  bool find_intervals(long nseg, std::vector<size_t> &seg_begin, std::vector<size_t> &seg_end) const {
    const uint8_t *d = r.s.data;
    const size_t size = r.s.size;
    seg_begin.assign((size_t)nseg, 0);
    seg_end.assign((size_t)nseg, 0);
    size_t i = r.s.pos;
    long k = 0;
    seg_begin[0] = i;
    for (;;) {
      const void *f = i < size ? memchr(d + i, 0xFF, size - i) : nullptr;
      if (!f) { seg_end[(size_t)k] = size; break; }
      i = (size_t)((const uint8_t *)f - d);
      const uint8_t m = i + 1 < size ? d[i + 1] : (uint8_t)M_EOI;
      if (m == 0) { i += 2; continue; }
      seg_end[(size_t)k] = i;
      if (m >= M_RST0 && m <= M_RST7) {
        if (m != (uint8_t)(M_RST0 + (k & 7)) || k + 1 >= nseg) return false;
        seg_begin[(size_t)++k] = i + 2;
        i += 2;
        continue;
      }
      break;  // the marker that ends the scan (or one the serial decoder will complain about)
    }
    return k == nseg - 1;
  }

  // ---- the scan handed to the device as compressed bytes (gidcuda_job_scan) -----------------------
  // Set by the caller that owns the job; returns false when the device path does not apply.
  struct DeviceScan {
    const uint8_t *data;
    size_t len;
    std::vector<uint32_t> begin, end;  // relative to data
    int restart_interval;              // 0: a scan without restart markers; data = `unstuffed`
    std::vector<uint8_t> unstuffed;
    const std::vector<uint8_t> *dc_bits[3], *dc_vals[3], *ac_bits[3], *ac_vals[3];
  };
  std::function<bool(const DeviceScan &)> device_scan;
  bool device_scanned = false;

  void device_scan_tables(DeviceScan &ds) {
    for (int c = 0; c < 3; c++) {
      ds.dc_bits[c] = ds.dc_vals[c] = ds.ac_bits[c] = ds.ac_vals[c] = nullptr;
      if (!im.info[c].present) continue;
      table(0, ht_dc[c]);  // raises like the serial loop when a table is missing
      table(1, ht_ac[c]);
      ds.dc_bits[c] = &im.huff_bits[0][ht_dc[c]];
      ds.dc_vals[c] = &im.huff_vals[0][ht_dc[c]];
      ds.ac_bits[c] = &im.huff_bits[1][ht_ac[c]];
      ds.ac_vals[c] = &im.huff_vals[1][ht_ac[c]];
    }
  }

  // A scan without restart markers: the bytes up to the marker that ends it, stuffing removed
  // (FF 00 -> FF, :640-662), go to the device's self-synchronising decoder.
  bool device_whole_scan() {
    if (!device_scan || im.restart_interval > 0) return false;
    const uint8_t *d = r.s.data;
    const size_t size = r.s.size;
    DeviceScan ds;
    ds.restart_interval = 0;
    size_t i = r.s.pos;
    if (size - i < 4096) return false;  // a small scan is decoded on the host before the extra launches are queued
    ds.unstuffed.reserve(size - i);
    for (;;) {
      const void *f = i < size ? memchr(d + i, 0xFF, size - i) : nullptr;
      const size_t j = f ? (size_t)((const uint8_t *)f - d) : size;
      ds.unstuffed.insert(ds.unstuffed.end(), d + i, d + j);
      if (!f) break;
      const uint8_t m = j + 1 < size ? d[j + 1] : (uint8_t)M_EOI;
      if (m != 0) {
        if (m >= M_RST0 && m <= M_RST7) return false;  // restart markers without DRI: the serial loop's case
        break;
      }
      ds.unstuffed.push_back(0xFF);
      i = j + 2;
    }
    if (ds.unstuffed.empty()) return false;
    ds.data = ds.unstuffed.data();
    ds.len = ds.unstuffed.size();
    device_scan_tables(ds);
    if (!device_scan(ds)) return false;
    device_scanned = true;
    emit_feedback(50);
    return true;
  }

  bool device_baseline_scan() {
    const int ri = im.restart_interval;
    const long nmcu = (long)mcu_count_image_h * mcu_count_image_v;
    if (!device_scan || ri <= 0) return false;
    const long nseg = (nmcu + ri - 1) / ri;
    if (nseg < 64) return false;  // too few independent streams to occupy a GPU
    std::vector<size_t> seg_begin, seg_end;
    if (!find_intervals(nseg, seg_begin, seg_end)) return false;
    DeviceScan ds;
    const size_t base = seg_begin[0];
    ds.data = r.s.data + base;
    ds.len = seg_end[(size_t)nseg - 1] - base;
    if (ds.len == 0 || ds.len > 0xFFFFFFF0u) return false;
    ds.begin.resize((size_t)nseg);
    ds.end.resize((size_t)nseg);
    for (long k = 0; k < nseg; k++) {
      ds.begin[(size_t)k] = (uint32_t)(seg_begin[(size_t)k] - base);
      ds.end[(size_t)k] = (uint32_t)(seg_end[(size_t)k] - base);
    }
    ds.restart_interval = ri;
    device_scan_tables(ds);
    if (!device_scan(ds)) return false;
    device_scanned = true;
    emit_feedback(50);
    return true;
  }

  bool parallel_baseline_scan() {
    const int ri = im.restart_interval;
    const long nmcu = (long)mcu_count_image_h * mcu_count_image_v;
    if (ri <= 0 || parallel_threads < 2 || nmcu < 2048) return false;
    const long nseg = (nmcu + ri - 1) / ri;
    const int T = (int)std::min<long>(parallel_threads, nseg / 4);
    if (T < 2) return false;
    std::vector<size_t> seg_begin, seg_end;
    if (!find_intervals(nseg, seg_begin, seg_end)) return false;
    const uint8_t *d = r.s.data;
    // -- workers
    int ncomp_blocks[3] = {0, 0, 0};
    const HuffTable *tdc[3] = {nullptr, nullptr, nullptr}, *tac[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < 3; c++)
      if (im.info[c].present) {
        ncomp_blocks[c] = im.info[c].samples_hor * im.info[c].samples_ver;
        tdc[c] = &table(0, ht_dc[c]);
        tac[c] = &table(1, ht_ac[c]);
      }
    std::vector<Run> out((size_t)T);
    std::vector<std::array<int, 3>> rows((size_t)T, std::array<int, 3>{{0, 0, 0}});
    std::vector<std::array<uint32_t, 3>> mags((size_t)T, std::array<uint32_t, 3>{{0, 0, 0}});
    std::atomic<bool> failed{false};
    const int mh = mcu_count_image_h;
    auto work = [&](int w) {
      try {
        const long s0 = nseg * w / T, s1 = nseg * (w + 1) / T;
        RecordSink loc[3];
        for (int c = 0; c < 3; c++)
          if (im.info[c].present && sink[c].on) {
            const uint64_t seq0 = (uint64_t)s0 * ri * ncomp_blocks[c];
            const uint64_t seq1 = std::min<uint64_t>((uint64_t)s1 * ri, (uint64_t)nmcu) * ncomp_blocks[c];
            // a slice of the record array in proportion to the blocks (32 records per block)
            const size_t b0 = (size_t)(sink[c].cap / sink[c].nblocks * seq0);
            const size_t b1 = (size_t)(sink[c].cap / sink[c].nblocks * seq1);
            loc[c].first = sink[c].first;
            loc[c].rec = sink[c].rec + b0;
            loc[c].cap = b1 - b0;
            loc[c].seq = (uint32_t)seq0;
            loc[c].on = true;
            out[(size_t)w].begin[c] = b0;
          }
        for (long sg = s0; sg < s1 && !failed.load(std::memory_order_relaxed); sg++) {
          WideBits wb(d, seg_begin[(size_t)sg], seg_end[(size_t)sg]);
          int pred[3] = {0, 0, 0};
          const long m1 = std::min<long>((sg + 1) * ri, nmcu);
          for (long m = sg * ri; m < m1; m++) {
            const int mx = (int)(m % mh), my = (int)(m / mh);
            for (int c = 0; c < 3; c++) {
              if (!im.info[c].present) continue;
              const Component_Info &ci = im.info[c];
              RecordSink *k = loc[c].on ? &loc[c] : nullptr;
              if (k && k->n + 64 * (size_t)ncomp_blocks[c] > k->cap) { failed = true; return; }
              for (int sy = 0; sy < ci.samples_ver; sy++)
                for (int sx = 0; sx < ci.samples_hor; sx++)
                  baseline_block(wb, *tdc[c], *tac[c], pred[c], k,
                                 k ? nullptr : plane[c].block(mx * ci.samples_hor + sx, my * ci.samples_ver + sy),
                                 rows[(size_t)w][(size_t)c], mags[(size_t)w][(size_t)c]);
            }
          }
          // every interval but the last must end where its RSTn begins: all bytes loaded, less
          // than a byte of real bits left (the serial check tolerates more; let it decide then)
          // (an interval that needed more bits than it has was fed FF bytes here, not what follows it in the file)
          if (wb.overrun) { failed = true; return; }
          if (sg != nseg - 1) {
            wb.drop(wb.len & 7);
            if (wb.pos < wb.size || wb.real() >= 8) { failed = true; return; }  // (pos < size: bytes left over, or a marker inside the interval)
          }
        }
        for (int c = 0; c < 3; c++) out[(size_t)w].count[c] = loc[c].n;
      } catch (...) {
        failed = true;
      }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < T; w++) pool.emplace_back(work, w);
    work(0);
    for (std::thread &t : pool) t.join();
    (failed ? detail::parallel_fallbacks : detail::parallel_scans).fetch_add(1, std::memory_order_relaxed);
    if (failed) {
      for (int c = 0; c < 3; c++)  // rewind: the serial loop starts from clean planes
        if (im.info[c].present && !sink[c].on && plane[c].base)
          memset(plane[c].base, 0, (size_t)plane[c].blocks_x * (size_t)plane[c].blocks_y * 64 * sizeof(int16_t));
      return false;
    }
    // -- stitch: first[] numbers the records of the concatenated runs
    for (int c = 0; c < 3; c++) {
      if (!im.info[c].present) continue;
      for (int w = 0; w < T; w++) mag[c] |= mags[(size_t)w][(size_t)c];
      if (!sink[c].on) {
        for (int w = 0; w < T; w++) last_row[c] = std::max(last_row[c], rows[(size_t)w][(size_t)c]);
        continue;
      }
      size_t base = 0;
      for (int w = 0; w < T; w++) {
        const long s0 = nseg * w / T, s1 = nseg * (w + 1) / T;
        const uint64_t seq0 = (uint64_t)s0 * ri * ncomp_blocks[c];
        const uint64_t seq1 = std::min<uint64_t>((uint64_t)s1 * ri, (uint64_t)nmcu) * ncomp_blocks[c];
        if (base)
          for (uint64_t q = seq0; q < seq1; q++) sink[c].first[q] += (uint32_t)base;
        base += out[(size_t)w].count[c];
      }
      sink[c].n = base;
      sink[c].seq = sink[c].nblocks;
    }
    runs = out;
    emit_feedback(50);
    return true;
  }

  void baseline_scan() {  // :1179-1220; loops over FRAME components like the reference (:1190-1191)
    const int mh = mcu_count_image_h, mv = mcu_count_image_v;
    if (device_baseline_scan() || device_whole_scan()) return;
    ensure_host_output();
    if (parallel_baseline_scan()) return;
    rst_count = im.restart_interval;
    next_rst = 0;
    WideBits wb(r.s.data, r.s.pos, r.s.size);
    const HuffTable *tdc[3] = {nullptr, nullptr, nullptr}, *tac[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < 3; c++)
      if (im.info[c].present) {
        tdc[c] = &table(0, ht_dc[c]);
        tac[c] = &table(1, ht_ac[c]);
      }
    for (int my = 0; my < mv; my++) {
      for (int mx = 0; mx < mh; mx++) {
        for (int c = 0; c < 3; c++) {
          if (!im.info[c].present) continue;
          const Component_Info &ci = im.info[c];
          for (int sy = 0; sy < ci.samples_ver; sy++)
            for (int sx = 0; sx < ci.samples_hor; sx++) {
              RecordSink &k = sink[c];
              if (k.on && k.n + 64 > k.cap) spill(c);
              baseline_block(wb, *tdc[c], *tac[c], dc_pred[c], k.on ? &k : nullptr,
                             k.on ? nullptr : plane[c].block(mx * ci.samples_hor + sx, my * ci.samples_ver + sy),
                             last_row[c], mag[c]);
            }
        }
        if (!(my == mv - 1 && mx == mh - 1)) check_restart_wide(wb);
      }
      emit_feedback((50 * (my + 1)) / mv);
    }
    r.s.pos = wb.pos;
    bits.memo_marker = wb.memo_marker;
  }

  // ---- progressive (:1243-1747; ITU T.81 annex G) ----

  template <class B> void dc_first(B &bits, int c, int16_t *blk, int al) {
    const int s = bits.symbol(table(0, ht_dc[c]));
    dc_pred[c] += bits.receive_extend(s & 15);
    mag[c] |= magnitude(dc_pred[c] * (1 << al));
    blk[0] = (int16_t)(dc_pred[c] * (1 << al));
  }
  template <class B> void dc_refine(B &bits, int c, int16_t *blk, int al) {
    if (bits.get(1)) {  // (added, not OR-ed: :1320-1326)
      const int v = blk[0] + (1 << al);
      mag[c] |= magnitude(v);
      blk[0] = (int16_t)v;
    }
  }
  // Which AC coefficients of a block are non-zero, bit k = zig-zag position k: kept beside the planes of
  // a progressive frame so that the refinement scans -- which visit every position of the band in every
  // block to ask "is it non-zero?" (:1560-1612, :1650-1677) -- read a register instead of walking the
  // 128-byte block (the luma refinement scan was 10 of the 28 ms a 2048 x 2048 file took to decode).
  // The bit is recomputed from the stored int16 at every store, so it always equals (coefficient != 0).
  std::vector<uint64_t> nzmask[3];
  uint64_t &mask_of(int c, const int16_t *blk) { return nzmask[c][(size_t)(blk - plane[c].base) >> 6]; }

  template <class B> void ac_first(B &bits, int c, int16_t *blk, int ss, int se, int al) {
    if (eobrun > 0) { eobrun--; return; }
    const HuffTable &ac = table(1, ht_ac[c]);
    uint64_t &mask = mask_of(c, blk);
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
        const int value = bits.receive_extend(s) * (1 << al);
        mag[c] |= magnitude(value);
        blk[nat] = (int16_t)value;
        mask = blk[nat] != 0 ? mask | (1ull << k) : mask & ~(1ull << k);
        if ((nat >> 3) > last_row[c]) last_row[c] = nat >> 3;
        k++;
      }
    }
  }
  // Progressive_AC_Scan, refining (:1515-1689), one block's share of it.  The reference's own control flow, so
  // that streams no encoder writes come out as they do there:
  //   * a new coefficient may have any size s (its value is Bin_Twos_Complement of s bits, times 2**al; :1580-1616);
  //   * a zero run is walked without regard to the end of the band (:1561-1573): it may carry the index past
  //     it, and raises Constraint_Error (zag_zig's range) past coefficient 63;
  //   * a correction bit is ADDED to the magnitude of the coefficient, whatever its bits (:1433-1447);
  //   * the coefficients an end-of-band run passes wait in a list of 2000 for their bits (Append, :1402-1415):
  //     a longer run raises Constraint_Error.
  // The order of the reads is the reference's: the walk reads nothing, then the bits of the new value, then one
  // bit per coefficient passed.
  int eob_points = 0;  // coefficients passed by the end-of-band run under way
  template <class B> void ac_refine(B &bits, int c, int16_t *blk, int ss, int se, int al) {
    const int p1 = 1 << al;
    int k = ss;
    uint64_t &mask = mask_of(c, blk);
    auto correct = [&](int kk) {  // Refine_Batch, one point
      if (!bits.get(1)) return;
      int16_t &z = blk[kZigZag[kk]];
      const int v = z > 0 ? z + p1 : z - p1;
      mag[c] |= magnitude(v);
      z = (int16_t)v;
      mask = z != 0 ? mask | (1ull << kk) : mask & ~(1ull << kk);
    };
    if (eobrun <= 0) {
      const HuffTable &ac = table(1, ht_ac[c]);
      while (k <= se) {
        const int rs = bits.symbol(ac);
        const int rr = rs >> 4, s = rs & 15;
        if (s == 0 && rr < 15) {
          eobrun = 1 << rr;
          if (rr) eobrun += bits.get(rr);
          eob_points = 0;
          break;
        }
        uint8_t passed[64];
        int npassed = 0;
        for (int zero_run = s == 0 ? 16 : rr; zero_run > 0; k++) {
          if (k > 63) throw constraint_error("JPEG: progressive: zag_zig index out of range");
          if ((mask >> k) & 1ull) passed[npassed++] = (uint8_t)k;
          else zero_run--;
        }
        if (s != 0) {
          const int value = BitReader::extend(bits.get(s), s) * p1;
          if (k > 63) throw constraint_error("JPEG: progressive: zag_zig index out of range");
          while ((mask >> k) & 1ull) {
            passed[npassed++] = (uint8_t)k;
            if (++k > 63) throw error_in_image_data("JPEG, progressive: zig-zag index overflow");
          }
          mag[c] |= magnitude(value);
          int16_t &z = blk[kZigZag[k]];
          z = (int16_t)value;
          if (z != 0) mask |= 1ull << k;
          if ((kZigZag[k] >> 3) > last_row[c]) last_row[c] = kZigZag[k] >> 3;
          k++;
        }
        for (int i = 0; i < npassed; i++) correct(passed[i]);
      }
    }
    if (eobrun > 0) {
      // the rest of the band: only the non-zero coefficients take a correction bit, in zig-zag order
      if (k <= se) {
        uint64_t m = mask & (~0ull << k) & (se >= 63 ? ~0ull : ((1ull << (se + 1)) - 1ull));
        eob_points += __builtin_popcountll(m);
        if (eob_points > 2000) throw constraint_error("Progressive JPEG: refining buffer capacity 2000 exceeded");
        while (m) {
          const int kk = __builtin_ctzll(m);
          m &= m - 1;
          correct(kk);
        }
        k = se + 1;
      }
      eobrun--;
    }
  }

  // ---- the scans of a progressive frame handed to the device (gidcuda_job_pscan*, include/gidcuda.h) --------
  // First scans and DC refinement scans go there as compressed bytes; AC refinement scans are decoded here
  // from the non-zero history alone (no coefficient values are needed to know which bits a refinement scan
  // reads, :1560-1690) and answered with three bit masks per block.  Decided at the frame's first scan; if a
  // later scan does not fit the device path the whole frame is handed back (device_declined) and decoded here.
  struct device_declined {};
  struct DevicePscan {
    int ncomps, comps[3], ss, se, ah, al;
    const uint8_t *data;
    size_t len;
    const std::vector<uint8_t> *dc_bits[3], *dc_vals[3], *ac_bits[3], *ac_vals[3];
  };
  std::function<bool(size_t)> device_pscan_begin;                 // all entropy bytes of the frame fit in this many
  std::function<bool(const DevicePscan &)> device_pscan;          // false: declined
  std::function<const uint64_t *(int, int32_t *, int32_t *)> device_masks;  // (c, blocks_x, nblocks): waits; throws on error
  std::function<void(int, uint64_t **, uint64_t **, uint64_t **)> device_refine_begin;
  std::function<void(int, int)> device_refine_commit;             // (c, al)
  size_t device_pscan_min_bytes = 96 * 1024;  // smaller files: the launches cost more than the host decoder
  bool device_progressive = false;
  std::vector<uint64_t> hist[3];              // non-zero history per component, indexed by block of the padded plane
  int hist_bx[3] = {0, 0, 0};
  bool hist_valid[3] = {false, false, false};
  std::vector<uint8_t> unstuffed;

  // AC refinement of one block from its non-zero history (the mirror's ac_refine with the coefficient values
  // left out): which positions read a correction bit of 1, which became +-(1 << al).
  template <class B> void ac_refine_history(B &bits, int c, uint64_t &mask, uint64_t &corr, uint64_t &newpos, uint64_t &newneg,
                                            int ss, int se) {
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
            eob_points = 0;
            break;
          }
        } else {
          // (what no encoder writes -- a size other than 1, a run that leaves the band, an end-of-band run over
          // more than 2000 coefficients -- is the host decoder's: it follows the reference's control flow there)
          if (s != 1) throw device_declined();
          value = bits.get(1) ? 1 : -1;
        }
        for (; k <= se; k++) {
          if ((mask >> k) & 1u) {
            if (bits.get(1)) corr |= 1ull << k;
          } else {
            if (rr == 0) {
              if (value) {
                newpos |= 1ull << k;
                if (value < 0) newneg |= 1ull << k;
                mask |= 1ull << k;
              }
              break;
            }
            rr--;
          }
        }
        if (k > se) throw device_declined();  // the run went past the end of the band
      }
    }
    if (eobrun > 0) {
      if (k <= se) {
        uint64_t m = mask & (~0ull << k) & (se >= 63 ? ~0ull : ((1ull << (se + 1)) - 1ull));
        eob_points += __builtin_popcountll(m);
        if (eob_points > 2000) throw device_declined();
        while (m) {
          const int kk = __builtin_ctzll(m);
          m &= m - 1;
          if (bits.get(1)) corr |= 1ull << kk;
        }
      }
      eobrun--;
    }
  }

  // The same with the walks over the band done by bit arithmetic (x86 BMI2: pdep finds the n-th zero position,
  // popcount counts the correction bits, which are then read together and dealt out by pdep).  Same bits read
  // in the same order, same results (tests: gidhost_selftest_refine, and the device path against the oracle).
#if defined(__x86_64__)
  template <class B>
  __attribute__((target("bmi2,popcnt"))) void read_corrections(B &bits, uint64_t m, uint64_t &corr) {
    while (m) {
      const uint64_t chunk = __builtin_ia32_pdep_di(0xFFFFull, m);  // its 16 lowest positions
      const int n = __builtin_popcountll(chunk);
      const uint32_t v = (uint32_t)bits.get_serial(n);             // first bit read = most significant
      uint32_t rev = 0;                                             // ... = lowest position
      for (uint32_t x = v, i = 0; i < (uint32_t)n; i++, x >>= 1) rev = (rev << 1) | (x & 1u);
      corr |= __builtin_ia32_pdep_di((uint64_t)rev, chunk);
      m &= ~chunk;
    }
  }
  template <class B>
  __attribute__((target("bmi2,popcnt"))) void ac_refine_history_bmi2(B &bits, int c, uint64_t &mask, uint64_t &corr,
                                                                     uint64_t &newpos, uint64_t &newneg, int ss, int se) {
    const uint64_t band = (se >= 63 ? ~0ull : ((1ull << (se + 1)) - 1ull)) & (~0ull << ss);
    int k = ss;
    if (eobrun <= 0) {
      const HuffTable &ac = table(1, ht_ac[c]);
      while (k <= se) {
        const int rs = bits.symbol(ac);
        const int rr = rs >> 4, s = rs & 15;
        int value = 0;
        if (s == 0) {
          if (rr < 15) {
            eobrun = 1 << rr;
            if (rr) eobrun += bits.get(rr);
            eob_points = 0;
            break;
          }
        } else {
          if (s != 1) throw device_declined();  // (see ac_refine_history)
          value = bits.get(1) ? 1 : -1;
        }
        // pass rr zero positions and stop at the next one; every non-zero position on the way reads a bit
        const uint64_t from = band & (~0ull << k);
        const uint64_t target = __builtin_ia32_pdep_di(1ull << rr, ~mask & from);
        if (!target) throw device_declined();   // the run went past the end of the band
        const int kt = target ? __builtin_ctzll(target) : se + 1;
        read_corrections(bits, mask & from & (kt > 63 ? ~0ull : ((1ull << kt) - 1ull)), corr);
        if (target && value) {
          newpos |= target;
          if (value < 0) newneg |= target;
          mask |= target;
        }
        k = kt + 1;
      }
    }
    if (eobrun > 0) {
      if (k <= se) {
        const uint64_t m = mask & band & (~0ull << k);
        eob_points += __builtin_popcountll(m);
        if (eob_points > 2000) throw device_declined();
        read_corrections(bits, m, corr);
      }
      eobrun--;
    }
  }
#endif
  template <class B> void ac_refine_history_auto(B &bits, int c, uint64_t &mask, uint64_t &corr, uint64_t &newpos,
                                                 uint64_t &newneg, int ss, int se) {
#if defined(__x86_64__)
    static const bool fast = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    if (fast && !force_portable_refine) {
      ac_refine_history_bmi2(bits, c, mask, corr, newpos, newneg, ss, se);
      return;
    }
#endif
    ac_refine_history(bits, c, mask, corr, newpos, newneg, ss, se);
  }
  bool force_portable_refine = false;
  int8_t coded_al[3][64] = {};    // device path: the bit down to which each coefficient of a component has been coded (-1: not yet), see device_progressive_scan
  bool selftest_refine = false;   // see progressive_scan
  size_t selftest_pos = 0;
  uint8_t selftest_marker = 0;
  int selftest_scans = 0;

  void device_progressive_scan(const int *comps, int ncomps, int ss, int se, int ah, int al) {
    if (im.restart_interval != 0) throw device_declined();
    if (ss > se || se > 63) throw error_in_image_data("JPEG: progressive: bad spectral selection");
    if (ss == 0 && se != 0) throw error_in_image_data("JPEG: progressive: DC and AC in the same scan");
    if (ah != 0 && ah - al != 1)  // :1716-1727
      throw error_in_image_data("JPEG, progressive: precision improvement has to be by one bit.");
    if (ss > 0 && ncomps != 1) throw error_in_image_data("JPEG: progressive: AC scan with several components");
    int ncomp_image = 0;
    for (const Component_Info &ci : im.info) ncomp_image += ci.present ? 1 : 0;
    if (ncomps != 1 && ncomps != ncomp_image) throw device_declined();
    rst_count = 0;
    next_rst = 0;
    eobrun = 0;
    // The device refines by setting one bit (DC) and by the libjpeg rule (AC); the reference ADDS the bit
    // (:1320-1326, :1433-1447).  The two agree when the bit is clear, which it is when every coefficient the
    // scan touches was last coded down to bit al + 1 -- the successive approximation every encoder writes.
    for (int i = 0; i < ncomps; i++)
      for (int k = ss; k <= se; k++) {
        int8_t &p = coded_al[comps[i]][k];
        if (ah > 0 && (ah != al + 1 || p != ah)) throw device_declined();
        p = (int8_t)al;
      }
    const uint8_t *d = r.s.data;
    const size_t size = r.s.size, start = r.s.pos;
    if (ss > 0 && ah > 0) {
      // AC refinement: decoded here from the history, through the wide bit buffer over the raw bytes
      const int c = comps[0];
      const auto t0 = std::chrono::steady_clock::now();
      if (!hist_valid[c]) {
        int32_t bx = 0, nb = 0;
        const uint64_t *m = device_masks(c, &bx, &nb);
        hist[c].assign(m, m + nb);
        hist_bx[c] = bx;
        hist_valid[c] = true;
      }
      const auto t1 = std::chrono::steady_clock::now();
      detail::ns_masks_wait.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count(), std::memory_order_relaxed);
      uint64_t *corr = nullptr, *newpos = nullptr, *newneg = nullptr;
      device_refine_begin(c, &corr, &newpos, &newneg);
      const Component_Info &ci = im.info[c];
      const int wc = (im.width * ci.samples_hor + im.max_samples_hor - 1) / im.max_samples_hor;
      const int hc = (im.height * ci.samples_ver + im.max_samples_ver - 1) / im.max_samples_ver;
      const int bw = (wc + 7) / 8, bh = (hc + 7) / 8;
      WideBits wb(d, start, size);
      wb.track = true;
      try {
        for (int by = 0; by < bh; by++)
          for (int bx = 0; bx < bw; bx++) {
            const size_t pb = (size_t)by * (size_t)hist_bx[c] + (size_t)bx;
            ac_refine_history_auto(wb, c, hist[c][pb], corr[pb], newpos[pb], newneg[pb], ss, se);
          }
      } catch (const gid_error &) {
        throw device_declined();  // whatever the stream raises, the host decoder raises it (or carries on as the reference does)
      }
      if (eobrun > 0) throw device_declined();  // an end-of-band run that reaches past the last block
      device_refine_commit(c, al);
      detail::ns_refine_host.fetch_add(
          std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t1).count(), std::memory_order_relaxed);
      uint8_t marker = 0;
      r.s.pos = wb.reference_position(start, &marker);
      bits.memo_marker = marker;
      return;
    }
    // a first scan or a DC refinement scan: its bytes up to the marker that ends it, stuffing removed
    const auto tq = std::chrono::steady_clock::now();
    unstuffed.clear();
    size_t i = start;
    uint8_t m = 0;
    for (;;) {
      const void *f = i < size ? memchr(d + i, 0xFF, size - i) : nullptr;
      if (!f) throw device_declined();  // no marker ends the scan: a truncated file is the host decoder's case
      const size_t j = (size_t)((const uint8_t *)f - d);
      unstuffed.insert(unstuffed.end(), d + i, d + j);
      if (j + 1 >= size) throw device_declined();
      m = d[j + 1];
      if (m != 0) {
        if (m >= M_RST0 && m <= M_RST7) throw device_declined();
        i = j;
        break;
      }
      unstuffed.push_back(0xFF);
      i = j + 2;
    }
    if (unstuffed.empty()) throw device_declined();
    const bool ok = m == M_EOI || m == M_COM || m == M_DHT || m == M_DRI || m == M_SOS;
    if (!ok) throw device_declined();  // (the serial decoder raises on it: let it)
    DevicePscan ds;
    ds.ncomps = ncomps;
    for (int k = 0; k < 3; k++) ds.comps[k] = k < ncomps ? comps[k] : 0;
    ds.ss = ss; ds.se = se; ds.ah = ah; ds.al = al;
    ds.data = unstuffed.data();
    ds.len = unstuffed.size();
    for (int c = 0; c < 3; c++) {
      ds.dc_bits[c] = ds.dc_vals[c] = ds.ac_bits[c] = ds.ac_vals[c] = nullptr;
      bool used = false;
      for (int k = 0; k < ncomps; k++) used = used || comps[k] == c;
      if (!used || !im.info[c].present) continue;
      if (ss == 0 && ah == 0) {
        table(0, ht_dc[c]);  // raises like the serial loop when the table is missing
        ds.dc_bits[c] = &im.huff_bits[0][ht_dc[c]];
        ds.dc_vals[c] = &im.huff_vals[0][ht_dc[c]];
      } else if (ss > 0) {
        table(1, ht_ac[c]);
        ds.ac_bits[c] = &im.huff_bits[1][ht_ac[c]];
        ds.ac_vals[c] = &im.huff_vals[1][ht_ac[c]];
      }
    }
    if (!device_pscan(ds)) throw device_declined();
    detail::ns_pscan_queue.fetch_add(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - tq).count(), std::memory_order_relaxed);
    if (ss > 0) hist_valid[comps[0]] = false;  // new non-zero coefficients: the history of the component is the device's
    r.s.pos = i + 2;                           // past the marker that ended the scan, which the segment loop takes up
    bits.memo_marker = m;
    device_scanned = true;
    emit_feedback(std::min(49, 5 * (scans + 1)));
  }

  void progressive_scan(const int *comps, int ncomps, int ss, int se, int ah, int al) {
    if (scans == 0) {
      memset(coded_al, -1, sizeof coded_al);
      // the whole frame on the device, or none of it: decided here, at its first scan
      device_progressive = device_pscan && device_pscan_begin && im.restart_interval == 0 &&
                           r.s.size >= device_pscan_min_bytes && device_pscan_begin(r.s.size - r.s.pos);
    }
    if (device_progressive) {
      device_progressive_scan(comps, ncomps, ss, se, ah, al);
      return;
    }
    // Without restart markers the scan is read through the 64-bit buffer; afterwards the stream
    // position and the remembered marker are set to what the byte-wise buffer would have left
    // behind, so that the segment loop (and its errors on damaged files) carry on identically.
    if (wide_progressive && im.restart_interval == 0) {
      const size_t start = r.s.pos;
      // Self-test of the device path's host half (no device needed): decode the AC refinement scan a first
      // time from the non-zero history alone, apply its verdict to a copy of the plane the way
      // pscan_apply_kernel does, and hold the copy against what the decoder proper makes of the scan.
      std::vector<int16_t> shadow;
      int shadow_c = -1;
      if (selftest_refine && ss > 0 && ah > 0 && ncomps == 1 && plane[comps[0]].base) {
        const int c = comps[0];
        const Component_Info &ci = im.info[c];
        const size_t nb = (size_t)plane[c].blocks_x * (size_t)plane[c].blocks_y;
        if (nzmask[c].size() != nb) nzmask[c].assign(nb, 0);
        shadow.assign(plane[c].base, plane[c].base + nb * 64);
        std::vector<uint64_t> h(nzmask[c]), corr(nb, 0), np(nb, 0), nn(nb, 0);
        const int wc = (im.width * ci.samples_hor + im.max_samples_hor - 1) / im.max_samples_hor;
        const int hc = (im.height * ci.samples_ver + im.max_samples_ver - 1) / im.max_samples_ver;
        const int bw = (wc + 7) / 8, bh = (hc + 7) / 8;
        WideBits w2(r.s.data, start, r.s.size);
        w2.track = true;
        eobrun = 0;
        bool usual = ah == al + 1;
        try {
          for (int by = 0; by < bh && usual; by++)
            for (int bx = 0; bx < bw; bx++) {
              const size_t pb = (size_t)by * (size_t)plane[c].blocks_x + (size_t)bx;
              ac_refine_history_auto(w2, c, h[pb], corr[pb], np[pb], nn[pb], ss, se);
            }
          if (eobrun > 0) usual = false;
        } catch (const gid_error &) {
          usual = false;   // (a damaged scan: the decoder proper below raises what there is to raise)
        } catch (const device_declined &) {
          usual = false;   // (a scan the device path hands to the host decoder)
        }
        const int p1 = 1 << al, m1 = -(1 << al);
        for (size_t b = 0; b < nb; b++) {
          int16_t *blk = shadow.data() + b * 64;
          for (uint64_t m = corr[b]; m; m &= m - 1) {
            int16_t &z = blk[kZigZag[__builtin_ctzll(m)]];
            if ((z & p1) == 0) z = (int16_t)(z >= 0 ? z + p1 : z + m1);
          }
          for (uint64_t m = np[b]; m; m &= m - 1) {
            const int k = __builtin_ctzll(m);
            blk[kZigZag[k]] = (int16_t)(((nn[b] >> k) & 1ull) ? m1 : p1);
          }
        }
        selftest_pos = w2.reference_position(start, &selftest_marker);
        shadow_c = usual ? c : -1;
      }
      WideBits wb(r.s.data, start, r.s.size);
      wb.track = true;
      progressive_scan_with(wb, comps, ncomps, ss, se, ah, al);
      if (shadow_c >= 0) {
        const size_t nb = (size_t)plane[shadow_c].blocks_x * (size_t)plane[shadow_c].blocks_y;
        uint8_t mk = 0;
        const size_t pos = wb.reference_position(start, &mk);
        if (memcmp(shadow.data(), plane[shadow_c].base, nb * 64 * sizeof(int16_t)) != 0 || pos != selftest_pos || mk != selftest_marker)
          throw constraint_error("self-test: the AC refinement scan decoded from the non-zero history differs from the decoder's");
        selftest_scans++;
      }
      uint8_t marker = 0;
      r.s.pos = wb.reference_position(start, &marker);
      bits.memo_marker = marker;
      return;
    }
    progressive_scan_with(bits, comps, ncomps, ss, se, ah, al);
  }

  template <class B>
  void progressive_scan_with(B &bits, const int *comps, int ncomps, int ss, int se, int ah, int al) {
    if (ss > se || se > 63) throw error_in_image_data("JPEG: progressive: bad spectral selection");
    if (ss == 0 && se != 0) throw error_in_image_data("JPEG: progressive: DC and AC in the same scan");
    if (ah != 0 && ah - al != 1)  // :1716-1727
      throw error_in_image_data("JPEG, progressive: precision improvement has to be by one bit.");
    if (ss > 0 && ncomps != 1) throw error_in_image_data("JPEG: progressive: AC scan with several components");
    ensure_host_output();
    for (int i = 0; i < ncomps; i++) {
      const int c = comps[i];
      const size_t nb = (size_t)plane[c].blocks_x * (size_t)plane[c].blocks_y;
      if (nzmask[c].size() != nb) nzmask[c].assign(nb, 0);
      if (ss == 0 && ah > 0) mag[c] |= 1u << al;  // DC refinement ORs bit al into the DC terms
    }
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
                if (ah == 0) dc_first(bits, c, blk, al);
                else dc_refine(bits, c, blk, al);
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
          if (ah == 0) dc_first(bits, c, blk, al);
          else dc_refine(bits, c, blk, al);
        } else if (ah == 0) {
          ac_first(bits, c, blk, ss, se, al);
        } else {
          ac_refine(bits, c, blk, ss, se, al);
        }
        if (!(by == bh - 1 && bx == bw - 1)) check_restart(true);
      }
      if (first_scan) emit_feedback((50 * (by + 1)) / bh);
    }
    // An end-of-band run of a refining scan walks on through the blocks it covers as soon as it is read
    // (:1650-1675): past the component's last block that is rows of the image array no scan has written
    // (8 * max_samples_ver * MCU rows high for every component, Create_Image_Array :1857-1883) -- nothing to
    // refine there -- and past the array's last row Constraint_Error.
    if (ss > 0 && ah > 0 && eobrun > 0) {
      const long rows = (long)mcu_count_image_v * im.max_samples_ver;
      if ((long)eobrun > (rows - bh) * (long)bw) throw constraint_error("JPEG: progressive: image_array index out of range");
    }
  }

  // Read_SOS (:1855-2034): scan header, then the scan's entropy-coded data.
  void read_sos(const Segment &) {
    const int n = r.byte();
    int ncomp_image = 0;
    for (const Component_Info &ci : im.info) ncomp_image += ci.present ? 1 : 0;
    if (n > ncomp_image) throw error_in_image_data("JPEG: Scan segment has more color components than the image");
    if (n < 1) throw error_in_image_data("JPEG: Scan segment without components");
    // (the segment's length field is not looked at: Read_SOS reads what it needs, :1895-1940)
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
    if (!frame_begun) {
      begin_frame();
      frame_begun = true;
    }
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

}  // namespace gid
