// gid.hpp -- C++ host mirror of GID's public API for the JPEG path.
//
// GID is an Ada library; no Ada compiler exists in this repository's build image,
// so this header + gid_jpeg_host.cpp are the host side that is actually built and
// tested above the C ABI (include/gidcuda.h).  Names, argument meaning and error
// behaviour follow the reference (file:line relative to the GID 013 tree):
//
//   gid::Load_Image_Header      GID.Load_Image_Header              gid.ads:64-67
//   gid::Load_Image_Contents    generic GID.Load_Image_Contents    gid.ads:117-143
//       Primary_Color_Range     type ... is mod <>                 gid.ads:118   (2**8 or 2**16 only, :920-937)
//       set_x_y (x, y)          procedure Set_X_Y                  gid.ads:125   (y = 0 is the BOTTOM row)
//       put_pixel (r, g, b, a)  procedure Put_Pixel                gid.ads:131
//       feedback (percents)     procedure Feedback                 gid.ads:137
//   gid::Pixel_Width / Pixel_Height / Display_Orientation / ...    gid.ads:85-94, :158-171
//   exceptions                                                      gid.ads:73-77
//
// What differs from the reference is only WHERE the pixel reconstruction runs:
// the entropy decoder (baseline + progressive, in gid_jpeg_host.cpp) stores
// un-dequantised int16 coefficients into the planes lent by gidcuda_job_plane,
// the B200 kernel does dequant + IDCT + upsampling + colour, and the rows that
// come back are replayed through set_x_y / put_pixel exactly as the patched
// GID.Decoding_JPG would (gid_b200/ada/gid-decoding_jpg-cuda_patch.ada).
// There is no CPU fallback: without a CUDA device Load_Image_Contents raises
// gid::cuda_device_error.
#pragma once

#include <cstddef>
#include <cstdint>
#include <atomic>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

struct gidcuda_frame_desc;

namespace gid {

// ---- exceptions (gid.ads:73-77 + the predefined ones the JPEG path can raise) ----
struct gid_error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct unknown_image_format : gid_error { using gid_error::gid_error; };
struct known_but_unsupported_image_format : gid_error { using gid_error::gid_error; };
struct unsupported_image_subformat : gid_error { using gid_error::gid_error; };
struct error_in_image_data : gid_error { using gid_error::gid_error; };
struct invalid_primary_color_range : gid_error { using gid_error::gid_error; };
struct constraint_error : gid_error { using gid_error::gid_error; };   // Ada Constraint_Error (:1408)
struct cuda_device_error : gid_error { using gid_error::gid_error; };  // GID.CUDA.Device_Error

enum class Image_Format_Type { BMP, FITS, GIF, JPEG, PNG, PNM, QOI, RIFF, TGA, TIFF };  // gid.ads; only JPEG is decoded here
enum class Orientation { Unchanged, Rotation_90, Rotation_180, Rotation_270 };  // gid.ads:88-92
enum class Display_Mode { fast, redundant };                                       // gid.ads (unused by JPEG)

// Input: GID reads an Ada stream through GID.Buffering; here, a memory stream.
struct Stream {
  const uint8_t *data = nullptr;
  size_t size = 0, pos = 0;
  Stream() = default;
  Stream(const uint8_t *d, size_t n) : data(d), size(n) {}
};

// One quantisation table as read (zig-zag order, gid.ads:271-272)
struct Quantization_Table { uint16_t q[64]; bool defined = false; };

struct Component_Info {  // JPEG_Defs.Info_per_Component_A, gid.ads:285-296
  bool present = false;
  int samples_hor = 0, samples_ver = 0, up_factor_x = 0, up_factor_y = 0, qt_assoc = 0;
};

// Image_Descriptor (gid.ads:339-372): filled by Load_Image_Header, consumed by
// Load_Image_Contents.  Only the JPEG fields exist here.
struct Image_Descriptor {
  Image_Format_Type format = Image_Format_Type::JPEG;
  std::string detailed_format;
  int subformat_id = 0;  // number of components
  int width = 0, height = 0;
  int bits_per_pixel = 0;
  bool greyscale = false, progressive = false;
  Orientation display_orientation = Orientation::Unchanged;
  // JPEG_stuff (gid.ads:318-333)
  Component_Info info[3];  // Y, Cb, Cr
  int max_samples_hor = 0, max_samples_ver = 0;
  Quantization_Table qt_list[8];
  int restart_interval = 0;
  struct Huffman;  // 65 536-entry direct tables in the reference (gid.ads:316); canonical tables here
  std::vector<uint8_t> huff_bits[2][8], huff_vals[2][8];  // [DC=0 / AC=1][table id]: counts (16) and symbols
  // the stream the header was read from (GID keeps it attached, gid.ads:345)
  Stream *stream = nullptr;
  int device = 0;  // CUDA device used by Load_Image_Contents (not in the reference)
  // Coefficients reach the device as sparse records (gidcuda_job_records: one record per
  // non-zero coefficient, the entropy decoder's natural output) unless this is false: then, or
  // for a component too dense for the record buffer, as dense int16 planes.
  bool sparse_input = true;
  // Host threads for the entropy decode of ONE image (not in the reference, which is serial): a
  // baseline scan with restart intervals (DRI) is decoded interval-parallel when > 1
  // (0 = one per hardware thread).  Scans without restart markers, and progressive scans,
  // stay serial.  The decoded coefficients are the same either way.
  int entropy_threads = 1;
  // Baseline scans with restart intervals (at least 64 of them) are entropy-decoded ON THE DEVICE,
  // one thread per interval (gidcuda_job_scan): only the compressed bytes cross the host link.
  // Streams the device decoder declines (damaged or unusual ones) are decoded on the host as
  // before, so results and error behaviour do not depend on this switch.
  bool device_entropy = true;
  // ... and so are the first scans and the DC refinement scans of PROGRESSIVE files of at least this many
  // bytes (gidcuda_job_pscan*); their AC refinement scans are decoded on the host from the non-zero history
  // the device hands back.  0 bytes: every progressive file; a negative value: none.
  long device_progressive_min_bytes = 96 * 1024;
  // Progressive scans (without restart markers) are read through the 64-bit bit buffer; false =
  // through the byte-wise one that mirrors the reference's.  Same planes, same stream position
  // afterwards, same exceptions (tests compare the two on damaged files).
  bool wide_bit_buffer = true;
  // The file batch driver only (Load_Image_Contents leaves the orientation to the client's Set_X_Y, as
  // GID does): write the pixels turned as Display_Orientation says (GIDCUDA_FLAG_ROTATE_*), so that
  // the caller's buffer holds the image as it is to be displayed -- H x W for Rotation_90 / _270.
  bool apply_orientation = false;
  size_t stream_pos_after_header = 0;  // where Load_Image_Contents starts reading
};

// gid.ads:64-67.  try_tga is accepted for signature compatibility and ignored.
void Load_Image_Header(Image_Descriptor &image, Stream &from, bool try_tga = false);

inline int Pixel_Width(const Image_Descriptor &i) { return i.width; }                       // gid.ads:85
inline int Pixel_Height(const Image_Descriptor &i) { return i.height; }                     // gid.ads:86
inline Orientation Display_Orientation(const Image_Descriptor &i) { return i.display_orientation; }  // :94
inline Image_Format_Type Format(const Image_Descriptor &i) { return i.format; }             // gid.ads:158
inline const std::string &Detailed_Format(const Image_Descriptor &i) { return i.detailed_format; }
inline int Subformat(const Image_Descriptor &i) { return i.subformat_id; }
inline int Bits_per_Pixel(const Image_Descriptor &i) { return i.bits_per_pixel; }
inline bool Is_RLE_Encoded(const Image_Descriptor &) { return false; }
inline bool Is_Progressive(const Image_Descriptor &i) { return i.progressive; }
inline bool Is_Interlaced(const Image_Descriptor &i) { return i.progressive; }
inline bool Greyscale(const Image_Descriptor &i) { return i.greyscale; }
inline bool Has_Palette(const Image_Descriptor &) { return false; }
inline bool Expect_Transparency(const Image_Descriptor &) { return false; }

namespace detail {

// How many baseline scans were decoded restart-interval-parallel / fell back to the serial loop
// after the parallel attempt met something unusual (process-wide; tests and statistics).
extern std::atomic<long> parallel_scans, parallel_fallbacks;
// progressive frames on the device: host time (ns) spent waiting for the non-zero history / decoding AC
// refinement scans / queueing first scans (gidhost_counter "ns_masks_wait", "ns_refine_host", "ns_pscan_queue")
extern std::atomic<long> ns_masks_wait, ns_refine_host, ns_pscan_queue, selftest_refine_scans;
extern int g_selftest_refine;

// Result of the device reconstruction: top-down RGB8 rows in pinned host memory
// owned by the library until release().
struct Reconstruction {
  const uint8_t *pixels = nullptr;
  size_t stride = 0;
  void *job = nullptr, *ctx = nullptr;
};

// Entropy-decodes every scan of the image into the job's int16 planes, then
// submits the job and waits.  feedback receives 0 .. 50 during the entropy pass.
Reconstruction decode_and_reconstruct(Image_Descriptor &image, const std::function<void(int)> &feedback);
void release(Reconstruction &r);

// The two halves of the above for callers that keep several frames in flight on one context
// (gid_batch.cpp): entropy-decode into a job of `ctx` and submit it (asynchronous; returns the
// gidcuda_job), later gidcuda_job_wait / gidcuda_job_release.
// out (optional): the caller's top-down packed RGB8 buffer; when it is pinned host memory the
// pixels land there directly (gidcuda_job_output) and gidcuda_job_wait returns that address.
void *decode_and_submit(Image_Descriptor &image, void *ctx, const std::function<void(int)> &feedback,
                        uint8_t *out = nullptr, size_t out_capacity = 0);
// gidcuda_job_wait for a job of decode_and_submit.  When the device entropy decoder left the scan
// to the host (GIDCUDA_ERR_DATA), the image is decoded again on the host cores into a new job
// (*job is replaced) and that one is waited for.  Raises GID's exceptions.
const uint8_t *wait_pixels(Image_Descriptor &image, void *ctx, void **job, size_t *stride, uint8_t *out = nullptr,
                           size_t out_capacity = 0);

// statistics (process-wide): scans decoded on the device / handed back to the host by it
extern std::atomic<long> device_scans, device_scan_fallbacks;

// Entropy decode only, into calloc'd planes (see gidhost_decode_planes).
void decode_planes(Image_Descriptor &image, struct ::gidcuda_frame_desc *desc, int16_t *planes[3], int32_t blocks_x[3],
                   int32_t blocks_y[3]);

}  // namespace detail

// gid.ads:117-143.  Primary_Color_Range: uint8_t (values 0 .. 255) or uint16_t
// (values scaled by 257, Out_Pixel_8 gid-decoding_jpg.adb:909-938); any other
// modulus raises invalid_primary_color_range, as the reference does at run time.
template <class Primary_Color_Range, class Set_X_Y, class Put_Pixel, class Feedback>
void Load_Image_Contents(Image_Descriptor &image, double &next_frame, Set_X_Y set_x_y, Put_Pixel put_pixel,
                         Feedback feedback, Display_Mode /*mode*/ = Display_Mode::fast) {
  static_assert(std::is_unsigned<Primary_Color_Range>::value, "Primary_Color_Range is a modular type");
  if (sizeof(Primary_Color_Range) != 1 && sizeof(Primary_Color_Range) != 2)
    throw invalid_primary_color_range("JPEG: color range not supported");
  detail::Reconstruction rec = detail::decode_and_reconstruct(image, [&](int p) { feedback(p); });
  try {
    const int W = image.width, H = image.height;
    const Primary_Color_Range alpha = static_cast<Primary_Color_Range>(~static_cast<Primary_Color_Range>(0));
    for (int row = 0; row < H; row++) {
      set_x_y(0, H - 1 - row);  // y counts from the bottom (gid-decoding_jpg.adb:1024)
      const uint8_t *p = rec.pixels + (size_t)row * rec.stride;
      for (int x = 0; x < W; x++, p += 3) {
        if (sizeof(Primary_Color_Range) == 1)
          put_pixel((Primary_Color_Range)p[0], (Primary_Color_Range)p[1], (Primary_Color_Range)p[2], alpha);
        else
          put_pixel((Primary_Color_Range)(p[0] * 257), (Primary_Color_Range)(p[1] * 257),
                    (Primary_Color_Range)(p[2] * 257), alpha);
      }
      feedback(50 + (50 * (row + 1)) / H);
    }
  } catch (...) {
    detail::release(rec);
    throw;
  }
  detail::release(rec);
  next_frame = 0.0;  // gid.adb:149: JPEG has a single frame
}

}  // namespace gid
