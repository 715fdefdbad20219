// depth_image.cpp -- deplex::utils::DepthImage: PNG -> 16-bit grey samples -> organized cloud.
//
// Reference: cpp/deplex/src/deplex/utils/depth_image.cpp:30-78, which decodes through the vendored
// stb_image (`stbi_load_16(path, &w, &h, &channels, STBI_grey)`).  This reader is written against the PNG
// specification (RFC 2083) over zlib's inflate and produces what that call produces for the files deplex
// is used with, in stb's order of operations: colour is reduced to luma as (77 r + 150 g + 29 b) >> 8 at the file's own
// sample width FIRST, and an 8-bit result is then widened as v * 257; alpha is dropped; grey samples of 1, 2 and 4 bits
// are scaled to 8 bits (x 255, 85, 17) and palette images are expanded through PLTE before that.  Interlaced (Adam7)
// PNGs are rejected with the reference's "Couldn't read image" error (stb decodes them; depth maps are never interlaced).
#include "deplex/utils/depth_image.h"

#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace deplex {
namespace utils {
namespace {

uint32_t be32(const unsigned char* p) {
  return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) | p[3];
}

int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  if (pa <= pb && pa <= pc) return a;
  return pb <= pc ? b : c;
}

// Decodes `path` into 16-bit grey.  Returns false on any failure (stb reports failures as a null pointer).
bool decode_png_grey16(std::string const& path, std::vector<uint16_t>* out, int32_t* width, int32_t* height) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) return false;
  std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  static const unsigned char kSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 8 + 25 || std::memcmp(file.data(), kSig, 8) != 0) return false;

  size_t pos = 8;
  uint32_t w = 0, h = 0;
  int depth = 0, color = -1, interlace = 0;
  std::vector<unsigned char> idat, plte;
  bool seen_end = false;
  while (!seen_end && pos + 12 <= file.size()) {
    const uint32_t len = be32(&file[pos]);
    const unsigned char* type = &file[pos + 4];
    if (pos + 12 + static_cast<size_t>(len) > file.size()) return false;
    const unsigned char* body = &file[pos + 8];
    if (std::memcmp(type, "IHDR", 4) == 0) {
      if (len != 13) return false;
      w = be32(body);
      h = be32(body + 4);
      depth = body[8];
      color = body[9];
      if (body[10] != 0 || body[11] != 0) return false;
      interlace = body[12];
    } else if (std::memcmp(type, "PLTE", 4) == 0) {
      plte.assign(body, body + len);
    } else if (std::memcmp(type, "IDAT", 4) == 0) {
      idat.insert(idat.end(), body, body + len);
    } else if (std::memcmp(type, "IEND", 4) == 0) {
      seen_end = true;
    }
    pos += 12 + static_cast<size_t>(len);
  }
  if (w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24) || idat.empty()) return false;
  if (interlace != 0) return false;
  int channels;
  switch (color) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;  // palette indices
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return false;
  }
  // legal sample widths (RFC 2083 section 3.1.1)
  const bool sub_byte = depth == 1 || depth == 2 || depth == 4;
  if (!(depth == 8 || (depth == 16 && color != 3) || (sub_byte && (color == 0 || color == 3)))) return false;
  if (color == 3 && (plte.empty() || plte.size() % 3 != 0)) return false;
  const size_t bpp = sub_byte ? 1 : static_cast<size_t>(channels) * (depth / 8);  // filter distance in bytes
  const size_t stride = (static_cast<size_t>(w) * channels * depth + 7) / 8;
  std::vector<unsigned char> raw((stride + 1) * h);
  uLongf raw_len = static_cast<uLongf>(raw.size());
  if (uncompress(raw.data(), &raw_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK) return false;
  if (raw_len != raw.size()) return false;

  // undo the per-scanline filters in place (RFC 2083 section 6)
  std::vector<unsigned char> img(stride * h);
  for (uint32_t y = 0; y < h; ++y) {
    const unsigned char* src = &raw[(stride + 1) * y];
    const int filter = src[0];
    ++src;
    unsigned char* cur = &img[stride * y];
    const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= bpp ? cur[i - bpp] : 0;
      const int b = up ? up[i] : 0;
      const int c = (up && i >= bpp) ? up[i - bpp] : 0;
      int v = src[i];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return false;
      }
      cur[i] = static_cast<unsigned char>(v);
    }
  }

  out->resize(static_cast<size_t>(w) * h);
  const uint32_t scale = depth == 1 ? 0xffu : depth == 2 ? 0x55u : depth == 4 ? 0x11u : 1u;  // stb's depth_scale_table
  for (uint32_t y = 0; y < h; ++y) {
    const unsigned char* row = &img[stride * y];
    for (uint32_t x = 0; x < w; ++x) {
      uint32_t ch[4] = {0, 0, 0, 0};
      int n_ch = channels, bits = depth;
      if (depth < 8 || color == 3) {
        uint32_t v = row[x];
        if (depth < 8) {
          const size_t bit = static_cast<size_t>(x) * depth;
          v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
        }
        bits = 8;
        if (color == 3) {
          if (static_cast<size_t>(v) * 3 + 2 >= plte.size()) return false;
          ch[0] = plte[3 * v]; ch[1] = plte[3 * v + 1]; ch[2] = plte[3 * v + 2];
          n_ch = 3;
        } else {
          ch[0] = v * scale;
        }
      } else {
        for (int c = 0; c < channels; ++c) {
          const unsigned char* s = row + static_cast<size_t>(x) * channels * (depth / 8) + static_cast<size_t>(c) * (depth / 8);
          ch[c] = depth == 16 ? static_cast<uint32_t>((s[0] << 8) | s[1]) : static_cast<uint32_t>(s[0]);
        }
      }
      // stb: reduce to grey at the file's sample width (stbi__compute_y / _16), then widen 8 -> 16 bits as v * 257
      uint32_t grey = ch[0];
      if (n_ch >= 3) grey = (ch[0] * 77u + ch[1] * 150u + ch[2] * 29u) >> 8;
      if (bits == 8) grey = (grey & 0xffu) * 257u;
      (*out)[static_cast<size_t>(y) * w + x] = static_cast<uint16_t>(grey);
    }
  }
  *width = static_cast<int32_t>(w);
  *height = static_cast<int32_t>(h);
  return true;
}

}  // namespace

DepthImage::DepthImage() : width_(0), height_(0) {}

DepthImage::DepthImage(std::string const& image_path) : width_(0), height_(0) { reset(image_path); }

void DepthImage::reset(std::string const& image_path) {
  std::vector<uint16_t> img;
  int32_t w = 0, h = 0;
  if (!decode_png_grey16(image_path, &img, &w, &h)) {
    throw std::runtime_error("Error: Couldn't read image " + image_path);
  }
  image_.swap(img);
  width_ = w;
  height_ = h;
}

int32_t DepthImage::getWidth() const { return width_; }
int32_t DepthImage::getHeight() const { return height_; }
uint16_t const* DepthImage::data() const { return image_.data(); }

std::vector<float> DepthImage::toPointCloudRowMajor(Intrinsics const& k) const {
  const float fx = k[0], cx = k[2], fy = k[4], cy = k[5];
  std::vector<float> pts(static_cast<size_t>(width_) * height_ * 3);
  size_t i = 0;
  for (int32_t r = 0; r < height_; ++r) {
    const float fr = static_cast<float>(r);
    for (int32_t c = 0; c < width_; ++c, ++i) {
      const float z = static_cast<float>(image_[i]);
      // depth_image.cpp:70-73: (indices - c) * z / f, evaluated left to right, each step rounded to fp32
      volatile float tx = (static_cast<float>(c) - cx) * z;
      volatile float ty = (fr - cy) * z;
      pts[3 * i + 0] = tx / fx;
      pts[3 * i + 1] = ty / fy;
      pts[3 * i + 2] = z;
    }
  }
  return pts;
}

}  // namespace utils
}  // namespace deplex
