// io_formats.cpp -- decoders of the per-frame input files of the reference's driver.
//
// example/vdo_slam.cc reads, per frame (:98-141, file names :186-194):
//   image_0/NNNNNN.png   8-bit colour or grey PNG        cv::imread(..., CV_LOAD_IMAGE_UNCHANGED)           (:105)
//   depth/NNNNNN.png     16-bit grey PNG (disparity*256) cv::imread UNCHANGED, then convertTo(CV_32F)       (:106-110)
//   flow/NNNNNN.flo      Middlebury .flo                 cv::optflow::readOpticalFlow -> CV_32FC2            (:117)
//   semantic/NNNNNN.txt  one text row of ints per image row, written into a CV_32SC1 (LoadMask, :253-450; only the
//                        non-zero entries are stored by the reference -- the zeros are whatever the fresh cv::Mat held;
//                        here they are 0)
// OpenCV's decoders are not in the reference tree; the PNG path below is the published format (RFC 2083: chunks, zlib
// stream, the five scan-line filters, big-endian 16-bit samples) on top of zlib's inflate, the .flo path the Middlebury
// layout ("PIEH" float tag 202021.25, int32 width, int32 height, then row-major (u, v) float pairs).  Pixel order follows
// cv::imread: BGR / BGRA for colour.  Host-only; tests/test_io_formats.py pins every decoder against cv2 4.13.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/vdo_b200.h"

namespace {

bool read_file(const char* path, std::vector<unsigned char>& buf) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 0 || (unsigned long)sz > ((unsigned long)1 << 31)) { std::fclose(f); return false; }    // directories, pipes, > 2 GiB: not an input file of this path
  try { buf.resize((size_t)sz); } catch (...) { std::fclose(f); return false; }
  const size_t got = sz ? std::fread(buf.data(), 1, (size_t)sz, f) : 0;
  std::fclose(f);
  return got == (size_t)sz;
}
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct PngHeader { int w = 0, h = 0, depth = 0, color = 0, interlace = 0, channels = 0; };

// parses the chunk list; on success idat holds the concatenated IDAT payload
int png_parse(const std::vector<unsigned char>& f, PngHeader& H, std::vector<unsigned char>* idat) {
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (f.size() < 33 || std::memcmp(f.data(), sig, 8) != 0) return VDO_ERR_ARG;
  size_t p = 8;
  bool have_hdr = false;
  while (p + 12 <= f.size()) {
    const uint32_t len = be32(&f[p]);
    const unsigned char* type = &f[p + 4];
    if (p + 12 + (size_t)len > f.size()) return VDO_ERR_ARG;
    const unsigned char* data = &f[p + 8];
    if (std::memcmp(type, "IHDR", 4) == 0) {
      if (len < 13) return VDO_ERR_ARG;
      H.w = (int)be32(data); H.h = (int)be32(data + 4); H.depth = data[8]; H.color = data[9]; H.interlace = data[12];
      if (data[10] != 0 || data[11] != 0) return VDO_ERR_UNSUPPORTED;
      H.channels = H.color == 0 ? 1 : H.color == 2 ? 3 : H.color == 4 ? 2 : H.color == 6 ? 4 : 0;
      have_hdr = true;
    } else if (std::memcmp(type, "IDAT", 4) == 0) {
      if (idat) idat->insert(idat->end(), data, data + len);
    } else if (std::memcmp(type, "IEND", 4) == 0) {
      break;
    }
    p += 12 + (size_t)len;
  }
  if (!have_hdr || H.w <= 0 || H.h <= 0) return VDO_ERR_ARG;
  if (H.channels == 0 || H.color == 4 || H.interlace != 0 || (H.depth != 8 && H.depth != 16)) return VDO_ERR_UNSUPPORTED;   // palette, grey+alpha, Adam7, < 8 bit
  return VDO_OK;
}

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// decodes into `out` (h rows of w*channels samples, samples in file order RGB[A], 16-bit big-endian as in the file)
int png_decode(const char* path, PngHeader& H, std::vector<unsigned char>& out) {
  std::vector<unsigned char> file, idat;
  if (!read_file(path, file)) return VDO_ERR_ARG;
  int rc = png_parse(file, H, &idat);
  if (rc != VDO_OK) return rc;
  const size_t bpp = (size_t)H.channels * H.depth / 8, row = (size_t)H.w * bpp;
  // the header's size is untrusted: deflate expands at most ~1032x, so an image that claims more bytes than the compressed stream could
  // possibly hold (or an absurd size) is rejected before anything is allocated
  if ((size_t)H.w > 65536 || (size_t)H.h > 65536 || (row + 1) * (size_t)H.h > idat.size() * 1100 + 4096) return VDO_ERR_ARG;
  std::vector<unsigned char> raw;
  try { raw.resize((row + 1) * (size_t)H.h); out.reserve(row * (size_t)H.h); } catch (...) { return VDO_ERR_ARG; }
  uLongf n = (uLongf)raw.size();
  if (uncompress(raw.data(), &n, idat.data(), (uLong)idat.size()) != Z_OK || n != raw.size()) return VDO_ERR_ARG;
  out.assign(row * (size_t)H.h, 0);
  for (int y = 0; y < H.h; ++y) {
    const unsigned char* in = &raw[(row + 1) * (size_t)y];
    const int ft = in[0];
    ++in;
    unsigned char* cur = &out[row * (size_t)y];
    const unsigned char* up = y ? cur - row : nullptr;
    for (size_t x = 0; x < row; ++x) {
      const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
      int v = in[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return VDO_ERR_ARG;
      }
      cur[x] = (unsigned char)v;
    }
  }
  return VDO_OK;
}

}  // namespace

extern "C" {

int vdo_io_png_info(const char* path, int* w, int* h, int* channels, int* bit_depth) {
  if (!path) return VDO_ERR_ARG;
  std::vector<unsigned char> file;
  if (!read_file(path, file)) return VDO_ERR_ARG;
  PngHeader H;
  const int rc = png_parse(file, H, nullptr);
  if (rc != VDO_OK) return rc;
  if (w) *w = H.w;
  if (h) *h = H.h;
  if (channels) *channels = H.channels;
  if (bit_depth) *bit_depth = H.depth;
  return VDO_OK;
}

// dst: h x w x channels samples (uint8 or host-endian uint16), colour in OpenCV order (BGR / BGRA) like cv::imread UNCHANGED
int vdo_io_read_png(const char* path, void* dst, size_t dst_bytes) {
  if (!path || !dst) return VDO_ERR_ARG;
  PngHeader H;
  std::vector<unsigned char> px;
  const int rc = png_decode(path, H, px);
  if (rc != VDO_OK) return rc;
  if (dst_bytes < px.size()) return VDO_ERR_ARG;
  const size_t n = (size_t)H.w * H.h;
  const int ch = H.channels;
  if (H.depth == 8) {
    unsigned char* o = (unsigned char*)dst;
    if (ch == 1) std::memcpy(o, px.data(), n);
    else
      for (size_t i = 0; i < n; ++i) {
        o[ch * i] = px[ch * i + 2]; o[ch * i + 1] = px[ch * i + 1]; o[ch * i + 2] = px[ch * i];
        if (ch == 4) o[4 * i + 3] = px[4 * i + 3];
      }
  } else {
    uint16_t* o = (uint16_t*)dst;
    auto s16 = [&](size_t k) { return (uint16_t)((px[2 * k] << 8) | px[2 * k + 1]); };
    if (ch == 1) for (size_t i = 0; i < n; ++i) o[i] = s16(i);
    else
      for (size_t i = 0; i < n; ++i) {
        o[ch * i] = s16(ch * i + 2); o[ch * i + 1] = s16(ch * i + 1); o[ch * i + 2] = s16(ch * i);
        if (ch == 4) o[4 * i + 3] = s16(4 * i + 3);
      }
  }
  return VDO_OK;
}

// single-channel PNG (8 or 16 bit) -> float, the example's `imD.convertTo(imD_f, CV_32F)` (example/vdo_slam.cc:110)
int vdo_io_read_png_gray_f32(const char* path, float* dst, int w, int h) {
  if (!path || !dst) return VDO_ERR_ARG;
  PngHeader H;
  std::vector<unsigned char> px;
  const int rc = png_decode(path, H, px);
  if (rc != VDO_OK) return rc;
  if (H.channels != 1 || H.w != w || H.h != h) return VDO_ERR_ARG;
  const size_t n = (size_t)w * h;
  if (H.depth == 8) for (size_t i = 0; i < n; ++i) dst[i] = (float)px[i];
  else for (size_t i = 0; i < n; ++i) dst[i] = (float)((px[2 * i] << 8) | px[2 * i + 1]);
  return VDO_OK;
}

int vdo_io_flo_info(const char* path, int* w, int* h) {
  if (!path) return VDO_ERR_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return VDO_ERR_ARG;
  float tag = 0; int32_t wh[2] = {0, 0};
  const bool ok = std::fread(&tag, 4, 1, f) == 1 && std::fread(wh, 4, 2, f) == 2;
  std::fclose(f);
  if (!ok || tag != 202021.25f || wh[0] <= 0 || wh[1] <= 0) return VDO_ERR_ARG;
  if (w) *w = wh[0];
  if (h) *h = wh[1];
  return VDO_OK;
}
// dst: h x w x 2 floats (u, v interleaved) = the CV_32FC2 matrix of cv::optflow::readOpticalFlow
int vdo_io_read_flo(const char* path, float* dst, size_t dst_floats) {
  if (!path || !dst) return VDO_ERR_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return VDO_ERR_ARG;
  float tag = 0; int32_t wh[2] = {0, 0};
  int rc = VDO_OK;
  if (std::fread(&tag, 4, 1, f) != 1 || std::fread(wh, 4, 2, f) != 2 || tag != 202021.25f || wh[0] <= 0 || wh[1] <= 0) rc = VDO_ERR_ARG;
  const size_t n = rc == VDO_OK ? 2 * (size_t)wh[0] * (size_t)wh[1] : 0;
  if (rc == VDO_OK && (dst_floats < n || std::fread(dst, 4, n, f) != n)) rc = VDO_ERR_ARG;
  std::fclose(f);
  return rc;
}

// LoadMask (example/vdo_slam.cc:253-450 minus the display colouring): text row r holds the w labels of image row r.
// Rows beyond h are an error, missing rows / entries stay 0.
int vdo_io_read_mask_txt(const char* path, int32_t* dst, int w, int h) {
  if (!path || !dst || w <= 0 || h <= 0) return VDO_ERR_ARG;
  std::vector<unsigned char> buf;
  if (!read_file(path, buf)) return VDO_ERR_ARG;
  buf.push_back(0);
  std::memset(dst, 0, sizeof(int32_t) * (size_t)w * h);
  const char* p = (const char*)buf.data();
  int row = 0;
  while (*p) {
    // one line
    int col = 0;
    bool any = false;
    while (*p && *p != '\n') {
      while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
      if (!*p || *p == '\n') break;
      bool neg = false;
      if (*p == '-') { neg = true; ++p; }
      if (*p < '0' || *p > '9') return VDO_ERR_ARG;
      long v = 0;
      while (*p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; }
      any = true;
      if (col < w) { if (row >= h) return VDO_ERR_ARG; dst[(size_t)row * w + col] = (int32_t)(neg ? -v : v); }
      ++col;
    }
    if (*p == '\n') ++p;
    if (any) ++row;                      // empty lines are skipped like the reference's `if(!s.empty())`
  }
  return VDO_OK;
}

}  // extern "C"
