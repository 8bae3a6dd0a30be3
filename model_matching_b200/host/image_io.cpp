#include "image_io.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace imgio {
namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  if (pa <= pb && pa <= pc) return a;
  return pb <= pc ? b : c;
}
bool fail(std::string* err, const char* m) { if (err) *err = m; return false; }
}  // namespace

bool read_png(const std::string& path, Image& out, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return fail(err, "cannot open file");
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  if (buf.size() < 8 || memcmp(buf.data(), sig, 8) != 0) return fail(err, "not a PNG");
  size_t p = 8;
  int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat;
  while (p + 12 <= buf.size()) {
    uint32_t len = be32(&buf[p]);
    const char* type = (const char*)&buf[p + 4];
    const uint8_t* d = &buf[p + 8];
    if (p + 12 + len > buf.size()) return fail(err, "truncated chunk");
    if (!memcmp(type, "IHDR", 4)) {
      W = (int)be32(d); H = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), d, d + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    p += 12 + len;
  }
  if (W <= 0 || H <= 0) return fail(err, "missing IHDR");
  if (interlace) return fail(err, "interlaced PNG not supported");
  int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 6 ? 4 : ctype == 4 ? 2 : 0;
  if (!ch || (depth != 8 && depth != 16)) return fail(err, "unsupported colour type / bit depth");
  const int bpp = ch * depth / 8;
  const size_t stride = (size_t)W * bpp;
  std::vector<uint8_t> raw((stride + 1) * H);
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size())
    return fail(err, "zlib inflate failed");
  std::vector<uint8_t> img(stride * H);
  for (int y = 0; y < H; ++y) {
    const uint8_t* s = &raw[(stride + 1) * y];
    uint8_t* cur = &img[stride * y];
    const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
    const int ft = s[0];
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
      int v = s[1 + x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) / 2; break;
        case 4: v += paeth(a, b, c); break;
        default: return fail(err, "bad filter type");
      }
      cur[x] = (uint8_t)v;
    }
  }
  out = Image();
  out.width = W; out.height = H; out.bit_depth = depth;
  if (depth == 16) {
    // keep the first (gray) channel only
    out.channels = 1;
    out.u16.resize((size_t)W * H);
    for (size_t i = 0; i < (size_t)W * H; ++i) out.u16[i] = (uint16_t)((img[i * bpp] << 8) | img[i * bpp + 1]);
  } else {
    out.channels = ch == 2 ? 1 : ch;
    out.u8.resize((size_t)W * H * out.channels);
    if (ch == 2) for (size_t i = 0; i < (size_t)W * H; ++i) out.u8[i] = img[2 * i];
    else out.u8 = img;
  }
  return true;
}

bool write_png_gray8(const std::string& path, const uint8_t* data, int W, int H) {
  std::vector<uint8_t> raw((size_t)(W + 1) * H);
  for (int y = 0; y < H; ++y) { raw[(size_t)(W + 1) * y] = 0; memcpy(&raw[(size_t)(W + 1) * y + 1], data + (size_t)W * y, W); }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  fwrite(sig, 1, 8, f);
  auto chunk = [&](const char* type, const uint8_t* d, uint32_t len) {
    uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                      (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
    fwrite(hdr, 1, 8, f);
    if (len) fwrite(d, 1, len, f);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (len) crc = crc32(crc, d, len);
    uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    fwrite(c, 1, 4, f);
  };
  uint8_t ihdr[13] = {(uint8_t)(W >> 24), (uint8_t)(W >> 16), (uint8_t)(W >> 8), (uint8_t)W,
                      (uint8_t)(H >> 24), (uint8_t)(H >> 16), (uint8_t)(H >> 8), (uint8_t)H, 8, 0, 0, 0, 0};
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", comp.data(), (uint32_t)clen);
  chunk("IEND", nullptr, 0);
  fclose(f);
  return true;
}

bool load_bgr8(const std::string& path, std::vector<uint8_t>& bgr, int& W, int& H) {
  Image im;
  if (!read_png(path, im) || im.bit_depth != 8) return false;
  W = im.width; H = im.height;
  bgr.resize((size_t)W * H * 3);
  for (size_t i = 0; i < (size_t)W * H; ++i) {
    uint8_t r, g, b;
    if (im.channels == 1) r = g = b = im.u8[i];
    else { r = im.u8[i * im.channels]; g = im.u8[i * im.channels + 1]; b = im.u8[i * im.channels + 2]; }
    bgr[3 * i] = b; bgr[3 * i + 1] = g; bgr[3 * i + 2] = r;
  }
  return true;
}

bool load_gray16(const std::string& path, std::vector<uint16_t>& g, int& W, int& H) {
  Image im;
  if (!read_png(path, im)) return false;
  W = im.width; H = im.height;
  g.resize((size_t)W * H);
  if (im.bit_depth == 16) g = im.u16;
  else for (size_t i = 0; i < (size_t)W * H; ++i) g[i] = im.u8[i * im.channels];
  return true;
}

bool load_gray8(const std::string& path, std::vector<uint8_t>& g, int& W, int& H) {
  Image im;
  if (!read_png(path, im)) return false;
  W = im.width; H = im.height;
  g.resize((size_t)W * H);
  if (im.bit_depth == 16) for (size_t i = 0; i < (size_t)W * H; ++i) g[i] = (uint8_t)(im.u16[i] >> 8);
  else if (im.channels == 1) g = im.u8;
  else for (size_t i = 0; i < (size_t)W * H; ++i) {  // cv gray conversion
    const uint8_t* p = &im.u8[i * im.channels];
    g[i] = (uint8_t)((p[0] * 299 + p[1] * 587 + p[2] * 114 + 500) / 1000);
  }
  return true;
}

}  // namespace imgio
