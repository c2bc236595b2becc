// image_io.cpp — display conversion and file output of the accumulation buffer
// (reference: MinimalOptiX::updateContent / saveCurrentFrame, MinimalOptiX.cpp:43-84; Qt's
// QImage/QColor replaced by a dependency-free PNG/PPM writer).
#include "image_io.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace moxh {

// QColor::setRedF stores round(v * 65535) as 16 bit; QImage RGB888 keeps the high byte.
static inline uint8_t quantise(float v) {
  v = fmaxf(0.f, fminf(v, 1.f));
  return (uint8_t)(((uint32_t)floorf(v * 65535.0f + 0.5f)) >> 8);
}

void accumToRgb8(const float* accum, uint32_t W, uint32_t H, float nAccum, uint8_t* out) {
  for (uint32_t i = 0; i < H; ++i)
    for (uint32_t j = 0; j < W; ++j) {
      const float* src = accum + 3 * ((size_t)i * W + j);
      uint8_t* dst = out + 3 * ((size_t)(H - 1 - i) * W + j);  // accumulator row 0 is the image bottom
      dst[0] = quantise(src[0] / nAccum);
      dst[1] = quantise(src[1] / nAccum);
      dst[2] = quantise(src[2] / nAccum);
    }
}

namespace {
uint32_t crcTable[256];
bool crcInit = false;
uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n) {
  if (!crcInit) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = c & 1 ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      crcTable[i] = c;
    }
    crcInit = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = crcTable[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}
void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(FILE* f, const char* type, const std::vector<uint8_t>& data) {
  std::vector<uint8_t> buf;
  put32(buf, (uint32_t)data.size());
  fwrite(buf.data(), 1, 4, f);
  std::vector<uint8_t> body(type, type + 4);
  body.insert(body.end(), data.begin(), data.end());
  fwrite(body.data(), 1, body.size(), f);
  buf.clear();
  put32(buf, crc32(0, body.data(), body.size()));
  fwrite(buf.data(), 1, 4, f);
}
}  // namespace

bool writeImage(const std::string& path, const uint8_t* rgb, uint32_t W, uint32_t H, std::string& err) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot open " + path; return false; }
  bool ppm = path.size() > 4 && path.substr(path.size() - 4) == ".ppm";
  if (ppm) {
    fprintf(f, "P6\n%u %u\n255\n", W, H);
    fwrite(rgb, 1, (size_t)W * H * 3, f);
    fclose(f);
    return true;
  }
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  fwrite(sig, 1, 8, f);
  std::vector<uint8_t> ihdr;
  put32(ihdr, W); put32(ihdr, H);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk(f, "IHDR", ihdr);
  // zlib stream of stored (uncompressed) deflate blocks over the filtered scanlines.
  std::vector<uint8_t> raw;
  raw.reserve((size_t)H * (3 * W + 1));
  for (uint32_t y = 0; y < H; ++y) { raw.push_back(0); raw.insert(raw.end(), rgb + (size_t)y * W * 3, rgb + (size_t)(y + 1) * W * 3); }
  std::vector<uint8_t> z;
  z.push_back(0x78); z.push_back(0x01);
  size_t pos = 0;
  uint32_t a = 1, b = 0;
  while (pos < raw.size() || raw.empty()) {
    size_t n = std::min<size_t>(65535, raw.size() - pos);
    bool last = pos + n >= raw.size();
    z.push_back(last ? 1 : 0);
    z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    for (size_t i = pos; i < pos + n; ++i) { a = (a + raw[i]) % 65521; b = (b + a) % 65521; }
    pos += n;
    if (last) break;
  }
  put32(z, (b << 16) | a);
  chunk(f, "IDAT", z);
  chunk(f, "IEND", {});
  fclose(f);
  return true;
}

bool writeAccum(const std::string& path, const float* accum, uint32_t W, uint32_t H, uint64_t launches, std::string& err) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot open " + path; return false; }
  fwrite("MOXA", 1, 4, f);
  fwrite(&W, 4, 1, f); fwrite(&H, 4, 1, f); fwrite(&launches, 8, 1, f);
  fwrite(accum, sizeof(float), (size_t)W * H * 3, f);
  fclose(f);
  return true;
}

bool readAccum(const std::string& path, float* accum, uint32_t W, uint32_t H, uint64_t* launches, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  char magic[4]; uint32_t w = 0, h = 0; uint64_t n = 0;
  bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "MOXA", 4) == 0 && fread(&w, 4, 1, f) == 1 && fread(&h, 4, 1, f) == 1 &&
            fread(&n, 8, 1, f) == 1 && w == W && h == H && fread(accum, sizeof(float), (size_t)W * H * 3, f) == (size_t)W * H * 3;
  fclose(f);
  if (!ok) { err = "bad accumulator file " + path; return false; }
  if (launches) *launches = n;
  return true;
}

}  // namespace moxh
