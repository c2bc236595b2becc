// jpeg_decode.cpp — see jpeg_decode.h.  ITU-T T.81 Huffman decoding (sequential and progressive) and
// libjpeg's reconstruction arithmetic, written from the published algorithms:
//   * inverse DCT: Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits between the passes
//     ("islow"), samples wrapped to 10 bits and clamped;
//   * chroma upsampling: 3/4-1/4 triangle filter per direction ("fancy" upsampling) over the real
//     (un-padded) downsampled extent, edge samples replicated; components no wider than two samples
//     are box-replicated;
//   * YCbCr -> RGB: 16-bit fixed point, rounding folded into the Cr->R, Cb->B and Cb->G terms.
#include "jpeg_decode.h"

#include <algorithm>
#include <cstring>

namespace moxh {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  uint8_t vals[256];
  int maxcode[17], mincode[17], valptr[17];
  void build(const uint8_t counts[17]) {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += counts[l];
      k += counts[l];
      maxcode[l] = counts[l] ? code - 1 : -1;
      code <<= 1;
    }
    present = true;
  }
};

struct BitReader {
  const uint8_t* p = nullptr;
  size_t n = 0, pos = 0;
  uint32_t buf = 0;
  int cnt = 0;
  bool atMarker = false;
  void reset() { buf = 0; cnt = 0; atMarker = false; }
  void fill() {
    while (cnt <= 24) {
      uint32_t c = 0;
      if (!atMarker) {
        if (pos >= n) atMarker = true;
        else {
          c = p[pos];
          if (c == 0xff) {
            uint8_t c2 = pos + 1 < n ? p[pos + 1] : 0xd9;
            if (c2 == 0) pos += 2;                       // stuffed zero
            else { atMarker = true; c = 0; }             // a marker ends the segment: feed zeros from here on
          } else ++pos;
        }
      }
      buf |= c << (24 - cnt);
      cnt += 8;
    }
  }
  int bit() {
    if (cnt < 1) fill();
    int b = (int)(buf >> 31);
    buf <<= 1; --cnt;
    return b;
  }
  int bits(int k) {
    if (k == 0) return 0;
    if (cnt < k) fill();
    int v = (int)(buf >> (32 - k));
    buf <<= k; cnt -= k;
    return v;
  }
  int decode(const HuffTable& t) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
      code = (code << 1) | bit();
      if (t.maxcode[l] >= 0 && code <= t.maxcode[l]) return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    return 0;
  }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;          // tables of the current scan
  int dsW = 0, dsH = 0;        // real downsampled extent
  int bw = 0, bh = 0;          // blocks stored (padded to whole MCUs)
  int pred = 0;
  bool quantLatched = false;
  uint16_t quant[64];
  std::vector<int16_t> coef;
  std::vector<uint8_t> plane;  // bw*8 x bh*8 samples after the inverse DCT
};

const int64_t F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
              F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;

inline int64_t descale(int64_t x, int n) { return (x + ((int64_t)1 << (n - 1))) >> n; }

// one 1-D pass over 8 values spaced `stride` apart; shiftEven applies to the results
inline void idct1d(const int64_t* in, int stride, int64_t* out, int ostride, int shift) {
  int64_t z2 = in[2 * stride], z3 = in[6 * stride];
  int64_t z1 = (z2 + z3) * F_0_541;
  int64_t tmp2 = z1 - z3 * F_1_847, tmp3 = z1 + z2 * F_0_765;
  z2 = in[0]; z3 = in[4 * stride];
  int64_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
  const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7 * stride]; tmp1 = in[5 * stride]; tmp2 = in[3 * stride]; tmp3 = in[1 * stride];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int64_t z4 = tmp1 + tmp3;
  const int64_t z5 = (z3 + z4) * F_1_175;
  tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
  z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  out[0 * ostride] = descale(tmp10 + tmp3, shift); out[7 * ostride] = descale(tmp10 - tmp3, shift);
  out[1 * ostride] = descale(tmp11 + tmp2, shift); out[6 * ostride] = descale(tmp11 - tmp2, shift);
  out[2 * ostride] = descale(tmp12 + tmp1, shift); out[5 * ostride] = descale(tmp12 - tmp1, shift);
  out[3 * ostride] = descale(tmp13 + tmp0, shift); out[4 * ostride] = descale(tmp13 - tmp0, shift);
}

void idctBlock(const int16_t* coef, const uint16_t* quant, uint8_t* dst, int dstStride) {
  int64_t in[64], ws[64], o[64];
  for (int i = 0; i < 64; ++i) in[i] = (int64_t)coef[i] * quant[i];
  for (int c = 0; c < 8; ++c) idct1d(in + c, 8, ws + c, 8, 13 - 2);       // columns
  for (int r = 0; r < 8; ++r) idct1d(ws + 8 * r, 1, o + 8 * r, 1, 13 + 2 + 3);  // rows
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) {
      int y = (int)(o[8 * r + c] & 1023);
      if (y >= 512) y -= 1024;
      dst[r * dstStride + c] = (uint8_t)std::min(255, std::max(0, y + 128));
    }
}

struct Decoder {
  BitReader br;
  int W = 0, H = 0, ncomp = 0, hmax = 1, vmax = 1, mcusX = 0, mcusY = 0;
  bool progressive = false, sawSof = false;
  int restartInterval = 0;
  int adobeTransform = -1;
  bool sawJfif = false;
  uint16_t qt[4][64];
  bool qtPresent[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  Component comp[3];
  uint32_t eobrun = 0;
  std::string err;

  bool fail(const char* m) { if (err.empty()) err = m; return false; }

  bool parseSof(const uint8_t* s, size_t len) {
    if (len < 6 || s[0] != 8) return fail("unsupported JPEG sample precision (need 8 bit)");
    H = (s[1] << 8) | s[2]; W = (s[3] << 8) | s[4]; ncomp = s[5];
    if (W <= 0 || H <= 0) return fail("bad JPEG dimensions");
    if ((int64_t)W * H > ((int64_t)1 << 28)) return fail("JPEG larger than 2^28 pixels");
    if (ncomp != 1 && ncomp != 3) return fail("unsupported JPEG component count (need 1 or 3)");
    if (len < (size_t)6 + 3 * ncomp) return fail("truncated JPEG frame header");
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      c.id = s[6 + 3 * i]; c.h = s[7 + 3 * i] >> 4; c.v = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i] & 3;
      if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4) return fail("bad JPEG sampling factors");
      hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
    }
    if (ncomp == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }  // a lone component is never subsampled
    mcusX = (W + 8 * hmax - 1) / (8 * hmax); mcusY = (H + 8 * vmax - 1) / (8 * vmax);
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      if (hmax % c.h || vmax % c.v) return fail("unsupported JPEG sampling ratio");
      c.dsW = (W * c.h + hmax - 1) / hmax; c.dsH = (H * c.v + vmax - 1) / vmax;
      c.bw = mcusX * c.h; c.bh = mcusY * c.v;
      c.coef.assign((size_t)c.bw * c.bh * 64, 0);
    }
    sawSof = true;
    return true;
  }

  bool parseDht(const uint8_t* s, size_t len) {
    size_t i = 0;
    while (i + 17 <= len) {
      int tc = s[i] >> 4, th = s[i] & 15;
      if (tc > 1 || th > 3) return fail("bad JPEG Huffman table id");
      uint8_t counts[17] = {0};
      int total = 0;
      for (int l = 1; l <= 16; ++l) { counts[l] = s[i + l]; total += counts[l]; }
      i += 17;
      if (total > 256 || i + total > len) return fail("bad JPEG Huffman table");
      HuffTable& t = tc ? ac[th] : dc[th];
      memset(t.vals, 0, sizeof t.vals);
      memcpy(t.vals, s + i, total);
      t.build(counts);
      i += total;
    }
    return true;
  }

  bool parseDqt(const uint8_t* s, size_t len) {
    size_t i = 0;
    while (i < len) {
      int pq = s[i] >> 4, tq = s[i] & 15;
      ++i;
      if (tq > 3 || pq > 1 || i + (pq ? 128 : 64) > len) return fail("bad JPEG quantisation table");
      for (int k = 0; k < 64; ++k) {
        qt[tq][kZigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
        i += pq ? 2 : 1;
      }
      qtPresent[tq] = true;
    }
    return true;
  }

  bool restart() {
    br.reset();
    size_t& p = br.pos;
    while (p < br.n && br.p[p] != 0xff) ++p;  // tolerate garbage before the marker
    while (p + 1 < br.n && br.p[p + 1] == 0xff) ++p;
    if (p + 1 >= br.n || br.p[p + 1] < 0xd0 || br.p[p + 1] > 0xd7) return fail("missing JPEG restart marker");
    p += 2;
    for (int i = 0; i < ncomp; ++i) comp[i].pred = 0;
    eobrun = 0;
    return true;
  }

  // ---- block decoders (T.81 F.2.2, G.1.2) ----
  void blockSequential(Component& c, int16_t* b) {
    const int t = std::min(br.decode(dc[c.td]), 16);  // categories above 16 only occur in damaged tables
    c.pred = (int)((uint32_t)c.pred + (uint32_t)(t ? extend(br.bits(t), t) : 0));
    b[0] = (int16_t)c.pred;
    for (int k = 1; k < 64;) {
      int rs = br.decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
      if (s) {
        k += r;
        if (k > 63) break;
        b[kZigzag[k++]] = (int16_t)extend(br.bits(s), s);
      } else {
        if (r != 15) break;
        k += 16;
      }
    }
  }
  void blockDcFirst(Component& c, int16_t* b, int al) {
    const int t = std::min(br.decode(dc[c.td]), 16);  // categories above 16 only occur in damaged tables
    c.pred = (int)((uint32_t)c.pred + (uint32_t)(t ? extend(br.bits(t), t) : 0));
    b[0] = (int16_t)((uint32_t)c.pred << al);
  }
  void blockDcRefine(int16_t* b, int al) {
    if (br.bit()) b[0] |= (int16_t)(1 << al);
  }
  void blockAcFirst(Component& c, int16_t* b, int ss, int se, int al) {
    if (eobrun) { --eobrun; return; }
    for (int k = ss; k <= se;) {
      int rs = br.decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
      if (s) {
        k += r;
        if (k > 63) break;
        b[kZigzag[k++]] = (int16_t)(extend(br.bits(s), s) * (1 << al));
      } else {
        if (r != 15) { eobrun = (1u << r) - 1u + (r ? (uint32_t)br.bits(r) : 0u); break; }
        k += 16;
      }
    }
  }
  void blockAcRefine(Component& c, int16_t* b, int ss, int se, int al) {
    const int p1 = 1 << al, m1 = -(1 << al);
    int k = ss;
    auto correct = [&](int16_t& v) {
      if (br.bit() && (v & p1) == 0) v = (int16_t)(v + (v >= 0 ? p1 : m1));
    };
    if (eobrun == 0) {
      for (; k <= se; ++k) {
        int rs = br.decode(ac[c.ta]), r = rs >> 4, s = rs & 15;
        int value = 0;
        if (s) value = br.bit() ? p1 : m1;
        else if (r != 15) { eobrun = (1u << r) + (r ? (uint32_t)br.bits(r) : 0u); break; }
        // skip r still-zero coefficients, correcting the non-zero ones passed on the way
        while (k <= se) {
          int16_t& v = b[kZigzag[k]];
          if (v != 0) correct(v);
          else if (--r < 0) break;
          ++k;
        }
        if (value && k <= se) b[kZigzag[k]] = (int16_t)value;
      }
    }
    if (eobrun > 0) {
      for (; k <= se; ++k) {
        int16_t& v = b[kZigzag[k]];
        if (v != 0) correct(v);
      }
      --eobrun;
    }
  }

  bool decodeScan(const uint8_t* s, size_t len) {
    if (!sawSof) return fail("JPEG scan before frame header");
    if (len < 1) return fail("truncated JPEG scan header");
    int ns = s[0];
    if (ns < 1 || ns > ncomp || len < (size_t)4 + 2 * ns) return fail("bad JPEG scan header");
    Component* sc[3];
    for (int i = 0; i < ns; ++i) {
      int id = s[1 + 2 * i];
      sc[i] = nullptr;
      for (int j = 0; j < ncomp; ++j) if (comp[j].id == id) sc[i] = &comp[j];
      if (!sc[i]) return fail("JPEG scan names an unknown component");
      sc[i]->td = s[2 + 2 * i] >> 4; sc[i]->ta = s[2 + 2 * i] & 15;
      if (sc[i]->td > 3 || sc[i]->ta > 3) return fail("bad JPEG table selector");
      if (!sc[i]->quantLatched) {
        if (!qtPresent[sc[i]->tq]) return fail("JPEG quantisation table missing");
        memcpy(sc[i]->quant, qt[sc[i]->tq], sizeof qt[0]);
        sc[i]->quantLatched = true;
      }
    }
    int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
    if (!progressive) { ss = 0; se = 63; ah = al = 0; }
    else if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1) || al > 13) return fail("bad progressive JPEG scan parameters");
    const bool needDc = ss == 0 && ah == 0, needAc = !progressive || ss != 0;
    for (int i = 0; i < ns; ++i) {
      if (needDc && !dc[sc[i]->td].present) return fail("JPEG DC Huffman table missing");
      if (needAc && !ac[sc[i]->ta].present) return fail("JPEG AC Huffman table missing");
      sc[i]->pred = 0;
    }
    eobrun = 0;
    br.reset();
    int todo = restartInterval;
    auto block = [&](Component& c, int bx, int by) {
      int16_t* b = &c.coef[((size_t)by * c.bw + bx) * 64];
      if (!progressive) blockSequential(c, b);
      else if (ss == 0) { if (ah == 0) blockDcFirst(c, b, al); else blockDcRefine(b, al); }
      else { if (ah == 0) blockAcFirst(c, b, ss, se, al); else blockAcRefine(c, b, ss, se, al); }
    };
    if (ns == 1) {
      Component& c = *sc[0];
      const int nbx = (c.dsW + 7) / 8, nby = (c.dsH + 7) / 8;
      for (int by = 0; by < nby; ++by)
        for (int bx = 0; bx < nbx; ++bx) {
          if (restartInterval) { if (todo == 0) { if (!restart()) return false; todo = restartInterval; } --todo; }
          block(c, bx, by);
        }
    } else {
      for (int my = 0; my < mcusY; ++my)
        for (int mx = 0; mx < mcusX; ++mx) {
          if (restartInterval) { if (todo == 0) { if (!restart()) return false; todo = restartInterval; } --todo; }
          for (int i = 0; i < ns; ++i)
            for (int v = 0; v < sc[i]->v; ++v)
              for (int h = 0; h < sc[i]->h; ++h) block(*sc[i], mx * sc[i]->h + h, my * sc[i]->v + v);
        }
    }
    return true;
  }

  // ---- reconstruction ----
  void inverseDct() {
    for (int i = 0; i < ncomp; ++i) {
      Component& c = comp[i];
      const int stride = c.bw * 8;
      c.plane.assign((size_t)stride * c.bh * 8, 0);
      for (int by = 0; by < c.bh; ++by)
        for (int bx = 0; bx < c.bw; ++bx)
          idctBlock(&c.coef[((size_t)by * c.bw + bx) * 64], c.quant, &c.plane[(size_t)by * 8 * stride + bx * 8], stride);
      std::vector<int16_t>().swap(c.coef);
    }
  }

  // full-resolution plane (W x H) of one component
  void upsample(const Component& c, std::vector<uint8_t>& out) const {
    out.resize((size_t)W * H);
    const int fx = hmax / c.h, fy = vmax / c.v, stride = c.bw * 8;
    auto row = [&](int r) { return &c.plane[(size_t)std::min(std::max(r, 0), c.dsH - 1) * stride]; };
    const bool fancy = c.dsW > 2;
    std::vector<int> sum(c.dsW + 2);
    std::vector<uint8_t> line((size_t)c.dsW * 2 + 2);
    for (int y = 0; y < H; ++y) {
      uint8_t* o = &out[(size_t)y * W];
      if (fx == 1 && fy == 1) { memcpy(o, row(y), W); continue; }
      if (fx == 2 && fy == 1 && fancy) {
        const uint8_t* in = row(y);
        const int n = c.dsW;
        line[0] = in[0]; line[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
        for (int x = 1; x < n - 1; ++x) {
          line[2 * x] = (uint8_t)((in[x] * 3 + in[x - 1] + 1) >> 2);
          line[2 * x + 1] = (uint8_t)((in[x] * 3 + in[x + 1] + 2) >> 2);
        }
        line[2 * n - 2] = (uint8_t)((in[n - 1] * 3 + in[n - 2] + 1) >> 2); line[2 * n - 1] = in[n - 1];
        memcpy(o, line.data(), W);
        continue;
      }
      if (fx == 2 && fy == 2 && fancy) {
        const int r = y >> 1;
        const uint8_t *in0 = row(r), *in1 = row((y & 1) ? r + 1 : r - 1);
        const int n = c.dsW;
        for (int x = 0; x < n; ++x) sum[x] = in0[x] * 3 + in1[x];
        line[0] = (uint8_t)((sum[0] * 4 + 8) >> 4); line[1] = (uint8_t)((sum[0] * 3 + sum[1] + 7) >> 4);
        for (int x = 1; x < n - 1; ++x) {
          line[2 * x] = (uint8_t)((sum[x] * 3 + sum[x - 1] + 8) >> 4);
          line[2 * x + 1] = (uint8_t)((sum[x] * 3 + sum[x + 1] + 7) >> 4);
        }
        line[2 * n - 2] = (uint8_t)((sum[n - 1] * 3 + sum[n - 2] + 8) >> 4); line[2 * n - 1] = (uint8_t)((sum[n - 1] * 4 + 7) >> 4);
        memcpy(o, line.data(), W);
        continue;
      }
      if (fx == 1 && fy == 2) {
        const int r = y >> 1;
        const uint8_t *in0 = row(r), *in1 = row((y & 1) ? r + 1 : r - 1);
        const int bias = (y & 1) ? 2 : 1;
        for (int x = 0; x < W; ++x) o[x] = (uint8_t)((in0[x] * 3 + in1[x] + bias) >> 2);
        continue;
      }
      const uint8_t* in = row(y / fy);  // box replication
      for (int x = 0; x < W; ++x) o[x] = in[x / fx];
    }
  }

  bool run(const uint8_t* data, size_t size, int& w, int& h, std::vector<uint8_t>& rgb) {
    if (size < 4 || data[0] != 0xff || data[1] != 0xd8) return fail("not a JPEG file");
    br.p = data; br.n = size;
    size_t pos = 2;
    bool done = false;
    while (!done) {
      while (pos < size && data[pos] != 0xff) ++pos;
      while (pos < size && data[pos] == 0xff) ++pos;
      if (pos >= size) break;
      const uint8_t m = data[pos++];
      if (m == 0xd9) break;
      if (m == 0x00 || m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;  // stuffed byte left over from a scan, TEM, RSTn
      if (pos + 2 > size) return fail("truncated JPEG");
      const size_t len = ((size_t)data[pos] << 8) | data[pos + 1];
      if (len < 2 || pos + len > size) return fail("truncated JPEG segment");
      const uint8_t* body = data + pos + 2;
      const size_t blen = len - 2;
      pos += len;
      switch (m) {
        case 0xc0: case 0xc1: case 0xc2:
          if (sawSof) return fail("JPEG with several frames");
          progressive = m == 0xc2;
          if (!parseSof(body, blen)) return false;
          break;
        case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
          return fail("unsupported JPEG process (lossless, hierarchical or arithmetic coding)");
        case 0xc4: if (!parseDht(body, blen)) return false; break;
        case 0xdb: if (!parseDqt(body, blen)) return false; break;
        case 0xdd: if (blen >= 2) restartInterval = (body[0] << 8) | body[1]; break;
        case 0xe0: if (blen >= 5 && !memcmp(body, "JFIF", 5)) sawJfif = true; break;
        case 0xee: if (blen >= 12 && !memcmp(body, "Adobe", 5)) adobeTransform = body[11]; break;
        case 0xda:
          br.pos = pos;
          if (!decodeScan(body, blen)) return false;
          pos = br.pos;
          break;
        default: break;
      }
    }
    if (!sawSof) return fail("JPEG without a frame header");
    for (int i = 0; i < ncomp; ++i) if (!comp[i].quantLatched) return fail("JPEG component without a scan");
    inverseDct();
    w = W; h = H;
    rgb.resize((size_t)W * H * 3);
    std::vector<uint8_t> pl[3];
    for (int i = 0; i < ncomp; ++i) upsample(comp[i], pl[i]);
    if (ncomp == 1) {
      for (size_t i = 0; i < (size_t)W * H; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = pl[0][i];
      return true;
    }
    bool isRgb = false;
    if (sawJfif) isRgb = false;
    else if (adobeTransform >= 0) isRgb = adobeTransform == 0;
    else isRgb = comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B';
    for (size_t i = 0; i < (size_t)W * H; ++i) {
      const int y = pl[0][i], cb = pl[1][i] - 128, cr = pl[2][i] - 128;
      int r, g, b;
      if (isRgb) { r = pl[0][i]; g = pl[1][i]; b = pl[2][i]; }
      else {
        r = y + (int)((91881 * (int64_t)cr + 32768) >> 16);
        b = y + (int)((116130 * (int64_t)cb + 32768) >> 16);
        g = y + (int)((-22554 * (int64_t)cb + 32768 - 46802 * (int64_t)cr) >> 16);
      }
      rgb[3 * i] = (uint8_t)std::min(255, std::max(0, r));
      rgb[3 * i + 1] = (uint8_t)std::min(255, std::max(0, g));
      rgb[3 * i + 2] = (uint8_t)std::min(255, std::max(0, b));
    }
    return true;
  }
};

}  // namespace

bool decodeJpeg(const uint8_t* data, size_t size, int& w, int& h, std::vector<uint8_t>& rgb, std::string& err) {
  Decoder d;
  if (d.run(data, size, w, h, rgb)) return true;
  err = d.err.empty() ? "corrupt JPEG" : d.err;
  return false;
}

}  // namespace moxh
