// JPEG texture decoding for the loader.  The reference hands texture files to SDL2_image
// (Texture::LoadFromFile, texture.cc:60-109), whose JPEG path is libjpeg with its defaults.  This is an own decoder
// of the same streams - baseline / extended sequential and progressive Huffman JPEG, 8-bit samples, greyscale or three
// components, restart intervals, interleaved and non-interleaved scans - that follows libjpeg's DEFAULT arithmetic step
// by step, so that the texels are the bytes libjpeg (6b / libjpeg-turbo) produces, not merely close to them:
//   * inverse DCT: the "islow" integer transform (13-bit constants, two passes, jidctint.c);
//   * chroma upsampling: "fancy" triangle filters for 2h1v and 2h2v (jdsample.c), plain replication when the
//     subsampled plane is at most two samples wide;
//   * YCbCr -> RGB: 16-bit fixed-point tables (jdcolor.c).
// tests/test_host_logic.py compares with fixtures decoded by libjpeg-turbo (tests/golden/jpeg/).  CMYK / YCCK,
// arithmetic coding, 12-bit and lossless JPEG fail the load, as any undecodable texture does upstream
// (objreader.cc:467-469).
#include <cstdint>
#include <cstring>
#include <vector>

#include "scene_build.h"

namespace mtb {
namespace {

constexpr int kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  uint8_t vals[256];
  int mincode[17], maxcode[17], valptr[17];  // per code length 1..16 (maxcode -1: no codes of this length)
  uint16_t look[512];                        // 9-bit prefix -> (length << 8) | value, 0: longer code

  void Build(const uint8_t counts[16], const uint8_t *symbols, int n) {
    memcpy(vals, symbols, (size_t)n);
    memset(look, 0, sizeof(look));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
      valptr[len] = k;
      mincode[len] = code;
      for (int i = 0; i < counts[len - 1]; i++, k++, code++) {
        if (len <= 9) {
          const int first = code << (9 - len);
          for (int j = 0; j < (1 << (9 - len)); j++) look[first + j] = (uint16_t)((len << 8) | vals[k]);
        }
      }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    present = true;
  }
};

// Entropy-coded segment reader: removes the 0xFF00 stuffing, stops feeding at a marker (zero bits from there on).
struct JpegBits {
  const uint8_t *p, *end;
  uint32_t acc = 0;
  int n = 0;
  int marker = 0;  // the marker the segment ended at (0: none seen yet)

  void Fill() {
    while (n <= 24) {
      uint32_t byte = 0;
      if (marker == 0 && p < end) {
        byte = *p++;
        if (byte == 0xff) {
          while (p < end && *p == 0xff) p++;  // fill bytes
          const int m = p < end ? *p++ : 0xd9;
          if (m != 0) {
            marker = m;
            byte = 0;
          }
        }
      } else if (marker == 0) {
        marker = 0xd9;
      }
      acc |= byte << (24 - n);
      n += 8;
    }
  }
  int Peek(int k) {
    if (n < k) Fill();
    return (int)(acc >> (32 - k));
  }
  void Skip(int k) {
    acc <<= k;
    n -= k;
  }
  int Get(int k) {
    if (k == 0) return 0;
    const int v = Peek(k);
    Skip(k);
    return v;
  }
  int Decode(const HuffTable &h) {
    const int e = h.look[Peek(9)];
    if (e != 0) {
      Skip(e >> 8);
      return e & 255;
    }
    int code = Peek(16);
    for (int len = 10; len <= 16; len++) {
      const int c = code >> (16 - len);
      if (h.maxcode[len] >= 0 && c <= h.maxcode[len] && c >= h.mincode[len]) {
        Skip(len);
        return h.vals[h.valptr[len] + c - h.mincode[len]];
      }
    }
    Skip(16);
    return -1;
  }
  // restart: drop the partial byte, take the RSTn marker
  bool Restart() {
    acc = 0;
    n = 0;
    if (marker == 0) {  // the marker has not been run into yet: it must be next
      while (p < end && *p != 0xff) p++;
      while (p < end && *p == 0xff) p++;
      if (p < end) marker = *p++;
    }
    const bool ok = marker >= 0xd0 && marker <= 0xd7;
    marker = 0;
    return ok;
  }
};

inline int Extend(int v, int s) { return s == 0 ? 0 : (v < (1 << (s - 1)) ? v - (1 << s) + 1 : v); }

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;             // tables of the current scan
  int blocks_w = 0, blocks_h = 0;  // blocks that carry image data (non-interleaved scans walk these)
  int alloc_w = 0, alloc_h = 0;    // padded to whole MCUs
  int width = 0, height = 0;       // downsampled_width / downsampled_height of libjpeg
  int pred = 0;
  std::vector<int16_t> coef;
  std::vector<uint8_t> plane;  // alloc_w * 8 samples per row
};

// jidctint.c: jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2), one block, with libjpeg's range-limit table semantics.
inline uint8_t RangeLimit(int64_t x) {
  const int i = (int)(x & 1023);  // sample_range_limit + CENTERJSAMPLE, indexed & RANGE_MASK
  if (i < 128) return (uint8_t)(128 + i);
  if (i < 512) return 255;
  if (i < 896) return 0;
  return (uint8_t)(i - 896);
}
inline int64_t Descale(int64_t x, int n) { return (x + ((int64_t)1 << (n - 1))) >> n; }

void IdctIslow(const int16_t *coef, const uint16_t *quant, uint8_t *out, size_t stride) {
  constexpr int64_t F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633, F_1_501 = 12299,
                    F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
  int64_t ws[64];
  for (int c = 0; c < 8; c++) {
    const int16_t *in = coef + c;
    const uint16_t *q = quant + c;
    if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0) {
      const int64_t dc = ((int64_t)in[0] * q[0]) * 4;  // << PASS1_BITS
      for (int r = 0; r < 8; r++) ws[r * 8 + c] = dc;
      continue;
    }
    int64_t z2 = (int64_t)in[16] * q[16], z3 = (int64_t)in[48] * q[48];
    int64_t z1 = (z2 + z3) * F_0_541;
    int64_t tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    z2 = (int64_t)in[0] * q[0];
    z3 = (int64_t)in[32] * q[32];
    int64_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = (int64_t)in[56] * q[56];
    tmp1 = (int64_t)in[40] * q[40];
    tmp2 = (int64_t)in[24] * q[24];
    tmp3 = (int64_t)in[8] * q[8];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298;
    tmp1 *= F_2_053;
    tmp2 *= F_3_072;
    tmp3 *= F_1_501;
    z1 *= -F_0_899;
    z2 *= -F_2_562;
    z3 *= -F_1_961;
    z4 *= -F_0_390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    ws[0 * 8 + c] = Descale(tmp10 + tmp3, 11);
    ws[7 * 8 + c] = Descale(tmp10 - tmp3, 11);
    ws[1 * 8 + c] = Descale(tmp11 + tmp2, 11);
    ws[6 * 8 + c] = Descale(tmp11 - tmp2, 11);
    ws[2 * 8 + c] = Descale(tmp12 + tmp1, 11);
    ws[5 * 8 + c] = Descale(tmp12 - tmp1, 11);
    ws[3 * 8 + c] = Descale(tmp13 + tmp0, 11);
    ws[4 * 8 + c] = Descale(tmp13 - tmp0, 11);
  }
  for (int r = 0; r < 8; r++) {
    const int64_t *w = ws + r * 8;
    uint8_t *o = out + (size_t)r * stride;
    int64_t z2 = w[2], z3 = w[6];
    int64_t z1 = (z2 + z3) * F_0_541;
    int64_t tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    int64_t tmp0 = (w[0] + w[4]) * 8192, tmp1 = (w[0] - w[4]) * 8192;
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7];
    tmp1 = w[5];
    tmp2 = w[3];
    tmp3 = w[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298;
    tmp1 *= F_2_053;
    tmp2 *= F_3_072;
    tmp3 *= F_1_501;
    z1 *= -F_0_899;
    z2 *= -F_2_562;
    z3 *= -F_1_961;
    z4 *= -F_0_390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    o[0] = RangeLimit(Descale(tmp10 + tmp3, 18));
    o[7] = RangeLimit(Descale(tmp10 - tmp3, 18));
    o[1] = RangeLimit(Descale(tmp11 + tmp2, 18));
    o[6] = RangeLimit(Descale(tmp11 - tmp2, 18));
    o[2] = RangeLimit(Descale(tmp12 + tmp1, 18));
    o[5] = RangeLimit(Descale(tmp12 - tmp1, 18));
    o[3] = RangeLimit(Descale(tmp13 + tmp0, 18));
    o[4] = RangeLimit(Descale(tmp13 - tmp0, 18));
  }
}

struct JpegDecoder {
  const std::vector<uint8_t> &d;
  int width = 0, height = 0, ncomp = 0, max_h = 1, max_v = 1, mcus_x = 0, mcus_y = 0;
  bool progressive = false, have_frame = false, saw_jfif = false, saw_adobe = false;
  int adobe_transform = 0, restart_interval = 0;
  uint16_t quant[4][64];
  bool have_quant[4] = {false, false, false, false};
  HuffTable dc[4], ac[4];
  Component comp[3];

  explicit JpegDecoder(const std::vector<uint8_t> &data) : d(data) { memset(quant, 0, sizeof(quant)); }

  // ---- one block of a scan ----
  bool BlockSequential(JpegBits &br, Component &c, int16_t *blk) {
    const HuffTable &hd = dc[c.td], &ha = ac[c.ta];
    const int s = br.Decode(hd);
    if (s < 0 || s > 15) return false;
    c.pred += Extend(br.Get(s), s);
    blk[0] = (int16_t)c.pred;
    for (int k = 1; k < 64;) {
      const int rs = br.Decode(ha);
      if (rs < 0) return false;
      const int r = rs >> 4, sz = rs & 15;
      if (sz == 0) {
        if (r != 15) break;
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) return false;
      blk[kZigzag[k]] = (int16_t)Extend(br.Get(sz), sz);
      k++;
    }
    return true;
  }
  bool BlockDcFirst(JpegBits &br, Component &c, int16_t *blk, int al) {
    const int s = br.Decode(dc[c.td]);
    if (s < 0 || s > 15) return false;
    c.pred += Extend(br.Get(s), s);
    blk[0] = (int16_t)(c.pred * (1 << al));
    return true;
  }
  static void BlockDcRefine(JpegBits &br, int16_t *blk, int al) {
    if (br.Get(1)) blk[0] = (int16_t)(blk[0] | (1 << al));
  }
  bool BlockAcFirst(JpegBits &br, Component &c, int16_t *blk, int ss, int se, int al, int *eobrun) {
    if (*eobrun > 0) {
      (*eobrun)--;
      return true;
    }
    const HuffTable &ha = ac[c.ta];
    for (int k = ss; k <= se;) {
      const int rs = br.Decode(ha);
      if (rs < 0) return false;
      const int r = rs >> 4, sz = rs & 15;
      if (sz == 0) {
        if (r == 15) {
          k += 16;
          continue;
        }
        *eobrun = (1 << r) - 1;
        if (r) *eobrun += br.Get(r);
        break;
      }
      k += r;
      if (k > 63) return false;
      blk[kZigzag[k]] = (int16_t)(Extend(br.Get(sz), sz) * (1 << al));
      k++;
    }
    return true;
  }
  // jdphuff.c: decode_mcu_AC_refine
  bool BlockAcRefine(JpegBits &br, Component &c, int16_t *blk, int ss, int se, int al, int *eobrun) {
    const int p1 = 1 << al, m1 = -(1 << al);
    const HuffTable &ha = ac[c.ta];
    int k = ss;
    if (*eobrun == 0) {
      for (; k <= se; k++) {
        const int rs = br.Decode(ha);
        if (rs < 0) return false;
        int r = rs >> 4, s = rs & 15;
        if (s != 0) {
          s = br.Get(1) ? p1 : m1;
        } else if (r != 15) {
          *eobrun = 1 << r;
          if (r) *eobrun += br.Get(r);
          break;
        }
        do {
          int16_t *cp = blk + kZigzag[k];
          if (*cp != 0) {
            if (br.Get(1) && (*cp & p1) == 0) *cp = (int16_t)(*cp >= 0 ? *cp + p1 : *cp + m1);
          } else if (--r < 0) {
            break;
          }
          k++;
        } while (k <= se);
        if (s != 0) {
          if (k > 63) return false;
          blk[kZigzag[k]] = (int16_t)s;
        }
      }
    }
    if (*eobrun > 0) {
      for (; k <= se; k++) {
        int16_t *cp = blk + kZigzag[k];
        if (*cp != 0 && br.Get(1) && (*cp & p1) == 0) *cp = (int16_t)(*cp >= 0 ? *cp + p1 : *cp + m1);
      }
      (*eobrun)--;
    }
    return true;
  }

  bool DecodeScan(size_t *pos) {
    size_t p = *pos;
    if (p + 2 > d.size()) return false;
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    if (len < 6 || p + len > d.size()) return false;
    const int ns = d[p + 2];
    if (ns < 1 || ns > ncomp || len != (size_t)(6 + 2 * ns)) return false;
    Component *sc[3];
    for (int i = 0; i < ns; i++) {
      sc[i] = nullptr;
      for (int k = 0; k < ncomp; k++) {
        if (comp[k].id == d[p + 3 + 2 * i]) sc[i] = &comp[k];
      }
      if (sc[i] == nullptr) return false;
      sc[i]->td = d[p + 4 + 2 * i] >> 4;
      sc[i]->ta = d[p + 4 + 2 * i] & 15;
      if (sc[i]->td > 3 || sc[i]->ta > 3) return false;
    }
    const int ss = d[p + 3 + 2 * ns], se = d[p + 4 + 2 * ns], ah = d[p + 5 + 2 * ns] >> 4, al = d[p + 5 + 2 * ns] & 15;
    if (progressive) {
      if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) return false;
    } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
      return false;
    }
    for (int i = 0; i < ns; i++) {
      const bool need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || ss > 0;
      if ((need_dc && !dc[sc[i]->td].present) || (need_ac && !ac[sc[i]->ta].present)) return false;
      sc[i]->pred = 0;
    }
    JpegBits br;
    br.p = d.data() + p + len;
    br.end = d.data() + d.size();
    int eobrun = 0;
    const bool interleaved = ns > 1;
    const int units_x = interleaved ? mcus_x : sc[0]->blocks_w, units_y = interleaved ? mcus_y : sc[0]->blocks_h;
    int countdown = restart_interval;
    for (int uy = 0; uy < units_y; uy++) {
      for (int ux = 0; ux < units_x; ux++) {
        if (restart_interval > 0 && countdown == 0) {
          if (!br.Restart()) return false;
          for (int i = 0; i < ns; i++) sc[i]->pred = 0;
          eobrun = 0;
          countdown = restart_interval;
        }
        for (int i = 0; i < ns; i++) {
          Component &c = *sc[i];
          const int nbx = interleaved ? c.h : 1, nby = interleaved ? c.v : 1;
          for (int by = 0; by < nby; by++) {
            for (int bx = 0; bx < nbx; bx++) {
              const int x = interleaved ? ux * c.h + bx : ux, y = interleaved ? uy * c.v + by : uy;
              int16_t *blk = &c.coef[((size_t)y * c.alloc_w + x) * 64];
              bool ok;
              if (!progressive) {
                ok = BlockSequential(br, c, blk);
              } else if (ss == 0) {
                ok = true;
                if (ah == 0) {
                  ok = BlockDcFirst(br, c, blk, al);
                } else {
                  BlockDcRefine(br, blk, al);
                }
              } else {
                ok = ah == 0 ? BlockAcFirst(br, c, blk, ss, se, al, &eobrun) : BlockAcRefine(br, c, blk, ss, se, al, &eobrun);
              }
              if (!ok) return false;
            }
          }
        }
        countdown--;
      }
    }
    // continue behind the entropy-coded data: at the marker the reader ran into, or search for the next one
    if (br.marker != 0) {
      *pos = (size_t)(br.p - d.data()) - 2;
    } else {
      const uint8_t *q = br.p;
      while (q + 1 < br.end && !(q[0] == 0xff && q[1] != 0 && q[1] != 0xff && !(q[1] >= 0xd0 && q[1] <= 0xd7))) q++;
      *pos = (size_t)(q - d.data());
    }
    return true;
  }

  bool ParseFrame(size_t p, size_t len) {
    if (len < 8 || d[p + 2] != 8) return false;  // 8-bit samples only
    height = (d[p + 3] << 8) | d[p + 4];
    width = (d[p + 5] << 8) | d[p + 6];
    ncomp = d[p + 7];
    if (width <= 0 || height <= 0 || width > 30000 || height > 30000) return false;
    if ((ncomp != 1 && ncomp != 3) || len != (size_t)(8 + 3 * ncomp)) return false;
    for (int i = 0; i < ncomp; i++) {
      Component &c = comp[i];
      c.id = d[p + 8 + 3 * i];
      c.h = d[p + 9 + 3 * i] >> 4;
      c.v = d[p + 9 + 3 * i] & 15;
      c.tq = d[p + 10 + 3 * i];
      if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return false;
      max_h = c.h > max_h ? c.h : max_h;
      max_v = c.v > max_v ? c.v : max_v;
    }
    if (ncomp == 1) comp[0].h = comp[0].v = max_h = max_v = 1;  // a single component is never subsampled
    mcus_x = (width + 8 * max_h - 1) / (8 * max_h);
    mcus_y = (height + 8 * max_v - 1) / (8 * max_v);
    for (int i = 0; i < ncomp; i++) {
      Component &c = comp[i];
      if (max_h % c.h != 0 || max_v % c.v != 0) return false;
      c.width = (width * c.h + max_h - 1) / max_h;
      c.height = (height * c.v + max_v - 1) / max_v;
      c.blocks_w = (c.width + 7) / 8;
      c.blocks_h = (c.height + 7) / 8;
      c.alloc_w = mcus_x * c.h;
      c.alloc_h = mcus_y * c.v;
      c.coef.assign((size_t)c.alloc_w * c.alloc_h * 64, 0);
    }
    have_frame = true;
    return true;
  }

  bool Decode(LoadedTexture *tex) {
    if (d.size() < 4 || d[0] != 0xff || d[1] != 0xd8) return false;
    size_t pos = 2;
    bool done = false, any_scan = false;
    while (!done) {
      // next marker
      while (pos < d.size() && d[pos] != 0xff) pos++;
      while (pos < d.size() && d[pos] == 0xff) pos++;
      if (pos >= d.size()) break;
      const int m = d[pos++];
      if (m == 0xd9) break;
      if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
      if (pos + 2 > d.size()) return false;
      const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
      if (len < 2 || pos + len > d.size()) return false;
      switch (m) {
        case 0xc0: case 0xc1: case 0xc2:
          if (have_frame) return false;
          progressive = m == 0xc2;
          if (!ParseFrame(pos, len)) return false;
          break;
        case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
          return false;  // lossless, differential, arithmetic
        case 0xc4: {     // DHT
          size_t q = pos + 2;
          while (q < pos + len) {
            if (q + 17 > pos + len) return false;
            const int tc = d[q] >> 4, th = d[q] & 15;
            if (tc > 1 || th > 3) return false;
            int n = 0;
            for (int i = 0; i < 16; i++) n += d[q + 1 + i];
            if (n > 256 || q + 17 + n > pos + len) return false;
            (tc == 0 ? dc[th] : ac[th]).Build(&d[q + 1], &d[q + 17], n);
            q += 17 + (size_t)n;
          }
          break;
        }
        case 0xdb: {  // DQT (zigzag order in the file)
          size_t q = pos + 2;
          while (q < pos + len) {
            const int pq = d[q] >> 4, tq = d[q] & 15;
            if (tq > 3 || pq > 1 || q + 1 + (size_t)(pq ? 128 : 64) > pos + len) return false;
            for (int i = 0; i < 64; i++) {
              quant[tq][kZigzag[i]] = pq ? (uint16_t)((d[q + 1 + 2 * i] << 8) | d[q + 2 + 2 * i]) : d[q + 1 + i];
            }
            have_quant[tq] = true;
            q += 1 + (size_t)(pq ? 128 : 64);
          }
          break;
        }
        case 0xdd:
          if (len != 4) return false;
          restart_interval = (d[pos + 2] << 8) | d[pos + 3];
          break;
        case 0xe0:
          if (len >= 7 && memcmp(&d[pos + 2], "JFIF\0", 5) == 0) saw_jfif = true;
          break;
        case 0xee:
          if (len >= 14 && memcmp(&d[pos + 2], "Adobe", 5) == 0) {
            saw_adobe = true;
            adobe_transform = d[pos + 13];
          }
          break;
        case 0xda: {
          if (!have_frame) return false;
          if (!DecodeScan(&pos)) return false;
          any_scan = true;
          continue;  // pos already points at the next marker
        }
        default: break;
      }
      pos += len;
    }
    if (!have_frame || !any_scan) return false;

    // ---- coefficients -> sample planes ----
    for (int i = 0; i < ncomp; i++) {
      Component &c = comp[i];
      if (!have_quant[c.tq]) return false;
      const size_t stride = (size_t)c.alloc_w * 8;
      c.plane.assign(stride * c.alloc_h * 8, 0);
      for (int by = 0; by < c.alloc_h; by++) {
        for (int bx = 0; bx < c.alloc_w; bx++) {
          IdctIslow(&c.coef[((size_t)by * c.alloc_w + bx) * 64], quant[c.tq], &c.plane[(size_t)by * 8 * stride + (size_t)bx * 8], stride);
        }
      }
      std::vector<int16_t>().swap(c.coef);
    }

    // ---- upsampling to full resolution (jdsample.c), then colour conversion (jdcolor.c) ----
    std::vector<uint8_t> full[3];
    for (int i = 0; i < ncomp; i++) {
      Component &c = comp[i];
      const int hx = max_h / c.h, vx = max_v / c.v;
      const size_t stride = (size_t)c.alloc_w * 8;
      const int fw = c.width * hx;  // >= width
      full[i].assign((size_t)fw * height, 0);
      const bool fancy = c.width > 2;
      for (int y = 0; y < height; y++) {
        uint8_t *out = &full[i][(size_t)y * fw];
        const int r = y / vx;  // source row
        const uint8_t *in0 = &c.plane[(size_t)r * stride];
        if (hx == 1 && vx == 1) {
          memcpy(out, in0, (size_t)c.width);
        } else if (hx == 2 && vx == 1) {
          if (!fancy) {
            for (int x = 0; x < c.width; x++) out[2 * x] = out[2 * x + 1] = in0[x];
          } else {  // h2v1_fancy_upsample
            const int n = c.width;
            out[0] = in0[0];
            out[1] = (uint8_t)((in0[0] * 3 + in0[1] + 2) >> 2);
            for (int x = 1; x < n - 1; x++) {
              out[2 * x] = (uint8_t)((in0[x] * 3 + in0[x - 1] + 1) >> 2);
              out[2 * x + 1] = (uint8_t)((in0[x] * 3 + in0[x + 1] + 2) >> 2);
            }
            out[2 * n - 2] = (uint8_t)((in0[n - 1] * 3 + in0[n - 2] + 1) >> 2);
            out[2 * n - 1] = in0[n - 1];
          }
        } else if (hx == 2 && vx == 2) {
          if (!fancy) {
            for (int x = 0; x < c.width; x++) out[2 * x] = out[2 * x + 1] = in0[x];
          } else {  // h2v2_fancy_upsample: the nearer neighbour row, clamped at the plane's real first / last row
            int rn = (y & 1) == 0 ? r - 1 : r + 1;
            rn = rn < 0 ? 0 : (rn > c.height - 1 ? c.height - 1 : rn);
            const uint8_t *in1 = &c.plane[(size_t)rn * stride];
            const int n = c.width;
            int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
            out[0] = (uint8_t)((thiscol * 4 + 8) >> 4);
            out[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
            lastcol = thiscol;
            thiscol = nextcol;
            for (int x = 1; x < n - 1; x++) {
              nextcol = in0[x + 1] * 3 + in1[x + 1];
              out[2 * x] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
              out[2 * x + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
              lastcol = thiscol;
              thiscol = nextcol;
            }
            out[2 * n - 2] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
            out[2 * n - 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
          }
        } else if (hx == 1 && vx == 2) {  // libjpeg-turbo's h1v2_fancy_upsample
          int rn = (y & 1) == 0 ? r - 1 : r + 1;
          rn = rn < 0 ? 0 : (rn > c.height - 1 ? c.height - 1 : rn);
          const uint8_t *in1 = &c.plane[(size_t)rn * stride];
          const int bias = (y & 1) == 0 ? 1 : 2;
          for (int x = 0; x < c.width; x++) out[x] = (uint8_t)((in0[x] * 3 + in1[x] + bias) >> 2);
        } else {  // any other integral ratio (4:1:1 ...): int_upsample, plain replication
          for (int x = 0; x < fw; x++) out[x] = in0[x / hx];
        }
      }
      std::vector<uint8_t>().swap(c.plane);
    }
    // colour space as libjpeg deduces it (jdapimin.c: default_decompress_parms)
    bool ycc = true;
    if (ncomp == 3) {
      if (saw_jfif) {
        ycc = true;
      } else if (saw_adobe) {
        ycc = adobe_transform != 0;
      } else {
        ycc = !(comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B');
      }
    }
    tex->width = width;
    tex->height = height;
    tex->rgba.assign((size_t)width * height * 4, 255);
    const int fw0 = comp[0].width * (max_h / comp[0].h);
    for (int y = 0; y < height; y++) {
      uint8_t *dst = &tex->rgba[(size_t)y * width * 4];
      const uint8_t *p0 = &full[0][(size_t)y * fw0];
      if (ncomp == 1) {
        for (int x = 0; x < width; x++) dst[4 * x] = dst[4 * x + 1] = dst[4 * x + 2] = p0[x];
        continue;
      }
      const uint8_t *p1 = &full[1][(size_t)y * (comp[1].width * (max_h / comp[1].h))];
      const uint8_t *p2 = &full[2][(size_t)y * (comp[2].width * (max_h / comp[2].h))];
      for (int x = 0; x < width; x++) {
        if (!ycc) {
          dst[4 * x] = p0[x];
          dst[4 * x + 1] = p1[x];
          dst[4 * x + 2] = p2[x];
          continue;
        }
        const int yy = p0[x], cb = p1[x] - 128, cr = p2[x] - 128;
        // FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768
        const int r = yy + (int)((91881 * (int64_t)cr + 32768) >> 16);
        const int g = yy + (int)((-22554 * (int64_t)cb + 32768 - 46802 * (int64_t)cr) >> 16);
        const int b = yy + (int)((116130 * (int64_t)cb + 32768) >> 16);
        dst[4 * x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        dst[4 * x + 1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
        dst[4 * x + 2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
      }
    }
    return true;
  }
};

}  // namespace

bool DecodeJpeg(const std::vector<uint8_t> &d, LoadedTexture *tex) {
  JpegDecoder dec(d);
  return dec.Decode(tex);
}

}  // namespace mtb
