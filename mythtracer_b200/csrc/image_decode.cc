// Texture file decoding for the loader.  The reference hands the file to SDL2_image and
// converts whatever comes back to RGBA32 (Texture::LoadFromFile, texture.cc:60-109); only the R, G, B bytes are
// ever read afterwards (px / 255.0, texture.cc:100-104).  SDL2_image is not available offline, so the formats
// OBJ/MTL assets usually reference are decoded here, to the same RGBA32 bytes:
//   * PPM  P6, maxval 255
//   * PNG  non-interlaced or Adam7; greyscale, greyscale+alpha, RGB, RGBA, palette; 1/2/4/8 bits (16 bits keep the high byte,
//          as libpng's strip_16); no gamma / colour management (libpng's default as well)
//   * BMP  uncompressed 24 / 32 bits and 8-bit palette, bottom-up or top-down
//   * TGA  true-colour (type 2), greyscale (3) and run-length true-colour (10), 24 / 32 bits, either origin
//   * JPEG baseline / extended sequential / progressive, 8 bits, grey or three components (jpeg_decode.cc)
// Anything else fails the load, as an undecodable file does upstream (objreader.cc:467-469).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "scene_build.h"

namespace mtb {
bool DecodeJpeg(const std::vector<uint8_t> &d, LoadedTexture *tex);  // jpeg_decode.cc

namespace {

bool ReadWholeFile(const std::string &path, std::vector<uint8_t> *out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (f == nullptr) return false;
  std::vector<uint8_t> buf;
  uint8_t chunk[65536];
  size_t n;
  while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) {
    buf.insert(buf.end(), chunk, chunk + n);
    if (buf.size() > (size_t)4 << 30) break;
  }
  fclose(f);
  out->swap(buf);
  return true;
}

// the reference's sanity window (texture.cc:74-78)
bool SizeOk(int64_t w, int64_t h) { return w > 0 && h > 0 && w <= 30000 && h <= 30000; }

void SetSize(LoadedTexture *tex, int w, int h) {
  tex->width = w;
  tex->height = h;
  tex->rgba.assign((size_t)w * (size_t)h * 4, 255);
}

// ---------------------------------------------------------------------------------------------------
// PPM
// ---------------------------------------------------------------------------------------------------
bool PpmInt(const std::vector<uint8_t> &d, size_t *pos, int *out) {
  for (;;) {  // whitespace and comments
    if (*pos >= d.size()) return false;
    const uint8_t c = d[*pos];
    if (c == '#') {
      while (*pos < d.size() && d[*pos] != '\n') (*pos)++;
    } else if (c == ' ' || c == '\t' || c == '\r' || c == '\n') {
      (*pos)++;
    } else {
      break;
    }
  }
  int64_t v = 0;
  bool any = false;
  while (*pos < d.size() && d[*pos] >= '0' && d[*pos] <= '9') {
    v = v * 10 + (d[*pos] - '0');
    if (v > 100000000) return false;
    any = true;
    (*pos)++;
  }
  *out = (int)v;
  return any;
}

bool DecodePpm(const std::vector<uint8_t> &d, LoadedTexture *tex) {
  size_t pos = 2;
  int w = 0, h = 0, maxval = 0;
  if (!PpmInt(d, &pos, &w) || !PpmInt(d, &pos, &h) || !PpmInt(d, &pos, &maxval) || maxval != 255) return false;
  if (!SizeOk(w, h) || pos >= d.size()) return false;
  pos++;  // the single whitespace byte after maxval
  const size_t n = (size_t)w * (size_t)h;
  if (d.size() - pos < n * 3) return false;
  SetSize(tex, w, h);
  for (size_t i = 0; i < n; i++) {
    tex->rgba[i * 4 + 0] = d[pos + i * 3 + 0];
    tex->rgba[i * 4 + 1] = d[pos + i * 3 + 1];
    tex->rgba[i * 4 + 2] = d[pos + i * 3 + 2];
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// inflate (RFC 1951) for PNG's zlib stream
// ---------------------------------------------------------------------------------------------------
struct BitReader {
  const uint8_t *p;
  size_t n, pos = 0;
  uint32_t bits = 0;
  int count = 0;
  bool ok = true;
  uint32_t Get(int k) {
    while (count < k) {
      if (pos >= n) {
        ok = false;
        return 0;
      }
      bits |= (uint32_t)p[pos++] << count;
      count += 8;
    }
    const uint32_t v = bits & ((k == 32) ? 0xffffffffu : ((1u << k) - 1u));
    bits = k == 32 ? 0 : bits >> k;
    count -= k;
    return v;
  }
};

struct Huffman {
  uint16_t count[16];
  uint16_t symbol[320];
  bool Build(const uint8_t *lengths, int n) {
    memset(count, 0, sizeof(count));
    for (int i = 0; i < n; i++) count[lengths[i]]++;
    if (count[0] == n) return true;  // no codes: legal for an unused distance tree
    int left = 1;
    for (int len = 1; len < 16; len++) {
      left <<= 1;
      left -= count[len];
      if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + count[len];
    for (int i = 0; i < n; i++) {
      if (lengths[i] != 0) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    return true;
  }
  int Decode(BitReader *br) const {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; len++) {
      code |= (int)br->Get(1);
      if (!br->ok) return -1;
      const int c = count[len];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};

bool Inflate(const uint8_t *src, size_t n, std::vector<uint8_t> *out, size_t expect) {
  static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint16_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint16_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t kOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  BitReader br{src, n};
  out->clear();
  out->reserve(expect);
  for (;;) {
    const uint32_t last = br.Get(1), type = br.Get(2);
    if (!br.ok) return false;
    if (type == 0) {
      br.bits = 0;
      br.count = 0;
      if (br.pos + 4 > n) return false;
      const uint32_t len = src[br.pos] | (src[br.pos + 1] << 8), nlen = src[br.pos + 2] | (src[br.pos + 3] << 8);
      br.pos += 4;
      if ((len ^ 0xffffu) != nlen || br.pos + len > n) return false;
      out->insert(out->end(), src + br.pos, src + br.pos + len);
      br.pos += len;
    } else if (type == 1 || type == 2) {
      Huffman lit, dist;
      uint8_t lengths[320];
      if (type == 1) {
        for (int i = 0; i < 144; i++) lengths[i] = 8;
        for (int i = 144; i < 256; i++) lengths[i] = 9;
        for (int i = 256; i < 280; i++) lengths[i] = 7;
        for (int i = 280; i < 288; i++) lengths[i] = 8;
        if (!lit.Build(lengths, 288)) return false;
        for (int i = 0; i < 30; i++) lengths[i] = 5;
        if (!dist.Build(lengths, 30)) return false;
      } else {
        const int nlen = (int)br.Get(5) + 257, ndist = (int)br.Get(5) + 1, ncode = (int)br.Get(4) + 4;
        if (!br.ok || nlen > 286 || ndist > 30) return false;
        uint8_t cl[19];
        memset(cl, 0, sizeof(cl));
        for (int i = 0; i < ncode; i++) cl[kOrder[i]] = (uint8_t)br.Get(3);
        Huffman lencode;
        if (!br.ok || !lencode.Build(cl, 19)) return false;
        int idx = 0;
        while (idx < nlen + ndist) {
          const int sym = lencode.Decode(&br);
          if (sym < 0) return false;
          if (sym < 16) {
            lengths[idx++] = (uint8_t)sym;
          } else {
            int rep, val = 0;
            if (sym == 16) {
              if (idx == 0) return false;
              val = lengths[idx - 1];
              rep = 3 + (int)br.Get(2);
            } else if (sym == 17) {
              rep = 3 + (int)br.Get(3);
            } else {
              rep = 11 + (int)br.Get(7);
            }
            if (!br.ok || idx + rep > nlen + ndist) return false;
            while (rep-- > 0) lengths[idx++] = (uint8_t)val;
          }
        }
        if (lengths[256] == 0) return false;
        if (!lit.Build(lengths, nlen) || !dist.Build(lengths + nlen, ndist)) return false;
      }
      for (;;) {
        const int sym = lit.Decode(&br);
        if (sym < 0) return false;
        if (sym < 256) {
          out->push_back((uint8_t)sym);
        } else if (sym == 256) {
          break;
        } else {
          const int li = sym - 257;
          if (li >= 29) return false;
          const int len = kLenBase[li] + (int)br.Get(kLenExtra[li]);
          const int ds = dist.Decode(&br);
          if (ds < 0 || ds >= 30) return false;
          const size_t d = kDistBase[ds] + br.Get(kDistExtra[ds]);
          if (!br.ok || d > out->size()) return false;
          const size_t from = out->size() - d;
          for (int i = 0; i < len; i++) out->push_back((*out)[from + i]);
        }
        if (out->size() > expect + 65536) return false;  // more data than the image can hold
      }
    } else {
      return false;
    }
    if (last) break;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// PNG
// ---------------------------------------------------------------------------------------------------
uint32_t Be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool DecodePng(const std::vector<uint8_t> &d, LoadedTexture *tex) {
  size_t pos = 8;
  uint32_t w = 0, h = 0;
  int depth = 0, color = 0, interlace = 0;
  std::vector<uint8_t> idat, palette;
  bool have_ihdr = false, done = false;
  while (!done && pos + 12 <= d.size()) {
    const uint32_t len = Be32(&d[pos]);
    const uint8_t *type = &d[pos + 4];
    if (len > d.size() || pos + 12 + len > d.size()) return false;
    const uint8_t *body = &d[pos + 8];
    if (memcmp(type, "IHDR", 4) == 0 && len >= 13) {
      w = Be32(body);
      h = Be32(body + 4);
      depth = body[8];
      color = body[9];
      interlace = body[12];
      have_ihdr = true;
    } else if (memcmp(type, "PLTE", 4) == 0) {
      palette.assign(body, body + len);
    } else if (memcmp(type, "IDAT", 4) == 0) {
      idat.insert(idat.end(), body, body + len);
    } else if (memcmp(type, "IEND", 4) == 0) {
      done = true;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || !SizeOk(w, h) || idat.size() < 6) return false;
  int channels;
  switch (color) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return false;
  }
  if (!(depth == 8 || depth == 16 || ((color == 0 || color == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
  if (color == 3 && (depth == 16 || palette.size() < 3)) return false;
  const size_t bpp_bits = (size_t)channels * depth;
  const size_t bpp = (bpp_bits + 7) / 8;  // filter unit, at least one byte
  // The image arrives as one pass (non-interlaced) or as the seven Adam7 passes, each a complete filtered sub-image
  // of the pixels (x0 + i dx, y0 + j dy); a pass without pixels has no bytes at all.
  struct Pass {
    uint32_t x0, y0, dx, dy;
  };
  static const Pass kAdam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
  static const Pass kWhole[1] = {{0, 0, 1, 1}};
  if (interlace > 1) return false;
  const Pass *passes = interlace ? kAdam7 : kWhole;
  const int n_passes = interlace ? 7 : 1;
  size_t expect = 0;
  for (int k = 0; k < n_passes; k++) {
    const Pass &ps = passes[k];
    if (w <= ps.x0 || h <= ps.y0) continue;
    const size_t pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
    expect += ((pw * bpp_bits + 7) / 8 + 1) * ph;
  }
  std::vector<uint8_t> raw;
  // zlib header (2 bytes) + deflate stream + adler32 (4 bytes, not checked)
  if ((idat[0] & 0x0f) != 8 || (idat[1] & 0x20) != 0) return false;
  if (!Inflate(idat.data() + 2, idat.size() - 2, &raw, expect)) return false;
  if (raw.size() < expect) return false;
  SetSize(tex, (int)w, (int)h);
  size_t at = 0;
  for (int k = 0; k < n_passes; k++) {
    const Pass &ps = passes[k];
    if (w <= ps.x0 || h <= ps.y0) continue;
    const size_t pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
    const size_t stride = (pw * bpp_bits + 7) / 8;
    // undo the scanline filters in place (the row above the first row of a pass is all zero)
    std::vector<uint8_t> prev(stride, 0);
    for (size_t py = 0; py < ph; py++) {
      uint8_t *row = &raw[at + py * (stride + 1) + 1];
      const uint8_t filter = row[-1];
      for (size_t i = 0; i < stride; i++) {
        const int a = i >= bpp ? row[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
        int pred;
        switch (filter) {
          case 0: pred = 0; break;
          case 1: pred = a; break;
          case 2: pred = b; break;
          case 3: pred = (a + b) >> 1; break;
          case 4: {
            const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
            pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            break;
          }
          default: return false;
        }
        row[i] = (uint8_t)(row[i] + pred);
      }
      memcpy(prev.data(), row, stride);
      const size_t y = ps.y0 + py * ps.dy;
      for (size_t px_i = 0; px_i < pw; px_i++) {
        uint8_t *dst = &tex->rgba[(y * w + ps.x0 + px_i * ps.dx) * 4];
        uint8_t v[4] = {0, 0, 0, 255};
        if (depth >= 8) {
          const size_t step = depth / 8;  // 16 bits: the high (first) byte, like png_set_strip_16
          const uint8_t *px = row + px_i * channels * step;
          for (int ch = 0; ch < channels; ch++) v[ch] = px[ch * step];
        } else {
          const size_t bit = px_i * depth;
          const int sample = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
          v[0] = (uint8_t)(color == 3 ? sample : sample * 255 / ((1 << depth) - 1));
        }
        if (color == 3) {
          const size_t idx = (size_t)v[0] * 3;
          if (idx + 3 > palette.size()) return false;
          dst[0] = palette[idx];
          dst[1] = palette[idx + 1];
          dst[2] = palette[idx + 2];
        } else if (color == 0 || color == 4) {
          dst[0] = dst[1] = dst[2] = v[0];
          if (color == 4) dst[3] = v[1];
        } else {
          dst[0] = v[0];
          dst[1] = v[1];
          dst[2] = v[2];
          if (color == 6) dst[3] = v[3];
        }
      }
    }
    at += (stride + 1) * ph;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// BMP
// ---------------------------------------------------------------------------------------------------
uint32_t Le32(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t Le16(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8); }

bool DecodeBmp(const std::vector<uint8_t> &d, LoadedTexture *tex) {
  if (d.size() < 54) return false;
  const uint32_t data_off = Le32(&d[10]), hdr = Le32(&d[14]);
  if (hdr < 40) return false;
  const int32_t w = (int32_t)Le32(&d[18]);
  int32_t h = (int32_t)Le32(&d[22]);
  const uint32_t bits = Le16(&d[28]), comp = Le32(&d[30]);
  const bool top_down = h < 0;
  if (top_down) h = -h;
  if (!SizeOk(w, h) || !(comp == 0 || (comp == 3 && bits == 32)) || !(bits == 8 || bits == 24 || bits == 32)) return false;
  const size_t stride = (((size_t)w * bits + 31) / 32) * 4;
  if ((size_t)data_off + stride * (size_t)h > d.size()) return false;
  const uint8_t *pal = &d[14 + hdr];
  uint32_t n_pal = Le32(&d[46]);
  if (bits == 8) {
    if (n_pal == 0) n_pal = 256;
    if ((size_t)14 + hdr + (size_t)n_pal * 4 > d.size()) return false;
  }
  SetSize(tex, w, h);
  for (int32_t y = 0; y < h; y++) {
    const uint8_t *row = &d[data_off + stride * (size_t)(top_down ? y : h - 1 - y)];
    uint8_t *dst = &tex->rgba[(size_t)y * w * 4];
    for (int32_t x = 0; x < w; x++) {
      if (bits == 8) {
        const uint32_t idx = row[x];
        if (idx >= n_pal) return false;
        dst[x * 4 + 0] = pal[idx * 4 + 2];
        dst[x * 4 + 1] = pal[idx * 4 + 1];
        dst[x * 4 + 2] = pal[idx * 4 + 0];
      } else {
        const uint8_t *px = row + (size_t)x * (bits / 8);
        dst[x * 4 + 0] = px[2];
        dst[x * 4 + 1] = px[1];
        dst[x * 4 + 2] = px[0];
      }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// TGA
// ---------------------------------------------------------------------------------------------------
bool DecodeTga(const std::vector<uint8_t> &d, LoadedTexture *tex) {
  if (d.size() < 18) return false;
  const int id_len = d[0], cmap_type = d[1], type = d[2];
  const int w = (int)Le16(&d[12]), h = (int)Le16(&d[14]), bits = d[16], desc = d[17];
  if (cmap_type != 0 || !(type == 2 || type == 3 || type == 10)) return false;
  if (!SizeOk(w, h)) return false;
  if (!((type == 3 && bits == 8) || (type != 3 && (bits == 24 || bits == 32)))) return false;
  const int bpp = bits / 8;
  size_t pos = 18 + (size_t)id_len;
  const size_t n = (size_t)w * (size_t)h;
  std::vector<uint8_t> px(n * (size_t)bpp);
  if (type == 10) {
    size_t out = 0;
    while (out < n) {
      if (pos >= d.size()) return false;
      const int head = d[pos++];
      const size_t run = (size_t)(head & 127) + 1;
      if (out + run > n) return false;
      if (head & 128) {
        if (pos + (size_t)bpp > d.size()) return false;
        for (size_t i = 0; i < run; i++) memcpy(&px[(out + i) * bpp], &d[pos], (size_t)bpp);
        pos += (size_t)bpp;
      } else {
        if (pos + run * bpp > d.size()) return false;
        memcpy(&px[out * bpp], &d[pos], run * bpp);
        pos += run * bpp;
      }
      out += run;
    }
  } else {
    if (pos + px.size() > d.size()) return false;
    memcpy(px.data(), &d[pos], px.size());
  }
  const bool top = (desc & 0x20) != 0, right = (desc & 0x10) != 0;
  SetSize(tex, w, h);
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      const uint8_t *s = &px[((size_t)(top ? y : h - 1 - y) * w + (size_t)(right ? w - 1 - x : x)) * bpp];
      uint8_t *dst = &tex->rgba[((size_t)y * w + x) * 4];
      if (bpp == 1) {
        dst[0] = dst[1] = dst[2] = s[0];
      } else {
        dst[0] = s[2];
        dst[1] = s[1];
        dst[2] = s[0];
        if (bpp == 4) dst[3] = s[3];
      }
    }
  }
  return true;
}

bool EndsWithNoCase(const std::string &s, const char *suffix) {
  const size_t n = strlen(suffix);
  if (s.size() < n) return false;
  for (size_t i = 0; i < n; i++) {
    char c = s[s.size() - n + i];
    if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    if (c != suffix[i]) return false;
  }
  return true;
}

}  // namespace

bool DecodeImageFile(const std::string &path, LoadedTexture *tex) {
  std::vector<uint8_t> d;
  if (!ReadWholeFile(path, &d) || d.size() < 4) return false;
  static const uint8_t kPngMagic[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (d.size() >= 8 && memcmp(d.data(), kPngMagic, 8) == 0) return DecodePng(d, tex);
  if (d[0] == 0xff && d[1] == 0xd8 && d[2] == 0xff) return DecodeJpeg(d, tex);
  if (d[0] == 'P' && d[1] == '6') return DecodePpm(d, tex);
  if (d[0] == 'B' && d[1] == 'M') return DecodeBmp(d, tex);
  if (EndsWithNoCase(path, ".tga")) return DecodeTga(d, tex);  // TGA has no magic number
  return false;
}

}  // namespace mtb
