// Host-side OBJ + MTL (+ texture) reader with the observable behaviour of the reference loader
// (reference VerStarting/objreader.cc), so that the same files give the same triangles, in the same
// order, with the same debug line numbers:
//   * lines are consumed in 127-byte pieces (char line[128] + fgets, objreader.cc:233-235) and every piece
//     advances the 0-based line counter (objreader.cc:210,233);
//   * the text after the LAST '\r' / '\n' is cut (objreader.cc:239-247);
//   * a face keeps a token only when more input follows it (`s >> token; if (s.eof()) break;`,
//     objreader.cc:111-115): "f 1 2 3" loses its third vertex and is rejected, "f 1 2 3 " is a triangle;
//   * face indices are read with %i (objreader.cc:117-125), 1-based, no relative indices;
//   * quads become (0,1,2) and (2,3,0) with the quad's line number on both (objreader.cc:141-151,180);
//   * normals / texture coordinates are taken only when all three indices are present (objreader.cc:159-174);
//   * `usemtl` of an unknown name selects "no material" and parsing goes on (objreader.cc:85-90);
//   * MTL keys read: newmtl Ka Kd Ks Ns Ni Tr Tf Refl map_Ka; d illum Ke map_Kd are ignored
//     (objreader.cc:487-503); a texture that fails to load fails the whole load (objreader.cc:467-469).
// Texture files are decoded by image_decode.cc (PPM, PNG, BMP, TGA -> RGBA32): SDL2_image, which the reference
// uses (texture.cc:60-109), is not available offline.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "scene_build.h"

namespace mtb {
namespace {

struct FileCloser {
  FILE *f;
  ~FileCloser() {
    if (f != nullptr) fclose(f);
  }
};

std::string DirOf(const std::string &path) {
  const size_t pos = path.find_last_of("/\\");
  return pos == std::string::npos ? std::string() : path.substr(0, pos);
}

std::string Join(const std::string &dir, const char *name) { return dir.empty() ? std::string(name) : dir + "/" + name; }

void ChopLineEnd(char *line) {
  char *p = strrchr(line, '\r');
  if (p != nullptr) *p = '\0';
  p = strrchr(line, '\n');
  if (p != nullptr) *p = '\0';
}

// sscanf("%lf") == strtod after optional white space; these helpers avoid the format interpreter on the
// millions of `v` / `vn` / `vt` / `f` lines of a big model while keeping sscanf's accept / reject behaviour.
inline bool ScanDouble(const char **p, double *out) {
  char *end = nullptr;
  const double v = strtod(*p, &end);
  if (end == *p) return false;
  *out = v;
  *p = end;
  return true;
}

// sscanf("%i"): optional sign, 0x / 0 prefixes select base 16 / 8 (objreader.cc:117-125 reads indices so)
inline bool ScanInt(const char **p, int *out) {
  char *end = nullptr;
  const long v = strtol(*p, &end, 0);
  if (end == *p) return false;
  *out = (int)v;
  *p = end;
  return true;
}

// The reference tries "%i/%i/%i", "%i//%i", "%i/%i", "%i" in turn on one face token (objreader.cc:117-125).
// Same outcome in one pass: v is mandatory; "v/vt/vn", "v//vn", "v/vt" fill what parses; anything else keeps v.
inline bool ScanFaceToken(const char *tok, int *v, int *vt, int *vn) {
  *v = *vt = *vn = 0;
  const char *p = tok;
  if (!ScanInt(&p, v)) return false;
  if (*p != '/') return true;
  p++;
  if (*p == '/') {  // v//vn
    p++;
    int n;
    if (ScanInt(&p, &n)) *vn = n;
    return true;
  }
  int t;
  if (!ScanInt(&p, &t)) return true;  // "v/junk": only "%i" matches
  *vt = t;
  if (*p == '/') {
    p++;
    int n;
    if (ScanInt(&p, &n)) *vn = n;
  }
  return true;
}

inline bool IsSpace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }



class MtlParser {
 public:
  MtlParser(LoadedScene *scene, std::string *err) : scene_(scene), err_(err) {}

  bool Parse(const std::string &path) {
    FileCloser fc{fopen(path.c_str(), "r")};
    if (fc.f == nullptr) {
      *err_ = "file \"" + path + "\" not found";
      return false;
    }
    dir_ = DirOf(path);
    char line[128];
    while (fgets(line, sizeof(line), fc.f) != nullptr) {
      ChopLineEnd(line);
      char key[16] = {0};
      if (sscanf(line, "%15s", key) != 1 || key[0] == '#') continue;
      if (!Handle(key, line)) return false;
    }
    return true;
  }

 private:
  bool NeedCurrent() {
    if (cur_ >= 0) return true;
    *err_ = "material not ready; missing newmtl";
    return false;
  }

  bool Triple(const char *line, const char *fmt, double *dst, const char *what) {
    if (!NeedCurrent()) return false;
    double r, g, b;
    if (sscanf(line, fmt, &r, &g, &b) != 3) {
      *err_ = std::string("unsupported ") + what + " format \"" + line + "\"";
      return false;
    }
    dst[0] = r;
    dst[1] = g;
    dst[2] = b;
    return true;
  }

  bool Single(const char *line, const char *fmt, double *dst, const char *what) {
    if (!NeedCurrent()) return false;
    double v;
    if (sscanf(line, fmt, &v) != 1) {
      *err_ = std::string("unsupported ") + what + " format \"" + line + "\"";
      return false;
    }
    *dst = v;
    return true;
  }

  bool Handle(const std::string &key, const char *line) {
    if (key == "newmtl") {
      char name[128];
      if (sscanf(line, "newmtl %127s", name) != 1) {
        *err_ = "unsupported newmtl format";
        return false;
      }
      // materials[name] = ... (objreader.cc:280): a repeated name replaces the earlier definition.
      int found = -1;
      for (size_t i = 0; i < scene_->material_names.size(); i++) {
        if (scene_->material_names[i] == name) found = (int)i;
      }
      if (found < 0) {
        scene_->material_names.push_back(name);
        scene_->materials.emplace_back();
        found = (int)scene_->materials.size() - 1;
      }
      mtb_material fresh;
      memset(&fresh, 0, sizeof(fresh));
      fresh.texture = -1;
      scene_->materials[(size_t)found] = fresh;
      cur_ = found;
      return true;
    }
    if (key == "d" || key == "illum" || key == "Ke" || key == "map_Kd") return true;
    if (key == "Ka") return Triple(line, " Ka %lf %lf %lf", scene_->materials[Cur()].ambient, "Ka");
    if (key == "Kd") return Triple(line, " Kd %lf %lf %lf", scene_->materials[Cur()].diffuse, "Kd");
    if (key == "Ks") return Triple(line, " Ks %lf %lf %lf", scene_->materials[Cur()].specular, "Ks");
    if (key == "Tf") return Triple(line, " Tf %lf %lf %lf", scene_->materials[Cur()].transmission_filter, "Tf");
    if (key == "Ns") return Single(line, " Ns %lf", &scene_->materials[Cur()].specular_exp, "Ns");
    if (key == "Ni") return Single(line, " Ni %lf", &scene_->materials[Cur()].refraction_index, "Ni");
    if (key == "Tr") return Single(line, " Tr %lf", &scene_->materials[Cur()].transparency, "Tr");
    if (key == "Refl") return Single(line, " Refl %lf", &scene_->materials[Cur()].reflectance, "Refl");
    if (key == "map_Ka") {
      if (!NeedCurrent()) return false;
      char fname[256];
      if (sscanf(line, " map_Ka %255[^\n]", fname) != 1) {
        *err_ = std::string("unsupported map_ka format \"") + line + "\"";
        return false;
      }
      int tex = -1;
      for (size_t i = 0; i < scene_->texture_names.size(); i++) {
        if (scene_->texture_names[i] == fname) tex = (int)i;
      }
      if (tex < 0) {
        LoadedTexture t;
        if (!DecodeImageFile(Join(dir_, fname), &t)) {
          *err_ = std::string("cannot load texture \"") + fname + "\"";
          return false;
        }
        scene_->textures.push_back(std::move(t));
        scene_->texture_names.push_back(fname);
        tex = (int)scene_->textures.size() - 1;
      }
      scene_->materials[Cur()].texture = tex;
      return true;
    }
    fprintf(stderr, "warning: unknown MTL feature \"%s\"\n", key.c_str());
    return true;
  }

  size_t Cur() const { return cur_ < 0 ? 0 : (size_t)cur_; }

  LoadedScene *scene_;
  std::string *err_;
  std::string dir_;
  int cur_ = -1;
};

}  // namespace

bool LoadMtlFile(const char *path, LoadedScene *scene, std::string *err) {
  MtlParser mp(scene, err);
  return mp.Parse(path);
}

// ---------------------------------------------------------------------------------------------------
// OBJ reader.  The reference reads line by line (objreader.cc:201-274); what it observably does is kept (see the
// list at the top of this file), but the work is organised for a many-core host - the parse was 0.35 s of the 0.61 s
// a 500 k-triangle scene needed to reach its first frame in round 1:
//   phase 1 (parallel): the file is cut at newline boundaries into chunks; every chunk is consumed in the same
//            127-byte pieces fgets would deliver (a piece ends at a newline, so every chunk start is a piece start),
//            and parsed into chunk-local vertex / normal / texcoord arrays, face records and an ordered list of
//            events (usemtl, mtllib, warnings, the first error);
//   phase 2 (sequential, tiny): chunks in file order - running piece counts give the 0-based line numbers
//            (objreader.cc:210,233), events resolve materials with the names known AT THAT POINT of the file, the
//            first error in file order ends the load;
//   phase 3 (parallel): triangles are assembled into their final positions; an index is valid only against the
//            vertices defined BEFORE its face, as in a sequential read.
// ---------------------------------------------------------------------------------------------------
namespace {

struct FaceRec {
  int32_t vi[5], ti[5], ni[5];
  int32_t count;               // 3, or 5 for a quad (0,1,2),(2,3,0)
  int32_t piece;               // piece index inside the chunk
  int32_t n_pos, n_nrm, n_tex; // chunk-local element counts in front of this face
  int32_t material;            // filled in phase 2
};

enum EventKind { kEvUseMtl, kEvMtlLib, kEvWarnFeature, kEvError };
struct ObjEvent {
  EventKind kind;
  std::string text;
  int32_t face_index;  // faces of the chunk in front of this event
  int32_t piece;
};

struct ChunkOut {
  std::vector<double> pos, nrm, tex;  // xyz triples
  std::vector<FaceRec> faces;
  std::vector<ObjEvent> events;
  int32_t n_pieces = 0;
  int32_t n_tris = 0;
  // phase 3
  int32_t bad_face = -1;
  std::string bad_text;
};

// One piece (what fgets(line, 128) would have delivered, NUL-terminated, line end chopped).  Returns false on an
// error (an error event has been recorded).
bool ParsePiece(char *line, int32_t piece, ChunkOut *out) {
  char key_buf[16] = {0};
  {
    const char *q = line;
    while (IsSpace(*q)) q++;
    int n = 0;
    while (*q != '\0' && !IsSpace(*q) && n < 15) key_buf[n++] = *q++;
    if (n == 0 || key_buf[0] == '#') return true;
  }
  auto fail = [&](const std::string &text) {
    out->events.push_back(ObjEvent{kEvError, text, (int32_t)out->faces.size(), piece});
    return false;
  };
  const std::string key(key_buf);
  if (key == "v" || key == "vn") {
    double x, y, z;
    // sscanf(line, "v %lf %lf %lf"): the literal must start the line (no leading blanks), then three doubles
    const char *p = line + key.size();
    if (strncmp(line, key.c_str(), key.size()) != 0 || !ScanDouble(&p, &x) || !ScanDouble(&p, &y) || !ScanDouble(&p, &z)) {
      return fail(std::string("unsupported ") + (key == "v" ? "vertex" : "normal") + " format \"" + line + "\"");
    }
    std::vector<double> &dst = key == "v" ? out->pos : out->nrm;
    dst.push_back(x);
    dst.push_back(y);
    dst.push_back(z);
  } else if (key == "vt") {
    double u, v, w = 0.0;
    const char *p = line + 2;
    if (strncmp(line, "vt", 2) != 0 || !ScanDouble(&p, &u) || !ScanDouble(&p, &v)) {
      return fail(std::string("unsupported texcoord format \"") + line + "\"");
    }
    ScanDouble(&p, &w);  // optional third coordinate
    out->tex.push_back(u);
    out->tex.push_back(v);
    out->tex.push_back(w);
  } else if (key == "mtllib") {
    char fname[256];
    if (sscanf(line, "mtllib %255[^\n]", fname) != 1) return fail(std::string("unsupported mtllib format \"") + line + "\"");
    out->events.push_back(ObjEvent{kEvMtlLib, fname, (int32_t)out->faces.size(), piece});
  } else if (key == "usemtl") {
    char name[128];
    if (sscanf(line, "usemtl %127s", name) != 1) return fail("unsupported usemtl format");
    out->events.push_back(ObjEvent{kEvUseMtl, name, (int32_t)out->faces.size(), piece});
  } else if (key == "f") {
    // `s >> token; if (s.eof()) break;` keeps a token only when at least one more character follows it
    // (objreader.cc:109-115): the first token ("f") is skipped, a token that ends the line is dropped.
    FaceRec fr;
    int count = 0;
    const char *p = line;
    while (IsSpace(*p)) p++;
    while (*p != '\0' && !IsSpace(*p)) p++;  // the "f" itself
    for (;;) {
      while (IsSpace(*p)) p++;
      if (*p == '\0') break;
      const char *start = p;
      while (*p != '\0' && !IsSpace(*p)) p++;
      if (*p == '\0') break;  // the quirk: nothing follows this token
      char tokbuf[128];
      const size_t len = (size_t)(p - start);
      memcpy(tokbuf, start, len);
      tokbuf[len] = '\0';
      int v = 0, vt = 0, vn = 0;
      if (!ScanFaceToken(tokbuf, &v, &vt, &vn)) return fail(std::string("unsupported face format \"") + tokbuf + "\"");
      if (count < 4) {
        fr.vi[count] = v - 1;
        fr.ti[count] = vt - 1;
        fr.ni[count] = vn - 1;
      }
      count++;
    }
    if (count != 3 && count != 4) return fail("unsupported face count (" + std::to_string(count) + ")\n  " + line);
    if (count == 4) {
      fr.vi[4] = fr.vi[0];
      fr.ti[4] = fr.ti[0];
      fr.ni[4] = fr.ni[0];
      count = 5;
    }
    fr.count = count;
    fr.piece = piece;
    fr.n_pos = (int32_t)(out->pos.size() / 3);
    fr.n_nrm = (int32_t)(out->nrm.size() / 3);
    fr.n_tex = (int32_t)(out->tex.size() / 3);
    fr.material = -1;
    out->faces.push_back(fr);
    out->n_tris += count == 5 ? 2 : 1;
  } else if (key == "s" || key == "g" || key == "o") {
    return true;
  } else {
    out->events.push_back(ObjEvent{kEvWarnFeature, key_buf, (int32_t)out->faces.size(), piece});
  }
  return true;
}

// Consumes data[begin, end) the way `while (fgets(line, 128, f))` does: pieces of at most 127 bytes, a piece ends
// behind a newline; the C string a piece holds ends at its first NUL byte.
void ParseChunk(const char *data, size_t begin, size_t end, ChunkOut *out) {
  size_t p = begin;
  char line[128];
  while (p < end) {
    size_t len = end - p < 127 ? end - p : 127;
    const void *nl = memchr(data + p, '\n', len);
    if (nl != nullptr) len = (size_t)(static_cast<const char *>(nl) - (data + p)) + 1;
    memcpy(line, data + p, len);
    line[len] = '\0';
    p += len;
    const int32_t piece = out->n_pieces++;
    ChopLineEnd(line);
    if (!ParsePiece(line, piece, out)) return;  // the sequential reader stops at its first error
  }
}

}  // namespace

bool LoadObjFile(const char *path, LoadedScene *scene, std::string *err) {
  std::vector<char> data;
  {
    FileCloser fc{fopen(path, "rb")};
    if (fc.f == nullptr) {
      *err = std::string("file \"") + path + "\" not found";
      return false;
    }
    char buf[1 << 16];
    size_t got;
    if (fseek(fc.f, 0, SEEK_END) == 0) {
      const long size = ftell(fc.f);
      if (size > 0) data.reserve((size_t)size);
      fseek(fc.f, 0, SEEK_SET);
    }
    while ((got = fread(buf, 1, sizeof(buf), fc.f)) > 0) data.insert(data.end(), buf, buf + got);
  }
  const std::string dir = DirOf(path);
  const size_t size = data.size();

  // ---- phase 1: chunks that start right behind a newline ----
  unsigned n_threads = std::thread::hardware_concurrency();
  if (n_threads == 0) n_threads = 1;
  if (n_threads > 32) n_threads = 32;
  if (size < (1u << 20)) n_threads = 1;
  if (const char *env = getenv("MTB_LOADER_THREADS")) {  // tests: the result must not depend on the chunking
    if (atoi(env) >= 1) n_threads = (unsigned)std::min(atoi(env), 64);
  }
  const size_t n_chunks = n_threads == 1 ? 1 : (size_t)n_threads * 4;
  std::vector<size_t> cut(n_chunks + 1, size);
  cut[0] = 0;
  for (size_t c = 1; c < n_chunks; c++) {
    size_t p = size / n_chunks * c;
    if (p < cut[c - 1]) p = cut[c - 1];
    const void *nl = p < size ? memchr(data.data() + p, '\n', size - p) : nullptr;
    cut[c] = nl != nullptr ? (size_t)(static_cast<const char *>(nl) - data.data()) + 1 : size;
  }
  std::vector<ChunkOut> chunks(n_chunks);
  {
    std::atomic<size_t> next(0);
    auto worker = [&]() {
      for (;;) {
        const size_t c = next.fetch_add(1);
        if (c >= n_chunks) return;
        ParseChunk(data.data(), cut[c], cut[c + 1], &chunks[c]);
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; t++) pool.emplace_back(worker);
    worker();
    for (std::thread &t : pool) t.join();
  }

  // ---- phase 2: file order ----
  std::vector<int64_t> pos_base(n_chunks + 1, 0), nrm_base(n_chunks + 1, 0), tex_base(n_chunks + 1, 0), line_base(n_chunks + 1, 0),
      tri_base(n_chunks + 1, 0);
  int material = -1;
  size_t used_chunks = n_chunks;
  bool failed = false;
  std::string fail_text;
  for (size_t c = 0; c < n_chunks && !failed; c++) {
    ChunkOut &ch = chunks[c];
    size_t next_event = 0;
    for (size_t f = 0; f <= ch.faces.size(); f++) {
      while (next_event < ch.events.size() && (size_t)ch.events[next_event].face_index == f) {
        const ObjEvent &ev = ch.events[next_event++];
        if (ev.kind == kEvUseMtl) {
          material = -1;
          for (size_t i = 0; i < scene->material_names.size(); i++) {
            if (scene->material_names[i] == ev.text) material = (int)i;
          }
          if (material < 0) fprintf(stderr, "warning: material \"%s\" not found\n", ev.text.c_str());
        } else if (ev.kind == kEvMtlLib) {
          MtlParser mp(scene, err);
          if (!mp.Parse(Join(dir, ev.text.c_str()))) {
            failed = true;
            fail_text = *err;
          }
        } else if (ev.kind == kEvWarnFeature) {
          fprintf(stderr, "warning: unknown OBJ feature \"%s\"\n", ev.text.c_str());
        } else {
          failed = true;
          fail_text = ev.text;
        }
        if (failed) break;
      }
      if (failed) {
        // everything in front of the failing piece still counts (an out-of-range index there is the earlier error)
        ch.faces.resize(f);
        ch.n_tris = 0;
        for (const FaceRec &fr : ch.faces) ch.n_tris += fr.count == 5 ? 2 : 1;
        break;
      }
      if (f < ch.faces.size()) ch.faces[f].material = material;
    }
    pos_base[c + 1] = pos_base[c] + (int64_t)ch.pos.size();
    nrm_base[c + 1] = nrm_base[c] + (int64_t)ch.nrm.size();
    tex_base[c + 1] = tex_base[c] + (int64_t)ch.tex.size();
    line_base[c + 1] = line_base[c] + ch.n_pieces;
    tri_base[c + 1] = tri_base[c] + ch.n_tris;
    if (failed) used_chunks = c + 1;
  }

  // ---- phase 3: global attribute arrays, then the triangles in their final places ----
  std::vector<double> pos((size_t)pos_base[used_chunks]), nrm((size_t)nrm_base[used_chunks]), tex((size_t)tex_base[used_chunks]);
  const size_t tri0 = scene->triangles.size();
  scene->triangles.resize(tri0 + (size_t)tri_base[used_chunks]);
  {
    std::atomic<size_t> next(0);
    auto gather = [&]() {
      for (;;) {
        const size_t c = next.fetch_add(1);
        if (c >= used_chunks) return;
        const ChunkOut &ch = chunks[c];
        if (!ch.pos.empty()) memcpy(&pos[(size_t)pos_base[c]], ch.pos.data(), ch.pos.size() * sizeof(double));
        if (!ch.nrm.empty()) memcpy(&nrm[(size_t)nrm_base[c]], ch.nrm.data(), ch.nrm.size() * sizeof(double));
        if (!ch.tex.empty()) memcpy(&tex[(size_t)tex_base[c]], ch.tex.data(), ch.tex.size() * sizeof(double));
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; t++) pool.emplace_back(gather);
    gather();
    for (std::thread &t : pool) t.join();
  }
  {
    std::atomic<size_t> next(0);
    auto assemble = [&]() {
      for (;;) {
        const size_t c = next.fetch_add(1);
        if (c >= used_chunks) return;
        ChunkOut &ch = chunks[c];
        mtb_triangle *dst = scene->triangles.data() + tri0 + (size_t)tri_base[c];
        for (size_t f = 0; f < ch.faces.size(); f++) {
          const FaceRec &fr = ch.faces[f];
          // what a sequential reader had seen when it reached this face
          const size_t have_pos = (size_t)(pos_base[c] / 3 + fr.n_pos), have_nrm = (size_t)(nrm_base[c] / 3 + fr.n_nrm),
                       have_tex = (size_t)(tex_base[c] / 3 + fr.n_tex);
          for (int base = 0; base + 3 <= fr.count; base += 2) {
            mtb_triangle tr;
            memset(&tr, 0, sizeof(tr));
            const char *bad = nullptr;
            for (int j = 0; j < 3; j++) {
              const int idx = fr.vi[base + j];
              // The reference indexes its vectors unchecked (objreader.cc:155); out-of-range is refused here.
              if (idx < 0 || (size_t)idx >= have_pos) {
                bad = "vertex";
                break;
              }
              memcpy(tr.vertex + j * 3, &pos[(size_t)idx * 3], 3 * sizeof(double));
            }
            if (bad == nullptr && fr.ni[base] != -1 && fr.ni[base + 1] != -1 && fr.ni[base + 2] != -1) {
              for (int j = 0; j < 3; j++) {
                const int idx = fr.ni[base + j];
                if (idx < 0 || (size_t)idx >= have_nrm) {
                  bad = "normal";
                  break;
                }
                memcpy(tr.normal + j * 3, &nrm[(size_t)idx * 3], 3 * sizeof(double));
              }
            }
            if (bad == nullptr && fr.ti[base] != -1 && fr.ti[base + 1] != -1 && fr.ti[base + 2] != -1) {
              for (int j = 0; j < 3; j++) {
                const int idx = fr.ti[base + j];
                if (idx < 0 || (size_t)idx >= have_tex) {
                  bad = "texcoord";
                  break;
                }
                memcpy(tr.uvw + j * 3, &tex[(size_t)idx * 3], 3 * sizeof(double));
              }
            }
            if (bad != nullptr) {
              if (ch.bad_face < 0) {
                ch.bad_face = (int32_t)f;
                ch.bad_text = std::string(bad) + " index out of range in line " + std::to_string(line_base[c] + fr.piece);
              }
              break;
            }
            tr.material = fr.material;
            tr.line_no = (int32_t)(line_base[c] + fr.piece);
            *dst++ = tr;
          }
          if (ch.bad_face >= 0) break;
        }
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; t++) pool.emplace_back(assemble);
    assemble();
    for (std::thread &t : pool) t.join();
  }
  for (size_t c = 0; c < used_chunks; c++) {
    if (chunks[c].bad_face >= 0) {  // the earliest error in file order
      *err = chunks[c].bad_text;
      scene->triangles.resize(tri0);
      return false;
    }
  }
  if (failed) {
    *err = fail_text;
    return false;
  }
  return true;
}

}  // namespace mtb
