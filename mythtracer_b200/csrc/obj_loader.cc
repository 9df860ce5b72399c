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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>

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

bool LoadObjFile(const char *path, LoadedScene *scene, std::string *err) {
  FileCloser fc{fopen(path, "r")};
  if (fc.f == nullptr) {
    *err = std::string("file \"") + path + "\" not found";
    return false;
  }
  const std::string dir = DirOf(path);
  std::vector<double> pos, nrm, tex;  // xyz triples
  int material = -1;
  char line[128];
  for (int line_no = 0; fgets(line, sizeof(line), fc.f) != nullptr; line_no++) {
    ChopLineEnd(line);
    // sscanf(line, "%15s", key): skip white space, then up to 15 non-space characters (done by hand: the format
    // interpreter was a quarter of the load time of a 500 k-triangle model)
    char key_buf[16] = {0};
    {
      const char *q = line;
      while (IsSpace(*q)) q++;
      int n = 0;
      while (*q != '\0' && !IsSpace(*q) && n < 15) key_buf[n++] = *q++;
      if (n == 0 || key_buf[0] == '#') continue;
    }
    const std::string key(key_buf);
    if (key == "v" || key == "vn") {
      double x, y, z;
      // sscanf(line, "v %lf %lf %lf"): the literal must start the line (no leading blanks), then three doubles
      const char *p = line + key.size();
      if (strncmp(line, key.c_str(), key.size()) != 0 || !ScanDouble(&p, &x) || !ScanDouble(&p, &y) || !ScanDouble(&p, &z)) {
        *err = std::string("unsupported ") + (key == "v" ? "vertex" : "normal") + " format \"" + line + "\"";
        return false;
      }
      std::vector<double> &dst = key == "v" ? pos : nrm;
      dst.push_back(x);
      dst.push_back(y);
      dst.push_back(z);
    } else if (key == "vt") {
      double u, v, w = 0.0;
      const char *p = line + 2;
      if (strncmp(line, "vt", 2) != 0 || !ScanDouble(&p, &u) || !ScanDouble(&p, &v)) {
        *err = std::string("unsupported texcoord format \"") + line + "\"";
        return false;
      }
      ScanDouble(&p, &w);  // optional third coordinate
      tex.push_back(u);
      tex.push_back(v);
      tex.push_back(w);
    } else if (key == "mtllib") {
      char fname[256];
      if (sscanf(line, "mtllib %255[^\n]", fname) != 1) {
        *err = std::string("unsupported mtllib format \"") + line + "\"";
        return false;
      }
      MtlParser mp(scene, err);
      if (!mp.Parse(Join(dir, fname))) return false;
    } else if (key == "usemtl") {
      char name[128];
      if (sscanf(line, "usemtl %127s", name) != 1) {
        *err = "unsupported usemtl format";
        return false;
      }
      material = -1;
      for (size_t i = 0; i < scene->material_names.size(); i++) {
        if (scene->material_names[i] == name) material = (int)i;
      }
      if (material < 0) fprintf(stderr, "warning: material \"%s\" not found\n", name);
    } else if (key == "f") {
      // `s >> token; if (s.eof()) break;` keeps a token only when at least one more character follows it
      // (objreader.cc:109-115): the first token ("f") is skipped, a token that ends the line is dropped.
      int vi[5], ti[5], ni[5];
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
        if (!ScanFaceToken(tokbuf, &v, &vt, &vn)) {
          *err = std::string("unsupported face format \"") + tokbuf + "\"";
          return false;
        }
        if (count < 4) {
          vi[count] = v - 1;
          ti[count] = vt - 1;
          ni[count] = vn - 1;
        }
        count++;
      }
      if (count != 3 && count != 4) {
        *err = "unsupported face count (" + std::to_string(count) + ")\n  " + line;
        return false;
      }
      if (count == 4) {
        vi[4] = vi[0];
        ti[4] = ti[0];
        ni[4] = ni[0];
        count = 5;
      }
      for (int base = 0; base + 3 <= count; base += 2) {
        mtb_triangle tr;
        memset(&tr, 0, sizeof(tr));
        for (int j = 0; j < 3; j++) {
          const int idx = vi[base + j];
          // The reference indexes its vectors unchecked (objreader.cc:155); out-of-range is refused here.
          if (idx < 0 || (size_t)idx * 3 + 2 >= pos.size()) {
            *err = std::string("vertex index out of range in \"") + line + "\"";
            return false;
          }
          memcpy(tr.vertex + j * 3, &pos[(size_t)idx * 3], 3 * sizeof(double));
        }
        if (ni[base] != -1 && ni[base + 1] != -1 && ni[base + 2] != -1) {
          for (int j = 0; j < 3; j++) {
            const int idx = ni[base + j];
            if (idx < 0 || (size_t)idx * 3 + 2 >= nrm.size()) {
              *err = std::string("normal index out of range in \"") + line + "\"";
              return false;
            }
            memcpy(tr.normal + j * 3, &nrm[(size_t)idx * 3], 3 * sizeof(double));
          }
        }
        if (ti[base] != -1 && ti[base + 1] != -1 && ti[base + 2] != -1) {
          for (int j = 0; j < 3; j++) {
            const int idx = ti[base + j];
            if (idx < 0 || (size_t)idx * 3 + 2 >= tex.size()) {
              *err = std::string("texcoord index out of range in \"") + line + "\"";
              return false;
            }
            memcpy(tr.uvw + j * 3, &tex[(size_t)idx * 3], 3 * sizeof(double));
          }
        }
        tr.material = material;
        tr.line_no = line_no;
        scene->triangles.push_back(tr);
      }
    } else if (key == "s" || key == "g" || key == "o") {
      continue;
    } else {
      fprintf(stderr, "warning: unknown OBJ feature \"%s\"\n", key_buf);
    }
  }
  return true;
}

}  // namespace mtb
