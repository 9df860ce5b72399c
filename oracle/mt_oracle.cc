// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of MythTracer's per-pixel ray-casting path.
//
// This file is the parity CHECKER for the CUDA path, never the thing shipped or measured as the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs load it.
// It restates, in plain FP64 scalar C++ (no FMA contraction: build with -ffp-contract=off and no -march),
// what the reference computes; every function cites the reference file:line it follows.  Its own parity
// is pinned against the unmodified reference (oracle/_ref) and against fixtures the reference produced
// (see mt_oracle.h).  On top of what the reference exposes it taps per-pixel decision signatures, ray
// counts and work counters, which the reference cannot report without being patched.
#include "mt_oracle.h"

#include <omp.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------
// math3d.h:31-136 -- operation order matters for bit parity, so every helper spells it out.
// ---------------------------------------------------------------------------------------------------
struct V3 {
  double v[3];
};

inline V3 mk(double x, double y, double z) { return V3{{x, y, z}}; }
inline V3 add(const V3 &a, const V3 &b) { return mk(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
inline V3 sub(const V3 &a, const V3 &b) { return mk(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
inline V3 neg(const V3 &a) { return mk(-a.v[0], -a.v[1], -a.v[2]); }
inline V3 mulv(const V3 &a, const V3 &b) { return mk(a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2]); }
inline V3 muls(const V3 &a, double s) { return mk(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
inline V3 divs(const V3 &a, double s) { return mk(a.v[0] / s, a.v[1] / s, a.v[2] / s); }
// math3d.h:116-118  (this = b, argument = a: a0*b0 + a1*b1 + a2*b2; products commute, sums do not)
inline double dot(const V3 &a, const V3 &b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
// math3d.h:120-126  this.Cross(a)
inline V3 cross(const V3 &t, const V3 &a) {
  return mk(t.v[1] * a.v[2] - t.v[2] * a.v[1], t.v[2] * a.v[0] - t.v[0] * a.v[2], t.v[0] * a.v[1] - t.v[1] * a.v[0]);
}
inline double sqrlen(const V3 &a) { return a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]; }
// math3d.h:105-114
inline double sqrdist(const V3 &t, const V3 &a) {
  const double dx = a.v[0] - t.v[0], dy = a.v[1] - t.v[1], dz = a.v[2] - t.v[2];
  return dx * dx + dy * dy + dz * dz;
}
inline double dist(const V3 &t, const V3 &a) { return std::sqrt(sqrdist(t, a)); }
// math3d.h:128-131
inline V3 normalized(const V3 &a) {
  const double l = std::sqrt(sqrlen(a));
  return mk(a.v[0] / l, a.v[1] / l, a.v[2] / l);
}

// std::min / std::max as libstdc++ defines them; the NaN behaviour is part of the contract
// (SURVEY.md fact 9, appendix A.7).
inline double smin(double a, double b) { return (b < a) ? b : a; }
inline double smax(double a, double b) { return (a < b) ? b : a; }
inline double smin3(double a, double b, double c) {
  double r = a;
  if (b < r) r = b;
  if (c < r) r = c;
  return r;
}
inline double smax3(double a, double b, double c) {
  double r = a;
  if (r < b) r = b;
  if (r < c) r = c;
  return r;
}

struct Box {
  V3 lo, hi;
};

// math3d.h:184-293: 4x4 row-major matrices, only what camera.cc uses.
struct M4 {
  double m[4][4];
};
inline M4 m4mul(const M4 &a, const M4 &b) {  // math3d.h:188-200
  M4 r;
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++)
      r.m[j][i] = a.m[j][0] * b.m[0][i] + a.m[j][1] * b.m[1][i] + a.m[j][2] * b.m[2][i] + a.m[j][3] * b.m[3][i];
  return r;
}
inline V3 m4v(const M4 &a, const V3 &p) {  // math3d.h:210-216 (all three rows add m[0][3])
  return mk(a.m[0][0] * p.v[0] + a.m[0][1] * p.v[1] + a.m[0][2] * p.v[2] + a.m[0][3],
            a.m[1][0] * p.v[0] + a.m[1][1] * p.v[1] + a.m[1][2] * p.v[2] + a.m[0][3],
            a.m[2][0] * p.v[0] + a.m[2][1] * p.v[1] + a.m[2][2] * p.v[2] + a.m[0][3]);
}
inline double deg2rad(double a) { return (a * M_PI) / 180.0; }  // math3d.h:177-179
inline M4 rot_x(double deg) {  // math3d.h:226-233
  const double a = deg2rad(deg);
  return M4{{{1.0, 0.0, 0.0, 0.0}, {0.0, cos(a), -sin(a), 0.0}, {0.0, sin(a), cos(a), 0.0}, {0.0, 0.0, 0.0, 1.0}}};
}
inline M4 rot_y(double deg) {  // math3d.h:235-242
  const double a = deg2rad(deg);
  return M4{{{cos(a), 0.0, sin(a), 0.0}, {0.0, 1.0, 0.0, 0.0}, {-sin(a), 0.0, cos(a), 0.0}, {0.0, 0.0, 0.0, 1.0}}};
}
inline M4 rot_z(double deg) {  // math3d.h:244-251
  const double a = deg2rad(deg);
  return M4{{{cos(a), -sin(a), 0.0, 0.0}, {sin(a), cos(a), 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}, {0.0, 0.0, 0.0, 1.0}}};
}

struct Sensor {
  V3 start, d_scan, d_pixel, origin;
};

// camera.cc:27-63
Sensor make_sensor(const mto_camera &cam, int width, int height) {
  const double aov_vertical = (double(height) / double(width)) * cam.aov;
  const M4 rot_left = rot_y(cam.aov / 2.0);
  const M4 rot_right = rot_y(-cam.aov / 2.0);
  const M4 rot_top = rot_z(aov_vertical / 2.0);
  const M4 rot_bottom = rot_z(-aov_vertical / 2.0);
  const M4 left_top = m4mul(rot_top, rot_left);
  const M4 right_top = m4mul(rot_bottom, rot_right);  // sic: camera.cc:38
  const M4 left_bottom = m4mul(rot_bottom, rot_left);
  const V3 fwd = mk(0.0, 0.0, 1.0);
  V3 tl = m4v(left_top, fwd);
  V3 tr = m4v(right_top, fwd);
  V3 bl = m4v(left_bottom, fwd);
  const M4 frustum = m4mul(m4mul(rot_y(cam.yaw), rot_x(cam.pitch)), rot_z(cam.roll));
  tl = m4v(frustum, tl);
  tr = m4v(frustum, tr);
  bl = m4v(frustum, bl);
  Sensor s;
  s.d_scan = divs(sub(bl, tl), double(height));
  s.d_pixel = divs(sub(tr, tl), double(width));
  s.start = tl;
  s.origin = mk(cam.origin[0], cam.origin[1], cam.origin[2]);
  return s;
}

// camera.cc:65-69
inline V3 sensor_dir(const Sensor &s, int x, int y) {
  return normalized(add(add(s.start, muls(s.d_scan, double(y))), muls(s.d_pixel, double(x))));
}

// ---------------------------------------------------------------------------------------------------
// scene data
// ---------------------------------------------------------------------------------------------------
struct Tri {
  V3 vertex[3], normal[3], uvw[3];
  Box box;  // Triangle::CacheAABB, primitive_triangle.cc:18-24
  int32_t material, line_no;
};

struct Tex {
  size_t width = 0, height = 0;
  std::vector<V3> colors;
};

struct Node {
  std::vector<int32_t> prims;  // insertion order preserved (octtree.cc:106-129)
  int32_t child = -1;          // index of child 0; the 8 children are contiguous (octtree.cc:58)
  Box box;
};

struct Ray {
  V3 o, d, inv;
};

struct Counters {
  mto_stats s;
};

}  // namespace

struct mto_scene {
  std::vector<Tri> tris;
  std::vector<mto_material> mtls;
  std::vector<Tex> texs;
  std::vector<mto_light> lights;
  std::vector<Node> nodes;  // nodes[0] = root
  int64_t depth = 0;
};

namespace {

int g_threads = 0;

// aabb.cc:29-33 (closed intervals)
inline bool box_has_point(const Box &b, const V3 &p) {
  return p.v[0] >= b.lo.v[0] && p.v[0] <= b.hi.v[0] && p.v[1] >= b.lo.v[1] && p.v[1] <= b.hi.v[1] &&
         p.v[2] >= b.lo.v[2] && p.v[2] <= b.hi.v[2];
}
// aabb.cc:5-7
inline bool box_fully_contains(const Box &outer, const Box &inner) {
  return box_has_point(outer, inner.lo) && box_has_point(outer, inner.hi);
}

// octtree.cc:46-135.  Children are appended to s->nodes, so `idx` must be re-read after the resize.
void attempt_split(mto_scene *s, int32_t idx, int64_t level) {
  if (level > s->depth) s->depth = level;
  if (s->nodes[idx].prims.size() < 16) return;  // SPLIT_BOUNDARY, octtree.h:43
  const Box b = s->nodes[idx].box;
  V3 c;  // CalcCenter, octtree.cc:46-50
  for (int i = 0; i < 3; i++) c.v[i] = b.lo.v[i] + (b.hi.v[i] - b.lo.v[i]) / 2.0;
  const int32_t first = (int32_t)s->nodes.size();
  s->nodes.resize(s->nodes.size() + 8);
  s->nodes[idx].child = first;
  // octtree.cc:61-100: bit0 = +x, bit1 = +z, bit2 = +y
  for (int k = 0; k < 8; k++) {
    Box cb;
    cb.lo.v[0] = (k & 1) ? c.v[0] : b.lo.v[0];
    cb.hi.v[0] = (k & 1) ? b.hi.v[0] : c.v[0];
    cb.lo.v[2] = (k & 2) ? c.v[2] : b.lo.v[2];
    cb.hi.v[2] = (k & 2) ? b.hi.v[2] : c.v[2];
    cb.lo.v[1] = (k & 4) ? c.v[1] : b.lo.v[1];
    cb.hi.v[1] = (k & 4) ? b.hi.v[1] : c.v[1];
    s->nodes[first + k].box = cb;
  }
  std::vector<int32_t> remaining;
  for (int32_t p : s->nodes[idx].prims) {  // octtree.cc:106-122: first child that fully contains wins
    bool placed = false;
    for (int k = 0; k < 8 && !placed; k++) {
      if (box_fully_contains(s->nodes[first + k].box, s->tris[p].box)) {
        s->nodes[first + k].prims.push_back(p);
        placed = true;
      }
    }
    if (!placed) remaining.push_back(p);
  }
  s->nodes[idx].prims.swap(remaining);
  s->nodes[idx].prims.shrink_to_fit();
  for (int k = 0; k < 8; k++) attempt_split(s, first + k, level + 1);  // octtree.cc:132-134
}

// The slab test shared by Node::NodeIntersectRay (octtree.cc:138-167) and the triangle pre-test
// (primitive_triangle.cc:85-108).
inline bool slab(const Box &b, const Ray &r, double *tmin_out) {
  const double t1 = (b.lo.v[0] - r.o.v[0]) * r.inv.v[0];
  const double t2 = (b.hi.v[0] - r.o.v[0]) * r.inv.v[0];
  const double t3 = (b.lo.v[1] - r.o.v[1]) * r.inv.v[1];
  const double t4 = (b.hi.v[1] - r.o.v[1]) * r.inv.v[1];
  const double t5 = (b.lo.v[2] - r.o.v[2]) * r.inv.v[2];
  const double t6 = (b.hi.v[2] - r.o.v[2]) * r.inv.v[2];
  const double tmax = smin3(smax(t1, t2), smax(t3, t4), smax(t5, t6));
  if (tmax < 0.0) return false;
  const double tmin = smax3(smin(t1, t2), smin(t3, t4), smin(t5, t6));
  if (tmin > tmax) return false;
  *tmin_out = tmin;
  return true;
}

// primitive_triangle.cc:81-143
inline bool tri_intersect(const Tri &tr, const Ray &r, double *t_out, Counters *cn) {
  cn->s.n_triaabb++;
  double unused;
  if (!slab(tr.box, r, &unused)) return false;
  cn->s.n_mt++;
  const V3 e1 = sub(tr.vertex[1], tr.vertex[0]);
  const V3 e2 = sub(tr.vertex[2], tr.vertex[0]);
  const V3 pvec = cross(r.d, e2);
  const double det = dot(e1, pvec);
  if (det >= -0.00000001 && det < 0.00000001) return false;
  const double inv_det = 1.0 / det;
  const V3 tvec = sub(r.o, tr.vertex[0]);
  const double u = dot(tvec, pvec) * inv_det;
  if (u < 0.0 || u > 1.0) return false;
  const V3 qvec = cross(tvec, e1);
  const double v = dot(r.d, qvec) * inv_det;
  if (v < 0.0 || u + v > 1.0) return false;
  const double t = dot(e2, qvec) * inv_det;
  if (t < 0.0) return false;
  cn->s.n_hit++;
  *t_out = t;
  return true;
}

// std::sort on <= 16 elements is libstdc++'s insertion sort (bits/stl_algo.h __insertion_sort): an
// element smaller than the first goes to the front, otherwise it is moved left while it compares less
// than its predecessor.  Restated literally so that NaN keys land where the reference puts them.
void sort_children(int32_t *idx, double *key, int n) {
  for (int i = 1; i < n; i++) {
    const int32_t vi = idx[i];
    const double vk = key[i];
    if (vk < key[0]) {
      for (int j = i; j > 0; j--) {
        idx[j] = idx[j - 1];
        key[j] = key[j - 1];
      }
      idx[0] = vi;
      key[0] = vk;
    } else {
      int j = i;
      while (vk < key[j - 1]) {
        idx[j] = idx[j - 1];
        key[j] = key[j - 1];
        j--;
      }
      idx[j] = vi;
      key[j] = vk;
    }
  }
}

// octtree.cc:169-257.  Returns the triangle index or -1.
int32_t node_query(const mto_scene *s, int32_t idx, const Ray &r, double *t_out, Counters *cn) {
  cn->s.n_visit++;
  const Node &n = s->nodes[idx];
  int32_t best = -1;
  double best_t = 0.0;
  for (int32_t p : n.prims) {  // octtree.cc:177-196
    double t;
    if (!tri_intersect(s->tris[p], r, &t, cn)) continue;
    if (best != -1 && t > best_t) continue;
    best = p;
    best_t = t;
  }
  if (n.child >= 0) {
    int32_t order[8];
    double key[8];
    int cnt = 0;
    for (int k = 0; k < 8; k++) {  // octtree.cc:204-211
      double d;
      cn->s.n_slab++;
      if (!slab(s->nodes[n.child + k].box, r, &d)) continue;
      order[cnt] = n.child + k;
      key[cnt] = d;
      cnt++;
    }
    sort_children(order, key, cnt);  // octtree.cc:213-216
    for (int i = 0; i < cnt; i++) {  // octtree.cc:219-247
      double t;
      const int32_t p = node_query(s, order[i], r, &t, cn);
      if (p == -1) continue;
      if (best != -1 && t > best_t) continue;
      best = p;
      best_t = t;
      break;
    }
  }
  if (best == -1) return -1;
  *t_out = best_t;
  return best;
}

// octtree.cc:26-40
int32_t tree_intersect(const mto_scene *s, const V3 &o, const V3 &d, V3 *point, double *t_out, Counters *cn) {
  cn->s.rays++;
  Ray r;
  r.o = o;
  r.d = d;
  r.inv = mk(1.0 / d.v[0], 1.0 / d.v[1], 1.0 / d.v[2]);
  double unused;
  cn->s.n_slab++;
  if (!slab(s->nodes[0].box, r, &unused)) return -1;
  double t;
  const int32_t p = node_query(s, 0, r, &t, cn);
  if (p == -1) return -1;
  *t_out = t;
  *point = add(o, muls(d, t));  // primitive_triangle.cc:141 (same expression, same inputs)
  return p;
}

// primitive_triangle.cc:27-40
inline double heron(double a, double b, double c) {
  const double p = (a + b + c) / 2.0;
  const double area_sqr = p * (p - a) * (p - b) * (p - c);
  if (area_sqr < 0.0) return 0.0;
  return std::sqrt(area_sqr);
}

// primitive_triangle.cc:43-79 (GetNormal and GetUVW share the weights)
inline V3 bary_interp(const Tri &tr, const V3 &point, const V3 attr[3]) {
  const double a = dist(tr.vertex[0], tr.vertex[1]);
  const double b = dist(tr.vertex[1], tr.vertex[2]);
  const double c = dist(tr.vertex[2], tr.vertex[0]);
  const double p0 = dist(point, tr.vertex[0]);
  const double p1 = dist(point, tr.vertex[1]);
  const double p2 = dist(point, tr.vertex[2]);
  const double n0 = heron(b, p2, p1);
  const double n1 = heron(c, p0, p2);
  const double n2 = heron(a, p1, p0);
  const double n = n0 + n1 + n2;
  return divs(add(add(muls(attr[0], n0), muls(attr[1], n1)), muls(attr[2], n2)), n);
}

// texture.cc:11-58
V3 tex_sample(const Tex &tx, double u, double v) {
  u = fmod(u, 1.0);
  v = fmod(v, 1.0);
  if (u < 0.0) u += 1.0;
  if (v < 0.0) v += 1.0;
  v = 1.0 - v;
  const double x = u * (double)(tx.width - 1);
  const double y = v * (double)(tx.height - 1);
  const size_t bx = (size_t)x;
  const size_t by = (size_t)y;
  const size_t bx1 = (bx + 1 == tx.width) ? bx : bx + 1;
  const size_t by1 = (by + 1 == tx.height) ? by : by + 1;
  const V3 c0 = tx.colors.at(bx + by * tx.width);
  const V3 c1 = tx.colors.at(bx1 + by * tx.width);
  const V3 c2 = tx.colors.at(bx + by1 * tx.width);
  const V3 c3 = tx.colors.at(bx1 + by1 * tx.width);
  const double dx = fmod(x, 1.0);
  const double dy = fmod(y, 1.0);
  const double a0 = (1.0 - dx) * (1.0 - dy);
  const double a1 = dx * (1.0 - dy);
  const double a2 = (1.0 - dx) * dy;
  const double a3 = dx * dy;
  return add(add(add(muls(c0, a0), muls(c1, a1)), muls(c2, a2)), muls(c3, a3));
}

inline uint64_t mix64(uint64_t path, uint64_t kind, uint64_t value) {
  uint64_t z = path * 0x9E3779B97F4A7C15ull + kind * 0xC2B2AE3D27D4EB4Full + value * 0x165667B19E3779F9ull +
               0x27D4EB2F165667C5ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct PixelTap {
  uint64_t sig_hits = 0, sig_shadow = 0;
  uint32_t n_rays = 0;
  int32_t dbg_line = -1;
  V3 dbg_point;
  bool want_dbg = false;
};

// mythtracer.cc:13-228
V3 trace(const mto_scene *s, const V3 &ro, const V3 &rd, int level, bool in_object, double coef, int max_depth,
         uint64_t path, PixelTap *tap, Counters *cn) {
  V3 P;
  double hit_t;
  tap->n_rays++;
  const int32_t prim = tree_intersect(s, ro, rd, &P, &hit_t, cn);
  if (prim == -1) {
    if (level == 0 && tap->want_dbg) {  // mythtracer.cc:24-27
      tap->dbg_line = -1;
      tap->dbg_point = mk(NAN, NAN, NAN);
    }
    return mk(0.0, 0.0, 0.0);
  }
  const Tri &tr = s->tris[prim];
  if (level == 0 && tap->want_dbg) {  // mythtracer.cc:33-36
    tap->dbg_line = tr.line_no;
    tap->dbg_point = P;
  }
  tap->sig_hits += mix64(path, 1, (uint64_t)(int64_t)tr.line_no);
  cn->s.n_shade++;

  V3 normal = bary_interp(tr, P, tr.normal);  // mythtracer.cc:38
  const V3 towards_camera = neg(rd);
  double normal_ray_dot = dot(towards_camera, normal);  // normal.Dot(towards_camera)
  if (normal_ray_dot < 0.0) {
    normal = neg(normal);
    normal_ray_dot = dot(towards_camera, normal);
  }
  if (tr.material < 0) {  // mythtracer.cc:49-52
    normal_ray_dot = (normal_ray_dot + 1.0) * 0.5;
    return mk(normal_ray_dot, normal_ray_dot, normal_ray_dot);
  }
  const mto_material &m = s->mtls[tr.material];
  V3 surface = mk(m.ambient[0], m.ambient[1], m.ambient[2]);
  if (m.texture >= 0) {  // mythtracer.cc:59-64
    const V3 uvw = bary_interp(tr, P, tr.uvw);
    const V3 tc = tex_sample(s->texs[m.texture], uvw.v[0], uvw.v[1]);
    surface = mulv(surface, tc);
  }
  // mythtracer.cc:68-69: ray.direction - normal * (2 * ray.direction.Dot(normal))
  const V3 reflected = sub(rd, muls(normal, 2 * dot(normal, rd)));
  const V3 refl_origin = add(P, muls(reflected, 0.0001));  // mythtracer.cc:72

  const V3 Kd = mk(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
  const V3 Ks = mk(m.specular[0], m.specular[1], m.specular[2]);
  V3 color = mk(0.0, 0.0, 0.0);
  for (size_t li = 0; li < s->lights.size(); li++) {  // mythtracer.cc:78
    const mto_light &lt = s->lights[li];
    const V3 lpos = mk(lt.position[0], lt.position[1], lt.position[2]);
    const V3 lamb = mk(lt.ambient[0], lt.ambient[1], lt.ambient[2]);
    const V3 ldir = normalized(sub(lpos, P));
    color = add(color, mulv(lamb, surface));  // mythtracer.cc:83-84

    V3 power = mk(1.0, 1.0, 1.0);
    bool in_shadow = false;
    bool through = false;
    uint64_t segments = 0;
    for (V3 start = P;;) {  // mythtracer.cc:94-156
      const V3 so = add(start, muls(ldir, 0.00001));
      const double light_distance = dist(start, lpos);
      V3 sp;
      double sd;
      tap->n_rays++;
      cn->s.shadow++;
      segments++;
      const int32_t sprim = tree_intersect(s, so, ldir, &sp, &sd, cn);
      if (sprim == -1) break;
      if (sd > light_distance) break;
      const int32_t smtl = s->tris[sprim].material;
      // The reference dereferences mtl unconditionally (mythtracer.cc:121); scenes handed to the
      // oracle give every triangle a material.  A missing material is treated as opaque here.
      const double str = smtl >= 0 ? s->mtls[smtl].transparency : 0.0;
      if (str == 0.0) {
        power = mk(0.0, 0.0, 0.0);
        in_shadow = true;
        break;
      }
      if (!through) {  // mythtracer.cc:129-132: light_power *= Tf * Tr
        const mto_material &sm = s->mtls[smtl];
        const V3 tf = mk(sm.transmission_filter[0], sm.transmission_filter[1], sm.transmission_filter[2]);
        power = mulv(power, muls(tf, sm.transparency));
      }
      through = !through;
      start = add(sp, muls(ldir, 0.0000001));  // mythtracer.cc:137
      if (sqrdist(P, start) > sqrdist(P, lpos)) break;  // mythtracer.cc:141-145
      if (power.v[0] <= 0.001 && power.v[1] <= 0.001 && power.v[2] <= 0.001) {
        power = mk(0.0, 0.0, 0.0);
        in_shadow = true;
        break;
      }
    }
    tap->sig_shadow += mix64(path, 2 + li, (in_shadow ? 1u : 0u) | (segments << 1));
    // mythtracer.cc:159-161
    power.v[0] = smax(power.v[0], lamb.v[0]);
    power.v[1] = smax(power.v[1], lamb.v[1]);
    power.v[2] = smax(power.v[2], lamb.v[2]);
    // mythtracer.cc:163-167: ((((Kd * surface) * L.N) * light.diffuse) * power)
    const V3 ldif = mk(lt.diffuse[0], lt.diffuse[1], lt.diffuse[2]);
    color = add(color, mulv(mulv(muls(mulv(Kd, surface), dot(normal, ldir)), ldif), power));
    if (!in_shadow) {  // mythtracer.cc:169-177
      const double refl_dot = dot(towards_camera, reflected);
      if (refl_dot > 0) {
        const V3 lspec = mk(lt.specular[0], lt.specular[1], lt.specular[2]);
        color = add(color, mulv(muls(mulv(Ks, surface), pow(refl_dot, m.specular_exp)), lspec));
      }
    }
  }

  if (level < max_depth && m.reflectance > 0.0 && coef > 0.01 && !in_object) {  // mythtracer.cc:181-189
    cn->s.reflect++;
    const V3 c = trace(s, refl_origin, reflected, level + 1, in_object, coef * m.reflectance, max_depth, path * 2,
                       tap, cn);
    color = add(color, muls(c, m.reflectance));
  }
  if (level < max_depth && m.transparency > 0.0) {  // mythtracer.cc:192-225
    const V3 refr = normalized(rd);                  // the bending formula is commented out upstream
    const V3 refr_origin = add(P, muls(refr, 0.00001));
    cn->s.refract++;
    const V3 c = trace(s, refr_origin, refr, level + 1, !in_object, coef, max_depth, path * 2 + 1, tap, cn);
    const V3 tf = mk(m.transmission_filter[0], m.transmission_filter[1], m.transmission_filter[2]);
    color = add(color, muls(mulv(c, tf), m.transparency));  // (c * Tf) * Tr
  }
  return color;
}

// mythtracer.cc:235-241.  (uint8_t)(double) is undefined for NaN; x86-64 yields 0, which is restated.
inline void quantize(const V3 &c, uint8_t *rgb) {
  for (int i = 0; i < 3; i++) {
    const double x = c.v[i];
    if (x > 1.0) rgb[i] = 255;
    else if (x < 0.0) rgb[i] = 0;
    else if (x != x) rgb[i] = 0;
    else rgb[i] = (uint8_t)(x * 255);
  }
}

void add_stats(mto_stats *dst, const mto_stats &src) {
  uint64_t *d = reinterpret_cast<uint64_t *>(dst);
  const uint64_t *q = reinterpret_cast<const uint64_t *>(&src);
  for (size_t i = 0; i < sizeof(mto_stats) / sizeof(uint64_t); i++) d[i] += q[i];
}

int render_impl(const mto_scene *s, const mto_camera *cam, int image_w, int image_h, int chunk_x, int chunk_y,
                int chunk_w, int chunk_h, int max_depth, uint8_t *rgb, double *color_out, int32_t *dbg_line_no,
                double *dbg_point, const mto_taps *taps, mto_stats *stats) {
  if (s == nullptr || cam == nullptr || chunk_w <= 0 || chunk_h <= 0) return -1;
  const Sensor sensor = make_sensor(*cam, image_w, image_h);  // mythtracer.cc:289-290: full image size
  mto_stats total;
  memset(&total, 0, sizeof(total));
  const int nthreads = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
  {
    Counters cn;
    memset(&cn, 0, sizeof(cn));
#pragma omp for schedule(dynamic, 1)
    for (int j = 0; j < chunk_h; j++) {  // mythtracer.cc:295-302
      for (int i = 0; i < chunk_w; i++) {
        const size_t px = (size_t)j * chunk_w + i;
        PixelTap tap;
        tap.want_dbg = (dbg_line_no != nullptr || dbg_point != nullptr);
        cn.s.primary++;
        const V3 d = sensor_dir(sensor, chunk_x + i, chunk_y + j);
        const V3 c = trace(s, sensor.origin, d, 0, false, 1.0, max_depth, 1, &tap, &cn);
        if (rgb != nullptr) quantize(c, rgb + px * 3);
        if (color_out != nullptr) memcpy(color_out + px * 3, c.v, sizeof(c.v));
        if (dbg_line_no != nullptr) dbg_line_no[px] = tap.dbg_line;
        if (dbg_point != nullptr) memcpy(dbg_point + px * 3, tap.dbg_point.v, sizeof(tap.dbg_point.v));
        if (taps != nullptr) {
          if (taps->sig_hits) taps->sig_hits[px] = tap.sig_hits;
          if (taps->sig_shadow) taps->sig_shadow[px] = tap.sig_shadow;
          if (taps->n_rays) taps->n_rays[px] = tap.n_rays;
        }
      }
    }
#pragma omp critical
    add_stats(&total, cn.s);
  }
  if (stats != nullptr) *stats = total;
  return 0;
}

}  // namespace

extern "C" {

mto_scene *mto_create(const mto_triangle *tris, int64_t n_tris, const mto_material *mtls, int32_t n_mtls,
                      const mto_texture *texs, int32_t n_texs) {
  mto_scene *s = new mto_scene;
  s->tris.resize((size_t)n_tris);
  s->nodes.resize(1);
  // OctTree root box starts as {0,0,0}-{0,0,0} (math3d.h:141) and is only extended (octtree.cc:12-13).
  s->nodes[0].box.lo = mk(0.0, 0.0, 0.0);
  s->nodes[0].box.hi = mk(0.0, 0.0, 0.0);
  for (int64_t i = 0; i < n_tris; i++) {
    Tri &t = s->tris[(size_t)i];
    for (int k = 0; k < 3; k++) {
      t.vertex[k] = mk(tris[i].vertex[k * 3], tris[i].vertex[k * 3 + 1], tris[i].vertex[k * 3 + 2]);
      t.normal[k] = mk(tris[i].normal[k * 3], tris[i].normal[k * 3 + 1], tris[i].normal[k * 3 + 2]);
      t.uvw[k] = mk(tris[i].uvw[k * 3], tris[i].uvw[k * 3 + 1], tris[i].uvw[k * 3 + 2]);
    }
    t.material = tris[i].material;
    t.line_no = tris[i].line_no;
    // CacheAABB (primitive_triangle.cc:18-24) via AABB::Extend (aabb.cc:42-47): std::min/std::max
    t.box.lo = t.vertex[0];
    t.box.hi = t.vertex[0];
    for (int k = 1; k < 3; k++)
      for (int a = 0; a < 3; a++) {
        t.box.lo.v[a] = smin(t.box.lo.v[a], t.vertex[k].v[a]);
        t.box.hi.v[a] = smax(t.box.hi.v[a], t.vertex[k].v[a]);
      }
    // AddPrimitive (octtree.cc:8-14)
    for (int a = 0; a < 3; a++) {
      s->nodes[0].box.lo.v[a] = smin(s->nodes[0].box.lo.v[a], t.box.lo.v[a]);
      s->nodes[0].box.hi.v[a] = smax(s->nodes[0].box.hi.v[a], t.box.hi.v[a]);
    }
    s->nodes[0].prims.push_back((int32_t)i);
  }
  s->mtls.assign(mtls, mtls + n_mtls);
  s->texs.resize((size_t)n_texs);
  for (int32_t i = 0; i < n_texs; i++) {  // texture.cc:94-106
    Tex &tx = s->texs[(size_t)i];
    tx.width = (size_t)texs[i].width;
    tx.height = (size_t)texs[i].height;
    tx.colors.resize(tx.width * tx.height);
    const uint8_t *px = texs[i].rgba;
    for (size_t k = 0; k < tx.colors.size(); k++, px += 4)
      tx.colors[k] = mk((double)px[0] / 255.0, (double)px[1] / 255.0, (double)px[2] / 255.0);
  }
  attempt_split(s, 0, 0);  // Finalize, octtree.cc:16-24
  return s;
}

void mto_destroy(mto_scene *s) { delete s; }

void mto_set_lights(mto_scene *s, const mto_light *lights, int32_t n) { s->lights.assign(lights, lights + n); }

void mto_set_threads(int32_t n) { g_threads = n; }
int32_t mto_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

void mto_tree_info(const mto_scene *s, int64_t out[5]) {
  int64_t biggest = 0, interior = 0;
  for (const Node &n : s->nodes) {
    if ((int64_t)n.prims.size() > biggest) biggest = (int64_t)n.prims.size();
    if (n.child >= 0) interior += (int64_t)n.prims.size();
  }
  out[0] = (int64_t)s->nodes.size();
  out[1] = s->depth;
  out[2] = biggest;
  out[3] = (int64_t)s->nodes[0].prims.size();
  out[4] = interior;
}

void mto_scene_aabb(const mto_scene *s, double out6[6]) {
  for (int a = 0; a < 3; a++) {
    out6[a] = s->nodes[0].box.lo.v[a];
    out6[3 + a] = s->nodes[0].box.hi.v[a];
  }
}

static void walk_nodes(const mto_scene *s, int32_t idx, int32_t depth, double *node_box, int32_t *node_depth) {
  const Node &n = s->nodes[idx];
  for (int32_t p : n.prims) {
    if (node_depth != nullptr) node_depth[p] = depth;
    if (node_box != nullptr) {
      memcpy(node_box + (size_t)p * 6, n.box.lo.v, 24);
      memcpy(node_box + (size_t)p * 6 + 3, n.box.hi.v, 24);
    }
  }
  if (n.child >= 0)
    for (int k = 0; k < 8; k++) walk_nodes(s, n.child + k, depth + 1, node_box, node_depth);
}

void mto_triangle_nodes(const mto_scene *s, double *node_box, int32_t *node_depth) {
  walk_nodes(s, 0, 0, node_box, node_depth);
}

int mto_render(const mto_scene *s, const mto_camera *cam, int image_w, int image_h, int chunk_x, int chunk_y,
               int chunk_w, int chunk_h, int max_depth, uint8_t *rgb, int32_t *dbg_line_no, double *dbg_point,
               const mto_taps *taps, mto_stats *stats) {
  return render_impl(s, cam, image_w, image_h, chunk_x, chunk_y, chunk_w, chunk_h, max_depth, rgb, nullptr,
                     dbg_line_no, dbg_point, taps, stats);
}

int mto_render_color(const mto_scene *s, const mto_camera *cam, int image_w, int image_h, int chunk_x,
                     int chunk_y, int chunk_w, int chunk_h, int max_depth, double *color) {
  return render_impl(s, cam, image_w, image_h, chunk_x, chunk_y, chunk_w, chunk_h, max_depth, nullptr, color,
                     nullptr, nullptr, nullptr, nullptr);
}

int mto_intersect(const mto_scene *s, int64_t n, const double *origins, const double *dirs, int32_t *tri_index,
                  double *t, double *point, mto_stats *stats) {
  mto_stats total;
  memset(&total, 0, sizeof(total));
  const int nthreads = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
  {
    Counters cn;
    memset(&cn, 0, sizeof(cn));
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; i++) {
      V3 p;
      double d = 0.0;
      const int32_t hit = tree_intersect(s, mk(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]),
                                         mk(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]), &p, &d, &cn);
      tri_index[i] = hit;
      if (hit >= 0) {
        if (t != nullptr) t[i] = d;
        if (point != nullptr) memcpy(point + i * 3, p.v, sizeof(p.v));
      }
    }
#pragma omp critical
    add_stats(&total, cn.s);
  }
  if (stats != nullptr) *stats = total;
  return 0;
}

int mto_intersect_brute(const mto_scene *s, int64_t n, const double *origins, const double *dirs,
                        int32_t *tri_index, double *t) {
  const int nthreads = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
  for (int64_t i = 0; i < n; i++) {
    Counters cn;
    Ray r;
    r.o = mk(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]);
    r.d = mk(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]);
    r.inv = mk(1.0 / r.d.v[0], 1.0 / r.d.v[1], 1.0 / r.d.v[2]);
    int32_t best = -1;
    double best_t = 0.0, unused;
    if (slab(s->nodes[0].box, r, &unused)) {  // the root gate, octtree.cc:34-37
      for (size_t p = 0; p < s->tris.size(); p++) {
        double tt;
        if (!tri_intersect(s->tris[p], r, &tt, &cn)) continue;
        if (best != -1 && tt > best_t) continue;
        best = (int32_t)p;
        best_t = tt;
      }
    }
    tri_index[i] = best;
    if (best >= 0 && t != nullptr) t[i] = best_t;
  }
  return 0;
}

void mto_camera_ray(const mto_camera *cam, int image_w, int image_h, int x, int y, double dir[3]) {
  const Sensor s = make_sensor(*cam, image_w, image_h);
  const V3 d = sensor_dir(s, x, y);
  memcpy(dir, d.v, sizeof(d.v));
}

void mto_camera_sensor(const mto_camera *cam, int image_w, int image_h, double out9[9]) {
  const Sensor s = make_sensor(*cam, image_w, image_h);
  memcpy(out9, s.start.v, 24);
  memcpy(out9 + 3, s.d_scan.v, 24);
  memcpy(out9 + 6, s.d_pixel.v, 24);
}

void mto_texture_sample(const mto_scene *s, int32_t tex, double u, double v, double out[3]) {
  const V3 c = tex_sample(s->texs[(size_t)tex], u, v);
  memcpy(out, c.v, sizeof(c.v));
}

void mto_triangle_normal(const mto_scene *s, int64_t tri, const double point[3], double out[3]) {
  const Tri &t = s->tris[(size_t)tri];
  const V3 r = bary_interp(t, mk(point[0], point[1], point[2]), t.normal);
  memcpy(out, r.v, sizeof(r.v));
}

void mto_triangle_uvw(const mto_scene *s, int64_t tri, const double point[3], double out[3]) {
  const Tri &t = s->tris[(size_t)tri];
  const V3 r = bary_interp(t, mk(point[0], point[1], point[2]), t.uvw);
  memcpy(out, r.v, sizeof(r.v));
}

void mto_quantize(const double color[3], uint8_t rgb[3]) { quantize(mk(color[0], color[1], color[2]), rgb); }

void mto_math3d(const double a[3], const double b[3], double out[9]) {
  const V3 va = mk(a[0], a[1], a[2]), vb = mk(b[0], b[1], b[2]);
  out[0] = std::sqrt(sqrlen(va));  // math3d.h:101-103
  out[1] = dist(va, vb);
  out[2] = dot(vb, va);
  const V3 c = cross(va, vb);
  const V3 n = normalized(va);
  memcpy(out + 3, c.v, 24);
  memcpy(out + 6, n.v, 24);
}

uint64_t mto_mix64(uint64_t path, uint64_t kind, uint64_t value) { return mix64(path, kind, value); }

}  // extern "C"
