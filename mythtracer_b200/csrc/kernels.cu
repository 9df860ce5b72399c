// Hand-written sm_100a kernels for MythTracer's ray-casting path.
//
// Everything that DECIDES a result (which triangle is hit, shadowed or not, the hit point) is computed in
// FP64 with the reference's operation order and without FMA contraction (build with --fmad=false; the
// reference binary contains no FMA), so hit ids, shadow decisions and colours are
// reproduced bit for bit except for pow() (mythtracer.cc:174; CUDA's and glibc's differ by <= 2 ulp).
//
// Reference functions replaced here:
//   OctTree::IntersectRay                      octtree.cc:26-40        -> TraceRegular / TraceLiteral
//   Node::NodeIntersectRay                     octtree.cc:138-167      -> the slab tests inside them
//   Node::PrimitiveIntersectRay                octtree.cc:169-257      -> the explicit frame stack
//   Triangle::IntersectRay                     primitive_triangle.cc:81-143 -> TestSlot*
//   Triangle::GetNormal / GetUVW               primitive_triangle.cc:27-79  -> Barycentric
//   Texture::GetColorAt                        texture.cc:11-58        -> SampleTexture
//   Sensor::GetRay                             camera.cc:65-69         -> PixelDirection
//   MythTracer::TraceRayWorker + shadow walk   mythtracer.cc:13-228    -> the state machine of RenderMega
//   MythTracer::V3DtoRGB                       mythtracer.cc:235-241   -> Quantize
//   the OpenMP row loop                        mythtracer.cc:292-305   -> the CUDA grid (8x8 pixel tiles)
//
// Two traversals exist.  Rays whose direction has a zero / non-finite component ("irregular": the NaN
// producing cases of SURVEY.md fact 9) take TraceLiteral, a literal restatement including std::min/max
// NaN behaviour and libstdc++'s insertion sort.  All other rays take TraceRegular, which may use any
// evaluation order that yields the same VALUES when no NaN can occur: sign-selected near/far planes
// instead of pairwise min/max, the three shared planes of the eight children, skipping empty subtrees,
// and a conservative threaded BVH over long node lists whose candidates are then decided by the exact
// reference tests with the reference's tie rule (later list entry wins on equal t).
#include <math_constants.h>

#include "device_scene.h"

namespace mtb {
namespace {

constexpr int kMaxTreeStack = MTB_MAX_TREE_DEPTH + 2;
constexpr int kMaxRayStack = MTB_MAX_RAY_DEPTH + 2;
constexpr int kTile = 8;            // 8x8 pixel tiles, one per 64-thread block; a warp covers 8x4 pixels
constexpr int kBlockThreads = 64;

// ---------------------------------------------------------------------------------------------------
// math3d.h:31-136 with the operand order spelled out
// ---------------------------------------------------------------------------------------------------
struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 Mk(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 Add(const D3 &a, const D3 &b) { return Mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 Sub(const D3 &a, const D3 &b) { return Mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 Neg(const D3 &a) { return Mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 MulV(const D3 &a, const D3 &b) { return Mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ D3 MulS(const D3 &a, double s) { return Mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 DivS(const D3 &a, double s) { return Mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ double Dot(const D3 &a, const D3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// this.Cross(a), math3d.h:120-126
__device__ __forceinline__ D3 Cross(const D3 &t, const D3 &a) {
  return Mk(t.y * a.z - t.z * a.y, t.z * a.x - t.x * a.z, t.x * a.y - t.y * a.x);
}
__device__ __forceinline__ double SqrLen(const D3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double SqrDist(const D3 &t, const D3 &a) {
  const double dx = a.x - t.x, dy = a.y - t.y, dz = a.z - t.z;
  return dx * dx + dy * dy + dz * dz;
}
__device__ __forceinline__ double Dist(const D3 &t, const D3 &a) { return sqrt(SqrDist(t, a)); }
__device__ __forceinline__ D3 Normalized(const D3 &a) {
  const double l = sqrt(SqrLen(a));
  return Mk(a.x / l, a.y / l, a.z / l);
}
__device__ __forceinline__ D3 Load3(const double *p) { return Mk(p[0], p[1], p[2]); }

// std::min / std::max of libstdc++ (NaN behaviour is contract)
__device__ __forceinline__ double SMin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double SMax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double SMin3(double a, double b, double c) {
  double r = a;
  if (b < r) r = b;
  if (c < r) r = c;
  return r;
}
__device__ __forceinline__ double SMax3(double a, double b, double c) {
  double r = a;
  if (r < b) r = b;
  if (r < c) r = c;
  return r;
}

template <bool DBG>
__device__ __forceinline__ void Count(unsigned long long *cnt, int which, unsigned long long n = 1) {
  if (DBG) cnt[which] += n;
}

// 128-bit read-only loads of the 16-byte aligned records
__device__ __forceinline__ double2 Ld2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

struct Ray {
  D3 o, d, inv;
  bool sx, sy, sz;  // inv component negative (regular rays only)
};

// ---------------------------------------------------------------------------------------------------
// Literal slab test: Node::NodeIntersectRay (octtree.cc:138-167) == the triangle pre-test
// (primitive_triangle.cc:85-108).  Used for irregular rays.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool SlabLiteral(const double lo[3], const double hi[3], const Ray &r, double *tmin_out) {
  const double t1 = (lo[0] - r.o.x) * r.inv.x;
  const double t2 = (hi[0] - r.o.x) * r.inv.x;
  const double t3 = (lo[1] - r.o.y) * r.inv.y;
  const double t4 = (hi[1] - r.o.y) * r.inv.y;
  const double t5 = (lo[2] - r.o.z) * r.inv.z;
  const double t6 = (hi[2] - r.o.z) * r.inv.z;
  const double tmax = SMin3(SMax(t1, t2), SMax(t3, t4), SMax(t5, t6));
  if (tmax < 0.0) return false;
  const double tmin = SMax3(SMin(t1, t2), SMin(t3, t4), SMin(t5, t6));
  if (tmin > tmax) return false;
  *tmin_out = tmin;
  return true;
}

// Regular rays (finite non-zero inverse direction, finite origin): no NaN can arise, and since
// lo <= hi and FP64 subtraction / multiplication are monotonic, min(t_lo, t_hi) is t of the plane the
// direction sign selects.  Same values as SlabLiteral, fewer FP64 compares.
__device__ __forceinline__ bool SlabRegular(double lox, double loy, double loz, double hix, double hiy, double hiz,
                                            const Ray &r, double *tmin_out) {
  const double nx = ((r.sx ? hix : lox) - r.o.x) * r.inv.x;
  const double fx = ((r.sx ? lox : hix) - r.o.x) * r.inv.x;
  const double ny = ((r.sy ? hiy : loy) - r.o.y) * r.inv.y;
  const double fy = ((r.sy ? loy : hiy) - r.o.y) * r.inv.y;
  const double nz = ((r.sz ? hiz : loz) - r.o.z) * r.inv.z;
  const double fz = ((r.sz ? loz : hiz) - r.o.z) * r.inv.z;
  const double tmax = SMin3(fx, fy, fz);
  if (tmax < 0.0) return false;
  const double tmin = SMax3(nx, ny, nz);
  if (tmin > tmax) return false;
  *tmin_out = tmin;
  return true;
}

// Moller-Trumbore exactly as primitive_triangle.cc:111-142.
__device__ __forceinline__ bool MollerTrumbore(const double *vert, const Ray &r, double *t_out) {
  const double2 a = Ld2(vert + 0), b = Ld2(vert + 2), c = Ld2(vert + 4), d = Ld2(vert + 6);
  const double v22 = __ldg(vert + 8);
  const D3 v0 = Mk(a.x, a.y, b.x), v1 = Mk(b.y, c.x, c.y), v2 = Mk(d.x, d.y, v22);
  const D3 e1 = Sub(v1, v0);
  const D3 e2 = Sub(v2, v0);
  const D3 pvec = Cross(r.d, e2);
  const double det = Dot(e1, pvec);
  if (det >= -0.00000001 && det < 0.00000001) return false;
  const double inv_det = 1.0 / det;
  const D3 tvec = Sub(r.o, v0);
  const double u = Dot(tvec, pvec) * inv_det;
  if (u < 0.0 || u > 1.0) return false;
  const D3 qvec = Cross(tvec, e1);
  const double v = Dot(r.d, qvec) * inv_det;
  if (v < 0.0 || u + v > 1.0) return false;
  const double t = Dot(e2, qvec) * inv_det;
  if (t < 0.0) return false;
  *t_out = t;
  return true;
}

// One list entry for a regular ray; candidates arrive in arbitrary order, so the reference's sequential
// "replace unless strictly farther" (octtree.cc:186-195) becomes: nearer wins, on equal t the entry that is
// later in the reference's list (= larger insertion index) wins.
template <bool DBG>
__device__ __forceinline__ void TestSlotRegular(const DeviceScene &sc, int slot, const Ray &r, double *best_t,
                                                int *best_slot, unsigned long long *cnt) {
  const SlotRec *rec = sc.slots + slot;
  const double2 b0 = Ld2(rec->box + 0), b1 = Ld2(rec->box + 2), b2 = Ld2(rec->box + 4);
  Count<DBG>(cnt, kTriAabb);
  double unused;
  if (!SlabRegular(b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, r, &unused)) return;
  Count<DBG>(cnt, kMt);
  double t;
  if (!MollerTrumbore(rec->vert, r, &t)) return;
  Count<DBG>(cnt, kHit);
  if (*best_slot >= 0) {
    if (t > *best_t) return;
    if (t == *best_t && __ldg(&rec->tri) < __ldg(&sc.slots[*best_slot].tri)) return;
  }
  *best_t = t;
  *best_slot = slot;
}

// ---------------------------------------------------------------------------------------------------
// Regular traversal
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__device__ int TraceRegular(const DeviceScene &sc, const Ray &r, double *t_out, unsigned long long *cnt) {
  int f_child[kMaxTreeStack];
  unsigned f_order[kMaxTreeStack];
  double f_t[kMaxTreeStack];
  int f_slot[kMaxTreeStack];

  {  // the root gate (octtree.cc:34-37)
    const NodeRec *root = sc.nodes;
    double unused;
    Count<DBG>(cnt, kSlab);
    if (!SlabRegular(__ldg(&root->planes[0]), __ldg(&root->planes[1]), __ldg(&root->planes[2]), __ldg(&root->planes[6]),
                     __ldg(&root->planes[7]), __ldg(&root->planes[8]), r, &unused)) {
      return -1;
    }
  }

  int sp = 0;
  int cur = 0;
  // current frame in registers
  double c_t = 0.0;
  int c_slot = -1;
  unsigned c_order = 0;
  int c_child = -1;
  for (;;) {
    // ---- enter node `cur`: own list first (octtree.cc:177-196) ----
    const NodeRec *node = sc.nodes + cur;
    const int4 info = __ldg(reinterpret_cast<const int4 *>(&node->first_child));  // first_child, list_first, list_count, bvh_root
    const int2 info2 = __ldg(reinterpret_cast<const int2 *>(&node->child_mask));  // child_mask, bvh_end
    Count<DBG>(cnt, kVisit);
    c_t = 0.0;
    c_slot = -1;
    c_order = 0;
    c_child = info.x;
    if (info.w < 0) {
      for (int s = info.y, e = info.y + info.z; s < e; s++) TestSlotRegular<DBG>(sc, s, r, &c_t, &c_slot, cnt);
    } else {
      int i = info.w;
      const int end = info2.y;
      while (i < end) {
        const BvhRec *b = sc.bvh + i;
        const double2 b0 = Ld2(b->box + 0), b1 = Ld2(b->box + 2), b2 = Ld2(b->box + 4);
        const int4 bi = __ldg(reinterpret_cast<const int4 *>(&b->skip));
        Count<DBG>(cnt, kBvh);
        double unused;
        if (SlabRegular(b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, r, &unused)) {
          for (int s = bi.y, e = bi.y + bi.z; s < e; s++) TestSlotRegular<DBG>(sc, s, r, &c_t, &c_slot, cnt);
          i = i + 1;
        } else {
          i = bi.x;
        }
      }
    }
    // ---- children that the ray enters, ordered by entry distance (octtree.cc:200-216) ----
    const unsigned mask = (unsigned)info2.x;
    if (info.x >= 0 && mask != 0u) {
      const double2 p0 = Ld2(node->planes + 0), p1 = Ld2(node->planes + 2), p2 = Ld2(node->planes + 4), p3 = Ld2(node->planes + 6);
      const double hz = __ldg(&node->planes[8]);
      // planes: lo = (p0.x p0.y p1.x), c = (p1.y p2.x p2.y), hi = (p3.x p3.y hz)
      const double tx0 = (p0.x - r.o.x) * r.inv.x, tx1 = (p1.y - r.o.x) * r.inv.x, tx2 = (p3.x - r.o.x) * r.inv.x;
      const double ty0 = (p0.y - r.o.y) * r.inv.y, ty1 = (p2.x - r.o.y) * r.inv.y, ty2 = (p3.y - r.o.y) * r.inv.y;
      const double tz0 = (p1.x - r.o.z) * r.inv.z, tz1 = (p2.y - r.o.z) * r.inv.z, tz2 = (hz - r.o.z) * r.inv.z;
      // near / far per axis for the lower [lo,c] and upper [c,hi] halves
      const double nx[2] = {r.sx ? tx1 : tx0, r.sx ? tx2 : tx1}, fx[2] = {r.sx ? tx0 : tx1, r.sx ? tx1 : tx2};
      const double ny[2] = {r.sy ? ty1 : ty0, r.sy ? ty2 : ty1}, fy[2] = {r.sy ? ty0 : ty1, r.sy ? ty1 : ty2};
      const double nz[2] = {r.sz ? tz1 : tz0, r.sz ? tz2 : tz1}, fz[2] = {r.sz ? tz0 : tz1, r.sz ? tz1 : tz2};
      double key[8];
      unsigned pass = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int xh = k & 1, zh = (k >> 1) & 1, yh = (k >> 2) & 1;  // octtree.cc:61-100
        const double tmax = SMin3(fx[xh], fy[yh], fz[zh]);
        const double tmin = SMax3(nx[xh], ny[yh], nz[zh]);
        key[k] = tmin;
        if ((mask >> k) & 1u) {
          Count<DBG>(cnt, kSlab);
          if (!(tmax < 0.0) && !(tmin > tmax)) pass |= 1u << k;
        }
      }
      // stable order by key: child j precedes child k (j < k) iff key[j] <= key[k]
      unsigned rank[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int k = j + 1; k < 8; k++) {
          const bool both = ((pass >> j) & (pass >> k) & 1u) != 0u;
          const bool j_first = key[j] <= key[k];
          rank[k] += (both && j_first) ? 1u : 0u;
          rank[j] += (both && !j_first) ? 1u : 0u;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if ((pass >> k) & 1u) c_order |= (8u | (unsigned)k) << (4u * rank[k]);
      }
    }
    // ---- visit children nearest first; unwind finished frames (octtree.cc:219-256) ----
    for (;;) {
      const unsigned e = c_order & 15u;
      if (e != 0u) {
        c_order >>= 4;
        f_child[sp] = c_child;
        f_order[sp] = c_order;
        f_t[sp] = c_t;
        f_slot[sp] = c_slot;
        sp++;
        cur = c_child + (int)(e & 7u);
        break;  // enter the child
      }
      if (sp == 0) {
        if (c_slot < 0) return -1;
        *t_out = c_t;
        return c_slot;
      }
      // return (c_t, c_slot) to the parent frame
      sp--;
      const double rt = c_t;
      const int rs = c_slot;
      c_child = f_child[sp];
      c_order = f_order[sp];
      c_t = f_t[sp];
      c_slot = f_slot[sp];
      if (rs >= 0 && !(c_slot >= 0 && rt > c_t)) {
        c_t = rt;
        c_slot = rs;
        c_order = 0;  // `break`: nodes were sorted by distance (octtree.cc:244-246)
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Literal traversal for irregular rays: no shortcut of any kind.
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__device__ __noinline__ int TraceLiteral(const DeviceScene &sc, const Ray &r, double *t_out, unsigned long long *cnt) {
  int f_node[kMaxTreeStack];
  unsigned f_order[kMaxTreeStack];
  double f_t[kMaxTreeStack];
  int f_slot[kMaxTreeStack];
  Count<DBG>(cnt, kLiteral);
  {
    const NodeRec *root = sc.nodes;
    double unused;
    Count<DBG>(cnt, kSlab);
    if (!SlabLiteral(root->planes, root->planes + 6, r, &unused)) return -1;
  }
  int sp = 0;
  int cur = 0;
  for (;;) {
    const NodeRec *node = sc.nodes + cur;
    Count<DBG>(cnt, kVisit);
    double bt = 0.0;
    int bs = -1;
    for (int k = 0; k < node->list_count; k++) {  // reference list order
      const int slot = sc.list_order[node->list_first + k];
      const SlotRec *rec = sc.slots + slot;
      double unused, t;
      Count<DBG>(cnt, kTriAabb);
      if (!SlabLiteral(rec->box, rec->box + 3, r, &unused)) continue;
      Count<DBG>(cnt, kMt);
      if (!MollerTrumbore(rec->vert, r, &t)) continue;
      Count<DBG>(cnt, kHit);
      if (bs >= 0 && t > bt) continue;
      bt = t;
      bs = slot;
    }
    unsigned order = 0;
    if (node->first_child >= 0) {
      int idx[8];
      double key[8];
      int n = 0;
      for (int k = 0; k < 8; k++) {
        const NodeRec *ch = sc.nodes + node->first_child + k;
        double d;
        Count<DBG>(cnt, kSlab);
        if (!SlabLiteral(ch->planes, ch->planes + 6, r, &d)) continue;
        idx[n] = k;
        key[n] = d;
        n++;
      }
      // libstdc++ __insertion_sort (std::sort on <= 16 elements), comparator a.second < b.second
      for (int i = 1; i < n; i++) {
        const int vi = idx[i];
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
      for (int i = 0; i < n; i++) order |= (8u | (unsigned)idx[i]) << (4u * i);
    }
    f_node[sp] = cur;
    f_order[sp] = order;
    f_t[sp] = bt;
    f_slot[sp] = bs;
    for (;;) {
      const unsigned e = f_order[sp] & 15u;
      if (e != 0u) {
        f_order[sp] >>= 4;
        cur = sc.nodes[f_node[sp]].first_child + (int)(e & 7u);
        sp++;
        break;
      }
      if (sp == 0) {
        if (f_slot[0] < 0) return -1;
        *t_out = f_t[0];
        return f_slot[0];
      }
      const double rt = f_t[sp];
      const int rs = f_slot[sp];
      sp--;
      if (rs >= 0 && !(f_slot[sp] >= 0 && rt > f_t[sp])) {
        f_t[sp] = rt;
        f_slot[sp] = rs;
        f_order[sp] = 0;
      }
    }
  }
}

// OctTree::IntersectRay (octtree.cc:26-40): inverse direction, then one of the two traversals.
template <bool DBG>
__device__ __forceinline__ int Trace(const DeviceScene &sc, const D3 &o, const D3 &d, double *t_out,
                                     unsigned long long *cnt) {
  Ray r;
  r.o = o;
  r.d = d;
  r.inv = Mk(1.0 / d.x, 1.0 / d.y, 1.0 / d.z);
  r.sx = r.inv.x < 0.0;
  r.sy = r.inv.y < 0.0;
  r.sz = r.inv.z < 0.0;
  Count<DBG>(cnt, kRays);
  const bool regular = isfinite(r.inv.x) && isfinite(r.inv.y) && isfinite(r.inv.z) && r.inv.x != 0.0 &&
                       r.inv.y != 0.0 && r.inv.z != 0.0 && isfinite(o.x) && isfinite(o.y) && isfinite(o.z);
  if (regular) return TraceRegular<DBG>(sc, r, t_out, cnt);
  return TraceLiteral<DBG>(sc, r, t_out, cnt);
}

// ---------------------------------------------------------------------------------------------------
// shading helpers
// ---------------------------------------------------------------------------------------------------
// primitive_triangle.cc:27-40
__device__ __forceinline__ double Heron(double a, double b, double c) {
  const double p = (a + b + c) / 2.0;
  const double area_sqr = p * (p - a) * (p - b) * (p - c);
  if (area_sqr < 0.0) return 0.0;
  return sqrt(area_sqr);
}

struct BaryWeights {
  double n0, n1, n2, n;
};
// primitive_triangle.cc:45-57 (shared by GetNormal and GetUVW)
__device__ __forceinline__ BaryWeights Barycentric(const D3 &v0, const D3 &v1, const D3 &v2, const D3 &point) {
  const double a = Dist(v0, v1);
  const double b = Dist(v1, v2);
  const double c = Dist(v2, v0);
  const double p0 = Dist(point, v0);
  const double p1 = Dist(point, v1);
  const double p2 = Dist(point, v2);
  BaryWeights w;
  w.n0 = Heron(b, p2, p1);
  w.n1 = Heron(c, p0, p2);
  w.n2 = Heron(a, p1, p0);
  w.n = w.n0 + w.n1 + w.n2;
  return w;
}

// texture.cc:11-58 with point fetches of the 8-bit texels; px / 255.0 as texture.cc:100-104.
__device__ __forceinline__ D3 Texel(cudaTextureObject_t tex, size_t x, size_t y) {
  const uchar4 p = tex2D<uchar4>(tex, (float)x + 0.5f, (float)y + 0.5f);
  return Mk((double)p.x / 255.0, (double)p.y / 255.0, (double)p.z / 255.0);
}
__device__ D3 SampleTexture(cudaTextureObject_t tex, int2 dim, double u, double v) {
  u = fmod(u, 1.0);
  v = fmod(v, 1.0);
  if (u < 0.0) u += 1.0;
  if (v < 0.0) v += 1.0;
  v = 1.0 - v;
  const size_t width = (size_t)dim.x, height = (size_t)dim.y;
  const double x = u * (double)(width - 1);
  const double y = v * (double)(height - 1);
  size_t bx = (size_t)x;
  size_t by = (size_t)y;
  // (size_t)NaN is undefined upstream (vector::at would throw); stay inside the texture here
  if (bx >= width) bx = width - 1;
  if (by >= height) by = height - 1;
  const size_t bx1 = (bx + 1 == width) ? bx : bx + 1;
  const size_t by1 = (by + 1 == height) ? by : by + 1;
  const D3 c0 = Texel(tex, bx, by), c1 = Texel(tex, bx1, by), c2 = Texel(tex, bx, by1), c3 = Texel(tex, bx1, by1);
  const double dx = fmod(x, 1.0);
  const double dy = fmod(y, 1.0);
  const double a0 = (1.0 - dx) * (1.0 - dy);
  const double a1 = dx * (1.0 - dy);
  const double a2 = (1.0 - dx) * dy;
  const double a3 = dx * dy;
  return Add(Add(Add(MulS(c0, a0), MulS(c1, a1)), MulS(c2, a2)), MulS(c3, a3));
}

// mythtracer.cc:235-241.  (uint8_t)(NaN * 255) is 0 on x86-64; restated explicitly.
__device__ __forceinline__ unsigned char QuantizeChannel(double v) {
  if (v > 1.0) return 255;
  if (v < 0.0) return 0;
  if (v != v) return 0;
  return (unsigned char)(int)(v * 255);
}

__device__ __forceinline__ unsigned long long Mix64(unsigned long long path, unsigned long long kind,
                                                    unsigned long long value) {
  unsigned long long z = path * 0x9E3779B97F4A7C15ull + kind * 0xC2B2AE3D27D4EB4Full +
                         value * 0x165667B19E3779F9ull + 0x27D4EB2F165667C5ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// One suspended TraceRayWorker activation waiting for its reflection / refraction child.
struct ShadeFrame {
  D3 color;
  D3 point;
  D3 dir;        // ray.direction of this activation (the refraction child continues along it)
  double coef;   // current_reflection_coef
  unsigned long long path;
  int material;
  unsigned char stage;      // 0: reflection child pending, 1: refraction child pending
  unsigned char do_refract;
  unsigned char in_object;
  unsigned char pad_;
};

// ---------------------------------------------------------------------------------------------------
// RenderMega: one thread = one pixel = the whole TraceRay recursion, evaluated in the reference's
// post-order so that colour sums associate identically.  Every loop iteration issues exactly one
// OctTree::IntersectRay-equivalent query (a primary / reflection / refraction ray or one shadow segment),
// so the lanes of a warp reconverge at the single Trace call site.
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__global__ void __launch_bounds__(kBlockThreads) RenderMega(DeviceScene sc, RenderParams rp) {
  // block -> 8x8 tile of one of this launch's strips (a strip = 8 image rows; strips are interleaved
  // across devices / processes, the in-process form of the reference's master/worker tiling)
  const int strip = rp.strip_first + ((int)blockIdx.x / rp.tiles_x) * rp.strip_stride;
  const int px = ((int)blockIdx.x % rp.tiles_x) * kTile + (int)(threadIdx.x & 7u);
  const int py = strip * kTile + (int)(threadIdx.x >> 3);
  const bool live = px < rp.chunk_w && py < rp.chunk_h;

  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int i = 0; i < kNumCounters; i++) cnt[i] = 0;
  }
  unsigned long long sig_hits = 0, sig_shadow = 0;
  unsigned n_rays = 0;

  if (live) {
    ShadeFrame stack[kMaxRayStack];
    int sp = 0;

    // Sensor::GetRay (camera.cc:65-69) with full-image pixel coordinates (mythtracer.cc:298)
    const D3 start = Load3(rp.sensor), d_scan = Load3(rp.sensor + 3), d_pixel = Load3(rp.sensor + 6);
    D3 m_o = Load3(rp.origin);
    D3 m_d = Normalized(Add(Add(start, MulS(d_scan, (double)(rp.chunk_y + py))), MulS(d_pixel, (double)(rp.chunk_x + px))));
    int level = 0;
    bool in_object = false;
    double coef = 1.0;
    unsigned long long path = 1;
    Count<DBG>(cnt, kPrimary);

    // shading context of the current activation (valid while its shadow rays are traced)
    D3 P = Mk(0, 0, 0), normal = Mk(0, 0, 0), surface = Mk(0, 0, 0), reflected = Mk(0, 0, 0), color = Mk(0, 0, 0);
    int material = -1;
    // shadow walk state (mythtracer.cc:86-156)
    int li = 0;
    D3 ldir = Mk(0, 0, 0), lpos = Mk(0, 0, 0), power = Mk(0, 0, 0), seg_start = Mk(0, 0, 0);
    bool in_shadow = false, through = false;
    unsigned segments = 0;
    bool shadow_mode = false;
    D3 final_color = Mk(0, 0, 0);

    for (;;) {
      D3 to, td;
      double light_distance = 0.0;
      if (shadow_mode) {
        to = Add(seg_start, MulS(ldir, 0.00001));  // mythtracer.cc:95-99
        td = ldir;
        light_distance = Dist(seg_start, lpos);    // mythtracer.cc:101-102
        Count<DBG>(cnt, kShadow);
      } else {
        to = m_o;
        td = m_d;
      }
      double t = 0.0;
      const int slot = Trace<DBG>(sc, to, td, &t, cnt);
      n_rays++;

      bool have_ret = false;
      D3 ret = Mk(0.0, 0.0, 0.0);
      if (shadow_mode) {
        bool light_done = false;
        segments++;
        if (slot < 0) {
          light_done = true;  // mythtracer.cc:109-112
        } else if (t > light_distance) {
          light_done = true;  // mythtracer.cc:115-118
        } else {
          const int smtl = __ldg(&sc.shade[slot].material);
          // mtl is dereferenced unconditionally upstream (mythtracer.cc:121); a missing material acts opaque
          const double str = smtl >= 0 ? __ldg(&sc.materials[smtl].transparency) : 0.0;
          if (str == 0.0) {
            power = Mk(0.0, 0.0, 0.0);
            in_shadow = true;
            light_done = true;
          } else {
            if (!through) {  // light_power *= Tf * Tr (mythtracer.cc:129-132)
              const D3 tf = Load3(sc.materials[smtl].transmission_filter);
              power = MulV(power, MulS(tf, str));
            }
            through = !through;
            const D3 hit_point = Add(to, MulS(td, t));               // primitive_triangle.cc:141
            seg_start = Add(hit_point, MulS(ldir, 0.0000001));       // mythtracer.cc:137
            if (SqrDist(P, seg_start) > SqrDist(P, lpos)) {          // mythtracer.cc:141-145
              light_done = true;
            } else if (power.x <= 0.001 && power.y <= 0.001 && power.z <= 0.001) {
              power = Mk(0.0, 0.0, 0.0);
              in_shadow = true;
              light_done = true;
            }
          }
        }
        if (!light_done) continue;  // next segment of the same light

        // ---- this light is settled: Phong terms (mythtracer.cc:159-177) ----
        const mtb_light *lt = sc.lights + li;
        const mtb_material *m = sc.materials + material;
        sig_shadow += Mix64(path, 2ull + (unsigned long long)li, (in_shadow ? 1ull : 0ull) | ((unsigned long long)segments << 1));
        const D3 lamb = Load3(lt->ambient);
        power.x = SMax(power.x, lamb.x);
        power.y = SMax(power.y, lamb.y);
        power.z = SMax(power.z, lamb.z);
        color = Add(color, MulV(MulV(MulS(MulV(Load3(m->diffuse), surface), Dot(normal, ldir)), Load3(lt->diffuse)), power));
        if (!in_shadow) {
          const double refl_dot = Dot(Neg(m_d), reflected);
          if (refl_dot > 0) {
            color = Add(color, MulV(MulS(MulV(Load3(m->specular), surface), pow(refl_dot, m->specular_exp)), Load3(lt->specular)));
          }
        }
        li++;
      } else {
        // ---- result of a primary / reflection / refraction ray (mythtracer.cc:13-76) ----
        if (slot < 0) {
          if (level == 0 && rp.dbg != nullptr) {
            mtb_debug *dbg = rp.dbg + (size_t)py * rp.chunk_w + px;
            dbg->line_no = -1;
            dbg->pad_ = 0;
            dbg->point[0] = dbg->point[1] = dbg->point[2] = CUDART_NAN;
          }
          have_ret = true;  // background colour {0,0,0}
        } else {
          const ShadeRec *sh = sc.shade + slot;
          const SlotRec *sr = sc.slots + slot;
          P = Add(to, MulS(td, t));
          const int line_no = __ldg(&sh->line_no);
          if (level == 0 && rp.dbg != nullptr) {
            mtb_debug *dbg = rp.dbg + (size_t)py * rp.chunk_w + px;
            dbg->line_no = line_no;
            dbg->pad_ = 0;
            dbg->point[0] = P.x;
            dbg->point[1] = P.y;
            dbg->point[2] = P.z;
          }
          sig_hits += Mix64(path, 1ull, (unsigned long long)(long long)line_no);
          Count<DBG>(cnt, kShade);
          const D3 v0 = Load3(sr->vert), v1 = Load3(sr->vert + 3), v2 = Load3(sr->vert + 6);
          const BaryWeights w = Barycentric(v0, v1, v2, P);
          // (normal[0]*n0 + normal[1]*n1 + normal[2]*n2) / n, not normalised (primitive_triangle.cc:60)
          normal = DivS(Add(Add(MulS(Load3(sh->normal), w.n0), MulS(Load3(sh->normal + 3), w.n1)), MulS(Load3(sh->normal + 6), w.n2)), w.n);
          const D3 towards_camera = Neg(m_d);
          double normal_ray_dot = Dot(towards_camera, normal);
          if (normal_ray_dot < 0.0) {
            normal = Neg(normal);
            normal_ray_dot = Dot(towards_camera, normal);
          }
          material = __ldg(&sh->material);
          if (material < 0) {  // mythtracer.cc:49-52
            normal_ray_dot = (normal_ray_dot + 1.0) * 0.5;
            ret = Mk(normal_ray_dot, normal_ray_dot, normal_ray_dot);
            have_ret = true;
          } else {
            const mtb_material *m = sc.materials + material;
            surface = Load3(m->ambient);
            const int tex = m->texture;
            if (tex >= 0) {  // mythtracer.cc:59-64
              const double u = (sh->uv[0] * w.n0 + sh->uv[2] * w.n1 + sh->uv[4] * w.n2) / w.n;
              const double v = (sh->uv[1] * w.n0 + sh->uv[3] * w.n1 + sh->uv[5] * w.n2) / w.n;
              surface = MulV(surface, SampleTexture(sc.textures[tex], sc.texture_dim[tex], u, v));
            }
            // ray.direction - normal * (2 * ray.direction.Dot(normal)) (mythtracer.cc:68-69)
            reflected = Sub(m_d, MulS(normal, 2 * Dot(normal, m_d)));
            color = Mk(0.0, 0.0, 0.0);
            li = 0;
          }
        }
      }

      if (!have_ret) {
        if (li < sc.n_lights) {
          // ---- start the shadow walk of light li (mythtracer.cc:79-94) ----
          const mtb_light *lt = sc.lights + li;
          lpos = Load3(lt->position);
          ldir = Normalized(Sub(lpos, P));
          color = Add(color, MulV(Load3(lt->ambient), surface));  // mythtracer.cc:83-84
          power = Mk(1.0, 1.0, 1.0);
          in_shadow = false;
          through = false;
          segments = 0;
          seg_start = P;
          shadow_mode = true;
          continue;
        }
        // ---- all lights done: secondary rays (mythtracer.cc:181-225) ----
        shadow_mode = false;
        const mtb_material *m = sc.materials + material;
        const double refl = m->reflectance, tr = m->transparency;
        const bool do_reflect = level < rp.max_depth && refl > 0.0 && coef > 0.01 && !in_object;
        const bool do_refract = level < rp.max_depth && tr > 0.0;
        if (do_reflect || do_refract) {
          ShadeFrame &f = stack[sp];
          f.color = color;
          f.point = P;
          f.dir = m_d;
          f.coef = coef;
          f.path = path;
          f.material = material;
          f.stage = do_reflect ? 0 : 1;
          f.do_refract = do_refract ? 1 : 0;
          f.in_object = in_object ? 1 : 0;
          sp++;
          level++;
          if (do_reflect) {
            Count<DBG>(cnt, kReflect);
            m_o = Add(P, MulS(reflected, 0.0001));  // mythtracer.cc:70-75
            m_d = reflected;
            coef = coef * refl;
            path = path * 2ull;
          } else {
            Count<DBG>(cnt, kRefract);
            const D3 rdir = Normalized(m_d);       // mythtracer.cc:208-212
            m_o = Add(P, MulS(rdir, 0.00001));     // mythtracer.cc:214-218
            m_d = rdir;
            in_object = !in_object;
            path = path * 2ull + 1ull;
          }
          continue;
        }
        ret = color;
      }

      // ---- an activation returned `ret`: fold it into suspended parents (mythtracer.cc:185-189,220-224) ----
      bool finished = false;
      for (;;) {
        if (sp == 0) {
          final_color = ret;
          finished = true;
          break;
        }
        ShadeFrame &f = stack[sp - 1];
        const mtb_material *m = sc.materials + f.material;
        if (f.stage == 0) {
          f.color = Add(f.color, MulS(ret, m->reflectance));
          if (f.do_refract) {
            f.stage = 1;
            Count<DBG>(cnt, kRefract);
            const D3 rdir = Normalized(f.dir);
            m_o = Add(f.point, MulS(rdir, 0.00001));
            m_d = rdir;
            in_object = !(f.in_object != 0);
            coef = f.coef;
            path = f.path * 2ull + 1ull;
            level = sp;
            shadow_mode = false;
            break;  // trace the refraction child
          }
          ret = f.color;
          sp--;
        } else {
          // (c * Tf) * Tr (mythtracer.cc:224)
          f.color = Add(f.color, MulS(MulV(ret, Load3(m->transmission_filter)), m->transparency));
          ret = f.color;
          sp--;
        }
      }
      if (finished) break;
    }

    unsigned char *out = rp.rgb + ((size_t)py * rp.chunk_w + px) * 3;
    out[0] = QuantizeChannel(final_color.x);
    out[1] = QuantizeChannel(final_color.y);
    out[2] = QuantizeChannel(final_color.z);
    const size_t pix = (size_t)py * rp.chunk_w + px;
    if (rp.sig_hits != nullptr) rp.sig_hits[pix] = sig_hits;
    if (rp.sig_shadow != nullptr) rp.sig_shadow[pix] = sig_shadow;
    if (rp.n_rays != nullptr) rp.n_rays[pix] = n_rays;
  }

  if (DBG && rp.counters != nullptr) {
    for (int i = 0; i < kNumCounters; i++) {
      unsigned long long v = cnt[i];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(rp.counters + i, v);
    }
  } else if (rp.counters != nullptr) {
    // the fast build still reports the ray count (the metric's numerator)
    unsigned long long v = n_rays;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(rp.counters + kRays, v);
  }
}

// Batched OctTree::IntersectRay (octtree.cc:26-40), one ray per thread.
template <bool DBG>
__global__ void __launch_bounds__(128) IntersectKernel(DeviceScene sc, IntersectParams ip) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  if (i < ip.n) {
    const D3 o = Load3(ip.origins + i * 3), d = Load3(ip.dirs + i * 3);
    double t = 0.0;
    const int slot = Trace<DBG>(sc, o, d, &t, cnt);
    if (slot < 0) {
      ip.tri_index[i] = -1;
    } else {
      ip.tri_index[i] = sc.slots[slot].tri;
      if (ip.t != nullptr) ip.t[i] = t;
      if (ip.point != nullptr) {
        const D3 p = Add(o, MulS(d, t));
        ip.point[i * 3 + 0] = p.x;
        ip.point[i * 3 + 1] = p.y;
        ip.point[i * 3 + 2] = p.z;
      }
    }
  }
  if (DBG && ip.counters != nullptr) {
    for (int k = 0; k < kNumCounters; k++) {
      unsigned long long v = cnt[k];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(ip.counters + k, v);
    }
  }
}

}  // namespace

void LaunchRenderMega(const DeviceScene &sc, const RenderParams &rp, int n_blocks, bool debug_build,
                      cudaStream_t stream) {
  if (n_blocks <= 0) return;
  if (debug_build) {
    RenderMega<true><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
  } else {
    RenderMega<false><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
  }
}

void LaunchIntersect(const DeviceScene &sc, const IntersectParams &ip, bool debug_build, cudaStream_t stream) {
  if (ip.n <= 0) return;
  const int blocks = (int)((ip.n + 127) / 128);
  if (debug_build) {
    IntersectKernel<true><<<blocks, 128, 0, stream>>>(sc, ip);
  } else {
    IntersectKernel<false><<<blocks, 128, 0, stream>>>(sc, ip);
  }
}

}  // namespace mtb
