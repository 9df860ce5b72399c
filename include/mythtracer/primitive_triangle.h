// reference VerStarting/primitive_triangle.h:10-30, primitive_triangle.cc:18-143
#pragma once
#include <algorithm>
#include <cmath>
#include <memory>
#include <string>

#include "primitive.h"

namespace raytracer {

class Triangle : public Primitive {
 public:
  ~Triangle() override {}
  AABB GetAABB() const override { return cached_aabb; }
  std::string Serialize() const override { return "nope"; }  // a stub upstream too (primitive_triangle.cc:145-147)
  static bool Deserialize(std::unique_ptr<Triangle> *, const std::string &) { return false; }

  void CacheAABB() {  // primitive_triangle.cc:18-24
    AABB box{vertex[0], vertex[0]};
    box.Extend(vertex[1]);
    box.Extend(vertex[2]);
    cached_aabb = box;
  }

  // primitive_triangle.cc:81-143 on the host: slab pre-test against the cached box, then Moller-Trumbore with the
  // reference's thresholds.  Upstream reads the ray's private inv_direction, which only OctTree::IntersectRay fills
  // in (octtree.cc:30-32); a direct caller gets it computed here the same way.  The renderer does not come through
  // here - the CUDA kernels run the same arithmetic (csrc/device_core.cuh: SlabRegular / SlabLiteral, MollerTrumbore).
  bool IntersectRay(const Ray &ray, V3D *point, V3D::basetype *distance) const override {
    using T = V3D::basetype;
    T near_t[3], far_t[3];
    for (int a = 0; a < 3; a++) {
      const T inv = 1.0 / ray.direction.v[a];
      const T lo = (cached_aabb.min.v[a] - ray.origin.v[a]) * inv, hi = (cached_aabb.max.v[a] - ray.origin.v[a]) * inv;
      near_t[a] = std::min(lo, hi);  // std::min / std::max keep the first argument on NaN, as upstream
      far_t[a] = std::max(lo, hi);
    }
    const T leave = std::min({far_t[0], far_t[1], far_t[2]});
    if (leave < 0.0) return false;  // the box lies behind the origin
    const T enter = std::max({near_t[0], near_t[1], near_t[2]});
    if (enter > leave) return false;

    const V3D edge1 = vertex[1] - vertex[0], edge2 = vertex[2] - vertex[0];
    const V3D p = ray.direction.Cross(edge2);
    const T det = edge1.Dot(p);
    if (det >= -0.00000001 && det < 0.00000001) return false;  // parallel to the plane
    const T inv_det = 1.0 / det;
    const V3D from_v0 = ray.origin - vertex[0];
    const T u = from_v0.Dot(p) * inv_det;
    if (u < 0.0 || u > 1.0) return false;
    const V3D q = from_v0.Cross(edge1);
    const T v = ray.direction.Dot(q) * inv_det;
    if (v < 0.0 || u + v > 1.0) return false;
    const T t = edge2.Dot(q) * inv_det;
    if (t < 0.0) return false;  // behind the origin
    *distance = t;
    *point = ray.origin + ray.direction * t;
    return true;
  }

  // primitive_triangle.cc:42-79: both interpolate with sub-triangle areas from Heron's formula (not normalised)
  V3D GetNormal(const V3D &point) const override { return Blend(normal, point); }
  V3D GetUVW(const V3D &point) const override { return Blend(uvw, point); }

  V3D vertex[3]{};
  V3D normal[3]{};
  V3D uvw[3]{};
  AABB cached_aabb;

 private:
  static V3D::basetype Heron(V3D::basetype a, V3D::basetype b, V3D::basetype c) {  // primitive_triangle.cc:27-40
    const V3D::basetype s = (a + b + c) / 2.0;
    const V3D::basetype sq = s * (s - a) * (s - b) * (s - c);
    return sq < 0.0 ? 0.0 : std::sqrt(sq);  // collinear points can give a tiny negative value
  }
  V3D Blend(const V3D attr[3], const V3D &point) const {
    const V3D::basetype side01 = vertex[0].Distance(vertex[1]), side12 = vertex[1].Distance(vertex[2]), side20 = vertex[2].Distance(vertex[0]);
    const V3D::basetype d0 = point.Distance(vertex[0]), d1 = point.Distance(vertex[1]), d2 = point.Distance(vertex[2]);
    const V3D::basetype w0 = Heron(side12, d2, d1), w1 = Heron(side20, d0, d2), w2 = Heron(side01, d1, d0);
    const V3D::basetype total = w0 + w1 + w2;
    return (attr[0] * w0 + attr[1] * w1 + attr[2] * w2) / total;
  }
};

}  // namespace raytracer
