// reference VerStarting/primitive.h:13-44.  On the device intersection, normal and UV evaluation are devirtualised
// (Triangle is the only concrete primitive); the host-side interface keeps the reference's
// virtuals so that code written against Primitive* keeps compiling - they evaluate on the host, in the reference's
// FP64 operation order (build the caller with -ffp-contract=off to get the reference's bits: its binary has no FMA).
#pragma once
#include <string>

#include "aabb.h"
#include "material.h"
#include "math3d.h"
#include "ray.h"

namespace raytracer {

class Primitive {
 public:
  virtual ~Primitive() {}
  virtual AABB GetAABB() const = 0;
  // primitive.h:20-24: true when the ray hits, with the hit point and its distance from the ray origin
  virtual bool IntersectRay(const Ray &ray, V3D *point, V3D::basetype *distance) const = 0;
  virtual V3D GetNormal(const V3D &point) const = 0;  // primitive.h:27
  virtual V3D GetUVW(const V3D &point) const = 0;     // primitive.h:30
  virtual std::string Serialize() const = 0;

  Material *mtl = nullptr;  // not owned
  int debug_line_no = 0;
};

}  // namespace raytracer
