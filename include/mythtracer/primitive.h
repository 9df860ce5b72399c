// reference VerStarting/primitive.h:13-44.  Intersection, normal and UV evaluation are devirtualised on
// the device (Triangle is the only concrete primitive); the host-side interface keeps
// what callers and the octree builder need.
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
  virtual std::string Serialize() const = 0;

  Material *mtl = nullptr;  // not owned
  int debug_line_no = 0;
};

}  // namespace raytracer
