// reference VerStarting/primitive_triangle.h:10-30
#pragma once
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

  V3D vertex[3]{};
  V3D normal[3]{};
  V3D uvw[3]{};
  AABB cached_aabb;
};

}  // namespace raytracer
