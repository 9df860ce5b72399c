// reference VerStarting/aabb.h:7-17, aabb.cc:5-47
#pragma once
#include <cmath>
#include <utility>

#include "math3d.h"

namespace raytracer {
using math3d::V3D;

class AABB {
 public:
  bool Contains(const V3D &p) const {  // closed intervals
    for (int i = 0; i < 3; i++)
      if (!(p.v[i] >= min.v[i] && p.v[i] <= max.v[i])) return false;
    return true;
  }
  bool FullyContains(const AABB &o) const { return Contains(o.min) && Contains(o.max); }
  bool Contains(const AABB &o) const {  // overlap test on centres and extents
    const auto a = GetCenterWHD(), b = o.GetCenterWHD();
    for (int i = 0; i < 3; i++)
      if (!(std::fabs(a.first.v[i] - b.first.v[i]) * 2.0 <= a.second.v[i] + b.second.v[i])) return false;
    return true;
  }
  void Extend(const V3D &p) {
    for (int i = 0; i < 3; i++) {
      min.v[i] = (p.v[i] < min.v[i]) ? p.v[i] : min.v[i];
      max.v[i] = (max.v[i] < p.v[i]) ? p.v[i] : max.v[i];
    }
  }
  void Extend(const AABB &o) {
    Extend(o.min);
    Extend(o.max);
  }
  std::pair<V3D, V3D> GetCenterWHD() const { return {min + (max - min) / 2, max - min}; }

  V3D min, max;
};

}  // namespace raytracer
