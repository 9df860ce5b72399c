// reference VerStarting/camera.h:12-46, camera.cc:9-96
#pragma once
#include <stdint.h>

#include <cstring>
#include <vector>

#include "../mythtracer_b200.h"
#include "math3d.h"
#include "ray.h"

namespace raytracer {
using math3d::M4D;
using math3d::V3D;

class Camera {
 public:
  class Sensor {
   public:
    // camera.cc:65-69: full-image pixel coordinates
    Ray GetRay(int x, int y) const {
      V3D direction = start_point + (delta_scanline * y) + (delta_pixel * x);
      direction.Norm();
      return {origin_, direction};
    }

   private:
    V3D delta_scanline, delta_pixel, start_point, origin_;
    int width = 0, height = 0;
    friend Camera;
  };

  V3D origin;
  V3D::basetype pitch, yaw, roll;  // degrees about X, Y, Z
  V3D::basetype aov;               // angle of view, degrees

  V3D GetDirection() const {  // camera.cc:9-15
    V3D dir{0.0, 0.0, 1.0};
    return M4D::RotationYDeg(yaw) * M4D::RotationXDeg(pitch) * dir;
  }

  // camera.cc:17-63: the three sensor vectors come from the library, so host and device agree bit for bit
  Sensor GetSensor(int width, int height) const {
    Sensor s;
    s.width = width;
    s.height = height;
    s.origin_ = origin;
    const mtb_camera c = AsMtb();
    double v[9];
    mtb_camera_sensor(&c, width, height, v);
    s.start_point = V3D{v[0], v[1], v[2]};
    s.delta_scanline = V3D{v[3], v[4], v[5]};
    s.delta_pixel = V3D{v[6], v[7], v[8]};
    return s;
  }

  static const size_t kSerializedSize = sizeof(V3D) + 4 * sizeof(V3D::basetype);  // 56 bytes

  void Serialize(std::vector<uint8_t> *bytes) {  // camera.cc:71-81
    const mtb_camera c = AsMtb();
    bytes->resize(kSerializedSize);
    memcpy(bytes->data(), &c, kSerializedSize);
  }
  bool Deserialize(const std::vector<uint8_t> &bytes) {  // camera.cc:83-96
    if (bytes.size() != kSerializedSize) return false;
    mtb_camera c;
    memcpy(&c, bytes.data(), kSerializedSize);
    origin = V3D{c.origin[0], c.origin[1], c.origin[2]};
    pitch = c.pitch;
    yaw = c.yaw;
    roll = c.roll;
    aov = c.aov;
    return true;
  }

  mtb_camera AsMtb() const { return mtb_camera{{origin.v[0], origin.v[1], origin.v[2]}, pitch, yaw, roll, aov}; }
};
static_assert(sizeof(mtb_camera) == 56, "Camera wire form");

}  // namespace raytracer
