// reference VerStarting/material.h:12-50
#pragma once
#include <memory>
#include <string>
#include <unordered_map>

#include "math3d.h"
#include "texture.h"

namespace raytracer {
using math3d::V3D;

class Material {
 public:
  V3D ambient{}, diffuse{}, specular{};  // Ka Kd Ks
  Texture *tex = nullptr;                // map_Ka, not owned
  V3D::basetype specular_exp = 0.0;      // Ns
  V3D::basetype reflectance = 0.0;       // Refl
  V3D::basetype transparency = 0.0;      // Tr
  V3D transmission_filter{};             // Tf
  V3D::basetype refraction_index = 0.0;  // Ni
};

typedef std::unordered_map<std::string, std::unique_ptr<Material>> MaterialMap;

}  // namespace raytracer
