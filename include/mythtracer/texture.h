// reference VerStarting/texture.h:13-25.  Texels are kept as 8-bit RGBA (what SDL hands the reference,
// texture.cc:81-104) and as the px / 255.0 doubles of Texture::colors.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "math3d.h"

namespace raytracer {
using math3d::V3D;

class Texture {
 public:
  size_t width = 0;
  size_t height = 0;
  std::vector<V3D> colors;
  std::vector<uint8_t> rgba;  // extension: the 8-bit source texels the device samples
};

typedef std::unordered_map<std::string, std::unique_ptr<Texture>> TextureMap;

}  // namespace raytracer
