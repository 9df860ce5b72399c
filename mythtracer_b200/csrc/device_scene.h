// Structures shared by the host API (api.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mythtracer_b200.h"
#include "scene_build.h"

namespace mtb {

// Everything the kernels read, resident in the HBM of one device.
struct DeviceScene {
  const NodeRec *nodes;
  const SlotRec *slots;
  const ShadeRec *shade;
  const BvhRec *bvh;
  const int32_t *list_order;
  const mtb_material *materials;
  const cudaTextureObject_t *textures;  // uchar4 point-sampled texture objects
  const int2 *texture_dim;
  const mtb_light *lights;
  int32_t n_lights;
  int32_t n_materials;
};

// Work counters, one slot per field of mtb_stats' integer part (same order).
enum Counter {
  kRays = 0, kPrimary, kShadow, kReflect, kRefract, kSlab, kVisit, kTriAabb, kMt, kHit, kShade, kBvh, kLiteral,
  kNumCounters
};

struct RenderParams {
  double sensor[9];   // start_point, delta_scanline, delta_pixel (camera.cc:56-62)
  double origin[3];
  int32_t image_w, image_h;
  int32_t chunk_x, chunk_y, chunk_w, chunk_h;
  int32_t max_depth;
  // The chunk is cut into strips of 8 rows; this launch renders strips strip_first + i * strip_stride
  // (multi-GPU / multi-process interleave).  Block b <-> 8x8 tile (b % tiles_x) of local strip b / tiles_x.
  int32_t tiles_x, strip_first, strip_stride;
  uint8_t *rgb;        // chunk-local RGB24, stride chunk_w*3 (may be peer memory of device 0)
  mtb_debug *dbg;      // nullable
  uint64_t *sig_hits;  // nullable taps
  uint64_t *sig_shadow;
  uint32_t *n_rays;
  unsigned long long *counters;  // kNumCounters, nullable unless the counting build runs
};

struct IntersectParams {
  int64_t n;
  const double *origins, *dirs;
  int32_t *tri_index;
  double *t, *point;
  unsigned long long *counters;
};

// kernels.cu
void LaunchRenderMega(const DeviceScene &sc, const RenderParams &rp, int n_blocks, bool debug_build,
                      cudaStream_t stream);
void LaunchIntersect(const DeviceScene &sc, const IntersectParams &ip, bool debug_build, cudaStream_t stream);

}  // namespace mtb
