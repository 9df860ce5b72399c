// reference VerStarting/octtree.h:14-39.  AddPrimitive / Finalize / IntersectRay / GetAABB keep their
// signatures; Finalize hands the primitives to mtb_scene_upload (which builds the reference's octree and
// keeps it in HBM) and IntersectRay is served by the batched CUDA query mtb_intersect_rays.
#pragma once
#include <cstdio>
#include <list>
#include <memory>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../mythtracer_b200.h"
#include "math3d.h"
#include "primitive.h"
#include "primitive_triangle.h"

namespace raytracer {
using math3d::V3D;

class OctTree {
 public:
  OctTree() {}
  ~OctTree() {
    if (ctx_ != nullptr) mtb_destroy(ctx_);
  }
  OctTree(const OctTree &) = delete;
  OctTree &operator=(const OctTree &) = delete;

  // Takes ownership of `p` (octtree.h:18-22).  The root box only ever grows from {0,0,0} (octtree.cc:12-13).
  void AddPrimitive(Primitive *p) {
    primitives_.push_back(std::unique_ptr<Primitive>(p));
    root_aabb_.Extend(p->GetAABB());
    finalized_ = false;
  }

  // octtree.cc:16-24.  Standalone trees (no Scene) upload triangles without materials.
  void Finalize() { Finalize(nullptr, nullptr); }

  // octtree.cc:26-40; nullptr on a miss or when no CUDA device is available (there is no CPU fallback).
  const Primitive *IntersectRay(const Ray &ray, V3D *point, V3D::basetype *distance) const {
    if (!finalized_ || ctx_ == nullptr) return nullptr;
    int32_t tri = -1;
    double t = 0.0, p[3] = {0, 0, 0};
    if (mtb_intersect_rays(ctx_, 1, ray.origin.v, ray.direction.v, &tri, &t, p, nullptr) != MTB_OK || tri < 0) return nullptr;
    *point = V3D{p[0], p[1], p[2]};
    *distance = t;
    return index_[(size_t)tri];
  }

  // Batched form of the above (extension): tri_index[i] = position in AddPrimitive order or -1.
  bool IntersectRays(int64_t n, const double *origins, const double *dirs, int32_t *tri_index, double *t, double *point) const {
    return finalized_ && ctx_ != nullptr && mtb_intersect_rays(ctx_, n, origins, dirs, tri_index, t, point, nullptr) == MTB_OK;
  }

  AABB GetAABB() const { return root_aabb_; }

  // ---- extensions used by Scene / MythTracer ----
  void Clear() {
    primitives_.clear();
    index_.clear();
    root_aabb_ = AABB{};
    finalized_ = false;
  }
  size_t size() const { return primitives_.size(); }
  const Primitive *at(size_t i) const { return index_.empty() ? nullptr : index_[i]; }
  bool finalized() const { return finalized_; }
  mtb_context *context() const { return ctx_; }
  const char *last_error() const { return ctx_ != nullptr ? mtb_last_error(ctx_) : mtb_last_error(nullptr); }
  // CUDA devices the tree (and the renderer on top of it) uses; call before the first Finalize.
  void SetDevices(const std::vector<int> &devices) { devices_ = devices; }
  void SetFlags(uint32_t flags) {
    flags_ = flags;
    if (ctx_ != nullptr) mtb_set_flags(ctx_, flags_);
  }

  // Finalize with the scene's material / texture tables (MythTracer::RayTrace's lazy Finalize,
  // mythtracer.cc:281-285).  Returns false when no device / upload fails.
  bool Finalize(const MaterialMap *materials, const TextureMap *textures) {
    printf("Triangles: %u\n", (unsigned int)primitives_.size());  // as octtree.cc:17
    if (ctx_ == nullptr) {
      if (mtb_create(&ctx_, devices_.empty() ? nullptr : devices_.data(), (int)devices_.size()) != MTB_OK) {
        fprintf(stderr, "error: %s\n", mtb_last_error(nullptr));
        ctx_ = nullptr;
        return false;
      }
      mtb_set_flags(ctx_, flags_);
    }
    std::unordered_map<const Texture *, int32_t> tex_index;
    std::vector<mtb_texture> texs;
    if (textures != nullptr) {
      for (const auto &kv : *textures) {
        tex_index[kv.second.get()] = (int32_t)texs.size();
        texs.push_back(mtb_texture{(int32_t)kv.second->width, (int32_t)kv.second->height, kv.second->rgba.data()});
      }
    }
    std::unordered_map<const Material *, int32_t> mtl_index;
    std::vector<mtb_material> mtls;
    if (materials != nullptr) {
      for (const auto &kv : *materials) {
        const Material &m = *kv.second;
        mtb_material o{};
        for (int c = 0; c < 3; c++) {
          o.ambient[c] = m.ambient.v[c];
          o.diffuse[c] = m.diffuse.v[c];
          o.specular[c] = m.specular.v[c];
          o.transmission_filter[c] = m.transmission_filter.v[c];
        }
        o.specular_exp = m.specular_exp;
        o.reflectance = m.reflectance;
        o.transparency = m.transparency;
        o.refraction_index = m.refraction_index;
        const auto it = m.tex != nullptr ? tex_index.find(m.tex) : tex_index.end();
        o.texture = it != tex_index.end() ? it->second : -1;
        mtl_index[&m] = (int32_t)mtls.size();
        mtls.push_back(o);
      }
    }
    std::vector<mtb_triangle> tris;
    tris.reserve(primitives_.size());
    index_.clear();
    index_.reserve(primitives_.size());
    for (const auto &p : primitives_) {
      const Triangle *t = dynamic_cast<const Triangle *>(p.get());
      if (t == nullptr) continue;  // Triangle is the only primitive the renderer knows
      mtb_triangle o{};
      for (int k = 0; k < 3; k++) {
        for (int c = 0; c < 3; c++) {
          o.vertex[k * 3 + c] = t->vertex[k].v[c];
          o.normal[k * 3 + c] = t->normal[k].v[c];
          o.uvw[k * 3 + c] = t->uvw[k].v[c];
        }
      }
      const auto it = t->mtl != nullptr ? mtl_index.find(t->mtl) : mtl_index.end();
      o.material = it != mtl_index.end() ? it->second : -1;
      o.line_no = t->debug_line_no;
      tris.push_back(o);
      index_.push_back(p.get());
    }
    const int rc = mtb_scene_upload(ctx_, tris.data(), (int64_t)tris.size(), mtls.data(), (int32_t)mtls.size(), texs.data(),
                                    (int32_t)texs.size());
    if (rc != MTB_OK) {
      fprintf(stderr, "error: %s\n", mtb_last_error(ctx_));
      return false;
    }
    finalized_ = true;
    return true;
  }

 private:
  std::list<std::unique_ptr<Primitive>> primitives_;
  std::vector<const Primitive *> index_;  // AddPrimitive order -> primitive (device hit index -> pointer)
  AABB root_aabb_;
  bool finalized_ = false;
  mtb_context *ctx_ = nullptr;
  std::vector<int> devices_;
  uint32_t flags_ = 0;
};

}  // namespace raytracer
