// reference VerStarting/objreader.h:13-64.  Parsing is done by the library's loader (same observable
// behaviour as objreader.cc, see mythtracer_b200/csrc/obj_loader.cc); these classes fill a Scene from it.
#pragma once
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../mythtracer_b200.h"
#include "primitive_triangle.h"
#include "scene.h"

namespace raytracer {

namespace detail {
// Copies materials / textures of a host-only loader context into the Scene; returns index -> Material*.
inline std::vector<Material *> ImportMaterials(const mtb_context *ctx, Scene *scene) {
  mtb_scene_summary info;
  mtb_scene_info(ctx, &info);
  std::vector<Texture *> tex_by_index;
  for (int32_t i = 0; i < info.n_textures; i++) {
    mtb_texture t;
    mtb_scene_texture(ctx, i, &t);
    std::unique_ptr<Texture> tex(new Texture);
    tex->width = (size_t)t.width;
    tex->height = (size_t)t.height;
    tex->rgba.assign(t.rgba, t.rgba + (size_t)t.width * t.height * 4);
    tex->colors.resize((size_t)t.width * t.height);
    for (size_t k = 0; k < tex->colors.size(); k++) {  // texture.cc:100-104
      tex->colors[k] = V3D{(double)t.rgba[k * 4] / 255.0, (double)t.rgba[k * 4 + 1] / 255.0, (double)t.rgba[k * 4 + 2] / 255.0};
    }
    tex_by_index.push_back(tex.get());
    scene->textures[mtb_scene_texture_name(ctx, i)] = std::move(tex);
  }
  std::vector<mtb_material> mtls((size_t)info.n_materials);
  mtb_scene_read(ctx, nullptr, mtls.data());
  std::vector<Material *> out;
  for (int32_t i = 0; i < info.n_materials; i++) {
    const mtb_material &m = mtls[(size_t)i];
    std::unique_ptr<Material> mat(new Material);
    mat->ambient = V3D{m.ambient[0], m.ambient[1], m.ambient[2]};
    mat->diffuse = V3D{m.diffuse[0], m.diffuse[1], m.diffuse[2]};
    mat->specular = V3D{m.specular[0], m.specular[1], m.specular[2]};
    mat->transmission_filter = V3D{m.transmission_filter[0], m.transmission_filter[1], m.transmission_filter[2]};
    mat->specular_exp = m.specular_exp;
    mat->reflectance = m.reflectance;
    mat->transparency = m.transparency;
    mat->refraction_index = m.refraction_index;
    mat->tex = m.texture >= 0 ? tex_by_index[(size_t)m.texture] : nullptr;
    out.push_back(mat.get());
    scene->materials[mtb_scene_material_name(ctx, i)] = std::move(mat);
  }
  return out;
}
}  // namespace detail

class MtlFileReader {
 public:
  bool ReadMtlFile(Scene *scene, const char *fname) {  // objreader.cc:472-549
    mtb_context *ctx = nullptr;
    if (mtb_create_host(&ctx) != MTB_OK) return false;
    const bool ok = mtb_load_mtl(ctx, fname) == MTB_OK;
    if (ok) detail::ImportMaterials(ctx, scene);
    mtb_destroy(ctx);
    return ok;
  }
};

class ObjFileReader {
 public:
  bool ReadObjFile(Scene *scene, const char *fname) {  // objreader.cc:201-274
    mtb_context *ctx = nullptr;
    if (mtb_create_host(&ctx) != MTB_OK) return false;
    if (mtb_load_obj(ctx, fname) != MTB_OK) {
      mtb_destroy(ctx);
      return false;
    }
    const std::vector<Material *> mtl = detail::ImportMaterials(ctx, scene);
    mtb_scene_summary info;
    mtb_scene_info(ctx, &info);
    std::vector<mtb_triangle> tris((size_t)info.n_triangles);
    mtb_scene_read(ctx, tris.data(), nullptr);
    for (const mtb_triangle &t : tris) {
      Triangle *tr = new Triangle();
      for (int k = 0; k < 3; k++) {
        tr->vertex[k] = V3D{t.vertex[k * 3], t.vertex[k * 3 + 1], t.vertex[k * 3 + 2]};
        tr->normal[k] = V3D{t.normal[k * 3], t.normal[k * 3 + 1], t.normal[k * 3 + 2]};
        tr->uvw[k] = V3D{t.uvw[k * 3], t.uvw[k * 3 + 1], t.uvw[k * 3 + 2]};
      }
      tr->mtl = t.material >= 0 ? mtl[(size_t)t.material] : nullptr;
      tr->debug_line_no = t.line_no;
      tr->CacheAABB();
      scene->tree.AddPrimitive(tr);  // the tree owns it from here (objreader.cc:186)
    }
    mtb_destroy(ctx);
    return true;
  }
};

}  // namespace raytracer
