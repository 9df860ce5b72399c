/* TEST INFRASTRUCTURE ONLY (oracle/): C interface of the CPU restatement of MythTracer's ray-casting path.
 *
 * This library is the parity CHECKER.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it; the product (mythtracer_b200/) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement bit-for-bit (RGB bytes,
 * primary-hit line numbers, hit points, t values) against the unmodified reference compiled into
 * oracle/_ref/libmythtracer_ref.so, against the reference's own known-answer vectors
 * (octtree_test.cc:35-69, math3d_test.cc:68-89) and against committed fixtures in tests/golden/ that were
 * produced by the reference itself (tests/golden/make_golden.py).
 *
 * The struct layouts are deliberately the same as the product's C ABI (include/mythtracer_b200.h) so the
 * same numpy buffers can be handed to both sides of a parity test.
 */
#ifndef MT_ORACLE_H_
#define MT_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mto_scene mto_scene;

/* reference primitive_triangle.h:26-29 + primitive.h:41-42 */
typedef struct {
  double vertex[9];
  double normal[9];
  double uvw[9];
  int32_t material; /* index into the material array, -1 = mtl == nullptr */
  int32_t line_no;  /* Primitive::debug_line_no */
} mto_triangle;

/* reference material.h:12-48 */
typedef struct {
  double ambient[3], diffuse[3], specular[3];
  double specular_exp, reflectance, transparency;
  double transmission_filter[3];
  double refraction_index;
  int32_t texture; /* index into the texture array, -1 = none */
  int32_t pad_;
} mto_material;

/* RGBA32 texels as the reference holds them after SDL conversion (texture.cc:81-104) */
typedef struct {
  int32_t width, height;
  const uint8_t *rgba;
} mto_texture;

/* reference light.h:8-14 */
typedef struct {
  double position[3], ambient[3], diffuse[3], specular[3];
} mto_light;

/* reference camera.h:31-33 (the 7 doubles Camera::Serialize writes, camera.cc:71-81) */
typedef struct {
  double origin[3];
  double pitch, yaw, roll, aov;
} mto_camera;

/* Work counters; the per-ray figures SURVEY.md section 8(d) builds the roofline from. */
typedef struct {
  uint64_t rays;          /* OctTree::IntersectRay calls = primary + shadow segments + reflect + refract */
  uint64_t primary, shadow, reflect, refract;
  uint64_t n_slab;        /* Node::NodeIntersectRay calls              (octtree.cc:138) */
  uint64_t n_visit;       /* Node::PrimitiveIntersectRay calls         (octtree.cc:169) */
  uint64_t n_triaabb;     /* Triangle::IntersectRay calls (AABB pre-test, primitive_triangle.cc:85) */
  uint64_t n_mt;          /* Moller-Trumbore evaluations               (primitive_triangle.cc:111) */
  uint64_t n_hit;         /* triangle tests that returned true */
  uint64_t n_shade;       /* shaded hits                               (mythtracer.cc:38) */
} mto_stats;

/* Per-pixel decision taps (all optional).  Signatures are order independent sums of mto_mix64 events:
 *   sig_hits   += mix(path, 1, tri_line_no)            for every non-shadow ray that hit (path: root = 1,
 *                                                       reflection child = 2p, refraction child = 2p+1)
 *   sig_shadow += mix(path, 2 + light, in_shadow | segments << 1)  for every (shaded hit, light)
 *   n_rays     = OctTree::IntersectRay calls spent on the pixel */
typedef struct {
  uint64_t *sig_hits;
  uint64_t *sig_shadow;
  uint32_t *n_rays;
} mto_taps;

mto_scene *mto_create(const mto_triangle *tris, int64_t n_tris, const mto_material *mtls, int32_t n_mtls,
                      const mto_texture *texs, int32_t n_texs);
void mto_destroy(mto_scene *s);
void mto_set_lights(mto_scene *s, const mto_light *lights, int32_t n);
void mto_set_threads(int32_t n); /* 0 = all */
int32_t mto_get_threads(void);

/* octree shape: nodes, depth, biggest list, root list length, triangles kept in interior nodes */
void mto_tree_info(const mto_scene *s, int64_t out[5]);
void mto_scene_aabb(const mto_scene *s, double out6[6]);
/* per triangle: box (lo.xyz, hi.xyz) and depth of the octree node whose list holds it */
void mto_triangle_nodes(const mto_scene *s, double *node_box, int32_t *node_depth);

/* MythTracer::RayTrace(WorkChunk*) (mythtracer.cc:280-312) with MAX_RECURSION_LEVEL = max_depth.
 * rgb: chunk_w*chunk_h*3.  dbg_line_no / dbg_point: PerPixelDebugInfo of the primary hit (may be NULL). */
int mto_render(const mto_scene *s, const mto_camera *cam, int image_w, int image_h, int chunk_x, int chunk_y,
               int chunk_w, int chunk_h, int max_depth, uint8_t *rgb, int32_t *dbg_line_no, double *dbg_point,
               const mto_taps *taps, mto_stats *stats);

/* Unquantised colour of the same pixels (3 doubles per pixel), for tolerance-free colour comparisons. */
int mto_render_color(const mto_scene *s, const mto_camera *cam, int image_w, int image_h, int chunk_x,
                     int chunk_y, int chunk_w, int chunk_h, int max_depth, double *color);

/* OctTree::IntersectRay (octtree.cc:26-40), batched.  tri_index = insertion index, -1 on a miss. */
int mto_intersect(const mto_scene *s, int64_t n, const double *origins, const double *dirs, int32_t *tri_index,
                  double *t, double *point, mto_stats *stats);

/* Brute force over all triangles in insertion order with the reference's own per-triangle test and
 * replace-unless-strictly-farther rule: used to show the octree is a pure
 * accelerator on the test scenes. */
int mto_intersect_brute(const mto_scene *s, int64_t n, const double *origins, const double *dirs,
                        int32_t *tri_index, double *t);

/* Camera::GetSensor + Sensor::GetRay (camera.cc:17-69): direction of pixel (x, y) of a w x h image. */
void mto_camera_ray(const mto_camera *cam, int image_w, int image_h, int x, int y, double dir[3]);
/* The three sensor vectors start_point, delta_scanline, delta_pixel (camera.cc:56-62). */
void mto_camera_sensor(const mto_camera *cam, int image_w, int image_h, double out9[9]);

/* Texture::GetColorAt (texture.cc:11-58) on texture `tex` of the scene. */
void mto_texture_sample(const mto_scene *s, int32_t tex, double u, double v, double out[3]);

/* Triangle::GetNormal / GetUVW (primitive_triangle.cc:43-79) of triangle `tri` at `point`. */
void mto_triangle_normal(const mto_scene *s, int64_t tri, const double point[3], double out[3]);
void mto_triangle_uvw(const mto_scene *s, int64_t tri, const double point[3], double out[3]);

/* MythTracer::V3DtoRGB (mythtracer.cc:235-241). */
void mto_quantize(const double color[3], uint8_t rgb[3]);

/* math3d.h known answers (math3d_test.cc:68-89): out = Length(a), a.Distance(b), a.Dot(b), a.Cross(b)[3],
 * a.DupNorm()[3]  (9 doubles). */
void mto_math3d(const double a[3], const double b[3], double out[9]);

uint64_t mto_mix64(uint64_t path, uint64_t kind, uint64_t value);

#ifdef __cplusplus
}
#endif
#endif /* MT_ORACLE_H_ */
