// A caller written against the REFERENCE's interface (same call pattern as VerStarting/main_local.cc:24-155
// and VerStarting/octtree_test.cc:14-72), compiled against include/mythtracer/*.h.  Usage:
//   shim_check host <scene.obj>                       loader / wire-form checks, no GPU needed
//   shim_check render <scene.obj> <out.raw> <w> <h>   full path on the GPU, raw RGB24 like main_local.cc:127-132
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <mythtracer/mythtracer.h>

using math3d::V3D;
using raytracer::AABB;
using raytracer::Camera;
using raytracer::Light;
using raytracer::MythTracer;
using raytracer::OctTree;
using raytracer::PerPixelDebugInfo;
using raytracer::Primitive;
using raytracer::Ray;
using raytracer::Triangle;
using raytracer::WorkChunk;

static int Fail(const char *what) {
  fprintf(stderr, "FAIL: %s\n", what);
  return 1;
}

// tri <in.bin> <out.bin>: records of 15 doubles (3 vertices, ray origin, ray direction) -> records of 5 doubles
// (hit flag, distance, hit point) from Triangle::IntersectRay called through Primitive*.
static int TriangleBatch(const char *in_path, const char *out_path) {
  FILE *in = fopen(in_path, "rb"), *out = fopen(out_path, "wb");
  if (!in || !out) return Fail("tri: files");
  double rec[15];
  while (fread(rec, sizeof(double), 15, in) == 15) {
    Triangle tri;
    for (int v = 0; v < 3; v++) tri.vertex[v] = {rec[3 * v], rec[3 * v + 1], rec[3 * v + 2]};
    tri.CacheAABB();
    const Primitive *prim = &tri;
    V3D point{};
    double res[5] = {0, 0, 0, 0, 0};
    if (prim->IntersectRay(Ray({rec[9], rec[10], rec[11]}, {rec[12], rec[13], rec[14]}), &point, &res[1])) {
      res[0] = 1.0;
      res[2] = point.v[0], res[3] = point.v[1], res[4] = point.v[2];
    }
    fwrite(res, sizeof(double), 5, out);
  }
  fclose(in);
  fclose(out);
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 4 && strcmp(argv[1], "tri") == 0) return TriangleBatch(argv[2], argv[3]);
  if (argc < 3) return Fail("usage");
  MythTracer mt;
  if (!mt.LoadObj(argv[2])) return Fail("LoadObj");
  AABB aabb = mt.GetScene()->tree.GetAABB();
  printf("%f %f %f x %f %f %f\n", aabb.min.v[0], aabb.min.v[1], aabb.min.v[2], aabb.max.v[0], aabb.max.v[1], aabb.max.v[2]);
  printf("triangles %zu materials %zu textures %zu\n", mt.GetScene()->tree.size(), mt.GetScene()->materials.size(),
         mt.GetScene()->textures.size());

  // wire forms (mythtracer.cc:314-429, camera.cc:71-96)
  Camera cam{{301.37, 57.21, 161.13}, 4.0, 243.0, 0.0, 110.0};
  std::vector<uint8_t> blob;
  cam.Serialize(&blob);
  Camera back{};
  if (blob.size() != 56 || !back.Deserialize(blob) || back.yaw != 243.0 || back.origin.v[2] != 161.13) return Fail("Camera wire form");
  WorkChunk wc{1920, 1080, 128, 256, 128, 128, cam, {}, {}};
  wc.SerializeInput(&blob);
  WorkChunk wd{};
  if (blob.size() != 24 || !wd.DeserializeInput(blob) || wd.chunk_y != 256) return Fail("WorkChunk input wire form");
  blob[8] = 0xff, blob[9] = 0xff;  // chunk_x beyond the image
  if (wd.DeserializeInput(blob)) return Fail("WorkChunk bounds check");
  Ray r = cam.GetSensor(1920, 1080).GetRay(960, 540);
  if (!(std::fabs(r.direction.Length() - 1.0) < 1e-12)) return Fail("Sensor::GetRay");
  V3D a{1, 2, 3}, b{5, 4, 3};  // math3d_test.cc:68-89
  if (a.Dot(b) != 22.0 || a.Cross(b).v[1] != 12.0 || std::fabs(a.Length() - 3.7416573867739413) > 1e-12) return Fail("math3d");
  {  // the Primitive virtuals (primitive.h:20-30), called the way reference code would: through Primitive*
    Triangle tri;
    tri.vertex[0] = {0, 0, 0}, tri.vertex[1] = {1, 0, 0}, tri.vertex[2] = {0, 1, 0};
    tri.normal[0] = {1, 0, 0}, tri.normal[1] = {0, 1, 0}, tri.normal[2] = {0, 0, 1};
    tri.uvw[0] = {0, 0, 0}, tri.uvw[1] = {1, 0, 0}, tri.uvw[2] = {0, 1, 0};
    tri.CacheAABB();
    const Primitive *prim = &tri;
    V3D hit{};
    double dist = -1.0;
    if (!prim->IntersectRay(Ray({0.25, 0.25, 1.0}, {0, 0, -1}), &hit, &dist) || dist != 1.0 || hit.v[0] != 0.25 || hit.v[1] != 0.25 ||
        hit.v[2] != 0.0)
      return Fail("Triangle::IntersectRay hit");
    if (prim->IntersectRay(Ray({0.75, 0.75, 1.0}, {0, 0, -1}), &hit, &dist)) return Fail("Triangle::IntersectRay outside (u + v > 1)");
    if (prim->IntersectRay(Ray({0.25, 0.25, 1.0}, {0, 0, 1}), &hit, &dist)) return Fail("Triangle::IntersectRay behind the origin");
    if (prim->IntersectRay(Ray({0.25, 0.25, 1.0}, {1, 0, 0}), &hit, &dist)) return Fail("Triangle::IntersectRay parallel");
    const V3D uvw = prim->GetUVW({0.25, 0.5, 0.0}), nrm = prim->GetNormal({0.25, 0.5, 0.0});
    if (std::fabs(uvw.v[0] - 0.25) > 1e-12 || std::fabs(uvw.v[1] - 0.5) > 1e-12 || std::fabs(nrm.v[0] - 0.25) > 1e-12 ||
        std::fabs(nrm.v[1] - 0.25) > 1e-12 || std::fabs(nrm.v[2] - 0.5) > 1e-12)
      return Fail("Triangle::GetUVW / GetNormal");
  }
  if (strcmp(argv[1], "host") == 0) {
    puts("host ok");
    return 0;
  }

  if (argc < 6) return Fail("usage: render <obj> <out.raw> <w> <h>");
  const int W = atoi(argv[4]), H = atoi(argv[5]);
  mt.SetMaxRecursionLevel(3);
  mt.GetScene()->lights.clear();  // main_local.cc:79-110
  mt.GetScene()->lights.push_back(Light{{231.82174, 81.69966, 27.78259}, {0.3, 0.3, 0.3}, {1.0, 1.0, 1.0}, {1.0, 1.0, 1.0}});
  mt.GetScene()->lights.push_back(Light{{200, 95.0, 160}, {0.0, 0.0, 0.0}, {0.3, 0.3, 0.3}, {0.3, 0.3, 0.3}});
  std::vector<uint8_t> bitmap;
  if (!mt.RayTrace(W, H, &cam, &bitmap)) return Fail("RayTrace(w, h, cam, out)");
  if (bitmap.size() != (size_t)W * H * 3) return Fail("bitmap size");
  // the same frame as 2 WorkChunks with debug taps (main_net_worker.cc:147-150)
  std::vector<uint8_t> tiled((size_t)W * H * 3);
  for (int part = 0; part < 2; part++) {
    WorkChunk chunk{W, H, 0, part * (H / 2), W, part == 0 ? H / 2 : H - H / 2, cam, {}, {}};
    chunk.output_bitmap.resize((size_t)chunk.chunk_width * chunk.chunk_height * 3);
    chunk.output_debug.resize((size_t)chunk.chunk_width * chunk.chunk_height);
    if (!mt.RayTrace(&chunk)) return Fail("RayTrace(WorkChunk*)");
    memcpy(&tiled[(size_t)chunk.chunk_y * W * 3], chunk.output_bitmap.data(), chunk.output_bitmap.size());
    const PerPixelDebugInfo &d = chunk.output_debug[0];
    const Ray pr = cam.GetSensor(W, H).GetRay(0, chunk.chunk_y);
    V3D p;
    double dist = 0;
    const Primitive *hit = mt.GetScene()->tree.IntersectRay(pr, &p, &dist);
    if ((hit == nullptr) != (d.line_no < 0)) return Fail("debug tap vs OctTree::IntersectRay");
    if (hit != nullptr && (hit->debug_line_no != d.line_no || p.v[0] != d.point.v[0])) return Fail("debug tap mismatch");
  }
  if (tiled != bitmap) return Fail("tiles != frame");
  FILE *f = fopen(argv[3], "wb");
  fwrite(&bitmap[0], bitmap.size(), 1, f);
  fclose(f);

  // octtree_test.cc:14-72 (standalone tree, CacheAABB as the loader does)
  OctTree tree;
  Triangle *tr0 = new Triangle();
  tr0->vertex[0] = {1, 1, 0}, tr0->vertex[1] = {1, 0, 0}, tr0->vertex[2] = {0, 0, 0};
  tr0->CacheAABB();
  tree.AddPrimitive(tr0);
  Triangle *tr1 = new Triangle();
  tr1->vertex[0] = {1, 1, 1}, tr1->vertex[1] = {1, 0, 1}, tr1->vertex[2] = {0, 0, 1};
  tr1->CacheAABB();
  tree.AddPrimitive(tr1);
  tree.Finalize();
  V3D point;
  V3D::basetype distance;
  if (tree.IntersectRay(Ray{{0.9, 0.9, -10.0}, {0.0, 0.0, 1.0}}, &point, &distance) != tr0) return Fail("front ray");
  if (tree.IntersectRay(Ray{{0.9, 0.9, 10.0}, {0.0, 0.0, -1.0}}, &point, &distance) != tr1) return Fail("back ray");
  if (tree.IntersectRay(Ray{{5.0, 5.0, 5.0}, {0.0, 0.0, 1.0}}, &point, &distance) != nullptr) return Fail("missing ray");
  puts("render ok");
  return 0;
}
