// The reference's single-process animation driver (VerStarting/main_local.cc:24-155) on the B200 renderer
// (SURVEY.md section 8 f3): orbiting camera, lights re-set every frame, raw RGB24 frames to
// anim/dump_%05i.raw.  The model path, resolution and frame range are arguments instead of constants.
//
//   mythtracer_local_b200 scene.obj [width height first_frame last_frame out_dir]
#include <sys/stat.h>
#include <sys/types.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <mythtracer/mythtracer.h>

using raytracer::AABB;
using raytracer::Camera;
using raytracer::Light;
using raytracer::MythTracer;

int main(int argc, char **argv) {
  if (argc < 2) {
    puts("usage: mythtracer_local_b200 scene.obj [width height first_frame last_frame out_dir]");
    return 1;
  }
  const int W = argc > 2 ? atoi(argv[2]) : 1920 / 4, H = argc > 3 ? atoi(argv[3]) : 1080 / 4;
  const int first = argc > 4 ? atoi(argv[4]) : 74, last = argc > 5 ? atoi(argv[5]) : 180;
  const char *out_dir = argc > 6 ? argv[6] : "anim";
  printf("Creating %s/ directory\n", out_dir);
  mkdir(out_dir, 0700);
  printf("Resolution: %u %u\n", W, H);

  MythTracer mt;
  if (!mt.LoadObj(argv[1])) return 1;
  const AABB aabb = mt.GetScene()->tree.GetAABB();
  printf("%f %f %f x %f %f %f\n", aabb.min.v[0], aabb.min.v[1], aabb.min.v[2], aabb.max.v[0], aabb.max.v[1], aabb.max.v[2]);

  // Two pinned frames in rotation: while frame k+1 renders (queued without waiting), frame k is written to disk.
  uint8_t *frames[2] = {MythTracer::AllocFrame(W, H), MythTracer::AllocFrame(W, H)};
  if (frames[0] == nullptr || frames[1] == nullptr) return 1;
  const size_t frame_bytes = (size_t)W * H * 3;
  auto write_frame = [&](int number, const uint8_t *pixels) {
    puts("Writing");
    char fname[512];
    snprintf(fname, sizeof(fname), "%s/dump_%.5i.raw", out_dir, number);
    FILE *f = fopen(fname, "wb");
    if (f == nullptr) return false;
    fwrite(pixels, frame_bytes, 1, f);
    fclose(f);
    return true;
  };
  int frame = 0;
  int rendered = 0, pending = -1;
  const auto t0 = std::chrono::steady_clock::now();
  for (double angle = 0.0; angle <= 360.0; angle += 2.0, frame++) {
    if (frame < first || frame > last) continue;
    Camera cam{{300.0, 107.0, 40.0}, 30.0, angle + 90, 0.0, 110.0};  // main_local.cc:72-76
    auto &lights = mt.GetScene()->lights;                              // main_local.cc:79-110
    lights.clear();
    lights.push_back(Light{{231.82174, 81.69966, -27.78259}, {0.3, 0.3, 0.3}, {1.0, 1.0, 1.0}, {1.0, 1.0, 1.0}});
    for (double z : {0.0, 80.0, 160.0}) lights.push_back(Light{{200, 80.0, z}, {0.0, 0.0, 0.0}, {0.3, 0.3, 0.3}, {0.3, 0.3, 0.3}});
    puts("Rendering.");
    if (!mt.RayTraceAsync(W, H, &cam, frames[rendered & 1])) return 1;
    if (pending >= 0 && !write_frame(pending, frames[(rendered + 1) & 1])) return 1;  // overlaps the render
    if (!mt.Wait()) return 1;
    pending = frame;
    rendered++;
  }
  if (pending >= 0 && !write_frame(pending, frames[(rendered + 1) & 1])) return 1;
  const double total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  MythTracer::FreeFrame(frames[0]);
  MythTracer::FreeFrame(frames[1]);
  printf("Done: %d frames, %.2f ms per frame (render + copies + file writes, double-buffered)\n", rendered, rendered ? total_ms / rendered : 0.0);
  return 0;
}
