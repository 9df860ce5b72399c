// Wavefront pipeline: the recursive TraceRayWorker (mythtracer.cc:13-228) as an iterative, level-by-level
// sequence of kernels over ray queues in HBM.
//
//   level L:  WfTrace       one thread per queued ray: OctTree::IntersectRay, hit point, interpolated normal,
//                           surface colour, reflected direction (mythtracer.cc:18-76) -> activation record; then, in
//                           the same kernel, the reflection / refraction children of the hit are appended to the queue
//                           of level L+1 with a warp-aggregated (ballot + prefix count) slot allocation
//                           (mythtracer.cc:181-225).  Whether a child exists does not depend on the lights.
//     side streams:
//             WfShadow      one thread per (hit, light): the whole shadow walk through transparent surfaces
//                           (mythtracer.cc:86-156); lights are independent of each other, only the order in
//                           which their terms are summed matters, and that order is kept by WfLight
//             WfLight       one thread per hit: Phong sum over the lights in scene order (mythtracer.cc:78-178)
//   The trace chain of the levels is the critical path (each link ends with the slowest ray of its level); the
//   shadow / light kernels of level L only need level L's activation records, so they run on side streams and fill
//   the machine while the chain advances.
//   finally:  WfFold        deepest level first: parent += child * Refl, then parent += (child * Tf) * Tr --
//                           the same two additions, in the same order, as the recursion performs on return
//             WfResolve     V3DtoRGB (mythtracer.cc:235-241) into the chunk-local RGB24 buffer
//             WfCommit      work counters of the frame -> the context's counters
//
// Nothing is read back to the host while a frame is in flight (round 1 read one counter per level): how many rays
// a level holds is a device-side number (WfBuffers::level_n), every kernel sizes its own loop from it, and the host
// only chooses grid sizes - from the previous frame's counts when it has them, else from the capacity bound.  A
// queue that overflows sets WfBuffers::ctrl[0]; every later kernel of the frame then exits at once and a RenderMega
// launch that is always queued behind the frame (RenderParams::run_if) renders the same pixels instead - both
// pipelines produce the same bytes - while the host enlarges the queues for the next frame.
//
// Because every activation keeps its own colour and the fold replays the reference's additions in the
// reference's order, the result is bit-identical to the megakernel (and to the reference, up to pow()).
#include <cstdlib>

#include "device_core.cuh"

namespace mtb {
namespace {

constexpr int kWfBlock = 128;
#ifndef MTB_WF_MIN_BLOCKS
#define MTB_WF_MIN_BLOCKS 8  // measured on B200 (C3): 1 -> 107 ms, 4 -> 90, 6 -> 79, 8 -> 70 (64 registers, 32 warps/SM)
#endif

__device__ __forceinline__ void Store3(double *p, const D3 &v) {
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}

template <bool DBG>
__device__ __forceinline__ void FlushCounters(unsigned long long *cnt, unsigned long long *global, unsigned n_rays) {
  if (global == nullptr) return;
  if (DBG) {
    for (int i = 0; i < kNumCounters; i++) {
      unsigned long long v = cnt[i];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(global + i, v);
    }
  } else {
    unsigned long long v = n_rays;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(global + kRays, v);
  }
}

// Rays queued for `level` and the activation id of its first ray (activations are numbered level by level).
struct LevelRange {
  int n, base;
};
__device__ __forceinline__ LevelRange WfLevel(const WfBuffers &wf, int level) {
  LevelRange r;
  r.base = 0;
  for (int k = 0; k < level; k++) r.base += (int)wf.level_n[k];
  r.n = (int)wf.level_n[level];
  return r;
}

// Sub-warp packing.  A warp runs as long as its slowest ray and serialises what its lanes do differently,
// so a level with few rays (deep levels, or a small share of the frame on one of 8 GPUs) finishes sooner
// when its rays are spread over more, emptier warps: only the first `lanes` lanes of each warp get a ray.
// The loops below are warp-uniform: a warp handles items [first, first + lanes), then advances by the grid.
struct WarpSpan {
  long long first, step;
  int lane;
};
__device__ __forceinline__ WarpSpan WfSpan(int lanes) {
  WarpSpan s;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  s.lane = (int)(threadIdx.x & 31u);
  s.first = warp * lanes;
  s.step = (((long long)gridDim.x * blockDim.x) >> 5) * lanes;
  return s;
}

// ---------------------------------------------------------------------------------------------------
// WfTrace
// ---------------------------------------------------------------------------------------------------
// Pixel slot -> tile of this launch.  Hybrid frames (rp.heavy_k): slot >> 6 is a position in the launch order and
// only the first *heavy_k positions belong to the wavefront (-1 beyond them).
__device__ __forceinline__ int WfTileOfSlot(const RenderParams &rp, int slot) {
  const int pos = slot >> 6;
  if (rp.heavy_k == nullptr) return pos;
  if (pos >= __ldg(rp.heavy_k)) return -1;
  return __ldg(rp.tile_order + pos);
}

// What a tile cost, for the next frame's launch order and split (the megakernel does the same per block).
__device__ __forceinline__ void WfChargeTile(const RenderParams &rp, int pixel, unsigned rays) {
  if (rp.tile_cost == nullptr || rays == 0u || pixel < 0) return;
  const int px = pixel % rp.chunk_w, py = pixel / rp.chunk_w;
  const int local_strip = ((py >> 3) - rp.strip_first) / rp.strip_stride;
  atomicAdd(rp.tile_cost + local_strip * rp.tiles_x + (px >> 3), rays);
}

// First kernel of a frame: level 0 holds one ray per pixel slot, nothing else is queued, no overflow, no work yet.
__global__ void WfBegin(WfBuffers wf, int slots) {
  const int i = (int)threadIdx.x;
  if (i <= MTB_MAX_RAY_DEPTH + 1) wf.level_n[i] = i == 0 ? (uint32_t)slots : 0u;
  if (i < 2) wf.ctrl[i] = 0u;
  if (i < kNumCounters) wf.work[i] = 0ull;
}

template <bool DBG>
__global__ void __launch_bounds__(kWfBlock, MTB_WF_MIN_BLOCKS) WfTrace(DeviceScene sc, RenderParams rp, WfBuffers wf, int level, int lanes) {
  MTB_DECLARE_FAST_CTX(kWfBlock);
  if (wf.ctrl[0] != 0u) return;  // an earlier level overflowed: the frame is rendered by the repair launch
  const LevelRange lr = WfLevel(wf, level);
  const int n = lr.n < wf.queue_cap ? lr.n : wf.queue_cap, act_base = lr.base;
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  unsigned traced = 0;
  const int q = level & 1, qn = q ^ 1;
  const WarpSpan span = WfSpan(lanes);
  for (long long first = span.first; first < n; first += span.step) {
    // i: position in the level's queue; activation id = act_base + i
    const int i = (int)first + span.lane;
    const bool mine = span.lane < lanes && i < n;
    bool do_reflect = false, do_refract = false, in_object = false;
    double coef = 1.0, refl = 0.0;
    D3 P = Mk(0.0, 0.0, 0.0), child_refl = Mk(0.0, 0.0, 0.0), d = Mk(0.0, 0.0, 0.0);
    int pixel = -1;
    unsigned long long path = 1ull;
    const int act = act_base + i;
    if (mine) {
      D3 o;
      bool live = true;
      if (level == 0) {
        // 8x8 pixel tiles (a warp = 8x4 pixels) of this launch's strips, as in RenderMega; in a hybrid frame the
        // slots are the first *heavy_k tiles of the launch order
        const int tile = WfTileOfSlot(rp, i), t = i & 63;
        const int strip = rp.strip_first + (tile / rp.tiles_x) * rp.strip_stride;
        const int px = (tile % rp.tiles_x) * 8 + (t & 7);
        const int py = strip * 8 + (t >> 3);
        live = tile >= 0 && px < rp.chunk_w && py < rp.chunk_h;
        pixel = live ? py * rp.chunk_w + px : -1;
        const D3 start = Load3(rp.sensor), d_scan = Load3(rp.sensor + 3), d_pixel = Load3(rp.sensor + 6);
        o = Load3(rp.origin);
        d = Normalized(Add(Add(start, MulS(d_scan, (double)(rp.chunk_y + py))), MulS(d_pixel, (double)(rp.chunk_x + px))));
        if (live) Count<DBG>(cnt, kPrimary);
      } else {
        o = Load3(wf.rq_o[q] + (size_t)i * 3);
        d = Load3(wf.rq_d[q] + (size_t)i * 3);
        pixel = wf.rq_pixel[q][i];
        path = wf.rq_path[q][i];
        coef = wf.rq_coef[q][i];
        in_object = wf.rq_inobj[q][i] != 0;
      }
      wf.act_refl[act] = -1;
      wf.act_refr[act] = -1;
      wf.act_pixel[act] = pixel;
      wf.act_path[act] = path;
      int act_mtl = -2;
      D3 color = Mk(0.0, 0.0, 0.0);
      if (live) {
        double t = 0.0;
        const int slot = Trace<DBG>(sc, o, d, CUDART_INF, &t, cnt, fctx);
        traced++;
        if (rp.n_rays != nullptr) atomicAdd(rp.n_rays + pixel, 1u);
        WfChargeTile(rp, pixel, 1u);
        if (slot < 0) {
          if (level == 0 && rp.dbg != nullptr) {
            mtb_debug *dbg = rp.dbg + pixel;
            dbg->line_no = -1;
            dbg->pad_ = 0;
            dbg->point[0] = dbg->point[1] = dbg->point[2] = CUDART_NAN;
          }
        } else {
          const ShadeRec *sh = sc.shade + slot;
          const SlotRec *sr = sc.slots + slot;
          P = Add(o, MulS(d, t));
          const int line_no = __ldg(&sh->line_no);
          if (level == 0 && rp.dbg != nullptr) {
            mtb_debug *dbg = rp.dbg + pixel;
            dbg->line_no = line_no;
            dbg->pad_ = 0;
            dbg->point[0] = P.x;
            dbg->point[1] = P.y;
            dbg->point[2] = P.z;
          }
          if (rp.sig_hits != nullptr) {
            atomicAdd(reinterpret_cast<unsigned long long *>(rp.sig_hits) + pixel, Mix64(path, 1ull, (unsigned long long)(long long)line_no));
          }
          Count<DBG>(cnt, kShade);
          const D3 v0 = Load3(sr->vert), v1 = Load3(sr->vert + 3), v2 = Load3(sr->vert + 6);
          const BaryWeights w = Barycentric(v0, v1, v2, P);
          D3 normal = DivS(Add(Add(MulS(Load3(sh->normal), w.n0), MulS(Load3(sh->normal + 3), w.n1)), MulS(Load3(sh->normal + 6), w.n2)), w.n);
          const D3 towards_camera = Neg(d);
          double normal_ray_dot = Dot(towards_camera, normal);
          if (normal_ray_dot < 0.0) {
            normal = Neg(normal);
            normal_ray_dot = Dot(towards_camera, normal);
          }
          const int material = __ldg(&sh->material);
          if (material < 0) {  // mythtracer.cc:49-52
            normal_ray_dot = (normal_ray_dot + 1.0) * 0.5;
            color = Mk(normal_ray_dot, normal_ray_dot, normal_ray_dot);
          } else {
            const mtb_material *m = sc.materials + material;
            D3 surface = Load3(m->ambient);
            const int tex = m->texture;
            if (tex >= 0) {
              const double u = (sh->uv[0] * w.n0 + sh->uv[2] * w.n1 + sh->uv[4] * w.n2) / w.n;
              const double v = (sh->uv[1] * w.n0 + sh->uv[3] * w.n1 + sh->uv[5] * w.n2) / w.n;
              surface = MulV(surface, SampleTexture(sc.tex_atlas, tex, sc.texture_dim[tex], u, v));
            }
            child_refl = Sub(d, MulS(normal, 2 * Dot(normal, d)));
            act_mtl = material;
            Store3(wf.act_point + (size_t)act * 3, P);
            Store3(wf.act_normal + (size_t)act * 3, normal);
            Store3(wf.act_surface + (size_t)act * 3, surface);
            Store3(wf.act_reflected + (size_t)act * 3, child_refl);
            Store3(wf.act_dir + (size_t)act * 3, d);
            if (level < rp.max_depth) {
              refl = m->reflectance;
              do_reflect = refl > 0.0 && coef > 0.01 && !in_object;  // mythtracer.cc:181-184
              do_refract = m->transparency > 0.0;                    // mythtracer.cc:192
            }
          }
        }
      }
      wf.act_mtl[act] = act_mtl;
      Store3(wf.act_color + (size_t)act * 3, color);
    }
    // ---- children of the warp's hits -> queue of the next level: warp ballots + prefix counts, one atomic per
    // warp.  The warp's reflection children are stored first, then its refraction children, so that neighbouring
    // queue entries (= the lanes of a warp in the next level) are rays of the same kind from neighbouring pixels ----
    __syncwarp();
    const unsigned refl_mask = __ballot_sync(0xffffffffu, do_reflect);
    const unsigned refr_mask = __ballot_sync(0xffffffffu, do_refract);
    const unsigned n_refl = (unsigned)__popc(refl_mask), n_refr = (unsigned)__popc(refr_mask);
    const unsigned total = n_refl + n_refr;
    if (total == 0u) continue;
    const unsigned below = (1u << span.lane) - 1u;
    unsigned base = 0;
    if (span.lane == 0) base = atomicAdd(wf.level_n + level + 1, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (do_reflect || do_refract) {
      const int next_base = act_base + lr.n;
      if (base + total > (unsigned)wf.queue_cap || (long long)next_base + base + total > (long long)wf.act_cap) {
        wf.ctrl[0] = 1u;  // overflow: this frame is rendered by the repair launch, the next one gets larger queues
      } else {
        if (do_reflect) {
          const unsigned pos = base + (unsigned)__popc(refl_mask & below);
          Count<DBG>(cnt, kReflect);
          Store3(wf.rq_o[qn] + (size_t)pos * 3, Add(P, MulS(child_refl, 0.0001)));  // mythtracer.cc:70-75
          Store3(wf.rq_d[qn] + (size_t)pos * 3, child_refl);
          wf.rq_coef[qn][pos] = coef * refl;
          wf.rq_path[qn][pos] = path * 2ull;
          wf.rq_pixel[qn][pos] = pixel;
          wf.rq_inobj[qn][pos] = in_object ? 1 : 0;
          wf.act_refl[act] = next_base + (int)pos;
        }
        if (do_refract) {
          const unsigned pos = base + n_refl + (unsigned)__popc(refr_mask & below);
          Count<DBG>(cnt, kRefract);
          const D3 rdir = Normalized(d);                                       // mythtracer.cc:208-212
          Store3(wf.rq_o[qn] + (size_t)pos * 3, Add(P, MulS(rdir, 0.00001)));  // mythtracer.cc:214-218
          Store3(wf.rq_d[qn] + (size_t)pos * 3, rdir);
          wf.rq_coef[qn][pos] = coef;
          wf.rq_path[qn][pos] = path * 2ull + 1ull;
          wf.rq_pixel[qn][pos] = pixel;
          wf.rq_inobj[qn][pos] = in_object ? 0 : 1;
          wf.act_refr[act] = next_base + (int)pos;
        }
      }
    }
  }
  FlushCounters<DBG>(cnt, wf.work, traced);
}

// ---------------------------------------------------------------------------------------------------
// WfShadow: thread = (light, activation).  Light-major task order keeps the rays of a warp aimed at one light.
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__global__ void __launch_bounds__(kWfBlock, MTB_WF_MIN_BLOCKS) WfShadow(DeviceScene sc, RenderParams rp, WfBuffers wf, int level, int lanes) {
  MTB_DECLARE_FAST_CTX(kWfBlock);
  if (wf.ctrl[0] != 0u) return;
  const LevelRange lr = WfLevel(wf, level);
  const int n = lr.n < wf.queue_cap ? lr.n : wf.queue_cap;
  const long long tasks = (long long)n * sc.n_lights;
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  unsigned traced = 0;
  const WarpSpan span = WfSpan(lanes);
  for (long long first = span.first; first < tasks; first += span.step) {
    const long long task = first + span.lane;
    if (span.lane >= lanes || task >= tasks) continue;
    const int li = (int)(task / n);
    const int act = lr.base + (int)(task - (long long)li * n);
    if (wf.act_mtl[act] < 0) continue;
    const D3 P = Load3(wf.act_point + (size_t)act * 3);
    const D3 lpos = Load3(sc.lights[li].position);
    const D3 ldir = Normalized(Sub(lpos, P));
    D3 power = Mk(1.0, 1.0, 1.0);
    bool in_shadow = false, through = false;
    unsigned segments = 0;
    D3 seg_start = P;
    for (;;) {  // mythtracer.cc:94-156
      const D3 to = Add(seg_start, MulS(ldir, 0.00001));
      const double light_distance = Dist(seg_start, lpos);
      double t = 0.0;
      Count<DBG>(cnt, kShadow);
      const int slot = Trace<DBG>(sc, to, ldir, light_distance, &t, cnt, fctx);
      segments++;
      if (slot < 0) break;
      if (t > light_distance) break;
      const int smtl = __ldg(&sc.shade[slot].material);
      const double str = smtl >= 0 ? __ldg(&sc.materials[smtl].transparency) : 0.0;
      if (str == 0.0) {
        power = Mk(0.0, 0.0, 0.0);
        in_shadow = true;
        break;
      }
      if (!through) power = MulV(power, MulS(Load3(sc.materials[smtl].transmission_filter), str));
      through = !through;
      seg_start = Add(Add(to, MulS(ldir, t)), MulS(ldir, 0.0000001));
      if (SqrDist(P, seg_start) > SqrDist(P, lpos)) break;
      if (power.x <= 0.001 && power.y <= 0.001 && power.z <= 0.001) {
        power = Mk(0.0, 0.0, 0.0);
        in_shadow = true;
        break;
      }
    }
    traced += segments;
    Store3(wf.sh_power + ((size_t)li * wf.act_cap + act) * 3, power);
    wf.sh_flags[(size_t)li * wf.act_cap + act] = (in_shadow ? 1u : 0u) | (segments << 1);
    if (rp.n_rays != nullptr) atomicAdd(rp.n_rays + wf.act_pixel[act], segments);
    WfChargeTile(rp, wf.act_pixel[act], segments);
  }
  FlushCounters<DBG>(cnt, wf.work, traced);
}

// ---------------------------------------------------------------------------------------------------
// WfLight: Phong sum over the lights in scene order (mythtracer.cc:78-178)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWfBlock) WfLight(DeviceScene sc, RenderParams rp, WfBuffers wf, int level) {
  if (wf.ctrl[0] != 0u) return;
  const LevelRange lr = WfLevel(wf, level);
  const int n = lr.n < wf.queue_cap ? lr.n : wf.queue_cap;
  for (int k = (int)(blockIdx.x * blockDim.x + threadIdx.x); k < n; k += (int)(gridDim.x * blockDim.x)) {
    const int act = lr.base + k;
    const int material = wf.act_mtl[act];
    if (material < 0) continue;
    const mtb_material *m = sc.materials + material;
    const D3 P = Load3(wf.act_point + (size_t)act * 3);
    const D3 normal = Load3(wf.act_normal + (size_t)act * 3);
    const D3 surface = Load3(wf.act_surface + (size_t)act * 3);
    const D3 reflected = Load3(wf.act_reflected + (size_t)act * 3);
    const D3 m_d = Load3(wf.act_dir + (size_t)act * 3);
    const unsigned long long path = wf.act_path[act];
    D3 color = Mk(0.0, 0.0, 0.0);
    unsigned long long sig = 0;
    for (int li = 0; li < sc.n_lights; li++) {
      const mtb_light *lt = sc.lights + li;
      const D3 ldir = Normalized(Sub(Load3(lt->position), P));
      const D3 lamb = Load3(lt->ambient);
      color = Add(color, MulV(lamb, surface));
      D3 power = Load3(wf.sh_power + ((size_t)li * wf.act_cap + act) * 3);
      const unsigned flags = wf.sh_flags[(size_t)li * wf.act_cap + act];
      const bool in_shadow = (flags & 1u) != 0u;
      sig += Mix64(path, 2ull + (unsigned long long)li, (unsigned long long)flags);
      power.x = SMax(power.x, lamb.x);
      power.y = SMax(power.y, lamb.y);
      power.z = SMax(power.z, lamb.z);
      color = Add(color, MulV(MulV(MulS(MulV(Load3(m->diffuse), surface), Dot(normal, ldir)), Load3(lt->diffuse)), power));
      if (!in_shadow) {
        const double refl_dot = Dot(Neg(m_d), reflected);
        if (refl_dot > 0) {
          color = Add(color, MulV(MulS(MulV(Load3(m->specular), surface), pow(refl_dot, m->specular_exp)), Load3(lt->specular)));
        }
      }
    }
    if (rp.sig_shadow != nullptr && sc.n_lights > 0) {
      atomicAdd(reinterpret_cast<unsigned long long *>(rp.sig_shadow) + wf.act_pixel[act], sig);
    }
    Store3(wf.act_color + (size_t)act * 3, color);
  }
}

// parent += child * Refl ; parent += (child * Tf) * Tr   (mythtracer.cc:185-189, 220-224)
__global__ void __launch_bounds__(256) WfFold(DeviceScene sc, WfBuffers wf, int level) {
  if (wf.ctrl[0] != 0u) return;
  const LevelRange lr = WfLevel(wf, level);
  const int n = lr.n < wf.queue_cap ? lr.n : wf.queue_cap;
  for (int k = (int)(blockIdx.x * blockDim.x + threadIdx.x); k < n; k += (int)(gridDim.x * blockDim.x)) {
    const int a = lr.base + k;
    const int rc = wf.act_refl[a], tc = wf.act_refr[a];
    if (rc < 0 && tc < 0) continue;
    const mtb_material *m = sc.materials + wf.act_mtl[a];
    D3 color = Load3(wf.act_color + (size_t)a * 3);
    if (rc >= 0) color = Add(color, MulS(Load3(wf.act_color + (size_t)rc * 3), m->reflectance));
    if (tc >= 0) color = Add(color, MulS(MulV(Load3(wf.act_color + (size_t)tc * 3), Load3(m->transmission_filter)), m->transparency));
    Store3(wf.act_color + (size_t)a * 3, color);
  }
}

// One thread = 8 bytes of a tile row (a row of 8 pixels = 24 bytes = the 24 consecutive doubles of act_color that
// belong to slots tile * 64 + row * 8 + 0..7): the frame may be the peer-mapped frame of another GPU, where one
// aligned 8-byte store per thread is what the link likes (RenderMega writes its tiles the same way).
__global__ void __launch_bounds__(256) WfResolve(RenderParams rp, WfBuffers wf, int n_slots) {
  if (wf.ctrl[0] != 0u) return;
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);  // (tile position, row, segment)
  const int pos = i / 24, row = (i % 24) / 3, seg = i % 3;
  if (pos * 64 >= n_slots) return;
  // same slot -> pixel mapping as level 0 of WfTrace (the level-0 queue itself has been reused by now)
  const int tile = WfTileOfSlot(rp, pos * 64);
  if (tile < 0) return;
  const int strip = rp.strip_first + (tile / rp.tiles_x) * rp.strip_stride;
  const int px0 = (tile % rp.tiles_x) * 8, py = strip * 8 + row;
  if (py >= rp.chunk_h) return;
  const double *c = wf.act_color + ((size_t)pos * 64 + (size_t)row * 8) * 3 + seg * 8;
  unsigned char b[8];
#pragma unroll
  for (int k = 0; k < 8; k++) b[k] = QuantizeChannel(c[k]);
  unsigned char *out = rp.rgb + ((size_t)py * rp.chunk_w + px0) * 3 + seg * 8;
  if (px0 + 8 <= rp.chunk_w && (rp.chunk_w & 7) == 0 && (reinterpret_cast<uintptr_t>(rp.rgb) & 7u) == 0u) {
    uint2 v;
    v.x = (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24);
    v.y = (unsigned)b[4] | ((unsigned)b[5] << 8) | ((unsigned)b[6] << 16) | ((unsigned)b[7] << 24);
    *reinterpret_cast<uint2 *>(out) = v;
  } else {
    const int valid = (rp.chunk_w - px0) * 3 - seg * 8;  // bytes of this segment that lie inside the chunk
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < valid) out[k] = b[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Queue pipeline: ONE persistent kernel for all levels
// ---------------------------------------------------------------------------------------------------
// The level-by-level pipeline above pays, per level, one kernel boundary on its critical path (a level's trace kernel
// ends with its slowest ray before the first ray of the next level starts) and writes the whole shading context of
// every activation to HBM for the shadow / light kernels.  Here a task is ONE ACTIVATION from beginning to end:
// a warp takes up to 32 consecutive entries of a single ray queue (entry id = activation id; ids below `slots` are
// the primary rays), traces them, shades the hits, appends their reflection / refraction children to the queue
// (warp ballots + prefix counts, one atomic per warp) and publishes them BEFORE it walks its own shadow segments -
// so the chain primary -> child -> grandchild runs ahead on other warps while the shadow walks fill the machine -
// and then does the shadow walks and the Phong sum of its hits in registers (nothing but colour, material and the two
// child links of an activation is ever stored).  The lanes of a warp are in phase by construction: all trace, then
// all walk towards light 0, then light 1, ... (the megakernel's lanes drift apart because every pixel follows its own
// tree).  The fold happens afterwards, per pixel, in the reference's order (WfResolveTree).
//
// Queue protocol.  qctl[kQTail] is advanced by producers (reservation), then the entries are written, then - after a
// __threadfence() - act_ready[id] = epoch marks each one complete.  A consumer takes [head, head + n) with one
// compare-and-swap on qctl[kQHead] (never beyond the reserved tail), and a lane whose entry is not complete yet spins
// on its flag; the producer never waits for anything in between, so the wait is short.  Queue fields are read with
// ld.global.cg: they were written by another SM during this kernel and L1 is not coherent.  qctl[kQPending] counts
// activations that are not finished; a warp that finds the queue empty leaves when it is zero.  Overflow of the
// activation table, or a wait that lasts absurdly long, sets ctrl[0]: every warp leaves and the repair launch
// (RenderMega, same bytes) renders the frame.
#ifndef MTB_QUEUE_MIN_BLOCKS
#define MTB_QUEUE_MIN_BLOCKS 16  // 64-thread blocks at 64 registers, the megakernel's shape (14 / 16 / 18 / 20 blocks: C3 9.77 / 9.36 / 9.38 / 9.42 ms)
#endif
constexpr unsigned kQueueSpinLimit = 1u << 22;
#ifndef MTB_QUEUE_MIN_TICKET
#define MTB_QUEUE_MIN_TICKET 4
#endif
#ifndef MTB_QUEUE_DRAIN
#define MTB_QUEUE_DRAIN 1  // 0: tickets are always 32 wide
#endif

__device__ __forceinline__ unsigned LdVolatile(const uint32_t *p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ D3 LdCg3(const double *p) { return Mk(__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)); }

// What an activation keeps across a traversal (local memory, explicit; cf. PixelState in megakernel.cu).
struct alignas(16) QueueState {
  double m_o[3], m_d[3];                                       // its ray
  double P[3], normal[3], surface[3], reflected[3], color[3];  // shading context (mythtracer.cc:38-76)
  double ldir[3], seg_start[3], power[3];                      // shadow walk of the current light (mythtracer.cc:86-156)
  double coef, light_distance;
  unsigned long long path, sig;
  int level, pixel, material, child_refl, child_refr, in_object;
  unsigned segments, rays;
};
#define MTB_QSTATE_BARRIER() asm volatile("" : : "l"(&st) : "memory")
__device__ __forceinline__ D3 Ld3q(const double *p) { return Mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void St3q(double *p, const D3 &v) {
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}

__global__ void WfQueueBegin(WfBuffers wf, int slots) {
  const int i = (int)threadIdx.x;
  if (i <= MTB_MAX_RAY_DEPTH + 1) wf.level_n[i] = i == 0 ? (uint32_t)slots : 0u;
  if (i < 2) wf.ctrl[i] = 0u;
  if (i < kNumCounters) wf.work[i] = 0ull;
  if (i < kQWords) wf.qctl[i] = (i == kQTail || i == kQPending) ? (uint32_t)slots : 0u;
}

template <bool DBG>
__global__ void __launch_bounds__(kBlockThreads, MTB_QUEUE_MIN_BLOCKS) WfQueue(DeviceScene sc, RenderParams rp, WfBuffers wf, int slots, unsigned epoch) {
  MTB_DECLARE_FAST_CTX(kBlockThreads);
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  unsigned traced = 0;
  // The warp's ticket: it owns activations [ticket, ticket + tsize) and has handed `done` of them to its lanes so far.
  // Tickets are 32 wide (primary rays: whole groups of 8x4 pixels) until the queue runs dry at the very end of the
  // frame: when fewer rays are queued than MTB_QUEUE_DRAIN per warp of the grid, a warp takes 16, 8 or 4, so that the
  // last rays spread over more, emptier warps - a warp runs as long as its slowest ray, and the kernel as long as its
  // last warp.  Measured on C3, share 1/8 of the frame: always 32 wide 1.67 ms; this rule 1.50; thresholds 4x higher
  // 1.67, 6x higher 1.86 (small tickets any earlier cost more than they gain); tickets down to 2 or 1: 1.49.
  const unsigned n_warps = gridDim.x * (kBlockThreads / 32);
  unsigned ticket = 0, tsize = 32, done = 32;
  for (;;) {
    if (done >= tsize) {
      if (lane == 0u) {
        const unsigned head = LdVolatile(wf.qctl + kQHead);
        tsize = 32u;
        if (head >= (unsigned)slots) {  // (the head only grows: a ticket taken now lies behind the primary rays too)
          const unsigned tail = LdVolatile(wf.qctl + kQTail);
          const unsigned backlog = tail > head ? tail - head : 0u;
          const unsigned unit = n_warps * MTB_QUEUE_DRAIN;
          tsize = backlog >= 4u * unit ? 32u : (backlog >= 2u * unit ? 16u : (backlog >= unit ? 8u : 4u));
          if (MTB_QUEUE_MIN_TICKET < 4 && backlog < unit / 2u) tsize = backlog >= unit / 4u ? 2u : (unsigned)MTB_QUEUE_MIN_TICKET;
        }
        ticket = atomicAdd(wf.qctl + kQHead, tsize);  // fetch-and-add: never fails, never retried
      }
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      tsize = __shfl_sync(0xffffffffu, tsize, 0);
      done = 0;
    }
    if (ticket >= (unsigned)wf.act_cap) break;  // beyond the table: such entries are never produced (overflow is flagged)
    // ---- how many of the ticket's entries exist by now?  Primary rays always do; queued ones once the tail has
    // passed them.  A partly filled ticket is not waited for for long: what is there goes out to the first lanes
    // (late in a frame that spreads the few remaining rays over many warps by itself) ----
    unsigned n = tsize - done;
    bool over = false;
    if (ticket >= (unsigned)slots) {
      unsigned waited = 0;
      for (;;) {
        unsigned tail = 0, pending = 1, stop = 0;
        if (lane == 0u) {
          tail = LdVolatile(wf.qctl + kQTail);
          pending = LdVolatile(wf.qctl + kQPending);
          stop = LdVolatile(wf.ctrl);
          if (pending == 0u) tail = LdVolatile(wf.qctl + kQTail);  // nothing is in flight any more: this tail is final
        }
        tail = __shfl_sync(0xffffffffu, tail, 0);
        pending = __shfl_sync(0xffffffffu, pending, 0);
        stop = __shfl_sync(0xffffffffu, stop, 0);
        const unsigned first = ticket + done;
        const unsigned avail = tail > first ? (tail - first < tsize - done ? tail - first : tsize - done) : 0u;
        if (stop != 0u || (avail == 0u && pending == 0u)) {
          over = true;
          break;
        }
        if (avail == tsize - done || (avail > 0u && waited >= 2u)) {
          n = avail;
          break;
        }
        __nanosleep(waited < 16u ? 250u : 2000u);
        if (++waited > kQueueSpinLimit) {
          if (lane == 0u) wf.ctrl[0] = 1u;
          over = true;
          break;
        }
      }
    }
    if (over) break;

    // ---- the batch: lane k < n owns activation ticket + done + k ----
    const bool mine = lane < n;
    const int act = (int)(ticket + done + lane);
    done += n;
    bool live = mine, failed = false;
    // what has to survive a traversal lives in `st` (local memory, explicit - see PixelState in megakernel.cu for why),
    // including the scalars that are only touched between two traversals: a register each across the node loop is
    // what pushes other values into spill slots at 64 registers
    QueueState st;
    st.in_object = 0;
    st.level = 0;
    st.pixel = -1;
    st.coef = 1.0;
    st.path = 1ull;
    D3 o = Mk(0.0, 0.0, 0.0), d = Mk(0.0, 0.0, 0.0);
    if (mine) {
      if (act < slots) {
        // 8x8 pixel tiles (a warp = 8x4 pixels) of this launch's strips, as in RenderMega / WfTrace level 0
        const int tile = WfTileOfSlot(rp, act), t = act & 63;
        const int strip = rp.strip_first + (tile / rp.tiles_x) * rp.strip_stride;
        const int px = (tile % rp.tiles_x) * 8 + (t & 7);
        const int py = strip * 8 + (t >> 3);
        live = tile >= 0 && px < rp.chunk_w && py < rp.chunk_h;
        st.pixel = live ? py * rp.chunk_w + px : -1;
        const D3 start = Load3(rp.sensor), d_scan = Load3(rp.sensor + 3), d_pixel = Load3(rp.sensor + 6);
        o = Load3(rp.origin);
        d = Normalized(Add(Add(start, MulS(d_scan, (double)(rp.chunk_y + py))), MulS(d_pixel, (double)(rp.chunk_x + px))));
        if (live) Count<DBG>(cnt, kPrimary);
      } else {
        // reserved, and being written by the warp that spawned it: a short wait
        unsigned spins = 0;
        while (LdVolatile(wf.act_ready + act) != epoch) {
          if (++spins > kQueueSpinLimit || ((spins & 255u) == 0u && LdVolatile(wf.ctrl) != 0u)) {
            failed = true;
            break;
          }
          __nanosleep(40u);
        }
        __threadfence();
        if (!failed) {
          o = LdCg3(wf.act_point + (size_t)act * 3);
          d = LdCg3(wf.act_dir + (size_t)act * 3);
          st.coef = __ldcg(wf.act_coef + act);
          st.path = __ldcg(wf.act_path + act);
          st.pixel = __ldcg(wf.act_pixel + act);
          const int info = __ldcg(wf.act_info + act);
          st.level = info & 0xff;
          st.in_object = (info >> 8) != 0;
        }
      }
    }
    if (__any_sync(0xffffffffu, failed)) {
      if (lane == 0u) wf.ctrl[0] = 1u;
      break;
    }

    // ---- one traversal per loop iteration, as in RenderMega: iteration 0 traces the activations' own rays, then the
    // hits are shaded and their children queued (warp-wide), then every further iteration is one shadow segment of
    // the lanes that still walk.  What has to survive a traversal lives in `st` (local memory, explicit - see
    // PixelState in megakernel.cu for why); a handful of scalars stay in registers ----
    St3q(st.m_o, o);
    St3q(st.m_d, d);
    St3q(st.color, Mk(0.0, 0.0, 0.0));
    bool walking = live, main_phase = true, lit = false, in_shadow = false, through = false;
    int li = 0;
    st.material = -2;
    st.child_refl = -1;
    st.child_refr = -1;
    st.segments = 0;
    st.rays = 0;
    st.sig = 0ull;
    bool stop_all = false;
    for (;;) {
      int slot = -1;
      double t = 0.0;
      if (walking) {
        D3 to, td;
        double limit = CUDART_INF;
        if (main_phase) {
          to = Ld3q(st.m_o);
          td = Ld3q(st.m_d);
        } else {
          const D3 seg = Ld3q(st.seg_start);
          td = Ld3q(st.ldir);
          to = Add(seg, MulS(td, 0.00001));                                     // mythtracer.cc:95-99
          limit = Dist(seg, Load3(sc.lights[li].position));   // mythtracer.cc:101-102
          st.light_distance = limit;
          Count<DBG>(cnt, kShadow);
        }
        MTB_QSTATE_BARRIER();
        slot = Trace<DBG>(sc, to, td, limit, &t, cnt, fctx);
        MTB_QSTATE_BARRIER();
        st.rays++;
      }
      if (main_phase) {
        // ---- results of the activations' own rays (mythtracer.cc:13-76); every lane of the warp is here ----
        main_phase = false;
        bool do_reflect = false, do_refract = false;
        double refl = 0.0;
        if (walking) {
          walking = false;
          if (slot < 0) {
            if (st.level == 0 && rp.dbg != nullptr) {
              mtb_debug *dbg = rp.dbg + st.pixel;
              dbg->line_no = -1;
              dbg->pad_ = 0;
              dbg->point[0] = dbg->point[1] = dbg->point[2] = CUDART_NAN;
            }
          } else {
            const ShadeRec *sh = sc.shade + slot;
            const SlotRec *sr = sc.slots + slot;
            const D3 m_d = Ld3q(st.m_d);
            const D3 P = Add(Ld3q(st.m_o), MulS(m_d, t));
            const int line_no = __ldg(&sh->line_no);
            if (st.level == 0 && rp.dbg != nullptr) {
              mtb_debug *dbg = rp.dbg + st.pixel;
              dbg->line_no = line_no;
              dbg->pad_ = 0;
              dbg->point[0] = P.x;
              dbg->point[1] = P.y;
              dbg->point[2] = P.z;
            }
            if (rp.sig_hits != nullptr) {
              atomicAdd(reinterpret_cast<unsigned long long *>(rp.sig_hits) + st.pixel, Mix64(st.path, 1ull, (unsigned long long)(long long)line_no));
            }
            Count<DBG>(cnt, kShade);
            const D3 v0 = Load3(sr->vert), v1 = Load3(sr->vert + 3), v2 = Load3(sr->vert + 6);
            const BaryWeights w = Barycentric(v0, v1, v2, P);
            D3 normal = DivS(Add(Add(MulS(Load3(sh->normal), w.n0), MulS(Load3(sh->normal + 3), w.n1)), MulS(Load3(sh->normal + 6), w.n2)), w.n);
            const D3 towards_camera = Neg(m_d);
            double normal_ray_dot = Dot(towards_camera, normal);
            if (normal_ray_dot < 0.0) {
              normal = Neg(normal);
              normal_ray_dot = Dot(towards_camera, normal);
            }
            const int mtl = __ldg(&sh->material);
            if (mtl < 0) {  // mythtracer.cc:49-52
              normal_ray_dot = (normal_ray_dot + 1.0) * 0.5;
              St3q(st.color, Mk(normal_ray_dot, normal_ray_dot, normal_ray_dot));
            } else {
              const mtb_material *m = sc.materials + mtl;
              D3 surface = Load3(m->ambient);
              const int tex = m->texture;
              if (tex >= 0) {
                const double u = (sh->uv[0] * w.n0 + sh->uv[2] * w.n1 + sh->uv[4] * w.n2) / w.n;
                const double v = (sh->uv[1] * w.n0 + sh->uv[3] * w.n1 + sh->uv[5] * w.n2) / w.n;
                surface = MulV(surface, SampleTexture(sc.tex_atlas, tex, sc.texture_dim[tex], u, v));
              }
              const D3 reflected = Sub(m_d, MulS(normal, 2 * Dot(normal, m_d)));
              St3q(st.P, P);
              St3q(st.normal, normal);
              St3q(st.surface, surface);
              St3q(st.reflected, reflected);
              st.material = mtl;
              lit = true;
              if (st.level < rp.max_depth) {
                refl = m->reflectance;
                do_reflect = refl > 0.0 && st.coef > 0.01 && !st.in_object;  // mythtracer.cc:181-184
                do_refract = m->transparency > 0.0;                    // mythtracer.cc:192
              }
            }
          }
        }
        // ---- children -> queue, published before this warp walks its shadow segments.  The warp's reflection
        // children are stored first, then its refraction children: neighbouring entries (= the lanes of some warp
        // later) are rays of the same kind from neighbouring pixels ----
        const unsigned refl_mask = __ballot_sync(0xffffffffu, do_reflect);
        const unsigned refr_mask = __ballot_sync(0xffffffffu, do_refract);
        const unsigned n_refl = (unsigned)__popc(refl_mask), total = n_refl + (unsigned)__popc(refr_mask);
        if (total != 0u) {
          unsigned base = 0;
          if (lane == 0u) {
            atomicAdd(wf.qctl + kQPending, total);  // counted before they can be seen: the count never runs low
            base = atomicAdd(wf.qctl + kQTail, total);
          }
          base = __shfl_sync(0xffffffffu, base, 0);
          if ((unsigned long long)base + total > (unsigned long long)wf.act_cap) {
            stop_all = true;
          } else {
            if (do_reflect || do_refract) {
              const D3 P = Ld3q(st.P), m_d = Ld3q(st.m_d);
              if (do_reflect) {
                const int c = (int)(base + (unsigned)__popc(refl_mask & below));
                const D3 reflected = Ld3q(st.reflected);
                Count<DBG>(cnt, kReflect);
                Store3(wf.act_point + (size_t)c * 3, Add(P, MulS(reflected, 0.0001)));  // mythtracer.cc:70-75
                Store3(wf.act_dir + (size_t)c * 3, reflected);
                wf.act_coef[c] = st.coef * refl;
                wf.act_path[c] = st.path * 2ull;
                wf.act_pixel[c] = st.pixel;
                wf.act_info[c] = (st.level + 1) | (st.in_object ? 256 : 0);
                st.child_refl = c;
              }
              if (do_refract) {
                const int c = (int)(base + n_refl + (unsigned)__popc(refr_mask & below));
                Count<DBG>(cnt, kRefract);
                const D3 rdir = Normalized(m_d);                                      // mythtracer.cc:208-212
                Store3(wf.act_point + (size_t)c * 3, Add(P, MulS(rdir, 0.00001)));    // mythtracer.cc:214-218
                Store3(wf.act_dir + (size_t)c * 3, rdir);
                wf.act_coef[c] = st.coef;
                wf.act_path[c] = st.path * 2ull + 1ull;
                wf.act_pixel[c] = st.pixel;
                wf.act_info[c] = (st.level + 1) | (st.in_object ? 0 : 256);
                st.child_refr = c;
              }
              __threadfence();
              if (st.child_refl >= 0) asm volatile("st.volatile.global.u32 [%0], %1;" : : "l"(wf.act_ready + st.child_refl), "r"(epoch) : "memory");
              if (st.child_refr >= 0) asm volatile("st.volatile.global.u32 [%0], %1;" : : "l"(wf.act_ready + st.child_refr), "r"(epoch) : "memory");
            }
          }
        }
        if (stop_all) break;
        // this batch's activations spawn nothing else: they leave the count (their children are in it already)
        if (lane == 0u) atomicAdd(wf.qctl + kQPending, 0u - n);
        // the links and the material are final now
        if (mine) {
          wf.act_mtl[act] = st.material;
          wf.act_refl[act] = st.child_refl;
          wf.act_refr[act] = st.child_refr;
        }
        walking = lit && sc.n_lights > 0;
      } else if (walking) {
        // ---- one shadow segment came back (mythtracer.cc:104-156) ----
        bool light_done = false;
        st.segments++;
        if (slot < 0) {
          light_done = true;
        } else if (t > st.light_distance) {
          light_done = true;
        } else {
          const int smtl = __ldg(&sc.shade[slot].material);
          const double str = smtl >= 0 ? __ldg(&sc.materials[smtl].transparency) : 0.0;
          if (str == 0.0) {
            St3q(st.power, Mk(0.0, 0.0, 0.0));
            in_shadow = true;
            light_done = true;
          } else {
            D3 power = Ld3q(st.power);
            if (!through) {
              power = MulV(power, MulS(Load3(sc.materials[smtl].transmission_filter), str));
              St3q(st.power, power);
            }
            through = !through;
            const D3 ldir = Ld3q(st.ldir);
            const D3 to = Add(Ld3q(st.seg_start), MulS(ldir, 0.00001));
            const D3 seg_start = Add(Add(to, MulS(ldir, t)), MulS(ldir, 0.0000001));  // mythtracer.cc:137
            St3q(st.seg_start, seg_start);
            const D3 P = Ld3q(st.P);
            if (SqrDist(P, seg_start) > SqrDist(P, Load3(sc.lights[li].position))) {  // mythtracer.cc:141-145
              light_done = true;
            } else if (power.x <= 0.001 && power.y <= 0.001 && power.z <= 0.001) {
              St3q(st.power, Mk(0.0, 0.0, 0.0));
              in_shadow = true;
              light_done = true;
            }
          }
        }
        if (light_done) {
          // ---- this light is settled: Phong terms (mythtracer.cc:159-177) ----
          const mtb_light *lt = sc.lights + li;
          const mtb_material *m = sc.materials + st.material;
          st.sig += Mix64(st.path, 2ull + (unsigned long long)li, (in_shadow ? 1ull : 0ull) | ((unsigned long long)st.segments << 1));
          const D3 lamb = Load3(lt->ambient);
          D3 power = Ld3q(st.power);
          power.x = SMax(power.x, lamb.x);
          power.y = SMax(power.y, lamb.y);
          power.z = SMax(power.z, lamb.z);
          const D3 surface = Ld3q(st.surface);
          D3 color = Ld3q(st.color);
          color = Add(color, MulV(MulV(MulS(MulV(Load3(m->diffuse), surface), Dot(Ld3q(st.normal), Ld3q(st.ldir))), Load3(lt->diffuse)), power));
          if (!in_shadow) {
            const double refl_dot = Dot(Neg(Ld3q(st.m_d)), Ld3q(st.reflected));
            if (refl_dot > 0) {
              color = Add(color, MulV(MulS(MulV(Load3(m->specular), surface), pow(refl_dot, m->specular_exp)), Load3(lt->specular)));
            }
          }
          St3q(st.color, color);
          li++;
          walking = li < sc.n_lights;
        } else {
          continue;  // next segment of the same light
        }
      }
      if (!walking) break;
      // ---- start the shadow walk of light li (mythtracer.cc:79-94) ----
      {
        const mtb_light *lt = sc.lights + li;
        const D3 P = Ld3q(st.P);
        St3q(st.ldir, Normalized(Sub(Load3(lt->position), P)));
        St3q(st.color, Add(Ld3q(st.color), MulV(Load3(lt->ambient), Ld3q(st.surface))));  // mythtracer.cc:83-84
        St3q(st.power, Mk(1.0, 1.0, 1.0));
        St3q(st.seg_start, P);
        in_shadow = false;
        through = false;
        st.segments = 0;
      }
    }
    if (__any_sync(0xffffffffu, stop_all)) {
      if (lane == 0u) wf.ctrl[0] = 1u;
      break;
    }
    if (mine) {
      Store3(wf.act_color + (size_t)act * 3, Ld3q(st.color));
      if (live) {
        traced += st.rays;
        if (lit && rp.sig_shadow != nullptr && sc.n_lights > 0) atomicAdd(reinterpret_cast<unsigned long long *>(rp.sig_shadow) + st.pixel, st.sig);
        if (rp.n_rays != nullptr) atomicAdd(rp.n_rays + st.pixel, st.rays);
        WfChargeTile(rp, st.pixel, st.rays);
      }
    }
  }
  FlushCounters<DBG>(cnt, wf.work, traced);
}

// Per-pixel fold of the activation tree in the reference's order - parent += child * Refl, then
// parent += (child * Tf) * Tr (mythtracer.cc:185-189, 220-224), children before parents - followed by V3DtoRGB
// (mythtracer.cc:235-241).  One block = one 8x8 tile (slots pos * 64 ..), written like RenderMega writes its tiles.
__global__ void __launch_bounds__(kBlockThreads) WfResolveTree(DeviceScene sc, RenderParams rp, WfBuffers wf, int slots) {
  __shared__ __align__(8) unsigned char s_rgb[kTile * kTile * 3];
  if (wf.ctrl[0] != 0u) return;
  const int pos = (int)blockIdx.x, root = pos * 64 + (int)threadIdx.x;
  const int tile = WfTileOfSlot(rp, pos * 64);
  if (tile < 0 || root >= slots) return;
  const int strip = rp.strip_first + (tile / rp.tiles_x) * rp.strip_stride;
  const int tx = (int)(threadIdx.x & 7u), ty = (int)(threadIdx.x >> 3);
  const int px = (tile % rp.tiles_x) * kTile + tx, py = strip * kTile + ty;
  const bool live = px < rp.chunk_w && py < rp.chunk_h;
  {
    int f_act[MTB_MAX_RAY_DEPTH + 2];
    unsigned char f_stage[MTB_MAX_RAY_DEPTH + 2];  // 0: nothing folded yet, 1: reflection child in progress / done, 2: refraction child
    D3 f_color[MTB_MAX_RAY_DEPTH + 2];
    int sp = 0;
    D3 ret = Mk(0.0, 0.0, 0.0);
    if (live) {
      f_act[0] = root;
      f_stage[0] = 0;
      f_color[0] = Load3(wf.act_color + (size_t)root * 3);
      sp = 1;
    }
    while (sp > 0) {
      const int a = f_act[sp - 1];
      int child = -1;
      if (f_stage[sp - 1] == 0) {
        f_stage[sp - 1] = 1;
        child = wf.act_refl[a];
      }
      if (child < 0 && f_stage[sp - 1] == 1) {
        f_stage[sp - 1] = 2;
        child = wf.act_refr[a];
      }
      if (child >= 0 && sp < MTB_MAX_RAY_DEPTH + 2) {
        f_act[sp] = child;
        f_stage[sp] = 0;
        f_color[sp] = Load3(wf.act_color + (size_t)child * 3);
        sp++;
        continue;
      }
      ret = f_color[sp - 1];
      sp--;
      if (sp == 0) break;
      const mtb_material *m = sc.materials + wf.act_mtl[f_act[sp - 1]];
      if (f_stage[sp - 1] == 1) {
        f_color[sp - 1] = Add(f_color[sp - 1], MulS(ret, m->reflectance));
      } else {
        f_color[sp - 1] = Add(f_color[sp - 1], MulS(MulV(ret, Load3(m->transmission_filter)), m->transparency));
      }
    }
    unsigned char *out = s_rgb + ((ty * kTile) + tx) * 3;
    out[0] = QuantizeChannel(ret.x);
    out[1] = QuantizeChannel(ret.y);
    out[2] = QuantizeChannel(ret.z);
  }
  __syncwarp();
  {
    const int px0 = (tile % rp.tiles_x) * kTile, py0 = strip * kTile + (int)(threadIdx.x >> 5) * 4;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned char *half = s_rgb + (threadIdx.x >> 5) * (4 * kTile * 3);
    const bool whole_rows = px0 + kTile <= rp.chunk_w && (rp.chunk_w & 7) == 0 && (reinterpret_cast<uintptr_t>(rp.rgb) & 7u) == 0u;
    if (whole_rows) {
      if (lane < 12u) {
        const int row = (int)lane / 3, seg = (int)lane % 3;
        if (py0 + row < rp.chunk_h) {
          const uint2 v = *reinterpret_cast<const uint2 *>(half + row * 24 + seg * 8);
          *reinterpret_cast<uint2 *>(rp.rgb + ((size_t)(py0 + row) * rp.chunk_w + px0) * 3 + seg * 8) = v;
        }
      }
    } else if (live) {
      unsigned char *dst = rp.rgb + ((size_t)py * rp.chunk_w + px) * 3;
      const unsigned char *src = s_rgb + ((ty * kTile) + tx) * 3;
      dst[0] = src[0];
      dst[1] = src[1];
      dst[2] = src[2];
    }
  }
}

// Last kernel of a frame.  The work counters of a frame that overflowed are dropped (the repair launch counts its
// own rays); host_copy (pinned, nullable) receives level_n[] and the overflow flag for the next frame's grid sizes
// and queue capacities - written by the device, never waited for by the host.
__global__ void WfCommit(WfBuffers wf, unsigned long long *global, uint32_t *host_copy) {
  const int i = (int)threadIdx.x;
  const bool overflow = wf.ctrl[0] != 0u;
  if (!overflow && global != nullptr && i < kNumCounters && wf.work[i] != 0ull) atomicAdd(global + i, wf.work[i]);
  if (host_copy != nullptr) {
    if (i <= MTB_MAX_RAY_DEPTH + 1) host_copy[i] = wf.level_n[i];
    if (i == 0) host_copy[kWfHostOverflow] = overflow ? 1u : 0u;
    __syncwarp();
    __threadfence_system();
    if (i == 0) host_copy[kWfHostSequence] += 1u;  // frames completed so far
  }
}

}  // namespace

// Lanes per warp for a queue of `n` items: full warps once the queue fills the machine (148 SMs x 32
// resident warps), otherwise the largest power of two that still spreads it over all warp slots.
static int PackLanes(long long n) {
  static const int min_lanes = []() {
    const char *env = getenv("MTB_WF_MIN_LANES");  // development knob (A/B of the packing)
    const int v = env != nullptr ? atoi(env) : 4;
    return v >= 1 && v <= 32 ? v : 4;
  }();
  const long long slots = 148LL * 32;
  int lanes = 32;
  while (lanes > min_lanes && n < slots * lanes) lanes >>= 1;
  return lanes;
}

// Blocks for `items` loop items at `lanes` items per warp, at most what `bound` items need; never zero (the real
// item count is read on the device, the loops are grid-strided).
static int BlocksFor(long long items, long long bound, int lanes, int block) {
  if (items > bound) items = bound;
  if (items < 1) items = 1;
  const long long warps = (items + lanes - 1) / lanes;
  long long blocks = (warps * 32 + block - 1) / block;
  if (blocks > 0x3fffffff) blocks = 0x3fffffff;
  return (int)blocks;
}

void LaunchWfBegin(const WfBuffers &wf, int slots, cudaStream_t stream) { WfBegin<<<1, 32, 0, stream>>>(wf, slots); }

// `expect`: how many rays the level is expected to hold (previous frame, with headroom; or the capacity bound)
void LaunchWfTrace(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int level, long long expect, bool debug_build,
                   cudaStream_t stream) {
  const int lanes = level == 0 ? 32 : PackLanes(expect);  // level 0 maps warps to 8x4 pixel tiles
  const int blocks = BlocksFor(expect, wf.queue_cap, lanes, kWfBlock);
  if (debug_build) {
    WfTrace<true><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, lanes);
  } else {
    WfTrace<false><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, lanes);
  }
}

void LaunchWfShadow(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int level, long long expect, bool debug_build,
                    cudaStream_t stream) {
  if (sc.n_lights <= 0) return;
  const long long tasks = expect * sc.n_lights;
  const int lanes = PackLanes(tasks);
  const int blocks = BlocksFor(tasks, (long long)wf.queue_cap * sc.n_lights, lanes, kWfBlock);
  if (debug_build) {
    WfShadow<true><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, lanes);
  } else {
    WfShadow<false><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, lanes);
  }
}

void LaunchWfLight(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int level, long long expect, cudaStream_t stream) {
  WfLight<<<BlocksFor(expect, wf.queue_cap, 32, kWfBlock), kWfBlock, 0, stream>>>(sc, rp, wf, level);
}

void LaunchWfFold(const DeviceScene &sc, const WfBuffers &wf, int level, long long expect, cudaStream_t stream) {
  WfFold<<<BlocksFor(expect, wf.queue_cap, 32, 256), 256, 0, stream>>>(sc, wf, level);
}

void LaunchWfResolve(const RenderParams &rp, const WfBuffers &wf, int n_slots, cudaStream_t stream) {
  if (n_slots <= 0) return;
  const int threads = (n_slots / 64) * 24;
  WfResolve<<<(threads + 255) / 256, 256, 0, stream>>>(rp, wf, n_slots);
}

void LaunchWfQueueBegin(const WfBuffers &wf, int slots, cudaStream_t stream) { WfQueueBegin<<<1, 32, 0, stream>>>(wf, slots); }

void LaunchWfQueue(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int slots, unsigned epoch, int sm_count, bool debug_build,
                   cudaStream_t stream) {
  // persistent: what the device can hold, but not more warps than the frame has groups of 32 primary rays
  long long blocks = (long long)sm_count * MTB_QUEUE_MIN_BLOCKS;
  const long long need = ((long long)slots / 32 + 1) / 2;
  if (blocks > need) blocks = need < 1 ? 1 : need;
  if (debug_build) {
    WfQueue<true><<<(int)blocks, kBlockThreads, 0, stream>>>(sc, rp, wf, slots, epoch);
  } else {
    WfQueue<false><<<(int)blocks, kBlockThreads, 0, stream>>>(sc, rp, wf, slots, epoch);
  }
}

void LaunchWfResolveTree(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int slots, cudaStream_t stream) {
  if (slots <= 0) return;
  WfResolveTree<<<slots / 64, kBlockThreads, 0, stream>>>(sc, rp, wf, slots);
}

void LaunchWfCommit(const WfBuffers &wf, unsigned long long *global, uint32_t *host_copy, cudaStream_t stream) {
  WfCommit<<<1, 32, 0, stream>>>(wf, global, host_copy);
}

}  // namespace mtb
