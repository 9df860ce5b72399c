// Wavefront pipeline: the recursive TraceRayWorker (mythtracer.cc:13-228) as an iterative, level-by-level
// sequence of kernels over ray queues in HBM.
//
//   level L:  WfTraceMain   one thread per queued ray: OctTree::IntersectRay, hit point, interpolated normal,
//                           surface colour, reflected direction  (mythtracer.cc:18-76) -> activation record
//             WfSpawn       one thread per hit: the reflection / refraction children are appended to the queue
//                           of level L+1 with a warp-aggregated (ballot + prefix sum) slot allocation
//                           (mythtracer.cc:181-225).  Whether a child exists does not depend on the lights.
//     stream 2:
//             WfShadow      one thread per (hit, light): the whole shadow walk through transparent surfaces
//                           (mythtracer.cc:86-156); lights are independent of each other, only the order in
//                           which their terms are summed matters, and that order is kept by WfLight
//             WfLight       one thread per hit: Phong sum over the lights in scene order (mythtracer.cc:78-178)
//   The trace -> spawn -> trace chain of the levels is the critical path (each link ends with the slowest ray
//   of its level); the shadow / light kernels of level L only need level L's activation records, so they run
//   on a second stream and fill the machine while the chain advances.
//   finally:  WfFold        deepest level first: parent += child * Refl, then parent += (child * Tf) * Tr --
//                           the same two additions, in the same order, as the recursion performs on return
//             WfResolve     V3DtoRGB (mythtracer.cc:235-241) into the chunk-local RGB24 buffer
//
// Because every activation keeps its own colour and the fold replays the reference's additions in the
// reference's order, the result is bit-identical to the megakernel (and to the reference, up to pow()).
// Compared with the megakernel the traversal kernels need ~half the registers (the shading state lives in
// HBM between kernels), rays of one kind run together, and finished pixels do not idle lanes.
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

// Coherence key of a queued ray: the direction octant and the Morton code of the origin's cell in a
// 32^3 grid over the scene box.  Rays with equal keys start close together and head the same way, so
// they walk the same octree nodes and list-BVH records.
__device__ __forceinline__ unsigned Spread5(unsigned v) {  // abcde -> a00b00c00d00e
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}
__device__ __forceinline__ unsigned RayKey(const WfBuffers &wf, const D3 &o, const D3 &d) {
  const float fx = ((float)o.x - wf.cell_lo[0]) * wf.cell_scale[0];
  const float fy = ((float)o.y - wf.cell_lo[1]) * wf.cell_scale[1];
  const float fz = ((float)o.z - wf.cell_lo[2]) * wf.cell_scale[2];
  const unsigned cx = (unsigned)fminf(fmaxf(fx, 0.0f), 31.0f);  // NaN -> 0
  const unsigned cy = (unsigned)fminf(fmaxf(fy, 0.0f), 31.0f);
  const unsigned cz = (unsigned)fminf(fmaxf(fz, 0.0f), 31.0f);
  const unsigned oct = (d.x < 0.0 ? 1u : 0u) | (d.y < 0.0 ? 2u : 0u) | (d.z < 0.0 ? 4u : 0u);
  return (oct << 15) | Spread5(cx) | (Spread5(cy) << 1) | (Spread5(cz) << 2);
}

// Counting sort of a level's queue by RayKey: histogram, exclusive scan, scatter.
__global__ void WfSortHistogram(const uint32_t *__restrict__ keys, uint32_t *hist, int n) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < n) atomicAdd(hist + keys[i], 1u);
}
__global__ void __launch_bounds__(1024) WfSortScan(uint32_t *hist) {
  // one block: 1024 threads x 256 consecutive bins = 2^18 bins
  __shared__ uint32_t partial[1024];
  constexpr int kPer = (1 << kWfSortBits) / 1024;
  uint32_t *mine = hist + (size_t)threadIdx.x * kPer;
  uint32_t sum = 0;
  for (int k = 0; k < kPer; k++) sum += mine[k];
  partial[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t v = threadIdx.x >= (unsigned)off ? partial[threadIdx.x - off] : 0u;
    __syncthreads();
    partial[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = partial[threadIdx.x] - sum;
  for (int k = 0; k < kPer; k++) {
    const uint32_t c = mine[k];
    mine[k] = run;
    run += c;
  }
}
__global__ void WfSortScatter(const uint32_t *__restrict__ keys, uint32_t *offsets, int32_t *perm, int n) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i < n) perm[atomicAdd(offsets + keys[i], 1u)] = i;
}

// Sub-warp packing.  A warp runs as long as its slowest ray and serialises what its lanes do differently,
// so a level with few rays (deep levels, or a small share of the frame on one of 8 GPUs) finishes sooner
// when its rays are spread over more, emptier warps: only the first `lanes` lanes of each warp get a ray.
__device__ __forceinline__ long long PackedIndex64(long long n, int lanes) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = (int)(threadIdx.x & 31u);
  return lane < lanes ? warp * lanes + lane : n;
}
__device__ __forceinline__ int PackedIndex(int n, int lanes) { return (int)PackedIndex64(n, lanes); }

// ---------------------------------------------------------------------------------------------------
// WfTraceMain
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

template <bool DBG>
__global__ void __launch_bounds__(kWfBlock, MTB_WF_MIN_BLOCKS) WfTraceMain(DeviceScene sc, RenderParams rp, WfBuffers wf, int level, int n,
                                                        int act_base, const int32_t *__restrict__ perm, int lanes) {
#ifdef MTB_SMEM_TOP
  __shared__ NodeRec top_store[kTopNodes];
  const NodeRec *top = top_store;
  const int top_n = sc.n_nodes < kTopNodes ? sc.n_nodes : kTopNodes;
  StageTopNodes(sc, top_store, sc.n_nodes);
#endif
  // j: position in processing order; i: position in the level's queue; activation id = act_base + i.
  // Only the first `lanes` lanes of a warp carry a ray (see PackedIndex).
  const int j = PackedIndex(n, lanes);
  const int i = (perm != nullptr && j < n) ? perm[j] : j;
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  unsigned traced = 0;
  const int q = level & 1;
  if (j < n) {
    D3 o, d;
    int pixel;
    unsigned long long path;
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
      path = 1ull;
      wf.rq_coef[0][i] = 1.0;
      wf.rq_inobj[0][i] = 0;
      if (live) Count<DBG>(cnt, kPrimary);
    } else {
      o = Load3(wf.rq_o[q] + (size_t)i * 3);
      d = Load3(wf.rq_d[q] + (size_t)i * 3);
      pixel = wf.rq_pixel[q][i];
      path = wf.rq_path[q][i];
    }
    const int act = act_base + i;
    wf.act_refl[act] = -1;
    wf.act_refr[act] = -1;
    wf.act_pixel[act] = pixel;
    wf.act_path[act] = path;
    int act_mtl = -2;
    D3 color = Mk(0.0, 0.0, 0.0);
    if (live) {
      double t = 0.0;
      const int slot = Trace<DBG>(sc, o, d, CUDART_INF, &t, cnt MTB_TOP_ARGS);
      traced = 1;
      if (rp.n_rays != nullptr) atomicAdd(rp.n_rays + pixel, 1u);
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
        const D3 P = Add(o, MulS(d, t));
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
          const D3 reflected = Sub(d, MulS(normal, 2 * Dot(normal, d)));
          act_mtl = material;
          Store3(wf.act_point + (size_t)act * 3, P);
          Store3(wf.act_normal + (size_t)act * 3, normal);
          Store3(wf.act_surface + (size_t)act * 3, surface);
          Store3(wf.act_reflected + (size_t)act * 3, reflected);
          Store3(wf.act_dir + (size_t)act * 3, d);
        }
      }
    }
    wf.act_mtl[act] = act_mtl;
    Store3(wf.act_color + (size_t)act * 3, color);
    WfChargeTile(rp, pixel, traced);
  }
  FlushCounters<DBG>(cnt, rp.counters, traced);
}

// ---------------------------------------------------------------------------------------------------
// WfSpawn: children of the level's hits -> queue of the next level (mythtracer.cc:181-225)
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__global__ void __launch_bounds__(kWfBlock) WfSpawn(DeviceScene sc, RenderParams rp, WfBuffers wf, int level, int n, int act_base) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  const int q = level & 1, qn = q ^ 1;
  const int act = act_base + i;
  bool do_reflect = false, do_refract = false;
  double coef = 0.0, refl = 0.0;
  bool in_object = false;
  int material = -1;
  if (i < n) material = wf.act_mtl[act];
  if (material >= 0 && level < rp.max_depth) {
    const mtb_material *m = sc.materials + material;
    coef = wf.rq_coef[q][i];
    in_object = wf.rq_inobj[q][i] != 0;
    refl = m->reflectance;
    do_reflect = refl > 0.0 && coef > 0.01 && !in_object;  // mythtracer.cc:181-184
    do_refract = m->transparency > 0.0;                    // mythtracer.cc:192
  }
  // ---- queue compaction: warp ballots + prefix counts, one atomic per warp.  The warp's reflection
  // children are stored first, then its refraction children, so that neighbouring queue entries (= the
  // lanes of a warp in the next level) are rays of the same kind from neighbouring pixels ----
  const unsigned lane = threadIdx.x & 31u;
  const unsigned refl_mask = __ballot_sync(0xffffffffu, do_reflect);
  const unsigned refr_mask = __ballot_sync(0xffffffffu, do_refract);
  const unsigned n_refl = (unsigned)__popc(refl_mask), n_refr = (unsigned)__popc(refr_mask);
  const unsigned total = n_refl + n_refr;
  const unsigned below = (1u << lane) - 1u;
  unsigned base = 0;
  if (total > 0u) {
    if (lane == 0u) base = atomicAdd(wf.counters + 0, total);
    base = __shfl_sync(0xffffffffu, base, 0);
  }
  if (do_reflect || do_refract) {
    const unsigned pos_refl = base + (unsigned)__popc(refl_mask & below);
    const unsigned pos_refr = base + n_refl + (unsigned)__popc(refr_mask & below);
    const int next_base = act_base + n;
    if (base + total > (unsigned)wf.queue_cap || (long long)next_base + base + total > (long long)wf.act_cap) {
      wf.counters[1] = 1u;  // overflow: the host retries the frame with larger buffers
    } else {
      const D3 P = Load3(wf.act_point + (size_t)act * 3);
      const unsigned long long path = wf.act_path[act];
      const int pixel = wf.act_pixel[act];
      if (do_reflect) {
        const unsigned pos = pos_refl;
        Count<DBG>(cnt, kReflect);
        const D3 reflected = Load3(wf.act_reflected + (size_t)act * 3);
        const D3 ro = Add(P, MulS(reflected, 0.0001));  // mythtracer.cc:70-75
        Store3(wf.rq_o[qn] + (size_t)pos * 3, ro);
        Store3(wf.rq_d[qn] + (size_t)pos * 3, reflected);
        wf.sort_key[qn][pos] = RayKey(wf, ro, reflected);
        wf.rq_coef[qn][pos] = coef * refl;
        wf.rq_path[qn][pos] = path * 2ull;
        wf.rq_pixel[qn][pos] = pixel;
        wf.rq_inobj[qn][pos] = in_object ? 1 : 0;
        wf.act_refl[act] = next_base + (int)pos;
      }
      if (do_refract) {
        const unsigned pos = pos_refr;
        Count<DBG>(cnt, kRefract);
        const D3 rdir = Normalized(Load3(wf.act_dir + (size_t)act * 3));  // mythtracer.cc:208-212
        const D3 ro = Add(P, MulS(rdir, 0.00001));                        // mythtracer.cc:214-218
        Store3(wf.rq_o[qn] + (size_t)pos * 3, ro);
        Store3(wf.rq_d[qn] + (size_t)pos * 3, rdir);
        wf.sort_key[qn][pos] = RayKey(wf, ro, rdir);
        wf.rq_coef[qn][pos] = coef;
        wf.rq_path[qn][pos] = path * 2ull + 1ull;
        wf.rq_pixel[qn][pos] = pixel;
        wf.rq_inobj[qn][pos] = in_object ? 0 : 1;
        wf.act_refr[act] = next_base + (int)pos;
      }
    }
  }
  if (DBG) FlushCounters<true>(cnt, rp.counters, 0);
}

// ---------------------------------------------------------------------------------------------------
// WfShadow: thread = (light, activation).  Light-major task order keeps the rays of a warp aimed at one light.
// ---------------------------------------------------------------------------------------------------
template <bool DBG>
__global__ void __launch_bounds__(kWfBlock, MTB_WF_MIN_BLOCKS) WfShadow(DeviceScene sc, RenderParams rp, WfBuffers wf, int act_begin, int n,
                                                     int lanes) {
#ifdef MTB_SMEM_TOP
  __shared__ NodeRec top_store[kTopNodes];
  const NodeRec *top = top_store;
  const int top_n = sc.n_nodes < kTopNodes ? sc.n_nodes : kTopNodes;
  StageTopNodes(sc, top_store, sc.n_nodes);
#endif
  const long long task = PackedIndex64((long long)n * sc.n_lights, lanes);
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  unsigned traced = 0;
  if (task < (long long)n * sc.n_lights) {
    const int li = (int)(task / n);
    const int act = act_begin + (int)(task - (long long)li * n);
    if (wf.act_mtl[act] >= 0) {
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
        const int slot = Trace<DBG>(sc, to, ldir, light_distance, &t, cnt MTB_TOP_ARGS);
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
      traced = segments;
      Store3(wf.sh_power + ((size_t)li * wf.act_cap + act) * 3, power);
      wf.sh_flags[(size_t)li * wf.act_cap + act] = (in_shadow ? 1u : 0u) | (segments << 1);
      if (rp.n_rays != nullptr) atomicAdd(rp.n_rays + wf.act_pixel[act], segments);
      WfChargeTile(rp, wf.act_pixel[act], segments);
    }
  }
  FlushCounters<DBG>(cnt, rp.counters, traced);
}

// ---------------------------------------------------------------------------------------------------
// WfLight: Phong sum over the lights in scene order (mythtracer.cc:78-178)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWfBlock) WfLight(DeviceScene sc, RenderParams rp, WfBuffers wf, int act_begin, int n) {
  const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= n) return;
  const int act = act_begin + k;
  const int material = wf.act_mtl[act];
  if (material < 0) return;
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

// parent += child * Refl ; parent += (child * Tf) * Tr   (mythtracer.cc:185-189, 220-224)
__global__ void __launch_bounds__(256) WfFold(DeviceScene sc, WfBuffers wf, int begin, int end) {
  const int a = begin + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (a >= end) return;
  const int rc = wf.act_refl[a], tc = wf.act_refr[a];
  if (rc < 0 && tc < 0) return;
  const mtb_material *m = sc.materials + wf.act_mtl[a];
  D3 color = Load3(wf.act_color + (size_t)a * 3);
  if (rc >= 0) color = Add(color, MulS(Load3(wf.act_color + (size_t)rc * 3), m->reflectance));
  if (tc >= 0) color = Add(color, MulS(MulV(Load3(wf.act_color + (size_t)tc * 3), Load3(m->transmission_filter)), m->transparency));
  Store3(wf.act_color + (size_t)a * 3, color);
}

__global__ void __launch_bounds__(256) WfResolve(RenderParams rp, WfBuffers wf, int n_slots) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n_slots) return;
  // same slot -> pixel mapping as level 0 of WfTraceMain (the level-0 queue itself has been reused by now)
  const int tile = WfTileOfSlot(rp, i), t = i & 63;
  if (tile < 0) return;
  const int strip = rp.strip_first + (tile / rp.tiles_x) * rp.strip_stride;
  const int px = (tile % rp.tiles_x) * 8 + (t & 7);
  const int py = strip * 8 + (t >> 3);
  if (px >= rp.chunk_w || py >= rp.chunk_h) return;
  const int pixel = py * rp.chunk_w + px;
  const D3 c = Load3(wf.act_color + (size_t)i * 3);
  unsigned char *out = rp.rgb + (size_t)pixel * 3;
  out[0] = QuantizeChannel(c.x);
  out[1] = QuantizeChannel(c.y);
  out[2] = QuantizeChannel(c.z);
}

}  // namespace

// Lanes per warp for a queue of `n` items: full warps once the queue fills the machine (148 SMs x 32
// resident warps), otherwise the largest power of two that still spreads it over all warp slots.
static int PackLanes(long long n) {
  const long long slots = 148LL * 32;
  int lanes = 32;
  while (lanes > 4 && n < slots * lanes) lanes >>= 1;
  return lanes;
}

void LaunchWfSort(const WfBuffers &wf, int level, int n, cudaStream_t stream) {
  if (n <= 0) return;
  cudaMemsetAsync(wf.sort_hist, 0, sizeof(uint32_t) << kWfSortBits, stream);
  const uint32_t *keys = wf.sort_key[level & 1];
  WfSortHistogram<<<(n + 255) / 256, 256, 0, stream>>>(keys, wf.sort_hist, n);
  WfSortScan<<<1, 1024, 0, stream>>>(wf.sort_hist);
  WfSortScatter<<<(n + 255) / 256, 256, 0, stream>>>(keys, wf.sort_hist, wf.perm, n);
}

void LaunchWfTraceMain(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int level, int n,
                       int act_base, bool sorted, bool debug_build, cudaStream_t stream) {
  if (n <= 0) return;
  const int lanes = level == 0 ? 32 : PackLanes(n);  // level 0 maps warps to 8x4 pixel tiles
  const long long warps = ((long long)n + lanes - 1) / lanes;
  const int blocks = (int)((warps * 32 + kWfBlock - 1) / kWfBlock);
  const int32_t *perm = sorted ? wf.perm : nullptr;
  if (debug_build) {
    WfTraceMain<true><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, n, act_base, perm, lanes);
  } else {
    WfTraceMain<false><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, n, act_base, perm, lanes);
  }
}

void LaunchWfSpawn(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int level, int n,
                   int act_base, bool debug_build, cudaStream_t stream) {
  if (n <= 0) return;
  const int blocks = (n + kWfBlock - 1) / kWfBlock;
  if (debug_build) {
    WfSpawn<true><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, n, act_base);
  } else {
    WfSpawn<false><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, level, n, act_base);
  }
}

void LaunchWfShadow(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int act_begin, int n,
                    bool debug_build, cudaStream_t stream) {
  const long long tasks = (long long)n * sc.n_lights;
  if (tasks <= 0) return;
  const int lanes = PackLanes(tasks);
  const long long warps = (tasks + lanes - 1) / lanes;
  const int blocks = (int)((warps * 32 + kWfBlock - 1) / kWfBlock);
  if (debug_build) {
    WfShadow<true><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, act_begin, n, lanes);
  } else {
    WfShadow<false><<<blocks, kWfBlock, 0, stream>>>(sc, rp, wf, act_begin, n, lanes);
  }
}

void LaunchWfLight(const DeviceScene &sc, const RenderParams &rp, const WfBuffers &wf, int act_begin, int n,
                   cudaStream_t stream) {
  if (n <= 0) return;
  WfLight<<<(n + kWfBlock - 1) / kWfBlock, kWfBlock, 0, stream>>>(sc, rp, wf, act_begin, n);
}

void LaunchWfFold(const DeviceScene &sc, const WfBuffers &wf, int begin, int end, cudaStream_t stream) {
  if (end <= begin) return;
  WfFold<<<(end - begin + 255) / 256, 256, 0, stream>>>(sc, wf, begin, end);
}

void LaunchWfResolve(const RenderParams &rp, const WfBuffers &wf, int n_slots, cudaStream_t stream) {
  if (n_slots <= 0) return;
  WfResolve<<<(n_slots + 255) / 256, 256, 0, stream>>>(rp, wf, n_slots);
}

}  // namespace mtb
