// RenderMega (per-pixel megakernel) and the batched intersect kernel; see device_core.cuh for the traversal.
#include <atomic>

#include "device_core.cuh"

namespace mtb {
namespace {

// One suspended TraceRayWorker activation waiting for its reflection / refraction child.
struct ShadeFrame {
  D3 color;
  D3 point;
  D3 dir;        // ray.direction of this activation (the refraction child continues along it)
  double coef;   // current_reflection_coef
  unsigned long long path;
  int material;
  unsigned char stage;      // 0: reflection child pending, 1: refraction child pending
  unsigned char do_refract;
  unsigned char in_object;
  unsigned char pad_;
};

// ---------------------------------------------------------------------------------------------------
// RenderMega: one thread = one pixel = the whole TraceRay recursion, evaluated in the reference's
// post-order so that colour sums associate identically.  Every loop iteration issues exactly one
// OctTree::IntersectRay-equivalent query (a primary / reflection / refraction ray or one shadow segment),
// so the lanes of a warp reconverge at the single Trace call site.
// ---------------------------------------------------------------------------------------------------
#ifndef MTB_MEGA_MIN_BLOCKS
#define MTB_MEGA_MIN_BLOCKS 16  // measured on B200 (C3): 1 -> 86 ms, 8 -> 80, 12 -> 66, 16 -> 64 (spills stay in L1)
#endif
// PERSIST: the grid is sized to the machine (SMs x resident blocks) and every LANE draws its next pixel from a
// global counter the moment its current pixel is finished, so a warp never idles lanes behind its most
// expensive pixel (ncu on the one-tile-per-block form: 16 of 32 lanes alive at the Trace call site).  The fetch
// is a branch at the top of the single ray loop - not an outer loop - so freshly fetched lanes trace their
// primary ray in the same Trace call as their neighbours' shadow / secondary rays.  Work items are numbered
// tile by tile in launch order (item >> 6 = position in tile_order, item & 63 = pixel of the 8x8 tile), so a
// warp's 32 consecutive items start as an 8x4 patch.
//
// PACK (kPackThreads = 128 threads = a 16x8 pixel tile): the rays of one loop iteration are traced by the FIRST
// `total` threads of the block instead of by their owners.  Every thread that needs a ray traced writes it to
// shared memory at its rank among the needy threads (warp ballot + per-warp counts), the block synchronises,
// thread i traces ray i, the block synchronises again and the owners pick up (slot, t).  A tile whose pixels are
// mostly finished, or mostly cheap, then keeps one or two warps busy at close to 32 lanes instead of four warps at
// a few lanes each (ncu on the tile-per-block form: 14 of 32 lanes per instruction).  Results cannot change: the
// same rays are traced by the same code, only by different threads.
constexpr int kPackThreads = 128;
constexpr int kPackTileW = 16;
constexpr int kModeTile = 0, kModePersist = 1, kModePack = 2, kModeSync = 3, kModeResume = 4;

template <bool DBG, int MODE>
__global__ void __launch_bounds__(MODE == kModePack ? kPackThreads : kBlockThreads,
                                  MODE == kModePack ? MTB_MEGA_MIN_BLOCKS * kBlockThreads / kPackThreads : MTB_MEGA_MIN_BLOCKS)
    RenderMega(DeviceScene sc, RenderParams rp) {
  constexpr bool PERSIST = MODE == kModePersist;
  constexpr bool PACK = MODE == kModePack;
  // SYNC: the tile-per-block form with the warp re-converged by force at the top of every iteration: a lane whose
  // pixel is finished idles in the loop until the whole warp is, so that __syncwarp() can gather all 32 lanes in
  // front of the Trace call.  (Per-pixel ray counts are balanced - a warp's lanes need 93 % of its maximum on C3 -
  // yet ncu shows 14 of 32 lanes per instruction: without the barrier the lanes that come back from the shadow
  // branch and from the secondary-ray branch walk the traversal as separate groups.)
  // RESUME: SYNC plus suspendable walks (TraceBegin / TraceRun / TraceEnd, device_core.cuh): a lane whose ray is
  // finished shades and starts its next ray while its neighbours' longer rays are parked, instead of waiting.
  constexpr bool RESUME = MODE == kModeResume;
  constexpr bool SYNC = MODE == kModeSync || RESUME;
  FastWalk walk_store;
  unsigned long long walk_stack[RESUME ? kFastStack : 1];
  FastWalk *walk_ptr = &walk_store;
  if (RESUME) asm volatile("" : "+l"(walk_ptr));  // opaque: the parked walk stays in memory between TraceRun calls
  FastWalk &walk = *walk_ptr;
  walk.node = kFastExit;
  bool in_flight = false;
  __shared__ double s_ray[PACK ? 7 * kPackThreads : 1];   // o.xyz, d.xyz, t_limit of the rays of this iteration
  __shared__ double s_res_t[PACK ? kPackThreads : 1];
  __shared__ int s_res_slot[PACK ? kPackThreads : 1];
  __shared__ int s_warp_count[PACK ? kPackThreads / 32 : 1];
#ifdef MTB_SMEM_TOP
  __shared__ NodeRec top_store[kTopNodes];
  const NodeRec *top = top_store;
  const int top_n = sc.n_nodes < kTopNodes ? sc.n_nodes : kTopNodes;
  StageTopNodes(sc, top_store, sc.n_nodes);
#endif
  // hybrid frame: the leading (most expensive) tiles of the order belong to the wavefront pipeline
  if (MODE == kModeTile && rp.heavy_k != nullptr && (int)blockIdx.x < __ldg(rp.heavy_k)) return;
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int i = 0; i < kNumCounters; i++) cnt[i] = 0;
  }
  unsigned long long sig_hits = 0, sig_shadow = 0;
  unsigned n_rays = 0;        // rays of the current pixel
  unsigned rays_total = 0;    // rays of every pixel this lane rendered

  // block -> 8x8 tile of one of this launch's strips (a strip = 8 image rows; strips are interleaved
  // across devices / processes, the in-process form of the reference's master/worker tiling)
  int tile_id = 0, px = 0, py = 0;
  bool active = false;
  bool drawn = false;  // !PERSIST: this thread's one pixel has been taken

  {
    ShadeFrame stack[kMaxRayStack];
    int sp = 0;

    const D3 start = Load3(rp.sensor), d_scan = Load3(rp.sensor + 3), d_pixel = Load3(rp.sensor + 6);
    D3 m_o = Mk(0, 0, 0), m_d = Mk(0, 0, 0);
    int level = 0;
    bool in_object = false;
    double coef = 1.0;
    unsigned long long path = 1;

    // shading context of the current activation (valid while its shadow rays are traced)
    D3 P = Mk(0, 0, 0), normal = Mk(0, 0, 0), surface = Mk(0, 0, 0), reflected = Mk(0, 0, 0), color = Mk(0, 0, 0);
    int material = -1;
    // shadow walk state (mythtracer.cc:86-156)
    int li = 0;
    D3 ldir = Mk(0, 0, 0), lpos = Mk(0, 0, 0), power = Mk(0, 0, 0), seg_start = Mk(0, 0, 0);
    bool in_shadow = false, through = false;
    unsigned segments = 0;
    bool shadow_mode = false;
    D3 final_color = Mk(0, 0, 0);

    for (;;) {
      if (SYNC) {
        __syncwarp();
        if (__all_sync(0xffffffffu, drawn && !active)) break;
      }
      if (!active) {
        // ---- next pixel of this lane ----
        unsigned item;
        bool fresh = !PACK;  // a new pixel was taken: set up its primary ray
        if (SYNC) {
          fresh = false;
          if (!drawn) {
            drawn = true;
            item = blockIdx.x * (unsigned)kBlockThreads + threadIdx.x;
            const int pos = (int)(item >> 6);
            tile_id = rp.tile_order != nullptr ? rp.tile_order[pos] : pos;
            const int strip = rp.strip_first + (tile_id / rp.tiles_x) * rp.strip_stride;
            px = (tile_id % rp.tiles_x) * kTile + (int)(item & 7u);
            py = strip * kTile + (int)((item >> 3) & 7u);
            active = fresh = px < rp.chunk_w && py < rp.chunk_h;
          }
          item = 0;
        } else if (PACK) {
          if (!drawn) {
            drawn = true;
            const int pos = (int)blockIdx.x;
            tile_id = rp.tile_order != nullptr ? rp.tile_order[pos] : pos;
            const int strip = rp.strip_first + (tile_id / rp.tiles_x) * rp.strip_stride;
            px = (tile_id % rp.tiles_x) * kPackTileW + (int)(threadIdx.x & 15u);
            py = strip * kTile + (int)(threadIdx.x >> 4);
            active = fresh = px < rp.chunk_w && py < rp.chunk_h;
          }
          item = 0;
        } else if (PERSIST) {
          const unsigned peers = __activemask();
          const unsigned lane = threadIdx.x & 31u;
          const int leader = __ffs((int)peers) - 1;
          unsigned base = 0;
          if ((int)lane == leader) base = atomicAdd(rp.work_counter, (unsigned)__popc(peers));
          base = __shfl_sync(peers, base, leader);
          item = base + (unsigned)__popc(peers & ((1u << lane) - 1u));
        } else {
          if (drawn) break;
          drawn = true;
          item = blockIdx.x * (unsigned)kBlockThreads + threadIdx.x;
        }
        if (!PACK && !SYNC) {
          if (item >= rp.n_items) break;
          const int pos = (int)(item >> 6);
          tile_id = rp.tile_order != nullptr ? rp.tile_order[pos] : pos;
          const int strip = rp.strip_first + (tile_id / rp.tiles_x) * rp.strip_stride;
          px = (tile_id % rp.tiles_x) * kTile + (int)(item & 7u);
          py = strip * kTile + (int)((item >> 3) & 7u);
          if (!(px < rp.chunk_w && py < rp.chunk_h)) continue;
          active = true;
        }
        if (fresh) {  // (PACK: a thread comes by here every iteration once its pixel is finished or was never live)
        // Sensor::GetRay (camera.cc:65-69) with full-image pixel coordinates (mythtracer.cc:298)
        m_o = Load3(rp.origin);
        m_d = Normalized(Add(Add(start, MulS(d_scan, (double)(rp.chunk_y + py))), MulS(d_pixel, (double)(rp.chunk_x + px))));
        sp = 0;
        level = 0;
        in_object = false;
        coef = 1.0;
        path = 1;
        shadow_mode = false;
        sig_hits = 0;
        sig_shadow = 0;
        n_rays = 0;
        Count<DBG>(cnt, kPrimary);
        }
      }
      D3 to, td;
      double light_distance = 0.0;
      if (shadow_mode) {
        to = Add(seg_start, MulS(ldir, 0.00001));  // mythtracer.cc:95-99
        td = ldir;
        light_distance = Dist(seg_start, lpos);    // mythtracer.cc:101-102
        if (!RESUME || !in_flight) Count<DBG>(cnt, kShadow);
      } else {
        to = m_o;
        td = m_d;
      }
      double t = 0.0;
      int slot;
      if (PACK) {
        // ---- hand the ray to the block: thread i traces the i-th ray of this iteration ----
        const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        const unsigned needy = __ballot_sync(0xffffffffu, active);
        if (lane == 0u) s_warp_count[warp] = __popc(needy);
        __syncthreads();
        int mine = __popc(needy & ((1u << lane) - 1u)), total = 0;
#pragma unroll
        for (int w = 0; w < kPackThreads / 32; w++) {
          const int c = s_warp_count[w];
          mine += (unsigned)w < warp ? c : 0;
          total += c;
        }
        if (total == 0) break;  // every pixel of the tile is finished (uniform across the block)
        if (active) {
          s_ray[0 * kPackThreads + mine] = to.x;
          s_ray[1 * kPackThreads + mine] = to.y;
          s_ray[2 * kPackThreads + mine] = to.z;
          s_ray[3 * kPackThreads + mine] = td.x;
          s_ray[4 * kPackThreads + mine] = td.y;
          s_ray[5 * kPackThreads + mine] = td.z;
          s_ray[6 * kPackThreads + mine] = shadow_mode ? light_distance : CUDART_INF;
        }
        __syncthreads();
        if ((int)threadIdx.x < total) {
          const int i = (int)threadIdx.x;
          const D3 ro = Mk(s_ray[0 * kPackThreads + i], s_ray[1 * kPackThreads + i], s_ray[2 * kPackThreads + i]);
          const D3 rd = Mk(s_ray[3 * kPackThreads + i], s_ray[4 * kPackThreads + i], s_ray[5 * kPackThreads + i]);
          double rt = 0.0;
          s_res_slot[i] = Trace<DBG>(sc, ro, rd, s_ray[6 * kPackThreads + i], &rt, cnt MTB_TOP_ARGS);
          s_res_t[i] = rt;
        }
        __syncthreads();
        if (!active) continue;
        slot = s_res_slot[mine];
        t = s_res_t[mine];
      } else if (RESUME) {
        if (active && !in_flight) {
          TraceBegin<DBG>(sc, to, td, shadow_mode ? light_distance : CUDART_INF, &walk, cnt);
          in_flight = true;
        }
        __syncwarp();
        TraceRun<DBG>(sc, &walk, walk_stack, cnt);
        if (!active || walk.node != kFastExit) continue;  // idle, or this lane's ray is parked
        in_flight = false;
        slot = TraceEnd<DBG>(sc, &walk, &t, cnt MTB_TOP_ARGS);
      } else if (SYNC) {
        __syncwarp();
        if (!active) continue;
        slot = Trace<DBG>(sc, to, td, shadow_mode ? light_distance : CUDART_INF, &t, cnt MTB_TOP_ARGS);
      } else {
        slot = Trace<DBG>(sc, to, td, shadow_mode ? light_distance : CUDART_INF, &t, cnt MTB_TOP_ARGS);
      }
      n_rays++;

      bool have_ret = false;
      D3 ret = Mk(0.0, 0.0, 0.0);
      if (shadow_mode) {
        bool light_done = false;
        segments++;
        if (slot < 0) {
          light_done = true;  // mythtracer.cc:109-112
        } else if (t > light_distance) {
          light_done = true;  // mythtracer.cc:115-118
        } else {
          const int smtl = __ldg(&sc.shade[slot].material);
          // mtl is dereferenced unconditionally upstream (mythtracer.cc:121); a missing material acts opaque
          const double str = smtl >= 0 ? __ldg(&sc.materials[smtl].transparency) : 0.0;
          if (str == 0.0) {
            power = Mk(0.0, 0.0, 0.0);
            in_shadow = true;
            light_done = true;
          } else {
            if (!through) {  // light_power *= Tf * Tr (mythtracer.cc:129-132)
              const D3 tf = Load3(sc.materials[smtl].transmission_filter);
              power = MulV(power, MulS(tf, str));
            }
            through = !through;
            const D3 hit_point = Add(to, MulS(td, t));               // primitive_triangle.cc:141
            seg_start = Add(hit_point, MulS(ldir, 0.0000001));       // mythtracer.cc:137
            if (SqrDist(P, seg_start) > SqrDist(P, lpos)) {          // mythtracer.cc:141-145
              light_done = true;
            } else if (power.x <= 0.001 && power.y <= 0.001 && power.z <= 0.001) {
              power = Mk(0.0, 0.0, 0.0);
              in_shadow = true;
              light_done = true;
            }
          }
        }
        if (!light_done) continue;  // next segment of the same light

        // ---- this light is settled: Phong terms (mythtracer.cc:159-177) ----
        const mtb_light *lt = sc.lights + li;
        const mtb_material *m = sc.materials + material;
        sig_shadow += Mix64(path, 2ull + (unsigned long long)li, (in_shadow ? 1ull : 0ull) | ((unsigned long long)segments << 1));
        const D3 lamb = Load3(lt->ambient);
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
        li++;
      } else {
        // ---- result of a primary / reflection / refraction ray (mythtracer.cc:13-76) ----
        if (slot < 0) {
          if (level == 0 && rp.dbg != nullptr) {
            mtb_debug *dbg = rp.dbg + (size_t)py * rp.chunk_w + px;
            dbg->line_no = -1;
            dbg->pad_ = 0;
            dbg->point[0] = dbg->point[1] = dbg->point[2] = CUDART_NAN;
          }
          have_ret = true;  // background colour {0,0,0}
        } else {
          const ShadeRec *sh = sc.shade + slot;
          const SlotRec *sr = sc.slots + slot;
          P = Add(to, MulS(td, t));
          const int line_no = __ldg(&sh->line_no);
          if (level == 0 && rp.dbg != nullptr) {
            mtb_debug *dbg = rp.dbg + (size_t)py * rp.chunk_w + px;
            dbg->line_no = line_no;
            dbg->pad_ = 0;
            dbg->point[0] = P.x;
            dbg->point[1] = P.y;
            dbg->point[2] = P.z;
          }
          sig_hits += Mix64(path, 1ull, (unsigned long long)(long long)line_no);
          Count<DBG>(cnt, kShade);
          const D3 v0 = Load3(sr->vert), v1 = Load3(sr->vert + 3), v2 = Load3(sr->vert + 6);
          const BaryWeights w = Barycentric(v0, v1, v2, P);
          // (normal[0]*n0 + normal[1]*n1 + normal[2]*n2) / n, not normalised (primitive_triangle.cc:60)
          normal = DivS(Add(Add(MulS(Load3(sh->normal), w.n0), MulS(Load3(sh->normal + 3), w.n1)), MulS(Load3(sh->normal + 6), w.n2)), w.n);
          const D3 towards_camera = Neg(m_d);
          double normal_ray_dot = Dot(towards_camera, normal);
          if (normal_ray_dot < 0.0) {
            normal = Neg(normal);
            normal_ray_dot = Dot(towards_camera, normal);
          }
          material = __ldg(&sh->material);
          if (material < 0) {  // mythtracer.cc:49-52
            normal_ray_dot = (normal_ray_dot + 1.0) * 0.5;
            ret = Mk(normal_ray_dot, normal_ray_dot, normal_ray_dot);
            have_ret = true;
          } else {
            const mtb_material *m = sc.materials + material;
            surface = Load3(m->ambient);
            const int tex = m->texture;
            if (tex >= 0) {  // mythtracer.cc:59-64
              const double u = (sh->uv[0] * w.n0 + sh->uv[2] * w.n1 + sh->uv[4] * w.n2) / w.n;
              const double v = (sh->uv[1] * w.n0 + sh->uv[3] * w.n1 + sh->uv[5] * w.n2) / w.n;
              surface = MulV(surface, SampleTexture(sc.tex_atlas, tex, sc.texture_dim[tex], u, v));
            }
            // ray.direction - normal * (2 * ray.direction.Dot(normal)) (mythtracer.cc:68-69)
            reflected = Sub(m_d, MulS(normal, 2 * Dot(normal, m_d)));
            color = Mk(0.0, 0.0, 0.0);
            li = 0;
          }
        }
      }

      if (!have_ret) {
        if (li < sc.n_lights) {
          // ---- start the shadow walk of light li (mythtracer.cc:79-94) ----
          const mtb_light *lt = sc.lights + li;
          lpos = Load3(lt->position);
          ldir = Normalized(Sub(lpos, P));
          color = Add(color, MulV(Load3(lt->ambient), surface));  // mythtracer.cc:83-84
          power = Mk(1.0, 1.0, 1.0);
          in_shadow = false;
          through = false;
          segments = 0;
          seg_start = P;
          shadow_mode = true;
          continue;
        }
        // ---- all lights done: secondary rays (mythtracer.cc:181-225) ----
        shadow_mode = false;
        const mtb_material *m = sc.materials + material;
        const double refl = m->reflectance, tr = m->transparency;
        const bool do_reflect = level < rp.max_depth && refl > 0.0 && coef > 0.01 && !in_object;
        const bool do_refract = level < rp.max_depth && tr > 0.0;
        if (do_reflect || do_refract) {
          ShadeFrame &f = stack[sp];
          f.color = color;
          f.point = P;
          f.dir = m_d;
          f.coef = coef;
          f.path = path;
          f.material = material;
          f.stage = do_reflect ? 0 : 1;
          f.do_refract = do_refract ? 1 : 0;
          f.in_object = in_object ? 1 : 0;
          sp++;
          level++;
          if (do_reflect) {
            Count<DBG>(cnt, kReflect);
            m_o = Add(P, MulS(reflected, 0.0001));  // mythtracer.cc:70-75
            m_d = reflected;
            coef = coef * refl;
            path = path * 2ull;
          } else {
            Count<DBG>(cnt, kRefract);
            const D3 rdir = Normalized(m_d);       // mythtracer.cc:208-212
            m_o = Add(P, MulS(rdir, 0.00001));     // mythtracer.cc:214-218
            m_d = rdir;
            in_object = !in_object;
            path = path * 2ull + 1ull;
          }
          continue;
        }
        ret = color;
      }

      // ---- an activation returned `ret`: fold it into suspended parents (mythtracer.cc:185-189,220-224) ----
      bool finished = false;
      for (;;) {
        if (sp == 0) {
          final_color = ret;
          finished = true;
          break;
        }
        ShadeFrame &f = stack[sp - 1];
        const mtb_material *m = sc.materials + f.material;
        if (f.stage == 0) {
          f.color = Add(f.color, MulS(ret, m->reflectance));
          if (f.do_refract) {
            f.stage = 1;
            Count<DBG>(cnt, kRefract);
            const D3 rdir = Normalized(f.dir);
            m_o = Add(f.point, MulS(rdir, 0.00001));
            m_d = rdir;
            in_object = !(f.in_object != 0);
            coef = f.coef;
            path = f.path * 2ull + 1ull;
            level = sp;
            shadow_mode = false;
            break;  // trace the refraction child
          }
          ret = f.color;
          sp--;
        } else {
          // (c * Tf) * Tr (mythtracer.cc:224)
          f.color = Add(f.color, MulS(MulV(ret, Load3(m->transmission_filter)), m->transparency));
          ret = f.color;
          sp--;
        }
      }
      if (!finished) continue;

      // ---- the pixel is done (mythtracer.cc:301) ----
      const size_t pix = (size_t)py * rp.chunk_w + px;
      unsigned char *out = rp.rgb + pix * 3;
      out[0] = QuantizeChannel(final_color.x);
      out[1] = QuantizeChannel(final_color.y);
      out[2] = QuantizeChannel(final_color.z);
      if (rp.sig_hits != nullptr) rp.sig_hits[pix] = sig_hits;
      if (rp.sig_shadow != nullptr) rp.sig_shadow[pix] = sig_shadow;
      if (rp.n_rays != nullptr) rp.n_rays[pix] = n_rays;
      rays_total += n_rays;
      // what this tile cost, for the next frame's launch order
      if (PERSIST && rp.tile_cost != nullptr) atomicAdd(rp.tile_cost + tile_id, n_rays);
      active = false;
      if (PACK || SYNC) shadow_mode = false;
    }
  }

  if (!PERSIST && rp.tile_cost != nullptr) {
    unsigned v = rays_total;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31u) == 0u) atomicAdd(rp.tile_cost + (rp.tile_order != nullptr ? rp.tile_order[blockIdx.x] : (int)blockIdx.x), v);
  }
  if (DBG && rp.counters != nullptr) {
    for (int i = 0; i < kNumCounters; i++) {
      unsigned long long v = cnt[i];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(rp.counters + i, v);
    }
  } else if (rp.counters != nullptr) {
    // the fast build still reports the ray count (the metric's numerator)
    unsigned long long v = rays_total;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(rp.counters + kRays, v);
  }
}

// Batched OctTree::IntersectRay (octtree.cc:26-40), one ray per thread.
template <bool DBG>
__global__ void __launch_bounds__(128) IntersectKernel(DeviceScene sc, IntersectParams ip) {
#ifdef MTB_SMEM_TOP
  __shared__ NodeRec top_store[kTopNodes];
  const NodeRec *top = top_store;
  const int top_n = sc.n_nodes < kTopNodes ? sc.n_nodes : kTopNodes;
  StageTopNodes(sc, top_store, sc.n_nodes);
#endif
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cnt_store[DBG ? kNumCounters : 1];
  unsigned long long *cnt = cnt_store;
  if (DBG) {
    for (int k = 0; k < kNumCounters; k++) cnt[k] = 0;
  }
  if (i < ip.n) {
    const D3 o = Load3(ip.origins + i * 3), d = Load3(ip.dirs + i * 3);
    double t = 0.0;
    const int slot = Trace<DBG>(sc, o, d, CUDART_INF, &t, cnt MTB_TOP_ARGS);
    if (slot < 0) {
      ip.tri_index[i] = -1;
    } else {
      ip.tri_index[i] = sc.slots[slot].tri;
      if (ip.t != nullptr) ip.t[i] = t;
      if (ip.point != nullptr) {
        const D3 p = Add(o, MulS(d, t));
        ip.point[i * 3 + 0] = p.x;
        ip.point[i * 3 + 1] = p.y;
        ip.point[i * 3 + 2] = p.z;
      }
    }
  }
  if (DBG && ip.counters != nullptr) {
    for (int k = 0; k < kNumCounters; k++) {
      unsigned long long v = cnt[k];
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31u) == 0u && v != 0ull) atomicAdd(ip.counters + k, v);
    }
  }
}

// Counting sort of the tiles by cost bucket (log2 of the ray count, most expensive first).  One block.
__global__ void __launch_bounds__(1024) BuildTileOrder(uint32_t *tile_cost, int32_t *tile_order, int n_tiles, int32_t *heavy_k, int k_max, int heavy_factor) {
  __shared__ unsigned hist[33];
  __shared__ unsigned offset[33];
  __shared__ unsigned long long total;
  if (threadIdx.x < 33) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  unsigned long long mine = 0;
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const unsigned c = tile_cost[t];
    mine += c;
    atomicAdd(&hist[c == 0u ? 0 : 32 - __clz((int)c)], 1u);
  }
  atomicAdd(&total, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int b = 32; b >= 0; b--) {
      offset[b] = run;
      run += hist[b];
    }
    if (heavy_k != nullptr) {
      // heavy = whole buckets [2^(b-1), 2^b) whose lower bound is at least heavy_factor x the mean cost, most
      // expensive bucket first, as long as they fit into k_max tiles
      const unsigned long long mean = n_tiles > 0 ? total / (unsigned long long)n_tiles : 0;
      int k = 0;
      for (int b = 32; b >= 2; b--) {
        if ((1ull << (b - 1)) < (unsigned long long)heavy_factor * mean || mean == 0) break;
        if (k + (int)hist[b] > k_max) break;
        k += (int)hist[b];
      }
      *heavy_k = k;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const unsigned c = tile_cost[t];
    const unsigned pos = atomicAdd(&offset[c == 0u ? 0 : 32 - __clz((int)c)], 1u);
    tile_order[pos] = t;
    tile_cost[t] = 0;
  }
}

}  // namespace

void LaunchBuildTileOrder(uint32_t *tile_cost, int32_t *tile_order, int n_tiles, int32_t *heavy_k, int k_max, int heavy_factor,
                          cudaStream_t stream) {
  if (n_tiles <= 0) return;
  BuildTileOrder<<<1, 1024, 0, stream>>>(tile_cost, tile_order, n_tiles, heavy_k, k_max, heavy_factor);
}

int MegaResidentBlocks(int device) {
  static std::atomic<int> cached[64];  // devices are driven by one host thread each (api.cu)
  if (device >= 0 && device < 64 && cached[device].load() > 0) return cached[device].load();
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, RenderMega<false, kModePersist>, kBlockThreads, 0) != cudaSuccess) per_sm = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) sms = 0;
  const int n = per_sm > 0 && sms > 0 ? per_sm * sms : 148 * MTB_MEGA_MIN_BLOCKS;
  if (device >= 0 && device < 64) cached[device].store(n);
  return n;
}

int MegaTileWidth(int mode) { return mode == kModePack ? kPackTileW : kTile; }

// n_blocks = number of tiles (8x8 pixels; 16x8 in pack mode).  mode 1 (persistent): `persistent_blocks` blocks
// (never more than there are tiles) and rp.work_counter must point at a zeroed counter.
void LaunchRenderMega(const DeviceScene &sc, const RenderParams &rp_in, int n_blocks, int mode, int persistent_blocks,
                      bool debug_build, cudaStream_t stream) {
  if (n_blocks <= 0) return;
  RenderParams rp = rp_in;
  rp.n_items = (uint32_t)n_blocks * (uint32_t)kBlockThreads;
  if (mode == kModePack) {
    if (debug_build) {
      RenderMega<true, kModePack><<<n_blocks, kPackThreads, 0, stream>>>(sc, rp);
    } else {
      RenderMega<false, kModePack><<<n_blocks, kPackThreads, 0, stream>>>(sc, rp);
    }
  } else if (mode == kModePersist) {
    const int grid = persistent_blocks < n_blocks ? persistent_blocks : n_blocks;
    if (debug_build) {
      RenderMega<true, kModePersist><<<grid, kBlockThreads, 0, stream>>>(sc, rp);
    } else {
      RenderMega<false, kModePersist><<<grid, kBlockThreads, 0, stream>>>(sc, rp);
    }
  } else if (mode == kModeSync) {
    if (debug_build) {
      RenderMega<true, kModeSync><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
    } else {
      RenderMega<false, kModeSync><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
    }
  } else if (mode == kModeResume) {
    if (debug_build) {
      RenderMega<true, kModeResume><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
    } else {
      RenderMega<false, kModeResume><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
    }
  } else if (debug_build) {
    RenderMega<true, kModeTile><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
  } else {
    RenderMega<false, kModeTile><<<n_blocks, kBlockThreads, 0, stream>>>(sc, rp);
  }
}

void LaunchIntersect(const DeviceScene &sc, const IntersectParams &ip, bool debug_build, cudaStream_t stream) {
  if (ip.n <= 0) return;
  const int blocks = (int)((ip.n + 127) / 128);
  if (debug_build) {
    IntersectKernel<true><<<blocks, 128, 0, stream>>>(sc, ip);
  } else {
    IntersectKernel<false><<<blocks, 128, 0, stream>>>(sc, ip);
  }
}

}  // namespace mtb
