// Device-side build of the scene BVH: the acceleration structure of the certified fast
// traversal is built on the GPU from the triangle reference boxes, next to the data it indexes, instead of on the host.
//
// Algorithm: PLOC - parallel locally-ordered clustering (Meister & Bittner 2017), a bottom-up agglomerative build:
//   1. the references are sorted along a 63-bit Morton curve of their box centres (cub radix sort);
//   2. every cluster looks for its best partner - the one whose union box has the smallest surface area - among its
//      2 x radius neighbours in curve order; clusters that choose EACH OTHER are merged; the survivors are
//      compacted (prefix sums); repeated until one cluster is left (a constant fraction merges per round).
// Two single-reference clusters that merge become one two-reference LEAF (the scene BVH's leaf size, scene_build.h);
// every other merge creates a Bvh2Node, which holds exactly what PLOC has at that moment: the two children's boxes.
// Boxes are unions of the exact FP64 reference boxes (no rounding on the way up); a stored child box is that union
// grown by the FP32 slab test's padding and rounded outwards to float, the same formula as the host builder
// (SceneBvhBuilder::SetChild).  Which tree is built can only change the SPEED of a traversal, never its result:
// the leaves are decided by the reference's exact tests.
//
// Everything is deterministic: merges are mutual-nearest pairs with index tie breaks, node numbers and leaf positions
// come from prefix sums.  Nodes are renumbered so that node 0 is the root and parents precede their children.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "device_scene.h"

namespace mtb {
namespace {

constexpr int kPlocRadiusDefault = 64;  // C3 frame time with the tree it gives (B200): see DESIGN.md section 10, f1
constexpr int kPlocRadiusMax = 128;
constexpr int kPlocBlock = 256;

struct Cluster {
  double lo[3], hi[3];
  int32_t id;      // >= 0: node (creation order); < 0: leaf, ~id = first reference
  int32_t ref2;    // second reference of a two-reference leaf, -1: none
  int32_t height;  // 0 for a leaf
  int32_t pad_;
};

__device__ __forceinline__ unsigned long long Spread21(unsigned long long v) {  // 21 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void PlocMorton(const double *__restrict__ box, int n, double3 lo, double3 inv_ext, unsigned long long *keys, int32_t *vals) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const double *b = box + (size_t)i * 6;
  const double cx = ((b[0] + b[3]) * 0.5 - lo.x) * inv_ext.x, cy = ((b[1] + b[4]) * 0.5 - lo.y) * inv_ext.y,
               cz = ((b[2] + b[5]) * 0.5 - lo.z) * inv_ext.z;
  const double s = 2097151.0;  // 2^21 - 1
  const unsigned long long x = (unsigned long long)fmin(fmax(cx * s, 0.0), s), y = (unsigned long long)fmin(fmax(cy * s, 0.0), s),
                           z = (unsigned long long)fmin(fmax(cz * s, 0.0), s);
  keys[i] = (Spread21(x) << 2) | (Spread21(y) << 1) | Spread21(z);
  vals[i] = i;
}

__global__ void PlocInit(const double *__restrict__ box, const int32_t *__restrict__ order, int n, Cluster *out) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const int32_t r = order[i];
  const double *b = box + (size_t)r * 6;
  Cluster c;
  for (int a = 0; a < 3; a++) {
    c.lo[a] = b[a];
    c.hi[a] = b[3 + a];
  }
  c.id = ~r;
  c.ref2 = -1;
  c.height = 0;
  c.pad_ = 0;
  out[i] = c;
}

// Best partner of every cluster among its neighbours in curve order (smallest surface area of the union; the lower
// index wins ties).  The boxes of a block's clusters and their halo are staged in shared memory as floats: the
// choice is a heuristic, the unions themselves are made from the exact boxes.
__global__ void __launch_bounds__(kPlocBlock) PlocNearest(const Cluster *__restrict__ c, int n, int radius, int32_t *nn) {
  __shared__ float s_box[kPlocBlock + 2 * kPlocRadiusMax][6];
  const int base = (int)(blockIdx.x * kPlocBlock) - radius;
  for (int k = (int)threadIdx.x; k < kPlocBlock + 2 * radius; k += kPlocBlock) {
    const int j = base + k;
    if (j >= 0 && j < n) {
      for (int a = 0; a < 3; a++) {
        s_box[k][a] = (float)c[j].lo[a];
        s_box[k][3 + a] = (float)c[j].hi[a];
      }
    }
  }
  __syncthreads();
  const int i = (int)(blockIdx.x * kPlocBlock + threadIdx.x);
  if (i >= n) return;
  const int me = (int)threadIdx.x + radius;
  float best = INFINITY;
  int best_j = -1;
  for (int d = -radius; d <= radius; d++) {
    const int j = i + d;
    if (d == 0 || j < 0 || j >= n) continue;
    const int k = me + d;
    const float dx = fmaxf(s_box[me][3], s_box[k][3]) - fminf(s_box[me][0], s_box[k][0]);
    const float dy = fmaxf(s_box[me][4], s_box[k][4]) - fminf(s_box[me][1], s_box[k][1]);
    const float dz = fmaxf(s_box[me][5], s_box[k][5]) - fminf(s_box[me][2], s_box[k][2]);
    const float area = dx * dy + dy * dz + dz * dx;
    if (area < best) {  // strict: the lower index keeps a tie
      best = area;
      best_j = j;
    }
  }
  nn[i] = best_j;
}

// flags[i].x = 1 if cluster i survives this round (everything but the higher index of a merging pair),
// flags[i].y = 1 if it is the lower index of a pair whose merge creates a NODE (not a two-reference leaf).
__global__ void PlocFlags(const Cluster *__restrict__ c, const int32_t *__restrict__ nn, int n, int32_t *keep, int32_t *makes_node) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const int j = nn[i];
  const bool mutual = j >= 0 && nn[j] == i;
  keep[i] = (mutual && j < i) ? 0 : 1;
  int node = 0;
  if (mutual && i < j) {
    const bool both_single = c[i].id < 0 && c[i].ref2 < 0 && c[j].id < 0 && c[j].ref2 < 0;
    node = both_single ? 0 : 1;
  }
  makes_node[i] = node;
}

__device__ __forceinline__ float BoxDown(double x) { return __double2float_rd(x); }
__device__ __forceinline__ float BoxUp(double x) { return __double2float_ru(x); }

// Raw node of the build: final boxes, children still in build encoding (leaf: the cluster's references).
struct RawNode {
  float lbox[6], rbox[6];
  int32_t lid, lref2, rid, rref2;  // id >= 0: node in creation order; < 0: leaf with references ~id (and ref2)
};

__global__ void PlocMerge(const Cluster *__restrict__ c, const int32_t *__restrict__ nn, const int32_t *__restrict__ keep_pos,
                          const int32_t *__restrict__ node_pos, const int32_t *__restrict__ keep, const int32_t *__restrict__ makes_node,
                          int n, int node_base, double pad, Cluster *out, RawNode *nodes) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n || keep[i] == 0) return;
  const int j = nn[i];
  const bool merge = j > i && nn[j] == i;
  Cluster r = c[i];
  if (merge) {
    const Cluster o = c[j];
    if (makes_node[i]) {
      const int k = node_base + node_pos[i];
      RawNode nd;
      for (int a = 0; a < 3; a++) {
        nd.lbox[a] = BoxDown(r.lo[a] - pad);
        nd.lbox[3 + a] = BoxUp(r.hi[a] + pad);
        nd.rbox[a] = BoxDown(o.lo[a] - pad);
        nd.rbox[3 + a] = BoxUp(o.hi[a] + pad);
      }
      nd.lid = r.id;
      nd.lref2 = r.ref2;
      nd.rid = o.id;
      nd.rref2 = o.ref2;
      nodes[k] = nd;
      r.id = k;
      r.ref2 = -1;
      r.height = (r.height > o.height ? r.height : o.height) + 1;
    } else {
      r.ref2 = ~o.id;  // two single references -> one leaf
    }
    for (int a = 0; a < 3; a++) {
      r.lo[a] = fmin(r.lo[a], o.lo[a]);
      r.hi[a] = fmax(r.hi[a], o.hi[a]);
    }
  }
  out[keep_pos[i]] = r;
}

// Leaf children -> number of references (0 for inner children), 2 entries per raw node (left, right).
__global__ void PlocLeafCounts(const RawNode *__restrict__ nodes, int n_nodes, int32_t *counts) {
  const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= n_nodes) return;
  // leaf positions are handed out in FINAL node order (root first): raw node k becomes node n_nodes - 1 - k
  const int f = n_nodes - 1 - k;
  const RawNode &nd = nodes[k];
  counts[2 * f + 0] = nd.lid < 0 ? (nd.lref2 >= 0 ? 2 : 1) : 0;
  counts[2 * f + 1] = nd.rid < 0 ? (nd.rref2 >= 0 ? 2 : 1) : 0;
}

__global__ void PlocFinalize(const RawNode *__restrict__ nodes, int n_nodes, const int32_t *__restrict__ first, Bvh2Node *out, int32_t *leaf_ref) {
  const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (k >= n_nodes) return;
  const int f = n_nodes - 1 - k;
  const RawNode &nd = nodes[k];
  Bvh2Node o;
  for (int a = 0; a < 6; a++) {
    o.lbox[a] = nd.lbox[a];
    o.rbox[a] = nd.rbox[a];
  }
  o.pad_[0] = o.pad_[1] = 0;
  if (nd.lid >= 0) {
    o.left = n_nodes - 1 - nd.lid;
  } else {
    const int pos = first[2 * f + 0], cnt = nd.lref2 >= 0 ? 2 : 1;
    leaf_ref[pos] = ~nd.lid;
    if (cnt == 2) leaf_ref[pos + 1] = nd.lref2;
    o.left = ~(int32_t)(((uint32_t)pos << 3) | (uint32_t)cnt);
  }
  if (nd.rid >= 0) {
    o.right = n_nodes - 1 - nd.rid;
  } else {
    const int pos = first[2 * f + 1], cnt = nd.rref2 >= 0 ? 2 : 1;
    leaf_ref[pos] = ~nd.rid;
    if (cnt == 2) leaf_ref[pos + 1] = nd.rref2;
    o.right = ~(int32_t)(((uint32_t)pos << 3) | (uint32_t)cnt);
  }
  out[f] = o;
}

// gslots[pos] = slots[ref_slot[leaf_ref[pos]]]: the 128-byte triangle records in leaf order, gathered on the device.
__global__ void GatherLeafSlots(const SlotRec *__restrict__ slots, const int32_t *__restrict__ ref_slot, const int32_t *__restrict__ leaf_ref,
                                int n_positions, SlotRec *gslots) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte word per thread
  const long long pos = t >> 3;
  if (pos >= n_positions) return;
  const int w = (int)(t & 7);
  const uint4 *src = reinterpret_cast<const uint4 *>(slots + ref_slot[leaf_ref[pos]]);
  reinterpret_cast<uint4 *>(gslots + pos)[w] = src[w];
}

template <typename T>
struct Scratch {
  T *p = nullptr;
  cudaError_t Alloc(size_t n) { return cudaMalloc(reinterpret_cast<void **>(&p), (n > 0 ? n : 1) * sizeof(T)); }
  ~Scratch() {
    if (p != nullptr) cudaFree(p);
  }
};

#define MTB_TRY(expr)                    \
  do {                                   \
    cudaError_t e__ = (expr);            \
    if (e__ != cudaSuccess) return e__;  \
  } while (0)

}  // namespace

cudaError_t BuildSceneBvhOnDevice(const double *h_ref_box, int64_t n_refs, const double scene_box[6], double pad, cudaStream_t stream,
                                  Bvh2Node **d_nodes_out, int32_t *n_nodes_out, int32_t **d_leaf_ref_out, int32_t *depth_out) {
  *d_nodes_out = nullptr;
  *d_leaf_ref_out = nullptr;
  *n_nodes_out = 0;
  *depth_out = 0;
  if (n_refs < 3 || n_refs > 0x0fffffff) return cudaErrorInvalidValue;  // (tiny scenes: the host builder's single-leaf forms)
  const int n = (int)n_refs;
  // one allocation for all the scratch arrays (a dozen cudaMalloc calls cost more than the build itself)
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int32_t *)nullptr,
                                  (int32_t *)nullptr, n, 0, 63, stream);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int32_t *)nullptr, (int32_t *)nullptr, 2 * n + 2, stream);
  const size_t tmp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  struct Carver {
    size_t used = 0;
    char *base = nullptr;
    size_t Take(size_t bytes) {
      const size_t at = used;
      used += (bytes + 255) & ~(size_t)255;
      return at;
    }
  } carve;
  const size_t o_box = carve.Take((size_t)n * 6 * sizeof(double)), o_keys = carve.Take((size_t)n * 8), o_keys2 = carve.Take((size_t)n * 8),
               o_vals = carve.Take((size_t)n * 4), o_order = carve.Take((size_t)n * 4), o_nn = carve.Take((size_t)n * 4),
               o_keep = carve.Take((size_t)n * 4), o_makes = carve.Take((size_t)n * 4), o_keep_pos = carve.Take((size_t)n * 4),
               o_node_pos = carve.Take((size_t)n * 4), o_ca = carve.Take((size_t)n * sizeof(Cluster)), o_cb = carve.Take((size_t)n * sizeof(Cluster)),
               o_raw = carve.Take((size_t)n * sizeof(RawNode)), o_counts = carve.Take((2 * (size_t)n + 2) * 4),
               o_first = carve.Take((2 * (size_t)n + 2) * 4), o_tmp = carve.Take(tmp_bytes);
  Scratch<char> arena;
  MTB_TRY(arena.Alloc(carve.used));
  struct {
    double *p;
  } box{reinterpret_cast<double *>(arena.p + o_box)};
  struct {
    unsigned long long *p;
  } keys{reinterpret_cast<unsigned long long *>(arena.p + o_keys)}, keys_sorted{reinterpret_cast<unsigned long long *>(arena.p + o_keys2)};
  struct I32 {
    int32_t *p;
  };
  const I32 vals{reinterpret_cast<int32_t *>(arena.p + o_vals)}, order{reinterpret_cast<int32_t *>(arena.p + o_order)},
      nn{reinterpret_cast<int32_t *>(arena.p + o_nn)}, keep{reinterpret_cast<int32_t *>(arena.p + o_keep)},
      makes_node{reinterpret_cast<int32_t *>(arena.p + o_makes)}, keep_pos{reinterpret_cast<int32_t *>(arena.p + o_keep_pos)},
      node_pos{reinterpret_cast<int32_t *>(arena.p + o_node_pos)}, counts{reinterpret_cast<int32_t *>(arena.p + o_counts)},
      first{reinterpret_cast<int32_t *>(arena.p + o_first)};
  struct {
    Cluster *p;
  } ca{reinterpret_cast<Cluster *>(arena.p + o_ca)}, cb{reinterpret_cast<Cluster *>(arena.p + o_cb)};
  struct {
    RawNode *p;
  } raw{reinterpret_cast<RawNode *>(arena.p + o_raw)};
  struct {
    unsigned char *p;
  } tmp{reinterpret_cast<unsigned char *>(arena.p + o_tmp)};

  MTB_TRY(cudaMemcpyAsync(box.p, h_ref_box, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, stream));
  const int blocks = (n + kPlocBlock - 1) / kPlocBlock;
  double3 lo = make_double3(scene_box[0], scene_box[1], scene_box[2]);
  double3 inv = make_double3(scene_box[3] > scene_box[0] ? 1.0 / (scene_box[3] - scene_box[0]) : 0.0,
                             scene_box[4] > scene_box[1] ? 1.0 / (scene_box[4] - scene_box[1]) : 0.0,
                             scene_box[5] > scene_box[2] ? 1.0 / (scene_box[5] - scene_box[2]) : 0.0);
  PlocMorton<<<blocks, kPlocBlock, 0, stream>>>(box.p, n, lo, inv, keys.p, vals.p);
  size_t bytes = tmp_bytes;
  MTB_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys_sorted.p, vals.p, order.p, n, 0, 63, stream));
  PlocInit<<<blocks, kPlocBlock, 0, stream>>>(box.p, order.p, n, ca.p);

  int radius = kPlocRadiusDefault;
  if (const char *env = getenv("MTB_PLOC_RADIUS")) {  // development knob
    const int v = atoi(env);
    if (v >= 1 && v <= kPlocRadiusMax) radius = v;
  }
  Cluster *cur = ca.p, *nxt = cb.p;
  int count = n, n_nodes = 0;
  for (int round = 0; count > 1; round++) {
    if (round > 4096) return cudaErrorUnknown;  // (cannot happen: every round merges at least the globally best pair)
    const int b = (count + kPlocBlock - 1) / kPlocBlock;
    PlocNearest<<<b, kPlocBlock, 0, stream>>>(cur, count, radius, nn.p);
    PlocFlags<<<b, kPlocBlock, 0, stream>>>(cur, nn.p, count, keep.p, makes_node.p);
    bytes = tmp_bytes;
    MTB_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, keep.p, keep_pos.p, count, stream));
    bytes = tmp_bytes;
    MTB_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, makes_node.p, node_pos.p, count, stream));
    PlocMerge<<<b, kPlocBlock, 0, stream>>>(cur, nn.p, keep_pos.p, node_pos.p, keep.p, makes_node.p, count, n_nodes, pad, nxt, raw.p);
    int32_t tail[4];  // last elements of the two flag arrays and of their scans -> the new counts
    MTB_TRY(cudaMemcpyAsync(tail + 0, keep.p + count - 1, 4, cudaMemcpyDeviceToHost, stream));
    MTB_TRY(cudaMemcpyAsync(tail + 1, keep_pos.p + count - 1, 4, cudaMemcpyDeviceToHost, stream));
    MTB_TRY(cudaMemcpyAsync(tail + 2, makes_node.p + count - 1, 4, cudaMemcpyDeviceToHost, stream));
    MTB_TRY(cudaMemcpyAsync(tail + 3, node_pos.p + count - 1, 4, cudaMemcpyDeviceToHost, stream));
    MTB_TRY(cudaStreamSynchronize(stream));
    const int new_count = tail[0] + tail[1];
    n_nodes += tail[2] + tail[3];
    if (new_count >= count) return cudaErrorUnknown;
    count = new_count;
    Cluster *t = cur;
    cur = nxt;
    nxt = t;
  }
  // the last cluster is the root
  Cluster root;
  MTB_TRY(cudaMemcpyAsync(&root, cur, sizeof(Cluster), cudaMemcpyDeviceToHost, stream));
  MTB_TRY(cudaStreamSynchronize(stream));
  if (root.id < 0 || n_nodes < 1 || root.id != n_nodes - 1) return cudaErrorUnknown;

  Bvh2Node *d_nodes = nullptr;
  int32_t *d_leaf_ref = nullptr;
  MTB_TRY(cudaMalloc(reinterpret_cast<void **>(&d_nodes), (size_t)n_nodes * sizeof(Bvh2Node)));
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&d_leaf_ref), (size_t)n * sizeof(int32_t));
  if (e != cudaSuccess) {
    cudaFree(d_nodes);
    return e;
  }
  const int nb = (n_nodes + kPlocBlock - 1) / kPlocBlock;
  PlocLeafCounts<<<nb, kPlocBlock, 0, stream>>>(raw.p, n_nodes, counts.p);
  bytes = tmp_bytes;
  e = cub::DeviceScan::ExclusiveSum(tmp.p, bytes, counts.p, first.p, 2 * n_nodes, stream);
  if (e == cudaSuccess) {
    PlocFinalize<<<nb, kPlocBlock, 0, stream>>>(raw.p, n_nodes, first.p, d_nodes, d_leaf_ref);
    e = cudaStreamSynchronize(stream);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(d_nodes);
    cudaFree(d_leaf_ref);
    return e;
  }
  *d_nodes_out = d_nodes;
  *d_leaf_ref_out = d_leaf_ref;
  *n_nodes_out = n_nodes;
  *depth_out = root.height + 1;
  return cudaSuccess;
}

void LaunchGatherLeafSlots(const SlotRec *slots, const int32_t *ref_slot, const int32_t *leaf_ref, int64_t n_positions, SlotRec *gslots,
                           cudaStream_t stream) {
  if (n_positions <= 0) return;
  const long long threads = n_positions * 8;
  GatherLeafSlots<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(slots, ref_slot, leaf_ref, (int)n_positions, gslots);
}

}  // namespace mtb
