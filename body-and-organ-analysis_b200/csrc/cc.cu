// Connected-component post-processing of the body-composition label maps on the device.
//
// Replaces the skimage / OpenCV calls of
//   postprocess_region_segmentation  (_external/body_composition_analysis/body_regions/postprocess.py:8-40:
//                                     skimage.measure.label + regionprops, every component but the largest -> 255)
//   postprocess_part_segmentation    (_external/body_composition_analysis/body_parts/postprocess.py:7-60: slice-wise
//                                     cv2.findContours(RETR_EXTERNAL) + drawContours(FILLED), then
//                                     skimage.morphology.remove_small_objects on the objects and on the holes)
// which the reference runs between the networks and the tissue rules (infer/infer.py:81-89).
//
// One primitive: union-find labelling (label = smallest linear index of the component, so the component that a raster
// scan meets first has the smallest label - the order skimage.measure.label numbers them in, which decides ties of
// "largest").  Connectivity: 26 neighbours in 3-D (skimage's default `connectivity = ndim`, and `connectivity=3` of
// remove_small_objects) or 4 neighbours inside each z-slice (the background of an 8-connected OpenCV contour).
// Sizes can be weighted per slice: the label maps of the 5 mm networks are replicated along z when they go back to the
// input grid (order-0 zoom), which maps components one to one, so the post-processing runs on the 5 mm grid with
// weight[z] = number of output slices that source slice z becomes - the voxel counts are those of the full-size volume.
#include <string.h>
#include "common.cuh"

namespace boa {

// Root of v with path halving: every visited node is re-pointed at its grandparent.  Pointers only ever move to an
// ancestor (trees merge by hanging the larger root under the smaller), so the plain stores are safe next to the
// concurrent atomicMin of cc_union - at worst a node keeps a slightly older ancestor.
// (no const / __restrict__ on L: the tree is updated concurrently, its loads must not take the read-only path)
__device__ __forceinline__ int cc_find(int* L, int v) {
  int r = v;
  while (true) {
    const int p = L[r];
    if (p == r) return r;
    const int gp = L[p];
    if (gp == p) return p;
    L[r] = gp;
    r = gp;
  }
}

__device__ __forceinline__ void cc_union(int* L, int a, int b) {
  while (true) {
    a = cc_find(L, a);
    b = cc_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);  // hang the larger root under the smaller one
    if (old == b) return;
    b = old;                              // somebody else re-rooted b meanwhile: retry from there
  }
}

// The voxel set being labelled: {v : set.in[seg[v]] != 0}, complemented when invert != 0.
struct CcLabelSet {
  uint8_t in[256];
};

// Initial forest: every voxel of the set points at the first voxel of its x-run inside the 32-voxel segment its warp
// covers (ballot + bit scan), so the links along x - the bulk of all links in a solid region - never go through the
// union-find; runs that continue across a segment boundary are joined by one union in the merge pass.
__global__ void __launch_bounds__(256)
cc_init_kernel(const uint8_t* __restrict__ seg, size_t n, int W, CcLabelSet set, int invert, int* __restrict__ L) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_round = (n + 31) / 32 * 32;
  const unsigned lane = threadIdx.x & 31u;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_round; v += stride) {
    const bool in = v < n && ((set.in[seg[v]] != 0) != (invert != 0));
    const bool row_start = v < n && (v % (size_t)W) == 0;
    const unsigned bits = __ballot_sync(0xffffffffu, in);
    const unsigned breaks = __ballot_sync(0xffffffffu, row_start);
    if (v >= n) continue;
    if (!in) { L[v] = -1; continue; }
    const unsigned below = (1u << lane) - 1u;
    const unsigned zeros = ~bits & below;                      // lanes below me that are outside the set
    const unsigned stops = breaks & (below | (1u << lane));    // row starts at or below me
    const int after_zero = zeros ? 32 - __clz(zeros) : 0;      // first lane after the highest such gap
    const int at_break = stops ? 31 - __clz(stops) : 0;
    const int start = after_zero > at_break ? after_zero : at_break;
    L[v] = (int)(v - (lane - (unsigned)start));
  }
}

// Every voxel of the set joins the neighbours that precede it in raster order (half of the neighbourhood), except
// where the link is implied by the links of its left neighbour v-1 (same run):
//   row centre c in the set:      v ~ c   is implied when v-1 and c-1 are in the set  (v-1 ~ c-1, runs along x)
//   centre outside, c-1 in set:   v ~ c-1 is implied when v-1 is in the set           (c-1 is the centre of v-1)
//   centre outside, c+1 in set:   always joined
// (centre in the set: its own run links make the diagonals redundant).  Joins therefore happen where runs start to
// overlap - on the surface of a region, not in its volume.
__global__ void __launch_bounds__(256)
cc_merge_kernel(int* L, int D, int H, int W, int mode /*0: 26-conn 3-D, 1: 4-conn per slice*/) {
  const size_t n = (size_t)D * H * W;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (L[i] < 0) continue;
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((size_t)W * H));
    const int v = (int)i;
    const bool left_in = x > 0 && L[i - 1] >= 0;
    // a run that crosses the boundary of its warp's 32-voxel segment
    if (left_in && (i & 31) == 0) cc_union(L, v, v - 1);
    auto row = [&](size_t r, bool diagonals) {
      const bool centre = L[r] >= 0;
      const bool left = x > 0 && L[r - 1] >= 0;
      if (centre) {
        if (!(left_in && left)) cc_union(L, v, (int)r);
      } else if (diagonals) {
        if (left && !left_in) cc_union(L, v, (int)r - 1);
        if (x + 1 < W && L[r + 1] >= 0) cc_union(L, v, (int)r + 1);
      }
    };
    if (y > 0) row(i - W, mode == 0);
    if (mode == 0 && z > 0) {
      const size_t pz = i - (size_t)W * H;
      if (y > 0) row(pz - W, true);
      row(pz, true);
      if (y + 1 < H) row(pz + W, true);
    }
  }
}

// Final pass: every node points straight at its root.  The walk is read-only (no halving here: a late halving store of
// another thread could replace a node's final root pointer by an intermediate ancestor); the only store to L[v] is its
// owner's.
__global__ void __launch_bounds__(256) cc_compress_kernel(int* L, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
    int r = L[v];
    if (r < 0) continue;
    while (true) {
      const int p = L[r];
      if (p == r) break;
      r = p;
    }
    L[v] = r;
  }
}

// sizes[root] += weight[z] for every voxel of the set.  A solid region is ONE root for tens of millions of voxels, so
// the adds are combined before they reach memory: a thread keeps a running sum while consecutive voxels of its
// grid-stride walk share the root, and what is left at the end is combined across the warp (one atomic per distinct
// root per warp).  border != nullptr: border[root] = 1 when a voxel of the component lies on the edge of its slice.
__global__ void __launch_bounds__(256)
cc_sizes_kernel(const int* __restrict__ L, int D, int H, int W, const int* __restrict__ weight, int* __restrict__ sizes,
                int* __restrict__ border) {
  const size_t n = (size_t)D * H * W;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int cur = -1, sum = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int root = L[i];
    if (root < 0) continue;
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((size_t)W * H));
    if (border && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) border[root] = 1;
    const int w = weight ? weight[z] : 1;
    if (root != cur) {
      if (cur >= 0) atomicAdd(&sizes[cur], sum);
      cur = root;
      sum = 0;
    }
    sum += w;
  }
  // every thread arrives here: combine the leftovers of the warp per root
  const unsigned peers = __match_any_sync(0xffffffffu, cur);
  if (cur >= 0) {
    int total = 0;
    for (unsigned m = peers; m; m &= m - 1) total += __shfl_sync(peers, sum, __ffs(m) - 1);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sizes[cur], total);
  }
}

// best = max over roots of (size << 32 | ~root): the largest component, the first one in raster order on ties
__global__ void __launch_bounds__(256)
cc_largest_kernel(const int* __restrict__ L, const int* __restrict__ sizes, size_t n, unsigned long long* __restrict__ best,
                  int* __restrict__ n_components) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
    if (L[v] == (int)v) {
      atomicMax(best, ((unsigned long long)(unsigned)sizes[v] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)v));
      atomicAdd(n_components, 1);
    }
}

// op 0: every voxel of the set outside the largest component -> seg = fill_value           (regions, :13-16)
// op 1: every voxel of the set whose component has size <= threshold -> seg = fill_value   (remove_small_objects)
// op 2: every voxel of the set whose component does not touch its slice border -> fill     (contour fill)
__global__ void __launch_bounds__(256)
cc_apply_kernel(const int* __restrict__ L, const int* __restrict__ sizes, const int* __restrict__ border,
                const unsigned long long* __restrict__ best, size_t n, int op, int threshold, int fill_value,
                uint8_t* __restrict__ seg) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int best_root = -1;
  if (op == 0) best_root = (int)(0xFFFFFFFFu - (unsigned)(*best & 0xFFFFFFFFull));
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
    const int r = L[v];
    if (r < 0) continue;
    bool hit;
    if (op == 0) hit = r != best_root;
    else if (op == 1) hit = sizes[r] <= threshold;
    else hit = border[r] == 0;
    if (hit) seg[v] = (uint8_t)fill_value;
  }
}

// out[v] = label where mask[v] != 0 (body_parts/postprocess.py:50: `out[filled] = label`)
__global__ void __launch_bounds__(256)
cc_paint_kernel(const uint8_t* __restrict__ mask, size_t n, int label, uint8_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
    if (mask[v]) out[v] = (uint8_t)label;
}

}  // namespace boa

using namespace boa;

// Label the components of {v : h_label_set[seg[v]] != 0} (256 host bytes), complemented when invert != 0, and apply
// `op` to seg in place (see cc_apply_kernel).  d_labels / d_sizes / d_border: int32 [V] scratch each (d_border
// only for op 2); d_best: 16 bytes of scratch.  d_slice_weight: int32 [D] or null (all 1).
extern "C" int boa_cc_filter(uint8_t* d_seg, const int32_t* shape, const uint8_t* h_label_set, int invert, int mode, int op,
                             int threshold, int fill_value, const int32_t* d_slice_weight, int32_t* d_labels,
                             int32_t* d_sizes, int32_t* d_border, void* d_best, void* stream) {
  BOA_REQUIRE(d_seg && shape && h_label_set && d_labels && d_sizes && d_best, "boa_cc_filter: null pointer");
  CcLabelSet set;
  memcpy(set.in, h_label_set, 256);
  BOA_REQUIRE(mode == 0 || mode == 1, "boa_cc_filter: mode must be 0 (26-connected) or 1 (4-connected per slice)");
  BOA_REQUIRE(op >= 0 && op <= 2 && (op != 2 || d_border), "boa_cc_filter: bad op %d", op);
  const int D = shape[0], H = shape[1], W = shape[2];
  BOA_REQUIRE(D > 0 && H > 0 && W > 0, "boa_cc_filter: bad shape");
  const size_t n = (size_t)D * H * W;
  BOA_REQUIRE(n < ((size_t)1 << 31), "boa_cc_filter: more than 2^31 voxels");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, 256);
  cc_init_kernel<<<grid, 256, 0, s>>>(d_seg, n, W, set, invert, d_labels);
  BOA_CHECK_LAUNCH();
  cc_merge_kernel<<<grid, 256, 0, s>>>(d_labels, D, H, W, mode);
  BOA_CHECK_LAUNCH();
  cc_compress_kernel<<<grid, 256, 0, s>>>(d_labels, n);
  BOA_CHECK_LAUNCH();
  BOA_CUDA(cudaMemsetAsync(d_sizes, 0, n * sizeof(int32_t), s));
  if (op == 2) BOA_CUDA(cudaMemsetAsync(d_border, 0, n * sizeof(int32_t), s));
  BOA_CUDA(cudaMemsetAsync(d_best, 0, 16, s));
  cc_sizes_kernel<<<grid, 256, 0, s>>>(d_labels, D, H, W, d_slice_weight, d_sizes, op == 2 ? d_border : nullptr);
  BOA_CHECK_LAUNCH();
  if (op == 0) {
    cc_largest_kernel<<<grid, 256, 0, s>>>(d_labels, d_sizes, n, static_cast<unsigned long long*>(d_best),
                                           reinterpret_cast<int*>(static_cast<unsigned long long*>(d_best) + 1));
    BOA_CHECK_LAUNCH();
  }
  cc_apply_kernel<<<grid, 256, 0, s>>>(d_labels, d_sizes, d_border, static_cast<const unsigned long long*>(d_best), n, op,
                                       threshold, fill_value, d_seg);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}

extern "C" int boa_paint_label(const uint8_t* d_mask, size_t n, int label, uint8_t* d_out, void* stream) {
  BOA_REQUIRE(d_mask && d_out, "boa_paint_label: null pointer");
  cc_paint_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_mask, n, label, d_out);
  BOA_CHECK_LAUNCH();
  return BOA_OK;
}
