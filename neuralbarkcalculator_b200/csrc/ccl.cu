// K5 -- small-region removal by GPU union-find connected-component labelling, fused with the class histogram.
//
// Replaces remove_small_zones (utils.py:135-148: skimage remove_small_holes + remove_small_objects, threshold
// 150, connectivity=2), the --exclude_nodes relabel (models.py:273-276) and the per-class pixel counts behind the
// CSV percentages (models.py:323-332).  Per image, 8-connected:
//   stage A: components of the foreground (mask != 0) smaller than T become background
//   stage B: components of the resulting background smaller than T become foreground
//   out = background ? 0 : (mask == 0 ? 1 : mask)            (filled background islands are always class 1)
// Labelling = label-equivalence union-find on pixel indices: horizontal runs inside a warp get their run start as
// initial label (ballot), vertical / diagonal neighbours are merged with atomicMin unions, one flatten pass,
// then component sizes are counted with warp-aggregated atomics that saturate at T (only "size < T" matters).
#include "common.cuh"

namespace nbc {

// src[p] != 0 <=> pixel p belongs to the set being labelled.  Labels are indices into the whole [N,H,W] buffer
// and always satisfy labels[p] <= p, -1 for pixels outside the set.
// vh (optional): ragged batch, image n has vh[n] valid rows of the H-row canvas; pixels below are outside the image.
// Every kernel runs on a (pixel blocks, image) grid; a block that starts below the last valid row of its image
// exits at once, so the dead part of a ragged canvas costs nothing and is never read or written.
__device__ __forceinline__ int valid_rows(int n, int H, const int* vh) { return vh == nullptr ? H : min(H, __ldg(vh + n)); }

__global__ void __launch_bounds__(256) ccl_init(const uint8_t* __restrict__ src, int H, int W, int* __restrict__ labels,
                                                int* __restrict__ sizes, const int* __restrict__ vh) {
  const int n = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t live = (int64_t)valid_rows(n, H, vh) * W;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x;
  if (i0 >= live) return;   // block-uniform
  const int64_t i = i0 + threadIdx.x;
  const int64_t p = (int64_t)n * HW + i;
  const int lane = threadIdx.x & 31;
  const bool inb = i < live;
  const bool in = inb && src[p] != 0;
  const int x = inb ? (int)(i % W) : 0;
  const unsigned m = __ballot_sync(0xffffffffu, in);
  const unsigned rowstart = __ballot_sync(0xffffffffu, x == 0);
  // link[k] = lanes k and k-1 are both in the set and in the same image row
  const unsigned link = m & (m << 1) & ~rowstart;
  if (inb) {
    int lab = -1;
    if (in) lab = (int)p - __clz(~link << (31 - lane));  // run of set lanes directly below this one
    labels[p] = lab;
    // a root is always a pixel that started as its own label (unions only ever lower the label of a root), so only those
    // counters can be incremented or read later: zeroing them alone saves 4 bytes of writes per pixel
    if (lab == (int)p) sizes[p] = 0;
  }
}

__device__ __forceinline__ int find_root(const int* labels, int a) {
  int l = __ldcg(labels + a);
  while (l != a) {
    a = l;
    l = __ldcg(labels + a);
  }
  return a;
}

__device__ __forceinline__ void unite(int* labels, int a, int b) {
  bool done;
  do {
    a = find_root(labels, a);
    b = find_root(labels, b);
    if (a < b) {
      const int old = atomicMin(labels + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(labels + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// Row-to-row unions.  Pixels of a horizontal run are already equivalent (ccl_init inside a warp, the lane-0 union
// below across warp boundaries), so a union with the row above is only needed ONCE per pair of touching runs: at
// the first pixel where the pair starts to touch.  With l = left, u = up, ul = up-left, ur = up-right:
//   u in set : unite(p, u) unless (l and ul are in the set)   -- then l made the same union (l~p, ul~u by runs)
//   u not    : unite(p, ul) unless l is in the set            -- then l united with ul (its own "up")
//              unite(p, ur) unless r is in the set            -- then r unites with ur (its own "up")
// which turns ~one union per pixel into ~one per touching run pair.
__global__ void __launch_bounds__(256) ccl_merge(const uint8_t* __restrict__ src, int H, int W, int* __restrict__ labels,
                                                 const int* __restrict__ vh) {
  const int n = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t live = (int64_t)valid_rows(n, H, vh) * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= live) return;
  const int64_t p = (int64_t)n * HW + i;
  if (src[p] == 0) return;
  const int lane = threadIdx.x & 31;
  const int x = (int)(i % W);
  const int y = (int)(i / W);
  const int ip = (int)p;
  const bool l = x > 0 && src[p - 1] != 0;
  if (l && lane == 0) unite(labels, ip, ip - 1);  // run cut by the warp boundary
  if (y > 0) {
    const int64_t up = p - W;
    const bool ul = x > 0 && src[up - 1] != 0;
    if (src[up]) {
      if (!(l && ul)) unite(labels, ip, (int)up);
    } else {
      if (ul && !l) unite(labels, ip, (int)up - 1);
      if (x < W - 1 && src[up + 1] != 0 && src[p + 1] == 0) unite(labels, ip, (int)up + 1);
    }
  }
}

__global__ void __launch_bounds__(256) ccl_flatten_count(int H, int W, int threshold, int* __restrict__ labels,
                                                         int* __restrict__ sizes, const int* __restrict__ vh) {
  const int n = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t live = (int64_t)valid_rows(n, H, vh) * W;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x;
  if (i0 >= live) return;   // block-uniform
  const int64_t i = i0 + threadIdx.x;
  const int64_t p = (int64_t)n * HW + i;
  const int lane = threadIdx.x & 31;
  int root = -1;
  if (i < live && labels[p] >= 0) {
    root = find_root(labels, (int)p);
    labels[p] = root;
  }
  const bool in = root >= 0;
  const unsigned active = __ballot_sync(0xffffffffu, in);
  if (in) {
    const unsigned peers = __match_any_sync(active, root);
    if (lane == __ffs(peers) - 1) {
      // saturating count: once a component is known to have >= threshold pixels its exact size is irrelevant
      if (__ldcg(sizes + root) < threshold) atomicAdd(sizes + root, __popc(peers));
    }
  }
}

// stage A result + stage B initialisation in one pass: setB[p] = background after removing small foreground components,
// and at once the run labels of that background set (what ccl_init would compute from setB in another pass over it) in
// the stage-B label / counter arrays.
__global__ void __launch_bounds__(256) ccl_stage_a_apply_init_b(const uint8_t* __restrict__ mask, const int* __restrict__ labels,
                                                                const int* __restrict__ sizes, int threshold,
                                                                uint8_t* __restrict__ setB, int* __restrict__ labelsB,
                                                                int* __restrict__ sizesB, int H, int W,
                                                                const int* __restrict__ vh) {
  const int n = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t live = (int64_t)valid_rows(n, H, vh) * W;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x;
  if (i0 >= live) return;   // block-uniform
  const int64_t i = i0 + threadIdx.x;
  const int64_t p = (int64_t)n * HW + i;
  const int lane = threadIdx.x & 31;
  const bool inb = i < live;
  const bool in = inb && ((mask[p] == 0) || (__ldcg(sizes + labels[p]) < threshold));
  const int x = inb ? (int)(i % W) : 0;
  const unsigned m = __ballot_sync(0xffffffffu, in);
  const unsigned rowstart = __ballot_sync(0xffffffffu, x == 0);
  const unsigned link = m & (m << 1) & ~rowstart;
  if (inb) {
    setB[p] = in ? 1 : 0;
    int lab = -1;
    if (in) lab = (int)p - __clz(~link << (31 - lane));
    labelsB[p] = lab;
    if (lab == (int)p) sizesB[p] = 0;
  }
}

__global__ void __launch_bounds__(256) ccl_final(uint8_t* __restrict__ mask, const uint8_t* __restrict__ setB,
                                                 const int* __restrict__ labels, const int* __restrict__ sizes,
                                                 int threshold, int exclude_nodes, int H, int W,
                                                 int* __restrict__ counts, const int* __restrict__ vh) {
  __shared__ int s_cnt[3];
  const int n = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t live = (int64_t)valid_rows(n, H, vh) * W;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x;
  if (i0 >= live) return;   // block-uniform
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = i0 + threadIdx.x;
  int cls = -1;
  if (i < live) {
    const int64_t p = (int64_t)n * HW + i;
    const uint8_t m = mask[p];
    const bool bg = setB[p] && (__ldcg(sizes + labels[p]) >= threshold);
    cls = bg ? 0 : (m == 0 ? 1 : m);
    if (exclude_nodes && cls == 2) cls = 1;
    if (cls != m) mask[p] = (uint8_t)cls;
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int k = __popc(__ballot_sync(0xffffffffu, cls == c));
    if (lane == 0 && k) atomicAdd(&s_cnt[c], k);
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(counts + n * 3 + threadIdx.x, s_cnt[threadIdx.x]);
}

}  // namespace nbc

using namespace nbc;

extern "C" size_t nbc_ccl_workspace_bytes(int N, int H, int W) {
  const size_t total = (size_t)N * H * W;
  return align_up(total * 4, 256) * 4 + align_up(total, 256);   // labels + counters of both stages, the stage-B set
}

static int remove_small_zones_impl(uint8_t* mask, int N, int H, int W, int threshold, int exclude_nodes, int32_t* counts,
                                   void* workspace, size_t workspace_bytes, const int* vh, cudaStream_t stream) {
  NBC_REQUIRE(mask && counts && workspace, "nbc_remove_small_zones: null pointer");
  NBC_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "nbc_remove_small_zones: bad shape");
  const int64_t total = (int64_t)N * H * W;
  NBC_REQUIRE(total < (1ll << 31), "nbc_remove_small_zones: N*H*W must be < 2^31");
  if (workspace_bytes < nbc_ccl_workspace_bytes(N, H, W)) {
    set_error("nbc_remove_small_zones: workspace %zu < %zu", workspace_bytes, nbc_ccl_workspace_bytes(N, H, W));
    return NBC_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  int* labels = reinterpret_cast<int*>(ws);
  int* sizes = reinterpret_cast<int*>(ws + align_up((size_t)total * 4, 256));
  int* labelsB = reinterpret_cast<int*>(ws + 2 * align_up((size_t)total * 4, 256));
  int* sizesB = reinterpret_cast<int*>(ws + 3 * align_up((size_t)total * 4, 256));
  uint8_t* setB = reinterpret_cast<uint8_t*>(ws + 4 * align_up((size_t)total * 4, 256));
  const int64_t HW = (int64_t)H * W;
  const dim3 grid((unsigned)ceil_div64(HW, 256), N);
  NBC_CUDA(cudaMemsetAsync(counts, 0, (size_t)N * 3 * sizeof(int32_t), stream));
  // stage A: foreground components
  ccl_init<<<grid, 256, 0, stream>>>(mask, H, W, labels, sizes, vh);
  NBC_CHECK_LAUNCH();
  ccl_merge<<<grid, 256, 0, stream>>>(mask, H, W, labels, vh);
  NBC_CHECK_LAUNCH();
  ccl_flatten_count<<<grid, 256, 0, stream>>>(H, W, threshold, labels, sizes, vh);
  NBC_CHECK_LAUNCH();
  // stage B: background components of the stage-A result (its own label / counter arrays: the fused kernel still reads
  // stage A's while it writes them)
  ccl_stage_a_apply_init_b<<<grid, 256, 0, stream>>>(mask, labels, sizes, threshold, setB, labelsB, sizesB, H, W, vh);
  NBC_CHECK_LAUNCH();
  ccl_merge<<<grid, 256, 0, stream>>>(setB, H, W, labelsB, vh);
  NBC_CHECK_LAUNCH();
  ccl_flatten_count<<<grid, 256, 0, stream>>>(H, W, threshold, labelsB, sizesB, vh);
  NBC_CHECK_LAUNCH();
  ccl_final<<<grid, 256, 0, stream>>>(mask, setB, labelsB, sizesB, threshold, exclude_nodes, H, W, counts, vh);
  NBC_CHECK_LAUNCH();
  return 0;
}

extern "C" int nbc_remove_small_zones(uint8_t* mask, int N, int H, int W, int threshold, int exclude_nodes,
                                      int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  return remove_small_zones_impl(mask, N, H, W, threshold, exclude_nodes, counts, workspace, workspace_bytes, nullptr,
                                 reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int nbc_remove_small_zones_ragged(uint8_t* mask, int N, int Hc, int W, const int32_t* heights, int threshold,
                                             int exclude_nodes, int32_t* counts, void* workspace, size_t workspace_bytes,
                                             void* stream) {
  NBC_REQUIRE(heights, "nbc_remove_small_zones_ragged: null heights");
  return remove_small_zones_impl(mask, N, Hc, W, threshold, exclude_nodes, counts, workspace, workspace_bytes, heights,
                                 reinterpret_cast<cudaStream_t>(stream));
}
