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

// ------------------------------------------------------------------------------------------------------------
// Run-based variant (the default for W <= 1024, i.e. every processed image): the same two-stage labelling on BITMASKS.
// A row is 32-bit words (bit i of word L = pixel 32 L + i is in the set), one word per lane, one warp per row; the units
// of the union-find are whole horizontal RUNS -- labels / counters live at the pixel index of a run's first pixel, so
// find_root / unite above are used unchanged -- and every pass touches 32 pixels per lane with a handful of bit operations
// instead of one thread per pixel: the per-pixel kernels move ~50 bytes and several hundred instructions per pixel, this
// one reads the mask once per stage (1 byte per pixel) and works on ~1/30 as many elements.
//   rows_bits_init : mask bytes -> set words, labels[start] = start for every run start
//   rows_merge     : the unions between runs of neighbouring rows.  Same rule as ccl_merge, as word masks: with C / U the
//                    words of this row / the row above and L1, R1, UL, UR their shifts by one pixel,
//                      E1 = C & U & ~(L1 & UL)   -> unite(run of x, run of up x)         (first pixel where the pair touches)
//                      E2 = C & ~U & UL & ~L1    -> unite(run of x, run of up x - 1)
//                      E3 = C & ~U & UR & ~R1    -> unite(run of x, run starting at up x + 1)
//   rows_flatten   : per run segment (a run cut at word boundaries) root lookup + saturating size count, path compression
//   rows_apply_a   : stage A result -> stage B set words (+ their run starts)
//   rows_final     : stage B result -> mask bytes (only the pixels that change are rewritten) + class counts
// ------------------------------------------------------------------------------------------------------------
constexpr unsigned kFull = 0xffffffffu;
constexpr int kRunRows = 8;      // rows (warps) per block

// 4 bytes -> 4 bits: bit k = (byte k != 0)
__device__ __forceinline__ uint32_t nz_nibble(uint32_t v) {
  const uint32_t m = (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
  return ((m >> 7) * 0x01020408u) >> 24;
}
// 4 bits -> 4 byte masks (0xFF where the bit is set)
__device__ __forceinline__ uint32_t nibble_bytes(uint32_t nib) { return (((nib & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu; }

struct RowBytes {
  uint32_t v[8];     // the lane's 32 mask bytes (aligned path only)
  uint32_t nz, one, two;
  bool vec;
};
// this lane's 32 pixels of a mask row: words of (byte != 0), (byte == 1), (byte == 2); bits at or beyond W are 0
__device__ __forceinline__ RowBytes load_row_bytes(const uint8_t* __restrict__ row, int W, int lane) {
  RowBytes r;
  r.nz = r.one = r.two = 0u;
  r.vec = false;
  const int x0 = lane * 32;
  if (x0 >= W) return r;
  if (x0 + 32 <= W && (reinterpret_cast<uintptr_t>(row + x0) & 15) == 0) {
    r.vec = true;
    const uint4 a = *reinterpret_cast<const uint4*>(row + x0), b = *reinterpret_cast<const uint4*>(row + x0 + 16);
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r.nz |= nz_nibble(r.v[j]) << (4 * j);
      r.one |= nz_nibble(__vcmpeq4(r.v[j], 0x01010101u)) << (4 * j);
      r.two |= nz_nibble(__vcmpeq4(r.v[j], 0x02020202u)) << (4 * j);
    }
  } else {
    for (int i = 0; i < 32 && x0 + i < W; ++i) {
      const uint32_t c = row[x0 + i];
      r.nz |= (uint32_t)(c != 0) << i, r.one |= (uint32_t)(c == 1) << i, r.two |= (uint32_t)(c == 2) << i;
    }
  }
  return r;
}
__device__ __forceinline__ uint32_t valid_word(int W, int lane) {
  const int x0 = lane * 32;
  return x0 >= W ? 0u : (x0 + 32 <= W ? kFull : ((1u << (W - x0)) - 1u));
}

// Warp-collective: does the run covering bit 0 of this lane's word continue from the lane below, and where does it start
// (x position in the row)?  Every lane of the warp must call this.
struct RowCarry {
  bool enter;
  int enter_start;
};
__device__ __forceinline__ RowCarry row_carry(uint32_t word, int lane) {
  RowCarry c;
  const uint32_t prev_msb = __shfl_up_sync(kFull, word >> 31, 1);
  c.enter = lane > 0 && (word & 1u) && prev_msb;
  const uint32_t full = __ballot_sync(kFull, word == kFull);
  const uint32_t lower = ~full & ((1u << lane) - 1u);      // lanes below this one whose word has a zero bit
  const int lp = lower ? 31 - __clz(lower) : 0;
  const uint32_t wp = __shfl_sync(kFull, word, lp);
  // the entering run began right after the highest zero bit of the nearest lower word that has one (x = 0 if none has)
  c.enter_start = lower ? lp * 32 + (32 - __clz(~wp)) : 0;
  return c;
}
// x position of the first pixel of the run that contains bit i of this lane's word (bit i must be set)
__device__ __forceinline__ int run_start(uint32_t word, int lane, int i, const RowCarry& c) {
  const uint32_t z = ~word & ((1u << i) - 1u);
  if (z) return lane * 32 + (32 - __clz(z));
  return c.enter ? c.enter_start : lane * 32;
}
// first maximal segment of ones in `rem` (non-zero): start bit and mask
__device__ __forceinline__ uint32_t first_segment(uint32_t rem, int& i, int& len) {
  i = __ffs(rem) - 1;
  const uint32_t t = ~(rem >> i);
  len = t ? __ffs(t) - 1 : 32;
  return len == 32 ? kFull : (((1u << len) - 1u) << i);
}

__device__ __forceinline__ bool run_row(int H, const int* vh, int& n, int& y) {
  n = blockIdx.y;
  y = blockIdx.x * kRunRows + (threadIdx.x >> 5);
  return y < valid_rows(n, H, vh);      // warp-uniform
}

// stage A set = (mask != 0)
__global__ void __launch_bounds__(32 * kRunRows) rows_bits_init(const uint8_t* __restrict__ mask, int H, int W,
                                                                uint32_t* __restrict__ bits, int* __restrict__ labels,
                                                                int* __restrict__ sizes, const int* __restrict__ vh) {
  int n, y;
  if (!run_row(H, vh, n, y)) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)n * H + y;
  const uint32_t word = load_row_bytes(mask + row * W, W, lane).nz;
  bits[row * 32 + lane] = word;
  const uint32_t carry = __shfl_up_sync(kFull, word >> 31, 1);
  uint32_t starts = word & ~((word << 1) | (lane > 0 ? carry : 0u));
  const int base = (int)(row * W) + lane * 32;
  while (starts) {
    const int i = __ffs(starts) - 1;
    starts &= starts - 1;
    labels[base + i] = base + i;
    sizes[base + i] = 0;
  }
}

__global__ void __launch_bounds__(32 * kRunRows) rows_merge(const uint32_t* __restrict__ bits, int H, int W, int* __restrict__ labels,
                                                            const int* __restrict__ vh) {
  int n, y;
  if (!run_row(H, vh, n, y) || y == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)n * H + y;
  const uint32_t C = bits[row * 32 + lane], U = bits[(row - 1) * 32 + lane];
  // neighbour bits across the word boundaries (every lane executes every shuffle; the row ends contribute zeros)
  uint32_t c_prev = __shfl_up_sync(kFull, C >> 31, 1), u_prev = __shfl_up_sync(kFull, U >> 31, 1);
  uint32_t c_next = __shfl_down_sync(kFull, C & 1u, 1), u_next = __shfl_down_sync(kFull, U & 1u, 1);
  if (lane == 0) c_prev = u_prev = 0u;
  if (lane == 31) c_next = u_next = 0u;
  const uint32_t L1 = (C << 1) | c_prev, R1 = (C >> 1) | (c_next << 31);
  const uint32_t UL = (U << 1) | u_prev, UR = (U >> 1) | (u_next << 31);
  const uint32_t E1 = C & U & ~(L1 & UL), E2 = C & ~U & UL & ~L1, E3 = C & ~U & UR & ~R1;
  const RowCarry cc = row_carry(C, lane), cu = row_carry(U, lane);
  // start of the run that contains bit 31 of the lane below, in the row above (target of an E2 event at bit 0)
  const int s31 = (U >> 31) ? run_start(U, lane, 31, cu) : 0;
  const int s31_below = __shfl_up_sync(kFull, s31, 1);
  uint32_t ev = E1 | E2 | E3;
  const int cur0 = (int)(row * W), up0 = (int)((row - 1) * W);
  while (ev) {
    const int i = __ffs(ev) - 1;
    ev &= ev - 1;
    const int a = cur0 + run_start(C, lane, i, cc);
    if ((E1 >> i) & 1u) {
      unite(labels, a, up0 + run_start(U, lane, i, cu));
    } else {
      if ((E2 >> i) & 1u) unite(labels, a, up0 + (i > 0 ? run_start(U, lane, i - 1, cu) : s31_below));
      if ((E3 >> i) & 1u) unite(labels, a, up0 + lane * 32 + i + 1);      // up x is outside the set: that run starts at x + 1
    }
  }
}

__global__ void __launch_bounds__(32 * kRunRows) rows_flatten(const uint32_t* __restrict__ bits, int H, int W, int threshold,
                                                              int* __restrict__ labels, int* __restrict__ sizes,
                                                              const int* __restrict__ vh) {
  int n, y;
  if (!run_row(H, vh, n, y)) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)n * H + y;
  const uint32_t word = bits[row * 32 + lane];
  const RowCarry c = row_carry(word, lane);
  const int row0 = (int)(row * W);
  uint32_t rem = word;
  while (rem) {
    int i, len;
    rem &= ~first_segment(rem, i, len);
    const bool starts_here = !(i == 0 && c.enter);
    const int idx = row0 + (starts_here ? lane * 32 + i : c.enter_start);
    const int root = find_root(labels, idx);
    if (starts_here) labels[idx] = root;
    // saturating count: once a component is known to have >= threshold pixels its exact size is irrelevant
    if (__ldcg(sizes + root) < threshold) atomicAdd(sizes + root, len);
  }
}

// "small" bits of a row: the pixels of set segments whose component has fewer than `threshold` pixels
__device__ __forceinline__ uint32_t small_bits(uint32_t word, int lane, const RowCarry& c, int row0, const int* __restrict__ labels,
                                               const int* __restrict__ sizes, int threshold) {
  uint32_t rem = word, small = 0u;
  while (rem) {
    int i, len;
    const uint32_t seg = first_segment(rem, i, len);
    rem &= ~seg;
    const int idx = row0 + ((i == 0 && c.enter) ? c.enter_start : lane * 32 + i);
    if (__ldcg(sizes + __ldcg(labels + idx)) < threshold) small |= seg;      // labels[run start] is the root after rows_flatten
  }
  return small;
}

// stage A result -> stage B set: background or a small foreground component; + the run starts of that set
__global__ void __launch_bounds__(32 * kRunRows) rows_apply_a(const uint32_t* __restrict__ bitsA, const int* __restrict__ labelsA,
                                                              const int* __restrict__ sizesA, int threshold, int H, int W,
                                                              uint32_t* __restrict__ bitsB, int* __restrict__ labelsB,
                                                              int* __restrict__ sizesB, const int* __restrict__ vh) {
  int n, y;
  if (!run_row(H, vh, n, y)) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)n * H + y;
  const uint32_t F = bitsA[row * 32 + lane];
  const RowCarry c = row_carry(F, lane);
  const int row0 = (int)(row * W);
  const uint32_t B = (~F | small_bits(F, lane, c, row0, labelsA, sizesA, threshold)) & valid_word(W, lane);
  bitsB[row * 32 + lane] = B;
  const uint32_t carry = __shfl_up_sync(kFull, B >> 31, 1);
  uint32_t starts = B & ~((B << 1) | (lane > 0 ? carry : 0u));
  const int base = row0 + lane * 32;
  while (starts) {
    const int i = __ffs(starts) - 1;
    starts &= starts - 1;
    labelsB[base + i] = base + i;
    sizesB[base + i] = 0;
  }
}

// out = background ? 0 : (mask == 0 ? 1 : mask), background = stage-B set pixel whose component is not small; class 2 -> 1 with
// exclude_nodes; counts per class.  Only pixels whose value changes are rewritten.
__global__ void __launch_bounds__(32 * kRunRows) rows_final(uint8_t* __restrict__ mask, const uint32_t* __restrict__ bitsB,
                                                            const int* __restrict__ labelsB, const int* __restrict__ sizesB,
                                                            int threshold, int exclude_nodes, int H, int W,
                                                            int* __restrict__ counts, const int* __restrict__ vh) {
  __shared__ int s_cnt[3];
  int n, y;
  n = blockIdx.y;
  if (blockIdx.x * kRunRows >= valid_rows(n, H, vh)) return;      // block-uniform
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (run_row(H, vh, n, y)) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)n * H + y;
    const uint32_t B = bitsB[row * 32 + lane];
    const RowCarry c = row_carry(B, lane);
    const int row0 = (int)(row * W);
    const uint32_t valid = valid_word(W, lane);
    const uint32_t bg = B & ~small_bits(B, lane, c, row0, labelsB, sizesB, threshold);
    uint8_t* rowp = mask + row * W;
    RowBytes r = load_row_bytes(rowp, W, lane);
    const uint32_t to0 = bg & r.nz;                                              // foreground pixel that ends as background
    const uint32_t to1 = ~bg & valid & (~r.nz | (exclude_nodes ? r.two : 0u));    // filled island, or node -> bark
    if (to0 | to1) {
      if (r.vec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t m0 = nibble_bytes(to0 >> (4 * j)), m1 = nibble_bytes(to1 >> (4 * j));
          r.v[j] = (r.v[j] & ~(m0 | m1)) | (m1 & 0x01010101u);
        }
        uint4* dst = reinterpret_cast<uint4*>(rowp + lane * 32);
        dst[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        dst[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
      } else {
        uint32_t ch = to0 | to1;
        while (ch) {
          const int i = __ffs(ch) - 1;
          ch &= ch - 1;
          rowp[lane * 32 + i] = (uint8_t)((to1 >> i) & 1u);
        }
      }
    }
    int c0 = __popc(bg & valid);
    int c2 = exclude_nodes ? 0 : __popc(r.two & ~bg);
    int c1 = __popc(~bg & valid & (~r.nz | r.one | (exclude_nodes ? r.two : 0u)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(kFull, c0, o), c1 += __shfl_xor_sync(kFull, c1, o), c2 += __shfl_xor_sync(kFull, c2, o);
    }
    if (lane == 0) {
      if (c0) atomicAdd(&s_cnt[0], c0);
      if (c1) atomicAdd(&s_cnt[1], c1);
      if (c2) atomicAdd(&s_cnt[2], c2);
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(counts + n * 3 + threadIdx.x, s_cnt[threadIdx.x]);
}

}  // namespace nbc

using namespace nbc;

extern "C" size_t nbc_ccl_workspace_bytes(int N, int H, int W) {
  const size_t total = (size_t)N * H * W;
  // labels + counters of both stages, the stage-B set (pixel kernels) or the two stages' set words (run kernels: 32 words per row)
  const size_t words = align_up((size_t)N * H * 32 * 4, 256) * 2;
  return align_up(total * 4, 256) * 4 + (align_up(total, 256) > words ? align_up(total, 256) : words);
}

// NBC_CCL_RUNS=0 selects the per-pixel kernels (A/B measurements, cross-check); rows wider than 1024 pixels always use them
static int ccl_runs_default() {
  static const int v = [] {
    const char* e = getenv("NBC_CCL_RUNS");
    return (e && *e) ? atoi(e) : 1;
  }();
  return v;
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
  if (W <= 1024 && ccl_runs_default()) {
    uint32_t* bitsA = reinterpret_cast<uint32_t*>(setB);
    uint32_t* bitsB = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(setB) + align_up((size_t)N * H * 32 * 4, 256));
    const dim3 rg((unsigned)ceil_div(H, kRunRows), N);
    const int rt = 32 * kRunRows;
    rows_bits_init<<<rg, rt, 0, stream>>>(mask, H, W, bitsA, labels, sizes, vh);
    NBC_CHECK_LAUNCH();
    rows_merge<<<rg, rt, 0, stream>>>(bitsA, H, W, labels, vh);
    NBC_CHECK_LAUNCH();
    rows_flatten<<<rg, rt, 0, stream>>>(bitsA, H, W, threshold, labels, sizes, vh);
    NBC_CHECK_LAUNCH();
    rows_apply_a<<<rg, rt, 0, stream>>>(bitsA, labels, sizes, threshold, H, W, bitsB, labelsB, sizesB, vh);
    NBC_CHECK_LAUNCH();
    rows_merge<<<rg, rt, 0, stream>>>(bitsB, H, W, labelsB, vh);
    NBC_CHECK_LAUNCH();
    rows_flatten<<<rg, rt, 0, stream>>>(bitsB, H, W, threshold, labelsB, sizesB, vh);
    NBC_CHECK_LAUNCH();
    rows_final<<<rg, rt, 0, stream>>>(mask, bitsB, labelsB, sizesB, threshold, exclude_nodes, H, W, counts, vh);
    NBC_CHECK_LAUNCH();
    return 0;
  }
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
