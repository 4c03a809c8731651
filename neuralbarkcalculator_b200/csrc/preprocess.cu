// K1 -- 4x cubic resize (4096^2 -> 1024^2) + dark-band trim, all in integer arithmetic on u8.
//
// Replaces Preprocessor._preprocess_image (models.py:191-203: skimage resize order=3 + trim_black + imsave's
// float->u8) and trim_black (models.py:157-166).  At exactly 4x the cubic (Catmull-Rom) sample point sits at
// 4i+1.5, i.e. a separable 4-tap filter [-1, 9, 9, -1]/16 over the pixel's own 4x4 block, so
//     S = sum_{p,q} w_p w_q * in[4i+p, 4j+q]   (exact int, = 256 * resized value)
//     out = clamp((S + 128) >> 8, min(in), max(in))        -- skimage clips to the input range, then u8 rounding
//     nondark(i,j) = sum_c clip(S_c) / (256*255) > 1e-3   <=>  min(in) >= 1  or  sum_c max(S_c, 0) >= 66
// Pass 1 reads the raw image once (coalesced 16-byte loads, 48 B per thread per row), writes (S+128)>>8 as int16,
// and reduces the input min / max and the per-row non-dark counts.  A one-block kernel turns the counts into the
// [first, last) rows (keep a row iff > 85 % of its pixels are non-dark).  Pass 2 clamps and compacts the rows.
// The BMP pixel array (bottom-up, BGR) is consumed directly: flags bit0 = BGR, bit1 = bottom-up.
#include "common.cuh"

namespace nbc {

struct PreHeader {
  int inv_min;  // 255 - min(in), reduced with atomicMax so a zero-filled header is the identity
  int max;
  int pad[2];
};

__device__ __forceinline__ int byte_of(uint32_t v, int i) { return (v >> (8 * i)) & 0xFF; }

template <bool kAligned>
__global__ void __launch_bounds__(256) resize4x_pass1(const uint8_t* __restrict__ raw, int H, int W, int64_t pitch,
                                                      int flags, int16_t* __restrict__ r16, PreHeader* hdr,
                                                      int* __restrict__ rowcount) {
  const int Wo = W >> 2;
  const int groups = (Wo + 3) >> 2;  // 4 output pixels per thread
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int ho = blockIdx.y;
  const bool active = g < groups;
  int nondark = 0;
  uint32_t vmin = 0xFFFFFFFFu, vmax = 0u;
  if (active) {
    const int wo0 = g << 2;
    const int npx = min(4, Wo - wo0);
    const int nbytes = npx * 12;
    int V[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) V[i] = 0;
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      const int r = 4 * ho + pr;
      const int64_t rr = (flags & 2) ? (int64_t)(H - 1 - r) : (int64_t)r;
      const uint8_t* src = raw + rr * pitch + (int64_t)wo0 * 12;
      uint32_t words[12];
      if (kAligned && npx == 4) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const uint4 v = __ldg(s4 + i);
          words[4 * i] = v.x, words[4 * i + 1] = v.y, words[4 * i + 2] = v.z, words[4 * i + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          uint32_t v = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int idx = 4 * i + b;
            // replicate the first byte into the padding so it does not disturb min / max
            const uint32_t byte = (idx < nbytes) ? src[idx] : src[0];
            v |= byte << (8 * b);
          }
          words[i] = v;
        }
      }
      const int wgt = (pr == 0 || pr == 3) ? -1 : 9;
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        vmin = __vminu4(vmin, words[i]);
        vmax = __vmaxu4(vmax, words[i]);
#pragma unroll
        for (int b = 0; b < 4; ++b) V[4 * i + b] += wgt * byte_of(words[i], b);
      }
    }
    // horizontal pass: output pixel px uses input pixels 4px..4px+3 (3 bytes each)
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      if (px < npx) {
        int S[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int b = px * 12 + c;
          S[c] = -V[b] + 9 * V[b + 3] + 9 * V[b + 6] - V[b + 9];
        }
        const int s0 = (flags & 1) ? S[2] : S[0], s2 = (flags & 1) ? S[0] : S[2];  // BGR -> RGB
        int16_t* o = r16 + ((int64_t)ho * Wo + wo0 + px) * 3;
        o[0] = (int16_t)((s0 + 128) >> 8);
        o[1] = (int16_t)((S[1] + 128) >> 8);
        o[2] = (int16_t)((s2 + 128) >> 8);
        nondark += (max(S[0], 0) + max(S[1], 0) + max(S[2], 0)) >= 66 ? 1 : 0;
      }
    }
  }
  // reductions: bytes -> scalar min/max, then warp, then one atomic per warp
  int mn = min(min(byte_of(vmin, 0), byte_of(vmin, 1)), min(byte_of(vmin, 2), byte_of(vmin, 3)));
  int mx = max(max(byte_of(vmax, 0), byte_of(vmax, 1)), max(byte_of(vmax, 2), byte_of(vmax, 3)));
  if (!active) mn = 255, mx = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nondark += __shfl_xor_sync(0xffffffffu, nondark, o);
  }
  // one (min, max, count) per block; the two range atomics all land on the same address, so they are skipped when
  // they cannot change it (the values only grow, a stale read just means one redundant atomic)
  __shared__ int s_mn[8], s_mx[8], s_nd[8];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s_mn[wid] = mn, s_mx[wid] = mx, s_nd[wid] = nondark;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 1; i < nw; ++i) mn = min(mn, s_mn[i]), mx = max(mx, s_mx[i]), nondark += s_nd[i];
    if (255 - mn > __ldcg(&hdr->inv_min)) atomicMax(&hdr->inv_min, 255 - mn);
    if (mx > __ldcg(&hdr->max)) atomicMax(&hdr->max, mx);
    if (nondark) atomicAdd(&rowcount[ho], nondark);
  }
}

// rows of a u8 image that needs no resize: nondark <=> any channel non-zero (sum/255 > 1e-3)
__global__ void __launch_bounds__(256) rowcount_u8(const uint8_t* __restrict__ img, int H, int W,
                                                    int* __restrict__ rowcount) {
  const int h = blockIdx.x;
  int cnt = 0;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const uint8_t* p = img + ((int64_t)h * W + w) * 3;
    cnt += ((int)p[0] + p[1] + p[2]) >= 1 ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&rowcount[h], cnt);
}

// trim_black row rule (models.py:160-164): keep[r] = count/W > 0.85; first = argmax(keep);
// last = H - argmax(keep[::-1]); nothing kept -> (0, H).  all_nondark: min(in) >= 1 makes every pixel non-dark.
__global__ void __launch_bounds__(1024) trim_rows_kernel(const int* __restrict__ rowcount, const PreHeader* hdr,
                                                         int Ho, int Wo, int square, int32_t* first_last) {
  __shared__ int s_first, s_last_from_end;
  if (threadIdx.x == 0) s_first = INT_MAX, s_last_from_end = INT_MAX;
  __syncthreads();
  const bool all_nondark = hdr && (255 - hdr->inv_min) >= 1;
  for (int r = threadIdx.x; r < Ho; r += blockDim.x) {
    const int c = all_nondark ? Wo : rowcount[r];
    const bool keep = ((double)c / (double)Wo) > 0.85;
    if (keep) {
      atomicMin(&s_first, r);
      atomicMin(&s_last_from_end, Ho - 1 - r);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int first = 0, last = Ho;
    if (square && s_first != INT_MAX) {
      first = s_first;
      last = Ho - s_last_from_end;
    }
    first_last[0] = first;
    first_last[1] = last;
  }
}

__global__ void __launch_bounds__(256) resize4x_pass2(const int16_t* __restrict__ r16, const PreHeader* hdr, int Wo,
                                                      const int32_t* __restrict__ first_last,
                                                      uint8_t* __restrict__ out) {
  const int first = first_last[0], last = first_last[1];
  const int lo = 255 - hdr->inv_min, hi = hdr->max;
  const int64_t total = (int64_t)(last - first) * Wo * 3;
  const int16_t* src = r16 + (int64_t)first * Wo * 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < total; i += stride) {
    if (i + 8 <= total && ((reinterpret_cast<uintptr_t>(src + i) & 15) == 0) &&
        ((reinterpret_cast<uintptr_t>(out + i) & 7) == 0)) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[2] = {0, 0};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int s = (int)(int16_t)((w[k >> 1] >> ((k & 1) * 16)) & 0xFFFF);
        o[k >> 2] |= (uint32_t)min(max(s, lo), hi) << ((k & 3) * 8);
      }
      *reinterpret_cast<uint2*>(out + i) = make_uint2(o[0], o[1]);
    } else {
      for (int64_t j = i; j < min(i + 8, total); ++j) out[j] = (uint8_t)min(max((int)src[j], lo), hi);
    }
  }
}

__global__ void __launch_bounds__(256) copy_rows_u8(const uint8_t* __restrict__ img, int W,
                                                    const int32_t* __restrict__ first_last, uint8_t* __restrict__ out) {
  const int first = first_last[0], last = first_last[1];
  const int64_t total = (int64_t)(last - first) * W * 3;
  const uint8_t* src = img + (int64_t)first * W * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = src[i];
}

static size_t pre_header_bytes(int Ho) { return align_up(sizeof(PreHeader) + (size_t)Ho * sizeof(int), 256); }

}  // namespace nbc

using namespace nbc;

extern "C" size_t nbc_preprocess_workspace_bytes(int H, int W) {
  const int Ho = H / 4, Wo = W / 4;
  return pre_header_bytes(Ho > H ? Ho : H) + align_up((size_t)Ho * Wo * 3 * sizeof(int16_t), 256);
}

extern "C" int nbc_preprocess_4x_u8(const uint8_t* raw, int H, int W, int64_t pitch, int flags, uint8_t* out,
                                    int32_t* first_last, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NBC_REQUIRE(raw && out && first_last && workspace, "nbc_preprocess_4x_u8: null pointer");
  NBC_REQUIRE(H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, "nbc_preprocess_4x_u8: H and W must be multiples of 4 (got %dx%d)", H, W);
  NBC_REQUIRE(pitch >= (int64_t)W * 3, "nbc_preprocess_4x_u8: pitch %lld < 3*W", (long long)pitch);
  if (workspace_bytes < nbc_preprocess_workspace_bytes(H, W)) {
    set_error("nbc_preprocess_4x_u8: workspace %zu < %zu", workspace_bytes, nbc_preprocess_workspace_bytes(H, W));
    return NBC_ERR_WORKSPACE;
  }
  const int Ho = H / 4, Wo = W / 4;
  char* ws = reinterpret_cast<char*>(workspace);
  PreHeader* hdr = reinterpret_cast<PreHeader*>(ws);
  int* rowcount = reinterpret_cast<int*>(ws + sizeof(PreHeader));
  int16_t* r16 = reinterpret_cast<int16_t*>(ws + pre_header_bytes(H));
  NBC_CUDA(cudaMemsetAsync(ws, 0, sizeof(PreHeader) + (size_t)Ho * sizeof(int), stream));
  const int groups = (Wo + 3) / 4;
  dim3 grid(ceil_div(groups, 256), Ho);
  const bool aligned = (reinterpret_cast<uintptr_t>(raw) % 16 == 0) && (pitch % 16 == 0);
  if (aligned)
    resize4x_pass1<true><<<grid, 256, 0, stream>>>(raw, H, W, pitch, flags, r16, hdr, rowcount);
  else
    resize4x_pass1<false><<<grid, 256, 0, stream>>>(raw, H, W, pitch, flags, r16, hdr, rowcount);
  NBC_CHECK_LAUNCH();
  trim_rows_kernel<<<1, 1024, 0, stream>>>(rowcount, hdr, Ho, Wo, Ho == Wo ? 1 : 0, first_last);
  NBC_CHECK_LAUNCH();
  const int64_t total8 = ceil_div64((int64_t)Ho * Wo * 3, 8);
  const int blocks = (int)(total8 < 256 ? 1 : (ceil_div64(total8, 256) < 4 * 148 ? ceil_div64(total8, 256) : 4 * 148));
  resize4x_pass2<<<blocks, 256, 0, stream>>>(r16, hdr, Wo, first_last, out);
  NBC_CHECK_LAUNCH();
  return 0;
}

extern "C" int nbc_trim_u8(const uint8_t* img, int H, int W, uint8_t* out, int32_t* first_last, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NBC_REQUIRE(img && out && first_last && workspace, "nbc_trim_u8: null pointer");
  NBC_REQUIRE(H > 0 && W > 0, "nbc_trim_u8: bad shape");
  if (workspace_bytes < pre_header_bytes(H)) {
    set_error("nbc_trim_u8: workspace too small");
    return NBC_ERR_WORKSPACE;
  }
  char* ws = reinterpret_cast<char*>(workspace);
  int* rowcount = reinterpret_cast<int*>(ws + sizeof(PreHeader));
  NBC_CUDA(cudaMemsetAsync(ws, 0, sizeof(PreHeader) + (size_t)H * sizeof(int), stream));
  rowcount_u8<<<H, 256, 0, stream>>>(img, H, W, rowcount);
  NBC_CHECK_LAUNCH();
  trim_rows_kernel<<<1, 1024, 0, stream>>>(rowcount, nullptr, H, W, H == W ? 1 : 0, first_last);
  NBC_CHECK_LAUNCH();
  copy_rows_u8<<<4 * 148, 256, 0, stream>>>(img, W, first_last, out);
  NBC_CHECK_LAUNCH();
  return 0;
}
