// Head tail + K3.
//  * nbc_head_1x1: Dropout (identity in eval) + Conv2d(512, 3, 1) with bias (models.py:113-124) on the bf16
//    NHWC features, f32 weights, f32 planar logits [N,3,h*w].  HBM-bound: one warp per pixel, 32 B per lane.
//  * nbc_upsample_argmax / nbc_upsample_bicubic: F.interpolate(mode='bicubic', align_corners=False)
//    (models.py:38-41) fused with torch.argmax(dim=1) (models.py:270).  The 12.6 MB f32 full-resolution logits
//    are never written in the argmax variant.  Index / weight arithmetic follows torch's upsample_bicubic2d in
//    f32 with explicit IEEE operations (no fma contraction) in the order of oracle/model.py
//    ``upsample_bicubic_restated`` so the result is bit-identical to that restatement.
#include "common.cuh"
#include "cubic.cuh"

namespace nbc {

// kIters = Cin / 256: every lane owns 8 consecutive channels of each 256-channel slab and keeps its 3 x 8 x kIters
// weights in registers, so the inner loop is one 16-byte load + 24 FMAs per slab.
template <int kIters>
__global__ void __launch_bounds__(256) head1x1_kernel(const __nv_bfloat16* __restrict__ x, int P, int N, int Cin,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      float* __restrict__ logits, int f16, const int* __restrict__ valid_h,
                                                      int row_w) {
  const int lane = threadIdx.x & 31;
  // 32-bit pixel indices (the host checks N * P < 2^31): the first version's 64-bit divisions per pixel cost more than the
  // 48 FMAs of the pixel
  const int warp_global = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int M = N * P;
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
  float wr[kIters][3][8];
#pragma unroll
  for (int it = 0; it < kIters; ++it)
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) wr[it][k][e] = __ldg(w + k * Cin + it * 256 + lane * 8 + e);
  // kPix pixels per warp and iteration: their loads are all issued before the first FMA (one pixel at a time left the
  // kernel at 1.5 TB/s -- too little memory-level parallelism)
  constexpr int kPix = 4;
  for (int m0 = warp_global; m0 < M; m0 += nwarps * kPix) {
    uint4 v[kPix][kIters];
    bool on[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      const int m = m0 + j * nwarps;
      on[j] = m < M;
      if (on[j] && valid_h != nullptr) {   // ragged batch: rows at or below an image's last valid one feed nothing (K3 clamps its taps)
        const int img = m / P;
        on[j] = (m - img * P) / row_w < __ldg(valid_h + img);
      }
      if (on[j]) {
        const __nv_bfloat16* xr = x + (int64_t)m * Cin + lane * 8;
#pragma unroll
        for (int it = 0; it < kIters; ++it) v[j][it] = __ldg(reinterpret_cast<const uint4*>(xr + it * 256));
      }
    }
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      if (!on[j]) continue;      // warp-uniform
      const int m = m0 + j * nwarps;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int it = 0; it < kIters; ++it) {
        const uint32_t u[4] = {v[j][it].x, v[j][it].y, v[j][it].z, v[j][it].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float f0 = lo16(u[k], f16), f1 = hi16(u[k], f16);
          a0 = fmaf(f0, wr[it][0][2 * k], a0), a0 = fmaf(f1, wr[it][0][2 * k + 1], a0);
          a1 = fmaf(f0, wr[it][1][2 * k], a1), a1 = fmaf(f1, wr[it][1][2 * k + 1], a1);
          a2 = fmaf(f0, wr[it][2][2 * k], a2), a2 = fmaf(f1, wr[it][2][2 * k + 1], a2);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (lane == 0) {
        const int img = m / P, pix = m - img * P;
        float* o = logits + (int64_t)img * 3 * P + pix;
        o[0] = a0 + b0;
        o[P] = a1 + b1;
        o[2 * P] = a2 + b2;
      }
    }
  }
}

__device__ __forceinline__ float cubic_sample(const float* __restrict__ plane, int w, const int (&iy)[4],
                                              const float (&wy)[4], const int (&ix)[4], const float (&wx)[4]) {
  float out = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* r = plane + (int64_t)iy[i] * w;
    float inner = __fmul_rn(__ldg(r + ix[0]), wx[0]);
    inner = __fadd_rn(inner, __fmul_rn(__ldg(r + ix[1]), wx[1]));
    inner = __fadd_rn(inner, __fmul_rn(__ldg(r + ix[2]), wx[2]));
    inner = __fadd_rn(inner, __fmul_rn(__ldg(r + ix[3]), wx[3]));
    const float term = __fmul_rn(inner, wy[i]);
    out = (i == 0) ? term : __fadd_rn(out, term);
  }
  return out;
}

template <bool kArgmax>
__global__ void __launch_bounds__(256) upsample_kernel(const float* __restrict__ logits, int N, int C, int h, int w,
                                                       int H, int W, float scale_y, float scale_x,
                                                       uint8_t* __restrict__ mask, float* __restrict__ out) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  const int n = blockIdx.z;
  if (X >= W) return;
  int ix[4], iy[4];
  float wx[4], wy[4];
  cubic_taps(X, scale_x, w, ix, wx);
  cubic_taps(Y, scale_y, h, iy, wy);
  const int64_t plane = (int64_t)h * w;
  if (kArgmax) {
    float best = 0.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      const float v = cubic_sample(logits + ((int64_t)n * C + c) * plane, w, iy, wy, ix, wx);
      if (c == 0 || v > best) best = v, arg = c;  // ties -> lowest index (torch.argmax)
    }
    mask[((int64_t)n * H + Y) * W + X] = (uint8_t)arg;
  } else {
    for (int c = 0; c < C; ++c)
      out[(((int64_t)n * C + c) * H + Y) * W + X] =
          cubic_sample(logits + ((int64_t)n * C + c) * plane, w, iy, wy, ix, wx);
  }
}

// Ragged batch: image n has H_n valid rows in a canvas of Hc rows; its logits occupy ceil(H_n/8) rows of the hc-row
// logits canvas -- per-image in / out heights (scale = h_n / H_n), otherwise the arithmetic of upsample_kernel<true>.
// K3 with the x pass SHARED by the output rows of a group.  An output pixel is a 4 x 4 cubic sample of each class plane:
// the inner (x) sums depend on (source row, X) only, and kRows = 8 consecutive output rows of an 8x upsample touch at
// most 6 distinct source rows -- so a thread (one X) computes the <= 6 x 3 inner sums once, keeps them in its own column
// of shared memory (no synchronisation: a thread only reads what it wrote) and finishes its 8 outputs from them: 9
// global loads per output pixel instead of 48.  Exactly the arithmetic of cubic_sample (same operations, same order), so
// the mask is bit-identical; a group that would need more than kSrc source rows (upscaling by less than ~2x) takes the
// per-pixel path.  heights == nullptr: dense batch (every image H rows, h logits rows).
constexpr int kK3Rows = 8, kK3Src = 6;
__global__ void __launch_bounds__(256) upsample_argmax_rows_kernel(const float* __restrict__ logits, int hc, int w, int Hc, int W,
                                                                   int h_dense, float scale_x, const int* __restrict__ heights,
                                                                   uint8_t* __restrict__ mask) {
  __shared__ float s_inner[3][kK3Src][256];
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y0 = blockIdx.y * kK3Rows;
  const int n = blockIdx.z;
  const int H = heights != nullptr ? min(max(__ldg(heights + n), 1), Hc) : Hc;
  if (X >= W || Y0 >= H) return;
  const int h = heights != nullptr ? ((((H - 1) / 2 + 1) - 1) / 2 + 1 - 1) / 2 + 1 : h_dense;
  const float scale_y = __fdiv_rn((float)h, (float)H);
  const int rows = min(kK3Rows, H - Y0);
  int ix[4];
  float wx[4];
  cubic_taps(X, scale_x, w, ix, wx);
  int iy[4];
  float wy[4];
  cubic_taps(Y0, scale_y, h, iy, wy);
  const int r_min = iy[0];
  cubic_taps(Y0 + rows - 1, scale_y, h, iy, wy);
  const int nsrc = iy[3] - r_min + 1;      // tap rows are non-decreasing in Y and in k
  const int64_t plane = (int64_t)hc * w;
  const float* base = logits + (int64_t)n * 3 * plane;
  uint8_t* out = mask + ((int64_t)n * Hc + Y0) * W + X;
  if (nsrc > kK3Src) {
    for (int y = 0; y < rows; ++y) {
      cubic_taps(Y0 + y, scale_y, h, iy, wy);
      float best = 0.f;
      int arg = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = cubic_sample(base + c * plane, w, iy, wy, ix, wx);
        if (c == 0 || v > best) best = v, arg = c;
      }
      out[(int64_t)y * W] = (uint8_t)arg;
    }
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < nsrc; ++r) {
      const float* rp = base + c * plane + (int64_t)(r_min + r) * w;
      float inner = __fmul_rn(__ldg(rp + ix[0]), wx[0]);
      inner = __fadd_rn(inner, __fmul_rn(__ldg(rp + ix[1]), wx[1]));
      inner = __fadd_rn(inner, __fmul_rn(__ldg(rp + ix[2]), wx[2]));
      inner = __fadd_rn(inner, __fmul_rn(__ldg(rp + ix[3]), wx[3]));
      s_inner[c][r][threadIdx.x] = inner;
    }
  }
  for (int y = 0; y < rows; ++y) {
    cubic_taps(Y0 + y, scale_y, h, iy, wy);
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fmul_rn(s_inner[c][iy[0] - r_min][threadIdx.x], wy[0]);
#pragma unroll
      for (int k = 1; k < 4; ++k) v = __fadd_rn(v, __fmul_rn(s_inner[c][iy[k] - r_min][threadIdx.x], wy[k]));
      if (c == 0 || v > best) best = v, arg = c;
    }
    out[(int64_t)y * W] = (uint8_t)arg;
  }
}

__global__ void heights_kernel(const int* __restrict__ first_last, int N, int* __restrict__ heights) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) heights[n] = first_last[2 * n + 1] - first_last[2 * n];
}

}  // namespace nbc

using namespace nbc;

extern "C" int nbc_heights_from_first_last(const int32_t* first_last, int N, int32_t* heights, void* stream) {
  NBC_REQUIRE(first_last && heights && N > 0, "nbc_heights_from_first_last: bad argument");
  heights_kernel<<<ceil_div(N, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(first_last, N, heights);
  NBC_CHECK_LAUNCH();
  return 0;
}

extern "C" int nbc_upsample_argmax_ragged(const float* logits, int N, int hc, int w, int Hc, int W, const int32_t* heights,
                                          uint8_t* mask, void* stream) {
  NBC_REQUIRE(logits && heights && mask, "nbc_upsample_argmax_ragged: null pointer");
  NBC_REQUIRE(N > 0 && hc > 0 && w > 0 && Hc > 0 && W > 0 && Hc <= 65535 && N <= 65535, "nbc_upsample_argmax_ragged: bad shape");
  dim3 grid(ceil_div(W, 256), ceil_div(Hc, kK3Rows), N);
  upsample_argmax_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, hc, w, Hc, W, 0,
                                                                                         (float)w / (float)W, heights, mask);
  NBC_CHECK_LAUNCH();
  return 0;
}

namespace nbc {
int head_1x1(const void* x_bf16, int64_t pixels_per_image, int N, int Cin, const float* w3xC, const float* bias3,
             float* logits_planar, int f16, cudaStream_t stream, const int* valid_h, int row_w) {
  NBC_REQUIRE(x_bf16 && w3xC && bias3 && logits_planar, "nbc_head_1x1: null pointer");
  NBC_REQUIRE((Cin == 256 || Cin == 512 || Cin == 1024) && N > 0 && pixels_per_image > 0,
              "nbc_head_1x1: Cin must be 256, 512 or 1024 (got %d)", Cin);
  const int64_t M = (int64_t)N * pixels_per_image;
  NBC_REQUIRE(M < (1ll << 31) - (1 << 24), "nbc_head_1x1: N * pixels_per_image must be < 2^31");
  const int64_t want = ceil_div64(M, 8 * 4);
  const int blocks = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  if (Cin == 256)
    head1x1_kernel<1><<<blocks, 256, 0, stream>>>(xb, (int)pixels_per_image, N, Cin, w3xC, bias3, logits_planar, f16, valid_h, row_w);
  else if (Cin == 512)
    head1x1_kernel<2><<<blocks, 256, 0, stream>>>(xb, (int)pixels_per_image, N, Cin, w3xC, bias3, logits_planar, f16, valid_h, row_w);
  else
    head1x1_kernel<4><<<blocks, 256, 0, stream>>>(xb, (int)pixels_per_image, N, Cin, w3xC, bias3, logits_planar, f16, valid_h, row_w);
  NBC_CHECK_LAUNCH();
  return 0;
}
}  // namespace nbc

extern "C" int nbc_head_1x1(const void* x_bf16, int64_t pixels_per_image, int N, int Cin, int f16, const float* w3xC,
                            const float* bias3, float* logits_planar, void* stream) {
  return head_1x1(x_bf16, pixels_per_image, N, Cin, w3xC, bias3, logits_planar, f16 ? 1 : 0,
                  reinterpret_cast<cudaStream_t>(stream), nullptr, 1);
}

static int upsample_common(const float* logits, int N, int C, int h, int w, int H, int W, uint8_t* mask, float* out,
                           cudaStream_t stream) {
  NBC_REQUIRE(logits && (mask || out), "nbc_upsample: null pointer");
  NBC_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0 && H <= 65535 && N <= 65535, "nbc_upsample: bad shape");
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;  // IEEE f32 division, as torch
  dim3 grid(ceil_div(W, 256), H, N);
  if (mask && C == 3)
    upsample_argmax_rows_kernel<<<dim3(ceil_div(W, 256), ceil_div(H, kK3Rows), N), 256, 0, stream>>>(logits, h, w, H, W, h, sx, nullptr,
                                                                                                    mask);
  else if (mask)
    upsample_kernel<true><<<grid, 256, 0, stream>>>(logits, N, C, h, w, H, W, sy, sx, mask, nullptr);
  else
    upsample_kernel<false><<<grid, 256, 0, stream>>>(logits, N, C, h, w, H, W, sy, sx, nullptr, out);
  NBC_CHECK_LAUNCH();
  return 0;
}

extern "C" int nbc_upsample_argmax(const float* logits, int N, int h, int w, int H, int W, uint8_t* mask,
                                   void* stream) {
  NBC_REQUIRE(mask, "nbc_upsample_argmax: null mask");
  return upsample_common(logits, N, 3, h, w, H, W, mask, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int nbc_upsample_bicubic(const float* logits, int N, int C, int h, int w, int H, int W, float* out,
                                    void* stream) {
  NBC_REQUIRE(out, "nbc_upsample_bicubic: null out");
  return upsample_common(logits, N, C, h, w, H, W, nullptr, out, reinterpret_cast<cudaStream_t>(stream));
}
