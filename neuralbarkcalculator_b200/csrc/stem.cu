// Stem of the network: ToTensor + Normalize (dataset.py:181-190, models.py:233-237) fused into
// conv1 7x7/2 pad 3 + bn1 + ReLU (torchvision resnet50 stem, models.py:127-131), then maxpool 3x3/2 pad 1.
//
// The input is the u8 NHWC image as decoded from the processed PNG; normalisation is done in f32 *before* the
// zero padding exactly as the reference does ((u8/255 - mean)/std, IEEE div/sub/div), so the border is right.
// K = 7*7*3 = 147 is too ragged for a TMA/UMMA tile; this first version runs on CUDA cores in f32:
// one CTA = 16x16 output pixels x 64 channels, input tile and all weights staged in shared memory.
#include "common.cuh"
#include "conv.h"

namespace nbc {

constexpr int ST_T = 16;                    // output tile edge
constexpr int ST_IN = 2 * ST_T + 5;         // 37 input rows / cols
constexpr int ST_W_FLOATS = 147 * 64;       // [tap*3+c][oc]
constexpr int ST_IN_FLOATS = ST_IN * ST_IN * 3;
constexpr int ST_SMEM = (ST_W_FLOATS + ST_IN_FLOATS) * 4;

struct StemParams {
  const uint8_t* img;   // u8 NHWC (kF32 == false)
  const float* xf;      // f32 NCHW, already normalised (kF32 == true)
  const float* w;     // [64][7][7][3], BN folded
  const float* bias;  // [64]
  __nv_bfloat16* out; // [N][Ho][Wo][64]
  int N, H, W, Ho, Wo;
  float mean[3], std[3];
  int f16;  // 16-bit output format: 0 bf16, 1 fp16
};

template <bool kF32>
__global__ void __launch_bounds__(256) stem_kernel(const StemParams p) {
  extern __shared__ float smem_f[];
  float* sw = smem_f;                 // [147][64]
  float* sin = smem_f + ST_W_FLOATS;  // [37][37][3]
  const int tid = threadIdx.x;
  const int img = blockIdx.z;
  const int oh0 = blockIdx.y * ST_T, ow0 = blockIdx.x * ST_T;
  for (int i = tid; i < ST_W_FLOATS; i += 256) {
    const int oc = i / 147, k = i - oc * 147;  // source order [oc][k]
    sw[k * 64 + oc] = __ldg(p.w + i);
  }
  const int ih0 = 2 * oh0 - 3, iw0 = 2 * ow0 - 3;
  for (int i = tid; i < ST_IN * ST_IN; i += 256) {
    const int r = i / ST_IN, c = i - r * ST_IN;
    const int ih = ih0 + r, iw = iw0 + c;
    float v[3] = {0.f, 0.f, 0.f};
    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
      if (kF32) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = __ldg(p.xf + (((int64_t)img * 3 + ch) * p.H + ih) * p.W + iw);
      } else {
        const uint8_t* s = p.img + (((int64_t)img * p.H + ih) * p.W + iw) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          v[ch] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[ch], 255.f), p.mean[ch]), p.std[ch]);
      }
    }
    sin[i * 3] = v[0], sin[i * 3 + 1] = v[1], sin[i * 3 + 2] = v[2];
  }
  __syncthreads();

  const int px = tid & 15, py = tid >> 4;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  for (int ky = 0; ky < 7; ++ky) {
    const float* row = sin + ((2 * py + ky) * ST_IN + 2 * px) * 3;
    const float* wk = sw + ky * 21 * 64;
#pragma unroll 3
    for (int j = 0; j < 21; ++j) {  // kx*3 + c
      const float v = row[j];
      const float4* w4 = reinterpret_cast<const float4*>(wk + j * 64);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 w = w4[q];
        acc[4 * q] = fmaf(v, w.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
      }
    }
  }
  const int oh = oh0 + py, ow = ow0 + px;
  if (oh < p.Ho && ow < p.Wo) {
    uint4* o = reinterpret_cast<uint4*>(p.out + (((int64_t)img * p.Ho + oh) * p.Wo + ow) * 64);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = q * 8 + 2 * k;
        r[k] = pack16x2(fmaxf(acc[c] + __ldg(p.bias + c), 0.f), fmaxf(acc[c + 1] + __ldg(p.bias + c + 1), 0.f), p.f16);
      }
      o[q] = make_uint4(r[0], r[1], r[2], r[3]);
    }
  }
}

// ---- tensor-core stem: staging pass -----------------------------------------------------------------------------
// Writes the normalised image as bf16 [N][Hp][Wp][4] (channel 3 = 0) with an explicit zero border of 3 pixels, so
// that tap row ky of output (ho, wo) is the 64 contiguous bytes starting at padded pixel (2*ho + ky, 2*wo).
// Hp = 2*Ho + 5, Wp = 2*Wo + 6.  The implicit GEMM itself is conv_tc_kernel<64, 32> (conv_tc.cu).
template <bool kF32>
__global__ void __launch_bounds__(256) stem_pad_kernel(const StemParams p, int Hp, int Wp, uint2* __restrict__ padded,
                                                       const int* __restrict__ valid_h) {
  // grid (padded columns / 256, padded rows, images): plain 32-bit index arithmetic (the first version's grid-stride loop
  // spent its time in 64-bit divisions)
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  if (xp >= Wp) return;
  const int x = xp - 3, y = (int)blockIdx.y - 3, img = blockIdx.z;
  const int vh = valid_h ? min(p.H, __ldg(valid_h + img)) : p.H;   // ragged batch: rows >= vh are zero padding
  // ... of which the stem reads at most the 7-row window of its last valid output row (2 * ceil(vh / 2) + 3 < vh + 5);
  // the dead part of the canvas below is never read and need not be written
  if (valid_h && y >= vh + 8) return;
  float v[3] = {0.f, 0.f, 0.f};
  if (y >= 0 && y < vh && x >= 0 && x < p.W) {
    if (kF32) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) v[ch] = __ldg(p.xf + (((int64_t)img * 3 + ch) * p.H + y) * p.W + x);
    } else {
      const uint8_t* s = p.img + (((int64_t)img * p.H + y) * p.W + x) * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        v[ch] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[ch], 255.f), p.mean[ch]), p.std[ch]);
    }
  }
  padded[((int64_t)img * Hp + blockIdx.y) * Wp + xp] = make_uint2(pack16x2(v[0], v[1], p.f16), pack16x2(v[2], 0.f, p.f16));
}

// stem weights f32 [64][7][7][3] (BN folded) -> bf16 [64][7][8][4], zero for kx == 7 or c == 3  (K = 224)
__global__ void stem_pack_kernel(const float* __restrict__ w, unsigned short* __restrict__ out, int f16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 224) return;
  const int c = i & 3, kx = (i >> 2) & 7, ky = (i >> 5) % 7, oc = i / 224;
  float v = 0.f;
  if (c < 3 && kx < 7) v = w[((oc * 7 + ky) * 7 + kx) * 3 + c];
  out[i] = cvt16(v, f16);
}

// maxpool 3x3 stride 2 pad 1 (padding = -inf), 16-bit NHWC, 8 channels per thread; grid (Wo * C/8 / 256, Ho, N).  The
// maximum is taken on the packed 16-bit pairs (HMNMX2): the result is one of the stored values, bit for bit.
__device__ __forceinline__ uint32_t max16x2(uint32_t a, uint32_t b, int f16) {
  if (f16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__global__ void __launch_bounds__(256) maxpool_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                                      int Ho, int Wo, __nv_bfloat16* __restrict__ y,
                                                      const int* __restrict__ valid_h, int f16) {
  const int cg = C >> 3;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Wo * cg) return;
  const int c8 = t % cg, wo = t / cg, ho = blockIdx.y, n = blockIdx.z;
  uint4* dst = reinterpret_cast<uint4*>(y + (((int64_t)n * Ho + ho) * Wo + wo) * C + c8 * 8);
  if (valid_h != nullptr) {   // ragged batch: zero halo rows after the valid ones, nothing beyond
    const int vh = __ldg(valid_h + n);
    if (ho >= vh) {
      if (ho < vh + 4) *dst = make_uint4(0, 0, 0, 0);
      return;
    }
  }
  const uint32_t neg_inf = f16 ? 0xFC00FC00u : 0xFF80FF80u;
  uint32_t m[4] = {neg_inf, neg_inf, neg_inf, neg_inf};
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int h = 2 * ho + dy;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int w = 2 * wo + dx;
      if (w < 0 || w >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)n * H + h) * W + w) * C + c8 * 8));
      m[0] = max16x2(m[0], v.x, f16), m[1] = max16x2(m[1], v.y, f16), m[2] = max16x2(m[2], v.z, f16), m[3] = max16x2(m[3], v.w, f16);
    }
  }
  *dst = make_uint4(m[0], m[1], m[2], m[3]);
}

}  // namespace nbc

using namespace nbc;

static int stem_launch(const uint8_t* img, const float* xf, int N, int H, int W, const float* mean3, const float* std3,
                       const float* w_stem, const float* bias, void* out, cudaStream_t stream, int f16 = 0) {
  NBC_REQUIRE((img || xf) && w_stem && bias && out, "nbc_stem: null pointer");
  NBC_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "nbc_stem: bad shape");
  static bool attr_set = false;
  if (!attr_set) {
    NBC_CUDA(cudaFuncSetAttribute(stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    NBC_CUDA(cudaFuncSetAttribute(stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    attr_set = true;
  }
  StemParams p;
  p.img = img, p.xf = xf, p.w = w_stem, p.bias = bias, p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.N = N, p.H = H, p.W = W, p.Ho = (H + 6 - 7) / 2 + 1, p.Wo = (W + 6 - 7) / 2 + 1;
  p.f16 = f16;
  for (int i = 0; i < 3; ++i) p.mean[i] = mean3 ? mean3[i] : 0.f, p.std[i] = std3 ? std3[i] : 1.f;
  dim3 grid(ceil_div(p.Wo, ST_T), ceil_div(p.Ho, ST_T), N);
  if (xf)
    stem_kernel<true><<<grid, 256, ST_SMEM, stream>>>(p);
  else
    stem_kernel<false><<<grid, 256, ST_SMEM, stream>>>(p);
  NBC_CHECK_LAUNCH();
  return 0;
}

extern "C" int nbc_stem_u8(const uint8_t* img, int N, int H, int W, const float* mean3_host, const float* std3_host,
                           const float* w_stem, const float* bias, void* out, void* stream) {
  NBC_REQUIRE(img && mean3_host && std3_host, "nbc_stem_u8: null pointer");
  return stem_launch(img, nullptr, N, H, W, mean3_host, std3_host, w_stem, bias, out,
                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int nbc_stem_f32(const float* x_nchw, int N, int H, int W, const float* w_stem, const float* bias, void* out,
                            void* stream) {
  NBC_REQUIRE(x_nchw, "nbc_stem_f32: null pointer");
  return stem_launch(nullptr, x_nchw, N, H, W, nullptr, nullptr, w_stem, bias, out,
                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" size_t nbc_stem_tc_workspace_bytes(int N, int H, int W) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  return align_up((size_t)N * (2 * Ho + 5) * (2 * Wo + 6) * 8, 1024);
}

extern "C" int nbc_stem_pack_weights(const float* w_stem_f32, int f16, void* w224_bf16, void* stream) {
  NBC_REQUIRE(w_stem_f32 && w224_bf16, "nbc_stem_pack_weights: null pointer");
  stem_pack_kernel<<<ceil_div(64 * 224, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w_stem_f32, reinterpret_cast<unsigned short*>(w224_bf16), f16 ? 1 : 0);
  NBC_CHECK_LAUNCH();
  return 0;
}

namespace nbc {
int stem_tc_pad(const void* input, int input_kind, int N, int H, int W, const float* mean3, const float* std3, void* padded,
                cudaStream_t stream, const int* valid_h, int f16) {
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.f16 = f16;
  p.img = input_kind == 0 ? reinterpret_cast<const uint8_t*>(input) : nullptr;
  p.xf = input_kind == 1 ? reinterpret_cast<const float*>(input) : nullptr;
  p.N = N, p.H = H, p.W = W, p.Ho = (H - 1) / 2 + 1, p.Wo = (W - 1) / 2 + 1;
  for (int i = 0; i < 3; ++i) p.mean[i] = mean3 ? mean3[i] : 0.f, p.std[i] = std3 ? std3[i] : 1.f;
  const int Hp = 2 * p.Ho + 5, Wp = 2 * p.Wo + 6;
  NBC_REQUIRE(Hp <= 65535 && N <= 65535, "stem staging: image too tall / batch too large");
  const dim3 grid(ceil_div(Wp, 256), Hp, N);
  if (input_kind == 1)
    stem_pad_kernel<true><<<grid, 256, 0, stream>>>(p, Hp, Wp, reinterpret_cast<uint2*>(padded), valid_h);
  else
    stem_pad_kernel<false><<<grid, 256, 0, stream>>>(p, Hp, Wp, reinterpret_cast<uint2*>(padded), valid_h);
  NBC_CHECK_LAUNCH();
  return 0;
}
}  // namespace nbc

extern "C" int nbc_stem_tc(const void* input, int input_kind, int N, int H, int W, const float* mean3_host,
                           const float* std3_host, const void* w224_bf16, const float* bias, int f16, void* workspace,
                           size_t workspace_bytes, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NBC_REQUIRE(input && w224_bf16 && bias && workspace && out, "nbc_stem_tc: null pointer");
  NBC_REQUIRE(input_kind == 0 || input_kind == 1, "nbc_stem_tc: input_kind must be 0 (u8 NHWC) or 1 (f32 NCHW)");
  NBC_REQUIRE(input_kind == 1 || (mean3_host && std3_host), "nbc_stem_tc: mean/std required for u8 input");
  NBC_REQUIRE(N > 0 && H > 0 && W > 0, "nbc_stem_tc: bad shape");
  NBC_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "nbc_stem_tc: workspace must be 16-byte aligned");
  if (workspace_bytes < nbc_stem_tc_workspace_bytes(N, H, W)) {
    set_error("nbc_stem_tc: workspace too small");
    return NBC_ERR_WORKSPACE;
  }
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  int rc = stem_tc_pad(input, input_kind, N, H, W, mean3_host, std3_host, workspace, stream, nullptr, f16 ? 1 : 0);
  if (rc) return rc;
  ConvTcPrepared prep;
  rc = conv_tc_prepare_stem(N, Ho, Wo, 2 * Ho + 5, 2 * Wo + 6, workspace, w224_bf16, bias, out, &prep, nullptr, f16 ? 1 : 0);
  if (rc) return rc;
  return conv_tc_run(&prep, stream);
}

extern "C" int nbc_maxpool3x3s2_bf16(const void* x, int N, int H, int W, int C, int f16, void* y, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NBC_REQUIRE(x && y, "nbc_maxpool3x3s2_bf16: null pointer");
  NBC_REQUIRE(C % 8 == 0 && N > 0 && H > 0 && W > 0, "nbc_maxpool3x3s2_bf16: bad shape");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  NBC_REQUIRE(Ho <= 65535 && N <= 65535, "nbc_maxpool3x3s2_bf16: image too tall / batch too large");
  maxpool_kernel<<<dim3(ceil_div(Wo * (C / 8), 256), Ho, N), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C, Ho,
                                                                              Wo, reinterpret_cast<__nv_bfloat16*>(y), nullptr,
                                                                              f16 ? 1 : 0);
  NBC_CHECK_LAUNCH();
  return 0;
}

namespace nbc {
int maxpool_ragged(const void* x, int N, int H, int W, int C, void* y, const int* valid_h, cudaStream_t stream, int f16) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool_kernel<<<dim3(ceil_div(Wo * (C / 8), 256), Ho, N), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C, Ho,
                                                                              Wo, reinterpret_cast<__nv_bfloat16*>(y), valid_h, f16);
  NBC_CHECK_LAUNCH();
  return 0;
}

// per-image valid rows at the four resolution levels of the network: levels[0..3][N] = H, H/2, H/4, H/8 (ceil chain);
// heights come either as an int[N] array or as the {first,last} pairs K1 writes (first_last != nullptr)
__global__ void ragged_levels_kernel(const int* __restrict__ heights, const int* __restrict__ first_last, int N, int Hc,
                                     int* __restrict__ levels) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int h = first_last ? first_last[2 * n + 1] - first_last[2 * n] : heights[n];
  h = max(1, min(h, Hc));
  levels[n] = h;
  h = (h - 1) / 2 + 1;
  levels[N + n] = h;
  h = (h - 1) / 2 + 1;
  levels[2 * N + n] = h;
  h = (h - 1) / 2 + 1;
  levels[3 * N + n] = h;
}
int ragged_levels(const int* heights, const int* first_last, int N, int Hc, int* levels, cudaStream_t stream) {
  ragged_levels_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(heights, first_last, N, Hc, levels);
  NBC_CHECK_LAUNCH();
  return 0;
}
}  // namespace nbc
