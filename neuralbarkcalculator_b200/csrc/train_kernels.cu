// Training-path kernels (row a14: the step of __main__.py:231-269 -- train-mode forward, weighted CE, backward, Adam).
//
// All activations are bf16 NHWC, all statistics / gradients of parameters f32.  Per conv-BN unit the forward is
//   z = conv(x, w)            (conv_tc.cu, raw weights, no bias)
//   mean/var over N*H*W       (bn_stats_* : deterministic two-stage reduction)
//   y = relu?(gamma (z-mean) invstd + beta (+ residual))       (bn_apply)
// and the backward
//   g  = dy * (y > 0)                                          (ReLU mask, fused)
//   s1 = sum g, s2 = sum g * xhat                              (bn_bwd_reduce, two-stage)
//   dz = gamma invstd (g - s1/M - xhat s2/M);  dgamma = s2, dbeta = s1      (bn_bwd_apply)
//   dx = conv(dz, flip(w)^T)   (conv_tc.cu again, with dgrad-packed weights)   dW = wgrad(dz, x)   (wgrad_mma)
// torch's BatchNorm2d (momentum 0.1, eps 1e-5, unbiased running_var) and Adam (L2 weight decay added to the
// gradient) semantics are followed exactly; see oracle/train.py.
#include "common.cuh"
#include "conv.h"
#include "train.h"

namespace nbc {

// ------------------------------------------------------------------------------------------------ BN reductions
// z: [M][C] 16-bit.  Each block owns a slab of rows.  Thread t covers the 8 channels of 16-byte vector (t % VL) on
// the rows rl, rl + RL, ... of the slab (VL = min(C/8, 256) vectors per row pass, RL = 256 / VL row lanes), so every
// warp instruction reads whole 128-byte lines whatever C is; the RL row lanes are then summed through shared memory.
// partial: [nblocks][C][2] f32.
constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 148 * 8;

static int bn_blocks(int64_t M) {
  const int64_t nb = ceil_div64(M, 32);
  return (int)(nb < kBnMaxBlocks ? (nb < 1 ? 1 : nb) : kBnMaxBlocks);
}

// sums the RL row-lane copies of the block's 16 * VL accumulators and writes them to partial[blockIdx.x]
__device__ __forceinline__ void bn_block_reduce(float (&acc)[16], float* red /* [RL][VL*16] smem */, int VL, int RL, int v0,
                                                int rl, bool active, int v /* vector index this pass */, int C,
                                                float* __restrict__ partial) {
  if (RL == 1) {
    if (active) {
      float4* o = reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.x * C + (int64_t)v * 8) * 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
    }
    return;
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int k = 0; k < 16; ++k) red[(rl * VL + v0) * 16 + k] = acc[k];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < VL * 16; o += kBnThreads) {
    float sum = 0.f;
    for (int r = 0; r < RL; ++r) sum += red[r * VL * 16 + o];
    partial[(int64_t)blockIdx.x * C * 2 + o] = sum;   // o = (v * 8 + e) * 2 + which, v < VL = C / 8
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_stats_partial(const uint4* __restrict__ z, int64_t M, int C,
                                                               int64_t rows_per_block, float* __restrict__ partial) {
  __shared__ float red[kBnThreads * 16];
  const int vecs = C >> 3;
  const int VL = vecs < kBnThreads ? vecs : kBnThreads, RL = kBnThreads / VL;
  const int v0 = threadIdx.x % VL, rl = threadIdx.x / VL;
  const bool active = rl < RL;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  for (int v = v0; v < vecs; v += VL) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    if (active) {
#pragma unroll 4
      for (int64_t r = r0 + rl; r < r1; r += RL) {
        const uint4 q = __ldg(z + r * vecs + v);
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = bf16lo(u[k]), b = bf16hi(u[k]);
          acc[4 * k] += a, acc[4 * k + 1] = fmaf(a, a, acc[4 * k + 1]);
          acc[4 * k + 2] += b, acc[4 * k + 3] = fmaf(b, b, acc[4 * k + 3]);
        }
      }
    }
    bn_block_reduce(acc, red, VL, RL, v0, rl, active, v, C, partial);
  }
}

// block-wide sum (128 threads, one block per channel) of the per-block partials of one channel (two values), in double,
// in a fixed order; every thread returns the totals
constexpr int kBnFinThreads = 128;
__device__ __forceinline__ void bn_sum_partials(const float* __restrict__ partial, int nblocks, int C, int c, double& s,
                                                double& q) {
  __shared__ double sh[2][kBnFinThreads / 32];
  s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += kBnFinThreads) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((int64_t)i * C + c) * 2));
    s += (double)v.x, q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s, sh[1][threadIdx.x >> 5] = q;
  __syncthreads();
  s = 0.0, q = 0.0;
#pragma unroll
  for (int w = 0; w < kBnFinThreads / 32; ++w) s += sh[0][w], q += sh[1][w];
}

// one block per channel: sums the partials in double, produces the affine a = gamma*invstd, b = beta - mean*a,
// saves mean / invstd for the backward and updates the running statistics like torch (momentum, unbiased var)
__global__ void __launch_bounds__(kBnFinThreads) bn_stats_finalize(const float* __restrict__ partial, int nblocks, int C, int64_t M,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                  float* __restrict__ running_mean, float* __restrict__ running_var,
                                  float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ a_out,
                                  float* __restrict__ b_out) {
  const int c = blockIdx.x;
  double s, q;
  bn_sum_partials(partial, nblocks, C, c, s, q);
  if (threadIdx.x != 0) return;
  const double mean = s / (double)M;
  double var = q / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  save_mean[c] = (float)mean;
  save_invstd[c] = invstd;
  const float a = gamma[c] * invstd;
  a_out[c] = a;
  b_out[c] = beta[c] - (float)mean * a;
  if (running_mean != nullptr) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = relu?(z * a[c] + b[c] (+ residual)), 8 channels per thread
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ a,
                                                       const float* __restrict__ b, const uint4* __restrict__ residual,
                                                       int64_t total8, int C8, int relu, uint4* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const uint4 v = __ldg(z + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t rr[4] = {0, 0, 0, 0};
    if (residual != nullptr) {
      const uint4 rv = __ldg(residual + i);
      rr[0] = rv.x, rr[1] = rv.y, rr[2] = rv.z, rr[3] = rv.w;
    }
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(a + c)), a1 = __ldg(reinterpret_cast<const float4*>(a + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f0 = fmaf(bf16lo(u[k]), av[2 * k], bv[2 * k]);
      float f1 = fmaf(bf16hi(u[k]), av[2 * k + 1], bv[2 * k + 1]);
      if (residual != nullptr) f0 += bf16lo(rr[k]), f1 += bf16hi(rr[k]);
      if (relu) f0 = fmaxf(f0, 0.f), f1 = fmaxf(f1, 0.f);
      o[k] = pack_bf16x2(f0, f1);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------ BN backward
// partial: [nblocks][C][2] = (sum g, sum g * xhat) with g = dy * (y > 0 if relu), xhat = (z - mean) * invstd
__global__ void __launch_bounds__(kBnThreads) bn_bwd_partial(const uint4* __restrict__ dy, const uint4* __restrict__ y,
                                                             const uint4* __restrict__ z, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, int64_t M, int C, int relu,
                                                             int64_t rows_per_block, float* __restrict__ partial) {
  __shared__ float red[kBnThreads * 16];
  const int vecs = C >> 3;
  const int VL = vecs < kBnThreads ? vecs : kBnThreads, RL = kBnThreads / VL;
  const int v0 = threadIdx.x % VL, rl = threadIdx.x / VL;
  const bool active = rl < RL;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  for (int v = v0; v < vecs; v += VL) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    if (active) {
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean + v * 8)), m1 = __ldg(reinterpret_cast<const float4*>(mean + v * 8 + 4));
      const float4 i0 = __ldg(reinterpret_cast<const float4*>(invstd + v * 8)), i1 = __ldg(reinterpret_cast<const float4*>(invstd + v * 8 + 4));
      const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      const float iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll 2
      for (int64_t r = r0 + rl; r < r1; r += RL) {
        const uint4 gq = __ldg(dy + r * vecs + v), zq = __ldg(z + r * vecs + v);
        uint4 yq = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);   // "positive" when there is no ReLU
        if (relu) yq = __ldg(y + r * vecs + v);
        const uint32_t gu[4] = {gq.x, gq.y, gq.z, gq.w}, zu[4] = {zq.x, zq.y, zq.z, zq.w}, yu[4] = {yq.x, yq.y, yq.z, yq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float g0 = bf16lo(gu[k]), g1 = bf16hi(gu[k]);
          if (!(bf16lo(yu[k]) > 0.f)) g0 = 0.f;
          if (!(bf16hi(yu[k]) > 0.f)) g1 = 0.f;
          acc[4 * k] += g0, acc[4 * k + 1] = fmaf(g0, (bf16lo(zu[k]) - mv[2 * k]) * iv[2 * k], acc[4 * k + 1]);
          acc[4 * k + 2] += g1, acc[4 * k + 3] = fmaf(g1, (bf16hi(zu[k]) - mv[2 * k + 1]) * iv[2 * k + 1], acc[4 * k + 3]);
        }
      }
    }
    bn_block_reduce(acc, red, VL, RL, v0, rl, active, v, C, partial);
  }
}

// one block per channel: sums[c] = (s1, s2); accumulates dgamma += s2, dbeta += s1
__global__ void __launch_bounds__(kBnFinThreads) bn_bwd_finalize(const float* __restrict__ partial, int nblocks, int C,
                                                       float* __restrict__ sums, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta) {
  const int c = blockIdx.x;
  double s, t;
  bn_sum_partials(partial, nblocks, C, c, s, t);
  if (threadIdx.x != 0) return;
  sums[2 * c] = (float)s, sums[2 * c + 1] = (float)t;
  dgamma[c] += (float)t;
  dbeta[c] += (float)s;
}

// dz = gamma*invstd * (g - s1/M - xhat*s2/M);  optionally also writes g (the gradient that flows into the skip path)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y,
                                                           const uint4* __restrict__ z, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sums, float inv_m, int64_t total8,
                                                           int C8, int relu, uint4* __restrict__ dz, uint4* __restrict__ g_out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const uint4 dv = __ldg(dy + i), zv = __ldg(z + i);
    uint4 yv = make_uint4(0, 0, 0, 0);
    if (relu) yv = __ldg(y + i);
    const uint32_t du[4] = {dv.x, dv.y, dv.z, dv.w}, zu[4] = {zv.x, zv.y, zv.z, zv.w}, yu[4] = {yv.x, yv.y, yv.z, yv.w};
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean + c)), m1 = __ldg(reinterpret_cast<const float4*>(mean + c + 4));
    const float4 i0 = __ldg(reinterpret_cast<const float4*>(invstd + c)), i1 = __ldg(reinterpret_cast<const float4*>(invstd + c + 4));
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sums + 2 * c)), s1 = __ldg(reinterpret_cast<const float4*>(sums + 2 * c + 4));
    const float4 s2 = __ldg(reinterpret_cast<const float4*>(sums + 2 * c + 8)), s3 = __ldg(reinterpret_cast<const float4*>(sums + 2 * c + 12));
    const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const float iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float sa[8] = {s0.x, s0.z, s1.x, s1.z, s2.x, s2.z, s3.x, s3.z};   // sum g
    const float sb[8] = {s0.y, s0.w, s1.y, s1.w, s2.y, s2.w, s3.y, s3.w};   // sum g * xhat
    uint32_t o[4], go[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg[2] = {bf16lo(du[k]), bf16hi(du[k])};
      const float zz[2] = {bf16lo(zu[k]), bf16hi(zu[k])};
      const float yy[2] = {bf16lo(yu[k]), bf16hi(yu[k])};
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = 2 * k + e;
        if (relu && !(yy[e] > 0.f)) gg[e] = 0.f;
        const float xhat = (zz[e] - mv[cc]) * iv[cc];
        r[e] = gv[cc] * iv[cc] * (gg[e] - sa[cc] * inv_m - xhat * sb[cc] * inv_m);
      }
      o[k] = pack_bf16x2(r[0], r[1]);
      go[k] = pack_bf16x2(gg[0], gg[1]);
    }
    dz[i] = make_uint4(o[0], o[1], o[2], o[3]);
    if (g_out != nullptr) g_out[i] = make_uint4(go[0], go[1], go[2], go[3]);
  }
}

// ------------------------------------------------------------------------------------------------ weight packs
// master weights live in the forward layout [Cout][kh][kw][Cin] (f32).  fwd: plain cast; dgrad: [Cin][kh'][kw'][Cout]
// with the taps flipped (kh' = kh-1-ky), so that dx = conv(dz, w_dgrad) with the same padding / dilation.
__global__ void __launch_bounds__(256) cast_pack_kernel(const float* __restrict__ w, int64_t n, __nv_bfloat16* __restrict__ o) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = __float2bfloat16_rn(w[i]);
}
__global__ void __launch_bounds__(256) dgrad_pack_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw,
                                                         __nv_bfloat16* __restrict__ o) {
  const int64_t n = (int64_t)Cout * Cin * kh * kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    int64_t t = i / Cout;
    const int kx = (int)(t % kw);
    t /= kw;
    const int ky = (int)(t % kh);
    const int ci = (int)(t / kh);
    o[i] = __float2bfloat16_rn(w[(((int64_t)co * kh + (kh - 1 - ky)) * kw + (kw - 1 - kx)) * Cin + ci]);
  }
}
// OIHW f32 (torch state_dict) <-> [Cout][kh][kw][Cin] f32 (internal master layout)
__global__ void __launch_bounds__(256) oihw_to_ohwi_kernel(const float* __restrict__ src, int Cout, int Cin, int kh, int kw,
                                                           float* __restrict__ dst, int reverse) {
  const int64_t n = (int64_t)Cout * Cin * kh * kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t t = i / Cin;
    const int kx = (int)(t % kw);
    t /= kw;
    const int ky = (int)(t % kh);
    const int co = (int)(t / kh);
    const int64_t j = (((int64_t)co * Cin + ci) * kh + ky) * kw + kx;
    if (reverse)
      dst[j] = src[i];
    else
      dst[i] = src[j];
  }
}

// ------------------------------------------------------------------------------------------------ wgrad (mma.sync)
// dW[co][tap][ci] += sum over a chunk of output pixels p of dz[p][co] * x[p shifted by tap][ci].
// CTA tile 128 (co) x 64 (ci), K step = 32 pixels; both operands sit in smem as [pixel][channel] and are fed to
// mma.sync through ldmatrix.trans.  grid = (co tiles * ci tiles, taps, splitK); f32 atomics into dW.
struct WgradParams {
  const __nv_bfloat16* dz;  // [N][Ho][Wo][Cout]
  const __nv_bfloat16* x;   // [N][H][W][Cin]
  float* dw;                // [Cout][kh][kw][Cin]
  int N, H, W, Cin, Cout, Ho, Wo, kh, kw, stride, pad, dil;
  int64_t M;                // N*Ho*Wo
  int64_t chunk;            // pixels per split
  int ci_tiles;
};
constexpr int WG_LDA = 128 + 8, WG_LDB = 64 + 8;

__device__ __forceinline__ void cp_async16_z(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) wgrad_mma_kernel(const WgradParams p) {
  __shared__ __align__(16) __nv_bfloat16 sA[2][32][WG_LDA];  // [pixel][co]
  __shared__ __align__(16) __nv_bfloat16 sB[2][32][WG_LDB];  // [pixel][ci]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int co0 = (blockIdx.x / p.ci_tiles) * 128, ci0 = (blockIdx.x % p.ci_tiles) * 64;
  const int tap = blockIdx.y;
  const int ky = tap / p.kw, kx = tap - ky * p.kw;
  const int64_t m_begin = (int64_t)blockIdx.z * p.chunk;
  const int64_t m_end = min(p.M, m_begin + p.chunk);
  if (m_begin >= m_end) return;
  const int steps = (int)((m_end - m_begin + 31) / 32);

  auto load = [&](int st, int buf) {
    const int64_t mb = m_begin + (int64_t)st * 32;
    // A: 32 pixels x 128 co = 512 16-byte pieces, two per thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int piece = tid + i * 256;
      const int r = piece >> 4, c = (piece & 15) * 8;
      const int64_t m = mb + r;
      const bool ok = m < m_end && (co0 + c) < p.Cout;
      const __nv_bfloat16* src = ok ? p.dz + m * p.Cout + co0 + c : p.dz;
      cp_async16_z(smem_u32(&sA[buf][r][c]), src, ok ? 16 : 0);
    }
    // B: 32 pixels x 64 ci = 256 pieces, one per thread, pixel shifted by the tap (zero outside the image)
    {
      const int r = tid >> 3, c = (tid & 7) * 8;
      const int64_t m = mb + r;
      bool ok = m < m_end;
      const __nv_bfloat16* src = p.x;
      if (ok) {
        const int wo = (int)(m % p.Wo);
        const int64_t t = m / p.Wo;
        const int ho = (int)(t % p.Ho);
        const int n = (int)(t / p.Ho);
        const int hi = ho * p.stride + ky * p.dil - p.pad, wi = wo * p.stride + kx * p.dil - p.pad;
        ok = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
        if (ok) src = p.x + (((int64_t)n * p.H + hi) * p.W + wi) * p.Cin + ci0 + c;
      }
      cp_async16_z(smem_u32(&sB[buf][r][c]), src, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  load(0, 0);
  for (int st = 0; st < steps; ++st) {
    const int buf = st & 1;
    if (st + 1 < steps) {
      load(st + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 32; ks += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        // A fragment (m16 x k16) from smem [k][m]: matrices (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi)
        const int kr = ks + ((lane >> 4) << 3) + (lane & 7);
        const int mc = wm * 32 + i * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(smem_u32(&sA[buf][kr][mc]), a[i][0], a[i][1], a[i][2], a[i][3]);
      }
      uint32_t b[4][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // B fragments for two n8 tiles from smem [k][n]: matrices (k lo, n0), (k hi, n0), (k lo, n0+8), (k hi, n0+8)
        const int kr = ks + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int nc = wn * 32 + j * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(smem_u32(&sB[buf][kr][nc]), b[2 * j][0], b[2 * j][1], b[2 * j + 1][0], b[2 * j + 1][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_16816(acc[i][j], a[i], b[j][0], b[j][1]);
    }
    __syncthreads();
  }
  const int64_t taps = (int64_t)p.kh * p.kw;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int co = co0 + wm * 32 + i * 16 + (lane >> 2) + half * 8;
      if (co >= p.Cout) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = ci0 + wn * 32 + j * 8 + (lane & 3) * 2;
        float* d = p.dw + ((int64_t)co * taps + tap) * p.Cin + ci;
        atomicAdd(d, acc[i][j][half * 2]);
        atomicAdd(d + 1, acc[i][j][half * 2 + 1]);
      }
    }
}

int wgrad_mma(const ConvGeom& g, const void* dz, const void* x, float* dw, cudaStream_t stream) {
  NBC_REQUIRE(g.Cin % 64 == 0 && g.Cout % 64 == 0, "wgrad: Cin and Cout must be multiples of 64");
  WgradParams p;
  p.dz = reinterpret_cast<const __nv_bfloat16*>(dz), p.x = reinterpret_cast<const __nv_bfloat16*>(x), p.dw = dw;
  p.N = g.N, p.H = g.H, p.W = g.W, p.Cin = g.Cin, p.Cout = g.Cout, p.Ho = g.Ho(), p.Wo = g.Wo();
  p.kh = g.kh, p.kw = g.kw, p.stride = g.stride, p.pad = g.pad, p.dil = g.dil;
  p.M = (int64_t)g.N * p.Ho * p.Wo;
  p.ci_tiles = g.Cin / 64;
  const int co_tiles = ceil_div(g.Cout, 128);
  const int base = co_tiles * p.ci_tiles * g.kh * g.kw;
  int split = ceil_div(148 * 4, base);
  const int64_t max_split = ceil_div64(p.M, 2048);
  if (split > max_split) split = (int)max_split;
  if (split < 1) split = 1;
  p.chunk = ceil_div64(ceil_div64(p.M, split), 32) * 32;
  dim3 grid(co_tiles * p.ci_tiles, g.kh * g.kw, (unsigned)ceil_div64(p.M, p.chunk));
  wgrad_mma_kernel<<<grid, 256, 0, stream>>>(p);
  NBC_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ host wrappers
static int grid_for(int64_t n, int per_thread = 1) {
  const int64_t b = ceil_div64(n, 256LL * per_thread);
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

size_t bn_partial_bytes(int64_t M, int C) { return (size_t)bn_blocks(M) * C * 2 * sizeof(float); }

int bn_forward_train(const void* z, int64_t M, int C, const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, float* save_mean, float* save_invstd, float* a, float* b,
                     float* partial, const void* residual, int relu, void* y, cudaStream_t stream) {
  NBC_REQUIRE(C % 8 == 0, "bn: C must be a multiple of 8");
  const int nb = bn_blocks(M);
  const int64_t rpb = ceil_div64(M, nb);
  bn_stats_partial<<<nb, kBnThreads, 0, stream>>>(reinterpret_cast<const uint4*>(z), M, C, rpb, partial);
  NBC_CHECK_LAUNCH();
  bn_stats_finalize<<<C, kBnFinThreads, 0, stream>>>(partial, nb, C, M, gamma, beta, eps, momentum, running_mean,
                                                        running_var, save_mean, save_invstd, a, b);
  NBC_CHECK_LAUNCH();
  const int64_t total8 = M * C / 8;
  bn_apply_kernel<<<grid_for(total8), 256, 0, stream>>>(reinterpret_cast<const uint4*>(z), a, b,
                                                        reinterpret_cast<const uint4*>(residual), total8, C / 8, relu,
                                                        reinterpret_cast<uint4*>(y));
  NBC_CHECK_LAUNCH();
  return 0;
}

int bn_backward(const void* dy, const void* y, const void* z, int64_t M, int C, const float* gamma, const float* save_mean,
                const float* save_invstd, int relu, float* partial, float* sums, float* dgamma, float* dbeta, void* dz,
                void* g_out, cudaStream_t stream) {
  NBC_REQUIRE(C % 8 == 0, "bn: C must be a multiple of 8");
  const int nb = bn_blocks(M);
  const int64_t rpb = ceil_div64(M, nb);
  bn_bwd_partial<<<nb, kBnThreads, 0, stream>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(y),
                                                reinterpret_cast<const uint4*>(z), save_mean, save_invstd, M, C, relu, rpb,
                                                partial);
  NBC_CHECK_LAUNCH();
  bn_bwd_finalize<<<C, kBnFinThreads, 0, stream>>>(partial, nb, C, sums, dgamma, dbeta);
  NBC_CHECK_LAUNCH();
  const int64_t total8 = M * C / 8;
  bn_bwd_apply_kernel<<<grid_for(total8), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(y), reinterpret_cast<const uint4*>(z), save_mean,
      save_invstd, gamma, sums, (float)(1.0 / (double)M), total8, C / 8, relu, reinterpret_cast<uint4*>(dz),
      reinterpret_cast<uint4*>(g_out));
  NBC_CHECK_LAUNCH();
  return 0;
}

int cast_pack(const float* w, int64_t n, void* out, cudaStream_t stream) {
  cast_pack_kernel<<<grid_for(n), 256, 0, stream>>>(w, n, reinterpret_cast<__nv_bfloat16*>(out));
  NBC_CHECK_LAUNCH();
  return 0;
}
int dgrad_pack(const float* w, int Cout, int Cin, int kh, int kw, void* out, cudaStream_t stream) {
  dgrad_pack_kernel<<<grid_for((int64_t)Cout * Cin * kh * kw), 256, 0, stream>>>(w, Cout, Cin, kh, kw,
                                                                              reinterpret_cast<__nv_bfloat16*>(out));
  NBC_CHECK_LAUNCH();
  return 0;
}
int oihw_ohwi(const float* src, int Cout, int Cin, int kh, int kw, float* dst, int reverse, cudaStream_t stream) {
  oihw_to_ohwi_kernel<<<grid_for((int64_t)Cout * Cin * kh * kw), 256, 0, stream>>>(src, Cout, Cin, kh, kw, dst, reverse);
  NBC_CHECK_LAUNCH();
  return 0;
}

}  // namespace nbc
