// K2 (general-shape path) -- implicit-GEMM convolution with warp-level mma.sync.m16n8k16 bf16.
//
// Handles what the tcgen05 kernel does not (arbitrary kernel size / stride, Cin % 32 == 0) and serves as the
// on-device cross-check of the tcgen05 kernel in tests.  Same math as conv_tc.cu:
//   y = act(conv(x, w) + bias (+ residual)),  x NHWC bf16, w [Cout][kh][kw][Cin] bf16, f32 accumulate.
// CTA tile 128 pixels x 64 channels, K chunk 32, cp.async double buffering with zero-fill for the padding.
#include "common.cuh"
#include "conv.h"

namespace nbc {

struct ConvMmaParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  int N, H, W, Cin, Cout, Ho, Wo;
  int kh, kw, stride, pad, dil, relu, f16;
  int64_t M;  // N*Ho*Wo
};

constexpr int MM_BM = 128, MM_BN = 64, MM_BK = 32, MM_LD = 40;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) conv_mma_kernel(const ConvMmaParams p) {
  __shared__ __align__(16) __nv_bfloat16 sA[2][MM_BM][MM_LD];
  __shared__ __align__(16) __nv_bfloat16 sB[2][MM_BN][MM_LD];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps, warp tile 32 x 32
  const int64_t m0 = (int64_t)blockIdx.x * MM_BM;
  const int n0 = blockIdx.y * MM_BN;

  // the two A rows (pixels) this thread copies, fixed for the whole K loop
  int a_img[2], a_ho[2], a_wo[2];
  bool a_ok[2];
  const int a_part = tid & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = (tid >> 2) + i * 64;
    const int64_t m = m0 + row;
    a_ok[i] = m < p.M;
    const int64_t mm = a_ok[i] ? m : 0;
    a_wo[i] = (int)(mm % p.Wo);
    const int64_t t = mm / p.Wo;
    a_ho[i] = (int)(t % p.Ho);
    a_img[i] = (int)(t / p.Ho);
  }
  const int b_row = tid >> 2, b_part = tid & 3;
  const int cchunks = p.Cin / MM_BK;
  const int kchunks = p.kh * p.kw * cchunks;
  const int64_t Ktot = (int64_t)p.kh * p.kw * p.Cin;

  auto load_chunk = [&](int kc, int buf) {
    const int tap = kc / cchunks, c0 = (kc - tap * cchunks) * MM_BK;
    const int ky = tap / p.kw, kx = tap - ky * p.kw;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = (tid >> 2) + i * 64;
      const int hi = a_ho[i] * p.stride + ky * p.dil - p.pad;
      const int wi = a_wo[i] * p.stride + kx * p.dil - p.pad;
      const bool ok = a_ok[i] && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
      const __nv_bfloat16* src =
          ok ? p.x + (((int64_t)a_img[i] * p.H + hi) * p.W + wi) * p.Cin + c0 + a_part * 8 : p.x;
      cp_async16(smem_u32(&sA[buf][row][a_part * 8]), src, ok ? 16 : 0);
    }
    const __nv_bfloat16* srcb = p.w + (int64_t)(n0 + b_row) * Ktot + (int64_t)tap * p.Cin + c0 + b_part * 8;
    cp_async16(smem_u32(&sB[buf][b_row][b_part * 8]), srcb, 16);
    cp_async_commit();
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  load_chunk(0, 0);
  for (int kc = 0; kc < kchunks; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < kchunks) {
      load_chunk(kc + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < MM_BK; ks += 16) {
      uint32_t a[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = wm * 32 + i * 16 + (lane & 15);
        const int c = ks + (lane >> 4) * 8;
        ldmatrix_x4(smem_u32(&sA[buf][r][c]), a[i][0], a[i][1], a[i][2], a[i][3]);
      }
      uint32_t b[4][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = wn * 32 + j * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = ks + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(smem_u32(&sB[buf][r][c]), b[2 * j][0], b[2 * j][1], b[2 * j + 1][0], b[2 * j + 1][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (p.f16)
            mma_f16(acc[i][j], a[i], b[j][0], b[j][1]);
          else
            mma_bf16(acc[i][j], a[i], b[j][0], b[j][1]);
        }
    }
    __syncthreads();
  }

  // epilogue: c0,c1 -> (row = lane/4, col = 2*(lane%4)), c2,c3 -> row + 8
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = wm * 32 + i * 16 + (lane >> 2) + half * 8;
      const int64_t m = m0 + row;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + wn * 32 + j * 8 + (lane & 3) * 2;
        float v0 = acc[i][j][half * 2] + __ldg(p.bias + col);
        float v1 = acc[i][j][half * 2 + 1] + __ldg(p.bias + col + 1);
        const int64_t off = m * p.Cout + col;
        if (p.residual) {
          const uint32_t rv = __ldg(reinterpret_cast<const uint32_t*>(p.residual + off));
          v0 += lo16(rv, p.f16);
          v1 += hi16(rv, p.f16);
        }
        if (p.relu) {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        }
        *reinterpret_cast<uint32_t*>(p.y + off) = pack16x2(v0, v1, p.f16);
      }
    }
}

bool conv_mma_supported(const ConvGeom& g) { return g.Cin % 32 == 0 && g.Cout % 64 == 0 && g.Ho() > 0 && g.Wo() > 0; }

int conv_mma(const ConvGeom& g, const void* x, const void* w, const float* bias, const void* residual, void* y,
             cudaStream_t stream) {
  if (!conv_mma_supported(g)) {
    set_error("conv_mma: unsupported shape Cin=%d Cout=%d", g.Cin, g.Cout);
    return NBC_ERR_INVALID;
  }
  ConvMmaParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.N = g.N, p.H = g.H, p.W = g.W, p.Cin = g.Cin, p.Cout = g.Cout, p.Ho = g.Ho(), p.Wo = g.Wo();
  p.kh = g.kh, p.kw = g.kw, p.stride = g.stride, p.pad = g.pad, p.dil = g.dil, p.relu = g.relu, p.f16 = g.f16;
  p.M = (int64_t)g.N * p.Ho * p.Wo;
  dim3 grid((unsigned)ceil_div64(p.M, MM_BM), (unsigned)(g.Cout / MM_BN));
  conv_mma_kernel<<<grid, 256, 0, stream>>>(p);
  NBC_CHECK_LAUNCH();
  return 0;
}

}  // namespace nbc
