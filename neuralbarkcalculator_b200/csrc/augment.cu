// N4 -- training-time augmentation on the GPU, one fused gather pass per batch (reference __main__.py:153-176
// get_loader_for_crop_batch: pad_resize -> ColorJitter(saturation, brightness) -> RandomCrop -> RandomHorizontalFlip ->
// RandomVerticalFlip -> ToTensor, applied with the SAME random draw to the image and to its label image,
// dataset.py:171-179; then target = round(target * 2), dataset.py:184-193).
//
// The random draws are made by the caller (host); this kernel applies them.  For output sample b, pixel (y, x):
//   geometry : flips and crop offset map (y, x) to a row / column of the reflect-padded source (utils.py:242-247
//              pad_resize with 'reflect' padding; only the case where padding alone reaches the target size is built --
//              an odd difference would need PIL's antialiased resize as well and is rejected by the host wrapper)
//   colour   : PIL's ImageEnhance arithmetic in explicitly rounded f32 operations, in the drawn order:
//                brightness: out = blend(0, v, f)              saturation: out = blend(L, v, f),
//                L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16    (PIL's ITU-R 601-2 luma)
//                blend(a, v, f) = (uint8) trunc(a + f * (v - a)), clipped to [0, 255] when f is outside [0, 1]
//   label    : the dual image value (0 / 127 / 255) goes through the SAME brightness blend (saturation leaves a grey
//              image unchanged), then class = round(value / 255 * 2) as dataset.py:190-193
// Bytes: reads and writes 4 bytes per output pixel -- a pure HBM gather; the crop is tiny next to a training step.
#include "common.cuh"

namespace nbc {

struct AugmentParams {   // one per output sample, in device memory (matches nbc_augment_params in nbc.h)
  int32_t src, x0, y0, hflip, vflip, order;   // order: 0 = brightness then saturation, 1 = saturation then brightness
  float brightness, saturation;               // factor; <= 0 disables (a factor of exactly 1 is applied like PIL does)
};

__device__ __forceinline__ int pil_blend(int a, int v, float f) {
  const float t = __fadd_rn(__int2float_rn(a), __fmul_rn(f, __int2float_rn(v - a)));
  if (f >= 0.f && f <= 1.f) return (int)t;          // interpolation: plain truncation (Pillow Blend.c)
  if (t <= 0.f) return 0;
  if (t >= 255.f) return 255;
  return (int)t;
}

__device__ __forceinline__ int reflect_index(int i, int n) {   // numpy / torchvision 'reflect': edge not repeated
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}

__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ images, const uint8_t* __restrict__ duals,
                                                      int Hs, int Ws, int pad_top, int pad_left, int crop,
                                                      const AugmentParams* __restrict__ params, uint8_t* __restrict__ out_img,
                                                      uint8_t* __restrict__ out_cls) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= crop) return;
  const AugmentParams p = params[b];
  // output (y, x) <- flipped crop coordinates <- padded image coordinates <- source pixel
  const int cx = p.hflip ? crop - 1 - x : x, cy = p.vflip ? crop - 1 - y : y;
  const int sx = reflect_index(p.x0 + cx - pad_left, Ws), sy = reflect_index(p.y0 + cy - pad_top, Hs);
  const int64_t spix = ((int64_t)p.src * Hs + sy) * Ws + sx;
  const uint8_t* s = images + spix * 3;
  int r = s[0], g = s[1], bl = s[2];
  int d = duals != nullptr ? (int)duals[spix] : 0;
#pragma unroll
  for (int step = 0; step < 2; ++step) {
    const bool do_brightness = (step == 0) == (p.order == 0);
    if (do_brightness) {
      if (p.brightness > 0.f) {
        r = pil_blend(0, r, p.brightness), g = pil_blend(0, g, p.brightness), bl = pil_blend(0, bl, p.brightness);
        d = pil_blend(0, d, p.brightness);
      }
    } else if (p.saturation > 0.f) {
      const int L = (19595 * r + 38470 * g + 7471 * bl + 0x8000) >> 16;
      r = pil_blend(L, r, p.saturation), g = pil_blend(L, g, p.saturation), bl = pil_blend(L, bl, p.saturation);
    }
  }
  const int64_t opix = ((int64_t)b * crop + y) * crop + x;
  uint8_t* o = out_img + opix * 3;
  o[0] = (uint8_t)r, o[1] = (uint8_t)g, o[2] = (uint8_t)bl;
  if (out_cls != nullptr) {
    // ToTensor: d / 255 (f32); dataset.py:190-191: * 2, round (half to even, torch.round_)
    const float t = __fmul_rn(__fdiv_rn(__int2float_rn(d), 255.f), 2.f);
    out_cls[opix] = (uint8_t)__float2int_rn(t);
  }
}

}  // namespace nbc

using namespace nbc;

extern "C" int nbc_augment_batch(const uint8_t* images, const uint8_t* duals, int M, int Hs, int Ws, int target_h, int target_w,
                                 int crop, const nbc_augment_params* params, int B, uint8_t* out_images, uint8_t* out_classes,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static_assert(sizeof(nbc_augment_params) == sizeof(AugmentParams), "parameter struct mismatch");
  NBC_REQUIRE(images && params && out_images, "nbc_augment_batch: null pointer");
  NBC_REQUIRE((duals == nullptr) == (out_classes == nullptr), "nbc_augment_batch: duals and out_classes go together");
  NBC_REQUIRE(M > 0 && Hs > 0 && Ws > 0 && B > 0 && crop > 0 && B <= 65535 && crop <= 65535, "nbc_augment_batch: bad shape");
  NBC_REQUIRE(target_h >= Hs && target_w >= Ws && (target_h - Hs) % 2 == 0 && (target_w - Ws) % 2 == 0,
              "nbc_augment_batch: pad_resize(%d, %d) of a %dx%d image needs a real resize (odd or negative difference); only "
              "reflect padding is built", target_w, target_h, Ws, Hs);
  NBC_REQUIRE(crop <= target_h && crop <= target_w, "nbc_augment_batch: crop %d larger than the padded image", crop);
  NBC_REQUIRE((target_h - Hs) / 2 < Hs && (target_w - Ws) / 2 < Ws, "nbc_augment_batch: reflect padding wider than the image");
  dim3 grid(ceil_div(crop, 256), crop, B);
  augment_kernel<<<grid, 256, 0, stream>>>(images, duals, Hs, Ws, (target_h - Hs) / 2, (target_w - Ws) / 2, crop,
                                           reinterpret_cast<const AugmentParams*>(params), out_images, out_classes);
  NBC_CHECK_LAUNCH();
  return 0;
}
