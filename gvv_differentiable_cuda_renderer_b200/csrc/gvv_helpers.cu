// gvv_helpers.cu -- the loss-side helpers that sit right next to the op.
//
//   gvv_gaussian_smooth   python/utils/GaussianSmoothingGpu.py:12-37 (smoothImage): depthwise (2s+1)^2 Gaussian
//                         with zero "SAME" padding over [V,H,W,3] images.  The 2-D kernel there is
//                         outer(vals, vals) / sum, i.e. separable into two normalised 1-D passes.
//   gvv_image_gradient    imageGradient (cpp/src/Utils/RendererUtil.h:566-620): the target-image gradient the
//                         backward's model-to-data term evaluates per covered pixel on EVERY call although the
//                         target is constant during a fit; computed once here, handed to the backward with
//                         gvv_set_target_gradient.
#include "gvv_internal.h"

namespace gvv {

constexpr int kMaxTaps = 65;
struct Taps { float w[kMaxTaps]; int half; };

// one thread per float of the image; dir = 0: along x (stride 3 floats), 1: along y (stride 3*W floats)
__global__ void __launch_bounds__(256)
gauss_pass_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int W, int H, int dir, Taps t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long pix = i / 3;
  const int x = (int)(pix % W), y = (int)((pix / W) % H);
  const int pos = dir == 0 ? x : y, lim = dir == 0 ? W : H;
  const long long stride = dir == 0 ? 3 : 3ll * W;
  float s = 0.f;
  for (int k = -t.half; k <= t.half; ++k) {
    const int q = pos + k;
    if (q >= 0 && q < lim) s = fmaf(t.w[k + t.half], __ldg(in + i + k * stride), s);   // cross-correlation, zero padding
  }
  out[i] = s;
}

// one thread per pixel of one view; same window, weights and normalisation as RendererUtil.h:566-620
__global__ void __launch_bounds__(256)
image_gradient_kernel(const float* __restrict__ img, float* __restrict__ du, float* __restrict__ dv, long long nPix, int W, int H, int fs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nPix) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  float ux = 0.f, uy = 0.f, uz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
  if (x >= fs + 1 && y >= fs + 1 && x < W - (fs + 1) && y < H - (fs + 1)) {
    float norm = 0.f;
    for (int yy = -fs; yy <= fs; ++yy)
      for (int xx = -fs; xx <= fs; ++xx) {
        const float* I = img + 3 * (i + (long long)yy * W + xx);
        const float den = (float)(xx * xx + yy * yy);
        float Gu = 0.f, Gv = 0.f;
        if (den != 0.f) { Gu = (float)xx / den; Gv = (float)yy / den; }
        const float r = __ldg(I), g = __ldg(I + 1), b = __ldg(I + 2);
        ux += Gu * r; uy += Gu * g; uz += Gu * b;
        vx += Gv * r; vy += Gv * g; vz += Gv * b;
        norm += fabsf(Gu);
      }
    const float inorm = 1.f / norm;
    ux *= inorm; uy *= inorm; uz *= inorm; vx *= inorm; vy *= inorm; vz *= inorm;
  }
  du[3 * i] = ux; du[3 * i + 1] = uy; du[3 * i + 2] = uz;
  dv[3 * i] = vx; dv[3 * i + 1] = vy; dv[3 * i + 2] = vz;
}

}  // namespace gvv

using namespace gvv;

extern "C" int gvv_gaussian_smooth(int32_t device, int64_t images, int32_t height, int32_t width, int32_t half_size,
                                   const float* taps, const float* in, float* tmp, float* out, void* stream) {
  if (images < 0 || height <= 0 || width <= 0 || half_size < 0 || 2 * half_size + 1 > kMaxTaps || !taps || !in || !tmp || !out) return GVV_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return GVV_ECUDA;
  Taps t;
  t.half = half_size;
  for (int i = 0; i < 2 * half_size + 1; ++i) t.w[i] = taps[i];
  const long long n = (long long)images * height * width * 3;
  if (n == 0) return GVV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 255) / 256);
  gauss_pass_kernel<<<grid, 256, 0, st>>>(in, tmp, n, width, height, 0, t);
  gauss_pass_kernel<<<grid, 256, 0, st>>>(tmp, out, n, width, height, 1, t);
  return cudaGetLastError() == cudaSuccess ? GVV_OK : GVV_ECUDA;
}

extern "C" int gvv_image_gradient(int32_t device, int64_t images, int32_t height, int32_t width, int32_t filter_size,
                                  const float* image, float* d_du, float* d_dv, void* stream) {
  if (images < 0 || height <= 0 || width <= 0 || filter_size < 0 || !image || !d_du || !d_dv) return GVV_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return GVV_ECUDA;
  const long long n = (long long)images * height * width;
  if (n == 0) return GVV_OK;
  image_gradient_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(image, d_du, d_dv, n, width, height, filter_size);
  return cudaGetLastError() == cudaSuccess ? GVV_OK : GVV_ECUDA;
}
