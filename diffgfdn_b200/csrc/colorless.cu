// Colorless (spectral flatness) loss of the lossless sub-FDN responses, forward and backward.
// Replaces mse_loss / amse_loss of the reference (diff_gfdn/colorless_fdn/losses.py:20-73) applied per group
// against a target of ones (trainer.py:298-303):  loss[g] = mean_k (|H[k,g]| - 1)^p,  p = 4 where asym and
// |H|-1 > 1, else 2. One block per group, float64 accumulation, fixed reduction order.
#include "common.cuh"

namespace dgfdn {
namespace {
constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads) colorless_fwd_kernel(int g, int64_t k, const float2* __restrict__ h,
                                                                 int asym, double* __restrict__ loss) {
  __shared__ double red[kThreads / 32];
  const int gi = blockIdx.x;
  double acc = 0.0;
  for (int64_t kk = threadIdx.x; kk < k; kk += kThreads) {
    const float2 v = h[kk * g + gi];
    const double d = hypot((double)v.x, (double)v.y) - 1.0;
    const double d2 = d * d;
    acc += (asym && d > 1.0) ? d2 * d2 : d2;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    loss[gi] = s / (double)k;
  }
}

__global__ void colorless_bwd_kernel(int g, int64_t k, const float2* __restrict__ h, int asym,
                                     const double* __restrict__ coef, float2* __restrict__ gh) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * g) return;
  const int gi = (int)(i % g);
  const float2 v = h[i];
  const double a = hypot((double)v.x, (double)v.y);
  const double d = a - 1.0;
  double dfd = (asym && d > 1.0) ? 4.0 * d * d * d : 2.0 * d;
  double s = (a > 0.0) ? coef[gi] * dfd / ((double)k * a) : 0.0;
  gh[i] = make_float2((float)(s * v.x), (float)(s * v.y));
}
}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_colorless_fwd(int g, int64_t k, const void* h_sub, int asym, double* loss, void* stream) {
  DGFDN_CHECK(g >= 1 && k >= 1 && h_sub && loss, "colorless_fwd: bad arguments");
  colorless_fwd_kernel<<<g, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g, k, static_cast<const float2*>(h_sub),
                                                                             asym, loss);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_colorless_bwd(int g, int64_t k, const void* h_sub, int asym, const double* coef, void* gh,
                                   void* stream) {
  DGFDN_CHECK(g >= 1 && k >= 1 && h_sub && coef && gh, "colorless_bwd: bad arguments");
  const int64_t tot = k * g;
  colorless_bwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, k, static_cast<const float2*>(h_sub), asym, coef, static_cast<float2*>(gh));
  DGFDN_LAUNCH_CHECK();
  return 0;
}
