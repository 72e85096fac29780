// K3e: energy-decay-relief loss on the STFT of the achieved response, forward and backward.
//
//   EDR[f,m] = sum_{m' >= m} |S[f,m']|^2 ;  L = sum_b ( sum_{f,m} |T_dB[b,f,m] - 10 log10(EDR_b[f,m] + eps)| ) / den[b]
//
// Replaces get_edr_from_stft's Python loop over frames, the dB conversion and the normalised L1 reduction of
// edr_loss.forward (reference diff_gfdn/losses.py:447-495, 556-575; utils.py:16-40) and their autograd (abs, pow,
// flip, cumsum, flip, log10, clip, sub, abs, sum, div: ~40 launches and eight (B,F,T_f) temporaries per call).
// The STFT itself stays cuFFT (batched R2C of the hann-windowed frames); S arrives as [rows, T_f, F] complex64, the
// layout torch.fft.rfft produces for frames [rows, T_f, win], and the target EDR is kept in the same layout, so a
// thread owns one (row, frequency) pair, walks the frames from the last to the first with the running energy in a
// register and every access is coalesced along f. HBM: 8 B (S) + 4 B (target) per (row, frame, bin) forward;
// backward adds the 8 B of dL/dS. float64 inside, fixed-order reductions.
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kThreads = 256;
constexpr double kEps = 1.1920928955078125e-07;  // torch.finfo(float32).eps (utils.py:35)
constexpr double kDbFactor = 4.342944819032518;   // 10 / ln(10)

__device__ __forceinline__ double power_of(float2 v) { return (double)v.x * (double)v.x + (double)v.y * (double)v.y; }

__global__ void __launch_bounds__(kThreads) edr_db_kernel(int64_t tf, int64_t f, const float2* __restrict__ s,
                                                          float* __restrict__ out) {
  const int64_t fi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (fi >= f) return;
  const int64_t base = (int64_t)blockIdx.y * tf * f + fi;
  double run = 0.0;
  for (int64_t m = tf - 1; m >= 0; --m) {
    run += power_of(s[base + m * f]);
    const double d = 10.0 * log10(run + kEps);
    out[base + m * f] = (float)(d < -200.0 ? -200.0 : d);
  }
}

// part[row * gridDim.x + blockIdx.x] = sum over this block's bins and all frames of |target - EDR_dB|
__global__ void __launch_bounds__(kThreads) edr_loss_fwd_kernel(int64_t tf, int64_t f, const float2* __restrict__ s,
                                                                const float* __restrict__ tdb,
                                                                double* __restrict__ part) {
  __shared__ double red[kThreads / 32];
  const int64_t fi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t base = (int64_t)blockIdx.y * tf * f + fi;
  double acc = 0.0;
  if (fi < f) {
    double run = 0.0;
    for (int64_t m = tf - 1; m >= 0; --m) {
      run += power_of(s[base + m * f]);
      double d = 10.0 * log10(run + kEps);
      d = d < -200.0 ? -200.0 : d;
      acc += fabs((double)tdb[base + m * f] - d);
    }
  }
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) t += red[i];
    part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void edr_finalize_kernel(int64_t rows, int nblk, const double* __restrict__ part,
                                    const double* __restrict__ den, double* __restrict__ loss) {
  __shared__ double red[kThreads / 32];
  double acc = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += kThreads) {
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += part[r * nblk + b];
    acc += t / den[r];
  }
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) t += red[i];
    loss[0] = t;
  }
}

// gS[b,m,f] = (gloss / den[b]) * 2 S[b,m,f] * sum_{m'' <= m} dL/dEDR[b,f,m''] ,
// dL/dEDR[m''] = -sign(T - E_dB) (10/ln10) / (EDR + eps)  (0 where the dB value is clipped).
// Pass 1 (last frame -> first) parks dL/dEDR in gS.x, pass 2 (first -> last) turns it into the prefix sum.
__global__ void __launch_bounds__(kThreads) edr_loss_bwd_kernel(int64_t tf, int64_t f, const float2* __restrict__ s,
                                                                const float* __restrict__ tdb,
                                                                const double* __restrict__ den,
                                                                const double* __restrict__ gloss,
                                                                float2* __restrict__ gs) {
  const int64_t fi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (fi >= f) return;
  const int64_t base = (int64_t)blockIdx.y * tf * f + fi;
  const double scale = 2.0 * gloss[0] / den[blockIdx.y];
  double run = 0.0;
  for (int64_t m = tf - 1; m >= 0; --m) {
    run += power_of(s[base + m * f]);
    const double x = run + kEps;
    const double d = 10.0 * log10(x);
    double g = 0.0;
    if (d >= -200.0) {
      const double diff = (double)tdb[base + m * f] - d;
      g = (diff > 0.0 ? -1.0 : (diff < 0.0 ? 1.0 : 0.0)) * kDbFactor / x;
    }
    gs[base + m * f].x = (float)g;
  }
  double pre = 0.0;
  for (int64_t m = 0; m < tf; ++m) {
    pre += (double)gs[base + m * f].x;
    const float2 v = s[base + m * f];
    const double w = scale * pre;
    gs[base + m * f] = make_float2((float)(w * (double)v.x), (float)(w * (double)v.y));
  }
}

int check(int64_t rows, int64_t tf, int64_t f) {
  DGFDN_CHECK(rows >= 0 && rows <= 65535 && tf >= 1 && f >= 1, "edr: bad sizes rows=%lld frames=%lld bins=%lld (rows <= 65535)",
              (long long)rows, (long long)tf, (long long)f);
  return 0;
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_edr_db(int64_t rows, int64_t tf, int64_t f, const void* s, float* out_db, void* stream) {
  if (check(rows, tf, f)) return 1;
  DGFDN_CHECK(s && out_db, "edr_db: null pointer");
  if (rows == 0) return 0;
  const dim3 grid((unsigned)((f + kThreads - 1) / kThreads), (unsigned)rows);
  edr_db_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(tf, f, static_cast<const float2*>(s), out_db);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t dgfdn_edr_ws_bytes(int64_t rows, int64_t f) {
  if (rows < 1 || f < 1) return 0;
  return rows * ((f + kThreads - 1) / kThreads) * (int64_t)sizeof(double);
}

extern "C" int dgfdn_edr_loss_fwd(int64_t rows, int64_t tf, int64_t f, const void* s, const float* target_db,
                                  const double* den, double* loss, void* ws, void* stream) {
  if (check(rows, tf, f)) return 1;
  DGFDN_CHECK(s && target_db && den && loss && ws, "edr_loss_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = (int)((f + kThreads - 1) / kThreads);
  if (rows > 0) {
    const dim3 grid((unsigned)nblk, (unsigned)rows);
    edr_loss_fwd_kernel<<<grid, kThreads, 0, st>>>(tf, f, static_cast<const float2*>(s), target_db,
                                                   static_cast<double*>(ws));
    DGFDN_LAUNCH_CHECK();
  }
  edr_finalize_kernel<<<1, kThreads, 0, st>>>(rows, nblk, static_cast<const double*>(ws), den, loss);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_edr_loss_bwd(int64_t rows, int64_t tf, int64_t f, const void* s, const float* target_db,
                                  const double* den, const double* gloss, void* gs, void* stream) {
  if (check(rows, tf, f)) return 1;
  DGFDN_CHECK(s && target_db && den && gloss && gs, "edr_loss_bwd: null pointer");
  if (rows == 0) return 0;
  const dim3 grid((unsigned)((f + kThreads - 1) / kThreads), (unsigned)rows);
  edr_loss_bwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      tf, f, static_cast<const float2*>(s), target_db, den, gloss, static_cast<float2*>(gs));
  DGFDN_LAUNCH_CHECK();
  return 0;
}
