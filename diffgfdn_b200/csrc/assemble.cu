// Coupled feedback matrix of the Grouped FDN and its adjoint, one launch each:
//
//   Phi = ND_Unitary(clamp(alpha, -pi, pi))      G(G-1)/2 Givens angles, U_n = R_{n-2} ... R_0 [[U_{n-1}, 0], [0, 1]],
//                                                R_i rotating the (i, n-1) plane (feedback_loop.py:39-87, 406-412)
//   A[iL+a, jL+b] = Phi[i,j] (U_i U_j)[a,b]      block_M o (Phi (x) 1_{LxL}), diagonal blocks U_i^2 (quirk Q4)
//                                                (feedback_loop.py:393-404, 424-455)
//
// The reference builds this with ~45 tiny torch kernels forward and ~90 backward (entry-wise rotation matrices, a chain
// of GxG matmuls, kron, einsum); on a 3 ms training step that is a tenth of a millisecond of launch-bound work. Here
// everything is float64 inside one CTA. The adjoint differentiates the rotation product in forward mode, one thread
// per angle (G <= 8: at most 28 angles and 8x8 matrices in local memory), and takes dL/dA in float64 from the K1
// adjoint: dL/dalpha cancels to ~1e-4 of the size of its terms.
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kMaxG = DGFDN_MAX_GROUPS;
constexpr double kPi = (double)3.14159274101257324f;  // torch.clamp on the float32 angles: the bound is float32(pi)

// Phi (and, when t >= 0, dPhi/dalpha_t) of ND_Unitary; row-major G x G in `phi` / `dphi`.
__device__ void nd_unitary(const float* alpha, int g, int t, double* phi, double* dphi) {
  double x[kMaxG * kMaxG], dx[kMaxG * kMaxG];
  for (int i = 0; i < kMaxG * kMaxG; ++i) x[i] = dx[i] = 0.0;
  x[0] = 1.0;  // U_1 = [1]
  for (int n = 2; n <= g; ++n) {
    // embed: [[U_{n-1}, 0], [0, 1]] (tangent: zero corner). Stored with leading dimension kMaxG, so only the corner moves.
    for (int i = 0; i < n - 1; ++i) {
      x[i * kMaxG + n - 1] = x[(n - 1) * kMaxG + i] = 0.0;
      dx[i * kMaxG + n - 1] = dx[(n - 1) * kMaxG + i] = 0.0;
    }
    x[(n - 1) * kMaxG + n - 1] = 1.0;
    dx[(n - 1) * kMaxG + n - 1] = 0.0;
    const int start = (n - 1) * (n - 2) / 2;
    for (int i = 0; i < n - 1; ++i) {  // R_0 first: rot = R_i @ rot
      double a = (double)alpha[start + i];
      const bool clamped = a < -kPi || a > kPi;
      a = a < -kPi ? -kPi : (a > kPi ? kPi : a);
      const double cs = cos(a), sn = sin(a);
      const bool mine = (start + i == t) && !clamped;
      for (int col = 0; col < n; ++col) {
        const double ri = x[i * kMaxG + col], rl = x[(n - 1) * kMaxG + col];
        const double di = dx[i * kMaxG + col], dl = dx[(n - 1) * kMaxG + col];
        x[i * kMaxG + col] = cs * ri - sn * rl;
        x[(n - 1) * kMaxG + col] = sn * ri + cs * rl;
        dx[i * kMaxG + col] = cs * di - sn * dl + (mine ? (-sn * ri - cs * rl) : 0.0);
        dx[(n - 1) * kMaxG + col] = sn * di + cs * dl + (mine ? (cs * ri - sn * rl) : 0.0);
      }
    }
  }
  for (int i = 0; i < g; ++i)
    for (int j = 0; j < g; ++j) {
      if (phi) phi[i * g + j] = x[i * kMaxG + j];
      if (dphi) dphi[i * g + j] = dx[i * kMaxG + j];
    }
}

__global__ void __launch_bounds__(256) coupled_feedback_fwd_kernel(int g, int l, const float* __restrict__ u,
                                                                   const float* __restrict__ alpha,
                                                                   double* __restrict__ a_out, double* __restrict__ phi_out) {
  __shared__ double s_phi[kMaxG * kMaxG];
  extern __shared__ double s_u[];  // [g][l][l]
  const int n = g * l;
  for (int i = threadIdx.x; i < g * l * l; i += blockDim.x) s_u[i] = (double)u[i];
  if (threadIdx.x == 0) nd_unitary(alpha, g, -1, s_phi, nullptr);
  __syncthreads();
  for (int i = threadIdx.x; i < g * g; i += blockDim.x) phi_out[i] = s_phi[i];
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int r = e / n, c = e % n;
    const int bi = r / l, aa = r % l, bj = c / l, bb = c % l;
    double acc = 0.0;
    for (int k = 0; k < l; ++k) acc += s_u[(bi * l + aa) * l + k] * s_u[(bj * l + k) * l + bb];
    a_out[e] = s_phi[bi * g + bj] * acc;
  }
}

__global__ void __launch_bounds__(256) coupled_feedback_bwd_kernel(int g, int l, const float* __restrict__ u,
                                                                   const float* __restrict__ alpha,
                                                                   const double* __restrict__ ga, float* __restrict__ gu,
                                                                   float* __restrict__ galpha) {
  __shared__ double s_phi[kMaxG * kMaxG], s_gphi[kMaxG * kMaxG];
  extern __shared__ double s_dyn[];  // u [g l l] | ga [n n]
  const int n = g * l, na = g * (g - 1) / 2;
  double* s_u = s_dyn;
  double* s_ga = s_dyn + g * l * l;
  for (int i = threadIdx.x; i < g * l * l; i += blockDim.x) s_u[i] = (double)u[i];
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) s_ga[i] = ga[i];
  if (threadIdx.x == 0) nd_unitary(alpha, g, -1, s_phi, nullptr);
  __syncthreads();
  // dL/dPhi[i,j] = sum_ab gA[iL+a, jL+b] (U_i U_j)[a,b]
  for (int e = threadIdx.x; e < g * g; e += blockDim.x) {
    const int bi = e / g, bj = e % g;
    double acc = 0.0;
    for (int aa = 0; aa < l; ++aa)
      for (int bb = 0; bb < l; ++bb) {
        double prod = 0.0;
        for (int k = 0; k < l; ++k) prod += s_u[(bi * l + aa) * l + k] * s_u[(bj * l + k) * l + bb];
        acc += s_ga[(bi * l + aa) * n + bj * l + bb] * prod;
      }
    s_gphi[e] = acc;
  }
  // dL/dU_m[p,q] = sum_j Phi[m,j] sum_b gA[mL+p, jL+b] U_j[q,b] + sum_i Phi[i,m] sum_a U_i[a,p] gA[iL+a, mL+q]
  if (gu != nullptr) {
    for (int e = threadIdx.x; e < g * l * l; e += blockDim.x) {
      const int m = e / (l * l), p = (e / l) % l, q = e % l;
      double acc = 0.0;
      for (int j = 0; j < g; ++j) {
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < l; ++k) {
          t1 += s_ga[(m * l + p) * n + j * l + k] * s_u[(j * l + q) * l + k];
          t2 += s_u[(j * l + k) * l + p] * s_ga[(j * l + k) * n + m * l + q];
        }
        acc += s_phi[m * g + j] * t1 + s_phi[j * g + m] * t2;
      }
      gu[e] = (float)acc;
    }
  }
  __syncthreads();
  if (galpha != nullptr && threadIdx.x < na) {
    double dphi[kMaxG * kMaxG];
    nd_unitary(alpha, g, threadIdx.x, nullptr, dphi);
    double acc = 0.0;
    for (int i = 0; i < g * g; ++i) acc += s_gphi[i] * dphi[i];
    galpha[threadIdx.x] = (float)acc;
  }
}

int check(int g, int l) {
  DGFDN_CHECK(g >= 1 && g <= kMaxG && l >= 1 && g * l <= DGFDN_MAX_LINES, "coupled_feedback: %d groups of %d lines out of range", g,
              l);
  return 0;
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_coupled_feedback_fwd(int g, int l, const float* u, const float* alpha, double* a, double* phi,
                                          void* stream) {
  if (check(g, l)) return 1;
  DGFDN_CHECK(u && a && phi && (alpha || g == 1), "coupled_feedback_fwd: null pointer");
  const size_t smem = (size_t)g * l * l * sizeof(double);
  coupled_feedback_fwd_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(g, l, u, alpha, a, phi);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_coupled_feedback_bwd(int g, int l, const float* u, const float* alpha, const double* ga, float* gu,
                                          float* galpha, void* stream) {
  if (check(g, l)) return 1;
  DGFDN_CHECK(u && ga && (alpha || g == 1), "coupled_feedback_bwd: null pointer");
  const size_t smem = ((size_t)g * l * l + (size_t)g * l * g * l) * sizeof(double);
  coupled_feedback_bwd_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(g, l, u, alpha, ga, gu, galpha);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
