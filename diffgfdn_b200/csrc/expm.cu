// Orthogonal parametrisation of the per-group mixing matrices: U_g = expm(triu(M_g,1) - triu(M_g,1)^T), forward
// and backward, one CTA per matrix, no host synchronisation.
//
// Replaces Skew + MatrixExponential (= torch.matrix_exp) of the reference (diff_gfdn/feedback_loop.py:16-36, 270,
// 401-403). torch.matrix_exp picks its Pade/Taylor degree from a norm it reads back on the HOST, i.e. it blocks
// the stream twice per training step (forward and Frechet-derivative backward) and cannot be captured in a CUDA
// graph. Here the scaling exponent is chosen on the device: B = A / 2^s with ||B||_1 <= 1/2, a degree-14 Taylor
// polynomial in Horner form (remainder < 2e-17 ||B||), then s squarings, all in float64 in shared memory.
// The backward uses the block identity  expm([[A^T, G], [0, A^T]]) = [[e^{A^T}, L(A^T, G)], [0, e^{A^T}]]  for the
// adjoint of the Frechet derivative and folds in the adjoint of the skew map: dL/dM = triu(Q - Q^T, 1), Q = L(A^T,G).
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kMaxDim = 32;  // 2 L <= 32
constexpr int kLd = kMaxDim + 1;
constexpr int kTaylor = 14;

// P <- expm(X) for the n x n matrix in s_x (row stride kLd); s_p receives the result, s_t is scratch.
// All 32 x 32 threads of the CTA call this; thread (i, j) owns element (i, j).
__device__ void expm_block(int n, double* s_x, double* s_p, double* s_t, double* s_red) {
  const int i = threadIdx.y, j = threadIdx.x;
  const bool in = i < n && j < n;
  // ||X||_1 = max column sum; thread row 0 scans column j
  if (i == 0) {
    double c = 0.0;
    if (j < n)
      for (int r = 0; r < n; ++r) c += fabs(s_x[r * kLd + j]);
    s_red[j] = c;
  }
  __syncthreads();
  double norm = 0.0;
  for (int c = 0; c < n; ++c) norm = fmax(norm, s_red[c]);
  int s = 0;
  if (norm > 0.5) {
    int e;
    frexp(norm / 0.5, &e);  // norm/0.5 = f 2^e, f in [0.5, 1)  =>  2^e >= norm/0.5
    s = e;
  }
  const double scale = ldexp(1.0, -s);
  if (in) {
    s_x[i * kLd + j] *= scale;
    s_p[i * kLd + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  // Horner: P = I + B P / q, q = kTaylor .. 1
  for (int q = kTaylor; q >= 1; --q) {
    double acc = 0.0;
    if (in) {
      for (int k = 0; k < n; ++k) acc = fma(s_x[i * kLd + k], s_p[k * kLd + j], acc);
      acc = acc / (double)q + ((i == j) ? 1.0 : 0.0);
    }
    __syncthreads();
    if (in) s_p[i * kLd + j] = acc;
    __syncthreads();
  }
  for (int r = 0; r < s; ++r) {
    double acc = 0.0;
    if (in)
      for (int k = 0; k < n; ++k) acc = fma(s_p[i * kLd + k], s_p[k * kLd + j], acc);
    __syncthreads();
    if (in) s_p[i * kLd + j] = acc;
    __syncthreads();
  }
  (void)s_t;
}

__global__ void __launch_bounds__(kMaxDim* kMaxDim) skew_expm_fwd_kernel(int l, const float* __restrict__ m,
                                                                         float* __restrict__ u) {
  __shared__ double s_x[kMaxDim * kLd], s_p[kMaxDim * kLd], s_red[kMaxDim];
  const int i = threadIdx.y, j = threadIdx.x;
  const float* mg = m + (size_t)blockIdx.x * l * l;
  if (i < l && j < l) {
    double v = 0.0;
    if (j > i) v = (double)mg[i * l + j];
    if (j < i) v = -(double)mg[j * l + i];
    s_x[i * kLd + j] = v;
  }
  __syncthreads();
  expm_block(l, s_x, s_p, nullptr, s_red);
  if (i < l && j < l) u[(size_t)blockIdx.x * l * l + i * l + j] = (float)s_p[i * kLd + j];
}

__global__ void __launch_bounds__(kMaxDim* kMaxDim) skew_expm_bwd_kernel(int l, const float* __restrict__ m,
                                                                         const float* __restrict__ gu,
                                                                         float* __restrict__ gm) {
  __shared__ double s_x[kMaxDim * kLd], s_p[kMaxDim * kLd], s_red[kMaxDim];
  const int i = threadIdx.y, j = threadIdx.x;
  const int n = 2 * l;
  const float* mg = m + (size_t)blockIdx.x * l * l;
  const float* gg = gu + (size_t)blockIdx.x * l * l;
  if (i < n && j < n) {
    // [[A^T, G], [0, A^T]] with A = T - T^T, T = triu(M, 1):  A^T[i][j] = A[j][i]
    double v = 0.0;
    const int bi = i % l, bj = j % l;
    if ((i < l) == (j < l)) {
      if (bi > bj) v = (double)mg[bj * l + bi];    // A[bj][bi], bj < bi
      if (bi < bj) v = -(double)mg[bi * l + bj];   // A[bj][bi] = -M[bi][bj], bj > bi
    } else if (i < l) {
      v = (double)gg[bi * l + bj];
    }
    s_x[i * kLd + j] = v;
  }
  __syncthreads();
  expm_block(n, s_x, s_p, nullptr, s_red);
  // Q = upper-right block = dL/dA; dL/dM = triu(Q - Q^T, 1)
  if (i < l && j < l) {
    double v = 0.0;
    if (j > i) v = s_p[i * kLd + (l + j)] - s_p[j * kLd + (l + i)];
    gm[(size_t)blockIdx.x * l * l + i * l + j] = (float)v;
  }
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_skew_expm_fwd(int g, int l, const float* m, float* u, void* stream) {
  DGFDN_CHECK(g >= 1 && l >= 1 && m && u, "skew_expm_fwd: bad arguments");
  DGFDN_CHECK(2 * l <= kMaxDim, "skew_expm_fwd: l=%d exceeds %d", l, kMaxDim / 2);
  skew_expm_fwd_kernel<<<g, dim3(kMaxDim, kMaxDim), 0, static_cast<cudaStream_t>(stream)>>>(l, m, u);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_skew_expm_bwd(int g, int l, const float* m, const float* gu, float* gm, void* stream) {
  DGFDN_CHECK(g >= 1 && l >= 1 && m && gu && gm, "skew_expm_bwd: bad arguments");
  DGFDN_CHECK(2 * l <= kMaxDim, "skew_expm_bwd: l=%d exceeds %d", l, kMaxDim / 2);
  skew_expm_bwd_kernel<<<g, dim3(kMaxDim, kMaxDim), 0, static_cast<cudaStream_t>(stream)>>>(l, m, gu, gm);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
