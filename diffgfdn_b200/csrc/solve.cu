// K1 / K1^T: per-bin build + complex LU solve of the Grouped-FDN feedback loop, one warp per bin.
//
//   M_k = diag(z_k^{m_i} / gamma_i) - A      x_k = M_k^{-1} b      y[k,g] = sum_{n in g} c_n x_k[n]
//
// Replaces FeedbackLoop.forward (reference diff_gfdn/feedback_loop.py:326-391: diag_embed + repeat +
// torch.linalg.inv on K dense NxN complex128 matrices) and the einsums of model.py:615-619 / :1083 /
// :237-250. The reference inverts, then contracts with c (per receiver!) and b; b is shared by every
// receiver, so one solve per bin is all that is needed and the (K,N,N) inverse never exists.
//
// Arithmetic is float64 (the reference inverts in complex128, feedback_loop.py:391); outputs are
// complex64 like the reference's P. The matrix lives in shared memory (column-major, one column per
// lane), elimination runs row-per-lane, pivots are found with warp shuffles.
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kWarps = 4;

struct SolveParams {
  int n, g, l;
  int64_t k;
  const double2* z;
  const int32_t* delays;
  const float* a;
  int transpose_a;
  const float* gamma;
  const float2* gamma_z;
  const float* b;
  const float* c;
  // forward outputs
  float2* x;
  float2* y;
  // backward inputs / outputs
  const float2* xin;
  const float2* gy;
  const float2* gx;
  double* ws;
};

// shared-memory carve-up (doubles). Block-wide: A (n*n), invgamma (n), b (n), c (n), delays as double (n).
// Per warp: mat (2*n*n), rhs (2*n), xs (2*n), lam (2*n), acc (n*n, backward only).
// Both sizes are rounded up to an even count so that every double2 array stays 16-byte aligned for odd n.
__host__ __device__ inline size_t block_doubles(int n) { return ((size_t)n * n + 4 * (size_t)n + 1) & ~(size_t)1; }
__host__ __device__ inline size_t warp_doubles(int n, bool bwd) {
  return (2 * (size_t)n * n + 6 * (size_t)n + (bwd ? (size_t)n * n : 0) + 1) & ~(size_t)1;
}

// z^m * invgamma for this lane's delay line, float64. Also returns z^m alone through zm.
__device__ __forceinline__ double2 diag_entry(const SolveParams& p, int64_t bin, int lane, const double* s_invg,
                                              const double* s_delay, double2* zm) {
  double2 zk = p.z[bin];
  double r = hypot(zk.x, zk.y);
  double th = atan2(zk.y, zk.x);
  double m = s_delay[lane];
  double mag = pow(r, m);
  double sn, cs;
  sincos(m * th, &sn, &cs);
  double2 v = make_double2(mag * cs, mag * sn);
  *zm = v;
  if (p.gamma_z != nullptr) {
    float2 gz = p.gamma_z[(int64_t)lane * p.k + bin];
    return cdiv(v, make_double2((double)gz.x, (double)gz.y));
  }
  double ig = s_invg[lane];
  return make_double2(v.x * ig, v.y * ig);
}

// Build column `lane` of M (or of M^H when adjoint) in shared memory. s_a holds the effective A
// (already transposed on load when transpose_a is set), row-major.
__device__ __forceinline__ void build_column(double2* mat, const double* s_a, int n, int lane, double2 dz,
                                             bool adjoint) {
  if (lane < n) {
    for (int r = 0; r < n; ++r) {
      double av = adjoint ? s_a[lane * n + r] : s_a[r * n + lane];
      mat[r + n * lane] = make_double2(-av, 0.0);
    }
    double2 d = mat[lane + n * lane];
    d.x += dz.x;
    d.y += adjoint ? -dz.y : dz.y;
    mat[lane + n * lane] = d;
  }
}

// Gaussian elimination with partial pivoting + back substitution on the warp's shared-memory system.
// On return lane r (< n) holds x_r.
__device__ __forceinline__ double2 warp_solve(double2* mat, double2* rhs, int n, int lane) {
  for (int kk = 0; kk < n; ++kk) {
    double key = (lane >= kk && lane < n) ? cnorm(mat[lane + n * kk]) : -1.0;
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ok = __shfl_xor_sync(0xffffffffu, key, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ok > key || (ok == key && oi < idx)) {
        key = ok;
        idx = oi;
      }
    }
    const int piv = idx;
    if (piv != kk) {
      if (lane >= kk && lane < n) {
        double2 t = mat[kk + n * lane];
        mat[kk + n * lane] = mat[piv + n * lane];
        mat[piv + n * lane] = t;
      }
      if (lane == 0) {
        double2 t = rhs[kk];
        rhs[kk] = rhs[piv];
        rhs[piv] = t;
      }
    }
    __syncwarp();
    const double2 pinv = cinv(mat[kk + n * kk]);
    if (lane > kk && lane < n) {
      const double2 f = cmul(mat[lane + n * kk], pinv);
      for (int j = kk + 1; j < n; ++j) {
        double2 pj = mat[kk + n * j];
        double2 v = mat[lane + n * j];
        v.x -= f.x * pj.x - f.y * pj.y;
        v.y -= f.x * pj.y + f.y * pj.x;
        mat[lane + n * j] = v;
      }
      double2 pr = rhs[kk];
      double2 v = rhs[lane];
      v.x -= f.x * pr.x - f.y * pr.y;
      v.y -= f.x * pr.y + f.y * pr.x;
      rhs[lane] = v;
    }
    __syncwarp();
  }
  double2 mine = make_double2(0.0, 0.0);
  for (int kk = n - 1; kk >= 0; --kk) {
    const double2 xk = cdiv(rhs[kk], mat[kk + n * kk]);
    __syncwarp();
    if (lane < kk) {
      double2 m = mat[lane + n * kk];
      double2 v = rhs[lane];
      v.x -= m.x * xk.x - m.y * xk.y;
      v.y -= m.x * xk.y + m.y * xk.x;
      rhs[lane] = v;
    }
    if (lane == kk) mine = xk;
    __syncwarp();
  }
  return mine;
}

__device__ __forceinline__ void load_block_constants(const SolveParams& p, double* s_a, double* s_invg, double* s_b,
                                                     double* s_c, double* s_delay) {
  const int n = p.n;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    int r = i / n, c = i % n;
    s_a[i] = (double)(p.transpose_a ? p.a[c * n + r] : p.a[i]);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_invg[i] = p.gamma ? 1.0 / (double)p.gamma[i] : 1.0;
    s_b[i] = p.b ? (double)p.b[i] : 0.0;
    s_c[i] = p.c ? (double)p.c[i] : 0.0;
    s_delay[i] = (double)p.delays[i];
  }
}

__global__ void __launch_bounds__(kWarps * 32) solve_fwd_kernel(SolveParams p) {
  extern __shared__ double smem[];
  const int n = p.n;
  double* s_a = smem;
  double* s_invg = s_a + (size_t)n * n;
  double* s_b = s_invg + n;
  double* s_c = s_b + n;
  double* s_delay = s_c + n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + block_doubles(n) + (size_t)warp * warp_doubles(n, false);
  double2* mat = reinterpret_cast<double2*>(wbase);
  double2* rhs = mat + (size_t)n * n;

  load_block_constants(p, s_a, s_invg, s_b, s_c, s_delay);
  __syncthreads();

  for (int64_t bin = (int64_t)blockIdx.x * kWarps + warp; bin < p.k; bin += (int64_t)gridDim.x * kWarps) {
    double2 zm;
    double2 dz = make_double2(0.0, 0.0);
    if (lane < n) dz = diag_entry(p, bin, lane, s_invg, s_delay, &zm);
    build_column(mat, s_a, n, lane, dz, false);
    if (lane < n) rhs[lane] = make_double2(s_b[lane], 0.0);
    __syncwarp();
    double2 xr = warp_solve(mat, rhs, n, lane);
    if (p.x != nullptr && lane < n) p.x[bin * n + lane] = make_float2((float)xr.x, (float)xr.y);
    if (p.y != nullptr) {
      const double cr = (lane < n) ? s_c[lane] : 0.0;
      const int grp = (lane < n) ? lane / p.l : -1;
      for (int gi = 0; gi < p.g; ++gi) {
        double re = warp_sum(grp == gi ? cr * xr.x : 0.0);
        double im = warp_sum(grp == gi ? cr * xr.y : 0.0);
        if (lane == 0) p.y[bin * p.g + gi] = make_float2((float)re, (float)im);
      }
    }
    __syncwarp();
  }
}

// Backward: adjoint solve per bin + accumulation of the parameter gradients. Each block writes one
// row of partial sums to ws; solve_bwd_reduce_kernel adds the rows in a fixed order.
__global__ void __launch_bounds__(kWarps * 32) solve_bwd_kernel(SolveParams p) {
  extern __shared__ double smem[];
  const int n = p.n;
  double* s_a = smem;
  double* s_invg = s_a + (size_t)n * n;
  double* s_b = s_invg + n;
  double* s_c = s_b + n;
  double* s_delay = s_c + n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + block_doubles(n) + (size_t)warp * warp_doubles(n, true);
  double2* mat = reinterpret_cast<double2*>(wbase);
  double2* rhs = mat + (size_t)n * n;
  double2* xs = rhs + n;
  double2* lam = xs + n;
  double* acc = reinterpret_cast<double*>(lam + n);

  load_block_constants(p, s_a, s_invg, s_b, s_c, s_delay);
  for (int i = lane; i < n * n; i += 32) acc[i] = 0.0;
  __syncthreads();

  double gb_acc = 0.0, gc_acc = 0.0, gig_acc = 0.0;
  for (int64_t bin = (int64_t)blockIdx.x * kWarps + warp; bin < p.k; bin += (int64_t)gridDim.x * kWarps) {
    double2 zm = make_double2(0.0, 0.0);
    double2 dz = make_double2(0.0, 0.0);
    double2 xr = make_double2(0.0, 0.0);
    double2 gyr = make_double2(0.0, 0.0);
    if (lane < n) {
      dz = diag_entry(p, bin, lane, s_invg, s_delay, &zm);
      float2 xv = p.xin[bin * n + lane];
      xr = make_double2((double)xv.x, (double)xv.y);
      xs[lane] = xr;
      double2 r = make_double2(0.0, 0.0);
      if (p.gy != nullptr) {
        float2 gv = p.gy[bin * p.g + lane / p.l];
        gyr = make_double2((double)gv.x, (double)gv.y);
        r.x = s_c[lane] * gyr.x;
        r.y = s_c[lane] * gyr.y;
      }
      if (p.gx != nullptr) {
        float2 gv = p.gx[bin * n + lane];
        r.x += (double)gv.x;
        r.y += (double)gv.y;
      }
      rhs[lane] = r;
    }
    build_column(mat, s_a, n, lane, dz, true);
    __syncwarp();
    double2 lr = warp_solve(mat, rhs, n, lane);
    if (lane < n) lam[lane] = lr;
    __syncwarp();
    if (lane < n) {
      // dL/dA_eff[i][j] = Re(lambda_i conj(x_j)); the reduce kernel transposes back when A_eff = A^T.
      for (int j = 0; j < n; ++j) {
        double2 o = xs[j];
        acc[lane + n * j] += lr.x * o.x + lr.y * o.y;
      }
      gb_acc += lr.x;
      gc_acc += xr.x * gyr.x + xr.y * gyr.y;
      // grad wrt dz_i is -lambda_i conj(x_i); dz_i = zm_i * invgamma_i (real invgamma)
      double2 t = cmulc(lr, xr);  // lambda * conj(x)
      gig_acc -= zm.x * t.x + zm.y * t.y;
    }
    __syncwarp();
  }
  // block reduction: warp partials -> ws[blockIdx.x]
  __syncthreads();
  const size_t per = (size_t)n * n + 3 * (size_t)n;
  double* out = p.ws + (size_t)blockIdx.x * per;
  double* acc0 = smem + block_doubles(n) + 2 * (size_t)n * n + 6 * (size_t)n;  // warp 0 acc
  const size_t wstride = warp_doubles(n, true);
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += acc0[(size_t)w * wstride + i];
    out[i] = s;  // layout acc[i + n*j] : row i, col j  ->  index i + n*j
  }
  // per-lane scalars: stash in each warp's rhs area (2n doubles) + xs area
  double* stash = reinterpret_cast<double*>(rhs);  // 6n doubles available (rhs, xs, lam)
  if (lane < n) {
    stash[lane] = gb_acc;
    stash[n + lane] = gc_acc;
    stash[2 * n + lane] = gig_acc;
  }
  __syncthreads();
  double* stash0 = smem + block_doubles(n) + 2 * (size_t)n * n;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += stash0[(size_t)w * wstride + i];
    out[(size_t)n * n + i] = s;
  }
}

__global__ void solve_bwd_reduce_kernel(const double* ws, int nblocks, int n, int transpose_a, double* ga, double* gb,
                                        double* gc, double* gig) {
  const int per = n * n + 3 * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += ws[(size_t)b * per + i];
    if (i < n * n) {
      // partial index i = row + n*col of dL/dA_eff
      int row = i % n, col = i / n;
      if (transpose_a) {
        int t = row;
        row = col;
        col = t;
      }
      if (ga) ga[row * n + col] = s;
    } else {
      int j = i - n * n;
      if (j < n) {
        if (gb) gb[j] = s;
      } else if (j < 2 * n) {
        if (gc) gc[j - n] = s;
      } else {
        if (gig) gig[j - 2 * n] = s;
      }
    }
  }
}

int grid_blocks(int64_t k) {
  int64_t want = (k + kWarps - 1) / kWarps;
  int64_t cap = (int64_t)sm_count() * 4;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

int check_common(int n, int g, int64_t k) {
  DGFDN_CHECK(n >= 1 && n <= DGFDN_MAX_LINES, "solve: n=%d out of range [1,%d]", n, DGFDN_MAX_LINES);
  DGFDN_CHECK(g >= 1 && g <= DGFDN_MAX_GROUPS && n % g == 0, "solve: g=%d must divide n=%d and be <= %d", g, n,
              DGFDN_MAX_GROUPS);
  DGFDN_CHECK(k >= 1, "solve: k=%lld must be positive", (long long)k);
  return 0;
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_solve_fwd(int n, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                               int transpose_a, const float* gamma, const void* gamma_z, const float* b,
                               const float* c, void* x, void* y, void* stream) {
  if (check_common(n, g, k)) return 1;
  DGFDN_CHECK(z && delays && a && b, "solve_fwd: null input pointer");
  DGFDN_CHECK(y == nullptr || c != nullptr, "solve_fwd: y requested without c");
  SolveParams p{};
  p.n = n;
  p.g = g;
  p.l = n / g;
  p.k = k;
  p.z = static_cast<const double2*>(z);
  p.delays = delays;
  p.a = a;
  p.transpose_a = transpose_a;
  p.gamma = gamma;
  p.gamma_z = static_cast<const float2*>(gamma_z);
  p.b = b;
  p.c = c;
  p.x = static_cast<float2*>(x);
  p.y = static_cast<float2*>(y);
  size_t smem = (block_doubles(n) + kWarps * warp_doubles(n, false)) * sizeof(double);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  solve_fwd_kernel<<<grid_blocks(k), kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t dgfdn_solve_bwd_ws_bytes(int n) {
  return (int64_t)sm_count() * 4 * ((int64_t)n * n + 3 * (int64_t)n) * (int64_t)sizeof(double);
}

extern "C" int dgfdn_solve_bwd(int n, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                               int transpose_a, const float* gamma, const void* gamma_z, const float* c,
                               const void* x, const void* gy, const void* gx, double* ga, double* gb, double* gc,
                               double* ginvgamma, void* ws, void* stream) {
  if (check_common(n, g, k)) return 1;
  DGFDN_CHECK(z && delays && a && x && ws, "solve_bwd: null input pointer");
  DGFDN_CHECK(gy || gx, "solve_bwd: need gy or gx");
  DGFDN_CHECK(gy == nullptr || c != nullptr, "solve_bwd: gy given without c");
  SolveParams p{};
  p.n = n;
  p.g = g;
  p.l = n / g;
  p.k = k;
  p.z = static_cast<const double2*>(z);
  p.delays = delays;
  p.a = a;
  p.transpose_a = transpose_a;
  p.gamma = gamma;
  p.gamma_z = static_cast<const float2*>(gamma_z);
  p.b = nullptr;
  p.c = c;
  p.xin = static_cast<const float2*>(x);
  p.gy = static_cast<const float2*>(gy);
  p.gx = static_cast<const float2*>(gx);
  p.ws = static_cast<double*>(ws);
  const int blocks = grid_blocks(k);
  size_t smem = (block_doubles(n) + kWarps * warp_doubles(n, true)) * sizeof(double);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  solve_bwd_kernel<<<blocks, kWarps * 32, smem, st>>>(p);
  DGFDN_LAUNCH_CHECK();
  solve_bwd_reduce_kernel<<<8, 128, 0, st>>>(p.ws, blocks, n, transpose_a, ga, gb, gc, ginvgamma);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
