// K1 / K1^T: per-bin build + complex solve of the Grouped-FDN feedback loop, one warp per bin.
//
//   M_k = diag(z_k^{m_i} / gamma_i) - A      x_k = M_k^{-1} b      y[k,g] = sum_{n in g} c_n x_k[n]
//
// Replaces FeedbackLoop.forward (reference diff_gfdn/feedback_loop.py:326-391: diag_embed + repeat +
// torch.linalg.inv on K dense NxN complex128 matrices) and the einsums of model.py:615-619 / :1083 /
// :237-250. The reference inverts, then contracts with c (per receiver!) and b; b is shared by every
// receiver, so one solve per bin is all that is needed and the (K,N,N) inverse never exists.
//
// Layout: lane i of the warp owns row i of the system in REGISTERS (NP complex doubles, NP = N rounded up to a
// multiple of 4, all loops unrolled so every index is static). Gauss-Jordan elimination with partial pivoting:
// the pivot lane is found with a shuffle arg-max, it publishes its row through a small shared-memory line and
// every other lane eliminates against the broadcast values. Rows are never swapped -- a lane remembers which
// column it was pivot for and ends up holding that component of the solution.
// Arithmetic is float64 (the reference inverts in complex128, feedback_loop.py:391); outputs are complex64 like
// the reference's P. The adjoint kernel solves M^H lambda = g the same way and reduces the parameter gradients
// with a fixed-order two-stage reduction (deterministic).
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

// Shared memory (in doubles). Block-wide constants: A_eff row-major [NP*NP], A_eff transposed [NP*NP],
// invgamma/b/c/delay [4*NP]. Per warp: pivot row line 2*(NP+1), solution line 2*NP, saved-x line 2*NP and (backward)
// the gradient accumulator NP*NP.
template <int NP>
struct Smem {
  static constexpr size_t kBlock = 2 * (size_t)NP * NP + 4 * NP;
  static constexpr size_t kWarpFwd = 2 * (NP + 1) + 4 * NP;
  static constexpr size_t kWarpBwd = kWarpFwd + (size_t)NP * NP;
  static_assert(kBlock % 2 == 0 && kWarpFwd % 2 == 0 && kWarpBwd % 2 == 0, "double2 alignment");
};

// z^m * invgamma for this lane's delay line, float64. Also returns z^m alone through zm.
__device__ __forceinline__ double2 diag_entry(const SolveParams& p, int64_t bin, int line, double invg, double delay,
                                              double2* zm) {
  const double2 zk = p.z[bin];
  const double r = hypot(zk.x, zk.y);
  const double th = atan2(zk.y, zk.x);
  const double mag = (fabs(r - 1.0) < 4e-16) ? 1.0 : pow(r, delay);
  double sn, cs;
  sincos(delay * th, &sn, &cs);
  const double2 v = make_double2(mag * cs, mag * sn);
  *zm = v;
  if (p.gamma_z != nullptr) {
    const float2 gz = p.gamma_z[(int64_t)line * p.k + bin];
    return cdiv(v, make_double2((double)gz.x, (double)gz.y));
  }
  return make_double2(v.x * invg, v.y * invg);
}

// Gauss-Jordan with partial pivoting on a register-resident system: lane `lane` holds row `lane` in m[] and its
// right-hand side in rhs. Returns the solution component this lane ends up owning; *col is its index
// (a permutation of 0..NP-1 over the lanes < NP; lanes >= NP return col = -1).
// The elimination step is a template over the column so that every register-array index is a compile-time
// constant (a plain `#pragma unroll` over the column left the array in local memory for NP = 20, 24, 28).
struct GJState {
  double2 rhs;
  double2 diag;
  int mycol;
  bool used;
};

template <int NP, int K>
struct GJStep {
  static __device__ __forceinline__ void run(double2 (&m)[NP], GJState& st, int lane, double2* line) {
    double key = st.used ? -1.0 : cnorm(m[K]);
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, key, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ok > key || (ok == key && oi < idx)) {
        key = ok;
        idx = oi;
      }
    }
    const int piv = idx;
    if (lane == piv) {
#pragma unroll
      for (int j = K; j < NP; ++j) line[j] = m[j];
      line[NP] = st.rhs;
      st.used = true;
      st.mycol = K;
      st.diag = m[K];
    }
    __syncwarp();
    if (lane != piv && lane < NP) {
      const double2 f = cmul(m[K], cinv(line[K]));
#pragma unroll
      for (int j = K + 1; j < NP; ++j) {
        const double2 pj = line[j];
        m[j].x -= f.x * pj.x - f.y * pj.y;
        m[j].y -= f.x * pj.y + f.y * pj.x;
      }
      const double2 pr = line[NP];
      st.rhs.x -= f.x * pr.x - f.y * pr.y;
      st.rhs.y -= f.x * pr.y + f.y * pr.x;
    }
    __syncwarp();
    if constexpr (K + 1 < NP) GJStep<NP, K + 1>::run(m, st, lane, line);
  }
};

template <int NP>
__device__ __forceinline__ double2 gauss_jordan(double2 (&m)[NP], double2 rhs, int lane, double2* line, int* col) {
  GJState st;
  st.rhs = rhs;
  st.diag = make_double2(1.0, 0.0);
  st.mycol = -1;
  st.used = lane >= NP;
  GJStep<NP, 0>::run(m, st, lane, line);
  *col = st.mycol;
  return cdiv(st.rhs, st.diag);
}

template <int NP>
__device__ __forceinline__ void load_block_constants(const SolveParams& p, double* s_a, double* s_at, double* s_invg,
                                                     double* s_b, double* s_c, double* s_delay) {
  const int n = p.n;
  for (int i = threadIdx.x; i < NP * NP; i += blockDim.x) {
    const int r = i / NP, c = i % NP;
    double v = 0.0;
    if (r < n && c < n) v = (double)(p.transpose_a ? p.a[c * n + r] : p.a[r * n + c]);
    s_a[r * NP + c] = v;
    s_at[c * NP + r] = v;
  }
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    const bool in = i < n;
    s_invg[i] = (in && p.gamma) ? 1.0 / (double)p.gamma[i] : 1.0;
    s_b[i] = (in && p.b) ? (double)p.b[i] : 0.0;
    s_c[i] = (in && p.c) ? (double)p.c[i] : 0.0;
    s_delay[i] = in ? (double)p.delays[i] : 0.0;
  }
}

// Row `lane` of M (adjoint = false) or of M^H (adjoint = true). Padded rows/columns (>= n) form an identity block.
template <int NP>
__device__ __forceinline__ void build_row(double2 (&m)[NP], const double* s_a, const double* s_at, int n, int lane,
                                          double2 dz, bool adjoint) {
  // M[i][j] = delta_ij dz_i - A[i][j]          -> needs A[lane][j]  = s_at[j*NP + lane]  (lane-contiguous)
  // M^H[i][j] = delta_ij conj(dz_i) - A[j][i]  -> needs A[j][lane]  = s_a [j*NP + lane]
  const double* src = adjoint ? s_a : s_at;
  const int li = lane < NP ? lane : 0;
  const double dre = lane < n ? dz.x : 1.0;
  const double dim = lane < n ? (adjoint ? -dz.y : dz.y) : 0.0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const bool on_diag = (j == lane);
    m[j] = make_double2((on_diag ? dre : 0.0) - src[j * NP + li], on_diag ? dim : 0.0);
  }
}

template <int NP>
__global__ void __launch_bounds__(kWarps * 32) solve_fwd_kernel(SolveParams p) {
  extern __shared__ double smem[];
  const int n = p.n;
  double* s_a = smem;
  double* s_at = s_a + NP * NP;
  double* s_invg = s_at + NP * NP;
  double* s_b = s_invg + NP;
  double* s_c = s_b + NP;
  double* s_delay = s_c + NP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + Smem<NP>::kBlock + (size_t)warp * Smem<NP>::kWarpFwd;
  double2* line = reinterpret_cast<double2*>(wbase);

  load_block_constants<NP>(p, s_a, s_at, s_invg, s_b, s_c, s_delay);
  __syncthreads();
  const int li = lane < NP ? lane : 0;
  const double my_invg = s_invg[li], my_delay = s_delay[li], my_b = s_b[li];

  for (int64_t bin = (int64_t)blockIdx.x * kWarps + warp; bin < p.k; bin += (int64_t)gridDim.x * kWarps) {
    double2 zm;
    double2 dz = make_double2(0.0, 0.0);
    if (lane < n) dz = diag_entry(p, bin, lane, my_invg, my_delay, &zm);
    double2 m[NP];
    build_row<NP>(m, s_a, s_at, n, lane, dz, false);
    int col;
    const double2 xr = gauss_jordan<NP>(m, make_double2(my_b, 0.0), lane, line, &col);
    const bool live = col >= 0 && col < n;
    if (p.x != nullptr && live) p.x[bin * n + col] = make_float2((float)xr.x, (float)xr.y);
    if (p.y != nullptr) {
      const double cr = live ? s_c[col] : 0.0;
      const int grp = live ? col / p.l : -1;
      for (int gi = 0; gi < p.g; ++gi) {
        const double re = warp_sum(grp == gi ? cr * xr.x : 0.0);
        const double im = warp_sum(grp == gi ? cr * xr.y : 0.0);
        if (lane == 0) p.y[bin * p.g + gi] = make_float2((float)re, (float)im);
      }
    }
  }
}

// Backward: adjoint solve per bin + accumulation of the parameter gradients. Each block writes one row of partial
// sums to ws; solve_bwd_reduce_kernel adds the rows in a fixed order.
template <int NP>
__global__ void __launch_bounds__(kWarps * 32) solve_bwd_kernel(SolveParams p) {
  extern __shared__ double smem[];
  const int n = p.n;
  double* s_a = smem;
  double* s_at = s_a + NP * NP;
  double* s_invg = s_at + NP * NP;
  double* s_b = s_invg + NP;
  double* s_c = s_b + NP;
  double* s_delay = s_c + NP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wbase = smem + Smem<NP>::kBlock + (size_t)warp * Smem<NP>::kWarpBwd;
  double2* line = reinterpret_cast<double2*>(wbase);
  double2* lam = line + (NP + 1);
  double2* xs = lam + NP;
  double* acc = reinterpret_cast<double*>(xs + NP);

  load_block_constants<NP>(p, s_a, s_at, s_invg, s_b, s_c, s_delay);
  for (int i = lane; i < NP * NP; i += 32) acc[i] = 0.0;
  __syncthreads();
  const int li = lane < NP ? lane : 0;
  const double my_invg = s_invg[li], my_delay = s_delay[li], my_c = s_c[li];

  double gb_acc = 0.0, gc_acc = 0.0, gig_acc = 0.0;
  for (int64_t bin = (int64_t)blockIdx.x * kWarps + warp; bin < p.k; bin += (int64_t)gridDim.x * kWarps) {
    double2 zm = make_double2(0.0, 0.0);
    double2 dz = make_double2(0.0, 0.0);
    double2 xr = make_double2(0.0, 0.0);
    double2 gyr = make_double2(0.0, 0.0);
    double2 rhs = make_double2(0.0, 0.0);
    if (lane < n) {
      dz = diag_entry(p, bin, lane, my_invg, my_delay, &zm);
      const float2 xv = p.xin[bin * n + lane];
      xr = make_double2((double)xv.x, (double)xv.y);
      if (p.gy != nullptr) {
        const float2 gv = p.gy[bin * p.g + lane / p.l];
        gyr = make_double2((double)gv.x, (double)gv.y);
        rhs.x = my_c * gyr.x;
        rhs.y = my_c * gyr.y;
      }
      if (p.gx != nullptr) {
        const float2 gv = p.gx[bin * n + lane];
        rhs.x += (double)gv.x;
        rhs.y += (double)gv.y;
      }
    }
    if (lane < NP) xs[lane] = xr;
    double2 m[NP];
    build_row<NP>(m, s_a, s_at, n, lane, dz, true);
    int col;
    const double2 sol = gauss_jordan<NP>(m, rhs, lane, line, &col);
    if (col >= 0) lam[col] = sol;
    __syncwarp();
    if (lane < n) {
      const double2 lr = lam[lane];
      // dL/dA_eff[i][j] = Re(lambda_i conj(x_j)); the reduce kernel transposes back when A_eff = A^T.
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const double2 o = xs[j];
        acc[lane + NP * j] += lr.x * o.x + lr.y * o.y;
      }
      gb_acc += lr.x;
      gc_acc += xr.x * gyr.x + xr.y * gyr.y;
      // grad wrt dz_i is -lambda_i conj(x_i); dz_i = zm_i * invgamma_i (real invgamma)
      const double2 t = cmulc(lr, xr);
      gig_acc -= zm.x * t.x + zm.y * t.y;
    }
    __syncwarp();
  }
  // block reduction: warp partials -> ws[blockIdx.x]
  __syncthreads();
  const size_t per = (size_t)n * n + 3 * (size_t)n;
  double* out = p.ws + (size_t)blockIdx.x * per;
  double* acc0 = smem + Smem<NP>::kBlock + Smem<NP>::kWarpFwd;  // warp 0 accumulator
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int row = i % n, colj = i / n;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += acc0[(size_t)w * Smem<NP>::kWarpBwd + row + NP * colj];
    out[i] = s;  // index row + n*col of dL/dA_eff
  }
  double* stash = reinterpret_cast<double*>(line);  // 2(NP+1) + 4 NP doubles available per warp
  if (lane < n) {
    stash[lane] = gb_acc;
    stash[n + lane] = gc_acc;
    stash[2 * n + lane] = gig_acc;
  }
  __syncthreads();
  double* stash0 = smem + Smem<NP>::kBlock;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += stash0[(size_t)w * Smem<NP>::kWarpBwd + i];
    out[(size_t)n * n + i] = s;
  }
}

// One warp per output element: lanes stride over the per-block partial rows, fixed-order shuffle reduction.
__global__ void solve_bwd_reduce_kernel(const double* ws, int nblocks, int n, int transpose_a, double* ga, double* gb,
                                        double* gc, double* gig) {
  const int per = n * n + 3 * n;
  const int lane = threadIdx.x & 31;
  const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= per) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += ws[(size_t)b * per + i];
  s = warp_sum(s);
  if (lane != 0) return;
  if (i < n * n) {
    int row = i % n, col = i / n;
    if (transpose_a) {
      const int t = row;
      row = col;
      col = t;
    }
    if (ga) ga[row * n + col] = s;
  } else {
    const int j = i - n * n;
    if (j < n) {
      if (gb) gb[j] = s;
    } else if (j < 2 * n) {
      if (gc) gc[j - n] = s;
    } else {
      if (gig) gig[j - 2 * n] = s;
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

template <int NP>
int launch_fwd(const SolveParams& p, cudaStream_t st) {
  const size_t smem = (Smem<NP>::kBlock + kWarps * Smem<NP>::kWarpFwd) * sizeof(double);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_fwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  solve_fwd_kernel<NP><<<grid_blocks(p.k), kWarps * 32, smem, st>>>(p);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

template <int NP>
int launch_bwd(const SolveParams& p, int blocks, cudaStream_t st) {
  const size_t smem = (Smem<NP>::kBlock + kWarps * Smem<NP>::kWarpBwd) * sizeof(double);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_bwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  solve_bwd_kernel<NP><<<blocks, kWarps * 32, smem, st>>>(p);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

#define DGFDN_DISPATCH_NP(n, CALL)  \
  do {                              \
    const int np__ = ((n) + 3) & ~3; \
    switch (np__) {                 \
      case 4: return CALL(4);       \
      case 8: return CALL(8);       \
      case 12: return CALL(12);     \
      case 16: return CALL(16);     \
      case 20: return CALL(20);     \
      case 24: return CALL(24);     \
      case 28: return CALL(28);     \
      default: return CALL(32);     \
    }                               \
  } while (0)

int dispatch_fwd(const SolveParams& p, cudaStream_t st) {
#define CALL_FWD(NP) launch_fwd<NP>(p, st)
  DGFDN_DISPATCH_NP(p.n, CALL_FWD);
#undef CALL_FWD
}

int dispatch_bwd(const SolveParams& p, int blocks, cudaStream_t st) {
#define CALL_BWD(NP) launch_bwd<NP>(p, blocks, st)
  DGFDN_DISPATCH_NP(p.n, CALL_BWD);
#undef CALL_BWD
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
  return dispatch_fwd(p, static_cast<cudaStream_t>(stream));
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dispatch_bwd(p, blocks, st)) return 1;
  const int per = n * n + 3 * n;
  solve_bwd_reduce_kernel<<<(per * 32 + 255) / 256, 256, 0, st>>>(p.ws, blocks, n, transpose_a, ga, gb, gc, ginvgamma);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
