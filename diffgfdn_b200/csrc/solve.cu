// K1 / K1^T: per-bin build + complex solve of the Grouped-FDN feedback loop.
//
//   M_k = diag(z_k^{m_i} / gamma_i) - A      x_k = M_k^{-1} b      y[k,g] = sum_{n in g} c_n x_k[n]
//
// Replaces FeedbackLoop.forward (reference diff_gfdn/feedback_loop.py:326-391: diag_embed + repeat +
// torch.linalg.inv on K dense NxN complex128 matrices) and the einsums of model.py:615-619 / :1083 /
// :237-250. The reference inverts, then contracts with c (per receiver!) and b; b is shared by every
// receiver, so one solve per bin is all that is needed and the (K,N,N) inverse never exists.
//
// Layout: a group of W lanes (W = 4, 8, 16 or 32, the power of two >= the padded system size NP) owns one system;
// a warp therefore works on 32/W systems at once (the G lossless 8x8 sub-FDN systems of a bin pack four to a
// warp). Lane i of the group holds row i in REGISTERS (NP complex doubles, every loop unrolled so each index is
// static). Gauss-Jordan elimination with partial pivoting: the pivot lane is found with ONE warp-reduce
// (REDUX.MAX on a key made of the high word of |m|^2 and the lane number; a log2(W)-step shuffle butterfly for
// sub-warp groups), it publishes its row through a double-buffered shared-memory line and every other lane
// eliminates against it. Rows are never swapped -- a lane remembers which column it was pivot for and ends up
// holding that component of the solution. z^m is formed by binary powering in float64 (22 complex products,
// relative error < 1e-12 for m < 4096).
// Arithmetic is float64 (the reference inverts in complex128, feedback_loop.py:391); outputs are complex64 like
// the reference's P. The adjoint kernel solves M^H lambda = g the same way and reduces the parameter gradients
// with a fixed-order two-stage reduction (deterministic).
//
// "Group mode" (nsys = G > 1) solves the G independent LxL systems diag(z^{m_g}) - M_g of
// DiffGFDN.sub_fdn_output (model.py:209-252) instead of one block-diagonal NxN system: 1/G^2 of the flops.
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kWarps = 4;

struct SolveParams {
  int n;      // rows of one system (N in coupled mode, L in group mode)
  int nsys;   // systems per bin (1 in coupled mode, G in group mode)
  int ntot;   // total number of delay lines N = n * nsys
  int g, l;   // groups and lines per group (for y)
  int64_t k;
  const double2* z;
  const int32_t* delays;
  const float* a;  // coupled: [N,N]; group mode: [G,L,L]
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
  // saved elimination (optional): multipliers of every step, and per lane its pivot value / z^m / pivot column
  float2* fac;   // [bins * nsys][NP][W]
  float4* rec;   // [bins * nsys][W]   (pivot.x, pivot.y, zm.x, zm.y)
  int* pcol;     // [bins * nsys][W]   column this lane was pivot for (-1: padding lane)
  // FIR coupling (filter_matrix, reference feedback_loop.py:362-373): A(z_k) = sum_p taps[p] z_k^-p, taps [P][N][N] real;
  // null for a constant A. The adjoint kernel then also writes lambda [K, N] (the tap gradients are a DFT-weighted sum of
  // lambda x^H over the bins, formed by the caller).
  const float* taps;
  int ntaps;
  float2* lam;
};

// Shared memory (in doubles). Block-wide constants per system type q: A_eff row-major [NP*NP] and transposed
// [NP*NP]; invgamma/b/c/delay [4*ntot_pad]. Per lane group: two pivot lines of (NP+1) double2, and (backward)
// lambda[NP], x[NP] double2 and the gradient accumulator NP*NP.
template <int NP>
struct Smem {
  static constexpr size_t kConstPerSys = 2 * (size_t)NP * NP + 4 * NP;
  static constexpr size_t kGroupFwd = 2 * 2 * (NP + 1);
  static constexpr size_t kGroupBwd = kGroupFwd + 4 * NP + (size_t)NP * NP;
  // mixed-precision forward: two float2 pivot lines (NP + 1 entries, pitch NP + 2: 16-byte aligned for 128-bit accesses) |
  // solution by column, double2 [NP] | pivot lane per step, int [NP]
  // (an even number of doubles: the double2 part stays 16-byte aligned from group to group)
  static constexpr size_t kGroupMixed = (2 * (NP + 2) + 2 * NP + (NP + 1) / 2 + 2) & ~(size_t)1;
};

__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = r * (2.0 - d * r);
  r = r * (2.0 - d * r);
  return r;
}
__device__ __forceinline__ double2 fast_cinv(double2 b) {
  const double s = fast_rcp(b.x * b.x + b.y * b.y);
  return make_double2(b.x * s, -b.y * s);
}

// z^m by binary powering; nbits = bit length of the largest exponent in the warp (uniform trip count).
__device__ __forceinline__ double2 cpow_int(double2 z, int m, int nbits) {
  double2 r = make_double2(1.0, 0.0);
  double2 p = z;
  for (int bit = 0; bit < nbits; ++bit) {
    if (m & (1 << bit)) r = cmul(r, p);
    p = cmul(p, p);
  }
  return r;
}

template <int W>
__device__ __forceinline__ int pivot_sublane(unsigned key) {
  if (W == 32) {
    key = __reduce_max_sync(0xffffffffu, key);
  } else {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
      const unsigned other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
  }
  return (W - 1) - (int)(key & (unsigned)(W - 1));
}

struct GJState {
  double2 rhs;
  double2 diag;
  int mycol;
  bool used;
};

// One elimination step per template instance, so that every register-array index is a compile-time constant.
template <int NP, int W, int K>
struct GJStep {
  static __device__ __forceinline__ void run(double2 (&m)[NP], GJState& st, int sl, double2* line0, double2* line1,
                                             float2* fac, double2* fm) {
    double2* line = (K & 1) ? line1 : line0;
    unsigned key = 0u;
    const double nrm = cnorm(m[K]);
    if (!st.used) {
      const unsigned hi = (unsigned)__double2hiint(nrm);  // monotone in |m|^2 for non-negative doubles
      key = 0x80000000u | (hi & ~(unsigned)(W - 1)) | (unsigned)(W - 1 - sl);
    }
    // Every lane takes the reciprocal of its own candidate while the pivot search is in flight: the winner publishes
    // 1 / pivot in the slot of the pivot value, so the reciprocal (rcp + 2 Newton steps) leaves the critical path
    // search -> broadcast -> multiplier. Same instruction count: the other lanes no longer invert what they read.
    const double rn = fast_rcp(nrm);
    const double2 myinv = make_double2(m[K].x * rn, -m[K].y * rn);
    const int piv = pivot_sublane<W>(key);
    if (sl == piv) {
      line[K] = myinv;
#pragma unroll
      for (int j = K + 1; j < NP; ++j) line[j] = m[j];
      line[NP] = st.rhs;
      st.used = true;
      st.mycol = K;
      st.diag = m[K];
    }
    __syncwarp();
    float2 fsave = make_float2(0.f, 0.f);
    if (fm != nullptr) fm[K] = make_double2(0.0, 0.0);  // (constant index after inlining: stays in registers)
    if (sl != piv && sl < NP) {
      const double2 f = cmul(m[K], line[K]);
      fsave = make_float2((float)f.x, (float)f.y);
      if (fm != nullptr) fm[K] = f;
      const double nfx = -f.x, nfy = -f.y;
#pragma unroll
      for (int j = K + 1; j < NP; ++j) {  // m[j] -= f * p[j]: four fused multiply-adds per complex element
        const double2 pj = line[j];
        m[j].x = fma(f.y, pj.y, fma(nfx, pj.x, m[j].x));
        m[j].y = fma(nfy, pj.x, fma(nfx, pj.y, m[j].y));
      }
      const double2 pr = line[NP];
      st.rhs.x = fma(f.y, pr.y, fma(nfx, pr.x, st.rhs.x));
      st.rhs.y = fma(nfy, pr.x, fma(nfx, pr.y, st.rhs.y));
    }
    // the multiplier of this (step, row): what the adjoint replay needs (0 for the pivot row and the padding lanes)
    if (fac != nullptr) fac[K * W + sl] = fsave;
    // no second barrier: step K+1 writes the other line; step K+2 reuses this one only after the barrier of K+1
    if constexpr (K + 1 < NP) GJStep<NP, W, K + 1>::run(m, st, sl, line0, line1, fac, fm);
  }
};

// Returns the solution component this lane ends up owning; *col is its index (a permutation of 0..NP-1 over the
// sub-lanes < NP; sub-lanes >= NP return col = -1).
template <int NP, int W>
__device__ __forceinline__ double2 gauss_jordan(double2 (&m)[NP], double2 rhs, int sl, double2* line0, double2* line1,
                                                int* col, float2* fac = nullptr, double2* pivot = nullptr,
                                                double2* fm = nullptr) {
  GJState st;
  st.rhs = rhs;
  st.diag = make_double2(1.0, 0.0);
  st.mycol = -1;
  st.used = sl >= NP;
  GJStep<NP, W, 0>::run(m, st, sl, line0, line1, fac, fm);
  __syncwarp();  // the lines may be rewritten by the caller's next system
  *col = st.mycol;
  if (pivot != nullptr) *pivot = st.diag;
  return cmul(st.rhs, fast_cinv(st.diag));
}

template <int NP>
__device__ __forceinline__ void load_block_constants(const SolveParams& p, double* s_const, double* s_vec) {
  const int n = p.n;
  for (int i = threadIdx.x; i < p.nsys * NP * NP; i += blockDim.x) {
    const int q = i / (NP * NP), rc = i % (NP * NP);
    const int r = rc / NP, c = rc % NP;
    double v = 0.0;
    if (r < n && c < n) {
      const float* aq = p.a + (size_t)q * n * n;
      v = (double)(p.transpose_a ? aq[c * n + r] : aq[r * n + c]);
    }
    double* base = s_const + (size_t)q * 2 * NP * NP;
    base[r * NP + c] = v;            // A_eff row-major
    base[NP * NP + c * NP + r] = v;  // A_eff transposed
  }
  // per delay line: invgamma, b, c, delay  (index q*NP + i)
  for (int i = threadIdx.x; i < p.nsys * NP; i += blockDim.x) {
    const int q = i / NP, r = i % NP;
    const bool in = r < n;
    const int line = q * n + r;
    s_vec[0 * p.nsys * NP + i] = (in && p.gamma) ? 1.0 / (double)p.gamma[line] : 1.0;
    s_vec[1 * p.nsys * NP + i] = (in && p.b) ? (double)p.b[line] : 0.0;
    s_vec[2 * p.nsys * NP + i] = (in && p.c) ? (double)p.c[line] : 0.0;
    s_vec[3 * p.nsys * NP + i] = in ? (double)p.delays[line] : 0.0;
  }
}

// Row `sl` of M (adjoint = false) or of M^H (adjoint = true). Padded rows/columns (>= n) form an identity block.
template <int NP>
__device__ __forceinline__ void build_row(double2 (&m)[NP], const double* s_a, const double* s_at, int n, int sl,
                                          double2 dz, bool adjoint) {
  // M[i][j] = delta_ij dz_i - A[i][j]          -> needs A[sl][j]  = s_at[j*NP + sl]  (lane-contiguous)
  // M^H[i][j] = delta_ij conj(dz_i) - A[j][i]  -> needs A[j][sl]  = s_a [j*NP + sl]
  const double* src = adjoint ? s_a : s_at;
  const int li = sl < NP ? sl : 0;
  const double dre = sl < n ? dz.x : 1.0;
  const double dim = sl < n ? (adjoint ? -dz.y : dz.y) : 0.0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const bool on_diag = (j == sl);
    m[j] = make_double2((on_diag ? dre : 0.0) - src[j * NP + li], on_diag ? dim : 0.0);
  }
}

// FIR coupling: the taps staged in shared memory as s_taps[p][j * NP + i] = A_eff,p[i][j] (forward) or A_eff,p[j][i] (adjoint),
// A_eff = A^T when transpose_a; padded rows / columns are zero.
template <int NP>
__device__ __forceinline__ void stage_taps(const SolveParams& p, float* s_taps, bool adjoint) {
  const int n = p.n;
  for (int i = threadIdx.x; i < p.ntaps * NP * NP; i += blockDim.x) {
    const int t = i / (NP * NP), rc = i % (NP * NP);
    const int j = rc / NP, l = rc % NP;  // lane l reads entry (row r, column c) of A_eff
    const int r = adjoint ? j : l, c = adjoint ? l : j;
    float v = 0.f;
    if (r < n && c < n) v = p.transpose_a ? p.taps[((size_t)t * n + c) * n + r] : p.taps[((size_t)t * n + r) * n + c];
    s_taps[i] = v;
  }
}

// Row `sl` of M_k = D_k - sum_p A_p z_k^-p (adjoint = false) or of M_k^H (adjoint = true); zinv = 1 / z_k.
template <int NP>
__device__ __forceinline__ void build_row_fir(double2 (&m)[NP], const float* s_taps, int ntaps, int n, int sl, double2 dz,
                                              double2 zinv, bool adjoint) {
  const int li = sl < NP ? sl : 0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const bool on_diag = (j == sl);
    m[j] = make_double2(on_diag ? (sl < n ? dz.x : 1.0) : 0.0, on_diag && sl < n ? (adjoint ? -dz.y : dz.y) : 0.0);
  }
  double2 w = make_double2(1.0, 0.0);  // z^-p (conjugated in the adjoint)
  if (adjoint) zinv.y = -zinv.y;
  for (int t = 0; t < ntaps; ++t) {
    const float* tp = s_taps + (size_t)t * NP * NP;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double a = (double)tp[j * NP + li];
      m[j].x = fma(-a, w.x, m[j].x);
      m[j].y = fma(-a, w.y, m[j].y);
    }
    w = cmul(w, zinv);
  }
}

// z^{m_i} / gamma_i for this lane's delay line; also returns z^m alone through zm.
__device__ __forceinline__ double2 diag_entry(const SolveParams& p, int64_t bin, int line, double invg, int delay,
                                              int nbits, double2* zm) {
  const double2 v = cpow_int(p.z[bin], delay, nbits);
  *zm = v;
  if (p.gamma_z != nullptr) {
    const float2 gz = p.gamma_z[(int64_t)line * p.k + bin];
    return cmul(v, fast_cinv(make_double2((double)gz.x, (double)gz.y)));
  }
  return make_double2(v.x * invg, v.y * invg);
}

// Lane-group bookkeeping shared by both kernels: which system type q this group serves and which bins.
template <int W>
struct GroupIndex {
  int sl, q;
  int64_t sgid, bin0, bin_stride;
  __device__ __forceinline__ GroupIndex(const SolveParams& p) {
    constexpr int kSpw = 32 / W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sl = lane % W;
    sgid = ((int64_t)blockIdx.x * kWarps + warp) * kSpw + lane / W;
    const int64_t total = (int64_t)gridDim.x * kWarps * kSpw;
    q = (int)(sgid % p.nsys);
    bin0 = sgid / p.nsys;
    bin_stride = (total - q + p.nsys - 1) / p.nsys;  // number of lane groups serving system type q
  }
};

template <int NP, int W>
__global__ void __launch_bounds__(kWarps * 32, (NP <= 24 ? 3 : 2)) solve_fwd_kernel(SolveParams p) {
  extern __shared__ double smem[];
  constexpr int kSpw = 32 / W;
  const int n = p.n;
  double* s_const = smem;
  double* s_vec = s_const + (size_t)p.nsys * 2 * NP * NP;
  double* s_groups = s_vec + (size_t)p.nsys * 4 * NP;
  const GroupIndex<W> gi(p);
  const int sl = gi.sl, q = gi.q;
  const int lg = (threadIdx.x >> 5) * kSpw + (threadIdx.x & 31) / W;  // lane group inside the block
  double2* line0 = reinterpret_cast<double2*>(s_groups + (size_t)lg * Smem<NP>::kGroupFwd);
  double2* line1 = line0 + (NP + 1);
  float* s_taps = reinterpret_cast<float*>(s_groups + (size_t)kWarps * kSpw * Smem<NP>::kGroupFwd);

  load_block_constants<NP>(p, s_const, s_vec);
  if (p.taps != nullptr) stage_taps<NP>(p, s_taps, false);
  __syncthreads();
  const double* s_a = s_const + (size_t)q * 2 * NP * NP;
  const double* s_at = s_a + NP * NP;
  const int li = q * NP + (sl < NP ? sl : 0);
  const double my_invg = s_vec[li], my_b = s_vec[p.nsys * NP + li], my_c = s_vec[2 * p.nsys * NP + li];
  const int my_delay = (int)s_vec[3 * p.nsys * NP + li];
  const int my_line = q * n + sl;
  const int nbits = 32 - __clz((int)__reduce_max_sync(0xffffffffu, (unsigned)(sl < n ? my_delay : 0)));

  // every lane of a warp runs the same number of iterations (shuffles are warp wide); idle groups clamp the bin
  const int64_t iters = (p.k + gi.bin_stride - 1) / gi.bin_stride;
  int64_t iters_max = iters;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t other = __shfl_xor_sync(0xffffffffu, iters_max, o);
    iters_max = other > iters_max ? other : iters_max;
  }
  for (int64_t it = 0; it < iters_max; ++it) {
    int64_t bin = gi.bin0 + it * gi.bin_stride;
    const bool live_bin = bin < p.k;
    if (!live_bin) bin = p.k - 1;
    double2 zm = make_double2(0.0, 0.0);
    double2 dz = make_double2(0.0, 0.0);
    if (sl < n) dz = diag_entry(p, bin, my_line, my_invg, my_delay, nbits, &zm);
    double2 m[NP];
    if (p.taps != nullptr)
      build_row_fir<NP>(m, s_taps, p.ntaps, n, sl, dz, fast_cinv(p.z[bin]), false);
    else
      build_row<NP>(m, s_a, s_at, n, sl, dz, false);
    int col;
    double2 pivot;
    const bool save = p.fac != nullptr && live_bin;
    const int64_t slot = bin * p.nsys + q;  // saved-elimination slot of this (bin, system)
    const double2 xr = gauss_jordan<NP, W>(m, make_double2(my_b, 0.0), sl, line0, line1, &col,
                                           save ? p.fac + slot * (NP * W) : nullptr, &pivot);
    if (save) {
      p.rec[slot * W + sl] = make_float4((float)pivot.x, (float)pivot.y, (float)zm.x, (float)zm.y);
      p.pcol[slot * W + sl] = col;
    }
    const bool live = col >= 0 && col < n;
    if (p.x != nullptr && live && live_bin) p.x[bin * p.ntot + q * n + col] = make_float2((float)xr.x, (float)xr.y);
    if (p.y != nullptr) {
      // c_col x_col through the (now idle) pivot line; one lane per output group adds its entries in order
      if (live) {
        const double cr = s_vec[2 * p.nsys * NP + q * NP + col];
        line0[col] = make_double2(cr * xr.x, cr * xr.y);
      }
      __syncwarp();
      if (p.nsys == 1) {
        if (sl < p.g && live_bin) {
          double re = 0.0, im = 0.0;
          for (int j = 0; j < p.l; ++j) {
            const double2 v = line0[sl * p.l + j];
            re += v.x;
            im += v.y;
          }
          p.y[bin * p.g + sl] = make_float2((float)re, (float)im);
        }
      } else if (sl == 0 && live_bin) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < n; ++j) {
          const double2 v = line0[j];
          re += v.x;
          im += v.y;
        }
        p.y[bin * p.g + q] = make_float2((float)re, (float)im);
      }
      __syncwarp();
    }
  }
  (void)my_c;
}

// ---- mixed-precision forward: float32 elimination + one float64 residual-refinement step ---------------------------
// The forward solve of the coupled system is bound by instruction issue / latency, not by the FP64 pipe's peak (35 % of it):
// 24 dependent elimination steps per bin with 96 registers of row state per lane. Here the ROW lives in float32 (half the
// registers, FFMA latency, 128-bit pivot-line reads carry two complex entries) while everything that fixes the accuracy
// stays float64: z^m by binary powering, the diagonal z^m / gamma, and the residual r = b - M x0 of the float32 solution
// x0, formed against the float64 A in shared memory. The correction M delta = r is solved by replaying the saved
// elimination (the multipliers are still in registers): delta = E^-1 T_N ... T_1 r, one shuffle + one complex FMA per
// step. x = x0 + delta is then accurate to ~(cond(M) eps_32)^2 -- the same multipliers, stored as float32, are what the
// adjoint replay (solve_bwd_replay_kernel) has always used. DGFDN_SOLVE_MIXED=0 keeps the all-float64 kernel.
__device__ __forceinline__ float rcp_nr(float d) {  // rcp.approx + one Newton step (~1 ulp; IEEE rounding is not needed here)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}

struct GJStateF {
  float2 rhs;
  float2 diag;
  int mycol;
  bool used;
};

// The float32 row is held as NP / 2 float4 (two complex entries each): the pivot lane publishes its row with 128-bit
// stores straight from those register quads. (Held as float2 pairs, the compiler still fused the stores into 128-bit ones
// and paid two MOVs per entry to pack them -- 12 % of the kernel's instructions, in a branch only one lane takes.)
__device__ __forceinline__ float2 row_get(const float4& q, int j) { return (j & 1) ? make_float2(q.z, q.w) : make_float2(q.x, q.y); }

template <int NP, int W, int K>
struct GJStepF {
  static __device__ __forceinline__ void run(float4 (&m4)[NP / 2], GJStateF& st, int sl, float2* line0, float2* line1, float2* fac,
                                             float2 (&fm)[NP]) {
    float2* line = (K & 1) ? line1 : line0;  // NP + 1 complex entries, 16-byte aligned
    unsigned key = 0u;
    const float2 mk = row_get(m4[K / 2], K);
    const float nrm = mk.x * mk.x + mk.y * mk.y;
    if (!st.used)  // monotone in |m|^2 for non-negative floats; the low bits carry the lane
      key = 0x80000000u | ((__float_as_uint(nrm) >> 1) & ~(unsigned)(W - 1)) | (unsigned)(W - 1 - sl);
    const float rn = rcp_nr(nrm);  // every lane inverts its own candidate while the search is in flight
    const float2 myinv = make_float2(mk.x * rn, -mk.y * rn);
    const int piv = pivot_sublane<W>(key);
    if (sl == piv) {
      line[K] = myinv;
      if constexpr ((K & 1) == 0) line[K + 1] = make_float2(m4[K / 2].z, m4[K / 2].w);  // the other half of K's quad
#pragma unroll
      for (int q = K / 2 + 1; q < NP / 2; ++q) reinterpret_cast<float4*>(line)[q] = m4[q];
      line[NP] = st.rhs;
      st.used = true;
      st.mycol = K;
      st.diag = mk;
    }
    __syncwarp();
    float2 f = make_float2(0.f, 0.f);
    if (sl != piv && sl < NP) {
      f = cmulf(mk, line[K]);
      const float nfx = -f.x, nfy = -f.y;
      if constexpr ((K & 1) == 0) {
        const float2 pj = line[K + 1];
        m4[K / 2].z = fmaf(f.y, pj.y, fmaf(nfx, pj.x, m4[K / 2].z));
        m4[K / 2].w = fmaf(nfy, pj.x, fmaf(nfx, pj.y, m4[K / 2].w));
      }
#pragma unroll
      for (int q = K / 2 + 1; q < NP / 2; ++q) {
        const float4 pq = reinterpret_cast<const float4*>(line)[q];
        m4[q].x = fmaf(f.y, pq.y, fmaf(nfx, pq.x, m4[q].x));
        m4[q].y = fmaf(nfy, pq.x, fmaf(nfx, pq.y, m4[q].y));
        m4[q].z = fmaf(f.y, pq.w, fmaf(nfx, pq.z, m4[q].z));
        m4[q].w = fmaf(nfy, pq.z, fmaf(nfx, pq.w, m4[q].w));
      }
      const float2 pr = line[NP];
      st.rhs.x = fmaf(f.y, pr.y, fmaf(nfx, pr.x, st.rhs.x));
      st.rhs.y = fmaf(nfy, pr.x, fmaf(nfx, pr.y, st.rhs.y));
    }
    fm[K] = f;
    if (fac != nullptr) fac[K * W + sl] = f;
    if constexpr (K + 1 < NP) GJStepF<NP, W, K + 1>::run(m4, st, sl, line0, line1, fac, fm);
  }
};

// delta <- T_K delta for the saved steps in order: component i != p_K loses f[K][i] * delta[p_K]
template <int NP, int W, int K>
struct CorrStepF {
  static __device__ __forceinline__ void run(const float2 (&fm)[NP], float2& r, int sl, int base, const int* ptab) {
    const int src = ptab[K];
    const float vx = __shfl_sync(0xffffffffu, r.x, base + src);
    const float vy = __shfl_sync(0xffffffffu, r.y, base + src);
    if (sl != src) {  // (fm[K] is zero for the pivot lane and the padding lanes anyway)
      r.x = fmaf(fm[K].y, vy, fmaf(-fm[K].x, vx, r.x));
      r.y = fmaf(-fm[K].y, vx, fmaf(-fm[K].x, vy, r.y));
    }
    if constexpr (K + 1 < NP) CorrStepF<NP, W, K + 1>::run(fm, r, sl, base, ptab);
  }
};

template <int NP, int W>
__global__ void __launch_bounds__(kWarps * 32, 4) solve_fwd_mixed_kernel(SolveParams p) {
  extern __shared__ double smem[];
  constexpr int kSpw = 32 / W;
  const int n = p.n;
  double* s_const = smem;
  double* s_vec = s_const + (size_t)p.nsys * 2 * NP * NP;
  double* s_groups = s_vec + (size_t)p.nsys * 4 * NP;
  const GroupIndex<W> gi(p);
  const int sl = gi.sl, q = gi.q;
  const int lane = threadIdx.x & 31;
  const int lg = (threadIdx.x >> 5) * kSpw + lane / W;
  // per lane group: two float2 pivot lines of NP + 1 | the solution by column as double2 [NP] | pivot lane of every step
  double* gbase = s_groups + (size_t)lg * Smem<NP>::kGroupMixed;
  float2* line0 = reinterpret_cast<float2*>(gbase);
  float2* line1 = line0 + (NP + 2);
  double2* xcol = reinterpret_cast<double2*>(gbase + 2 * (NP + 2));
  int* ptab = reinterpret_cast<int*>(gbase + 2 * (NP + 2) + 2 * NP);

  load_block_constants<NP>(p, s_const, s_vec);
  __syncthreads();
  const double* s_a = s_const + (size_t)q * 2 * NP * NP;
  const double* s_at = s_a + NP * NP;
  const int li = q * NP + (sl < NP ? sl : 0);
  const double my_invg = s_vec[li], my_b = s_vec[p.nsys * NP + li];
  const int my_delay = (int)s_vec[3 * p.nsys * NP + li];
  const int my_line = q * n + sl;
  const int nbits = 32 - __clz((int)__reduce_max_sync(0xffffffffu, (unsigned)(sl < n ? my_delay : 0)));
  const int base = lane - sl;

  const int64_t iters = (p.k + gi.bin_stride - 1) / gi.bin_stride;
  int64_t iters_max = iters;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t other = __shfl_xor_sync(0xffffffffu, iters_max, o);
    iters_max = other > iters_max ? other : iters_max;
  }
  for (int64_t it = 0; it < iters_max; ++it) {
    int64_t bin = gi.bin0 + it * gi.bin_stride;
    const bool live_bin = bin < p.k;
    if (!live_bin) bin = p.k - 1;
    double2 zm = make_double2(0.0, 0.0);
    double2 dz = make_double2(0.0, 0.0);
    if (sl < n) dz = diag_entry(p, bin, my_line, my_invg, my_delay, nbits, &zm);
    // row `sl` of M in float32 (padded rows / columns: identity)
    float4 m4[NP / 2];
    float2 fm[NP];
    {
      const int lr = sl < NP ? sl : 0;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const bool on_diag = (j == sl);
        const float re = (float)((on_diag ? (sl < n ? dz.x : 1.0) : 0.0) - s_at[j * NP + lr]);
        const float im = on_diag && sl < n ? (float)dz.y : 0.f;
        if (j & 1)
          m4[j / 2].z = re, m4[j / 2].w = im;
        else
          m4[j / 2].x = re, m4[j / 2].y = im;
      }
    }
    const bool save = p.fac != nullptr && live_bin;
    const int64_t slot = bin * p.nsys + q;
    GJStateF st;
    st.rhs = make_float2((float)my_b, 0.f);
    st.diag = make_float2(1.f, 0.f);
    st.mycol = -1;
    st.used = sl >= NP;
    GJStepF<NP, W, 0>::run(m4, st, sl, line0, line1, save ? p.fac + slot * (NP * W) : nullptr, fm);
    __syncwarp();
    const int col = st.mycol;
    const float dn = rcp_nr(st.diag.x * st.diag.x + st.diag.y * st.diag.y);
    const float2 dinv = make_float2(st.diag.x * dn, -st.diag.y * dn);
    const float2 x0 = cmulf(st.rhs, dinv);
    if (save) {
      p.rec[slot * W + sl] = make_float4(st.diag.x, st.diag.y, (float)zm.x, (float)zm.y);
      p.pcol[slot * W + sl] = col;
    }
    // ---- float64 residual of row sl:  r = b - (dz x0[sl] - sum_j A[sl][j] x0[j])
    if (col >= 0) {
      xcol[col] = make_double2((double)x0.x, (double)x0.y);
      ptab[col] = sl;
    }
    __syncwarp();
    float2 r = make_float2(0.f, 0.f);
    if (sl < NP) {
      const int lr = sl;
      double ax = 0.0, ay = 0.0;
#pragma unroll 8
      for (int j = 0; j < NP; ++j) {
        const double a = s_at[j * NP + lr];
        const double2 xj = xcol[j];
        ax = fma(a, xj.x, ax);
        ay = fma(a, xj.y, ay);
      }
      const double2 xs = xcol[sl];
      double rx, ry;
      if (sl < n) {
        rx = my_b - (dz.x * xs.x - dz.y * xs.y) + ax;
        ry = -(dz.x * xs.y + dz.y * xs.x) + ay;
      } else {  // identity padding row: x = b = 0 exactly
        rx = -xs.x;
        ry = -xs.y;
      }
      r = make_float2((float)rx, (float)ry);
    }
    // ---- correction by the saved elimination, then x = x0 + delta in float64
    CorrStepF<NP, W, 0>::run(fm, r, sl, base, ptab);
    const float2 delta = cmulf(r, dinv);
    const double2 xr = make_double2((double)x0.x + (double)delta.x, (double)x0.y + (double)delta.y);
    __syncwarp();  // xcol / ptab are rewritten below and by the next bin
    const bool live = col >= 0 && col < n;
    if (p.x != nullptr && live && live_bin) p.x[bin * p.ntot + q * n + col] = make_float2((float)xr.x, (float)xr.y);
    if (p.y != nullptr) {
      if (live) {
        const double cr = s_vec[2 * p.nsys * NP + q * NP + col];
        xcol[col] = make_double2(cr * xr.x, cr * xr.y);
      }
      __syncwarp();
      if (p.nsys == 1) {
        if (sl < p.g && live_bin) {
          double re = 0.0, im = 0.0;
          for (int j = 0; j < p.l; ++j) {
            const double2 v = xcol[sl * p.l + j];
            re += v.x;
            im += v.y;
          }
          p.y[bin * p.g + sl] = make_float2((float)re, (float)im);
        }
      } else if (sl == 0 && live_bin) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < n; ++j) {
          const double2 v = xcol[j];
          re += v.x;
          im += v.y;
        }
        p.y[bin * p.g + q] = make_float2((float)re, (float)im);
      }
      __syncwarp();
    }
  }
}

// Backward: adjoint solve per bin + accumulation of the parameter gradients. Each lane group writes one row of
// partial sums to ws; solve_bwd_reduce_kernel adds the rows of each system type in a fixed order.
template <int NP, int W>
__global__ void __launch_bounds__(kWarps * 32, (NP <= 24 ? 3 : 2)) solve_bwd_kernel(SolveParams p) {
  extern __shared__ double smem[];
  constexpr int kSpw = 32 / W;
  const int n = p.n;
  double* s_const = smem;
  double* s_vec = s_const + (size_t)p.nsys * 2 * NP * NP;
  double* s_groups = s_vec + (size_t)p.nsys * 4 * NP;
  const GroupIndex<W> gi(p);
  const int sl = gi.sl, q = gi.q;
  const int lg = (threadIdx.x >> 5) * kSpw + (threadIdx.x & 31) / W;
  double* gbase = s_groups + (size_t)lg * Smem<NP>::kGroupBwd;
  double2* line0 = reinterpret_cast<double2*>(gbase);
  double2* line1 = line0 + (NP + 1);
  double2* lam = line1 + (NP + 1);
  double2* xs = lam + NP;
  double* acc = reinterpret_cast<double*>(xs + NP);
  float* s_taps = reinterpret_cast<float*>(s_groups + (size_t)kWarps * kSpw * Smem<NP>::kGroupBwd);

  load_block_constants<NP>(p, s_const, s_vec);
  if (p.taps != nullptr) stage_taps<NP>(p, s_taps, true);
  if (sl < NP)
    for (int j = 0; j < NP; ++j) acc[sl + NP * j] = 0.0;
  __syncthreads();
  const double* s_a = s_const + (size_t)q * 2 * NP * NP;
  const double* s_at = s_a + NP * NP;
  const int li = q * NP + (sl < NP ? sl : 0);
  const double my_invg = s_vec[li], my_c = s_vec[2 * p.nsys * NP + li];
  const int my_delay = (int)s_vec[3 * p.nsys * NP + li];
  const int my_line = q * n + sl;
  const int my_group = p.nsys == 1 ? sl / p.l : q;
  const int nbits = 32 - __clz((int)__reduce_max_sync(0xffffffffu, (unsigned)(sl < n ? my_delay : 0)));

  const int64_t iters = (p.k + gi.bin_stride - 1) / gi.bin_stride;
  int64_t iters_max = iters;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t other = __shfl_xor_sync(0xffffffffu, iters_max, o);
    iters_max = other > iters_max ? other : iters_max;
  }
  double gb_acc = 0.0, gc_acc = 0.0, gig_acc = 0.0;
  for (int64_t it = 0; it < iters_max; ++it) {
    int64_t bin = gi.bin0 + it * gi.bin_stride;
    const bool live_bin = bin < p.k;
    if (!live_bin) bin = p.k - 1;
    double2 zm = make_double2(0.0, 0.0);
    double2 dz = make_double2(0.0, 0.0);
    double2 xr = make_double2(0.0, 0.0);
    double2 gyr = make_double2(0.0, 0.0);
    double2 rhs = make_double2(0.0, 0.0);
    if (sl < n) {
      dz = diag_entry(p, bin, my_line, my_invg, my_delay, nbits, &zm);
      if (live_bin) {
        const float2 xv = p.xin[bin * p.ntot + my_line];
        xr = make_double2((double)xv.x, (double)xv.y);
        if (p.gy != nullptr) {
          const float2 gv = p.gy[bin * p.g + my_group];
          gyr = make_double2((double)gv.x, (double)gv.y);
          rhs.x = my_c * gyr.x;
          rhs.y = my_c * gyr.y;
        }
        if (p.gx != nullptr) {
          const float2 gv = p.gx[bin * p.ntot + my_line];
          rhs.x += (double)gv.x;
          rhs.y += (double)gv.y;
        }
      }
    }
    if (sl < NP) xs[sl] = xr;
    double2 m[NP];
    if (p.taps != nullptr)
      build_row_fir<NP>(m, s_taps, p.ntaps, n, sl, dz, fast_cinv(p.z[bin]), true);
    else
      build_row<NP>(m, s_a, s_at, n, sl, dz, true);
    int col;
    const double2 sol = gauss_jordan<NP, W>(m, rhs, sl, line0, line1, &col);
    if (col >= 0) lam[col] = sol;
    __syncwarp();
    if (sl < n) {  // a dead bin contributes zeros: its rhs and x are zero
      const double2 lr = lam[sl];
      if (p.lam != nullptr && live_bin) p.lam[bin * p.ntot + my_line] = make_float2((float)lr.x, (float)lr.y);
      // dL/dA_eff[i][j] = Re(lambda_i conj(x_j)); the reduce kernel transposes back when A_eff = A^T.
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const double2 o = xs[j];
        acc[sl + NP * j] += lr.x * o.x + lr.y * o.y;
      }
      gb_acc += lr.x;
      gc_acc += xr.x * gyr.x + xr.y * gyr.y;
      // grad wrt dz_i is -lambda_i conj(x_i); dz_i = zm_i * invgamma_i (real invgamma)
      const double2 t = cmulc(lr, xr);
      gig_acc -= zm.x * t.x + zm.y * t.y;
    }
    __syncwarp();
  }
  // this lane group's partial sums -> ws[sgid]
  const size_t per = (size_t)n * n + 3 * (size_t)n;
  double* out = p.ws + (size_t)gi.sgid * per;
  if (sl < n) {
    for (int j = 0; j < n; ++j) out[sl + n * j] = acc[sl + NP * j];  // index row + n*col of dL/dA_eff
    out[(size_t)n * n + sl] = gb_acc;
    out[(size_t)n * n + n + sl] = gc_acc;
    out[(size_t)n * n + 2 * n + sl] = gig_acc;
  }
}

template <int W>
__device__ __forceinline__ double2 group_sum(double2 v) {  // sum over the W lanes of a group, every lane gets it
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// Backward WITHOUT a second factorisation: the forward kernel saved, per (bin, system), the multipliers f[k][i] of
// its Gauss-Jordan steps T_k = I - f_k e_{p_k}^T (T_N ... T_1 M = E, E[p_k, k] = pivot_k). The adjoint solve
// lambda = M^-H g = T_1^H ... T_N^H E^-H g is then N rank-one updates: w[p_k] = g_k / conj(pivot_k), and for
// k = N-1 .. 0 only component p_k changes, w[p_k] -= sum_i conj(f[k][i]) w[i] (one group-wide complex sum per step).
// ~1/4 of the instructions of a fresh elimination and a handful of registers, so the latency chain is hidden by
// occupancy. The gradient accumulation (outer products in shared memory, fixed-order two-stage reduction) is the
// one of solve_bwd_kernel.
template <int NP, int W>
__global__ void __launch_bounds__(kWarps * 32) solve_bwd_replay_kernel(SolveParams p) {
  extern __shared__ double smem[];
  constexpr int kSpw = 32 / W;
  const int n = p.n;
  double* s_c = smem;                                  // [nsys * NP] output gains per line
  double* s_groups = s_c + (size_t)p.nsys * NP;
  const GroupIndex<W> gi(p);
  const int sl = gi.sl, q = gi.q;
  const int lg = (threadIdx.x >> 5) * kSpw + (threadIdx.x & 31) / W;
  double* gbase = s_groups + (size_t)lg * (2 * NP + (size_t)NP * NP);
  double2* xs = reinterpret_cast<double2*>(gbase);
  double* acc = gbase + 2 * NP;
  for (int i = threadIdx.x; i < p.nsys * NP; i += blockDim.x) {
    const int qq = i / NP, r = i % NP;
    s_c[i] = (r < n && p.c) ? (double)p.c[qq * n + r] : 0.0;
  }
  if (sl < NP)
    for (int j = 0; j < NP; ++j) acc[sl + NP * j] = 0.0;
  __syncthreads();
  const int my_line = q * n + sl;
  const int my_group = p.nsys == 1 ? sl / p.l : q;

  const int64_t iters = (p.k + gi.bin_stride - 1) / gi.bin_stride;
  int64_t iters_max = iters;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t other = __shfl_xor_sync(0xffffffffu, iters_max, o);
    iters_max = other > iters_max ? other : iters_max;
  }
  double gb_acc = 0.0, gc_acc = 0.0, gig_acc = 0.0;
  for (int64_t it = 0; it < iters_max; ++it) {
    int64_t bin = gi.bin0 + it * gi.bin_stride;
    const bool live_bin = bin < p.k;
    if (!live_bin) bin = p.k - 1;
    const int64_t slot = bin * p.nsys + q;
    const float4 rec = p.rec[slot * W + sl];
    const int col = p.pcol[slot * W + sl];
    double2 xr = make_double2(0.0, 0.0), gyr = make_double2(0.0, 0.0), w = make_double2(0.0, 0.0);
    if (live_bin) {
      if (sl < n) {
        const float2 xv = p.xin[bin * p.ntot + my_line];
        xr = make_double2((double)xv.x, (double)xv.y);
        if (p.gy != nullptr) {
          const float2 gv = p.gy[bin * p.g + my_group];
          gyr = make_double2((double)gv.x, (double)gv.y);
        }
      }
      if (col >= 0 && col < n) {  // right-hand side of the column this lane was pivot for
        double2 g = make_double2(0.0, 0.0);
        if (p.gy != nullptr) {
          const float2 gv = p.gy[bin * p.g + (p.nsys == 1 ? col / p.l : q)];
          const double cc = s_c[q * NP + col];
          g = make_double2(cc * (double)gv.x, cc * (double)gv.y);
        }
        if (p.gx != nullptr) {
          const float2 gv = p.gx[bin * p.ntot + q * n + col];
          g.x += (double)gv.x;
          g.y += (double)gv.y;
        }
        const double px = (double)rec.x, py = (double)rec.y;
        const double inv = fast_rcp(px * px + py * py);
        w = make_double2((g.x * px - g.y * py) * inv, (g.x * py + g.y * px) * inv);  // g / conj(pivot)
      }
    }
    if (sl < NP) xs[sl] = xr;
    const float2* fac = p.fac + slot * (NP * W) + sl;
    {
      // The recursion itself runs in float32: the multipliers ARE float32, the group-wide sum is half the shuffles, and
      // lambda only needs the ~1e-6 this leaves (the gradient gate is 1e-3; the products below accumulate in float64).
      float2 wf = make_float2((float)w.x, (float)w.y);
#pragma unroll 4
      for (int k = NP - 1; k >= 0; --k) {
        const float2 f = fac[k * W];
        float sx = fmaf(f.x, wf.x, f.y * wf.y), sy = fmaf(f.x, wf.y, -f.y * wf.x);  // conj(f) w
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) {
          sx += __shfl_xor_sync(0xffffffffu, sx, o);
          sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (col == k) {
          wf.x -= sx;
          wf.y -= sy;
        }
      }
      w = make_double2((double)wf.x, (double)wf.y);
    }
    __syncwarp();
    if (sl < n) {  // a dead bin contributes zeros: its w and x are zero
      const double2 lr = w;
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const double2 o = xs[j];
        acc[sl + NP * j] += lr.x * o.x + lr.y * o.y;
      }
      gb_acc += lr.x;
      gc_acc += xr.x * gyr.x + xr.y * gyr.y;
      const double2 t = cmulc(lr, xr);
      gig_acc -= (double)rec.z * t.x + (double)rec.w * t.y;
    }
    __syncwarp();
  }
  const size_t per = (size_t)n * n + 3 * (size_t)n;
  double* out = p.ws + (size_t)gi.sgid * per;
  if (sl < n) {
    for (int j = 0; j < n; ++j) out[sl + n * j] = acc[sl + NP * j];
    out[(size_t)n * n + sl] = gb_acc;
    out[(size_t)n * n + n + sl] = gc_acc;
    out[(size_t)n * n + 2 * n + sl] = gig_acc;
  }
}

// Adjoint replay with the multipliers in registers (compile-time step index).
template <int NP, int W, int K>
struct ReplayStep {
  static __device__ __forceinline__ void run(const double2 (&fm)[NP], double2& w, int col) {
    const double2 f = fm[K];
    const double2 s = group_sum<W>(make_double2(f.x * w.x + f.y * w.y, f.x * w.y - f.y * w.x));  // conj(f) w
    if (col == K) {
      w.x -= s.x;
      w.y -= s.y;
    }
    if constexpr (K > 0) ReplayStep<NP, W, K - 1>::run(fm, w, col);
  }
};

// K1c: the colorless branch in ONE pass per bin. For every lossless LxL sub-FDN system: eliminate, y = c^T x, the
// spectral-flatness term (|y| - 1)^p (p = 4 where asym and |y| - 1 > 1, else 2; colorless_fdn/losses.py:20-73) and its
// gradient dL/dy, then the adjoint solve with the elimination still in registers (ReplayStep) and the gradient outer
// products. Replaces solve_fwd(groups) + colorless_fwd + colorless_bwd + solve_bwd(groups): the second elimination,
// the H_sub round trip and three launches disappear. loss_part[sgid] holds the group's partial sum of the loss terms.
template <int NP, int W>
__global__ void __launch_bounds__(kWarps * 32, (NP <= 8 ? 4 : 3)) solve_colorless_kernel(SolveParams p, int asym, double* loss_part) {
  extern __shared__ double smem[];
  constexpr int kSpw = 32 / W;
  const int n = p.n;
  double* s_const = smem;
  double* s_vec = s_const + (size_t)p.nsys * 2 * NP * NP;
  double* s_groups = s_vec + (size_t)p.nsys * 4 * NP;
  const GroupIndex<W> gi(p);
  const int sl = gi.sl, q = gi.q;
  const int lg = (threadIdx.x >> 5) * kSpw + (threadIdx.x & 31) / W;
  double* gbase = s_groups + (size_t)lg * Smem<NP>::kGroupBwd;
  double2* line0 = reinterpret_cast<double2*>(gbase);
  double2* line1 = line0 + (NP + 1);
  double2* xs = line1 + (NP + 1) + NP;  // (the lam slot of the backward layout is unused here)
  double* acc = reinterpret_cast<double*>(xs + NP);

  load_block_constants<NP>(p, s_const, s_vec);
  if (sl < NP)
    for (int j = 0; j < NP; ++j) acc[sl + NP * j] = 0.0;
  __syncthreads();
  const double* s_a = s_const + (size_t)q * 2 * NP * NP;
  const double* s_at = s_a + NP * NP;
  const double* s_c = s_vec + 2 * p.nsys * NP + q * NP;
  const int li = q * NP + (sl < NP ? sl : 0);
  const double my_invg = s_vec[li], my_b = s_vec[p.nsys * NP + li];
  const int my_delay = (int)s_vec[3 * p.nsys * NP + li];
  const int my_line = q * n + sl;
  const int nbits = 32 - __clz((int)__reduce_max_sync(0xffffffffu, (unsigned)(sl < n ? my_delay : 0)));
  const double inv_k = 1.0 / (double)p.k;

  const int64_t iters = (p.k + gi.bin_stride - 1) / gi.bin_stride;
  int64_t iters_max = iters;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t other = __shfl_xor_sync(0xffffffffu, iters_max, o);
    iters_max = other > iters_max ? other : iters_max;
  }
  double gb_acc = 0.0, gc_acc = 0.0, loss_acc = 0.0;
  for (int64_t it = 0; it < iters_max; ++it) {
    int64_t bin = gi.bin0 + it * gi.bin_stride;
    const bool live_bin = bin < p.k;
    if (!live_bin) bin = p.k - 1;
    double2 zm = make_double2(0.0, 0.0), dz = make_double2(0.0, 0.0);
    if (sl < n) dz = diag_entry(p, bin, my_line, my_invg, my_delay, nbits, &zm);
    double2 m[NP], fm[NP];
    build_row<NP>(m, s_a, s_at, n, sl, dz, false);
    int col;
    double2 pivot;
    const double2 xr = gauss_jordan<NP, W>(m, make_double2(my_b, 0.0), sl, line0, line1, &col, nullptr, &pivot, fm);
    const bool live = col >= 0 && col < n;
    if (col >= 0) xs[col] = live ? xr : make_double2(0.0, 0.0);
    __syncwarp();
    double2 y = make_double2(0.0, 0.0);
    for (int j = 0; j < n; ++j) {  // every lane of the group forms the same y (fixed order)
      const double2 v = xs[j];
      y.x += s_c[j] * v.x;
      y.y += s_c[j] * v.y;
    }
    const float2 yf = make_float2((float)y.x, (float)y.y);  // the loss sees the complex64 response, like the module path
    const double a = hypot((double)yf.x, (double)yf.y);
    const double d = a - 1.0;
    const double d2 = d * d;
    const bool quartic = asym && d > 1.0;
    const double dfd = quartic ? 4.0 * d2 * d : 2.0 * d;
    const double sc = (a > 0.0 && live_bin) ? dfd * inv_k / a : 0.0;
    const double2 gy = make_double2(sc * (double)yf.x, sc * (double)yf.y);
    if (sl == 0 && live_bin) loss_acc += quartic ? d2 * d2 : d2;
    // adjoint: w[p_k] = g_k / conj(pivot_k), g_k = c_k gy; then the saved steps backwards
    double2 w = make_double2(0.0, 0.0);
    if (live) {
      const double cc = s_c[col];
      const double inv = fast_rcp(pivot.x * pivot.x + pivot.y * pivot.y);
      const double gx = cc * gy.x, gyy = cc * gy.y;
      w = make_double2((gx * pivot.x - gyy * pivot.y) * inv, (gx * pivot.y + gyy * pivot.x) * inv);
    }
    ReplayStep<NP, W, NP - 1>::run(fm, w, col);
    if (sl < n) {
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const double2 o = xs[j];
        acc[sl + NP * j] += w.x * o.x + w.y * o.y;
      }
      gb_acc += w.x;
      const double2 xo = xs[sl];
      gc_acc += xo.x * gy.x + xo.y * gy.y;
    }
    __syncwarp();
  }
  const size_t per = (size_t)n * n + 3 * (size_t)n;
  double* out = p.ws + (size_t)gi.sgid * per;
  if (sl < n) {
    for (int j = 0; j < n; ++j) out[sl + n * j] = acc[sl + NP * j];
    out[(size_t)n * n + sl] = gb_acc;
    out[(size_t)n * n + n + sl] = gc_acc;
    out[(size_t)n * n + 2 * n + sl] = 0.0;
  }
  if (sl == 0) loss_part[gi.sgid] = loss_acc;
}

// loss[q] = (1/K) sum over the lane groups that served system type q (fixed order); one block of 256 threads per q
__global__ void colorless_loss_reduce_kernel(const double* part, int64_t ngroups, int nsys, int64_t k, double* loss) {
  __shared__ double red[8];
  const int q = blockIdx.x;
  const int t = threadIdx.x;
  double s = 0.0;
  for (int64_t sg = q + (int64_t)t * nsys; sg < ngroups; sg += (int64_t)blockDim.x * nsys) s += part[sg];
  s = warp_sum(s);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  if (t == 0) {
    double v = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[i];
    loss[q] = v / (double)k;
  }
}

// First stage of the gradient reduction for large grids: block (q, c) adds the rows of chunk c among the lane groups
// that served system type q (rows q, q + nsys, ...), thread i the elements i, i + blockDim, ... of a row (coalesced).
// out row c * nsys + q, so that solve_bwd_reduce_kernel finishes on `chunks * nsys` rows with the same row -> q rule.
constexpr int kReduceChunks = 64;
__global__ void solve_bwd_prereduce_kernel(const double* __restrict__ ws, int64_t ngroups, int per, int nsys,
                                           double* __restrict__ out) {
  const int q = blockIdx.x, c = blockIdx.y;
  const int64_t rows_q = (ngroups - q + nsys - 1) / nsys;  // rows of this system type
  const int64_t chunk = (rows_q + kReduceChunks - 1) / kReduceChunks;
  const int64_t r0 = (int64_t)c * chunk, r1 = r0 + chunk < rows_q ? r0 + chunk : rows_q;
  const int i = blockIdx.z * blockDim.x + threadIdx.x;
  if (i >= per) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four loads in flight; the order of the sum stays fixed
  int64_t r = r0;
  for (; r + 4 <= r1; r += 4) {
    s0 += ws[(size_t)(q + r * nsys) * per + i];
    s1 += ws[(size_t)(q + (r + 1) * nsys) * per + i];
    s2 += ws[(size_t)(q + (r + 2) * nsys) * per + i];
    s3 += ws[(size_t)(q + (r + 3) * nsys) * per + i];
  }
  for (; r < r1; ++r) s0 += ws[(size_t)(q + r * nsys) * per + i];
  out[(size_t)(c * nsys + q) * per + i] = (s0 + s1) + (s2 + s3);
}

// One warp per output element: lanes stride over the partial rows of the lane groups that served system type q
// (sgid = q, q + nsys, ...), fixed-order shuffle reduction.
__global__ void solve_bwd_reduce_kernel(const double* ws, int64_t ngroups, int n, int nsys, int transpose_a, double* ga,
                                        double* gb, double* gc, double* gig) {
  const int per = n * n + 3 * n;
  const int lane = threadIdx.x & 31;
  const int idx = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (idx >= per * nsys) return;
  const int q = idx / per, i = idx % per;
  double s = 0.0;
  for (int64_t sg = q + (int64_t)lane * nsys; sg < ngroups; sg += 32 * (int64_t)nsys) s += ws[(size_t)sg * per + i];
  s = warp_sum(s);
  if (lane != 0) return;
  if (i < n * n) {
    int row = i % n, col = i / n;
    if (transpose_a) {
      const int t = row;
      row = col;
      col = t;
    }
    if (ga) ga[(size_t)q * n * n + row * n + col] = s;
  } else {
    const int j = i - n * n;
    if (j < n) {
      if (gb) gb[q * n + j] = s;
    } else if (j < 2 * n) {
      if (gc) gc[q * n + j - n] = s;
    } else {
      if (gig) gig[q * n + j - 2 * n] = s;
    }
  }
}

// Reduction of `groups` partial rows of (n^2 + 3 n) doubles per system type: directly for small grids, through the
// chunked first stage (scratch `pre`, kReduceChunks * nsys rows) for large ones.
void launch_bwd_reduce(const double* ws, double* pre, int64_t groups, int n, int nsys, int transpose_a, double* ga,
                       double* gb, double* gc, double* gig, cudaStream_t st) {
  const int per = n * n + 3 * n;
  const double* rows = ws;
  if (groups > 4 * (int64_t)kReduceChunks * nsys) {
    solve_bwd_prereduce_kernel<<<dim3(nsys, kReduceChunks, (per + 127) / 128), 128, 0, st>>>(ws, groups, per, nsys, pre);
    rows = pre;
    groups = (int64_t)kReduceChunks * nsys;
  }
  solve_bwd_reduce_kernel<<<(per * nsys * 32 + 255) / 256, 256, 0, st>>>(rows, groups, n, nsys, transpose_a, ga, gb, gc, gig);
}

constexpr int lanes_for(int np) { return np <= 4 ? 4 : (np <= 8 ? 8 : (np <= 16 ? 16 : 32)); }

constexpr int kMaxBlocksPerSm = 12;  // the backward workspace is sized for this many blocks per SM in a grid

// One resident wave: `per_sm` is the occupancy of the instantiation being launched (a grid of 4 blocks per SM with
// only 3 resident leaves a 148-block second wave that runs at a third of the SM's throughput).
int grid_blocks(int64_t systems, int w, int per_sm) {
  const int per_block = kWarps * (32 / w);
  int64_t want = (systems + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : (per_sm > kMaxBlocksPerSm ? kMaxBlocksPerSm : per_sm));
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

// Resident blocks per SM of a kernel at its dynamic shared-memory size (queried once per instantiation and size).
template <class K>
int blocks_per_sm(K kern, size_t smem) {
  static size_t cached_smem = (size_t)-1;
  static int cached = 0;
  if (cached_smem != smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kWarps * 32, smem) != cudaSuccess) n = 1;
    cached = n < 1 ? 1 : n;
    cached_smem = smem;
  }
  return cached;
}

int check_common(int n, int nsys, int g, int64_t k) {
  DGFDN_CHECK(n >= 1 && n <= DGFDN_MAX_LINES, "solve: n=%d out of range [1,%d]", n, DGFDN_MAX_LINES);
  DGFDN_CHECK(nsys >= 1 && n * nsys <= DGFDN_MAX_LINES, "solve: %d systems of %d lines exceed %d lines", nsys, n,
              DGFDN_MAX_LINES);
  DGFDN_CHECK(g >= 1 && g <= DGFDN_MAX_GROUPS && (n * nsys) % g == 0, "solve: g=%d must divide N=%d and be <= %d", g,
              n * nsys, DGFDN_MAX_GROUPS);
  DGFDN_CHECK(nsys == 1 || nsys == g, "solve: group mode needs one system per group");
  DGFDN_CHECK(k >= 1, "solve: k=%lld must be positive", (long long)k);
  return 0;
}

template <int NP>
size_t smem_bytes(const SolveParams& p, bool bwd) {
  constexpr int W = lanes_for(NP);
  const size_t groups = kWarps * (32 / W);
  const size_t taps = p.taps != nullptr ? (((size_t)p.ntaps * NP * NP * sizeof(float) + 7) & ~(size_t)7) : 0;
  return ((size_t)p.nsys * Smem<NP>::kConstPerSys + groups * (bwd ? Smem<NP>::kGroupBwd : Smem<NP>::kGroupFwd)) *
             sizeof(double) + taps;
}

template <int NP>
int launch_fwd(const SolveParams& p, cudaStream_t st) {
  constexpr int W = lanes_for(NP);
  const size_t smem = smem_bytes<NP>(p, false);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_fwd_kernel<NP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = grid_blocks(p.k * p.nsys, W, blocks_per_sm(solve_fwd_kernel<NP, W>, smem));
  solve_fwd_kernel<NP, W><<<blocks, kWarps * 32, smem, st>>>(p);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

template <int NP>
int launch_fwd_mixed(const SolveParams& p, cudaStream_t st) {
  constexpr int W = lanes_for(NP);
  const size_t smem = ((size_t)p.nsys * Smem<NP>::kConstPerSys + (size_t)kWarps * (32 / W) * Smem<NP>::kGroupMixed) * sizeof(double);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_fwd_mixed_kernel<NP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = grid_blocks(p.k * p.nsys, W, blocks_per_sm(solve_fwd_mixed_kernel<NP, W>, smem));
  solve_fwd_mixed_kernel<NP, W><<<blocks, kWarps * 32, smem, st>>>(p);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

template <int NP>
int launch_bwd(const SolveParams& p, int* blocks_out, cudaStream_t st) {
  constexpr int W = lanes_for(NP);
  const size_t smem = smem_bytes<NP>(p, true);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_bwd_kernel<NP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = grid_blocks(p.k * p.nsys, W, blocks_per_sm(solve_bwd_kernel<NP, W>, smem));
  *blocks_out = blocks;
  solve_bwd_kernel<NP, W><<<blocks, kWarps * 32, smem, st>>>(p);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

template <int NP>
int launch_bwd_replay(const SolveParams& p, int* blocks_out, cudaStream_t st) {
  constexpr int W = lanes_for(NP);
  const size_t groups = kWarps * (32 / W);
  const size_t smem = ((size_t)p.nsys * NP + groups * (2 * NP + (size_t)NP * NP)) * sizeof(double);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_bwd_replay_kernel<NP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = grid_blocks(p.k * p.nsys, W, blocks_per_sm(solve_bwd_replay_kernel<NP, W>, smem));
  *blocks_out = blocks;
  solve_bwd_replay_kernel<NP, W><<<blocks, kWarps * 32, smem, st>>>(p);
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

int dispatch_fwd_mixed(const SolveParams& p, cudaStream_t st) {
#define CALL_FWD(NP) launch_fwd_mixed<NP>(p, st)
  DGFDN_DISPATCH_NP(p.n, CALL_FWD);
#undef CALL_FWD
}

int dispatch_bwd(const SolveParams& p, int* blocks, cudaStream_t st) {
#define CALL_BWD(NP) launch_bwd<NP>(p, blocks, st)
  DGFDN_DISPATCH_NP(p.n, CALL_BWD);
#undef CALL_BWD
}

int dispatch_bwd_replay(const SolveParams& p, int* blocks, cudaStream_t st) {
#define CALL_BWD(NP) launch_bwd_replay<NP>(p, blocks, st)
  DGFDN_DISPATCH_NP(p.n, CALL_BWD);
#undef CALL_BWD
}

// The saved elimination of one forward call: [fac | rec | pcol], W lanes per system
struct FactorLayout {
  size_t off_rec, off_pcol, total;
};
FactorLayout factor_layout(int n, int nsys, int64_t k) {
  const int np = (n + 3) & ~3;
  const int w = lanes_for(np > 32 ? 32 : np);
  const size_t slots = (size_t)k * nsys;
  FactorLayout f;
  size_t o = slots * np * w * sizeof(float2);
  o = (o + 255) & ~(size_t)255;
  f.off_rec = o;
  o += slots * w * sizeof(float4);
  o = (o + 255) & ~(size_t)255;
  f.off_pcol = o;
  o += slots * w * sizeof(int);
  f.total = o;
  return f;
}

void bind_factors(SolveParams& p, void* factors, int n, int nsys, int64_t k) {
  if (factors == nullptr) return;
  const FactorLayout f = factor_layout(n, nsys, k);
  unsigned char* base = static_cast<unsigned char*>(factors);
  p.fac = reinterpret_cast<float2*>(base);
  p.rec = reinterpret_cast<float4*>(base + f.off_rec);
  p.pcol = reinterpret_cast<int*>(base + f.off_pcol);
}

int lanes_runtime(int n) { return lanes_for((n + 3) & ~3); }

int fill_params(SolveParams& p, int n, int nsys, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                int transpose_a, const float* gamma, const void* gamma_z, const float* b, const float* c) {
  p.n = n;
  p.nsys = nsys;
  p.ntot = n * nsys;
  p.g = g;
  p.l = p.ntot / g;
  p.k = k;
  p.z = static_cast<const double2*>(z);
  p.delays = delays;
  p.a = a;
  p.transpose_a = transpose_a;
  p.gamma = gamma;
  p.gamma_z = static_cast<const float2*>(gamma_z);
  p.b = b;
  p.c = c;
  return 0;
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

static int64_t bwd_rows_bytes(int n);

static int solve_fwd_impl(int n, int nsys, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                          int transpose_a, const float* gamma, const void* gamma_z, const float* b, const float* c,
                          void* x, void* y, void* factors, void* stream, const float* taps = nullptr, int ntaps = 0) {
  if (check_common(n, nsys, g, k)) return 1;
  DGFDN_CHECK(z && delays && a && b, "solve_fwd: null input pointer");
  DGFDN_CHECK(y == nullptr || c != nullptr, "solve_fwd: y requested without c");
  SolveParams p{};
  fill_params(p, n, nsys, g, k, z, delays, a, transpose_a, gamma, gamma_z, b, c);
  p.x = static_cast<float2*>(x);
  p.y = static_cast<float2*>(y);
  if (taps != nullptr) {  // FIR coupling: per-bin complex A(z_k) in the float64 kernel, no saved elimination
    DGFDN_CHECK(nsys == 1 && ntaps >= 1 && ntaps <= 64 && factors == nullptr, "solve_fir_fwd: 1..64 taps, coupled mode, no factors");
    p.taps = taps;
    p.ntaps = ntaps;
    return dispatch_fwd(p, static_cast<cudaStream_t>(stream));
  }
  bind_factors(p, factors, n, nsys, k);
  // coupled systems with scalar absorption: float32 elimination + float64 refinement (solve_fwd_mixed_kernel).
  // Filter absorption (gamma_z), the small sub-FDN systems (group mode) and DGFDN_SOLVE_MIXED=0 take the float64 kernel.
  static const bool mixed_on = [] {
    const char* e = getenv("DGFDN_SOLVE_MIXED");
    return e == nullptr || atoi(e) != 0;
  }();
  bool mixed = mixed_on && nsys == 1 && n >= 12 && gamma_z == nullptr;
  if (const char* e = getenv("DGFDN_SOLVE_MIXED_FORCE")) mixed = atoi(e) != 0 && nsys == 1;
  if (mixed) return dispatch_fwd_mixed(p, static_cast<cudaStream_t>(stream));
  return dispatch_fwd(p, static_cast<cudaStream_t>(stream));
}

static int solve_bwd_impl(int n, int nsys, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                          int transpose_a, const float* gamma, const void* gamma_z, const float* c, const void* x,
                          const void* gy, const void* gx, double* ga, double* gb, double* gc, double* ginvgamma,
                          void* ws, const void* factors, void* stream, const float* taps = nullptr, int ntaps = 0,
                          void* lam = nullptr) {
  if (check_common(n, nsys, g, k)) return 1;
  DGFDN_CHECK(z && delays && a && x && ws, "solve_bwd: null input pointer");
  DGFDN_CHECK(gy || gx, "solve_bwd: need gy or gx");
  DGFDN_CHECK(gy == nullptr || c != nullptr, "solve_bwd: gy given without c");
  DGFDN_CHECK(taps == nullptr || (nsys == 1 && ntaps >= 1 && ntaps <= 64 && factors == nullptr && lam != nullptr),
              "solve_fir_bwd: 1..64 taps, coupled mode, no factors, lambda output required");
  SolveParams p{};
  fill_params(p, n, nsys, g, k, z, delays, a, transpose_a, gamma, gamma_z, nullptr, c);
  p.taps = taps;
  p.ntaps = ntaps;
  p.lam = static_cast<float2*>(lam);
  p.xin = static_cast<const float2*>(x);
  p.gy = static_cast<const float2*>(gy);
  p.gx = static_cast<const float2*>(gx);
  p.ws = static_cast<double*>(ws);
  const int w = lanes_runtime(n);
  int blocks = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bind_factors(p, const_cast<void*>(factors), n, nsys, k);
  if (factors != nullptr ? dispatch_bwd_replay(p, &blocks, st) : dispatch_bwd(p, &blocks, st)) return 1;
  const int64_t groups = (int64_t)blocks * kWarps * (32 / w);  // lane groups of the grid that was launched
  launch_bwd_reduce(p.ws, reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + bwd_rows_bytes(n)), groups, n, nsys,
                    transpose_a, ga, gb, gc, ginvgamma, st);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_solve_fwd(int n, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                               int transpose_a, const float* gamma, const void* gamma_z, const float* b,
                               const float* c, void* x, void* y, void* factors, void* stream) {
  return solve_fwd_impl(n, 1, g, k, z, delays, a, transpose_a, gamma, gamma_z, b, c, x, y, factors, stream);
}

extern "C" int dgfdn_solve_fir_fwd(int n, int g, int ntaps, int64_t k, const void* z, const int32_t* delays,
                                   const float* taps, int transpose_a, const float* gamma, const void* gamma_z,
                                   const float* b, const float* c, void* x, void* y, void* stream) {
  DGFDN_CHECK(taps != nullptr, "solve_fir_fwd: null taps");
  return solve_fwd_impl(n, 1, g, k, z, delays, taps, transpose_a, gamma, gamma_z, b, c, x, y, nullptr, stream, taps, ntaps);
}

extern "C" int dgfdn_solve_fir_bwd(int n, int g, int ntaps, int64_t k, const void* z, const int32_t* delays,
                                   const float* taps, int transpose_a, const float* gamma, const void* gamma_z,
                                   const float* c, const void* x, const void* gy, const void* gx, void* lam, double* gb,
                                   double* gc, double* ginvgamma, void* ws, void* stream) {
  DGFDN_CHECK(taps != nullptr && lam != nullptr, "solve_fir_bwd: null taps / lambda");
  return solve_bwd_impl(n, 1, g, k, z, delays, taps, transpose_a, gamma, gamma_z, c, x, gy, gx, nullptr, gb, gc, ginvgamma, ws,
                        nullptr, stream, taps, ntaps, lam);
}

extern "C" int64_t dgfdn_solve_factors_bytes(int n, int64_t k) {
  if (n < 1 || n > DGFDN_MAX_LINES || k < 1) return 0;
  return (int64_t)factor_layout(n, 1, k).total;
}
extern "C" int64_t dgfdn_solve_groups_factors_bytes(int l, int g, int64_t k) {
  if (l < 1 || g < 1 || l * g > DGFDN_MAX_LINES || k < 1) return 0;
  return (int64_t)factor_layout(l, g, k).total;
}

// one row of (n^2 + 3n) doubles per lane group of the largest grid the backward kernel launches
static int64_t bwd_rows_bytes(int n) {
  const int64_t groups = (int64_t)sm_count() * kMaxBlocksPerSm * kWarps * (32 / lanes_runtime(n));
  return groups * ((int64_t)n * n + 3 * (int64_t)n) * (int64_t)sizeof(double);
}
// ... followed by the rows of the first reduction stage (kReduceChunks per system type)
static int64_t bwd_ws_bytes(int n) {
  if (n < 1 || n > DGFDN_MAX_LINES) return 0;
  return bwd_rows_bytes(n) + (int64_t)kReduceChunks * DGFDN_MAX_GROUPS * ((int64_t)n * n + 3 * (int64_t)n) * (int64_t)sizeof(double);
}

template <int NP>
int launch_colorless(const SolveParams& p, int asym, double* loss_part, int max_sms, int* blocks_out, cudaStream_t st) {
  constexpr int W = lanes_for(NP);
  const size_t smem = smem_bytes<NP>(p, true);
  DGFDN_CUDA(cudaFuncSetAttribute(solve_colorless_kernel<NP, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // max_sms > 0: a grid of at most max_sms resident waves' worth of blocks. The fused step runs K1c in the shadow of the
  // receiver kernel on the SMs its clusters cannot use; a grid sized for the whole chip would leave blocks waiting that
  // flood the SMs the receiver kernel frees and hold back the adjoint chain behind them.
  const int bpsm = blocks_per_sm(solve_colorless_kernel<NP, W>, smem);
  int blocks = grid_blocks(p.k * p.nsys, W, bpsm);
  if (max_sms > 0 && blocks > max_sms * bpsm) blocks = max_sms * bpsm;
  *blocks_out = blocks;
  solve_colorless_kernel<NP, W><<<blocks, kWarps * 32, smem, st>>>(p, asym, loss_part);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t dgfdn_solve_colorless_ws_bytes(int l) {
  const int64_t b = bwd_ws_bytes(l);
  if (b == 0) return 0;
  const int64_t groups = (int64_t)sm_count() * kMaxBlocksPerSm * kWarps * (32 / lanes_runtime(l));
  return b + groups * (int64_t)sizeof(double);
}

extern "C" int dgfdn_solve_colorless(int l, int g, int64_t k, const void* z, const int32_t* delays, const float* m_raw,
                                     const float* gamma, const float* b, const float* c, int asym, int max_sms,
                                     double* loss, double* gm, double* gb, double* gc, void* ws, void* stream) {
  if (check_common(l, g, g, k)) return 1;
  DGFDN_CHECK(z && delays && m_raw && b && c && loss && gm && gb && gc && ws, "solve_colorless: null pointer");
  SolveParams p{};
  fill_params(p, l, g, g, k, z, delays, m_raw, 0, gamma, nullptr, b, c);
  p.ws = static_cast<double*>(ws);
  double* loss_part = reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + bwd_ws_bytes(l));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blocks = 0;
  const int np = (l + 3) & ~3;
  int rc = 1;
  switch (np) {
    case 4: rc = launch_colorless<4>(p, asym, loss_part, max_sms, &blocks, st); break;
    case 8: rc = launch_colorless<8>(p, asym, loss_part, max_sms, &blocks, st); break;
    case 12: rc = launch_colorless<12>(p, asym, loss_part, max_sms, &blocks, st); break;
    case 16: rc = launch_colorless<16>(p, asym, loss_part, max_sms, &blocks, st); break;
    default:
      set_error("solve_colorless: at most 16 lines per group (got %d)", l);
      return 1;
  }
  if (rc) return rc;
  const int w = lanes_runtime(l);
  const int64_t groups = (int64_t)blocks * kWarps * (32 / w);
  launch_bwd_reduce(p.ws, reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + bwd_rows_bytes(l)), groups, l, g, 0,
                    gm, gb, gc, nullptr, st);
  DGFDN_LAUNCH_CHECK();
  colorless_loss_reduce_kernel<<<g, 256, 0, st>>>(loss_part, groups, g, k, loss);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t dgfdn_solve_bwd_ws_bytes(int n) { return bwd_ws_bytes(n); }
extern "C" int64_t dgfdn_solve_groups_bwd_ws_bytes(int l) { return bwd_ws_bytes(l); }

extern "C" int dgfdn_solve_bwd(int n, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                               int transpose_a, const float* gamma, const void* gamma_z, const float* c,
                               const void* x, const void* gy, const void* gx, double* ga, double* gb, double* gc,
                               double* ginvgamma, void* ws, const void* factors, void* stream) {
  return solve_bwd_impl(n, 1, g, k, z, delays, a, transpose_a, gamma, gamma_z, c, x, gy, gx, ga, gb, gc, ginvgamma, ws,
                        factors, stream);
}

extern "C" int dgfdn_solve_groups_fwd(int l, int g, int64_t k, const void* z, const int32_t* delays, const float* m_raw,
                                      const float* gamma, const float* b, const float* c, void* x, void* y,
                                      void* factors, void* stream) {
  return solve_fwd_impl(l, g, g, k, z, delays, m_raw, 0, gamma, nullptr, b, c, x, y, factors, stream);
}

extern "C" int dgfdn_solve_groups_bwd(int l, int g, int64_t k, const void* z, const int32_t* delays, const float* m_raw,
                                      const float* gamma, const float* c, const void* x, const void* gy, const void* gx,
                                      double* gm, double* gb, double* gc, double* ginvgamma, void* ws,
                                      const void* factors, void* stream) {
  return solve_bwd_impl(l, g, g, k, z, delays, m_raw, 0, gamma, nullptr, c, x, gy, gx, gm, gb, gc, ginvgamma, ws, factors,
                        stream);
}
