// K2: receiver projection  H[r,k] = sum_g s[r,g] y[k,g] + d[r,k]  and its adjoint, plus the SH-domain
// variant and the SH -> direction channel mix.
//
// Replaces the (B,N,K) complex expansions and einsums of the reference (diff_gfdn/model.py:583-619,
// 1056-1088; trainer.py:853-865). The contraction length is G (3) -- this is a streaming kernel bound
// by HBM: 8 B read (d) + 8 B written (H) per receiver.bin; y (K*G*8 B, a few MB) stays in L2.
// Each thread owns two consecutive bins (one 128-bit access per row) and walks over a block of rows,
// so y is loaded once per thread and reused from registers.
#include <type_traits>

#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerBlock = 8;

// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kThreads) project_fwd_kernel(int64_t rows, int64_t k, const float* __restrict__ s,
                                                               const float2* __restrict__ y,
                                                               const float2* __restrict__ d, int64_t ldd,
                                                               float2* __restrict__ h, int64_t ldh, bool vec) {
  __shared__ float s_s[kRowsPerBlock * G];
  const int64_t r0 = (int64_t)blockIdx.y * kRowsPerBlock;
  const int nr = (int)min((int64_t)kRowsPerBlock, rows - r0);
  for (int i = threadIdx.x; i < nr * G; i += kThreads) s_s[i] = s[r0 * G + i];
  __syncthreads();
  const int64_t k0 = 2 * ((int64_t)blockIdx.x * kThreads + threadIdx.x);
  if (k0 >= k) return;
  const bool two = (k0 + 1 < k);
  float2 ya[G], yb[G];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    ya[gi] = y[k0 * G + gi];
    yb[gi] = two ? y[(k0 + 1) * G + gi] : make_float2(0.f, 0.f);
  }
  for (int r = 0; r < nr; ++r) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d != nullptr) {
      const float2* dp = d + (r0 + r) * ldd + k0;
      if (vec && two) {
        acc = ld_stream(reinterpret_cast<const float4*>(dp));
      } else {
        float2 a = dp[0];
        acc.x = a.x;
        acc.y = a.y;
        if (two) {
          float2 b = dp[1];
          acc.z = b.x;
          acc.w = b.y;
        }
      }
    }
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const float sv = s_s[r * G + gi];
      acc.x = fmaf(sv, ya[gi].x, acc.x);
      acc.y = fmaf(sv, ya[gi].y, acc.y);
      acc.z = fmaf(sv, yb[gi].x, acc.z);
      acc.w = fmaf(sv, yb[gi].y, acc.w);
    }
    float2* hp = h + (r0 + r) * ldh + k0;
    if (vec && two) {
      *reinterpret_cast<float4*>(hp) = acc;
    } else {
      hp[0] = make_float2(acc.x, acc.y);
      if (two) hp[1] = make_float2(acc.z, acc.w);
    }
  }
}

// gy[k,g] (+)= sum_r s[r,g] gh[r,k]: one thread per bin pair, loop over all rows (coalesced row reads).
template <int G>
__global__ void __launch_bounds__(kThreads) project_bwd_gy_kernel(int64_t rows, int64_t k,
                                                                  const float* __restrict__ s,
                                                                  const float2* __restrict__ gh, int64_t ldh,
                                                                  float2* __restrict__ gy, int accumulate, bool vec) {
  extern __shared__ float s_all[];  // [chunk rows, G]
  const int64_t k0 = 2 * ((int64_t)blockIdx.x * kThreads + threadIdx.x);
  const bool active = k0 < k;
  const bool two = (k0 + 1 < k);
  float2 aa[G], ab[G];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) aa[gi] = ab[gi] = make_float2(0.f, 0.f);
  constexpr int kChunk = 512;
  for (int64_t rc = 0; rc < rows; rc += kChunk) {
    const int nr = (int)min((int64_t)kChunk, rows - rc);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * G; i += kThreads) s_all[i] = s[rc * G + i];
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        const float2* gp = gh + (rc + r) * ldh + k0;
        float4 v;
        if (vec && two) {
          v = ld_stream(reinterpret_cast<const float4*>(gp));
        } else {
          float2 a = gp[0];
          float2 b = two ? gp[1] : make_float2(0.f, 0.f);
          v = make_float4(a.x, a.y, b.x, b.y);
        }
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          const float sv = s_all[r * G + gi];
          aa[gi].x = fmaf(sv, v.x, aa[gi].x);
          aa[gi].y = fmaf(sv, v.y, aa[gi].y);
          ab[gi].x = fmaf(sv, v.z, ab[gi].x);
          ab[gi].y = fmaf(sv, v.w, ab[gi].y);
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    float2* o = gy + k0 * G + gi;
    if (accumulate) {
      float2 p = *o;
      aa[gi].x += p.x;
      aa[gi].y += p.y;
    }
    *o = aa[gi];
    if (two) {
      float2* o2 = gy + (k0 + 1) * G + gi;
      if (accumulate) {
        float2 p = *o2;
        ab[gi].x += p.x;
        ab[gi].y += p.y;
      }
      *o2 = ab[gi];
    }
  }
}

// gs[r,g] = Re sum_k conj(y[k,g]) gh[r,k]: one block per row, float64 accumulation, fixed-order reduce.
template <int G>
__global__ void __launch_bounds__(kThreads) project_bwd_gs_kernel(int64_t k, const float2* __restrict__ y,
                                                                  const float2* __restrict__ gh, int64_t ldh,
                                                                  float* __restrict__ gs) {
  __shared__ double s_red[kThreads / 32][G];
  const int64_t r = blockIdx.x;
  const float2* gp = gh + r * ldh;
  double acc[G];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) acc[gi] = 0.0;
  for (int64_t kk = threadIdx.x; kk < k; kk += kThreads) {
    const float2 v = gp[kk];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      const float2 yv = y[kk * G + gi];
      acc[gi] += (double)(yv.x * v.x + yv.y * v.y);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    double v = warp_sum(acc[gi]);
    if (lane == 0) s_red[warp][gi] = v;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) v += s_red[w][threadIdx.x];
    gs[r * G + threadIdx.x] = (float)v;
  }
}

// ---------------------------------------------------------------------------------------------
// SH projection: H_sh[r,l,k] = sum_g cw[r,g,l] x[k, gL+l].  Block = 128 bins; the x tile is staged in
// shared memory once and reused for every row handled by the block.
constexpr int kShBins = 128;

__global__ void __launch_bounds__(kShBins) project_sh_fwd_kernel(int g, int l, int64_t rows, int64_t k,
                                                                 const float* __restrict__ cw,
                                                                 const float2* __restrict__ x,
                                                                 float2* __restrict__ h, int rows_per_block) {
  extern __shared__ float2 s_x[];  // [kShBins][n+1]
  const int n = g * l;
  const int64_t kb = (int64_t)blockIdx.x * kShBins;
  const int nb = (int)min((int64_t)kShBins, k - kb);
  for (int i = threadIdx.x; i < nb * n; i += kShBins) {
    int bi = i / n, ni = i % n;
    s_x[bi * (n + 1) + ni] = x[(kb + bi) * n + ni];
  }
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  if (threadIdx.x >= nb) return;
  const float2* xr = s_x + threadIdx.x * (n + 1);
  for (int64_t r = r0; r < r1; ++r) {
    const float* cwr = cw + r * n;
    for (int li = 0; li < l; ++li) {
      float2 acc = make_float2(0.f, 0.f);
      for (int gi = 0; gi < g; ++gi) {
        const float w = __ldg(cwr + gi * l + li);
        const float2 xv = xr[gi * l + li];
        acc.x = fmaf(w, xv.x, acc.x);
        acc.y = fmaf(w, xv.y, acc.y);
      }
      h[(r * l + li) * k + kb + threadIdx.x] = acc;
    }
  }
}

// gx[k,n] (+)= sum_r cw[r,n] gh[r,l(n),k]: thread per bin, loops rows; x-tile style staging of the result.
__global__ void __launch_bounds__(kShBins) project_sh_bwd_gx_kernel(int g, int l, int64_t rows, int64_t k,
                                                                    const float* __restrict__ cw,
                                                                    const float2* __restrict__ gh,
                                                                    float2* __restrict__ gx, int accumulate) {
  extern __shared__ float2 s_g[];  // [kShBins][n+1]
  const int n = g * l;
  const int64_t kb = (int64_t)blockIdx.x * kShBins;
  const int nb = (int)min((int64_t)kShBins, k - kb);
  float2* mine = s_g + threadIdx.x * (n + 1);
  for (int ni = 0; ni < n; ++ni) mine[ni] = make_float2(0.f, 0.f);
  if (threadIdx.x < nb) {
    for (int64_t r = 0; r < rows; ++r) {
      const float* cwr = cw + r * n;
      for (int li = 0; li < l; ++li) {
        const float2 v = gh[(r * l + li) * k + kb + threadIdx.x];
        for (int gi = 0; gi < g; ++gi) {
          const float w = __ldg(cwr + gi * l + li);
          float2 a = mine[gi * l + li];
          a.x = fmaf(w, v.x, a.x);
          a.y = fmaf(w, v.y, a.y);
          mine[gi * l + li] = a;
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb * n; i += kShBins) {
    int bi = i / n, ni = i % n;
    float2 v = s_g[bi * (n + 1) + ni];
    float2* o = gx + (kb + bi) * n + ni;
    if (accumulate) {
      float2 p = *o;
      v.x += p.x;
      v.y += p.y;
    }
    *o = v;
  }
}

// gcw[r,g,l] = Re sum_k conj(x[k,gL+l]) gh[r,l,k]: block per (row, l); float64 accumulation.
__global__ void __launch_bounds__(kThreads) project_sh_bwd_gcw_kernel(int g, int l, int64_t k,
                                                                      const float2* __restrict__ x,
                                                                      const float2* __restrict__ gh,
                                                                      float* __restrict__ gcw) {
  __shared__ double s_red[kThreads / 32][DGFDN_MAX_GROUPS];
  const int n = g * l;
  const int64_t r = blockIdx.x / l;
  const int li = (int)(blockIdx.x % l);
  const float2* gp = gh + (r * l + li) * k;
  double acc[DGFDN_MAX_GROUPS];
#pragma unroll
  for (int gi = 0; gi < DGFDN_MAX_GROUPS; ++gi) acc[gi] = 0.0;
  for (int64_t kk = threadIdx.x; kk < k; kk += kThreads) {
    const float2 v = gp[kk];
#pragma unroll
    for (int gi = 0; gi < DGFDN_MAX_GROUPS; ++gi) {
      if (gi < g) {
        const float2 xv = x[kk * n + gi * l + li];
        acc[gi] += (double)(xv.x * v.x + xv.y * v.y);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int gi = 0; gi < DGFDN_MAX_GROUPS; ++gi) {
    double v = warp_sum(acc[gi]);
    if (lane == 0) s_red[warp][gi] = v;
  }
  __syncthreads();
  if (threadIdx.x < g) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) v += s_red[w][threadIdx.x];
    gcw[r * n + threadIdx.x * l + li] = (float)v;
  }
}

// out[r,j,k] = sum_l w[j,l] in[r,l,k]; thread per bin, all channels of one row in registers.
constexpr int kMaxChan = 32;
__global__ void __launch_bounds__(kThreads) mix_channels_kernel(int cin, int cout, int64_t rows, int64_t k,
                                                                const float* __restrict__ w,
                                                                const float2* __restrict__ in,
                                                                float2* __restrict__ out) {
  extern __shared__ float s_w[];  // [cout, cin]
  for (int i = threadIdx.x; i < cin * cout; i += kThreads) s_w[i] = w[i];
  __syncthreads();
  const int64_t kk = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t r = blockIdx.y;
  if (kk >= k) return;
  float2 v[kMaxChan];
#pragma unroll
  for (int c = 0; c < kMaxChan; ++c)
    if (c < cin) v[c] = in[(r * cin + c) * k + kk];
  for (int j = 0; j < cout; ++j) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < kMaxChan; ++c) {
      if (c < cin) {
        const float ww = s_w[j * cin + c];
        acc.x = fmaf(ww, v[c].x, acc.x);
        acc.y = fmaf(ww, v[c].y, acc.y);
      }
    }
    out[(r * cout + j) * k + kk] = acc;
  }
}

template <typename F>
int dispatch_g(int g, F&& f) {
  switch (g) {
    case 1: f(std::integral_constant<int, 1>{}); return 0;
    case 2: f(std::integral_constant<int, 2>{}); return 0;
    case 3: f(std::integral_constant<int, 3>{}); return 0;
    case 4: f(std::integral_constant<int, 4>{}); return 0;
    case 5: f(std::integral_constant<int, 5>{}); return 0;
    case 6: f(std::integral_constant<int, 6>{}); return 0;
    case 7: f(std::integral_constant<int, 7>{}); return 0;
    case 8: f(std::integral_constant<int, 8>{}); return 0;
    default: set_error("project: g=%d out of range [1,%d]", g, DGFDN_MAX_GROUPS); return 1;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_project_fwd(int g, int64_t rows, int64_t k, const float* s, const void* y, const void* d,
                                 int64_t ldd, void* h, int64_t ldh, void* stream) {
  if (rows == 0) return 0;  // an empty batch: nothing to write (empty tensors carry null pointers)
  DGFDN_CHECK(rows >= 0 && k >= 1 && s && y && h, "project_fwd: bad arguments");
  if (rows == 0) return 0;
  DGFDN_CHECK(ldh >= k && (d == nullptr || ldd >= k), "project_fwd: row stride smaller than k");
  const bool vec = aligned16(h) && (ldh % 2 == 0) && (d == nullptr || (aligned16(d) && ldd % 2 == 0));
  dim3 grid((unsigned)((k + 2 * kThreads - 1) / (2 * kThreads)), (unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dispatch_g(g, [&](auto gc) {
        project_fwd_kernel<decltype(gc)::value><<<grid, kThreads, 0, st>>>(
            rows, k, s, static_cast<const float2*>(y), static_cast<const float2*>(d), ldd, static_cast<float2*>(h),
            ldh, vec);
      }))
    return 1;
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_project_bwd(int g, int64_t rows, int64_t k, const float* s, const void* y, const void* gh,
                                 int64_t ldh, void* gy, int accumulate_gy, float* gs, void* stream) {
  DGFDN_CHECK(rows >= 0 && k >= 1 && gh, "project_bwd: bad arguments");
  DGFDN_CHECK(ldh >= k, "project_bwd: row stride smaller than k");
  if (rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = aligned16(gh) && (ldh % 2 == 0);
  if (gy != nullptr) {
    DGFDN_CHECK(s != nullptr, "project_bwd: gy needs s");
    unsigned blocks = (unsigned)((k + 2 * kThreads - 1) / (2 * kThreads));
    if (dispatch_g(g, [&](auto gc) {
          constexpr int G = decltype(gc)::value;
          project_bwd_gy_kernel<G><<<blocks, kThreads, 512 * G * sizeof(float), st>>>(
              rows, k, s, static_cast<const float2*>(gh), ldh, static_cast<float2*>(gy), accumulate_gy, vec);
        }))
      return 1;
    DGFDN_LAUNCH_CHECK();
  }
  if (gs != nullptr) {
    DGFDN_CHECK(y != nullptr, "project_bwd: gs needs y");
    if (dispatch_g(g, [&](auto gc) {
          project_bwd_gs_kernel<decltype(gc)::value><<<(unsigned)rows, kThreads, 0, st>>>(
              k, static_cast<const float2*>(y), static_cast<const float2*>(gh), ldh, gs);
        }))
      return 1;
    DGFDN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int dgfdn_project_sh_fwd(int g, int l, int64_t rows, int64_t k, const float* cw, const void* x,
                                    void* h_sh, void* stream) {
  DGFDN_CHECK(g >= 1 && g <= DGFDN_MAX_GROUPS && l >= 1 && g * l <= DGFDN_MAX_LINES, "project_sh_fwd: bad g/l");
  DGFDN_CHECK(rows >= 0 && k >= 1 && cw && x && h_sh, "project_sh_fwd: bad arguments");
  if (rows == 0) return 0;
  const int n = g * l;
  const int rpb = 16;
  dim3 grid((unsigned)((k + kShBins - 1) / kShBins), (unsigned)((rows + rpb - 1) / rpb));
  size_t smem = (size_t)kShBins * (n + 1) * sizeof(float2);
  project_sh_fwd_kernel<<<grid, kShBins, smem, static_cast<cudaStream_t>(stream)>>>(
      g, l, rows, k, cw, static_cast<const float2*>(x), static_cast<float2*>(h_sh), rpb);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_project_sh_bwd(int g, int l, int64_t rows, int64_t k, const float* cw, const void* x,
                                    const void* gh, void* gx, int accumulate_gx, float* gcw, void* stream) {
  DGFDN_CHECK(g >= 1 && g <= DGFDN_MAX_GROUPS && l >= 1 && g * l <= DGFDN_MAX_LINES, "project_sh_bwd: bad g/l");
  DGFDN_CHECK(rows >= 0 && k >= 1 && gh, "project_sh_bwd: bad arguments");
  if (rows == 0) return 0;
  const int n = g * l;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gx != nullptr) {
    DGFDN_CHECK(cw != nullptr, "project_sh_bwd: gx needs cw");
    size_t smem = (size_t)kShBins * (n + 1) * sizeof(float2);
    project_sh_bwd_gx_kernel<<<(unsigned)((k + kShBins - 1) / kShBins), kShBins, smem, st>>>(
        g, l, rows, k, cw, static_cast<const float2*>(gh), static_cast<float2*>(gx), accumulate_gx);
    DGFDN_LAUNCH_CHECK();
  }
  if (gcw != nullptr) {
    DGFDN_CHECK(x != nullptr, "project_sh_bwd: gcw needs x");
    project_sh_bwd_gcw_kernel<<<(unsigned)(rows * l), kThreads, 0, st>>>(g, l, k, static_cast<const float2*>(x),
                                                                         static_cast<const float2*>(gh), gcw);
    DGFDN_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int dgfdn_mix_channels(int cin, int cout, int64_t rows, int64_t k, const float* w, const void* in,
                                  void* out, void* stream) {
  DGFDN_CHECK(cin >= 1 && cin <= kMaxChan && cout >= 1, "mix_channels: cin=%d must be in [1,%d]", cin, kMaxChan);
  DGFDN_CHECK(rows >= 0 && k >= 1 && w && in && out, "mix_channels: bad arguments");
  if (rows == 0) return 0;
  DGFDN_CHECK(rows <= 65535, "mix_channels: rows=%lld exceeds grid.y limit; tile the call", (long long)rows);
  dim3 grid((unsigned)((k + kThreads - 1) / kThreads), (unsigned)rows);
  mix_channels_kernel<<<grid, kThreads, (size_t)cin * cout * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      cin, cout, rows, k, w, static_cast<const float2*>(in), static_cast<float2*>(out));
  DGFDN_LAUNCH_CHECK();
  return 0;
}
