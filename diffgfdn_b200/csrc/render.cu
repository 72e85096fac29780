// K6: block-recursive time-domain renderer of the Grouped FDN (late reverberation tail).
//
// The reference has no recursive renderer: it renders h = irfft(H) (diff_gfdn/utils.py:169) per receiver, sums
// octave bands (run_subband_training_treble.py:316-358) and cross-fades block convolutions for a moving listener
// (sound_examples.py:163-226). The transfer function it samples, H(z) = c^T (D Gamma^-1 - A)^-1 b, is the
// recursion
//     x_i[t] = gamma_i ( sum_j A_ij x_j[t - m_i] + b_i u[t - m_i] ),     q_g[t] = sum_{i in g} c_i x_i[t],
// whose state is receiver independent; a listener only mixes the G group signals with its gains s[r,g].
// With a block of L = min_i m_i samples every right-hand side refers to earlier blocks, so a block is one
// dense (N x N) * (N x L) product. render_groups advances one CTA per (octave) band; render_mix streams the
// listener outputs (4 B written per listener.sample, the only HBM-heavy part).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads) render_groups_kernel(int n, int g, int64_t tlen,
                                                                 const int32_t* __restrict__ delays,
                                                                 const float* __restrict__ a,
                                                                 const float* __restrict__ gamma,
                                                                 const float* __restrict__ b,
                                                                 const float* __restrict__ c,
                                                                 const float* __restrict__ u, float* hist,
                                                                 float* __restrict__ q) {
  __shared__ float s_a[DGFDN_MAX_LINES * DGFDN_MAX_LINES];
  __shared__ float s_gamma[DGFDN_MAX_LINES], s_b[DGFDN_MAX_LINES], s_c[DGFDN_MAX_LINES];
  __shared__ int s_m[DGFDN_MAX_LINES];
  const int band = blockIdx.x;
  const int l = n / g;
  for (int i = threadIdx.x; i < n * n; i += kThreads) s_a[i] = a[(size_t)band * n * n + i];
  for (int i = threadIdx.x; i < n; i += kThreads) {
    s_gamma[i] = gamma ? gamma[band * n + i] : 1.f;
    s_b[i] = b[band * n + i];
    s_c[i] = c[band * n + i];
    s_m[i] = delays[band * n + i];
  }
  __syncthreads();
  int blk = s_m[0];
  for (int i = 1; i < n; ++i) blk = min(blk, s_m[i]);
  float* hb = hist + (size_t)band * tlen * n;
  float* qb = q + (size_t)band * tlen * g;
  for (int64_t s0 = 0; s0 < tlen; s0 += blk) {
    const int len = (int)min((int64_t)blk, tlen - s0);
    // one work item per (sample, line)
    for (int w = threadIdx.x; w < len * n; w += kThreads) {
      const int i = w % n;
      const int64_t t = s0 + w / n;
      const int64_t src = t - s_m[i];
      float v = 0.f;
      if (src >= 0) {
        const float* xr = hb + src * n;
        float acc = s_b[i] * (u ? u[src] : (src == 0 ? 1.f : 0.f));
        for (int j = 0; j < n; ++j) acc = fmaf(s_a[i * n + j], xr[j], acc);
        v = s_gamma[i] * acc;
      }
      hb[t * n + i] = v;
    }
    __syncthreads();  // block-scope ordering of the global writes above
    for (int w = threadIdx.x; w < len * g; w += kThreads) {
      const int gi = w % g;
      const int64_t t = s0 + w / g;
      float acc = 0.f;
      for (int j = 0; j < l; ++j) acc = fmaf(s_c[gi * l + j], hb[t * n + gi * l + j], acc);
      qb[t * g + gi] = acc;
    }
  }
}

// The same recursion with a thread-block CLUSTER per band: the (sample, line) items of a block of min(m) samples are
// spread over kRenderCluster CTAs (8 SMs per band instead of 1), the state history lives in global memory (L2) and one
// barrier.cluster (release / acquire: orders the global writes at cluster scope) separates a block from the next.
// History reads bypass L1 (ld.global.cg): the lines were written by other SMs.
constexpr int kRenderCluster = 8;

__global__ void __launch_bounds__(kThreads) render_groups_cluster_kernel(int n, int g, int64_t tlen,
                                                                         const int32_t* __restrict__ delays,
                                                                         const float* __restrict__ a,
                                                                         const float* __restrict__ gamma,
                                                                         const float* __restrict__ b,
                                                                         const float* __restrict__ c,
                                                                         const float* __restrict__ u, float* hist,
                                                                         float* __restrict__ q) {
  __shared__ float s_a[DGFDN_MAX_LINES * DGFDN_MAX_LINES];
  __shared__ float s_gamma[DGFDN_MAX_LINES], s_b[DGFDN_MAX_LINES], s_c[DGFDN_MAX_LINES];
  __shared__ int s_m[DGFDN_MAX_LINES];
  const int band = blockIdx.x / kRenderCluster;
  const int rank = blockIdx.x % kRenderCluster;
  const int l = n / g;
  for (int i = threadIdx.x; i < n * n; i += kThreads) s_a[i] = a[(size_t)band * n * n + i];
  for (int i = threadIdx.x; i < n; i += kThreads) {
    s_gamma[i] = gamma ? gamma[band * n + i] : 1.f;
    s_b[i] = b[band * n + i];
    s_c[i] = c[band * n + i];
    s_m[i] = delays[band * n + i];
  }
  __syncthreads();
  int blk = s_m[0];
  for (int i = 1; i < n; ++i) blk = min(blk, s_m[i]);
  float* hb = hist + (size_t)band * tlen * n;
  float* qb = q + (size_t)band * tlen * g;
  const int gtid = rank * kThreads + threadIdx.x;
  constexpr int kStride = kRenderCluster * kThreads;
  for (int64_t s0 = 0; s0 < tlen; s0 += blk) {
    const int len = (int)min((int64_t)blk, tlen - s0);
    for (int w = gtid; w < len * n; w += kStride) {
      const int i = w % n;
      const int64_t t = s0 + w / n;
      const int64_t src = t - s_m[i];
      float v = 0.f;
      if (src >= 0) {
        const float* xr = hb + src * n;
        float acc = s_b[i] * (u ? u[src] : (src == 0 ? 1.f : 0.f));
        if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(xr) & 15u) == 0) {  // a state vector is n/4 128-bit L2 loads
          for (int j = 0; j < n; j += 4) {
            const float4 x4 = __ldcg(reinterpret_cast<const float4*>(xr + j));
            acc = fmaf(s_a[i * n + j], x4.x, acc);
            acc = fmaf(s_a[i * n + j + 1], x4.y, acc);
            acc = fmaf(s_a[i * n + j + 2], x4.z, acc);
            acc = fmaf(s_a[i * n + j + 3], x4.w, acc);
          }
        } else {
          for (int j = 0; j < n; ++j) acc = fmaf(s_a[i * n + j], __ldcg(xr + j), acc);
        }
        v = s_gamma[i] * acc;
      }
      hb[t * n + i] = v;
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int w = gtid; w < len * g; w += kStride) {
      const int gi = w % g;
      const int64_t t = s0 + w / g;
      float acc = 0.f;
      for (int j = 0; j < l; ++j) acc = fmaf(s_c[gi * l + j], __ldcg(hb + t * n + gi * l + j), acc);
      qb[t * g + gi] = acc;
    }
  }
}

// out[r,t] = sum_band sum_g s[band, traj[r, t/hop], g] q[band, t, g]; four samples per thread (128-bit stores).
__global__ void __launch_bounds__(256) render_mix_kernel(int bands, int g, int64_t tlen, int64_t positions,
                                                         int64_t hop, int64_t nhops, const float* __restrict__ s,
                                                         const int32_t* __restrict__ traj,
                                                         const float* __restrict__ q, float* __restrict__ out) {
  const int64_t r = blockIdx.y;
  const int64_t t0 = 4 * ((int64_t)blockIdx.x * 256 + threadIdx.x);
  if (t0 >= tlen) return;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t t = t0 + e;
    if (t < tlen) {
      const int64_t pos = traj[r * nhops + t / hop];
      float acc = 0.f;
      for (int bd = 0; bd < bands; ++bd) {
        const float* sp = s + ((size_t)bd * positions + pos) * g;
        const float* qp = q + ((size_t)bd * tlen + t) * g;
        for (int gi = 0; gi < g; ++gi) acc = fmaf(__ldg(sp + gi), __ldg(qp + gi), acc);
      }
      v[e] = acc;
    }
  }
  float* o = out + r * tlen + t0;
  if (t0 + 3 < tlen && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
    st_stream(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
  } else {
    for (int e = 0; e < 4 && t0 + e < tlen; ++e) o[e] = v[e];
  }
}

// Tiled listener mix (hop % 4 == 0): a block owns up to 1024 samples of ONE hop and kMixTile listeners. A thread keeps
// 4 samples x kMixTile listeners of accumulators in registers, reads the 4 x G group samples of a band with G 128-bit
// loads (q is [band][t][g], so 4 consecutive samples are 4 G contiguous floats) and reuses them for every listener
// of the tile; the listeners' gains for this hop sit in shared memory (one broadcast LDS per gain). Per output
// sample that is the bands x G FMAs the sum needs plus ~1/kMixTile of a load: the kernel streams its 4 B per
// listener.sample at HBM write speed instead of re-reading q and s from L2 for every listener.
constexpr int kMixTile = 8;

template <int G>
__global__ void __launch_bounds__(256) render_mix_tiled_kernel(int bands, int64_t tlen, int64_t listeners,
                                                               int64_t positions, int64_t hop, int64_t nhops,
                                                               int blocks_per_hop, const float* __restrict__ s,
                                                               const int32_t* __restrict__ traj,
                                                               const float* __restrict__ q, float* __restrict__ out) {
  extern __shared__ float s_tile[];  // [kMixTile][bands * G]
  const int bg = bands * G;
  const int64_t hidx = blockIdx.x / blocks_per_hop;
  const int sub = blockIdx.x % blocks_per_hop;
  const int64_t r0 = (int64_t)blockIdx.y * kMixTile;
  const int nl = (int)min((int64_t)kMixTile, listeners - r0);
  for (int i = threadIdx.x; i < kMixTile * bg; i += 256) {
    const int l = i / bg, j = i % bg;
    float v = 0.f;
    if (l < nl) {
      const int64_t pos = traj[(r0 + l) * nhops + hidx];
      v = s[((size_t)(j / G) * positions + pos) * G + (j % G)];
    }
    s_tile[i] = v;
  }
  __syncthreads();
  const int64_t hop_end = min(tlen, (hidx + 1) * hop);
  const int64_t t0 = hidx * hop + 4 * ((int64_t)sub * 256 + threadIdx.x);
  if (t0 >= hop_end) return;
  const bool full = t0 + 3 < hop_end;
  float acc[kMixTile][4];
#pragma unroll
  for (int l = 0; l < kMixTile; ++l)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[l][e] = 0.f;
  for (int bd = 0; bd < bands; ++bd) {
    float v[4 * G];
    const float* qp = q + ((size_t)bd * tlen + t0) * G;
    if (full && (reinterpret_cast<uintptr_t>(qp) & 15u) == 0) {  // (a band starts 16-byte aligned only if tlen G % 4 == 0)
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(qp) + i);
        v[4 * i] = w.x, v[4 * i + 1] = w.y, v[4 * i + 2] = w.z, v[4 * i + 3] = w.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4 * G; ++i) v[i] = (t0 + i / G < hop_end) ? __ldg(qp + i) : 0.f;
    }
#pragma unroll
    for (int l = 0; l < kMixTile; ++l) {
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const float w = s_tile[l * bg + bd * G + gi];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[l][e] = fmaf(w, v[e * G + gi], acc[l][e]);
      }
    }
  }
  for (int l = 0; l < nl; ++l) {
    float* o = out + (r0 + l) * tlen + t0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
      st_stream(reinterpret_cast<float4*>(o), make_float4(acc[l][0], acc[l][1], acc[l][2], acc[l][3]));
    } else {
      for (int e = 0; e < 4 && t0 + e < hop_end; ++e) o[e] = acc[l][e];
    }
  }
}

template <int G>
int launch_mix_tiled(int bands, int64_t t, int64_t listeners, int64_t positions, int64_t hop, int64_t nhops,
                     const float* s, const int32_t* traj, const float* q, float* out, cudaStream_t st) {
  const int blocks_per_hop = (int)((std::min(hop, t) + 1023) / 1024);
  const dim3 grid((unsigned)(nhops * blocks_per_hop), (unsigned)((listeners + kMixTile - 1) / kMixTile));
  const size_t smem = (size_t)kMixTile * bands * G * sizeof(float);
  render_mix_tiled_kernel<G><<<grid, 256, smem, st>>>(bands, t, listeners, positions, hop, nhops, blocks_per_hop, s, traj,
                                                      q, out);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_render_groups(int bands, int n, int g, int64_t t, const int32_t* delays, const float* a,
                                   const float* gamma, const float* b, const float* c, const float* u, float* hist,
                                   float* q, void* stream) {
  DGFDN_CHECK(bands >= 1 && n >= 1 && n <= DGFDN_MAX_LINES && g >= 1 && n % g == 0 && t >= 1,
              "render_groups: bad sizes");
  DGFDN_CHECK(delays && a && b && c && hist && q, "render_groups: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (getenv("DGFDN_RENDER_SINGLE_CTA") == nullptr) {  // one cluster of 8 CTAs per band
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kRenderCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(bands * kRenderCluster));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DGFDN_CUDA(cudaLaunchKernelEx(&cfg, render_groups_cluster_kernel, n, g, t, delays, a, gamma, b, c, u, hist, q));
    return 0;
  }
  render_groups_kernel<<<bands, kThreads, 0, st>>>(n, g, t, delays, a, gamma, b, c, u, hist, q);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_render_mix(int bands, int g, int64_t t, int64_t listeners, int64_t positions, int64_t hop,
                                const float* s, const int32_t* traj, const float* q, float* out, void* stream) {
  DGFDN_CHECK(bands >= 1 && g >= 1 && t >= 1 && listeners >= 0 && positions >= 1 && hop >= 1,
              "render_mix: bad sizes");
  if (listeners == 0) return 0;
  DGFDN_CHECK(s && traj && q && out, "render_mix: null pointer");
  DGFDN_CHECK(listeners <= 65535, "render_mix: listeners=%lld exceeds grid.y limit; tile the call",
              (long long)listeners);
  const int64_t nhops = (t + hop - 1) / hop;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tiled kernel: hops of a multiple of 4 samples (128-bit accesses never straddle a hop), 16-byte aligned q, and a
  // grid.x that fits; anything else takes the per-sample kernel below
  const int64_t gx = nhops * ((std::min(hop, t) + 1023) / 1024);
  if (hop % 4 == 0 && g <= 4 && (reinterpret_cast<uintptr_t>(q) & 15u) == 0 && gx < (int64_t)1 << 31 &&
      (size_t)kMixTile * bands * g * sizeof(float) <= 48 * 1024) {
    switch (g) {
      case 1: return launch_mix_tiled<1>(bands, t, listeners, positions, hop, nhops, s, traj, q, out, st);
      case 2: return launch_mix_tiled<2>(bands, t, listeners, positions, hop, nhops, s, traj, q, out, st);
      case 3: return launch_mix_tiled<3>(bands, t, listeners, positions, hop, nhops, s, traj, q, out, st);
      default: return launch_mix_tiled<4>(bands, t, listeners, positions, hop, nhops, s, traj, q, out, st);
    }
  }
  dim3 grid((unsigned)((t + 1023) / 1024), (unsigned)listeners);
  render_mix_kernel<<<grid, 256, 0, st>>>(bands, g, t, positions, hop, nhops, s, traj, q, out);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
