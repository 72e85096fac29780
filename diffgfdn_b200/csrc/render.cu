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
// 4 samples x kMixTile listeners of accumulators in registers (as packed float2 pairs of samples: FFMA2), reads the
// 4 x G group samples of a band with G 128-bit loads (q is [band][t][g], so 4 consecutive samples are 4 G contiguous
// floats) and reuses them for every listener of the tile; the listeners' gains for this hop sit in shared memory as
// [band * G + g][listener], so the 8 gains one (band, g) needs are two 128-bit broadcast loads. Per output sample that
// is bands x G / 2 packed FMAs plus ~1/20 of a load. Measured at BASELINE configs[4] (profiles/r02_render_mix.txt):
// 8 listeners x scalar FMAs 2.75 ms -> FFMA2 over listener pairs 2.53 -> 16 listeners per tile, persistent blocks with
// prefetched gains and group samples 1.86 ms; staging the group samples of 64 listeners in shared memory by TMA
// (a quarter of the L2 reads) measured 2.0 ms and was dropped. The packed-FMA floor of the 24-term sum is 0.87 ms.
constexpr int kMixTile = 16;

template <int G>
__global__ void __launch_bounds__(256) render_mix_tiled_kernel(int bands, int64_t tlen, int64_t listeners,
                                                               int64_t positions, int64_t hop, int64_t nhops,
                                                               const float* __restrict__ s,
                                                               const int32_t* __restrict__ traj,
                                                               const float* __restrict__ q, float* __restrict__ out) {
  // PERSISTENT: a block walks over (hop, listener tile) items, hop-major (concurrent blocks share a hop's q slice in
  // L2). The gains of the NEXT item -- a dependent traj -> s gather, ~2 us of latency -- are fetched into registers
  // while the current item is computed, then parked in the other half of shared memory.
  extern __shared__ __align__(16) float s_tile[];  // [2][bands * G][kMixTile]
  const int bg = bands * G;
  const int64_t ntl = (listeners + kMixTile - 1) / kMixTile;
  const int64_t nitems = nhops * ntl;
  const int tile_floats = kMixTile * bg;
  auto gather = [&](int64_t item, int i) -> float {  // gain i = (band g, listener) of an item
    const int64_t hidx = item / ntl, r0 = (item % ntl) * kMixTile;
    const int j = i / kMixTile, l = i % kMixTile;
    if (r0 + l >= listeners) return 0.f;
    const int64_t pos = __ldg(traj + (r0 + l) * nhops + hidx);
    return __ldg(s + ((size_t)(j / G) * positions + pos) * G + (j % G));
  };
  const int nthreads = blockDim.x;  // chosen on the host so that the threads x iterations cover a hop with little waste
  int cur = 0;
  if ((int64_t)blockIdx.x < nitems)
    for (int i = threadIdx.x; i < tile_floats; i += nthreads) s_tile[i] = gather(blockIdx.x, i);
  __syncthreads();
  const bool aligned = (tlen * G) % 4 == 0 && (reinterpret_cast<uintptr_t>(q) & 15u) == 0;  // every band / row 16-byte aligned
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t next = item + gridDim.x;
    float pre[4] = {0.f, 0.f, 0.f, 0.f};  // tile_floats <= 4 x blockDim (checked on the host)
    if (next < nitems) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if ((int)threadIdx.x + nthreads * u < tile_floats) pre[u] = gather(next, threadIdx.x + nthreads * u);
    }
    const float* tile = s_tile + cur * tile_floats;
    const int64_t hidx = item / ntl, r0 = (item % ntl) * kMixTile;
    const int nl = (int)min((int64_t)kMixTile, listeners - r0);
    const int64_t hop_end = min(tlen, (hidx + 1) * hop);
    for (int64_t t0 = hidx * hop + 4 * (int64_t)threadIdx.x; t0 < hop_end; t0 += 4 * nthreads) {
      const bool full = aligned && t0 + 3 < hop_end;
      auto load_q = [&](int bd, float* v) {  // the 4 x G group samples of band bd at t0..t0+3
        const float* qp = q + ((size_t)bd * tlen + t0) * G;
        if (full) {
#pragma unroll
          for (int i = 0; i < G; ++i) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(qp) + i);
            v[4 * i] = w.x, v[4 * i + 1] = w.y, v[4 * i + 2] = w.z, v[4 * i + 3] = w.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4 * G; ++i) v[i] = (t0 + i / G < hop_end) ? __ldg(qp + i) : 0.f;
        }
      };
      float2 acc[kMixTile / 2][4];  // [listener pair][sample]: the pair's gains are adjacent in shared memory
#pragma unroll
      for (int l = 0; l < kMixTile / 2; ++l)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[l][e] = make_float2(0.f, 0.f);
      float vn[4 * G];
      load_q(0, vn);
      for (int bd = 0; bd < bands; ++bd) {
        float v[4 * G];
#pragma unroll
        for (int i = 0; i < 4 * G; ++i) v[i] = vn[i];
        if (bd + 1 < bands) load_q(bd + 1, vn);  // the next band's samples travel while this band is accumulated
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          const float4* wp = reinterpret_cast<const float4*>(tile + (bd * G + gi) * kMixTile);
          float2 w2[kMixTile / 2];
#pragma unroll
          for (int l = 0; l < kMixTile / 4; ++l) {
            const float4 wv = wp[l];
            w2[2 * l] = make_float2(wv.x, wv.y);
            w2[2 * l + 1] = make_float2(wv.z, wv.w);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x = v[e * G + gi];
            const float2 x2 = make_float2(x, x);
#pragma unroll
            for (int l = 0; l < kMixTile / 2; ++l) acc[l][e] = __ffma2_rn(w2[l], x2, acc[l][e]);
          }
        }
      }
#pragma unroll
      for (int l = 0; l < kMixTile; ++l) {
        if (l < nl) {
          float* o = out + (r0 + l) * tlen + t0;
          const float a0 = (l & 1) ? acc[l >> 1][0].y : acc[l >> 1][0].x, a1 = (l & 1) ? acc[l >> 1][1].y : acc[l >> 1][1].x;
          const float a2 = (l & 1) ? acc[l >> 1][2].y : acc[l >> 1][2].x, a3 = (l & 1) ? acc[l >> 1][3].y : acc[l >> 1][3].x;
          if (full && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
            st_stream(reinterpret_cast<float4*>(o), make_float4(a0, a1, a2, a3));
          } else {
            if (t0 < hop_end) o[0] = a0;
            if (t0 + 1 < hop_end) o[1] = a1;
            if (t0 + 2 < hop_end) o[2] = a2;
            if (t0 + 3 < hop_end) o[3] = a3;
          }
        }
      }
    }
    float* nxt = s_tile + (cur ^ 1) * tile_floats;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if ((int)threadIdx.x + nthreads * u < tile_floats) nxt[threadIdx.x + nthreads * u] = pre[u];
    __syncthreads();
    cur ^= 1;
  }
}

template <int G>
int launch_mix_tiled(int bands, int64_t t, int64_t listeners, int64_t positions, int64_t hop, int64_t nhops,
                     const float* s, const int32_t* traj, const float* q, float* out, cudaStream_t st) {
  const int64_t nitems = nhops * ((listeners + kMixTile - 1) / kMixTile);
  // threads per block: a multiple of 32 <= 256 such that threads x iterations covers the hop's 128-bit words with
  // the least idle lanes (3200 samples: 160 threads x 5 iterations exactly, instead of 256 x 4 with 22 % idle)
  const int64_t n4 = (std::min(hop, t) + 3) / 4;
  int threads = 256;
  double best = 1e30;
  for (int64_t it = (n4 + 255) / 256; it <= (n4 + 255) / 256 + 3; ++it) {
    const int64_t th = std::min<int64_t>(256, std::max<int64_t>(32, ((n4 + it - 1) / it + 31) / 32 * 32));
    const double waste = (double)(th * ((n4 + th - 1) / th)) / (double)n4 + 0.002 * (256 - th) / 32;  // prefer fuller blocks on ties
    if (waste < best) best = waste, threads = (int)th;
  }
  DGFDN_CHECK(kMixTile * bands * G <= 4 * threads, "render_mix: too many (band, group) gains for the tile prefetch");
  const size_t smem = (size_t)2 * kMixTile * bands * G * sizeof(float);
  int per_sm = 0;
  DGFDN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_mix_tiled_kernel<G>, threads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t resident = (int64_t)sm_count() * per_sm;
  render_mix_tiled_kernel<G><<<(unsigned)std::min(nitems, resident), threads, smem, st>>>(bands, t, listeners, positions, hop,
                                                                                          nhops, s, traj, q, out);
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
  if (hop % 4 == 0 && g <= 4 && kMixTile * bands * g <= 4 * 128) {
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
