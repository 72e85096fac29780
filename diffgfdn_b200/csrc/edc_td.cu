// K3c: receiver step in the TIME domain -- mix, energy-decay curve, dB loss and its whole backward in one kernel.
//
// The reference evaluates, per receiver r, H_r = sum_g s[r,g] y_g + d_r over every bin (diff_gfdn/model.py:583-619),
// takes irfft(H_r, n=K)[mix:max_len] and compares Schroeder EDCs in dB (losses.py:187-238). The inverse DFT is
// linear and the receiver enters only through the G real gains s[r,:], so
//
//     h_r[t] = sum_g s[r,g] hy_g[t] + hd_r[t],     hy_g = irfft(y_g)[window],   hd_r = irfft(d_r)[window]
//
// where hy is G rows per step (receiver independent) and hd_r is a constant of the data set (like the target
// EDC), precomputed once. Per receiver and step nothing is left in the frequency domain: this kernel reads
// hd_r and the target EDC in dB (4 B + 4 B per sample), and produces the row loss, dL/ds[r,:] and dL/dh_r.
// The adjoint closes with one G-row contraction (td_contract) and one G-row inverse-DFT adjoint.
//
// Persistent kernel, one CTA (512 threads) per SM, one receiver row at a time per CTA. Pass 1 walks the row late ->
// early in chunks of 8192 samples (four independent 4-sample segments per thread): h (kept in shared memory),
// suffix scan of h^2 (float32 inside a segment and a warp, float64 across warps and chunks), dB,
// |target - achieved|, dL/dEDC. Pass 2 walks early -> late: prefix
// scan of dL/dEDC, dL/dh = 2 h cumsum, dot products with hy for dL/ds; it reads no HBM, so it prefetches the CTA's
// next row into L2 meanwhile. Reduction order is fixed (deterministic).
#include <type_traits>

#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kSeg = 4;                    // samples per segment (one 128-bit access)
constexpr int kNSeg = 4;                   // segments per thread and iteration (independent chains: ILP)
constexpr int kSub = kThreads * kSeg;      // 2048 samples: one segment of every thread
constexpr int kChunk = kNSeg * kSub;       // 8192 samples per iteration, one block scan
constexpr float kEpsF = 1.1920928955078125e-07f;   // torch.finfo(float32).eps (reference utils.py:35)
constexpr float kDbPerLog2 = 3.0102999566398120f;  // 10 / log2(10)
constexpr double kDbFactor = 4.342944819032518;    // 10 / ln(10)

struct ScanSmem {
  double in[2][kNSeg][kWarps];   // [parity][segment][warp]
  double out[2][kNSeg][kWarps];
};

// Exclusive scan over the samples of one chunk, given the per-thread totals v[j] of the thread's kNSeg segments
// (segment j covers samples [j kSub + 4 tid, +4) of the chunk). Inside a warp the scan runs in float32 (128
// samples), across warps, sub-chunks and chunks in float64; all kNSeg scans share two barriers. REVERSE sums over
// later samples (suffix), otherwise over earlier ones. off[j] is what to add to the partial sums inside segment j;
// `carry` holds everything outside the chunk and is advanced by the chunk total. `parity` alternates between
// consecutive calls (double-buffered shared memory instead of a third barrier).
template <bool REVERSE>
__device__ __forceinline__ void block_scan_n(const float (&v)[kNSeg], ScanSmem& sm, int parity, double& carry,
                                             float (&off)[kNSeg]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float inc[kNSeg];
#pragma unroll
  for (int j = 0; j < kNSeg; ++j) inc[j] = v[j];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const bool take = REVERSE ? (lane + o < 32) : (lane >= o);
#pragma unroll
    for (int j = 0; j < kNSeg; ++j) {
      const float t = REVERSE ? __shfl_down_sync(0xffffffffu, inc[j], o) : __shfl_up_sync(0xffffffffu, inc[j], o);
      if (take) inc[j] += t;
    }
  }
  if (lane == (REVERSE ? 0 : 31)) {
#pragma unroll
    for (int j = 0; j < kNSeg; ++j) sm.in[parity][j][warp] = (double)inc[j];
  }
  __syncthreads();
  if (warp < kNSeg) {  // warp j scans the kWarps warp totals of segment j
    double w = lane < kWarps ? sm.in[parity][warp][lane] : 0.0;
#pragma unroll
    for (int o = 1; o < kWarps; o <<= 1) {
      const double t = REVERSE ? __shfl_down_sync(0xffffffffu, w, o) : __shfl_up_sync(0xffffffffu, w, o);
      if (REVERSE ? (lane + o < 32) : (lane >= o)) w += t;
    }
    if (lane < kWarps) sm.out[parity][warp][lane] = w;
  }
  __syncthreads();
  double base = carry;
#pragma unroll
  for (int jj = 0; jj < kNSeg; ++jj) {
    const int j = REVERSE ? kNSeg - 1 - jj : jj;  // sub-chunks in scan order
    double w, tot;
    if (REVERSE) {
      w = (warp < kWarps - 1) ? sm.out[parity][j][warp + 1] : 0.0;
      tot = sm.out[parity][j][0];
    } else {
      w = (warp > 0) ? sm.out[parity][j][warp - 1] : 0.0;
      tot = sm.out[parity][j][kWarps - 1];
    }
    off[j] = (float)(w + base) + (inc[j] - v[j]);
    base += tot;
  }
  carry = base;
}

struct TdParams {
  int64_t rows;
  int tn;
  const float* s;     // [rows, G]
  const float* hy;    // [G, tn]
  const float* hd;    // [rows, ldhd] or null
  int64_t ldhd;
  const float* tdb;   // [rows, ldt]
  int64_t ldt;
  const float* mask;  // [tn] or null
  double coef;
  double* row_sum;    // [rows]
  float* gs;          // [rows, G]
  float* gh;          // [rows, ldg]
  int64_t ldg;
};

// One instruction pulls `bytes` (multiple of 16) starting at p (16-byte aligned) into L2.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

enum LoadKind { kStream, kOwn, kRead };  // read-once HBM stream / re-read of this thread's own store / read-only

// One segment (4 consecutive samples) at q[0..3]; `t` is its first sample index. Samples at or beyond tn read as 0.
// VEC: tn % 4 == 0 and every row is 16-byte aligned, so a segment is either fully inside or fully outside.
// FULL: the caller guarantees t + 4 <= tn (implies VEC): no bounds logic at all.
template <bool VEC, bool FULL, LoadKind KIND>
__device__ __forceinline__ float4 load_seg(const float* __restrict__ q, int t, int tn) {
  if (VEC) {
    if (!FULL && t >= tn) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    return KIND == kStream ? ld_stream(q4) : (KIND == kOwn ? __ldcg(q4) : __ldg(q4));
  }
  float4 v;
  v.x = (t < tn) ? __ldcg(q) : 0.f;
  v.y = (t + 1 < tn) ? __ldcg(q + 1) : 0.f;
  v.z = (t + 2 < tn) ? __ldcg(q + 2) : 0.f;
  v.w = (t + 3 < tn) ? __ldcg(q + 3) : 0.f;
  return v;
}

template <bool VEC, bool FULL>
__device__ __forceinline__ void store_seg(float* __restrict__ q, int t, int tn, float4 v) {
  if (VEC) {
    if (FULL || t < tn) __stcg(reinterpret_cast<float4*>(q), v);
    return;
  }
  if (t < tn) __stcg(q, v.x);
  if (t + 1 < tn) __stcg(q + 1, v.y);
  if (t + 2 < tn) __stcg(q + 2, v.z);
  if (t + 3 < tn) __stcg(q + 3, v.w);
}

__device__ __forceinline__ float4 suffix4(float4 h) {  // suffix sums of h^2 inside a segment
  float4 s;
  s.w = h.w * h.w;
  s.z = fmaf(h.z, h.z, s.w);
  s.y = fmaf(h.y, h.y, s.z);
  s.x = fmaf(h.x, h.x, s.y);
  return s;
}
__device__ __forceinline__ float4 prefix4(float4 g) {
  float4 p;
  p.x = g.x;
  p.y = p.x + g.y;
  p.z = p.y + g.z;
  p.w = p.z + g.w;
  return p;
}

// dB of the EDC samples of one segment, masked |target - dB| and dL/dEDC. `suf` holds the suffix sums of h^2 inside
// the segment, `offe` everything later than the segment plus eps. 10 log10(EDC + eps) >= -69.2 dB, so the
// reference's clip at -200 dB (utils.py:38-40) can never bind and is not evaluated.
template <bool MASKED>
__device__ __forceinline__ float4 db_loss_seg(float4 suf, float offe, float4 td, float4 mk, float cf, float& acc) {
  float4 ge;
#define DGFDN_DB_ONE(C)                                                                         \
  {                                                                                             \
    const float x = suf.C + offe;                                                               \
    const float diff = td.C - kDbPerLog2 * __log2f(x);                                          \
    const float w = MASKED ? mk.C : 1.f;                                                        \
    acc = fmaf(w, fabsf(diff), acc);                                                            \
    const float g = __fdividef(MASKED ? w * cf : cf, x);                                        \
    /* -sign(diff) g: flip the sign of g where diff > 0 */                                      \
    const float sg = __int_as_float(__float_as_int(g) ^ (~__float_as_int(diff) & 0x80000000));  \
    ge.C = diff == 0.f ? 0.f : sg;                                                              \
  }
  DGFDN_DB_ONE(x) DGFDN_DB_ONE(y) DGFDN_DB_ONE(z) DGFDN_DB_ONE(w)
#undef DGFDN_DB_ONE
  return ge;
}

// Per-row, per-thread base pointers (already offset by 4 tid); chunk c / segment j add c kChunk + j kSub floats.
template <int G>
struct RowPtrs {
  const float* hd;  // may be null
  const float* td;
  float* gr;
  const float* hy[G];
  const float* mask;  // may be null
};

template <int G, bool VEC, bool FULL>
__device__ __forceinline__ void mix_chunk(const RowPtrs<G>& rp, const float (&sv)[G], int base, int t0, int tn,
                                          float4 (&h)[kNSeg]) {
#pragma unroll
  for (int j = 0; j < kNSeg; ++j)
    h[j] = rp.hd ? load_seg<VEC, FULL, kStream>(rp.hd + base + j * kSub, t0 + j * kSub, tn)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int j = 0; j < kNSeg; ++j) {
      const float4 y = load_seg<VEC, FULL, kRead>(rp.hy[g] + base + j * kSub, t0 + j * kSub, tn);
      h[j].x = fmaf(sv[g], y.x, h[j].x);
      h[j].y = fmaf(sv[g], y.y, h[j].y);
      h[j].z = fmaf(sv[g], y.z, h[j].z);
      h[j].w = fmaf(sv[g], y.w, h[j].w);
    }
  }
}

template <int G, bool H_IN_SMEM, bool VEC, bool FULL>
__device__ __forceinline__ void pass1_chunk(const RowPtrs<G>& rp, const float (&sv)[G], int c, int tn, float cf,
                                            float4* s_h4, ScanSmem& sm, double& carry, float& acc) {
  const int base = c * kChunk;                // offset of the chunk from the per-thread row pointers
  const int t0 = base + threadIdx.x * kSeg;   // first sample of segment 0
  if (VEC && c > 0 && threadIdx.x == 0) {     // next iteration's HBM inputs -> L2 (one bulk prefetch per stream)
    if (rp.hd) prefetch_l2_bulk(rp.hd + base - kChunk, kChunk * 4);
    prefetch_l2_bulk(rp.td + base - kChunk, kChunk * 4);
  }
  float4 h[kNSeg], td[kNSeg];
  mix_chunk<G, VEC, FULL>(rp, sv, base, t0, tn, h);
#pragma unroll
  for (int j = 0; j < kNSeg; ++j) td[j] = load_seg<VEC, FULL, kStream>(rp.td + base + j * kSub, t0 + j * kSub, tn);
  float4 suf[kNSeg];
  float tot[kNSeg], off[kNSeg];
#pragma unroll
  for (int j = 0; j < kNSeg; ++j) {
    if (H_IN_SMEM && (FULL || t0 + j * kSub < tn)) s_h4[(t0 + j * kSub) >> 2] = h[j];  // zeros beyond tn in a ragged tail
    suf[j] = suffix4(h[j]);
    tot[j] = suf[j].x;
  }
  block_scan_n<true>(tot, sm, c & 1, carry, off);
#pragma unroll
  for (int j = 0; j < kNSeg; ++j) {
    const int t = t0 + j * kSub;
    float4 ge;
    if (FULL && rp.mask == nullptr) {
      ge = db_loss_seg<false>(suf[j], off[j] + kEpsF, td[j], td[j], cf, acc);
    } else {
      float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
      if (rp.mask != nullptr) mk = load_seg<VEC, FULL, kRead>(rp.mask + base + j * kSub, t, tn);
      if (!FULL) {  // samples beyond tn must not contribute
        if (t >= tn) mk.x = 0.f;
        if (t + 1 >= tn) mk.y = 0.f;
        if (t + 2 >= tn) mk.z = 0.f;
        if (t + 3 >= tn) mk.w = 0.f;
      }
      ge = db_loss_seg<true>(suf[j], off[j] + kEpsF, td[j], mk, cf, acc);
    }
    store_seg<VEC, FULL>(rp.gr + base + j * kSub, t, tn, ge);
  }
}

template <int G, bool H_IN_SMEM, bool VEC, bool FULL>
__device__ __forceinline__ void pass2_chunk(const RowPtrs<G>& rp, const float* hdn, const float* tdn,
                                            const float (&sv)[G], int c, int nchunks, int tn, bool want_gs,
                                            const float4* s_h4, ScanSmem& sm, double& carry, float (&gsacc)[G]) {
  const int base = c * kChunk;
  const int t0 = base + threadIdx.x * kSeg;
  if (VEC && threadIdx.x == 0) {
    // pull the chunk of the CTA's NEXT row that its pass 1 will consume at the mirrored position into L2
    // (hdn / tdn are that row's un-offset pointers; pass 2 itself reads no HBM)
    const int pb = (nchunks - 1 - c) * kChunk;
    const unsigned bytes = (unsigned)(min(kChunk, tn - pb) * 4);
    if (tdn != nullptr) prefetch_l2_bulk(tdn + pb, bytes);
    if (hdn != nullptr) prefetch_l2_bulk(hdn + pb, bytes);
  }
  float4 ge[kNSeg], h[kNSeg], pre[kNSeg];
  float tot[kNSeg], off[kNSeg];
#pragma unroll
  for (int j = 0; j < kNSeg; ++j)  // written by this same thread in pass 1
    ge[j] = load_seg<VEC, FULL, kOwn>(rp.gr + base + j * kSub, t0 + j * kSub, tn);
  if (H_IN_SMEM) {
#pragma unroll
    for (int j = 0; j < kNSeg; ++j)
      h[j] = (FULL || t0 + j * kSub < tn) ? s_h4[(t0 + j * kSub) >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    mix_chunk<G, VEC, FULL>(rp, sv, base, t0, tn, h);
  }
#pragma unroll
  for (int j = 0; j < kNSeg; ++j) {
    pre[j] = prefix4(ge[j]);
    tot[j] = pre[j].w;
  }
  block_scan_n<false>(tot, sm, c & 1, carry, off);
#pragma unroll
  for (int j = 0; j < kNSeg; ++j) {
    float4 o;
    o.x = 2.f * h[j].x * (pre[j].x + off[j]);
    o.y = 2.f * h[j].y * (pre[j].y + off[j]);
    o.z = 2.f * h[j].z * (pre[j].z + off[j]);
    o.w = 2.f * h[j].w * (pre[j].w + off[j]);
    store_seg<VEC, FULL>(rp.gr + base + j * kSub, t0 + j * kSub, tn, o);
    h[j] = o;  // keep dL/dh for the dot products below
  }
  if (want_gs) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float a = gsacc[g];
#pragma unroll
      for (int j = 0; j < kNSeg; ++j) {
        const float4 y = load_seg<VEC, FULL, kRead>(rp.hy[g] + base + j * kSub, t0 + j * kSub, tn);
        a = fmaf(h[j].x, y.x, a);
        a = fmaf(h[j].y, y.y, a);
        a = fmaf(h[j].z, y.z, a);
        a = fmaf(h[j].w, y.w, a);
      }
      gsacc[g] = a;
    }
  }
}

// Persistent: CTA b handles rows b, b + gridDim.x, ...  H_IN_SMEM keeps the row's h (tn floats) in shared memory
// between the two passes; otherwise pass 2 rebuilds it from hd and hy.
template <int G, bool H_IN_SMEM, bool VEC>
__global__ void __launch_bounds__(kThreads, 1) td_edc_step_kernel(TdParams p) {
  extern __shared__ float4 s_h4[];  // [ceil(tn/4)] segments when H_IN_SMEM
  __shared__ ScanSmem sm;
  __shared__ double red[kWarps][G + 1];
  const int tn = p.tn;
  const int tid = threadIdx.x;
  const int nchunks = (tn + kChunk - 1) / kChunk;
  const int nfull = VEC ? tn / kChunk : 0;  // chunks that need no bounds logic
  const float cf = (float)(p.coef * kDbFactor);

  for (int64_t r = blockIdx.x; r < p.rows; r += gridDim.x) {
    float sv[G];
    RowPtrs<G> rp;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sv[g] = p.s[r * G + g];
      rp.hy[g] = p.hy + (int64_t)g * tn + tid * kSeg;
    }
    rp.hd = p.hd ? p.hd + r * p.ldhd + tid * kSeg : nullptr;
    rp.td = p.tdb + r * p.ldt + tid * kSeg;
    rp.gr = p.gh + r * p.ldg + tid * kSeg;
    rp.mask = p.mask ? p.mask + tid * kSeg : nullptr;
    // the row this CTA handles next: its inputs are pulled into L2 while pass 2 (which reads no HBM) runs
    const int64_t rn = r + gridDim.x;
    const float* hdn = (rn < p.rows && p.hd) ? p.hd + rn * p.ldhd : nullptr;
    const float* tdn = (rn < p.rows) ? p.tdb + rn * p.ldt : nullptr;

    // ---- pass 1, late -> early: EDC[t] = sum_{tau >= t} h^2, loss, dL/dEDC -> gr
    double carry = 0.0;
    float acc = 0.f;
    for (int c = nchunks - 1; c >= 0; --c) {
      if (c < nfull)
        pass1_chunk<G, H_IN_SMEM, VEC, VEC>(rp, sv, c, tn, cf, s_h4, sm, carry, acc);
      else
        pass1_chunk<G, H_IN_SMEM, VEC, false>(rp, sv, c, tn, cf, s_h4, sm, carry, acc);
    }
    __syncthreads();  // s_h4 complete; scan buffers of either parity idle

    // ---- pass 2, early -> late: dL/dh[tau] = 2 h[tau] sum_{t <= tau} dL/dEDC[t]; dL/ds_g = <dL/dh, hy_g>
    float gsacc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) gsacc[g] = 0.f;
    carry = 0.0;
    const bool want_gs = p.gs != nullptr;
    for (int c = 0; c < nchunks; ++c) {
      if (c < nfull)
        pass2_chunk<G, H_IN_SMEM, VEC, VEC>(rp, hdn, tdn, sv, c, nchunks, tn, want_gs, s_h4, sm, carry, gsacc);
      else
        pass2_chunk<G, H_IN_SMEM, VEC, false>(rp, hdn, tdn, sv, c, nchunks, tn, want_gs, s_h4, sm, carry, gsacc);
    }

    // ---- block reduction of the row loss and of dL/ds[r, :] (float64, fixed order)
    const int lane = tid & 31, warp = tid >> 5;
    {
      const double v = warp_sum((double)acc);
      if (lane == 0) red[warp][G] = v;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const double v = warp_sum((double)gsacc[g]);
      if (lane == 0) red[warp][g] = v;
    }
    __syncthreads();
    if (tid <= G) {
      double v = 0.0;
      for (int w = 0; w < kWarps; ++w) v += red[w][tid];
      if (tid == G) {
        if (p.row_sum) p.row_sum[r] = v;
      } else if (p.gs != nullptr) {
        p.gs[r * G + tid] = (float)v;
      }
    }
    __syncthreads();  // red / s_h4 / scan buffers are reused by the next row
  }
}

// h[r,t] = sum_g s[r,g] hy[g,t] + hd[r,t]  (the receiver's late RIR window itself; inference / validation)
template <int G>
__global__ void __launch_bounds__(256) td_mix_kernel(int64_t rows, int64_t tn, const float* __restrict__ s,
                                                     const float* __restrict__ hy, const float* __restrict__ hd,
                                                     int64_t ldhd, float* __restrict__ h, int64_t ldh) {
  const int64_t r = blockIdx.y;
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= tn) return;
  float acc = hd ? hd[r * ldhd + t] : 0.f;
#pragma unroll
  for (int g = 0; g < G; ++g) acc = fmaf(__ldg(s + r * G + g), __ldg(hy + (int64_t)g * tn + t), acc);
  h[r * ldh + t] = acc;
}

// partial[split][g][t] = sum_{r in split} s[r,g] gh[r,t]; one thread per 4 samples, rows streamed once.
constexpr int kCThreads = 256;
constexpr int kCRows = 256;
template <int G>
__global__ void __launch_bounds__(kCThreads) td_contract_kernel(int64_t rows, int64_t tn, const float* __restrict__ s,
                                                                const float* __restrict__ gh, int64_t ldg,
                                                                float* __restrict__ partial, int64_t rows_per_split,
                                                                bool vec) {
  __shared__ float s_s[kCRows * G];
  const int64_t t0 = 4 * ((int64_t)blockIdx.x * kCThreads + threadIdx.x);
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(rows, r_begin + rows_per_split);
  const bool active = t0 < tn;
  const bool full = vec && (t0 + 4 <= tn);
  float4 acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t rc = r_begin; rc < r_end; rc += kCRows) {
    const int nr = (int)min((int64_t)kCRows, r_end - rc);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * G; i += kCThreads) s_s[i] = s[rc * G + i];
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        const float* gp = gh + (rc + r) * ldg + t0;
        float4 v;
        if (full) {
          v = ld_stream(reinterpret_cast<const float4*>(gp));
        } else {
          v.x = gp[0];
          v.y = (t0 + 1 < tn) ? gp[1] : 0.f;
          v.z = (t0 + 2 < tn) ? gp[2] : 0.f;
          v.w = (t0 + 3 < tn) ? gp[3] : 0.f;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float sv = s_s[r * G + g];
          acc[g].x = fmaf(sv, v.x, acc[g].x);
          acc[g].y = fmaf(sv, v.y, acc[g].y);
          acc[g].z = fmaf(sv, v.z, acc[g].z);
          acc[g].w = fmaf(sv, v.w, acc[g].w);
        }
      }
    }
  }
  if (!active) return;
  float* out = partial + (int64_t)blockIdx.y * G * tn;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float* o = out + (int64_t)g * tn + t0;
    o[0] = acc[g].x;
    if (t0 + 1 < tn) o[1] = acc[g].y;
    if (t0 + 2 < tn) o[2] = acc[g].z;
    if (t0 + 3 < tn) o[3] = acc[g].w;
  }
}

__global__ void td_contract_reduce_kernel(int64_t total, int nsplit, const float* __restrict__ partial,
                                          float* __restrict__ ghy, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = accumulate ? ghy[i] : 0.f;
  for (int sp = 0; sp < nsplit; ++sp) v += partial[(int64_t)sp * total + i];
  ghy[i] = v;
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
    default: set_error("td: g=%d out of range [1,%d]", g, DGFDN_MAX_GROUPS); return 1;
  }
}

inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int contract_splits(int64_t rows, int64_t tn) {
  const int64_t tblocks = (tn + 4 * kCThreads - 1) / (4 * kCThreads);
  int64_t want = (4 * (int64_t)sm_count() + tblocks - 1) / tblocks;
  if (want > 64) want = 64;
  const int64_t max_by_rows = (rows + 63) / 64;  // at least 64 rows per split
  if (want > max_by_rows) want = max_by_rows;
  return (int)(want < 1 ? 1 : want);
}

constexpr size_t kMaxHSmem = 200 * 1024;

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_td_edc_step(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd,
                                 int64_t ldhd, const float* target_db, int64_t ldt, const float* mask, double coef,
                                 double* row_sum, float* gs, float* gh, int64_t ldg, void* stream) {
  DGFDN_CHECK(rows >= 0 && tn >= 1 && s && hy && target_db && gh, "td_edc_step: bad arguments");
  DGFDN_CHECK(tn < ((int64_t)1 << 30), "td_edc_step: tn too large");
  DGFDN_CHECK(ldt >= tn && ldg >= tn && (hd == nullptr || ldhd >= tn), "td_edc_step: row stride smaller than tn");
  if (rows == 0) return 0;
  TdParams p{};
  p.rows = rows;
  p.tn = (int)tn;
  p.s = s;
  p.hy = hy;
  p.hd = hd;
  p.ldhd = ldhd;
  p.tdb = target_db;
  p.ldt = ldt;
  p.mask = mask;
  p.coef = coef;
  p.row_sum = row_sum;
  p.gs = gs;
  p.gh = gh;
  p.ldg = ldg;
  const bool vec = aligned16(hy) && aligned16(hd) && aligned16(target_db) && aligned16(gh) && aligned16(mask) &&
                   tn % 4 == 0 && ldt % 4 == 0 && ldg % 4 == 0 && (hd == nullptr || ldhd % 4 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t hbytes = (size_t)((tn + 3) / 4) * sizeof(float4);
  const bool in_smem = hbytes <= kMaxHSmem;
  const unsigned grid = (unsigned)(rows < (int64_t)sm_count() ? rows : sm_count());  // persistent, one CTA per SM
  int rc = 0;
  auto launch = [&](auto kern, size_t smem) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxHSmem) != cudaSuccess) {
      set_error("td_edc_step: cannot raise the shared-memory limit");
      rc = 1;
      return;
    }
    kern<<<grid, kThreads, smem, st>>>(p);
  };
  if (dispatch_g(g, [&](auto gc) {
        constexpr int G = decltype(gc)::value;
        if (in_smem && vec) launch(td_edc_step_kernel<G, true, true>, hbytes);
        else if (in_smem) launch(td_edc_step_kernel<G, true, false>, hbytes);
        else if (vec) launch(td_edc_step_kernel<G, false, true>, 0);
        else launch(td_edc_step_kernel<G, false, false>, 0);
      }))
    return 1;
  if (rc) return rc;
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_td_mix(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd,
                            int64_t ldhd, float* h, int64_t ldh, void* stream) {
  DGFDN_CHECK(rows >= 0 && tn >= 1 && s && hy && h, "td_mix: bad arguments");
  DGFDN_CHECK(ldh >= tn && (hd == nullptr || ldhd >= tn), "td_mix: row stride smaller than tn");
  if (rows == 0) return 0;
  DGFDN_CHECK(rows <= 65535, "td_mix: rows=%lld exceeds grid.y limit; tile the call", (long long)rows);
  dim3 grid((unsigned)((tn + 255) / 256), (unsigned)rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dispatch_g(g, [&](auto gc) {
        td_mix_kernel<decltype(gc)::value><<<grid, 256, 0, st>>>(rows, tn, s, hy, hd, ldhd, h, ldh);
      }))
    return 1;
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t dgfdn_td_contract_ws_bytes(int g, int64_t rows, int64_t tn) {
  if (g < 1 || rows < 1 || tn < 1) return 0;
  return (int64_t)contract_splits(rows, tn) * g * tn * (int64_t)sizeof(float);
}

extern "C" int dgfdn_td_contract(int g, int64_t rows, int64_t tn, const float* s, const float* gh, int64_t ldg,
                                 float* ghy, int accumulate, void* ws, void* stream) {
  DGFDN_CHECK(rows >= 0 && tn >= 1 && s && gh && ghy && ws, "td_contract: bad arguments");
  DGFDN_CHECK(ldg >= tn, "td_contract: row stride smaller than tn");
  if (rows == 0) return 0;
  const int nsplit = contract_splits(rows, tn);
  const int64_t rps = (rows + nsplit - 1) / nsplit;
  const bool vec = aligned16(gh) && ldg % 4 == 0;
  dim3 grid((unsigned)((tn + 4 * kCThreads - 1) / (4 * kCThreads)), (unsigned)nsplit);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(ws);
  if (dispatch_g(g, [&](auto gc) {
        td_contract_kernel<decltype(gc)::value><<<grid, kCThreads, 0, st>>>(rows, tn, s, gh, ldg, partial, rps, vec);
      }))
    return 1;
  DGFDN_LAUNCH_CHECK();
  const int64_t total = (int64_t)g * tn;
  td_contract_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, nsplit, partial, ghy, accumulate);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
