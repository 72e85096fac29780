// K3t: the receiver step (same arithmetic as K3c / K3d, see edc_td_fused.cu) as a TIME-SLICED persistent kernel.
//
//   h_r[t]  = sum_g s[r,g] hy_g[t] + hd_r[t]                     (model.py:583-619 by linearity of irfft)
//   EDC_r[t] = sum_{tau >= t} h_r[tau]^2 ;  L += sum_t w[t] |target_dB[r,t] - 10 log10(EDC_r[t] + eps)|
//                                                               (losses.py:187-238, utils.py:16-40)
//   dL/dh_r[t] = 2 h_r[t] sum_{t' <= t} dL/dEDC_r[t']           (adjoint of the reversed cumsum)
//   gs[r,g]  = <dL/dh_r, hy_g> ;   ghy[g,t] (+)= sum_r s[r,g] dL/dh_r[t]
//
// K3d gives a receiver row to a cluster of 8 CTAs; only 15 such clusters are co-resident on a B200 (120 of 148 SMs)
// and every row pair costs two CTA barriers + two cluster exchanges in lock step. Here the TIME axis is cut into
// NS <= #SM slices of S samples and CTA c owns slice c for EVERY row: its ghy accumulators are G x S floats (in
// registers), its hy slice is constant (registers), and nothing in the CTA is ever synchronised across warps inside
// the row loop. What a row needs from the other slices are two scalars -- the sum of h^2 over the later slices and the
// sum of dL/dEDC over the earlier ones. Every CTA publishes its slice totals in L2 and sums the ones it needs:
//
//   A  h = hd + s.hy (kept in the hd slot), slice total of h^2                              -> T1[r][c]
//   B  EDC = sum_{c' > c} T1[r][c'] + local suffix, dB, |.|, dL/dEDC (local scan kept in the target slot); total -> T2[r][c]
//   C  dL/dh = h (sum_{c' < c} T2[r][c'] + local prefix); ghy accumulators; <dL/dh, hy_g>   -> part_gs[r][c]
//
// A value doubles as its own flag (0xFFFFFFFF = not there yet; the finalize kernel puts the flags back). A worker warp
// handles RPW = 32/LPR rows per task (LPR lanes per row, NV float4 per lane, in-lane serial scans + one shuffle scan
// over the LPR lanes) and runs its tasks as a dynamically scheduled pipeline: each trip of the loop issues the
// (non-blocking) reads of the totals the oldest B and C candidates wait for, runs A of the next loaded task meanwhile,
// then B / C if their totals have all arrived. A hop through L2 takes ~2 us under load; the rings (6 hd slots, 4
// target-dB slots per warp, filled by 1-D TMA bulk copies as soon as C frees them) hold the rows in flight meanwhile.
//
// HBM traffic: the algorithmic 8 B per receiver.sample plus ~4 KB of totals / partials per row (< 1.5 %).
// Deterministic: every reduction has a fixed order. The grid is launched cooperatively (all NS CTAs co-resident).
#include <type_traits>

#include "common.cuh"
#include "edc_td_sliced.cuh"

namespace dgfdn {
namespace {

constexpr uint32_t kFlag = 0xFFFFFFFFu;             // "not published yet" (a NaN pattern; totals are sanitised)
constexpr float kEpsF = 1.1920928955078125e-07f;    // torch.finfo(float32).eps (reference utils.py:35)
constexpr float kDbPerLog2 = 3.0102999566398120f;   // 10 / log2(10)
constexpr double kDbFactor = 4.342944819032518;     // 10 / ln(10)

// W worker warps; HD / TD slots of the per-warp hd / target-dB rings; HYREG: hy slice in registers (else shared memory)
template <int G_, int LPR_, int NV_, int W_, int HD_, int TD_, bool HYREG_>
struct Shape {
  static constexpr int G = G_, LPR = LPR_, NV = NV_, W = W_;
  static constexpr int RPW = 32 / LPR;            // rows per warp task
  static constexpr int S4 = LPR * NV;             // 128-bit segments per slice
  static constexpr int HD_DEPTH = HD_, TD_DEPTH = TD_;
  static constexpr int THREADS = 32 * W;
  static constexpr int MAXQ = (kSlicedNsp / 4 + LPR - 1) / LPR;  // 128-bit reads per lane of a row's totals
  static constexpr bool HYREG = HYREG_;
  static_assert(NV % 2 == 1, "odd segment count per lane: conflict-free 128-bit shared-memory accesses");
  static_assert(LPR == 8 || LPR == 16 || LPR == 32, "a quarter warp (one 128-bit access phase) must stay inside a row");
  static_assert(G <= HD_DEPTH + TD_DEPTH, "the epilogue reuses the rings for G x S4 partials per row group");
};

struct SlicedParams {
  int64_t rows;
  int tn4;            // tn / 4
  int ns;             // slices = CTAs
  const float* s;     // [rows, G]
  const float* hy;    // [G, tn]
  const float* hd;    // [rows, ldhd] or null
  int64_t ldhd;
  const float* tdb;   // [rows, ldt]
  int64_t ldt;
  const float* mask;  // [tn] or null
  double coef;
  float* rec;         // per-row records: [rows][kSlicedRecFloats] = T1[Nsp] | T2[Nsp] | part_gs[Nsp] (float4)
  double* part_loss;  // [ns]
  float* ghy;         // [G, tn]
  int accumulate;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lg2_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {  // has the phase with this parity completed?
  uint32_t ok;  // (the same answer in every lane: the result feeds warp-uniform branches, see the vote at the call sites)
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint4 ld_l2_v4(const float* p) {  // GPU-scope relaxed load: served by L2, never by a stale L1 line
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_l2(float* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t sane_bits(float v) {  // a NaN total must not look like the flag (nor stall anyone)
  return __float_as_uint(v == v ? v : __int_as_float(0x7f800000));
}

// ---- packed fp32x2 arithmetic on float4 (FFMA2 / FMUL2 / FADD2) ------------------------------------------------
__device__ __forceinline__ float2 lo(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float4 fma4(float a, float4 y, float4 c) {  // a y + c
  const float2 a2 = make_float2(a, a);
  return cat(__ffma2_rn(a2, lo(y), lo(c)), __ffma2_rn(a2, hi(y), hi(c)));
}
__device__ __forceinline__ float4 mul4(float4 x, float4 y) {
  return cat(__fmul2_rn(lo(x), lo(y)), __fmul2_rn(hi(x), hi(y)));
}
__device__ __forceinline__ float4 rsub4(float a, float4 q) {  // a - q
  const float2 a2 = make_float2(a, a), m1 = make_float2(-1.f, -1.f);
  return cat(__ffma2_rn(m1, lo(q), a2), __ffma2_rn(m1, hi(q), a2));
}
__device__ __forceinline__ void dot4(float2& acc, float4 x, float4 y) {  // acc.x + acc.y accumulates <x, y>
  acc = __ffma2_rn(lo(x), lo(y), acc);
  acc = __ffma2_rn(hi(x), hi(y), acc);
}

template <class T>
struct SlicedSmem {  // static part; the rings follow in dynamic shared memory
  unsigned long long bar_hd[T::W][T::HD_DEPTH];
  unsigned long long bar_td[T::W][T::TD_DEPTH];
  float exc1[T::W][T::HD_DEPTH][32];  // per lane: sum of h^2 over the later lanes of its row (lives with the hd slot)
  float exc2[T::W][T::TD_DEPTH][32];  // per lane: sum of dL/dEDC over the earlier lanes of its row + its own (td slot)
  float sv[T::W][T::HD_DEPTH][T::RPW][4];  // receiver gains of the rows in the hd slots (stage A -> stage C)
  double red[T::W];
};

template <class T>
__host__ __device__ constexpr size_t sliced_dyn_smem() {
  // hd ring + td ring + weights (+ hy slice when it is not register resident)
  return (size_t)16 * T::S4 * ((size_t)T::W * T::RPW * (T::HD_DEPTH + T::TD_DEPTH) + 1 + (T::HYREG ? 0 : T::G));
}

// The totals a row group needs from the other slices: up to MAXQ 128-bit words per lane. Word 0 of lane sub == 0 is the
// one that straddles the own slice: `edge` clears its components on the wrong side (all ones for every other word);
// words past the last slice read 0.0 (workspace initialisation), words that do not exist are not loaded (`on` bit).
template <int MAXQ>
struct Totals {
  uint4 q[MAXQ];
  // word k0 + kstep m of the row's totals; a word that does not exist for this lane is replaced by the LAST word of
  // the padded row, which always reads 0.0 (slices >= ns are never published; kSlicedNsp - 4 >= ns is checked on the host)
  __device__ __forceinline__ void load(const float* base, int k0, int kstep) {
#pragma unroll
    for (int m = 0; m < MAXQ; ++m) {
      const int k = k0 + kstep * m;
      q[m] = ld_l2_v4(base + 4 * ((k < 0 || k >= kSlicedNsp / 4) ? kSlicedNsp / 4 - 1 : k));
    }
  }
  // applied when the words are first looked at (not at the load: the loads stay in flight during stage A)
  __device__ __forceinline__ void mask_edge(uint4 edge) { q[0].x &= edge.x, q[0].y &= edge.y, q[0].z &= edge.z, q[0].w &= edge.w; }
  __device__ __forceinline__ bool complete() const {  // no word still reads "not published" (the largest uint32)
    uint32_t mx = 0u;
#pragma unroll
    for (int m = 0; m < MAXQ; ++m) mx = __vimax3_u32(mx, __vimax3_u32(q[m].x, q[m].y, q[m].z), q[m].w);
    return mx != kFlag;
  }
  __device__ __forceinline__ float sum() const {  // fixed order
    float a = 0.f;
#pragma unroll
    for (int m = 0; m < MAXQ; ++m)
      a += (__uint_as_float(q[m].x) + __uint_as_float(q[m].y)) + (__uint_as_float(q[m].z) + __uint_as_float(q[m].w));
    return a;
  }
};

__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp; lets TMA operands stay uniform
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}

// Where a pipeline stage of a worker warp stands: the task it handles next, that task's ring slots and the phase
// parities of their barriers, this lane's row and its record. Advanced incrementally (no divisions in the loop).
struct Cursor {
  int n;              // tasks done
  int hslot, tslot;   // ring slots of task n
  uint32_t hpar, tpar;
  int64_t row;        // this lane's row of task n (may lie past the last row in the last task)
  float* rec;         // record of min(row, rows - 1): a lane without a row reads along with the last valid one
  template <int HD, int TD>
  __device__ __forceinline__ void advance(int64_t row_step, float* rec0, int64_t last_row) {
    ++n;
    if (++hslot == HD) hslot = 0, hpar ^= 1u;
    if (++tslot == TD) tslot = 0, tpar ^= 1u;
    row += row_step;
    rec = rec0 + min(row, last_row) * kSlicedRecFloats;
  }
};

template <class T>
__global__ void __launch_bounds__(T::THREADS, 1) td_sliced_kernel(SlicedParams p) {
  constexpr int G = T::G, LPR = T::LPR, NV = T::NV, W = T::W, RPW = T::RPW, S4 = T::S4, MAXQ = T::MAXQ;
  constexpr int HD = T::HD_DEPTH, TD = T::TD_DEPTH;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ SlicedSmem<T> sm;
  float4* hd_ring = reinterpret_cast<float4*>(dyn_smem);      // [W][HD][RPW][S4]
  float4* td_ring = hd_ring + (size_t)W * HD * RPW * S4;      // [W][TD][RPW][S4]
  float4* wt_s = td_ring + (size_t)W * TD * RPW * S4;         // [S4] loss weights (mask x inside-the-row)
  float4* hy_s = wt_s + S4;                                   // [G][S4] when !HYREG

  const int tid = threadIdx.x, lane = tid & 31;
  const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp index, provably warp-uniform for the compiler
  const int c = blockIdx.x, ns = p.ns;
  const int tn4 = p.tn4;
  const int seg0 = c * S4;
  const int len4 = max(0, min(S4, tn4 - seg0));
  const uint32_t slot_bytes = (uint32_t)len4 * 16u;
  const bool has_hd = p.hd != nullptr;
  const float cf2 = (float)(2.0 * p.coef * kDbFactor);  // the factor 2 of d(h^2) rides on dL/dEDC
  const bool uniw = p.mask == nullptr && len4 == S4;     // no mask and a full slice: every sample weighs 1

  // ---- one-time set-up --------------------------------------------------------------------------------------
  for (int i = tid; i < W * (HD + TD) * RPW * S4; i += T::THREADS) hd_ring[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < S4; i += T::THREADS) {
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < len4) wv = p.mask ? __ldg(reinterpret_cast<const float4*>(p.mask) + seg0 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    wt_s[i] = wv;
    if (!T::HYREG) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        hy_s[g * S4 + i] = i < len4 ? __ldg(reinterpret_cast<const float4*>(p.hy) + (int64_t)g * tn4 + seg0 + i)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (tid < W) {
    for (int d = 0; d < HD; ++d) mbar_init(smem_u32(&sm.bar_hd[tid][d]), 1);
    for (int d = 0; d < TD; ++d) mbar_init(smem_u32(&sm.bar_td[tid][d]), 1);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero-fill / barrier init before TMA writes
  __syncthreads();

  double loss_acc = 0.0;
  float loss_f = 0.f;
  float4 acc[G][NV];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[g][j] = make_float4(0.f, 0.f, 0.f, 0.f);

  {
    const int sub = lane & (LPR - 1), rw = lane / LPR;
    const int64_t ntask = (p.rows + RPW - 1) / RPW;
    const int nmine = ntask > w ? (int)((ntask - w + W - 1) / W) : 0;  // tasks w, w + W, ...
    float4 hyr[T::HYREG ? G : 1][T::HYREG ? NV : 1];
    if (T::HYREG) {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int i = sub * NV + j;
          hyr[g][j] = i < len4 ? __ldg(reinterpret_cast<const float4*>(p.hy) + (int64_t)g * tn4 + seg0 + i)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    auto hyv = [&](int g, int j) -> float4 { return T::HYREG ? hyr[T::HYREG ? g : 0][T::HYREG ? j : 0] : hy_s[g * S4 + sub * NV + j]; };
    float4* my_hd = hd_ring + (size_t)w * HD * RPW * S4;
    float4* my_td = td_ring + (size_t)w * TD * RPW * S4;

    // which 128-bit words of a row's totals this lane reads: the suffix side walks up from the word holding slice
    // c + 1, the prefix side walks down from the word holding slice c - 1 (the straddling word is word 0 of lane 0)
    const int kb1 = (c + 1) / 4 + sub, kb2 = c >= 1 ? (c - 1) / 4 - sub : -1;
    uint4 edge1 = make_uint4(~0u, ~0u, ~0u, ~0u), edge2 = edge1;
    if (sub == 0) {
      const int f1 = c - 4 * ((c + 1) / 4);  // components j <= f1 of the first suffix word are slices <= c
      edge1 = make_uint4(f1 >= 0 ? 0u : ~0u, f1 >= 1 ? 0u : ~0u, f1 >= 2 ? 0u : ~0u, f1 >= 3 ? 0u : ~0u);
      const int f2 = c - 4 * ((c - 1) / 4);  // components j >= f2 of the first prefix word are slices >= c
      edge2 = make_uint4(f2 <= 0 ? 0u : ~0u, f2 <= 1 ? 0u : ~0u, f2 <= 2 ? 0u : ~0u, f2 <= 3 ? 0u : ~0u);
    }

    // TMA refills. Task i of this warp covers rows (w + W i) RPW .. + RPW. One elected lane issues; every operand
    // is warp-uniform (running pointers), so the copies take the uniform datapath without operand waterfalls.
    const int64_t row_step = (int64_t)W * RPW;
    const int64_t last_row = p.rows - 1;
    const uint32_t hd_base = smem_u32(my_hd), td_base = smem_u32(my_td);
    const uint32_t bar_hd0 = smem_u32(&sm.bar_hd[w][0]), bar_td0 = smem_u32(&sm.bar_td[w][0]);
    auto issue = [&](bool is_hd, int i, int slot) {  // converged warp, uniform arguments
      if (i >= nmine || (is_hd && !has_hd)) return;
      const int64_t r0 = (w + (int64_t)W * i) * RPW;
      const int nvalid = (int)min((int64_t)RPW, p.rows - r0);
      const int64_t ld = is_hd ? p.ldhd : p.ldt;
      const float* src = (is_hd ? p.hd : p.tdb) + r0 * ld + (int64_t)seg0 * 4;
      const uint32_t bar = (is_hd ? bar_hd0 : bar_td0) + 8u * (uint32_t)slot;
      const uint32_t dst = (is_hd ? hd_base : td_base) + (uint32_t)(slot * RPW * S4) * 16u;
      if (elect_one()) {
        mbar_expect_tx(bar, slot_bytes * (uint32_t)nvalid);
#pragma unroll
        for (int q = 0; q < RPW; ++q)
          if (q < nvalid) tma_load_1d(dst + (uint32_t)(q * S4) * 16u, src + q * ld, slot_bytes, bar);
      }
    };
    for (int i = 0; i < HD; ++i) issue(true, i, i);
    for (int i = 0; i < TD; ++i) issue(false, i, i);

    Cursor ca, cb, cc;  // stage A / B / C
    ca.n = 0, ca.hslot = ca.tslot = 0, ca.hpar = ca.tpar = 0u;
    ca.row = (int64_t)w * RPW + rw;
    ca.rec = p.rec + min(ca.row, last_row) * kSlicedRecFloats;
    cb = cc = ca;
    float sa[G];  // receiver gains of stage A's next row, fetched one task ahead
#pragma unroll
    for (int g = 0; g < G; ++g) sa[g] = ca.row <= last_row ? __ldg(p.s + ca.row * G + g) : 0.f;
    uint32_t idle = 0;
    while (cc.n < nmine) {
      // ---- non-blocking reads of the totals the oldest B and C candidates wait for (looked at after stage A) -----
      const bool wantB = cb.n < ca.n && cb.n < cc.n + TD, wantC = cc.n < cb.n;
      Totals<MAXQ> t1, t2;
      if (wantB) t1.load(cb.rec, kb1, LPR);
      if (wantC) t2.load(cc.rec + kSlicedNsp, kb2, -LPR);
      bool did = false;

      // ---------------- A: h = hd + s.hy (in place), slice total of h^2 ------------------------------------------
      if (ca.n < nmine && ca.n < cc.n + HD && (!has_hd || __all_sync(0xffffffffu, mbar_try(bar_hd0 + 8u * ca.hslot, ca.hpar)))) {
        const bool valid = ca.row <= last_row;
        float4* hs = my_hd + (ca.hslot * RPW + rw) * S4 + sub * NV;
        float2 sq0 = make_float2(0.f, 0.f), sq1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          float4 v = has_hd ? hs[j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int g = 0; g < G; ++g) v = fma4(sa[g], hyv(g, j), v);
          hs[j] = v;
          sq0 = __ffma2_rn(lo(v), lo(v), sq0);
          sq1 = __ffma2_rn(hi(v), hi(v), sq1);
        }
        if (sub < G) sm.sv[w][ca.hslot][rw][sub] = sub == 0 ? sa[0] : sub == 1 ? sa[G > 1 ? 1 : 0] : sub == 2 ? sa[G > 2 ? 2 : 0] : sa[G > 3 ? 3 : 0];
        const float tot = (sq0.x + sq0.y) + (sq1.x + sq1.y);
        float inc = tot;  // inclusive suffix sum over the lanes of this row (later lanes = later samples)
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) {
          const float t = __shfl_down_sync(0xffffffffu, inc, o, LPR);
          if (sub + o < LPR) inc += t;
        }
        const float nb1 = __shfl_down_sync(0xffffffffu, inc, 1, LPR);
        sm.exc1[w][ca.hslot][lane] = sub < LPR - 1 ? nb1 : 0.f;
        if (valid && sub == 0) st_l2(ca.rec + c, sane_bits(inc));
        ca.advance<HD, TD>(row_step, p.rec, last_row);
#pragma unroll
        for (int g = 0; g < G; ++g) sa[g] = ca.row <= last_row ? __ldg(p.s + ca.row * G + g) : 0.f;
        did = true;
      }

      // ---------------- B: EDC, dB, |.|, dL/dEDC and its in-lane exclusive suffix sums (in place) ----------------
      bool readyB = false;
      if (wantB) {
        t1.mask_edge(edge1);
        readyB = __all_sync(0xffffffffu, t1.complete() && mbar_try(bar_td0 + 8u * cb.tslot, cb.tpar));
      }
      if (readyB) {
        const bool validB = cb.row <= last_row;
        float carry = t1.sum();
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) carry += __shfl_xor_sync(0xffffffffu, carry, o, LPR);
        const float4* hs = my_hd + (cb.hslot * RPW + rw) * S4 + sub * NV;
        float4* ts = my_td + (cb.tslot * RPW + rw) * S4 + sub * NV;
        const float wrow = validB ? 1.f : 0.f;
        const float run0 = (carry + sm.exc1[w][cb.hslot][lane]) + kEpsF;  // EDC + eps of the sample after this lane's last
        float qrun = 0.f, la0 = 0.f, la1 = 0.f;
        auto stage_b = [&](auto uniform_weights) {
          constexpr bool UNIW = decltype(uniform_weights)::value;  // every sample of the slice weighs 1
          const float4* ws = wt_s + sub * NV;
          const float2 nk2 = make_float2(-kDbPerLog2, -kDbPerLog2);
          const float2 cf = make_float2(cf2 * wrow, cf2 * wrow);
          float run = run0;
#pragma unroll
          for (int j = NV - 1; j >= 0; --j) {
            const float4 h = hs[j];
            float4 x;
            x.w = fmaf(h.w, h.w, run);
            x.z = fmaf(h.z, h.z, x.w);
            x.y = fmaf(h.y, h.y, x.z);
            x.x = fmaf(h.x, h.x, x.y);
            run = x.x;
            const float4 td = ts[j];
            // d = target_dB - 10 log10(x) ; dL/dEDC = -sign(d) w 2 coef (10/ln10) / x. An exact tie d == 0 (where
            // torch's abs' is 0) is not special-cased (see edc_td_fused.cu).
            const float2 d01 = __ffma2_rn(nk2, make_float2(lg2_ftz(x.x), lg2_ftz(x.y)), lo(td));
            const float2 d23 = __ffma2_rn(nk2, make_float2(lg2_ftz(x.z), lg2_ftz(x.w)), hi(td));
            float2 c01 = cf, c23 = cf;
            if (UNIW) {
              la0 += fabsf(d01.x);
              la1 += fabsf(d01.y);
              la0 += fabsf(d23.x);
              la1 += fabsf(d23.y);
            } else {
              const float4 wt = ws[j];
              la0 = fmaf(wt.x, fabsf(d01.x), la0);
              la1 = fmaf(wt.y, fabsf(d01.y), la1);
              la0 = fmaf(wt.z, fabsf(d23.x), la0);
              la1 = fmaf(wt.w, fabsf(d23.y), la1);
              c01 = __fmul2_rn(lo(wt), cf);
              c23 = __fmul2_rn(hi(wt), cf);
            }
            const float2 g01 = __fmul2_rn(c01, make_float2(rcp_ftz(x.x), rcp_ftz(x.y)));
            const float2 g23 = __fmul2_rn(c23, make_float2(rcp_ftz(x.z), rcp_ftz(x.w)));
            float4 ge;
            ge.x = __int_as_float(__float_as_int(g01.x) ^ (~__float_as_int(d01.x) & 0x80000000));
            ge.y = __int_as_float(__float_as_int(g01.y) ^ (~__float_as_int(d01.y) & 0x80000000));
            ge.z = __int_as_float(__float_as_int(g23.x) ^ (~__float_as_int(d23.x) & 0x80000000));
            ge.w = __int_as_float(__float_as_int(g23.y) ^ (~__float_as_int(d23.y) & 0x80000000));
            float4 q;  // sum of dL/dEDC over the LATER samples of this lane
            q.w = qrun;
            q.z = q.w + ge.w;
            q.y = q.z + ge.z;
            q.x = q.y + ge.y;
            qrun = q.x + ge.x;
            ts[j] = q;
          }
        };
        if (uniw)
          stage_b(std::true_type{});
        else
          stage_b(std::false_type{});
        loss_f += (la0 + la1) * wrow;
        float inc = qrun;  // inclusive prefix sum over the lanes of this row (earlier lanes = earlier samples)
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, inc, o, LPR);
          if (sub >= o) inc += t;
        }
        // P[t] = (earlier slices) + (earlier lanes) + (this lane's samples <= t) = carry2 + inc - q[t]
        sm.exc2[w][cb.tslot][lane] = inc;
        if (validB && sub == LPR - 1) st_l2(cb.rec + kSlicedNsp + c, sane_bits(inc));
        cb.advance<HD, TD>(row_step, p.rec, last_row);
        if ((cb.n & 15) == 0) loss_acc += (double)loss_f, loss_f = 0.f;  // float32 partial over <= 16 tasks x NV x 4 samples
        did = true;
      }

      // ---------------- C: dL/dh = h P ; ghy accumulators ; <dL/dh, hy_g> -----------------------------------------
      bool readyC = false;
      if (wantC) {
        t2.mask_edge(edge2);
        readyC = __all_sync(0xffffffffu, t2.complete());
      }
      if (readyC) {
        const bool validC = cc.row <= last_row;
        float carry = t2.sum();
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) carry += __shfl_xor_sync(0xffffffffu, carry, o, LPR);
        const float4 s4 = *reinterpret_cast<const float4*>(sm.sv[w][cc.hslot][rw]);
        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
        const float4* hs = my_hd + (cc.hslot * RPW + rw) * S4 + sub * NV;
        const float4* ts = my_td + (cc.tslot * RPW + rw) * S4 + sub * NV;
        const float cst = carry + sm.exc2[w][cc.tslot][lane];
        float2 dg[G];
#pragma unroll
        for (int g = 0; g < G; ++g) dg[g] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const float4 gh = mul4(hs[j], rsub4(cst, ts[j]));
#pragma unroll
          for (int g = 0; g < G; ++g) {
            acc[g][j] = fma4(sv[g], gh, acc[g][j]);
            dot4(dg[g], gh, hyv(g, j));
          }
        }
        float4 gsv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float d = dg[g].x + dg[g].y;
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o, LPR);
          (&gsv.x)[g] = d;
        }
        if (validC && sub == 0) *reinterpret_cast<float4*>(cc.rec + 2 * kSlicedNsp + 4 * c) = gsv;
        // both slots of this task are free: refill them with the tasks that map to them next
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(true, cc.n + HD, cc.hslot);
        issue(false, cc.n + TD, cc.tslot);
        cc.advance<HD, TD>(row_step, p.rec, last_row);
        did = true;
      }
      if (did) {
        idle = 0;
      } else {
        __nanosleep(100);
        if (++idle > (1u << 25)) __trap();  // a lost total or copy must abort the launch, never hang the device
      }
    }
    loss_acc += (double)loss_f;
  }

  // ---- epilogue: ghy slice = sum over (worker warp, row-in-task) of the register accumulators, fixed order ------
  __syncthreads();  // every TMA copy has been consumed; the rings are free
  float4* part = reinterpret_cast<float4*>(dyn_smem);  // [W * RPW][G][S4]
  {
    const int sub = lane & (LPR - 1), rw = lane / LPR;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int j = 0; j < NV; ++j) part[((size_t)(w * RPW + rw) * G + g) * S4 + sub * NV + j] = acc[g][j];
  }
  {
    const double v = warp_sum(loss_acc);
    if (lane == 0) sm.red[w] = v;
  }
  __syncthreads();
  for (int i = tid; i < G * S4; i += T::THREADS) {
    const int g = i / S4, seg = i % S4;
    if (seg >= len4) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < W * RPW; ++q) {
      const float4 a = part[((size_t)q * G + g) * S4 + seg];
      v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
    }
    float4* dst = reinterpret_cast<float4*>(p.ghy) + (int64_t)g * tn4 + seg0 + seg;
    if (p.accumulate) {
      const float4 o = *dst;
      v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
    }
    *dst = v;
  }
  if (tid == 0) {
    double t = 0.0;
    for (int i = 0; i < W; ++i) t += sm.red[i];
    p.part_loss[c] = t;
  }
}

// gs[r,g] = sum_c part_gs[r][c][g] (one warp per row, fixed order) ; loss (+)= sum_c part_loss[c] ; the totals of the
// row go back to "not published" for the next launch
__global__ void td_sliced_finalize_kernel(int64_t rows, int g, int ns, float* __restrict__ rec, float* __restrict__ gs,
                                          const double* __restrict__ part_loss, double* __restrict__ loss, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r < rows) {
    float* row = rec + r * kSlicedRecFloats;
    const float4* pg = reinterpret_cast<const float4*>(row + 2 * kSlicedNsp);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = lane; c < ns; c += 32) {
      const float4 a = pg[c];
      v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
      row[c] = __uint_as_float(kFlag);
      row[kSlicedNsp + c] = __uint_as_float(kFlag);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
      v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
    }
    if (gs != nullptr && lane < g) gs[r * g + lane] = lane == 0 ? v.x : lane == 1 ? v.y : lane == 2 ? v.z : v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss != nullptr) {
    double t = accumulate ? loss[0] : 0.0;
    for (int c = 0; c < ns; ++c) t += part_loss[c];
    loss[0] = t;
  }
}

// ---- shapes ---------------------------------------------------------------------------------------------------
struct ShapeInfo {
  int lpr, nv;
};
constexpr int kNumShapes = 5;
constexpr ShapeInfo kShapes[kNumShapes] = {{8, 1}, {32, 1}, {16, 3}, {16, 5}, {32, 5}};  // S = 32, 128, 192, 320, 640

// Pipeline configurations (DGFDN_TD_SLICED_CFG selects one for tuning; 0 is the default):
//   0: 8 warps (2 per SM sub-partition, <= 255 registers), hy slice in registers when it fits, rings 6 + 4
//   1: 12 warps (3 per sub-partition, <= 168 registers), hy slice in shared memory, rings 4 + 2
template <int G, int LPR, int NV, int CFG>
struct Config;
template <int G, int LPR, int NV>
struct Config<G, LPR, NV, 0> {
  using type = Shape<G, LPR, NV, 8, 6, 4, (G * NV <= 15)>;
};
template <int G, int LPR, int NV>
struct Config<G, LPR, NV, 1> {
  using type = Shape<G, LPR, NV, 12, 4, 2, false>;
};

template <class T>
int launch_sliced(const SlicedParams& p, cudaStream_t st, bool query_only) {
  auto kern = td_sliced_kernel<T>;
  static bool configured[64] = {false};
  int dev = 0;
  DGFDN_CUDA(cudaGetDevice(&dev));
  DGFDN_CHECK(dev >= 0 && dev < 64, "td_edc_sliced: device index out of range");
  const size_t smem = sliced_dyn_smem<T>();
  if (!configured[dev]) {
    DGFDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    DGFDN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, T::THREADS, smem));
    cudaFuncAttributes fa;
    DGFDN_CUDA(cudaFuncGetAttributes(&fa, kern));
    DGFDN_CHECK(nb >= 1, "td_edc_sliced: a CTA of %d threads x %d registers / %zu + %zu B shared memory does not fit an SM",
                T::THREADS, fa.numRegs, smem, fa.sharedSizeBytes);
    configured[dev] = true;
  }
  if (query_only) return 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all NS CTAs co-resident: they wait for one another through L2
  attr[0].val.cooperative = 1;
  cfg.gridDim = dim3((unsigned)p.ns);
  cfg.blockDim = dim3(T::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SlicedParams pp = p;
  DGFDN_CUDA(cudaLaunchKernelEx(&cfg, kern, pp));
  return 0;
}

template <int G, int CFG>
int by_shape(int shape, const SlicedParams& p, cudaStream_t st, bool q) {
  switch (shape) {
    case 0: return launch_sliced<typename Config<G, 8, 1, CFG>::type>(p, st, q);
    case 1: return launch_sliced<typename Config<G, 32, 1, CFG>::type>(p, st, q);
    case 2: return launch_sliced<typename Config<G, 16, 3, CFG>::type>(p, st, q);
    case 3: return launch_sliced<typename Config<G, 16, 5, CFG>::type>(p, st, q);
    case 4: return launch_sliced<typename Config<G, 32, 5, CFG>::type>(p, st, q);
  }
  return 1;
}

int sliced_cfg() {
  if (const char* e = getenv("DGFDN_TD_SLICED_CFG")) return atoi(e) == 1 ? 1 : 0;
  return 0;
}

int dispatch_sliced(int g, int shape, const SlicedParams& p, cudaStream_t st, bool q) {
  const int cfg = sliced_cfg();
  switch (g) {
    case 1: return by_shape<1, 0>(shape, p, st, q);
    case 2: return by_shape<2, 0>(shape, p, st, q);
    case 3: return cfg == 1 ? by_shape<3, 1>(shape, p, st, q) : by_shape<3, 0>(shape, p, st, q);
    case 4: return by_shape<4, 0>(shape, p, st, q);
  }
  return 1;
}

}  // namespace

bool sliced_shape(int g, int64_t tn, SlicedShape* out) {
  if (g < 1 || g > 4 || tn < 4 || tn % 4 != 0) return false;
  const int sms = sm_count();
  const int64_t tn4 = tn / 4;
  for (int i = 0; i < kNumShapes; ++i) {
    const int64_t s4 = (int64_t)kShapes[i].lpr * kShapes[i].nv;
    const int64_t ns = (tn4 + s4 - 1) / s4;
    if (ns <= sms && ns <= kSlicedNsp - 4) {  // the last 128-bit word of a padded row of totals stays 0.0
      if (out) {
        out->shape = i;
        out->ns = (int)ns;
        out->samples_per_slice = (int)(4 * s4);
        out->threads = 256;
      }
      return true;
    }
  }
  return false;
}

size_t sliced_ws_bytes(int64_t rows) { return (size_t)kSlicedHeaderBytes + (size_t)rows * kSlicedRecFloats * sizeof(float); }

namespace {
// T1 / T2 words of every row: "not published" for the slices of a launch, 0.0 past them (read as part of a 128-bit word)
__global__ void td_sliced_ws_init_kernel(float* rec, int64_t rows, int ns) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * kSlicedNsp) return;
  const int64_t r = i / (2 * kSlicedNsp);
  const int j = (int)(i % (2 * kSlicedNsp));
  rec[r * kSlicedRecFloats + j] = (j % kSlicedNsp) < ns ? __uint_as_float(kFlag) : 0.f;
}
}  // namespace

int sliced_ws_init(void* ws, int g, int64_t rows, int64_t tn, cudaStream_t st) {
  DGFDN_CUDA(cudaMemsetAsync(ws, 0xFF, sliced_ws_bytes(rows), st));
  SlicedShape shp;
  if (!sliced_shape(g, tn, &shp)) return 0;
  const int64_t n = rows * 2 * kSlicedNsp;
  td_sliced_ws_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + kSlicedHeaderBytes), rows, shp.ns);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

int sliced_launch(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd, int64_t ldhd,
                  const float* target_db, int64_t ldt, const float* mask, double coef, double* loss_sum, float* gs, float* ghy,
                  int accumulate, void* ws, cudaStream_t st) {
  SlicedShape shp;
  DGFDN_CHECK(sliced_shape(g, tn, &shp), "td_edc_sliced: unsupported shape");
  unsigned char* base = static_cast<unsigned char*>(ws);
  SlicedParams p{};
  p.rows = rows;
  p.tn4 = (int)(tn / 4);
  p.ns = shp.ns;
  p.s = s;
  p.hy = hy;
  p.hd = hd;
  p.ldhd = ldhd;
  p.tdb = target_db;
  p.ldt = ldt;
  p.mask = mask;
  p.coef = coef;
  p.part_loss = reinterpret_cast<double*>(base);
  p.rec = reinterpret_cast<float*>(base + kSlicedHeaderBytes);
  p.ghy = ghy;
  p.accumulate = accumulate;
  if (int rc = dispatch_sliced(g, shp.shape, p, st, false)) return rc;
  const int warps_per_block = 8;
  td_sliced_finalize_kernel<<<(unsigned)((rows + warps_per_block - 1) / warps_per_block), 32 * warps_per_block, 0, st>>>(
      rows, g, shp.ns, p.rec, gs, p.part_loss, loss_sum, accumulate);
  DGFDN_LAUNCH_CHECK();
  return 0;
}

}  // namespace dgfdn
