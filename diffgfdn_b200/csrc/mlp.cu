// K7: position -> gain network of a receiver shard, forward and backward, one launch each.
//
//   enc(p)   = [sin(f_i pi p_xyz), cos(f_i pi p_xyz)]_i                       (reference dnn.py:89-126)
//   a_0      = enc ;  a_l = relu(LN_l(W_l a_{l-1} + b_l)) (+ a_{l-1} in the skip-connection variant, l >= 1)
//   out      = W_out a_last + b_out                                           (reference dnn.py:284-400)
//   gains    = lo + (hi - lo) / (1 + exp(-out))   when final_act = 1          (dnn.py:21-36, gain_filters.py:497-536)
//
// The reference runs this as ~25 torch kernels forward and ~80 backward (cuBLAS GEMMs with M = receivers, N = K =
// 128, LayerNorm gamma/beta reductions, bias reductions, ...); at 12 500 receivers those cost more than the whole
// receiver kernel. Here a CTA owns 96 consecutive receivers and walks the layers with the activations in shared
// memory: FP32 FFMA2 register tiles (6 rows x 8 columns per thread), LayerNorm statistics by half-warp shuffles.
// The forward saves the normalised pre-activations x^_l (and 1/sigma_l), the backward recomputes a_l from them,
// back-propagates, and writes this CTA's partial parameter gradients; a second kernel adds the partials in a fixed
// order (deterministic), accumulating in float64.
#include "common.cuh"

namespace dgfdn {
namespace {

constexpr int kMT = 256;          // threads per CTA: 16 row groups (ty) x 16 column groups (tx)
constexpr int kTR = 6;            // rows per thread
constexpr int kRowsPerCta = 96;   // 16 * kTR
constexpr int kLd = 132;          // shared-memory pitch of an activation row (floats): 128 + 4, keeps float4 alignment
constexpr int kMaxLayers = 8;     // LayerNorm layers (1 + hidden)
constexpr int kMaxOut = 32;
constexpr float kLnEps = 1e-5f;   // torch.nn.LayerNorm default

struct MlpParams {
  int64_t rows;
  int in_dim, in_pad;   // 6 F and its padding to a multiple of 64
  int nfeat;            // F
  int nl;               // LayerNorm layers
  int out_dim;
  int residual;
  int final_act;
  int pos_is_double;
  float lo, hi;
  const void* pos;      // [rows, 3] float or double
  const void* freq;     // [F] f_i pi in the position dtype
  const float* w[kMaxLayers];
  const float* b[kMaxLayers];
  const float* gam[kMaxLayers];
  const float* bet[kMaxLayers];
  const float* wout;
  const float* bout;
  float* out;           // [rows, out_dim]  (post final activation)
  float* xhat;          // [nl, rows, H]
  float* rstd;          // [nl, rows]
  float* asave;         // [nl, rows, H] post-activation (residual variant only) or null
  // backward
  const float* gout;    // [rows, out_dim]
  float* part;          // [grid, nparams] partial gradients of this CTA
  int64_t nparams;
  int64_t part_stride;  // nparams rounded up to a multiple of 4 floats (128-bit stores into a CTA's slice)
};

__device__ __forceinline__ float2 ffma2(float a, float2 w, float2 c) { return __ffma2_rn(make_float2(a, a), w, c); }

// acc[i][j] (+)= sum_k A[r_i][k] * W[k][c_j]: A row-major in shared memory with pitch kLd (r_i = ty*6 + i), W[k][n] in
// shared memory with pitch ldw; this thread's columns are tx*4 + 64*j' + {0..3}.
template <int NC4>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ as, const float* __restrict__ ws, int ldw, int kdim,
                                          int ty, int tx, float4 (&acc)[kTR][NC4]) {
  const float* arow = as + (ty * kTR) * kLd;
  for (int k4 = 0; k4 < kdim; k4 += 4) {
    float4 a[kTR];
#pragma unroll
    for (int i = 0; i < kTR; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * kLd + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float4 w[NC4];
#pragma unroll
      for (int j = 0; j < NC4; ++j) w[j] = *reinterpret_cast<const float4*>(ws + (k4 + kk) * ldw + tx * 4 + 64 * j);
#pragma unroll
      for (int i = 0; i < kTR; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
        for (int j = 0; j < NC4; ++j) {
          const float2 lo2 = ffma2(av, make_float2(w[j].x, w[j].y), make_float2(acc[i][j].x, acc[i][j].y));
          const float2 hi2 = ffma2(av, make_float2(w[j].z, w[j].w), make_float2(acc[i][j].z, acc[i][j].w));
          acc[i][j] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
        }
      }
    }
  }
}

__device__ __forceinline__ float half_warp_sum(float v) {  // the 16 lanes that share a row group
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// enc rows of this CTA into as[r][0..in_pad): sin / cos in the position dtype, cast to float (dnn.py:120-126)
template <typename T>
__device__ __forceinline__ void encode_tile(const MlpParams& p, int64_t row0, int nvalid, float* as) {
  const T* pos = static_cast<const T*>(p.pos);
  const T* freq = static_cast<const T*>(p.freq);
  const int per_row = p.nfeat * 3;
  for (int idx = threadIdx.x; idx < kRowsPerCta * per_row; idx += kMT) {
    const int r = idx / per_row, e = idx % per_row;
    const int f = e / 3, ax = e % 3;
    float s = 0.f, c = 0.f;
    if (r < nvalid) {
      const T arg = freq[f] * pos[(row0 + r) * 3 + ax];
      s = (float)sin(arg);
      c = (float)cos(arg);
    }
    as[r * kLd + f * 6 + ax] = s;
    as[r * kLd + f * 6 + 3 + ax] = c;
  }
  for (int idx = threadIdx.x; idx < kRowsPerCta * (p.in_pad - p.in_dim); idx += kMT) {
    const int r = idx / (p.in_pad - p.in_dim), e = idx % (p.in_pad - p.in_dim);
    as[r * kLd + p.in_dim + e] = 0.f;
  }
}

// ws[k][n] = W[n][k] for k < kdim_pad (zero beyond kdim), n < H: lanes run over n so the shared-memory stores are
// conflict free; the strided global reads hit L2 (a layer is 64 KB).
template <int H>
__device__ __forceinline__ void stage_w_transposed(const float* __restrict__ w, int kdim, int kdim_pad, float* ws) {
  for (int idx = threadIdx.x; idx < kdim_pad * H; idx += kMT) {
    const int k = idx / H, n = idx % H;
    ws[k * H + n] = k < kdim ? __ldg(w + (size_t)n * kdim + k) : 0.f;
  }
}

template <int H>
__global__ void __launch_bounds__(kMT, 1) mlp_fwd_kernel(MlpParams p) {
  constexpr int NC4 = H / 64;
  extern __shared__ __align__(16) float sm[];
  float* buf0 = sm;                          // [96][kLd]
  float* buf1 = buf0 + kRowsPerCta * kLd;    // [96][kLd]
  float* ws = buf1 + kRowsPerCta * kLd;      // [128][H]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta;
  const int nvalid = (int)min((int64_t)kRowsPerCta, p.rows - row0);

  if (p.pos_is_double) encode_tile<double>(p, row0, nvalid, buf0);
  else encode_tile<float>(p, row0, nvalid, buf0);
  float* a_in = buf0;
  float* a_out = buf1;
  for (int l = 0; l < p.nl; ++l) {
    const int kdim = l == 0 ? p.in_dim : H, kpad = l == 0 ? p.in_pad : H;
    __syncthreads();  // a_in complete; ws free
    stage_w_transposed<H>(p.w[l], kdim, kpad, ws);
    __syncthreads();
    float4 acc[kTR][NC4];
#pragma unroll
    for (int j = 0; j < NC4; ++j) {
      const float4 bias = __ldg(reinterpret_cast<const float4*>(p.b[l] + tx * 4 + 64 * j));
#pragma unroll
      for (int i = 0; i < kTR; ++i) acc[i][j] = bias;
    }
    gemm_rows<NC4>(a_in, ws, H, kpad, ty, tx, acc);
    float4 gam[NC4], bet[NC4];
#pragma unroll
    for (int j = 0; j < NC4; ++j) {
      gam[j] = __ldg(reinterpret_cast<const float4*>(p.gam[l] + tx * 4 + 64 * j));
      bet[j] = __ldg(reinterpret_cast<const float4*>(p.bet[l] + tx * 4 + 64 * j));
    }
#pragma unroll
    for (int i = 0; i < kTR; ++i) {
      const int r = ty * kTR + i;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NC4; ++j) s += (acc[i][j].x + acc[i][j].y) + (acc[i][j].z + acc[i][j].w);
      const float mean = half_warp_sum(s) * (1.f / H);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NC4; ++j) {
        acc[i][j].x -= mean, acc[i][j].y -= mean, acc[i][j].z -= mean, acc[i][j].w -= mean;
        q += (acc[i][j].x * acc[i][j].x + acc[i][j].y * acc[i][j].y) + (acc[i][j].z * acc[i][j].z + acc[i][j].w * acc[i][j].w);
      }
      const float rstd = rsqrtf(half_warp_sum(q) * (1.f / H) + kLnEps);
      const bool live = r < nvalid;
      if (live && tx == 0) p.rstd[(int64_t)l * p.rows + row0 + r] = rstd;
#pragma unroll
      for (int j = 0; j < NC4; ++j) {
        const int c = tx * 4 + 64 * j;
        float4 xh = make_float4(acc[i][j].x * rstd, acc[i][j].y * rstd, acc[i][j].z * rstd, acc[i][j].w * rstd);
        if (live) *reinterpret_cast<float4*>(p.xhat + ((int64_t)l * p.rows + row0 + r) * H + c) = xh;
        float4 y;
        y.x = fmaxf(fmaf(xh.x, gam[j].x, bet[j].x), 0.f);
        y.y = fmaxf(fmaf(xh.y, gam[j].y, bet[j].y), 0.f);
        y.z = fmaxf(fmaf(xh.z, gam[j].z, bet[j].z), 0.f);
        y.w = fmaxf(fmaf(xh.w, gam[j].w, bet[j].w), 0.f);
        if (p.residual && l > 0) {
          const float4 prev = *reinterpret_cast<const float4*>(a_in + r * kLd + c);
          y.x += prev.x, y.y += prev.y, y.z += prev.z, y.w += prev.w;
        }
        *reinterpret_cast<float4*>(a_out + r * kLd + c) = y;
        if (p.asave != nullptr && live) *reinterpret_cast<float4*>(p.asave + ((int64_t)l * p.rows + row0 + r) * H + c) = y;
      }
    }
    float* t = a_in;
    a_in = a_out;
    a_out = t;
  }
  __syncthreads();
  // output layer: out[r][j] = b_out[j] + <a[r], W_out[j]>; W_out staged as is ([out_dim][H])
  for (int idx = tid; idx < p.out_dim * H; idx += kMT) ws[idx] = __ldg(p.wout + idx);
  __syncthreads();
  for (int idx = tid; idx < kRowsPerCta * p.out_dim; idx += kMT) {
    const int r = idx / p.out_dim, j = idx % p.out_dim;
    if (r >= nvalid) continue;
    const float4* a4 = reinterpret_cast<const float4*>(a_in + r * kLd);
    const float4* w4 = reinterpret_cast<const float4*>(ws + j * H);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < H / 4; ++k) {
      const float4 a = a4[k], w = w4[k];
      s0 = fmaf(a.x, w.x, s0), s1 = fmaf(a.y, w.y, s1), s0 = fmaf(a.z, w.z, s0), s1 = fmaf(a.w, w.w, s1);
    }
    float v = (s0 + s1) + __ldg(p.bout + j);
    if (p.final_act == 1) v = p.lo + (p.hi - p.lo) * (1.f / (1.f + expf(-v)));
    p.out[(row0 + r) * p.out_dim + j] = v;
  }
}

// Flat layout of the parameter gradients: per layer W_l [H, in_l], b_l [H], gamma_l [H], beta_l [H]; W_out, b_out.
__device__ __host__ inline int64_t layer_offset(int l, int in_dim, int h) {
  if (l == 0) return 0;
  return (int64_t)h * in_dim + 3 * h + (int64_t)(l - 1) * ((int64_t)h * h + 3 * h);
}

template <int H>
__global__ void __launch_bounds__(kMT, 1) mlp_bwd_kernel(MlpParams p) {
  constexpr int NC4 = H / 64;
  constexpr int NN = H / 16;  // rows of dW per thread (ty*4 + 64*j' + {0..3})
  extern __shared__ __align__(16) float sm[];
  float* gz = sm;                           // [96][kLd] grad wrt the linear output of the current layer
  float* ap = gz + kRowsPerCta * kLd;       // [96][kLd] input activations of the current layer
  float* ws = ap + kRowsPerCta * kLd;       // [H][H] W_l as stored ([n][k])
  float* red = ws + H * H;                  // [3][16][H] column partial sums of the row groups
  float* gy_s = red + 3 * 16 * H;           // [96][kMaxOut + 1]
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta;
  const int nvalid = (int)min((int64_t)kRowsPerCta, p.rows - row0);
  float* part = p.part + (int64_t)blockIdx.x * p.part_stride;
  const int64_t off_out = layer_offset(p.nl, p.in_dim, H);

  // activation entering layer l (l = nl: the output layer) into ap
  auto load_act = [&](int l) {
    if (l == 0) {
      if (p.pos_is_double) encode_tile<double>(p, row0, nvalid, ap);
      else encode_tile<float>(p, row0, nvalid, ap);
      return;
    }
    const int src = l - 1;
    for (int idx = tid; idx < kRowsPerCta * (H / 4); idx += kMT) {
      const int r = idx / (H / 4), c = (idx % (H / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nvalid) {
        if (p.asave != nullptr) {
          v = __ldg(reinterpret_cast<const float4*>(p.asave + ((int64_t)src * p.rows + row0 + r) * H + c));
        } else {
          const float4 xh = __ldg(reinterpret_cast<const float4*>(p.xhat + ((int64_t)src * p.rows + row0 + r) * H + c));
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gam[src] + c));
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bet[src] + c));
          v.x = fmaxf(fmaf(xh.x, g.x, b.x), 0.f);
          v.y = fmaxf(fmaf(xh.y, g.y, b.y), 0.f);
          v.z = fmaxf(fmaf(xh.z, g.z, b.z), 0.f);
          v.w = fmaxf(fmaf(xh.w, g.w, b.w), 0.f);
        }
      }
      *reinterpret_cast<float4*>(ap + r * kLd + c) = v;
    }
  };

  // ---- output layer -------------------------------------------------------------------------------------
  load_act(p.nl);
  for (int idx = tid; idx < kRowsPerCta * p.out_dim; idx += kMT) {
    const int r = idx / p.out_dim, j = idx % p.out_dim;
    float g = 0.f;
    if (r < nvalid) {
      g = __ldg(p.gout + (row0 + r) * p.out_dim + j);
      if (p.final_act == 1) {
        const float sg = (__ldg(p.out + (row0 + r) * p.out_dim + j) - p.lo) / (p.hi - p.lo);
        g *= (p.hi - p.lo) * sg * (1.f - sg);
      }
    }
    gy_s[r * (kMaxOut + 1) + j] = g;
  }
  for (int idx = tid; idx < p.out_dim * H; idx += kMT) ws[idx] = __ldg(p.wout + idx);
  __syncthreads();
  for (int idx = tid; idx < p.out_dim * H; idx += kMT) {  // dW_out[j][k] = sum_r gy[r][j] a[r][k]
    const int j = idx / H, k = idx % H;
    float s = 0.f;
    for (int r = 0; r < kRowsPerCta; ++r) s = fmaf(gy_s[r * (kMaxOut + 1) + j], ap[r * kLd + k], s);
    part[off_out + idx] = s;
  }
  if (tid < p.out_dim) {
    float s = 0.f;
    for (int r = 0; r < kRowsPerCta; ++r) s += gy_s[r * (kMaxOut + 1) + tid];
    part[off_out + (int64_t)p.out_dim * H + tid] = s;
  }
  float4 ga[kTR][NC4];  // grad wrt a_l for this thread's tile
#pragma unroll
  for (int i = 0; i < kTR; ++i)
#pragma unroll
    for (int j = 0; j < NC4; ++j) ga[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int o = 0; o < p.out_dim; ++o) {
    float4 w[NC4];
#pragma unroll
    for (int j = 0; j < NC4; ++j) w[j] = *reinterpret_cast<const float4*>(ws + o * H + tx * 4 + 64 * j);
#pragma unroll
    for (int i = 0; i < kTR; ++i) {
      const float g = gy_s[(ty * kTR + i) * (kMaxOut + 1) + o];
#pragma unroll
      for (int j = 0; j < NC4; ++j) {
        ga[i][j].x = fmaf(g, w[j].x, ga[i][j].x), ga[i][j].y = fmaf(g, w[j].y, ga[i][j].y);
        ga[i][j].z = fmaf(g, w[j].z, ga[i][j].z), ga[i][j].w = fmaf(g, w[j].w, ga[i][j].w);
      }
    }
  }

  // ---- LayerNorm layers, last to first --------------------------------------------------------------------
  for (int l = p.nl - 1; l >= 0; --l) {
    const int kdim = l == 0 ? p.in_dim : H, kpad = l == 0 ? p.in_pad : H;
    const int64_t off = layer_offset(l, p.in_dim, H);
    const int64_t off_b = off + (int64_t)H * kdim, off_g = off_b + H, off_be = off_g + H;
    float4 gam[NC4], bet[NC4], cg[NC4], cb[NC4], cz[NC4];
#pragma unroll
    for (int j = 0; j < NC4; ++j) {
      gam[j] = __ldg(reinterpret_cast<const float4*>(p.gam[l] + tx * 4 + 64 * j));
      bet[j] = __ldg(reinterpret_cast<const float4*>(p.bet[l] + tx * 4 + 64 * j));
      cg[j] = cb[j] = cz[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();  // previous layer's GEMMs are done with gz / ap / ws
#pragma unroll
    for (int i = 0; i < kTR; ++i) {
      const int r = ty * kTR + i;
      const bool live = r < nvalid;
      const float rstd = live ? __ldg(p.rstd + (int64_t)l * p.rows + row0 + r) : 0.f;
      float4 xh[NC4], gx[NC4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NC4; ++j) {
        const int c = tx * 4 + 64 * j;
        xh[j] = live ? __ldg(reinterpret_cast<const float4*>(p.xhat + ((int64_t)l * p.rows + row0 + r) * H + c))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
#define DGFDN_LN_ONE(C)                                                                       \
  {                                                                                           \
    const float gyv = fmaf(xh[j].C, gam[j].C, bet[j].C) > 0.f ? ga[i][j].C : 0.f;             \
    cg[j].C = fmaf(gyv, xh[j].C, cg[j].C);                                                    \
    cb[j].C += gyv;                                                                           \
    gx[j].C = gyv * gam[j].C;                                                                 \
    s1 += gx[j].C;                                                                            \
    s2 = fmaf(gx[j].C, xh[j].C, s2);                                                          \
  }
        DGFDN_LN_ONE(x) DGFDN_LN_ONE(y) DGFDN_LN_ONE(z) DGFDN_LN_ONE(w)
#undef DGFDN_LN_ONE
      }
      const float m1 = half_warp_sum(s1) * (1.f / H), m2 = half_warp_sum(s2) * (1.f / H);
#pragma unroll
      for (int j = 0; j < NC4; ++j) {
        float4 z;
        z.x = rstd * (gx[j].x - m1 - xh[j].x * m2);
        z.y = rstd * (gx[j].y - m1 - xh[j].y * m2);
        z.z = rstd * (gx[j].z - m1 - xh[j].z * m2);
        z.w = rstd * (gx[j].w - m1 - xh[j].w * m2);
        cz[j].x += z.x, cz[j].y += z.y, cz[j].z += z.z, cz[j].w += z.w;
        *reinterpret_cast<float4*>(gz + r * kLd + tx * 4 + 64 * j) = z;
      }
    }
#pragma unroll
    for (int j = 0; j < NC4; ++j) {
      const int c = tx * 4 + 64 * j;
      *reinterpret_cast<float4*>(red + (0 * 16 + ty) * H + c) = cg[j];
      *reinterpret_cast<float4*>(red + (1 * 16 + ty) * H + c) = cb[j];
      *reinterpret_cast<float4*>(red + (2 * 16 + ty) * H + c) = cz[j];
    }
    load_act(l);
    for (int idx = tid; idx < (H * kpad) / 4; idx += kMT) {  // W_l as stored, [n][k] with pitch kpad
      const int n = (idx * 4) / kpad, k = (idx * 4) % kpad;
      float4 v;
      if (kdim == kpad) {
        v = __ldg(reinterpret_cast<const float4*>(p.w[l] + (size_t)n * kdim + k));
      } else {
        v.x = k + 0 < kdim ? __ldg(p.w[l] + (size_t)n * kdim + k + 0) : 0.f;
        v.y = k + 1 < kdim ? __ldg(p.w[l] + (size_t)n * kdim + k + 1) : 0.f;
        v.z = k + 2 < kdim ? __ldg(p.w[l] + (size_t)n * kdim + k + 2) : 0.f;
        v.w = k + 3 < kdim ? __ldg(p.w[l] + (size_t)n * kdim + k + 3) : 0.f;
      }
      *reinterpret_cast<float4*>(ws + n * kpad + k) = v;
    }
    __syncthreads();
    for (int idx = tid; idx < 3 * H; idx += kMT) {  // gamma, beta, bias partials of this CTA (fixed order)
      const int which = idx / H, c = idx % H;
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) s += red[(which * 16 + g) * H + c];
      part[(which == 0 ? off_g : which == 1 ? off_be : off_b) + c] = s;
    }
    // dW_l[n][k] = sum_r gz[r][n] ap[r][k]: this thread owns n = ty*4 + 64 jn + {0..3}, k = tx*4 + 64 jk + {0..3}
    {
      const int nk4 = kpad / 64;  // 1 or 2 column groups
      float4 acc[NN][NC4];
#pragma unroll
      for (int a = 0; a < NN; ++a)
#pragma unroll
        for (int j = 0; j < NC4; ++j) acc[a][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < kRowsPerCta; ++r) {
        float4 g4[NC4], a4[NC4];
#pragma unroll
        for (int j = 0; j < NC4; ++j) {
          g4[j] = *reinterpret_cast<const float4*>(gz + r * kLd + ty * 4 + 64 * j);
          a4[j] = j < nk4 ? *reinterpret_cast<const float4*>(ap + r * kLd + tx * 4 + 64 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int jn = 0; jn < NC4; ++jn) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float gv = e == 0 ? g4[jn].x : e == 1 ? g4[jn].y : e == 2 ? g4[jn].z : g4[jn].w;
#pragma unroll
            for (int jk = 0; jk < NC4; ++jk) {
              float4& c = acc[jn * 4 + e][jk];
              const float2 lo2 = ffma2(gv, make_float2(a4[jk].x, a4[jk].y), make_float2(c.x, c.y));
              const float2 hi2 = ffma2(gv, make_float2(a4[jk].z, a4[jk].w), make_float2(c.z, c.w));
              c = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
            }
          }
        }
      }
#pragma unroll
      for (int jn = 0; jn < NC4; ++jn)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = ty * 4 + 64 * jn + e;
#pragma unroll
          for (int jk = 0; jk < NC4; ++jk) {
            const int k = tx * 4 + 64 * jk;
            if (jk >= nk4) continue;
            const float4 v = acc[jn * 4 + e][jk];
            float* dst = part + off + (int64_t)n * kdim + k;
            if (kdim == kpad) {
              *reinterpret_cast<float4*>(dst) = v;
            } else {
              if (k + 0 < kdim) dst[0] = v.x;
              if (k + 1 < kdim) dst[1] = v.y;
              if (k + 2 < kdim) dst[2] = v.z;
              if (k + 3 < kdim) dst[3] = v.w;
            }
          }
        }
    }
    if (l > 0) {  // grad wrt a_{l-1}: gz W_l (+ the skip path)
      float4 nga[kTR][NC4];
#pragma unroll
      for (int i = 0; i < kTR; ++i)
#pragma unroll
        for (int j = 0; j < NC4; ++j) nga[i][j] = p.residual ? ga[i][j] : make_float4(0.f, 0.f, 0.f, 0.f);
      gemm_rows<NC4>(gz, ws, H, H, ty, tx, nga);
#pragma unroll
      for (int i = 0; i < kTR; ++i)
#pragma unroll
        for (int j = 0; j < NC4; ++j) ga[i][j] = nga[i][j];
    }
  }
}

// grad[i] = sum_c part[c][i] in CTA order, float64 accumulation
__global__ void mlp_reduce_kernel(const float* __restrict__ part, int ncta, int64_t stride, int64_t nparams,
                                  float* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nparams) return;
  double s = 0.0;
  for (int c = 0; c < ncta; ++c) s += (double)part[(int64_t)c * stride + i];
  grad[i] = (float)s;
}

int64_t nparams_of(int in_dim, int h, int nl, int out_dim) { return layer_offset(nl, in_dim, h) + (int64_t)out_dim * h + out_dim; }

bool shape_ok(int in_dim, int nfeat, int h, int nl, int out_dim) {
  return (h == 64 || h == 128) && in_dim == 6 * nfeat && in_dim >= 6 && in_dim <= h && in_dim % 4 == 0 && nl >= 1 &&
         nl <= kMaxLayers && out_dim >= 1 && out_dim <= kMaxOut;
}

size_t fwd_smem(int h) { return (size_t)(2 * kRowsPerCta * kLd + 128 * h) * sizeof(float); }
size_t bwd_smem(int h) {
  return (size_t)(2 * kRowsPerCta * kLd + h * h + 3 * 16 * h + kRowsPerCta * (kMaxOut + 1)) * sizeof(float);
}

int fill_params(MlpParams* p, int64_t rows, int in_dim, int nfeat, int h, int nl, int out_dim, int residual, int final_act,
                float lo, float hi, int pos_is_double, const void* pos, const void* freq, const float* const* params_host) {
  DGFDN_CHECK(shape_ok(in_dim, nfeat, h, nl, out_dim),
              "mlp: unsupported shape (neurons 64|128, in = 6F <= neurons, <= %d LayerNorm layers, out <= %d)", kMaxLayers,
              kMaxOut);
  DGFDN_CHECK(rows >= 0 && pos && freq && params_host, "mlp: bad arguments");
  DGFDN_CHECK(final_act == 0 || (final_act == 1 && hi != lo), "mlp: bad final activation");
  p->rows = rows;
  p->in_dim = in_dim;
  p->in_pad = (in_dim + 63) / 64 * 64;
  p->nfeat = nfeat;
  p->nl = nl;
  p->out_dim = out_dim;
  p->residual = residual;
  p->final_act = final_act;
  p->pos_is_double = pos_is_double;
  p->lo = lo;
  p->hi = hi;
  p->pos = pos;
  p->freq = freq;
  for (int l = 0; l < nl; ++l) {
    p->w[l] = params_host[4 * l + 0];
    p->b[l] = params_host[4 * l + 1];
    p->gam[l] = params_host[4 * l + 2];
    p->bet[l] = params_host[4 * l + 3];
    DGFDN_CHECK(p->w[l] && p->b[l] && p->gam[l] && p->bet[l], "mlp: null parameter pointer");
    DGFDN_CHECK(((uintptr_t)p->w[l] | (uintptr_t)p->b[l] | (uintptr_t)p->gam[l] | (uintptr_t)p->bet[l]) % 16 == 0,
                "mlp: parameters must be 16-byte aligned");
  }
  p->wout = params_host[4 * nl];
  p->bout = params_host[4 * nl + 1];
  DGFDN_CHECK(p->wout && p->bout, "mlp: null parameter pointer");
  return 0;
}

}  // namespace
}  // namespace dgfdn

using namespace dgfdn;

extern "C" int dgfdn_mlp_supported(int in_dim, int nfeat, int neurons, int nl, int out_dim) {
  return shape_ok(in_dim, nfeat, neurons, nl, out_dim) ? 1 : 0;
}

extern "C" int64_t dgfdn_mlp_num_params(int in_dim, int neurons, int nl, int out_dim) {
  return nparams_of(in_dim, neurons, nl, out_dim);
}

extern "C" int64_t dgfdn_mlp_bwd_ws_bytes(int64_t rows, int in_dim, int neurons, int nl, int out_dim) {
  const int64_t ncta = (rows + kRowsPerCta - 1) / kRowsPerCta;
  return ncta * ((nparams_of(in_dim, neurons, nl, out_dim) + 3) / 4 * 4) * (int64_t)sizeof(float);
}

extern "C" int dgfdn_mlp_fwd(int64_t rows, int in_dim, int nfeat, int neurons, int nl, int out_dim, int residual,
                             int final_act, float lo, float hi, int pos_is_double, const void* pos, const void* freq,
                             const float* const* params_host, float* out, float* xhat, float* rstd, float* asave,
                             void* stream) {
  MlpParams p{};
  if (fill_params(&p, rows, in_dim, nfeat, neurons, nl, out_dim, residual, final_act, lo, hi, pos_is_double, pos, freq,
                  params_host))
    return 1;
  DGFDN_CHECK(out && xhat && rstd && (!residual || asave), "mlp_fwd: missing output buffers");
  if (rows == 0) return 0;
  p.out = out;
  p.xhat = xhat;
  p.rstd = rstd;
  p.asave = residual ? asave : nullptr;
  const unsigned grid = (unsigned)((rows + kRowsPerCta - 1) / kRowsPerCta);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = fwd_smem(neurons);
  if (neurons == 128) {
    DGFDN_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_fwd_kernel<128><<<grid, kMT, smem, st>>>(p);
  } else {
    DGFDN_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_fwd_kernel<64><<<grid, kMT, smem, st>>>(p);
  }
  DGFDN_LAUNCH_CHECK();
  return 0;
}

extern "C" int dgfdn_mlp_bwd(int64_t rows, int in_dim, int nfeat, int neurons, int nl, int out_dim, int residual,
                             int final_act, float lo, float hi, int pos_is_double, const void* pos, const void* freq,
                             const float* const* params_host, const float* out, const float* xhat, const float* rstd,
                             const float* asave, const float* gout, float* grad, void* ws, void* stream) {
  MlpParams p{};
  if (fill_params(&p, rows, in_dim, nfeat, neurons, nl, out_dim, residual, final_act, lo, hi, pos_is_double, pos, freq,
                  params_host))
    return 1;
  DGFDN_CHECK(out && xhat && rstd && gout && grad && ws && (!residual || asave), "mlp_bwd: missing buffers");
  p.nparams = nparams_of(in_dim, neurons, nl, out_dim);
  p.part_stride = (p.nparams + 3) / 4 * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    DGFDN_CUDA(cudaMemsetAsync(grad, 0, (size_t)p.nparams * sizeof(float), st));
    return 0;
  }
  p.out = const_cast<float*>(out);
  p.xhat = const_cast<float*>(xhat);
  p.rstd = const_cast<float*>(rstd);
  p.asave = residual ? const_cast<float*>(asave) : nullptr;
  p.gout = gout;
  p.part = static_cast<float*>(ws);
  const unsigned grid = (unsigned)((rows + kRowsPerCta - 1) / kRowsPerCta);
  const size_t smem = bwd_smem(neurons);
  if (neurons == 128) {
    DGFDN_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_bwd_kernel<128><<<grid, kMT, smem, st>>>(p);
  } else {
    DGFDN_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_bwd_kernel<64><<<grid, kMT, smem, st>>>(p);
  }
  DGFDN_LAUNCH_CHECK();
  mlp_reduce_kernel<<<(unsigned)((p.nparams + 255) / 256), 256, 0, st>>>(p.part, (int)grid, p.part_stride, p.nparams, grad);
  DGFDN_LAUNCH_CHECK();
  return 0;
}
