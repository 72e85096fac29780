/*
 * diffgfdn_b200 -- C ABI of the B200 (sm_100a) kernels for the DiffGFDN hot path.
 *
 * The reference (orchidas/DiffGFDN) is pure PyTorch: it has no FFI of its own. The entry points below
 * are what a binding for its hot path would call; each one names the reference code it replaces
 * (paths relative to /root/reference/src). INTEGRATION.md shows the ctypes stub on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless its name ends in _host;
 *   - complex64  = interleaved (re, im) float  pairs (torch.complex64 layout), "c64" below;
 *   - complex128 = interleaved (re, im) double pairs (torch.complex128 layout), "c128";
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     except the *_host entry points and plan creation;
 *   - every function returns 0 on success, non-zero on error; dgfdn_last_error() returns the message
 *     of the last failure on the calling thread. There is no CPU fallback anywhere.
 *   - complex gradients follow the torch convention: g = dL/dRe + i dL/dIm.
 */
#ifndef DIFFGFDN_B200_H
#define DIFFGFDN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DGFDN_API __attribute__((visibility("default")))
#else
#define DGFDN_API
#endif

#define DGFDN_MAX_LINES 32  /* N: delay lines handled by one warp */
#define DGFDN_MAX_GROUPS 8  /* G */

DGFDN_API const char* dgfdn_last_error(void);
DGFDN_API int dgfdn_version(void);
/* number of SMs of the current device (grid sizing on the host side) */
DGFDN_API int dgfdn_sm_count(void);
/* Strided host -> device copy on `stream`: the first width_bytes of each of `rows` rows of a (pinned) host array
 * (cudaMemcpy2DAsync). The end-to-end path uses it to ship only bins 0..K/2 of the reference-layout (B, K) arrays
 * produced by the data loader (dataloader.py:250,320-325, 674-704) -- the bins irfft(X, n=K) reads (quirk Q3). */
DGFDN_API int dgfdn_copy_rows_h2d(void* dst, int64_t dst_pitch_bytes, const void* src_host, int64_t src_pitch_bytes,
                        int64_t width_bytes, int64_t rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Orthogonal parametrisation of the mixing matrices.  Replaces Skew + MatrixExponential (torch.matrix_exp) of
 * feedback_loop.py:16-36, 270, 401-403, which synchronises with the host twice per step:
 *   U_g = expm(triu(M_g,1) - triu(M_g,1)^T),   m, u [G,L,L] float32, 2 L <= 32, one CTA per matrix, float64 inside.
 * bwd: gm = d<gu, U>/dM  (adjoint of the Frechet derivative through the 2L x 2L block identity). */
DGFDN_API int dgfdn_skew_expm_fwd(int g, int l, const float* m, float* u, void* stream);
DGFDN_API int dgfdn_skew_expm_bwd(int g, int l, const float* m, const float* gu, float* gm, void* stream);

/* Coupled feedback matrix and its adjoint in one launch each (feedback_loop.py:39-87, 393-412, 424-455):
 *   Phi = ND_Unitary(clamp(alpha, -pi, pi)) [G,G];   A[iL+a, jL+b] = Phi[i,j] (U_i U_j)[a,b]   (diagonal blocks U_i^2)
 * u [G,L,L] float32 (the orthogonal mixing matrices), alpha [G(G-1)/2] float32 (NULL when G = 1), a [N,N] and phi [G,G]
 * float64 out. bwd: ga [N,N] float64 -> gu [G,L,L], galpha [G(G-1)/2] float32 (either may be NULL). G <= 8. */
DGFDN_API int dgfdn_coupled_feedback_fwd(int g, int l, const float* u, const float* alpha, double* a, double* phi, void* stream);
DGFDN_API int dgfdn_coupled_feedback_bwd(int g, int l, const float* u, const float* alpha, const double* ga, float* gu,
                               float* galpha, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1: per-bin build + solve.   Replaces FeedbackLoop.forward (diff_gfdn/feedback_loop.py:326-391),
 * the two einsums of DiffGFDNVarReceiverPos.forward (model.py:615-619) up to the receiver gains,
 * the state of DiffDirectionalFDNVarReceiverPos.forward (model.py:1083, transpose_a = 1) and
 * DiffGFDN.sub_fdn_output (model.py:209-252, with A = blockdiag(M_raw), gamma = NULL).
 *
 *   M_k = diag(z_k^{m_i} / gamma_i) - A            (transpose_a: - A^T)
 *   x_k = M_k^{-1} b                               Gauss-Jordan with partial pivoting, float64
 *   y[k,g] = sum_{n in group g} c_n x_k[n]
 *
 * z       [K]    c128 sample points (dataloader.py:552-566)
 * delays  [N]    int32, 0 <= m_i < 2^30 (z^m is formed by binary powering)
 * a       [N,N]  float32 row-major
 * gamma   [N]    float32 or NULL (= 1);  gamma_z [N,K] c64 or NULL (per-bin filter response, overrides gamma)
 * b, c    [N]    float32
 * x       [K,N]  c64 out (may be NULL)     y [K,G] c64 out (may be NULL)
 * factors NULL, or dgfdn_solve_factors_bytes(n, k) bytes that receive the elimination itself (the multipliers of every
 *         Gauss-Jordan step, the pivots, z^m): handed to dgfdn_solve_bwd, the adjoint solve then replays them as N
 *         rank-one updates instead of factorising M^H again (the reference's autograd re-solves with the saved LU too).
 */
DGFDN_API int64_t dgfdn_solve_factors_bytes(int n, int64_t k);
DGFDN_API int dgfdn_solve_fwd(int n, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                    int transpose_a, const float* gamma, const void* gamma_z, const float* b, const float* c,
                    void* x, void* y, void* factors, void* stream);

/* Adjoint of dgfdn_solve_fwd (replaces autograd through torch.linalg.inv + einsum, trainer.py:473-474).
 *   lambda_k = M_k^{-H} (c o gy[k, g(.)] + gx[k, .])
 *   ga  = Re sum_k lambda_k x_k^H (transposed back if transpose_a),  gb = Re sum_k lambda_k,
 *   gc_n = Re sum_k conj(x_k[n]) gy[k,g(n)],   ginvgamma_i = -Re sum_k conj(z_k^{m_i}) lambda_k[i] conj(x_k[i])
 * x [K,N] c64 is the state saved by the forward call; gy [K,G] c64 and gx [K,N] c64 may be NULL (not both).
 * Outputs are float64: ga [N,N], gb [N], gc [N], ginvgamma [N] (gradient w.r.t. 1/gamma_i; ignored if NULL).
 * ws: scratch of dgfdn_solve_bwd_ws_bytes(n) bytes. Reduction order is fixed (deterministic).
 * factors: the buffer filled by the forward call with the same (z, a, gamma, ...) or NULL (fresh elimination of M^H).
 */
DGFDN_API int64_t dgfdn_solve_bwd_ws_bytes(int n);
DGFDN_API int dgfdn_solve_bwd(int n, int g, int64_t k, const void* z, const int32_t* delays, const float* a,
                    int transpose_a, const float* gamma, const void* gamma_z, const float* c, const void* x,
                    const void* gy, const void* gx, double* ga, double* gb, double* gc, double* ginvgamma,
                    void* ws, const void* factors, void* stream);

/* K1 with FIR coupling (coupling_matrix_type: filter_matrix, feedback_loop.py:90-143, 362-373, 447-453): the feedback matrix is
 * a polynomial A(z) = sum_p A_p z^-p, so every bin has its own complex M_k = diag(z_k^m / gamma) - sum_p A_p z_k^-p.
 * taps [P,N,N] float32 (A_p row-major), 1 <= P <= 64; everything else as dgfdn_solve_fwd / dgfdn_solve_bwd (float64
 * elimination, the adjoint re-eliminates M_k^H). The adjoint also returns lambda [K,N] c64: the tap gradients are
 *   gtaps[p,i,j] = Re sum_k lambda[k,i] conj(x[k,j]) conj(z_k^-p)        (transposed back if transpose_a)
 * a P-term DFT-weighted sum of outer products that the caller forms. ws: dgfdn_solve_bwd_ws_bytes(n) bytes. */
DGFDN_API int dgfdn_solve_fir_fwd(int n, int g, int ntaps, int64_t k, const void* z, const int32_t* delays, const float* taps,
                        int transpose_a, const float* gamma, const void* gamma_z, const float* b, const float* c,
                        void* x, void* y, void* stream);
DGFDN_API int dgfdn_solve_fir_bwd(int n, int g, int ntaps, int64_t k, const void* z, const int32_t* delays, const float* taps,
                        int transpose_a, const float* gamma, const void* gamma_z, const float* c, const void* x,
                        const void* gy, const void* gx, void* lam, double* gb, double* gc, double* ginvgamma, void* ws,
                        void* stream);

/* Group mode of K1: the G independent lossless LxL systems of DiffGFDN.sub_fdn_output (model.py:209-252, quirk Q1:
 * RAW mixing matrices, normally no absorption) solved as G small systems per bin -- four 8x8 systems to a warp --
 * instead of one block-diagonal NxN system.
 *   x[k, g L + i] = ((diag(z_k^{m_g} / gamma_g) - M_g)^{-1} b_g)[i],   y[k,g] = sum_i c[g L + i] x[k, g L + i]
 * m_raw [G,L,L] float32; delays, gamma (or NULL), b, c [G*L]; x [K, G*L] c64 (may be NULL), y [K,G] c64.
 * The adjoint returns gm [G,L,L], gb, gc, ginvgamma [G*L] (float64); ws: dgfdn_solve_groups_bwd_ws_bytes(l) bytes.
 * factors: as for dgfdn_solve_fwd/bwd, dgfdn_solve_groups_factors_bytes(l, g, k) bytes or NULL. */
DGFDN_API int64_t dgfdn_solve_groups_factors_bytes(int l, int g, int64_t k);
DGFDN_API int dgfdn_solve_groups_fwd(int l, int g, int64_t k, const void* z, const int32_t* delays, const float* m_raw,
                           const float* gamma, const float* b, const float* c, void* x, void* y, void* factors,
                           void* stream);
DGFDN_API int64_t dgfdn_solve_groups_bwd_ws_bytes(int l);
DGFDN_API int dgfdn_solve_groups_bwd(int l, int g, int64_t k, const void* z, const int32_t* delays, const float* m_raw,
                           const float* gamma, const float* c, const void* x, const void* gy, const void* gx,
                           double* gm, double* gb, double* gc, double* ginvgamma, void* ws, const void* factors,
                           void* stream);

/* K1c: the colorless branch of a training step in one pass per bin (trainer.py:298-308 with DiffGFDN.sub_fdn_output,
 * model.py:209-252, and mse_loss / amse_loss, colorless_fdn/losses.py:20-73):
 *   loss[g] = mean_k (|y[k,g]| - 1)^p,  y[k,g] = c_g^T (diag(z_k^{m_g} / gamma_g) - M_g)^{-1} b_g,  p = 4 where asym and
 *   |y| - 1 > 1, else 2 -- together with dloss[g]/dM_g, /db, /dc (float64: gm [G,L,L], gb, gc [G*L]), the adjoint solved with
 * the elimination still in registers. L <= 16. ws: dgfdn_solve_colorless_ws_bytes(l) bytes.
 * max_sms: 0 = a grid for the whole device; > 0 = at most that many SMs' worth of resident blocks (the fused step runs this
 * kernel next to the receiver kernel, on the SMs its thread-block clusters leave idle). */
DGFDN_API int64_t dgfdn_solve_colorless_ws_bytes(int l);
DGFDN_API int dgfdn_solve_colorless(int l, int g, int64_t k, const void* z, const int32_t* delays, const float* m_raw,
                          const float* gamma, const float* b, const float* c, int asym, int max_sms, double* loss,
                          double* gm, double* gb, double* gc, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2: receiver projection.  Replaces the (B,N,K) expansion + einsums of model.py:583-619.
 *   H[r,k] = sum_g s[r,g] y[k,g] + d[r,k]
 * s [R,G] float32, y [K,G] c64, d [R,K] c64 or NULL, h [R,K] c64 out (row stride ldh elements).
 */
DGFDN_API int dgfdn_project_fwd(int g, int64_t rows, int64_t k, const float* s, const void* y, const void* d, int64_t ldd,
                      void* h, int64_t ldh, void* stream);
/* Adjoint:  gy[k,g] (+)= sum_r s[r,g] gh[r,k] ;  gs[r,g] = Re sum_k conj(y[k,g]) gh[r,k].
 * accumulate_gy != 0 adds into gy (receiver tiles processed one after another). gs/gy may be NULL. */
DGFDN_API int dgfdn_project_bwd(int g, int64_t rows, int64_t k, const float* s, const void* y, const void* gh, int64_t ldh,
                      void* gy, int accumulate_gy, float* gs, void* stream);

/* K2s: receiver projection through SVF output filters (use_svfs: True).  Replaces SVF_from_MLP.forward's loop over
 * (receiver, group, section) with its (B,N,K) complex filter tensor (gain_filters.py:383-401), SOSFilter.forward
 * (gain_filters.py:221-241) and the einsums of model.py:583-619:
 *   F[r,g,k] = prod_s (b0 + b1 z_k^-1 + b2 z_k^-2) / (a0 + a1 z_k^-1 + a2 z_k^-2) ;  H[r,k] = sum_g F[r,g,k] y[k,g] + d[r,k]
 * coef [R,G,S,6] float64 = (b0 b1 b2 a0 a1 a2) per section (S <= 16), z [K] c128, y [K,G] c64, d [R,K] c64 or NULL,
 * h [R,K] c64 out. Coefficients and response are float64 (the reference evaluates on its complex128 z grid; the
 * denominators cancel to ~4 f_c^2 near DC, which float32 coefficients resolve to 1e-3 only). */
DGFDN_API int dgfdn_project_svf_fwd(int g, int nsec, int64_t rows, int64_t k, const double* coef, const void* z, const void* y,
                          const void* d, int64_t ldd, void* h, int64_t ldh, void* stream);
/* Adjoint: gy[k,g] = sum_r conj(F[r,g,k]) gh[r,k] ;  gcoef[r,g,s,:] = Re sum_k conj(gh[r,k]) dH[r,k]/dcoef (deterministic
 * two-stage reduction over bins; ws of dgfdn_project_svf_bwd_ws_bytes bytes). gy / gcoef may be NULL. rows <= 65535. */
DGFDN_API int64_t dgfdn_project_svf_bwd_ws_bytes(int g, int nsec, int64_t rows, int64_t k);
DGFDN_API int dgfdn_project_svf_bwd(int g, int nsec, int64_t rows, int64_t k, const double* coef, const void* z, const void* y,
                          const void* gh, int64_t ldh, double* gcoef, void* gy, void* ws, void* stream);

/* Directional (SH) projection, model.py:1056-1088:
 *   H_sh[r,l,k] = sum_g cw[r,g,l] x[k, g L + l]          cw = w o c  [R,G,L] float32, x [K,N] c64
 * adjoint: gcw[r,g,l] = Re sum_k conj(x[k,gL+l]) gh[r,l,k] ; gx[k,gL+l] (+)= sum_r cw[r,g,l] gh[r,l,k] */
DGFDN_API int dgfdn_project_sh_fwd(int g, int l, int64_t rows, int64_t k, const float* cw, const void* x, void* h_sh,
                         void* stream);
DGFDN_API int dgfdn_project_sh_bwd(int g, int l, int64_t rows, int64_t k, const float* cw, const void* x, const void* gh,
                         void* gx, int accumulate_gx, float* gcw, void* stream);

/* Channel mix, trainer.py:853-865 (SH -> directions):  out[r,j,k] = sum_l w[j,l] in[r,l,k].
 * The adjoint is the same call with w transposed. w [cout,cin] float32 row-major, in/out c64. */
DGFDN_API int dgfdn_mix_channels(int cin, int cout, int64_t rows, int64_t k, const float* w, const void* in, void* out,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3a: real inverse DFT of arbitrary length n on a time window, as a chirp-z transform over
 * power-of-two cuFFT C2C passes.  Replaces torch.fft.irfft(X, n)[..., t0:t0+tn] at losses.py:207-213
 * (n = K, quirk Q3), :344-346 (n = 2(K-1)) and :442-445, and utils.py:169.
 *   out[r,t] = (1/n) Re sum_{k=0}^{n/2} w_k filt[k] X[r,k] e^{+2 pi i k (t0+t)/n},  w = 1 (DC, Nyquist) or 2
 * Only bins 0..n/2 of each row are read; Im of DC (and Nyquist for even n) is ignored, as pocketfft/cuFFT C2R do.
 */
typedef struct dgfdn_czt_plan dgfdn_czt_plan;
DGFDN_API int dgfdn_czt_plan_create(int64_t n, int64_t t0, int64_t tn, dgfdn_czt_plan** plan);
DGFDN_API int dgfdn_czt_plan_destroy(dgfdn_czt_plan* plan);
/* length (complex elements) of one scratch row; callers allocate rows * mc c64 of scratch */
DGFDN_API int64_t dgfdn_czt_plan_mc(const dgfdn_czt_plan* plan);
/* x [rows, ldx] c64 (>= n/2+1 valid bins per row); filt [n/2+1] c64 or NULL; scratch [rows, mc] c64;
 * out [rows, tn] float32 */
DGFDN_API int dgfdn_irfft_window_fwd(dgfdn_czt_plan* plan, const void* x, int64_t ldx, int64_t rows, const void* filt,
                           void* scratch, float* out, void* stream);
/* adjoint: gx[r,k] = conj(filt[k]) w_k/n sum_t gout[r,t] e^{-2 pi i k (t0+t)/n} for k <= n/2 (imag zeroed at DC /
 * Nyquist), zero for n/2 < k < kx. gx [rows, ldx] c64 with kx columns written per row. */
DGFDN_API int dgfdn_irfft_window_bwd(dgfdn_czt_plan* plan, const float* gout, int64_t rows, const void* filt, void* scratch,
                           void* gx, int64_t ldx, int64_t kx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3b: Schroeder energy-decay curve + dB loss.  Replaces schroeder_backward_integral + db + mean|.|
 * (losses.py:187-199, 217-238; utils.py:16-40) and losses.py:349-369.
 *   EDC[r,t] = sum_{tau>=t} h[r,tau]^2 ;  dB = max(10 log10(EDC + eps_f32), -200)
 */
/* curve_db [rows, tn] float32 out */
DGFDN_API int dgfdn_edc_db(const float* h, int64_t rows, int64_t tn, float* curve_db, void* stream);
/* row_sum[r] = sum_t mask[t] | target_db[r,t] - dB(EDC[r,t]) |  (float64; mask [tn] float32 or NULL = all ones) */
DGFDN_API int dgfdn_edc_loss_fwd(const float* h, const float* target_db, const float* mask, int64_t rows, int64_t tn,
                       double* row_sum, void* stream);
/* gh[r,t] = coef * d(sum_r row_sum[r]) / d h[r,t]   (gh [rows, tn] float32 out) */
DGFDN_API int dgfdn_edc_loss_bwd(const float* h, const float* target_db, const float* mask, int64_t rows, int64_t tn,
                       double coef, float* gh, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3e: energy-decay-relief loss on an STFT.  Replaces get_edr_from_stft + db + the normalised L1 reduction of
 * edr_loss.forward (losses.py:447-495, 556-575; utils.py:16-40) and their autograd:
 *   EDR[f,m] = sum_{m' >= m} |S[f,m']|^2 ;  loss = sum_b ( sum_{f,m} |T_dB[b,m,f] - 10 log10(EDR_b[f,m] + eps)| ) / den[b]
 * s [R, T_f, F] c64 (torch.fft.rfft of the hann-windowed frames [R, T_f, win]: the STFT stays cuFFT), target_db and
 * out_db [R, T_f, F] float32 in the same layout, den [R] float64 (= sum |T_dB|), loss [1] float64 out, ws of
 * dgfdn_edr_ws_bytes bytes. bwd: gs [R, T_f, F] c64 = gloss[0] * dloss/dS (torch convention). R <= 65535. */
DGFDN_API int dgfdn_edr_db(int64_t rows, int64_t tf, int64_t f, const void* s, float* out_db, void* stream);
DGFDN_API int64_t dgfdn_edr_ws_bytes(int64_t rows, int64_t f);
DGFDN_API int dgfdn_edr_loss_fwd(int64_t rows, int64_t tf, int64_t f, const void* s, const float* target_db, const double* den,
                       double* loss, void* ws, void* stream);
DGFDN_API int dgfdn_edr_loss_bwd(int64_t rows, int64_t tf, int64_t f, const void* s, const float* target_db, const double* den,
                       const double* gloss, void* gs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3c: receiver step in the time domain.  irfft is linear and a receiver enters H_r = sum_g s[r,g] y_g + d_r
 * (model.py:583-619) only through its G real gains, so irfft(H_r)[window] = sum_g s[r,g] hy[g,:] + hd[r,:] with
 * hy = irfft(y_g)[window] (G rows per step, dgfdn_irfft_window_fwd) and hd[r,:] = irfft(d_r)[window] a constant of
 * the data set. dgfdn_td_edc_step replaces, per receiver, model.py:583-619 + losses.py:207-238 and their autograd
 * backward: it reads hd and the target EDC in dB (8 B per receiver.sample) and never touches the frequency domain.
 *   h[r,t]   = sum_g s[r,g] hy[g,t] + hd[r,t]                     (hd may be NULL)
 *   row_sum[r] = sum_t mask[t] | target_db[r,t] - dB(EDC(h[r,:])[t]) |        float64
 *   gh[r,t]  = coef * d(sum_r row_sum[r]) / d h[r,t]               float32 [rows, ldg]
 *   gs[r,g]  = sum_t gh[r,t] hy[g,t]                               float32 (may be NULL)
 * s [rows,G], hy [G,tn], hd [rows,ldhd], target_db [rows,ldt], mask [tn] or NULL (= ones).
 */
DGFDN_API int dgfdn_td_edc_step(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd,
                      int64_t ldhd, const float* target_db, int64_t ldt, const float* mask, double coef,
                      double* row_sum, float* gs, float* gh, int64_t ldg, void* stream);
/* h[r,t] = sum_g s[r,g] hy[g,t] + hd[r,t]: the late RIR window of each receiver (utils.py:169 get_response +
 * irfft, restricted to the window), float32 [rows, ldh]. */
DGFDN_API int dgfdn_td_mix(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd, int64_t ldhd,
                 float* h, int64_t ldh, void* stream);
/* ghy[g,t] (+)= sum_r s[r,g] gh[r,t]  (adjoint of the mix w.r.t. hy; fixed-order two-stage reduction).
 * ws: scratch of dgfdn_td_contract_ws_bytes(g, rows, tn) bytes. */
DGFDN_API int64_t dgfdn_td_contract_ws_bytes(int g, int64_t rows, int64_t tn);
DGFDN_API int dgfdn_td_contract(int g, int64_t rows, int64_t tn, const float* s, const float* gh, int64_t ldg, float* ghy,
                      int accumulate, void* ws, void* stream);

/* K3d: the same receiver step with NO per-receiver output (replaces dgfdn_td_edc_step + dgfdn_td_contract, i.e. per
 * receiver model.py:583-619 + losses.py:207-238 + their autograd backward, when only the parameter-side gradients
 * are wanted). A cluster of 8 CTAs owns a row (one time slice each, scan carries over distributed shared memory),
 * inputs arrive by TMA bulk copies, and the ghy accumulators stay in registers across rows: dL/dh is never stored.
 * HBM traffic: hd 4 B + target_db 4 B per receiver.sample.
 *   loss_sum[0] (+)= sum_{r,t} mask[t] | target_db[r,t] - dB(EDC(h[r,:])[t]) |        float64 (may be NULL)
 *   gs[r,g]      = coef * d loss_sum / d s[r,g]                                     float32 [rows, G] (may be NULL)
 *   ghy[g,t]    (+)= coef * d loss_sum / d hy[g,t]                                  float32 [G, tn]
 * accumulate != 0 adds to loss_sum / ghy instead of overwriting them (receiver tiles). Requirements, reported by
 * dgfdn_td_edc_fused_supported(g, tn): 1 <= g <= 4, tn % 4 == 0, tn <= 94720; all rows 16-byte aligned.
 * Two kernels sit behind this entry point: K3d (a cluster of 8 CTAs per row, tn <= 55296) and K3t (time-sliced
 * persistent kernel: CTA c owns time slice c of EVERY row, slice totals travel through L2), which takes the longer
 * windows; DGFDN_TD_KERNEL=cluster|sliced forces one where both take the shape.
 * ws: scratch of dgfdn_td_edc_fused_ws_bytes(g, rows, tn) bytes, initialised ONCE with dgfdn_td_edc_fused_ws_init
 * (K3t's carry words double as their own "not published" flags; every launch leaves them reset). A workspace serves
 * launches of <= rows rows, one at a time. */
DGFDN_API int dgfdn_td_edc_fused_supported(int g, int64_t tn);
DGFDN_API int64_t dgfdn_td_edc_fused_ws_bytes(int g, int64_t rows, int64_t tn);
DGFDN_API int dgfdn_td_edc_fused_ws_init(void* ws, int g, int64_t rows, int64_t tn, void* stream);
DGFDN_API int dgfdn_td_edc_fused(int g, int64_t rows, int64_t tn, const float* s, const float* hy, const float* hd,
                       int64_t ldhd, const float* target_db, int64_t ldt, const float* mask, double coef,
                       double* loss_sum, float* gs, float* ghy, int accumulate, void* ws, void* stream);
/* Diagnostic: which tile variant dgfdn_td_edc_fused uses for (g, tn) on the current device [-1: unsupported], its
 * cluster size, threads per CTA, and how many clusters are co-resident (= the launch grid / cluster size). The
 * environment variable DGFDN_TD_VARIANT=<id> forces a variant when the row fits its slices (tuning only). */
DGFDN_API int dgfdn_td_edc_fused_info(int g, int64_t tn, int* variant, int* cluster_size, int* threads, int* clusters);

/* ------------------------------------------------------------------------------------------------
 * K7: position -> gain network of a receiver shard.  Replaces SinusoidalEncoding + MLP / MLP_SkipConnections
 * + ScaledSigmoid (dnn.py:21-36, 89-126, 284-400) as called by Gains_from_MLP.forward (gain_filters.py:497-536)
 * and Directional_Beamforming_Weights_from_MLP.forward (spatial_sampling/model.py:169-190):
 *   a_0 = [sin(f_i pi p), cos(f_i pi p)]_i ;  a_l = relu(LayerNorm_l(W_l a_{l-1} + b_l)) (+ a_{l-1}, residual, l >= 1)
 *   out = W_out a_last + b_out ;  final_act = 1: out <- lo + (hi - lo) / (1 + exp(-out)).
 * pos [rows,3] and freq [nfeat] (= f_i pi) are float32 (pos_is_double = 0) or float64; in_dim = 6 nfeat.
 * params_host: HOST array of 4 nl + 2 DEVICE pointers {W_l [neurons,in_l], b_l, ln_gamma_l, ln_beta_l}_l, W_out
 * [out_dim,neurons], b_out (the torch parameter layouts), 16-byte aligned. neurons in {64,128}, nl <= 8, out_dim <= 32.
 * fwd: out [rows,out_dim]; saved for the backward: xhat [nl,rows,neurons], rstd [nl,rows], asave [nl,rows,neurons]
 * (residual variant only, else NULL).
 * bwd: gout [rows,out_dim] -> grad: flat float32 [dgfdn_mlp_num_params] in the order of params_host; ws: scratch of
 * dgfdn_mlp_bwd_ws_bytes. Deterministic (fixed-order two-stage reduction). */
DGFDN_API int dgfdn_mlp_supported(int in_dim, int nfeat, int neurons, int nl, int out_dim);
DGFDN_API int64_t dgfdn_mlp_num_params(int in_dim, int neurons, int nl, int out_dim);
DGFDN_API int64_t dgfdn_mlp_bwd_ws_bytes(int64_t rows, int in_dim, int neurons, int nl, int out_dim);
DGFDN_API int dgfdn_mlp_fwd(int64_t rows, int in_dim, int nfeat, int neurons, int nl, int out_dim, int residual,
                  int final_act, float lo, float hi, int pos_is_double, const void* pos, const void* freq,
                  const float* const* params_host, float* out, float* xhat, float* rstd, float* asave, void* stream);
DGFDN_API int dgfdn_mlp_bwd(int64_t rows, int in_dim, int nfeat, int neurons, int nl, int out_dim, int residual,
                  int final_act, float lo, float hi, int pos_is_double, const void* pos, const void* freq,
                  const float* const* params_host, const float* out, const float* xhat, const float* rstd,
                  const float* asave, const float* gout, float* grad, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Colorless (spectral flatness) loss of the lossless sub-FDNs, colorless_fdn/losses.py:20-73 with
 * y_true = 1:   loss[g] = mean_k (|H[k,g]| - 1)^p ,  p = 2, or 4 where |H|-1 > 1 when asym != 0.
 * h_sub [K,G] c64; loss [G] float64 out. bwd: gh[k,g] = coef[g] * dloss[g]/dH[k,g]  (c64 out). */
DGFDN_API int dgfdn_colorless_fwd(int g, int64_t k, const void* h_sub, int asym, double* loss, void* stream);
DGFDN_API int dgfdn_colorless_bwd(int g, int64_t k, const void* h_sub, int asym, const double* coef, void* gh, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6: block-recursive time-domain renderer (no reference implementation exists; ground truth is
 * irfft(H), utils.py:169, and the band sum of run_subband_training_treble.py:316-358).
 *   x_i[t] = gamma_i ( sum_j A_ij x_j[t - m_i] + b_i u[t - m_i] ),   q[band, t, g] = sum_{i in g} c_i x_i[t]
 * One CTA per band advances min(m) samples per step with the state history in shared/global memory.
 * a [bands,N,N], gamma/b/c [bands,N] float32, delays [bands,N] int32, u [T] float32 (NULL = unit impulse),
 * hist: scratch [bands, T, N] float32, q [bands, T, G] float32 out.
 */
DGFDN_API int dgfdn_render_groups(int bands, int n, int g, int64_t t, const int32_t* delays, const float* a,
                        const float* gamma, const float* b, const float* c, const float* u, float* hist, float* q,
                        void* stream);
/* listener mix: out[r,t] = sum_band sum_g s[band, pos(r,t), g] q[band,t,g]; pos = traj[r, t / hop] indexes the
 * receiver-gain table s [bands, P, G]; traj [R, ceil(T/hop)] int32; out [R,T] float32. */
DGFDN_API int dgfdn_render_mix(int bands, int g, int64_t t, int64_t listeners, int64_t positions, int64_t hop,
                     const float* s, const int32_t* traj, const float* q, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exchange steps of the bin-sharded multi-GPU step over NVLink peer memory (csrc/peer.cu): every rank of the node owns one
 * symmetric buffer (same layout everywhere) and holds the device pointers of all of them (peer_ptrs[world], host array;
 * P2P mappings from the caller, e.g. torch's symmetric-memory rendezvous). Layout of a buffer: flags at flags_off
 * (dgfdn_peer_flags_bytes() bytes, zero before first use), and per channel two data regions (parity of the channel's sequence
 * number) of region_bytes at data_off, the slot of source rank r at r * slot_stride inside a region.
 *   dgfdn_peer_push    stores src (nbytes) into this rank's slot in EVERY rank's buffer, then raises this rank's flag there;
 *   dgfdn_peer_gather  waits for every rank's flag of the channel's current sequence number, out <- the world slots in rank order;
 *   dgfdn_peer_reduce  the same wait, out[i] = sum over the slots in rank order (float32; deterministic, identical on all ranks).
 * Replaces dist.all_gather_into_tensor / dist.all_reduce for the three small exchanges of a step (no counterpart in the
 * reference, which is single-process). state: dgfdn_peer_state_bytes() bytes of zeroed local device memory, one per
 * buffer. All sizes / offsets multiples of 16 bytes. Stream ordered, no host synchronisation (CUDA-graph capturable). */
DGFDN_API int64_t dgfdn_peer_state_bytes(void);
DGFDN_API int64_t dgfdn_peer_flags_bytes(void);
DGFDN_API int dgfdn_peer_push(int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off, int64_t data_off,
                    int64_t region_bytes, int64_t slot_stride, void* state, const void* src, int64_t nbytes, void* stream);
DGFDN_API int dgfdn_peer_gather(int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off,
                      int64_t data_off, int64_t region_bytes, int64_t slot_stride, void* state, void* out, int64_t nbytes,
                      void* stream);
DGFDN_API int dgfdn_peer_reduce(int world, int rank, int channel, const void* const* peer_ptrs, int64_t flags_off,
                      int64_t data_off, int64_t region_bytes, int64_t slot_stride, void* state, void* out, int64_t nbytes,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFGFDN_B200_H */
