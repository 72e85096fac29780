"""CPU oracle for the DiffGFDN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64, torch-CPU, autograd-differentiable restatement of the reference algorithm
(orchidas/DiffGFDN @ f63c4bf) for the frequency-sampled Grouped-FDN transfer function, its
EDC / EDR / directional-EDC / colorless losses and the loss composition of the trainer.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / baseline -- never on the product path.

Parity pinning: the reference ships NO golden vectors for this path, so
this oracle is pinned against outputs of the reference itself: `oracle/gen_golden.py` imports
the unmodified reference (via `oracle/ref_shim.py`) in the build container and freezes its
outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks every function here
against those fixtures (and, when /root/reference is present, against the live reference).

Every function cites the reference file:line (relative to /root/reference/src) it follows.
All quirks listed in SURVEY.md section 8a (Q1-Q6) are reproduced on purpose.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

F64 = torch.float64
C128 = torch.complex128
EPS_F32 = float(np.finfo(np.float32).eps)


def _t(x, dtype=F64):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# --------------------------------------------------------------------------------------
# parameter pre-processing (a-2 .. a-5)
# --------------------------------------------------------------------------------------
def db2lin(x):
    """diff_gfdn/utils.py:43-59 (root-power quantity: 10**(x/20))."""
    return torch.pow(torch.tensor(10.0, dtype=F64), _t(x) * 0.05)


def decay_times_to_gain_per_sample(t60_per_group: Sequence[float], delays: Sequence[int], fs: float,
                                   num_groups: int) -> torch.Tensor:
    """diff_gfdn/absorption_filters.py:40-53, used by model.py:155-163.

    gamma_i = 10^(-3 m_i / (fs T60_g(i))): a gain for the WHOLE delay line i."""
    delays = _t(delays)
    n = delays.numel()
    per = n // num_groups
    t60 = _t(t60_per_group).reshape(-1).repeat_interleave(per)
    return db2lin(-60.0 * delays / (fs * t60))


def ortho_param(m: torch.Tensor) -> torch.Tensor:
    """feedback_loop.py:16-36,270: expm(triu(M,1) - triu(M,1)^T)."""
    a = m.triu(1)
    return torch.matrix_exp(a - a.transpose(-1, -2))


def _unit(n, r, c, dtype):
    e = torch.zeros(n, n, dtype=dtype)
    e[r, c] = 1.0
    return e


def nd_unitary(alpha: torch.Tensor, n: int) -> torch.Tensor:
    """feedback_loop.py:60-87: U_n = R_{n-2} ... R_0 [[U_{n-1}, 0], [0, 1]], R_i a planar rotation of
    rows/cols (i, n-1) (feedback_loop.py:42-58). Written without in-place writes so autograd sees alpha."""
    assert alpha.numel() == n * (n - 1) // 2
    dt = alpha.dtype
    if n == 1:
        return torch.ones(1, 1, dtype=dt)
    start = (n - 1) * (n - 2) // 2
    cur = alpha[start:]
    rot = torch.eye(n, dtype=dt)
    for i in range(n - 1):
        c, s = torch.cos(cur[i]), torch.sin(cur[i])
        diag = _unit(n, i, i, dt) + _unit(n, n - 1, n - 1, dt)
        r = torch.eye(n, dtype=dt) - diag + c * diag - s * _unit(n, i, n - 1, dt) + s * _unit(n, n - 1, i, dt)
        rot = r @ rot
    big = torch.nn.functional.pad(nd_unitary(alpha[:start], n - 1), (0, 1, 0, 1)) + _unit(n, n - 1, n - 1, dt)
    return rot @ big


def coupled_feedback_matrix(m_raw: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """feedback_loop.py:393-455 (SCALAR coupling).

    block_M[i,j] = U_i U_j (diagonal blocks are U_i^2: quirk Q4), Phi = nd_unitary(clamp(alpha)),
    A = block_M o (Phi (x) 1_{LxL})."""
    g, l, _ = m_raw.shape
    u = ortho_param(m_raw)
    rows = []
    for i in range(g):
        rows.append(torch.cat([u[i] @ u[j] for j in range(g)], dim=1))
    block_m = torch.cat(rows, dim=0)
    phi = nd_unitary(alpha.clamp(min=-math.pi, max=math.pi), g)
    return block_m * torch.kron(phi, torch.ones(l, l, dtype=m_raw.dtype))


# --------------------------------------------------------------------------------------
# frequency grid (a-1)
# --------------------------------------------------------------------------------------
def z_grid(nfft: int, radius: float = 1.0) -> torch.Tensor:
    """dataloader.py:552-566: z_k = r exp(j 2 pi k / nfft), k = 0..nfft/2 (complex128)."""
    w = torch.as_tensor(np.fft.rfftfreq(nfft) * 2.0 * np.pi, dtype=F64)
    return torch.polar(torch.full_like(w, radius), w)


# --------------------------------------------------------------------------------------
# feedback loop (a-6) and transfer functions (a-8, a-8b, a-9)
# --------------------------------------------------------------------------------------
def paraunitary_coupling(unitary_matrix: torch.Tensor, unit_vectors: torch.Tensor, eps: float = 1e-9) -> torch.Tensor:
    """feedback_loop.py:90-143 (FIRParaunitary) as called by construct_coupling_matrix (:414-421):
    Phi(z) = H_{P-2}(z) ... H_0(z) U,  H_k(z) = (I - v_k v_k^T) + v_k v_k^T z^-1 with v_k the k-th column of
    unit_vectors normalised to unit length (+ eps), U = expm(skew(unitary_matrix)). Returns (G, G, P) real taps."""
    g = unitary_matrix.shape[0]
    v = unit_vectors / (torch.norm(unit_vectors, dim=0, keepdim=True) + eps)
    poly = [torch.eye(g, dtype=unitary_matrix.dtype)]  # taps of the running product
    for k in range(unit_vectors.shape[1]):
        vv = torch.outer(v[:, k], v[:, k])
        h0, h1 = torch.eye(g, dtype=vv.dtype) - vv, vv
        nxt = [torch.zeros(g, g, dtype=vv.dtype) for _ in range(len(poly) + 1)]
        for i, tap in enumerate(poly):  # matrix_convolution(H_k, poly): H_k on the left (utils.py:216-239)
            nxt[i] = nxt[i] + h0 @ tap
            nxt[i + 1] = nxt[i + 1] + h1 @ tap
        poly = nxt
    u = ortho_param(unitary_matrix)
    return torch.stack([tap @ u for tap in poly], dim=-1)


def coupled_feedback_matrix_filter(m_raw: torch.Tensor, phi: torch.Tensor) -> torch.Tensor:
    """feedback_loop.py:447-453: A[..., p] = block_M o (Phi[..., p] (x) 1_{LxL}); (N, N, P) real taps of A(z)."""
    g, l, _ = m_raw.shape
    u = ortho_param(m_raw)
    block_m = torch.cat([torch.cat([u[i] @ u[j] for j in range(g)], dim=1) for i in range(g)], dim=0)
    ones = torch.ones(l, l, dtype=m_raw.dtype)
    return torch.stack([block_m * torch.kron(phi[..., p], ones) for p in range(phi.shape[-1])], dim=-1)


def feedback_loop_inverse(z: torch.Tensor, delays: torch.Tensor, gamma: torch.Tensor,
                          a: torch.Tensor) -> torch.Tensor:
    """feedback_loop.py:326-391: P_k = (diag(z_k^m) Gamma^-1 - A)^-1  (quirk Q5: D Gamma^-1, not D - A Gamma).

    gamma is (N,) real (scalar absorption) or (N, K) complex (filter absorption Gamma_i(z_k)). a is (N, N), or
    (N, N, P) taps of a FIR coupling A(z_k) = sum_p a[..., p] z_k^-p (filter_matrix coupling, :362-373)."""
    z = z.to(C128)
    d = z.unsqueeze(-1)**delays.to(F64)  # (K, N)                      feedback_loop.py:330
    if gamma.dim() == 1:
        dd = d / gamma.to(C128).unsqueeze(0)  #                        feedback_loop.py:385-386
    else:
        dd = d / gamma.to(C128).transpose(0, 1)  #                     feedback_loop.py:378-381
    if a.dim() == 3:
        zp = z.unsqueeze(-1)**(-torch.arange(a.shape[-1], dtype=F64))  # (K, P)
        # the reference casts A(z) to complex64 before the subtraction (:373)
        az = torch.einsum('nmp,kp->knm', a.to(C128), zp).to(torch.complex64).to(C128)
    else:
        az = a.to(C128).unsqueeze(0)
    m = torch.diag_embed(dd) - az
    return torch.linalg.inv(m)  #                                      feedback_loop.py:391


def gfdn_state(z, delays, gamma, a, b) -> torch.Tensor:
    """x_k = P_k b, (K, N) complex128. Same arithmetic as einsum over P (model.py:615-619, 1083)."""
    p = feedback_loop_inverse(z, delays, gamma, a)
    return torch.einsum('knm,m->kn', p, b.to(C128))


def omni_response(z, delays, gamma, a, b, c, s, d=None) -> torch.Tensor:
    """model.py:569-625 (DiffGFDNVarReceiverPos.forward).

    H[r,k] = sum_{n,m} (s[r,g(n)] c_n) P_k[n,m] b_m + d[r,k];  s is (B, G) receiver gains."""
    p = feedback_loop_inverse(z, delays, gamma, a)
    n = delays.numel()
    g = s.shape[1]
    cfull = (s.to(F64).repeat_interleave(n // g, dim=1) * c.to(F64).unsqueeze(0)).to(C128)  # (B, N)
    htemp = torch.einsum('bn,knm->bmk', cfull, p)  #                   model.py:615-616
    h = torch.einsum('bmk,m->bk', htemp, b.to(C128))  #                model.py:619
    if d is not None:
        h = h + d.to(C128)
    return h


def directional_response(z, delays, gamma, a, b, c, w) -> torch.Tensor:
    """model.py:1043-1094 (DiffDirectionalFDNVarReceiverPos.forward).

    H_sh[r,l,k] = sum_g w[r,g,l] c[g,l] (P_k^T b)[g L + l]; w is (B, G, L) (already L2-normalised).
    NOTE (quirk Q11): einsum('knm,bnk->bmk', P, B) at model.py:1083 contracts the FIRST matrix index of P,
    i.e. the state is P_k^T b, not P_k b as in the omni model."""
    p = feedback_loop_inverse(z, delays, gamma, a)
    x = torch.einsum('knm,n->km', p, b.to(C128))  # (K, N)            model.py:1083
    bsz, g, l = w.shape
    xg = x.transpose(0, 1).reshape(g, l, -1)  #                       model.py:1084-1085
    cg = c.to(F64).reshape(g, l)
    cw = (w.to(F64) * cg.unsqueeze(0)).to(C128)  #                    model.py:1056-1073
    return torch.einsum('bgl,glk->blk', cw, xg)  #                    model.py:1088


def sub_fdn_output(z, delays, m_raw, b, c) -> Tuple[torch.Tensor, torch.Tensor]:
    """model.py:209-252: per-group lossless response with the RAW M_g (quirk Q1) and no absorption.

    Returns Hout (K, G) and Hout_per_del (N, K, G)."""
    z = z.to(C128)
    g, l, _ = m_raw.shape
    n = g * l
    k = z.numel()
    hout = torch.zeros(k, g, dtype=C128)
    hper = torch.zeros(n, k, g, dtype=C128)
    cols, pers = [], []
    for gi in range(g):
        idx = slice(gi * l, (gi + 1) * l)
        dmat = torch.diag_embed(z.unsqueeze(-1)**delays[idx].to(F64))  # model.py:238-239
        p = torch.linalg.inv(dmat - m_raw[gi].to(C128).unsqueeze(0))  # model.py:237,240
        xs = torch.einsum('knm,m->kn', p, b[idx].to(C128))
        per = xs * c[idx].to(C128).unsqueeze(0)  # (K, L)               model.py:243-246
        cols.append(per.sum(dim=1))  #                                model.py:249-250
        pers.append(per)
    hout = torch.stack(cols, dim=1)
    blocks = []
    for gi in range(g):
        blk = torch.zeros(l, k, g, dtype=C128)
        onehot = torch.zeros(g, dtype=C128)
        onehot[gi] = 1.0
        blk = pers[gi].transpose(0, 1).unsqueeze(-1) * onehot
        blocks.append(blk)
    hper = torch.cat(blocks, dim=0)
    return hout, hper


# --------------------------------------------------------------------------------------
# receiver-gain generators (a-7, a-7c)
# --------------------------------------------------------------------------------------
def sinusoidal_encoding(pos: torch.Tensor, num_fourier_features: int) -> torch.Tensor:
    """dnn.py:103-126: [sin(f_i pi p), cos(f_i pi p)] for f = exp(linspace(ln 1, ln 32, F))."""
    pos = pos.to(F64)
    freqs = torch.exp(torch.linspace(math.log(1.0), math.log(32.0), num_fourier_features, dtype=F64))
    chunks = []
    for k in range(num_fourier_features):
        chunks.append(torch.sin(freqs[k] * math.pi * pos))
        chunks.append(torch.cos(freqs[k] * math.pi * pos))
    return torch.cat(chunks, dim=-1)


def mlp_forward(x: torch.Tensor, weights: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """dnn.py:331-400 (MLP): Linear -> LayerNorm -> ReLU repeated, final Linear.

    `weights` is a state_dict; `prefix` e.g. 'output_scalars.mlp.model.'; layer indices 0,1,3,4,...,last."""
    idx = sorted({int(k[len(prefix):].split('.')[0]) for k in weights if k.startswith(prefix)})
    lin = [i for i in idx if weights[f'{prefix}{i}.weight'].dim() == 2]
    h = x.to(F64)
    for j, i in enumerate(lin):
        w = weights[f'{prefix}{i}.weight'].to(F64)
        bb = weights[f'{prefix}{i}.bias'].to(F64)
        h = h @ w.t() + bb
        if j < len(lin) - 1:
            lw = weights[f'{prefix}{i + 1}.weight'].to(F64)
            lb = weights[f'{prefix}{i + 1}.bias'].to(F64)
            h = torch.nn.functional.layer_norm(h, (h.shape[-1], ), lw, lb, 1e-5)
            h = torch.relu(h)
    return h


def mlp_skip_forward(x: torch.Tensor, weights: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """dnn.py:262-328 (MLP_SkipConnections): input_layer, residual blocks, output_layer."""
    h = x.to(F64)
    h = h @ weights[f'{prefix}input_layer.0.weight'].to(F64).t() + weights[f'{prefix}input_layer.0.bias'].to(F64)
    h = torch.nn.functional.layer_norm(h, (h.shape[-1], ), weights[f'{prefix}input_layer.1.weight'].to(F64),
                                       weights[f'{prefix}input_layer.1.bias'].to(F64), 1e-5)
    h = torch.relu(h)
    i = 0
    while f'{prefix}hidden_layers.{i}.linear.weight' in weights:
        r = h
        o = h @ weights[f'{prefix}hidden_layers.{i}.linear.weight'].to(F64).t() + \
            weights[f'{prefix}hidden_layers.{i}.linear.bias'].to(F64)
        o = torch.nn.functional.layer_norm(o, (o.shape[-1], ), weights[f'{prefix}hidden_layers.{i}.norm.weight'].to(F64),
                                           weights[f'{prefix}hidden_layers.{i}.norm.bias'].to(F64), 1e-5)
        h = torch.relu(o) + r
        i += 1
    return h @ weights[f'{prefix}output_layer.weight'].to(F64).t() + weights[f'{prefix}output_layer.bias'].to(F64)


def scaled_sigmoid(x, lo, hi):
    """dnn.py:21-36."""
    return lo + (hi - lo) * (1.0 / (1.0 + torch.exp(-x)))


def gains_from_mlp(norm_pos, weights, num_fourier_features, num_groups,
                   prefix='output_scalars.mlp.model.') -> torch.Tensor:
    """gain_filters.py:497-536: s[r,g] = -1 + 2 sigmoid(MLP(enc(p_r)))."""
    enc = sinusoidal_encoding(norm_pos, num_fourier_features)
    out = mlp_forward(enc, weights, prefix).reshape(norm_pos.shape[0], num_groups)
    return scaled_sigmoid(out, -1.0, 1.0)


# --------------------------------------------------------------------------------------
# SVF output filters (a-7b)
# --------------------------------------------------------------------------------------
def svf_cutoffs(fs: float) -> torch.Tensor:
    """gain_filters.py:369-374 with filters/geq.py:9-56 (eq_freqs / octave_bands, defaults 31.25 Hz .. 16 kHz):
    pi * [f_1/sqrt(2), f_1 .. f_9, f_9 sqrt(2)] / fs with f_i = 62.5 * 2^(i-1) -- used as the SVF 'g' directly
    (no tan pre-warping)."""
    centre = []
    c = 31.25
    while c < 16000:
        c = c * 2.0
        centre.append(c)
    freqs = [centre[0] / math.sqrt(2.0)] + centre + [centre[-1] * math.sqrt(2.0)]
    return math.pi * torch.tensor(freqs, dtype=F64) / fs


def svf_params_from_mlp(pos, weights, num_fourier_features, num_groups, num_biquads=11,
                        prefix='output_filters.mlp.model.') -> torch.Tensor:
    """gain_filters.py:376-382,403-421: MLP(enc(listener_position)) -> (B, G, S, 2); [...,0] resonance through
    ScaledSigmoid(1e-6, 1), [...,1] gain in dB through ScaledSigmoid(-6, 6)."""
    enc = sinusoidal_encoding(pos, num_fourier_features)
    out = mlp_forward(enc, weights, prefix).reshape(pos.shape[0], num_groups, num_biquads, 2)
    return torch.stack([scaled_sigmoid(out[..., 0], 1e-6, 1.0), scaled_sigmoid(out[..., 1], -6.0, 6.0)], dim=-1)


def svf_to_biquads(svf_params: torch.Tensor, cutoffs: torch.Tensor, pole_factor: float = 1.0) -> torch.Tensor:
    """gain_filters.py:20-102 (SVF mixing coefficients: section 0 low shelf, last high shelf, others peaking) and
    :116-151 (BiquadCascade.from_svf_coeffs). svf_params (..., S, 2) -> (..., S, 6) = [b0 b1 b2 a0 a1 a2]."""
    res, gdb = svf_params[..., 0].to(F64), svf_params[..., 1].to(F64)
    gain = torch.pow(torch.tensor(10.0, dtype=F64), gdb * 0.05)
    f = cutoffs.to(F64)
    ns = f.numel()
    one = torch.ones_like(gain)
    kind = torch.zeros(ns, dtype=torch.long)
    kind[0], kind[-1] = 1, 2  # 0 peaking, 1 low shelf, 2 high shelf
    m_lp = torch.where(kind == 1, gain, one)
    m_hp = torch.where(kind == 2, gain, one)
    m_bp = torch.where(kind == 0, 2.0 * res * gain, 2.0 * res * torch.sqrt(gain))
    r = pole_factor
    b0 = f * f * m_lp + f * m_bp + m_hp
    b1 = (2.0 * f * f * m_lp - 2.0 * m_hp) * r
    b2 = (f * f * m_lp - f * m_bp + m_hp) * r * r
    a0 = f * f + 2.0 * res * f + 1.0
    a1 = (2.0 * f * f - 2.0) * r * one
    a2 = (f * f - 2.0 * res * f + 1.0) * r * r
    return torch.stack([b0, b1, b2, a0, a1, a2], dim=-1)


def sos_response(z: torch.Tensor, coef: torch.Tensor) -> torch.Tensor:
    """gain_filters.py:221-241 (SOSFilter.forward): prod_s (b0 + b1 z^-1 + b2 z^-2) / (a0 + a1 z^-1 + a2 z^-2).
    coef (..., S, 6) -> (..., K) complex128."""
    zi = (1.0 / z.to(C128))
    zi2 = zi * zi
    c = coef.to(C128).unsqueeze(-1)  # (..., S, 6, 1)
    num = c[..., 0, :] + c[..., 1, :] * zi + c[..., 2, :] * zi2
    den = c[..., 3, :] + c[..., 4, :] * zi + c[..., 5, :] * zi2
    return torch.prod(num / den, dim=-2)


def absorption_filter_response(z: torch.Tensor, delay_filters: torch.Tensor) -> torch.Tensor:
    """feedback_loop.py:236-255, 333-340: Gamma_i(z_k) of the GEQ absorption filters. delay_filters is the
    reference's buffer (N, S, 3, 2): [..., 0] numerators, [..., 1] denominators of S biquads per delay line
    (model.py:131-147). -> (N, K) complex128 (the reference multiplies the sections in complex64)."""
    coef = torch.cat([delay_filters[..., 0], delay_filters[..., 1]], dim=-1)  # (N, S, 6)
    return sos_response(z, coef)


def omni_response_svf(z, delays, gamma, a, b, c, coef, d=None) -> torch.Tensor:
    """model.py:583-619 with use_svf_in_output: C[r,n,k] = F[r,g(n),k] c_n (gain_filters.py:388-401, every delay
    line of a group shares the group's filter); coef is (B, G, S, 6)."""
    p = feedback_loop_inverse(z, delays, gamma, a)
    n = delays.numel()
    g = coef.shape[1]
    filt = sos_response(z, coef)  # (B, G, K)
    cfull = filt.repeat_interleave(n // g, dim=1) * c.to(C128).reshape(1, n, 1)  # (B, N, K)
    htemp = torch.einsum('bnk,knm->bmk', cfull, p)
    h = torch.einsum('bmk,m->bk', htemp, b.to(C128))
    if d is not None:
        h = h + d.to(C128)
    return h


def source_receiver_response(z, delays, gamma, a, b, c, s_rx, s_src, d=None) -> torch.Tensor:
    """model.py:402-452 (DiffGFDNVarSourceReceiverPos.forward):
    H[r,k] = sum_{n,m} (C[r,g(n),k] c_n) P_k[n,m] (B[r,g(m),k] b_m) + d[r,k].
    s_rx / s_src are (B, G) real gains (Gains_from_MLP) or (B, G, K) complex filter responses (SVF_from_MLP,
    use_svf_in_output / use_svf_in_input: every delay line of a group shares the group's cascade)."""
    p = feedback_loop_inverse(z, delays, gamma, a)
    n = delays.numel()

    def expand(f, v):  # (B, N, K) complex
        g = f.shape[1]
        if f.dim() == 2:
            f = f.to(C128).unsqueeze(-1).expand(-1, -1, z.numel())
        return f.to(C128).repeat_interleave(n // g, dim=1) * v.to(C128).reshape(1, n, 1)

    cfull, bfull = expand(s_rx, c), expand(s_src, b)
    htemp = torch.einsum('bnk,knm->bmk', cfull, p)  #                  model.py:442-443
    h = torch.einsum('bmk,bmk->bk', htemp, bfull)  #                   model.py:449
    if d is not None:
        h = h + d.to(C128)
    return h


def single_position_response(z, delays, gamma, a, b, c, f_out, f_in, d=None) -> torch.Tensor:
    """model.py:779-836 (DiffGFDNSinglePos.forward): one source-receiver pair. f_out / f_in are the per-group
    receiver / source factors: (G,) learnable scalars or (G, K) filter responses (get_filter, :838-908).
    H[k] = sum_{n,m} f_out[g(n),k] c_n P_k[n,m] f_in[g(m),k] b_m + d[k]."""
    p = feedback_loop_inverse(z, delays, gamma, a)
    n = delays.numel()
    k = z.numel()

    def expand(f):
        f = f.to(C128)
        if f.dim() == 1:
            f = f.unsqueeze(-1).expand(-1, k)
        return f.repeat_interleave(n // f.shape[0], dim=0)  # (N, K)

    cz = expand(f_out) * c.to(C128).unsqueeze(-1)
    bz = expand(f_in) * b.to(C128).unsqueeze(-1)
    h = torch.einsum('nk,knm,mk->k', cz, p, bz)
    if d is not None:
        h = h + d.to(C128)
    return h


def sh_gains_from_mlp(norm_pos, weights, num_fourier_features, num_groups, num_sh,
                      prefix='sh_output_scalars.mlp.', skip=False, normalise=True) -> torch.Tensor:
    """spatial_sampling/model.py:169-190 with normalise_weights (:78-80): w / (||w||_2 + 1e-6) over l."""
    enc = sinusoidal_encoding(norm_pos, num_fourier_features)
    if skip:
        out = mlp_skip_forward(enc, weights, prefix)
    else:
        out = mlp_forward(enc, weights, prefix + 'model.')
    w = out.reshape(norm_pos.shape[0], num_groups, num_sh)
    if normalise:
        w = w / (torch.norm(w, dim=-1, keepdim=True) + 1e-6)
    return w


# --------------------------------------------------------------------------------------
# losses (a-12 .. a-15)
# --------------------------------------------------------------------------------------
def db(x: torch.Tensor, is_squared: bool = False, min_value: float = -200.0) -> torch.Tensor:
    """utils.py:16-40: factor*log10(|x| + eps_f32), clipped from below."""
    factor = 10.0 if is_squared else 20.0
    return (factor * torch.log10(torch.abs(x) + EPS_F32)).clip(min=min_value)


def schroeder(sig: torch.Tensor) -> torch.Tensor:
    """losses.py:187-199: flip(cumsum(flip(sig**2)))."""
    return torch.flip(torch.cumsum(torch.flip(sig**2, dims=[-1]), dim=-1), dims=[-1])


def ms_to_samps(ms: float, fs: float) -> int:
    """utils.py:62-80."""
    return int(ms * 1e-3 * fs)


def irfft_q3(x: torch.Tensor) -> torch.Tensor:
    """losses.py:207-213 / :442-445: torch.fft.irfft(X, n=K) with K = X.shape[-1] (quirk Q3).

    Only the first K//2+1 bins are used and K samples are returned."""
    return torch.fft.irfft(x, x.shape[-1])


def edc_curves_db(resp: torch.Tensor, max_ir_len_samps: int, mixing_time_samps: int) -> torch.Tensor:
    """losses.py:204-218 + db(): the EDC in dB of one response, (..., T)."""
    max_len = min(max_ir_len_samps, resp.shape[-1])
    rir = irfft_q3(resp)[..., mixing_time_samps:max_len]
    return db(schroeder(rir), is_squared=True)


def edc_loss(target: torch.Tensor, achieved: torch.Tensor, max_ir_len_ms: float, fs: float,
             mixing_time_ms: float = 20.0, mask_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """losses.py:201-238 (broadband branch). `mask_index` replaces the random Bernoulli mask of :221-223."""
    mx = ms_to_samps(max_ir_len_ms, fs)
    mix = ms_to_samps(mixing_time_ms, fs)
    t = edc_curves_db(target.to(C128), mx, mix)
    a = edc_curves_db(achieved.to(C128), mx, mix)
    if mask_index is not None:
        t = t[..., mask_index]
        a = a[..., mask_index]
    return torch.mean(torch.abs(t - a))


def stft_q(rir: torch.Tensor, win: int, hop: int) -> torch.Tensor:
    """losses.py:501-553: zero-pad to a hop multiple, hann(win) (periodic), center=False, onesided."""
    t = rir.shape[-1]
    if t % hop != 0:
        rir = torch.nn.functional.pad(rir, (0, hop * int(np.ceil(t / hop)) - t))
    window = torch.hann_window(win, dtype=rir.dtype)
    return torch.stft(rir, win, hop_length=hop, win_length=win, window=window, center=False, normalized=False,
                      onesided=True, return_complex=True)


def edr_db(s: torch.Tensor) -> torch.Tensor:
    """losses.py:556-575: EDR[f,m] = sum_{m'>=m} |S[f,m']|^2, in dB (float32 buffer in the reference)."""
    p = torch.abs(s)**2
    e = torch.flip(torch.cumsum(torch.flip(p, dims=[-1]), dim=-1), dims=[-1])
    return db(e, is_squared=True)


def edr_loss(target: torch.Tensor, achieved: torch.Tensor, win: int = 4096, hop: int = 2048,
             reduced_pole_radius: Optional[float] = None) -> torch.Tensor:
    """losses.py:430-495 (no ERB grouping, no frequency weights)."""
    trir = irfft_q3(target.to(C128))
    arir = irfft_q3(achieved.to(C128))
    if reduced_pole_radius is not None:
        arir = arir * torch.pow(torch.tensor(1.0 / reduced_pole_radius, dtype=F64),
                                torch.arange(arir.shape[-1], dtype=F64))
    te = edr_db(stft_q(trir, win, hop))
    ae = edr_db(stft_q(arir, win, hop))
    fl = torch.abs(te - ae).sum(dim=-1)
    if te.dim() == 3:
        return (fl.sum(dim=-1) / torch.abs(te).sum(dim=(-1, -2))).sum()
    return fl.sum() / torch.abs(te).sum()


def decay_kernel(t60s: np.ndarray, time: np.ndarray, fs: float) -> np.ndarray:
    """submodules/slope2noise/slope2noise/utils.py:173-210 with normalize_envelope=True, add_noise=False.

    t60s: (n, b) ; returns (n, t, b)."""
    tau = np.log(10**6) / t60s
    e = np.exp(-np.einsum('nb,t->ntb', tau, time))
    return np.einsum('ntb,nb->ntb', e, np.sqrt(1 - np.exp(-2 * tau / fs)))


def directional_envelopes(common_decay_times: np.ndarray, edc_len_ms: float, fs: float) -> torch.Tensor:
    """losses.py:303-317: (num_slopes, edc_len_samps) float32 envelopes."""
    n = ms_to_samps(edc_len_ms, fs)
    num_slopes = common_decay_times.shape[-1]
    t = np.linspace(0, (n - 1) / fs, n)
    env = torch.zeros(num_slopes, n, dtype=torch.float32)
    for k in range(num_slopes):
        env[k] = torch.tensor(decay_kernel(np.expand_dims(common_decay_times[:, k], -1), t, fs)).squeeze()
    return env


def sh_to_directional(h_sh: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """trainer.py:853-865: H_dir[b,j,k] = sum_l Y[j,l] H_sh[b,l,k]."""
    return torch.einsum('jl,blk->bjk', y.to(h_sh.dtype), h_sh)


def directional_edc_loss(h_dir: torch.Tensor, amps: torch.Tensor, envelopes: torch.Tensor, edc_len_samps: int,
                         mixing_time_samps: int, mask_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """losses.py:333-371: irfft with DEFAULT n = 2(K-1), slice [mix : mix+edc_len], EDC vs sum_s amps*env."""
    rir = torch.fft.irfft(h_dir.to(C128))[..., mixing_time_samps:edc_len_samps + mixing_time_samps]
    pred = schroeder(rir)
    true = torch.einsum('bjk,kt->bjt', amps.to(F64), envelopes.to(F64))
    if mask_index is not None:
        pred = pred[..., mask_index]
        true = true[..., mask_index]
    return torch.mean(torch.abs(db(true, is_squared=True) - db(pred, is_squared=True)))


def mse_loss(y_pred: torch.Tensor) -> torch.Tensor:
    """colorless_fdn/losses.py:20-41 against y_true = 1 (1-D case)."""
    return torch.mean((torch.abs(y_pred) - 1.0)**2, dim=-1)


def amse_loss(y_pred: torch.Tensor) -> torch.Tensor:
    """colorless_fdn/losses.py:44-73 against y_true = 1 (1-D case): exponent 4 where |y|-1 > 1."""
    diff = torch.abs(y_pred) - 1.0
    g = 2.0 + 2.0 * (diff > 1.0).to(F64)
    return torch.mean(torch.pow(diff, g), dim=0)


def sparsity_loss(a: torch.Tensor) -> torch.Tensor:
    """colorless_fdn/losses.py:7-17."""
    n = a.shape[-1]
    return -(torch.sum(torch.abs(a)) - n * np.sqrt(n)) / (n * (np.sqrt(n) - 1))


def colorless_losses(h_sub: torch.Tensor, m_raw: torch.Tensor, spectral_w: float, sparsity_w: float,
                     asym: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """trainer.py:295-313: spectral loss `+=` over groups; sparsity `=` (LAST group only: quirk Q2)."""
    spec = torch.zeros((), dtype=F64)
    spars = torch.zeros((), dtype=F64)
    for k in range(m_raw.shape[0]):
        spec = spec + spectral_w * (amse_loss(h_sub[..., k]) if asym else mse_loss(h_sub[..., k]))
        spars = sparsity_w * sparsity_loss(ortho_param(m_raw[k]))
    return spec, spars


def normalize_energy(h_sub: torch.Tensor) -> torch.Tensor:
    """trainer.py:317-332: E_g = mean_k |H_sub[k,g]|^2 ; caller divides b_g, c_g by E_g^(1/4)."""
    return torch.mean(torch.abs(h_sub)**2, dim=0)


# --------------------------------------------------------------------------------------
# time-domain ground truth for the renderer (a-17)
# --------------------------------------------------------------------------------------
def impulse_response(h: torch.Tensor) -> torch.Tensor:
    """utils.py:169: h = irfft(H) with default n = 2(K-1)."""
    return torch.fft.irfft(h.to(C128), dim=-1)


def fdn_time_domain(delays, gamma, a, b, c, num_samples: int) -> torch.Tensor:
    """Sample-by-sample recursion implied by H(z) of feedback_loop.py:326-391:

        x_i[n] = gamma_i ( sum_j A_ij x_j[n - m_i] + b_i u[n - m_i] ),  u = unit impulse.
    Returns the per-line states weighted by c: q[n, i] = c_i x_i[n]  (num_samples, N). Pure-python loop over
    blocks of min(m) samples -- small cases only."""
    delays = [int(v) for v in delays]
    n = len(delays)
    gam = _t(gamma)
    a = _t(a)
    bb = _t(b)
    cc = _t(c)
    x = torch.zeros(num_samples, n, dtype=F64)
    u = torch.zeros(num_samples, dtype=F64)
    u[0] = 1.0
    blk = min(delays)
    for s in range(0, num_samples, blk):
        e = min(s + blk, num_samples)
        for i in range(n):
            lo, hi = s - delays[i], e - delays[i]
            if hi <= 0:
                continue
            src_lo = max(lo, 0)
            off = src_lo - lo
            past = x[src_lo:hi]  # (len, N) all earlier than s because m_i >= blk
            val = gam[i] * (past @ a[i] + bb[i] * u[src_lo:hi])
            x[s + off:e, i] = val
    return x * cc.unsqueeze(0)
