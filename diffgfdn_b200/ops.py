"""torch.autograd bindings of the sm_100a kernels (called through the C ABI in include/diffgfdn_b200.h).

Every op takes and returns CUDA tensors, runs on torch's current stream and raises on CPU tensors -- there is no
CPU or eager fallback. Shapes follow the reference (orchidas/DiffGFDN): K frequency bins, N delay lines, G groups,
R receivers (rows)."""
import ctypes
import os
import math
from typing import Optional, Tuple

import torch

from . import _lib

C64 = torch.complex64
C128 = torch.complex128


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda(name, t: Optional[torch.Tensor], dtype=None, optional=False):
    if t is None:
        if optional:
            return None
        raise RuntimeError(f"{name}: tensor required")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor, got device {t.device} "
                           "(diffgfdn_b200 has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


# ----------------------------------------------------------------------------------------------------------
# orthogonal parametrisation (matrix exponential of the skew part), sync free
# ----------------------------------------------------------------------------------------------------------
SKEW_EXPM_MAX_L = 16


class _SkewExpm(torch.autograd.Function):
    """U = expm(triu(M,1) - triu(M,1)^T) for a batch (G, L, L) of raw mixing matrices (reference
    feedback_loop.py:16-36). Same value as torch.matrix_exp(Skew(M)) without its device->host read-backs."""

    @staticmethod
    def forward(ctx, m):
        m_ = _cuda("M", m, torch.float32)
        if m_.dim() < 2 or m_.shape[-1] != m_.shape[-2] or m_.shape[-1] > SKEW_EXPM_MAX_L:
            raise RuntimeError(f"skew_expm: expected (..., L, L) with L <= {SKEW_EXPM_MAX_L}")
        l = m_.shape[-1]
        g = m_.numel() // (l * l)
        u = torch.empty_like(m_)
        with torch.cuda.device(m_.device):
            _lib.call("dgfdn_skew_expm_fwd", g, l, _ptr(m_), _ptr(u), _stream())
        ctx.save_for_backward(m_)
        return u

    @staticmethod
    def backward(ctx, gu):
        (m_, ) = ctx.saved_tensors
        l = m_.shape[-1]
        g = m_.numel() // (l * l)
        gu_ = _cuda("gU", gu, torch.float32)
        gm = torch.empty_like(m_)
        with torch.cuda.device(m_.device):
            _lib.call("dgfdn_skew_expm_bwd", g, l, _ptr(m_), _ptr(gu_), _ptr(gm), _stream())
        return gm


def skew_expm(m: torch.Tensor) -> torch.Tensor:
    return _SkewExpm.apply(m)


# ----------------------------------------------------------------------------------------------------------
# K1: per-bin solve
# ----------------------------------------------------------------------------------------------------------
class _CoupledFeedback(torch.autograd.Function):
    """A = block(U_i U_j) o (ND_Unitary(alpha) (x) 1), float64 (reference feedback_loop.py:39-87, 393-455)."""

    @staticmethod
    def forward(ctx, u, alpha):
        u_ = _cuda("U", u, torch.float32)
        g, l, _ = u_.shape
        alpha_ = _cuda("alpha", alpha, torch.float32) if g > 1 else None
        if g > 1 and alpha_.numel() != g * (g - 1) // 2:
            raise RuntimeError("coupled_feedback: alpha must hold G(G-1)/2 angles")
        a = torch.empty(g * l, g * l, dtype=torch.float64, device=u_.device)
        phi = torch.empty(g, g, dtype=torch.float64, device=u_.device)
        with torch.cuda.device(u_.device):
            _lib.call("dgfdn_coupled_feedback_fwd", g, l, _ptr(u_), _ptr(alpha_), _ptr(a), _ptr(phi), _stream())
        ctx.save_for_backward(u_, alpha_)
        ctx.mark_non_differentiable(phi)
        return a, phi

    @staticmethod
    def backward(ctx, ga, _gphi):
        u_, alpha_ = ctx.saved_tensors
        g, l, _ = u_.shape
        ga_ = _cuda("gA", ga, torch.float64)
        gu = torch.empty_like(u_) if ctx.needs_input_grad[0] else None
        galpha = torch.empty_like(alpha_) if (alpha_ is not None and ctx.needs_input_grad[1]) else None
        with torch.cuda.device(u_.device):
            _lib.call("dgfdn_coupled_feedback_bwd", g, l, _ptr(u_), _ptr(alpha_), _ptr(ga_), _ptr(gu), _ptr(galpha),
                      _stream())
        return gu, galpha


def coupled_feedback(u: torch.Tensor, alpha: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(A (N,N) float64, Phi (G,G) float64) from the orthogonal mixing matrices U (G,L,L) and the coupling angles."""
    return _CoupledFeedback.apply(u, alpha)


def _factor_buffer(wanted: bool, size_fn: str, size_args, device) -> Optional[torch.Tensor]:
    """Buffer for the saved elimination of a K1 forward call (None when no gradient will be needed, or when
    DGFDN_SOLVE_REPLAY=0 asks for the fresh adjoint elimination)."""
    if not wanted or os.environ.get("DGFDN_SOLVE_REPLAY", "1") == "0":
        return None
    nbytes = int(getattr(_lib.load(), size_fn)(*size_args))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


class _GFDNSolve(torch.autograd.Function):
    """x_k = (diag(z_k^m / gamma) - A)^-1 b ;  y[k,g] = sum_{n in g} c_n x_k[n].

    Replaces FeedbackLoop.forward + einsums (reference feedback_loop.py:326-391, model.py:615-619, 1083, 237-250)."""

    @staticmethod
    def forward(ctx, z, delays, a, gamma, b, c, num_groups, transpose_a, gamma_z):
        z = _cuda("z", z, C128)
        delays = _cuda("delays", delays, torch.int32)
        a_ = _cuda("A", a, torch.float32)
        gamma_ = _cuda("gamma", gamma, torch.float32, optional=True)
        gamma_z_ = _cuda("gamma_z", gamma_z, C64, optional=True)
        b_ = _cuda("b", b.reshape(-1), torch.float32)
        c_ = _cuda("c", c.reshape(-1), torch.float32)
        n = a_.shape[0]
        k = z.shape[0]
        if a_.shape != (n, n) or delays.numel() != n or b_.numel() != n or c_.numel() != n:
            raise RuntimeError("gfdn_solve: inconsistent shapes")
        if gamma_z_ is not None and tuple(gamma_z_.shape) != (n, k):
            raise RuntimeError("gfdn_solve: gamma_z must be (N, K)")
        x = torch.empty(k, n, dtype=C64, device=z.device)
        y = torch.empty(k, num_groups, dtype=C64, device=z.device)
        with torch.cuda.device(z.device):
            # when a gradient will be asked for, keep the elimination: the adjoint solve replays it (no 2nd factorisation)
            factors = _factor_buffer(any(ctx.needs_input_grad), "dgfdn_solve_factors_bytes", (n, k), z.device)
            _lib.call("dgfdn_solve_fwd", n, num_groups, k, _ptr(z), _ptr(delays), _ptr(a_), int(transpose_a),
                      _ptr(gamma_), _ptr(gamma_z_), _ptr(b_), _ptr(c_), _ptr(x), _ptr(y), _ptr(factors), _stream())
        ctx.factors = factors
        ctx.save_for_backward(z, delays, a_, gamma_, gamma_z_, c_, x)
        ctx.meta = (n, num_groups, k, int(transpose_a), b.shape, c.shape)
        ctx.a_dtype = a.dtype  # float64 callers (FeedbackLoop.solve) get dL/dA back in float64
        return x, y

    @staticmethod
    def backward(ctx, gx, gy):
        z, delays, a_, gamma_, gamma_z_, c_, x = ctx.saved_tensors
        n, g, k, tr, bshape, cshape = ctx.meta
        if gx is None and gy is None:
            return (None, ) * 9
        gx_ = _cuda("gx", gx, C64, optional=True)
        gy_ = _cuda("gy", gy, C64, optional=True)
        dev = z.device
        out = torch.empty(n * n + 3 * n, dtype=torch.float64, device=dev)
        ga, gb, gc, gig = out[:n * n], out[n * n:n * n + n], out[n * n + n:n * n + 2 * n], out[n * n + 2 * n:]
        with torch.cuda.device(dev):
            ws = torch.empty(_lib.load().dgfdn_solve_bwd_ws_bytes(n) // 8, dtype=torch.float64, device=dev)
            _lib.call("dgfdn_solve_bwd", n, g, k, _ptr(z), _ptr(delays), _ptr(a_), tr, _ptr(gamma_), _ptr(gamma_z_),
                      _ptr(c_), _ptr(x), _ptr(gy_), _ptr(gx_), _ptr(ga), _ptr(gb), _ptr(gc), _ptr(gig), _ptr(ws),
                      _ptr(ctx.factors), _stream())
        g_a = ga.reshape(n, n).to(ctx.a_dtype) if ctx.needs_input_grad[2] else None
        g_gamma = None
        if gamma_ is not None and ctx.needs_input_grad[3]:
            g_gamma = (-gig / gamma_.to(torch.float64)**2).to(torch.float32)
        g_b = gb.to(torch.float32).reshape(bshape) if ctx.needs_input_grad[4] else None
        g_c = gc.to(torch.float32).reshape(cshape) if ctx.needs_input_grad[5] else None
        return None, None, g_a, g_gamma, g_b, g_c, None, None, None


def gfdn_solve(z: torch.Tensor, delays: torch.Tensor, a: torch.Tensor, gamma: Optional[torch.Tensor],
               b: torch.Tensor, c: torch.Tensor, num_groups: int, transpose_a: bool = False,
               gamma_z: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (x [K,N] c64, y [K,G] c64). Differentiable w.r.t. a, gamma (scalar gains), b, c."""
    return _GFDNSolve.apply(z, delays, a, gamma, b, c, num_groups, transpose_a, gamma_z)


class _GFDNSolveFIR(torch.autograd.Function):
    """K1 with FIR coupling: x_k = (diag(z_k^m / gamma) - sum_p A_p z_k^-p)^-1 b (reference feedback_loop.py:362-373 with
    coupling_matrix_type filter_matrix). taps (N, N, P) real."""

    @staticmethod
    def forward(ctx, z, delays, taps, gamma, b, c, num_groups, transpose_a, gamma_z):
        z = _cuda("z", z, C128)
        delays = _cuda("delays", delays, torch.int32)
        if taps.dim() != 3 or taps.shape[0] != taps.shape[1]:
            raise RuntimeError("gfdn_solve_fir: taps must be (N, N, P)")
        taps_ = _cuda("taps", taps.permute(2, 0, 1), torch.float32)  # (P, N, N)
        gamma_ = _cuda("gamma", gamma, torch.float32, optional=True)
        gamma_z_ = _cuda("gamma_z", gamma_z, C64, optional=True)
        b_ = _cuda("b", b.reshape(-1), torch.float32)
        c_ = _cuda("c", c.reshape(-1), torch.float32)
        ntaps, n, _ = taps_.shape
        k = z.shape[0]
        if delays.numel() != n or b_.numel() != n or c_.numel() != n:
            raise RuntimeError("gfdn_solve_fir: inconsistent shapes")
        x = torch.empty(k, n, dtype=C64, device=z.device)
        y = torch.empty(k, num_groups, dtype=C64, device=z.device)
        with torch.cuda.device(z.device):
            _lib.call("dgfdn_solve_fir_fwd", n, num_groups, ntaps, k, _ptr(z), _ptr(delays), _ptr(taps_), int(transpose_a),
                      _ptr(gamma_), _ptr(gamma_z_), _ptr(b_), _ptr(c_), _ptr(x), _ptr(y), _stream())
        ctx.save_for_backward(z, delays, taps_, gamma_, gamma_z_, c_, x)
        ctx.meta = (n, num_groups, ntaps, k, int(transpose_a), b.shape, c.shape, taps.dtype)
        return x, y

    @staticmethod
    def backward(ctx, gx, gy):
        z, delays, taps_, gamma_, gamma_z_, c_, x = ctx.saved_tensors
        n, g, ntaps, k, tr, bshape, cshape, tdtype = ctx.meta
        if gx is None and gy is None:
            return (None, ) * 9
        gx_ = _cuda("gx", gx, C64, optional=True)
        gy_ = _cuda("gy", gy, C64, optional=True)
        dev = z.device
        out = torch.empty(3 * n, dtype=torch.float64, device=dev)
        gb, gc, gig = out[:n], out[n:2 * n], out[2 * n:]
        lam = torch.empty(k, n, dtype=C64, device=dev)
        with torch.cuda.device(dev):
            ws = torch.empty(_lib.load().dgfdn_solve_bwd_ws_bytes(n) // 8, dtype=torch.float64, device=dev)
            _lib.call("dgfdn_solve_fir_bwd", n, g, ntaps, k, _ptr(z), _ptr(delays), _ptr(taps_), tr, _ptr(gamma_),
                      _ptr(gamma_z_), _ptr(c_), _ptr(x), _ptr(gy_), _ptr(gx_), _ptr(lam), _ptr(gb), _ptr(gc), _ptr(gig),
                      _ptr(ws), _stream())
        g_taps = None
        if ctx.needs_input_grad[2]:
            # gtaps[i,j,p] = Re sum_k lambda[k,i] conj(x[k,j]) conj(z_k^-p): P small (N x K) x (K x N) products
            lam128, xc = lam.to(C128), x.to(C128).conj()
            zinv = (1.0 / z).conj()
            w = torch.ones_like(zinv)
            cols = []
            for _ in range(ntaps):
                cols.append(((lam128 * w.unsqueeze(-1)).transpose(0, 1) @ xc).real)
                w = w * zinv
            g_taps = torch.stack(cols, dim=-1)  # gradient w.r.t. A_eff: transposed back when the solve ran on A^T
            if tr:
                g_taps = g_taps.transpose(0, 1)
            g_taps = g_taps.to(tdtype)
        g_gamma = None
        if gamma_ is not None and ctx.needs_input_grad[3]:
            g_gamma = (-gig / gamma_.to(torch.float64)**2).to(torch.float32)
        g_b = gb.to(torch.float32).reshape(bshape) if ctx.needs_input_grad[4] else None
        g_c = gc.to(torch.float32).reshape(cshape) if ctx.needs_input_grad[5] else None
        return None, None, g_taps, g_gamma, g_b, g_c, None, None, None


def gfdn_solve_fir(z: torch.Tensor, delays: torch.Tensor, taps: torch.Tensor, gamma: Optional[torch.Tensor],
                   b: torch.Tensor, c: torch.Tensor, num_groups: int, transpose_a: bool = False,
                   gamma_z: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """gfdn_solve with a FIR feedback matrix A(z) = sum_p taps[..., p] z^-p; differentiable w.r.t. taps, gamma, b, c."""
    return _GFDNSolveFIR.apply(z, delays, taps, gamma, b, c, num_groups, transpose_a, gamma_z)


class _GFDNSolveGroups(torch.autograd.Function):
    """G independent LxL systems per bin: x[k, gL+i] = ((diag(z_k^{m_g}/gamma_g) - M_g)^-1 b_g)[i],
    y[k,g] = sum_i c[gL+i] x[k, gL+i]  -- DiffGFDN.sub_fdn_output (reference model.py:209-252)."""

    @staticmethod
    def forward(ctx, z, delays, m, gamma, b, c):
        z = _cuda("z", z, C128)
        delays = _cuda("delays", delays, torch.int32)
        m_ = _cuda("M", m, torch.float32)
        gamma_ = _cuda("gamma", gamma, torch.float32, optional=True)
        b_ = _cuda("b", b.reshape(-1), torch.float32)
        c_ = _cuda("c", c.reshape(-1), torch.float32)
        if m_.dim() != 3 or m_.shape[1] != m_.shape[2]:
            raise RuntimeError("gfdn_solve_groups: M must be (G, L, L)")
        g, l, _ = m_.shape
        n, k = g * l, z.shape[0]
        if delays.numel() != n or b_.numel() != n or c_.numel() != n or (gamma_ is not None and gamma_.numel() != n):
            raise RuntimeError("gfdn_solve_groups: inconsistent shapes")
        x = torch.empty(k, n, dtype=C64, device=z.device)
        y = torch.empty(k, g, dtype=C64, device=z.device)
        with torch.cuda.device(z.device):
            # the small LxL systems re-eliminate in the backward by default: saving L^2 multipliers per system costs
            # more HBM time than the O(L^3) elimination it would replace (measured: 0.42 vs 0.41 ms at L = 8)
            want = any(ctx.needs_input_grad) and os.environ.get("DGFDN_SOLVE_REPLAY_GROUPS", "0") == "1"
            factors = _factor_buffer(want, "dgfdn_solve_groups_factors_bytes", (l, g, k), z.device)
            _lib.call("dgfdn_solve_groups_fwd", l, g, k, _ptr(z), _ptr(delays), _ptr(m_), _ptr(gamma_), _ptr(b_),
                      _ptr(c_), _ptr(x), _ptr(y), _ptr(factors), _stream())
        ctx.factors = factors
        ctx.save_for_backward(z, delays, m_, gamma_, c_, x)
        ctx.meta = (g, l, k, b.shape, c.shape)
        return x, y

    @staticmethod
    def backward(ctx, gx, gy):
        z, delays, m_, gamma_, c_, x = ctx.saved_tensors
        g, l, k, bshape, cshape = ctx.meta
        if gx is None and gy is None:
            return (None, ) * 6
        gx_ = _cuda("gx", gx, C64, optional=True)
        gy_ = _cuda("gy", gy, C64, optional=True)
        n = g * l
        dev = z.device
        out = torch.empty(g * l * l + 3 * n, dtype=torch.float64, device=dev)
        gm, gb, gc, gig = out[:g * l * l], out[g * l * l:g * l * l + n], out[g * l * l + n:g * l * l + 2 * n], \
            out[g * l * l + 2 * n:]
        with torch.cuda.device(dev):
            ws = torch.empty(_lib.load().dgfdn_solve_groups_bwd_ws_bytes(l) // 8, dtype=torch.float64, device=dev)
            _lib.call("dgfdn_solve_groups_bwd", l, g, k, _ptr(z), _ptr(delays), _ptr(m_), _ptr(gamma_), _ptr(c_), _ptr(x),
                      _ptr(gy_), _ptr(gx_), _ptr(gm), _ptr(gb), _ptr(gc), _ptr(gig), _ptr(ws), _ptr(ctx.factors),
                      _stream())
        g_m = gm.reshape(g, l, l).to(torch.float32) if ctx.needs_input_grad[2] else None
        g_gamma = None
        if gamma_ is not None and ctx.needs_input_grad[3]:
            g_gamma = (-gig / gamma_.to(torch.float64)**2).to(torch.float32)
        g_b = gb.to(torch.float32).reshape(bshape) if ctx.needs_input_grad[4] else None
        g_c = gc.to(torch.float32).reshape(cshape) if ctx.needs_input_grad[5] else None
        return None, None, g_m, g_gamma, g_b, g_c


def gfdn_solve_groups(z: torch.Tensor, delays: torch.Tensor, m: torch.Tensor, gamma: Optional[torch.Tensor],
                      b: torch.Tensor, c: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (x [K, G*L] c64, y [K, G] c64) of the G decoupled systems. Differentiable w.r.t. m, gamma, b, c."""
    return _GFDNSolveGroups.apply(z, delays, m, gamma, b, c)


# ----------------------------------------------------------------------------------------------------------
# K2: receiver projection
# ----------------------------------------------------------------------------------------------------------
class _ReceiverProject(torch.autograd.Function):
    """H[r,k] = sum_g s[r,g] y[k,g] + d[r,k]  (reference model.py:583-619)."""

    @staticmethod
    def forward(ctx, s, y, d):
        s_ = _cuda("s", s, torch.float32)
        y_ = _cuda("y", y, C64)
        d_ = _cuda("d", d, C64, optional=True)
        rows, g = s_.shape
        k = y_.shape[0]
        if y_.shape[1] != g or (d_ is not None and tuple(d_.shape) != (rows, k)):
            raise RuntimeError("receiver_project: inconsistent shapes")
        h = torch.empty(rows, k, dtype=C64, device=s_.device)
        with torch.cuda.device(s_.device):
            _lib.call("dgfdn_project_fwd", g, rows, k, _ptr(s_), _ptr(y_), _ptr(d_), k, _ptr(h), k, _stream())
        ctx.save_for_backward(s_, y_)
        ctx.d_dtype = None if d is None else d.dtype
        return h

    @staticmethod
    def backward(ctx, gh):
        s_, y_ = ctx.saved_tensors
        rows, g = s_.shape
        k = y_.shape[0]
        gh_ = _cuda("gh", gh, C64)
        gs = torch.empty_like(s_) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y_) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(s_.device):
            _lib.call("dgfdn_project_bwd", g, rows, k, _ptr(s_), _ptr(y_), _ptr(gh_), k, _ptr(gy), 0, _ptr(gs),
                      _stream())
        gd = gh_.to(ctx.d_dtype) if (ctx.d_dtype is not None and ctx.needs_input_grad[2]) else None
        return gs, gy, gd


def receiver_project(s: torch.Tensor, y: torch.Tensor, d: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _ReceiverProject.apply(s, y, d)


class _SVFProject(torch.autograd.Function):
    """H[r,k] = sum_g F[r,g,k] y[k,g] + d[r,k], F the biquad-cascade response of coef[r,g] (reference
    gain_filters.py:221-241, 383-401; model.py:583-619 with use_svf_in_output)."""

    @staticmethod
    def forward(ctx, coef, z, y, d):
        coef_ = _cuda("coef", coef, torch.float64)
        z_ = _cuda("z", z, torch.complex128)
        y_ = _cuda("y", y, C64)
        d_ = _cuda("d", d, C64, optional=True)
        if coef_.dim() != 4 or coef_.shape[3] != 6:
            raise RuntimeError("svf_project: coef must be (rows, groups, sections, 6)")
        rows, g, nsec, _ = coef_.shape
        k = y_.shape[0]
        if y_.shape[1] != g or z_.numel() != k or (d_ is not None and tuple(d_.shape) != (rows, k)):
            raise RuntimeError("svf_project: inconsistent shapes")
        h = torch.empty(rows, k, dtype=C64, device=coef_.device)
        with torch.cuda.device(coef_.device):
            _lib.call("dgfdn_project_svf_fwd", g, nsec, rows, k, _ptr(coef_), _ptr(z_), _ptr(y_), _ptr(d_), k, _ptr(h), k,
                      _stream())
        ctx.save_for_backward(coef_, z_, y_)
        ctx.d_dtype = None if d is None else d.dtype
        ctx.coef_dtype = coef.dtype
        return h

    @staticmethod
    def backward(ctx, gh):
        coef_, z_, y_ = ctx.saved_tensors
        rows, g, nsec, _ = coef_.shape
        k = y_.shape[0]
        gh_ = _cuda("gh", gh, C64)
        gcoef = torch.empty_like(coef_) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y_) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(coef_.device):
            ws = None
            if gcoef is not None:
                nbytes = _lib.load().dgfdn_project_svf_bwd_ws_bytes(g, nsec, rows, k)
                ws = torch.empty(max(1, nbytes // 8), dtype=torch.float64, device=coef_.device)
            _lib.call("dgfdn_project_svf_bwd", g, nsec, rows, k, _ptr(coef_), _ptr(z_), _ptr(y_), _ptr(gh_), k,
                      _ptr(gcoef), _ptr(gy), _ptr(ws), _stream())
        gd = gh_.to(ctx.d_dtype) if (ctx.d_dtype is not None and ctx.needs_input_grad[3]) else None
        return (None if gcoef is None else gcoef.to(ctx.coef_dtype)), None, gy, gd


def svf_project(coef: torch.Tensor, z: torch.Tensor, y: torch.Tensor, d: Optional[torch.Tensor] = None) -> torch.Tensor:
    """coef (R,G,S,6) biquad coefficients (evaluated in float64), z (K,) complex128, y (K,G) complex64, d (R,K) complex64
    or None."""
    return _SVFProject.apply(coef, z, y, d)


class _SHProject(torch.autograd.Function):
    """H_sh[r,l,k] = sum_g cw[r,g,l] x[k, gL+l]  (reference model.py:1056-1088)."""

    @staticmethod
    def forward(ctx, cw, x):
        cw_ = _cuda("cw", cw, torch.float32)
        x_ = _cuda("x", x, C64)
        rows, g, l = cw_.shape
        k = x_.shape[0]
        if x_.shape[1] != g * l:
            raise RuntimeError("sh_project: inconsistent shapes")
        h = torch.empty(rows, l, k, dtype=C64, device=x_.device)
        with torch.cuda.device(x_.device):
            _lib.call("dgfdn_project_sh_fwd", g, l, rows, k, _ptr(cw_), _ptr(x_), _ptr(h), _stream())
        ctx.save_for_backward(cw_, x_)
        return h

    @staticmethod
    def backward(ctx, gh):
        cw_, x_ = ctx.saved_tensors
        rows, g, l = cw_.shape
        k = x_.shape[0]
        gh_ = _cuda("gh", gh, C64)
        gcw = torch.empty_like(cw_) if ctx.needs_input_grad[0] else None
        gx = torch.empty_like(x_) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x_.device):
            _lib.call("dgfdn_project_sh_bwd", g, l, rows, k, _ptr(cw_), _ptr(x_), _ptr(gh_), _ptr(gx), 0, _ptr(gcw),
                      _stream())
        return gcw, gx


def sh_project(cw: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return _SHProject.apply(cw, x)


class _MixChannels(torch.autograd.Function):
    """out[r,j,k] = sum_l w[j,l] in[r,l,k]  (reference trainer.py:853-865); w is a constant matrix."""

    @staticmethod
    def forward(ctx, w, x):
        w_ = _cuda("w", w, torch.float32)
        x_ = _cuda("x", x, C64)
        rows, cin, k = x_.shape
        cout = w_.shape[0]
        if w_.shape[1] != cin:
            raise RuntimeError("mix_channels: inconsistent shapes")
        out = torch.empty(rows, cout, k, dtype=C64, device=x_.device)
        with torch.cuda.device(x_.device):
            for r0 in range(0, rows, 32768):
                r1 = min(rows, r0 + 32768)
                _lib.call("dgfdn_mix_channels", cin, cout, r1 - r0, k, _ptr(w_), _ptr(x_[r0:r1]), _ptr(out[r0:r1]),
                          _stream())
        ctx.save_for_backward(w_)
        return out

    @staticmethod
    def backward(ctx, gout):
        (w_, ) = ctx.saved_tensors
        gx = _MixChannels.apply(w_.t().contiguous(), gout) if ctx.needs_input_grad[1] else None
        return None, gx


def mix_channels(w: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    return _MixChannels.apply(w, x)


# ----------------------------------------------------------------------------------------------------------
# K3a: windowed inverse real DFT of arbitrary length (chirp-z over cuFFT)
# ----------------------------------------------------------------------------------------------------------
class CZTPlan:
    """Device-side plan for out[t] = irfft(X, n)[t0 + t], t in [0, tn). Owns the chirp tables and cuFFT plans."""

    def __init__(self, n: int, t0: int, tn: int, device: torch.device):
        self.n, self.t0, self.tn, self.device = int(n), int(t0), int(tn), torch.device(device)
        self.num_bins = self.n // 2 + 1
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.call("dgfdn_czt_plan_create", self.n, self.t0, self.tn, ctypes.byref(h))
        self._h = h
        self.mc = int(_lib.load().dgfdn_czt_plan_mc(h))

    @property
    def handle(self):
        return self._h

    def rows_per_call(self, scratch_bytes: int = 1 << 30) -> int:
        return int(max(1, min(32768, scratch_bytes // (self.mc * 8))))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().dgfdn_czt_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


_PLANS = {}


def get_czt_plan(n: int, t0: int, tn: int, device) -> CZTPlan:
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (int(n), int(t0), int(tn), device.index)
    plan = _PLANS.get(key)
    if plan is None:
        plan = CZTPlan(n, t0, tn, device)
        _PLANS[key] = plan
    return plan


class _IrfftWindow(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, plan, filt):
        x_ = _cuda("X", x, C64)
        filt_ = _cuda("filt", filt, C64, optional=True)
        kx = x_.shape[-1]
        if kx < plan.num_bins:
            raise RuntimeError(f"irfft_window: need at least {plan.num_bins} bins, got {kx}")
        if filt_ is not None and filt_.numel() < plan.num_bins:
            raise RuntimeError("irfft_window: filter shorter than the number of bins used")
        lead = x_.shape[:-1]
        x2 = x_.reshape(-1, kx)
        rows = x2.shape[0]
        out = torch.empty(rows, plan.tn, dtype=torch.float32, device=x_.device)
        step = plan.rows_per_call()
        with torch.cuda.device(x_.device):
            scratch = torch.empty(min(step, rows) * plan.mc, dtype=C64, device=x_.device)
            for r0 in range(0, rows, step):
                r1 = min(rows, r0 + step)
                _lib.call("dgfdn_irfft_window_fwd", plan.handle, _ptr(x2[r0:r1]), kx, r1 - r0, _ptr(filt_),
                          _ptr(scratch), _ptr(out[r0:r1]), _stream())
        ctx.plan = plan
        ctx.kx = kx
        ctx.lead = lead
        ctx.save_for_backward(filt_)
        return out.reshape(*lead, plan.tn)

    @staticmethod
    def backward(ctx, gout):
        plan, kx = ctx.plan, ctx.kx
        (filt_, ) = ctx.saved_tensors
        g2 = _cuda("gout", gout, torch.float32).reshape(-1, plan.tn)
        rows = g2.shape[0]
        gx = torch.empty(rows, kx, dtype=C64, device=g2.device)
        step = plan.rows_per_call()
        with torch.cuda.device(g2.device):
            scratch = torch.empty(min(step, rows) * plan.mc, dtype=C64, device=g2.device)
            for r0 in range(0, rows, step):
                r1 = min(rows, r0 + step)
                _lib.call("dgfdn_irfft_window_bwd", plan.handle, _ptr(g2[r0:r1]), r1 - r0, _ptr(filt_), _ptr(scratch),
                          _ptr(gx[r0:r1]), kx, kx, _stream())
        return gx.reshape(*ctx.lead, kx), None, None


def irfft_window(x: torch.Tensor, n: int, t0: int, tn: int, filt: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.fft.irfft(filt * x, n)[..., t0:t0+tn] as float32, differentiable w.r.t. x (complex64 rows)."""
    plan = get_czt_plan(n, t0, tn, x.device)
    return _IrfftWindow.apply(x, plan, filt)


# ----------------------------------------------------------------------------------------------------------
# K3b: energy decay curves and the dB loss
# ----------------------------------------------------------------------------------------------------------
def edc_db(h: torch.Tensor) -> torch.Tensor:
    """10 log10(reverse-cumsum(h^2) + eps) clipped at -200 (reference losses.py:187-199 + utils.py:16-40). No grad."""
    h_ = _cuda("h", h, torch.float32)
    tn = h_.shape[-1]
    h2 = h_.reshape(-1, tn)
    out = torch.empty_like(h2)
    with torch.cuda.device(h_.device):
        _lib.call("dgfdn_edc_db", _ptr(h2), h2.shape[0], tn, _ptr(out), _stream())
    return out.reshape(h_.shape)


class _EDCLoss(torch.autograd.Function):
    """sum_{r,t} mask[t] |target_db[r,t] - dB(EDC(h)[r,t])| as a float64 scalar."""

    @staticmethod
    def forward(ctx, h, target_db, mask):
        h_ = _cuda("h", h, torch.float32)
        t_ = _cuda("target_db", target_db, torch.float32)
        m_ = _cuda("mask", mask, torch.float32, optional=True)
        tn = h_.shape[-1]
        h2 = h_.reshape(-1, tn)
        t2 = t_.reshape(-1, tn)
        if t2.shape != h2.shape or (m_ is not None and m_.numel() != tn):
            raise RuntimeError("edc_loss: inconsistent shapes")
        rows = h2.shape[0]
        row_sum = torch.empty(rows, dtype=torch.float64, device=h_.device)
        with torch.cuda.device(h_.device):
            _lib.call("dgfdn_edc_loss_fwd", _ptr(h2), _ptr(t2), _ptr(m_), rows, tn, _ptr(row_sum), _stream())
        ctx.save_for_backward(h2, t2, m_)
        ctx.shape = h_.shape
        return row_sum.sum()

    @staticmethod
    def backward(ctx, g):
        h2, t2, m_ = ctx.saved_tensors
        rows, tn = h2.shape
        gh = torch.empty_like(h2)
        with torch.cuda.device(h2.device):
            # the upstream scalar is folded in afterwards to avoid a device->host sync here
            _lib.call("dgfdn_edc_loss_bwd", _ptr(h2), _ptr(t2), _ptr(m_), rows, tn, 1.0, _ptr(gh), _stream())
        return (gh * g.to(torch.float32)).reshape(ctx.shape), None, None


def edc_abs_db_sum(h: torch.Tensor, target_db: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _EDCLoss.apply(h, target_db, mask)


# ----------------------------------------------------------------------------------------------------------
# K3c: receiver step in the time domain (mix + EDC + dB loss + backward in one kernel)
# ----------------------------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------------------------
# K3e: energy decay relief on an STFT
# ----------------------------------------------------------------------------------------------------------
def edr_db(s: torch.Tensor) -> torch.Tensor:
    """10 log10(sum_{m' >= m} |S[..., m', f]|^2 + eps) of an STFT S (R, T_f, F) complex64 -> (R, T_f, F) float32."""
    s_ = _cuda("S", s, C64)
    rows, tf, f = s_.shape
    out = torch.empty(rows, tf, f, dtype=torch.float32, device=s_.device)
    with torch.cuda.device(s_.device):
        _lib.call("dgfdn_edr_db", rows, tf, f, _ptr(s_), _ptr(out), _stream())
    return out


class _EDRLoss(torch.autograd.Function):
    """sum_b sum_{f,m} |T_dB - EDR_dB(S_b)| / den[b]  (reference losses.py:447-495)."""

    @staticmethod
    def forward(ctx, s, target_db, den):
        s_ = _cuda("S", s, C64)
        t_ = _cuda("target_db", target_db, torch.float32)
        den_ = _cuda("den", den, torch.float64)
        rows, tf, f = s_.shape
        if tuple(t_.shape) != (rows, tf, f) or den_.numel() != rows:
            raise RuntimeError("edr_loss: inconsistent shapes")
        loss = torch.empty(1, dtype=torch.float64, device=s_.device)
        with torch.cuda.device(s_.device):
            ws = torch.empty(max(1, _lib.load().dgfdn_edr_ws_bytes(rows, f) // 8), dtype=torch.float64, device=s_.device)
            _lib.call("dgfdn_edr_loss_fwd", rows, tf, f, _ptr(s_), _ptr(t_), _ptr(den_), _ptr(loss), _ptr(ws), _stream())
        ctx.save_for_backward(s_, t_, den_)
        return loss[0]

    @staticmethod
    def backward(ctx, gloss):
        s_, t_, den_ = ctx.saved_tensors
        rows, tf, f = s_.shape
        g_ = _cuda("gloss", gloss.reshape(1), torch.float64)
        gs = torch.empty_like(s_)
        with torch.cuda.device(s_.device):
            _lib.call("dgfdn_edr_loss_bwd", rows, tf, f, _ptr(s_), _ptr(t_), _ptr(den_), _ptr(g_), _ptr(gs), _stream())
        return gs, None, None


def edr_l1_normalised(s: torch.Tensor, target_db: torch.Tensor, den: torch.Tensor) -> torch.Tensor:
    return _EDRLoss.apply(s, target_db, den)


def td_contract_workspace(num_groups: int, rows: int, tn: int, device) -> torch.Tensor:
    nbytes = int(_lib.load().dgfdn_td_contract_ws_bytes(num_groups, rows, tn))
    return torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=device)


@torch.no_grad()
def td_mix(s: torch.Tensor, hy: torch.Tensor, hd: Optional[torch.Tensor] = None) -> torch.Tensor:
    """h[r,t] = sum_g s[r,g] hy[g,t] + hd[r,t]: late RIR window of every receiver (float32, no grad)."""
    s_ = _cuda("s", s, torch.float32)
    hy_ = _cuda("hy", hy, torch.float32)
    hd_ = _cuda("hd", hd, torch.float32, optional=True)
    rows, g = s_.shape
    tn = hy_.shape[-1]
    if hy_.shape[0] != g or (hd_ is not None and tuple(hd_.shape) != (rows, tn)):
        raise RuntimeError("td_mix: inconsistent shapes")
    out = torch.empty(rows, tn, dtype=torch.float32, device=s_.device)
    with torch.cuda.device(s_.device):
        for r0 in range(0, rows, 32768):
            r1 = min(rows, r0 + 32768)
            _lib.call("dgfdn_td_mix", g, r1 - r0, tn, _ptr(s_[r0:r1]), _ptr(hy_), _ptr(None if hd_ is None else hd_[r0:r1]),
                      tn, _ptr(out[r0:r1]), tn, _stream())
    return out


class _TDEDCLoss(torch.autograd.Function):
    """sum_{r,t} mask[t] |target_db[r,t] - dB(EDC(sum_g s[r,g] hy[g,:] + hd[r,:])[t])| as a float64 scalar.

    Same value as edc_abs_db_sum(irfft_window(receiver_project(s, y, d))) of the frequency-domain path (reference
    model.py:583-619 + losses.py:207-238), by linearity of the inverse DFT; differentiable w.r.t. s and hy."""

    @staticmethod
    def forward(ctx, s, hy, hd, target_db, mask, tile_rows):
        s_ = _cuda("s", s, torch.float32)
        hy_ = _cuda("hy", hy, torch.float32)
        hd_ = _cuda("hd", hd, torch.float32, optional=True)
        t_ = _cuda("target_db", target_db, torch.float32)
        m_ = _cuda("mask", mask, torch.float32, optional=True)
        rows, g = s_.shape
        tn = hy_.shape[-1]
        if hy_.shape[0] != g or tuple(t_.shape) != (rows, tn) or (hd_ is not None and tuple(hd_.shape) != (rows, tn)) \
                or (m_ is not None and m_.numel() != tn):
            raise RuntimeError("td_edc_loss: inconsistent shapes")
        dev = s_.device
        tile = max(1, min(int(tile_rows), rows))
        row_sum = torch.empty(rows, dtype=torch.float64, device=dev)
        gs = torch.empty(rows, g, dtype=torch.float32, device=dev)
        ghy = torch.empty(g, tn, dtype=torch.float32, device=dev)
        gh = torch.empty(tile, tn, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = td_contract_workspace(g, tile, tn, dev)
            for i, r0 in enumerate(range(0, rows, tile)):
                r1 = min(rows, r0 + tile)
                _lib.call("dgfdn_td_edc_step", g, r1 - r0, tn, _ptr(s_[r0:r1]), _ptr(hy_),
                          _ptr(None if hd_ is None else hd_[r0:r1]), tn, _ptr(t_[r0:r1]), tn, _ptr(m_), 1.0,
                          _ptr(row_sum[r0:r1]), _ptr(gs[r0:r1]), _ptr(gh), tn, _stream())
                _lib.call("dgfdn_td_contract", g, r1 - r0, tn, _ptr(s_[r0:r1]), _ptr(gh), tn, _ptr(ghy), int(i > 0),
                          _ptr(ws), _stream())
        ctx.save_for_backward(gs, ghy)
        return row_sum.sum()

    @staticmethod
    def backward(ctx, g):
        gs, ghy = ctx.saved_tensors
        gf = g.to(torch.float32)
        return (gs * gf if ctx.needs_input_grad[0] else None, ghy * gf if ctx.needs_input_grad[1] else None, None, None,
                None, None)


def td_edc_abs_db_sum(s: torch.Tensor, hy: torch.Tensor, hd: Optional[torch.Tensor], target_db: torch.Tensor,
                      mask: Optional[torch.Tensor] = None, tile_rows: int = 296) -> torch.Tensor:
    return _TDEDCLoss.apply(s, hy, hd, target_db, mask, tile_rows)


def td_fused_supported(num_groups: int, tn: int) -> bool:
    """True when the cluster-fused receiver kernel (K3d, dgfdn_td_edc_fused) handles this shape."""
    return bool(_lib.load().dgfdn_td_edc_fused_supported(int(num_groups), int(tn)))


def td_fused_workspace(num_groups: int, rows: int, tn: int, device) -> torch.Tensor:
    """Scratch of dgfdn_td_edc_fused for launches of up to `rows` rows, initialised (K3t's carry words read 'not
    published')."""
    nbytes = int(_lib.load().dgfdn_td_edc_fused_ws_bytes(num_groups, rows, tn))
    ws = torch.empty(max((nbytes + 3) // 4, 1), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.call("dgfdn_td_edc_fused_ws_init", _ptr(ws), int(num_groups), int(rows), int(tn), _stream())
    return ws


class _TDEDCLossFused(torch.autograd.Function):
    """Same value and gradients as _TDEDCLoss through dgfdn_td_edc_fused: one launch over all rows, dL/dh never
    stored (clusters of 8 CTAs per row, TMA-staged inputs, register-resident ghy accumulators)."""

    @staticmethod
    def forward(ctx, s, hy, hd, target_db, mask):
        s_ = _cuda("s", s, torch.float32)
        hy_ = _cuda("hy", hy, torch.float32)
        hd_ = _cuda("hd", hd, torch.float32, optional=True)
        t_ = _cuda("target_db", target_db, torch.float32)
        m_ = _cuda("mask", mask, torch.float32, optional=True)
        rows, g = s_.shape
        tn = hy_.shape[-1]
        if hy_.shape[0] != g or tuple(t_.shape) != (rows, tn) or (hd_ is not None and tuple(hd_.shape) != (rows, tn)) \
                or (m_ is not None and m_.numel() != tn):
            raise RuntimeError("td_edc_loss_fused: inconsistent shapes")
        dev = s_.device
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        gs = torch.empty(rows, g, dtype=torch.float32, device=dev)
        ghy = torch.empty(g, tn, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = td_fused_workspace(g, rows, tn, dev)
            _lib.call("dgfdn_td_edc_fused", g, rows, tn, _ptr(s_), _ptr(hy_), _ptr(hd_), tn, _ptr(t_), tn, _ptr(m_), 1.0,
                      _ptr(loss), _ptr(gs), _ptr(ghy), 0, _ptr(ws), _stream())
        ctx.save_for_backward(gs, ghy)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        gs, ghy = ctx.saved_tensors
        gf = g.to(torch.float32)
        return (gs * gf if ctx.needs_input_grad[0] else None, ghy * gf if ctx.needs_input_grad[1] else None, None, None,
                None)


def td_edc_abs_db_sum_fused(s: torch.Tensor, hy: torch.Tensor, hd: Optional[torch.Tensor], target_db: torch.Tensor,
                            mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _TDEDCLossFused.apply(s, hy, hd, target_db, mask)


# ----------------------------------------------------------------------------------------------------------
# K7: position -> gain network (encoding + MLP + final activation), one launch forward, two backward
# ----------------------------------------------------------------------------------------------------------
def mlp_supported(in_dim: int, nfeat: int, neurons: int, num_ln_layers: int, out_dim: int) -> bool:
    return bool(_lib.load().dgfdn_mlp_supported(int(in_dim), int(nfeat), int(neurons), int(num_ln_layers), int(out_dim)))


class _PositionMLP(torch.autograd.Function):
    """out = final_act(MLP(enc(pos))) for the reference's MLP / MLP_SkipConnections (dnn.py:284-400) with the
    sinusoidal encoding (dnn.py:89-126) and the scaled sigmoid (dnn.py:21-36). params = (W_l, b_l, ln_w_l, ln_b_l)
    per LayerNorm layer, then (W_out, b_out). Differentiable w.r.t. the parameters (positions are data)."""

    @staticmethod
    def forward(ctx, pos, freq, residual, final_act, lo, hi, *params):
        nl = (len(params) - 2) // 4
        pos_ = _cuda("pos", pos)
        if pos_.dtype not in (torch.float32, torch.float64):
            pos_ = pos_.to(torch.float32)
        freq_ = _cuda("freq", freq, pos_.dtype)
        ps = [_cuda("param", q, torch.float32) for q in params]
        rows = pos_.shape[0]
        neurons, in_dim = ps[0].shape
        out_dim = ps[-2].shape[0]
        nfeat = freq_.numel()
        dev = pos_.device
        ptrs = (ctypes.c_void_p * len(ps))(*[q.data_ptr() for q in ps])
        out = torch.empty(rows, out_dim, dtype=torch.float32, device=dev)
        xhat = torch.empty(nl, rows, neurons, dtype=torch.float32, device=dev)
        rstd = torch.empty(nl, rows, dtype=torch.float32, device=dev)
        asave = torch.empty(nl, rows, neurons, dtype=torch.float32, device=dev) if residual else None
        meta = (rows, in_dim, nfeat, neurons, nl, out_dim, int(residual), int(final_act), float(lo), float(hi),
                int(pos_.dtype == torch.float64))
        with torch.cuda.device(dev):
            _lib.call("dgfdn_mlp_fwd", *meta, _ptr(pos_), _ptr(freq_), ptrs, _ptr(out), _ptr(xhat), _ptr(rstd),
                      _ptr(asave), _stream())
        ctx.meta = meta
        ctx.shapes = [q.shape for q in params]
        ctx.save_for_backward(pos_, freq_, out, xhat, rstd, *( [asave] if residual else []), *ps)
        return out

    @staticmethod
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        rows, in_dim, nfeat, neurons, nl, out_dim, residual = ctx.meta[:7]
        pos_, freq_, out, xhat, rstd = saved[:5]
        asave = saved[5] if residual else None
        ps = saved[6 if residual else 5:]
        gout_ = _cuda("gout", gout, torch.float32)
        dev = pos_.device
        lib = _lib.load()
        nparams = int(lib.dgfdn_mlp_num_params(in_dim, neurons, nl, out_dim))
        grad = torch.empty(nparams, dtype=torch.float32, device=dev)
        ptrs = (ctypes.c_void_p * len(ps))(*[q.data_ptr() for q in ps])
        with torch.cuda.device(dev):
            ws = torch.empty(max(1, int(lib.dgfdn_mlp_bwd_ws_bytes(rows, in_dim, neurons, nl, out_dim)) // 4),
                             dtype=torch.float32, device=dev)
            _lib.call("dgfdn_mlp_bwd", *ctx.meta, _ptr(pos_), _ptr(freq_), ptrs, _ptr(out), _ptr(xhat), _ptr(rstd),
                      _ptr(asave), _ptr(gout_), _ptr(grad), _ptr(ws), _stream())
        grads, off = [], 0
        for i, shp in enumerate(ctx.shapes):
            n = math.prod(shp)
            grads.append(grad[off:off + n].view(shp) if ctx.needs_input_grad[6 + i] else None)
            off += n
        return (None, None, None, None, None, None, *grads)


def position_mlp(pos: torch.Tensor, freq: torch.Tensor, params, residual: bool = False, final_act: int = 0,
                 lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
    """(rows, out_dim) float32. params: flat list [W_0, b_0, ln_w_0, ln_b_0, ..., W_out, b_out]."""
    return _PositionMLP.apply(pos, freq, bool(residual), int(final_act), float(lo), float(hi), *params)


# ----------------------------------------------------------------------------------------------------------
# colorless loss
# ----------------------------------------------------------------------------------------------------------
class _Colorless(torch.autograd.Function):
    """loss[g] = mean_k (|H[k,g]| - 1)^p  (reference colorless_fdn/losses.py:20-73 with y_true = 1)."""

    @staticmethod
    def forward(ctx, h_sub, asym):
        h_ = _cuda("h_sub", h_sub, C64)
        k, g = h_.shape
        loss = torch.empty(g, dtype=torch.float64, device=h_.device)
        with torch.cuda.device(h_.device):
            _lib.call("dgfdn_colorless_fwd", g, k, _ptr(h_), int(asym), _ptr(loss), _stream())
        ctx.save_for_backward(h_)
        ctx.asym = int(asym)
        return loss

    @staticmethod
    def backward(ctx, gl):
        (h_, ) = ctx.saved_tensors
        k, g = h_.shape
        coef = _cuda("coef", gl, torch.float64)
        gh = torch.empty_like(h_)
        with torch.cuda.device(h_.device):
            _lib.call("dgfdn_colorless_bwd", g, k, _ptr(h_), ctx.asym, _ptr(coef), _ptr(gh), _stream())
        return gh, None


class _ColorlessSolve(torch.autograd.Function):
    """loss[g] = mean_k (|c_g^T (diag(z_k^{m_g}) - M_g)^-1 b_g| - 1)^p and its parameter gradients in ONE kernel pass
    (K1c): sub_fdn_output + mse/amse loss + both backward passes of the module path (reference model.py:209-252,
    colorless_fdn/losses.py:20-73, trainer.py:298-303)."""

    @staticmethod
    def forward(ctx, z, delays, m, b, c, asym, max_sms=0):
        z = _cuda("z", z, C128)
        delays = _cuda("delays", delays, torch.int32)
        m_ = _cuda("M", m, torch.float32)
        b_ = _cuda("b", b.reshape(-1), torch.float32)
        c_ = _cuda("c", c.reshape(-1), torch.float32)
        g, l, _ = m_.shape
        n, k = g * l, z.shape[0]
        if delays.numel() != n or b_.numel() != n or c_.numel() != n:
            raise RuntimeError("colorless_solve_loss: inconsistent shapes")
        dev = z.device
        loss = torch.empty(g, dtype=torch.float64, device=dev)
        out = torch.empty(g * l * l + 2 * n, dtype=torch.float64, device=dev)
        gm, gb, gc = out[:g * l * l], out[g * l * l:g * l * l + n], out[g * l * l + n:]
        with torch.cuda.device(dev):
            ws = torch.empty(_lib.load().dgfdn_solve_colorless_ws_bytes(l) // 8, dtype=torch.float64, device=dev)
            _lib.call("dgfdn_solve_colorless", l, g, k, _ptr(z), _ptr(delays), _ptr(m_), None, _ptr(b_), _ptr(c_),
                      int(asym), int(max_sms), _ptr(loss), _ptr(gm), _ptr(gb), _ptr(gc), _ptr(ws), _stream())
        ctx.save_for_backward(gm.reshape(g, l, l), gb.reshape(g, l), gc.reshape(g, l))
        ctx.shapes = (b.shape, c.shape)
        return loss

    @staticmethod
    def backward(ctx, gl):
        gm, gb, gc = ctx.saved_tensors
        bshape, cshape = ctx.shapes
        gl = gl.to(torch.float64)
        g_m = (gm * gl.view(-1, 1, 1)).to(torch.float32) if ctx.needs_input_grad[2] else None
        g_b = (gb * gl.view(-1, 1)).to(torch.float32).reshape(bshape) if ctx.needs_input_grad[3] else None
        g_c = (gc * gl.view(-1, 1)).to(torch.float32).reshape(cshape) if ctx.needs_input_grad[4] else None
        return None, None, g_m, g_b, g_c, None, None


def colorless_solve_loss(z: torch.Tensor, delays: torch.Tensor, m: torch.Tensor, b: torch.Tensor, c: torch.Tensor,
                         asym: bool, max_sms: int = 0) -> torch.Tensor:
    """Per-group colorless loss (G,) float64 of the lossless sub-FDNs, differentiable w.r.t. m (G,L,L), b, c.
    max_sms > 0 bounds the grid to that many SMs' worth of resident blocks (see dgfdn_solve_colorless)."""
    return _ColorlessSolve.apply(z, delays, m, b, c, bool(asym), int(max_sms))


def colorless_loss_per_group(h_sub: torch.Tensor, asym: bool) -> torch.Tensor:
    return _Colorless.apply(h_sub, asym)


# ----------------------------------------------------------------------------------------------------------
# K6: renderer
# ----------------------------------------------------------------------------------------------------------
@torch.no_grad()
def render_groups(delays: torch.Tensor, a: torch.Tensor, gamma: torch.Tensor, b: torch.Tensor, c: torch.Tensor,
                  num_groups: int, num_samples: int, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Block-recursive FDN: returns q [bands, T, G] float32 (group signals, weighted by c). Inputs have a leading
    band axis: delays [bands,N] int32, a [bands,N,N], gamma/b/c [bands,N]."""
    a_ = _cuda("a", a, torch.float32)
    bands, n, _ = a_.shape
    delays_ = _cuda("delays", delays.reshape(bands, n), torch.int32)
    gamma_ = _cuda("gamma", gamma.reshape(bands, n), torch.float32)
    b_ = _cuda("b", b.reshape(bands, n), torch.float32)
    c_ = _cuda("c", c.reshape(bands, n), torch.float32)
    u_ = _cuda("u", u, torch.float32, optional=True)
    if u_ is not None and u_.numel() < num_samples:
        raise RuntimeError("render_groups: input signal shorter than num_samples")
    hist = torch.empty(bands, num_samples, n, dtype=torch.float32, device=a_.device)
    q = torch.empty(bands, num_samples, num_groups, dtype=torch.float32, device=a_.device)
    with torch.cuda.device(a_.device):
        _lib.call("dgfdn_render_groups", bands, n, num_groups, num_samples, _ptr(delays_), _ptr(a_), _ptr(gamma_),
                  _ptr(b_), _ptr(c_), _ptr(u_), _ptr(hist), _ptr(q), _stream())
    return q


@torch.no_grad()
def render_mix(s: torch.Tensor, traj: torch.Tensor, q: torch.Tensor, hop: int) -> torch.Tensor:
    """out[r,t] = sum_band sum_g s[band, traj[r, t//hop], g] q[band,t,g]; s [bands,P,G], traj [R, ceil(T/hop)] int32."""
    s_ = _cuda("s", s, torch.float32)
    q_ = _cuda("q", q, torch.float32)
    traj_ = _cuda("traj", traj, torch.int32)
    bands, t, g = q_.shape
    positions = s_.shape[1]
    listeners = traj_.shape[0]
    nhops = (t + hop - 1) // hop
    if traj_.shape[1] != nhops or s_.shape[0] != bands or s_.shape[2] != g:
        raise RuntimeError("render_mix: inconsistent shapes")
    out = torch.empty(listeners, t, dtype=torch.float32, device=q_.device)
    with torch.cuda.device(q_.device):
        for r0 in range(0, listeners, 32768):
            r1 = min(listeners, r0 + 32768)
            _lib.call("dgfdn_render_mix", bands, g, t, r1 - r0, positions, hop, _ptr(s_), _ptr(traj_[r0:r1]), _ptr(q_),
                      _ptr(out[r0:r1]), _stream())
    return out


def is_finite_scalar(x: float) -> bool:
    return not (math.isnan(x) or math.isinf(x))


def td_fused_info(num_groups: int, tn: int) -> dict:
    """Tile variant, cluster size, threads per CTA and co-resident clusters of dgfdn_td_edc_fused for this shape."""
    import ctypes
    v = [ctypes.c_int(0) for _ in range(4)]
    _lib.call("dgfdn_td_edc_fused_info", int(num_groups), int(tn), *[ctypes.byref(x) for x in v])
    return dict(variant=v[0].value, cluster_size=v[1].value, threads=v[2].value, clusters=v[3].value)
