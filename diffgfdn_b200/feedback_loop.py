"""Feedback loop of the Grouped FDN: orthogonal coupled feedback matrix + the per-bin solve.

Mirrors reference diff_gfdn/feedback_loop.py (class and parameter names, state_dict keys `M`, `alpha`). The
parameter pre-processing (matrix exponential, Givens rotations, Kronecker mask) is O(N^2) and stays in PyTorch on
the device, differentiable for free; the O(K N^3) part -- the reference's `torch.linalg.inv` over K dense matrices
(feedback_loop.py:391) -- is the sm_100a kernel behind `ops.gfdn_solve`."""
import os
import warnings
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import ops
from .absorption_filters import decay_times_to_gain_per_sample
from .config.config import CouplingMatrixType


class Skew(nn.Module):

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        A = X.triu(1)
        return A - A.transpose(-1, -2)


class MatrixExponential(nn.Module):

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        return torch.matrix_exp(X)


class OrthoParam(nn.Sequential):
    """Skew() -> MatrixExponential(), the reference's `ortho_param` (feedback_loop.py:270). CUDA inputs with
    L <= 16 take the fused, host-sync-free kernel (ops.skew_expm); anything else the stock torch sequence."""

    def __init__(self):
        super().__init__(Skew(), MatrixExponential())

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if X.is_cuda and X.dtype == torch.float32 and X.shape[-1] <= ops.SKEW_EXPM_MAX_L:
            return ops.skew_expm(X)
        return super().forward(X)


class ND_Unitary(nn.Module):
    """N-D rotation from N(N-1)/2 Givens angles: U_n = R_{n-2}...R_0 [[U_{n-1},0],[0,1]] with R_i rotating the
    (i, n-1) plane (reference feedback_loop.py:39-87). Built without in-place writes, on alpha's device; the constant
    selector matrices are cached per (N, device, dtype) so a step issues no host->device copies (CUDA-graph safe)."""

    def __init__(self):
        super().__init__()
        self._consts = {}

    def _constants(self, n, like):
        """c0[i] + cos(a_i) c1[i] + sin(a_i) c2[i] is the rotation of the (i, n-1) plane; `corner` completes the
        embedding of U_{n-1} into n dimensions."""
        key = (n, like.device, like.dtype)
        c = self._consts.get(key)
        if c is None:
            eye = torch.eye(n)
            c0, c1, c2 = torch.zeros(n - 1, n, n), torch.zeros(n - 1, n, n), torch.zeros(n - 1, n, n)
            for i in range(n - 1):
                c1[i, i, i] = 1.0
                c1[i, n - 1, n - 1] = 1.0
                c0[i] = eye - c1[i]
                c2[i, n - 1, i] = 1.0
                c2[i, i, n - 1] = -1.0
            corner = torch.zeros(n, n)
            corner[n - 1, n - 1] = 1.0
            c = tuple(t.to(device=like.device, dtype=like.dtype) for t in (c0, c1, c2, corner))
            self._consts[key] = c
        return c

    def forward(self, alpha: torch.Tensor, N: int) -> torch.Tensor:
        assert len(alpha) == N * (N - 1) // 2
        if N == 1:
            return torch.ones(1, 1, dtype=alpha.dtype, device=alpha.device)
        start = (N - 1) * (N - 2) // 2
        cur = alpha[start:]
        c0, c1, c2, corner = self._constants(N, alpha)
        # every rotation of this level in four launches (the reference writes them entry by entry, :60-87)
        rots = c0 + torch.cos(cur).view(-1, 1, 1) * c1 + torch.sin(cur).view(-1, 1, 1) * c2
        rot = rots[0]
        for i in range(1, N - 1):
            rot = rots[i] @ rot
        big = nn.functional.pad(self.forward(alpha[:start], N - 1), (0, 1, 0, 1)) + corner
        return rot @ big


class FeedbackLoop(nn.Module):

    def __init__(self,
                 sample_rate: float,
                 num_groups: int,
                 num_delay_lines_per_group: int,
                 delays: torch.Tensor,
                 use_absorption_filters: bool,
                 coupling_matrix_type: CouplingMatrixType = None,
                 use_zero_coupling: bool = True,
                 coupling_matrix_order: Optional[int] = None,
                 colorless_feedback_matrix: Optional[torch.Tensor] = None,
                 gains: Optional[torch.Tensor] = None,
                 common_decay_times: Optional[List] = None,
                 device: torch.device = 'cpu'):
        super().__init__()
        self.sample_rate = sample_rate
        self.num_groups = num_groups
        self.num_delay_lines_per_group = num_delay_lines_per_group
        self.delays = torch.as_tensor(delays, dtype=torch.float32, device=device)
        self.num_delays = len(self.delays)
        self.use_absorption_filters = use_absorption_filters
        self.use_zero_coupling = use_zero_coupling
        self.device = device
        self.coupling_matrix_type = coupling_matrix_type or CouplingMatrixType.SCALAR
        self.coupling_matrix_order = coupling_matrix_order
        self._eps = 1e-9
        if self.coupling_matrix_type == CouplingMatrixType.FILTER and not coupling_matrix_order or \
                (self.coupling_matrix_type == CouplingMatrixType.FILTER and coupling_matrix_order < 2):
            raise ValueError("filter_matrix coupling needs coupling_matrix_order (pu_matrix_order) >= 2")
        self._init_absorption(gains, common_decay_times)
        self._init_feedback_matrix(colorless_feedback_matrix)

    # ---- absorption (reference feedback_loop.py:193-258) -------------------------------------------------
    def _init_absorption(self, gains, common_decay_times):
        self.delay_line_gain_response = None  # (N, K) complex per-bin response when filters are used
        if gains is None:
            if self.use_absorption_filters:
                raise NotImplementedError("learnable absorption filters do not exist in the reference either")
            if common_decay_times is None:
                t60 = 0.1 + 1.9 * torch.rand(self.num_groups)
            else:
                t60 = torch.as_tensor(np.asarray(common_decay_times).squeeze(), dtype=torch.float32)
            self.common_decay_times = nn.Parameter(t60.to(self.device))
            per = self.num_delay_lines_per_group
            # computed once at construction like the reference (quirk Q9: never refreshed). The reference keeps the graph
            # of this one-off computation alive (loss.backward(retain_graph=learn_common_decay_times)), so ITS optimizer
            # moves `common_decay_times` although no forward pass ever reads the new value; here the gains are detached:
            # identical outputs, but the parameter (and the checkpoint entry of that name) stays at its initial value.
            warnings.warn("FeedbackLoop: common_decay_times is a parameter without effect -- the delay-line gains are "
                          "computed once at construction (as in the reference) and detached; it receives no gradient",
                          stacklevel=2)
            self.delay_line_gains = torch.cat([
                decay_times_to_gain_per_sample(self.common_decay_times[i], self.delays[i * per:(i + 1) * per],
                                               torch.tensor(self.sample_rate)) for i in range(self.num_groups)
            ]).detach().to(self.device)
        elif self.use_absorption_filters:
            # Frequency-dependent absorption (reference :236-255): `gains` holds the filter coefficients, designed at
            # init time on the host (GEQ: (N, S, 3, 2) biquads [.., 0] numerators / [.., 1] denominators; Prony:
            # (N, order, 2)). Their responses Gamma_i(z_k) are evaluated on the z grid of the first solve and cached.
            g = torch.as_tensor(gains, device=self.device)
            if g.ndim not in (3, 4):
                raise RuntimeError("absorption filters: expected (N, S, 3, 2) biquads or (N, order, 2) IIR coefficients")
            self.absorption_coeffs = g
            self.delay_line_gains = g
            self._gamma_cache = None
        else:
            self.delay_line_gains = torch.as_tensor(gains, dtype=torch.float32, device=self.device)

    def set_absorption_response(self, gamma_z: torch.Tensor):
        """Use per-bin complex delay-line gains Gamma_i(z_k), shape (N, K) (reference feedback_loop.py:333-344)."""
        self.delay_line_gain_response = gamma_z.to(torch.complex64)

    @torch.no_grad()
    def absorption_response(self, z: torch.Tensor) -> Optional[torch.Tensor]:
        """Gamma_i(z_k) (N, K) complex64 of the absorption filters on this z grid, or None for scalar gains.
        GEQ biquads go through the cascade kernel (K2s, float64 inside; reference :334-340 multiplies the sections
        in complex64); Prony filters are sum_k b_k z^-k / (sum_k a_k z^-k + 1e-9) (reference gain_filters.py:176-203).
        Constant over training: cached per (grid, coefficient version)."""
        if self.delay_line_gain_response is not None:
            if self.delay_line_gain_response.shape[-1] != z.numel():
                raise RuntimeError("absorption response was set for a different number of bins")
            return self.delay_line_gain_response
        coeffs = getattr(self, "absorption_coeffs", None)
        if coeffs is None:
            return None
        key = (z.data_ptr(), z.numel(), coeffs.data_ptr(), coeffs._version)
        if self._gamma_cache is not None and self._gamma_cache[0] == key:
            return self._gamma_cache[1]
        if not bool(coeffs.abs().sum() > 0):
            raise RuntimeError("absorption filters are all zero: the GEQ design runs on the host at init time and is "
                               "outside this package -- load a reference checkpoint (buffer 'delay_filters') or call "
                               "DiffGFDN.set_absorption_filters(coefficients)")
        from .gain_filters import cascade_response
        zc = z.to(device=coeffs.device, dtype=torch.complex128)
        if coeffs.ndim == 4:
            coef = torch.cat([coeffs[..., 0], coeffs[..., 1]], dim=-1).to(torch.float64).unsqueeze(1)  # (N, 1, S, 6)
            gamma = cascade_response(coef, zc)[:, 0]
        else:
            k = torch.arange(coeffs.shape[1], device=coeffs.device, dtype=torch.float64)
            zp = zc.unsqueeze(0)**(-k.unsqueeze(-1))  # (order, K)
            num = coeffs[..., 0].to(torch.complex128) @ zp
            den = coeffs[..., 1].to(torch.complex128) @ zp
            gamma = (num / (den + 1e-9)).to(torch.complex64)
        self._gamma_cache = (key, gamma, z)  # z is kept alive: its address cannot be handed to another grid
        return gamma

    # ---- feedback matrix (reference feedback_loop.py:260-324) --------------------------------------------
    def _init_feedback_matrix(self, colorless_feedback_matrix):
        self.ortho_param = OrthoParam()
        L = self.num_delay_lines_per_group
        if self.coupling_matrix_type == CouplingMatrixType.RANDOM:
            # any orthogonal N x N matrix, no group structure (reference :272-277):  A = expm(skew(R))
            n = self.num_delays
            self.random_feedback_matrix = nn.Parameter(((2 * torch.rand(n, n) - 1) / np.sqrt(L)).to(self.device))
            return
        if colorless_feedback_matrix is not None:
            self.M = colorless_feedback_matrix.clone().detach().to(self.device)
        else:
            self.M = nn.Parameter(((2 * torch.rand(self.num_groups, L, L) - 1) / np.sqrt(L)).to(self.device))
        self.nd_unitary = ND_Unitary()
        if self.coupling_matrix_type == CouplingMatrixType.FILTER:
            # paraunitary FIR coupling Phi(z): order - 1 Householder vectors and the order-0 unitary factor
            # (reference feedback_loop.py:311-324; same parameter names and shapes)
            self.unit_vectors = nn.Parameter(torch.randn(self.num_groups, self.coupling_matrix_order - 1).to(self.device))
            self.unitary_matrix = nn.Parameter(((2 * torch.rand(self.num_groups, self.num_groups) - 1) /
                                                np.sqrt(self.num_groups)).to(self.device))
            return
        n_alpha = self.num_groups * (self.num_groups - 1) // 2
        if self.use_zero_coupling:
            self.register_buffer("alpha", torch.zeros(n_alpha, device=self.device))
        else:
            self.alpha = nn.Parameter((np.pi / 4 * torch.rand(n_alpha, dtype=torch.float32)).to(self.device))

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.delays = fn(self.delays)
        self.delay_line_gains = fn(self.delay_line_gains)
        if hasattr(self, "M") and not isinstance(self.M, nn.Parameter):
            self.M = fn(self.M)
        if self.delay_line_gain_response is not None:
            self.delay_line_gain_response = fn(self.delay_line_gain_response)
        if getattr(self, "absorption_coeffs", None) is not None:
            self.absorption_coeffs = self.delay_line_gains  # same tensor (moved above)
            self._gamma_cache = None
        return self

    def construct_block_mixing_matrix(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """block_M[i,j] = U_i U_j with U = expm(skew(M)); diagonal blocks are U_i^2 (reference :393-404, Q4)."""
        U = self.ortho_param(self.M)  # (G, L, L)
        if dtype is not None:
            U = U.to(dtype)
        G, L, _ = U.shape
        blocks = torch.einsum('iab,jbc->iajc', U, U)  # (G, L, G, L)
        return blocks.reshape(G * L, G * L)

    def construct_paraunitary_coupling(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Phi(z) = H_{P-2}(z) ... H_0(z) U as (G, G, P) real taps: H_k(z) = (I - v_k v_k^T) + v_k v_k^T z^-1 with v_k
        the k-th column of `unit_vectors` scaled to unit length, U = expm(skew(unitary_matrix)) (reference
        FIRParaunitary, feedback_loop.py:90-143, called at :414-421). The polynomial product is a handful of G x G
        matmuls (the reference convolves entry by entry, utils.py:216-239)."""
        dt = dtype or self.unit_vectors.dtype
        uv = self.unit_vectors.to(dt)
        v = uv / (torch.norm(uv, dim=0, keepdim=True) + self._eps)
        eye = torch.eye(self.num_groups, dtype=dt, device=uv.device)
        taps = [eye]
        for k in range(v.shape[1]):
            vv = torch.outer(v[:, k], v[:, k])
            h0 = eye - vv
            nxt = [h0 @ taps[0]] + [h0 @ taps[i] + vv @ taps[i - 1] for i in range(1, len(taps))] + [vv @ taps[-1]]
            taps = nxt
        u = self.ortho_param(self.unitary_matrix.to(dt))
        return torch.stack([t @ u for t in taps], dim=-1)

    def coupled_feedback_taps(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Taps of A(z) = sum_p A_p z^-p, (N, N, P) real: A_p = block_M o (Phi_p (x) 1_{LxL}) (reference :447-453)."""
        block_M = self.construct_block_mixing_matrix(dtype)
        phi = self.construct_paraunitary_coupling(dtype)
        self.phi = phi.detach()
        G, L = self.num_groups, self.num_delay_lines_per_group
        taps = block_M.view(G, L, G, L, 1) * phi.view(G, 1, G, 1, -1)
        return taps.reshape(G * L, G * L, -1)

    def _solve_filter_coupling(self, z: torch.Tensor, b: torch.Tensor, c: torch.Tensor, transpose: bool):
        """filter_matrix coupling: the system matrix D(z_k) Gamma^-1 - A(z_k) has a different complex A per bin
        (reference :362-373). K1 builds A(z_k) = sum_p A_p z_k^-p from the taps in shared memory per bin
        (ops.gfdn_solve_fir); the adjoint hands back lambda and the tap gradients are P small products."""
        taps = self.coupled_feedback_taps(torch.float64)
        self.coupled_feedback_matrix = taps.detach()
        gamma_z = self.absorption_response(z)
        gamma = None if gamma_z is not None else self.delay_line_gains
        return ops.gfdn_solve_fir(z, self.delays.to(torch.int32), taps, gamma, b, c, self.num_groups,
                                  transpose_a=transpose, gamma_z=gamma_z)

    def construct_coupling_matrix(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        if self.coupling_matrix_type == CouplingMatrixType.FILTER:
            return self.construct_paraunitary_coupling(dtype)
        alpha = self.alpha.clamp(min=-np.pi, max=np.pi)
        if dtype is not None:
            alpha = alpha.to(dtype)
        return self.nd_unitary(alpha, self.num_groups)

    def coupled_feedback_matrix_real(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        if self.coupling_matrix_type == CouplingMatrixType.RANDOM:
            a = self.ortho_param(self.random_feedback_matrix)
            return a if dtype is None else a.to(dtype)
        return self._structured_feedback_matrix(dtype)

    def _structured_feedback_matrix(self, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """A = block_M o (Phi (x) 1_{LxL}), real (N, N) float32 (reference :424-455 before to_complex).

        dtype=torch.float64 runs the small product / Givens graph in float64: dL/dalpha is a sum of O(1) terms of
        dL/dA that cancels to ~1e-4 of their size, which a float32 graph resolves to 1e-3 only (the K1 adjoint hands
        dL/dA over in float64; the forward value is rounded to float32 by K1 either way)."""
        block_M = self.construct_block_mixing_matrix(dtype)
        phi = self.construct_coupling_matrix(dtype)
        self.phi = phi.detach()  # kept for get_parameters()/get_param_dict(); detached (no graph outlives the step)
        G, L = self.num_groups, self.num_delay_lines_per_group
        # block_M o (Phi (x) 1_{LxL}) as one broadcast multiply over the (G, L, G, L) view
        return (block_M.view(G, L, G, L) * phi.view(G, 1, G, 1)).reshape(G * L, G * L)

    def _assembled(self) -> torch.Tensor:
        """A (N, N) float64 for the solves: one fused kernel forward, one backward (ops.coupled_feedback) instead of
        the ~135 tiny launches of the torch graph in coupled_feedback_matrix_real (which stays the public, float32
        form). DGFDN_FUSED_ASSEMBLY=0 falls back to the torch graph."""
        if self.coupling_matrix_type == CouplingMatrixType.RANDOM:
            a = self.ortho_param(self.random_feedback_matrix)
            self.coupled_feedback_matrix = a.detach()
            return a
        if os.environ.get("DGFDN_FUSED_ASSEMBLY", "1") == "0" or self.num_groups > 8:
            a = self.coupled_feedback_matrix_real(torch.float64)
        else:
            a, phi = ops.coupled_feedback(self.ortho_param(self.M), self.alpha)
            self.phi = phi
        self.coupled_feedback_matrix = a.detach()
        return a

    def get_coupled_feedback_matrix(self) -> torch.Tensor:
        if self.coupling_matrix_type == CouplingMatrixType.FILTER:  # (N, N, order) taps of A(z)
            a = self.coupled_feedback_taps()
        else:
            a = self.coupled_feedback_matrix_real()
        return torch.complex(a, torch.zeros_like(a))

    def solve(self, z: torch.Tensor, b: torch.Tensor, c: torch.Tensor, transpose: bool = False):
        """x_k = (D(z_k) Gamma^-1 - A)^-1 b and y[k,g] = sum_{n in g} c_n x_k[n] on the GPU (one warp per bin)."""
        if self.coupling_matrix_type == CouplingMatrixType.FILTER:
            return self._solve_filter_coupling(z, b, c, transpose)
        a = self._assembled()
        gamma_z = self.absorption_response(z)
        gamma = None if gamma_z is not None else self.delay_line_gains
        return ops.gfdn_solve(z, self.delays.to(torch.int32), a, gamma, b, c, self.num_groups, transpose_a=transpose,
                              gamma_z=gamma_z)

    def transfer_matrix(self, z: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> List[torch.Tensor]:
        """Group-to-group transfer functions T[g'][k,g] = sum_{n in g, m in g'} c_n P_k[n,m] b_m: one K1 solve per
        source group g' (b masked to that group), A assembled once. The model variants with source-side gains or
        filters (reference model.py:402-452, 779-836) are bilinear in the per-group factors of both sides, so these
        G x G functions per bin are all they need of the feedback loop."""
        L = self.num_delay_lines_per_group
        b = b.reshape(-1)
        group_of_line = torch.arange(self.num_delays, device=b.device) // L
        if self.coupling_matrix_type == CouplingMatrixType.FILTER:
            return [self._solve_filter_coupling(z, b * (group_of_line == gp).to(b.dtype), c, False)[1]
                    for gp in range(self.num_groups)]
        a = self._assembled()
        gamma_z = self.absorption_response(z)
        gamma = None if gamma_z is not None else self.delay_line_gains
        delays = self.delays.to(torch.int32)
        out = []
        for gp in range(self.num_groups):
            mask = (group_of_line == gp).to(b.dtype)
            out.append(ops.gfdn_solve(z, delays, a, gamma, b * mask, c, self.num_groups, gamma_z=gamma_z)[1])
        return out

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """Dense P[k] = (D Gamma^-1 - A)^-1, (K, N, N) complex64 -- API compatibility with reference :326-391.
        The models never call this (they solve against b directly); it runs N single-RHS solves."""
        n = self.num_delays
        eye = torch.eye(n, dtype=torch.float32, device=self.delays.device)
        cols = [self.solve(z, eye[j], eye[j])[0] for j in range(n)]
        return torch.stack(cols, dim=-1)

    def get_parameters(self):
        if self.coupling_matrix_type == CouplingMatrixType.RANDOM:  # reference :465-466
            return self.ortho_param(self.random_feedback_matrix)
        M = [self.ortho_param(self.M[i]) for i in range(self.num_groups)]
        coupled = self.get_coupled_feedback_matrix()
        return (M, self.phi, None, None, coupled, self.delay_line_gains)

    @torch.no_grad()
    def get_param_dict(self):
        coupled = self.get_coupled_feedback_matrix()
        if self.coupling_matrix_type == CouplingMatrixType.RANDOM:  # reference :490-494
            d = {'delay_line_gains': self.delay_line_gains, 'coupled_feedback_matrix': coupled.squeeze().cpu().numpy()}
            if hasattr(self, 'common_decay_times'):
                d['common_decay_times'] = self.common_decay_times
            return d
        d = {'delay_line_gains': self.delay_line_gains, 'coupling_matrix': self.phi.squeeze().cpu().numpy(),
             'individual_mixing_matrix': self.M.squeeze().cpu().numpy(),
             'coupled_feedback_matrix': coupled.squeeze().cpu().numpy()}
        if hasattr(self, 'common_decay_times'):
            d['common_decay_times'] = self.common_decay_times
        if self.coupling_matrix_type == CouplingMatrixType.FILTER:  # reference :481-488
            uv = self.unit_vectors / (torch.norm(self.unit_vectors, dim=0, keepdim=True) + self._eps)
            d['unitary_matrix'] = self.ortho_param(self.unitary_matrix).squeeze().cpu().numpy()
            d['unit_vectors'] = uv.squeeze().cpu().numpy()
            return d
        if not self.use_zero_coupling:
            d['coupling_coefficient'] = self.alpha.squeeze().cpu().numpy()
        return d
