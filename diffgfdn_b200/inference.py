"""Inference / auralisation chain of the reference on the B200 kernels (SURVEY f-4).

What the reference does with trained models (all on the CPU, per position):

  * `get_response`: h = irfft(H) per receiver (diff_gfdn/utils.py:149-179);
  * sub-band synthesis: every octave-band model's RIR is convolved with that band's FIR of the amplitude-preserving
    filterbank, `fftconvolve(h, fir, 'full')`, and the bands are added (run_subband_training_treble.py:316-358);
  * moving listener: the stimulus is cut into hops, hop k is convolved with the (recursively smoothed) RIR of the
    k-th position and the blocks are overlap-added with a linear cross-fade over the first `fade_len` samples of a
    block (sound_examples.py:163-226 `filter_overlap_add`, fades :118-127, stimulus tiling :149-161).

Here the tail is produced by the block-recursive time-domain renderer K6 (`ops.render_groups`: the FDN state is
receiver independent, a receiver enters only through its G gains per band), and everything after it uses that
linearity:

    h_p           = sum_c s[c, p] qf_c ,   qf_(band,g) = q_(band,g) * fir_band          (C = bands x G channels)
    block k of r  = sum_c S'[r,k,c] (u_k * qf_c) ,   S'[r,k] = alpha s[:, p_k] + (1 - alpha) S'[r,k-1]
    out[r, t]     = sum_k sum_c S'[r,k,c] W[k,c,t - k hop]

with W the cross-fade-weighted block responses (receiver independent: built once) -- so R moving listeners cost
ONE GEMM-shaped contraction per output hop, (R x nb C) x (nb C x hop) with nb = blocks a response spans, instead of
R x num_pos FFT convolutions. That contraction is a plain dense GEMM and runs on the tensor cores through cuBLAS as
three TF32 products of a hi/lo split of both operands (float32-grade accuracy: the 1e-5-of-peak bar); the FIR and
block convolutions are batched cuFFT transforms; the recursion and the static mix are this package's kernels."""
from typing import List, Optional

import torch

from . import ops


def _next_pow2(n: int) -> int:
    return 1 << max(0, (int(n) - 1).bit_length())


def _tf32_split(x: torch.Tensor):
    """x = hi + lo with hi exactly representable in TF32 (10 explicit mantissa bits): low 13 bits cleared."""
    hi = (x.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, x - hi


def matmul_3xtf32(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a @ b on the tensor cores with float32-grade accuracy: a_hi b_hi + a_hi b_lo + a_lo b_hi (the dropped a_lo b_lo
    term is 2^-22 relative). cuBLAS TF32 GEMMs (a library GEMM for a plain dense product)."""
    keep = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a_hi, a_lo = _tf32_split(a)
        b_hi, b_lo = _tf32_split(b)
        res = torch.matmul(a_hi, b_lo)
        res.addmm_(a_lo, b_hi)
        res.addmm_(a_hi, b_hi)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = keep
    if out is not None:
        out.copy_(res)
        return out
    return res


class GFDNAuraliser:
    """Late-reverberation renderer of `bands` trained (sub-band) Grouped FDNs.

    delays (bands, N) int, a (bands, N, N) coupled feedback matrices, gamma / b / c (bands, N) delay-line gains and
    input / output gains -- e.g. from `DiffGFDN.feedback_loop.coupled_feedback_matrix_real()`, `.delay_line_gains`,
    `input_gains`, `output_gains` of each band's model; band_firs (bands, L) the filterbank FIRs (None: no sub-band
    filtering, a single full-band model)."""

    def __init__(self, delays, a, gamma, b, c, num_groups: int, band_firs: Optional[torch.Tensor] = None, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GFDNAuraliser runs on a CUDA device (no CPU fallback)")
        f32 = dict(device=self.device, dtype=torch.float32)
        self.a = torch.as_tensor(a).to(**f32).contiguous()
        self.bands, self.n = self.a.shape[0], self.a.shape[1]
        self.delays = torch.as_tensor(delays).to(device=self.device, dtype=torch.int32).reshape(self.bands, self.n).contiguous()
        self.gamma = torch.as_tensor(gamma).to(**f32).reshape(self.bands, self.n).contiguous()
        self.b = torch.as_tensor(b).to(**f32).reshape(self.bands, self.n).contiguous()
        self.c = torch.as_tensor(c).to(**f32).reshape(self.bands, self.n).contiguous()
        self.g = int(num_groups)
        self.firs = None if band_firs is None else torch.as_tensor(band_firs).to(**f32).reshape(self.bands, -1).contiguous()

    @classmethod
    def from_models(cls, nets: List, band_firs: Optional[torch.Tensor] = None):
        """One trained DiffGFDN per band (scalar absorption), as the reference's sub-band inference loads them
        (run_subband_training_treble.py:283-305)."""
        with torch.no_grad():
            a = torch.stack([n.feedback_loop.coupled_feedback_matrix_real().float() for n in nets])
            gam = torch.stack([n.feedback_loop.delay_line_gains.float() for n in nets])
            b = torch.stack([n.input_gains.reshape(-1).float() for n in nets])
            c = torch.stack([n.output_gains.reshape(-1).float() for n in nets])
            dl = torch.stack([n.delays.to(torch.int32) for n in nets])
        return cls(dl, a, gam, b, c, nets[0].num_groups, band_firs, device=nets[0].device)

    # ---- receiver-independent part ---------------------------------------------------------------------
    @torch.no_grad()
    def group_responses(self, num_samples: int) -> torch.Tensor:
        """q (bands, T, G): impulse responses of the group outputs (K6 block recursion)."""
        return ops.render_groups(self.delays, self.a, self.gamma, self.b, self.c, self.g, int(num_samples))

    @torch.no_grad()
    def filtered_group_responses(self, num_samples: int) -> torch.Tensor:
        """qf (bands, T + L - 1, G): every band's group responses convolved ('full') with the band's FIR -- the
        sub-band filtering of run_subband_training_treble.py:316-321 moved in front of the receiver mix (both are
        linear, the FIR is the same for every receiver)."""
        q = self.group_responses(num_samples)
        if self.firs is None:
            return q
        lf = self.firs.shape[1]
        n = _next_pow2(num_samples + lf - 1)
        spec = torch.fft.rfft(q, n=n, dim=1) * torch.fft.rfft(self.firs, n=n, dim=1).unsqueeze(-1)
        return torch.fft.irfft(spec, n=n, dim=1)[:, :num_samples + lf - 1].contiguous()

    # ---- static receivers --------------------------------------------------------------------------------
    @torch.no_grad()
    def static_rirs(self, s: torch.Tensor, num_samples: int) -> torch.Tensor:
        """h (P, T + L - 1) = sum_band fftconvolve(h_band[p], fir_band) for P positions with gains s (bands, P, G)
        (run_subband_training_treble.py:316-358: filtered band RIRs summed per position)."""
        qf = self.filtered_group_responses(num_samples)
        p = s.shape[1]
        traj = torch.arange(p, device=self.device, dtype=torch.int32).unsqueeze(1)
        return ops.render_mix(s.to(self.device, torch.float32), traj, qf, qf.shape[1])

    # ---- moving listeners --------------------------------------------------------------------------------
    @torch.no_grad()
    def moving_listeners(self, stimulus: torch.Tensor, s: torch.Tensor, traj: torch.Tensor, hop: int, rir_len: int,
                         fade_len: int, alpha: float = 0.5, max_block_bytes: int = 1 << 31) -> torch.Tensor:
        """out (R, num_pos * hop): `filter_overlap_add` (sound_examples.py:163-226) for R listeners at once. stimulus
        (any length, tiled to num_pos * hop like :149-161), s (bands, P, G) position gains, traj (R, num_pos) position
        of listener r during hop k, rir_len the rendered tail length T (the RIRs are T + L - 1 samples long)."""
        dev = self.device
        s = s.to(dev, torch.float32)
        traj = traj.to(dev, torch.long)
        r_n, npos = traj.shape
        total = npos * hop
        c_n = self.bands * self.g
        qf = self.filtered_group_responses(rir_len)                      # (bands, Lh, G)
        lh = qf.shape[1]
        qf = qf.permute(0, 2, 1).reshape(c_n, lh)                        # channel c = band * G + g
        # stimulus tiled to the simulation length (float32 buffer, like the reference's), cut into hops
        stim = stimulus.to(dev, torch.float32).reshape(-1)
        ext = stim.repeat((total + stim.numel() - 1) // stim.numel())[:total].reshape(npos, hop)
        # ---- receiver-independent block responses v[k, c] = u_k * qf_c, truncated at the end of the simulation
        ylen = hop + lh - 1
        nf = _next_pow2(ylen)
        qf_spec = torch.fft.rfft(qf, n=nf)                               # (C, nf/2+1)
        # W rows are long enough for a block's response plus the cross-fade tails that ride on its gains
        nb = (ylen + fade_len + hop - 1) // hop + (fade_len + hop - 1) // hop + 1
        nw = nb * hop
        w = torch.zeros(npos, c_n, nw, device=dev, dtype=torch.float32)
        per = max(1, int(max_block_bytes // (c_n * (nf // 2 + 1) * 8)))
        lens = [min(ylen, total - k * hop) for k in range(npos)]
        for k0 in range(0, npos, per):
            k1 = min(npos, k0 + per)
            v = torch.fft.irfft(torch.fft.rfft(ext[k0:k1], n=nf).unsqueeze(1) * qf_spec.unsqueeze(0), n=nf)
            w[k0:k1, :, :ylen] = v[:, :, :ylen]
        for k in range(npos):
            if lens[k] < ylen:
                w[k, :, lens[k]:ylen] = 0.0
        # ---- cross-fade (sound_examples.py:204-223), expressed on the block responses: block k > 0 adds
        # head * fade_in over its first ol samples, plus prev_tail * fade_out, where prev_tail is made of the LAST
        # samples of what earlier blocks produced -- pieces that scale with THOSE blocks' gains, so they are added
        # to those blocks' rows of W at the offset of block k
        ramp = torch.linspace(-1.0, 1.0, fade_len, device=dev, dtype=torch.float64)
        f_in, f_out = (0.5 * (1 + ramp)).float(), (0.5 * (1 - ramp)).float()
        tail = []  # [(source block m, piece (C, fade_len))]: prev_tail = sum_m gains_m . piece_m
        for k in range(npos):
            ol = min(fade_len, lens[k])
            if k > 0:
                for m, piece in tail:
                    off = (k - m) * hop
                    w[m, :, off:off + ol] += piece[:, :ol] * f_out[:ol]
            raw = w[k, :, :lens[k]].clone() if lens[k] < fade_len else w[k, :, lens[k] - fade_len:lens[k]].clone()
            if k > 0:
                w[k, :, :ol] *= f_in[:ol]
            if lens[k] >= fade_len:      # :217-219
                tail = [(k, raw)]
            else:                        # :220-223: only the first len samples of prev_tail are replaced
                kept = []
                for m, piece in tail:
                    piece = piece.clone()
                    piece[:, :lens[k]] = 0.0
                    kept.append((m, piece))
                new = torch.zeros(c_n, fade_len, device=dev)
                new[:, :lens[k]] = raw
                tail = kept + [(k, new)]
        # ---- per-listener gains of every hop, recursively smoothed (:189-192)
        s_c = s.permute(1, 0, 2).reshape(s.shape[1], c_n)               # (P, C)
        gains = s_c[traj]                                               # (R, npos, C)
        for k in range(1, npos):
            gains[:, k] = alpha * gains[:, k] + (1.0 - alpha) * gains[:, k - 1]
        # ---- overlap-add as one GEMM per output hop: (R x nk C) @ (nk C x hop), 3 x TF32 on the tensor cores
        out = torch.empty(r_n, total, device=dev, dtype=torch.float32)
        g_hi, g_lo = _tf32_split(gains)
        keep = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            for j in range(npos):
                k0 = max(0, j - nb + 1)
                nk = j + 1 - k0
                # rows (k, c) of W at the offset of output hop j: element (kk, c, t') sits at
                # (k0 + kk) C nw + c nw + (j - k0 - kk) hop + t' -- one strided gather
                b_j = torch.as_strided(w, (nk, c_n, hop), (c_n * nw - hop, nw, 1),
                                       k0 * c_n * nw + (j - k0) * hop).reshape(nk * c_n, hop)
                b_hi, b_lo = _tf32_split(b_j)
                a_hi = g_hi[:, k0:j + 1].reshape(r_n, -1)
                a_lo = g_lo[:, k0:j + 1].reshape(r_n, -1)
                res = torch.mm(a_hi, b_lo)
                res.addmm_(a_lo, b_hi)
                res.addmm_(a_hi, b_hi)
                out[:, j * hop:(j + 1) * hop] = res
        finally:
            torch.backends.cuda.matmul.allow_tf32 = keep
        return out


def srir_to_brir(srirs: torch.Tensor, hrir_sh: torch.Tensor, rotations: torch.Tensor) -> torch.Tensor:
    """Spatial (ambisonic) RIRs -> binaural RIRs for a set of head orientations (reference sofa_parser.py:452-505
    `convert_srir_to_brir`). srirs (R, C, T) with C = (order + 1)^2 SH channels; hrir_sh (C, 2, Th): the SH representation
    of the HRIR set (the reference gets it from the SOFA file through spaudiopy); rotations (O, C, C): the real SH rotation
    matrices of the NEGATED head orientations (spaudiopy `sh_rotation_matrix`, sofa_parser.py:488-493). Both are inputs
    here like the other third-party filter designs. Returns (R, O, nfft, 2) with nfft = next power of two >= T.

        BRTF[r,o,f,e] = sum_n conj(HRTF_sh[n,e,f]) sum_m Rot[o,n,m] SRTF[r,m,f]

    The reference loops over receivers and orientations on the host; this is one batched contraction between two real
    FFTs on whatever device the inputs live on (cuFFT + a cuBLAS complex GEMM per batch: plain library work, no own kernel)."""
    if srirs.dim() != 3 or hrir_sh.dim() != 3 or rotations.dim() != 3:
        raise RuntimeError("srir_to_brir: srirs (R, C, T), hrir_sh (C, 2, Th), rotations (O, C, C)")
    c = srirs.shape[1]
    if hrir_sh.shape[0] != c or hrir_sh.shape[1] != 2 or tuple(rotations.shape[1:]) != (c, c):
        raise RuntimeError("srir_to_brir: inconsistent number of SH channels")
    nfft = _next_pow2(srirs.shape[-1])
    rtf = torch.fft.rfft(srirs, n=nfft, dim=-1)  # (R, C, F)
    htf = torch.fft.rfft(hrir_sh.to(srirs.device, srirs.dtype), n=nfft, dim=-1)  # (C, 2, F)
    rot = rotations.to(srirs.device, srirs.dtype)
    # fold the HRTFs into the (real) rotation first: W[o,e,m,f] = sum_n conj(H[n,e,f]) Rot[o,n,m] is receiver independent.
    # Real and imaginary parts are contracted as real tensors (plain GEMMs; no complex-einsum code path is involved).
    wr = torch.einsum('nef,onm->oemf', htf.real, rot)
    wi = -torch.einsum('nef,onm->oemf', htf.imag, rot)
    ar, ai = rtf.real, rtf.imag
    re = torch.einsum('oemf,rmf->rofe', wr, ar) - torch.einsum('oemf,rmf->rofe', wi, ai)
    im = torch.einsum('oemf,rmf->rofe', wr, ai) + torch.einsum('oemf,rmf->rofe', wi, ar)
    return torch.fft.irfft(torch.complex(re, im), n=nfft, dim=2)
