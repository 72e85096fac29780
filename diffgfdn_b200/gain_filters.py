"""Receiver/source gain and filter generators (reference diff_gfdn/gain_filters.py).

`Gains_from_MLP` maps a (normalised) position to one scalar gain per group. The reference then repeats those gains
to a dense (B, N, K) tensor (gain_filters.py:526-536) which the model multiplies bin by bin; here the kernels
consume the (B, G) table directly (`gains()`), and `forward()` returns a stride-0 *view* with the reference's shape
for callers that want it."""
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .config.config import FeatureEncodingType
from .dnn import MLP, ScaledSigmoid, SinusoidalEncoding, fused_position_mlp


def svf_cutoff_frequencies(sample_rate: float) -> torch.Tensor:
    """pi f_c / fs for the low shelf, the nine octave-band peaking sections (62.5 Hz .. 16 kHz) and the high shelf
    (reference gain_filters.py:369-374 with filters/geq.py:9-56 at its defaults)."""
    centre = []
    c = 31.25
    while c < 16000:
        c = c * 2.0
        centre.append(c)
    freqs = [centre[0] / math.sqrt(2.0)] + centre + [centre[-1] * math.sqrt(2.0)]
    return math.pi * torch.tensor(freqs, dtype=torch.float64) / sample_rate


def svf_to_biquads(svf_params: torch.Tensor, cutoffs: torch.Tensor, compress_pole_factor: float = 1.0) -> torch.Tensor:
    """(..., S, 2) constrained (resonance, gain dB) -> (..., S, 6) biquad coefficients (b0 b1 b2 a0 a1 a2), all
    receivers / groups / sections at once and differentiable. Section 0 is a low shelf, the last a high shelf, the
    others peaking (reference gain_filters.py:405-414); mixing coefficients of SVF.__post_init__ (:59-102) and the
    bilinear form of BiquadCascade.from_svf_coeffs (:116-151). Computed and returned in float64: the denominator
    a0 + a1 z^-1 + a2 z^-2 cancels to ~4 f_c^2 near DC, so coefficients held in float32 (the reference's) already
    perturb the response by 1e-3 there."""
    svf_params = svf_params.to(torch.float64)
    res, gain_db = svf_params[..., 0], svf_params[..., 1]
    gain = torch.pow(10.0, gain_db * 0.05)
    f = cutoffs if cutoffs.device == svf_params.device and cutoffs.dtype == svf_params.dtype else \
        cutoffs.to(device=svf_params.device, dtype=svf_params.dtype)
    ns = f.numel()
    sec = torch.arange(ns, device=svf_params.device)
    kind = (sec == 0).long() + 2 * (sec == ns - 1).long()  # 1: low shelf, 2: high shelf, 0: peaking (no host scalars)
    one = torch.ones_like(gain)
    m_lp = torch.where(kind == 1, gain, one)
    m_hp = torch.where(kind == 2, gain, one)
    m_bp = torch.where(kind == 0, 2 * res * gain, 2 * res * torch.sqrt(gain))
    r = compress_pole_factor
    b0 = f**2 * m_lp + f * m_bp + m_hp
    b1 = (2 * f**2 * m_lp - 2 * m_hp) * r
    b2 = (f**2 * m_lp - f * m_bp + m_hp) * r**2
    a0 = f**2 + 2 * res * f + 1
    a1 = ((2 * f**2 - 2) * r) * one
    a2 = (f**2 - 2 * res * f + 1) * r**2
    return torch.stack([b0, b1, b2, a0, a1, a2], dim=-1)


@dataclass
class BiquadCascade:
    """num_sos second-order sections: numerator / denominator coefficients (num_sos, 3) (reference :105-115)."""
    num_sos: int
    num_coeffs: torch.Tensor
    den_coeffs: torch.Tensor


class SOSFilter(nn.Module):
    """Frequency response of a biquad cascade on a z grid (reference gain_filters.py:206-241), evaluated by the
    K2s kernel (float64 inside, complex64 out)."""

    def __init__(self, num_biquads: int, biquad_cascade: Optional[BiquadCascade] = None,
                 device: Optional[torch.device] = 'cpu'):
        super().__init__()
        self.device = device
        self.num_biquads = num_biquads
        if biquad_cascade is not None:
            self.biquad_cascade = biquad_cascade

    def forward(self, z: torch.Tensor, biquad_cascade: Optional[BiquadCascade] = None) -> torch.Tensor:
        bc = self.biquad_cascade if biquad_cascade is None else biquad_cascade
        coef = torch.cat([bc.num_coeffs, bc.den_coeffs], dim=-1).reshape(1, 1, self.num_biquads, 6)
        return cascade_response(coef, z)[0, 0]


def cascade_response(coef: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """F[r,g,k] of every cascade in coef (R, G, S, 6): the projection kernel with one-hot group states."""
    rows, g = coef.shape[:2]
    z = z.to(device=coef.device, dtype=torch.complex128)
    eye = torch.eye(g, dtype=torch.complex64, device=coef.device)
    k = z.numel()
    out = [ops.svf_project(coef, z, eye[i].expand(k, g).contiguous(), None) for i in range(g)]
    return torch.stack(out, dim=1)


class SVF_from_MLP(nn.Module):
    """MLP(position) -> constrained SVF parameters -> one biquad cascade per (receiver, group) (reference
    gain_filters.py:318-430). The reference evaluates every cascade on the z grid into a dense (B, N, K) complex
    tensor with a Python loop over (receiver, group, section); here `coefficients()` returns the (B, G, S, 6)
    biquad table (vectorised, differentiable) that the projection kernel consumes, and `forward()` materialises
    the reference's (B, N, K) tensor only for callers that ask for it."""

    def __init__(self,
                 sample_rate: float,
                 num_groups: int,
                 num_delay_lines_per_group: int,
                 num_fourier_features: int,
                 num_hidden_layers: int,
                 num_neurons: int,
                 encoding_type: FeatureEncodingType = FeatureEncodingType.SINE,
                 compress_pole_factor: Optional[float] = 1.0,
                 position_type: str = "output_gains",
                 device: Optional[torch.device] = 'cpu'):
        super().__init__()
        self.num_groups = num_groups
        self.num_delay_lines_per_group = num_delay_lines_per_group
        self.num_delay_lines = num_groups * num_delay_lines_per_group
        self.position_type = position_type
        self.encoding_type = encoding_type
        self.compress_pole_factor = compress_pole_factor
        self.device = device
        if self.encoding_type != FeatureEncodingType.SINE:
            # unreachable in the reference too: its collate function never puts 'mesh_2D' into a batch
            # (dataloader.py:674-704), which the meshgrid branch of forward() reads (gain_filters.py:353, 511)
            raise NotImplementedError("meshgrid position encoding: the reference's own batches never carry 'mesh_2D', so "
                                      "that branch cannot run there either; only the sinusoidal encoding is built")
        self.svf_cutoff_freqs = svf_cutoff_frequencies(sample_rate)
        self.num_biquads = self.svf_cutoff_freqs.numel()
        self.encoder = SinusoidalEncoding(num_fourier_features)
        self.mlp = MLP(3 * num_fourier_features * 2, num_hidden_layers, num_neurons, self.num_groups,
                       self.num_biquads, num_params=2)
        self.sos_filter = SOSFilter(self.num_biquads, device=self.device)
        self.scaled_res = ScaledSigmoid(lower_limit=1e-6, upper_limit=1.0)   # resonance
        self.scaled_gains = ScaledSigmoid(lower_limit=-6, upper_limit=6)     # gain in dB

    def svf_parameters(self, x: Dict) -> torch.Tensor:
        """(B, G, S, 2): resonance in (1e-6, 1), gain in (-6, 6) dB (reference :376-421)."""
        position = x['listener_position'] if self.position_type == "output_gains" else x['source_position']
        param = next(self.mlp.parameters())
        position = position.to(param.device)
        self.batch_size = position.shape[0]
        raw = fused_position_mlp(self.encoder, self.mlp, position, final_act=0)
        if raw is None:
            raw = self.mlp(self.encoder(position).to(param.dtype))
        raw = raw.reshape(self.batch_size, self.num_groups, self.num_biquads, 2)
        return torch.stack([self.scaled_res(raw[..., 0]), self.scaled_gains(raw[..., 1])], dim=-1)

    def coefficients(self, x: Dict) -> torch.Tensor:
        """(B, G, S, 6) biquad coefficients of this batch; remembers the detached tables for get_parameters()."""
        svf = self.svf_parameters(x)
        if self.svf_cutoff_freqs.device != svf.device:  # moved once: no host -> device copy inside a step
            self.svf_cutoff_freqs = self.svf_cutoff_freqs.to(svf.device)
        coef = svf_to_biquads(svf, self.svf_cutoff_freqs, self.compress_pole_factor)
        self.svf_params = svf.detach()
        self.biquad_coeffs_ = coef.detach()
        return coef

    def forward(self, x: Dict) -> torch.Tensor:
        """(B, N, K) complex64 filter responses, every delay line of a group sharing the group's filter."""
        f = cascade_response(self.coefficients(x), x['z_values'])
        return f.repeat_interleave(self.num_delay_lines_per_group, dim=1)

    @property
    def biquad_cascade(self) -> List[List[BiquadCascade]]:
        c = self.biquad_coeffs_
        return [[BiquadCascade(self.num_biquads, c[b, g, :, :3], c[b, g, :, 3:]) for g in range(self.num_groups)]
                for b in range(c.shape[0])]

    def get_parameters(self) -> Tuple:
        c = self.biquad_coeffs_
        coeffs = [[c[b, g].cpu().numpy() for g in range(self.num_groups)] for b in range(c.shape[0])]
        return (self.svf_params, coeffs)

    @torch.no_grad()
    def get_param_dict(self, x: Dict) -> Dict:
        self.coefficients(x)
        svf, coeffs = self.get_parameters()
        return {'svf_params': svf.squeeze().cpu().numpy(), 'biquad_coeffs': coeffs}


class Gains_from_MLP(nn.Module):

    def __init__(self,
                 num_groups: int,
                 num_delay_lines_per_group: int,
                 num_fourier_features: int,
                 num_hidden_layers: int,
                 num_neurons: int,
                 encoding_type: FeatureEncodingType = FeatureEncodingType.SINE,
                 position_type: str = "output_gains",
                 device: Optional[torch.device] = 'cpu',
                 gain_limits: Optional[Tuple] = None):
        super().__init__()
        self.num_groups = num_groups
        self.num_delay_lines_per_group = num_delay_lines_per_group
        self.position_type = position_type
        self.encoding_type = encoding_type
        self.device = device
        if self.encoding_type != FeatureEncodingType.SINE:
            # unreachable in the reference too: its collate function never puts 'mesh_2D' into a batch
            # (dataloader.py:674-704), which the meshgrid branch of forward() reads (gain_filters.py:353, 511)
            raise NotImplementedError("meshgrid position encoding: the reference's own batches never carry 'mesh_2D', so "
                                      "that branch cannot run there either; only the sinusoidal encoding is built")
        self.encoder = SinusoidalEncoding(num_fourier_features)
        self.mlp = MLP(3 * num_fourier_features * 2, num_hidden_layers, num_neurons, self.num_groups,
                       num_biquads_in_cascade=1, num_params=1)
        lo, hi = (-1.0, 1.0) if gain_limits is None else gain_limits
        self.scaled_sigmoid = ScaledSigmoid(lower_limit=lo, upper_limit=hi)

    def gains(self, x: Dict) -> torch.Tensor:
        """(B, G) gains: scaled_sigmoid(MLP(enc(position)))  (reference gain_filters.py:502-524)."""
        position = x['norm_listener_position'] if self.position_type == "output_gains" else x['source_position']
        param = next(self.mlp.parameters())
        position = position.to(param.device)
        self.batch_size = position.shape[0]
        gains = fused_position_mlp(self.encoder, self.mlp, position, final_act=1, lo=self.scaled_sigmoid.lower_limit,
                                   hi=self.scaled_sigmoid.upper_limit)  # K7: one kernel forward, two backward
        if gains is None:
            out = self.mlp(self.encoder(position))
            gains = self.scaled_sigmoid(out.view(-1)).view(self.batch_size, self.num_groups)
        self.gains_ = gains.detach()  # for get_parameters(); detached so no autograd graph outlives the step
        return gains

    def forward(self, x: Dict) -> torch.Tensor:
        """(B, N, K) expansion with the reference's shape; a view, nothing is materialised."""
        g = self.gains(x)
        k = len(x['z_values'])
        return g.repeat_interleave(self.num_delay_lines_per_group, dim=1).unsqueeze(-1).expand(-1, -1, k)

    def get_parameters(self):
        return self.gains_

    @torch.no_grad()
    def get_param_dict(self, x: Dict) -> Dict:
        return {'gains': self.gains(x).squeeze().cpu().numpy()}
