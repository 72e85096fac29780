"""Differentiable Grouped-FDN models with the reference's nn.Module surface, on B200 kernels.

Class names, constructor signatures, `forward(x: dict)` contract, parameter/buffer names (state_dict keys) and
helper methods follow reference diff_gfdn/model.py; the arithmetic is restructured:

    reference: P_k = inv(D_k Gamma^-1 - A) for every bin, expand receiver gains to (B, N, K), two einsums
               -> O(B K N^2) work and several (B, N, K) complex temporaries (model.py:583-619)
    here:      x_k = (D_k Gamma^-1 - A)^-1 b once per bin (b is shared by all receivers), fold c and the group
               structure into y[k,g], then H[r,k] = sum_g s[r,g] y[k,g] + d[r,k] -- a length-G contraction
               that streams at HBM bandwidth.

Outputs are complex64 (the reference returns complex128 only because its `d` input is complex128, quirk Q6).
Inputs that live on the host are copied to the module's device; the kernels themselves have no CPU path."""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import ops
from .absorption_filters import decay_times_to_gain_filters_geq, decay_times_to_gain_per_sample
from .config.config import CouplingMatrixType, FeedbackLoopConfig, OutputFilterConfig
from .feedback_loop import FeedbackLoop
from .dnn import ScaledSigmoid
from .gain_filters import (Gains_from_MLP, SOSFilter, SVF_from_MLP, cascade_response, svf_cutoff_frequencies,
                           svf_to_biquads)
from .sh_gains import Directional_Beamforming_Weights_from_MLP


def _resolve_device(device) -> torch.device:
    if device is None or str(device) in ("gpu", "cuda"):
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    return torch.device(device)


class DiffGFDN(nn.Module):
    """Parent module (reference model.py:24-299)."""

    def __init__(self,
                 sample_rate: int,
                 num_groups: int,
                 delays: List[int],
                 device: torch.device,
                 feedback_loop_config: FeedbackLoopConfig,
                 use_absorption_filters: bool,
                 learn_common_decay_times: bool,
                 common_decay_times: Optional[List] = None,
                 band_centre_hz: Optional[List] = None,
                 colorless_fdn_params: Optional[List] = None,
                 use_colorless_loss: bool = False):
        super().__init__()
        self.sample_rate = sample_rate
        self.device = _resolve_device(device)
        self.num_groups = num_groups
        self.num_delay_lines = len(delays)
        self.num_delay_lines_per_group = int(self.num_delay_lines / self.num_groups)
        self.use_absorption_filters = use_absorption_filters
        self.band_centre_hz = band_centre_hz
        self.common_decay_times = common_decay_times
        self.learn_common_decay_times = learn_common_decay_times
        self.use_colorless_loss = use_colorless_loss
        # whether forward() also builds the (N, K, G) per-delay-line tensor of reference model.py:243-246;
        # nothing in the training path reads it, so the fused trainer switches it off
        self.return_per_delay_outputs = True
        self.delays = torch.tensor(delays, dtype=torch.float32, device=self.device)
        self.register_buffer('delay_buffer', self.delays)
        per = self.num_delay_lines_per_group
        self.delays_by_group = [self.delays[i:i + per] for i in range(0, self.num_delay_lines, per)]
        self._init_io_gains(colorless_fdn_params)
        self._init_absorption(band_centre_hz)
        self._init_feedback(feedback_loop_config, colorless_fdn_params)

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.delays = self.delay_buffer
        per = self.num_delay_lines_per_group
        self.delays_by_group = [self.delays[i:i + per] for i in range(0, self.num_delay_lines, per)]
        self.device = self.delay_buffer.device
        fl = getattr(self, "feedback_loop", None)
        if fl is not None and getattr(fl, "absorption_coeffs", None) is not None:
            # the feedback loop reads the SAME tensor as the 'delay_filters' buffer (load_state_dict copies in place)
            fl.absorption_coeffs = fl.delay_line_gains = self.delay_filters
            fl._gamma_cache = None
        return self

    def _init_io_gains(self, colorless_fdn_params=None):
        """reference model.py:95-122"""
        n = self.num_delay_lines
        if colorless_fdn_params is None:
            self.input_gains = nn.Parameter(((2 * torch.randn(n, 1) - 1) / n).to(self.device))
            self.output_gains = nn.Parameter(((2 * torch.randn(n, 1) - 1) / n).to(self.device))
        else:
            self.input_gains = torch.tensor([colorless_fdn_params[i].opt_input_gains.tolist()
                                             for i in range(self.num_groups)], device=self.device).view(-1, 1)
            self.output_gains = torch.tensor([colorless_fdn_params[i].opt_output_gains.tolist()
                                              for i in range(self.num_groups)], device=self.device).view(-1, 1)

    def _init_absorption(self, band_centre_hz=None):
        """reference model.py:124-166 (broadband branch; GEQ filter design is init-time host code, see
        FeedbackLoop.set_absorption_response)."""
        if self.common_decay_times is None or self.learn_common_decay_times:
            self.gain_per_sample = None
            return
        if self.use_absorption_filters:
            # One graphic equaliser per delay line (reference absorption_filters.py:108-155 -> filters/geq.py: an LBFGS
            # fit per line, ~1 s each; here the exact least-squares minimiser, init-time host code). The buffer keeps
            # the reference's name and (N, bands + 3, 3, 2) layout so its checkpoints load unchanged (load_state_dict
            # or set_absorption_filters replace the design). The per-bin responses are evaluated on the GPU.
            if band_centre_hz is None:
                raise RuntimeError("use_absorption_filters=True needs band_centre_hz (one T60 per band and group)")
            t60 = np.squeeze(np.asarray(self.common_decay_times))  # (bands, G)
            per_group = [decay_times_to_gain_filters_geq(band_centre_hz, t60[:, i] if t60.ndim > 1 else t60,
                                                         self.delays_by_group[i].cpu().numpy(), self.sample_rate)
                         for i in range(self.num_groups)]  # each (bands + 3, L, 3, 2)
            stacked = torch.stack(per_group).permute(0, 2, 1, 3, 4)  # (G, L, bands + 3, 3, 2), reference :141-147
            self.gain_per_sample = stacked.reshape(self.num_delay_lines, len(band_centre_hz) + 3, 3, 2).to(self.device)
            self.filter_order = 3
            self.register_buffer('delay_filters', self.gain_per_sample)
            return
        t60 = np.squeeze(np.asarray(self.common_decay_times))
        gains = [decay_times_to_gain_per_sample(float(t60[i]), self.delays_by_group[i].cpu().numpy(),
                                                self.sample_rate).tolist() for i in range(self.num_groups)]
        self.gain_per_sample = torch.flatten(torch.tensor(gains, device=self.device)).to(torch.float32)
        self.register_buffer('delay_filters', self.gain_per_sample)

    def _init_feedback(self, feedback_loop_config: FeedbackLoopConfig, colorless_fdn_params=None):
        """reference model.py:168-207"""
        kw = dict(gains=self.gain_per_sample, use_zero_coupling=feedback_loop_config.use_zero_coupling,
                  common_decay_times=self.common_decay_times,
                  coupling_matrix_type=feedback_loop_config.coupling_matrix_type, device=self.device)
        if colorless_fdn_params is None:
            kw['coupling_matrix_order'] = feedback_loop_config.pu_matrix_order
        else:
            kw['colorless_feedback_matrix'] = torch.stack(
                [torch.from_numpy(colorless_fdn_params[i].opt_feedback_matrix) for i in range(self.num_groups)], dim=0)
        self.feedback_loop = FeedbackLoop(self.sample_rate, self.num_groups, self.num_delay_lines_per_group,
                                          self.delays, self.use_absorption_filters, **kw)

    @torch.no_grad()
    def set_absorption_filters(self, coeffs: torch.Tensor):
        """Absorption-filter coefficients from an external design (decay_times_to_gain_filters_geq of the reference,
        (N, S, 3, 2)), copied into the 'delay_filters' buffer."""
        self.delay_filters.copy_(torch.as_tensor(coeffs).to(self.delay_filters))

    # ---- helpers ---------------------------------------------------------------------------------------
    def _on_device(self, t: torch.Tensor, dtype=None) -> torch.Tensor:
        if t.device != self.device or (dtype is not None and t.dtype != dtype):
            t = t.to(device=self.device, dtype=dtype, non_blocking=True)
        return t

    def _gains_vec(self, g) -> torch.Tensor:
        return g.reshape(-1)

    def sub_fdn_output(self, z: torch.Tensor) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """Response of each lossless sub-FDN, with the RAW mixing matrices M_g and no absorption (reference
        model.py:209-252, quirk Q1): G small solves per bin (group mode of K1). Returns Hout (K, G) and Hout_per_del (N, K, G)."""
        z = self._on_device(z, torch.complex128)
        xs, hout = ops.gfdn_solve_groups(z, self.delays.to(torch.int32), self.feedback_loop.M, None,
                                         self._gains_vec(self.input_gains), self._gains_vec(self.output_gains))
        if not self.return_per_delay_outputs:
            return hout, None
        per = (xs * self._gains_vec(self.output_gains).to(xs.dtype)).transpose(0, 1)  # (N, K)
        idx = torch.arange(self.num_delay_lines, device=per.device)
        groups = torch.arange(self.num_groups, device=per.device)
        # one-hot of each delay line's group, built from comparisons (no host scalar: this runs inside graph captures)
        onehot = ((idx // self.num_delay_lines_per_group).unsqueeze(-1) == groups).to(per.dtype).unsqueeze(1)
        return hout, per.unsqueeze(-1) * onehot

    @torch.no_grad()
    def get_param_dict(self) -> Dict:
        """reference model.py:254-299"""
        fl = self.feedback_loop
        d = {'delays': self.delays.squeeze().cpu().numpy(),
             'gains_per_sample': fl.delay_line_gains.squeeze().cpu().numpy(),
             'input_gains': self.input_gains.squeeze().cpu().numpy(),
             'output_gains': self.output_gains.squeeze().cpu().numpy(),
             'coupled_feedback_matrix': fl.get_coupled_feedback_matrix().squeeze().cpu().numpy()}
        if fl.coupling_matrix_type != CouplingMatrixType.RANDOM:  # the unstructured matrix has no group / coupling parts
            d['individual_mixing_matrix'] = fl.M.squeeze().cpu().numpy()
            d['coupling_matrix'] = fl.nd_unitary(fl.alpha, self.num_groups).squeeze().cpu().numpy()
        return d


class DiffGFDNVarReceiverPos(DiffGFDN):
    """GFDN for a grid of receiver positions with MLP-driven receiver gains (reference model.py:502-661)."""

    def __init__(self,
                 sample_rate: int,
                 num_groups: int,
                 delays: List[int],
                 device: torch.device,
                 feedback_loop_config: FeedbackLoopConfig,
                 output_filter_config: OutputFilterConfig,
                 use_absorption_filters: bool,
                 learn_common_decay_times: Optional[bool] = False,
                 common_decay_times: Optional[List] = None,
                 band_centre_hz: Optional[List] = None,
                 colorless_fdn_params: Optional[List] = None,
                 use_colorless_loss: bool = False):
        super().__init__(sample_rate, num_groups, delays, device, feedback_loop_config, use_absorption_filters,
                         learn_common_decay_times, common_decay_times, band_centre_hz, colorless_fdn_params,
                         use_colorless_loss)
        self.use_svf_in_output = output_filter_config.use_svfs
        self.input_scalars = torch.ones(self.num_groups, 1)
        if self.use_svf_in_output:
            self.output_filters = SVF_from_MLP(self.sample_rate, self.num_groups, self.num_delay_lines_per_group,
                                               output_filter_config.num_fourier_features,
                                               output_filter_config.num_hidden_layers,
                                               output_filter_config.num_neurons_per_layer,
                                               output_filter_config.encoding_type,
                                               output_filter_config.compress_pole_factor,
                                               device=self.device).to(self.device)
        else:
            self.output_scalars = Gains_from_MLP(self.num_groups, self.num_delay_lines_per_group,
                                                 output_filter_config.num_fourier_features,
                                                 output_filter_config.num_hidden_layers,
                                                 output_filter_config.num_neurons_per_layer,
                                                 output_filter_config.encoding_type, device=self.device).to(self.device)

    def forward(self, x: Dict, output_scalars: Optional[torch.Tensor] = None):
        """H(z) = c^T (D Gamma^-1 - A)^-1 b + d for every receiver of the batch and every bin.

        x: 'z_values' (K,) complex, 'listener_position' (B,3), 'norm_listener_position' (B,3),
           'target_early_response' (B,K) complex.  Returns H (B,K) complex64, or (H, (H_sub, H_sub_per_del))."""
        z = self._on_device(x['z_values'], torch.complex128)
        self.batch_size = x['listener_position'].shape[0]
        coef = s = None
        if self.use_svf_in_output:  # the provided-scalars override only exists on the gains branch (model.py:589-605)
            coef = self.output_filters.coefficients(x)
        elif output_scalars is None:
            s = self.output_scalars.gains(x)
        else:
            assert output_scalars.shape == (self.batch_size, self.num_groups)
            s = self._on_device(output_scalars, torch.float32)
        _, y = self.feedback_loop.solve(z, self._gains_vec(self.input_gains), self._gains_vec(self.output_gains))
        d = x.get('target_early_response')
        d = None if d is None else self._on_device(d, torch.complex64)
        H = ops.svf_project(coef, z, y, d) if coef is not None else ops.receiver_project(s, y, d)
        if self.use_colorless_loss:
            return H, self.sub_fdn_output(z)
        return H

    def get_parameters(self) -> Tuple:
        """reference model.py:627-636 (only defined for the SVF variant there)."""
        fl = self.feedback_loop
        svf_params, biquad_coeffs = self.output_filters.get_parameters()
        return (self.delays, fl.delay_line_gains, self.input_gains, fl.M, fl.nd_unitary(fl.alpha, self.num_groups),
                fl.get_coupled_feedback_matrix(), svf_params, biquad_coeffs)

    @torch.no_grad()
    def get_param_dict_inference(self, data: Dict) -> Dict:
        if self.use_svf_in_output:
            out = self.output_filters.get_param_dict(data)
            return {'output_svf_params': out['svf_params'], 'output_biquad_coeffs': out['biquad_coeffs']}
        return {'output_scalars': self.output_scalars.get_param_dict(data)['gains']}

    @torch.no_grad()
    def get_param_dict(self) -> Dict:
        d = super().get_param_dict()
        d['input_scalars'] = self.input_scalars.squeeze().cpu().numpy()
        return d


class DiffGFDNVarSourceReceiverPos(DiffGFDN):
    """GFDN for a grid of source AND receiver positions: MLP-driven factors on both sides (reference model.py:303-500).

    H[r,k] = sum_{g,g'} C_g(r,k) T[k,g,g'] B_g'(r,k) + d[r,k] with T the group-to-group transfer functions
    (FeedbackLoop.transfer_matrix): the receiver-side projection kernel runs once per source group with the row
    gains (or the first biquad numerator) scaled by that group's source gain, chaining through its additive input."""

    def __init__(self,
                 sample_rate: int,
                 num_groups: int,
                 delays: List[int],
                 device: torch.device,
                 feedback_loop_config: FeedbackLoopConfig,
                 output_filter_config: OutputFilterConfig,
                 input_filter_config: OutputFilterConfig,
                 use_absorption_filters: bool,
                 learn_common_decay_times: bool = False,
                 common_decay_times: Optional[List] = None,
                 band_centre_hz: Optional[List] = None,
                 colorless_fdn_params: Optional[List] = None,
                 use_colorless_loss: bool = False):
        super().__init__(sample_rate, num_groups, delays, device, feedback_loop_config, use_absorption_filters,
                         learn_common_decay_times, common_decay_times, band_centre_hz, colorless_fdn_params,
                         use_colorless_loss)
        self.use_svf_in_output = output_filter_config.use_svfs
        self.use_svf_in_input = input_filter_config.use_svfs
        if self.use_svf_in_output:
            self.output_filters = SVF_from_MLP(self.sample_rate, self.num_groups, self.num_delay_lines_per_group,
                                               output_filter_config.num_fourier_features,
                                               output_filter_config.num_hidden_layers,
                                               output_filter_config.num_neurons_per_layer,
                                               output_filter_config.encoding_type,
                                               output_filter_config.compress_pole_factor,
                                               position_type="output_gains", device=self.device).to(self.device)
        else:
            self.output_scalars = Gains_from_MLP(self.num_groups, self.num_delay_lines_per_group,
                                                 output_filter_config.num_fourier_features,
                                                 output_filter_config.num_hidden_layers,
                                                 output_filter_config.num_neurons_per_layer,
                                                 output_filter_config.encoding_type, position_type="output_gains",
                                                 device=self.device).to(self.device)
        if self.use_svf_in_input:  # reference model.py:376-388: cascades driven by x['source_position']
            self.input_filters = SVF_from_MLP(self.sample_rate, self.num_groups, self.num_delay_lines_per_group,
                                              input_filter_config.num_fourier_features,
                                              input_filter_config.num_hidden_layers,
                                              input_filter_config.num_neurons_per_layer,
                                              input_filter_config.encoding_type,
                                              input_filter_config.compress_pole_factor,
                                              position_type="input_gains", device=self.device).to(self.device)
        else:
            self.input_scalars = Gains_from_MLP(self.num_groups, self.num_delay_lines_per_group,
                                                input_filter_config.num_fourier_features,
                                                input_filter_config.num_hidden_layers,
                                                input_filter_config.num_neurons_per_layer,
                                                input_filter_config.encoding_type, position_type="input_gains",
                                                device=self.device).to(self.device)

    def forward(self, x: Dict):
        """H[r,k] = sum_{g,g'} C[r,g,k] T[k,g,g'] B[r,g',k] + d[r,k] with T the G x G group-to-group transfer functions of
        the loop (one K1 solve per source group) and C / B the receiver / source factors: (B, G) gains or the responses
        of the (receiver, group) SVF cascades (reference model.py:402-452 expands both to (B, N, K))."""
        z = self._on_device(x['z_values'], torch.complex128)
        self.batch_size = x['listener_position'].shape[0]
        coef = self.output_filters.coefficients(x) if self.use_svf_in_output else None
        s_rx = None if self.use_svf_in_output else self.output_scalars.gains(x)
        if self.use_svf_in_input:  # source-side cascade responses F_in[r,g',k], (B, G, K) complex64 (K2s, differentiable)
            f_src = cascade_response(self.input_filters.coefficients(x), z)
        else:
            s_src = self.input_scalars.gains(x)  # (B, G) from x['source_position']
        T = self.feedback_loop.transfer_matrix(z, self._gains_vec(self.input_gains), self._gains_vec(self.output_gains))
        d = x.get('target_early_response')
        H = None if d is None else self._on_device(d, torch.complex64)
        for gp in range(self.num_groups):
            if self.use_svf_in_input:  # receiver side of source group gp, then the per-receiver, per-bin source factor
                part = ops.svf_project(coef, z, T[gp], None) if coef is not None else ops.receiver_project(s_rx, T[gp], None)
                part = part * f_src[:, gp]
                H = part if H is None else H + part
                continue
            w = s_src[:, gp]
            if coef is not None:  # scale the cascade by the source gain through its first numerator
                scale = torch.ones_like(coef)
                scale[:, :, 0, :3] = w.reshape(-1, 1, 1)
                H = ops.svf_project(coef * scale, z, T[gp], H)
            else:
                H = ops.receiver_project(s_rx * w.unsqueeze(-1), T[gp], H)
        if self.use_colorless_loss:
            return H, self.sub_fdn_output(z)
        return H

    @torch.no_grad()
    def get_param_dict_inference(self, data: Dict) -> Dict:
        if self.use_svf_in_input:
            i = self.input_filters.get_param_dict(data)
            out = {'input_svf_params': i['svf_params'], 'input_biquad_coeffs': i['biquad_coeffs']}
        else:
            out = {'input_scalars': self.input_scalars.get_param_dict(data)['gains']}
        if self.use_svf_in_output:
            o = self.output_filters.get_param_dict(data)
            out.update({'output_svf_params': o['svf_params'], 'output_biquad_coeffs': o['biquad_coeffs']})
        else:
            out['output_scalars'] = self.output_scalars.get_param_dict(data)['gains']
        return out


class DiffGFDNSinglePos(DiffGFDN):
    """GFDN for ONE source-receiver pair: learnable per-group scalars or SVF cascades on either side (reference
    model.py:667-960). H[k] = sum_{g,g'} C_g(z_k) T[k,g,g'] B_g'(z_k) + d[k]."""

    def __init__(self,
                 sample_rate: int,
                 num_groups: int,
                 delays: List[int],
                 device: torch.device,
                 feedback_loop_config: FeedbackLoopConfig,
                 output_filter_config: OutputFilterConfig,
                 use_absorption_filters: bool,
                 learn_common_decay_times: Optional[bool] = False,
                 common_decay_times: Optional[List] = None,
                 band_centre_hz: Optional[List] = None,
                 colorless_fdn_params: Optional[List] = None,
                 use_colorless_loss: bool = False,
                 input_filter_config: Optional[OutputFilterConfig] = None):
        super().__init__(sample_rate, num_groups, delays, device, feedback_loop_config, use_absorption_filters,
                         learn_common_decay_times, common_decay_times, band_centre_hz, colorless_fdn_params,
                         use_colorless_loss)
        self.use_svf_in_input = input_filter_config.use_svfs if input_filter_config is not None else False
        self.use_svf_in_output = output_filter_config.use_svfs
        if self.use_svf_in_output or self.use_svf_in_input:
            self.svf_cutoff_freqs = svf_cutoff_frequencies(self.sample_rate).to(self.device)
            self.num_biquads = self.svf_cutoff_freqs.numel()
            self.compress_pole_factor = output_filter_config.compress_pole_factor
        for side, svf in (("input", self.use_svf_in_input), ("output", self.use_svf_in_output)):
            if svf:  # random resonance, 0 dB gains (reference :744-776)
                init = torch.randn(self.num_groups, self.num_biquads, 2)
                init[..., 1] = 0.0
                setattr(self, f"{side}_svf_params", nn.Parameter(init.to(self.device)))
                setattr(self, f"{side}_filters", SOSFilter(self.num_biquads, device=self.device))
                setattr(self, f"{side}_scaled_res", ScaledSigmoid(lower_limit=1e-6, upper_limit=1.0))
                setattr(self, f"{side}_scaled_gains", ScaledSigmoid(lower_limit=-6.0, upper_limit=6.0))
            else:
                setattr(self, f"{side}_scalars",
                        nn.Parameter((torch.ones(self.num_groups, 1) / np.sqrt(self.num_groups)).to(self.device)))

    def biquad_coefficients(self, filt_type: str = 'output') -> torch.Tensor:
        """(G, S, 6) biquads of the learnable cascades of one side (reference get_filter :838-893)."""
        side = 'output' if filt_type == 'output' else 'input'
        raw = getattr(self, f"{side}_svf_params")
        svf = torch.stack([getattr(self, f"{side}_scaled_res")(raw[..., 0]),
                           getattr(self, f"{side}_scaled_gains")(raw[..., 1])], dim=-1)
        coef = svf_to_biquads(svf, self.svf_cutoff_freqs, self.compress_pole_factor)
        setattr(self, f"{side}_biquad_coeffs_", coef.detach())
        return coef

    def get_filter(self, z_values: torch.Tensor, filt_type: str = 'output') -> torch.Tensor:
        """(N, K) complex64 filter responses, every delay line of a group sharing the group's filter."""
        f = cascade_response(self.biquad_coefficients(filt_type).unsqueeze(0), self._on_device(z_values, torch.complex128))
        return f[0].repeat_interleave(self.num_delay_lines_per_group, dim=0)

    def forward(self, x: Dict):
        z = self._on_device(x['z_values'], torch.complex128)
        T = self.feedback_loop.transfer_matrix(z, self._gains_vec(self.input_gains), self._gains_vec(self.output_gains))
        if self.use_svf_in_input:
            f_in = cascade_response(self.biquad_coefficients('input').unsqueeze(0), z)[0]  # (G, K) complex64
            y = sum(T[gp] * f_in[gp].unsqueeze(-1) for gp in range(self.num_groups))
        else:
            y = sum(T[gp] * self.input_scalars[gp, 0] for gp in range(self.num_groups))
        d = x.get('target_early_response')
        d = None if d is None else self._on_device(d, torch.complex64).reshape(1, -1)
        if self.use_svf_in_output:
            H = ops.svf_project(self.biquad_coefficients('output').unsqueeze(0), z, y, d)[0]
        else:
            H = ops.receiver_project(self.output_scalars.reshape(1, -1), y, d)[0]
        if self.use_colorless_loss:
            return H, self.sub_fdn_output(z)
        return H

    @torch.no_grad()
    def get_param_dict(self) -> Dict:
        d = super().get_param_dict()
        d['absorption_coeffs'] = self.feedback_loop.delay_line_gains
        for side, svf in (("input", self.use_svf_in_input), ("output", self.use_svf_in_output)):
            if svf:
                coef = self.biquad_coefficients(side)
                raw = getattr(self, f"{side}_svf_params")
                d[f'{side}_svf_params'] = torch.stack([getattr(self, f"{side}_scaled_res")(raw[..., 0]),
                                                       getattr(self, f"{side}_scaled_gains")(raw[..., 1])],
                                                      dim=-1).squeeze().cpu().numpy()
                d[f'{side}_biquad_coeffs'] = [coef[g].cpu().numpy() for g in range(self.num_groups)]
            else:
                d[f'{side}_scalars'] = getattr(self, f"{side}_scalars").squeeze().cpu().numpy()
        return d


class DiffDirectionalFDNVarReceiverPos(DiffGFDN):
    """Directional FDN: one delay line per (group, SH channel), MLP-driven SH gains (reference model.py:975-1126)."""

    def __init__(self,
                 sample_rate: int,
                 num_groups: int,
                 delays: List[int],
                 device: torch.device,
                 feedback_loop_config: FeedbackLoopConfig,
                 output_filter_config: OutputFilterConfig,
                 ambi_order: int,
                 desired_directions: Optional[np.ndarray],
                 use_absorption_filters: bool = False,
                 learn_common_decay_times: Optional[bool] = False,
                 common_decay_times: Optional[List] = None,
                 band_centre_hz: Optional[List] = None,
                 colorless_fdn_params: Optional[List] = None,
                 use_colorless_loss: bool = False,
                 analysis_matrix: Optional[np.ndarray] = None):
        super().__init__(sample_rate, num_groups, delays, device, feedback_loop_config, use_absorption_filters,
                         learn_common_decay_times, common_decay_times, band_centre_hz, colorless_fdn_params,
                         use_colorless_loss)
        self.ambi_order = ambi_order
        assert self.num_delay_lines_per_group == (self.ambi_order + 1)**2, \
            "Number of delay lines per group must be equal to the number of ambisonics channels"
        self.input_scalars = torch.ones(self.num_groups, 1)
        self.use_svf_in_output = False
        self.sh_output_scalars = Directional_Beamforming_Weights_from_MLP(
            self.num_groups, self.ambi_order, output_filter_config.num_fourier_features,
            output_filter_config.num_hidden_layers, output_filter_config.num_neurons_per_layer,
            desired_directions=desired_directions, device=self.device,
            beamformer_type=output_filter_config.beamformer_type,
            use_skip_connections=output_filter_config.use_skip_connections,
            analysis_matrix=analysis_matrix).to(self.device)

    def forward(self, x: Dict):
        """H_sh[r,l,k] = sum_g w[r,g,l] c[g,l] (P_k^T b)[gL+l]  -> (B, (N_sp+1)^2, K) complex64.

        The reference contracts the FIRST index of P with b here (einsum 'knm,bnk->bmk', model.py:1083), i.e. the
        state is P^T b: the solve runs on A^T (quirk Q11 in DESIGN.md)."""
        z = self._on_device(x['z_values'], torch.complex128)
        self.batch_size = x['listener_position'].shape[0]
        w = self.sh_output_scalars(x, normalise_weights=True)  # (B, G, L)
        c = self._gains_vec(self.output_gains).reshape(1, self.num_groups, self.num_delay_lines_per_group)
        xs, _ = self.feedback_loop.solve(z, self._gains_vec(self.input_gains), self._gains_vec(self.output_gains),
                                         transpose=True)
        H = ops.sh_project(w * c, xs)
        if self.use_colorless_loss:
            return H, self.sub_fdn_output(z)
        return H

    @torch.no_grad()
    def get_param_dict_inference(self, data: Dict, normalise_weights: bool = False) -> Dict:
        return {'output_scalars': self.sh_output_scalars.get_param_dict(data, normalise_weights)['beamformer_weights']}

    @torch.no_grad()
    def get_param_dict(self) -> Dict:
        d = super().get_param_dict()
        d['input_scalars'] = self.input_scalars.squeeze().cpu().numpy()
        return d


__all__ = ["DiffGFDN", "DiffGFDNVarReceiverPos", "DiffGFDNVarSourceReceiverPos", "DiffGFDNSinglePos",
           "DiffDirectionalFDNVarReceiverPos", "CouplingMatrixType"]
