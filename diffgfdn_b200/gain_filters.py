"""Receiver/source gain generators (reference diff_gfdn/gain_filters.py:437-555).

`Gains_from_MLP` maps a (normalised) position to one scalar gain per group. The reference then repeats those gains
to a dense (B, N, K) tensor (gain_filters.py:526-536) which the model multiplies bin by bin; here the kernels
consume the (B, G) table directly (`gains()`), and `forward()` returns a stride-0 *view* with the reference's shape
for callers that want it."""
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from .config.config import FeatureEncodingType
from .dnn import MLP, ScaledSigmoid, SinusoidalEncoding, fused_position_mlp


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
            raise NotImplementedError("only the sinusoidal position encoding is on the B200 hot path "
                                      "(no shipped config uses the meshgrid encoding)")
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
