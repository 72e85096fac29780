"""SH-domain receiver gains for the directional FDN (reference spatial_sampling/model.py:17-190).

The analysis matrix Y (directions x SH channels) comes from spaudiopy's `design_sph_filterbank` in the reference,
a third-party package that is not part of the hot path: it is used when importable, otherwise the caller supplies
`analysis_matrix` (parity tests inject a synthetic one into both sides)."""
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from .config.config import BeamformerType
from .dnn import MLP, MLP_SkipConnections, Sigmoid, SinusoidalEncoding, fused_position_mlp


def design_analysis_matrix(ambi_order: int, desired_directions: np.ndarray,
                           beamformer_type: Optional[BeamformerType]) -> np.ndarray:
    """reference spatial_sampling/model.py:49-72 (requires spaudiopy)."""
    try:
        import spaudiopy as sp
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise RuntimeError("spaudiopy is needed to design the SH analysis matrix; pass analysis_matrix= instead") from e
    if beamformer_type == BeamformerType.MAX_DI:
        w = sp.sph.cardioid_modal_weights(ambi_order)
    elif beamformer_type == BeamformerType.MAX_RE:
        w = sp.sph.maxre_modal_weights(ambi_order)
    elif beamformer_type == BeamformerType.BUTTER:
        w = sp.sph.butterworth_modal_weights(ambi_order, k=5, n_c=3)
    else:
        w = np.ones(ambi_order + 1)
    y, _ = sp.sph.design_sph_filterbank(ambi_order, desired_directions[0, :], np.pi / 2 - desired_directions[1, :], w,
                                        mode='energy', sh_type='real')
    return np.asarray(y)


class Directional_Beamforming_Weights_from_MLP(nn.Module):

    def __init__(self,
                 num_groups: int,
                 ambi_order: int,
                 num_fourier_features: int,
                 num_hidden_layers: int,
                 num_neurons: int,
                 desired_directions: Optional[np.ndarray] = None,
                 device: Optional[torch.device] = 'cpu',
                 beamformer_type: Optional[BeamformerType] = None,
                 use_skip_connections: Optional[bool] = False,
                 analysis_matrix: Optional[np.ndarray] = None):
        super().__init__()
        self.num_groups = num_groups
        self.device = device
        self.ambi_order = ambi_order
        self.num_fourier_features = num_fourier_features
        self.num_out_features = (ambi_order + 1)**2
        if analysis_matrix is None:
            analysis_matrix = design_analysis_matrix(ambi_order, desired_directions, beamformer_type)
        # plain attribute (not a buffer) so that state_dict keys match the reference
        self.analysis_matrix = torch.as_tensor(np.asarray(analysis_matrix), dtype=torch.float32, device=device)
        self.scaling = Sigmoid()
        self.encoder = SinusoidalEncoding(num_fourier_features)
        cls = MLP_SkipConnections if use_skip_connections else MLP
        self.mlp = cls(3 * num_fourier_features * 2, num_hidden_layers, num_neurons, self.num_groups,
                       num_biquads_in_cascade=1, num_params=self.num_out_features)

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.analysis_matrix = fn(self.analysis_matrix)
        return self

    def normalise_weights(self, weights: torch.Tensor) -> torch.Tensor:
        """reference spatial_sampling/model.py:78-80"""
        return weights / (torch.norm(weights, dim=-1, keepdim=True) + 1e-6)

    def forward(self, x: Dict, normalise_weights: bool = False) -> torch.Tensor:
        """(B, G, (N_sp+1)^2) weights (reference spatial_sampling/model.py:169-190)."""
        position = x['norm_listener_position'].to(next(self.mlp.parameters()).device)
        self.batch_size = position.shape[0]
        w = fused_position_mlp(self.encoder, self.mlp, position)  # K7 kernels
        if w is None:
            w = self.mlp(self.encoder(position))
        w = w.reshape(self.batch_size, self.num_groups, self.num_out_features)
        if normalise_weights:
            w = self.normalise_weights(w)
        self.weights = w
        return w

    def get_directional_amplitudes(self) -> torch.Tensor:
        return self.scaling(torch.einsum('jn, bkn-> bjk', self.analysis_matrix, self.weights))

    def get_parameters(self):
        return self.weights

    @torch.no_grad()
    def get_param_dict(self, x: Dict, normalise_weights: bool = False) -> Dict:
        self.forward(x, normalise_weights=normalise_weights)
        return {'beamformer_weights': self.weights.squeeze().cpu().numpy(),
                'directional_weights': self.get_directional_amplitudes().squeeze().cpu().numpy()}
