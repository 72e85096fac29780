"""Position -> gain networks (reference diff_gfdn/dnn.py). Module and parameter names match the reference so that
its checkpoints load (`...mlp.model.{0,1,3,4,...}` and the skip-connection layout). On CUDA the whole chain
encoding -> MLP -> final activation runs in the K7 kernels (`fused_position_mlp`, csrc/mlp.cu); the nn.Module
`forward`s below are the same arithmetic in stock PyTorch, kept for shapes K7 does not cover (neurons not in
{64, 128}) and as the float32 reference of the kernel tests."""
from typing import Optional

import numpy as np
import torch
from torch import nn
from torch.nn import init

from . import ops


class Sigmoid(nn.Module):

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return 1.0 / (1 + torch.exp(-x))


class ScaledSigmoid(nn.Module):
    """lower + (upper - lower) * sigmoid(x)  (reference dnn.py:21-36)."""

    def __init__(self, lower_limit: float, upper_limit: float):
        super().__init__()
        self.lower_limit = lower_limit
        self.upper_limit = upper_limit
        self.sigmoid = Sigmoid()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.lower_limit + (self.upper_limit - self.lower_limit) * self.sigmoid(x)


class SinusoidalEncoding(nn.Module):
    """[sin(f_i pi p), cos(f_i pi p)]_i, f = exp(linspace(ln 1, ln 32, F)), float32 output of width 6F
    (reference dnn.py:89-126). Vectorised: one broadcasted multiply instead of a Python loop over features."""

    def __init__(self, num_fourier_features: int):
        super().__init__()
        self.num_fourier_features = num_fourier_features

    def frequencies(self, device, dtype) -> torch.Tensor:
        """f_i pi in the position dtype (float32 table, then cast -- like the reference's broadcast)."""
        key = (str(device), dtype)
        if getattr(self, "_freq_key", None) != key:
            f = torch.exp(torch.linspace(np.log(1.0), np.log(32.0), self.num_fourier_features, device=device))
            self._freq, self._freq_key = (f * np.pi).to(dtype), key
        return self._freq

    def forward(self, pos_coords: torch.Tensor) -> torch.Tensor:
        arg = self.frequencies(pos_coords.device, pos_coords.dtype).view(1, -1, 1) * pos_coords.unsqueeze(1)  # (P, F, 3)
        enc = torch.cat((torch.sin(arg), torch.cos(arg)), dim=-1)  # (P, F, 6): [sin xyz, cos xyz] per feature
        return enc.reshape(pos_coords.shape[0], -1).to(torch.float32)


def _init_linear(module: nn.Module):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            init.kaiming_uniform_(m.weight, nonlinearity='relu')
            if m.bias is not None:
                init.constant_(m.bias, 0)


class MLP(nn.Module):
    """Linear -> LayerNorm -> ReLU, (1 + hidden) times, then Linear (reference dnn.py:331-400)."""

    def __init__(self, num_pos_features: int, num_hidden_layers: int, num_neurons: int, num_groups: int,
                 num_biquads_in_cascade: int, num_params: int):
        super().__init__()
        self.num_biquads = num_biquads_in_cascade
        self.num_groups = num_groups
        self.num_params = num_params
        layers = [nn.Linear(num_pos_features, num_neurons), nn.LayerNorm(num_neurons), nn.ReLU()]
        for _ in range(num_hidden_layers):
            layers += [nn.Linear(num_neurons, num_neurons), nn.LayerNorm(num_neurons), nn.ReLU()]
        layers.append(nn.Linear(num_neurons, num_groups * num_params * num_biquads_in_cascade))
        self.model = nn.Sequential(*layers)
        _init_linear(self.model)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x).view(x.shape[0], self.num_groups, self.num_biquads, self.num_params)


class ResidualBlock(nn.Module):

    def __init__(self, num_neurons: int):
        super().__init__()
        self.linear = nn.Linear(num_neurons, num_neurons)
        self.norm = nn.LayerNorm(num_neurons)
        self.activation = nn.ReLU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.activation(self.norm(self.linear(x))) + x


class MLP_SkipConnections(nn.Module):
    """MLP with residual hidden blocks (reference dnn.py:284-328)."""

    def __init__(self, num_pos_features: int, num_hidden_layers: int, num_neurons: int, num_groups: int,
                 num_biquads_in_cascade: int, num_params: int):
        super().__init__()
        self.num_biquads = num_biquads_in_cascade
        self.num_groups = num_groups
        self.num_params = num_params
        self.input_layer = nn.Sequential(nn.Linear(num_pos_features, num_neurons), nn.LayerNorm(num_neurons),
                                         nn.ReLU())
        self.hidden_layers = nn.ModuleList([ResidualBlock(num_neurons) for _ in range(num_hidden_layers)])
        self.output_layer = nn.Linear(num_neurons, num_groups * num_params * num_biquads_in_cascade)
        _init_linear(self)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b = x.shape[0]
        x = self.input_layer(x)
        for layer in self.hidden_layers:
            x = layer(x)
        return self.output_layer(x).view(b, self.num_groups, self.num_biquads, self.num_params)


def fused_position_mlp(encoder: SinusoidalEncoding, mlp: nn.Module, position: torch.Tensor, final_act: int = 0,
                       lo: float = 0.0, hi: float = 1.0) -> Optional[torch.Tensor]:
    """final_act(mlp(encoder(position))) as (P, out_dim) float32 through the K7 kernels, or None when the shape is
    outside their range (the caller then uses the nn.Module path)."""
    if not position.is_cuda or position.dtype not in (torch.float32, torch.float64):
        return None
    if isinstance(mlp, MLP):
        mods = list(mlp.model)
        lins = [m for m in mods if isinstance(m, nn.Linear)]
        norms = [m for m in mods if isinstance(m, nn.LayerNorm)]
        residual = False
    elif isinstance(mlp, MLP_SkipConnections):
        lins = [mlp.input_layer[0]] + [blk.linear for blk in mlp.hidden_layers] + [mlp.output_layer]
        norms = [mlp.input_layer[1]] + [blk.norm for blk in mlp.hidden_layers]
        residual = True
    else:
        return None
    neurons, in_dim = lins[0].weight.shape
    if not ops.mlp_supported(in_dim, encoder.num_fourier_features, neurons, len(norms), lins[-1].weight.shape[0]):
        return None
    if any(abs(n.eps - 1e-5) > 0 for n in norms):
        return None
    params = []
    for lin, norm in zip(lins[:-1], norms):
        params += [lin.weight, lin.bias, norm.weight, norm.bias]
    params += [lins[-1].weight, lins[-1].bias]
    return ops.position_mlp(position.contiguous(), encoder.frequencies(position.device, position.dtype), params,
                            residual=residual, final_act=final_act, lo=lo, hi=hi)
