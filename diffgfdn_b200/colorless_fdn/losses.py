"""Colorless-FDN losses with the reference's callables (diff_gfdn/colorless_fdn/losses.py:7-73).

`mse_loss` / `amse_loss` on a 1-D complex response against ones run in one fused sm_100a kernel (forward and
backward); other shapes / targets fall back to the same formula in torch ops on the device (never on the CPU)."""
import numpy as np
import torch
from torch import nn

from .. import ops


class sparsity_loss(nn.Module):
    """-(sum|A| - N sqrt(N)) / (N (sqrt(N) - 1)) on an (N, N) orthogonal matrix -- O(N^2), stays in torch."""

    def forward(self, A: torch.Tensor):
        N = A.shape[-1]
        return -(torch.sum(torch.abs(A)) - (N * np.sqrt(N))) / (N * (np.sqrt(N) - 1))


def _is_ones(y_true: torch.Tensor) -> bool:
    return bool(torch.all(y_true == 1))


class mse_loss(nn.Module):
    """mean_k (|y_pred| - |y_true|)^2"""
    asym = False

    def forward(self, y_pred: torch.Tensor, y_true: torch.Tensor = None, assume_unit_target: bool = False):
        if y_pred.is_complex() and y_pred.ndim == 1 and (y_true is None or assume_unit_target or _is_ones(y_true)):
            return ops.colorless_loss_per_group(y_pred.unsqueeze(-1), self.asym)[0]
        diff = torch.abs(y_pred) - torch.abs(y_true)
        if self.asym:
            g = 2.0 + 2.0 * (diff > 1).to(diff.dtype)
            loss = torch.mean(torch.pow(diff, g), dim=0)
        else:
            loss = torch.mean(diff**2, dim=0 if y_pred.ndim > 1 else -1)
        return torch.mean(loss) if y_pred.ndim > 1 else loss


class amse_loss(mse_loss):
    """Asymmetric version: exponent 4 where |y_pred| - |y_true| > 1 (reference colorless_fdn/losses.py:44-73)."""
    asym = True
