"""Small DSP helpers with the reference's names and semantics (diff_gfdn/utils.py:16-179)."""
from typing import Dict, Optional, Union

import numpy as np
import torch
from torch import nn

EPS_F32 = float(torch.finfo(torch.float32).eps)


def db(x, is_squared: bool = False, min_value: float = -200):
    """factor * log10(|x| + eps_f32), clipped from below (reference utils.py:16-40)."""
    factor = 10.0 if is_squared else 20.0
    if torch.is_tensor(x):
        return (factor * torch.log10(torch.abs(x) + EPS_F32)).clip(min=min_value)
    return (factor * np.log10(np.abs(x) + np.finfo(np.float32).eps)).clip(min=min_value)


def db2lin(x, is_squared: bool = False):
    e = 0.1 if is_squared else 0.05
    if torch.is_tensor(x):
        return torch.pow(10.0, x * e)
    return np.power(10.0, x * e)


def ms_to_samps(ms, fs: float):
    """reference utils.py:62-80 (truncating int conversion)."""
    if isinstance(ms, torch.Tensor):
        return (ms * 1e-3 * torch.tensor(fs)).int()
    samp = ms * 1e-3 * fs
    return int(samp) if np.isscalar(samp) else samp.astype(np.int32)


def samps_to_ms(samps, fs: float):
    if isinstance(samps, torch.Tensor):
        return samps.float() / torch.tensor(fs) * 1e3
    return float(samps) / fs * 1e3


def get_frequency_samples(num: int, device: Optional[torch.device] = None):
    """num points on the upper unit semicircle (reference utils.py:128-141)."""
    angle = torch.linspace(0, 1, steps=num, device=device)
    return torch.polar(torch.ones(num, device=device), angle * np.pi)


def unit_circle_grid(nfft: int, radius: float = 1.0, device=None) -> torch.Tensor:
    """z_k = r exp(j 2 pi k / nfft), k = 0..nfft/2, complex128 (reference dataloader.py:552-566)."""
    w = torch.as_tensor(np.fft.rfftfreq(nfft) * 2.0 * np.pi, dtype=torch.float64, device=device)
    return torch.polar(torch.full_like(w, radius), w)


def to_complex(x: torch.Tensor):
    return torch.complex(x, torch.zeros_like(x))


@torch.no_grad()
def get_response(x: Union[Dict, torch.Tensor], net: nn.Module, output_scalars: Optional[torch.Tensor] = None):
    """Forward pass + impulse response h = irfft(H) (reference utils.py:149-179). The irfft here is the plain
    power-of-two-friendly default-n transform (cuFFT through torch.fft); it is an inference-side helper."""
    if getattr(net, "use_colorless_loss", False):
        H, H_sub = net(x, output_scalars) if output_scalars is not None else net(x)
        return H, H_sub, torch.fft.irfft(H, dim=-1)
    H = net(x)
    return H, torch.fft.irfft(H, dim=-1)


class TensorKeyedCache:
    """Small FIFO cache of values derived from constant input tensors (targets, z grids), keyed by storage identity.

    A key made of data_ptr alone is NOT an identity: once a batch tensor is freed the caching allocator hands the same
    address to the next batch of the same shape, and a stale entry would be returned for different data (a loader
    that gathers a fresh batch every step hits this within a few steps). Every entry therefore keeps a reference to
    its source tensor: while the entry lives the address cannot be reused, so pointer + shape + version is exact."""

    def __init__(self, max_entries: int = 16):
        self.max_entries = max_entries
        self._entries = {}

    @staticmethod
    def _key(t: torch.Tensor, extra=None):
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version, str(t.device), extra)

    def get(self, t: torch.Tensor, extra=None):
        hit = self._entries.get(self._key(t, extra))
        return None if hit is None else hit[1]

    def put(self, t: torch.Tensor, value, extra=None):
        while len(self._entries) >= self.max_entries:
            self._entries.pop(next(iter(self._entries)))
        self._entries[self._key(t, extra)] = (t, value)
        return value

    def clear(self):
        self._entries.clear()
