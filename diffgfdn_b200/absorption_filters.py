"""Delay-line absorption (reference diff_gfdn/absorption_filters.py:40-53)."""
from typing import Sequence, Union

import numpy as np
import torch

from .utils import db2lin


def decay_times_to_gain_per_sample(common_decay_times: Union[float, torch.Tensor],
                                   delay_length_samp: Union[Sequence[int], torch.Tensor], fs: float):
    """gamma_i = 10^(-3 m_i / (fs T60)): the gain of the whole delay line i for a broadband decay time."""
    if isinstance(common_decay_times, torch.Tensor):
        return db2lin(-60 * delay_length_samp / (fs * common_decay_times))
    return db2lin(-60 * np.array(delay_length_samp) / (fs * common_decay_times))


# ----------------------------------------------------------------------------------------------------------------
# Graphic-equaliser absorption filters (init-time host code; reference absorption_filters.py:108-155 with
# filters/geq.py:59-172 and filters/functional.py:220-374). One cascade of (bands + 3) biquads per delay line whose
# magnitude follows the per-band attenuation 10^(-3 m / (fs T60(f))) of that line. The reference finds the section gains
# with 100 LBFGS steps on a linear least-squares objective; the objective is quadratic in the gains, so the minimiser
# is taken directly (float64 least squares, bound handling by a projected active set) -- same design, to the precision
# of the reference's float32 optimiser.
# ----------------------------------------------------------------------------------------------------------------
_GEQ_R = 2.7            # bandwidth factor of the peaking sections: Q = sqrt(R) / (R - 1)
_GEQ_NFFT = 2**16       # grid on which the prototype sections are probed
_GEQ_PROTO_DB = 10.0    # prototype gain; command gains are bounded by twice this value


def _geq_sections(center_hz: np.ndarray, shelving_hz: np.ndarray, gain_db: np.ndarray, fs: float) -> np.ndarray:
    """(bands + 3, 6) biquads [b0 b1 b2 a0 a1 a2]: broadband gain, low shelf, one peaking filter per band, high shelf."""
    g = 10.0**(np.asarray(gain_db, dtype=np.float64).reshape(-1) / 20.0)
    nsec = len(center_hz) + 3
    assert g.size == nsec
    sos = np.zeros((nsec, 6))
    sos[0] = [g[0], 0.0, 0.0, 1.0, 0.0, 0.0]

    def shelf(fc, gain, high):
        t = np.tan(np.pi * fc / fs)
        g2, g4 = gain**0.5, gain**0.25
        b = g2 * np.array([g2 * t * t + np.sqrt(2.0) * t * g4 + 1.0, 2.0 * g2 * t * t - 2.0,
                           g2 * t * t - np.sqrt(2.0) * t * g4 + 1.0])
        a = np.array([g2 + np.sqrt(2.0) * t * g4 + t * t, 2.0 * t * t - 2.0 * g2, g2 - np.sqrt(2.0) * t * g4 + t * t])
        return (a * gain, b) if high else (b, a)

    sos[1, :3], sos[1, 3:] = shelf(shelving_hz[0], g[1], False)
    sos[-1, :3], sos[-1, 3:] = shelf(shelving_hz[1], g[-1], True)
    q = np.sqrt(_GEQ_R) / (_GEQ_R - 1.0)
    for i, fc in enumerate(center_hz):
        gain = g[i + 2]
        w = 2.0 * np.pi * fc / fs
        t = np.tan(w / q / 2.0)
        rg = np.sqrt(gain)
        sos[i + 2] = [rg + gain * t, -2.0 * rg * np.cos(w), rg - gain * t, rg + t, -2.0 * rg * np.cos(w), rg - t]
    return sos


def _interp1(x: np.ndarray, xp: np.ndarray, fp: np.ndarray) -> np.ndarray:
    """Piecewise-linear interpolation, constant beyond the ends (what the reference's 1-D grid interpolator does)."""
    return np.interp(x, xp, fp)


def design_geq(target_gain_db: np.ndarray, center_hz: np.ndarray, shelving_hz: np.ndarray, fs: float) -> np.ndarray:
    """Section gains of the graphic equaliser that matches `target_gain_db` (given at [1 Hz, bands..., fs/2.1]) in the
    least-squares sense on 101 log-spaced control frequencies; returns the (bands + 3, 6) biquads. The reference
    probes its prototype sections in float32 (~5e-3 dB of rounding noise in the interaction matrix of the low shelf
    and the lowest bands) and stops its float32 LBFGS after 100 steps; adjacent sections trade gain against each
    other, so its individual section gains differ from the exact minimiser's by up to ~1 % while the CASCADE responses
    agree to 0.02 dB (tests/test_host_logic_cpu.py)."""
    center_hz = np.asarray(center_hz, dtype=np.float64)
    nsec = len(center_hz) + 3
    control = np.round(np.logspace(0.0, np.log10(fs / 2.1), 101))
    target = _interp1(control, np.concatenate(([1.0], center_hz, [fs / 2.1])), np.asarray(target_gain_db, dtype=np.float64))
    proto = _geq_sections(center_hz, shelving_hz, np.full(nsec, _GEQ_PROTO_DB), fs)
    proto = proto / proto[:, 3:4]
    freqs = np.fft.rfftfreq(_GEQ_NFFT, 1.0 / fs)
    resp = np.fft.rfft(proto[:, :3], _GEQ_NFFT, axis=-1) / (np.fft.rfft(proto[:, 3:], _GEQ_NFFT, axis=-1) + 1e-10)
    mag_db = 20.0 * np.log10(np.abs(resp))
    inter = np.stack([_interp1(control, freqs, mag_db[s]) for s in range(nsec)], axis=1) / _GEQ_PROTO_DB  # (101, nsec)
    hi = np.concatenate(([np.inf], np.full(nsec - 1, 2.0 * _GEQ_PROTO_DB)))
    gains = np.linalg.lstsq(inter, target, rcond=None)[0]
    free = np.ones(nsec, dtype=bool)
    for _ in range(nsec):  # bounded least squares: clamp the violators, re-solve for the rest
        bad = free & (np.abs(gains) > hi)
        if not bad.any():
            break
        gains[bad] = np.sign(gains[bad]) * hi[bad]
        free &= ~bad
        if free.any():
            rhs = target - inter[:, ~free] @ gains[~free]
            gains[free] = np.linalg.lstsq(inter[:, free], rhs, rcond=None)[0]
    return _geq_sections(center_hz, shelving_hz, gains, fs)


def decay_times_to_gain_filters_geq(band_centre_hz: Sequence[float], common_decay_times: Sequence[float],
                                    delay_length_samp: Sequence[int], fs: float) -> torch.Tensor:
    """(bands + 3, num_delays, 3, 2) float32: [..., 0] numerators, [..., 1] denominators of the absorption cascade of
    every delay line of ONE group (reference absorption_filters.py:108-155; the layout the model reshapes to
    (N, bands + 3, 3, 2), model.py:131-147)."""
    bands = np.asarray(band_centre_hz, dtype=np.float64)
    t60 = np.asarray(common_decay_times, dtype=np.float64).reshape(-1)
    delays = np.asarray(delay_length_samp, dtype=np.float64).reshape(-1)
    shelving = np.array([bands[0] / np.sqrt(2.0), bands[-1] * np.sqrt(2.0)])
    per_sample = 10.0**(-3.0 / fs / t60)
    lin = per_sample[:, None]**delays[None, :]                       # (bands, delays)
    lin = np.concatenate([0.5 * lin[:1], lin, 0.5 * lin[-1:]], axis=0)   # the shelves aim at half the edge gains
    eps = float(np.finfo(np.float32).eps)
    out = np.zeros((len(bands) + 3, len(delays), 3, 2))
    for i in range(len(delays)):
        target_db = np.clip(20.0 * np.log10(np.abs(lin[:, i]) + eps), -200.0, None)
        sos = design_geq(target_db, bands, shelving, fs)
        out[:, i, :, 0] = sos[:, :3]
        out[:, i, :, 1] = sos[:, 3:]
    return torch.tensor(out, dtype=torch.float32)
