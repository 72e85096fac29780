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
