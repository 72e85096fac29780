"""Config schema of the DiffGFDN hot path: accepts the reference's YAML files verbatim.

Field names, defaults and validation mirror the reference schema (diff_gfdn/config/config.py:17-282 and
spatial_sampling/config.py:9-14) so that every shipped `data/config/**/*.yml` and the dict built by
`run_subband_training_treble.create_config` validate unchanged (`extra="forbid"` is kept).

Deliberate difference (documented in DESIGN.md): the reference's `device` validator returns None for every input
(config.py:163-170, quirk Q10), which silently trains on CPU. Here `device` is kept as given; the runner maps
'gpu'/'cuda' to `cuda:{LOCAL_RANK}` and there is no CPU execution path for the kernels.
"""
from enum import Enum
from typing import List, Optional, Tuple

import numpy as np
from pydantic import BaseModel, ConfigDict, Field, computed_field, model_validator


class BeamformerType(Enum):
    BUTTER = 'butterworth'
    MAX_DI = 'max_directivity'
    MAX_RE = 'max_re'


class CouplingMatrixType(Enum):
    SCALAR = "scalar_matrix"
    FILTER = "filter_matrix"
    RANDOM = "random_matrix"

    def __repr__(self) -> str:
        return str(self.value)


class FeatureEncodingType(Enum):
    SINE = "sinusoidal"
    MESHGRID = "meshgrid"

    def __repr__(self) -> str:
        return str(self.value)


class FeedbackLoopConfig(BaseModel):
    pu_matrix_order: int = 2**5
    coupling_matrix_type: CouplingMatrixType = CouplingMatrixType.SCALAR
    use_zero_coupling: bool = True


class MLPTuningConfig(BaseModel):
    tune_hyperparameters: bool = True
    min_layers: int = 1
    max_layers: int = 20
    min_neurons: int = 2**4
    max_neurons: int = 2**7
    step_size: int = 2**4
    num_trials: int = 50


class SubbandProcessingConfig(BaseModel):
    centre_frequency: float
    frequency_range: Tuple
    num_fraction_octaves: int = 3
    use_amp_preserving_filterbank: bool = True


class OutputFilterConfig(BaseModel):
    use_svfs: bool = True
    compress_pole_factor: float = 1.0
    mlp_tuning_config: Optional[MLPTuningConfig] = None
    num_hidden_layers: int = 3
    num_neurons_per_layer: int = 2**7
    num_fourier_features: int = 10
    encoding_type: FeatureEncodingType = FeatureEncodingType.SINE
    beamformer_type: Optional[BeamformerType] = None
    use_skip_connections: bool = False


class DecayFilterConfig(BaseModel):
    use_absorption_filters: bool = True
    learn_common_decay_times: bool = False
    initialise_with_opt_values: bool = True


class TestSetConfig(BaseModel):
    __test__ = False  # not a pytest class
    seed: int = 4314
    ratio: float = 0.1


class TrainerConfig(BaseModel):
    batch_size: int = 32
    num_freq_bins: Optional[int] = None
    device: Optional[str] = 'cpu'
    train_valid_split: Optional[float] = 0.8
    hold_out_test_set: Optional[TestSetConfig] = None
    grid_resolution_m: Optional[float] = None
    max_epochs: int = 5
    lr: float = 0.01
    io_lr: float = 0.01
    coupling_angle_lr: float = 0.01
    output_filt_ir_len_ms: float = 500
    use_reg_loss: bool = False
    use_erb_edr_loss: bool = False
    use_colorless_loss: bool = False
    use_asym_spectral_loss: bool = False
    edc_loss_weight: float = 1.0
    edr_loss_weight: float = 1.0
    spectral_loss_weight: float = 1.0
    sparsity_loss_weight: float = 1.0
    use_edc_mask: bool = False
    use_frequency_weighting: bool = False
    subband_process_config: Optional[SubbandProcessingConfig] = None
    train_dir: str = "output/cpu/"
    ir_dir: str = "audio/cpu/"
    save_true_irs: bool = False
    alias_attenuation_db: Optional[int] = None
    reduced_pole_radius: float = Field(default=1.0)

    @model_validator(mode='after')
    def calculate_reduced_pole_radius(self):
        """reference config.py:172-182"""
        if self.alias_attenuation_db is not None and self.num_freq_bins is not None:
            self.reduced_pole_radius = 10**(-abs(self.alias_attenuation_db) / self.num_freq_bins / 20)
        return self


class ColorlessFDNConfig(BaseModel):
    use_colorless_prototype: bool = False
    batch_size: int = 2000
    max_epochs: int = 20
    train_valid_split: float = 0.8
    lr: float = 0.01
    alpha: float = 1
    saved_param_path: Optional[str] = None

    @computed_field
    @property
    def load_fixed_parameters(self) -> bool:
        return self.saved_param_path is not None


def _primes_in(lo: int, hi: int) -> np.ndarray:
    """Primes p with lo <= p < hi (what sympy.primerange yields)."""
    sieve = np.ones(max(hi, 2), dtype=bool)
    sieve[:2] = False
    for i in range(2, int(hi**0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = False
    idx = np.nonzero(sieve)[0]
    return idx[idx >= lo].astype(np.int32)


def _next_prime(n: int) -> int:
    """Smallest prime strictly greater than n (sympy.nextprime)."""
    c = int(n) + 1
    while True:
        if c >= 2 and all(c % q for q in range(2, int(c**0.5) + 1)):
            return c
        c += 1


class DiffGFDNConfig(BaseModel):
    seed: int = 46434
    room_dataset_path: str = 'resources/Georg_3room_FDTD/srirs.pkl'
    num_groups: int = 3
    ir_path: Optional[str] = None
    sample_rate: float = 32000.0
    trainer_config: TrainerConfig = TrainerConfig()
    delay_range_ms: List[float] = [20.0, 50.0]
    ambi_order: Optional[int] = None
    num_delay_lines: Optional[int] = 12
    feedback_loop_config: FeedbackLoopConfig = FeedbackLoopConfig()
    decay_filter_config: DecayFilterConfig = DecayFilterConfig()
    output_filter_config: OutputFilterConfig = OutputFilterConfig()
    input_filter_config: Optional[OutputFilterConfig] = OutputFilterConfig()
    colorless_fdn_config: ColorlessFDNConfig = ColorlessFDNConfig()

    model_config = ConfigDict(extra="forbid")

    @model_validator(mode="after")
    def set_num_delay_lines(self):
        """reference config.py:242-247"""
        if self.ambi_order is not None:
            self.num_delay_lines = ((self.ambi_order + 1)**2) * self.num_groups
        return self

    @model_validator(mode='after')
    def set_train_valid_ratio(self):
        """reference config.py:250-260"""
        if self.trainer_config.grid_resolution_m is not None:
            if self.ambi_order is None:
                raise AttributeError("Only use grid resolution for directional reverberation training!")
            self.trainer_config.train_valid_split = None
        return self

    @computed_field
    @property
    def delay_length_samps(self) -> List[int]:
        """Co-prime delay lengths: a seeded permutation of the primes in the delay range plus the first prime
        above it (reference config.py:262-279; same numpy RNG calls, so the same delays for the same seed)."""
        rng_ms = np.asarray(self.delay_range_ms)
        lo, hi = (rng_ms * 1e-3 * self.sample_rate).astype(np.int32)
        primes = _primes_in(int(lo), int(hi))
        np.random.seed(self.seed)
        shuffled = primes[np.random.permutation(len(primes))]
        return np.array(np.r_[shuffled[:self.num_delay_lines - 1], _next_prime(int(hi))], dtype=np.int32).tolist()
