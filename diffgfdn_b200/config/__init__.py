from .config import (BeamformerType, ColorlessFDNConfig, CouplingMatrixType, DecayFilterConfig, DiffGFDNConfig,
                     FeatureEncodingType, FeedbackLoopConfig, MLPTuningConfig, OutputFilterConfig,
                     SubbandProcessingConfig, TestSetConfig, TrainerConfig)
from .config_loader import dump_config_to_pickle, load_and_validate_config, load_yaml_config

__all__ = [
    "BeamformerType", "ColorlessFDNConfig", "CouplingMatrixType", "DecayFilterConfig", "DiffGFDNConfig",
    "FeatureEncodingType", "FeedbackLoopConfig", "MLPTuningConfig", "OutputFilterConfig", "SubbandProcessingConfig",
    "TestSetConfig", "TrainerConfig", "dump_config_to_pickle", "load_and_validate_config", "load_yaml_config"
]
