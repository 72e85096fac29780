"""YAML -> pydantic config (same entry points as reference diff_gfdn/config/config_loader.py:11-46)."""
from pathlib import Path
import pickle
from typing import Dict, Type

from pydantic import BaseModel
import yaml


def load_yaml_config(file_path: str):
    with open(Path(file_path).resolve(), 'r', encoding='utf-8') as f:
        return yaml.safe_load(f)


def load_and_validate_config(file_path: str, config_class: Type[BaseModel]):
    return config_class(**load_yaml_config(file_path))


def dump_config_to_pickle(config_data: Dict, output_path: str):
    with open(Path(output_path).resolve(), 'wb') as f:
        pickle.dump(config_data, f)
