"""BASELINE configs[4] render (K6 recursion + listener mix) alone: the 'render' object of bench.py.
usage (GPU box): python scripts/time_render.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from diffgfdn_b200 import _lib, build  # noqa: E402

build.build()
_lib.load()
dev = torch.device("cuda")
hbm, _ = bench.peaks()
print(json.dumps(bench.render_metric(dev, hbm)))
