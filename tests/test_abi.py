"""CPU checks of the C-ABI boundary: the library builds, loads and exports exactly what include/*.h declares,
and the ctypes table mirrors the header. No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "diffgfdn_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"DGFDN_API[^;(]*?\b(dgfdn_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 20
    for must in ("dgfdn_solve_fwd", "dgfdn_solve_bwd", "dgfdn_project_fwd", "dgfdn_project_bwd",
                 "dgfdn_irfft_window_fwd", "dgfdn_irfft_window_bwd", "dgfdn_edc_loss_fwd", "dgfdn_edc_loss_bwd",
                 "dgfdn_render_groups", "dgfdn_render_mix"):
        assert must in syms


def test_library_builds_loads_and_exports_every_symbol():
    from diffgfdn_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header():
    from diffgfdn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    src = open(HEADER).read()
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"DGFDN_API[^;(]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", src, re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params == "void" else len([p for p in params.split(",") if p.strip()])
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table has {len(args)}"


def test_ops_reject_cpu_tensors():
    import torch
    from diffgfdn_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.receiver_project(torch.zeros(2, 3), torch.zeros(8, 3, dtype=torch.complex64), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.edc_db(torch.zeros(2, 16))
