"""Import shim for the UNMODIFIED reference (orchidas/DiffGFDN) -- TEST INFRASTRUCTURE ONLY.

This file is used only in the build container (where /root/reference exists) by
`oracle/gen_golden.py` to produce the golden fixtures under `tests/golden/` and by
`tests/test_oracle_vs_reference.py` (auto-skipped when /root/reference is absent).
Nothing in `diffgfdn_b200/` may import it.

It stubs third-party packages the reference imports at module level but that are not
installed here (matplotlib, pyfar, librosa, spaudiopy, ...) and patches one pydantic
incompatibility in `spatial_sampling/config.py:43-44`.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("DIFFGFDN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "diff_gfdn"))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def _stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            mod = _Stub(n)
            mod.__path__ = []
            sys.modules[n] = mod
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], mod)


_installed = False


def install():
    """Make `import diff_gfdn...` resolve to the reference sources."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    for n in ["matplotlib", "matplotlib.pyplot", "IPython", "IPython.display", "pyfar", "pyfar.dsp",
              "pyfar.dsp.filter", "librosa", "librosa.filters", "spaudiopy", "spaudiopy.sph", "soundfile",
              "optuna", "sofar", "h5py", "DecayFitNet", "DecayFitNet.python", "DecayFitNet.python.toolbox",
              "DecayFitNet.python.toolbox.DecayFitNetToolbox", "DecayFitNet.python.toolbox.core",
              "DecayFitNet.python.toolbox.utils"]:
        try:
            __import__(n)
        except Exception:
            _stub(n)
    sys.path[:0] = [os.path.join(REF_ROOT, "src"), os.path.join(REF_ROOT, "submodules", "slope2noise")]
    src = open(os.path.join(REF_ROOT, "src", "spatial_sampling", "config.py")).read()
    src = src.replace("Optional[MLPConfig()]", "Optional[MLPConfig]").replace("Optional[CNNConfig()]",
                                                                              "Optional[CNNConfig]")
    pkg = types.ModuleType("spatial_sampling")
    pkg.__path__ = [os.path.join(REF_ROOT, "src", "spatial_sampling")]
    sys.modules["spatial_sampling"] = pkg
    cfg = types.ModuleType("spatial_sampling.config")
    sys.modules["spatial_sampling.config"] = cfg
    exec(compile(src, "spatial_sampling/config.py", "exec"), cfg.__dict__)
    _installed = True
