"""Build libdiffgfdn_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m diffgfdn_b200.build [--force]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with the snapshot."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdiffgfdn_b200.so")
SOURCES = ["common.cu", "expm.cu", "solve.cu", "project.cu", "czt.cu", "edc.cu", "edc_td.cu", "edc_td_fused.cu", "colorless.cu", "render.cu"]


def _nvcc():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "diffgfdn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64")
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", INCLUDE, "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", LIB_PATH, "-L", cuda_lib, "-lcufft", "-lcudart", "-Xlinker", f"-rpath,{cuda_lib}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
