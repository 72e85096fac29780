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
SOURCES = ["common.cu", "expm.cu", "solve.cu", "project.cu", "czt.cu", "edc.cu", "edc_td.cu", "edc_td_fused.cu", "edc_td_sliced.cu", "colorless.cu", "render.cu",
           "mlp.cu", "svf.cu", "edr.cu", "assemble.cu", "peer.cu"]


def _nvcc():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(INCLUDE, "diffgfdn_b200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                               if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per .cu (compiled in parallel, only when its source or a header changed), then one link."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "lib64")
    base = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
            "-fPIC,-fvisibility=hidden", "-I", INCLUDE, "-I", CSRC]
    if verbose:
        base += ["-Xptxas", "-v"]
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(obj_dir, s[:-3] + ".o")
        if force or _obj_stale(src, obj):
            jobs.append((s, base + ["-c", src, "-o", obj]))

    def run(job):
        name, cmd = job
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n" + res.stdout + res.stderr)
        return name, res.stdout + res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            for name, log in pool.map(run, jobs):
                if verbose:
                    print(f"== {name}\n{log}")
    objs = [os.path.join(obj_dir, s[:-3] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [_nvcc(), "--shared", "-o", LIB_PATH] + objs + ["-L", cuda_lib, "-lcufft", "-lcudart", "-Xlinker",
                                                             f"-rpath,{cuda_lib}"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
