"""Golden vectors of the auralisation chain from the UNMODIFIED reference (build container only):

    python oracle/gen_golden_auralisation.py   ->  tests/golden/auralisation_cases.npz

`dynamic_rendering_moving_receiver.filter_overlap_add` (src/sound_examples.py:163-226) is called as it is, on a
stand-in object carrying exactly the attributes it reads; the sub-band sum uses scipy.signal.fftconvolve like
src/run_subband_training_treble.py:316-321."""
import os
import sys
import types

import numpy as np
from scipy.signal import fftconvolve

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
ref_shim._stub("pyloudnorm")
import sound_examples  # noqa: E402  (the reference module)

Ref = sound_examples.dynamic_rendering_moving_receiver


def run_reference(stimulus, rirs, hop, fs, fade_ms, alpha):
    num_pos = rirs.shape[0]
    obj = types.SimpleNamespace(sample_rate=fs, num_pos=num_pos, hop_size=hop, total_sim_len=num_pos * hop, rirs=rirs,
                                late_rirs=rirs, stimulus=stimulus, get_fade_windows=Ref.get_fade_windows)
    obj.extended_stimulus = Ref.create_extended_stimulus(obj)
    out = Ref.filter_overlap_add(obj, use_whole_rir=True, alpha=alpha, fade_len_ms=fade_ms)
    return obj.extended_stimulus, out


def main():
    rng = np.random.default_rng(2024)
    fs = 8000.0
    cases = {}
    for name, (num_pos, hop, rir_len, stim_len, fade_ms, alpha) in {
            "a": (12, 80, 300, 500, 5.0, 0.5),      # RIR spans several hops, fade 40 < hop
            "b": (9, 30, 200, 77, 5.0, 1.0),        # fade 40 > hop: the short-tail branch (:221-223), no smoothing
            "c": (5, 64, 50, 1000, 2.0, 0.25),      # RIR shorter than a hop
    }.items():
        decay = np.exp(-np.arange(rir_len) / (0.3 * rir_len))
        rirs = rng.standard_normal((num_pos, rir_len)) * decay
        stim = rng.standard_normal(stim_len).astype(np.float32)
        ext, out = run_reference(stim, rirs, hop, fs, fade_ms, alpha)
        for k, v in dict(rirs=rirs, stimulus=stim, ext=ext, out=out, hop=hop, fs=fs, fade_ms=fade_ms, alpha=alpha).items():
            cases[f"{name}/{k}"] = np.asarray(v)
    # sub-band synthesis: 4 bands, 3 positions
    band_rirs = rng.standard_normal((4, 3, 240)) * np.exp(-np.arange(240) / 60.0)
    firs = rng.standard_normal((4, 33)) * np.hanning(33)
    total = sum(np.stack([fftconvolve(band_rirs[b, p], firs[b], mode='full') for p in range(3)]) for b in range(4))
    cases.update({"sub/band_rirs": band_rirs, "sub/firs": firs, "sub/out": total})
    out_path = os.path.join(os.path.dirname(HERE), "tests", "golden", "auralisation_cases.npz")
    np.savez_compressed(out_path, **cases)
    print("wrote", out_path, {k: v.shape for k, v in cases.items() if k.endswith("/out")})


if __name__ == "__main__":
    main()
