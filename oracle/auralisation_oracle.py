"""CPU oracle of the reference's inference / auralisation chain -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement (loops, float64) of what the reference does with a trained model at inference time:

  * sub-band synthesis (src/run_subband_training_treble.py:316-358): the RIR of every octave-band model is convolved
    with that band's FIR of the amplitude-preserving filterbank (`fftconvolve(h, fir, mode='full')`, :318-321) and the
    filtered band RIRs of a position are added (`sum_arrays`, :355-358);
  * moving listener (src/sound_examples.py:163-226 `dynamic_rendering_moving_receiver.filter_overlap_add`, fades
    :118-127, stimulus tiling :149-161): the stimulus is cut into hops of `update_ms`, hop k is convolved with the
    (recursively smoothed, :189-192) RIR of the k-th position, and the convolution outputs are overlap-added; the
    first `fade_len` samples a block adds are `prev_tail * fade_out + head * fade_in`, where prev_tail is the LAST
    `fade_len` samples of what the previous block added (:204-223 -- the reference's own definition, kept as is).

Pinned against the unmodified reference: `oracle/gen_golden_auralisation.py` calls the reference's
`filter_overlap_add` itself (imported through `oracle/ref_shim.py`, with a stand-in `self`) and SciPy's
`fftconvolve` on seeded inputs and freezes the outputs in `tests/golden/auralisation_*.npz`;
`tests/test_oracle_golden.py` checks this file against them. Only tests/, smoke() and bench.py's CPU legs may
import this module."""
import numpy as np


def full_convolve(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """scipy.signal.fftconvolve(x, h, mode='full') on the last axis, as a direct sum (float64)."""
    x = np.asarray(x, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    out = np.zeros(x.shape[:-1] + (x.shape[-1] + h.shape[-1] - 1, ))
    for i in range(h.shape[-1]):
        out[..., i:i + x.shape[-1]] += x * h[..., i:i + 1]
    return out


def subband_sum_rir(band_rirs: np.ndarray, band_firs: np.ndarray) -> np.ndarray:
    """run_subband_training_treble.py:316-321, 355-358: sum_b fftconvolve(h_b, fir_b, 'full').
    band_rirs (bands, ..., T), band_firs (bands, L) -> (..., T + L - 1)."""
    acc = 0.0
    for b in range(band_rirs.shape[0]):
        acc = acc + full_convolve(band_rirs[b], band_firs[b])
    return acc


def fade_windows(win_len: int, fade_out: bool) -> np.ndarray:
    """sound_examples.py:118-127 (linear, correlated fades)."""
    n = np.linspace(start=-1, stop=1, num=win_len)
    return 0.5 * (1 + (1 - 2 * int(fade_out)) * n)


def extend_stimulus(stimulus: np.ndarray, total_len: int) -> np.ndarray:
    """sound_examples.py:149-161: the stimulus repeated up to num_pos * hop samples (float32 buffer)."""
    out = np.zeros(total_len, dtype=np.float32)
    n = len(stimulus)
    for rep in range(int(np.ceil(total_len / n))):
        lo, hi = rep * n, min((rep + 1) * n, total_len)
        out[lo:hi] = stimulus[:hi - lo]
    return out


def filter_overlap_add(stimulus_ext: np.ndarray, rirs: np.ndarray, hop: int, fade_len: int, alpha: float = 0.5) -> np.ndarray:
    """sound_examples.py:163-226. stimulus_ext (num_pos * hop,), rirs (num_pos, L) -> (num_pos * hop,)."""
    num_pos = rirs.shape[0]
    total = len(stimulus_ext)
    out = np.zeros(total, dtype=np.float64)
    f_out, f_in = fade_windows(fade_len, True), fade_windows(fade_len, False)
    prev_tail = np.zeros(fade_len)
    prev_filter = None
    for k in range(num_pos):
        lo, hi = k * hop, min((k + 1) * hop, total)
        cur = np.asarray(rirs[k], dtype=np.float64)
        if prev_filter is not None:  # :189-192
            cur = alpha * cur + (1 - alpha) * prev_filter
        prev_filter = cur
        y = full_convolve(np.asarray(stimulus_ext[lo:hi], dtype=np.float64), cur)
        end = min(lo + len(y), total)
        y = y[:end - lo]
        if k > 0:  # :204-214
            ol = min(fade_len, len(y))
            out[lo:lo + ol] += prev_tail[:ol] * f_out[:ol] + y[:ol] * f_in[:ol]
            out[lo + ol:end] += y[ol:]
        else:
            out[lo:end] += y
        if len(y) >= fade_len:  # :217-223
            prev_tail[:fade_len] = y[-fade_len:]
        else:
            prev_tail[:len(y)] = y
    return out


def convert_srir_to_brir(srirs: np.ndarray, hrir_sh: np.ndarray, rotations: np.ndarray) -> np.ndarray:
    """sofa_parser.py:452-505 with the two third-party inputs made explicit: hrir_sh = hrtf_reader.
    get_spherical_harmonic_representation(order) (C, 2, Th) and rotations[o] = spaudiopy sh_rotation_matrix(order,
    -azimuth_o, -elevation_o, 0, 'real') (:488-493). Same loops and einsum as the reference. Returns (R, O, nfft, 2)."""
    num_receivers = srirs.shape[0]
    nfft = 2**int(np.ceil(np.log2(srirs.shape[-1])))  #                        :466
    ambi_rtfs = np.fft.rfft(srirs, nfft, axis=-1)  #                            :470
    ambi_hrtfs = np.fft.rfft(hrir_sh, n=nfft, axis=-1)  #                       :473
    brirs = np.zeros((num_receivers, rotations.shape[0], nfft, 2))
    for r in range(num_receivers):
        cur = ambi_rtfs[r]  # (C, F)
        for o in range(rotations.shape[0]):
            rotated = cur.T @ rotations[o].T  #                                 :495  (F, C)
            brtf = np.einsum('nrf, fn -> fr', np.conj(ambi_hrtfs), rotated)  #  :498
            brirs[r, o] = np.fft.irfft(brtf, n=nfft, axis=0)  #                 :501
    return brirs
