"""CPU checks of host-side logic that needs no GPU: the reference-holding tensor cache, the batched Givens product
against the oracle, the SVF coefficient formulas against the oracle, the config schema of the shipped YAMLs."""
import numpy as np
import torch

from oracle import gfdn_oracle as O


def test_tensor_keyed_cache_holds_sources_and_evicts_fifo():
    from diffgfdn_b200.utils import TensorKeyedCache
    cache = TensorKeyedCache(max_entries=3)
    keep = []
    for i in range(3):
        t = torch.full((4, ), float(i))
        keep.append(t.data_ptr())
        cache.put(t, i)
        del t  # the cache keeps the tensor alive: a new tensor cannot get its address
    fresh = torch.full((4, ), 9.0)
    assert fresh.data_ptr() not in keep
    assert cache.get(fresh) is None
    a = torch.zeros(4)
    cache.put(a, "a")  # fourth entry: the first one is evicted
    assert len(cache._entries) == 3 and cache.get(a) == "a"
    a.add_(1.0)  # in-place change bumps the version: the stale value is not served
    assert cache.get(a) is None
    b = torch.zeros(8)[::2]
    cache.put(b, "strided", extra=("x", 1))
    assert cache.get(b, extra=("x", 1)) == "strided" and cache.get(b) is None


def test_batched_givens_product_matches_oracle_and_is_orthogonal():
    from diffgfdn_b200.feedback_loop import ND_Unitary
    torch.manual_seed(0)
    for n in (1, 2, 3, 5, 8):
        alpha = (torch.rand(n * (n - 1) // 2, dtype=torch.float64) * 2 - 1).requires_grad_(True)
        u = ND_Unitary()(alpha, n)
        assert float((u - O.nd_unitary(alpha.detach(), n)).abs().max()) < 1e-14
        assert float((u @ u.t() - torch.eye(n, dtype=torch.float64)).abs().max()) < 1e-13
        if n > 1:
            u.sum().backward()
            assert torch.isfinite(alpha.grad).all()


def test_svf_coefficients_match_oracle():
    from diffgfdn_b200.gain_filters import svf_cutoff_frequencies, svf_to_biquads
    torch.manual_seed(1)
    svf = torch.stack([torch.rand(4, 3, 11) * 0.98 + 1e-6, torch.rand(4, 3, 11) * 12 - 6], dim=-1).to(torch.float32)
    for fs, pf in ((32000.0, 1.0), (48000.0, 0.997)):
        cut = svf_cutoff_frequencies(fs)
        assert float((cut - O.svf_cutoffs(fs)).abs().max()) < 1e-15 and cut.numel() == 11
        got = svf_to_biquads(svf, cut, pf)
        want = O.svf_to_biquads(svf, O.svf_cutoffs(fs), pf)
        assert got.dtype == torch.float64 and float((got - want).abs().max()) < 1e-13


def test_shipped_yaml_shapes_validate():
    """The dictionaries run_model.py / run_subband_training_treble.py build must pass the (extra='forbid') schema."""
    from diffgfdn_b200.config import DiffGFDNConfig
    cfg = DiffGFDNConfig(**{
        "sample_rate": 32000.0, "num_delay_lines": 12, "num_groups": 3,
        "decay_filter_config": {"use_absorption_filters": False},
        "feedback_loop_config": {"coupling_matrix_type": "scalar_matrix", "use_zero_coupling": False},
        "output_filter_config": {"use_svfs": True, "num_hidden_layers": 3, "num_neurons_per_layer": 128,
                                 "num_fourier_features": 20, "compress_pole_factor": 0.998},
        "trainer_config": {"max_epochs": 5, "batch_size": 32, "use_colorless_loss": True, "num_freq_bins": 131072,
                           "train_dir": "out/", "ir_dir": "ir/"}})
    assert len(cfg.delay_length_samps) == 12 and cfg.output_filter_config.use_svfs
    try:
        DiffGFDNConfig(not_a_key=1)
    except Exception:
        pass
    else:
        raise AssertionError("unknown keys must be rejected")


def test_geq_absorption_design_matches_the_reference_cascades():
    """decay_times_to_gain_filters_geq (absorption_filters.py:108-155 / filters/geq.py) against coefficients produced
    by the reference's own design (tests/golden/geq_design.npz): same layout, cascade responses within 0.05 dB (the
    reference's section gains carry its float32 probing / LBFGS noise, see design_geq), and the target attenuation
    10^(-3 m / (fs T60(f))) is met at the band centres."""
    import os
    from diffgfdn_b200.absorption_filters import decay_times_to_gain_filters_geq
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geq_design.npz"))
    ref = g["coeffs"]
    fs = float(g["fs"])
    got = decay_times_to_gain_filters_geq(list(g["bands"]), g["t60"], list(g["delays"]), fs).numpy()
    assert got.shape == ref.shape == (len(g["bands"]) + 3, len(g["delays"]), 3, 2)

    def cascade_db(c, w):
        z = np.exp(-1j * w)
        num = c[..., 0, 0, None] + c[..., 1, 0, None] * z + c[..., 2, 0, None] * z * z
        den = c[..., 0, 1, None] + c[..., 1, 1, None] * z + c[..., 2, 1, None] * z * z
        return 20 * np.log10(np.abs(np.prod(num / den, axis=0)))

    w = np.linspace(0.001, np.pi, 4000)
    assert np.abs(cascade_db(got, w) - cascade_db(ref, w)).max() < 0.05
    wb = 2 * np.pi * g["bands"] / fs
    want = 20 * np.log10((10.0**(-3.0 / fs / g["t60"]))[None, :]**g["delays"][:, None])  # (delays, bands)
    assert np.abs(cascade_db(got, wb) - want).max() < 0.5  # a graphic equaliser meets its band targets to a fraction of a dB


def test_bin_slices_of_the_sharded_step_cover_every_bin_once():
    """shard_bins: rank r solves bins [r * per, min(k, (r + 1) * per)), per = ceil(k / world) rounded up to an even count
    (16-byte units of the complex64 x G rows moved by the peer-memory gather)."""
    from diffgfdn_b200.fused import ShardedEDCStep
    for k, world in [(65537, 8), (65537, 2), (131073, 4), (2049, 2), (7, 8), (16, 3), (1, 2)]:
        seen = []
        for rank in range(world):
            step = ShardedEDCStep.__new__(ShardedEDCStep)
            step.world_size, step.rank = world, rank
            per = step._bins_per_rank(k)
            lo, hi = step._bin_slice(k)
            assert per % 2 == 0 and per * world >= k and 0 <= lo <= hi <= k and hi - lo <= per
            assert lo == min(k, rank * per)
            seen.extend(range(lo, hi))
        assert seen == list(range(k))


def test_srir_to_brir_matches_the_reference_loops():
    """SH -> binaural step of the auralisation chain (reference sofa_parser.py:452-505) against its restatement with the
    reference's loops and einsum (oracle/auralisation_oracle.py); device-agnostic torch code, float64 here."""
    from diffgfdn_b200.inference import srir_to_brir
    from oracle import auralisation_oracle as A
    rng = np.random.default_rng(5)
    r, order, t, th, o = 3, 2, 300, 64, 4
    c = (order + 1)**2
    srirs = rng.standard_normal((r, c, t)) * np.exp(-np.arange(t) / 60.0)
    hrir_sh = rng.standard_normal((c, 2, th)) * np.exp(-np.arange(th) / 10.0)
    rot = np.stack([np.linalg.qr(rng.standard_normal((c, c)))[0] for _ in range(o)])
    want = A.convert_srir_to_brir(srirs, hrir_sh, rot)
    got = srir_to_brir(torch.tensor(srirs), torch.tensor(hrir_sh), torch.tensor(rot)).numpy()
    assert got.shape == want.shape == (r, o, 512, 2)
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
