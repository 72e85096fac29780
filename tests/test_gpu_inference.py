"""Inference / auralisation chain (diffgfdn_b200/inference.py) against the CPU oracle restatement of the reference
(oracle/auralisation_oracle.py, pinned to the reference's own `filter_overlap_add` by tests/test_oracle_golden.py):
sub-band FIR + band sum of run_subband_training_treble.py:316-358 and the moving-listener overlap-add with linear
cross-fade of sound_examples.py:163-226. Tolerance: 1e-5 of peak (BASELINE.json)."""
import numpy as np
import pytest
import torch

from oracle import auralisation_oracle as A
from oracle import gfdn_oracle as O

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _bands(bands, delays, g, gen, fs, t60):
    n = len(delays)
    a, gam, b, c = [], [], [], []
    for bd in range(bands):
        m_raw = (2 * torch.rand(g, n // g, n // g, dtype=F64, generator=gen) - 1) / np.sqrt(n // g)
        a.append(O.coupled_feedback_matrix(m_raw, np.pi / 4 * torch.rand(g * (g - 1) // 2, dtype=F64, generator=gen)).float())
        gam.append(O.decay_times_to_gain_per_sample([t * (1 - 0.1 * bd) for t in t60], delays, fs, g).float())
        b.append((torch.randn(n, dtype=F64, generator=gen) / n).float())
        c.append((torch.randn(n, dtype=F64, generator=gen) / n).float())
    return a, gam, b, c


def _oracle_group_irs(delays, a, gam, b, c, g, t):
    n = len(delays)
    return np.stack([O.fdn_time_domain(delays, gam[bd].double(), a[bd].double(), b[bd].double(), c[bd].double(), t)
                     .reshape(t, g, n // g).sum(-1).numpy() for bd in range(len(a))])  # (bands, T, G)


@pytest.mark.parametrize("hop,fade,alpha,rir_len,npos", [(80, 40, 0.5, 300, 12), (30, 40, 1.0, 200, 9), (64, 16, 0.25, 40, 6)])
def test_moving_listeners_match_filter_overlap_add(hop, fade, alpha, rir_len, npos):
    from diffgfdn_b200.inference import GFDNAuraliser
    fs, g, bands = 8000.0, 3, 2
    delays = [17, 19, 23, 29, 31, 37]
    gen = torch.Generator().manual_seed(7)
    a, gam, b, c = _bands(bands, delays, g, gen, fs, [0.02, 0.03, 0.04])
    firs = torch.randn(bands, 21, generator=gen) * torch.hann_window(21, periodic=False)
    positions, listeners = 5, 4
    s = 2 * torch.rand(bands, positions, g, generator=gen) - 1
    traj = torch.randint(0, positions, (listeners, npos), generator=gen)
    stim = torch.randn(333, generator=gen)
    aur = GFDNAuraliser(torch.tensor([delays] * bands), torch.stack(a), torch.stack(gam), torch.stack(b), torch.stack(c), g, firs)
    out = aur.moving_listeners(stim, s, traj, hop, rir_len, fade, alpha).cpu().double().numpy()
    # oracle: static RIR of every position (recursion -> receiver mix per band -> FIR -> band sum), then the OLA per listener
    q = _oracle_group_irs(delays, a, gam, b, c, g, rir_len)
    band_rirs = np.einsum('bpg,btg->bpt', s.double().numpy(), q)  # (bands, P, T)
    rirs = A.subband_sum_rir(band_rirs, firs.double().numpy())  # (P, T + L - 1)
    ext = A.extend_stimulus(stim.numpy(), npos * hop)
    for r in range(listeners):
        ref = A.filter_overlap_add(ext, rirs[traj[r].numpy()], hop, fade, alpha)
        err = np.abs(out[r] - ref).max() / np.abs(ref).max()
        assert err < 1e-5, (r, err)
    # the static path: sub-band synthesis of every position
    h = aur.static_rirs(s, rir_len).cpu().double().numpy()
    assert np.abs(h - rirs).max() / np.abs(rirs).max() < 1e-5


def test_auraliser_from_trained_models_matches_irfft_of_the_model_response():
    """`GFDNAuraliser.from_models` on a DiffGFDNVarReceiverPos: the rendered static RIR equals irfft(H) of the model's
    own frequency-sampled response (get_response, utils.py:149-179) once the tail has decayed inside nfft."""
    import numpy as np

    from diffgfdn_b200.config import DiffGFDNConfig, FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.inference import GFDNAuraliser
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    from diffgfdn_b200.utils import unit_circle_grid
    torch.manual_seed(0)
    fs, nfft = 32000.0, 2**15
    delays = DiffGFDNConfig(seed=235265, num_delay_lines=12).delay_length_samps
    net = DiffGFDNVarReceiverPos(fs, 3, delays, 'cuda', FeedbackLoopConfig(use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16,
                                                    num_fourier_features=4), use_absorption_filters=False,
                                 common_decay_times=np.array([[0.05, 0.08, 0.1]]), use_colorless_loss=False)
    pos = torch.rand(3, 3)
    data = dict(z_values=unit_circle_grid(nfft), listener_position=pos, norm_listener_position=pos)
    with torch.no_grad():
        H = net(data)
        s = net.output_scalars.gains({'norm_listener_position': pos.cuda()})
    h_ref = torch.fft.irfft(H.to(torch.complex128), n=nfft)
    aur = GFDNAuraliser.from_models([net])
    h = aur.static_rirs(s.unsqueeze(0), nfft).double()
    assert float((h - h_ref).abs().max() / h_ref.abs().max()) < 1e-4


def test_srir_to_brir_on_the_device_matches_the_reference_loops():
    """SH -> binaural step (reference sofa_parser.py:452-505) in float32 on the GPU against the restatement with the
    reference's own loops; 1e-5 of peak."""
    from diffgfdn_b200.inference import srir_to_brir
    rng = np.random.default_rng(9)
    r, order, t, th, o = 5, 2, 1000, 128, 6
    c = (order + 1)**2
    srirs = rng.standard_normal((r, c, t)) * np.exp(-np.arange(t) / 150.0)
    hrir_sh = rng.standard_normal((c, 2, th)) * np.exp(-np.arange(th) / 20.0)
    rot = np.stack([np.linalg.qr(rng.standard_normal((c, c)))[0] for _ in range(o)])
    want = A.convert_srir_to_brir(srirs, hrir_sh, rot)
    got = srir_to_brir(torch.tensor(srirs, dtype=torch.float32).cuda(), torch.tensor(hrir_sh, dtype=torch.float32).cuda(),
                       torch.tensor(rot, dtype=torch.float32).cuda())
    assert got.is_cuda and tuple(got.shape) == want.shape
    assert np.abs(got.cpu().numpy() - want).max() < 1e-5 * np.abs(want).max()
