"""Parity of the TIMED paths against the CPU oracle at the BASELINE shapes.

(a) `ShardedEDCStep` (what bench.py times) at configs[3]'s shape -- N = 24 delay lines, nfft = 2^18 (K = 131 073
    bins), T60 = (0.3, 0.8, 1.5) s, tn = 47 360 samples per receiver, the K3d tile that the bench line uses -- against
    `oracle/gfdn_oracle.py` (float64, the reference's own formulation: dense inverse per bin, projection of every
    receiver over every bin, one irfft per receiver) on the SAME parameters: EDC within 0.01 dB, every parameter
    gradient within 1e-3 relative (BASELINE.json tolerances; reference trainer.py:259-315, 452-477, losses.py:201-238).
(b) The renderer at configs[4]'s shape -- N = 12 with the real prime delays 641..1601, 8 bands, 320 000 samples --
    against the oracle's time-domain recursion and against irfft(H) of the frequency-sampled model, 1e-5 of peak.
"""
import os

import numpy as np
import pytest
import torch

from oracle import gfdn_oracle as O

pytestmark = pytest.mark.gpu
F64 = torch.float64
FS = 32000.0
T60 = (0.3, 0.8, 1.5)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _oracle_step(net, z, pos, early, target, edc_w, feats, subband=None, edr_w=0.0):
    """Loss terms and parameter gradients of the oracle on net's parameters (float64 copies)."""
    p = {k: v.detach().cpu().to(F64).requires_grad_(v.dtype.is_floating_point and k in dict(net.named_parameters()))
         for k, v in net.state_dict().items()}
    g = net.num_groups
    delays = net.delays.cpu().to(F64)
    gamma = O.decay_times_to_gain_per_sample(T60, delays.tolist(), FS, g)
    a = O.coupled_feedback_matrix(p["feedback_loop.M"], p["feedback_loop.alpha"])
    b, c = p["input_gains"].reshape(-1), p["output_gains"].reshape(-1)
    s = O.gains_from_mlp(pos.cpu().to(F64), p, feats, g)
    zc = z.cpu()
    H = O.omni_response(zc, delays, gamma, a, b, c, s, early.cpu().to(torch.complex128))
    if subband is not None:
        H = H * subband.cpu().to(torch.complex128)
    tgt = target.cpu().to(torch.complex128)
    edc = O.edc_loss(tgt, H, max(T60) * 1e3, FS)
    edr = O.edr_loss(tgt, H) if edr_w else torch.zeros((), dtype=F64)
    hs, _ = O.sub_fdn_output(zc, delays, p["feedback_loop.M"], b, c)
    spec, spars = O.colorless_losses(hs, p["feedback_loop.M"], 1.0, 1.0, asym=True)
    (edc_w * edc + edr_w * edr + spec + spars).backward()
    grads = {k: p[k].grad for k, _ in net.named_parameters()}
    return dict(edc=float(edc), edr=float(edr), spec=float(spec), spars=float(spars)), grads


@pytest.mark.parametrize("replay", ["1", "0"])
def test_fused_step_matches_oracle_at_baseline_shape(replay, monkeypatch):
    """replay=1: K1 adjoint by replaying the saved float32 elimination (the timed path); 0: fresh elimination of M^H."""
    import bench
    from diffgfdn_b200 import ops
    from diffgfdn_b200.fused import ShardedEDCStep
    from diffgfdn_b200.utils import unit_circle_grid
    monkeypatch.setenv("DGFDN_SOLVE_REPLAY", replay)
    dev = torch.device("cuda")
    rows, nfft = 6, 2**18
    net = bench.build_net(dev, seed=1234)
    z = unit_circle_grid(nfft, device=dev)
    gen = torch.Generator(device=dev).manual_seed(5)
    pos = torch.rand(rows, 3, device=dev, generator=gen)
    early, target = bench.synth_responses(rows, nfft, dev, 17)
    step = ShardedEDCStep(net, max(T60) * 1e3, edc_weight=10.0)
    step.attach(z, pos, None, None)
    assert step.tn == 47360 and step.k == 131073
    step.attach(z, pos, step.precompute_early_window(early), step.precompute_target_db(target))
    assert step.use_fused, "the bench line's receiver kernel must be the one under test"
    out = step.step()
    torch.cuda.synchronize()
    feats = net.output_scalars.encoder.num_fourier_features
    ref, g_ref = _oracle_step(net, z, pos, early, target, 10.0, feats)
    assert abs(float(out["edc_loss"]) / 10.0 - ref["edc"]) < 0.01, (float(out["edc_loss"]) / 10.0, ref["edc"])
    assert abs(float(out["spectral_loss"]) - ref["spec"]) < 1e-4 * abs(ref["spec"])
    assert abs(float(out["sparsity_loss"]) - ref["spars"]) < 1e-4 * max(1.0, abs(ref["spars"]))
    worst = {}
    for k, p in net.named_parameters():
        worst[k] = _rel(p.grad.detach().cpu().to(F64), g_ref[k])
    bad = {k: v for k, v in worst.items() if not v < 1e-3}
    assert not bad, f"gradient mismatch vs oracle: {bad}"
    # the receiver kernel variant of the bench line
    info = ops.td_fused_info(3, step.tn)
    assert info["variant"] >= 0


def test_renderer_matches_oracle_at_configs4_shape():
    """BASELINE configs[4]: N = 12 (prime delays 641..1601 at 32 kHz), G = 3, 8 bands, 10 s = 320 000 samples: the
    block-recursive renderer (500 dependent blocks, float32 state) against the oracle's float64 recursion on every
    band, the moving-listener mix against its restatement, and a static listener against irfft(H) (utils.py:169)."""
    from diffgfdn_b200 import ops
    from diffgfdn_b200.config import DiffGFDNConfig
    delays = DiffGFDNConfig(seed=235265, num_delay_lines=12).delay_length_samps
    assert min(delays) >= 641 and max(delays) <= 1601
    g, bands, t = 3, 8, 320000
    gen = torch.Generator().manual_seed(9)
    a, gam, b, c = [], [], [], []
    for bd in range(bands):
        m_raw = (2 * torch.rand(g, 4, 4, dtype=F64, generator=gen) - 1) / 2.0
        a.append(O.coupled_feedback_matrix(m_raw, np.pi / 4 * torch.rand(3, dtype=F64, generator=gen)))
        t60 = [T60[i] * (1.0 - 0.05 * bd) for i in range(g)]
        gam.append(O.decay_times_to_gain_per_sample(t60, delays, FS, g))
        b.append((2 * torch.randn(12, dtype=F64, generator=gen) - 1) / 12)
        c.append((2 * torch.randn(12, dtype=F64, generator=gen) - 1) / 12)
    dl = torch.tensor([delays] * bands, dtype=torch.int32)
    cu = lambda v: torch.stack(v).float().cuda()  # noqa: E731
    q = ops.render_groups(dl.cuda(), cu(a), cu(gam), cu(b), cu(c), g, t)
    qc = q.cpu().to(F64)
    for bd in range(bands):
        qo = O.fdn_time_domain(delays, gam[bd].float().to(F64), a[bd].float().to(F64), b[bd].float().to(F64),
                               c[bd].float().to(F64), t).reshape(t, g, 4).sum(-1)
        err = float((qc[bd] - qo).abs().max() / qo.abs().max())
        assert err < 1e-5, (bd, err)
    # moving listeners, 100 ms hops (sound_examples.py:87)
    hop, listeners, positions = 3200, 6, 11
    s = 2 * torch.rand(bands, positions, g, generator=gen) - 1
    traj = torch.randint(0, positions, (listeners, (t + hop - 1) // hop), generator=gen, dtype=torch.int32)
    out = ops.render_mix(s.cuda(), traj.cuda(), q, hop).cpu().to(F64)
    ref = torch.zeros(listeners, t, dtype=F64)
    for r in range(listeners):
        gains = s[:, traj[r].long()].to(F64)  # (bands, hops, g)
        per_sample = gains.repeat_interleave(hop, dim=1)[:, :t]  # (bands, t, g)
        ref[r] = (per_sample * qc).sum(dim=(0, 2))
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-5
    # static listener == irfft(H) of the frequency-sampled model, nfft = 2^19 >= the rendered length
    nfft = 2**19
    H = O.omni_response(O.z_grid(nfft), torch.tensor(delays, dtype=F64), gam[0].float().to(F64), a[0].float().to(F64),
                        b[0].float().to(F64), c[0].float().to(F64), s[0, :1].to(F64))
    h = O.impulse_response(H)[0, :t]
    hr = torch.einsum('g,tg->t', s[0, 0].to(F64), qc[0])
    assert float((hr - h).abs().max() / h.abs().max()) < 1e-5
