"""Generate golden fixtures by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference has no golden vectors of its own for this path, so these
fixtures -- outputs of the reference's own modules on seeded synthetic inputs -- are what pins
the oracle (`oracle/gfdn_oracle.py`) and, through it, the CUDA kernels.
Inputs follow SURVEY.md section 8(d) at sizes that keep each fixture well under 1 MB.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from diff_gfdn.colorless_fdn.losses import amse_loss, mse_loss, sparsity_loss  # noqa: E402
from diff_gfdn.config.config import (DiffGFDNConfig, FeedbackLoopConfig, OutputFilterConfig,  # noqa: E402
                                     TrainerConfig)
from diff_gfdn.losses import directional_edc_loss, edc_loss, edr_loss  # noqa: E402
from diff_gfdn.model import (DiffDirectionalFDNVarReceiverPos, DiffGFDNSinglePos, DiffGFDNVarReceiverPos,  # noqa: E402
                             DiffGFDNVarSourceReceiverPos)
from diff_gfdn.trainer import (DirectionalFDNVarReceiverPosTrainer, SinglePosTrainer,  # noqa: E402
                               VarReceiverPosTrainer)
from diff_gfdn.utils import get_response  # noqa: E402
import spaudiopy  # noqa: E402  (stub)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FS = 32000.0


def synth_batch(nfft, bsz, seed, early=True, radius=1.0, early_scale=1.0):
    """Synthetic receiver batch: decaying-noise target RIRs, 20 ms early part."""
    rng = np.random.default_rng(seed)
    k = nfft // 2 + 1
    w = np.fft.rfftfreq(nfft) * 2 * np.pi
    z = torch.polar(torch.full((k, ), radius, dtype=torch.float64), torch.tensor(w))
    pos = torch.tensor(rng.uniform(0, 1, (bsz, 3)))
    t = np.arange(nfft // 2)
    tau = rng.uniform(0.02, 0.05, (bsz, 1)) * FS
    rir = rng.standard_normal((bsz, nfft // 2)) * np.exp(-t[None, :] / tau)
    target = torch.tensor(np.fft.rfft(rir, n=nfft, axis=-1))
    e = rir.copy() * early_scale
    e[:, 640:] = 0.0
    e[:, 560:640] *= np.hanning(160)[80:][None, :]
    d = torch.tensor(np.fft.rfft(e, n=nfft, axis=-1)) if early else torch.zeros(bsz, k, dtype=torch.complex128)
    return dict(z_values=z, listener_position=pos, norm_listener_position=pos.clone(),
                source_position=torch.zeros(bsz, 3, dtype=torch.float64), target_early_response=d,
                target_rir_response=target)


def np_state(net):
    return {f"param/{k}": v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}


def grads_of(net):
    return {f"grad/{k}": p.grad.detach().cpu().numpy().copy() for k, p in net.named_parameters() if p.grad is not None}


def make_trainer(cls, net, **kw):
    tmp = tempfile.mkdtemp(prefix="dgfdn_golden_")
    cfg = TrainerConfig(train_dir=os.path.join(tmp, "out"), ir_dir=os.path.join(tmp, "ir"), **kw)
    return cls(net, cfg)


def case_omni(name, n_lines, nfft, bsz, t60, seed, hidden, neurons, feats, radius=1.0, subband=False, steps=2,
              early_scale=1.0, svf=False, pole_factor=1.0, geq_bands=None):
    cfg = DiffGFDNConfig(seed=235265, num_delay_lines=n_lines)
    delays = cfg.delay_length_samps
    torch.manual_seed(seed)
    np.random.seed(seed)
    net = DiffGFDNVarReceiverPos(FS, 3, delays, 'cpu', FeedbackLoopConfig(use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs=svf, num_hidden_layers=hidden,
                                                    num_neurons_per_layer=neurons, num_fourier_features=feats,
                                                    compress_pole_factor=pole_factor),
                                 use_absorption_filters=geq_bands is not None,
                                 common_decay_times=np.array([t60]) if geq_bands is None else np.asarray(t60),
                                 band_centre_hz=geq_bands, use_colorless_loss=True)
    # early_scale < 1 brings the direct path d down towards the level of the late (GFDN) part: with the short T60s
    # that fit these small fixtures a random initialisation leaves the late part ~1e4 below |d|, which no float32
    # FFT can resolve to 1e-3 (the reference runs this path in float64)
    data = synth_batch(nfft, bsz, seed + 1, radius=radius, early_scale=early_scale)
    trainer = make_trainer(VarReceiverPosTrainer, net, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=nfft, io_lr=0.01, lr=0.01)
    out = {"meta/delays": np.array(delays), "meta/nfft": nfft, "meta/fs": FS, "meta/t60": np.array(t60),
           "meta/radius": radius, "meta/feats": feats, "meta/edc_w": 10.0, "meta/edr_w": 1.0,
           "meta/max_ir_len_ms": float(np.max(t60) * 1e3)}
    for key in ("listener_position", "norm_listener_position", "target_early_response", "target_rir_response"):
        out[f"data/{key}"] = data[key].numpy()
    out.update(np_state(net))

    if subband:
        k = nfft // 2 + 1
        rng = np.random.default_rng(seed + 7)
        fir = np.hanning(257) * np.sinc(np.linspace(-8, 8, 257)) * np.cos(np.linspace(-8, 8, 257) * 9.0)
        filt = torch.fft.rfft(torch.tensor(fir + 0.01 * rng.standard_normal(257)), n=nfft)
        assert filt.shape[0] == k
        trainer.subband_process_config = types.SimpleNamespace()  # any non-None value enables trainer.py:457-461
        trainer.subband_filter_freq_resp = filt
        out["data/subband_filter"] = filt.numpy()

    # forward + individual losses + backward, straight through the reference trainer code path
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    Huse = H * trainer.subband_filter_freq_resp if subband else H
    losses = trainer.calculate_losses(data, Huse, (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    out["out/H"] = H.detach().numpy()
    out["out/H_sub"] = Hs.detach().numpy()
    out["out/H_sub_per_del_s16"] = Hsd.detach().numpy()[:, ::16, :]
    out["out/A"] = net.feedback_loop.get_coupled_feedback_matrix().detach().numpy()
    out["out/phi"] = net.feedback_loop.phi.detach().numpy()
    if svf:  # SVF output filters (gain_filters.py:334-402): constrained (resonance, gain dB) and the biquads
        out["out/svf_params"] = net.output_filters.svf_params.detach().numpy()
        out["out/biquads"] = np.stack([np.stack([np.concatenate([c.num_coeffs.detach().numpy(),
                                                                 c.den_coeffs.detach().numpy()], axis=-1)
                                                 for c in row]) for row in net.output_filters.biquad_cascade])
        out["meta/pole_factor"] = pole_factor
    else:
        out["out/s"] = net.output_scalars.gains.detach().numpy()
    if geq_bands is None:
        out["out/gamma"] = net.feedback_loop.delay_line_gains.detach().numpy()
    else:  # GEQ absorption filters (absorption_filters.py:108-155): responses on a decimated grid
        out["meta/band_centre_hz"] = np.asarray(geq_bands, dtype=np.float64)
        out["out/gamma_z_s8"] = np.stack([f(data['z_values'][::8]).detach().numpy()
                                          for f in net.feedback_loop.delay_line_gains])
    for kk, v in losses.items():
        out[f"loss/{kk}"] = float(v.detach())
    out["loss/total"] = float(total.detach())
    out.update(grads_of(net))
    # stand-alone loss callables (no weights)
    mx = float(np.max(t60) * 1e3)
    out["loss/edc_raw"] = float(edc_loss(mx, FS)(data['target_rir_response'], Huse.detach()))
    out["loss/edr_raw"] = float(edr_loss(FS)(data['target_rir_response'], Huse.detach()))
    out["loss/mse_g0"] = float(mse_loss()(Hs.detach()[..., 0], torch.ones_like(Hs.detach()[..., 0])))
    out["loss/amse_g0"] = float(amse_loss()(Hs.detach()[..., 0], torch.ones_like(Hs.detach()[..., 0])))
    # masked EDC with an explicit index (the reference draws it from the global RNG: losses.py:221-223)
    crit = edc_loss(mx, FS, use_mask=True)
    torch.manual_seed(99)
    masked = float(crit(data['target_rir_response'], Huse.detach()))
    torch.manual_seed(99)
    tlen = min(int(mx * 1e-3 * FS), nfft // 2 + 1) - 640
    probs = torch.empty(tlen).uniform_(0, 1)
    idx = torch.argwhere(torch.bernoulli(probs)).squeeze(-1)
    out["data/edc_mask_index"] = idx.numpy()
    out["loss/edc_masked_raw"] = masked
    # time-domain response (utils.py:169)
    _, _, h = get_response(data, net)
    out["out/h_s8"] = h.numpy()[:, ::8]

    # a few optimisation steps exactly as trainer.py:373-379 (normalize + train_step)
    step_losses = []
    for _ in range(steps):
        trainer.normalize(data)
        loss, _ = trainer.train_step(data)
        step_losses.append(loss)
    out["steps/loss"] = np.array(step_losses)
    for kname, v in net.state_dict().items():
        out[f"steps/param/{kname}"] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "total", out["loss/total"], {k: v for k, v in out.items() if k.startswith("loss/")}, step_losses)


def _finish_variant(name, net, trainer, data, out, t60, nfft):
    """forward + trainer losses + backward of a model variant; common tail of the a-8c cases"""
    out.update(np_state(net))
    net.zero_grad()
    if getattr(net, "use_svf_in_output", False) and isinstance(net, DiffGFDNSinglePos):
        # the reference deep-copies its non-leaf biquad tensors (model.py:903-906), which raises under autograd with
        # this torch: the SVF branch of DiffGFDNSinglePos only runs without a graph -- forward values only
        with torch.no_grad():
            H, (Hs, Hsd) = net(data)
            losses = trainer.calculate_losses(data, H, (Hs, Hsd))
            total = sum(losses.values())
        for tag, casc in (("in", net.input_biquad_cascade), ("out", net.output_biquad_cascade)):
            out[f"out/biquads_{tag}"] = np.stack([np.concatenate([c.num_coeffs.numpy(), c.den_coeffs.numpy()], axis=-1)
                                                  for c in casc])
    else:
        H, (Hs, Hsd) = net(data)
        losses = trainer.calculate_losses(data, H, (Hs, Hsd))
        total = sum(losses.values())
        total.backward()
    out["out/H"] = H.detach().numpy()
    out["out/H_sub"] = Hs.detach().numpy()
    out["out/A"] = net.feedback_loop.get_coupled_feedback_matrix().detach().numpy()
    for kk, v in losses.items():
        out[f"loss/{kk}"] = float(v.detach())
    out["loss/total"] = float(total.detach())
    out.update(grads_of(net))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: v for k, v in out.items() if k.startswith("loss/")})


def case_src_rx(name, n_lines, nfft, bsz, t60, seed, hidden, neurons, feats, svf_in=False, svf_out=False,
                pole_factor=0.998):
    """DiffGFDNVarSourceReceiverPos (model.py:303-452): MLP-driven gains -- or SVF cascades (svf_in / svf_out) -- on the
    source AND the receiver side."""
    cfg = DiffGFDNConfig(seed=235265, num_delay_lines=n_lines)
    delays = cfg.delay_length_samps
    torch.manual_seed(seed)
    np.random.seed(seed)
    ofc_out = OutputFilterConfig(use_svfs=svf_out, num_hidden_layers=hidden, num_neurons_per_layer=neurons,
                                 num_fourier_features=feats, compress_pole_factor=pole_factor)
    ofc_in = OutputFilterConfig(use_svfs=svf_in, num_hidden_layers=hidden, num_neurons_per_layer=neurons,
                                num_fourier_features=feats, compress_pole_factor=pole_factor)
    net = DiffGFDNVarSourceReceiverPos(FS, 3, delays, 'cpu', FeedbackLoopConfig(use_zero_coupling=False), ofc_out, ofc_in,
                                       use_absorption_filters=False, learn_common_decay_times=False,
                                       common_decay_times=np.array([t60]), use_colorless_loss=True)
    data = synth_batch(nfft, bsz, seed + 1, early_scale=1e-3)
    rng = np.random.default_rng(seed + 5)
    data['source_position'] = torch.tensor(rng.uniform(0, 1, (bsz, 3)))
    trainer = make_trainer(VarReceiverPosTrainer, net, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=nfft)
    out = {"meta/delays": np.array(delays), "meta/nfft": nfft, "meta/fs": FS, "meta/t60": np.array(t60),
           "meta/radius": 1.0, "meta/feats": feats, "meta/edc_w": 10.0, "meta/edr_w": 1.0,
           "meta/max_ir_len_ms": float(np.max(t60) * 1e3), "meta/svf_in": svf_in, "meta/svf_out": svf_out,
           "meta/pole_factor": pole_factor}
    if svf_in or svf_out:  # the float32 cascades the reference builds for this batch, (B, G, S, 6) each
        with torch.no_grad():
            for tag, on in (("in", svf_in), ("out", svf_out)):
                if on:
                    filt = net.input_filters if tag == "in" else net.output_filters
                    filt(data)
                    out[f"out/biquads_{tag}"] = np.stack([np.stack([np.concatenate(
                        [c.num_coeffs.detach().numpy(), c.den_coeffs.detach().numpy()], axis=-1) for c in row])
                        for row in filt.biquad_cascade])
    for key in ("listener_position", "norm_listener_position", "source_position", "target_early_response",
                "target_rir_response"):
        out[f"data/{key}"] = data[key].numpy()
    _finish_variant(name, net, trainer, data, out, t60, nfft)


def case_single(name, n_lines, nfft, t60, seed, svf_in, svf_out, pole_factor=0.998):
    """DiffGFDNSinglePos (model.py:667-960): one source-receiver pair, learnable per-group scalars or SVF cascades."""
    cfg = DiffGFDNConfig(seed=235265, num_delay_lines=n_lines)
    delays = cfg.delay_length_samps
    torch.manual_seed(seed)
    np.random.seed(seed)
    net = DiffGFDNSinglePos(FS, 3, delays, 'cpu', FeedbackLoopConfig(use_zero_coupling=False),
                            OutputFilterConfig(use_svfs=svf_out, compress_pole_factor=pole_factor),
                            use_absorption_filters=False, common_decay_times=np.array([t60]), use_colorless_loss=True,
                            input_filter_config=OutputFilterConfig(use_svfs=svf_in, compress_pole_factor=pole_factor))
    with torch.no_grad():  # move the SVF gains off their 0 dB initialisation so every coefficient path is exercised
        for nm, prm in net.named_parameters():
            if 'svf_params' in nm:
                prm[..., 1] = torch.randn_like(prm[..., 1])
            if nm in ('input_scalars', 'output_scalars'):
                prm.mul_(1.0 + 0.3 * torch.randn_like(prm))
    batch = synth_batch(nfft, 1, seed + 1, early_scale=1e-3)
    data = {k: (v[0] if k in ("target_early_response", "target_rir_response") else v) for k, v in batch.items()}
    trainer = make_trainer(lambda n, c: SinglePosTrainer(n, c, "golden"), net, use_colorless_loss=True,
                           use_asym_spectral_loss=True, edc_loss_weight=10.0, num_freq_bins=nfft)
    out = {"meta/delays": np.array(delays), "meta/nfft": nfft, "meta/fs": FS, "meta/t60": np.array(t60),
           "meta/radius": 1.0, "meta/edc_w": 10.0, "meta/edr_w": 1.0, "meta/max_ir_len_ms": float(np.max(t60) * 1e3),
           "meta/svf_in": svf_in, "meta/svf_out": svf_out, "meta/pole_factor": pole_factor}
    for key in ("target_early_response", "target_rir_response"):
        out[f"data/{key}"] = data[key].numpy()
    _finish_variant(name, net, trainer, data, out, t60, nfft)


def case_directional(name, nfft, bsz, t60, seed, hidden, neurons, feats, skip):
    ambi = 2
    lsh = (ambi + 1)**2
    cfg = DiffGFDNConfig(seed=235265, ambi_order=ambi)
    delays = cfg.delay_length_samps
    assert len(delays) == 3 * lsh
    rng = np.random.default_rng(seed + 3)
    nj = 12
    ymat = (rng.standard_normal((nj, lsh)) / 3.0).astype(np.float32)
    spaudiopy.sph.design_sph_filterbank = lambda *a, **k: (ymat, ymat.T)
    torch.manual_seed(seed)
    np.random.seed(seed)
    dirs = rng.uniform(0, np.pi, (2, nj))
    net = DiffDirectionalFDNVarReceiverPos(FS, 3, delays, 'cpu', FeedbackLoopConfig(use_zero_coupling=False),
                                           OutputFilterConfig(use_svfs=False, num_hidden_layers=hidden,
                                                              num_neurons_per_layer=neurons,
                                                              num_fourier_features=feats,
                                                              use_skip_connections=skip),
                                           ambi, dirs, common_decay_times=np.array([t60]), use_colorless_loss=True)
    with torch.no_grad():  # lift |H| so that the predicted EDC sits above the eps floor of utils.db
        net.input_gains.mul_(6.0)
        net.output_gains.mul_(6.0)
    data = synth_batch(nfft, bsz, seed + 1, early=False)
    amps = torch.tensor(rng.uniform(1e-4, 1.0, (bsz, nj, 3)))
    data['target_common_slope_amps'] = amps
    trainer = make_trainer(DirectionalFDNVarReceiverPosTrainer, net, use_colorless_loss=True,
                           use_asym_spectral_loss=False, edc_loss_weight=2.0, num_freq_bins=nfft)
    out = {"meta/delays": np.array(delays), "meta/nfft": nfft, "meta/fs": FS, "meta/t60": np.array(t60),
           "meta/feats": feats, "meta/skip": skip, "meta/edc_w": 2.0, "meta/edc_len_ms": float(np.max(t60) * 1e3),
           "data/Y": ymat, "data/amps": amps.numpy(), "data/norm_listener_position": data['norm_listener_position'].numpy()}
    out.update(np_state(net))
    net.zero_grad()
    H_sh, (Hs, Hsd) = net(data)
    H_dir = trainer.convert_ambi_rir_to_directional_rir(H_sh)
    losses = trainer.calculate_losses(data, H_dir, (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    out["out/H_sh"] = H_sh.detach().numpy()
    out["out/H_dir_s4"] = H_dir.detach().numpy()[..., ::4]
    out["out/H_sub"] = Hs.detach().numpy()
    out["out/w"] = net.sh_output_scalars.weights.detach().numpy()
    out["out/A"] = net.feedback_loop.get_coupled_feedback_matrix().detach().numpy()
    out["out/envelopes_s16"] = trainer.criterion[0].envelopes.numpy()[:, ::16]
    for kk, v in losses.items():
        out[f"loss/{kk}"] = float(v.detach())
    out["loss/total"] = float(total.detach())
    out.update(grads_of(net))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: v for k, v in out.items() if k.startswith("loss/")})


def case_random_coupling(name, n_lines, nfft, bsz, t60, seed):
    """coupling_matrix_type: random_matrix (feedback_loop.py:272-277, 350-352): A = expm(skew(R)) with an unstructured
    learnable R (N x N) -- the single-room sub-band YAML. No colorless loss (there are no group matrices)."""
    from diff_gfdn.config.config import CouplingMatrixType
    cfg = DiffGFDNConfig(seed=235265, num_delay_lines=n_lines)
    delays = cfg.delay_length_samps
    torch.manual_seed(seed)
    np.random.seed(seed)
    g = len(t60)
    net = DiffGFDNVarReceiverPos(FS, g, delays, 'cpu', FeedbackLoopConfig(coupling_matrix_type=CouplingMatrixType.RANDOM),
                                 OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16,
                                                    num_fourier_features=4),
                                 use_absorption_filters=False, common_decay_times=np.array([t60]), use_colorless_loss=False)
    data = synth_batch(nfft, bsz, seed + 1, early_scale=1e-3)
    trainer = make_trainer(VarReceiverPosTrainer, net, use_colorless_loss=False, edc_loss_weight=10.0, num_freq_bins=nfft)
    out = {"meta/delays": np.array(delays), "meta/nfft": nfft, "meta/fs": FS, "meta/t60": np.array(t60),
           "meta/feats": 4, "meta/edc_w": 10.0, "meta/max_ir_len_ms": float(np.max(t60) * 1e3)}
    for key in ("listener_position", "norm_listener_position", "target_early_response", "target_rir_response"):
        out[f"data/{key}"] = data[key].numpy()
    out.update(np_state(net))
    net.zero_grad()
    H = net(data)
    losses = trainer.calculate_losses(data, H)
    total = sum(losses.values())
    total.backward()
    out["out/H"] = H.detach().numpy()
    out["out/A"] = net.feedback_loop.coupled_feedback_matrix.detach().numpy()
    for kk, v in losses.items():
        out[f"loss/{kk}"] = float(v.detach())
    out["loss/total"] = float(total.detach())
    out.update(grads_of(net))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: v for k, v in out.items() if k.startswith("loss/")})


def case_filter_coupling(name, n_lines, nfft, bsz, t60, seed, order):
    """coupling_matrix_type: filter_matrix (feedback_loop.py:90-143, 311-324, 362-373, 414-421, 447-453): the coupling
    between groups is a paraunitary FIR matrix Phi(z) of `order` taps built from Householder stages, A(z) per bin."""
    from diff_gfdn.config.config import CouplingMatrixType
    cfg = DiffGFDNConfig(seed=235265, num_delay_lines=n_lines)
    delays = cfg.delay_length_samps
    torch.manual_seed(seed)
    np.random.seed(seed)
    g = len(t60)
    net = DiffGFDNVarReceiverPos(FS, g, delays, 'cpu',
                                 FeedbackLoopConfig(coupling_matrix_type=CouplingMatrixType.FILTER, pu_matrix_order=order,
                                                    use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16,
                                                    num_fourier_features=4),
                                 use_absorption_filters=False, common_decay_times=np.array([t60]), use_colorless_loss=True)
    data = synth_batch(nfft, bsz, seed + 1, early_scale=1e-3)
    trainer = make_trainer(VarReceiverPosTrainer, net, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=nfft)
    out = {"meta/delays": np.array(delays), "meta/nfft": nfft, "meta/fs": FS, "meta/t60": np.array(t60),
           "meta/feats": 4, "meta/edc_w": 10.0, "meta/edr_w": 1.0, "meta/order": order,
           "meta/max_ir_len_ms": float(np.max(t60) * 1e3)}
    for key in ("listener_position", "norm_listener_position", "target_early_response", "target_rir_response"):
        out[f"data/{key}"] = data[key].numpy()
    out.update(np_state(net))
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    losses = trainer.calculate_losses(data, H, (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    out["out/H"] = H.detach().numpy()
    out["out/H_sub"] = Hs.detach().numpy()
    out["out/A"] = net.feedback_loop.coupled_feedback_matrix.detach().numpy()  # (N, N, order) complex
    out["out/phi"] = net.feedback_loop.phi.detach().numpy()  # (G, G, order)
    for kk, v in losses.items():
        out[f"loss/{kk}"] = float(v.detach())
    out["loss/total"] = float(total.detach())
    out.update(grads_of(net))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: v for k, v in out.items() if k.startswith("loss/")})


def case_dataloader(name):
    """Data boundary (dataloader.py:185-325, 515-600, 674-745): a small two-room grid through the reference classes."""
    from diff_gfdn.dataloader import (MultiRIRDataset, RIRData, RoomDataset, SingleRIRDataset, create_fixed_test_split,
                                      custom_collate)
    rng = np.random.default_rng(77)
    fs, nrec, tlen, nfft = 8000.0, 7, 1500, 2048
    rirs = rng.standard_normal((nrec, tlen)) * np.exp(-np.arange(tlen) / 300.0)
    rec = rng.uniform(0.0, 5.0, (nrec, 3))
    src = np.array([1.0, 2.0, 1.5])
    kw = dict(num_rooms=2, sample_rate=fs, source_position=src, receiver_position=rec,
              common_decay_times=np.array([[0.2, 0.4]]), room_dims=[[3.0, 2.0, 2.5], [2.0, 2.0, 2.5]],
              room_start_coord=[[0.0, 0.0, 0.0], [3.0, 0.0, 0.0]], mixing_time_ms=20.0, nfft=nfft)
    out = {"in/rirs": rirs.copy(), "in/receiver_position": rec.copy(), "in/source_position": src.copy(),
           "meta/fs": fs, "meta/nfft": nfft, "meta/radius": 1.0005}
    room = RoomDataset(rirs=rirs.copy(), **kw)
    ds = MultiRIRDataset('cpu', room, new_sampling_radius=1.0005)
    out["out/rir_mag_response"] = np.asarray(room.rir_mag_response)
    out["out/early_rir_mag_response"] = np.asarray(room.early_rir_mag_response)
    out["out/late_rir_mag_response"] = np.asarray(room.late_rir_mag_response)
    out["out/norm_receiver_position"] = room.norm_receiver_position
    out["out/z_values"] = ds.z_values.numpy()
    test_set, rest = create_fixed_test_split(ds, test_ratio=0.3, seed=4314)
    out["out/test_indices"] = np.asarray(test_set.indices)
    out["out/rest_indices"] = np.asarray(rest.indices)
    batch = custom_collate([ds[int(i)] for i in rest.indices[:3]])
    for k, v in batch.items():
        out[f"batch/{k}"] = v.numpy()
    # multi-source layout (index pairs) and the single-RIR classes
    src2 = np.stack([src, src + 0.5])
    rirs2 = rng.standard_normal((2, nrec, tlen)) * np.exp(-np.arange(tlen) / 250.0)
    out["in/rirs2"] = rirs2.copy()
    out["in/source_position2"] = src2.copy()
    room2 = RoomDataset(rirs=rirs2.copy(), **{**kw, "source_position": src2})
    ds2 = MultiRIRDataset('cpu', room2)
    batch2 = custom_collate([ds2[i] for i in (1, 9, 13)])
    for k, v in batch2.items():
        out[f"batch2/{k}"] = v.numpy()
    rd = RIRData(np.array([[0.2, 0.4]]), None, mixing_time_ms=20.0, nfft=nfft, rir=rirs[0].copy(), sample_rate=fs)
    sd = SingleRIRDataset('cpu', rd)
    out["single/rir_mag_response"] = sd.rir_mag_response.numpy()
    out["single/early_rir_mag_response"] = sd.early_rir_mag_response.numpy()
    out["single/late_rir_mag_response"] = sd.late_rir_mag_response.numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: np.asarray(v).shape for k, v in out.items() if k.startswith(("out/", "batch"))})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    case_omni("omni_n12", 12, 8192, 4, [0.05, 0.08, 0.12], 11, 1, 32, 6)
    case_omni("omni_n12_subband_r", 12, 8192, 3, [0.04, 0.09, 0.11], 12, 2, 16, 4, radius=1.00002, subband=True,
              early_scale=1e-3)
    case_omni("omni_n24", 24, 4096, 3, [0.03, 0.05, 0.06], 13, 1, 16, 4, steps=1)
    case_omni("omni_n12_svf", 12, 8192, 3, [0.05, 0.08, 0.12], 14, 1, 32, 6, svf=True, pole_factor=0.998,
              early_scale=1e-3)
    bands = [63.0, 125.0, 250.0, 500.0, 1000.0, 2000.0, 4000.0, 8000.0]
    t60_bands = np.stack([np.linspace(0.12, 0.05, 8), np.linspace(0.09, 0.06, 8), np.linspace(0.06, 0.04, 8)], axis=1)
    case_omni("omni_n12_geq_svf", 12, 8192, 2, t60_bands, 18, 1, 16, 4, svf=True, pole_factor=0.998, early_scale=1e-3,
              geq_bands=bands, steps=1)
    case_src_rx("src_rx_n12", 12, 8192, 3, [0.05, 0.08, 0.12], 15, 1, 16, 4)
    case_src_rx("src_rx_n12_svf", 12, 4096, 2, [0.04, 0.06, 0.09], 24, 1, 16, 4, svf_in=True, svf_out=True)
    case_src_rx("src_rx_n12_svf_in", 12, 4096, 2, [0.04, 0.06, 0.09], 25, 1, 16, 4, svf_in=True, svf_out=False)
    case_single("single_n12", 12, 8192, [0.05, 0.08, 0.12], 16, False, False)
    case_single("single_n12_svf", 12, 8192, [0.05, 0.08, 0.12], 17, True, True)
    case_random_coupling("random_coupling_n8", 8, 8192, 3, [0.06, 0.11], 19)
    case_filter_coupling("filter_coupling_n12", 12, 8192, 3, [0.05, 0.08, 0.12], 23, 4)
    case_dataloader("dataloader_small")
    case_directional("directional_n27", 8192, 2, [0.05, 0.08, 0.1], 21, 1, 16, 4, skip=False)
    case_directional("directional_n27_skip", 4096, 2, [0.03, 0.04, 0.05], 22, 2, 16, 4, skip=True)
