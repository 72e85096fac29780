"""GPU parity of the drop-in nn.Module / loss / trainer surface against fixtures produced by the UNMODIFIED
reference (tests/golden/*.npz) and against the CPU oracle.

Tolerances (BASELINE.json): |H| 1e-4 relative, EDC 0.01 dB, gradients 1e-3 relative."""
import numpy as np
import pytest
import torch

from golden_util import load, oracle_directional, oracle_omni, params_of

pytestmark = pytest.mark.gpu

OMNI = {  # name: (hidden, neurons, fourier features, steps)
    "omni_n12": (1, 32, 6, 2),
    "omni_n12_subband_r": (2, 16, 4, 2),
    "omni_n24": (1, 16, 4, 1),
    "omni_n12_svf": (1, 32, 6, 2),  # SVF output filters (use_svfs: True), compress_pole_factor 0.998
    "omni_n12_geq_svf": (1, 16, 4, 1),  # + GEQ absorption filters (use_absorption_filters: True): the full-band YAML
}
DIRECTIONAL = {"directional_n27": (1, 16, 4, False), "directional_n27_skip": (2, 16, 4, True)}


def rel(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a))
    b = torch.as_tensor(np.asarray(b.detach().cpu() if torch.is_tensor(b) else b))
    return float((a - b).abs().max() / b.abs().max())


def build_omni(g, hidden, neurons, feats):
    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    geq = "meta/band_centre_hz" in g
    net = DiffGFDNVarReceiverPos(float(g["meta/fs"]), 3, [int(v) for v in g["meta/delays"]], 'cuda',
                                 FeedbackLoopConfig(use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs="out/svf_params" in g, num_hidden_layers=hidden,
                                                    num_neurons_per_layer=neurons, num_fourier_features=feats,
                                                    compress_pole_factor=float(g.get("meta/pole_factor", 1.0))),
                                 use_absorption_filters=geq,
                                 common_decay_times=g["meta/t60"] if geq else np.array([g["meta/t60"]]),
                                 band_centre_hz=list(g["meta/band_centre_hz"]) if geq else None,
                                 use_colorless_loss=True)
    state = {k[len("param/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("param/")}
    net.load_state_dict(state, strict=True)  # reference checkpoint layout must load unchanged
    return net


def omni_data(g):
    from diffgfdn_b200.utils import unit_circle_grid
    data = {k[len("data/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("data/")}
    data["z_values"] = unit_circle_grid(int(g["meta/nfft"]), float(g["meta/radius"]))
    data["target_rir_response"] = data["target_rir_response"].cuda()
    return data


def make_trainer(cls, net, tmp_path, **kw):
    from diffgfdn_b200.config import TrainerConfig
    return cls(net, TrainerConfig(train_dir=str(tmp_path / "out"), ir_dir=str(tmp_path / "ir"), **kw))


@pytest.mark.parametrize("name", list(OMNI))
def test_omni_forward_losses_grads_match_reference(name, tmp_path):
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    g = load(name)
    hidden, neurons, feats, _ = OMNI[name]
    net = build_omni(g, hidden, neurons, feats)
    data = omni_data(g)
    trainer = make_trainer(VarReceiverPosTrainer, net, tmp_path, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]))
    if "subband_filter" in data:
        trainer.set_subband_filter(data["subband_filter"])
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    d = g["data/target_early_response"]
    assert H.dtype == torch.complex64 and tuple(H.shape) == d.shape
    # the late (GFDN) part alone: forward without the direct path d (complex64 H cannot resolve a late part that
    # sits 1e5 below |d|, which is the case in this fixture; the reference keeps H in complex128 only because its
    # d input is complex128, quirk Q6)
    with torch.no_grad():
        H_late, _ = net({k: v for k, v in data.items() if k != "target_early_response"})
    # SVF case: the biquad coefficients are float32 on both sides and a0 + a1 z^-1 + a2 z^-2 cancels to ~4 f_c^2 near
    # DC, so one ulp of difference between the CPU and the GPU pow/sqrt shows up as ~1e-3 there (see
    # tests/test_oracle_golden.py); the kernel itself is checked to 1e-5 against the oracle in test_gpu_kernels.py
    assert rel(H_late.cpu().to(torch.complex128).numpy(), g["out/H"] - d) < (5e-3 if net.use_svf_in_output else 1e-4)
    if net.use_svf_in_output:
        assert rel(net.output_filters.svf_params, g["out/svf_params"]) < 1e-5
        assert rel(net.output_filters.biquad_coeffs_, g["out/biquads"]) < 1e-5
    assert rel(np.abs(H.detach().cpu().numpy()), np.abs(g["out/H"])) < (5e-3 if net.use_svf_in_output else 1e-4)
    assert rel(Hs, g["out/H_sub"]) < 1e-4
    assert rel(Hsd.detach().cpu().numpy()[:, ::16, :], g["out/H_sub_per_del_s16"]) < 1e-4
    assert rel(net.feedback_loop.coupled_feedback_matrix_real(), np.real(g["out/A"])) < 1e-4
    losses = trainer.calculate_losses(data, trainer.apply_subband_filter(H), (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    w_edc = float(g["meta/edc_w"])
    assert abs(float(losses["edc_loss"]) / w_edc - g["loss/edc_raw"]) < 0.01  # dB
    assert abs(float(losses["edr_loss"]) - g["loss/edr_loss"]) < 2e-3 * g["loss/edr_loss"]
    assert abs(float(losses["spectral_loss"]) - g["loss/spectral_loss"]) < 1e-4 * g["loss/spectral_loss"]
    assert abs(float(losses["sparsity_loss"]) - g["loss/sparsity_loss"]) < 1e-5
    assert abs(float(total) - g["loss/total"]) < 2e-3 * g["loss/total"]
    for k, p in net.named_parameters():
        assert rel(p.grad, g[f"grad/{k}"]) < 1e-3, k


@pytest.mark.parametrize("name", list(OMNI))
def test_omni_loss_callables_and_mask(name):
    from diffgfdn_b200.colorless_fdn.losses import amse_loss, mse_loss
    from diffgfdn_b200.losses import edc_loss, edr_loss
    g = load(name)
    o = oracle_omni(g, params_of(g))
    H = o["Huse"].to(torch.complex64).cuda()
    tgt = o["tgt"].to(torch.complex64).cuda()
    mx, fs = float(g["meta/max_ir_len_ms"]), float(g["meta/fs"])
    assert abs(float(edc_loss(mx, fs)(tgt, H)) - g["loss/edc_raw"]) < 0.01
    idx = torch.tensor(g["data/edc_mask_index"])
    assert abs(float(edc_loss(mx, fs)(tgt, H, mask_index=idx)) - g["loss/edc_masked_raw"]) < 0.01
    assert abs(float(edr_loss(fs)(tgt, H)) - g["loss/edr_raw"]) < 2e-3 * g["loss/edr_raw"]
    hs = o["H_sub"][:, 0].to(torch.complex64).cuda()
    assert abs(float(mse_loss()(hs, torch.ones_like(hs))) - g["loss/mse_g0"]) < 1e-5
    assert abs(float(amse_loss()(hs, torch.ones_like(hs))) - g["loss/amse_g0"]) < 1e-5
    # the random mask path draws from torch's CPU generator exactly like the reference (losses.py:221-223)
    torch.manual_seed(99)
    masked = float(edc_loss(mx, fs, use_mask=True)(tgt, H))
    assert abs(masked - g["loss/edc_masked_raw"]) < 0.01


@pytest.mark.parametrize("name", list(OMNI))
def test_omni_normalize_and_adam_steps_match_reference(name, tmp_path):
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    g = load(name)
    hidden, neurons, feats, steps = OMNI[name]
    net = build_omni(g, hidden, neurons, feats)
    data = omni_data(g)
    trainer = make_trainer(VarReceiverPosTrainer, net, tmp_path, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]), io_lr=0.01, lr=0.01)
    if "subband_filter" in data:
        trainer.set_subband_filter(data["subband_filter"])
    got = []
    for _ in range(steps):
        trainer.normalize(data)
        loss, _ = trainer.train_step(data)
        got.append(loss)
    assert np.allclose(got, g["steps/loss"], rtol=2e-3)
    for k, v in net.state_dict().items():
        ref = g[f"steps/param/{k}"]
        # Adam's first steps move every entry by ~lr * sign(grad): entries whose gradient is at the level of the
        # reference's float32 biquad-coefficient noise (SVF case, see above) may differ by a fraction of lr = 0.01
        slack = 1e-3 if net.use_svf_in_output else 2e-4
        assert float(np.abs(v.cpu().numpy() - ref).max()) < 5e-3 * max(1e-3, float(np.abs(ref).max())) + slack, k


@pytest.mark.parametrize("name", list(DIRECTIONAL))
def test_directional_forward_losses_grads_match_reference(name, tmp_path):
    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffDirectionalFDNVarReceiverPos
    from diffgfdn_b200.trainer import DirectionalFDNVarReceiverPosTrainer
    from diffgfdn_b200.utils import unit_circle_grid
    g = load(name)
    hidden, neurons, feats, skip = DIRECTIONAL[name]
    net = DiffDirectionalFDNVarReceiverPos(float(g["meta/fs"]), 3, [int(v) for v in g["meta/delays"]], 'cuda',
                                           FeedbackLoopConfig(use_zero_coupling=False),
                                           OutputFilterConfig(use_svfs=False, num_hidden_layers=hidden,
                                                              num_neurons_per_layer=neurons,
                                                              num_fourier_features=feats, use_skip_connections=skip),
                                           2, None, common_decay_times=np.array([g["meta/t60"]]),
                                           use_colorless_loss=True, analysis_matrix=g["data/Y"])
    net.load_state_dict({k[len("param/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("param/")},
                        strict=True)
    pos = torch.tensor(g["data/norm_listener_position"])
    data = dict(z_values=unit_circle_grid(int(g["meta/nfft"])), listener_position=pos, norm_listener_position=pos,
                target_common_slope_amps=torch.tensor(g["data/amps"]))
    trainer = make_trainer(DirectionalFDNVarReceiverPosTrainer, net, tmp_path, use_colorless_loss=True,
                           use_asym_spectral_loss=False, edc_loss_weight=float(g["meta/edc_w"]),
                           num_freq_bins=int(g["meta/nfft"]))
    net.zero_grad()
    H_sh, (Hs, _) = net(data)
    assert rel(H_sh, g["out/H_sh"]) < 1e-4
    H_dir = trainer.convert_ambi_rir_to_directional_rir(H_sh)
    assert rel(H_dir.detach().cpu().numpy()[..., ::4], g["out/H_dir_s4"]) < 1e-4
    losses = trainer.calculate_losses(data, H_dir, (Hs, None))
    total = sum(losses.values())
    total.backward()
    assert abs(float(losses["edc_loss"]) - g["loss/edc_loss"]) < 0.01 * float(g["meta/edc_w"])
    assert abs(float(total) - g["loss/total"]) < 1e-3 * g["loss/total"]
    for k, p in net.named_parameters():
        assert rel(p.grad, g[f"grad/{k}"]) < 1e-3, k
    # oracle agrees on the same inputs (ties the oracle, the reference fixture and the kernels together)
    o = oracle_directional(g, params_of(g))
    assert rel(H_sh, o["H_sh"]) < 1e-4


@pytest.mark.parametrize("name", ["src_rx_n12", "single_n12", "single_n12_svf"])
def test_source_receiver_and_single_position_variants_match_reference(name, tmp_path):
    """SURVEY a-8c: DiffGFDNVarSourceReceiverPos (model.py:402-452) and DiffGFDNSinglePos (:779-908) on the K1
    group-to-group transfer functions + the projection kernels."""
    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNSinglePos, DiffGFDNVarSourceReceiverPos
    from diffgfdn_b200.trainer import SinglePosTrainer, VarReceiverPosTrainer
    from diffgfdn_b200.utils import unit_circle_grid
    from golden_util import oracle_variant
    g = load(name)
    delays = [int(v) for v in g["meta/delays"]]
    common = dict(use_absorption_filters=False, common_decay_times=np.array([g["meta/t60"]]), use_colorless_loss=True)
    if name.startswith("src_rx"):
        ofc = OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16, num_fourier_features=4)
        net = DiffGFDNVarSourceReceiverPos(float(g["meta/fs"]), 3, delays, 'cuda', FeedbackLoopConfig(use_zero_coupling=False),
                                           ofc, ofc, learn_common_decay_times=False, **common)
        tcls = VarReceiverPosTrainer
    else:
        pf = float(g["meta/pole_factor"])
        net = DiffGFDNSinglePos(float(g["meta/fs"]), 3, delays, 'cuda', FeedbackLoopConfig(use_zero_coupling=False),
                                OutputFilterConfig(use_svfs=bool(g["meta/svf_out"]), compress_pole_factor=pf),
                                input_filter_config=OutputFilterConfig(use_svfs=bool(g["meta/svf_in"]),
                                                                       compress_pole_factor=pf), **common)
        tcls = SinglePosTrainer
    net.load_state_dict({k[len("param/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("param/")},
                        strict=True)
    data = {k[len("data/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("data/")}
    data["z_values"] = unit_circle_grid(int(g["meta/nfft"]))
    data["target_rir_response"] = data["target_rir_response"].cuda()
    trainer = make_trainer(tcls, net, tmp_path, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]))
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    d = g["data/target_early_response"]
    svf = name.endswith("_svf")
    assert tuple(H.shape) == d.shape and H.dtype == torch.complex64
    assert rel(H.detach().cpu().to(torch.complex128).numpy() - d, g["out/H"] - d) < (5e-3 if svf else 1e-4)
    assert rel(Hs, g["out/H_sub"]) < 1e-4
    po = params_of(g, requires_grad=True)
    o = oracle_variant(g, po)
    assert rel(H.detach().cpu().to(torch.complex128).numpy() - d, o["H"].detach().numpy() - d) < 1e-4  # float64 oracle
    losses = trainer.calculate_losses(data, H, (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    assert abs(float(losses["edc_loss"]) - g["loss/edc_loss"]) < 0.01 * float(g["meta/edc_w"])
    assert abs(float(total) - g["loss/total"]) < 2e-3 * g["loss/total"]
    if svf:  # the reference cannot differentiate this branch (oracle/gen_golden.py): gradients against the oracle
        o["total"].backward()
    for k, p in net.named_parameters():
        ref = po[k].grad if svf else g[f"grad/{k}"]
        assert rel(p.grad, ref) < 1e-3, k


@pytest.mark.parametrize("name", ["src_rx_n12_svf", "src_rx_n12_svf_in"])
def test_source_receiver_svf_cascades_match_reference(name, tmp_path):
    """DiffGFDNVarSourceReceiverPos with SVF cascades on the source side (reference model.py:376-388, 432-449), with
    gains or cascades on the receiver side: forward, losses and gradients vs the reference golden and the oracle."""
    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarSourceReceiverPos
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    from diffgfdn_b200.utils import unit_circle_grid
    from golden_util import oracle_variant
    g = load(name)
    pf = float(g["meta/pole_factor"])
    mk = lambda svf: OutputFilterConfig(use_svfs=bool(svf), num_hidden_layers=1, num_neurons_per_layer=16,  # noqa: E731
                                        num_fourier_features=int(g["meta/feats"]), compress_pole_factor=pf)
    net = DiffGFDNVarSourceReceiverPos(float(g["meta/fs"]), 3, [int(v) for v in g["meta/delays"]], 'cuda',
                                       FeedbackLoopConfig(use_zero_coupling=False), mk(g["meta/svf_out"]),
                                       mk(g["meta/svf_in"]), use_absorption_filters=False, learn_common_decay_times=False,
                                       common_decay_times=np.array([g["meta/t60"]]), use_colorless_loss=True)
    net.load_state_dict({k[len("param/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("param/")},
                        strict=True)
    data = {k[len("data/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("data/")}
    data["z_values"] = unit_circle_grid(int(g["meta/nfft"]))
    data["target_rir_response"] = data["target_rir_response"].cuda()
    trainer = make_trainer(VarReceiverPosTrainer, net, tmp_path, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]))
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    d = g["data/target_early_response"]
    assert tuple(H.shape) == d.shape and H.dtype == torch.complex64
    Hn = H.detach().cpu().to(torch.complex128).numpy()
    assert rel(Hn - d, g["out/H"] - d) < 5e-3  # the reference's float32 biquad coefficients
    po = params_of(g, requires_grad=True)
    o = oracle_variant(g, po)
    assert rel(Hn - d, o["H"].detach().numpy() - d) < 1e-4  # float64 oracle
    assert rel(Hs, g["out/H_sub"]) < 1e-4
    losses = trainer.calculate_losses(data, H, (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    assert abs(float(total) - g["loss/total"]) < 2e-3 * g["loss/total"]
    o["total"].backward()
    for k, p in net.named_parameters():
        assert rel(p.grad, po[k].grad) < 1e-3, k          # float64 oracle
        assert rel(p.grad, g[f"grad/{k}"]) < 5e-3, k      # reference (float32 coefficient noise)
    pd = net.get_param_dict_inference(data)
    assert "input_svf_params" in pd and "input_biquad_coeffs" in pd


def test_random_coupling_matches_reference(tmp_path):
    """coupling_matrix_type: random_matrix (the single-room sub-band YAML): forward, losses and gradients vs the
    reference golden; checkpoint key 'feedback_loop.random_feedback_matrix'."""
    from diffgfdn_b200.config import CouplingMatrixType, FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    g = load("random_coupling_n8")
    net = DiffGFDNVarReceiverPos(float(g["meta/fs"]), len(g["meta/t60"]), [int(v) for v in g["meta/delays"]], 'cuda',
                                 FeedbackLoopConfig(coupling_matrix_type=CouplingMatrixType.RANDOM),
                                 OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16,
                                                    num_fourier_features=4),
                                 use_absorption_filters=False, common_decay_times=np.array([g["meta/t60"]]),
                                 use_colorless_loss=False)
    net.load_state_dict({k[len("param/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("param/")},
                        strict=True)
    data = omni_data({**g, "meta/radius": 1.0})
    trainer = make_trainer(VarReceiverPosTrainer, net, tmp_path, use_colorless_loss=False, edc_loss_weight=10.0,
                           num_freq_bins=int(g["meta/nfft"]))
    net.zero_grad()
    H = net(data)
    d = g["data/target_early_response"]
    assert rel(H.detach().cpu().to(torch.complex128).numpy() - d, g["out/H"] - d) < 1e-4
    assert rel(net.feedback_loop.coupled_feedback_matrix, g["out/A"]) < 1e-5
    losses = trainer.calculate_losses(data, H)
    total = sum(losses.values())
    total.backward()
    assert abs(float(losses["edc_loss"]) - g["loss/edc_loss"]) < 0.01 * float(g["meta/edc_w"])
    assert abs(float(total) - g["loss/total"]) < 2e-3 * g["loss/total"]
    for k, p in net.named_parameters():
        assert rel(p.grad, g[f"grad/{k}"]) < 1e-3, k
    assert set(net.feedback_loop.get_param_dict()) >= {"coupled_feedback_matrix", "delay_line_gains"}


def test_filter_coupling_matches_reference(tmp_path):
    """coupling_matrix_type: filter_matrix (reference feedback_loop.py:90-143, 362-373, 447-453): paraunitary FIR coupling,
    A(z_k) per bin. Forward, losses and every gradient vs the reference golden; reference state_dict loads strictly."""
    from diffgfdn_b200.config import CouplingMatrixType, FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    g = load("filter_coupling_n12")
    net = DiffGFDNVarReceiverPos(float(g["meta/fs"]), len(g["meta/t60"]), [int(v) for v in g["meta/delays"]], 'cuda',
                                 FeedbackLoopConfig(coupling_matrix_type=CouplingMatrixType.FILTER,
                                                    pu_matrix_order=int(g["meta/order"]), use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16,
                                                    num_fourier_features=4),
                                 use_absorption_filters=False, common_decay_times=np.array([g["meta/t60"]]),
                                 use_colorless_loss=True)
    net.load_state_dict({k[len("param/"):]: torch.tensor(v) for k, v in g.items() if k.startswith("param/")},
                        strict=True)
    data = omni_data({**g, "meta/radius": 1.0})
    trainer = make_trainer(VarReceiverPosTrainer, net, tmp_path, use_colorless_loss=True, use_asym_spectral_loss=True,
                           edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]))
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    d = g["data/target_early_response"]
    assert rel(H.detach().cpu().to(torch.complex128).numpy() - d, g["out/H"] - d) < 1e-4
    assert rel(Hs, g["out/H_sub"]) < 1e-4
    assert rel(net.feedback_loop.phi, g["out/phi"]) < 1e-5
    assert rel(net.feedback_loop.coupled_feedback_matrix, g["out/A"].real) < 1e-5
    losses = trainer.calculate_losses(data, H, (Hs, Hsd))
    total = sum(losses.values())
    total.backward()
    assert abs(float(losses["edc_loss"]) - g["loss/edc_loss"]) < 0.01 * float(g["meta/edc_w"])
    assert abs(float(total) - g["loss/total"]) < 2e-3 * g["loss/total"]
    for k, p in net.named_parameters():
        assert rel(p.grad, g[f"grad/{k}"]) < 1e-3, k
    pd = net.feedback_loop.get_param_dict()
    assert set(pd) >= {"coupled_feedback_matrix", "unitary_matrix", "unit_vectors", "coupling_matrix"}
    assert pd["coupled_feedback_matrix"].shape == g["out/A"].shape


def test_feedback_loop_dense_inverse_api():
    """FeedbackLoop.forward(z) keeps the reference's (K, N, N) inverse for API compatibility."""
    from oracle import gfdn_oracle as O
    g = load("omni_n12")
    net = build_omni(g, *OMNI["omni_n12"][:3])
    z = O.z_grid(256)
    P = net.feedback_loop(z.cuda())
    a = net.feedback_loop.coupled_feedback_matrix_real().detach().cpu().to(torch.float64)
    Po = O.feedback_loop_inverse(z, torch.tensor(g["meta/delays"], dtype=torch.float64),
                                 net.feedback_loop.delay_line_gains.cpu().to(torch.float64), a)
    assert rel(P, Po) < 1e-5
