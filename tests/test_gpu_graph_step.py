"""The CUDA-graph replay of a whole module-path training step (diffgfdn_b200/trainer.py: normalize + forward + losses +
backward + Adam captured per batch signature) must reproduce the eager steps: same loss trajectory, same parameters.
Fixtures and trainer configuration are those of the reference-generated goldens (tests/golden/*.npz)."""
import copy

import numpy as np
import pytest
import torch

import test_gpu_model_golden as G
from golden_util import load

pytestmark = pytest.mark.gpu


def _run(name, tmp_path, monkeypatch, graph, steps=6, batches=2):
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    monkeypatch.setenv("DGFDN_GRAPH_STEP", "1" if graph else "0")
    g = load(name)
    hidden, neurons, feats, _ = G.OMNI[name]
    net = G.build_omni(g, hidden, neurons, feats)
    data = G.omni_data(g)
    trainer = G.make_trainer(VarReceiverPosTrainer, net, tmp_path, use_colorless_loss=True, use_asym_spectral_loss=True,
                             edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]), io_lr=0.01, lr=0.01)
    assert trainer.use_cuda_graph == graph
    if "subband_filter" in data:
        trainer.set_subband_filter(data["subband_filter"])
    # two different batches of the same shape, alternating: a replay must see the batch it was given
    gen = torch.Generator().manual_seed(1)
    other = dict(data)
    for k in ("listener_position", "norm_listener_position"):
        other[k] = torch.rand(data[k].shape, generator=gen, dtype=data[k].dtype)
    other["target_rir_response"] = torch.roll(data["target_rir_response"], 1, dims=0).clone()
    other["target_early_response"] = torch.roll(data["target_early_response"], 1, dims=0).clone()
    losses, parts = [], []
    for i in range(steps):
        d = data if i % batches == 0 else other
        loss, all_losses = trainer.train_step(d, with_norm=not net.use_svf_in_output)
        losses.append(loss)
        parts.append({k: float(v) for k, v in all_losses.items()})
    return losses, parts, {k: v.detach().clone() for k, v in net.state_dict().items()}, trainer


@pytest.mark.parametrize("name", ["omni_n12", "omni_n12_subband_r", "omni_n12_svf"])
def test_graph_replayed_steps_equal_eager_steps(name, tmp_path, monkeypatch):
    l_e, p_e, s_e, _ = _run(name, tmp_path / "e", monkeypatch, graph=False)
    l_g, p_g, s_g, tr = _run(name, tmp_path / "g", monkeypatch, graph=True)
    assert any(isinstance(v, dict) for v in tr._graphs.values()), "no step was captured"
    assert np.allclose(l_g, l_e, rtol=1e-4), (l_g, l_e)
    for a, b in zip(p_g, p_e):
        for k in b:
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k])), k
    for k in s_e:
        assert float((s_g[k] - s_e[k]).abs().max()) <= 1e-4 * max(1e-2, float(s_e[k].abs().max())), k


def test_learning_rate_change_recaptures(tmp_path, monkeypatch):
    """StepLR changes the learning rates (python floats baked into the captured Adam kernels): the next step must not
    replay the old graph."""
    l_g, _, _, tr = _run("omni_n12", tmp_path, monkeypatch, graph=True, steps=4, batches=1)
    key = next(k for k, v in tr._graphs.items() if isinstance(v, dict))
    old = tr._graphs[key]["graph"]
    for grp in tr.optimizer.param_groups:
        grp["lr"] = grp["lr"] * 0.1
    g = load("omni_n12")
    data = G.omni_data(g)
    before = {k: v.detach().clone() for k, v in tr.net.state_dict().items()}
    tr.train_step(data, with_norm=True)   # signature known, rates changed: eager step on static buffers + re-capture
    assert tr._graphs[key]["graph"] is not old
    step = max(float((v - before[k]).abs().max()) for k, v in tr.net.state_dict().items() if v.dtype.is_floating_point)
    assert np.isfinite(step) and step < 5e-2  # an Adam update at the reduced rates (+ the normalisation of b, c)
