"""K7 (csrc/mlp.cu): position -> gain network kernels against the float64 oracle (oracle/gfdn_oracle.py, a
restatement of reference dnn.py / gain_filters.py / spatial_sampling/model.py) and against the same arithmetic in
stock float32 PyTorch, forward and every parameter gradient."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    # (rows, fourier features, hidden layers, neurons, skip connections, position dtype)
    return [(1, 10, 3, 128, False, torch.float64), (97, 10, 3, 128, False, torch.float32),
            (500, 10, 1, 64, False, torch.float64), (193, 4, 2, 64, False, torch.float32),
            (300, 10, 2, 128, True, torch.float64), (96, 8, 0, 128, False, torch.float64)]


def _randomise(mod, seed):
    """LayerNorm affine parameters and biases away from their (1, 0) initial values so every gradient path is live."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in mod.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight.copy_(1.0 + 0.3 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.3 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, torch.nn.Linear):
                m.bias.copy_(0.3 * torch.randn(m.bias.shape, generator=g))


@pytest.mark.parametrize("rows,nfeat,hidden,neurons,skip,pdtype", _cases())
def test_gains_from_mlp_matches_torch_and_oracle(rows, nfeat, hidden, neurons, skip, pdtype):
    from diffgfdn_b200 import dnn
    from diffgfdn_b200.gain_filters import Gains_from_MLP
    from diffgfdn_b200.sh_gains import Directional_Beamforming_Weights_from_MLP
    from oracle import gfdn_oracle as O
    torch.manual_seed(rows + nfeat)
    dev = torch.device("cuda")
    groups = 3
    if skip:
        mod = Directional_Beamforming_Weights_from_MLP(groups, 2, nfeat, hidden, neurons, device=dev,
                                                       use_skip_connections=True,
                                                       analysis_matrix=np.ones((12, 9), dtype=np.float32)).to(dev)
    else:
        mod = Gains_from_MLP(groups, 4, nfeat, hidden, neurons, device=dev).to(dev)
    _randomise(mod, 5)
    # seed 13: seeds 1 and 11 put one ReLU input of a float32-position case (97 / 193 rows) within rounding of zero,
    # and the float32 formulation itself (stock torch on the CPU included: 6e-3) then differs from the float64
    # oracle by a flipped unit; seed 13 keeps every case 1e3 x clear of that (checked with float32 torch on the CPU)
    pos = torch.rand(rows, 3, dtype=pdtype, generator=torch.Generator().manual_seed(13)).to(dev)
    x = {"norm_listener_position": pos}
    wgt = torch.randn(rows, groups * (9 if skip else 1), generator=torch.Generator().manual_seed(2)).to(dev)

    def run(fused):
        for p in mod.parameters():
            p.grad = None
        saved = dnn.fused_position_mlp
        if not fused:
            import diffgfdn_b200.gain_filters as gf
            import diffgfdn_b200.sh_gains as sg
            gf.fused_position_mlp = sg.fused_position_mlp = lambda *a, **k: None
        try:
            out = mod(x).reshape(rows, -1) if skip else mod.gains(x)
        finally:
            if not fused:
                gf.fused_position_mlp = sg.fused_position_mlp = saved
        (out * wgt).sum().backward()
        return out.detach(), {n: p.grad.clone() for n, p in mod.named_parameters()}

    out_k, g_k = run(True)
    out_t, g_t = run(False)
    assert out_k.shape == out_t.shape
    # oracle (float64 CPU)
    w64 = {n: v.detach().cpu().to(torch.float64).requires_grad_(True) for n, v in mod.state_dict().items()}
    if skip:
        out_o = O.sh_gains_from_mlp(pos.cpu(), w64, nfeat, groups, 9, prefix="mlp.", skip=True, normalise=False)
    else:
        out_o = O.gains_from_mlp(pos.cpu(), w64, nfeat, groups, prefix="mlp.model.")
    out_o = out_o.reshape(rows, -1)
    (out_o * wgt.cpu().to(torch.float64)).sum().backward()
    scale = float(out_o.abs().max())
    err_k = float((out_k.cpu().to(torch.float64) - out_o).abs().max()) / scale
    err_t = float((out_t.cpu().to(torch.float64) - out_o).abs().max()) / scale
    assert err_k < 1e-4, f"kernel forward vs oracle: {err_k}"  # north_star: 1e-4 relative
    assert err_k < 4 * err_t + 2e-6, f"kernel forward is less accurate than float32 torch: {err_k} vs {err_t}"
    for n, go in ((n, v.grad) for n, v in w64.items() if v.grad is not None):
        den = float(go.abs().max()) + 1e-30
        ek = float((g_k[n].cpu().to(torch.float64) - go).abs().max()) / den
        et = float((g_t[n].cpu().to(torch.float64) - go).abs().max()) / den
        ekt = float((g_k[n] - g_t[n]).abs().max()) / den
        assert ekt < 2e-5, f"grad {n}: kernel vs float32 torch {ekt}"
        assert ek < 1e-3, f"grad {n}: kernel vs oracle {ek}"  # north_star: gradients within 1e-3 relative
        assert ek < 8 * et + 2e-5, f"grad {n}: kernel {ek} vs float32 torch {et}"


def test_mlp_kernel_is_used_and_deterministic():
    from diffgfdn_b200.gain_filters import Gains_from_MLP
    dev = torch.device("cuda")
    torch.manual_seed(0)
    mod = Gains_from_MLP(3, 8, 10, 3, 128, device=dev).to(dev)
    pos = torch.rand(1000, 3, device=dev, dtype=torch.float64)
    outs, grads = [], []
    for _ in range(2):
        for p in mod.parameters():
            p.grad = None
        s = mod.gains({"norm_listener_position": pos})
        assert type(s.grad_fn).__name__.startswith("_PositionMLP"), "the K7 kernel must be on the path"
        s.square().sum().backward()
        outs.append(s.detach().clone())
        grads.append(torch.cat([p.grad.reshape(-1) for p in mod.parameters()]))
    assert torch.equal(outs[0], outs[1]) and torch.equal(grads[0], grads[1])
    assert float(s.min()) > -1.0 and float(s.max()) < 1.0
