"""Diagnostic (GPU): per-loss gradient error of the CUDA path vs the float64 oracle on a golden fixture."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, warnings
warnings.filterwarnings("ignore")
from golden_util import load, params_of
from oracle import gfdn_oracle as O
from test_gpu_model_golden import build_omni, omni_data, OMNI
from diffgfdn_b200.losses import edc_loss, edr_loss

name = sys.argv[1] if len(sys.argv) > 1 else "omni_n12_subband_r"
g = load(name)
fs = float(g["meta/fs"]); mx = float(g["meta/max_ir_len_ms"])
F = torch.tensor(g["data/subband_filter"]) if "data/subband_filter" in g else None

def oracle_grads(which):
    p = params_of(g, requires_grad=True)
    nfft = int(g["meta/nfft"])
    z = O.z_grid(nfft, float(g["meta/radius"])); delays = torch.tensor(g["meta/delays"], dtype=torch.float64)
    gamma = O.decay_times_to_gain_per_sample(g["meta/t60"], g["meta/delays"], fs, 3)
    A = O.coupled_feedback_matrix(p["feedback_loop.M"], p["feedback_loop.alpha"])
    b = p["input_gains"].reshape(-1); c = p["output_gains"].reshape(-1)
    s = O.gains_from_mlp(torch.tensor(g["data/norm_listener_position"]), p, int(g["meta/feats"]), 3)
    H = O.omni_response(z, delays, gamma, A, b, c, s, torch.tensor(g["data/target_early_response"]))
    if F is not None: H = H * F
    tgt = torch.tensor(g["data/target_rir_response"])
    loss = O.edc_loss(tgt, H, mx, fs) if which == "edc" else O.edr_loss(tgt, H)
    loss.backward()
    return float(loss), {k: v.grad for k, v in p.items() if v.grad is not None}

def gpu_grads(which):
    net = build_omni(g, *OMNI[name][:3]); data = omni_data(g)
    H, _ = net(data)
    if F is not None: H = H * F.to(torch.complex64).cuda()
    tgt = data["target_rir_response"].to(torch.complex64)
    loss = edc_loss(mx, fs)(tgt, H) if which == "edc" else edr_loss(fs)(tgt, H)
    loss.backward()
    return float(loss), {k: p.grad.cpu().double() for k, p in net.named_parameters() if p.grad is not None}

for which in ("edc", "edr"):
    lo, go = oracle_grads(which); lg, gg = gpu_grads(which)
    print(which, "loss oracle", lo, "gpu", lg)
    for k in ("feedback_loop.alpha", "feedback_loop.M", "input_gains", "output_gains", "output_scalars.mlp.model.0.weight"):
        e = float((gg[k] - go[k]).abs().max()); m = float(go[k].abs().max())
        print(f"   {k:40s} abs err {e:.3e}  max|g| {m:.3e}  rel {e/m:.3e}")
