"""Per-parameter gradient errors of the module path against the reference goldens (debug aid, GPU)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, tempfile, pathlib
from golden_util import load, oracle_omni, params_of
import test_gpu_model_golden as T
from diffgfdn_b200.trainer import VarReceiverPosTrainer
for name in ("omni_n12", "omni_n12_svf"):
    g = load(name)
    hidden, neurons, feats, _ = T.OMNI[name]
    net = T.build_omni(g, hidden, neurons, feats)
    data = T.omni_data(g)
    tr = T.make_trainer(VarReceiverPosTrainer, net, pathlib.Path(tempfile.mkdtemp()), use_colorless_loss=True,
                        use_asym_spectral_loss=True, edc_loss_weight=10.0, num_freq_bins=int(g["meta/nfft"]))
    net.zero_grad()
    H, (Hs, Hsd) = net(data)
    losses = tr.calculate_losses(data, H, (Hs, Hsd))
    sum(losses.values()).backward()
    p = params_of(g, requires_grad=True)
    o = oracle_omni(g, p)
    o["total"].backward()
    for k, prm in net.named_parameters():
        print(name, k, "vs ref %.2e" % T.rel(prm.grad, g[f"grad/{k}"]), "vs oracle %.2e" % T.rel(prm.grad, p[k].grad),
              "oracle vs ref %.2e" % T.rel(p[k].grad, g[f"grad/{k}"]))
