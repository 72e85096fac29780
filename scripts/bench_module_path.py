"""Module-path step at the reference's own C1 shape (BASELINE configs[0]: N=12, G=3, B=32 receivers per batch,
nfft=131072 => K=65537 bins, EDC + EDR + colorless losses, Adam): time per train_step and a kernel table.
    python scripts/bench_module_path.py [--svf] > gpurun_out/module_path.txt"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--svf", action="store_true")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--nfft", type=int, default=131072)
    a = ap.parse_args()
    from diffgfdn_b200.config import DiffGFDNConfig, FeedbackLoopConfig, OutputFilterConfig, TrainerConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    from diffgfdn_b200.trainer import VarReceiverPosTrainer
    from diffgfdn_b200.utils import unit_circle_grid
    torch.manual_seed(0)
    fs, t60 = 32000.0, np.array([[0.3, 0.8, 1.5]])
    delays = DiffGFDNConfig(seed=235265, num_delay_lines=12).delay_length_samps
    net = DiffGFDNVarReceiverPos(fs, 3, delays, "cuda", FeedbackLoopConfig(use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs=a.svf, num_hidden_layers=3, num_neurons_per_layer=128,
                                                    num_fourier_features=20), use_absorption_filters=False,
                                 common_decay_times=t60, use_colorless_loss=True)
    tmp = tempfile.mkdtemp()
    tr = VarReceiverPosTrainer(net, TrainerConfig(train_dir=tmp + "/o", ir_dir=tmp + "/i", num_freq_bins=a.nfft,
                                                  use_colorless_loss=True, use_asym_spectral_loss=True, edc_loss_weight=10.0))
    k = a.nfft // 2 + 1
    gen = torch.Generator(device="cuda").manual_seed(1)
    t = torch.arange(a.nfft // 2, device="cuda")
    rir = torch.randn(a.batch, a.nfft // 2, device="cuda", generator=gen) * torch.exp(-t / 6000.0)
    early = rir.clone()
    early[:, 640:] = 0
    pos = torch.rand(a.batch, 3, device="cuda", generator=gen)
    data = dict(z_values=unit_circle_grid(a.nfft).cuda(), listener_position=pos, norm_listener_position=pos,
                target_early_response=torch.fft.rfft(early, n=a.nfft).to(torch.complex64),
                target_rir_response=torch.fft.rfft(rir, n=a.nfft).to(torch.complex64))
    for _ in range(3):
        tr.normalize(data)
        tr.train_step(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        tr.normalize(data)
        tr.train_step(data)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"module path, svf={a.svf}: {ms:.3f} ms per normalize+train_step, {a.batch * k / (ms * 1e-3):.3e} receiver*bin evals/s "
          f"(B={a.batch}, K={k}, N=12)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            tr.normalize(data)
            tr.train_step(data)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=80))


if __name__ == "__main__":
    main()
