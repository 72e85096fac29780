"""Kernel-level table of a few training steps (torch.profiler / CUPTI) on the bench workload. Run on the GPU box:
    python scripts/profile_step.py [--receivers N] [--tile-rows R] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--receivers", type=int, default=12500)
    ap.add_argument("--nfft", type=int, default=bench.NFFT)
    ap.add_argument("--tile-rows", type=int, default=12500)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    from diffgfdn_b200.fused import ShardedEDCStep
    from diffgfdn_b200.utils import unit_circle_grid
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    net = bench.build_net(dev)
    step = ShardedEDCStep(net, max(bench.T60) * 1e3, tile_rows=a.tile_rows, edc_weight=10.0)
    z = unit_circle_grid(a.nfft, device=dev)
    pos = torch.rand(a.receivers, 3, device=dev)
    step.attach(z, pos, None, None)
    early, tgt = bench.synth_responses(min(a.receivers, 256), a.nfft, dev, 1)
    reps = (a.receivers + early.shape[0] - 1) // early.shape[0]
    hd = step.precompute_early_window(early).repeat(reps, 1)[:a.receivers].contiguous()
    tdb = step.precompute_target_db(tgt).repeat(reps, 1)[:a.receivers].contiguous()
    step.attach(z, pos, hd, tdb)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    for _ in range(3):
        step.step()
        opt.step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step.step()
            opt.step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
