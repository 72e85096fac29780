"""Per-parameter gradient parity of the fused step against the oracle after K captured Adam steps at the BASELINE
shape (diagnostic for bench.py's cpu_baseline.parity_vs_gpu_step).  usage: python scripts/diag_parity_after_training.py [steps] [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from diffgfdn_b200.fused import ShardedEDCStep  # noqa: E402
from diffgfdn_b200.utils import unit_circle_grid  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda")
    nfft = 2**18
    net = bench.build_net(dev)
    z = unit_circle_grid(nfft, device=dev)
    gen = torch.Generator(device=dev).manual_seed(100)
    pos = torch.rand(1024, 3, device=dev, generator=gen)
    early, target = bench.synth_responses(1024, nfft, dev, 200)
    step = ShardedEDCStep(net, 1500.0, edc_weight=10.0)
    step.attach(z, pos, None, None)
    step.attach(z, pos, step.precompute_early_window(early), step.precompute_target_db(target))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
    if steps:
        step.capture(optimizer=opt, warmup=1)
        for _ in range(steps):
            step.replay()
        torch.cuda.synchronize()
        step.release_graph()
    params = bench.oracle_params(net)
    losses, secs = bench.oracle_step(params, net.delays.cpu().double(), z.cpu(), pos[:rows].cpu().double(),
                                     early[:rows].cpu().to(torch.complex128), target[:rows].cpu().to(torch.complex128),
                                     net.output_scalars.encoder.num_fourier_features)
    for replay in ("1", "0"):
        os.environ["DGFDN_SOLVE_REPLAY"] = replay
        sub = ShardedEDCStep(net, 1500.0, edc_weight=10.0)
        sub.attach(z, pos[:rows], None, None)
        sub.attach(z, pos[:rows], sub.precompute_early_window(early[:rows]), sub.precompute_target_db(target[:rows]))
        out = sub.step()
        torch.cuda.synchronize()
        print(f"replay={replay} edc gpu {float(out['edc_loss']) / 10:.6f} oracle {losses['edc']:.6f}")
        for k, q in net.named_parameters():
            g, o = q.grad.detach().cpu().double(), params[k].grad
            print(f"   {k:45s} rel {float((g - o).abs().max() / o.abs().max()):.2e}  |g|max {float(o.abs().max()):.3e}")


if __name__ == "__main__":
    main()
