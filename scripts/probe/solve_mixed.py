"""K1 forward: all-float64 kernel vs float32 elimination + float64 refinement (DGFDN_SOLVE_MIXED_FORCE=0/1) at the bench
shape: accuracy against a complex128 torch.linalg.solve on the device, and time per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from diffgfdn_b200 import ops  # noqa: E402
from diffgfdn_b200.utils import unit_circle_grid  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    net = bench.build_net(dev)
    fl = net.feedback_loop
    z = unit_circle_grid(bench.NFFT, device=dev)[:bench.NFFT // 4 + 1]
    k = z.numel()
    with torch.no_grad():
        a = fl._assembled().to(torch.float32).contiguous()
        gamma = fl.delay_line_gains
        b, c = net.input_gains.reshape(-1), net.output_gains.reshape(-1)
        delays = fl.delays.to(torch.int32)
        # float64 truth on a subset of bins (every 16th)
        idx = torch.arange(0, k, 16, device=dev)
        zz = z[idx].to(torch.complex128)
        d = zz.unsqueeze(-1)**fl.delays.to(torch.float64) / gamma.to(torch.complex128)
        m = torch.diag_embed(d) - a.to(torch.complex128)
        xt = torch.linalg.solve(m, b.to(torch.complex128).expand(idx.numel(), -1).unsqueeze(-1)).squeeze(-1)
        cond = torch.linalg.cond(m)
        print(f"bins {k}, N {a.shape[0]}, cond(M): median {float(cond.median()):.1f}, max {float(cond.max()):.1f}")
        for mode in ("0", "1"):
            os.environ["DGFDN_SOLVE_MIXED_FORCE"] = mode
            x, y = ops.gfdn_solve(z, delays, a, gamma, b, c, net.num_groups)
            err = (x[idx].to(torch.complex128) - xt).abs().amax(-1) / xt.abs().amax(-1)
            for _ in range(3):
                ops.gfdn_solve(z, delays, a, gamma, b, c, net.num_groups)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.gfdn_solve(z, delays, a, gamma, b, c, net.num_groups)
            e1.record()
            torch.cuda.synchronize()
            print(f"mixed={mode}: x rel err vs float64 solve: median {float(err.median()):.2e}, max {float(err.max()):.2e}; "
                  f"{e0.elapsed_time(e1) / 10:.3f} ms per launch (no saved factors)")


if __name__ == "__main__":
    main()
