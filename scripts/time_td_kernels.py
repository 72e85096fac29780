"""Time dgfdn_td_edc_fused alone (no other stream) with the time-sliced kernel K3t and the cluster kernel K3d at the
BASELINE shard shape.  usage (GPU box): python scripts/time_td_kernels.py [rows] [tn] [kernels=sliced,cluster] [reps]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffgfdn_b200 import _lib, ops  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
    tn = int(sys.argv[2]) if len(sys.argv) > 2 else 47360
    kernels = sys.argv[3].split(",") if len(sys.argv) > 3 else ["sliced:0", "sliced:1", "cluster"]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    g = 3
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(3)
    decay = torch.exp(-torch.arange(tn, device=dev) / (0.25 * tn))
    hy = torch.randn(g, tn, device=dev, generator=gen) * decay
    s = torch.randn(rows, g, device=dev, generator=gen)
    hd = torch.randn(rows, tn, device=dev, generator=gen) * decay * 0.1
    tdb = ops.edc_db(torch.randn(rows, tn, device=dev, generator=gen) * decay)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    res = {}
    for kern in kernels:
        os.environ["DGFDN_TD_KERNEL"] = kern.split(":")[0]  # "sliced:1" = sliced kernel, pipeline configuration 1
        os.environ["DGFDN_TD_SLICED_CFG"] = kern.split(":")[1] if ":" in kern else "0"
        info = ops.td_fused_info(g, tn)
        ws = ops.td_fused_workspace(g, rows, tn, dev)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        gs = torch.empty(rows, g, device=dev)
        ghy = torch.empty(g, tn, device=dev)

        def launch():
            _lib.call("dgfdn_td_edc_fused", g, rows, tn, p(s), p(hy), p(hd), tn, p(tdb), tn, None, ctypes.c_double(1e-6),
                      p(loss), p(gs), p(ghy), 0, p(ws), st)

        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[kern] = dict(info=info, ms=ms, gbps=8.0 * rows * tn / ms / 1e6, loss=float(loss), gs=float(gs.abs().sum()),
                         ghy=float(ghy.abs().sum()))
        print(kern, json.dumps(res[kern]), flush=True)


if __name__ == "__main__":
    main()
