"""Time every tile variant of dgfdn_td_edc_fused (K3d) at the BASELINE shard shape and check it against K3c.
usage (GPU box): python scripts/tune_td_fused.py [rows] [tn]"""
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
    g = 3
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(3)
    decay = torch.exp(-torch.arange(tn, device=dev) / (0.25 * tn))
    hy = torch.randn(g, tn, device=dev, generator=gen) * decay
    s = torch.randn(rows, g, device=dev, generator=gen)
    hd = torch.randn(rows, tn, device=dev, generator=gen) * decay * 0.1
    tdb = ops.edc_db(torch.randn(rows, tn, device=dev, generator=gen) * decay)
    # reference: K3c on the first 64 rows
    n_ref = min(rows, 64)
    s_o, hy_o = s[:n_ref].clone().requires_grad_(True), hy.clone().requires_grad_(True)
    ref = ops.td_edc_abs_db_sum(s_o, hy_o, hd[:n_ref], tdb[:n_ref], None, tile_rows=32)
    ref.backward()
    out = {}
    for v in range(12):
        os.environ["DGFDN_TD_VARIANT"] = str(v)
        info = ops.td_fused_info(g, tn)
        if info["variant"] != v:
            continue
        s_c, hy_c = s[:n_ref].clone().requires_grad_(True), hy.clone().requires_grad_(True)
        val = ops.td_edc_abs_db_sum_fused(s_c, hy_c, hd[:n_ref], tdb[:n_ref], None)
        val.backward()
        err = dict(loss=abs(float(val) - float(ref)) / abs(float(ref)),
                   gs=float((s_c.grad - s_o.grad).abs().max() / s_o.grad.abs().max()),
                   ghy=float((hy_c.grad - hy_o.grad).abs().max() / hy_o.grad.abs().max()))
        ws = ops.td_fused_workspace(g, rows, tn, dev)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        gs = torch.empty(rows, g, device=dev)
        ghy = torch.empty(g, tn, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

        def launch():
            _lib.call("dgfdn_td_edc_fused", g, rows, tn, p(s), p(hy), p(hd), tn, p(tdb), tn, None, ctypes.c_double(1e-6),
                      p(loss), p(gs), p(ghy), 0, p(ws), st)

        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[v] = dict(info=info, ms=ms, gbps=8.0 * rows * tn / ms / 1e6, err=err)
        print(v, json.dumps(out[v]), flush=True)
    os.environ.pop("DGFDN_TD_VARIANT", None)


if __name__ == "__main__":
    main()
