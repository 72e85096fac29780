"""Kernel timeline (start offset, duration, stream) of ONE replay of the captured training step (CUPTI via
torch.profiler): shows what overlaps with what.   python scripts/timeline_step.py > gpurun_out/timeline.txt"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


MIN_US = float(os.environ.get("TIMELINE_MIN_US", "4"))


def main():
    from diffgfdn_b200.fused import ShardedEDCStep
    from diffgfdn_b200.utils import unit_circle_grid
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))  # under torchrun: the bin-sharded multi-GPU step, rank 0 prints
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = bench.build_net(dev)
    rows = 12500
    step = ShardedEDCStep(net, max(bench.T60) * 1e3, tile_rows=rows, edc_weight=10.0, world_size=world,
                          total_receivers=rows * world, shard_bins=world > 1)
    z = unit_circle_grid(bench.NFFT, device=dev)
    pos = torch.rand(rows, 3, device=dev)
    step.attach(z, pos, None, None)
    early, tgt = bench.synth_responses(256, bench.NFFT, dev, 1)
    reps = (rows + 255) // 256
    hd = step.precompute_early_window(early).repeat(reps, 1)[:rows].contiguous()
    tdb = step.precompute_target_db(tgt).repeat(reps, 1)[:rows].contiguous()
    step.attach(z, pos, hd, tdb)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)
    step.capture(optimizer=opt, warmup=2)
    for _ in range(3):
        step.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.replay()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        if dist.get_rank() != 0:
            step.release_graph()
            dist.destroy_process_group()
            return
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in ev)
    print(f"# one replay: {len(ev)} device activities, {(end - t0) / 1e3:.3f} ms from first start to last end")
    for e in ev:
        if e["dur"] >= MIN_US:
            print(f"{(e['ts'] - t0) / 1e3:8.3f} ms  +{e['dur'] / 1e3:7.3f} ms  stream {e['args'].get('stream', '?'):>3}  {e['name'][:90]}")
    if world > 1:
        step.release_graph()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
