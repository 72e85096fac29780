"""Worker of tests/test_gpu_dataparallel.py: one rank of a receiver-sharded ShardedEDCStep (not collected by pytest).

    python -m torch.distributed.run --nproc-per-node W tests/dp_worker.py OUT_DIR ROWS NFFT [strong]

Every rank builds the SAME net and the SAME full synthetic data set (seeded), keeps its contiguous block of receivers,
runs one step with world_size = W and writes loss terms + the flat gradient to OUT_DIR/rank{r}.pt. Ranks use NCCL when
the box has a GPU per rank, else they share cuda:0 and all-reduce over gloo (same host logic, same kernels)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_problem(rows, nfft, device, n_lines=12, t60=(0.03, 0.05, 0.06), seed=0):
    from diffgfdn_b200.config import DiffGFDNConfig, FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    from diffgfdn_b200.utils import unit_circle_grid
    torch.manual_seed(seed)
    delays = DiffGFDNConfig(seed=235265, num_delay_lines=n_lines).delay_length_samps
    net = DiffGFDNVarReceiverPos(32000.0, 3, delays, device, FeedbackLoopConfig(use_zero_coupling=False),
                                 OutputFilterConfig(use_svfs=False, num_hidden_layers=1, num_neurons_per_layer=16,
                                                    num_fourier_features=4), use_absorption_filters=False,
                                 common_decay_times=np.array([list(t60)]), use_colorless_loss=True)
    rng = np.random.default_rng(seed)
    rir = rng.standard_normal((rows, nfft // 2)) * np.exp(-np.arange(nfft // 2)[None, :] / 500.0)
    target = torch.tensor(np.fft.rfft(rir, n=nfft, axis=-1)).to(torch.complex64).to(device)
    early = torch.tensor(np.fft.rfft(rir[:, :640], n=nfft, axis=-1)).to(torch.complex64).to(device)
    pos = torch.tensor(rng.uniform(0, 1, (rows, 3))).to(device)
    return net, max(t60) * 1e3, unit_circle_grid(nfft).to(device), pos, early, target


def run_step(net, max_ms, z, pos, early, target, world, total, pg=None, graph=False, **kw):
    from diffgfdn_b200.fused import ShardedEDCStep
    step = ShardedEDCStep(net, max_ms, edc_weight=10.0, world_size=world, total_receivers=total, process_group=pg, **kw)
    step.attach(z, pos, None, None)
    step.attach(z, pos, step.precompute_early_window(early), step.precompute_target_db(target))
    if graph:
        state = {k: v.clone() for k, v in net.state_dict().items()}
        step.capture(optimizer=None, warmup=1)
        net.load_state_dict(state)
        out = step.replay()
    else:
        out = step.step()
    torch.cuda.synchronize()
    flat = torch.cat([p.grad.reshape(-1).float() for p in net.parameters()])
    run_step.peer_active = step.peer is not None  # exchanges over NVLink peer memory (else NCCL / gloo collectives)
    return {k: float(v) for k, v in out.items()}, flat.cpu()


def main():
    out_dir, rows, nfft = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    mode = sys.argv[4] if len(sys.argv) > 4 else "weak"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    one_gpu_each = torch.cuda.device_count() >= world
    dev = torch.device("cuda", local if one_gpu_each else 0)
    torch.cuda.set_device(dev)
    if one_gpu_each:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    net, max_ms, z, pos, early, target = make_problem(rows, nfft, dev)
    per = (rows + world - 1) // world
    sl = slice(rank * per, min(rows, (rank + 1) * per))
    kw = dict(shard_bins=True) if mode in ("strong", "strong_graph") else {}
    losses, flat = run_step(net, max_ms, z, pos[sl], early[sl], target[sl], world, rows, graph=mode.endswith("graph"), **kw)
    torch.save(dict(losses=losses, flat=flat, backend=dist.get_backend(), peer=run_step.peer_active),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
