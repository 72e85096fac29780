"""World-size-2 gloo test (CPU) of the data-parallel host logic: the flat-bucket gradient all-reduce and the
loss normalisation convention of diffgfdn_b200/fused.py (sum over ranks of per-rank partial means == global mean)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffgfdn_b200.fused import ShardedEDCStep
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    x = torch.arange(32, dtype=torch.float32).reshape(8, 4) / 10.0
    shard = x[rank * 4:(rank + 1) * 4]
    # per-rank partial of a global mean: divide by the TOTAL number of rows, then SUM over ranks
    loss = net(shard).pow(2).sum() / x.shape[0]
    loss.backward()
    step = ShardedEDCStep.__new__(ShardedEDCStep)
    step.net, step.pg, step.world_size = net, None, world
    step.allreduce_grads()
    ret[rank] = [p.grad.clone() for p in net.parameters()]
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    x = torch.arange(32, dtype=torch.float32).reshape(8, 4) / 10.0
    (net(x).pow(2).sum() / x.shape[0]).backward()
    for r in range(world):
        for g, p in zip(ret[r], net.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6)
