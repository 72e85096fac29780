"""the N-rank receiver-sharded step gives the same loss and the same (all-reduced) gradients
as one rank on the concatenated shards. Runs 2 ranks under torch.distributed.run: NCCL when the box has 2 GPUs, else
both ranks on cuda:0 with the all-reduce over gloo (the sharding logic and every kernel are the same)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(tmp_path, rows, nfft, mode, world=2):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py"), str(tmp_path), str(rows),
           str(nfft), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]


@pytest.mark.parametrize("peer", ["1", "0"])
@pytest.mark.parametrize("mode", ["weak", "strong", "graph", "strong_graph"])
def test_two_ranks_equal_one_rank_on_the_concatenation(tmp_path, mode, peer, monkeypatch):
    """mode weak: every rank solves all bins (K1 replicated); strong: K1 sharded over bins with all-gather of y /
    reduce of the adjoint right-hand sides; graph / strong_graph: the step captured in a CUDA graph (exchanges inside).
    peer 1: the exchanges run over NVLink peer memory (csrc/peer.cu) when every rank has its own GPU, 0: NCCL."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dp_worker
    if mode.endswith("graph") and torch.cuda.device_count() < 2:
        pytest.skip("a captured exchange needs one GPU per rank")
    if peer == "1" and torch.cuda.device_count() < 2:
        pytest.skip("peer memory needs one GPU per rank")
    monkeypatch.setenv("DGFDN_PEER", peer)
    rows, nfft = 10, 4096
    net, max_ms, z, pos, early, target = dp_worker.make_problem(rows, nfft, torch.device("cuda"))
    losses1, flat1 = dp_worker.run_step(net, max_ms, z, pos, early, target, 1, rows)
    ranks = _launch(tmp_path, rows, nfft, mode)
    edc = sum(r["losses"]["edc_loss"] for r in ranks)  # per-rank partial of the global mean
    assert abs(edc - losses1["edc_loss"]) < 1e-5 * abs(losses1["edc_loss"])
    if torch.cuda.device_count() >= 2:
        assert all(r["peer"] == (peer == "1") for r in ranks), [r["peer"] for r in ranks]
    if mode.startswith("strong"):  # bin-sharded colorless branch: the ranks' shares add up
        spec = sum(r["losses"]["spectral_loss"] for r in ranks)
        assert abs(spec - losses1["spectral_loss"]) < 1e-5 * abs(losses1["spectral_loss"])
    for r in ranks:
        if not mode.startswith("strong"):
            assert abs(r["losses"]["spectral_loss"] - losses1["spectral_loss"]) < 1e-5 * abs(losses1["spectral_loss"])
        err = float((r["flat"] - flat1).abs().max() / flat1.abs().max())
        assert err < 1e-4, (mode, r["backend"], err)
    assert torch.equal(ranks[0]["flat"], ranks[1]["flat"])
