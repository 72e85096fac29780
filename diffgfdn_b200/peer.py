"""Exchange steps of the bin-sharded multi-GPU step over NVLink peer memory (csrc/peer.cu).

One symmetric buffer per rank (torch's symmetric-memory rendezvous hands every process the P2P device pointers of all
of them), one channel per exchange of the step; an exchange is a push kernel (posted stores into every rank's buffer +
flags) and a wait-and-copy / wait-and-add kernel -- no collective library call, capturable in the step's CUDA graph.
The reference is single-process: this has no counterpart there."""
import ctypes
from typing import Dict

import torch
import torch.distributed as dist

from . import _lib


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class PeerExchange:
    """channels: name -> bytes a rank contributes per exchange (fixed for the life of the object)."""

    def __init__(self, group, device: torch.device, channels: Dict[str, int]):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerExchange: one node, at most 8 ranks")
        lib = _lib.load()
        self.flags_off = 0
        off = _align(int(lib.dgfdn_peer_flags_bytes()), 4096)
        self.layout = {}
        for i, (name, nbytes) in enumerate(channels.items()):
            if nbytes <= 0 or nbytes % 16:
                raise RuntimeError(f"PeerExchange: channel {name}: {nbytes} bytes is not a positive multiple of 16")
            stride = _align(nbytes, 256)
            region = stride * self.world
            self.layout[name] = dict(channel=i, nbytes=nbytes, data_off=off, region=region, stride=stride)
            off += 2 * region
        self.buf = symm.empty(off, dtype=torch.uint8, device=device)
        self.buf.zero_()
        hdl = symm.rendezvous(self.buf, self.group)
        ptrs = list(hdl.buffer_ptrs)
        if len(ptrs) != self.world or hdl.rank != self.rank:
            raise RuntimeError("PeerExchange: symmetric-memory rendezvous does not match the process group")
        self._hdl = hdl
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self.state = torch.zeros(int(lib.dgfdn_peer_state_bytes()) // 4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)  # every buffer is zeroed before anybody pushes

    def _call(self, fn: str, name: str, t: torch.Tensor):
        lay = self.layout[name]
        if t.numel() * t.element_size() != lay["nbytes"] or not t.is_contiguous() or t.device != self.buf.device:
            raise RuntimeError(f"PeerExchange.{fn}({name}): expected a contiguous tensor of {lay['nbytes']} bytes on "
                               f"{self.buf.device}")
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)
        return lay, stream

    def all_gather(self, name: str, mine: torch.Tensor) -> torch.Tensor:
        """(world * mine.shape[0], ...) <- every rank's `mine`, in rank order."""
        lay, stream = self._call("all_gather", name, mine)
        out = torch.empty((self.world * mine.shape[0], ) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        args = (self.world, self.rank, lay["channel"], self.ptrs, self.flags_off, lay["data_off"], lay["region"], lay["stride"],
                ctypes.c_void_p(self.state.data_ptr()))
        with torch.cuda.device(mine.device):
            _lib.call("dgfdn_peer_push", *args, ctypes.c_void_p(mine.data_ptr()), lay["nbytes"], stream)
            _lib.call("dgfdn_peer_gather", *args, ctypes.c_void_p(out.data_ptr()), lay["nbytes"], stream)
        return out

    def all_reduce_(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """t <- sum over the ranks (float32), added in rank order: bit-identical on every rank."""
        if t.dtype != torch.float32:
            raise RuntimeError("PeerExchange.all_reduce_: float32 only")
        lay, stream = self._call("all_reduce_", name, t)
        args = (self.world, self.rank, lay["channel"], self.ptrs, self.flags_off, lay["data_off"], lay["region"], lay["stride"],
                ctypes.c_void_p(self.state.data_ptr()))
        with torch.cuda.device(t.device):
            _lib.call("dgfdn_peer_push", *args, ctypes.c_void_p(t.data_ptr()), lay["nbytes"], stream)
            _lib.call("dgfdn_peer_reduce", *args, ctypes.c_void_p(t.data_ptr()), lay["nbytes"], stream)
        return t
