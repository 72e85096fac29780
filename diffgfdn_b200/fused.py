"""Receiver-sharded, tiled training step for large receiver counts (BASELINE config 4: 100k receivers x 2^17 bins).

The reference can only hold ~32 receivers per step because it materialises (B, N, K) complex tensors
(model.py:583-619). Here the per-bin solve is receiver independent and the projection + loss pipeline streams over
receivers, so a step walks the rank's receiver shard in tiles:

    s = MLP(positions)                      torch (cuBLAS), autograd graph kept            (B, G)
    x, y = solve(z, A, gamma, b, c)         K1 kernel, autograd graph kept                 (K, G)
    for each tile of R receivers:           raw C-ABI calls on preallocated buffers, no graph
        H   = project(s, y, d)              K2          (R, K) -- full spectrum, as the reference's forward
        h   = irfft(H, n=K)[mix:max_len]    K3a         chirp-z over cuFFT
        l  += sum |EDC_dB(target) - EDC_dB(h)|          K3b forward
        gh  = dl/dh                         K3b backward
        gH  = irfft^T(gh)                   K3a adjoint (bins 0..K/2 only: the others have zero gradient, Q3)
        gy += s^T gH ; gs = Re(gH y^H)      K2 adjoint
    backward([y, s], [gy, gs]) + colorless losses -> parameter gradients     K1^T kernel + torch autograd
    all-reduce of the flat gradient bucket over NCCL when world_size > 1

Losses are means over ALL receivers of the job, so every rank scales by 1/B_total; the colorless losses are
receiver independent, computed on every rank and divided by world_size before the SUM all-reduce.
Receivers are independent given the shared parameters: sharding needs no data-path collective (weak scaling)."""
import ctypes
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .losses import edc_loss
from .model import DiffGFDNVarReceiverPos

C64 = torch.complex64


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class ShardedEDCStep:
    """One data-parallel training step (EDC + colorless losses) over this rank's receiver shard."""

    def __init__(self, net: DiffGFDNVarReceiverPos, max_ir_len_ms: float, tile_rows: int = 128,
                 edc_weight: float = 1.0, spectral_weight: float = 1.0, sparsity_weight: float = 1.0,
                 asym_spectral: bool = True, mixing_time_ms: float = 20.0, world_size: int = 1,
                 total_receivers: Optional[int] = None, process_group=None):
        self.net = net
        self.dev = net.device
        self.crit = edc_loss(max_ir_len_ms, net.sample_rate, mixing_time_ms=mixing_time_ms)
        self.tile_rows = int(tile_rows)
        self.w_edc, self.w_spec, self.w_spars = edc_weight, spectral_weight, sparsity_weight
        self.asym = asym_spectral
        self.world_size = world_size
        self.total_receivers = total_receivers
        self.pg = process_group
        self.kernel_launches = 0
        self._bufs = None
        self._flat = None

    # ---- data ------------------------------------------------------------------------------------------
    def attach(self, z: torch.Tensor, positions: torch.Tensor, d: Optional[torch.Tensor],
               target_db: Optional[torch.Tensor]):
        """Device-resident shard: z (K,) c128, positions (B,3), d (B,K) c64 or None, target_db (B,tn) f32."""
        self.z = z.to(self.dev, torch.complex128)
        self.positions = positions.to(self.dev)
        self.d = d
        self.target_db = target_db
        self.k = self.z.numel()
        self.n_fft, self.t0, self.tn = self.crit.window(self.k)
        self.kx = self.n_fft // 2 + 1
        self.plan = ops.get_czt_plan(self.n_fft, self.t0, self.tn, self.dev)
        self.rows = self.positions.shape[0]
        if self.total_receivers is None:
            self.total_receivers = self.rows * self.world_size
        r = min(self.tile_rows, self.rows)
        dev = self.dev
        self._bufs = dict(h_tile=torch.empty(r, self.k, dtype=C64, device=dev),
                          scratch=torch.empty(r * self.plan.mc, dtype=C64, device=dev),
                          h=torch.empty(r, self.tn, dtype=torch.float32, device=dev),
                          gh=torch.empty(r, self.tn, dtype=torch.float32, device=dev),
                          g_tile=torch.empty(r, self.kx, dtype=C64, device=dev),
                          row_sum=torch.empty(self.rows, dtype=torch.float64, device=dev))

    @torch.no_grad()
    def precompute_target_db(self, target_response: torch.Tensor) -> torch.Tensor:
        """EDC of the targets in dB for the loss window (done once per dataset; targets never change)."""
        out = torch.empty(target_response.shape[0], self.tn, dtype=torch.float32, device=self.dev)
        step = self.plan.rows_per_call(1 << 29)
        for r0 in range(0, target_response.shape[0], step):
            h = ops.irfft_window(target_response[r0:r0 + step].to(self.dev, C64), self.n_fft, self.t0, self.tn)
            out[r0:r0 + step] = ops.edc_db(h)
        return out

    # ---- the tile pipeline -----------------------------------------------------------------------------
    @torch.no_grad()
    def _tile(self, r0, r1, s_d, y_d, d_tile, tdb_tile, gy_acc, gs, coef, stream, accumulate):
        b = self._bufs
        g = self.net.num_groups
        rows = r1 - r0
        k, kx, tn = self.k, self.kx, self.tn
        lib_call = _lib.call
        lib_call("dgfdn_project_fwd", g, rows, k, _p(s_d[r0:r1]), _p(y_d), _p(d_tile), k, _p(b["h_tile"]), k, stream)
        lib_call("dgfdn_irfft_window_fwd", self.plan.handle, _p(b["h_tile"]), k, rows, None, _p(b["scratch"]),
                 _p(b["h"]), stream)
        lib_call("dgfdn_edc_loss_fwd", _p(b["h"]), _p(tdb_tile), None, rows, tn, _p(b["row_sum"][r0:r1]), stream)
        lib_call("dgfdn_edc_loss_bwd", _p(b["h"]), _p(tdb_tile), None, rows, tn, ctypes.c_double(coef), _p(b["gh"]),
                 stream)
        lib_call("dgfdn_irfft_window_bwd", self.plan.handle, _p(b["gh"]), rows, None, _p(b["scratch"]),
                 _p(b["g_tile"]), kx, kx, stream)
        lib_call("dgfdn_project_bwd", g, rows, kx, _p(s_d[r0:r1]), _p(y_d), _p(b["g_tile"]), kx, _p(gy_acc),
                 1 if accumulate else 0, _p(gs[r0:r1]), stream)
        # own kernels per tile: project 1, czt fwd 3, edc fwd 1, edc bwd 1, czt bwd 3, project bwd 2 (+4 cuFFT)
        self.kernel_launches += 11

    def step(self, host_d: Optional[torch.Tensor] = None, host_target: Optional[torch.Tensor] = None) -> Dict:
        """Forward + backward over the shard; leaves gradients in net.parameters().grad and returns the losses.

        With host_d / host_target (pinned host tensors, (P, K) complex64 pools that are cycled over the shard) the
        step streams its inputs host -> device tile by tile on a copy stream and rebuilds the target EDC on the fly:
        this is the end-to-end mode."""
        net = self.net
        for p in net.parameters():
            p.grad = None
        s = net.output_scalars.gains({'norm_listener_position': self.positions})
        _, y = net.feedback_loop.solve(self.z, net.input_gains.reshape(-1), net.output_gains.reshape(-1))
        keep = net.return_per_delay_outputs
        net.return_per_delay_outputs = False
        h_sub, _ = net.sub_fdn_output(self.z)
        net.return_per_delay_outputs = keep
        per_group = ops.colorless_loss_per_group(h_sub, self.asym)
        spectral = self.w_spec * per_group.sum()
        sparsity = self.w_spars * self._sparsity(net.feedback_loop.ortho_param(net.feedback_loop.M[net.num_groups - 1]))
        self.kernel_launches += 2 + 1  # two solves, colorless forward

        s_d = s.detach().contiguous()
        y_d = y.detach().contiguous()
        gy_acc = torch.zeros_like(y_d)
        gs = torch.empty_like(s_d)
        coef = self.w_edc / (self.total_receivers * self.tn)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        r = self._bufs["h_tile"].shape[0]
        if host_d is None:
            for i, r0 in enumerate(range(0, self.rows, r)):
                r1 = min(self.rows, r0 + r)
                d_tile = None if self.d is None else self.d[r0:r1]
                self._tile(r0, r1, s_d, y_d, d_tile, self.target_db[r0:r1], gy_acc, gs, coef, stream, i > 0)
        else:
            self._stream_tiles(host_d, host_target, s_d, y_d, gy_acc, gs, coef, stream, r)
        edc = self._bufs["row_sum"].sum() * coef
        aux = (spectral + sparsity.to(spectral.dtype)) / self.world_size
        torch.autograd.backward([y, s, aux], [gy_acc, gs, torch.ones_like(aux)])
        self.kernel_launches += 2 * 2 + 1  # two adjoint solves (+ reduce each), colorless backward
        if self.world_size > 1:
            self.allreduce_grads()
        return {'edc_loss': edc, 'spectral_loss': spectral.detach(), 'sparsity_loss': sparsity.detach()}

    @staticmethod
    def _sparsity(a: torch.Tensor) -> torch.Tensor:
        n = a.shape[-1]
        return -(torch.sum(torch.abs(a)) - n * n**0.5) / (n * (n**0.5 - 1))

    # ---- end-to-end mode: inputs come from pinned host memory every step ---------------------------------
    def _stream_tiles(self, host_d, host_target, s_d, y_d, gy_acc, gs, coef, stream, r):
        dev = self.dev
        if "stage" not in self._bufs:
            self._bufs["stage"] = [dict(d=torch.empty(r, self.k, dtype=C64, device=dev),
                                        t=torch.empty(r, self.k, dtype=C64, device=dev),
                                        tdb=torch.empty(r, self.tn, dtype=torch.float32, device=dev),
                                        ht=torch.empty(r, self.tn, dtype=torch.float32, device=dev),
                                        ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
            self._bufs["copy_stream"] = torch.cuda.Stream(device=dev)
        stage = self._bufs["stage"]
        copy_stream = self._bufs["copy_stream"]
        main = torch.cuda.current_stream()
        pool = host_d.shape[0]
        self.h2d_bytes = 0
        tiles = list(range(0, self.rows, r))

        def issue(i):
            r0 = tiles[i]
            r1 = min(self.rows, r0 + r)
            st = stage[i % 2]
            p0 = r0 % pool
            n = r1 - r0
            if p0 + n > pool:
                p0 = 0
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(st["free"])
                st["d"][:n].copy_(host_d[p0:p0 + n], non_blocking=True)
                st["t"][:n].copy_(host_target[p0:p0 + n], non_blocking=True)
                st["ready"].record(copy_stream)
            self.h2d_bytes += 2 * n * self.k * 8

        for st in stage:
            st["free"].record(main)
        issue(0)
        for i, r0 in enumerate(tiles):
            r1 = min(self.rows, r0 + r)
            n = r1 - r0
            if i + 1 < len(tiles):
                issue(i + 1)
            st = stage[i % 2]
            main.wait_event(st["ready"])
            # target EDC of this tile, rebuilt every step in this mode
            _lib.call("dgfdn_irfft_window_fwd", self.plan.handle, _p(st["t"]), self.k, n, None,
                      _p(self._bufs["scratch"]), _p(st["ht"]), stream)
            _lib.call("dgfdn_edc_db", _p(st["ht"]), n, self.tn, _p(st["tdb"]), stream)
            self.kernel_launches += 4
            self._tile(r0, r1, s_d, y_d, st["d"][:n], st["tdb"][:n], gy_acc, gs, coef, stream, i > 0)
            st["free"].record(main)

    # ---- data parallel -----------------------------------------------------------------------------------
    def allreduce_grads(self):
        """One NCCL all-reduce (SUM) over a single flat float32 bucket holding every parameter gradient
        (< 1 MB: latency bound over NVLink 5 / NVSwitch)."""
        params = [p for p in self.net.parameters() if p.requires_grad]
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
