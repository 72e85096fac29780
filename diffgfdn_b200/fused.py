"""Receiver-sharded training step for large receiver counts (BASELINE config 4: 100k receivers x 2^17 bins).

The reference can only hold ~32 receivers per step because it materialises (B, N, K) complex tensors
(model.py:583-619) and runs one irfft per receiver (losses.py:207-213). Two linearities remove all per-receiver
frequency-domain work:

  * b is shared by every receiver, so the per-bin system is solved ONCE per bin:  y[k,g] = c_g^T (D G^-1 - A)^-1 b;
  * irfft is linear and receiver r enters H_r = sum_g s[r,g] y_g + d_r only through its G real gains, so
        h_r = irfft(H_r)[window] = sum_g s[r,g] hy_g + hd_r,   hy_g = irfft(y_g)[window]   (G rows per step)
    and hd_r = irfft(d_r)[window] is a constant of the data set, precomputed once like the target EDC.

One step:

    s  = MLP(positions)                        torch (cuBLAS), autograd graph kept            (B, G)
    y  = solve(z, A, gamma, b, c)              K1 kernel, autograd graph kept                 (K, G)
    hy = irfft(y^T, n=K)[mix:max_len]          K3a chirp-z, G rows, autograd graph kept       (G, tn)
    for each tile of receivers:                raw C-ABI calls on preallocated buffers, no graph
        row loss, dL/ds, dL/dh = td_edc_step(s, hy, hd, target_db)      K3c: mix + EDC + dB loss + backward
        dL/dhy += s^T dL/dh                                             K3c contraction
    backward([hy, s], [dL/dhy, dL/ds]) + colorless losses -> parameter gradients   K3a^T + K1^T + torch autograd
    all-reduce of the flat gradient bucket over NCCL when world_size > 1

HBM traffic per receiver and time sample: hd 4 B + target dB 4 B (+ 4 B of dL/dh that stay in L2 when the tile is
small). Losses are means over ALL receivers of the job, so every rank scales by 1/B_total; the colorless losses are
receiver independent, computed on every rank and divided by world_size before the SUM all-reduce. Receivers are
independent given the shared parameters: sharding needs no data-path collective (weak scaling).

End-to-end mode (`step(host_d=..., host_target=...)`): the early and target responses arrive every step in the
reference's layout -- (B, K) complex64 frequency-domain arrays in pinned host memory (dataloader.py:674-704) -- and
are moved host -> device tile by tile on a copy stream (only bins 0..K/2, the ones irfft(X, n=K) reads), transformed
with the chirp-z kernels and fed to the same time-domain kernels."""
import ctypes
import os
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .losses import edc_loss
from .model import DiffGFDNVarReceiverPos

C64 = torch.complex64


class _GatherBins(torch.autograd.Function):
    """y (K, G) from this rank's bin slice y_loc = y[lo:hi] (slices of `per` bins, the last ones shorter): over NVLink
    peer memory (PeerExchange) when the ranks share a node, else an all-gather of the padded slices over NCCL, else (other
    backends) a zero-padded SUM all-reduce. The caller sums the gradient of everything downstream (dL/dhy) over the ranks
    BEFORE the backward, so the backward is the slice of the total gradient."""

    @staticmethod
    def forward(ctx, y_loc, lo, hi, k, per, group, peer):
        world = dist.get_world_size(group)
        ctx.lo, ctx.hi = lo, hi
        if peer is not None or dist.get_backend(group) == "nccl":
            mine = y_loc
            if hi - lo < per:
                mine = torch.zeros(per, y_loc.shape[1], dtype=y_loc.dtype, device=y_loc.device)
                mine[:hi - lo] = y_loc
            mine = mine.contiguous()
            if peer is not None:
                return peer.all_gather("y", mine)[:k]
            buf = torch.empty(world * per, y_loc.shape[1], dtype=y_loc.dtype, device=y_loc.device)
            dist.all_gather_into_tensor(torch.view_as_real(buf), torch.view_as_real(mine), group=group)
            return buf[:k]
        y = torch.zeros(k, y_loc.shape[1], dtype=y_loc.dtype, device=y_loc.device)
        y[lo:hi] = y_loc
        dist.all_reduce(torch.view_as_real(y), op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, gy):
        return gy[ctx.lo:ctx.hi].contiguous(), None, None, None, None, None, None


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class ShardedEDCStep:
    """One data-parallel training step (EDC + colorless losses) over this rank's receiver shard."""

    def __init__(self, net: DiffGFDNVarReceiverPos, max_ir_len_ms: float, tile_rows: int = 296,
                 edc_weight: float = 1.0, spectral_weight: float = 1.0, sparsity_weight: float = 1.0,
                 asym_spectral: bool = True, mixing_time_ms: float = 20.0, world_size: int = 1,
                 total_receivers: Optional[int] = None, process_group=None, e2e_tile_rows: int = 128,
                 shard_bins: bool = False, subband_filter: Optional[torch.Tensor] = None):
        self.net = net
        self.dev = net.device
        self.crit = edc_loss(max_ir_len_ms, net.sample_rate, mixing_time_ms=mixing_time_ms)
        self.tile_rows = int(tile_rows)
        self.e2e_tile_rows = int(e2e_tile_rows)
        self.w_edc, self.w_spec, self.w_spars = edc_weight, spectral_weight, sparsity_weight
        self.asym = asym_spectral
        self.world_size = world_size
        self.total_receivers = total_receivers
        self.pg = process_group
        # shard_bins: the receiver-independent per-bin work (coupled solve K1 + its adjoint, colorless branch K1c) is
        # split over the ranks by bins instead of repeated on every rank: y is gathered before the inverse DFT and
        # dL/dhy summed before the adjoint -- two small all-reduces (8 K G and 4 G tn bytes) on the critical path
        self.shard_bins = bool(shard_bins) and world_size > 1
        self.rank = dist.get_rank(process_group) if self.shard_bins else 0
        # Sub-band training (reference trainer.py:457-461, 804: H * subband_filter_freq_resp before the losses): the band
        # filter F[k] is receiver independent, so it folds into y[k,g] before the inverse DFT of the G rows and into the
        # early responses d when their windows are (re)built -- nothing per receiver and bin is added.
        self.subband_filter = None if subband_filter is None else subband_filter.to(net.device, C64).reshape(-1)
        self.kernel_launches = 0
        self.h2d_bytes = 0
        self._bufs = None
        self.mask = None
        self._mask_count = None
        self.events = None  # set to a dict of lists to collect per-kernel CUDA events (bench.py)
        self.use_side_stream = os.environ.get("DGFDN_SIDE_STREAM", "1") != "0"
        # exchanges of the multi-GPU step over NVLink peer memory (diffgfdn_b200/peer.py) instead of NCCL; set up at attach()
        self.use_peer = world_size > 1 and os.environ.get("DGFDN_PEER", "1") != "0"
        self.peer = None
        self.use_fused_colorless = os.environ.get("DGFDN_FUSED_COLORLESS", "1") != "0"
        self._side = None
        self._side2 = None

    def _side_stream(self) -> torch.cuda.Stream:
        if self._side is None:
            # between the main chain (captured at the highest priority) and the colorless branch: the position network's
            # backward has a big grid that would otherwise sit in front of the first kernels of the adjoint chain
            self._side = torch.cuda.Stream(device=self.dev, priority=-1)
        return self._side

    def _colorless_sms(self, bins: int) -> int:
        """Grid bound of K1c: the SMs the receiver kernel leaves idle when the branch can finish in that kernel's shadow
        (estimates from the measured rates: K1c 0.27 ms on 148 SMs for 131 073 bins x 3 groups of 8 lines; receiver
        kernel 3.4 TB/s of its 8 B per receiver.sample), else 0 = the whole device."""
        if not self._shadow_sms:
            return 0
        net = self.net
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        work = bins * net.num_groups * net.num_delay_lines_per_group**3 / (131073 * 3 * 512)
        t_k1c = 0.27e-3 * work * sms / self._shadow_sms
        t_rx = self.rows * self.tn * 8 / 3.4e12
        return self._shadow_sms if t_k1c <= 1.15 * t_rx else 0

    def _colorless_stream(self) -> torch.cuda.Stream:
        """The colorless branch has its own stream: on the position network's stream its 1.4 ms in the shadow of the
        receiver kernel would hold back that network's backward, which only waits for dL/ds."""
        if self._side2 is None:
            self._side2 = torch.cuda.Stream(device=self.dev, priority=0)  # lowest: it only fills what the main chain leaves
        return self._side2

    # ---- data ------------------------------------------------------------------------------------------
    def attach(self, z: torch.Tensor, positions: torch.Tensor, early_window: Optional[torch.Tensor],
               target_db: Optional[torch.Tensor]):
        """Device-resident shard: z (K,) c128, positions (B,3), early_window (B,tn) f32 or None (= precompute_
        early_window(d)), target_db (B,tn) f32 (= precompute_target_db(target))."""
        self.z = z.to(self.dev, torch.complex128)
        self.positions = positions.to(self.dev)
        self.k = self.z.numel()
        self.n_fft, self.t0, self.tn = self.crit.window(self.k)
        self.kx = self.n_fft // 2 + 1
        self.plan = ops.get_czt_plan(self.n_fft, self.t0, self.tn, self.dev)
        self.rows = self.positions.shape[0]
        if self.total_receivers is None:
            self.total_receivers = self.rows * self.world_size
        for name, t in (("early_window", early_window), ("target_db", target_db)):
            if t is not None and (tuple(t.shape) != (self.rows, self.tn) or t.dtype != torch.float32 or not t.is_cuda):
                raise RuntimeError(f"attach: {name} must be a CUDA float32 tensor of shape ({self.rows}, {self.tn})")
        self.hd = None if early_window is None else early_window.contiguous()
        self.target_db = None if target_db is None else target_db.contiguous()
        r = max(1, min(self.tile_rows, self.rows))
        g = self.net.num_groups
        dev = self.dev
        # K3d (cluster-fused kernel, no dL/dh in memory) when the shape allows it; DGFDN_TD_FUSED=0 forces K3c
        self.use_fused = ops.td_fused_supported(g, self.tn) and os.environ.get("DGFDN_TD_FUSED", "1") != "0"
        # SMs the receiver kernel's clusters cannot use (15 clusters of 8 CTAs on 148 SMs leave 28): the colorless branch
        # runs there with a grid bounded to them, so none of its blocks is left waiting for the SMs that kernel frees
        self._shadow_sms = 0
        if self.use_fused:
            info = ops.td_fused_info(g, self.tn)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            idle = sms - info["clusters"] * info["cluster_size"]
            if info["cluster_size"] > 1 and idle >= 16:
                self._shadow_sms = idle
            self._bufs = dict(fws=ops.td_fused_workspace(g, self.rows, self.tn, dev),
                              loss_sum=torch.zeros(1, dtype=torch.float64, device=dev))
        else:
            self._bufs = dict(gh=torch.empty(r, self.tn, dtype=torch.float32, device=dev),
                              ws=ops.td_contract_workspace(g, r, self.tn, dev),
                              row_sum=torch.empty(self.rows, dtype=torch.float64, device=dev))
        self._setup_peer()

    def set_mask(self, mask: Optional[torch.Tensor]):
        """Fixed 0/1 sample mask of the EDC loss (reference losses.py:221-238 draws one per call; here it is an
        input). The masked loss is a mean over the KEPT samples, so the normalisation uses mask.sum()."""
        if mask is None:
            self.mask, self._mask_count = None, None
            return
        m = mask.to(self.dev, torch.float32).contiguous()
        if m.numel() != self.tn:
            raise RuntimeError(f"set_mask: mask must have tn = {self.tn} samples")
        self.mask, self._mask_count = m, float(m.sum().item())

    def _window_rows(self, resp: torch.Tensor, to_db: bool) -> torch.Tensor:
        out = torch.empty(resp.shape[0], self.tn, dtype=torch.float32, device=self.dev)
        step = self.plan.rows_per_call(1 << 29)
        for r0 in range(0, resp.shape[0], step):
            h = ops.irfft_window(resp[r0:r0 + step].to(self.dev, C64), self.n_fft, self.t0, self.tn)
            out[r0:r0 + step] = ops.edc_db(h) if to_db else h
        return out

    @torch.no_grad()
    def precompute_target_db(self, target_response: torch.Tensor) -> torch.Tensor:
        """EDC of the targets in dB on the loss window (done once per data set; targets never change)."""
        return self._window_rows(target_response, True)

    @torch.no_grad()
    def precompute_early_window(self, early_response: torch.Tensor) -> torch.Tensor:
        """hd = irfft(d, n=K)[mix:max_len] of the early (direct-path) responses, (B, tn) float32. d is an input of
        the data set ('target_early_response'), constant over training, so its window is too. With a sub-band filter
        the model response is (sum_g s y_g + d) F, so d is filtered here."""
        if self.subband_filter is not None:
            f = self.subband_filter
            early_response = torch.cat([early_response[r0:r0 + 1024].to(self.dev, C64) * f[:early_response.shape[1]]
                                        for r0 in range(0, early_response.shape[0], 1024)])
        return self._window_rows(early_response, False)

    # ---- one step ----------------------------------------------------------------------------------------
    def step(self, host_d: Optional[torch.Tensor] = None, host_target: Optional[torch.Tensor] = None) -> Dict:
        """Forward + backward over the shard; leaves gradients in net.parameters().grad and returns the losses.

        With host_d / host_target (pinned host tensors, (P, K) complex64 pools that are cycled over the shard) the
        step streams its inputs host -> device tile by tile on a copy stream and rebuilds the early window and the
        target EDC on the fly: this is the end-to-end mode."""
        net = self.net
        g = net.num_groups
        for p in net.parameters():
            p.grad = None
        ev = self.events
        sec = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if ev is not None else None
        if sec:
            sec[0].record()
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.use_side_stream else main
        side2 = self._colorless_stream() if self.use_side_stream else main
        start = torch.cuda.Event()
        start.record(main)
        # irfft(X, n=K) reads bins 0..K/2 only (reference losses.py:207-213, quirk Q3), so the coupled system is
        # solved on those kx bins; the other bins of H reach no loss term (the colorless loss has its own solve)
        z_edc = self.z if net.feedback_loop.delay_line_gain_response is not None else self.z[:self.kx]
        if self.shard_bins:  # this rank's bins of the coupled solve; y is completed over the ranks
            ke = z_edc.shape[0]
            lo, hi = self._bin_slice(ke)
            _, y_loc = net.feedback_loop.solve(z_edc[lo:hi], net.input_gains.reshape(-1), net.output_gains.reshape(-1))
            y = _GatherBins.apply(y_loc, lo, hi, ke, self._bins_per_rank(ke), self.pg, self.peer)
            self.kernel_launches += 2 if self.peer is not None else 0  # push + wait-and-gather
        else:
            _, y = net.feedback_loop.solve(z_edc, net.input_gains.reshape(-1), net.output_gains.reshape(-1))
        # the graph is cut at y: the backward below runs the chirp-z adjoint first, on its own (see there)
        y_cut = y.detach().requires_grad_(True)
        y_f = y_cut if self.subband_filter is None else y_cut * self.subband_filter[:y_cut.shape[0]].unsqueeze(-1)
        hy = ops.irfft_window(y_f.transpose(0, 1), self.n_fft, self.t0, self.tn)  # (G, tn)
        # own kernels of the front: position network, 2 x skew-expm (coupled matrix, sparsity term), matrix assembly,
        # coupled solve, chirp-z (pre, mul, post; the 2 cuFFT launches are not counted), colorless branch
        fused_cl = self.use_fused_colorless and net.num_delay_lines_per_group <= 16
        self.kernel_launches += 1 + 2 + 1 + 1 + 3 + (3 if fused_cl else 2)  # K1c + its two reductions | groups solve + loss

        # Side stream, enqueued AFTER the coupled solve so that the solve chain leads the step: the position -> gain
        # network only meets that chain at the receiver kernel (and, autograd keeping a node's backward on its
        # forward's stream, its backward runs next to the adjoint solve); the sparsity term is a handful of tiny
        # kernels on one mixing matrix.
        side.wait_event(start)
        with torch.cuda.stream(side):
            s = net.output_scalars.gains({'norm_listener_position': self.positions})
            s_ready = torch.cuda.Event()
            s_ready.record(side)
        main.wait_event(s_ready)
        s_d = s.detach().contiguous()
        hy_d = hy.detach().contiguous()
        ghy = torch.empty_like(hy_d)
        gs = torch.empty_like(s_d)
        if self.mask is not None and self._mask_count is None:
            raise RuntimeError("step: attach the EDC mask with set_mask() (the loss is normalised by mask.sum())")
        coef = self.w_edc / (self.total_receivers * (self.tn if self.mask is None else self._mask_count))
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        pre_rx = torch.cuda.Event()
        pre_rx.record(main)
        if sec:
            sec[1].record()
        if host_d is None:
            if self.target_db is None:
                raise RuntimeError("step: attach() a target_db (and early_window) first, or pass host buffers")
            b = self._bufs
            if self.use_fused:  # the whole shard in one launch: nothing per receiver is written, so no tiling
                self._td_tile(0, self.rows, s_d, hy_d, self.hd, self.target_db, None, None, ghy, gs, coef, stream, False)
            else:
                r = b["gh"].shape[0]
                for i, r0 in enumerate(range(0, self.rows, r)):
                    r1 = min(self.rows, r0 + r)
                    self._td_tile(r0, r1, s_d, hy_d, None if self.hd is None else self.hd[r0:r1],
                                  self.target_db[r0:r1], b["gh"], b["ws"], ghy, gs, coef, stream, i > 0)
        else:
            self._stream_tiles(host_d, host_target, s_d, hy_d, ghy, gs, coef, stream)
        if sec:
            sec[2].record()
        # The colorless branch is receiver independent and shares nothing with the EDC branch but the parameters. It
        # is enqueued on the side stream AFTER the receiver kernel, gated only on what precedes that kernel: K3d's
        # 15 clusters of 8 CTAs take their 120 SMs first and K1c runs on the 28 SMs they cannot use (0.29 ms of work
        # for the whole chip = 1.5 ms on 28 SMs, the length of K3d), instead of competing with the coupled solve.
        side2.wait_event(pre_rx)
        with torch.cuda.stream(side2):
            sparsity = self.w_spars * self._sparsity(net.feedback_loop.ortho_param(net.feedback_loop.M[g - 1]))
            zc, share = self.z, 1.0 / self.world_size
            if self.shard_bins:  # mean over ALL bins = sum over ranks of (bins of the rank / K) x mean over its bins
                lo_c, hi_c = self._bin_slice(self.k)
                zc, share = self.z[lo_c:hi_c], (hi_c - lo_c) / self.k
            if self.use_fused_colorless and net.num_delay_lines_per_group <= 16:
                # K1c: solve, loss, dL/dy and the adjoint in one pass per bin (no H_sub, no second elimination)
                per_group = ops.colorless_solve_loss(zc, net.delays.to(torch.int32), net.feedback_loop.M,
                                                     net.input_gains.reshape(-1), net.output_gains.reshape(-1), self.asym,
                                                     max_sms=self._colorless_sms(zc.numel()) if host_d is None else 0)
            else:
                keep = net.return_per_delay_outputs
                net.return_per_delay_outputs = False
                h_sub, _ = net.sub_fdn_output(zc)
                net.return_per_delay_outputs = keep
                per_group = ops.colorless_loss_per_group(h_sub, self.asym)
            spectral = self.w_spec * per_group.sum() * share  # this rank's share: the ranks' values add up to the loss
            aux = spectral + sparsity.to(spectral.dtype) / self.world_size
        gs_ready = torch.cuda.Event()  # dL/ds (and dL/dhy of this rank) are complete on the main stream
        gs_ready.record(main)
        edc = (self._bufs["loss_sum"][0] if self.use_fused else self._bufs["row_sum"].sum()) * coef
        if self.shard_bins:  # every rank needs the TOTAL dL/dhy for the adjoint solve of its bins
            if self.peer is not None:
                self.peer.all_reduce_("ghy", ghy)
                self.kernel_launches += 2  # push + wait-and-add
            else:
                dist.all_reduce(ghy, op=dist.ReduceOp.SUM, group=self.pg)
        # no join before the backward: the engine runs each node on its forward's stream and orders producers and
        # consumers itself, so the adjoint solve of the EDC branch does not wait for the tail of the colorless branch
        # two calls: the EDC branch first, so that its nodes are enqueued (and captured) ahead of the colorless tail --
        # the branches share nothing but leaf parameters
        # three calls for the EDC branch: (1) the chirp-z adjoint (a chain of 5-15 us kernels right behind the receiver
        # kernel), (2) the position network's backward on its stream, gated on (1): its big grid otherwise shares the SMs
        # with those latency-bound kernels and doubles their time; behind them it overlaps the adjoint solve (FP64 /
        # shuffle bound against FP32 FMA work), (3) the adjoint solve chain from dL/dy on.
        # (With the per-bin work sharded over the ranks the adjoint chain is short and the position network's backward is
        # what finishes last: it then starts as soon as dL/ds exists. Measured at N = 8: 1.84 ms ungated, 1.87 ms gated.)
        torch.autograd.backward([hy], [ghy])
        if not self.shard_bins:
            after_czt = torch.cuda.Event()
            after_czt.record(main)
            side.wait_event(after_czt)
        else:
            side.wait_event(gs_ready)  # (the engine takes the gradients handed to backward() as ready on the CALLING stream)
        with torch.cuda.stream(side):
            torch.autograd.backward([s], [gs])
        torch.autograd.backward([y], [y_cut.grad])
        with torch.cuda.stream(side2):
            torch.autograd.backward([aux], [torch.ones_like(aux)])
        main.wait_stream(side)  # the engine joins the streams of the leaves; this makes the join explicit for capture
        main.wait_stream(side2)
        # chirp-z adjoint, coupled adjoint solve (+ reduce), assembly bwd, 2 x skew-expm bwd, position network bwd
        # (+ reduce); the separate colorless path adds its adjoint solve (+ reduce) and the loss backward
        self.kernel_launches += 3 + 2 + 1 + 2 + 2 + (0 if fused_cl else 3)
        if self.world_size > 1:
            self.allreduce_grads()
        if sec:
            sec[3].record()
            ev.setdefault("front(MLP+solves+irfft of G rows)", []).append((sec[0], sec[1]))
            ev.setdefault("receiver tiles", []).append((sec[1], sec[2]))
            ev.setdefault("back(irfft^T+adjoint solves+autograd)", []).append((sec[2], sec[3]))
        # losses as reported: 'edc_loss' and (with shard_bins) 'spectral_loss' are this rank's share -- they add up over the
        # ranks; without shard_bins every rank evaluates the whole spectral loss
        spectral_rep = spectral.detach() if self.shard_bins else spectral.detach() * self.world_size
        return {'edc_loss': edc, 'spectral_loss': spectral_rep, 'sparsity_loss': sparsity.detach()}

    # ---- CUDA graph: the whole resident step (+ optimizer) as one launch ----------------------------------
    def capture(self, optimizer: Optional[torch.optim.Optimizer] = None, warmup: int = 3):
        """Capture step() (+ optimizer.step()) of the resident mode into a CUDA graph. Every kernel of the step --
        cuBLAS MLP, expm, per-bin solves, cuFFT, the receiver kernels, the NCCL all-reduce -- is stream ordered and
        free of host synchronisation, so a replay costs one launch. The optimizer must be capturable (e.g.
        torch.optim.Adam(..., capturable=True)). Gradients live in the graph's memory pool: parameters keep
        pointing at them, so code that reads p.grad after replay() sees the fresh values."""
        if self.target_db is None:
            raise RuntimeError("capture: attach() resident inputs first")
        self.events = None
        # the captured main chain runs at high stream priority (kernel nodes keep the priority of the stream they were
        # captured on): when the receiver kernel frees its SMs, the adjoint chain is placed ahead of the waiting blocks of
        # the low-priority colorless branch
        side = torch.cuda.Stream(device=self.dev, priority=-3)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.step()
                if optimizer is not None:
                    optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        before = self.kernel_launches
        self._graph = torch.cuda.CUDAGraph()
        # same stream as the warm-up: the parameters' AccumulateGrad nodes are bound to the stream they were
        # created on, and a backward that crosses streams would invalidate the capture
        with torch.cuda.graph(self._graph, stream=side):
            self._static_losses = self.step()
            if optimizer is not None:
                optimizer.step()
        self.launches_per_replay = self.kernel_launches - before
        return self._graph

    def release_graph(self):
        """Drop the captured graph and its static outputs (before tearing down a process group it holds nodes of)."""
        self._graph = None
        self._static_losses = None

    def replay(self) -> Dict:
        """Run the captured step once; returns the (static) loss tensors, overwritten by every replay."""
        self._graph.replay()
        self.kernel_launches += self.launches_per_replay
        return self._static_losses

    @torch.no_grad()
    def _td_tile(self, r0, r1, s_d, hy_d, hd_tile, tdb_tile, gh, ws, ghy, gs, coef, stream, accumulate):
        g = self.net.num_groups
        rows, tn = r1 - r0, self.tn
        ev = self.events
        if ev is not None:
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            marks[0].record()
        if self.use_fused:
            _lib.call("dgfdn_td_edc_fused", g, rows, tn, _p(s_d[r0:r1]), _p(hy_d), _p(hd_tile), tn, _p(tdb_tile), tn,
                      _p(self.mask), ctypes.c_double(coef), _p(self._bufs["loss_sum"]), _p(gs[r0:r1]), _p(ghy),
                      1 if accumulate else 0, _p(self._bufs["fws"]), stream)
            if ev is not None:
                marks[1].record()
                ev.setdefault("td_edc_fused", []).append((marks[0], marks[1]))
            self.kernel_launches += 2  # td_fused, td_fused_finalize
            return
        _lib.call("dgfdn_td_edc_step", g, rows, tn, _p(s_d[r0:r1]), _p(hy_d), _p(hd_tile), tn, _p(tdb_tile), tn,
                  _p(self.mask), ctypes.c_double(coef), _p(self._bufs["row_sum"][r0:r1]), _p(gs[r0:r1]), _p(gh), tn,
                  stream)
        if ev is not None:
            marks[1].record()
        _lib.call("dgfdn_td_contract", g, rows, tn, _p(s_d[r0:r1]), _p(gh), tn, _p(ghy), 1 if accumulate else 0, _p(ws),
                  stream)
        if ev is not None:
            marks[2].record()
            ev.setdefault("td_edc_step", []).append((marks[0], marks[1]))
            ev.setdefault("td_contract", []).append((marks[1], marks[2]))
        self.kernel_launches += 3  # td_edc_step, td_contract, td_contract_reduce

    def _bins_per_rank(self, k: int) -> int:
        """Bins of a rank's slice: ceil(k / world) rounded up to an even count (16-byte granularity of complex64 x G rows
        for the peer-memory exchange); the last ranks' slices are shorter."""
        per = (k + self.world_size - 1) // self.world_size
        return per + (per & 1)

    def _bin_slice(self, k: int):
        """Contiguous bins [lo, hi) of this rank out of k."""
        per = self._bins_per_rank(k)
        lo = min(k, self.rank * per)
        return lo, min(k, lo + per)

    def _setup_peer(self):
        """One symmetric buffer per rank with a channel per exchange of the step (y slices, dL/dhy, gradient bucket)."""
        if not self.use_peer or self.peer is not None:
            return
        try:
            if dist.get_backend(self.pg) != "nccl":
                raise RuntimeError("ranks do not own one GPU each")
            from .peer import PeerExchange
            g = self.net.num_groups
            n_grad = sum(p.numel() for p in self.net.parameters() if p.requires_grad)
            channels = {"grads": 4 * ((n_grad + 3) // 4 * 4)}
            if self.shard_bins:
                ke = self.k if self.net.feedback_loop.delay_line_gain_response is not None else self.kx
                channels["y"] = self._bins_per_rank(ke) * g * 8
                channels["ghy"] = g * self.tn * 4
            self.peer = PeerExchange(self.pg, self.dev, channels)
        except Exception as exc:  # NCCL carries the exchanges then (same results, a library collective per exchange)
            import warnings
            warnings.warn(f"ShardedEDCStep: peer-memory exchange unavailable ({exc}); using NCCL collectives")
            self.use_peer = False

    @staticmethod
    def _sparsity(a: torch.Tensor) -> torch.Tensor:
        n = a.shape[-1]
        return -(torch.sum(torch.abs(a)) - n * n**0.5) / (n * (n**0.5 - 1))

    # ---- end-to-end mode: inputs come from pinned host memory every step ---------------------------------
    def _stream_tiles(self, host_d, host_target, s_d, hy_d, ghy, gs, coef, stream):
        if self.subband_filter is not None:
            raise RuntimeError("step(host_d, host_target): the host-streamed mode rebuilds the early windows from the raw "
                               "early responses; with a sub-band filter use the resident mode (precompute_early_window "
                               "applies the filter) or hand over early responses that are already band filtered")
        dev = self.dev
        r = max(1, min(self.e2e_tile_rows, self.rows))
        kx, tn, g = self.kx, self.tn, self.net.num_groups
        for name, t in (("host_d", host_d), ("host_target", host_target)):
            if t.dtype != C64 or t.dim() != 2 or t.shape[1] < kx or not t.is_pinned() or not t.is_contiguous():
                raise RuntimeError(f"step: {name} must be a contiguous pinned complex64 (P, >= {kx}) host tensor")
        if "stage" not in self._bufs:
            self._bufs["stage"] = [dict(d=torch.empty(r, kx, dtype=C64, device=dev),
                                        t=torch.empty(r, kx, dtype=C64, device=dev),
                                        ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
            self._bufs["copy_stream"] = torch.cuda.Stream(device=dev)
            self._bufs["e2e"] = dict(scratch=torch.empty(r * self.plan.mc, dtype=C64, device=dev),
                                     hd=torch.empty(r, tn, dtype=torch.float32, device=dev),
                                     ht=torch.empty(r, tn, dtype=torch.float32, device=dev),
                                     tdb=torch.empty(r, tn, dtype=torch.float32, device=dev),
                                     gh=None if self.use_fused else torch.empty(r, tn, dtype=torch.float32, device=dev),
                                     ws=None if self.use_fused else ops.td_contract_workspace(g, r, tn, dev))
        stage = self._bufs["stage"]
        e = self._bufs["e2e"]
        copy_stream = self._bufs["copy_stream"]
        cs = ctypes.c_void_p(copy_stream.cuda_stream)
        main = torch.cuda.current_stream()
        pool = host_d.shape[0]
        if host_target.shape[0] != pool or pool < min(r, self.rows):
            raise RuntimeError(f"step: host pools must have the same row count and hold at least one tile "
                               f"({min(r, self.rows)} rows); got {pool} and {host_target.shape[0]}")
        self.h2d_bytes = 0
        tiles = list(range(0, self.rows, r))

        def issue(i):
            r0 = tiles[i]
            n = min(self.rows, r0 + r) - r0
            st = stage[i % 2]
            p0 = r0 % pool
            if p0 + n > pool:
                p0 = 0
            copy_stream.wait_event(st["free"])
            for key, host in (("d", host_d), ("t", host_target)):
                _lib.call("dgfdn_copy_rows_h2d", _p(st[key]), kx * 8, ctypes.c_void_p(host[p0].data_ptr()),
                          host.shape[1] * 8, kx * 8, n, cs)
            st["ready"].record(copy_stream)
            self.h2d_bytes += 2 * n * kx * 8

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
            # early window and target EDC of this tile, rebuilt every step in this mode
            _lib.call("dgfdn_irfft_window_fwd", self.plan.handle, _p(st["d"]), kx, n, None, _p(e["scratch"]), _p(e["hd"]),
                      stream)
            _lib.call("dgfdn_irfft_window_fwd", self.plan.handle, _p(st["t"]), kx, n, None, _p(e["scratch"]), _p(e["ht"]),
                      stream)
            st["free"].record(main)
            _lib.call("dgfdn_edc_db", _p(e["ht"]), n, tn, _p(e["tdb"]), stream)
            self.kernel_launches += 7
            self._td_tile(r0, r1, s_d, hy_d, e["hd"], e["tdb"], e["gh"], e["ws"], ghy, gs, coef, stream, i > 0)

    # ---- data parallel -----------------------------------------------------------------------------------
    def allreduce_grads(self):
        """One NCCL all-reduce (SUM) over a single flat float32 bucket holding every parameter gradient
        (< 1 MB: latency bound over NVLink 5 / NVSwitch)."""
        params = [p for p in self.net.parameters() if p.requires_grad]
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        parts = [p.grad.reshape(-1).to(torch.float32) for p in params]
        n = sum(t.numel() for t in parts)
        peer = getattr(self, "peer", None)
        if peer is not None and (n + 3) // 4 * 4 > n:  # the peer exchange moves 16-byte units
            parts.append(torch.zeros((n + 3) // 4 * 4 - n, dtype=torch.float32, device=parts[0].device))
        flat = torch.cat(parts)
        if peer is not None:
            peer.all_reduce_("grads", flat)
            self.kernel_launches += 2
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
        views, off = [], 0
        for p in params:
            n = p.numel()
            views.append(flat[off:off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_([p.grad for p in params], views)  # one multi-tensor kernel, not one copy per parameter
