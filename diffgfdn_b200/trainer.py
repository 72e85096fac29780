"""Training-step contract of the reference (diff_gfdn/trainer.py) on the B200 kernels.

Kept from the reference: Adam parameter groups by parameter name and StepLR(10, 0.1) (trainer.py:152-228), the
per-step energy normalisation of b and c (trainer.py:317-332), the loss composition including its quirks
(spectral loss `+=` over groups, sparsity loss `=` last group only: trainer.py:295-313, Q2), sub-band filtering of
H before the losses (trainer.py:457-461, 804), the SH -> direction projection (trainer.py:853-865), per-epoch
`state_dict` checkpoints (trainer.py:249-257) and early stopping. Not ported: wav export (`save_ir`) and the pyfar
filterbank design -- the sub-band filter response is an input (`set_subband_filter`)."""
import os
from pathlib import Path
import time
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .colorless_fdn.losses import amse_loss, mse_loss, sparsity_loss
from .config.config import TrainerConfig
from .losses import directional_edc_loss, edc_loss, edr_loss
from .model import DiffGFDN
from .utils import TensorKeyedCache


class Trainer:

    def __init__(self, net: DiffGFDN, trainer_config: TrainerConfig):
        self.net = net
        self.device = net.device
        self.max_epochs = trainer_config.max_epochs
        self.patience = 5
        self.early_stop = 0
        self.train_dir = Path(trainer_config.train_dir).resolve()
        self.ir_dir = Path(trainer_config.ir_dir).resolve()
        self.use_reg_loss = trainer_config.use_reg_loss
        self.use_colorless_loss = trainer_config.use_colorless_loss
        self.reduced_pole_radius = trainer_config.reduced_pole_radius
        self.subband_process_config = trainer_config.subband_process_config
        self.subband_filter_freq_resp = None
        self.use_directional_fdn = getattr(self.net, "ambi_order", None) is not None
        if self.use_reg_loss:
            # the reference's own branch cannot run: its trainer calls torch.cat(tensor, tensor) at construction
            # (trainer.py:89-100) and the loss reads net.biquad_cascade, which VarReceiverPos models do not have (:290-293)
            raise NotImplementedError("use_reg_loss: the reference's reg_loss branch raises at trainer construction "
                                      "(torch.cat on a tensor, trainer.py:99); there is no behaviour to mirror")
        self.init_scheduler(trainer_config)
        if self.net.common_decay_times is None:
            max_ir_len_ms = 2000
        else:
            max_ir_len_ms = float(self.net.common_decay_times.max() * 1e3)
        if self.use_directional_fdn:
            self.criterion = [directional_edc_loss(self.net.common_decay_times, max_ir_len_ms, self.net.sample_rate,
                                                   use_mask=trainer_config.use_edc_mask).to(self.device)]
            self.loss_weights = [trainer_config.edc_loss_weight]
        else:
            self.criterion = [
                edr_loss(self.net.sample_rate, reduced_pole_radius=self.reduced_pole_radius,
                         use_erb_grouping=trainer_config.use_erb_edr_loss,
                         use_weight_fn=trainer_config.use_frequency_weighting),
                edc_loss(max_ir_len_ms, self.net.sample_rate, use_mask=trainer_config.use_edc_mask),
            ]
            self.loss_weights = [trainer_config.edr_loss_weight, trainer_config.edc_loss_weight]
        if self.use_colorless_loss:
            self.colorless_criterion = [amse_loss() if trainer_config.use_asym_spectral_loss else mse_loss(),
                                        sparsity_loss()]
            self.colorless_loss_weights = [trainer_config.spectral_loss_weight, trainer_config.sparsity_loss_weight]
        # The module path is ~120 kernel launches per step at batch size 32 and was launch bound (1.9 ms of kernels in
        # 5.2 ms of wall time): normalize + forward + losses + backward + Adam of a batch shape are captured once in a
        # CUDA graph and replayed (DGFDN_GRAPH_STEP=0 keeps every step eager). The random EDC mask is drawn on the
        # host per step (reference losses.py:221-223), which a replay cannot do: such configs stay eager.
        self.use_cuda_graph = (os.environ.get("DGFDN_GRAPH_STEP", "1") != "0" and not trainer_config.use_edc_mask
                               and torch.device(self.device).type == "cuda")
        self._graphs = {}

    def set_subband_filter(self, freq_resp: torch.Tensor):
        """Frequency response F[k] of the octave-band filter applied to H before the losses (trainer.py:112-150
        derive it from pyfar; here it is an input)."""
        self.subband_filter_freq_resp = freq_resp.to(device=self.device, dtype=torch.complex64)

    def init_scheduler(self, trainer_config: TrainerConfig):
        """Adam with per-name learning rates + StepLR(10, 0.1) (reference trainer.py:152-228)."""
        named = list(self.net.named_parameters())

        def pick(pred):
            return [p for n, p in named if pred(n)]

        keys = ('feedback_loop.alpha', 'input_gains', 'output_gains', 'output_svf_params', 'output_scalars',
                'sh_output_scalars', 'input_scalars')
        groups = [
            {'params': pick(lambda n: 'feedback_loop.alpha' in n), 'lr': trainer_config.coupling_angle_lr},
            {'params': pick(lambda n: 'output_gains' in n), 'lr': trainer_config.io_lr},
            {'params': pick(lambda n: 'input_gains' in n), 'lr': trainer_config.io_lr},
            {'params': pick(lambda n: 'output_svf_params' in n), 'lr': trainer_config.io_lr},
            {'params': pick(lambda n: 'input_scalars' in n), 'lr': trainer_config.io_lr},
            {'params': pick(lambda n: 'output_scalars' in n or 'sh_output_scalars' in n), 'lr': trainer_config.io_lr},
        ]
        others = pick(lambda n: not any(k in n for k in keys))
        if others:
            groups.append({'params': others, 'lr': trainer_config.lr})
        # capturable: the step counters live on the device, so optimizer.step() can sit inside a CUDA graph
        cuda = torch.device(self.device).type == "cuda"
        self.optimizer = torch.optim.Adam([g for g in groups if g['params']], capturable=cuda)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=10, gamma=0.1)

    def save_model(self, e: int):
        d = os.path.join(self.train_dir, 'checkpoints')
        os.makedirs(d, exist_ok=True)
        torch.save(self.net.state_dict(), os.path.join(d, f'model_e{e}.pt'))

    def apply_subband_filter(self, H: torch.Tensor) -> torch.Tensor:
        if self.subband_process_config is not None or self.subband_filter_freq_resp is not None:
            if self.subband_filter_freq_resp is None:
                raise RuntimeError("subband_process_config is set: provide the filter with set_subband_filter()")
            return H * self.subband_filter_freq_resp
        return H

    def calculate_losses(self, data: Dict, H: torch.Tensor, H_sub_fdn: Optional[Tuple] = None) -> Dict:
        """reference trainer.py:259-315"""
        if self.use_directional_fdn:
            amps = data['target_common_slope_amps']
            all_losses = {'edc_loss': self.loss_weights[0] * self.criterion[0](H, amps)}
        else:
            target = self._device_c64(data['target_rir_response'])
            edr_val = self.loss_weights[0] * self.criterion[0](target, H)
            edc_val = self.loss_weights[1] * self.criterion[1](target, H)
            all_losses = {'edc_loss': edc_val, 'edr_loss': edr_val}
        if self.use_colorless_loss:
            spectral = 0.0
            sparsity = 0.0
            per_group = ops.colorless_loss_per_group(H_sub_fdn[0], self.colorless_criterion[0].asym)
            for k in range(self.net.num_groups):
                spectral = spectral + self.colorless_loss_weights[0] * per_group[k]
                # `=`, not `+=`: only the last group's sparsity survives (reference trainer.py:305, quirk Q2)
                sparsity = self.colorless_loss_weights[1] * self.colorless_criterion[1](
                    self.net.feedback_loop.ortho_param(self.net.feedback_loop.M[k]))
            all_losses.update({'spectral_loss': spectral, 'sparsity_loss': sparsity})
        return all_losses

    def _device_c64(self, t: torch.Tensor) -> torch.Tensor:
        """Targets as complex64 device tensors, moved / converted once per SOURCE tensor (host or complex128 dataset
        fields are constant over training): the loss callables key their target-EDC / EDR caches on the tensor they are
        handed, so a fresh `.to(device)` copy per step would miss them every time and pile up stale entries."""
        if t.is_cuda and t.dtype == torch.complex64:
            return t
        cache = self.__dict__.get("_c64_cache")
        if cache is None:
            cache = self.__dict__["_c64_cache"] = TensorKeyedCache(max_entries=2)
        out = cache.get(t)
        if out is None:
            out = cache.put(t, t.to(device=self.device, dtype=torch.complex64))
        return out

    @torch.no_grad()
    def normalize(self, data: Dict):
        """Unit-energy normalisation of every sub-FDN: b_g, c_g /= (mean_k |H_sub[k,g]|^2)^(1/4) (reference
        trainer.py:317-332). Only the receiver-independent colorless solve is needed, not a full forward."""
        if not self.use_colorless_loss:
            return
        keep = self.net.return_per_delay_outputs
        self.net.return_per_delay_outputs = False
        try:
            h_sub, _ = self.net.sub_fdn_output(data['z_values'])
        finally:
            self.net.return_per_delay_outputs = keep
        energy = torch.mean(torch.abs(h_sub)**2, dim=0)  # (G,)
        scale = torch.pow(energy, 0.25).repeat_interleave(self.net.num_delay_lines_per_group).view(-1, 1)
        for name, prm in self.net.named_parameters():
            if name in ('input_gains', 'output_gains'):
                prm.data /= scale.to(prm.dtype)

    def _forward_losses(self, data: Dict):
        out = self.net(data)
        H, H_sub = out if self.use_colorless_loss else (out, None)
        H = self.apply_subband_filter(H)
        if self.use_directional_fdn:
            H = self.convert_ambi_rir_to_directional_rir(H)
        return self.calculate_losses(data, H, H_sub)

    def _step_body(self, data: Dict, with_norm: bool):
        if with_norm:
            self.normalize(data)
        self.optimizer.zero_grad(set_to_none=True)
        all_losses = self._forward_losses(data)
        loss = sum(all_losses.values())
        loss.backward()
        self.optimizer.step()
        return loss, all_losses

    def train_step(self, data: Dict, with_norm: bool = False):
        """reference trainer.py:452-477 / 795-825. with_norm=True runs normalize(data) first (what the epoch loop does
        before every step, trainer.py:366-377) inside the same captured graph."""
        if not self.use_cuda_graph:
            loss, all_losses = self._step_body(data, with_norm)
            return loss.item(), all_losses
        return self._graphed_step(data, with_norm)

    # ---- CUDA-graph replay of a whole step ---------------------------------------------------------------
    def _clear_target_caches(self):
        for crit in self.criterion:
            cache = getattr(crit, "_target_cache", None)
            if cache is not None:
                cache.clear()
        cache = self.__dict__.get("_c64_cache")
        if cache is not None:
            cache.clear()

    def _graphed_step(self, data: Dict, with_norm: bool):
        """First sighting of a batch signature: eager. Second: the batch is copied into static device buffers, the step
        runs eagerly on them (warming every cache keyed on the constant inputs) and is then captured. From the third on:
        copy + one graph launch. A change of learning rate (StepLR) re-captures."""
        tensors = {k: v for k, v in data.items() if torch.is_tensor(v)}
        sig = tuple(sorted((k, tuple(v.shape), str(v.dtype)) for k, v in tensors.items()))
        lrs = tuple(float(g['lr']) for g in self.optimizer.param_groups)
        key = (sig, with_norm)
        entry = self._graphs.get(key)
        if entry is not None and entry != "seen" and entry["lrs"] != lrs:
            entry = self._graphs[key] = "seen"  # the captured Adam kernels hold the old learning rates
        # Every eager step of a graphed trainer runs on ONE side stream, the stream of the later capture: autograd binds
        # the parameters' AccumulateGrad nodes to the stream they are first used on, and a captured backward that has
        # to synchronise with another (the default) stream invalidates the capture.
        if self.__dict__.get("_stream") is None:
            self._stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        if entry is None:
            self._graphs[key] = "seen"
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                loss, all_losses = self._step_body(data, with_norm)
            cur.wait_stream(self._stream)
            return loss.item(), all_losses
        if entry == "seen":
            static = dict(data)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                for k, v in tensors.items():
                    static[k] = v.to(self.device, copy=True)
                loss, all_losses = self._step_body(static, with_norm)  # this call's real step, and the warm-up
                result = loss.item(), {k: v.detach().clone() if torch.is_tensor(v) else v for k, v in all_losses.items()}
                del loss, all_losses  # nothing may keep this step's autograd graph (and its AccumulateGrad nodes) alive
                self._clear_target_caches()  # the capture must recompute everything derived from the batch
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                self.optimizer.zero_grad(set_to_none=True)
                with torch.cuda.graph(graph, stream=self._stream):
                    g_loss, g_all = self._step_body(static, with_norm)
                self._clear_target_caches()  # they now point into the graph's memory pool
            cur.wait_stream(self._stream)
            self._graphs[key] = dict(graph=graph, static=static, loss=g_loss, all=g_all, lrs=lrs,
                                     src={k: None for k in tensors})
            return result
        static, src = entry["static"], entry["src"]
        for k, v in tensors.items():
            ident = (v.data_ptr(), v._version)
            if k == 'z_values' and src[k] == ident:
                continue  # the frequency grid is the same tensor every step
            static[k].copy_(v, non_blocking=True)
            src[k] = ident
        entry["graph"].replay()
        return entry["loss"].item(), entry["all"]

    @torch.no_grad()
    def valid_step(self, data: Dict):
        all_losses = self._forward_losses(data)
        return sum(all_losses.values()).item(), all_losses

    def convert_ambi_rir_to_directional_rir(self, H_sh: torch.Tensor) -> torch.Tensor:
        """H_dir[b,j,k] = sum_l Y[j,l] H_sh[b,l,k] (reference trainer.py:853-865)."""
        return ops.mix_channels(self.net.sh_output_scalars.analysis_matrix, H_sh)

    def train(self, train_dataset, valid_dataset):
        """Epoch loop of reference trainer.py:345-450 (without the wav export at the end)."""
        self.train_loss, self.valid_loss = [], []
        self.individual_train_loss, self.individual_valid_loss = [], []
        st = time.time()
        self.save_model(-1)
        for epoch in range(self.max_epochs):
            et = time.time()
            tot, parts = 0.0, {}
            svf = getattr(self.net, "use_svf_in_output", False)
            if svf:  # full-band (SVF) models normalise b, c once per epoch, the others before every step (:366-377)
                self.normalize(next(iter(train_dataset)))
            for data in train_dataset:
                cur, cur_all = self.train_step(data, with_norm=not svf)
                tot += cur
                for k, v in cur_all.items():
                    parts[k] = parts.get(k, 0.0) + float(v)
            vtot, vparts = 0.0, {}
            for data in valid_dataset:
                cur, cur_all = self.valid_step(data)
                vtot += cur
                for k, v in cur_all.items():
                    vparts[k] = vparts.get(k, 0.0) + float(v)
            self.scheduler.step()
            self.train_loss.append(tot / max(1, len(train_dataset)))
            self.individual_train_loss.append({k: v / max(1, len(train_dataset)) for k, v in parts.items()})
            self.valid_loss.append(vtot / max(1, len(valid_dataset)))
            self.individual_valid_loss.append({k: v / max(1, len(valid_dataset)) for k, v in vparts.items()})
            self.save_model(epoch)
            print(f"epoch {epoch:3d}, train_loss {self.train_loss[-1]:.4f}, valid_loss {self.valid_loss[-1]:.4f}, "
                  f"time {time.time() - et:.3f}s")
            if epoch >= 1:
                self.early_stop = self.early_stop + 1 if abs(self.valid_loss[-2] - self.valid_loss[-1]) <= 1e-3 else 0
            if self.early_stop == self.patience:
                break
        print(f"Training time: {time.time() - st:.3f}s")


class VarReceiverPosTrainer(Trainer):
    """Omni GFDN over a grid of receivers (reference trainer.py:338-564)."""


class SinglePosTrainer(Trainer):
    """One measured RIR, one source-receiver pair (reference trainer.py:570-688): no batching, no validation split,
    energy-matched initial gains."""

    def __init__(self, net, trainer_config: TrainerConfig, filename: str = "ir"):
        super().__init__(net, trainer_config)
        self.filename = filename

    @torch.no_grad()
    def normalize(self, data: Dict):
        """Sub-FDN normalisation, then scale the per-group scalars so that the mean energy of H matches the target
        (reference :633-648; there `H` is only defined when the colorless loss is off -- a NameError otherwise --
        here the response is evaluated in both cases)."""
        super().normalize(data)
        out = self.net(data)
        H = out[0] if self.use_colorless_loss else out
        target = data['target_rir_response'].to(self.device)
        ratio = torch.mean(torch.abs(H)**2) / torch.mean(torch.abs(target)**2)
        for name, prm in self.net.named_parameters():
            if name in ('input_scalars', 'output_scalars'):
                prm.data /= torch.pow(ratio, 0.25).to(prm.dtype)

    def train(self, train_dataset, valid_dataset=None):
        data = next(iter(train_dataset))
        self.normalize(data)
        self.train_loss, self.individual_train_loss = [], []
        st = time.time()
        for epoch in range(self.max_epochs):
            et = time.time()
            for data in train_dataset:
                epoch_loss, parts = self.train_step(data)
            self.scheduler.step()
            self.train_loss.append(epoch_loss)
            self.individual_train_loss.append(parts)
            self.save_model(epoch)
            print(f"epoch {epoch:3d}, train_loss {epoch_loss:.4f}, time {time.time() - et:.3f}s")
            if epoch >= 1:
                self.early_stop = self.early_stop + 1 if abs(self.train_loss[-2] - self.train_loss[-1]) <= 1e-4 else 0
            if self.early_stop == self.patience:
                break
        print(f"Training time: {time.time() - st:.3f}s")


class DirectionalFDNVarReceiverPosTrainer(Trainer):
    """Directional FDN over a grid of receivers (reference trainer.py:690-921)."""
