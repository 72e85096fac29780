"""Benchmark of the DiffGFDN hot path on B200: receiver.bin evals/s of one fwd+bwd training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], per-GPU shard): N = 24 delay lines in G = 3 groups, nfft = 2^18
(K = 2^17 + 1 bins), 12 500 receivers per GPU (100 000 on 8 GPUs -> weak scaling), scalar absorption, MLP receiver
gains, EDC (weight 10) + colorless losses, fs = 32 kHz, T60 = (0.3, 0.8, 1.5) s. One step = forward (MLP gains,
per-bin solve, colorless sub-FDN solve, projection of every receiver over every bin, irfft, EDC loss, colorless
loss) + backward to every parameter gradient (+ NCCL all-reduce of the flat gradient when N > 1).

The step never forms H per receiver: the per-bin system is solved once per bin, the inverse DFT runs on G rows, and
every receiver is handled in the time domain (diffgfdn_b200/fused.py) -- same loss and
gradients as the reference's project-then-irfft formulation (tests/test_gpu_fused.py).

`value`  : inputs resident in HBM in the data layer's device format (early-response windows hd = irfft(d)[window]
           and target EDC curves in dB, both float32 (B, tn), built once from the frequency-domain responses),
           timed with CUDA events.
`e2e`    : the same step with its inputs streamed every step from pinned host memory in the reference's layout
           (early + target responses, (B, K) complex64) and the loss read back -- host<->device copies and the
           per-receiver transforms inside the timed region.
`--impl reference`: the CPU port of the reference algorithm (oracle/gfdn_oracle.py) on the host cores, on a
           bounded sample of the same workload.
Prints ONE JSON line on rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 32000.0
T60 = (0.3, 0.8, 1.5)
N_LINES, N_GROUPS = 24, 3
NFFT = 2**18
SURVEY_BYTES_PER_EVAL = 64.0  # bytes per receiver.bin of the un-restructured pipeline
TD_BYTES_PER_SAMPLE = 8.0  # td_edc_step reads hd (4 B) + target dB (4 B) per receiver.sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--receivers", type=int, default=12500, help="receivers per GPU")
    ap.add_argument("--nfft", type=int, default=NFFT)
    ap.add_argument("--tile-rows", type=int, default=int(os.environ.get("DGFDN_TILE_ROWS", "296")))
    ap.add_argument("--e2e-tile-rows", type=int, default=int(os.environ.get("DGFDN_E2E_TILE_ROWS", "128")))
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps to time (0: the same count as --steps)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --receivers per GPU; strong: --total-receivers split over the GPUs")
    ap.add_argument("--total-receivers", type=int, default=100000, help="receivers of the whole job (--scaling strong)")
    ap.add_argument("--no-shard-bins", action="store_true",
                    help="N > 1: every rank solves all bins (K1 / K1c replicated) instead of sharding them")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true", help="skip the secondary metric (BASELINE configs[4] renderer)")
    ap.add_argument("--config", default="c4", choices=["c1", "c1svf", "c2", "c3", "c4"],
                    help="c4 (default): BASELINE configs[3], the headline; c1 / c1svf / c2 / c3: one JSON line for the "
                         "reference's own batch-32 configurations (configs[0..2]) through the drop-in trainer")
    ap.add_argument("--no-configs", action="store_true", help="default run: skip the configs[0..2] lines under 'configs'")
    ap.add_argument("--cpu-sample-receivers", type=int, default=32,
                    help="receivers of one CPU-baseline step (32 = the reference's own batch size, trainer batch_size)")
    args = ap.parse_args()
    if args.scaling == "strong":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.receivers = (args.total_receivers + world - 1) // world
    if args.e2e_steps <= 0:
        args.e2e_steps = args.steps
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower() == "active"
                                                         for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------
def delays_for(n_lines):
    from diffgfdn_b200.config import DiffGFDNConfig
    return DiffGFDNConfig(seed=235265, num_delay_lines=n_lines).delay_length_samps


def build_net(device, seed=1234):
    """The configs[3] model. device == "cpu_params_only": the same initialisation drawn on the CPU as float64 oracle
    parameters (no CUDA, no product code on the reference arm)."""
    if device == "cpu_params_only":
        return _oracle_init(seed)
    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    torch.manual_seed(seed)
    return DiffGFDNVarReceiverPos(FS, N_GROUPS, delays_for(N_LINES), device, FeedbackLoopConfig(use_zero_coupling=False),
                                  OutputFilterConfig(use_svfs=False), use_absorption_filters=False,
                                  common_decay_times=np.array([T60]), use_colorless_loss=True)


def _oracle_init(seed, hidden=3, neurons=128, feats=10):
    """Reference initialisation (model.py:100-106, feedback_loop.py:288-310, dnn.py:331-400 Kaiming-uniform linears,
    unit LayerNorms) as float64 leaves keyed like the reference state_dict."""
    gen = torch.Generator().manual_seed(seed)
    g, n = N_GROUPS, N_LINES
    l = n // g
    f64 = torch.float64
    p = {"feedback_loop.M": (2 * torch.rand(g, l, l, dtype=f64, generator=gen) - 1) / np.sqrt(l),
         "feedback_loop.alpha": np.pi / 4 * torch.rand(g * (g - 1) // 2, dtype=f64, generator=gen),
         "input_gains": (2 * torch.randn(n, 1, dtype=f64, generator=gen) - 1) / n,
         "output_gains": (2 * torch.randn(n, 1, dtype=f64, generator=gen) - 1) / n}
    dims = [6 * feats] + [neurons] * (hidden + 1)
    idx = 0
    for i in range(hidden + 1):
        bound = np.sqrt(6.0 / dims[i])
        p[f"output_scalars.mlp.model.{idx}.weight"] = (2 * torch.rand(dims[i + 1], dims[i], dtype=f64, generator=gen) - 1) * bound
        p[f"output_scalars.mlp.model.{idx}.bias"] = torch.zeros(dims[i + 1], dtype=f64)
        p[f"output_scalars.mlp.model.{idx + 1}.weight"] = torch.ones(dims[i + 1], dtype=f64)
        p[f"output_scalars.mlp.model.{idx + 1}.bias"] = torch.zeros(dims[i + 1], dtype=f64)
        idx += 3
    p[f"output_scalars.mlp.model.{idx}.weight"] = (2 * torch.rand(g, neurons, dtype=f64, generator=gen) - 1) * np.sqrt(6.0 / neurons)
    p[f"output_scalars.mlp.model.{idx}.bias"] = torch.zeros(g, dtype=f64)
    for v in p.values():
        v.requires_grad_(True)
    return {"params": p, "delays": torch.tensor(delays_for(n), dtype=f64), "feats": feats}


def render_metric(device, hbm_peak):
    """Secondary metric of BASELINE.json (configs[4]): block-recursive time-domain render of 10 s of late tail for 4096
    moving listeners x 8 octave bands from GFDN parameters (N = 12, G = 3 per band): K6 render_groups (the recursion,
    min(m) samples per block) + render_mix (per-listener gains switched every 100 ms, bands summed). Output resident in
    HBM: 4 B written per listener.sample."""
    from diffgfdn_b200 import ops
    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig
    from diffgfdn_b200.model import DiffGFDNVarReceiverPos
    bands, n_lines, listeners, seconds, positions = 8, 12, 4096, 10.0, 838
    t = int(seconds * FS)
    hop = int(0.1 * FS)
    gen = torch.Generator(device=device).manual_seed(77)
    pos = torch.rand(positions, 3, device=device, generator=gen)
    a, gam, b, c, s, dl = [], [], [], [], [], []
    with torch.no_grad():
        for bd in range(bands):
            torch.manual_seed(500 + bd)
            net = DiffGFDNVarReceiverPos(FS, N_GROUPS, delays_for(n_lines), device, FeedbackLoopConfig(use_zero_coupling=False),
                                         OutputFilterConfig(use_svfs=False), use_absorption_filters=False,
                                         common_decay_times=np.array([T60]), use_colorless_loss=False)
            a.append(net.feedback_loop.coupled_feedback_matrix_real().float())
            gam.append(net.feedback_loop.delay_line_gains.float())
            b.append(net.input_gains.reshape(-1).float())
            c.append(net.output_gains.reshape(-1).float())
            s.append(net.output_scalars.gains({'norm_listener_position': pos}).float())
            dl.append(net.delays.to(torch.int32))
    a, gam, b, c, s, dl = (torch.stack(v).contiguous() for v in (a, gam, b, c, s, dl))
    traj = torch.randint(0, positions, (listeners, (t + hop - 1) // hop), device=device, generator=gen, dtype=torch.int32)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(3):  # two warm-up passes, the third is timed
        ev[0].record()
        q = ops.render_groups(dl, a, gam, b, c, N_GROUPS, t)
        ev[1].record()
        out = ops.render_mix(s, traj, q, hop)
        ev[2].record()
        if it < 2:
            del out
    torch.cuda.synchronize()
    ms_groups, ms_mix = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    peak = float(out.abs().max())
    total = listeners * t
    # the reference's own moving-listener semantics (sound_examples.py:163-226 filter_overlap_add + the sub-band FIRs of
    # run_subband_training_treble.py:316-358): a 10 s stimulus through 2 s tails, hops of 100 ms cross-faded over 50 ms
    from diffgfdn_b200.inference import GFDNAuraliser
    firs = torch.randn(bands, 1025, device=device, generator=gen) * torch.hann_window(1025, periodic=False, device=device)
    aur = GFDNAuraliser(dl, a, gam, b, c, N_GROUPS, firs, device=device)
    stim = torch.randn(int(2.5 * FS), device=device, generator=gen)
    traj_ola = traj[:, :t // hop].long()
    ola_ms = []
    for it in range(2):
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record()
        y = aur.moving_listeners(stim, s, traj_ola, hop, int(2.0 * FS), int(0.05 * FS), 0.5)
        ev2[1].record()
        torch.cuda.synchronize()
        ola_ms.append(ev2[0].elapsed_time(ev2[1]))
    ola = {"metric": "moving-listener overlap-add, reference semantics (filter_overlap_add + sub-band FIRs)",
           "value": float(y.numel()) / (ola_ms[-1] * 1e-3), "unit": "listener*samples/s", "ms": ola_ms[-1],
           "finite": bool(torch.isfinite(y).all()),
           "note": "block recursion + batched cuFFT block responses + one (listeners x blocks.channels) x (blocks.channels x hop) "
                   "GEMM per output hop as 3 x TF32 on the tensor cores (cuBLAS); the reference convolves per listener and hop"}
    del y
    return {"metric": "render listener*samples/s (10 s late tail, 4096 moving listeners x 8 octave bands)",
            "moving_listeners_reference_semantics": ola,
            "value": total / ((ms_groups + ms_mix) * 1e-3), "unit": "listener*samples/s",
            "ms": {"render_groups(recursion, 8 bands)": ms_groups, "render_mix(4096 listeners)": ms_mix},
            "roofline": {"bound": "hbm", "kernel": "render_mix", "achieved": 4.0 * total / (ms_mix * 1e-3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": 4.0 * total / (ms_mix * 1e-3) / 1e9 / hbm_peak,
                         "note": "4 B written per listener.sample; the recursion is latency bound (T / min(m) dependent blocks)"},
            "config": {"workload": "BASELINE configs[4]", "bands": bands, "delay_lines": n_lines, "listeners": listeners,
                       "samples": t, "hop": hop, "positions": positions}, "finite": bool(np.isfinite(peak))}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (best effort), so that the pinned host pool of
    the end-to-end mode is allocated next to the GPU's PCIe root. Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device() if local_rank is None else local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "gpu reports no NUMA node"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"node {node}: none of its CPUs is available to this process"
        os.sched_setaffinity(0, allowed)
        return f"node {node} ({len(allowed)} CPUs)"
    except Exception as e:  # no NVML / sysfs in the sandbox: run unbound
        return f"unbound ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[0..2]: the reference's own training configurations (batch 32, nfft 2^17) through the drop-in API
# ------------------------------------------------------------------------------------------------------------
SMALL_CONFIGS = {
    "c1": "configs[0] omni full-band, N=12 (3x4), B=32 x 65537 bins, EDC(w=10)+EDR+colorless losses, scalar receiver gains",
    "c1svf": "configs[0] as shipped (use_svfs: True): SVF output filters from the MLP (11 sections per group and receiver)",
    "c2": "configs[1] one octave band of the sub-band training: configs[0] with the band filter F[k] applied to H (8 such "
          "models are independent: one per GPU, replicas only)",
    "c3": "configs[2] directional FDN, ambisonic order 2 (N=27 = 3x9), J=12 directions, B=32 x 65537 bins, "
          "directional EDC + colorless losses",
}


def module_path_metric(name, device, hbm_peak, steps=20, warmup=5, batch=32, nfft=131072):
    """One training step (normalize + forward + losses + backward + Adam) of the reference's trainer API at a
    configs[0..2] shape: `Trainer.train_step(data, with_norm=True)`, i.e. the CUDA-graph replay of the module path.
    Timed with CUDA events around `steps` calls, batch resident on the device. Also timed with host-resident batches
    (the DataLoader case): e2e."""
    import tempfile

    from diffgfdn_b200.config import FeedbackLoopConfig, OutputFilterConfig, TrainerConfig
    from diffgfdn_b200.model import DiffDirectionalFDNVarReceiverPos, DiffGFDNVarReceiverPos
    from diffgfdn_b200.trainer import DirectionalFDNVarReceiverPosTrainer, VarReceiverPosTrainer
    from diffgfdn_b200.utils import unit_circle_grid
    torch.manual_seed(0)
    t60 = np.array([list(T60)])
    k = nfft // 2 + 1
    tmp = tempfile.mkdtemp()
    gen = torch.Generator(device=device).manual_seed(1)
    pos = torch.rand(batch, 3, device=device, generator=gen)
    z = unit_circle_grid(nfft).to(device)
    directional = name == "c3"
    if directional:
        n_lines = 27
        net = DiffDirectionalFDNVarReceiverPos(FS, N_GROUPS, delays_for(n_lines), device, FeedbackLoopConfig(use_zero_coupling=False),
                                               OutputFilterConfig(use_svfs=False, num_hidden_layers=3, num_neurons_per_layer=128,
                                                                  num_fourier_features=10), 2, None, common_decay_times=t60,
                                               use_colorless_loss=True,
                                               analysis_matrix=(torch.randn(12, 9, generator=torch.Generator().manual_seed(2)) / 3).numpy())
        tr = DirectionalFDNVarReceiverPosTrainer(net, TrainerConfig(train_dir=tmp + "/o", ir_dir=tmp + "/i", num_freq_bins=nfft,
                                                                    use_colorless_loss=True, edc_loss_weight=10.0))
        data = dict(z_values=z, listener_position=pos, norm_listener_position=pos,
                    target_common_slope_amps=1e-4 + torch.rand(batch, 12, N_GROUPS, device=device, generator=gen))
    else:
        n_lines = 12
        net = DiffGFDNVarReceiverPos(FS, N_GROUPS, delays_for(n_lines), device, FeedbackLoopConfig(use_zero_coupling=False),
                                     OutputFilterConfig(use_svfs=name == "c1svf", num_hidden_layers=3, num_neurons_per_layer=128,
                                                        num_fourier_features=10), use_absorption_filters=False,
                                     common_decay_times=t60, use_colorless_loss=True)
        tr = VarReceiverPosTrainer(net, TrainerConfig(train_dir=tmp + "/o", ir_dir=tmp + "/i", num_freq_bins=nfft,
                                                      use_colorless_loss=True, use_asym_spectral_loss=True, edc_loss_weight=10.0))
        early, target = synth_responses(batch, nfft, device, 300)
        data = dict(z_values=z, listener_position=pos, norm_listener_position=pos, target_early_response=early,
                    target_rir_response=target)
        if name == "c2":  # a synthetic octave-band filter: rfft of a hann-windowed sinc band-pass around 1 kHz
            taps = 2047
            n = torch.arange(taps, dtype=torch.float64) - taps // 2
            lo, hi = 707.0 / FS, 1414.0 / FS
            fir = (2 * hi * torch.sinc(2 * hi * n) - 2 * lo * torch.sinc(2 * lo * n)) * torch.hann_window(taps, periodic=False,
                                                                                                       dtype=torch.float64)
            tr.set_subband_filter(torch.fft.rfft(fir, n=nfft).to(torch.complex64))
    with_norm = not getattr(net, "use_svf_in_output", False)
    for _ in range(max(3, warmup)):  # the first two calls are eager (second one captures), then replays
        tr.train_step(data, with_norm=with_norm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, _ = tr.train_step(data, with_norm=with_norm)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # end to end: the batch arrives from (pinned) host memory every step, as from the reference's DataLoader
    host = {kk: (v.cpu().pin_memory() if torch.is_tensor(v) and kk != "z_values" else v) for kk, v in data.items()}
    h2d = sum(v.numel() * v.element_size() for kk, v in host.items() if torch.is_tensor(v) and kk != "z_values")
    for _ in range(3):
        tr.train_step(host, with_norm=with_norm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(host, with_norm=with_norm)
    torch.cuda.synchronize()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / steps
    evals = float(batch) * k
    graphed = any(isinstance(v, dict) for v in tr._graphs.values())
    return {"metric": "DiffGFDN receiver*bin evals/s fwd+bwd", "value": evals / (ms * 1e-3), "unit": "receiver*bin evals/s",
            "ms_per_step": ms, "steps": steps, "loss": float(loss), "cuda_graph": graphed,
            "config": {"workload": SMALL_CONFIGS[name], "batch": batch, "bins": k, "delay_lines": n_lines, "groups": N_GROUPS,
                       "api": "Trainer.train_step(data, with_norm=True): normalize + forward + losses + backward + Adam"},
            "e2e": {"value": evals / (ms_e2e * 1e-3), "unit": "receiver*bin evals/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
            "roofline": {"bound": "hbm", "achieved": evals * SURVEY_BYTES_PER_EVAL / (ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": evals * SURVEY_BYTES_PER_EVAL / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "note": "whole step against SURVEY 8(d)'s 64 B per receiver.bin of the project -> irfft -> loss "
                                 "pipeline; at batch 32 the step is ~100 small kernels (latency bound), not bandwidth bound"}}


def synth_responses(rows, nfft, device, seed):
    """Synthetic targets: 1 s of exponentially decaying noise per receiver -> rfft; the early
    response is the first 20 ms with a fade-out. Generated on the device in chunks. Returns complex64 tensors."""
    gen = torch.Generator(device=device).manual_seed(seed)
    k = nfft // 2 + 1
    tlen = int(FS)
    t = torch.arange(tlen, device=device, dtype=torch.float32)
    fade = torch.ones(640, device=device)
    fade[560:] = torch.hann_window(160, periodic=False, device=device)[80:]
    tgt = torch.empty(rows, k, dtype=torch.complex64, device=device)
    early = torch.empty(rows, k, dtype=torch.complex64, device=device)
    for r0 in range(0, rows, 256):
        r1 = min(rows, r0 + 256)
        tau = (0.1 + 0.2 * torch.rand(r1 - r0, 1, device=device, generator=gen)) * FS
        rir = torch.randn(r1 - r0, tlen, device=device, generator=gen) * torch.exp(-t / tau)
        tgt[r0:r1] = torch.fft.rfft(rir, n=nfft)
        early[r0:r1] = torch.fft.rfft(rir[:, :640] * fade, n=nfft)
    return early, tgt


def timed_steps(fn, steps, barrier):
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / steps, out


def event_breakdown(events, steps):
    """Per-step device time of each instrumented kernel / section from the CUDA events recorded inside the timed
    region (same stream as the launches)."""
    out = {}
    for name, pairs in events.items():
        ms = sum(a.elapsed_time(b) for a, b in pairs)
        out[name] = {"ms_per_step": ms / steps, "launches_per_step": len(pairs) / steps}
    return out


# ------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------------------------
def oracle_step(params, delays, z, pos, early, target, feats, edc_w=10.0):
    """One fwd + EDC/colorless loss + bwd of the reference algorithm (oracle/gfdn_oracle.py: float64 CPU port, the
    reference's own formulation -- dense inverse per bin, projection of every receiver over every bin, one irfft per
    receiver) on the given parameters (a float64 state_dict; leaves with requires_grad get .grad). Returns the loss
    terms, and the seconds it took."""
    from oracle import gfdn_oracle as O
    g = params["feedback_loop.M"].shape[0]
    t0 = time.perf_counter()
    gamma = O.decay_times_to_gain_per_sample(T60, delays.tolist(), FS, g)
    a = O.coupled_feedback_matrix(params["feedback_loop.M"], params["feedback_loop.alpha"])
    b, c = params["input_gains"].reshape(-1), params["output_gains"].reshape(-1)
    s = O.gains_from_mlp(pos, params, feats, g)
    H = O.omni_response(z, delays, gamma, a, b, c, s, early)
    h_sub, _ = O.sub_fdn_output(z, delays, params["feedback_loop.M"], b, c)
    edc = O.edc_loss(target, H, max(T60) * 1e3, FS)
    spec, spars = O.colorless_losses(h_sub, params["feedback_loop.M"], 1.0, 1.0, asym=True)
    (edc_w * edc + spec + spars).backward()
    return dict(edc=float(edc.detach()), spec=float(spec.detach()), spars=float(spars.detach())), time.perf_counter() - t0


def oracle_params(net):
    """float64 CPU copies of net's state_dict; the trainable ones are autograd leaves."""
    names = {k for k, _ in net.named_parameters()}
    return {k: v.detach().cpu().to(torch.float64).requires_grad_(k in names) for k, v in net.state_dict().items()
            if v.dtype.is_floating_point}


def cpu_reference_step(nfft, receivers, seed=7):
    """One CPU step on freshly drawn parameters and data (the --impl reference arm): seconds and receiver.bin evals."""
    from oracle import gfdn_oracle as O
    net = build_net("cpu_params_only", seed=seed)
    rng = np.random.default_rng(seed)
    tlen = int(FS)
    rir = rng.standard_normal((receivers, tlen)) * np.exp(-np.arange(tlen)[None, :] / (0.2 * FS))
    target = torch.tensor(np.fft.rfft(rir, n=nfft, axis=-1))
    early = torch.tensor(np.fft.rfft(rir[:, :640], n=nfft, axis=-1))
    pos = torch.tensor(rng.uniform(0, 1, (receivers, 3)))
    _, secs = oracle_step(net["params"], net["delays"], O.z_grid(nfft), pos, early, target, net["feats"])
    return secs, receivers * (nfft // 2 + 1)


def cpu_baseline(step, net, early_pool, target_pool, positions, args):
    """The oracle timed on the host cores on a bounded sample of the SAME workload -- the first cpu_sample_receivers
    receivers of this rank's shard, the GPU net's CURRENT parameters -- and, since both arms then computed the same
    numbers, the parity of the timed GPU path against it (EDC in dB, worst relative gradient error)."""
    from diffgfdn_b200.fused import ShardedEDCStep
    torch.set_num_threads(os.cpu_count() or 1)
    n = min(args.cpu_sample_receivers, early_pool.shape[0], positions.shape[0])
    sub = ShardedEDCStep(net, max(T60) * 1e3, edc_weight=10.0)
    sub.attach(step.z, positions[:n], None, None)
    sub.attach(step.z, positions[:n], sub.precompute_early_window(early_pool[:n]), sub.precompute_target_db(target_pool[:n]))
    out = sub.step()
    torch.cuda.synchronize()
    params = oracle_params(net)
    losses, secs = oracle_step(params, net.delays.cpu().to(torch.float64), step.z.cpu(), positions[:n].cpu().to(torch.float64),
                               early_pool[:n].cpu().to(torch.complex128), target_pool[:n].cpu().to(torch.complex128),
                               net.output_scalars.encoder.num_fourier_features)
    worst = max(float((q.grad.detach().cpu().double() - params[k].grad).abs().max() / params[k].grad.abs().max())
                for k, q in net.named_parameters())
    k = args.nfft // 2 + 1
    return {"value": n * k / secs, "unit": "receiver*bin evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} receivers x {k} bins, N={N_LINES}, one fwd+loss+bwd of oracle/gfdn_oracle.py (float64 "
                      f"torch-CPU port of the reference) on the GPU net's parameters, {secs:.1f} s",
            "parity_vs_gpu_step": {"receivers": n, "edc_db_gpu": float(out["edc_loss"]) / 10.0, "edc_db_oracle": losses["edc"],
                                   "edc_db_abs_diff": abs(float(out["edc_loss"]) / 10.0 - losses["edc"]),
                                   "spectral_rel_diff": abs(float(out["spectral_loss"]) - losses["spec"]) / abs(losses["spec"]),
                                   "worst_grad_rel_err": worst, "at": "initial parameters",
                                   "tolerances": "EDC 0.01 dB, gradients 1e-3 relative (BASELINE.json)"}}


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    evals = 0
    for i in range(args.warmup + args.steps):
        secs, evals = cpu_reference_step(args.nfft, args.cpu_sample_receivers, seed=7 + i)
        if i >= args.warmup:
            times.append(secs)
    ms = 1e3 * float(np.mean(times))
    val = evals / (ms / 1e3)
    base = {"value": val, "unit": "receiver*bin evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"each step = {args.cpu_sample_receivers} receivers x {args.nfft // 2 + 1} bins (bounded sample "
                      f"of the per-GPU shard of {args.receivers}), oracle/gfdn_oracle.py"}
    print(json.dumps({
        "impl": "reference", "metric": "DiffGFDN receiver*bin evals/s fwd+bwd", "value": val,
        "unit": "receiver*bin evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args), "cpu_baseline": base,
        "e2e": {"value": val, "unit": "receiver*bin evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(args):
    shard = (f"{args.total_receivers} receivers split over the GPUs ({args.receivers}/GPU)" if args.scaling == "strong"
             else f"per-GPU shard of {args.receivers} receivers")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    bins = "bins of the per-bin solves sharded over the ranks" if world > 1 and not args.no_shard_bins else "per-bin solves on every rank"
    return {"workload": f"BASELINE configs[3], {shard}: N={N_LINES} lines, G={N_GROUPS} groups, "
                        f"x {args.nfft // 2 + 1} bins (nfft={args.nfft}), EDC(w=10)+colorless losses, fwd+bwd+Adam",
            "receivers_per_gpu": args.receivers, "bins": args.nfft // 2 + 1, "delay_lines": N_LINES,
            "groups": N_GROUPS, "tile_rows": args.tile_rows,
            "parallelism": f"receiver-sharded dp{args.gpus}, {bins}",
            "l2_policy": "inputs larger than L2 (resident hd + target dB = 2 x receivers x tn x 4 B >> 126 MB, "
                         "streamed once per step), no explicit flush"}


# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    from diffgfdn_b200 import _lib, build
    build.build()
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the kernels have no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)  # before any pinned allocation: first touch decides where the pool lives
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    if args.config != "c4":  # one line for a configs[0..2] shape; N > 1 = independent replicas (sub-band models: one per GPU)
        hbm, how = peaks()
        with ClockSampler(local) as clocks:
            line = module_path_metric(args.config, device, hbm, steps=args.steps, warmup=args.warmup)
        vals = torch.tensor([line["value"], line["e2e"]["value"]], device=device, dtype=torch.float64)
        worst = torch.tensor([line["ms_per_step"]], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.SUM)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        if rank == 0:
            line.update({"value": float(vals[0]), "ms_per_step": float(worst[0]), "n_gpus": world, "warmup": args.warmup,
                         "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
                         "dtype": "f32 storage / f64 per-bin solve", "clocks": clocks.summary(), "gpu_launches": args.steps,
                         "parallelism": f"{world} independent replicas" if world > 1 else "single GPU"})
            line["e2e"]["value"] = float(vals[1])
            line["roofline"]["peak_source"] = how
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    from diffgfdn_b200.fused import ShardedEDCStep
    from diffgfdn_b200.utils import unit_circle_grid
    net = build_net(device)
    if world > 1:  # identical parameters on every rank
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    shard_bins = world > 1 and not args.no_shard_bins
    step = ShardedEDCStep(net, max(T60) * 1e3, tile_rows=args.tile_rows, edc_weight=10.0, world_size=world,
                          total_receivers=args.receivers * world, e2e_tile_rows=args.e2e_tile_rows, shard_bins=shard_bins)
    k = args.nfft // 2 + 1
    z = unit_circle_grid(args.nfft, device=device)
    gen = torch.Generator(device=device).manual_seed(100 + rank)
    positions = torch.rand(args.receivers, 3, device=device, generator=gen)
    step.attach(z, positions, None, None)
    # resident inputs: early-response windows and target EDC curves (float32 (B, tn)), built once from a pool of
    # synthetic frequency-domain responses
    pool_rows = min(args.receivers, 1024)
    early_pool, target_pool = synth_responses(pool_rows, args.nfft, device, 200 + rank)
    tdb_pool = step.precompute_target_db(target_pool)
    hd_pool = step.precompute_early_window(early_pool)
    reps = (args.receivers + pool_rows - 1) // pool_rows
    hd = hd_pool.repeat(reps, 1)[:args.receivers].contiguous()
    target_db = tdb_pool.repeat(reps, 1)[:args.receivers].contiguous()
    step.attach(z, positions, hd, target_db)
    # CPU baseline + parity of the timed path, on the INITIAL parameters (before any optimizer step), rank 0 only
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(step, net, early_pool, target_pool, positions, args)
    barrier()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True, fused=True)  # one multi-tensor kernel per step

    def eager_step():
        losses = step.step()
        opt.step()
        return losses

    for _ in range(args.warmup):
        eager_step()
    # (1) instrumented eager steps: CUDA events around the receiver kernels and the step sections (same stream),
    #     used for the per-kernel roofline and the stage breakdown
    step.events = {}
    ms_eager, _ = timed_steps(eager_step, args.steps, barrier)
    stages = event_breakdown(step.events, args.steps)
    step.events = None
    # (2) the headline: the same step (+ Adam) captured once in a CUDA graph and replayed -- no host work per step
    if args.no_graph:
        one_step = eager_step
    else:
        step.capture(optimizer=opt, warmup=1)
        one_step = step.replay
    for _ in range(args.warmup):
        one_step()
    step.kernel_launches = 0
    with ClockSampler(local) as clocks:
        ms, losses = timed_steps(one_step, args.steps, barrier)
        launches = step.kernel_launches
        loss_val = float(sum(v for v in losses.values()))  # of the last timed step (the replays below train on)
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # the timed region lasts a few milliseconds, one nvidia-smi poll takes longer: keep the SAME step running
        # (untimed) for ~1 s so that the clock / throttle samples are taken under this load. The count is derived from
        # the rank-agreed step time: every rank replays the same number of (all-reducing) steps.
        for _ in range(min(2000, max(20, int(1000.0 / max(ms, 0.5))))):
            one_step()
        torch.cuda.synchronize()
    evals_per_step = float(args.receivers) * world * k
    value = evals_per_step / (ms / 1e3)

    # end-to-end: inputs from pinned host memory every step, loss read back
    e2e = None
    if not args.no_e2e:
        host_d = early_pool.cpu().pin_memory()
        host_t = target_pool.cpu().pin_memory()

        def e2e_step():
            ls = step.step(host_d=host_d, host_target=host_t)
            opt.step()
            return float(sum(v for v in ls.values()))  # device -> host read of the loss

        e2e_step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        mine = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([mine], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = int(step.h2d_bytes) + int(positions.numel() * 4)
        e2e = {"value": evals_per_step / float(te.item()), "unit": "receiver*bin evals/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
               "ms_per_step": 1e3 * float(te.item()), "steps": args.e2e_steps,
               "h2d_GBps_per_rank": h2d / float(te.item()) / 1e9, "h2d_GBps_all_ranks": world * h2d / float(te.item()) / 1e9,
               "host_numa_binding": numa,
               "note": "early + target responses ((B, K) complex64, reference layout) streamed from a pinned host pool "
                       "every step (bins 0..K/2, the ones irfft(X, n=K) reads), early window + target EDC rebuilt on "
                       "the device, loss read back. Bound by host->device copies: one PCIe 5 x16 link per GPU (~55 GB/s) "
                       "at N <= 2; at N >= 4 GPUs behind the same PCIe switch share its uplink (see h2d_GBps_all_ranks)"}

    if rank == 0:
        hbm, how = peaks()
        # dominant receiver-scaling kernel: 8 algorithmic bytes per receiver.sample
        if step.use_fused:
            td_key, rows_per_launch = "td_edc_fused", args.receivers
            from diffgfdn_b200 import ops
            info = ops.td_fused_info(net.num_groups, step.tn)
            td_name = ("td_fused_kernel<G=3, variant %d: %d threads, cluster of %d> (K3d: cluster of 8 CTAs per receiver "
                       "row, TMA-staged inputs, three rows per iteration, mix + EDC + dB loss + whole backward, ghy "
                       "accumulators resident in tensor memory (tcgen05.ld/st); its finalize launch is inside the timed "
                       "pair)" % (info["variant"], info["threads"], info["cluster_size"]))
        else:
            td_key, rows_per_launch = "td_edc_step", min(args.tile_rows, args.receivers)
            td_name = "td_edc_step_kernel<3,true> (K3c: mix + EDC + dB loss + backward per receiver row)"
        td = stages[td_key]
        td_bytes = TD_BYTES_PER_SAMPLE * args.receivers * step.tn  # per step, this rank
        achieved = td_bytes / (td["ms_per_step"] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from `ncu --set full`
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(td_key)
        out = {
            "metric": "DiffGFDN receiver*bin evals/s fwd+bwd", "value": value, "unit": "receiver*bin evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 storage / f64 per-bin solve and scans",
            "data": "synthetic", "config": workload_config(args), "loss": loss_val, "clocks": clocks.summary(),
            "exchanges": ("none (one rank)" if world == 1 else
                          "own push / wait kernels over NVLink peer memory (csrc/peer.cu), inside the captured graph"
                          if step.peer is not None else "NCCL collectives inside the captured graph"),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "peak_source": how, "kernel": td_name,
                         "algorithmic_bytes_per_launch": TD_BYTES_PER_SAMPLE * rows_per_launch * step.tn,
                         "avg_launch_ms": td["ms_per_step"] / td["launches_per_step"],
                         "share_of_step": td["ms_per_step"] / ms,
                         "timing": "CUDA events around every launch of this kernel during %d instrumented eager steps "
                                   "of the same workload (events cannot be recorded inside a graph replay)" % args.steps,
                         "note": "8 B per receiver.sample (hd 4 + target dB 4); tn = %d samples per receiver" % step.tn},
            # the same throughput expressed with SURVEY.md 8(d)'s 64 B per receiver.bin of the project-then-irfft
            # pipeline this design replaces (can exceed 1: those bytes are no longer moved)
            "survey_64B_equivalent": {"GBps": value / world * SURVEY_BYTES_PER_EVAL / 1e9,
                                      "frac_of_peak": value / world * SURVEY_BYTES_PER_EVAL / 1e9 / hbm},
            "stages": stages, "ms_per_step_eager_instrumented": ms_eager,
            "cuda_graph": not args.no_graph,
            "e2e": e2e,
        }
        if not args.no_render:
            out["render"] = render_metric(device, hbm)
        if not args.no_configs:  # the reference's own batch-32 configurations through the drop-in trainer (graph replay)
            out["configs"] = {c: module_path_metric(c, device, hbm, steps=max(5, min(args.steps, 20))) for c in SMALL_CONFIGS}
        if cpu_base is not None:
            out["cpu_baseline"] = cpu_base
        print(json.dumps(out), flush=True)
    if world > 1:
        # The captured graph holds NCCL all-reduce nodes: release it (and everything that keeps its memory pool alive)
        # before the communicator goes away, then tear down in order.
        dist.barrier()
        torch.cuda.synchronize()
        step.release_graph()
        del opt
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        done = threading.Event()

        def _teardown():
            try:
                dist.destroy_process_group()
            finally:
                done.set()

        threading.Thread(target=_teardown, daemon=True).start()
        if not done.wait(30.0):  # a wedged communicator must not hold the box: the JSON line is out, leave
            os._exit(0)


if __name__ == "__main__":
    main()
