import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from diffgfdn_b200 import ops
from diffgfdn_b200.utils import unit_circle_grid
import bench
dev = torch.device("cuda", 0)
def try_capture(name, fn):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name)
    except Exception as e:
        print("FAIL", name, str(e).split("\n")[0][:150])
        torch.cuda.synchronize()
nfft = 8192
z = unit_circle_grid(nfft, device=dev)
k = z.numel()
n, g = 12, 3
delays = torch.tensor(bench.delays_for(n), dtype=torch.int32, device=dev)
a = torch.randn(n, n, device=dev).mul_(0.2).requires_grad_(True)
gamma = torch.full((n,), 0.9, device=dev, requires_grad=True)
b = torch.randn(n, device=dev, requires_grad=True); c = torch.randn(n, device=dev, requires_grad=True)
m = torch.randn(g, n // g, n // g, device=dev).mul_(0.3).requires_grad_(True)
def f_solve():
    x, y = ops.gfdn_solve(z, delays, a, gamma, b, c, g)
    return y
def fb(fn):
    def run():
        for t in (a, gamma, b, c, m): t.grad = None
        out = fn()
        out.abs().sum().backward() if out.is_complex() else out.sum().backward()
    return run
try_capture("solve fwd", f_solve)
try_capture("solve fwd+bwd", fb(f_solve))
try_capture("groups fwd+bwd", fb(lambda: ops.gfdn_solve_groups(z, delays, m, None, b, c)[1]))
try_capture("expm fwd+bwd", fb(lambda: ops.skew_expm(m)))
try_capture("irfft fwd+bwd", fb(lambda: ops.irfft_window(f_solve().transpose(0, 1), k, 100, 2000)))
try_capture("colorless fwd+bwd", fb(lambda: ops.colorless_loss_per_group(f_solve(), True)))

# ---- pieces of the fused step
net = bench.build_net(dev)
pos = torch.rand(64, 3, device=dev)
def zero():
    for p in net.parameters(): p.grad = None
def p_mlp():
    zero(); net.output_scalars.gains({'norm_listener_position': pos}).sum().backward()
def p_A():
    zero(); net.feedback_loop.coupled_feedback_matrix_real().sum().backward()
def p_phi():
    zero(); net.feedback_loop.construct_coupling_matrix().sum().backward()
def p_block():
    zero(); net.feedback_loop.construct_block_mixing_matrix().sum().backward()
def p_spars():
    zero()
    from diffgfdn_b200.fused import ShardedEDCStep
    ShardedEDCStep._sparsity(net.feedback_loop.ortho_param(net.feedback_loop.M[2])).backward()
try_capture("mlp", p_mlp)
try_capture("phi", p_phi)
try_capture("block", p_block)
try_capture("A", p_A)
try_capture("sparsity", p_spars)
