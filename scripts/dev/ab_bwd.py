"""A/B timing of the warp backward: staged (shared-memory accumulation) vs generic (global atomics).  Dev aid; GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tps_pp_b200 as T
from tps_pp_b200 import _native as N, functional as TF
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = T.TPS_PP().to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
fg = torch.randn((B, 64, 32, 128), device=dev, generator=gen).requires_grad_()
x = torch.randn((B, 64, 16, 64), device=dev, generator=gen).requires_grad_()
cp = (m.get_parameter("TPE.localization_fc2.bias").detach().view(1, 32, 2).repeat(B, 1, 1)
      + 0.002 * torch.randn((B, 32, 2), device=dev, generator=gen)).requires_grad_()
sc = torch.tanh(0.5 * torch.randn((B, 1024, 32), device=dev, generator=gen)).requires_grad_()
at = m.atten_tps
g0 = torch.randn((B, 64, 16, 64), device=dev, generator=gen)
g1 = torch.randn((B, 64, 16, 64), device=dev, generator=gen)
for variant, name in ((N.VARIANT_AUTO, "staged"), (N.VARIANT_GENERIC, "generic"), (N.VARIANT_AUTO, "staged")):
    TF.BWD_VARIANT = variant
    out, mp = TF.tps_warp(fg, x, cp, sc, at.P_hat, at.P, at.hat_C, (16, 64))
    ts = []
    for i in range(8):
        for t in (fg, x, cp, sc):
            t.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.autograd.backward([out, mp], [g0, g1], retain_graph=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    nbytes = B * (1048576 + 262144 + 131072 + 2 * 262144 + 1048576 + 262144 + 131072 + 131072)
    print(f"{name}: backward {ts[len(ts)//2]*1e3:.1f} us (min {ts[0]*1e3:.1f}); {nbytes / ts[len(ts)//2] / 1e6:.0f} GB/s algorithmic "
          f"(src + gout + pc_score x2 read, gsrc + g_pc_score written)")
