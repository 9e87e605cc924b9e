#!/usr/bin/env python
"""Per-kernel device times of the fused warp's backward at B = 256 (torch profiler / CUPTI)."""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tps_pp_b200 import constants as K, functional as TF  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hat, ph, P, _ = K.attention_tps_buffers((2, 16), (16, 64))
hat, ph, P = (torch.from_numpy(t).to(dev) for t in (hat, ph, P))
fg = torch.randn((B, 64, 32, 128), device=dev, generator=g).requires_grad_()
x = torch.randn((B, 64, 16, 64), device=dev, generator=g).requires_grad_()
s = torch.tanh(0.5 * torch.randn((B, 1024, 32), device=dev, generator=g)).requires_grad_()
base = torch.from_numpy(K.attention_init_bias((2, 16))).float().to(dev)
cp = (base[None] + 0.002 * torch.randn((B, 32, 2), device=dev, generator=g)).contiguous().requires_grad_()
o0, o1 = TF.tps_warp(fg, x, cp, s, ph, P, hat, (16, 64))
g0, g1 = torch.randn_like(o0), torch.randn_like(o1)


def bwd():
    torch.autograd.grad((o0, o1), (fg, x, cp, s), (g0, g1), retain_graph=True)


for _ in range(3):
    bwd()
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        bwd()
    torch.cuda.synchronize()
rows = sorted(((e.key, e.self_device_time_total / n, e.count // n) for e in prof.key_averages()), key=lambda r: -r[1])
print("# warp backward, B = %d: %.1f us device time per call" % (B, sum(r[1] for r in rows)))
for k, us, c in rows[:10]:
    if us > 0.5:
        print("%9.1f us %3d x  %s" % (us, c, k[:110]))
