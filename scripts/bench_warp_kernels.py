#!/usr/bin/env python
"""Micro-benchmark of the fused warp forward/backward at BASELINE config-2/4 sizes (CUDA events, L2-cold inputs
by construction: > 126 MB per call).  Prints one JSON line per case with algorithmic GB/s."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tps_pp_b200 import _native as N, constants as K, functional as TF  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    out = []
    # TPS++ geometry
    B = 256
    hat, ph, P, _ = K.attention_tps_buffers((2, 16), (16, 64))
    hat, ph, P = (torch.from_numpy(t).to(dev) for t in (hat, ph, P))
    fg = torch.randn((B, 64, 32, 128), device=dev, generator=g)
    x = torch.randn((B, 64, 16, 64), device=dev, generator=g)
    s = torch.tanh(0.5 * torch.randn((B, 1024, 32), device=dev, generator=g))
    base = torch.from_numpy(K.attention_init_bias((2, 16))).float().to(dev)
    cp = (base[None] + 0.002 * torch.randn((B, 32, 2), device=dev, generator=g)).contiguous()
    fwd_bytes = B * 1966336
    for name, variant in (("staged", N.VARIANT_STAGED), ("generic", N.VARIANT_GENERIC)):
        ms = timeit(lambda: TF.tps_warp(fg, x, cp, s, ph, P, hat, (16, 64), variant=variant))
        out.append(dict(case=f"tpspp_fwd_{name}", batch=B, ms=ms, gbs=fwd_bytes / ms / 1e6))
    fgr, xr, cpr, sr = fg.clone().requires_grad_(), x.clone().requires_grad_(), cp.clone().requires_grad_(), s.clone().requires_grad_()
    o0, o1 = TF.tps_warp(fgr, xr, cpr, sr, ph, P, hat, (16, 64))
    g0, g1 = torch.randn_like(o0), torch.randn_like(o1)

    def bwd():
        torch.autograd.grad((o0, o1), (fgr, xr, cpr, sr), (g0, g1), retain_graph=True)
    ms = timeit(bwd, iters=10, warm=3)
    bwd_bytes = B * (1966336 + 2 * 262144 + 1048576 + 262144 + 131072)   # + gout reads, gsrc writes, g_score
    out.append(dict(case="tpspp_bwd", batch=B, ms=ms, gbs=bwd_bytes / ms / 1e6))

    # classical high-res (config 4)
    for F_ in (20, 40):
        Bc = 1024
        inv, phc, _ = K.classical_tps_buffers(F_, (64, 256))
        inv, phc = torch.from_numpy(inv).to(dev), torch.from_numpy(phc).to(dev)
        img = torch.randn((Bc, 3, 64, 256), device=dev, generator=g)
        cb = torch.from_numpy(K.classical_init_bias(F_)).float().to(dev)
        cpc = (cb[None] + 0.02 * torch.randn((Bc, F_, 2), device=dev, generator=g)).contiguous()
        ms = timeit(lambda: TF.tps_warp(img, None, cpc, None, phc, None, inv, (64, 256), mode=N.MODE_CLASSICAL, theta=0.0),
                    iters=10, warm=3)
        out.append(dict(case=f"classical_64x256_F{F_}_fwd", batch=Bc, ms=ms, gbs=Bc * (2 * 3 * 64 * 256 * 4 + 8 * F_) / ms / 1e6,
                        img_per_s=Bc / ms * 1e3))
    Bc = 256
    inv, phc, _ = K.classical_tps_buffers(20, (32, 100))
    inv, phc = torch.from_numpy(inv).to(dev), torch.from_numpy(phc).to(dev)
    img = torch.randn((Bc, 3, 32, 100), device=dev, generator=g)
    cb = torch.from_numpy(K.classical_init_bias(20)).float().to(dev)
    cpc = (cb[None] + 0.02 * torch.randn((Bc, 20, 2), device=dev, generator=g)).contiguous()
    ms = timeit(lambda: TF.tps_warp(img, None, cpc, None, phc, None, inv, (32, 100), mode=N.MODE_CLASSICAL, theta=0.0))
    out.append(dict(case="classical_32x100_F20_fwd", batch=Bc, ms=ms, img_per_s=Bc / ms * 1e3))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
