"""A/B timing of the fused down kernel vs the four separate launches (dev aid; GPU)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tps_pp_b200 as T
from tps_pp_b200 import _native as N
sys.argv = sys.argv[:1]
import bench
dev = torch.device("cuda:0")
B = 256
gen = torch.Generator(device=dev).manual_seed(1234)
x = torch.randn((B, 64, 16, 64), device=dev, generator=gen)
o0 = torch.randn((B, 32, 32, 128), device=dev, generator=gen)
o1 = torch.randn((B, 32, 32, 128), device=dev, generator=gen)
m = T.TPS_PP().to(dev).eval()
bench._trained_like_(m)
lib = N.lib()
for flags in (0, N.HEAD_FLAG_UNFUSED_DOWN | N.HEAD_FLAG_UNFUSED_SCORE, 0):
    m.head_flags = flags
    with torch.no_grad():
        for _ in range(5):
            m(x, [o0, o1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m(x, [o0, o1])
        e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_float * 64)(); cnt = ctypes.c_int(0)
        lib.tpspp_launch_profile(1)
        acc = None
        for _ in range(5):
            m(x, [o0, o1])
            lib.tpspp_launch_profile_read(buf, 64, ctypes.byref(cnt))
            cur = [float(buf[i]) for i in range(cnt.value)]
            acc = cur if acc is None else [a + b for a, b in zip(acc, cur)]
        lib.tpspp_launch_profile(0)
    names = bench._launch_names(len(acc)) or [str(i) for i in range(len(acc))]
    print(f"flags={flags}: {e0.elapsed_time(e1) / 20:.4f} ms/step;", {n: round(v / 5, 4) for n, v in zip(names, acc)})
