"""Dev tool: prints the clock64 timeline written by an INSTRUMENTED build of libtpspp.so (-DTPSPP_TIMELINE: conv_pair_kernel
records events of CTA 0 for global chunk steps 128..191 into g_tl; `tpspp_dbg_read` copies it out).  The last
conv_pair_kernel launch of the forward (dec3) is what remains in the buffer.  Not part of the shipped library."""
import sys, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
import tps_pp_b200 as T
from tps_pp_b200 import _native as N
DEV = 'cuda:0'
m = T.TPS_PP().to(DEV).eval()
B = 256
g = torch.Generator(device=DEV).manual_seed(5)
x = torch.randn((B, 64, 16, 64), device=DEV, generator=g); o0 = torch.randn((B, 32, 32, 128), device=DEV, generator=g); o1 = torch.randn((B, 32, 32, 128), device=DEV, generator=g)
with torch.no_grad():
    for _ in range(3): m(x, [o0, o1])
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 2048)()
lib = N.lib()
print('rc', lib.tpspp_dbg_read(buf))
a = np.array(buf[:]).reshape(64, 32)
t0 = a[0, 0]
print('step | MMA: top waited issued | P warp0: lds a_empty st_issued st_waited arrived | P warp15: same')
for i in range(0, 40):
    r = a[i] - t0
    print(128 + i, r[0:3].tolist(), '|', r[8:13].tolist(), '|', r[16:21].tolist())
