"""Dev tool: prints the clock64 timeline written by an INSTRUMENTED build of libtpspp.so (a temporary `g_dbg` device
array + `tpspp_dbg_read` export added by hand to the kernel under study; see profiles/r01_conv_timeline.md).  It does
not work against the shipped library."""
import sys, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import tpspp_oracle as O
import tps_pp_b200 as T
DEV='cuda:0'
m = T.TPS_PP().to(DEV).eval(); m.load_state_dict(O.trained_like_state(3), strict=True)
B=256
g = torch.Generator(device=DEV).manual_seed(5)
x = torch.randn((B, 64, 16, 64), device=DEV, generator=g); o0 = torch.randn((B, 32, 32, 128), device=DEV, generator=g); o1 = torch.randn((B, 32, 32, 128), device=DEV, generator=g)
with torch.no_grad():
    for _ in range(3): m(x, [o0, o1])
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 4096)()
lib = ctypes.CDLL('/root/repo/tps_pp_b200/libtpspp.so')
print('rc', lib.tpspp_dbg_read(buf))
a = np.array(buf[:])
t0 = a[8 * 16 + 0]
print('gj | MMA: loop_top fc1_next_issued a2_full_seen fc2_c0_issued fc2_c1_issued | EPI: wait_d1 d1_seen gelu_done a2_empty_ok a2_full_arrived')
for gj in range(8, 20):
    r = a[gj * 16: gj * 16 + 13] - t0
    print(gj, r[0], r[1], r[2], r[3], r[4], '|', r[8], r[9], r[10], r[11], r[12])
print('a1_full seen per tile', (a[2048:2056] - t0).tolist())
