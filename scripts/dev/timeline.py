import sys, ctypes, torch, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import tpspp_oracle as O
import tps_pp_b200 as T
from tps_pp_b200 import _native as N
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
a = np.array(buf[:]).reshape(512, 8)
t0 = a[20, 0]
print('chunk  M:a_full  M:w_full  M:issued | P:start  P:a_empty_ok  P:st_done  P:arrived   (cycles rel.)')
for ch in range(20, 60):
    r = a[ch] - t0
    print(ch, r[0], r[1], r[2], '|', r[3], r[4], r[5], r[6])

b = np.array(buf[:])[2048:2048+512].reshape(128, 4)
print('gcc: before_issue  before_t_full_wait  after_t_full_wait (rel)')
for gc in range(2, 8):
    r = b[gc] - t0
    print(gc, r[0], r[1], r[2])
