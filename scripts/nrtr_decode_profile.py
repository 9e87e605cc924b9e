import sys, torch
sys.path.insert(0, '.')
import tps_pp_b200 as T
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = T.NRTRDecoder().to(dev).eval()
x = torch.randn((256, 64, 512), device=dev)
with torch.no_grad():
    for _ in range(2): m.forward_test(None, x, None)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.forward_test(None, x, None)
        torch.cuda.synchronize()
rows = sorted(((e.key, e.self_device_time_total, e.count) for e in prof.key_averages()), key=lambda r: -r[1])
print("total us", sum(r[1] for r in rows))
for k, us, n in rows[:14]:
    print("%9.1f us %5d x  %s" % (us, n, k[:110]))
