import torch, sys
sys.path.insert(0, '.')
import tps_pp_b200 as T
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
tp = T.TPSPreprocessor(num_fiducial=20, img_size=(64, 256), rectified_img_size=(64, 256), num_img_channel=3).to(dev).eval()
img = torch.randn((1024, 3, 64, 256), device=dev)
with torch.no_grad():
    for _ in range(3): tp(img)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): tp(img)
        torch.cuda.synchronize()
rows = sorted(((e.key, e.self_device_time_total / 3, e.count // 3) for e in prof.key_averages()), key=lambda r: -r[1])
for k, us, n in rows[:12]:
    print("%9.1f us %4d x  %s" % (us, n, k[:120]))
