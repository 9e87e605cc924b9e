"""Print |ours-ref64| / |ref32-ref64| for the whole-module parity cases (dev aid; GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import tps_pp_b200 as T
from tps_pp_b200 import _native as N
from oracle import tpspp_oracle as O
DEV = "cuda:0"
def mx(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))
g = np.load("tests/golden/tpspp_forward.npz")
sd = O.trained_like_state(3)
x, o0, o1 = O.synthetic_tpspp_inputs(2, 0)
for prec in (N.HEAD_TC, N.HEAD_FP32):
    m = T.TPS_PP().to(DEV).eval(); m.load_state_dict(sd, strict=True); m.head_precision = prec
    with torch.no_grad():
        r = m(torch.from_numpy(x).to(DEV), [torch.from_numpy(o0).to(DEV), torch.from_numpy(o1).to(DEV)])
    print("golden prec", prec, "out", mx(r["output"], g["ref64_output"]), "floor", mx(g["ref32_output"], g["ref64_output"]),
          "mp", mx(r["mp_img"], g["ref64_mp_img"]), "floor", mx(g["ref32_mp_img"], g["ref64_mp_img"]),
          "vs ref32 out", mx(r["output"], g["ref32_output"]))
for seed, stock in ((5, True), (1, False), (0, False), (7, True)):
    if stock:
        torch.manual_seed(0); m = T.TPS_PP().to(DEV).eval(); sdd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    else:
        m = T.TPS_PP().to(DEV).eval(); m.load_state_dict(sd, strict=True); sdd = sd
    xx, a0, a1 = O.synthetic_tpspp_inputs(2, seed)
    r64 = O.tps_pp_forward(sdd, xx, [a0, a1], dtype=torch.float64, sampler="numpy")
    r32 = O.tps_pp_forward(sdd, xx, [a0, a1], dtype=torch.float32)
    for prec in (N.HEAD_TC, N.HEAD_FP32):
        m.head_precision = prec
        with torch.no_grad():
            r = m(torch.from_numpy(xx).to(DEV), [torch.from_numpy(a0).to(DEV), torch.from_numpy(a1).to(DEV)])
        print("seed", seed, "stock", stock, "prec", prec, "out", mx(r["output"], r64["output"]), "floor", mx(r32["output"], r64["output"]),
              "mp", mx(r["mp_img"], r64["mp_img"]), "floor", mx(r32["mp_img"], r64["mp_img"]), "grid err", mx(r["pc_score"], r64["pc_score"]))
g = np.load("tests/golden/nrtr_argmax.npz")
xg, g0, g1 = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "o0", "o1"))
for w in ("stock", "trained"):
    if w == "stock":
        torch.manual_seed(0); m = T.TPS_PP().to(DEV).eval()
    else:
        m = T.TPS_PP().to(DEV).eval(); m.load_state_dict(sd, strict=True)
    sdd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    r64 = O.tps_pp_forward(sdd, g["x"], [g["o0"], g["o1"]], dtype=torch.float64, sampler="numpy")
    for prec in (N.HEAD_TC, N.HEAD_FP32):
        m.head_precision = prec
        with torch.no_grad():
            r = m(xg, [g0, g1])
        print("nrtr", w, prec, "vs ref32", mx(r["output"], g[f"{w}_ref_output"]), "vs ref64", mx(r["output"], r64["output"]),
              "floor ref32-ref64", mx(g[f"{w}_ref_output"], r64["output"]), "out scale", float(np.abs(g[f"{w}_ref_output"]).max()))
