"""GPU: the native backbone stage in front of the rectifier (tpspp_stage_fwd: stem + layer1 + layer2)
against the oracle, and the image-to-rectified-features chain through the drop-in ``ResNetABI_v2_large`` + ``TPS_PP``.

Acceptance rule as for the head: as close to the fp64 twin as the reference's own fp32 path is,
err(ours, ref64) <= 4 * err(ref32, ref64) + 5e-5 * scale (tensor-core accumulators truncate)."""
import numpy as np
import pytest
import torch

import tps_pp_b200 as T
from oracle import tpspp_oracle as O
from tps_pp_b200 import _native as N
from tps_pp_b200 import functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def mx(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))


def _backbone(seed=5):
    m = T.ResNetABI_v2_large(arch_settings=[3, 4, 6, 6, 3], strides=[1, 2, 2, 1, 2]).to(DEV).eval()
    sd = O.trained_like_backbone_state(seed)
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    return m, sd


@pytest.mark.parametrize("batch,seed", [(2, 1234), (1, 7), (5, 11), (19, 3)])
def test_stage_vs_oracle(native_lib, batch, seed):
    m, sd = _backbone()
    img = O.synthetic_images(batch, seed)
    with torch.no_grad():
        o0, o1, x, _ = TF.stage_forward(torch.from_numpy(img).to(DEV), m.stage_tensors())
    torch.cuda.synchronize()
    x32, outs32 = O.backbone_stage_forward(sd, img, torch.float32)
    x64, outs64 = O.backbone_stage_forward(sd, img, torch.float64)
    report = []
    for name, got, r32, r64 in (("o0", o0, outs32[0], outs64[0]), ("o1", o1, outs32[1], outs64[1]), ("x", x, x32, x64)):
        scale = float(r64.abs().max())
        floor = mx(r32, r64)
        err = mx(got, r64)
        report.append(f"{name}: |ours-ref64|={err:.2e} |ref32-ref64|={floor:.2e} scale={scale:.2f}")
        assert err <= 4 * floor + 5e-5 * max(scale, 1.0), "\n".join(report)
    print("\n".join(report))


def test_stage_matches_committed_reference_subsample(native_lib, golden):
    g = golden("backbone_stage.npz")
    m, _ = _backbone()
    img = O.synthetic_images(int(g["batch"]))
    with torch.no_grad():
        o0, o1, x, _ = TF.stage_forward(torch.from_numpy(img).to(DEV), m.stage_tensors())
    s = int(g["stride"])
    for name, got in (("x", x), ("o0", o0), ("o1", o1)):
        ref = g[name]
        err = mx(got.cpu().numpy().reshape(-1)[::s], ref)
        assert err <= 1e-4 * max(1.0, float(np.abs(ref).max())), (name, err)


def test_module_native_stage_equals_library_stage(native_lib):
    """The drop-in backbone: native stage (no_grad, eval) against its own library-op stage, and the reference's
    forward contract -- ``tpsnet(x, outs)`` in front of layer3, ``output`` replaces ``x``, dict(output, img_ref)."""
    m, _ = _backbone()
    img = torch.from_numpy(O.synthetic_images(3, 21)).to(DEV)
    seen = {}

    def tpsnet(x, outs, **kw):
        seen.update(x=x, outs=list(outs))
        return {"output": x * 0.5}
    with torch.no_grad():
        r_nat = m(img, tpsnet, True)
        assert m._last_stage_native
        nat = dict(x=seen["x"].clone(), o0=seen["outs"][0].clone(), o1=seen["outs"][1].clone())
        m.stage_impl = "library"
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            r_lib = m(img, tpsnet, True)
        assert not m._last_stage_native
    for k, ref in (("x", seen["x"]), ("o0", seen["outs"][0]), ("o1", seen["outs"][1])):
        assert mx(nat[k], ref) <= 5e-5 * max(1.0, float(ref.abs().max())), k
    assert set(r_nat) == {"output", "img_ref"} and r_nat["output"].shape == r_lib["output"].shape == (3, 512, 4, 16)
    assert mx(r_nat["img_ref"], r_lib["img_ref"]) <= 1e-3
    m.stage_impl = "native"
    with pytest.raises(RuntimeError):
        m(img, tpsnet, True)             # autograd recording: the native stage has no backward


def test_image_to_rectified_features_vs_oracle(native_lib):
    """img -> native stage -> native TPS_PP against the oracle chain (reference resnet_v2_large.py:176-191 + tps_pp.py:564-625)."""
    m, sd_b = _backbone()
    tps = T.TPS_PP().to(DEV).eval()
    sd_t = O.trained_like_state(3)
    tps.load_state_dict(sd_t, strict=True)
    img = O.synthetic_images(4, 99)
    with torch.no_grad():
        r = m(torch.from_numpy(img).to(DEV), tps, True)
    assert m._last_stage_native and all(tps.native_stages.values())
    x32, outs32 = O.backbone_stage_forward(sd_b, img, torch.float32)
    x64, outs64 = O.backbone_stage_forward(sd_b, img, torch.float64)
    r32 = O.tps_pp_forward(sd_t, x32.numpy(), [o.numpy() for o in outs32], dtype=torch.float32)
    r64 = O.tps_pp_forward(sd_t, x64.numpy(), [o.numpy() for o in outs64], dtype=torch.float64)
    floor = mx(r32["output"], r64["output"])
    err = mx(r["img_ref"], r64["output"])
    print(f"img -> output: |ours - ref64| = {err:.3e}, reference's own |ref32 - ref64| = {floor:.3e}")
    assert err <= max(1e-5, 1.5 * floor)


def test_pipelined_streams_match_serial(native_lib):
    """The end-to-end pattern of bench.py: uploads on a copy stream into two device buffer sets, compute on the main stream behind
    an event, downloads on a third stream.  The chain's kernels are launched with programmatic stream serialization (next grid's
    prologue overlaps the previous grid's drain, csrc/common.cuh); cross-stream event dependencies must stay full dependencies --
    every pipelined step has to reproduce the serial result of ITS OWN input bit for bit."""
    m, _ = _backbone()
    tps = T.TPS_PP().to(DEV).eval()
    tps.load_state_dict(O.trained_like_state(3), strict=True)
    steps, batch = 8, 24
    host_in = [torch.from_numpy(O.synthetic_images(batch, 500 + i)).pin_memory() for i in range(steps)]
    with torch.no_grad():
        serial = []
        for h in host_in:
            xx, outs = m.stage(h.to(DEV))
            serial.append(tps(xx, outs)["output"].cpu())
            torch.cuda.synchronize()
        dev = torch.device(DEV)
        dbuf = [torch.empty((batch, 3, 32, 128), device=dev) for _ in range(2)]
        s_in, s_out, main = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.current_stream(dev)
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        host_out = [torch.empty_like(serial[0]).pin_memory() for _ in range(steps)]
        for k in range(2):
            ev_free[k].record(main)

        def upload(i):
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_free[i & 1])
                dbuf[i & 1].copy_(host_in[i], non_blocking=True)
                ev_in[i & 1].record(s_in)

        upload(0)
        for i in range(steps):
            k = i & 1
            if i + 1 < steps:
                upload(i + 1)
            main.wait_event(ev_in[k])
            xx, outs = m.stage(dbuf[k])
            out = tps(xx, outs)["output"]
            done = torch.cuda.Event()
            done.record(main)
            ev_free[k].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                host_out[i].copy_(out, non_blocking=True)
            out.record_stream(s_out)
        torch.cuda.synchronize()
    for i in range(steps):
        assert torch.equal(host_out[i], serial[i]), f"pipelined step {i} differs from its serial result"
