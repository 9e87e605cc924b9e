"""GPU: the classical preprocessor's localisation network on native kernels (tpspp_locnet_fwd) against the
oracle's restatement of tps_preprocessor.py:96-156 (pinned to the reference's LocalizationNetwork by oracle/make_golden.py) in
fp64, and through the drop-in TPSPreprocessor module."""
import numpy as np
import pytest
import torch

import tps_pp_b200 as T
from oracle import tpspp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _trained_like(m, seed):
    """Random but trained-looking localisation weights: BN statistics away from (0, 1), a non-zero fc2 weight."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in list(m.LocalizationNetwork.named_parameters()) + list(m.LocalizationNetwork.named_buffers()):
            if name.endswith("running_mean"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif name.endswith("num_batches_tracked"):
                continue
            elif "conv" in name and name.endswith(".weight") and p.dim() == 1:      # BN gamma (some negative: the fold must not assume s > 0)
                p.copy_(0.5 + torch.rand(p.shape, generator=g) * (torch.rand(p.shape, generator=g) > 0.1).float() * 1.0 - 0.2)
            elif "conv" in name and name.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name == "localization_fc2.weight":
                p.copy_(torch.randn(p.shape, generator=g) * 0.01)
    return m


@pytest.mark.parametrize("batch,chans,h,w,f", [(3, 3, 64, 256, 20), (2, 1, 64, 128, 40), (5, 3, 32, 256, 20)])
def test_locnet_vs_oracle_fp64(native_lib, batch, chans, h, w, f):
    m = _trained_like(T.TPSPreprocessor(num_fiducial=f, img_size=(h, w), rectified_img_size=(h, w), num_img_channel=chans), 11 + f)
    m = m.to(DEV).eval()
    g = torch.Generator().manual_seed(batch * 7 + w)
    img = torch.randn((batch, chans, h, w), generator=g)
    with torch.no_grad():
        cp = m.localize(img.to(DEV))
    assert m._last_locnet_native, "the native localisation network did not run"
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref64 = O.classical_localization({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, img.double())
    ref32 = O.classical_localization(sd, img)
    err = float((cp.cpu().double() - ref64).abs().max())
    floor = float((ref32.double() - ref64).abs().max())
    print(f"locnet {batch}x{chans}x{h}x{w} F={f}: |C' - oracle64| = {err:.3e} (fp32 reference-path floor {floor:.3e})")
    assert err <= max(1e-5, 4 * floor)               # north_star: fp32 control points within 1e-4 of the reference


def test_locnet_module_forward_and_fallbacks(native_lib):
    m = _trained_like(T.TPSPreprocessor(num_fiducial=20, img_size=(64, 256), rectified_img_size=(64, 256), num_img_channel=3), 5)
    m = m.to(DEV).eval()
    img = torch.randn((4, 3, 64, 256), device=DEV)
    with torch.no_grad():
        out_native = m(img)
        assert m._last_locnet_native
        m.locnet_impl = "library"
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            out_lib = m(img)
        assert not m._last_locnet_native
    assert float((out_native - out_lib).abs().max()) <= 2e-4     # the rectified image through two fp32-level localisation paths
    m.locnet_impl = "auto"
    out_grad = m(img)                                            # autograd recording: library stack (BatchNorm in eval mode still)
    assert not m._last_locnet_native and out_grad.shape == out_native.shape
    # a geometry the tensor-core tiling does not take (the recogniser configs' 32x100) stays on the library stack
    m2 = T.TPSPreprocessor(num_fiducial=20, img_size=(32, 100), rectified_img_size=(32, 100), num_img_channel=1).to(DEV).eval()
    with torch.no_grad():
        m2(torch.randn((2, 1, 32, 100), device=DEV))
    assert not m2._last_locnet_native


def test_moran_drop_in_vs_reference_golden(native_lib, golden):
    """MORAN (moran.py:13-103) with its two grid_sample calls on the native sampling core: state_dict keys of the reference,
    output against the reference's own output (fixture written by oracle/make_golden.py) and against the fp64 oracle."""
    g = golden("moran.npz")
    m = T.MORAN(num_img_channel=3, img_size=(32, 128), maxBatch=4)
    assert list(m.state_dict().keys()) == [str(k) for k in g["state_keys"]]
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    missing = m.load_state_dict(sd, strict=False)
    assert set(missing.missing_keys) <= {"grid"} | {k for k in m.state_dict() if k.endswith("num_batches_tracked")}
    assert not missing.unexpected_keys
    m = m.to(DEV).eval()
    x = torch.from_numpy(g["x"]).to(DEV)
    with torch.no_grad():
        out0 = m(x)
        assert m._last_sampler_native
        out1 = m(x, enhance=1)
    assert out0.shape == (2, 3, 32, 128)
    # (a) the sampler itself: the same module with ATen's grid_sample (autograd recording) -- identical offsets, so the
    #     difference is the native sampling core against ATen: a few fp32 ulps of the blend
    out_grad = m(x.clone().requires_grad_())
    assert not m._last_sampler_native
    assert float((out_grad - out0).abs().max()) <= 5e-6
    # (b) the whole module against the reference's own output and the fp64 oracle: the offset CNN is a cuDNN fp32 stack whose
    #     ~1e-6 rounding differences from the reference's CPU convolutions are multiplied by the image gradient along y
    #     (an N(0,1) noise image: O(1) per source pixel x 23.5 source pixels per unit offset)
    assert float((out0.cpu() - torch.from_numpy(g["ref32_output"])).abs().max()) <= 2e-4
    assert float((out1.cpu() - torch.from_numpy(g["ref32_output_enhance1"])).abs().max()) <= 4e-4
    ref64 = O.moran_forward({k: v.double() for k, v in sd.items()}, torch.from_numpy(g["x"]).double(), (32, 128))
    assert float((out0.cpu().double() - ref64).abs().max()) <= 2e-4
    with pytest.raises(RuntimeError):
        m.cpu()(torch.zeros(1, 3, 32, 128))
