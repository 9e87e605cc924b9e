"""GPU: the training-path convolution (tpspp_conv_fwd / tpspp_conv_bwd through ``functional.conv_relu``) against torch autograd
of ``relu(conv2d(.))`` in fp64 on the device -- forward, data gradient (the forward tcgen05 engine run over the zero-inserted
output gradient with the transposed / mirrored weight image), weight gradient and bias gradient, for every geometry the
head's 14 ConvModules use (reference tps_pp.py:126-131,149-154,538-548)."""
import pytest
import torch
import torch.nn.functional as F

from tps_pp_b200 import functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (cin, h, w, k, stride): down0/1, down2, down_feat, down0_1/1_1, enc0, enc1, enc2, enc3, dec0, dec1, dec2/3
GEOMS = [(32, 32, 128, 1, 1), (64, 16, 64, 1, 1), (192, 32, 128, 1, 1), (64, 32, 128, 3, 2), (192, 16, 64, 3, 1),
         (64, 16, 64, 3, 2), (64, 8, 32, 3, 2), (64, 4, 16, 3, (2, 1)), (64, 4, 16, 3, 1), (64, 8, 32, 3, 1), (64, 16, 64, 3, 1)]


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("cin,h,w,k,stride", GEOMS)
@pytest.mark.parametrize("relu", [True, False])
def test_conv_relu_forward_and_gradients(native_lib, cin, h, w, k, stride, relu):
    if not relu and (k, stride) != (3, 1):
        pytest.skip("the linear variant is covered on one geometry per kernel size")
    g = torch.Generator(device=DEV).manual_seed(1000 + cin + h + k)
    b = 8
    x = torch.randn((b, cin, h, w), device=DEV, generator=g)
    wt = torch.randn((64, cin, k, k), device=DEV, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn(64, device=DEV, generator=g) * 0.1
    xs, ws, bs = (t.clone().requires_grad_() for t in (x, wt, bias))
    y = TF.conv_relu(xs, ws, bs, stride, relu)
    gy = torch.randn(y.shape, device=DEV, generator=g)
    y.backward(gy)
    xd, wd, bd = (t.double().requires_grad_() for t in (x, wt, bias))
    yd = F.conv2d(xd, wd, bd, stride=stride, padding=k // 2)
    if relu:
        # the ReLU mask is taken from OUR forward so that pixels within rounding of zero do not decide the comparison
        yd = yd * (y.detach() > 0).double()
    yd.backward(gy.double())
    assert rel(y.detach(), yd.detach()) < 2e-5
    assert rel(xs.grad, xd.grad) < 1e-4
    assert rel(ws.grad, wd.grad) < 1e-4
    assert rel(bs.grad, bd.grad) < 1e-4


def test_conv_relu_partial_grads_and_errors(native_lib):
    x = torch.randn((4, 64, 8, 32), device=DEV)
    w = torch.randn((64, 64, 3, 3), device=DEV, requires_grad=True)
    b = torch.zeros(64, device=DEV)
    y = TF.conv_relu(x, w, b, 1)             # only the weight needs a gradient: no data-gradient launches
    y.sum().backward()
    assert w.grad is not None and torch.isfinite(w.grad).all()
    with pytest.raises(RuntimeError):
        TF.conv_relu(torch.randn((4, 48, 8, 32), device=DEV), torch.randn((64, 48, 3, 3), device=DEV), b, 1)
    with pytest.raises(RuntimeError):
        TF.conv_relu(x.cpu(), w.detach().cpu(), b.cpu(), 1)


# fused torch.cat / F.interpolate sources (tps_pp.py:159-168 decoder upsampling, :583-585 down_feat, :560-562 feat_cat):
# (per-source channel shape, ups, h, w of the conv input, k)
FUSED = [
    ([64, 64, 64], [1, 1, 1], 16, 64, 3, "enc0: cat(down0_1, down1_1, f2)"),
    ([64, 64, 64], [1, 1, 2], 32, 128, 1, "down_feat: cat(f0, f1, up2(f2))"),
    ([64], [(2, 1)], 4, 16, 3, "dec0: up(2,1)"),
    ([64], [2], 8, 32, 3, "dec1: up2"),
    ([64], [2], 16, 64, 3, "dec2: up2"),
]


@pytest.mark.parametrize("chans,ups,h,w,k,what", FUSED, ids=[f[5] for f in FUSED])
def test_conv_relu_fused_cat_and_upsample(native_lib, chans, ups, h, w, k, what):
    g = torch.Generator(device=DEV).manual_seed(77 + h + k)
    b = 8
    nups = [(u, u) if isinstance(u, int) else u for u in ups]
    srcs = [torch.randn((b, c, h // u[0], w // u[1]), device=DEV, generator=g) for c, u in zip(chans, nups)]
    cin = sum(chans)
    wt = torch.randn((64, cin, k, k), device=DEV, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn(64, device=DEV, generator=g) * 0.1
    xs = [t.clone().requires_grad_() for t in srcs]
    ws, bs = wt.clone().requires_grad_(), bias.clone().requires_grad_()
    assert TF.conv_relu_supported(xs, ws, 1, ups)
    y = TF.conv_relu(xs, ws, bs, 1, True, ups)
    gy = torch.randn(y.shape, device=DEV, generator=g)
    y.backward(gy)
    xd = [t.double().requires_grad_() for t in srcs]
    wd, bd = wt.double().requires_grad_(), bias.double().requires_grad_()
    full = torch.cat([F.interpolate(t, scale_factor=tuple(float(v) for v in u), mode="nearest") if u != (1, 1) else t
                      for t, u in zip(xd, nups)], dim=1)
    yd = F.conv2d(full, wd, bd, padding=k // 2) * (y.detach() > 0).double()
    yd.backward(gy.double())
    assert rel(y.detach(), yd.detach()) < 2e-5
    for a, r in zip(xs, xd):
        assert a.grad.shape == r.grad.shape and rel(a.grad, r.grad) < 1e-4, what
    assert rel(ws.grad, wd.grad) < 1e-4
    assert rel(bs.grad, bd.grad) < 1e-4
