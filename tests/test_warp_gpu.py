"""GPU parity tests of the fused warp kernels, called through the C ABI (ctypes -> libtpspp.so).

Tiers (DESIGN.md "Parity"):
  1. sampler, given the reference's own fp32 grid: == ATen grid_sample to <= 1e-6 (fp32 coordinate math)
  2. grid: ours vs the fp64 twin of the reference <= 1e-6 normalised; vs the reference's fp32 grid <= 5e-5
  3. pixels: ours vs fp64 twin <= 1e-5 absolute (tolerance from BASELINE.json north_star)
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import tpspp_oracle as O
from tps_pp_b200 import _native as N
from tps_pp_b200 import functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PIX_TOL = 1e-5      # north_star: rectified pixels within 1e-5 absolute
GRID_TOL = 1e-6


def mx(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device=DEV, dtype=dtype)


@pytest.fixture(scope="module")
def consts(native_lib):
    c = O.tpspp_constants()
    return c, cu(c["hat_C"]), cu(c["P_hat"]), cu(c["P"])


def _inputs(B, seed, C=64, amp=0.05):
    rs = np.random.RandomState(seed)
    cp = O.smooth_c_prime(O.tpspp_init_bias(), B, seed=seed + 1, amp=amp)
    s = np.tanh(0.5 * rs.standard_normal((B, 1024, 32))).astype(np.float32)
    fg = rs.standard_normal((B, C, 32, 128)).astype(np.float32)
    x = rs.standard_normal((B, C, 16, 64)).astype(np.float32)
    return cp, s, fg, x


# ------------------------------------------------------------------ tier 1
def test_sampler_matches_aten_given_reference_grid(golden, native_lib):
    g = golden("warp_tpspp.npz")
    b = g["ref_grid32"].shape[0]
    grid = cu(g["ref_grid32"].reshape(b, 16, 64, 2))
    out, mp = TF.grid_sample_border(cu(g["fg_ch"]), grid, cu(g["x_ch"]))
    assert mx(out, g["ref_out_ch"]) <= 1e-6
    assert mx(mp, g["ref_mp_ch"]) <= 1e-6
    # against torch's own CUDA kernel on this device: bit-identical arithmetic
    ref = torch.nn.functional.grid_sample(cu(g["fg_ch"]), grid, padding_mode="border", align_corners=True)
    assert mx(out, ref) <= 1e-6
    assert float((out != ref).float().mean()) < 1e-3


def test_sampler_edge_grids(native_lib):
    """out-of-range, exactly-on-border, NaN and inf coordinates behave like ATen's CUDA kernel."""
    src = cu(np.random.RandomState(0).standard_normal((1, 2, 5, 7)))
    vals = [-3.0, -1.0, -0.999999, 0.0, 0.3333, 1.0, 1.000001, 7.5, float("inf"), float("-inf"), float("nan")]
    gx, gy = np.meshgrid(vals, vals)
    grid = cu(np.stack([gx, gy], -1)[None])
    out = TF.grid_sample_border(src, grid)
    ref = torch.nn.functional.grid_sample(src, grid, padding_mode="border", align_corners=True)
    assert torch.isfinite(out).all()
    assert mx(out, ref) <= 1e-6


# ------------------------------------------------------------------ tiers 2/3
@pytest.mark.parametrize("variant", [N.VARIANT_GENERIC, N.VARIANT_STAGED])
def test_fused_warp_vs_oracle(consts, variant):
    c, hat, ph, P = consts
    B = 3
    cp, s, fg, x = _inputs(B, 21, C=16)
    out, mp = TF.tps_warp(cu(fg), cu(x), cu(cp), cu(s), ph, P, hat, (16, 64), variant=variant)
    o64, m64, g64 = O.tpspp_warp(fg, x, cp, s, c, dtype=np.float64)
    assert mx(out, o64) <= PIX_TOL
    assert mx(mp, m64) <= PIX_TOL
    if variant == N.VARIANT_GENERIC:
        grid = TF.tps_grid(cu(cp), cu(s), ph, P, hat, (16, 64))
        assert mx(grid, g64) <= GRID_TOL


@pytest.mark.parametrize("variant", [N.VARIANT_GENERIC, N.VARIANT_STAGED])
def test_fused_warp_vs_reference_golden(golden, consts, variant):
    """Against vectors produced by the unmodified reference (fp32 path and its fp64 twin)."""
    c, hat, ph, P = consts
    g = golden("warp_tpspp.npz")
    out, mp = TF.tps_warp(cu(g["fg_ch"]), cu(g["x_ch"]), cu(g["c_prime"]), cu(g["pc_score"]), ph, P, hat, (16, 64),
                          variant=variant)
    assert mx(out, g["ref_out64_ch"]) <= PIX_TOL
    floor = mx(g["ref_out_ch"], g["ref_out64_ch"])           # the reference's own fp32 error (F6)
    assert mx(out, g["ref_out_ch"]) <= floor + PIX_TOL
    if variant == N.VARIANT_GENERIC:
        grid = TF.tps_grid(cu(g["c_prime"]), cu(g["pc_score"]), ph, P, hat, (16, 64))
        assert mx(grid, g["ref_grid64"]) <= GRID_TOL
        assert mx(grid, g["ref_grid32"]) <= 5e-5


def test_staged_equals_generic_bitwise_full_size(consts):
    """BASELINE config 2 size (B=256, 64 ch): the two kernels share the arithmetic -> identical bits."""
    c, hat, ph, P = consts
    B = 256
    gen = torch.Generator(device=DEV).manual_seed(0)
    fg = torch.randn((B, 64, 32, 128), device=DEV, generator=gen)
    x = torch.randn((B, 64, 16, 64), device=DEV, generator=gen)
    s = torch.tanh(0.5 * torch.randn((B, 1024, 32), device=DEV, generator=gen))
    cp = cu(O.smooth_c_prime(O.tpspp_init_bias(), B, seed=3))
    a0, a1 = TF.tps_warp(fg, x, cp, s, ph, P, hat, (16, 64), variant=N.VARIANT_GENERIC)
    b0, b1 = TF.tps_warp(fg, x, cp, s, ph, P, hat, (16, 64), variant=N.VARIANT_STAGED)
    assert torch.equal(a0, b0) and torch.equal(a1, b1)
    # lean (single source) staged variant
    c0, none = TF.tps_warp(fg, None, cp, s, ph, P, hat, (16, 64), variant=N.VARIANT_STAGED)
    assert none is None and torch.equal(c0, a0)
    # size-independent properties: linearity in the source, identity transform
    d0, _ = TF.tps_warp(2.0 * fg, x, cp, s, ph, P, hat, (16, 64))
    assert torch.equal(d0, 2.0 * a0)
    ident = cu(np.broadcast_to(c["C"].astype(np.float32), (B, 32, 2)))
    e0, e1 = TF.tps_warp(fg, x, ident, torch.zeros_like(s), ph, P, hat, (16, 64))
    gridP = P.view(1, 16, 64, 2).expand(B, -1, -1, -1).contiguous()
    r0 = torch.nn.functional.grid_sample(fg, gridP, padding_mode="border", align_corners=True)
    # identity C' reproduces P to ~1.5e-5 normalised (SURVEY C-3) = 1e-3 source px -> few e-3 on N(0,1) noise
    assert mx(e0, r0) <= 2e-2


@pytest.mark.parametrize("B,C,n_hw", [(1, 1, (16, 64)), (5, 3, (16, 64)), (2, 7, (10, 50)), (300, 2, (16, 64))])
def test_ragged_shapes(native_lib, B, C, n_hw):
    """ragged channel counts / batches that do not divide the CTA partition; non-default rectified size."""
    from tps_pp_b200 import constants as K
    hat, ph, P, _ = K.attention_tps_buffers((2, 16), n_hw)
    n = n_hw[0] * n_hw[1]
    rs = np.random.RandomState(B * 7 + C)
    cp = O.smooth_c_prime(O.tpspp_init_bias(), B, seed=B)
    s = np.tanh(rs.standard_normal((B, n, 32))).astype(np.float32)
    fg = rs.standard_normal((B, C, 20, 44)).astype(np.float32)
    x = rs.standard_normal((B, C, 9, 12)).astype(np.float32)
    consts = dict(hat_C=hat, P_hat=ph, P=P)
    o64, m64, _ = O.tpspp_warp(fg[:4], x[:4], cp[:4], s[:4], consts, dtype=np.float64, rect_size=n_hw)
    for variant in (N.VARIANT_GENERIC, N.VARIANT_STAGED):
        out, mp = TF.tps_warp(cu(fg), cu(x), cu(cp), cu(s), cu(ph), cu(P), cu(hat), n_hw, variant=variant)
        assert mx(out[:4], o64) <= PIX_TOL and mx(mp[:4], m64) <= PIX_TOL


def test_empty_batch_and_bad_args(consts):
    c, hat, ph, P = consts
    e = torch.empty((0, 4, 32, 128), device=DEV)
    out, _ = TF.tps_warp(e, None, torch.empty((0, 32, 2), device=DEV), torch.empty((0, 1024, 32), device=DEV),
                         ph, P, hat, (16, 64))
    assert out.shape == (0, 4, 16, 64)
    with pytest.raises(RuntimeError, match="P_hat"):
        TF.tps_warp(torch.zeros((1, 1, 4, 4), device=DEV), None, torch.zeros((1, 32, 2), device=DEV),
                    torch.zeros((1, 1024, 32), device=DEV), ph[:, :5], P, hat, (16, 64))
    with pytest.raises(RuntimeError, match="staged"):
        TF.tps_warp(torch.zeros((1, 1, 4, 4), device=DEV), None, torch.zeros((1, 20, 2), device=DEV), None,
                    torch.zeros((3200, 23), device=DEV), None, torch.zeros((23, 23), device=DEV), (32, 100),
                    mode=N.MODE_CLASSICAL, variant=N.VARIANT_STAGED)


def test_nonfinite_control_points_do_not_crash(consts):
    c, hat, ph, P = consts
    cp, s, fg, x = _inputs(2, 4, C=2)
    cp[0, 3, 0] = np.nan
    cp[1, 5, 1] = np.inf
    for variant in (N.VARIANT_GENERIC, N.VARIANT_STAGED):
        out, mp = TF.tps_warp(cu(fg), cu(x), cu(cp), cu(s), ph, P, hat, (16, 64), variant=variant)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()      # NaN/inf coordinates clamp to the border like ATen CUDA


# ------------------------------------------------------------------ classical
def test_classical_vs_reference_golden(golden, native_lib):
    g = golden("warp_classical.npz")
    for f, rs in ((20, (32, 100)), (6, (8, 12))):
        cc = O.classical_constants(f, rs)
        img, cp = g[f"F{f}_img"], g[f"F{f}_c_prime"]
        out, _ = TF.tps_warp(cu(img), None, cu(cp), None, cu(cc["P_hat"]), None, cu(cc["inv_delta_C"]), rs,
                             mode=N.MODE_CLASSICAL, theta=0.0)
        o64, g64 = O.classical_warp(img, cp, cc, rs, dtype=np.float64)
        assert mx(out, o64) <= PIX_TOL
        floor = mx(g[f"F{f}_ref_out"], o64)
        assert mx(out, g[f"F{f}_ref_out"]) <= floor + PIX_TOL
        grid = TF.tps_grid(cu(cp), None, cu(cc["P_hat"]), None, cu(cc["inv_delta_C"]), rs, mode=N.MODE_CLASSICAL)
        assert mx(grid, g64) <= GRID_TOL and mx(grid, g[f"F{f}_ref_grid32"]) <= 2e-5


@pytest.mark.parametrize("F_", [20, 40])
def test_classical_highres_properties(native_lib, F_):
    """BASELINE config 4 (64x256, F=20/40): too big for the numpy oracle at B=1024, so check a slice
    against the oracle and the whole batch against size-independent properties."""
    rs_ = (64, 256)
    cc = O.classical_constants(F_, rs_)
    B = 64
    rs = np.random.RandomState(F_)
    cp = O.smooth_c_prime(O.classical_init_bias(F_), B, seed=2, amp=0.1, centre=0.0)
    img = rs.standard_normal((B, 3, 64, 256)).astype(np.float32)
    timg = cu(img)
    args = (cu(cp), None, cu(cc["P_hat"]), None, cu(cc["inv_delta_C"]), rs_)
    out, _ = TF.tps_warp(timg, None, *args, mode=N.MODE_CLASSICAL, theta=0.0)
    o64, _ = O.classical_warp(img[:2], cp[:2], cc, rs_, dtype=np.float64)
    assert mx(out[:2], o64) <= PIX_TOL
    out2, _ = TF.tps_warp(timg * -0.5, None, *args, mode=N.MODE_CLASSICAL, theta=0.0)
    assert torch.equal(out2, out * -0.5)
    const = torch.full_like(timg, 3.25)
    outc, _ = TF.tps_warp(const, None, *args, mode=N.MODE_CLASSICAL, theta=0.0)
    assert mx(outc, 3.25) <= 1e-5        # bilinear weights sum to one
    # the P_hat-stationary kernel against the plain per-pixel kernel, including control points that push the
    # grid far outside the image and non-finite ones (clipped to the border like ATen)
    cp_bad = cp.copy()
    cp_bad[1] *= 7.0
    cp_bad[2, :, 0] += 3.0
    cp_bad[3, 1, 0] = np.nan
    cp_bad[4, 2, 1] = np.inf
    bad = (cu(cp_bad),) + args[1:]
    a, _ = TF.tps_warp(timg, None, *bad, mode=N.MODE_CLASSICAL, theta=0.0, variant=N.VARIANT_TILED)
    assert N.last_launch_count() == 2            # T kernel + P_hat-stationary kernel
    b, _ = TF.tps_warp(timg, None, *bad, mode=N.MODE_CLASSICAL, theta=0.0, variant=N.VARIANT_GENERIC)
    assert N.last_launch_count() == 1
    assert torch.isfinite(a).all() and mx(a, b) <= 2e-6
    t, _ = TF.tps_warp(timg, None, *args, mode=N.MODE_CLASSICAL, theta=0.0, variant=N.VARIANT_TILED)
    assert mx(t[:2], o64) <= PIX_TOL


def test_classical_tiled_auto_dispatch_and_ragged(native_lib):
    """AUTO picks the P_hat-stationary kernel from batch x pixels >= 4 Mi; ragged tile (n % 256 != 0), a
    ragged batch slice and bf16 pixels go through it too."""
    F_, rs_ = 20, (36, 100)                      # 3600 pixels: 14 full tiles + 16 pixels
    cc = O.classical_constants(F_, rs_)
    B = 1201
    cp = O.smooth_c_prime(O.classical_init_bias(F_), B, seed=5, amp=0.08, centre=0.0)
    g = torch.Generator(device=DEV).manual_seed(1)
    timg = torch.randn((B, 1, 40, 120), device=DEV, generator=g)
    args = (cu(cp), None, cu(cc["P_hat"]), None, cu(cc["inv_delta_C"]), rs_)
    a, _ = TF.tps_warp(timg, None, *args, mode=N.MODE_CLASSICAL, theta=0.0)
    assert N.last_launch_count() == 2
    b, _ = TF.tps_warp(timg, None, *args, mode=N.MODE_CLASSICAL, theta=0.0, variant=N.VARIANT_GENERIC)
    assert mx(a, b) <= 2e-6
    idx = [0, 600, B - 1]
    o64, _ = O.classical_warp(timg[idx].cpu().numpy(), cp[idx], cc, rs_, dtype=np.float64)
    assert mx(a[idx], o64) <= PIX_TOL
    ah, _ = TF.tps_warp(timg.bfloat16(), None, *args, mode=N.MODE_CLASSICAL, theta=0.0)
    bh, _ = TF.tps_warp(timg.bfloat16(), None, *args, mode=N.MODE_CLASSICAL, theta=0.0, variant=N.VARIANT_GENERIC)
    assert ah.dtype == torch.bfloat16 and mx(ah.float(), bh.float()) <= 1e-2
    small, _ = TF.tps_warp(timg[:8], None, cu(cp[:8]), *args[1:], mode=N.MODE_CLASSICAL, theta=0.0)
    assert N.last_launch_count() == 1            # small batch: per-pixel kernel
    assert torch.equal(small, b[:8])


# ------------------------------------------------------------------ backward
def _oracle_bwd(fg, x, cp, s, c, go0, go1):
    o, m, g = O.tpspp_warp(fg, x, cp, s, c, dtype=np.float64)
    B = g.shape[0]
    g4 = g.reshape(B, 16, 64, 2)
    gs0, gg0 = O.grid_sample_backward(fg, g4, go0, dtype=np.float64)
    gs1, gg1 = O.grid_sample_backward(x, g4, go1, dtype=np.float64)
    dC, ds = O.tpspp_grid_backward((gg0 + gg1).reshape(B, 1024, 2), cp, s, c["hat_C"], c["P"], c["P_hat"])
    return gs0, gs1, dC, ds


def test_backward_vs_oracle(consts):
    c, hat, ph, P = consts
    B = 2
    cp, s, fg, x = _inputs(B, 33, C=5)
    rs = np.random.RandomState(1)
    go0 = rs.standard_normal((B, 5, 16, 64)).astype(np.float32)
    go1 = rs.standard_normal((B, 5, 16, 64)).astype(np.float32)
    tfg, tx, tcp, ts = (cu(fg).requires_grad_(), cu(x).requires_grad_(), cu(cp).requires_grad_(), cu(s).requires_grad_())
    out, mp = TF.tps_warp(tfg, tx, tcp, ts, ph, P, hat, (16, 64))
    (out * cu(go0)).sum().backward(retain_graph=True)
    g_only0 = [t.grad.clone() for t in (tfg, tcp, ts)]
    for t in (tfg, tx, tcp, ts):
        t.grad = None
    ((out * cu(go0)).sum() + (mp * cu(go1)).sum()).backward()
    gs0, gs1, dC, ds = _oracle_bwd(fg, x, cp, s, c, go0, go1)
    assert mx(tfg.grad, gs0) <= 1e-5
    assert mx(tx.grad, gs1) <= 1e-5
    # gradients through the solve are large (SURVEY A-3): relative tolerance
    assert mx(tcp.grad, dC) <= 2e-4 * float(np.abs(dC).max())
    assert mx(ts.grad, ds) <= 2e-4 * float(np.abs(ds).max())
    # single-output graph (mp_img unused downstream, as in the recogniser): no gout1
    gs0b, _, dCb, dsb = _oracle_bwd(fg, x, cp, s, c, go0, np.zeros_like(go1))
    assert mx(g_only0[0], gs0b) <= 1e-5
    assert mx(g_only0[1], dCb) <= 2e-4 * float(np.abs(dCb).max())
    assert mx(g_only0[2], dsb) <= 2e-4 * float(np.abs(dsb).max())


def test_backward_classical_vs_oracle(native_lib):
    f, rs_ = 20, (32, 100)
    cc = O.classical_constants(f, rs_)
    B = 2
    rs = np.random.RandomState(8)
    cp = O.smooth_c_prime(O.classical_init_bias(f), B, seed=5, amp=0.1, centre=0.0)
    img = rs.standard_normal((B, 3, 32, 100)).astype(np.float32)
    go = rs.standard_normal((B, 3, 32, 100)).astype(np.float32)
    timg, tcp = cu(img).requires_grad_(), cu(cp).requires_grad_()
    out, _ = TF.tps_warp(timg, None, tcp, None, cu(cc["P_hat"]), None, cu(cc["inv_delta_C"]), rs_,
                         mode=N.MODE_CLASSICAL, theta=0.0)
    (out * cu(go)).sum().backward()
    o64, g64 = O.classical_warp(img, cp, cc, rs_, dtype=np.float64)
    gs, gg = O.grid_sample_backward(img, g64.reshape(B, 32, 100, 2), go, dtype=np.float64)
    dC = O.classical_grid_backward(gg.reshape(B, -1, 2), cc["inv_delta_C"], cc["P_hat"], f)
    assert mx(timg.grad, gs) <= 1e-5
    assert mx(tcp.grad, dC) <= 2e-4 * float(np.abs(dC).max())


def test_backward_matches_torch_autograd_on_device(consts):
    """Same check against the ops the reference itself calls (bmm + grid_sample), in fp64 on the GPU."""
    c, hat, ph, P = consts
    B = 4
    cp, s, fg, x = _inputs(B, 40, C=8)
    tfg, tcp, ts = cu(fg).requires_grad_(), cu(cp).requires_grad_(), cu(s).requires_grad_()
    out, _ = TF.tps_warp(tfg, None, tcp, ts, ph, P, hat, (16, 64))
    w = torch.randn_like(out)
    (out * w).sum().backward()
    d = torch.float64
    rfg, rcp, rs_ = (cu(fg, d).requires_grad_(), cu(cp, d).requires_grad_(), cu(s, d).requires_grad_())
    phi = torch.cat([torch.ones(B, 1024, 1, dtype=d, device=DEV), P.double()[None].repeat(B, 1, 1),
                     ph.double()[None] * (rs_ * 0.5 + 1)], 2)
    T = torch.bmm(hat.double()[None].repeat(B, 1, 1), torch.cat([rcp, torch.zeros(B, 3, 2, dtype=d, device=DEV)], 1))
    grid = torch.bmm(phi, T).reshape(B, 16, 64, 2)
    ref = torch.nn.functional.grid_sample(rfg, grid, padding_mode="border", align_corners=True)
    (ref * w.double()).sum().backward()
    assert mx(out, ref) <= PIX_TOL
    assert mx(tfg.grad, rfg.grad) <= 1e-5
    assert mx(tcp.grad, rcp.grad) <= 2e-4 * float(rcp.grad.abs().max())
    assert mx(ts.grad, rs_.grad) <= 2e-4 * float(rs_.grad.abs().max())


def test_bf16_features(consts):
    """bf16 feature I/O, fp32 coordinates: error is the bf16 rounding of the 4 taps + output."""
    c, hat, ph, P = consts
    cp, s, fg, x = _inputs(2, 50, C=4)
    fgb, xb = cu(fg, torch.bfloat16), cu(x, torch.bfloat16)
    out, mp = TF.tps_warp(fgb, xb, cu(cp), cu(s), ph, P, hat, (16, 64))
    assert out.dtype == torch.bfloat16
    o64, m64, _ = O.tpspp_warp(fgb.float().cpu().numpy(), xb.float().cpu().numpy(), cp, s, c, dtype=np.float64)
    assert mx(out.float(), o64) <= 2 ** -8 * max(1.0, float(np.abs(o64).max()))   # one bf16 ulp of the output


def test_staged_backward_equals_generic_full_geometry(consts):
    """The shared-memory staged backward (default for the TPS++ geometry) against the per-pixel global-atomics kernel at
    the module's real plane sizes (64 channels, 32x128 + 16x64 sources), B = 7: 448 planes over 148 persistent CTAs, so
    image segments split across CTAs and the accumulation-buffer ring wraps."""
    c, hat, ph, P = consts
    B = 7
    cp, s, fg, x = _inputs(B, 77, C=64)
    rs = np.random.RandomState(2)
    go0 = cu(rs.standard_normal((B, 64, 16, 64)).astype(np.float32))
    go1 = cu(rs.standard_normal((B, 64, 16, 64)).astype(np.float32))
    grads = {}
    for variant in (N.VARIANT_AUTO, N.VARIANT_GENERIC):
        TF.BWD_VARIANT = variant
        try:
            tfg, tx, tcp, ts = (cu(fg).requires_grad_(), cu(x).requires_grad_(), cu(cp).requires_grad_(), cu(s).requires_grad_())
            out, mp = TF.tps_warp(tfg, tx, tcp, ts, ph, P, hat, (16, 64))
            ((out * go0).sum() + (mp * go1).sum()).backward()
            grads[variant] = [t.grad.clone() for t in (tfg, tx, tcp, ts)]
            launches = N.last_launch_count()
        finally:
            TF.BWD_VARIANT = N.VARIANT_AUTO
        if variant == N.VARIANT_AUTO:
            pass
    a, g = grads[N.VARIANT_AUTO], grads[N.VARIANT_GENERIC]
    assert mx(a[0], g[0]) <= 2e-5 and mx(a[1], g[1]) <= 2e-5          # d src: same taps, different summation order
    assert mx(a[2], g[2]) <= 2e-4 * float(g[2].abs().max())
    assert mx(a[3], g[3]) <= 2e-4 * float(g[3].abs().max())
    gs0, gs1, dC, ds = _oracle_bwd(fg[:2], x[:2], cp[:2], s[:2], c, go0[:2].cpu().numpy(), go1[:2].cpu().numpy())
    assert mx(a[0][:2], gs0) <= 2e-5 and mx(a[1][:2], gs1) <= 2e-5
    assert mx(a[2][:2], dC) <= 2e-4 * float(np.abs(dC).max())


def test_src0_bf16_staged_warp(consts):
    """TPSPP_SRC0_BF16 (bf16 feat_grid, fp32 x, fp32 outputs: the hand-over of the head's bf16 mode): the staged kernel on bf16
    planes must equal the fp32 kernel on the same (bf16-rounded) values bit for bit; forward only; staged kernel only."""
    c, hat, ph, P = consts
    B = 37
    gen = torch.Generator(device=DEV).manual_seed(5)
    fg = torch.randn((B, 64, 32, 128), device=DEV, generator=gen).bfloat16()
    x = torch.randn((B, 64, 16, 64), device=DEV, generator=gen)
    s = torch.tanh(0.5 * torch.randn((B, 1024, 32), device=DEV, generator=gen))
    cp = cu(O.smooth_c_prime(O.tpspp_init_bias(), B, seed=3))
    a0, a1 = TF.tps_warp(fg.float(), x, cp, s, ph, P, hat, (16, 64))
    b0, b1 = TF.tps_warp(fg, x, cp, s, ph, P, hat, (16, 64))
    assert b0.dtype == torch.float32 and b1.dtype == torch.float32
    assert torch.equal(a0, b0) and torch.equal(a1, b1)
    with pytest.raises(RuntimeError):                       # no generic / classical form of the mixed layout
        TF.tps_warp(fg, x, cp, s, ph, P, hat, (16, 64), variant=N.VARIANT_GENERIC)
    with pytest.raises(RuntimeError):                       # inference layout: no backward
        TF.tps_warp(fg, x.clone().requires_grad_(), cp, s, ph, P, hat, (16, 64))
