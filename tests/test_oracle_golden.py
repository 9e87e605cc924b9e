"""CPU: the oracle restatement against the committed reference vectors (tests/golden, produced by
oracle/make_golden.py from the unmodified reference in the build container)."""
import numpy as np
import pytest
import torch

from oracle import tpspp_oracle as O


def mx(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def test_constants_bit_exact(golden):
    g = golden("constants.npz")
    c = O.tpspp_constants()
    assert np.array_equal(c["hat_C"], g["tpspp_hat_C"])
    assert np.array_equal(c["P_hat"], g["tpspp_P_hat"])
    assert np.array_equal(c["P"], g["tpspp_P"])
    for f, rs in ((20, (32, 100)), (6, (8, 12))):
        cc = O.classical_constants(f, rs)
        assert np.array_equal(cc["inv_delta_C"], g[f"classical_F{f}_{rs[0]}x{rs[1]}_inv_delta_C"])
        assert np.array_equal(cc["P_hat"], g[f"classical_F{f}_{rs[0]}x{rs[1]}_P_hat"])


def test_warp_grid_and_sampler_vs_reference(golden):
    g = golden("warp_tpspp.npz")
    c = O.tpspp_constants()
    g64 = O.tpspp_grid(g["c_prime"], g["pc_score"], c["hat_C"], c["P"], c["P_hat"], dtype=np.float64)
    assert mx(g64, g["ref_grid64"]) < 1e-12
    assert mx(g64, g["ref_grid32"]) < 5e-5          # the reference's own fp32 noise floor (F6)
    b = g64.shape[0]
    out = O.grid_sample(g["fg_ch"], g["ref_grid32"].reshape(b, 16, 64, 2), dtype=np.float32)
    assert mx(out, g["ref_out_ch"]) < 2e-6
    mp = O.grid_sample(g["x_ch"], g["ref_grid32"].reshape(b, 16, 64, 2), dtype=np.float32)
    assert mx(mp, g["ref_mp_ch"]) < 2e-6
    out64 = O.grid_sample(g["fg_ch"], g["ref_grid64"].reshape(b, 16, 64, 2), dtype=np.float64)
    assert mx(out64, g["ref_out64_ch"]) < 1e-12


def test_classical_warp_vs_reference(golden):
    g = golden("warp_classical.npz")
    for f, rs in ((20, (32, 100)), (6, (8, 12))):
        cc = O.classical_constants(f, rs)
        out, grid = O.classical_warp(g[f"F{f}_img"], g[f"F{f}_c_prime"], cc, rs, dtype=np.float64)
        assert mx(grid, g[f"F{f}_ref_grid32"]) < 2e-5
        o32 = O.grid_sample(g[f"F{f}_img"], g[f"F{f}_ref_grid32"].reshape(-1, rs[0], rs[1], 2), dtype=np.float32)
        assert mx(o32, g[f"F{f}_ref_out"]) < 2e-6


def test_full_forward_vs_reference(golden):
    g = golden("tpspp_forward.npz")
    sd = O.trained_like_state(int(g["state_seed"]))
    assert np.allclose([float(v.double().abs().sum()) for v in sd.values()], g["state_digest"], rtol=0, atol=0)
    x, o0, o1 = O.synthetic_tpspp_inputs(int(g["batch"]), int(g["input_seed"]))
    r32 = O.tps_pp_forward(sd, x, [o0, o1], dtype=torch.float32)
    r64 = O.tps_pp_forward(sd, x, [o0, o1], dtype=torch.float64)
    assert mx(r32["control_point"], g["ref32_control_point"]) < 1e-6
    assert mx(r64["control_point"], g["ref64_control_point"]) < 1e-12
    assert mx(r32["pc_score"], g["ref32_pc_score"]) < 1e-5
    assert mx(r64["grid"], g["ref64_grid"]) < 1e-10
    assert mx(r64["output"], g["ref64_output"]) < 1e-6       # fixture stored as fp32
    assert mx(r64["mp_img"], g["ref64_mp_img"]) < 1e-6
    # fp32 vs fp32: only as close as the reference is to its own fp64 twin
    floor_out = mx(g["ref32_output"], g["ref64_output"])
    assert mx(r32["output"], g["ref32_output"]) < 4 * floor_out + 1e-5
    assert mx(r32["en_feat"], g["ref32_en_feat"]) < 1e-5


def test_backward_restatement_matches_autograd():
    """fp64: oracle backward formulas == torch autograd through bmm + grid_sample."""
    import torch.nn.functional as F
    rs = np.random.RandomState(5)
    c = O.tpspp_constants()
    B = 2
    cp = O.smooth_c_prime(O.tpspp_init_bias(), B, seed=9)
    s = np.tanh(rs.standard_normal((B, 1024, 32))).astype(np.float32)
    src = rs.standard_normal((B, 3, 32, 128)).astype(np.float32)
    gout = rs.standard_normal((B, 3, 16, 64))
    cpt = torch.from_numpy(cp).double().requires_grad_(True)
    st = torch.from_numpy(s).double().requires_grad_(True)
    srct = torch.from_numpy(src).double().requires_grad_(True)
    hat = torch.from_numpy(c["hat_C"]).double(); ph = torch.from_numpy(c["P_hat"]).double()
    P = torch.from_numpy(c["P"]).double()
    phi = torch.cat([torch.ones(B, 1024, 1, dtype=torch.float64), P[None].repeat(B, 1, 1), ph[None] * (st * 0.5 + 1)], 2)
    T = torch.bmm(hat[None].repeat(B, 1, 1), torch.cat([cpt, torch.zeros(B, 3, 2, dtype=torch.float64)], 1))
    grid = torch.bmm(phi, T).reshape(B, 16, 64, 2)
    out = F.grid_sample(srct, grid, padding_mode="border", align_corners=True)
    (out * torch.from_numpy(gout)).sum().backward()
    gsrc, gg = O.grid_sample_backward(src, grid.detach().numpy(), gout, dtype=np.float64)
    dC, ds = O.tpspp_grid_backward(gg.reshape(B, 1024, 2), cp, s, c["hat_C"], c["P"], c["P_hat"])
    assert mx(gsrc, srct.grad) < 1e-10
    assert mx(dC, cpt.grad) / float(cpt.grad.abs().max()) < 1e-10
    assert mx(ds, st.grad) / float(st.grad.abs().max()) < 1e-10


def test_sampler_restatement_vs_aten_on_random_and_edge_grids():
    """The numpy sampler against torch's own grid_sample (the third-party code the reference calls at tps_pp.py:606-615),
    including coordinates far outside the image, exactly on the border, and the (0,1)-coordinates-into-a-[-1,1]-sampler
    quirk of TPS_PP.  Runs anywhere torch does -- no reference tree needed."""
    rs = np.random.RandomState(0)
    for (b, c, h, w, hr, wr) in ((2, 3, 5, 7, 4, 6), (1, 2, 16, 64, 16, 64), (3, 1, 1, 9, 2, 3)):
        src = rs.standard_normal((b, c, h, w))
        grid = rs.uniform(-1.6, 1.6, size=(b, hr, wr, 2))
        grid[0, 0, 0] = (-1.0, -1.0); grid[0, 0, 1] = (1.0, 1.0); grid[0, 1, 0] = (7.5, -9.0); grid[0, 1, 1] = (0.0, 0.0)
        for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
            ours = O.grid_sample(src, grid, dtype=dt)
            ref = torch.nn.functional.grid_sample(torch.from_numpy(src.astype(dt)), torch.from_numpy(grid.astype(dt)),
                                                  mode="bilinear", padding_mode="border", align_corners=True).numpy()
            assert mx(ours, ref) <= tol, (b, c, h, w, dt)
        # backward restatement against autograd in fp64
        ts = torch.from_numpy(src).requires_grad_()
        tg = torch.from_numpy(grid).requires_grad_()
        out = torch.nn.functional.grid_sample(ts, tg, mode="bilinear", padding_mode="border", align_corners=True)
        go = rs.standard_normal(out.shape)
        out.backward(torch.from_numpy(go))
        gs, gg = O.grid_sample_backward(src, grid, go, dtype=np.float64)
        assert mx(gs, ts.grad.numpy()) <= 1e-12 and mx(gg, tg.grad.numpy()) <= 1e-12


def test_nrtr_fixture_weights_are_reproducible(golden):
    """tests/golden/nrtr_argmax.npz carries no rectifier weights: "stock" = the drop-in module under seed 0
    (bit-equal to the reference ctor, mmcv ConvModule kaiming init included), "trained" = trained_like_state(3)."""
    import torch
    import tps_pp_b200 as T
    g = golden("nrtr_argmax.npz")
    torch.manual_seed(0)
    m = T.TPS_PP()
    digest = np.array([float(v.double().abs().sum()) for v in m.state_dict().values()])
    assert np.allclose(digest, g["stock_state_digest"], rtol=1e-6, atol=0)
    digest = np.array([float(v.double().abs().sum()) for v in O.trained_like_state(3).values()])
    assert np.allclose(digest, g["trained_state_digest"], rtol=1e-6, atol=0)
    assert g["x"].shape == (2, 64, 16, 64) and g["o0"].shape == (2, 32, 32, 128)
    assert float(g["stock_safe_delta"]) >= 1e-3 and float(g["trained_safe_delta"]) >= 1e-3 and int(g["stock_flips"][0]) == 0 and int(g["trained_flips"][0]) == 0


def test_backbone_stage_vs_reference(golden):
    """the oracle's stem + layer1 + layer2 against what the unmodified reference backbone handed to
    ``tpsnet(x, outs)`` (strided subsample of the full tensors, oracle/make_golden.py::backbone_fixture)."""
    g = golden("backbone_stage.npz")
    sd = O.trained_like_backbone_state(5)
    digest = [float(v.double().abs().sum()) for v in sd.values()]
    assert np.allclose(digest, g["state_digest"], rtol=0, atol=0)
    img = O.synthetic_images(int(g["batch"]))
    x, outs = O.backbone_stage_forward(sd, img, torch.float32)
    s = int(g["stride"])
    assert mx(x.numpy().reshape(-1)[::s], g["x"]) < 2e-5
    assert mx(outs[0].numpy().reshape(-1)[::s], g["o0"]) < 2e-5
    assert mx(outs[1].numpy().reshape(-1)[::s], g["o1"]) < 2e-5
    assert tuple(x.shape) == (2, 64, 16, 64) and tuple(outs[0].shape) == tuple(outs[1].shape) == (2, 32, 32, 128)


def test_moran_restatement_vs_reference(golden):
    """oracle.moran_forward (moran.py:66-103) against the unmodified reference's output on the committed fixture."""
    g = golden("moran.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    x = torch.from_numpy(g["x"])
    for enh, key in ((0, "ref32_output"), (1, "ref32_output_enhance1")):
        out = O.moran_forward(sd, x, (32, 128), enhance=enh)
        assert float((out - torch.from_numpy(g[key])).abs().max()) <= 1e-5


def test_nrtr_decoder_restatement_and_dropin_init(golden):
    """oracle.nrtr_forward_test (nrtr_decoder.py:153-177) against the reference's own greedy decode on the committed
    reduced-configuration fixture; the drop-in NRTRDecoder reproduces the reference's keys and seeded initial weights."""
    import tps_pp_b200 as T
    g = golden("nrtr_decoder.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    probs = O.nrtr_forward_test(sd, torch.from_numpy(g["out_enc"]), [float(v) for v in g["valid_ratios"]], n_head=2,
                                max_seq_len=12, start_idx=1, padding_idx=36)
    assert float((probs - torch.from_numpy(g["ref32_probs"])).abs().max()) <= 1e-6
    torch.manual_seed(0)
    m = T.NRTRDecoder()
    assert list(m.state_dict().keys()) == [str(k) for k in g["full_state_keys"]]
    digest = np.array([float(v.double().abs().sum()) for v in m.state_dict().values()])
    assert np.array_equal(digest, g["full_init_digest"])
    small = T.NRTRDecoder(n_layers=2, d_embedding=128, n_head=2, d_model=128, d_inner=64, n_position=64, num_classes=37,
                          max_seq_len=12, start_idx=1, padding_idx=36)
    small.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError):
        small.forward_test(None, torch.from_numpy(g["out_enc"]), None)        # CUDA only: no CPU fallback
