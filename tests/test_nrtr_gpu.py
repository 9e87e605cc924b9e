"""GPU: the NRTR decoder's greedy decode as a native incremental (KV-cache) decode (tps_pp_b200.NRTRDecoder.forward_test over
tpspp_linear_fwd + tpspp_attn_decode) against the reference's own output (reduced-configuration fixture
written by oracle/make_golden.py from the unmodified reference class) and against the oracle on the full configuration."""
import numpy as np
import pytest
import torch

import tps_pp_b200 as T
from oracle import tpspp_oracle as O
from tps_pp_b200 import functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("b,heads,cap,length", [(3, 2, 12, 5), (7, 8, 40, 40), (260, 8, 64, 33), (2, 8, 256, 256)])
def test_attn_decode_vs_torch(native_lib, b, heads, cap, length):
    g = torch.Generator(device=DEV).manual_seed(b * 100 + cap)
    d = heads * 64
    q = torch.randn((b, d), device=DEV, generator=g)
    k = torch.randn((b, cap, d), device=DEV, generator=g)
    v = torch.randn((b, cap, d), device=DEV, generator=g)
    lens = torch.randint(1, length + 1, (b,), device=DEV, generator=g).int()
    for use_lens in (False, True):
        out = TF.attn_decode(q, k, v, heads, length, 8.0, kv_lens=lens if use_lens else None)
        qq = q.double().view(b, heads, 1, 64) / 8.0
        kk = k.double().view(b, cap, heads, 64).transpose(1, 2)
        vv = v.double().view(b, cap, heads, 64).transpose(1, 2)
        a = torch.matmul(qq, kk.transpose(2, 3))
        t = torch.arange(cap, device=DEV)[None, None, None, :]
        limit = lens.view(b, 1, 1, 1) if use_lens else length
        a = a.masked_fill(t >= limit, float("-inf"))
        ref = torch.matmul(torch.softmax(a, -1), vv).transpose(1, 2).reshape(b, d)
        assert float((out.double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max().clamp_min(1.0))
    # append form: q / k_new / v_new as column slices of one fused projection; the kernel writes the step's rows into the cache
    fused = torch.randn((b, 3 * d), device=DEV, generator=g)
    kc, vc = k.clone(), v.clone()
    out = TF.attn_decode(fused[:, :d], kc, vc, heads, length, 8.0, k_new=fused[:, d:2 * d], v_new=fused[:, 2 * d:])
    k2, v2 = k.clone(), v.clone()
    k2[:, length - 1] = fused[:, d:2 * d]
    v2[:, length - 1] = fused[:, 2 * d:]
    assert torch.equal(kc, k2) and torch.equal(vc, v2)
    assert torch.equal(out, TF.attn_decode(fused[:, :d].contiguous(), k2, v2, heads, length, 8.0))
    # head-major caches [B, heads, capacity, 64]: same arithmetic, contiguous per (image, head)
    kh = k.view(b, cap, heads, 64).permute(0, 2, 1, 3).contiguous()
    vh = v.view(b, cap, heads, 64).permute(0, 2, 1, 3).contiguous()
    outh = TF.attn_decode(fused[:, :d], kh, vh, heads, length, 8.0, k_new=fused[:, d:2 * d], v_new=fused[:, 2 * d:], head_major=True)
    assert torch.equal(outh, out)
    assert torch.equal(kh.permute(0, 2, 1, 3).reshape(b, cap, d), k2)


def test_nrtr_decoder_native_decode_vs_reference_golden(native_lib, golden):
    g = golden("nrtr_decoder.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m = T.NRTRDecoder(n_layers=2, d_embedding=128, n_head=2, d_model=128, d_inner=64, n_position=64, num_classes=37,
                      max_seq_len=12, start_idx=1, padding_idx=36)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    out_enc = torch.from_numpy(g["out_enc"]).to(DEV)
    metas = [dict(valid_ratio=float(r)) for r in g["valid_ratios"]]
    probs = m.forward_test(None, out_enc, metas)
    assert m.last_test_native and probs.shape == (3, 12, 36)
    ref = torch.from_numpy(g["ref32_probs"])
    err = float((probs.cpu() - ref).abs().max())
    floor = float(np.abs(g["ref32_probs"] - g["ref64_probs"]).max())
    print(f"NRTR incremental decode: |probs - ref32| = {err:.3e} (reference fp32 vs fp64 twin {floor:.3e}); "
          f"min top-1 margin of the fixture {float(g['min_margin']):.2f}")
    assert torch.equal(probs.cpu().argmax(-1), ref.argmax(-1))               # identical decodes
    assert err <= 5e-5
    with torch.no_grad():
        lib = m.forward_test_library(None, out_enc, metas)                     # the reference's algorithm on torch ops
    assert float((lib.cpu() - ref).abs().max()) <= 5e-5
    through_forward = m(None, out_enc, None, metas, train_mode=False)
    assert torch.equal(through_forward, probs)


def test_nrtr_decoder_full_config_vs_oracle(native_lib):
    """nrtr_tps++.py's decoder (6 layers, 512 wide, 8 heads, 40 steps, 92 classes), seeded weights: the native incremental
    decode against the oracle (pinned to the reference on this configuration by make_golden) -- identical arg-maxes wherever
    the oracle's own top-1 margin is decisive, probabilities at fp32 level."""
    torch.manual_seed(0)
    m = T.NRTRDecoder().eval()
    with torch.no_grad():
        m.classifier.weight.mul_(8.0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    out_enc = torch.randn((3, 24, 512), generator=g)
    ref = O.nrtr_forward_test(sd, out_enc, [1.0, 0.5, 0.75])
    m = m.to(DEV)
    probs = m.forward_test(None, out_enc.to(DEV), [dict(valid_ratio=1.0), dict(valid_ratio=0.5), dict(valid_ratio=0.75)]).cpu()
    top2 = ref.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    # greedy decode: a flipped near-tie would change everything behind it, so compare up to the first indecisive step per image
    for i in range(ref.shape[0]):
        weak = (margin[i] < 1e-3).nonzero()
        upto = int(weak[0]) if len(weak) else ref.shape[1]
        assert upto >= 1
        assert torch.equal(probs[i, :upto].argmax(-1), ref[i, :upto].argmax(-1))
        assert float((probs[i, :upto] - ref[i, :upto]).abs().max()) <= 2e-4
