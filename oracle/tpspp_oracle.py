"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the TPS++ rectifier hot path.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package ``tps_pp_b200`` never
does, and fails loudly when its CUDA library is missing.

Parity status: **pinned against the reference run in the build container**
(``oracle/make_golden.py`` imports the unmodified reference sources through
``oracle/ref_loader.py`` and checks every function below against them; the
resulting vectors are committed under ``tests/golden/``).  The reference's own
test-suite holds no golden vectors for this path, so the
reference *code* is the anchor.

All ``file:line`` citations are relative to the reference tree
(``mmocr/models/textrecog/...``):

* ``tps_pp.py``           = ``backbones/tps_pp/tps_pp.py``
* ``DGAB.py``             = ``backbones/tps_pp/DGAB.py``
* ``tps_preprocessor.py`` = ``preprocessor/tps_preprocessor.py``
* ATen sampler semantics  = ``torch/include/ATen/native/GridSampler.h:27-85`` and
  ``ATen/native/cuda/GridSampler.cuh:20-180`` (bilinear, border, align_corners=True)

Two styles live here on purpose:

* explicit numpy loops/vector code for the parts our CUDA kernels re-implement
  (constants, grid generator, bilinear sampler, their backward) -- independent
  of ATen so that ATen itself can be cross-checked;
* ``torch`` CPU functional ops for the dense head (convs, linears, LayerNorm),
  parameterised by dtype so an fp64 twin of every quantity is available.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-6      # tps_pp.py:340, tps_preprocessor.py:171
THETA = 0.5     # tps_pp.py:341 ("thela")


# --------------------------------------------------------------------------
# A-1 constants (host, float64) ---------------------------------------------
# --------------------------------------------------------------------------
def tpspp_C(point_size: Tuple[int, int]) -> np.ndarray:
    """Canonical control points in (0,1), row-major over (py, px); tps_pp.py:368-380."""
    py, px = point_size
    xs = np.linspace(0.5, px - 0.5, num=int(px)) / px
    ys = np.linspace(0.5, py - 0.5, num=int(py)) / py
    gx, gy = np.meshgrid(xs, ys)               # [py, px]
    return np.stack([gx, gy], axis=2).reshape(-1, 2)


def tpspp_P(rect_size: Tuple[int, int]) -> np.ndarray:
    """Target pixel centres in (0,1); tps_pp.py:437-450."""
    hr, wr = rect_size
    xs = np.linspace(0.5, wr - 0.5, num=int(wr)) / wr
    ys = np.linspace(0.5, hr - 0.5, num=int(hr)) / hr
    gx, gy = np.meshgrid(xs, ys)
    return np.stack([gx, gy], axis=2).reshape(-1, 2)


def classical_C(num_fiducial: int) -> np.ndarray:
    """RARE control points on the top/bottom edges in [-1,1]; tps_preprocessor.py:199-211."""
    half = int(num_fiducial / 2)
    xs = np.linspace(-1.0, 1.0, half)
    top = np.stack([xs, -np.ones(half)], axis=1)
    bot = np.stack([xs, np.ones(half)], axis=1)
    return np.concatenate([top, bot], axis=0)


def classical_P(rect_size: Tuple[int, int]) -> np.ndarray:
    """Target pixel centres in (-1,1); tps_preprocessor.py:238-253."""
    hr, wr = rect_size
    xs = (np.arange(-wr, wr, 2) + 1.0) / wr
    ys = (np.arange(-hr, hr, 2) + 1.0) / hr
    gx, gy = np.meshgrid(xs, ys)
    return np.stack([gx, gy], axis=2).reshape(-1, 2)


def inv_delta_C(C: np.ndarray) -> np.ndarray:
    """inv of [[1,C,R],[0,C^T],[0,1]] with R=r^2 ln r (diag r:=1);
    tps_pp.py:382-405 == tps_preprocessor.py:213-236."""
    f = C.shape[0]
    r = np.zeros((f, f), dtype=float)
    for i in range(f):
        for j in range(i, f):
            d = np.linalg.norm(C[i] - C[j])
            r[i, j] = d
            r[j, i] = d
    np.fill_diagonal(r, 1)
    rbf = (r ** 2) * np.log(r)
    top = np.concatenate([np.ones((f, 1)), C, rbf], axis=1)
    mid = np.concatenate([np.zeros((2, 3)), C.T], axis=1)
    bot = np.concatenate([np.zeros((1, 3)), np.ones((1, f))], axis=1)
    return np.linalg.inv(np.concatenate([top, mid, bot], axis=0))


def rbf_P_hat(C: np.ndarray, P: np.ndarray, eps: float = EPS) -> np.ndarray:
    """d^2 ln(d+eps), d=|P_p - C_k|; tps_pp.py:452-465 (TPS++ buffer = these columns only)."""
    diff = P[:, None, :] - C[None, :, :]
    d = np.linalg.norm(diff, ord=2, axis=2)
    return np.square(d) * np.log(d + eps)


def classical_P_hat(C: np.ndarray, P: np.ndarray, eps: float = EPS) -> np.ndarray:
    """[1, P, rbf]; tps_preprocessor.py:255-268."""
    n = P.shape[0]
    return np.concatenate([np.ones((n, 1)), P, rbf_P_hat(C, P, eps)], axis=1)


def tpspp_constants(point_size=(2, 16), rect_size=(16, 64)) -> Dict[str, np.ndarray]:
    """Buffers exactly as the reference registers them: float64 math, ``.float()`` storage
    (tps_pp.py:353-366).  ``P`` is what ``torch.tensor(self.P).float()`` yields (tps_pp.py:472)."""
    C = tpspp_C(point_size)
    P = tpspp_P(rect_size)
    return dict(
        C=C, P64=P,
        hat_C=inv_delta_C(C).astype(np.float32),
        P_hat=rbf_P_hat(C, P).astype(np.float32),
        P=P.astype(np.float32))


def classical_constants(num_fiducial=20, rect_size=(32, 100)) -> Dict[str, np.ndarray]:
    C = classical_C(num_fiducial)
    P = classical_P(rect_size)
    return dict(
        C=C, P64=P,
        inv_delta_C=inv_delta_C(C).astype(np.float32),
        P_hat=classical_P_hat(C, P).astype(np.float32))


# --------------------------------------------------------------------------
# A-2 grid generator ---------------------------------------------------------
# --------------------------------------------------------------------------
def tpspp_grid(c_prime: np.ndarray, pc_score: np.ndarray, hat_C: np.ndarray,
               P: np.ndarray, P_hat: np.ndarray, theta: float = THETA,
               dtype=np.float64) -> np.ndarray:
    """grid[b,p,:] = [1, P_p, P_hat[p,:]*(1+theta*s[b,p,:])] . (hat_C . [C';0]).

    tps_pp.py:467-479 (Phi) and :481-496 (two bmm).  Inputs are the fp32 buffers /
    activations; ``dtype`` selects the arithmetic (float32 reproduces the
    reference's precision class, float64 is the error-free twin)."""
    c_prime = np.asarray(c_prime, dtype=dtype)
    s = np.asarray(pc_score, dtype=dtype)
    b, f, _ = c_prime.shape
    n = P.shape[0]
    hat = np.asarray(hat_C, dtype=dtype)
    T = np.einsum('ij,bjk->bik', hat[:, :f], c_prime)          # zeros pad rows drop out (:489-494)
    rbf = np.asarray(P_hat, dtype=dtype)[None] * (s * dtype(theta) + dtype(1))   # :474
    phi = np.concatenate([np.ones((b, n, 1), dtype=dtype),
                          np.broadcast_to(np.asarray(P, dtype=dtype)[None], (b, n, 2)),
                          rbf], axis=2)                                     # :477
    return np.einsum('bpk,bkc->bpc', phi, T)                                 # :495


def classical_grid(c_prime: np.ndarray, inv_dC: np.ndarray, P_hat: np.ndarray,
                   dtype=np.float64) -> np.ndarray:
    """tps_preprocessor.py:270-282."""
    c_prime = np.asarray(c_prime, dtype=dtype)
    f = c_prime.shape[1]
    T = np.einsum('ij,bjk->bik', np.asarray(inv_dC, dtype=dtype)[:, :f], c_prime)
    return np.einsum('pk,bkc->bpc', np.asarray(P_hat, dtype=dtype), T)


# --------------------------------------------------------------------------
# A-3 bilinear sampler (border padding, align_corners=True) -----------------
# --------------------------------------------------------------------------
def _source_index(coord: np.ndarray, size: int):
    """unnormalise + clip; returns (clipped coord, d clipped / d normalised).

    GridSampler.cuh:23-31 (``((coord + 1.f) / 2) * (size - 1)``), :58-81 (border clip whose
    gradient is zero where the unclipped coordinate is <=0 or >=size-1)."""
    dt = coord.dtype.type
    un = ((coord + dt(1)) / dt(2)) * dt(size - 1)
    lo, hi = dt(0), dt(size - 1)
    inside = (un > lo) & (un < hi)
    clipped = np.minimum(hi, np.maximum(un, lo))
    # CUDA fmaxf/fminf drop NaNs: max(NaN,0)=0 (GridSampler.cuh:52-56)
    clipped = np.where(np.isnan(un), lo, clipped)
    mult = np.where(inside, dt(size - 1) / dt(2), dt(0))
    return clipped, mult


def grid_sample(src: np.ndarray, grid: np.ndarray, dtype=None) -> np.ndarray:
    """src [B,C,H,W], grid [B,Hr,Wr,2] (x,y) -> [B,C,Hr,Wr].

    Taps accumulate in the order nw, ne, sw, se with weights
    nw=(x1-ix)(y1-iy), ne=(ix-x0)(y1-iy), sw=(x1-ix)(iy-y0), se=(ix-x0)(iy-y0)
    (GridSampler.cuh grid_sampler_2d_kernel; out-of-range taps contribute 0)."""
    dtype = dtype or src.dtype
    src = np.asarray(src, dtype=dtype)
    grid = np.asarray(grid, dtype=dtype)
    b, c, h, w = src.shape
    ix, _ = _source_index(grid[..., 0], w)
    iy, _ = _source_index(grid[..., 1], h)
    x0 = np.floor(ix); y0 = np.floor(iy)
    x1 = x0 + 1; y1 = y0 + 1
    wnw = (x1 - ix) * (y1 - iy)
    wne = (ix - x0) * (y1 - iy)
    wsw = (x1 - ix) * (iy - y0)
    wse = (ix - x0) * (iy - y0)
    out = np.zeros((b, c) + grid.shape[1:3], dtype=dtype)
    bi = np.arange(b)[:, None, None]
    for (xx, yy, ww) in ((x0, y0, wnw), (x1, y0, wne), (x0, y1, wsw), (x1, y1, wse)):
        xi = xx.astype(np.int64); yi = yy.astype(np.int64)
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        xi = np.clip(xi, 0, w - 1); yi = np.clip(yi, 0, h - 1)
        v = src[bi, :, yi, xi]                 # [B,Hr,Wr,C]
        v = np.where(ok[..., None], v, 0)
        out = out + np.moveaxis(v * ww[..., None], 3, 1)
    return out


def grid_sample_backward(src: np.ndarray, grid: np.ndarray, gout: np.ndarray, dtype=None):
    """Returns (gsrc [B,C,H,W], ggrid [B,Hr,Wr,2]); SURVEY App. A-3,
    GridSampler.cuh grid_sampler_2d_backward_kernel."""
    dtype = dtype or src.dtype
    src = np.asarray(src, dtype=dtype); grid = np.asarray(grid, dtype=dtype)
    gout = np.asarray(gout, dtype=dtype)
    b, c, h, w = src.shape
    ix, mx = _source_index(grid[..., 0], w)
    iy, my = _source_index(grid[..., 1], h)
    x0 = np.floor(ix); y0 = np.floor(iy)
    x1 = x0 + 1; y1 = y0 + 1
    gsrc = np.zeros_like(src)
    gix = np.zeros_like(ix); giy = np.zeros_like(iy)
    bi = np.broadcast_to(np.arange(b)[:, None, None], ix.shape)
    g = np.moveaxis(gout, 1, 3)                # [B,Hr,Wr,C]
    taps = ((x0, y0, (x1 - ix) * (y1 - iy), -(y1 - iy), -(x1 - ix)),
            (x1, y0, (ix - x0) * (y1 - iy), (y1 - iy), -(ix - x0)),
            (x0, y1, (x1 - ix) * (iy - y0), -(iy - y0), (x1 - ix)),
            (x1, y1, (ix - x0) * (iy - y0), (iy - y0), (ix - x0)))
    for xx, yy, ww, dwx, dwy in taps:
        xi = xx.astype(np.int64); yi = yy.astype(np.int64)
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        xc = np.clip(xi, 0, w - 1); yc = np.clip(yi, 0, h - 1)
        v = np.where(ok[..., None], src[bi, :, yc, xc], 0)          # [B,Hr,Wr,C]
        gv = np.sum(v * g, axis=3)
        gix = gix + gv * dwx
        giy = giy + gv * dwy
        contrib = np.where(ok[..., None], g * ww[..., None], 0)    # [B,Hr,Wr,C]
        for ch in range(c):
            np.add.at(gsrc[:, ch], (bi, yc, xc), contrib[..., ch])
    ggrid = np.stack([gix * mx, giy * my], axis=-1)
    return gsrc, ggrid


def tpspp_grid_backward(ggrid: np.ndarray, c_prime: np.ndarray, pc_score: np.ndarray,
                        hat_C: np.ndarray, P: np.ndarray, P_hat: np.ndarray,
                        theta: float = THETA, dtype=np.float64):
    """(dC' [B,F,2], dpc_score [B,n,F]) from dgrid [B,n,2]; SURVEY App. A-3:
    dT = Phi^T dgrid, dC' = hat_C[:, :F]^T dT, dpc_score = theta * P_hat o (dgrid T[3:]^T)."""
    ggrid = np.asarray(ggrid, dtype=dtype)
    c_prime = np.asarray(c_prime, dtype=dtype)
    s = np.asarray(pc_score, dtype=dtype)
    b, f, _ = c_prime.shape
    n = P.shape[0]
    hat = np.asarray(hat_C, dtype=dtype)[:, :f]
    T = np.einsum('ij,bjk->bik', hat, c_prime)
    ph = np.asarray(P_hat, dtype=dtype)
    phi = np.concatenate([np.ones((b, n, 1), dtype=dtype),
                          np.broadcast_to(np.asarray(P, dtype=dtype)[None], (b, n, 2)),
                          ph[None] * (s * dtype(theta) + dtype(1))], axis=2)
    dT = np.einsum('bpk,bpc->bkc', phi, ggrid)
    dC = np.einsum('kf,bkc->bfc', hat, dT)
    dphi_rbf = np.einsum('bpc,bkc->bpk', ggrid, T[:, 3:, :])
    ds = dtype(theta) * ph[None] * dphi_rbf
    return dC, ds


def classical_grid_backward(ggrid, inv_dC, P_hat, num_fiducial, dtype=np.float64):
    ggrid = np.asarray(ggrid, dtype=dtype)
    hat = np.asarray(inv_dC, dtype=dtype)[:, :num_fiducial]
    dT = np.einsum('pk,bpc->bkc', np.asarray(P_hat, dtype=dtype), ggrid)
    return np.einsum('kf,bkc->bfc', hat, dT)


# --------------------------------------------------------------------------
# fused warp = grid generator + sampler(s) ----------------------------------
# --------------------------------------------------------------------------
def tpspp_warp(feat_grid, x, c_prime, pc_score, consts, dtype=np.float64,
               rect_size=(16, 64)):
    """tps_pp.py:597-615: one grid, two ``grid_sample``s.  The grid is rounded to
    float32 before sampling when dtype is float32 (that is what the reference holds)."""
    g = tpspp_grid(c_prime, pc_score, consts['hat_C'], consts['P'], consts['P_hat'], dtype=dtype)
    b = g.shape[0]
    g4 = g.reshape(b, rect_size[0], rect_size[1], 2)
    out = grid_sample(np.asarray(feat_grid), g4, dtype=dtype)
    mp = grid_sample(np.asarray(x), g4, dtype=dtype) if x is not None else None
    return out, mp, g


def classical_warp(img, c_prime, consts, rect_size, dtype=np.float64):
    """tps_preprocessor.py:70-85."""
    g = classical_grid(c_prime, consts['inv_delta_C'], consts['P_hat'], dtype=dtype)
    g4 = g.reshape(g.shape[0], rect_size[0], rect_size[1], 2)
    return grid_sample(np.asarray(img), g4, dtype=dtype), g


# --------------------------------------------------------------------------
# A-4 head (torch CPU functional, dtype-parametric) -------------------------
# --------------------------------------------------------------------------
def _t(state, key, dtype):
    v = state[key]
    if isinstance(v, np.ndarray):
        v = torch.from_numpy(v)
    return v.detach().to(dtype=dtype, device='cpu')


def _conv_relu(state, prefix, x, stride=1, padding=0):
    """mmcv ConvModule(norm_cfg=None) = conv(bias) + ReLU; tps_pp.py:126-131,149-154,538-548."""
    w = _t(state, prefix + '.conv.weight', x.dtype)
    b = _t(state, prefix + '.conv.bias', x.dtype)
    return F.relu(F.conv2d(x, w, b, stride=stride, padding=padding))


def _linear(state, prefix, x, bias=True):
    w = _t(state, prefix + '.weight', x.dtype)
    b = _t(state, prefix + '.bias', x.dtype) if bias else None
    return F.linear(x, w, b)


def down_stage(state, x, o0, o1):
    """tps_pp.py:581-585 with :560-562 -> (feat_cat [B,192,h,w], feat_grid [B,64,2h,2w])."""
    f0 = _conv_relu(state, 'down0', o0)
    f1 = _conv_relu(state, 'down1', o1)
    f2 = _conv_relu(state, 'down2', x)
    feat_cat = torch.cat((_conv_relu(state, 'down0_1', f0, 2, 1),
                          _conv_relu(state, 'down1_1', f1, 2, 1), f2), dim=1)
    up = F.interpolate(f2, scale_factor=2, mode='nearest')
    feat_grid = _conv_relu(state, 'down_feat', torch.cat((f0, f1, up), dim=1))
    return feat_cat, feat_grid


def cbam(state, prefix, x):
    """tps_pp.py:27-82 (ratio=16): channel gate then spatial gate."""
    w0 = _t(state, prefix + '.channel_attention.shared_MLP.0.weight', x.dtype)
    w2 = _t(state, prefix + '.channel_attention.shared_MLP.2.weight', x.dtype)

    def mlp(v):
        return F.conv2d(F.relu(F.conv2d(v, w0)), w2)

    avg = x.mean(dim=(2, 3), keepdim=True)
    mx = x.amax(dim=(2, 3), keepdim=True)
    out = torch.sigmoid(mlp(avg) + mlp(mx)) * x
    sw = _t(state, prefix + '.spatial_attention.conv2d.weight', x.dtype)
    sb = _t(state, prefix + '.spatial_attention.conv2d.bias', x.dtype)
    sp = torch.cat([out.mean(dim=1, keepdim=True), out.amax(dim=1, keepdim=True)], dim=1)
    return torch.sigmoid(F.conv2d(sp, sw, sb, padding=1)) * out


def msfa(state, feat_cat, p_stride=2):
    """tps_pp.py:156-169: 4 encoder convs (strides 1,2,p_stride,(2,1)), CBAM on the
    deepest map, 4 decoder (nearest-up + conv) with skip adds.  Returns
    (en_feat = pre-CBAM encoder output, de_feat)."""
    strides = [1, 2, p_stride, (2, 1)]
    k = feat_cat
    feats = []
    for i, s in enumerate(strides):
        k = _conv_relu(state, f'MSFA.conv.k_encoder.{i}', k, s, 1)
        feats.append(k)
    point = feats[-1]
    k = cbam(state, 'MSFA.conv.atten', point)
    scales = [(2, 1), p_stride, 2, 1]
    for i in range(4):
        if scales[i] != 1:                      # nn.Upsample(scale_factor=1) is the identity
            k = F.interpolate(k, scale_factor=scales[i], mode='nearest')
        k = _conv_relu(state, f'MSFA.conv.k_decoder.{i}.1', k, 1, 1)
        if i < 3:
            k = k + feats[2 - i]
    return point, k


def dgab(state, prefix, x, y):
    """DGAB.py:74-77 / :39-55 / :17-23.  x [B,C,H,W], y [B,F,C]."""
    h, w = x.shape[2], x.shape[3]
    n1w = _t(state, prefix + '.norm1.weight', x.dtype); n1b = _t(state, prefix + '.norm1.bias', x.dtype)
    u = F.layer_norm(x, (h, w), n1w, n1b)
    yt = y.transpose(1, 2)                                            # b c t
    lw = F.linear(torch.cat([u.mean(2), yt], 2), _t(state, prefix + '.attn.mlp_w.0.weight', x.dtype))
    v_w = lw[:, :, :-1].softmax(dim=-1).unsqueeze(2)
    lh = F.linear(torch.cat([u.mean(3), yt], 2), _t(state, prefix + '.attn.mlp_h.0.weight', x.dtype))
    v_h = lh[:, :, :-1].softmax(dim=-1).unsqueeze(3)
    a = v_h * u * lh[:, :, -1].unsqueeze(-1).unsqueeze(-1) + v_w * u * lw[:, :, -1].unsqueeze(-1).unsqueeze(-1)
    x = x + _linear(state, prefix + '.attn.proj', a)
    n2w = _t(state, prefix + '.norm2.weight', x.dtype); n2b = _t(state, prefix + '.norm2.bias', x.dtype)
    v = F.layer_norm(x, (h, w), n2w, n2b)
    v = _linear(state, prefix + '.mlp.fc2', F.gelu(_linear(state, prefix + '.mlp.fc1', v)))
    return x + v


def tpe(state, en_feat, de_feat, scale=64 ** -0.5):
    """tps_pp.py:315-325 (+ :293-312).  Returns (C' [B,F,2], pc_score [B,hw,F], de' )."""
    b = en_feat.shape[0]
    en = en_feat.flatten(2).transpose(1, 2)                           # b (h w) c
    de = dgab(state, 'TPE.atten.0', de_feat, en)
    z = F.relu(_linear(state, 'TPE.localization_fc1.2',
                       F.relu(_linear(state, 'TPE.localization_fc1.0', en))))
    cp = _linear(state, 'TPE.localization_fc2', z.reshape(b, -1)).view(b, -1, 2)
    feat = de.flatten(2).transpose(1, 2)
    p1 = _linear(state, 'TPE.p_linear.1', _linear(state, 'TPE.p_linear.0', en))
    f = _linear(state, 'TPE.feat_linear.1', _linear(state, 'TPE.feat_linear.0', feat))
    score = torch.tanh(torch.einsum('bmc,bnc->bmn', f, p1) * scale)
    return cp, score, de


def tps_pp_forward(state, x, outs, dtype=torch.float32, point_size=(2, 16),
                   rect_size=(16, 64), p_stride=2, sampler='torch', consts=None):
    """Whole ``TPS_PP.forward`` (tps_pp.py:564-625).  ``sampler='numpy'`` uses the explicit
    sampler above, ``'torch'`` calls ``F.grid_sample`` like the reference does."""
    consts = consts or tpspp_constants(point_size, rect_size)
    x = torch.as_tensor(x).to(dtype); o0 = torch.as_tensor(outs[0]).to(dtype); o1 = torch.as_tensor(outs[1]).to(dtype)
    feat_cat, feat_grid = down_stage(state, x, o0, o1)
    en_feat, de_feat = msfa(state, feat_cat, p_stride)
    cp, score, de2 = tpe(state, en_feat, de_feat, scale=x.shape[1] ** -0.5)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    hat_C = state.get('atten_tps.hat_C', consts['hat_C'])
    P_hat = state.get('atten_tps.P_hat', consts['P_hat'])
    hat_C = hat_C.numpy() if torch.is_tensor(hat_C) else hat_C
    P_hat = P_hat.numpy() if torch.is_tensor(P_hat) else P_hat
    g = tpspp_grid(cp.numpy(), score.numpy(), hat_C, consts['P'], P_hat, dtype=npdt)
    b = g.shape[0]
    g4 = g.reshape(b, rect_size[0], rect_size[1], 2)
    if sampler == 'numpy':
        out = torch.from_numpy(grid_sample(feat_grid.numpy(), g4, dtype=npdt))
        mp = torch.from_numpy(grid_sample(x.numpy(), g4, dtype=npdt))
    else:
        tg = torch.from_numpy(g4)
        out = F.grid_sample(feat_grid, tg, padding_mode='border', align_corners=True)
        mp = F.grid_sample(x, tg, padding_mode='border', align_corners=True)
    return dict(output=out, logits=None, mp_img=mp, pc_score=score,
                control_point=cp, grid=torch.from_numpy(g), feat_grid=feat_grid,
                en_feat=en_feat, de_feat=de_feat, de_feat_dgab=de2, feat_cat=feat_cat)


# --------------------------------------------------------------------------
# classical localisation network (stays torch in the product too) -----------
# --------------------------------------------------------------------------
def classical_localization(state, img, eps=1e-5):
    """tps_preprocessor.py:101-156, eval-mode BatchNorm."""
    x = img
    for ci, bi, pool in ((0, 1, True), (4, 5, True), (8, 9, True), (12, 13, False)):
        p = f'LocalizationNetwork.conv.{ci}'
        q = f'LocalizationNetwork.conv.{bi}'
        x = F.conv2d(x, _t(state, p + '.weight', x.dtype), None, 1, 1)
        x = F.batch_norm(x, _t(state, q + '.running_mean', x.dtype), _t(state, q + '.running_var', x.dtype),
                         _t(state, q + '.weight', x.dtype), _t(state, q + '.bias', x.dtype), False, 0.0, eps)
        x = F.relu(x)
        if pool:
            x = F.max_pool2d(x, 2, 2)
    x = x.mean(dim=(2, 3))
    x = F.relu(_linear(state, 'LocalizationNetwork.localization_fc1.0', x))
    x = _linear(state, 'LocalizationNetwork.localization_fc2', x)
    return x.view(x.shape[0], -1, 2)


def nrtr_decoder_attention(state, trg_seq, src, src_mask, n_head, padding_idx, dtype=torch.float32):
    """NRTRDecoder._attention (nrtr_decoder.py:93-112) over pre-norm TFDecoderLayers (transformer_layers.py:152-165),
    MultiHeadAttention / ScaledDotProductAttention (transformer_module.py:24-34,74-98), eval mode (dropout = identity)."""
    emb = _t(state, 'trg_word_emb.weight', dtype)
    d = emb.shape[1]
    dk = d // n_head
    x = emb[trg_seq] + _t(state, 'position_enc.position_table', dtype)[:, :trg_seq.shape[1]]
    ls = trg_seq.shape[1]
    causal = (1 - torch.triu(torch.ones((ls, ls)), diagonal=1)).unsqueeze(0).bool()
    trg_mask = (trg_seq != padding_idx).unsqueeze(-2) & causal

    def lin(prefix, v):
        w = _t(state, prefix + '.weight', dtype)
        b = _t(state, prefix + '.bias', dtype) if (prefix + '.bias') in state else None
        return F.linear(v, w, b)

    def mha(prefix, q, k, v, mask):
        b, lq, lk = q.shape[0], q.shape[1], k.shape[1]
        qq = lin(prefix + '.linear_q', q).view(b, lq, n_head, dk).transpose(1, 2)
        kk = lin(prefix + '.linear_k', k).view(b, lk, n_head, dk).transpose(1, 2)
        vv = lin(prefix + '.linear_v', v).view(b, lk, n_head, dk).transpose(1, 2)
        a = torch.matmul(qq / dk ** 0.5, kk.transpose(2, 3))
        if mask is not None:
            m4 = mask.unsqueeze(1) if mask.dim() == 3 else mask.unsqueeze(1).unsqueeze(1)
            a = a.masked_fill(m4 == 0, float('-inf'))
        out = torch.matmul(F.softmax(a, dim=-1), vv).transpose(1, 2).contiguous().view(b, lq, n_head * dk)
        return lin(prefix + '.fc', out)

    i = 0
    while f'layer_stack.{i}.norm1.weight' in state:
        p = f'layer_stack.{i}.'
        ln = lambda nm, v: F.layer_norm(v, (d,), _t(state, p + nm + '.weight', dtype), _t(state, p + nm + '.bias', dtype), 1e-5)
        h = ln('norm1', x)
        x = x + mha(p + 'self_attn', h, h, h, trg_mask)
        x = x + mha(p + 'enc_attn', ln('norm2', x), src, src, src_mask)
        x = x + lin(p + 'mlp.w_2', F.gelu(lin(p + 'mlp.w_1', ln('norm3', x))))
        i += 1
    return F.layer_norm(x, (d,), _t(state, 'layer_norm.weight', dtype), _t(state, 'layer_norm.bias', dtype), 1e-6)


def nrtr_forward_test(state, out_enc, valid_ratios=None, n_head=8, max_seq_len=40, start_idx=1, padding_idx=92):
    """NRTRDecoder.forward_test (nrtr_decoder.py:153-177): greedy decode that re-runs the decoder over the whole padded
    sequence at every step; ``valid_ratios`` -> the source mask of ``_get_mask`` (:114-127).  Returns probabilities [N, T, C-1]."""
    import math
    dtype = out_enc.dtype
    n, t, _ = out_enc.shape
    src_mask = None
    if valid_ratios is not None:
        src_mask = torch.zeros((n, t), dtype=dtype)
        for i, vr in enumerate(valid_ratios):
            src_mask[i, :min(t, math.ceil(t * vr))] = 1
    seq = torch.full((n, max_seq_len + 1), padding_idx, dtype=torch.long)
    seq[:, 0] = start_idx
    outs = []
    for step in range(max_seq_len):
        dec = nrtr_decoder_attention(state, seq, out_enc, src_mask, n_head, padding_idx, dtype)
        logits = F.linear(dec[:, step, :], _t(state, 'classifier.weight', dtype), _t(state, 'classifier.bias', dtype))
        probs = F.softmax(logits, dim=-1)
        outs.append(probs)
        seq[:, step + 1] = probs.argmax(dim=-1)
    return torch.stack(outs, dim=1)


def moran_grid(target=(32, 128)) -> np.ndarray:
    """The identity grid [H, W, 2] of moran.py:51-64 (x from the width axis, y from the height axis, align_corners)."""
    h = np.arange(target[0]) * 2. / (target[0] - 1) - 1
    w = np.arange(target[1]) * 2. / (target[1] - 1) - 1
    g = np.stack(np.meshgrid(w, h, indexing='ij'), axis=-1)
    return np.transpose(g, (1, 0, 2)).astype(np.float32)


def moran_forward(state, x, target=(32, 128), enhance=0, eps=1e-5):
    """moran.py:66-103, eval-mode BatchNorm: offset CNN on the bilinearly down-scaled image, MaxPool(2,1) of the positive
    and negative parts, the offset map and then the image resampled with border padding / align_corners."""
    dt = x.dtype
    grid = torch.from_numpy(moran_grid(target)).to(dt)[None].repeat(x.shape[0], 1, 1, 1)
    gx, gy = grid[..., 0:1], grid[..., 1:2]

    def cnn(img):
        y = F.max_pool2d(img, 2, 2)
        for ci, pool, relu in ((1, True, True), (5, True, True), (9, False, True), (12, False, True), (15, False, False)):
            y = F.conv2d(y, _t(state, f'cnn.{ci}.weight', dt), _t(state, f'cnn.{ci}.bias', dt), 1, 1)
            q = f'cnn.{ci + 1}'
            y = F.batch_norm(y, _t(state, q + '.running_mean', dt), _t(state, q + '.running_var', dt), _t(state, q + '.weight', dt),
                             _t(state, q + '.bias', dt), False, 0.0, eps)
            if relu:
                y = F.relu(y)
            if pool:
                y = F.max_pool2d(y, 2, 2)
        return y

    def offsets_of(img):
        off = cnn(img)
        pooled = F.max_pool2d(F.relu(off), 2, 1) - F.max_pool2d(F.relu(-off), 2, 1)
        return F.grid_sample(pooled, grid, padding_mode='border', align_corners=True).permute(0, 2, 3, 1)

    small = F.interpolate(x, size=tuple(target), mode='bilinear', align_corners=True)
    og = offsets_of(small)
    out = F.grid_sample(x, torch.cat([gx, gy + og], 3), padding_mode='border', align_corners=True)
    for _ in range(enhance):
        og = og + offsets_of(out)
        out = F.grid_sample(x, torch.cat([gx, gy + og], 3), padding_mode='border', align_corners=True)
    return out


# --------------------------------------------------------------------------
# deterministic synthetic inputs (numpy legacy RNG: stable across versions) --
# --------------------------------------------------------------------------
def smooth_c_prime(base: np.ndarray, batch: int, seed: int = 7, amp: float = 0.05,
                   centre: float = 0.5) -> np.ndarray:
    """"trained-like" control points: a smooth per-image field
    (scale + shear + sinusoidal baseline) applied to the init lattice ``base`` [F,2]."""
    rs = np.random.RandomState(seed)
    out = np.empty((batch,) + base.shape, dtype=np.float64)
    for i in range(batch):
        a, sh, sc = rs.uniform(-amp, amp, 3)
        fr = rs.uniform(0.5, 2.0); ph = rs.uniform(0, 2 * math.pi)
        x = base[:, 0]; y = base[:, 1]
        out[i, :, 0] = (x - centre) * (1 + sc) + centre + sh * (y - centre)
        out[i, :, 1] = y + a * np.sin(2 * math.pi * fr * x + ph)
    return out.astype(np.float32)


def tpspp_init_bias(point_size=(2, 16)) -> np.ndarray:
    """localization_fc2.bias lattice, tps_pp.py:280-285."""
    py, px = point_size
    xs = np.linspace(0.1, px - 0.1, num=int(px)) / px
    ys = np.linspace(0.1, py - 0.1, num=int(py)) / py
    gx, gy = np.meshgrid(xs, ys)
    return np.stack([gx, gy], axis=2).reshape(-1, 2)


def classical_init_bias(num_fiducial=20) -> np.ndarray:
    """tps_preprocessor.py:132-141."""
    half = int(num_fiducial / 2)
    xs = np.linspace(-1.0, 1.0, half)
    top = np.stack([xs, np.linspace(0.0, -1.0, num=half)], axis=1)
    bot = np.stack([xs, np.linspace(1.0, 0.0, num=half)], axis=1)
    return np.concatenate([top, bot], axis=0)


TPSPP_PARAM_SHAPES: Sequence[Tuple[str, Tuple[int, ...]]] = (
    # SURVEY App. A-5, in the reference's construction order (tps_pp.py:533-548)
    *[(f'MSFA.conv.k_encoder.{i}.conv.weight', (64, 192 if i == 0 else 64, 3, 3)) for i in range(4)],
    ('MSFA.conv.atten.channel_attention.shared_MLP.0.weight', (4, 64, 1, 1)),
    ('MSFA.conv.atten.channel_attention.shared_MLP.2.weight', (64, 4, 1, 1)),
    ('MSFA.conv.atten.spatial_attention.conv2d.weight', (1, 2, 3, 3)),
    *[(f'MSFA.conv.k_decoder.{i}.1.conv.weight', (64, 64, 3, 3)) for i in range(4)],
    ('TPE.p_linear.0.weight', (32, 64)), ('TPE.p_linear.1.weight', (128, 32)),
    ('TPE.feat_linear.0.weight', (32, 64)), ('TPE.feat_linear.1.weight', (128, 32)),
    ('TPE.atten.0.attn.mlp_h.0.weight', (17, 48)), ('TPE.atten.0.attn.mlp_w.0.weight', (65, 96)),
    ('TPE.atten.0.attn.proj.weight', (64, 64)),
    ('TPE.atten.0.mlp.fc1.weight', (256, 64)), ('TPE.atten.0.mlp.fc2.weight', (64, 256)),
    ('TPE.localization_fc1.0.weight', (256, 64)), ('TPE.localization_fc1.2.weight', (2, 256)),
    ('down0.conv.weight', (64, 32, 1, 1)), ('down1.conv.weight', (64, 32, 1, 1)),
    ('down2.conv.weight', (64, 64, 1, 1)), ('down0_1.conv.weight', (64, 64, 3, 3)),
    ('down1_1.conv.weight', (64, 64, 3, 3)), ('down_feat.conv.weight', (64, 192, 1, 1)),
)
_NO_BIAS = {'MSFA.conv.atten.channel_attention.shared_MLP.0.weight',
            'MSFA.conv.atten.channel_attention.shared_MLP.2.weight',
            'TPE.atten.0.attn.mlp_h.0.weight', 'TPE.atten.0.attn.mlp_w.0.weight'}


def trained_like_state(seed: int = 3, fc2_sigma: float = 5e-5, ln_sigma: float = 0.1,
                       point_size=(2, 16), rect_size=(16, 64), field_amp: float = 0.05
                       ) -> Dict[str, torch.Tensor]:
    """A full TPS_PP state_dict (58 tensors + 2 buffers, App. A-5) drawn from numpy's legacy
    RNG, scaled so that activations stay O(1) through the head (He-uniform for layers feeding
    a ReLU, Xavier-like otherwise), LayerNorm ~ (1,0)+N(0,ln_sigma).  SURVEY F8: the stock
    init has ``localization_fc2.weight == 0`` and an affine bias lattice, so neither the
    features nor ``pc_score`` reach the grid.  Here ``fc2.weight ~ N(0, fc2_sigma^2)`` (C' moves
    by a few 1e-4 with the features -- more is numerically meaningless, F8/C-13), the ReLU before
    it is kept alive by a +0.5 bias, and ``fc2.bias`` = init lattice + one smooth non-affine
    field so the RBF weights are non-zero and ``pc_score`` moves samples by several pixels."""
    rs = np.random.RandomState(seed)
    st: Dict[str, torch.Tensor] = {}
    relu_fed = ('conv.weight', 'shared_MLP.0', 'localization_fc1', 'mlp.fc1')
    for name, shape in TPSPP_PARAM_SHAPES:
        fan_in = int(np.prod(shape[1:]))
        gain = 6.0 if any(t in name for t in relu_fed) else 3.0
        bound = math.sqrt(gain / fan_in)
        st[name] = torch.from_numpy(rs.uniform(-bound, bound, shape).astype(np.float32))
        if name not in _NO_BIAS:
            st[name[:-6] + 'bias'] = torch.from_numpy(rs.uniform(-0.1, 0.1, shape[:1]).astype(np.float32))
    st['TPE.localization_fc1.2.bias'] = st['TPE.localization_fc1.2.bias'] + 0.5
    hh, ww = rect_size
    for nm in ('norm1', 'norm2'):
        st[f'TPE.atten.0.{nm}.weight'] = torch.from_numpy((1 + ln_sigma * rs.standard_normal((hh, ww))).astype(np.float32))
        st[f'TPE.atten.0.{nm}.bias'] = torch.from_numpy((ln_sigma * rs.standard_normal((hh, ww))).astype(np.float32))
    f = point_size[0] * point_size[1]
    st['TPE.localization_fc2.weight'] = torch.from_numpy((fc2_sigma * rs.standard_normal((2 * f, 2 * f))).astype(np.float32))
    lattice = smooth_c_prime(tpspp_init_bias(point_size), 1, seed=seed + 100, amp=field_amp)[0]
    st['TPE.localization_fc2.bias'] = torch.from_numpy(lattice.reshape(-1).astype(np.float32))
    c = tpspp_constants(point_size, rect_size)
    st['atten_tps.hat_C'] = torch.from_numpy(c['hat_C'])
    st['atten_tps.P_hat'] = torch.from_numpy(c['P_hat'])
    return st


def synthetic_tpspp_inputs(batch: int, seed: int = 0, h: int = 16, w: int = 64):
    """x [B,64,h,w], outs 2x[B,32,2h,2w] ~ N(0,1) (BASELINE config 2), numpy legacy RNG."""
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((batch, 64, h, w)).astype(np.float32)
    o0 = rs.standard_normal((batch, 32, 2 * h, 2 * w)).astype(np.float32)
    o1 = rs.standard_normal((batch, 32, 2 * h, 2 * w)).astype(np.float32)
    return x, o0, o1


# --------------------------------------------------------------------------
# Backbone stage in front of the call: stem + layer1 + layer2 of ResNetABI_v2_large --------------
# --------------------------------------------------------------------------
# (layer, blocks, in planes, planes, stride of the first block) for strides=[1,2,...] (the only setting under
# which TPS_PP's shape contract holds); backbones/resnet_v2_large.py:44-96
BACKBONE_STAGE_LAYERS = (('layer1', 3, 32, 32, 1), ('layer2', 4, 32, 64, 2))
BN_EPS = 1e-5   # nn.BatchNorm2d default


def backbone_stage_keys():
    """state_dict keys/shapes of the stage, in the reference module's order (resnet_v2_large.py:131-135,109-129;
    layers/conv_layer.py:12-33 on top of mmcv BasicBlock: conv1 = 1x1 stride 1, conv2 = 3x3 with the block's stride)."""
    def bn(prefix, c):
        return [(prefix + '.weight', (c,)), (prefix + '.bias', (c,)), (prefix + '.running_mean', (c,)),
                (prefix + '.running_var', (c,)), (prefix + '.num_batches_tracked', ())]
    keys = [('conv1.weight', (32, 3, 3, 3)), ('conv1.bias', (32,))] + bn('bn1', 32)
    for layer, blocks, cin, planes, stride in BACKBONE_STAGE_LAYERS:
        for i in range(blocks):
            p = f'{layer}.{i}.'
            ci = cin if i == 0 else planes
            keys += [(p + 'conv1.weight', (planes, ci, 1, 1))] + bn(p + 'bn1', planes)
            keys += [(p + 'conv2.weight', (planes, planes, 3, 3))] + bn(p + 'bn2', planes)
            if i == 0 and (stride != 1 or ci != planes):
                keys += [(p + 'downsample.0.weight', (planes, ci, 1, 1))] + bn(p + 'downsample.1', planes)
    return keys


def trained_like_backbone_state(seed: int = 5) -> Dict[str, torch.Tensor]:
    """Deterministic (numpy legacy RNG) weights for the stage: He-uniform convolutions, BatchNorm with non-trivial
    affine parameters and running statistics so that folding BN into the convolution is actually exercised."""
    rs = np.random.RandomState(seed)
    st: Dict[str, torch.Tensor] = {}
    for name, shape in backbone_stage_keys():
        if name.endswith('num_batches_tracked'):
            st[name] = torch.tensor(100, dtype=torch.int64)
        elif name.endswith('running_var'):
            st[name] = torch.from_numpy(rs.uniform(0.5, 1.5, shape).astype(np.float32))
        elif name.endswith('running_mean'):
            st[name] = torch.from_numpy((0.1 * rs.standard_normal(shape)).astype(np.float32))
        elif len(shape) == 4:
            bound = math.sqrt(6.0 / int(np.prod(shape[1:])))
            st[name] = torch.from_numpy(rs.uniform(-bound, bound, shape).astype(np.float32))
        elif name.endswith('.weight'):      # BN gamma; the residual branch (bn2) is damped so activations stay O(1..10)
            lo, hi = (0.1, 0.3) if '.bn2.' in name else (0.6, 1.4)
            st[name] = torch.from_numpy(rs.uniform(lo, hi, shape).astype(np.float32))
        else:                               # conv1.bias, BN beta
            st[name] = torch.from_numpy((0.1 * rs.standard_normal(shape)).astype(np.float32))
    return st


def synthetic_images(batch: int, seed: int = 1234, h: int = 32, w: int = 128) -> np.ndarray:
    """img [B,3,32,128] ~ N(0,1) (BASELINE configs 1/5), numpy legacy RNG."""
    return np.random.RandomState(seed).standard_normal((batch, 3, h, w)).astype(np.float32)


def _bn_eval(state, prefix, x):
    return F.batch_norm(x, _t(state, prefix + '.running_mean', x.dtype), _t(state, prefix + '.running_var', x.dtype),
                        _t(state, prefix + '.weight', x.dtype), _t(state, prefix + '.bias', x.dtype), False, 0.0, BN_EPS)


def backbone_stage_forward(state, img, dtype=torch.float32):
    """What ``ResNetABI_v2_large.forward`` computes before it calls ``tpsnet(x, outs)`` (resnet_v2_large.py:176-191), eval
    mode: stem conv+BN+ReLU, then layer1 and layer2; ``outs`` = [stem output, layer1 output].  Block = mmcv BasicBlock
    forward with the reference's 1x1/3x3 pair (conv_layer.py:30-33): relu(bn1(conv1 x)) -> bn2(conv2 .) -> + identity
    (through ``downsample`` = 1x1 stride-s conv + BN when present) -> relu.  Returns (x, [o0, o1])."""
    x = torch.as_tensor(np.asarray(img)).to(dtype)
    x = F.relu(_bn_eval(state, 'bn1', F.conv2d(x, _t(state, 'conv1.weight', dtype), _t(state, 'conv1.bias', dtype), padding=1)))
    outs = []
    for layer, blocks, cin, planes, stride in BACKBONE_STAGE_LAYERS:
        outs.append(x)
        for i in range(blocks):
            p = f'{layer}.{i}.'
            s = stride if i == 0 else 1
            out = F.relu(_bn_eval(state, p + 'bn1', F.conv2d(x, _t(state, p + 'conv1.weight', dtype))))
            out = _bn_eval(state, p + 'bn2', F.conv2d(out, _t(state, p + 'conv2.weight', dtype), stride=s, padding=1))
            idn = x
            if (p + 'downsample.0.weight') in state:
                idn = _bn_eval(state, p + 'downsample.1', F.conv2d(x, _t(state, p + 'downsample.0.weight', dtype), stride=s))
            x = F.relu(out + idn)
    return x, outs


def head_intermediates(state, x, outs, dtype=torch.float32, p_stride=2):
    """Every intermediate of the head, named like the native workspace slots (include/tpspp.h TPSPP_WS_*),
    computed with the same stage functions as :func:`tps_pp_forward` (tps_pp.py:581-594, 156-169;
    DGAB.py:39-55,74-77)."""
    x = torch.as_tensor(x).to(dtype); o0 = torch.as_tensor(outs[0]).to(dtype); o1 = torch.as_tensor(outs[1]).to(dtype)
    r = {}
    r['f0'] = _conv_relu(state, 'down0', o0)
    r['f1'] = _conv_relu(state, 'down1', o1)
    r['f2'] = _conv_relu(state, 'down2', x)
    r['a0'] = _conv_relu(state, 'down0_1', r['f0'], 2, 1)
    r['a1'] = _conv_relu(state, 'down1_1', r['f1'], 2, 1)
    up = F.interpolate(r['f2'], scale_factor=2, mode='nearest')
    r['feat_grid'] = _conv_relu(state, 'down_feat', torch.cat((r['f0'], r['f1'], up), dim=1))
    k = torch.cat((r['a0'], r['a1'], r['f2']), dim=1)
    for i, s in enumerate([1, 2, p_stride, (2, 1)]):
        k = _conv_relu(state, f'MSFA.conv.k_encoder.{i}', k, s, 1)
        r[f'e{i}'] = k
    k = cbam(state, 'MSFA.conv.atten', r['e3'])
    r['cbam'] = k
    skips = [r['e2'], r['e1'], r['e0']]
    for i, sc in enumerate([(2, 1), p_stride, 2, 1]):
        if sc != 1:
            k = F.interpolate(k, scale_factor=sc, mode='nearest')
        k = _conv_relu(state, f'MSFA.conv.k_decoder.{i}.1', k, 1, 1)
        if i < 3:
            k = k + skips[i]
            r[f'd{i}'] = k
    r['de'] = k
    # DGAB split at the same points as the native kernels
    pre = 'TPE.atten.0'
    en = r['e3'].flatten(2).transpose(1, 2)
    h, w = k.shape[2], k.shape[3]
    u = F.layer_norm(k, (h, w), _t(state, pre + '.norm1.weight', dtype), _t(state, pre + '.norm1.bias', dtype))
    yt = en.transpose(1, 2)
    lw = F.linear(torch.cat([u.mean(2), yt], 2), _t(state, pre + '.attn.mlp_w.0.weight', dtype))
    lh = F.linear(torch.cat([u.mean(3), yt], 2), _t(state, pre + '.attn.mlp_h.0.weight', dtype))
    v_w = lw[:, :, :-1].softmax(dim=-1).unsqueeze(2)
    v_h = lh[:, :, :-1].softmax(dim=-1).unsqueeze(3)
    a = v_h * u * lh[:, :, -1].unsqueeze(-1).unsqueeze(-1) + v_w * u * lw[:, :, -1].unsqueeze(-1).unsqueeze(-1)
    r['x1'] = k + _linear(state, pre + '.attn.proj', a)
    r['v'] = F.layer_norm(r['x1'], (h, w), _t(state, pre + '.norm2.weight', dtype), _t(state, pre + '.norm2.bias', dtype))
    r['de2'] = r['x1'] + _linear(state, pre + '.mlp.fc2', F.gelu(_linear(state, pre + '.mlp.fc1', r['v'])))
    r['p1'] = _linear(state, 'TPE.p_linear.1', _linear(state, 'TPE.p_linear.0', en))
    cp, score, de2 = tpe(state, r['e3'], r['de'], scale=x.shape[1] ** -0.5)
    r['c_prime'] = cp
    r['pc_score'] = score
    return r
