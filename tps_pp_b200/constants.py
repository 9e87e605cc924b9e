"""Host-side (numpy float64) construction of the TPS constants, stored as float32 buffers.

These reproduce, bit for bit, the buffers the reference registers at construction time:
``Attention_Enhanced_TPS`` (tps_pp.py:353-405,437-465: ``hat_C`` = inverse of delta_C, ``P_hat`` =
rbf columns only, ``P`` re-created every forward at tps_pp.py:472) and the classical
``GridGenerator`` (tps_preprocessor.py:176-268: ``inv_delta_C``, ``P_hat`` = [1, P, rbf]).
tests/test_constants.py checks them against the committed golden buffers.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

EPS = 1e-6


def _cell_centres(count: int) -> np.ndarray:
    """(0.5, 1.5, ..., count-0.5)/count as the reference's linspace builds it."""
    return np.linspace(0.5, count - 0.5, num=int(count)) / count


def _lattice(xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """Row-major (y outer, x inner) list of (x, y) pairs."""
    out = np.empty((ys.size * xs.size, 2), dtype=np.float64)
    out[:, 0] = np.tile(xs, ys.size)
    out[:, 1] = np.repeat(ys, xs.size)
    return out


def _pair_dist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    d = a[:, None, :] - b[None, :, :]
    return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])


def _inverse_delta_c(c: np.ndarray) -> np.ndarray:
    f = c.shape[0]
    r = _pair_dist(c, c)
    r[np.arange(f), np.arange(f)] = 1.0
    delta = np.zeros((f + 3, f + 3), dtype=np.float64)
    delta[:f, 0] = 1.0
    delta[:f, 1:3] = c
    delta[:f, 3:] = (r ** 2) * np.log(r)
    delta[f:f + 2, 3:] = c.T
    delta[f + 2, 3:] = 1.0
    return np.linalg.inv(delta)


def _rbf(c: np.ndarray, p: np.ndarray) -> np.ndarray:
    d = _pair_dist(p, c)
    return np.square(d) * np.log(d + EPS)


def attention_tps_buffers(point_size: Tuple[int, int], rect_size: Tuple[int, int]):
    """-> (hat_C [F+3,F+3] f32, P_hat [n,F] f32, P [n,2] f32, C [F,2] f64) in the (0,1) convention."""
    py, px = point_size
    hr, wr = rect_size
    c = _lattice(_cell_centres(px), _cell_centres(py))
    p = _lattice(_cell_centres(wr), _cell_centres(hr))
    return (_inverse_delta_c(c).astype(np.float32), _rbf(c, p).astype(np.float32),
            p.astype(np.float32), c)


def classical_tps_buffers(num_fiducial: int, rect_size: Tuple[int, int]):
    """-> (inv_delta_C [F+3,F+3] f32, P_hat [n,F+3] f32, C [F,2] f64) in the [-1,1] convention."""
    half = int(num_fiducial / 2)
    xs = np.linspace(-1.0, 1.0, half)
    c = np.concatenate([np.stack([xs, -np.ones(half)], 1), np.stack([xs, np.ones(half)], 1)], 0)
    hr, wr = rect_size
    gx = (np.arange(-wr, wr, 2) + 1.0) / wr
    gy = (np.arange(-hr, hr, 2) + 1.0) / hr
    p = _lattice(gx, gy)
    p_hat = np.concatenate([np.ones((p.shape[0], 1)), p, _rbf(c, p)], axis=1)
    return _inverse_delta_c(c).astype(np.float32), p_hat.astype(np.float32), c


def attention_init_bias(point_size: Tuple[int, int]) -> np.ndarray:
    """Initial ``localization_fc2.bias`` lattice of TPS++ (tps_pp.py:280-285), [F,2] f64."""
    py, px = point_size
    xs = np.linspace(0.1, px - 0.1, num=int(px)) / px
    ys = np.linspace(0.1, py - 0.1, num=int(py)) / py
    return _lattice(xs, ys)


def classical_init_bias(num_fiducial: int) -> np.ndarray:
    """Initial ``localization_fc2.bias`` of the RARE localisation net (tps_preprocessor.py:132-141)."""
    half = int(num_fiducial / 2)
    xs = np.linspace(-1.0, 1.0, half)
    top = np.stack([xs, np.linspace(0.0, -1.0, num=half)], axis=1)
    bot = np.stack([xs, np.linspace(1.0, 0.0, num=half)], axis=1)
    return np.concatenate([top, bot], axis=0)
