"""Drop-in ``TPSPreprocessor`` / ``BasePreprocessor`` (registry ``PREPROCESSOR``).

Interface kept from the reference (preprocessor/tps_preprocessor.py:24-85,
preprocessor/base_preprocessor.py:7-12): ctor kwargs ``num_fiducial=20, img_size=(32,100),
rectified_img_size=(32,100), num_img_channel=1, init_cfg=None`` with its five asserts,
``forward(batch_img[B,C,H,W]) -> Tensor[B,C,Hr,Wr]``, state_dict keys
``LocalizationNetwork.conv.{0,4,8,12}.weight``, BN ``{1,5,9,13}.*``,
``LocalizationNetwork.localization_fc1.0.*``, ``LocalizationNetwork.localization_fc2.*`` and the
buffers ``GridGenerator.inv_delta_C`` / ``GridGenerator.P_hat``.

The grid generator + sampler is the fused native kernel (classical mode).  The localisation network (97 % of the
preprocessor's time) runs on the native kernels too (``tpspp_locnet_fwd``) whenever autograd is
not recording, the module is in eval mode (BatchNorm folded) and the image geometry tiles (64x256, 64x128, 32x256, ...);
otherwise -- training, or e.g. the 32x100 default of the recogniser configs -- it stays the cuDNN library stack.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from . import constants as K
from . import functional as TF
from .rectifier import _BaseModule
from .registry import PREPROCESSOR


@PREPROCESSOR.register_module()
class BasePreprocessor(_BaseModule):
    """Identity preprocessor (base_preprocessor.py:7-12)."""

    def forward(self, x, **kwargs):
        return x


def _localization_convs(cin: int) -> nn.Sequential:
    layers = []
    chans = (cin, 64, 128, 256, 512)
    for i in range(4):
        layers += [nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1, bias=False), nn.BatchNorm2d(chans[i + 1]), nn.ReLU(True)]
        layers.append(nn.MaxPool2d(2, 2) if i < 3 else nn.AdaptiveAvgPool2d(1))
    return nn.Sequential(*layers)


@PREPROCESSOR.register_module()
class TPSPreprocessor(BasePreprocessor):
    def __init__(self, num_fiducial=20, img_size=(32, 100), rectified_img_size=(32, 100),
                 num_img_channel=1, init_cfg=None):
        super().__init__(init_cfg=init_cfg)
        assert isinstance(num_fiducial, int)
        assert num_fiducial > 0
        assert isinstance(img_size, tuple)
        assert isinstance(rectified_img_size, tuple)
        assert isinstance(num_img_channel, int)
        self.num_fiducial = num_fiducial
        self.img_size = img_size
        self.rectified_img_size = rectified_img_size
        self.num_img_channel = num_img_channel

        loc = nn.Module()
        self.add_module("LocalizationNetwork", loc)
        loc.add_module("conv", _localization_convs(num_img_channel))
        loc.add_module("localization_fc1", nn.Sequential(nn.Linear(512, 256), nn.ReLU(True)))
        loc.add_module("localization_fc2", nn.Linear(256, num_fiducial * 2))
        with torch.no_grad():
            loc.localization_fc2.weight.fill_(0)
            loc.localization_fc2.bias.copy_(torch.from_numpy(K.classical_init_bias(num_fiducial)).float().view(-1))

        inv_dc, p_hat, _ = K.classical_tps_buffers(num_fiducial, rectified_img_size)
        gen = nn.Module()
        self.add_module("GridGenerator", gen)
        gen.register_buffer("inv_delta_C", torch.from_numpy(inv_dc))
        gen.register_buffer("P_hat", torch.from_numpy(p_hat))
        self.warp_variant = N.VARIANT_AUTO
        # "auto": native localisation network where it applies (see the module docstring); "library": always cuDNN / cuBLAS
        self.locnet_impl = "auto"
        self._locnet_ws = {}
        self._last_locnet_native = None

    def _use_native_locnet(self, batch_img: torch.Tensor) -> bool:
        if self.locnet_impl == "library" or self.training or torch.is_grad_enabled():
            return False
        if batch_img.dtype != torch.float32 or batch_img.dim() != 4:
            return False
        _, c, h, w = batch_img.shape
        return c == self.num_img_channel and TF.locnet_supported(c, h, w, self.num_fiducial)

    def localize(self, batch_img: torch.Tensor) -> torch.Tensor:
        """C' [B,F,2] (tps_preprocessor.py:143-156)."""
        loc = self.LocalizationNetwork
        self._last_locnet_native = self._use_native_locnet(batch_img)
        if self._last_locnet_native:
            sd = dict(loc.named_parameters())
            sd.update(dict(loc.named_buffers()))
            params = [sd[k] for k in TF.locnet_param_keys()]
            key = (batch_img.device.index, torch.cuda.current_stream(batch_img.device).cuda_stream)
            stamp = (batch_img.shape[0], tuple((p.data_ptr(), p._version) for p in params))
            ws, ws_stamp = self._locnet_ws.get(key, (None, None))
            cp, ws = TF.locnet_forward(batch_img, params, self.num_fiducial, ws, weights_cached=(ws_stamp == stamp))
            if len(self._locnet_ws) >= 8 and key not in self._locnet_ws:
                self._locnet_ws.clear()
            self._locnet_ws[key] = (ws, stamp)
            return cp
        feats = loc.conv(batch_img).view(batch_img.size(0), -1)
        return loc.localization_fc2(loc.localization_fc1(feats)).view(batch_img.size(0), self.num_fiducial, 2)

    def forward(self, batch_img: torch.Tensor, **kwargs) -> torch.Tensor:
        if not batch_img.is_cuda:
            raise RuntimeError("tps_pp_b200.TPSPreprocessor runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        c_prime = self.localize(batch_img)
        gen = self.GridGenerator
        out, _ = TF.tps_warp(batch_img, None, c_prime.float(), None, gen.P_hat, None, gen.inv_delta_C,
                             self.rectified_img_size, N.MODE_CLASSICAL, 0.0, self.warp_variant)
        return out


@PREPROCESSOR.register_module()
class MORAN(BasePreprocessor):
    """Drop-in MORAN rectifier (reference preprocessor/moran.py:13-103): a small offset CNN on the down-scaled image, the
    offset map resampled to the target size and the image resampled along ``y + offset`` -- two
    ``F.grid_sample(padding_mode='border', align_corners=True)`` calls per pass (``moran.py:89-103``), which run on the
    native sampling core (``tpspp_sample_fwd``) whenever autograd is not recording; the offset CNN (five small
    convolutions on a 16 x 64 map) stays a library stack.  Same ctor kwargs, ``cnn.*`` state_dict keys and ``grid`` buffer."""

    def __init__(self, num_img_channel=3, img_size=(32, 128), maxBatch=256):
        super().__init__()
        self.targetH = img_size[0]
        self.targetW = img_size[1]
        self.maxBatch = maxBatch
        self.cnn = nn.Sequential(
            nn.MaxPool2d(2, 2),
            nn.Conv2d(num_img_channel, 64, 3, 1, 1), nn.BatchNorm2d(64), nn.ReLU(True), nn.MaxPool2d(2, 2),
            nn.Conv2d(64, 128, 3, 1, 1), nn.BatchNorm2d(128), nn.ReLU(True), nn.MaxPool2d(2, 2),
            nn.Conv2d(128, 64, 3, 1, 1), nn.BatchNorm2d(64), nn.ReLU(True),
            nn.Conv2d(64, 16, 3, 1, 1), nn.BatchNorm2d(16), nn.ReLU(True),
            nn.Conv2d(16, 1, 3, 1, 1), nn.BatchNorm2d(1))
        self.pool = nn.MaxPool2d(2, 1)
        self.register_buffer("grid", torch.tensor(self.grid_process(maxBatch)).float())
        self._last_sampler_native = None

    def grid_process(self, maxBatch):
        """The identity sampling grid [maxBatch, H, W, 2] (moran.py:51-64)."""
        h_list = np.arange(self.targetH) * 2. / (self.targetH - 1) - 1
        w_list = np.arange(self.targetW) * 2. / (self.targetW - 1) - 1
        grid = np.stack(np.meshgrid(w_list, h_list, indexing="ij"), axis=-1)
        grid = np.transpose(grid, (1, 0, 2))
        return np.tile(np.expand_dims(grid, 0), [maxBatch, 1, 1, 1])

    def _sample(self, src, grid):
        if self._last_sampler_native:
            return TF.grid_sample_border(src, grid)
        return F.grid_sample(src, grid, padding_mode="border", align_corners=True)

    def forward(self, x, test=None, enhance=0, debug=False):
        if not x.is_cuda:
            raise RuntimeError("tps_pp_b200.MORAN runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        assert x.size(0) <= self.maxBatch
        self._last_sampler_native = (not torch.is_grad_enabled()) and x.dtype == torch.float32
        b = x.size(0)
        grid = self.grid[:b]
        grid_x = grid[:, :, :, 0].unsqueeze(3)
        grid_y = grid[:, :, :, 1].unsqueeze(3)
        x_small = F.interpolate(x, size=(self.targetH, self.targetW), mode="bilinear", align_corners=True)

        def offsets_of(img):
            # the offset CNN feeds sampling coordinates: keep the library convolutions in true fp32 (cuDNN's default TF32
            # moves the offsets by ~1e-3, i.e. the rectified pixels by the same order)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                off = self.cnn(img)
            pooled = self.pool(F.relu(off)) - self.pool(F.relu(-off))
            return self._sample(pooled.contiguous(), grid.contiguous()).permute(0, 2, 3, 1).contiguous()

        offsets_grid = offsets_of(x_small)
        x_rectified = self._sample(x, torch.cat([grid_x, grid_y + offsets_grid], 3))
        for _ in range(enhance):
            offsets_grid = offsets_grid + offsets_of(x_rectified)
            x_rectified = self._sample(x, torch.cat([grid_x, grid_y + offsets_grid], 3))
        return x_rectified
