"""TEST INFRASTRUCTURE ONLY -- loads the reference's NRTR recogniser (backbone, encoder, decoder, convertor)
by file path, on top of ``oracle.ref_loader``'s stubs, for the offline "identical argmax decodes" check
(``scripts/nrtr_argmax_check.py``).  Build container only: needs ``/root/reference``.

Extra third-party stubs, all with their documented behaviour:

* ``mmcv.cnn.resnet.conv3x3`` / ``BasicBlock`` (mmcv-full 1.3.x ``mmcv/cnn/resnet.py``): conv3x3-bn-relu-conv3x3-bn,
  ``out += residual`` (through ``downsample`` when given), relu.  Used by ``layers/conv_layer.py:3-4``.
* ``mmocr.models.builder.build_activation_layer`` -> ``nn.GELU`` for ``'mmcv.GELU'`` (``transformer_module.py:117``).
* ``mmocr.utils.is_type_list`` / ``list_from_file``; dummy names for the missing ``backbones/tps.py`` and
  ``tools.data.textrecog.visual_feat.draw_feature_map``.
"""
from __future__ import annotations

import sys
import types

import torch.nn as nn

from . import ref_loader as R


def _conv3x3(in_planes, out_planes, stride=1, dilation=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=dilation, dilation=dilation, bias=False)


class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style="pytorch", with_cp=False):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride, dilation)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out += residual
        return self.relu(out)


def _is_type_list(x, t):
    return isinstance(x, list) and all(isinstance(i, t) for i in x)


def _list_from_file(path, encoding="utf-8"):
    with open(path, encoding=encoding) as fh:
        return [ln.rstrip("\n\r") for ln in fh]


def _build_activation_layer(cfg):
    kind = dict(cfg)["type"]
    if kind in ("mmcv.GELU", "GELU"):
        return nn.GELU()
    if kind == "ReLU":
        return nn.ReLU()
    raise KeyError(kind)


def load_nrtr():
    """Namespace with the reference classes needed for config 1 (NRTR + TPS_PP forward)."""
    ref = R.load_reference()
    m = sys.modules
    R._mod("mmcv.cnn.resnet", conv3x3=_conv3x3, BasicBlock=_BasicBlock)
    b = m["mmocr.models.builder"]
    for n in ("ENCODERS", "DECODERS", "CONVERTORS"):
        if not hasattr(b, n):
            setattr(b, n, R._Registry(n))
    b.build_activation_layer = _build_activation_layer
    R._mod("mmocr.utils", is_type_list=_is_type_list, list_from_file=_list_from_file)
    m["mmocr"].utils = m["mmocr.utils"]
    dummies = {n: type(n, (), {}) for n in ("U_TPSnet", "Deform_net", "DAttentionBaseline", "UDAT_Net", "TPSnet",
                                             "TPSnet_Warp", "TPSnetv2")}
    R._mod("mmocr.models.textrecog.backbones.tps", **dummies)
    R._mod("tools"); R._mod("tools.data"); R._mod("tools.data.textrecog")
    R._mod("tools.data.textrecog.visual_feat", draw_feature_map=lambda *a, **k: None)
    base = "mmocr/models/"
    conv_layer = R._load("mmocr.models.textrecog.layers.conv_layer", base + "textrecog/layers/conv_layer.py")
    R._mod("mmocr.models.textrecog.layers", BasicBlock=conv_layer.BasicBlock)
    bb = R._load("mmocr.models.textrecog.backbones.resnet_v2_large", base + "textrecog/backbones/resnet_v2_large.py")
    R._mod("mmocr.models.common")
    tm = R._load("mmocr.models.common.modules.transformer_module", base + "common/modules/transformer_module.py")
    R._mod("mmocr.models.common.modules", MultiHeadAttention=tm.MultiHeadAttention, PositionalEncoding=tm.PositionalEncoding,
           PositionwiseFeedForward=tm.PositionwiseFeedForward, ScaledDotProductAttention=tm.ScaledDotProductAttention)
    tl = R._load("mmocr.models.common.layers.transformer_layers", base + "common/layers/transformer_layers.py")
    common = m["mmocr.models.common"]
    for n in ("TFEncoderLayer", "TFDecoderLayer"):
        setattr(common, n, getattr(tl, n))
    common.PositionalEncoding = tm.PositionalEncoding
    R._mod("mmocr.models.textrecog.encoders"); R._mod("mmocr.models.textrecog.decoders"); R._mod("mmocr.models.textrecog.convertors")
    R._load("mmocr.models.textrecog.encoders.base_encoder", base + "textrecog/encoders/base_encoder.py")
    enc = R._load("mmocr.models.textrecog.encoders.nrtr_encoder", base + "textrecog/encoders/nrtr_encoder.py")
    R._load("mmocr.models.textrecog.decoders.base_decoder", base + "textrecog/decoders/base_decoder.py")
    dec = R._load("mmocr.models.textrecog.decoders.nrtr_decoder", base + "textrecog/decoders/nrtr_decoder.py")
    R._load("mmocr.models.textrecog.convertors.base", base + "textrecog/convertors/base.py")
    attn = R._load("mmocr.models.textrecog.convertors.attn", base + "textrecog/convertors/attn.py")
    return types.SimpleNamespace(ref=ref, TPS_PP=ref.TPS_PP, ResNetABI_v2_large=bb.ResNetABI_v2_large,
                                 NRTREncoder=enc.NRTREncoder, NRTRDecoder=dec.NRTRDecoder, AttnConvertor=attn.AttnConvertor)
