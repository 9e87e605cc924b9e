"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference sources by file path.

Only usable where ``/root/reference`` exists (the build container).  It is used
by ``oracle/make_golden.py`` to pin ``oracle/tpspp_oracle.py`` against the real
reference and to emit the fixtures under ``tests/golden/``.  Nothing in the
product package (``tps_pp_b200``), in the ``-m gpu`` tests, in ``smoke()`` or in
``bench.py`` may import this module: the reference tree does not travel to the
GPU box.

The reference package cannot be imported as shipped (missing
modules, mmcv/mmdet/timm absent), so the handful of third-party symbols its
hot-path files use are stubbed here with their documented behaviour
(SURVEY.md App. B):

* ``mmcv.cnn.ConvModule``  -> ``self.conv = Conv2d(bias=True)``, ``self.activate = ReLU(inplace=True)``,
  followed by mmcv's ctor-time ``init_weights()`` = ``kaiming_normal_(fan_out, relu)`` + zero bias
  (mmcv-full 1.3.8-1.5.0 with ``norm_cfg=None``; used at ``tps_pp.py:126-131,149-154,538-548``)
* ``mmcv.runner.BaseModule`` -> ``nn.Module`` + ``init_weights()``
* ``timm.models.layers.DropPath`` -> identity (``DGAB.py:67``, drop_path=0)
* ``mmocr.models.builder.{BACKBONES,PREPROCESSOR}`` -> decorator registries
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("TPSPP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(
        REF_ROOT, "mmocr/models/textrecog/backbones/tps_pp/tps_pp.py"))


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg)


class _ConvModule(nn.Module):
    """mmcv ConvModule with norm_cfg=None, act_cfg=dict(type='ReLU') (defaults)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size,
                              stride=stride, padding=padding, bias=True)
        self.activate = nn.ReLU(inplace=True)
        self.init_weights()

    def init_weights(self):
        # mmcv ConvModule.init_weights (called at the end of its ctor and again by BaseModule.init_weights):
        # kaiming_init(conv, a=0, mode='fan_out', nonlinearity='relu', distribution='normal'), bias = 0
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        return self.activate(self.conv(x))


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        # mmcv BaseModule.init_weights with init_cfg=None: direct children that define init_weights, once
        if getattr(self, "_is_init", False):
            return
        for m in self.children():
            if hasattr(m, "init_weights"):
                m.init_weights()
        self._is_init = True


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


_LOADED = {}


def _install_stubs():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_tpspp_stub", False):
        return
    _mod("mmcv", _tpspp_stub=True)
    _mod("mmcv.cnn", ConvModule=_ConvModule)
    _mod("mmcv.runner", BaseModule=_BaseModule, ModuleList=nn.ModuleList,
         Sequential=nn.Sequential)
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=nn.Identity)
    regs = {n: _Registry(n) for n in ("BACKBONES", "PREPROCESSOR")}
    _mod("mmocr")
    _mod("mmocr.models")
    _mod("mmocr.models.builder", **regs)
    _mod("mmocr.models.textrecog")
    _mod("mmocr.models.textrecog.backbones")
    _mod("mmocr.models.textrecog.backbones.tps_pp")
    _mod("mmocr.models.textrecog.preprocessor")


def _load(modname, relpath):
    if modname in _LOADED:
        return _LOADED[modname]
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    _LOADED[modname] = mod
    return mod


def load_reference():
    """Return a namespace with the reference's hot-path classes."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    dgab = _load("mmocr.models.textrecog.backbones.tps_pp.DGAB",
                 "mmocr/models/textrecog/backbones/tps_pp/DGAB.py")
    tps_pp = _load("mmocr.models.textrecog.backbones.tps_pp.tps_pp",
                   "mmocr/models/textrecog/backbones/tps_pp/tps_pp.py")
    _load("mmocr.models.textrecog.preprocessor.base_preprocessor",
          "mmocr/models/textrecog/preprocessor/base_preprocessor.py")
    tpsp = _load("mmocr.models.textrecog.preprocessor.tps_preprocessor",
                 "mmocr/models/textrecog/preprocessor/tps_preprocessor.py")
    moran = _load("mmocr.models.textrecog.preprocessor.moran", "mmocr/models/textrecog/preprocessor/moran.py")
    return types.SimpleNamespace(
        MORAN=moran.MORAN,
        TPS_PP=tps_pp.TPS_PP,
        Attention_Enhanced_TPS=tps_pp.Attention_Enhanced_TPS,
        DGAB=dgab.DGAB,
        TPSPreprocessor=tpsp.TPSPreprocessor,
        GridGenerator=tpsp.GridGenerator,
        tps_pp_module=tps_pp, dgab_module=dgab, tps_preprocessor_module=tpsp)


def build_quiet(ctor, *args, seed=0, **kwargs):
    """Construct a reference module under a fixed seed with its prints silenced
    (``tps_pp.py:212,268,291,558`` print parameter counts)."""
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return ctor(*args, **kwargs)
