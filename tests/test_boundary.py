"""CPU: the drop-in boundary -- registries, ctor contracts, state_dict layout, the C-ABI library
loads and exports every symbol include/tpspp.h declares.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

import tps_pp_b200 as T
from tps_pp_b200 import _native, constants as K, functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "tpspp.h")).read()
    declared = re.findall(r"TPSPP_API\s+[\w\s\*]+?\b(tpspp_\w+)\s*\(", header)
    assert len(declared) >= 8
    for name in declared:
        assert hasattr(native_lib, name), f"libtpspp.so does not export {name}"
    assert set(declared) == set(_native.exported_symbols())
    assert native_lib.tpspp_version() == 2


def test_cfg_struct_matches_header():
    header = open(os.path.join(ROOT, "include", "tpspp.h")).read()
    body = re.search(r"typedef struct tpspp_warp_cfg \{(.*?)\} tpspp_warp_cfg;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _native.WarpCfg._fields_]
    body = re.search(r"typedef struct tpspp_head_cfg \{(.*?)\} tpspp_head_cfg;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _native.HeadCfg._fields_]


def test_workspace_query_is_host_only(native_lib):
    import ctypes
    cfg = _native.WarpCfg(256, 64, 32, 128, 64, 16, 64, 16, 64, 32, 0, 0.5, 0, 0)
    nbytes = native_lib.tpspp_warp_workspace_bytes(ctypes.byref(cfg))
    assert nbytes >= 256 * 1024 * 2 * 4
    bad = _native.WarpCfg(1, 0, 32, 128, 0, 0, 0, 16, 64, 32, 0, 0.5, 0, 0)
    assert native_lib.tpspp_warp_workspace_bytes(ctypes.byref(bad)) == 0
    assert native_lib.tpspp_warp_fwd_workspace_bytes(ctypes.byref(cfg)) == 0          # attention mode: none
    cls = _native.WarpCfg(1024, 3, 64, 256, 0, 0, 0, 64, 256, 40, 1, 0.0, 0, 0)
    assert native_lib.tpspp_warp_fwd_workspace_bytes(ctypes.byref(cls)) == 1024 * 43 * 2 * 8
    assert b"src0 geometry" in native_lib.tpspp_last_error()


def test_registry_builds_reference_names():
    m = T.build_backbone(dict(type="TPS_PP"))
    assert isinstance(m, T.TPS_PP)
    p = T.build_preprocessor(dict(type="TPSPreprocessor", num_fiducial=20, img_size=(32, 100),
                                  rectified_img_size=(32, 100), num_img_channel=3))
    assert isinstance(p, T.TPSPreprocessor)
    with pytest.raises(KeyError):
        T.build_backbone(dict(type="NoSuchThing"))


def test_ctor_contracts_like_reference_tests():
    # tests/test_models/test_ocr_preprocessor.py:10-17 of the reference
    with pytest.raises(AssertionError):
        T.TPSPreprocessor(num_fiducial=-1)
    with pytest.raises(AssertionError):
        T.TPSPreprocessor(img_size=32)
    with pytest.raises(AssertionError):
        T.TPSPreprocessor(rectified_img_size=100)
    with pytest.raises(AssertionError):
        T.TPSPreprocessor(num_img_channel="bgr")
    with pytest.raises(AssertionError):
        T.TPS_PP(img_size=[16, 64])
    with pytest.raises(AssertionError):
        T.TPS_PP(rectified_img_size=64)
    pre = T.BasePreprocessor()
    pre.init_weights()
    x = torch.randn(1, 1, 32, 100)
    assert pre(x).shape == x.shape


def test_state_dict_layout_and_init_match_reference(golden):
    g = golden("constants.npz")
    torch.manual_seed(0)
    m = T.TPS_PP()
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["init_keys"]]
    assert len(sd) == 60 and sum(p.numel() for p in m.parameters()) == 546597
    sums = np.array([float(v.double().sum()) for v in sd.values()])
    abss = np.array([float(v.double().abs().sum()) for v in sd.values()])
    assert np.array_equal(sums, g["init_sums"]) and np.array_equal(abss, g["init_abs_sums"])
    assert float(sd["TPE.localization_fc2.weight"].abs().max()) == 0.0
    # the 14 ConvModule convs carry mmcv's ctor-time kaiming-normal(fan_out, relu) init with zero bias
    w = sd["down0_1.conv.weight"]
    assert abs(float(w.std()) - (2.0 / (64 * 9)) ** 0.5) < 0.05 * (2.0 / (64 * 9)) ** 0.5
    assert float(sd["down0_1.conv.bias"].abs().max()) == 0.0 and float(sd["MSFA.conv.k_encoder.0.conv.bias"].abs().max()) == 0.0
    # BaseModule.init_weights(): the six direct ConvModule children are drawn again, once
    m.init_weights()
    sd2 = m.state_dict()
    sums2 = np.array([float(v.double().sum()) for v in sd2.values()])
    assert np.array_equal(sums2, g["init2_sums"])
    m.init_weights()
    assert np.array_equal(np.array([float(v.double().sum()) for v in m.state_dict().values()]), g["init2_sums"])


def test_head_param_shapes_table_matches_module():
    """functional.head_param_shapes is what head_forward validates the 58 tensors against before handing raw
    pointers to the kernels (a module built for another geometry must raise, not read out of bounds)."""
    from tps_pp_b200 import functional as TF
    m = T.TPS_PP()
    assert [tuple(p.shape) for p in m.parameters()] == TF.head_param_shapes(16, 64, 32)
    assert TF.head_param_shapes(8, 64, 16) != TF.head_param_shapes(16, 64, 32)     # LayerNorm / gate weights follow (h, F)


def test_constants_match_reference_buffers(golden):
    g = golden("constants.npz")
    hat, ph, p, _ = K.attention_tps_buffers((2, 16), (16, 64))
    assert np.array_equal(hat, g["tpspp_hat_C"]) and np.array_equal(ph, g["tpspp_P_hat"]) and np.array_equal(p, g["tpspp_P"])
    for f, rs in ((20, (32, 100)), (6, (8, 12))):
        inv, pc, _ = K.classical_tps_buffers(f, rs)
        assert np.array_equal(inv, g[f"classical_F{f}_{rs[0]}x{rs[1]}_inv_delta_C"])
        assert np.array_equal(pc, g[f"classical_F{f}_{rs[0]}x{rs[1]}_P_hat"])
    c = T.TPSPreprocessor(20, (32, 100), (32, 100), 1)
    keys = list(c.state_dict().keys())
    assert "GridGenerator.inv_delta_C" in keys and "GridGenerator.P_hat" in keys
    assert "LocalizationNetwork.conv.13.running_var" in keys and "LocalizationNetwork.localization_fc2.bias" in keys


def test_no_cpu_fallback():
    m = T.TPS_PP()
    x = torch.randn(1, 64, 16, 64)
    o = torch.randn(1, 32, 32, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, [o, o])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.TPSPreprocessor()(torch.randn(1, 1, 32, 100))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tps_pp_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{fn} mentions the oracle"


def test_launch_profile_is_a_no_op_without_work(native_lib):
    """Host-only part of the measurement aid: enabling it and reading before any head call yields zero launches."""
    import ctypes
    buf = (ctypes.c_float * 8)()
    cnt = ctypes.c_int(-1)
    assert native_lib.tpspp_launch_profile(1) == 0
    assert native_lib.tpspp_launch_profile_read(buf, 8, ctypes.byref(cnt)) == 0 and cnt.value == 0
    assert native_lib.tpspp_launch_profile(0) == 0
    assert native_lib.tpspp_launch_profile_read(None, 8, ctypes.byref(cnt)) != 0       # bad arguments are reported
    assert b"tpspp_launch_profile_read" in native_lib.tpspp_last_error()



def test_every_cfg_struct_has_the_header_layout(tmp_path):
    """sizeof / field offsets of the ctypes mirrors against include/tpspp.h compiled as plain C (the header must also BE plain C:
    the reference-side bindings are cgo / ctypes style)."""
    import ctypes
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    pairs = [("tpspp_warp_cfg", _native.WarpCfg), ("tpspp_head_cfg", _native.HeadCfg), ("tpspp_stage_cfg", _native.StageCfg),
             ("tpspp_conv_cfg", _native.ConvCfg), ("tpspp_linear_cfg", _native.LinearCfg), ("tpspp_locnet_cfg", _native.LocnetCfg),
             ("tpspp_attn_cfg", _native.AttnCfg)]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "tpspp.h"', "int main(void) {"]
    for cname, mirror in pairs:
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for (cname, mirror), line in zip(pairs, out):
        parts = line.split()
        assert parts[0] == cname
        assert int(parts[1]) == ctypes.sizeof(mirror), cname
        offs = [int(v) for v in parts[2:]]
        assert offs == [getattr(mirror, f[0]).offset for f in mirror._fields_], cname


def test_training_and_locnet_workspace_queries_are_host_only(native_lib):
    """tpspp_conv / tpspp_linear / tpspp_locnet workspace queries validate their cfg on the host (no GPU needed)."""
    import ctypes
    conv = _native.ConvCfg(8, 192, 16, 64, 3, 1, 1, 1, 3)
    for i in range(3):
        conv.up_h[i] = conv.up_w[i] = 1
    assert native_lib.tpspp_conv_workspace_bytes(ctypes.byref(conv)) > 0
    bad = _native.ConvCfg(8, 48, 16, 64, 3, 1, 1, 1, 1)
    assert native_lib.tpspp_conv_workspace_bytes(ctypes.byref(bad)) == 0 and b"cin" in native_lib.tpspp_last_error()
    lin = _native.LinearCfg(4096, 64, 256, 1)
    assert native_lib.tpspp_linear_workspace_bytes(ctypes.byref(lin)) > 0
    assert native_lib.tpspp_linear_workspace_bytes(ctypes.byref(_native.LinearCfg(100, 64, 256, 3))) == 0      # rows % batches
    assert native_lib.tpspp_locnet_workspace_bytes(ctypes.byref(_native.LocnetCfg(4, 3, 64, 256, 20, 0))) > 0
    assert native_lib.tpspp_locnet_workspace_bytes(ctypes.byref(_native.LocnetCfg(4, 1, 32, 100, 20, 0))) == 0  # 100 % 8
    assert native_lib.tpspp_locnet_workspace_bytes(ctypes.byref(_native.LocnetCfg(4, 1, 32, 128, 20, 0))) == 0  # 4x16 map does not tile
    assert b"not supported" in native_lib.tpspp_last_error()
    assert native_lib.tpspp_locnet_workspace_bytes(ctypes.byref(_native.LocnetCfg(4, 2, 64, 256, 20, 0))) == 0  # 2 channels


def test_locnet_param_table_matches_module():
    """The 24-tensor table tpspp_locnet_fwd takes = the LocalizationNetwork slice of the reference state_dict order."""
    m = T.TPSPreprocessor(num_fiducial=20, img_size=(64, 256), rectified_img_size=(64, 256), num_img_channel=3)
    keys = [k for k in m.state_dict() if k.startswith("LocalizationNetwork.") and not k.endswith("num_batches_tracked")]
    assert keys == ["LocalizationNetwork." + k for k in TF.locnet_param_keys()]
    sd = m.state_dict()
    assert [tuple(sd[k].shape) for k in keys] == TF.locnet_param_shapes(3, 20)
    assert len(keys) == _native.LP_COUNT
