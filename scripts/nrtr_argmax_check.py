#!/usr/bin/env python
"""Offline check of the north-star clause "identical NRTR argmax decodes on the synthetic batch" (BASELINE config 1).

The reference recogniser (ResNetABI_v2_large -> TPS_PP -> NRTR encoder/decoder, ``encode_decode_recognizer.py:107-122``)
only exists in the build container and the CUDA rectifier only runs on the GPU box, so the check has three stages:

  prepare  (build container, CPU)  run the unmodified reference on a seeded synthetic batch, capture the tensors it
                                   hands to ``tpsnet(x, outs)`` (``resnet_v2_large.py:189-191``), the tpsnet weights, the
                                   reference rectifier output and the decoded strings -> tests/golden/_nrtr/stage.npz
  gpu      (B200)                  run tps_pp_b200.TPS_PP on the captured tensors with the same weights
                                   -> gpurun_out/nrtr_ours.npz
  compare  (build container, CPU)  replay our rectified features through the rest of the reference network (layer3-5,
                                   encoder, autoregressive decoder) and compare argmax strings / class probabilities
                                   -> profiles/r01_nrtr_argmax.md

Two weight sets for the rectifier: the stock init (localization_fc2.weight = 0) and the trained-like synthetic
state of ``oracle.trained_like_state`` (C' depends on the features).  TEST INFRASTRUCTURE: imports ``oracle/``.
"""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGE = os.path.join(ROOT, "tests", "golden", "_nrtr", "stage.npz")
OURS = os.path.join(ROOT, "gpurun_out", "nrtr_ours.npz")
B = 8


def _stack():
    from oracle import nrtr_loader as L
    ns = L.load_nrtr()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        tps = ns.TPS_PP()
        bb = ns.ResNetABI_v2_large(arch_settings=[3, 4, 6, 6, 3], strides=[1, 2, 2, 1, 2])
        enc = ns.NRTREncoder()
        dec = ns.NRTRDecoder(num_classes=93, start_idx=91, padding_idx=92, max_seq_len=40)
    conv = ns.AttnConvertor("DICT90", with_unknown=True, max_seq_len=40)
    for m in (tps, bb, enc, dec):
        m.eval()
    return tps, bb, enc, dec, conv


def _decode(bb, enc, dec, conv, img, tpsnet):
    metas = [{"valid_ratio": 1.0}] * img.shape[0]
    with torch.no_grad():
        f = bb(img, tpsnet, True)["output"]
        probs = dec(f, enc(f, metas), None, metas, train_mode=False)
    idx, _ = conv.tensor2idx(probs, metas)
    return conv.idx2str(idx), probs


def _weight_sets(tps):
    from oracle import tpspp_oracle as O
    stock = {k: v.detach().clone() for k, v in tps.state_dict().items()}
    trained = {k: v.detach().clone() for k, v in O.trained_like_state().items()}
    assert trained.keys() == stock.keys()
    return {"stock": stock, "trained": trained}


def prepare():
    tps, bb, enc, dec, conv = _stack()
    img = torch.randn(B, 3, 32, 128, generator=torch.Generator().manual_seed(1234))
    out = {"img": img.numpy()}
    for name, sd in _weight_sets(tps).items():
        tps.load_state_dict(sd, strict=True)
        cap = {}

        def tapped(x, outs, **kw):
            r = tps(x, outs, **kw)
            cap.update(x=x.detach().clone(), o0=outs[0].detach().clone(), o1=outs[1].detach().clone(),
                       output=r["output"].detach().clone())
            return r
        strings, probs = _decode(bb, enc, dec, conv, img, tapped)
        out.update({"x": cap["x"].numpy(), "o0": cap["o0"].numpy(), "o1": cap["o1"].numpy(),
                    f"{name}_ref_output": cap["output"].numpy(), f"{name}_ref_probs": probs.numpy(),
                    f"{name}_ref_strings": np.array(strings)})
        for k, v in sd.items():
            out[f"{name}_sd/{k}"] = v.numpy()
        print(name, strings[:2])
    os.makedirs(os.path.dirname(STAGE), exist_ok=True)
    np.savez(STAGE, **out)
    print("wrote", STAGE, os.path.getsize(STAGE) >> 20, "MiB")


def gpu():
    import tps_pp_b200
    from tps_pp_b200 import _native as N
    z = np.load(STAGE)
    dev = "cuda:0"
    x, o0, o1 = (torch.from_numpy(z[k]).to(dev) for k in ("x", "o0", "o1"))
    res = {}
    for name in ("stock", "trained"):
        sd = {k.split("/", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{name}_sd/")}
        for prec, tag in ((N.HEAD_TC, "tc"), (N.HEAD_FP32, "fp32")):
            m = tps_pp_b200.TPS_PP().to(dev).eval()
            m.head_precision = prec
            m.load_state_dict(sd, strict=True)
            with torch.no_grad():
                r = m(x, [o0, o1])
            assert all(m.native_stages.values()), m.native_stages
            res[f"{name}_{tag}_output"] = r["output"].cpu().numpy()
            print(name, tag, float(np.abs(res[f"{name}_{tag}_output"] - z[f"{name}_ref_output"]).max()))
    os.makedirs(os.path.dirname(OURS), exist_ok=True)
    np.savez(OURS, **res)


def compare():
    z, ours = np.load(STAGE), np.load(OURS)
    tps, bb, enc, dec, conv = _stack()
    img = torch.from_numpy(z["img"])
    rows = []
    for name in ("stock", "trained"):
        ref_strings = [str(s) for s in z[f"{name}_ref_strings"]]
        ref_probs = z[f"{name}_ref_probs"]
        for tag in ("tc", "fp32"):
            o = torch.from_numpy(ours[f"{name}_{tag}_output"])
            strings, probs = _decode(bb, enc, dec, conv, img, lambda x, outs, **kw: {"output": o})
            same = sum(a == b for a, b in zip(strings, ref_strings))
            ref_arg = ref_probs.argmax(-1)
            rows.append(dict(weights=name, head=tag, strings_equal=f"{same}/{len(strings)}",
                             argmax_positions_equal=f"{int((probs.numpy().argmax(-1) == ref_arg).sum())}/{ref_arg.size}",
                             max_abs_output_diff=float(np.abs(o.numpy() - z[f"{name}_ref_output"]).max()),
                             max_abs_prob_diff=float(np.abs(probs.numpy() - ref_probs).max()),
                             example=strings[0]))
    print(json.dumps(rows, indent=1))
    with open(os.path.join(ROOT, "profiles", "r01_nrtr_argmax.md"), "w") as fh:
        fh.write("# r01 — NRTR argmax decodes with the B200 rectifier swapped in (BASELINE config 1)\n\n"
                 "`scripts/nrtr_argmax_check.py prepare | gpu | compare`: batch 8 of N(0,1) images [8,3,32,128] (seed 1234),\n"
                 "reference ResNetABI_v2_large(strides=[1,2,2,1,2]) + NRTR encoder/decoder, random init under seed 0,\n"
                 "greedy decode of 40 steps (`nrtr_decoder.py:153-177`).  The rectified features fed to layer3-5 come from\n"
                 "the reference TPS_PP (CPU fp32) or from `tps_pp_b200.TPS_PP` on a B200 given the same `(x, outs)` and weights.\n\n"
                 "| rectifier weights | head mode | decoded strings equal | argmax positions equal | max abs diff `output` | max abs diff class probs |\n"
                 "|---|---|---|---|---|---|\n")
        for r in rows:
            fh.write(f"| {r['weights']} | {r['head']} | {r['strings_equal']} | {r['argmax_positions_equal']} | "
                     f"{r['max_abs_output_diff']:.2e} | {r['max_abs_prob_diff']:.2e} |\n")
        fh.write("\nExample decode (random-init network, so the text is noise): `" + rows[0]["example"] + "`\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("stage", choices=["prepare", "gpu", "compare"])
    {"prepare": prepare, "gpu": gpu, "compare": compare}[ap.parse_args().stage]()
