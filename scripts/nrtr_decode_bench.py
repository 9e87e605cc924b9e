#!/usr/bin/env python
"""NRTR decoder greedy decode (reference decoders/nrtr_decoder.py:153-177; BASELINE configs[4]'s "+ NRTR inference" part):
the native incremental decode of tps_pp_b200.NRTRDecoder.forward_test against the reference's algorithm (full recompute of
the padded prefix at each of the 40 steps) on torch / cuBLAS fp32 ops, same module, same weights, same GPU.
python scripts/nrtr_decode_bench.py [batch] [src_len]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tps_pp_b200 as T  # noqa: E402


def timeit(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    Tsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 64          # 4 x 16 feature map of a 32 x 128 image after layer5
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = T.NRTRDecoder().to(dev).eval()
    torch.backends.cuda.matmul.allow_tf32 = False
    out_enc = torch.randn((B, Tsrc, 512), device=dev)
    with torch.no_grad():
        p_nat = m.forward_test(None, out_enc, None)
        p_lib = m.forward_test_library(None, out_enc, None)
        same = float((p_nat.argmax(-1) == p_lib.argmax(-1)).float().mean())
        m.decode_graph = False
        native_ms = timeit(lambda: m.forward_test(None, out_enc, None), 3, 1)
        m.decode_graph = True
        graph_ms = timeit(lambda: m.forward_test(None, out_enc, None), 5, 2)
        lib_ms = timeit(lambda: m.forward_test_library(None, out_enc, None), 2, 1)
    best = graph_ms if graph_ms is not None else native_ms
    print(json.dumps({"metric": "nrtr_decoder_greedy_decode", "batch": B, "src_len": Tsrc, "steps": m.max_seq_len,
                      "native_ms": native_ms, "native_graph_ms": graph_ms, "library_reference_algorithm_ms": lib_ms,
                      "img_per_s": B / (best * 1e-3), "speedup_vs_library": lib_ms / best,
                      "argmax_agreement_native_vs_library": same,
                      "what": "6 layers, d_model 512, 8 heads, 40 greedy steps, 92 classes, random-init weights; native = KV-cache "
                              "incremental decode (tpspp_linear_fwd + tpspp_attn_decode); library = the reference's full-prefix "
                              "recompute on torch / cuBLAS fp32 ops (TF32 off)"}))


if __name__ == "__main__":
    main()
