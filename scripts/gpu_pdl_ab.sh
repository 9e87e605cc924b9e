#!/bin/bash
# A/B of programmatic dependent launch along the head / stage / warp chain (TPSPP_PDL=0|1), plus the parity tests with it on.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_head_gpu.py tests/test_module_gpu.py tests/test_stage_gpu.py tests/test_warp_gpu.py -q -m gpu --timeout=300 --no-header 2>&1 | tail -4 > gpurun_out/pytest_pdl.log; tail -2 gpurun_out/pytest_pdl.log
for rep in 1 2; do for v in 0 1; do
  TPSPP_PDL=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_pdl$v.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdl$v.json").read())
print("PDL=$v", "ms/step", round(d["ms_per_step"],4), "head", round(d["roofline_head"]["avg_head_ms"],4), "warp", round(d["roofline"]["avg_launch_ms"],4), "from_image", round(d["from_image"]["ms_per_step"],4), "stage", round(d["from_image"]["stage_ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
PY
done; done
