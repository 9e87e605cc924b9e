#!/bin/bash
# A/B pass: head + module + stage tests, then the bench in the default mode and with all-tf32 3xTF32 convolutions.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest head/module/stage"; timeout 1200 python -m pytest tests/test_head_gpu.py tests/test_module_gpu.py tests/test_stage_gpu.py -q -m gpu --timeout=300 --no-header -rA -s 2>&1 | grep -v "^$" | tail -300 > gpurun_out/pytest_head.log; tail -3 gpurun_out/pytest_head.log
run() { name=$1; shift; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>&1 | tail -1 > gpurun_out/bench_$name.json
python - $name <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{f}.json").read())
    print(f, "%.4f ms"%d["ms_per_step"], d["roofline_dominant"]["per_launch_ms"])
    if d.get("from_image"): print("   image: %.4f ms/step, e2e %.0f img/s, stage launches %s" % (d["from_image"]["ms_per_step"], d["e2e"]["value"], d["from_image"]["stage_per_launch_ms"]))
except Exception as e: print(f, "ERR", e, open(f"gpurun_out/bench_{f}.json").read()[-800:])
PY
}
run tc --head tc
run tc3x --head tc3x
