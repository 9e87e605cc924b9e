#!/bin/bash
# Round-2 quick GPU pass: all GPU tests, smoke, both bench arms.  Logs -> gpurun_out/.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=600 --no-header -rA -s 2>&1 | grep -v "^$" | tail -120 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
