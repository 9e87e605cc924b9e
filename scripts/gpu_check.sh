#!/bin/bash
# One GPU-box pass: tests, smoke, bench, ncu launch list.  Logs go to gpurun_out/.
# usage: scripts/gpu_check.sh [quick|full|prof]
mode=${1:-full}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python - <<'PY' > gpurun_out/env.txt 2>&1
import os, torch
print("cores", os.cpu_count(), "torch", torch.__version__, "cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0))
PY
if [ "$mode" != head ]; then
echo "== pytest -m gpu (warp)"; timeout 900 python -m pytest tests/test_warp_gpu.py -q -m gpu --timeout=300 -x --no-header -rA 2>&1 | tail -60 | tee gpurun_out/pytest_warp.log
fi
echo "== pytest -m gpu (head)"; timeout 400 python -m pytest tests/test_head_gpu.py -q -m gpu --timeout=300 --no-header -rA -s 2>&1 | tail -80 | tee gpurun_out/pytest_head.log
[ "$mode" = head ] && exit 0
echo "== pytest -m gpu (module)"; timeout 900 python -m pytest tests/test_module_gpu.py -q -m gpu --timeout=300 --no-header -rA -s 2>&1 | tail -60 | tee gpurun_out/pytest_module.log
if [ "$mode" = quick ]; then exit 0; fi
if [ "$mode" = head ]; then exit 0; fi
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -5 | tee gpurun_out/bench.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.log
if [ "$mode" = prof ]; then
  echo "== ncu launches"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
  echo "== ncu full (warp kernel)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_staged -s 3 -c 2 -f -o gpurun_out/warp_staged \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
