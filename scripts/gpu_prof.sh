#!/bin/bash
# ncu launch list (+ optional full capture of kernels matching $1) for the bench step.
# usage: gpu_prof.sh [kernel-regex [skip [count]]]   (env HEAD=tc|fp32|library)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
HEAD=${HEAD:-tc}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --head $HEAD > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
if [ -n "$1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-4} -c ${3:-2} -f -o gpurun_out/prof_$1 \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --head $HEAD > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log | cut -c1-200
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --head $HEAD 2>&1 | tail -1 | tee gpurun_out/bench.log
