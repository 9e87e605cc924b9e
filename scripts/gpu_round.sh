#!/bin/bash
# Round-end evidence run on one B200: all GPU tests, smoke, both bench arms, ncu launch list and full captures
# of the warp kernel (HBM roofline) and the largest tensor-core conv (tensor-pipe utilisation).
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu --timeout=600 --no-header 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== bench fp32 head"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --head fp32 2>&1 | tail -1 | tee gpurun_out/bench_fp32.json
echo "== bench library head"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --head library 2>&1 | tail -1 | tee gpurun_out/bench_library.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: warp kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_staged -s 3 -c 1 -f -o gpurun_out/prof_warp \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_warp.log 2>&1
echo "== bench bf16 conv operands"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --head bf16 2>&1 | tail -1 | tee gpurun_out/bench_bf16.json
echo "== ncu full: the 11 TMA-staged convolutions of one step (enc0 = 7th) and the 4 linear-layer launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 33 -c 11 -f -o gpurun_out/prof_conv_tma \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lin_tma_kernel|mlp_fused_kernel|dgab_warp_kernel|loc_p1_kernel" -s 15 -c 5 -f -o gpurun_out/prof_lin \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_lin.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv | tee gpurun_out/smi_end.txt
