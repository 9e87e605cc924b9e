#!/bin/bash
# Round-end evidence run on one B200: all GPU tests, smoke, both bench arms, the secondary workloads, the training step, ncu
# launch list and full captures of the warp kernel (HBM roofline) and the 3x3 convolutions.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout=600 --no-header 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke.log
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== bench tc3x (all-tf32 3xTF32 convolutions)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --head tc3x 2>&1 | tail -1 > gpurun_out/bench_tc3x.json
echo "== bench bf16 (operands + storage)"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --head bf16 2>&1 | tail -1 > gpurun_out/bench_bf16.json
echo "== bench classical"; for F in 20 40; do timeout 900 python bench.py --workload classical64x256 --F $F --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_classical_F$F.json; done
echo "== bench global batch 2048 (configs[4] shard at N=1)"; timeout 900 python bench.py --global-batch 2048 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_gb2048.json
echo "== training step (native, CUDA graph) + per-kernel table; library arm"
timeout 600 python scripts/train_step_bench.py --profile 2>&1 | grep -v Warning | tail -75 > gpurun_out/train_kernels.txt; tail -1 gpurun_out/train_kernels.txt > gpurun_out/train_native.json
timeout 600 python scripts/train_step_bench.py --library-convs 2>&1 | tail -1 > gpurun_out/train_library.json
echo "== NRTR greedy decode (native incremental decode vs the reference's algorithm on cuBLAS fp32)"
timeout 600 python scripts/nrtr_decode_bench.py 256 64 2>&1 | tail -1 > gpurun_out/nrtr_decode.json
timeout 600 python scripts/nrtr_decode_bench.py 1024 64 2>&1 | tail -1 > gpurun_out/nrtr_decode_b1024.json
echo "== warp backward per kernel"; timeout 300 python scripts/warp_bwd_profile.py 2>&1 | grep -v Warn | tail -8 > gpurun_out/warp_bwd_kernels.txt
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: warp kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warp_fwd_staged -s 3 -c 1 -f -o gpurun_out/prof_warp \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_warp.log 2>&1
echo "== ncu full: the convolutions of one rectifier step (3x3: down0_1, down1_1, enc0, enc1, dec1, dec2, dec3) and the fused kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tma_kernel|down_fused_kernel" -s 24 -c 8 -f -o gpurun_out/prof_conv_tma \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_conv.log 2>&1
echo "== ncu full: bf16 mode convolutions"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tma_kernel" -s 24 -c 8 -f -o gpurun_out/prof_conv_bf16 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --head bf16 > gpurun_out/ncu_full_conv_bf16.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv | tee gpurun_out/smi_end.txt
