#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
TPSPP_LIB=$PWD/tps_pp_b200/libtpspp_tl.so python scripts/dev/timeline.py > gpurun_out/timeline.log 2>&1; sed -n 1,30p gpurun_out/timeline.log
timeout 600 python -m pytest tests/test_head_gpu.py -q -m gpu --timeout=300 --no-header -x 2>&1 | tail -3
run() { name=$1; shift; timeout 900 env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline $HEADARG 2>&1 | tail -1 > gpurun_out/bench_$name.json
python - $name <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{f}.json").read())
    pl=d["roofline_dominant"]["per_launch_ms"]
    print(f, "%.4f ms"%d["ms_per_step"], {k:pl[k] for k in ("down0_1","enc0","enc1","dec1","dec2","dec3")})
except Exception as e: print(f, "ERR", e, open(f"gpurun_out/bench_{f}.json").read()[-800:])
PY
}
HEADARG="--head tc" run pair_mix X=1
HEADARG="--head tc3x" run pair_3x X=1
HEADARG="--head tc" run single_mix TPSPP_CONV_PAIR=0
