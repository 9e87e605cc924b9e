#!/bin/bash
# timing experiments on conv_tma_kernel (TPSPP_CONV_DBG bits; results are numerically wrong by design)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
for head in tc tc3x; do
for dbg in 0 1 2 4 8 16 32 3 12 14 6 5; do
  TPSPP_CONV_DBG=$dbg timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head $head 2>/dev/null | tail -1 > gpurun_out/dbg.json
  python - "$head" "$dbg" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/dbg.json").read())
    pl=d["roofline_dominant"]["per_launch_ms"]
    print(sys.argv[1], "dbg", sys.argv[2], "step %.3f"%d["ms_per_step"], {k:pl[k] for k in ("down0_1","enc0","enc1","dec2","dec3")})
except Exception as e: print(sys.argv[1], sys.argv[2], "ERR", e)
PY
done; done 2>&1 | tee gpurun_out/conv_dbg.log
