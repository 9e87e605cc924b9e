#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
for dbg in 0 1 16 2 4 8 6 22 ; do
  TPSPP_CONV_DBG=$dbg timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --head tc 2>/dev/null | tail -1 > gpurun_out/dbg.json
  python - "$dbg" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/dbg.json").read())
    pl=d["roofline_dominant"]["per_launch_ms"]
    print("pair dbg", sys.argv[1], "step %.3f"%d["ms_per_step"], {k:pl[k] for k in ("down0_1","enc0","enc1","dec2","dec3")})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done 2>&1 | tee gpurun_out/pair_dbg.log
