#!/bin/bash
# Dev tool: decode bench (device-resident) over IR budgets per chunk buffer.  usage: SECONDS_PER_STREAM=30 tools/ab_dec_ir.sh MB...
for v in "$@"; do
  CB200_IR_MB=$v python bench.py --seconds ${SECONDS_PER_STREAM:-30} --steps 2 --warmup 2 --no-e2e --no-cpu --no-encode --no-mixed --no-parity > gpurun_out/ab_ir.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_ir.log").read().strip().splitlines()[-1])
    print("IR_MB=$v", round(d["value"]), {k:(v["launches"], round(v["ms_per_launch"],1)) for k,v in d["roofline"]["stages"].items()})
except Exception as e:
    print("IR_MB=$v", "FAILED", e)
PY
done
