#!/bin/bash
# Dev tool: A/B the decoder stages over library variants in build_variants/ (run on the GPU box via gpurun).
# usage: tools/ab_dec.sh variant...
for v in "$@"; do
  CB200_LIB=/root/repo/build_variants/$v.so python bench.py --seconds 6 --steps 3 --warmup 2 --no-e2e --no-cpu --no-encode > gpurun_out/ab_dec_$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_dec_$v.log").read().strip().splitlines()[-1])
    print("$v", round(d["value"]), {k:round(v["ms_per_launch"],1) for k,v in d["roofline"]["stages"].items()})
except Exception as e:
    print("$v", "FAILED", e)
PY
done
