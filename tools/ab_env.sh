#!/bin/bash
# Dev tool: sweep an environment knob of the library over the decode bench (run on the GPU box via gpurun).
# usage: tools/ab_env.sh VAR value...        (E2E=1 tools/ab_env.sh ... also times the host-buffer path, 60 s streams)
var=$1; shift
for v in "$@"; do
  if [ -n "$E2E" ]; then args="--steps 1 --warmup 1 --no-cpu --no-encode --no-mixed --no-parity"; else args="--seconds 12 --steps 2 --warmup 2 --no-e2e --no-cpu --no-encode --no-mixed --no-parity"; fi
  env $var=$v python bench.py $args > gpurun_out/ab_env.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_env.log").read().strip().splitlines()[-1])
    print("$var=$v", round(d["value"]), "e2e", d.get("e2e") and round(d["e2e"]["value"]), {k:round(v["ms_total"],1) for k,v in d["roofline"]["stages"].items()})
except Exception as e:
    print("$var=$v", "FAILED", e)
PY
done
