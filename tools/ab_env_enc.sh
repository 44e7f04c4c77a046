#!/bin/bash
# Dev tool: sweep an environment knob of the library over the device-resident encode bench (run on the GPU box via gpurun).
# usage: tools/ab_env_enc.sh "VAR=value [VAR2=value2]" ...     ("-" = no knob)
for kv in "$@"; do
  if [ "$kv" = "-" ]; then kv=""; fi
  echo "[$kv] $(env $kv python tools/enc_bench_dev.py 4096 ${FRAMES:-50} 2>&1 | grep -E 'best|rror' | head -2 | tr '\n' ' ')"
done
