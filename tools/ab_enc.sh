#!/bin/bash
# Dev tool: A/B the encoder kernel over library variants in build_variants/ (run on the GPU box via gpurun).
# usage: tools/ab_enc.sh variant...     ("cur" = the in-tree library)
for v in "$@"; do
  if [ "$v" = cur ]; then lib=/root/repo/concentus_b200/libconcentus_b200.so; else lib=/root/repo/build_variants/$v.so; fi
  echo "$v: $(CB200_LIB=$lib python tools/enc_bench_dev.py 4096 ${FRAMES:-50} 2>&1 | grep -E "best" | tr '\n' ' ')"
done
