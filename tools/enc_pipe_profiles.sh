#!/bin/bash
# One `ncu --set full` capture per kernel of the encoder pipeline (run via gpurun).  usage: tools/enc_pipe_profiles.sh <tag> [kernels...]
tag=${1:-r2}; shift
ks=${@:-pipe_prepass_kernel pipe_fe2_kernel pipe_comb_kernel pipe_transform_kernel pipe_decide_kernel pipe_prep_kernel pipe_spec_kernel pipe_leaves_kernel pipe_exact_kernel}
for k in $ks; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip ${SKIP:-12} --launch-count 1 -f -o gpurun_out/${tag}_$k \
    python tools/enc_bench.py 4096 20 > gpurun_out/${tag}_$k.log 2>&1
done
