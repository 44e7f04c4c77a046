#!/bin/bash
# Round-end evidence on one GPU box (run via gpurun): plain bench lines, the ncu launch list of the same command, and one
# `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/; tools/summarize_ncu.py turns the reports into profiles/.
set -x
tag=${1:-r1}
timeout 600 python bench.py > gpurun_out/bench_${tag}.log 2> gpurun_out/bench_${tag}.err || exit 1
timeout 600 python bench.py --impl reference > gpurun_out/bench_${tag}_ref.log 2> gpurun_out/bench_${tag}_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
    python bench.py --no-cpu --no-parity > gpurun_out/${tag}_bench_under_ncu.log 2>&1
for k in parse_kernel synth_kernel deemph_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 1 --launch-count 1 -f -o gpurun_out/${tag}_dec_$k \
    python bench.py --streams 4096 --seconds 3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-encode --no-parity --base 16 > gpurun_out/${tag}_dec_$k.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:encode_span_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/${tag}_enc \
    python tools/enc_bench.py 2072 4 > gpurun_out/${tag}_enc.log 2>&1
tail -c 600 gpurun_out/bench_${tag}_ref.log
