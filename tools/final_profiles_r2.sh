#!/bin/bash
# Round-2 evidence on one GPU box (run via gpurun, two calls: `bench` and `ncu`).  Everything lands in gpurun_out/ (<= 64 MiB per call:
# the ncu reports are summarised ON the box and only the two dominant kernels' reports travel back).
#   tools/final_profiles_r2.sh bench r2f   plain bench lines of both arms + the ncu launch list of the same bench command
#   tools/final_profiles_r2.sh ncu r2f     one `ncu --set full` capture per hot kernel (decoder stages, every encoder pipeline kernel)
set -x
what=${1:-bench}; tag=${2:-r2f}
if [ "$what" = bench ]; then
  timeout 900 python bench.py > gpurun_out/bench_${tag}.log 2> gpurun_out/bench_${tag}.err || exit 1
  timeout 600 python bench.py --impl reference > gpurun_out/bench_${tag}_ref.log 2> gpurun_out/bench_${tag}_ref.err
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
      python bench.py --no-cpu --no-parity --no-mixed > gpurun_out/${tag}_bench_under_ncu.log 2>&1
  exit 0
fi
summarise() {  # report title command keep
  python tools/summarize_ncu.py $1.ncu-rep "$2" "$3" > $1_ncu_summary.md 2> /dev/null
  python tools/ncu_by_line.py $1.ncu-rep 40 > $1_by_function.md 2> /dev/null
  if [ "$4" != keep ]; then rm -f $1.ncu-rep; fi
}
DEC="python bench.py --streams 4096 --seconds 3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-encode --no-parity --no-mixed"
for k in parse_kernel synth_kernel deemph_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 1 --launch-count 1 -f -o gpurun_out/${tag}_dec_$k \
    $DEC > gpurun_out/${tag}_dec_$k.log 2>&1
  keep=no; [ $k = parse_kernel ] && keep=keep
  summarise gpurun_out/${tag}_dec_$k "ncu --set full, decoder $k (round-2 final build), B200" "$DEC   (4,096 streams x 150 packets in one launch per stage; second launch captured)" $keep
done
ENC="python tools/enc_bench_dev.py 4096 16"
for k in pipe_walk_kernel pipe_transform_kernel pipe_prep_kernel pipe_decide_kernel pipe_transient2_kernel pipe_comb_kernel pipe_head_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 20 --launch-count 1 -f -o gpurun_out/${tag}_enc_$k \
    $ENC > gpurun_out/${tag}_enc_$k.log 2>&1
  keep=no; [ $k = pipe_walk_kernel ] && keep=keep
  summarise gpurun_out/${tag}_enc_$k "ncu --set full, encoder pipeline $k (round-2 final build), B200" "$ENC   (4,096 streams, one frame step; launch 21 of the kernel captured)" $keep
done
for k in pipe_prepass2_kernel pipe_fe1_kernel pipe_fe2_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 2 --launch-count 1 -f -o gpurun_out/${tag}_enc_$k \
    $ENC > gpurun_out/${tag}_enc_$k.log 2>&1
  summarise gpurun_out/${tag}_enc_$k "ncu --set full, encoder pipeline $k (round-2 final build), B200" "$ENC   (4,096 streams x 8 frames per launch; launch 3 captured)" no
done
du -sh gpurun_out
