#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench first, then ncu passes of the same command.
# Outputs land in gpurun_out/ (copied to profiles/ by hand after reading them).
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --parity-reads 20000"
OUT=gpurun_out
mkdir -p $OUT
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.log || { echo "plain run failed"; tail -5 $OUT/prof_plain.log; exit 1; }
# launch list of one timed step (my kernels only; 56 launches x 10 sub-batches per step)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mtsv -s 1760 -c 620 --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
# full captures of the heaviest kernels
for K in verify_kernel seed_search_kernel coalesce_kernel select_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 32 -c 2 -f -o $OUT/prof_$K \
      $CMD > $OUT/ncu_$K.log 2>&1
done
ls -la $OUT
