#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench first, then ncu passes of the same command.
# Usage: [CONFIG=cfg2] [TAG=r02] tools/profile.sh [kernel ...]   (default: launch list + verify_warp_kernel + seed_search_kernel)
# Each kernel is captured for the three launches of ONE step (after the parity-gate launch and one warm-up step).
set -u
CONFIG=${CONFIG:-cfg2}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --parity-reads 20000 --only-main --config $CONFIG"
OUT=gpurun_out
mkdir -p $OUT
$CMD > $OUT/prof_plain_$CONFIG.json 2> $OUT/prof_plain_$CONFIG.log || { echo "plain run failed"; tail -5 $OUT/prof_plain_$CONFIG.log; exit 1; }
# launch list of about two steps: this library's kernels only (the build / parity-gate / warm-up launches come first)
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(?!.*(at::|elementwise|vectorized|distribution|index|cunn|Philox|rs_|sfx_|ms_|fm_|ktab|sa_|text_pack|normalise|bwt_|occ_|scan_))" \
    -s ${LAUNCH_SKIP:-300} -c ${LAUNCH_COUNT:-260} --csv --log-file $OUT/launches_$CONFIG.csv $CMD > $OUT/ncu_launches_$CONFIG.log 2>&1
KERNELS="${@:-verify_warp_kernel seed_search_kernel}"
for K in $KERNELS; do
  ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed.sum \
      --clock-control none --import-source on -k regex:^$K -s ${SKIP:-4} -c ${COUNT:-3} -f -o $OUT/prof_${CONFIG}_$K \
      $CMD > $OUT/ncu_${CONFIG}_$K.log 2>&1
done
ls -la $OUT | grep -i "prof_\|launches" | head -30
