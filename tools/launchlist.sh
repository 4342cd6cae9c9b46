#!/bin/bash
# launch list only (one timed step), prints per-kernel totals
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --parity-reads 20000"
OUT=gpurun_out
MYK=$(grep "^MYK=" tools/profile.sh | head -1 | cut -d"'" -f2)
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.log || { echo "plain run failed"; tail -5 $OUT/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MYK" -s 2400 -c 760 --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/launches.csv
