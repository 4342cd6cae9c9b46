#!/bin/bash
# usage: tools/prof_one.sh <kernel-regex> [skip] : one ncu --set full capture (2 launches) of the default bench
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --parity-reads 20000"
K=$1; S=${2:-32}
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.log || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:^$K -s $S -c 2 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
ls -la gpurun_out/prof_$K.ncu-rep
