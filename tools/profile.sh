#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench first, then ncu passes of the same command.
# Usage: tools/profile.sh [kernel ...]   (default: launch list + verify_kernel + seed_search_kernel)
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --parity-reads 20000"
OUT=gpurun_out
mkdir -p $OUT
MYK='regex:^(count_slots|expand_slots|encode_reads|encode_fwd|seed_search|seed_select|locate|sort_classify|sort_warp|sort_medium|sort_large|coalesce|coalesce_heavy|coalesce_monster|rank_emit|cand_class|cand_order|verify|verify_warp|select|gather_hits)_kernel|^scan_(tile_sums|sums_inplace|apply|empty)'
$CMD > $OUT/prof_plain.json 2> $OUT/prof_plain.log || { echo "plain run failed"; tail -5 $OUT/prof_plain.log; exit 1; }
# launch list of about two steps (this library's kernels only; ~37 launches x 3 sub-batches per step; the parity
# gate and the warm-up steps come first and do the same work per step)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MYK" -s 150 -c 230 --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
KERNELS="${@:-verify_warp_kernel seed_search_kernel}"
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:^$K -s 4 -c 2 -f -o $OUT/prof_$K \
      $CMD > $OUT/ncu_$K.log 2>&1
done
ls -la $OUT | head -30
