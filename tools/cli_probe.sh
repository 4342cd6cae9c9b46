#!/bin/bash
# GPU box: time the mtsv-binner binary on a plain FASTQ of 2 M reads with per-batch debug lines (-v)
set -u
python - <<'PY'
import numpy as np, sys, os, json, glob
sys.path.insert(0, os.getcwd())
import bench
from mtsv_tools_b200 import synth
cfg = bench.CONFIGS["cfg2"]
path, _ = bench.ensure_index_file("cfg2", cfg, 0, 0, lambda: None)
text, _b, ref_off = bench.index_file_text_and_bins(path)
import torch
ref_t = torch.from_numpy(np.array(text[:-1])).cuda()
r, o = bench.make_reads(cfg, ref_t, ref_off, 2_000_000, 4, "cuda:0")
bench.write_fastq("/tmp/cli_probe.fq", r.cpu().numpy(), 2_000_000, 150)
print(path)
PY
IDX=/tmp/mtsv_b200_cache/cfg2_seed3.index
for t in 16 4; do
  MTSV_B200_TRACE=${TRACE:-} mtsv_tools_b200/bin/mtsv-binner -v --fastq /tmp/cli_probe.fq --index $IDX --results /tmp/cli_probe.res --force-overwrite --threads $t 2>&1 | grep -v "trace\]   " | tail -12
done
