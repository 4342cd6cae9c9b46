import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from mtsv_tools_b200 import MGIndex, Params, synth
cfg = dict(bench.CONFIGS["cfg2"], seed=5)
parts = bench.get_index_parts("cfg2_chunk0", cfg, "cuda:0", 0, 1, lambda: None)
g = MGIndex.from_parts(parts["text"], parts["bins"], parts["bwt"], parts["sa_sample"], 32)
ref_t = torch.from_numpy(parts["text"][:-1]).cuda()
for n in (1000000, 3000000):
    d = synth.make_reads_torch(ref_t, parts["ref_off"], n, 150, 4, "cuda:0")
    off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 150
    g.set_profiling(True)
    try:
        r = g.bin_reads_device(d.data_ptr(), off.data_ptr(), n, Params())
        print(n, "ok hits", r[2], g.last_batch_stats())
    except Exception as e:
        print(n, "ERR", e)
