import os, sys, numpy as np, torch, time
sys.path.insert(0, os.getcwd())
import bench
from mtsv_tools_b200 import MGIndex, Params
cfg = bench.CONFIGS["cfg2"]
path, _ = bench.ensure_index_file("cfg2", cfg, 0, 0, lambda: None)
g = MGIndex.from_file(path, device=0)
text, _b, ref_off = bench.index_file_text_and_bins(path)
ref_t = torch.from_numpy(np.array(text[:-1])).cuda()
r, o = bench.make_reads(cfg, ref_t, ref_off, 10_000_000, 4, "cuda:0")
h = torch.empty(r.numel(), dtype=torch.uint8, pin_memory=True); h.copy_(r)
ho = torch.empty(o.numel(), dtype=torch.int64, pin_memory=True); ho.copy_(o)
torch.cuda.synchronize()
hr, hoff = h.numpy(), ho.numpy().view(np.uint64)
for i in range(4):
    t = time.perf_counter(); g.bin_reads_pinned((hr, hoff), Params()); print("call %.2f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
