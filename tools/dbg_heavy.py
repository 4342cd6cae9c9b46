import sys, torch, numpy as np
sys.path.insert(0, ".")
import bench
from mtsv_tools_b200 import synth
from oracle import pyoracle
cfg = bench.CONFIGS["cfg2"]
parts = bench.get_index_parts("cfg2", cfg, "cuda:0", 0, 1, lambda: None)
oix = pyoracle.Index.from_parts(parts["text"], parts["bins"], parts["bwt"], parts["sa_sample"], 32)
ref_t = torch.from_numpy(parts["text"][:-1]).cuda()
n = 200000
d = synth.make_reads_torch(ref_t, parts["ref_off"], n, 150, 4, "cuda:0").cpu().numpy().reshape(n, 150)
ncount = (d == ord("N")).sum(1)
print("reads with >=10 N:", int((ncount >= 10).sum()), "of", n)
idx = np.nonzero(ncount >= 10)[0][:400]
rows = []
for i in idx:
    for strand in (0, 1):
        seq = bytes(d[i]) if strand == 0 else synth.revcomp(bytes(d[i]))
        c = pyoracle.Counters()
        oix.matching_tax_ids(seq, pyoracle.default_params(), c)
        rows.append((c.rows_located, c.candidates, int(ncount[i])))
rows = np.array(rows)
print("per strand rows_located: mean %.0f median %.0f p90 %.0f max %d" % (rows[:,0].mean(), np.median(rows[:,0]), np.percentile(rows[:,0], 90), rows[:,0].max()))
print("per strand candidates:   mean %.0f median %.0f p90 %.0f max %d" % (rows[:,1].mean(), np.median(rows[:,1]), np.percentile(rows[:,1], 90), rows[:,1].max()))
# all reads sample: distribution of hits per strand
c_all = []
for i in range(3000):
    c = pyoracle.Counters(); oix.matching_tax_ids(bytes(d[i]), pyoracle.default_params(), c); c_all.append(c.rows_located)
c_all = np.array(c_all)
print("all reads fwd strand rows_located: mean %.1f p50 %d p90 %d p99 %d max %d; frac>16: %.3f" % (c_all.mean(), np.median(c_all), np.percentile(c_all,90), np.percentile(c_all,99), c_all.max(), (c_all>16).mean()))
