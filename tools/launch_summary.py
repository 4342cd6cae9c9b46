import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mn = hdr.index("Metric Name")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = r[kn].split("(")[0]
    v = float(r[mv].replace(",", ""))
    agg[name][0] += 1; agg[name][1] += v; agg[name][2] = max(agg[name][2], v)
tot = sum(v[1] for v in agg.values())
print("total %.1f ms, launches %d" % (tot / 1e6, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-46s n=%4d  %9.3f ms  %5.1f%%  avg %.3f max %.3f ms" % (k[:46], v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e6, v[2] / 1e6))
