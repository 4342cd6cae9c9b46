import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    print("%s: value %.4g reads/s  e2e %.4g  ms/step %.1f  e2e ms %.1f  cpu %s" % (
        path, d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"].get("ms_per_step", 0),
        d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None))
    st = d.get("stages_ms_per_step", {})
    print("   stages:", {k: round(v, 2) for k, v in st.items() if v}, "sum %.1f" % sum(st.values()))
    print("   ktab", d["config"].get("ktab_k"), "rank sectors/seed %.2f" % (d["work_per_step"]["rank_queries"] / max(1, d["work_per_step"]["n_seed_slots"])))
