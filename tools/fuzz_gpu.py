#!/usr/bin/env python
"""Randomized differential campaign on a GPU box: the CUDA path through the C ABI against the oracle (which runs
the reference's own ssw.c) on adversarial small cases.  Widens tests/fuzz_cases.py: read lengths 0..420 (all verifier
widths, uniform and ragged batches, the >= 254-base SW re-check), tiny device sub-batches (many slices, both lanes,
adaptive launch groups).   usage: python tools/fuzz_gpu.py --seconds 300 [--seed N]"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mtsv_tools_b200 import MGIndex, Params  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from mtsv_tools_b200.index import pack_reads_planes  # noqa: E402
from tests.fuzz_cases import heavy_case, long_case, rand_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=int(time.time()))
    a = ap.parse_args()
    rng = random.Random(a.seed)
    t0 = time.time()
    n = n_hits = n_long = n_heavy = 0
    while time.time() - t0 < a.seconds:
        case_seed = rng.getrandbits(48)
        crng = random.Random(case_seed)
        u = crng.random()
        kind = "long" if u < 0.45 else ("heavy" if u < 0.65 else "small")
        ix, reads, p = long_case(crng) if kind == "long" else (heavy_case(crng) if kind == "heavy" else rand_case(crng))
        # two-round verification: left to the library's own rule, forced on, forced off
        gv = crng.choice([None, "1", "0"])
        if gv is None:
            os.environ.pop("MTSV_B200_GROUP_VERIFY", None)
        else:
            os.environ["MTSV_B200_GROUP_VERIFY"] = gv
        h1, o1 = ix.bin_reads(reads, p)
        pg = Params(edit_rate=p.edit_rate, seed_size=p.seed_size, seed_gap=p.seed_gap, min_seed=p.min_seed,
                    max_hits=p.max_hits, tune_max_hits=p.tune_max_hits,
                    max_candidates=None if p.max_candidates < 0 else p.max_candidates,
                    max_assignments=None if p.max_assignments < 0 else p.max_assignments)
        opts = dict(sa_rate=crng.choice([1, 2, 32]), ktab_k=crng.choice([0, 0xFFFFFFFF, 2, 5]),
                    batch_reads=crng.choice([0, 0, 1, 3, 7, 16]), max_batch_hits=crng.choice([0, 0, 50]))
        if opts["sa_rate"] > ix.sa_sample_rate:
            opts["sa_rate"] = 1
        with MGIndex.from_parts(ix.text, ix.bins(), ix.bwt, ix.sa_sample, ix.sa_sample_rate, **opts) as g:
            for api in ("bin_reads", "bin_reads_pinned", "bin_reads_packed"):
                try:
                    if api == "bin_reads_packed":
                        pk, pko = pack_reads_planes(reads)
                        h2, o2 = g.bin_reads_packed(pk, pko, pg)
                    else:
                        h2, o2 = getattr(g, api)(po.pack_seqs(reads) if api == "bin_reads_pinned" else reads, pg)
                except Exception as e:  # a refused case (limits) must be refused loudly, not silently wrong
                    if "cap" in str(e) or "limit" in str(e).lower():
                        continue
                    raise
                ok = np.array_equal(o1, o2) and all(np.array_equal(h1[f], h2[f]) for f in ("tax_id", "gi", "offset", "edit"))
                if not ok:
                    print("MISMATCH campaign seed %d case seed %d kind %s api %s opts %s" % (a.seed, case_seed, kind, api, opts))
                    sys.exit(1)
        n += 1
        n_hits += len(h1)
        n_long += kind == "long"
        n_heavy += kind == "heavy"
    print("fuzz_gpu: %d cases (%d long, %d heavy) x 3 entry points in %.0f s, %d hits compared, campaign seed %d: all bit-exact"
          % (n, n_long, n_heavy, time.time() - t0, n_hits, a.seed))


if __name__ == "__main__":
    main()
