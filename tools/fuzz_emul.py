#!/usr/bin/env python
"""CPU twin of tools/fuzz_gpu.py: the per-item device logic (csrc/core.cuh compiled with g++, tests/emul/) against
the oracle on the same randomized adversarial cases.  Needs no GPU; covers the arithmetic (k-mer table incl. direct
entries, seed rule, windows, bounded edit distance, the SW re-check of long reads, selection), not the kernels'
parallel structure.   usage: python tools/fuzz_emul.py --seconds 600 [--seed N]"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests import emul_api as em  # noqa: E402
from tests.fuzz_cases import long_case, rand_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=int(time.time()))
    a = ap.parse_args()
    rng = random.Random(a.seed)
    t0 = time.time()
    n = n_hits = 0
    while time.time() - t0 < a.seconds:
        case_seed = rng.getrandbits(48)
        crng = random.Random(case_seed)
        kind = "long" if crng.random() < 0.6 else "small"
        ix, reads, p = long_case(crng) if kind == "long" else rand_case(crng)
        h1, o1 = ix.bin_reads(reads, p)
        sa_rate = crng.choice([1, 2, 32])
        if sa_rate > ix.sa_sample_rate:
            sa_rate = 1
        ktab_k = crng.choice([0, 2, 5, 7, 9])
        e = em.EmulIndex(ix, sa_rate=sa_rate, ktab_k=ktab_k)
        cat, off = po.pack_seqs(reads)
        h2, o2 = e.bin_reads(cat, off, p)
        ok = np.array_equal(o1, o2) and all(np.array_equal(h1[f], h2[f]) for f in ("tax_id", "gi", "offset", "edit"))
        if not ok:
            print("MISMATCH campaign seed %d case seed %d kind %s sa_rate %d ktab_k %d" % (a.seed, case_seed, kind, sa_rate, ktab_k))
            sys.exit(1)
        n += 1
        n_hits += len(h1)
    print("fuzz_emul: %d cases in %.0f s, %d hits compared, campaign seed %d: all bit-exact" % (n, time.time() - t0, n_hits, a.seed))


if __name__ == "__main__":
    main()
