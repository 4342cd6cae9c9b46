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
from tests.fuzz_cases import rand_case  # noqa: E402


def mutate(rng, s, n):
    s = bytearray(s)
    for _ in range(n):
        if not s:
            break
        i = rng.randrange(len(s))
        r = rng.random()
        if r < 0.45:
            s[i] = rng.choice(b"ACGTNacgtnx")
        elif r < 0.75:
            for _ in range(rng.randint(1, 3)):
                s.insert(i, rng.choice(b"ACGT"))
        else:
            del s[i:i + rng.randint(1, 3)]
    return bytes(s)


def long_case(rng):
    """Bigger references and reads of up to 420 bases, uniform or ragged."""
    nseq = rng.randint(1, 5)
    seqs = []
    for _ in range(nseq):
        L = rng.randint(400, 3000)
        s = bytes(rng.choice(b"ACGT") for _ in range(L))
        if seqs and rng.random() < 0.4:
            s = mutate(rng, rng.choice(seqs), rng.randint(0, 30))
        if rng.random() < 0.3:
            p = rng.randrange(len(s) - 60)
            s = s[:p] + b"N" * rng.randint(5, 50) + s[p + 50:]
        seqs.append(s)
    ix = po.Index.build(seqs, list(range(10, 10 + nseq)), [rng.randint(1, 3) for _ in range(nseq)],
                        rng.choice([7, 64]), rng.choice([3, 32]))
    text = bytes(ix.text)
    uniform = rng.random() < 0.5
    L0 = rng.choice([30, 64, 65, 100, 128, 129, 150, 192, 193, 250, 253, 254, 256, 257, 300, 420])
    rate = rng.choice([0.02, 0.05, 0.13, 0.2, 0.3])
    reads = []
    for _ in range(rng.randint(1, 60)):
        L = L0 if uniform else rng.randint(0, 420)
        if rng.random() < 0.8:
            st = rng.randrange(0, max(1, len(text) - 2))
            s = text[st:st + L + 12].replace(b"$", b"A")
            s = mutate(rng, s, int(rng.random() * 1.3 * rate * L))[:L]
            if uniform and len(s) < L:
                s = s + bytes(rng.choice(b"ACGT") for _ in range(L - len(s)))
            if rng.random() < 0.5:
                s = bytes({65: 84, 67: 71, 71: 67, 84: 65}.get(c, c) for c in reversed(s))
        else:
            s = bytes(rng.choice(b"ACGTN") for _ in range(L))
        reads.append(s)
    p = po.default_params(edit_rate=rate, seed_size=rng.choice([12, 18, 18, 24]), seed_gap=rng.choice([3, 7, 15]),
                          min_seed=rng.choice([0.015, 0.3]), max_hits=rng.choice([20, 2000]),
                          tune_max_hits=rng.choice([2, 200]), max_candidates=rng.choice([-1, -1, 2]),
                          max_assignments=rng.choice([-1, -1, 1]))
    return ix, reads, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=int(time.time()))
    a = ap.parse_args()
    rng = random.Random(a.seed)
    t0 = time.time()
    n = n_hits = n_long = 0
    while time.time() - t0 < a.seconds:
        case_seed = rng.getrandbits(48)
        crng = random.Random(case_seed)
        kind = "long" if crng.random() < 0.6 else "small"
        ix, reads, p = long_case(crng) if kind == "long" else rand_case(crng)
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
            for api in ("bin_reads", "bin_reads_pinned"):
                try:
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
    print("fuzz_gpu: %d cases (%d long) in %.0f s, %d hits compared, campaign seed %d: all bit-exact"
          % (n, n_long, time.time() - t0, n_hits, a.seed))


if __name__ == "__main__":
    main()
