"""Adversarial small cases for the differential tests: tandem repeats, near-copies, N-rich references,
reads of length 0..120 with indels / lower case / garbage, and extreme parameters (seed 3..24, gap 1..20,
edit rate 0..1, min_seed up to 1, tiny max-hits / tune-max-hits, max_candidates / max_assignments 0..3)."""
from oracle import pyoracle as po


def rand_case(rng):
    nseq = rng.randint(1, 6)
    alpha = rng.choice([b"ACGT", b"ACGTN", b"AC", b"ACGTNNNN"])
    seqs = []
    motif = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(3, 40)))
    for _ in range(nseq):
        L = rng.randint(30, 600)
        mode = rng.random()
        if mode < 0.3:
            s = (motif * (L // len(motif) + 1))[:L]           # tandem repeats
            s = bytearray(s)
            for _ in range(rng.randint(0, 5)): s[rng.randrange(L)] = rng.choice(b"ACGT")
            s = bytes(s)
        elif mode < 0.5 and seqs:
            s = bytearray(rng.choice(seqs))                   # near copy of an earlier sequence
            for _ in range(rng.randint(0, 6)): s[rng.randrange(len(s))] = rng.choice(b"ACGTN")
            s = bytes(s)
        else:
            s = bytes(rng.choice(alpha) for _ in range(L))
        seqs.append(s)
    tax = [rng.randint(1, 3) for _ in range(nseq)]
    gi = list(range(10, 10 + nseq))
    ix = po.Index.build(seqs, gi, tax, rng.choice([1, 7, 64]), rng.choice([1, 3, 32]))
    text = bytes(ix.text)
    reads = []
    for _ in range(rng.randint(1, 40)):
        L = rng.randint(0, 120)
        if rng.random() < 0.75 and len(text) > 5:
            st = rng.randrange(0, max(1, len(text) - 2)); s = bytearray(text[st:st + L].replace(b"$", b"A"))
            for _ in range(rng.randint(0, 6)):
                if not s: break
                i = rng.randrange(len(s)); r = rng.random()
                if r < 0.5: s[i] = rng.choice(b"ACGTNacgtnx")
                elif r < 0.75: s.insert(i, rng.choice(b"ACGT"))
                else: del s[i]
            s = bytes(s)
            if rng.random() < 0.5: s = bytes({65:84,67:71,71:67,84:65}.get(c, c) for c in reversed(s))
        else:
            s = bytes(rng.choice(b"ACGTN") for _ in range(L))
        reads.append(s)
    p = po.default_params(edit_rate=rng.choice([0.0, 0.05, 0.13, 0.2, 0.34, 0.5, 0.7, 1.0]),
                          seed_size=rng.choice([rng.randint(3, 24)] * 6 + [40, 64, 65, 70, 100]), seed_gap=rng.randint(1, 20),
                          min_seed=rng.choice([0.015, 0.2, 0.5, 1.0]), max_hits=rng.choice([1, 3, 20, 2000]),
                          tune_max_hits=rng.choice([0, 1, 2, 10, 200]),
                          max_candidates=rng.choice([-1, -1, 0, 1, 3]), max_assignments=rng.choice([-1, -1, 0, 1, 2]))
    return ix, reads, p

