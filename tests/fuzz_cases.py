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


def heavy_case(rng):
    """Many near-identical strains under few TaxIDs: strands with dozens to hundreds of candidates — the warp-level
    selection with its hash set, the (strand, TaxID) group leaders of the two-round verification, both limits."""
    n_strains = rng.choice([20, 40, 80, 160, 300, 600])
    base_len = rng.randint(200, 900)
    base = bytes(rng.choice(b"ACGT") for _ in range(base_len))
    per_tax = rng.choice([1, 2, 5, 13])
    seqs, tax = [], []
    for i in range(n_strains):
        seqs.append(mutate(rng, base, rng.randint(0, max(1, base_len // 60))))
        tax.append(100 + (i // per_tax if rng.random() < 0.9 else rng.randint(0, n_strains)))
    order = list(range(n_strains))
    rng.shuffle(order)  # strains of one TaxID are not neighbours (the builder re-orders bins by TaxID anyway)
    seqs = [seqs[i] for i in order]
    tax = [tax[i] for i in order]
    ix = po.Index.build(seqs, list(range(1000, 1000 + n_strains)), tax, 64, rng.choice([4, 32]))
    L0 = rng.choice([40, 64, 75, 100, 150])
    rate = rng.choice([0.05, 0.13, 0.2])
    reads = []
    for _ in range(rng.randint(1, 12)):
        st = rng.randrange(0, max(1, base_len - L0))
        s = mutate(rng, base[st:st + L0 + 6], int(rng.random() * 1.5 * rate * L0))[:L0]
        if rng.random() < 0.5:
            s = bytes({65: 84, 67: 71, 71: 67, 84: 65}.get(c, c) for c in reversed(s))
        reads.append(s)
    p = po.default_params(edit_rate=rate, seed_size=rng.choice([12, 18]), seed_gap=rng.choice([5, 15]),
                          min_seed=rng.choice([0.015, 0.5]), max_hits=rng.choice([50, 2000, 100000]),
                          tune_max_hits=rng.choice([10, 200, 100000]), max_candidates=rng.choice([-1, -1, 0, 7, 40, 500]),
                          max_assignments=rng.choice([-1, -1, 0, 1, 9, 64]))
    return ix, reads, p
