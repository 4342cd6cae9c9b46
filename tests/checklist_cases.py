"""Constructed cases for SURVEY appendix C (parity checklist): the situations a random generator reaches only by
luck — seeds straddling two bins, reads longer than a bin (windows that are None in the middle of a run), equal
seed counts at two loci of one TaxID and of different TaxIDs, the same TaxID reachable on both strands with
different edits, tandem repeats that chain-merge windows, min_seeds >= 2, reads of S-1 / S / 0 bases."""
import random

import numpy as np

from mtsv_tools_b200 import synth
from oracle import pyoracle as po


def _rnd(rng, n):
    return bytes(rng.choice(b"ACGT") for _ in range(n))


def _mut(rng, s, n):
    s = bytearray(s)
    for _ in range(n):
        s[rng.randrange(len(s))] = rng.choice(b"ACGT")
    return bytes(s)


def build():
    rng = random.Random(2026)
    a = _rnd(rng, 1200)
    b = _rnd(rng, 900)
    short1, short2 = _rnd(rng, 90), _rnd(rng, 110)          # bins shorter than the reads
    dup = _rnd(rng, 700)
    dup_mut = _mut(rng, dup, 6)                                # near copy: same TaxID as dup -> ties / first-in-rank
    dup_other = _mut(rng, dup, 3)                              # near copy under another TaxID
    rc_src = _rnd(rng, 600)
    rc_copy = _mut(rng, synth.revcomp(rc_src), 5)              # same TaxID reachable on the other strand
    motif = _rnd(rng, 23)
    tandem = (motif * 40)[:800]                                # windows chain-merge far beyond L + 2k
    nrun = _rnd(rng, 300) + b"N" * 40 + _rnd(rng, 300)
    seqs = [a, short1, b, short2, dup, dup_mut, dup_other, rc_src, rc_copy, tandem, nrun]
    tax = [1, 2, 3, 2, 4, 4, 5, 6, 6, 7, 8]
    gi = list(range(100, 100 + len(seqs)))
    ix = po.Index.build(seqs, gi, tax, 64, 32)
    text = bytes(ix.text)
    gi_a, tax_a, st, en = ix.bins()
    starts = {int(g): int(s) for g, s in zip(gi_a, st)}
    ends = {int(g): int(e) for g, e in zip(gi_a, en)}
    reads = []

    def cut(p, L, nmut=0):
        s = text[p:p + L].replace(b"$", b"A")
        return _mut(rng, s, nmut) if nmut else s
    # seeds straddling two bins / reads spanning a junction, at every phase of the seed grid
    for g in gi:
        for d in range(-170, 30, 7):
            p = ends[g] + d
            if 0 <= p < len(text) - 160:
                reads.append(cut(p, 150, rng.choice([0, 0, 2])))
    # reads longer than the bins they start in (None windows between valid ones)
    for g in (101, 103):
        for d in range(-60, 60, 5):
            p = starts[g] + d
            if p >= 0:
                reads.append(cut(p, 150, rng.choice([0, 1])))
    # the duplicated family and the strand pair, forward and reverse-complemented, a few edits
    for g in (104, 105, 106, 107, 108):
        for _ in range(25):
            p = starts[g] + rng.randrange(0, ends[g] - starts[g] - 150)
            r = cut(p, 150, rng.choice([0, 1, 3, 8]))
            reads.append(r if rng.random() < 0.5 else synth.revcomp(r))
    # tandem repeats, N run
    for g in (109, 110):
        for _ in range(30):
            p = starts[g] + rng.randrange(0, ends[g] - starts[g] - 150)
            reads.append(cut(p, 150, rng.choice([0, 2])))
    # degenerate lengths around the seed size (18): 0, 1, 16, 17 (the reference panics below 17), 18, 19
    for L in (0, 1, 16, 17, 18, 19, 33):
        reads.append(cut(starts[100] + 50, L))
    reads.append(cut(starts[100] + 300, 150))                # and a duplicate of the next one (duplicate read ids are the CLI's business)
    reads.append(cut(starts[100] + 300, 150))
    flag_sets = [{}, dict(min_seed=0.4), dict(min_seed=1.0), dict(max_candidates=1), dict(max_candidates=2, max_assignments=1),
                 dict(tune_max_hits=1, max_hits=3), dict(tune_max_hits=3, max_hits=40), dict(edit_rate=0.0),
                 dict(edit_rate=0.51), dict(seed_size=12, seed_gap=5, min_seed=0.3), dict(seed_size=18, seed_gap=1)]
    return ix, reads, flag_sets
