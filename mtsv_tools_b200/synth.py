"""Deterministic synthetic references and reads (numpy only).

Shapes follow BASELINE.md §4 / SURVEY.md §8(d): i.i.d. ACGT genomes with a small fraction of N
runs, optional shared (diverged) segments between genomes, reads sampled from the reference with
substitutions / indels / N and a fraction of random reads, half of them reverse-complemented.
The reference ships no data generator (its tests use `random_database`, src/index.rs:604-642);
this module plays that role for the tests and for bench.py.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def make_reference(n_seqs, seq_len, seed, n_frac=0.001, shared_frac=0.0, divergence=0.01,
                   taxids=None, seqs_per_taxid=1):
    """Returns (cat uint8, off uint64[n_seqs+1], gi uint32, taxid uint32)."""
    rng = np.random.default_rng(seed)
    total = n_seqs * seq_len
    cat = _ACGT[rng.integers(0, 4, size=total, dtype=np.uint8)]
    if shared_frac > 0 and n_seqs > 1:
        seg = max(1, int(seq_len * shared_frac))
        for i in range(1, n_seqs):
            src = int(rng.integers(0, i))
            so = int(rng.integers(0, seq_len - seg + 1))
            do = int(rng.integers(0, seq_len - seg + 1))
            piece = cat[src * seq_len + so: src * seq_len + so + seg].copy()
            nmut = int(seg * divergence)
            if nmut:
                pos = rng.integers(0, seg, size=nmut)
                piece[pos] = _ACGT[rng.integers(0, 4, size=nmut, dtype=np.uint8)]
            cat[i * seq_len + do: i * seq_len + do + seg] = piece
    if n_frac > 0:
        n_target = int(total * n_frac)
        placed = 0
        while placed < n_target:
            run = int(rng.integers(10, 51))
            p = int(rng.integers(0, max(1, total - run)))
            cat[p:p + run] = ord("N")
            placed += run
    off = (np.arange(n_seqs + 1, dtype=np.uint64) * np.uint64(seq_len))
    gi = np.arange(1, n_seqs + 1, dtype=np.uint32)
    if taxids is None:
        taxids = 1000 + (np.arange(n_seqs, dtype=np.uint32) // np.uint32(seqs_per_taxid))
    taxids = np.asarray(taxids, dtype=np.uint32)
    return cat, off, gi, taxids


def make_reads(ref_cat, ref_off, n_reads, read_len, seed, frac_ref=0.9, sub=0.02, ins=0.0025,
               dele=0.0025, frac_n_reads=0.01, rc_frac=0.5):
    """Returns (cat uint8, off uint64[n_reads+1]); all reads have exactly read_len bases."""
    rng = np.random.default_rng(seed)
    n_ref = int(n_reads * frac_ref)
    n_seqs = len(ref_off) - 1
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    # random reads
    reads[n_ref:] = _ACGT[rng.integers(0, 4, size=(n_reads - n_ref, read_len), dtype=np.uint8)]
    if n_ref:
        span = read_len + 16  # slack so deletions can be compensated
        which = rng.integers(0, n_seqs, size=n_ref)
        lens = (ref_off[1:] - ref_off[:-1]).astype(np.int64)[which]
        usable = np.maximum(lens - span, 1)
        start = ref_off[:-1].astype(np.int64)[which] + (rng.random(n_ref) * usable).astype(np.int64)
        idx = start[:, None] + np.arange(span, dtype=np.int64)[None, :]
        idx = np.minimum(idx, len(ref_cat) - 1)
        src = ref_cat[idx]  # (n_ref, span)
        # substitutions
        m = rng.random((n_ref, span)) < sub
        src[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        out = src[:, :read_len].copy()
        # indels: applied row-wise only to the (few) affected reads
        p_indel = ins + dele
        if p_indel > 0:
            n_ev = rng.binomial(read_len, p_indel, size=n_ref)
            rows = np.nonzero(n_ev)[0]
            for r in rows:
                seq = list(src[r])
                for _ in range(int(n_ev[r])):
                    pos = int(rng.integers(1, read_len - 1))
                    if rng.random() < ins / p_indel:
                        seq.insert(pos, int(_ACGT[rng.integers(0, 4)]))
                    else:
                        del seq[pos]
                out[r] = np.asarray(seq[:read_len], dtype=np.uint8)
        # reverse complement
        rc = rng.random(n_ref) < rc_frac
        out[rc] = _COMP[out[rc][:, ::-1]]
        reads[:n_ref] = out
    # Ns in a fraction of all reads
    if frac_n_reads > 0:
        rows = np.nonzero(rng.random(n_reads) < frac_n_reads)[0]
        for r in rows:
            k = int(rng.integers(1, 4))
            reads[r, rng.integers(0, read_len, size=k)] = ord("N")
    perm = rng.permutation(n_reads)
    reads = reads[perm]
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return reads.reshape(-1), off


def revcomp(seq: bytes) -> bytes:
    return _COMP[np.frombuffer(seq, dtype=np.uint8)[::-1]].tobytes()
