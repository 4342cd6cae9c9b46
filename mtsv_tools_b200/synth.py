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


def _make_reads_chunk(ref_cat, ref_off, n_reads, read_len, rng, frac_ref, sub, ins, dele, frac_n_reads,
                      rc_frac):
    n_ref = int(n_reads * frac_ref)
    n_seqs = len(ref_off) - 1
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    reads[n_ref:] = _ACGT[rng.integers(0, 4, size=(n_reads - n_ref, read_len), dtype=np.uint8)]
    if n_ref:
        span = read_len + 8  # slack so deletions can be compensated
        which = rng.integers(0, n_seqs, size=n_ref)
        starts0 = ref_off[:-1].astype(np.int64)
        lens = (ref_off[1:] - ref_off[:-1]).astype(np.int64)[which]
        usable = np.maximum(lens - span, 1)
        start = starts0[which] + (rng.random(n_ref) * usable).astype(np.int64)
        idx = start[:, None] + np.arange(span, dtype=np.int64)[None, :]
        np.minimum(idx, len(ref_cat) - 1, out=idx)
        src = ref_cat[idx]  # (n_ref, span)
        del idx
        # substitutions
        m = rng.random((n_ref, span), dtype=np.float32) < sub
        src[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        del m
        # indels, vectorised: two rounds of "at most one event per read" so that the expected number of
        # events per read is read_len * (ins + dele)
        p_indel = ins + dele
        cols = np.arange(span, dtype=np.int64)[None, :]
        for _ in range(2):
            if p_indel <= 0:
                break
            ev = rng.random(n_ref) < min(1.0, read_len * p_indel / 2.0)
            is_ins = rng.random(n_ref) < (ins / p_indel)
            pos = rng.integers(1, read_len - 1, size=n_ref)[:, None]
            # deletion at pos: out[j] = src[j + (j >= pos)] ; insertion at pos: out[j] = src[j - (j > pos)]
            shift = np.where((ev & ~is_ins)[:, None], (cols >= pos).astype(np.int64),
                             np.where((ev & is_ins)[:, None], -(cols > pos).astype(np.int64), 0))
            gather = np.clip(cols + shift, 0, span - 1)
            src = np.take_along_axis(src, gather, axis=1)
            rows = np.nonzero(ev & is_ins)[0]
            src[rows, pos[rows, 0]] = _ACGT[rng.integers(0, 4, size=len(rows), dtype=np.uint8)]
        out = np.ascontiguousarray(src[:, :read_len])
        rc = rng.random(n_ref) < rc_frac
        out[rc] = _COMP[out[rc][:, ::-1]]
        reads[:n_ref] = out
    if frac_n_reads > 0:
        rows = np.nonzero(rng.random(n_reads) < frac_n_reads)[0]
        for k in range(3):
            sel = rows[rng.random(len(rows)) < (1.0 if k == 0 else 0.5)]
            reads[sel, rng.integers(0, read_len, size=len(sel))] = ord("N")
    return reads[rng.permutation(n_reads)]


def make_reads(ref_cat, ref_off, n_reads, read_len, seed, frac_ref=0.9, sub=0.02, ins=0.0025,
               dele=0.0025, frac_n_reads=0.01, rc_frac=0.5, chunk=1 << 19):
    """Returns (cat uint8, off uint64[n_reads+1]); all reads have exactly read_len bases.
    frac_ref of the reads are sampled from the reference (rc_frac of those reverse-complemented) with
    per-base substitution / insertion / deletion rates; frac_n_reads get 1-3 N; the rest are random."""
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    ref_off = np.asarray(ref_off)
    for ci, b in enumerate(range(0, n_reads, chunk)):
        e = min(n_reads, b + chunk)
        rng = np.random.default_rng([seed, ci])
        out[b:e] = _make_reads_chunk(ref_cat, ref_off, e - b, read_len, rng, frac_ref, sub, ins, dele,
                                     frac_n_reads, rc_frac)
    off = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return out.reshape(-1), off


def revcomp(seq: bytes) -> bytes:
    return _COMP[np.frombuffer(seq, dtype=np.uint8)[::-1]].tobytes()


def make_reads_torch(ref_cat_t, ref_off, n_reads, read_len, seed, device, frac_ref=0.9, sub=0.02,
                     ins=0.0025, dele=0.0025, frac_n_reads=0.01, rc_frac=0.5, chunk=1 << 20):
    """Same read model as make_reads, generated with torch on `device` (10M reads in seconds on a
    GPU).  ref_cat_t: uint8 tensor of the concatenated reference on `device`.  Returns a uint8
    tensor (n_reads * read_len) on `device`; offsets are i * read_len."""
    import torch
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.from_numpy(_COMP).to(device)
    starts0 = torch.from_numpy(np.asarray(ref_off[:-1]).astype(np.int64)).to(device)
    lens0 = torch.from_numpy((np.asarray(ref_off[1:]) - np.asarray(ref_off[:-1])).astype(np.int64)).to(device)
    n_seqs = len(ref_off) - 1
    out = torch.empty((n_reads, read_len), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    for ci, b in enumerate(range(0, n_reads, chunk)):
        e = min(n_reads, b + chunk)
        m = e - b
        g.manual_seed(seed * 1000003 + ci)
        n_ref = int(m * frac_ref)
        reads = torch.empty((m, read_len), dtype=torch.uint8, device=device)
        reads[n_ref:] = acgt[torch.randint(0, 4, (m - n_ref, read_len), generator=g, device=device)]
        if n_ref:
            span = read_len + 8
            which = torch.randint(0, n_seqs, (n_ref,), generator=g, device=device)
            usable = torch.clamp(lens0[which] - span, min=1)
            start = starts0[which] + (torch.rand(n_ref, generator=g, device=device, dtype=torch.float64)
                                      * usable).to(torch.int64)
            cols = torch.arange(span, device=device, dtype=torch.int64)[None, :]
            idx = torch.clamp(start[:, None] + cols, max=ref_cat_t.numel() - 1)
            src = ref_cat_t[idx]
            del idx
            msk = torch.rand((n_ref, span), generator=g, device=device) < sub
            rnd = acgt[torch.randint(0, 4, (n_ref, span), generator=g, device=device)]
            src = torch.where(msk, rnd, src)
            del msk, rnd
            p_indel = ins + dele
            for _ in range(2):
                if p_indel <= 0:
                    break
                ev = torch.rand(n_ref, generator=g, device=device) < min(1.0, read_len * p_indel / 2.0)
                is_ins = torch.rand(n_ref, generator=g, device=device) < (ins / p_indel)
                pos = torch.randint(1, read_len - 1, (n_ref, 1), generator=g, device=device)
                d_shift = (cols >= pos).to(torch.int64) * (ev & ~is_ins)[:, None]
                i_shift = (cols > pos).to(torch.int64) * (ev & is_ins)[:, None]
                gather = torch.clamp(cols + d_shift - i_shift, 0, span - 1)
                src = torch.gather(src, 1, gather)
                newb = acgt[torch.randint(0, 4, (n_ref, 1), generator=g, device=device)]
                src = torch.where(((cols == pos) & (ev & is_ins)[:, None]), newb, src)
            o = src[:, :read_len].contiguous()
            rc = torch.rand(n_ref, generator=g, device=device) < rc_frac
            o = torch.where(rc[:, None], comp[o.flip(1).to(torch.int64)], o)
            reads[:n_ref] = o
        if frac_n_reads > 0:
            sel = torch.rand(m, generator=g, device=device) < frac_n_reads
            for k in range(3):
                s2 = sel & (torch.rand(m, generator=g, device=device) < (1.0 if k == 0 else 0.5))
                posn = torch.randint(0, read_len, (m,), generator=g, device=device)
                hit = s2[:, None] & (torch.arange(read_len, device=device)[None, :] == posn[:, None])
                reads = torch.where(hit, torch.full_like(reads, ord("N")), reads)
        perm = torch.randperm(m, generator=g, device=device)
        out[b:e] = reads[perm]
    return out.reshape(-1)
