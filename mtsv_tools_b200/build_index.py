"""GPU restatement of ``mtsv-build`` (MGIndex::new, src/index.rs:491-582) for making indexes at
benchmark scale: reference sequences -> text + '$' -> suffix array -> BWT -> row-sampled SA.

The suffix array is built by prefix doubling (Manber & Myers) with ``torch.sort`` as the radix-sort
primitive, so a 1 Gbp index takes seconds on a B200 instead of minutes of single-threaded SA-IS.
This is offline scaffolding for tests and bench.py, not the hot path: the suffix
array of a '$'-terminated text is unique, so the result equals what any correct builder produces
(checked against the oracle's SA-IS in tests/test_build_index.py).  Works on CPU tensors too.
"""
import sys

import numpy as np
import torch

# byte order of the symbols in the text: $ < A < C < G < N < T
_CODE = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate(b"$ACGNT"):
    _CODE[_c] = _i
_NORM = np.full(256, ord("N"), dtype=np.uint8)  # src/index.rs:543-553
for _c in b"ACGTN":
    _NORM[_c] = _c
for _a, _b in zip(b"acgt", b"ACGT"):
    _NORM[_a] = _b


def concat_reference(cat, off, gi, taxid):
    """Bins in TaxID order, file order within a TaxID (parse_fasta_db's BTreeMap, src/io.rs:135-150),
    text normalised to ACGTN and terminated by '$'.  Returns (text uint8, (gi, tax, start, end))."""
    gi = np.asarray(gi, dtype=np.uint32)
    taxid = np.asarray(taxid, dtype=np.uint32)
    off = np.asarray(off, dtype=np.uint64)
    order = np.argsort(taxid, kind="stable")
    lens = (off[1:] - off[:-1])[order]
    starts = np.zeros(len(order), dtype=np.uint64)
    if len(order) > 1:
        starts[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    ends = starts + lens
    total = int(ends[-1]) if len(order) else 0
    if np.array_equal(order, np.arange(len(order))):
        text = np.empty(total + 1, dtype=np.uint8)
        text[:total] = cat[int(off[0]):int(off[0]) + total]
    else:
        text = np.empty(total + 1, dtype=np.uint8)
        for k, o in enumerate(order):
            text[int(starts[k]):int(ends[k])] = cat[int(off[o]):int(off[o + 1])]
    text[:total] = _NORM[text[:total]]
    text[total] = ord("$")
    return text, (gi[order], taxid[order], starts, ends)


def suffix_array(text_u8, device=None, verbose=False):
    """Suffix array (int64 tensor on `device`) of a '$'-terminated uint8 text (numpy)."""
    n = len(text_u8)
    if n >= (1 << 31):
        raise ValueError("this builder handles texts below 2^31 symbols")
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    codes = torch.from_numpy(_CODE[text_u8]).to(device)
    K = 21  # 3 bits per symbol -> 63-bit key
    padded = torch.cat([codes, torch.zeros(K, dtype=torch.uint8, device=device)])
    key = torch.zeros(n, dtype=torch.int64, device=device)
    for j in range(K):
        key.mul_(8).add_(padded[j:j + n])
    del padded, codes
    h = K
    rank = None
    rounds = 0
    while True:
        skey, sa = torch.sort(key)
        del key
        flag = torch.ones(n, dtype=torch.int64, device=device)
        flag[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
        del skey
        grp = torch.cumsum(flag, 0).sub_(1)
        del flag
        ngroups = int(grp[-1].item()) + 1
        if rank is None:
            rank = torch.empty(n, dtype=torch.int64, device=device)
        rank[sa] = grp
        del grp
        rounds += 1
        if verbose:
            print("  suffix_array: h=%d groups=%d/%d" % (h, ngroups, n), file=sys.stderr, flush=True)
        if ngroups == n:
            return sa
        del sa
        # key = (rank[i], rank[i+h] + 1) with 0 past the end
        key = rank << 32
        if h < n:
            key[: n - h] += rank[h:] + 1
        h *= 2


def build_index_parts(cat, off, gi, taxid, sa_sample=32, device=None, verbose=False):
    """Returns dict(text, bins, bwt, sa_sample, sa_rate) as numpy arrays — the fields of an MGIndex that
    mtsvgpu_index_from_parts / the oracle's from_parts take."""
    text, bins = concat_reference(cat, off, gi, taxid)
    n = len(text)
    sa = suffix_array(text, device=device, verbose=verbose)
    t = torch.from_numpy(text).to(sa.device)
    bwt = t[(sa - 1) % n]  # bwt[r] = text[SA[r]-1], text[n-1] when SA[r] == 0 (src/index.rs:567)
    sample = sa[::sa_sample].contiguous()
    out = dict(text=text, bins=bins, bwt=bwt.cpu().numpy(), sa_sample=sample.cpu().numpy().astype(np.uint64),
               sa_rate=sa_sample)
    del sa, t, bwt, sample
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return out
