"""Host mirror of ``mtsv-build`` (MGIndex::new, src/index.rs:491-582) over the C ABI.

``MGIndex.build`` (index.py) is the product path: reference sequences in, device-resident index out, with
the suffix array and BWT built by the library's own kernels (csrc/sufsort.cu, csrc/build.cu) and never
leaving the GPU.  The helpers here expose the intermediate fields of an MGIndex (text, bins, BWT,
row-sampled suffix array) as numpy arrays for the parity tests and for feeding the CPU oracle the very
index the GPU uses; `concat_reference` is the host restatement of the bin layout rule.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

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
    text = np.empty(total + 1, dtype=np.uint8)
    if np.array_equal(order, np.arange(len(order))):
        text[:total] = cat[int(off[0]):int(off[0]) + total]
    else:
        for k, o in enumerate(order):
            text[int(starts[k]):int(ends[k])] = cat[int(off[o]):int(off[o + 1])]
    text[:total] = _NORM[text[:total]]
    text[total] = ord("$")
    return text, (gi[order], taxid[order], starts, ends)


def suffix_array(text_u8, device=0, want_bwt=False):
    """``suffix_array(&seq)`` (src/index.rs:560) of a '$'-terminated uint8 text, on the GPU
    (mtsvgpu_suffix_array).  Returns the uint32 suffix array, and the BWT when asked."""
    L = _lib.load_library()
    text = np.ascontiguousarray(text_u8, dtype=np.uint8)
    n = len(text)
    sa = np.empty(n, dtype=np.uint32)
    bwt = np.empty(n, dtype=np.uint8) if want_bwt else None
    check(L.mtsvgpu_suffix_array(device, C.c_void_p(text.ctypes.data), n, C.c_void_p(sa.ctypes.data),
                                 C.c_void_p(bwt.ctypes.data) if want_bwt else None))
    return (sa, bwt) if want_bwt else sa


def build_index_parts(cat, off, gi, taxid, sa_sample=32, device=0, verbose=False):
    """Returns dict(text, bins, bwt, sa_sample, sa_rate) as numpy arrays — the fields of an MGIndex that
    mtsvgpu_index_from_parts / the oracle's from_parts take.  Suffix array and BWT come from the GPU."""
    if isinstance(device, str):
        device = int(device.split(":")[1]) if ":" in device else 0
    text, bins = concat_reference(cat, off, gi, taxid)
    sa, bwt = suffix_array(text, device=device, want_bwt=True)
    sample = sa[::sa_sample].astype(np.uint64)
    return dict(text=text, bins=bins, bwt=bwt, sa_sample=sample, sa_rate=sa_sample)
