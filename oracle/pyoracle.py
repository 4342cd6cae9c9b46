"""ctypes binding of oracle/liboracle.so.

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs.  Never by mtsv_tools_b200.
See oracle/mtsv_oracle.h for what each entry point restates (reference file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Hit(C.Structure):
    _fields_ = [("tax_id", C.c_uint32), ("gi", C.c_uint32), ("offset", C.c_uint64),
                ("edit", C.c_uint32), ("_pad", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("edit_rate", C.c_double), ("seed_size", C.c_uint32), ("seed_gap", C.c_uint32),
                ("min_seed", C.c_double), ("max_hits", C.c_uint64), ("tune_max_hits", C.c_uint64),
                ("max_candidates", C.c_int64), ("max_assignments", C.c_int64)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "seeds_searched", "bs_steps", "seeds_used", "rows_located", "lf_steps", "candidates",
        "sw_calls", "sw_cells", "ed_calls", "ed_cells", "window_bytes", "hits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


HIT_DTYPE = np.dtype([("tax_id", "<u4"), ("gi", "<u4"), ("offset", "<u8"), ("edit", "<u4"),
                      ("_pad", "<u4")])


def default_params(**kw):
    """Defaults of the mtsv-binner CLI (src/bin/mtsv-binner.rs:68-94)."""
    p = Params(0.13, 18, 15, 0.015, 2000, 200, -1, -1)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mtsv_oracle.cpp", "mtsv_oracle.h", "sais.hpp")]
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    ref_so = os.path.join(_HERE, "_ref", "libssw_ref.so")
    if stale or (not os.path.exists(ref_so) and os.path.exists("/root/reference/ssw/src/ssw.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    vp = C.c_void_p
    L.orc_index_build.restype = vp
    L.orc_index_build.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint64]
    L.orc_index_from_parts.restype = vp
    L.orc_index_from_parts.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, C.c_uint64, vp, vp, C.c_uint64,
                                       C.c_uint64, C.c_uint32]
    L.orc_index_write.argtypes = [vp, C.c_char_p]
    L.orc_index_read.restype = vp
    L.orc_index_read.argtypes = [C.c_char_p]
    L.orc_index_free.argtypes = [vp]
    L.orc_index_len.restype = C.c_uint64
    L.orc_index_len.argtypes = [vp]
    L.orc_index_text.restype = vp
    L.orc_index_text.argtypes = [vp]
    L.orc_index_bwt.restype = vp
    L.orc_index_bwt.argtypes = [vp]
    L.orc_index_nbins.restype = C.c_uint64
    L.orc_index_nbins.argtypes = [vp]
    L.orc_index_bins.argtypes = [vp, vp, vp, vp, vp]
    L.orc_index_sa_sample_rate.restype = C.c_uint64
    L.orc_index_sa_sample_rate.argtypes = [vp]
    L.orc_index_sa_sample_len.restype = C.c_uint64
    L.orc_index_sa_sample_len.argtypes = [vp]
    L.orc_index_sa_sample.restype = vp
    L.orc_index_sa_sample.argtypes = [vp]
    L.orc_index_occ_interval.restype = C.c_uint32
    L.orc_index_occ_interval.argtypes = [vp]
    L.orc_suffix_array.argtypes = [vp, C.c_uint64, vp]
    L.orc_backward_search.argtypes = [vp, vp, C.c_uint64, u64p, u64p, u64p]
    L.orc_occ.restype = C.c_uint64
    L.orc_occ.argtypes = [vp, C.c_uint64, C.c_uint8]
    L.orc_less.restype = C.c_uint64
    L.orc_less.argtypes = [vp, C.c_uint8]
    L.orc_locate.restype = C.c_uint64
    L.orc_locate.argtypes = [vp, C.c_uint64, u64p]
    L.orc_min_edit_distance.restype = C.c_uint32
    L.orc_min_edit_distance.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    L.orc_ssw_score.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_int]
    L.orc_candidate_indices.argtypes = [C.c_uint64] * 6 + [u64p, u64p]
    L.orc_matching_tax_ids.argtypes = [vp, C.c_char_p, C.c_uint64, C.POINTER(Params),
                                       C.POINTER(C.POINTER(Hit)), u64p, C.POINTER(Counters)]
    L.orc_bin_reads.argtypes = [vp, vp, vp, C.c_uint64, C.POINTER(Params), C.c_int,
                                C.POINTER(C.POINTER(Hit)), C.POINTER(u64p), C.POINTER(Counters)]
    L.orc_format_assignments.restype = C.c_int64
    L.orc_format_assignments.argtypes = [C.c_char_p, vp, C.c_uint64, C.c_int, C.c_char_p,
                                         C.c_uint64]
    L.orc_free.argtypes = [vp]
    _LIB = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_seqs(seqs):
    """list of bytes -> (uint8 concat, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    cat = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return cat, off


class Index:
    """Owns an orc_index*."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index handle is NULL")
        self.h = handle

    @classmethod
    def build(cls, seqs, gis, taxids, occ_interval=64, sa_sample=32):
        if isinstance(seqs, tuple):
            cat, off = seqs
        else:
            cat, off = pack_seqs(seqs)
        gi = np.asarray(gis, dtype=np.uint32)
        tx = np.asarray(taxids, dtype=np.uint32)
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return cls(lib().orc_index_build(_ptr(cat), _ptr(off), _ptr(gi), _ptr(tx), len(gi),
                                         occ_interval, sa_sample))

    @classmethod
    def from_parts(cls, text, bins, bwt, sa_sample, sa_rate, occ_interval=64):
        text = np.ascontiguousarray(text, dtype=np.uint8)
        bwt = np.ascontiguousarray(bwt, dtype=np.uint8)
        sa_sample = np.ascontiguousarray(sa_sample, dtype=np.uint64)
        gi, tx, st, en = [np.ascontiguousarray(a, dtype=d) for a, d in
                          zip(bins, (np.uint32, np.uint32, np.uint64, np.uint64))]
        return cls(lib().orc_index_from_parts(_ptr(text), len(text), _ptr(gi), _ptr(tx), _ptr(st), _ptr(en),
                                              len(gi), _ptr(bwt), _ptr(sa_sample), len(sa_sample), sa_rate,
                                              occ_interval))

    @classmethod
    def read(cls, path):
        return cls(lib().orc_index_read(os.fsencode(path)))

    def write(self, path):
        rc = lib().orc_index_write(self.h, os.fsencode(path))
        if rc != 0:
            raise IOError("orc_index_write failed: %d" % rc)

    def __del__(self):
        if getattr(self, "h", None) and _LIB is not None:
            _LIB.orc_index_free(self.h)
            self.h = None

    def __len__(self):
        return int(lib().orc_index_len(self.h))

    def _view(self, ptr, n, dtype):
        buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype)

    @property
    def text(self):
        return self._view(lib().orc_index_text(self.h), len(self), np.uint8)

    @property
    def bwt(self):
        return self._view(lib().orc_index_bwt(self.h), len(self), np.uint8)

    @property
    def sa_sample(self):
        return self._view(lib().orc_index_sa_sample(self.h),
                          int(lib().orc_index_sa_sample_len(self.h)), np.uint64)

    @property
    def sa_sample_rate(self):
        return int(lib().orc_index_sa_sample_rate(self.h))

    @property
    def occ_interval(self):
        return int(lib().orc_index_occ_interval(self.h))

    def bins(self):
        nb = int(lib().orc_index_nbins(self.h))
        gi = np.zeros(nb, np.uint32)
        tx = np.zeros(nb, np.uint32)
        st = np.zeros(nb, np.uint64)
        en = np.zeros(nb, np.uint64)
        lib().orc_index_bins(self.h, _ptr(gi), _ptr(tx), _ptr(st), _ptr(en))
        return gi, tx, st, en

    def backward_search(self, pat):
        lo, up, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
        pat = np.frombuffer(bytes(pat), dtype=np.uint8)
        r = lib().orc_backward_search(self.h, _ptr(pat), len(pat), C.byref(lo), C.byref(up),
                                      C.byref(st))
        return r, lo.value, up.value, st.value

    def locate(self, row):
        lf = C.c_uint64()
        return int(lib().orc_locate(self.h, row, C.byref(lf))), lf.value

    def occ(self, r, a):
        return int(lib().orc_occ(self.h, r, a))

    def less(self, a):
        return int(lib().orc_less(self.h, a))

    def matching_tax_ids(self, seq, params=None, counters=None):
        """One strand of an already normalised read -> list of (tax, gi, offset, edit)."""
        params = params or default_params()
        hp = C.POINTER(Hit)()
        n = C.c_uint64()
        lib().orc_matching_tax_ids(self.h, bytes(seq), len(seq), C.byref(params), C.byref(hp),
                                   C.byref(n), C.byref(counters) if counters is not None else None)
        out = [(hp[i].tax_id, hp[i].gi, hp[i].offset, hp[i].edit) for i in range(n.value)]
        lib().orc_free(hp)
        return out

    def bin_reads(self, seqs, params=None, threads=1, counters=None):
        """Batch, both strands (src/binner.rs:77-131). Returns (hits structured array, offsets)."""
        params = params or default_params()
        if isinstance(seqs, tuple):
            cat, off = seqs
        else:
            cat, off = pack_seqs(seqs)
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        hp = C.POINTER(Hit)()
        op = C.POINTER(C.c_uint64)()
        lib().orc_bin_reads(self.h, _ptr(cat), _ptr(off), n, C.byref(params), threads,
                            C.byref(hp), C.byref(op),
                            C.byref(counters) if counters is not None else None)
        offs = np.ctypeslib.as_array(op, shape=(n + 1,)).copy()
        total = int(offs[-1])
        if total:
            buf = (C.c_uint8 * (total * C.sizeof(Hit))).from_address(C.addressof(hp.contents))
            hits = np.frombuffer(buf, dtype=HIT_DTYPE).copy()
        else:
            hits = np.zeros(0, dtype=HIT_DTYPE)
        lib().orc_free(hp)
        lib().orc_free(op)
        return hits, offs


def suffix_array(text):
    """SA-IS suffix array (uint64) of a uint8 text ending in its unique smallest symbol."""
    t = np.ascontiguousarray(text, dtype=np.uint8)
    sa = np.zeros(len(t), dtype=np.uint64)
    if lib().orc_suffix_array(_ptr(t), len(t), _ptr(sa)) != 0:
        raise RuntimeError("orc_suffix_array failed")
    return sa


def min_edit_distance(p, t):
    return int(lib().orc_min_edit_distance(bytes(p), len(p), bytes(t), len(t)))


def ssw_score(read, ref, kind=0):
    return int(lib().orc_ssw_score(bytes(read), len(read), bytes(ref), len(ref), kind))


def ssw_ref_available():
    return bool(lib().orc_ssw_ref_available())


def candidate_indices(site, q_off, bin_start, bin_end, read_len, k):
    s, e = C.c_uint64(), C.c_uint64()
    ok = lib().orc_candidate_indices(site, q_off, bin_start, bin_end, read_len, k, C.byref(s),
                                     C.byref(e))
    return (s.value, e.value) if ok else None


def format_assignments(header, hits, long_format=False):
    """hits: iterable of (tax, gi, offset, edit) or structured array."""
    if isinstance(hits, np.ndarray) and hits.dtype == HIT_DTYPE:
        arr = np.ascontiguousarray(hits)
    else:
        hits = list(hits)
        arr = np.zeros(len(hits), dtype=HIT_DTYPE)
        for i, h in enumerate(hits):
            arr[i]["tax_id"], arr[i]["gi"], arr[i]["offset"], arr[i]["edit"] = h
    buf = C.create_string_buffer(64 + len(header) + 64 * max(1, len(arr)))
    n = lib().orc_format_assignments(header.encode(), _ptr(arr), len(arr), int(long_format), buf,
                                     len(buf))
    if n < 0:
        raise RuntimeError("format buffer too small")
    return buf.raw[:n].decode()


def results_lines(names, hits, offs, long_format=False):
    """Results text of a batch (one line per read with hits), as write_assignments would emit."""
    out = []
    for i, name in enumerate(names):
        a, b = int(offs[i]), int(offs[i + 1])
        if b > a:
            out.append(format_assignments(name, hits[a:b], long_format))
    return out


def collapse_taxid(parts):
    """mtsv-collapse, mode TaxId, for per-chunk hit lists of the SAME reads (src/collapse.rs:543-654): for
    equal read ids keep the minimum edit per TaxID (:597-602), listed by ascending TaxID
    (write_collapsed_taxid, :278-279).  parts: list of (hits structured array, offsets).  Returns
    (pairs uint32 [n,2], offsets uint64 [n_reads+1])."""
    n_reads = len(parts[0][1]) - 1
    pairs, offs = [], np.zeros(n_reads + 1, dtype=np.uint64)
    for r in range(n_reads):
        best = {}
        for hits, o in parts:
            for h in hits[int(o[r]):int(o[r + 1])]:
                t, e = int(h["tax_id"]), int(h["edit"])
                if t not in best or e < best[t]:
                    best[t] = e
        for t in sorted(best):
            pairs.append((t, best[t]))
        offs[r + 1] = len(pairs)
    return np.array(pairs, dtype=np.uint32).reshape(-1, 2), offs


def collapse_taxid_gi(parts):
    """mtsv-collapse, mode TaxIdGi (src/collapse.rs:603-625): per read and per (TaxID, GI) keep the smallest edit,
    ties to the smallest offset; listed by (TaxID, GI, edit, offset) (write_collapsed_taxid_gi, :311-318).
    parts as in collapse_taxid.  Returns (hits structured array HIT_DTYPE-like, offsets uint64 [n_reads+1])."""
    n_reads = len(parts[0][1]) - 1
    rows, offs = [], np.zeros(n_reads + 1, dtype=np.uint64)
    for r in range(n_reads):
        best = {}
        for hits, o in parts:
            for h in hits[int(o[r]):int(o[r + 1])]:
                key = (int(h["tax_id"]), int(h["gi"]))
                val = (int(h["edit"]), int(h["offset"]))
                if key not in best or val < best[key]:  # :622: edit <, or edit == and offset <
                    best[key] = val
        for key in sorted(best):
            rows.append((key[0], key[1], best[key][1], best[key][0]))
        offs[r + 1] = len(rows)
    out = np.zeros(len(rows), dtype=[("tax_id", "<u4"), ("gi", "<u4"), ("offset", "<u8"), ("edit", "<u4"),
                                     ("reserved", "<u4")])
    for i, (t, g, off, e) in enumerate(rows):
        out[i] = (t, g, off, e, 0)
    return out, offs
