"""mtsv-build on the GPU (csrc/sufsort.cu, csrc/build.cu) produces the same MGIndex fields — and the same
`.index` bytes — as the oracle's restated mtsv-build (SA-IS).  The bin layout rule is host logic and is
checked on the CPU; everything that launches kernels is in the GPU suite."""
import os

import numpy as np
import pytest

from mtsv_tools_b200 import synth
from mtsv_tools_b200.build_index import build_index_parts, concat_reference, suffix_array


def _cases():
    cat, off, gi, tax = synth.make_reference(6, 5000, seed=3, n_frac=0.01, shared_frac=0.5, divergence=0.002)
    tax = np.array([9, 3, 3, 7, 1, 9], dtype=np.uint32)  # forces the TaxID reordering
    yield "reordered", cat, off, gi, tax
    # highly repetitive text: many doubling rounds
    rep = np.frombuffer((b"ACGTACGTAC" * 300 + b"NNNNNNNNNN" * 20 + b"A" * 500), dtype=np.uint8).copy()
    yield "repetitive", rep, np.array([0, 2000, len(rep)], dtype=np.uint64), np.array([1, 2], np.uint32), np.array([5, 4], np.uint32)
    # lower case and IUPAC bytes are folded (src/index.rs:543-553); an empty sequence keeps its (empty) bin
    odd = np.frombuffer(b"acgtnRYKMacgtACGTNNNN" * 50, dtype=np.uint8).copy()
    yield "odd-bytes", odd, np.array([0, 100, 100, len(odd)], dtype=np.uint64), np.array([7, 8, 9], np.uint32), np.array([2, 2, 1], np.uint32)
    yield "tiny", np.frombuffer(b"A", dtype=np.uint8).copy(), np.array([0, 1], np.uint64), np.array([1], np.uint32), np.array([1], np.uint32)


def test_bin_layout_matches_oracle_cpu(oracle):
    for name, cat, off, gi, tax in _cases():
        ix = oracle.Index.build((cat, off), gi, tax, 64, 32)
        text, bins = concat_reference(cat, off, gi, tax)
        assert np.array_equal(text, ix.text), name
        for a, b in zip(bins, ix.bins()):
            assert np.array_equal(a, b), name


def _check_parts(oracle, cat, off, gi, tax):
    ix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    parts = build_index_parts(cat, off, gi, tax, 32, device=0)
    assert np.array_equal(parts["text"], ix.text)
    assert np.array_equal(parts["bwt"], ix.bwt)
    assert np.array_equal(parts["sa_sample"], ix.sa_sample)
    return ix, parts


@pytest.mark.gpu
def test_suffix_array_matches_oracle_gpu(oracle, monkeypatch):
    for name, cat, off, gi, tax in _cases():
        _check_parts(oracle, cat, off, gi, tax)
    cat, off, gi, tax = synth.make_reference(20, 100000, seed=4, n_frac=0.001, shared_frac=0.1)
    ix, parts = _check_parts(oracle, cat, off, gi, tax)
    want = oracle.suffix_array(ix.text)
    # short round-0 keys and tiny slabs: many doubling rounds, many slabs per round, groups larger than a slab
    for k, slab in ((1, 1000), (3, 50), (21, 7), (2, 1 << 30)):
        monkeypatch.setenv("MTSV_B200_SUFSORT_K", str(k))
        monkeypatch.setenv("MTSV_B200_SUFSORT_SLAB", str(slab))
        small = ix.text[-30001:] if k > 1 else ix.text[-6001:]
        got = suffix_array(small)
        assert np.array_equal(got.astype(np.uint64), oracle.suffix_array(small)), (k, slab)
    monkeypatch.delenv("MTSV_B200_SUFSORT_K")
    monkeypatch.delenv("MTSV_B200_SUFSORT_SLAB")
    assert np.array_equal(suffix_array(ix.text).astype(np.uint64), want)


@pytest.mark.gpu
def test_device_build_bins_like_from_parts_and_writes_the_same_index_file(oracle, tmp_path):
    from mtsv_tools_b200 import MGIndex, Params
    cat, off, gi, tax = synth.make_reference(12, 40000, seed=9, n_frac=0.002, shared_frac=0.2)
    tax = (tax[::-1]).copy()  # bins get reordered
    oix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    reads = synth.make_reads(cat, off, 3000, 150, seed=10)
    want_h, want_o = oix.bin_reads(reads, oracle.default_params(), threads=4)
    ref_file = str(tmp_path / "oracle.index")
    oix.write(ref_file)
    for opts in ({}, {"sa_rate": 8}, {"ktab_k": 6}):
        with MGIndex.build(cat, off, gi, tax, device=0, **opts) as g:
            info = g.info()
            assert info["text_len"] == len(oix.text) and info["build_seconds"] > 0
            hits, offs = g.bin_reads(reads, Params())
            assert np.array_equal(offs, want_o)
            for f in ("tax_id", "gi", "offset", "edit"):
                assert np.array_equal(hits[f], want_h[f]), (opts, f)
            # the `.index` it writes is byte-identical to the oracle's bincode dump, whatever the device layout
            out = str(tmp_path / "gpu.index")
            g.write(out, 64, 32)
            assert open(out, "rb").read() == open(ref_file, "rb").read(), opts
    # ... and a handle opened from a file re-serialises to the same bytes (other intervals too)
    with MGIndex.from_file(ref_file, device=0) as g:
        out = str(tmp_path / "again.index")
        g.write(out, 64, 32)
        assert open(out, "rb").read() == open(ref_file, "rb").read()
        oix2 = oracle.Index.build((cat, off), gi, tax, 128, 16)
        oix2.write(ref_file)
        g.write(out, 128, 16)
        assert open(out, "rb").read() == open(ref_file, "rb").read()


@pytest.mark.gpu
def test_index_beyond_2_pow_31_rows(oracle):
    """BASELINE config 3 has 4 Gbp chunks: rows and text positions above 2^31 must work everywhere (builder,
    FM ranks, k-mer table, suffix array, windows).  2.3 Gbp reference built on the device; a read sample drawn
    mostly from the upper half is compared with the oracle working on the same index fields."""
    import torch
    from mtsv_tools_b200 import MGIndex, Params
    free, _ = torch.cuda.mem_get_info(0)
    if free < 120e9:
        pytest.skip("needs ~70 GB of free device memory")
    n_seqs, seq_len = 46, 50_000_000
    total = n_seqs * seq_len
    assert total > (1 << 31)
    g = torch.Generator(device="cuda:0")
    g.manual_seed(77)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda:0")
    cat = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    for b in range(0, total, 1 << 28):
        e = min(total, b + (1 << 28))
        cat[b:e] = acgt[torch.randint(0, 4, (e - b,), generator=g, device="cuda:0")]
    # a segment of the last sequence repeats inside the first one: hits on both sides of 2^31 for one read
    cat[1_000_000:1_200_000] = cat[total - 3_000_000: total - 2_800_000]
    cat[total - 5_000_000: total - 4_999_960] = ord("N")
    off = np.arange(n_seqs + 1, dtype=np.uint64) * np.uint64(seq_len)
    gi = np.arange(1, n_seqs + 1, dtype=np.uint32)
    tax = (100 + np.arange(n_seqs)).astype(np.uint32)
    with MGIndex.build(cat.data_ptr(), off, gi, tax, device=0) as gix:
        info = gix.info()
        assert info["text_len"] == total + 1
        # reads: 4000 from the top 400 Mbp (positions > 2^31), 1000 from the repeated segment, 1000 anywhere
        hi0 = total - 400_000_000
        ref_off = off
        r_hi = synth.make_reads_torch(cat[hi0:], np.array([0, 400_000_000], np.uint64), 4000, 150, 5, "cuda:0")
        r_rep = synth.make_reads_torch(cat[total - 3_000_000: total - 2_800_000], np.array([0, 200_000], np.uint64),
                                       1000, 150, 6, "cuda:0")
        r_any = synth.make_reads_torch(cat, ref_off, 1000, 150, 7, "cuda:0")
        reads = torch.cat([r_hi, r_rep, r_any]).cpu().numpy()
        roff = np.arange(6001, dtype=np.uint64) * np.uint64(150)
        hits, offs = gix.bin_reads((reads, roff), Params())
        # the oracle gets the same fields: text from the host copy, BWT / samples through the written file
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "big.index")
            gix.write(path, 64, 32)
            del cat
            torch.cuda.empty_cache()
            oix = oracle.Index.read(path)
    want_h, want_o = oix.bin_reads((reads, roff), oracle.default_params(), threads=os.cpu_count())
    assert len(want_h) > 4000
    assert int(want_h["offset"].max()) > 0  # (bin-relative; the positions themselves are beyond 2^31)
    assert np.array_equal(offs, want_o)
    for f in ("tax_id", "gi", "offset", "edit"):
        assert np.array_equal(hits[f], want_h[f]), f
