"""The packed-read input format (include/mtsv_b200.h, mtsvgpu_bin_batch_packed): host packer against a numpy
statement of the format on the CPU; on the GPU, packed input must give exactly the results of raw input."""
import numpy as np
import pytest

from mtsv_tools_b200 import synth
from mtsv_tools_b200.index import pack_reads, pack_reads_planes


def _pack_numpy(cat, off):
    """Format by the book: per read lo / hi / nn planes of ceil(L/8) bytes, base j at bit j%8 of byte j/8;
    A C G T (either case) = 0 1 2 3, anything else sets nn only."""
    code = np.full(256, 4, np.uint8)
    for i, (u, l) in enumerate(zip(b"ACGT", b"acgt")):
        code[u] = code[l] = i
    out = []
    for r in range(len(off) - 1):
        c = code[cat[int(off[r]):int(off[r + 1])]]
        L = len(c)
        pb = (L + 7) // 8
        planes = np.zeros((3, pb * 8), np.uint8)
        planes[0, :L] = (c < 4) & ((c & 1) != 0)
        planes[1, :L] = (c < 4) & ((c & 2) != 0)
        planes[2, :L] = c >= 4
        out.append(np.packbits(planes.reshape(3, pb, 8), axis=2, bitorder="little").reshape(-1))
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


def _odd_reads(seed, n):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM.-*\x00\xff", dtype=np.uint8)
    reads = []
    for i in range(n):
        L = int(rng.choice([0, 1, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100, 150, 151, 255, 256, 257, 300, 1000]))
        p = np.ones(len(alphabet))
        p[:4] = 20
        reads.append(alphabet[rng.choice(len(alphabet), size=L, p=p / p.sum())].tobytes())
    return reads


def test_packer_matches_the_format():
    for seed in range(3):
        cat, off = pack_reads(_odd_reads(seed, 400))
        want = _pack_numpy(cat, off)
        for threads in (1, 3):
            got, off2 = pack_reads_planes((cat, off), threads=threads)
            assert np.array_equal(off2, off)
            assert np.array_equal(got, want), (seed, threads)
    # uniform lengths take the arithmetic layout; many reads so that several threads get a range each
    ref = synth.make_reference(2, 5000, seed=1)
    cat, off = synth.make_reads(ref[0], ref[1], 20000, 150, seed=2)
    got, _ = pack_reads_planes((cat, off), threads=4)
    assert len(got) == 20000 * 57
    assert np.array_equal(got, _pack_numpy(cat, off))
    # empty batch, a batch of empty reads, an empty read among others
    got, _ = pack_reads_planes((np.zeros(0, np.uint8), np.zeros(1, np.uint64)))
    assert len(got) == 0
    got, off = pack_reads_planes([b"", b""])
    assert len(got) == 0 and off.tolist() == [0, 0, 0]
    got, off = pack_reads_planes([b"", b"ACGTN", b""])
    assert got.tolist() == [0b01010, 0b01100, 0b10000]


def test_device_unpacker_gives_the_encoders_planes(emul):
    """core.cuh::unpack_word (what unpack_reads_kernel runs per read) on a packed record == the bit planes the device
    encodes from the raw bytes (encode_fwd_word), for records from core.cuh::pack_read and from the product's
    AVX2 / SWAR packer; garbage in a record's padding bits and lo/hi bits under an N are ignored."""
    import numpy as np
    for seed in range(2):
        for s in _odd_reads(seed + 20, 300):
            a, b = emul.packed_words(s)
            assert np.array_equal(a, b), (seed, len(s))
            rec, _ = pack_reads_planes([s])
            a2, b2 = emul.packed_words(s, rec)
            assert np.array_equal(a2, b2), (seed, len(s))
            if len(s) % 8 and len(rec):
                dirty = bytearray(rec.tobytes())
                pb = (len(s) + 7) // 8
                for pl in range(3):
                    dirty[pl * pb + pb - 1] |= (0xFF << (len(s) % 8)) & 0xFF  # set every padding bit
                a3, b3 = emul.packed_words(s, bytes(dirty))
                assert np.array_equal(a3, b3)


@pytest.mark.gpu
def test_packed_input_equals_raw_input(oracle):
    from mtsv_tools_b200 import MGIndex, Params
    cat, off, gi, tax = synth.make_reference(8, 30000, seed=21, n_frac=0.002, shared_frac=0.2)
    oix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    rng = np.random.default_rng(5)
    # uniform 150-base reads, ragged reads, garbage / empty / long reads mixed in
    uni = synth.make_reads(cat, off, 6000, 150, seed=22)
    r150 = [uni[0][int(uni[1][i]):int(uni[1][i + 1])].tobytes() for i in range(3000)]
    ragged = [s[: int(rng.integers(20, 151))] for s in r150]
    long_src = synth.make_reads(cat, off, 300, 400, seed=23)
    longs = [long_src[0][int(long_src[1][i]):int(long_src[1][i + 1])].tobytes() for i in range(300)]
    mixed = ragged[:1500] + _odd_reads(9, 200) + longs + [s.lower() for s in r150[:200]]
    order = rng.permutation(len(mixed))
    mixed = [mixed[i] for i in order]
    for name, reads, opts in (("uniform", uni, {}), ("ragged", pack_reads(ragged), {}),
                              ("mixed", pack_reads(mixed), {}), ("mixed-small-slices", pack_reads(mixed), {"batch_reads": 257}),
                              ("uniform-small-slices", uni, {"batch_reads": 1000})):
        with MGIndex.from_parts(oix.text, oix.bins(), oix.bwt, oix.sa_sample, oix.sa_sample_rate, device=0, **opts) as g:
            want_h, want_o = g.bin_reads(reads, Params())
            packed, poff = pack_reads_planes(reads)
            got_h, got_o = g.bin_reads_packed(packed, poff, Params())
            assert np.array_equal(got_o, want_o), name
            for f in ("tax_id", "gi", "offset", "edit"):
                assert np.array_equal(got_h[f], want_h[f]), (name, f)
            if name == "uniform":
                oh, oo = oix.bin_reads(reads, oracle.default_params(), threads=4)
                assert np.array_equal(got_o, oo) and np.array_equal(got_h["edit"], oh["edit"]) and len(oh) > 3000
                st = g.last_batch_stats()
                assert st["h2d_bytes"] == 6000 * 57  # planes only: offsets of equal-length slices are generated on the device
