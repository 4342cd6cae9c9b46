"""The mtsv-binner command-line driver (mtsv_tools_b200/csrc/mtsv_binner_main.cpp): reference flag surface,
FASTA/FASTQ(.gz) parsing, exit codes, results text and resume — src/bin/mtsv-binner.rs, src/binner.rs."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mtsv_tools_b200", "bin", "mtsv-binner")


def _run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True)


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mtsv_tools_b200", "csrc")])


def test_fastx_parsing(tmp_path):
    fa = tmp_path / "t.fa"
    fa.write_text(">r1 desc\nACGT\nACGT\n>r2\nNNNN\n\n>r3 x\nAC\n")
    out = _run("--dump-reads", "--fasta", str(fa))
    assert out.returncode == 0 and out.stdout == "r1\tACGTACGT\nr2\tNNNN\nr3\tAC\n"
    fq = tmp_path / "t.fq"
    fq.write_text("@q1 d\nACGT\n+\nIIII\n@q2\nAC\nGT\n+q2\n@III\n@q3\n\n+\n\n@q4\nA\n+\nI\n")
    want = "q1\tACGT\nq2\tACGT\nq3\t\nq4\tA\n"
    assert _run("--dump-reads", "--fastq", str(fq)).stdout == want
    gz = tmp_path / "t.fq.gz"
    with gzip.open(gz, "wt") as f:  # src/binner.rs:474-499: gz and plain read identically
        f.write(fq.read_text())
    assert _run("--dump-reads", "--fastq", str(gz)).stdout == want
    assert _run("--dump-reads", "--fastq", str(gz), "--read-offset", "2").stdout == "q3\t\nq4\tA\n"
    bad = tmp_path / "bad.fq"
    bad.write_text("@q1\nACGT\n+\nII\n")
    assert _run("--dump-reads", "--fastq", str(bad)).returncode == 12  # src/binner.rs:81-84


def test_exit_codes_without_gpu(tmp_path):
    fa = tmp_path / "t.fa"
    fa.write_text(">r1\nACGT\n")
    assert _run("--fasta", str(fa), "--index", "/nonexistent").returncode == 3  # no results path (:262-265)
    r = _run("--fasta", str(fa), "--index", "/nonexistent", "--results", str(tmp_path / "r.txt"))
    assert r.returncode == 2 and "cannot open" in r.stderr  # query error (:319-322)
    assert _run("--fasta", str(fa), "--index", "x", "--results", "y", "--edit-rate", "1.5").returncode == 101
    assert _run("--fasta", str(fa), "--index", "x", "--results", "y", "--min-seed", "0").returncode == 101


@pytest.mark.gpu
def test_cli_end_to_end(oracle, tmp_path):
    from mtsv_tools_b200 import synth
    cat, off, gi, tax = synth.make_reference(8, 20000, seed=1, n_frac=0.002, shared_frac=0.1, seqs_per_taxid=2)
    ix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    index_path = str(tmp_path / "ref.index")
    ix.write(index_path)
    rc, ro = synth.make_reads(cat, off, 3000, 150, seed=2)
    names = ["read_%d" % i for i in range(3000)]
    fq = tmp_path / "reads.fq.gz"
    with gzip.open(fq, "wt") as f:
        for i, n in enumerate(names):
            s = bytes(rc[int(ro[i]):int(ro[i + 1])]).decode()
            f.write("@%s some description\n%s\n+\n%s\n" % (n, s, "I" * len(s)))
    hits, offs = ix.bin_reads((rc, ro), oracle.default_params(), threads=4)
    for long in (False, True):
        res = tmp_path / ("res_%d.txt" % long)
        args = ["--fastq", str(fq), "--index", index_path, "--results", str(res), "--threads", "8"]
        if long:
            args += ["--output-format", "long"]
        r = _run(*args)
        assert r.returncode == 0, r.stderr
        want = oracle.results_lines(names, hits, offs, long)
        assert sorted(res.read_text().splitlines(True)) == sorted(want)
    # resume: results holding the first 1000 reads' lines -> the run continues after the last id present
    res = tmp_path / "resume.txt"
    first = oracle.results_lines(names[:1000], hits[:int(offs[1000])], offs[:1001], False)
    res.write_text("".join(first))
    r = _run("--fastq", str(fq), "--index", index_path, "--results", str(res))
    assert r.returncode == 0, r.stderr
    last_present = max(i for i in range(1000) if offs[i + 1] > offs[i])
    tail = oracle.results_lines(names[last_present + 1:], hits, offs[last_present + 1:], False)
    assert res.read_text() == "".join(first) + "".join(tail)
    # --force-overwrite starts over; parameters are honoured
    r = _run("--fastq", str(fq), "--index", index_path, "--results", str(res), "--force-overwrite",
             "--edit-rate", "0.05", "--seed-interval", "10", "--max-candidates", "2")
    assert r.returncode == 0, r.stderr
    h2, o2 = ix.bin_reads((rc, ro), oracle.default_params(edit_rate=0.05, seed_gap=10, max_candidates=2), threads=4)
    assert res.read_text() == "".join(oracle.results_lines(names, h2, o2, False))
