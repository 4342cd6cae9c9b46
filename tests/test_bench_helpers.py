"""Host-side helpers of bench.py that only ever run on a GPU box otherwise: the `.index` header reader, the FASTQ
writer of the CLI leg, the workload description both arms must share."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mtsv_tools_b200 import synth  # noqa: E402


def test_index_file_header_reader(oracle, tmp_path):
    cat, off, gi, tax = synth.make_reference(5, 3000, seed=8, n_frac=0.01)
    tax = np.array([9, 2, 2, 7, 1], dtype=np.uint32)
    ix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    path = str(tmp_path / "x.index")
    ix.write(path)
    text, bins, ref_off = bench.index_file_text_and_bins(path)
    assert np.array_equal(np.array(text), ix.text)
    g, t, s, e = ix.bins()
    assert np.array_equal(bins["gi"], g) and np.array_equal(bins["tax"], t)
    assert np.array_equal(bins["start"], s) and np.array_equal(bins["end"], e)
    assert np.array_equal(ref_off, np.concatenate([s, e[-1:]]))


def test_fastq_writer_round_trips_through_the_cli_parser(tmp_path):
    n, L = 1234, 150
    rng = np.random.default_rng(3)
    reads = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=n * L)]
    path = str(tmp_path / "r.fq")
    names = bench.write_fastq(path, reads, n, L)
    assert names[0] == "read_000000000" and names[-1] == "read_%09d" % (n - 1)
    exe = os.path.join(ROOT, "mtsv_tools_b200", "bin", "mtsv-binner")
    out = subprocess.run([exe, "--dump-reads", "--fastq", path], capture_output=True, text=True)
    assert out.returncode == 0
    lines = out.stdout.splitlines()
    assert len(lines) == n
    for i in (0, 1, 617, n - 1):
        name, seq = lines[i].split("\t")
        assert name == "read_%09d" % i and seq.encode() == reads[i * L:(i + 1) * L].tobytes()
    gz = str(tmp_path / "r.fq.gz")
    bench.write_fastq(gz, reads, n, L, gz=True)
    assert subprocess.run([exe, "--dump-reads", "--fastq", gz], capture_output=True, text=True).stdout == out.stdout


def test_both_arms_describe_the_same_workload():
    for name, cfg in bench.CONFIGS.items():
        a = bench.config_dict(name, cfg, cfg["reads"])
        b = bench.config_dict(name, dict(cfg), cfg["reads"])
        assert a == b and a["workload"] == cfg["label"] and a["reads_per_gpu_per_step"] == cfg["reads"]
    # every config the default run attaches exists
    for k in ("cfg1", "cfg2", "cfg4", "cfg4b", "cfg5"):
        assert k in bench.CONFIGS
