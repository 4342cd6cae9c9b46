"""Generates tests/golden/cfg1_small.npz: a scaled-down BASELINE config 1 (8 sequences x 20 kbp,
2000 x 150 bp reads, default flags) run through the oracle.  Run from the repo root:
    python -m tests.golden.make_golden
The reference itself (Rust) cannot run in this image, so the fixture records the oracle's output;
the oracle is pinned separately by tests/test_oracle.py."""
import os

import numpy as np


def build_case(oracle):
    from mtsv_tools_b200 import synth
    cat, off, gi, tax = synth.make_reference(8, 20000, seed=1, n_frac=0.002, shared_frac=0.1,
                                             seqs_per_taxid=2)
    ix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    reads = synth.make_reads(cat, off, 2000, 150, seed=2)
    return ix, reads, oracle.default_params()


def main():
    from oracle import pyoracle
    ix, reads, params = build_case(pyoracle)
    hits, offs = ix.bin_reads(reads, params, threads=4)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg1_small.npz")
    np.savez_compressed(out, hit_off=offs, tax_id=hits["tax_id"], gi=hits["gi"], offset=hits["offset"],
                        edit=hits["edit"])
    print("wrote", out, "reads", len(offs) - 1, "hits", len(hits))


if __name__ == "__main__":
    main()
