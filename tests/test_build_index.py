"""The GPU/torch index builder (mtsv_tools_b200/build_index.py) produces the same MGIndex fields as the
oracle's restated mtsv-build (SA-IS).  Runs on CPU tensors here; the GPU suite re-runs it on cuda."""
import random

import numpy as np
import pytest
import torch

from mtsv_tools_b200 import synth
from mtsv_tools_b200.build_index import build_index_parts


def _check(oracle, cat, off, gi, tax, device):
    ix = oracle.Index.build((cat, off), gi, tax, 64, 32)
    parts = build_index_parts(cat, off, gi, tax, 32, device=device)
    assert np.array_equal(parts["text"], ix.text)
    assert np.array_equal(parts["bwt"], ix.bwt)
    assert np.array_equal(parts["sa_sample"], ix.sa_sample)
    g, t, s, e = ix.bins()
    for a, b in zip(parts["bins"], (g, t, s, e)):
        assert np.array_equal(a, b)
    # an oracle index assembled from the parts answers queries identically
    ix2 = oracle.Index.from_parts(parts["text"], parts["bins"], parts["bwt"], parts["sa_sample"], 32)
    reads = synth.make_reads(cat, off, 300, 100, seed=5)
    h1, o1 = ix.bin_reads(reads, oracle.default_params())
    h2, o2 = ix2.bin_reads(reads, oracle.default_params())
    assert np.array_equal(h1, h2) and np.array_equal(o1, o2)


def test_builder_matches_oracle_cpu(oracle):
    cat, off, gi, tax = synth.make_reference(6, 5000, seed=3, n_frac=0.01, shared_frac=0.5, divergence=0.002)
    tax = np.array([9, 3, 3, 7, 1, 9], dtype=np.uint32)  # forces the TaxID reordering
    _check(oracle, cat, off, gi, tax, "cpu")
    # highly repetitive text: many doubling rounds
    rep = np.frombuffer((b"ACGTACGTAC" * 300 + b"NNNNNNNNNN" * 20 + b"A" * 500), dtype=np.uint8).copy()
    _check(oracle, rep, np.array([0, 2000, len(rep)], dtype=np.uint64), [1, 2], [5, 4], "cpu")


@pytest.mark.gpu
def test_builder_matches_oracle_gpu(oracle):
    cat, off, gi, tax = synth.make_reference(20, 100000, seed=4, n_frac=0.001, shared_frac=0.1)
    _check(oracle, cat, off, gi, tax, "cuda")
