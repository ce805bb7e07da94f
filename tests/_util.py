"""Shared helpers of the test-suite: golden fixtures, the oracle, stub embedding functions."""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_oracle as oracle  # noqa: E402  (tests may use the oracle; the product never does)


def golden_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def stub_vector(text: str, d: int) -> list:
    """Deterministic near-unit vector for a text (identical to oracle/make_golden.py)."""
    rng = np.random.default_rng(zlib.crc32(text.encode("utf-8")))
    v = rng.standard_normal(d)
    v /= np.sqrt((v * v).sum())
    return [float(x) for x in v]


def reference_import_path():
    """sys.path entry for the byte-compiled reference (oracle/_ref), or None."""
    p = os.path.join(ROOT, "oracle", "_ref", "svs_ref.bin")       # zip of the reference's .pyc files
    return p if os.path.isfile(p) else None
