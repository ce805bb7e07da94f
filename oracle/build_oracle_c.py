"""Compile oracle/svs_oracle_c.c (the plain-C restatement of the hot path) into oracle/_ref/libsvs_oracle_c.so and bind it
with ctypes.  TEST INFRASTRUCTURE ONLY: used by tests/test_oracle_c.py; never by svs_b200/."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "svs_oracle_c.c")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libsvs_oracle_c.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("gcc not found")
    subprocess.run([gcc, "-O2", "-std=c99", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


class OracleC:
    """ctypes face of the C restatement; arrays in, arrays / lists out."""

    def __init__(self):
        self.lib = C.CDLL(build())
        self.lib.svs_scores.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        self.lib.svs_get_top_k.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        self.lib.svs_get_top_k.restype = C.c_int64
        self.lib.svs_superheavy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.lib.svs_superheavy.restype = C.c_int64
        self.lib.svs_blob_to_row.argtypes = [C.c_char_p, C.c_int64, C.c_void_p]
        self.lib.svs_blob_to_row.restype = C.c_int64

    def scores(self, m: np.ndarray, q: np.ndarray) -> np.ndarray:
        m = np.ascontiguousarray(m, dtype=np.float32); q = np.ascontiguousarray(q, dtype=np.float32)
        assert m.ndim == 2 and q.shape == (m.shape[1],)
        x = np.empty(m.shape[0], dtype=np.float32)
        self.lib.svs_scores(m.ctypes.data, m.shape[0], m.shape[1], q.ctypes.data, x.ctypes.data)
        return x

    def get_top_k(self, scores: np.ndarray, k: int):
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        cap = max(0, min(int(k), len(scores)))
        s = np.empty(cap, dtype=np.float32); i = np.empty(cap, dtype=np.int64)
        c = self.lib.svs_get_top_k(scores.ctypes.data, len(scores), int(k), s.ctypes.data, i.ctypes.data)
        assert c == cap
        return [(float(a), int(b)) for a, b in zip(s, i)]

    def superheavy(self, m: np.ndarray, ids: np.ndarray, q: np.ndarray, k: int):
        m = np.ascontiguousarray(m, dtype=np.float32); q = np.ascontiguousarray(q, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        cap = max(0, min(int(k), len(ids)))
        s = np.empty(cap, dtype=np.float32); i = np.empty(cap, dtype=np.int64)
        c = self.lib.svs_superheavy(m.ctypes.data, ids.ctypes.data, len(ids), m.shape[1], q.ctypes.data, int(k), s.ctypes.data, i.ctypes.data)
        assert c == cap
        return [(float(a), int(b)) for a, b in zip(s, i)]

    def blob_to_row(self, blob: bytes):
        out = np.empty(len(blob) // 4, dtype=np.float32)
        c = self.lib.svs_blob_to_row(blob, len(blob), out.ctypes.data)
        if c < 0:
            raise AssertionError("blob length is not a multiple of 4")
        return out


if __name__ == "__main__":
    print(build(force=True))
