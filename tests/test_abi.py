"""The C-ABI shared library builds, loads and exports every symbol include/svsb200.h declares.
No compute is launched here (there is no GPU on the CPU box)."""
import ctypes
import os
import re
import subprocess

import pytest

from _util import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "svsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svsb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    syms = _declared_symbols()
    for must in ("svsb_create", "svsb_destroy", "svsb_load_begin", "svsb_load_rows", "svsb_load_end",
                 "svsb_invalidate", "svsb_is_loaded", "svsb_query", "svsb_query_batch", "svsb_last_error"):
        assert must in syms


def test_library_builds_loads_and_exports_every_declared_symbol():
    from svs_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/svsb200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    bound = _lib.load()
    assert b"sm_100a" in bound.svsb_version()
    assert bound.svsb_launch_count() >= 0


def test_library_contains_sm100a_code_and_tma_bulk_copies():
    from svs_b200 import build
    path = build.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "gemv_tma_kernel" in sass
    assert "UBLKCP" in sass, "TMA bulk copy (cp.async.bulk) missing from the GEMV kernel's SASS"
    assert "SYNCS.ARRIVE.TRANS64" in sass, "mbarrier expect_tx missing"


def test_no_device_fails_loudly_instead_of_falling_back():
    """On a box without a B200 the engine refuses to exist; nothing computes on the CPU."""
    import svs_b200
    from svs_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.svsb_create(None, 0, ctypes.byref(h))
    if rc == 0:                                   # running on the GPU box: nothing to check here
        lib.svsb_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == _lib.SVSB_E_NO_DEVICE
    assert "no CPU path" in _lib.last_error()
    with pytest.raises(svs_b200.EngineError):
        svs_b200.Engine()
    with pytest.raises(svs_b200.EngineError):
        svs_b200.DeviceEmbeddingsMatrix()._get_engine()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "svs_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "svs_oracle" not in src.replace("oracle/svs_oracle.py", ""), f
                assert "import oracle" not in src and "from oracle" not in src, f
