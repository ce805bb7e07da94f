"""Build the real reference into ``oracle/_ref/`` (git-ignored).  TEST INFRASTRUCTURE ONLY.

The reference (Rhobota/svs) is pure Python, so "building" it means byte-compiling its modules
from where they lie under ``/root/reference/src/svs`` into sourceless ``.pyc`` files under
``oracle/_ref/svs/``.  No reference source is copied into the repository; only compiled
artefacts are written, and only into ``oracle/_ref/`` (which travels to the GPU box with the
snapshot the way the engine's own built ``.so`` does).

(The other way to install it, ``pip install --target ... /root/reference``, needs the
``hatchling`` build backend, which is neither installed nor in /opt/wheelhouse.)

What it is used for:
  * ``bench.py``'s ``cpu_baseline`` / ``--impl reference``: timing the reference's own
    ``svs.util.get_top_k`` + ``np.dot`` (kind "reference" instead of the oracle "port");
  * ``tests/test_dropin_gpu.py``: driving the reference's unmodified ``svs.KB`` / ``svs.AsyncKB``
    with ``svs_b200.dropin.install()`` applied, on a GPU.
When ``oracle/_ref`` is absent those fall back to ``oracle/svs_oracle.py`` / skip.

Usage:  python oracle/build_ref.py [/root/reference]
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
ZIP = os.path.join(OUT, "svs_ref.bin")     # a zip archive; neutral extension so that snapshots keep it


def build(reference_root: str = "/root/reference") -> bool:
    src_pkg = os.path.join(reference_root, "src", "svs")
    if not os.path.isdir(src_pkg):
        return False
    dst_pkg = os.path.join(OUT, "svs")
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    for dirpath, _dirnames, filenames in os.walk(src_pkg):
        rel = os.path.relpath(dirpath, src_pkg)
        dst_dir = os.path.normpath(os.path.join(dst_pkg, rel))
        os.makedirs(dst_dir, exist_ok=True)
        for fn in filenames:
            if fn.endswith(".py"):
                # sourceless import layout: pkg/module.pyc next to where module.py would be
                py_compile.compile(os.path.join(dirpath, fn),
                                   cfile=os.path.join(dst_dir, fn + "c"),
                                   dfile=f"<reference>/src/svs/{rel}/{fn}",
                                   doraise=True)
    # The GPU-box snapshot drops *.pyc files, so the importable artefact is a zip of them
    # (zipimport loads sourceless .pyc); the loose tree is removed again.
    import zipfile
    with zipfile.ZipFile(ZIP, "w", zipfile.ZIP_DEFLATED) as z:
        for dirpath, _dirnames, filenames in os.walk(dst_pkg):
            for fn in filenames:
                full = os.path.join(dirpath, fn)
                z.write(full, os.path.relpath(full, OUT))
    shutil.rmtree(dst_pkg)
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as f:
        f.write(f"byte-compiled from {src_pkg} with python {sys.version.split()[0]}\n")
    return True


def import_path() -> str | None:
    """Entry to put on sys.path to import the built reference (a zip of .pyc files), or None if not built."""
    return ZIP if os.path.isfile(ZIP) else None


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    ok = build(root)
    print("built oracle/_ref" if ok else f"reference not found at {root}; nothing built")
