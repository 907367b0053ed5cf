"""TEST INFRASTRUCTURE ONLY -- recipe that BUILDS the reference for machines without ``/root/reference``.

The reference is interpreted Python 2 with no build system; "building" it means byte-compiling its hot-path
modules where they lie, after the same two textual Py2->3 substitutions ``oracle/ref_shim.py`` applies when it
executes them from source (``xrange(`` -> ``range(``, ``.iteritems()`` -> ``.items()``).  The outputs are ordinary
CPython bytecode files (the ``.pyc`` format, no source text; named ``*.bin`` because snapshot tools skip ``*.pyc``) under ``oracle/_ref/``, which is git-ignored but travels with a
``gpurun`` snapshot -- the GPU box runs the same interpreter, so ``bench.py --impl reference`` and the
``cpu_baseline`` leg can time the UNMODIFIED reference code there, and the reference's own unit tests can be run
against the drop-in classes.  No reference source is copied into this repository.

    python -m oracle.build_ref            # or __graft_entry__.build()
"""
from __future__ import annotations

import importlib.util
import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = os.environ.get("PARAKEET_REFERENCE_SRC", "/root/reference/src")
MODULES = ("matrix", "utils", "prkt_core_v2", "prkt_ros", "test_prkt_ros2")
PY2_SUBSTITUTIONS = (("xrange(", "range("), (".iteritems()", ".items()"))


def pyc_path(name: str) -> str:
    return os.path.join(OUT_DIR, "%s.cpython-%d%d.bin" % (name, sys.version_info[0], sys.version_info[1]))


def source_available(src_dir: str = REFERENCE_SRC) -> bool:
    return os.path.isfile(os.path.join(src_dir, "prkt_core_v2.py"))


def built() -> bool:
    return all(os.path.isfile(pyc_path(m)) for m in MODULES)


def build(src_dir: str = REFERENCE_SRC, force: bool = False) -> list:
    """Compile every module of ``MODULES`` found under ``src_dir``; returns the files written."""
    if not source_available(src_dir):
        return []
    os.makedirs(OUT_DIR, exist_ok=True)
    written = []
    for name in MODULES:
        src = os.path.join(src_dir, name + ".py")
        dst = pyc_path(name)
        if not os.path.isfile(src):
            continue
        if not force and os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        with open(src, "r") as fh:
            text = fh.read()
        for old, new in PY2_SUBSTITUTIONS:
            text = text.replace(old, new)
        code = compile(text, src, "exec", dont_inherit=True)
        tmp = dst + ".tmp.%d" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(importlib.util.MAGIC_NUMBER)
            fh.write(b"\x00" * 12)          # flags, mtime, size: unchecked (never imported by the import system)
            marshal.dump(code, fh)
        os.replace(tmp, dst)
        written.append(dst)
    return written


def load_code(name: str):
    """Code object of a compiled reference module, or None."""
    path = pyc_path(name)
    if not os.path.isfile(path):
        return None
    with open(path, "rb") as fh:
        if fh.read(4) != importlib.util.MAGIC_NUMBER:
            return None
        fh.read(12)
        return marshal.load(fh)


if __name__ == "__main__":
    out = build(force=True)
    print("compiled %d reference modules into %s" % (len(out), OUT_DIR) if out else
          "reference sources not found under %s; nothing built" % REFERENCE_SRC)
