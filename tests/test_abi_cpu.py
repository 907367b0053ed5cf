"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/parakeet_b200.h declares.  No compute calls: runs without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "parakeet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_typed():
    from parakeet_slam_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libparakeet_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert sorted(_lib.SIGNATURES) == names


def test_layout_queries_and_error_convention():
    from parakeet_slam_b200 import _lib
    lib = _lib.load()
    assert lib.pk_version() == 1
    assert lib.pk_hot_bytes(_lib.PK_DTYPE_F32) == 4
    assert lib.pk_cold_bytes(_lib.PK_DTYPE_F32) == 64 and lib.pk_cold_bytes(_lib.PK_DTYPE_F64) == 160
    assert lib.pk_block_bytes(64, _lib.PK_DTYPE_F32) == 64 * 68
    assert lib.pk_block_bytes(20, _lib.PK_DTYPE_F64) == 128 + 20 * 160   # key region padded to 64 B
    assert lib.pk_block_bytes(3, _lib.PK_DTYPE_F32) == 64 + 3 * 64
    assert lib.pk_block_bytes(64, _lib.dtype_with_orphans(_lib.PK_DTYPE_F32, 32)) == 64 * 68 + 64 + 32 * 64
    assert lib.pk_num_scan_blocks(1) == 1 and lib.pk_num_scan_blocks(1025) == 2
    p = _lib.default_params()
    assert (p.bearing_gate, p.color_gate, p.no_match_weight, p.qt_diag, p.promote_count) == (0.5, 300.0, 0.1, 0.1, 5)
    # argument errors are reported by return code + message, never by crashing
    rc = lib.pk_motion_update(None, 10, None, 0, 0, 0, 0.0, 0.0, 0.0, None)
    assert rc == -1 and b"pose4" in lib.pk_last_error()
    rc = lib.pk_measurement_update(None, None, None, None, 4, 0, 8, None, 200, ctypes.byref(p), None, None, None)
    assert rc == -1


def test_product_has_no_cpu_fallback_and_does_not_import_oracle():
    import torch
    from parakeet_slam_b200 import core, _lib
    if not torch.cuda.is_available():
        with pytest.raises(_lib.ParakeetLibraryError):
            core.FastSLAM()
    pkg = os.path.join(ROOT, "parakeet_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_python_constants_match_the_header():
    """Every PK_* constant the Python binding mirrors has the value include/parakeet_b200.h gives it."""
    from parakeet_slam_b200 import _lib
    text = open(os.path.join(ROOT, "include", "parakeet_b200.h")).read()
    defines = {}
    for name, val in re.findall(r"^#define\s+(PK_[A-Z0-9_]+)\s+([^\s/]+)", text, flags=re.M):
        v = val.rstrip("uUlL")
        try:
            defines[name] = int(v, 0)
        except ValueError:
            continue
    checked = 0
    for name, value in vars(_lib).items():
        if name.startswith("PK_") and isinstance(value, int) and name in defines:
            assert defines[name] == value, (name, defines[name], value)
            checked += 1
    assert checked >= 25
    assert _lib.PK_XP_RANK_LOC == 16 + _lib.PK_MAX_RANKS and _lib.PK_XPLAN_LONGS >= 16 + 2 * _lib.PK_MAX_RANKS


def test_layout_code_helpers():
    from parakeet_slam_b200 import _lib
    lib = _lib.load()
    base = lib.pk_block_bytes(256, _lib.PK_DTYPE_F32)
    assert base == 1024 + 256 * 64
    assert lib.pk_block_bytes(256, _lib.PK_DTYPE_F32 | _lib.PK_DTYPE_ARITH_F32) == base       # arithmetic flag: no layout change
    assert lib.pk_block_bytes(256, _lib.dtype_with_orphans(_lib.PK_DTYPE_F32, 16)) == base + 64 + 16 * 64
    assert lib.pk_particle_record_bytes(256, _lib.PK_DTYPE_F32) == 64 + base
    for cap in (1, 3, 20, 64, 100, 1024):            # records start on 64-byte (f32) / 32-byte (f64) boundaries
        assert lib.pk_block_bytes(cap, _lib.PK_DTYPE_F32) % 64 == 0
        assert lib.pk_block_bytes(cap, _lib.PK_DTYPE_F64) % 32 == 0
    assert lib.pk_obs_table_bytes() >= 6 * 64 * 8 + 64 * 4


def test_colour_screen_bound_contains_the_gate():
    """The screen of the fused kernel tests the squared key distance against B = floor(gate + 2 sqrt(3 gate) + 3) + 1
    (its -DPK_SCREEN_SAD=1 form: `sum |key difference| <= floor(sqrt(3 B))`; pk_measure.cu, measurement_common): both
    restated here and checked against the exact gate (prkt_core_v2.py:441,
    `abs(dr^2 + dg^2 + db^2) > gate` rejects) on random colour pairs, pairs on the gate's sphere in the all-equal direction
    (the worst case of an L1 bound) and pairs at the clamps of the 8-bit keys."""
    import math
    import numpy as np

    def key(c):
        return np.rint(np.clip(c, 0.0, 255.0)).astype(np.int64)

    rs = np.random.RandomState(3)
    for gate in (300.0, 0.0, 1.0, 75.0, 1200.0, 50000.0):
        B = math.floor(gate + 2.0 * math.sqrt(3.0 * gate) + 3.0) + 1
        t = math.isqrt(3 * B)
        a = rs.uniform(-5.0, 260.0, (200000, 3))
        r = math.sqrt(gate / 3.0)
        off = np.concatenate([rs.normal(0.0, 1.0, (100000, 3)) * (r + 1.0),
                              rs.choice([-1.0, 1.0], (100000, 3)) * rs.uniform(0.9 * r, 1.0 * r, (100000, 1))])
        b = a + off
        inside = np.abs(((a - b) ** 2).sum(1)) <= gate          # the reference does not reject
        l1 = np.abs(key(a) - key(b)).sum(1)
        l2 = ((key(a) - key(b)) ** 2).sum(1)
        assert inside.sum() > 1000
        assert np.all(l2[inside] <= B), gate
        assert np.all(l1[inside] <= t), gate
