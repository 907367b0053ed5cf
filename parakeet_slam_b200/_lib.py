"""ctypes binding of ``libparakeet_b200.so`` (the C ABI declared in ``include/parakeet_b200.h``).

There is no CPU fallback: if the shared library cannot be built or loaded, or the
device is not a B200 (sm_100), every entry point raises.  The library is built
in-tree (``parakeet_slam_b200/libparakeet_b200.so``) with

    nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3

so that it travels with a repository snapshot.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.environ.get("PARAKEET_B200_LIB") or os.path.join(PKG_DIR, "libparakeet_b200.so")
_LIB_OVERRIDDEN = bool(os.environ.get("PARAKEET_B200_LIB"))
SOURCES = ("pk_abi.cu", "pk_motion.cu", "pk_measure.cu", "pk_spawn.cu", "pk_resample.cu", "pk_probe.cu", "pk_sim.cu")
HEADERS = (os.path.join(CSRC_DIR, "pk_common.cuh"), os.path.join(CSRC_DIR, "pk_filter_math.cuh"),
           os.path.join(INCLUDE_DIR, "parakeet_b200.h"))

NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]

PK_DTYPE_F32 = 0
PK_DTYPE_F64 = 1
PK_MAX_OBS = 64
PK_SCAN_BLOCK = 1024
PK_NUM_STATS = 8
PK_PLAN_DOUBLES = 8
PK_META_COUNT_MASK = 0x00FFFFFF
PK_META_IMMUTABLE = 0x10000000
PK_META_POTENTIAL = 0x20000000
PK_STAT_MATCHED, PK_STAT_UNMATCHED, PK_STAT_EVALUATED, PK_STAT_FLAGS = 0, 1, 2, 3
PK_STAT_SAME_LANDMARK, PK_STAT_PROMOTED, PK_STAT_SPAWNED, PK_STAT_ORPHANED = 4, 5, 6, 7
PK_FLAG_SINGULAR_COV, PK_FLAG_NONFINITE_WEIGHT, PK_FLAG_REPROMOTED = 1, 2, 4
PK_FLAG_MAP_FULL, PK_FLAG_ORPHAN_EXPIRED, PK_FLAG_SPAWN_DEGENERATE = 8, 16, 32
PK_MAX_ORPHANS = 1024
PK_DTYPE_ARITH_F32 = 0x1000000
PK_MODEL_TEXTBOOK = 1
PK_MODEL_LOG_WEIGHTS = 2


def dtype_with_orphans(base: int, slots: int) -> int:
    """``PK_DTYPE_WITH_ORPHANS``: layout code = storage type | orphan slots << 8."""
    return base | (int(slots) << 8)
PK_MAX_RANKS = 32
PK_XPLAN_LONGS = 80
PK_PEER_HANDLE_BYTES = 64
PK_PEER_OVERFLOW, PK_PEER_TIMEOUT = 1, 2
PK_PEER_STATUS_WORDS = 4
(PK_XP_EMIT_LO, PK_XP_EMIT_N, PK_XP_N_LO, PK_XP_N_LOC, PK_XP_N_HI, PK_XP_N_BELOW, PK_XP_N_ABOVE,
 PK_XP_ABOVE_START, PK_XP_N_SEND, PK_XP_N_IN, PK_XP_OVERFLOW) = range(11)
PK_XP_RANK_LO = 16
PK_XP_RANK_LOC = 16 + PK_MAX_RANKS


class PkParams(ctypes.Structure):
    """``pk_params`` -- the reference's literals."""
    _fields_ = [("bearing_gate", ctypes.c_double), ("position_gate", ctypes.c_double),
                ("color_gate", ctypes.c_double), ("no_match_weight", ctypes.c_double),
                ("qt_diag", ctypes.c_double), ("promote_count", ctypes.c_int),
                ("model", ctypes.c_int)]


class ParakeetLibraryError(RuntimeError):
    pass


def _sources():
    return [os.path.join(CSRC_DIR, s) for s in SOURCES if os.path.exists(os.path.join(CSRC_DIR, s))]


def needs_build() -> bool:
    if _LIB_OVERRIDDEN:
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in list(_sources()) + list(HEADERS))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` into one shared library for sm_100a (cross-compiles on a
    machine without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise ParakeetLibraryError("nvcc not found; cannot build %s" % LIB_PATH)
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE_DIR, "-o", tmp] + _sources()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise ParakeetLibraryError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(proc.stdout)
    return LIB_PATH


_P = ctypes.c_void_p
_LL = ctypes.c_longlong
_ULL = ctypes.c_ulonglong
_I = ctypes.c_int
_D = ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/parakeet_b200.h declares
SIGNATURES = {
    "pk_version": (_I, []),
    "pk_last_error": (ctypes.c_char_p, []),
    "pk_default_params": (_I, [ctypes.POINTER(PkParams)]),
    "pk_check_device": (_I, []),
    "pk_hot_bytes": (_I, [_I]),
    "pk_cold_bytes": (_I, [_I]),
    "pk_block_bytes": (_LL, [_I, _I]),
    "pk_init_particles": (_I, [_P, _P, _P, _LL, _I, _I, _P]),
    "pk_map_broadcast": (_I, [_P, _I, _I, _LL, _LL, _I, _P, _P, _P, _P, _P, _P]),
    "pk_map_export": (_I, [_P, _I, _I, _P, _LL, _LL, _P, _P, _P, _P, _P, _P]),
    "pk_map_import": (_I, [_P, _I, _I, _P, _LL, _LL, _P, _P, _P, _P, _P, _P]),
    "pk_motion_update": (_I, [_P, _LL, _P, _ULL, _ULL, _LL, _D, _D, _D, _P]),
    "pk_measurement_update": (_I, [_P, _P, _P, _P, _I, _I, _LL, _P, _I, ctypes.POINTER(PkParams),
                                   _P, _P, _P]),
    "pk_obs_table_bytes": (_LL, []),
    "pk_measurement_update_dev": (_I, [_P, _P, _P, _P, _I, _I, _LL, _P, _I, ctypes.POINTER(PkParams),
                                       _P, _P, _P, _P]),
    "pk_spawn_update_dev": (_I, [_P, _P, _P, _P, _I, _I, _LL, _P, _I, _P, _D, _P, _P]),
    "pk_simulate_scan": (_I, [_P, _I, _D, _D, _D, _I, _P, _ULL, _ULL, _D, _D, _P, _P, _P, _P]),
    "pk_accuracy": (_I, [_P, _LL, _D, _D, _D, _P, _P, _P]),
    "pk_map_error": (_I, [_P, _I, _I, _P, _P, _LL, _P, _I, _P, _P, _P]),
    "pk_spawn_update": (_I, [_P, _P, _P, _P, _I, _I, _LL, _P, _I, _P, _D, _P, _P]),
    "pk_orphans_export": (_I, [_P, _I, _I, _P, _LL, _LL, _P, _P, _P]),
    "pk_num_scan_blocks": (_LL, [_LL]),
    "pk_weight_scan": (_I, [_P, _LL, _P, _P, _P]),
    "pk_resample_thresholds": (_I, [_P, _LL, _LL, _D, _P, _P, _P, _P]),
    "pk_resample_ancestors": (_I, [_P, _LL, _LL, _LL, _P, _P, _P, _LL, _LL, _LL, _P, _P, _P, _P, _P]),
    "pk_resample_plan": (_I, [_P, _LL, _LL, _LL, _P, _P, _P, _LL, _LL, _LL, _P, _P, _P, _P, _P]),
    "pk_gather_workspace_bytes": (_LL, [_LL]),
    "pk_resample_gather_planned": (_I, [_P, _LL, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "pk_resample_copy_blocks": (_I, [_P, _I, _I, _LL, _P, _P, _P]),
    "pk_resample_gather": (_I, [_P, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "pk_particle_record_bytes": (_LL, [_I, _I]),
    "pk_pack_particles": (_I, [_P, _LL, _LL, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "pk_resample_gather_sharded": (_I, [_P, _P, _P, _LL, _LL, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I,
                                        _P, _P, _P]),
    "pk_copy_blocks": (_I, [_P, _P, _I, _I, _P, _P, _P, _LL, _P, _P]),
    "pk_peer_alloc": (_I, [_LL, ctypes.POINTER(_P)]),
    "pk_peer_free": (_I, [_P]),
    "pk_peer_export": (_I, [_P, _P]),
    "pk_peer_open": (_I, [_P, ctypes.POINTER(_P)]),
    "pk_peer_close": (_I, [_P]),
    "pk_weight_scan_publish": (_I, [_P, _LL, _P, _P, _P, _I, _I, _P]),
    "pk_peer_barrier": (_I, [_P, _I, _I, _ULL, _D, _P, _P]),
    "pk_exchange_plan": (_I, [_P, _LL, _I, _I, _LL, _LL, _P, _P, _P]),
    "pk_exchange_plan_host": (_I, [_P, _I, _I, _LL, _LL, _P]),
    "pk_push_particles": (_I, [_P, _P, _LL, _I, _P, _P, _P, _P, _I, _I, _P, _LL, _P, _P]),
    "pk_resample_gather_peer": (_I, [_P, _P, _P, _P, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _LL, _P, _I, _I,
                                     _P, _P, _P, _I, _I, _ULL, _D, _P, _P, _P]),
    "pk_peer_post": (_I, [_P, _I, _I, _ULL, _P]),
    "pk_resample_thresholds_peer": (_I, [_P, _LL, _LL, _D, _P, _P, _P, _P, _I, _I, _ULL, _D, _LL, _LL, _P, _P, _P]),
    "pk_log_weights_max": (_I, [_P, _LL, _P, _P, _P]),
    "pk_log_weights_normalise": (_I, [_P, _LL, _P, _P, _P, _P]),
    "pk_summary_partial": (_I, [_P, _LL, _P, _P, _P]),
    "pk_best_particle": (_I, [_P, _LL, _P, _P, _P]),
    "pk_probe_likelihood": (_I, [_P, _P, _P, _P, _P, _P, _LL, ctypes.POINTER(PkParams), _P, _P]),
    "pk_probe_ekf": (_I, [_P, _P, _P, _P, _P, _P, _LL, ctypes.POINTER(PkParams), _P, _P, _P, _P, _P]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_needed: bool = True) -> ctypes.CDLL:
    """Load (building first if the sources are newer) and type every entry point."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_needed and needs_build():
            build()
        if not os.path.exists(LIB_PATH):
            raise ParakeetLibraryError("%s is missing and could not be built; there is no CPU fallback"
                                       % LIB_PATH)
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as exc:
            raise ParakeetLibraryError("cannot load %s: %s (no CPU fallback)" % (LIB_PATH, exc))
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                raise ParakeetLibraryError("%s does not export %s" % (LIB_PATH, name))
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().pk_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise ParakeetLibraryError("%s failed (%d): %s" % (what or "libparakeet_b200 call", rc, last_error()))


def default_params() -> PkParams:
    p = PkParams()
    check(load().pk_default_params(ctypes.byref(p)), "pk_default_params")
    return p


def require_device() -> None:
    """Fail loudly unless a B200-class (sm_100) CUDA device is current."""
    import torch
    if not torch.cuda.is_available():
        raise ParakeetLibraryError("no CUDA device: parakeet_slam_b200 has no CPU fallback")
    check(load().pk_check_device(), "pk_check_device")


def ptr(t) -> int:
    """Raw device (or host) address of a tensor / ndarray, or 0 for None."""
    if t is None:
        return 0
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
