#!/usr/bin/env python
"""Build-time variants of the library side by side, and their kernel times on one B200.

    python tools/k2_variants.py build            # here (no GPU): variants/<name>.so for every entry of VARIANTS
    python tools/k2_variants.py time [names...]  # on the GPU box: one subprocess per variant, one JSON line each

`build` compiles the working tree's csrc/ with the variant's -D flags (and `base` from `git show HEAD:` when the tree
is dirty, so a change can be timed against the last commit in the same call).  `time` runs BASELINE config 2 (2^20 x 64 x 8,
fp32 records, fp32 algebra) and the colour-ambiguous world through bench.Runner with the variant's library
(PARAKEET_B200_LIB) and prints motion / K2 / resampling medians.  variants/ is git-ignored but travels with gpurun.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "variants")

VARIANTS = {
    "new": [],
    "l2screen": ["-DPK_SCREEN_SAD=0"],
    "noring": ["-DPK_CLOOP_PREFETCH=0"],
    "reg112": ["-DPK_K2_REGCAP_F32=112"],
    "reg120w1": ["-DPK_K2_REGCAP_F32=120", "-DPK_MEASURE_WARPS=1"],
    "reg112w3": ["-DPK_K2_REGCAP_F32=112", "-DPK_MEASURE_WARPS=3"],
    "nometa": ["-DPK_COPY_META_CACHE=0"],
    "l2noring": ["-DPK_SCREEN_SAD=0", "-DPK_CLOOP_PREFETCH=0"],
    "sad": ["-DPK_SCREEN_SAD=1"],
}


def _compile(src_dir, inc_dir, out, flags):
    from parakeet_slam_b200 import _lib
    srcs = [os.path.join(src_dir, s) for s in _lib.SOURCES]
    cmd = ["nvcc"] + _lib.NVCC_FLAGS + list(flags) + ["-I", inc_dir, "-I", src_dir, "-o", out] + srcs
    subprocess.run(cmd, check=True)
    print("built", out, " ".join(flags))


def build(names):
    os.makedirs(VDIR, exist_ok=True)
    src, inc = os.path.join(ROOT, "parakeet_slam_b200", "csrc"), os.path.join(ROOT, "include")
    for name in names or list(VARIANTS):
        if name == "base":
            continue
        _compile(src, inc, os.path.join(VDIR, name + ".so"), VARIANTS[name])
    if not names or "base" in names:
        with tempfile.TemporaryDirectory() as tmp:
            for rel in ("parakeet_slam_b200/csrc", "include"):
                os.makedirs(os.path.join(tmp, rel))
                files = subprocess.run(["git", "ls-tree", "--name-only", "HEAD", rel + "/"], cwd=ROOT, check=True,
                                       capture_output=True, text=True).stdout.split()
                for f in files:
                    with open(os.path.join(tmp, f), "wb") as fh:
                        fh.write(subprocess.run(["git", "show", "HEAD:" + f], cwd=ROOT, check=True,
                                                capture_output=True).stdout)
            _compile(os.path.join(tmp, "parakeet_slam_b200/csrc"), os.path.join(tmp, "include"),
                     os.path.join(VDIR, "base.so"), [])


def _worker(name, what):
    import argparse
    import bench
    args = argparse.Namespace(exchange="peer")
    out = {"variant": name}
    steps, warm = 50, 5
    for tag, kw, total in (("c2", {}, 700.0), ("ambiguous6", {"num_colors": 6}, 300.0)):
        if what and tag not in what:
            continue
        R = bench.Runner(args, 1, 0, 1 << 20, 64, "f32", "f32", "peer", **kw)
        m = R.measure(R.step, steps, warm, min_total_ms=total)
        _, mf, ev, fd = R.stats()
        out[tag] = {"ms_per_step": m["median_ms"] / steps, **{k: round(v, 5) for k, v in m["kernel_ms"].items()},
                    "matched": round(mf, 5), "evals_pp": round(ev, 3), "f_dup": round(fd, 4), "blocks": len(m["block_ms"])}
        R.close()
    if not what or "sweep" in what:
        # the resampling chain at stated duplicate fractions (weights set by hand), block copies included
        import statistics
        import torch
        R = bench.Runner(args, 1, 0, 1 << 20, 64, "f32", "f32", "peer")
        for _ in range(3):
            R.step()
        fs, idx = R.fs, torch.arange(1 << 20, device="cuda")
        ev = lambda: torch.cuda.Event(enable_timing=True)
        sw = {}
        for label, dead in (("4%", idx % 25 == 0), ("10%", idx % 10 == 0), ("50%", idx % 2 == 1), ("94%", idx % 16 != 0)):
            ts = []
            for rep in range(6):
                fs.pose[:, 3] = torch.where(dead, 0.0, 1.0).to(torch.float64)
                a, b = ev(), ev()
                a.record()
                fs.low_variance_resample()
                fs.wait_blocks()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            sw[label] = round(statistics.median(ts[1:]), 4)
        out["resample_ms_at_f_dup"] = sw
        R.close()
    print("VARIANT " + json.dumps(out), flush=True)


def time_all(names):
    names = names or sorted(f[:-3] for f in os.listdir(VDIR) if f.endswith(".so"))
    for name in names:
        env = dict(os.environ, PARAKEET_B200_LIB=os.path.join(VDIR, name + ".so"))
        subprocess.run([sys.executable, os.path.abspath(__file__), "_worker", name], env=env, cwd=ROOT)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    if cmd == "build":
        build(sys.argv[2:])
    elif cmd == "time":
        time_all(sys.argv[2:])
    elif cmd == "_worker":
        _worker(sys.argv[2], sys.argv[3:])
    else:
        raise SystemExit(__doc__)
