"""In-process check that the particle-sharded filter reproduces the single-GPU filter bit for bit.

``bench.py`` runs it on every rank before it times anything at N > 1 (so a scaling number is never the
throughput of an unverified data plane) and ``tests/test_gpu_sharded.py`` / ``tests/multi_gpu_check.py``
cover more cases.  Both sides of the comparison are product code: ``ShardedFastSLAM`` over all ranks against
``FastSLAM`` holding all particles on this rank's GPU (cheap at the check's size).  Nothing from ``oracle/``
is involved -- parity with the REFERENCE is what the single-GPU parity tests establish; this establishes that
sharding does not change a bit of it.
"""
from __future__ import annotations

import random

import numpy as np


def _features(scn):
    from .core import Feature
    return [Feature(mean=np.array(row), covar=np.identity(5) * scn.preset_covar) for row in scn.landmarks]


def sharded_equals_single(particles_per_rank=8192, frames=8, exchange="peer", cases=None, group=None):
    """Run each case on the sharded and on the single-GPU filter and compare, per frame, poses, weights and ancestors
    of this rank's slice, and after the last frame the landmark maps (and the orphan readings in spawn mode).

    Returns ``dict(identical, migrations, cases)`` -- identical over ALL ranks (an all-reduce), migrations = particles
    that crossed a shard boundary, summed over cases, frames and ranks.
    """
    import torch
    import torch.distributed as dist
    from .core import FastSLAM
    from .rosless import Time, messages
    from .scenario import DT_NSEC, make_scenario
    from .sharded import ShardedFastSLAM

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Ml = int(particles_per_rank)
    M = Ml * world
    if cases is None:
        # (storage, arithmetic, skewed weights, spawn mode)
        cases = (("f32", "f32", True, False), ("f64", "f64", True, True))
    scn = make_scenario("c2", num_particles=M, num_landmarks=32, frames=frames, sigma_color=2.0, sigma_bearing=0.05)

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)

    def run(cls, dtype, arith, skew, spawn, **kw):
        clk, urng = Clk(), random.Random(99)
        if spawn:
            kw = dict(kw, spawn=True, capacity=32, orphan_capacity=16)
        fs = cls([] if spawn else _features(scn), num_particles=M, dtype=dtype, arithmetic=arith, noise="philox",
                 seed=7, uniform=urng.random, clock=clk, **kw)
        fs.keep_trace = True
        tw = messages.Twist()
        tw.linear.x, tw.angular.z = scn.v, scn.w
        fs.last_control = tw
        out, moved = [], 0
        for t in range(frames):
            clk.ns += DT_NSEC
            fs.motion_update(tw)
            fs.measurement_update(scn.observations[t])
            if skew:
                # starve the index ranges of alternating ranks: about half of the other ranks' offspring must move
                gidx = fs.particle_offset + torch.arange(fs.num_particles, device=fs.pose.device)
                odd = ((gidx // Ml) + t) % 2 == 1
                fs.pose[:, 3] *= torch.where(odd, 1e-3, 1.0).to(torch.float64)
            w = fs.pose[:, 3].clone()
            fs.low_variance_resample()
            if isinstance(fs, ShardedFastSLAM):
                plan = fs.last_plan
                moved += plan["n_lo"] + plan["n_hi"]
            out.append((fs.pose[:, :3].clone(), w, fs.last_ancestors.clone()))
        maps = list(fs.export_maps())
        # slots at or beyond n_live hold whatever the block's previous owner left there (copy-on-resample moves live
        # landmarks only): not part of the filter state
        dead = np.arange(maps[0].shape[1])[None, :] >= maps[5][:, None]
        for a in maps[:5]:
            a[dead] = 0
        if spawn:
            rows, totals = fs.export_orphans()
            maps += [totals, np.array([len(r) for r in rows])]
        if isinstance(fs, ShardedFastSLAM):
            fs.close()
        return out, maps, moved

    ok, migrations, report = True, 0, []
    lo, hi = rank * Ml, (rank + 1) * Ml
    for dtype, arith, skew, spawn in cases:
        out_s, maps_s, moved = run(ShardedFastSLAM, dtype, arith, skew, spawn, exchange=exchange, group=group)
        out_1, maps_1, _ = run(FastSLAM, dtype, arith, skew, spawn)
        same = True
        for (p_s, w_s, a_s), (p_1, w_1, a_1) in zip(out_s, out_1):
            same = same and torch.equal(p_s, p_1[lo:hi]) and torch.equal(w_s, w_1[lo:hi]) and torch.equal(a_s, a_1[lo:hi])
        for a, b in zip(maps_s, maps_1):
            same = same and np.array_equal(a, b[lo:hi])
        t = torch.tensor([0 if same else 1, moved], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, group=group)
        case_ok = int(t[0].item()) == 0 and (world == 1 or int(t[1].item()) > 0)
        ok = ok and case_ok
        migrations += int(t[1].item())
        report.append(dict(storage=dtype, arithmetic=arith, skewed=skew, spawn=spawn, identical=case_ok,
                           migrations=int(t[1].item())))
    return dict(identical=ok, migrations=migrations, particles=M, frames=frames, exchange=exchange, cases=report)
