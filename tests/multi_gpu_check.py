"""Run under torchrun (one rank per GPU): the sharded filter must reproduce the single-GPU filter
bit for bit (ancestors, poses, landmark state), whatever the number of ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tests/multi_gpu_check.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _live_only(maps):
    """Landmark slots at or beyond n_live hold whatever the block's previous owner left there (copy-on-resample moves
    live landmarks only): not part of the filter state, zeroed before comparing."""
    mean5, covp, covc, meta, ids, nlive = maps
    dead = np.arange(mean5.shape[1])[None, :] >= nlive[:, None]
    for a in (mean5, covp, covc, meta, ids):
        a[dead] = 0
    return mean5, covp, covc, meta, ids, nlive


def main():
    import torch
    import torch.distributed as dist
    from device_harness import make_features
    from parakeet_slam_b200.core import FastSLAM
    from parakeet_slam_b200.rosless import Time, messages
    from parakeet_slam_b200.scenario import DT_NSEC, make_scenario
    from parakeet_slam_b200.sharded import ShardedFastSLAM

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    M = 8192 * world if len(sys.argv) < 2 else int(sys.argv[1])
    frames = 12
    scn = make_scenario("c2", num_particles=M, num_landmarks=32, frames=frames, sigma_color=2.0, sigma_bearing=0.05)
    Ml = M // world

    class Clk(object):
        ns = 0

        def __call__(self):
            return Time(0, self.ns)

    def run(cls, dtype, skew=False, spawn=False, **kw):
        clk = Clk()
        urng = random.Random(99)
        if spawn:   # unknown map: the orphan region of every block must migrate with it
            kw = dict(kw, spawn=True, capacity=32, orphan_capacity=16)
        fs = cls([] if spawn else make_features(scn), num_particles=M, dtype=dtype, noise="philox", seed=7,
                 uniform=urng.random, clock=clk, **kw)
        fs.keep_trace = True
        tw = messages.Twist()
        tw.linear.x, tw.angular.z = scn.v, scn.w
        fs.last_control = tw
        out = []
        moved = 0
        for t in range(frames):
            clk.ns += DT_NSEC
            fs.motion_update(tw)
            fs.measurement_update(scn.observations[t])
            if skew:
                # starve the particles of every odd rank's index range: about half of each even rank's
                # offspring must move to a neighbour (large exchanges, both directions)
                gidx = fs.particle_offset + torch.arange(fs.num_particles, device="cuda")
                odd = ((gidx // Ml) + t) % 2 == 1
                fs.pose[:, 3] *= torch.where(odd, 1e-3, 1.0).to(torch.float64)
            w = fs.pose[:, 3].clone()
            fs.low_variance_resample()
            if isinstance(fs, ShardedFastSLAM):
                moved += fs.last_plan["n_lo"] + fs.last_plan["n_hi"]
            out.append((fs.pose[:, :3].clone(), w, fs.last_ancestors.clone(), fs.summary()))
        maps = _live_only(fs.export_maps())
        if spawn:
            rows, totals = fs.export_orphans()
            maps = tuple(maps) + (totals, np.concatenate([r.reshape(-1) for r in rows] + [np.zeros(0)]),
                                  np.array([len(r) for r in rows]))
        return fs, out, maps, moved

    ok = True
    single = {}
    for dtype, exchange, skew, spawn in (("f64", "peer", False, False), ("f32", "peer", False, False),
                                         ("f32", "peer", True, False), ("f64", "peer", True, True),
                                         ("f64", "nccl", False, False), ("f32", "nccl", True, True)):
        fs_s, out_s, maps_s, moved = run(ShardedFastSLAM, dtype, skew, spawn, exchange=exchange)
        moved_t = torch.tensor([moved], device="cuda")
        dist.all_reduce(moved_t)
        # single-GPU reference on every rank (cheap at this size), compared on the rank's own slice
        if (dtype, skew, spawn) not in single:
            single[(dtype, skew, spawn)] = run(FastSLAM, dtype, skew, spawn)
        fs_1, out_1, maps_1, _ = single[(dtype, skew, spawn)]
        lo, hi = rank * Ml, (rank + 1) * Ml
        for t in range(frames):
            p_s, w_s, a_s, sum_s = out_s[t]
            p_1, w_1, a_1, sum_1 = out_1[t]
            same = (torch.equal(p_s, p_1[lo:hi]) and torch.equal(w_s, w_1[lo:hi]) and torch.equal(a_s, a_1[lo:hi]))
            if not same:
                print("rank %d dtype %s %s skew=%s frame %d: sharded != single (pose %s weight %s anc %s)" % (
                    rank, dtype, exchange, skew, t, torch.equal(p_s, p_1[lo:hi]), torch.equal(w_s, w_1[lo:hi]),
                    torch.equal(a_s, a_1[lo:hi])), flush=True)
                ok = False
                break
            if max(abs(x - y) for x, y in zip(sum_s, sum_1)) > 1e-12:
                print("rank %d summary differs" % rank, sum_s, sum_1, flush=True)
                ok = False
        if spawn:
            # orphan readings: totals per particle, and the concatenated readings of the rank's slice
            counts_1 = maps_1[8]
            off = np.concatenate([[0], np.cumsum(counts_1)]) * 8
            maps_1 = tuple(maps_1[:6]) + (maps_1[6], maps_1[7][off[lo]:off[hi]], counts_1)
            maps_s = tuple(maps_s[:6]) + (maps_s[6], maps_s[7], maps_s[8])
            if not (np.array_equal(maps_s[6], maps_1[6][lo:hi]) and np.array_equal(maps_s[7], maps_1[7])
                    and np.array_equal(maps_s[8], maps_1[8][lo:hi])):
                print("rank %d dtype %s %s: orphan readings differ" % (rank, dtype, exchange), flush=True)
                ok = False
            if int(maps_1[6].max()) == 0:
                print("spawn run stored no reading: the test exercised nothing", flush=True)
                ok = False
        for name, a, b in zip(("mean", "covp", "covc", "meta", "ids", "nlive"), maps_s, maps_1):
            if not np.array_equal(a, b[lo:hi]):
                print("rank %d dtype %s %s skew=%s: landmark %s differs after %d frames" % (
                    rank, dtype, exchange, skew, name, frames), flush=True)
                ok = False
        b_s, b_1 = fs_s.best_particle(), fs_1.best_particle()
        if b_s != b_1:
            print("rank %d best particle differs" % rank, b_s, b_1, flush=True)
            ok = False
        if rank == 0:
            print("dtype %s exchange %s skew %s spawn %s: %d particles over %d ranks, %d frames, %d particle migrations, "
                  "identical=%s" % (dtype, exchange, skew, spawn, M, world, frames, int(moved_t.item()), ok), flush=True)
        if world > 1 and int(moved_t.item()) == 0:
            print("no particle crossed a shard boundary: the test exercised nothing", flush=True)
            ok = False
        if skew and int(moved_t.item()) < frames * M // 8:
            print("skewed run moved only %d particles" % int(moved_t.item()), flush=True)
            ok = False
        fs_s.close()
    # a receive buffer that is too small must be reported, not overrun
    if world > 1:
        from parakeet_slam_b200._lib import ParakeetLibraryError
        try:
            fs_o, _, _, _ = run(ShardedFastSLAM, "f32", True, exchange="peer", exchange_capacity=16)
            fs_o.check_exchange()
            print("rank %d: exchange overflow was not reported" % rank, flush=True)
            ok = False
        except ParakeetLibraryError as exc:
            if rank == 0:
                print("overflow reported as expected:", str(exc)[:80], flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
