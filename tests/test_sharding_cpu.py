"""Host-side logic of the particle-sharded filter on CPU: the exchange plan, alone and across a
world-size-2 gloo group (no GPU involved)."""
import os
import socket

import numpy as np
import pytest

from parakeet_slam_b200.sharded import plan_exchange


def _emitted(offspring, G):
    """emitted_before[g] from a global offspring-count vector split evenly over G ranks."""
    M = len(offspring)
    Ml = M // G
    cs = np.concatenate([[0], np.cumsum(offspring)])
    return [int(cs[g * Ml]) for g in range(G)] + [int(cs[-1])]


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", ["uniform", "skewed", "one_hot", "front", "back"])
def test_plan_is_consistent(G, kind):
    rs = np.random.RandomState(G * 7 + len(kind))
    Ml = 64
    M = G * Ml
    if kind == "uniform":
        off = np.ones(M, dtype=np.int64)
    elif kind == "skewed":
        w = rs.gamma(0.3, size=M)
        off = np.diff(np.floor(np.concatenate([[0], np.cumsum(w)]) / w.sum() * M + rs.uniform())).astype(np.int64)
        off[-1] += M - off.sum()
    elif kind == "one_hot":
        off = np.zeros(M, dtype=np.int64)
        off[rs.randint(M)] = M
    elif kind == "front":
        off = np.zeros(M, dtype=np.int64)
        off[:M // 4] = 4
    else:
        off = np.zeros(M, dtype=np.int64)
        off[-(M // 2):] = 2
    assert off.sum() == M and (off >= 0).all()
    E = _emitted(off, G)
    plans = [plan_exchange(E, Ml, g) for g in range(G)]
    anc = np.repeat(np.arange(M), off)              # global ancestor of every output slot
    for g, p in enumerate(plans):
        assert sum(p["send"]) == E[g + 1] - E[g]    # every offspring goes somewhere
        assert sum(p["recv"]) == Ml                  # every output slot is filled once
        for h in range(G):
            assert p["send"][h] == plans[h]["recv"][g]
            run = anc[E[g] + p["send_start"][h]: E[g] + p["send_start"][h] + p["send"][h]]
            # that run is exactly the part of rank h's output window whose ancestors live on rank g
            win = anc[h * Ml:(h + 1) * Ml]
            assert np.array_equal(run, win[(win >= g * Ml) & (win < (g + 1) * Ml)])
        assert (p["n_lo"], p["n_loc"], p["n_hi"]) == (sum(p["recv"][:g]), p["recv"][g], sum(p["recv"][g + 1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, seed, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(seed)          # same stream on every rank
        Ml = 1024
        M = world * Ml
        w = rs.gamma(0.2, size=M)
        u = rs.uniform()
        C = np.cumsum(w)
        anc = np.minimum(np.searchsorted(C, (u + np.arange(M)) * C[-1] / M, side="left"), M - 1)
        off = np.bincount(anc, minlength=M)
        E = _emitted(off, world)
        # the ranks only share their per-rank totals (here: emitted counts) -- check they agree
        mine = torch.tensor([E[rank + 1] - E[rank]], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert [int(t) for t in gathered] == [E[g + 1] - E[g] for g in range(world)]
        p = plan_exchange(E, Ml, rank)
        # payload: the global ancestor id of each migrating particle (stands for pose + map block)
        emit = anc[E[rank]:E[rank + 1]]
        send_counts = list(p["send"])
        send_counts[rank] = 0
        recv_counts = list(p["recv"])
        recv_counts[rank] = 0
        parts = [emit[p["send_start"][h]:p["send_start"][h] + send_counts[h]] for h in range(world)]
        send = torch.from_numpy(np.concatenate(parts).astype(np.int64))
        recv = torch.zeros(sum(recv_counts), dtype=torch.int64)
        dist.all_to_all_single(recv, send, recv_counts, send_counts)
        local = emit[p["send_start"][rank]:p["send_start"][rank] + p["n_loc"]]
        window = np.concatenate([recv.numpy()[:p["n_lo"]], local, recv.numpy()[p["n_lo"]:]])
        ok = np.array_equal(window, anc[rank * Ml:(rank + 1) * Ml])
        out.put((rank, bool(ok), int(p["n_lo"] + p["n_hi"])))
    finally:
        dist.destroy_process_group()


def test_exchange_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1234, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sum(m for _, _, m in res) > 0     # some particles really crossed the shard boundary
