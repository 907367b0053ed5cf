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


# ---- peer exchange: the device-resident plan (same C++ function on host and device) -------------------
def _xplan_host(E, Ml, rank, cap):
    import ctypes
    from parakeet_slam_b200 import _lib
    lib = _lib.load()
    G = len(E) - 1
    Ea = (ctypes.c_longlong * (G + 1))(*[int(v) for v in E])
    out = (ctypes.c_longlong * _lib.PK_XPLAN_LONGS)()
    _lib.check(lib.pk_exchange_plan_host(Ea, G, rank, Ml, cap, out), "pk_exchange_plan_host")
    return np.array(list(out), dtype=np.int64)


def _offspring_case(kind, M, rs):
    if kind == "uniform":
        return np.ones(M, dtype=np.int64)
    if kind == "skewed":
        w = rs.gamma(0.3, size=M)
        off = np.diff(np.floor(np.concatenate([[0], np.cumsum(w)]) / w.sum() * M + rs.uniform())).astype(np.int64)
        off[-1] += M - off.sum()
        return off
    off = np.zeros(M, dtype=np.int64)
    if kind == "one_hot":
        off[rs.randint(M)] = M
    elif kind == "front":
        off[:M // 4] = 4
    elif kind == "back":
        off[-(M // 2):] = 2
    elif kind == "alternate_eighths":    # every other eighth of the filter dies
        q = M // 8
        for b in range(0, 8, 2):
            off[b * q:(b + 1) * q] = 2
    return off


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("kind", ["uniform", "skewed", "one_hot", "front", "back", "alternate_eighths"])
def test_device_plan_matches_host_plan_and_push_order(G, kind):
    """pk_exchange_plan_host (the function the plan kernel runs) agrees with plan_exchange, and the
    push kernel's index arithmetic -- emulated here step by step -- fills every rank's receive buffer
    in exactly the order its output window expects."""
    from parakeet_slam_b200 import _lib
    rs = np.random.RandomState(G * 13 + len(kind))
    Ml = 64
    M = G * Ml
    off = _offspring_case(kind, M, rs)
    assert off.sum() == M
    E = _emitted(off, G)
    anc = np.repeat(np.arange(M), off)
    out_lo_global = np.concatenate([[0], np.cumsum(off)])[:-1]
    X = [_xplan_host(E, Ml, g, Ml) for g in range(G)]
    recv = [dict() for _ in range(G)]
    for g in range(G):
        x, p = X[g], plan_exchange(E, Ml, g)
        assert x[_lib.PK_XP_OVERFLOW] == 0
        assert (x[_lib.PK_XP_N_LO], x[_lib.PK_XP_N_LOC], x[_lib.PK_XP_N_HI]) == (p["n_lo"], p["n_loc"], p["n_hi"])
        assert (x[_lib.PK_XP_EMIT_LO], x[_lib.PK_XP_EMIT_N]) == (p["emit_lo"], p["emit_n"])
        assert x[_lib.PK_XP_N_BELOW] == sum(p["send"][:g]) and x[_lib.PK_XP_N_ABOVE] == sum(p["send"][g + 1:])
        assert x[_lib.PK_XP_N_SEND] == sum(p["send"]) - p["send"][g] and x[_lib.PK_XP_N_IN] == Ml - p["n_loc"]
        for h in range(G):
            ph = plan_exchange(E, Ml, h)
            assert x[_lib.PK_XP_RANK_LO + h] == ph["n_lo"] and x[_lib.PK_XP_RANK_LOC + h] == ph["n_loc"]
        # push_headers_kernel, one send item at a time
        out_lo = out_lo_global[g * Ml:(g + 1) * Ml]
        for j in range(int(x[_lib.PK_XP_N_SEND])):
            nb_ = int(x[_lib.PK_XP_N_BELOW])
            k = int(x[_lib.PK_XP_EMIT_LO]) + j if j < nb_ else int(x[_lib.PK_XP_ABOVE_START]) + (j - nb_)
            a = int(np.searchsorted(out_lo, k, side="right")) - 1           # upper_bound - 1
            h = k // Ml
            assert h != g
            k_local = k - h * Ml
            r = k_local if h > g else k_local - int(x[_lib.PK_XP_RANK_LOC + h])
            assert r not in recv[h]
            recv[h][r] = g * Ml + a
    for h in range(G):
        win = anc[h * Ml:(h + 1) * Ml]
        foreign = win[(win < h * Ml) | (win >= (h + 1) * Ml)]
        assert sorted(recv[h]) == list(range(len(foreign)))
        assert np.array_equal(np.array([recv[h][r] for r in range(len(foreign))], dtype=np.int64), foreign)


def test_device_plan_flags_overflow():
    from parakeet_slam_b200 import _lib
    Ml, G = 64, 4
    off = _offspring_case("front", G * Ml, np.random.RandomState(0))
    E = _emitted(off, G)
    need = max(Ml - plan_exchange(E, Ml, g)["n_loc"] for g in range(G))
    assert need > 8
    for g in range(G):
        x = _xplan_host(E, Ml, g, 8)
        assert x[_lib.PK_XP_OVERFLOW] == 1 and x[_lib.PK_XP_N_SEND] == 0 and x[_lib.PK_XP_N_IN] == 0
        assert _xplan_host(E, Ml, g, need)[_lib.PK_XP_OVERFLOW] in (0, 1)
    assert all(_xplan_host(E, Ml, g, Ml)[_lib.PK_XP_OVERFLOW] == 0 for g in range(G))


@pytest.mark.parametrize("seed", range(12))
def test_device_plan_random_offspring(seed):
    """Random rank counts (incl. odd ones), random heavy-tailed offspring: the device plan's window split and send runs
    add up on every rank and agree with the host plan."""
    from parakeet_slam_b200 import _lib
    rs = np.random.RandomState(100 + seed)
    G = int(rs.choice([2, 3, 5, 7, 8, 16]))
    Ml = int(rs.choice([32, 96, 256]))
    M = G * Ml
    w = rs.pareto(0.7, size=M) * (rs.uniform(size=M) < rs.uniform(0.05, 1.0))
    if w.sum() == 0:
        w[rs.randint(M)] = 1.0
    C = np.cumsum(w)
    anc = np.minimum(np.searchsorted(C, (rs.uniform() + np.arange(M)) * C[-1] / M, side="left"), M - 1)
    off = np.bincount(anc, minlength=M)
    E = _emitted(off, G)
    total_send = total_in = 0
    for g in range(G):
        x, p = _xplan_host(E, Ml, g, Ml), plan_exchange(E, Ml, g)
        assert x[_lib.PK_XP_OVERFLOW] == 0
        assert x[_lib.PK_XP_N_LO] + x[_lib.PK_XP_N_LOC] + x[_lib.PK_XP_N_HI] == Ml
        assert x[_lib.PK_XP_N_BELOW] + x[_lib.PK_XP_N_LOC] + x[_lib.PK_XP_N_ABOVE] == x[_lib.PK_XP_EMIT_N]
        assert (x[_lib.PK_XP_N_LO], x[_lib.PK_XP_N_LOC]) == (p["n_lo"], p["n_loc"])
        assert x[_lib.PK_XP_N_SEND] == sum(p["send"]) - p["send"][g]
        if x[_lib.PK_XP_N_ABOVE]:
            assert x[_lib.PK_XP_ABOVE_START] == max(E[g], (g + 1) * Ml)
        total_send += int(x[_lib.PK_XP_N_SEND])
        total_in += int(x[_lib.PK_XP_N_IN])
    assert total_send == total_in          # every migrating particle is sent once and received once
