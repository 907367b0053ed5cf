"""Particle-sharded FastSLAM: one process per GPU, particles split in contiguous index ranges.

Motion, association, EKF updates and weighting touch only a particle's own state, so they run
unchanged on every shard.  Two steps couple particles:

* resampling -- every rank needs the global prefix of the weights.  Ranks ``all_gather`` their
  per-block weight totals (8 B per 1024 particles) and each one folds ALL block totals in global
  order with the same fixed tree (``pk_resample_thresholds``), so the thresholds, and therefore the
  ancestors, are bit-identical to the single-GPU filter whatever the number of ranks.  Resampled
  particles whose output slot lives on another rank move with one ``all_to_all_single`` per payload
  (pose, aux, landmark blocks) straight out of / into the slot pools; the send / receive counts
  follow from the emitted-output counts at the rank boundaries, which every rank already holds --
  no extra count exchange;
* ``summary()`` / ``best_particle()`` -- an all-reduce of five doubles / an all-gather of two.

Two exchange engines implement that step with identical results:

``exchange="peer"`` (default)  every count stays on the device.  Each rank owns one CUDA-IPC shared
  allocation ``[flags | all block totals | receive buffer]``; the weight-scan kernel stores its block
  totals straight into every peer's copy (a fused all-gather), a flag barrier in peer memory replaces
  the collective's synchronisation, a one-thread kernel derives the exchange plan from the K3b
  output, and migrating particles are pushed by this library's TMA block mover directly into the
  destination rank's receive buffer over NVLink.  No NCCL call and no host round trip per frame, so
  the host keeps running ahead of the GPU exactly as in the single-GPU filter.
``exchange="nccl"``  ``all_gather`` of the block totals, one small D2H copy of the emitted-output
  counts, ``all_to_all_single`` of the packed records.  Unbounded exchange size; the host waits for
  the GPU once per frame.

``plan_exchange`` is the pure host-side arithmetic of the exchange; it is tested on CPU with a
world-size-2 ``gloo`` group and against the device plan (``pk_exchange_plan_host``) in
``tests/test_sharding_cpu.py``.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib
from .core import FastSLAM


def plan_exchange(emitted_before, particles_per_rank, rank):
    """Exchange plan of one resampling step.

    ``emitted_before[g]`` (length G+1) = number of output slots whose ancestor lives on a rank
    < g; rank g's offspring therefore occupy the global output slots
    ``[emitted_before[g], emitted_before[g+1])`` and output slot k is owned by rank
    ``k // particles_per_rank``.  Returns a dict with, for ``rank``:

    ``send[h]``  particles this rank sends to rank h (``send[rank]`` stay local),
    ``recv[g]``  particles it receives from rank g,
    ``send_start[h]`` offset of the run for rank h inside this rank's offspring list,
    ``n_lo`` / ``n_loc`` / ``n_hi`` the split of its own output window.
    """
    E = [int(v) for v in emitted_before]
    G = len(E) - 1
    Ml = int(particles_per_rank)

    def overlap(g, h):
        lo = max(E[g], h * Ml)
        hi = min(E[g + 1], (h + 1) * Ml)
        return max(0, hi - lo)

    send = [overlap(rank, h) for h in range(G)]
    recv = [overlap(g, rank) for g in range(G)]
    send_start = [max(E[rank], h * Ml) - E[rank] if send[h] else 0 for h in range(G)]
    n_lo = sum(recv[:rank])
    n_loc = recv[rank]
    n_hi = sum(recv[rank + 1:])
    assert n_lo + n_loc + n_hi == Ml, (E, Ml, rank)
    return dict(send=send, recv=recv, send_start=send_start, n_lo=n_lo, n_loc=n_loc, n_hi=n_hi,
                emit_lo=E[rank], emit_n=E[rank + 1] - E[rank])


class ShardedFastSLAM(FastSLAM):
    """``FastSLAM`` over ``torch.distributed`` (NCCL): ``num_particles`` is the GLOBAL particle count,
    each rank holds ``num_particles / world_size`` of them (a multiple of 1024)."""

    def __init__(self, preset_features=[], *, num_particles, group=None, exchange="peer", exchange_capacity=None,
                 barrier_timeout_s=20.0, **kw):
        import torch.distributed as dist

        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        import os as _os
        self._timing_on = _os.environ.get("PK_TIMING", "") == "1"
        self._timing = []

        self._dist = dist
        self._group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        M_total = int(num_particles)
        if M_total % self.world_size or (M_total // self.world_size) % _lib.PK_SCAN_BLOCK:
            raise ValueError("num_particles must be a multiple of world_size * %d" % _lib.PK_SCAN_BLOCK)
        self.num_particles_total = M_total
        super().__init__(preset_features, num_particles=M_total // self.world_size, **kw)
        self.particle_offset = self.rank * self.num_particles
        torch = self._torch
        # runs of more than 16 equal ancestors are listed for the fill kernel: a window of n_out output slots holds
        # at most n_out / 17 of them, and here n_out goes up to the GLOBAL particle count (include/parakeet_b200.h)
        self._big_runs = torch.zeros((4 + 3 * min(self.num_particles, M_total // 17 + 2),), dtype=torch.int64,
                                     device=self._device)
        G, nb = self.world_size, self._nb
        dev = self._device
        self._all_sums = torch.zeros((G * nb,), dtype=torch.float64, device=dev)
        self._all_prefix = torch.zeros((G * nb, 2), dtype=torch.float64, device=dev)
        self._all_count = torch.zeros((G * nb + 1,), dtype=torch.int64, device=dev)
        self._emit = torch.zeros((max(2 * self.num_particles, 1),), dtype=torch.int64, device=dev)
        self._record_bytes = int(self._lib.pk_particle_record_bytes(self.capacity, self._dt))
        self._exchange_cap = 0
        self._send_buf = self._recv_buf = self._pack_ws = None
        self._rank_idx = torch.arange(0, G * nb + 1, nb, device=dev)
        self._last_plan = None
        self._barrier_timeout_s = float(barrier_timeout_s)
        self._peer = None
        if exchange == "peer":
            if G > _lib.PK_MAX_RANKS:
                raise ValueError("the peer exchange serves at most %d ranks" % _lib.PK_MAX_RANKS)
            self._setup_peer(exchange_capacity)

    # -- peer memory -------------------------------------------------------------------------------
    def _setup_peer(self, exchange_capacity):
        """One CUDA-IPC allocation per rank: [flags 4 KiB | all block totals | receive buffer], mapped by
        every other rank; device tables hold each region's address on every rank."""
        torch, lib, dist = self._torch, self._lib, self._dist
        G, me, nb, Ml, dev = self.world_size, self.rank, self._nb, self.num_particles, self._device
        rec = self._record_bytes
        if exchange_capacity is None:
            # Records a rank may receive per frame.  A receive buffer that holds the WHOLE shard can never overflow
            # (a rank cannot receive more than its own output window), so that is the default whenever it fits into
            # 80 % of the memory still free beside the landmark pool (config 4: 36.5 GB beside a 36.5 GB pool); only a
            # shard too big for that gets a smaller buffer, and only then can a frame report PK_PEER_OVERFLOW.
            free, _total = torch.cuda.mem_get_info(dev)
            exchange_capacity = min(Ml, max(4096, int(0.8 * free) // rec))
        cap = int(min(max(int(exchange_capacity), 1), Ml))
        self.exchange_capacity = cap
        off_sums = 4096
        off_recv = off_sums + ((G * nb * 8 + 255) // 256) * 256
        total = off_recv + cap * rec
        with self._on_device():
            base = ctypes.c_void_p()
            _lib.check(lib.pk_peer_alloc(total, ctypes.byref(base)), "pk_peer_alloc")
            handle = (ctypes.c_ubyte * _lib.PK_PEER_HANDLE_BYTES)()
            _lib.check(lib.pk_peer_export(base, handle), "pk_peer_export")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            every = torch.zeros((G, _lib.PK_PEER_HANDLE_BYTES), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(every.view(-1), mine, group=self._group)
            every = every.cpu().numpy()
            bases, opened, failure = [], [], None
            for g in range(G):
                if g == me:
                    bases.append(int(base.value))
                    continue
                h = (ctypes.c_ubyte * _lib.PK_PEER_HANDLE_BYTES)(*[int(v) for v in every[g]])
                p = ctypes.c_void_p()
                if failure is None and lib.pk_peer_open(h, ctypes.byref(p)) != 0:
                    failure = "pk_peer_open(rank %d): %s" % (g, _lib.last_error())
                if failure is None:
                    opened.append(p)
                bases.append(int(p.value or 0))
            # all ranks agree on the outcome before anyone touches peer memory
            ok = torch.tensor([0 if failure else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self._group)
            if int(ok.item()) == 0:
                for p in opened:
                    lib.pk_peer_close(p)
                dist.barrier(group=self._group)
                lib.pk_peer_free(base)
                raise _lib.ParakeetLibraryError("peer memory could not be mapped on every rank (%s); use "
                                                "exchange='nccl'" % (failure or "another rank failed"))
            i64 = torch.int64
            tab = lambda off: torch.tensor([b + off for b in bases], dtype=i64, device=dev)
            self._peer = dict(base=base, opened=opened, bases=bases, total=total,
                              flags_tab=tab(0), sums_tab=tab(off_sums), recv_tab=tab(off_recv),
                              sums_ptr=bases[me] + off_sums, recv_ptr=bases[me] + off_recv, epoch=0)
            self._xplan = torch.zeros((_lib.PK_XPLAN_LONGS,), dtype=i64, device=dev)
            self._peer_status = torch.zeros((_lib.PK_PEER_STATUS_WORDS,), dtype=i64, device=dev)
            self._anc_window = torch.zeros((Ml,), dtype=i64, device=dev)
            self._send_capacity = max(1, (G - 1) * cap)   # one particle may own every output slot
            self._push_ws = torch.zeros((4 * self._send_capacity,), dtype=torch.int32, device=dev)
            self._push_stream = torch.cuda.Stream(device=dev)
            self._ev_plan = torch.cuda.Event()
            self._ev_push = torch.cuda.Event()
            torch.cuda.synchronize(dev)
        dist.barrier(group=self._group)   # every mapping exists and every flag is zero before the first frame

    def close(self):
        """Unmap the peers' allocations and free this rank's (collective: call on every rank)."""
        if self._peer is None:
            return
        torch, lib = self._torch, self._lib
        with self._on_device():
            torch.cuda.synchronize(self._device)
            self._dist.barrier(group=self._group)
            for p in self._peer["opened"]:
                lib.pk_peer_close(p)
            self._dist.barrier(group=self._group)
            lib.pk_peer_free(self._peer["base"])
        self._peer = None

    def _barrier(self, st):
        pr = self._peer
        pr["epoch"] += 1
        _lib.check(self._lib.pk_peer_barrier(_lib.ptr(pr["flags_tab"]), self.rank, self.world_size, pr["epoch"],
                                             self._barrier_timeout_s, _lib.ptr(self._peer_status), st),
                   "pk_peer_barrier")

    def check_exchange(self):
        """Raise if an earlier frame overflowed the exchange capacity or a peer missed a barrier (the
        status word is sticky; reading it synchronises the stream)."""
        if self._peer is None:
            return
        bits = int(self._peer_status[0].item())
        if bits & _lib.PK_PEER_OVERFLOW:
            raise _lib.ParakeetLibraryError(
                "a resampling step had to move more than exchange_capacity=%d particles between ranks; the filter "
                "state is void -- rebuild with a larger exchange_capacity or exchange='nccl'" % self.exchange_capacity)
        if bits & _lib.PK_PEER_TIMEOUT:
            raise _lib.ParakeetLibraryError("a rank did not reach a resampling barrier within %.0f s"
                                            % self._barrier_timeout_s)

    @property
    def last_plan(self):
        """Exchange plan of the last resampling step as a dict (peer mode: read back from the device)."""
        if self._peer is None or self._frame == 0:
            return self._last_plan
        x = self._xplan.cpu().numpy()
        G = self.world_size
        return dict(n_lo=int(x[_lib.PK_XP_N_LO]), n_loc=int(x[_lib.PK_XP_N_LOC]), n_hi=int(x[_lib.PK_XP_N_HI]),
                    emit_lo=int(x[_lib.PK_XP_EMIT_LO]), emit_n=int(x[_lib.PK_XP_EMIT_N]),
                    n_below=int(x[_lib.PK_XP_N_BELOW]), n_above=int(x[_lib.PK_XP_N_ABOVE]),
                    overflow=bool(x[_lib.PK_XP_OVERFLOW]),
                    rank_n_lo=[int(v) for v in x[_lib.PK_XP_RANK_LO:_lib.PK_XP_RANK_LO + G]],
                    rank_n_loc=[int(v) for v in x[_lib.PK_XP_RANK_LOC:_lib.PK_XP_RANK_LOC + G]])

    def _tick(self, name):
        """PK_TIMING=1: CUDA events between the launches of the peer resampling chain (debugging aid; timing_report()
        averages them).  A no-op otherwise."""
        if not self._timing_on:
            return
        ev = self._torch.cuda.Event(enable_timing=True)
        ev.record()
        self._timing.append((name, ev))

    def timing_report(self):
        """Average milliseconds per segment of the peer resampling chain (needs PK_TIMING=1)."""
        self._torch.cuda.synchronize(self._device)
        acc, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(self._timing, self._timing[1:]):
            if n1 == "start":
                continue
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        return {k: acc[k] / cnt[k] for k in acc}

    def barrier_wait_report(self, reset=True):
        """Average milliseconds per frame this rank spent inside the two flag barriers of the peer resampling chain
        (from kernel entry to all flags seen; accumulated on the device by the kernels that hold the barriers)."""
        if self._peer is None:
            return {}
        s = self._peer_status.cpu().numpy()
        n = max(1, int(s[3]))
        out = dict(frames=int(s[3]), totals_barrier_ms=float(s[1]) / n * 1e-6, pushes_barrier_ms=float(s[2]) / n * 1e-6)
        if reset:
            self._peer_status[1:] = 0
        return out

    def _all_reduce_weight_stat(self, tensor, op):
        """The weight normaliser across shards: NCCL all-reduce of the maximum log weight / of sum w and sum w^2."""
        dist = self._dist
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self._group)

    def _resample_peer(self):
        """low_variance_resample with every count on the device and both exchanges over peer memory."""
        torch, lib, dist = self._torch, self._lib, self._dist
        Ml, Mt, G, me, nb = self.num_particles, self.num_particles_total, self.world_size, self.rank, self._nb
        pr = self._peer
        with self._lock, self._on_device():
            u01 = float(self._uniform())  # every rank must draw the same value (same seed / same source)
            st = self._stream()
            cur, nxt = self._cur, 1 - self._cur
            tick = self._tick
            tick("start")
            if self.weights == "log":
                self._normalise_log_weights()
            pose_in, aux_in, slot_in = self._pose[cur], self._aux[cur], self._slot[cur]
            status = _lib.ptr(self._peer_status)
            # K3a + all-gather of the block totals in one kernel (stores into every rank's copy)
            _lib.check(lib.pk_weight_scan_publish(_lib.ptr(pose_in), Ml, _lib.ptr(self._cumsum),
                                                  _lib.ptr(self._block_sums), _lib.ptr(pr["sums_tab"]), me, G, st),
                       "pk_weight_scan_publish")
            tick("scan+publish")
            # one single-CTA kernel: flag barrier (all ranks' totals have arrived), K3b over all ranks' totals, and
            # the exchange plan derived from the emitted-output counts at the rank boundaries
            pr["epoch"] += 1
            _lib.check(lib.pk_resample_thresholds_peer(pr["sums_ptr"], G * nb, Mt, u01, _lib.ptr(self._plan),
                                                       _lib.ptr(self._all_prefix), _lib.ptr(self._all_count),
                                                       _lib.ptr(pr["flags_tab"]), me, G, pr["epoch"],
                                                       self._barrier_timeout_s, Ml, self.exchange_capacity,
                                                       _lib.ptr(self._xplan), status, st),
                       "pk_resample_thresholds_peer")
            tick("barrier+thresholds+plan")
            # ancestors of my own output window (entries owned by other ranks' particles stay untouched), each local
            # particle's offspring inside the window and the dead-particle scan, in one kernel
            _lib.check(lib.pk_resample_plan(_lib.ptr(self._cumsum), Ml, self.particle_offset, me * nb,
                                            _lib.ptr(self._plan), _lib.ptr(self._all_prefix),
                                            _lib.ptr(self._all_count), Mt, me * Ml, Ml,
                                            _lib.ptr(self._out_lo), _lib.ptr(self._offspring),
                                            _lib.ptr(self._anc_window), _lib.ptr(self._gather_ws), st),
                       "pk_resample_plan")
            tick("resample_plan")
            # offspring that live on other ranks: header + landmark block straight into their receive buffers -- on a
            # side stream, so that they overlap the free list and the slot assignment of the gather (the copies of the
            # local duplicates wait for them: a sender that is dead locally donates its block)
            main = torch.cuda.current_stream(self._device)
            side = self._push_stream
            self._ev_plan.record(main)
            side.wait_event(self._ev_plan)
            _lib.check(lib.pk_push_particles(_lib.ptr(self._xplan), _lib.ptr(self._out_lo), Ml, me, _lib.ptr(pose_in),
                                             _lib.ptr(aux_in), _lib.ptr(slot_in), _lib.ptr(self._pool), self.capacity,
                                             self._dt, _lib.ptr(pr["recv_tab"]), self._send_capacity,
                                             _lib.ptr(self._push_ws), side.cuda_stream), "pk_push_particles")
            # split-phase second barrier: the flag "my pushes have landed" is posted right behind them on the side
            # stream, the wait sits inside the gather in front of the part that reads the receive buffer -- a rank
            # with many local duplicates to copy does not hold the others up
            pr["epoch"] += 1
            _lib.check(lib.pk_peer_post(_lib.ptr(pr["flags_tab"]), me, G, pr["epoch"], side.cuda_stream), "pk_peer_post")
            self._ev_push.record(side)
            _lib.check(lib.pk_resample_gather_peer(
                _lib.ptr(self._xplan), _lib.ptr(self._anc_window), _lib.ptr(self._out_lo), _lib.ptr(self._offspring), Ml,
                self.particle_offset, _lib.ptr(pose_in), _lib.ptr(self._pose[nxt]), _lib.ptr(aux_in),
                _lib.ptr(self._aux[nxt]), _lib.ptr(slot_in), _lib.ptr(self._slot[nxt]), pr["recv_ptr"],
                self.exchange_capacity, _lib.ptr(self._pool), self.capacity, self._dt, _lib.ptr(self._gather_ws),
                _lib.ptr(self._n_copied), _lib.ptr(pr["flags_tab"]), me, G, pr["epoch"], self._barrier_timeout_s,
                status, self._ev_push.cuda_event, st), "pk_resample_gather_peer")
            tick("barrier+gather")
            self._cur = nxt
            if self.keep_trace:
                # debugging / parity traces only: every rank scatters its offspring into a global list
                anc_global = torch.zeros((Mt,), dtype=torch.int64, device=self._device)
                _lib.check(lib.pk_resample_ancestors(_lib.ptr(self._cumsum), Ml, self.particle_offset, me * nb,
                                                     _lib.ptr(self._plan), _lib.ptr(self._all_prefix),
                                                     _lib.ptr(self._all_count), Mt, 0, Mt,
                                                     _lib.ptr(self._out_lo), _lib.ptr(self._offspring),
                                                     _lib.ptr(anc_global), _lib.ptr(self._big_runs), st),
                           "pk_resample_ancestors(trace)")
                dist.all_reduce(anc_global, group=self._group)
                self.last_ancestors = anc_global[me * Ml:(me + 1) * Ml].clone()

    # noise for the parity mode: every rank draws the GLOBAL block and keeps its slice, so the
    # stream consumed is the one a single-process filter would consume
    def _draw_noise(self, M):
        if self._noise == "philox":
            return None
        lo = self.particle_offset
        if self._noise == "numpy":
            z = np.random.standard_normal((self.num_particles_total, 3))
        else:
            z = np.ascontiguousarray(self._noise(self.num_particles_total), dtype=np.float64)
        return np.ascontiguousarray(z[lo:lo + M])

    def low_variance_resample(self):
        """``FastSLAM.low_variance_resample`` (``prkt_core_v2.py:210-252``) over all shards."""
        if self._peer is not None:
            return self._resample_peer()
        torch, lib, dist = self._torch, self._lib, self._dist
        Ml, Mt, G, me, nb = self.num_particles, self.num_particles_total, self.world_size, self.rank, self._nb
        dev = self._device
        with self._lock, self._on_device():
            u01 = float(self._uniform())  # every rank must draw the same value (same seed / same source)
            st = self._stream()
            cur, nxt = self._cur, 1 - self._cur
            pose_in, aux_in, slot_in = self._pose[cur], self._aux[cur], self._slot[cur]
            if self.weights == "log":
                self._normalise_log_weights()
            _lib.check(lib.pk_weight_scan(_lib.ptr(pose_in), Ml, _lib.ptr(self._cumsum), _lib.ptr(self._block_sums), st),
                       "pk_weight_scan")
            # weight normaliser: all ranks obtain all block totals, then fold them identically
            dist.all_gather_into_tensor(self._all_sums, self._block_sums, group=self._group)
            _lib.check(lib.pk_resample_thresholds(_lib.ptr(self._all_sums), G * nb, Mt, u01, _lib.ptr(self._plan),
                                                  _lib.ptr(self._all_prefix), _lib.ptr(self._all_count), st),
                       "pk_resample_thresholds")
            E = self._all_count[self._rank_idx].cpu().tolist()   # outputs emitted before each rank (one small D2H)
            plan = plan_exchange(E, Ml, me)
            self._last_plan = plan
            if plan["emit_n"] > self._emit.numel():
                self._emit = torch.zeros((plan["emit_n"],), dtype=torch.int64, device=dev)
            # my particles' offspring: emit[k - emit_lo] = global ancestor index, ascending
            _lib.check(lib.pk_resample_ancestors(_lib.ptr(self._cumsum), Ml, self.particle_offset, me * nb,
                                                 _lib.ptr(self._plan), _lib.ptr(self._all_prefix),
                                                 _lib.ptr(self._all_count), Mt, plan["emit_lo"], max(plan["emit_n"], 1),
                                                 _lib.ptr(self._out_lo), _lib.ptr(self._offspring), _lib.ptr(self._emit),
                                                 _lib.ptr(self._big_runs), st), "pk_resample_ancestors")
            emit = self._emit
            # ---- pack what leaves this rank (before any block is overwritten) -------------------
            send_counts = list(plan["send"])
            send_counts[me] = 0
            recv_counts = list(plan["recv"])
            recv_counts[me] = 0
            n_send, n_in = sum(send_counts), sum(recv_counts)
            rec = self._record_bytes
            self._ensure_exchange_buffers(n_send, n_in)
            # the runs for the ranks below me are contiguous at the head of the emit list and those for the
            # ranks above me at its tail (ancestors ascend), so two pack calls cover every destination
            n_below = sum(send_counts[:me])
            n_above = sum(send_counts[me + 1:])
            for count, start, off in ((n_below, 0, 0), (n_above, plan["emit_n"] - n_above, n_below)):
                if count:
                    _lib.check(lib.pk_pack_particles(_lib.ptr(emit[start:]), count, self.particle_offset,
                                                     _lib.ptr(pose_in), _lib.ptr(aux_in), _lib.ptr(slot_in),
                                                     _lib.ptr(self._pool), self.capacity, self._dt,
                                                     self._send_buf.data_ptr() + off * rec, _lib.ptr(self._pack_ws), st),
                               "pk_pack_particles")
            if G > 1:
                # cross-shard resampled particles: one record per particle, rank to rank over NVLink
                dist.all_to_all_single(self._recv_buf[:n_in], self._send_buf[:n_send], recv_counts, send_counts,
                                       group=self._group)
            # ---- local assignment: survivors keep their block, duplicates and arrivals take freed ones
            n_lo, n_loc = plan["n_lo"], plan["n_loc"]
            local_run = emit[plan["send_start"][me]:] if n_loc else None
            _lib.check(lib.pk_resample_gather_sharded(
                _lib.ptr(local_run), _lib.ptr(self._out_lo), _lib.ptr(self._offspring), Ml, self.particle_offset,
                n_lo, n_loc, _lib.ptr(pose_in), _lib.ptr(self._pose[nxt]), _lib.ptr(aux_in), _lib.ptr(self._aux[nxt]),
                _lib.ptr(slot_in), _lib.ptr(self._slot[nxt]), _lib.ptr(self._recv_buf), _lib.ptr(self._pool),
                self.capacity, self._dt, _lib.ptr(self._gather_ws), _lib.ptr(self._n_copied), st),
                "pk_resample_gather_sharded")
            self._cur = nxt
            if self.keep_trace:
                # debugging / parity traces only: assemble the global ancestor list and keep my window
                anc_global = torch.zeros((Mt,), dtype=torch.int64, device=dev)
                anc_global[plan["emit_lo"]:plan["emit_lo"] + plan["emit_n"]] = emit[:plan["emit_n"]]
                dist.all_reduce(anc_global, group=self._group)
                self.last_ancestors = anc_global[me * Ml:(me + 1) * Ml].clone()

    def _ensure_exchange_buffers(self, n_send, n_in):
        torch = self._torch
        need = max(n_send, n_in, 1)
        if need > self._exchange_cap:
            cap = max(need, 2 * self._exchange_cap, 256)
            self._exchange_cap = cap
            self._send_buf = torch.empty((cap, self._record_bytes), dtype=torch.uint8, device=self._device)
            self._recv_buf = torch.empty((cap, self._record_bytes), dtype=torch.uint8, device=self._device)
            self._pack_ws = torch.empty((3 * cap,), dtype=torch.int32, device=self._device)

    def summary(self):
        torch, lib, dist = self._torch, self._lib, self._dist
        with self._lock, self._on_device():
            _lib.check(lib.pk_summary_partial(_lib.ptr(self.pose), self.num_particles, _lib.ptr(self._out5),
                                              _lib.ptr(self._red_ws), self._stream()), "pk_summary_partial")
            dist.all_reduce(self._out5, group=self._group)
            s = self._out5.cpu().numpy()
            self.check_exchange()
        count = float(self.num_particles_total)
        return (float(s[0] / count), float(s[1] / count), math.atan2(float(s[2]), float(s[3])),)

    def best_particle(self):
        torch, lib, dist = self._torch, self._lib, self._dist
        with self._lock, self._on_device():
            _lib.check(lib.pk_best_particle(_lib.ptr(self.pose), self.num_particles, _lib.ptr(self._best2),
                                            _lib.ptr(self._red_ws), self._stream()), "pk_best_particle")
            mine = self._best2.clone()
            mine[1] += self.particle_offset
            allb = torch.zeros((self.world_size, 2), dtype=torch.float64, device=self._device)
            dist.all_gather_into_tensor(allb.view(-1), mine, group=self._group)
            b = allb.cpu().numpy()
        best = max(range(self.world_size), key=lambda g: (b[g, 0], -b[g, 1]))
        return int(b[best, 1]), float(b[best, 0])

    def stats(self):
        """Counters summed over ranks (a small all-reduce)."""
        s = super().stats()
        t = self._torch.tensor([s[k] for k in ("matched", "unmatched", "evaluated", "same_landmark", "promoted",
                                               "blocks_copied")], dtype=self._torch.int64, device=self._device)
        self._dist.all_reduce(t, group=self._group)
        out = dict(zip(("matched", "unmatched", "evaluated", "same_landmark", "promoted", "blocks_copied"),
                       [int(v) for v in t.cpu()]))
        out["flags"] = s["flags"]
        plan = self.last_plan
        out["migrated_in"] = int(plan["n_lo"] + plan["n_hi"]) if plan else 0
        self.check_exchange()
        return out
