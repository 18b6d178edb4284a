"""CPU-only, world_size 2 over gloo: the host-side logic of the sharded bundle adjuster.

  * ptam_bundle_shard_plan (a pure host function of the product library): contiguous, covering,
    measurement-balanced point ranges, same answer on every rank;
  * the exchange step: the reduced camera system is additive over point shards, so the all-reduced
    sum of per-shard partial (S, vE) equals the oracle's full system;
  * the distributed order statistic: a multi-pass MSB radix select whose digit histograms are
    all-reduced finds the exact floor(n/2)-th smallest of the union of the shards' errors (the
    protocol k_ba_hist_pick run on the device).
The 2-GPU NCCL run of the same path is tests/test_bundle_sharded_gpu.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Bundle, product_lib, shard_plan


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _subgraph(g, lo, hi):
    keep = (g["meas_point"] >= lo) & (g["meas_point"] < hi)
    h = dict(g)
    for k in ("meas_cam", "meas_point", "meas_uv", "meas_sigma_sq"):
        h[k] = g[k][keep]
    return h


def _radix_select_allreduce(local_vals, k_of_n):
    """exact k-th smallest of the union of all ranks' non-negative doubles; k_of_n(n) -> k."""
    keys = np.ascontiguousarray(local_vals, np.float64).view(np.uint64)
    prefix, k = np.uint64(0), None
    for p in range(6):  # digit widths 11,11,11,11,11,9 as in k_ba_hist_pick
        shift = np.uint64(53 - 11 * p if p < 5 else 0)
        width = np.uint64(11 if p < 5 else 9)
        sel = keys if p == 0 else keys[(keys >> (shift + width)) == (prefix >> (shift + width))]
        hist = np.bincount(((sel >> shift) & ((np.uint64(1) << width) - np.uint64(1))).astype(np.int64), minlength=2048)
        t = torch.from_numpy(hist.astype(np.int64))
        dist.all_reduce(t)
        hist = t.numpy()
        if p == 0:
            k = k_of_n(int(hist.sum()))
        cum = np.cumsum(hist)
        b = int(np.searchsorted(cum, k, side="right"))
        k -= int(cum[b - 1]) if b else 0
        prefix |= np.uint64(b) << shift
    return np.array([prefix], np.uint64).view(np.float64)[0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from test_oracle_bundle import _dense_reduced_system
        from oracle.binding import oracle_lib
        g = synth.make_ba_graph(5, 60, 200, seed=11)
        plan = shard_plan(product_lib(), len(g["points"]), g["meas_point"], world)
        lo, hi = int(plan[rank]), int(plan[rank + 1])
        # the oracle's full system (every rank computes it; it is the checker)
        b = Bundle(oracle_lib(), g["width"], g["height"])
        b.add_graph(g)
        b.begin()
        b.lm_step()
        st = b.stats()
        n = 6 * int((g["cam_fixed"] == 0).sum())
        # reduced_system re-assembles with the lambda of the last trial
        S_ref, vE_ref = b.reduced_system(n)
        # the state the step started from is the input graph: partial systems from it
        lam_trial = None
        # find the lambda of the last trial: lambda after a good step = trial * 0.3
        lam = 1e-4
        trials = st.lambda_trials
        f = 2.0
        for _ in range(trials - 1):
            lam *= f; f *= 2
        S_part, vE_part = _dense_reduced_system(_subgraph(g, lo, hi), st.sigma_squared, lam)
        ts, tv = torch.from_numpy(S_part.copy()), torch.from_numpy(vE_part.copy())
        dist.all_reduce(ts); dist.all_reduce(tv)
        err_S = float(np.abs(ts.numpy() - S_ref).max() / np.abs(S_ref).max())
        err_v = float(np.abs(tv.numpy() - vE_ref).max() / np.abs(vE_ref).max())
        # distributed median against a sort of the union
        rng = np.random.default_rng(100 + rank)
        vals = np.abs(rng.normal(0, 1, 5000 + 37 * rank)) ** 2
        vals[:50] = 0.25  # ties
        med = _radix_select_allreduce(vals, lambda n_: n_ // 2)
        gathered = [None] * world
        dist.all_gather_object(gathered, vals)
        allv = np.sort(np.concatenate(gathered))
        q.put((rank, plan.tolist(), err_S, err_v, med == allv[len(allv) // 2]))
    finally:
        dist.destroy_process_group()


def test_shard_plan_properties():
    lib = product_lib()
    rng = np.random.default_rng(0)
    for P, M, world in ((100, 700, 2), (1000, 6000, 8), (3, 9, 8), (50, 0, 4), (1, 5, 2)):
        mp_ = np.sort(rng.integers(0, P, M)).astype(np.int32)
        plan = shard_plan(lib, P, mp_, world)
        assert plan[0] == 0 and plan[-1] == P and np.all(np.diff(plan) >= 0)
        if M >= 50 * world:
            per = np.array([((mp_ >= plan[r]) & (mp_ < plan[r + 1])).sum() for r in range(world)])
            assert per.max() - per.min() <= 2 * np.bincount(mp_).max() + 1
    assert np.array_equal(shard_plan(lib, 10, np.arange(10, dtype=np.int32), 1), [0, 10])
    with pytest.raises(Exception):
        shard_plan(lib, 4, np.array([5], np.int32), 2)


def test_two_rank_exchange_and_select_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out[0][1] == out[1][1]  # same plan on both ranks
    for rank, plan, err_S, err_v, med_ok in out:
        assert err_S < 1e-5 and err_v < 1e-5, (err_S, err_v)  # numeric differentiation in the numpy derivation
        assert med_ok
