"""Multi-rank path on CPU: world_size-2 ``gloo`` run of shard -> solve -> single gather -> unpack.  The solver
inside each rank is the oracle-backed stand-in (tests only); what is under test is the partitioning, the packed
result layout and the one-collective gather of vehicle_motion_planning..._b200/sharding.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import obca_testlib as common
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import scenario as sc, sharding


def test_shard_range_covers_batch():
    for B in (1, 7, 8, 1000, 8192, 65536):
        for G in (1, 2, 3, 4, 8):
            seen = []
            for r in range(G):
                lo, hi, per = sharding.shard_range(B, r, G)
                assert hi - lo <= per and 0 <= lo <= hi <= B
                seen += list(range(lo, hi))
            assert seen == list(range(B))


def test_packed_outputs_views_alias_one_buffer():
    p = sharding.PackedOutputs(5, 6, 10, 3)
    assert p.views["x"].shape == (5, 7, 3) and p.views["lam"].shape == (5, 7, 10) and p.views["mu"].shape == (5, 7, 12)
    assert p.views["status"].dtype == torch.int32 and p.views["status"].shape == (5,)
    for k, v in p.views.items():
        assert v.is_contiguous()
        v.fill_(3 if v.dtype == torch.int32 else 1.5)
    # every word of the buffer is covered by exactly the views (tail padding of the int32 pairs aside)
    assert (p.buf[:p.offsets["status"]] == 1.5).all()
    assert p.nbytes == p.words * 8
    assert (p.views_of(p.buf)["iters"] == 3).all()


class _TorchOracleSolver(common.OracleSolver):
    """``solve`` with torch CPU tensors in / packed views out (what BatchSolver.solve does on the GPU)."""

    def solve(self, x0, u0, xref, A, b0, db=None, T_max=None, term=None, uref=None, out=None, stream=None, Ts=None):
        n = lambda t: None if t is None else t.numpy()
        r = self.solve_host(n(x0), n(u0), n(xref), n(A), n(b0), n(db), T_max=n(T_max), term=n(term))
        for k, v in r.items():
            out[k].copy_(torch.as_tensor(v))
        return out


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = sc.make_batch(2, B)
        prm, a = common.batch_arrays(b)
        solver = _TorchOracleSolver(prm, a["edge_ptr"], B, nthreads=2)
        out, packed = sharding.solve_sharded(solver, a, B, rank, world, "cpu", dst=0)
        if rank == 0:
            q.put({k: v.numpy().copy() for k, v in out.items()})
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("B", [10, 13])
def test_two_rank_gloo_equals_single_rank(B):
    """sharded result == single-rank result bit for bit (instances are independent), ragged last shard included"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = sc.make_batch(2, B)
    prm, a = common.batch_arrays(b)
    ref = common.OracleSolver(prm, a["edge_ptr"], B).solve_host(a["x0"], a["u0"], a["xref"], a["A"], a["b0"], a["db"],
                                                                T_max=a["T_max"])
    for k in ("x", "u", "lam", "mu", "T", "obj", "status", "iters"):
        assert got[k].shape == ref[k].shape, k
        assert np.array_equal(got[k], ref[k]), k


def _cl_driver(dyn):
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import closed_loop as cl, demo_setting as ds
    s = ds.problemSetting("demo9"); s.senseDis = 8
    return cl.ClosedLoopBatch(s, dyn, N=5, Q_free=0.5, sense=8.0, max_steps=4,
                              solver_factory=lambda prm, ep, cap: common.OracleSolver(prm, ep, cap, nthreads=2))


def _cl_dyn(B):
    from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import closed_loop as cl
    dyn = cl.demo9_monte_carlo(B)
    dyn[:, 1] = np.linspace(12, 24, B); dyn[:, 6] = 0          # close enough to be detected within the first steps
    return dyn


def _cl_worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = sharding.closed_loop_sharded(_cl_driver, _cl_dyn(B), rank, world, dst=0)
        if rank == 0:
            q.put(out)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 1])
def test_closed_loop_shards_by_scenario(B):
    """cfg 4 across two ranks: scenario shards run independently, one gather of the packed logs; equal to one rank
    (B = 1: the second rank's shard is empty)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cl_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    drv = _cl_driver(_cl_dyn(B))
    ref = drv.run()
    one = sharding.closed_loop_sharded(_cl_driver, _cl_dyn(B), 0, 1)
    for k in ("traj", "mode", "x", "u", "Ts_opt", "steps", "failed", "reached"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
        assert np.array_equal(one[k], ref[k], equal_nan=True), k
    assert got["solves"] == ref["solves"] and got["launches"] >= ref["launches"]


def test_packed_outputs_over_an_existing_buffer():
    """rank 0 of a peer-to-peer gather lays its result views over its row of the gathered buffer"""
    import torch
    ref = sharding.PackedOutputs(5, 4, 6, 2)
    full = torch.zeros((2, 3, ref.words), dtype=torch.float64)
    p = sharding.PackedOutputs(5, 4, 6, 2, buf=full[1, 0])
    assert p.buf.data_ptr() == full[1, 0].data_ptr() and p.words == ref.words
    p.views["x"][...] = 1.5; p.views["status"][...] = 7
    assert float(full[1, 0].sum()) != 0.0 and float(full[0].abs().sum()) == 0.0 and float(full[1, 1:].abs().sum()) == 0.0
    q = ref.views_of(full[1, 0])
    assert torch.equal(q["x"], p.views["x"]) and int(q["status"][3]) == 7
    with pytest.raises(ValueError):
        sharding.PackedOutputs(5, 4, 6, 2, buf=torch.zeros(ref.words - 1, dtype=torch.float64))
