"""
world_size-2 coverage of the multi-GPU host logic on the gloo backend (no GPU):
row / grid-point sharding bounds, the all-gather that reassembles sharded
predictions, the score-table reduction of the SplineCV job deal, and -- with the
CPU oracle standing in for the device kernels -- that summing per-rank partial
normal equations reproduces the single-rank system (the property the NCCL
all-reduce of the Gram matrix relies on).
"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as td
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    import verde_oracle as vo

    from verde_b200 import dist

    assert dist.world() == (rank, world) and dist.sharding_active()
    # 1. sharded "predict": each rank computes its contiguous slice, all-gather reassembles
    rng = np.random.RandomState(5)
    e, n = rng.uniform(0, 100, 1001), rng.uniform(0, 100, 1001)
    fe, fn, f = rng.uniform(0, 100, 37), rng.uniform(0, 100, 37), rng.normal(size=37)
    a, b = dist.shard_bounds(e.size)
    local = vo.predict(e[a:b], n[a:b], fe, fn, 0.0, f)
    full = dist.allgather_concat(local, e.size)
    np.testing.assert_array_equal(full, vo.predict(e, n, fe, fn, 0.0, f))
    # 2. row-sharded normal equations: partial sums + all-reduce == the full system
    d = rng.normal(size=e.size)
    w = rng.uniform(0.5, 2, e.size)
    jac = vo.jacobian(e[a:b], n[a:b], fe, fn, 0.0)
    parts = vo.normal_equations(jac, d[a:b], w[a:b])
    packed = np.concatenate([p.ravel() for p in parts])
    summed = dist.allreduce_sum_numpy(packed)
    ref = np.concatenate([p.ravel() for p in vo.normal_equations(vo.jacobian(e, n, fe, fn, 0.0), d, w)])
    np.testing.assert_allclose(summed, ref, rtol=1e-13)
    # 3. SplineCV job deal: every job scored by exactly one rank, table summed
    jobs = list(range(7))
    table = np.zeros(len(jobs))
    for j in dist.round_robin(len(jobs)):
        table[j] = 10.0 + j
    table = dist.allreduce_sum_numpy(table)
    np.testing.assert_array_equal(table, 10.0 + np.arange(7))
    # 4. panel-cyclic Cholesky: the schedule + broadcast the GPU driver (GramSystem.solve_distributed)
    #    uses, with numpy standing in for the panel / update kernels on a dense SPD matrix
    import torch

    nb, blk, pw = 11, 8, 3  # 11 block columns of 8, outer panels of 3 block columns
    m = nb * blk
    x = np.random.RandomState(9).normal(size=(3 * m, m))
    a_full = x.T @ x + 1e-3 * np.eye(m)
    work = torch.from_numpy(a_full.copy())  # every rank starts from the all-reduced matrix
    mat = work.numpy()
    schedule = dist.panel_cyclic_schedule(nb, pw, world)
    assert [o for _, _, o in schedule] == [0, 1, 0, 1] and schedule[-1][:2] == (9, 2)
    for p, (k0, width, owner) in enumerate(schedule):
        c0, c1 = k0 * blk, (k0 + width) * blk
        if rank == owner:  # "chol_panel": factor the diagonal block, solve the rows below
            mat[c0:c1, c0:c1] = np.linalg.cholesky(mat[c0:c1, c0:c1])
            mat[c1:, c0:c1] = np.linalg.solve(mat[c0:c1, c0:c1], mat[c1:, c0:c1].T).T
        panel = work[:, c0:c1].contiguous()
        dist.broadcast_(panel, owner)
        work[:, c0:c1] = panel
        for q0, qwidth, qowner in schedule[p + 1:]:
            if qowner == rank:  # "chol_update" of the block columns this rank owns
                j0, j1 = q0 * blk, (q0 + qwidth) * blk
                mat[j0:, j0:j1] -= mat[j0:, c0:c1] @ mat[j0:j1, c0:c1].T
    np.testing.assert_allclose(np.tril(mat), np.linalg.cholesky(a_full), rtol=0, atol=1e-10)
    # 4b. the LOOK-AHEAD order of GramSystem.solve_distributed (the owner of panel p+1 updates and factors it first,
    #     as soon as panel p has arrived; everything else it owns is updated afterwards through dist.owned_ranges)
    work = torch.from_numpy(a_full.copy())
    mat = work.numpy()

    def factor(c0, c1):
        mat[c0:c1, c0:c1] = np.linalg.cholesky(mat[c0:c1, c0:c1])
        mat[c1:, c0:c1] = np.linalg.solve(mat[c0:c1, c0:c1], mat[c1:, c0:c1].T).T

    def update(c0, c1, j0, j1):
        mat[j0:, j0:j1] -= mat[j0:, c0:c1] @ mat[j0:j1, c0:c1].T

    if rank == schedule[0][2]:
        factor(0, schedule[0][1] * blk)
    for p, (k0, width, owner) in enumerate(schedule):
        c0, c1 = k0 * blk, (k0 + width) * blk
        panel = work[:, c0:c1].contiguous()
        dist.broadcast_(panel, owner)
        work[:, c0:c1] = panel
        later = schedule[p + 1:]
        if later and later[0][2] == rank:
            q0, qwidth, _ = later[0]
            update(c0, c1, q0 * blk, (q0 + qwidth) * blk)
            factor(q0 * blk, (q0 + qwidth) * blk)
            later = later[1:]
        for lo, hi in dist.owned_ranges(later, rank):
            update(c0, c1, lo * blk, hi * blk)
    np.testing.assert_allclose(np.tril(mat), np.linalg.cholesky(a_full), rtol=0, atol=1e-10)
    assert dist.owned_ranges(schedule, 0) == [(0, 3), (6, 9)] and dist.owned_ranges(schedule[1:], 1) == [(3, 6), (9, 11)]
    assert dist.panel_tiles(2) == 8 and dist.panel_tiles(4) == 4 and dist.panel_tiles(8) == 4
    assert dist.use_distributed_cholesky(391, 8) and not dist.use_distributed_cholesky(24, 8)
    dist.set_cholesky("replicated")
    assert not dist.use_distributed_cholesky(391, 8)
    dist.set_cholesky("auto")
    # a status code seen by one rank only (failed pivot on the panel owner) becomes collective
    assert dist.agree_on_status(7 if rank == 1 else 0) == (0, 7)
    assert dist.agree_on_status(-2 if rank == 0 else 0) == (-2, 0)
    # 5. turning sharding off makes the helpers local again
    dist.set_sharding(False)
    assert not dist.sharding_active()
    dist.set_sharding("auto")
    open(os.path.join(out_dir, "ok_%d" % rank), "w").write("ok")
    td.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok_0", "ok_1"]


def test_shard_bounds_cover_everything():
    from verde_b200 import dist

    for total in (0, 1, 7, 50_000, 4_000_001):
        for size in (1, 2, 3, 8):
            spans = [dist.shard_bounds(total, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
                assert b0 == a1 and b0 - a0 >= b1 - a1 >= 0
    assert dist.round_robin(20, 3, 8) == [3, 11, 19]
    assert dist.world() == (0, 1) and not dist.sharding_active()
    with pytest.raises(ValueError):
        dist.set_sharding("maybe")
    with pytest.raises(ValueError):
        dist.set_cholesky("sometimes")
    assert len(dist.panel_cyclic_schedule(392, 4, 8)) == 98 and dist.panel_tiles(1) == 8
    for mode in ("serial", "lookahead", "peer"):
        dist.set_panel_pipeline(mode)
        assert dist.panel_pipeline() == mode
    with pytest.raises(ValueError):
        dist.set_panel_pipeline("eager")
    dist.set_merged_updates(False)
    assert not dist.merged_updates()
    dist.set_merged_updates(True)
    sched = dist.panel_cyclic_schedule(391, 8, 8)
    assert len(sched) == 49 and sched[0] == (0, 8, 0) and sched[-1] == (384, 7, 0) and sched[9] == (72, 8, 1)
    assert sum(w for _, w, _ in sched) == 391
    assert not dist.use_distributed_cholesky(391, 8, 1)
