"""N>1 host logic with world_size 2 over gloo (no GPU): contiguous ray sharding, the packed
surface-table broadcast and the moment all-reduce.  The per-rank trace itself is the CUDA
kernel on a GPU box; here each rank evaluates its shard with the oracle so that the sharded
result can be compared with the single-process one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optiland_b200.distributed import broadcast_table, global_rms_spot_radius, shard_range
from tests._util import Case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import trace_oracle as O

        c = Case("hubble_c4")
        table = broadcast_table(c.table if rank == 0 else None, src=0)
        lo, hi = shard_range(c.n, rank, world)
        sub = {k: v[lo:hi] for k, v in c.rays.items()}
        out, rec, _ = O.trace(table, sub)
        rms = global_rms_spot_radius(torch.from_numpy(rec["x"][-1]), torch.from_numpy(rec["y"][-1]),
                                     torch.from_numpy(rec["intensity"][-1]))
        q.put((rank, lo, hi, rec["y"][-1], rms, table.num_surfaces))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10, 1001):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gloo_table_broadcast_and_moments():
    from oracle import trace_oracle as O

    c = Case("hubble_c4")
    _, rec, _ = O.trace(c.table, c.rays)
    x, y, i = rec["x"][-1], rec["y"][-1], rec["intensity"][-1]
    m = (i > 0) & np.isfinite(x) & np.isfinite(y)
    ref_rms = np.sqrt(np.mean((x[m] - x[m].mean()) ** 2 + (y[m] - y[m].mean()) ** 2))

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    ys = np.concatenate([g[3] for g in got])
    np.testing.assert_array_equal(ys, y)  # shards concatenate into the single-process record row
    for g in got:
        assert g[5] == c.table.num_surfaces
        assert g[4] == pytest.approx(ref_rms, rel=1e-12)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from optiland_b200.distributed import sharded_rms_spot_loss_and_grad

        n = 1001
        lo, hi = shard_range(n, rank, world)
        px, py, params0 = _toy_problem(n)
        params = params0.clone().requires_grad_(True)
        loss, g = sharded_rms_spot_loss_and_grad(lambda p: _toy_trace(p, px[lo:hi], py[lo:hi]), params)
        q.put((rank, loss, g.numpy()))
    finally:
        dist.destroy_process_group()


def _toy_problem(n):
    g = torch.Generator().manual_seed(5)
    px, py = torch.rand(n, generator=g, dtype=torch.float64) - 0.5, torch.rand(n, generator=g, dtype=torch.float64) - 0.5
    params = torch.tensor([[0.3, -0.2, 1.5], [0.7, 0.1, -0.4]], dtype=torch.float64)
    return px, py, params


def _toy_trace(p, px, py):
    """A smooth stand-in for the per-rank differentiable trace (the real one is the CUDA forward + adjoint kernels)."""
    x = p[0, 0] * px + p[0, 1] * py ** 2 + p[1, 2] * torch.sin(p[0, 2] * px) + 12.0
    y = p[1, 0] * py + p[1, 1] * px * py + 0.1 * p[0, 2] * px ** 3 - 7.0
    return x, y


@pytest.mark.timeout(120)
def test_two_rank_sharded_gradient_step_equals_single_process():
    """The all-reduce of the parameter gradients (sharded config-3 step): 2 ranks over gloo == 1 process."""
    px, py, params0 = _toy_problem(1001)
    params = params0.clone().requires_grad_(True)
    x, y = _toy_trace(params, px, py)
    loss = torch.sqrt(torch.mean((x - x.mean()) ** 2 + (y - y.mean()) ** 2))
    loss.backward()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=100) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for _, l, g in got:
        assert l == pytest.approx(float(loss), rel=1e-12)
        np.testing.assert_allclose(g, params.grad.numpy(), rtol=1e-10, atol=1e-14)
