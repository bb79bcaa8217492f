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
