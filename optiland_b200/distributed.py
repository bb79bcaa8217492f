"""Multi-GPU plumbing for the trace path: one process per GPU (``torch.distributed``).

Rays are independent: each rank traces a contiguous range of the batch,
so there is NO data-path collective.  The only exchanges are
  * a broadcast of the packed surface table (a few KiB) from the rank that owns the live
    Optiland objects, and
  * optional all-reduces of analysis moments (spot centroid / RMS radius) -- a handful of
    scalars, latency-bound on NVLink/NVSwitch.
Works with the ``nccl`` backend (CUDA tensors) and with ``gloo`` (CPU tensors; used by the
world_size-2 tests that run without GPUs).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import table as T


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n rays owned by ``rank``; keeps field/pupil ordering so the
    per-rank record rows concatenate into the reference's (S, N) layout."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _comm_device(group=None) -> torch.device:
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def broadcast_table(table: T.SurfaceTable | None, src: int = 0, group=None) -> T.SurfaceTable:
    """Broadcast a packed ``SurfaceTable`` from ``src`` to every rank (others pass ``None``)."""
    dev = _comm_device(group)
    rank = dist.get_rank(group)
    if rank == src:
        surf, pool = table.pack()
        wl = np.ascontiguousarray(table.wavelengths, dtype=np.float64)
        raw = np.concatenate([surf.view(np.uint8).ravel(), pool.view(np.uint8).ravel(), wl.view(np.uint8).ravel()])
        sizes = torch.tensor([surf.size, pool.size, wl.size], dtype=torch.int64, device=dev)
    else:
        sizes = torch.zeros(3, dtype=torch.int64, device=dev)
    dist.broadcast(sizes, src=src, group=group)
    ns, npool, nwl = (int(v) for v in sizes.cpu())
    total = ns * T.OLB_SURFACE_DTYPE.itemsize + 8 * npool + 8 * nwl
    buf = torch.from_numpy(raw.copy()).to(dev) if rank == src else torch.empty(total, dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src=src, group=group)
    raw = buf.cpu().numpy()
    a = ns * T.OLB_SURFACE_DTYPE.itemsize
    b = a + 8 * npool
    return T.SurfaceTable.unpack(raw[:a].view(T.OLB_SURFACE_DTYPE).copy(), raw[a:b].view(np.float64).copy(),
                                 raw[b:].view(np.float64).copy())


def _allreduce_sum(v: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized():
        v = v.to(_comm_device(group))
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v.cpu()


def global_rms_spot_radius(x: torch.Tensor, y: torch.Tensor, intensity: torch.Tensor, group=None) -> float:
    """RMS spot radius about the centroid over ALL ranks' rays (the rms_spot_size operand,
    optiland/optimization/operand/ray.py:299-342, sharded).  Rays with i <= 0 or non-finite
    intercepts are masked as in optiland/analysis/spot_diagram/core.py:471-472.  Two 3-/1-scalar
    all-reduces (centroid first, then the centred second moment: the one-pass form loses 8
    digits on an off-axis field); both are latency-bound."""
    m = (intensity > 0) & torch.isfinite(x) & torch.isfinite(y)
    zero = torch.zeros((), dtype=torch.float64, device=x.device)
    xd = torch.where(m, x.double(), zero)
    yd = torch.where(m, y.double(), zero)
    cnt, sx, sy = (float(v) for v in _allreduce_sum(torch.stack([m.sum().double(), xd.sum(), yd.sum()]), group))
    if cnt == 0:
        return float("nan")
    cx, cy = sx / cnt, sy / cnt
    d2 = torch.where(m, (x.double() - cx) ** 2 + (y.double() - cy) ** 2, zero).sum().reshape(1)
    return float(np.sqrt(float(_allreduce_sum(d2, group)[0]) / cnt))


def sharded_rms_spot_loss_and_grad(trace_fn, params: torch.Tensor, group=None):
    """One step of the sharded autograd configuration (config 3 over several GPUs): every rank traces
    ITS shard of the rays differentiably -- ``trace_fn(params) -> (x, y)`` image-surface intercepts that are autograd
    outputs of ``params`` (``optiland_b200.autograd.trace_differentiable`` on the rank's rays) -- and gets back the
    GLOBAL loss = RMS spot radius about the global centroid (optimization/operand/ray.py:299-342) and the GLOBAL
    dLoss/dparams.

    Exchange: two scalar all-reduces for the loss (n, sum x, sum y; then the centred second moment) and one all-reduce(sum) of
    the (S, GP_COUNT) parameter-gradient block -- a few hundred doubles, latency-bound on NVLink.  No per-ray data
    crosses GPUs.  The chain rule through the global centroid needs no extra exchange: with V = mean |p - c|^2 and
    c = mean p, dV/dp_i = 2 (p_i - c) / n (the dependence through c cancels), so each rank back-propagates
    (p_i - c) / (n L) through its own shard."""
    x, y = trace_fn(params)
    xd, yd = x.detach().double(), y.detach().double()
    m = torch.isfinite(xd) & torch.isfinite(yd)
    zero = torch.zeros((), dtype=torch.float64, device=xd.device)
    sums = torch.stack([m.sum().double(), torch.where(m, xd, zero).sum(), torch.where(m, yd, zero).sum()])
    n, sx, sy = (float(v) for v in _allreduce_sum(sums, group))
    cx, cy = sx / n, sy / n
    # centred second moment in a second pass (the one-pass form sum(x^2 + y^2)/n - c^2 loses 6 digits on an off-axis
    # spot: 0.57 mm^2 against a variance of 5e-7 mm^2)
    d2 = torch.where(m, (xd - cx) ** 2 + (yd - cy) ** 2, zero).sum().reshape(1)
    var = float(_allreduce_sum(d2, group)[0]) / n
    loss = var ** 0.5
    mask = m.to(x.dtype)
    surrogate = (mask * ((x - cx) ** 2 + (y - cy) ** 2)).sum() / (2.0 * n * loss)
    (g,) = torch.autograd.grad(surrogate, params, allow_unused=True)
    g = torch.zeros_like(params) if g is None else g
    if dist.is_available() and dist.is_initialized():
        gd = g.to(_comm_device(group)).contiguous()
        dist.all_reduce(gd, op=dist.ReduceOp.SUM, group=group)
        g = gd.to(params.device)
    return loss, g


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> int | None:
    """NUMA node the GPU's PCIe root sits on (sysfs), or None when the platform does not say."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process (one process per GPU) to the CPUs of its GPU's NUMA node BEFORE it allocates pinned host
    buffers: first-touch then places the staging memory on the socket whose PCIe root complex the GPU hangs off, so
    the H2D / D2H legs of the host-buffer entry points (olb_trace_host_*) do not cross the inter-socket link and the
    8 ranks of one node do not all draw on one socket's memory controllers.  torchrun does not do this (it only sets
    OMP_NUM_THREADS=1).  Returns what was done, for the bench line."""
    import os

    node = gpu_numa_node(device_index)
    info = {"numa_node": node, "cpus": None}
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except (OSError, ValueError):
        pass
    return info
