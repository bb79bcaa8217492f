"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL).  Checks, on hardware:
  1. rays sharded by `shard_range` over the ranks, table broadcast over NCCL: the shards' records, gathered and
     concatenated, are BIT-IDENTICAL to the records of a single-GPU trace of the whole batch (fp32 and fp64; C2
     Double-Gauss and the polarized C5 system through the fused launch);
  2. `global_rms_spot_radius` over NCCL == the single-GPU value;
  3. the sharded config-3 gradient step (`sharded_rms_spot_loss_and_grad`: all-reduce of the parameter gradients)
     == the single-GPU forward + backward.
Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from optiland_b200 import autograd as AG
    from optiland_b200.distributed import (broadcast_table, global_rms_spot_radius, shard_range,
                                           sharded_rms_spot_loss_and_grad)
    from optiland_b200.launch import launch_from_affine, pupil_affine, pupil_affine_fields
    from optiland_b200.trace import DeviceTable, RealRays, trace_pupil_device
    from tests._util import Case

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    def gather_rows(t):
        """(rows, n_local) tensors of every rank -> (rows, N) on rank 0 (variable shard sizes)."""
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[-1]], dtype=torch.int64, device=dev))
        mx = int(max(int(s) for s in sizes))
        pad = torch.zeros(t.shape[:-1] + (mx,), dtype=t.dtype, device=dev)
        pad[..., :t.shape[-1]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        return torch.cat([b[..., :int(s)] for b, s in zip(bufs, sizes)], dim=-1)

    # ---- 1 + 2: sharded forward == single GPU, bit for bit ---------------------------------------------
    N = 1_000_003
    j = torch.arange(N, device=dev, dtype=torch.float64)
    r = torch.sqrt((j + 0.5) / N)
    th = j * (np.pi * (3.0 - np.sqrt(5.0)))
    for case, dtype in (("dgauss_c2", torch.float32), ("dgauss_c2", torch.float64), ("hubble_c4", torch.float64)):
        c = Case(case)
        sc = {k[9:]: float(c.z[k]) for k in c.z.files if k.startswith("x_launch_")}
        table = broadcast_table(c.table if rank == 0 else None, src=0)
        dtab = DeviceTable(table, dev)
        Px, Py = (r * torch.cos(th)).to(dtype), (r * torch.sin(th)).to(dtype)
        lo, hi = shard_range(N, rank, world)
        S = table.num_surfaces
        _, rec = trace_pupil_device(dtab, Px[lo:hi].contiguous(), Py[lo:hi].contiguous(), pupil_affine(sc), 0, S)
        rms = global_rms_spot_radius(rec["x"][-1], rec["y"][-1], rec["intensity"][-1])
        full = {k: gather_rows(rec[k].contiguous()) for k in ("x", "y", "opd", "L", "intensity")}
        if rank == 0:
            _, one = trace_pupil_device(dtab, Px, Py, pupil_affine(sc), 0, S)
            same = all(torch.equal(full[k].view(torch.int64 if dtype == torch.float64 else torch.int32),
                                   one[k].contiguous().view(torch.int64 if dtype == torch.float64 else torch.int32)) for k in full)
            x, y, i = (one[k][-1].double() for k in ("x", "y", "intensity"))
            m = (i > 0) & torch.isfinite(x) & torch.isfinite(y)
            rms1 = float(torch.sqrt(((x[m] - x[m].mean()) ** 2 + (y[m] - y[m].mean()) ** 2).mean()))
            out[f"{case}_{str(dtype)[6:]}"] = {"bit_identical": bool(same), "rms_nccl": rms, "rms_single": rms1,
                                              "rms_rel_err": abs(rms - rms1) / rms1}
    # polarized C5 call shape, per-ray fields + wavelengths
    c = Case("generic_polarized_c5")
    sc = {k[9:]: float(c.z[k]) for k in c.z.files if k.startswith("x_launch_")}
    table = broadcast_table(c.table if rank == 0 else None, src=0)
    dtab = DeviceTable(table, dev)
    n5 = 300_000
    g = torch.Generator(device=dev).manual_seed(7)
    idx = torch.randint(0, c.n, (n5,), device=dev, generator=g)
    arr = {k: torch.from_numpy(np.ascontiguousarray(c.extra(k))).to(dev)[idx] for k in ("Px", "Py", "Hx", "Hy")}
    w = torch.from_numpy(c.rays["w"]).to(dev)[idx]
    lo, hi = shard_range(n5, rank, world)
    sl = {k: v[lo:hi].contiguous() for k, v in arr.items()}
    rays, rec = trace_pupil_device(dtab, sl["Px"], sl["Py"], pupil_affine_fields(sc, sl["Hx"], sl["Hy"]), 0, table.num_surfaces,
                                   wavelength=w[lo:hi].contiguous(), polarization=None)
    full_i = gather_rows(rays.i.reshape(1, -1).contiguous())
    full_p = gather_rows(torch.view_as_real(rays.p).reshape(-1, 18).T.contiguous())
    if rank == 0:
        r1, _ = trace_pupil_device(dtab, arr["Px"], arr["Py"], pupil_affine_fields(sc, arr["Hx"], arr["Hy"]), 0,
                                   table.num_surfaces, wavelength=w, polarization=None)
        out["c5_polarized_f64"] = {"intensity_bit_identical": bool(torch.equal(full_i[0], r1.i)),
                                   "P_bit_identical": bool(torch.equal(full_p, torch.view_as_real(r1.p).reshape(-1, 18).T))}

    # ---- 3: sharded gradient step ------------------------------------------------------------------------
    c = Case("telephoto_c3_tol1e-10")
    table = broadcast_table(c.table if rank == 0 else None, src=0)
    n3 = 400_000
    g = torch.Generator(device=dev).manual_seed(3)
    idx = torch.randint(0, c.n, (n3,), device=dev, generator=g)
    rr = {k: torch.from_numpy(v).to(dev)[idx] for k, v in c.rays.items()}
    lo, hi = shard_range(n3, rank, world)

    def mk(a, b):
        return RealRays(*[rr[k][a:b] for k in ("x", "y", "z", "L", "M", "N", "i", "w")], dtype=torch.float64, device=dev)

    params = AG.table_to_params(table).to(dev).requires_grad_(True)

    def trace_fn(p, a=lo, b=hi):
        rec = AG.trace_differentiable(table, p, mk(a, b), rows=(-1,))
        return rec["x"], rec["y"]

    loss, grad = sharded_rms_spot_loss_and_grad(trace_fn, params)
    if rank == 0:
        p1 = AG.table_to_params(table).to(dev).requires_grad_(True)
        x, y = trace_fn(p1, 0, n3)
        l1 = torch.sqrt(torch.mean((x - x.mean()) ** 2 + (y - y.mean()) ** 2))
        l1.backward()
        scale = float(p1.grad.abs().max())
        out["c3_sharded_gradient"] = {"loss_rel_err": abs(loss - float(l1)) / float(l1),
                                      "grad_max_abs_err_over_scale": float((grad - p1.grad).abs().max()) / scale,
                                      "grad_entries": int((p1.grad != 0).sum())}
        print("MULTI_GPU_RESULT " + json.dumps(out), flush=True)
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
