"""What bounds the end-to-end (host-buffer) path when all GPUs of the box copy at once: per-GPU pinned D2H / H2D
bandwidth with 1 .. N GPUs active concurrently (one process per GPU, each bound to its GPU's NUMA node as bench.py does),
plus the box's `nvidia-smi topo -m`.   torchrun --nproc-per-node N scripts/pcie_concurrent.py"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200.distributed import bind_to_gpu_numa  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 320_000_000
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    devb = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
        for name, (src, dst) in (("d2h", (devb, host)), ("h2d", (host, devb))):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            gbps = 0.0
            if rank < active:
                dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                gbps = 5 * nbytes / (time.perf_counter() - t0) / 1e9
            t = torch.tensor([gbps], device="cuda")
            if world > 1:
                g = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(g, t)
                vals = [round(float(v), 1) for v in g][:active]
            else:
                vals = [round(gbps, 1)]
            out[f"{name}_GBps_per_gpu_{active}_active"] = vals
    if rank == 0:
        out["numa_rank0"] = numa
        try:
            out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        except Exception as e:
            out["topo"] = str(e)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
