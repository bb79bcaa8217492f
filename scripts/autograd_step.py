"""A few forward+backward steps of the config-3 autograd workload (for ncu launch lists / timing)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200 import autograd as AG  # noqa: E402
from optiland_b200.trace import RealRays  # noqa: E402
from scripts.bench_configs import resample  # noqa: E402
from tests._util import Case  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
dtype = torch.float32 if (len(sys.argv) < 3 or sys.argv[2] == "f32") else torch.float64
c = Case("telephoto_c3_tol1e-6")
base = resample(c, n, dtype)
params = AG.table_to_params(c.table).requires_grad_(True)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for it in range(6):
    e = [ev() for _ in range(4)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e[0].record()
    rr = RealRays.__new__(RealRays)
    rr.__dict__.update(base.__dict__)
    rec = AG.trace_differentiable(c.table, params, rr, rows=(-1,))
    e[1].record()
    x, y = rec["x"], rec["y"]
    loss = torch.sqrt(torch.mean((x - x.mean()) ** 2 + (y - y.mean()) ** 2))
    e[2].record()
    params.grad = None
    loss.backward()
    e[3].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"step {it}: fwd {e[0].elapsed_time(e[1]):.3f} ms, loss {e[1].elapsed_time(e[2]):.3f} ms, "
          f"bwd {e[2].elapsed_time(e[3]):.3f} ms, wall {wall:.3f} ms")
