"""Host-buffer path (olb_trace_host_f32) timing vs chunk size; plus raw pinned H2D / D2H bandwidth."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WAVELENGTH, load_case  # noqa: E402
from optiland_b200.launch import launch_infinite_angle  # noqa: E402
from optiland_b200.trace import DeviceTable, RealRays, trace_host  # noqa: E402

n = 10_000_000
c, sc = load_case()
dev = torch.device("cuda:0")
dtab = DeviceTable(c.table, dev)
S = c.table.num_surfaces
g = torch.Generator(device=dev).manual_seed(0)
r = torch.rand(n, generator=g, device=dev, dtype=torch.float64).sqrt()
th = 6.283185307179586 * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
base = RealRays(*launch_infinite_angle(r * torch.cos(th), r * torch.sin(th), sc), 1.0, WAVELENGTH, dtype=torch.float32, device=dev)
h_in = {k: getattr(base, k).cpu().pin_memory() for k in ("x", "y", "z", "L", "M", "N", "i")}
h_out = {k: torch.empty(n, dtype=torch.float32).pin_memory() for k in ("x", "y", "z", "L", "M", "N", "i", "opd")}
rec = torch.empty((8, S, n), dtype=torch.float32, device=dev)
# raw PCIe
d = torch.empty(n * 8, dtype=torch.float32, device=dev)
hh = torch.empty(n * 8, dtype=torch.float32).pin_memory()
for name, fn in (("h2d", lambda: d.copy_(hh, non_blocking=True)), ("d2h", lambda: hh.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print(json.dumps({name + "_GBps": round(5 * n * 32 / (time.perf_counter() - t0) / 1e9, 1)}))
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    scratch = None
    for _ in range(2):
        scratch = trace_host(dtab, h_in, h_out, n, torch.float32, chunk=chunk, scratch=scratch, rec=rec)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 8
    for _ in range(K):
        trace_host(dtab, h_in, h_out, n, torch.float32, chunk=chunk, scratch=scratch, rec=rec)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / K * 1e3
    print(json.dumps({"chunk": chunk, "ms": round(ms, 3), "Grs_per_s": round(n * 12 / ms / 1e6, 2)}))
