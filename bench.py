#!/usr/bin/env python
"""bench.py -- ray-surface intersections / s on BASELINE.json's headline configuration.

Workload (config 2): the Double-Gauss sample (13 surfaces, 12 traced: 8 spherical + 4 plane),
10 M rays per GPU, single field / wavelength, full per-surface records (the reference's
semantics: every surface stores x,y,z,L,M,N,intensity,opd), fp32.

One "step" = one pass of the hot path (`SurfaceGroup.trace`) over the whole batch = ONE launch
of the persistent sm_100a kernel.  `value` is measured with the launch rays resident in HBM;
`e2e` goes through the C ABI's host-buffer entry point with pinned host arrays (H2D + kernel +
D2H inside the timed region).  `--impl reference` times the CPU oracle port (the reference is
pure Python and does not travel to the GPU box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ray-surface intersections/sec"
UNIT = "ray-surfaces/s"
WORKLOAD = "double_gauss_13surf_10Mrays_full_records"
N_RAYS = 10_000_000
WAVELENGTH = 0.5876


def load_case():
    from tests._util import Case

    c = Case("dgauss_c2")
    sc = {k[9:]: float(c.z[k]) for k in c.z.files if k.startswith("x_launch_")}
    return c, sc


def pupil_numpy(n, seed):
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    return r * np.cos(th), r * np.sin(th)


# --------------------------------------------------------------------------------------
# CPU leg: the oracle port (NumPy fp64), chunked over a thread pool
# --------------------------------------------------------------------------------------

def cpu_trace_throughput(table, sc, n_rays, threads, chunk=100_000, seed=123):
    """Trace `n_rays` launch rays through the table with the NumPy oracle; return
    (ray-surfaces/s, seconds).  NumPy ufuncs release the GIL, so chunks run on `threads` cores."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import trace_oracle as O
    from optiland_b200.launch import launch_infinite_angle

    Px, Py = pupil_numpy(n_rays, seed)
    x0, y0, z0, L, M, N = launch_infinite_angle(Px, Py, sc)
    rays = dict(x=x0, y=y0, z=z0, L=L, M=M, N=N, i=np.ones(n_rays), w=np.full(n_rays, WAVELENGTH))
    n_traced = table.num_surfaces - 1

    def work(lo):
        sub = {k: v[lo:lo + chunk] for k, v in rays.items()}
        out, rec, _ = O.trace(table, sub)
        return float(out["x"][0])

    starts = list(range(0, n_rays, chunk))
    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, starts))
    else:
        for lo in starts:
            work(lo)
    dt = time.perf_counter() - t0
    return n_rays * n_traced / dt, dt


def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c, sc = load_case()
    threads = min(os.cpu_count() or 1, 32)
    sample = 200_000
    for _ in range(max(args.warmup, 0) and 1):
        cpu_trace_throughput(c.table, sc, 100_000, threads)
    t_total, work = 0.0, 0
    for _ in range(args.steps):
        v, dt = cpu_trace_throughput(c.table, sc, sample, threads)
        t_total += dt
        work += sample * (c.table.num_surfaces - 1)
    value = work / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": sample, "surfaces": c.table.num_surfaces,
                   "note": "bounded sample of the same workload per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} rays x {c.table.num_surfaces - 1} surfaces per step, NumPy fp64 oracle "
                                   f"port, {threads} threads over 100k-ray chunks"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# clocks sampling
# --------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_lo=None, t_hi=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.lines:
            if t_lo is not None and not (t_lo <= t <= t_hi + 0.05):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    from optiland_b200 import _lib
    from optiland_b200 import table as T
    from optiland_b200.launch import launch_infinite_angle
    from optiland_b200.trace import DeviceTable, RealRays, trace_device, trace_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- surface table: rank 0 packs, NCCL-broadcasts the (tiny) packed table ----------
    c, sc = load_case()
    if world > 1:
        from optiland_b200.distributed import broadcast_table

        table = broadcast_table(c.table if rank == 0 else None, src=0)
    else:
        table = c.table
    dtab = DeviceTable(table, dev)
    lib = dtab.lib
    S = table.num_surfaces
    n_traced = S - 1
    n = args.rays
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    es = 4 if args.dtype == "f32" else 8

    # ---- launch rays (each rank its own pupil sample: weak scaling, no exchange) -------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    r = torch.rand(n, generator=g, device=dev, dtype=torch.float64).sqrt()
    th = 2 * np.pi * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    x0, y0, z0, L, M, N = launch_infinite_angle(r * torch.cos(th), r * torch.sin(th), sc)
    del r, th
    base = RealRays(x0, y0, z0, L, M, N, 1.0, WAVELENGTH, dtype=dtype, device=dev)
    del x0, y0, z0, L, M, N
    torch.cuda.empty_cache()

    def fresh():
        """A RealRays view over the SAME launch arrays (the trace with records does not modify them)."""
        rr = RealRays.__new__(RealRays)
        rr.__dict__.update(base.__dict__)
        return rr

    def step():
        rr = fresh()
        rec = trace_device(dtab, rr, 0, S, record=True)
        return rr, rec

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        rr, rec = step()
    del rr, rec
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = lib.olb_launch_count()
    barrier()
    t_lo = time.perf_counter()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        rr, rec = step()
        ev[k][1].record()
    t_end.record()
    barrier()
    t_hi = time.perf_counter()
    launches = lib.olb_launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    # nvidia-smi samples every 20 ms: a short timed region (small --steps) may hold no sample at all; then the
    # SAME step keeps running, untimed, until the load window is >= 0.1 s, and the clocks line says so
    clock_extra = 0
    t_clock_hi = t_hi
    while t_clock_hi - t_lo < 0.1:
        rr, rec = step()
        torch.cuda.synchronize()
        clock_extra += 1
        t_clock_hi = time.perf_counter()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    # sanity of the last step's result (not timed): image-surface centroid must be finite
    chk = float(rec["x"][-1].mean().item())
    assert np.isfinite(chk)
    del rr, rec

    # ---- e2e: pinned host arrays -> C ABI host entry point -> pinned host result --------
    # (a) the Optic.trace-shaped call: the per-ray inputs are the pupil samples (Px, Py); the launch
    #     state (paraxial aiming) is generated on the device (olb_trace_host_pupil_*);
    # (b) the SurfaceGroup.trace-shaped call: the full launch state arrays cross PCIe (olb_trace_host_*).
    from optiland_b200.launch import pupil_affine_infinite_angle

    aff = pupil_affine_infinite_angle(sc)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    r = torch.rand(n, generator=g, device=dev, dtype=torch.float64).sqrt()
    th = 2 * np.pi * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    h_pupil = {"Px": (r * torch.cos(th)).to(dtype).cpu().pin_memory(), "Py": (r * torch.sin(th)).to(dtype).cpu().pin_memory()}
    del r, th
    keys_in = ("x", "y", "z", "L", "M", "N", "i")
    h_in = {k: getattr(base, k).cpu().pin_memory() for k in keys_in}
    h_out = {k: torch.empty(n, dtype=dtype).pin_memory() for k in ("x", "y", "z", "L", "M", "N", "i", "opd")}
    vec = 4 if es == 4 else 2
    stride = n if n % vec == 0 else (n + 63) // 64 * 64
    rec_buf = torch.empty((8, S, stride), dtype=dtype, device=dev)
    chunk = 1 << 20
    e2e_steps = max(3, min(args.steps, 10))

    def time_host(inputs, affine):
        scratch = None
        for _ in range(2):
            scratch = trace_host(dtab, inputs, h_out, n, dtype, chunk=chunk, scratch=scratch, rec=rec_buf, affine=affine)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            trace_host(dtab, inputs, h_out, n, dtype, chunk=chunk, scratch=scratch, rec=rec_buf, affine=affine)
            metric_val = float(h_out["x"][:1024].mean())  # touch the result on the host
        torch.cuda.synchronize(dev)
        assert np.isfinite(metric_val)
        return (time.perf_counter() - t0) / e2e_steps

    # (c) SpotDiagram-shaped call (f-1 + f-2): pupil samples in, RMS spot radius out -- the trace writes NO
    #     per-ray output; 64 bytes come back.  Chunked H2D on 3 streams overlapping the kernels.
    from optiland_b200.trace import moments_to_spot, trace_moments_device

    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    d_px = [torch.empty(chunk, dtype=dtype, device=dev) for _ in range(3)]
    d_py = [torch.empty(chunk, dtype=dtype, device=dev) for _ in range(3)]

    def spot_step():
        mom = torch.zeros(8, dtype=torch.float64, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        for ci, lo in enumerate(range(0, n, chunk)):
            m = min(chunk, n - lo)
            st = streams[ci % 3]
            with torch.cuda.stream(st):
                d_px[ci % 3][:m].copy_(h_pupil["Px"][lo:lo + m], non_blocking=True)
                d_py[ci % 3][:m].copy_(h_pupil["Py"][lo:lo + m], non_blocking=True)
                trace_moments_device(dtab, m, dtype, pupil=(d_px[ci % 3][:m], d_py[ci % 3][:m], aff), moments=mom)
        for st in streams:
            st.synchronize()
        return moments_to_spot(mom)["rms_centroid"]

    for _ in range(2):
        rms_val = spot_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rms_val = spot_step()
    torch.cuda.synchronize(dev)
    spot_s = (time.perf_counter() - t0) / e2e_steps
    assert np.isfinite(rms_val)
    del d_px, d_py

    e2e_s = time_host(h_pupil, aff)
    e2e_state_s = time_host(h_in, None)
    h2d = 2 * es * n
    h2d_state = len(keys_in) * es * n
    d2h = 8 * es * n
    del rec_buf

    # ---- max over ranks ----------------------------------------------------------------
    times = torch.tensor([total_ms, kern_ms, e2e_s * 1e3, e2e_state_s * 1e3, spot_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_ms, e2e_state_ms, spot_ms = (float(v) for v in times.cpu())

    if rank == 0:
        clocks = sampler.stop(t_lo, t_clock_hi)
        clocks["window"] = ("timed region" if clock_extra == 0 else
                            f"timed region + {clock_extra} further untimed steps of the same launch (the timed region "
                            f"of {total_ms:.1f} ms is too short for nvidia-smi's 20 ms sampling)")
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        n_loads = 8 if table.n_wl == 1 else 9  # x,y,z,L,M,N,i,opd (+w when several wavelengths)
        bytes_per_ray = es * (n_loads + 8 * S)  # full records, final state aliased to the last row
        achieved = bytes_per_ray * n / (kern_ms * 1e-3) / 1e9
        value = world * n * n_traced / (total_ms * 1e-3 / args.steps)
        cpu_threads = min(os.cpu_count() or 1, 32)
        cpu_n = 1_000_000
        cpu_val, cpu_dt = cpu_trace_throughput(table, sc, cpu_n, cpu_threads)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.dtype)
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "system": "DoubleGauss (optiland.samples), 13 surfaces / 12 traced",
                       "rays_per_gpu": n, "records": "full (8 arrays x 13 surfaces)", "wavelengths": 1,
                       "l2_policy": f"inputs+records {bytes_per_ray * n / 1e9:.2f} GB per step >> 126 MB L2",
                       "parallelism": f"rays sharded over {world} GPU(s), table broadcast, no exchange"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                         "algorithmic_bytes_per_ray": bytes_per_ray, "kernel_ms": kern_ms,
                         "kernel": "olb::trace_kernel<%s,%d,0>" % (("float", 4) if es == 4 else ("double", 1))},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                             "sample": f"{cpu_n} rays x {n_traced} surfaces in {cpu_dt:.1f} s, NumPy fp64 oracle port, "
                                       f"{cpu_threads} threads over 100k-ray chunks (os.cpu_count={os.cpu_count()})"},
            "e2e": {"value": world * n * n_traced / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "what": "Optic.trace-shaped call through olb_trace_host_pupil_*: pinned host pupil samples (Px, Py) "
                            "-> H2D -> launch state generated in-kernel -> trace (records stay in HBM) -> D2H of the "
                            "final ray state (x,y,z,L,M,N,i,opd); 1 Mi-ray chunks on 3 streams"},
            "e2e_launch_arrays": {"value": world * n * n_traced / (e2e_state_ms * 1e-3), "unit": UNIT,
                                  "h2d_bytes_per_step": h2d_state, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_state_ms,
                                  "what": "SurfaceGroup.trace-shaped call through olb_trace_host_*: the 7 launch-state "
                                          "arrays cross PCIe instead of the 2 pupil arrays"},
            "e2e_spot_rms": {"value": world * n * n_traced / (spot_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                             "d2h_bytes_per_step": 64, "ms_per_step": spot_ms, "rms_spot_radius_mm": rms_val,
                             "what": "SpotDiagram-shaped call (next rows f-1 + f-2): pinned host pupil samples -> H2D -> "
                                     "launch generation + trace + spot moments fused in one kernel, NO per-ray output "
                                     "(no records) -> D2H of 8 doubles; not the headline (it skips the records)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
