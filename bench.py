#!/usr/bin/env python
"""bench.py -- ray-surface intersections / s on BASELINE.json's headline configuration.

Workload (config 2): the Double-Gauss sample (13 surfaces, 12 traced: 8 spherical + 4 plane),
10 M rays per GPU, single field / wavelength, full per-surface records (the reference's
semantics: every surface stores x,y,z,L,M,N,intensity,opd), fp32.

One "step" = one pass of the hot path (`SurfaceGroup.trace`) over the whole batch = ONE launch
of the sm_100a kernel.  `value` is measured with the launch rays resident in HBM; `e2e` goes
through the C ABI's host-buffer entry point with pinned host arrays (H2D + kernel + D2H inside
the timed region); `e2e_optic_trace` is the same through the UNMODIFIED reference's own
`Optic.trace` with the plugin installed (live Optiland objects; needs oracle/_ref, see
scripts/make_ref.sh).  Before the line is printed a strided sample of the timed batch is traced
through the CPU oracle and compared (`parity`); the line is refused if that fails.

`--impl reference` times the reference's own CPU implementation of the path: the stock NumPy
backend, `optic.surfaces.trace(rays)` on pre-generated launch arrays, one
process per host core (the reference's Python layer holds the GIL, so threads do not scale);
the NumPy oracle port is the labelled fallback when the reference is not on the box.

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
REF_CHUNK = 100_000          # rays per work item of the CPU arms


def load_case(name="dgauss_c2"):
    from tests._util import Case

    c = Case(name)
    sc = {k[9:]: float(c.z[k]) for k in c.z.files if k.startswith("x_launch_")}
    return c, sc


def workload_config(n_surfaces, n_rays, world):
    """The `config` object: identical in both arms (the reference arm traces a bounded SAMPLE of it per step and
    says so in `cpu_baseline.sample`)."""
    return {"workload": WORKLOAD, "system": f"DoubleGauss (optiland.samples), {n_surfaces} surfaces / {n_surfaces - 1} traced",
            "rays_per_gpu": n_rays, "records": f"full (8 arrays x {n_surfaces} surfaces)", "wavelengths": 1,
            "field": "Hx = Hy = 0", "pupil": "uniform random in the unit disk"}


def pupil_numpy(n, seed):
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    return r * np.cos(th), r * np.sin(th)


# --------------------------------------------------------------------------------------
# CPU arm 1: the oracle port (NumPy fp64), chunked over a thread pool
# --------------------------------------------------------------------------------------

def cpu_port_throughput(table, sc, n_rays, threads, chunk=REF_CHUNK, seed=123):
    """Trace `n_rays` launch rays through the table with the NumPy oracle port; return (ray-surfaces/s, seconds,
    threads that had work).  NumPy ufuncs release the GIL, so chunks run on `threads` cores."""
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
    used = max(1, min(threads, len(starts)))
    t0 = time.perf_counter()
    if used > 1:
        with ThreadPoolExecutor(used) as ex:
            list(ex.map(work, starts))
    else:
        for lo in starts:
            work(lo)
    dt = time.perf_counter() - t0
    return n_rays * n_traced / dt, dt, used


# --------------------------------------------------------------------------------------
# CPU arm 2: the STOCK reference (NumPy backend), one process per core
# --------------------------------------------------------------------------------------

_W: dict = {}


def _ref_worker_init():
    """Runs once in every worker process (forked after the parent imported the reference)."""
    import optiland.backend as be
    from optiland.samples.objectives import DoubleGauss

    be.set_backend("numpy")
    _W["lens"] = DoubleGauss()
    _W["launch"] = None


def _ref_worker_step(job):
    """One work item: REF_CHUNK launch rays (the reference's own RayGenerator, made once per worker) through
    `optic.surfaces.trace(rays)` -- the unmodified reference's hot path.  Returns (seconds in trace, checksum)."""
    from optiland.rays import RealRays

    idx, chunk = job
    lens = _W["lens"]
    if _W["launch"] is None or _W["launch"][0].size != chunk:
        Px, Py = pupil_numpy(chunk, 1000 + os.getpid() % 1000)
        z = np.zeros(chunk)
        r = lens.ray_tracer.ray_generator.generate_rays(z, z, Px, Py, WAVELENGTH)
        _W["launch"] = tuple(np.array(getattr(r, k)) for k in ("x", "y", "z", "L", "M", "N", "i", "w"))
    a = _W["launch"]
    rays = RealRays(*(v.copy() for v in a))
    t0 = time.perf_counter()
    lens.surfaces.trace(rays)
    dt = time.perf_counter() - t0
    return dt, float(rays.y[0]), int(lens.surfaces.num_surfaces)


class StockReferencePool:
    """`workers` processes, each holding its own unmodified `DoubleGauss()`; a step = every worker traces ONE chunk."""

    def __init__(self, workers):
        import multiprocessing as mp

        from oracle.ref_import import import_reference

        import_reference()          # in the parent, so that the forked workers inherit the imported modules
        self.workers = workers
        self.pool = mp.get_context("fork").Pool(workers, initializer=_ref_worker_init)

    def step(self, chunk=REF_CHUNK):
        t0 = time.perf_counter()
        res = self.pool.map(_ref_worker_step, [(i, chunk) for i in range(self.workers)], chunksize=1)
        dt = time.perf_counter() - t0
        n_surf = res[0][2]
        assert all(np.isfinite(r[1]) for r in res)
        return self.workers * chunk * (n_surf - 1), dt, n_surf

    def close(self):
        self.pool.terminate()
        self.pool.join()


def stock_reference_available():
    try:
        from oracle.ref_import import reference_available

        return reference_available()
    except Exception:
        return False


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c, sc = load_case()
    S = c.table.num_surfaces
    world = int(os.environ.get("WORLD_SIZE", "1"))
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))      # the CPU arm may use every core of the box
    except (OSError, AttributeError, ValueError):
        pass
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if stock_reference_available():
        workers = max(1, min(ncpu, 64))
        pool = StockReferencePool(workers)
        try:
            for _ in range(min(max(args.warmup, 1), 2)):
                pool.step()
            work = t_total = 0.0
            for _ in range(args.steps):
                w, dt, _ = pool.step()
                work += w
                t_total += dt
        finally:
            pool.close()
        kind, cores = "reference", workers
        sample = (f"{workers} x {REF_CHUNK} rays x {S - 1} surfaces per step: the UNMODIFIED reference "
                  f"(optiland NumPy backend, fp64), optic.surfaces.trace(rays) on pre-generated launch arrays, one "
                  f"process per core ({workers} of {ncpu} usable CPUs), wall clock over the whole step")
    else:
        threads = max(1, min(ncpu, 32))
        n_chunks = 2 * threads
        for _ in range(1):
            cpu_port_throughput(c.table, sc, threads * REF_CHUNK, threads)
        work = t_total = 0.0
        cores = threads
        for _ in range(args.steps):
            v, dt, used = cpu_port_throughput(c.table, sc, n_chunks * REF_CHUNK, threads)
            work += n_chunks * REF_CHUNK * (S - 1)
            t_total += dt
            cores = used
        kind = "port"
        sample = (f"{n_chunks} x {REF_CHUNK} rays x {S - 1} surfaces per step, NumPy fp64 oracle port (the reference is "
                  f"not on this box: run scripts/make_ref.sh), {cores} threads")
    value = work / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(S, args.rays, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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

    def wait_first(self, timeout):
        """Block (bounded) until nvidia-smi has delivered its first sample."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def count(self, t_lo):
        return sum(1 for t, _ in list(self.lines) if t >= t_lo) if self.proc is not None else 1 << 30

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
# parity gate: the timed batch against the CPU oracle
# --------------------------------------------------------------------------------------

# fp32: about 3x the errors the kernel achieves against the fp64 reference on this system (max |dx| 8.4e-6 mm,
# max |dOPD| 6.9e-5 mm on the goldens); fp64: 1e-11 x the 188 mm system scale with headroom
PARITY_TOL = {"f32": {"xy": 3e-5, "opd": 3e-4, "dir": 3e-6}, "f64": {"xy": 1e-9, "opd": 1e-9, "dir": 1e-11}}


def parity_gate(table, base, rec, n, dtype_name, sample=4096):
    """Trace a strided sample of the launch rays the timed step used through the NumPy oracle (fp64) and compare
    EVERY record row with what the kernel wrote in that step.  Returns the `parity` object; raises on failure."""
    import torch

    from oracle import trace_oracle as O

    stride = max(1, n // sample)
    idx = torch.arange(0, n, stride, device=base.x.device)[:sample]
    sub = {k: getattr(base, k)[idx].double().cpu().numpy() for k in ("x", "y", "z", "L", "M", "N", "i")}
    sub["w"] = np.full(idx.numel(), float(table.wavelengths[0]))
    _, orec, _ = O.trace(table, sub)
    err = {}
    for k in ("x", "y", "z", "L", "M", "N", "opd", "intensity"):
        got = rec[k][:, idx].double().cpu().numpy()
        want = orec[k]
        if not np.array_equal(np.isnan(got), np.isnan(want)):
            raise AssertionError(f"parity gate: NaN pattern of '{k}' differs from the oracle")
        err[k] = float(np.nanmax(np.abs(got - want))) if np.isfinite(want).any() else 0.0
    tol = PARITY_TOL[dtype_name]
    out = {"checked_rays": int(idx.numel()), "rows": int(rec["x"].shape[0]), "against": "NumPy fp64 oracle on the same launch arrays",
           "max_dx": err["x"], "max_dy": err["y"], "max_dz": err["z"], "max_dopd": err["opd"],
           "max_ddir": max(err["L"], err["M"], err["N"]), "max_dintensity": err["intensity"],
           "tol": {"xyz_mm": tol["xy"], "opd_mm": tol["opd"], "dir": tol["dir"]}}
    ok = (max(err["x"], err["y"], err["z"]) <= tol["xy"] and err["opd"] <= tol["opd"] and out["max_ddir"] <= tol["dir"]
          and err["intensity"] <= 1e-6)
    out["ok"] = bool(ok)
    if not ok:
        raise AssertionError(f"parity gate failed: {json.dumps(out)}")
    return out


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

def spiral_pupil(lo, hi, n_total, device, dtype):
    """Pupil samples as a function of the GLOBAL ray index (Vogel spiral: uniform over the unit disk): a rank's shard
    [lo, hi) of an N-ray batch holds the same rays whatever the number of ranks."""
    import torch

    j = torch.arange(lo, hi, device=device, dtype=torch.float64)
    r = torch.sqrt((j + 0.5) / n_total)
    th = j * (np.pi * (3.0 - np.sqrt(5.0)))
    return (r * torch.cos(th)).to(dtype), (r * torch.sin(th)).to(dtype)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from optiland_b200 import _lib
    from optiland_b200.distributed import bind_to_gpu_numa, shard_range
    from optiland_b200.launch import launch_infinite_angle, pupil_affine
    from optiland_b200.trace import DeviceTable, RealRays, trace_device, trace_host, trace_pupil_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: run on the CPUs (and allocate the pinned staging buffers from the memory) of the NUMA node
    # this GPU hangs off, before any pinned allocation is made
    numa = bind_to_gpu_numa(local) if not args.no_numa else {"numa_node": None, "cpus": None}
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

    def make_base(dt, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        r = torch.rand(n, generator=g, device=dev, dtype=torch.float64).sqrt()
        th = 2 * np.pi * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        x0, y0, z0, L, M, N = launch_infinite_angle(r * torch.cos(th), r * torch.sin(th), sc)
        del r, th
        b = RealRays(x0, y0, z0, L, M, N, 1.0, WAVELENGTH, dtype=dt, device=dev)
        del x0, y0, z0, L, M, N
        torch.cuda.empty_cache()
        return b

    # ---- launch rays (each rank its own pupil sample: weak scaling, no exchange) -------
    base = make_base(dtype, 1234 + rank)

    def fresh(b):
        """A RealRays view over the SAME launch arrays (the trace with records does not modify them)."""
        rr = RealRays.__new__(RealRays)
        rr.__dict__.update(b.__dict__)
        return rr

    def step(b=None):
        rr = fresh(base if b is None else b)
        rec = trace_device(dtab, rr, 0, S, record=True)
        return rr, rec

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def timed_steps(fn, steps, warmup=3):
        """K steps of `fn` bracketed by barrier + synchronize; per-step CUDA events on the launching stream.
        Returns (total ms over the K steps, mean per-step kernel ms, last result)."""
        out = None
        for _ in range(warmup):
            out = fn()
        del out
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0e.record()
        for k in range(steps):
            ev[k][0].record()
            out = fn()
            ev[k][1].record()
        t1e.record()
        barrier()
        return t0e.elapsed_time(t1e), float(np.mean([a.elapsed_time(b) for a, b in ev])), out

    # nvidia-smi needs up to a second or two to enumerate the GPUs of an 8-GPU box before its first sample: it is
    # started BEFORE the warm-up so that it is already streaming when the timed region begins
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        rr, rec = step()
    del rr, rec
    if rank == 0:
        sampler.wait_first(3.0)
    barrier()
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
    while t_clock_hi - t_lo < 0.1 or (rank == 0 and sampler.count(t_lo) < 3 and t_clock_hi - t_lo < 3.0):
        rr, rec = step()
        torch.cuda.synchronize()
        clock_extra += 1
        t_clock_hi = time.perf_counter()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    # ---- parity gate on the result of the last step (not timed) ---------------------------
    parity = parity_gate(table, base, rec, n, args.dtype)
    del rr, rec

    # ---- the other precision, driver-timed in the same run ----------------------------------
    other = "f64" if args.dtype == "f32" else "f32"
    odt = torch.float64 if other == "f64" else torch.float32
    oes = 8 if other == "f64" else 4
    obase = make_base(odt, 1234 + rank)
    o_steps = max(3, min(args.steps, 50))
    o_total, o_kern, (orr, orec) = timed_steps(lambda: step(obase), o_steps)
    o_parity = parity_gate(table, obase, orec, n, other)
    del orr, orec, obase
    torch.cuda.empty_cache()

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

    # ---- e2e through the UNMODIFIED reference's Optic.trace with the plugin installed ----------
    optic_trace = {"unavailable": "the reference is not on this box (scripts/make_ref.sh stages it under oracle/_ref)"}
    ot_ms = float("nan")
    if stock_reference_available() and not args.no_optic_trace:
        try:
            ot_ms, optic_trace = time_optic_trace(dev, n, dtype, h_pupil, h_out, e2e_steps, barrier, lib, S)
        except Exception as e:  # noqa: BLE001  (reported in the line, never silently)
            optic_trace = {"unavailable": f"{type(e).__name__}: {e}"}
    del base
    torch.cuda.empty_cache()

    # ---- strong scaling and the sharded north-star configurations (fixed TOTAL work split over the ranks) ------
    sharded = {}
    s_steps = max(3, min(args.steps, 20))

    def run_sharded(tag, case_name, total, dt, what):
        cc, ssc = load_case(case_name)
        tab = cc.table
        if world > 1:
            from optiland_b200.distributed import broadcast_table

            tab = broadcast_table(tab if rank == 0 else None, src=0)
        dtb = DeviceTable(tab, dev)
        lo, hi = shard_range(total, rank, world)
        Px, Py = spiral_pupil(lo, hi, total, dev, dt)
        affc = pupil_affine(ssc)
        Sx = tab.num_surfaces
        tot_ms, k_ms, _ = timed_steps(lambda: trace_pupil_device(dtb, Px, Py, affc, 0, Sx), s_steps)
        sharded[tag] = (tot_ms / s_steps, total, Sx, what)

    run_sharded("c2_strong_10M", "dgauss_c2", 10_000_000, dtype,
                "config 2 with the TOTAL fixed: 10 M rays (global Vogel-spiral pupil) split into contiguous shards, "
                "launch state generated in-kernel, full records per shard")
    run_sharded("c4_hubble_16M_f64", "hubble_c4", 16_000_000, torch.float64,
                "config 4: Hubble (2 conic mirrors + obscuration), 16 M rays fp64 split over the ranks, full records")

    # config 5 at its stated size and call shape: 5 fields x 3 wavelengths, per-ray (Hx, Hy, Px, Py, lambda) arrays,
    # Zernike freeform + Fresnel coatings + unpolarized PolarizedRays, 32 M rays split over the ranks; the ray's field,
    # wavelength and pupil point are functions of its GLOBAL index
    def run_c5(total=32_000_000):
        from optiland_b200.launch import pupil_affine_fields

        cc, ssc = load_case("generic_polarized_c5")
        tab = cc.table
        if world > 1:
            from optiland_b200.distributed import broadcast_table

            tab = broadcast_table(tab if rank == 0 else None, src=0)
        dtb = DeviceTable(tab, dev)
        lo, hi = shard_range(total, rank, world)
        dt = torch.float64
        j = torch.arange(lo, hi, device=dev, dtype=torch.int64)
        block = total // 15
        blk = torch.clamp(j // block, max=14)
        fields = torch.tensor([(0.0, 0.0), (0.0, 0.5), (0.0, 1.0), (0.5, 0.5), (-0.7, 0.3)], device=dev, dtype=dt)
        wl3 = torch.tensor([float(v) for v in tab.wavelengths], device=dev, dtype=dt)
        Hx, Hy = fields[blk // 3, 0].contiguous(), fields[blk // 3, 1].contiguous()
        w = wl3[blk % 3].contiguous()
        jj = (j - blk * block).double()
        rr_ = 0.95 * torch.sqrt((jj + 0.5) / block).clamp(max=1.0)
        th_ = jj * (np.pi * (3.0 - np.sqrt(5.0)))
        Px, Py = (rr_ * torch.cos(th_)).contiguous(), (rr_ * torch.sin(th_)).contiguous()
        del j, blk, jj, rr_, th_
        affc = pupil_affine_fields(ssc, Hx, Hy)
        Sx = tab.num_surfaces
        tot_ms, _, _ = timed_steps(lambda: trace_pupil_device(dtb, Px, Py, affc, 0, Sx, wavelength=w, polarization=None), s_steps)
        sharded["c5_polarized_zernike_32M_f64"] = (
            tot_ms / s_steps, total, Sx,
            "config 5: Zernike freeform + Fresnel coatings + unpolarized PolarizedRays, 5 fields x 3 wavelengths as per-ray "
            "arrays (trace_generic's call shape), 32 M rays fp64 split over the ranks: fused launch, full records, P matrices "
            "and the update_intensity epilogue")

    run_c5()
    torch.cuda.empty_cache()

    # config 3's differentiable step over a fixed 4 M rays: forward + adjoint kernels per shard, 4-scalar all-reduce for
    # the loss, all-reduce(sum) of the parameter gradients
    def run_c3_grad(total=4_000_000):
        from optiland_b200 import autograd as AG
        from optiland_b200.distributed import sharded_rms_spot_loss_and_grad

        cc, _ = load_case("telephoto_c3_tol1e-6")
        tab = cc.table
        if world > 1:
            from optiland_b200.distributed import broadcast_table

            tab = broadcast_table(tab if rank == 0 else None, src=0)
        lo, hi = shard_range(total, rank, world)
        gen = torch.Generator(device=dev).manual_seed(99)
        idx = torch.randint(0, cc.n, (total,), device=dev, generator=gen)[lo:hi]
        rr_ = {k: torch.from_numpy(v).to(dev)[idx] for k, v in cc.rays.items()}
        rays = RealRays(rr_["x"], rr_["y"], rr_["z"], rr_["L"], rr_["M"], rr_["N"], rr_["i"], rr_["w"], dtype=torch.float32, device=dev)
        params = AG.table_to_params(tab).to(dev).requires_grad_(True)

        def trace_fn(p):
            rec = AG.trace_differentiable(tab, p, fresh(rays), rows=(-1,))
            return rec["x"], rec["y"]

        tot_ms, _, _ = timed_steps(lambda: sharded_rms_spot_loss_and_grad(trace_fn, params), s_steps)
        sharded["c3_grad_step_4M_f32"] = (
            tot_ms / s_steps, total, tab.num_surfaces,
            "config 3: reverse telephoto with 2 even aspheres, forward + adjoint kernels over a fixed 4 M rays split over the "
            "ranks + all-reduce of the loss moments and of the (S, 28) parameter-gradient block; wall time of the whole step")

    run_c3_grad()
    torch.cuda.empty_cache()

    # ---- max over ranks ----------------------------------------------------------------
    names = list(sharded)
    vals = [total_ms, kern_ms, e2e_s * 1e3, e2e_state_s * 1e3, spot_s * 1e3, o_total, o_kern, ot_ms] + [sharded[k][0] for k in names]
    times = torch.tensor([v if np.isfinite(v) else -1.0 for v in vals], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    tl = [float(v) for v in times.cpu()]
    total_ms, kern_ms, e2e_ms, e2e_state_ms, spot_ms, o_total, o_kern, ot_ms = tl[:8]
    for k, v in zip(names, tl[8:]):
        sharded[k] = (v,) + sharded[k][1:]

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

        def roof(esz, k_ms, dname):
            bpr = esz * (n_loads + 8 * S)  # full records, final state aliased to the last row
            ach = bpr * n / (k_ms * 1e-3) / 1e9
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dname)
            except (OSError, ValueError):
                pass
            return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "static: profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                      "`ncu --set full` capture of this kernel at this size; not measured in this run)",
                    "peak_source": peak_kind, "algorithmic_bytes_per_ray": bpr, "kernel_ms": k_ms,
                    "kernel": "olb::trace_kernel<%s,%d,0>" % (("float", 4) if esz == 4 else ("double", 1))}

        value = world * n * n_traced / (total_ms * 1e-3 / args.steps)
        # CPU baseline: the stock reference when it is on the box (bounded sample), the port beside it
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        port_threads = max(1, min(ncpu, 32))
        port_n = 32 * REF_CHUNK
        port_val, port_dt, port_used = cpu_port_throughput(table, sc, port_n, port_threads)
        cpu_port = {"value": port_val, "unit": UNIT, "cores": port_used, "kind": "port",
                    "sample": f"{port_n} rays x {n_traced} surfaces in {port_dt:.1f} s, NumPy fp64 oracle port (the reference's "
                              f"arithmetic without its per-call material cache-key hashing), {port_used} threads over "
                              f"{port_n // REF_CHUNK} chunks of {REF_CHUNK} rays"}
        cpu_baseline = cpu_port
        if world == 1 and stock_reference_available() and not args.no_cpu_reference:
            # the stock reference in its own process (a clean interpreter: no CUDA context to fork, every core usable)
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                                      "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600).stdout
                ref_line = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
                cpu_baseline = ref_line["cpu_baseline"]
            except Exception as e:  # noqa: BLE001
                cpu_baseline = dict(cpu_port, note=f"stock reference arm failed ({type(e).__name__}: {e}); port reported")
        cfg = workload_config(S, n, world)
        cfg.update({"l2_policy": f"inputs+records {es * (n_loads + 8 * S) * n / 1e9:.2f} GB per step >> 126 MB L2",
                    "parallelism": f"rays sharded over {world} GPU(s), table broadcast, no exchange",
                    "numa_binding": numa})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "roofline": roof(es, kern_ms, args.dtype),
            "parity": parity,
            other: {"ms_per_step": o_total / o_steps, "steps": o_steps, "value": world * n * n_traced / (o_total * 1e-3 / o_steps),
                    "roofline": roof(oes, o_kern, other), "parity": o_parity,
                    "what": f"the same workload in {other}, timed in the same run (CUDA events, barrier + synchronize)"},
            "cpu_baseline": cpu_baseline,
            "cpu_baseline_port": cpu_port,
            "e2e": {"value": world * n * n_traced / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "what": "Optic.trace-shaped call through olb_trace_host_pupil_*: pinned host pupil samples (Px, Py) "
                            "-> H2D -> launch state generated in-kernel -> trace (records stay in HBM) -> D2H of the "
                            "final ray state (x,y,z,L,M,N,i,opd); 1 Mi-ray chunks on 3 streams"},
            "e2e_optic_trace": dict(optic_trace, **({"value": world * n * n_traced / (ot_ms * 1e-3), "unit": UNIT,
                                                     "ms_per_step": ot_ms} if ot_ms > 0 else {})),
            "e2e_launch_arrays": {"value": world * n * n_traced / (e2e_state_ms * 1e-3), "unit": UNIT,
                                  "h2d_bytes_per_step": h2d_state, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_state_ms,
                                  "what": "SurfaceGroup.trace-shaped call through olb_trace_host_*: the 7 launch-state "
                                          "arrays cross PCIe instead of the 2 pupil arrays"},
            "e2e_spot_rms": {"value": world * n * n_traced / (spot_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                             "d2h_bytes_per_step": 64, "ms_per_step": spot_ms, "rms_spot_radius_mm": rms_val,
                             "what": "SpotDiagram-shaped call (next rows f-1 + f-2): pinned host pupil samples -> H2D -> "
                                     "launch generation + trace + spot moments fused in one kernel, NO per-ray output "
                                     "(no records) -> D2H of 8 doubles; not the headline (it skips the records)"},
            "sharded_fixed_total": {k: {"rays_total": tot, "ms_per_step": ms, "value": tot * (Sx - 1) / (ms * 1e-3), "unit": UNIT,
                                        "scaling": "strong", "steps": s_steps, "what": what}
                                    for k, (ms, tot, Sx, what) in sharded.items()},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_optic_trace(dev, n, dtype, h_pupil, h_out, steps, barrier, lib, S):
    """`Optic.trace` of the unmodified reference (live `DoubleGauss()` object, torch backend on the device) with the
    plugin installed: per step the pinned host pupil samples are copied to the device, `lens.trace(...)` is called
    exactly as a user would (optic/optic.py:715-740), and the final ray state is copied back to pinned host memory."""
    import torch

    from oracle.ref_import import import_reference

    import_reference()
    import optiland.backend as be
    from optiland.distribution import BaseDistribution
    from optiland.samples.objectives import DoubleGauss

    from optiland_b200 import plugin as P

    class HostPupil(BaseDistribution):
        def generate_points(self, num_points=None):
            self.x = h_pupil["Px"].to(dev, non_blocking=True)
            self.y = h_pupil["Py"].to(dev, non_blocking=True)

    be.set_backend("torch")
    be.set_device("cuda")
    be.set_precision("float32" if dtype == torch.float32 else "float64")
    be.grad_mode.disable()
    P.install()
    try:
        lens = DoubleGauss()
        dist_obj = HostPupil()

        def one():
            dist_obj.generate_points()
            rays = lens.trace(Hx=0.0, Hy=0.0, wavelength=WAVELENGTH, num_rays=n, distribution=dist_obj)
            for k in ("x", "y", "z", "L", "M", "N", "i", "opd"):
                h_out[k].copy_(getattr(rays, k), non_blocking=True)
            torch.cuda.synchronize(dev)
            return float(h_out["x"][:1024].mean())

        for _ in range(2):
            one()
        P.stats(reset=True)
        l0 = lib.olb_launch_count()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            v = one()
        torch.cuda.synchronize(dev)
        ms = 1e3 * (time.perf_counter() - t0) / steps
        launches = lib.olb_launch_count() - l0
        declines = P.stats()
        assert np.isfinite(v)
        if declines or launches < steps:
            raise RuntimeError(f"Optic.trace did not run on the kernel: declines={declines}, launches={launches}")
        es = 4 if dtype == torch.float32 else 8
        info = {"h2d_bytes_per_step": 2 * es * n, "d2h_bytes_per_step": 8 * es * n, "olb_launches_per_step": launches / steps,
                "declines": declines,
                "what": "the UNMODIFIED reference's Optic.trace(Hx, Hy, wavelength, num_rays, distribution) on a live "
                        "DoubleGauss() with optiland_b200.plugin installed (torch backend, device cuda): pinned host pupil "
                        "samples -> H2D -> RealRayTracer.trace -> ONE fused launch (launch generation + 13 surfaces + "
                        "records handed back to the Surface objects) -> D2H of the final ray state; sequential, not chunked"}
        return ms, info
    finally:
        P.uninstall()
        be.set_device("cpu")
        be.set_backend("numpy")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--no-optic-trace", action="store_true", help="skip the e2e_optic_trace leg (live Optiland objects)")
    ap.add_argument("--no-cpu-reference", action="store_true", help="cpu_baseline from the port only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
