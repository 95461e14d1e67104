#!/usr/bin/env python
"""cuSten-B200 benchmark: 2D stencil Gpoints/s and HBM GB/s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload (every N): BASELINE.json configs[3], the 2D XY periodic function-pointer stencil (nonlinear
c^3 - c term through a 3x3 Laplacian, cuPentCahnADI.cu:164-188) on a 32768^2 FP64 grid — the configuration the
multi-GPU target is quoted on; it fits one B200 (2 x 8 GiB), so N = 1 runs the same grid and the scaling is strong.
A step is one TIME STEP of the whole grid: Compute + Swap (the output of a step is the input of the next), so at N > 1
every sweep depends on the neighbours' previous one and the halo rows it reads over NVLink are fresh every step.
After the timed loop every rank checks the rows either side of its seams against the oracle (`parity` key; a
mismatch fails the run).  At N = 1 the line also carries a per-variant table on 16384^2 (the single-GPU target size)
and the configs[1] / configs[2] workloads.

`value` is timed with inputs resident in HBM; `e2e` goes through the same C ABI with the grid in pinned HOST
memory (the numTiles out-of-core path: H2D of every tile and D2H of every result inside the timed region).
--impl reference times the reference's own CPU implementation of this path (serialCahnADI.c nonlinearRHS,
compiled from /root/reference into oracle/_ref/) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "2d_stencil_gpoints_per_s"
UNIT = "Gpoints/s"
ALG_BYTES_PER_POINT = 16.0  # one FP64 read + one FP64 write per grid point per sweep

WORKLOADS = {
    # name: (variant, n, tiles, description)
    "xy_p_fun_32768": ("XYpFun", 32768, 1, "configs[3]: 2D XY periodic Fun stencil (c^3-c, 3x3), 32768^2 FP64"),
    "x_p_8192": ("Xp", 8192, 1, "configs[1]: 2D X periodic 9-pt 8th-order d2/dx2 (2d_x_p example), 8192^2 FP64"),
    "xy_np_16384_t4": ("XYnp", 16384, 4, "configs[2]: 2D XY non-periodic 3x3 cross stencil, 16384^2 FP64, numTiles=4"),
    "xy_p_fun_16384": ("XYpFun", 16384, 1, "2D XY periodic Fun stencil (c^3-c, 3x3), 16384^2 FP64"),
    "xy_p_16384": ("XYp", 16384, 1, "2D XY periodic 3x3 cross stencil (2d_xy_p example), 16384^2 FP64"),
}


SEED = 0x5EED


def _w_d2_8th(h):
    """9-point 8th-order second derivative (examples/src/2d_x_p.cu:99-114)."""
    return np.array([-1.0 / 560, 8.0 / 315, -1.0 / 5, 8.0 / 5, -205.0 / 72, 8.0 / 5, -1.0 / 5, 8.0 / 315, -1.0 / 560]) / (h * h)


def _w_cross_xy(dx, dy):
    """3x3 cross derivative d2/dxdy (examples/src/2d_xy_p.cu:112-120)."""
    s = 1.0 / (4.0 * dx * dy)
    return np.array([s, 0.0, -s, 0.0, 0.0, 0.0, -s, 0.0, s])


def _w_laplace5(sig):
    """3x3 five-point Laplacian times sigma (cuPentCahnADI.cu:511-513)."""
    return np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=np.float64) * sig


def stencil_args(variant, n, time_stepping=False):
    """Coefficients / window of the named workloads (SURVEY.md section 8d 'synthetic inputs')."""
    h = 2 * np.pi / n
    d = "XY" if variant.startswith("XY") else variant[0]
    fun = None
    if d == "X":
        coef, kw = _w_d2_8th(h), dict(H=9, L=4, R=4)
        if variant.endswith("Fun"):
            fun, kw["numCoe"] = "weighted9_x", 9
    elif d == "Y":
        coef, kw = _w_d2_8th(h), dict(V=9, T=4, B=4)
        if variant.endswith("Fun"):
            fun = "weighted9_y"
            if variant == "YpFun":
                kw["numCoe"] = 9
    elif variant.endswith("Fun"):
        kw, fun = dict(H=3, L=1, R=1, V=3, T=1, B=1), "cubic_xy"
        if time_stepping:
            # the same user function (sum of coe * (c^3 - c) over the 3x3 window) with coefficients that make the
            # map c -> f(c) a bounded iteration: c' = h + eps Lap5(h), h = c - c^3 (an explicit Allen-Cahn-like step).
            # With the solver's own sigma_N (~43 at this size) the bare stencil is not a time stepper: it overflows
            # within a few applications.  Arithmetic per point is identical.
            eps = 0.05
            coef = np.array([0, -eps, 0, -eps, -1 + 4 * eps, -eps, 0, -eps, 0], dtype=np.float64)
        else:
            # sigma_N * 5-point Laplacian applied to c^3 - c (cuPentCahnADI.cu:496-516), dt = 0.1 dx, D = 1
            dx = 16 * np.pi / n
            coef = _w_laplace5((0.1 * dx / 3.0) * 2.0 / dx ** 2)
    else:
        coef, kw = _w_cross_xy(h, h), dict(H=3, L=1, R=1, V=3, T=1, B=1)
        if time_stepping:
            coef = coef / np.abs(coef).sum()
    kw["fun"] = fun
    return np.ascontiguousarray(coef, dtype=np.float64), kw


def hash_rows(row0, rows, nx, seed, lo, hi):
    """numpy twin of custen_fill_hash (custen_b200/csrc/api_c.cu): rows [row0, row0 + rows) of the synthetic field."""
    idx = (np.uint64(row0) * np.uint64(nx) + np.arange(rows * nx, dtype=np.uint64))
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + (hi - lo) * u).reshape(rows, nx)


def seam_parity(variant, coef, kw, n, total_steps, row_lo, row_hi, final_rows_of, margin=2):
    """Checker (the oracle, oracle/custen_oracle.c - never on the timed path): after `total_steps` time steps, the
    `margin` rows either side of each end of this rank's slab [row_lo, row_hi) must equal what the oracle gets by
    time-stepping just the band of the GLOBAL initial field those rows depend on (x wraps, the band shrinks by T + B
    rows per step).  final_rows_of(global_row_list) returns this rank's final values of those rows.
    Returns (rows checked, doubles differing)."""
    import oracle_lib as ol
    T, B = kw.get("T", 0), kw.get("B", 0)
    fun = kw.get("fun")
    okw = {k: v for k, v in kw.items() if k in ("H", "L", "R", "V")}
    checked = differing = 0
    targets = [(row_lo, min(row_lo + margin, row_hi)), (max(row_hi - margin, row_lo), row_hi)]
    for g0, g1 in targets:
        b0, b1 = g0 - total_steps * T, g1 + total_steps * B
        rows = [(r % n) for r in range(b0, b1)]
        # contiguous runs of the wrapped row list, regenerated from the counter-based hash
        parts, start = [], 0
        for i in range(1, len(rows) + 1):
            if i == len(rows) or rows[i] != rows[i - 1] + 1:
                parts.append(hash_rows(rows[start], i - start, n, SEED, -0.1, 0.1))
                start = i
        band = np.ascontiguousarray(np.vstack(parts))
        got_band, v0, v1 = ol.oracle_evolve_band(variant, band, coef, total_steps, T, B, fun=fun, **okw)
        want = got_band[v0:v1]
        assert want.shape[0] == g1 - g0
        got = final_rows_of(list(range(g0, g1)))
        differing += ol.count_diff(got, want)
        checked += g1 - g0
    return checked, differing


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_from=None, t_to=None):
        """Summary of the samples taken in [t_from, t_to] (perf_counter seconds; default: all of them)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        self.th.join(timeout=2)
        rows = [r for t, r in self.rows if (t_from is None or t >= t_from) and (t_to is None or t <= t_to)]
        sm = [float(r[0]) for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's serial CPU code (oracle/_ref/libserialcahn.so)
# ------------------------------------------------------------------------------------------------------------

def _serial_lib():
    import oracle_lib as ol
    lib = ol.serial()
    if lib is None:
        raise RuntimeError("oracle/_ref/libserialcahn.so not available")
    return lib


def cpu_sweeps(n, threads, repeats):
    """`threads` host threads, each sweeping its own n x n field `repeats` times with the reference's
    nonlinearRHS (serialCahnADI.c:553-622).  Returns seconds for the whole batch."""
    lib = _serial_lib()
    _dp = ctypes.POINTER(ctypes.c_double)
    dx = 16 * np.pi / n
    w = np.ascontiguousarray(_w_laplace5((0.1 * dx / 3.0) * 2.0 / dx ** 2))
    rng = np.random.default_rng(1)
    bufs = [(np.ascontiguousarray(rng.uniform(-0.1, 0.1, (n, n))), np.zeros((n, n))) for _ in range(threads)]

    def work(i):
        a, b = bufs[i]
        for _ in range(repeats):
            lib.nonlinearRHS(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), w.ctypes.data_as(_dp), 3, 3, 1, 1, n)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n = 2048
    for _ in range(args.warmup):
        cpu_sweeps(n, cores, 1)
    t = cpu_sweeps(n, cores, args.steps)
    value = cores * n * n * args.steps / t / 1e9
    variant, size, tiles, desc = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "xy_p_fun_%d_per_core" % n, "sample_of": args.workload, "description": desc,
                   "same_config_as_gpu_arm": False,
                   "note": "reference CPU implementation of the path: serialCahnADI.c nonlinearRHS (3x3 c^3-c stencil, periodic); "
                           "each host thread sweeps its own %d^2 periodic field (that function only takes square grids and is "
                           "serial), a bounded and cache-friendlier sample of the 32768^2 workload, so the ratio against it is "
                           "conservative; the reference's CUDA library on the same B200 is in profiles/r1_reference_gpu_16384.json" % n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{cores} threads x {args.steps} sweeps of an independent {n}^2 periodic field each"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------

def time_resident(cs, st, steps, warmup, pre_step=None):
    """Median-free contract timing: W warm-ups, then exactly K steps between two events on the launching stream."""
    lib = cs.load()
    h = ctypes.addressof(st.handle)
    for _ in range(warmup):
        if pre_step:
            pre_step()
        st.compute(cs.DEVICE)
    cs.device_synchronize()
    e0, e1 = lib.custen_event_create(), lib.custen_event_create()
    time_resident.launches = -cs.launch_count()
    lib.custen_event_record(e0, h, 0)
    for _ in range(steps):
        if pre_step:
            pre_step()
        st.compute(cs.DEVICE)
    lib.custen_event_record(e1, h, 0)
    lib.custen_event_synchronize(e1)
    time_resident.launches += cs.launch_count()
    ms = lib.custen_event_elapsed_ms(e0, e1)
    lib.custen_event_destroy(e0)
    lib.custen_event_destroy(e1)
    return ms


def variant_table(cs, torch, n, steps, warmup, peak):
    """All 12 variants on an n x n grid, device-resident (the '>= 80 % of HBM peak on 16384^2' target)."""
    table = {}
    inp = (torch.rand((n, n), device="cuda", dtype=torch.float64) * 0.2 - 0.1)
    out = torch.zeros_like(inp)
    for v in cs.VARIANTS:
        coef, kw = stencil_args(v, n)
        tc = torch.from_numpy(np.ascontiguousarray(coef)).cuda()
        st = cs.Stencil2D(v, n, n, out, inp, tc, **kw)
        ms = time_resident(cs, st, steps, warmup) / steps
        gpts = n * n / ms / 1e6
        table[v] = {"gpoints_per_s": round(gpts, 2), "hbm_gbs": round(gpts * ALG_BYTES_PER_POINT, 1),
                    "frac_of_peak": round(gpts * ALG_BYTES_PER_POINT / peak, 4), "path": st.path}
        if v.endswith("Fun"):  # the same call through the opaque device pointer (unregistered user function)
            cs.set_tuning(force_opaque=1)
            ms = time_resident(cs, st, steps, warmup) / steps
            cs.set_tuning()
            table[v]["opaque_pointer_gpoints_per_s"] = round(n * n / ms / 1e6, 2)
        st.destroy()
    # 13th variant: WENO5 advection reads phi, u, v and writes one field: 32 B per point.  Two inputs: the
    # fields of the reference's own program (examples/src/2d_xyWENOADV_p.cu:97-101: smooth phi, rotating velocity) and
    # uniform random phi / velocities (the upwind side changes from point to point)
    lib = cs.load()
    x = torch.arange(n, device="cuda", dtype=torch.float64) * (2 * torch.pi / n)
    fields = {
        "example": lambda: ((torch.cos(x)[None, :] * torch.sin(x)[:, None]).contiguous(),
                            torch.sin(x)[:, None].expand(n, n).contiguous(), (-torch.sin(x))[None, :].expand(n, n).contiguous()),
        "random": lambda: (torch.rand((n, n), device="cuda", dtype=torch.float64),
                           torch.rand((n, n), device="cuda", dtype=torch.float64) * 2 - 1,
                           torch.rand((n, n), device="cuda", dtype=torch.float64) * 2 - 1),
    }
    res = {}
    for fname, make in fields.items():
        phi, u, v = make()
        h = cs.cuSten_t()
        cs.cuStenCreate2DXYWENOADVp(h, torch.cuda.current_device(), 1, n, n, 32, 32, 2 * np.pi / n, 2 * np.pi / n, u, v, out, phi)
        hp = ctypes.addressof(h)
        for _ in range(warmup):
            cs.cuStenCompute2DXYWENOADVp(h, cs.DEVICE)
        cs.device_synchronize()
        e0, e1 = lib.custen_event_create(), lib.custen_event_create()
        lib.custen_event_record(e0, hp, 0)
        for _ in range(steps):
            cs.cuStenCompute2DXYWENOADVp(h, cs.DEVICE)
        lib.custen_event_record(e1, hp, 0)
        lib.custen_event_synchronize(e1)
        ms = lib.custen_event_elapsed_ms(e0, e1) / steps
        res[fname] = (n * n / ms / 1e6, cs.last_path(h))
        cs.cuStenDestroy2DXYWENOADVp(h)
        lib.custen_event_destroy(e0)
        lib.custen_event_destroy(e1)
        del phi, u, v
    gpts = res["random"][0]
    table["XYWENOADVp"] = {"gpoints_per_s": round(gpts, 2), "hbm_gbs": round(gpts * 32.0, 1),
                           "frac_of_peak": round(gpts * 32.0 / peak, 4), "path": res["random"][1],
                           "example_fields_gpoints_per_s": round(res["example"][0], 2),
                           "note": "32 B/point algorithmic (phi, u, v in; one field out); compute-bound (18 single-precision "
                                   "powf per point, kept for bit parity with the reference kernel); headline = random fields, "
                                   "the workload of round 1 and of the reference kernel's figure "
                                   "(profiles/r1_reference_gpu_16384.json); example_fields = the reference example's own fields"}
    return table


def cahn_on_slabs(cs, torch, dist, rank, world, local_rank, n=4096, steps=40, check_steps=4):
    """Config 5 on `world` GPUs (custen_cahn_slab_*: no collective on the data path, NCCL only moves IPC handles).
    Parity inside the run: every rank also steps the whole grid with the single-GPU solver on its own GPU and compares
    its slab's rows bit for bit (partitions are solved with the same arithmetic wherever they live); rank 0 adds the
    distance of that road from the bit-identical one (= the reference's GPU solver, tests/test_cahn_gpu.py)."""
    from custen_b200.cahn import CahnHilliard, CahnHilliardSlab
    rows = n // world
    c0 = np.random.default_rng(0).uniform(-0.1, 0.1, (n, n))
    mine = c0[rank * rows:(rank + 1) * rows]
    slab = CahnHilliardSlab(n, device=local_rank)
    slab.set_field(mine)
    slab.step(check_steps)
    got = slab.field()
    single = CahnHilliard(n, device=local_rank, solver=2)
    single.set_field(c0)
    single.step(check_steps)
    whole = single.field()
    single.destroy()
    torch.cuda.set_device(local_rank)
    differing = int(np.count_nonzero(got.view(np.int64) != whole[rank * rows:(rank + 1) * rows].view(np.int64)))
    rel = None
    if rank == 0:
        exact = CahnHilliard(n, device=local_rank, solver=0)
        exact.set_field(c0)
        exact.step(check_steps)
        ref = exact.field()
        exact.destroy()
        torch.cuda.set_device(local_rank)
        rel = float(np.max(np.abs(whole - ref)) / np.max(np.abs(ref)))
    slab.set_field(mine)
    slab.step(5)
    slab.synchronize()
    dist.barrier()
    ms = slab.time_steps(steps)
    t = torch.tensor([ms, float(differing), float(slab.error())], device="cuda", dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    slab.destroy()
    ms = float(tmax[0].item())
    return {"n_gpus": world, "ms_per_step": round(ms, 4), "mpoint_steps_per_s": round(n * n / ms / 1e3, 1), "solver": 2,
            "steps_timed": steps, "timing": "CUDA events on every slab's stream, max over ranks",
            "parity": {"bits_differing_vs_single_gpu": int(t[1].item()), "neighbour_wait_timeouts": int(t[2].item()),
                       "steps_compared": check_steps, "rel_vs_bit_identical_road": rel},
            "exchange": "per step and GPU: 2 x 2 halo rows of c and c(t - dt) and the 4 interface values per system of "
                        "the y-partitions within reach of the seam, read in place over NVLink; no all-to-all",
            "note": "tolerance-mode road (custen_b200/csrc/cahn_part.cu); the distance to the bit-identical road at 4096^2 is "
                    "the reference solver's own rounding error (tests/test_pent_part_cpu.py)"}


def run_ours(args, rank, world, local_rank):
    import torch
    import custen_b200 as cs
    import custen_b200.slab as slab

    torch.cuda.set_device(local_rank)
    # keep this rank (and the pinned buffers it allocates) on the CPUs / memory next to its GPU
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    variant, n, tiles, desc = WORKLOADS[args.workload]
    coef_ts, kw = stencil_args(variant, n, time_stepping=True)
    coef, _ = stencil_args(variant, n)
    rows = n // world
    peak, peak_src = peaks()
    lib = cs.load()
    tcoef = torch.from_numpy(coef).cuda()
    T, B = (0, 0) if variant[0] == "X" and not variant.startswith("XY") else (kw.get("T", 0), kw.get("B", 0))

    # ---- resident timing: K time steps (Compute + Swap) through the C slab layer -----------------------------------
    # world == 1 is the same code with a single slab whose halo rows are its own far edge (plain periodic wrap).
    ss = slab.SlabStencil(variant, n, n, coef_ts, transport=args.transport, numTiles=tiles, **kw)
    lib.custen_fill_hash(ss.input.data_ptr(), ss.row0, rows, n, SEED, -0.1, 0.1)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)   # started ahead of the warm-up: nvidia-smi needs a moment to deliver its first line
    ss.run(args.warmup)
    ss.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = cs.launch_count()
    t_timed0 = time.perf_counter()
    if ss.slab:
        ms = float(lib.custen_slab_time_run(ss.slab, args.steps))     # events on the slab's stream, synchronises
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ss.run(args.steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    launches_timed = cs.launch_count() - l0
    t_timed1 = time.perf_counter()
    if dist:
        dist.barrier()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    path = ss.path
    timed_out = ss.error()

    # ---- parity: the rows either side of this rank's seams (and one interior row) against the oracle ---------------
    parity = None
    if not args.no_parity:
        total = args.warmup + args.steps
        final = ss.input

        def final_rows_of(global_rows):
            idx = torch.tensor([r - ss.row0 for r in global_rows], device="cuda")
            return final.index_select(0, idx).cpu().numpy()

        t0p = time.perf_counter()
        checked, differing = seam_parity(variant, coef_ts, kw, n, total, ss.row0, ss.row0 + rows, final_rows_of)
        mid = ss.row0 + rows // 2
        c2, d2 = seam_parity(variant, coef_ts, kw, n, total, mid, mid + 1, final_rows_of, margin=1)
        checked, differing = checked + c2 // 2, differing + d2 // 2   # the one-row target is visited as both of its ends
        pv = torch.tensor([checked, differing, int(timed_out)], device="cuda", dtype=torch.int64)
        if dist:
            dist.all_reduce(pv)
        parity = {"rows_checked": int(pv[0]), "bits_differing": int(pv[1]), "neighbour_wait_timeouts": int(pv[2]),
                  "time_steps_compared": total, "checker_seconds": round(time.perf_counter() - t0p, 1),
                  "what": "final rows next to every slab seam (2 either side) and one interior row per rank, against "
                          "oracle/custen_oracle.c time-stepping the band of the regenerated global initial field they "
                          "depend on; doubles compared bit for bit"}

    # halo traffic: (T + B) rows of nx doubles per GPU per sweep; the same rows timed on their own as an NCCL exchange
    halo = None
    if world > 1 and T + B:
        periodic = not variant.replace("Fun", "").endswith("np")
        top = torch.empty((max(T, 1), n), device="cuda", dtype=torch.float64)
        bot = torch.empty((max(B, 1), n), device="cuda", dtype=torch.float64)
        inp = ss.input
        for _ in range(3):
            slab.exchange_halos(inp, T, B, top[:T], bot[:B], rank, world, periodic)
        torch.cuda.synchronize()
        dist.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(20):
            slab.exchange_halos(inp, T, B, top[:T], bot[:B], rank, world, periodic)
        h1.record()
        torch.cuda.synchronize()
        th = torch.tensor([h0.elapsed_time(h1) / 20], device="cuda", dtype=torch.float64)
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
        hb = (T + B) * n * 8
        halo = {"bytes_received_per_gpu_per_sweep": hb, "transport_in_timed_region": args.transport,
                "nccl_exchange_us": round(float(th.item()) * 1e3, 1),
                "nccl_exchange_nvlink_gbs_per_gpu": round(hb / (float(th.item()) * 1e-3) / 1e9, 2),
                "in_sweep_gbs_per_gpu": round(hb / (ms / args.steps * 1e-3) / 1e9, 2),
                "note": "peer transport: the sweep's TMA producer reads the rows from the neighbours' memory and waits for "
                        "the neighbour inside the kernel, in front of the halo rows only - no exchange step, no barrier "
                        "kernel; the NCCL figure is the same rows sent with send/recv on their own (latency-bound)"}
        del top, bot
    # clocks and throttle reasons under this load.  K steps take 7 - 60 ms here, less than nvidia-smi's sampling period, so
    # when the timed region was too short to be sampled reliably the same loop is repeated UNTIMED for half a second (after
    # the parity check has read its rows) and sampled there; the same number of extra steps on every rank.
    if ms >= 400.0:
        clocks = sampler.stop(t_timed0, t_timed1)
        clocks["window"] = "timed region"
    else:
        extra = int(min(20000, max(args.steps, 500.0 / (ms / args.steps))))
        if dist:
            dist.barrier()
        t_rep0 = time.perf_counter()
        ss.run(extra)
        ss.synchronize()
        t_rep1 = time.perf_counter()
        clocks = sampler.stop(t_rep0 + 0.05, t_rep1)
        clocks["window"] = ("untimed repeat of the timed loop (%d more steps, %.0f ms) right after it: the timed region, %.1f ms, "
                            "is shorter than nvidia-smi's sampling period" % (extra, 1e3 * (t_rep1 - t_rep0), ms))
    ss.destroy()
    ms_per_step = ms / args.steps
    value = n * n / ms_per_step / 1e6  # Gpoints/s, whole job

    # ---- end to end: grid in pinned host memory, through the out-of-core tile scheduler -----------------------
    e2e = None
    if not args.no_e2e:
        nbytes = rows * n * 8
        node = ctypes.c_int(-1)
        h_in = lib.custen_host_alloc_near(nbytes, local_rank, ctypes.byref(node))
        h_out = lib.custen_host_alloc_near(nbytes, local_rank, None)
        near = bool(h_in and h_out)
        if not near:   # mmap / registration refused: ordinary pinned memory
            h_in, h_out = lib.custen_host_alloc(nbytes), lib.custen_host_alloc(nbytes)
        v_in = np.ctypeslib.as_array((ctypes.c_double * (rows * n)).from_address(h_in))
        tmp = torch.empty((rows, n), device="cuda", dtype=torch.float64)
        lib.custen_fill_hash(tmp.data_ptr(), rank * rows, rows, n, SEED, -0.1, 0.1)
        torch.from_numpy(v_in).copy_(tmp.view(-1))
        torch.cuda.synchronize()
        del tmp
        torch.cuda.empty_cache()
        # what the host link carries with plain copies both ways at once, all ranks at the same time
        if dist:
            dist.barrier()
        lib.custen_link_probe(h_in, h_out, nbytes, 1, local_rank)
        if dist:
            dist.barrier()
        probe_ms = float(lib.custen_link_probe(h_in, h_out, nbytes, 2, local_rank)) / 2
        tp = torch.tensor([probe_ms], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        probe_ms = float(tp.item())
        e2e_tiles = max(tiles, args.e2e_tiles)
        st = cs.Stencil2D(variant, n, rows, h_out, h_in, tcoef, numTiles=e2e_tiles, deviceNum=local_rank, **kw)
        halo_bytes = 0
        if world > 1 and T + B:
            top = torch.empty((max(T, 1), n), device="cuda", dtype=torch.float64)
            bot = torch.empty((max(B, 1), n), device="cuda", dtype=torch.float64)
            periodic = not variant.replace("Fun", "").endswith("np")
            st.set_slab(top, bot, rank == 0, rank == world - 1)
            h_view = torch.from_numpy(v_in).view(rows, n)
            halo_bytes = (T + B) * n * 8

            def pre():
                # edge rows of the host-resident slab travel H2D, then over NVLink to the neighbours
                first_rows = h_view[:max(B, 1)].cuda(non_blocking=True)
                last_rows = h_view[-max(T, 1):].cuda(non_blocking=True)
                # exchange_halos only touches local[:B] and local[-T:], so a two-piece stand-in is enough
                standin = torch.cat([first_rows, last_rows])
                slab.exchange_halos(standin, T, B, top[:T], bot[:B], rank, world, periodic)
        else:
            pre = None

        def e2e_step():
            if pre:
                pre()
            st.compute(cs.HOST)

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        cs.device_synchronize()
        if dist:
            dist.barrier()
        k = max(1, min(args.steps, args.e2e_steps))
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(k):
            e2e_step()
        t1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([t0.elapsed_time(t1)], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / k
        nodes = [None] * world
        info = {"gpu_numa_node": int(lib.custen_device_numa_node(local_rank)), "buffers_bound_to_node": int(node.value),
                "cpus": sorted(os.sched_getaffinity(0))[:: max(1, len(os.sched_getaffinity(0)) // 4)][:4]}
        if dist:
            dist.all_gather_object(nodes, info)
        else:
            nodes = [info]
        ceiling_gbs = world * 2 * nbytes / (probe_ms * 1e-3) / 1e9      # both directions, all ranks
        moved_gbs = world * (2 * nbytes + halo_bytes) / (e2e_ms * 1e-3) / 1e9
        e2e = {"value": n * n / e2e_ms / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(world * (nbytes + halo_bytes)),
               "d2h_bytes_per_step": int(world * nbytes), "ms_per_step": e2e_ms, "steps": k, "numTiles": e2e_tiles,
               "host_link_gbs": round(moved_gbs, 1), "host_link_ceiling_gbs": round(ceiling_gbs, 1),
               "link_frac": round(moved_gbs / ceiling_gbs, 3),
               "ceiling_how": "custen_link_probe: cudaMemcpyAsync H2D and D2H of the same pinned buffers at once, every rank at "
                              "the same time, max over ranks",
               "host_buffers": "mmap + mbind to the GPU's NUMA node + cudaHostRegister" if near else "cudaHostAlloc",
               "ranks": nodes,
               "path": "custenCompute2D%s(HOST) on pinned host buffers: staged tile pipeline" % variant}
        st.destroy()
        cs.device_synchronize()
        if near:
            lib.custen_host_free_near(h_in, nbytes)
            lib.custen_host_free_near(h_out, nbytes)
        else:
            lib.custen_host_free(h_in)
            lib.custen_host_free(h_out)

    # ---- BASELINE.json config 5 at N > 1: the Cahn-Hilliard step on y-slabs (every rank takes part) ---------------
    cahn_slabs = None
    if world > 1 and not args.no_cahn:
        try:
            cahn_slabs = cahn_on_slabs(cs, torch, dist, rank, world, local_rank)
        except Exception as ex:   # noqa: BLE001 - a bench line with the error is worth more than no line
            cahn_slabs = {"error": repr(ex)}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- rank 0 extras: per-variant table, other configs, cpu baseline -----------------------------------------
    torch.cuda.empty_cache()
    extras = {}
    if cahn_slabs is not None:
        extras["cahn_hilliard_4096"] = cahn_slabs
    if world == 1 and not args.no_table:
        extras["variants_16384"] = variant_table(cs, torch, 16384, 10, 3, peak)
        for wname in ("x_p_8192", "xy_np_16384_t4"):
            v2, n2, t2, d2 = WORKLOADS[wname]
            c2, k2 = stencil_args(v2, n2)
            a = torch.rand((n2, n2), device="cuda", dtype=torch.float64)
            b = torch.zeros_like(a)
            s2 = cs.Stencil2D(v2, n2, n2, b, a, torch.from_numpy(np.ascontiguousarray(c2)).cuda(), numTiles=t2, **k2)
            m2 = time_resident(cs, s2, 20, 3) / 20
            extras[wname] = {"gpoints_per_s": round(n2 * n2 / m2 / 1e6, 2),
                             "frac_of_peak": round(n2 * n2 * ALG_BYTES_PER_POINT / m2 / 1e6 / peak, 4), "path": s2.path}
            s2.destroy()
            del a, b

    if world == 1 and not args.no_table:
        # BASELINE.json config 3 as the reference's callers run it: unified memory, numTiles = 4, through the
        # prefetch pipeline; offload = DEVICE (tiles stay on the GPU) and offload = HOST (every tile is sent back to
        # the CPU after its sweep, so each step crosses the host link in both directions)
        try:
            v3, n3, t3, _ = WORKLOADS["xy_np_16384_t4"]
            c3, k3 = stencil_args(v3, n3)
            lib = cs.load()
            cnt = n3 * n3
            res3 = {}
            # policy 0 (the default) = nothing is prefetched when the grid is already where the call wants it (DEVICE),
            # and HOST sweeps the CPU-resident grid in place over the host link instead of migrating every tile both
            # ways; policy 1 = the reference's prefetch pipeline on every call.  Each policy gets fresh buffers: the
            # unified-memory driver throttles pages that bounced between CPU and GPU a moment ago.
            for pol, pname in ((0, "default"), (1, "reference_pipeline")):
                m_in, m_out, m_w = lib.custen_managed_alloc(cnt * 8), lib.custen_managed_alloc(cnt * 8), lib.custen_managed_alloc(9 * 8)
                # filled by the CPU, like the reference's example programs do (examples/src/2d_xy_np.cu:88-101): the
                # first DEVICE call then migrates the grid to the GPU
                h_in = np.ctypeslib.as_array((ctypes.c_double * cnt).from_address(m_in))
                h_in[:] = np.random.default_rng(5).uniform(-1.0, 1.0, cnt)
                np.ctypeslib.as_array((ctypes.c_double * cnt).from_address(m_out))[:] = 0.0
                np.ctypeslib.as_array((ctypes.c_double * 9).from_address(m_w))[:] = np.ascontiguousarray(c3).ravel()
                del h_in
                s3 = cs.Stencil2D(v3, n3, n3, m_out, m_in, m_w, numTiles=t3, **k3)
                cs.set_managed_policy(pol)
                rp = {}
                for name, off, reps in (("offload_DEVICE", cs.DEVICE, 10), ("offload_HOST", cs.HOST, 3)):
                    s3.compute(off)
                    s3.compute(off)
                    cs.device_synchronize()
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        s3.compute(off)
                    cs.device_synchronize()
                    dt3 = (time.perf_counter() - t0) / reps
                    rp[name] = {"gpoints_per_s": round(n3 * n3 / dt3 / 1e9, 2), "ms_per_step": round(dt3 * 1e3, 3),
                                "mode": s3.mode}
                per_way = (2 if pol == 1 else 1) * cnt * 8  # pipeline: in and out tiles both migrate, both ways
                rp["offload_HOST"]["host_link_gbs_each_way"] = round(per_way / (rp["offload_HOST"]["ms_per_step"] * 1e-3) / 1e9, 1)
                res3[pname] = rp
                cs.set_managed_policy(0)
                s3.destroy()
                cs.device_synchronize()
                for pm in (m_in, m_out, m_w):
                    lib.custen_managed_free(pm)
            res3["note"] = "cudaMallocManaged buffers, numTiles = 4 (wall clock around Compute + device sync)"
            extras["xy_np_16384_t4_unified_memory"] = res3
        except Exception as ex:
            extras["xy_np_16384_t4_unified_memory"] = {"error": str(ex)}

        # BASELINE.json config 5 (and the 512^2 size of config 1): Cahn-Hilliard ADI steps on the re-hosted solver
        from custen_b200.cahn import CahnHilliard
        for ncahn in (512, 4096):
            res = {}
            # default: fused right-hand-side pass + partitioned tolerance-mode solve (within 1e-13 of the reference);
            # bit_identical: the same pass with the TMA-fed solve in the reference's operation order; engine path:
            # findCBar, cuStenCompute2DXYp / XYpFun and findRHS as separate passes (the reference driver's structure)
            # with the cp.async ring solve
            for key, fused, solver in (("ms_per_step", 1, 2), ("bit_identical_ms_per_step", 1, 0),
                                       ("engine_path_ms_per_step", 0, 1)):
                sol = CahnHilliard(ncahn, device=local_rank, solver=solver, fused=fused)
                sol.set_field(np.random.default_rng(0).uniform(-0.1, 0.1, (ncahn, ncahn)))
                sol.step(3)
                res[key] = round(sol.time_steps(20), 4)
                if key == "ms_per_step":
                    res["solver"] = sol.solver
                sol.destroy()
            res["mpoint_steps_per_s"] = round(ncahn * ncahn / res["ms_per_step"] / 1e3, 1)
            fields = []
            for solver in (2, 0):
                sol = CahnHilliard(ncahn, device=local_rank, solver=solver)
                sol.set_field(np.random.default_rng(0).uniform(-0.1, 0.1, (ncahn, ncahn)))
                sol.step(4)
                fields.append(sol.field())
                sol.destroy()
            res["parity"] = {"steps_compared": 4, "rel_vs_bit_identical_road": float(
                np.max(np.abs(fields[0] - fields[1])) / np.max(np.abs(fields[1])))}
            res["note"] = ("custen_cahn_step: right-hand side (2 stencils) + 2 cyclic pentadiagonal ADI solves per step; "
                           "solver 2 = partitioned solve (per step within 1e-13 of the reference GPU solver up to n = 1024 "
                           "and within the reference solver's own rounding error beyond); the other two roads are "
                           "bit-identical to the reference (tests/test_cahn_gpu.py)")
            extras[f"cahn_hilliard_{ncahn}"] = res

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            ncpu, reps = 4096, 0
            t = 0.0
            while t < 10.0 and reps < 64:
                t += cpu_sweeps(ncpu, 1, 1)
                reps += 1
            cpu = {"value": ncpu * ncpu * reps / t / 1e9, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"{reps} sweeps of a {ncpu}^2 periodic field by the reference's serial nonlinearRHS "
                             f"(serialCahnADI.c:553-622), {t:.1f} s on 1 of {len(os.sched_getaffinity(0))} host threads"}
        except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
        # BASELINE.json config 1: the reference's serial CPU Cahn-Hilliard program, 512^2, as shipped (T = 10)
        exe = os.path.join(ROOT, "oracle", "_ref", "serialCahnADI")
        if os.path.exists(exe) and not args.no_serial:
            try:
                r = subprocess.run([exe, "512"], capture_output=True, text=True, timeout=240)
                secs = float(r.stdout.split()[0])
                nsteps = 0
                t, dt = 0.0, 0.1 * (16.0 * np.pi / 512)
                while t < 10.0:  # the program's own loop condition (serialCahnADI.c:1010)
                    t += dt
                    nsteps += 1
                extras["config1_serial_cpu_cahn_512"] = {
                    "seconds": secs, "steps": nsteps, "mpoint_steps_per_s": round(512 * 512 * nsteps / secs / 1e6, 2),
                    "cores": 1, "host_threads_available": len(os.sched_getaffinity(0)),
                    "note": "oracle/_ref/serialCahnADI 512 = cuPentSpeedUp/serialCahnADITiming/serialCahnADI.c built as "
                            "compile.sh:18 minus the unused HDF5 flags; its own clock() print"}
            except Exception as ex:
                extras["config1_serial_cpu_cahn_512"] = {"error": str(ex)}

    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload if world == 1 else "", None)
        traffic_src = "stored_from_profiles (ncu --set full capture of this kernel, profiles/r2_ncu_full_summary.md); not measured in this run"
    achieved = value / world * ALG_BYTES_PER_POINT  # GB/s per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "grid": [n, n], "rows_per_gpu": rows,
                   "numTiles": tiles, "step": "one time step = cuSten Compute + Swap on every slab (custen_slab_run)",
                   "parallelism": f"y-slabs x{world}" + (f", halo transport: {args.transport}" if world > 1 else ""),
                   "l2": "no flush: every sweep streams 2 x %.1f GiB per GPU, far larger than the 126 MB L2" % (rows * n * 8 / 2 ** 30),
                   "kernel_family": path},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches_timed),
        "parity": parity,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "note": "algorithmic 16 B/point x points per launch / mean launch duration (CUDA events on the "
                             "launching stream over the timed region); per GPU"},
        "cpu_baseline": cpu,
    }
    if halo:
        line["halo_exchange"] = halo
    line.update(extras)
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    if parity and (parity["bits_differing"] or parity["neighbour_wait_timeouts"]):
        sys.stderr.write("PARITY FAILURE: %s\n" % json.dumps(parity))
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="xy_p_fun_32768", choices=sorted(WORKLOADS))
    ap.add_argument("--transport", default="peer", choices=["exchange", "peer"],
                    help="halo transport at N > 1: 'peer' reads the neighbours' edge rows in place over NVLink (IPC), "
                         "'exchange' sends them with NCCL every step")
    ap.add_argument("--e2e-tiles", type=int, default=32)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the seam rows after the timed loop")
    ap.add_argument("--no-table", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cahn", action="store_true", help="skip the Cahn-Hilliard (config 5) measurement at N > 1")
    ap.add_argument("--no-serial", action="store_true", help="skip the ~30 s serial CPU Cahn-Hilliard baseline (config 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
