"""The C slab layer (custen_b200/csrc/slab.cu, custen_mg_*) time-stepping a global grid cut into y-slabs, against the
oracle's whole-grid time stepping, bit for bit.

Runs on ONE GPU too: several slabs may share a device (they then share a stream), which exercises everything the
multi-GPU layer does on the host and in the kernels - peer pointers into the neighbours' buffers, Swap re-aliasing of
both buffers and seams, the per-slab sweep counters, the guard-row bookkeeping, the fallback road - except the actual
spinning.  With two or more GPUs the same cases also run one slab per device (true concurrency, kernels waiting on
each other over NVLink) and through the graph replay.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle_lib as ol  # noqa: E402
import custen_b200 as cs  # noqa: E402

NGPU = torch.cuda.device_count()

CASES = [
    ("XYp", dict(H=3, L=1, R=1, V=3, T=1, B=1), None),
    ("XYp", dict(H=5, L=2, R=2, V=5, T=2, B=2), None),
    ("XYnp", dict(H=5, L=2, R=2, V=5, T=2, B=2), None),
    ("XYnp", dict(H=3, L=1, R=1, V=3, T=1, B=1), None),
    ("Yp", dict(V=9, T=4, B=4), None),
    ("Ynp", dict(V=5, T=2, B=2), None),
    ("XYpFun", dict(H=3, L=1, R=1, V=3, T=1, B=1), "cubic_xy"),
    ("XYnpFun", dict(H=3, L=1, R=1, V=3, T=1, B=1), "weighted_xy"),
    ("YpFun", dict(V=9, T=4, B=4), "weighted9_y"),
    ("Xp", dict(H=9, L=4, R=4), None),
    ("Xnp", dict(H=5, L=2, R=2), None),
]


def _layouts():
    out = [[0], [0, 0], [0, 0, 0, 0]]
    if NGPU >= 2:
        out.append([0, 1])
        out.append(list(range(min(NGPU, 4))))
    return out


def _coef(variant, kw, fun):
    n = kw.get("H", 1) * kw.get("V", 1)
    c = np.random.default_rng(3).uniform(-1, 1, n)
    if fun == "cubic_xy":
        # h + eps Lap(h), h = c - c^3: stays bounded under iteration (explicit Allen-Cahn-like step)
        eps = 0.05
        c = np.array([0, -eps, 0, -eps, -1 + 4 * eps, -eps, 0, -eps, 0], dtype=np.float64)
    else:
        c = c / np.abs(c).sum()  # non-expansive: the iterates stay O(1)
    return c


def _run_mg(devices, variant, kw, fun, nx, ny, steps, plain_steps=0):
    lib = cs.load()
    coef = _coef(variant, kw, fun)
    full = cases.field("random", nx, ny)
    devs = (ctypes.c_int * len(devices))(*devices)
    mg = lib.custen_mg_create(len(devices), ctypes.addressof(devs), cs.VARIANTS.index(variant), nx, ny, coef.ctypes.data,
                              coef.size, kw.get("H", 1), kw.get("L", 0), kw.get("R", 0), kw.get("V", 1), kw.get("T", 0),
                              kw.get("B", 0), fun.encode() if fun else None, None)
    lib.custen_mg_scatter(mg, full.ctypes.data)
    lib.custen_mg_fill_output(mg, cases.SENTINEL)
    for _ in range(plain_steps):
        lib.custen_mg_compute(mg)
        lib.custen_mg_swap(mg)
    lib.custen_mg_run(mg, steps - plain_steps)
    lib.custen_mg_synchronize(mg)
    assert lib.custen_mg_error(mg) == 0, "a neighbour wait timed out"
    got, other = np.empty_like(full), np.empty_like(full)
    lib.custen_mg_gather(mg, got.ctypes.data, 0)
    lib.custen_mg_gather(mg, other.ctypes.data, 1)
    path = cs.api.PATH_NAMES[lib.custen_slab_last_path(lib.custen_mg_slab(mg, 0))]
    lib.custen_mg_destroy(mg)
    want, want_other = ol.oracle_time_steps(variant, full, np.full_like(full, cases.SENTINEL), coef, steps, fun=fun, **kw)
    return ol.count_diff(got, want), ol.count_diff(other, want_other), path


@pytest.mark.parametrize("devices", _layouts(), ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("variant,kw,fun", CASES, ids=lambda v: v if isinstance(v, str) else None)
def test_mg_time_stepping_matches_whole_grid_oracle(devices, variant, kw, fun):
    nx, ny, steps = 1024, 64 * len(devices) * (2 if len(devices) < 4 else 1), 5
    d0, d1, path = _run_mg(devices, variant, kw, fun, nx, ny, steps)
    assert (d0, d1) == (0, 0), f"{variant} on {devices}: {d0} / {d1} points differ from the oracle after {steps} steps"
    assert path.startswith("stream"), path


@pytest.mark.parametrize("devices", _layouts()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_mg_fallback_road_and_many_steps(devices):
    # odd nx: the plain-load kernel with the standalone wait / signal kernels around it
    d0, d1, path = _run_mg(devices, "XYp", dict(H=3, L=1, R=1, V=3, T=1, B=1), None, 515, 48 * len(devices), 4)
    assert (d0, d1, path) == (0, 0, "fallback")
    # enough steps for the graph replay (pairs) with an odd tail, after a plain first step
    d0, d1, path = _run_mg(devices, "XYpFun", dict(H=3, L=1, R=1, V=3, T=1, B=1), "cubic_xy", 2048, 96 * len(devices), 11,
                           plain_steps=2)
    assert (d0, d1) == (0, 0)


def test_mg_static_input_repeated_compute():
    # Compute without Swap (the reference's examples): the counters advance, nobody waits for long, same result each time
    lib = cs.load()
    devices = [0, 0] if NGPU < 2 else [0, 1]
    variant, kw = "XYp", dict(H=3, L=1, R=1, V=3, T=1, B=1)
    nx, ny = 512, 256
    coef = _coef(variant, kw, None)
    full = cases.field("random", nx, ny)
    devs = (ctypes.c_int * 2)(*devices)
    mg = lib.custen_mg_create(2, ctypes.addressof(devs), cs.VARIANTS.index(variant), nx, ny, coef.ctypes.data, 9, 3, 1, 1, 3, 1,
                              1, None, None)
    lib.custen_mg_scatter(mg, full.ctypes.data)
    for _ in range(3):
        lib.custen_mg_compute(mg)
    got = np.empty_like(full)
    lib.custen_mg_gather(mg, got.ctypes.data, 1)
    assert lib.custen_mg_error(mg) == 0
    lib.custen_mg_destroy(mg)
    want = ol.oracle_sweep(variant, full, np.zeros_like(full), coef, **kw)
    assert ol.count_diff(got, want) == 0


def test_fill_hash_matches_numpy_twin():
    lib = cs.load()
    rows, nx, row0 = 37, 256, 1000003
    t = torch.empty((rows, nx), device="cuda", dtype=torch.float64)
    lib.custen_fill_hash(t.data_ptr(), row0, rows, nx, 0x5EED, -0.1, 0.1)
    torch.cuda.synchronize()
    assert ol.count_diff(t.cpu().numpy(), cases.hash_field(row0, rows, nx, 0x5EED, -0.1, 0.1)) == 0
