"""snapshot files and the coarsening analysis that reads them (reference: Print_Out in
cuPentCahnADI/src/cuPentCahnADI.cu:103-140 at the cadence of :592-601, statistics of cuPentCahnADI/plotting.py:41-77).
On the CPU the fields come from the reference's serial twin (oracle/_ref/libserialcahn.so) on a 64^2 grid; the writer used
here is the numpy twin of custen_cahn_write_snapshot (the GPU test compares the two byte for byte)."""
import importlib.util
import math
import os

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("cahn_analysis", os.path.join(ROOT, "examples", "cahn_analysis.py"))
ca = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ca)


def test_snapshot_round_trip_and_naming(tmp_path):
    f = np.random.default_rng(0).uniform(-1, 1, (24, 40))
    path = ca.write_snapshot(tmp_path, 1.25, f)
    assert os.path.basename(path) == "cahn_hilliard_1.2500000000.bin"   # "%0.10lf", cuPentCahnADI.cu:108-110
    t, g = ca.read_snapshot(path)
    assert t == 1.25 and g.shape == f.shape and np.array_equal(g, f)
    with open(path, "r+b") as fh:
        fh.truncate(100)
    with pytest.raises(ValueError):
        ca.read_snapshot(path)


def test_simpson_matches_scipy_and_integrates_cubics_exactly():
    from scipy.integrate import simpson as sp_simpson
    x = np.linspace(0.0, 2.0 * math.pi, 65)
    y = np.random.default_rng(1).uniform(-1, 1, (7, 65))
    assert np.allclose(ca.simpson(y, x), sp_simpson(y, x=x), rtol=1e-13)
    assert abs(ca.simpson(x ** 3, x) - (2.0 * math.pi) ** 4 / 4.0) < 1e-10
    xe = np.linspace(0.0, 1.0, 64)                      # even number of samples: trapezoid on the last interval
    assert abs(ca.simpson(xe ** 2, xe) - 1.0 / 3.0) < 1e-4


def test_statistics_of_known_fields():
    n = 64
    # c = +-1 everywhere (fully separated): <c^2> -> 1, s(t) diverges; c = 0.5: s = 1 / (1 - 0.25)
    s, _ = ca.statistics(np.full((n, n), 0.5) + 1e-9 * np.random.default_rng(2).standard_normal((n, n)))
    assert abs(s - 1.0 / 0.75) < 1e-6
    # a single mode (kx, ky) = (3, 4): 1 / k1 = 1 / 5
    x = 2.0 * math.pi * np.arange(n) / n
    c = np.cos(3 * x)[None, :] * np.cos(4 * x)[:, None]
    _, inv_k1 = ca.statistics(c)
    assert abs(inv_k1 - 0.2) < 1e-12


def test_analysis_of_a_64_squared_run(tmp_path):
    """The reference's loop on a 64^2 grid with its serial CPU twin: a snapshot every 100 steps and one at the end;
    the analysis orders them by time and sees the domains coarsen."""
    n, every, total = 64, 100, 450
    lx = 16.0 * math.pi
    dt = 0.1 * lx / n
    c = np.random.default_rng(3).uniform(-0.1, 0.1, (n, n))
    c_old, done, time = None, 0, 0.0
    if ol.serial() is None:
        pytest.skip("reference serial twin not built")
    fields = ol.serial_cahn_trajectory(c, [100, 200, 300, 400, 450], lx)
    for steps, f in zip([100, 200, 300, 400, 450], fields):
        time = 0.0
        for _ in range(steps):
            time += dt
        if steps % every == 0:
            ca.write_snapshot(tmp_path, time, f)
    ca.write_snapshot(tmp_path, time, fields[-1])      # the reference's final Print_Out
    rows = ca.analyse(tmp_path)
    assert len(rows) == 5
    t = [r[0] for r in rows]
    assert t == sorted(t) and abs(t[-1] - 450 * dt) < 1e-9
    s = [r[1] for r in rows]
    inv_k1 = [r[2] for r in rows]
    assert all(np.isfinite(s)) and all(v > 1.0 for v in s)
    assert s[-1] > s[0]                                 # phase separation proceeds: <c^2> grows
    assert inv_k1[-1] > inv_k1[0] > 0.0                 # ... and the dominant length scale with it
