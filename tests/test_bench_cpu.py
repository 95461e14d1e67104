"""bench.py's own host logic, checked without a GPU: the counter-based synthetic field, and the seam-row parity
checker (it must accept the oracle's own whole-grid time stepping and notice a single wrong double)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import cases  # noqa: E402
import oracle_lib as ol  # noqa: E402


def test_hash_rows_twins_agree_and_rows_are_independent_of_the_slab_cut():
    a = bench.hash_rows(0, 16, 32, bench.SEED, -0.1, 0.1)
    b = np.vstack([bench.hash_rows(r, 4, 32, bench.SEED, -0.1, 0.1) for r in (0, 4, 8, 12)])
    assert ol.count_diff(a, b) == 0
    assert ol.count_diff(a, cases.hash_field(0, 16, 32, bench.SEED, -0.1, 0.1)) == 0
    assert a.min() >= -0.1 and a.max() < 0.1 and np.unique(a).size == a.size


def test_seam_parity_accepts_the_whole_grid_result_and_sees_one_flipped_bit():
    n, steps, world = 64, 7, 4
    for variant in ("XYpFun", "XYp"):
        coef, kw = bench.stencil_args(variant, n, time_stepping=True)
        full0 = bench.hash_rows(0, n, n, bench.SEED, -0.1, 0.1)
        okw = {k: v for k, v in kw.items() if k != "fun"}
        final, _ = ol.oracle_time_steps(variant, full0, np.zeros_like(full0), coef, steps, fun=kw["fun"], **okw)
        assert np.isfinite(final).all() and np.abs(final).max() < 1.0
        rows = n // world
        for rank in range(world):
            lo, hi = rank * rows, (rank + 1) * rows
            checked, diff = bench.seam_parity(variant, coef, kw, n, steps, lo, hi, lambda gr: final[gr])
            assert (checked, diff) == (4, 0)
        bad = final.copy()
        bad[rows, 5] = np.nextafter(bad[rows, 5], 1.0)
        checked, diff = bench.seam_parity(variant, coef, kw, n, steps, rows, 2 * rows, lambda gr: bad[gr])
        assert diff == 1


def test_time_stepping_workload_stays_bounded():
    # 23 applications (3 warm-up + 20 timed steps) of the bench's time-stepping map keep the field O(0.1)
    n = 96
    coef, kw = bench.stencil_args("XYpFun", n, time_stepping=True)
    f = bench.hash_rows(0, n, n, bench.SEED, -0.1, 0.1)
    okw = {k: v for k, v in kw.items() if k != "fun"}
    final, _ = ol.oracle_time_steps("XYpFun", f, np.zeros_like(f), coef, 23, fun="cubic_xy", **okw)
    assert 1e-3 < np.abs(final).max() < 0.2


def test_clock_sampler_summarises_only_the_samples_of_the_requested_window():
    class FakeProc:
        def terminate(self):
            pass

    class FakeThread:
        def join(self, timeout=None):
            pass

    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.proc, s.th = FakeProc(), FakeThread()
    idle = ["210", "1965", "140.0", "Not Active", "Not Active", "Not Active", "Not Active"]
    load = ["1905", "1965", "950.0", "Not Active", "Not Active", "Not Active", "Active"]
    s.rows = [(0.10, idle), (0.20, idle), (1.00, load), (1.05, load), (1.10, ["1935"] + load[1:]), (2.0, idle)]
    got = s.stop(0.95, 1.2)
    assert got["samples"] == 3 and got["sm_mhz"] == 1905.0 and got["sm_max_mhz"] == 1965.0
    assert got["reasons"] == ["sw_power_cap"]
    s.rows = s.rows[:2]
    assert s.stop(0.95, 1.2) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
