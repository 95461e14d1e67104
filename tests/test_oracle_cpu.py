"""CPU tests of the oracle itself (no GPU): the restatement must agree with the reference's own CPU code and
with the golden vectors produced by the reference's CUDA kernels, and honour the documented write masks."""
import ctypes
import glob
import os

import numpy as np
import pytest

import cases
import oracle_lib as ol

_dp = ctypes.POINTER(ctypes.c_double)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def _serial_rhs(fn_name, inp, weights, V, H, L, T):
    lib = ol.serial()
    if lib is None:
        pytest.skip("reference serial twin not available (no /root/reference and no prebuilt oracle/_ref)")
    n = inp.shape[0]
    out = np.zeros_like(inp)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    getattr(lib, fn_name)(inp.ctypes.data_as(_dp), out.ctypes.data_as(_dp), w.ctypes.data_as(_dp), V, H, L, T, n)
    return out


@pytest.mark.parametrize("n", [32, 96])
def test_oracle_matches_reference_serial_linear_rhs(n):
    """XYp 5x5 vs serialCahnADI.c:478-547 (linearRHS).  That file is compiled without FMA, so: to rounding."""
    inp = cases.field("random", n, n, seed=n)
    w = cases.weights_biharmonic(0.0123)
    ref = _serial_rhs("linearRHS", inp, w, 5, 5, 2, 2)
    got = ol.oracle_sweep("XYp", inp, np.zeros_like(inp), w, H=5, L=2, R=2, V=5, T=2, B=2)
    assert _rel(got, ref) < 1e-13


@pytest.mark.parametrize("n", [32, 96])
def test_oracle_matches_reference_serial_nonlinear_rhs(n):
    """XYpFun with the c^3 - c function vs serialCahnADI.c:553-622 (nonlinearRHS)."""
    inp = cases.field("random", n, n, seed=n + 1) * 0.1
    w = cases.weights_laplace5(0.731)
    ref = _serial_rhs("nonlinearRHS", inp, w, 3, 3, 1, 1)
    got = ol.oracle_sweep("XYpFun", inp, np.zeros_like(inp), w, H=3, L=1, R=1, V=3, T=1, B=1, fun="cubic_xy")
    assert _rel(got, ref) < 1e-13


def test_weights_and_weighted_fun_agree_bitwise():
    """A generic weighted user function is the weights variant by another road: identical FMA chain."""
    inp = cases.field("random", 64, 48)
    w = np.random.default_rng(3).uniform(-1, 1, 9)
    a = ol.oracle_sweep("XYp", inp, np.zeros_like(inp), w, H=3, L=1, R=1, V=3, T=1, B=1)
    b = ol.oracle_sweep("XYpFun", inp, np.zeros_like(inp), w, H=3, L=1, R=1, V=3, T=1, B=1, fun="weighted_xy")
    assert ol.bits_equal(a, b)
    w9 = np.random.default_rng(4).uniform(-1, 1, 9)
    a = ol.oracle_sweep("Xp", inp, np.zeros_like(inp), w9, H=9, L=4, R=4)
    b = ol.oracle_sweep("XpFun", inp, np.zeros_like(inp), w9, H=9, L=4, R=4, fun="weighted9_x")
    assert ol.bits_equal(a, b)
    a = ol.oracle_sweep("Yp", inp, np.zeros_like(inp), w9, V=9, T=4, B=4)
    b = ol.oracle_sweep("YpFun", inp, np.zeros_like(inp), w9, V=9, T=4, B=4, fun="weighted9_y")
    assert ol.bits_equal(a, b)


def test_x_is_y_transposed():
    inp = cases.field("random", 40, 56)
    w = np.random.default_rng(5).uniform(-1, 1, 5)
    a = ol.oracle_sweep("Xp", inp, np.zeros_like(inp), w, H=5, L=2, R=2)
    b = ol.oracle_sweep("Yp", np.ascontiguousarray(inp.T), np.zeros((40, 56)), w, V=5, T=2, B=2)
    assert ol.bits_equal(a, np.ascontiguousarray(b.T))


def test_analytic_answers_of_the_examples():
    """The example programs' own sanity check: d2/dx2 sin = -sin etc., to truncation error."""
    n = 256
    h = 2 * np.pi / n
    f = cases.field("sinx", n, 8)
    got = ol.oracle_sweep("Xp", f, np.zeros_like(f), cases.weights_d2_8th(h), H=9, L=4, R=4)
    assert np.max(np.abs(got + f)) < 1e-9          # examples/src/2d_x_p.cu:87
    f = cases.field("siny", 8, n)
    got = ol.oracle_sweep("Yp", f, np.zeros_like(f), cases.weights_d2_2nd(h), V=3, T=1, B=1)
    assert np.max(np.abs(got + f)) < 1e-3          # examples/src/2d_y_p.cu (2nd order)
    f = cases.field("sinxcosy", n, n)
    x = np.arange(n) * h
    ans = -np.cos(x)[None, :] * np.sin(x)[:, None]  # examples/src/2d_xy_p.cu:88
    got = ol.oracle_sweep("XYp", f, np.zeros_like(f), cases.weights_cross_xy(h, h), H=3, L=1, R=1, V=3, T=1, B=1)
    assert np.max(np.abs(got - ans)) < 1e-3


def test_non_periodic_write_masks():
    S = cases.SENTINEL
    inp = cases.field("random", 32, 24)
    w = np.ones(5)
    out = ol.oracle_sweep("Xnp", inp, np.full_like(inp, S), w, H=5, L=2, R=2)
    assert np.all(out[:, :2] == S) and np.all(out[:, -2:] == 0.0) and not np.any(out[:, 2:-2] == S)
    out = ol.oracle_sweep("XnpFun", inp, np.full_like(inp, S), [1.0], H=3, L=1, R=1, fun="second_diff_x")
    assert np.all(out[:, :1] == S) and np.all(out[:, -1:] == S) and not np.any(out[:, 1:-1] == S)
    out = ol.oracle_sweep("Ynp", inp, np.full_like(inp, S), w, V=5, T=2, B=2)
    assert np.all(out[:2] == S) and np.all(out[-2:] == S) and not np.any(out[2:-2] == S)
    w9 = np.ones(9)
    out = ol.oracle_sweep("XYnp", inp, np.full_like(inp, S), w9, H=3, L=1, R=1, V=3, T=1, B=1)
    assert np.all(out[0] == S) and np.all(out[-1] == S) and np.all(out[:, 0] == S) and np.all(out[:, -1] == S)
    assert not np.any(out[1:-1, 1:-1] == S)
    out = ol.oracle_sweep("XYp", inp, np.full_like(inp, S), w9, H=3, L=1, R=1, V=3, T=1, B=1)
    assert not np.any(out == S)


def test_periodic_wrap_is_a_roll():
    """Periodic variants commute with cyclic shifts of the grid (size-independent property)."""
    inp = cases.field("random", 48, 40)
    w = np.random.default_rng(11).uniform(-1, 1, 25)
    a = ol.oracle_sweep("XYp", inp, np.zeros_like(inp), w, H=5, L=2, R=2, V=5, T=2, B=2)
    sh = np.roll(inp, (7, -5), axis=(0, 1))
    b = ol.oracle_sweep("XYp", np.ascontiguousarray(sh), np.zeros_like(inp), w, H=5, L=2, R=2, V=5, T=2, B=2)
    assert ol.bits_equal(np.roll(a, (7, -5), axis=(0, 1)), b)


def test_slab_mode_reassembles_the_global_sweep():
    """periodic bits = 1 (x wraps, y does not) on [halo; slab; halo] reproduces the global periodic sweep."""
    inp = cases.field("random", 32, 64)
    w = np.random.default_rng(12).uniform(-1, 1, 9)
    kw = dict(H=3, L=1, R=1, V=3, T=1, B=1)
    full = ol.oracle_sweep("XYp", inp, np.zeros_like(inp), w, **kw)
    parts = []
    for g in range(4):
        lo, hi = g * 16, (g + 1) * 16
        ext = np.ascontiguousarray(np.take(inp, range(lo - 1, hi + 1), axis=0, mode="wrap"))
        o = ol.oracle_sweep("XYp", ext, np.zeros_like(ext), w, periodic_bits=1, **kw)
        parts.append(o[1:-1])
    assert ol.bits_equal(full, np.vstack(parts))


GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("path", GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def test_oracle_reproduces_reference_gpu_golden_vectors(path):
    """tests/golden/*.npz hold outputs of the reference's CUDA kernels (sm_100 rebuild) on seeded inputs."""
    g = np.load(path)
    c = {k: (g[k].item() if g[k].shape == () else g[k]) for k in g.files}
    fun = str(c["fun"]) if str(c["fun"]) else None
    out = np.full_like(c["inp"], cases.SENTINEL)
    got = ol.oracle_sweep(str(c["variant"]), c["inp"], out, c["coef"], H=int(c["H"]), L=int(c["L"]), R=int(c["R"]),
                          V=int(c["V"]), T=int(c["T"]), B=int(c["B"]), fun=fun)
    assert ol.count_diff(got, c["out"]) == 0


def test_golden_vectors_exist():
    assert len(GOLDEN_FILES) >= 11, "golden vectors missing: run tests/golden/make_golden.py on a GPU box"
