"""ctypes access to the oracle libraries under oracle/_ref/ — test infrastructure only.

  libcusten_oracle.so  CPU restatement of the reference's stencil semantics (oracle/custen_oracle.c)
  libserialcahn.so     the reference's serial CPU twin, compiled from /root/reference (CPU pin of the oracle)
  libcusten_ref.so     the reference's CUDA library rebuilt for sm_100 + oracle/ref_shim.cu (GPU pin)
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

DIRS = {"X": 0, "Y": 1, "XY": 2}
FUNS = {None: 0, "second_diff_x": 1, "weighted9_x": 2, "weighted9_y": 3, "weighted3_y": 4, "weighted_xy": 5,
        "cubic_xy": 6}
VARIANT_IDS = {v: i for i, v in enumerate(
    ("Xp", "Xnp", "XpFun", "XnpFun", "Yp", "Ynp", "YpFun", "YnpFun", "XYp", "XYnp", "XYpFun", "XYnpFun"))}

_dp = ctypes.POINTER(ctypes.c_double)


def _ensure_built(name):
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=False, capture_output=True)
    return path if os.path.exists(path) else None


def variant_parts(variant):
    d = "XY" if variant.startswith("XY") else variant[0]
    rest = variant[len(d):]
    return d, rest.startswith("p"), rest.endswith("Fun")


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = _ensure_built("libcusten_oracle.so")
        if path is None:
            raise RuntimeError("oracle/_ref/libcusten_oracle.so missing and could not be built (make -C oracle)")
        lib = ctypes.CDLL(path)
        lib.custen_oracle_sweep.argtypes = [ctypes.c_int] * 3 + [_dp, _dp, ctypes.c_int, ctypes.c_int, _dp] + [ctypes.c_int] * 6
        lib.custen_oracle_sweep.restype = ctypes.c_int
        _oracle = lib
    return _oracle


def oracle_sweep(variant, inp, out, coef, H=1, L=0, R=0, V=1, T=0, B=0, fun=None, periodic_bits=None):
    """CPU oracle for one sweep.  `out` is modified in place (pre-fill it) and returned."""
    d, periodic, is_fun = variant_parts(variant)
    assert is_fun == (fun is not None)
    inp = np.ascontiguousarray(inp, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    assert out.dtype == np.float64 and out.flags.c_contiguous
    ny, nx = inp.shape
    bits = (3 if periodic else 0) if periodic_bits is None else periodic_bits
    rc = oracle().custen_oracle_sweep(DIRS[d], bits, FUNS[fun], inp.ctypes.data_as(_dp), out.ctypes.data_as(_dp), nx, ny,
                                      coef.ctypes.data_as(_dp), H, L, R, V, T, B)
    assert rc == 0
    return out


def oracle_time_steps(variant, in0, out0, coef, steps, **kw):
    """`steps` x (Compute + Swap) on a whole grid, the way a cuSten caller time-steps (the output buffer of one step is
    the input of the next and vice versa; regions a variant does not write keep what the buffer held).  Returns
    (latest result, the other buffer)."""
    a = np.ascontiguousarray(in0, dtype=np.float64).copy()
    b = np.ascontiguousarray(out0, dtype=np.float64).copy()
    for _ in range(steps):
        oracle_sweep(variant, a, b, coef, **kw)
        a, b = b, a
    return a, b


def oracle_evolve_band(variant, band0, coef, steps, T, B, **kw):
    """Time-step a horizontal band of a y-periodic grid without its neighbours: x wraps, y does not, so every step
    eats T rows at the top and B at the bottom.  Returns (band after `steps` steps, first valid row, end valid row);
    rows outside [first, end) are garbage.  Used to check a few rows of a grid too large for a whole-grid oracle."""
    d, periodic, is_fun = variant_parts(variant)
    assert periodic
    a = np.ascontiguousarray(band0, dtype=np.float64).copy()
    b = np.zeros_like(a)
    for _ in range(steps):
        oracle_sweep(variant, a, b, coef, T=T, B=B, periodic_bits=1, **kw)
        a, b = b, a
    return a, steps * T, a.shape[0] - steps * B


_serial = None


def serial():
    """The reference's serial CPU twin as a library, or None when it is not available (no /root/reference)."""
    global _serial
    if _serial is None:
        path = _ensure_built("libserialcahn.so")
        if path is None:
            return None
        lib = ctypes.CDLL(path)
        for name in ("linearRHS", "nonlinearRHS"):
            f = getattr(lib, name)
            f.argtypes = [_dp, _dp, _dp] + [ctypes.c_int] * 5
            f.restype = None
        _serial = lib
    return _serial


_ref = None


def ref_gpu():
    """The reference's CUDA kernels rebuilt for sm_100 (needs a GPU to call)."""
    global _ref
    if _ref is None:
        path = _ensure_built("libcusten_ref.so")
        if path is None:
            return None
        lib = ctypes.CDLL(path)
        lib.ref_sweep.argtypes = ([ctypes.c_int, _dp, _dp] + [ctypes.c_int] * 5 + [_dp] + [ctypes.c_int] * 7
                                  + [ctypes.c_char_p, ctypes.c_int])
        lib.ref_sweep.restype = ctypes.c_int
        lib.ref_time.argtypes = ([ctypes.c_int] * 6 + [_dp] + [ctypes.c_int] * 7 + [ctypes.c_char_p, ctypes.c_int,
                                                                                   ctypes.c_int])
        lib.ref_time.restype = ctypes.c_double
        _ref = lib
    return _ref


def ref_sweep(variant, inp, out, coef, H=1, L=0, R=0, V=1, T=0, B=0, fun=None, tiles=1, block=(32, 32), offload=0):
    """Run the reference's own CUDA kernels.  Returns `out` (modified in place) or None if the reference has no
    working implementation of the variant (XpFun)."""
    lib = ref_gpu()
    if lib is None:
        import pytest
        pytest.skip("oracle/_ref/libcusten_ref.so not built (needs /root/reference at build time)")
    inp = np.ascontiguousarray(inp, dtype=np.float64)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    ny, nx = inp.shape
    rc = lib.ref_sweep(VARIANT_IDS[variant], inp.ctypes.data_as(_dp), out.ctypes.data_as(_dp), nx, ny, tiles, block[0],
                       block[1], coef.ctypes.data_as(_dp), coef.size, H, L, R, V, T, B,
                       fun.encode() if fun else None, offload)
    return out if rc == 0 else None


def ref_time(variant, nx, ny, coef, H=1, L=0, R=0, V=1, T=0, B=0, fun=None, tiles=1, block=(32, 32), warmup=3, iters=10):
    lib = ref_gpu()
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    return lib.ref_time(VARIANT_IDS[variant], nx, ny, tiles, block[0], block[1], coef.ctypes.data_as(_dp), coef.size,
                        H, L, R, V, T, B, fun.encode() if fun else None, warmup, iters)


def oracle_weno(inp, u, v, dx, dy):
    lib = oracle()
    lib.custen_oracle_weno.argtypes = [_dp] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double]
    lib.custen_oracle_weno.restype = ctypes.c_int
    inp, u, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (inp, u, v))
    out = np.zeros_like(inp)
    ny, nx = inp.shape
    lib.custen_oracle_weno(inp.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp), out.ctypes.data_as(_dp),
                           nx, ny, dx, dy)
    return out


def ref_weno(inp, u, v, dx, dy, tiles=1, block=(32, 32), out_init=None):
    """The reference's WENO kernel (sm_100 rebuild)."""
    lib = ref_gpu()
    if lib is None:
        import pytest
        pytest.skip("oracle/_ref/libcusten_ref.so not built (needs /root/reference at build time)")
    lib.ref_weno.argtypes = [_dp] * 4 + [ctypes.c_int] * 5 + [ctypes.c_double, ctypes.c_double]
    lib.ref_weno.restype = ctypes.c_int
    inp, u, v = (np.ascontiguousarray(a, dtype=np.float64) for a in (inp, u, v))
    out = np.zeros_like(inp) if out_init is None else out_init.copy()
    ny, nx = inp.shape
    lib.ref_weno(inp.ctypes.data_as(_dp), u.ctypes.data_as(_dp), v.ctypes.data_as(_dp), out.ctypes.data_as(_dp), nx, ny,
                 tiles, block[0], block[1], dx, dy)
    return out




_cahn_libs = {}


def ref_cahn_run(c0, nsteps, lx, warm=0, engine="reference"):
    """The reference's GPU Cahn-Hilliard solver (timing twin + BatchHyper + cuPentBatch, oracle/ref_cahn_shim.cu):
    returns (final field after warm + nsteps steps, ms per timed step).
    engine = "reference": linked against the reference's own cuSten library (libcahn_ref.so, the oracle);
    engine = "new": the SAME translation units linked against this repo's libcuSten.a (libcahn_new.so) - the
    zero-source-change drop-in proof for config 5."""
    if engine not in _cahn_libs:
        path = _ensure_built("libcahn_ref.so" if engine == "reference" else "libcahn_new.so")
        if path is None:
            return None
        lib = ctypes.CDLL(path)
        lib.ref_cahn_run2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, _dp, _dp]
        lib.ref_cahn_run2.restype = ctypes.c_int
        _cahn_libs[engine] = lib
    lib = _cahn_libs[engine]
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    n = c0.shape[0]
    out = np.empty_like(c0)
    ms = ctypes.c_double(0.0)
    rc = lib.ref_cahn_run2(n, warm, nsteps, lx, c0.ctypes.data_as(_dp), out.ctypes.data_as(_dp), ctypes.byref(ms))
    assert rc == 0
    return out, ms.value


def serial_cahn_trajectory(c0, stops, lx, D=1.0, gamma=0.01):
    """serial_cahn_run, returning a copy of the field after each of the (increasing) step counts in `stops`."""
    return serial_cahn_run(c0, stops[-1], lx, D, gamma, stops=list(stops))


def serial_cahn_run(c0, nsteps, lx, D=1.0, gamma=0.01, stops=None):
    """The reference's serial CPU twin, driven function by function in the order of its own main loop
    (serialCahnADI.c:1010-1047) from a caller-supplied field.  Returns the final field, or None if unavailable."""
    lib = serial()
    if lib is None:
        return None
    c0 = np.ascontiguousarray(c0, dtype=np.float64)
    n = c0.shape[0]
    m = n - 2
    dx = lx / n
    dt = 0.1 * dx
    P = lambda a: a.ctypes.data_as(_dp)  # noqa: E731
    dbl, cint = ctypes.c_double, ctypes.c_int
    sig_l = 2.0 * dt * D * gamma / (3.0 * (dx * dx * dx * dx))
    a, b, c, d, e = sig_l, -4 * sig_l, 1 + 6 * sig_l, -4 * sig_l, sig_l
    ds, dl, diag, du, dw = (np.zeros(m) for _ in range(5))
    lib.setLHS.argtypes = [_dp] * 5 + [dbl] * 5 + [cint]
    lib.pentFactor.argtypes = [_dp] * 5 + [cint]
    lib.findOmega.argtypes = [_dp] * 3 + [dbl] * 5 + [cint]
    lib.findCBar.argtypes = [_dp] * 3 + [cint]
    lib.findRHS.argtypes = [_dp] * 4 + [cint]
    lib.cyclicInv.argtypes = [_dp] * 9 + [dbl] * 4 + [cint, cint]
    lib.transpose.argtypes = [_dp, _dp, cint]
    lib.findNew.argtypes = [_dp] * 3 + [cint]
    for f in (lib.setLHS, lib.pentFactor, lib.findOmega, lib.findCBar, lib.findRHS, lib.cyclicInv, lib.transpose, lib.findNew):
        f.restype = None
    lib.setLHS(P(ds), P(dl), P(diag), P(du), P(dw), a, b, c, d, e, m)
    lib.pentFactor(P(ds), P(dl), P(diag), P(du), P(dw), m)
    omega, inv1, inv2 = np.zeros(4), np.zeros(m), np.zeros(m)
    lib.findOmega(P(omega), P(inv1), P(inv2), a, b, c, d, e, m)
    w_lin = np.array([0, 0, -1, 0, 0, 0, -2, 8, -2, 0, -1, 8, -20, 8, -1, 0, -2, 8, -2, 0, 0, 0, -1, 0, 0], dtype=np.float64) * sig_l
    sig_n = (dt / 3.0) * D * (2.0 / (dx * dx))
    w_non = np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=np.float64) * sig_n
    c_old, c_cur = c0.copy(), c0.copy()
    c_bar, c_half, c_non = np.zeros_like(c0), np.zeros_like(c0), np.zeros_like(c0)
    kept = []
    for it in range(nsteps):
        lib.findCBar(P(c_old), P(c_cur), P(c_bar), n)
        lib.linearRHS(P(c_bar), P(c_half), P(w_lin), 5, 5, 2, 2, n)
        lib.nonlinearRHS(P(c_cur), P(c_non), P(w_non), 3, 3, 1, 1, n)
        lib.findRHS(P(c_old), P(c_cur), P(c_half), P(c_non), n)
        for i in range(n):
            row = ctypes.cast(c_half.ctypes.data + i * n * 8, _dp)
            lib.cyclicInv(P(ds), P(dl), P(diag), P(du), P(dw), P(inv1), P(inv2), P(omega), row, a, b, d, e, m, n)
        lib.transpose(P(c_half), P(c_cur), n)
        for i in range(n):
            row = ctypes.cast(c_cur.ctypes.data + i * n * 8, _dp)
            lib.cyclicInv(P(ds), P(dl), P(diag), P(du), P(dw), P(inv1), P(inv2), P(omega), row, a, b, d, e, m, n)
        lib.transpose(P(c_cur), P(c_half), n)
        lib.findNew(P(c_cur), P(c_bar), P(c_half), n)
        if stops is not None and it + 1 in stops:
            kept.append(c_cur.copy())
    return c_cur if stops is None else kept


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.int64), np.ascontiguousarray(b).view(np.int64))


def count_diff(a, b):
    return int(np.count_nonzero(np.ascontiguousarray(a).view(np.int64) != np.ascontiguousarray(b).view(np.int64)))
