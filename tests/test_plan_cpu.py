"""Host logic without a GPU: the plan's tiling, seam pointers and write masks (custen_debug_bands builds the plan
of a variant without touching CUDA and reports the bands Compute would launch).  Each band is 'executed' here by the
ORACLE on [top rows; band rows; bottom rows] and scattered through the band's masks; the pieces must reassemble the
oracle's global sweep bit for bit — for every variant, tile count, merged / per-tile launch, and y-slab position."""
import numpy as np
import pytest

import cases
import oracle_lib as ol
import custen_b200._lib as L

MARK = -777.25  # pre-fill of per-band scratch outputs: must never leak into the assembled result

SHAPES = {
    "X": dict(H=5, L=2, R=2), "Y": dict(V=5, T=2, B=2), "XY": dict(H=3, L=1, R=1, V=5, T=2, B=2),
}
FUNS = {"XpFun": ("weighted9_x", dict(H=9, L=4, R=4)), "XnpFun": ("weighted9_x", dict(H=9, L=4, R=4)),
        "YpFun": ("weighted9_y", dict(V=9, T=4, B=4)), "YnpFun": ("weighted9_y", dict(V=9, T=4, B=4)),
        "XYpFun": ("cubic_xy", dict(H=3, L=1, R=1, V=3, T=1, B=1)), "XYnpFun": ("cubic_xy", dict(H=3, L=1, R=1, V=3, T=1, B=1))}


def _kw(variant):
    if variant in FUNS:
        fun, kw = FUNS[variant]
        return dict(kw, fun=fun)
    d = ol.variant_parts(variant)[0]
    return dict(SHAPES[d], fun=None)


def _run_band(variant, b, inp_local, halo_top, halo_bottom, coef, kw, out_local):
    """Execute one band with the oracle and scatter it into out_local through the band's masks."""
    nx = b.nx
    r0 = b.in_off // nx
    assert b.in_off % nx == 0 and b.out_off == b.in_off
    parts, t0 = [], 0
    if b.top_kind == 1:
        parts.append(inp_local[b.top_off // nx: b.top_off // nx + b.T])
    elif b.top_kind == 2:
        parts.append(halo_top)
    if b.top_kind:
        t0 = b.T
    parts.append(inp_local[r0: r0 + b.rows])
    if b.bottom_kind == 1:
        parts.append(inp_local[b.bottom_off // nx: b.bottom_off // nx + b.B])
    elif b.bottom_kind == 2:
        parts.append(halo_bottom)
    ext = np.ascontiguousarray(np.vstack(parts))
    # raw values everywhere the window exists: x periodic iff the band wraps, y never wraps inside a band
    raw = np.full_like(ext, MARK)
    bits = 1 if b.wrap_x else 0
    okw = {k: v for k, v in kw.items() if k != "fun"}
    ol.oracle_sweep(variant, ext, raw, coef, fun=kw["fun"], periodic_bits=bits, **okw)
    band = raw[t0: t0 + b.rows]
    ys = slice(b.ylo, b.yhi)
    out_local[r0: r0 + b.rows][ys, b.xlo:b.xhi] = band[ys, b.xlo:b.xhi]
    if b.zero_right:
        out_local[r0: r0 + b.rows][ys, b.xhi:] = 0.0


@pytest.mark.parametrize("merged", [False, True])
@pytest.mark.parametrize("tiles", [1, 2, 3, 4])
@pytest.mark.parametrize("variant", L.VARIANTS)
def test_bands_reassemble_the_global_sweep(variant, tiles, merged):
    nx, ny = 40, 48
    kw = _kw(variant)
    okw = {k: v for k, v in kw.items() if k != "fun"}
    inp = cases.field("random", nx, ny, seed=tiles)
    coef = np.random.default_rng(1).uniform(-1, 1, max(9, okw.get("H", 1) * okw.get("V", 1)))
    want = ol.oracle_sweep(variant, inp, np.full_like(inp, cases.SENTINEL), coef, fun=kw["fun"], **okw)
    bands = L.debug_bands(variant, tiles, nx, ny, merged=merged, **okw)
    assert len(bands) == (1 if merged else tiles)
    assert all(b.contiguous for b in bands)
    got = np.full_like(inp, cases.SENTINEL)
    for b in bands:
        _run_band(variant, b, inp, None, None, coef, kw, got)
    assert not np.any(got == MARK)
    assert ol.count_diff(got, want) == 0


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("tiles", [1, 2])
@pytest.mark.parametrize("variant", ["Yp", "Ynp", "XYp", "XYnp", "XYpFun", "XYnpFun", "YnpFun"])
def test_slab_bands_reassemble_the_global_sweep(variant, tiles, world):
    nx, ny = 24, 32 * world
    kw = _kw(variant)
    okw = {k: v for k, v in kw.items() if k != "fun"}
    periodic = ol.variant_parts(variant)[1]
    inp = cases.field("random", nx, ny, seed=world)
    coef = np.random.default_rng(2).uniform(-1, 1, max(9, okw.get("H", 1) * okw.get("V", 1)))
    want = ol.oracle_sweep(variant, inp, np.full_like(inp, cases.SENTINEL), coef, fun=kw["fun"], **okw)
    got = np.full_like(inp, cases.SENTINEL)
    rows = ny // world
    T, B = okw.get("T", 0), okw.get("B", 0)
    for g in range(world):
        lo, hi = g * rows, (g + 1) * rows
        local_in = np.ascontiguousarray(inp[lo:hi])
        local_out = got[lo:hi]
        halo_top = np.take(inp, range(lo - T, lo), axis=0, mode="wrap")
        halo_bot = np.take(inp, range(hi, hi + B), axis=0, mode="wrap")
        bands = L.debug_bands(variant, tiles, nx, rows, slab=(g == 0, g == world - 1), **okw)
        for b in bands:
            if not periodic:  # a slab at the physical edge must not look at a halo
                assert not (g == 0 and b.in_off == 0 and b.top_kind)
                assert not (g == world - 1 and b.in_off // nx + b.rows == rows and b.bottom_kind)
            _run_band(variant, b, local_in, halo_top, halo_bot, coef, kw, local_out)
    assert not np.any(got == MARK)
    assert ol.count_diff(got, want) == 0


def test_public_fields_and_seams_follow_the_reference_formulas():
    """Seam offsets as custenCreateDestroy2DXYp.cu:194-228 computes them (3 tiles of 16 rows, T = B = 2)."""
    nx, ny, T = 40, 48, 2
    bands = L.debug_bands("XYp", 3, nx, ny, H=5, L=2, R=2, V=5, T=2, B=2)
    assert [b.in_off for b in bands] == [0, 16 * nx, 32 * nx]
    assert [b.top_off for b in bands] == [(ny - T) * nx, (16 - T) * nx, (32 - T) * nx]
    assert [b.bottom_off for b in bands] == [16 * nx, 32 * nx, 0]
    assert all(b.wrap_x and b.top_kind == 1 and b.bottom_kind == 1 for b in bands)
    np_bands = L.debug_bands("XYnp", 3, nx, ny, H=5, L=2, R=2, V=5, T=2, B=2)
    assert [(b.ylo, b.yhi) for b in np_bands] == [(2, 16), (0, 16), (0, 14)]
    assert all((b.xlo, b.xhi, b.wrap_x, b.zero_right) == (2, nx - 2, 0, 0) for b in np_bands)
    xnp = L.debug_bands("Xnp", 2, nx, ny, H=5, L=2, R=2)
    assert all(b.zero_right == 1 and b.top_kind == 0 and b.bottom_kind == 0 for b in xnp)
