"""Parity cases shared by the CPU tests, the GPU tests and the golden-vector generator.

Stencils and launch parameters are those of the reference's example programs, at
sizes an oracle sweep finishes in well under a second, on analytic fields (as the examples use) and on seeded
random fields.  Every case is valid for the reference (nx % BLOCK_X == 0, nyTile % BLOCK_Y == 0, at least two
blocks per direction, L,R <= BLOCK_X, T,B <= BLOCK_Y, XYnp with L == R).
"""
import numpy as np

SENTINEL = 12345.678  # pre-fill of `out`, so untouched regions are compared too


def weights_d2_8th(h):
    """9-point 8th-order second derivative (examples/src/2d_x_p.cu:99-114)."""
    s = 1.0 / (h * h)
    w = np.array([-1.0 / 560, 8.0 / 315, -1.0 / 5, 8.0 / 5, -205.0 / 72, 8.0 / 5, -1.0 / 5, 8.0 / 315, -1.0 / 560])
    return w * s


def weights_d2_4th(h):
    """5-point 4th-order second derivative (BASELINE.json config 2, literal reading)."""
    return np.array([-1.0 / 12, 4.0 / 3, -5.0 / 2, 4.0 / 3, -1.0 / 12]) / (h * h)


def weights_d2_2nd(h):
    """3-point second derivative (examples/src/2d_y_p.cu:99-108)."""
    return np.array([1.0, -2.0, 1.0]) / (h * h)


def weights_cross_xy(dx, dy):
    """3x3 cross derivative d2/dxdy (examples/src/2d_xy_p.cu:112-120)."""
    s = 1.0 / (4.0 * dx * dy)
    return np.array([s, 0.0, -s, 0.0, 0.0, 0.0, -s, 0.0, s])


def weights_biharmonic(sig):
    """5x5 13-point biharmonic times -sigma (cuPentCahnADI/src/cuPentCahnADI.cu:463-476)."""
    w = np.array([0, 0, -1, 0, 0,
                  0, -2, 8, -2, 0,
                  -1, 8, -20, 8, -1,
                  0, -2, 8, -2, 0,
                  0, 0, -1, 0, 0], dtype=np.float64)
    return w * sig


def weights_laplace5(sig):
    """3x3 five-point Laplacian times sigma (cuPentCahnADI/src/cuPentCahnADI.cu:511-513)."""
    return np.array([0, 1, 0, 1, -4, 1, 0, 1, 0], dtype=np.float64) * sig


def field(kind, nx, ny, seed=0x5EED):
    if kind == "random":
        return np.random.default_rng(seed).uniform(-1.0, 1.0, size=(ny, nx))
    x = np.arange(nx) * (2 * np.pi / nx)
    y = np.arange(ny) * (2 * np.pi / ny)
    if kind == "sinx":
        return np.tile(np.sin(x), (ny, 1))
    if kind == "siny":
        return np.tile(np.sin(y)[:, None], (1, nx))
    if kind == "sinxcosy":
        return np.sin(x)[None, :] * np.cos(y)[:, None]
    raise ValueError(kind)


def hash_field(row0, rows, nx, seed, lo, hi):
    """numpy twin of custen_fill_hash (custen_b200/csrc/api_c.cu): rows [row0, row0 + rows) of the synthetic field whose
    point (r, c) is lo + (hi - lo) * u(seed, r * nx + c), u = top 53 bits of splitmix64's finaliser / 2^53."""
    idx = (np.uint64(row0) * np.uint64(nx) + np.arange(rows * nx, dtype=np.uint64))
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + (hi - lo) * u).reshape(rows, nx)


def _c(name, variant, nx, ny, tiles, block, coef, H=1, L=0, R=0, V=1, T=0, B=0, fun=None, numCoe=None, fld="random"):
    return dict(name=name, variant=variant, nx=nx, ny=ny, tiles=tiles, block=block, coef=np.asarray(coef, float),
                H=H, L=L, R=R, V=V, T=T, B=B, fun=fun, numCoe=numCoe, field=fld)


def build_cases():
    cs = []
    dx = lambda n: 2 * np.pi / n  # noqa: E731
    rng = np.random.default_rng(7)
    # --- X ---
    cs.append(_c("x_p_9pt_example", "Xp", 2048, 1024, 2, (32, 32), weights_d2_8th(dx(2048)), H=9, L=4, R=4, fld="sinx"))
    cs.append(_c("x_p_9pt_random", "Xp", 1024, 256, 1, (32, 32), weights_d2_8th(dx(1024)), H=9, L=4, R=4))
    cs.append(_c("x_p_5pt", "Xp", 1024, 128, 1, (32, 32), weights_d2_4th(dx(1024)), H=5, L=2, R=2))
    cs.append(_c("x_p_3pt_oddL", "Xp", 512, 128, 2, (32, 16), weights_d2_2nd(dx(512)), H=3, L=1, R=1))
    cs.append(_c("x_p_7pt_tile_family", "Xp", 512, 64, 1, (32, 16), rng.uniform(-1, 1, 7), H=7, L=3, R=3))
    cs.append(_c("x_p_4pt_asym", "Xp", 256, 64, 1, (32, 16), rng.uniform(-1, 1, 4), H=4, L=1, R=2))
    cs.append(_c("x_np_9pt_example", "Xnp", 1024, 512, 1, (32, 16), weights_d2_8th(dx(1024)), H=9, L=4, R=4, fld="sinx"))
    cs.append(_c("x_np_5pt_tiles", "Xnp", 512, 256, 4, (32, 16), weights_d2_4th(dx(512)), H=5, L=2, R=2))
    cs.append(_c("x_np_fun_example", "XnpFun", 1024, 512, 4, (32, 16), [1.0 / dx(1024) ** 2], H=3, L=1, R=1,
                 fun="second_diff_x", numCoe=1, fld="sinx"))
    cs.append(_c("x_np_fun_9pt", "XnpFun", 512, 128, 1, (32, 16), weights_d2_8th(dx(512)), H=9, L=4, R=4,
                 fun="weighted9_x", numCoe=9))
    cs.append(_c("x_p_fun_3pt", "XpFun", 512, 128, 2, (32, 32), [1.0 / dx(512) ** 2], H=3, L=1, R=1,
                 fun="second_diff_x", numCoe=1))
    cs.append(_c("x_p_fun_9pt", "XpFun", 512, 128, 1, (32, 32), weights_d2_8th(dx(512)), H=9, L=4, R=4,
                 fun="weighted9_x", numCoe=9))
    # --- Y ---
    cs.append(_c("y_p_3pt_example", "Yp", 512, 512, 1, (8, 8), weights_d2_2nd(dx(512)), V=3, T=1, B=1, fld="siny"))
    cs.append(_c("y_p_9pt_tiles", "Yp", 256, 512, 4, (32, 16), weights_d2_8th(dx(512)), V=9, T=4, B=4))
    cs.append(_c("y_p_5pt", "Yp", 256, 256, 2, (32, 16), weights_d2_4th(dx(256)), V=5, T=2, B=2))
    cs.append(_c("y_np_9pt_example", "Ynp", 512, 512, 2, (8, 8), weights_d2_8th(dx(512)), V=9, T=4, B=4, fld="siny"))
    cs.append(_c("y_np_3pt", "Ynp", 256, 256, 1, (32, 16), weights_d2_2nd(dx(256)), V=3, T=1, B=1))
    cs.append(_c("y_p_fun_example", "YpFun", 64, 64, 4, (4, 4), weights_d2_8th(dx(64)), V=9, T=4, B=4,
                 fun="weighted9_y", numCoe=9, fld="siny"))
    cs.append(_c("y_p_fun_3pt", "YpFun", 512, 256, 1, (32, 16), weights_d2_2nd(dx(256)), V=3, T=1, B=1,
                 fun="weighted3_y", numCoe=3))
    cs.append(_c("y_np_fun_example", "YnpFun", 512, 512, 2, (8, 8), weights_d2_8th(dx(512)), V=9, T=4, B=4,
                 fun="weighted9_y", fld="siny"))
    # --- XY ---
    cs.append(_c("xy_p_cross_example", "XYp", 1024, 1024, 1, (32, 32), weights_cross_xy(dx(1024), dx(1024)),
                 H=3, L=1, R=1, V=3, T=1, B=1, fld="sinxcosy"))
    cs.append(_c("xy_p_cross_tiles", "XYp", 512, 512, 4, (32, 32), weights_cross_xy(dx(512), dx(512)),
                 H=3, L=1, R=1, V=3, T=1, B=1))
    cs.append(_c("xy_p_biharmonic", "XYp", 512, 512, 1, (32, 32), weights_biharmonic(0.37), H=5, L=2, R=2, V=5, T=2, B=2))
    cs.append(_c("xy_p_3x5_tile_family", "XYp", 256, 256, 2, (32, 16), rng.uniform(-1, 1, 15), H=3, L=1, R=1, V=5, T=2, B=2))
    cs.append(_c("xy_np_cross_example", "XYnp", 128, 128, 4, (4, 4), weights_cross_xy(dx(128), dx(128)),
                 H=3, L=1, R=1, V=3, T=1, B=1, fld="sinxcosy"))
    cs.append(_c("xy_np_5x5", "XYnp", 512, 256, 2, (32, 16), weights_biharmonic(0.11), H=5, L=2, R=2, V=5, T=2, B=2))
    cs.append(_c("xy_p_fun_example", "XYpFun", 1024, 1024, 1, (16, 32), weights_cross_xy(dx(1024), dx(1024)),
                 H=3, L=1, R=1, V=3, T=1, B=1, fun="weighted_xy", fld="sinxcosy"))
    cs.append(_c("xy_p_fun_cubic", "XYpFun", 512, 512, 2, (8, 8), weights_laplace5(0.21), H=3, L=1, R=1, V=3, T=1, B=1,
                 fun="cubic_xy"))
    cs.append(_c("xy_p_fun_5x5", "XYpFun", 256, 256, 1, (32, 16), weights_biharmonic(0.05), H=5, L=2, R=2, V=5, T=2, B=2,
                 fun="weighted_xy"))
    cs.append(_c("xy_np_fun_example", "XYnpFun", 128, 128, 1, (8, 8), weights_cross_xy(dx(128), dx(128)),
                 H=3, L=1, R=1, V=3, T=1, B=1, fun="weighted_xy", fld="sinxcosy"))
    cs.append(_c("xy_np_fun_cubic_tiles", "XYnpFun", 256, 256, 4, (16, 16), weights_laplace5(0.3), H=3, L=1, R=1, V=3, T=1,
                 B=1, fun="cubic_xy"))
    return cs


CASES = build_cases()
CASE_IDS = [c["name"] for c in CASES]


def case_input(c):
    return field(c["field"], c["nx"], c["ny"])


def case_kwargs(c):
    return dict(H=c["H"], L=c["L"], R=c["R"], V=c["V"], T=c["T"], B=c["B"], fun=c["fun"])


def build_golden_cases():
    """Small cases whose reference-GPU outputs are committed under tests/golden/ (one or two per variant that has
    a working reference; XpFun has none)."""
    rng = np.random.default_rng(2024)
    g = []
    u = lambda n: rng.uniform(-1, 1, n)  # noqa: E731
    g.append(_c("g_xp_9", "Xp", 64, 32, 2, (16, 8), u(9), H=9, L=4, R=4))
    g.append(_c("g_xp_3", "Xp", 64, 32, 1, (16, 8), u(3), H=3, L=1, R=1))
    g.append(_c("g_xnp_5", "Xnp", 64, 32, 1, (16, 8), u(5), H=5, L=2, R=2))
    g.append(_c("g_xnpfun_3", "XnpFun", 64, 32, 2, (16, 8), u(1), H=3, L=1, R=1, fun="second_diff_x", numCoe=1))
    g.append(_c("g_yp_5", "Yp", 32, 64, 2, (8, 8), u(5), V=5, T=2, B=2))
    g.append(_c("g_ynp_9", "Ynp", 32, 64, 2, (8, 8), u(9), V=9, T=4, B=4))
    g.append(_c("g_ypfun_9", "YpFun", 32, 64, 1, (8, 8), u(9), V=9, T=4, B=4, fun="weighted9_y", numCoe=9))
    g.append(_c("g_ynpfun_9", "YnpFun", 32, 64, 2, (8, 8), u(9), V=9, T=4, B=4, fun="weighted9_y"))
    g.append(_c("g_xyp_3x3", "XYp", 64, 48, 1, (16, 8), u(9), H=3, L=1, R=1, V=3, T=1, B=1))
    g.append(_c("g_xyp_5x5", "XYp", 64, 64, 2, (16, 8), u(25), H=5, L=2, R=2, V=5, T=2, B=2))
    g.append(_c("g_xynp_3x3", "XYnp", 64, 48, 3, (16, 8), u(9), H=3, L=1, R=1, V=3, T=1, B=1))
    g.append(_c("g_xypfun_w", "XYpFun", 64, 48, 1, (16, 8), u(9), H=3, L=1, R=1, V=3, T=1, B=1, fun="weighted_xy"))
    g.append(_c("g_xypfun_cubic", "XYpFun", 64, 48, 2, (8, 8), u(9), H=3, L=1, R=1, V=3, T=1, B=1, fun="cubic_xy"))
    g.append(_c("g_xynpfun_cubic", "XYnpFun", 64, 48, 1, (8, 8), u(9), H=3, L=1, R=1, V=3, T=1, B=1, fun="cubic_xy"))
    return g


GOLDEN_CASES = build_golden_cases()
