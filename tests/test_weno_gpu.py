"""13th public variant, cuStenCreate/Compute/Swap/Destroy2DXYWENOADVp: periodic fifth-order WENO advection.
Bit-level parity is against the reference's own CUDA kernel (it squares through single-precision powf, which no CPU
libm reproduces); the CPU oracle is a sanity check at single-precision tolerance."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import custen_b200 as cs  # noqa: E402


def _fields(nx, ny, seed=0):
    x = np.arange(nx) * (2 * np.pi / nx)
    y = np.arange(ny) * (2 * np.pi / ny)
    rng = np.random.default_rng(seed)
    phi = np.sin(x)[None, :] * np.cos(y)[:, None] + 0.05 * rng.uniform(-1, 1, (ny, nx))
    u = np.cos(y)[:, None] * np.ones((1, nx)) + 0.1 * rng.uniform(-1, 1, (ny, nx))   # both signs
    v = -np.sin(x)[None, :] * np.ones((ny, 1)) + 0.1 * rng.uniform(-1, 1, (ny, nx))
    return phi, u, v


def _ours(phi, u, v, dx, dy, tiles=1, kind="device", fallback=False):
    ny, nx = phi.shape
    cs.set_tuning(force_fallback=1 if fallback else 0)
    h = cs.cuSten_t()
    if kind == "device":
        t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (phi, u, v)]
        out = torch.zeros_like(t[0])
        cs.cuStenCreate2DXYWENOADVp(h, 0, tiles, nx, ny, 32, 16, dx, dy, t[1], t[2], out, t[0])
        cs.cuStenCompute2DXYWENOADVp(h, cs.DEVICE)
        cs.device_synchronize()
        res = out.cpu().numpy()
    elif kind == "pageable":
        a = [np.ascontiguousarray(x, dtype=np.float64).copy() for x in (phi, u, v)]
        res = np.zeros_like(a[0])
        cs.cuStenCreate2DXYWENOADVp(h, 0, tiles, nx, ny, 32, 16, dx, dy, a[1].ctypes.data, a[2].ctypes.data, res.ctypes.data,
                                    a[0].ctypes.data)
        cs.cuStenCompute2DXYWENOADVp(h, cs.HOST)
        cs.device_synchronize()
    else:
        import ctypes
        lib = cs.load()
        n = nx * ny
        alloc, free = ((lib.custen_managed_alloc, lib.custen_managed_free) if kind == "managed"
                       else (lib.custen_host_alloc, lib.custen_host_free))
        ptrs = [alloc(n * 8) for _ in range(4)]
        views = [np.ctypeslib.as_array((ctypes.c_double * n).from_address(p)) for p in ptrs]
        for vw, a in zip(views[:3], (phi, u, v)):
            vw[:] = a.ravel()
        views[3][:] = 0.0
        cs.cuStenCreate2DXYWENOADVp(h, 0, tiles, nx, ny, 32, 16, dx, dy, ptrs[1], ptrs[2], ptrs[3], ptrs[0])
        cs.cuStenCompute2DXYWENOADVp(h, cs.HOST)
        cs.device_synchronize()
        res = views[3].reshape(ny, nx).copy()
    path = cs.last_path(h)
    cs.cuStenDestroy2DXYWENOADVp(h)
    if kind in ("managed", "pinned"):
        for p in ptrs:
            free(p)
    cs.set_tuning()
    return res, path


@pytest.mark.parametrize("nx,ny,tiles", [(256, 128, 1), (512, 512, 4), (1024, 256, 2)])
def test_bit_exact_against_reference_weno_kernel(nx, ny, tiles):
    phi, u, v = _fields(nx, ny, seed=nx)
    dx, dy = 2 * np.pi / nx, 2 * np.pi / ny
    ref = ol.ref_weno(phi, u, v, dx, dy, tiles=tiles, block=(32, 16))
    got, path = _ours(phi, u, v, dx, dy, tiles=tiles)
    assert path == "stream_tile"
    assert ol.count_diff(got, ref) == 0
    got_fb, path = _ours(phi, u, v, dx, dy, tiles=tiles, fallback=True)
    assert path == "fallback" and ol.count_diff(got_fb, ref) == 0
    got_m, _ = _ours(phi, u, v, dx, dy, tiles=tiles, kind="managed")
    assert ol.count_diff(got_m, ref) == 0
    # host-resident fields and velocities: the staged out-of-core road (tile + halo rows + the tile's velocities per slot)
    for kind in ("pinned", "pageable"):
        got_h, _ = _ours(phi, u, v, dx, dy, tiles=tiles, kind=kind)
        assert ol.count_diff(got_h, ref) == 0, kind


def test_windows_outside_the_verified_powf_range_match_the_reference_kernel():
    """Flat and linear regions (exact-zero smoothness arguments), huge and tiny amplitudes, an infinity and a NaN:
    the windows whose powf arguments leave the range in which the restated powf was verified take the library call."""
    nx, ny = 512, 256
    phi, u, v = _fields(nx, ny, seed=3)
    phi[:64, :] = 0.75                                    # flat: every difference is exactly zero
    phi[64:96, :] = np.arange(nx)[None, :] * 0.125        # linear in x with exactly representable steps
    phi[96:128, :] = np.arange(32)[:, None] * 0.5         # linear in y
    phi[128:160, :128] *= 1e30                            # (float) of the second differences overflows 2^60
    phi[128:160, 128:256] *= 1e-30                        # ... and underflows 2^-60
    phi[128:160, 256:384] *= 1e200                        # infinite after the conversion to float
    phi[200, 100] = np.inf
    phi[220, 300] = np.nan
    dx, dy = 2 * np.pi / nx, 2 * np.pi / ny
    with np.errstate(all="ignore"):
        ref = ol.ref_weno(phi, u, v, dx, dy, tiles=2, block=(32, 16))
        got, path = _ours(phi, u, v, dx, dy, tiles=2)
    assert path == "stream_tile"
    nan_r, nan_g = np.isnan(ref), np.isnan(got)
    assert (nan_r == nan_g).all() and nan_r.any()
    assert ol.count_diff(np.where(nan_g, 0.0, got), np.where(nan_r, 0.0, ref)) == 0
    got_fb, path = _ours(phi, u, v, dx, dy, tiles=2, fallback=True)
    assert path == "fallback" and (np.isnan(got_fb) == nan_r).all()
    assert ol.count_diff(np.where(nan_r, 0.0, got_fb), np.where(nan_r, 0.0, ref)) == 0


def test_cpu_oracle_agrees_to_single_precision():
    nx, ny = 128, 96
    phi, u, v = _fields(nx, ny, seed=5)
    dx, dy = 2 * np.pi / nx, 2 * np.pi / ny
    got, _ = _ours(phi, u, v, dx, dy)
    want = ol.oracle_weno(phi, u, v, dx, dy)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-5
    # and it approximates u dphi/dx + v dphi/dy of the smooth part
    assert np.isfinite(got).all()


def test_swap_time_stepping():
    nx, ny = 256, 128
    phi, u, v = _fields(nx, ny, seed=9)
    dx, dy = 2 * np.pi / nx, 2 * np.pi / ny
    t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (phi, u, v)]
    out = torch.zeros_like(t[0])
    h = cs.cuSten_t()
    cs.cuStenCreate2DXYWENOADVp(h, 0, 2, nx, ny, 32, 16, dx, dy, t[1], t[2], out, t[0])
    cs.cuStenCompute2DXYWENOADVp(h, cs.DEVICE)
    cs.device_synchronize()
    cs.cuStenSwap2DXYWENOADVp(h, out)
    cs.cuStenCompute2DXYWENOADVp(h, cs.DEVICE)   # now t[0] <- weno(out)
    cs.device_synchronize()
    step1 = ol.ref_weno(phi, u, v, dx, dy, tiles=2, block=(32, 16))
    step2 = ol.ref_weno(step1, u, v, dx, dy, tiles=2, block=(32, 16))
    assert ol.count_diff(out.cpu().numpy(), step1) == 0
    assert ol.count_diff(t[0].cpu().numpy(), step2) == 0
    cs.cuStenDestroy2DXYWENOADVp(h)
