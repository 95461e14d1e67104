"""The partitioned pentadiagonal solve's DEVICE kernels (custen_b200/csrc/pent_part.cu: k_part_cols, k_part_rows,
k_spike_reduce) on caller-supplied systems, against the host restatement of the same arithmetic (pinned against dense
solves in tests/test_pent_part_cpu.py): bit for bit, including stiff systems whose interface coupling reaches several
partitions."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import custen_b200 as cs  # noqa: E402
import oracle_lib as ol  # noqa: E402


def _sigma(n, lx=16.0 * math.pi):
    dx = lx / n
    return 2.0 * (0.1 * dx) * 0.01 / (3.0 * dx ** 4)


def _coef(sig):
    return np.array([sig, -4 * sig, 1 + 6 * sig, -4 * sig, sig])


def _host(n, npart, co, rhs_rows):
    """rhs_rows[sys][i] -> x[sys][i] through custen_pent_part_host"""
    lib = cs.load()
    out = np.empty_like(rhs_rows)
    nb = 0
    for k in range(rhs_rows.shape[0]):
        r = np.ascontiguousarray(rhs_rows[k])
        x = np.empty(n)
        nb = lib.custen_pent_part_host(n, npart, co.ctypes.data, r.ctypes.data, x.ctypes.data)
        out[k] = x
    return nb, out


def _device(n, npart, co, rhs, nsys, layout):
    lib = cs.load()
    rhs = np.ascontiguousarray(rhs)
    x = np.empty_like(rhs)
    nb = lib.custen_pent_part_device(n, npart, co.ctypes.data, nsys, rhs.ctypes.data, x.ctypes.data, layout)
    return nb, x


@pytest.mark.parametrize("n,npart,sig", [(256, 32, None), (256, 64, 0.05), (512, 128, None), (512, 256, 2.0),
                                         (1024, 128, 45.0), (1024, 64, 360.0), (2048, 128, None), (4096, 128, None),
                                         (4096, 64, None), (4096, 256, None), (4096, 32, None)])
@pytest.mark.parametrize("layout", [0, 1])
def test_device_kernels_match_the_host_restatement(n, npart, sig, layout):
    co = _coef(sig if sig is not None else _sigma(n))
    nsys = 64
    rhs_rows = np.random.default_rng(n + npart + layout).uniform(-1, 1, (nsys, n))
    nb_h, want = _host(n, npart, co, rhs_rows)
    nb_d, got = _device(n, npart, co, rhs_rows if layout == 0 else rhs_rows.T, nsys, layout)
    assert nb_d == nb_h >= 1
    got = got if layout == 0 else got.T
    assert ol.count_diff(np.ascontiguousarray(got), want) == 0, (nb_h, np.max(np.abs(got - want)))


@pytest.mark.parametrize("n,npart,sig", [(128, 32, None), (256, 64, 3.0), (512, 128, 360.0), (1024, 128, None), (2048, 128, None)])
def test_adi_pair_with_the_correction_applied_on_load(n, npart, sig):
    """layout 2: solve along x, then along y with the x-correction applied while the y-tiles are staged."""
    co = _coef(sig if sig is not None else _sigma(n))
    a = np.random.default_rng(n).uniform(-1, 1, (n, n))
    _, x1 = _host(n, npart, co, a)                       # rows of a: systems along x
    _, x2t = _host(n, npart, co, np.ascontiguousarray(x1.T))   # columns: systems along y
    nb, got = _device(n, npart, co, a, n, 2)
    assert nb >= 1
    assert ol.count_diff(got, np.ascontiguousarray(x2t.T)) == 0, np.max(np.abs(got - x2t.T))
