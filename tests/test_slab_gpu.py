"""Multi-GPU y-slab layer on real GPUs (needs >= 2), one process per GPU: both halo transports time-stepping a global
grid (Compute + Swap, so every step depends on the neighbours' previous one) against the oracle's whole-grid stepping.
The single-process twin of these cases (tests/test_slab_c_gpu.py) also runs on a one-GPU box."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs at least two CUDA devices", allow_module_level=True)

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle_lib as ol  # noqa: E402

STEPS = 6


def _coef(kw, fun):
    if fun == "cubic_xy":
        eps = 0.05
        return np.array([0, -eps, 0, -eps, -1 + 4 * eps, -eps, 0, -eps, 0], dtype=np.float64)
    c = np.random.default_rng(3).uniform(-1, 1, kw.get("H", 1) * kw.get("V", 1))
    return c / np.abs(c).sum()


def _worker(rank, world, port, variant, kw, fun, transport, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from custen_b200 import slab
        nx, ny = 1024, 256 * world
        full = cases.field("random", nx, ny)
        coef = _coef(kw, fun)
        lo, hi = slab.slab_rows(ny, world, rank)
        ss = slab.SlabStencil(variant, nx, ny, coef, transport=transport, fun=fun, **kw)
        ss.input.copy_(torch.from_numpy(full[lo:hi].copy()))
        ss.output.fill_(cases.SENTINEL)
        torch.cuda.synchronize()
        dist.barrier()
        ss.run(STEPS)
        ss.synchronize()
        dist.barrier()
        got = ss.input.cpu().numpy()
        err = ss.error()
        want, _ = ol.oracle_time_steps(variant, full, np.full_like(full, cases.SENTINEL), coef, STEPS, fun=fun, **kw)
        q.put((rank, ol.count_diff(got, want[lo:hi]) + (10 ** 9 if err else 0)))
        ss.destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["exchange", "peer"])
@pytest.mark.parametrize("variant,kw,fun", [
    ("XYp", dict(H=3, L=1, R=1, V=3, T=1, B=1), None),
    ("XYnp", dict(H=5, L=2, R=2, V=5, T=2, B=2), None),
    ("Yp", dict(V=9, T=4, B=4), None),
    ("XYpFun", dict(H=3, L=1, R=1, V=3, T=1, B=1), "cubic_xy"),
    ("Xp", dict(H=9, L=4, R=4), None),
])
def test_slabs_time_step_like_the_whole_grid(variant, kw, fun, transport):
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (abs(hash((variant, transport))) % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, kw, fun, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: 0 for r in range(world)}
