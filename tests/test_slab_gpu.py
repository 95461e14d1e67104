"""Multi-GPU y-slab layer on real GPUs (needs >= 2): both halo transports against the oracle's global sweep."""
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


def _worker(rank, world, port, variant, kw, transport, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import custen_b200 as cs
        from custen_b200 import slab
        nx, ny = 512, 256 * world
        full = cases.field("random", nx, ny)
        H, V = kw.get("H", 1), kw.get("V", 1)
        coef = np.random.default_rng(3).uniform(-1, 1, H * V)
        lo, hi = slab.slab_rows(ny, world, rank)
        inp = torch.from_numpy(full[lo:hi].copy()).cuda()
        out = torch.full_like(inp, cases.SENTINEL)
        tcoef = torch.from_numpy(coef).cuda()
        ss = slab.SlabStencil(variant, nx, ny, inp, out, tcoef, transport=transport, **kw)
        dist.barrier()
        ss.step()
        cs.device_synchronize()
        dist.barrier()
        want = ol.oracle_sweep(variant, full, np.full_like(full, cases.SENTINEL), coef, **kw)[lo:hi]
        q.put((rank, ol.count_diff(out.cpu().numpy(), want)))
        ss.destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["exchange", "peer"])
@pytest.mark.parametrize("variant,kw", [
    ("XYp", dict(H=3, L=1, R=1, V=3, T=1, B=1)),
    ("XYnp", dict(H=5, L=2, R=2, V=5, T=2, B=2)),
    ("Yp", dict(V=9, T=4, B=4)),
    ("XYpFun", dict(H=3, L=1, R=1, V=3, T=1, B=1, fun="cubic_xy")),
    ("Xp", dict(H=9, L=4, R=4)),
])
def test_slabs_reassemble_global_sweep(variant, kw, transport):
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (abs(hash((variant, transport))) % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, kw, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: 0 for r in range(world)}
