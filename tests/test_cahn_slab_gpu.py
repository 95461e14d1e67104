"""BASELINE.json config 5 on several GPUs: y-slabs, peer halos for the two stencils, two all-to-all transposes per step.
Must be bit-identical to the single-GPU solver (which is bit-identical to the reference's GPU solver)."""
import math
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


def _worker(rank, world, port, n, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle_lib as ol
        from custen_b200.cahn import CahnHilliard, CahnHilliardSlab
        c0 = np.random.default_rng(11).uniform(-0.1, 0.1, (n, n))
        rows = n // world
        single = CahnHilliard(n, device=rank)
        single.set_field(c0)
        single.step(steps)
        want = single.field()[rank * rows:(rank + 1) * rows]
        single.destroy()
        slab = CahnHilliardSlab(n)
        slab.set_field(c0[rank * rows:(rank + 1) * rows])
        slab.step(steps)
        got = slab.field()
        slab.destroy()
        q.put((rank, ol.count_diff(got, want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,steps", [(128, 6), (512, 12)])
def test_slab_solver_is_bit_identical_to_single_gpu(n, steps):
    world = min(torch.cuda.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + n % 97
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: 0 for r in range(world)}
