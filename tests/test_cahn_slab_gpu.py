"""BASELINE.json config 5 on y-slabs (custen_cahn_slab_* / custen_cahn_mg_*, custen_b200/csrc/cahn_part.cu): every slab
reads only halo rows and partition interface values from its two neighbours.  Must be bit-identical to the single-GPU
tolerance-mode solver (partitions are solved with the same arithmetic wherever they live), which is within 1e-13 per
step of the reference's GPU solver (tests/test_cahn_gpu.py).

The slab logic (halo pointers, interface exchange, device-side ordering, graph replay) is also exercised on ONE GPU:
several slabs of one grid on the same device, each on its own stream - the driver's single-GPU box runs these."""
import math
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from custen_b200.cahn import CahnHilliard, CahnHilliardMultiGpu  # noqa: E402

NGPU = torch.cuda.device_count()


def _initial(n, seed=0):
    return np.random.default_rng(seed).uniform(-0.1, 0.1, size=(n, n))


def _single(c0, steps, np_rows=None, device=0):
    import custen_b200 as cs
    if np_rows:
        cs.load().custen_cahn_set_partition_rows(np_rows)
    try:
        s = CahnHilliard(c0.shape[0], solver=2, device=device)
        assert s.solver == 2
        s.set_field(c0)
        s.step(steps)
        out = s.field()
        s.destroy()
    finally:
        cs.load().custen_cahn_set_partition_rows(128)
    return out


def _mg(c0, steps, devices, np_rows=None, graph=1, chunks=None):
    import custen_b200 as cs
    if np_rows:
        cs.load().custen_cahn_set_partition_rows(np_rows)
    try:
        m = CahnHilliardMultiGpu(c0.shape[0], len(devices), devices=devices)
    finally:
        cs.load().custen_cahn_set_partition_rows(128)
    cs.load().custen_cahn_mg_set_graph(m.h, graph)
    m.set_field(c0)
    for k in (chunks or [steps]):
        m.step(k)
    out = m.field()
    err = m.error()
    m.destroy()
    assert err == 0, f"{err} neighbour waits timed out"
    return out


@pytest.mark.parametrize("n,slabs,np_rows,steps", [(256, 2, 64, 7), (256, 4, 32, 6), (512, 2, 128, 9), (512, 4, 64, 12),
                                                    (1024, 8, 32, 5), (1024, 4, 128, 6), (2048, 8, 128, 4)])
def test_slabs_on_one_gpu_match_the_single_slab_solver(n, slabs, np_rows, steps):
    c0 = _initial(n, seed=n + slabs)
    want = _single(c0, steps, np_rows)
    got = _mg(c0, steps, [0] * slabs, np_rows)
    assert ol.count_diff(got, want) == 0


@pytest.mark.parametrize("graph", [0, 1])
def test_slab_steps_compose_and_graph_replay_matches(graph):
    c0 = _initial(256, seed=5)
    want = _single(c0, 11, 64)
    got = _mg(c0, 11, [0, 0], 64, graph=graph, chunks=[1, 3, 2, 4, 1])
    assert ol.count_diff(got, want) == 0


def test_layouts_the_slabs_cannot_take_are_refused():
    with pytest.raises(ValueError):
        CahnHilliardMultiGpu(256, 3, devices=[0, 0, 0])      # 256 / 3
    with pytest.raises(ValueError):
        CahnHilliardMultiGpu(128, 4, devices=[0, 0, 0, 0])   # one 32-row partition per slab: coupling reaches further


@pytest.mark.skipif(NGPU < 2, reason="needs at least two CUDA devices")
@pytest.mark.parametrize("n,steps", [(512, 12), (2048, 6), (4096, 4)])
def test_one_process_several_gpus(n, steps):
    world = 8 if NGPU >= 8 else 4 if NGPU >= 4 else 2
    c0 = _initial(n, seed=n)
    want = _single(c0, steps)
    got = _mg(c0, steps, list(range(world)))
    assert ol.count_diff(got, want) == 0


def _worker(rank, world, port, n, steps, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from custen_b200.cahn import CahnHilliardSlab
        c0 = _initial(n, seed=11)
        rows = n // world
        want = _single(c0, steps, device=rank)[rank * rows:(rank + 1) * rows]
        torch.cuda.set_device(rank)
        slab = CahnHilliardSlab(n)
        slab.set_field(c0[rank * rows:(rank + 1) * rows])
        slab.step(steps)
        got = slab.field()
        err = slab.error()
        slab.destroy()
        q.put((rank, ol.count_diff(got, want) + 1000000 * err))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(NGPU < 2, reason="needs at least two CUDA devices")
@pytest.mark.parametrize("n,steps", [(512, 12), (2048, 7)])
def test_one_process_per_gpu_over_cuda_ipc(n, steps):
    import torch.multiprocessing as mp
    world = min(NGPU, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + n % 97
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: 0 for r in range(world)}
