"""Host-side logic of the multi-GPU y-slab layer on CPU: two gloo ranks exchange halos with the same code the GPU
path uses (custen_b200.slab.exchange_halos); each rank's slab + halos is then swept by the ORACLE (tests only)
and the pieces must reassemble the oracle's global sweep bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import oracle_lib as ol  # noqa: E402
from custen_b200 import slab  # noqa: E402


def test_slab_rows_and_neighbours():
    assert [slab.slab_rows(32768, 8, r) for r in (0, 7)] == [(0, 4096), (28672, 32768)]
    with pytest.raises(ValueError):
        slab.slab_rows(100, 8, 0)
    assert slab.neighbours(0, 8, True) == (7, 1) and slab.neighbours(7, 8, True) == (6, 0)
    assert slab.neighbours(0, 8, False) == (None, 1) and slab.neighbours(7, 8, False) == (6, None)
    assert slab.neighbours(0, 1, True) == (0, 0) and slab.neighbours(0, 1, False) == (None, None)


def _worker(rank, world, port, variant, kw, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny = 48, 64
        periodic = not variant.replace("Fun", "").endswith("np")
        full = cases.field("random", nx, ny)
        coef = np.random.default_rng(3).uniform(-1, 1, kw.get("H", 1) * kw.get("V", 1))
        lo, hi = slab.slab_rows(ny, world, rank)
        local = torch.from_numpy(full[lo:hi].copy())
        T, B = kw.get("T", 0), kw.get("B", 0)
        top, bot = torch.zeros((T, nx), dtype=torch.float64), torch.zeros((B, nx), dtype=torch.float64)
        slab.exchange_halos(local, T, B, top, bot, rank, world, periodic)
        up, down = slab.neighbours(rank, world, periodic)
        parts = ([top.numpy()] if up is not None else []) + [local.numpy()] + ([bot.numpy()] if down is not None else [])
        ext = np.ascontiguousarray(np.vstack(parts))
        out = np.full_like(ext, cases.SENTINEL)
        ol.oracle_sweep(variant, ext, out, coef, periodic_bits=1 if periodic else 0, **kw)
        t0 = T if up is not None else 0
        mine = out[t0:t0 + (hi - lo)]
        want = ol.oracle_sweep(variant, full, np.full_like(full, cases.SENTINEL), coef, **kw)[lo:hi]
        q.put((rank, ol.count_diff(mine, want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant,kw", [
    ("XYp", dict(H=3, L=1, R=1, V=3, T=1, B=1)),
    ("XYp", dict(H=5, L=2, R=2, V=5, T=2, B=2)),
    ("XYnp", dict(H=3, L=1, R=1, V=3, T=1, B=1)),
    ("Yp", dict(V=9, T=4, B=4)),
    ("Ynp", dict(V=5, T=2, B=2)),
])
def test_two_rank_halo_exchange_reassembles_global_sweep(variant, kw):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (abs(hash((variant, tuple(sorted(kw.items()))))) % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, variant, kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {0: 0, 1: 0}
