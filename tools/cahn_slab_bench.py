"""Cahn-Hilliard ADI on y-slabs: ms per step (run under torchrun).  python -m torch.distributed.run ... tools/cahn_slab_bench.py n steps"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custen_b200.cahn import CahnHilliardSlab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = n // world
s = CahnHilliardSlab(n)
s.set_field(np.random.default_rng(rank).uniform(-0.1, 0.1, (rows, n)))
s.step(3)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s.step(steps)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item()) / steps
    print({"n": n, "gpus": world, "ms_per_step": ms, "mpoint_steps_per_s": n * n / ms / 1e3})
s.destroy()
dist.destroy_process_group()
