"""Cahn-Hilliard ADI on y-slabs, one process per GPU: ms per step (run under torchrun; events on the slab's own stream,
max over ranks).  python -m torch.distributed.run --nproc-per-node G ... tools/cahn_slab_bench.py [n] [steps] [np]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custen_b200 as cs  # noqa: E402
from custen_b200.cahn import CahnHilliardSlab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if len(sys.argv) > 3:
    cs.load().custen_cahn_set_partition_rows(int(sys.argv[3]))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = n // world
s = CahnHilliardSlab(n)
s.set_field(np.random.default_rng(rank).uniform(-0.1, 0.1, (rows, n)))
s.step(5)
s.synchronize()
dist.barrier()
ms = s.time_steps(steps)
t = torch.tensor([ms], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
err = s.error()
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"n": n, "gpus": world, "ms_per_step": ms, "mpoint_steps_per_s": n * n / ms / 1e3, "wait_timeouts": err,
                      "driver": "one process per GPU (CUDA IPC)"}))
s.destroy()
dist.destroy_process_group()
