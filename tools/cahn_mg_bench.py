"""Cahn-Hilliard ADI on y-slabs driven by ONE process (custen_cahn_mg_*): ms per step on 1 .. all GPUs of the box.
python tools/cahn_mg_bench.py [n] [steps] [np]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custen_b200 as cs  # noqa: E402
from custen_b200.cahn import CahnHilliardMultiGpu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if len(sys.argv) > 3:
    cs.load().custen_cahn_set_partition_rows(int(sys.argv[3]))
c0 = np.random.default_rng(0).uniform(-0.1, 0.1, (n, n))
g = 1
while g <= torch.cuda.device_count():
    try:
        m = CahnHilliardMultiGpu(n, g)
    except ValueError as ex:
        print(json.dumps({"n": n, "gpus": g, "error": str(ex)}))
        g *= 2
        continue
    m.set_field(c0)
    m.step(5)
    ms = m.time_steps(steps)
    print(json.dumps({"n": n, "gpus": g, "ms_per_step": ms, "mpoint_steps_per_s": n * n / ms / 1e3, "wait_timeouts": m.error(),
                      "driver": "one process (custen_cahn_mg_*)"}), flush=True)
    m.destroy()
    g *= 2
