"""Run a few Cahn-Hilliard steps (for ncu launch lists and quick timings):
python tools/cahn_steps.py [n] [steps] [solver] [partition_rows]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import custen_b200 as cs  # noqa: E402
from custen_b200.cahn import CahnHilliard  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
solver = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if len(sys.argv) > 4:
    cs.load().custen_cahn_set_partition_rows(int(sys.argv[4]))
s = CahnHilliard(n, solver=solver)
s.set_field(np.random.default_rng(0).uniform(-0.1, 0.1, (n, n)))
s.step(2)
print("n", n, "solver", s.solver, "np", sys.argv[4] if len(sys.argv) > 4 else "default", "ms/step", s.time_steps(steps))
s.destroy()
