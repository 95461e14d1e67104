"""Run a few Cahn-Hilliard steps (for ncu launch lists): python tools/cahn_steps.py [n] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custen_b200.cahn import CahnHilliard  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
s = CahnHilliard(n)
s.set_field(np.random.default_rng(0).uniform(-0.1, 0.1, (n, n)))
s.step(2)
print("ms/step", s.time_steps(steps))
s.destroy()
