"""Python binding of the Cahn-Hilliard ADI solver re-hosted on the engine (custen_b200/csrc/cahn.cu; reference
program cuPentCahnADI/src/cuPentCahnADI.cu).  All arithmetic happens in the CUDA library."""
import math

import numpy as np

from . import _lib


class CahnHilliard:
    """2D Cahn-Hilliard, periodic n x n grid of side lx, dt = dt_over_dx * lx / n (reference: D = 1, gamma = 0.01,
    dt_over_dx = 0.1; lx = 2 pi in the demo, 16 pi in the timing twins)."""

    def __init__(self, n, D=1.0, gamma=0.01, lx=16.0 * math.pi, dt_over_dx=0.1, device=0):
        self.lib = _lib.load()
        self.n = n
        self.h = self.lib.custen_cahn_create(n, D, gamma, lx, dt_over_dx, device)

    def set_field(self, c0):
        c0 = np.ascontiguousarray(c0, dtype=np.float64)
        assert c0.shape == (self.n, self.n)
        self.lib.custen_cahn_set_field(self.h, c0.ctypes.data)

    def step(self, nsteps=1):
        self.lib.custen_cahn_step(self.h, nsteps)

    def field(self):
        out = np.empty((self.n, self.n))
        self.lib.custen_cahn_get_field(self.h, out.ctypes.data)
        return out

    def time_steps(self, nsteps):
        """Milliseconds per step over `nsteps` steps (CUDA events)."""
        return self.lib.custen_cahn_time_steps(self.h, nsteps) / nsteps

    def destroy(self):
        if self.h:
            self.lib.custen_cahn_destroy(self.h)
            self.h = None
