"""Python binding of the Cahn-Hilliard ADI solver re-hosted on the engine (custen_b200/csrc/cahn.cu; reference
program cuPentCahnADI/src/cuPentCahnADI.cu).  All arithmetic happens in the CUDA library."""
import math

import numpy as np

from . import _lib


class CahnHilliard:
    """2D Cahn-Hilliard, periodic n x n grid of side lx, dt = dt_over_dx * lx / n (reference: D = 1, gamma = 0.01,
    dt_over_dx = 0.1; lx = 2 pi in the demo, 16 pi in the timing twins)."""

    def __init__(self, n, D=1.0, gamma=0.01, lx=16.0 * math.pi, dt_over_dx=0.1, device=0, solver=None, fused=None,
                 graph=None):
        """solver: 2 partitioned tolerance-mode pentadiagonal solve (default; within 1e-13 of the reference's solver),
        0 / 1 the two solves that keep the reference's operation order (bit-identical results).  fused / graph: see
        include/custen_c.h.  None leaves the library default (custen_cahn_set_*)."""
        self.lib = _lib.load()
        self.n = n
        self.h = self.lib.custen_cahn_create(n, D, gamma, lx, dt_over_dx, device)
        for key, value in ((0, solver), (1, fused), (2, graph)):
            if value is not None:
                self.lib.custen_cahn_config(self.h, key, int(value))

    @property
    def solver(self):
        """The pentadiagonal solve in force (2 falls back to 0 where the partitioned layout cannot take the grid)."""
        return self.lib.custen_cahn_config(self.h, 0, -1)

    def set_field(self, c0, c_old=None):
        """c(t = 0); c_old = c(t = -dt) (default: the same field, like the reference's initial condition)."""
        c0 = np.ascontiguousarray(c0, dtype=np.float64)
        assert c0.shape == (self.n, self.n)
        if c_old is None:
            self.lib.custen_cahn_set_field(self.h, c0.ctypes.data)
        else:
            c_old = np.ascontiguousarray(c_old, dtype=np.float64)
            self.lib.custen_cahn_set_fields(self.h, c0.ctypes.data, c_old.ctypes.data)

    def step(self, nsteps=1):
        self.lib.custen_cahn_step(self.h, nsteps)

    def field(self):
        out = np.empty((self.n, self.n))
        self.lib.custen_cahn_get_field(self.h, out.ctypes.data)
        return out

    def time_steps(self, nsteps):
        """Milliseconds per step over `nsteps` steps (CUDA events)."""
        return self.lib.custen_cahn_time_steps(self.h, nsteps) / nsteps

    @property
    def dt(self):
        return self.lib.custen_cahn_dt(self.h)

    def write_snapshot(self, directory, time):
        """c(t) into <directory>/cahn_hilliard_<time>.bin (reference: Print_Out, cuPentCahnADI.cu:103-140)."""
        if self.lib.custen_cahn_write_snapshot(self.h, str(directory).encode(), float(time)) != 0:
            raise OSError(f"cannot write a snapshot into {directory}")

    def destroy(self):
        if self.h:
            self.lib.custen_cahn_destroy(self.h)
            self.h = None


class CahnHilliardSlab:
    """The tolerance-mode solver on y-slabs over several GPUs, one process per GPU (torch.distributed initialised; it is
    used for the rendezvous only: 64-byte CUDA IPC handles and two barriers - no collective on the data path).

    Rank g owns rows [g n/world, (g+1) n/world).  Halo rows and the y-direction partitions' interface values are read in
    place from the two neighbours' memory; the slabs order themselves on the device (include/custen_c.h,
    custen_cahn_slab_*).  Results are bit-identical to CahnHilliard(solver=2) on one GPU.
    """

    def __init__(self, n, D=1.0, gamma=0.01, lx=16.0 * math.pi, dt_over_dx=0.1, group=None, device=None):
        import ctypes
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.lib = lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if n % self.world:
            raise ValueError("n must be divisible by the number of ranks")
        self.n, self.rows = n, n // self.world
        dev = torch.cuda.current_device() if device is None else device
        self.h = lib.custen_cahn_slab_create(n, self.rank, self.world, D, gamma, lx, dt_over_dx, dev)
        if not self.h:
            raise ValueError(f"the partitioned layout cannot take n = {n} on {self.world} slabs")
        if self.world > 1:
            mine = (ctypes.c_char * 64)()
            lib.custen_cahn_slab_export(self.h, ctypes.addressof(mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine), group=group)
            up = (ctypes.c_char * 64).from_buffer_copy(handles[(self.rank - 1) % self.world])
            down = (ctypes.c_char * 64).from_buffer_copy(handles[(self.rank + 1) % self.world])
            lib.custen_cahn_slab_connect(self.h, ctypes.addressof(up), ctypes.addressof(down))
            dist.barrier(group=group)

    def set_field(self, c0_rows, c_old_rows=None):
        c0_rows = np.ascontiguousarray(c0_rows, dtype=np.float64)
        assert c0_rows.shape == (self.rows, self.n)
        if c_old_rows is None:
            self.lib.custen_cahn_slab_set_field(self.h, c0_rows.ctypes.data)
        else:
            c_old_rows = np.ascontiguousarray(c_old_rows, dtype=np.float64)
            self.lib.custen_cahn_slab_set_fields(self.h, c0_rows.ctypes.data, c_old_rows.ctypes.data)
        if self.world > 1:
            self.dist.barrier(group=self.group)   # the neighbours' halo rows are in place before anyone steps

    def step(self, nsteps=1):
        self.lib.custen_cahn_slab_step(self.h, nsteps)

    def time_steps(self, nsteps):
        """Milliseconds per step over `nsteps` steps on this rank (CUDA events on the slab's stream)."""
        return self.lib.custen_cahn_slab_time_steps(self.h, nsteps) / nsteps

    def synchronize(self):
        self.lib.custen_cahn_slab_synchronize(self.h)

    def error(self):
        return self.lib.custen_cahn_slab_error(self.h)

    def field(self):
        out = np.empty((self.rows, self.n))
        self.lib.custen_cahn_slab_get_field(self.h, out.ctypes.data)
        return out

    def destroy(self):
        if self.h:
            self.lib.custen_cahn_slab_synchronize(self.h)
            if self.world > 1:
                self.dist.barrier(group=self.group)   # nobody unmaps memory a neighbour may still be reading
            self.lib.custen_cahn_slab_destroy(self.h)
            self.h = None


class CahnHilliardMultiGpu:
    """The same slabs driven by ONE process over `ngpus` GPUs (custen_cahn_mg_*)."""

    def __init__(self, n, ngpus, D=1.0, gamma=0.01, lx=16.0 * math.pi, dt_over_dx=0.1, devices=None):
        import ctypes
        self.lib = _lib.load()
        self.n, self.ngpus = n, ngpus
        devs = (ctypes.c_int * ngpus)(*(devices if devices is not None else range(ngpus)))
        self.h = self.lib.custen_cahn_mg_create(n, ngpus, ctypes.addressof(devs), D, gamma, lx, dt_over_dx)
        if not self.h:
            raise ValueError(f"the partitioned layout cannot take n = {n} on {ngpus} slabs")

    def set_field(self, c0):
        c0 = np.ascontiguousarray(c0, dtype=np.float64)
        assert c0.shape == (self.n, self.n)
        self.lib.custen_cahn_mg_set_field(self.h, c0.ctypes.data)

    def step(self, nsteps=1):
        self.lib.custen_cahn_mg_step(self.h, nsteps)

    def time_steps(self, nsteps):
        return self.lib.custen_cahn_mg_time_steps(self.h, nsteps) / nsteps

    def error(self):
        return self.lib.custen_cahn_mg_error(self.h)

    def field(self):
        out = np.empty((self.n, self.n))
        self.lib.custen_cahn_mg_get_field(self.h, out.ctypes.data)
        return out

    def destroy(self):
        if self.h:
            self.lib.custen_cahn_mg_destroy(self.h)
            self.h = None
