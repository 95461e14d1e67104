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


class _DeviceBuffer:
    """Zero-copy torch view of a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class CahnHilliardSlab:
    """The same solver on y-slabs over several GPUs (one process per GPU, torch.distributed initialised, NCCL).

    Rank g owns rows [g n/world, (g+1) n/world).  Stencil halo rows are read in place from the neighbours' memory
    (CUDA IPC, custen_set_slab); the y-direction solve needs whole columns, so each step makes two all-to-all
    transposes (torch.distributed.all_to_all_single) between the phases of custen_cahn_slab_phase.  Results are
    bit-identical to the single-GPU solver: every system is solved by one thread in the same order.
    """

    def __init__(self, n, D=1.0, gamma=0.01, lx=16.0 * math.pi, dt_over_dx=0.1, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.ctypes = torch, dist, ctypes
        self.lib = lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if n % self.world or (n // self.world) < 4:
            raise ValueError("n must be divisible by the number of ranks")
        self.n, self.rows = n, n // self.world
        dev = torch.cuda.current_device()
        self.h = lib.custen_cahn_slab_create(n, self.rank, self.world, D, gamma, lx, dt_over_dx, dev)
        count = n * self.rows
        buf = lambda k: lib.custen_cahn_slab_buffer(self.h, k)  # noqa: E731
        self.t_send1 = torch.as_tensor(_DeviceBuffer(buf(3), count), device=f"cuda:{dev}")
        self.t_recv = torch.as_tensor(_DeviceBuffer(buf(4), count), device=f"cuda:{dev}")
        self.t_send2 = torch.as_tensor(_DeviceBuffer(buf(5), count), device=f"cuda:{dev}")
        # neighbours' field / cBar buffers and barrier flags through CUDA IPC
        up, down = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        self._mapped, opened = [], {}

        def exchange(ptr):
            hd, off = (ctypes.c_char * 64)(), ctypes.c_size_t(0)
            lib.custen_ipc_export(ptr, ctypes.addressof(hd), ctypes.byref(off))
            hs, offs = [None] * self.world, [None] * self.world
            dist.all_gather_object(hs, bytes(hd), group=group)
            dist.all_gather_object(offs, int(off.value), group=group)

            def peer(r):
                key = hs[r]
                if key not in opened:
                    b = (ctypes.c_char * 64).from_buffer_copy(hs[r])
                    opened[key] = lib.custen_ipc_open(ctypes.addressof(b))
                    self._mapped.append(opened[key])
                return opened[key] + offs[r]
            return peer(up), peer(down)

        row = n * 8
        for which, T in ((0, 1), (1, 1), (2, 2)):   # nonlinear term: 3 x 3 (T = B = 1); linear term: 5 x 5 (T = B = 2)
            p_up, p_down = exchange(buf(which))
            handle = lib.custen_cahn_slab_handle(self.h, which)
            lib.custen_set_slab(handle, p_up + (self.rows - T) * row, p_down, 0, 0)
        self._flags = lib.custen_device_alloc(16)
        self._up_flags, self._down_flags = exchange(self._flags)
        self._epoch = 0
        dist.barrier(group=group)

    def set_field(self, c0_rows):
        c0_rows = np.ascontiguousarray(c0_rows, dtype=np.float64)
        assert c0_rows.shape == (self.rows, self.n)
        self.lib.custen_cahn_slab_set_field(self.h, c0_rows.ctypes.data)
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    def step(self, nsteps=1):
        lib, dist = self.lib, self.dist
        for _ in range(nsteps):
            lib.custen_cahn_slab_phase(self.h, 0)
            self._epoch += 1
            lib.custen_peer_barrier(None, self._up_flags, self._down_flags, self._flags, self._epoch)
            lib.custen_cahn_slab_phase(self.h, 1)
            dist.all_to_all_single(self.t_recv, self.t_send1, group=self.group)
            lib.custen_cahn_slab_phase(self.h, 2)
            dist.all_to_all_single(self.t_recv, self.t_send2, group=self.group)
            lib.custen_cahn_slab_phase(self.h, 3)

    def field(self):
        out = np.empty((self.rows, self.n))
        self.lib.custen_cahn_slab_get_field(self.h, out.ctypes.data)
        return out

    def destroy(self):
        if self.h:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            self.lib.custen_cahn_destroy(self.h)
            for p in self._mapped:
                self.lib.custen_ipc_close(p)
            self.lib.custen_device_free(self._flags)
            self.h = None
