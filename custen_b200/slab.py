"""Multi-GPU y-slab layer (new; the reference is single-GPU) - Python side.

One process per GPU.  The global ny x nx grid is cut into `world` contiguous y-slabs; rank g owns rows
[g*ny/world, (g+1)*ny/world) of both field buffers.  A sweep needs only the T rows above and the B rows below each
slab - the role the reference gives its per-tile boundaryTop / boundaryBottom pointers
(custenCreateDestroy2DXYp.cu:194-228).  Two transports:

  peer       (default) everything happens in the CUDA library (custen_b200/csrc/slab.cu, `custen_slab_*`): the
             neighbours' buffers are mapped once through CUDA IPC, the stencil kernel's TMA producer reads the halo rows
             straight out of peer memory over NVLink, and the wait for the neighbour sits inside the sweep kernel, in
             front of the halo rows only.  Python's part is the rendezvous: moving 64-byte IPC handles between ranks.
  exchange   every step the edge rows travel with torch.distributed P2P ops (NCCL send/recv on GPUs, gloo in the CPU
             tests) into local halo buffers; a ring for periodic variants, an open line otherwise.  Kept as the
             library-collective baseline the peer transport is measured against.

X variants need neither (rows are independent).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


def slab_rows(ny, world, rank):
    """Row range [lo, hi) of `rank`'s slab.  ny must divide evenly (like the reference's ny % numTiles == 0)."""
    if ny % world:
        raise ValueError(f"ny={ny} is not divisible by the number of slabs {world}")
    n = ny // world
    return rank * n, (rank + 1) * n


def neighbours(rank, world, periodic):
    """(rank above, rank below) or None where the slab touches a physical edge of a non-periodic grid."""
    up = rank - 1 if rank > 0 else (world - 1 if periodic else None)
    down = rank + 1 if rank < world - 1 else (0 if periodic else None)
    return up, down


def exchange_halos(local, T, B, top_buf, bottom_buf, rank, world, periodic, group=None):
    """Fill top_buf (T rows, from the slab above) and bottom_buf (B rows, from the slab below).

    `local` is this rank's slab (rows x nx, contiguous).  Works on CPU tensors (gloo) and CUDA tensors (nccl).
    With world == 1 and a periodic grid the halos are the slab's own far edges.
    """
    up, down = neighbours(rank, world, periodic)
    if world == 1:
        if periodic:
            if T:
                top_buf.copy_(local[-T:])
            if B:
                bottom_buf.copy_(local[:B])
        return
    ops = []
    # my last T rows are the top halo of the slab below; my first B rows the bottom halo of the slab above
    if down is not None and T:
        ops.append(dist.P2POp(dist.isend, local[-T:].contiguous(), down, group))
    if up is not None and B:
        ops.append(dist.P2POp(dist.isend, local[:B].contiguous(), up, group))
    if up is not None and T:
        ops.append(dist.P2POp(dist.irecv, top_buf, up, group))
    if down is not None and B:
        ops.append(dist.P2POp(dist.irecv, bottom_buf, down, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class _DeviceView:
    """Zero-copy torch view of a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_view(ptr, shape, device):
    return torch.as_tensor(_DeviceView(ptr, shape), device=device)


class SlabStencil:
    """This rank's slab of a global grid, time-stepped with Compute + Swap.

    variant, coef, H..B, fun: as custen_b200.Stencil2D (coef: numpy array).  The slab owns both field buffers;
    `input` / `output` are torch views of the current roles, `step()` = one sweep, `swap()` trades the roles,
    `run(n)` = n x (sweep + swap).
    """

    def __init__(self, variant, nx, ny_global, coef, transport="peer", group=None, device=None, H=1, L=0, R=0, V=1, T=0,
                 B=0, fun=None, numCoe=None, numTiles=1):
        from . import api, _lib
        self.lib = lib = _lib.load()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.variant = variant
        self.periodic = not variant.replace("Fun", "").endswith("np")
        self.is_x = not variant.startswith("XY") and variant.startswith("X")
        lo, hi = slab_rows(ny_global, self.world, self.rank)
        self.row0, self.rows, self.nx = lo, hi - lo, nx
        self.T, self.B = (0, 0) if self.is_x else (T, B)
        self.transport = transport
        dev = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", dev)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        ncoef = int(numCoe) if numCoe is not None else coef.size
        fp = api.builtin_fun(fun) if isinstance(fun, str) else fun
        self.slab = None
        up, down = neighbours(self.rank, self.world, self.periodic)
        if transport == "peer":
            self.slab = lib.custen_slab_create(_lib.VARIANTS.index(variant), dev, self.rank, self.world, nx, ny_global,
                                               coef.ctypes.data, ncoef, H, L, R, V, T, B, fp)
            if self.world > 1:
                mine, off = (ctypes.c_char * 64)(), ctypes.c_size_t(0)
                lib.custen_slab_export(self.slab, ctypes.addressof(mine), ctypes.byref(off))
                handles, offs = [None] * self.world, [None] * self.world
                dist.all_gather_object(handles, bytes(mine), group=group)
                dist.all_gather_object(offs, int(off.value), group=group)
                keep = []

                def arg(r):
                    if r is None:
                        return None, 0
                    buf = (ctypes.c_char * 64).from_buffer_copy(handles[r])
                    keep.append(buf)
                    return ctypes.addressof(buf), offs[r]

                (uh, uo), (dh, do) = arg(up), arg(down)
                lib.custen_slab_connect(self.slab, uh, uo, dh, do)
                dist.barrier(group=group)
        elif transport == "exchange":
            self._buf = [torch.zeros((self.rows, nx), dtype=torch.float64, device=self.device) for _ in range(2)]
            self._cur = 0
            self._coef = torch.from_numpy(coef).to(self.device)
            self.st = api.Stencil2D(variant, nx, self.rows, self._buf[1], self._buf[0], self._coef, deviceNum=dev, H=H, L=L,
                                    R=R, V=V, T=T, B=B, fun=fun, numCoe=numCoe, numTiles=numTiles)
            if not self.is_x:
                self.top = torch.empty((max(self.T, 1), nx), dtype=torch.float64, device=self.device)
                self.bottom = torch.empty((max(self.B, 1), nx), dtype=torch.float64, device=self.device)
                self.st.set_slab(self.top, self.bottom, self.rank == 0, self.rank == self.world - 1)
        else:
            raise ValueError(transport)

    # ---- buffers ---------------------------------------------------------------------------------------------------
    def _field(self, which):
        if self.slab:
            return device_view(self.lib.custen_slab_field(self.slab, which), (self.rows, self.nx), self.device)
        return self._buf[self._cur ^ which]

    @property
    def input(self):
        return self._field(0)

    @property
    def output(self):
        return self._field(1)

    # ---- stepping --------------------------------------------------------------------------------------------------
    def step(self):
        """One sweep over the global grid (this rank's share), halos included."""
        if self.slab:
            self.lib.custen_slab_compute(self.slab)
            return
        if not self.is_x and (self.T or self.B):
            exchange_halos(self.input, self.T, self.B, self.top[: self.T], self.bottom[: self.B], self.rank, self.world,
                           self.periodic, self.group)
        self.st.compute(0)

    def swap(self):
        if self.slab:
            self.lib.custen_slab_swap(self.slab)
            return
        self.st.swap(self._buf[self._cur ^ 1])
        self._cur ^= 1

    def run(self, nsteps):
        """nsteps x (sweep + swap); the peer transport replays pairs of steps from a CUDA graph."""
        if self.slab:
            self.lib.custen_slab_run(self.slab, nsteps)
            return
        for _ in range(nsteps):
            self.step()
            self.swap()

    def synchronize(self):
        if self.slab:
            self.lib.custen_slab_synchronize(self.slab)
        else:
            torch.cuda.synchronize()

    @property
    def path(self):
        from .api import PATH_NAMES
        if self.slab:
            return PATH_NAMES.get(self.lib.custen_slab_last_path(self.slab), "?")
        return self.st.path

    def error(self):
        return bool(self.slab and self.lib.custen_slab_error(self.slab))

    def destroy(self):
        torch.cuda.synchronize()
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)  # nobody unmaps memory a neighbour may still be reading
        if self.slab:
            self.lib.custen_slab_destroy(self.slab)
            self.slab = None
        elif getattr(self, "st", None):
            self.st.destroy()
            self.st = None
