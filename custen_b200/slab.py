"""Multi-GPU y-slab layer (new; the reference is single-GPU).

One process per GPU.  The global ny x nx grid is cut into `world` contiguous y-slabs; rank g owns rows
[g*ny/world, (g+1)*ny/world).  A sweep needs only the T rows above and the B rows below each slab — the role
the reference gives its per-tile boundaryTop / boundaryBottom pointers (custenCreateDestroy2DXYp.cu:194-228).
Two transports, both feeding `custen_set_slab`:

  exchange   every step, the edge rows travel with torch.distributed P2P ops (NCCL send/recv on GPUs, gloo in
             the CPU tests) into local halo buffers; a ring for periodic variants, an open line otherwise.
  peer       the neighbours' arrays are mapped once through CUDA IPC and the stencil kernel's TMA producer
             reads the halo rows straight out of peer memory over NVLink, so the transfer is folded into the
             sweep itself; a step then only needs a barrier.

X variants need neither (rows are independent).
"""
import ctypes

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


class SlabStencil:
    """A cuSten handle over this rank's slab of a global grid, with its halo transport.

    variant, coef, H..B, fun: as custen_b200.Stencil2D.  `inp` / `out` are this rank's CUDA tensors (rows x nx).
    transport: "exchange" (NCCL send/recv into halo buffers) or "peer" (IPC-mapped neighbour memory).
    """

    def __init__(self, variant, nx, ny_global, inp, out, coef, transport="exchange", group=None, **kw):
        from . import api, _lib
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.periodic = not variant.replace("Fun", "").endswith("np")
        self.is_x = not variant.startswith("XY") and variant.startswith("X")
        lo, hi = slab_rows(ny_global, self.world, self.rank)
        self.rows, self.nx = hi - lo, nx
        self.T, self.B = (0, 0) if self.is_x else (kw.get("T", 0), kw.get("B", 0))
        self.inp, self.out = inp, out
        self.transport = transport
        self.st = api.Stencil2D(variant, nx, self.rows, out, inp, coef, deviceNum=inp.device.index or 0, **kw)
        self._mapped = []
        up, down = neighbours(self.rank, self.world, self.periodic)
        first, last = self.rank == 0, self.rank == self.world - 1
        if self.is_x:
            return
        if transport == "exchange" or self.world == 1:
            self.top = torch.empty((max(self.T, 1), nx), dtype=torch.float64, device=inp.device)
            self.bottom = torch.empty((max(self.B, 1), nx), dtype=torch.float64, device=inp.device)
            self.st.set_slab(self.top, self.bottom, first, last)
        elif transport == "peer":
            lib = _lib.load()
            mine = (ctypes.c_char * 64)()
            off = ctypes.c_size_t(0)
            # IPC handles name the allocation a pointer lives in; `off` is the tensor's byte offset inside it
            lib.custen_ipc_export(inp.data_ptr(), ctypes.addressof(mine), ctypes.byref(off))
            handles, offs = [None] * self.world, [None] * self.world
            dist.all_gather_object(handles, bytes(mine), group=group)
            dist.all_gather_object(offs, int(off.value), group=group)

            opened = {}

            def peer_ptr(r):
                if r not in opened:  # with two ranks the slab above and the slab below are the same peer
                    buf = (ctypes.c_char * 64).from_buffer_copy(handles[r])
                    opened[r] = lib.custen_ipc_open(ctypes.addressof(buf))
                    self._mapped.append(opened[r])
                return opened[r] + offs[r]

            row = nx * 8
            top = peer_ptr(up) + (self.rows - self.T) * row if up is not None else 0
            bottom = peer_ptr(down) if down is not None else 0
            self.st.set_slab(top, bottom, first, last)
            # neighbour barrier flags (2 x u64 per rank), exchanged the same way
            self._flags = lib.custen_device_alloc(16)
            fh = (ctypes.c_char * 64)()
            lib.custen_ipc_export(self._flags, ctypes.addressof(fh), None)
            fhandles = [None] * self.world
            dist.all_gather_object(fhandles, bytes(fh), group=group)
            fopened = {}

            def flag_ptr(r):
                if r is None:
                    return None
                if r not in fopened:
                    buf = (ctypes.c_char * 64).from_buffer_copy(fhandles[r])
                    fopened[r] = lib.custen_ipc_open(ctypes.addressof(buf))
                    self._mapped.append(fopened[r])
                return fopened[r]

            self._up_flags, self._down_flags = flag_ptr(up), flag_ptr(down)
            self._epoch = 0
            dist.barrier(group=group)
        else:
            raise ValueError(transport)

    def step(self):
        """One sweep over the global grid (this rank's share), halos included."""
        if not self.is_x and (self.T or self.B):
            if self.transport == "exchange" or self.world == 1:
                exchange_halos(self.inp, self.T, self.B, self.top[: self.T], self.bottom[: self.B], self.rank,
                               self.world, self.periodic, self.group)
            else:
                # peer transport: the sweep reads the neighbours' edge rows in place, so all a step needs is to
                # know that the neighbours have finished the previous one (and are done reading my rows)
                from . import _lib
                self._epoch += 1
                _lib.load().custen_peer_barrier(ctypes.addressof(self.st.handle), self._up_flags, self._down_flags,
                                                self._flags, self._epoch)
        self.st.compute(0)

    def destroy(self):
        from . import _lib
        from .api import device_synchronize
        device_synchronize()
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)  # nobody unmaps memory a neighbour may still be reading
        self.st.destroy()
        for p in self._mapped:
            _lib.load().custen_ipc_close(p)
        self._mapped = []
        if getattr(self, "_flags", None):
            _lib.load().custen_device_free(self._flags)
            self._flags = None
