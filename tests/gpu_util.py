"""Helpers for the GPU tests: run the new engine through its C ABI on buffers of every memory kind."""
import ctypes

import numpy as np
import torch

import custen_b200 as cs
import cases


def _view(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))


class Buffers:
    """in / out / coef buffers of one memory kind: device | managed | pinned | pageable."""

    def __init__(self, kind, inp, out_init, coef):
        self.kind = kind
        lib = cs.load()
        n = inp.size
        self._free = []
        if kind == "device":
            self.t_in = torch.from_numpy(inp).cuda()
            self.t_out = torch.from_numpy(out_init).cuda()
            self.t_coef = torch.from_numpy(np.ascontiguousarray(coef, dtype=np.float64)).cuda()
            self.inp, self.out, self.coef = self.t_in.data_ptr(), self.t_out.data_ptr(), self.t_coef.data_ptr()
        elif kind in ("managed", "pinned"):
            alloc, free = ((lib.custen_managed_alloc, lib.custen_managed_free) if kind == "managed"
                           else (lib.custen_host_alloc, lib.custen_host_free))
            self.inp, self.out, self.coef = alloc(n * 8), alloc(n * 8), alloc(max(coef.size, 1) * 8)
            self._free = [(free, p) for p in (self.inp, self.out, self.coef)]
            _view(self.inp, n)[:] = inp.ravel()
            _view(self.out, n)[:] = out_init.ravel()
            _view(self.coef, coef.size)[:] = coef
        elif kind == "pageable":
            self.a_in = np.ascontiguousarray(inp).copy()
            self.a_out = np.ascontiguousarray(out_init).copy()
            self.a_coef = np.ascontiguousarray(coef, dtype=np.float64).copy()
            self.inp, self.out, self.coef = self.a_in.ctypes.data, self.a_out.ctypes.data, self.a_coef.ctypes.data
        else:
            raise ValueError(kind)
        self.shape = inp.shape

    def result(self, which="out"):
        cs.device_synchronize()
        n = self.shape[0] * self.shape[1]
        if self.kind == "device":
            return (self.t_out if which == "out" else self.t_in).cpu().numpy()
        if self.kind == "pageable":
            return (self.a_out if which == "out" else self.a_in).copy()
        return _view(self.out if which == "out" else self.inp, n).reshape(self.shape).copy()

    def free(self):
        cs.device_synchronize()
        for f, p in self._free:
            f(p)
        self._free = []


def run_ours(c, inp, kind="device", tiles=None, offload=cs.DEVICE, out_init=None, return_path=False):
    """One sweep of case `c` through the C ABI; returns the output grid (and the kernel family used)."""
    if out_init is None:
        out_init = np.full_like(inp, cases.SENTINEL)
    buf = Buffers(kind, inp, out_init, c["coef"])
    st = cs.Stencil2D(c["variant"], c["nx"], c["ny"], buf.out, buf.inp, buf.coef, H=c["H"], L=c["L"], R=c["R"], V=c["V"],
                      T=c["T"], B=c["B"], fun=c["fun"], numCoe=c["numCoe"], numTiles=tiles or c["tiles"],
                      block=c["block"])
    st.compute(offload)
    res = buf.result()
    path, mode = st.path, st.mode
    st.destroy()
    buf.free()
    return (res, path, mode) if return_path else res
