"""Python mirror of the reference's cuSten 2D API, bound to the C ABI of libcusten_b200.so.

Function names, argument order and meaning are the reference's (cuSten/src/struct/cuSten_struct_functions.h,
cuSten/src/kernels/stencil_kernels.h): `cuStenCreate2DXYp(handle, deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y,
dataOutput, dataInput, weights, numStenHoriz, numStenLeft, numStenRight, numStenVert, numStenTop, numStenBottom)`
and so on for the 12 variants; `cuStenCompute2D<V>(handle, offload)` with `HOST` / `DEVICE`; Swap; Destroy.
Data arguments are raw addresses: pass an int, a torch tensor, a numpy array or anything with `data_ptr()`.
The handle is a `cuSten_t` ctypes structure laid out like the reference's struct, so public fields are readable.
All compute happens in the CUDA library; nothing here touches grid data.
"""
import ctypes

from . import _lib

DEVICE = 0  # cuSten/cuSten.h:30
HOST = 1    # cuSten/cuSten.h:31

PATH_NAMES = {0: "none", 1: "stream_acc", 2: "stream_tile", 3: "fallback", 4: "stream_inline"}
MODE_NAMES = {0: "resident", 1: "resident_per_tile", 2: "managed_pipeline", 3: "staged", 4: "managed_resident",
              5: "managed_zero_copy"}


class cuSten_t(ctypes.Structure):
    """Field-for-field mirror of the reference handle (cuSten/src/struct/cuSten_struct_type.h:84-122)."""
    _fields_ = (
        [(n, ctypes.c_int) for n in (
            "deviceNum", "numStreams", "numTiles", "nx", "ny", "nyTile", "numSten", "numStenLeft", "numStenRight",
            "numStenTop", "numStenBottom", "numStenHoriz", "numStenVert", "BLOCK_X", "BLOCK_Y", "xGrid", "yGrid",
            "mem_shared")]
        + [(n, ctypes.c_void_p) for n in ("dataInput", "dataOutput", "uVel", "vVel", "weights", "coe")]
        + [("coeDx", ctypes.c_double), ("coeDy", ctypes.c_double)]
        + [(n, ctypes.c_int) for n in ("numCoe", "nxLocal", "nyLocal")]
        + [(n, ctypes.c_void_p) for n in ("boundaryTop", "boundaryBottom")]
        + [("numBoundaryTop", ctypes.c_int), ("numBoundaryBottom", ctypes.c_int)]
        + [(n, ctypes.c_void_p) for n in ("streams", "events", "devFunc")]
    )


def ptr(x):
    """Raw address of a buffer argument."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    if isinstance(x, ctypes.c_void_p):
        return x.value
    raise TypeError(f"cannot take the address of {type(x)!r}")


def _h(handle):
    return ctypes.addressof(handle)


def _make(variant):
    def create(handle, deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y, dataOutput, dataInput, *tail):
        lib = _lib.load()
        assert ctypes.sizeof(handle) == lib.custen_handle_size()
        args = [ptr(a) if i == 0 or (variant.endswith("Fun") and i == len(tail) - 1) else int(a)
                for i, a in enumerate(tail)]
        getattr(lib, f"custenCreate2D{variant}")(_h(handle), deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y,
                                                 ptr(dataOutput), ptr(dataInput), *args)

    def compute(handle, offload):
        getattr(_lib.load(), f"custenCompute2D{variant}")(_h(handle), int(offload))

    def swap(handle, dataInput):
        getattr(_lib.load(), f"custenSwap2D{variant}")(_h(handle), ptr(dataInput))

    def destroy(handle):
        getattr(_lib.load(), f"custenDestroy2D{variant}")(_h(handle))

    create.__name__, compute.__name__ = f"cuStenCreate2D{variant}", f"cuStenCompute2D{variant}"
    swap.__name__, destroy.__name__ = f"cuStenSwap2D{variant}", f"cuStenDestroy2D{variant}"
    return create, compute, swap, destroy


for _v in _lib.VARIANTS:
    _c, _k, _s, _d = _make(_v)
    globals()[f"cuStenCreate2D{_v}"] = _c
    globals()[f"cuStenCompute2D{_v}"] = _k
    globals()[f"cuStenSwap2D{_v}"] = _s
    globals()[f"cuStenDestroy2D{_v}"] = _d
del _v, _c, _k, _s, _d


def cuStenCreate2DXYWENOADVp(handle, deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y, dx, dy, u, v, dataOutput, dataInput):
    """13th variant: periodic fifth-order WENO advection u dphi/dx + v dphi/dy (cuSten_struct_functions.h:309)."""
    _lib.load().custenCreate2DXYWENOADVp(_h(handle), deviceNum, numTiles, nx, ny, BLOCK_X, BLOCK_Y, float(dx), float(dy),
                                         ptr(u), ptr(v), ptr(dataOutput), ptr(dataInput))


def cuStenCompute2DXYWENOADVp(handle, offload):
    _lib.load().custenCompute2DXYWENOADVp(_h(handle), int(offload))


def cuStenSwap2DXYWENOADVp(handle, dataInput):
    _lib.load().custenSwap2DXYWENOADVp(_h(handle), ptr(dataInput))


def cuStenDestroy2DXYWENOADVp(handle):
    _lib.load().custenDestroy2DXYWENOADVp(_h(handle))


def checkError(action):
    _lib.load().custenCheckError(action.encode())


def device_synchronize():
    _lib.load().custen_device_synchronize()


def builtin_fun(name):
    """Device function pointer of a user-function fixture linked into the library (see builtin_funs.cuh)."""
    p = _lib.load().custen_builtin_fun(name.encode())
    if not p:
        raise KeyError(f"no built-in device function named {name!r}")
    return p


def last_path(handle):
    return PATH_NAMES.get(_lib.load().custen_last_path(_h(handle)), "?")


def last_mode(handle):
    return MODE_NAMES.get(_lib.load().custen_last_mode(_h(handle)), "?")


def launch_count():
    return int(_lib.load().custen_launch_count())


def set_tuning(force_fallback=0, force_tile=0, chunk_rows=0, ctas_per_sm=0, force_opaque=0):
    _lib.load().custen_set_tuning(force_fallback, force_tile, chunk_rows, ctas_per_sm, force_opaque)


def set_managed_policy(policy=0):
    """0: unified-memory grids take the resident / zero-copy roads when they apply; 1: always the prefetch pipeline."""
    _lib.load().custen_set_managed_policy(int(policy))


def set_slab(handle, top, bottom, is_first, is_last):
    _lib.load().custen_set_slab(_h(handle), ptr(top), ptr(bottom), int(is_first), int(is_last))


class Stencil2D:
    """Convenience wrapper used by the tests and bench: one handle, one variant, uniform arguments.

    variant  one of custen_b200.VARIANTS            coef  weights or coe buffer (address-able)
    H, L, R  window width / taps left / right       V, T, B  window height / taps above / below
    fun      name of a built-in device function or a raw device function pointer (Fun variants)
    """

    def __init__(self, variant, nx, ny, out, inp, coef, H=1, L=0, R=0, V=1, T=0, B=0, fun=None, numCoe=None,
                 numTiles=1, block=(32, 32), deviceNum=0):
        assert variant in _lib.VARIANTS, variant
        self.variant, self.handle = variant, cuSten_t()
        create = globals()[f"cuStenCreate2D{variant}"]
        is_fun = variant.endswith("Fun")
        fp = builtin_fun(fun) if isinstance(fun, str) else fun
        if is_fun and fp is None:
            raise ValueError("Fun variants need a device function")
        head = (self.handle, deviceNum, numTiles, nx, ny, block[0], block[1], out, inp, coef)
        d = variant[0] if not variant.startswith("XY") else "XY"
        ncoe = numCoe if numCoe is not None else H * V
        if d == "X":
            tail = (H, L, R) + ((ncoe, fp) if is_fun else ())
        elif d == "Y":
            tail = (V, T, B) + (((ncoe, fp) if variant == "YpFun" else (fp,)) if is_fun else ())
        else:
            tail = (H, L, R, V, T, B) + ((fp,) if is_fun else ())
        create(*head, *tail)
        self._alive = True

    def compute(self, offload=DEVICE):
        globals()[f"cuStenCompute2D{self.variant}"](self.handle, offload)

    def swap(self, dataInput):
        globals()[f"cuStenSwap2D{self.variant}"](self.handle, dataInput)

    def destroy(self):
        if self._alive:
            globals()[f"cuStenDestroy2D{self.variant}"](self.handle)
            self._alive = False

    def set_slab(self, top, bottom, is_first, is_last):
        set_slab(self.handle, top, bottom, is_first, is_last)

    @property
    def path(self):
        return last_path(self.handle)

    @property
    def mode(self):
        return last_mode(self.handle)
