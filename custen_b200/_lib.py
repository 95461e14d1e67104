"""ctypes loader for libcusten_b200.so (the C ABI declared in include/custen_c.h).

There is no CPU fallback: if the CUDA library has not been built, or cannot be loaded, importing the
API raises.  Build it with `make lib` (or `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcusten_b200.so")

VARIANTS = ("Xp", "Xnp", "XpFun", "XnpFun", "Yp", "Ynp", "YpFun", "YnpFun", "XYp", "XYnp", "XYpFun", "XYnpFun")

_c_int, _c_dbl_p, _c_void_p = ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p
_PREFIX = [_c_void_p] + [_c_int] * 6 + [_c_dbl_p, _c_dbl_p]

# trailing Create parameters after dataInput, per variant (include/custen_c.h)
_CREATE_TAIL = {
    "Xp": [_c_dbl_p, _c_int, _c_int, _c_int],
    "Xnp": [_c_dbl_p, _c_int, _c_int, _c_int],
    "XpFun": [_c_dbl_p, _c_int, _c_int, _c_int, _c_int, _c_dbl_p],
    "XnpFun": [_c_dbl_p, _c_int, _c_int, _c_int, _c_int, _c_dbl_p],
    "Yp": [_c_dbl_p, _c_int, _c_int, _c_int],
    "Ynp": [_c_dbl_p, _c_int, _c_int, _c_int],
    "YpFun": [_c_dbl_p, _c_int, _c_int, _c_int, _c_int, _c_dbl_p],
    "YnpFun": [_c_dbl_p, _c_int, _c_int, _c_int, _c_dbl_p],
    "XYp": [_c_dbl_p] + [_c_int] * 6,
    "XYnp": [_c_dbl_p] + [_c_int] * 6,
    "XYpFun": [_c_dbl_p] + [_c_int] * 6 + [_c_dbl_p],
    "XYnpFun": [_c_dbl_p] + [_c_int] * 6 + [_c_dbl_p],
}

# every symbol include/custen_c.h declares (tests check the library exports all of them)
EXPORTED = (
    [f"custen{op}2D{v}" for v in VARIANTS + ("XYWENOADVp",) for op in ("Create", "Swap", "Destroy", "Compute")]
    + ["custenCheckError", "custen_handle_size", "custen_device_synchronize", "custen_builtin_fun",
       "custen_last_path", "custen_last_mode", "custen_launch_count", "custen_set_tuning", "custen_set_managed_policy", "custen_set_slab",
       "custen_ipc_export", "custen_ipc_open", "custen_ipc_close", "custen_event_create", "custen_event_record",
       "custen_event_synchronize", "custen_event_elapsed_ms", "custen_event_destroy", "custen_host_alloc",
       "custen_host_free", "custen_managed_alloc", "custen_managed_free", "custen_peer_barrier", "custen_device_alloc",
       "custen_device_free", "custen_cahn_create", "custen_cahn_set_field", "custen_cahn_step", "custen_cahn_get_field",
       "custen_cahn_time_steps", "custen_cahn_destroy", "custen_cahn_set_table_rows", "custen_cahn_set_solver", "custen_cahn_set_fused", "custen_cahn_set_graph", "custen_debug_bands", "custen_cahn_slab_create",
       "custen_cahn_slab_export", "custen_cahn_slab_connect", "custen_cahn_slab_connect_local", "custen_cahn_slab_set_field",
       "custen_cahn_slab_set_fields", "custen_cahn_slab_get_field", "custen_cahn_slab_step", "custen_cahn_slab_time_steps",
       "custen_cahn_slab_synchronize", "custen_cahn_slab_error", "custen_cahn_slab_set_timeout", "custen_cahn_slab_set_graph",
       "custen_cahn_slab_partition_rows", "custen_cahn_slab_destroy", "custen_cahn_mg_create", "custen_cahn_mg_set_field",
       "custen_cahn_mg_get_field", "custen_cahn_mg_step", "custen_cahn_mg_time_steps", "custen_cahn_mg_error",
       "custen_cahn_mg_set_graph", "custen_cahn_mg_destroy", "custen_cahn_set_fields", "custen_cahn_write_snapshot", "custen_cahn_dt", "custen_cahn_set_rhs_stream",
       "custen_set_handle_managed_policy", "custen_mem_advise", "custen_mem_prefetch", "custen_fill_hash",
       "custen_slab_create", "custen_slab_export", "custen_slab_connect", "custen_slab_field", "custen_slab_rows",
       "custen_slab_compute", "custen_slab_swap", "custen_slab_run", "custen_slab_run_plain", "custen_slab_time_run",
       "custen_slab_synchronize", "custen_slab_error", "custen_slab_set_timeout", "custen_slab_last_path",
       "custen_slab_destroy", "custen_mg_create", "custen_mg_scatter", "custen_mg_gather", "custen_mg_fill_output",
       "custen_mg_compute", "custen_mg_swap", "custen_mg_run", "custen_mg_synchronize", "custen_mg_error", "custen_mg_slab",
       "custen_mg_destroy", "custen_device_numa_node", "custen_host_alloc_near", "custen_host_free_near", "custen_link_probe",
       "custen_cahn_set_partition_rows", "custen_cahn_config", "custen_pent_part_host", "custen_pent_part_choose_np", "custen_pent_part_device"]
)

_lib = None


def load():
    """Load the CUDA library (once) and declare the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: cuSten-B200 has no CPU fallback; build the CUDA library with `make lib`")
    lib = ctypes.CDLL(LIB_PATH)
    for v in VARIANTS:
        f = getattr(lib, f"custenCreate2D{v}")
        f.argtypes, f.restype = _PREFIX + _CREATE_TAIL[v], None
        f = getattr(lib, f"custenSwap2D{v}")
        f.argtypes, f.restype = [_c_void_p, _c_dbl_p], None
        f = getattr(lib, f"custenDestroy2D{v}")
        f.argtypes, f.restype = [_c_void_p], None
        f = getattr(lib, f"custenCompute2D{v}")
        f.argtypes, f.restype = [_c_void_p, _c_int], None
    f = lib.custenCreate2DXYWENOADVp
    f.argtypes = [_c_void_p] + [_c_int] * 6 + [ctypes.c_double, ctypes.c_double] + [_c_dbl_p] * 4
    f.restype = None
    lib.custenSwap2DXYWENOADVp.argtypes, lib.custenSwap2DXYWENOADVp.restype = [_c_void_p, _c_dbl_p], None
    lib.custenDestroy2DXYWENOADVp.argtypes, lib.custenDestroy2DXYWENOADVp.restype = [_c_void_p], None
    lib.custenCompute2DXYWENOADVp.argtypes, lib.custenCompute2DXYWENOADVp.restype = [_c_void_p, _c_int], None
    lib.custenCheckError.argtypes, lib.custenCheckError.restype = [ctypes.c_char_p], None
    lib.custen_handle_size.argtypes, lib.custen_handle_size.restype = [], ctypes.c_size_t
    lib.custen_device_synchronize.argtypes, lib.custen_device_synchronize.restype = [], None
    lib.custen_builtin_fun.argtypes, lib.custen_builtin_fun.restype = [ctypes.c_char_p], ctypes.c_void_p
    lib.custen_last_path.argtypes, lib.custen_last_path.restype = [_c_void_p], _c_int
    lib.custen_last_mode.argtypes, lib.custen_last_mode.restype = [_c_void_p], _c_int
    lib.custen_launch_count.argtypes, lib.custen_launch_count.restype = [], ctypes.c_uint64
    lib.custen_set_tuning.argtypes, lib.custen_set_tuning.restype = [_c_int] * 5, None
    lib.custen_set_managed_policy.argtypes, lib.custen_set_managed_policy.restype = [_c_int], None
    lib.custen_set_slab.argtypes, lib.custen_set_slab.restype = [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int], None
    lib.custen_ipc_export.argtypes, lib.custen_ipc_export.restype = [_c_void_p, _c_void_p, _c_void_p], None
    lib.custen_ipc_open.argtypes, lib.custen_ipc_open.restype = [_c_void_p], ctypes.c_void_p
    lib.custen_ipc_close.argtypes, lib.custen_ipc_close.restype = [_c_void_p], None
    lib.custen_event_create.argtypes, lib.custen_event_create.restype = [], ctypes.c_void_p
    lib.custen_event_record.argtypes, lib.custen_event_record.restype = [_c_void_p, _c_void_p, _c_int], None
    lib.custen_event_synchronize.argtypes, lib.custen_event_synchronize.restype = [_c_void_p], None
    lib.custen_event_elapsed_ms.argtypes, lib.custen_event_elapsed_ms.restype = [_c_void_p, _c_void_p], ctypes.c_float
    lib.custen_event_destroy.argtypes, lib.custen_event_destroy.restype = [_c_void_p], None
    lib.custen_host_alloc.argtypes, lib.custen_host_alloc.restype = [ctypes.c_size_t], ctypes.c_void_p
    lib.custen_host_free.argtypes, lib.custen_host_free.restype = [_c_void_p], None
    lib.custen_peer_barrier.argtypes = [_c_void_p, _c_void_p, _c_void_p, _c_void_p, ctypes.c_uint64]
    lib.custen_peer_barrier.restype = None
    lib.custen_device_alloc.argtypes, lib.custen_device_alloc.restype = [ctypes.c_size_t], ctypes.c_void_p
    lib.custen_device_free.argtypes, lib.custen_device_free.restype = [_c_void_p], None
    lib.custen_cahn_create.argtypes = [_c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _c_int]
    lib.custen_cahn_create.restype = ctypes.c_void_p
    lib.custen_cahn_set_field.argtypes, lib.custen_cahn_set_field.restype = [_c_void_p, _c_void_p], None
    lib.custen_cahn_step.argtypes, lib.custen_cahn_step.restype = [_c_void_p, _c_int], None
    lib.custen_cahn_get_field.argtypes, lib.custen_cahn_get_field.restype = [_c_void_p, _c_void_p], None
    lib.custen_cahn_time_steps.argtypes, lib.custen_cahn_time_steps.restype = [_c_void_p, _c_int], ctypes.c_float
    lib.custen_cahn_destroy.argtypes, lib.custen_cahn_destroy.restype = [_c_void_p], None
    lib.custen_cahn_set_table_rows.argtypes, lib.custen_cahn_set_table_rows.restype = [_c_int], None
    lib.custen_cahn_set_solver.argtypes, lib.custen_cahn_set_solver.restype = [_c_int], None
    lib.custen_cahn_set_fused.argtypes, lib.custen_cahn_set_fused.restype = [_c_int], None
    lib.custen_cahn_set_graph.argtypes, lib.custen_cahn_set_graph.restype = [_c_int], None
    lib.custen_cahn_slab_create.argtypes = [_c_int, _c_int, _c_int] + [ctypes.c_double] * 4 + [_c_int]
    lib.custen_cahn_slab_create.restype = ctypes.c_void_p
    lib.custen_cahn_slab_export.argtypes, lib.custen_cahn_slab_export.restype = [_c_void_p, _c_void_p], None
    lib.custen_cahn_slab_connect.argtypes, lib.custen_cahn_slab_connect.restype = [_c_void_p, _c_void_p, _c_void_p], None
    lib.custen_cahn_slab_connect_local.argtypes, lib.custen_cahn_slab_connect_local.restype = [_c_void_p] * 3, None
    lib.custen_cahn_slab_set_fields.argtypes, lib.custen_cahn_slab_set_fields.restype = [_c_void_p] * 3, None
    lib.custen_cahn_slab_step.argtypes, lib.custen_cahn_slab_step.restype = [_c_void_p, _c_int], None
    lib.custen_cahn_slab_time_steps.argtypes, lib.custen_cahn_slab_time_steps.restype = [_c_void_p, _c_int], ctypes.c_float
    lib.custen_cahn_slab_synchronize.argtypes, lib.custen_cahn_slab_synchronize.restype = [_c_void_p], None
    lib.custen_cahn_slab_error.argtypes, lib.custen_cahn_slab_error.restype = [_c_void_p], _c_int
    lib.custen_cahn_slab_set_timeout.argtypes, lib.custen_cahn_slab_set_timeout.restype = [_c_void_p, ctypes.c_double], None
    lib.custen_cahn_slab_set_graph.argtypes, lib.custen_cahn_slab_set_graph.restype = [_c_void_p, _c_int], None
    lib.custen_cahn_slab_partition_rows.argtypes, lib.custen_cahn_slab_partition_rows.restype = [_c_void_p], _c_int
    lib.custen_cahn_slab_destroy.argtypes, lib.custen_cahn_slab_destroy.restype = [_c_void_p], None
    lib.custen_cahn_mg_create.argtypes = [_c_int, _c_int, _c_void_p] + [ctypes.c_double] * 4
    lib.custen_cahn_mg_create.restype = ctypes.c_void_p
    lib.custen_cahn_mg_set_field.argtypes, lib.custen_cahn_mg_set_field.restype = [_c_void_p, _c_void_p], None
    lib.custen_cahn_mg_get_field.argtypes, lib.custen_cahn_mg_get_field.restype = [_c_void_p, _c_void_p], None
    lib.custen_cahn_mg_step.argtypes, lib.custen_cahn_mg_step.restype = [_c_void_p, _c_int], None
    lib.custen_cahn_mg_time_steps.argtypes, lib.custen_cahn_mg_time_steps.restype = [_c_void_p, _c_int], ctypes.c_float
    lib.custen_cahn_mg_error.argtypes, lib.custen_cahn_mg_error.restype = [_c_void_p], _c_int
    lib.custen_cahn_mg_set_graph.argtypes, lib.custen_cahn_mg_set_graph.restype = [_c_void_p, _c_int], None
    lib.custen_cahn_mg_destroy.argtypes, lib.custen_cahn_mg_destroy.restype = [_c_void_p], None
    lib.custen_cahn_set_fields.argtypes, lib.custen_cahn_set_fields.restype = [_c_void_p] * 3, None
    lib.custen_cahn_write_snapshot.argtypes = [_c_void_p, ctypes.c_char_p, ctypes.c_double]
    lib.custen_cahn_write_snapshot.restype = _c_int
    lib.custen_cahn_set_rhs_stream.argtypes, lib.custen_cahn_set_rhs_stream.restype = [_c_int], None
    lib.custen_cahn_dt.argtypes, lib.custen_cahn_dt.restype = [_c_void_p], ctypes.c_double
    lib.custen_cahn_slab_set_field.argtypes, lib.custen_cahn_slab_set_field.restype = [_c_void_p, _c_void_p], None
    lib.custen_cahn_slab_get_field.argtypes, lib.custen_cahn_slab_get_field.restype = [_c_void_p, _c_void_p], None
    lib.custen_debug_bands.argtypes = [_c_int] * 14 + [_c_void_p, _c_int]
    lib.custen_debug_bands.restype = _c_int
    lib.custen_managed_alloc.argtypes, lib.custen_managed_alloc.restype = [ctypes.c_size_t], ctypes.c_void_p
    lib.custen_managed_free.argtypes, lib.custen_managed_free.restype = [_c_void_p], None
    lib.custen_set_handle_managed_policy.argtypes, lib.custen_set_handle_managed_policy.restype = [_c_void_p, _c_int], None
    lib.custen_mem_advise.argtypes, lib.custen_mem_advise.restype = [_c_void_p, ctypes.c_size_t, _c_int, _c_int], _c_int
    lib.custen_mem_prefetch.argtypes, lib.custen_mem_prefetch.restype = [_c_void_p, ctypes.c_size_t, _c_int], _c_int
    lib.custen_fill_hash.argtypes = [_c_void_p, ctypes.c_longlong, _c_int, _c_int, ctypes.c_uint64, ctypes.c_double, ctypes.c_double]
    lib.custen_fill_hash.restype = None
    lib.custen_slab_create.argtypes = [_c_int] * 6 + [_c_void_p] + [_c_int] * 7 + [_c_void_p]
    lib.custen_slab_create.restype = ctypes.c_void_p
    lib.custen_slab_export.argtypes, lib.custen_slab_export.restype = [_c_void_p, _c_void_p, _c_void_p], None
    lib.custen_slab_connect.argtypes = [_c_void_p, _c_void_p, ctypes.c_size_t, _c_void_p, ctypes.c_size_t]
    lib.custen_slab_connect.restype = None
    lib.custen_slab_field.argtypes, lib.custen_slab_field.restype = [_c_void_p, _c_int], ctypes.c_void_p
    lib.custen_slab_rows.argtypes, lib.custen_slab_rows.restype = [_c_void_p], _c_int
    for name in ("compute", "swap", "synchronize", "destroy"):
        f = getattr(lib, f"custen_slab_{name}")
        f.argtypes, f.restype = [_c_void_p], None
    lib.custen_slab_run.argtypes, lib.custen_slab_run.restype = [_c_void_p, _c_int], None
    lib.custen_slab_run_plain.argtypes, lib.custen_slab_run_plain.restype = [_c_void_p, _c_int], None
    lib.custen_slab_time_run.argtypes, lib.custen_slab_time_run.restype = [_c_void_p, _c_int], ctypes.c_float
    lib.custen_slab_error.argtypes, lib.custen_slab_error.restype = [_c_void_p], _c_int
    lib.custen_slab_set_timeout.argtypes, lib.custen_slab_set_timeout.restype = [_c_void_p, ctypes.c_double], None
    lib.custen_slab_last_path.argtypes, lib.custen_slab_last_path.restype = [_c_void_p], _c_int
    lib.custen_mg_create.argtypes = [_c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p] + [_c_int] * 7 + [ctypes.c_char_p, _c_void_p]
    lib.custen_mg_create.restype = ctypes.c_void_p
    lib.custen_mg_scatter.argtypes, lib.custen_mg_scatter.restype = [_c_void_p, _c_void_p], None
    lib.custen_mg_gather.argtypes, lib.custen_mg_gather.restype = [_c_void_p, _c_void_p, _c_int], None
    lib.custen_mg_fill_output.argtypes, lib.custen_mg_fill_output.restype = [_c_void_p, ctypes.c_double], None
    for name in ("compute", "swap", "synchronize", "destroy"):
        f = getattr(lib, f"custen_mg_{name}")
        f.argtypes, f.restype = [_c_void_p], None
    lib.custen_mg_run.argtypes, lib.custen_mg_run.restype = [_c_void_p, _c_int], None
    lib.custen_mg_error.argtypes, lib.custen_mg_error.restype = [_c_void_p], _c_int
    lib.custen_mg_slab.argtypes, lib.custen_mg_slab.restype = [_c_void_p, _c_int], ctypes.c_void_p
    lib.custen_device_numa_node.argtypes, lib.custen_device_numa_node.restype = [_c_int], _c_int
    lib.custen_host_alloc_near.argtypes, lib.custen_host_alloc_near.restype = [ctypes.c_size_t, _c_int, _c_void_p], ctypes.c_void_p
    lib.custen_host_free_near.argtypes, lib.custen_host_free_near.restype = [_c_void_p, ctypes.c_size_t], None
    lib.custen_link_probe.argtypes = [_c_void_p, _c_void_p, ctypes.c_size_t, _c_int, _c_int]
    lib.custen_link_probe.restype = ctypes.c_float
    lib.custen_cahn_set_partition_rows.argtypes, lib.custen_cahn_set_partition_rows.restype = [_c_int], None
    lib.custen_cahn_config.argtypes, lib.custen_cahn_config.restype = [_c_void_p, _c_int, _c_int], _c_int
    lib.custen_pent_part_host.argtypes, lib.custen_pent_part_host.restype = [_c_int, _c_int, _c_void_p, _c_void_p, _c_void_p], _c_int
    lib.custen_pent_part_device.argtypes = [_c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int]
    lib.custen_pent_part_device.restype = _c_int
    lib.custen_pent_part_choose_np.argtypes, lib.custen_pent_part_choose_np.restype = [_c_int, _c_int], _c_int
    _lib = lib
    return lib


class BandDesc(ctypes.Structure):
    """Record written by custen_debug_bands (custen_b200/csrc/plan.h BandDesc)."""
    _fields_ = ([(n, ctypes.c_longlong) for n in ("in_off", "out_off", "top_off", "bottom_off")]
                + [(n, ctypes.c_int) for n in ("top_kind", "bottom_kind", "rows", "nx", "L", "R", "T", "B", "H", "V", "wrap_x",
                                               "xlo", "xhi", "ylo", "yhi", "zero_right", "contiguous")])


def debug_bands(variant, numTiles, nx, ny, H=1, L=0, R=0, V=1, T=0, B=0, merged=False, slab=None):
    """Bands Compute would launch (no CUDA).  slab = None or (is_first, is_last)."""
    lib = load()
    arr = (BandDesc * 64)()
    n = lib.custen_debug_bands(VARIANTS.index(variant), numTiles, nx, ny, H, L, R, V, T, B, int(merged),
                               0 if slab is None else 1, 0 if slab is None else int(slab[0]),
                               0 if slab is None else int(slab[1]), ctypes.addressof(arr), 64)
    return [arr[i] for i in range(n)]
