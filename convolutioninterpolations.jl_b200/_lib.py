"""ctypes binding of libcib200.so (the C-ABI declared in include/cib200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module raises,
and every compute entry point fails loudly when no CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (CIB200_SO_NAME: load an A/B build of the library, development only - see build.py)
SO_PATH = os.path.join(_HERE, os.environ.get("CIB200_SO_NAME", "libcib200.so"))

OK, EOOB, ENAN, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM = range(7)
F64, F32 = 0, 1
HOST, DEVICE = 0, 1
KERNEL_IDS = {"b5": 0, "b7": 1, "b9": 2, "b11": 3, "b13": 4, "a3": 5, "a4": 6, "a5": 7, "a7": 8}
BC_IDS = {"detect": 0, "poly": 1, "linear": 2, "quadratic": 3}
EXTRAP_IDS = {"throw": 0, "flat": 1, "line": 2}
SUBGRID_IDS = {"cubic": 0, "quintic": 1, "linear": 2}
EVAL_MODES = {"auto": 0, "gather": 1, "binned": 2}
STRATEGIES = {0: "none", 1: "gather", 2: "binned", 3: "cellrun", 4: "lanequery", 5: "cellmma", 6: "cellmma_block"}  # cib200_strategy

# every symbol include/cib200.h declares (tests assert the library exports all of them)
SYMBOLS = [
    "cib200_version", "cib200_last_error", "cib200_launch_count", "cib200_kernel_table",
    "cib200_build_coefs", "cib200_extend_knots", "cib200_itp_create", "cib200_itp_destroy",
    "cib200_itp_set_eval_mode", "cib200_itp_eval", "cib200_itp_eval_grid", "cib200_itp_scratch_bytes",
    "cib200_resample", "cib200_resample_sg", "cib200_smooth", "cib200_smooth_halfwidth", "cib200_host_register",
    "cib200_host_unregister", "cib200_itp_set_profiling", "cib200_itp_last_pass_ms", "cib200_itp_last_strategy",
    "cib200_build_coefs_device", "cib200_itp_get_coefs", "cib200_itp_eval_multi", "cib200_device_count",
    "cib200_set_device", "cib200_itpnd_create", "cib200_itpnd_destroy", "cib200_itpnd_eval", "cib200_itpnd_eval_grid",
    "cib200_itp_create_tables", "cib200_itpnd_create_tables", "cib200_resample_tables",
    "cib200_itgnd_create", "cib200_itgnd_destroy", "cib200_itgnd_eval", "cib200_itgnd_eval_grid",
    "cib200_itg_create", "cib200_itg_create_tables", "cib200_itg_destroy", "cib200_itg_eval", "cib200_itg_eval_grid", "cib200_itg_get",
]


class ItpDesc(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("dtype", C.c_int32), ("coef_shape", C.c_int64 * 3),
                ("coefs", C.c_void_p), ("knots", C.c_void_p * 3), ("h", C.c_double * 3),
                ("kernel", C.c_int32 * 3), ("deriv", C.c_int32 * 3), ("extrap", C.c_int32),
                ("perdim", C.c_int32), ("subgrid", C.c_int32 * 3)]


class CIB200Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


class DomainError(CIB200Error):
    """extrap=Throw and a query is outside the domain (ErrorException in the reference)."""


class InexactError(CIB200Error):
    """NaN / Inf coordinate (InexactError from floor(Int, x) in the reference)."""


class UnsupportedModeError(CIB200Error):
    """A mode that stays on the reference's own Julia path (never silently approximated here)."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for this path.")
    L = C.CDLL(SO_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.cib200_version.restype = i32
    L.cib200_last_error.restype = C.c_char_p
    L.cib200_launch_count.restype = i64
    L.cib200_kernel_table.restype = C.POINTER(dbl)
    L.cib200_kernel_table.argtypes = [i32, i32, C.POINTER(i32)]
    L.cib200_build_coefs.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.cib200_build_coefs_device.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.cib200_extend_knots.argtypes = [vp, i64, i32, vp]
    L.cib200_itp_create.argtypes = [C.POINTER(ItpDesc), C.POINTER(vp)]
    L.cib200_itp_create_tables.argtypes = [C.POINTER(ItpDesc), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp)]
    L.cib200_itp_destroy.argtypes = [vp]
    L.cib200_itp_get_coefs.argtypes = [vp, vp]
    L.cib200_itp_set_eval_mode.argtypes = [vp, i32]
    L.cib200_itp_eval.argtypes = [vp, i64, C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itp_eval_grid.argtypes = [vp, C.POINTER(i64), C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_device_count.restype = i32
    L.cib200_set_device.argtypes = [i32]
    L.cib200_itp_eval_multi.argtypes = [C.POINTER(vp), i32, i64, C.POINTER(vp), vp, C.POINTER(i64)]
    L.cib200_itpnd_create.argtypes = [i32, i32, vp, vp, C.POINTER(vp), vp, vp, vp, i32, i32, C.POINTER(vp)]
    L.cib200_itpnd_create_tables.argtypes = [i32, i32, vp, vp, C.POINTER(vp), vp, vp, vp, i32, i32, C.POINTER(vp),
                                             C.POINTER(i32), C.POINTER(vp)]
    L.cib200_itpnd_destroy.argtypes = [vp]
    L.cib200_itpnd_eval.argtypes = [vp, i64, C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itpnd_eval_grid.argtypes = [vp, C.POINTER(i64), C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itgnd_create.argtypes = [i32, vp, vp, C.POINTER(vp), vp, vp, vp, i32, C.POINTER(vp), vp, vp, C.POINTER(vp)]
    L.cib200_itgnd_destroy.argtypes = [vp]
    L.cib200_itgnd_eval.argtypes = [vp, i64, C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itgnd_eval_grid.argtypes = [vp, C.POINTER(i64), C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itg_create.argtypes = [C.POINTER(ItpDesc), C.POINTER(vp)]
    L.cib200_itg_create_tables.argtypes = [C.POINTER(ItpDesc), C.POINTER(vp), C.POINTER(i32), C.POINTER(vp)]
    L.cib200_itg_destroy.argtypes = [vp]
    L.cib200_itg_eval.argtypes = [vp, i64, C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itg_eval_grid.argtypes = [vp, C.POINTER(i64), C.POINTER(vp), vp, i32, vp, C.POINTER(i64)]
    L.cib200_itg_get.argtypes = [vp, i32, i32, vp]
    L.cib200_itp_set_profiling.argtypes = [vp, i32]
    L.cib200_itp_last_pass_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.cib200_itp_last_strategy.argtypes = [vp]
    L.cib200_itp_scratch_bytes.restype = i64
    L.cib200_itp_scratch_bytes.argtypes = [vp, i64]
    L.cib200_resample.argtypes = [i32, i32, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, i32, vp]
    L.cib200_resample_sg.argtypes = [i32, i32, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp, i32, vp]
    L.cib200_resample_tables.argtypes = [i32, i32, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp,
                                         C.POINTER(vp), C.POINTER(i32), i32, vp]
    L.cib200_smooth.argtypes = [i32, i32, vp, vp, vp, vp, dbl, i32, vp]
    L.cib200_smooth_halfwidth.restype = i32
    L.cib200_smooth_halfwidth.argtypes = [dbl]
    L.cib200_host_register.argtypes = [vp, C.c_size_t]
    L.cib200_host_unregister.argtypes = [vp]
    _lib = L
    return L


def check(status):
    if status == OK:
        return
    msg = lib().cib200_last_error().decode("utf-8", "replace")
    cls = {EOOB: DomainError, ENAN: InexactError, EUNSUPPORTED: UnsupportedModeError}.get(status, CIB200Error)
    raise cls(status, msg or f"cib200 status {status}")


def launch_count():
    return int(lib().cib200_launch_count())
