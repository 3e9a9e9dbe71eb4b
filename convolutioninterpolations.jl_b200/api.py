"""Host-side mirror of the reference's public API for the accelerated path.

Same names, keyword arguments, defaults and error behaviour as
  convolution_interpolation   src/convolution/convolution_interpolation_function.jl:79-157
  convolution_resample        src/convolution/convolution_resample.jl:65-238
  convolution_smooth          src/convolution/convolution_gaussian_smoothing.jl:45-131
(Julia symbols become strings: :b5 -> "b5"; `itp.(xs...)` becomes `itp.eval(xs...)`,
`itp(xs, ys)` with vectors is the Cartesian product exactly as extrapolation_operator.jl:66-99.)

This file only normalises arguments and resolves modes; all numerical work happens behind the
C-ABI (include/cib200.h) in libcib200.so.  Modes the CUDA path does not implement raise
UnsupportedModeError - in the Julia glue those calls keep dispatching to the reference's own
methods; nothing is silently approximated and nothing falls back to the CPU here.

Inputs may be numpy arrays (host buffers, copied through the library's pipelined H2D/D2H path) or
CUDA torch tensors (device-resident, zero-copy, asynchronous on torch's current stream).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (BC_IDS, EXTRAP_IDS, KERNEL_IDS, SUBGRID_IDS, CIB200Error, UnsupportedModeError, check, lib)

# data_structures/degree_to_equations.jl:13-25
DEGREE_TO_EQUATIONS = {"a0": 1, "a1": 1, "a3": 2, "a4": 3, "a5": 3, "a7": 4, "b5": 5, "b7": 6, "b9": 7,
                       "b11": 8, "b13": 9}
# fast_convolution_interpolation_constructor.jl:400-412
MAX_SMOOTH_DERIVATIVE = {"a0": -1, "a1": 0, "a3": 1, "a4": 1, "a5": 1, "a7": 1, "b5": 3, "b7": 4, "b9": 5,
                         "b11": 6, "b13": 6}
# get_precomputed_kernel_and_range.jl:33-40
MAX_SHIPPED_DERIVATIVE = {"a0": 0, "a1": 0, "a3": 0, "a4": 0, "a5": 0, "a7": 0, "b5": 2, "b7": 3, "b9": 4,
                          "b11": 5, "b13": 5}
# convolution_interpolation_function.jl:183-195
MINIMUM_POLYNOMIAL_BC_POINTS = {"a0": 2, "a1": 2, "a3": 4, "a4": 4, "a5": 4, "a7": 4, "b5": 6, "b7": 8,
                                "b9": 8, "b11": 8, "b13": 8}
# highest K^(d) table shipped per kernel (precomputed_kernels/kernel_<k>.jl): the quintic subgrid needs K^(d+2)
MAX_TABLE_DERIVATIVE = {"a3": 1, "a4": 1, "a5": 1, "a7": 1, "b5": 3, "b7": 4, "b9": 5, "b11": 6, "b13": 6}
ACCELERATED_KERNELS = tuple(KERNEL_IDS)
MAX_ND = 8  # CIB200_MAX_ND: point evaluation in 4..8 dimensions goes through cib200_itpnd_*


class AbstractExtrapolation:
    name = None

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self))

    def __repr__(self):
        return f"{type(self).__name__}()"


class Throw(AbstractExtrapolation):
    name = "throw"


class Flat(AbstractExtrapolation):
    name = "flat"


class Line(AbstractExtrapolation):
    name = "line"


class Natural(AbstractExtrapolation):
    name = "natural"


def _sym(s):
    if isinstance(s, str):
        return s[1:] if s.startswith(":") else s
    return s


def _extrap_type(s):
    """convolution_interpolation_function.jl:197-211"""
    if isinstance(s, AbstractExtrapolation):
        return s
    s = _sym(s)
    if s == "throw":
        return Throw()
    if s == "line":
        return Line()
    if s == "flat":
        return Flat()
    raise ValueError(f"Unknown extrapolation type: {s}, must be :throw, :line, :flat or :natural.")


def _ntuple(v, n, scalar, what):
    if isinstance(v, scalar):
        return tuple(_sym(v) for _ in range(n))
    v = tuple(_sym(x) for x in v)
    if len(v) == n:
        return v
    if len(v) == 1:
        return v * n
    raise ValueError(f"Invalid {what} specification: {v}.")


def _norm_bc(bc, n):
    if isinstance(bc, str):
        return tuple((_sym(bc), _sym(bc)) for _ in range(n))
    bc = tuple(bc)
    if len(bc) == 2 and all(isinstance(b, str) for b in bc):
        return tuple((_sym(bc[0]), _sym(bc[1])) for _ in range(n))
    if len(bc) == n and all(len(tuple(b)) == 2 for b in bc):
        return tuple((_sym(b[0]), _sym(b[1])) for b in bc)
    if len(bc) == 1:
        return tuple((_sym(bc[0][0]), _sym(bc[0][1])) for _ in range(n))
    raise ValueError(f"Invalid bc specification: {bc}.")


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _np_dtype(dtype_id):
    return np.float32 if dtype_id == _lib.F32 else np.float64


def _knots_tuple(knots, ndim, what="knots"):
    if ndim == 1 and not (isinstance(knots, (tuple, list)) and len(knots) == 1 and np.ndim(knots[0]) == 1):
        knots = (knots,)
    knots = tuple(knots)
    if len(knots) != ndim:
        raise ValueError(f"Invalid {what} specification: expected {ndim} knot vectors, got {len(knots)}.")
    return knots


def is_uniform_grid(k, rtol=1e-10):
    """convolution_kernels/nonuniform.jl:13-25"""
    h0 = k[1] - k[0]
    if h0 == 0:
        return False
    d = np.diff(k)
    return bool(np.all(np.abs(d - h0) <= rtol * abs(h0)))


def _build_subgrids(kernels, derivs, subgrids):
    """fast_convolution_interpolation_constructor.jl:456-494"""
    n = len(kernels)
    low = any(k in ("a0", "a1") for k in kernels)
    all_cubic = len(set(subgrids)) == 1 and subgrids[0] == "cubic"
    has_int = any(d == -1 for d in derivs)
    out = []
    for k, dv, sg in zip(kernels, derivs, subgrids):
        if dv == -1:
            if k in ("a0", "a1") or n >= 4:
                out.append("linear")
            elif k in ("a3", "a4", "a5", "a7"):
                out.append("cubic")
            else:
                out.append(sg)
        elif dv == 0:
            out.append("linear" if (low or (n == 3 and not all_cubic) or n >= 4 or has_int) else sg)
        else:
            avail = MAX_SMOOTH_DERIVATIVE.get(k, 0) - dv
            out.append("linear" if (avail <= 0 or (n == 3 and not all_cubic) or n >= 4 or low or has_int) else sg)
    return tuple(out)


class LinearSubgridTableWarning(UserWarning):
    """linear_rows=101 was requested: a :linear-subgrid axis is evaluated on the 101-row shipped table, whereas the
    reference generates >= 10 000 rows for such axes (precompute_actual = max(precompute, 10_000):
    fast_convolution_interpolation_constructor.jl:117-118, convolution_resample.jl:228-229) - the default here too.  The
    101-row result is ~1e-4 relative away from the reference's (DESIGN.md 1, row f)."""


def _warn_linear_rows(subgrids, kernels, what):
    if any(sg == "linear" and k not in ("a0", "a1") for sg, k in zip(subgrids, kernels)):
        import warnings
        warnings.warn(f"{what}: :linear subgrid axes {tuple(subgrids)} are evaluated on 101-row tables, the reference "
                      "uses >= 10 000 rows for them (agreement ~1e-5 relative, not 1e-12)", LinearSubgridTableWarning,
                      stacklevel=3)


def _effective_subgrids(subgrids, ndim):
    """What the evaluators do with a resolved subgrid tuple: in 3-D anything but all-cubic is evaluated with linear
    blending on every axis (eager_3d.jl:139-176; perdim_3d.jl:161 "fallback to all linear")."""
    if ndim == 3 and tuple(subgrids) != ("cubic",) * 3:
        return ("linear",) * 3
    return tuple(subgrids)


def _require_accelerated(kernels, derivs, subgrids, ndim, what, perdim=True):
    """Raise UnsupportedModeError for modes that stay on the reference's Julia path.  With precompute = 101 the tables
    of every subgrid are the shipped K^(d) (the generated ones of the :linear path are the same exact rationals rounded
    once): :linear needs K^(d), :cubic also K^(d+1), :quintic also K^(d+2)."""
    if ndim > MAX_ND and what == "convolution_interpolation":
        raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: N={ndim} > {MAX_ND} is not on the accelerated path")
    for d, (k, dv, sg) in enumerate(zip(kernels, derivs, subgrids)):
        if k not in KERNEL_IDS:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: kernel :{k} (axis {d + 1}) is not on the "
                                       f"accelerated path ({ACCELERATED_KERNELS})")
        if dv < 0:
            # integral direction: FastIntegralOrder (every axis, convolution_fast_integration_{1d,2d,3d}.jl) or
            # FastMixedIntegralOrder (some axes, convolution_fast_mixed_*.jl) on the shipped K~ / K / K' tables; the other
            # axes of a mixed interpolant are ordinary :linear axes (checked below).  N >= 4 resolves the integral
            # directions to :linear K~ tables and has no tails: reference path.
            if what != "convolution_interpolation":
                raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what} does not support antiderivatives")
            if dv != -1:
                raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: derivative={tuple(derivs)} is not a valid order")
            if ndim > 3:
                # N >= 4: every axis on the :linear subgrid with generated tables, three functors
                # (convolution_fast_integration_nd.jl, convolution_fast_mixed_3_integral_dims.jl,
                # convolution_fast_mixed_N_integral_dims.jl): cib200_itgnd_*.  One or two integral axes in N >= 4
                # dimensions (the N-D instances of convolution_fast_mixed_{1_integral_dim,2_integral_dims}.jl) stay on
                # the reference path.
                n_int = sum(1 for x in derivs if x == -1)
                if n_int < 3:
                    raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: derivative={tuple(derivs)}: {n_int} integral "
                                               f"axes in {ndim} dimensions are not on the accelerated path (3 or more are)")
                if sg != "linear":
                    raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: N >= 4 antiderivatives use the :linear subgrid")
                continue
            if sg not in ("cubic", "quintic"):
                raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: antiderivative with the :{sg} subgrid needs "
                                           "generated tables")
            continue
        if sg not in ("cubic", "quintic", "linear"):
            raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: unknown subgrid {subgrids}")
        need = dv + {"linear": 0, "cubic": 1, "quintic": 2}[sg]
        if need > MAX_TABLE_DERIVATIVE[k]:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: derivative {dv} of :{k} with the :{sg} subgrid needs "
                                       f"K^({need}), which is not a shipped table")
    if what == "convolution_interpolation" and not any(dv == -1 for dv in derivs):
        # (the antiderivative evaluators take any per-axis mix of cubic / quintic: eval_sg_k, eval_subgrid_helpers.jl:2-19)
        # eager_2d.jl handles (linear, linear), (cubic, cubic), (quintic, quintic) only; the mixed pairs exist on the
        # per-dim path (perdim_2d.jl); N = 3 never resolves to quintic (_build_subgrids)
        if "quintic" in subgrids and ndim > 2:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: subgrid {tuple(subgrids)} has no 3-D evaluator")
        if ndim == 2 and not perdim and len(set(subgrids)) != 1:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, f"{what}: subgrid {tuple(subgrids)} has no evaluator for "
                                       "one kernel and one derivative order on all axes (eager_2d.jl:91-311)")


def _stream_ptr(device_tensor):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device_tensor.device).cuda_stream)


class FastConvolutionInterpolation:
    """Numeric content of the reference struct (data_structures/fast_convolution_interpolation.jl:56-102)
    for the eager cubic-subgrid case, plus the device handle."""

    def __init__(self, knots, values, kernel="b5", precompute=101, bc="detect", derivative=0, subgrid="cubic",
                 lazy=False, boundary_fallback=False, dtype=None, eval_mode="auto", _extrap="throw", coefs_on="host",
                 devices=None, linear_rows=None):
        values = np.asarray(values)
        if values.dtype not in (np.float32, np.float64):
            values = values.astype(np.float64)
        if dtype is not None:
            values = values.astype(np.dtype(dtype))
        nd = values.ndim
        self.ndim = nd
        self.dtype = values.dtype
        self.dtype_id = _lib.F32 if values.dtype == np.float32 else _lib.F64
        knots = _knots_tuple(knots, nd)
        # T.(knots): knots are converted to the element type (constructor.jl:60-62)
        self.knots_in = tuple(np.asarray(k).astype(self.dtype).astype(np.float64) for k in knots)
        for d in range(nd):
            if len(self.knots_in[d]) != values.shape[d]:
                raise ValueError(f"knots[{d}] has {len(self.knots_in[d])} points but values has {values.shape[d]}")
        self.kernel = _ntuple(kernel, nd, str, "kernel")
        self.derivative = _ntuple(derivative, nd, (int, np.integer), "derivative")
        sg_req = _ntuple(subgrid, nd, str, "subgrid")
        self.bc = _norm_bc(bc, nd)
        if any(k == "n3" for k in self.kernel):
            raise ValueError("The :n3 kernel is not supported by FastConvolutionInterpolation.")
        if lazy:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, "lazy=true stays on the reference path (ghost-padded grids "
                                       "are built once on the host and uploaded)")
        for k in self.kernel:
            if k not in DEGREE_TO_EQUATIONS:
                raise ValueError(f"Unknown kernel :{k}")
        self.subgrid = _build_subgrids(self.kernel, self.derivative, sg_req)
        if precompute != 101:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, "precompute != 101 needs generated tables (reference path)")
        # _build_kernel_sym (constructor.jl:378-398): HigherOrderKernel vs HigherOrderMixedKernel
        self.perdim = not (len(set(self.kernel)) == 1 and len(set(self.derivative)) == 1)
        self._itg = any(d == -1 for d in self.derivative)
        if not self._itg:
            self.subgrid = _effective_subgrids(self.subgrid, nd)
        _require_accelerated(self.kernel, self.derivative, self.subgrid, nd, "convolution_interpolation", self.perdim)
        # :linear axes: precompute_actual = max(precompute, 10_000) rows, generated from exact rationals
        # (constructor.jl:117-118); linear_rows=101 evaluates them on the shipped tables instead (NOT the reference's
        # result: ~1e-4 relative away; kept for the kernels' own tests and for comparisons)
        from . import tables as _tables
        self.linear_rows = _tables.LINEAR_ROWS if linear_rows is None else int(linear_rows)
        self._gen_tables = [None] * nd
        if self.linear_rows != 101:
            for d in range(nd):
                if self.subgrid[d] == "linear" and self.derivative[d] >= 0:
                    self._gen_tables[d] = np.asfortranarray(
                        _tables.generated_table(self.kernel[d], int(self.derivative[d]), self.linear_rows))
        else:
            _warn_linear_rows(self.subgrid, self.kernel, "convolution_interpolation")
        if self._itg and self.dtype_id != _lib.F64:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, "the antiderivative path is Float64 only")
        self.eqs = tuple(DEGREE_TO_EQUATIONS[k] for k in self.kernel)
        self.h = tuple(float(k[1] - k[0]) for k in self.knots_in)  # constructor.jl:120
        L = lib()
        # ---- subsystem 1: ghost-padded coefficients + extended knots, on the host ----
        v = np.asfortranarray(values)
        cshape = tuple(s + 2 * (e - 1) for s, e in zip(v.shape, self.eqs))
        n = (C.c_int64 * nd)(*v.shape)
        hh = (C.c_double * nd)(*self.h)
        kid = (C.c_int32 * nd)(*[KERNEL_IDS[k] for k in self.kernel])
        bl = (C.c_int32 * nd)(*[BC_IDS[b[0]] for b in self.bc])
        br = (C.c_int32 * nd)(*[BC_IDS[b[1]] for b in self.bc])
        if coefs_on not in ("host", "device"):
            raise ValueError("coefs_on must be 'host' or 'device'")
        self.coefs_on = coefs_on
        self._d_coefs = None
        if coefs_on == "device":
            # values go up once, the ghost fill runs on the GPU (cib200_build_coefs_device) and the
            # padded grid never visits the host unless .coefs is read
            import torch
            tdt = torch.float32 if self.dtype_id == _lib.F32 else torch.float64
            d_vals = torch.from_numpy(v.reshape(-1, order="F").copy()).cuda()  # column-major image of the values
            self._d_coefs = torch.empty(int(np.prod(cshape)), dtype=tdt, device=d_vals.device)
            check(L.cib200_build_coefs_device(nd, self.dtype_id, n, C.c_void_p(d_vals.data_ptr()), hh, kid, bl, br,
                                              C.c_void_p(self._d_coefs.data_ptr()), _lib.DEVICE, _stream_ptr(d_vals)))
            torch.cuda.current_stream(d_vals.device).synchronize()
            del d_vals
            self._coefs = None
            self._cshape = cshape
            coefs_ptr = self._d_coefs.data_ptr()
        else:
            self._coefs = np.zeros(cshape, dtype=self.dtype, order="F")
            check(L.cib200_build_coefs(nd, self.dtype_id, n, v.ctypes.data, hh, kid, bl, br, self._coefs.ctypes.data))
            coefs_ptr = self._coefs.ctypes.data
        self.knots = []
        for d in range(nd):
            ext = np.zeros(cshape[d])
            check(L.cib200_extend_knots(self.knots_in[d].ctypes.data, len(self.knots_in[d]), self.eqs[d] - 1,
                                        ext.ctypes.data))
            self.knots.append(ext)
        self.knots = tuple(self.knots)
        # ---- upload ----
        if nd >= 4 and self._itg:
            # antiderivative / mixed interpolants in N >= 4 dimensions: convolution_fast_integration_nd.jl (every axis
            # integral: its loops stop at domain_size[d]), convolution_fast_mixed_3_integral_dims.jl,
            # convolution_fast_mixed_N_integral_dims.jl - cib200_itgnd_*
            self._nd = True
            self._itgnd = True
            rows = self.linear_rows
            tabs = [np.asfortranarray(_tables.generated_table(self.kernel[d], int(self.derivative[d]), rows))
                    if rows != 101 else None for d in range(nd)]
            self._gen_tables = tabs
            cs = (C.c_int64 * nd)(*cshape)
            kp = (C.c_void_p * nd)(*[k.ctypes.data for k in self.knots])
            dv = (C.c_int32 * nd)(*[int(d) for d in self.derivative])
            tp = (C.c_void_p * nd)(*[None if t is None else t.ctypes.data for t in tabs])
            npre = (C.c_int32 * nd)(*[0 if t is None else t.shape[0] for t in tabs])
            all_int = all(d == -1 for d in self.derivative)
            ext = (C.c_int64 * nd)(*(v.shape if all_int else cshape))
            handle = C.c_void_p()
            if coefs_on == "device":
                coefs_ptr = self.coefs.ctypes.data
            check(L.cib200_itgnd_create(nd, cs, coefs_ptr, kp, hh, kid, dv, EXTRAP_IDS[_extrap], tp, npre, ext,
                                        C.byref(handle)))
            self._handle = handle
            self._handles = [handle]
            self.devices = None
            self.eval_mode = "integral-nd"
            self.anchor = tuple(float(k[e - 1]) if d == -1 else 0.0 for k, e, d in zip(self.knots, self.eqs, self.derivative))
            return
        if nd >= 4:
            # HigherDimension evaluators (eager_nd.jl:97-140 / perdim_nd.jl:4-36): every axis on the :linear subgrid
            self._nd = True
            cs = (C.c_int64 * nd)(*cshape)
            kp = (C.c_void_p * nd)(*[k.ctypes.data for k in self.knots])
            dv = (C.c_int32 * nd)(*[int(d) for d in self.derivative])
            handle = C.c_void_p()
            if coefs_on == "device":
                coefs_ptr = self.coefs.ctypes.data  # the N-D constructor copies from host memory
            tp = (C.c_void_p * nd)(*[None if t is None else t.ctypes.data for t in self._gen_tables])
            npre = (C.c_int32 * nd)(*[0 if t is None else t.shape[0] for t in self._gen_tables])
            check(L.cib200_itpnd_create_tables(nd, self.dtype_id, cs, coefs_ptr, kp, hh, kid, dv,
                                               EXTRAP_IDS[_extrap], int(self.perdim), tp, npre, C.byref(handle)))
            if devices is not None and len(devices) > 1:
                raise UnsupportedModeError(_lib.EUNSUPPORTED, "devices=[...] (single-process multi-GPU) covers 1-3 dimensions")
            self._handle = handle
            self._handles = [handle]
            self.devices = None
            self.eval_mode = "nd"
            return
        self._nd = False
        desc = _lib.ItpDesc()
        desc.ndim = nd
        desc.dtype = self.dtype_id
        for d in range(nd):
            desc.coef_shape[d] = cshape[d]
            desc.knots[d] = self.knots[d].ctypes.data
            desc.h[d] = self.h[d]
            desc.kernel[d] = KERNEL_IDS[self.kernel[d]]
            desc.deriv[d] = int(self.derivative[d])
            desc.subgrid[d] = SUBGRID_IDS[self.subgrid[d]]
        desc.coefs = coefs_ptr
        desc.extrap = EXTRAP_IDS[_extrap]
        desc.perdim = int(self.perdim)
        if self._itg:
            # antiderivative on all or some axes: prefix-sum tails along the integral axes, built on the device
            # (constructor.jl:176-364)
            if devices is not None and len(devices) > 1:
                raise UnsupportedModeError(_lib.EUNSUPPORTED, "devices=[...] is not available for antiderivatives")
            handle = C.c_void_p()
            tp = (C.c_void_p * 3)(*[None if t is None else t.ctypes.data for t in self._gen_tables] + [None] * (3 - nd))
            npre = (C.c_int32 * 3)(*[0 if t is None else t.shape[0] for t in self._gen_tables] + [0] * (3 - nd))
            check(L.cib200_itg_create_tables(C.byref(desc), tp, npre, C.byref(handle)))
            self._handle = handle
            self._handles = [handle]
            self.devices = None
            self.eval_mode = "integral"
            # itp.anchor (constructor.jl:139): the result is zero at the first interior knot of each axis
            self.anchor = tuple(float(k[e - 1]) if dv == -1 else 0.0
                                for k, e, dv in zip(self.knots, self.eqs, self.derivative))
            return
        # one handle per device (the grid is replicated, query batches are split by cib200_itp_eval_multi);
        # devices=None: the current device only
        self._handles = []
        tp = (C.c_void_p * 3)(*[None if t is None else t.ctypes.data for t in self._gen_tables] + [None] * (3 - nd))
        npre = (C.c_int32 * 3)(*[0 if t is None else t.shape[0] for t in self._gen_tables] + [0] * (3 - nd))
        if devices is None:
            handle = C.c_void_p()
            check(L.cib200_itp_create_tables(C.byref(desc), tp, npre, C.byref(handle)))
            self._handles.append(handle)
        else:
            import torch
            if not devices:
                raise ValueError("devices must name at least one CUDA device")
            for dev in devices:
                with torch.cuda.device(int(dev)):
                    handle = C.c_void_p()
                    check(L.cib200_itp_create_tables(C.byref(desc), tp, npre, C.byref(handle)))
                    self._handles.append(handle)
        self.devices = None if devices is None else tuple(int(d) for d in devices)
        self._handle = self._handles[0]
        self._d_coefs = None  # the handles hold their own copies
        self.set_eval_mode(eval_mode)

    @property
    def coefs(self):
        """itp.coefs (ghost-padded grid); with coefs_on="device" it is downloaded on first access"""
        if self._coefs is None:
            import torch
            if self._d_coefs is not None:
                self._coefs = self._d_coefs.cpu().numpy().reshape(self._cshape, order="F")
            else:
                self._coefs = np.zeros(self._cshape, dtype=self.dtype, order="F")
                check(lib().cib200_itp_get_coefs(self._handle, self._coefs.ctypes.data))
        return self._coefs

    def left_values(self, d):
        """itp.left_values[d + 1] of an antiderivative interpolant (constructor.jl:414-448)"""
        out = np.empty(self.knots[d].size)
        check(lib().cib200_itg_get(self._handle, 0, d, out.ctypes.data))
        return out

    def tail(self, mask):
        """prefix-sum array saturated on the axes of the bit mask (see cib200_itg_get): tail1_left / tail2_ll / ..."""
        out = np.empty(tuple(k.size for k in self.knots), order="F")
        check(lib().cib200_itg_get(self._handle, 1, int(mask), out.ctypes.data))
        return out

    def _entry(self, name):
        """the C-ABI entry point of this interpolant's family: cib200_itp_* / cib200_itpnd_* / cib200_itg_*"""
        fam = "itgnd" if getattr(self, "_itgnd", False) else ("itg" if getattr(self, "_itg", False) else ("itpnd" if self._nd else "itp"))
        return getattr(lib(), f"cib200_{fam}_{name}")

    def set_eval_mode(self, mode):
        if self._nd or self._itg:
            return  # one strategy in N >= 4 dimensions and for antiderivatives
        for h in self._handles:
            check(lib().cib200_itp_set_eval_mode(h, _lib.EVAL_MODES[mode]))
        self.eval_mode = mode

    def set_profiling(self, enable=True):
        """Measurement hook: bracket the passes of every evaluation with CUDA events (cib200_itp_set_profiling)."""
        if self._nd or self._itg:
            raise UnsupportedModeError(_lib.EUNSUPPORTED, "pass timing exists for 1-3 dimensional interpolation only")
        check(lib().cib200_itp_set_profiling(self._handle, 1 if enable else 0))

    def last_pass_ms(self):
        """Device time of the most recent profiled evaluation: {binning, cell sort, evaluation kernel, spill} in ms."""
        ms = (C.c_float * 4)()
        check(lib().cib200_itp_last_pass_ms(self._handle, ms))
        return dict(zip(("binning", "cellsort", "eval_kernel", "spill"), (float(v) for v in ms)))

    def last_strategy(self):
        """Which evaluation kernel the most recent 3-D call ran: gather | binned | cellrun | lanequery | cellmma
        (cib200_itp_last_strategy; measurement / test hook)."""
        return _lib.STRATEGIES[int(lib().cib200_itp_last_strategy(self._handle))]

    def close(self):
        hs = getattr(self, "_handles", None) or ([self._handle] if getattr(self, "_handle", None) else [])
        self._handle, self._handles = None, []
        for h in hs:
            self._entry("destroy")(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConvolutionExtrapolation:
    """data_structures/convolution_extrapolation.jl:28-31 + the call operators of
    convolution_extrapolate/extrapolation_operator.jl."""

    def __init__(self, itp, et):
        self.itp = itp
        self.et = et

    # -- itp.(xs, ys, zs): element-wise over equally shaped coordinate arrays --------------------
    def eval(self, *xs, out=None, check_finite=False):
        """itp.(xs...).  numpy arrays (host buffers): synchronous, a NaN / Inf coordinate raises InexactError and - with
        extrap=Throw - a query outside the domain DomainError, as the reference does.  CUDA torch tensors: the call is
        stream-ordered and returns without a host synchronisation; with extrap=Throw the status is fetched (and raised) as
        for host buffers, with Flat / Line a NaN / Inf coordinate yields NaN in the result unless `check_finite=True` asks the
        library to fetch the status word (one synchronisation) and raise like the host path."""
        itp = self.itp
        if len(xs) != itp.ndim:
            raise ValueError(f"expected {itp.ndim} coordinate arrays, got {len(xs)}")
        L = lib()
        bad = C.c_int64(-1)
        if _is_torch(xs[0]):
            import torch
            tdt = torch.float32 if itp.dtype_id == _lib.F32 else torch.float64
            xs = [x.to(tdt).contiguous() for x in xs]
            if not all(x.is_cuda for x in xs):
                raise ValueError("torch inputs must be CUDA tensors (use numpy arrays for host buffers)")
            shape = xs[0].shape
            nq = xs[0].numel()
            if any(x.numel() != nq for x in xs):
                raise ValueError("coordinate tensors must have the same number of elements")
            if out is None:
                out = torch.empty(shape, dtype=tdt, device=xs[0].device)
            ptrs = (C.c_void_p * itp.ndim)(*[x.data_ptr() for x in xs])
            check(itp._entry("eval")(itp._handle, nq, ptrs, C.c_void_p(out.data_ptr()), _lib.DEVICE, _stream_ptr(xs[0]),
                                     C.byref(bad) if check_finite else None))
            return out
        ndt = _np_dtype(itp.dtype_id)
        xs = np.broadcast_arrays(*[np.asarray(x, dtype=ndt) for x in xs])
        shape = xs[0].shape
        xs = [np.ascontiguousarray(x).reshape(-1) for x in xs]
        nq = xs[0].size
        if out is None:
            out = np.empty(nq, dtype=ndt)
        ptrs = (C.c_void_p * itp.ndim)(*[x.ctypes.data for x in xs])
        if len(itp._handles) > 1:
            hs = (C.c_void_p * len(itp._handles))(*[h.value for h in itp._handles])
            check(L.cib200_itp_eval_multi(hs, len(itp._handles), nq, ptrs, out.ctypes.data, C.byref(bad)))
            return out.reshape(shape)
        check(itp._entry("eval")(itp._handle, nq, ptrs, out.ctypes.data, _lib.HOST, None, C.byref(bad)))
        return out.reshape(shape)

    # -- itp(x, y, z): scalars, or vectors -> Cartesian product ---------------------------------
    def __call__(self, *xs):
        itp = self.itp
        if len(xs) != itp.ndim:
            raise ValueError(f"expected {itp.ndim} coordinates, got {len(xs)}")
        if all(np.ndim(x) == 0 and not _is_torch(x) for x in xs):
            return float(self.eval(*[np.array([x]) for x in xs])[0])
        L = lib()
        bad = C.c_int64(-1)
        ndt = _np_dtype(itp.dtype_id)
        vs = [np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=ndt))) for x in xs]
        if any(v.ndim != 1 for v in vs):
            raise ValueError("itp(xs...) takes scalars or vectors; use itp.eval(...) for element-wise evaluation")
        n = (C.c_int64 * itp.ndim)(*[v.size for v in vs])
        out = np.zeros(tuple(v.size for v in vs), dtype=ndt, order="F")
        ptrs = (C.c_void_p * itp.ndim)(*[v.ctypes.data for v in vs])
        check(itp._entry("eval_grid")(itp._handle, n, ptrs, out.ctypes.data, _lib.HOST, None, C.byref(bad)))
        return out


def convolution_interpolation(knots, values, kernel="b5", fast=True, precompute=101, extrap=None, bc="detect",
                              derivative=0, subgrid="cubic", lazy=False, boundary_fallback=False, dtype=None,
                              eval_mode="auto", coefs_on="host", devices=None, linear_rows=None):
    """convolution_interpolation(knots, values; kernel, fast, precompute, extrap, bc, derivative, subgrid,
    lazy, boundary_fallback)  -  convolution_interpolation_function.jl:79-157"""
    values = np.asarray(values)
    nd = values.ndim
    knots_t = _knots_tuple(knots, nd)
    kernels = _ntuple(kernel, nd, str, "kernel")
    derivs = _ntuple(derivative, nd, (int, np.integer), "derivative")
    bcs = _norm_bc(bc, nd)
    extrap = Throw() if extrap is None else extrap
    is_integral = any(d == -1 for d in derivs)
    # is_uniform_grid runs on T.(knots) with rtol 1e-5 for Float32, 1e-10 otherwise (nonuniform.jl:13-25)
    vdt = np.dtype(dtype) if dtype is not None else (values.dtype if values.dtype == np.float32 else np.dtype(np.float64))
    rtol = 1e-5 if vdt == np.float32 else 1e-10
    is_nonuniform = any(not is_uniform_grid(np.asarray(k).astype(vdt).astype(np.float64), rtol) for k in knots_t) or \
        any(k == "n3" for k in kernels)
    if is_nonuniform:
        raise UnsupportedModeError(_lib.EUNSUPPORTED, "non-uniform grids stay on the reference path (fast=false)")
    if is_integral and lazy and fast:
        raise ValueError("Lazy mode not supported in fast mode for antiderivatives.")  # :114-117
    if not fast:
        raise UnsupportedModeError(_lib.EUNSUPPORTED, "fast=false (direct kernel evaluation) stays on the reference path")
    # :poly -> :linear when a dimension has too few points (:135-143)
    fixed = []
    for d in range(nd):
        b = bcs[d]
        if "poly" in b and values.shape[d] < MINIMUM_POLYNOMIAL_BC_POINTS.get(kernels[d], 2):
            b = ("linear", "linear") if b == ("poly", "poly") else (("linear", b[1]) if b[0] == "poly" else (b[0], "linear"))
        fixed.append(b)
    if isinstance(extrap, Natural) or _sym(extrap) == "natural":
        raise UnsupportedModeError(_lib.EUNSUPPORTED, "extrap=:natural is not on the accelerated path")
    et = _extrap_type(extrap)
    itp = FastConvolutionInterpolation(knots_t, values, kernel=kernels, precompute=precompute, bc=tuple(fixed),
                                       derivative=derivs, subgrid=subgrid, lazy=lazy,
                                       boundary_fallback=boundary_fallback, dtype=dtype, eval_mode=eval_mode,
                                       _extrap=et.name, coefs_on=coefs_on, devices=devices, linear_rows=linear_rows)
    return ConvolutionExtrapolation(itp, et)


def _sweep_io(values, dtype):
    """returns (dtype_id, memkind, in_ptr, make_out(shape), stream, keepalive)"""
    if _is_torch(values):
        import torch
        if not values.is_cuda:
            raise ValueError("torch inputs must be CUDA tensors (use numpy arrays for host buffers)")
        tdt = values.dtype if values.dtype in (torch.float32, torch.float64) else torch.float64
        if dtype is not None:
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
        # Julia order = column-major: a torch tensor indexed [i1, i2, ...] is laid out with axis 0 contiguous
        v = values.to(tdt)
        vt = v.permute(*reversed(range(v.dim()))).contiguous()  # memory == column-major of v
        did = _lib.F32 if tdt == torch.float32 else _lib.F64

        def make_out(shape):
            o = torch.empty(tuple(reversed(shape)), dtype=tdt, device=values.device)
            return o, C.c_void_p(o.data_ptr()), lambda: o.permute(*reversed(range(o.dim())))
        return did, _lib.DEVICE, C.c_void_p(vt.data_ptr()), make_out, _stream_ptr(values), vt
    v = np.asarray(values)
    if v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float64)
    if dtype is not None:
        v = v.astype(np.dtype(dtype))
    v = np.asfortranarray(v)
    did = _lib.F32 if v.dtype == np.float32 else _lib.F64

    def make_out(shape):
        o = np.zeros(shape, dtype=v.dtype, order="F")
        return o, C.c_void_p(o.ctypes.data), lambda: o
    return did, _lib.HOST, C.c_void_p(v.ctypes.data), make_out, None, v


_UNSET = object()


def convolution_resample(knots_in, knots_out, values, precompute=_UNSET, kernel=_UNSET, derivative=_UNSET,
                         bc=_UNSET, subgrid=_UNSET, dtype=None, linear_rows=None):
    """convolution_resample(knots_in, knots_out, values; precompute, kernel, derivative, bc, subgrid)
    convolution_resample.jl:65-238 (front-ends) -> :240-386 (internal, replaced by cib200_resample).

    As in Julia, a call WITHOUT keyword arguments dispatches to the kw-less wrappers (:65-172), which pass
    :b13 / :cubic / 101 / :detect straight to the internal routine for every N <= 10; a call WITH keyword
    arguments goes through the kw method (:174-238), which runs _build_subgrids (N >= 4 -> :linear)."""
    kwless = all(a is _UNSET for a in (precompute, kernel, derivative, bc, subgrid))
    precompute = 101 if precompute is _UNSET else precompute
    kernel = "b13" if kernel is _UNSET else kernel
    derivative = 0 if derivative is _UNSET else derivative
    bc = "detect" if bc is _UNSET else bc
    subgrid = "cubic" if subgrid is _UNSET else subgrid
    nd = values.dim() if _is_torch(values) else np.ndim(values)
    shape_in = tuple(values.shape)
    kin = _knots_tuple(knots_in, nd, "knots_in")
    kout = _knots_tuple(knots_out, nd, "knots_out")
    kernels = _ntuple(kernel, nd, str, "kernel")
    derivs = _ntuple(derivative, nd, (int, np.integer), "derivative")
    sgs = _ntuple(subgrid, nd, str, "subgrid")
    bcs = _norm_bc(bc, nd)
    for d in range(nd):  # :213-224
        if sgs[d] != "linear" and kernels[d] in ("a0", "a1"):
            raise ValueError(f"Cannot use {sgs[d]} subgrid with {kernels[d]} kernel. Use :linear instead.")
    for d in range(nd):
        if derivs[d] < 0:
            raise ValueError("convolution_resample does not support antiderivatives (derivative=-1). "
                             "Use convolution_interpolation with derivative=-1 instead.")
    if precompute != 101:
        raise UnsupportedModeError(_lib.EUNSUPPORTED, "precompute != 101 needs generated tables (reference path)")
    sgs_valid = sgs if kwless else _build_subgrids(kernels, derivs, sgs)  # :226 (kw method only)
    if kwless and nd > 10:
        raise TypeError("no method matching convolution_resample for N > 10 without keyword arguments")
    _require_accelerated(kernels, derivs, sgs_valid, nd, "convolution_resample")
    # kw method: precompute_actual = max(precompute, 10_000) rows for :linear axes (:228-229), generated from exact
    # rationals; the kw-less wrappers pass 101 straight through.  linear_rows=101 keeps the shipped tables (not the
    # reference's result, ~1e-6 .. 1e-4 relative away)
    from . import tables as _tables
    rows = _tables.LINEAR_ROWS if linear_rows is None else int(linear_rows)
    gen = [None] * nd
    if not kwless and rows != 101:
        for d in range(nd):
            if sgs_valid[d] == "linear" and kernels[d] not in ("a0", "a1"):
                gen[d] = np.asfortranarray(_tables.generated_table(kernels[d], int(derivs[d]), rows))
    elif not kwless:
        _warn_linear_rows(sgs_valid, kernels, "convolution_resample")
    did, memkind, in_ptr, make_out, stream, keep = _sweep_io(values, dtype)
    ndt = _np_dtype(did)
    # T.(knots): :183-188
    kin = [np.ascontiguousarray(np.asarray(k).astype(ndt).astype(np.float64)) for k in kin]
    kout = [np.ascontiguousarray(np.asarray(k).astype(ndt).astype(np.float64)) for k in kout]
    for d in range(nd):
        if len(kin[d]) != shape_in[d]:
            raise ValueError(f"knots_in[{d}] has {len(kin[d])} points but values has {shape_in[d]}")
    shape_out = tuple(len(k) for k in kout)
    out, out_ptr, view = make_out(shape_out)
    n_in = (C.c_int64 * nd)(*shape_in)
    n_out = (C.c_int64 * nd)(*shape_out)
    kinp = (C.c_void_p * nd)(*[k.ctypes.data for k in kin])
    koutp = (C.c_void_p * nd)(*[k.ctypes.data for k in kout])
    kid = (C.c_int32 * nd)(*[KERNEL_IDS[k] for k in kernels])
    dv = (C.c_int32 * nd)(*[int(d) for d in derivs])
    bl = (C.c_int32 * nd)(*[BC_IDS[b[0]] for b in bcs])
    br = (C.c_int32 * nd)(*[BC_IDS[b[1]] for b in bcs])
    sg = (C.c_int32 * nd)(*[SUBGRID_IDS[x] for x in sgs_valid])
    tp = (C.c_void_p * nd)(*[None if t is None else t.ctypes.data for t in gen])
    npre = (C.c_int32 * nd)(*[0 if t is None else t.shape[0] for t in gen])
    check(lib().cib200_resample_tables(nd, did, n_in, n_out, in_ptr, out_ptr, kinp, koutp, kid, dv, bl, br, sg, tp, npre,
                                       memkind, stream))
    del keep
    return view()


def convolution_smooth(knots, values, B, dtype=None):
    """convolution_smooth(knots, values, B) - convolution_gaussian_smoothing.jl:45-131"""
    nd = values.dim() if _is_torch(values) else np.ndim(values)
    shape = tuple(values.shape)
    ks = _knots_tuple(knots, nd)
    B = float(B)
    did, memkind, in_ptr, make_out, stream, keep = _sweep_io(values, dtype)
    ndt = _np_dtype(did)
    E = lib().cib200_smooth_halfwidth(B)
    for d in range(nd):  # :50-56
        n = len(ks[d])
        if E >= n:
            raise ValueError(f"Gaussian stencil half-width ({E} points) exceeds grid size (n={n}) in dimension {d + 1}. "
                             f"Use a larger B (currently B={B}) or more grid points (need n > {E}).")
        if n != shape[d]:
            raise ValueError(f"knots[{d}] has {n} points but values has {shape[d]}")
    h = (C.c_double * nd)(*[float(ndt(k[1]) - ndt(k[0])) for k in ks])  # T(knots[d][2] - knots[d][1]), :72
    out, out_ptr, view = make_out(shape)
    n = (C.c_int64 * nd)(*shape)
    check(lib().cib200_smooth(nd, did, n, in_ptr, out_ptr, h, B, memkind, stream))
    del keep
    return view()
