"""ctypes front-end of the CPU ORACLE (oracle/cib_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
`--impl reference` legs of bench.py - never by the product package.

The Python part restates the reference's *host* logic that sits in front of the hot loops
(argument normalisation, mode resolution, knot extension); each helper cites the reference
file:line it follows (paths relative to /root/reference/src).  Parity status: unpinned against
real Julia output (Julia is not available); pinned against the shipped kernel tables, the
integer ghost matrices and the analytic properties of the reference's test-suite.

Array convention: `values[i1, i2, ...]` is indexed like the Julia array; memory order is free
(converted to column-major before the call).
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# generated reference constants (tools/gen_tables.py): one copy in the tree, shared with the library's build
_PKG = os.path.join(os.path.dirname(_HERE), "convolutioninterpolations.jl_b200")
_TABLES_INC = os.path.join(_PKG, "csrc", "cib200_tables.inc")
_POLYS_JSON = os.path.join(_PKG, "kernel_polys.json")
_LIB = None

KERNEL_IDS = {"b5": 0, "b7": 1, "b9": 2, "b11": 3, "b13": 4, "a3": 5, "a4": 6, "a5": 7, "a7": 8}
KERNEL_EQS = {"b5": 5, "b7": 6, "b9": 7, "b11": 8, "b13": 9, "a3": 2, "a4": 3, "a5": 3, "a7": 4}        # degree_to_equations.jl:13-25
MAX_SMOOTH_DERIVATIVE = {"b5": 3, "b7": 4, "b9": 5, "b11": 6, "b13": 6, "a3": 1, "a4": 1, "a5": 1, "a7": 1}  # constructor.jl:400-412
MAX_SHIPPED_DERIVATIVE = {"b5": 2, "b7": 3, "b9": 4, "b11": 5, "b13": 5, "a3": 0, "a4": 0, "a5": 0, "a7": 0}  # get_precomputed...jl:33-40
MIN_POLY_BC_POINTS = {"b5": 6, "b7": 8, "b9": 8, "b11": 8, "b13": 8, "a3": 4, "a4": 4, "a5": 4, "a7": 4}  # convolution_interpolation_function.jl:183-195
MAX_TABLE_DERIVATIVE = {"b5": 3, "b7": 4, "b9": 5, "b11": 6, "b13": 6, "a3": 1, "a4": 1, "a5": 1, "a7": 1}  # shipped K^(d)
SUBGRID_IDS = {"cubic": 0, "quintic": 1, "linear": 2}
BC_IDS = {"detect": 0, "poly": 1, "linear": 2, "quadratic": 3}
EXTRAP_IDS = {"throw": 0, "flat": 1, "line": 2}


class OracleError(RuntimeError):
    pass


def build(force=False):
    so = os.path.join(_HERE, "libcib_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cib_oracle.c", "cib_oracle.h")] + [_TABLES_INC]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "libcib_oracle.so"], check=True, capture_output=True)
    return so


class _Itp(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("cshape", C.c_int64 * 3), ("coefs", C.c_void_p),
                ("knots", C.c_void_p * 3), ("h", C.c_double * 3), ("kernel", C.c_int32 * 3),
                ("deriv", C.c_int32 * 3), ("extrap", C.c_int32), ("perdim", C.c_int32),
                ("subgrid", C.c_int32 * 3), ("tab", C.c_void_p * 3), ("n_pre", C.c_int32 * 3)]


class _ItpNd(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("cshape", C.c_int64 * 8), ("coefs", C.c_void_p),
                ("knots", C.c_void_p * 8), ("h", C.c_double * 8), ("kernel", C.c_int32 * 8),
                ("deriv", C.c_int32 * 8), ("extrap", C.c_int32), ("perdim", C.c_int32),
                ("tab", C.c_void_p * 8), ("n_pre", C.c_int32 * 8)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_table.restype = C.POINTER(C.c_double)
        L.orc_table.argtypes = [C.c_int, C.c_int]
        L.orc_itp_eval.argtypes = [C.POINTER(_Itp), C.c_int64, C.POINTER(C.c_void_p), C.c_void_p,
                                   C.POINTER(C.c_int64)]
        L.orc_itp_eval_nd.argtypes = [C.POINTER(_ItpNd), C.c_int64, C.POINTER(C.c_void_p), C.c_void_p,
                                      C.POINTER(C.c_int64)]
        L.orc_create_coefs.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_resample.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_void_p)] * 2 + [C.c_void_p] * 4
        L.orc_resample_sg.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_void_p)] * 2 + [C.c_void_p] * 5
        L.orc_resample_tables.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_void_p)] * 2 + [C.c_void_p] * 5 + \
            [C.POINTER(C.c_void_p), C.c_void_p]
        L.orc_smooth.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_smooth_halfwidth.argtypes = [C.c_double]
        L.orc_ghost_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_table_i1.restype = C.POINTER(C.c_double)
        L.orc_table_i1.argtypes = [C.c_int]
        L.orc_itg_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_itg_create_mixed.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int,
                                           C.POINTER(C.c_void_p)]
        L.orc_itg_eval.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_itg_destroy.argtypes = [C.c_void_p]
        L.orc_itg_right_tail_maxabs.restype = C.c_double
        L.orc_itg_right_tail_maxabs.argtypes = [C.c_void_p]
        L.orc_itg_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def set_threads(n):
    lib().orc_set_threads(int(n))


def table(kernel, d):
    """K^(d) table of `kernel` as a (101, 2*eqs) array (copy)."""
    p = lib().orc_table(KERNEL_IDS[kernel], int(d))
    if not p:
        raise OracleError(f"no table for {kernel} d={d}")
    e = KERNEL_EQS[kernel]
    return np.ctypeslib.as_array(p, shape=(101, 2 * e)).copy()


def table_i1(kernel):
    """antiderivative table K~ of `kernel` (KERNEL_PRE_<k>_i1) as a (101, 2*eqs) array (copy)."""
    p = lib().orc_table_i1(KERNEL_IDS[kernel])
    e = KERNEL_EQS[kernel]
    return np.ctypeslib.as_array(p, shape=(101, 2 * e)).copy()


LINEAR_ROWS = 10_000  # precompute_actual of a :linear axis: max(precompute, 10_000), constructor.jl:117-118
_GENERATED = {}


def generated_table(kernel, d, n_pre=LINEAR_ROWS):
    """precompute_kernel(F, degree, derivative, eqs, precompute, pre_range_exact), precompute_kernel_and_range.jl:150-157:
    kernel_pre[r, c] = F(K^(d)((r-1)//(n_pre-1) - eqs + c - 1)) in exact rational arithmetic.  Every entry is the ratio of
    two integers (sum of a_k/b_k * (m/Q)^k over the piece's coefficients, brought to the denominator lcm(b) * Q^n),
    rounded once by Python's correctly rounded int / int.  Returns an (n_pre, 2 eqs) C-ordered array (cached)."""
    key = (kernel, d, n_pre)
    if key in _GENERATED:
        return _GENERATED[key]
    import json
    from fractions import Fraction
    info = json.load(open(_POLYS_JSON))[kernel]
    eqs, Q = info["eqs"], n_pre - 1
    pieces = []
    for piece in info["pieces"]:
        c = [Fraction(x) for x in piece]
        dc = [c[k] * math.factorial(k) / math.factorial(k - d) for k in range(d, len(c))] or [Fraction(0)]
        lcm = 1
        for x in dc:
            lcm = lcm * x.denominator // math.gcd(lcm, x.denominator)
        pieces.append(([int(x * lcm) for x in dc], lcm))
    t = np.zeros((n_pre, 2 * eqs))
    for col in range(2 * eqs):
        for r in range(n_pre):
            m = r + (col - eqs) * Q  # s = m / Q
            M = abs(m)
            if M >= eqs * Q:
                continue
            A, lcm = pieces[M // Q]
            n = len(A) - 1
            num = sum(a * M ** k * Q ** (n - k) for k, a in enumerate(A))
            if d % 2 == 1 and m < 0:
                num = -num
            t[r, col] = num / (lcm * Q ** n)
    _GENERATED[key] = t
    return t


def ghost_matrix(which, kernel, rows):
    m = {"poly": MIN_POLY_BC_POINTS[kernel], "linear": 2, "quadratic": 3}[which]
    out = np.zeros((rows, m))
    lib().orc_ghost_matrix({"poly": 0, "linear": 1, "quadratic": 2}[which], KERNEL_IDS[kernel], rows,
                           out.ctypes.data)
    return out


# ---------------------------------------------------------------------------------------------
# argument normalisation: convolution_interpolation_function.jl:88-107
# ---------------------------------------------------------------------------------------------
def _ntuple(v, n, scalar_types):
    if isinstance(v, scalar_types):
        return (v,) * n
    v = tuple(v)
    if len(v) == 1:
        return v * n
    if len(v) != n:
        raise OracleError(f"Invalid specification: {v}")
    return v


def _norm_bc(bc, n):
    if isinstance(bc, str):
        return tuple((bc, bc) for _ in range(n))
    bc = tuple(bc)
    if len(bc) == 2 and all(isinstance(b, str) for b in bc):
        return tuple(tuple(bc) for _ in range(n))
    if len(bc) == 1:
        return tuple(tuple(bc[0]) for _ in range(n))
    if len(bc) != n:
        raise OracleError(f"Invalid bc specification: {bc}")
    return tuple(tuple(b) for b in bc)


def _knots_tuple(knots, ndim):
    if ndim == 1 and not isinstance(knots, (tuple, list)):
        knots = (knots,)
    if isinstance(knots, np.ndarray) and knots.ndim == 1 and ndim == 1:
        knots = (knots,)
    return tuple(np.ascontiguousarray(k, dtype=np.float64) for k in knots)


def build_subgrids(kernels, derivs, subgrids):
    """fast_convolution_interpolation_constructor.jl:456-494 (no a0/a1, no integral dims)."""
    n = len(kernels)
    all_cubic = len(set(subgrids)) == 1 and subgrids[0] == "cubic"
    out = []
    for k, dv, sg in zip(kernels, derivs, subgrids):
        if dv == 0:
            out.append("linear" if (n == 3 and not all_cubic) or n >= 4 else sg)
        else:
            avail = MAX_SMOOTH_DERIVATIVE[k] - dv
            out.append("linear" if avail <= 0 or (n == 3 and not all_cubic) or n >= 4 else sg)
    return tuple(out)


def effective_subgrids(subgrids, ndim):
    """What the evaluators actually do with a resolved subgrid tuple: in 3-D anything but all-cubic is evaluated with
    linear blending on every axis (eager_3d.jl:139-176; perdim_3d.jl:161 "fallback to all linear")."""
    if ndim == 3 and tuple(subgrids) != ("cubic",) * 3:
        return ("linear",) * 3
    return tuple(subgrids)


def check_subgrids(kernels, derivs, subgrids, ndim, perdim, what):
    """Which resolved (kernel, derivative, subgrid) combinations the shipped 101-row tables can serve (precompute = 101:
    the generated tables of the :linear path are then the shipped K^(d), both are exact rationals rounded once):
    :linear needs K^(d); :cubic also K^(d+1); :quintic also K^(d+2) and exists in 1-D / 2-D only."""
    for d in range(ndim):
        k, dv, sg = kernels[d], derivs[d], subgrids[d]
        if dv < 0 or sg not in ("cubic", "quintic", "linear"):
            raise OracleError(f"{what}: mode outside the oracle's scope: subgrid={subgrids} deriv={derivs}")
        need = dv + {"linear": 0, "cubic": 1, "quintic": 2}[sg]
        if need > MAX_TABLE_DERIVATIVE[k]:
            raise OracleError(f"{what}: K^({need}) of :{k} is not shipped ({sg} subgrid, derivative {dv})")
    if what == "interpolation":
        if "quintic" in subgrids and (ndim > 2 or (ndim == 2 and not perdim and subgrids != ("quintic", "quintic"))):
            raise OracleError(f"{what}: subgrid {subgrids} has no evaluator on this path")
        if ndim == 2 and not perdim and len(set(subgrids)) != 1:
            raise OracleError(f"{what}: subgrid {subgrids} has no evaluator on the single-kernel 2-D path")


def extend_vector(x, n_extra):
    """convolution/extend_vector.jl:10-18"""
    step_start = x[1] - x[0]
    step_end = x[-1] - x[-2]
    start = np.array([x[0] - i * step_start for i in range(n_extra, 0, -1)])
    end = np.array([x[-1] + i * step_end for i in range(1, n_extra + 1)])
    return np.concatenate([start, x, end])


def create_coefs(values, h, kernels, bcs):
    """create_convolutional_coefs.jl:141-160 -> padded array (Julia index order)."""
    v = np.asfortranarray(values, dtype=np.float64)
    nd = v.ndim
    n = (C.c_int64 * nd)(*v.shape)
    hh = (C.c_double * nd)(*h)
    kid = (C.c_int32 * nd)(*[KERNEL_IDS[k] for k in kernels])
    bl = (C.c_int32 * nd)(*[BC_IDS[b[0]] for b in bcs])
    br = (C.c_int32 * nd)(*[BC_IDS[b[1]] for b in bcs])
    cshape = tuple(s + 2 * (KERNEL_EQS[k] - 1) for s, k in zip(v.shape, kernels))
    c = np.zeros(cshape, order="F")
    rc = lib().orc_create_coefs(nd, n, v.ctypes.data, hh, kid, bl, br, c.ctypes.data)
    if rc:
        raise OracleError(f"orc_create_coefs rc={rc}")
    return c


class Interpolant:
    """convolution_interpolation(knots, values; kernel, extrap, bc, derivative) - fast eager path,
    cubic subgrid (convolution_interpolation_function.jl:79-170,
    fast_convolution_interpolation_constructor.jl:99-174)."""

    def __init__(self, knots, values, kernel="b5", extrap="throw", bc="detect", derivative=0,
                 subgrid="cubic", linear_rows=LINEAR_ROWS):
        values = np.asarray(values, dtype=np.float64)
        nd = values.ndim
        if nd < 1 or nd > 8:
            raise OracleError("oracle covers 1-8 D")
        self.ndim = nd
        self.knots_in = _knots_tuple(knots, nd)
        self.kernels = _ntuple(kernel, nd, str)
        self.derivs = _ntuple(derivative, nd, int)
        sg_req = _ntuple(subgrid, nd, str)
        bcs = _norm_bc(bc, nd)
        # :poly -> :linear downgrade, convolution_interpolation_function.jl:135-143
        fixed = []
        for d in range(nd):
            b = bcs[d]
            if "poly" in b and values.shape[d] < MIN_POLY_BC_POINTS[self.kernels[d]]:
                if b == ("poly", "poly"):
                    b = ("linear", "linear")
                elif b[0] == "poly":
                    b = ("linear", b[1])
                else:
                    b = (b[0], "linear")
            fixed.append(b)
        self.bc = tuple(fixed)
        self.subgrid = build_subgrids(self.kernels, self.derivs, sg_req)
        self.perdim = not (len(set(self.kernels)) == 1 and len(set(self.derivs)) == 1)
        self.subgrid = effective_subgrids(self.subgrid, nd)
        check_subgrids(self.kernels, self.derivs, self.subgrid, nd, self.perdim, "interpolation")
        self.eqs = tuple(KERNEL_EQS[k] for k in self.kernels)
        self.h = tuple(float(k[1] - k[0]) for k in self.knots_in)  # constructor.jl:120
        self.coefs = create_coefs(values, self.h, self.kernels, self.bc)
        self.knots = tuple(extend_vector(k, e - 1) for k, e in zip(self.knots_in, self.eqs))
        self.extrap = extrap
        # _build_kernel_sym, constructor.jl:378-398
        self.perdim = not (len(set(self.kernels)) == 1 and len(set(self.derivs)) == 1)
        it = _Itp() if nd <= 3 else _ItpNd()  # N >= 4: HigherDimension evaluators (eager_nd.jl / perdim_nd.jl)
        it.ndim = nd
        for d in range(nd):
            it.cshape[d] = self.coefs.shape[d]
            it.knots[d] = self.knots[d].ctypes.data
            it.h[d] = self.h[d]
            it.kernel[d] = KERNEL_IDS[self.kernels[d]]
            it.deriv[d] = self.derivs[d]
            if nd <= 3:
                it.subgrid[d] = SUBGRID_IDS[self.subgrid[d]]
        # :linear axes read the generated >= 10 000-row table (constructor.jl:117-118); linear_rows=101 keeps the shipped
        # one (what the reference would compute from a 101-row table; only used to cross-check the two)
        self._tabs = []
        for d in range(nd):
            if self.subgrid[d] == "linear" and linear_rows != 101:
                t = generated_table(self.kernels[d], self.derivs[d], linear_rows)
                self._tabs.append(t)
                it.tab[d] = t.ctypes.data
                it.n_pre[d] = linear_rows
        it.coefs = self.coefs.ctypes.data
        it.extrap = EXTRAP_IDS[extrap]
        it.perdim = int(self.perdim)
        self._it = it

    def eval_points(self, *xs):
        """itp.(xs...) : element-wise over equally long coordinate vectors."""
        xs = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), np.broadcast(*xs).shape)).ravel()
              for x in xs]
        assert len(xs) == self.ndim
        nq = xs[0].size
        out = np.empty(nq)
        ptrs = (C.c_void_p * self.ndim)(*[x.ctypes.data for x in xs])
        bad = C.c_int64(-1)
        fn = lib().orc_itp_eval if self.ndim <= 3 else lib().orc_itp_eval_nd
        rc = fn(C.byref(self._it), nq, ptrs, out.ctypes.data, C.byref(bad))
        if rc == 1:
            raise OracleError(f"query {bad.value} outside the domain with extrap=throw")
        if rc == 2:
            raise OracleError(f"query {bad.value}: InexactError (NaN/Inf coordinate)")
        if rc:
            raise OracleError(f"orc_itp_eval rc={rc}")
        return out

    eval = eval_points

    def __call__(self, *xs):
        if all(np.ndim(x) == 0 for x in xs):
            return float(self.eval_points(*[np.array([x]) for x in xs])[0])
        # extrapolation_operator.jl:66-99: Cartesian product of coordinate vectors
        grids = np.meshgrid(*[np.asarray(x, dtype=np.float64) for x in xs], indexing="ij")
        return self.eval_points(*[g.ravel() for g in grids]).reshape(grids[0].shape)


def build_subgrids_integral(kernels, subgrids, ndim):
    """_build_subgrids for integral directions, fast_convolution_interpolation_constructor.jl:464-472"""
    out = []
    for k, sg in zip(kernels, subgrids):
        out.append("linear" if ndim >= 4 else ("cubic" if k in ("a3", "a4", "a5", "a7") else sg))
    return tuple(out)


class IntegralInterpolant:
    """convolution_interpolation(knots, values; kernel, derivative=-1, ...) with EVERY axis an integral direction, 1-D to
    3-D, fast path: FastConvolutionInterpolation with FastIntegralOrder (fast_convolution_interpolation_constructor.jl:
    99-174, tails :176-364; evaluators convolution_fast_integration_{1d,2d,3d}.jl).  Result anchored to zero at the first
    interior knot of each axis (`anchor`, constructor.jl:139)."""

    def __init__(self, knots, values, kernel="b5", extrap="throw", bc="detect", subgrid="cubic", derivative=-1,
                 linear_rows=LINEAR_ROWS):
        values = np.asarray(values, dtype=np.float64)
        nd = values.ndim
        if nd < 1 or nd > 3:
            raise OracleError("antiderivative oracle covers 1-3 D")
        self.ndim = nd
        self.knots_in = _knots_tuple(knots, nd)
        self.kernels = _ntuple(kernel, nd, str)
        self.derivs = _ntuple(derivative, nd, int)  # -1: integral axis; >= 0: :linear axis of a mixed interpolant
        bcs = _norm_bc(bc, nd)
        fixed = []
        for d in range(nd):  # :poly -> :linear downgrade, convolution_interpolation_function.jl:135-143
            b = bcs[d]
            if "poly" in b and values.shape[d] < MIN_POLY_BC_POINTS[self.kernels[d]]:
                b = ("linear", "linear") if b == ("poly", "poly") else (("linear", b[1]) if b[0] == "poly" else (b[0], "linear"))
            fixed.append(b)
        self.bc = tuple(fixed)
        sg_int = build_subgrids_integral(self.kernels, _ntuple(subgrid, nd, str), nd)
        # non-integral axes of a mixed interpolant resolve to :linear (has_integral_dim, constructor.jl:474-491)
        self.subgrid = tuple(sg_int[d] if self.derivs[d] == -1 else "linear" for d in range(nd))
        for k, sg, dv in zip(self.kernels, self.subgrid, self.derivs):
            if dv == -1 and (sg == "linear" or (sg == "quintic" and MAX_TABLE_DERIVATIVE[k] < 1)):
                raise OracleError(f"antiderivative: subgrid {self.subgrid} needs generated tables")
            if dv < -1 or dv > MAX_TABLE_DERIVATIVE[k]:
                raise OracleError(f"K^({dv}) of :{k} does not exist")
        self.eqs = tuple(KERNEL_EQS[k] for k in self.kernels)
        self.h = tuple(float(k[1] - k[0]) for k in self.knots_in)
        self.coefs = create_coefs(values, self.h, self.kernels, self.bc)
        self.knots = tuple(extend_vector(k, e - 1) for k, e in zip(self.knots_in, self.eqs))
        # anchor = knots_new[d][eqs[d]] on the integral axes, zero(T) elsewhere (constructor.jl:139)
        self.anchor = tuple(float(k[e - 1]) if dv == -1 else 0.0 for k, e, dv in zip(self.knots, self.eqs, self.derivs))
        self.extrap = extrap
        cs = (C.c_int64 * nd)(*self.coefs.shape)
        kp = (C.c_void_p * nd)(*[k.ctypes.data for k in self.knots])
        hh = (C.c_double * nd)(*self.h)
        kid = (C.c_int32 * nd)(*[KERNEL_IDS[k] for k in self.kernels])
        sg = (C.c_int32 * nd)(*[SUBGRID_IDS[x] for x in self.subgrid])
        dvs = (C.c_int32 * nd)(*self.derivs)
        self._tabs = [generated_table(self.kernels[d], self.derivs[d], linear_rows)
                      if (self.derivs[d] >= 0 and linear_rows != 101) else None for d in range(nd)]
        tp = (C.c_void_p * nd)(*[None if t is None else t.ctypes.data for t in self._tabs])
        npre = (C.c_int32 * nd)(*[0 if t is None else t.shape[0] for t in self._tabs])
        h = C.c_void_p()
        rc = lib().orc_itg_create_mixed(nd, cs, self.coefs.ctypes.data, kp, hh, kid, sg, dvs, tp, npre,
                                        EXTRAP_IDS[extrap], C.byref(h))
        if rc:
            raise OracleError(f"orc_itg_create rc={rc}")
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_itg_destroy(self._h)
            self._h = None

    def right_tail_maxabs(self):
        return float(lib().orc_itg_right_tail_maxabs(self._h))

    def left_values(self, d):
        out = np.empty(self.coefs.shape[d])
        lib().orc_itg_get(self._h, 0, d, out.ctypes.data)
        return out

    def tail(self, code):
        out = np.empty(self.coefs.shape, order="F")
        if lib().orc_itg_get(self._h, 1, code, out.ctypes.data):
            raise OracleError(f"no tail {code}")
        return out

    def eval_points(self, *xs):
        xs = [np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), np.broadcast(*xs).shape)).ravel()
              for x in xs]
        assert len(xs) == self.ndim
        nq = xs[0].size
        out = np.empty(nq)
        ptrs = (C.c_void_p * self.ndim)(*[x.ctypes.data for x in xs])
        bad = C.c_int64(-1)
        rc = lib().orc_itg_eval(self._h, nq, ptrs, out.ctypes.data, C.byref(bad))
        if rc == 1:
            raise OracleError(f"query {bad.value} outside the domain with extrap=throw")
        if rc == 2:
            raise OracleError(f"query {bad.value}: InexactError (NaN/Inf coordinate)")
        if rc:
            raise OracleError(f"orc_itg_eval rc={rc}")
        return out

    eval = eval_points

    def __call__(self, *xs):
        if all(np.ndim(x) == 0 for x in xs):
            return float(self.eval_points(*[np.array([x]) for x in xs])[0])
        grids = np.meshgrid(*[np.asarray(x, dtype=np.float64) for x in xs], indexing="ij")
        return self.eval_points(*[g.ravel() for g in grids]).reshape(grids[0].shape)


def convolution_interpolation(knots, values, **kw):
    dv = kw.get("derivative", 0)
    nd = np.ndim(values)
    dvs = (dv,) * nd if isinstance(dv, int) else tuple(dv) * (nd if len(tuple(dv)) == 1 else 1)
    if any(d == -1 for d in dvs):
        # pure (FastIntegralOrder) or mixed (FastMixedIntegralOrder: the other axes on :linear, generated tables)
        return IntegralInterpolant(knots, values, **kw)
    return Interpolant(knots, values, **kw)


_UNSET = object()


def convolution_resample(knots_in, knots_out, values, kernel=_UNSET, derivative=_UNSET, bc=_UNSET,
                         subgrid=_UNSET, linear_rows=LINEAR_ROWS):
    """convolution_resample.jl:65-172 (kw-less wrappers: no _build_subgrids) / :174-238 (kw method) + :240-386."""
    kwless = all(a is _UNSET for a in (kernel, derivative, bc, subgrid))
    kernel = "b13" if kernel is _UNSET else kernel
    derivative = 0 if derivative is _UNSET else derivative
    bc = "detect" if bc is _UNSET else bc
    subgrid = "cubic" if subgrid is _UNSET else subgrid
    v = np.asfortranarray(values, dtype=np.float64)
    nd = v.ndim
    kin = _knots_tuple(knots_in, nd)
    kout = _knots_tuple(knots_out, nd)
    kernels = _ntuple(kernel, nd, str)
    derivs = _ntuple(derivative, nd, int)
    bcs = _norm_bc(bc, nd)
    sgs = _ntuple(subgrid, nd, str) if kwless else build_subgrids(kernels, derivs, _ntuple(subgrid, nd, str))
    for d in range(nd):
        if derivs[d] < 0:
            raise OracleError("convolution_resample does not support antiderivatives")
    check_subgrids(kernels, derivs, sgs, nd, True, "resample")
    n_in = (C.c_int64 * nd)(*v.shape)
    n_out = (C.c_int64 * nd)(*[len(k) for k in kout])
    out = np.zeros(tuple(len(k) for k in kout), order="F")
    kinp = (C.c_void_p * nd)(*[k.ctypes.data for k in kin])
    koutp = (C.c_void_p * nd)(*[k.ctypes.data for k in kout])
    kid = (C.c_int32 * nd)(*[KERNEL_IDS[k] for k in kernels])
    dv = (C.c_int32 * nd)(*derivs)
    bl = (C.c_int32 * nd)(*[BC_IDS[b[0]] for b in bcs])
    br = (C.c_int32 * nd)(*[BC_IDS[b[1]] for b in bcs])
    sg = (C.c_int32 * nd)(*[SUBGRID_IDS[x] for x in sgs])
    # kw method: precompute_actual = max(precompute, 10_000) for :linear axes (:228-229); the kw-less wrappers pass 101
    tabs = [generated_table(kernels[d], derivs[d], linear_rows) if (sgs[d] == "linear" and not kwless and linear_rows != 101)
            else None for d in range(nd)]
    tp = (C.c_void_p * nd)(*[None if t is None else t.ctypes.data for t in tabs])
    npre = (C.c_int32 * nd)(*[0 if t is None else t.shape[0] for t in tabs])
    rc = lib().orc_resample_tables(nd, n_in, n_out, v.ctypes.data, out.ctypes.data, kinp, koutp, kid, dv, bl, br, sg,
                                   tp, npre)
    if rc:
        raise OracleError(f"orc_resample rc={rc}")
    return out


def smooth_halfwidth(B):
    return lib().orc_smooth_halfwidth(float(B))


def convolution_smooth(knots, values, B):
    """convolution_gaussian_smoothing.jl:45-131"""
    v = np.asfortranarray(values, dtype=np.float64)
    nd = v.ndim
    ks = _knots_tuple(knots, nd)
    E = smooth_halfwidth(B)
    for d in range(nd):
        if E >= len(ks[d]):
            raise OracleError(f"Gaussian stencil half-width ({E} points) exceeds grid size (n={len(ks[d])}) "
                              f"in dimension {d + 1}.")
    n = (C.c_int64 * nd)(*v.shape)
    h = (C.c_double * nd)(*[float(k[1] - k[0]) for k in ks])
    out = np.zeros(v.shape, order="F")
    rc = lib().orc_smooth(nd, n, v.ctypes.data, out.ctypes.data, h, float(B))
    if rc:
        raise OracleError(f"orc_smooth rc={rc}")
    return out


def theta(B, terms):
    q = math.exp(-B)
    return 1.0 + sum(2 * q ** (n * n) for n in range(1, terms + 1))


# ---------------------------------------------------------------------------------------------
# SECOND oracle: the reference's DIRECT evaluators.  They share nothing with the fast path but
# the ghost-padded coefficient grid: no tables, no Hermite blending - every stencil weight is the kernel polynomial
# itself, evaluated at the query by Horner's rule.  The reference asserts fast ~ direct to 1e-6
# (test/test_constructors.jl:4-49); the table + cubic-Hermite scheme is good to ~1e-9, which is what pins the fast
# restatement (and the CUDA path) against an independent statement of the same interpolant.
# ---------------------------------------------------------------------------------------------
_POLY_CACHE = {}


def kernel_poly_coefs(kernel, d):
    """Float64 coefficient vectors of K^(d) per piece, as the reference stores them: the exact rational
    k (k-1) ... (k-d+1) c_k (convolution_kernel_b5.jl:9-60: `b5_coefs_d1 = [0//1*c0, 1//1*c1, ...]`), converted once by
    T(coef[i]) inside horner (horner.jl:44-52).  The d leading (zero) coefficients are kept: Horner skips them."""
    key = (kernel, d)
    if key not in _POLY_CACHE:
        import json
        from fractions import Fraction
        info = json.load(open(_POLYS_JSON))[kernel]
        out = []
        for piece in info["pieces"]:
            c = [Fraction(x) for x in piece]
            dc = [c[k] * (math.factorial(k) // math.factorial(k - d)) if k >= d else Fraction(0) for k in range(len(c))]
            out.append(np.array([float(x) for x in dc]))
        _POLY_CACHE[key] = (info["eqs"], out)
    return _POLY_CACHE[key]


def kernel_value(kernel, d, s):
    """(::ConvolutionKernel{kernel, DO})(s), e.g. convolution_kernel_b5.jl:62-92: piece by floor(|s|), Horner from the
    highest coefficient down to index 1 + DO (horner.jl:48-50), times sign(s) for odd DO, zero for |s| >= eqs."""
    eqs, pieces = kernel_poly_coefs(kernel, d)
    s = np.asarray(s, dtype=np.float64)
    sa = np.abs(s)
    out = np.zeros_like(sa)
    piece = np.floor(sa).astype(np.int64)
    for p, coef in enumerate(pieces):
        m = piece == p
        if not m.any():
            continue
        x = sa[m]
        r = np.full_like(x, coef[-1])
        for i in range(len(coef) - 2, d - 1, -1):
            r = r * x + coef[i]
        out[m] = r
    if d % 2 == 1:
        out = out * np.sign(s)
    return out


def direct_eval(itp, *xs):
    """The direct evaluators convolution_kernel_interpolation_eager_{1d,2d,3d}.jl (3-D: :1-29) on the ghost-padded grid
    of `itp` (an Interpolant): result = sum_{l,m,n} coefs[i+l, j+m, k+n] K1(s1) K2(s2) K3(s3), s_d = (x_d - knots_d[.]) / h_d,
    times prod_d 1 / h_d^DO_d.  In-domain queries only (the extrapolation wrapper is the fast path's, not restated here)."""
    nd = itp.ndim
    assert 1 <= nd <= 3 and len(xs) == nd
    xs = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in xs]
    nq = xs[0].size
    idx, wts = [], []
    for d in range(nd):
        k, h, e = itp.knots[d], itp.h[d], itp.eqs[d]
        i_float = (xs[d] - k[0]) / h + 1.0
        i = np.clip(np.floor(i_float).astype(np.int64), e, len(k) - e)  # 1-based
        taps = np.arange(-(e - 1), e + 1)
        pos = (i[:, None] - 1) + taps[None, :]  # 0-based indices of the 2 eqs taps
        s = (xs[d][:, None] - k[pos]) / h
        idx.append(pos)
        wts.append(kernel_value(itp.kernels[d], itp.derivs[d], s))
    c = itp.coefs
    res = np.zeros(nq)
    if nd == 1:
        for l in range(wts[0].shape[1]):
            res += c[idx[0][:, l]] * wts[0][:, l]
    elif nd == 2:
        for l in range(wts[0].shape[1]):
            for m in range(wts[1].shape[1]):
                res += c[idx[0][:, l], idx[1][:, m]] * wts[0][:, l] * wts[1][:, m]
    else:
        for l in range(wts[0].shape[1]):
            for m in range(wts[1].shape[1]):
                wlm = wts[0][:, l] * wts[1][:, m]
                for n in range(wts[2].shape[1]):
                    res += c[idx[0][:, l], idx[1][:, m], idx[2][:, n]] * wlm * wts[2][:, n]
    scale = 1.0
    for d in range(nd):
        scale *= 1.0 / itp.h[d] ** itp.derivs[d]
    return res * scale


# ---------------------------------------------------------------------------------------------
# Antiderivative / mixed interpolants in N >= 4 dimensions (the remainder):
#   convolution_fast_integration_nd.jl:1-28          every axis integral (HigherDimension, no tails)
#   convolution_fast_mixed_3_integral_dims.jl:1-232  exactly three integral axes (27-region tail decomposition)
#   convolution_fast_mixed_N_integral_dims.jl:15-82  any other count of integral axes >= 4 (brute force over the grid)
# restated in numpy / plain Python loops (small cases only; test infrastructure).  For N >= 4 every axis is on the
# :linear subgrid (_build_subgrids, constructor.jl:456-494) and reads generated tables of >= 10 000 rows.
# ---------------------------------------------------------------------------------------------
_GENERATED_I1 = {}


def generated_table_i1(kernel, n_pre=LINEAR_ROWS):
    """kernel_pre of an integral axis on the :linear subgrid: K~((r-1)//(n_pre-1) - eqs + c - 1) = int_0^s K(u) du in exact
    rational arithmetic, rounded once (precompute_kernel_and_range.jl:150-157 with the _i1 polynomials,
    convolution_kernel_b5.jl:52-58); odd in s, +-1/2 outside the support."""
    key = (kernel, n_pre)
    if key in _GENERATED_I1:
        return _GENERATED_I1[key]
    import json
    from fractions import Fraction
    info = json.load(open(_POLYS_JSON))[kernel]
    eqs, Q = info["eqs"], n_pre - 1
    pieces, below = [], Fraction(0)
    for p, piece in enumerate(info["pieces"]):
        c = [Fraction(x) for x in piece]
        a = [Fraction(0)] + [c[k] / (k + 1) for k in range(len(c))]
        val = lambda s, a=a: sum(ak * s ** k for k, ak in enumerate(a))
        a[0] = below - val(Fraction(p))
        below += val(Fraction(p + 1)) - val(Fraction(p))
        lcm = 1
        for x in a:
            lcm = lcm * x.denominator // math.gcd(lcm, x.denominator)
        pieces.append(([int(x * lcm) for x in a], lcm))
    t = np.zeros((n_pre, 2 * eqs))
    for col in range(2 * eqs):
        for r in range(n_pre):
            m = r + (col - eqs) * Q
            M = abs(m)
            if M >= eqs * Q:
                t[r, col] = 0.5 if m > 0 else -0.5
                continue
            A, lcm = pieces[M // Q]
            n = len(A) - 1
            num = sum(a * M ** k * Q ** (n - k) for k, a in enumerate(A))
            t[r, col] = (num if m >= 0 else -num) / (lcm * Q ** n)
    _GENERATED_I1[key] = t
    return t


def _suffix_sum(a, axis):
    """_suffix_sum, constructor.jl:450-454"""
    return np.flip(np.cumsum(np.flip(a, axis=axis), axis=axis), axis=axis)


class IntegralInterpolantND:
    """convolution_interpolation(knots, values; kernel, derivative) with derivative[d] = -1 on some or all axes, N >= 4
    (also constructible for N < 4 to cross-check the functors against each other).  Every axis on the :linear subgrid."""

    def __init__(self, knots, values, kernel="b5", derivative=-1, bc="detect", extrap="throw", linear_rows=LINEAR_ROWS,
                 functor=None):
        values = np.asarray(values, dtype=np.float64)
        nd = values.ndim
        self.ndim = nd
        self.knots_in = _knots_tuple(knots, nd)
        self.kernels = _ntuple(kernel, nd, str)
        self.derivs = _ntuple(derivative, nd, int)
        bcs = _norm_bc(bc, nd)
        fixed = []
        for d in range(nd):
            b = bcs[d]
            if "poly" in b and values.shape[d] < MIN_POLY_BC_POINTS[self.kernels[d]]:
                b = ("linear", "linear") if b == ("poly", "poly") else (("linear", b[1]) if b[0] == "poly" else (b[0], "linear"))
            fixed.append(b)
        self.eqs = tuple(KERNEL_EQS[k] for k in self.kernels)
        self.h = tuple(float(k[1] - k[0]) for k in self.knots_in)
        self.coefs = create_coefs(values, self.h, self.kernels, tuple(fixed))
        self.knots = tuple(extend_vector(k, e - 1) for k, e in zip(self.knots_in, self.eqs))
        self.domain_size = values.shape
        self.extrap = extrap
        self.tables = tuple(generated_table_i1(k, linear_rows) if dv == -1 else generated_table(k, dv, linear_rows)
                            for k, dv in zip(self.kernels, self.derivs))
        self.n_pre = linear_rows
        self.int_dims = [d for d in range(nd) if self.derivs[d] == -1]
        ni = len(self.int_dims)
        # functor selection: integral_dimension = Val(n_integral) for <= 3, HigherDimension otherwise (constructor.jl:150)
        self.functor = functor or ("nd" if ni == nd else ("mixed3" if ni == 3 else "mixedN"))
        self.left_values = tuple(self._left_values(d) if self.derivs[d] == -1 else np.zeros(1) for d in range(nd))
        if self.functor == "mixed3":
            self._build_tails3()

    # _compute_left_values, constructor.jl:414-448 (linear subgrid branch)
    def _left_values(self, d):
        eqs, n = self.eqs[d], self.coefs.shape[d]
        tab, npre = self.tables[d], self.n_pre
        lv = np.zeros(n)
        for j in range(1, n + 1):
            sj = float(eqs - j)
            if abs(sj) >= eqs:
                lv[j - 1] = 0.5 * np.sign(sj)
                continue
            col_float = eqs + sj
            col = min(max(int(math.floor(col_float)) + 1, 1), 2 * eqs)
            cidx = (col_float - (col - 1)) * (npre - 1) + 1.0
            idx = min(max(int(math.floor(cidx)), 1), npre - 1)
            t = cidx - idx
            lv[j - 1] = (1.0 - t) * tab[idx - 1, col - 1] + t * tab[idx, col - 1]
        return lv

    # K~ lookup of an integral axis: eval_sg_kt (eval_subgrid_helpers.jl:22-32), linear subgrid
    def _kt(self, d, j, x):
        eqs, h = self.eqs[d], self.h[d]
        xjd = self.knots[d][eqs - 1] + (j - eqs) * h
        sj = (x - xjd) / h
        if abs(sj) >= eqs:
            return 0.5 * np.sign(sj)
        col_float = eqs + sj
        col = min(max(int(math.floor(col_float)) + 1, 1), 2 * eqs)
        t = (col_float - (col - 1)) * (self.n_pre - 1) + 1.0
        idx = min(max(int(math.floor(t)), 1), self.n_pre - 1)
        t -= idx
        tab = self.tables[d]
        return (1.0 - t) * tab[idx - 1, col - 1] + t * tab[idx, col - 1]

    # K^(DO) lookup of a derivative axis at padded index j (mixed_N :52-68); None outside the stencil
    def _kd(self, d, j, x):
        eqs, h, k = self.eqs[d], self.h[d], self.knots[d]
        pos = min(max(int(math.floor((x - k[0]) / h + 1.0)), eqs), len(k) - eqs)
        col = j - pos + eqs
        if col < 1 or col > 2 * eqs:
            return None
        diff_right = 1.0 - (x - k[pos - 1]) / h
        cidx = diff_right * (self.n_pre - 1) + 1.0
        i = min(max(int(math.floor(cidx)), 1), self.n_pre - 1)
        t = cidx - i
        tab = self.tables[d]
        return (1.0 - t) * tab[i - 1, col - 1] + t * tab[i, col - 1]

    def _scale(self):
        s = 1.0
        for d in range(self.ndim):
            s *= self.h[d] if self.derivs[d] == -1 else (-1.0 / self.h[d]) ** self.derivs[d]
        return s

    def _eval_nd(self, x):
        """convolution_fast_integration_nd.jl:16-28: the loops run over 1:domain_size[d] - the INTERIOR extents - while
        coefs and left_values are indexed in padded numbering: the last 2 (eqs - 1) padded entries of every axis are
        never visited (restated literally)."""
        nd = self.ndim
        w = [np.array([self._kt(d, j, x[d]) - self.left_values[d][j - 1] for j in range(1, self.domain_size[d] + 1)])
             for d in range(nd)]
        sub = self.coefs[tuple(slice(0, n) for n in self.domain_size)]
        acc = sub
        for d in range(nd - 1, -1, -1):
            acc = np.tensordot(acc, w[d], axes=([d], [0]))
        return float(acc) * float(np.prod(self.h))

    def _eval_mixed_n(self, x):
        """convolution_fast_mixed_N_integral_dims.jl:24-82: brute force over the whole padded grid"""
        nd = self.ndim
        w = []
        for d in range(nd):
            n = self.coefs.shape[d]
            if self.derivs[d] == -1:
                w.append(np.array([self._kt(d, j, x[d]) - self.left_values[d][j - 1] for j in range(1, n + 1)]))
            else:
                w.append(np.array([self._kd(d, j, x[d]) or 0.0 for j in range(1, n + 1)]))
        acc = self.coefs
        for d in range(nd - 1, -1, -1):
            acc = np.tensordot(acc, w[d], axes=([d], [0]))
        return float(acc) * self._scale()

    def _build_tails3(self):
        """_build_tails_dispatch(..., ::Val{3}), constructor.jl:277-334: the arrays the functor reads"""
        c = self.coefs
        nd = self.ndim
        d1, d2, d3 = self.int_dims
        shp = lambda d: tuple(c.shape[d] if i == d else 1 for i in range(nd))
        lv = {d: self.left_values[d].reshape(shp(d)) for d in (d1, d2, d3)}
        wl = {d: 0.5 - lv[d] for d in lv}
        wr = {d: -0.5 - lv[d] for d in lv}
        cs, ss = (lambda a, d: np.cumsum(a, axis=d)), _suffix_sum
        self.tail1_left = {d: cs(c * wl[d], d) for d in lv}
        self.tail1_right = {d: ss(c * wr[d], d) for d in lv}
        # tail3_edge_xx[k]: free in int_dim k, saturated in the other two (1-based k = 1, 2, 3 as in the reference)
        self.edge_ll = {1: cs(cs(c * wl[d2] * wl[d3], d2), d3), 2: cs(cs(c * wl[d1] * wl[d3], d1), d3), 3: cs(cs(c * wl[d1] * wl[d2], d1), d2)}
        self.edge_rl = {1: ss(cs(c * wr[d2] * wl[d3], d3), d2), 2: ss(cs(c * wr[d1] * wl[d3], d3), d1), 3: ss(cs(c * wr[d1] * wl[d2], d2), d1)}
        self.edge_lr = {1: cs(ss(c * wl[d2] * wr[d3], d3), d2), 2: cs(ss(c * wl[d1] * wr[d3], d3), d1), 3: cs(ss(c * wl[d1] * wr[d2], d2), d1)}
        self.edge_rr = {1: ss(ss(c * wr[d2] * wr[d3], d3), d2), 2: ss(ss(c * wr[d1] * wr[d3], d3), d1), 3: ss(ss(c * wr[d1] * wr[d2], d2), d1)}
        self.corner = {
            "lll": cs(cs(cs(c * wl[d1] * wl[d2] * wl[d3], d1), d2), d3),
            "rll": ss(cs(cs(c * wr[d1] * wl[d2] * wl[d3], d2), d3), d1),
            "lrl": cs(ss(cs(c * wl[d1] * wr[d2] * wl[d3], d3), d2), d1),
            "llr": cs(cs(ss(c * wl[d1] * wl[d2] * wr[d3], d3), d1), d2),
            "rrl": ss(ss(cs(c * wr[d1] * wr[d2] * wl[d3], d3), d2), d1),
            "rlr": ss(cs(ss(c * wr[d1] * wl[d2] * wr[d3], d2), d1), d3),
            "lrr": cs(ss(ss(c * wl[d1] * wr[d2] * wr[d3], d3), d2), d1),
            "rrr": ss(ss(ss(c * wr[d1] * wr[d2] * wr[d3], d3), d2), d1),
        }

    def _eval_mixed3(self, x):
        """convolution_fast_mixed_3_integral_dims.jl:8-225 (1-based indices kept; arrays indexed with -1)"""
        import itertools
        nd, c = self.ndim, self.coefs
        d1, d2, d3 = self.int_dims
        isint = [d in (d1, d2, d3) for d in range(nd)]
        i = [min(max(int(math.floor((x[d] - self.knots[d][0]) / self.h[d] + 1.0)), self.eqs[d]), c.shape[d] - self.eqs[d])
             for d in range(nd)]
        l = [i[d] - self.eqs[d] for d in range(nd)]
        r = [i[d] + self.eqs[d] + 1 for d in range(nd)]
        l_ok = [l[d] >= 1 for d in range(nd)]
        r_ok = [r[d] <= c.shape[d] for d in range(nd)]
        rng = lambda d: range(i[d] - self.eqs[d] + 1, i[d] + self.eqs[d] + 1)
        lvw = lambda d, j: self._kt(d, j, x[d]) - self.left_values[d][j - 1]
        at = lambda arr, idx: arr[tuple(k - 1 for k in idx)]

        def setidx(idx, **kw):
            idx = list(idx)
            for d, v in kw.items():
                idx[int(d[1:])] = v
            return idx

        result = 0.0
        deriv_ranges = [range(1, 2) if isint[d] else rng(d) for d in range(nd)]
        for idx_d in itertools.product(*deriv_ranges):
            kt_prod, skip = 1.0, False
            for d in range(nd):
                if isint[d]:
                    continue
                kv = self._kd(d, idx_d[d], x[d])
                if kv is None:
                    skip = True
                    break
                kt_prod *= kv
            if skip:
                continue
            S = lambda idx, **kw: setidx(idx, **{f"d{k}": v for k, v in kw.items()})
            for j3 in rng(d3):
                lv3 = lvw(d3, j3)
                idx3 = S(idx_d, **{str(d3): j3})
                for j2 in rng(d2):
                    lv2 = lvw(d2, j2)
                    lv23 = lv2 * lv3
                    idx23 = S(idx3, **{str(d2): j2})
                    for j1 in rng(d1):
                        lv1 = lvw(d1, j1)
                        result += at(c, S(idx23, **{str(d1): j1})) * lv1 * lv23 * kt_prod
                    tl1 = at(self.tail1_left[d1], S(idx23, **{str(d1): l[d1]})) if l_ok[d1] else 0.0
                    tr1 = at(self.tail1_right[d1], S(idx23, **{str(d1): r[d1]})) if r_ok[d1] else 0.0
                    result += (tl1 + tr1) * lv23 * kt_prod
                for j1 in rng(d1):
                    lv1 = lvw(d1, j1)
                    idx13 = S(idx3, **{str(d1): j1})
                    tl2 = at(self.tail1_left[d2], S(idx13, **{str(d2): l[d2]})) if l_ok[d2] else 0.0
                    tr2 = at(self.tail1_right[d2], S(idx13, **{str(d2): r[d2]})) if r_ok[d2] else 0.0
                    result += (tl2 + tr2) * lv1 * lv3 * kt_prod
                for (a1, ok1, a2, ok2, arr) in ((l, l_ok, l, l_ok, self.edge_ll[3]), (r, r_ok, l, l_ok, self.edge_rl[3]),
                                                (l, l_ok, r, r_ok, self.edge_lr[3]), (r, r_ok, r, r_ok, self.edge_rr[3])):
                    v = at(arr, S(idx3, **{str(d1): a1[d1], str(d2): a2[d2]})) if (ok1[d1] and ok2[d2]) else 0.0
                    result += v * lv3 * kt_prod
            for j2 in rng(d2):
                lv2 = lvw(d2, j2)
                for j1 in rng(d1):
                    lv1 = lvw(d1, j1)
                    idx12 = S(idx_d, **{str(d1): j1, str(d2): j2})
                    tl3 = at(self.tail1_left[d3], S(idx12, **{str(d3): l[d3]})) if l_ok[d3] else 0.0
                    tr3 = at(self.tail1_right[d3], S(idx12, **{str(d3): r[d3]})) if r_ok[d3] else 0.0
                    result += (tl3 + tr3) * lv1 * lv2 * kt_prod
            for j2 in rng(d2):
                lv2 = lvw(d2, j2)
                idx2 = S(idx_d, **{str(d2): j2})
                for (a1, ok1, a3, ok3, arr) in ((l, l_ok, l, l_ok, self.edge_ll[2]), (r, r_ok, l, l_ok, self.edge_rl[2]),
                                                (l, l_ok, r, r_ok, self.edge_lr[2]), (r, r_ok, r, r_ok, self.edge_rr[2])):
                    v = at(arr, S(idx2, **{str(d1): a1[d1], str(d3): a3[d3]})) if (ok1[d1] and ok3[d3]) else 0.0
                    result += v * lv2 * kt_prod
            for j1 in rng(d1):
                lv1 = lvw(d1, j1)
                idx1 = S(idx_d, **{str(d1): j1})
                for (a2, ok2, a3, ok3, arr) in ((l, l_ok, l, l_ok, self.edge_ll[1]), (r, r_ok, l, l_ok, self.edge_rl[1]),
                                                (l, l_ok, r, r_ok, self.edge_lr[1]), (r, r_ok, r, r_ok, self.edge_rr[1])):
                    v = at(arr, S(idx1, **{str(d2): a2[d2], str(d3): a3[d3]})) if (ok2[d2] and ok3[d3]) else 0.0
                    result += v * lv1 * kt_prod
            for name, arr in self.corner.items():
                sel = [(l, l_ok) if ch == "l" else (r, r_ok) for ch in name]
                if sel[0][1][d1] and sel[1][1][d2] and sel[2][1][d3]:
                    result += at(arr, S(idx_d, **{str(d1): sel[0][0][d1], str(d2): sel[1][0][d2], str(d3): sel[2][0][d3]})) * kt_prod
        return result * self._scale()

    def _core(self, x):
        return {"nd": self._eval_nd, "mixed3": self._eval_mixed3, "mixedN": self._eval_mixed_n}[self.functor](x)

    def eval_points(self, *xs):
        """the extrapolation wrapper (extrapolation_operator.jl:31-63, extrapolate_point.jl:98-160) around the functor"""
        xs = [np.asarray(x, dtype=np.float64).ravel() for x in xs]
        lo = np.array([k[e - 1] for k, e in zip(self.knots, self.eqs)])
        hi = np.array([k[-e] for k, e in zip(self.knots, self.eqs)])
        hmin = min(self.h)
        out = np.empty(xs[0].size)
        for q in range(xs[0].size):
            x = np.array([a[q] for a in xs])
            if np.all((x >= lo) & (x <= hi)):
                out[q] = self._core(x)
                continue
            if self.extrap == "throw":
                raise OracleError(f"query {q} outside the domain with extrap=throw")
            xc = np.clip(x, lo, hi)
            if self.extrap == "flat":
                out[q] = self._core(xc)
                continue
            dx = x - xc
            d2 = 0.0
            for v in dx:
                d2 += v * v
            if d2 < 2.220446049250313e-16:
                out[q] = self._core(xc)
                continue
            eps, dist = hmin / 100.0, math.sqrt(d2)
            fb = self._core(xc)
            fi = self._core(xc - eps * dx / dist)
            out[q] = fb + (fb - fi) / eps * dist
        return out

    eval = eval_points
