# CIB200Glue.jl - the thin `ccall` glue a maintainer adds to ConvolutionInterpolations.jl to route the
# uniform-grid hot path to libcib200.so (C-ABI: include/cib200.h).  UNTESTED: Julia is not installed in this
# repository's build container, so this file has never been executed.  What is checked mechanically
# (tests/test_glue_consistency.py): every `ccall` names an exported symbol of include/cib200.h and passes the
# prototype's argument count and C types, the ItpDesc mirror has the fields of cib200_itp_desc in order, and every
# reference name used below is imported and exists in /root/reference/src.  The same ABI is exercised end to end by
# the ctypes binding in convolutioninterpolations.jl_b200/_lib.py.
#
# Nothing of the public API changes: the methods below are added on the reference's own types.
#   include("CIB200Glue.jl") at the end of src/ConvolutionInterpolations.jl, `using .CIB200Glue`.
module CIB200Glue

using ..ConvolutionInterpolations: ConvolutionExtrapolation, FastConvolutionInterpolation,
    HigherOrderKernel, HigherOrderMixedKernel, Throw, Flat, Line, DerivativeOrder,
    FastIntegralOrder, FastMixedIntegralOrder, get_precomputed_kernel_and_range

const LIB = get(ENV, "CIB200_LIB", "libcib200.so")
const KERNEL_ID = Dict(:b5 => 0, :b7 => 1, :b9 => 2, :b11 => 3, :b13 => 4, :a3 => 5, :a4 => 6, :a5 => 7, :a7 => 8)
const BC_ID = Dict(:detect => 0, :poly => 1, :linear => 2, :quadratic => 3)
const SUBGRID_ID = Dict(:cubic => 0, :quintic => 1, :linear => 2)
const MAX_TABLE = Dict(:b5 => 3, :b7 => 4, :b9 => 5, :b11 => 6, :b13 => 6, :a3 => 1, :a4 => 1, :a5 => 1, :a7 => 1)  # shipped K^(d)
const MAX_DERIV = Dict(:b5 => 2, :b7 => 3, :b9 => 4, :b11 => 5, :b13 => 5, :a3 => 0, :a4 => 0, :a5 => 0, :a7 => 0)   # _max_shipped_derivative

# mirror of cib200_itp_desc (include/cib200.h) -------------------------------------------------------
struct ItpDesc
    ndim::Int32
    dtype::Int32
    coef_shape::NTuple{3,Int64}
    coefs::Ptr{Cvoid}
    knots::NTuple{3,Ptr{Float64}}
    h::NTuple{3,Float64}
    kernel::NTuple{3,Int32}
    deriv::NTuple{3,Int32}
    extrap::Int32
    perdim::Int32
    subgrid::NTuple{3,Int32}
end

last_error() = unsafe_string(ccall((:cib200_last_error, LIB), Cstring, ()))

function check(status::Int32)
    status == 0 && return
    msg = last_error()
    status == 1 && error(msg)                    # ErrorException, as extrapolate_point.jl:156-159
    status == 2 && throw(InexactError(:floor, Int, NaN))  # what floor(Int, NaN) raises in the reference
    error("cib200 status $status: $msg")
end

pad3(t::NTuple{N,T}, fill) where {N,T} = ntuple(d -> d <= N ? t[d] : fill, 3)

# ---- which interpolants take the accelerated path (everything else keeps the Julia methods) ----------
kernel_symbols(::HigherOrderKernel{KS}) where {KS} = KS        # struct_constructors.jl:76-80
kernel_symbols(::HigherOrderMixedKernel{KS}) where {KS} = KS
kernel_symbols(_) = nothing

# antiderivative on every axis (FastIntegralOrder, convolution_fast_integration_{1d,2d,3d}.jl) or on some of them
# (FastMixedIntegralOrder{DO}, convolution_fast_mixed_{1_integral_dim,2_integral_dims}.jl): cib200_itg_*
is_integral(itp::FastConvolutionInterpolation) =
    itp.derivative_order isa FastIntegralOrder || itp.derivative_order isa FastMixedIntegralOrder
integral_derivs(itp::FastConvolutionInterpolation{T,N}) where {T,N} =
    itp.derivative_order isa FastIntegralOrder ? ntuple(_ -> -1, N) : typeof(itp.derivative_order).parameters[1]

function accelerated(itp::FastConvolutionInterpolation{T,N}) where {T,N}
    (T === Float64 || T === Float32) || return false
    1 <= N <= 8 || return false                                          # CIB200_MAX_ND
    ks = kernel_symbols(itp.kernel_sym)
    ks === nothing && return false
    all(k -> haskey(KERNEL_ID, k), ks) || return false
    itp.lazy === Val{false}() || return false
    if is_integral(itp) && N >= 4
        # N >= 4: every axis :linear on generated tables; three or more integral axes (convolution_fast_integration_nd.jl,
        # convolution_fast_mixed_3_integral_dims.jl, convolution_fast_mixed_N_integral_dims.jl): cib200_itgnd_*
        dvi = integral_derivs(itp)
        return T === Float64 && count(==(-1), dvi) >= 3 && all(==(:linear), subgrid_tuple(itp)) &&
               all(d -> dvi[d] == -1 || 0 <= dvi[d] <= MAX_TABLE[ks[d]], 1:N)
    end
    if is_integral(itp)
        # shipped K~ / K / K' tables only: Float64, N <= 3, cubic or quintic subgrid on every axis (mixed integral /
        # interpolation axes resolve to :linear on generated tables, _build_subgrids)
        dvi = integral_derivs(itp)
        sgi = subgrid_tuple(itp)
        return T === Float64 && N <= 3 &&
               all(d -> dvi[d] == -1 ? (length(itp.pre_range[d]) == 101 && (sgi[d] === :cubic || sgi[d] === :quintic)) :
                                       (sgi[d] === :linear && 0 <= dvi[d] <= MAX_TABLE[ks[d]]), 1:N)
    end
    itp.derivative_order isa DerivativeOrder || return false
    dv = derivative_tuple(itp)
    sg = effective_subgrids(subgrid_tuple(itp), N)
    # cubic / quintic axes read the shipped 101-row tables; :linear axes carry generated tables of >= 10 000 rows
    # (constructor.jl:117-118) which are handed to the library as they are (cib200_itp_create_tables)
    all(d -> sg[d] === :linear || length(itp.pre_range[d]) == 101, 1:N) || return false
    all(s -> haskey(SUBGRID_ID, s), sg) || return false
    # tables a subgrid needs: K^(d) (:linear), + K^(d+1) (:cubic), + K^(d+2) (:quintic)
    need(d) = dv[d] + (sg[d] === :linear ? 0 : sg[d] === :cubic ? 1 : 2)
    all(d -> dv[d] >= 0 && need(d) <= MAX_TABLE[ks[d]], 1:N) || return false
    # evaluators: eager_2d.jl has (linear, linear), (cubic, cubic), (quintic, quintic); the mixed pairs exist on the
    # per-dim path (perdim_2d.jl); :quintic never in 3-D (_build_subgrids)
    any(==(:quintic), sg) && N > 2 && return false
    N >= 4 && !all(==(:linear), sg) && return false                      # HigherDimension: :linear only (_build_subgrids)
    N == 2 && !(itp.kernel_sym isa HigherOrderMixedKernel) && !allequal(sg) && return false
    return true
end
# in 3-D anything but all-cubic is evaluated with linear blending on every axis (eager_3d.jl:139-176, perdim_3d.jl:161)
effective_subgrids(sg, N) = (N == 3 && any(!=(:cubic), sg)) ? ntuple(_ -> :linear, 3) : sg
subgrid_tuple(itp) = typeof(itp.subgrid).parameters[1]               # Val{SG}
derivative_tuple(itp) = typeof(itp.derivative_order).parameters[1]   # DerivativeOrder{DO} (struct_constructors.jl:182)

# itp.kernel_pre[d] of the :linear axes (n_pre x 2 eqs, column-major: passed as it lies in memory), C_NULL elsewhere
function linear_tables(itp::FastConvolutionInterpolation{T,N}, sg) where {T,N}
    tabs64 = [sg[d] === :linear ? Matrix{Float64}(itp.kernel_pre[d]) : Matrix{Float64}(undef, 0, 0) for d in 1:N]
    tptr = [sg[d] === :linear ? pointer(tabs64[d]) : Ptr{Float64}(0) for d in 1:N]
    npre = Int32[sg[d] === :linear ? size(tabs64[d], 1) : 0 for d in 1:N]
    return tabs64, tptr, npre
end

# ---- handle cache: one upload per interpolant, freed by a finalizer --------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    replicas::Vector{Ptr{Cvoid}}   # one handle per GPU (1-3 D): cib200_itp_eval_multi splits host batches over them
end
Handle(p::Ptr{Cvoid}) = Handle(p, Ptr{Cvoid}[p])
# GPUs to use from this process: ENV["CIB200_DEVICES"] = "0,1,2,3" (default: the current device only)
devices() = haskey(ENV, "CIB200_DEVICES") ? parse.(Int32, split(ENV["CIB200_DEVICES"], ",")) : Int32[]
# ConvolutionExtrapolation and FastConvolutionInterpolation are immutable structs (no finalizers, no WeakKeyDict keys):
# the cache is keyed on the identity of the one mutable object an interpolant owns for its whole life, its coefficient
# array, plus the extrapolation mode; a finalizer on that array drops the entry and frees the device memory.
const HANDLES = Dict{Tuple{UInt,Int32},Handle}()
const HANDLES_LOCK = ReentrantLock()
function release_handles(coefs)
    # finalizers must not block: if the lock is busy, try again at the next collection (Julia manual, "finalizer")
    if trylock(HANDLES_LOCK)
        try
            id = objectid(coefs)
            for key in [k for k in keys(HANDLES) if k[1] == id]
                finalize(pop!(HANDLES, key))
            end
        finally
            unlock(HANDLES_LOCK)
        end
    else
        finalizer(release_handles, coefs)
    end
    return nothing
end

extrap_id(::Throw) = Int32(0)
extrap_id(::Flat) = Int32(1)
extrap_id(::Line) = Int32(2)

function handle(etp::ConvolutionExtrapolation{T,N}) where {T,N}
    lock(HANDLES_LOCK) do
        key = (objectid(etp.itp.coefs), extrap_id(etp.et))
        fresh = !any(k -> k[1] == key[1], keys(HANDLES))
        hd = get!(HANDLES, key) do
            itp = etp.itp
            ks = kernel_symbols(itp.kernel_sym)
            dv = derivative_tuple(itp)
            knots64 = ntuple(d -> Vector{Float64}(itp.knots[d]), N)
            if N >= 4 && is_integral(itp)
                # antiderivative / mixed interpolants in N >= 4 dimensions: cib200_itgnd_*.  left_values and tails of the
                # Julia struct are not read; the every-axis-integral functor visits 1:domain_size[d] only
                dvi = integral_derivs(itp)
                out = Ref{Ptr{Cvoid}}(C_NULL)
                cshape = Int64[size(itp.coefs)...]
                kptr = [pointer(knots64[d]) for d in 1:N]
                hh = Float64[itp.h...]
                kid = Int32[KERNEL_ID[ks[d]] for d in 1:N]
                dvv = Int32[dvi[d] for d in 1:N]
                tabs64, tptr, npre = linear_tables(itp, ntuple(_ -> :linear, N))
                ext = Int64[(itp.derivative_order isa FastIntegralOrder ? itp.domain_size[d] : size(itp.coefs, d)) for d in 1:N]
                coefs64 = Array{Float64,N}(itp.coefs)
                GC.@preserve coefs64 knots64 cshape kptr hh kid dvv tabs64 tptr npre ext check(ccall((:cib200_itgnd_create, LIB), Int32,
                    (Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Int32,
                     Ptr{Ptr{Float64}}, Ptr{Int32}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                    N, cshape, coefs64, kptr, hh, kid, dvv, extrap_id(etp.et), tptr, npre, ext, out))
                h = Handle(out[])
                finalizer(x -> ccall((:cib200_itgnd_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
                return h
            end
            if N >= 4
                # HigherDimension evaluators (eager_nd.jl:97-140 / perdim_nd.jl:4-36): cib200_itpnd_*
                out = Ref{Ptr{Cvoid}}(C_NULL)
                cshape = Int64[size(itp.coefs)...]
                kptr = [pointer(knots64[d]) for d in 1:N]
                hh = Float64[itp.h...]
                kid = Int32[KERNEL_ID[ks[d]] for d in 1:N]
                dvv = Int32[dv[d] for d in 1:N]
                tabs64, tptr, npre = linear_tables(itp, ntuple(_ -> :linear, N))
                GC.@preserve itp knots64 cshape kptr hh kid dvv tabs64 tptr npre check(ccall((:cib200_itpnd_create_tables, LIB), Int32,
                    (Int32, Int32, Ptr{Int64}, Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32},
                     Int32, Int32, Ptr{Ptr{Float64}}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                    N, T === Float32 ? 1 : 0, cshape, pointer(itp.coefs), kptr, hh, kid, dvv, extrap_id(etp.et),
                    Int32(itp.kernel_sym isa HigherOrderMixedKernel), tptr, npre, out))
                h = Handle(out[])
                finalizer(x -> ccall((:cib200_itpnd_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
                return h
            end
            if is_integral(itp)
                # the library builds left_values and the prefix-sum tails itself (constructor.jl:176-364, :414-448) from
                # itp.coefs; itp.tail* of the Julia struct are not read
                dvi = integral_derivs(itp)
                desc = Ref(ItpDesc(Int32(N), Int32(0), pad3(Int64.(size(itp.coefs)), Int64(1)), pointer(itp.coefs),
                                   pad3(ntuple(d -> pointer(knots64[d]), N), Ptr{Float64}(0)),
                                   pad3(Float64.(itp.h), 1.0), pad3(ntuple(d -> Int32(KERNEL_ID[ks[d]]), N), Int32(0)),
                                   pad3(ntuple(d -> Int32(dvi[d]), N), Int32(0)), extrap_id(etp.et), Int32(0),
                                   pad3(ntuple(d -> Int32(SUBGRID_ID[subgrid_tuple(itp)[d]]), N), Int32(0))))
                # the :linear (non-integral) axes of a mixed interpolant bring their generated tables
                tabs64, tptr, npre = linear_tables(itp, ntuple(d -> dvi[d] == -1 ? :integral : :linear, N))
                out = Ref{Ptr{Cvoid}}(C_NULL)
                GC.@preserve itp knots64 tabs64 tptr npre check(ccall((:cib200_itg_create_tables, LIB), Int32,
                    (Ref{ItpDesc}, Ptr{Ptr{Float64}}, Ptr{Int32}, Ref{Ptr{Cvoid}}), desc, tptr, npre, out))
                h = Handle(out[])
                finalizer(x -> ccall((:cib200_itg_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
                return h
            end
            desc = Ref(ItpDesc(Int32(N), T === Float32 ? Int32(1) : Int32(0),
                               pad3(Int64.(size(itp.coefs)), Int64(1)), pointer(itp.coefs),
                               pad3(ntuple(d -> pointer(knots64[d]), N), Ptr{Float64}(0)),
                               pad3(Float64.(itp.h), 1.0), pad3(ntuple(d -> Int32(KERNEL_ID[ks[d]]), N), Int32(0)),
                               pad3(ntuple(d -> Int32(dv[d]), N), Int32(0)), extrap_id(etp.et),
                               Int32(itp.kernel_sym isa HigherOrderMixedKernel),
                               pad3(ntuple(d -> Int32(SUBGRID_ID[effective_subgrids(subgrid_tuple(itp), N)[d]]), N), Int32(0))))
            tabs64, tptr, npre = linear_tables(itp, effective_subgrids(subgrid_tuple(itp), N))
            devs = devices()
            ptrs = Ptr{Cvoid}[]
            for dev in (isempty(devs) ? Int32[-1] : devs)
                dev >= 0 && check(ccall((:cib200_set_device, LIB), Int32, (Int32,), dev))
                out = Ref{Ptr{Cvoid}}(C_NULL)
                GC.@preserve itp knots64 tabs64 tptr npre check(ccall((:cib200_itp_create_tables, LIB), Int32,
                    (Ref{ItpDesc}, Ptr{Ptr{Float64}}, Ptr{Int32}, Ref{Ptr{Cvoid}}), desc, tptr, npre, out))
                push!(ptrs, out[])
            end
            h = Handle(ptrs[1], ptrs)
            finalizer(x -> foreach(p -> ccall((:cib200_itp_destroy, LIB), Int32, (Ptr{Cvoid},), p), x.replicas), h)
            h
        end
        fresh && finalizer(release_handles, etp.itp.coefs)
        hd
    end
end

# ---- query / result buffers -------------------------------------------------------------------------
# A coordinate vector that already is a dense Vector{T} is passed as it lies in memory (no copy); anything else
# (ranges, views, other element types) is materialised once.  The library moves pageable host memory through its own
# pinned staging ring (cib200_itp_eval, memkind HOST), so nothing has to be page-locked for a one-off call.  Buffers
# that are re-used across many calls can be page-locked once with `pin!` (cib200_host_register: the copies then run
# straight from the caller's memory at PCIe rate); `unpin!` before the array is resized or dropped.
as_dense(::Type{T}, x::Vector{T}) where {T} = x
as_dense(::Type{T}, x::AbstractVector) where {T} = Vector{T}(x)
function pin!(x::Array)
    GC.@preserve x check(ccall((:cib200_host_register, LIB), Int32, (Ptr{Cvoid}, Csize_t), x, sizeof(x)))
    return x
end
function unpin!(x::Array)
    GC.@preserve x check(ccall((:cib200_host_unregister, LIB), Int32, (Ptr{Cvoid},), x))
    return x
end

# ---- itp.(xs...) : replaces the per-element loop over extrapolation_operator.jl:31-63 ---------------
function Base.Broadcast.broadcasted(etp::ConvolutionExtrapolation{T,N}, xs::Vararg{AbstractVector{<:Real},N}) where {T,N}
    accelerated(etp.itp) || return Base.Broadcast.broadcasted(Base.Broadcast.DefaultArrayStyle{1}(), etp, xs...)
    nq = length(xs[1])
    all(x -> length(x) == nq, xs) || throw(DimensionMismatch("coordinate vectors must have equal length"))
    xq = ntuple(d -> as_dense(T, xs[d]), N)
    out = Vector{T}(undef, nq)
    ptrs = [Ptr{Cvoid}(pointer(xq[d])) for d in 1:N]
    bad = Ref{Int64}(-1)
    hd = handle(etp)
    if length(hd.replicas) > 1     # several GPUs: contiguous slices, one pipeline per device
        GC.@preserve xq out ptrs check(ccall((:cib200_itp_eval_multi, LIB), Int32,
            (Ptr{Ptr{Cvoid}}, Int32, Int64, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ref{Int64}),
            hd.replicas, length(hd.replicas), nq, ptrs, out, bad))
        return out
    end
    hp = hd.ptr
    if is_integral(etp.itp) && N >= 4
        GC.@preserve xq out ptrs check(ccall((:cib200_itgnd_eval, LIB), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
            hp, nq, ptrs, out, 0, C_NULL, bad))
        return out
    end
    if is_integral(etp.itp)
        GC.@preserve xq out ptrs check(ccall((:cib200_itg_eval, LIB), Int32,
            (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
            hp, nq, ptrs, out, 0, C_NULL, bad))
        return out
    end
    GC.@preserve xq out ptrs check(N >= 4 ?   # ccall needs a literal symbol: one call site per entry point
        ccall((:cib200_itpnd_eval, LIB), Int32,
              (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
              hp, nq, ptrs, out, 0, C_NULL, bad) :
        ccall((:cib200_itp_eval, LIB), Int32,
              (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
              hp, nq, ptrs, out, 0, C_NULL, bad))
    return out
end

# ---- itp(xs, ys, zs) with vectors: Cartesian product, extrapolation_operator.jl:66-99 ----------------
function (etp::ConvolutionExtrapolation{T,N})(xs::Vararg{AbstractVector{<:Real},N}) where {T,N}
    accelerated(etp.itp) || return invoke(etp, Tuple{Vararg{AbstractArray,N}}, xs...)
    xq = ntuple(d -> as_dense(T, xs[d]), N)
    n = Int64[length(x) for x in xq]
    out = zeros(T, n...)
    ptrs = [Ptr{Cvoid}(pointer(xq[d])) for d in 1:N]
    bad = Ref{Int64}(-1)
    hp = handle(etp).ptr
    if is_integral(etp.itp) && N >= 4
        GC.@preserve xq out ptrs check(ccall((:cib200_itgnd_eval_grid, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
            hp, n, ptrs, out, 0, C_NULL, bad))
        return out
    end
    if is_integral(etp.itp)
        GC.@preserve xq out ptrs check(ccall((:cib200_itg_eval_grid, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
            hp, n, ptrs, out, 0, C_NULL, bad))
        return out
    end
    GC.@preserve xq out ptrs check(N >= 4 ?
        ccall((:cib200_itpnd_eval_grid, LIB), Int32,
              (Ptr{Cvoid}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
              hp, n, ptrs, out, 0, C_NULL, bad) :
        ccall((:cib200_itp_eval_grid, LIB), Int32,
              (Ptr{Cvoid}, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ref{Int64}),
              hp, n, ptrs, out, 0, C_NULL, bad))
    return out
end

# ---- convolution_resample_internal (convolution_resample.jl:240-386) ---------------------------------
function resample_accelerated(values::AbstractArray{T,N}, kernels, derivatives, subgrids, precompute) where {T,N}
    (T === Float64 || T === Float32) && N <= 8 && all(k -> haskey(KERNEL_ID, k), kernels) &&
        all(s -> haskey(SUBGRID_ID, s), subgrids) &&
        all(d -> subgrids[d] === :linear || precompute[d] == 101, 1:N) &&   # :linear axes: generated tables, passed below
        all(d -> derivatives[d] >= 0 && derivatives[d] + (subgrids[d] === :linear ? 0 : subgrids[d] === :cubic ? 1 : 2) <=
                 MAX_TABLE[kernels[d]], 1:N)
end

function cib200_resample(knots_in::NTuple{N}, knots_out::NTuple{N}, values::Array{T,N},
                         kernels, derivatives, bcs, subgrids, precompute) where {T,N}
    kin = ntuple(d -> Vector{Float64}(knots_in[d]), N)
    kout = ntuple(d -> Vector{Float64}(knots_out[d]), N)
    out = Array{T,N}(undef, ntuple(d -> length(kout[d]), N))
    n_in = Int64[size(values)...]; n_out = Int64[size(out)...]
    pin = [pointer(k) for k in kin]; pout = [pointer(k) for k in kout]
    kid = Int32[KERNEL_ID[k] for k in kernels]; dv = Int32[derivatives...]
    bl = Int32[BC_ID[b[1]] for b in bcs]; br = Int32[BC_ID[b[2]] for b in bcs]
    sg = Int32[SUBGRID_ID[s] for s in subgrids]
    # :linear axes: the table convolution_resample_internal itself would fetch (:279-280; precompute[d] >= 10 000 from
    # the kw method, 101 from the kw-less wrappers), column-major as it lies in memory
    tabs = [subgrids[d] === :linear ?
            Matrix{Float64}(get_precomputed_kernel_and_range(kernels[d], precompute[d], T, derivatives[d], :linear)[2]) :
            Matrix{Float64}(undef, 0, 0) for d in 1:N]
    tptr = [subgrids[d] === :linear ? pointer(tabs[d]) : Ptr{Float64}(0) for d in 1:N]
    npre = Int32[subgrids[d] === :linear ? size(tabs[d], 1) : 0 for d in 1:N]
    GC.@preserve kin kout values out tabs tptr npre check(ccall((:cib200_resample_tables, LIB), Int32,
        (Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Float64}}, Ptr{Ptr{Float64}},
         Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Ptr{Float64}}, Ptr{Int32}, Int32, Ptr{Cvoid}),
        N, T === Float32 ? 1 : 0, n_in, n_out, values, out, pin, pout, kid, dv, bl, br, sg, tptr, npre, 0, C_NULL))
    return out
end
# in convolution_resample_internal, first line:
#   resample_accelerated(values, kernels, derivatives, subgrids, precompute) &&
#       return CIB200Glue.cib200_resample(knots_in, knots_out, values, kernels, derivatives, bcs, subgrids, precompute)

# ---- convolution_smooth (convolution_gaussian_smoothing.jl:45-126) -----------------------------------
function cib200_smooth(knots::NTuple{N,AbstractVector}, values::Array{T,N}, B::Float64) where {T,N}
    out = similar(values)
    n = Int64[size(values)...]
    h = Float64[T(knots[d][2] - knots[d][1]) for d in 1:N]
    st = GC.@preserve values out ccall((:cib200_smooth, LIB), Int32,
        (Int32, Int32, Ptr{Int64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Float64, Int32, Ptr{Cvoid}),
        N, T === Float32 ? 1 : 0, n, values, out, h, B, 0, C_NULL)
    st == 3 && error(last_error())   # "Gaussian stencil half-width (E points) exceeds grid size ..." (:50-56)
    check(st)
    return out
end
# in convolution_smooth(knots::NTuple{N}, values::AbstractArray{T,N}, B), first line:
#   (T === Float64 || T === Float32) && N <= 8 && return CIB200Glue.cib200_smooth(knots, Array(values), B)

end # module
