# julia_dump.jl - runs the UNMODIFIED reference (ConvolutionInterpolations.jl) on the seeded cases written by
# tools/julia_cases.py and writes its outputs next to them, so that any box with Julia turns the parity of this
# repository from "against a restatement" into "against the reference itself".
#
#   python tools/julia_cases.py write cases/            # inputs: knots, values, queries (raw little-endian Float64)
#   julia --project=/path/to/ConvolutionInterpolations.jl tools/julia_dump.jl cases/
#   python tools/julia_cases.py check cases/ [--gpu]    # oracle (and CUDA library) against out.f64 of every case
#
# No package beyond the reference is needed (the case files are key=value text + raw arrays).  Never executed in the
# build container (no Julia there); tools/julia_stub.py plays this script's role in the tests so that the file format,
# the loader and the comparison are exercised end to end.
using ConvolutionInterpolations

readvec(path) = reinterpret(Float64, read(path)) |> collect
function readspec(path)
    d = Dict{String,String}()
    for ln in eachline(path)
        isempty(strip(ln)) && continue
        k, v = split(ln, "="; limit=2)
        d[strip(k)] = strip(v)
    end
    return d
end
ints(s) = parse.(Int, split(s, ","))
syms(s) = Symbol.(split(s, ","))
tup(v) = length(v) == 1 ? v[1] : Tuple(v)

function run_case(dir)
    sp = readspec(joinpath(dir, "spec.txt"))
    nd = parse(Int, sp["ndim"])
    n = ints(sp["n"])
    knots = ntuple(d -> readvec(joinpath(dir, "knots_$(d-1).f64")), nd)
    values = reshape(readvec(joinpath(dir, "values.f64")), n...)       # column-major, as Julia stores it
    kind = sp["kind"]
    out = if kind == "interp"
        extrap = Dict("throw" => Throw(), "flat" => Flat(), "line" => Line())[sp["extrap"]]
        itp = convolution_interpolation(nd == 1 ? knots[1] : knots, values; kernel=tup(syms(sp["kernel"])),
                                        derivative=tup(ints(sp["derivative"])), bc=Symbol(sp["bc"]), extrap=extrap,
                                        subgrid=tup(syms(sp["subgrid"])))
        xs = ntuple(d -> readvec(joinpath(dir, "x_$(d-1).f64")), nd)
        itp.(xs...)
    elseif kind == "resample"
        kout = ntuple(d -> readvec(joinpath(dir, "knots_out_$(d-1).f64")), nd)
        convolution_resample(knots, kout, values; kernel=tup(syms(sp["kernel"])), derivative=tup(ints(sp["derivative"])),
                             bc=Symbol(sp["bc"]), subgrid=tup(syms(sp["subgrid"])))
    elseif kind == "smooth"
        convolution_smooth(knots, values, parse(Float64, sp["B"]))
    else
        error("unknown kind $kind")
    end
    write(joinpath(dir, "out.f64"), Float64.(vec(out)))
    # the ghost-padded grid and the extended knots as the reference built them (pins the constructor separately)
    if kind == "interp"
        itp0 = convolution_interpolation(nd == 1 ? knots[1] : knots, values; kernel=tup(syms(sp["kernel"])),
                                         derivative=tup(ints(sp["derivative"])), bc=Symbol(sp["bc"]),
                                         subgrid=tup(syms(sp["subgrid"])))
        write(joinpath(dir, "coefs.f64"), Float64.(vec(itp0.itp.coefs)))
    end
    println("  wrote ", joinpath(dir, "out.f64"), " (", length(out), " values)")
end

root = length(ARGS) >= 1 ? ARGS[1] : "cases"
for name in sort(readdir(root))
    dir = joinpath(root, name)
    isfile(joinpath(dir, "spec.txt")) || continue
    println(name)
    run_case(dir)
end
write(joinpath(root, "JULIA_VERSION"), string(VERSION, " ConvolutionInterpolations ", pkgversion(ConvolutionInterpolations)))
