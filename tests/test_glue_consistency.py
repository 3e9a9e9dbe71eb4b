"""The Julia glue (julia/CIB200Glue.jl) cannot be executed here (no Julia in the image).  What can be checked without
Julia is checked mechanically: every `ccall` in the glue names a prototype of include/cib200.h and passes the same
number of arguments with the Julia type that corresponds to each C type; the `ItpDesc` mirror has the fields of
`cib200_itp_desc` in order; every reference name the glue uses is imported and defined in the reference sources."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cib200.h")
GLUE = os.path.join(ROOT, "julia", "CIB200Glue.jl")
REF = "/root/reference/src"


def strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def c_prototypes():
    src = strip_c_comments(open(HEADER).read())
    src = src[src.index('extern "C"'):]
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(cib200_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if "typedef" in ret:
            continue
        arglist = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, arglist)
    return protos


def c_type(arg):
    """C declaration of one parameter -> canonical type string without the parameter name."""
    a = re.sub(r"\b(const)\b", "", arg)
    a = a.replace("*", " * ")
    toks = a.split()
    if toks and re.fullmatch(r"[A-Za-z_]\w*", toks[-1]) and len(toks) > 1 and toks[-1] not in ("int32_t", "int64_t", "double", "void", "size_t", "char"):
        toks = toks[:-1]  # drop the parameter name
    return " ".join(toks)


JULIA_OF_C = {
    "int32_t": {"Int32"},
    "int64_t": {"Int64"},
    "double": {"Float64"},
    "size_t": {"Csize_t"},
    "int32_t *": {"Ptr{Int32}", "Ref{Int32}"},
    "int64_t *": {"Ptr{Int64}", "Ref{Int64}"},
    "double *": {"Ptr{Float64}"},
    "float *": {"Ptr{Float32}"},
    "void *": {"Ptr{Cvoid}"},
    "void * *": {"Ptr{Ptr{Cvoid}}"},
    "double * *": {"Ptr{Ptr{Float64}}"},
    "cib200_itp_desc *": {"Ref{ItpDesc}", "Ptr{ItpDesc}"},
    "cib200_itp *": {"Ptr{Cvoid}"},
    "cib200_itpnd *": {"Ptr{Cvoid}"},
    "cib200_itg *": {"Ptr{Cvoid}"},
    "cib200_itgnd *": {"Ptr{Cvoid}"},
    "cib200_itgnd * *": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "cib200_itp * *": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "cib200_itpnd * *": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "cib200_itg * *": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "char *": {"Cstring"},
}


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    src = re.sub(r"#[^\n]*", "", open(GLUE).read())
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\)\s*,", src):
        i, depth = m.start() + len("ccall"), 0
        j = i
        while True:
            if src[j] == "(":
                depth += 1
            elif src[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        parts = split_top(src[i + 1:j])
        # parts: (:name, LIB) | RetType | (ArgTypes) | args...
        ret = parts[1]
        argt = parts[2].strip()
        assert argt.startswith("(") and argt.endswith(")"), (m.group(1), argt)
        types = split_top(argt[1:-1])
        calls.append((m.group(1), ret, types, parts[3:]))
    return calls


def test_every_ccall_matches_a_header_prototype():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 15
    for name, ret, types, args in calls:
        assert name in protos, f"{name} is not declared in include/cib200.h"
        cret, cargs = protos[name]
        assert len(types) == len(cargs), f"{name}: {len(types)} Julia argument types for {len(cargs)} C parameters"
        assert len(args) == len(types), f"{name}: {len(args)} values for {len(types)} argument types"
        want_ret = JULIA_OF_C[c_type(cret + " x") if cret.endswith("*") else cret.replace("const", "").strip()
                              if cret.replace("const", "").strip() in JULIA_OF_C else c_type(cret + " x")]
        assert ret in want_ret, f"{name}: return {ret}, header says {cret}"
        for k, (jt, ca) in enumerate(zip(types, cargs)):
            ct = c_type(ca)
            assert ct in JULIA_OF_C, f"{name}: no mapping for C type '{ct}' ({ca})"
            assert jt in JULIA_OF_C[ct], f"{name} argument {k} ({ca}): Julia passes {jt}, expected one of {sorted(JULIA_OF_C[ct])}"


def test_itp_desc_mirror_matches_the_header_struct():
    hdr = strip_c_comments(open(HEADER).read())
    body = re.search(r"typedef struct cib200_itp_desc \{(.*?)\} cib200_itp_desc;", hdr, flags=re.S).group(1)
    cfields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(.*?)(\w+)(\[CIB200_MAX_ITP_DIM\])?$", decl)
        cfields.append((m.group(2), c_type(m.group(1) + " x"), bool(m.group(3))))
    glue = re.sub(r"#[^\n]*", "", open(GLUE).read())
    jbody = re.search(r"struct ItpDesc\n(.*?)\nend", glue, flags=re.S).group(1)
    jfields = [tuple(x.strip() for x in ln.split("::")) for ln in jbody.strip().splitlines()]
    assert [f[0] for f in jfields] == [f[0] for f in cfields]
    scalar = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "void *": "Ptr{Cvoid}", "double *": "Ptr{Float64}"}
    for (jn, jt), (cn, ct, arr) in zip(jfields, cfields):
        want = scalar[ct]
        assert jt == (f"NTuple{{3,{want}}}" if arr else want), f"field {cn}: {jt} vs {ct}{'[3]' if arr else ''}"


def test_glue_imports_every_reference_name_it_uses():
    glue = re.sub(r"#[^\n]*", "", open(GLUE).read())
    imp = re.search(r"using \.\.ConvolutionInterpolations:(.*?)\n\n", glue, flags=re.S).group(1)
    imported = {x.strip() for x in imp.replace("\n", " ").split(",") if x.strip()}
    # names of the reference the glue mentions anywhere
    used = set(re.findall(r"\b(ConvolutionExtrapolation|FastConvolutionInterpolation|HigherOrderKernel|HigherOrderMixedKernel|"
                          r"Throw|Flat|Line|DerivativeOrder|FastIntegralOrder|FastMixedIntegralOrder|"
                          r"get_precomputed_kernel_and_range)\b", glue))
    assert used <= imported, f"used but not imported: {sorted(used - imported)}"
    # no WeakKeyDict on immutable reference structs (finalizers cannot be attached to them)
    assert "WeakKeyDict" not in glue
    if not os.path.isdir(REF):
        pytest.skip("reference sources not present on this box")
    defined = ""
    for dp, _, fs in os.walk(REF):
        for f in fs:
            if f.endswith(".jl"):
                defined += open(os.path.join(dp, f), errors="ignore").read()
    for name in sorted(imported):
        assert re.search(rf"\b(struct|function|abstract type)\s+{name}\b|\b{name}\(", defined), f"{name} not found in the reference sources"
    # fields of FastConvolutionInterpolation / ConvolutionExtrapolation the glue reads
    fields = set(re.findall(r"\bitp\.(\w+)", glue)) | {"itp", "et"}
    struct_src = open(os.path.join(REF, "data_structures", "fast_convolution_interpolation.jl")).read() + \
        open(os.path.join(REF, "data_structures", "convolution_extrapolation.jl")).read()
    for f in sorted(fields):
        assert re.search(rf"\b{f}::", struct_src), f"field {f} is not a field of the reference structs"
