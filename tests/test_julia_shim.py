"""Julia is not installed in the build image, so julia/GeostatInversionB200.jl cannot be run here.
This test parses every `ccall` of the shim and checks it against include/gsi_b200.h: the symbol
exists, the return type is Int32 (Cstring for gsi_last_error_string), and the argument count and
argument classes (int32 / int64 / double / pointer) match the C declaration -- so the shim stays a
mechanical transcription of the header.  It also checks that the shim covers the call surface
SURVEY.md §8(b) lists."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsi_b200.h")
SHIM = os.path.join(ROOT, "julia", "GeostatInversionB200.jl")


def _split_top(s):
    """Split on top-level commas (ignores commas inside (), {} and [])."""
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


def _c_class(t):
    t = t.strip()
    if "*" in t:
        return "ptr"
    t = re.sub(r"\bconst\b", "", t).split()
    base = t[0] if t else ""
    return {"int32_t": "i32", "int64_t": "i64", "double": "f64", "void": "void"}.get(base, "?" + base)


def header_decls():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(gsi_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if args in ("", "void"):
            classes = []
        else:
            classes = []
            for a in _split_top(args):
                a = re.sub(r"\s+", " ", a)
                # drop the parameter name: keep everything up to the last identifier
                tm = re.match(r"(.*?)(\b\w+)$", a)
                typ = tm.group(1) if tm and ("*" in tm.group(1) or " " in a) else a
                classes.append(_c_class(typ))
        decls[name] = ("ptr" if "*" in ret else _c_class(ret), classes)
    return decls


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring",):
        return "ptr"
    return {"Int32": "i32", "Int64": "i64", "Float64": "f64", "Cvoid": "void"}.get(t, "?" + t)


def shim_ccalls():
    src = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(gsi_\w+),\s*LIB\),\s*", src):
        # scan the balanced parenthesis of this ccall
        i = src.index("(", m.start())
        depth, j = 0, i
        while True:
            if src[j] == "(":
                depth += 1
            elif src[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        parts = _split_top(src[i + 1:j])
        ret, argt, vals = parts[1], parts[2], parts[3:]
        assert argt.startswith("(") and argt.endswith(")"), (m.group(1), argt)
        types = [t for t in _split_top(argt[1:-1]) if t]
        calls.append((m.group(1), ret.strip(), types, vals))
    return calls


def test_every_ccall_matches_the_header():
    decls = header_decls()
    assert len(decls) >= 40
    calls = shim_ccalls()
    assert len(calls) >= 30
    for name, ret, types, vals in calls:
        assert name in decls, f"{name}: not declared in gsi_b200.h"
        cret, cargs = decls[name]
        assert _jl_class(ret) == cret, f"{name}: return {ret} vs C {cret}"
        assert len(types) == len(cargs) == len(vals), f"{name}: {len(types)} Julia types, {len(vals)} values, {len(cargs)} C parameters"
        for k, (jt, ct) in enumerate(zip(types, cargs)):
            assert _jl_class(jt) == ct, f"{name}: argument {k + 1} is {jt} in the shim, {ct} in the header"


def test_shim_binds_the_surface_of_survey_8b():
    bound = {c[0] for c in shim_ccalls()}
    needed = {"gsi_ctx_create", "gsi_ctx_destroy", "gsi_comm_unique_id", "gsi_buf_alloc", "gsi_buf_free", "gsi_buf_upload",
              "gsi_buf_download", "gsi_op_dense", "gsi_op_lowrankcov", "gsi_op_kernelcov", "gsi_op_kernelcov_grid",
              "gsi_op_free", "gsi_op_size", "gsi_op_apply", "gsi_rangefinder_fixed", "gsi_randsvd",
              "gsi_rangefinder_adaptive", "gsi_rangefinder_adaptive_blocked", "gsi_eig_nystrom", "gsi_fftrf_powerlaw",
              "gsi_pcga_lowrank_matvec", "gsi_pcga_lsqr_solve", "gsi_pcga_direct_solve", "gsi_pcga_update",
              "gsi_pcga_paramstorun", "gsi_sketch_apply", "gsi_sketch_cov", "gsi_last_error_string"}
    assert needed <= bound, sorted(needed - bound)
    src = open(SHIM).read()
    for sig in ("function randsvd(A, K::Int, p::Int, q::Int", "function rangefinder(A, l::Int64, numiterations::Int64)",
                "function rangefinder(A; epsilon::Float64=1e-8, r::Int=10", "function eig_nystrom(A, Q::Matrix{Float64})",
                "function getxis(::Type{Val{:iwantfields}}, samplefield", "function getxis(samplefield",
                "pcgalsqr(forwardmodel::Function", "pcgadirect(forwardmodel::Function", "const pcga = pcgadirect",
                "function rga(forwardmodel::Function", "Base.:*(Bt::LinearAlgebra.Adjoint{Float64, Matrix{Float64}}, op::Operator)",
                "Base.:*(op::Operator, x::Vector{Float64})"):
        assert sig in src, sig
    # rga must call a user-supplied pcgafunc with exactly the reference's keywords (no library-specific ones)
    call = src[src.index("return pcgafunc("):]
    call = call[:call.index("\nend")]
    assert "ctx=" not in call and "maxiters=maxiters, delta=delta, xtol=xtol, callback=callback" in call
