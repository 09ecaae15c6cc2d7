"""The reference-side binding julia/GalerkinToolkitGPUAssemblyExt.jl cannot be executed here (no `julia` in the image), so
what CAN be checked is checked statically:
  * every `GT.<name>` it uses is a name GalerkinToolkit v0.6.3 defines (snapshot tests/golden/reference_symbols.txt, made by
    tests/golden/make_reference_symbols.py from /root/reference/src; re-derived live when the reference tree is present);
  * every `ccall` names an exported symbol of include/gtk_assembly.h with the declared arity, argument and return types;
  * the isbits mirror of gtk_form_params has the header's fields in the header's order."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "julia", "GalerkinToolkitGPUAssemblyExt.jl")
HEADER = os.path.join(ROOT, "include", "gtk_assembly.h")
SNAPSHOT = os.path.join(ROOT, "tests", "golden", "reference_symbols.txt")


def _code(path):
    """source without comments (Julia '#' comments; string literals keep their content)"""
    out = []
    for line in open(path, encoding="utf-8"):
        out.append(re.sub(r"#.*$", "", line) if '"' not in line else line.split(" # ")[0])
    return "".join(out)


def test_every_GT_identifier_exists_in_the_reference():
    used = set(re.findall(r"\bGT\.([A-Za-z_∫][\w!]*)", _code(SHIM)))
    assert len(used) > 40, used
    known = set(open(SNAPSHOT, encoding="utf-8").read().split())
    missing = sorted(used - known)
    assert not missing, f"the shim uses GT names the reference does not define: {missing}"
    if os.path.isdir("/root/reference/src"):            # the snapshot is not stale
        import importlib.util
        spec = importlib.util.spec_from_file_location("mrs", os.path.join(ROOT, "tests", "golden", "make_reference_symbols.py"))
        mrs = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mrs)
        live = mrs.symbols("/root/reference/src")
        assert used <= live, sorted(used - live)
        assert known == live


def test_extended_methods_match_reference_signatures():
    """the generic functions the shim adds methods to take, in the reference, the positional arguments the shim declares"""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present")
    asm = open("/root/reference/src/assembly.jl", encoding="utf-8").read()
    comp = open("/root/reference/src/compiler.jl", encoding="utf-8").read()
    # the model implementation the shim mirrors (COOAssembly): same positional arity
    assert re.search(r"function counter\(s::COOAssembly,::Type\{T\},dofs_i;", asm)
    assert re.search(r"function counter\(s::COOAssembly,::Type\{T\},dofs_i,dofs_j;", asm)
    assert "function do_loop(counter::COOCounter)" in asm and "if ! do_loop(counter)" in asm
    assert "function contribute!(alloc::COOMatrixAllocation,v,i,j,field_i,field_j)" in asm
    assert "function contribute!(alloc::COOVectorAllocation,v,i,field_i)" in asm
    assert "function compress(alloc::COOMatrixAllocation;reuse=Val(false))" in asm
    assert "function compress!(alloc::COOMatrixAllocation,A,cache)" in asm
    assert re.search(r"function generate_assemble_matrix\(contribution::DomainContribution,space_trial::AbstractSpace,space_test::AbstractSpace;parameters=\(\),optimize_options=nothing\)", comp)
    assert re.search(r"function generate_assemble_vector\(contribution::DomainContribution,space::AbstractSpace;parameters=\(\),optimize_options=nothing\)", comp)
    shim = _code(SHIM)
    for sig in ("GT.counter(::GPUAssembly, ::Type{T}, dofs_i;", "GT.counter(::GPUAssembly, ::Type{T}, dofs_i, dofs_j;",
                "GT.do_loop(::GPUCounter) = false", "GT.contribute!(::GPUMatrixAllocation, v, i, j, field_i, field_j)",
                "GT.contribute!(::GPUVectorAllocation, v, i, field_i)", "GT.compress(a::GPUMatrixAllocation{T,Ti}; reuse = Val(false))",
                "GT.compress!(a::GPUMatrixAllocation, A, cache)",
                "GT.generate_assemble_matrix(c::GT.DomainContribution{A,<:GPUQuadrature}, space_trial::GT.AbstractSpace,",
                "GT.generate_assemble_vector(c::GT.DomainContribution{A,<:GPUQuadrature}, space::GT.AbstractSpace;"):
        assert sig in shim, sig
    # DomainContribution's second type parameter IS the quadrature (what the dispatch relies on)
    prob = open("/root/reference/src/problems.jl", encoding="utf-8").read()
    assert re.search(r"struct DomainContribution\{A,B,C\}\s+integrand::A\s+quadrature::B\s+coefficient::C", prob)
    assert "function integrate(f,quadrature::AbstractQuadrature)" in prob


C2J = {"int32_t": "Cint", "int64_t": "Int64", "double": "Cdouble", "void": "Cvoid"}


def _header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER, encoding="utf-8").read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int32_t|int64_t|const char\*)\s+(gtk_\w+)\s*\(([^)]*)\)\s*;", text):
        params = []
        for a in [x.strip() for x in args.split(",") if x.strip() and x.strip() != "void"]:
            a = a.replace("const ", "")
            m = re.match(r"(\w+)\s*(\*{0,2})\s*\w*$", a)
            assert m, (name, a)
            base, stars = m.group(1), m.group(2)
            base = {"gtk_ctx": "void", "gtk_form_params": "FormParams"}.get(base, base)
            j = C2J.get(base, base)
            for _ in stars:
                j = f"Ptr{{{j}}}"
            params.append(j)
        protos[name] = ("Cstring" if "char" in ret else C2J[ret], params)
    return protos


def test_every_ccall_matches_the_header():
    protos = _header_prototypes()
    shim = _code(SHIM)
    calls = re.findall(r"ccall\(\(:(gtk_\w+), LIB\),\s*(\w+),\s*\(([^)]*)\)", shim)
    assert len(calls) >= 12
    for name, ret, argt in calls:
        assert name in protos, f"{name} is not declared in include/gtk_assembly.h"
        pret, pargs = protos[name]
        norm = lambda t: t.replace("Int32", "Cint").replace("Float64", "Cdouble")     # the same types in Julia
        got = [norm(a.strip()) for a in argt.split(",") if a.strip()]
        assert ret == pret, (name, ret, pret)
        assert got == pargs, (name, got, pargs)
    assert {"gtk_create", "gtk_destroy", "gtk_set_mesh", "gtk_set_space", "gtk_set_tabulation", "gtk_matrix_symbolic",
            "gtk_matrix_pattern", "gtk_matrix_numeric_device", "gtk_vector_assemble_device", "gtk_copy_nzval",
            "gtk_copy_vector", "gtk_field_set_values"} <= {c[0] for c in calls}


def test_form_params_mirror_has_the_header_layout():
    text = open(HEADER, encoding="utf-8").read()
    body = re.search(r"typedef struct gtk_form_params \{(.*?)\} gtk_form_params;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    cfields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?\w+\s*\*?", "", decl)
        cfields += [re.sub(r"\[\d+\]", "", n).strip(" *") for n in names.split(",")]
    shim = open(SHIM, encoding="utf-8").read()
    jbody = re.search(r"struct FormParams\n(.*?)\nend", shim, flags=re.S).group(1)
    jfields = [ln.split("::")[0].strip() for ln in jbody.splitlines() if "::" in ln]
    assert jfields == cfields, (jfields, cfields)


def test_block_struct_mirrors_have_the_header_layout():
    """gtk_part / gtk_block in the Julia file mirror the C structs field for field (product-space entry point)"""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER, encoding="utf-8").read(), flags=re.S)
    shim = open(SHIM, encoding="utf-8").read()
    for name in ("gtk_part", "gtk_block"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
        cfields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"^(const\s+)?\w+\s*\*?", "", decl)
            cfields += [re.sub(r"\[\d+\]", "", n).strip(" *") for n in names.split(",")]
        jbody = re.search(r"struct %s\n(.*?)\nend" % name, shim, flags=re.S).group(1)
        jfields = [ln.split("::")[0].strip() for ln in jbody.splitlines() if "::" in ln]
        assert jfields == cfields, (name, jfields, cfields)
    calls = {c for c in re.findall(r"ccall\(\(:(gtk_\w+), LIB\)", _code(SHIM))}
    assert {"gtk_set_parts", "gtk_matrix_numeric_blocks_device"} <= calls
