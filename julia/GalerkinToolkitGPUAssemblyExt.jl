# GalerkinToolkitGPUAssemblyExt.jl — reference-side binding of libgtkasm (include/gtk_assembly.h) for GalerkinToolkit v0.6.3.
#
# NOT executed in this repository's CI: the build image has no `julia`.  What IS checked (tests/test_julia_shim.py):
# every `GT.<name>` used below is defined in /root/reference/src, and every `ccall` matches the prototype in
# include/gtk_assembly.h (symbol, arity, argument and return types).
#
# The seam is the reference's own (all file:line relative to /root/reference/src):
#   * an `assembly_method` object (assembly.jl:11-25): `gpu_assembly()` implements counter / do_loop / allocate /
#     contribute! / reset! / compress / compress! exactly like `COOAssembly` does (assembly.jl:428-588); `do_loop` is false,
#     so allocate_matrix / allocate_vector skip their CPU counting loops (assembly.jl:52-54, 120-122);
#   * the loop generators `generate_assemble_matrix` / `generate_assemble_vector` (compiler.jl:1097-1136), specialised —
#     by ordinary dispatch, no method is overwritten — on contributions whose quadrature is a `GPUQuadrature`, the
#     measure `gpu_measure(Ω, degree)` returns.  They keep the calling convention `params_loop(parameters...)(alloc)`
#     that assemble_matrix / update_matrix! use (problems.jl:337-342, 352-361).
#
# User code changes in two places and nowhere else:
#     dΩ = GPU.gpu_measure(Ω, 2)                                              # was GT.measure(Ω, 2)
#     A, cache = GT.assemble_matrix(a, Float64, V, V; reuse = Val(true),
#                                   assembly_method = (; matrix = GPU.gpu_assembly(), vector = GPU.gpu_assembly()))
#     GT.update_matrix!(A, cache)                                             # numeric re-assembly on the cached pattern
# Forms the engine does not recognise raise an error (GTK_ERR_UNSUPPORTED_FORM); nothing falls back to the CPU loop.
module GalerkinToolkitGPUAssemblyExt

import GalerkinToolkit as GT
import ForwardDiff
import LinearAlgebra
using LinearAlgebra: ⋅
using SparseArrays
using StaticArrays

const LIB = get(ENV, "GTK_LIBGTKASM", "libgtkasm.so")

# ---------------------------------------------------------------------------------------------------------------------
# C ABI (include/gtk_assembly.h)
# ---------------------------------------------------------------------------------------------------------------------
const GTK_OK = Cint(0)
const GTK_FREE = Cint(1)
const GTK_DIRICHLET = Cint(2)
const FORM_LAPLACE = Cint(1)
const FORM_MASS = Cint(2)
const FORM_PLAPLACE_JACOBIAN = Cint(4)
const FORM_SOURCE_CONST = Cint(101)
const FORM_SOURCE_QP = Cint(103)
const FORM_PLAPLACE_RESIDUAL = Cint(104)

# mirrors `gtk_form_params` field for field (isbits, same layout as the C struct)
struct FormParams
    alpha::Cdouble
    lambda::Cdouble
    mu::Cdouble
    f_const::NTuple{3,Cdouble}
    f_nodal::Ptr{Cdouble}
    f_qp::Ptr{Cdouble}
    coef_nodal::Ptr{Cdouble}
    coef_qp::Ptr{Cdouble}
    accumulate::Cint
    exponent::Cdouble
end
FormParams(; alpha = 1.0, f_const = 0.0, f_qp = C_NULL, exponent = 0.0) =
    FormParams(alpha, 0.0, 0.0, (f_const, 0.0, 0.0), C_NULL, f_qp, C_NULL, C_NULL, Cint(0), exponent)

mutable struct Engine
    handle::Ptr{Cvoid}
    function Engine(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:gtk_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, out)
        rc == GTK_OK || error("gtk_create failed ($rc): libgtkasm needs a CUDA GPU, there is no CPU fallback")
        e = new(out[])
        finalizer(x -> ccall((:gtk_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), e)
        e
    end
end

function check(e::Engine, rc::Cint)
    rc == GTK_OK && return nothing
    msg = unsafe_string(ccall((:gtk_last_error, LIB), Cstring, (Ptr{Cvoid},), e.handle))
    error("libgtkasm error $rc: $msg")
end

# ---------------------------------------------------------------------------------------------------------------------
# the assembly_method object and its allocations (assembly.jl:428-588 is the model)
# ---------------------------------------------------------------------------------------------------------------------
struct GPUAssembly <: GT.AbstractType end
gpu_assembly() = GPUAssembly()

struct GPUVectorCounter{T} <: GT.AbstractType
    nrows::Int
end
struct GPUMatrixCounter{T,Ti} <: GT.AbstractType
    nrows::Int
    ncols::Int
end
const GPUCounter = Union{GPUVectorCounter,GPUMatrixCounter}

function GT.counter(::GPUAssembly, ::Type{T}, dofs_i; index_type = Int32, eltype = T, vector_type = Vector{eltype}) where T
    eltype === Float64 || error("libgtkasm assembles Float64 only")
    GPUVectorCounter{eltype}(length(dofs_i))
end
function GT.counter(::GPUAssembly, ::Type{T}, dofs_i, dofs_j; eltype = T, index_type = Int32,
                    matrix_type = SparseMatrixCSC{eltype,index_type}) where T
    (eltype === Float64 && index_type in (Int32, Int64) && matrix_type === SparseMatrixCSC{Float64,index_type}) ||
        error("libgtkasm produces SparseMatrixCSC{Float64,Int32 | Int64}; assembly_options $((; eltype, index_type, matrix_type)) are not supported")
    GPUMatrixCounter{eltype,index_type}(length(dofs_i), length(dofs_j))
end

GT.do_loop(::GPUCounter) = false          # allocate_matrix / allocate_vector skip their counting loops
GT.reset!(c::GPUCounter) = c

# the allocation owns the engine context; the generated "loop" fills it, compress reads it back
mutable struct GPUMatrixAllocation{T,Ti} <: GT.AbstractType
    counter::GPUMatrixCounter{T,Ti}
    engine::Union{Nothing,Engine}
    nnz::Int
end
mutable struct GPUVectorAllocation{T} <: GT.AbstractType
    counter::GPUVectorCounter{T}
    engine::Union{Nothing,Engine}
    n_integrals::Int                      # integrals already summed into the device vector (problems.jl:258-266)
end
const GPUAllocation = Union{GPUMatrixAllocation,GPUVectorAllocation}

GT.allocate(c::GPUMatrixCounter{T,Ti}) where {T,Ti} = GPUMatrixAllocation{T,Ti}(c, nothing, 0)
GT.allocate(c::GPUVectorCounter{T}) where T = GPUVectorAllocation{T}(c, nothing, 0)
Base.eltype(::GPUMatrixAllocation{T}) where T = T
Base.eltype(::GPUVectorAllocation{T}) where T = T

function GT.reset!(a::GPUVectorAllocation)
    a.n_integrals = 0
    a
end
GT.reset!(a::GPUMatrixAllocation) = a

# a CPU-generated loop must never fill a GPU allocation: explicit error instead of a silent fallback
GT.contribute!(::GPUVectorAllocation, v, i, field_i) =
    error("gpu_assembly() received contributions from a CPU loop: integrate over GPU.gpu_measure(Ω, degree), not GT.measure")
GT.contribute!(::GPUMatrixAllocation, v, i, j, field_i, field_j) =
    error("gpu_assembly() received contributions from a CPU loop: integrate over GPU.gpu_measure(Ω, degree), not GT.measure")

function GT.compress(a::GPUMatrixAllocation{T,Ti}; reuse = Val(false)) where {T,Ti}
    e = a.engine
    e === nothing && error("compress before any integral was assembled")
    (; nrows, ncols) = a.counter
    colptr = Vector{Ti}(undef, ncols + 1)
    rowval = Vector{Ti}(undef, a.nnz)
    nzval = Vector{Float64}(undef, a.nnz)
    GC.@preserve colptr rowval nzval begin
        if Ti === Int32
            check(e, ccall((:gtk_matrix_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), e.handle, colptr, rowval))
        else    # assembly_options = (; index_type = Int64) (assembly.jl:434-445); also lifts the 2^31 limit on nnz
            check(e, ccall((:gtk_matrix_pattern_i64, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), e.handle, colptr, rowval))
        end
        check(e, ccall((:gtk_copy_nzval, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), e.handle, nzval))
    end
    A = SparseMatrixCSC{Float64,Ti}(nrows, ncols, colptr, rowval, nzval)
    GT.val_parameter(reuse) ? (A, e) : A
end

function GT.compress!(a::GPUMatrixAllocation, A, cache)
    nz = nonzeros(A)
    GC.@preserve nz check(a.engine, ccall((:gtk_copy_nzval, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), a.engine.handle, nz))
    A
end

function GT.compress(a::GPUVectorAllocation; reuse = Val(false))
    b = Vector{Float64}(undef, a.counter.nrows)
    GC.@preserve b check(a.engine, ccall((:gtk_copy_vector, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), a.engine.handle, b))
    GT.val_parameter(reuse) ? (b, nothing) : b
end

function GT.compress!(a::GPUVectorAllocation, b, cache)
    GC.@preserve b check(a.engine, ccall((:gtk_copy_vector, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), a.engine.handle, b))
    b
end

# ---------------------------------------------------------------------------------------------------------------------
# the measure whose integrals go to the GPU
# ---------------------------------------------------------------------------------------------------------------------
struct GPUQuadrature{Q} <: GT.AbstractQuadrature
    parent::Q
end
gpu_measure(domain::GT.AbstractDomain, degree) = GPUQuadrature(GT.measure(domain, degree))

GT.domain(q::GPUQuadrature) = GT.domain(q.parent)
GT.mesh(q::GPUQuadrature) = GT.mesh(q.parent)
GT.reference_quadratures(q::GPUQuadrature) = GT.reference_quadratures(q.parent)
GT.face_reference_id(q::GPUQuadrature) = GT.face_reference_id(q.parent)
GT.coordinate_quantity(q::GPUQuadrature) = GT.coordinate_quantity(q.parent)
GT.weight_quantity(q::GPUQuadrature) = GT.weight_quantity(q.parent)

# ---------------------------------------------------------------------------------------------------------------------
# inputs of the engine from the reference's own accessors (the flat arrays of GalerkinToolkitExamples/src/poisson.jl:319-333)
# ---------------------------------------------------------------------------------------------------------------------
function upload_problem!(e::Engine, V::GT.AbstractSpace, q::GPUQuadrature)
    Ω = GT.domain(q)
    mesh = GT.mesh(Ω)
    D = GT.num_dims(mesh)
    GT.num_dims(Ω) == D || error("libgtkasm: volume integrals only on this path (boundary terms: gtk_set_manifold_dim)")
    length(GT.reference_spaces(V)) == 1 || error("libgtkasm: one reference element per mesh")
    cells = GT.faces(Ω)
    cells == 1:GT.num_faces(mesh, D) || error("libgtkasm: the domain must be the whole interior of the mesh")
    xyz = GT.node_coordinates(mesh)                        # Vector{SVector{D,Float64}} = [n_nodes][D] in memory
    cell_nodes = GT.face_nodes(mesh, D)                    # JaggedArray{Int32}: .data is [n_cells][n_lnodes]
    cell_dofs = GT.face_dofs(V)                            # JaggedArray{Int32}: negative = Dirichlet id (assembly.jl:155-157)
    n_cells = length(cell_nodes)
    n_lnodes = length(cell_nodes[1])
    n_ldofs = length(cell_dofs[1])
    n_free = length(GT.free_dofs(V))
    n_diri = length(GT.dirichlet_dofs(V))
    points = GT.coordinates(GT.reference_quadratures(q)[1])
    w = collect(Float64, GT.weights(GT.reference_quadratures(q)[1]))
    refspace = GT.reference_spaces(V)[1]
    refcell = GT.reference_spaces(mesh, Val(D))[1]
    ∇ = ForwardDiff.gradient
    # tabulator(...)(f, x) is [point, dof]; the engine wants Julia's Matrix[dof, point] memory order (accessors.jl:486-496)
    N = collect(permutedims(GT.tabulator(refspace)(GT.value, points)))
    dN = collect(permutedims(GT.tabulator(refspace)(∇, points)))
    M = collect(permutedims(GT.tabulator(refcell)(GT.value, points)))
    dM = collect(permutedims(GT.tabulator(refcell)(∇, points)))
    nd, cd = cell_nodes.data, cell_dofs.data
    GC.@preserve xyz nd cd w N dN M dM begin
        check(e, ccall((:gtk_set_mesh, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cdouble}, Int64, Cint, Ptr{Int32}),
                       e.handle, D, length(xyz), pointer(reinterpret(Float64, xyz)), n_cells, n_lnodes, nd))
        check(e, ccall((:gtk_set_space, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Int64, Int64),
                       e.handle, n_ldofs, 1, cd, n_free, n_diri))
        check(e, ccall((:gtk_set_tabulation, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                       e.handle, length(w), w, N, pointer(reinterpret(Float64, dN)), M, pointer(reinterpret(Float64, dM))))
    end
    e
end

fd_code(x) = x == GT.FREE ? GTK_FREE : GTK_DIRICHLET

# `parameters = (uh,)`: the DiscreteField goes to the engine's field slot (problems.jl:276-285, 352-361)
function upload_field!(e::Engine, uh)
    fv = collect(Float64, GT.free_values(uh))
    dv = collect(Float64, GT.dirichlet_values(uh))
    GC.@preserve fv dv check(e, ccall((:gtk_field_set_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), e.handle, fv, dv))
end

# ---------------------------------------------------------------------------------------------------------------------
# form recognition on the reference's term IR (compiler.jl:60-1028): the optimised integrand term of a contribution
# ---------------------------------------------------------------------------------------------------------------------
# Named flux functions for GT.call (an anonymous closure cannot be looked into): callable on the CPU path as well.
struct PLaplaceFlux
    q::Int
end
(f::PLaplaceFlux)(∇u) = LinearAlgebra.norm(∇u)^(f.q - 2) * ∇u
struct PLaplaceDFlux
    q::Int
end
(f::PLaplaceDFlux)(∇du, ∇u) = (f.q - 2) * LinearAlgebra.norm(∇u)^(f.q - 4) * (∇u ⋅ ∇du) * ∇u + LinearAlgebra.norm(∇u)^(f.q - 2) * ∇du

callee_value(t) = (t isa GT.CallTerm && t.callee isa GT.LeafTerm) ? t.callee.value : nothing

# all terms of a tree, depth first (GT.dependencies is the IR's generic child accessor)
function walk(f, t)
    f(t)
    for d in GT.dependencies(t)
        d isa GT.AbstractTerm && walk(f, d)
    end
end

struct FormShape
    n_grad_args::Int        # tabulated gradients of form arguments
    n_value_args::Int       # tabulated values of form arguments
    n_fields::Int           # DiscreteField evaluations
    calls::Vector{Any}      # callee values of every CallTerm
end

function shape_of(term)
    ng = nv = nf = 0
    calls = Any[]
    walk(term) do t
        if t isa GT.TabulatedTerm && t.parent isa GT.FormArgumentTerm
            # the tabulated function is the first dependency of the form argument: GT.value or ForwardDiff.gradient
            f = GT.dependencies(t.parent)[1]
            (f isa GT.LeafTerm && f.value === ForwardDiff.gradient) ? (ng += 1) : (nv += 1)
        elseif t isa GT.DiscreteFieldTerm
            nf += 1
        elseif t isa GT.CallTerm
            push!(calls, callee_value(t))
        end
    end
    FormShape(ng, nv, nf, calls)
end

unsupported(what) = error("GTK_ERR_UNSUPPORTED_FORM: $what is not one of the forms libgtkasm assembles " *
                          "(∫∇u⋅∇v, ∫u v, ∫f v, p-Laplacian residual/Jacobian through PLaplaceFlux/PLaplaceDFlux); no CPU fallback")

function recognise_bilinear(c::GT.DomainContribution)
    term = GT.optimize(GT.term(c, GT.index(Val(2))))
    s = shape_of(term)
    flux = findfirst(x -> x isa PLaplaceDFlux, s.calls)
    if flux !== nothing && s.n_grad_args == 2 && s.n_fields == 1
        return (FORM_PLAPLACE_JACOBIAN, FormParams(alpha = Float64(GT.coefficient(c)), exponent = Float64(s.calls[flux].q)))
    elseif s.n_fields == 0 && s.n_grad_args == 2 && s.n_value_args == 0 && any(x -> x === LinearAlgebra.dot, s.calls)
        return (FORM_LAPLACE, FormParams(alpha = Float64(GT.coefficient(c))))
    elseif s.n_fields == 0 && s.n_grad_args == 0 && s.n_value_args == 2
        return (FORM_MASS, FormParams(alpha = Float64(GT.coefficient(c))))
    end
    unsupported("this bilinear form")
end

function recognise_linear(c::GT.DomainContribution)
    term = GT.optimize(GT.term(c, GT.index(Val(1))))
    s = shape_of(term)
    flux = findfirst(x -> x isa PLaplaceFlux, s.calls)
    if flux !== nothing && s.n_grad_args == 1 && s.n_value_args == 1 && s.n_fields == 1
        # ∇(v,x)⋅GT.call(flux,∇(u,x)) - v(x): source f ≡ 1 (test/problems_ext_tests.jl:162)
        return (FORM_PLAPLACE_RESIDUAL, FormParams(alpha = Float64(GT.coefficient(c)), f_const = 1.0, exponent = Float64(s.calls[flux].q)))
    elseif s.n_fields == 0 && s.n_grad_args == 0 && s.n_value_args == 1 && all(x -> x === Base.:*, s.calls)
        return (FORM_SOURCE_CONST, FormParams(alpha = Float64(GT.coefficient(c)), f_const = 1.0))
    end
    unsupported("this linear form")
end

# ---------------------------------------------------------------------------------------------------------------------
# the loop generators (compiler.jl:1097-1136): same signature, same `params_loop(parameters...)(alloc)` convention
# ---------------------------------------------------------------------------------------------------------------------
function gpu_allocation(alloc, what)
    inner = alloc.allocation              # MatrixAllocation / VectorAllocation (assembly.jl:159-169)
    inner isa GPUAllocation || error("GPU.gpu_measure needs assembly_method = (; $what = GPU.gpu_assembly())")
    inner
end

function GT.generate_assemble_matrix(c::GT.DomainContribution{A,<:GPUQuadrature}, space_trial::GT.AbstractSpace,
                                     space_test::GT.AbstractSpace; parameters = (), optimize_options = nothing) where A
    space_trial === space_test || unsupported("a form with different trial and test spaces")
    form, params = recognise_bilinear(c)
    q = GT.quadrature(c)
    params_loop = (params_now...) -> function (alloc::GT.MatrixAllocation)
        a = gpu_allocation(alloc, "matrix")
        if a.engine === nothing                                   # first assembly: inputs + symbolic phase
            a.engine = upload_problem!(Engine(), space_test, q)
            nnz = Ref{Int64}(0)
            fr, fc = alloc.free_or_dirichlet
            check(a.engine, ccall((:gtk_matrix_symbolic, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}),
                                  a.engine.handle, fd_code(fr), fd_code(fc), nnz))
            a.nnz = nnz[]
        end
        isempty(params_now) || upload_field!(a.engine, params_now[1])
        p = Ref(params)
        check(a.engine, ccall((:gtk_matrix_numeric_device, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{FormParams}), a.engine.handle, form, p))
        alloc
    end
    params_loop
end

function GT.generate_assemble_vector(c::GT.DomainContribution{A,<:GPUQuadrature}, space::GT.AbstractSpace;
                                     parameters = (), optimize_options = nothing) where A
    form, params = recognise_linear(c)
    q = GT.quadrature(c)
    params_loop = (params_now...) -> function (alloc::GT.VectorAllocation)
        a = gpu_allocation(alloc, "vector")
        if a.engine === nothing
            a.engine = upload_problem!(Engine(), space, q)
            check(a.engine, ccall((:gtk_vector_symbolic, LIB), Cint, (Ptr{Cvoid}, Cint), a.engine.handle, fd_code(alloc.free_or_dirichlet)))
        end
        isempty(params_now) || upload_field!(a.engine, params_now[1])
        # a sum of integrals is ONE COO vector in the reference (problems.jl:258-266): later integrals accumulate
        p = Ref(FormParams(params.alpha, params.lambda, params.mu, params.f_const, params.f_nodal, params.f_qp,
                           params.coef_nodal, params.coef_qp, Cint(a.n_integrals > 0), params.exponent))
        check(a.engine, ccall((:gtk_vector_assemble_device, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{FormParams}), a.engine.handle, form, p))
        a.n_integrals += 1
        alloc
    end
    params_loop
end

# ---------------------------------------------------------------------------------------------------------------------
# Product spaces (V × Q): explicit-block entry point (SURVEY §8 f4).
#
# The engine assembles a CartesianProductSpace as ONE super element per cell (all fields' dofs, shifted by the field's block
# offset — exactly what MonolithicAssemblyAllocation does per push, assembly.jl:321-333, 386-416).  Transparent dispatch from
# `GT.assemble_matrix(a, T, VxQ, VxQ)` would need block recognition on the multi-field term IR (the `field == the_field`
# masks of compiler.jl:728-739); that recogniser exists in this repository only in its Python mirror (gt.py: recognise_blocks),
# so the Julia side offers the explicit form: the caller names the blocks.
#
#     VxQ = V × Q;  dΩ = GPU.gpu_measure(Ω, 4)
#     blocks = [GPU.block(1, 1, GPU.BLOCK_LAPLACE, 1.0),        # ∇v⋅∇u        (u field 1, v field 1)
#               GPU.block(2, 1, GPU.BLOCK_VALU_DIVV, -1.0),     # -div(v) p    (u field 2, v field 1)
#               GPU.block(1, 2, GPU.BLOCK_DIVU_VALV, 1.0)]      # q div(u)     (u field 1, v field 2)
#     A = GPU.assemble_product_matrix(VxQ, dΩ, blocks)          # SparseMatrixCSC{Float64,Int32}, all field blocks stored
# ---------------------------------------------------------------------------------------------------------------------
const BLOCK_MASS = Cint(1)
const BLOCK_LAPLACE = Cint(2)
const BLOCK_VALU_DIVV = Cint(3)
const BLOCK_DIVU_VALV = Cint(4)

# mirrors of the C structs (isbits, field for field)
struct gtk_part
    n_lshape::Cint
    n_comp::Cint
    side::Cint
    N::Ptr{Cdouble}
    dN::Ptr{Cdouble}
end
struct gtk_block
    part_u::Cint
    part_v::Cint
    form::Cint
    alpha::Cdouble
    c::NTuple{3,Cdouble}
end
block(field_u, field_v, form, alpha) = gtk_block(Cint(field_u - 1), Cint(field_v - 1), form, Float64(alpha), (0.0, 0.0, 0.0))

function assemble_product_matrix(VxQ::GT.AbstractSpace, q::GPUQuadrature, blocks::Vector{gtk_block})
    flds = GT.fields(VxQ)
    Ω = GT.domain(q)
    mesh = GT.mesh(Ω)
    D = GT.num_dims(mesh)
    GT.num_dims(Ω) == D || error("libgtkasm: assemble_product_matrix integrates over the interior of the mesh")
    xyz = GT.node_coordinates(mesh)
    cell_nodes = GT.face_nodes(mesh, D)
    n_cells = length(cell_nodes)
    # block offsets (assembly.jl:321-333) and the super dof table: field-major, free ids + free offset, Dirichlet ids - Dirichlet offset
    nfree = [length(GT.free_dofs(f)) for f in flds]
    ndiri = [length(GT.dirichlet_dofs(f)) for f in flds]
    off_f = cumsum(vcat(0, nfree))[1:end-1]
    off_d = cumsum(vcat(0, ndiri))[1:end-1]
    nld = [length(GT.face_dofs(f)[1]) for f in flds]
    L = sum(nld)
    super = Vector{Int32}(undef, n_cells * L)
    for cell in 1:n_cells
        k = (cell - 1) * L
        for (i, f) in enumerate(flds)
            for d in GT.face_dofs(f)[cell]
                k += 1
                super[k] = d > 0 ? Int32(d + off_f[i]) : Int32(d - off_d[i])
            end
        end
    end
    points = GT.coordinates(GT.reference_quadratures(q)[1])
    w = collect(Float64, GT.weights(GT.reference_quadratures(q)[1]))
    refcell = GT.reference_spaces(mesh, Val(D))[1]
    ∇ = ForwardDiff.gradient
    M = collect(permutedims(GT.tabulator(refcell)(GT.value, points)))
    dM = collect(permutedims(GT.tabulator(refcell)(∇, points)))
    # per field: SCALAR shape functions of its reference element (vector-valued fields: n_comp components per node)
    tabs = map(flds) do f
        reffe = GT.reference_spaces(f)[1]
        ts = GT.tensor_size(reffe)                      # :scalar or a tuple such as (D,)  (space.jl:1061, 1185-1187)
        ncomp = ts === :scalar ? 1 : prod(ts)
        scalar = ncomp == 1 ? reffe : GT.lagrange_space(GT.domain(reffe), GT.order(reffe))
        N = collect(permutedims(GT.tabulator(scalar)(GT.value, points)))
        dN = collect(permutedims(GT.tabulator(scalar)(∇, points)))
        (; ncomp, N, dN)
    end
    e = Engine()
    nd = cell_nodes.data
    GC.@preserve xyz nd super w M dM tabs begin
        check(e, ccall((:gtk_set_mesh, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cdouble}, Int64, Cint, Ptr{Int32}),
                       e.handle, D, length(xyz), pointer(reinterpret(Float64, xyz)), n_cells, length(cell_nodes[1]), nd))
        check(e, ccall((:gtk_set_space, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Int64, Int64),
                       e.handle, L, 1, super, sum(nfree), sum(ndiri)))
        parts = [gtk_part(Cint(size(t.N, 1)), Cint(t.ncomp), Cint(0), pointer(t.N), pointer(reinterpret(Float64, t.dN))) for t in tabs]
        check(e, ccall((:gtk_set_parts, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{gtk_part}, Cint, Cint, Ptr{Int32}),
                       e.handle, length(w), w, M, pointer(reinterpret(Float64, dM)), length(parts), parts, 1, 1, C_NULL))
    end
    nnz = Ref{Int64}(0)
    check(e, ccall((:gtk_matrix_symbolic, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}), e.handle, GTK_FREE, GTK_FREE, nnz))
    check(e, ccall((:gtk_matrix_numeric_blocks_device, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{gtk_block}), e.handle, length(blocks), blocks))
    n = sum(nfree)
    colptr = Vector{Int32}(undef, n + 1)
    rowval = Vector{Int32}(undef, nnz[])
    nzval = Vector{Float64}(undef, nnz[])
    check(e, ccall((:gtk_matrix_pattern, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), e.handle, colptr, rowval))
    check(e, ccall((:gtk_copy_nzval, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), e.handle, nzval))
    SparseArrays.SparseMatrixCSC(n, n, colptr, rowval, nzval)
end

end # module
