/* gtk_assembly.h — C ABI of libgtkasm, the B200-native assembly engine for
 * GalerkinToolkit.jl's hot path (the cell loop behind GT.assemble_matrix /
 * GT.assemble_vector).
 *
 * The reference has no FFI for this path (pure Julia).  Each entry point below
 * names the reference interface it replaces (file:line relative to
 * /root/reference/src) — that is what a Julia `ccall` shim binds; see
 * INTEGRATION.md for the shim.
 *
 * Conventions:
 *  - plain pointers and sizes only; every function returns an int32 status
 *    (GTK_OK = 0, < 0 = error) and never throws; gtk_last_error() gives text;
 *  - all indices on the ABI are 1-based Int32 exactly as Julia stores them
 *    (cell->node ids, cell->dof ids with NEGATIVE = Dirichlet id, colptr/rowval);
 *  - input pointers are HOST pointers borrowed for the duration of the call
 *    (the Julia side wraps the call in GC.@preserve); the engine copies to HBM;
 *  - outputs go to CALLER-allocated host arrays (two-phase: query nnz, allocate,
 *    fill) so Julia owns the SparseMatrixCSC storage; *_device variants leave
 *    results in HBM and hand out device pointers valid until the next call
 *    that re-assembles, or gtk_destroy;
 *  - a ctx is bound to one GPU and is not re-entrant; calls are synchronous
 *    unless noted; there is NO CPU fallback: without a usable GPU gtk_create fails.
 */
#ifndef GTK_ASSEMBLY_H
#define GTK_ASSEMBLY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gtk_ctx gtk_ctx;

/* status codes */
enum {
  GTK_OK = 0,
  GTK_ERR_INVALID = -1,            /* bad argument */
  GTK_ERR_CUDA = -2,               /* CUDA runtime error (text in gtk_last_error) */
  GTK_ERR_UNSUPPORTED_FORM = -3,   /* form/element not recognised: explicit error, never a CPU fallback */
  GTK_ERR_STATE = -4,              /* call order violated (e.g. numeric before symbolic) */
  GTK_ERR_TOO_LARGE = -5,          /* index range exceeded (Int32 colptr, COO index) */
  GTK_ERR_NCCL = -6
};

/* field.jl:136-142 FREE / DIRICHLET; selects rows/cols like `free_or_dirichlet`
 * of allocate_matrix (assembly.jl:75-104) */
enum { GTK_FREE = 1, GTK_DIRICHLET = 2 };

/* Recognised forms (what compiler.jl/passes.jl would pattern-match).
 * Bilinear:  be[r,c] = sum_q (alpha * integrand(u=phi_r, v=phi_c)) * dV_q
 * Linear:    be[i]   = sum_q (alpha * (f(x_q) . phi_i)) * dV_q                 */
enum {
  GTK_FORM_LAPLACE = 1,          /* ∫ ∇u·∇v            (scalar space)                      */
  GTK_FORM_MASS = 2,             /* ∫ u v  (component-wise u·v for vector spaces)           */
  GTK_FORM_ELASTICITY_ISO = 3,   /* ∫ σ(ε(u)):ε(v), σ = λ tr(ε) I + 2 μ ε (n_comp == D)     */
  GTK_FORM_PLAPLACE_JACOBIAN = 4,/* ∫ ∇v·dflux(∇du,∇u_h), dflux(∇du,∇u) = (q-2)|∇u|^(q-4)(∇u·∇du)∇u + |∇u|^(q-2)∇du: the Jacobian of
                                    the p-Laplacian about the DiscreteField parameter u_h (gtk_field_set_values), q = `exponent`
                                    (test/problems_ext_tests.jl:159-163, test/assembly_tests.jl:677-686; update_matrix! with
                                    parameters=(uh,), problems.jl:352-361; field evaluation accessors.jl:1489-1563)            */
  GTK_FORM_SOURCE_CONST = 101,   /* ∫ f·v, f constant (f_const[0..n_comp))                  */
  GTK_FORM_SOURCE_NODAL = 102,   /* ∫ f_h·v, f_h = Σ_node f_node M_node (mesh nodes)         */
  GTK_FORM_SOURCE_QP = 103,      /* ∫ f·v, f given per (cell, quadrature point)             */
  GTK_FORM_PLAPLACE_RESIDUAL = 104 /* ∫ ∇v·flux(∇u_h) − f v, flux(∇u) = |∇u|^(q-2)∇u, f = f_qp if given else f_const[0]: the residual of
                                    the p-Laplacian about u_h (update_vector! with parameters=(uh,), problems.jl:276-285)       */
};

/* Scalar integrals of the DiscreteField parameter (assemble_scalar, problems.jl:173-199: `∫(...) |> sum`):
 *   VOLUME ∫ 1     L2SQ ∫ abs2(u_h − g)     H1SQ ∫ (∇u_h − ∇g)·(∇u_h − ∇g)
 * with g (or ∇g) sampled by the host at the quadrature points through f_qp ([n_cells][n_q] values for L2SQ,
 * [n_cells][n_q][D] gradients for H1SQ; NULL = 0): the norms and manufactured-solution errors of the reference's tests
 * (test/problems_tests.jl:96-105, test/problems_ext_tests.jl:139-171). */
enum { GTK_SCALAR_VOLUME = 200, GTK_SCALAR_L2SQ = 201, GTK_SCALAR_H1SQ = 202 };

typedef struct gtk_form_params {
  double alpha;            /* scalar in front of the integral (problems.jl:29-33, 102-108) */
  double lambda, mu;       /* Lamé parameters (ELASTICITY_ISO) */
  double f_const[3];       /* SOURCE_CONST */
  const double* f_nodal;   /* SOURCE_NODAL: host [n_nodes][n_comp] */
  const double* f_qp;      /* SOURCE_QP:    host [n_cells][n_q][n_comp] */
  const double* coef_nodal;/* bilinear forms LAPLACE / MASS: scalar coefficient κ in front of the integrand, ∫ κ ∇u·∇v or ∫ κ u v,
                              as a nodal field on the mesh nodes, host [n_nodes] (κ_q = Σ_node κ_node M_node(ξ_q): a Q1/P1
                              DiscreteField passed as `parameters`, problems.jl:352-361, accessors.jl:1489-1563) … */
  const double* coef_qp;   /* … or sampled by the host at the quadrature points, host [n_cells][n_q] (AnalyticalField
                              coefficients, field.jl:17-58).  At most one of the two; NULL = no coefficient. */
  int32_t accumulate;      /* linear forms: != 0 adds this integral to the vector already on the device (gtk_set_vector or a
                              previous assembly) instead of starting from zeros — a sum of integrals such as
                              ∫_Ω f v + ∫_Γ g v is one COO vector in the reference (problems.jl:258-266) */
  double exponent;         /* PLAPLACE_*: the q of flux(∇u) = |∇u|^(q-2) ∇u */
} gtk_form_params;

/* ---- life cycle ------------------------------------------------------------ */
int32_t gtk_version(void);
/* One engine context on CUDA device `device`.  Fails (GTK_ERR_CUDA) if there is no GPU. */
int32_t gtk_create(int32_t device, gtk_ctx** out);
int32_t gtk_destroy(gtk_ctx* ctx);
const char* gtk_last_error(const gtk_ctx* ctx);
/* Launch everything on this cudaStream_t (default: the legacy default stream) so a
 * host that owns streams/events (Julia CUDA.jl, torch) can order and time the work. */
int32_t gtk_set_stream(gtk_ctx* ctx, void* cuda_stream);

/* ---- inputs ---------------------------------------------------------------- */
/* node_coordinates(mesh) :: Vector{SVector{D,Float64}} = AoS [n_nodes][D];
 * face_nodes(mesh,D) JaggedArray .data with constant n_lnodes per cell
 * (cartesian_mesh.jl:213-263; GalerkinToolkitExamples/src/poisson.jl:323-325). */
int32_t gtk_set_mesh(gtk_ctx* ctx, int32_t D, int64_t n_nodes, const double* xyz,
                     int64_t n_cells, int32_t n_lnodes, const int32_t* cell_nodes);
/* The synthetic inputs of the benchmark configurations generated in HBM instead of uploaded: GT.cartesian_mesh(domain,
 * cells) with hexahedra (cartesian_mesh.jl:213-263) restricted to the node layers [kz0, kz1] of the last direction, and
 * V = lagrange_space(interior(mesh), 1; dirichlet_boundary = boundary(mesh)) (vertex ids topology.jl:1034-1097, dof
 * numbering space.jl:327-417, 910-920).  Equivalent to gtk_set_mesh + gtk_set_space with the arrays the host would have
 * produced, bit for bit (coordinates: one multiplication and one addition per component, like the reference).
 * slab_local_numbering = 0: the reference's numbering of the WHOLE mesh (kz0 = 0, kz1 = cells[2]);
 * slab_local_numbering = 1: a z-slab of the multi-GPU partition: local node ids, free dofs numbered lexicographically among
 * the local free nodes (a node is free iff it is interior to the WHOLE box), Dirichlet ids lexicographically among the rest. */
int32_t gtk_set_cartesian_q1_problem(gtk_ctx* ctx, const double* domain6, const int64_t* cells3, int64_t kz0, int64_t kz1,
                                     int32_t slab_local_numbering, int64_t* n_free_out, int64_t* n_dirichlet_out);
/* Cells whose reference space has dimension d < D: boundary faces of a D-dimensional mesh handed over as a mesh of their
 * own (face nodes = face_nodes(mesh, D-1) of the faces of a boundary domain, domain.jl:705-753; `cell_dofs` = the space's
 * dofs on each face; tabulations on the (D-1)-dimensional reference face, gradients with d components).  dV is then
 * sqrt(det(JᵀJ)) w with the D x d Jacobian (quadrature.jl:4-6, accessors.jl:1000-1007) — how the reference integrates
 * Neumann / Robin terms ∫_Γ g v dΓ.  Call after gtk_set_mesh (which resets d = D) and before gtk_set_tabulation.
 * Forms that need physical gradients are not available for d < D (GTK_ERR_UNSUPPORTED_FORM). */
int32_t gtk_set_manifold_dim(gtk_ctx* ctx, int32_t d);
/* Cells [first, first+count) (0-based) take part in the NUMERIC assembly; all cells take part in the symbolic
 * phase.  Used by the multi-GPU partition: a rank adds the neighbour's boundary cell layer to its mesh so that the
 * pattern of its own rows is complete, but assembles only the cells it owns.  Default: all cells. */
int32_t gtk_set_active_cells(gtk_ctx* ctx, int64_t first, int64_t count);
/* Replace coordinates only (geometry update before a numeric re-assembly). */
int32_t gtk_update_coordinates(gtk_ctx* ctx, const double* xyz);
/* face_dofs(V).data (space.jl:81-87, 372-417): [n_cells][n_ldofs], 1-based, < 0 = Dirichlet id
 * (assembly.jl:155-157).  n_comp > 1: dof = (node-1)*n_comp + c (space.jl:1267-1271). */
int32_t gtk_set_space(gtk_ctx* ctx, int32_t n_ldofs, int32_t n_comp, const int32_t* cell_dofs,
                      int64_t n_free, int64_t n_dirichlet);
/* weights(quadrature) and the tabulated reference shape functions
 * (quadrature.jl:78-95; accessors.jl:486-496; poisson.jl:321-324).
 * N  [n_q][n_ldofs/n_comp]      = Julia Matrix[dof,point] memory order
 * dN [n_q][n_ldofs/n_comp][D]   M/dM: same for the n_lnodes geometry functions. */
int32_t gtk_set_tabulation(gtk_ctx* ctx, int32_t n_q, const double* w,
                           const double* N, const double* dN,
                           const double* M, const double* dM);

/* ---- matrix: allocate_matrix + compress pattern (assembly.jl:75-153, 571-575) -- */
/* Builds colptr/rowval of SparseMatrixCSC{Float64,Int32} and the device-side
 * assembly plan.  Replaces the counting loop (assembly.jl:119-153) and the
 * symbolic half of PartitionedArrays.sparse_matrix (assembly.jl:574). */
int32_t gtk_matrix_symbolic(gtk_ctx* ctx, int32_t rows_free_or_dirichlet,
                            int32_t cols_free_or_dirichlet, int64_t* nnz_out);
/* Copy the pattern out: colptr[n_cols+1], rowval[nnz], 1-based Int32. */
int32_t gtk_matrix_pattern(gtk_ctx* ctx, int32_t* colptr, int32_t* rowval);
/* The same pattern with 64-bit indices: `assembly_options = (; index_type = Int64)` of the reference
 * (assembly.jl:434-445: counter(...; index_type, matrix_type = SparseMatrixCSC{eltype,index_type})), i.e. the colptr / rowval
 * of a SparseMatrixCSC{Float64,Int64}.  Also the way out for patterns with nnz >= 2^31, which the Int32 variant refuses
 * (GTK_ERR_TOO_LARGE) — e.g. BASELINE config 5 assembled on one GPU (3.59e9 nonzeros). */
int32_t gtk_matrix_pattern_i64(gtk_ctx* ctx, int64_t* colptr, int64_t* rowval);
/* colptr entries of the selected matrix for a few columns (0-based columns, n_cols allowed; values 0-based positions in nzval):
 * what a host needs to count the stored entries of a column range — e.g. the rows a rank owns in a partitioned assembly —
 * without copying the whole colptr out. */
int32_t gtk_matrix_colptr_at(gtk_ctx* ctx, int32_t n, const int64_t* cols, int64_t* out);
/* Numeric assembly = generated loop (compiler.jl:1826-1923) + contribute!
 * (assembly.jl:189-208) + compress/compress! (assembly.jl:571-588).  Re-callable:
 * second and later calls are GT.update_matrix! (problems.jl:352-361). */
int32_t gtk_matrix_numeric(gtk_ctx* ctx, int32_t form_id, const gtk_form_params* p, double* nzval);
int32_t gtk_matrix_numeric_device(gtk_ctx* ctx, int32_t form_id, const gtk_form_params* p);

/* ---- vector: assemble_vector (problems.jl:244-274; assembly.jl:175-187, 558-569) -- */
int32_t gtk_vector_symbolic(gtk_ctx* ctx, int32_t free_or_dirichlet);
/* Upload b (host, length n_rows of the vector selection) as the device vector that an `accumulate` assembly adds to. */
int32_t gtk_set_vector(gtk_ctx* ctx, const double* b);
int32_t gtk_vector_assemble(gtk_ctx* ctx, int32_t form_id, const gtk_form_params* p, double* b);
int32_t gtk_vector_assemble_device(gtk_ctx* ctx, int32_t form_id, const gtk_form_params* p);

/* ---- matrix + vector in one pass: assemble_matrix_and_vector (problems.jl:391-404) -- */
int32_t gtk_assemble_matrix_and_vector(gtk_ctx* ctx, int32_t matrix_form, const gtk_form_params* pm,
                                       int32_t vector_form, const gtk_form_params* pv,
                                       double* nzval, double* b);
int32_t gtk_assemble_matrix_and_vector_device(gtk_ctx* ctx, int32_t matrix_form, const gtk_form_params* pm,
                                              int32_t vector_form, const gtk_form_params* pv);

/* ---- several matrices side by side + linear-problem right-hand side (problems.jl:363-387, 439-453) ------------ */
/* The reference keeps A (free x free) and Ad (free x Dirichlet) as two matrices with their own caches
 * (assemble_matrix_with_free_and_dirichlet_columns, problems.jl:363-380).  A ctx holds up to 4 matrices; `slot`
 * (0..3, default 0) selects the one gtk_matrix_symbolic / _pattern / _numeric / gtk_copy_nzval / gtk_device_pointer
 * work on.  Each keeps its pattern, plans and values, so update_matrix! on either stays a numeric-only call. */
int32_t gtk_select_matrix(gtk_ctx* ctx, int32_t slot);
/* b = beta*b + alpha*M*x with M the selected matrix, b the ctx's assembled vector (same row selection), x a host
 * vector of length n_cols: Julia's mul!(b, M, x, alpha, beta) for SparseMatrixCSC, same summation order per row
 * (increasing column) and same roundings (separate multiply and add), so `mul!(b, Ad, xd, -1, 1)` of
 * PartitionedSolvers_linear_problem (problems.jl:447) is reproduced bitwise given the same Ad, xd, b. */
int32_t gtk_matvec_add_device(gtk_ctx* ctx, double alpha, const double* x, double beta);
int32_t gtk_matvec_add(gtk_ctx* ctx, double alpha, const double* x, double beta, double* b);

/* ---- DiscreteField parameter u_h (field.jl:93-125) and nodal interpolation (space.jl:1876-1897, 2000-2060) -------- */
/* The ctx holds ONE discrete field of the current space: free values [n_free] and Dirichlet values [n_dirichlet] in HBM.
 * It is what `parameters=(uh,)` hands to the generated loops of the reference (problems.jl:465-497): the PLAPLACE_* forms
 * and gtk_scalar_assemble read it per cell (dof > 0 -> free value, dof < 0 -> Dirichlet value; accessors.jl:1489-1510).
 * gtk_field_set_values uploads host arrays (either may be NULL = leave as is; a field that was never set is zero):
 * `solution_field!(uh, x)` (problems.jl:519-526) is gtk_field_set_values(ctx, x, NULL).  The _device variant takes device
 * pointers (device-to-device copy on the ctx stream), so x and u_h need not leave HBM between Newton steps. */
int32_t gtk_field_set_values(gtk_ctx* ctx, const double* free_values, const double* dirichlet_values);
int32_t gtk_field_set_values_device(gtk_ctx* ctx, const double* d_free_values, const double* d_dirichlet_values);
int32_t gtk_field_get_values(gtk_ctx* ctx, double* free_values, double* dirichlet_values);
/* free values += a * dx  (Newton update x <- x + a dx; dx host [n_free]) */
int32_t gtk_field_axpy_free(gtk_ctx* ctx, double a, const double* dx);
/* Coordinates of the nodes the dofs sit on, as `interpolate!` sees them: node_coordinates(::LagrangeMeshSpace)
 * (space.jl:1876-1897) loops over the cells, x = Σ_lmnode tab[lnode,lmnode]·x_mnode from zero in local-mesh-node order, and
 * the LAST cell that holds a node wins.  `M_at_nodes` [n_ldofs/n_comp][n_lnodes] is that tabulator (geometry shape functions
 * at the reference nodes of the space; identity for order 1).  Results: x_free [n_free][D], x_dirichlet [n_dirichlet][D]
 * (every component dof of a node gets the node's coordinate), on the host and/or left in HBM
 * (gtk_device_pointer 6, 7).  The host evaluates g there (a Julia closure cannot cross a C ABI; with CUDA.jl it can be
 * broadcast over the device array) and hands the values to gtk_field_set_values[_device]: that IS
 * interpolate_free! / interpolate_dirichlet! (field.jl:352-373, space.jl:2000-2060: v = fun(node_x[node]) per dof). */
int32_t gtk_space_dof_coordinates(gtk_ctx* ctx, const double* M_at_nodes, double* x_free, double* x_dirichlet);
/* assemble_scalar over the current measure (problems.jl:173-190): *out = Σ_cells Σ_q integrand·dV, summed in a fixed
 * tree order (bit-reproducible).  kind: GTK_SCALAR_*. */
int32_t gtk_scalar_assemble(gtk_ctx* ctx, int32_t kind, const gtk_form_params* p, double* out);

/* ---- multi-field spaces and skeleton integrals ---------------------------------------------------- */
/* The generated loops of the reference run, per integration face, over (field of v, field of u, cell around for v, cell
 * around for u) and push one element matrix per combination through MonolithicAssemblyAllocation, which adds the field's
 * block offset to the row / column ids (compiler.jl:1826-1923, assembly.jl:321-333, 386-416; every block is pushed, also
 * the identically zero ones: block_mask defaults to all true).  Skeleton integrals (`∫(…, measure(skeleton(mesh), degree))`)
 * have two cells around every face, boundary and volume integrals one (accessors.jl:394-473).
 *
 * Engine view: the SUPER element of a face is the concatenation — field-major, then cell-around-major — of its "parts"
 * (field, side).  The host passes
 *   gtk_set_mesh          the integration faces (volume cells, or the (D-1)-faces of a skeleton / boundary with
 *                         gtk_set_manifold_dim(D-1): dV comes from the face's own geometry, accessors.jl:1000-1007),
 *   gtk_set_space         the super dof table [n_faces][L]: for every part the dofs of that field on that cell around, free ids
 *                         shifted by the field's free offset, Dirichlet ids by its Dirichlet offset (kept negative),
 *                         n_free / n_dirichlet = totals over the fields, n_comp = 1,
 *   gtk_set_parts         the quadrature weights, the geometry tabulation of the faces (as gtk_set_tabulation) and one
 *                         descriptor per part, in the order of the super dof table.
 * gtk_matrix_symbolic / gtk_vector_symbolic, patterns, slots, free / Dirichlet selections then work unchanged. */
#define GTK_MAX_PARTS 8
typedef struct gtk_part {
  int32_t n_lshape;   /* scalar shape functions of the field on one cell                                             */
  int32_t n_comp;     /* components; the part holds n_lshape * n_comp consecutive local dofs, dof = shape * n_comp + c */
  int32_t side;       /* which cell around the face (0-based; volume integrals: 0)                                   */
  const double* N;    /* host [n_var][n_q][n_lshape]: the cell's shape functions at the face's quadrature points, one table
                         per (local face, permutation) variant (accessors.jl:498-522, reference_map :1914-1943); volume: n_var = 1 */
  const double* dN;   /* host [n_var][n_q][n_lshape][D] reference gradients, or NULL (values only); on skeleton faces they are used
                         by GTK_BLOCK_IP together with gtk_set_skeleton_cells */
} gtk_part;
/* face_var: host [n_faces][n_sides], 0-based variant of the tabulation for the cell around on each side (NULL if n_var ==
 * n_sides == 1).  Replaces gtk_set_tabulation for this mesh / space; a later gtk_set_mesh / gtk_set_space drops the parts. */
int32_t gtk_set_parts(gtk_ctx* ctx, int32_t n_q, const double* w, const double* M, const double* dM, int32_t n_parts,
                      const gtk_part* parts, int32_t n_sides, int32_t n_var, const int32_t* face_var);
/* Block integrands.  Row index = shape function of u (part_u), column = shape function of v (part_v), like the single-field
 * forms (be[r,c] = Σ_q (alpha · integrand(u = φ_r, v = φ_c)) · dV_q).  alpha carries the integral's scalar, the sign of the
 * term and the side weights of jump (-1 on side 0, +1 on side 1: v[2](p) - v[1](p)) or mean (1/2) — exact factors. */
enum {
  GTK_BLOCK_ZERO = 0,       /* nothing recognised in this block: zeros are pushed (they are stored, as in the reference)   */
  GTK_BLOCK_MASS = 1,       /* u(x) * v(x)  (component-wise for vector parts; the two parts may be different fields / sides) */
  GTK_BLOCK_LAPLACE = 2,    /* ∇(v,x) ⋅ ∇(u,x)  (component-wise: the Frobenius product of the Jacobians for vector parts)   */
  GTK_BLOCK_VALU_DIVV = 3,  /* u(x) * div(v,x): u a scalar part (pressure), v a vector part with n_comp == D                */
  GTK_BLOCK_DIVU_VALV = 4,  /* v(x) * div(u,x): v a scalar part, u a vector part                                           */
  GTK_BLOCK_IP = 5,         /* skeleton faces, scalar parts, u on the cell around s_u, v on s_v (needs gtk_set_skeleton_cells):
                               c[0] ((1/h) v n_sv)⋅(u n_su) + c[1] (v n_sv)⋅∇u + c[2] ∇v⋅(u n_su), n = unit normals of the two cells
                               (accessors.jl:1009-1035), h = diameter of the face (field.jl:488-492, accessors.jl:907-921): the
                               interior-penalty terms (γ/h) jump(v,n)⋅jump(u,n) - jump(v,n)⋅mean(∇u) - mean(∇v)⋅jump(u,n) of
                               test/assembly_tests.jl:329-340 are c = (γ, -1/2, -1/2) on all four (side, side) blocks            */
  GTK_BLOCK_IP_NOH = 6      /* the same with c[0] NOT divided by h: c[0] (v n_sv)⋅(u n_su) + c[1] (v n_sv)⋅∇u + c[2] ∇v⋅(u n_su) — on a boundary
                               face (n⋅n = 1) the Nitsche terms v u - v n⋅∇u - n⋅∇v u of test/issue_224.jl:73-76 are c = (1, -1, -1)      */
};
typedef struct gtk_block { int32_t part_u, part_v, form; double alpha; double c[3]; } gtk_block;
/* Cells around the faces of a skeleton measure, for blocks with gradients / normals there (call after gtk_set_parts):
 * cell_nodes [n_cells][n_lnodes] = face_nodes(mesh, D).data, side_cells [n_faces][n_sides] (1-based; boundary faces have one cell around), dM_cell [n_var][n_q][n_lnodes][D]
 * = the cell's geometry gradients at the face points mapped into the cell per (local face, permutation) variant, ref_normals
 * [n_var][D] = normals(mesh(domain(refface)))[ldface] of the variant's local face (domain.jl:226, 258).  The parts' dN tables are
 * then [n_var][n_q][n_lshape][D] at the same mapped points. */
int32_t gtk_set_skeleton_cells(gtk_ctx* ctx, int64_t n_cells, int32_t n_lnodes, const int32_t* cell_nodes, const int32_t* side_cells,
                               const double* dM_cell, const double* ref_normals);
/* Numeric assembly of Σ blocks on the pattern of gtk_matrix_symbolic (blocks not listed are GTK_BLOCK_ZERO); re-callable
 * like gtk_matrix_numeric (update_matrix!).  E.g. Stokes a((u,p),(v,q)) = ∫ ∇v⋅∇u - div(v) p + q div(u)
 * (docs/src/src_jl/example_stokes.jl) is {(u,v,LAPLACE,1), (p,v,VALU_DIVV,-1), (u,q,DIVU_VALV,1)}; the reference's test
 * ∫_Λ jump(u)·jump(v) (test/assembly_tests.jl:397-401) is MASS on the four (side, side) blocks with alpha = ±1. */
int32_t gtk_matrix_numeric_blocks(gtk_ctx* ctx, int32_t n_blocks, const gtk_block* blocks, double* nzval);
int32_t gtk_matrix_numeric_blocks_device(gtk_ctx* ctx, int32_t n_blocks, const gtk_block* blocks);
/* Linear forms: per part ∫ alpha (f · v) with f constant — ∫_Λ jump(v) (test/assembly_tests.jl:420-424) is alpha = -1 / +1
 * on the two sides, f = 1.  accumulate as in gtk_form_params. */
typedef struct gtk_vblock { int32_t part; double alpha; double f_const[3]; double c[3]; } gtk_vblock;
int32_t gtk_vector_assemble_blocks(gtk_ctx* ctx, int32_t n, const gtk_vblock* vblocks, int32_t accumulate, double* b);
int32_t gtk_vector_assemble_blocks_device(gtk_ctx* ctx, int32_t n, const gtk_vblock* vblocks, int32_t accumulate);
/* Linear forms with data g sampled by the host at the faces' quadrature points (g_qp: host [n_faces][n_q]; scalar parts):
 *   per part  ∫ alpha g (c[0] v + (c[1]/h) v + c[2] n⋅∇v)     (f_const unused)
 * — the Nitsche right-hand side (γ/h) v g - n⋅∇v g of docs/src/src_jl/example_hello_world_dg.jl:83-87 is c = (0, γ, -1) on the
 * faces of the Dirichlet boundary (one cell around: gtk_set_skeleton_cells with n_sides = 1), a volume source v f is c = (1, 0, 0). */
int32_t gtk_vector_assemble_blocks_data(gtk_ctx* ctx, int32_t n, const gtk_vblock* vblocks, const double* g_qp, int32_t accumulate, double* b);
int32_t gtk_vector_assemble_blocks_data_device(gtk_ctx* ctx, int32_t n, const gtk_vblock* vblocks, const double* g_qp, int32_t accumulate);

/* ---- sums of integrals over different domains in ONE matrix --------------------------------------------------------- */
/* a(u,v) = ∫_Ω … dΩ + ∫_Γ … dΓ + ∫_Λ … dΛ: the reference lets every contribution of the form push into the same COO
 * allocation and compresses once (problems.jl:319-350), so the matrix has the UNION pattern and, per stored entry, the sum of
 * all triplets.  The engine assembles one integral per context (each has its own integration faces: cells, boundary faces,
 * interior faces); `ctx` (a context of its own, on the same GPU, no mesh needed) then holds the merged matrix:
 *   gtk_matrix_sum_symbolic  union colptr / rowval of the sources' patterns (same row / column selection and sizes) and, per
 *                            source, the position of each of its nonzeros in the union; afterwards gtk_matrix_pattern[_i64],
 *                            gtk_copy_nzval, gtk_device_pointer, gtk_select_matrix / gtk_matvec_add work on `ctx` as usual;
 *   gtk_matrix_sum_numeric   nzval = Σ_k nzval_k in the order of `src` (call after the sources' numeric assemblies; re-callable:
 *                            update_matrix!).  No float atomics; per entry the sum is (Σ integral 1) + (Σ integral 2) + …,
 *                            the reference adds the later integrals' triplets one by one instead (O(1e-16) relative apart). */
int32_t gtk_matrix_sum_symbolic(gtk_ctx* ctx, int32_t n, gtk_ctx** src, int64_t* nnz_out);
int32_t gtk_matrix_sum_numeric(gtk_ctx* ctx, int32_t n, gtk_ctx** src, double* nzval);
int32_t gtk_matrix_sum_numeric_device(gtk_ctx* ctx, int32_t n, gtk_ctx** src);

/* ---- device-resident results -------------------------------------------------- */
/* which: 0 nzval (double[nnz]) 1 b (double[n_rows]) 2 colptr (int64[n_cols+1], 0-based)
 *        3 rowval (int32[nnz], 1-based) 4 field free values (double[n_free]) 5 field Dirichlet values (double[n_dirichlet])
 *        6 / 7 coordinates of the free / Dirichlet dofs (double[n][D], after gtk_space_dof_coordinates)
 *        8 node coordinates (double[n_nodes][D])  9 cell_nodes (int32[n_cells][n_lnodes])  10 cell_dofs (int32[n_cells][n_ldofs]) */
int32_t gtk_device_pointer(gtk_ctx* ctx, int32_t which, void** dptr, int64_t* count);
/* Host copy of the first `bytes` bytes of the array gtk_device_pointer(which) names (diagnostics, tests). */
int32_t gtk_copy_device_array(gtk_ctx* ctx, int32_t which, void* host, int64_t bytes);
int32_t gtk_copy_nzval(gtk_ctx* ctx, double* nzval);
int32_t gtk_copy_vector(gtk_ctx* ctx, double* b);

/* ---- introspection (bench / tests) --------------------------------------------- */
/* key: 0 kernels launched by the last numeric call   1 total kernels launched
 *      2 device bytes held   3 nnz   4 n_coo (valid triplets)   5 fast-path id of last numeric call */
int64_t gtk_info(const gtk_ctx* ctx, int32_t key);

/* Per-kernel device timing of the LAST numeric call (CUDA events recorded on the ctx stream around
 * every kernel launch while profiling is on).  gtk_profile_get synchronises the stream. */
int32_t gtk_set_profiling(gtk_ctx* ctx, int32_t on);
int32_t gtk_profile_count(const gtk_ctx* ctx);
int32_t gtk_profile_get(gtk_ctx* ctx, int32_t i, char* name64, double* milliseconds);

/* ---- multi-GPU: ghost-row summation over NCCL ------------------- */
/* 128-byte ncclUniqueId produced on rank 0; the host broadcasts it by its own means. */
int32_t gtk_comm_unique_id(void* id128);
int32_t gtk_comm_init(gtk_ctx* ctx, int32_t rank, int32_t n_ranks, const void* id128);
/* Exchange plan with one peer rank, computed by the host (PartitionedArrays-style local_to_owner /
 * local_to_global bookkeeping; see galerkintoolkit.jl_b200/partition.py):
 *   send_nz[n_send_nz]  0-based positions in nzval of the ghost-row entries this rank sends to `peer`
 *   send_rows[n_send_b] 0-based rows of b sent to `peer`
 *   recv_nz / recv_rows where the values received from `peer` are ADDED (each position at most once per peer;
 *                       the peer's send order must match this receive order).
 * Must be called after gtk_matrix_symbolic; replaces any previous plan for that peer. */
int32_t gtk_comm_set_exchange(gtk_ctx* ctx, int32_t peer, int64_t n_send_nz, const int64_t* send_nz,
                              int64_t n_send_b, const int32_t* send_rows, int64_t n_recv_nz, const int64_t* recv_nz,
                              int64_t n_recv_b, const int32_t* recv_rows);
/* The same plan built ON THE DEVICE for a block row partition (PartitionedArrays' variable_partition data model,
 * docs/src/src_jl/manual_mesh_partitioning.jl:14-35): local free row i (0-based) has global id gid0 + i, rank p owns the
 * global ids [own_start[p], own_start[p+1]) (own_start: host, n_ranks + 1 entries).  Rows of the local pattern that another
 * rank owns and that the ACTIVE cells contribute to are ghost rows: their stored entries (CSC order) and b rows are sent to
 * the owner, which locates them in its own pattern.  Collective over the communicator of gtk_comm_init (counts:
 * ncclAllGather; (row, column) keys: ncclSend/ncclRecv); replaces every previous plan.  Needs the free x free pattern of
 * slot 0.  Errors if an announced entry is missing from the owner's pattern or lies in a row the receiver does not own. */
int32_t gtk_comm_build_exchange(gtk_ctx* ctx, int64_t gid0, const int64_t* own_start);
/* Collective: swaps the CUDA IPC handles of all receive blocks over NCCL, imports them, and agrees on the transport (peer
 * memory if EVERY rank mapped every block, else NCCL for everybody).  Call on every rank after the plan is set. */
int32_t gtk_comm_connect_peer_memory(gtk_ctx* ctx);
/* Exchange + add ghost-row nzval and b contributions after a numeric call. */
int32_t gtk_comm_sum_ghost_rows(gtk_ctx* ctx);
/* gtk_assemble_matrix_and_vector_device + gtk_comm_sum_ghost_rows in one call, with the exchange overlapped with the
 * assembly: the part of the sweep that produces the values a peer waits for runs first, pack + ncclSend/ncclRecv proceed
 * on a private side stream while the rest of the sweep runs, the received partial sums are added at the end.  Bitwise the
 * same result as the two separate calls (which is also what it does when the structured sweep kernels do not apply). */
int32_t gtk_assemble_and_sum_ghost_rows_device(gtk_ctx* ctx, int32_t matrix_form, const gtk_form_params* pm,
                                               int32_t vector_form, const gtk_form_params* pv);
/* Peer-memory transport (NVLink / NVSwitch boxes): after gtk_comm_set_exchange, every rank exports for each peer the
 * 64-byte CUDA IPC handle of the block it receives that peer's values in; the host carries it to the peer (like the NCCL
 * unique id), which imports it.  Once all peers of a rank are imported, the exchange is one kernel per peer that gathers
 * the ghost entries and stores them straight into the owner's buffer over NVLink, plus system-scope sequence flags; NCCL
 * (gtk_comm_init) is then only the fallback (GTK_DISABLE_P2P=1).  Results are bitwise the same on both transports. */
int32_t gtk_comm_p2p_export(gtk_ctx* ctx, int32_t peer, void* handle64);
int32_t gtk_comm_p2p_import(gtk_ctx* ctx, int32_t peer, const void* handle64);
/* key: 0 ghost nz entries sent per exchange  1 ghost nz entries received  2 bytes moved per exchange
 *      3 transport of the next exchange (1 peer memory, 0 NCCL) */
int64_t gtk_comm_ghost_info(const gtk_ctx* ctx, int32_t key);

#ifdef __cplusplus
}
#endif
#endif /* GTK_ASSEMBLY_H */
