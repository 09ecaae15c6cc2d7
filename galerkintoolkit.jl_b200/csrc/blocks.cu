// multi-field spaces (CartesianProductSpace) and skeleton integrals.
//
// The reference's generated matrix loop (compiler.jl:1826-1923) runs, per integration face and point, over
//   field_1 (argument v) ⊃ field_2 (argument u) ⊃ face_around_1 ⊃ face_around_2 ⊃ dof_1 ⊃ dof_2
// into one element matrix per (field pair, pair of cells around the face), and pushes every one of them through
// MonolithicAssemblyAllocation (assembly.jl:386-416: row / column offsets per field, every block — also the identically
// zero ones — because block_mask defaults to all true, assembly.jl:321-333).
//
// Engine view: a face carries a SUPER element whose local dofs are the concatenation, field-major then cell-around-major,
// of the "parts" (field, side); the host hands over the super dof table through gtk_set_space (free ids offset by the
// field's block offset, Dirichlet ids likewise and negative).  For one (row, column) key all duplicates of a face come
// from different (side_u, side_v) pairs, and the column-major order of the super matrix (side_v outer, side_u inner)
// is the reference's push order for them; different fields never share a key.  So the sort-based symbolic phase and the
// fixed-order reduction of the single-field path (symbolic.cu, k_reduce_nz / k_reduce_rows) are used unchanged.
//
//   k_elem_blocks<D,d>  phase A: one thread per (face, point): J and dV of the integration face (its own geometry,
//                                accessors.jl:1000-1007) and, for volume cells, the physical gradients of every part;
//                       phase B: one thread per super-matrix entry: the block's integrand summed over the points in
//                                the reference's order, α and the side weights (jump: -1 / +1, mean: 1/2) folded into the
//                                block's scalar (exact: they are powers of two or signs).
//   k_elem_vblocks<D,d> the same for linear forms.
//
// On skeleton / boundary faces the shape functions of the cells around are tabulated at the face's quadrature points
// mapped into the cell for every (local face, permutation) variant (accessors.jl:498-522, reference_map :1914-1943); the
// host passes those tables and the variant of every (face, side).
#include <algorithm>
#include <string>
#include "gtk_internal.h"
#include "elem_math.cuh"

int32_t gtk_reduce_nz_launch(gtk_ctx* ctx);
int32_t gtk_reduce_rows_launch(gtk_ctx* ctx, int accumulate);

namespace {

constexpr int BP_MAX = GTK_MAX_PARTS;

struct BlockArgs {
  const double* xyz;
  const int32_t* cell_nodes;
  int64_t n_cells;
  int nln, nq, n_parts, L, n_sides, nls_total, need_grad;
  const double *w, *dM;
  int p_nls[BP_MAX], p_ncomp[BP_MAX], p_off[BP_MAX], p_side[BP_MAX], p_goff[BP_MAX];
  const double* p_N[BP_MAX];
  const double* p_dN[BP_MAX];
  const int32_t* face_var;               // [n_cells][n_sides] 0-based tabulation variant, or null (variant 0)
  int b_form[BP_MAX][BP_MAX];            // [part of u = row][part of v = column]
  double b_alpha[BP_MAX][BP_MAX];
  double v_alpha[BP_MAX], v_f[BP_MAX][3];   // linear forms: per part
  double v_c[BP_MAX][3];                 // linear forms with data g(x_q): k0 v g + (k1/h) v g + k2 (n⋅∇v) g
  const double* v_g;                     // [n_faces][nq] data sampled at the faces' quadrature points, or null
  // skeleton faces with gradient / normal terms (GTK_BLOCK_IP): geometry of the cells around
  const int32_t* cellD_nodes;            // [n_Dcells][nlnD] 1-based
  const int32_t* side_cells;             // [n_faces][n_sides] 1-based cell around
  const double* dMc;                     // [n_var][nq][nlnD][D] geometry gradients of the cell at the mapped face points
  const double* nref;                    // [n_var][D] reference normal of the variant's local face
  int nlnD, skel;
  double b_c[BP_MAX][BP_MAX][3];         // GTK_BLOCK_IP coefficients
  double* out;
  int cb;
  int64_t act0, act1;
};

__device__ __forceinline__ int part_of(const BlockArgs& a, int i) {
  int p = 0;
  while (p + 1 < a.n_parts && i >= a.p_off[p + 1]) ++p;
  return p;
}

__device__ __forceinline__ int variant_of(const BlockArgs& a, int64_t cell, int part) {
  return a.face_var ? a.face_var[cell * a.n_sides + a.p_side[part]] : 0;
}

// phase A shared by both kernels
template <int D, int d>
__device__ __forceinline__ void blocks_phase_a(const BlockArgs& a, int64_t cell0, int ncb, double* G, double* dV, double* Nrm = nullptr,
                                               double* hF = nullptr) {
  const int nq = a.nq;
  if constexpr (D != d) if (a.skel && hF) {
    // diameter(::MeshFace) (accessors.jl:907-921): the largest distance between two nodes of the face
    for (int cl = threadIdx.x; cl < ncb; cl += blockDim.x) {
      const int32_t* nd = a.cell_nodes + (cell0 + cl) * a.nln;
      double diam = 0.0;
      for (int i = 0; i < a.nln; ++i)
        for (int j = 0; j < a.nln; ++j) {
          double s2 = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) { const double dx = a.xyz[(size_t)(nd[i] - 1) * D + k] - a.xyz[(size_t)(nd[j] - 1) * D + k]; s2 += dx * dx; }
          diam = fmax(diam, sqrt(s2));
        }
      hF[cl] = diam;
    }
  }
  for (int t = threadIdx.x; t < ncb * nq; t += blockDim.x) {
    const int cl = t / nq, q = t - cl * nq;
    const int64_t cell = cell0 + cl;
    double J[D][d];
    gtkmath::jacobian_from<D, d>(a.xyz, a.cell_nodes + cell * a.nln, a.nln, a.dM + (size_t)q * a.nln * d, J);
    dV[t] = gtkmath::change_of_measure<D, d>(J) * a.w[q];
    if constexpr (D != d) if (a.skel && Nrm) {
      // the cells around: Jacobian of the CELL's geometry at the mapped face point (unit_normal, accessors.jl:1009-1035;
      // shape_functions(gradient, …) :1312-1333), physical gradients of the side's parts, unit normal J^-T n_ref / |…|
      for (int side = 0; side < a.n_sides; ++side) {
        const int64_t c = a.side_cells[cell * a.n_sides + side] - 1;
        const int var = a.face_var[cell * a.n_sides + side];
        double Jc[D][D];
        gtkmath::jacobian_from<D, D>(a.xyz, a.cellD_nodes + c * a.nlnD, a.nlnD, a.dMc + ((size_t)var * nq + q) * a.nlnD * D, Jc);
        double JT[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) JT[i][j] = Jc[j][i];
        const double dt = gtkmath::det_mat<D>(JT);
        for (int p = 0; p < a.n_parts; ++p) {
          if (a.p_side[p] != side || !a.p_dN[p]) continue;
          const int nls = a.p_nls[p];
          const double* dNq = a.p_dN[p] + ((size_t)var * nq + q) * nls * D;
          double* g = G + ((size_t)t * a.nls_total + a.p_goff[p]) * D;
          for (int s = 0; s < nls; ++s) gtkmath::solve_JT<D>(Jc, dt, dNq + s * D, g + s * D);
        }
        double v[D];
        gtkmath::solve_JT<D>(Jc, dt, a.nref + (size_t)var * D, v);
        double m2 = v[0] * v[0];
#pragma unroll
        for (int k = 1; k < D; ++k) m2 += v[k] * v[k];
        const double m = sqrt(m2);
        double* nr = Nrm + ((size_t)t * a.n_sides + side) * D;
#pragma unroll
        for (int k = 0; k < D; ++k) nr[k] = m < 2.220446049250313e-16 ? 0.0 : v[k] / m;      // map_unit_normal (accessors.jl:1026-1035)
      }
    }
    if constexpr (D == d) if (a.need_grad) {
      double JT[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) JT[i][j] = J[j][i];
      const double dt = gtkmath::det_mat<D>(JT);
      for (int p = 0; p < a.n_parts; ++p) {
        if (!a.p_dN[p]) continue;
        const int nls = a.p_nls[p];
        const double* dNq = a.p_dN[p] + ((size_t)variant_of(a, cell, p) * nq + q) * nls * D;
        double* g = G + ((size_t)t * a.nls_total + a.p_goff[p]) * D;
        for (int s = 0; s < nls; ++s) gtkmath::solve_JT<D>(J, dt, dNq + s * D, g + s * D);
      }
    }
  }
}

template <int D, int d>
__global__ void __launch_bounds__(128) k_elem_blocks(BlockArgs a) {
  extern __shared__ double smem[];
  const int nq = a.nq, L = a.L;
  double* dV = smem;                                   // [cb][nq]
  double* G = dV + (size_t)a.cb * nq;                  // [cb][nq][nls_total][D]   (blocks with gradients only)
  double* Nrm = G + (a.need_grad ? (size_t)a.cb * nq * a.nls_total * D : 0);   // [cb][nq][n_sides][D] unit normals (skeleton, IP blocks)
  double* hF = Nrm + (a.skel ? (size_t)a.cb * nq * a.n_sides * D : 0);         // [cb] face diameters
  const int64_t cell0 = (int64_t)blockIdx.x * a.cb;
  const int ncb = (int)min((int64_t)a.cb, a.n_cells - cell0);
  blocks_phase_a<D, d>(a, cell0, ncb, G, dV, Nrm, hF);
  __syncthreads();
  const int L2 = L * L;
  for (int t = threadIdx.x; t < ncb * L2; t += blockDim.x) {
    const int cl = t / L2, rem = t - cl * L2;
    const int c = rem / L, r = rem - c * L;             // column: shape function of v, row: shape function of u
    const int pu = part_of(a, r), pv = part_of(a, c);
    const int form = a.b_form[pu][pv];
    const int64_t cell = cell0 + cl;
    double acc = 0.0;
    const int ru = r - a.p_off[pu], cv = c - a.p_off[pv];
    const int ncu = a.p_ncomp[pu], ncv = a.p_ncomp[pv];
    const int ra = ru / ncu, ri = ru - ra * ncu;
    const int ca = cv / ncv, cj = cv - ca * ncv;
    // component-wise forms couple equal components only: the other entries of the block are (stored) zeros
    const bool off_component = (form == GTK_BLOCK_MASS || form == GTK_BLOCK_LAPLACE) && ri != cj;
    if (form != GTK_BLOCK_ZERO && !off_component && cell >= a.act0 && cell < a.act1) {
      const double alpha = a.b_alpha[pu][pv];
      const int nlu = a.p_nls[pu], nlv = a.p_nls[pv];
      const double* Nu = a.p_N[pu] + (size_t)variant_of(a, cell, pu) * nq * nlu + ra;
      const double* Nv = a.p_N[pv] + (size_t)variant_of(a, cell, pv) * nq * nlv + ca;
      // one loop per form (the form is fixed per entry): Σ_q (alpha * integrand) * dV in the reference's order
      const double* dv = dV + cl * nq;
      const size_t gstride = (size_t)a.nls_total * D;
      const double* gu = G + (size_t)cl * nq * gstride + (size_t)(a.p_goff[pu] + ra) * D;
      const double* gv = G + (size_t)cl * nq * gstride + (size_t)(a.p_goff[pv] + ca) * D;
      if (form == GTK_BLOCK_MASS) {
        for (int q = 0; q < nq; ++q) acc += (alpha * (Nu[q * nlu] * Nv[q * nlv])) * dv[q];
      } else if (form == GTK_BLOCK_IP || form == GTK_BLOCK_IP_NOH) {
        // interior-penalty terms on a skeleton face, u on side su, v on side sv (test/assembly_tests.jl:329-340):
        //   c0 ((1/h) v n_sv)⋅(u n_su) + c1 (v n_sv)⋅∇u + c2 ∇v⋅(u n_su)
        if constexpr (D != d) {
          const int su = a.p_side[pu], sv = a.p_side[pv];
          const double c0 = form == GTK_BLOCK_IP ? a.b_c[pu][pv][0] / hF[cl] : a.b_c[pu][pv][0], c1 = a.b_c[pu][pv][1], c2 = a.b_c[pu][pv][2];
          for (int q = 0; q < nq; ++q, gu += gstride, gv += gstride) {
            const double* nu = Nrm + ((size_t)(cl * nq + q) * a.n_sides + su) * D;
            const double* nv = Nrm + ((size_t)(cl * nq + q) * a.n_sides + sv) * D;
            const double fu = Nu[q * nlu], fv = Nv[q * nlv];
            double t0 = (c0 * (fv * nv[0])) * (fu * nu[0]), t1 = (fv * nv[0]) * gu[0], t2 = gv[0] * (fu * nu[0]);
#pragma unroll
            for (int k = 1; k < D; ++k) {
              t0 += (c0 * (fv * nv[k])) * (fu * nu[k]);
              t1 += (fv * nv[k]) * gu[k];
              t2 += gv[k] * (fu * nu[k]);
            }
            acc += (alpha * ((t0 + c1 * t1) + c2 * t2)) * dv[q];
          }
        }
      } else if constexpr (D == d) {
        if (form == GTK_BLOCK_LAPLACE) {
          for (int q = 0; q < nq; ++q, gu += gstride, gv += gstride) {
            double dt = gv[0] * gu[0];
#pragma unroll
            for (int k = 1; k < D; ++k) dt += gv[k] * gu[k];
            acc += (alpha * dt) * dv[q];
          }
        } else if (form == GTK_BLOCK_VALU_DIVV) {       // u(x) * div(v)(x): u scalar part, v vector part
          for (int q = 0; q < nq; ++q, gv += gstride) acc += (alpha * (gv[cj] * Nu[q * nlu])) * dv[q];
        } else if (form == GTK_BLOCK_DIVU_VALV) {       // v(x) * div(u)(x): v scalar part, u vector part
          for (int q = 0; q < nq; ++q, gu += gstride) acc += (alpha * (Nv[q * nlv] * gu[ri])) * dv[q];
        }
      }
    }
    a.out[cell * (int64_t)L2 + rem] = acc;
  }
}

template <int D, int d>
__global__ void __launch_bounds__(128) k_elem_vblocks(BlockArgs a) {
  extern __shared__ double smem[];
  const int nq = a.nq, L = a.L;
  double* dV = smem;                                   // [cb][nq]
  double* G = dV + (size_t)a.cb * nq;                  // [cb][nq][nls_total][D]   (faces with gradient / normal terms only)
  double* Nrm = G + (a.need_grad ? (size_t)a.cb * nq * a.nls_total * D : 0);
  double* hF = Nrm + (a.skel ? (size_t)a.cb * nq * a.n_sides * D : 0);
  const int64_t cell0 = (int64_t)blockIdx.x * a.cb;
  const int ncb = (int)min((int64_t)a.cb, a.n_cells - cell0);
  blocks_phase_a<D, d>(a, cell0, ncb, a.skel ? G : nullptr, dV, a.skel ? Nrm : nullptr, a.skel ? hF : nullptr);
  __syncthreads();
  for (int t = threadIdx.x; t < ncb * L; t += blockDim.x) {
    const int cl = t / L, i = t - cl * L;
    const int p = part_of(a, i);
    const int64_t cell = cell0 + cl;
    double acc = 0.0;
    if (a.v_alpha[p] != 0.0 && cell >= a.act0 && cell < a.act1) {
      const int il = i - a.p_off[p], nc = a.p_ncomp[p], nls = a.p_nls[p];
      const int ia = il / nc, ic = il - ia * nc;
      const double* Np = a.p_N[p] + (size_t)variant_of(a, cell, p) * nq * nls + ia;
      const double alpha = a.v_alpha[p];
      if (!a.v_g) {
        const double f = a.v_f[p][ic];
        for (int q = 0; q < nq; ++q) acc += (alpha * (f * Np[q * nls])) * dV[cl * nq + q];
      } else {
        // data terms (Nitsche right-hand side, docs/src/src_jl/example_hello_world_dg.jl:83-87): g sampled at the face points
        const double k0 = a.v_c[p][0], k1 = a.skel ? a.v_c[p][1] / hF[cl] : 0.0, k2 = a.v_c[p][2];
        const double* gq = a.v_g + cell * nq;
        for (int q = 0; q < nq; ++q) {
          const double fv = Np[q * nls];
          double t = (k0 * fv) * gq[q] + (k1 * fv) * gq[q];
          if constexpr (D != d) if (a.skel && k2 != 0.0) {
            const double* gv = G + ((size_t)(cl * nq + q) * a.nls_total + a.p_goff[p] + ia) * D;
            const double* nv = Nrm + ((size_t)(cl * nq + q) * a.n_sides + a.p_side[p]) * D;
            double dn = nv[0] * gv[0];
#pragma unroll
            for (int k = 1; k < D; ++k) dn += nv[k] * gv[k];
            t += k2 * (dn * gq[q]);
          }
          acc += (alpha * t) * dV[cl * nq + q];
        }
      }
    }
    a.out[cell * (int64_t)L + i] = acc;
  }
}

struct PartsState {
  int n_parts = 0, n_sides = 1, n_var = 1, L = 0, nls_total = 0;
  int nls[BP_MAX], ncomp[BP_MAX], side[BP_MAX], off[BP_MAX], goff[BP_MAX];
  bool has_dN[BP_MAX];
  double* tab = nullptr;  size_t tab_n = 0;          // all N then all dN tables, one allocation
  size_t N_at[BP_MAX], dN_at[BP_MAX];
  int32_t* face_var = nullptr; size_t face_var_n = 0;
  int64_t n_cells = 0;
  int nq = 0;
  // cells around skeleton faces (gtk_set_skeleton_cells)
  bool skel = false;
  int nlnD = 0;
  int64_t n_Dcells = 0;
  int32_t* cellD_nodes = nullptr; size_t cellD_n = 0;
  int32_t* side_cells = nullptr;  size_t side_n = 0;
  double* skel_tab = nullptr;     size_t skel_tab_n = 0;   // dMc then nref
  size_t nref_at = 0;
};

inline PartsState* parts(gtk_ctx* ctx) { return static_cast<PartsState*>(ctx->parts); }

int32_t ensure_d(gtk_ctx* ctx, double** p, size_t* cap, size_t n) {
  if (*cap >= n && *p) return GTK_OK;
  if (*p) gtk_dev_free(ctx, *p, *cap * sizeof(double));
  *p = nullptr; *cap = 0;
  if (n == 0) n = 1;
  int32_t rc = gtk_dev_alloc(ctx, (void**)p, n * sizeof(double));
  if (rc == GTK_OK) *cap = n;
  return rc;
}

int32_t fill(gtk_ctx* ctx, BlockArgs& a) {
  PartsState* ps = parts(ctx);
  if (!ps) GTK_FAIL(GTK_ERR_STATE, "gtk_set_parts must be called first (after gtk_set_mesh and gtk_set_space)");
  if (!ctx->xyz || !ctx->cell_dofs || ps->n_cells != ctx->n_cells || ps->L != ctx->nld)
    GTK_FAIL(GTK_ERR_STATE, "mesh or space changed after gtk_set_parts: call it again");
  a.xyz = ctx->xyz; a.cell_nodes = ctx->cell_nodes; a.n_cells = ctx->n_cells;
  a.nln = ctx->nln; a.nq = ctx->nq; a.n_parts = ps->n_parts; a.L = ps->L; a.n_sides = ps->n_sides;
  a.nls_total = ps->nls_total; a.need_grad = 0;
  a.w = ctx->w; a.dM = ctx->dM;
  for (int p = 0; p < BP_MAX; ++p) {
    const bool on = p < ps->n_parts;
    a.p_nls[p] = on ? ps->nls[p] : 0; a.p_ncomp[p] = on ? ps->ncomp[p] : 1; a.p_off[p] = on ? ps->off[p] : 0;
    a.p_side[p] = on ? ps->side[p] : 0; a.p_goff[p] = on ? ps->goff[p] : 0;
    a.p_N[p] = on ? ps->tab + ps->N_at[p] : nullptr;
    a.p_dN[p] = on && ps->has_dN[p] ? ps->tab + ps->dN_at[p] : nullptr;
    a.v_alpha[p] = 0.0;
    for (int k = 0; k < 3; ++k) a.v_c[p][k] = 0.0;
    for (int k = 0; k < 3; ++k) a.v_f[p][k] = 0.0;
    for (int q = 0; q < BP_MAX; ++q) { a.b_form[p][q] = GTK_BLOCK_ZERO; a.b_alpha[p][q] = 0.0; }
  }
  a.face_var = ps->face_var;
  a.v_g = nullptr;
  a.skel = ps->skel ? 1 : 0;
  a.cellD_nodes = ps->cellD_nodes; a.side_cells = ps->side_cells; a.nlnD = ps->nlnD;
  a.dMc = ps->skel_tab; a.nref = ps->skel_tab ? ps->skel_tab + ps->nref_at : nullptr;
  for (int p = 0; p < BP_MAX; ++p) for (int q = 0; q < BP_MAX; ++q) for (int k = 0; k < 3; ++k) a.b_c[p][q][k] = 0.0;
  a.act0 = ctx->act_count < 0 ? 0 : ctx->act_first;
  a.act1 = ctx->act_count < 0 ? ctx->n_cells : ctx->act_first + ctx->act_count;
  return GTK_OK;
}

int32_t pick(gtk_ctx* ctx, size_t per_cell_bytes, int* cb, size_t* smem) {
  int c = (int)((96 * 1024) / per_cell_bytes);
  if (c > 32) c = 32;
  if (c < 1) c = (int)((ctx->smem_optin - 2048) / per_cell_bytes);
  if (c < 1) GTK_FAIL(GTK_ERR_TOO_LARGE, "super element too large for the shared-memory block kernel");
  *cb = c;
  *smem = c * per_cell_bytes;
  return GTK_OK;
}

}  // namespace

void gtk_parts_release(gtk_ctx* ctx) {
  PartsState* ps = parts(ctx);
  if (!ps) return;
  if (ps->tab) gtk_dev_free(ctx, ps->tab, ps->tab_n * sizeof(double));
  if (ps->face_var) gtk_dev_free(ctx, ps->face_var, ps->face_var_n * sizeof(int32_t));
  if (ps->cellD_nodes) gtk_dev_free(ctx, ps->cellD_nodes, ps->cellD_n * sizeof(int32_t));
  if (ps->side_cells) gtk_dev_free(ctx, ps->side_cells, ps->side_n * sizeof(int32_t));
  if (ps->skel_tab) gtk_dev_free(ctx, ps->skel_tab, ps->skel_tab_n * sizeof(double));
  delete ps;
  ctx->parts = nullptr;
}

extern "C" int32_t gtk_set_parts(gtk_ctx* ctx, int32_t n_q, const double* w, const double* M, const double* dM, int32_t n_parts,
                                 const gtk_part* pd, int32_t n_sides, int32_t n_var, const int32_t* face_var) {
  if (!ctx) return GTK_ERR_INVALID;
  if (n_q < 1 || !w || !M || !dM || !pd || n_parts < 1 || n_parts > BP_MAX || n_sides < 1 || n_sides > 2 || n_var < 1)
    GTK_FAIL(GTK_ERR_INVALID, "gtk_set_parts: bad arguments (1..8 parts, 1 or 2 sides)");
  if (!ctx->xyz || !ctx->cell_dofs) GTK_FAIL(GTK_ERR_STATE, "gtk_set_parts: set mesh and space first");
  if ((n_var > 1 || n_sides > 1) && !face_var) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_parts: face_var is required with several variants or sides");
  GTK_CK(cudaSetDevice(ctx->device));
  gtk_parts_release(ctx);
  {   // validate everything before any state is created: a failed call leaves the context without parts
    int total = 0;
    for (int p = 0; p < n_parts; ++p) {
      if (pd[p].n_lshape < 1 || pd[p].n_comp < 1 || pd[p].n_comp > 3 || pd[p].side < 0 || pd[p].side >= n_sides || !pd[p].N)
        GTK_FAIL(GTK_ERR_INVALID, "gtk_set_parts: bad part descriptor " + std::to_string(p));
      total += pd[p].n_lshape * pd[p].n_comp;
    }
    if (total != ctx->nld)
      GTK_FAIL(GTK_ERR_INVALID, "gtk_set_parts: the parts hold " + std::to_string(total) + " local dofs, the dof table of gtk_set_space " + std::to_string(ctx->nld));
    if (face_var) {
      const size_t n = (size_t)ctx->n_cells * n_sides;
      for (size_t i = 0; i < n; ++i)
        if (face_var[i] < 0 || face_var[i] >= n_var) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_parts: face_var out of range");
    }
  }
  PartsState* ps = new PartsState();
  ctx->parts = ps;
  const int D = ctx->D, dm = ctx->dman;
  ps->n_parts = n_parts; ps->n_sides = n_sides; ps->n_var = n_var; ps->n_cells = ctx->n_cells; ps->nq = n_q;
  size_t at = 0;
  int off = 0, goff = 0;
  for (int p = 0; p < n_parts; ++p) {
    ps->nls[p] = pd[p].n_lshape; ps->ncomp[p] = pd[p].n_comp; ps->side[p] = pd[p].side;
    ps->off[p] = off; ps->goff[p] = goff; ps->has_dN[p] = pd[p].dN != nullptr;
    off += pd[p].n_lshape * pd[p].n_comp; goff += pd[p].n_lshape;
    ps->N_at[p] = at; at += (size_t)n_var * n_q * pd[p].n_lshape;
    ps->dN_at[p] = at; if (pd[p].dN) at += (size_t)n_var * n_q * pd[p].n_lshape * D;
  }
  ps->L = off; ps->nls_total = goff;
  std::vector<double> h(at);
  for (int p = 0; p < n_parts; ++p) {
    std::copy(pd[p].N, pd[p].N + (size_t)n_var * n_q * pd[p].n_lshape, h.begin() + ps->N_at[p]);
    if (pd[p].dN) std::copy(pd[p].dN, pd[p].dN + (size_t)n_var * n_q * pd[p].n_lshape * D, h.begin() + ps->dN_at[p]);
  }
  int32_t rc = gtk_dev_alloc(ctx, (void**)&ps->tab, std::max<size_t>(at, 1) * sizeof(double));
  if (rc) return rc;
  ps->tab_n = std::max<size_t>(at, 1);
  GTK_CK(cudaMemcpyAsync(ps->tab, h.data(), at * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (face_var) {
    const size_t n = (size_t)ctx->n_cells * n_sides;
    if ((rc = gtk_dev_alloc(ctx, (void**)&ps->face_var, std::max<size_t>(n, 1) * sizeof(int32_t)))) return rc;
    ps->face_var_n = std::max<size_t>(n, 1);
    GTK_CK(cudaMemcpyAsync(ps->face_var, face_var, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  // geometry of the integration faces: weights and the tabulated geometry functions, as gtk_set_tabulation keeps them
  ctx->nq = n_q;
  const size_t nM = (size_t)n_q * ctx->nln;
  auto up = [&](double** dst, size_t* have, const double* src, size_t n) -> int32_t {
    if (*dst && *have != n) { gtk_dev_free(ctx, *dst, *have * sizeof(double)); *dst = nullptr; *have = 0; }
    if (!*dst) { int32_t r = gtk_dev_alloc(ctx, (void**)dst, n * sizeof(double)); if (r) return r; *have = n; }
    GTK_CK(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return GTK_OK;
  };
  if ((rc = up(&ctx->w, &ctx->sz.w, w, (size_t)n_q))) return rc;
  if ((rc = up(&ctx->M, &ctx->sz.M, M, nM))) return rc;
  if ((rc = up(&ctx->dM, &ctx->sz.dM, dM, nM * dm))) return rc;
  ctx->h_w.assign(w, w + n_q);
  ctx->h_M.assign(M, M + nM);
  ctx->h_dM.assign(dM, dM + nM * dm);
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

extern "C" int32_t gtk_set_skeleton_cells(gtk_ctx* ctx, int64_t n_cells, int32_t n_lnodes, const int32_t* cell_nodes, const int32_t* side_cells,
                                          const double* dM_cell, const double* ref_normals) {
  if (!ctx) return GTK_ERR_INVALID;
  PartsState* ps = parts(ctx);
  if (!ps) GTK_FAIL(GTK_ERR_STATE, "gtk_set_skeleton_cells: call gtk_set_parts first");
  if (ctx->dman != ctx->D - 1) GTK_FAIL(GTK_ERR_STATE, "gtk_set_skeleton_cells: the integration faces must be (D-1)-faces (skeleton: two cells around, boundary: one)");
  if (n_cells < 1 || n_lnodes < 2 || !cell_nodes || !side_cells || !dM_cell || !ref_normals) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_skeleton_cells: bad arguments");
  GTK_CK(cudaSetDevice(ctx->device));
  const int D = ctx->D;
  const size_t nf2 = (size_t)ctx->n_cells * ps->n_sides;
  for (size_t i = 0; i < nf2; ++i)
    if (side_cells[i] < 1 || side_cells[i] > n_cells) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_skeleton_cells: side_cells out of range");
  for (size_t i = 0; i < (size_t)n_cells * n_lnodes; ++i)
    if (cell_nodes[i] < 1 || cell_nodes[i] > ctx->n_nodes) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_skeleton_cells: cell_nodes out of range");
  if (ps->cellD_nodes) { gtk_dev_free(ctx, ps->cellD_nodes, ps->cellD_n * sizeof(int32_t)); ps->cellD_nodes = nullptr; }
  if (ps->side_cells) { gtk_dev_free(ctx, ps->side_cells, ps->side_n * sizeof(int32_t)); ps->side_cells = nullptr; }
  if (ps->skel_tab) { gtk_dev_free(ctx, ps->skel_tab, ps->skel_tab_n * sizeof(double)); ps->skel_tab = nullptr; }
  ps->skel = false;
  int32_t rc;
  ps->cellD_n = (size_t)n_cells * n_lnodes; ps->side_n = nf2 > 0 ? nf2 : 1;
  const size_t ndm = (size_t)ps->n_var * ps->nq * n_lnodes * D, nnr = (size_t)ps->n_var * D;
  ps->skel_tab_n = ndm + nnr; ps->nref_at = ndm;
  if ((rc = gtk_dev_alloc(ctx, (void**)&ps->cellD_nodes, ps->cellD_n * sizeof(int32_t)))) return rc;
  if ((rc = gtk_dev_alloc(ctx, (void**)&ps->side_cells, ps->side_n * sizeof(int32_t)))) return rc;
  if ((rc = gtk_dev_alloc(ctx, (void**)&ps->skel_tab, ps->skel_tab_n * sizeof(double)))) return rc;
  GTK_CK(cudaMemcpyAsync(ps->cellD_nodes, cell_nodes, ps->cellD_n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (nf2) GTK_CK(cudaMemcpyAsync(ps->side_cells, side_cells, nf2 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  GTK_CK(cudaMemcpyAsync(ps->skel_tab, dM_cell, ndm * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GTK_CK(cudaMemcpyAsync(ps->skel_tab + ndm, ref_normals, nnr * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  ps->nlnD = n_lnodes; ps->n_Dcells = n_cells; ps->skel = true;
  return GTK_OK;
}

static int32_t launch_blocks(gtk_ctx* ctx, BlockArgs& a, bool matrix) {
  const int D = ctx->D, dm = ctx->dman;
  size_t per_cell = (size_t)ctx->nq * sizeof(double);
  if (matrix && a.need_grad) per_cell += (size_t)ctx->nq * a.nls_total * D * sizeof(double);
  if (!matrix && a.skel) per_cell += (size_t)ctx->nq * a.nls_total * D * sizeof(double);
  if (a.skel) per_cell += ((size_t)ctx->nq * a.n_sides * D + 1) * sizeof(double);
  size_t smem;
  int32_t rc = pick(ctx, per_cell, &a.cb, &smem);
  if (rc) return rc;
  const int grid = (int)((ctx->n_cells + a.cb - 1) / a.cb);
  cudaStream_t st = ctx->stream;
#define LAUNCH_B(DD, dd)                                                                                            \
  do {                                                                                                              \
    if (matrix) {                                                                                                   \
      GTK_CK(cudaFuncSetAttribute(k_elem_blocks<DD, dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
      { GtkProf pr_(ctx, "k_elem_blocks"); k_elem_blocks<DD, dd><<<grid, 128, smem, st>>>(a); }                      \
    } else {                                                                                                        \
      GTK_CK(cudaFuncSetAttribute(k_elem_vblocks<DD, dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      { GtkProf pr_(ctx, "k_elem_vblocks"); k_elem_vblocks<DD, dd><<<grid, 128, smem, st>>>(a); }                    \
    }                                                                                                               \
  } while (0)
  if (D == 1 && dm == 1) LAUNCH_B(1, 1); else if (D == 2 && dm == 2) LAUNCH_B(2, 2); else if (D == 3 && dm == 3) LAUNCH_B(3, 3);
  else if (D == 2 && dm == 1) LAUNCH_B(2, 1); else if (D == 3 && dm == 2) LAUNCH_B(3, 2); else if (D == 3 && dm == 1) LAUNCH_B(3, 1);
  else GTK_FAIL(GTK_ERR_INVALID, "D must be 1, 2 or 3 and 1 <= manifold dimension <= D");
#undef LAUNCH_B
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return GTK_OK;
}

extern "C" int32_t gtk_matrix_numeric_blocks_device(gtk_ctx* ctx, int32_t n_blocks, const gtk_block* blocks) {
  if (!ctx) return GTK_ERR_INVALID;
  if (n_blocks < 0 || (n_blocks && !blocks)) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_numeric_blocks: bad arguments");
  if (!ctx->ms.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_symbolic must be called before gtk_matrix_numeric_blocks");
  GTK_CK(cudaSetDevice(ctx->device));
  BlockArgs a;
  int32_t rc = fill(ctx, a);
  if (rc) return rc;
  PartsState* ps = parts(ctx);
  for (int b = 0; b < n_blocks; ++b) {
    const gtk_block& k = blocks[b];
    if (k.part_u < 0 || k.part_u >= ps->n_parts || k.part_v < 0 || k.part_v >= ps->n_parts)
      GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_numeric_blocks: part index out of range");
    const int cu = ps->ncomp[k.part_u], cv = ps->ncomp[k.part_v];
    const bool grad = k.form == GTK_BLOCK_LAPLACE || k.form == GTK_BLOCK_VALU_DIVV || k.form == GTK_BLOCK_DIVU_VALV;
    bool ok = false;
    if (k.form == GTK_BLOCK_ZERO) ok = true;
    else if (k.form == GTK_BLOCK_IP || k.form == GTK_BLOCK_IP_NOH) {
      if (!ps->skel || ctx->dman == ctx->D)
        GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "interior-penalty blocks need a skeleton measure and gtk_set_skeleton_cells; no CPU fallback");
      ok = cu == 1 && cv == 1 && ps->has_dN[k.part_u] && ps->has_dN[k.part_v];
    }
    else if (k.form == GTK_BLOCK_MASS) ok = cu == cv;
    else if (k.form == GTK_BLOCK_LAPLACE) ok = cu == cv && ps->has_dN[k.part_u] && ps->has_dN[k.part_v];
    else if (k.form == GTK_BLOCK_VALU_DIVV) ok = cu == 1 && cv == ctx->D && ps->has_dN[k.part_v];
    else if (k.form == GTK_BLOCK_DIVU_VALV) ok = cv == 1 && cu == ctx->D && ps->has_dN[k.part_u];
    if (!ok)
      GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "block form " + std::to_string(k.form) + " is not available for parts (" + std::to_string(k.part_u) + ", " +
                                             std::to_string(k.part_v) + ") (components / gradient tables do not fit); no CPU fallback");
    if (grad && ctx->dman != ctx->D)
      GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "blocks with gradients are supported on volume cells only; no CPU fallback");
    if (a.b_form[k.part_u][k.part_v] != GTK_BLOCK_ZERO)
      GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "one term per (part_u, part_v) block; sums inside a block are not recognised; no CPU fallback");
    a.b_form[k.part_u][k.part_v] = k.form;
    a.b_alpha[k.part_u][k.part_v] = k.alpha;
    for (int c = 0; c < 3; ++c) a.b_c[k.part_u][k.part_v][c] = k.c[c];
    if (grad || k.form == GTK_BLOCK_IP || k.form == GTK_BLOCK_IP_NOH) a.need_grad = 1;
  }
  {
    bool any_ip = false;
    for (int b = 0; b < n_blocks; ++b) any_ip |= blocks[b].form == GTK_BLOCK_IP || blocks[b].form == GTK_BLOCK_IP_NOH;
    if (!any_ip) a.skel = 0;
  }
  ctx->launches_last = 0;
  ctx->fast_path_last = 7;
  gtk_prof_reset(ctx);
  MatSym& m = ctx->ms;
  if (!m.generic_plan && (rc = gtk_symbolic_generic_plan(ctx))) return rc;
  if ((rc = ensure_d(ctx, &ctx->KE, &ctx->KE_cap, (size_t)m.n_full))) return rc;
  if ((rc = ensure_d(ctx, &ctx->nzval, &ctx->nzval_cap, (size_t)m.nnz))) return rc;
  if (ctx->n_cells == 0 || m.nnz == 0) return GTK_OK;
  a.out = ctx->KE;
  if ((rc = launch_blocks(ctx, a, true))) return rc;
  return gtk_reduce_nz_launch(ctx);
}

extern "C" int32_t gtk_matrix_numeric_blocks(gtk_ctx* ctx, int32_t n_blocks, const gtk_block* blocks, double* nzval) {
  int32_t rc = gtk_matrix_numeric_blocks_device(ctx, n_blocks, blocks);
  if (rc) return rc;
  if (!nzval) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_numeric_blocks: nzval is null");
  return gtk_copy_nzval(ctx, nzval);
}

extern "C" int32_t gtk_vector_assemble_blocks_data_device(gtk_ctx* ctx, int32_t n, const gtk_vblock* vb, const double* g_qp, int32_t accumulate) {
  if (!ctx) return GTK_ERR_INVALID;
  if (n < 0 || (n && !vb)) GTK_FAIL(GTK_ERR_INVALID, "gtk_vector_assemble_blocks: bad arguments");
  GTK_CK(cudaSetDevice(ctx->device));
  BlockArgs a;
  int32_t rc = fill(ctx, a);
  if (rc) return rc;
  PartsState* ps = parts(ctx);
  for (int b = 0; b < n; ++b) {
    if (vb[b].part < 0 || vb[b].part >= ps->n_parts) GTK_FAIL(GTK_ERR_INVALID, "gtk_vector_assemble_blocks: part index out of range");
    if (a.v_alpha[vb[b].part] != 0.0) GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "one term per part; no CPU fallback");
    a.v_alpha[vb[b].part] = vb[b].alpha;
    for (int k = 0; k < 3; ++k) { a.v_f[vb[b].part][k] = vb[b].f_const[k]; a.v_c[vb[b].part][k] = vb[b].c[k]; }
    if (g_qp && ps->ncomp[vb[b].part] != 1) GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "data terms need scalar parts; no CPU fallback");
    if (g_qp && (vb[b].c[1] != 0.0 || vb[b].c[2] != 0.0) && (!ps->skel || !ps->has_dN[vb[b].part]))
      GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "face-diameter / normal-derivative data terms need gtk_set_skeleton_cells and gradient tables; no CPU fallback");
  }
  if (g_qp) {
    const size_t ng = (size_t)ctx->n_cells * ctx->nq;
    if ((rc = ensure_d(ctx, &ctx->f_dev, &ctx->f_cap, ng))) return rc;
    GTK_CK(cudaMemcpyAsync(ctx->f_dev, g_qp, ng * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    GTK_CK(cudaStreamSynchronize(ctx->stream));   // borrowed host buffer
    a.v_g = ctx->f_dev;
    a.need_grad = a.skel;
  } else {
    a.skel = 0;
  }
  if (!ctx->vs.ready && (rc = gtk_symbolic_vector_impl(ctx, GTK_FREE))) return rc;
  VecSym& v = ctx->vs;
  if (!v.generic_plan && (rc = gtk_symbolic_vector_generic_plan(ctx))) return rc;
  ctx->launches_last = 0;
  ctx->fast_path_last = 7;
  gtk_prof_reset(ctx);
  if ((rc = ensure_d(ctx, &ctx->BE, &ctx->BE_cap, (size_t)v.n_full))) return rc;
  if ((rc = ensure_d(ctx, &ctx->bvec, &ctx->bvec_cap, (size_t)v.n_rows))) return rc;
  if (!accumulate) GTK_CK(cudaMemsetAsync(ctx->bvec, 0, sizeof(double) * (size_t)(v.n_rows > 0 ? v.n_rows : 1), ctx->stream));
  if (ctx->n_cells == 0 || v.n_urows == 0) return GTK_OK;
  a.out = ctx->BE;
  if ((rc = launch_blocks(ctx, a, false))) return rc;
  return gtk_reduce_rows_launch(ctx, accumulate ? 1 : 0);
}

extern "C" int32_t gtk_vector_assemble_blocks_device(gtk_ctx* ctx, int32_t n, const gtk_vblock* vb, int32_t accumulate) {
  return gtk_vector_assemble_blocks_data_device(ctx, n, vb, nullptr, accumulate);
}

extern "C" int32_t gtk_vector_assemble_blocks_data(gtk_ctx* ctx, int32_t n, const gtk_vblock* vb, const double* g_qp, int32_t accumulate, double* b) {
  int32_t rc = gtk_vector_assemble_blocks_data_device(ctx, n, vb, g_qp, accumulate);
  if (rc) return rc;
  if (!b) GTK_FAIL(GTK_ERR_INVALID, "gtk_vector_assemble_blocks_data: b is null");
  return gtk_copy_vector(ctx, b);
}

extern "C" int32_t gtk_vector_assemble_blocks(gtk_ctx* ctx, int32_t n, const gtk_vblock* vb, int32_t accumulate, double* b) {
  int32_t rc = gtk_vector_assemble_blocks_device(ctx, n, vb, accumulate);
  if (rc) return rc;
  if (!b) GTK_FAIL(GTK_ERR_INVALID, "gtk_vector_assemble_blocks: b is null");
  return gtk_copy_vector(ctx, b);
}
