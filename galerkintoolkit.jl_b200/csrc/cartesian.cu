// GT.cartesian_mesh(domain, cells) + lagrange_space(Ω, 1; dirichlet_boundary = boundary(mesh)) generated IN HBM
// (gtk_set_cartesian_q1_problem) — the synthetic inputs of BASELINE configs 2 and 5 without a host round trip.
//
//   cartesian_mesh.jl:213-263   node id 1 + i + (n1+1) j + (n1+1)(n2+1) k, x = pmin + h .* (i,j,k), h = (pmax-pmin) ./ cells;
//                               cells x-fastest, local nodes in tensor order
//   topology.jl:1034-1097       vertex ids: the 2^D box corners first (lexicographic), then the other nodes in node order
//   space.jl:327-417, 910-920   Q1: dof = vertex id; Dirichlet dofs (whole boundary) -> -(1..ndiri) in increasing old id
//                               (corners -1..-8, then the other boundary nodes lexicographically); free dofs 1..nfree in
//                               increasing old id = interior nodes lexicographically
// A z-slab [kz0, kz1] of node layers (multi-GPU partition, partition.py: slab_problem) uses slab-local ids: free dofs
// lexicographic among the local free nodes, Dirichlet ids lexicographic among the local boundary nodes; whether a node is
// free is decided by its GLOBAL position.  Everything is closed-form per node / per cell: no scan, no sort.
// The coordinates are computed as numpy / Julia do (one multiplication, one addition, no FMA contraction), so they are
// bit-identical to the host-generated arrays (tests/test_gpu_cartesian.py).
#include "gtk_internal.h"

void gtk_fastq1_release(gtk_ctx* ctx);

namespace {

struct CartArgs {
  int64_t n1, n2, n3;        // cells of the WHOLE mesh per direction
  int64_t kz0, kz1;          // node layers [kz0, kz1] present locally
  double pmin[3], h[3];
  int mode;                  // 0 reference numbering (whole mesh), 1 slab-local numbering
};

__device__ __forceinline__ int64_t clampi(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

// dof id of node (i, j, kg) with kg the GLOBAL layer
__device__ __forceinline__ int32_t node_dof(const CartArgs& a, int64_t i, int64_t j, int64_t kg) {
  const int64_t n1 = a.n1, n2 = a.n2, n3 = a.n3;
  const bool ii = i > 0 && i < n1, ji = j > 0 && j < n2, ki = kg > 0 && kg < n3;
  const int64_t k0 = a.mode == 1 ? a.kz0 : 0;                      // first layer that counts
  // interior nodes strictly before (i,j,kg) in lexicographic order (x fastest), counted from layer k0
  const int64_t int_layers_before = clampi(kg - 1, 0, n3 - 1) - clampi(k0 - 1, 0, n3 - 1);   // interior layers in [k0, kg)
  int64_t before_int = int_layers_before * (n1 - 1) * (n2 - 1);
  if (ki) before_int += clampi(j - 1, 0, n2 - 1) * (n1 - 1) + (ji ? clampi(i - 1, 0, n1 - 1) : 0);
  if (ii && ji && ki) return (int32_t)(before_int + 1);
  const int64_t before_all = i + (n1 + 1) * j + (n1 + 1) * (n2 + 1) * (kg - k0);
  const int64_t before_bnd = before_all - before_int;
  if (a.mode == 1) return (int32_t)(-(before_bnd + 1));
  const bool ci = i == 0 || i == n1, cj = j == 0 || j == n2, ck = kg == 0 || kg == n3;
  if (ci && cj && ck) return -(int32_t)(1 + (i > 0) + 2 * (j > 0) + 4 * (kg > 0));
  // corners strictly before this node
  int64_t cb = 0;
  for (int c = 0; c < 8; ++c) {
    const int64_t x = (c & 1) ? n1 : 0, y = (c & 2) ? n2 : 0, z = (c & 4) ? n3 : 0;
    cb += (x + (n1 + 1) * y + (n1 + 1) * (n2 + 1) * z) < before_all;
  }
  return (int32_t)(-(8 + (before_bnd - cb) + 1));
}

__global__ void k_cart_nodes(CartArgs a, double* __restrict__ xyz) {
  const int64_t npl = (a.n1 + 1) * (a.n2 + 1), n_nodes = npl * (a.kz1 - a.kz0 + 1);
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_nodes; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = v / npl, r = v - k * npl, j = r / (a.n1 + 1), i = r - j * (a.n1 + 1);
    xyz[3 * v + 0] = __dadd_rn(a.pmin[0], __dmul_rn(a.h[0], (double)i));
    xyz[3 * v + 1] = __dadd_rn(a.pmin[1], __dmul_rn(a.h[1], (double)j));
    xyz[3 * v + 2] = __dadd_rn(a.pmin[2], __dmul_rn(a.h[2], (double)(k + a.kz0)));
  }
}

// thread per (cell, local node): coalesced 4-byte stores
__global__ void k_cart_cells(CartArgs a, int32_t* __restrict__ cell_nodes, int32_t* __restrict__ cell_dofs) {
  const int64_t n_cells = a.n1 * a.n2 * (a.kz1 - a.kz0), n = n_cells * 8;
  const int64_t sx = 1, sy = a.n1 + 1, sz = (a.n1 + 1) * (a.n2 + 1);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = e >> 3;
    const int ln = (int)(e & 7);
    const int64_t ck = cell / (a.n1 * a.n2), r = cell - ck * a.n1 * a.n2, cj = r / a.n1, ci = r - cj * a.n1;
    const int64_t i = ci + (ln & 1), j = cj + ((ln >> 1) & 1), k = ck + ((ln >> 2) & 1);
    cell_nodes[e] = (int32_t)(1 + i * sx + j * sy + k * sz);
    cell_dofs[e] = node_dof(a, i, j, k + a.kz0);
  }
}

template <class T>
int32_t fresh(gtk_ctx* ctx, T** p, size_t* old_n, size_t n) {
  if (*p && *old_n != n) { gtk_dev_free(ctx, *p, *old_n * sizeof(T)); *p = nullptr; }
  if (!*p) { int32_t rc = gtk_dev_alloc(ctx, (void**)p, n * sizeof(T)); if (rc) return rc; }
  *old_n = n;
  return GTK_OK;
}

}  // namespace

extern "C" int32_t gtk_set_cartesian_q1_problem(gtk_ctx* ctx, const double* domain6, const int64_t* cells3, int64_t kz0, int64_t kz1,
                                                int32_t slab_local_numbering, int64_t* n_free_out, int64_t* n_dirichlet_out) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!domain6 || !cells3) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_cartesian_q1_problem: null argument");
  const int64_t n1 = cells3[0], n2 = cells3[1], n3 = cells3[2];
  if (n1 < 2 || n2 < 2 || n3 < 2 || kz0 < 0 || kz1 > n3 || kz1 <= kz0)
    GTK_FAIL(GTK_ERR_INVALID, "gtk_set_cartesian_q1_problem: needs >= 2 cells per direction and node layers 0 <= kz0 < kz1 <= cells[2]");
  if (!slab_local_numbering && (kz0 != 0 || kz1 != n3))
    GTK_FAIL(GTK_ERR_INVALID, "gtk_set_cartesian_q1_problem: the reference numbering needs the whole mesh (kz0 = 0, kz1 = cells[2])");
  const int64_t npl = (n1 + 1) * (n2 + 1), n_nodes = npl * (kz1 - kz0 + 1), n_cells = n1 * n2 * (kz1 - kz0);
  // free nodes: interior in x, y and (globally) z
  const int64_t lo = kz0 < 1 ? 1 : kz0, hi = kz1 > n3 - 1 ? n3 - 1 : kz1;
  const int64_t n_free = (hi >= lo ? hi - lo + 1 : 0) * (n1 - 1) * (n2 - 1);
  if (n_nodes >= 0x7FFFFFFFll || n_cells * 8 < 0) GTK_FAIL(GTK_ERR_TOO_LARGE, "node ids are Int32 on this ABI");
  GTK_CK(cudaSetDevice(ctx->device));
  auto& sz = ctx->sz;
  int32_t rc;
  if ((rc = fresh(ctx, &ctx->xyz, &sz.xyz, (size_t)n_nodes * 3))) return rc;
  if ((rc = fresh(ctx, &ctx->cell_nodes, &sz.cell_nodes, (size_t)n_cells * 8))) return rc;
  if ((rc = fresh(ctx, &ctx->cell_dofs, &sz.cell_dofs, (size_t)n_cells * 8))) return rc;
  ctx->D = 3; ctx->dman = 3; ctx->n_nodes = n_nodes; ctx->n_cells = n_cells; ctx->nln = 8;
  ctx->act_first = 0; ctx->act_count = -1;
  if (ctx->nls != 8) ctx->nq = 0;
  ctx->nld = 8; ctx->ncomp = 1; ctx->nls = 8;
  ctx->n_free = n_free; ctx->n_diri = n_nodes - n_free;
  gtk_field_release(ctx);
  gtk_release_all_matrices(ctx); gtk_vecsym_release(ctx);
  CartArgs a;
  a.n1 = n1; a.n2 = n2; a.n3 = n3; a.kz0 = kz0; a.kz1 = kz1; a.mode = slab_local_numbering ? 1 : 0;
  for (int d = 0; d < 3; ++d) {
    a.pmin[d] = domain6[2 * d];
    a.h[d] = (domain6[2 * d + 1] - domain6[2 * d]) / (double)cells3[d];   // (pmax - pmin) ./ cells
  }
  auto grid = [&](int64_t n) { int64_t g = (n + 255) / 256; int64_t cap = (int64_t)ctx->sm_count * 16; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); };
  k_cart_nodes<<<grid(n_nodes), 256, 0, ctx->stream>>>(a, ctx->xyz);
  k_cart_cells<<<grid(n_cells * 8), 256, 0, ctx->stream>>>(a, ctx->cell_nodes, ctx->cell_dofs);
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx, 2);
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  if (n_free_out) *n_free_out = n_free;
  if (n_dirichlet_out) *n_dirichlet_out = n_nodes - n_free;
  return GTK_OK;
}
