// The DiscreteField parameter u_h of a context and what interpolation needs from the device.
//
//   field.jl:93-125      DiscreteField = (space, free_values, dirichlet_values)
//   problems.jl:465-497  nonlinear problems re-assemble residual / Jacobian with parameters=(uh,)
//   problems.jl:519-526  solution_field!(uh, x): free values <- x
//   space.jl:1876-1897   node_coordinates(::LagrangeMeshSpace): loop over cells, x = Σ tab[lnode,lmnode] x_mnode from zero,
//                        the last cell holding a node wins
//   space.jl:2000-2060   interpolate_impl!: v = fun(node_x[node]) per dof, free / Dirichlet by sign
//
// The values live in HBM next to the pattern; the PLAPLACE_* kernels (numeric.cu) and the scalar integrals gather them per
// cell.  "Last cell wins" is made deterministic with an integer atomicMax over (cell, local dof) keys — integer atomics
// only, the coordinates themselves are computed once per dof by the winning (cell, local node).
#include "gtk_internal.h"

namespace {

int32_t ensure_zero(gtk_ctx* ctx, double** p, size_t* cap, size_t n) {
  if (*p && *cap == (n ? n : 1)) return GTK_OK;
  if (*p) gtk_dev_free(ctx, *p, *cap * sizeof(double));
  *p = nullptr; *cap = 0;
  const size_t m = n ? n : 1;
  int32_t rc = gtk_dev_alloc(ctx, (void**)p, m * sizeof(double));
  if (rc) return rc;
  *cap = m;
  GTK_CK(cudaMemsetAsync(*p, 0, m * sizeof(double), ctx->stream));
  return GTK_OK;
}

__global__ void k_axpy(double* __restrict__ y, const double* __restrict__ x, double a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = y[i] + a * x[i];
}

// key[dof] = max over (cell, ldof) holding it of cell*nld + ldof: the LAST writer of the reference's cell loop
__global__ void k_last_holder(const int32_t* __restrict__ cell_dofs, int64_t n_full, unsigned long long* __restrict__ key_free,
                              unsigned long long* __restrict__ key_diri) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_full; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = cell_dofs[e];
    if (d > 0) atomicMax(key_free + (d - 1), (unsigned long long)e + 1ull);
    else if (d < 0) atomicMax(key_diri + (-d - 1), (unsigned long long)e + 1ull);
  }
}

// x[dof] = Σ_lmnode tab[lnode,lmnode] * x_mnode, sequential from zero (space.jl:1888-1892), by the winning cell
__global__ void k_dof_coordinates(const unsigned long long* __restrict__ key, int64_t n_dofs, int nld, int ncomp, int nln, int D,
                                  const double* __restrict__ Mn, const double* __restrict__ xyz,
                                  const int32_t* __restrict__ cell_nodes, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_dofs; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = key[i];
    if (k == 0ull) { for (int a = 0; a < D; ++a) out[i * D + a] = 0.0; continue; }   // dof held by no cell
    const int64_t e = (int64_t)(k - 1ull);
    const int64_t cell = e / nld;
    const int lnode = (int)(e - cell * nld) / ncomp;
    const int32_t* nodes = cell_nodes + cell * nln;
    for (int a = 0; a < D; ++a) {
      double x = 0.0;
      for (int m = 0; m < nln; ++m) x += Mn[lnode * nln + m] * xyz[(size_t)(nodes[m] - 1) * D + a];
      out[i * D + a] = x;
    }
  }
}

inline int grid_for(int64_t n, int block, int sm) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int32_t gtk_field_ensure(gtk_ctx* ctx) {
  if (!ctx->cell_dofs) GTK_FAIL(GTK_ERR_STATE, "the discrete field needs a space: call gtk_set_space first");
  int32_t rc = ensure_zero(ctx, &ctx->u_free, &ctx->u_free_cap, (size_t)ctx->n_free);
  if (rc) return rc;
  return ensure_zero(ctx, &ctx->u_diri, &ctx->u_diri_cap, (size_t)ctx->n_diri);
}

void gtk_field_release(gtk_ctx* ctx) {
  if (ctx->u_free) gtk_dev_free(ctx, ctx->u_free, ctx->u_free_cap * sizeof(double));
  if (ctx->u_diri) gtk_dev_free(ctx, ctx->u_diri, ctx->u_diri_cap * sizeof(double));
  if (ctx->xdof_free) gtk_dev_free(ctx, ctx->xdof_free, ctx->xdof_free_cap * sizeof(double));
  if (ctx->xdof_diri) gtk_dev_free(ctx, ctx->xdof_diri, ctx->xdof_diri_cap * sizeof(double));
  ctx->u_free = ctx->u_diri = ctx->xdof_free = ctx->xdof_diri = nullptr;
  ctx->u_free_cap = ctx->u_diri_cap = ctx->xdof_free_cap = ctx->xdof_diri_cap = 0;
}

extern "C" {

static int32_t field_set(gtk_ctx* ctx, const double* fv, const double* dv, cudaMemcpyKind kind) {
  if (!ctx) return GTK_ERR_INVALID;
  GTK_CK(cudaSetDevice(ctx->device));
  int32_t rc = gtk_field_ensure(ctx);
  if (rc) return rc;
  if (fv && ctx->n_free) GTK_CK(cudaMemcpyAsync(ctx->u_free, fv, sizeof(double) * (size_t)ctx->n_free, kind, ctx->stream));
  if (dv && ctx->n_diri) GTK_CK(cudaMemcpyAsync(ctx->u_diri, dv, sizeof(double) * (size_t)ctx->n_diri, kind, ctx->stream));
  // host buffers are only borrowed for the duration of the call (they may be pinned: the copy is then truly asynchronous)
  if (kind == cudaMemcpyHostToDevice) GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_field_set_values(gtk_ctx* ctx, const double* fv, const double* dv) { return field_set(ctx, fv, dv, cudaMemcpyHostToDevice); }

int32_t gtk_field_set_values_device(gtk_ctx* ctx, const double* fv, const double* dv) { return field_set(ctx, fv, dv, cudaMemcpyDeviceToDevice); }

int32_t gtk_field_get_values(gtk_ctx* ctx, double* fv, double* dv) {
  if (!ctx) return GTK_ERR_INVALID;
  GTK_CK(cudaSetDevice(ctx->device));
  int32_t rc = gtk_field_ensure(ctx);
  if (rc) return rc;
  if (fv && ctx->n_free) GTK_CK(cudaMemcpyAsync(fv, ctx->u_free, sizeof(double) * (size_t)ctx->n_free, cudaMemcpyDeviceToHost, ctx->stream));
  if (dv && ctx->n_diri) GTK_CK(cudaMemcpyAsync(dv, ctx->u_diri, sizeof(double) * (size_t)ctx->n_diri, cudaMemcpyDeviceToHost, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_field_axpy_free(gtk_ctx* ctx, double a, const double* dx) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!dx && ctx->n_free) GTK_FAIL(GTK_ERR_INVALID, "gtk_field_axpy_free: dx is null");
  GTK_CK(cudaSetDevice(ctx->device));
  int32_t rc = gtk_field_ensure(ctx);
  if (rc) return rc;
  if (ctx->n_free == 0) return GTK_OK;
  const size_t n = (size_t)ctx->n_free;
  if (ctx->xvec_cap < n) {
    if (ctx->xvec) gtk_dev_free(ctx, ctx->xvec, ctx->xvec_cap * sizeof(double));
    ctx->xvec = nullptr; ctx->xvec_cap = 0;
    if ((rc = gtk_dev_alloc(ctx, (void**)&ctx->xvec, n * sizeof(double)))) return rc;
    ctx->xvec_cap = n;
  }
  GTK_CK(cudaMemcpyAsync(ctx->xvec, dx, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_axpy<<<grid_for((int64_t)n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->u_free, ctx->xvec, a, (int64_t)n);
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_space_dof_coordinates(gtk_ctx* ctx, const double* M_at_nodes, double* x_free, double* x_dirichlet) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->xyz || !ctx->cell_dofs) GTK_FAIL(GTK_ERR_STATE, "gtk_space_dof_coordinates: set mesh and space first");
  if (!M_at_nodes) GTK_FAIL(GTK_ERR_INVALID, "gtk_space_dof_coordinates: M_at_nodes is null");
  GTK_CK(cudaSetDevice(ctx->device));
  const int D = ctx->D;
  const size_t nf = (size_t)ctx->n_free, nd = (size_t)ctx->n_diri;
  int32_t rc;
  if ((rc = ensure_zero(ctx, &ctx->xdof_free, &ctx->xdof_free_cap, nf * D))) return rc;
  if ((rc = ensure_zero(ctx, &ctx->xdof_diri, &ctx->xdof_diri_cap, nd * D))) return rc;
  unsigned long long *kf = nullptr, *kd = nullptr;
  double* Mn = nullptr;
  const size_t nMn = (size_t)ctx->nls * ctx->nln;
  GTK_CK(gtk_cuda_malloc(ctx, &kf, (nf ? nf : 1) * sizeof(unsigned long long)));
  GTK_CK(gtk_cuda_malloc(ctx, &kd, (nd ? nd : 1) * sizeof(unsigned long long)));
  GTK_CK(gtk_cuda_malloc(ctx, &Mn, nMn * sizeof(double)));
  cudaStream_t st = ctx->stream;
  GTK_CK(cudaMemsetAsync(kf, 0, (nf ? nf : 1) * sizeof(unsigned long long), st));
  GTK_CK(cudaMemsetAsync(kd, 0, (nd ? nd : 1) * sizeof(unsigned long long), st));
  GTK_CK(cudaMemcpyAsync(Mn, M_at_nodes, nMn * sizeof(double), cudaMemcpyHostToDevice, st));
  const int64_t n_full = ctx->n_cells * (int64_t)ctx->nld;
  if (n_full) {
    k_last_holder<<<grid_for(n_full, 256, ctx->sm_count), 256, 0, st>>>(ctx->cell_dofs, n_full, kf, kd);
    gtk_count_launch(ctx);
  }
  if (nf) {
    k_dof_coordinates<<<grid_for((int64_t)nf, 128, ctx->sm_count), 128, 0, st>>>(kf, (int64_t)nf, ctx->nld, ctx->ncomp, ctx->nln, D, Mn, ctx->xyz,
                                                                                 ctx->cell_nodes, ctx->xdof_free);
    gtk_count_launch(ctx);
  }
  if (nd) {
    k_dof_coordinates<<<grid_for((int64_t)nd, 128, ctx->sm_count), 128, 0, st>>>(kd, (int64_t)nd, ctx->nld, ctx->ncomp, ctx->nln, D, Mn, ctx->xyz,
                                                                                 ctx->cell_nodes, ctx->xdof_diri);
    gtk_count_launch(ctx);
  }
  GTK_CK(cudaGetLastError());
  if (x_free && nf) GTK_CK(cudaMemcpyAsync(x_free, ctx->xdof_free, nf * D * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (x_dirichlet && nd) GTK_CK(cudaMemcpyAsync(x_dirichlet, ctx->xdof_diri, nd * D * sizeof(double), cudaMemcpyDeviceToHost, st));
  GTK_CK(cudaStreamSynchronize(st));
  gtk_cuda_free(ctx, kf); gtk_cuda_free(ctx, kd); gtk_cuda_free(ctx, Mn);
  return GTK_OK;
}

int32_t gtk_scalar_assemble(gtk_ctx* ctx, int32_t kind, const gtk_form_params* p, double* out) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!out) GTK_FAIL(GTK_ERR_INVALID, "gtk_scalar_assemble: out is null");
  GTK_CK(cudaSetDevice(ctx->device));
  return gtk_scalar_impl(ctx, kind, p, out);
}

}  // extern "C"
