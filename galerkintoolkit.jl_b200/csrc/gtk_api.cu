// C ABI of libgtkasm (include/gtk_assembly.h): argument checks, HBM residency of the inputs,
// dispatch to symbolic.cu / numeric.cu / fastq1.cu / comm.cu.  No CPU compute path exists here:
// every numeric entry point launches CUDA kernels or returns an error.
#include <algorithm>
#include <cstring>
#include "gtk_internal.h"

cudaError_t gtk_cuda_malloc(gtk_ctx* ctx, void** p, size_t bytes) {
  *p = nullptr;
  const size_t sz = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
  auto& pool = ctx->pool;
  auto it = pool.cached.find(sz);
  if (it != pool.cached.end()) {
    *p = it->second;
    pool.cached.erase(it);
    pool.cached_bytes -= sz;
  } else {
    cudaError_t e = cudaMalloc(p, sz);
    if (e != cudaSuccess) {   // out of memory: give the cached blocks back to the driver and retry once
      cudaGetLastError();
      gtk_pool_flush(ctx);
      e = cudaMalloc(p, sz);
      if (e != cudaSuccess) { *p = nullptr; return e; }
    }
  }
  pool.live[*p] = sz;
  return cudaSuccess;
}

void gtk_cuda_free(gtk_ctx* ctx, void* p) {
  if (!p) return;
  auto& pool = ctx->pool;
  auto it = pool.live.find(p);
  if (it == pool.live.end()) { cudaFree(p); return; }   // not ours (should not happen)
  const size_t sz = it->second;
  pool.live.erase(it);
  if (pool.cached_bytes + sz <= pool.cap_bytes) {
    pool.cached.emplace(sz, p);
    pool.cached_bytes += sz;
  } else {
    cudaFree(p);
  }
}

void gtk_pool_flush(gtk_ctx* ctx) {
  for (auto& kv : ctx->pool.cached) cudaFree(kv.second);
  ctx->pool.cached.clear();
  ctx->pool.cached_bytes = 0;
}

int32_t gtk_dev_alloc(gtk_ctx* ctx, void** p, size_t bytes) {
  GTK_CK(gtk_cuda_malloc(ctx, p, bytes));
  ctx->bytes_held += (int64_t)bytes;
  return GTK_OK;
}

void gtk_dev_free(gtk_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  gtk_cuda_free(ctx, p);
  ctx->bytes_held -= (int64_t)bytes;
}

void gtk_fastq1_release(gtk_ctx* ctx);
void gtk_fastq1_coords_changed(gtk_ctx* ctx);
void gtk_comm_release(gtk_ctx* ctx);

namespace {
template <class T>
int32_t upload(gtk_ctx* ctx, T** dst, size_t* old_n, const T* src, size_t n) {
  if (*dst && *old_n != n) { gtk_dev_free(ctx, *dst, *old_n * sizeof(T)); *dst = nullptr; }
  if (!*dst) {
    int32_t rc = gtk_dev_alloc(ctx, (void**)dst, n * sizeof(T));
    if (rc) return rc;
  }
  *old_n = n;
  if (n) GTK_CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return GTK_OK;
}
}  // namespace

extern "C" {

int32_t gtk_version(void) { return 200; }   // round 2: gtk_part / gtk_block / gtk_vblock layouts, block, sum and field entry points

int32_t gtk_create(int32_t device, gtk_ctx** out) {
  if (!out) return GTK_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return GTK_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return GTK_ERR_CUDA;
  gtk_ctx* ctx = new gtk_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return GTK_ERR_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  *out = ctx;
  return GTK_OK;
}

int32_t gtk_destroy(gtk_ctx* ctx) {
  if (!ctx) return GTK_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  gtk_comm_release(ctx);
  gtk_parts_release(ctx);
  gtk_sumplan_release(ctx);
  gtk_release_all_matrices(ctx);
  gtk_vecsym_release(ctx);
  for (auto& sl : ctx->slots) { gtk_cuda_free(ctx, sl.nzval); sl.nzval = nullptr; }
  gtk_cuda_free(ctx, ctx->xvec);
  for (void* q : {(void*)ctx->xyz, (void*)ctx->cell_nodes, (void*)ctx->cell_dofs, (void*)ctx->w, (void*)ctx->N, (void*)ctx->dN,
                  (void*)ctx->M, (void*)ctx->dM, (void*)ctx->KE, (void*)ctx->BE, (void*)ctx->nzval, (void*)ctx->bvec,
                  (void*)ctx->f_dev, (void*)ctx->coef_dev, (void*)ctx->Cm, (void*)ctx->u_free, (void*)ctx->u_diri,
                  (void*)ctx->xdof_free, (void*)ctx->xdof_diri, (void*)ctx->scal_part})
    gtk_cuda_free(ctx, q);
  gtk_pool_flush(ctx);
  for (auto& kv : ctx->pool.live) cudaFree(kv.first);   // anything still registered
  delete ctx;
  return GTK_OK;
}

const char* gtk_last_error(const gtk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t gtk_set_stream(gtk_ctx* ctx, void* s) {
  if (!ctx) return GTK_ERR_INVALID;
  // everything issued so far (and every pooled block's last use) is ordered on the old stream
  if (ctx->stream != (cudaStream_t)s) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  ctx->stream = (cudaStream_t)s;
  return GTK_OK;
}

int32_t gtk_set_mesh(gtk_ctx* ctx, int32_t D, int64_t n_nodes, const double* xyz, int64_t n_cells,
                     int32_t n_lnodes, const int32_t* cell_nodes) {
  if (!ctx) return GTK_ERR_INVALID;
  if (D < 1 || D > 3 || n_nodes < 0 || n_cells < 0 || n_lnodes < 1 || (!xyz && n_nodes) || (!cell_nodes && n_cells))
    GTK_FAIL(GTK_ERR_INVALID, "gtk_set_mesh: bad arguments");
  GTK_CK(cudaSetDevice(ctx->device));
  auto& sz = ctx->sz;
  ctx->D = D; ctx->dman = D; ctx->n_nodes = n_nodes; ctx->n_cells = n_cells; ctx->nln = n_lnodes;
  ctx->act_first = 0; ctx->act_count = -1;
  // the space and the tabulation described the previous mesh: they must be handed over again
  if (ctx->cell_dofs) { gtk_dev_free(ctx, ctx->cell_dofs, sz.cell_dofs * sizeof(int32_t)); ctx->cell_dofs = nullptr; sz.cell_dofs = 0; }
  ctx->nld = 0; ctx->nls = 0; ctx->ncomp = 1; ctx->nq = 0;
  gtk_field_release(ctx);
  gtk_parts_release(ctx);
  int32_t rc = upload(ctx, &ctx->xyz, &sz.xyz, xyz, (size_t)n_nodes * D);
  if (rc) return rc;
  rc = upload(ctx, &ctx->cell_nodes, &sz.cell_nodes, cell_nodes, (size_t)n_cells * n_lnodes);
  if (rc) return rc;
  gtk_release_all_matrices(ctx); gtk_vecsym_release(ctx);
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_set_manifold_dim(gtk_ctx* ctx, int32_t d) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->D) GTK_FAIL(GTK_ERR_STATE, "gtk_set_manifold_dim: set the mesh first");
  if (d < 1 || d > ctx->D) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_manifold_dim: 1 <= d <= D");
  if (d != ctx->dman) { ctx->nq = 0; gtk_parts_release(ctx); }   // gradients were tabulated with another number of components
  ctx->dman = d;
  return GTK_OK;
}

int32_t gtk_set_vector(gtk_ctx* ctx, const double* b) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->vs.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_set_vector: call gtk_vector_symbolic first");
  if (!b && ctx->vs.n_rows) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_vector: b is null");
  GTK_CK(cudaSetDevice(ctx->device));
  const size_t n = (size_t)(ctx->vs.n_rows > 0 ? ctx->vs.n_rows : 1);
  if (ctx->bvec_cap < n || !ctx->bvec) {
    if (ctx->bvec) gtk_dev_free(ctx, ctx->bvec, ctx->bvec_cap * sizeof(double));
    ctx->bvec = nullptr; ctx->bvec_cap = 0;
    int32_t rc = gtk_dev_alloc(ctx, (void**)&ctx->bvec, n * sizeof(double));
    if (rc) return rc;
    ctx->bvec_cap = n;
  }
  if (ctx->vs.n_rows) GTK_CK(cudaMemcpyAsync(ctx->bvec, b, sizeof(double) * (size_t)ctx->vs.n_rows, cudaMemcpyHostToDevice, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_set_active_cells(gtk_ctx* ctx, int64_t first, int64_t count) {
  if (!ctx) return GTK_ERR_INVALID;
  if (first < 0 || count < 0 || first + count > ctx->n_cells) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_active_cells: range outside the mesh");
  ctx->act_first = first;
  ctx->act_count = count;
  // the affine classification only inspected the previously active layers: every slot's plan must classify again
  const int keep = ctx->cur_slot;
  for (int s = 0; s < gtk_ctx::N_SLOTS; ++s) { gtk_select_matrix_impl(ctx, s); gtk_fastq1_coords_changed(ctx); }
  gtk_select_matrix_impl(ctx, keep);
  return GTK_OK;
}

int32_t gtk_update_coordinates(gtk_ctx* ctx, const double* xyz) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->xyz || !xyz) GTK_FAIL(GTK_ERR_STATE, "gtk_update_coordinates: set the mesh first");
  GTK_CK(cudaSetDevice(ctx->device));
  GTK_CK(cudaMemcpyAsync(ctx->xyz, xyz, sizeof(double) * (size_t)ctx->n_nodes * ctx->D, cudaMemcpyHostToDevice, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  gtk_fastq1_coords_changed(ctx);
  return GTK_OK;
}

int32_t gtk_set_space(gtk_ctx* ctx, int32_t n_ldofs, int32_t n_comp, const int32_t* cell_dofs,
                      int64_t n_free, int64_t n_dirichlet) {
  if (!ctx) return GTK_ERR_INVALID;
  if (n_ldofs < 1 || n_comp < 1 || n_comp > 3 || n_ldofs % n_comp || n_free < 0 || n_dirichlet < 0 ||
      (!cell_dofs && ctx->n_cells))
    GTK_FAIL(GTK_ERR_INVALID, "gtk_set_space: bad arguments");
  if (n_free >= 0x7FFFFFFFll || n_dirichlet >= 0x7FFFFFFFll)
    GTK_FAIL(GTK_ERR_TOO_LARGE, "dof ids are Int32 on this ABI");
  GTK_CK(cudaSetDevice(ctx->device));
  auto& sz = ctx->sz;
  if (ctx->nls != n_ldofs / n_comp) ctx->nq = 0;   // N / dN were sized for another element
  ctx->nld = n_ldofs; ctx->ncomp = n_comp; ctx->nls = n_ldofs / n_comp;
  ctx->n_free = n_free; ctx->n_diri = n_dirichlet;
  gtk_field_release(ctx);   // the discrete field belongs to the previous space
  if (ctx->parts) { gtk_parts_release(ctx); ctx->nq = 0; }
  int32_t rc = upload(ctx, &ctx->cell_dofs, &sz.cell_dofs, cell_dofs, (size_t)ctx->n_cells * n_ldofs);
  if (rc) return rc;
  gtk_release_all_matrices(ctx); gtk_vecsym_release(ctx);
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_set_tabulation(gtk_ctx* ctx, int32_t n_q, const double* w, const double* N, const double* dN,
                           const double* M, const double* dM) {
  if (!ctx) return GTK_ERR_INVALID;
  if (n_q < 1 || !w || !N || !dN || !M || !dM) GTK_FAIL(GTK_ERR_INVALID, "gtk_set_tabulation: bad arguments");
  if (!ctx->D || !ctx->nls) GTK_FAIL(GTK_ERR_STATE, "gtk_set_tabulation: set mesh and space first");
  GTK_CK(cudaSetDevice(ctx->device));
  gtk_parts_release(ctx);   // a plain tabulation replaces the part tables of gtk_set_parts
  auto& sz = ctx->sz;
  ctx->nq = n_q;
  const int D = ctx->dman;   // reference-space dimension of the tabulated gradients
  size_t nN = (size_t)n_q * ctx->nls, nM = (size_t)n_q * ctx->nln;
  ctx->h_w.assign(w, w + n_q);
  ctx->h_N.assign(N, N + nN);
  ctx->h_dN.assign(dN, dN + nN * D);
  ctx->h_M.assign(M, M + nM);
  ctx->h_dM.assign(dM, dM + nM * D);
  int32_t rc;
  if ((rc = upload(ctx, &ctx->w, &sz.w, w, (size_t)n_q))) return rc;
  if ((rc = upload(ctx, &ctx->N, &sz.N, N, nN))) return rc;
  if ((rc = upload(ctx, &ctx->dN, &sz.dN, dN, nN * D))) return rc;
  if ((rc = upload(ctx, &ctx->M, &sz.M, M, nM))) return rc;
  if ((rc = upload(ctx, &ctx->dM, &sz.dM, dM, nM * D))) return rc;
  // the sweep plan depends on mesh + space only; whether the tabulation is the Q1/Gauss-2 one is checked per numeric call
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_matrix_symbolic(gtk_ctx* ctx, int32_t rfd, int32_t cfd, int64_t* nnz_out) {
  if (!ctx) return GTK_ERR_INVALID;
  if ((rfd != GTK_FREE && rfd != GTK_DIRICHLET) || (cfd != GTK_FREE && cfd != GTK_DIRICHLET))
    GTK_FAIL(GTK_ERR_INVALID, "free_or_dirichlet must be GTK_FREE or GTK_DIRICHLET");
  if (!ctx->cell_dofs) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_symbolic: set mesh and space first");
  GTK_CK(cudaSetDevice(ctx->device));
  gtk_fastq1_release(ctx);
  gtk_sumplan_release(ctx);
  int32_t rc = gtk_symbolic_matrix_impl(ctx, rfd, cfd);
  if (rc) return rc;
  if (nnz_out) *nnz_out = ctx->ms.nnz;
  return GTK_OK;
}

int32_t gtk_matrix_colptr_at(gtk_ctx* ctx, int32_t n, const int64_t* cols, int64_t* out) {
  if (!ctx) return GTK_ERR_INVALID;
  MatSym& m = ctx->ms;
  if (!m.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_colptr_at: no symbolic result");
  if (n < 0 || (n && (!cols || !out))) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_colptr_at: bad arguments");
  GTK_CK(cudaSetDevice(ctx->device));
  for (int i = 0; i < n; ++i) {
    if (cols[i] < 0 || cols[i] > m.n_cols) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_colptr_at: column out of range");
    GTK_CK(cudaMemcpyAsync(out + i, m.colptr + cols[i], sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_matrix_pattern(gtk_ctx* ctx, int32_t* colptr, int32_t* rowval) {
  if (!ctx) return GTK_ERR_INVALID;
  MatSym& m = ctx->ms;
  if (!m.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_pattern: no symbolic result");
  if (m.nnz >= 0x7FFFFFFFll) GTK_FAIL(GTK_ERR_TOO_LARGE, "nnz does not fit the Int32 colptr of SparseMatrixCSC{Float64,Int32}");
  GTK_CK(cudaSetDevice(ctx->device));
  if (colptr) {
    std::vector<int64_t> cp((size_t)m.n_cols + 1);
    GTK_CK(cudaMemcpyAsync(cp.data(), m.colptr, sizeof(int64_t) * cp.size(), cudaMemcpyDeviceToHost, ctx->stream));
    GTK_CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < cp.size(); ++i) colptr[i] = (int32_t)(cp[i] + 1);
  }
  if (rowval && m.nnz) {
    GTK_CK(cudaMemcpyAsync(rowval, m.rowval, sizeof(int32_t) * (size_t)m.nnz, cudaMemcpyDeviceToHost, ctx->stream));
    GTK_CK(cudaStreamSynchronize(ctx->stream));
  }
  return GTK_OK;
}

namespace {
__global__ void k_widen_rows(const int32_t* __restrict__ src, int64_t n, int64_t* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

int32_t gtk_matrix_pattern_i64(gtk_ctx* ctx, int64_t* colptr, int64_t* rowval) {
  if (!ctx) return GTK_ERR_INVALID;
  MatSym& m = ctx->ms;
  if (!m.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_pattern_i64: no symbolic result");
  GTK_CK(cudaSetDevice(ctx->device));
  if (colptr) {
    GTK_CK(cudaMemcpyAsync(colptr, m.colptr, sizeof(int64_t) * (size_t)(m.n_cols + 1), cudaMemcpyDeviceToHost, ctx->stream));
    GTK_CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i <= m.n_cols; ++i) colptr[i] += 1;   // 1-based like Julia
  }
  if (rowval && m.nnz) {   // widened on the device, chunk by chunk (row ids stay Int32 in HBM)
    const int64_t chunk = (int64_t)16 << 20;
    int64_t* tmp = nullptr;
    GTK_CK(gtk_cuda_malloc(ctx, &tmp, sizeof(int64_t) * (size_t)std::min(chunk, m.nnz)));
    for (int64_t i0 = 0; i0 < m.nnz; i0 += chunk) {
      const int64_t n = std::min(chunk, m.nnz - i0);
      k_widen_rows<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(m.rowval + i0, n, tmp);
      GTK_CK(cudaGetLastError());
      GTK_CK(cudaMemcpyAsync(rowval + i0, tmp, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
      GTK_CK(cudaStreamSynchronize(ctx->stream));
    }
    gtk_cuda_free(ctx, tmp);
  }
  return GTK_OK;
}

int32_t gtk_matrix_numeric_device(gtk_ctx* ctx, int32_t form, const gtk_form_params* p) {
  if (!ctx) return GTK_ERR_INVALID;
  GTK_CK(cudaSetDevice(ctx->device));
  return gtk_numeric_matrix_impl(ctx, form, p);
}

int32_t gtk_copy_nzval(gtk_ctx* ctx, double* nzval) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->ms.ready || !ctx->nzval) GTK_FAIL(GTK_ERR_STATE, "no assembled matrix");
  if (ctx->ms.nnz && nzval)
    GTK_CK(cudaMemcpyAsync(nzval, ctx->nzval, sizeof(double) * (size_t)ctx->ms.nnz, cudaMemcpyDeviceToHost, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_copy_vector(gtk_ctx* ctx, double* b) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->vs.ready || !ctx->bvec) GTK_FAIL(GTK_ERR_STATE, "no assembled vector");
  if (ctx->vs.n_rows && b)
    GTK_CK(cudaMemcpyAsync(b, ctx->bvec, sizeof(double) * (size_t)ctx->vs.n_rows, cudaMemcpyDeviceToHost, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_matrix_numeric(gtk_ctx* ctx, int32_t form, const gtk_form_params* p, double* nzval) {
  int32_t rc = gtk_matrix_numeric_device(ctx, form, p);
  if (rc) return rc;
  return gtk_copy_nzval(ctx, nzval);
}

int32_t gtk_vector_symbolic(gtk_ctx* ctx, int32_t fd) {
  if (!ctx) return GTK_ERR_INVALID;
  if (fd != GTK_FREE && fd != GTK_DIRICHLET) GTK_FAIL(GTK_ERR_INVALID, "free_or_dirichlet must be GTK_FREE or GTK_DIRICHLET");
  if (!ctx->cell_dofs) GTK_FAIL(GTK_ERR_STATE, "gtk_vector_symbolic: set mesh and space first");
  GTK_CK(cudaSetDevice(ctx->device));
  return gtk_symbolic_vector_impl(ctx, fd);
}

int32_t gtk_vector_assemble_device(gtk_ctx* ctx, int32_t form, const gtk_form_params* p) {
  if (!ctx) return GTK_ERR_INVALID;
  GTK_CK(cudaSetDevice(ctx->device));
  return gtk_numeric_vector_impl(ctx, form, p);
}

int32_t gtk_vector_assemble(gtk_ctx* ctx, int32_t form, const gtk_form_params* p, double* b) {
  int32_t rc = gtk_vector_assemble_device(ctx, form, p);
  if (rc) return rc;
  return gtk_copy_vector(ctx, b);
}

int32_t gtk_assemble_matrix_and_vector_device(gtk_ctx* ctx, int32_t mform, const gtk_form_params* pm,
                                              int32_t vform, const gtk_form_params* pv) {
  if (!ctx) return GTK_ERR_INVALID;
  GTK_CK(cudaSetDevice(ctx->device));
  return gtk_numeric_both_impl(ctx, mform, pm, vform, pv);
}

int32_t gtk_assemble_matrix_and_vector(gtk_ctx* ctx, int32_t mform, const gtk_form_params* pm, int32_t vform,
                                       const gtk_form_params* pv, double* nzval, double* b) {
  int32_t rc = gtk_assemble_matrix_and_vector_device(ctx, mform, pm, vform, pv);
  if (rc) return rc;
  if (ctx->ms.nnz && nzval)
    GTK_CK(cudaMemcpyAsync(nzval, ctx->nzval, sizeof(double) * (size_t)ctx->ms.nnz, cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->vs.n_rows && b)
    GTK_CK(cudaMemcpyAsync(b, ctx->bvec, sizeof(double) * (size_t)ctx->vs.n_rows, cudaMemcpyDeviceToHost, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_device_pointer(gtk_ctx* ctx, int32_t which, void** dptr, int64_t* count) {
  if (!ctx || !dptr) return GTK_ERR_INVALID;
  switch (which) {
    case 0: *dptr = ctx->nzval; if (count) *count = ctx->ms.nnz; break;
    case 1: *dptr = ctx->bvec; if (count) *count = ctx->vs.n_rows; break;
    case 2: *dptr = ctx->ms.colptr; if (count) *count = ctx->ms.n_cols + 1; break;
    case 3: *dptr = ctx->ms.rowval; if (count) *count = ctx->ms.nnz; break;
    case 4: { int32_t rc = gtk_field_ensure(ctx); if (rc) return rc; *dptr = ctx->u_free; if (count) *count = ctx->n_free; break; }
    case 5: { int32_t rc = gtk_field_ensure(ctx); if (rc) return rc; *dptr = ctx->u_diri; if (count) *count = ctx->n_diri; break; }
    case 6: *dptr = ctx->xdof_free; if (count) *count = ctx->xdof_free ? ctx->n_free * ctx->D : 0; break;
    case 7: *dptr = ctx->xdof_diri; if (count) *count = ctx->xdof_diri ? ctx->n_diri * ctx->D : 0; break;
    case 8: *dptr = ctx->xyz; if (count) *count = ctx->n_nodes * ctx->D; break;
    case 9: *dptr = ctx->cell_nodes; if (count) *count = ctx->n_cells * ctx->nln; break;
    case 10: *dptr = ctx->cell_dofs; if (count) *count = ctx->cell_dofs ? ctx->n_cells * ctx->nld : 0; break;
    default: GTK_FAIL(GTK_ERR_INVALID, "gtk_device_pointer: unknown selector");
  }
  return GTK_OK;
}

int32_t gtk_copy_device_array(gtk_ctx* ctx, int32_t which, void* host, int64_t bytes) {
  if (!ctx) return GTK_ERR_INVALID;
  void* d = nullptr; int64_t count = 0;
  int32_t rc = gtk_device_pointer(ctx, which, &d, &count);
  if (rc) return rc;
  const int64_t elem = (which == 2) ? 8 : ((which == 3 || which == 9 || which == 10) ? 4 : 8);
  if (bytes < 0 || bytes > count * elem || (bytes && (!host || !d))) GTK_FAIL(GTK_ERR_INVALID, "gtk_copy_device_array: bad size or array not present");
  GTK_CK(cudaSetDevice(ctx->device));
  if (bytes) GTK_CK(cudaMemcpyAsync(host, d, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  return GTK_OK;
}

int32_t gtk_set_profiling(gtk_ctx* ctx, int32_t on) {
  if (!ctx) return GTK_ERR_INVALID;
  ctx->profiling = on != 0;
  return GTK_OK;
}

int32_t gtk_profile_count(const gtk_ctx* ctx) { return ctx ? (int32_t)ctx->prof.size() : 0; }

int32_t gtk_profile_get(gtk_ctx* ctx, int32_t i, char* name64, double* ms) {
  if (!ctx) return GTK_ERR_INVALID;
  if (i < 0 || i >= (int32_t)ctx->prof.size()) GTK_FAIL(GTK_ERR_INVALID, "gtk_profile_get: index out of range");
  GTK_CK(cudaEventSynchronize(ctx->prof[i].b));
  float f = 0.f;
  GTK_CK(cudaEventElapsedTime(&f, ctx->prof[i].a, ctx->prof[i].b));
  if (ms) *ms = (double)f;
  if (name64) { strncpy(name64, ctx->prof[i].name, 63); name64[63] = 0; }
  return GTK_OK;
}

int64_t gtk_info(const gtk_ctx* ctx, int32_t key) {
  if (!ctx) return -1;
  switch (key) {
    case 0: return ctx->launches_last;
    case 1: return ctx->launches_total;
    case 2: return ctx->bytes_held;
    case 3: return ctx->ms.nnz;
    case 4: return ctx->ms.n_valid;
    case 5: return ctx->fast_path_last;
    default: return -1;
  }
}

}  // extern "C"
