// Matrix slots and b = beta b + alpha A x on device.
//
// What it replaces: the linear-problem right-hand side of the reference,
//   A, Ad, b = assemble_matrix_and_vector_with_free_and_dirichlet_columns(...)   problems.jl:413-430
//   mul!(b, Ad, xd, -1, 1)                                                       problems.jl:447
// The reference assembles A (free x free) and Ad (free x Dirichlet) as two matrices that live side by side
// (problems.jl:363-380), so the engine keeps up to four assembled matrices: gtk_select_matrix(slot) makes one of them
// the target of gtk_matrix_symbolic / gtk_matrix_numeric / gtk_matrix_pattern; each keeps its pattern, plans and values
// for later update_matrix! calls.
//
// gtk_matvec_add is Julia's 5-argument mul!(C, A::SparseMatrixCSC, B, alpha, beta) (SparseArrays `_spmatmul!`):
//   C .*= beta (skipped for beta == 1, zero-fill for beta == 0);  for each column k: axk = B[k]*alpha;
//   for each stored entry j of the column: C[rowval[j]] += nzval[j]*axk
// so every C[i] receives its terms in increasing column order with separately rounded multiply and add.  The kernel
// reproduces exactly that order and rounding (no FMA contraction, no atomics): one thread per row walks the row's
// entries through a row-major index built once per pattern (stable sort of the nz positions by row).
#include <cub/cub.cuh>
#include "gtk_internal.h"

namespace {

__global__ void k_expand_cols(const int64_t* __restrict__ colptr, int64_t n_cols, int64_t nnz, const int32_t* __restrict__ rowval,
                              int32_t* __restrict__ keys, uint32_t* __restrict__ pos, int32_t* __restrict__ col) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n_cols;   // largest c with colptr[c] <= p
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (colptr[mid] <= p) lo = mid; else hi = mid;
    }
    keys[p] = rowval[p] - 1;
    pos[p] = (uint32_t)p;
    col[p] = (int32_t)lo;
  }
}

__global__ void k_gather_cols(const uint32_t* __restrict__ pos, const int32_t* __restrict__ col_by_pos, int64_t nnz,
                              int32_t* __restrict__ col_sorted) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nnz; s += (int64_t)gridDim.x * blockDim.x)
    col_sorted[s] = col_by_pos[pos[s]];
}

// csr_ptr[i] = first sorted position with row >= i
__global__ void k_row_ptr(const int32_t* __restrict__ rows_sorted, int64_t nnz, int64_t n_rows, int64_t* __restrict__ ptr) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (rows_sorted[mid] < i) lo = mid + 1; else hi = mid;
    }
    ptr[i] = lo;
  }
}

__global__ void k_matvec_add(const int64_t* __restrict__ ptr, const uint32_t* __restrict__ pos, const int32_t* __restrict__ col,
                             const double* __restrict__ nzval, const double* __restrict__ x, double alpha, double beta,
                             int64_t n_rows, double* __restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = beta == 1.0 ? b[i] : (beta == 0.0 ? 0.0 : __dmul_rn(b[i], beta));
    for (int64_t s = ptr[i]; s < ptr[i + 1]; ++s)
      acc = __dadd_rn(acc, __dmul_rn(nzval[pos[s]], __dmul_rn(x[col[s]], alpha)));
    b[i] = acc;
  }
}

inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

int32_t build_csr(gtk_ctx* ctx) {
  MatSym& m = ctx->ms;
  if (m.csr_ready) return GTK_OK;
  if (m.nnz >= (int64_t)0x7FFFFFFFll) GTK_FAIL(GTK_ERR_TOO_LARGE, "nnz exceeds the 31-bit position index of the row-major view");
  cudaStream_t st = ctx->stream;
  int32_t rc;
  if ((rc = gtk_alloc(ctx, &m.csr_ptr, (size_t)m.n_rows + 1))) return rc;
  if (m.nnz == 0) {
    GTK_CK(cudaMemsetAsync(m.csr_ptr, 0, sizeof(int64_t) * (size_t)(m.n_rows + 1), st));
    m.csr_ready = true;
    return GTK_OK;
  }
  if ((rc = gtk_alloc(ctx, &m.csr_pos, (size_t)m.nnz))) return rc;
  if ((rc = gtk_alloc(ctx, &m.csr_col, (size_t)m.nnz))) return rc;
  int32_t *k0 = nullptr, *k1 = nullptr, *colp = nullptr;
  uint32_t* v0 = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { gtk_cuda_free(ctx, k0); gtk_cuda_free(ctx, k1); gtk_cuda_free(ctx, colp); gtk_cuda_free(ctx, v0); gtk_cuda_free(ctx, tmp); };
#define CKM(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return GTK_ERR_CUDA; } } while (0)
  CKM(gtk_cuda_malloc(ctx, &k0, sizeof(int32_t) * (size_t)m.nnz));
  CKM(gtk_cuda_malloc(ctx, &k1, sizeof(int32_t) * (size_t)m.nnz));
  CKM(gtk_cuda_malloc(ctx, &colp, sizeof(int32_t) * (size_t)m.nnz));
  CKM(gtk_cuda_malloc(ctx, &v0, sizeof(uint32_t) * (size_t)m.nnz));
  k_expand_cols<<<grid_for(m.nnz, 256), 256, 0, st>>>(m.colptr, m.n_cols, m.nnz, m.rowval, k0, v0, colp);
  CKM(cudaGetLastError());
  int bits = 1;
  while (bits < 32 && ((uint64_t)m.n_rows >> bits) != 0) ++bits;
  size_t tb = 0;
  // stable LSD radix sort by row: inside a row the entries keep their CSC order = increasing column
  CKM(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, m.csr_pos, (int)m.nnz, 0, bits, st));
  CKM(gtk_cuda_malloc(ctx, &tmp, tb));
  CKM(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, m.csr_pos, (int)m.nnz, 0, bits, st));
  k_gather_cols<<<grid_for(m.nnz, 256), 256, 0, st>>>(m.csr_pos, colp, m.nnz, m.csr_col);
  k_row_ptr<<<grid_for(m.n_rows + 1, 256), 256, 0, st>>>(k1, m.nnz, m.n_rows, m.csr_ptr);
  CKM(cudaGetLastError());
  CKM(cudaStreamSynchronize(st));
  cleanup();
#undef CKM
  m.csr_ready = true;
  return GTK_OK;
}

}  // namespace

int32_t gtk_select_matrix_impl(gtk_ctx* ctx, int slot) {
  if (slot < 0 || slot >= gtk_ctx::N_SLOTS) GTK_FAIL(GTK_ERR_INVALID, "gtk_select_matrix: slot must be 0..3");
  if (slot == ctx->cur_slot) return GTK_OK;
  MatSlot& out = ctx->slots[ctx->cur_slot];
  out.ms = ctx->ms; out.nzval = ctx->nzval; out.nzval_cap = ctx->nzval_cap;
  MatSlot& in = ctx->slots[slot];
  ctx->ms = in.ms; ctx->nzval = in.nzval; ctx->nzval_cap = in.nzval_cap;
  in = MatSlot();
  ctx->cur_slot = slot;
  return GTK_OK;
}

void gtk_release_all_matrices(gtk_ctx* ctx) {
  const int keep = ctx->cur_slot;
  for (int s = 0; s < gtk_ctx::N_SLOTS; ++s) {
    gtk_select_matrix_impl(ctx, s);
    gtk_matsym_release(ctx);
  }
  gtk_select_matrix_impl(ctx, keep);
}

extern "C" {

int32_t gtk_select_matrix(gtk_ctx* ctx, int32_t slot) {
  if (!ctx) return GTK_ERR_INVALID;
  return gtk_select_matrix_impl(ctx, slot);
}

int32_t gtk_matvec_add_device(gtk_ctx* ctx, double alpha, const double* x, double beta) {
  if (!ctx) return GTK_ERR_INVALID;
  MatSym& m = ctx->ms;
  if (!m.ready || (m.nnz && !ctx->nzval)) GTK_FAIL(GTK_ERR_STATE, "gtk_matvec_add: the selected matrix is not assembled");
  if (!ctx->vs.ready || !ctx->bvec || ctx->vs.n_rows != m.n_rows)
    GTK_FAIL(GTK_ERR_STATE, "gtk_matvec_add: assemble a vector with the matrix's row selection first");
  if (!x && m.n_cols) GTK_FAIL(GTK_ERR_INVALID, "gtk_matvec_add: x is null");
  GTK_CK(cudaSetDevice(ctx->device));
  int32_t rc = build_csr(ctx);
  if (rc) return rc;
  const size_t nx = (size_t)(m.n_cols > 0 ? m.n_cols : 1);
  if (ctx->xvec_cap < nx) {
    if (ctx->xvec) gtk_dev_free(ctx, ctx->xvec, ctx->xvec_cap * sizeof(double));
    ctx->xvec = nullptr; ctx->xvec_cap = 0;
    if ((rc = gtk_dev_alloc(ctx, (void**)&ctx->xvec, nx * sizeof(double)))) return rc;
    ctx->xvec_cap = nx;
  }
  if (m.n_cols) GTK_CK(cudaMemcpyAsync(ctx->xvec, x, sizeof(double) * (size_t)m.n_cols, cudaMemcpyHostToDevice, ctx->stream));
  if (m.n_rows) {
    GtkProf pr_(ctx, "k_matvec_add");
    k_matvec_add<<<grid_for(m.n_rows, 256), 256, 0, ctx->stream>>>(m.csr_ptr, m.csr_pos, m.csr_col, ctx->nzval, ctx->xvec, alpha, beta,
                                                                 m.n_rows, ctx->bvec);
  }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  GTK_CK(cudaStreamSynchronize(ctx->stream));   // x is a borrowed host pointer
  return GTK_OK;
}

int32_t gtk_matvec_add(gtk_ctx* ctx, double alpha, const double* x, double beta, double* b) {
  int32_t rc = gtk_matvec_add_device(ctx, alpha, x, beta);
  if (rc) return rc;
  return gtk_copy_vector(ctx, b);
}

}  // extern "C"
