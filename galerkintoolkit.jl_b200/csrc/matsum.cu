// Sums of integrals over DIFFERENT domains in one matrix:  a(u,v) = ∫_Ω … dΩ + ∫_Γ … dΓ + ∫_Λ … dΛ.
//
// The reference loops over the contributions of the form and lets every one of them push into the SAME COO allocation
// (problems.jl:319-350: `for contribution in contributions(a(u,v))` around generate_assemble_matrix; one compress at
// the end), so the result has the union pattern and, per stored entry, the left-to-right sum of all triplets.
// The engine assembles one integral per context (each has its own integration faces); this file merges the assembled
// matrices on the device:
//   symbolic  keys col * n_rows + row of every source pattern, concatenated -> CUB radix sort (key, index) -> heads ->
//             union colptr / rowval and, per source, the position of each of its nonzeros in the union;
//   numeric   nzval_union = 0, then for the sources in the order of the sum: nzval_union[map_k[p]] += nzval_k[p]
//             (a source holds every (row, column) once, so no two threads of a pass touch the same entry: no atomics,
//             bit-reproducible).
// Order of the floating-point sum per entry: (Σ triplets of integral 1) + (Σ of integral 2) + … — the reference adds the
// triplets of integral 2 one by one to the running sum instead; the difference is O(1e-16) relative (inside the 1e-12 gate),
// the pattern is identical.
#include <cub/cub.cuh>
#include <vector>
#include "gtk_internal.h"

void gtk_matsym_release(gtk_ctx* ctx);

namespace {

struct SumPlan {
  int n = 0;
  std::vector<int64_t> nnz;        // per source
  std::vector<uint32_t*> map;      // per source: position of its p-th nonzero in the union
};

inline SumPlan* plan_of(gtk_ctx* ctx) { return static_cast<SumPlan*>(ctx->sumplan); }

inline int grid_for(int64_t n, int block, int sm) {
  int64_t g = (n + block - 1) / block, cap = (int64_t)sm * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// keys of one source pattern: thread per column
__global__ void k_sum_keys(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowval, int64_t n_cols, uint64_t n_rows,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t idx0) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x)
    for (int64_t p = colptr[c]; p < colptr[c + 1]; ++p) {
      keys[p] = (uint64_t)c * n_rows + (uint64_t)(rowval[p] - 1);
      idx[p] = idx0 + (uint32_t)p;
    }
}

__global__ void k_sum_heads(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ head) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    head[s] = (s == 0 || keys[s] != keys[s - 1]) ? 1u : 0u;
}

// pos[s] = inclusive scan of head - 1: the union position of sorted entry s
__global__ void k_sum_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head,
                              const uint32_t* __restrict__ incl, int64_t n, uint64_t n_rows, uint32_t* __restrict__ map_all,
                              uint64_t* __restrict__ ukeys, int32_t* __restrict__ rowval) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t pos = incl[s] - 1u;
    map_all[idx[s]] = pos;
    if (head[s]) {
      ukeys[pos] = keys[s];
      rowval[pos] = (int32_t)(keys[s] % n_rows) + 1;
    }
  }
}

// colptr[c] = first union position whose key >= c * n_rows
__global__ void k_sum_colptr(const uint64_t* __restrict__ ukeys, int64_t nnz, int64_t n_cols, uint64_t n_rows, int64_t* __restrict__ colptr) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c <= n_cols; c += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t want = (uint64_t)c * n_rows;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (ukeys[mid] < want) lo = mid + 1; else hi = mid;
    }
    colptr[c] = lo;
  }
}

__global__ void k_sum_add(const double* __restrict__ src, const uint32_t* __restrict__ map, int64_t n, double* __restrict__ dst) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) dst[map[p]] += src[p];
}

int32_t check_sources(gtk_ctx* ctx, int32_t n, gtk_ctx** src) {
  if (n < 1 || n > 16 || !src) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_sum: 1..16 source contexts");
  for (int k = 0; k < n; ++k) {
    if (!src[k] || src[k] == ctx) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_sum: null source, or the destination among the sources");
    if (src[k]->device != ctx->device) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_sum: all contexts must live on the same GPU");
    if (!src[k]->ms.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_sum: source " + std::to_string(k) + " has no pattern (gtk_matrix_symbolic)");
    if (src[k]->ms.n_rows != src[0]->ms.n_rows || src[k]->ms.n_cols != src[0]->ms.n_cols)
      GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_sum: the sources are matrices of different sizes");
  }
  return GTK_OK;
}

}  // namespace

void gtk_sumplan_release(gtk_ctx* ctx) {
  SumPlan* sp = plan_of(ctx);
  if (!sp) return;
  for (int k = 0; k < sp->n; ++k)
    if (sp->map[k]) gtk_dev_free(ctx, sp->map[k], sizeof(uint32_t) * (size_t)(sp->nnz[k] > 0 ? sp->nnz[k] : 1));
  delete sp;
  ctx->sumplan = nullptr;
}

extern "C" int32_t gtk_matrix_sum_symbolic(gtk_ctx* ctx, int32_t n, gtk_ctx** src, int64_t* nnz_out) {
  if (!ctx) return GTK_ERR_INVALID;
  int32_t rc = check_sources(ctx, n, src);
  if (rc) return rc;
  GTK_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (int k = 0; k < n; ++k) GTK_CK(cudaStreamSynchronize(src[k]->stream));   // their patterns were built on their own streams
  gtk_sumplan_release(ctx);
  gtk_matsym_release(ctx);
  const int64_t n_rows = src[0]->ms.n_rows, n_cols = src[0]->ms.n_cols;
  int64_t total = 0;
  for (int k = 0; k < n; ++k) total += src[k]->ms.nnz;
  if (total >= (int64_t)0x7FFFFFFFll) GTK_FAIL(GTK_ERR_TOO_LARGE, "gtk_matrix_sum: more than 2^31 source nonzeros");
  SumPlan* sp = new SumPlan();
  ctx->sumplan = sp;
  sp->n = n;
  sp->nnz.resize(n); sp->map.assign(n, nullptr);
  MatSym& m = ctx->ms;
  m.rows_fd = src[0]->ms.rows_fd; m.cols_fd = src[0]->ms.cols_fd;
  m.n_rows = n_rows; m.n_cols = n_cols;
  const size_t T = (size_t)(total > 0 ? total : 1);
  uint64_t *keys = nullptr, *keys2 = nullptr, *ukeys = nullptr;
  uint32_t *idx = nullptr, *idx2 = nullptr, *head = nullptr, *incl = nullptr, *map_all = nullptr;
  void* tmp = nullptr;
  int32_t* rowval_tmp = nullptr;
  auto cleanup = [&]() {
    for (void* q : {(void*)keys, (void*)keys2, (void*)ukeys, (void*)idx, (void*)idx2, (void*)head, (void*)incl, (void*)map_all, tmp, (void*)rowval_tmp})
      gtk_cuda_free(ctx, q);
  };
#define CKS(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return GTK_ERR_CUDA; } } while (0)
  CKS(gtk_cuda_malloc(ctx, &keys, sizeof(uint64_t) * T));
  CKS(gtk_cuda_malloc(ctx, &keys2, sizeof(uint64_t) * T));
  CKS(gtk_cuda_malloc(ctx, &ukeys, sizeof(uint64_t) * T));
  CKS(gtk_cuda_malloc(ctx, &idx, sizeof(uint32_t) * T));
  CKS(gtk_cuda_malloc(ctx, &idx2, sizeof(uint32_t) * T));
  CKS(gtk_cuda_malloc(ctx, &head, sizeof(uint32_t) * T));
  CKS(gtk_cuda_malloc(ctx, &incl, sizeof(uint32_t) * T));
  CKS(gtk_cuda_malloc(ctx, &map_all, sizeof(uint32_t) * T));
  CKS(gtk_cuda_malloc(ctx, &rowval_tmp, sizeof(int32_t) * T));
  int64_t off = 0;
  for (int k = 0; k < n; ++k) {
    const MatSym& s = src[k]->ms;
    sp->nnz[k] = s.nnz;
    if (s.nnz > 0) {
      k_sum_keys<<<grid_for(n_cols, 128, ctx->sm_count), 128, 0, st>>>(s.colptr, s.rowval, n_cols, (uint64_t)n_rows, keys + off, idx + off, (uint32_t)off);
      CKS(cudaGetLastError());
    }
    off += s.nnz;
  }
  int64_t nnz = 0;
  if (total > 0) {
    int end_bit = 1;
    { const uint64_t maxkey = (uint64_t)n_cols * (uint64_t)n_rows; while (end_bit < 64 && (maxkey >> end_bit) != 0) ++end_bit; }
    size_t tb = 0, tb2 = 0;
    CKS(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, idx, idx2, (int)total, 0, end_bit, st));
    CKS(cub::DeviceScan::InclusiveSum(nullptr, tb2, head, incl, (int)total, st));
    tb = tb > tb2 ? tb : tb2;
    CKS(gtk_cuda_malloc(ctx, &tmp, tb));
    CKS(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, idx, idx2, (int)total, 0, end_bit, st));
    k_sum_heads<<<grid_for(total, 256, ctx->sm_count), 256, 0, st>>>(keys2, total, head);
    CKS(cudaGetLastError());
    CKS(cub::DeviceScan::InclusiveSum(tmp, tb, head, incl, (int)total, st));
    uint32_t last = 0;
    CKS(cudaMemcpyAsync(&last, incl + (total - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CKS(cudaStreamSynchronize(st));
    nnz = (int64_t)last;
    k_sum_scatter<<<grid_for(total, 256, ctx->sm_count), 256, 0, st>>>(keys2, idx2, head, incl, total, (uint64_t)n_rows, map_all, ukeys, rowval_tmp);
    CKS(cudaGetLastError());
  }
  m.nnz = nnz;
  if ((rc = gtk_alloc(ctx, &m.colptr, (size_t)n_cols + 1))) { cleanup(); return rc; }
  if ((rc = gtk_alloc(ctx, &m.rowval, (size_t)(nnz > 0 ? nnz : 1)))) { cleanup(); return rc; }
  k_sum_colptr<<<grid_for(n_cols + 1, 256, ctx->sm_count), 256, 0, st>>>(ukeys, nnz, n_cols, (uint64_t)n_rows, m.colptr);
  CKS(cudaGetLastError());
  if (nnz > 0) CKS(cudaMemcpyAsync(m.rowval, rowval_tmp, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
  off = 0;
  for (int k = 0; k < n; ++k) {
    const size_t nk = (size_t)(sp->nnz[k] > 0 ? sp->nnz[k] : 1);
    if ((rc = gtk_dev_alloc(ctx, (void**)&sp->map[k], sizeof(uint32_t) * nk))) { cleanup(); return rc; }
    if (sp->nnz[k] > 0) CKS(cudaMemcpyAsync(sp->map[k], map_all + off, sizeof(uint32_t) * (size_t)sp->nnz[k], cudaMemcpyDeviceToDevice, st));
    off += sp->nnz[k];
  }
  CKS(cudaStreamSynchronize(st));
  cleanup();
#undef CKS
  m.ready = true;
  m.generic_plan = false;
  m.n_full = 0; m.n_valid = total;
  if (nnz_out) *nnz_out = nnz;
  return GTK_OK;
}

extern "C" int32_t gtk_matrix_sum_numeric_device(gtk_ctx* ctx, int32_t n, gtk_ctx** src) {
  if (!ctx) return GTK_ERR_INVALID;
  SumPlan* sp = plan_of(ctx);
  if (!sp || !ctx->ms.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_sum_symbolic must be called first");
  int32_t rc = check_sources(ctx, n, src);
  if (rc) return rc;
  if (n != sp->n) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_sum_numeric: not the sources of the symbolic call");
  for (int k = 0; k < n; ++k)
    if (src[k]->ms.nnz != sp->nnz[k] || !src[k]->nzval) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_sum_numeric: source " + std::to_string(k) + " changed its pattern or holds no values");
  GTK_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (int k = 0; k < n; ++k)
    if (src[k]->stream != st) GTK_CK(cudaStreamSynchronize(src[k]->stream));   // their values were assembled on their own streams
  const size_t nz = (size_t)(ctx->ms.nnz > 0 ? ctx->ms.nnz : 1);
  if (ctx->nzval_cap < nz || !ctx->nzval) {
    if (ctx->nzval) gtk_dev_free(ctx, ctx->nzval, ctx->nzval_cap * sizeof(double));
    ctx->nzval = nullptr; ctx->nzval_cap = 0;
    if ((rc = gtk_dev_alloc(ctx, (void**)&ctx->nzval, nz * sizeof(double)))) return rc;
    ctx->nzval_cap = nz;
  }
  ctx->launches_last = 0;
  gtk_prof_reset(ctx);
  GTK_CK(cudaMemsetAsync(ctx->nzval, 0, nz * sizeof(double), st));
  for (int k = 0; k < n; ++k) {
    if (sp->nnz[k] == 0) continue;
    { GtkProf pr_(ctx, "k_sum_add"); k_sum_add<<<grid_for(sp->nnz[k], 256, ctx->sm_count), 256, 0, st>>>(src[k]->nzval, sp->map[k], sp->nnz[k], ctx->nzval); }
    GTK_CK(cudaGetLastError());
    gtk_count_launch(ctx);
  }
  return GTK_OK;
}

extern "C" int32_t gtk_matrix_sum_numeric(gtk_ctx* ctx, int32_t n, gtk_ctx** src, double* nzval) {
  int32_t rc = gtk_matrix_sum_numeric_device(ctx, n, src);
  if (rc) return rc;
  if (!nzval) GTK_FAIL(GTK_ERR_INVALID, "gtk_matrix_sum_numeric: nzval is null");
  return gtk_copy_nzval(ctx, nzval);
}
