// Symbolic phase: CSC sparsity pattern + deterministic reduction plan, on device.
//
// Replaces (i) the counting loop of allocate_matrix (assembly.jl:119-153), (ii) the COO
// index push of contribute! (assembly.jl:545-556) and (iii) the symbolic half of
// PartitionedArrays.sparse_matrix = Julia sparse(I,J,V,m,n) (assembly.jl:571-575).
//
// Every COO slot e = cell*nld^2 + c*nld + r (the reference's push order: cell-major, column c
// outer, row r inner, assembly.jl:195-207) gets the 64-bit key col*n_rows+row, or an INVALID key
// if the reference would skip it by sign (assembly.jl:155-157).  A stable LSD radix sort of
// (key, e) puts duplicates of one (row,col) next to each other *in reference push order*, so the
// later fixed-order segmented sum reproduces Julia's left-to-right duplicate combine and is
// bit-reproducible (no float atomics anywhere).
#include <cub/cub.cuh>
#include "gtk_internal.h"

namespace {

__device__ __forceinline__ bool skip_dof(int d, int fd) {  // assembly.jl:155-157
  return (d > 0 && fd == GTK_DIRICHLET) || (d < 0 && fd == GTK_FREE);
}

__global__ void k_matrix_keys(const int32_t* __restrict__ cell_dofs, int64_t n_full, int nld,
                              int rows_fd, int cols_fd, uint64_t n_rows, uint64_t invalid,
                              uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int nld2 = nld * nld;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_full;
       e += (int64_t)gridDim.x * blockDim.x) {
    int64_t cell = e / nld2;
    int rem = (int)(e - cell * nld2);
    int c = rem / nld, r = rem - c * nld;
    int dr = cell_dofs[cell * nld + r];
    int dc = cell_dofs[cell * nld + c];
    bool valid = !(skip_dof(dr, rows_fd) || skip_dof(dc, cols_fd)) && dr != 0 && dc != 0;
    uint64_t row = (uint64_t)(abs(dr) - 1), col = (uint64_t)(abs(dc) - 1);
    keys[e] = valid ? col * n_rows + row : invalid;
    vals[e] = (uint32_t)e;
  }
}

__global__ void k_vector_keys(const int32_t* __restrict__ cell_dofs, int64_t n_full, int fd,
                              uint32_t invalid, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_full;
       e += (int64_t)gridDim.x * blockDim.x) {
    int d = cell_dofs[e];
    bool valid = !skip_dof(d, fd) && d != 0;
    keys[e] = valid ? (uint32_t)(abs(d) - 1) : invalid;
    vals[e] = (uint32_t)e;
  }
}

// first index s with keys[s] >= invalid  (keys sorted)
template <class K>
__global__ void k_lower_bound(const K* __restrict__ keys, int64_t n, K invalid, int64_t* out) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < invalid) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}

template <class K>
struct HeadPred {
  const K* keys;
  __device__ __forceinline__ bool operator()(const uint32_t& s) const {
    return s == 0 || keys[s] != keys[s - 1];
  }
};

__global__ void k_pattern(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ nzptr,
                          int64_t nnz, uint64_t n_rows, int64_t n_cols,
                          int64_t* __restrict__ colptr, int32_t* __restrict__ rowval) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * blockDim.x) {
    uint64_t key = keys[nzptr[p]];
    int64_t col = (int64_t)(key / n_rows);
    rowval[p] = (int32_t)(key - (uint64_t)col * n_rows) + 1;
    int64_t prev = p > 0 ? (int64_t)(keys[nzptr[p - 1]] / n_rows) : -1;
    for (int64_t c = prev + 1; c <= col; ++c) colptr[c] = p;   // also fills empty columns
    if (p == nnz - 1)
      for (int64_t c = col + 1; c <= n_cols; ++c) colptr[c] = nnz;
  }
}

__global__ void k_urows(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rowptr,
                        int64_t n_urows, int32_t* __restrict__ urow) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n_urows;
       u += (int64_t)gridDim.x * blockDim.x)
    urow[u] = (int32_t)keys[rowptr[u]];
}

inline int grid_for(int64_t n, int block, int sm) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int bits_for(uint64_t maxval) {
  int b = 1;
  while (b < 64 && (maxval >> b) != 0) ++b;
  return b;
}

}  // namespace

void gtk_fastq1_release(gtk_ctx* ctx);   // fastq1.cu

void gtk_comm_release_plan(gtk_ctx* ctx);   // comm.cu

void gtk_matsym_release(gtk_ctx* ctx) {
  MatSym& m = ctx->ms;
  if (m.ready && ctx->cur_slot == 0) gtk_comm_release_plan(ctx);   // the ghost plan indexes the nzval of slot 0's pattern
  gtk_fastq1_release(ctx);
  gtk_free(ctx, m.dest, (size_t)m.n_full);
  gtk_free(ctx, m.multi, (size_t)m.n_multi);
  gtk_free(ctx, m.csr_ptr, (size_t)m.n_rows + 1);
  gtk_free(ctx, m.csr_pos, (size_t)m.nnz);
  gtk_free(ctx, m.csr_col, (size_t)m.nnz);
  gtk_free(ctx, m.colptr, (size_t)m.n_cols + 1);
  gtk_free(ctx, m.rowval, (size_t)m.nnz);
  gtk_free(ctx, m.perm, (size_t)m.n_valid);
  gtk_free(ctx, m.nzptr, (size_t)m.nnz + 1);
  m = MatSym();
}

void gtk_vecsym_release(gtk_ctx* ctx) {
  VecSym& v = ctx->vs;
  gtk_free(ctx, v.perm, (size_t)v.n_valid);
  gtk_free(ctx, v.rowptr, (size_t)v.n_urows + 1);
  gtk_free(ctx, v.urow, (size_t)v.n_urows);
  v = VecSym();
}

int32_t gtk_fastq1_symbolic(gtk_ctx* ctx, bool* handled);   // fastq1.cu

namespace {
struct MultiPred {
  const uint32_t* nzptr;
  __device__ __forceinline__ bool operator()(const uint32_t& p) const { return nzptr[p + 1] - nzptr[p] > 1u; }
};

__global__ void k_direct_dest(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ nzptr, int64_t nnz,
                              uint32_t* __restrict__ dest) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t s0 = nzptr[p], s1 = nzptr[p + 1];
    if (s1 - s0 == 1u) dest[perm[s0]] = 0x80000000u | (uint32_t)p;
    else for (uint32_t s = s0; s < s1; ++s) dest[perm[s]] = 0u;   // staged
  }
}
}  // namespace

// Where every COO slot's value goes when the producer can write results itself (elemgemm.cu): most nonzeros of a
// high-order matrix have exactly ONE contribution (Q3 hex: 82 %) — those skip the staging round trip entirely.
int32_t gtk_symbolic_direct_plan(gtk_ctx* ctx) {
  MatSym& m = ctx->ms;
  if (m.direct_ready) return GTK_OK;
  if (!m.generic_plan) { int32_t rc = gtk_symbolic_generic_plan(ctx); if (rc) return rc; }
  if (m.nnz >= (int64_t)0x7FFFFFFFll) GTK_FAIL(GTK_ERR_TOO_LARGE, "nnz exceeds the 31-bit nonzero id of the direct-write plan");
  cudaStream_t st = ctx->stream;
  int32_t rc;
  if ((rc = gtk_alloc(ctx, &m.dest, (size_t)(m.n_full > 0 ? m.n_full : 1)))) return rc;
  GTK_CK(cudaMemsetAsync(m.dest, 0xFF, sizeof(uint32_t) * (size_t)(m.n_full > 0 ? m.n_full : 1), st));
  m.n_multi = 0;
  if (m.nnz > 0) {
    k_direct_dest<<<grid_for(m.nnz, 256, ctx->sm_count), 256, 0, st>>>(m.perm, m.nzptr, m.nnz, m.dest);
    GTK_CK(cudaGetLastError());
    uint32_t* sel = nullptr;
    int64_t* d_n = nullptr;
    void* tmp = nullptr;
    auto cleanup = [&]() { gtk_cuda_free(ctx, sel); gtk_cuda_free(ctx, d_n); gtk_cuda_free(ctx, tmp); };
#define CKD(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return GTK_ERR_CUDA; } } while (0)
    CKD(gtk_cuda_malloc(ctx, &sel, sizeof(uint32_t) * (size_t)m.nnz));
    CKD(gtk_cuda_malloc(ctx, &d_n, sizeof(int64_t)));
    cub::CountingInputIterator<uint32_t> cnt(0);
    MultiPred pred{m.nzptr};
    size_t tb = 0;
    CKD(cub::DeviceSelect::If(nullptr, tb, cnt, sel, d_n, (int)m.nnz, pred, st));
    CKD(gtk_cuda_malloc(ctx, &tmp, tb));
    CKD(cub::DeviceSelect::If(tmp, tb, cnt, sel, d_n, (int)m.nnz, pred, st));
    int64_t nm = 0;
    CKD(cudaMemcpyAsync(&nm, d_n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CKD(cudaStreamSynchronize(st));
    m.n_multi = nm;
    if (nm > 0) {
      if ((rc = gtk_alloc(ctx, &m.multi, (size_t)nm))) { cleanup(); return rc; }
      CKD(cudaMemcpyAsync(m.multi, sel, sizeof(uint32_t) * (size_t)nm, cudaMemcpyDeviceToDevice, st));
      CKD(cudaStreamSynchronize(st));
    }
    cleanup();
#undef CKD
  }
  m.direct_ready = true;
  return GTK_OK;
}

int32_t gtk_symbolic_matrix_impl(gtk_ctx* ctx, int rows_fd, int cols_fd) {
  gtk_matsym_release(ctx);
  MatSym& m = ctx->ms;
  m.rows_fd = rows_fd;
  m.cols_fd = cols_fd;
  m.n_rows = rows_fd == GTK_FREE ? ctx->n_free : ctx->n_diri;
  m.n_cols = cols_fd == GTK_FREE ? ctx->n_free : ctx->n_diri;
  const int nld = ctx->nld;
  m.n_full = ctx->n_cells * (int64_t)nld * nld;
  cudaStream_t st = ctx->stream;
  GTK_CK(gtk_cuda_malloc(ctx, &m.colptr, sizeof(int64_t) * (size_t)(m.n_cols + 1)));
  ctx->bytes_held += sizeof(int64_t) * (m.n_cols + 1);
  GTK_CK(cudaMemsetAsync(m.colptr, 0, sizeof(int64_t) * (size_t)(m.n_cols + 1), st));
  if (rows_fd == GTK_FREE && cols_fd == GTK_FREE) {
    // structured Q1-hex meshes: pattern + sweep plan straight from the node lattice, generic plan deferred
    bool handled = false;
    int32_t rc = gtk_fastq1_symbolic(ctx, &handled);
    if (rc) return rc;
    if (handled) { m.ready = true; m.generic_plan = false; return GTK_OK; }
  }
  return gtk_symbolic_generic_plan(ctx);
}

// The sort-based symbolic phase: pattern + reduction plan (perm, nzptr) for the staged generic numeric kernels.  Also
// called lazily when the structured phase above produced the pattern and a form outside the sweep kernels is assembled
// (the pattern it rebuilds is the same one, by construction and by test).
int32_t gtk_symbolic_generic_plan(gtk_ctx* ctx) {
  MatSym& m = ctx->ms;
  const int rows_fd = m.rows_fd, cols_fd = m.cols_fd;
  const int nld = ctx->nld;
  if (m.n_full >= (int64_t)0xFFFFFFFFll)
    GTK_FAIL(GTK_ERR_TOO_LARGE, "n_cells*n_ldofs^2 exceeds the 32-bit COO slot index of this build");
  cudaStream_t st = ctx->stream;
  const int64_t n = m.n_full;
  const uint64_t invalid = (uint64_t)m.n_rows * (uint64_t)m.n_cols;
  if (m.rowval) { ctx->bytes_held -= sizeof(int32_t) * m.nnz; gtk_cuda_free(ctx, m.rowval); m.rowval = nullptr; }
  m.generic_plan = true;
  if (n == 0 || invalid == 0) {
    m.ready = true;
    GTK_CK(cudaStreamSynchronize(st));
    return GTK_OK;
  }

  uint64_t *k0 = nullptr, *k1 = nullptr;
  uint32_t *v0 = nullptr, *v1 = nullptr;
  void* tmp = nullptr;
  int64_t* d_scalar = nullptr;
  auto cleanup = [&]() {
    gtk_cuda_free(ctx, k0); gtk_cuda_free(ctx, k1); gtk_cuda_free(ctx, v0); gtk_cuda_free(ctx, v1); gtk_cuda_free(ctx, tmp); gtk_cuda_free(ctx, d_scalar);
  };
#define CKL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return GTK_ERR_CUDA; } } while (0)
  CKL(gtk_cuda_malloc(ctx, &k0, sizeof(uint64_t) * n));
  CKL(gtk_cuda_malloc(ctx, &k1, sizeof(uint64_t) * n));
  CKL(gtk_cuda_malloc(ctx, &v0, sizeof(uint32_t) * n));
  CKL(gtk_cuda_malloc(ctx, &v1, sizeof(uint32_t) * n));
  CKL(gtk_cuda_malloc(ctx, &d_scalar, sizeof(int64_t) * 2));

  k_matrix_keys<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(
      ctx->cell_dofs, n, nld, rows_fd, cols_fd, (uint64_t)m.n_rows, invalid, k0, v0);
  CKL(cudaGetLastError());

  cub::DoubleBuffer<uint64_t> dk(k0, k1);
  cub::DoubleBuffer<uint32_t> dv(v0, v1);
  size_t tmp_bytes = 0;
  const int end_bit = bits_for(invalid);
  CKL(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, end_bit, st));
  CKL(gtk_cuda_malloc(ctx, &tmp, tmp_bytes));
  CKL(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, n, 0, end_bit, st));
  const uint64_t* ks = dk.Current();
  const uint32_t* es = dv.Current();

  k_lower_bound<uint64_t><<<1, 1, 0, st>>>(ks, n, invalid, d_scalar);
  CKL(cudaGetLastError());
  int64_t n_valid = 0;
  CKL(cudaMemcpyAsync(&n_valid, d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CKL(cudaStreamSynchronize(st));
  m.n_valid = n_valid;

  if (n_valid > 0) {
    // segment heads -> nzptr, nnz
    uint32_t* heads = nullptr;   // worst case n_valid entries; shrunk afterwards
    CKL(gtk_cuda_malloc(ctx, &heads, sizeof(uint32_t) * (size_t)(n_valid + 1)));
    cub::CountingInputIterator<uint32_t> cnt(0);
    size_t tb2 = 0;
    HeadPred<uint64_t> pred{ks};
    int64_t* d_nsel = d_scalar + 1;
    gtk_cuda_free(ctx, tmp); tmp = nullptr;
    CKL(cub::DeviceSelect::If(nullptr, tb2, cnt, heads, d_nsel, n_valid, pred, st));
    CKL(gtk_cuda_malloc(ctx, &tmp, tb2));
    CKL(cub::DeviceSelect::If(tmp, tb2, cnt, heads, d_nsel, n_valid, pred, st));
    int64_t nnz = 0;
    CKL(cudaMemcpyAsync(&nnz, d_nsel, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CKL(cudaStreamSynchronize(st));
    m.nnz = nnz;
    CKL(gtk_cuda_malloc(ctx, &m.nzptr, sizeof(uint32_t) * (size_t)(nnz + 1)));
    CKL(cudaMemcpyAsync(m.nzptr, heads, sizeof(uint32_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
    uint32_t nv32 = (uint32_t)n_valid;
    CKL(cudaMemcpyAsync(m.nzptr + nnz, &nv32, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CKL(cudaStreamSynchronize(st));
    gtk_cuda_free(ctx, heads);
    CKL(gtk_cuda_malloc(ctx, &m.rowval, sizeof(int32_t) * (size_t)nnz));
    CKL(gtk_cuda_malloc(ctx, &m.perm, sizeof(uint32_t) * (size_t)n_valid));
    ctx->bytes_held += sizeof(uint32_t) * (nnz + 1) + sizeof(int32_t) * nnz + sizeof(uint32_t) * n_valid;
    k_pattern<<<grid_for(nnz, 256, ctx->sm_count), 256, 0, st>>>(ks, m.nzptr, nnz, (uint64_t)m.n_rows,
                                                                 m.n_cols, m.colptr, m.rowval);
    CKL(cudaGetLastError());
    CKL(cudaMemcpyAsync(m.perm, es, sizeof(uint32_t) * (size_t)n_valid, cudaMemcpyDeviceToDevice, st));
    CKL(cudaStreamSynchronize(st));
  }
  cleanup();
#undef CKL
  m.ready = true;
  return GTK_OK;
}

bool gtk_fastq1_plan_ok(const gtk_ctx* ctx);   // fastq1.cu

int32_t gtk_symbolic_vector_impl(gtk_ctx* ctx, int fd) {
  gtk_vecsym_release(ctx);
  VecSym& v = ctx->vs;
  v.fd = fd;
  v.n_rows = fd == GTK_FREE ? ctx->n_free : ctx->n_diri;
  v.n_full = ctx->n_cells * (int64_t)ctx->nld;
  if (v.n_rows >= (int64_t)0x7FFFFFFFll) GTK_FAIL(GTK_ERR_TOO_LARGE, "row count exceeds Int32");
  if (fd == GTK_FREE && gtk_fastq1_plan_ok(ctx)) {   // the sweep kernels write b themselves: generic plan deferred
    v.ready = true;
    v.generic_plan = false;
    return GTK_OK;
  }
  return gtk_symbolic_vector_generic_plan(ctx);
}

int32_t gtk_symbolic_vector_generic_plan(gtk_ctx* ctx) {
  VecSym& v = ctx->vs;
  const int fd = v.fd;
  if (v.n_full >= (int64_t)0xFFFFFFFFll)
    GTK_FAIL(GTK_ERR_TOO_LARGE, "n_cells*n_ldofs exceeds the 32-bit slot index of this build");
  cudaStream_t st = ctx->stream;
  const int64_t n = v.n_full;
  v.generic_plan = true;
  if (n == 0 || v.n_rows == 0) { v.ready = true; return GTK_OK; }
  const uint32_t invalid = (uint32_t)v.n_rows;
  uint32_t *k0 = nullptr, *k1 = nullptr, *v0 = nullptr, *v1 = nullptr, *heads = nullptr;
  void* tmp = nullptr;
  int64_t* d_scalar = nullptr;
  auto cleanup = [&]() {
    gtk_cuda_free(ctx, k0); gtk_cuda_free(ctx, k1); gtk_cuda_free(ctx, v0); gtk_cuda_free(ctx, v1); gtk_cuda_free(ctx, tmp); gtk_cuda_free(ctx, d_scalar); gtk_cuda_free(ctx, heads);
  };
#define CKL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return GTK_ERR_CUDA; } } while (0)
  CKL(gtk_cuda_malloc(ctx, &k0, sizeof(uint32_t) * n));
  CKL(gtk_cuda_malloc(ctx, &k1, sizeof(uint32_t) * n));
  CKL(gtk_cuda_malloc(ctx, &v0, sizeof(uint32_t) * n));
  CKL(gtk_cuda_malloc(ctx, &v1, sizeof(uint32_t) * n));
  CKL(gtk_cuda_malloc(ctx, &d_scalar, sizeof(int64_t) * 2));
  k_vector_keys<<<grid_for(n, 256, ctx->sm_count), 256, 0, st>>>(ctx->cell_dofs, n, fd, invalid, k0, v0);
  CKL(cudaGetLastError());
  cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
  size_t tb = 0;
  const int end_bit = bits_for(invalid);
  CKL(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, end_bit, st));
  CKL(gtk_cuda_malloc(ctx, &tmp, tb));
  CKL(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, n, 0, end_bit, st));
  const uint32_t* ks = dk.Current();
  const uint32_t* es = dv.Current();
  k_lower_bound<uint32_t><<<1, 1, 0, st>>>(ks, n, invalid, d_scalar);
  CKL(cudaGetLastError());
  int64_t n_valid = 0;
  CKL(cudaMemcpyAsync(&n_valid, d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CKL(cudaStreamSynchronize(st));
  v.n_valid = n_valid;
  if (n_valid > 0) {
    CKL(gtk_cuda_malloc(ctx, &heads, sizeof(uint32_t) * (size_t)(n_valid + 1)));
    cub::CountingInputIterator<uint32_t> cnt(0);
    HeadPred<uint32_t> pred{ks};
    size_t tb2 = 0;
    int64_t* d_nsel = d_scalar + 1;
    gtk_cuda_free(ctx, tmp); tmp = nullptr;
    CKL(cub::DeviceSelect::If(nullptr, tb2, cnt, heads, d_nsel, n_valid, pred, st));
    CKL(gtk_cuda_malloc(ctx, &tmp, tb2));
    CKL(cub::DeviceSelect::If(tmp, tb2, cnt, heads, d_nsel, n_valid, pred, st));
    int64_t nu = 0;
    CKL(cudaMemcpyAsync(&nu, d_nsel, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CKL(cudaStreamSynchronize(st));
    v.n_urows = nu;
    CKL(gtk_cuda_malloc(ctx, &v.rowptr, sizeof(uint32_t) * (size_t)(nu + 1)));
    CKL(gtk_cuda_malloc(ctx, &v.urow, sizeof(int32_t) * (size_t)nu));
    CKL(gtk_cuda_malloc(ctx, &v.perm, sizeof(uint32_t) * (size_t)n_valid));
    ctx->bytes_held += sizeof(uint32_t) * (nu + 1) + sizeof(int32_t) * nu + sizeof(uint32_t) * n_valid;
    CKL(cudaMemcpyAsync(v.rowptr, heads, sizeof(uint32_t) * (size_t)nu, cudaMemcpyDeviceToDevice, st));
    uint32_t nv32 = (uint32_t)n_valid;
    CKL(cudaMemcpyAsync(v.rowptr + nu, &nv32, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    k_urows<<<grid_for(nu, 256, ctx->sm_count), 256, 0, st>>>(ks, v.rowptr, nu, v.urow);
    CKL(cudaGetLastError());
    CKL(cudaMemcpyAsync(v.perm, es, sizeof(uint32_t) * (size_t)n_valid, cudaMemcpyDeviceToDevice, st));
    CKL(cudaStreamSynchronize(st));
  }
  cleanup();
#undef CKL
  v.ready = true;
  return GTK_OK;
}
