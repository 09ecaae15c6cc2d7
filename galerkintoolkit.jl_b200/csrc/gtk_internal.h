// Internal engine state (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/gtk_assembly.h"

#define GTK_CK(call)                                                                      \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ +   \
                 ":" + std::to_string(__LINE__) + ")";                                    \
      return GTK_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define GTK_FAIL(code, msg) \
  do {                      \
    ctx->err = (msg);       \
    return (code);          \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;  // elements allocated
};

// Sparse-pattern + assembly plan of one (rows, cols) selection.
struct MatSym {
  bool ready = false;
  bool generic_plan = false;   // perm / nzptr built (sort-based plan); false after the structured symbolic phase
  int rows_fd = GTK_FREE, cols_fd = GTK_FREE;
  int64_t n_rows = 0, n_cols = 0;
  int64_t n_full = 0;    // n_cells * n_ldofs^2  (COO slots incl. skipped ones)
  int64_t n_valid = 0;   // N_coo: triplets the reference would push (assembly.jl:545-556)
  int64_t nnz = 0;
  int64_t* colptr = nullptr;   // [n_cols+1], 0-based, device
  int32_t* rowval = nullptr;   // [nnz], 1-based, device
  uint32_t* perm = nullptr;    // [n_valid] sorted position -> COO slot e = cell*nld^2 + c*nld + r
  uint32_t* nzptr = nullptr;   // [nnz+1]  first sorted position of every stored nonzero
  // Q1-hex structured fast path ("tile plan"), see plan.cu
  void* plan = nullptr;
  // direct-write plan of the element-GEMM path (elemgemm.cu), built on first use: where each COO slot's value goes
  bool direct_ready = false;
  uint32_t* dest = nullptr;      // [n_full] 0x80000000|p: the slot is the ONLY contribution of nonzero p -> nzval[p];
                                 //          0xFFFFFFFF: skipped slot; anything else: staged in KE[e] and reduced
  uint32_t* multi = nullptr;     // [n_multi] nonzeros with >= 2 contributions
  int64_t n_multi = 0;
  // row-major view for b = beta b + alpha A x (matvec.cu), built on first use
  bool csr_ready = false;
  int64_t* csr_ptr = nullptr;    // [n_rows+1]
  uint32_t* csr_pos = nullptr;   // [nnz] position in nzval, rows ascending, columns ascending inside a row
  int32_t* csr_col = nullptr;    // [nnz] 0-based column of that entry
};

// one assembled matrix kept next to the current one (gtk_select_matrix): pattern, plans and values
struct MatSlot {
  MatSym ms;
  double* nzval = nullptr;
  size_t nzval_cap = 0;
};

struct VecSym {
  bool ready = false;
  bool generic_plan = false;   // perm / rowptr / urow built
  int fd = GTK_FREE;
  int64_t n_rows = 0;
  int64_t n_full = 0;   // n_cells * n_ldofs
  int64_t n_valid = 0;
  int64_t n_urows = 0;  // rows that receive at least one contribution
  uint32_t* perm = nullptr;    // [n_valid] sorted position -> e = cell*nld + i
  uint32_t* rowptr = nullptr;  // [n_urows+1]
  int32_t* urow = nullptr;     // [n_urows] 0-based row id
};

// Ghost-row exchange fused into the persistent sweep kernel (fastq1.cu + comm.cu): the warps that drew all (patch,
// z-segment) items of the top layers have written what a peer waits for; PUSH items then gather those entries and store
// them into the owner's buffer over NVLink, UNPACK items add what the peers pushed — work items of the SAME kernel, ordered
// by device-side counters.  All counters only ever grow; the host passes this launch's targets.
struct GtkCommPeerDev {
  const int64_t* send_nz; const int32_t* send_rows; long long n_send_nz, n_send_b;
  double* remote_buf; unsigned long long* remote_ready; const unsigned long long* local_ack;
  const int64_t* recv_nz; const int32_t* recv_rows; long long n_recv_nz, n_recv_b;
  const double* recv_buf; const unsigned long long* local_ready; unsigned long long* remote_ack;
  unsigned long long seq;                       // exchange number with this peer
  unsigned long long push_target, unpack_target;   // counter values at which the last PUSH / UNPACK item of this launch finishes
  // mode 2 (exchange fused into the sweep's copy-out): per in-plane lattice node, where the ghost entries of its columns
  // sit in the send / receive buffer (-1: none).  tbl = [snd_off T-1][snd_off T][snd_boff][rcv_off B-1][rcv_off B][rcv_boff],
  // each s2 = (n1+1)(n2+1) ints; T = node layer of the rows sent to this peer, B = node layer of the rows received from it
  const int32_t* tbl;
  int T, B;
};
struct GtkCommDev {
  int on;                                       // 0: plain sweep, 1: PUSH / UNPACK work items, 2: exchange inside the copy-out
  int n_peers;
  int top_layer, bot_layer;                     // sweep items with kz0 >= top_layer feed the PUSH, with kz1 <= bot_layer the UNPACK
  unsigned long long* cnt;                      // [0] top items done [1] bottom items done [2+p] PUSH items done [4+p] UNPACK items done
  unsigned long long top_target, bot_target;
  unsigned long long* dbg;                      // optional [8]: ns spent waiting (GTK_COMM_TIMING), else nullptr
  int dbg_skip;                                 // experiments only: 1 skip the remote stores, 2 skip the receive-side work, 4 skip the fences
  GtkCommPeerDev peer[2];
};

struct gtk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 148;
  size_t smem_optin = 0;

  // mesh
  int D = 0, nln = 0;
  int dman = 0;          // dimension of the cells' reference space (= D for volume cells, D-1 for boundary faces)
  int64_t n_nodes = 0, n_cells = 0;
  double* xyz = nullptr;
  int32_t* cell_nodes = nullptr;
  int64_t act_first = 0, act_count = -1;   // numeric-active cell range (-1 = all)
  // space
  int nld = 0, ncomp = 1, nls = 0;
  int64_t n_free = 0, n_diri = 0;
  int32_t* cell_dofs = nullptr;
  // tabulation (device + host copies)
  int nq = 0;
  double *w = nullptr, *N = nullptr, *dN = nullptr, *M = nullptr, *dM = nullptr;
  std::vector<double> h_w, h_N, h_dN, h_M, h_dM;

  MatSym ms;            // the SELECTED matrix (gtk_select_matrix); the others wait in `slots`
  VecSym vs;
  static constexpr int N_SLOTS = 4;
  MatSlot slots[N_SLOTS];
  int cur_slot = 0;
  double* xvec = nullptr; size_t xvec_cap = 0;   // uploaded x of gtk_matvec_add
  // staging / results
  double* KE = nullptr;  size_t KE_cap = 0;   // [n_cells][nld][nld] element matrices, e-indexed
  double* BE = nullptr;  size_t BE_cap = 0;   // [n_cells][nld]
  double* nzval = nullptr; size_t nzval_cap = 0;
  double* bvec = nullptr;  size_t bvec_cap = 0;
  double* f_dev = nullptr; size_t f_cap = 0;  // uploaded f_nodal / f_qp
  double* coef_dev = nullptr; size_t coef_cap = 0;   // uploaded coef_nodal / coef_qp
  double* Cm = nullptr;  size_t Cm_cap = 0;   // [n_cells][n_q padded][6] per-point metric (elemgemm.cu)
  // DiscreteField parameter u_h of the current space (field.cu): free / Dirichlet values, zero until set
  double* u_free = nullptr;  size_t u_free_cap = 0;
  double* u_diri = nullptr;  size_t u_diri_cap = 0;
  double* xdof_free = nullptr; size_t xdof_free_cap = 0;   // [n_free][D] dof-node coordinates (gtk_space_dof_coordinates)
  double* xdof_diri = nullptr; size_t xdof_diri_cap = 0;   // [n_dirichlet][D]
  double* scal_part = nullptr; size_t scal_part_cap = 0;   // block partial sums of gtk_scalar_assemble

  struct { size_t xyz = 0, cell_nodes = 0, cell_dofs = 0, w = 0, N = 0, dN = 0, M = 0, dM = 0; } sz;  // uploaded element counts

  int64_t launches_last = 0, launches_total = 0;
  int64_t bytes_held = 0;
  int fast_path_last = 0;
  // sweep kernels: node-layer subset of the next launch (comm.cu overlap): 0 all, 1 layers >= seg_layer (hold what goes to
  // a peer), 3 layers < seg_lo (hold what a peer adds to), 2 the layers in between
  int seg_mode = 0, seg_layer = 0, seg_lo = 0;
  // fused exchange (comm.cu asks, fastq1.cu answers): want = run the exchange inside the sweep launch if it can; done = it did
  bool fuse_comm_want = false, fuse_comm_done = false;

  // per-kernel profiling (events around each launch of the last numeric call)
  bool profiling = false;
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;       // records of the last numeric call
  std::vector<ProfRec> prof_pool;  // reusable events

  // device-memory pool: freed blocks are kept and handed out again for requests of the same size, so that a repeated
  // symbolic phase (same mesh, new space / new selection) does not pay cudaMalloc/cudaFree (ms each, driver-dependent).
  // Everything the engine does is ordered on ctx->stream, so immediate reuse is safe.
  struct Pool {
    std::unordered_map<void*, size_t> live;
    std::multimap<size_t, void*> cached;
    size_t cached_bytes = 0;
    size_t cap_bytes = (size_t)8 << 30;
  } pool;

  // multi-GPU
  void* comm = nullptr;   // ncclComm_t
  int rank = 0, n_ranks = 1;
  void* ghost = nullptr;  // GhostPlan*
  void* sumplan = nullptr;   // SumPlan* (matsum.cu): this context holds the merged matrix of a sum of integrals
  void* parts = nullptr;  // PartsState* (blocks.cu): parts of a product space / skeleton integral; replaces N / dN
};

// ---- helpers implemented in gtk_api.cu ----
cudaError_t gtk_cuda_malloc(gtk_ctx* ctx, void** p, size_t bytes);   // pooled cudaMalloc (does not touch bytes_held)
template <class T>
inline cudaError_t gtk_cuda_malloc(gtk_ctx* ctx, T** p, size_t bytes) { return gtk_cuda_malloc(ctx, reinterpret_cast<void**>(p), bytes); }
void gtk_cuda_free(gtk_ctx* ctx, void* p);                            // back to the pool
void gtk_pool_flush(gtk_ctx* ctx);                                    // cudaFree every cached block
int32_t gtk_dev_alloc(gtk_ctx* ctx, void** p, size_t bytes);
void gtk_dev_free(gtk_ctx* ctx, void* p, size_t bytes);
template <class T>
inline int32_t gtk_alloc(gtk_ctx* ctx, T** p, size_t n) {
  return gtk_dev_alloc(ctx, reinterpret_cast<void**>(p), n * sizeof(T));
}
template <class T>
inline void gtk_free(gtk_ctx* ctx, T*& p, size_t n) {
  if (p) gtk_dev_free(ctx, p, n * sizeof(T));
  p = nullptr;
}
inline void gtk_count_launch(gtk_ctx* ctx, int n = 1) {
  ctx->launches_last += n;
  ctx->launches_total += n;
}
// Brackets one kernel launch with events when profiling is on:  { GtkProf p(ctx, "name"); kernel<<<>>>(); }
struct GtkProf {
  gtk_ctx* ctx; int idx = -1;
  GtkProf(gtk_ctx* c, const char* name) : ctx(c) {
    if (!c->profiling) return;
    gtk_ctx::ProfRec r;
    if (!c->prof_pool.empty()) { r = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
    r.name = name;
    cudaEventRecord(r.a, c->stream);
    c->prof.push_back(r);
    idx = (int)c->prof.size() - 1;
  }
  ~GtkProf() { if (idx >= 0) cudaEventRecord(ctx->prof[idx].b, ctx->stream); }
};
inline void gtk_prof_reset(gtk_ctx* ctx) {
  for (auto& r : ctx->prof) ctx->prof_pool.push_back(r);
  ctx->prof.clear();
}

// ---- symbolic.cu ----
int32_t gtk_symbolic_matrix_impl(gtk_ctx* ctx, int rows_fd, int cols_fd);
int32_t gtk_symbolic_vector_impl(gtk_ctx* ctx, int fd);
int32_t gtk_symbolic_generic_plan(gtk_ctx* ctx);          // matrix: sort-based pattern + reduction plan
int32_t gtk_symbolic_direct_plan(gtk_ctx* ctx);           // matrix: dest / multi on top of the generic plan
int32_t gtk_symbolic_vector_generic_plan(gtk_ctx* ctx);   // vector: same
void gtk_matsym_release(gtk_ctx* ctx);                   // the selected matrix: pattern, plans (values stay allocated)
void gtk_release_all_matrices(gtk_ctx* ctx);             // every slot (mesh / space changed)
int32_t gtk_select_matrix_impl(gtk_ctx* ctx, int slot);
void gtk_vecsym_release(gtk_ctx* ctx);

// ---- numeric.cu ----
int32_t gtk_numeric_matrix_impl(gtk_ctx* ctx, int form, const gtk_form_params* p);
int32_t gtk_numeric_vector_impl(gtk_ctx* ctx, int form, const gtk_form_params* p);
// coefficient of a bilinear form: uploads p->coef_nodal / p->coef_qp; *mode = 0 none, 1 nodal, 2 per quadrature point
int32_t gtk_upload_coefficient(gtk_ctx* ctx, int form, const gtk_form_params* p, int* mode);
int32_t gtk_numeric_both_impl(gtk_ctx* ctx, int mform, const gtk_form_params* pm, int vform,
                              const gtk_form_params* pv);
int32_t gtk_scalar_impl(gtk_ctx* ctx, int kind, const gtk_form_params* p, double* out);

// ---- matsum.cu ----
void gtk_sumplan_release(gtk_ctx* ctx);

// ---- blocks.cu ----
void gtk_parts_release(gtk_ctx* ctx);       // mesh / space / manifold dimension changed: the part tables are void

// ---- field.cu ----
int32_t gtk_field_ensure(gtk_ctx* ctx);     // allocates (zero-filled) u_free / u_diri for the current space if missing
void gtk_field_release(gtk_ctx* ctx);       // the space changed: drop values and dof coordinates
