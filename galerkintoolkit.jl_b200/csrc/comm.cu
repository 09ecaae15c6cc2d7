// Multi-GPU data path: ghost-row summation over NCCL.
//
// One process per GPU.  Every cell is assembled exactly once, by the rank that owns it; rows whose
// owner is another rank ("ghost rows", PartitionedArrays vocabulary) hold partial sums that are sent
// to their owner and added there — PartitionedArrays.assemble! semantics.  The reference has no
// distributed assembly (docs/src/manual/introduction.md:22-28; space.jl:2776-2827 is commented out),
// so this mirrors the data model of docs/src/src_jl/manual_mesh_partitioning.jl:14-35 instead.
//
// The HOST computes the exchange plan (which nzval / b entries go to which peer, and where the
// received values are added) — galerkintoolkit.jl_b200/partition.py, testable on CPU over gloo.
// This file only moves bytes, two ways:
//  * peer-memory path (default once the host has exchanged the IPC handles, gtk_comm_p2p_export / _import):
//    ONE kernel per peer gathers the ghost entries and stores them straight into the owner's receive buffer over
//    NVLink (k_pack_push), then raises a sequence flag in the owner's memory; the owner's k_wait_unpack_add spins on
//    that flag (system-scope acquire), adds the values in increasing peer rank (deterministic; no float atomics) and
//    acknowledges, which is what the sender's next push waits for before it overwrites the buffer.  No NCCL kernel,
//    no staging copy: at 8 ranks ncclSend/ncclRecv of the 2 x 20 MB per rank of BASELINE config 5 took 0.49 ms per
//    step, the push takes the NVLink time of the payload.
//  * NCCL path (fallback; GTK_DISABLE_P2P=1): pack kernel -> ncclSend / ncclRecv in one group -> add kernels.
// NCCL is dlopen'ed so that single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <time.h>
#include <nccl.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cstring>
#include "gtk_internal.h"

int32_t gtk_fastq1_min_layer(gtk_ctx* ctx, const int64_t* d_nz_pos, int64_t n, const int32_t* d_rows, int64_t nb, int* layer,
                             int* max_layer);   // fastq1.cu
bool gtk_fastq1_plan_ok(const gtk_ctx* ctx);
int gtk_fastq1_affine_state(gtk_ctx* ctx);   // 1: every active cell is exactly affine (classifies on first use), 0: not, -1: no plan
int32_t gtk_fastq1_comm_tables(gtk_ctx* ctx, const int64_t* send_nz, int64_t n_send_nz, const int32_t* send_rows, int64_t n_send_b,
                               const int64_t* recv_nz, int64_t n_recv_nz, const int32_t* recv_rows, int64_t n_recv_b,
                               int send_layer, int recv_layer, int32_t** tbl_out, bool* ok);
int64_t gtk_fastq1_plane_nodes(const gtk_ctx* ctx);

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  // a copy already loaded by the host process (e.g. torch's bundled one) wins; else the system library
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
#define LOAD(field, sym) api.field = (decltype(api.field))dlsym(api.handle, sym)
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(Send, "ncclSend");
  LOAD(Recv, "ncclRecv");
  LOAD(AllGather, "ncclAllGather");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.AllGather && api.GroupStart &&
           api.GroupEnd && api.GetErrorString;
  return api;
}

struct Peer {
  int rank = -1;
  int64_t n_send_nz = 0, n_send_b = 0, n_recv_nz = 0, n_recv_b = 0;
  int64_t* send_nz = nullptr;   // device: nz positions (0-based) whose values go to the peer
  int32_t* send_rows = nullptr; // device: rows of b (0-based) that go to the peer
  int64_t* recv_nz = nullptr;   // device: nz positions the received values are added to
  int32_t* recv_rows = nullptr;
  double* send_buf = nullptr;   // [n_send_nz + n_send_b]  (NCCL path)
  double* recv_buf = nullptr;   // = ipc_block + P2P_HDR doubles: [n_recv_nz + n_recv_b]
  int min_send_layer = -1;      // lowest lattice node layer whose sweep segment writes a value sent to this peer (-1: unknown)
  int max_recv_layer = -1;      // highest node layer whose columns hold a position this peer's values are added to (-1: unknown / none)
  // peer-memory path.  ipc_block (raw cudaMalloc, exported over CUDA IPC) = header of P2P_HDR doubles + receive buffer:
  //   header word 0: `ready` sequence number, written by the PEER's k_pack_push after its data landed here
  //   header word 1: `ack` sequence number, written by the PEER's k_wait_unpack_add after it consumed what WE pushed
  double* ipc_block = nullptr;
  size_t ipc_bytes = 0;
  double* remote_block = nullptr;   // the peer's block for us, mapped with cudaIpcOpenMemHandle
  unsigned int* done_ctr = nullptr; // [2] block counters of the push / unpack kernels of this peer
  // exchange counter of the peer-memory path WITH THIS PEER: starts at 0 together with the zero-filled flag header of a
  // freshly created ipc_block, so replacing the plan of a pair (both sides re-export / re-import) restarts the protocol
  unsigned long long seq = 0;
  // exchange fused into the sweep's copy-out (GtkCommDev mode 2): verified per-node buffer offsets, or nullptr
  int32_t* tbl = nullptr;
  size_t tbl_bytes = 0;
  int send_row_layer = -1, recv_row_layer = -1;
};
constexpr int P2P_HDR = 32;   // doubles (256 B) in front of the receive buffer

struct GhostPlan {
  std::vector<Peer> peers;   // sorted by rank
  cudaStream_t side = nullptr;          // pack + NCCL run here while the rest of the sweep runs on ctx->stream
  cudaEvent_t ev_first = nullptr, ev_xchg = nullptr;
  bool p2p_off = false;                 // the ranks agreed to stay on NCCL (gtk_comm_build_exchange could not map every block)
  bool assign_untouched = false;        // device-built plan: entries of columns without local contributions are ASSIGNED by the
                                        // unpack, so the sweep need not write (zero) those columns at all
  bool p2p_ready() const {
    if (peers.empty() || p2p_off || getenv("GTK_DISABLE_P2P")) return false;
    for (auto& p : peers) if (!p.remote_block || !p.ipc_block) return false;
    return true;
  }
};

__global__ void k_pack(const double* __restrict__ nzval, const int64_t* __restrict__ idx, int64_t n,
                       const double* __restrict__ b, const int32_t* __restrict__ rows, int64_t nb,
                       double* __restrict__ buf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + nb; i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = i < n ? nzval[idx[i]] : b[rows[i - n]];
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Sender: gather the ghost entries and store them into the OWNER's receive buffer (peer memory over NVLink); the last
// block to finish raises the owner's `ready` flag.  Before overwriting the buffer every block waits until the owner has
// acknowledged the previous exchange (`ack` flag in OUR block, written by the owner's k_wait_unpack_add).
__global__ void __launch_bounds__(256) k_pack_push(const double* __restrict__ nzval, const int64_t* __restrict__ idx, int64_t n,
                                                   const double* __restrict__ b, const int32_t* __restrict__ rows, int64_t nb,
                                                   double* remote_buf, unsigned long long* remote_ready,
                                                   const unsigned long long* local_ack, unsigned long long seq, unsigned int* ctr) {
  if (threadIdx.x == 0) while (ld_acquire_sys(local_ack) + 1 < seq) __nanosleep(64);
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + nb; i += (int64_t)gridDim.x * blockDim.x)
    remote_buf[i] = i < n ? nzval[idx[i]] : b[rows[i - n]];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ctr, 1u) == gridDim.x - 1) {   // every block's stores are fenced: publish
      *ctr = 0;
      __threadfence_system();
      st_release_sys(remote_ready, seq);
    }
  }
}

// Owner: wait for the peer's data of exchange `seq`, add it (each target position at most once per peer, so no
// atomics), then acknowledge in the PEER's block so that its next push may overwrite the buffer.
__global__ void __launch_bounds__(256) k_wait_unpack_add(double* __restrict__ nzval, const int64_t* __restrict__ idx, int64_t n,
                                                         double* __restrict__ b, const int32_t* __restrict__ rows, int64_t nb,
                                                         const double* buf, const unsigned long long* local_ready,
                                                         unsigned long long* remote_ack, unsigned long long seq, unsigned int* ctr) {
  if (threadIdx.x == 0) while (ld_acquire_sys(local_ready) < seq) __nanosleep(64);
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + nb; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ldcg(buf + i);   // written by another GPU while this kernel runs: not through L1
    if (i < n) {
      const int64_t p = idx[i];
      if (p >= 0) nzval[p] += v; else nzval[~p] = v;   // ~p: a column no local cell contributes to — the value IS the peer's
    } else b[rows[i - n]] += v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ctr, 1u) == gridDim.x - 1) {
      *ctr = 0;
      __threadfence_system();
      st_release_sys(remote_ack, seq);
    }
  }
}

// each target position appears at most once per peer (checked on the host), so no atomics are needed
__global__ void k_unpack_add(double* __restrict__ nzval, const int64_t* __restrict__ idx, int64_t n,
                             double* __restrict__ b, const int32_t* __restrict__ rows, int64_t nb,
                             const double* __restrict__ buf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + nb; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n) {
      const int64_t p = idx[i];
      if (p >= 0) nzval[p] += buf[i]; else nzval[~p] = buf[i];
    } else b[rows[i - n]] += buf[i];
  }
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

void free_peer(gtk_ctx* ctx, Peer& p) {
  gtk_dev_free(ctx, p.send_nz, sizeof(int64_t) * p.n_send_nz);
  gtk_dev_free(ctx, p.send_rows, sizeof(int32_t) * p.n_send_b);
  gtk_dev_free(ctx, p.recv_nz, sizeof(int64_t) * p.n_recv_nz);
  gtk_dev_free(ctx, p.recv_rows, sizeof(int32_t) * p.n_recv_b);
  gtk_dev_free(ctx, p.send_buf, sizeof(double) * (p.n_send_nz + p.n_send_b));
  if (p.remote_block) cudaIpcCloseMemHandle(p.remote_block);
  if (p.ipc_block) { cudaFree(p.ipc_block); ctx->bytes_held -= (int64_t)p.ipc_bytes; }
  if (p.done_ctr) cudaFree(p.done_ctr);
  if (p.tbl) gtk_dev_free(ctx, p.tbl, p.tbl_bytes);
  p.tbl = nullptr;
  p.remote_block = p.ipc_block = p.recv_buf = nullptr; p.done_ctr = nullptr;
}

template <class T>
int32_t upload_idx(gtk_ctx* ctx, T** dst, const T* src, int64_t n) {
  *dst = nullptr;
  if (n == 0) return GTK_OK;
  int32_t rc = gtk_dev_alloc(ctx, (void**)dst, sizeof(T) * n);
  if (rc) return rc;
  GTK_CK(cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, ctx->stream));
  return GTK_OK;
}

}  // namespace

// Takes ownership of the four device index arrays of `p` (allocated with gtk_dev_alloc), allocates its buffers and makes
// it the plan for p.rank (replacing a previous one: the pair then restarts its peer-memory protocol, both sides must
// re-export / re-import).
static int32_t install_peer(gtk_ctx* ctx, Peer p) {
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  if (!g) { g = new GhostPlan(); ctx->ghost = g; }
  for (size_t i = 0; i < g->peers.size(); ++i)
    if (g->peers[i].rank == p.rank) { free_peer(ctx, g->peers[i]); g->peers.erase(g->peers.begin() + i); break; }
  int32_t rc;
  const int64_t n_send = p.n_send_nz + p.n_send_b, n_recv = p.n_recv_nz + p.n_recv_b;
  if (n_send) if ((rc = gtk_dev_alloc(ctx, (void**)&p.send_buf, sizeof(double) * n_send))) return rc;
  // flag header + TWO receive buffers (exchange k uses buffer k & 1: a sender may run one exchange ahead of the owner's
  // unpack) in ONE raw allocation (not pooled): it is exported to the peer over CUDA IPC
  p.ipc_bytes = sizeof(double) * (size_t)(P2P_HDR + 2 * n_recv);
  GTK_CK(cudaMalloc(&p.ipc_block, p.ipc_bytes));
  ctx->bytes_held += (int64_t)p.ipc_bytes;
  GTK_CK(cudaMemsetAsync(p.ipc_block, 0, p.ipc_bytes, ctx->stream));
  p.recv_buf = p.ipc_block + P2P_HDR;
  GTK_CK(cudaMalloc(&p.done_ctr, 2 * sizeof(unsigned int)));
  GTK_CK(cudaMemsetAsync(p.done_ctr, 0, 2 * sizeof(unsigned int), ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  int max_send_layer = -1;
  if ((rc = gtk_fastq1_min_layer(ctx, p.send_nz, p.n_send_nz, p.send_rows, p.n_send_b, &p.min_send_layer, &max_send_layer))) return rc;
  { int lo_unused = -1;
    if ((rc = gtk_fastq1_min_layer(ctx, p.recv_nz, p.n_recv_nz, p.recv_rows, p.n_recv_b, &lo_unused, &p.max_recv_layer))) return rc; }
  {   // tables of the copy-out fused exchange: the exchanged rows of a slab interface lie in the highest layer of the columns
      // that hold them (send: the top node layer, receive: the first owned layer)
    bool ok = false;
    if ((rc = gtk_fastq1_comm_tables(ctx, p.send_nz, p.n_send_nz, p.send_rows, p.n_send_b, p.recv_nz, p.n_recv_nz, p.recv_rows, p.n_recv_b,
                                     max_send_layer, p.max_recv_layer, &p.tbl, &ok))) return rc;
    p.tbl_bytes = ok ? sizeof(int32_t) * 6 * (size_t)gtk_fastq1_plane_nodes(ctx) : 0;
    p.send_row_layer = max_send_layer; p.recv_row_layer = p.max_recv_layer;
  }
  g->peers.push_back(p);
  std::sort(g->peers.begin(), g->peers.end(), [](const Peer& a, const Peer& b) { return a.rank < b.rank; });
  return GTK_OK;
}

#define NCCL_CK(call)                                                                          \
  do {                                                                                         \
    ncclResult_t r_ = (call);                                                                  \
    if (r_ != ncclSuccess) {                                                                   \
      ctx->err = std::string(#call) + ": " + nccl().GetErrorString(r_);                        \
      return GTK_ERR_NCCL;                                                                     \
    }                                                                                          \
  } while (0)

// true when the unpack assigns the entries of columns without local contributions (device-built plan): the sweep kernels
// then restrict themselves to the node layers their active cells touch
bool gtk_comm_assigns_untouched(const gtk_ctx* ctx) {
  const GhostPlan* g = (const GhostPlan*)ctx->ghost;
  return g && g->assign_untouched && !getenv("GTK_SWEEP_HALO");
}

// Descriptor of one fused exchange (called by the sweep launcher when ctx->fuse_comm_want): pointers, sequence numbers,
// layer bounds.  Returns false when the plan does not qualify (no peer memory, more than two peers, layers unknown or
// overlapping); a true return COMMITS the exchange (sequence numbers advance) — the caller must launch.
bool gtk_comm_fused_begin(gtk_ctx* ctx, GtkCommDev* d, int n_layers) {
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  d->on = 0;
  if (!g || g->peers.empty() || g->peers.size() > 2 || !g->p2p_ready() || getenv("GTK_DISABLE_FUSED_COMM")) return false;
  // mode 2 (the exchange rides in the copy-out of the sweep: ghost entries go from registers to the peer's buffer, received
  // ones are added in registers) needs the verified per-node tables of every peer; mode 1 (PUSH / UNPACK work items) pays
  // for scattered 8-byte accesses and only wins for small interfaces (profiles/r02_fused_exchange.txt)
  bool mode2 = !getenv("GTK_DISABLE_FUSED_COPYOUT");
  for (auto& p : g->peers) mode2 = mode2 && p.tbl != nullptr;
  if (!mode2) {
    int64_t total = 0;
    for (auto& p : g->peers) total += p.n_send_nz + p.n_send_b + p.n_recv_nz + p.n_recv_b;
    const char* e = getenv("GTK_FUSED_COMM_MAX_MB");
    const double max_mb = e ? atof(e) : 8.0;
    if (total * 8.0 > max_mb * 1048576.0) return false;
  }
  int top = n_layers, bot = 0;
  for (auto& p : g->peers) {
    if (p.n_send_nz + p.n_send_b) { if (p.min_send_layer < 0) return false; top = std::min(top, p.min_send_layer); }
    if (p.n_recv_nz + p.n_recv_b) { if (p.max_recv_layer < 0) return false; bot = std::max(bot, p.max_recv_layer + 1); }
    if (p.n_send_b && !ctx->bvec) return false;
  }
  if (bot > top) return false;
  d->n_peers = (int)g->peers.size();
  d->top_layer = top; d->bot_layer = bot;
  for (int i = 0; i < d->n_peers; ++i) {
    Peer& p = g->peers[i];
    ++p.seq;
    GtkCommPeerDev& q = d->peer[i];
    const int64_t ns = p.n_send_nz + p.n_send_b, nr = p.n_recv_nz + p.n_recv_b;
    q.send_nz = p.send_nz; q.send_rows = p.send_rows; q.n_send_nz = p.n_send_nz; q.n_send_b = p.n_send_b;
    q.remote_buf = p.remote_block + P2P_HDR + (p.seq & 1) * ns;
    q.remote_ready = reinterpret_cast<unsigned long long*>(p.remote_block) + 0;
    q.local_ack = reinterpret_cast<const unsigned long long*>(p.ipc_block) + 1;
    q.recv_nz = p.recv_nz; q.recv_rows = p.recv_rows; q.n_recv_nz = p.n_recv_nz; q.n_recv_b = p.n_recv_b;
    q.recv_buf = p.recv_buf + (p.seq & 1) * nr;
    q.local_ready = reinterpret_cast<const unsigned long long*>(p.ipc_block) + 0;
    q.remote_ack = reinterpret_cast<unsigned long long*>(p.remote_block) + 1;
    q.seq = p.seq;
    q.push_target = q.unpack_target = 0;   // filled by the launcher (it knows the item counts)
    q.tbl = p.tbl; q.T = p.send_row_layer; q.B = p.recv_row_layer;
  }
  d->on = mode2 ? 2 : 1;
  return true;
}

// The exchange plan indexes nzval / b of ONE pattern: it dies with that pattern (gtk_set_mesh, gtk_set_space,
// gtk_matrix_symbolic) so that a later gtk_comm_sum_ghost_rows cannot scatter through stale positions.  The NCCL
// communicator survives.
void gtk_comm_release_plan(gtk_ctx* ctx) {
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  if (!g) return;
  cudaStreamSynchronize(ctx->stream);
  if (g->side) cudaStreamSynchronize(g->side);
  for (auto& p : g->peers) free_peer(ctx, p);
  if (g->side) cudaStreamDestroy(g->side);
  if (g->ev_first) cudaEventDestroy(g->ev_first);
  if (g->ev_xchg) cudaEventDestroy(g->ev_xchg);
  delete g;
  ctx->ghost = nullptr;
}

void gtk_comm_release(gtk_ctx* ctx) {
  gtk_comm_release_plan(ctx);
  if (ctx->comm && nccl().ok) nccl().CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
}

extern "C" {

int32_t gtk_comm_unique_id(void* id128) {
  if (!id128 || !nccl().ok) return GTK_ERR_NCCL;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return GTK_ERR_NCCL;
  memcpy(id128, &id, sizeof(id));
  return GTK_OK;
}

int32_t gtk_comm_init(gtk_ctx* ctx, int32_t rank, int32_t n_ranks, const void* id128) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!id128 || rank < 0 || rank >= n_ranks) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_init: bad arguments");
  if (!nccl().ok) GTK_FAIL(GTK_ERR_NCCL, "libnccl.so.2 could not be loaded");
  GTK_CK(cudaSetDevice(ctx->device));
  if (ctx->comm) { nccl().CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  NCCL_CK(nccl().CommInitRank(&comm, n_ranks, id, rank));
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->n_ranks = n_ranks;
  // NCCL sets its rings and peer-to-peer channels up lazily, on the first collective / first send-recv of a pair (hundreds
  // of ms each): do that here, as part of creating the communicator, with one tiny all-gather and one exchange with the
  // two neighbouring ranks — the pairs a slab partition talks to — so that building an exchange plan costs what it costs
  if (n_ranks > 1 && !getenv("GTK_NO_NCCL_WARMUP")) {
    int64_t* d = nullptr;
    GTK_CK(gtk_cuda_malloc(ctx, &d, sizeof(int64_t) * (size_t)(n_ranks + 4)));
    GTK_CK(cudaMemsetAsync(d, 0, sizeof(int64_t) * (size_t)(n_ranks + 4), ctx->stream));
    NCCL_CK(nccl().AllGather(d + n_ranks, d, 1, ncclInt64, comm, ctx->stream));
    NCCL_CK(nccl().GroupStart());
    for (int dr = -1; dr <= 1; dr += 2) {
      const int q = rank + dr;
      if (q < 0 || q >= n_ranks) continue;
      NCCL_CK(nccl().Send(d + n_ranks, 1, ncclInt64, q, comm, ctx->stream));
      NCCL_CK(nccl().Recv(d + n_ranks + 2 + (dr > 0), 1, ncclInt64, q, comm, ctx->stream));
    }
    NCCL_CK(nccl().GroupEnd());
    GTK_CK(cudaStreamSynchronize(ctx->stream));
    gtk_cuda_free(ctx, d);
  }
  return GTK_OK;
}

int32_t gtk_comm_set_exchange(gtk_ctx* ctx, int32_t peer, int64_t n_send_nz, const int64_t* send_nz, int64_t n_send_b,
                              const int32_t* send_rows, int64_t n_recv_nz, const int64_t* recv_nz, int64_t n_recv_b,
                              const int32_t* recv_rows) {
  if (!ctx) return GTK_ERR_INVALID;
  if (peer < 0 || peer == ctx->rank || (ctx->comm && peer >= ctx->n_ranks) || n_send_nz < 0 || n_send_b < 0 ||
      n_recv_nz < 0 || n_recv_b < 0 || (n_send_nz && !send_nz) || (n_send_b && !send_rows) || (n_recv_nz && !recv_nz) ||
      (n_recv_b && !recv_rows))
    GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_set_exchange: bad arguments");
  if (!ctx->ms.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_comm_set_exchange: call gtk_matrix_symbolic first (the plan indexes its nzval)");
  {   // the kernels scatter through these positions unchecked: validate them once, here
    const int64_t nnz = ctx->ms.nnz, nr = ctx->ms.n_rows;
    for (int64_t i = 0; i < n_send_nz; ++i) if (send_nz[i] < 0 || send_nz[i] >= nnz) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_set_exchange: send_nz position outside [0, nnz)");
    for (int64_t i = 0; i < n_recv_nz; ++i) if (recv_nz[i] < 0 || recv_nz[i] >= nnz) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_set_exchange: recv_nz position outside [0, nnz)");
    for (int64_t i = 0; i < n_send_b; ++i) if (send_rows[i] < 0 || send_rows[i] >= nr) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_set_exchange: send_rows entry outside [0, n_rows)");
    for (int64_t i = 0; i < n_recv_b; ++i) if (recv_rows[i] < 0 || recv_rows[i] >= nr) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_set_exchange: recv_rows entry outside [0, n_rows)");
  }
  GTK_CK(cudaSetDevice(ctx->device));
  Peer p;
  p.rank = peer;
  p.n_send_nz = n_send_nz; p.n_send_b = n_send_b; p.n_recv_nz = n_recv_nz; p.n_recv_b = n_recv_b;
  int32_t rc;
  if ((rc = upload_idx(ctx, &p.send_nz, send_nz, n_send_nz))) return rc;
  if ((rc = upload_idx(ctx, &p.send_rows, send_rows, n_send_b))) return rc;
  if ((rc = upload_idx(ctx, &p.recv_nz, recv_nz, n_recv_nz))) return rc;
  if ((rc = upload_idx(ctx, &p.recv_rows, recv_rows, n_recv_b))) return rc;
  return install_peer(ctx, p);
}

}  // extern "C"

// pack + send/recv on stream `st`
static int32_t exchange_on(gtk_ctx* ctx, GhostPlan* g, cudaStream_t st) {
  if (g->p2p_ready()) {   // gather + store into the owner's buffer over NVLink + flag, one kernel per peer
    for (auto& p : g->peers) {
      ++p.seq;   // one exchange = one push and one unpack per peer; both kernels of this exchange use the same number
      const int64_t n = p.n_send_nz + p.n_send_b;
      if (n == 0) continue;
      if (p.n_send_b && !ctx->bvec) GTK_FAIL(GTK_ERR_STATE, "ghost rows of b requested but no vector assembled");
      unsigned long long* rhdr = reinterpret_cast<unsigned long long*>(p.remote_block);
      const unsigned long long* lhdr = reinterpret_cast<const unsigned long long*>(p.ipc_block);
      { GtkProf pr_(ctx, "k_pack_push");
        k_pack_push<<<std::min(grid_for(n), 4 * ctx->sm_count), 256, 0, st>>>(ctx->nzval, p.send_nz, p.n_send_nz, ctx->bvec, p.send_rows, p.n_send_b,
                                                 p.remote_block + P2P_HDR + (p.seq & 1) * n, rhdr + 0, lhdr + 1, p.seq, p.done_ctr + 0); }
      GTK_CK(cudaGetLastError());
      gtk_count_launch(ctx);
    }
    return GTK_OK;
  }
  if (!ctx->comm) GTK_FAIL(GTK_ERR_STATE, "ghost-row exchange: neither peer memory (gtk_comm_p2p_import) nor NCCL (gtk_comm_init) is set up");
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  for (auto& p : g->peers) {
    const int64_t n = p.n_send_nz + p.n_send_b;
    if (n == 0) continue;
    if (p.n_send_b && !ctx->bvec) GTK_FAIL(GTK_ERR_STATE, "ghost rows of b requested but no vector assembled");
    { GtkProf pr_(ctx, "k_pack"); k_pack<<<grid_for(n), 256, 0, st>>>(ctx->nzval, p.send_nz, p.n_send_nz, ctx->bvec, p.send_rows, p.n_send_b, p.send_buf); }
    GTK_CK(cudaGetLastError());
    gtk_count_launch(ctx);
  }
  NCCL_CK(nccl().GroupStart());
  for (auto& p : g->peers) {
    if (p.n_send_nz + p.n_send_b) NCCL_CK(nccl().Send(p.send_buf, (size_t)(p.n_send_nz + p.n_send_b), ncclFloat64, p.rank, comm, st));
    if (p.n_recv_nz + p.n_recv_b) NCCL_CK(nccl().Recv(p.recv_buf, (size_t)(p.n_recv_nz + p.n_recv_b), ncclFloat64, p.rank, comm, st));
  }
  NCCL_CK(nccl().GroupEnd());
  return GTK_OK;
}

// add what the peers sent, in increasing peer rank: fixed summation order
static int32_t unpack_on(gtk_ctx* ctx, GhostPlan* g, cudaStream_t st, bool alone = true) {
  if (g->p2p_ready()) {
    for (auto& p : g->peers) {
      const int64_t n = p.n_recv_nz + p.n_recv_b;
      if (n == 0) continue;
      const unsigned long long* lhdr = reinterpret_cast<const unsigned long long*>(p.ipc_block);
      unsigned long long* rhdr = reinterpret_cast<unsigned long long*>(p.remote_block);
      { GtkProf pr_(ctx, "k_wait_unpack_add");
        // sharing the GPU with the sweep: at most 2 blocks per SM (a block may spin for the peer's flag); alone: the
        // scattered read-modify-writes are latency-bound, so as many threads as there are entries to hide it
        k_wait_unpack_add<<<alone ? std::min((int)((n + 255) / 256), 16 * ctx->sm_count) : std::min(grid_for(n), 2 * ctx->sm_count), 256, 0, st>>>(ctx->nzval, p.recv_nz, p.n_recv_nz, ctx->bvec, p.recv_rows, p.n_recv_b,
                                                       p.recv_buf + (p.seq & 1) * n, lhdr + 0, rhdr + 1, p.seq, p.done_ctr + 1); }
      GTK_CK(cudaGetLastError());
      gtk_count_launch(ctx);
    }
    return GTK_OK;
  }
  for (auto& p : g->peers) {
    const int64_t n = p.n_recv_nz + p.n_recv_b;
    if (n == 0) continue;
    { GtkProf pr_(ctx, "k_unpack_add"); k_unpack_add<<<grid_for(n), 256, 0, st>>>(ctx->nzval, p.recv_nz, p.n_recv_nz, ctx->bvec, p.recv_rows, p.n_recv_b, p.recv_buf); }
    GTK_CK(cudaGetLastError());
    gtk_count_launch(ctx);
  }
  return GTK_OK;
}

extern "C" {

int32_t gtk_comm_sum_ghost_rows(gtk_ctx* ctx) {
  if (!ctx) return GTK_ERR_INVALID;
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  if (!g || g->peers.empty()) return GTK_OK;
  if (!ctx->comm && !g->p2p_ready()) GTK_FAIL(GTK_ERR_STATE, "gtk_comm_sum_ghost_rows: call gtk_comm_init (NCCL) or gtk_comm_p2p_import (peer memory) first");
  if (!ctx->nzval) GTK_FAIL(GTK_ERR_STATE, "gtk_comm_sum_ghost_rows: nothing assembled yet");
  GTK_CK(cudaSetDevice(ctx->device));
  int32_t rc = exchange_on(ctx, g, ctx->stream);
  if (rc) return rc;
  return unpack_on(ctx, g, ctx->stream);
}

// Numeric assembly + ghost-row summation with the exchange hidden behind the sweep: the z-segments that produce the
// values a peer waits for run first; their pack + ncclSend/ncclRecv go to a side stream while the remaining segments
// run on the main stream; the received partial sums are added once both are done.  Same values, same order of additions
// as gtk_assemble_matrix_and_vector_device + gtk_comm_sum_ghost_rows (bitwise), which is also what runs when the sweep
// kernels do not apply.
int32_t gtk_assemble_and_sum_ghost_rows_device(gtk_ctx* ctx, int32_t mform, const gtk_form_params* pm, int32_t vform,
                                               const gtk_form_params* pv) {
  if (!ctx) return GTK_ERR_INVALID;
  GTK_CK(cudaSetDevice(ctx->device));
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  int32_t rc;
  // top part: node layers >= `layer` hold what goes to a peer (none on a rank that only receives: empty top part)
  int layer = 0x3FFFFFFF;
  if (g) for (auto& p : g->peers) if (p.n_send_nz + p.n_send_b) layer = p.min_send_layer < 0 ? -1 : (layer < 0 ? -1 : (p.min_send_layer < layer ? p.min_send_layer : layer));
  const bool overlap = g && !g->peers.empty() && (ctx->comm || g->p2p_ready()) && layer >= 0 && gtk_fastq1_plan_ok(ctx) &&
                       !getenv("GTK_DISABLE_OVERLAP");
  if (!overlap) {
    if ((rc = gtk_numeric_both_impl(ctx, mform, pm, vform, pv))) return rc;
    return gtk_comm_sum_ghost_rows(ctx);
  }
  // First choice: ONE persistent kernel that sweeps, pushes the ghost entries over NVLink and adds the received ones
  // (work items of the same launch, see GtkCommDev) — when the exactly-affine sweep kernel applies and the transport is
  // peer memory with at most two peers (a slab partition).  Anything else takes the multi-launch overlap below.
  if (g->p2p_ready() && g->peers.size() <= 2 && !getenv("GTK_DISABLE_FUSED_COMM") && gtk_fastq1_affine_state(ctx) == 1) {
    ctx->fuse_comm_want = true; ctx->fuse_comm_done = false;
    rc = gtk_numeric_both_impl(ctx, mform, pm, vform, pv);
    ctx->fuse_comm_want = false;
    if (rc) return rc;
    if (ctx->fuse_comm_done) return GTK_OK;
    return gtk_comm_sum_ghost_rows(ctx);   // the launcher declined (form / tabulation / layer bounds): serial exchange, same bits
  }
  if (!g->side) {
    // highest priority: the block scheduler otherwise keeps feeding the (much larger) sweep grid launched right after
    // and the pack / NCCL kernels would only start once that grid has been dispatched completely
    int least = 0, greatest = 0;
    GTK_CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    GTK_CK(cudaStreamCreateWithPriority(&g->side, cudaStreamNonBlocking, greatest));
    GTK_CK(cudaEventCreateWithFlags(&g->ev_first, cudaEventDisableTiming));
    GTK_CK(cudaEventCreateWithFlags(&g->ev_xchg, cudaEventDisableTiming));
  }
  static const bool timing = getenv("GTK_COMM_TIMING") != nullptr;   // debug: device timeline of one overlapped step
  cudaEvent_t te[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (timing) { for (auto& e : te) cudaEventCreate(&e); cudaEventRecord(te[0], ctx->stream); }
  // bottom part: the node layers whose columns hold the positions the peers' values are added to.  Sweeping them early
  // lets the unpack run on the side stream too, concurrently with the middle of the sweep.
  int lo = 0;
  bool lo_known = true;
  for (auto& p : g->peers) if (p.n_recv_nz + p.n_recv_b) {
    if (p.max_recv_layer < 0) lo_known = false;
    else if (p.max_recv_layer + 1 > lo) lo = p.max_recv_layer + 1;
  }
  // Measured on 2 GPUs: with 1.2 MB received per step (128^3 slabs) the early unpack gains 2 % (0.1531 -> 0.1498 ms); with
  // 19.8 MB (512x512x64 slabs) it LOSES 3 % (0.998 -> 1.026 ms): the sweep runs at 5 CTAs/SM with the register file full,
  // so every block of the spinning / adding unpack kernel that is resident displaces one sweep CTA of its SM for as long as
  // it lives.  Hence: early only for small messages (GTK_EARLY_UNPACK=0/1 overrides).
  int64_t recv_total = 0;
  for (auto& p : g->peers) recv_total += p.n_recv_nz + p.n_recv_b;
  const char* eu = getenv("GTK_EARLY_UNPACK");
  const bool want_early = eu ? atoi(eu) != 0 : recv_total * 8 <= (4 << 20);
  const bool early_unpack = lo_known && lo <= layer && want_early && !getenv("GTK_DISABLE_EARLY_UNPACK");
  if (!early_unpack) lo = 0;
  ctx->seg_mode = 1; ctx->seg_layer = layer; ctx->seg_lo = lo;
  rc = gtk_numeric_both_impl(ctx, mform, pm, vform, pv);
  ctx->seg_mode = 0;
  if (rc) return rc;
  if (timing) cudaEventRecord(te[1], ctx->stream);
  int64_t launches_acc = ctx->launches_last;
  if (ctx->fast_path_last != 1 && ctx->fast_path_last != 2)   // the sweep declined (form, tabulation): everything is assembled already
    return gtk_comm_sum_ghost_rows(ctx);
  if (lo > 0) {
    ctx->seg_mode = 3; ctx->seg_layer = layer; ctx->seg_lo = lo;
    rc = gtk_numeric_both_impl(ctx, mform, pm, vform, pv);
    ctx->seg_mode = 0;
    if (rc) return rc;
    launches_acc += ctx->launches_last;
  }
  GTK_CK(cudaEventRecord(g->ev_first, ctx->stream));
  GTK_CK(cudaStreamWaitEvent(g->side, g->ev_first, 0));
  ctx->launches_last = 0;
  if ((rc = exchange_on(ctx, g, g->side))) return rc;
  if (early_unpack && (rc = unpack_on(ctx, g, g->side, false))) return rc;
  GTK_CK(cudaEventRecord(g->ev_xchg, g->side));
  if (timing) cudaEventRecord(te[2], g->side);
  launches_acc += ctx->launches_last;
  ctx->seg_mode = 2; ctx->seg_layer = layer; ctx->seg_lo = lo;
  rc = gtk_numeric_both_impl(ctx, mform, pm, vform, pv);
  ctx->seg_mode = 0;
  if (rc) return rc;
  launches_acc += ctx->launches_last;
  ctx->launches_last = 0;
  if (timing) cudaEventRecord(te[3], ctx->stream);
  GTK_CK(cudaStreamWaitEvent(ctx->stream, g->ev_xchg, 0));
  if (!early_unpack) rc = unpack_on(ctx, g, ctx->stream);
  ctx->launches_last += launches_acc;   // numeric_both_impl restarts the per-call counter
  if (timing) {
    cudaEventRecord(te[4], ctx->stream);
    cudaEventSynchronize(te[4]);
    cudaEventSynchronize(te[2]);
    float t[5] = {0, 0, 0, 0, 0};
    for (int i = 1; i < 5; ++i) cudaEventElapsedTime(&t[i], te[0], te[i]);
    fprintf(stderr, "[gtk rank %d] overlap timeline (ms after start): first segments %.4f | exchange done %.4f | rest of sweep %.4f | unpack done %.4f\n",
            ctx->rank, t[1], t[2], t[3], t[4]);
    for (auto& e : te) cudaEventDestroy(e);
  }
  return rc;
}

// Peer-memory transport: every rank exports, per peer, the block it receives that peer's values in (64-byte CUDA IPC
// handle); the host carries the handles to the peers by its own means (as it does for the NCCL unique id) and each rank
// imports the handle of the block its peer keeps for it.  Once every peer of a rank is imported the exchange uses direct
// NVLink stores + flags instead of NCCL.  Needs peer access between the GPUs (NVLink / NVSwitch box).
int32_t gtk_comm_p2p_export(gtk_ctx* ctx, int32_t peer, void* handle64) {
  if (!ctx || !handle64) return GTK_ERR_INVALID;
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  if (g) for (auto& p : g->peers) if (p.rank == peer) {
    GTK_CK(cudaSetDevice(ctx->device));
    GTK_CK(cudaStreamSynchronize(ctx->stream));   // the block's zero-fill must be done before a peer can write to it
    cudaIpcMemHandle_t h;
    GTK_CK(cudaIpcGetMemHandle(&h, p.ipc_block));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(handle64, &h, sizeof(h));
    return GTK_OK;
  }
  GTK_FAIL(GTK_ERR_STATE, "gtk_comm_p2p_export: no exchange plan for this peer (gtk_comm_set_exchange first)");
}

int32_t gtk_comm_p2p_import(gtk_ctx* ctx, int32_t peer, const void* handle64) {
  if (!ctx || !handle64) return GTK_ERR_INVALID;
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  if (g) for (auto& p : g->peers) if (p.rank == peer) {
    GTK_CK(cudaSetDevice(ctx->device));
    if (p.remote_block) { cudaIpcCloseMemHandle(p.remote_block); p.remote_block = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* ptr = nullptr;
    GTK_CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p.remote_block = (double*)ptr;
    return GTK_OK;
  }
  GTK_FAIL(GTK_ERR_STATE, "gtk_comm_p2p_import: no exchange plan for this peer (gtk_comm_set_exchange first)");
}

int64_t gtk_comm_ghost_info(const gtk_ctx* ctx, int32_t key) {
  if (!ctx) return -1;
  const GhostPlan* g = (const GhostPlan*)ctx->ghost;
  int64_t s = 0, r = 0;
  if (g) for (auto& p : g->peers) { s += p.n_send_nz + p.n_send_b; r += p.n_recv_nz + p.n_recv_b; }
  switch (key) {
    case 0: return s;
    case 1: return r;
    case 2: return 8 * (s + r);
    case 3: return g && g->p2p_ready() ? 1 : 0;   // transport of the next exchange: 1 peer memory, 0 NCCL
    default: return -1;
  }
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Exchange plan built ON THE DEVICE (gtk_comm_build_exchange).  Block row partition, PartitionedArrays'
// variable_partition data model: local free row i has global id gid0 + i and rank p owns the global ids
// [own_start[p], own_start[p+1]).  What the host numpy version (partition.py: ghost_send_lists / match_received /
// build_exchange_plan) does with Python objects over torch.distributed happens here with three kernels, CUB stream
// compaction and NCCL (counts: ncclAllGather; (row, column) keys: ncclSend/ncclRecv between the ranks that share rows).
namespace {

__global__ void k_touch_rows(const int32_t* __restrict__ cell_dofs, int64_t e0, int64_t e1, uint8_t* __restrict__ touched) {
  for (int64_t e = e0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < e1; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = cell_dofs[e];
    if (d > 0) touched[d - 1] = 1;     // same value from every writer
  }
}

struct GhostNzPred {   // nz position -> stored in a row of [lo, hi) that the active cells touch
  const int32_t* rowval; const uint8_t* touched; int64_t lo, hi;
  __device__ __forceinline__ bool operator()(const int64_t& p) const {
    const int64_t r = (int64_t)rowval[p] - 1;
    return r >= lo && r < hi && touched[r];
  }
};
struct GhostRowPred {
  const uint8_t* touched;
  __device__ __forceinline__ bool operator()(const int32_t& r) const { return touched[r] != 0; }
};

// keys of the selected entries: [0,n) global row, [n,2n) global column, [2n, 2n+nb) global row of the b entries
__global__ void k_ghost_keys(const int64_t* __restrict__ nz_pos, int64_t n, const int32_t* __restrict__ b_rows, int64_t nb,
                             const int32_t* __restrict__ rowval, const int64_t* __restrict__ colptr, int64_t n_cols, int64_t gid0,
                             int64_t* __restrict__ keys) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + nb; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= n) { keys[2 * n + (i - n)] = gid0 + b_rows[i - n]; continue; }
    const int64_t p = nz_pos[i];
    int64_t lo = 0, hi = n_cols;            // last column c with colptr[c] <= p
    while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if (colptr[mid] <= p) lo = mid; else hi = mid - 1; }
    keys[i] = gid0 + (int64_t)rowval[p] - 1;
    keys[n + i] = gid0 + lo;
  }
}

// owner side: where the announced entries are added.  err bits: 1 row/column not local, 2 entry missing from the pattern,
// 4 row not owned by this rank
__global__ void k_match_keys(const int64_t* __restrict__ keys, int64_t n, int64_t nb, const int32_t* __restrict__ rowval,
                             const int64_t* __restrict__ colptr, int64_t n_rows, int64_t gid0, int64_t own_lo, int64_t own_hi,
                             const uint8_t* __restrict__ touched, int64_t* __restrict__ recv_nz, int32_t* __restrict__ recv_rows,
                             int* __restrict__ err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + nb; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= n) {
      const int64_t g = keys[2 * n + (i - n)];
      if (g < own_lo || g >= own_hi || g - gid0 < 0 || g - gid0 >= n_rows) { atomicOr(err, 4); recv_rows[i - n] = 0; }
      else recv_rows[i - n] = (int32_t)(g - gid0);
      continue;
    }
    const int64_t gr = keys[i], gc = keys[n + i];
    const int64_t lr = gr - gid0, lc = gc - gid0;
    if (lr < 0 || lr >= n_rows || lc < 0 || lc >= n_rows) { atomicOr(err, 1); recv_nz[i] = 0; continue; }
    if (gr < own_lo || gr >= own_hi) atomicOr(err, 4);
    int64_t lo = colptr[lc], hi = colptr[lc + 1];
    const int32_t target = (int32_t)(lr + 1);
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (rowval[mid] < target) lo = mid + 1; else hi = mid; }
    if (lo >= colptr[lc + 1] || rowval[lo] != target) { atomicOr(err, 2); recv_nz[i] = 0; }
    else recv_nz[i] = touched[lc] ? lo : ~lo;   // no active cell touches column lc: assign instead of add
  }
}

struct Iota64 {   // counting iterator without thrust
  using value_type = int64_t; using difference_type = int64_t; using pointer = const int64_t*; using reference = int64_t;
  using iterator_category = std::random_access_iterator_tag;
  int64_t v;
  __host__ __device__ int64_t operator[](int64_t i) const { return v + i; }
  __host__ __device__ int64_t operator*() const { return v; }
  __host__ __device__ Iota64 operator+(int64_t i) const { return Iota64{v + i}; }
};
struct Iota32 {
  using value_type = int32_t; using difference_type = int64_t; using pointer = const int32_t*; using reference = int32_t;
  using iterator_category = std::random_access_iterator_tag;
  int64_t v;
  __host__ __device__ int32_t operator[](int64_t i) const { return (int32_t)(v + i); }
  __host__ __device__ int32_t operator*() const { return (int32_t)v; }
  __host__ __device__ Iota32 operator+(int64_t i) const { return Iota32{v + i}; }
};

}  // namespace

static double wall_ms() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return 1e3 * t.tv_sec + 1e-6 * t.tv_nsec; }
#define GTK_LAP(label) do { if (timing) { cudaStreamSynchronize(st); const double t_ = wall_ms(); fprintf(stderr, "[gtk rank %d] %s: %.2f ms\n", ctx->rank, label, t_ - t_last); t_last = t_; } } while (0)

extern "C" int32_t gtk_comm_build_exchange(gtk_ctx* ctx, int64_t gid0, const int64_t* own_start) {
  if (!ctx) return GTK_ERR_INVALID;
  const bool timing = getenv("GTK_COMM_TIMING") != nullptr;
  double t_last = wall_ms();
  if (!own_start) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_build_exchange: own_start is null");
  if (!ctx->comm) GTK_FAIL(GTK_ERR_STATE, "gtk_comm_build_exchange: call gtk_comm_init first");
  MatSym& m = ctx->ms;
  if (!m.ready || ctx->cur_slot != 0 || m.rows_fd != GTK_FREE || m.cols_fd != GTK_FREE)
    GTK_FAIL(GTK_ERR_STATE, "gtk_comm_build_exchange: needs the free x free pattern of slot 0 (gtk_matrix_symbolic)");
  const int W = ctx->n_ranks, me = ctx->rank;
  const int64_t n = m.n_rows;
  for (int p = 0; p < W; ++p) if (own_start[p] > own_start[p + 1]) GTK_FAIL(GTK_ERR_INVALID, "gtk_comm_build_exchange: own_start must be non-decreasing");
  GTK_CK(cudaSetDevice(ctx->device));
  gtk_comm_release_plan(ctx);
  cudaStream_t st = ctx->stream;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  int32_t rc = GTK_OK;
  // 1. rows the active cells contribute to
  uint8_t* touched = nullptr;
  GTK_CK(gtk_cuda_malloc(ctx, &touched, (size_t)(n ? n : 1)));
  GTK_CK(cudaMemsetAsync(touched, 0, (size_t)(n ? n : 1), st));
  {
    const int64_t a0 = ctx->act_count < 0 ? 0 : ctx->act_first, a1 = ctx->act_count < 0 ? ctx->n_cells : ctx->act_first + ctx->act_count;
    if (a1 > a0) k_touch_rows<<<grid_for((a1 - a0) * ctx->nld), 256, 0, st>>>(ctx->cell_dofs, a0 * ctx->nld, a1 * ctx->nld, touched);
  }
  // 2. per peer: nz positions (CSC order) and b rows stored in the rows that peer owns
  struct Send { int peer; int64_t lo, hi, n_nz = 0, n_b = 0; int64_t* nz = nullptr; int32_t* rows = nullptr; int64_t* keys = nullptr; };
  std::vector<Send> sends;
  int64_t* d_cnt = nullptr;
  GTK_CK(gtk_cuda_malloc(ctx, &d_cnt, 2 * sizeof(int64_t)));
  void* tmp = nullptr; size_t tmp_bytes = 0;
  for (int p = 0; p < W; ++p) {
    if (p == me) continue;
    Send sd; sd.peer = p;
    sd.lo = std::max<int64_t>(own_start[p] - gid0, 0); sd.hi = std::min<int64_t>(own_start[p + 1] - gid0, n);
    if (sd.hi <= sd.lo) continue;
    int64_t* cand_nz = nullptr; int32_t* cand_rows = nullptr;
    GTK_CK(gtk_cuda_malloc(ctx, &cand_nz, sizeof(int64_t) * (size_t)(m.nnz ? m.nnz : 1)));
    GTK_CK(gtk_cuda_malloc(ctx, &cand_rows, sizeof(int32_t) * (size_t)(sd.hi - sd.lo)));
    size_t need1 = 0, need2 = 0;
    GhostNzPred pn{m.rowval, touched, sd.lo, sd.hi};
    GhostRowPred pr{touched};
    cub::DeviceSelect::If(nullptr, need1, Iota64{0}, cand_nz, d_cnt, m.nnz, pn, st);
    cub::DeviceSelect::If(nullptr, need2, Iota32{sd.lo}, cand_rows, d_cnt + 1, sd.hi - sd.lo, pr, st);
    const size_t need = std::max(need1, need2);
    if (need > tmp_bytes) { gtk_cuda_free(ctx, tmp); GTK_CK(gtk_cuda_malloc(ctx, &tmp, need)); tmp_bytes = need; }
    GTK_CK(cub::DeviceSelect::If(tmp, need1, Iota64{0}, cand_nz, d_cnt, m.nnz, pn, st));
    GTK_CK(cub::DeviceSelect::If(tmp, need2, Iota32{sd.lo}, cand_rows, d_cnt + 1, sd.hi - sd.lo, pr, st));
    int64_t h_cnt[2];
    GTK_CK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    GTK_CK(cudaStreamSynchronize(st));
    sd.n_nz = h_cnt[0]; sd.n_b = h_cnt[1];
    if (sd.n_nz + sd.n_b) {   // exact-size copies owned by the plan (allocated the way free_peer releases them)
      if (sd.n_nz) { if ((rc = gtk_dev_alloc(ctx, (void**)&sd.nz, sizeof(int64_t) * sd.n_nz))) return rc;
                     GTK_CK(cudaMemcpyAsync(sd.nz, cand_nz, sizeof(int64_t) * sd.n_nz, cudaMemcpyDeviceToDevice, st)); }
      if (sd.n_b) { if ((rc = gtk_dev_alloc(ctx, (void**)&sd.rows, sizeof(int32_t) * sd.n_b))) return rc;
                    GTK_CK(cudaMemcpyAsync(sd.rows, cand_rows, sizeof(int32_t) * sd.n_b, cudaMemcpyDeviceToDevice, st)); }
      GTK_CK(gtk_cuda_malloc(ctx, &sd.keys, sizeof(int64_t) * (size_t)(2 * sd.n_nz + sd.n_b)));
      k_ghost_keys<<<grid_for(sd.n_nz + sd.n_b), 256, 0, st>>>(sd.nz, sd.n_nz, sd.rows, sd.n_b, m.rowval, m.colptr, m.n_cols, gid0, sd.keys);
      GTK_CK(cudaGetLastError());
      sends.push_back(sd);
    }
    GTK_CK(cudaStreamSynchronize(st));
    gtk_cuda_free(ctx, cand_nz); gtk_cuda_free(ctx, cand_rows);
  }
  gtk_cuda_free(ctx, tmp); gtk_cuda_free(ctx, d_cnt);
  GTK_LAP("build_exchange: select ghost entries + keys");
  // 3. everybody learns how much it receives from whom
  std::vector<int64_t> row(2 * W, 0), mat((size_t)2 * W * W, 0);
  for (auto& sd : sends) { row[2 * sd.peer] = sd.n_nz; row[2 * sd.peer + 1] = sd.n_b; }
  int64_t *d_row = nullptr, *d_mat = nullptr;
  GTK_CK(gtk_cuda_malloc(ctx, &d_row, sizeof(int64_t) * 2 * W));
  GTK_CK(gtk_cuda_malloc(ctx, &d_mat, sizeof(int64_t) * 2 * W * W));
  GTK_CK(cudaMemcpyAsync(d_row, row.data(), sizeof(int64_t) * 2 * W, cudaMemcpyHostToDevice, st));
  NCCL_CK(nccl().AllGather(d_row, d_mat, (size_t)2 * W, ncclInt64, comm, st));
  GTK_CK(cudaMemcpyAsync(mat.data(), d_mat, sizeof(int64_t) * 2 * W * W, cudaMemcpyDeviceToHost, st));
  GTK_CK(cudaStreamSynchronize(st));
  gtk_cuda_free(ctx, d_row); gtk_cuda_free(ctx, d_mat);
  GTK_LAP("build_exchange: all-gather of counts");
  // 4. keys travel to the owners
  struct Recv { int peer; int64_t n_nz, n_b; int64_t* keys = nullptr; };
  std::vector<Recv> recvs;
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    const int64_t rn = mat[((size_t)q * W + me) * 2], rb = mat[((size_t)q * W + me) * 2 + 1];
    if (rn + rb == 0) continue;
    Recv rv{q, rn, rb};
    GTK_CK(gtk_cuda_malloc(ctx, &rv.keys, sizeof(int64_t) * (size_t)(2 * rn + rb)));
    recvs.push_back(rv);
  }
  NCCL_CK(nccl().GroupStart());
  for (auto& sd : sends) NCCL_CK(nccl().Send(sd.keys, (size_t)(2 * sd.n_nz + sd.n_b), ncclInt64, sd.peer, comm, st));
  for (auto& rv : recvs) NCCL_CK(nccl().Recv(rv.keys, (size_t)(2 * rv.n_nz + rv.n_b), ncclInt64, rv.peer, comm, st));
  NCCL_CK(nccl().GroupEnd());
  GTK_LAP("build_exchange: send/recv of keys");
  // 5. owners locate the announced entries in their own pattern
  int* d_err = nullptr;
  GTK_CK(gtk_cuda_malloc(ctx, &d_err, sizeof(int)));
  GTK_CK(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  std::vector<Peer> peers;
  auto peer_of = [&](int r) -> Peer& {
    for (auto& p : peers) if (p.rank == r) return p;
    peers.emplace_back(); peers.back().rank = r; return peers.back();
  };
  for (auto& sd : sends) { Peer& p = peer_of(sd.peer); p.n_send_nz = sd.n_nz; p.n_send_b = sd.n_b; p.send_nz = sd.nz; p.send_rows = sd.rows; }
  for (auto& rv : recvs) {
    Peer& p = peer_of(rv.peer);
    p.n_recv_nz = rv.n_nz; p.n_recv_b = rv.n_b;
    if (rv.n_nz) if ((rc = gtk_dev_alloc(ctx, (void**)&p.recv_nz, sizeof(int64_t) * rv.n_nz))) return rc;
    if (rv.n_b) if ((rc = gtk_dev_alloc(ctx, (void**)&p.recv_rows, sizeof(int32_t) * rv.n_b))) return rc;
    k_match_keys<<<grid_for(rv.n_nz + rv.n_b), 256, 0, st>>>(rv.keys, rv.n_nz, rv.n_b, m.rowval, m.colptr, n, gid0, own_start[me], own_start[me + 1],
                                                            touched, p.recv_nz, p.recv_rows, d_err);
    GTK_CK(cudaGetLastError());
  }
  int h_err = 0;
  GTK_CK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  GTK_CK(cudaStreamSynchronize(st));
  gtk_cuda_free(ctx, d_err); gtk_cuda_free(ctx, touched);
  for (auto& sd : sends) gtk_cuda_free(ctx, sd.keys);
  for (auto& rv : recvs) gtk_cuda_free(ctx, rv.keys);
  if (h_err) {
    for (auto& p : peers) free_peer(ctx, p);
    GTK_FAIL(GTK_ERR_INVALID, std::string("gtk_comm_build_exchange: a peer announced a ghost-row entry that ") +
                                  ((h_err & 1) ? "is not a local dof of the owner; " : "") + ((h_err & 2) ? "is missing from the owner's sparsity pattern; " : "") +
                                  ((h_err & 4) ? "lies in a row the receiver does not own; " : ""));
  }
  GTK_LAP("build_exchange: match keys");
  std::sort(peers.begin(), peers.end(), [](const Peer& a, const Peer& b) { return a.rank < b.rank; });
  for (auto& p : peers) if ((rc = install_peer(ctx, p))) return rc;
  GTK_LAP("build_exchange: install peers (buffers, min/max layers)");
  if (ctx->ghost) {
    ((GhostPlan*)ctx->ghost)->assign_untouched = true;
    // columns no active cell touches are not swept any more: they hold what the unpack assigns, or zeros (rows nobody sends)
    if (ctx->nzval) GTK_CK(cudaMemsetAsync(ctx->nzval, 0, sizeof(double) * ctx->nzval_cap, st));
  }
  return GTK_OK;   // transport: gtk_comm_connect_peer_memory (collective, also for a rank without peers)
}

// Second collective half of gtk_comm_build_exchange, separate so that EVERY rank (also one without peers) reaches the
// agreement AllGather: swaps the CUDA IPC handles of the receive blocks over NCCL and imports them; if any rank failed
// to map a block all ranks stay on NCCL.
extern "C" int32_t gtk_comm_connect_peer_memory(gtk_ctx* ctx) {
  if (!ctx) return GTK_ERR_INVALID;
  if (!ctx->comm) GTK_FAIL(GTK_ERR_STATE, "gtk_comm_connect_peer_memory: call gtk_comm_init first");
  GTK_CK(cudaSetDevice(ctx->device));
  GhostPlan* g = (GhostPlan*)ctx->ghost;
  cudaStream_t st = ctx->stream;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  const int W = ctx->n_ranks;
  const bool timing = getenv("GTK_COMM_TIMING") != nullptr;
  double t_last = wall_ms();
  int ok = getenv("GTK_DISABLE_P2P") ? 0 : 1;
  const size_t np = g ? g->peers.size() : 0;
  unsigned char* d_h = nullptr;   // [np] mine, [np] theirs, 64 bytes each
  GTK_CK(gtk_cuda_malloc(ctx, &d_h, 128 * (np ? np : 1)));
  std::vector<unsigned char> mine(64 * (np ? np : 1), 0), theirs(64 * (np ? np : 1), 0);
  for (size_t i = 0; i < np; ++i) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, g->peers[i].ipc_block) != cudaSuccess) { cudaGetLastError(); ok = 0; memset(&h, 0, sizeof(h)); }
    memcpy(mine.data() + 64 * i, &h, 64);
  }
  if (np) GTK_CK(cudaMemcpyAsync(d_h, mine.data(), 64 * np, cudaMemcpyHostToDevice, st));
  NCCL_CK(nccl().GroupStart());
  for (size_t i = 0; i < np; ++i) {
    NCCL_CK(nccl().Send(d_h + 64 * i, 64, ncclUint8, g->peers[i].rank, comm, st));
    NCCL_CK(nccl().Recv(d_h + 64 * (np + i), 64, ncclUint8, g->peers[i].rank, comm, st));
  }
  NCCL_CK(nccl().GroupEnd());
  if (np) GTK_CK(cudaMemcpyAsync(theirs.data(), d_h + 64 * np, 64 * np, cudaMemcpyDeviceToHost, st));
  GTK_CK(cudaStreamSynchronize(st));
  GTK_LAP("connect_peer_memory: export + swap handles");
  if (ok) for (size_t i = 0; i < np; ++i) {
    Peer& p = g->peers[i];
    if (p.remote_block) { cudaIpcCloseMemHandle(p.remote_block); p.remote_block = nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, theirs.data() + 64 * i, 64);
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    p.remote_block = (double*)ptr;
  }
  GTK_LAP("connect_peer_memory: cudaIpcOpenMemHandle");
  // agreement: one int per rank
  int *d_ok = nullptr, *d_all = nullptr;
  GTK_CK(gtk_cuda_malloc(ctx, &d_ok, sizeof(int)));
  GTK_CK(gtk_cuda_malloc(ctx, &d_all, sizeof(int) * W));
  GTK_CK(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
  NCCL_CK(nccl().AllGather(d_ok, d_all, 1, ncclInt32, comm, st));
  std::vector<int> all(W, 0);
  GTK_CK(cudaMemcpyAsync(all.data(), d_all, sizeof(int) * W, cudaMemcpyDeviceToHost, st));
  GTK_CK(cudaStreamSynchronize(st));
  gtk_cuda_free(ctx, d_h); gtk_cuda_free(ctx, d_ok); gtk_cuda_free(ctx, d_all);
  bool all_ok = true;
  for (int v : all) all_ok = all_ok && v != 0;
  if (g) g->p2p_off = !all_ok;
  GTK_LAP("connect_peer_memory: agreement");
  return GTK_OK;
}
