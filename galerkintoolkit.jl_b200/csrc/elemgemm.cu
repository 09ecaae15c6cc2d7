// Numeric phase, high-order path: element matrices as a batched dense GEMM on the FP64 tensor cores
// (DMMA.8x8x4, `mma.sync.aligned.m8n8k4.f64`), BASELINE config 3 (3D Poisson, Q3 hexahedra).
//
// What it replaces: the generated cell loop (compiler.jl:1865-1917) for the Laplacian integrand
//   be[r,c] = Σ_q (α ∇φ_r·∇φ_c) dV_q ,  ∇φ_i = Jᵀ \ ∇̂φ_i  (accessors.jl:1365-1368),  dV = sqrt(det JᵀJ) w_q
// (accessors.jl:1000-1007, quadrature.jl:4-6).  Written as linear algebra per cell:
//   Ke = Ĝᵀ · Y ,   Y[(q,a), j] = Σ_b C_q[a][b] Ĝ[(q,b), j] ,   C_q = α w_q adj(J_q) adj(J_q)ᵀ / |det J_q|
// with Ĝ[(q,a), i] = ∂_a φ̂_i(ξ_q) the tabulated reference gradients (accessors.jl:486-496) — THE SAME matrix for
// every cell.  So the whole mesh is one GEMM with M = n_ldofs, K = 3 n_q, N = n_ldofs · n_cells whose A operand is
// constant: its DMMA fragments sit in a table in shared memory for the life of the kernel, and the B operand (Y) is
// produced in REGISTERS by the very lanes that need it as fragments, from the 6 numbers per point of C_q — nothing but
// C_q (TMA) is read per cell, nothing but the finished tiles is written.  Ke is symmetric: 36 of the 64 tiles are computed.
//
//   k_cell_metric         : one thread per (cell, point): J = Σ_n x_n ⊗ ∇̂M_n (accessors.jl:941-968), C_q -> HBM
//                           (48 B per point; 0.8 GB at config 3, 5 % of the step's traffic)
//   k_elem_laplace_dmma   : persistent CTAs of 2·(n_ldofs/8) independent warps, see the comment at the kernel: K is ordered
//                           k = 4·(3·(q/4) + a) + q%4 so that one k-step of 4 is one direction a of 4 consecutive points;
//                           lane (r, c) of the warp owning column block cb computes Y[(q,·), 8cb+r] for q ≡ c (mod 4),
//                           which is exactly its B fragment of the three k-steps of that point group.  Result tiles go
//                           straight from the accumulator fragments to the e-indexed staging array KE (and mirrored).
// The scatter (compress) is the generic fixed-order segmented sum k_reduce_nz of numeric.cu: no float atomics.
//
// Roofline: FP64 tensor pipe.  F_alg = n_cells · 2 · n_ldofs² · 3 n_q; measured DMMA peak on this
// GPU is 16 cycles per DMMA.8x8x4 per SM sub-partition = 37 TFLOP/s (profiles/r01_dmma_peak.txt) — the same rate
// as the DFMA pipe, so tensor cores here buy instruction-issue and register bandwidth, not a higher flop peak.
#include "gtk_internal.h"

int32_t gtk_reduce_nz_launch(gtk_ctx* ctx);      // numeric.cu
int32_t gtk_reduce_multi_launch(gtk_ctx* ctx);   // numeric.cu

namespace {

struct MetricArgs {
  const double* xyz;
  const int32_t* cell_nodes;
  const double* dM;   // [nq][nln][3]
  const double* w;    // [nq]
  int64_t n_cells;
  int nln, nq, nqp;   // nqp = padded points per cell in the output
  double alpha;
  int64_t act0, act1;
  const double* coef; // coefficient κ: nullptr, nodal [n_nodes] (coef_mode 1) or per point [n_cells][nq] (2)
  const double* M;    // [nq][nln] geometry shape values (nodal coefficient)
  int coef_mode;
  double* C;          // [n_cells][nqp][6]
};

__global__ void __launch_bounds__(256) k_cell_metric(MetricArgs a) {
  __shared__ double2 sC[256 * 3];   // staged so that the 48-byte records leave as fully coalesced 16-byte stores
  extern __shared__ double sdM[];   // [nln*3][nq]: consecutive lanes (points) read consecutive addresses
  for (int i = threadIdx.x; i < a.nq * a.nln * 3; i += 256) {
    const int q = i / (a.nln * 3), nj = i - q * (a.nln * 3);
    sdM[nj * a.nq + q] = a.dM[i];
  }
  __syncthreads();
  const int64_t total = a.n_cells * a.nqp;
  for (int64_t base = blockIdx.x * (int64_t)256; base < total; base += (int64_t)gridDim.x * 256) {
    const int64_t t = base + threadIdx.x;
    const int64_t cell = t / a.nqp;
    const int q = (int)(t - cell * a.nqp);
    double c[6] = {0, 0, 0, 0, 0, 0};
    if (t < total && q < a.nq && cell >= a.act0 && cell < a.act1) {
      double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      const int32_t* nodes = a.cell_nodes + cell * a.nln;
      for (int n = 0; n < a.nln; ++n) {   // local-node order, as accessors.jl:941-948
        const double* x = a.xyz + (size_t)(nodes[n] - 1) * 3;
        const double x0 = x[0], x1 = x[1], x2 = x[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double dm = sdM[(n * 3 + j) * a.nq + q];
          J[0][j] += x0 * dm; J[1][j] += x1 * dm; J[2][j] += x2 * dm;
        }
      }
      // adj(J): J^{-1} = adj / det
      double A[3][3];
      A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
      A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
      A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
      A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
      A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
      double s = a.alpha * a.w[q] / fabs(det);
      if (a.coef_mode == 2) s *= a.coef[cell * a.nq + q];
      else if (a.coef_mode == 1) {
        double kq = 0.0;
        for (int n = 0; n < a.nln; ++n) kq += a.coef[nodes[n] - 1] * a.M[q * a.nln + n];
        s *= kq;
      }
      // (J^{-1} J^{-T})[a][b] dV = s · Σ_k adj[a][k] adj[b][k]
      c[0] = s * (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]);
      c[1] = s * (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]);
      c[2] = s * (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]);
      c[3] = s * (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
      c[4] = s * (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]);
      c[5] = s * (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
    }
    sC[threadIdx.x * 3 + 0] = make_double2(c[0], c[1]);
    sC[threadIdx.x * 3 + 1] = make_double2(c[2], c[3]);
    sC[threadIdx.x * 3 + 2] = make_double2(c[4], c[5]);
    __syncthreads();
    const int64_t left = total - base;
    const int n2 = (int)(left < 256 ? left : 256) * 3;
    double2* out = reinterpret_cast<double2*>(a.C + (size_t)base * 6);
    for (int i = threadIdx.x; i < n2; i += 256) out[i] = sC[i];
    __syncthreads();
  }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct GemmArgs {
  const double* dN;   // [nq][nld][3]
  const double* C;    // [n_cells][4*NG][6]
  int64_t n_cells;
  int nld, nq;
  double* KE;         // [n_cells][nld][nld] staging of the slots whose nonzero has several contributions
  const uint32_t* dest;   // [n_cells][nld][nld] MatSym::dest, or nullptr: stage everything
  double* nzval;
};

// one element-matrix entry -> its place: straight into nzval when it is the nonzero's only contribution
__device__ __forceinline__ void put_entry(const GemmArgs& a, size_t e, double v) {
  if (!a.dest) { a.KE[e] = v; return; }
  const uint32_t d = a.dest[e];
  if (d & 0x80000000u) { if (d != 0xFFFFFFFFu) a.nzval[d & 0x7FFFFFFFu] = v; }
  else a.KE[e] = v;
}

// MT: 8x8 tiles per side of the element matrix (padded n_ldofs = 8 MT);  NG: groups of 4 quadrature points.
// WPB warps share one column block (they take alternate cells), so a CTA has WPB·MT warps.
template <int MT, int NG>
struct GemmCfg {
  static constexpr int WPB = 2;
  static constexpr int NW = WPB * MT;
  static constexpr int NLDP = 8 * MT;
  static constexpr int NQP = 4 * NG;
  static constexpr int KS = 3 * NG;          // k-steps of 4
  static constexpr int THREADS = 32 * NW;
  static constexpr int NTW = MT / 2 + 1;     // tiles per warp (circulant split of the upper triangle)
  static constexpr int CD = NQP * 6;         // doubles of one cell's metric
  static constexpr size_t G_BYTES = (size_t)MT * KS * 32 * sizeof(double);   // fragment table of Ĝ
  static constexpr size_t C_BYTES = (size_t)CD * sizeof(double);
  static constexpr size_t SMEM = G_BYTES + (size_t)NW * 2 * C_BYTES + (size_t)NW * 2 * sizeof(uint64_t);
};

// A warp owns COLUMN block cb of Ke for every WPB-th cell of its CTA.  Its B fragments are the columns j = 8cb+r of Y,
// which lane (r, c) computes in registers from C_q (q ≡ c mod 4) and the 3 values Ĝ[(q,·), j] — Y never touches shared
// memory.  All fragments of Ĝ (A operands, and the production's inputs, which are the diagonal tile's A fragments) come
// from a constant table in shared memory, Gs[row block][k-step][lane], built once per CTA: conflict-free 8-byte loads.
// Ke is symmetric (C_q is), so only the upper triangle of tiles is computed: column block cb takes the row blocks
// (cb + d) mod MT, d = 0 .. MT/2, the last one only for cb < MT/2 — every unordered pair of blocks exactly once, 4 or 5
// tiles per warp for MT = 8.  Each off-diagonal tile is stored twice (as computed and transposed), which also makes the
// assembled matrix bitwise symmetric.
// Warps never synchronise with each other after the table is built: each streams through its cells on its own, fetching
// C_q two cells ahead with its own TMA bulk copies (cp.async.bulk + a private pair of mbarriers); 4 warps per SM
// sub-partition keep the DMMA pipe fed while others wait for operands or store.
template <int MT, int NG>
__global__ void __launch_bounds__(GemmCfg<MT, NG>::THREADS, 1) k_elem_laplace_dmma(GemmArgs a) {
  using Cfg = GemmCfg<MT, NG>;
  constexpr int NTW = Cfg::NTW, KS = Cfg::KS, CD = Cfg::CD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Gs = reinterpret_cast<double*>(smem_raw);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  const int cb = w % MT, sub = w / MT;   // column block, position among the warps sharing it
  double* Cw = reinterpret_cast<double*>(smem_raw + Cfg::G_BYTES) + (size_t)w * 2 * CD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + Cfg::G_BYTES + (size_t)Cfg::NW * 2 * Cfg::C_BYTES) + 2 * w;
  const int nld = a.nld;

  // fragment table: Gs[(m*KS + 3g + d)*32 + 4r + c] = Ĝ[(q = 4g+c, d), i = 8m + r], zero outside the element's real size
  for (int idx = threadIdx.x; idx < MT * KS * 32; idx += Cfg::THREADS) {
    const int l = idx & 31, s = (idx >> 5) % KS, m = idx / (32 * KS);
    const int q = 4 * (s / 3) + (l & 3), i = 8 * m + (l >> 2);
    Gs[idx] = (q < a.nq && i < nld) ? a.dN[((size_t)q * nld + i) * 3 + (s % 3)] : 0.0;
  }
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const double* gt[NTW];   // fragment base of this warp's d-th tile (row block (cb+d) % MT); d = 0 is its own block
#pragma unroll
  for (int d = 0; d < NTW; ++d) gt[d] = Gs + (size_t)(((cb + d) % MT) * KS) * 32 + lane;
  const bool last_tile = cb < MT / 2;   // warp-uniform

  // cells of this warp: first + k*stride, k = 0 .. K-1
  const int64_t stride = (int64_t)gridDim.x * Cfg::WPB, first = blockIdx.x + (int64_t)sub * gridDim.x;
  if (first >= a.n_cells) return;
  const int64_t K = (a.n_cells - first + stride - 1) / stride;
  auto issue = [&](int64_t k) {   // lane 0: C of the warp's k-th cell -> its buffer k&1
    const int b = (int)(k & 1);
    mbar_expect_tx(&bars[b], (uint32_t)Cfg::C_BYTES);
    tma_load_1d(Cw + b * CD, a.C + (size_t)(first + k * stride) * CD, (uint32_t)Cfg::C_BYTES, &bars[b]);
  };
  if (lane == 0) {
    issue(0);
    if (K > 1) issue(1);
  }
  const int jown = 8 * cb + 2 * c;   // first of the two Ke columns in this lane's accumulator fragments

  for (int64_t k = 0; k < K; ++k) {
    const double* Cq = Cw + (k & 1) * CD;
    mbar_wait(&bars[k & 1], (uint32_t)((k >> 1) & 1));

    double acc[NTW][2];
#pragma unroll
    for (int d = 0; d < NTW; ++d) acc[d][0] = acc[d][1] = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      // Y[(q,·), j] = C_q Ĝ[(q,·), j] for q = 4g + c: the B fragments of k-steps 3g .. 3g+2
      const double2* cp = reinterpret_cast<const double2*>(Cq + (4 * g + c) * 6);
      const double2 c01 = cp[0], c23 = cp[1], c45 = cp[2];   // xx xy | xz yy | yz zz
      const double g0 = gt[0][(3 * g) * 32], g1 = gt[0][(3 * g + 1) * 32], g2 = gt[0][(3 * g + 2) * 32];
      double y[3];
      y[0] = c01.x * g0 + c01.y * g1 + c23.x * g2;
      y[1] = c01.y * g0 + c23.y * g1 + c45.x * g2;
      y[2] = c23.x * g0 + c45.x * g1 + c45.y * g2;
#pragma unroll
      for (int d3 = 0; d3 < 3; ++d3) {
        const int s = 3 * g + d3;
        dmma884(acc[0][0], acc[0][1], d3 == 0 ? g0 : (d3 == 1 ? g1 : g2), y[d3]);
#pragma unroll
        for (int d = 1; d < NTW - 1; ++d) dmma884(acc[d][0], acc[d][1], gt[d][s * 32], y[d3]);
        if (NTW > 1 && last_tile) dmma884(acc[NTW - 1][0], acc[NTW - 1][1], gt[NTW - 1][s * 32], y[d3]);
      }
    }
    __syncwarp();
    if (lane == 0 && k + 2 < K) issue(k + 2);   // every lane is done reading buffer k&1

    // ---- store: D[i = 8m+r][j = jown+{0,1}] -> KE[cell][c_slot = i][r_slot = j] and its transpose ----
    const size_t e0 = (size_t)(first + k * stride) * nld * nld;
#pragma unroll
    for (int d = 0; d < NTW; ++d) {
      if (d == NTW - 1 && d > 0 && !last_tile) break;
      const int i = 8 * ((cb + d) % MT) + r;
      if (nld == Cfg::NLDP) {
        const size_t e = e0 + (size_t)i * nld + jown;
        if (!a.dest) {
          *reinterpret_cast<double2*>(a.KE + e) = make_double2(acc[d][0], acc[d][1]);
        } else {
          const uint2 dd = *reinterpret_cast<const uint2*>(a.dest + e);   // e is even: 8-byte aligned
          if (dd.x == 0u && dd.y == 0u) {
            *reinterpret_cast<double2*>(a.KE + e) = make_double2(acc[d][0], acc[d][1]);
          } else {
            if (dd.x & 0x80000000u) { if (dd.x != 0xFFFFFFFFu) a.nzval[dd.x & 0x7FFFFFFFu] = acc[d][0]; } else a.KE[e] = acc[d][0];
            if (dd.y & 0x80000000u) { if (dd.y != 0xFFFFFFFFu) a.nzval[dd.y & 0x7FFFFFFFu] = acc[d][1]; } else a.KE[e + 1] = acc[d][1];
          }
        }
        if (d > 0) {
          put_entry(a, e0 + (size_t)jown * nld + i, acc[d][0]);
          put_entry(a, e0 + (size_t)(jown + 1) * nld + i, acc[d][1]);
        }
      } else if (i < nld) {
        if (jown < nld) { put_entry(a, e0 + (size_t)i * nld + jown, acc[d][0]); if (d > 0) put_entry(a, e0 + (size_t)jown * nld + i, acc[d][0]); }
        if (jown + 1 < nld) { put_entry(a, e0 + (size_t)i * nld + jown + 1, acc[d][1]); if (d > 0) put_entry(a, e0 + (size_t)(jown + 1) * nld + i, acc[d][1]); }
      }
    }
  }
}

template <int MT, int NG>
int32_t launch_gemm(gtk_ctx* ctx, const GemmArgs& a) {
  using Cfg = GemmCfg<MT, NG>;
  auto kern = k_elem_laplace_dmma<MT, NG>;
  GTK_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  int per_sm = 0;
  GTK_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::THREADS, Cfg::SMEM));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)ctx->sm_count * per_sm;
  if (grid > a.n_cells) grid = a.n_cells;
  { GtkProf pr_(ctx, "k_elem_laplace_dmma"); kern<<<(int)grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(a); }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return GTK_OK;
}

}  // namespace

// Returns GTK_OK with *handled = false when the element/form is outside this path (the caller goes on to the generic
// kernel).  Eligible: D = 3, scalar space, LAPLACE, n_ldofs ≤ 64, n_q ≤ 64.
int32_t gtk_elemgemm_try(gtk_ctx* ctx, int form, const gtk_form_params* p, bool* handled) {
  *handled = false;
  if (getenv("GTK_DISABLE_DMMA")) return GTK_OK;
  if (form != GTK_FORM_LAPLACE || ctx->D != 3 || ctx->dman != 3 || ctx->ncomp != 1) return GTK_OK;
  const int nld = ctx->nld, nq = ctx->nq;
  int mt, ng;
  if (nld <= 8 && nq <= 8) { mt = 1; ng = 2; }
  else if (nld <= 16 && nq <= 12) { mt = 2; ng = 3; }
  else if (nld <= 32 && nq <= 28) { mt = 4; ng = 7; }
  else if (nld <= 64 && nq <= 64) { mt = 8; ng = 16; }
  else return GTK_OK;
  MatSym& m = ctx->ms;
  const int nqp = 4 * ng;
  int32_t rc;
  auto ensure = [&](double** ptr, size_t* cap, size_t n) -> int32_t {
    if (*cap >= n && *ptr) return GTK_OK;
    if (*ptr) gtk_dev_free(ctx, *ptr, *cap * sizeof(double));
    *ptr = nullptr; *cap = 0;
    int32_t r2 = gtk_dev_alloc(ctx, (void**)ptr, (n ? n : 1) * sizeof(double));
    if (r2 == GTK_OK) *cap = n ? n : 1;
    return r2;
  };
  // Opt-in experiment, measured SLOWER at config 3 (profiles/r01_q3_direct_write.txt): writing the 82 % single-contribution
  // entries straight into nzval saves their staging round trip (reduction 5.3 -> 3.3 ms) but turns the epilogue's 64-byte
  // row stores into 32-byte runs at unaligned places plus a dependent 4-byte index load per entry, and the GEMM kernel goes
  // from 8.2 to 12.6 ms.  Kept for the bitwise A/B test; default is staging everything.
  const bool direct = getenv("GTK_ENABLE_DIRECT_WRITE") != nullptr;
  if (direct && (rc = gtk_symbolic_direct_plan(ctx))) return rc;
  if ((rc = ensure(&ctx->KE, &ctx->KE_cap, (size_t)m.n_full))) return rc;
  if ((rc = ensure(&ctx->nzval, &ctx->nzval_cap, (size_t)m.nnz))) return rc;
  if ((rc = ensure(&ctx->Cm, &ctx->Cm_cap, (size_t)ctx->n_cells * nqp * 6))) return rc;
  *handled = true;
  ctx->fast_path_last = 3;
  if (ctx->n_cells == 0 || m.nnz == 0) return GTK_OK;

  MetricArgs ma;
  ma.xyz = ctx->xyz; ma.cell_nodes = ctx->cell_nodes; ma.dM = ctx->dM; ma.w = ctx->w;
  ma.n_cells = ctx->n_cells; ma.nln = ctx->nln; ma.nq = nq; ma.nqp = nqp;
  ma.alpha = p ? p->alpha : 1.0;
  ma.act0 = ctx->act_count < 0 ? 0 : ctx->act_first;
  ma.act1 = ctx->act_count < 0 ? ctx->n_cells : ctx->act_first + ctx->act_count;
  if ((rc = gtk_upload_coefficient(ctx, form, p, &ma.coef_mode))) return rc;
  ma.coef = ctx->coef_dev; ma.M = ctx->M;
  ma.C = ctx->Cm;
  {
    int64_t total = ctx->n_cells * nqp;
    int64_t g = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    const size_t dm_bytes = (size_t)nq * ctx->nln * 3 * sizeof(double);
    if (dm_bytes + 12288 > ctx->smem_optin) GTK_FAIL(GTK_ERR_TOO_LARGE, "geometry tabulation too large for k_cell_metric");
    GTK_CK(cudaFuncSetAttribute(k_cell_metric, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm_bytes));
    GtkProf pr_(ctx, "k_cell_metric");
    k_cell_metric<<<(int)(g > cap ? cap : g), 256, dm_bytes, ctx->stream>>>(ma);
  }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);

  GemmArgs ga;
  ga.dN = ctx->dN; ga.C = ctx->Cm; ga.n_cells = ctx->n_cells; ga.nld = nld; ga.nq = nq; ga.KE = ctx->KE;
  ga.dest = direct ? m.dest : nullptr; ga.nzval = ctx->nzval;
  if (mt == 1) rc = launch_gemm<1, 2>(ctx, ga);
  else if (mt == 2) rc = launch_gemm<2, 3>(ctx, ga);
  else if (mt == 4) rc = launch_gemm<4, 7>(ctx, ga);
  else rc = launch_gemm<8, 16>(ctx, ga);
  if (rc) return rc;
  return direct ? gtk_reduce_multi_launch(ctx) : gtk_reduce_nz_launch(ctx);
}
