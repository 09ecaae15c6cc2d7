// Numeric phase, high-order path: element matrices as a batched dense GEMM on the FP64 tensor cores
// (DMMA.8x8x4, `mma.sync.aligned.m8n8k4.f64`), BASELINE config 3 (3D Poisson, Q3 hexahedra).
//
// What it replaces: the generated cell loop (compiler.jl:1865-1917) for the Laplacian integrand
//   be[r,c] = Σ_q (α ∇φ_r·∇φ_c) dV_q ,  ∇φ_i = Jᵀ \ ∇̂φ_i  (accessors.jl:1365-1368),  dV = sqrt(det JᵀJ) w_q
// (accessors.jl:1000-1007, quadrature.jl:4-6).  Written as linear algebra per cell:
//   Ke = Ĝᵀ · Y ,   Y[(q,a), j] = Σ_b C_q[a][b] Ĝ[(q,b), j] ,   C_q = α w_q adj(J_q) adj(J_q)ᵀ / |det J_q|
// with Ĝ[(q,a), i] = ∂_a φ̂_i(ξ_q) the tabulated reference gradients (accessors.jl:486-496) — THE SAME matrix for
// every cell.  So the whole mesh is one GEMM with M = n_ldofs, K = 3 n_q, N = n_ldofs · n_cells whose A operand is
// constant: each warp keeps its 8 rows of Ĝᵀ in REGISTERS for the life of the kernel (48 doubles for Q3) and only
// the B operand (Y, produced per cell from the 6 numbers per point of C_q) goes through shared memory.
//
//   k_cell_metric         : one thread per (cell, point): J = Σ_n x_n ⊗ ∇̂M_n (accessors.jl:941-968), C_q -> HBM
//                           (48 B per point; 0.8 GB at config 3, 5 % of the step's traffic)
//   k_elem_laplace_dmma   : persistent CTAs, one cell at a time, MT = n_ldofs/8 warps.  C_q of the NEXT cell arrives by
//                           a TMA bulk copy (cp.async.bulk + mbarrier) while this cell's GEMM runs.  Thread (warp w,
//                           lane = 4r+c) owns Ĝ[(q,·), i = 8w+r] for q ≡ c (mod 4): exactly its DMMA A fragments
//                           AND the operands it needs to produce column j = 8w+r of Y.  K is ordered
//                           k = 4·(3·(q/4) + a) + q%4 so that one k-step of 4 is one direction a of 4 consecutive points.
//                           Y lives in shared memory with row stride n_ldofs+4 doubles: the B-fragment loads (lane reads
//                           row c, column 8t+r) and the production stores are bank-conflict-free.
//                           Result tile D[i][j] goes straight from the accumulator fragments to the e-indexed staging
//                           array KE (slot c = i, r = j; 16-byte stores, full 32-byte sectors).
// The scatter (compress) is the generic fixed-order segmented sum k_reduce_nz of numeric.cu: no float atomics.
//
// Roofline: FP64 tensor pipe.  F_alg = n_cells · 2 · n_ldofs² · 3 n_q; measured DMMA peak on this
// GPU is 16 cycles per DMMA.8x8x4 per SM sub-partition = 37 TFLOP/s (profiles/r01_dmma_peak.txt) — the same rate
// as the DFMA pipe, so tensor cores here buy instruction-issue and register bandwidth, not a higher flop peak.
#include "gtk_internal.h"

int32_t gtk_reduce_nz_launch(gtk_ctx* ctx);   // numeric.cu

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
  double* C;          // [n_cells][nqp][6]
};

__global__ void __launch_bounds__(256) k_cell_metric(MetricArgs a) {
  const int64_t total = a.n_cells * a.nqp;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = t / a.nqp;
    const int q = (int)(t - cell * a.nqp);
    double c[6] = {0, 0, 0, 0, 0, 0};
    if (q < a.nq && cell >= a.act0 && cell < a.act1) {
      double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      const int32_t* nodes = a.cell_nodes + cell * a.nln;
      const double* dMq = a.dM + (size_t)q * a.nln * 3;
      for (int n = 0; n < a.nln; ++n) {   // local-node order, as accessors.jl:941-948
        const double* x = a.xyz + (size_t)(nodes[n] - 1) * 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) J[i][j] += x[i] * dMq[n * 3 + j];
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
      const double s = a.alpha * a.w[q] / fabs(det);
      // (J^{-1} J^{-T})[a][b] dV = s · Σ_k adj[a][k] adj[b][k]
      c[0] = s * (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]);
      c[1] = s * (A[0][0] * A[1][0] + A[0][1] * A[1][1] + A[0][2] * A[1][2]);
      c[2] = s * (A[0][0] * A[2][0] + A[0][1] * A[2][1] + A[0][2] * A[2][2]);
      c[3] = s * (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
      c[4] = s * (A[1][0] * A[2][0] + A[1][1] * A[2][1] + A[1][2] * A[2][2]);
      c[5] = s * (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
    }
    double2* out = reinterpret_cast<double2*>(a.C + (size_t)t * 6);
    out[0] = make_double2(c[0], c[1]);
    out[1] = make_double2(c[2], c[3]);
    out[2] = make_double2(c[4], c[5]);
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
  double* KE;         // [n_cells][nld][nld]
};

// MT: 8-row tiles of the element matrix (padded n_ldofs = 8 MT) = warps per CTA;  NG: groups of 4 quadrature points.
template <int MT, int NG>
struct GemmCfg {
  static constexpr int NLDP = 8 * MT;
  static constexpr int NQP = 4 * NG;
  static constexpr int KS = 3 * NG;          // k-steps of 4
  static constexpr int SJ = NLDP + 4;        // row stride of Y in doubles (≡ 4 mod 16 when NLDP ≡ 0 mod 16; see below)
  static constexpr int THREADS = 32 * MT;
  static constexpr size_t Y_BYTES = (size_t)4 * KS * SJ * sizeof(double);
  static constexpr size_t C_BYTES = (size_t)NQP * 6 * sizeof(double);
  static constexpr size_t SMEM = Y_BYTES + 2 * C_BYTES + 16;
};

template <int MT, int NG>
__global__ void __launch_bounds__(32 * MT, 1) k_elem_laplace_dmma(GemmArgs a) {
  using Cfg = GemmCfg<MT, NG>;
  constexpr int SJ = Cfg::SJ, KS = Cfg::KS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Ys = reinterpret_cast<double*>(smem_raw);
  double* Cs = reinterpret_cast<double*>(smem_raw + Cfg::Y_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + Cfg::Y_BYTES + 2 * Cfg::C_BYTES);

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
  const int nld = a.nld;
  const int iown = 8 * w + r;   // row of Ĝᵀ (A fragment) = column of Y this thread produces

  // A fragments: Ĝ[(q = 4g+c, a), i = iown], zero outside the element's real size
  double A[NG][3];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int q = 4 * g + c;
    const bool in = q < a.nq && iown < nld;
#pragma unroll
    for (int d = 0; d < 3; ++d) A[g][d] = in ? a.dN[((size_t)q * nld + iown) * 3 + d] : 0.0;
  }

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int64_t cell = blockIdx.x;
  if (threadIdx.x == 0 && cell < a.n_cells) {
    mbar_expect_tx(&bars[0], (uint32_t)Cfg::C_BYTES);
    tma_load_1d(Cs, a.C + (size_t)cell * Cfg::NQP * 6, (uint32_t)Cfg::C_BYTES, &bars[0]);
  }
  uint32_t it = 0;
  for (; cell < a.n_cells; cell += gridDim.x, ++it) {
    const int par = it & 1;
    const int64_t next = cell + gridDim.x;
    if (threadIdx.x == 0 && next < a.n_cells) {   // buffer par^1 was last read before the previous iteration's barriers
      mbar_expect_tx(&bars[par ^ 1], (uint32_t)Cfg::C_BYTES);
      tma_load_1d(Cs + (par ^ 1) * Cfg::NQP * 6, a.C + (size_t)next * Cfg::NQP * 6, (uint32_t)Cfg::C_BYTES, &bars[par ^ 1]);
    }
    mbar_wait(&bars[par], (it >> 1) & 1);

    // ---- produce Y[(q,a), j = iown] = Σ_b C_q[a][b] Ĝ[(q,b), j] for q ≡ c (mod 4) ----
    const double* Cq = Cs + par * Cfg::NQP * 6;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const double2* cp = reinterpret_cast<const double2*>(Cq + (4 * g + c) * 6);
      const double2 c01 = cp[0], c23 = cp[1], c45 = cp[2];   // xx xy | xz yy | yz zz
      const double g0 = A[g][0], g1 = A[g][1], g2 = A[g][2];
      const double y0 = c01.x * g0 + c01.y * g1 + c23.x * g2;
      const double y1 = c01.y * g0 + c23.y * g1 + c45.x * g2;
      const double y2 = c23.x * g0 + c45.x * g1 + c45.y * g2;
      double* yp = Ys + (size_t)(4 * (3 * g) + c) * SJ + iown;
      yp[0] = y0;
      yp[4 * SJ] = y1;
      yp[8 * SJ] = y2;
    }
    __syncthreads();

    // ---- Ke rows [8w, 8w+8) = Ĝᵀ Y : KS k-steps × MT column tiles of DMMA.8x8x4 ----
    double acc[MT][2];
#pragma unroll
    for (int t = 0; t < MT; ++t) acc[t][0] = acc[t][1] = 0.0;
    const double* yb = Ys + (size_t)c * SJ + r;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      const double af = A[s / 3][s % 3];
#pragma unroll
      for (int t = 0; t < MT; ++t) dmma884(acc[t][0], acc[t][1], af, yb[(size_t)(4 * s) * SJ + 8 * t]);
    }

    // ---- store: D[i = 8w+r][j = 8t+2c+{0,1}] -> KE[cell][c_slot = i][r_slot = j] ----
    double* out = a.KE + (size_t)cell * nld * nld + (size_t)iown * nld;
    if (nld == Cfg::NLDP) {
#pragma unroll
      for (int t = 0; t < MT; ++t)
        *reinterpret_cast<double2*>(out + 8 * t + 2 * c) = make_double2(acc[t][0], acc[t][1]);
    } else if (iown < nld) {
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int j = 8 * t + 2 * c;
        if (j < nld) out[j] = acc[t][0];
        if (j + 1 < nld) out[j + 1] = acc[t][1];
      }
    }
    __syncthreads();   // Y is rewritten by the next iteration
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
  if (form != GTK_FORM_LAPLACE || ctx->D != 3 || ctx->ncomp != 1) return GTK_OK;
  const int nld = ctx->nld, nq = ctx->nq;
  int mt, ng;
  if (nld <= 16 && nq <= 12) { mt = 2; ng = 3; }
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
  ma.C = ctx->Cm;
  {
    int64_t total = ctx->n_cells * nqp;
    int64_t g = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    GtkProf pr_(ctx, "k_cell_metric");
    k_cell_metric<<<(int)(g > cap ? cap : g), 256, 0, ctx->stream>>>(ma);
  }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);

  GemmArgs ga;
  ga.dN = ctx->dN; ga.C = ctx->Cm; ga.n_cells = ctx->n_cells; ga.nld = nld; ga.nq = nq; ga.KE = ctx->KE;
  if (mt == 2) rc = launch_gemm<2, 3>(ctx, ga);
  else if (mt == 4) rc = launch_gemm<4, 7>(ctx, ga);
  else rc = launch_gemm<8, 16>(ctx, ga);
  if (rc) return rc;
  return gtk_reduce_nz_launch(ctx);
}
