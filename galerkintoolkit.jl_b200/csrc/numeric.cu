// Numeric phase, generic path: any Lagrange element the host tabulates.
//
//   k_elem_matrix / k_elem_vector : the generated cell loop (compiler.jl:1865-1917, 1958-1990)
//       with the per-point arithmetic of accessors.jl (J :941-968, dV :1000-1007 +
//       quadrature.jl:4-6, ∇N = Jᵀ\∇̂N :1365-1368), two phases per CTA through shared memory:
//       A) one thread per (cell, point): geometry + physical gradients -> smem
//       B) one thread per element-matrix entry: Σ_q integrand·dV in the reference's q order,
//          written to the e-indexed staging array (coalesced, 8 B/thread consecutive).
//   k_reduce_nz / k_reduce_rows   : compress (assembly.jl:571-588): fixed-order segmented sum of
//       duplicates in reference push order (no atomics; bit-reproducible).
//
// The structured Q1 fast path lives in fastq1.cu and is tried first.
#include <algorithm>
#include "gtk_internal.h"
#include "q1hex_math.cuh"
#include "elem_math.cuh"

bool gtk_fastq1_tabulation_ok(const gtk_ctx* ctx);   // fastq1.cu: the tabulation is the Q1 / 2x2x2 Gauss one
int32_t gtk_fastq1_try(gtk_ctx* ctx, int mform, const gtk_form_params* pm, int vform,
                       const gtk_form_params* pv, bool* handled);
int32_t gtk_elemgemm_try(gtk_ctx* ctx, int form, const gtk_form_params* p, bool* handled);   // elemgemm.cu

namespace {

struct ElemArgs {
  const double* xyz;
  const int32_t* cell_nodes;
  int64_t n_cells;
  int nln, nls, ncomp, nld, nq;
  const double *w, *N, *dN, *M, *dM;
  int form;
  double alpha, lambda, mu;
  double f_const[3];
  const double* f_ptr;
  const double* coef_ptr;   // coefficient κ of a bilinear form
  int coef_mode;            // 0 none, 1 nodal [n_nodes], 2 per quadrature point [n_cells][nq]
  double* out;
  int cb;
  int64_t act0, act1;   // numeric-active cells [act0, act1); others contribute zeros
  // DiscreteField parameter u_h (PLAPLACE_* forms, scalar integrals): gathered per cell by the sign of the dof id
  const int32_t* cell_dofs;
  const double *u_free, *u_diri;
  double expo;          // q of flux(∇u) = |∇u|^(q-2) ∇u
};

// x^e as Julia evaluates Float64^Int for the small integer exponents of the p-Laplacian tests (x^1 = x, x^0 = 1,
// x^-1 = inv(x)); pow() otherwise
__device__ __forceinline__ double powq(double x, double e) {
  if (e == 0.0) return 1.0;
  if (e == 1.0) return x;
  if (e == 2.0) return x * x;
  if (e == -1.0) return 1.0 / x;
  if (e == -2.0) { const double i = 1.0 / x; return i * i; }
  return pow(x, e);
}

__device__ __forceinline__ double field_value(const ElemArgs& a, int dof) {   // accessors.jl:1496-1508
  return dof > 0 ? a.u_free[dof - 1] : a.u_diri[-dof - 1];
}

using gtkmath::det_mat;
using gtkmath::change_of_measure;
using gtkmath::solve_JT;

template <int D, int d>
__device__ __forceinline__ void jacobian_at(const ElemArgs& a, int64_t cell, int q, double (&J)[D][d]) {
  gtkmath::jacobian_from<D, d>(a.xyz, a.cell_nodes + cell * a.nln, a.nln, a.dM + (size_t)q * a.nln * d, J);
}

template <int D, int d>
__global__ void __launch_bounds__(128) k_elem_matrix(ElemArgs a) {
  extern __shared__ double smem[];
  const int nq = a.nq, nls = a.nls, nld = a.nld, ncomp = a.ncomp;
  double* G = smem;                                  // [cb][nq][nls][D]
  double* dV = G + (size_t)a.cb * nq * nls * D;      // [cb][nq]
  double* FU = dV + (size_t)a.cb * nq;               // [cb][nq][D+2]: ∇u_h, (q-2)|∇u_h|^(q-4), |∇u_h|^(q-2)  (PLAPLACE_JACOBIAN)
  const int64_t cell0 = (int64_t)blockIdx.x * a.cb;
  const int ncb = (int)min((int64_t)a.cb, a.n_cells - cell0);
  const bool need_grad = a.form != GTK_FORM_MASS;
  for (int t = threadIdx.x; t < ncb * nq; t += blockDim.x) {
    int cl = t / nq, q = t - cl * nq;
    double J[D][d];
    jacobian_at<D, d>(a, cell0 + cl, q, J);
    double dv = change_of_measure<D, d>(J) * a.w[q];
    if (a.coef_mode == 2) dv *= a.coef_ptr[(cell0 + cl) * nq + q];
    else if (a.coef_mode == 1) {   // κ_q = Σ_node κ_node M_node(ξ_q), sequential in local-node order
      const int32_t* nodes = a.cell_nodes + (cell0 + cl) * a.nln;
      double kq = 0.0;
      for (int n = 0; n < a.nln; ++n) kq += a.coef_ptr[nodes[n] - 1] * a.M[q * a.nln + n];
      dv *= kq;
    }
    dV[t] = dv;
    if constexpr (D == d) if (need_grad) {
      double JT[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) JT[i][j] = J[j][i];
      double d = det_mat<D>(JT);
      double* g = G + (size_t)t * nls * D;
      const double* dNq = a.dN + (size_t)q * nls * D;
      for (int s = 0; s < nls; ++s) solve_JT<D>(J, d, dNq + s * D, g + s * D);
      if (a.form == GTK_FORM_PLAPLACE_JACOBIAN) {
        // ∇u_h = sum(i -> x[i]*s[i]; init = zero), sequential in local-dof order (accessors.jl:1549-1556)
        const int32_t* dofs = a.cell_dofs + (cell0 + cl) * nld;
        double gu[D];
#pragma unroll
        for (int k = 0; k < D; ++k) gu[k] = 0.0;
        for (int s = 0; s < nls; ++s) {
          const double us = field_value(a, dofs[s]);
#pragma unroll
          for (int k = 0; k < D; ++k) gu[k] += us * g[s * D + k];
        }
        double n2 = gu[0] * gu[0];
#pragma unroll
        for (int k = 1; k < D; ++k) n2 += gu[k] * gu[k];
        const double nrm = sqrt(n2);
        double* fu = FU + (size_t)t * (D + 2);
#pragma unroll
        for (int k = 0; k < D; ++k) fu[k] = gu[k];
        fu[D] = (a.expo - 2.0) * powq(nrm, a.expo - 4.0);
        fu[D + 1] = powq(nrm, a.expo - 2.0);
      }
    }
  }
  __syncthreads();
  const int nld2 = nld * nld;
  for (int t = threadIdx.x; t < ncb * nld2; t += blockDim.x) {
    int cl = t / nld2, rem = t - cl * nld2;
    int c = rem / nld, r = rem - c * nld;
    int ra = r / ncomp, ri = r - ra * ncomp;
    int ca = c / ncomp, cj = c - ca * ncomp;
    double acc = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double* gq = G + ((size_t)(cl * nq + q) * nls) * D;
      double v;
      if (a.form == GTK_FORM_LAPLACE) {
        double dt = 0.0;
        if (ri == cj) {
          dt = gq[ra * D] * gq[ca * D];
#pragma unroll
          for (int k = 1; k < D; ++k) dt += gq[ra * D + k] * gq[ca * D + k];
        }
        v = a.alpha * dt;
      } else if (a.form == GTK_FORM_MASS) {
        v = ri == cj ? a.alpha * (a.N[q * nls + ra] * a.N[q * nls + ca]) : 0.0;
      } else if (a.form == GTK_FORM_PLAPLACE_JACOBIAN) {
        // ∇v ⋅ ((q-2)*norm(∇u)^(q-4)*(∇u⋅∇du)*∇u + norm(∇u)^(q-2)*∇du), du = φ_r, v = φ_c, left to right
        const double* fu = FU + (size_t)(cl * nq + q) * (D + 2);
        double udu = fu[0] * gq[ra * D];
#pragma unroll
        for (int k = 1; k < D; ++k) udu += fu[k] * gq[ra * D + k];
        const double c1 = fu[D] * udu, c2 = fu[D + 1];
        double dt = gq[ca * D] * (c1 * fu[0] + c2 * gq[ra * D]);
#pragma unroll
        for (int k = 1; k < D; ++k) dt += gq[ca * D + k] * (c1 * fu[k] + c2 * gq[ra * D + k]);
        v = a.alpha * dt;
      } else {  // ELASTICITY_ISO: λ ∂_i s_a ∂_j s_b + μ ∂_j s_a ∂_i s_b + μ δ_ij ∇s_a·∇s_b
        double tt = a.lambda * (gq[ra * D + ri] * gq[ca * D + cj]) + a.mu * (gq[ra * D + cj] * gq[ca * D + ri]);
        if (ri == cj) {
          double dt = gq[ra * D] * gq[ca * D];
#pragma unroll
          for (int k = 1; k < D; ++k) dt += gq[ra * D + k] * gq[ca * D + k];
          tt += a.mu * dt;
        }
        v = a.alpha * tt;
      }
      acc += v * dV[cl * nq + q];
    }
    const int64_t cell = cell0 + cl;
    a.out[cell * (int64_t)nld2 + rem] = (cell >= a.act0 && cell < a.act1) ? acc : 0.0;
  }
}

template <int D, int d>
__global__ void __launch_bounds__(128) k_elem_vector(ElemArgs a) {
  extern __shared__ double smem[];
  const int nq = a.nq, nls = a.nls, nld = a.nld, ncomp = a.ncomp;
  double* dV = smem;                          // [cb][nq]
  double* F = dV + (size_t)a.cb * nq;         // [cb][nq][ncomp]
  const int64_t cell0 = (int64_t)blockIdx.x * a.cb;
  const int ncb = (int)min((int64_t)a.cb, a.n_cells - cell0);
  for (int t = threadIdx.x; t < ncb * nq; t += blockDim.x) {
    int cl = t / nq, q = t - cl * nq;
    int64_t cell = cell0 + cl;
    double J[D][d];
    jacobian_at<D, d>(a, cell, q, J);
    dV[t] = change_of_measure<D, d>(J) * a.w[q];
    for (int k = 0; k < ncomp; ++k) {
      double f;
      if (a.form == GTK_FORM_SOURCE_CONST) f = a.f_const[k];
      else if (a.form == GTK_FORM_SOURCE_QP) f = a.f_ptr[((size_t)cell * nq + q) * ncomp + k];
      else {  // nodal: Σ_node f_node M_node(ξ_q), sequential
        f = 0.0;
        const int32_t* nodes = a.cell_nodes + cell * a.nln;
        for (int n = 0; n < a.nln; ++n) f += a.f_ptr[(size_t)(nodes[n] - 1) * ncomp + k] * a.M[q * a.nln + n];
      }
      F[(size_t)t * ncomp + k] = f;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ncb * nld; t += blockDim.x) {
    int cl = t / nld, i = t - cl * nld;
    int ia = i / ncomp, ic = i - ia * ncomp;
    double acc = 0.0;
    for (int q = 0; q < nq; ++q)
      acc += (a.alpha * (F[(size_t)(cl * nq + q) * ncomp + ic] * a.N[q * nls + ia])) * dV[cl * nq + q];
    const int64_t cell = cell0 + cl;
    a.out[cell * (int64_t)nld + i] = (cell >= a.act0 && cell < a.act1) ? acc : 0.0;
  }
}

// Q1 hexahedra on an UNSTRUCTURED mesh (any cell order / node numbering: what a Gmsh mesh gives): one thread per cell
// computes the sum-factorised element matrix (36 unique entries, q1hex_math.cuh — the arithmetic of the structured sweep
// kernel) and the element vector and writes them to the staging arrays in one pass: replaces k_cell_metric +
// k_elem_laplace_dmma<1,2> + k_elem_vector (three kernels that each gather the cell's coordinates again).
__global__ void __launch_bounds__(128) k_q1hex_cells(const double* __restrict__ xyz, const int32_t* __restrict__ cell_nodes, int64_t n_cells,
                                                    int64_t act0, int64_t act1, double alpha, double fscale, int do_matrix, int do_vector,
                                                    double* __restrict__ KE, double* __restrict__ BE) {
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < n_cells; cell += (int64_t)gridDim.x * blockDim.x) {
    const bool active = cell >= act0 && cell < act1;
    double Ke[36], be[8];
#pragma unroll
    for (int e = 0; e < 36; ++e) Ke[e] = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) be[e] = 0.0;
    if (active) {
      const int4 n0 = __ldg(reinterpret_cast<const int4*>(cell_nodes + cell * 8));
      const int4 n1 = __ldg(reinterpret_cast<const int4*>(cell_nodes + cell * 8) + 1);
      const int nd[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
      double X[8][3];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double* x = xyz + (size_t)(nd[v] - 1) * 3;
        X[v][0] = __ldg(x); X[v][1] = __ldg(x + 1); X[v][2] = __ldg(x + 2);
      }
      q1hex::Cell<double> g;
      q1hex::geometry<double>(X, g);
      if (do_matrix) q1hex::laplace_ke<double>(g, alpha, Ke);
      if (do_vector) q1hex::source_be<double>(g, fscale, be);
    }
    if (do_matrix) {
      double2* out = reinterpret_cast<double2*>(KE + cell * 64);   // e = c * 8 + r, symmetric
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int r = 0; r < 8; r += 2) out[(c * 8 + r) / 2] = make_double2(Ke[q1hex::sym(r, c)], Ke[q1hex::sym(r + 1, c)]);
    }
    if (do_vector) {
      double2* ob = reinterpret_cast<double2*>(BE + cell * 8);
#pragma unroll
      for (int i = 0; i < 8; i += 2) ob[i / 2] = make_double2(be[i], be[i + 1]);
    }
  }
}

// K4: isotropic elasticity on 3D vector-valued elements with up to 10 scalar shape functions (P1 / P2
// tetrahedra, 3 components: 30 x 30 element matrices) — BASELINE config 4.  One WARP per cell, no block barrier:
//   A) lanes over (point, shape function): J_q = Σ x⊗∇̂M (local-node order), dV_q = sqrt(det JᵀJ) w_q, ∇s_a = Jᵀ\∇̂s_a
//      (the closed forms of the generic kernel, so both kernels agree bit for bit) -> the warp's shared memory;
//   B) lanes over the 55 node pairs a <= b: the 3 x 3 block  Σ_q (α (λ ∂_i s_a ∂_j s_b + μ ∂_j s_a ∂_i s_b + δ_ij μ ∇s_a·∇s_b)) dV_q
//      in the reference's q order; the block of (b, a) is its transpose BITWISE (products commute), so it is mirrored;
//   C) the 900 entries leave shared memory as one contiguous, coalesced 7.2 kB store into the staging array.
// Compute is ~35 kFLOP per cell, the kernel is bound by the staging write (HBM).
constexpr int K4_MAXS = 10, K4_MAXQ = 14;
constexpr int K4_WARP_D = K4_MAXQ * K4_MAXS * 3 + K4_MAXQ * 12 + 9 * K4_MAXS * K4_MAXS;   // doubles of shared memory per warp
__global__ void __launch_bounds__(256) k_elem_elasticity_w(ElemArgs a) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nq = a.nq, nls = a.nls, nld = 3 * nls, nld2 = nld * nld;
  double* G = smem + (size_t)w * K4_WARP_D;    // [nq][nls][3]
  double* JQ = G + K4_MAXQ * K4_MAXS * 3;      // [nq][12]: J (9), det Jᵀ, dV
  double* Ke = JQ + K4_MAXQ * 12;              // [nld][nld] column-major
  const int npair = nls * (nls + 1) / 2;
  // this lane's node pairs (ra <= ca), fixed for the whole kernel: pair p = lane and lane + 32
  int pra[2], pca[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h;
    int ca = 0, off = 0;
    while (off + ca + 1 <= p) { off += ca + 1; ++ca; }
    pra[h] = p - off; pca[h] = ca;
  }
  for (int64_t cell = (int64_t)blockIdx.x * nwarp + w; cell < a.n_cells; cell += (int64_t)gridDim.x * nwarp) {
    const bool active = cell >= a.act0 && cell < a.act1;
    if (active) {
      if (lane < nq) {   // geometry once per point
        double J[3][3];
        jacobian_at<3, 3>(a, cell, lane, J);
        double JT[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) { JT[i][j] = J[j][i]; JQ[lane * 12 + 3 * i + j] = J[i][j]; }
        JQ[lane * 12 + 9] = 1.0 / det_mat<3>(JT);   // ONE division per point; the gradients multiply by it (1 ulp from the reference's per-component division)
        JQ[lane * 12 + 10] = change_of_measure<3, 3>(J) * a.w[lane];
      }
      __syncwarp();
      for (int t = lane; t < nq * nls; t += 32) {
        const int q = t / nls;
        double J[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) J[i][j] = JQ[q * 12 + 3 * i + j];
        {
          const double* b = a.dN + (size_t)t * 3;
          double* g = G + (size_t)t * 3;
          const double rd = JQ[q * 12 + 9];
#define A_(i, j) J[(j)-1][(i)-1]
          g[0] = ((A_(2, 2) * A_(3, 3) - A_(2, 3) * A_(3, 2)) * b[0] + (A_(1, 3) * A_(3, 2) - A_(1, 2) * A_(3, 3)) * b[1] +
                  (A_(1, 2) * A_(2, 3) - A_(1, 3) * A_(2, 2)) * b[2]) * rd;
          g[1] = ((A_(2, 3) * A_(3, 1) - A_(2, 1) * A_(3, 3)) * b[0] + (A_(1, 1) * A_(3, 3) - A_(1, 3) * A_(3, 1)) * b[1] +
                  (A_(1, 3) * A_(2, 1) - A_(1, 1) * A_(2, 3)) * b[2]) * rd;
          g[2] = ((A_(2, 1) * A_(3, 2) - A_(2, 2) * A_(3, 1)) * b[0] + (A_(1, 2) * A_(3, 1) - A_(1, 1) * A_(3, 2)) * b[1] +
                  (A_(1, 1) * A_(2, 2) - A_(1, 2) * A_(2, 1)) * b[2]) * rd;
#undef A_
        }
      }
    }
    __syncwarp();
    if (active) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (lane + 32 * h >= npair) break;
        const int ra = pra[h], ca = pca[h];
        double acc[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) acc[i][j] = 0.0;
        for (int q = 0; q < nq; ++q) {
          const double* ga = G + ((size_t)q * nls + ra) * 3;
          const double* gb = G + ((size_t)q * nls + ca) * 3;
          const double dv = JQ[q * 12 + 10];
          const double ga0 = ga[0], ga1 = ga[1], ga2 = ga[2], gb0 = gb[0], gb1 = gb[1], gb2 = gb[2];
          // (α dV) folded into ∇s_a once per point: 33 FP64 instructions per point and pair instead of 55 (the kernel is
          // FP64-issue bound); rounding differs from the reference's ((α t) dV) order by ~1e-16 relative
          const double sdv = a.alpha * dv;
          const double A[3] = {ga0 * sdv, ga1 * sdv, ga2 * sdv};
          const double lA[3] = {a.lambda * A[0], a.lambda * A[1], a.lambda * A[2]};
          const double mA[3] = {a.mu * A[0], a.mu * A[1], a.mu * A[2]};
          const double gbb[3] = {gb0, gb1, gb2};
          double dtm = mA[0] * gb0;
          dtm = fma(mA[1], gb1, dtm);
          dtm = fma(mA[2], gb2, dtm);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              acc[i][j] = fma(lA[i], gbb[j], acc[i][j]);
              acc[i][j] = fma(mA[j], gbb[i], acc[i][j]);
              if (i == j) acc[i][j] += dtm;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            Ke[(3 * ca + j) * nld + 3 * ra + i] = acc[i][j];      // row (ra, i), column (ca, j)
            Ke[(3 * ra + i) * nld + 3 * ca + j] = acc[i][j];      // its mirror: row (ca, j), column (ra, i)
          }
      }
    }
    __syncwarp();
    double* out = a.out + cell * (int64_t)nld2;
    if ((nld2 & 1) == 0) {   // 16-byte stores (cell * nld2 is even whenever nld2 is)
      double2* o2 = reinterpret_cast<double2*>(out);
      const double2* k2 = reinterpret_cast<const double2*>(Ke);
      for (int e = lane; e < nld2 / 2; e += 32) o2[e] = active ? k2[e] : make_double2(0.0, 0.0);
    } else {
      for (int e = lane; e < nld2; e += 32) out[e] = active ? Ke[e] : 0.0;
    }
    __syncwarp();
  }
}

// Residual of the p-Laplacian about u_h (GTK_FORM_PLAPLACE_RESIDUAL):
// be[i] = Σ_q (α·(∇φ_i⋅flux(∇u_h) − f φ_i))·dV_q, flux(∇u) = |∇u|^(q-2) ∇u  (test/problems_ext_tests.jl:160-162)
template <int D>
__global__ void __launch_bounds__(128) k_elem_vector_field(ElemArgs a) {
  extern __shared__ double smem[];
  const int nq = a.nq, nls = a.nls, nld = a.nld;
  double* G = smem;                                  // [cb][nq][nls][D]
  double* dV = G + (size_t)a.cb * nq * nls * D;      // [cb][nq]
  double* FL = dV + (size_t)a.cb * nq;               // [cb][nq][D+1]: flux(∇u_h), f
  const int64_t cell0 = (int64_t)blockIdx.x * a.cb;
  const int ncb = (int)min((int64_t)a.cb, a.n_cells - cell0);
  for (int t = threadIdx.x; t < ncb * nq; t += blockDim.x) {
    int cl = t / nq, q = t - cl * nq;
    const int64_t cell = cell0 + cl;
    double J[D][D];
    jacobian_at<D, D>(a, cell, q, J);
    dV[t] = change_of_measure<D, D>(J) * a.w[q];
    double JT[D][D];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) JT[i][j] = J[j][i];
    const double dt = det_mat<D>(JT);
    double* g = G + (size_t)t * nls * D;
    const double* dNq = a.dN + (size_t)q * nls * D;
    const int32_t* dofs = a.cell_dofs + cell * nld;
    double gu[D];
#pragma unroll
    for (int k = 0; k < D; ++k) gu[k] = 0.0;
    for (int s = 0; s < nls; ++s) {
      solve_JT<D>(J, dt, dNq + s * D, g + s * D);
      const double us = field_value(a, dofs[s]);
#pragma unroll
      for (int k = 0; k < D; ++k) gu[k] += us * g[s * D + k];
    }
    double n2 = gu[0] * gu[0];
#pragma unroll
    for (int k = 1; k < D; ++k) n2 += gu[k] * gu[k];
    const double c = powq(sqrt(n2), a.expo - 2.0);
    double* fl = FL + (size_t)t * (D + 1);
#pragma unroll
    for (int k = 0; k < D; ++k) fl[k] = c * gu[k];
    fl[D] = a.f_ptr ? a.f_ptr[(size_t)cell * nq + q] : a.f_const[0];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ncb * nld; t += blockDim.x) {
    int cl = t / nld, i = t - cl * nld;
    double acc = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double* gq = G + ((size_t)(cl * nq + q) * nls + i) * D;
      const double* fl = FL + (size_t)(cl * nq + q) * (D + 1);
      double dt = gq[0] * fl[0];
#pragma unroll
      for (int k = 1; k < D; ++k) dt += gq[k] * fl[k];
      acc += (a.alpha * (dt - fl[D] * a.N[q * nls + i])) * dV[cl * nq + q];
    }
    const int64_t cell = cell0 + cl;
    a.out[cell * (int64_t)nld + i] = (cell >= a.act0 && cell < a.act1) ? acc : 0.0;
  }
}

// Scalar integrals of u_h (assemble_scalar, problems.jl:173-190): one thread per (cell, point) evaluates integrand·dV, the
// block sums its values in a fixed tree and writes one partial; k_sum_partials adds the partials in a fixed tree.
template <int D>
__global__ void __launch_bounds__(256) k_elem_scalar(ElemArgs a, int kind, double* __restrict__ part) {
  __shared__ double red[256];
  const int nq = a.nq, nls = a.nls, nld = a.nld;
  const int64_t n_pts = (a.act1 - a.act0) * nq;
  double acc = 0.0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_pts; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = a.act0 + t / nq;
    const int q = (int)(t % nq);
    double J[D][D];
    jacobian_at<D, D>(a, cell, q, J);
    const double dv = change_of_measure<D, D>(J) * a.w[q];
    double val = 1.0;
    const int32_t* dofs = a.cell_dofs + cell * nld;
    if (kind == GTK_SCALAR_L2SQ) {
      double u = 0.0;
      for (int s = 0; s < nls; ++s) u += field_value(a, dofs[s]) * a.N[q * nls + s];
      if (a.f_ptr) u -= a.f_ptr[(size_t)cell * nq + q];
      val = u * u;
    } else if (kind == GTK_SCALAR_H1SQ) {
      double JT[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) JT[i][j] = J[j][i];
      const double dt = det_mat<D>(JT);
      const double* dNq = a.dN + (size_t)q * nls * D;
      double gu[D];
#pragma unroll
      for (int k = 0; k < D; ++k) gu[k] = 0.0;
      for (int s = 0; s < nls; ++s) {
        double g[D];
        solve_JT<D>(J, dt, dNq + s * D, g);
        const double us = field_value(a, dofs[s]);
#pragma unroll
        for (int k = 0; k < D; ++k) gu[k] += us * g[k];
      }
      if (a.f_ptr) {
#pragma unroll
        for (int k = 0; k < D; ++k) gu[k] -= a.f_ptr[((size_t)cell * nq + q) * D + k];
      }
      val = gu[0] * gu[0];
#pragma unroll
      for (int k = 1; k < D; ++k) val += gu[k] * gu[k];
    }
    acc += val * dv;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += part[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// nzval[p] = Σ_{s in segment p} KE[perm[s]], left to right = reference push order.
__global__ void k_reduce_nz(const double* __restrict__ KE, const uint32_t* __restrict__ perm,
                            const uint32_t* __restrict__ nzptr, int64_t nnz, double* __restrict__ nzval) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz;
       p += (int64_t)gridDim.x * blockDim.x) {
    uint32_t s0 = nzptr[p], s1 = nzptr[p + 1];
    double acc = KE[perm[s0]];
    for (uint32_t s = s0 + 1; s < s1; ++s) acc += KE[perm[s]];
    nzval[p] = acc;
  }
}

__global__ void k_reduce_rows(const double* __restrict__ BE, const uint32_t* __restrict__ perm,
                              const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ urow,
                              int64_t n_urows, double* __restrict__ b, int accumulate) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n_urows;
       u += (int64_t)gridDim.x * blockDim.x) {
    uint32_t s0 = rowptr[u], s1 = rowptr[u + 1];
    // dense_vector starts from zeros (assembly.jl:562); with `accumulate` the COO entries of this integral follow the
    // ones already summed into b (one COO vector for all contributions of a sum of integrals, problems.jl:258-266)
    double acc = (accumulate ? b[urow[u]] : 0.0) + BE[perm[s0]];
    for (uint32_t s = s0 + 1; s < s1; ++s) acc += BE[perm[s]];
    b[urow[u]] = acc;
  }
}

inline int grid_for(int64_t n, int block, int sm) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int32_t ensure(gtk_ctx* ctx, double** p, size_t* cap, size_t n) {
  if (*cap >= n && *p) return GTK_OK;
  if (*p) gtk_dev_free(ctx, *p, *cap * sizeof(double));
  *p = nullptr; *cap = 0;
  if (n == 0) n = 1;
  int32_t rc = gtk_dev_alloc(ctx, (void**)p, n * sizeof(double));
  if (rc == GTK_OK) *cap = n;
  return rc;
}

int32_t fill_args(gtk_ctx* ctx, ElemArgs& a, int form, const gtk_form_params* p) {
  if (!ctx->xyz || !ctx->cell_dofs || !ctx->w || ctx->nq <= 0)
    GTK_FAIL(GTK_ERR_STATE, "mesh, space and tabulation must be set first (in this order; a new mesh or element invalidates the tabulation)");
  if (ctx->parts)
    GTK_FAIL(GTK_ERR_STATE, "this context holds the parts of a product space / skeleton integral (gtk_set_parts): use gtk_matrix_numeric_blocks / gtk_vector_assemble_blocks");
  a.xyz = ctx->xyz; a.cell_nodes = ctx->cell_nodes; a.n_cells = ctx->n_cells;
  a.nln = ctx->nln; a.nls = ctx->nls; a.ncomp = ctx->ncomp; a.nld = ctx->nld; a.nq = ctx->nq;
  a.w = ctx->w; a.N = ctx->N; a.dN = ctx->dN; a.M = ctx->M; a.dM = ctx->dM;
  a.form = form;
  a.alpha = p ? p->alpha : 1.0;
  a.lambda = p ? p->lambda : 0.0;
  a.mu = p ? p->mu : 0.0;
  for (int k = 0; k < 3; ++k) a.f_const[k] = p ? p->f_const[k] : 0.0;
  a.f_ptr = nullptr;
  a.coef_ptr = nullptr; a.coef_mode = 0;
  a.act0 = ctx->act_count < 0 ? 0 : ctx->act_first;
  a.act1 = ctx->act_count < 0 ? ctx->n_cells : ctx->act_first + ctx->act_count;
  a.cell_dofs = ctx->cell_dofs; a.u_free = ctx->u_free; a.u_diri = ctx->u_diri;
  a.expo = p ? p->exponent : 0.0;
  return GTK_OK;
}

// forms that read the DiscreteField parameter: scalar space, volume cells, values resident (zero until set)
int32_t field_form_check(gtk_ctx* ctx, ElemArgs& a, const char* what) {
  if (ctx->ncomp != 1) GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, std::string(what) + " needs a scalar space; no CPU fallback");
  if (ctx->dman != ctx->D) GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, std::string(what) + " is not available on cells of lower dimension than the space");
  int32_t rc = gtk_field_ensure(ctx);
  if (rc) return rc;
  a.u_free = ctx->u_free; a.u_diri = ctx->u_diri;
  return GTK_OK;
}

int32_t pick_cb(gtk_ctx* ctx, size_t per_cell_bytes, int* cb, size_t* smem) {
  size_t soft = 96 * 1024;
  int c = (int)(soft / per_cell_bytes);
  if (c > 32) c = 32;
  if (c < 1) c = (int)((ctx->smem_optin - 2048) / per_cell_bytes);
  if (c < 1) GTK_FAIL(GTK_ERR_TOO_LARGE, "element too large for the generic shared-memory kernel");
  *cb = c;
  *smem = c * per_cell_bytes;
  return GTK_OK;
}

}  // namespace

// nzval[p] for the nonzeros with several contributions only (the single-contribution ones were written by the producer)
__global__ void k_reduce_multi(const double* __restrict__ KE, const uint32_t* __restrict__ perm,
                               const uint32_t* __restrict__ nzptr, const uint32_t* __restrict__ multi, int64_t n_multi,
                               double* __restrict__ nzval) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_multi; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = multi[i];
    uint32_t s0 = nzptr[p], s1 = nzptr[p + 1];
    double acc = KE[perm[s0]];
    for (uint32_t s = s0 + 1; s < s1; ++s) acc += KE[perm[s]];   // reference push order, as k_reduce_nz
    nzval[p] = acc;
  }
}

int32_t gtk_reduce_multi_launch(gtk_ctx* ctx) {
  MatSym& m = ctx->ms;
  if (m.n_multi == 0) return GTK_OK;
  { GtkProf pr_(ctx, "k_reduce_multi"); k_reduce_multi<<<grid_for(m.n_multi, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->KE, m.perm, m.nzptr, m.multi, m.n_multi, ctx->nzval); }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return GTK_OK;
}

int32_t gtk_upload_coefficient(gtk_ctx* ctx, int form, const gtk_form_params* p, int* mode) {
  *mode = 0;
  if (!p || (!p->coef_nodal && !p->coef_qp)) return GTK_OK;
  if (p->coef_nodal && p->coef_qp) GTK_FAIL(GTK_ERR_INVALID, "give coef_nodal or coef_qp, not both");
  if (form != GTK_FORM_LAPLACE && form != GTK_FORM_MASS)
    GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "a scalar coefficient is supported for LAPLACE and MASS only; no CPU fallback");
  const size_t n = p->coef_nodal ? (size_t)ctx->n_nodes : (size_t)ctx->n_cells * ctx->nq;
  int32_t rc = ensure(ctx, &ctx->coef_dev, &ctx->coef_cap, n);
  if (rc) return rc;
  GTK_CK(cudaMemcpyAsync(ctx->coef_dev, p->coef_nodal ? p->coef_nodal : p->coef_qp, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // the host pointer is only borrowed for the duration of the call; with pinned memory the copy above is truly asynchronous
  GTK_CK(cudaStreamSynchronize(ctx->stream));
  *mode = p->coef_nodal ? 1 : 2;
  return GTK_OK;
}

// compress (assembly.jl:571-588) of the staged element matrices: shared by the generic and the DMMA path
int32_t gtk_reduce_nz_launch(gtk_ctx* ctx) {
  MatSym& m = ctx->ms;
  { GtkProf pr_(ctx, "k_reduce_nz"); k_reduce_nz<<<grid_for(m.nnz, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->KE, m.perm, m.nzptr, m.nnz, ctx->nzval); }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return GTK_OK;
}

int32_t gtk_reduce_rows_launch(gtk_ctx* ctx, int accumulate);

// Q1 hexahedra / 2x2x2 Gauss / LAPLACE (+ SOURCE_CONST) without lattice structure: the fused cell kernel, then the
// fixed-order reductions.  mform / vform == 0: that half is not wanted.
static int32_t q1cells_try(gtk_ctx* ctx, int mform, const gtk_form_params* pm, int vform, const gtk_form_params* pv, bool* handled) {
  *handled = false;
  if (getenv("GTK_DISABLE_Q1CELLS") || getenv("GTK_DISABLE_FASTPATH") || ctx->D != 3 || ctx->dman != 3 || ctx->nld != 8 || ctx->nln != 8 || ctx->ncomp != 1) return GTK_OK;
  if (mform && (mform != GTK_FORM_LAPLACE || (pm && (pm->coef_nodal || pm->coef_qp)))) return GTK_OK;
  if (vform && (vform != GTK_FORM_SOURCE_CONST || (pv && pv->accumulate))) return GTK_OK;
  if (!mform && !vform) return GTK_OK;
  if (!gtk_fastq1_tabulation_ok(ctx)) return GTK_OK;
  MatSym& m = ctx->ms;
  VecSym& v = ctx->vs;
  int32_t rc;
  if (mform) {
    if (!m.generic_plan && (rc = gtk_symbolic_generic_plan(ctx))) return rc;
    if ((rc = ensure(ctx, &ctx->KE, &ctx->KE_cap, (size_t)m.n_full))) return rc;
    if ((rc = ensure(ctx, &ctx->nzval, &ctx->nzval_cap, (size_t)m.nnz))) return rc;
  }
  if (vform) {
    if (!v.generic_plan && (rc = gtk_symbolic_vector_generic_plan(ctx))) return rc;
    if ((rc = ensure(ctx, &ctx->BE, &ctx->BE_cap, (size_t)v.n_full))) return rc;
    if ((rc = ensure(ctx, &ctx->bvec, &ctx->bvec_cap, (size_t)v.n_rows))) return rc;
    GTK_CK(cudaMemsetAsync(ctx->bvec, 0, sizeof(double) * (size_t)(v.n_rows > 0 ? v.n_rows : 1), ctx->stream));
  }
  *handled = true;
  ctx->fast_path_last = 5;
  if (ctx->n_cells == 0) return GTK_OK;
  const int64_t a0 = ctx->act_count < 0 ? 0 : ctx->act_first, a1 = ctx->act_count < 0 ? ctx->n_cells : ctx->act_first + ctx->act_count;
  { GtkProf pr_(ctx, "k_q1hex_cells");
    k_q1hex_cells<<<grid_for(ctx->n_cells, 128, ctx->sm_count), 128, 0, ctx->stream>>>(ctx->xyz, ctx->cell_nodes, ctx->n_cells, a0, a1, pm ? pm->alpha : 1.0,
                                                                                  pv ? pv->alpha * pv->f_const[0] : 0.0, mform != 0, vform != 0, ctx->KE, ctx->BE); }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  if (mform && m.nnz && (rc = gtk_reduce_nz_launch(ctx))) return rc;
  if (vform && v.n_urows && (rc = gtk_reduce_rows_launch(ctx, 0))) return rc;
  return GTK_OK;
}

int32_t gtk_numeric_matrix_generic(gtk_ctx* ctx, int form, const gtk_form_params* p) {
  MatSym& m = ctx->ms;
  {
    bool handled = false;
    int32_t rc = q1cells_try(ctx, form, p, 0, nullptr, &handled);
    if (rc || handled) return rc;
  }
  if (!m.generic_plan) {   // the structured symbolic phase deferred the sort-based plan
    int32_t rc = gtk_symbolic_generic_plan(ctx);
    if (rc) return rc;
  }
  {
    bool handled = false;
    int32_t rc = gtk_elemgemm_try(ctx, form, p, &handled);
    if (rc || handled) return rc;
  }
  if (form == GTK_FORM_ELASTICITY_ISO && ctx->ncomp != ctx->D)
    GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "ELASTICITY_ISO needs a vector space with n_comp == D");
  if (form != GTK_FORM_LAPLACE && form != GTK_FORM_MASS && form != GTK_FORM_ELASTICITY_ISO && form != GTK_FORM_PLAPLACE_JACOBIAN)
    GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "unsupported bilinear form id " + std::to_string(form) +
                                           " (supported: LAPLACE, MASS, ELASTICITY_ISO, PLAPLACE_JACOBIAN); no CPU fallback");
  ElemArgs a;
  int32_t rc = fill_args(ctx, a, form, p);
  if (rc) return rc;
  if (form == GTK_FORM_PLAPLACE_JACOBIAN && (rc = field_form_check(ctx, a, "PLAPLACE_JACOBIAN"))) return rc;
  if ((rc = gtk_upload_coefficient(ctx, form, p, &a.coef_mode))) return rc;
  a.coef_ptr = ctx->coef_dev;
  rc = ensure(ctx, &ctx->KE, &ctx->KE_cap, (size_t)m.n_full);
  if (rc) return rc;
  rc = ensure(ctx, &ctx->nzval, &ctx->nzval_cap, (size_t)m.nnz);
  if (rc) return rc;
  if (ctx->n_cells == 0 || m.nnz == 0) return GTK_OK;
  a.out = ctx->KE;
  const int D = ctx->D;
  if (form == GTK_FORM_ELASTICITY_ISO && D == 3 && ctx->dman == 3 && ctx->ncomp == 3 && ctx->nls <= K4_MAXS && ctx->nq <= K4_MAXQ &&
      !getenv("GTK_DISABLE_K4")) {
    const size_t smem = 8 * sizeof(double) * K4_WARP_D;
    GTK_CK(cudaFuncSetAttribute(k_elem_elasticity_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    GTK_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elem_elasticity_w, 256, smem));
    const int64_t blocks = std::min<int64_t>((ctx->n_cells + 7) / 8, (int64_t)ctx->sm_count * (occ < 1 ? 1 : occ));
    { GtkProf pr_(ctx, "k_elem_elasticity_w"); k_elem_elasticity_w<<<(int)blocks, 256, smem, ctx->stream>>>(a); }
    GTK_CK(cudaGetLastError());
    gtk_count_launch(ctx);
    ctx->fast_path_last = 4;
    return gtk_reduce_nz_launch(ctx);
  }
  size_t per_cell = ((size_t)ctx->nq * ctx->nls * D + ctx->nq + (size_t)ctx->nq * (D + 2)) * sizeof(double);
  size_t smem;
  rc = pick_cb(ctx, per_cell, &a.cb, &smem);
  if (rc) return rc;
  int grid = (int)((ctx->n_cells + a.cb - 1) / a.cb);
  cudaStream_t st = ctx->stream;
#define LAUNCH_M(DD, dd)                                                                                        \
  do {                                                                                                          \
    GTK_CK(cudaFuncSetAttribute(k_elem_matrix<DD, dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    { GtkProf pr_(ctx, "k_elem_matrix"); k_elem_matrix<DD, dd><<<grid, 128, smem, st>>>(a); }                    \
  } while (0)
  const int dm = ctx->dman;
  if (dm != D && form != GTK_FORM_MASS)
    GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "only MASS is supported on cells of lower dimension than the space (boundary faces); no CPU fallback");
  if (D == 1 && dm == 1) LAUNCH_M(1, 1); else if (D == 2 && dm == 2) LAUNCH_M(2, 2); else if (D == 3 && dm == 3) LAUNCH_M(3, 3);
  else if (D == 2 && dm == 1) LAUNCH_M(2, 1); else if (D == 3 && dm == 2) LAUNCH_M(3, 2); else if (D == 3 && dm == 1) LAUNCH_M(3, 1);
  else GTK_FAIL(GTK_ERR_INVALID, "D must be 1, 2 or 3 and 1 <= manifold dimension <= D");
#undef LAUNCH_M
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return gtk_reduce_nz_launch(ctx);
}

static int32_t upload_f(gtk_ctx* ctx, const double* host, size_t n) {
  if (!host) GTK_FAIL(GTK_ERR_INVALID, "source data pointer is null");
  int32_t rc = ensure(ctx, &ctx->f_dev, &ctx->f_cap, n);
  if (rc) return rc;
  GTK_CK(cudaMemcpyAsync(ctx->f_dev, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GTK_CK(cudaStreamSynchronize(ctx->stream));   // borrowed (possibly pinned) host buffer: do not return before it is read
  return GTK_OK;
}

int32_t gtk_numeric_vector_generic(gtk_ctx* ctx, int form, const gtk_form_params* p) {
  VecSym& v = ctx->vs;
  if (form != GTK_FORM_SOURCE_CONST && form != GTK_FORM_SOURCE_NODAL && form != GTK_FORM_SOURCE_QP && form != GTK_FORM_PLAPLACE_RESIDUAL)
    GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "unsupported linear form id " + std::to_string(form) +
                                           " (supported: SOURCE_CONST, SOURCE_NODAL, SOURCE_QP, PLAPLACE_RESIDUAL); no CPU fallback");
  if (!v.generic_plan) {
    int32_t rc = gtk_symbolic_vector_generic_plan(ctx);
    if (rc) return rc;
  }
  ElemArgs a;
  int32_t rc = fill_args(ctx, a, form, p);
  if (rc) return rc;
  if (form == GTK_FORM_SOURCE_NODAL) {
    rc = upload_f(ctx, p ? p->f_nodal : nullptr, (size_t)ctx->n_nodes * ctx->ncomp);
    if (rc) return rc;
    a.f_ptr = ctx->f_dev;
  } else if (form == GTK_FORM_SOURCE_QP) {
    rc = upload_f(ctx, p ? p->f_qp : nullptr, (size_t)ctx->n_cells * ctx->nq * ctx->ncomp);
    if (rc) return rc;
    a.f_ptr = ctx->f_dev;
  } else if (form == GTK_FORM_PLAPLACE_RESIDUAL) {
    if ((rc = field_form_check(ctx, a, "PLAPLACE_RESIDUAL"))) return rc;
    if (p && p->f_qp) {
      if ((rc = upload_f(ctx, p->f_qp, (size_t)ctx->n_cells * ctx->nq))) return rc;
      a.f_ptr = ctx->f_dev;
    }
  }
  rc = ensure(ctx, &ctx->BE, &ctx->BE_cap, (size_t)v.n_full);
  if (rc) return rc;
  rc = ensure(ctx, &ctx->bvec, &ctx->bvec_cap, (size_t)v.n_rows);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const int accumulate = p && p->accumulate ? 1 : 0;
  if (!accumulate) GTK_CK(cudaMemsetAsync(ctx->bvec, 0, sizeof(double) * (size_t)(v.n_rows > 0 ? v.n_rows : 1), st));
  if (ctx->n_cells == 0 || v.n_urows == 0) return GTK_OK;
  a.out = ctx->BE;
  const int D = ctx->D;
  size_t per_cell = ((size_t)ctx->nq * (1 + ctx->ncomp)) * sizeof(double);
  if (form == GTK_FORM_PLAPLACE_RESIDUAL) per_cell = ((size_t)ctx->nq * ctx->nls * D + ctx->nq + (size_t)ctx->nq * (D + 1)) * sizeof(double);
  size_t smem;
  rc = pick_cb(ctx, per_cell, &a.cb, &smem);
  if (rc) return rc;
  int grid = (int)((ctx->n_cells + a.cb - 1) / a.cb);
#define LAUNCH_VF(DD)                                                                                             \
  do {                                                                                                            \
    GTK_CK(cudaFuncSetAttribute(k_elem_vector_field<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    { GtkProf pr_(ctx, "k_elem_vector_field"); k_elem_vector_field<DD><<<grid, 128, smem, st>>>(a); }              \
  } while (0)
  if (form == GTK_FORM_PLAPLACE_RESIDUAL) {
    if (D == 1) LAUNCH_VF(1); else if (D == 2) LAUNCH_VF(2); else LAUNCH_VF(3);
  } else {
#define LAUNCH_V(DD, dd)                                                                                        \
  do {                                                                                                          \
    GTK_CK(cudaFuncSetAttribute(k_elem_vector<DD, dd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    { GtkProf pr_(ctx, "k_elem_vector"); k_elem_vector<DD, dd><<<grid, 128, smem, st>>>(a); }                    \
  } while (0)
  const int dm = ctx->dman;
  if (D == 1 && dm == 1) LAUNCH_V(1, 1); else if (D == 2 && dm == 2) LAUNCH_V(2, 2); else if (D == 3 && dm == 3) LAUNCH_V(3, 3);
  else if (D == 2 && dm == 1) LAUNCH_V(2, 1); else if (D == 3 && dm == 2) LAUNCH_V(3, 2); else if (D == 3 && dm == 1) LAUNCH_V(3, 1);
  else GTK_FAIL(GTK_ERR_INVALID, "D must be 1, 2 or 3 and 1 <= manifold dimension <= D");
#undef LAUNCH_V
  }
#undef LAUNCH_VF
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return gtk_reduce_rows_launch(ctx, accumulate);
}

int32_t gtk_reduce_rows_launch(gtk_ctx* ctx, int accumulate) {
  VecSym& v = ctx->vs;
  { GtkProf pr_(ctx, "k_reduce_rows"); k_reduce_rows<<<grid_for(v.n_urows, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->BE, v.perm, v.rowptr, v.urow,
                                                                         v.n_urows, ctx->bvec, accumulate); }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx);
  return GTK_OK;
}

int32_t gtk_numeric_matrix_impl(gtk_ctx* ctx, int form, const gtk_form_params* p) {
  if (ctx->parts) GTK_FAIL(GTK_ERR_STATE, "the context holds the parts of a product space / skeleton integral: use gtk_matrix_numeric_blocks / gtk_vector_assemble_blocks");
  if (!ctx->ms.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_symbolic must be called before gtk_matrix_numeric");
  ctx->launches_last = 0;
  ctx->fast_path_last = 0;
  gtk_prof_reset(ctx);
  bool handled = false;
  int32_t rc = gtk_fastq1_try(ctx, form, p, 0, nullptr, &handled);
  if (rc) return rc;
  if (handled) return GTK_OK;
  return gtk_numeric_matrix_generic(ctx, form, p);
}

int32_t gtk_numeric_vector_impl(gtk_ctx* ctx, int form, const gtk_form_params* p) {
  if (ctx->parts) GTK_FAIL(GTK_ERR_STATE, "the context holds the parts of a product space / skeleton integral: use gtk_matrix_numeric_blocks / gtk_vector_assemble_blocks");
  if (!ctx->vs.ready) {
    int32_t rc = gtk_symbolic_vector_impl(ctx, GTK_FREE);
    if (rc) return rc;
  }
  ctx->launches_last = 0;
  ctx->fast_path_last = 0;
  gtk_prof_reset(ctx);
  bool handled = false;
  int32_t rc = gtk_fastq1_try(ctx, 0, nullptr, form, p, &handled);
  if (rc) return rc;
  if (handled) return GTK_OK;
  return gtk_numeric_vector_generic(ctx, form, p);
}

int32_t gtk_numeric_both_impl(gtk_ctx* ctx, int mform, const gtk_form_params* pm, int vform,
                              const gtk_form_params* pv) {
  if (ctx->parts) GTK_FAIL(GTK_ERR_STATE, "the context holds the parts of a product space / skeleton integral: use gtk_matrix_numeric_blocks / gtk_vector_assemble_blocks");
  if (!ctx->ms.ready) GTK_FAIL(GTK_ERR_STATE, "gtk_matrix_symbolic must be called first");
  if (!ctx->vs.ready) {
    int32_t rc = gtk_symbolic_vector_impl(ctx, ctx->ms.rows_fd);
    if (rc) return rc;
  }
  ctx->launches_last = 0;
  ctx->fast_path_last = 0;
  gtk_prof_reset(ctx);
  bool handled = false;
  int32_t rc = gtk_fastq1_try(ctx, mform, pm, vform, pv, &handled);
  if (rc) return rc;
  if (handled) return GTK_OK;
  rc = q1cells_try(ctx, mform, pm, vform, pv, &handled);   // one cell pass for matrix + vector on unstructured Q1 hexahedra
  if (rc || handled) return rc;
  rc = gtk_numeric_matrix_generic(ctx, mform, pm);
  if (rc) return rc;
  return gtk_numeric_vector_generic(ctx, vform, pv);
}

int32_t gtk_scalar_impl(gtk_ctx* ctx, int kind, const gtk_form_params* p, double* out) {
  if (kind != GTK_SCALAR_VOLUME && kind != GTK_SCALAR_L2SQ && kind != GTK_SCALAR_H1SQ)
    GTK_FAIL(GTK_ERR_UNSUPPORTED_FORM, "unsupported scalar integral id " + std::to_string(kind) +
                                           " (supported: VOLUME, L2SQ, H1SQ); no CPU fallback");
  ElemArgs a;
  int32_t rc = fill_args(ctx, a, kind, p);
  if (rc) return rc;
  if ((rc = field_form_check(ctx, a, "a scalar integral of the discrete field"))) return rc;
  ctx->launches_last = 0;
  gtk_prof_reset(ctx);
  if (p && p->f_qp && kind != GTK_SCALAR_VOLUME) {
    if ((rc = upload_f(ctx, p->f_qp, (size_t)ctx->n_cells * ctx->nq * (kind == GTK_SCALAR_H1SQ ? ctx->D : 1)))) return rc;
    a.f_ptr = ctx->f_dev;
  }
  const int64_t n_pts = (a.act1 - a.act0) * ctx->nq;
  const int grid = grid_for(n_pts, 256, ctx->sm_count);
  if ((rc = ensure(ctx, &ctx->scal_part, &ctx->scal_part_cap, (size_t)ctx->sm_count * 16 + 1))) return rc;
  double* d_out = ctx->scal_part + (size_t)ctx->sm_count * 16;
  cudaStream_t st = ctx->stream;
  const int D = ctx->D;
  { GtkProf pr_(ctx, "k_elem_scalar");
    if (D == 1) k_elem_scalar<1><<<grid, 256, 0, st>>>(a, kind, ctx->scal_part);
    else if (D == 2) k_elem_scalar<2><<<grid, 256, 0, st>>>(a, kind, ctx->scal_part);
    else k_elem_scalar<3><<<grid, 256, 0, st>>>(a, kind, ctx->scal_part); }
  { GtkProf pr_(ctx, "k_sum_partials"); k_sum_partials<<<1, 256, 0, st>>>(ctx->scal_part, grid, d_out); }
  GTK_CK(cudaGetLastError());
  gtk_count_launch(ctx, 2);
  GTK_CK(cudaMemcpyAsync(out, d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
  GTK_CK(cudaStreamSynchronize(st));
  return GTK_OK;
}
