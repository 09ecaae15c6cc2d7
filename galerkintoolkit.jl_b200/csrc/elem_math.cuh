// Per-point arithmetic shared by the element kernels (numeric.cu, blocks.cu): the StaticArrays closed forms the reference
// evaluates (accessors.jl:941-968 J, quadrature.jl:4-6 change of measure, accessors.jl:1365-1368 physical gradients).
#pragma once
#include <cstdint>

namespace gtkmath {

template <int D>
__device__ __forceinline__ double det_mat(const double (&a)[D][D]) {
  if constexpr (D == 1) return a[0][0];
  if constexpr (D == 2) return a[0][0] * a[1][1] - a[0][1] * a[1][0];
  if constexpr (D == 3) {
    // StaticArrays: x0 . (x1 × x2) over columns
    double c0 = a[1][1] * a[2][2] - a[2][1] * a[1][2];
    double c1 = a[2][1] * a[0][2] - a[0][1] * a[2][2];
    double c2 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    return a[0][0] * c0 + a[1][0] * c1 + a[2][0] * c2;
  }
}

// sqrt(det(JᵀJ))  (quadrature.jl:4-6); J is D x d (d < D: boundary faces embedded in D dimensions)
template <int D, int d>
__device__ __forceinline__ double change_of_measure(const double (&J)[D][d]) {
  double G[d][d];
#pragma unroll
  for (int i = 0; i < d; ++i)
#pragma unroll
    for (int j = 0; j < d; ++j) {
      double s = J[0][i] * J[0][j];
#pragma unroll
      for (int k = 1; k < D; ++k) s += J[k][i] * J[k][j];
      G[i][j] = s;
    }
  return sqrt(det_mat<d>(G));
}

// g = a \ b with a = Jᵀ  (StaticArrays closed forms; accessors.jl:1365-1368)
template <int D>
__device__ __forceinline__ void solve_JT(const double (&J)[D][D], double d, const double* b, double* g) {
  if constexpr (D == 1) { g[0] = b[0] / J[0][0]; }
  if constexpr (D == 2) {
    // a[i][j] = J[j][i]
    g[0] = (J[1][1] * b[0] - J[1][0] * b[1]) / d;
    g[1] = (J[0][0] * b[1] - J[0][1] * b[0]) / d;
  }
  if constexpr (D == 3) {
#define A_(i, j) J[(j)-1][(i)-1]
    g[0] = ((A_(2, 2) * A_(3, 3) - A_(2, 3) * A_(3, 2)) * b[0] + (A_(1, 3) * A_(3, 2) - A_(1, 2) * A_(3, 3)) * b[1] +
            (A_(1, 2) * A_(2, 3) - A_(1, 3) * A_(2, 2)) * b[2]) / d;
    g[1] = ((A_(2, 3) * A_(3, 1) - A_(2, 1) * A_(3, 3)) * b[0] + (A_(1, 1) * A_(3, 3) - A_(1, 3) * A_(3, 1)) * b[1] +
            (A_(1, 3) * A_(2, 1) - A_(1, 1) * A_(2, 3)) * b[2]) / d;
    g[2] = ((A_(2, 1) * A_(3, 2) - A_(2, 2) * A_(3, 1)) * b[0] + (A_(1, 2) * A_(3, 1) - A_(1, 1) * A_(3, 2)) * b[1] +
            (A_(1, 1) * A_(2, 2) - A_(1, 2) * A_(2, 1)) * b[2]) / d;
#undef A_
  }
}

// J = Σ_node x_node ⊗ ∇̂M_node, sequential in local-node order (accessors.jl:941-948); nodes 1-based
template <int D, int d>
__device__ __forceinline__ void jacobian_from(const double* __restrict__ xyz, const int32_t* __restrict__ nodes, int nln,
                                              const double* __restrict__ dMq, double (&J)[D][d]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < d; ++j) J[i][j] = 0.0;
  for (int n = 0; n < nln; ++n) {
    const double* x = xyz + (size_t)(nodes[n] - 1) * D;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < d; ++j) J[i][j] += x[i] * dMq[n * d + j];
  }
}

}  // namespace gtkmath
