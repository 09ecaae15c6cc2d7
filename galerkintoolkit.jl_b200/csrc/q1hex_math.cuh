// Element arithmetic of the trilinear (Q1) hexahedron with the 2x2x2 Gauss rule, written once for
// device and host (the host instantiation is only used by tests/ to check the algebra on the CPU).
//
// What the reference computes per cell (compiler.jl:1879-1894 + accessors.jl:941-968, 1000-1007,
// 1365-1368):   Ke[r,c] = Σ_q (α ∇N_r·∇N_c) dV_q,  ∇N = J⁻ᵀ∇̂N,  dV = sqrt(det(JᵀJ)) w_q
//               be[i]   = Σ_q (α f N_i) dV_q
// Same numbers, far fewer FP64 operations (the FP64 pipe, not HBM, is what limits this kernel):
//   * the columns of J are bilinear in the two *other* reference coordinates, so each takes only 4
//     values over the 8 points and is built by two lerp stages from the 12 edge vectors;
//   * with r_0 = c1×c2, r_1 = c2×c0, r_2 = c0×c1 (J = [c0 c1 c2]):  det J = c0·r_0 and
//     ∇N_i·∇N_j dV = Σ_ab D_ab ∂̂_aN_i ∂̂_bN_j with the symmetric D_ab = w (r_a·r_b)/|det J|;
//   * ∂̂_aN_i are tensor products of 1-D values {a,b} and signs ±1, so Σ_q D_ab(q) ∂̂_aN_i ∂̂_bN_j is
//     contracted one direction at a time (sum factorisation): ≈600 instead of ≈1600 operations.
// Differences from the reference's operation order are O(1e-16) relative (tests: ≤1e-12).
#pragma once
#if defined(__CUDACC__)
#define GTK_HD __host__ __device__ __forceinline__
#else
#define GTK_HD inline
#endif

namespace q1hex {

// Gauss points on [0,1]: g0 = A, g1 = B.  n_0(g0) = B, n_1(g0) = A, n_0(g1) = A, n_1(g1) = B.
constexpr double GA = 0.21132486540518713;   // (1 - 1/sqrt(3)) / 2
constexpr double GB = 0.78867513459481287;   // 1 - GA
constexpr double PAA = GA * GA, PAB = GA * GB, PBB = GB * GB;
constexpr double W8 = 0.125;                  // weight of every point

// symmetric index of (r,c), r<=c, row-major upper triangle of an 8x8
GTK_HD constexpr int sym(int r, int c) { return r <= c ? r * 8 - (r * (r - 1)) / 2 + (c - r) : c * 8 - (c * (c - 1)) / 2 + (r - c); }

// value of n_k at Gauss point t
GTK_HD constexpr double nval(int k, int t) { return (k == t) ? GB : GA; }
// P_p(t) = n_k(t) n_l(t) with p = k + l
GTK_HD constexpr double pval(int p, int t) { return p == 1 ? PAB : ((p == 0) == (t == 0) ? PBB : PAA); }

// 1/x for normal positive x: MUFU.RCP64H seed (~2^-20 relative) + two Newton steps (-> ~1 ulp, not correctly
// rounded).  The IEEE division the compiler emits costs ~14 instructions with a slow-path call; parity with the
// reference is a 1e-12 tolerance, and the result is a pure function of x (bit-reproducible).
template <class T>
GTK_HD T fast_rcp(T x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"((double)x));
  double e = fma(-(double)x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-(double)x, y, 1.0);
  y = fma(y, e, y);
  return T(y);
#else
  return T(1) / x;
#endif
}

template <class T>
struct Cell {
  T D[6][8];   // D_ab at the 8 points (point index q1 + 2 q2 + 4 q3); ab order 00,11,22,01,02,12
  T dV[8];     // |det J| w
};

// X[v][k]: coordinates of local node v = v1 + 2 v2 + 4 v3 (tensor order, cartesian_mesh.jl:233-240)
template <class T>
GTK_HD void geometry(const T (&X)[8][3], Cell<T>& g) {
  // edge vectors along each reference direction
  T J0[2][2][3], J1[2][2][3], J2[2][2][3];   // J0[q2][q3], J1[q1][q3], J2[q1][q2]
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    {  // direction 0: edges (0,v2,v3)->(1,v2,v3)
      T e00 = X[1][k] - X[0][k], e10 = X[3][k] - X[2][k], e01 = X[5][k] - X[4][k], e11 = X[7][k] - X[6][k];
      T d0 = e10 - e00, d1 = e11 - e01;                 // lerp over v2
      T t00 = e00 + GA * d0, t10 = e00 + GB * d0;       // t[q2][v3=0]
      T t01 = e01 + GA * d1, t11 = e01 + GB * d1;       // t[q2][v3=1]
      T f0 = t01 - t00, f1 = t11 - t10;                 // lerp over v3
      J0[0][0][k] = t00 + GA * f0; J0[0][1][k] = t00 + GB * f0;
      J0[1][0][k] = t10 + GA * f1; J0[1][1][k] = t10 + GB * f1;
    }
    {  // direction 1: edges (v1,0,v3)->(v1,1,v3)
      T e00 = X[2][k] - X[0][k], e10 = X[3][k] - X[1][k], e01 = X[6][k] - X[4][k], e11 = X[7][k] - X[5][k];
      T d0 = e10 - e00, d1 = e11 - e01;                 // lerp over v1
      T t00 = e00 + GA * d0, t10 = e00 + GB * d0;
      T t01 = e01 + GA * d1, t11 = e01 + GB * d1;
      T f0 = t01 - t00, f1 = t11 - t10;                 // lerp over v3
      J1[0][0][k] = t00 + GA * f0; J1[0][1][k] = t00 + GB * f0;
      J1[1][0][k] = t10 + GA * f1; J1[1][1][k] = t10 + GB * f1;
    }
    {  // direction 2: edges (v1,v2,0)->(v1,v2,1)
      T e00 = X[4][k] - X[0][k], e10 = X[5][k] - X[1][k], e01 = X[6][k] - X[2][k], e11 = X[7][k] - X[3][k];
      T d0 = e10 - e00, d1 = e11 - e01;                 // lerp over v1
      T t00 = e00 + GA * d0, t10 = e00 + GB * d0;
      T t01 = e01 + GA * d1, t11 = e01 + GB * d1;
      T f0 = t01 - t00, f1 = t11 - t10;                 // lerp over v2
      J2[0][0][k] = t00 + GA * f0; J2[0][1][k] = t00 + GB * f0;
      J2[1][0][k] = t10 + GA * f1; J2[1][1][k] = t10 + GB * f1;
    }
  }
#pragma unroll
  for (int q3 = 0; q3 < 2; ++q3)
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2)
#pragma unroll
      for (int q1 = 0; q1 < 2; ++q1) {
        const int q = q1 + 2 * q2 + 4 * q3;
        const T* c0 = J0[q2][q3];
        const T* c1 = J1[q1][q3];
        const T* c2 = J2[q1][q2];
        T r0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
        T r1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
        T r2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
        T det = c0[0] * r0[0] + c0[1] * r0[1] + c0[2] * r0[2];
        T ad = det < T(0) ? -det : det;
        T s = T(W8) * fast_rcp<T>(ad);
        g.dV[q] = T(W8) * ad;
        g.D[0][q] = s * (r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
        g.D[1][q] = s * (r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
        g.D[2][q] = s * (r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
        g.D[3][q] = s * (r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2]);
        g.D[4][q] = s * (r0[0] * r2[0] + r0[1] * r2[1] + r0[2] * r2[2]);
        g.D[5][q] = s * (r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2]);
      }
}

// contraction helpers over one direction: values at the two Gauss points v0, v1
template <class T> GTK_HD T cn(int k, T v0, T v1) { return k == 0 ? GB * v0 + GA * v1 : GA * v0 + GB * v1; }        // Σ_t n_k(t) v_t
template <class T> GTK_HD T cp(int p, T v0, T v1) { return p == 1 ? PAB * (v0 + v1) : (p == 0 ? PBB * v0 + PAA * v1 : PAA * v0 + PBB * v1); }  // Σ_t P_p(t) v_t

// Ke (36 unique entries, index sym(r,c)) += alpha * Laplacian element matrix
template <class T>
GTK_HD void laplace_ke(const Cell<T>& g, T alpha, T (&Ke)[36]) {
#pragma unroll
  for (int i = 0; i < 36; ++i) Ke[i] = T(0);
  // ---- diagonal terms: H[p_u][p_v] over the two directions other than a ----
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // directions (u,v) = the other two, u < v; point index strides
    const int sa = 1 << a;
    const int u = a == 0 ? 1 : 0, v = a == 2 ? 1 : 2;
    const int su = 1 << u, sv = 1 << v;
    T E[2][2];   // summed over direction a: E[tu][tv]
#pragma unroll
    for (int tu = 0; tu < 2; ++tu)
#pragma unroll
      for (int tv = 0; tv < 2; ++tv) E[tu][tv] = g.D[a][tu * su + tv * sv] + g.D[a][tu * su + tv * sv + sa];
    T F[3][2];   // contracted over u: F[pu][tv]
#pragma unroll
    for (int pu = 0; pu < 3; ++pu)
#pragma unroll
      for (int tv = 0; tv < 2; ++tv) F[pu][tv] = cp<T>(pu, E[0][tv], E[1][tv]);
    T H[3][3];
#pragma unroll
    for (int pu = 0; pu < 3; ++pu)
#pragma unroll
      for (int pv = 0; pv < 3; ++pv) H[pu][pv] = alpha * cp<T>(pv, F[pu][0], F[pu][1]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = i; j < 8; ++j) {
        const int ia = (i >> a) & 1, ja = (j >> a) & 1;
        const int pu = ((i >> u) & 1) + ((j >> u) & 1), pv = ((i >> v) & 1) + ((j >> v) & 1);
        if (ia == ja) Ke[sym(i, j)] += H[pu][pv]; else Ke[sym(i, j)] -= H[pu][pv];
      }
  }
  // ---- mixed terms (a,b), a<b, c = the remaining direction ----
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const int a = m == 2 ? 1 : 0, b = m == 0 ? 1 : 2, c = 3 - a - b;
    const int sa = 1 << a, sb = 1 << b, sc = 1 << c;
    // G[ka][kb][pc] = Σ_q D_ab n_ka(q_a) n_kb(q_b) P_pc(q_c)
    //   product ∂̂_aN_i ∂̂_bN_j = s_{i_a} s_{j_b} n_{j_a}(q_a) n_{i_b}(q_b) P_{i_c+j_c}(q_c)
    T A1[2][2][2];   // [ka][tb][tc], contracted over a
#pragma unroll
    for (int ka = 0; ka < 2; ++ka)
#pragma unroll
      for (int tb = 0; tb < 2; ++tb)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc) A1[ka][tb][tc] = cn<T>(ka, g.D[3 + m][tb * sb + tc * sc], g.D[3 + m][tb * sb + tc * sc + sa]);
    T A2[2][2][2];   // [ka][kb][tc], contracted over b
#pragma unroll
    for (int ka = 0; ka < 2; ++ka)
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc) A2[ka][kb][tc] = cn<T>(kb, A1[ka][0][tc], A1[ka][1][tc]);
    T G[2][2][3];
#pragma unroll
    for (int ka = 0; ka < 2; ++ka)
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int pc = 0; pc < 3; ++pc) G[ka][kb][pc] = alpha * cp<T>(pc, A2[ka][kb][0], A2[ka][kb][1]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = i; j < 8; ++j) {
        const int ia = (i >> a) & 1, ib = (i >> b) & 1, ic = (i >> c) & 1;
        const int ja = (j >> a) & 1, jb = (j >> b) & 1, jc = (j >> c) & 1;
        // term 1: ∂̂_aN_i ∂̂_bN_j  sign s_{ia} s_{jb};  term 2: ∂̂_bN_i ∂̂_aN_j  sign s_{ib} s_{ja}
        if (ia == jb) Ke[sym(i, j)] += G[ja][ib][ic + jc]; else Ke[sym(i, j)] -= G[ja][ib][ic + jc];
        if (ib == ja) Ke[sym(i, j)] += G[ia][jb][ic + jc]; else Ke[sym(i, j)] -= G[ia][jb][ic + jc];
      }
  }
}

// be[i] = scale * Σ_q N_i(q) dV_q      (scale = α f)
template <class T>
GTK_HD void source_be(const Cell<T>& g, T scale, T (&be)[8]) {
  T A1[2][2][2], A2[2][2][2];
#pragma unroll
  for (int k1 = 0; k1 < 2; ++k1)
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
      for (int t3 = 0; t3 < 2; ++t3) A1[k1][t2][t3] = cn<T>(k1, g.dV[2 * t2 + 4 * t3], g.dV[1 + 2 * t2 + 4 * t3]);
#pragma unroll
  for (int k1 = 0; k1 < 2; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2)
#pragma unroll
      for (int t3 = 0; t3 < 2; ++t3) A2[k1][k2][t3] = cn<T>(k2, A1[k1][0][t3], A1[k1][1][t3]);
#pragma unroll
  for (int i = 0; i < 8; ++i) be[i] = scale * cn<T>((i >> 2) & 1, A2[i & 1][(i >> 1) & 1][0], A2[i & 1][(i >> 1) & 1][1]);
}

}  // namespace q1hex
