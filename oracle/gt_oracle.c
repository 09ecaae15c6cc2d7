/* ORACLE (C) — test infrastructure only; parity pinned through oracle/gt_oracle.py (see its header): this file is
 * cross-checked against the numpy oracle entry by entry (tests/test_oracle_invariants.py).
 *
 * Plain-C restatement of the reference's CPU assembly path, used (a) to cross-check the numpy
 * oracle and (b) as the timed CPU baseline ("kind": "port") of bench.py.  Never linked into the
 * product.  Follows, in order (paths relative to /root/reference/src):
 *   1. allocate_matrix counting loop                       assembly.jl:119-153, 440-445, 501-510
 *   2. generated cell loop (face -> point -> dof_c -> dof_r) compiler.jl:1865-1900
 *        J = sum_node x (x) grad M                         accessors.jl:941-968
 *        dV = sqrt(det(J'J)) w                             accessors.jl:1000-1007, quadrature.jl:4-6
 *        grad N = J' \ grad_ref N  (per dof)               accessors.jl:1365-1368
 *      contribute! -> COO push (col outer, row inner, skip by sign)  assembly.jl:189-208, 545-556
 *   3. compress -> sparse(I,J,V,m,n): CSC, rows sorted, duplicates summed in input order,
 *      explicit zeros kept                                  assembly.jl:571-575
 *   4. assemble_vector: COO (I,V) + dense_vector            assembly.jl:175-187, 535-543, 558-569
 *   5. update_matrix! / update_vector! (gto_reassemble): reset!, rerun the loops into the SAME COO arrays, then
 *      compress! = PartitionedArrays.sparse_matrix!(A,V,cache): nzval .= 0; nzval[K[k]] += V[k] with the nz index K
 *      cached by the first compress (no sort)               problems.jl:276-285, 352-361; assembly.jl:577-588
 * Scalar Lagrange spaces, D = 2 or 3, forms LAPLACE(1) / MASS(2), source f = const.
 * The reference is single-threaded; `nthreads` > 1 parallelises only the cell loops (pthreads; this image has no libgomp),
 * writing each cell's triplets at the offset the serial loop would use, so results are identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static double det2(const double a[2][2]) { return a[0][0] * a[1][1] - a[0][1] * a[1][0]; }
static double det3(const double a[3][3]) {
  /* StaticArrays: x0 . (x1 x x2) over columns */
  double c0 = a[1][1] * a[2][2] - a[2][1] * a[1][2];
  double c1 = a[2][1] * a[0][2] - a[0][1] * a[2][2];
  double c2 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
  return a[0][0] * c0 + a[1][0] * c1 + a[2][0] * c2;
}

/* element matrix + vector of one cell; be[c*nld + r] column-major like Julia */
/* Vector-valued spaces (ncomp > 1): local dof = node*ncomp + comp (space.jl:1267-1271); the tabulations are those of the
 * nls = nld/ncomp scalar shape functions.  form 3 = isotropic elasticity sigma(eps(u)):eps(v) with Lame parameters
 * (lam, mu): u = s_a e_i, v = s_b e_j  ->  lam d_i s_a d_j s_b + mu d_j s_a d_i s_b + mu delta_ij grad s_a . grad s_b. */
static int g_ncomp = 1;
static double g_lam = 0.0, g_mu = 0.0;
void gto_set_vector_space(int ncomp, double lam, double mu) { g_ncomp = ncomp < 1 ? 1 : ncomp; g_lam = lam; g_mu = mu; }

static void cell_kernel(int D, const double* xyz, const int32_t* nodes, int nln, int nld_full, int nq,
                        const double* w, const double* N, const double* dN, const double* dM, int form,
                        double alpha, double fconst, double* be, double* bv, double* g /* nld*3 scratch */) {
  const int ncomp = g_ncomp, nld = nld_full / ncomp;   /* nld: scalar shape functions from here on */
  if (ncomp > 1) {
    for (int i = 0; i < nld_full * nld_full; ++i) be[i] = 0.0;
    if (bv) for (int i = 0; i < nld_full; ++i) bv[i] = 0.0;
  }
  for (int i = 0; i < nld * nld; ++i) be[i] = 0.0;
  if (bv) for (int i = 0; i < nld; ++i) bv[i] = 0.0;
  for (int q = 0; q < nq; ++q) {
    double dV;
    if (D == 2) {
      double J[2][2] = {{0, 0}, {0, 0}};
      for (int n = 0; n < nln; ++n) {
        const double* x = xyz + (size_t)(nodes[n] - 1) * 2;
        const double* m = dM + ((size_t)q * nln + n) * 2;
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) J[i][j] += x[i] * m[j];
      }
      double G[2][2];
      for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) G[i][j] = J[0][i] * J[0][j] + J[1][i] * J[1][j];
      dV = sqrt(det2(G)) * w[q];
      if (form == 1 || form == 3) {
        double a[2][2] = {{J[0][0], J[1][0]}, {J[0][1], J[1][1]}};  /* a = J' */
        double d = det2(a);
        for (int s = 0; s < nld; ++s) {
          const double* b = dN + ((size_t)q * nld + s) * 2;
          g[s * 3 + 0] = (a[1][1] * b[0] - a[0][1] * b[1]) / d;
          g[s * 3 + 1] = (a[0][0] * b[1] - a[1][0] * b[0]) / d;
        }
      }
    } else {
      double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int n = 0; n < nln; ++n) {
        const double* x = xyz + (size_t)(nodes[n] - 1) * 3;
        const double* m = dM + ((size_t)q * nln + n) * 3;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[i][j] += x[i] * m[j];
      }
      double G[3][3];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        G[i][j] = (J[0][i] * J[0][j] + J[1][i] * J[1][j]) + J[2][i] * J[2][j];
      dV = sqrt(det3(G)) * w[q];
      if (form == 1 || form == 3) {
        double a[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = J[j][i];
        double d = det3(a);
#define A_(i, j) a[(i)-1][(j)-1]
        for (int s = 0; s < nld; ++s) {
          const double* b = dN + ((size_t)q * nld + s) * 3;
          g[s * 3 + 0] = ((A_(2, 2) * A_(3, 3) - A_(2, 3) * A_(3, 2)) * b[0] + (A_(1, 3) * A_(3, 2) - A_(1, 2) * A_(3, 3)) * b[1] +
                          (A_(1, 2) * A_(2, 3) - A_(1, 3) * A_(2, 2)) * b[2]) / d;
          g[s * 3 + 1] = ((A_(2, 3) * A_(3, 1) - A_(2, 1) * A_(3, 3)) * b[0] + (A_(1, 1) * A_(3, 3) - A_(1, 3) * A_(3, 1)) * b[1] +
                          (A_(1, 3) * A_(2, 1) - A_(1, 1) * A_(2, 3)) * b[2]) / d;
          g[s * 3 + 2] = ((A_(2, 1) * A_(3, 2) - A_(2, 2) * A_(3, 1)) * b[0] + (A_(1, 2) * A_(3, 1) - A_(1, 1) * A_(3, 2)) * b[1] +
                          (A_(1, 1) * A_(2, 2) - A_(1, 2) * A_(2, 1)) * b[2]) / d;
        }
#undef A_
      }
    }
    if (ncomp > 1) {
      const int nf = nld_full;
      for (int c = 0; c < nf; ++c)
        for (int r = 0; r < nf; ++r) {
          const int a = r / ncomp, i = r % ncomp, b = c / ncomp, j = c % ncomp;
          double t;
          if (form == 3) {
            t = g_lam * (g[a * 3 + i] * g[b * 3 + j]) + g_mu * (g[a * 3 + j] * g[b * 3 + i]);
            if (i == j) {
              double dt = g[a * 3] * g[b * 3];
              for (int k = 1; k < D; ++k) dt += g[a * 3 + k] * g[b * 3 + k];
              t = t + g_mu * dt;
            }
          } else if (form == 1) {
            t = 0.0;
            if (i == j) { t = g[a * 3] * g[b * 3]; for (int k = 1; k < D; ++k) t += g[a * 3 + k] * g[b * 3 + k]; }
          } else {
            t = i == j ? N[q * nld + a] * N[q * nld + b] : 0.0;
          }
          be[c * nf + r] += (alpha * t) * dV;
        }
      if (bv) for (int i = 0; i < nf; ++i) bv[i] += (1.0 * (fconst * N[q * nld + i / ncomp])) * dV;
      continue;
    }
    for (int c = 0; c < nld; ++c)
      for (int r = 0; r < nld; ++r) {
        double v;
        if (form == 1) {
          double dt = g[r * 3] * g[c * 3];
          for (int k = 1; k < D; ++k) dt += g[r * 3 + k] * g[c * 3 + k];
          v = (alpha * dt) * dV;
        } else {
          v = (alpha * (N[q * nld + r] * N[q * nld + c])) * dV;
        }
        be[c * nld + r] += v;
      }
    if (bv) for (int i = 0; i < nld; ++i) bv[i] += (1.0 * (fconst * N[q * nld + i])) * dV;
  }
}

typedef struct {
  int D; const double* xyz; const int32_t* cell_nodes; int nln, nld; const int32_t* cell_dofs; int nq;
  const double *w, *N, *dN, *dM; int form; double alpha, fconst;
  const int64_t *off, *voff; int32_t *I, *Jc; double* V; int32_t* VI; double* VV;
  int64_t c0, c1;
} loop_args;

/* 2. cell loop + contribute! for cells [c0, c1) */
static void* loop_worker(void* p) {
  loop_args* a = (loop_args*)p;
  const int nld = a->nld;
  double* be = (double*)malloc(sizeof(double) * nld * nld);
  double* bv = (double*)malloc(sizeof(double) * nld);
  double* g = (double*)malloc(sizeof(double) * nld * 3);
  for (int64_t cell = a->c0; cell < a->c1; ++cell) {
    const int32_t* dofs = a->cell_dofs + cell * nld;
    cell_kernel(a->D, a->xyz, a->cell_nodes + cell * a->nln, a->nln, nld, a->nq, a->w, a->N, a->dN, a->dM, a->form,
                a->alpha, a->fconst, be, a->VI ? bv : NULL, g);
    int64_t k = a->off[cell];
    for (int j = 0; j < nld; ++j) {
      if (dofs[j] < 0) continue;
      for (int i = 0; i < nld; ++i) {
        if (dofs[i] < 0) continue;
        a->I[k] = dofs[i]; a->Jc[k] = dofs[j]; a->V[k] = be[j * nld + i];
        ++k;
      }
    }
    if (a->VI) {
      int64_t kv = a->voff[cell];
      for (int i = 0; i < nld; ++i) {
        if (dofs[i] < 0) continue;
        a->VI[kv] = dofs[i]; a->VV[kv] = bv[i];
        ++kv;
      }
    }
  }
  free(be); free(bv); free(g);
  return NULL;
}

/* Returns 0 on success; -1 if `cap` (capacity of rowval/nzval) is too small (nnz_out is set).
 * t_phase[0..3]: seconds spent in count / loop / compress / vector (may be NULL). */
int gto_assemble(int D, int64_t n_nodes, const double* xyz, int64_t n_cells, int nln, const int32_t* cell_nodes,
                 int nld, const int32_t* cell_dofs, int64_t n_free, int nq, const double* w, const double* N,
                 const double* dN, const double* dM, int form, double alpha, double fconst, int32_t* colptr,
                 int32_t* rowval, double* nzval, int64_t cap, int64_t* nnz_out, double* b, int nthreads,
                 double* t_phase, int64_t* K_out /* NULL or [N_coo]: 0-based nz position of every COO entry (the cache of
                 sparse_matrix(...; reuse=true)) */) {
  (void)n_nodes;
  double t0 = now_s(), t1, t2, t3;
  /* 1. counting loop */
  int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_cells + 1));
  off[0] = 0;
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    const int32_t* dofs = cell_dofs + cell * nld;
    int64_t n = 0;
    for (int j = 0; j < nld; ++j) {
      if (dofs[j] < 0) continue;
      for (int i = 0; i < nld; ++i) {
        if (dofs[i] < 0) continue;
        n += 1;
      }
    }
    off[cell + 1] = off[cell] + n;
  }
  const int64_t ncoo = off[n_cells];
  int32_t* I = (int32_t*)calloc((size_t)(ncoo ? ncoo : 1), sizeof(int32_t));
  int32_t* Jc = (int32_t*)calloc((size_t)(ncoo ? ncoo : 1), sizeof(int32_t));
  double* V = (double*)calloc((size_t)(ncoo ? ncoo : 1), sizeof(double));
  /* vector COO */
  int64_t* voff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_cells + 1));
  voff[0] = 0;
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    int64_t n = 0;
    for (int i = 0; i < nld; ++i) n += cell_dofs[cell * nld + i] > 0;
    voff[cell + 1] = voff[cell] + n;
  }
  int32_t* VI = (int32_t*)calloc((size_t)(voff[n_cells] ? voff[n_cells] : 1), sizeof(int32_t));
  double* VV = (double*)calloc((size_t)(voff[n_cells] ? voff[n_cells] : 1), sizeof(double));
  t1 = now_s();
  /* 2. cell loop + contribute! */
  {
    loop_args la = {D, xyz, cell_nodes, nln, nld, cell_dofs, nq, w, N, dN, dM, form, alpha, fconst, off, voff, I, Jc, V,
                    b ? VI : NULL, VV, 0, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads == 1 || n_cells < 1024) {
      la.c0 = 0; la.c1 = n_cells;
      loop_worker(&la);
    } else {
      pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
      loop_args* as = (loop_args*)malloc(sizeof(loop_args) * nthreads);
      for (int t = 0; t < nthreads; ++t) {
        as[t] = la;
        as[t].c0 = n_cells * t / nthreads;
        as[t].c1 = n_cells * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, loop_worker, &as[t]);
      }
      for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
      free(th); free(as);
    }
  }
  t2 = now_s();
  /* 3. sparse(I,J,V): stable counting sort by row, then stable counting sort by column
   *    (= CSR build + transpose of SparseArrays.sparse!), then combine runs left to right. */
  int64_t n = n_free;
  int64_t* cnt = (int64_t*)calloc((size_t)(n + 2), sizeof(int64_t));
  int64_t* p1 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncoo ? ncoo : 1));
  for (int64_t k = 0; k < ncoo; ++k) cnt[I[k] + 1]++;
  for (int64_t r = 0; r <= n; ++r) cnt[r + 1] += cnt[r];
  for (int64_t k = 0; k < ncoo; ++k) p1[cnt[I[k]]++] = k;
  memset(cnt, 0, sizeof(int64_t) * (size_t)(n + 2));
  int64_t* p2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncoo ? ncoo : 1));
  for (int64_t k = 0; k < ncoo; ++k) cnt[Jc[k] + 1]++;
  for (int64_t c = 0; c <= n; ++c) cnt[c + 1] += cnt[c];
  for (int64_t t = 0; t < ncoo; ++t) { int64_t k = p1[t]; p2[cnt[Jc[k]]++] = k; }
  int64_t nnz = 0;
  int rc = 0;
  int64_t* colcnt = (int64_t*)calloc((size_t)(n + 1), sizeof(int64_t));
  int32_t pi = -1, pj = -1;
  for (int64_t t = 0; t < ncoo; ++t) {
    int64_t k = p2[t];
    if (I[k] == pi && Jc[k] == pj) {
      if (nnz <= cap) nzval[nnz - 1] += V[k];   /* duplicate: summed in input order */
    } else {
      if (nnz < cap) { rowval[nnz] = I[k]; nzval[nnz] = V[k]; }
      colcnt[Jc[k] - 1]++;
      nnz++;
      pi = I[k]; pj = Jc[k];
    }
    if (K_out) K_out[k] = nnz - 1;
  }
  if (nnz > cap) rc = -1;
  colptr[0] = 1;
  for (int64_t c = 0; c < n; ++c) colptr[c + 1] = (int32_t)(colptr[c] + colcnt[c]);
  free(colcnt);
  *nnz_out = nnz;
  t3 = now_s();
  /* 4. dense_vector */
  if (b) {
    for (int64_t i = 0; i < n; ++i) b[i] = 0.0;
    for (int64_t k = 0; k < voff[n_cells]; ++k) b[VI[k] - 1] += VV[k];
  }
  if (t_phase) { double t4 = now_s(); t_phase[0] = t1 - t0; t_phase[1] = t2 - t1; t_phase[2] = t3 - t2; t_phase[3] = t4 - t3; }
  free(off); free(I); free(Jc); free(V); free(voff); free(VI); free(VV); free(cnt); free(p1); free(p2);
  return rc;
}

/* N_coo of the matrix (what allocate_matrix counts, assembly.jl:119-153) and of the vector */
void gto_count(int64_t n_cells, int nld, const int32_t* cell_dofs, int64_t* ncoo_matrix, int64_t* ncoo_vector) {
  int64_t nm = 0, nv = 0;
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    int64_t nf = 0;
    for (int i = 0; i < nld; ++i) nf += cell_dofs[cell * nld + i] > 0;
    nm += nf * nf; nv += nf;
  }
  *ncoo_matrix = nm; *ncoo_vector = nv;
}

/* 5. update_matrix! + update_vector! on a cached pattern: I, Jc, V, VI, VV are the allocation's COO arrays (kept between
 * calls like the reference's `alloc`), K the cached nz index.  t_phase[0..2]: loop / compress! / dense_vector! seconds. */
int gto_reassemble(int D, const double* xyz, int64_t n_cells, int nln, const int32_t* cell_nodes, int nld,
                   const int32_t* cell_dofs, int64_t n_free, int nq, const double* w, const double* N, const double* dN,
                   const double* dM, int form, double alpha, double fconst, int32_t* I, int32_t* Jc, double* V,
                   int32_t* VI, double* VV, const int64_t* K, int64_t nnz, double* nzval, double* b, int nthreads,
                   double* t_phase) {
  double t0 = now_s(), t1, t2;
  int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_cells + 1));
  int64_t* voff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_cells + 1));
  off[0] = 0; voff[0] = 0;
  for (int64_t cell = 0; cell < n_cells; ++cell) {   /* reset!: the push cursor restarts; offsets per cell as the serial loop */
    int64_t nf = 0;
    for (int i = 0; i < nld; ++i) nf += cell_dofs[cell * nld + i] > 0;
    off[cell + 1] = off[cell] + nf * nf;
    voff[cell + 1] = voff[cell] + nf;
  }
  loop_args la = {D, xyz, cell_nodes, nln, nld, cell_dofs, nq, w, N, dN, dM, form, alpha, fconst, off, voff, I, Jc, V,
                  b ? VI : NULL, VV, 0, 0};
  if (nthreads < 1) nthreads = 1;
  if (nthreads == 1 || n_cells < 1024) {
    la.c0 = 0; la.c1 = n_cells;
    loop_worker(&la);
  } else {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    loop_args* as = (loop_args*)malloc(sizeof(loop_args) * nthreads);
    for (int t = 0; t < nthreads; ++t) {
      as[t] = la;
      as[t].c0 = n_cells * t / nthreads;
      as[t].c1 = n_cells * (t + 1) / nthreads;
      pthread_create(&th[t], NULL, loop_worker, &as[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th); free(as);
  }
  t1 = now_s();
  /* compress!(alloc, A, cache) -> sparse_matrix!(A, V, K) */
  const int64_t ncoo = off[n_cells];
  for (int64_t p = 0; p < nnz; ++p) nzval[p] = 0.0;
  for (int64_t k = 0; k < ncoo; ++k) nzval[K[k]] += V[k];
  t2 = now_s();
  if (b) {   /* dense_vector!(b, I, V) */
    for (int64_t i = 0; i < n_free; ++i) b[i] = 0.0;
    for (int64_t k = 0; k < voff[n_cells]; ++k) b[VI[k] - 1] += VV[k];
  }
  if (t_phase) { t_phase[0] = t1 - t0; t_phase[1] = t2 - t1; t_phase[2] = now_s() - t2; }
  free(off); free(voff);
  return 0;
}
