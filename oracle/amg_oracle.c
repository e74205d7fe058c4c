/* amg_oracle.c -- CPU ORACLE for the additive-AMG solve phase.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's solve-phase loops.  Every function cites the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libamg_b200.so)
 * never links, imports or calls it.
 *
 * Parity status: PINNED against the reference's own object code -- oracle/_ref compiles the
 * unmodified reference translation units SMEM_MatVec.cpp, SMEM_Smooth.cpp, SMEM_Sync_AMG.cpp,
 * SMEM_Async_AMG.cpp, SMEM_ExtendedSystem.cpp, SMEM_Solve.cpp, SEQ_*.cpp, Misc.cpp, DMEM_Mult.cpp, DMEM_Misc.cpp (the DMEM
 * files for ONE rank) against a hypre / MPI stub shim and
 * tests/test_oracle_golden.py compares both on the same inputs; the outputs are also
 * committed as fixtures under tests/golden/.  The reference ships no golden vectors of its
 * own.  The hierarchy (A_l, P_l) is an INPUT here -- hypre's setup is an
 * un-vendored third-party dependency and is not restated.
 *
 * Semantics chosen where the reference is ambiguous:
 *   - additive cycles follow the race-free sequential specification SEQ_Add_Vcycle
 *     (src/SEQ_AMG.cpp:110-235), smoothers dispatched as SMEM_Smooth does
 *     (src/SMEM_Solve.cpp:264-377);
 *   - the coarsest level contributes nothing in SMEM additive cycles (5.9c);
 *   - hybrid Jacobi/Gauss-Seidel takes the block list as an input (5.9e).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef struct {
   int nrows, ncols, nnz;
   const int *i;
   const int *j;
   const double *data;
} orc_csr;

/* enums: src/Main.hpp:47-75 */
#define ORC_JACOBI 0
#define ORC_HYBRID_JGS 2
#define ORC_L1_JACOBI 6
#define ORC_L1_HYBRID_JGS 12   /* L1_HYBRID_JACOBI_GAUSS_SEIDEL: Parfor branch only, divisor = l1 norms (src/SMEM_Smooth.cpp:253-263) */
#define ORC_MULT 0
#define ORC_AFACX 1
#define ORC_MULTADD 2
#define ORC_BPX 3

typedef struct {
   int num_levels;
   const orc_csr *A;        /* [L]   diag-first */
   const orc_csr *P;        /* [L-1] (smoothed for MULTADD) */
   const orc_csr *R;        /* [L-1] explicit restriction */
   const double *const *l1; /* [L] L1 row norms or NULL */
   int solver, smoother;
   double smooth_weight;
   int num_pre, num_post, fine_sweeps, coarse_sweeps;
   const int *const *jgs_blocks; /* [L] block boundaries (nblocks+1 ints) for hybrid JGS, or NULL */
   const int *jgs_nblocks;       /* [L] */
   int jgs_parfor_scale;         /* 1: divide by a_ii/w (Parfor variant, SMEM_Smooth.cpp:253-263) */
   int coarse_solve;             /* 1: DMEM convention -- direct solve on the coarsest level (hypre_GaussElimSolve,
                                    src/DMEM_Add.cpp:262-264, src/DMEM_Mult.cpp:393); 0: SMEM (contributes nothing) */
} orc_problem;

/* ---- SpMV family ------------------------------------------------------------------------ */
/* src/SMEM_MatVec.cpp:302-323 (SMEM_MatVec), src/SEQ_MatVec.cpp:3-24 */
void orc_matvec(const orc_csr *A, const double *x, double *y, int ns, int ne)
{
#pragma omp parallel for schedule(static)
   for (int i = ns; i < ne; i++) {
      double Axi = 0.0;
      for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) Axi += A->data[jj] * x[A->j[jj]];
      y[i] = Axi;
   }
}

/* y = alpha*A*x + beta*b   (src/SMEM_MatVec.cpp:123-259; the 12 alpha/beta special cases
 * there are algebraically this expression; residual is alpha=-1, beta=1 as at :95-103) */
void orc_spgemv(const orc_csr *A, const double *x, const double *b, double alpha, double beta, double *y)
{
#pragma omp parallel for schedule(static)
   for (int i = 0; i < A->nrows; i++) {
      double Axi = 0.0;
      for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) Axi += A->data[jj] * x[A->j[jj]];
      double v;
      if (alpha == -1.0 && beta == 1.0) v = b[i] - Axi;
      else if (alpha == 1.0 && beta == 1.0) v = b[i] + Axi;
      else if (beta == 0.0) v = alpha * Axi;
      else v = alpha * Axi + beta * b[i];
      y[i] = v;
   }
}

/* src/SMEM_MatVec.cpp:362-378 */
void orc_residual(const orc_csr *A, const double *b, const double *x, double *r)
{
   orc_spgemv(A, x, b, -1.0, 1.0, r);
}

/* y = A^T x   (src/SEQ_MatVec.cpp:26-45) */
void orc_matvecT(const orc_csr *A, const double *x, double *y)
{
   for (int i = 0; i < A->ncols; i++) y[i] = 0;
   for (int i = 0; i < A->nrows; i++)
      for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) y[A->j[jj]] += A->data[jj] * x[i];
}

/* sqrt(sum r_i^2)  (src/SMEM_Solve.cpp:199-203; src/Misc.cpp:296-309) */
double orc_norm2(const double *x, int n)
{
   double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
   for (int i = 0; i < n; i++) s += x[i] * x[i];
   return sqrt(s);
}

/* ---- smoothers ---------------------------------------------------------------------------- */
/* src/SMEM_Smooth.cpp:365-407 (SMEM_Sync_Jacobi) / src/SEQ_Smooth.cpp:4-44 */
void orc_jacobi(const orc_csr *A, const double *f, double *u, double *u_prev, double w, int sweeps, int zero_flag)
{
   const int n = A->nrows;
   for (int k = 0; k < sweeps; k++) {
      if (k == 0 && zero_flag) {
#pragma omp parallel for schedule(static)
         for (int i = 0; i < n; i++)
            if (A->data[A->i[i]] != 0.0) u[i] = w * f[i] / A->data[A->i[i]];
      } else {
         memcpy(u_prev, u, sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(static)
         for (int i = 0; i < n; i++)
            if (A->data[A->i[i]] != 0.0) {
               double res = f[i];
               for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) res -= A->data[jj] * u_prev[A->j[jj]];
               u[i] += w * res / A->data[A->i[i]];
            }
      }
   }
}

/* src/SMEM_Smooth.cpp:409-443 (SMEM_Sync_L1Jacobi) */
void orc_l1_jacobi(const orc_csr *A, const double *f, double *u, double *u_prev, const double *l1, int sweeps, int zero_flag)
{
   const int n = A->nrows;
   for (int k = 0; k < sweeps; k++) {
      if (k == 0 && zero_flag) {
#pragma omp parallel for schedule(static)
         for (int i = 0; i < n; i++) u[i] = f[i] / l1[i];
      } else {
         memcpy(u_prev, u, sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(static)
         for (int i = 0; i < n; i++) {
            double res = f[i];
            for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) res -= A->data[jj] * u_prev[A->j[jj]];
            u[i] += res / l1[i];
         }
      }
   }
}

/* src/SMEM_Smooth.cpp:533-586 (SMEM_Sync_HybridJacobiGaussSeidel): Gauss-Seidel inside each
 * block [ns,ne), Jacobi across blocks; weight forced to 1.  `scale`, when non-NULL, replaces
 * a_ii as the divisor (Parfor variant, src/SMEM_Smooth.cpp:253-263,276,297). */
void orc_hybrid_jgs(const orc_csr *A, const double *f, double *u, double *u_prev,
                    const int *blocks, int nblocks, const double *scale, int sweeps, int zero_flag)
{
   const int n = A->nrows;
   for (int k = 0; k < sweeps; k++) {
      if (k == 0 && zero_flag) {
#pragma omp parallel for schedule(dynamic, 16)
         for (int b = 0; b < nblocks; b++) {
            int ns = blocks[b], ne = blocks[b + 1];
            for (int i = ns; i < ne; i++) u[i] = 0.0;
            for (int i = ns; i < ne; i++)
               if (A->data[A->i[i]] != 0.0) {
                  double res = f[i];
                  for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) {
                     int ii = A->j[jj];
                     if (ii >= ns && ii < ne) res -= A->data[jj] * u[ii];
                  }
                  u[i] = res / (scale ? scale[i] : A->data[A->i[i]]);
               }
         }
      } else {
         memcpy(u_prev, u, sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(dynamic, 16)
         for (int b = 0; b < nblocks; b++) {
            int ns = blocks[b], ne = blocks[b + 1];
            for (int i = ns; i < ne; i++)
               if (A->data[A->i[i]] != 0.0) {
                  double res = f[i];
                  for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) {
                     int ii = A->j[jj];
                     if (ii >= ns && ii < ne) res -= A->data[jj] * u[ii];
                     else res -= A->data[jj] * u_prev[ii];
                  }
                  u[i] += res / (scale ? scale[i] : A->data[A->i[i]]);
               }
         }
      }
   }
}

/* src/SMEM_Smooth.cpp:643-702 (SMEM_Sync_SymmetricJacobi) / src/SEQ_Smooth.cpp:119-155:
 * u (+)= (2 M^-1 - M^-1 A M^-1) r,  M = D/w.  The arithmetic order of the reference lines
 * :663,:682-683 is kept. */
void orc_symmetric_jacobi(const orc_csr *A, const double *f, double *u, double *y, double *r,
                          double w, int sweeps, int zero_flag)
{
   const int n = A->nrows;
   int k = 0;
   if (zero_flag) memcpy(r, f, sizeof(double) * (size_t)n);
   else orc_residual(A, f, u, r);
   while (1) {
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) r[i] *= w / A->data[A->i[i]];
      orc_matvec(A, r, y, 0, n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) {
         double d = A->data[A->i[i]];
         r[i] = (2.0 * d * r[i] / w) - y[i];
         r[i] *= w / d;
         if (zero_flag) u[i] = r[i]; else u[i] += r[i];
      }
      k++;
      if (k == sweeps) break;
      orc_residual(A, f, u, r);
      /* the reference keeps zero_flags[level]==1 for later sweeps as well (:685-694), so a
       * second sweep overwrites u; mirrored. */
   }
}

/* src/SMEM_Smooth.cpp:704-762 (SMEM_Sync_SymmetricL1Jacobi) */
void orc_symmetric_l1_jacobi(const orc_csr *A, const double *f, double *u, double *y, double *r,
                             const double *l1, int sweeps, int zero_flag)
{
   const int n = A->nrows;
   int k = 0;
   if (zero_flag) memcpy(r, f, sizeof(double) * (size_t)n);
   else orc_residual(A, f, u, r);
   while (1) {
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) r[i] /= l1[i];
      orc_matvec(A, r, y, 0, n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; i++) {
         r[i] = (2.0 * l1[i] * r[i]) - y[i];
         r[i] /= l1[i];
         if (zero_flag) u[i] = r[i]; else u[i] += r[i];
      }
      k++;
      if (k == sweeps) break;
      orc_residual(A, f, u, r);
   }
}

/* smoother dispatch: src/SMEM_Solve.cpp:264-377, thread_part_type == ALL_LEVELS branch
 * (:277-323) for MULTADD/AFACX, Parfor branch (:324-345) for BPX.  zero_flags[level]==1
 * everywhere in additive cycles (src/SEQ_AMG.cpp:121, src/SMEM_Sync_AMG.cpp:217,457). */
static void orc_smooth(const orc_problem *pb, int level, const double *f, double *u,
                       double *scratch_y, double *scratch_r, int sweeps)
{
   const orc_csr *A = &pb->A[level];
   const int symm = (pb->solver == ORC_MULTADD) && pb->num_pre > 0 && pb->num_post > 0;
   if (pb->smoother == ORC_HYBRID_JGS || pb->smoother == ORC_L1_HYBRID_JGS) {
      double *scale = NULL;
      if (pb->jgs_parfor_scale || pb->solver == ORC_BPX || pb->smoother == ORC_L1_HYBRID_JGS) {
         /* SMEM_Sync_Parfor_HybridJacobiGaussSeidel, src/SMEM_Smooth.cpp:253-263: diag_scale = hypre's l1 norms for the L1
          * variant, A_diag = a_ii / w otherwise (the ALL_LEVELS dispatcher never reaches the L1 variant, src/SMEM_Solve.cpp:277-323) */
         scale = (double *)malloc(sizeof(double) * (size_t)A->nrows);
         for (int i = 0; i < A->nrows; i++)
            scale[i] = pb->smoother == ORC_L1_HYBRID_JGS ? pb->l1[level][i] : A->data[A->i[i]] / pb->smooth_weight;
      }
      orc_hybrid_jgs(A, f, u, scratch_y, pb->jgs_blocks[level], pb->jgs_nblocks[level], scale, sweeps, 1);
      free(scale);
   } else if (pb->smoother == ORC_L1_JACOBI) {
      if (symm) orc_symmetric_l1_jacobi(A, f, u, scratch_y, scratch_r, pb->l1[level], sweeps, 1);
      else orc_l1_jacobi(A, f, u, scratch_y, pb->l1[level], sweeps, 1);
   } else {
      if (symm) orc_symmetric_jacobi(A, f, u, scratch_y, scratch_r, pb->smooth_weight, sweeps, 1);
      else orc_jacobi(A, f, u, scratch_y, pb->smooth_weight, sweeps, 1);
   }
}

/* dense Gaussian elimination with partial pivoting on the coarsest operator: x = A^{-1} b (what
 * hypre_GaussElimSetup / hypre_GaussElimSolve, relax type 9, do -- third-party, restated) */
static void orc_dense_solve(const orc_csr *A, const double *b, double *x)
{
   const int n = A->nrows;
   double *M = (double *)calloc((size_t)n * (n + 1), sizeof(double));
   for (int i = 0; i < n; i++) {
      for (int p = A->i[i]; p < A->i[i + 1]; p++) M[(size_t)i * (n + 1) + A->j[p]] += A->data[p];
      M[(size_t)i * (n + 1) + n] = b[i];
   }
   for (int k = 0; k < n; k++) {
      int piv = k;
      for (int i = k + 1; i < n; i++)
         if (fabs(M[(size_t)i * (n + 1) + k]) > fabs(M[(size_t)piv * (n + 1) + k])) piv = i;
      if (piv != k)
         for (int j = 0; j <= n; j++) { double t = M[(size_t)k * (n + 1) + j]; M[(size_t)k * (n + 1) + j] = M[(size_t)piv * (n + 1) + j]; M[(size_t)piv * (n + 1) + j] = t; }
      const double d = M[(size_t)k * (n + 1) + k];
      for (int i = k + 1; i < n; i++) {
         const double fct = M[(size_t)i * (n + 1) + k] / d;
         if (fct != 0.0)
            for (int j = k; j <= n; j++) M[(size_t)i * (n + 1) + j] -= fct * M[(size_t)k * (n + 1) + j];
      }
   }
   for (int i = n - 1; i >= 0; i--) {
      double sum = M[(size_t)i * (n + 1) + n];
      for (int j = i + 1; j < n; j++) sum -= M[(size_t)i * (n + 1) + j] * x[j];
      x[i] = sum / M[(size_t)i * (n + 1) + i];
   }
   free(M);
}

/* ---- cycles --------------------------------------------------------------------------------- */
typedef struct {
   double **r, **e, **y, **s, **uc, **rf;
   int L;
} orc_work;

static orc_work *work_alloc(const orc_problem *pb)
{
   orc_work *w = (orc_work *)calloc(1, sizeof(orc_work));
   int L = pb->num_levels;
   w->L = L;
   double ***arr[6] = {&w->r, &w->e, &w->y, &w->s, &w->uc, &w->rf};
   for (int a = 0; a < 6; a++) {
      *arr[a] = (double **)calloc((size_t)L, sizeof(double *));
      for (int l = 0; l < L; l++) (*arr[a])[l] = (double *)calloc((size_t)pb->A[l].nrows, sizeof(double));
   }
   return w;
}

static void work_free(orc_work *w)
{
   double **arr[6] = {w->r, w->e, w->y, w->s, w->uc, w->rf};
   for (int a = 0; a < 6; a++) {
      for (int l = 0; l < w->L; l++) free(arr[a][l]);
      free(arr[a]);
   }
   free(w);
}

/* One additive cycle applied to residual w->r[0]; adds every level's correction into u.
 * Multadd / AFACx: src/SEQ_AMG.cpp:110-235 (sequential specification of
 * src/SMEM_Sync_AMG.cpp:408-621).  levels_done[l]++ mirrors local_num_correct. */
/* The correction of ONE level from the restricted residuals w->r[level] (and w->r[level + 1] for AFACx) into w->e[level]:
 * shared by the synchronous cycle (src/SEQ_AMG.cpp:110-235) and the asynchronous chain (src/SMEM_Async_AMG.cpp:109-207). */
static void orc_level_correction(const orc_problem *pb, orc_work *w, int level)
{
   const int L = pb->num_levels;
   double *uf = w->uc[level]; /* u_fine of this level */
   if (level == L - 1 && pb->coarse_solve && L > 1) {
      orc_dense_solve(&pb->A[level], w->r[level], w->e[level]);
   } else if (level == L - 1) {
      /* coarsest: solve commented out / result unused -> contributes 0 */
      memset(w->e[level], 0, sizeof(double) * (size_t)pb->A[level].nrows);
   } else if (pb->solver == ORC_MULTADD) {
      memset(w->e[level], 0, sizeof(double) * (size_t)pb->A[level].nrows);
      orc_smooth(pb, level, w->r[level], w->e[level], w->y[level], w->s[level], pb->fine_sweeps);
   } else { /* AFACx: src/SEQ_AMG.cpp:172-208, src/SMEM_Async_AMG.cpp:153-206 */
      int c = level + 1;
      double *ucoarse = w->rf[c];
      memset(ucoarse, 0, sizeof(double) * (size_t)pb->A[c].nrows);
      orc_smooth(pb, c, w->r[c], ucoarse, w->y[c], w->s[c], pb->coarse_sweeps);
      orc_matvec(&pb->P[level], ucoarse, w->e[level], 0, pb->P[level].nrows);
      /* r_fine = r - A e */
      orc_spgemv(&pb->A[level], w->e[level], w->r[level], -1.0, 1.0, w->y[level]);
      double *rfine = (double *)malloc(sizeof(double) * (size_t)pb->A[level].nrows);
      memcpy(rfine, w->y[level], sizeof(double) * (size_t)pb->A[level].nrows);
      memset(uf, 0, sizeof(double) * (size_t)pb->A[level].nrows);
      orc_smooth(pb, level, rfine, uf, w->y[level], w->s[level], pb->fine_sweeps);
      memcpy(w->e[level], uf, sizeof(double) * (size_t)pb->A[level].nrows);
      free(rfine);
   }
}

static void orc_add_vcycle(const orc_problem *pb, orc_work *w, double *u, int *levels_done)
{
   const int L = pb->num_levels;
   for (int l = 0; l < L - 1; l++) orc_matvec(&pb->R[l], w->r[l], w->r[l + 1], 0, pb->R[l].nrows);
   for (int level = 0; level < L; level++) {
      orc_level_correction(pb, w, level);
      if (levels_done) levels_done[level]++;
      /* prolong this level's correction to level 0 (src/SEQ_AMG.cpp:213-228) */
      for (int inner = level; inner > 0; inner--)
         orc_matvec(&pb->P[inner - 1], w->e[inner], w->e[inner - 1], 0, pb->P[inner - 1].nrows);
      const int n0 = pb->A[0].nrows;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n0; i++) u[i] += w->e[0][i];
   }
}

/* BPX: src/SMEM_Sync_AMG.cpp:147-294 (non-PAR_BPX branch): restrict with R = P^T, e_l = S_l r_l on
 * EVERY level incl. the coarsest (zero guess, num_pre sweeps, Parfor smoothers), e_l += P_l e_{l+1}
 * upward, u += e_0. */
static void orc_bpx_cycle(const orc_problem *pb, orc_work *w, double *u)
{
   const int L = pb->num_levels;
   for (int l = 0; l < L - 1; l++) orc_matvec(&pb->R[l], w->r[l], w->r[l + 1], 0, pb->R[l].nrows);
   for (int l = 0; l < L; l++) {
      memset(w->e[l], 0, sizeof(double) * (size_t)pb->A[l].nrows);
      orc_smooth(pb, l, w->r[l], w->e[l], w->y[l], w->s[l], pb->num_pre);
   }
   for (int l = L - 2; l >= 0; l--) orc_spgemv(&pb->P[l], w->e[l + 1], w->e[l], 1.0, 1.0, w->e[l]);
   const int n0 = pb->A[0].nrows;
#pragma omp parallel for schedule(static)
   for (int i = 0; i < n0; i++) u[i] += w->e[0][i];
}

/* One application of the selected cycle to residual r (length n0): u += B r.  Exposed for
 * per-cycle parity tests. */
/* Multiplicative V-cycle: SMEM_Sync_Parfor_Vcycle, src/SMEM_Sync_AMG.cpp:8-145 (non-preconditioner form).
 * u is level 0's solution (in/out), f its right-hand side.  w->e[l] (l >= 1) are the level solutions, w->r[l]
 * (l >= 1) the level right-hand sides.  Quirks mirrored: zero_flags is 1 on levels 1..L-2 on the way down (the
 * first sweep overwrites u_l), 0 on level 0 and on the way up; the coarsest level's zero flag is never raised, so its
 * num_pre + num_post sweeps start from whatever the previous cycle left in u_{L-1} (0 on the first cycle).
 * coarse_solve = 1: DMEM's comparator (DMEM_Mult / DMEM_MultCycle, src/DMEM_Mult.cpp:13-261, Jacobi branch `smoother == 0`): the same
 * V(1,1) with a direct solve on the coarsest level; the reference applies it to the residual from a zero guess and adds the
 * result to x, which is the same affine map. */
static void orc_mult_vcycle(const orc_problem *pb, orc_work *w, const double *f, double *u)
{
   const int L = pb->num_levels;
   const double om = pb->smooth_weight;
   for (int l = 0; l < L - 1; l++) {
      const double *fl = l == 0 ? f : w->r[l];
      double *ul = l == 0 ? u : w->e[l];
      if (pb->smoother == ORC_L1_JACOBI) orc_l1_jacobi(&pb->A[l], fl, ul, w->y[l], pb->l1[l], pb->num_pre, l > 0);
      else orc_jacobi(&pb->A[l], fl, ul, w->y[l], om, pb->num_pre, l > 0);
      orc_residual(&pb->A[l], fl, ul, w->rf[l]);
      orc_matvec(&pb->R[l], w->rf[l], w->r[l + 1], 0, pb->R[l].nrows);
   }
   {
      const int c = L - 1;
      const double *fc = c == 0 ? f : w->r[c];
      double *ucs = c == 0 ? u : w->e[c];
      /* DMEM convention (DMEM_MultCycle, src/DMEM_Mult.cpp:207: hypre_GaussElimSolve on the coarsest level) */
      if (pb->coarse_solve && L > 1) orc_dense_solve(&pb->A[c], fc, ucs);
      else if (pb->smoother == ORC_L1_JACOBI) orc_l1_jacobi(&pb->A[c], fc, ucs, w->y[c], pb->l1[c], pb->num_pre + pb->num_post, 0);
      else orc_jacobi(&pb->A[c], fc, ucs, w->y[c], om, pb->num_pre + pb->num_post, 0);
   }
   for (int l = L - 2; l >= 0; l--) {
      const double *fl = l == 0 ? f : w->r[l];
      double *ul = l == 0 ? u : w->e[l];
      orc_spgemv(&pb->P[l], w->e[l + 1], ul, 1.0, 1.0, w->s[l]);
      memcpy(ul, w->s[l], sizeof(double) * (size_t)pb->A[l].nrows);
      if (pb->smoother == ORC_L1_JACOBI) orc_l1_jacobi(&pb->A[l], fl, ul, w->y[l], pb->l1[l], pb->num_post, 0);
      else orc_jacobi(&pb->A[l], fl, ul, w->y[l], om, pb->num_post, 0);
   }
}

void orc_cycle(const orc_problem *pb, const double *r, double *u)
{
   orc_work *w = work_alloc(pb);
   memcpy(w->r[0], r, sizeof(double) * (size_t)pb->A[0].nrows);
   if (pb->solver == ORC_BPX) orc_bpx_cycle(pb, w, u);
   else orc_add_vcycle(pb, w, u, NULL);
   work_free(w);
}

/* Outer loop: src/SMEM_Solve.cpp:60-70 (r0), :128-240 (cycle, Chebyshev :169-188, residual
 * :192-197, norm :199-203, stop :222).  relres[0]=1, relres[k] after cycle k.  Returns the
 * number of cycles done.  cheby: 0 off; else mu, delta as from ChebySetup
 * (src/SMEM_Cheby.cpp:48-49). */
int orc_solve_sync(const orc_problem *pb, const double *f, double *u, double tol, int num_cycles,
                   int cheby_flag, double mu, double delta, double *relres, double *seconds)
{
   const int n0 = pb->A[0].nrows;
   orc_work *w = work_alloc(pb);
   double *r = w->r[0];
   double *u_outer = NULL, *y_outer = NULL, *c = NULL;
   double omega = 2.0, mu24 = 4.0 * mu * mu;
   orc_residual(&pb->A[0], f, u, r);
   const double r0 = orc_norm2(r, n0);
   relres[0] = 1.0;
   if (cheby_flag) {
      u_outer = (double *)calloc((size_t)n0, sizeof(double));
      y_outer = (double *)calloc((size_t)n0, sizeof(double));
      c = (double *)calloc((size_t)n0, sizeof(double));
   }
   int k, done = 0;
   double t0 = omp_get_wtime();
   for (k = 1; k <= num_cycles; k++) {
      if (cheby_flag) {
         /* precond form: cycle from zero guess on the current residual (src/SMEM_Sync_AMG.cpp:281-286
          * precond_flag branch), then the three-term recurrence src/SMEM_Solve.cpp:179-187 */
         memset(c, 0, sizeof(double) * (size_t)n0);
         if (pb->solver == ORC_BPX) orc_bpx_cycle(pb, w, c);
         /* MULT as a preconditioner (src/SMEM_Sync_AMG.cpp:29-36): the V-cycle on the right-hand side r from a zero guess */
         else if (pb->solver == ORC_MULT) orc_mult_vcycle(pb, w, r, c);
         else orc_add_vcycle(pb, w, c, NULL);
#pragma omp parallel for schedule(static)
         for (int i = 0; i < n0; i++) {
            double u_outer_prev = u_outer[i];
            u_outer[i] = y_outer[i] + omega * (delta * c[i] + u_outer[i] - y_outer[i]);
            y_outer[i] = u_outer_prev;
            u[i] = u_outer[i];
         }
         omega = 1.0 / (1.0 - omega / mu24);
      } else {
         if (pb->solver == ORC_MULT) orc_mult_vcycle(pb, w, f, u);
         else if (pb->solver == ORC_BPX) orc_bpx_cycle(pb, w, u);
         else orc_add_vcycle(pb, w, u, NULL);
      }
      orc_residual(&pb->A[0], f, u, r);
      double rn = orc_norm2(r, n0);
      relres[k] = rn / r0;
      done = k;
      if (rn / r0 < tol) break;
   }
   if (seconds) *seconds = omp_get_wtime() - t0;
   free(u_outer); free(y_outer); free(c);
   work_free(w);
   return done;
}

/* Sequential model of the asynchronous additive solve with no staleness: every level applies
 * its chain to the residual of the CURRENT shared u, one level after the other (the
 * num_threads-independent limit of src/SMEM_Async_AMG.cpp:79-352 when groups never overlap; with coarse_solve it is
 * DMEM_Add's asynchronous loop, src/DMEM_Add.cpp:101-130, grid after grid -- AddCycle :180-329 + DMEM_AddSmooth
 * src/DMEM_Smooth.cpp:574-638 -- which the reference's object code pins, tests/test_oracle_golden.py).
 * counts[l] = corrections. */
static int async_sequential(const orc_problem *pb, const double *f, double *u, int num_cycles, int *counts, double *final_relres,
                            int read_res)
{
   const int L = pb->num_levels, n0 = pb->A[0].nrows;
   orc_work *w = work_alloc(pb);
   orc_residual(&pb->A[0], f, u, w->r[0]);
   const double r0 = orc_norm2(w->r[0], n0);
   /* -read_type res: per-level accumulators of the corrections (level_vector[q].f[0]) */
   double **facc = NULL, *ye = NULL;
   if (read_res) {
      facc = (double **)malloc(sizeof(double *) * L);
      for (int q = 0; q < L; q++) facc[q] = (double *)calloc((size_t)n0, sizeof(double));
      ye = (double *)malloc(sizeof(double) * (size_t)n0);
   }
   for (int k = 0; k < num_cycles; k++)
      for (int q = 0; q < L; q++) {
         /* restriction chain: down to the group's level (Multadd) or one level further (AFACx), :84-108 */
         const int coarsest = pb->solver == ORC_MULTADD ? q : q + 1;
         for (int l = 0; l < coarsest && l < L - 1; l++) orc_matvec(&pb->R[l], w->r[l], w->r[l + 1], 0, pb->R[l].nrows);
         /* the level's correction: Multadd smoother / AFACx two-stage correction; the last level contributes 0 in SMEM and
          * solves directly in DMEM (AddCycle, src/DMEM_Add.cpp:262-264) */
         orc_level_correction(pb, w, q);
         for (int inner = q; inner > 0; inner--)
            orc_matvec(&pb->P[inner - 1], w->e[inner], w->e[inner - 1], 0, pb->P[inner - 1].nrows);
         counts[q]++;
         if (read_res) {
            /* :227-236, :288-293: y = A_0 e; f_q += e; r -= y (the shared residual is maintained incrementally and u is
             * assembled only at the end) */
            orc_matvec(&pb->A[0], w->e[0], ye, 0, n0);
            for (int i = 0; i < n0; i++) { facc[q][i] += w->e[0][i]; w->r[0][i] -= ye[i]; }
         } else {
            for (int i = 0; i < n0; i++) u[i] += w->e[0][i];
            orc_residual(&pb->A[0], f, u, w->r[0]);
         }
      }
   if (read_res) {
      /* :416-426: u += sum over the levels of their accumulated corrections, level after level */
      for (int q = 0; q < L; q++) {
         for (int i = 0; i < n0; i++) u[i] += facc[q][i];
         free(facc[q]);
      }
      free(facc); free(ye);
      orc_residual(&pb->A[0], f, u, w->r[0]);      /* SMEM_Solve's final residual, src/SMEM_Solve.cpp:82-91 */
   }
   *final_relres = orc_norm2(w->r[0], n0) / r0;
   work_free(w);
   return 0;
}

int orc_solve_async_sequential(const orc_problem *pb, const double *f, double *u, int num_cycles,
                               int *counts, double *final_relres)
{
   return async_sequential(pb, f, u, num_cycles, counts, final_relres, 0);
}

/* the same with `-read_type res` (READ_RES, FULL_ASYNC, local residual computation): src/SMEM_Async_AMG.cpp:227-236,285-296,416-426 */
int orc_solve_async_sequential_res(const orc_problem *pb, const double *f, double *u, int num_cycles,
                                   int *counts, double *final_relres)
{
   return async_sequential(pb, f, u, num_cycles, counts, final_relres, 1);
}

/* DMEM's accelerated synchronous additive solve: DMEM_SyncAddCorrect + DMEM_ChebyUpdate
 * (src/DMEM_Add.cpp:647-716, src/DMEM_Misc.cpp:612-666, ChebySetup src/DMEM_Setup.cpp:1901-1914).  Per cycle e = B r
 * (all grids' corrections accumulated), then cycle 0: d = e; later cycles: d = (omega-1) d + omega*delta*e with
 * omega = 2/(1+sqrt(1-mu^-2)) (accel 2, Richardson) or omega = 2 mu c_k / c_{k+1}, c_{k+1} = 2 mu c_k - c_{k-1},
 * c_0 = 1, c_1 = mu (accel 1, Chebyshev); x += d.  accel 0: x += e. */
int orc_solve_sync_dmem(const orc_problem *pb, const double *f, double *u, double tol, int num_cycles,
                        int accel, double mu, double delta, double *relres)
{
   const int n0 = pb->A[0].nrows;
   orc_work *w = work_alloc(pb);
   double *r = w->r[0];
   double *e = (double *)calloc((size_t)n0, sizeof(double)), *d = (double *)calloc((size_t)n0, sizeof(double));
   double c_prev = 1.0, c_cur = mu;
   orc_residual(&pb->A[0], f, u, r);
   const double r0 = orc_norm2(r, n0);
   relres[0] = 1.0;
   int done = 0;
   for (int k = 0; k < num_cycles; k++) {
      memset(e, 0, sizeof(double) * (size_t)n0);
      if (pb->solver == ORC_BPX) orc_bpx_cycle(pb, w, e);
      else orc_add_vcycle(pb, w, e, NULL);
      if (accel == 0) {
         for (int i = 0; i < n0; i++) u[i] += e[i];
      } else {
         if (k == 0) memcpy(d, e, sizeof(double) * (size_t)n0);
         else {
            double omega;
            if (accel == 2) omega = 2.0 / (1.0 + sqrt(1.0 - pow(mu, -2.0)));
            else {
               const double c_temp = c_cur;
               c_cur = 2.0 * mu * c_cur - c_prev;
               c_prev = c_temp;
               omega = 2.0 * mu * c_prev / c_cur;
            }
            for (int i = 0; i < n0; i++) d[i] = (omega - 1.0) * d[i] + omega * delta * e[i];
         }
         for (int i = 0; i < n0; i++) u[i] += d[i];
      }
      orc_residual(&pb->A[0], f, u, r);
      const double rn = orc_norm2(r, n0);
      relres[k + 1] = rn / r0;
      done = k + 1;
      if (rn / r0 < tol) break;
   }
   free(e); free(d);
   work_free(w);
   return done;
}

/* EigsPower (src/SMEM_Cheby.cpp:410-518): extreme eigenvalues of B*A by power iteration, B = one application of
 * the selected cycle from a zero guess (the reference hard-wires its hypre-vector BPXCycle for every solver other
 * than MULT; here B is the cycle `pb` names, identical for BPX).  Start vector all ones; `iters` normalise /
 * apply steps; eig_max = <v, BAv>; second pass deflated with u <- BAv - eig_max v; eig_min = <v, BAv>. */
void orc_eigs_power(const orc_problem *pb, int iters, double *eig_min, double *eig_max)
{
   const int n = pb->A[0].nrows;
   double *u = (double *)malloc(sizeof(double) * n), *e = (double *)malloc(sizeof(double) * n);
   double *f = (double *)malloc(sizeof(double) * n);
   double lam[2] = {0.0, 0.0};
   for (int pass = 0; pass < 2; pass++) {
      for (int i = 0; i < n; i++) u[i] = 1.0;
      for (int it = 1;; it++) {
         const double nu = orc_norm2(u, n);
         for (int i = 0; i < n; i++) { u[i] /= nu; e[i] = u[i]; }
         orc_matvec(&pb->A[0], u, f, 0, n);
         memset(u, 0, sizeof(double) * (size_t)n);      /* hypre_ParVectorSetConstantValues(u, 0.0), :452 */
         orc_cycle(pb, f, u);
         if (it == iters) break;
         if (pass == 1)
            for (int i = 0; i < n; i++) u[i] -= lam[0] * e[i];
      }
      double d = 0.0;
      for (int i = 0; i < n; i++) d += e[i] * u[i];
      lam[pass] = d;
   }
   *eig_max = lam[0];
   *eig_min = lam[1];
   free(u); free(e); free(f);
}

/* ---- implicit extended-system BPX solver (`-solver iebpx`), synchronous ----------------------------------------------
 * SMEM_ExtendedSystemSolve with IMPLICIT_EXTENDED_SYSTEM_BPX, src/SMEM_ExtendedSystem.cpp:9-836 (+ ExtendedSystemImplicitMatVec
 * :838-907).  Chebyshev-accelerated Jacobi on Griebel's semi-definite "generating system"  AA xx = bb  whose unknowns are the
 * per-level vectors u_0 .. u_{L-1}; block (k,l) of AA is  A_k P^{k<-l}  for l >= k  and  R^{k<-l} A_l  for l < k  (never
 * formed).  R = P^T, plain P (src/SMEM_Setup.cpp:262-274).  The solution of A x = f is x = sum_l P^{0<-l} u_l.
 *   set-up (:112-136): f_{l+1} = R_l f_l; r0_ext = sqrt(sum_l |f_l|^2); y_l = 0; u_l = delta f_l / a_ii (the raw diagonal);
 *                      e_l = A_l u_l
 *   every iteration (:392-636), omega_0 = 2:
 *     phase 1 (all levels from the same u, e):  z1_{L-1} = u_{L-1}, z1_k = P_k z1_{k+1} + u_k;
 *                                               z2_0 = 0, z2_{k+1} = R_k (z2_k + e_k)
 *     phase 2 (per level k):  z = A_k z1_k + z2_k;  r_k = f_k - z;  s = r_k ./ scale_k  (scale = a_ii/w, or the L1 norms);
 *                             u_k <- y_k + omega (delta s + u_k - y_k),  y_k <- old u_k;  e_k = A_k u_k
 *     omega <- 1 / (1 - omega / (2 mu)^2);  stop when the iteration counter reaches num_cycles, or -- from the second
 *     iteration on -- when sqrt(sum_k |r_k|^2) / r0_ext < tol (:636-658; the counter starts at 1 and the loop is skipped
 *     altogether for num_cycles <= 1, :271).
 *   finish (:777-817): extended residual of the final iterate, x = sum_l P^{0<-l} u_l, r = f - A_0 x.
 * ext_hist[it] (it >= 1) = sqrt(sum_k |r_k|^2)/r0_ext as measured in iteration it (the residual of the iterate that ENTERS
 * the iteration); returns the final value of the reference's loc_iters (= grid.local_num_correct). */
int orc_solve_iebpx(const orc_problem *pb, const double *f, double *x_out, double tol, int num_cycles, double mu, double delta,
                    double *ext_hist, double *ext_relres, double *relres)
{
   const int L = pb->num_levels;
   double **fl = (double **)malloc(sizeof(double *) * L), **u = (double **)malloc(sizeof(double *) * L);
   double **y = (double **)malloc(sizeof(double *) * L), **e = (double **)malloc(sizeof(double *) * L);
   double **z1 = (double **)malloc(sizeof(double *) * L), **z2 = (double **)malloc(sizeof(double *) * L);
   double **z = (double **)malloc(sizeof(double *) * L), **t = (double **)malloc(sizeof(double *) * L);
   for (int l = 0; l < L; l++) {
      const size_t n = (size_t)pb->A[l].nrows;
      fl[l] = (double *)calloc(n, sizeof(double)); u[l] = (double *)calloc(n, sizeof(double));
      y[l] = (double *)calloc(n, sizeof(double)); e[l] = (double *)calloc(n, sizeof(double));
      z1[l] = (double *)calloc(n, sizeof(double)); z2[l] = (double *)calloc(n, sizeof(double));
      z[l] = (double *)calloc(n, sizeof(double)); t[l] = (double *)calloc(n, sizeof(double));
   }
   const int n0 = pb->A[0].nrows;
   memcpy(fl[0], f, sizeof(double) * (size_t)n0);
   double ss = 0.0;
   for (int i = 0; i < n0; i++) ss += f[i] * f[i];
   const double r0 = sqrt(ss);
   for (int l = 0; l < L - 1; l++) {
      orc_matvec(&pb->R[l], fl[l], fl[l + 1], 0, pb->R[l].nrows);
      for (int i = 0; i < pb->A[l + 1].nrows; i++) ss += fl[l + 1][i] * fl[l + 1][i];
   }
   const double r0_ext = sqrt(ss);
   for (int l = 0; l < L; l++) {
      const orc_csr *A = &pb->A[l];
      for (int i = 0; i < A->nrows; i++) u[l][i] = delta * fl[l][i] / A->data[A->i[i]];
      orc_matvec(A, u[l], e[l], 0, A->nrows);
   }
   double omega = 2.0;
   const double mu22 = (2.0 * mu) * (2.0 * mu);
   int it = 1;
   /* phases 1 of the loop and of the final ExtendedSystemImplicitMatVec are the same computation */
#define ORC_EXT_PHASE1()                                                                         \
   do {                                                                                          \
      memcpy(z1[L - 1], u[L - 1], sizeof(double) * (size_t)pb->A[L - 1].nrows);                 \
      for (int k = L - 2; k >= 0; k--) {                                                         \
         orc_matvec(&pb->P[k], z1[k + 1], z1[k], 0, pb->P[k].nrows);                            \
         for (int i = 0; i < pb->A[k].nrows; i++) z1[k][i] += u[k][i];                          \
      }                                                                                          \
      memset(z2[0], 0, sizeof(double) * (size_t)n0);                                            \
      for (int k = 0; k < L - 1; k++) {                                                          \
         for (int i = 0; i < pb->A[k].nrows; i++) t[k][i] = z2[k][i] + e[k][i];                 \
         orc_matvec(&pb->R[k], t[k], z2[k + 1], 0, pb->R[k].nrows);                             \
      }                                                                                          \
   } while (0)
   if (num_cycles > 1)
      for (;;) {
         ORC_EXT_PHASE1();
         double rs = 0.0;
         for (int k = 0; k < L; k++) {
            const orc_csr *A = &pb->A[k];
            orc_matvec(A, z1[k], z[k], 0, A->nrows);
            for (int i = 0; i < A->nrows; i++) {
               z[k][i] += z2[k][i];
               const double r = fl[k][i] - z[k][i];
               const double scale = (pb->smoother == ORC_L1_JACOBI) ? pb->l1[k][i] : A->data[A->i[i]] / pb->smooth_weight;
               const double us = r / scale;
               const double up = u[k][i];
               u[k][i] = y[k][i] + omega * (delta * us + u[k][i] - y[k][i]);
               y[k][i] = up;
               rs += r * r;
            }
            orc_matvec(A, u[k], e[k], 0, A->nrows);
         }
         if (ext_hist) ext_hist[it] = sqrt(rs) / r0_ext;
         const int measured = it > 1;                /* check_resnorm_flag && loc_iters > 1 (:618) */
         omega = 1.0 / (1.0 - omega / mu22);
         it++;
         if (it == num_cycles) break;
         if (measured && sqrt(rs) / r0_ext < tol) break;
      }
   ORC_EXT_PHASE1();
   ss = 0.0;
   for (int k = 0; k < L; k++) {
      const orc_csr *A = &pb->A[k];
      orc_matvec(A, z1[k], z[k], 0, A->nrows);
      for (int i = 0; i < A->nrows; i++) { const double r = fl[k][i] - (z[k][i] + z2[k][i]); ss += r * r; }
   }
   if (ext_relres) *ext_relres = sqrt(ss) / r0_ext;
   for (int k = L - 2; k >= 0; k--) {
      orc_matvec(&pb->P[k], u[k + 1], e[k], 0, pb->P[k].nrows);
      for (int i = 0; i < pb->A[k].nrows; i++) u[k][i] += e[k][i];
   }
   memcpy(x_out, u[0], sizeof(double) * (size_t)n0);
   orc_residual(&pb->A[0], f, u[0], z[0]);
   if (relres) *relres = orc_norm2(z[0], n0) / r0;
   for (int l = 0; l < L; l++) { free(fl[l]); free(u[l]); free(y[l]); free(e[l]); free(z1[l]); free(z2[l]); free(z[l]); free(t[l]); }
   free(fl); free(u); free(y); free(e); free(z1); free(z2); free(z); free(t);
   return it;
}
#undef ORC_EXT_PHASE1

int orc_max_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int t) { omp_set_num_threads(t); }
