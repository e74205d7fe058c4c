// extended.cu -- implicit extended-system BPX solver (`-solver iebpx`) on the device.
//
// Replaces SMEM_ExtendedSystemSolve with IMPLICIT_EXTENDED_SYSTEM_BPX (src/SMEM_ExtendedSystem.cpp:9-836, finish :777-817,
// ExtendedSystemImplicitMatVec :838-907): Chebyshev-accelerated Jacobi on Griebel's semi-definite generating system whose
// unknowns are the per-level vectors u_0 .. u_{L-1}; block (k,l) of the operator is A_k P^{k<-l} for l >= k and
// R^{k<-l} A_l for l < k and is never formed.  Plain P, R = P^T (src/SMEM_Setup.cpp:262-274).
//
//   set-up      f_{l+1} = R_l f_l;  r0_ext = sqrt(sum_l |f_l|^2);  y_l = 0;  u_l = delta f_l ./ diag(A_l);  e_l = A_l u_l
//   iteration   phase 1 (all levels, from the same u, e):   z1_{L-1} = u_{L-1},  z1_k = P_k z1_{k+1} + u_k
//                                                           z2_0 = 0,           z2_{k+1} = R_k (z2_k + e_k)
//               phase 2 (level k):  r_k = f_k - (A_k z1_k + z2_k);  s = r_k ./ scale_k  (a_ii/w or the L1 norm)
//                                   u_k <- y_k + omega (delta s + u_k - y_k),  y_k <- old u_k;  e_k = A_k u_k
//               omega <- 1 / (1 - omega/(2 mu)^2)
//   finish      extended residual of the final iterate;  x = sum_l P^{0<-l} u_l;  r = f - A_0 x
//
// The reference gives every level a thread group that recomputes the z1 / z2 chains it needs (level k: L-1-k prolongations
// and k restrictions, sum over k = L(L-1) SpMVs per iteration); the chains are the same for all groups, so here they are
// computed once per iteration (2(L-1) SpMVs) and shared.  Everything is a composition of the kernels of kernels.cuh: the
// prolongation fuses "+ u_k", the residual fuses "f_k - z2_k - A_k z1_k" and its sum of squares.
//
// The EXPLICIT form (`-solver eebpx`, EXPLICIT_EXTENDED_SYSTEM_BPX, :84-110,295-365) is the same loop on a ONE-level
// context whose A_0 is the assembled extended matrix AA (BuildExtendedMatrix, src/SMEM_Setup.cpp:1426-1521; host side:
// amgh_build_extended_matrix) with smooth_weight = 1 (that branch divides by the raw diagonal): z1 = x, z2 = 0,
// r = bb - AA x, x <- y + omega (delta r ./ diag + x - y).  solver.ExtendedExplicitSolver drives it.
#include "ctx.h"
#include <algorithm>
#include <cmath>
#include <cstring>

struct ExtState {
   std::vector<double *> f, u, y, e, z1, z2, s, r, us;
};

void amgb_ext_teardown(amgb_ctx *c)
{
   delete c->ext;
   c->ext = nullptr;
}

static inline SpmvEpilogue epi(double alpha, double beta, const double *b, double gamma = 0.0, const double *cc = nullptr,
                               const double *rs = nullptr)
{
   SpmvEpilogue e;
   e.alpha = alpha; e.beta = beta; e.gamma = gamma; e.b = b; e.c = cc; e.rs = rs;
   return e;
}

static int ext_alloc(amgb_ctx *c)
{
   if (c->ext) return AMGB_OK;
   ExtState *x = new ExtState();
   const int L = c->L;
   std::vector<double *> *arrs[] = {&x->f, &x->u, &x->y, &x->e, &x->z1, &x->z2, &x->s, &x->r, &x->us};
   for (auto a : arrs) {
      a->assign(L, nullptr);
      for (int l = 0; l < L; l++) {
         int rc = amgb_dev_alloc_bytes(c, (void **)&(*a)[l], sizeof(double) * (size_t)std::max(1, c->A[l].nrows), true);
         if (rc) { delete x; return rc; }
      }
   }
   c->ext = x;
   return AMGB_OK;
}

// z1 / z2 chains from the current u, e (phase 1; also the first half of ExtendedSystemImplicitMatVec)
static void ext_phase1(amgb_ctx *c, ExtState *x)
{
   const int L = c->L;
   cudaMemcpyAsync(x->z1[L - 1], x->u[L - 1], sizeof(double) * (size_t)c->A[L - 1].nrows, cudaMemcpyDeviceToDevice, c->stream);
   for (int k = L - 2; k >= 0; k--) enq_spmv(c, c->P[k], false, x->z1[k + 1], x->z1[k], epi(1.0, 1.0, x->u[k]), false);
   // z2_0 stays zero (allocated zeroed, never written)
   for (int k = 0; k < L - 1; k++) {
      const double *in = x->e[k];
      if (k > 0) {   // s_k = z2_k + e_k: z2_k itself is still needed by phase 2
         cudaMemcpyAsync(x->s[k], x->z2[k], sizeof(double) * (size_t)c->A[k].nrows, cudaMemcpyDeviceToDevice, c->stream);
         c->launches += launch_add(c->cfg, c->stream, c->A[k].nrows, x->e[k], x->s[k]);
         in = x->s[k];
      }
      enq_spmv(c, c->R[k], false, in, x->z2[k + 1], epi(1.0, 0.0, nullptr), false);
   }
}

// r_k = f_k - z2_k - A_k z1_k for every level, d_scalars[1 + k] = |r_k|^2
static void ext_residuals(amgb_ctx *c, ExtState *x)
{
   for (int k = 0; k < c->L; k++) {
      SpmvEpilogue e = epi(-1.0, 1.0, x->f[k]);
      e.b2 = x->z2[k]; e.beta2 = -1.0;
      int grid = 0;
      c->launches += launch_spmv(c->cfg, c->stream, c->A[k], false, x->z1[k], x->r[k], e, c->partials, &grid);
      c->launches += launch_reduce_partials(c->stream, c->partials, grid, c->d_scalars + 1 + k);
   }
}

static int ext_fetch_sum(amgb_ctx *c, double *out)
{
   const int L = c->L;
   CUDA_OK(c, cudaMemcpyAsync(c->h_scalars + 1, c->d_scalars + 1, sizeof(double) * (size_t)L, cudaMemcpyDeviceToHost, c->stream));
   CUDA_OK(c, cudaStreamSynchronize(c->stream));
   double s = 0.0;
   for (int k = 0; k < L; k++) s += c->h_scalars[1 + k];
   *out = s;
   return AMGB_OK;
}

extern "C" int amgb_solve_extended(amgb_ctx *c, double tol, int num_cycles, double mu, double delta, double *ext_hist, int *iters,
                                   double *ext_relres, double *relres, double *solve_seconds)
{
   NEED_READY(c);
   const int L = c->L;
   if (L + 1 > 64) return amgb_fail(c, AMGB_EINVAL, "too many levels");
   if (num_cycles < 0 || mu == 0.0) return amgb_fail(c, AMGB_EINVAL, "bad arguments");
   if (c->dist) return amgb_fail(c, AMGB_EINVAL, "the extended-system solver runs on one GPU");
   const int sm = c->opt.smoother;
   if (sm != AMGB_SMOOTH_JACOBI && sm != AMGB_SMOOTH_L1_JACOBI)
      return amgb_fail(c, AMGB_EINVAL, "the extended-system solver runs with weighted or L1 Jacobi");
   if (c->opt.solver != AMGB_SOLVER_BPX && c->opt.solver != AMGB_SOLVER_IEBPX)
      return amgb_fail(c, AMGB_EINVAL, "the extended-system solver needs the plain transfers of BPX (solver BPX or IEBPX)");
   int rc;
   if ((rc = ext_alloc(c))) return rc;
   ExtState *x = c->ext;
   const int n0 = c->A[0].nrows;
   int grid = 0;
   // ---- set-up (src/SMEM_ExtendedSystem.cpp:112-136) ----
   CUDA_OK(c, cudaMemcpyAsync(x->f[0], c->f, sizeof(double) * (size_t)n0, cudaMemcpyDeviceToDevice, c->stream));
   for (int l = 0; l < L - 1; l++) enq_spmv(c, c->R[l], false, x->f[l], x->f[l + 1], epi(1.0, 0.0, nullptr), false);
   for (int l = 0; l < L; l++) {
      c->launches += launch_sumsq(c->cfg, c->stream, c->A[l].nrows, x->f[l], c->partials, &grid);
      c->launches += launch_reduce_partials(c->stream, c->partials, grid, c->d_scalars + 1 + l);
   }
   double ss;
   if ((rc = ext_fetch_sum(c, &ss))) return rc;
   const double r0 = sqrt(c->h_scalars[1]), r0_ext = sqrt(ss);
   for (int l = 0; l < L; l++) {
      const int n = c->A[l].nrows;
      CUDA_OK(c, cudaMemsetAsync(x->y[l], 0, sizeof(double) * (size_t)n, c->stream));
      // u = delta f ./ a_ii: ws = w/d, so f .* ws scaled by delta/w
      c->launches += launch_scale(c->cfg, c->stream, n, c->ws[l], x->f[l], x->u[l]);
      c->launches += launch_axpby(c->cfg, c->stream, n, 0.0, x->f[l], delta / c->opt.smooth_weight, x->u[l], nullptr);
      if (L > 1) enq_spmv(c, c->A[l], false, x->u[l], x->e[l], epi(1.0, 0.0, nullptr), false);   // e feeds the restriction chain only
   }
   double omega = 2.0;
   const double mu22 = (2.0 * mu) * (2.0 * mu);
   int it = 1;
   CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
   if (num_cycles > 1)
      for (;;) {
         ext_phase1(c, x);
         ext_residuals(c, x);
         for (int k = 0; k < L; k++) {
            const int n = c->A[k].nrows;
            const double *rs = (sm == AMGB_SMOOTH_L1_JACOBI) ? c->inv_l1[k] : c->ws[k];
            c->launches += launch_scale(c->cfg, c->stream, n, rs, x->r[k], x->us[k]);
            // u <- y + omega (delta s + u - y), y <- old u   (k_cheby; its third output goes to scratch)
            c->launches += launch_cheby(c->cfg, c->stream, n, omega, delta, x->us[k], x->u[k], x->y[k], x->s[k]);
            if (L > 1) enq_spmv(c, c->A[k], false, x->u[k], x->e[k], epi(1.0, 0.0, nullptr), false);
         }
         if ((rc = ext_fetch_sum(c, &ss))) return rc;
         const double rel = sqrt(ss) / r0_ext;
         if (ext_hist) ext_hist[it] = rel;
         const bool measured = it > 1;            // check_resnorm_flag && loc_iters > 1 (:618)
         omega = 1.0 / (1.0 - omega / mu22);
         it++;
         if (it == num_cycles) break;
         if (measured && rel < tol) break;
      }
   // ---- finish (:777-817) ----
   ext_phase1(c, x);
   ext_residuals(c, x);
   if ((rc = ext_fetch_sum(c, &ss))) return rc;
   if (ext_relres) *ext_relres = sqrt(ss) / r0_ext;
   for (int k = L - 2; k >= 0; k--) {
      // u_k += P_k u_{k+1}  (out of place through s_k: the kernel must not write the vector other rows' epilogues read)
      enq_spmv(c, c->P[k], false, x->u[k + 1], x->s[k], epi(1.0, 1.0, x->u[k]), false);
      CUDA_OK(c, cudaMemcpyAsync(x->u[k], x->s[k], sizeof(double) * (size_t)c->A[k].nrows, cudaMemcpyDeviceToDevice, c->stream));
   }
   CUDA_OK(c, cudaMemcpyAsync(c->u, x->u[0], sizeof(double) * (size_t)n0, cudaMemcpyDeviceToDevice, c->stream));
   CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
   enq_residual(c);
   if ((rc = amgb_fetch_scalar(c, &ss))) return rc;
   if (relres) *relres = sqrt(ss) / r0;
   float ms = 0;
   CUDA_OK(c, cudaEventSynchronize(c->ev1));
   cudaEventElapsedTime(&ms, c->ev0, c->ev1);
   if (solve_seconds) *solve_seconds = ms * 1e-3;
   if (iters) *iters = it;
   c->r0_norm = r0;
   CUDA_OK(c, cudaGetLastError());
   return AMGB_OK;
}
