// async_fact0.cu -- EXPERIMENTAL copy of the persistent asynchronous kernel (k_async_amg, async.cu) for
// amgb_options.factor_level0 with ASYNC_MULTADD.  Compiled, NOT yet run on hardware; kept in its own translation unit so that
// the object code of k_async_amg -- the kernel every measurement of round 1 used -- is not perturbed.
//
// P_0 / R_0 are the PLAIN transfers and the smoothing factors are applied on the fly:
//     Rbar_0 r = R_0 (r - A_0 diag(w/d) r),        Pbar_0 e = v - (w/d) o (A_0 v),  v = P_0 e
// -- two passes over the stencil A_0 (sliced ELL, HBM speed) and the 4-entry-per-row plain transfers replace the 3.5x denser,
// L1-bound products that every level group streams for every correction.  Only Multadd with the
// symmetrised (L1-)Jacobi smoother gets here (async_prepare in async.cu).  Everything else is k_async_amg verbatim.
#include "ctx.h"
#include "kernels.cuh"
#include "async_team.cuh"

namespace {

__global__ void __launch_bounds__(kABlock, 3) k_async_amg_fact0(const AsyncParams *__restrict__ pp)
{
   const AsyncParams &p = *pp;
   const int L = p.num_levels;
   // which level's group does this CTA belong to
   int q = 0;
   while (q + 1 < L && (int)blockIdx.x >= p.cta_begin[q + 1]) q++;
   Team tm;
   tm.cta = blockIdx.x - p.cta_begin[q];
   tm.nctas = p.cta_begin[q + 1] - p.cta_begin[q];
   tm.tid = tm.cta * kABlock + threadIdx.x;
   tm.size = tm.nctas * kABlock;
   tm.count = p.barrier_count + q;
   tm.gen = p.barrier_gen + q;

   const AsyncLevelVecs &v = p.g[q];
   double *t0 = p.t0[q];
   const bool multadd = p.solver == AMGB_SOLVER_ASYNC_MULTADD;
   const int n0 = p.A[0].nrows;
   __shared__ int s_stop;
   extern __shared__ __align__(128) unsigned char dyn_smem[];
   tm.smem = dyn_smem;

   // The coarsest level's correction is identically zero in the reference (direct solve commented
   // out, :112-131): its restrict / prolong / residual work adds exactly 0.0 to u, so this group only
   // keeps the correction count and the stop protocol.
   const bool idle = (q == L - 1);

   while (true) {
      // ---- restriction chain (src/SMEM_Async_AMG.cpp:93-108)
      const int coarsest = idle ? 0 : (multadd ? q : q + 1);
      for (int l = 0; l < coarsest; l++) {
         if (l == 0 && L > 1) {
            spmv_team<false, true>(p.A[0], v.r[0], t0, mk(-1.0, 1.0, v.r[0]), tm.tid, tm.size, false, tm.smem);   // t_0 = r_0 - A_0 diag(w/d) r_0
            group_barrier(tm);
            spmv_team<false, false>(p.R[0], t0, v.r[1], mk(1.0, 0.0, nullptr), tm.tid, tm.size, false, tm.smem);  // r_1 = R_0 t_0
            group_barrier(tm);
         } else if (l < L - 1) {
            spmv_team<false, false>(p.R[l], v.r[l], v.r[l + 1], mk(1.0, 0.0, nullptr), tm.tid, tm.size, false, tm.smem);
            group_barrier(tm);
         }
      }
      // ---- correction on the group's level (:134-207)
      if (q == L - 1) {
         // coarsest grid: the direct solve is commented out in the reference (:112-131); e stays 0
         group_barrier(tm);
      } else if (multadd) {
         team_smooth_zero(p, tm, q, v.r[q], v.e[q], v.t[q], p.fine_sweeps, p.symmetric != 0);
      } else {
         // AFACx (:153-206): u_c = S_{q+1} r_{q+1}; e = P u_c; r_f = r_q - A_q e; u_f = S_q r_f
         const int cl = q + 1;
         team_smooth_zero(p, tm, cl, v.r[cl], v.t[cl], v.w[cl], p.coarse_sweeps, false);
         spmv_team<false, false>(p.P[q], v.t[cl], v.t[q], mk(1.0, 0.0, nullptr), tm.tid, tm.size, false, tm.smem);
         group_barrier(tm);
         spmv_team<false, false>(p.A[q], v.t[q], v.w[q], mk(-1.0, 1.0, v.r[q]), tm.tid, tm.size, false, tm.smem);
         group_barrier(tm);
         team_smooth_zero(p, tm, q, v.w[q], v.e[q], v.t[q], p.fine_sweeps, false);
      }
      // ---- prolongation chain (:211-224)
      for (int l = idle ? -1 : q - 1; l >= 0; l--) {
         if (l == 0) {
            spmv_team<false, false>(p.P[0], v.e[1], t0, mk(1.0, 0.0, nullptr), tm.tid, tm.size, false, tm.smem);     // v = P_0 e_1
            group_barrier(tm);
            SpmvEpilogue fe = mk(-1.0, 0.0, nullptr, 0.0, nullptr, (p.smoother == AMGB_SMOOTH_L1_JACOBI) ? p.inv_l1[0] : p.ws[0]);
            fe.xs = t0; fe.xself = 1.0;
            spmv_team<false, false>(p.A[0], t0, v.e[0], fe, tm.tid, tm.size, false, tm.smem);                         // e_0 = v - (w/d) o (A_0 v)
            group_barrier(tm);
            continue;
         }
         spmv_team<false, false>(p.P[l], v.e[l + 1], v.e[l], mk(1.0, 0.0, nullptr), tm.tid, tm.size, false, tm.smem);
         group_barrier(tm);
      }
      // ---- u += e (atomic), private copy (:285-301)
      if (!idle)
         for (int i = tm.tid; i < n0; i += tm.size) {
            red_add_f64(p.u + i, ld_cg(v.e[0] + i));
            v.u_local[i] = ld_cg(p.u + i);
         }
      // ---- correction count and stop rule (:314-337)
      if (tm.tid == 0) {
         const int cnt = *((volatile int *)(p.num_correct + q)) + 1;
         *((volatile int *)(p.num_correct + q)) = cnt;
         __threadfence();
         if (p.converge_type == AMGB_CONVERGE_GLOBAL && q == 0 && *p.converge_flag == 0) {
            int all = 1;
            for (int l = 0; l < L; l++)
               if (*((volatile int *)(p.num_correct + l)) < p.num_cycles) { all = 0; break; }
            if (all) { *p.converge_flag = 1; __threadfence(); }
         }
      }
      group_barrier(tm);
      if (threadIdx.x == 0) {
         int stop;
         if (p.converge_type == AMGB_CONVERGE_LOCAL) stop = *((volatile int *)(p.num_correct + q)) >= p.num_cycles;
         else stop = *p.converge_flag;
         s_stop = stop;
      }
      // all CTAs of the group must take the same decision: publish the root CTA's view
      if (tm.nctas > 1 && p.converge_type != AMGB_CONVERGE_LOCAL) {
         // GLOBAL: the flag may flip between two CTAs' reads; the group root decides
         __syncthreads();
         if (tm.cta == 0 && threadIdx.x == 0) { *((volatile int *)(p.group_stop + q)) = s_stop; __threadfence(); }
         group_barrier(tm);
         if (threadIdx.x == 0) s_stop = *((volatile int *)(p.group_stop + q));
      }
      __syncthreads();
      const int stop = s_stop;
      // ---- private residual from the private copy (:338-351)
      if (!idle) spmv_team<false, false>(p.A[0], v.u_local, v.r[0], mk(-1.0, 1.0, p.f), tm.tid, tm.size, false, tm.smem);
      group_barrier(tm);
      if (stop) break;
   }
}

}  // namespace

int async_max_grid_fact0(int block)
{
   int dev = 0, sms = 0, per_sm = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   cudaFuncSetAttribute(k_async_amg_fact0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AMGB_TEAM_SMEM);
   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_async_amg_fact0, block, AMGB_TEAM_SMEM);
   return sms * per_sm;
}

int launch_async_fact0(const LaunchCfg &, cudaStream_t st, const AsyncParams *params_dev, int grid, int block,
                       const cudaAccessPolicyWindow *window)
{
   cudaLaunchConfig_t cfg = {};
   cfg.gridDim = dim3(grid);
   cfg.blockDim = dim3(block);
   cfg.dynamicSmemBytes = AMGB_TEAM_SMEM;
   cfg.stream = st;
   cudaLaunchAttribute attrs[2];
   int na = 0;
   attrs[na].id = cudaLaunchAttributeCooperative;
   attrs[na].val.cooperative = 1;
   na++;
   if (window && window->num_bytes > 0) {
      attrs[na].id = cudaLaunchAttributeAccessPolicyWindow;
      attrs[na].val.accessPolicyWindow = *window;
      na++;
   }
   cfg.attrs = attrs;
   cfg.numAttrs = na;
   cudaError_t e = cudaLaunchKernelEx(&cfg, k_async_amg_fact0, params_dev);
   return e == cudaSuccess ? 1 : -(int)e;
}
