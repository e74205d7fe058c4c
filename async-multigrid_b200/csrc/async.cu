// async.cu -- the asynchronous additive AMG solve as ONE persistent cooperative kernel.
//
// Replaces SMEM_Async_Add_AMG (src/SMEM_Async_AMG.cpp:7-437).  The reference runs one long-lived
// OpenMP parallel region in which every AMG level owns a group of threads; a group loops
//   restrict chain -> smooth on its level -> prolong chain -> atomic add into the shared u ->
//   private copy of u -> private residual r = f - A u_private
// with barriers only INSIDE the group (SMEM_LevelBarrier, src/Misc.cpp:485-533) and no barrier
// between groups.  Here a group is a contiguous range of CTAs of a cooperative launch (all CTAs
// co-resident, so the spin barriers below make progress); the shared u is updated with
// red.global.add.f64; the group barrier is an arrive counter + generation word in global memory.
// Termination mirrors the reference: LOCAL = a group stops after num_cycles own corrections
// (:317-322); GLOBAL = level 0's root raises converge_flag once every level has done num_cycles
// (CheckConverge, src/Misc.cpp:418-442) and every group sees it at its next barrier (:323-337).
#include "ctx.h"
#include "kernels.cuh"
#include "async_team.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

__global__ void __launch_bounds__(kABlock, 3) k_async_amg(const AsyncParams *__restrict__ pp)
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
         if (l < L - 1) {
            AMGB_TEAM_SPMV(false, p.R[l], v.r[l], v.r[l + 1], mk(1.0, 0.0, nullptr), tm);
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
         AMGB_TEAM_SPMV(false, p.P[q], v.t[cl], v.t[q], mk(1.0, 0.0, nullptr), tm);
         group_barrier(tm);
         AMGB_TEAM_SPMV(false, p.A[q], v.t[q], v.w[q], mk(-1.0, 1.0, v.r[q]), tm);
         group_barrier(tm);
         team_smooth_zero(p, tm, q, v.w[q], v.e[q], v.t[q], p.fine_sweeps, false);
      }
      // ---- prolongation chain (:211-224)
      for (int l = idle ? -1 : q - 1; l >= 0; l--) {
         AMGB_TEAM_SPMV(false, p.P[l], v.e[l + 1], v.e[l], mk(1.0, 0.0, nullptr), tm);
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
      if (!idle) AMGB_TEAM_SPMV(false, p.A[0], v.u_local, v.r[0], mk(-1.0, 1.0, p.f), tm);
      group_barrier(tm);
      if (stop) break;
   }
}

}  // namespace

int async_max_grid(int block)
{
   int dev = 0, sms = 0, per_sm = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
   cudaFuncSetAttribute(k_async_amg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AMGB_TEAM_SMEM);
   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_async_amg, block, AMGB_TEAM_SMEM);
   return sms * per_sm;
}

int launch_async(const LaunchCfg &, cudaStream_t st, const AsyncParams *params_dev, int grid, int block,
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
   cudaError_t e = cudaLaunchKernelEx(&cfg, k_async_amg, params_dev);
   return e == cudaSuccess ? 1 : -(int)e;
}

#ifndef AMGB_ASYNC_KERNEL_ONLY   // (async_ni.cu re-includes this file for the kernel and its two launch helpers only)
// EXPERIMENTAL switch AMGB_ASYNC_NOINLINE=1: run k_async_amg_ni (async_ni.cu), the same kernel with every SpMV behind a
// non-inlined call (code size 60 616 -> a few thousand instructions, profiles/README.md section 9)
static bool async_noinline()
{
   const char *e = getenv("AMGB_ASYNC_NOINLINE");
   return e && atoi(e) != 0;
}

// ---- host side: build the parameter block once, run ---------------------------------------------
static int async_prepare(amgb_ctx *c)
{
   if (c->async_ready) return AMGB_OK;
   const int L = c->L;
   const amgb_options &o = c->opt;
   const bool multadd = o.solver == AMGB_SOLVER_ASYNC_MULTADD || o.solver == AMGB_SOLVER_MULTADD;
   // factorised level-0 transfers (EXPERIMENTAL, k_async_amg_fact0 in async_fact0.cu): plain P_0 / R_0 were uploaded
   const bool fact0 = o.factor_level0 && multadd && c->symmetric && L >= 3;
   AsyncParams hp;
   memset(&hp, 0, sizeof(hp));
   hp.num_levels = L;
   hp.solver = multadd ? AMGB_SOLVER_ASYNC_MULTADD : AMGB_SOLVER_ASYNC_AFACX;
   hp.smoother = o.smoother;
   hp.symmetric = c->symmetric ? 1 : 0;
   hp.fine_sweeps = o.num_fine_smooth_sweeps;
   hp.coarse_sweeps = o.num_coarse_smooth_sweeps;
   hp.jgs_block_rows = o.jgs_block_rows;
   for (int l = 0; l < L; l++) {
      const double avg = c->A[l].nrows > 0 ? (double)c->A[l].nnz / c->A[l].nrows : 0.0;
      hp.jgs_lpb[l] = o.jgs_block_rows > AMGB_JGS_BMAX ? 0 : (avg <= 5.0 ? 4 : (avg <= 10.0 ? 8 : (avg <= 20.0 ? 16 : 32)));
   }
   int rc;
   for (int l = 0; l < L; l++) {
      hp.A[l] = c->A[l];
      if (l < L - 1) { hp.P[l] = c->P[l]; hp.R[l] = c->R[l]; }
      hp.ws[l] = c->ws[l];
      hp.inv_l1[l] = c->inv_l1[l];
   }
   // per-group vectors (level_vector[k].{r,e,u_prev,...}[l], l <= k+1: src/SMEM_Setup.cpp:314-341)
   const int n0 = c->A[0].nrows;
   for (int q = 0; q < L; q++) {
      for (int l = 0; l <= std::min(q + 1, L - 1); l++) {
         const size_t bytes = sizeof(double) * (size_t)c->A[l].nrows;
         if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.g[q].r[l], bytes, true))) return rc;
         if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.g[q].e[l], bytes, true))) return rc;
         if (l >= q) {
            if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.g[q].t[l], bytes, true))) return rc;
            if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.g[q].w[l], bytes, true))) return rc;
         }
      }
      if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.g[q].u_local, sizeof(double) * (size_t)n0, true))) return rc;
      if (fact0 && q >= 1 && (rc = amgb_dev_alloc_bytes(c, (void **)&hp.t0[q], sizeof(double) * (size_t)n0, true))) return rc;
   }
   // CTA groups proportional to the reference's work model (ComputeWork, src/SMEM_Setup.cpp:1083-1160)
   std::vector<double> work(L, 0.0);
   double tot = 0.0;
   for (int k = 0; k < L; k++) {
      double w = (double)c->A[0].nnz + n0;
      const int coarsest = multadd ? k : k + 1;
      for (int l = 0; l < coarsest && l < L - 1; l++) w += multadd ? (double)c->R[l].nnz : (k < L - 1 ? (double)l * c->R[l].nnz : 0.0);
      if (fact0 && k >= 1) w += 2.0 * c->A[0].nnz;      // the two extra passes over A_0 of the factorised level-0 transfers
      if (k == L - 1) w += c->A[k].nnz;
      else if (multadd) w += c->symmetric ? (double)o.num_fine_smooth_sweeps * (c->A[k].nnz + c->A[k].nrows) : (double)c->A[k].nrows;
      else w += (double)(o.num_coarse_smooth_sweeps - 1) * c->A[k + 1].nnz + c->P[k].nnz + c->A[k].nnz +
                (double)(o.num_fine_smooth_sweeps - 1) * c->A[k].nnz;
      for (int l = 0; l < k; l++) w += c->P[l].nnz;
      // the coarsest level's correction is identically zero (see k_async_amg): its group only keeps the count
      // and the stop protocol, so it gets the mandatory single CTA and no share of the rest
      if (k == L - 1 && L > 1) w = 0.0;
      work[k] = w;
      tot += w;
   }
   int grid = fact0 ? async_max_grid_fact0(kABlock) : (async_noinline() ? async_max_grid_ni(kABlock) : async_max_grid(kABlock));
   if (grid < L) return amgb_fail(c, AMGB_ECUDA, "cooperative grid %d smaller than the number of levels %d", grid, L);
   std::vector<int> ctas(L, 1);
   int left = grid - L;
   std::vector<double> want(L);
   for (int k = 0; k < L; k++) want[k] = work[k] / tot * grid;
   // largest-remainder distribution on top of the mandatory one CTA per level
   for (int k = 0; k < L; k++) {
      int extra = (int)std::floor(std::max(0.0, want[k] - 1.0));
      extra = std::min(extra, left);
      ctas[k] += extra;
      left -= extra;
   }
   for (int k = 0; left > 0; k = (k + 1) % L) {
      if (work[k] == 0.0 && L > 1) continue;          // never hand spare CTAs to the idle coarsest group
      ctas[k]++; left--;
   }
   hp.cta_begin[0] = 0;
   for (int k = 0; k < L; k++) hp.cta_begin[k + 1] = hp.cta_begin[k] + ctas[k];
   c->async_cta_begin.assign(hp.cta_begin, hp.cta_begin + L + 1);
   c->async_grid = grid;
   hp.f = c->f;
   hp.u = c->u;
   if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.barrier_count, sizeof(unsigned int) * AMGB_MAX_LEVELS, true))) return rc;
   unsigned int *gen;
   if ((rc = amgb_dev_alloc_bytes(c, (void **)&gen, sizeof(unsigned int) * AMGB_MAX_LEVELS, true))) return rc;
   hp.barrier_gen = gen;
   if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.num_correct, sizeof(int) * AMGB_MAX_LEVELS, true))) return rc;
   if ((rc = amgb_dev_alloc_bytes(c, (void **)&hp.group_stop, sizeof(int) * AMGB_MAX_LEVELS, true))) return rc;
   int *flag;
   if ((rc = amgb_dev_alloc_bytes(c, (void **)&flag, sizeof(int) * 4, true))) return rc;
   hp.converge_flag = flag;
   if (o.l2_persist && c->arena_used > 0) {
      // pin the coarse hierarchy in L2 for the persistent kernel (launch attribute, see launch_async)
      c->window.base_ptr = c->arena;
      c->window.num_bytes = std::min(c->arena_used, c->max_window);
      c->window.hitRatio = 1.0f;
      c->window.hitProp = cudaAccessPropertyPersisting;
      c->window.missProp = cudaAccessPropertyStreaming;
      c->window_valid = true;
   }
   c->async_host = new AsyncParams(hp);
   if ((rc = amgb_dev_alloc_bytes(c, &c->async_params_dev, sizeof(AsyncParams), false))) return rc;
   CUDA_OK(c, cudaStreamSynchronize(c->stream));
   c->async_ready = true;
   return AMGB_OK;
}

void amgb_async_teardown(amgb_ctx *c)
{
   delete c->async_host;
   c->async_host = nullptr;
}

extern "C" int amgb_solve_async(amgb_ctx *c, int num_cycles, int converge_type, int *corrections, double *relres,
                                double *solve_seconds)
{
   NEED_READY(c);
   if (num_cycles < 1) return amgb_fail(c, AMGB_EINVAL, "num_cycles < 1");
   if (c->opt.coarse_solve) return amgb_fail(c, AMGB_EINVAL, "coarse_solve (DMEM convention) is implemented for the synchronous cycles");
   for (int *b : c->jgs_bounds)
      if (b) return amgb_fail(c, AMGB_EINVAL, "explicit hybrid-JGS block lists are implemented for the synchronous cycles");
   int rc;
   if ((rc = async_prepare(c))) return rc;
   AsyncParams &hp = *c->async_host;
   const int L = c->L, n0 = c->A[0].nrows;
   hp.num_cycles = num_cycles;
   hp.converge_type = converge_type;
   // r0 and ||r0|| (src/SMEM_Solve.cpp:60-70); every group starts from a copy of r0
   // (src/SMEM_Async_AMG.cpp:10-15) and of u
   enq_residual(c);
   double ss;
   if ((rc = amgb_fetch_scalar(c, &ss))) return rc;
   const double r0 = sqrt(ss);
   for (int q = 0; q < L; q++) {
      CUDA_OK(c, cudaMemcpyAsync(hp.g[q].r[0], c->r[0], sizeof(double) * n0, cudaMemcpyDeviceToDevice, c->stream));
      CUDA_OK(c, cudaMemcpyAsync(hp.g[q].u_local, c->u, sizeof(double) * n0, cudaMemcpyDeviceToDevice, c->stream));
   }
   CUDA_OK(c, cudaMemsetAsync(hp.barrier_count, 0, sizeof(unsigned int) * AMGB_MAX_LEVELS, c->stream));
   CUDA_OK(c, cudaMemsetAsync((void *)hp.barrier_gen, 0, sizeof(unsigned int) * AMGB_MAX_LEVELS, c->stream));
   CUDA_OK(c, cudaMemsetAsync(hp.num_correct, 0, sizeof(int) * AMGB_MAX_LEVELS, c->stream));
   CUDA_OK(c, cudaMemsetAsync(hp.group_stop, 0, sizeof(int) * AMGB_MAX_LEVELS, c->stream));
   CUDA_OK(c, cudaMemsetAsync((void *)hp.converge_flag, 0, sizeof(int) * 4, c->stream));
   CUDA_OK(c, cudaMemcpyAsync(c->async_params_dev, &hp, sizeof(AsyncParams), cudaMemcpyHostToDevice, c->stream));
   CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
   const bool fact0 = L > 1 && hp.t0[1] != nullptr;
   int lr = fact0 ? launch_async_fact0(c->cfg, c->stream, (const AsyncParams *)c->async_params_dev, c->async_grid, kABlock,
                                       c->window_valid ? &c->window : nullptr)
                  : (async_noinline() ? launch_async_ni : launch_async)(c->cfg, c->stream, (const AsyncParams *)c->async_params_dev,
                                                                        c->async_grid, kABlock, c->window_valid ? &c->window : nullptr);
   if (lr < 0) return amgb_fail(c, AMGB_ECUDA, "cooperative launch failed: %s", cudaGetErrorString((cudaError_t)(-lr)));
   c->launches += 1;
   CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
   CUDA_OK(c, cudaEventSynchronize(c->ev1));
   float ms = 0;
   cudaEventElapsedTime(&ms, c->ev0, c->ev1);
   if (solve_seconds) *solve_seconds = ms * 1e-3;
   // final residual on the shared u (src/SMEM_Solve.cpp:82-91)
   enq_residual(c);
   if ((rc = amgb_fetch_scalar(c, &ss))) return rc;
   if (relres) *relres = sqrt(ss) / r0;
   if (corrections) {
      std::vector<int> h(AMGB_MAX_LEVELS);
      CUDA_OK(c, cudaMemcpy(h.data(), hp.num_correct, sizeof(int) * AMGB_MAX_LEVELS, cudaMemcpyDeviceToHost));
      for (int l = 0; l < L; l++) corrections[l] = h[l];
   }
   CUDA_OK(c, cudaGetLastError());
   return AMGB_OK;
}

extern "C" int amgb_async_groups(amgb_ctx *c, int *cta_begin /* num_levels+1 */, int *grid)
{
   NEED_READY(c);
   int rc;
   if ((rc = async_prepare(c))) return rc;
   for (int l = 0; l <= c->L; l++) cta_begin[l] = c->async_cta_begin[l];
   if (grid) *grid = c->async_grid;
   return AMGB_OK;
}
#endif   // AMGB_ASYNC_KERNEL_ONLY
