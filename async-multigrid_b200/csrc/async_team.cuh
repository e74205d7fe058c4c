// async_team.cuh -- device helpers shared by the persistent asynchronous kernels (async.cu, async_fact0.cu): the CTA-group
// ("team") descriptor, the group barrier (= the reference's SMEM_LevelBarrier) and the zero-guess smoother dispatch.
#pragma once
#include "ctx.h"
#include "kernels.cuh"

// every SpMV of the persistent kernels goes through this macro: the default expands to the inlined spmv_team (the measured
// round-1 kernels); async_ni.cu redefines it to a non-inlined call to keep the kernel's code inside the instruction cache
#ifndef AMGB_TEAM_SPMV
#define AMGB_TEAM_SPMV(SVAL, M, x, y, e, tm) spmv_team<false, SVAL>(M, x, y, e, (tm).tid, (tm).size, false, (tm).smem)
#endif

namespace {

constexpr int kABlock = 256;

struct Team {
   int tid, size;          // thread index / thread count within the level's CTA group
   int cta, nctas;         // CTA index / count within the group
   unsigned int *count;
   volatile unsigned int *gen;
   unsigned char *smem;    // AMGB_TEAM_SMEM bytes of shared memory (CSR-stream staging)
};

// barrier among the CTAs of one group (the reference's SMEM_LevelBarrier)
__device__ __forceinline__ void group_barrier(const Team &tm)
{
   __syncthreads();
   if (tm.nctas > 1) {
      if (threadIdx.x == 0) {
         __threadfence();
         const unsigned int g = *tm.gen;
         const unsigned int prev = atomicAdd(tm.count, 1u);
         if (prev == (unsigned int)tm.nctas - 1u) {
            atomicExch(tm.count, 0u);
            __threadfence();
            atomicAdd((unsigned int *)tm.gen, 1u);
         } else {
            while (*tm.gen == g) __nanosleep(32);
         }
         __threadfence();
      }
      __syncthreads();
   }
}

__device__ __forceinline__ SpmvEpilogue mk(double alpha, double beta, const double *b, double gamma = 0.0,
                                           const double *c = nullptr, const double *rs = nullptr)
{
   SpmvEpilogue e;
   e.alpha = alpha; e.beta = beta; e.gamma = gamma; e.b = b; e.c = c; e.rs = rs;
   return e;
}

// e = S_l f from a zero guess (the dispatch of SMEM_Smooth, src/SMEM_Solve.cpp:264-323), ends with
// a group barrier.  s1: scratch vector of level l.
__device__ void team_smooth_zero(const AsyncParams &p, const Team &tm, int l, const double *f, double *e,
                                 double *s1, int sweeps, bool symmetric)
{
   const DevCSR &A = p.A[l];
   const int n = A.nrows;
   if (p.smoother == AMGB_SMOOTH_ASYNC_GS || p.smoother == AMGB_SMOOTH_SEMI_ASYNC_GS) {
      for (int i = tm.tid; i < n; i += tm.size) st_cg(e + i, 0.0);
      group_barrier(tm);
      if (p.smoother == AMGB_SMOOTH_ASYNC_GS) {
         async_gs_team<false>(A, f, e, p.jgs_block_rows, sweeps, tm.tid, tm.size);
         group_barrier(tm);
      } else {
         for (int k = 0; k < sweeps; k++) {
            async_gs_team<false>(A, f, e, p.jgs_block_rows, 1, tm.tid, tm.size);
            group_barrier(tm);
         }
      }
      return;
   }
   if (p.smoother == AMGB_SMOOTH_HYBRID_JGS) {
      auto sweep = [&](const double *uprev, bool zero) {
         double *su = reinterpret_cast<double *>(tm.smem);
         switch (p.jgs_lpb[l]) {   // sub-warp per block (see hybrid_jgs_subwarp_team); 0: block longer than the staging slice
            case 4: hybrid_jgs_subwarp_team<false, 4>(A, f, e, uprev, nullptr, p.jgs_block_rows, zero, tm.tid, tm.size, su); break;
            case 8: hybrid_jgs_subwarp_team<false, 8>(A, f, e, uprev, nullptr, p.jgs_block_rows, zero, tm.tid, tm.size, su); break;
            case 16: hybrid_jgs_subwarp_team<false, 16>(A, f, e, uprev, nullptr, p.jgs_block_rows, zero, tm.tid, tm.size, su); break;
            case 32: hybrid_jgs_subwarp_team<false, 32>(A, f, e, uprev, nullptr, p.jgs_block_rows, zero, tm.tid, tm.size, su); break;
            default: hybrid_jgs_team<false>(A, f, e, uprev, nullptr, p.jgs_block_rows, zero, tm.tid, tm.size);
         }
      };
      sweep(nullptr, true);
      group_barrier(tm);
      for (int k = 1; k < sweeps; k++) {
         for (int i = tm.tid; i < n; i += tm.size) s1[i] = ld_cg(e + i);
         group_barrier(tm);
         sweep(s1, false);
         group_barrier(tm);
      }
      return;
   }
   const double *rs = (p.smoother == AMGB_SMOOTH_L1_JACOBI) ? p.inv_l1[l] : p.ws[l];
   if (symmetric) {
      AMGB_TEAM_SPMV(true, A, f, e, mk(-1.0, 2.0, f, 0.0, nullptr, rs), tm);
      group_barrier(tm);
      for (int k = 1; k < sweeps; k++) {
         AMGB_TEAM_SPMV(false, A, e, s1, mk(-1.0, 1.0, f), tm);
         group_barrier(tm);
         AMGB_TEAM_SPMV(true, A, s1, e, mk(-1.0, 2.0, s1, 0.0, nullptr, rs), tm);
         group_barrier(tm);
      }
      return;
   }
   double *cur = ((sweeps - 1) & 1) ? s1 : e;
   double *oth = (cur == e) ? s1 : e;
   for (int i = tm.tid; i < n; i += tm.size) cur[i] = __ldg(rs + i) * ld_cg(f + i);
   group_barrier(tm);
   for (int k = 1; k < sweeps; k++) {
      AMGB_TEAM_SPMV(false, A, cur, oth, mk(-1.0, 1.0, f, 1.0, cur, rs), tm);
      group_barrier(tm);
      double *tmp = cur; cur = oth; oth = tmp;
   }
}

}  // namespace
