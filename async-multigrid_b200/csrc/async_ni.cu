// async_ni.cu -- EXPERIMENTAL (no default path uses it; AMGB_ASYNC_NOINLINE=1 selects it): the persistent asynchronous kernel of
// async.cu with every SpMV of the chains behind ONE non-inlined device function per epilogue flavour.
//
// Why: k_async_amg inlines the whole storage-format dispatch of spmv_team at each of its call sites and comes out at
// 60 616 SASS instructions (~0.97 MB), several times the instruction cache that the 3 resident CTAs of different level
// groups share (profiles/README.md section 9).  Here the kernel body is the same source (this file re-includes async.cu with
// the call macro redefined), so the arithmetic and the protocol are identical; only the code layout differs.  Its own
// translation unit, so that the SASS of the measured round-1 kernel stays what it was.
#include "ctx.h"
#include "kernels.cuh"

namespace {

template <bool SVAL>
__device__ __noinline__ void spmv_team_ni(const DevCSR *M, const double *x, double *y, const SpmvEpilogue *e, int team_tid,
                                          int team_size, unsigned char *smem)
{
   spmv_team<false, SVAL>(*M, x, y, *e, team_tid, team_size, false, smem);
}

}  // namespace

#define AMGB_TEAM_SPMV(SVAL, M, x, y, e, tm)                                                               \
   do {                                                                                                   \
      const SpmvEpilogue amgb_epi_ = (e);                                                                 \
      spmv_team_ni<SVAL>(&(M), x, y, &amgb_epi_, (tm).tid, (tm).size, (tm).smem);                         \
   } while (0)

#define AMGB_ASYNC_KERNEL_ONLY
#define k_async_amg k_async_amg_ni
#define async_max_grid async_max_grid_ni
#define launch_async launch_async_ni
#include "async.cu"
