// dist.cu -- multi-GPU (DMEM replacement) entry points.  Round-1 state: the row-partitioned
// NCCL path is declared in include/amg_b200.h; until it lands these report AMGB_ESTATE so a
// caller can never mistake a missing implementation for a result.
#include "ctx.h"
#include <cstring>
#ifdef AMG_HAVE_NCCL
#include <nccl.h>
#endif

struct DistState { int rank = 0, nranks = 1; };

void amgb_dist_teardown(amgb_ctx *c)
{
   if (c && c->dist) { delete c->dist; c->dist = nullptr; }
}

extern "C" {
int amgb_dist_unique_id(unsigned char id128[128])
{
   (void)id128;
   return AMGB_ESTATE;
}
int amgb_dist_init(amgb_ctx *c, const unsigned char id128[128], int rank, int nranks)
{
   (void)id128; (void)rank; (void)nranks;
   return amgb_fail(c, AMGB_ESTATE, "distributed path not implemented in this build");
}
int amgb_dist_set_partition(amgb_ctx *c, int level, const int *row_starts)
{
   (void)level; (void)row_starts;
   return amgb_fail(c, AMGB_ESTATE, "distributed path not implemented in this build");
}
int amgb_dist_solve_sync(amgb_ctx *c, double tol, int max_cycles, double *relres_hist, int *n_cycles, double *solve_seconds)
{
   (void)tol; (void)max_cycles; (void)relres_hist; (void)n_cycles; (void)solve_seconds;
   return amgb_fail(c, AMGB_ESTATE, "distributed path not implemented in this build");
}
}
