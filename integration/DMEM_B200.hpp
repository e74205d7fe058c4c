// DMEM_B200.hpp -- the binding a maintainer of jwp3/async-multigrid adds to the reference's DMEM driver:
// one MPI rank drives one GPU.  The hierarchy is handed over once, after DMEM_Setup; DMEM_Add's loop (src/DMEM_Add.cpp:101-130)
// -- synchronous, accelerated or asynchronous -- becomes one call.
//
// A rank passes its LOCAL ROW BLOCK of every level: hypre_ParCSRMatrixDiag / Offd merged into one CSR whose columns are
// numbered in the rank's extended [ghost_lo | owned | ghost_hi] index space (col_map_offd names the ghost columns; DESIGN.md
// section 6), plus the ints of amgb_dist_set_level.  The merge is the maintainer's (it needs hypre's accessors); this file takes
// the result as plain arrays.
//
// This very file is compiled against the reference's own DMEM_Main.hpp (DMEM_AllData) by oracle/build_ref.sh, with the
// single-process MPI of oracle/ref_shim, and driven on one rank by tests/test_zz_gpu_extended.py.
#ifndef DMEM_B200_HPP
#define DMEM_B200_HPP
#include "DMEM_Main.hpp"
#include "amg_b200.h"

struct B200Layout {                 // one level on this rank (amgb_dist_set_level)
   int n_global, row_start, n_owned, halo_lo, halo_hi, distributed, send_lo, send_hi;
   const int *all_owned;            // n_owned of every rank
};
struct B200Csr { int nrows, ncols, nnz; const int *row_ptr, *col_idx; const double *values; };

static void DMEM_B200_Check(amgb_ctx *ctx, int rc){
   if (rc != AMGB_OK){ printf("amg_b200: %s\n", ctx ? amgb_last_error(ctx) : "no context"); exit(1); }
}

// call once after DMEM_Setup(dmem_all_data)
static amgb_ctx *DMEM_B200_Upload(DMEM_AllData *dmem_all_data, int gpu, int L, const B200Layout *lay,
                                  const B200Csr *A, const B200Csr *P, const B200Csr *R)
{
   int rank, size;
   MPI_Comm_rank(MPI_COMM_WORLD, &rank);
   MPI_Comm_size(MPI_COMM_WORLD, &size);
   unsigned char id[128];
   if (rank == 0) DMEM_B200_Check(NULL, amgb_dist_unique_id(id));
   MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD);                    // the NCCL id travels over the reference's own MPI
   amgb_ctx *ctx = NULL;
   DMEM_B200_Check(NULL, amgb_create(&ctx, gpu));
   DMEM_B200_Check(ctx, amgb_dist_init(ctx, id, rank, size));
   amgb_options opt; amgb_default_options(&opt);
   const int solver = dmem_all_data->input.solver;                     // src/Main.hpp:61-70
   opt.solver = (solver == AFACX || solver == ASYNC_AFACX || solver == SYNC_AFACX) ? AMGB_SOLVER_AFACX : AMGB_SOLVER_MULTADD;
   opt.smoother = dmem_all_data->input.smoother == L1_JACOBI ? AMGB_SMOOTH_L1_JACOBI : AMGB_SMOOTH_JACOBI;
   opt.smooth_weight = dmem_all_data->input.smooth_weight;
   opt.num_post_smooth_sweeps = dmem_all_data->input.simple_jacobi_flag == 0 ? 0 : 1;   // symmetrised smoother unless plain Jacobi was asked for
   opt.coarse_solve = opt.solver == AMGB_SOLVER_MULTADD;               // DMEM solves the coarsest grid directly (src/DMEM_Add.cpp:262-264)
   DMEM_B200_Check(ctx, amgb_set_options(ctx, &opt));
   DMEM_B200_Check(ctx, amgb_set_num_levels(ctx, L));
   for (int l = 0; l < L; l++)
      DMEM_B200_Check(ctx, amgb_dist_set_level(ctx, l, lay[l].n_global, lay[l].row_start, lay[l].n_owned, lay[l].halo_lo, lay[l].halo_hi,
                                               lay[l].distributed, lay[l].send_lo, lay[l].send_hi, lay[l].all_owned));
   for (int l = 0; l < L; l++){
      DMEM_B200_Check(ctx, amgb_set_matrix(ctx, AMGB_MAT_A, l, A[l].nrows, A[l].ncols, A[l].nnz, A[l].row_ptr, A[l].col_idx, A[l].values));
      if (l < L-1){
         DMEM_B200_Check(ctx, amgb_set_matrix(ctx, AMGB_MAT_P, l, P[l].nrows, P[l].ncols, P[l].nnz, P[l].row_ptr, P[l].col_idx, P[l].values));
         DMEM_B200_Check(ctx, amgb_set_matrix(ctx, AMGB_MAT_R, l, R[l].nrows, R[l].ncols, R[l].nnz, R[l].row_ptr, R[l].col_idx, R[l].values));
      }
   }
   DMEM_B200_Check(ctx, amgb_setup(ctx));
   DMEM_B200_Check(ctx, amgb_dist_setup(ctx));
   return ctx;
}

// replaces the loop of DMEM_Add(dmem_all_data)                      (src/DMEM_Add.cpp:101-130)
// f_local / u_local: this rank's rows; hist (may be NULL): num_cycles + 1 relative residual norms of the synchronous solve;
// corrections (may be NULL): per level, asynchronous solve.  Returns the global relative residual.
static double DMEM_Add_B200(DMEM_AllData *dmem_all_data, amgb_ctx *ctx, const double *f_local, double *u_local, double *hist, int *corrections)
{
   double relres = 0.0, secs = 0.0;
   int n = 0;
   DMEM_B200_Check(ctx, amgb_dist_set_rhs(ctx, f_local));
   if (dmem_all_data->input.async_flag){
      // every level group of every rank performs num_cycles corrections (-converge_test_type local); nobody waits for another group
      int cor[32];
      DMEM_B200_Check(ctx, amgb_dist_solve_async(ctx, dmem_all_data->input.num_cycles, cor, &relres, &secs));
      n = cor[0];
      if (corrections) for (int l = 0; l < dmem_all_data->grid.num_levels; l++) corrections[l] = cor[l];
   } else {
      // -cheby / -richard: DMEM_ChebyUpdate inside DMEM_SyncAddCorrect (src/DMEM_Add.cpp:706-711); accel_type 0: plain
      std::vector<double> h(dmem_all_data->input.num_cycles + 1, 0.0);
      DMEM_B200_Check(ctx, amgb_dist_solve_sync_accel(ctx, dmem_all_data->input.tol, dmem_all_data->input.num_cycles,
                                                      dmem_all_data->input.accel_type, dmem_all_data->cheby.mu, dmem_all_data->cheby.delta,
                                                      h.data(), &n, &secs));
      relres = h[n];
      if (hist) for (int k = 0; k <= n; k++) hist[k] = h[k];
   }
   DMEM_B200_Check(ctx, amgb_dist_get_solution(ctx, u_local));
   dmem_all_data->iter.cycle = n;
   dmem_all_data->output.solve_wtime = secs;                          // what DMEM_PrintOutput prints
   return relres;
}
#endif
