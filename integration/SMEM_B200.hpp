// SMEM_B200.hpp -- the binding a maintainer of jwp3/async-multigrid adds to the reference's SMEM driver:
// ~70 lines, no new dependencies.  SMEM_Solve is the only function that has to change: everything it calls per thread
// collapses into one call made by the master thread; the hierarchy is uploaded once, right after SMEM_Setup returns.
//
// This very file is compiled against the reference's own Main.hpp (AllData, hypre_CSRMatrix accessors) by
// oracle/build_ref.sh and driven by tests/test_zz_gpu_extended.py::test_reference_structs_drive_the_library.
#ifndef SMEM_B200_HPP
#define SMEM_B200_HPP
#include "Main.hpp"
#include "amg_b200.h"
#include <vector>

static amgb_ctx *b200 = NULL;
static std::vector<double> b200_hist;            // relative residual history of the last solve (-print_reshist)

static void B200_Check(int rc){
   if (rc != AMGB_OK){ printf("amg_b200: %s\n", amgb_last_error(b200)); MPI_Finalize(); exit(1); }
}

// call once after SMEM_Setup(all_data)                               (src/SMEM_Main.cpp:686)
static void SMEM_B200_Upload(AllData *all_data)
{
   int L = all_data->grid.num_levels;
   if (b200){ amgb_destroy(b200); b200 = NULL; }
   B200_Check(amgb_create(&b200, 0));
   amgb_options opt; amgb_default_options(&opt);
   opt.solver   = all_data->input.solver;          // same #define values (src/Main.hpp:47-75)
   opt.smoother = all_data->input.smoother;
   opt.smooth_weight = all_data->input.smooth_weight;
   opt.num_pre_smooth_sweeps    = all_data->input.num_pre_smooth_sweeps;
   opt.num_post_smooth_sweeps   = all_data->input.num_post_smooth_sweeps;
   opt.num_fine_smooth_sweeps   = all_data->input.num_fine_smooth_sweeps;
   opt.num_coarse_smooth_sweeps = all_data->input.num_coarse_smooth_sweeps;
   opt.async_type       = all_data->input.async_type == SEMI_ASYNC;      // -async_type        (src/SMEM_Main.cpp:472-481)
   opt.res_compute_type = all_data->input.res_compute_type == GLOBAL;    // -res_compute_type  (:458-467)
   opt.read_type        = all_data->input.read_type == READ_RES;         // -read_type         (:507-516)
   B200_Check(amgb_set_options(b200, &opt));
   B200_Check(amgb_set_num_levels(b200, L));
   for (int l = 0; l < L; l++){
      hypre_CSRMatrix *A = all_data->matrix.A[l];
      B200_Check(amgb_set_matrix(b200, AMGB_MAT_A, l, hypre_CSRMatrixNumRows(A), hypre_CSRMatrixNumCols(A),
                                 hypre_CSRMatrixNumNonzeros(A), hypre_CSRMatrixI(A), hypre_CSRMatrixJ(A),
                                 hypre_CSRMatrixData(A)));
      if (l < L-1){
         hypre_CSRMatrix *P = all_data->matrix.P[l], *R = all_data->matrix.R[l];   // P-bar / R-bar for MULTADD
         B200_Check(amgb_set_matrix(b200, AMGB_MAT_P, l, hypre_CSRMatrixNumRows(P), hypre_CSRMatrixNumCols(P),
                                    hypre_CSRMatrixNumNonzeros(P), hypre_CSRMatrixI(P), hypre_CSRMatrixJ(P),
                                    hypre_CSRMatrixData(P)));
         B200_Check(amgb_set_matrix(b200, AMGB_MAT_R, l, hypre_CSRMatrixNumRows(R), hypre_CSRMatrixNumCols(R),
                                    hypre_CSRMatrixNumNonzeros(R), hypre_CSRMatrixI(R), hypre_CSRMatrixJ(R),
                                    hypre_CSRMatrixData(R)));
      }
   }
   B200_Check(amgb_setup(b200));
}

// replaces the body of SMEM_Solve(all_data)                          (src/SMEM_Solve.cpp:11-262)
static void SMEM_Solve_B200(AllData *all_data)
{
   int cycles = 0;
   double relres = 0, secs = 0;
   b200_hist.assign(all_data->input.num_cycles + 1, 0.0);
   B200_Check(amgb_smem_solve(b200, all_data->vector.f[0], all_data->vector.u[0], all_data->input.tol,
                              all_data->input.num_cycles, b200_hist.data(), &cycles,
                              all_data->grid.local_num_correct, &relres, &secs));
   all_data->output.solve_wtime = secs;                      // what PrintOutput prints (src/Misc.cpp:141-188)
   all_data->output.num_cycles  = cycles;
   all_data->output.r_norm2     = relres;  all_data->output.r0_norm2 = 1.0;
   if (all_data->input.print_reshist_flag && !all_data->input.async_flag)
      for (int k = 0; k <= cycles; k++) printf("%d\t%e\n", k, b200_hist[k]);   // src/SMEM_Solve.cpp:232-239
}
#endif
