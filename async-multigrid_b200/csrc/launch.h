// launch.h -- host-side launchers of the stand-alone kernels (kernels.cu) and of the persistent
// asynchronous kernel (async.cu).  Every launcher returns the number of kernels it launched.
#pragma once
#include "common.cuh"

struct LaunchCfg {
   int num_sms = 148;
   int ctas_per_sm = 8;
   int stream_variant = 0;
};

// geometry of the CSR-stream kernel (kernels.cuh stream_rows_team): threads per CTA, entries per row block,
// staged-x window capacity (0 = gather x from global memory), bulk-copy stages.  amgb_options.stream_variant
// selects one; the row blocks are built for its cap / xcap at upload time.
struct StreamVariant { int nt, cap, xcap, stages; };
// stages == 0: the warp-granular variant (warp_stream_rows_team), cap = entries per warp chunk
// stages == -1: warp-granular with column-sorted chunks (DevCSR::pci/pva/pos)
#define AMGB_NUM_STREAM_VARIANTS 13
static const StreamVariant kStreamVariants[AMGB_NUM_STREAM_VARIANTS] = {
   {256, 2048, 0, 3}, {128, 1024, 0, 2}, {128, 1024, 0, 3}, {256, 2048, 1536, 2},
   {128, 1024, 1024, 2}, {256, 1024, 0, 2}, {128, 1024, 1024, 3}, {128, 2048, 0, 2},
   {256, 256, 0, 0}, {256, 128, 0, 0}, {256, 512, 0, 0}, {256, 256, 0, -1}, {256, 128, 0, -1},
};

// y = gamma*c + rs.*(beta*b + alpha*M*x); optional sum of y_i^2 into partial sums (one per CTA,
// `partials` must hold >= grid entries; *grid_out receives the grid used).
int launch_spmv(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &M, bool use_sval, const double *x, double *y,
                const SpmvEpilogue &e, double *partials, int *grid_out);
int spmv_units(const DevCSR &M);
int launch_spmv_units(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &M, int u0, int u1, bool use_sval, const double *x,
                      double *y, const SpmvEpilogue &e, double *partials, int *grid_out);
// out[0] = sum(partials[0..n)); if hist != nullptr: hist[k] = sqrt(out[0]) (and r0 handling on host)
int launch_reduce_partials(cudaStream_t st, const double *partials, int n, double *out);
// y = a.*x  (zero-guess Jacobi: u = (w/d).*f, src/SMEM_Smooth.cpp:381-389)
int launch_scale(const LaunchCfg &cfg, cudaStream_t st, int n, const double *a, const double *x, double *y);
// y += x
int launch_add(const LaunchCfg &cfg, cudaStream_t st, int n, const double *x, double *y);
// Chebyshev update (src/SMEM_Solve.cpp:179-187): uo_prev=uo; uo = yo + omega*(delta*c + uo - yo); yo = uo_prev; u = uo
int launch_cheby(const LaunchCfg &cfg, cudaStream_t st, int n, double omega, double delta, const double *c,
                 double *u_outer, double *y_outer, double *u);
// y = a*x + b*y [, z = y]; partial sums of x_i*y_i
int launch_axpby(const LaunchCfg &cfg, cudaStream_t st, int n, double a, const double *x, double b, double *y, double *z);
int launch_dot(const LaunchCfg &cfg, cudaStream_t st, int n, const double *x, const double *y, double *partials, int *grid_out);
// u += e here and on every peer GPU (IPC-mapped pointers)
#define AMGB_MAX_PEERS 15
struct PeerPtrs { double *p[AMGB_MAX_PEERS]; int n; };
int launch_push_correction(const LaunchCfg &cfg, cudaStream_t st, int n, const double *e, double *u, const PeerPtrs &peers);
// partial sums of x_i^2
int launch_sumsq(const LaunchCfg &cfg, cudaStream_t st, int n, const double *x, double *partials, int *grid_out);
// hybrid JGS sweep
int launch_hybrid_jgs(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &A, const double *f, double *u,
                      const double *u_prev, const double *scale, int block_rows, bool zero_guess);
// the same sweep over an explicit block list bounds[0..nblocks] (amgb_set_jgs_blocks: the reference's blocks are thread row ranges)
int launch_hybrid_jgs_list(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &A, const double *f, double *u,
                           const double *u_prev, const double *scale, const int *bounds, int nblocks, bool zero_guess);
// (semi-)asynchronous Gauss-Seidel sweeps on u (in place); y = M^T x
int launch_async_gs(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &A, const double *f, double *u, int block_rows,
                    int sweeps, bool semi);
int launch_spmv_transpose(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &M, const double *x, double *y);
// setup helpers: ws = w/d (0 where d == 0), l1 = sum |a_ij|, sval = va .* cs[col]
int launch_diag_scale(cudaStream_t st, const DevCSR &A, double w, double *ws, double *dow);
int launch_l1(cudaStream_t st, const DevCSR &A, double *l1, double *inv_l1);
int launch_colscale(cudaStream_t st, int nnz, const int *ci, const double *va, const double *cs, double *out);

// ---- persistent asynchronous kernel (async.cu) ---------------------------------------------------
#define AMGB_MAX_LEVELS 32
struct AsyncLevelVecs {           // per CTA group (the reference's level_vector[k], src/SMEM_Setup.cpp:314-341)
   double *r[AMGB_MAX_LEVELS];    // residual chain r_0 .. r_k
   double *e[AMGB_MAX_LEVELS];    // correction chain e_k .. e_0
   double *t[AMGB_MAX_LEVELS];    // scratch (u_prev / AFACx work)
   double *w[AMGB_MAX_LEVELS];    // scratch (AFACx r_fine)
   double *u_local;               // the group's private copy of the fine solution
};
struct AsyncParams {
   int num_levels;
   int solver;                    // AMGB_SOLVER_ASYNC_MULTADD / ASYNC_AFACX
   int smoother;
   int symmetric;
   int fine_sweeps, coarse_sweeps;
   int jgs_block_rows;
   int jgs_lpb[AMGB_MAX_LEVELS];  // lanes per hybrid-JGS block on every level (4/8/16/32 from the mean row length; 0: one thread per block)
   int num_cycles;
   int converge_type;
   DevCSR A[AMGB_MAX_LEVELS], P[AMGB_MAX_LEVELS], R[AMGB_MAX_LEVELS];
   const double *ws[AMGB_MAX_LEVELS];       // w/d
   const double *inv_l1[AMGB_MAX_LEVELS];   // 1/l1
   AsyncLevelVecs g[AMGB_MAX_LEVELS];
   int cta_begin[AMGB_MAX_LEVELS + 1];      // CTA range of every level's group
   const double *f;
   double *u;                               // shared fine solution (atomic adds)
   unsigned int *barrier_count;             // [L] arrive counters
   volatile unsigned int *barrier_gen;      // [L] generations
   int *num_correct;                        // [L] local_num_correct
   int *group_stop;                         // [L] the group root's stop decision (GLOBAL rule)
   volatile int *converge_flag;             // thread.converge_flag
   // (appended last so that the offsets the round-1 kernel reads stay what they were)
   double *t0[AMGB_MAX_LEVELS];             // per group: level-0 scratch of the factorised level-0 transfers (k_async_amg_fact0)
};
int launch_async(const LaunchCfg &cfg, cudaStream_t st, const AsyncParams *params_dev, int grid, int block,
                 const cudaAccessPolicyWindow *window);
int async_max_grid(int block);
// experimental copy with the level-0 transfers in factorised form (async_fact0.cu)
int launch_async_fact0(const LaunchCfg &cfg, cudaStream_t st, const AsyncParams *params_dev, int grid, int block,
                       const cudaAccessPolicyWindow *window);
int async_max_grid_fact0(int block);
// experimental copy with every SpMV behind a non-inlined call (async_ni.cu; AMGB_ASYNC_NOINLINE=1)
int launch_async_ni(const LaunchCfg &cfg, cudaStream_t st, const AsyncParams *params_dev, int grid, int block,
                    const cudaAccessPolicyWindow *window);
int async_max_grid_ni(int block);
