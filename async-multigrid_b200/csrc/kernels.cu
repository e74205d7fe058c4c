// kernels.cu -- stand-alone sm_100a kernels of the additive-AMG solve phase and their launchers.
// Reference loops replaced: SMEM_MatVec / SMEM_SpGEMV / SMEM_Residual (src/SMEM_MatVec.cpp:123-259,
// 302-378), SMEM_Sync_Jacobi / L1Jacobi / SymmetricJacobi / HybridJacobiGaussSeidel
// (src/SMEM_Smooth.cpp:365-443,533-586,643-762), the Chebyshev update and norm of SMEM_Solve
// (src/SMEM_Solve.cpp:179-187,199-203).  All of them are HBM-bound streaming kernels: no tensor
// cores (nothing is a dense contraction).
#include "kernels.cuh"
#include "launch.h"
#include <algorithm>

namespace {

constexpr int kBlock = 256;

// One kernel per storage scheme (separate register budgets): MODE 0 = vector-per-row CSR,
// 1 = sliced ELL (stencil levels), 2 = CSR-stream (row blocks through shared memory).
template <int MODE, bool SVAL>
__global__ void __launch_bounds__(kBlock) k_spmv(DevCSR M, const double *x, double *y, SpmvEpilogue e, double *partials)
{
   const int tid = blockIdx.x * kBlock + threadIdx.x;
   const int tsz = gridDim.x * kBlock;
   const bool norm = partials != nullptr;
   double ss;
   if (MODE == 1) ss = sell_rows_team<true, SVAL>(M, x, y, e, tid, tsz, norm);
   else if (MODE == 2) {
      extern __shared__ __align__(128) unsigned char dyn_smem[];
      ss = stream_rows_team<true, SVAL>(M, x, y, e, blockIdx.x, gridDim.x, dyn_smem, norm);
   } else ss = csr_rows_dispatch<true, SVAL>(M, x, y, e, tid, tsz, norm);
   if (norm) {
      ss = block_sum(ss);
      if (threadIdx.x == 0) partials[blockIdx.x] = ss;
   }
}

__global__ void __launch_bounds__(1024) k_reduce_partials(const double *partials, int n, double *out)
{
   double s = 0.0;
   for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
   s = block_sum(s);
   if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(kBlock) k_scale(int n, const double *__restrict__ a, const double *__restrict__ x, double *__restrict__ y)
{
   for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) y[i] = a[i] * x[i];
}

__global__ void __launch_bounds__(kBlock) k_add(int n, const double *__restrict__ x, double *__restrict__ y)
{
   for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) y[i] += x[i];
}

__global__ void __launch_bounds__(kBlock) k_cheby(int n, double omega, double delta, const double *__restrict__ c,
                                                   double *__restrict__ uo, double *__restrict__ yo, double *__restrict__ u)
{
   for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
      const double prev = uo[i];
      const double v = yo[i] + omega * (delta * c[i] + prev - yo[i]);
      uo[i] = v;
      yo[i] = prev;
      u[i] = v;
   }
}

__global__ void __launch_bounds__(kBlock) k_sumsq(int n, const double *__restrict__ x, double *partials)
{
   double s = 0.0;
   for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) s += x[i] * x[i];
   s = block_sum(s);
   if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kBlock) k_hybrid_jgs(DevCSR A, const double *f, double *u, const double *u_prev,
                                                        const double *scale, int B, int zero_guess)
{
   hybrid_jgs_team<true>(A, f, u, u_prev, scale, B, zero_guess != 0, blockIdx.x * kBlock + threadIdx.x, gridDim.x * kBlock);
}

__global__ void __launch_bounds__(kBlock) k_diag_scale(DevCSR A, double w, double *ws, double *dow)
{
   for (int i = blockIdx.x * kBlock + threadIdx.x; i < A.nrows; i += gridDim.x * kBlock) {
      const double d = (A.rp[i + 1] > A.rp[i]) ? A.va[A.rp[i]] : 0.0;
      ws[i] = (d != 0.0) ? w / d : 0.0;
      if (dow) dow[i] = d / w;
   }
}

__global__ void __launch_bounds__(kBlock) k_l1(DevCSR A, double *l1, double *inv_l1)
{
   for (int i = blockIdx.x * kBlock + threadIdx.x; i < A.nrows; i += gridDim.x * kBlock) {
      double s = 0.0;
      for (int p = A.rp[i]; p < A.rp[i + 1]; p++) s += fabs(A.va[p]);
      l1[i] = s;
      inv_l1[i] = (s != 0.0) ? 1.0 / s : 0.0;
   }
}

__global__ void __launch_bounds__(kBlock) k_colscale(int nnz, const int *__restrict__ ci, const double *__restrict__ va,
                                                      const double *__restrict__ cs, double *__restrict__ out)
{
   for (int p = blockIdx.x * kBlock + threadIdx.x; p < nnz; p += gridDim.x * kBlock) out[p] = va[p] * cs[ci[p]];
}

inline int grid_for(const LaunchCfg &cfg, long work_threads)
{
   long g = (work_threads + kBlock - 1) / kBlock;
   long cap = (long)cfg.num_sms * cfg.ctas_per_sm;
   return (int)std::max(1L, std::min(g, cap));
}

}  // namespace

template <int MODE, bool SVAL>
static int resident_grid(const LaunchCfg &cfg)
{
   static int per_sm = 0;   // co-resident CTAs per SM of this instantiation: grids are sized to exactly one wave
   if (per_sm == 0) {
      int v = 0;
      const size_t dyn = MODE == 2 ? AMGB_STREAM_SMEM : 0;
      if (dyn) cudaFuncSetAttribute(k_spmv<MODE, SVAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_spmv<MODE, SVAL>, kBlock, dyn) != cudaSuccess || v < 1) v = 2;
      per_sm = v;
   }
   return cfg.num_sms * per_sm;
}

template <int MODE>
static int launch_spmv_mode(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &M, bool use_sval, const double *x, double *y,
                            const SpmvEpilogue &e, double *partials, long ctas_of_work)
{
   const int cap = use_sval ? resident_grid<MODE, true>(cfg) : resident_grid<MODE, false>(cfg);
   const int grid = (int)std::max(1L, std::min(ctas_of_work, (long)cap));
   const size_t dyn = MODE == 2 ? AMGB_STREAM_SMEM : 0;
   if (use_sval) k_spmv<MODE, true><<<grid, kBlock, dyn, st>>>(M, x, y, e, partials);
   else k_spmv<MODE, false><<<grid, kBlock, dyn, st>>>(M, x, y, e, partials);
   return grid;
}

int launch_spmv(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &M, bool use_sval, const double *x, double *y,
                const SpmvEpilogue &e, double *partials, int *grid_out)
{
   int grid;
   if (M.sell_slices > 0) grid = launch_spmv_mode<1>(cfg, st, M, use_sval, x, y, e, partials, ((long)M.sell_slices * 32 + kBlock - 1) / kBlock);
   else if (M.nblk > 0) grid = launch_spmv_mode<2>(cfg, st, M, use_sval, x, y, e, partials, (long)M.nblk);
   else grid = launch_spmv_mode<0>(cfg, st, M, use_sval, x, y, e, partials, ((long)M.nrows * M.lpr + kBlock - 1) / kBlock);
   if (grid_out) *grid_out = grid;
   return 1;
}

int launch_reduce_partials(cudaStream_t st, const double *partials, int n, double *out)
{
   k_reduce_partials<<<1, 1024, 0, st>>>(partials, n, out);
   return 1;
}

int launch_scale(const LaunchCfg &cfg, cudaStream_t st, int n, const double *a, const double *x, double *y)
{
   k_scale<<<grid_for(cfg, n), kBlock, 0, st>>>(n, a, x, y);
   return 1;
}

int launch_add(const LaunchCfg &cfg, cudaStream_t st, int n, const double *x, double *y)
{
   k_add<<<grid_for(cfg, n), kBlock, 0, st>>>(n, x, y);
   return 1;
}

int launch_cheby(const LaunchCfg &cfg, cudaStream_t st, int n, double omega, double delta, const double *c,
                 double *u_outer, double *y_outer, double *u)
{
   k_cheby<<<grid_for(cfg, n), kBlock, 0, st>>>(n, omega, delta, c, u_outer, y_outer, u);
   return 1;
}

int launch_sumsq(const LaunchCfg &cfg, cudaStream_t st, int n, const double *x, double *partials, int *grid_out)
{
   int grid = grid_for(cfg, n);
   if (grid_out) *grid_out = grid;
   k_sumsq<<<grid, kBlock, 0, st>>>(n, x, partials);
   return 1;
}

int launch_hybrid_jgs(const LaunchCfg &cfg, cudaStream_t st, const DevCSR &A, const double *f, double *u,
                      const double *u_prev, const double *scale, int block_rows, bool zero_guess)
{
   long nblocks = ((long)A.nrows + block_rows - 1) / block_rows;
   k_hybrid_jgs<<<grid_for(cfg, nblocks), kBlock, 0, st>>>(A, f, u, u_prev, scale, block_rows, zero_guess ? 1 : 0);
   return 1;
}

int launch_diag_scale(cudaStream_t st, const DevCSR &A, double w, double *ws, double *dow)
{
   LaunchCfg cfg;
   k_diag_scale<<<grid_for(cfg, A.nrows), kBlock, 0, st>>>(A, w, ws, dow);
   return 1;
}

int launch_l1(cudaStream_t st, const DevCSR &A, double *l1, double *inv_l1)
{
   LaunchCfg cfg;
   k_l1<<<grid_for(cfg, A.nrows), kBlock, 0, st>>>(A, l1, inv_l1);
   return 1;
}

int launch_colscale(cudaStream_t st, int nnz, const int *ci, const double *va, const double *cs, double *out)
{
   LaunchCfg cfg;
   k_colscale<<<grid_for(cfg, nnz), kBlock, 0, st>>>(nnz, ci, va, cs, out);
   return 1;
}
