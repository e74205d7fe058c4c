// common.cuh -- shared device helpers for the sm_100a additive-AMG kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AMGB_WARP 32
#define AMGB_FULL 0xffffffffu

// Device-resident CSR matrix (row_ptr/col/val in HBM).  `sval` is an optional second value
// array with the same pattern: A*diag(w/d) (or A*diag(1/l1)), which turns the reference's
// three-pass symmetrised Jacobi (src/SMEM_Smooth.cpp:643-702) into ONE SpMV-shaped pass.
// `sell_*`: optional sliced-ELL (C = 32) copy used by the stencil-level kernels.
struct DevCSR {
   int nrows = 0, ncols = 0, nnz = 0;
   const int *rp = nullptr;
   const int *ci = nullptr;
   const double *va = nullptr;
   const double *sval = nullptr;
   // sliced ELL, slice height 32: slice s holds rows [32s, 32s+32); entry k of row r sits at
   // sell_off[s] + k*32 + (r & 31); padded entries have col = row's own index and val = 0.
   int sell_slices = 0;
   const int *sell_off = nullptr;    // [slices+1] offsets in units of entries
   const int *sell_ci = nullptr;
   const double *sell_va = nullptr;
   const double *sell_sval = nullptr;
   int lpr = 8;                      // lanes per row chosen for the CSR vector kernel
   // CSR-stream row blocks: CTA b owns rows [blk[b], blk[b+1]) whose entries (<= AMGB_STREAM_CAP,
   // counted from the 4-aligned start) are streamed with 128-bit loads into shared memory and then
   // reduced per row; a block made of ONE longer row is reduced by the whole CTA.
   int nblk = 0;
   const int *blk = nullptr;
};
#define AMGB_STREAM_CAP 2048

// y_i = gamma*c_i + rs_i * (beta*b_i + alpha * sum_j M_ij x_j)       (rs == nullptr -> 1)
// covers: MatVec (alpha=1), Residual (alpha=-1,beta=1,b=f), prolong-and-add (beta=1,b=y),
// general Jacobi sweep (rs=w/d, alpha=-1, beta=1, b=f, gamma=1, c=u_prev),
// symmetrised Jacobi from zero guess with M = A*diag(w/d) (rs=w/d, alpha=-1, beta=2, b=r).
struct SpmvEpilogue {
   double alpha, beta, gamma;
   const double *b;
   const double *c;
   const double *rs;
};

__device__ __forceinline__ double ld_stream(const double *p)
{
   // matrix values / indices are read exactly once: keep them out of L1
   double v;
   asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
   return v;
}
__device__ __forceinline__ int ld_stream(const int *p)
{
   int v;
   asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
   return v;
}
// 128-bit streaming loads (4 column indices / 2 values per instruction), 16-byte aligned
__device__ __forceinline__ int4 ld_stream4(const int *p)
{
   int4 v;
   asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
   return v;
}
__device__ __forceinline__ double2 ld_stream2(const double *p)
{
   double2 v;
   asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
   return v;
}
// coherent (L2) load for vectors that other CTA groups update concurrently inside the persistent
// asynchronous kernel
__device__ __forceinline__ double ld_cg(const double *p)
{
   double v;
   asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
   return v;
}
__device__ __forceinline__ void st_cg(double *p, double v)
{
   asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// fire-and-forget fp64 reduction into global memory (#pragma omp atomic, src/SMEM_Async_AMG.cpp:297)
__device__ __forceinline__ void red_add_f64(double *p, double v)
{
   asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <int W>
__device__ __forceinline__ double subwarp_sum(double v)
{
#pragma unroll
   for (int o = W / 2; o > 0; o >>= 1) v += __shfl_down_sync(AMGB_FULL, v, o, W);
   return v;
}

// RO: operands are read-only for the whole launch (stand-alone kernels); otherwise they may have
// been written by other CTAs earlier in the same persistent launch and are read from L2.
template <bool RO>
__device__ __forceinline__ double epilogue_apply(const SpmvEpilogue &e, int row, double ax)
{
   double t = e.alpha * ax;
   if (e.b) t += e.beta * (RO ? e.b[row] : ld_cg(e.b + row));
   if (e.rs) t *= __ldg(e.rs + row);
   if (e.c) t += e.gamma * (RO ? e.c[row] : ld_cg(e.c + row));
   return t;
}

// block-wide sum; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v)
{
   __shared__ double sm[32];
   v = subwarp_sum<32>(v);
   int w = threadIdx.x >> 5, l = threadIdx.x & 31;
   if (l == 0) sm[w] = v;
   __syncthreads();
   if (w == 0) {
      v = (l < (blockDim.x >> 5)) ? sm[l] : 0.0;
      v = subwarp_sum<32>(v);
   }
   __syncthreads();
   return v;
}
