// kernels.cuh -- device-side building blocks shared by the stand-alone kernels (kernels.cu) and
// the persistent asynchronous kernel (async.cu).  A "team" is the set of threads that cooperates
// on one operation: the whole grid for a stand-alone launch, one level's CTA group inside the
// persistent kernel (the reference's thread group of a level, src/SMEM_Setup.cpp:855-868).
#pragma once
#include "common.cuh"

// x loads: stand-alone kernels read x through the read-only path; inside the persistent kernel x
// was written by other CTAs of the team earlier in the same launch, so it is read with ld.cg (L2).
template <bool RO>
__device__ __forceinline__ double ld_x(const double *p)
{
   if (RO) return __ldg(p);
   return ld_cg(p);
}

// ---- CSR, LPR lanes per row (vector-per-row; LPR = 32 is warp-per-row) --------------------------
// reference loop: src/SMEM_MatVec.cpp:316-322 (+ the epilogues listed in common.cuh)
template <int LPR, bool RO, bool SVAL>
__device__ __forceinline__ double csr_rows_team(const DevCSR &M, const double *__restrict__ x,
                                                double *y, const SpmvEpilogue &e,
                                                int team_tid, int team_size, bool want_sumsq)
{
   const int lane = team_tid & (LPR - 1);
   constexpr int RPW = 32 / LPR;                       // rows per warp and iteration
   const int nsub = team_size / LPR;                   // rows per team iteration (multiple of RPW)
   const double *__restrict__ va = SVAL ? M.sval : M.va;
   double sumsq = 0.0;
   // the trip count is uniform per warp so that the full-mask shuffles below are legal
   for (int row0 = (team_tid >> 5) * RPW; row0 < M.nrows; row0 += nsub) {
      const int row = row0 + ((team_tid & 31) / LPR);
      const bool ok = row < M.nrows;
      const int s = ok ? __ldg(M.rp + row) : 0, t = ok ? __ldg(M.rp + row + 1) : 0;
      double acc = 0.0;
      for (int p = s + lane; p < t; p += LPR) acc += ld_stream(va + p) * ld_x<RO>(x + ld_stream(M.ci + p));
      acc = subwarp_sum<LPR>(acc);
      if (lane == 0 && ok) {
         double v = epilogue_apply<RO>(e, row, acc);
         y[row] = v;
         if (want_sumsq) sumsq += v * v;
      }
   }
   return sumsq;
}

// ---- sliced ELL (C = 32): one thread per row, one warp per slice, coalesced col/val streams -----
template <bool RO, bool SVAL>
__device__ __forceinline__ double sell_rows_team(const DevCSR &M, const double *__restrict__ x,
                                                 double *y, const SpmvEpilogue &e,
                                                 int team_tid, int team_size, bool want_sumsq)
{
   const int lane = team_tid & 31;
   const int nwarp = team_size >> 5;
   const double *__restrict__ va = SVAL ? M.sell_sval : M.sell_va;
   double sumsq = 0.0;
   for (int sl = team_tid >> 5; sl < M.sell_slices; sl += nwarp) {
      const int off = __ldg(M.sell_off + sl);
      const int width = (__ldg(M.sell_off + sl + 1) - off) >> 5;
      const int row = (sl << 5) + lane;
      const int *__restrict__ cp = M.sell_ci + off + lane;
      const double *__restrict__ vp = va + off + lane;
      double acc = 0.0;
#pragma unroll 4
      for (int k = 0; k < width; k++) acc += ld_stream(vp + (k << 5)) * ld_x<RO>(x + ld_stream(cp + (k << 5)));
      if (row < M.nrows) {
         double v = epilogue_apply<RO>(e, row, acc);
         y[row] = v;
         if (want_sumsq) sumsq += v * v;
      }
   }
   return sumsq;
}

// ---- CSR-stream: one CTA per row block --------------------------------------------------------
// Phase 1 streams the block's (col,val) pairs with 128-bit coalesced loads -- every lane busy and
// two independent 48-byte groups in flight per thread regardless of the row lengths -- multiplies
// by the gathered x and parks the products in shared memory.  Phase 2 sums each row's products
// with a sub-warp sized to the block's row count and applies the fused epilogue.  `sprod` holds
// AMGB_STREAM_CAP doubles.  Must be called by all threads of a 256-thread CTA.
template <bool RO, bool SVAL>
__device__ __forceinline__ double stream_block(const DevCSR &M, int b, const double *__restrict__ x, double *y,
                                               const SpmvEpilogue &e, double *sprod, bool want_sumsq)
{
   const int tid = threadIdx.x;
   const int r0 = __ldg(M.blk + b), r1 = __ldg(M.blk + b + 1);
   const int p0 = __ldg(M.rp + r0), p1 = __ldg(M.rp + r1);
   const int q0 = p0 & ~3;
   const double *__restrict__ va = SVAL ? M.sval : M.va;
   double sumsq = 0.0;
   if (p1 - q0 > AMGB_STREAM_CAP) {
      // a single long row: the whole CTA strides over it
      double acc = 0.0;
      for (int p = p0 + tid; p < p1; p += 256) acc += ld_stream(va + p) * ld_x<RO>(x + ld_stream(M.ci + p));
      acc = block_sum(acc);
      if (tid == 0) {
         const double v = epilogue_apply<RO>(e, r0, acc);
         y[r0] = v;
         if (want_sumsq) sumsq = v * v;
      }
      return sumsq;
   }
   const int ngroups = (p1 - q0 + 3) >> 2;            // <= 512: at most two groups per thread
   {
      const int g0 = tid, g1 = tid + 256;
      const bool h0 = g0 < ngroups, h1 = g1 < ngroups;
      int4 c0 = make_int4(0, 0, 0, 0), c1 = c0;
      double2 a0 = make_double2(0, 0), a1 = a0, b0 = a0, b1 = a0;
      if (h0) { const int p = q0 + 4 * g0; c0 = ld_stream4(M.ci + p); a0 = ld_stream2(va + p); a1 = ld_stream2(va + p + 2); }
      if (h1) { const int p = q0 + 4 * g1; c1 = ld_stream4(M.ci + p); b0 = ld_stream2(va + p); b1 = ld_stream2(va + p + 2); }
      if (h0) {
         const int p = q0 + 4 * g0;
         double2 o0, o1;
         o0.x = (p >= p0 && p < p1) ? a0.x * ld_x<RO>(x + c0.x) : 0.0;
         o0.y = (p + 1 >= p0 && p + 1 < p1) ? a0.y * ld_x<RO>(x + c0.y) : 0.0;
         o1.x = (p + 2 >= p0 && p + 2 < p1) ? a1.x * ld_x<RO>(x + c0.z) : 0.0;
         o1.y = (p + 3 >= p0 && p + 3 < p1) ? a1.y * ld_x<RO>(x + c0.w) : 0.0;
         *reinterpret_cast<double2 *>(sprod + 4 * g0) = o0;
         *reinterpret_cast<double2 *>(sprod + 4 * g0 + 2) = o1;
      }
      if (h1) {
         const int p = q0 + 4 * g1;
         double2 o0, o1;
         o0.x = (p < p1) ? b0.x * ld_x<RO>(x + c1.x) : 0.0;
         o0.y = (p + 1 < p1) ? b0.y * ld_x<RO>(x + c1.y) : 0.0;
         o1.x = (p + 2 < p1) ? b1.x * ld_x<RO>(x + c1.z) : 0.0;
         o1.y = (p + 3 < p1) ? b1.y * ld_x<RO>(x + c1.w) : 0.0;
         *reinterpret_cast<double2 *>(sprod + 4 * g1) = o0;
         *reinterpret_cast<double2 *>(sprod + 4 * g1 + 2) = o1;
      }
   }
   __syncthreads();
   const int nr = r1 - r0;
   // lanes per row: the largest power of two (<= 32) such that all rows fit in one pass
   int lpr = 1;
   while (lpr < 32 && nr * (lpr << 1) <= 256) lpr <<= 1;
   const int lane = tid & (lpr - 1);
   const int rows_per_pass = 256 / lpr;
   for (int base = 0; base < nr; base += rows_per_pass) {
      const int row = r0 + base + tid / lpr;
      const bool ok = row < r1;
      const int s = ok ? __ldg(M.rp + row) - q0 : 0, t = ok ? __ldg(M.rp + row + 1) - q0 : 0;
      double acc = 0.0;
      for (int i = s + lane; i < t; i += lpr) acc += sprod[i];
      for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_down_sync(AMGB_FULL, acc, o, 32);
      if (lane == 0 && ok) {
         const double v = epilogue_apply<RO>(e, row, acc);
         y[row] = v;
         if (want_sumsq) sumsq += v * v;
      }
   }
   return sumsq;
}

// all row blocks of M, dealt round-robin to the CTAs of a team; returns the thread's sum of y_i^2
template <bool RO, bool SVAL>
__device__ __forceinline__ double stream_rows_team(const DevCSR &M, const double *x, double *y, const SpmvEpilogue &e,
                                                   int team_cta, int team_nctas, double *sprod, bool want_sumsq)
{
   double sumsq = 0.0;
   for (int b = team_cta; b < M.nblk; b += team_nctas) {
      sumsq += stream_block<RO, SVAL>(M, b, x, y, e, sprod, want_sumsq);
      __syncthreads();                                 // sprod is reused by the next block
   }
   return sumsq;
}

// vector-per-row CSR with the lanes-per-row chosen at upload time
template <bool RO, bool SVAL>
__device__ __forceinline__ double csr_rows_dispatch(const DevCSR &M, const double *x, double *y, const SpmvEpilogue &e,
                                                    int team_tid, int team_size, bool want_sumsq)
{
   switch (M.lpr) {
      case 2: return csr_rows_team<2, RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
      case 4: return csr_rows_team<4, RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
      case 8: return csr_rows_team<8, RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
      case 16: return csr_rows_team<16, RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
      default: return csr_rows_team<32, RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
   }
}


// dispatch on storage + lanes per row
// (team_tid, team_size) = (team_cta * 256 + threadIdx.x, team_nctas * 256); sprod: AMGB_STREAM_CAP doubles
// of shared memory
template <bool RO, bool SVAL>
__device__ __forceinline__ double spmv_team(const DevCSR &M, const double *x, double *y, const SpmvEpilogue &e,
                                            int team_tid, int team_size, bool want_sumsq, double *sprod)
{
   if (M.sell_slices > 0) return sell_rows_team<RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
   if (M.nblk > 0) return stream_rows_team<RO, SVAL>(M, x, y, e, team_tid >> 8, team_size >> 8, sprod, want_sumsq);
   return csr_rows_dispatch<RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
}

// ---- hybrid Jacobi / Gauss-Seidel (src/SMEM_Smooth.cpp:533-586) ----------------------------------
// Gauss-Seidel inside a block of `B` consecutive rows (one thread walks one block in row order),
// Jacobi across blocks.  Zero-guess form (:549-562): out-of-block terms are skipped.  General form
// (:565-581): in-block columns read live u, others u_prev.  Divisor a_ii (weight forced to 1,
// :546) unless `scale` (= a_ii/omega, Parfor variant :253-263) is given.
template <bool RO>
__device__ __forceinline__ void hybrid_jgs_team(const DevCSR &A, const double *__restrict__ f,
                                                double *u, const double *__restrict__ u_prev,
                                                const double *__restrict__ scale, int B, bool zero_guess,
                                                int team_tid, int team_size)
{
   const int nblocks = (A.nrows + B - 1) / B;
   for (int blk = team_tid; blk < nblocks; blk += team_size) {
      const int ns = blk * B, ne = min(ns + B, A.nrows);
      if (zero_guess)
         for (int i = ns; i < ne; i++) u[i] = 0.0;
      for (int i = ns; i < ne; i++) {
         const int s = A.rp[i], t = A.rp[i + 1];
         const double d = A.va[s];
         if (d != 0.0) {
            double res = ld_x<RO>(f + i);
            for (int p = s; p < t; p++) {
               const int ii = A.ci[p];
               if (ii >= ns && ii < ne) res -= A.va[p] * u[ii];
               else if (!zero_guess) res -= A.va[p] * ld_x<RO>(u_prev + ii);
            }
            const double div = scale ? scale[i] : d;
            if (zero_guess) u[i] = res / div;
            else u[i] += res / div;
         }
      }
   }
}
