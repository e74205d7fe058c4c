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

// ---- CSR-stream: row blocks staged through shared memory by TMA bulk copies ---------------------
// A CTA owns every team_nctas-th row block.  One thread keeps AMGB_STREAM_STAGES blocks in flight:
// per block two cp.async.bulk copies (column indices, values; contiguous, 16-byte aligned segments of
// the CSR arrays) that complete on the stage's mbarrier -- no register staging, the HBM stream never
// stalls on the row structure.  Per block all 256 threads then (1) multiply the staged values by the
// gathered x in place, (2) sum each row's products with a sub-warp sized to the block's row count and
// apply the fused epilogue, whose row operands were loaded before the barrier wait.  A block made of
// ONE row longer than AMGB_STREAM_CAP is read straight from global memory by the whole CTA.
// `smem`: AMGB_STREAM_SMEM bytes, 16-byte aligned.  Must be called by all threads of a 256-thread CTA.
template <bool RO, bool SVAL>
__device__ __forceinline__ double stream_rows_team(const DevCSR &M, const double *__restrict__ x, double *y,
                                                   const SpmvEpilogue &e, int team_cta, int team_nctas,
                                                   unsigned char *smem, bool want_sumsq)
{
   constexpr int CAP = AMGB_STREAM_CAP, ST = AMGB_STREAM_STAGES;
   const int tid = threadIdx.x;
   const double *__restrict__ va = SVAL ? M.sval : M.va;
   int *scol = reinterpret_cast<int *>(smem);                                    // [ST][CAP]
   double *sval = reinterpret_cast<double *>(smem + (size_t)ST * CAP * 4);        // [ST][CAP]
   const uint32_t bar0 = smem_u32(smem + (size_t)ST * CAP * 12);
   const int nmine = team_cta < M.nblk ? (M.nblk - team_cta + team_nctas - 1) / team_nctas : 0;
   double sumsq = 0.0;
   uint64_t pol = 0;

   auto issue = [&](const int4 d, int stage) {       // thread 0 only
      const int q0 = d.z & ~3;
      if (d.w - q0 > CAP || d.w == d.z) return;       // long row / no entries: not staged
      const uint32_t n4 = (uint32_t)((d.w - q0 + 3) & ~3);
      const uint32_t bar = bar0 + 8u * stage;
      mbar_expect_tx(bar, n4 * 12u);
      tma_bulk_g2s(smem_u32(scol + stage * CAP), M.ci + q0, n4 * 4u, bar, pol);
      tma_bulk_g2s(smem_u32(sval + stage * CAP), va + q0, n4 * 8u, bar, pol);
   };

   __syncthreads();                                   // previous users of smem (and of the barriers) are done
   if (tid == 0) {
      pol = l2_evict_first_policy();
      for (int s = 0; s < ST; s++) mbar_init(bar0 + 8u * s, 1);
      mbar_init_fence();
      fence_proxy_async();
      for (int s = 0; s < ST && s < nmine; s++) issue(__ldg(M.blk + team_cta + s * team_nctas), s);
   }
   __syncthreads();
   uint32_t phase = 0;                                // bit s: parity to wait for on stage s
   int4 d = nmine > 0 ? __ldg(M.blk + team_cta) : make_int4(0, 0, 0, 0);
   for (int i = 0; i < nmine; i++) {
      const int b = team_cta + i * team_nctas;
      const int stage = i % ST;
      const int r0 = d.x, r1 = d.y, p0 = d.z, p1 = d.w, q0 = p0 & ~3, nr = r1 - r0;
      int4 dn = make_int4(0, 0, 0, 0), dp = dn;
      if (i + 1 < nmine) dn = __ldg(M.blk + b + team_nctas);
      if (tid == 0 && i + ST < nmine) dp = __ldg(M.blk + b + ST * team_nctas);
      if (p1 - q0 > CAP) {
         double acc = 0.0;
         for (int p = p0 + tid; p < p1; p += 256) acc += ld_stream(va + p) * ld_x<RO>(x + ld_stream(M.ci + p));
         acc = block_sum(acc);
         if (tid == 0) {
            const double v = epilogue_apply<RO>(e, r0, acc);
            y[r0] = v;
            if (want_sumsq) sumsq += v * v;
         }
         __syncthreads();
      } else {
         // lanes per row: the largest power of two (<= 32) such that all rows fit in one pass
         int lpr = 1;
         while (lpr < 32 && nr * (lpr << 1) <= 256) lpr <<= 1;
         const int lane = tid & (lpr - 1);
         const int rows_per_pass = 256 / lpr;
         // first pass: row pointers and epilogue operands loaded before the barrier wait
         const int row_a = r0 + tid / lpr;
         const bool ok_a = row_a < r1;
         int s_a = 0, t_a = 0;
         EpiOps o_a = {0.0, 0.0, 1.0};
         if (ok_a) {
            s_a = __ldg(M.rp + row_a) - q0;
            t_a = __ldg(M.rp + row_a + 1) - q0;
            if (lane == 0) o_a = epilogue_load<RO>(e, row_a);
         }
         if (p1 > p0) {
            mbar_wait(bar0 + 8u * stage, (phase >> stage) & 1u);
            phase ^= 1u << stage;
         }
         const int *cs = scol + stage * CAP;
         double *vs = sval + stage * CAP;
         // products in place.  Consecutive lanes take consecutive entries: neighbouring entries of a row
         // mostly point at neighbouring x, so one gather instruction touches few 128-byte lines (the
         // L1 wavefront count per gather, not HBM, is what limits a scattered mapping).
         const int first = p0 - q0, cnt = p1 - q0;
         double xv[AMGB_STREAM_CAP / 256];
#pragma unroll
         for (int k = 0; k < AMGB_STREAM_CAP / 256; k++) {
            const int q = tid + 256 * k;
            xv[k] = (q >= first && q < cnt) ? ld_x<RO>(x + cs[q]) : 0.0;
         }
#pragma unroll
         for (int k = 0; k < AMGB_STREAM_CAP / 256; k++) {
            const int q = tid + 256 * k;
            if (q < cnt) vs[q] = (q >= first) ? vs[q] * xv[k] : 0.0;
         }
         __syncthreads();
         {
            double acc = 0.0;
            for (int q = s_a + lane; q < t_a; q += lpr) acc += vs[q];
            for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_down_sync(AMGB_FULL, acc, o, 32);
            if (lane == 0 && ok_a) {
               const double v = epilogue_finish(e, o_a, acc);
               y[row_a] = v;
               if (want_sumsq) sumsq += v * v;
            }
         }
         for (int base = rows_per_pass; base < nr; base += rows_per_pass) {   // only when lpr == 1
            const int row = r0 + base + tid;
            if (row < r1) {
               const int s1 = __ldg(M.rp + row) - q0, t1 = __ldg(M.rp + row + 1) - q0;
               double acc = 0.0;
               for (int q = s1; q < t1; q++) acc += vs[q];
               const double v = epilogue_apply<RO>(e, row, acc);
               y[row] = v;
               if (want_sumsq) sumsq += v * v;
            }
         }
         __syncthreads();                              // the stage is free again
      }
      if (tid == 0 && i + ST < nmine) {
         fence_proxy_async();
         issue(dp, stage);
      }
      d = dn;
   }
   __syncthreads();
   if (tid == 0)
      for (int s = 0; s < ST; s++) mbar_inval(bar0 + 8u * s);   // the memory may be re-initialised by the next call
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
// (team_tid, team_size) = (team_cta * 256 + threadIdx.x, team_nctas * 256); smem: AMGB_STREAM_SMEM bytes
// of 16-byte aligned shared memory
template <bool RO, bool SVAL>
__device__ __forceinline__ double spmv_team(const DevCSR &M, const double *x, double *y, const SpmvEpilogue &e,
                                            int team_tid, int team_size, bool want_sumsq, unsigned char *smem)
{
   if (M.sell_slices > 0) return sell_rows_team<RO, SVAL>(M, x, y, e, team_tid, team_size, want_sumsq);
   if (M.nblk > 0) return stream_rows_team<RO, SVAL>(M, x, y, e, team_tid >> 8, team_size >> 8, smem, want_sumsq);
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
