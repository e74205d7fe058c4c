// spgemm.cu -- EXPERIMENTAL (compiled, not yet run on hardware; no default path calls it): smoothed-interpolant construction
// on the device, SURVEY.md 8f-1.
//
// Replaces SmoothTransfer (src/SMEM_Setup.cpp:1173-1254), whose two sparse products Pbar = G P and Rbar = P^T GT go through
// a serial Eigen triplet assembly + product (EigenMatMat, :1256-1339) and dominate the reference's Multadd setup time.
//   weighted Jacobi:  G_ii = 1 - w,          G_ij = -w a_ij / d_i;   GT_ii = 1 - w,          GT_ij = -w a_ij / d_j
//   L1:               G_ii = 1 - a_ii/l1_i,  G_ij = -a_ij / l1_i;    GT_ii = 1 - a_ii/l1_i,  GT_ij = -a_ij / l1_j
// Same boundary as the reference call: host CSR in (A_l diag-first, plain P_l), host CSR out, rows in the layout the
// reference gives every product (descending columns, then the entry with column == row swapped to the front, :1382-1423).
//
// Method: expand - sort - compress.  Every product term x_e * y_q becomes a (64-bit key = row * ncols + (ncols-1-col), value)
// pair in expansion order; a stable radix sort by key (thrust) groups the terms of one output entry, in order; a segmented
// reduction sums them; row pointers come from binary searches of the row boundaries in the sorted unique keys.  The
// transposed product needs no explicit transpose: the terms of Rbar are enumerated from the entries of P.  Deterministic
// pattern; values are sums of the same terms as the host restatement (amgh_smooth_transfer) in a different association.
#include "ctx.h"
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int kT = 256;
inline int blocks_for(long n) { return (int)std::max(1L, std::min((n + kT - 1) / kT, 1L << 20)); }

// s[r] = a_rr (Jacobi) or sum_j |a_rj| (L1)
__global__ void k_row_scale(int n, const int *rp, const double *va, int l1, double *s)
{
   for (int r = blockIdx.x * kT + threadIdx.x; r < n; r += gridDim.x * kT) {
      if (!l1) { s[r] = va[rp[r]]; continue; }
      double t = 0.0;
      for (int p = rp[r]; p < rp[r + 1]; p++) t += fabs(va[p]);
      s[r] = t;
   }
}

// values of G (by_col = 0: scaled by the ROW's s) or GT (by_col = 1: scaled by the COLUMN's s) on A's pattern; erow[e] = row of entry e
__global__ void k_smoother_values(int n, const int *rp, const int *ci, const double *va, const double *s, int l1, double w, int by_col,
                                  double *g, int *erow)
{
   for (int r = blockIdx.x * kT + threadIdx.x; r < n; r += gridDim.x * kT) {
      const int d = rp[r];
      g[d] = l1 ? 1.0 - va[d] / s[r] : 1.0 - w;
      erow[d] = r;
      for (int p = d + 1; p < rp[r + 1]; p++) {
         const double sc = s[by_col ? ci[p] : r];
         g[p] = l1 ? -va[p] / sc : -w * va[p] / sc;
         erow[p] = r;
      }
   }
}

__global__ void k_entry_rows(int n, const int *rp, int *erow)
{
   for (int r = blockIdx.x * kT + threadIdx.x; r < n; r += gridDim.x * kT)
      for (int p = rp[r]; p < rp[r + 1]; p++) erow[p] = r;
}

// number of product terms of X entry e: the length of the Y row it meets (transposed: Y row = X entry's ROW)
__global__ void k_term_counts(long nx, const int *xrow, const int *xci, int transposed, const int *yrp, long long *cnt)
{
   for (long e = blockIdx.x * (long)kT + threadIdx.x; e < nx; e += (long)gridDim.x * kT) {
      const int yr = transposed ? xrow[e] : xci[e];
      cnt[e] = yrp[yr + 1] - yrp[yr];
   }
}

__global__ void k_expand(long nx, const int *xrow, const int *xci, const double *xva, int transposed, const int *yrp, const int *yci,
                         const double *yva, const long long *off, long long ncols, long long *key, double *val)
{
   for (long e = blockIdx.x * (long)kT + threadIdx.x; e < nx; e += (long)gridDim.x * kT) {
      const int yr = transposed ? xrow[e] : xci[e];
      const long long orow = transposed ? xci[e] : xrow[e];
      const double xv = xva[e];
      long long o = off[e];
      for (int q = yrp[yr]; q < yrp[yr + 1]; q++, o++) {
         key[o] = orow * ncols + (ncols - 1 - yci[q]);      // ascending key = descending column inside a row
         val[o] = xv * yva[q];
      }
   }
}

// rp[i] = first unique key >= i * ncols
__global__ void k_row_pointers(int nrows, long long ncols, const long long *ukey, long long nu, int *rp)
{
   for (int i = blockIdx.x * kT + threadIdx.x; i <= nrows; i += gridDim.x * kT) {
      const long long target = (long long)i * ncols;
      long long lo = 0, hi = nu;
      while (lo < hi) {
         const long long mid = (lo + hi) >> 1;
         if (ukey[mid] < target) lo = mid + 1; else hi = mid;
      }
      rp[i] = (int)lo;
   }
}

__global__ void k_columns(long long nu, long long ncols, const long long *ukey, int *ci)
{
   for (long long j = blockIdx.x * (long long)kT + threadIdx.x; j < nu; j += (long long)gridDim.x * kT)
      ci[j] = (int)(ncols - 1 - ukey[j] % ncols);
}

// the reference's product layout: the entry with column == row goes to the front of its row (src/SMEM_Setup.cpp:1405-1418)
__global__ void k_diag_to_front(int nrows, const int *rp, int *ci, double *va)
{
   for (int r = blockIdx.x * kT + threadIdx.x; r < nrows; r += gridDim.x * kT) {
      const int s = rp[r];
      for (int p = s; p < rp[r + 1]; p++)
         if (ci[p] == r) {
            const int ct = ci[s]; ci[s] = ci[p]; ci[p] = ct;
            const double vt = va[s]; va[s] = va[p]; va[p] = vt;
            break;
         }
   }
}

struct DevBuf {
   std::vector<void *> p;
   ~DevBuf() { for (void *q : p) cudaFree(q); }
   template <class T> cudaError_t get(T **out, size_t n)
   {
      void *q = nullptr;
      cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
      if (e == cudaSuccess) p.push_back(q);
      *out = (T *)q;
      return e;
   }
};

// C = X * Y (transposed = 0) or X^T * Y (transposed = 1) on the device; X given with per-entry rows; result to malloc'ed host CSR
int esc_product(amgb_ctx *c, int out_rows, int out_cols, long nx, const int *xrow, const int *xci, const double *xva, int transposed,
                const int *yrp, const int *yci, const double *yva, amgb_host_csr *out)
{
   cudaStream_t st = c->stream;
   DevBuf buf;
   long long *cnt = nullptr, *key = nullptr, *ukey = nullptr;
   double *val = nullptr, *uval = nullptr;
   CUDA_OK(c, buf.get(&cnt, (size_t)nx + 1));
   k_term_counts<<<blocks_for(nx), kT, 0, st>>>(nx, xrow, xci, transposed, yrp, cnt);
   c->launches++;
   thrust::device_ptr<long long> dc(cnt);
   long long last_cnt = 0, last_off = 0;
   if (nx > 0) CUDA_OK(c, cudaMemcpyAsync(&last_cnt, cnt + nx - 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
   thrust::exclusive_scan(thrust::cuda::par.on(st), dc, dc + nx, dc);
   if (nx > 0) CUDA_OK(c, cudaMemcpyAsync(&last_off, cnt + nx - 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
   CUDA_OK(c, cudaStreamSynchronize(st));
   const long long terms = last_off + last_cnt;
   CUDA_OK(c, buf.get(&key, (size_t)terms));
   CUDA_OK(c, buf.get(&val, (size_t)terms));
   CUDA_OK(c, buf.get(&ukey, (size_t)terms));
   CUDA_OK(c, buf.get(&uval, (size_t)terms));
   k_expand<<<blocks_for(nx), kT, 0, st>>>(nx, xrow, xci, xva, transposed, yrp, yci, yva, cnt, (long long)out_cols, key, val);
   c->launches++;
   thrust::device_ptr<long long> dk(key), duk(ukey);
   thrust::device_ptr<double> dv(val), duv(uval);
   thrust::stable_sort_by_key(thrust::cuda::par.on(st), dk, dk + terms, dv);
   auto ends = thrust::reduce_by_key(thrust::cuda::par.on(st), dk, dk + terms, dv, duk, duv);
   const long long nu = ends.first - duk;
   if (nu > 2147483000LL) return amgb_fail(c, AMGB_EINVAL, "product has %lld entries: too many for int32 CSR", nu);
   int *rp = nullptr, *ci = nullptr;
   CUDA_OK(c, buf.get(&rp, (size_t)out_rows + 1));
   CUDA_OK(c, buf.get(&ci, (size_t)nu));
   k_row_pointers<<<blocks_for(out_rows + 1), kT, 0, st>>>(out_rows, (long long)out_cols, ukey, nu, rp);
   k_columns<<<blocks_for(nu), kT, 0, st>>>(nu, (long long)out_cols, ukey, ci);
   k_diag_to_front<<<blocks_for(out_rows), kT, 0, st>>>(out_rows, rp, ci, uval);
   c->launches += 3;
   out->nrows = out_rows; out->ncols = out_cols; out->nnz = (int)nu;
   out->row_ptr = (int *)malloc(sizeof(int) * ((size_t)out_rows + 1));
   out->col_idx = (int *)malloc(sizeof(int) * std::max<size_t>((size_t)nu, 1));
   out->values = (double *)malloc(sizeof(double) * std::max<size_t>((size_t)nu, 1));
   if (!out->row_ptr || !out->col_idx || !out->values) return amgb_fail(c, AMGB_ENOMEM, "host allocation of the product failed");
   CUDA_OK(c, cudaMemcpyAsync(out->row_ptr, rp, sizeof(int) * ((size_t)out_rows + 1), cudaMemcpyDeviceToHost, st));
   CUDA_OK(c, cudaMemcpyAsync(out->col_idx, ci, sizeof(int) * (size_t)nu, cudaMemcpyDeviceToHost, st));
   CUDA_OK(c, cudaMemcpyAsync(out->values, uval, sizeof(double) * (size_t)nu, cudaMemcpyDeviceToHost, st));
   CUDA_OK(c, cudaStreamSynchronize(st));
   CUDA_OK(c, cudaGetLastError());
   return AMGB_OK;
}

}  // namespace

extern "C" {

void amgb_host_csr_free(amgb_host_csr *m)
{
   if (!m) return;
   free(m->row_ptr); free(m->col_idx); free(m->values);
   memset(m, 0, sizeof(*m));
}

int amgb_smooth_transfer(amgb_ctx *c, int smooth_interp_type, double w, int n, const int *A_rp, const int *A_ci, const double *A_va,
                         int nc, const int *P_rp, const int *P_ci, const double *P_va, amgb_host_csr *Pbar, amgb_host_csr *Rbar)
{
   if (!c) return AMGB_EINVAL;
   if (n < 1 || nc < 1 || !A_rp || !A_ci || !A_va || !P_rp || !P_ci || !P_va || (!Pbar && !Rbar) || w == 0.0)
      return amgb_fail(c, AMGB_EINVAL, "bad arguments");
   if (smooth_interp_type != AMGB_SMOOTH_JACOBI && smooth_interp_type != AMGB_SMOOTH_L1_JACOBI)
      return amgb_fail(c, AMGB_EINVAL, "smooth_interp_type must be JACOBI or L1_JACOBI");
   for (int r = 0; r < n; r++)
      if (A_rp[r + 1] <= A_rp[r] || A_ci[A_rp[r]] != r) return amgb_fail(c, AMGB_EINVAL, "A must be diagonal-first (row %d)", r);
   CUDA_OK(c, cudaSetDevice(c->device));
   cudaStream_t st = c->stream;
   const int l1 = smooth_interp_type == AMGB_SMOOTH_L1_JACOBI;
   const long nnzA = A_rp[n], nnzP = P_rp[n];
   DevBuf buf;
   int *arp, *aci, *prp, *pci, *arow, *prow;
   double *ava, *pva, *s, *g;
   CUDA_OK(c, buf.get(&arp, (size_t)n + 1)); CUDA_OK(c, buf.get(&aci, (size_t)nnzA)); CUDA_OK(c, buf.get(&ava, (size_t)nnzA));
   CUDA_OK(c, buf.get(&prp, (size_t)n + 1)); CUDA_OK(c, buf.get(&pci, (size_t)nnzP)); CUDA_OK(c, buf.get(&pva, (size_t)nnzP));
   CUDA_OK(c, buf.get(&arow, (size_t)nnzA)); CUDA_OK(c, buf.get(&prow, (size_t)nnzP));
   CUDA_OK(c, buf.get(&s, (size_t)n)); CUDA_OK(c, buf.get(&g, (size_t)nnzA));
   CUDA_OK(c, cudaMemcpyAsync(arp, A_rp, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
   CUDA_OK(c, cudaMemcpyAsync(aci, A_ci, sizeof(int) * (size_t)nnzA, cudaMemcpyHostToDevice, st));
   CUDA_OK(c, cudaMemcpyAsync(ava, A_va, sizeof(double) * (size_t)nnzA, cudaMemcpyHostToDevice, st));
   CUDA_OK(c, cudaMemcpyAsync(prp, P_rp, sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
   CUDA_OK(c, cudaMemcpyAsync(pci, P_ci, sizeof(int) * (size_t)nnzP, cudaMemcpyHostToDevice, st));
   CUDA_OK(c, cudaMemcpyAsync(pva, P_va, sizeof(double) * (size_t)nnzP, cudaMemcpyHostToDevice, st));
   k_row_scale<<<blocks_for(n), kT, 0, st>>>(n, arp, ava, l1, s);
   c->launches++;
   int rc;
   if (Pbar) {
      // Pbar = G P: terms enumerated from the entries of G (A's pattern), meeting the rows of P
      k_smoother_values<<<blocks_for(n), kT, 0, st>>>(n, arp, aci, ava, s, l1, w, 0, g, arow);
      c->launches++;
      if ((rc = esc_product(c, n, nc, nnzA, arow, aci, g, 0, prp, pci, pva, Pbar))) return rc;
   }
   if (Rbar) {
      // Rbar = P^T GT: terms enumerated from the entries of P (k, c), meeting row k of GT; output row = c
      k_smoother_values<<<blocks_for(n), kT, 0, st>>>(n, arp, aci, ava, s, l1, w, 1, g, arow);
      k_entry_rows<<<blocks_for(n), kT, 0, st>>>(n, prp, prow);
      c->launches += 2;
      if ((rc = esc_product(c, nc, n, nnzP, prow, pci, pva, 1, arp, aci, g, Rbar))) return rc;
   }
   return AMGB_OK;
}

}  // extern "C"
