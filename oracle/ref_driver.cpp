// ref_driver.cpp -- TEST INFRASTRUCTURE.  Drives the reference's own, UNMODIFIED solve-phase
// object code (SMEM_Solve, SMEM_Sync_Add_Vcycle, SMEM_Sync_Parfor_{BPX,AFACx}cycle,
// SMEM_Async_Add_AMG, SMEM_Smooth*, SMEM_MatVec*, SEQ_*, Misc.cpp barriers/norms), compiled by
// oracle/build_ref.sh from /root/reference/src into oracle/_ref/libref_smem.so.
//
// What this file does is the part of SMEM_Setup the solve phase needs but that cannot be
// compiled here (it calls hypre / Eigen): it fills `AllData` from a hierarchy handed in as
// flat CSR arrays, following InitAlgebra's allocations (src/SMEM_Setup.cpp:280-419) and
// PartitionLevels / PartitionGrids' final assignment loops (src/SMEM_Setup.cpp:855-868,
// 940-1036).  The per-level thread counts come from the caller (hierarchy.balanced_threads
// restates src/SMEM_Setup.cpp:770-854).  It then calls the reference's InitSolve + SMEM_Solve.
//
// The per-iteration residual history is captured at full precision by routing SMEM_Solve.cpp's
// printf (only that TU, -Dprintf=ref_hook_printf) to a hook that reads all_data->output.
#include "ref_prelude.hpp"
#include "Misc.hpp"
#include "SMEM_Solve.hpp"
#include "SMEM_MatVec.hpp"
#include "SEQ_Smooth.hpp"
#include "SMEM_Sync_AMG.hpp"
#include "SMEM_ExtendedSystem.hpp"
#include "DMEM_Main.hpp"
#include "DMEM_Mult.hpp"
#include "DMEM_Misc.hpp"
#include "DMEM_Comm.hpp"
#include "DMEM_Add.hpp"
#include "DMEM_Smooth.hpp"
#include "SMEM_Cheby.hpp"
void AddCycle(DMEM_AllData *dmem_all_data);     // defined (non-static) in src/DMEM_Add.cpp:180, declared only inside that file
#include <cstdarg>

static AllData *g_all = nullptr;
static double *g_hist = nullptr;
static int g_hist_cap = 0;

extern "C" int ref_hook_printf(const char *fmt, ...)
{
   // SMEM_Solve.cpp:232-239 prints "%d\t%e\t\t%e\n" with (k, r_norm2/r0_norm2, rate)
   if (g_all && g_hist && strncmp(fmt, "%d\t%e", 5) == 0) {
      va_list ap;
      va_start(ap, fmt);
      int k = va_arg(ap, int);
      va_end(ap);
      if (k >= 0 && k < g_hist_cap) g_hist[k] = (k == 0) ? 1.0 : g_all->output.r_norm2 / g_all->output.r0_norm2;
   }
   return 0;
}

// ---- the few hypre entry points the solve phase calls -------------------------------------------
HYPRE_Int *hypre_LowerBound(HYPRE_Int *first, HYPRE_Int *last, HYPRE_Int value)
{
   return std::lower_bound(first, last, value);
}
// hypre csr_matrix.c: hypre_CSRMatrixGetLoadBalancedPartitionBoundary (published algorithm)
static HYPRE_Int lb_boundary(hypre_CSRMatrix *A, HYPRE_Int idx)
{
   HYPRE_Int nnz = A->num_nonzeros, n = A->num_rows, T = omp_get_num_threads();
   HYPRE_Int per = (nnz + T - 1) / T;
   if (idx <= 0) return 0;
   if (idx >= T) return n;
   return (HYPRE_Int)(hypre_LowerBound(A->i, A->i + n, per * idx) - A->i);
}
HYPRE_Int hypre_CSRMatrixGetLoadBalancedPartitionBegin(hypre_CSRMatrix *A) { return lb_boundary(A, omp_get_thread_num()); }
HYPRE_Int hypre_CSRMatrixGetLoadBalancedPartitionEnd(hypre_CSRMatrix *A) { return lb_boundary(A, omp_get_thread_num() + 1); }
// SMEM additive cycles call this on hypre's own F/U arrays, which the cycle never reads
// (SURVEY.md 5.9c): a no-op reproduces the reference result exactly.
// DMEM convention (driver flag functional_gauss_elim): U_array[level] = A_array[level]^{-1} F_array[level] by dense Gaussian
// elimination with partial pivoting (hypre_GaussElimSetup / Solve, relax types 9 / 99; un-vendored, restated).
HYPRE_Int hypre_GaussElimSolve(hypre_ParAMGData *amg, HYPRE_Int level, HYPRE_Int)
{
   if (!amg || !amg->functional_gauss_elim) return 0;
   hypre_CSRMatrix *A = amg->A_array[level]->diag;
   const int n = A->num_rows;
   std::vector<double> M((size_t)n * (n + 1), 0.0);
   const double *f = amg->F_array[level]->local_vector->data;
   double *u = amg->U_array[level]->local_vector->data;
   for (int i = 0; i < n; i++) {
      for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) M[(size_t)i * (n + 1) + A->j[jj]] += A->data[jj];
      M[(size_t)i * (n + 1) + n] = f[i];
   }
   for (int k = 0; k < n; k++) {
      int piv = k;
      for (int i = k + 1; i < n; i++) if (fabs(M[(size_t)i * (n + 1) + k]) > fabs(M[(size_t)piv * (n + 1) + k])) piv = i;
      if (piv != k) for (int j = 0; j <= n; j++) std::swap(M[(size_t)k * (n + 1) + j], M[(size_t)piv * (n + 1) + j]);
      for (int i = k + 1; i < n; i++) {
         const double fct = M[(size_t)i * (n + 1) + k] / M[(size_t)k * (n + 1) + k];
         if (fct != 0.0) for (int j = k; j <= n; j++) M[(size_t)i * (n + 1) + j] -= fct * M[(size_t)k * (n + 1) + j];
      }
   }
   for (int i = n - 1; i >= 0; i--) {
      double t = M[(size_t)i * (n + 1) + n];
      for (int j = i + 1; j < n; j++) t -= M[(size_t)i * (n + 1) + j] * u[j];
      u[i] = t / M[(size_t)i * (n + 1) + i];
   }
   return 0;
}
HYPRE_Int HYPRE_BoomerAMGSetPrintLevel(HYPRE_Solver, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_BoomerAMGSetMaxIter(HYPRE_Solver, HYPRE_Int) { return 0; }
HYPRE_Int hypre_ParVectorSetConstantValues(hypre_ParVector *v, HYPRE_Complex value)
{
   if (v && v->local_vector && v->local_vector->data)
      for (int i = 0; i < v->local_vector->size; i++) v->local_vector->data[i] = value;
   return 0;
}
// single-rank hypre vector / matrix operations named by the DMEM translation units (dmem_stub.h)
HYPRE_Int vecop_machine = HYPRE_MEMORY_HOST;            // src/DMEM_Main.cpp:10
HYPRE_Int hypre_ParCSRMatrixMatvecT(HYPRE_Complex alpha, hypre_ParCSRMatrix *A, hypre_ParVector *x, HYPRE_Complex beta, hypre_ParVector *y)
{
   hypre_CSRMatrix *M = A->diag;
   const double *xd = x->local_vector->data;
   double *yd = y->local_vector->data;
   for (int j = 0; j < M->num_cols; j++) yd[j] *= beta;
   for (int i = 0; i < M->num_rows; i++)
      for (int jj = M->i[i]; jj < M->i[i + 1]; jj++) yd[M->j[jj]] += alpha * M->data[jj] * xd[i];
   return 0;
}
HYPRE_Int hypre_BoomerAMGRelax(hypre_ParCSRMatrix *, hypre_ParVector *, HYPRE_Int *, HYPRE_Int, HYPRE_Int, HYPRE_Real, HYPRE_Real, HYPRE_Real *,
                               hypre_ParVector *, hypre_ParVector *, hypre_ParVector *) { abort(); }
HYPRE_Real hypre_SeqVectorInnerProd(hypre_Vector *x, hypre_Vector *y)
{
   double s = 0.0;
   for (int i = 0; i < x->size; i++) s += x->data[i] * y->data[i];
   return s;
}
HYPRE_Real hypre_ParVectorInnerProd(hypre_ParVector *x, hypre_ParVector *y) { return hypre_SeqVectorInnerProd(x->local_vector, y->local_vector); }
HYPRE_Int hypre_ParVectorScale(HYPRE_Complex alpha, hypre_ParVector *y)
{
   for (int i = 0; i < y->local_vector->size; i++) y->local_vector->data[i] *= alpha;
   return 0;
}
HYPRE_Int hypre_ParVectorAxpy(HYPRE_Complex alpha, hypre_ParVector *x, hypre_ParVector *y)
{
   for (int i = 0; i < y->local_vector->size; i++) y->local_vector->data[i] += alpha * x->local_vector->data[i];
   return 0;
}
HYPRE_Int hypre_CSRMatrixMatvec(HYPRE_Complex alpha, hypre_CSRMatrix *A, hypre_Vector *x, HYPRE_Complex beta, hypre_Vector *y)
{
   for (int i = 0; i < A->num_rows; i++) {
      double t = 0.0;
      for (int jj = A->i[i]; jj < A->i[i + 1]; jj++) t += A->data[jj] * x->data[A->j[jj]];
      y->data[i] = alpha * t + beta * y->data[i];
   }
   return 0;
}
// DMEM_Comm's message engine (src/DMEM_Comm.cpp; needs hypre's seq_mv.h and real MPI): one rank has no peer, the driver
// never reaches it
// SendRecv with NO peer in the list (one rank): the loop over comm_data->procs never runs and the function returns its initial
// return_flag = 0 (src/DMEM_Comm.cpp:92,95,347); with a peer it would need the real engine
int SendRecv(DMEM_AllData *, DMEM_CommData *comm_data, HYPRE_Real *, int) { if (!comm_data->procs.empty()) abort(); return 0; }
void CompleteRecv(DMEM_AllData *, DMEM_CommData *, HYPRE_Real *, int) { abort(); }
void CompleteInFlight(DMEM_AllData *, DMEM_CommData *) { abort(); }
void CheckInFlight(DMEM_AllData *, DMEM_CommData *, int) { abort(); }
void DMEM_ResetAllCommData(DMEM_AllData *) { abort(); }
hypre_ParVector *hypre_ParVectorCreate(MPI_Comm, HYPRE_BigInt, HYPRE_BigInt *) { abort(); }
HYPRE_Int hypre_ParVectorInitialize(hypre_ParVector *) { abort(); }
HYPRE_Int hypre_ParVectorDestroy(hypre_ParVector *) { abort(); }
HYPRE_Int hypre_ParVectorSetPartitioningOwner(hypre_ParVector *, HYPRE_Int) { abort(); }
// hypre ParCSR matvecs as far as SMEM_ExtendedSystem.cpp's EXPLICIT_EXTENDED_SYSTEM_BPX branch uses them (one rank: the
// diag block is the whole matrix): y = alpha A x + beta b  /  y = alpha A x + beta y  /  y = x
HYPRE_Int hypre_ParCSRMatrixMatvecOutOfPlace(HYPRE_Complex alpha, hypre_ParCSRMatrix *A, hypre_ParVector *x, HYPRE_Complex beta,
                                             hypre_ParVector *b, hypre_ParVector *y)
{
   hypre_CSRMatrix *M = A->diag;
   const double *xd = x->local_vector->data, *bd = b->local_vector->data;
   double *yd = y->local_vector->data;
   for (int i = 0; i < M->num_rows; i++) {
      double t = 0.0;
      for (int jj = M->i[i]; jj < M->i[i + 1]; jj++) t += M->data[jj] * xd[M->j[jj]];
      yd[i] = alpha * t + beta * bd[i];
   }
   return 0;
}
HYPRE_Int hypre_ParCSRMatrixMatvec(HYPRE_Complex alpha, hypre_ParCSRMatrix *A, hypre_ParVector *x, HYPRE_Complex beta, hypre_ParVector *y)
{
   return hypre_ParCSRMatrixMatvecOutOfPlace(alpha, A, x, beta, y, y);
}
HYPRE_Int hypre_ParVectorCopy(hypre_ParVector *x, hypre_ParVector *y)
{
   memcpy(y->local_vector->data, x->local_vector->data, sizeof(double) * x->local_vector->size);
   return 0;
}

// HYPRE_IJMatrix as far as ReadBinary_fread_HypreParCSR (src/Misc.cpp:800-915) uses it: rows are set one at a time
// and the assembled object is a ParCSR matrix whose (only) diag block is CSR.  As in hypre's IJ assembly the
// diagonal entry is moved to the front of its row; the other entries keep the order they were set in.
struct StubIJ {
   int n = 0;
   std::vector<std::vector<int>> cols;
   std::vector<std::vector<double>> vals;
   hypre_CSRMatrix csr;
   hypre_ParCSRMatrix par;
};
HYPRE_Int HYPRE_IJMatrixCreate(MPI_Comm, HYPRE_BigInt ilower, HYPRE_BigInt iupper, HYPRE_BigInt, HYPRE_BigInt, HYPRE_IJMatrix *m)
{
   StubIJ *ij = new StubIJ();
   ij->n = iupper - ilower + 1;
   ij->cols.resize(ij->n); ij->vals.resize(ij->n);
   *m = ij;
   return 0;
}
HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix, HYPRE_Int) { return 0; }
HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix) { return 0; }
HYPRE_Int HYPRE_IJMatrixSetValues(HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows, const HYPRE_BigInt *cols,
                                  const HYPRE_Complex *values)
{
   StubIJ *ij = (StubIJ *)m;
   int pos = 0;
   for (int r = 0; r < nrows; r++)
      for (int k = 0; k < ncols[r]; k++, pos++) { ij->cols[rows[r]].push_back(cols[pos]); ij->vals[rows[r]].push_back(values[pos]); }
   return 0;
}
HYPRE_Int HYPRE_IJMatrixAssemble(HYPRE_IJMatrix m)
{
   StubIJ *ij = (StubIJ *)m;
   size_t nnz = 0;
   for (auto &c : ij->cols) nnz += c.size();
   hypre_CSRMatrix &A = ij->csr;
   A.num_rows = A.num_cols = ij->n; A.num_nonzeros = (int)nnz; A.rownnz = nullptr; A.num_rownnz = ij->n;
   A.i = (int *)malloc(sizeof(int) * ((size_t)ij->n + 1));
   A.j = (int *)malloc(sizeof(int) * std::max<size_t>(nnz, 1));
   A.data = (double *)malloc(sizeof(double) * std::max<size_t>(nnz, 1));
   A.i[0] = 0;
   for (int r = 0; r < ij->n; r++) {
      int d = A.i[r];
      const int len = (int)ij->cols[r].size();
      for (int k = 0; k < len; k++) { A.j[d + k] = ij->cols[r][k]; A.data[d + k] = ij->vals[r][k]; }
      for (int k = 0; k < len; k++)
         if (A.j[d + k] == r) {
            const int jt = A.j[d + k]; const double vt = A.data[d + k];
            for (int q = k; q > 0; q--) { A.j[d + q] = A.j[d + q - 1]; A.data[d + q] = A.data[d + q - 1]; }
            A.j[d] = jt; A.data[d] = vt;
            break;
         }
      A.i[r + 1] = d + len;
   }
   ij->par.diag = &ij->csr; ij->par.global_num_rows = ij->n;
   return 0;
}
HYPRE_Int HYPRE_IJMatrixGetObject(HYPRE_IJMatrix m, void **object) { *object = &((StubIJ *)m)->par; return 0; }

// referenced by SMEM_Solve.cpp (async with one thread) but defined in a TU that is compiled too
// (SEQ_AMG.cpp); nothing else is missing.

struct RefCSR { int nrows, ncols, nnz; int *i; int *j; double *data; };

// hypre csr_matop.c: hypre_CSRMatrixTranspose (published algorithm: count the columns, prefix sum, scatter rows in order), as
// far as src/SMEM_Setup.cpp:253,268 uses it (explicit R = P^T for the solvers with plain transfers).  A row of AT lists the
// rows of A in ascending order.
HYPRE_Int hypre_CSRMatrixTranspose(hypre_CSRMatrix *A, hypre_CSRMatrix **AT, HYPRE_Int data)
{
   const int m = A->num_rows, n = A->num_cols, nnz = A->i[m];
   hypre_CSRMatrix *T = (hypre_CSRMatrix *)calloc(1, sizeof(hypre_CSRMatrix));
   T->num_rows = n; T->num_cols = m; T->num_nonzeros = nnz; T->num_rownnz = n;
   T->i = (HYPRE_Int *)calloc((size_t)n + 1, sizeof(HYPRE_Int));
   T->j = (HYPRE_Int *)calloc((size_t)nnz, sizeof(HYPRE_Int));
   T->data = data ? (HYPRE_Real *)calloc((size_t)nnz, sizeof(HYPRE_Real)) : nullptr;
   for (int p = 0; p < nnz; p++) T->i[A->j[p] + 1]++;
   for (int c = 0; c < n; c++) T->i[c + 1] += T->i[c];
   std::vector<int> next(T->i, T->i + n);
   for (int r = 0; r < m; r++)
      for (int p = A->i[r]; p < A->i[r + 1]; p++) {
         const int q = next[A->j[p]]++;
         T->j[q] = r;
         if (data) T->data[q] = A->data[p];
      }
   *AT = T;
   return 0;
}

// hypre's stencil generators (un-vendored), as called by src/BuildHypreMatrix.cpp: these stand-ins RECORD the coefficients the
// reference computed -- value[0] centre; GenerateLaplacian: [1] x, [2] y, [3] z neighbours; GenerateDifConv: [1] x-1, [2] y-1,
// [3] z-1, [4] x+1, [5] y+1, [6] z+1 (hypre par_laplace.c / par_difconv.c) -- and build nothing.
static double g_gen_values[8];
static int g_gen_count = 0;
static int g_gen_grid[3] = {0, 0, 0};      // the processor grid (P, Q, R) the reference chose (src/BuildHypreMatrix.cpp:36-76)
int amg_ref_num_procs = 1;
HYPRE_ParCSRMatrix GenerateLaplacian(MPI_Comm, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R, HYPRE_Int, HYPRE_Int, HYPRE_Int,
                                     HYPRE_Real *value)
{ memcpy(g_gen_values, value, sizeof(double) * 4); g_gen_count = 4; g_gen_grid[0] = P; g_gen_grid[1] = Q; g_gen_grid[2] = R; return nullptr; }
HYPRE_ParCSRMatrix GenerateLaplacian27pt(MPI_Comm, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int,
                                         HYPRE_Int, HYPRE_Real *value) { memcpy(g_gen_values, value, sizeof(double) * 2); g_gen_count = 2; return nullptr; }
HYPRE_ParCSRMatrix GenerateDifConv(MPI_Comm, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int,
                                   HYPRE_Real *value) { memcpy(g_gen_values, value, sizeof(double) * 7); g_gen_count = 7; return nullptr; }
HYPRE_ParCSRMatrix GenerateVarDifConv(MPI_Comm, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int, HYPRE_Int,
                                      HYPRE_Real, HYPRE_ParVector *) { abort(); }
void BuildHypreMatrix(AllData *all_data, HYPRE_ParCSRMatrix *A_ptr, HYPRE_ParVector *rhs_ptr, MPI_Comm comm, HYPRE_Int nx, HYPRE_Int ny,
                      HYPRE_Int nz, HYPRE_Real cx, HYPRE_Real cy, HYPRE_Real cz, HYPRE_Real ax, HYPRE_Real ay, HYPRE_Real az, HYPRE_Real eps,
                      int atype);
// The stencil coefficients src/BuildHypreMatrix.cpp:100-289 hands to hypre for test_problem = LAPLACE_3D7PT (1), LAPLACE_3D27PT (2)
// or DIFCONV_3D7PT (7); returns how many values were recorded.
extern "C" int ref_stencil_values(int test_problem, int nx, int ny, int nz, double cx, double cy, double cz, double ax, double ay, double az,
                                  int atype, double *out)
{
   AllData *ad = new AllData();
   memset((void *)&ad->input, 0, sizeof(ad->input));
   ad->input.test_problem = test_problem;
   HYPRE_ParCSRMatrix A = nullptr;
   HYPRE_ParVector rhs = nullptr;
   g_gen_count = 0;
   BuildHypreMatrix(ad, &A, &rhs, 0, nx, ny, nz, cx, cy, cz, ax, ay, az, 0.0, atype);
   memcpy(out, g_gen_values, sizeof(double) * g_gen_count);
   delete ad;
   return g_gen_count;
}
// The processor grid (P, Q, R) the reference's search (src/BuildHypreMatrix.cpp:36-76, the same code as src/DMEM_BuildMatrix.cpp:169-240)
// picks for `num_procs` ranks on an nx x ny x nz grid.
extern "C" void ref_processor_grid(int num_procs, int nx, int ny, int nz, int *pqr)
{
   double tmp[8];
   amg_ref_num_procs = num_procs;
   ref_stencil_values(LAPLACE_3D7PT, nx, ny, nz, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0, tmp);
   amg_ref_num_procs = 1;
   pqr[0] = g_gen_grid[0]; pqr[1] = g_gen_grid[1]; pqr[2] = g_gen_grid[2];
}

// named by SMEM_BuildMatrix (src/SMEM_Setup.cpp:1600-1660), which the driver never calls; src/Laplacian.cpp does not compile here (its 3-D
// half needs hypre's GenerateLaplacian)
void Laplacian_2D_5pt(HYPRE_IJMatrix *, int) { abort(); }
// hypre csr_matrix.c / csr_matop.c as far as BuildExtendedMatrix uses them.  hypre_CSRMatrixMultiply (published algorithm, serial
// path): row by row with a marker array; an entry of C appears where its column is FIRST touched (entries a_ik in stored order,
// then b_kj in stored order) and later touches accumulate into it; when A's row count equals B's column count the diagonal
// entry is placed first.  No sorting.
hypre_CSRMatrix *hypre_CSRMatrixCreate(HYPRE_Int num_rows, HYPRE_Int num_cols, HYPRE_Int num_nonzeros)
{
   hypre_CSRMatrix *M = (hypre_CSRMatrix *)calloc(1, sizeof(hypre_CSRMatrix));
   M->num_rows = num_rows; M->num_cols = num_cols; M->num_nonzeros = num_nonzeros; M->num_rownnz = num_rows;
   return M;
}
hypre_CSRMatrix *hypre_CSRMatrixMultiply(hypre_CSRMatrix *A, hypre_CSRMatrix *B)
{
   const int m = A->num_rows, n = B->num_cols;
   if (A->num_cols != B->num_rows) abort();
   const bool allsquare = (m == n);
   std::vector<int> ci, cp((size_t)m + 1, 0), marker((size_t)n, -1);
   std::vector<double> cv;
   for (int ic = 0; ic < m; ic++) {
      const int row_start = (int)ci.size();
      if (allsquare) { marker[ic] = (int)ci.size(); ci.push_back(ic); cv.push_back(0.0); }
      for (int ia = A->i[ic]; ia < A->i[ic + 1]; ia++) {
         const int ja = A->j[ia]; const double a = A->data[ia];
         for (int ib = B->i[ja]; ib < B->i[ja + 1]; ib++) {
            const int jb = B->j[ib]; const double b = B->data[ib];
            if (marker[jb] < row_start) { marker[jb] = (int)ci.size(); ci.push_back(jb); cv.push_back(a * b); }
            else cv[marker[jb]] += a * b;
         }
      }
      cp[ic + 1] = (int)ci.size();
   }
   hypre_CSRMatrix *Cm = hypre_CSRMatrixCreate(m, n, (int)ci.size());
   Cm->i = (HYPRE_Int *)malloc(sizeof(HYPRE_Int) * ((size_t)m + 1));
   Cm->j = (HYPRE_Int *)malloc(sizeof(HYPRE_Int) * std::max<size_t>(ci.size(), 1));
   Cm->data = (HYPRE_Real *)malloc(sizeof(HYPRE_Real) * std::max<size_t>(ci.size(), 1));
   memcpy(Cm->i, cp.data(), sizeof(int) * ((size_t)m + 1));
   if (!ci.empty()) { memcpy(Cm->j, ci.data(), sizeof(int) * ci.size()); memcpy(Cm->data, cv.data(), sizeof(double) * cv.size()); }
   return Cm;
}

void BuildExtendedMatrix(AllData *all_data, hypre_CSRMatrix **A_array, hypre_CSRMatrix **P_array, hypre_CSRMatrix **R_array,
                         hypre_CSRMatrix **B_ptr);
// src/SMEM_Setup.cpp (compiled unmodified against oracle/ref_shim/eigen_stub: Eigen is un-vendored)
void SmoothTransfer(AllData *all_data, hypre_CSRMatrix *P, hypre_CSRMatrix *R, int level);
void ComputeWork(AllData *all_data);
void PartitionLevels(AllData *all_data);
void PartitionGrids(AllData *all_data);

struct RefHandle {
   AllData all;
   std::vector<hypre_CSRMatrix> A, P, R;
   hypre_ParAMGData amg;
   std::vector<hypre_ParVector> pv;
   std::vector<hypre_Vector> lv;
   std::vector<hypre_ParVector *> Farr, Uarr;
   std::vector<hypre_ParCSRMatrix> parA;
   std::vector<hypre_ParCSRMatrix *> Aarr;
   std::vector<std::vector<double>> store;
   double *vec(size_t n) { store.emplace_back(n, 0.0); return store.back().data(); }
};

static void fill(hypre_CSRMatrix *m, const RefCSR &s)
{
   m->i = s.i; m->j = s.j; m->data = s.data;
   m->num_rows = s.nrows; m->num_cols = s.ncols; m->num_nonzeros = s.nnz;
   m->rownnz = nullptr; m->num_rownnz = s.nrows;
}

// ONE_LEVEL thread partition for the next ref_create whatever the solver: reaches SMEM_Sync_Parfor_AFACx_Vcycle
// (src/SMEM_Sync_AMG.cpp:296-406), which SMEM_Main.cpp:641-649 never selects for AFACX (it always takes ALL_LEVELS)
static int g_force_one_level = 0;
extern "C" void ref_force_one_level(int v) { g_force_one_level = v; }
// -read_type for the next ref_create (READ_SOL is the reference's default, src/SMEM_Main.cpp:93)
static int g_read_type = READ_SOL;
extern "C" void ref_set_read_type(int v) { g_read_type = v; }

extern "C" {

// threads_per_level[L]: BALANCED_THREADS result (caller).  thread_part_type is chosen as
// SMEM_Main.cpp:641-649 does (ONE_LEVEL for BPX/MULT, ALL_LEVELS otherwise).
void *ref_create(int L, const RefCSR *A, const RefCSR *P, const RefCSR *R, double *const *l1,
                 int solver, int smoother, double smooth_weight, int num_pre, int num_post,
                 int fine_sweeps, int coarse_sweeps, int num_threads, const int *threads_per_level,
                 const double *f)
{
   RefHandle *H = new RefHandle();
   AllData *ad = &H->all;
   memset((void *)&ad->input, 0, sizeof(ad->input));
   memset((void *)&ad->output, 0, sizeof(ad->output));
   memset((void *)&ad->matrix, 0, sizeof(ad->matrix));
   memset((void *)&ad->cheby, 0, sizeof(ad->cheby));
   // defaults of src/SMEM_Main.cpp:64-104
   ad->input.tol = 1e-9;
   ad->input.async_flag = (solver == ASYNC_MULTADD || solver == ASYNC_AFACX);
   ad->input.async_type = FULL_ASYNC;
   ad->input.check_resnorm_flag = 1;
   ad->input.converge_test_type = LOCAL;
   ad->input.res_compute_type = LOCAL;
   ad->input.thread_part_distr_type = BALANCED_THREADS;
   ad->input.num_pre_smooth_sweeps = num_pre;
   ad->input.num_post_smooth_sweeps = num_post;
   ad->input.num_fine_smooth_sweeps = fine_sweeps;
   ad->input.num_coarse_smooth_sweeps = coarse_sweeps;
   ad->input.num_threads = num_threads;
   ad->input.smooth_weight = smooth_weight;
   ad->input.smoother = smoother;
   ad->input.smooth_interp_type = JACOBI;
   ad->input.solver = solver;
   ad->input.read_type = g_read_type;
   ad->input.delay_type = DELAY_NONE;
   ad->input.construct_R_flag = 1;
   ad->input.print_reshist_flag = 1;
   ad->input.format_output_flag = 1;
   ad->input.thread_part_type = (g_force_one_level || solver == MULT || solver == BPX || solver == PAR_BPX) ? ONE_LEVEL : ALL_LEVELS;
   ad->cheby.mu = 1.0; ad->cheby.delta = 1.0;

   ad->grid.num_levels = L;
   ad->grid.n = (int *)malloc(sizeof(int) * L);
   H->A.resize(L); H->P.resize(L); H->R.resize(L);
   ad->matrix.A = (hypre_CSRMatrix **)malloc(sizeof(void *) * L);
   ad->matrix.P = (hypre_CSRMatrix **)malloc(sizeof(void *) * L);
   ad->matrix.R = (hypre_CSRMatrix **)malloc(sizeof(void *) * L);
   ad->matrix.L1_row_norm = (double **)malloc(sizeof(double *) * L);
   ad->matrix.A_diag = (double **)malloc(sizeof(double *) * L);
   for (int l = 0; l < L; l++) {
      fill(&H->A[l], A[l]);
      ad->matrix.A[l] = &H->A[l];
      ad->grid.n[l] = A[l].nrows;
      ad->matrix.L1_row_norm[l] = l1[l];
      // src/SMEM_Setup.cpp:233-238
      ad->matrix.A_diag[l] = H->vec(A[l].nrows);
      for (int i = 0; i < A[l].nrows; i++) ad->matrix.A_diag[l][i] = A[l].data[A[l].i[i]] / smooth_weight;
      if (l < L - 1) {
         fill(&H->P[l], P[l]); fill(&H->R[l], R[l]);
         ad->matrix.P[l] = &H->P[l]; ad->matrix.R[l] = &H->R[l];
      }
   }
   // hypre solver object: only the arrays SMEM_Solve.cpp:21-28 dereferences
   H->lv.resize(3); H->pv.resize(3);
   for (int k = 0; k < 3; k++) { H->lv[k].data = H->vec(A[0].nrows); H->lv[k].size = A[0].nrows; H->pv[k].local_vector = &H->lv[k]; }
   H->Farr.assign(L, &H->pv[0]); H->Uarr.assign(L, &H->pv[1]);
   H->parA.resize(L); H->Aarr.resize(L);
   for (int l = 0; l < L; l++) { H->parA[l].diag = &H->A[l]; H->parA[l].global_num_rows = A[l].nrows; H->Aarr[l] = &H->parA[l]; }
   H->amg.A_array = H->Aarr.data(); H->amg.P_array = nullptr; H->amg.R_array = nullptr;
   H->amg.F_array = H->Farr.data(); H->amg.U_array = H->Uarr.data();
   H->amg.Vtemp = &H->pv[2]; H->amg.Ztemp = &H->pv[2]; H->amg.l1_norms = (HYPRE_Real **)l1;
   ad->hypre.solver = (HYPRE_Solver)&H->amg;
   if (solver == PAR_BPX) {
      // src/SMEM_Setup.cpp:426-462: the level vectors concatenated, and the smoother's scale array over all levels
      ad->grid.disp = (int *)malloc((L + 1) * sizeof(int));
      ad->grid.N = 0; ad->grid.disp[0] = 0;
      for (int l = 0; l < L; l++) { ad->grid.disp[l + 1] = ad->grid.disp[l] + A[l].nrows; ad->grid.N += A[l].nrows; }
      ad->vector.xx = H->vec(ad->grid.N);
      ad->vector.rr = H->vec(ad->grid.N);
      double *ext = H->vec(ad->grid.N);
      const bool l1s = smoother == L1_JACOBI || smoother == L1_HYBRID_JACOBI_GAUSS_SEIDEL;
      for (int l = 0, k = 0; l < L; l++)
         for (int i = 0; i < A[l].nrows; i++, k++) ext[k] = l1s ? l1[l][i] : A[l].data[A[l].i[i]] / smooth_weight;
      if (l1s) ad->matrix.L1_row_norm_ext = ext; else ad->matrix.A_diag_ext = ext;
   }

   // vectors (src/SMEM_Setup.cpp:280-419)
   VectorData *v = &ad->vector;
   HYPRE_Real ***arrs[] = {&v->f, &v->u, &v->u_prev, &v->u_fine, &v->u_fine_prev, &v->u_coarse, &v->u_coarse_prev,
                           &v->y, &v->r, &v->r_fine, &v->r_coarse, &v->e, &v->z, &v->u_smooth};
   for (auto a : arrs) {
      *a = (HYPRE_Real **)calloc(L, sizeof(HYPRE_Real *));
      for (int l = 0; l < L; l++) (*a)[l] = H->vec(A[l].nrows);
   }
   v->y_expand = (HYPRE_Real **)calloc(L, sizeof(HYPRE_Real *));
   v->i.resize(L, vector<int>(0));
   memcpy(v->f[0], f, sizeof(double) * A[0].nrows);
   if (ad->input.thread_part_type == ALL_LEVELS) {
      ad->level_vector = (VectorData *)calloc(L, sizeof(VectorData));
      for (int level = 0; level < L; level++) {
         VectorData *lv = &ad->level_vector[level];
         HYPRE_Real ***la[] = {&lv->f, &lv->u, &lv->u_prev, &lv->u_coarse, &lv->u_coarse_prev, &lv->u_fine, &lv->u_fine_prev,
                               &lv->y, &lv->r, &lv->r_coarse, &lv->r_fine, &lv->e, &lv->z, &lv->z1, &lv->z2};
         for (auto a : la) {
            *a = (HYPRE_Real **)calloc(L, sizeof(HYPRE_Real *));
            // every inner level for the implicit extended system (InitVectors, src/Misc.cpp:571-577), else level + 2
            const int inner_end = solver == IMPLICIT_EXTENDED_SYSTEM_BPX ? L : level + 2;
            for (int inner = 0; inner < inner_end && inner < L; inner++) (*a)[inner] = H->vec(A[inner].nrows);
         }
      }
   }
   // src/SMEM_Setup.cpp:103-134
   ad->barrier.local_sense = (int *)calloc(num_threads, sizeof(int));
   ad->grid.global_smooth_flags = (int *)calloc(num_threads, sizeof(int));
   int **gi[] = {&ad->grid.zero_flags, &ad->grid.num_smooth_wait, &ad->grid.finest_num_res_compute,
                 &ad->grid.local_num_res_compute, &ad->grid.local_num_correct, &ad->grid.local_cycle_num_correct,
                 &ad->grid.last_read_correct, &ad->grid.last_read_cycle_correct};
   for (auto g : gi) *g = (int *)calloc(L, sizeof(int));
   ad->grid.mean_grid_wait = (double *)calloc(L, sizeof(double));
   ad->grid.max_grid_wait = (double *)calloc(L, sizeof(double));
   ad->grid.min_grid_wait = (double *)calloc(L, sizeof(double));
   double **od[] = {&ad->output.smooth_wtime, &ad->output.residual_wtime, &ad->output.restrict_wtime,
                    &ad->output.prolong_wtime, &ad->output.A_matvec_wtime, &ad->output.vec_wtime, &ad->output.innerprod_wtime};
   for (auto o : od) *o = (double *)calloc(num_threads, sizeof(double));
   ad->output.smooth_sweeps = (int *)calloc(num_threads, sizeof(int));

   // PartitionLevels, final assignment (src/SMEM_Setup.cpp:855-868 / 870-880)
   ThreadData *th = &ad->thread;
   th->thread_levels.resize(num_threads, vector<int>(0));
   th->level_threads.resize(L, vector<int>(0));
   th->barrier_flags = (int **)malloc(L * sizeof(int *));
   th->barrier_root = (int *)malloc(L * sizeof(int));
   th->global_barrier_flags = (int *)calloc(num_threads, sizeof(int));
   th->loc_sum = (double *)malloc(num_threads * sizeof(double));
   th->converge_flag = 0;
   for (int l = 0; l < L; l++) th->barrier_flags[l] = (int *)calloc(num_threads, sizeof(int));
   if (ad->input.thread_part_type == ALL_LEVELS) {
      int t = 0;
      for (int k = 0; k < L; k++) {
         th->barrier_root[k] = t;
         for (int tt = 0; tt < threads_per_level[k]; tt++) {
            th->thread_levels[t].push_back(k);
            th->level_threads[k].push_back(t);
            th->barrier_flags[k][t] = 0;
            if (t < num_threads - 1) t++;
         }
      }
   } else {
      for (int l = 0; l < L; l++) {
         for (int t = 0; t < num_threads; t++) { th->thread_levels[t].push_back(l); th->level_threads[l].push_back(t); }
         th->barrier_root[l] = 0;
      }
   }
   // PartitionGrids (src/SMEM_Setup.cpp:895-1036)
   int ***pa[] = {&th->A_ns, &th->A_ne, &th->R_ns, &th->R_ne, &th->P_ns, &th->P_ne, &th->row_ns, &th->row_ne};
   for (auto p : pa) {
      *p = (int **)malloc(L * sizeof(int *));
      for (int l = 0; l < L; l++) (*p)[l] = (int *)calloc(num_threads, sizeof(int));
   }
   auto bound = [](hypre_CSRMatrix *M, int per, int nt, int t, int *ns, int *ne) {
      int n = M->num_rows;
      *ns = (t == 0) ? 0 : (int)(hypre_LowerBound(M->i, M->i + n, per * t) - M->i);
      *ne = (t == nt - 1) ? n : (int)(hypre_LowerBound(M->i, M->i + n, per * (t + 1)) - M->i);
   };
   if (ad->input.thread_part_type == ALL_LEVELS) {
      for (int level = 0; level < L; level++) {
         int nlt = (int)th->level_threads[level].size();
         if (nlt == 0) continue;
         for (int inner = 0; inner < L; inner++)
            for (int i = 0; i < nlt; i++) {
               int t = th->level_threads[level][i];
               int st = t - th->level_threads[level][0];
               hypre_CSRMatrix *M = ad->matrix.A[inner];
               bound(M, (M->num_nonzeros + nlt - 1) / nlt, nlt, st, &th->A_ns[inner][t], &th->A_ne[inner][t]);
               if (inner < L - 1) {
                  M = ad->matrix.P[inner];
                  bound(M, (M->num_nonzeros + nlt - 1) / nlt, nlt, st, &th->P_ns[inner][t], &th->P_ne[inner][t]);
                  M = ad->matrix.R[inner];
                  bound(M, (M->num_nonzeros + nlt - 1) / nlt, nlt, st, &th->R_ns[inner][t], &th->R_ne[inner][t]);
               }
            }
         // row_ns / row_ne: rows dealt round-robin into contiguous parts (src/SMEM_Setup.cpp:982-1005)
         for (int inner = 0; inner < L; inner++) {
            const int n = ad->grid.n[inner];
            std::vector<int> parts(nlt, 0);
            for (int cnt = 0; cnt < n;)
               for (int i = 0; i < nlt && cnt < n; i++) { parts[i]++; cnt++; }
            int disp = 0;
            for (int i = 0; i < nlt; i++) {
               const int t = th->level_threads[level][i];
               th->row_ns[inner][t] = disp;
               disp += parts[i];
               th->row_ne[inner][t] = disp;
            }
         }
      }
   } else {
      for (int l = 0; l < L; l++) {
         int n = ad->grid.n[l];
         for (int t = 0; t < num_threads; t++) {
            int size = n / num_threads, rest = n - size * num_threads;
            if (t < rest) { th->A_ns[l][t] = t * size + t; th->A_ne[l][t] = (t + 1) * size + t + 1; }
            else { th->A_ns[l][t] = t * size + rest; th->A_ne[l][t] = (t + 1) * size + rest; }
         }
      }
   }
   return H;
}

// Runs InitSolve + SMEM_Solve of the reference.  hist[0..num_cycles] receives the relative
// residual after every cycle (sync); returns cycles done.  corrections[L] = local_num_correct.
int ref_solve(void *h, int num_cycles, double tol, int async_type, int cheby_flag, double mu, double delta,
              int precond_flag, double *u_out, double *hist, int *corrections, double *solve_seconds,
              double *final_relres)
{
   RefHandle *H = (RefHandle *)h;
   AllData *ad = &H->all;
   ad->input.num_cycles = num_cycles;
   ad->input.tol = tol;
   ad->input.async_type = async_type;
   ad->input.cheby_flag = cheby_flag;
   ad->input.precond_flag = precond_flag;
   ad->cheby.mu = mu; ad->cheby.delta = delta;
   omp_set_num_threads(ad->input.num_threads);
   InitSolve(ad);
   g_all = ad; g_hist = hist; g_hist_cap = num_cycles + 1;
   if (hist) for (int k = 0; k <= num_cycles; k++) hist[k] = -1.0;
   SMEM_Solve(ad);
   g_all = nullptr; g_hist = nullptr;
   int done = ad->output.num_cycles;
   if (hist && !ad->input.async_flag) { hist[0] = 1.0; hist[done] = ad->output.r_norm2 / ad->output.r0_norm2; }
   if (u_out) memcpy(u_out, ad->vector.u[0], sizeof(double) * ad->grid.n[0]);
   if (corrections) for (int l = 0; l < ad->grid.num_levels; l++) corrections[l] = ad->grid.local_num_correct[l];
   if (solve_seconds) *solve_seconds = ad->output.solve_wtime;
   if (final_relres) *final_relres = ad->output.r_norm2 / ad->output.r0_norm2;
   return done;
}

// Deterministic (race-free) run of the reference's grouped additive cycle.  SMEM_Solve's own
// loop (src/SMEM_Solve.cpp:128-240) starts the residual of iteration k without waiting for the
// other level groups to finish adding their corrections of iteration k (there is no barrier
// between SMEM_Sync_Add_Vcycle's "u += e" at src/SMEM_Sync_AMG.cpp:611-616 and
// SMEM_Sync_Residual at src/SMEM_Solve.cpp:192), and different groups add into the same u[i]
// without atomics.  This loop is SMEM_Solve's with ONE added "omp barrier"
// after the cycle and async_type = SEMI_ASYNC (the reference's own lock around the update);
// with one thread per level it is exactly the sequential specification.  The cycle, residual
// and smoothers executed are the reference's object code.
int ref_solve_sync_det(void *h, int num_cycles, double tol, double *u_out, double *hist, double *solve_seconds)
{
   RefHandle *H = (RefHandle *)h;
   AllData *ad = &H->all;
   ad->input.num_cycles = num_cycles;
   ad->input.tol = tol;
   ad->input.async_type = SEMI_ASYNC;
   ad->input.cheby_flag = 0;
   ad->input.precond_flag = 0;
   omp_set_num_threads(ad->input.num_threads);
   InitSolve(ad);
   HYPRE_Real *r = ad->vector.r[0];
   const int n0 = ad->grid.n[0];
#pragma omp parallel
   {
      SMEM_Sync_Residual(ad, ad->matrix.A[0], ad->vector.f[0], ad->vector.u[0], ad->vector.y[0], r);
   }
   ad->output.r0_norm2 = Parfor_Norm2(r, n0);
   hist[0] = 1.0;
   omp_init_lock(&ad->thread.lock);
   double r_inner_prod = 0;
   int done = 0;
   const double t_start = omp_get_wtime();      // the reference times exactly this loop (src/SMEM_Solve.cpp:107,245)
#pragma omp parallel
   {
      int tid = omp_get_thread_num();
      for (int k = 1; k <= num_cycles; k++) {
         SMEM_Sync_Add_Vcycle(ad);
#pragma omp barrier
         if (tid == 0) r_inner_prod = 0;
         SMEM_Sync_Residual(ad, ad->matrix.A[0], ad->vector.f[0], ad->vector.u[0], ad->vector.y[0], r);
#pragma omp for reduction(+ : r_inner_prod)
         for (int i = 0; i < n0; i++) r_inner_prod += r[i] * r[i];
         double r_norm2 = sqrt(r_inner_prod);
         if (tid == 0) { done = k; hist[k] = r_norm2 / ad->output.r0_norm2; }
#pragma omp barrier
         if (r_norm2 / ad->output.r0_norm2 < tol) break;
      }
   }
   if (solve_seconds) *solve_seconds = omp_get_wtime() - t_start;
   omp_destroy_lock(&ad->thread.lock);
   if (u_out) memcpy(u_out, ad->vector.u[0], sizeof(double) * n0);
   return done;
}

#ifdef REF_WITH_B200
// The binding of INTEGRATION.md (integration/SMEM_B200.hpp, compiled here against the reference's own Main.hpp) driven by the
// reference-side structs: InitSolve, then SMEM_Solve_B200 instead of SMEM_Solve.  hist: the library's residual history.
}   // extern "C"
#ifndef MPI_VERSION
static int MPI_Finalize() { return 0; }        // (the reference gets MPI through hypre's headers; this build is single-process)
#endif
#include "SMEM_B200.hpp"
extern "C" {
int ref_solve_b200(void *h, int num_cycles, double tol, int async_type, int res_compute_type, int read_type, double *u_out, double *hist,
                   int *corrections, double *final_relres)
{
   RefHandle *H = (RefHandle *)h;
   AllData *ad = &H->all;
   ad->input.num_cycles = num_cycles;
   ad->input.tol = tol;
   ad->input.async_type = async_type;
   ad->input.res_compute_type = res_compute_type;
   ad->input.read_type = read_type;
   ad->input.cheby_flag = 0;
   ad->input.precond_flag = 0;
   ad->input.print_reshist_flag = 0;
   InitSolve(ad);
   SMEM_B200_Upload(ad);
   SMEM_Solve_B200(ad);
   const int done = ad->output.num_cycles;
   if (hist) for (int k = 0; k <= done && k < (int)b200_hist.size(); k++) hist[k] = b200_hist[k];
   if (u_out) memcpy(u_out, ad->vector.u[0], sizeof(double) * ad->grid.n[0]);
   if (corrections) for (int l = 0; l < ad->grid.num_levels; l++) corrections[l] = ad->grid.local_num_correct[l];
   if (final_relres) *final_relres = ad->output.r_norm2 / ad->output.r0_norm2;
   amgb_destroy(b200);
   b200 = NULL;
   return done;
}

// The DMEM binding of INTEGRATION.md (integration/DMEM_B200.hpp, compiled here against the reference's own DMEM_Main.hpp) on ONE
// rank: every level is handed over whole (a one-rank "partition": nothing distributed), DMEM_AllData carries the options the
// way DMEM_Main's argument parser leaves them, DMEM_Add_B200 runs instead of DMEM_Add's loop.  R: the restriction operators
// themselves (rows = coarse points), as the library takes them.
}   // extern "C"
#include "DMEM_B200.hpp"
extern "C" {
int ref_dmem_solve_b200(int L, const RefCSR *A, const RefCSR *P, const RefCSR *R, double smooth_weight, int symmetrised, const double *b,
                        int num_cycles, double tol, int async_flag, double *x_out, double *hist, int *corrections, double *relres)
{
   DMEM_AllData *dm = new DMEM_AllData();
   dm->grid.num_levels = L;
   dm->input.solver = async_flag ? ASYNC_MULTADD : MULTADD;
   dm->input.smoother = JACOBI;
   dm->input.smooth_weight = smooth_weight;
   dm->input.simple_jacobi_flag = symmetrised ? -1 : 0;
   dm->input.async_flag = async_flag;
   dm->input.num_cycles = num_cycles;
   dm->input.tol = tol;
   dm->input.accel_type = 0;
   dm->cheby.mu = 1.0; dm->cheby.delta = 1.0;
   std::vector<B200Layout> lay(L);
   std::vector<B200Csr> bA(L), bP(L), bR(L);
   std::vector<int> owned(L);
   for (int l = 0; l < L; l++) {
      owned[l] = A[l].nrows;
      lay[l] = B200Layout{A[l].nrows, 0, A[l].nrows, 0, 0, 0, 0, 0, &owned[l]};
      bA[l] = B200Csr{A[l].nrows, A[l].ncols, A[l].nnz, A[l].i, A[l].j, A[l].data};
      if (l < L - 1) {
         bP[l] = B200Csr{P[l].nrows, P[l].ncols, P[l].nnz, P[l].i, P[l].j, P[l].data};
         bR[l] = B200Csr{R[l].nrows, R[l].ncols, R[l].nnz, R[l].i, R[l].j, R[l].data};
      }
   }
   amgb_ctx *ctx = DMEM_B200_Upload(dm, 0, L, lay.data(), bA.data(), bP.data(), bR.data());
   const double rr = DMEM_Add_B200(dm, ctx, b, x_out, hist, corrections);
   if (relres) *relres = rr;
   const int done = dm->iter.cycle;
   amgb_destroy(ctx);
   delete dm;
   return done;
}
#endif

// SMEM_ExtendedSystemSolve (src/SMEM_ExtendedSystem.cpp:9-836) for IMPLICIT_EXTENDED_SYSTEM_BPX, the way SMEM_Main runs it
// (InitSolve, then the solver: src/SMEM_Main.cpp:724-729).  The handle must have been created with solver = 16 and plain
// P / R = P^T.  Returns local_num_correct of thread 0 (the final loc_iters).
int ref_solve_iebpx(void *h, int num_cycles, double tol, double mu, double delta, double *u_out, double *ext_relres, double *relres)
{
   RefHandle *H = (RefHandle *)h;
   AllData *ad = &H->all;
   if (ad->input.solver != IMPLICIT_EXTENDED_SYSTEM_BPX) return -1;
   ad->input.num_cycles = num_cycles;
   ad->input.tol = tol;
   ad->input.async_flag = 0;
   ad->input.omp_parfor_flag = 0;
   ad->cheby.mu = mu; ad->cheby.delta = delta;
   omp_set_num_threads(ad->input.num_threads);
   InitSolve(ad);
   SMEM_ExtendedSystemSolve(ad);
   if (u_out) memcpy(u_out, ad->vector.u[0], sizeof(double) * ad->grid.n[0]);
   if (ext_relres) *ext_relres = ad->output.r_norm2_ext_sys / ad->output.r0_norm2_ext_sys;
   if (relres) *relres = ad->output.r_norm2 / ad->output.r0_norm2;
   return ad->grid.local_num_correct[0];
}

// SMEM_ExtendedSystemSolve for EXPLICIT_EXTENDED_SYSTEM_BPX (`-solver eebpx`), synchronous, on an assembled extended matrix
// AA (BuildExtendedMatrix is part of SMEM_Setup.cpp, which needs hypre: the caller assembles AA).  Fills exactly what that
// branch reads: matrix.AA, vector.xx/bb/rr/zz (src/SMEM_Setup.cpp:505-541), grid.disp, thread.AA_NS/NE (:543-556), and
// hypre's A / P / U / F arrays + Vtemp for the initial fine residual and the final prolongation sum (:736-775).
// x0 = 0 (InitVectors).  Returns loc_iters; x_out = fine-level solution, xx_out = the extended iterate.
int ref_solve_eebpx(int L, const RefCSR *A, const RefCSR *P, const RefCSR *AAin, const int *disp, const double *bb, int num_threads,
                    int num_cycles, double tol, double mu, double delta, double *x_out, double *xx_out, double *ext_relres, double *relres)
{
   AllData all;
   AllData *ad = &all;
   memset((void *)&ad->input, 0, sizeof(ad->input));
   memset((void *)&ad->output, 0, sizeof(ad->output));
   memset((void *)&ad->matrix, 0, sizeof(ad->matrix));
   memset((void *)&ad->cheby, 0, sizeof(ad->cheby));
   ad->input.solver = EXPLICIT_EXTENDED_SYSTEM_BPX;
   ad->input.num_threads = num_threads;
   ad->input.num_cycles = num_cycles;
   ad->input.tol = tol;
   ad->input.check_resnorm_flag = 1;
   ad->input.async_flag = 0;
   ad->input.omp_parfor_flag = 0;
   ad->input.delay_type = DELAY_NONE;
   ad->cheby.mu = mu; ad->cheby.delta = delta;
   ad->grid.num_levels = L;
   std::vector<int> n(L), dsp(disp, disp + L + 1), lnc(num_threads, 0);
   for (int l = 0; l < L; l++) n[l] = A[l].nrows;
   ad->grid.n = n.data(); ad->grid.disp = dsp.data(); ad->grid.local_num_correct = lnc.data();
   const int N = AAin->nrows;
   hypre_CSRMatrix AA; fill(&AA, *AAin);
   ad->matrix.AA = &AA;
   std::vector<double> xx(N, 0.0), rr(N, 0.0), zz(N, 0.0), b(bb, bb + N);
   ad->vector.xx = xx.data(); ad->vector.bb = b.data(); ad->vector.rr = rr.data(); ad->vector.zz = zz.data();
   std::vector<int> ns(num_threads), ne(num_threads);
   const int per = (AA.num_nonzeros + num_threads - 1) / num_threads;
   for (int t = 0; t < num_threads; t++) {
      ns[t] = (t == 0) ? 0 : (int)(hypre_LowerBound(AA.i, AA.i + N, per * t) - AA.i);
      ne[t] = (t == num_threads - 1) ? N : (int)(hypre_LowerBound(AA.i, AA.i + N, per * (t + 1)) - AA.i);
   }
   ad->thread.AA_NS = ns.data(); ad->thread.AA_NE = ne.data();
   std::vector<double> aw(num_threads, 0.0), vw(num_threads, 0.0), iw(num_threads, 0.0), w4(4 * (size_t)num_threads, 0.0);
   ad->output.A_matvec_wtime = aw.data(); ad->output.vec_wtime = vw.data(); ad->output.innerprod_wtime = iw.data();
   ad->output.smooth_wtime = w4.data(); ad->output.residual_wtime = w4.data() + num_threads;
   ad->output.restrict_wtime = w4.data() + 2 * num_threads; ad->output.prolong_wtime = w4.data() + 3 * num_threads;
   // hypre side: per-level A, P as ParCSR with one diag block; U / F vectors per level; Vtemp
   std::vector<hypre_CSRMatrix> hA(L), hP(L);
   std::vector<hypre_ParCSRMatrix> pA(L), pP(L);
   std::vector<hypre_ParCSRMatrix *> Aarr(L), Parr(L);
   std::vector<std::vector<double>> ud(L), fd(L);
   std::vector<hypre_Vector> uv(L), fv(L);
   std::vector<hypre_ParVector> up(L), fp(L);
   std::vector<hypre_ParVector *> Uarr(L), Farr(L);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); pA[l].diag = &hA[l]; pA[l].global_num_rows = n[l]; Aarr[l] = &pA[l];
      if (l < L - 1) { fill(&hP[l], P[l]); pP[l].diag = &hP[l]; pP[l].global_num_rows = n[l]; Parr[l] = &pP[l]; }
      ud[l].assign(n[l], 0.0); fd[l].assign(n[l], 0.0);
      uv[l].data = ud[l].data(); uv[l].size = n[l]; up[l].local_vector = &uv[l]; Uarr[l] = &up[l];
      fv[l].data = fd[l].data(); fv[l].size = n[l]; fp[l].local_vector = &fv[l]; Farr[l] = &fp[l];
   }
   memcpy(fd[0].data(), bb, sizeof(double) * n[0]);          // F_array[0] = f (bb's first block)
   std::vector<double> vd(n[0], 0.0);
   hypre_Vector vv; vv.data = vd.data(); vv.size = n[0];
   hypre_ParVector vp; vp.local_vector = &vv;
   hypre_ParAMGData amg;
   memset(&amg, 0, sizeof(amg));
   amg.A_array = Aarr.data(); amg.P_array = Parr.data(); amg.R_array = nullptr;
   amg.U_array = Uarr.data(); amg.F_array = Farr.data(); amg.Vtemp = &vp; amg.Ztemp = &vp;
   ad->hypre.solver = (HYPRE_Solver)&amg;
   omp_set_num_threads(num_threads);
   SMEM_ExtendedSystemSolve(ad);
   if (x_out) memcpy(x_out, ud[0].data(), sizeof(double) * n[0]);
   if (xx_out) memcpy(xx_out, xx.data(), sizeof(double) * N);
   if (ext_relres) *ext_relres = ad->output.r_norm2_ext_sys / ad->output.r0_norm2_ext_sys;
   if (relres) *relres = ad->output.r_norm2 / ad->output.r0_norm2;
   return lnc[0];
}

// ---- DMEM: the synchronous additive solve on all ranks, ONE rank -------------------------------------------------------
// DMEM_SyncAdd (src/DMEM_Mult.cpp:263-319) = DMEM_SyncAddCycle (:322-450) + residual + norm per cycle, the reference's
// object code.  hypre's solver object is filled the way DMEM_Setup leaves it for this path: A_array, P_array (the smoothed
// interpolants), R_array applied TRANSPOSED (hypre_ParCSRMatrixMatvecT; Rt[l] is n_l x n_{l+1}), add_rlx_wt = omega,
// simple = -1 for the symmetrised smoother (src/DMEM_Setup.cpp:465-483), GridRelaxType[1] = 0 (Jacobi), and a direct solve
// on the coarsest level.  x0 = 0, r = b.  hist[k] = ||b - A x_k|| / ||b||.  Returns the cycles done.
int ref_dmem_sync_add(int L, const RefCSR *A, const RefCSR *P, const RefCSR *Rt, double smooth_weight, int symmetrised,
                      const double *b, int num_cycles, double tol, double *x_out, double *hist)
{
   DMEM_AllData *dm = new DMEM_AllData();
   std::vector<hypre_CSRMatrix> hA(L), hP(L), hR(L);
   std::vector<hypre_ParCSRMatrix> pA(L), pP(L), pR(L);
   std::vector<hypre_ParCSRMatrix *> Aarr(L), Parr(L), Rarr(L);
   std::vector<std::vector<double>> ud(L), fd(L);
   std::vector<hypre_Vector> uv(L), fv(L);
   std::vector<hypre_ParVector> up(L), fp(L);
   std::vector<hypre_ParVector *> Uarr(L), Farr(L);
   std::vector<double *> l1(L, nullptr);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); memset(&pA[l], 0, sizeof(pA[l])); pA[l].diag = &hA[l]; pA[l].global_num_rows = A[l].nrows; Aarr[l] = &pA[l];
      if (l < L - 1) {
         fill(&hP[l], P[l]); memset(&pP[l], 0, sizeof(pP[l])); pP[l].diag = &hP[l]; pP[l].global_num_rows = P[l].nrows; Parr[l] = &pP[l];
         fill(&hR[l], Rt[l]); memset(&pR[l], 0, sizeof(pR[l])); pR[l].diag = &hR[l]; pR[l].global_num_rows = Rt[l].nrows; Rarr[l] = &pR[l];
      }
      ud[l].assign(A[l].nrows, 0.0); fd[l].assign(A[l].nrows, 0.0);
      uv[l].data = ud[l].data(); uv[l].size = A[l].nrows; up[l].local_vector = &uv[l]; Uarr[l] = &up[l];
      fv[l].data = fd[l].data(); fv[l].size = A[l].nrows; fp[l].local_vector = &fv[l]; Farr[l] = &fp[l];
   }
   const int n0 = A[0].nrows;
   std::vector<double> vt(n0, 0.0), xv(n0, 0.0), rv(b, b + n0), bv(b, b + n0), ev(n0, 0.0);
   hypre_Vector hv[5];
   hypre_ParVector pv[5];
   double *ptrs[5] = {vt.data(), xv.data(), rv.data(), bv.data(), ev.data()};
   for (int k = 0; k < 5; k++) { hv[k].data = ptrs[k]; hv[k].size = n0; pv[k].local_vector = &hv[k]; }
   hypre_ParAMGData amg;
   memset(&amg, 0, sizeof(amg));
   int relax_type[4] = {0, 0, 0, 0};
   amg.A_array = Aarr.data(); amg.P_array = Parr.data(); amg.R_array = Rarr.data(); amg.P_array_afacj = Parr.data();
   amg.F_array = Farr.data(); amg.U_array = Uarr.data(); amg.Vtemp = &pv[0]; amg.Ztemp = &pv[0];
   amg.l1_norms = l1.data(); amg.num_levels = L; amg.grid_relax_type = relax_type; amg.add_rlx_wt = smooth_weight;
   amg.simple = symmetrised ? -1 : 0; amg.functional_gauss_elim = 1;
   dm->hypre.solver = (HYPRE_Solver)&amg;
   dm->input.solver = SYNC_MULTADD;
   dm->input.tol = tol;
   dm->input.num_cycles = 1;                               // one cycle per call: the history is collected here
   dm->vector_fine.x = &pv[1]; dm->vector_fine.r = &pv[2]; dm->vector_fine.b = &pv[3]; dm->vector_fine.e = &pv[4];
   double r0 = 0.0;
   for (int i = 0; i < n0; i++) r0 += b[i] * b[i];
   r0 = sqrt(r0);
   dm->output.r0_norm2 = r0;
   if (hist) hist[0] = 1.0;
   int done = 0;
   for (int k = 1; k <= num_cycles; k++) {
      DMEM_SyncAdd(dm);                                    // cycle + residual into vector_fine.r (+ its norm, discarded)
      double rn = 0.0;
      for (int i = 0; i < n0; i++) rn += rv[i] * rv[i];
      done = k;
      if (hist) hist[k] = sqrt(rn) / r0;
      if (sqrt(rn) / r0 < tol) break;
   }
   if (x_out) memcpy(x_out, xv.data(), sizeof(double) * n0);
   delete dm;
   return done;
}

// DMEM_Mult / DMEM_MultCycle (src/DMEM_Mult.cpp:13-261), the reference's object code on one rank: the multiplicative V(1,1)
// comparator of the DMEM driver, Jacobi branch (grid_relax_type[1] == 0: u += relax_weight v / a_ii before the restriction and
// after the prolongation), hypre_GaussElimSolve on the coarsest level, R_array applied transposed.  One cycle per call of
// DMEM_Mult (it re-reads x and b on entry); DMEM_MultCycle repoints F_array[0] / U_array[0] at its arguments, so they are
// pointed back at hypre's own vectors before every call.
int ref_dmem_mult(int L, const RefCSR *A, const RefCSR *P, const RefCSR *Rt, double smooth_weight, const double *b, int num_cycles,
                  double tol, double *x_out, double *hist)
{
   DMEM_AllData *dm = new DMEM_AllData();
   std::vector<hypre_CSRMatrix> hA(L), hP(L), hR(L);
   std::vector<hypre_ParCSRMatrix> pA(L), pP(L), pR(L);
   std::vector<hypre_ParCSRMatrix *> Aarr(L), Parr(L), Rarr(L);
   std::vector<std::vector<double>> ud(L), fd(L);
   std::vector<hypre_Vector> uv(L), fv(L);
   std::vector<hypre_ParVector> up(L), fp(L);
   std::vector<hypre_ParVector *> Uarr(L), Farr(L);
   std::vector<double> rw(L, smooth_weight);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); memset(&pA[l], 0, sizeof(pA[l])); pA[l].diag = &hA[l]; pA[l].global_num_rows = A[l].nrows; Aarr[l] = &pA[l];
      if (l < L - 1) {
         fill(&hP[l], P[l]); memset(&pP[l], 0, sizeof(pP[l])); pP[l].diag = &hP[l]; pP[l].global_num_rows = P[l].nrows; Parr[l] = &pP[l];
         fill(&hR[l], Rt[l]); memset(&pR[l], 0, sizeof(pR[l])); pR[l].diag = &hR[l]; pR[l].global_num_rows = Rt[l].nrows; Rarr[l] = &pR[l];
      }
      ud[l].assign(A[l].nrows, 0.0); fd[l].assign(A[l].nrows, 0.0);
      uv[l].data = ud[l].data(); uv[l].size = A[l].nrows; up[l].local_vector = &uv[l]; Uarr[l] = &up[l];
      fv[l].data = fd[l].data(); fv[l].size = A[l].nrows; fp[l].local_vector = &fv[l]; Farr[l] = &fp[l];
   }
   const int n0 = A[0].nrows;
   std::vector<double> vt(n0, 0.0), xv(n0, 0.0), rv(b, b + n0), bv(b, b + n0), ev(n0, 0.0), dv(n0, 0.0);
   hypre_Vector hv[6];
   hypre_ParVector pv[6];
   double *ptrs[6] = {vt.data(), xv.data(), rv.data(), bv.data(), ev.data(), dv.data()};
   for (int k = 0; k < 6; k++) { hv[k].data = ptrs[k]; hv[k].size = n0; pv[k].local_vector = &hv[k]; }
   hypre_ParAMGData amg;
   memset(&amg, 0, sizeof(amg));
   int relax_type[4] = {0, 0, 0, 0};
   amg.A_array = Aarr.data(); amg.P_array = Parr.data(); amg.R_array = Rarr.data();
   amg.F_array = Farr.data(); amg.U_array = Uarr.data(); amg.Vtemp = &pv[0]; amg.Ztemp = &pv[0];
   amg.num_levels = L; amg.grid_relax_type = relax_type; amg.relax_weight = rw.data(); amg.functional_gauss_elim = 1;
   dm->hypre.solver = (HYPRE_Solver)&amg;
   dm->matrix.A_fine = &pA[0];
   dm->input.solver = MULT;
   dm->input.tol = tol;
   dm->input.num_cycles = 1;
   dm->input.accel_type = NO_ACCEL;
   dm->input.delay_flag = 0;
   dm->vector_fine.x = &pv[1]; dm->vector_fine.r = &pv[2]; dm->vector_fine.b = &pv[3]; dm->vector_fine.e = &pv[4];
   dm->vector_fine.d = &pv[5];
   double r0 = 0.0;
   for (int i = 0; i < n0; i++) r0 += b[i] * b[i];
   r0 = sqrt(r0);
   dm->output.r0_norm2 = r0;
   if (hist) hist[0] = 1.0;
   int done = 0;
   for (int k = 1; k <= num_cycles; k++) {
      Uarr[0] = &up[0]; Farr[0] = &fp[0]; Aarr[0] = &pA[0];
      DMEM_Mult(dm);
      double rn = 0.0;
      for (int i = 0; i < n0; i++) rn += rv[i] * rv[i];
      done = k;
      if (hist) hist[k] = sqrt(rn) / r0;
      if (sqrt(rn) / r0 < tol) break;
   }
   if (x_out) memcpy(x_out, xv.data(), sizeof(double) * n0);
   delete dm;
   return done;
}

// DMEM_ChebyUpdate (src/DMEM_Misc.cpp:612-666), synchronous branch: d and u of length n, `cycle` = iter.cycle; c / c_prev are
// the recurrence state (in / out)
void ref_dmem_cheby_update(int n, double *d, double *u, int cycle, double mu, double delta, int accel_type, double *c, double *c_prev)
{
   DMEM_AllData *dm = new DMEM_AllData();
   hypre_Vector dv, uv;
   hypre_ParVector dp, upv;
   dv.data = d; dv.size = n; dp.local_vector = &dv;
   uv.data = u; uv.size = n; upv.local_vector = &uv;
   dm->cheby.mu = mu; dm->cheby.delta = delta; dm->cheby.c = *c; dm->cheby.c_prev = *c_prev;
   dm->iter.cycle = cycle;
   dm->input.accel_type = accel_type;
   dm->input.solver = SYNC_MULTADD;
   dm->input.async_flag = 0;
   DMEM_ChebyUpdate(dm, &dp, &upv, n);
   *c = dm->cheby.c; *c_prev = dm->cheby.c_prev;
   delete dm;
}

// AddCycle (src/DMEM_Add.cpp:180-329) + DMEM_AddSmooth (src/DMEM_Smooth.cpp:574-638), the reference's object code, for every
// grid k in turn on ONE rank: r = b - A x, grid k's chain (restrict k times with R_array applied transposed, smooth with
// wJacobi_scale = d/omega -- symmetrised when simple_jacobi_flag = -1 -- or the direct solve on the last grid, prolong k
// times), x += U_array[0].  That is DMEM_Add's asynchronous loop (:101-130) when the grids never overlap.  `rounds` passes
// over all grids; hist[round] = ||b - A x|| / ||b||.
int ref_dmem_add_cycles(int L, const RefCSR *A, const RefCSR *P, const RefCSR *Rt, double smooth_weight, int symmetrised,
                        const double *b, int rounds, double *x_out, double *hist)
{
   DMEM_AllData *dm = new DMEM_AllData();
   std::vector<hypre_CSRMatrix> hA(L), hP(L), hR(L);
   std::vector<hypre_ParCSRMatrix> pA(L), pP(L), pR(L);
   std::vector<hypre_ParCSRMatrix *> Aarr(L), Parr(L), Rarr(L);
   std::vector<std::vector<double>> ud(L), fd(L), sc(L), ssc(L);
   std::vector<hypre_Vector> uv(L), fv(L);
   std::vector<hypre_ParVector> up(L), fp(L);
   std::vector<hypre_ParVector *> Uarr(L), Farr(L);
   std::vector<double *> scp(L), sscp(L);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); memset(&pA[l], 0, sizeof(pA[l])); pA[l].diag = &hA[l]; pA[l].global_num_rows = A[l].nrows; Aarr[l] = &pA[l];
      if (l < L - 1) {
         fill(&hP[l], P[l]); memset(&pP[l], 0, sizeof(pP[l])); pP[l].diag = &hP[l]; pP[l].global_num_rows = P[l].nrows; Parr[l] = &pP[l];
         fill(&hR[l], Rt[l]); memset(&pR[l], 0, sizeof(pR[l])); pR[l].diag = &hR[l]; pR[l].global_num_rows = Rt[l].nrows; Rarr[l] = &pR[l];
      }
      const int n = A[l].nrows;
      ud[l].assign(n, 0.0); fd[l].assign(n, 0.0); sc[l].resize(n); ssc[l].resize(n);
      for (int i = 0; i < n; i++) {            // src/DMEM_Setup.cpp:465-483
         const double d = A[l].data[A[l].i[i]];
         sc[l][i] = d == 0.0 ? 1.0 : d / smooth_weight;
         ssc[l][i] = -sc[l][i];
      }
      scp[l] = sc[l].data(); sscp[l] = ssc[l].data();
      uv[l].data = ud[l].data(); uv[l].size = n; up[l].local_vector = &uv[l]; Uarr[l] = &up[l];
      fv[l].data = fd[l].data(); fv[l].size = n; fp[l].local_vector = &fv[l]; Farr[l] = &fp[l];
   }
   const int n0 = A[0].nrows;
   std::vector<double> vt(n0, 0.0), x(n0, 0.0);
   hypre_Vector hv; hv.data = vt.data(); hv.size = n0;
   hypre_ParVector pvt; pvt.local_vector = &hv;
   hypre_ParAMGData amg;
   memset(&amg, 0, sizeof(amg));
   amg.A_array = Aarr.data(); amg.P_array = Parr.data(); amg.R_array = Rarr.data();
   amg.F_array = Farr.data(); amg.U_array = Uarr.data(); amg.Vtemp = &pvt; amg.Ztemp = &pvt;
   amg.num_levels = L; amg.functional_gauss_elim = 1;
   dm->hypre.solver_gridk = (HYPRE_Solver)&amg;
   dm->grid.num_levels = L;
   dm->input.solver = MULTADD;
   dm->input.smoother = JACOBI;
   dm->input.coarsest_mult_level = 0;
   dm->input.num_interpolants = ONE_INTERPOLANT + 1;          // anything but ONE_INTERPOLANT: level-by-level transfers
   dm->input.async_flag = 0;
   dm->input.simple_jacobi_flag = symmetrised ? -1 : 0;
   dm->matrix.wJacobi_scale_gridk = scp.data();
   dm->matrix.symmwJacobi_scale_gridk = sscp.data();
   dm->output.level_wtime.assign(L + 1, 0.0);
   double r0 = 0.0;
   for (int i = 0; i < n0; i++) r0 += b[i] * b[i];
   r0 = sqrt(r0);
   if (hist) hist[0] = 1.0;
   hypre_CSRMatrix *A0 = &hA[0];
   auto residual = [&](double *r) {
      double s = 0.0;
      for (int i = 0; i < n0; i++) {
         double t = 0.0;
         for (int jj = A0->i[i]; jj < A0->i[i + 1]; jj++) t += A0->data[jj] * x[A0->j[jj]];
         r[i] = b[i] - t;
         s += r[i] * r[i];
      }
      return sqrt(s);
   };
   double rn = residual(fd[0].data());
   for (int k = 1; k <= rounds; k++) {
      for (int grid = 0; grid < L; grid++) {
         dm->grid.my_grid = grid;
         AddCycle(dm);
         for (int i = 0; i < n0; i++) x[i] += ud[0][i];
         rn = residual(fd[0].data());
      }
      if (hist) hist[k] = rn / r0;
   }
   if (x_out) memcpy(x_out, x.data(), sizeof(double) * n0);
   delete dm;
   return rounds;
}

// DMEM_AsyncSmooth (src/DMEM_Smooth.cpp:16-313), the reference's object code, on ONE rank: the asynchronous fine-grid smoother
// with no neighbour and no other grid to hear from -- u = r ./ s, x += u, r -= A_diag u, `num_cycles` relaxations (LOCAL rule,
// AsyncSmoothCheckConverge :340-349).  smoother: ASYNC_JACOBI (8) with s = d / omega, ASYNC_L1_JACOBI (10) with s = l1.
int ref_dmem_async_smooth(const RefCSR *A, const double *b, double smooth_weight, const double *l1, int smoother, int num_cycles,
                          double *x_out, double *r_out)
{
   DMEM_AllData *dm = new DMEM_AllData();
   const int n = A->nrows;
   hypre_CSRMatrix hA, hOffd;
   fill(&hA, *A);
   std::vector<int> offd_i((size_t)n + 1, 0);
   memset(&hOffd, 0, sizeof(hOffd));
   hOffd.i = offd_i.data(); hOffd.num_rows = n; hOffd.num_cols = 0; hOffd.num_nonzeros = 0;
   hypre_ParCSRMatrix pA;
   memset(&pA, 0, sizeof(pA));
   pA.diag = &hA; pA.offd = &hOffd; pA.global_num_rows = n;
   hypre_ParCSRMatrix *Aarr[1] = {&pA};
   std::vector<double> u(n, 0.0), f(n, 0.0), vt(n, 0.0), x(n, 0.0), r(b, b + n), bb(b, b + n), e(n, 0.0), d(n, 0.0), scale(n), ghost(1, 0.0);
   for (int i = 0; i < n; i++) scale[i] = (smoother == ASYNC_L1_JACOBI) ? l1[i] : A->data[A->i[i]] / smooth_weight;
   double *ptrs[8] = {u.data(), f.data(), vt.data(), x.data(), r.data(), bb.data(), e.data(), d.data()};
   hypre_Vector hv[8];
   hypre_ParVector pv[8];
   for (int k = 0; k < 8; k++) { hv[k].data = ptrs[k]; hv[k].size = n; pv[k].local_vector = &hv[k]; }
   hypre_Vector gv[2];
   for (int k = 0; k < 2; k++) { gv[k].data = ghost.data(); gv[k].size = 0; }
   hypre_ParVector *Uarr[1] = {&pv[0]}, *Farr[1] = {&pv[1]};
   hypre_ParAMGData amg;
   memset(&amg, 0, sizeof(amg));
   amg.A_array = Aarr; amg.U_array = Uarr; amg.F_array = Farr; amg.Vtemp = &pv[2]; amg.Ztemp = &pv[2]; amg.num_levels = 1;
   dm->hypre.solver_gridk = (HYPRE_Solver)&amg;
   dm->vector_gridk.x = &pv[3]; dm->vector_gridk.r = &pv[4]; dm->vector_gridk.b = &pv[5]; dm->vector_gridk.e = &pv[6];
   dm->vector_gridk.d = &pv[7]; dm->vector_gridk.x_ghost = &gv[0]; dm->vector_gridk.x_ghost_prev = &gv[1];
   double *sp[1] = {scale.data()};
   dm->matrix.wJacobi_scale_gridk = sp;
   dm->matrix.L1_row_norm_gridk = sp;
   dm->input.smoother = smoother;
   dm->input.smooth_weight = smooth_weight;
   dm->input.converge_test_type = LOCAL_CONVERGE;
   dm->input.num_cycles = num_cycles;
   dm->input.accel_type = NO_ACCEL;
   dm->input.async_flag = 1;
   dm->iter.cycle = 0; dm->iter.relax = 0;
   DMEM_AsyncSmooth(dm, 0);
   const int relax = dm->iter.relax;
   if (x_out) memcpy(x_out, x.data(), sizeof(double) * n);
   if (r_out) memcpy(r_out, r.data(), sizeof(double) * n);
   delete dm;
   return relax;
}

// ChebySetup -> EigsPower -> BPXCycle (src/SMEM_Cheby.cpp:28-60,410-518,520-645), the reference's object code: power iteration on
// B A with B = its hypre-vector BPX cycle (restriction by hypre_ParCSRMatrixMatvecT with R_array = P_array, as hypre keeps
// it; (L1-)Jacobi / hybrid sweeps from zero with diag_scale = a_ii / omega or the L1 norms on EVERY level; prolongation with
// beta = 1).  out[4] = alpha (eig_min), beta (eig_max), mu, delta; f_after (may be null) receives F_array[0] as EigsPower
// leaves it (it restores the right-hand side it saved in vector.y[0], :512-516).
int ref_cheby_setup(int L, const RefCSR *A, const RefCSR *P, double *const *l1, int smoother, double smooth_weight, int num_sweeps,
                    int iters, int num_threads, const double *f0, double *out, double *f_after)
{
   AllData *ad = new AllData();
   memset((void *)&ad->input, 0, sizeof(ad->input));
   memset((void *)&ad->matrix, 0, sizeof(ad->matrix));
   memset((void *)&ad->cheby, 0, sizeof(ad->cheby));
   std::vector<hypre_CSRMatrix> hA(L), hP(L);
   std::vector<hypre_ParCSRMatrix> pA(L), pP(L);
   std::vector<hypre_ParCSRMatrix *> Aarr(L), Parr(L);
   std::vector<std::vector<double>> ud(L), fd(L), dow(L);
   std::vector<hypre_Vector> uv(L), fv(L);
   std::vector<hypre_ParVector> up(L), fp(L);
   std::vector<hypre_ParVector *> Uarr(L), Farr(L);
   std::vector<double *> dowp(L);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); memset(&pA[l], 0, sizeof(pA[l])); pA[l].diag = &hA[l]; pA[l].global_num_rows = A[l].nrows; Aarr[l] = &pA[l];
      if (l < L - 1) { fill(&hP[l], P[l]); memset(&pP[l], 0, sizeof(pP[l])); pP[l].diag = &hP[l]; pP[l].global_num_rows = P[l].nrows; Parr[l] = &pP[l]; }
      const int n = A[l].nrows;
      ud[l].assign(n, 0.0); fd[l].assign(n, 0.0); dow[l].resize(n);
      for (int i = 0; i < n; i++) dow[l][i] = A[l].data[A[l].i[i]] / smooth_weight;      // src/SMEM_Setup.cpp:233-238
      dowp[l] = dow[l].data();
      uv[l].data = ud[l].data(); uv[l].size = n; up[l].local_vector = &uv[l]; Uarr[l] = &up[l];
      fv[l].data = fd[l].data(); fv[l].size = n; fp[l].local_vector = &fv[l]; Farr[l] = &fp[l];
   }
   const int n0 = A[0].nrows;
   memcpy(fd[0].data(), f0, sizeof(double) * n0);
   std::vector<double> vt(n0, 0.0), e0(n0, 0.0), y0(n0, 0.0);
   hypre_Vector hv; hv.data = vt.data(); hv.size = n0;
   hypre_ParVector pvt; pvt.local_vector = &hv;
   hypre_ParAMGData amg;
   memset(&amg, 0, sizeof(amg));
   HYPRE_Int relax_types[4] = {0, 0, 0, 0};
   amg.A_array = Aarr.data(); amg.P_array = Parr.data(); amg.R_array = Parr.data();
   amg.F_array = Farr.data(); amg.U_array = Uarr.data(); amg.Vtemp = &pvt; amg.Ztemp = &pvt;
   amg.num_levels = L; amg.l1_norms = (HYPRE_Real **)l1; amg.grid_relax_type = relax_types;
   ad->hypre.solver = (HYPRE_Solver)&amg;
   ad->hypre.print_level = 0;
   ad->grid.num_levels = L;
   ad->input.solver = BPX;
   ad->input.smoother = smoother;
   ad->input.smooth_weight = smooth_weight;
   ad->input.num_pre_smooth_sweeps = num_sweeps;
   ad->input.num_threads = num_threads;
   ad->input.cheby_eig_type = CHEBY_EIG_POWER;
   ad->input.cheby_eig_max_iters = iters;
   ad->input.format_output_flag = 1;
   ad->input.num_cycles = 1;
   ad->matrix.A_diag = dowp.data();
   double *ep[1] = {e0.data()}, *yp[1] = {y0.data()};
   ad->vector.e = ep; ad->vector.y = yp;
   const int saved = omp_get_max_threads();
   omp_set_num_threads(num_threads);
   ChebySetup(ad);
   omp_set_num_threads(saved);
   out[0] = ad->cheby.alpha; out[1] = ad->cheby.beta; out[2] = ad->cheby.mu; out[3] = ad->cheby.delta;
   if (f_after) memcpy(f_after, fd[0].data(), sizeof(double) * n0);
   const int ok = (ad->cheby.omega != nullptr);
   delete ad;
   return ok ? 0 : 1;
}

// SmoothTransfer (src/SMEM_Setup.cpp:1173-1254: G = I - w D^-1 A / GT = I - w A D^-1 on A's pattern, or their L1 forms; Pbar = G P,
// Rbar = P^T GT through EigenMatMat :1256-1339 and the row layout of StdVector_to_CSR :1372-1424), the reference's object code,
// for ONE level: A (diag first), plain P (hypre's R_array is P as well).  out_P / out_R borrow arrays the reference malloc'ed
// (ref_free them); an output the reference leaves untouched (sweeps == 0) comes back with nrows = -1.
int ref_smooth_transfer(const RefCSR *A, const RefCSR *P, double *l1, double smooth_weight, int smooth_interp_type, int num_pre,
                        int num_post, int num_threads, RefCSR *out_P, RefCSR *out_R)
{
   AllData *ad = new AllData();
   memset((void *)&ad->input, 0, sizeof(ad->input));
   memset((void *)&ad->matrix, 0, sizeof(ad->matrix));
   hypre_CSRMatrix hA, hP;
   fill(&hA, *A); fill(&hP, *P);
   hypre_CSRMatrix *Aarr[1] = {&hA};
   hypre_CSRMatrix *Parr[1] = {(hypre_CSRMatrix *)calloc(1, sizeof(hypre_CSRMatrix))};
   hypre_CSRMatrix *Rarr[1] = {(hypre_CSRMatrix *)calloc(1, sizeof(hypre_CSRMatrix))};
   double *l1arr[1] = {l1};
   ad->matrix.A = Aarr; ad->matrix.P = Parr; ad->matrix.R = Rarr; ad->matrix.L1_row_norm = l1arr;
   ad->input.smooth_weight = smooth_weight;
   ad->input.smooth_interp_type = smooth_interp_type;
   ad->input.num_pre_smooth_sweeps = num_pre;
   ad->input.num_post_smooth_sweeps = num_post;
   ad->input.num_threads = num_threads;
   Parr[0]->num_rows = -1; Rarr[0]->num_rows = -1;
   omp_set_num_threads(num_threads);
   SmoothTransfer(ad, &hP, &hP, 0);
   auto give = [](hypre_CSRMatrix *m, RefCSR *o) {
      o->nrows = m->num_rows; o->ncols = m->num_cols; o->nnz = m->num_nonzeros; o->i = m->i; o->j = m->j; o->data = m->data;
      free(m);
   };
   give(Parr[0], out_P); give(Rarr[0], out_R);
   delete ad;
   return 0;
}

// ComputeWork (src/SMEM_Setup.cpp:1038-1170), PartitionLevels with BALANCED_THREADS (:590-868) and PartitionGrids (:895-1036), the
// reference's object code: the work model that sizes the level groups, the threads dealt to every level, and every thread's
// nnz-balanced row range of A on every level (= the hybrid smoother's Gauss-Seidel blocks).  Outputs: level_work[L],
// frac[L], threads_per_level[L], A_ns / A_ne [L * num_threads] (entry [l * num_threads + t]; -1 where the reference leaves
// the slot unset: thread t is not in a group that reaches it).
int ref_work_partition(int L, const RefCSR *A, const RefCSR *P, const RefCSR *R, int solver, int num_pre, int num_post, int fine_sweeps,
                       int coarse_sweeps, int num_threads, int *level_work, double *frac, int *threads_per_level, int *A_ns, int *A_ne)
{
   AllData *ad = new AllData();
   memset((void *)&ad->input, 0, sizeof(ad->input));
   memset((void *)&ad->matrix, 0, sizeof(ad->matrix));
   std::vector<hypre_CSRMatrix> hA(L), hP(L), hR(L);
   std::vector<hypre_CSRMatrix *> Aarr(L), Parr(L), Rarr(L);
   std::vector<int> n(L);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); Aarr[l] = &hA[l]; n[l] = A[l].nrows;
      if (l < L - 1) { fill(&hP[l], P[l]); fill(&hR[l], R[l]); Parr[l] = &hP[l]; Rarr[l] = &hR[l]; }
   }
   ad->matrix.A = Aarr.data(); ad->matrix.P = Parr.data(); ad->matrix.R = Rarr.data();
   ad->grid.num_levels = L; ad->grid.n = n.data();
   ad->input.solver = solver;
   ad->input.res_compute_type = LOCAL;
   ad->input.num_pre_smooth_sweeps = num_pre; ad->input.num_post_smooth_sweeps = num_post;
   ad->input.num_fine_smooth_sweeps = fine_sweeps; ad->input.num_coarse_smooth_sweeps = coarse_sweeps;
   ad->input.num_threads = num_threads;
   ad->input.thread_part_type = ALL_LEVELS;
   ad->input.thread_part_distr_type = BALANCED_THREADS;
   ad->input.construct_R_flag = 1;
   ComputeWork(ad);
   PartitionLevels(ad);
   // PartitionGrids deals the rows of every level with `while (count < n) for (i < num_level_threads) ...` (:1003-1011): a level that
   // BALANCED_THREADS left without a thread makes that loop spin forever -- the reference HANGS there (SURVEY.md 5.9d says "never
   // corrected"; it never gets that far).  Skip the call in that case and report the thread counts only.
   bool empty_level = false;
   for (int l = 0; l < L; l++) empty_level = empty_level || ad->thread.level_threads[l].empty();
   if (!empty_level) PartitionGrids(ad);
   for (int l = 0; l < L; l++) {
      level_work[l] = ad->grid.level_work[l];
      frac[l] = ad->grid.frac_level_work[l];
      threads_per_level[l] = (int)ad->thread.level_threads[l].size();
   }
   for (int i = 0; i < L * num_threads; i++) A_ns[i] = A_ne[i] = -1;
   // a thread's range on inner level l is set by the group it belongs to (PartitionGrids writes [inner_level][t])
   for (int k = 0; k < L && !empty_level; k++)
      for (int t : ad->thread.level_threads[k])
         for (int l = 0; l < L; l++) { A_ns[l * num_threads + t] = ad->thread.A_ns[l][t]; A_ne[l * num_threads + t] = ad->thread.A_ne[l][t]; }
   delete ad;
   return empty_level ? 1 : 0;
}

// BuildExtendedMatrix (src/SMEM_Setup.cpp:1426-1521), the reference's object code, EXPLICIT_EXTENDED_SYSTEM_BPX: blocks
// A_k P_k .. P_{l-1} / R_{l-1} .. R_k A_k^T through hypre_CSRMatrixMultiply / Transpose (stand-ins above), row layout by
// StdVector_to_CSR.  The result borrows arrays the reference calloc'ed (ref_free them).
int ref_build_extended_matrix(int L, const RefCSR *A, const RefCSR *P, const RefCSR *R, RefCSR *out, int *disp)
{
   AllData *ad = new AllData();
   memset((void *)&ad->input, 0, sizeof(ad->input));
   std::vector<hypre_CSRMatrix> hA(L), hP(L), hR(L);
   std::vector<hypre_CSRMatrix *> Aarr(L), Parr(L), Rarr(L);
   for (int l = 0; l < L; l++) {
      fill(&hA[l], A[l]); Aarr[l] = &hA[l];
      if (l < L - 1) { fill(&hP[l], P[l]); fill(&hR[l], R[l]); Parr[l] = &hP[l]; Rarr[l] = &hR[l]; }
   }
   ad->grid.num_levels = L;
   ad->input.solver = EXPLICIT_EXTENDED_SYSTEM_BPX;
   ad->input.construct_R_flag = 1;
   ad->input.format_output_flag = 1;
   hypre_CSRMatrix *B = nullptr;
   BuildExtendedMatrix(ad, Aarr.data(), Parr.data(), Rarr.data(), &B);
   out->nrows = B->num_rows; out->ncols = B->num_cols; out->nnz = B->num_nonzeros; out->i = B->i; out->j = B->j; out->data = B->data;
   for (int l = 0; l <= L; l++) disp[l] = ad->grid.disp[l];
   free(B);
   delete ad;
   return 0;
}

void ref_destroy(void *h) { delete (RefHandle *)h; }

// direct kernel entry points for unit parity (called SPMD over one [ns,ne) = all rows)
void ref_matvec(const RefCSR *A, double *x, double *y)
{
   hypre_CSRMatrix m; fill(&m, *A);
   SMEM_MatVec(nullptr, &m, x, y, 0, A->nrows);
}
// SMEM_Sync_Parfor_MatVecT (src/SMEM_MatVec.cpp:27-58; the -no_construct_R restriction): scatter into per-thread copies of y,
// then the copies are summed in thread order.
void ref_parfor_matvec_t(const RefCSR *A, double *x, double *y, int num_threads)
{
   hypre_CSRMatrix m; fill(&m, *A);
   AllData *ad = new AllData();
   ad->input.num_threads = num_threads;
   std::vector<double> ye((size_t)num_threads * A->ncols, 0.0);
   double *yep = ye.data();
   #pragma omp parallel num_threads(num_threads)
   {
      SMEM_Sync_Parfor_MatVecT(ad, &m, x, y, yep);
   }
   delete ad;
}
void ref_seq_symmetric_jacobi(const RefCSR *A, double *f, double *u, double w, int sweeps)
{
   hypre_CSRMatrix m; fill(&m, *A);
   AllData ad; ad.input.smooth_weight = w;
   std::vector<double> y(A->nrows), r(A->nrows);
   SEQ_SymmetricJacobi(&ad, &m, f, u, y.data(), r.data(), sweeps, 0);
}
void ref_seq_jacobi(const RefCSR *A, double *f, double *u, double w, int sweeps, int zero_flag)
{
   hypre_CSRMatrix m; fill(&m, *A);
   AllData ad; ad.input.smooth_weight = w;
   int zf = zero_flag; ad.grid.zero_flags = &zf;
   std::vector<double> up(A->nrows);
   SEQ_Jacobi(&ad, &m, f, u, up.data(), sweeps, 0);
}

// The reference's own matrix-file reader (`-problem file`, src/SMEM_Setup.cpp:1646-1650 -> ReadBinary_fread_HypreParCSR,
// src/Misc.cpp:800-915) on `path`; the caller frees the three arrays with ref_free.
int ref_read_matrix(const char *path, int symm_flag, int *nrows, int *nnz, int **ri, int **rj, double **rdata)
{
   FILE *fp = fopen(path, "rb");
   if (!fp) return 1;
   hypre_ParCSRMatrix *par = nullptr;
   ReadBinary_fread_HypreParCSR(fp, &par, symm_flag, 0);
   fclose(fp);
   hypre_CSRMatrix *A = hypre_ParCSRMatrixDiag(par);
   *nrows = A->num_rows; *nnz = A->num_nonzeros; *ri = A->i; *rj = A->j; *rdata = A->data;
   return 0;
}
void ref_free(void *p) { free(p); }
int ref_max_threads(void) { return omp_get_num_procs(); }   // (omp_get_max_threads follows the last omp_set_num_threads of a solve)
}
