/* hypre_stub.h -- TEST INFRASTRUCTURE.  Minimal stand-in for the hypre headers the reference's
 * solve-phase translation units include (src/Main.hpp:34-39).  hypre itself (a fork, version
 * unpinned) is an un-vendored dependency; only the handful of types / accessor macros the
 * solve phase touches are declared, laid out the way the driver (ref_driver.cpp) fills them.
 * Functions that are only *declared* here are never reached on the paths the driver runs
 * (or are defined in ref_driver.cpp). */
#ifndef AMG_REF_HYPRE_STUB_H
#define AMG_REF_HYPRE_STUB_H
#include <stddef.h>
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define hypre_MPI_COMM_WORLD 0
typedef int HYPRE_Int;
typedef int HYPRE_BigInt;
typedef double HYPRE_Real;
typedef double HYPRE_Complex;

typedef struct {
   HYPRE_Int *i;
   HYPRE_Int *j;
   HYPRE_Int num_rows;
   HYPRE_Int num_cols;
   HYPRE_Int num_nonzeros;
   HYPRE_Real *data;
   HYPRE_Int *rownnz;
   HYPRE_Int num_rownnz;
} hypre_CSRMatrix;
#define hypre_CSRMatrixData(m) ((m)->data)
#define hypre_CSRMatrixI(m) ((m)->i)
#define hypre_CSRMatrixJ(m) ((m)->j)
#define hypre_CSRMatrixNumRows(m) ((m)->num_rows)
#define hypre_CSRMatrixNumCols(m) ((m)->num_cols)
#define hypre_CSRMatrixNumNonzeros(m) ((m)->num_nonzeros)
#define hypre_CSRMatrixRownnz(m) ((m)->rownnz)
#define hypre_CSRMatrixNumRownnz(m) ((m)->num_rownnz)

typedef struct { HYPRE_Real *data; HYPRE_Int size; } hypre_Vector;
#define hypre_VectorData(v) ((v)->data)
#define hypre_VectorSize(v) ((v)->size)
typedef struct { hypre_Vector *local_vector; } hypre_ParVector;
#define hypre_ParVectorLocalVector(v) ((v)->local_vector)
typedef struct { hypre_CSRMatrix *diag; HYPRE_Int global_num_rows;
                 /* named by the DMEM translation units only (dmem_stub.h); never read on the driver's single-rank paths */
                 hypre_CSRMatrix *offd; HYPRE_BigInt *row_starts; HYPRE_BigInt *col_map_offd; MPI_Comm comm; } hypre_ParCSRMatrix;
#define hypre_ParCSRMatrixDiag(m) ((m)->diag)
#define hypre_ParCSRMatrixNumRows(m) ((m)->global_num_rows)
typedef struct {
   hypre_ParCSRMatrix **A_array, **P_array, **R_array;
   hypre_ParVector **F_array, **U_array;
   hypre_ParVector *Vtemp, *Ztemp;
   HYPRE_Real **l1_norms;
   /* DMEM side (src/DMEM_Mult.cpp): */
   HYPRE_Int num_levels; HYPRE_Int *grid_relax_type; HYPRE_Real add_rlx_wt; HYPRE_Int simple; HYPRE_Real *relax_weight;
   hypre_ParCSRMatrix **P_array_afacj;
   HYPRE_Int functional_gauss_elim;   /* driver flag: hypre_GaussElimSolve really solves (DMEM convention) */
} hypre_ParAMGData;
#define hypre_ParAMGDataAArray(d) ((d)->A_array)
#define hypre_ParAMGDataPArray(d) ((d)->P_array)
#define hypre_ParAMGDataRArray(d) ((d)->R_array)
#define hypre_ParAMGDataFArray(d) ((d)->F_array)
#define hypre_ParAMGDataUArray(d) ((d)->U_array)
#define hypre_ParAMGDataVtemp(d) ((d)->Vtemp)
#define hypre_ParAMGDataZtemp(d) ((d)->Ztemp)
#define hypre_ParAMGDataL1Norms(d) ((d)->l1_norms)

typedef void *HYPRE_Solver;
typedef hypre_ParCSRMatrix *HYPRE_ParCSRMatrix;
typedef hypre_ParVector *HYPRE_ParVector;
typedef void *HYPRE_IJMatrix;
typedef void *HYPRE_IJVector;
#define HYPRE_PARCSR 5555

/* nnz-balanced per-thread row partition: hypre csr_matrix.c
 * hypre_CSRMatrixGetLoadBalancedPartitionBoundary (published algorithm, restated in
 * ref_driver.cpp) */
HYPRE_Int hypre_CSRMatrixGetLoadBalancedPartitionBegin(hypre_CSRMatrix *A);
HYPRE_Int hypre_CSRMatrixGetLoadBalancedPartitionEnd(hypre_CSRMatrix *A);
HYPRE_Int *hypre_LowerBound(HYPRE_Int *first, HYPRE_Int *last, HYPRE_Int value);

/* declared only (not reached on the driver's paths, or no-ops defined in ref_driver.cpp) */
HYPRE_Int hypre_GaussElimSolve(hypre_ParAMGData *amg_data, HYPRE_Int level, HYPRE_Int relax_type);
HYPRE_Int HYPRE_BoomerAMGSetPrintLevel(HYPRE_Solver solver, HYPRE_Int print_level);
HYPRE_Int HYPRE_BoomerAMGSetMaxIter(HYPRE_Solver solver, HYPRE_Int max_iter);
HYPRE_Int hypre_ParVectorSetConstantValues(hypre_ParVector *v, HYPRE_Complex value);
HYPRE_Int hypre_ParCSRMatrixMatvecOutOfPlace(HYPRE_Complex alpha, hypre_ParCSRMatrix *A, hypre_ParVector *x, HYPRE_Complex beta, hypre_ParVector *b, hypre_ParVector *y);
HYPRE_Int hypre_ParCSRMatrixMatvec(HYPRE_Complex alpha, hypre_ParCSRMatrix *A, hypre_ParVector *x, HYPRE_Complex beta, hypre_ParVector *y);
HYPRE_Int hypre_ParVectorCopy(hypre_ParVector *x, hypre_ParVector *y);
HYPRE_Int HYPRE_IJMatrixCreate(MPI_Comm comm, HYPRE_BigInt ilower, HYPRE_BigInt iupper, HYPRE_BigInt jlower, HYPRE_BigInt jupper, HYPRE_IJMatrix *matrix);
HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix matrix, HYPRE_Int type);
HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix matrix);
HYPRE_Int HYPRE_IJMatrixSetValues(HYPRE_IJMatrix matrix, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows, const HYPRE_BigInt *cols, const HYPRE_Complex *values);
HYPRE_Int HYPRE_IJMatrixAssemble(HYPRE_IJMatrix matrix);
HYPRE_Int HYPRE_IJMatrixGetObject(HYPRE_IJMatrix matrix, void **object);
#include "dmem_stub.h"
#endif
