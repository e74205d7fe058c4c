/* BuildHypreMatrix.hpp -- TEST INFRASTRUCTURE.  src/SMEM_Setup.cpp:10 includes a header of this name which the reference repository
 * does not contain.  This stand-in declares what that translation unit names beyond oracle/ref_shim/hypre_stub.h:
 * hypre's set-up interface (HYPRE_BoomerAMG*, HYPRE_IJVector*: un-vendored, never reached by the driver -- they abort) and
 * hypre_CSRMatrixTranspose / Multiply / Create (functional, oracle/ref_driver.cpp).  Only SmoothTransfer / EigenMatMat / CSR_Transpose /
 * StdVector_to_CSR / ComputeWork / PartitionLevels / PartitionGrids of SMEM_Setup.cpp are ever called. */
#ifndef AMG_REF_SETUP_STUB_H
#define AMG_REF_SETUP_STUB_H
#include <stdlib.h>
#define AMG_REF_NEVER3(name) template <class... T> static inline int name(T...) { abort(); return 0; }
AMG_REF_NEVER3(HYPRE_BoomerAMGCreate) AMG_REF_NEVER3(HYPRE_BoomerAMGSetup) AMG_REF_NEVER3(HYPRE_BoomerAMGSetStrongThreshold)
AMG_REF_NEVER3(HYPRE_BoomerAMGSetSimple) AMG_REF_NEVER3(HYPRE_BoomerAMGSetRestriction) AMG_REF_NEVER3(HYPRE_BoomerAMGSetRelaxWt)
AMG_REF_NEVER3(HYPRE_BoomerAMGSetRelaxType) AMG_REF_NEVER3(HYPRE_BoomerAMGSetPostInterpType) AMG_REF_NEVER3(HYPRE_BoomerAMGSetPMaxElmts)
AMG_REF_NEVER3(HYPRE_BoomerAMGSetNumSweeps) AMG_REF_NEVER3(HYPRE_BoomerAMGSetNumFunctions) AMG_REF_NEVER3(HYPRE_BoomerAMGSetMeasureType)
AMG_REF_NEVER3(HYPRE_BoomerAMGSetMaxRowSum) AMG_REF_NEVER3(HYPRE_BoomerAMGSetMaxLevels) AMG_REF_NEVER3(HYPRE_BoomerAMGSetInterpType)
AMG_REF_NEVER3(HYPRE_BoomerAMGSetCycleRelaxType) AMG_REF_NEVER3(HYPRE_BoomerAMGSetCoarsenType) AMG_REF_NEVER3(HYPRE_BoomerAMGSetAggNumLevels)
AMG_REF_NEVER3(HYPRE_BoomerAMGSetAdditive) AMG_REF_NEVER3(HYPRE_BoomerAMGSetAddRelaxWt) AMG_REF_NEVER3(HYPRE_BoomerAMGSetAddRelaxType)
AMG_REF_NEVER3(HYPRE_IJVectorCreate) AMG_REF_NEVER3(HYPRE_IJVectorSetObjectType) AMG_REF_NEVER3(HYPRE_IJVectorInitialize)
AMG_REF_NEVER3(HYPRE_IJVectorSetValues) AMG_REF_NEVER3(HYPRE_IJVectorAssemble) AMG_REF_NEVER3(HYPRE_IJVectorGetObject)
AMG_REF_NEVER3(hypre_GaussElimSetup) AMG_REF_NEVER3(MPI_Finalize)
/* functional single-rank stand-ins (oracle/ref_driver.cpp), used by BuildExtendedMatrix (src/SMEM_Setup.cpp:1426-1521) */
hypre_CSRMatrix *hypre_CSRMatrixCreate(HYPRE_Int num_rows, HYPRE_Int num_cols, HYPRE_Int num_nonzeros);
hypre_CSRMatrix *hypre_CSRMatrixMultiply(hypre_CSRMatrix *A, hypre_CSRMatrix *B);
HYPRE_Int hypre_CSRMatrixTranspose(hypre_CSRMatrix *A, hypre_CSRMatrix **AT, HYPRE_Int data);
/* src/BuildHypreMatrix.cpp (its own header is the missing file this stand-in replaces): hypre's stencil generators are un-vendored;
 * the driver's versions RECORD the stencil coefficients the reference computed (its own arithmetic, :104-289) and build nothing */
void BuildHypreMatrix(AllData *all_data, HYPRE_ParCSRMatrix *A_ptr, HYPRE_ParVector *rhs_ptr, MPI_Comm comm, HYPRE_Int nx, HYPRE_Int ny,
                      HYPRE_Int nz, HYPRE_Real cx, HYPRE_Real cy, HYPRE_Real cz, HYPRE_Real ax, HYPRE_Real ay, HYPRE_Real az, HYPRE_Real eps,
                      int atype);
HYPRE_ParCSRMatrix GenerateLaplacian(MPI_Comm comm, HYPRE_Int nx, HYPRE_Int ny, HYPRE_Int nz, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R,
                                     HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real *value);
HYPRE_ParCSRMatrix GenerateLaplacian27pt(MPI_Comm comm, HYPRE_Int nx, HYPRE_Int ny, HYPRE_Int nz, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R,
                                         HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real *value);
HYPRE_ParCSRMatrix GenerateDifConv(MPI_Comm comm, HYPRE_Int nx, HYPRE_Int ny, HYPRE_Int nz, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R,
                                   HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real *value);
HYPRE_ParCSRMatrix GenerateVarDifConv(MPI_Comm comm, HYPRE_Int nx, HYPRE_Int ny, HYPRE_Int nz, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R,
                                      HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real eps, HYPRE_ParVector *rhs_ptr);
#ifndef hypre_CTAlloc
#define hypre_CTAlloc(type, count, location) ((type *)calloc((size_t)(count), sizeof(type)))
#define hypre_TFree(ptr, location) free(ptr)
#endif
#ifndef hypre_printf
#define hypre_printf printf
#endif
#ifndef AMG_REF_DMEM_STUB_H
static inline int hypre_MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
#endif
extern int amg_ref_num_procs;   /* oracle/ref_driver.cpp: the communicator size BuildHypreMatrix's processor-grid search sees (rank stays 0) */
static inline int hypre_MPI_Comm_size(MPI_Comm, int *s) { *s = amg_ref_num_procs; return 0; }
#endif
