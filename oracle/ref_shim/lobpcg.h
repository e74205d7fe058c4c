/* lobpcg.h -- TEST INFRASTRUCTURE.  Stand-in for hypre's LOBPCG / multivector / IJVector interfaces (un-vendored) so that
 * src/SMEM_Cheby.cpp compiles UNMODIFIED; EigsHypreLOBPCG (their only user) aborts if reached. */
#ifndef AMG_REF_LOBPCG_STUB_H
#define AMG_REF_LOBPCG_STUB_H
#include <stdlib.h>
typedef struct { int dummy; } utilities_FortranMatrix;
typedef struct { double absolute, relative; } lobpcg_Tolerance;
typedef struct { int dummy; } mv_InterfaceInterpreter;
typedef void *mv_MultiVectorPtr;
typedef void *HYPRE_Matrix;
typedef void *HYPRE_Vector;
typedef struct {
   void *(*MatvecCreate)(void *A, void *x);
   HYPRE_Int (*Matvec)(void *matvec_data, HYPRE_Complex alpha, void *A, void *x, HYPRE_Complex beta, void *y);
   HYPRE_Int (*MatvecDestroy)(void *matvec_data);
   void *(*MatMultiVecCreate)(void *A, void *x);
   HYPRE_Int (*MatMultiVec)(void *data, HYPRE_Complex alpha, void *A, void *x, HYPRE_Complex beta, void *y);
   HYPRE_Int (*MatMultiVecDestroy)(void *data);
} HYPRE_MatvecFunctions;
#define AMG_REF_NEVER2(name) template <class... T> static inline int name(T...) { abort(); return 0; }
AMG_REF_NEVER2(HYPRE_IJVectorCreate) AMG_REF_NEVER2(HYPRE_IJVectorSetObjectType) AMG_REF_NEVER2(HYPRE_IJVectorInitialize)
AMG_REF_NEVER2(HYPRE_IJVectorSetValues) AMG_REF_NEVER2(HYPRE_IJVectorAssemble) AMG_REF_NEVER2(HYPRE_IJVectorGetObject)
AMG_REF_NEVER2(HYPRE_IJVectorDestroy) AMG_REF_NEVER2(HYPRE_ParCSRSetupInterpreter) AMG_REF_NEVER2(mv_MultiVectorSetRandom)
AMG_REF_NEVER2(HYPRE_LOBPCGCreate) AMG_REF_NEVER2(HYPRE_LOBPCGSetMaxIter) AMG_REF_NEVER2(HYPRE_LOBPCGSetPrecondUsageMode)
AMG_REF_NEVER2(HYPRE_LOBPCGSetTol) AMG_REF_NEVER2(HYPRE_LOBPCGSetPrintLevel) AMG_REF_NEVER2(HYPRE_LOBPCGSetup)
AMG_REF_NEVER2(HYPRE_LOBPCGSolve) AMG_REF_NEVER2(HYPRE_LOBPCGDestroy) AMG_REF_NEVER2(HYPRE_BoomerAMGSolve)
AMG_REF_NEVER2(HYPRE_BoomerAMGSetNumSweeps)
template <class... T> static inline mv_MultiVectorPtr mv_MultiVectorCreateFromSampleVector(T...) { abort(); return 0; }
#endif
