/* slepceps.h -- TEST INFRASTRUCTURE.  Stand-in for SLEPc / PETSc (v3.15 in the reference's Makefile:195-196; absent here) so that
 * src/SMEM_Cheby.cpp compiles UNMODIFIED.  Only EigsPower / ChebySetup / BPXCycle of that file are ever called by the driver;
 * EigsSlepc (the only user of these names) aborts if reached. */
#ifndef AMG_REF_SLEPC_STUB_H
#define AMG_REF_SLEPC_STUB_H
#include <stdlib.h>
typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef void *Mat;
typedef void *Vec;
typedef void *EPS;
typedef const char *EPSType;
#define PETSC_COMM_WORLD 0
#define EPSARNOLDI "arnoldi"
enum { EPS_NHEP = 1, EPS_LARGEST_MAGNITUDE = 1, EPS_SMALLEST_MAGNITUDE = 2, MATOP_MULT = 3 };
#define AMG_REF_NEVER(name) template <class... T> static inline int name(T...) { abort(); return 0; }
AMG_REF_NEVER(SlepcInitialize) AMG_REF_NEVER(SlepcFinalize) AMG_REF_NEVER(MatCreateShell) AMG_REF_NEVER(MatShellSetOperation)
AMG_REF_NEVER(MatDestroy) AMG_REF_NEVER(MatShellGetContext) AMG_REF_NEVER(VecGetArray) AMG_REF_NEVER(VecRestoreArray)
AMG_REF_NEVER(VecGetArrayRead) AMG_REF_NEVER(VecRestoreArrayRead)
AMG_REF_NEVER(EPSCreate) AMG_REF_NEVER(EPSSetOperators) AMG_REF_NEVER(EPSSetProblemType) AMG_REF_NEVER(EPSSetTolerances)
AMG_REF_NEVER(EPSSetType) AMG_REF_NEVER(EPSSetFromOptions) AMG_REF_NEVER(EPSSetWhichEigenpairs) AMG_REF_NEVER(EPSSolve)
AMG_REF_NEVER(EPSGetConverged) AMG_REF_NEVER(EPSGetEigenpair) AMG_REF_NEVER(EPSGetIterationNumber) AMG_REF_NEVER(EPSDestroy)
#endif
