/* ref_prelude.hpp -- force-included (-include) before every reference translation unit.
 * src/Main.hpp is stale relative to the SMEM sources: the solve phase names
 * matrix.A_diag, L1_HYBRID_JACOBI_GAUSS_SEIDEL and CYCLE_PHASE_DOWN/UP, none of which it
 * declares (nor vector.u_smooth, which SMEM_ExtendedSystem.cpp uses).  The reference sources are compiled UNMODIFIED; this prelude injects the missing
 * member through the preprocessor (the token A_diag_ext in the struct body expands to
 * "A_diag_ext; double **A_diag") and supplies the three missing enumerators. */
#ifndef AMG_REF_PRELUDE_HPP
#define AMG_REF_PRELUDE_HPP
#define A_diag_ext A_diag_ext; double **A_diag; \
   double difconv_ax, difconv_ay, difconv_az, difconv_cx, difconv_cy, difconv_cz, vardifconv_eps; int difconv_atype   /* MatrixData: src/SMEM_Setup.cpp:1663-1666 */
#define add_P_max_elmts add_P_max_elmts; HYPRE_Int relax_type                                 /* HypreData: src/SMEM_Setup.cpp:1689-1693 */
#define smooth_interp_type smooth_interp_type; int simple_jacobi_flag; int hypre_memory       /* InputData: src/SMEM_Setup.cpp:1702, src/BuildHypreMatrix.cpp:117 */
#define z2 z2; HYPRE_Real **u_smooth   /* VectorData::u_smooth: used by src/SMEM_ExtendedSystem.cpp:374, allocated by src/SMEM_Setup.cpp:287 */
#include "Main.hpp"
#undef A_diag_ext
#undef add_P_max_elmts
#undef smooth_interp_type
#undef z2
#define L1_HYBRID_JACOBI_GAUSS_SEIDEL 12
#define CYCLE_PHASE_DOWN 0
#define CYCLE_PHASE_UP 1
#define MFEM_LAPLACE_AMR 9          /* named by src/SMEM_Setup.cpp:1606; never selected here */
/* src/SMEM_Solve.hpp:8-16 declares SMEM_Smooth with 10 parameters; the definition
 * (src/SMEM_Solve.cpp:264-273) and every caller use 11.  Pre-empt the stale header. */
#define SMEM_SOLVE_HPP
void SMEM_Solve(AllData *all_data);
void SMEM_Smooth(AllData *all_data, hypre_CSRMatrix *A, HYPRE_Real *f, HYPRE_Real *u, HYPRE_Real *y,
                 HYPRE_Real *r, int num_sweeps, int level, int cycle_phase, int ns, int ne);
#endif
