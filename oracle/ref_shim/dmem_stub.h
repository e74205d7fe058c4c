/* dmem_stub.h -- TEST INFRASTRUCTURE.  What the reference's DMEM translation units (src/DMEM_Mult.cpp, src/DMEM_Misc.cpp)
 * need beyond hypre_stub.h to compile UNMODIFIED for ONE rank: an MPI of exactly one process and a few more hypre accessors.
 * MPI and hypre are un-vendored dependencies absent from this image; only the entry points those two files name are
 * provided, with single-rank meaning (rank 0 of 1, reductions are copies), or declared and defined in ref_driver.cpp. */
#ifndef AMG_REF_DMEM_STUB_H
#define AMG_REF_DMEM_STUB_H
#include <string.h>
#include <omp.h>
namespace mfem {}
typedef int MPI_Request;
typedef int MPI_Status;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_INT 4
#define MPI_DOUBLE 8
#define hypre_MPI_SUM MPI_SUM
#define HYPRE_MPI_REAL MPI_DOUBLE
#define HYPRE_MEMORY_HOST 0
#define HYPRE_MEMORY_SHARED 1
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int hypre_MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline double MPI_Wtime(void) { return omp_get_wtime(); }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
#define MPI_BYTE 1
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }     /* one process: the buffer already holds the root's data */
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
/* never reached on one rank (no neighbour, no other grid): */
static inline int hypre_MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }
static inline int hypre_MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
static inline size_t amg_ref_mpi_size(MPI_Datatype t) { return t == MPI_INT ? sizeof(int) : sizeof(double); }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { memcpy(r, s, n * amg_ref_mpi_size(t)); return 0; }
static inline int hypre_MPI_Allreduce(void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { memcpy(r, s, n * amg_ref_mpi_size(t)); return 0; }
#define hypre_TMemcpy(dst, src, type, count, locdst, locsrc) memcpy((dst), (src), sizeof(type) * (size_t)(count))

/* more of hypre_ParAMGData / hypre_ParCSRMatrix (fields live in hypre_stub.h) */
#define hypre_ParAMGDataNumLevels(d) ((d)->num_levels)
#define hypre_ParAMGDataGridRelaxType(d) ((d)->grid_relax_type)
#define hypre_ParAMGDataAddRelaxWt(d) ((d)->add_rlx_wt)
#define hypre_ParAMGDataSimple(d) ((d)->simple)
#define hypre_ParAMGDataRelaxWeight(d) ((d)->relax_weight)
#define hypre_ParCSRMatrixComm(m) ((m)->comm)
#define hypre_ParCSRMatrixGlobalNumRows(m) ((m)->global_num_rows)
#define hypre_ParCSRMatrixGlobalNumCols(m) ((m)->global_num_rows)
#define hypre_ParCSRMatrixRowStarts(m) ((m)->row_starts)
#define hypre_ParCSRMatrixFirstRowIndex(m) (0)
#define hypre_ParCSRMatrixFirstColDiag(m) (0)
#define hypre_ParCSRMatrixOffd(m) ((m)->offd)
#define hypre_ParCSRMatrixColMapOffd(m) ((m)->col_map_offd)
#define hypre_ParCSRMatrixNumNonzeros(m) ((m)->diag->num_nonzeros)
#define hypre_ParCSRMatrixSetNumNonzeros(m) (0)

HYPRE_Int hypre_CSRMatrixMatvec(HYPRE_Complex alpha, hypre_CSRMatrix *A, hypre_Vector *x, HYPRE_Complex beta, hypre_Vector *y);
HYPRE_Int hypre_ParCSRMatrixMatvecT(HYPRE_Complex alpha, hypre_ParCSRMatrix *A, hypre_ParVector *x, HYPRE_Complex beta, hypre_ParVector *y);
HYPRE_Int hypre_BoomerAMGRelax(hypre_ParCSRMatrix *A, hypre_ParVector *f, HYPRE_Int *cf_marker, HYPRE_Int relax_type, HYPRE_Int relax_points,
                               HYPRE_Real relax_weight, HYPRE_Real omega, HYPRE_Real *l1_norms, hypre_ParVector *u, hypre_ParVector *Vtemp,
                               hypre_ParVector *Ztemp);
HYPRE_Real hypre_ParVectorInnerProd(hypre_ParVector *x, hypre_ParVector *y);
HYPRE_Real hypre_SeqVectorInnerProd(hypre_Vector *x, hypre_Vector *y);
HYPRE_Int hypre_ParVectorScale(HYPRE_Complex alpha, hypre_ParVector *y);
HYPRE_Int hypre_ParVectorAxpy(HYPRE_Complex alpha, hypre_ParVector *x, hypre_ParVector *y);
hypre_ParVector *hypre_ParVectorCreate(MPI_Comm comm, HYPRE_BigInt global_size, HYPRE_BigInt *partitioning);
HYPRE_Int hypre_ParVectorInitialize(hypre_ParVector *v);
HYPRE_Int hypre_ParVectorDestroy(hypre_ParVector *v);
HYPRE_Int hypre_ParVectorSetPartitioningOwner(hypre_ParVector *v, HYPRE_Int owns);
#endif
