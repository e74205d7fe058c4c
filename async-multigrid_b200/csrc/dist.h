// dist.h -- state of the row-partitioned (multi-GPU) path, shared by dist.cu and dist_async.cu (not part of the C ABI)
#pragma once
#include "ctx.h"
#include <vector>
#ifdef AMG_HAVE_NCCL
#include <nccl.h>
#endif

struct DistAsync;

struct DistLevel {
   int n_global = 0, row_start = 0, n_owned = 0, halo_lo = 0, halo_hi = 0, distributed = 0, send_lo = 0, send_hi = 0;
   bool set = false;
   std::vector<int> all_owned;   // n_owned of every rank (all-gather counts)
   int n_ext() const { return distributed ? halo_lo + n_owned + halo_hi : n_global; }
   int off() const { return distributed ? halo_lo : 0; }   // position of the first owned entry
};

struct DistState {
   int rank = 0, nranks = 1;
#ifdef AMG_HAVE_NCCL
   ncclComm_t comm = nullptr;
#endif
   std::vector<DistLevel> lv;
   std::vector<double *> ws, r, e;   // level layout (ws = w/d, or 1/l1 for the L1-Jacobi smoother)
   std::vector<double *> t, w;       // level layout: AFACx scratch (coarse-grid correction / its prolongation, fine residual)
   double *u = nullptr, *f = nullptr;   // u: level-0 layout; f: owned rows
   double *ecyc = nullptr, *dacc = nullptr;   // owned rows: cycle output and the accelerated increment (DMEM_ChebyUpdate)
   double *t0 = nullptr, *v0 = nullptr;  // level-0 layout: scratch of the factorised level-0 transfers (factor_level0)
   // halo exchange on its own stream, overlapped with the interior launch units of the SpMV that needs it
   cudaStream_t comm_stream = nullptr;
   cudaEvent_t ev_x = nullptr, ev_h = nullptr;
   double *partials3 = nullptr;          // 3 x npartials: interior / low boundary / high boundary launches
   bool overlap = true;
   bool ready = false;
   // one cycle + residual + norm captured as a CUDA graph, NCCL operations and the communication stream's fork / join
   // included, replayed per cycle.  Validated with a single-rank communicator only: with two ranks the replayed graph
   // DEADLOCKS on the B200 box (round 2, profiles/r2_call4_2gpu.log: both 2-GPU tests and `bench.py --gpus 2` hung until
   // their timeouts, while the per-operation path of the same build converged in 38 cycles) -- the captured ncclSend /
   // ncclRecv pairs of the two ranks never meet.  So the graph is the default for ONE rank and off otherwise;
   // AMGB_DIST_GRAPH=1 / 0 force it.
   bool use_graph = false;
   cudaGraphExec_t graph_exec = nullptr;
   bool graph_warm = false;              // one cycle has run with per-operation launches (NCCL's lazy connections exist)
   long long graph_kernels = 0, graph_halo_bytes = 0, graph_collectives = 0;
   // asynchronous fine-grid smoother across GPUs (DMEM_AsyncSmooth): the neighbours' level-0 solution vectors mapped
   // through CUDA IPC, and where in them this rank's boundary entries belong (their ghost slots)
   double *nbr_lo = nullptr, *nbr_hi = nullptr;   // rank-1 / rank+1
   long nbr_lo_off = 0;                            // first ghost_hi entry of rank-1 ( = its halo_lo + n_owned )
   double *sm_scratch = nullptr;
   long long halo_bytes = 0, collectives = 0;
   DistAsync *da = nullptr;              // row-partitioned asynchronous solve (dist_async.cu)
};

#ifdef AMG_HAVE_NCCL
#define NCCL_OK(c, call)                                                                                    \
   do {                                                                                                     \
      ncclResult_t r__ = (call);                                                                            \
      if (r__ != ncclSuccess)                                                                               \
         return amgb_fail((c), AMGB_ENCCL, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, ncclGetErrorString(r__)); \
   } while (0)
// ghosts of v (level layout) <- neighbours' boundary entries (st: nullptr = the context's stream)
int dist_halo(amgb_ctx *c, int l, double *v, cudaStream_t st = nullptr);
// r_0 = f - A_0 u on the owned rows, d_scalars[0] = global ||r||^2 (collective)
int dist_residual(amgb_ctx *c);
#endif
void amgb_dist_async_teardown(amgb_ctx *c);
