/* TEST INFRASTRUCTURE.  Compiles the reference's src/SMEM_Solve.cpp unmodified, with its
 * residual-history printf (src/SMEM_Solve.cpp:95-103,232-239) routed to ref_hook_printf so the
 * history can be read at full double precision.  All standard and reference headers are
 * included first so the macro touches only the body of SMEM_Solve.cpp. */
#include "ref_prelude.hpp"
#include "Misc.hpp"
#include "SEQ_MatVec.hpp"
#include "SEQ_AMG.hpp"
#include "SMEM_MatVec.hpp"
#include "SMEM_Sync_AMG.hpp"
#include "SMEM_Async_AMG.hpp"
#include "SMEM_Smooth.hpp"
#include "SEQ_Smooth.hpp"
extern "C" int ref_hook_printf(const char *fmt, ...);
#define printf ref_hook_printf
#include "SMEM_Solve.cpp"
