// async_team.cuh -- device helpers of the persistent asynchronous kernel (async.cu): the CTA-group ("team") descriptor and
// the group barrier (= the reference's SMEM_LevelBarrier).
#pragma once
#include "ctx.h"
#include "kernels.cuh"

namespace {

constexpr int kABlock = 256;

struct Team {
   int tid, size;          // thread index / thread count within the level's CTA group
   int cta, nctas;         // CTA index / count within the group
   unsigned int *count;
   volatile unsigned int *gen;
   unsigned char *smem;    // dynamic shared memory (hybrid-JGS staging of the HEAVY instantiation)
};

// barrier among the CTAs of one group (the reference's SMEM_LevelBarrier)
__device__ __forceinline__ void group_barrier(const Team &tm)
{
   __syncthreads();
   if (tm.nctas > 1) {
      if (threadIdx.x == 0) {
         __threadfence();
         const unsigned int g = *tm.gen;
         const unsigned int prev = atomicAdd(tm.count, 1u);
         if (prev == (unsigned int)tm.nctas - 1u) {
            atomicExch(tm.count, 0u);
            __threadfence();
            atomicAdd((unsigned int *)tm.gen, 1u);
         } else {
            while (*tm.gen == g) __nanosleep(32);
         }
         __threadfence();
      }
      __syncthreads();
   }
}

}  // namespace
