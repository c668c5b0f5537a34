// Record exchange between the co-resident CTAs of a panel-factorisation launch (lu.cu, qr.cu).
// Every CTA publishes ONE record per column step, all CTAs meet at a barrier, and every CTA reads
// the G records.  Records are double-buffered by column parity: a CTA that already publishes column
// k+1 (other parity) cannot overwrite what a slower CTA still reads for column k, and a buffer
// of a given parity is rewritten only after two barriers.
//
// Two transports (template parameter CL of the panel kernels):
//   CL = 0  any number of co-resident CTAs (cooperative launch): records in ctx->pxch (global memory,
//           read with __ldcg from L2), barrier = panel_grid_barrier below;
//   CL = 1  the whole launch is ONE thread-block cluster (<= 16 CTAs, i.e. short iterates: BASELINE
//           configs 1 and 2, TSQR blocks): every CTA pushes its record into the shared memory of all
//           cluster CTAs (DSMEM stores) and the barrier is the hardware cluster barrier -- a column step
//           costs ~1 us instead of the 4.5-5 us the global exchange costs even with 4-40 CTAs.
//
// panel_grid_barrier: two-level arrival counters instead of cooperative_groups' grid.sync().  With 148
// CTAs the single counter of grid.sync() serialises 148 atomics on one L2 line (~27 cycles each: 2 us);
// here CTA b arrives on leaf (b mod NL), the arrival that completes a leaf arrives on the root, and
// everybody polls the root: ~(G/NL + NL) serialised atomics.  Counters are monotonic within one
// factorisation (zeroed by a memset before its first panel launch; `epoch` = barriers so far + 1).
//
// Measured alternative (round 2, removed): per-CTA epoch flags (st.release of the epoch into the CTA's own flag, warp 0
// of every CTA polling the G packed flags with one strong load per lane and round) instead of the counters -- correct,
// but slower: LU 3.03 vs 2.80 ms and QR 6.81 vs 6.42 ms at 200 704 x 210 (148 pollers x 5 lines per round).
// Measured alternative (round 2, removed): stamped records polled by one thread per record
// instead of the grid barrier -- 206 us instead of 133 us per LU panel at C3 (the 148 x 148
// polling loads and per-thread fences cost more than the barrier's single counter).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace gsi {
namespace cg = cooperative_groups;

constexpr int kPxchClusterMax = 16;          // CTAs of a single-cluster panel launch (non-portable size, opt-in)
constexpr int kPbarStride = 32;              // unsigned ints between counters (one 128-byte line each)
constexpr int kPbarMaxLeaves = 16;

__host__ __device__ inline int pbar_leaves(int G) {
    int nl = 1;
    while (nl * nl < G) ++nl;                // ceil(sqrt(G))
    return nl > kPbarMaxLeaves ? kPbarMaxLeaves : nl;
}

// All threads of every CTA call it; cnt = ctx->pbar (root at [0], leaf i at [(1 + i) * kPbarStride]).
__device__ __forceinline__ void panel_grid_barrier(unsigned int* cnt, unsigned int epoch, int G, int b) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nl = pbar_leaves(G);
        const int leaf = b % nl;
        const unsigned int leaf_size = (unsigned int)((G - leaf + nl - 1) / nl);
        unsigned int* lp = cnt + (1 + leaf) * kPbarStride;
        unsigned int old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(lp) : "memory");
        if (old + 1 == epoch * leaf_size)
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(cnt) : "memory");
        const unsigned int target = epoch * (unsigned int)nl;
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

}  // namespace gsi
