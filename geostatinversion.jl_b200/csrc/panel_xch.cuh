// Record exchange between the co-resident CTAs of a panel-factorisation launch (lu.cu, qr.cu).
// Every CTA publishes ONE record per column step into ctx->pxch (an area nothing else uses), the
// grid meets at a cooperative-groups grid barrier, and every CTA reads the G records from L2
// (__ldcg).  Records are double-buffered by column parity: a CTA that already publishes column
// k+1 (other parity) cannot overwrite what a slower CTA still reads for column k, and a buffer
// of a given parity is rewritten only after two grid barriers.
//
// Measured alternative (round 2, removed): stamped records polled by one thread per record
// instead of the grid barrier -- 206 us instead of 133 us per LU panel at C3 (the 148 x 148
// polling loads and per-thread fences cost more than the barrier's single counter).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace gsi {
namespace cg = cooperative_groups;
}  // namespace gsi
