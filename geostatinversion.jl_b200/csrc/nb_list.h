// Instantiated widths of the TALL layout: NB = number of 8-column DMMA blocks a warp
// accumulates (lp = 8*NB columns, 2*NB FP64 accumulators per lane).
#pragma once
#define GSI_NB_LIST(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(20) X(24) X(27) X(28) X(32)
