// ab200_kernels_lane.cu -- the one-cell-per-LANE build of the evolution kernel (sources under lane/), for frontiers that are
// far larger than the device can hold at one warp per cell. A warp then carries 32 cells through the same instruction
// stream, which costs an order of magnitude fewer issued instructions per cell than the warp-per-cell build but gives every
// cell the latency of the slowest of its 32 neighbours and keeps ~10^5 scratch arenas in HBM; it only pays above a few tens
// of thousands of cells (measurements in profiles/). 128 registers, 16 resident warps per SM, 8-way merges, evict-first
// hints on the accumulator traffic of products.
#define AB_KW 1
#define AB_STREAMING_ACC 1
#define AB_VARIANT lane
#define AB_MINBLOCKS 8
#define AB_MAXSTREAMS 8
#define AB_MULSTREAMS 8
#define AB_EVOLVE_BLOCK 64
#include <cuda_runtime.h>
#include <stdio.h>
#include "lane/ab200_evolve.cuh"

namespace ab200 { inline namespace AB_VARIANT {
#include "lane/ab200_evolve_kernel.inc"
} }
