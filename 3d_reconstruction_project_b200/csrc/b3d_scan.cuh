// b3d_scan.cuh -- single-pass stream compaction (chained scan with decoupled look-back), hand-written.
// compact(pred, emit, n): for the j-th index i (ascending) with pred(i) true, calls emit(i, j); the total count is
// left in device memory. Used for: run heads of the sorted cell keys, valid-pixel compaction of the RGB-D
// deprojection, kept-index lists of the outlier filters. Order-preserving, deterministic.
#pragma once

#include "b3d_common.cuh"

namespace b3d {

constexpr int kScanBlock = 256;
#ifndef B3D_SCAN_ITEMS
#define B3D_SCAN_ITEMS 16  // per thread; 4 / 8 / 16 / 32: 0.71 / 0.43 / 0.34 / 0.65 ms of compaction per 64-pair step
#endif
constexpr int kScanItems = B3D_SCAN_ITEMS;
constexpr int kScanTile = kScanBlock * kScanItems;
constexpr unsigned long long kScanFlagAgg = 1ull << 62;
constexpr unsigned long long kScanFlagPrefix = 2ull << 62;
constexpr unsigned long long kScanValueMask = (1ull << 62) - 1;

// Items are dealt to the threads STRIPED (item k of thread t is tile_base + k * kScanBlock + t): every load of the predicate and
// every store of the emitter is a coalesced access of 32 consecutive indices / 32 consecutive output slots. (The first version
// gave each thread kScanItems CONSECUTIVE indices: order-preserving for free, but every warp access was strided by kScanItems
// elements and the emitters' stores touched one sector per lane -- the run-head and disparity compactions ran at a fifth of the
// copy bandwidth.) Output order = index order: the (k, warp) ballots are counted and scanned in tile order.
template <typename Pred, typename Emit>
__global__ void __launch_bounds__(kScanBlock) compact_kernel(Pred pred, Emit emit, int64_t n, unsigned long long* __restrict__ status,
                                                             unsigned int* __restrict__ ticket, int64_t* __restrict__ total) {
    constexpr int kWarps = kScanBlock / 32;
    constexpr int kCells = kScanItems * kWarps;  // (k, warp) groups of 32 consecutive indices, in tile order
    constexpr int kPer = kCells / 32;  // group counts per lane of the scanning warp
    static_assert(kCells % 32 == 0, "whole groups per lane");
    __shared__ unsigned int s_tile;
    __shared__ int s_cnt[kCells];  // counts, then exclusive offsets
    __shared__ long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)tile * kScanTile + threadIdx.x;
    unsigned int flags = 0;
    int rank[kScanItems];  // flagged items of the same (k, warp) group before this lane
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)k * kScanBlock;
        const bool f = i < n && pred(i);
        const unsigned int b = __ballot_sync(0xffffffffu, f);
        flags |= f ? 1u << k : 0u;
        rank[k] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_cnt[k * kWarps + warp] = __popc(b);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the group counts (lane l holds groups kPer * l ..), the tile's total, the look-back
        int c[kPer];
        int mine = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            c[j] = s_cnt[kPer * lane + j];
            mine += c[j];
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int tile_total = __shfl_sync(0xffffffffu, incl, 31);
        int run = incl - mine;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            s_cnt[kPer * lane + j] = run;
            run += c[j];
        }
        if (lane == 0) {
            long long prefix = 0;
            volatile unsigned long long* st = status;
            if (tile == 0) {
                st[0] = kScanFlagPrefix | (unsigned long long)tile_total;
            } else {
                st[tile] = kScanFlagAgg | (unsigned long long)tile_total;
                __threadfence();
                long long look = (long long)tile - 1;
                while (true) {
                    unsigned long long v = st[look];
                    if (v == 0) continue;  // predecessor not published yet
                    prefix += (long long)(v & kScanValueMask);
                    if (v & kScanFlagPrefix) break;
                    --look;
                }
                st[tile] = kScanFlagPrefix | (unsigned long long)(prefix + tile_total);
            }
            s_prefix = prefix;
            if ((int64_t)(tile + 1) * kScanTile >= n) *total = prefix + tile_total;
        }
    }
    __syncthreads();
    const int64_t tile_slot = s_prefix;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (flags & (1u << k)) emit(base + (int64_t)k * kScanBlock, tile_slot + s_cnt[k * kWarps + warp] + rank[k]);
    }
}

struct ScanScratch {
    DevBuf<unsigned long long> status;
    DevBuf<unsigned int> ticket;
};

// total_d: device int64 receiving the count (also written when n == 0)
template <typename Pred, typename Emit>
int compact(b3d_ctx* ctx, Pred pred, Emit emit, int64_t n, int64_t* total_d) {
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 0) {
        B3D_CUDA(cudaMemsetAsync(total_d, 0, sizeof(int64_t), ctx->stream));
        return B3D_OK;
    }
    ScanScratch sc;
    B3D_TRY(sc.status.alloc(ctx, (size_t)tiles));
    B3D_TRY(sc.ticket.alloc(ctx, 1));
    B3D_CUDA(cudaMemsetAsync(sc.status.p, 0, (size_t)tiles * sizeof(unsigned long long), ctx->stream));
    B3D_CUDA(cudaMemsetAsync(sc.ticket.p, 0, sizeof(unsigned int), ctx->stream));
    if (ctx->profiling) ctx->prof_begin("compact_kernel");
    compact_kernel<Pred, Emit><<<(unsigned int)tiles, kScanBlock, 0, ctx->stream>>>(pred, emit, n, sc.status.p, sc.ticket.p, total_d);
    ctx->launches += 1;
    if (ctx->profiling) ctx->prof_end();
    B3D_CUDA(cudaGetLastError());
    return B3D_OK;
}

}  // namespace b3d
