// b3d_scan.cuh -- single-pass stream compaction (chained scan with decoupled look-back), hand-written.
// compact(pred, emit, n): for the j-th index i (ascending) with pred(i) true, calls emit(i, j); the total count is
// left in device memory. Used for: run heads of the sorted cell keys, valid-pixel compaction of the RGB-D
// deprojection, kept-index lists of the outlier filters. Order-preserving, deterministic.
#pragma once

#include "b3d_common.cuh"

namespace b3d {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;
constexpr unsigned long long kScanFlagAgg = 1ull << 62;
constexpr unsigned long long kScanFlagPrefix = 2ull << 62;
constexpr unsigned long long kScanValueMask = (1ull << 62) - 1;

template <typename Pred, typename Emit>
__global__ void __launch_bounds__(kScanBlock) compact_kernel(Pred pred, Emit emit, int64_t n, unsigned long long* __restrict__ status,
                                                             unsigned int* __restrict__ ticket, int64_t* __restrict__ total) {
    __shared__ unsigned int s_tile;
    __shared__ int s_warp[kScanBlock / 32];
    __shared__ long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
    // each thread owns kScanItems consecutive indices (keeps the output order = index order)
    unsigned int flags = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < n && pred(i)) {
            flags |= 1u << k;
            ++cnt;
        }
    }
    // block exclusive scan of cnt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanBlock / 32; ++w) {
        const int v = s_warp[w];
        if (w < warp) warp_base += v;
        tile_total += v;
    }
    const int excl = warp_base + incl - cnt;
    if (threadIdx.x == 0) {
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
    __syncthreads();
    int64_t slot = s_prefix + excl;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (flags & (1u << k)) {
            emit(base + k, slot);
            ++slot;
        }
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
