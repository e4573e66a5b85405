// b3d_radix.cu -- hand-written stable LSD radix sorts, 8 bits per pass.
//
// radix_sort_keys (the hot path: voxel keys, search-grid cell keys, query chunk keys): SEGMENTED, PACKED, ONE KERNEL PER PASS.
//  * segmented: the clouds of a batch lie back to back and the cloud id sits in the key's top bits, so those bits are
//    already in order -- every cloud is sorted on its own low bits only (config 2: 29 instead of 36 bits, one pass less),
//    tiles never straddle clouds, and a cloud's result cannot depend on the rest of the batch.
//  * packed: the value of an element is its input position, which fits the 64-bit word next to the low key bits
//    (key << index_bits | index). One 8-byte stream per pass instead of key + value (12 bytes); the first pass packs while
//    it reads the caller's keys, the last pass unpacks into (sorted keys, order).
//  * one kernel per pass ("onesweep"): the digit histograms of ALL passes come from one read of the keys
//    (rs_hist_all_kernel); inside the scatter kernel a tile ranks its keys (match_any on top of per-warp digit counters),
//    publishes its 256 digit counts and finds its base offsets by decoupled look-back over the preceding tiles of its
//    cloud, then re-orders the tile in shared memory and writes every digit's run with consecutive threads.
// radix_sort_pairs (kept for key ranges that do not pack, > 64 bits of key + index): the round-1 three-kernel pass.
#include "b3d_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace b3d {
namespace {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                      // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;    // 4096 keys per block
constexpr int kRsWarpTile = 32 * kRsItems;        // 512 consecutive keys per warp
constexpr int kRsMaxPasses = 8;

// ---------------------------------------------------------------------------------------------------------------------
// segmented packed onesweep
// ---------------------------------------------------------------------------------------------------------------------
// Block shape of the segmented onesweep sort: kOsThreads x kOsItems keys per tile (measured on config 2, ms of sort per 64-pair
// step: 256 x 16: 2.98, 512 x 8: 3.34, 1024 x 4: 4.69 -- larger blocks wait longer at their barriers).
#ifndef B3D_OS_THREADS
#define B3D_OS_THREADS 256
#endif
#ifndef B3D_OS_ITEMS
#define B3D_OS_ITEMS 16
#endif
constexpr int kOsThreads = B3D_OS_THREADS;
constexpr int kOsWarps = kOsThreads / 32;
constexpr int kOsItems = B3D_OS_ITEMS;
constexpr int kOsTile = kOsThreads * kOsItems;
constexpr int kOsWarpTile = 32 * kOsItems;
static_assert(kOsThreads >= 256 && kOsTile % kRsThreads == 0, "one thread per digit; the histogram kernel walks a tile with kRsThreads threads");

struct RsSegView {
    const int32_t* seg_off;     // [B + 1] element offsets
    const int32_t* tile_first;  // [B + 1] first tile of every segment
    int B;
};

__device__ __forceinline__ int rs_find_segment(const RsSegView& sv, int tile) {
    int lo = 0, hi = sv.B;  // last b with tile_first[b] <= tile
    while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (sv.tile_first[m] <= tile) lo = m; else hi = m;
    }
    // empty segments share their first tile with the next one: move to the last segment starting here
    while (lo + 1 < sv.B && sv.tile_first[lo + 1] <= tile) ++lo;
    return lo;
}

// digit histograms of every pass from one read of the keys. Block b owns a contiguous range of tiles and flushes its
// shared-memory table whenever the segment changes. ghist: [B][passes][256]
__global__ void __launch_bounds__(kRsThreads) rs_hist_all_kernel(const uint64_t* __restrict__ keys, RsSegView sv, int n_tiles, int tiles_per_block,
                                                                 int passes, uint64_t low_mask, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[kRsMaxPasses][256];
    const int t0 = blockIdx.x * tiles_per_block, t1 = min(n_tiles, t0 + tiles_per_block);
    if (t0 >= t1) return;
    for (int p = 0; p < passes; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    int seg = rs_find_segment(sv, t0);
    for (int t = t0; t < t1; ++t) {
        int s = seg;
        while (s + 1 < sv.B && sv.tile_first[s + 1] <= t) ++s;
        if (s != seg) {
            __syncthreads();
            for (int p = 0; p < passes; ++p) {
                const uint32_t c = h[p][threadIdx.x];
                if (c) atomicAdd(&ghist[((int64_t)seg * passes + p) * 256 + threadIdx.x], c);
                h[p][threadIdx.x] = 0;
            }
            __syncthreads();
            seg = s;
        }
        const int64_t base = (int64_t)sv.seg_off[seg] + (int64_t)(t - sv.tile_first[seg]) * kOsTile;
        const int64_t end = sv.seg_off[seg + 1];
#pragma unroll 4
        for (int k = 0; k < kOsTile / kRsThreads; ++k) {
            const int64_t i = base + k * kRsThreads + threadIdx.x;
            if (i < end) {
                const uint64_t key = __ldg(keys + i) & low_mask;  // the segment bits above low_bits are not sorted
                for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(uint32_t)(key >> (8 * p)) & 255u], 1u);
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p) {
        const uint32_t c = h[p][threadIdx.x];
        if (c) atomicAdd(&ghist[((int64_t)seg * passes + p) * 256 + threadIdx.x], c);
    }
}

// ghist -> exclusive digit bases inside every segment (+ the segment's own offset). One block per (segment, pass).
__global__ void __launch_bounds__(256) rs_bases_kernel(uint32_t* __restrict__ ghist, const int32_t* __restrict__ seg_off, int passes) {
    __shared__ uint32_t s_w[8];
    const int seg = blockIdx.x / passes;
    uint32_t* h = ghist + (int64_t)blockIdx.x * 256;
    const uint32_t c = h[threadIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t wb = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (w < warp) wb += s_w[w];
    h[threadIdx.x] = (uint32_t)seg_off[seg] + wb + incl - c;
}

constexpr uint32_t kLbAgg = 1u << 30, kLbPrefix = 2u << 30, kLbMask = (1u << 30) - 1u;

// One pass. FIRST: reads the caller's keys and packs (key_low << ib | position); LAST: unpacks into keys_out / order_out.
// status: [n_tiles][256] look-back words of this pass (zeroed), ticket: this pass's tile counter (zeroed).
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kOsThreads) rs_onesweep_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t* __restrict__ order_out,
                                                                 RsSegView sv, int low_bits, int ib, int pass, int passes,
                                                                 const uint32_t* __restrict__ gbase, uint32_t* __restrict__ status,
                                                                 unsigned int* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(rs_smem);
    __shared__ uint32_t warp_cnt[kOsWarps][256];
    __shared__ uint32_t digit_base[256];   // global position of the tile's first key of each digit
    __shared__ uint32_t digit_start[256];  // tile-local position of the same
    __shared__ uint32_t s_wsum[kOsWarps];
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int k = threadIdx.x; k < kOsWarps * 256; k += kOsThreads) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const int tile = (int)s_tile;
    const int seg = rs_find_segment(sv, tile);
    const int tile_in_seg = tile - sv.tile_first[seg];
    const int64_t tbase = (int64_t)sv.seg_off[seg] + (int64_t)tile_in_seg * kOsTile;
    const int64_t seg_end = sv.seg_off[seg + 1];
    const int64_t wbase = tbase + (int64_t)warp * kOsWarpTile;
    const int tile_count = (int)min((int64_t)kOsTile, seg_end - tbase);
    const int shift = ib + 8 * pass;
    const uint64_t low_mask = low_bits >= 64 ? ~0ull : ((1ull << low_bits) - 1ull);
    uint64_t key[kOsItems];
    uint32_t rank[kOsItems];
    // all loads first (16 independent requests in flight per thread), then the ranking
#pragma unroll
    for (int s = 0; s < kOsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        uint64_t e = ~0ull;
        if (i < seg_end) {
            e = __ldg(in + i);
            if (FIRST) e = ((e & low_mask) << ib) | (uint64_t)i;
        }
        key[s] = e;
    }
    // rank 32 consecutive keys per step inside the warp (stable: steps in order, lanes in order)
#pragma unroll
    for (int s = 0; s < kOsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        const bool valid = i < seg_end;
        const uint32_t d = valid ? ((uint32_t)(key[s] >> shift) & 255u) : 256u;  // 256: the out-of-range lanes group together, unused
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const unsigned int lt = peers & ((1u << lane) - 1u);
        uint32_t pre = 0;
        if (valid) pre = warp_cnt[warp][d];
        __syncwarp();
        if (valid && lt == 0) warp_cnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[s] = pre + __popc(lt);
    }
    __syncthreads();
    // per-digit exclusive scan over the tile's warps, digit totals -> tile-local digit starts; look-back -> global bases
    // (thread d < 256 owns digit d)
    {
        const int d = threadIdx.x;
        const bool dig = d < 256;
        uint32_t acc = 0, incl = 0;
        if (dig) {
#pragma unroll
            for (int w = 0; w < kOsWarps; ++w) {
                const uint32_t c = warp_cnt[w][d];
                warp_cnt[w][d] = acc;
                acc += c;
            }
            // publish this tile's count of digit d, then sum the counts of the preceding tiles of the segment
            volatile uint32_t* st = status;
            uint32_t excl = 0;
            if (tile_in_seg == 0) {
                st[(int64_t)tile * 256 + d] = kLbPrefix | acc;
            } else {
                st[(int64_t)tile * 256 + d] = kLbAgg | acc;
                int look = tile - 1;
                const int first = tile - tile_in_seg;
                uint32_t spins = 0;
                while (true) {
                    const uint32_t v = st[(int64_t)look * 256 + d];
                    if (v == 0u) {  // predecessor not published yet (it holds an earlier ticket, so it is running)
                        if (++spins > (1u << 22)) __trap();  // never in a correct run: fail loudly instead of hanging the device
                        continue;
                    }
                    excl += v & kLbMask;
                    if ((v & kLbPrefix) || look == first) break;
                    --look;
                }
                st[(int64_t)tile * 256 + d] = kLbPrefix | (excl + acc);
            }
            digit_base[d] = gbase[((int64_t)seg * passes + pass) * 256 + d] + excl;
            incl = acc;  // block-wide exclusive scan of the digit totals (the first eight warps hold the digits)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
        }
        __syncthreads();
        if (dig) {
            uint32_t wb = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (w < warp) wb += s_wsum[w];
            digit_start[d] = wb + incl - acc;
        }
    }
    __syncthreads();
    // re-order the tile by digit in shared memory
#pragma unroll
    for (int s = 0; s < kOsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        if (i < seg_end) {
            const uint32_t d = (uint32_t)(key[s] >> shift) & 255u;
            skeys[digit_start[d] + warp_cnt[warp][d] + rank[s]] = key[s];
        }
    }
    __syncthreads();
    // coalesced writes: consecutive threads own consecutive positions of a digit's run
    const uint64_t idx_mask = (1ull << ib) - 1ull;
    const uint64_t high = (uint64_t)seg << low_bits;
    for (int k = threadIdx.x; k < tile_count; k += kOsThreads) {
        const uint64_t kk = skeys[k];
        const uint32_t d = (uint32_t)(kk >> shift) & 255u;
        const uint32_t pos = digit_base[d] + ((uint32_t)k - digit_start[d]);
        if (LAST) {
            out[pos] = high | (kk >> ib);
            order_out[pos] = (uint32_t)(kk & idx_mask);
        } else {
            out[pos] = kk;
        }
    }
}

// degenerate case (no key bit to sort): keys pass through, order = identity
__global__ void __launch_bounds__(256) rs_identity_kernel(const uint64_t* __restrict__ in, int64_t n, uint64_t* __restrict__ out, uint32_t* __restrict__ order) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        out[i] = in[i];
        order[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256) rs_iota_kernel(uint32_t* __restrict__ v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------------------------------------------
// (key, value) pairs, three kernels per pass (fallback for keys that do not pack)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist,
                                                             int n_tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        const int64_t i = base + k * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// single-pass chained exclusive scan (decoupled look-back) over a uint32 array, in place
constexpr int kSsItems = 8;
constexpr int kSsTile = kRsThreads * kSsItems;
__global__ void __launch_bounds__(kRsThreads) rs_scan_kernel(uint32_t* __restrict__ data, int64_t n, unsigned long long* __restrict__ status,
                                                             unsigned int* __restrict__ ticket) {
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_warp[kRsWarps];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * kSsTile + (int64_t)threadIdx.x * kSsItems;
    uint32_t v[kSsItems];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kSsItems; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0u;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
        const uint32_t t = s_warp[w];
        if (w < warp) warp_base += t;
        tile_total += t;
    }
    if (threadIdx.x == 0) {
        const unsigned long long kAgg = 1ull << 62, kPre = 2ull << 62, kMask = (1ull << 62) - 1;
        unsigned long long prefix = 0;
        volatile unsigned long long* st = status;
        if (tile == 0) {
            st[0] = kPre | tile_total;
        } else {
            st[tile] = kAgg | tile_total;
            __threadfence();
            long long look = (long long)tile - 1;
            while (true) {
                const unsigned long long s = st[look];
                if (s == 0) continue;
                prefix += s & kMask;
                if (s & kPre) break;
                --look;
            }
            st[tile] = kPre | (prefix + tile_total);
        }
        s_prefix = prefix;
    }
    __syncthreads();
    uint32_t run = (uint32_t)s_prefix + warp_base + incl - sum;
#pragma unroll
    for (int k = 0; k < kSsItems; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

// dynamic shared memory: the tile's pairs re-ordered by digit (keys then values)
__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                                const uint32_t* __restrict__ hist, int n_tiles) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(rs_smem);
    uint32_t* svals = reinterpret_cast<uint32_t*>(rs_smem + (size_t)kRsTile * sizeof(uint64_t));
    __shared__ uint32_t warp_cnt[kRsWarps][256];
    __shared__ uint32_t digit_base[256];   // global position of the tile's first key of each digit
    __shared__ uint32_t digit_start[256];  // tile-local position of the same
    __shared__ uint32_t s_wsum[kRsWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < kRsWarps * 256; k += kRsThreads) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const int64_t tbase = (int64_t)blockIdx.x * kRsTile;
    const int64_t wbase = tbase + (int64_t)warp * kRsWarpTile;
    const int tile_count = (int)min((int64_t)kRsTile, n - tbase);
    uint64_t key[kRsItems];
    uint32_t rank[kRsItems];
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        key[s] = i < n ? __ldg(keys_in + i) : ~0ull;
    }
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((uint32_t)(key[s] >> shift) & 255u) : 256u;
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const unsigned int lt = peers & ((1u << lane) - 1u);
        uint32_t pre = 0;
        if (valid) pre = warp_cnt[warp][d];
        __syncwarp();
        if (valid && lt == 0) warp_cnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[s] = pre + __popc(lt);
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = acc;
            acc += c;
        }
        uint32_t incl = acc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wb = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w)
            if (w < warp) wb += s_wsum[w];
        digit_start[d] = wb + incl - acc;
        digit_base[d] = hist[(int64_t)d * n_tiles + blockIdx.x];
    }
    __syncthreads();
    uint32_t val[kRsItems];
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        val[s] = i < n ? __ldg(vals_in + i) : 0u;
    }
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[s] >> shift) & 255u;
            const uint32_t lp = digit_start[d] + warp_cnt[warp][d] + rank[s];
            skeys[lp] = key[s];
            svals[lp] = val[s];
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < tile_count; k += kRsThreads) {
        const uint64_t kk = skeys[k];
        const uint32_t d = (uint32_t)(kk >> shift) & 255u;
        const uint32_t pos = digit_base[d] + ((uint32_t)k - digit_start[d]);
        keys_out[pos] = kk;
        vals_out[pos] = svals[k];
    }
}

int bits_needed(uint64_t count) {  // smallest b with 2^b >= count
    int b = 0;
    while (b < 63 && (1ull << b) < count) ++b;
    return b;
}

}  // namespace

// Sorts n pairs by key bits [0, end_bit). The result is in (keys_a, vals_a) when *result_in_a, else in (keys_b, vals_b);
// both buffer pairs must hold n elements (ping-pong).
int radix_sort_pairs(b3d_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int64_t n, int end_bit, bool* result_in_a) {
    *result_in_a = true;
    if (n <= 1 || end_bit <= 0) return B3D_OK;
    if (n >= (int64_t)1 << 32) return set_error(B3D_E_RANGE, "radix_sort_pairs: more than 2^32-1 elements");
    const int n_tiles = (int)((n + kRsTile - 1) / kRsTile);
    const int64_t hist_len = (int64_t)256 * n_tiles;
    const int64_t scan_tiles = (hist_len + kSsTile - 1) / kSsTile;
    DevBuf<uint32_t> hist;
    DevBuf<unsigned long long> status;
    DevBuf<unsigned int> ticket;
    B3D_TRY(hist.alloc(ctx, (size_t)hist_len));
    B3D_TRY(status.alloc(ctx, (size_t)scan_tiles));
    B3D_TRY(ticket.alloc(ctx, 1));
    const size_t kScatterSmem = (size_t)kRsTile * (sizeof(uint64_t) + sizeof(uint32_t));
    B3D_CUDA(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
    const int passes = (end_bit + 7) / 8;
    uint64_t* kin = keys_a;
    uint32_t* vin = vals_a;
    uint64_t* kout = keys_b;
    uint32_t* vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        B3D_LAUNCH(ctx, rs_hist_kernel, n_tiles, kRsThreads, 0, kin, n, shift, hist.p, n_tiles);
        B3D_CUDA(cudaMemsetAsync(status.p, 0, (size_t)scan_tiles * sizeof(unsigned long long), ctx->stream));
        B3D_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned int), ctx->stream));
        B3D_LAUNCH(ctx, rs_scan_kernel, (unsigned int)scan_tiles, kRsThreads, 0, hist.p, hist_len, status.p, ticket.p);
        B3D_LAUNCH(ctx, rs_scatter_kernel, n_tiles, kRsThreads, kScatterSmem, kin, vin, kout, vout, n, shift, hist.p, n_tiles);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    *result_in_a = (kin == keys_a);
    return B3D_OK;
}

// Segmented stable sort of composite keys (segment id << low_bits | low key): every segment [seg_off[b], seg_off[b+1]) is
// sorted on its low_bits low key bits; the bits above low_bits of a key must equal its segment index b. Outputs: the sorted
// keys and the input position of every sorted element (ascending inside equal keys). keys_in is left untouched.
int radix_sort_keys(b3d_ctx* ctx, const uint64_t* keys_in, int64_t n, int low_bits, const std::vector<int32_t>& seg_off_h, const int32_t* seg_off_d,
                    uint64_t* keys_out, uint32_t* order_out) {
    if (n <= 0) return B3D_OK;
    const int B = (int)seg_off_h.size() - 1;
    const int ib = std::max(1, bits_needed((uint64_t)n));
    int64_t longest = 0;
    for (int b = 0; b < B; ++b) longest = std::max<int64_t>(longest, seg_off_h[b + 1] - seg_off_h[b]);
    static const bool force_pairs = getenv("B3D_SORT_PAIRS") != nullptr;
    if (low_bits + ib > 64 || longest >= (int64_t)kLbMask || force_pairs) {
        // does not pack: whole-key pair sort (cloud bits included) through the three-kernel passes
        DevBuf<uint64_t> kb;
        DevBuf<uint32_t> vb;
        B3D_TRY(kb.alloc(ctx, (size_t)n));
        B3D_TRY(vb.alloc(ctx, (size_t)n));
        B3D_CUDA(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        B3D_LAUNCH(ctx, rs_iota_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, order_out, n);
        bool in_a = true;
        B3D_TRY(radix_sort_pairs(ctx, keys_out, order_out, kb.p, vb.p, n, low_bits + bits_needed((uint64_t)std::max(B, 1)), &in_a));
        if (!in_a) {
            B3D_CUDA(cudaMemcpyAsync(keys_out, kb.p, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
            B3D_CUDA(cudaMemcpyAsync(order_out, vb.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        return B3D_OK;
    }
    const int passes = (low_bits + 7) / 8;
    if (passes == 0) {
        B3D_LAUNCH(ctx, rs_identity_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, keys_in, n, keys_out, order_out);
        return B3D_OK;
    }
    if (passes > kRsMaxPasses) return set_error(B3D_E_RANGE, "radix_sort_keys: %d key bits need more than %d passes", low_bits, kRsMaxPasses);
    // tiles never straddle segments
    std::vector<int32_t> tile_first(B + 1, 0);
    for (int b = 0; b < B; ++b) tile_first[b + 1] = tile_first[b] + (int32_t)(((int64_t)seg_off_h[b + 1] - seg_off_h[b] + kOsTile - 1) / kOsTile);
    const int n_tiles = tile_first[B];
    DevBuf<int32_t> tile_first_d;
    B3D_TRY(tile_first_d.alloc(ctx, (size_t)B + 1));
    B3D_TRY(ctx->upload(tile_first_d.p, tile_first.data(), (size_t)(B + 1) * sizeof(int32_t)));
    RsSegView sv{seg_off_d, tile_first_d.p, B};
    // scratch: histograms / bases [B][passes][256], look-back words [passes][tiles][256], one ticket per pass
    DevBuf<uint32_t> ghist, status;
    DevBuf<unsigned int> tickets;
    const size_t hist_len = (size_t)B * passes * 256;
    const size_t status_len = (size_t)passes * n_tiles * 256;
    B3D_TRY(ghist.alloc(ctx, hist_len));
    B3D_TRY(status.alloc(ctx, status_len));
    B3D_TRY(tickets.alloc(ctx, (size_t)passes));
    B3D_CUDA(cudaMemsetAsync(ghist.p, 0, hist_len * sizeof(uint32_t), ctx->stream));
    B3D_CUDA(cudaMemsetAsync(status.p, 0, status_len * sizeof(uint32_t), ctx->stream));
    B3D_CUDA(cudaMemsetAsync(tickets.p, 0, (size_t)passes * sizeof(unsigned int), ctx->stream));
    {
        const int blocks = std::max(1, std::min(n_tiles, ctx->sm_count * 8));
        const int per = (n_tiles + blocks - 1) / blocks;
        B3D_LAUNCH(ctx, rs_hist_all_kernel, (n_tiles + per - 1) / per, kRsThreads, 0, keys_in, sv, n_tiles, per, passes,
                   low_bits >= 64 ? ~0ull : ((1ull << low_bits) - 1ull), ghist.p);
        ctx->prof_bytes(8 * n);
        B3D_LAUNCH(ctx, rs_bases_kernel, B * passes, 256, 0, ghist.p, seg_off_d, passes);
    }
    DevBuf<uint64_t> buf_a, buf_b;
    if (passes >= 2) B3D_TRY(buf_a.alloc(ctx, (size_t)n));
    if (passes >= 3) B3D_TRY(buf_b.alloc(ctx, (size_t)n));
    const size_t smem = (size_t)kOsTile * sizeof(uint64_t);
    B3D_CUDA(cudaFuncSetAttribute((rs_onesweep_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B3D_CUDA(cudaFuncSetAttribute((rs_onesweep_kernel<true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B3D_CUDA(cudaFuncSetAttribute((rs_onesweep_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B3D_CUDA(cudaFuncSetAttribute((rs_onesweep_kernel<false, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t* in = keys_in;
    for (int p = 0; p < passes; ++p) {
        const bool first = p == 0, last = p == passes - 1;
        uint64_t* out = last ? keys_out : ((p & 1) == 0 ? buf_a.p : buf_b.p);
        uint32_t* st = status.p + (size_t)p * n_tiles * 256;
        unsigned int* tk = tickets.p + p;
        if (first && last) {
            B3D_LAUNCH(ctx, (rs_onesweep_kernel<true, true>), n_tiles, kOsThreads, smem, in, out, order_out, sv, low_bits, ib, p, passes, ghist.p, st, tk);
        } else if (first) {
            B3D_LAUNCH(ctx, (rs_onesweep_kernel<true, false>), n_tiles, kOsThreads, smem, in, out, order_out, sv, low_bits, ib, p, passes, ghist.p, st, tk);
        } else if (last) {
            B3D_LAUNCH(ctx, (rs_onesweep_kernel<false, true>), n_tiles, kOsThreads, smem, in, out, order_out, sv, low_bits, ib, p, passes, ghist.p, st, tk);
        } else {
            B3D_LAUNCH(ctx, (rs_onesweep_kernel<false, false>), n_tiles, kOsThreads, smem, in, out, order_out, sv, low_bits, ib, p, passes, ghist.p, st, tk);
        }
        ctx->prof_bytes((last ? 20 : 16) * n);  // packed words in, packed words (or keys + order) out
        in = out;
    }
    return B3D_OK;
}

}  // namespace b3d
