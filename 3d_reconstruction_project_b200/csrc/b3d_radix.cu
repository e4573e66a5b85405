// b3d_radix.cu -- hand-written stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass.
// Per pass: (1) per-tile digit histograms (shared-memory atomics) written bin-major, (2) one single-pass chained
// exclusive scan over the 256 x tiles table, (3) a stable scatter: every warp ranks 32 consecutive keys per step with
// match_any (peers with the same digit) on top of per-warp digit counters, the per-warp counters are scanned across the
// tile's warps, and the pairs go to scanned-histogram + warp base + rank. Only the key bits that can differ are sorted
// (the callers pass end_bit = cell bits + cloud bits). This replaces the CUB device primitive on the hot path.
#include "b3d_common.cuh"

namespace b3d {
namespace {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                      // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;    // 4096 keys per block
constexpr int kRsWarpTile = 32 * kRsItems;        // 512 consecutive keys per warp

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
    // ---- rank 32 consecutive keys per step inside the warp (stable: steps in order, lanes in order)
    // all loads first (16 independent requests in flight per thread), then the ranking
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        key[s] = i < n ? __ldg(keys_in + i) : ~0ull;
    }
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const int64_t i = wbase + s * 32 + lane;
        const bool valid = i < n;
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
    // ---- per-digit exclusive scan over the tile's warps, digit totals -> tile-local digit starts, global bases
    {
        const int d = threadIdx.x;
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = acc;
            acc += c;
        }
        uint32_t incl = acc;  // block-wide exclusive scan of the digit totals
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
    // ---- re-order the tile by digit in shared memory
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
    // ---- coalesced writes: consecutive threads own consecutive positions of a digit's run
    for (int k = threadIdx.x; k < tile_count; k += kRsThreads) {
        const uint64_t kk = skeys[k];
        const uint32_t d = (uint32_t)(kk >> shift) & 255u;
        const uint32_t pos = digit_base[d] + ((uint32_t)k - digit_start[d]);
        keys_out[pos] = kk;
        vals_out[pos] = svals[k];
    }
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

}  // namespace b3d
