// b3d_chunks.cu -- compact warp chunks of query points (used by the staged searches of the ICP pass and of the normals).
#include "b3d_common.cuh"
#include "b3d_scan.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace b3d {
namespace {

// Hilbert index of (x, y, z) with `bits` bits per axis (Skilling's transpose form, then bit interleave): consecutive
// indices are face-adjacent cells, so runs of consecutive points are spatially tighter than along a Morton curve.
__device__ __forceinline__ unsigned long long hilbert3(uint32_t x, uint32_t y, uint32_t z, int bits) {
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (morton_spread3(X[0]) << 2) | (morton_spread3(X[1]) << 1) | morton_spread3(X[2]);
}

// Morton key of the (transformed) query on a quarter-cell lattice of its cloud's search grid: consecutive keys are
// spatially compact, so the 32 queries of a warp fit a small box (and stay compact under rigid updates).
__global__ void __launch_bounds__(256) chunk_key_kernel(const double* __restrict__ pts, const int32_t* __restrict__ off, const double* __restrict__ transforms,
                                                        int transform_stride, const Lattice* __restrict__ lat, int shift, int hilbert_bits, int sub,
                                                        uint64_t* __restrict__ keys) {
    const int cloud = blockIdx.y;
    const Lattice L = lat[cloud];
    double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (transforms != nullptr)
        for (int k = 0; k < 12; ++k) T[k] = transforms[(int64_t)cloud * transform_stride + k];
    const int32_t s0 = off[cloud], s1 = off[cloud + 1];
    const double q = L.cell / (double)sub;
    const double margin = (double)sub;  // one cell below the lattice origin
    for (int32_t i = s0 + blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += gridDim.x * blockDim.x) {
        const double x = pts[3 * (int64_t)i], y = pts[3 * (int64_t)i + 1], z = pts[3 * (int64_t)i + 2];
        const double px = T[0] * x + T[1] * y + T[2] * z + T[3];
        const double py = T[4] * x + T[5] * y + T[6] * z + T[7];
        const double pz = T[8] * x + T[9] * y + T[10] * z + T[11];
        const double hi = 2097151.0;  // 2^21 - 1
        // one cell of margin below the lattice origin; everything farther out clamps to the border
        const double ux = fmin(fmax(floor((px - L.ox) / q) + margin, 0.0), hi), uy = fmin(fmax(floor((py - L.oy) / q) + margin, 0.0), hi),
                     uz = fmin(fmax(floor((pz - L.oz) / q) + margin, 0.0), hi);
        const unsigned long long cap = (1ull << (shift / 3)) - 1ull;  // coordinates beyond the keyed range clamp to the border
        const unsigned long long ix = min((unsigned long long)ux, cap), iy = min((unsigned long long)uy, cap), iz = min((unsigned long long)uz, cap);
        const unsigned long long m = hilbert_bits > 0 ? hilbert3((uint32_t)ix, (uint32_t)iy, (uint32_t)iz, hilbert_bits)
                                                      : (morton_spread3(ix) << 2) | (morton_spread3(iy) << 1) | morton_spread3(iz);
        keys[i] = ((unsigned long long)cloud << shift) | (m & ((1ull << shift) - 1ull));
    }
}

// Gap-based chunking in ONE pass: runs of consecutive sorted points without a jump longer than tau (and without a change of
// cloud); a chunk starts every 32 points of a run. Two chained scans with decoupled look-back inside one kernel: first the index
// of the last run head at or before every point (a max-scan), from it the chunk-start flags, then their prefix count (the
// compaction). Replaces two compaction passes and a binary search per point.
constexpr int kCutBlock = 256;
constexpr int kCutItems = 8;
constexpr int kCutTile = kCutBlock * kCutItems;

__global__ void __launch_bounds__(kCutBlock) cut_chunks_kernel(const uint64_t* __restrict__ keys, const double4* __restrict__ pts, int shift, double tau2,
                                                               int32_t n, unsigned long long* __restrict__ st_head, unsigned long long* __restrict__ st_count,
                                                               unsigned int* __restrict__ ticket, int32_t* __restrict__ chunk_start,
                                                               int64_t* __restrict__ n_chunks) {
    __shared__ unsigned int s_tile;
    __shared__ int s_warp[kCutBlock / 32];
    __shared__ long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int32_t base = (int32_t)tile * kCutTile + (int32_t)threadIdx.x * kCutItems;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr unsigned long long kAgg = 1ull << 62, kPre = 2ull << 62, kMask = (1ull << 62) - 1ull;
    // ---- run heads of this thread's items; the last one (index + 1, 0 = none) ----------------------------------------------
    unsigned int head_bits = 0;
    int last = 0;
    {
        double4 prev = make_double4(0, 0, 0, 0);
        uint64_t prev_cloud = 0;
        if (base > 0 && base - 1 < n) {
            prev = pts[base - 1];
            prev_cloud = keys[base - 1] >> shift;
        }
#pragma unroll
        for (int k = 0; k < kCutItems; ++k) {
            const int32_t i = base + k;
            if (i < n) {
                const double4 a = pts[i];
                const uint64_t cl = keys[i] >> shift;
                const double dx = a.x - prev.x, dy = a.y - prev.y, dz = a.z - prev.z;
                const bool head = i == 0 || cl != prev_cloud || dx * dx + dy * dy + dz * dz > tau2;
                if (head) {
                    head_bits |= 1u << k;
                    last = i + 1;
                }
                prev = a;
                prev_cloud = cl;
            }
        }
    }
    // inclusive max-scan of `last` over the block -> last head at or before the END of every thread's items, within the tile
    int incl = last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = max(incl, v);
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;  // last head (index + 1) in the tile BEFORE this thread's items
    {
        int wmax = 0;
        for (int w = 0; w < warp; ++w) wmax = max(wmax, s_warp[w]);
        const int up = __shfl_up_sync(0xffffffffu, incl, 1);
        before = max(wmax, lane > 0 ? up : 0);
    }
    int tile_last = 0;
    for (int w = 0; w < kCutBlock / 32; ++w) tile_last = max(tile_last, s_warp[w]);
    __syncthreads();
    // look-back 1: last head before the tile
    if (threadIdx.x == 0) {
        volatile unsigned long long* st = st_head;
        long long prefix = 0;
        if (tile == 0) {
            st[0] = kPre | (unsigned long long)tile_last;
        } else {
            st[tile] = kAgg | (unsigned long long)tile_last;
            __threadfence();
            long long look = (long long)tile - 1;
            while (true) {
                const unsigned long long v = st[look];
                if (v == 0) continue;
                prefix = max(prefix, (long long)(v & kMask));
                if ((v & kPre) || prefix > 0) break;  // any head ends the search: earlier tiles cannot hold a later one
                --look;
            }
            st[tile] = kPre | (unsigned long long)max(prefix, (long long)tile_last);
        }
        s_prefix = prefix;
    }
    __syncthreads();
    int run_head = max(before, (int)s_prefix);  // index + 1 of the run head governing this thread's first item (>= 1: point 0 is a head)
    __syncthreads();
    // ---- chunk starts: every 32 points of a run ----------------------------------------------------------------------------
    unsigned int flags = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kCutItems; ++k) {
        const int32_t i = base + k;
        if (i < n) {
            if (head_bits & (1u << k)) run_head = i + 1;
            if (((i - (run_head - 1)) & 31) == 0) {
                flags |= 1u << k;
                ++cnt;
            }
        }
    }
    int cincl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, cincl, o);
        if (lane >= o) cincl += v;
    }
    if (lane == 31) s_warp[warp] = cincl;
    __syncthreads();
    int warp_base = 0, tile_total = 0;
    for (int w = 0; w < kCutBlock / 32; ++w) {
        const int v = s_warp[w];
        if (w < warp) warp_base += v;
        tile_total += v;
    }
    // look-back 2: chunk starts before the tile
    if (threadIdx.x == 0) {
        volatile unsigned long long* st = st_count;
        long long prefix = 0;
        if (tile == 0) {
            st[0] = kPre | (unsigned long long)tile_total;
        } else {
            st[tile] = kAgg | (unsigned long long)tile_total;
            __threadfence();
            long long look = (long long)tile - 1;
            while (true) {
                const unsigned long long v = st[look];
                if (v == 0) continue;
                prefix += (long long)(v & kMask);
                if (v & kPre) break;
                --look;
            }
            st[tile] = kPre | (unsigned long long)(prefix + tile_total);
        }
        s_prefix = prefix;
        if ((int64_t)(tile + 1) * kCutTile >= n) *n_chunks = prefix + tile_total;
    }
    __syncthreads();
    int64_t slot = s_prefix + warp_base + cincl - cnt;
#pragma unroll
    for (int k = 0; k < kCutItems; ++k)
        if (flags & (1u << k)) chunk_start[slot++] = base + k;
}

__global__ void chunk_sentinel_kernel(int32_t* chunk_start, const int64_t* n_chunks, int32_t n) { chunk_start[*n_chunks] = n; }

// chunk_off[b] = first chunk of cloud b (chunk_off[B] = n_chunks)
__global__ void chunk_ranges_kernel(const int32_t* __restrict__ chunk_start, const int64_t* __restrict__ n_chunks_d, const int32_t* __restrict__ off, int B,
                                    int32_t* __restrict__ chunk_off) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    const int32_t key = off[b];
    int lo = 0, hi = (int)*n_chunks_d;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (chunk_start[m] < key) lo = m + 1; else hi = m;
    }
    chunk_off[b] = lo;
}

__global__ void __launch_bounds__(256) chunk_gather_kernel(const double* __restrict__ pts, const uint32_t* __restrict__ order, int32_t n,
                                                           double4* __restrict__ out) {
    for (int32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = order[j];
        out[j] = make_double4(pts[3 * (int64_t)i], pts[3 * (int64_t)i + 1], pts[3 * (int64_t)i + 2], __longlong_as_double((long long)i));
    }
}

}  // namespace

// chunk_start / n_chunks from sorted keys (+ sorted points for the gap rule)
constexpr double kChunkGapQuery = 1.5, kChunkGapGrid = 1.5;

static double chunk_gap_cells(const char* own_env, double dflt) {
    if (const char* e = getenv(own_env)) return atof(e);
    if (const char* e = getenv("B3D_CHUNK_GAP")) return atof(e);
    return dflt;
}

static int cut_chunks(b3d_ctx* ctx, const uint64_t* keys, const double4* pts, int shift, double cell, double tau_cells, int32_t n,
                      int32_t* chunk_start, int64_t* n_chunks_d) {
    // runs of spatially consecutive points (no jump longer than tau_cells cells), then a chunk every 32 points of a run: chunks
    // are full except at the end of a run, and compact because the curve does not jump inside a run
    const double tau = tau_cells * cell;
    const int64_t tiles = ((int64_t)n + kCutTile - 1) / kCutTile;
    if (tiles == 0) {
        B3D_CUDA(cudaMemsetAsync(n_chunks_d, 0, sizeof(int64_t), ctx->stream));
        return B3D_OK;
    }
    DevBuf<unsigned long long> status;  // [2][tiles]: look-back words of the two scans
    DevBuf<unsigned int> ticket;
    B3D_TRY(status.alloc(ctx, (size_t)(2 * tiles)));
    B3D_TRY(ticket.alloc(ctx, 1));
    B3D_CUDA(cudaMemsetAsync(status.p, 0, (size_t)(2 * tiles) * sizeof(unsigned long long), ctx->stream));
    B3D_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(unsigned int), ctx->stream));
    B3D_LAUNCH(ctx, cut_chunks_kernel, (unsigned int)tiles, kCutBlock, 0, keys, pts, shift, tau * tau, n, status.p, status.p + tiles, ticket.p, chunk_start,
               n_chunks_d);
    ctx->prof_bytes((int64_t)n * 40);
    return B3D_OK;
}

int build_query_chunks(b3d_ctx* ctx, const double* pts, const int32_t* off_d, const std::vector<int32_t>& off_h, const SpatialSort& lattices,
                       const double* transforms, int transform_stride, QueryChunks* out) {
    const int B = (int)off_h.size() - 1;
    const int32_t n = off_h[B];
    out->chunk_off_h.assign(B + 1, 0);
    out->n_chunks = 0;
    out->most = 0;
    B3D_TRY(out->pts.alloc(ctx, (size_t)std::max(n, 1)));
    B3D_TRY(out->chunk_start.alloc(ctx, (size_t)n + 1));
    B3D_TRY(out->chunk_off.alloc(ctx, (size_t)B + 1));
    if (n == 0) {
        B3D_CUDA(cudaMemsetAsync(out->chunk_start.p, 0, sizeof(int32_t), ctx->stream));
        B3D_CUDA(cudaMemsetAsync(out->chunk_off.p, 0, (size_t)(B + 1) * sizeof(int32_t), ctx->stream));
        return B3D_OK;
    }
    // Morton bits: 3 x bits(4 * cells per axis + margin), capped at 3 x 21; coordinates beyond that wrap (still correct, less compact)
    int64_t max_axis = 1, longest = 0;
    for (int b = 0; b < B; ++b) {
        const Lattice& L = lattices.lat_h[b];
        max_axis = std::max<int64_t>(max_axis, std::max(std::max(L.nx, L.ny), L.nz));
        longest = std::max<int64_t>(longest, off_h[b + 1] - off_h[b]);
    }
    int axis_bits = 1;
    static const int sub = getenv("B3D_CHUNK_SUB") ? std::max(1, atoi(getenv("B3D_CHUNK_SUB"))) : 4;  // lattice steps per grid cell
    while (axis_bits < 21 && (1ll << axis_bits) < (int64_t)sub * max_axis + 2 * sub) ++axis_bits;
    int bbits = 0;
    while ((1ll << bbits) < B) ++bbits;
    while (3 * axis_bits + bbits > 63) --axis_bits;
    const int shift = 3 * axis_bits;
    DevBuf<uint64_t> k_in, k_out;
    DevBuf<uint32_t> o_out;
    B3D_TRY(k_in.alloc(ctx, n));
    B3D_TRY(k_out.alloc(ctx, n));
    B3D_TRY(o_out.alloc(ctx, n));
    const int kb = (int)std::min<int64_t>((longest + 255) / 256, std::max(1, ctx->sm_count * 16 / B));
    B3D_LAUNCH(ctx, chunk_key_kernel, dim3(std::max(1, kb), B), 256, 0, pts, off_d, transforms, transform_stride, lattices.lat.p, shift,
               getenv("B3D_CHUNK_MORTON") ? 0 : axis_bits, sub, k_in.p);
    (void)bbits;
    B3D_TRY(radix_sort_keys(ctx, k_in.p, n, shift, off_h, off_d, k_out.p, o_out.p));
    k_in.release();
    B3D_LAUNCH(ctx, chunk_gather_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, pts, o_out.p, n, out->pts.p);
    DevBuf<int64_t> n_chunks_d;
    B3D_TRY(n_chunks_d.alloc(ctx, 1));
    // gap of the query chunks (Hilbert order): measured on config 2, see DESIGN.md 3.5
    static const double gap_q = chunk_gap_cells("B3D_CHUNK_GAP_QUERY", kChunkGapQuery);
    B3D_TRY(cut_chunks(ctx, k_out.p, out->pts.p, shift, lattices.lat_h[0].cell, gap_q, n, out->chunk_start.p, n_chunks_d.p));
    B3D_LAUNCH(ctx, chunk_sentinel_kernel, 1, 1, 0, out->chunk_start.p, n_chunks_d.p, n);
    B3D_LAUNCH(ctx, chunk_ranges_kernel, (B + 1 + 127) / 128, 128, 0, out->chunk_start.p, n_chunks_d.p, off_d, B, out->chunk_off.p);
    B3D_TRY(ctx->download(out->chunk_off_h.data(), out->chunk_off.p, (size_t)(B + 1) * sizeof(int32_t)));
    out->n_chunks = out->chunk_off_h[B];
    for (int b = 0; b < B; ++b) out->most = std::max(out->most, out->chunk_off_h[b + 1] - out->chunk_off_h[b]);
    out->q = out->pts.p;
    return B3D_OK;
}

int chunks_from_grid(b3d_ctx* ctx, const Grid<double>& grid, const int32_t* off_d, const std::vector<int32_t>& off_h, QueryChunks* out) {
    const SpatialSort& ss = grid.sort;
    const int B = (int)off_h.size() - 1;
    const int32_t n = off_h[B];
    out->chunk_off_h.assign(B + 1, 0);
    out->n_chunks = 0;
    out->most = 0;
    out->pts.release();
    out->q = grid.pts.p;
    B3D_TRY(out->chunk_start.alloc(ctx, (size_t)n + 1));
    B3D_TRY(out->chunk_off.alloc(ctx, (size_t)B + 1));
    if (n == 0) {
        B3D_CUDA(cudaMemsetAsync(out->chunk_start.p, 0, sizeof(int32_t), ctx->stream));
        B3D_CUDA(cudaMemsetAsync(out->chunk_off.p, 0, (size_t)(B + 1) * sizeof(int32_t), ctx->stream));
        return B3D_OK;
    }
    DevBuf<int64_t> n_chunks_d;
    B3D_TRY(n_chunks_d.alloc(ctx, 1));
    // gap of the chunks cut from a grid's own (Morton) order: 1.0 / 1.25 / 1.5 / 2.0 / 3.0 cells measured 49.7 / 47.9 / 47.5 / 47.8 / 51.5 ms
    // per config-2 step in round 1, 1.0 / 1.5 / 2.5 cells 5.67 / 5.25 / 5.78 ms of normals in round 2
    static const double gap_g = chunk_gap_cells("B3D_CHUNK_GAP_GRID", kChunkGapGrid);
    B3D_TRY(cut_chunks(ctx, ss.keys.p, grid.pts.p, ss.shift, grid.cell, gap_g, n, out->chunk_start.p, n_chunks_d.p));
    B3D_LAUNCH(ctx, chunk_sentinel_kernel, 1, 1, 0, out->chunk_start.p, n_chunks_d.p, n);
    B3D_LAUNCH(ctx, chunk_ranges_kernel, (B + 1 + 127) / 128, 128, 0, out->chunk_start.p, n_chunks_d.p, off_d, B, out->chunk_off.p);
    B3D_TRY(ctx->download(out->chunk_off_h.data(), out->chunk_off.p, (size_t)(B + 1) * sizeof(int32_t)));
    out->n_chunks = out->chunk_off_h[B];
    for (int b = 0; b < B; ++b) out->most = std::max(out->most, out->chunk_off_h[b + 1] - out->chunk_off_h[b]);
    return B3D_OK;
}

}  // namespace b3d
