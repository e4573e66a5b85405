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

// gap-based chunking: runs of consecutive sorted points without a jump longer than tau; chunks = every 32 points of a run
struct GapPred {
    const double4* pts;
    const uint64_t* keys;
    int shift;
    double tau2;
    __device__ __forceinline__ bool operator()(int64_t i) const {
        if (i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift)) return true;  // first point of a cloud
        const double4 a = pts[i], b = pts[i - 1];
        const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        return dx * dx + dy * dy + dz * dz > tau2;
    }
};
struct RunChunkPred {
    const int32_t* heads;  // sorted run heads
    const int64_t* n_heads;
    __device__ __forceinline__ bool operator()(int64_t i) const {
        int lo = 0, hi = (int)*n_heads;  // last head <= i
        while (hi - lo > 1) {
            const int m = (lo + hi) >> 1;
            if ((int64_t)heads[m] <= i) lo = m; else hi = m;
        }
        return ((i - (int64_t)heads[lo]) & 31) == 0;
    }
};
struct ChunkEmit {
    int32_t* chunk_start;
    __device__ __forceinline__ void operator()(int64_t i, int64_t slot) const { chunk_start[slot] = (int32_t)i; }
};
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
static int cut_chunks(b3d_ctx* ctx, const uint64_t* keys, const double4* pts, int shift, double cell, int32_t n,
                      int32_t* chunk_start, int64_t* n_chunks_d) {
    // runs of spatially consecutive points (no jump longer than 1.5 cells), then a chunk every 32 points of a run: chunks
    // are full except at the end of a run, and compact because the curve does not jump inside a run
    DevBuf<int32_t> heads;
    DevBuf<int64_t> n_heads;
    B3D_TRY(heads.alloc(ctx, (size_t)n));
    B3D_TRY(n_heads.alloc(ctx, 1));
    double tau_cells = 1.5;  // measured on config 2 (1.0 / 1.25 / 1.5 / 2.0 / 3.0 cells: 49.7 / 47.9 / 47.5 / 47.8 / 51.5 ms per step)
    if (const char* e = getenv("B3D_CHUNK_GAP")) tau_cells = atof(e);
    const double tau = tau_cells * cell;
    B3D_TRY(compact(ctx, GapPred{pts, keys, shift, tau * tau}, ChunkEmit{heads.p}, n, n_heads.p));
    return compact(ctx, RunChunkPred{heads.p, n_heads.p}, ChunkEmit{chunk_start}, n, n_chunks_d);
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
    B3D_TRY(cut_chunks(ctx, k_out.p, out->pts.p, shift, lattices.lat_h[0].cell, n, out->chunk_start.p, n_chunks_d.p));
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
    B3D_TRY(cut_chunks(ctx, ss.keys.p, grid.pts.p, ss.shift, grid.cell, n, out->chunk_start.p, n_chunks_d.p));
    B3D_LAUNCH(ctx, chunk_sentinel_kernel, 1, 1, 0, out->chunk_start.p, n_chunks_d.p, n);
    B3D_LAUNCH(ctx, chunk_ranges_kernel, (B + 1 + 127) / 128, 128, 0, out->chunk_start.p, n_chunks_d.p, off_d, B, out->chunk_off.p);
    B3D_TRY(ctx->download(out->chunk_off_h.data(), out->chunk_off.p, (size_t)(B + 1) * sizeof(int32_t)));
    out->n_chunks = out->chunk_off_h[B];
    for (int b = 0; b < B; ++b) out->most = std::max(out->most, out->chunk_off_h[b + 1] - out->chunk_off_h[b]);
    return B3D_OK;
}

}  // namespace b3d
