// b3d_common.cuh -- shared host/device helpers of libb200recon.so (sm_100a only).
// Compiled with -fmad=false: voxel keys, squared distances and ordered sums are compared bit-exactly against the
// CPU oracle, so no multiply-add contraction anywhere in this library.
//
// Data model: every kernel works on a BATCH of clouds laid out back to back in one array ("segments"): cloud b owns
// the points [off[b], off[b+1]). A single cloud is a batch of one. Spatial keys carry the cloud id in their top bits,
// so one radix sort / one hash table / one launch serves the whole batch (BASELINE config 4: many frame pairs).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "b200recon.h"

namespace b3d {

int set_error(int code, const char* fmt, ...);

#define B3D_CUDA(expr)                                                                                             \
    do {                                                                                                           \
        cudaError_t e__ = (expr);                                                                                  \
        if (e__ != cudaSuccess)                                                                                    \
            return b3d::set_error(B3D_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define B3D_TRY(expr)                  \
    do {                               \
        int r__ = (expr);              \
        if (r__ != B3D_OK) return r__; \
    } while (0)

#define B3D_REQUIRE(cond, ...)                                       \
    do {                                                             \
        if (!(cond)) return b3d::set_error(B3D_E_INVALID, __VA_ARGS__); \
    } while (0)

// kernel launch on the context stream; counts the launch (bench.py reports it) and checks the launch status
// With profiling on (b3d_ctx_profile) every launch is bracketed by CUDA events on the same stream.
#define B3D_LAUNCH(ctx, kernel, grid, block, smem, ...)                  \
    do {                                                                 \
        if ((ctx)->profiling) (ctx)->prof_begin(#kernel);                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
        (ctx)->launches += 1;                                            \
        if ((ctx)->profiling) (ctx)->prof_end();                         \
        B3D_CUDA(cudaGetLastError());                                    \
    } while (0)

}  // namespace b3d

struct b3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // host -> device copies that overlap the kernels of `stream` (created on first use)
    int sm_count = 148;
    int64_t launches = 0;      // hand-written kernels
    void* pinned = nullptr;    // pinned host staging block (results, counters); valid until the next call that uses it
    size_t pinned_bytes = 0;
    // per-kernel timing (off by default): one event pair per launch, resolved by prof_report()
    bool profiling = false;
    struct ProfRec {
        const char* name;
        cudaEvent_t e0, e1;
        int64_t bytes;  // memory traffic the launch site declared for this launch (0: none declared)
    };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
    void prof_begin(const char* name);
    void prof_end();
    // declares the bytes the most recent launch reads + writes (profile report column 4); no-op unless profiling
    void prof_bytes(int64_t bytes) {
        if (profiling && !prof.empty()) prof.back().bytes = bytes;
    }

    // Scratch memory: a per-context caching allocator over cudaMalloc. The pipelines allocate the same sequence of sizes
    // every step; cudaMallocAsync's pool occasionally re-maps physical memory for large blocks (measured: 20-480 ms stalls,
    // profiles/r01e_alloc_stalls.txt), a size-keyed free list never does. All work of a context runs on ONE stream, so
    // handing a block to the next user is ordered after the previous user's kernels.
    std::multimap<size_t, void*> cache_free;
    std::unordered_map<void*, size_t> cache_live;
    size_t cache_total = 0;
    size_t cache_free_bytes = 0;
    size_t cache_cap_bytes = (size_t)16 << 30;
    void cache_release_all();

    int bind() const;
    int alloc_bytes(void** p, size_t bytes);
    void free_async(void* p);
    template <typename T>
    int alloc(T** p, size_t count) {
        return alloc_bytes(reinterpret_cast<void**>(p), (count ? count : 1) * sizeof(T));
    }
    int sync();
    int ensure_pinned(size_t bytes);
    // device -> pinned host staging -> (after sync) dst
    int download(void* dst_h, const void* src_d, size_t bytes);
    // small host -> device upload from pageable memory (staged by the runtime before the call returns)
    int upload(void* dst_d, const void* src_h, size_t bytes);
    // grid for a grid-stride kernel: enough blocks for n items, capped at blocks_per_sm full waves of the chip
    int grid_for(int64_t n, int block, int per_thread = 1, int blocks_per_sm = 8) const {
        int64_t need = (n + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
        int64_t cap = (int64_t)sm_count * blocks_per_sm;
        if (need < 1) need = 1;
        return (int)(need < cap ? need : cap);
    }
};

namespace b3d {

// RAII scratch buffer on the context's stream-ordered pool
template <typename T>
struct DevBuf {
    b3d_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t count = 0;
    DevBuf() = default;
    explicit DevBuf(b3d_ctx* c) : ctx(c) {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : ctx(o.ctx), p(o.p), count(o.count) { o.p = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; count = o.count;
            o.p = nullptr;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    int alloc(b3d_ctx* c, size_t n) {
        release();
        ctx = c;
        count = n;
        return ctx->alloc(&p, n);
    }
    void release() {
        if (p && ctx) ctx->free_async(p);
        p = nullptr;
    }
    operator T*() const { return p; }
};

// ---------------------------------------------------------------------------------------------------------
// device math shared by kernels
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ T dist2(T dx, T dy, T dz) {
    return (dx * dx + dy * dy) + dz * dz;  // fixed association, no FMA (-fmad=false)
}

template <typename T>
struct Vec3 {
    T x, y, z;
};
template <typename T>
__device__ __forceinline__ Vec3<T> cross3(const Vec3<T>& a, const Vec3<T>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T>
__device__ __forceinline__ T dot3(const Vec3<T>& a, const Vec3<T>& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
__device__ __forceinline__ float b3d_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double b3d_sqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float b3d_acos(float v) { return acosf(v); }
__device__ __forceinline__ double b3d_acos(double v) { return acos(v); }
__device__ __forceinline__ float b3d_cos(float v) { return cosf(v); }
__device__ __forceinline__ double b3d_cos(double v) { return cos(v); }
__device__ __forceinline__ float b3d_abs(float v) { return fabsf(v); }
__device__ __forceinline__ double b3d_abs(double v) { return fabs(v); }

// Symmetric 3x3 (a00 a01 a02 a11 a12 a22): eigenvector of the smallest eigenvalue with the closed-form
// (geometric-tools "robust eigen symmetric 3x3") solver Open3D's estimate_normals uses -- SURVEY.md A.4.
template <typename T>
struct Sym3 {
    T a00, a01, a02, a11, a12, a22;
};

template <typename T>
__device__ inline Vec3<T> sym3_eigvec0(const Sym3<T>& A, T ev) {
    Vec3<T> r0{A.a00 - ev, A.a01, A.a02};
    Vec3<T> r1{A.a01, A.a11 - ev, A.a12};
    Vec3<T> r2{A.a02, A.a12, A.a22 - ev};
    Vec3<T> c01 = cross3(r0, r1), c02 = cross3(r0, r2), c12 = cross3(r1, r2);
    T d0 = dot3(c01, c01), d1 = dot3(c02, c02), d2 = dot3(c12, c12);
    T dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    Vec3<T> c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    T s = b3d_sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    return {c.x / s, c.y / s, c.z / s};
}

template <typename T>
__device__ inline Vec3<T> sym3_eigvec1(const Sym3<T>& A, const Vec3<T>& e0, T ev1) {
    Vec3<T> U;
    if (b3d_abs(e0.x) > b3d_abs(e0.y)) {
        T inv = T(1) / b3d_sqrt(e0.x * e0.x + e0.z * e0.z);
        U = {-e0.z * inv, T(0), e0.x * inv};
    } else {
        T inv = T(1) / b3d_sqrt(e0.y * e0.y + e0.z * e0.z);
        U = {T(0), e0.z * inv, -e0.y * inv};
    }
    Vec3<T> V = cross3(e0, U);
    Vec3<T> AU{A.a00 * U.x + A.a01 * U.y + A.a02 * U.z, A.a01 * U.x + A.a11 * U.y + A.a12 * U.z,
               A.a02 * U.x + A.a12 * U.y + A.a22 * U.z};
    Vec3<T> AV{A.a00 * V.x + A.a01 * V.y + A.a02 * V.z, A.a01 * V.x + A.a11 * V.y + A.a12 * V.z,
               A.a02 * V.x + A.a12 * V.y + A.a22 * V.z};
    T m00 = U.x * AU.x + U.y * AU.y + U.z * AU.z - ev1;
    T m01 = U.x * AV.x + U.y * AV.y + U.z * AV.z;
    T m11 = V.x * AV.x + V.y * AV.y + V.z * AV.z - ev1;
    T a00 = b3d_abs(m00), a01 = b3d_abs(m01), a11 = b3d_abs(m11);
    if (a00 >= a11) {
        if ((a00 > a01 ? a00 : a01) > 0) {
            if (a00 >= a01) {
                m01 /= m00;
                m00 = T(1) / b3d_sqrt(T(1) + m01 * m01);
                m01 *= m00;
            } else {
                m00 /= m01;
                m01 = T(1) / b3d_sqrt(T(1) + m00 * m00);
                m00 *= m01;
            }
            return {m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
        }
        return U;
    }
    if ((a11 > a01 ? a11 : a01) > 0) {
        if (a11 >= a01) {
            m01 /= m11;
            m11 = T(1) / b3d_sqrt(T(1) + m01 * m01);
            m01 *= m11;
        } else {
            m11 /= m01;
            m01 = T(1) / b3d_sqrt(T(1) + m11 * m11);
            m11 *= m01;
        }
        return {m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
    }
    return U;
}

template <typename T>
__device__ inline Vec3<T> sym3_smallest_eigvec(Sym3<T> A) {
    T mx = A.a00;
    mx = A.a01 > mx ? A.a01 : mx;
    mx = A.a02 > mx ? A.a02 : mx;
    mx = A.a11 > mx ? A.a11 : mx;
    mx = A.a12 > mx ? A.a12 : mx;
    mx = A.a22 > mx ? A.a22 : mx;
    if (mx == 0) return {0, 0, 0};
    A.a00 /= mx; A.a01 /= mx; A.a02 /= mx; A.a11 /= mx; A.a12 /= mx; A.a22 /= mx;
    T norm = A.a01 * A.a01 + A.a02 * A.a02 + A.a12 * A.a12;
    if (norm > 0) {
        T q = (A.a00 + A.a11 + A.a22) / 3;
        T b00 = A.a00 - q, b11 = A.a11 - q, b22 = A.a22 - q;
        T p = b3d_sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        T c00 = b11 * b22 - A.a12 * A.a12;
        T c01 = A.a01 * b22 - A.a12 * A.a02;
        T c02 = A.a01 * A.a12 - b11 * A.a02;
        T det = (b00 * c00 - A.a01 * c01 + A.a02 * c02) / (p * p * p);
        T half = det * T(0.5);
        half = half < T(-1) ? T(-1) : (half > T(1) ? T(1) : half);
        T angle = b3d_acos(half) / T(3);
        const T two_thirds_pi = T(2.09439510239319549);
        T beta2 = b3d_cos(angle) * 2;
        T beta0 = b3d_cos(angle + two_thirds_pi) * 2;
        T beta1 = -(beta0 + beta2);
        T e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (half >= 0) {
            Vec3<T> v2 = sym3_eigvec0(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            Vec3<T> v1 = sym3_eigvec1(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross3(v1, v2);
        }
        Vec3<T> v0 = sym3_eigvec0(A, e0);
        if (e0 < e1 && e0 < e2) return v0;
        Vec3<T> v1 = sym3_eigvec1(A, v0, e1);
        if (e1 < e0 && e1 < e2) return v1;
        return cross3(v0, v1);
    }
    if (A.a00 < A.a11 && A.a00 < A.a22) return {1, 0, 0};
    if (A.a11 < A.a00 && A.a11 < A.a22) return {0, 1, 0};
    return {0, 0, 1};
}

// ---------------------------------------------------------------------------------------------------------
// spatial sort: points -> integer cells on a per-cloud lattice -> radix-sorted (cell key, point index) pairs ->
// run heads. Shared by voxel down-sampling (cells = voxels) and by the neighbour-search grid.
// ---------------------------------------------------------------------------------------------------------
struct Lattice {
    double ox, oy, oz;      // origin (legacy voxel: min_bound - vs/2 ; tensor voxel: 0 ; search grid: min_bound)
    double cell;            // cell edge
    int64_t kx0, ky0, kz0;  // integer coordinate bias (key coord = coord - k0), so key coords are >= 0
    int64_t nx, ny, nz;     // extents in cells
    int mode;               // 0: double lattice  floor((p - o) / cell)   (legacy voxel, search grid)
                            // 1: float lattice   floor(float(p) / float(cell))  (tensor voxel, origin 0)
    int morton;             // 1: cell keys are Morton codes of (cx, cy, cz) (search grids: consecutive cells are compact in
                            //    space, so the sorted points double as warp-chunk order); 0: linear (cx*ny + cy)*nz + cz (voxel
                            //    grids: ascending key order = ascending (ix, iy, iz), the canonical output order)
};

__host__ __device__ __forceinline__ unsigned long long morton_spread3(unsigned long long v) {  // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__host__ __device__ __forceinline__ unsigned long long morton_compact3(unsigned long long v) {  // every third bit -> 21 bits
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10c30c30c30c30c3ull;
    v = (v | (v >> 4)) & 0x100f00f00f00f00full;
    v = (v | (v >> 8)) & 0x1f0000ff0000ffull;
    v = (v | (v >> 16)) & 0x1f00000000ffffull;
    v = (v | (v >> 32)) & 0x1fffffull;
    return v;
}

// Key of a cell in the hashed cell table of a search grid: cloud and cell coordinates packed side by side (shift / 3 bits
// per axis -- search grids sort by Morton keys of exactly that width). Cheaper to form per probe than the Morton key.
__host__ __device__ __forceinline__ unsigned long long grid_slot_key(int shift, int cloud, long long x, long long y, long long z) {
    const int ab = shift / 3;
    return ((unsigned long long)cloud << shift) | ((unsigned long long)x << (2 * ab)) | ((unsigned long long)y << ab) | (unsigned long long)z;
}

// key of the cell (x, y, z) (coordinates already biased into [0, n)) inside its cloud
__host__ __device__ __forceinline__ unsigned long long lattice_key(const Lattice& L, long long x, long long y, long long z) {
    if (L.morton) return (morton_spread3((unsigned long long)x) << 2) | (morton_spread3((unsigned long long)y) << 1) | morton_spread3((unsigned long long)z);
    return (unsigned long long)((x * L.ny + y) * L.nz + z);
}

enum LatticeFlavour { kLatLegacyVoxel = 0, kLatTensorVoxel = 1, kLatSearch = 2 };

template <typename T>
__host__ __device__ __forceinline__ void lattice_coord(const Lattice& L, T x, T y, T z, int64_t& cx, int64_t& cy, int64_t& cz) {
    if (L.mode == 1) {
        float c = (float)L.cell;
        cx = (int64_t)floorf((float)x / c);
        cy = (int64_t)floorf((float)y / c);
        cz = (int64_t)floorf((float)z / c);
    } else {
        cx = (int64_t)floor(((double)x - L.ox) / L.cell);
        cy = (int64_t)floor(((double)y - L.oy) / L.cell);
        cz = (int64_t)floor(((double)z - L.oz) / L.cell);
    }
}

// A batch of clouds: B clouds, cloud b = points [off[b], off[b+1]). off_h is the host copy of the device array off.
struct Segments {
    int B = 1;
    const int32_t* off = nullptr;  // device [B+1]
    std::vector<int32_t> off_h;    // host   [B+1]
    int32_t total() const { return off_h.empty() ? 0 : off_h.back(); }
};

struct SpatialSort {
    DevBuf<uint64_t> keys;      // sorted composite keys [n]: (cloud << shift) | linear cell
    DevBuf<uint32_t> order;     // sorted point indices [n] (ascending inside a run: stable sort)
    DevBuf<int32_t> run_start;  // [n_runs + 1] first sorted position of each occupied cell (+ sentinel n)
    DevBuf<int32_t> run_off;    // [B + 1] first run of each cloud (+ sentinel n_runs)
    DevBuf<Lattice> lat;        // [B] device copy of the lattices
    std::vector<Lattice> lat_h;
    std::vector<int32_t> run_off_h;
    int64_t n = 0;
    int64_t n_runs = 0;
    int shift = 0;  // bits of the linear cell key
    int B = 1;
};

// hand-written stable LSD radix sort of (key, value) pairs by key bits [0, end_bit) (b3d_radix.cu); ping-pong buffers
int radix_sort_pairs(b3d_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int64_t n, int end_bit, bool* result_in_a);
// Segmented stable sort of composite keys (segment << low_bits | low key) by their low bits, one segment per cloud
// (packed key + position words, one look-back scatter kernel per 8-bit pass; b3d_radix.cu). order_out[j] = input position
// of the j-th sorted key (ascending among equal keys).
int radix_sort_keys(b3d_ctx* ctx, const uint64_t* keys_in, int64_t n, int low_bits, const std::vector<int32_t>& seg_off_h, const int32_t* seg_off_d,
                    uint64_t* keys_out, uint32_t* order_out);

// bounds_h: [B][6] (min xyz, max xyz) per cloud; clouds with no points get +-DBL_MAX
template <typename T>
int compute_bounds(b3d_ctx* ctx, const T* xyz, const Segments& seg, std::vector<double>* bounds_h);

// Builds the lattices (flavour-specific origin rule) from bounds and sorts.
template <typename T>
int spatial_sort(b3d_ctx* ctx, const T* xyz, const Segments& seg, double cell, int flavour, const std::vector<double>& bounds_h,
                 SpatialSort* out);

// makes a one-cloud Segments over n points (uploads the 2-entry offset array into `storage`)
int single_segment(b3d_ctx* ctx, int64_t n, DevBuf<int32_t>* storage, Segments* seg);
int upload_segments(b3d_ctx* ctx, const std::vector<int32_t>& off_h, DevBuf<int32_t>* storage, Segments* seg);

// ---------------------------------------------------------------------------------------------------------
// neighbour-search grid (device view, passed to kernels by value)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct PointT;
template <>
struct PointT<float> {
    using vec4 = float4;
};
template <>
struct PointT<double> {
    using vec4 = double4;
};

struct HashSlot {
    unsigned long long key;  // composite cell key, ~0ull = empty
    int32_t start, end;      // sorted positions [start, end)
};

template <typename T>
struct GridView {
    const typename PointT<T>::vec4* pts;  // sorted copy: xyz + original (cloud-local) index in .w's bit pattern
    const HashSlot* slots;
    uint32_t mask;
    const Lattice* lat;  // per cloud
    int shift;
    int32_t n;
    const int4* rec;     // float64 grids: sorted like pts, fixed-point coordinates in units of cell / 2^s + sorted position (b3d_stage2.cuh)
};

template <typename T>
struct Grid {
    DevBuf<typename PointT<T>::vec4> pts;
    DevBuf<int4> rec;    // float64 grids only
    DevBuf<HashSlot> slots;
    uint32_t mask = 0;
    SpatialSort sort;
    double cell = 0;
    GridView<T> view() const { return GridView<T>{pts.p, slots.p, mask, sort.lat.p, sort.shift, (int32_t)sort.n, rec.p}; }
};

template <typename T>
int grid_build(b3d_ctx* ctx, const T* xyz, const Segments& seg, double cell, const std::vector<double>* bounds_h, Grid<T>* out);

// Queries re-ordered along a space-filling curve (Hilbert on a quarter-cell lattice of a search grid, or the grid's own
// Morton cell order) and cut into warp chunks: the sorted points are split into runs wherever two consecutive points are
// more than two cells apart (the curve jumped) or a new cloud starts, and every run is cut every 32 points. Chunks are
// mostly full, spatially compact, never straddle clouds, and depend only on their own cloud (not on the rest of the batch).
struct QueryChunks {
    const double4* q = nullptr;   // the queries in chunk order (pts.p, or a search grid's own sorted points)
    DevBuf<double4> pts;          // [n] queries in Morton order, .w = original (batch-global) index (empty when q aliases a grid)
    DevBuf<int32_t> chunk_start;  // [n_chunks + 1]
    DevBuf<int32_t> chunk_off;    // [B + 1] first chunk of every cloud
    std::vector<int32_t> chunk_off_h;
    int32_t n_chunks = 0;
    int32_t most = 0;  // largest chunk count of a single cloud
};
// transforms: device array of B row-major 4x4 matrices applied before the key is taken (stride in doubles), or NULL
int build_query_chunks(b3d_ctx* ctx, const double* pts, const int32_t* off_d, const std::vector<int32_t>& off_h, const SpatialSort& lattices,
                       const double* transforms, int transform_stride, QueryChunks* out);
// chunks over a Morton-ordered search grid's own points (no extra sort, no copy): queries = the grid's points
template <typename T>
struct Grid;
int chunks_from_grid(b3d_ctx* ctx, const Grid<double>& grid, const int32_t* off_d, const std::vector<int32_t>& off_h, QueryChunks* out);

// Chooses the cell size for (k, radius) searches (one trial build measures the occupancy) and builds the grid.
// rmax_out: rings a query has to walk (radius searches), or the ring budget of a k-nearest walk.
template <typename T>
int build_search_grid(b3d_ctx* ctx, const T* xyz, const Segments& seg, int k, double radius, Grid<T>* out, int* rmax_out);

}  // namespace b3d
