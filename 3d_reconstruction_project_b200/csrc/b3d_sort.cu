// b3d_sort.cu -- spatial sort shared by voxel down-sampling and the neighbour-search grid:
// per-cloud bounds -> per-cloud lattice -> composite cell keys -> stable LSD radix sort of (key, point index) ->
// run heads -> hashed cell table. All hand-written (the radix sort is b3d_radix.cu).
#include "b3d_common.cuh"

#include <cstdlib>
#include "b3d_scan.cuh"
#include "b3d_search.cuh"
#include "b3d_stage2.cuh"

#include <cfloat>
#include <climits>
#include <cmath>

namespace b3d {

namespace {

constexpr int kBoundsBlock = 256;

// grid = (blocks per cloud, B). partial: [B][gridDim.x][6]
template <typename T>
__global__ void __launch_bounds__(kBoundsBlock) bounds_partial_kernel(const T* __restrict__ xyz, const int32_t* __restrict__ off,
                                                                     double* __restrict__ partial) {
    const int b = blockIdx.y;
    const int64_t s = off[b], e = off[b + 1];
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    int bad = 0;  // fmin/fmax drop NaNs silently: track them so that the host rejects the cloud
    for (int64_t i = s + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        double x = (double)xyz[3 * i], y = (double)xyz[3 * i + 1], z = (double)xyz[3 * i + 2];
        mn[0] = fmin(mn[0], x); mn[1] = fmin(mn[1], y); mn[2] = fmin(mn[2], z);
        mx[0] = fmax(mx[0], x); mx[1] = fmax(mx[1], y); mx[2] = fmax(mx[2], z);
        bad |= (x != x) | (y != y) | (z != z);
    }
    __shared__ double sm[6][kBoundsBlock / 32];
    const int s_nan = __syncthreads_or(bad);
    for (int d = 0; d < 3; ++d) {
        double a = mn[d], c = mx[d];
        for (int o = 16; o > 0; o >>= 1) {
            a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
            c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
        }
        if ((threadIdx.x & 31) == 0) {
            sm[d][threadIdx.x >> 5] = a;
            sm[3 + d][threadIdx.x >> 5] = c;
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[threadIdx.x][0];
        for (int w = 1; w < kBoundsBlock / 32; ++w) v = threadIdx.x < 3 ? fmin(v, sm[threadIdx.x][w]) : fmax(v, sm[threadIdx.x][w]);
        if (s_nan) v = nan("");
        partial[((int64_t)b * gridDim.x + blockIdx.x) * 6 + threadIdx.x] = v;
    }
}

// one block of six warps per cloud: warp t reduces component t (min x, y, z; max x, y, z) of the cloud's partial boxes, the lanes
// 32 apart (min / max are exact in any order; a single thread per component walked 592 dependent loads: 37 us for one large cloud)
__global__ void __launch_bounds__(192) bounds_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    const int b = blockIdx.x;
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* p = partial + (int64_t)b * nblocks * 6;
    double v = t < 3 ? INFINITY : -INFINITY;
    bool bad = false;
    for (int k = lane; k < nblocks; k += 32) {
        const double u = p[(int64_t)k * 6 + t];
        bad |= u != u;
        v = t < 3 ? fmin(v, u) : fmax(v, u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < 3 ? fmin(v, u) : fmax(v, u);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) out[b * 6 + t] = bad ? nan("") : v;
}

// grid = (blocks per cloud, B)
template <typename T>
__global__ void __launch_bounds__(256) cell_key_kernel(const T* __restrict__ xyz, const int32_t* __restrict__ off,
                                                       const Lattice* __restrict__ lat, int shift, uint64_t* __restrict__ keys) {
    const int b = blockIdx.y;
    const Lattice L = lat[b];
    const int64_t s = off[b], e = off[b + 1];
    const unsigned long long cloud_bits = (unsigned long long)b << shift;
    for (int64_t i = s + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t cx, cy, cz;
        lattice_coord<T>(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], cx, cy, cz);
        cx -= L.kx0; cy -= L.ky0; cz -= L.kz0;
        keys[i] = cloud_bits | lattice_key(L, cx, cy, cz);
    }
}

struct RunHeadPred {
    const uint64_t* keys;
    __device__ __forceinline__ bool operator()(int64_t i) const { return i == 0 || keys[i] != keys[i - 1]; }
};
struct RunHeadEmit {
    int32_t* run_start;
    __device__ __forceinline__ void operator()(int64_t i, int64_t slot) const { run_start[slot] = (int32_t)i; }
};

// run_off[b] = first run whose cloud id is >= b; run_off[B] = n_runs; also writes the run_start sentinel
__global__ void run_offsets_kernel(const uint64_t* __restrict__ keys, int32_t* run_start, const int64_t* __restrict__ n_runs_d, int shift, int B,
                                   int32_t n, int32_t* __restrict__ run_off) {
    const int64_t n_runs = *n_runs_d;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_runs; r += (int64_t)gridDim.x * blockDim.x) {
        const int cb = r < n_runs ? (int)(keys[run_start[r]] >> shift) : B;
        const int pb = r > 0 ? (int)(keys[run_start[r - 1]] >> shift) : -1;
        for (int c = pb + 1; c <= cb; ++c) run_off[c] = (int32_t)r;
        if (r == n_runs) run_start[n_runs] = n;
    }
}

// rec (float64 grids): the staged searches' 16-byte record of the point -- fixed-point coordinates in units of cell / 2^s
// relative to the lattice origin (b3d_stage2.cuh: unit_coord_of_point; record >> s == cell index), .w = sorted position
template <typename T>
__global__ void __launch_bounds__(256) gather_sorted_kernel(const T* __restrict__ xyz, const uint32_t* __restrict__ order, int32_t n,
                                                            typename PointT<T>::vec4* __restrict__ out, const uint64_t* __restrict__ keys,
                                                            const Lattice* __restrict__ lat, int shift, int4* __restrict__ rec) {
    const int us = grid_unit_shift(shift);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = order[i];
        typename PointT<T>::vec4 v;
        v.x = xyz[3 * (int64_t)p];
        v.y = xyz[3 * (int64_t)p + 1];
        v.z = xyz[3 * (int64_t)p + 2];
        v.w = index_as_w(T(0), (int)p);
        out[i] = v;
        if (rec != nullptr) {
            const Lattice L = lat[(int)(keys[i] >> shift)];
            rec[i] = make_int4(unit_coord_of_point((double)v.x, L.ox, L.cell, us), unit_coord_of_point((double)v.y, L.oy, L.cell, us),
                               unit_coord_of_point((double)v.z, L.oz, L.cell, us), (int)i);
        }
    }
}

__global__ void __launch_bounds__(256) hash_clear_kernel(HashSlot* __restrict__ slots, uint32_t n_slots) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
        slots[i].key = kEmptyKey;
        slots[i].start = 0;
        slots[i].end = 0;
    }
}

// keys: sorted composite Morton keys of a search grid; the table is keyed by grid_slot_key (cloud | x | y | z)
__global__ void __launch_bounds__(256) hash_insert_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ run_start, int64_t n_runs,
                                                          int shift, HashSlot* __restrict__ slots, uint32_t mask) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = run_start[r], e = run_start[r + 1];
        const unsigned long long mk = keys[s];
        const unsigned long long m = mk & ((1ull << shift) - 1ull);
        const unsigned long long key = grid_slot_key(shift, (int)(mk >> shift), (long long)morton_compact3(m >> 2), (long long)morton_compact3(m >> 1),
                                                     (long long)morton_compact3(m));
        uint32_t h = hash_key(key) & mask;
        while (true) {
            unsigned long long prev = atomicCAS(&slots[h].key, kEmptyKey, key);
            if (prev == kEmptyKey) {
                slots[h].start = s;
                slots[h].end = e;
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

int bits_for(unsigned __int128 count) {
    int b = 0;
    while (b < 127 && ((unsigned __int128)1 << b) < count) ++b;
    return b;
}

}  // namespace

int upload_segments(b3d_ctx* ctx, const std::vector<int32_t>& off_h, DevBuf<int32_t>* storage, Segments* seg) {
    B3D_TRY(storage->alloc(ctx, off_h.size()));
    B3D_TRY(ctx->upload(storage->p, off_h.data(), off_h.size() * sizeof(int32_t)));
    seg->B = (int)off_h.size() - 1;
    seg->off = storage->p;
    seg->off_h = off_h;
    return B3D_OK;
}

int single_segment(b3d_ctx* ctx, int64_t n, DevBuf<int32_t>* storage, Segments* seg) {
    if (n < 0 || n >= (int64_t)INT32_MAX) return set_error(B3D_E_RANGE, "cloud of %lld points exceeds 2^31-1", (long long)n);
    std::vector<int32_t> off = {0, (int32_t)n};
    return upload_segments(ctx, off, storage, seg);
}

template <typename T>
int compute_bounds(b3d_ctx* ctx, const T* xyz, const Segments& seg, std::vector<double>* bounds_h) {
    const int B = seg.B;
    int64_t longest = 0;
    for (int b = 0; b < B; ++b) longest = std::max<int64_t>(longest, seg.off_h[b + 1] - seg.off_h[b]);
    int blocks = ctx->grid_for(longest, kBoundsBlock, 8, 4);
    // keep the whole launch around a few waves of the chip
    int cap = std::max(1, ctx->sm_count * 8 / std::max(1, B));
    blocks = std::min(blocks, cap);
    DevBuf<double> partial, out;
    B3D_TRY(partial.alloc(ctx, (size_t)B * blocks * 6));
    B3D_TRY(out.alloc(ctx, (size_t)B * 6));
    B3D_LAUNCH(ctx, bounds_partial_kernel<T>, dim3(blocks, B), kBoundsBlock, 0, xyz, seg.off, partial.p);
    B3D_LAUNCH(ctx, bounds_final_kernel, B, 192, 0, partial.p, blocks, out.p);
    bounds_h->resize((size_t)B * 6);
    B3D_TRY(ctx->download(bounds_h->data(), out.p, (size_t)B * 6 * sizeof(double)));
    return B3D_OK;
}
template int compute_bounds<float>(b3d_ctx*, const float*, const Segments&, std::vector<double>*);
template int compute_bounds<double>(b3d_ctx*, const double*, const Segments&, std::vector<double>*);

static int make_lattice(const double* b, bool empty, double cell, int flavour, Lattice* out) {
    Lattice L{};
    L.cell = cell;
    L.mode = 0;
    if (empty) {
        L.nx = L.ny = L.nz = 1;
        *out = L;
        return B3D_OK;
    }
    for (int i = 0; i < 6; ++i)
        if (!std::isfinite(b[i])) return set_error(B3D_E_INVALID, "non-finite coordinates in the cloud");
    if (flavour == kLatLegacyVoxel) {
        // legacy VoxelDownSample: min_bound - vs/2, max_bound + vs/2, error if vs * INT_MAX < max extent
        L.ox = b[0] - cell * 0.5; L.oy = b[1] - cell * 0.5; L.oz = b[2] - cell * 0.5;
        double ext = 0;
        for (int d = 0; d < 3; ++d) ext = std::fmax(ext, (b[3 + d] + cell * 0.5) - (b[d] - cell * 0.5));
        if (cell * (double)INT_MAX < ext) return set_error(B3D_E_RANGE, "voxel_size is too small.");
    } else if (flavour == kLatTensorVoxel) {
        L.mode = 1;
    } else {
        L.ox = b[0]; L.oy = b[1]; L.oz = b[2];
    }
    int64_t* k0[3] = {&L.kx0, &L.ky0, &L.kz0};
    int64_t* nn[3] = {&L.nx, &L.ny, &L.nz};
    const double o[3] = {L.ox, L.oy, L.oz};
    for (int d = 0; d < 3; ++d) {
        double lo, up;
        if (L.mode == 1) {
            float c = (float)cell;
            lo = (double)std::floor((float)b[d] / c);
            up = (double)std::floor((float)b[3 + d] / c);
        } else {
            lo = std::floor((b[d] - o[d]) / cell);
            up = std::floor((b[3 + d] - o[d]) / cell);
        }
        if (!(std::fabs(lo) < 4.0e18) || !(std::fabs(up) < 4.0e18)) return set_error(B3D_E_RANGE, "voxel_size is too small.");
        *k0[d] = (int64_t)lo;
        *nn[d] = (int64_t)up - (int64_t)lo + 1;
    }
    *out = L;
    return B3D_OK;
}

template <typename T>
int spatial_sort(b3d_ctx* ctx, const T* xyz, const Segments& seg, double cell, int flavour, const std::vector<double>& bounds_h,
                 SpatialSort* out) {
    const int B = seg.B;
    const int64_t n = seg.total();
    if (n <= 0) return set_error(B3D_E_INVALID, "spatial_sort: empty batch");
    if (!(cell > 0)) return set_error(B3D_E_INVALID, "spatial_sort: cell size must be positive");
    std::vector<Lattice> lat(B);
    unsigned __int128 max_cells = 1;
    int64_t max_axis = 1;
    for (int b = 0; b < B; ++b) {
        B3D_TRY(make_lattice(&bounds_h[(size_t)b * 6], seg.off_h[b + 1] == seg.off_h[b], cell, flavour, &lat[b]));
        lat[b].morton = flavour == kLatSearch ? 1 : 0;
        unsigned __int128 t = (unsigned __int128)lat[b].nx * (unsigned __int128)lat[b].ny * (unsigned __int128)lat[b].nz;
        if (t > max_cells) max_cells = t;
        max_axis = std::max<int64_t>(max_axis, std::max(std::max(lat[b].nx, lat[b].ny), lat[b].nz));
    }
    int shift = std::max(1, bits_for(max_cells));
    if (flavour == kLatSearch) {
        // Morton keys: three interleaved axes of bits(max axis) bits each
        if (max_axis > (1ll << 21)) return set_error(B3D_E_RANGE, "search grid needs more than 2^21 cells on an axis (cell size too small)");
        shift = 3 * std::max(1, bits_for((unsigned __int128)max_axis));
    }
    const int cloud_bits = bits_for((unsigned __int128)B);
    if (shift + cloud_bits > 63) return set_error(B3D_E_RANGE, "voxel_size is too small. (cell lattice needs %d + %d key bits)", shift, cloud_bits);
    const int end_bit = shift + cloud_bits;

    out->B = B;
    out->shift = shift;
    out->lat_h = lat;
    B3D_TRY(out->lat.alloc(ctx, B));
    B3D_TRY(ctx->upload(out->lat.p, lat.data(), (size_t)B * sizeof(Lattice)));

    DevBuf<uint64_t> keys_in(ctx), keys_out(ctx);
    DevBuf<uint32_t> ord_out(ctx);
    B3D_TRY(keys_in.alloc(ctx, n));
    B3D_TRY(keys_out.alloc(ctx, n));
    B3D_TRY(ord_out.alloc(ctx, n));
    {
        int64_t longest = 0;
        for (int b = 0; b < B; ++b) longest = std::max<int64_t>(longest, seg.off_h[b + 1] - seg.off_h[b]);
        int blocks = (int)std::min<int64_t>((longest + 255) / 256, std::max(1, ctx->sm_count * 16 / B));
        blocks = std::max(1, blocks);
        B3D_LAUNCH(ctx, cell_key_kernel<T>, dim3(blocks, B), 256, 0, xyz, seg.off, out->lat.p, shift, keys_in.p);
    }
    // the cloud id in the top bits is already in order (clouds lie back to back): every cloud is sorted on its cell bits
    (void)end_bit;
    B3D_TRY(radix_sort_keys(ctx, keys_in.p, n, shift, seg.off_h, seg.off, keys_out.p, ord_out.p));
    keys_in.release();
    // run heads -> run_start[], count -> host
    DevBuf<int32_t> run_start(ctx);
    DevBuf<int64_t> n_runs_d(ctx);
    B3D_TRY(run_start.alloc(ctx, n + 1));
    B3D_TRY(n_runs_d.alloc(ctx, 1));
    B3D_TRY(compact(ctx, RunHeadPred{keys_out.p}, RunHeadEmit{run_start.p}, n, n_runs_d.p));
    B3D_TRY(out->run_off.alloc(ctx, B + 1));
    // the kernel covers n_runs + 1 entries; n is an upper bound on n_runs
    B3D_LAUNCH(ctx, run_offsets_kernel, ctx->grid_for(n + 1, 256, 1, 8), 256, 0, keys_out.p, run_start.p, n_runs_d.p, shift, B, (int32_t)n,
               out->run_off.p);
    int64_t n_runs = 0;
    B3D_TRY(ctx->download(&n_runs, n_runs_d.p, sizeof(int64_t)));
    out->run_off_h.resize(B + 1);
    B3D_TRY(ctx->download(out->run_off_h.data(), out->run_off.p, (size_t)(B + 1) * sizeof(int32_t)));
    out->keys = std::move(keys_out);
    out->order = std::move(ord_out);
    out->run_start = std::move(run_start);
    out->n = n;
    out->n_runs = n_runs;
    return B3D_OK;
}

template int spatial_sort<float>(b3d_ctx*, const float*, const Segments&, double, int, const std::vector<double>&, SpatialSort*);
template int spatial_sort<double>(b3d_ctx*, const double*, const Segments&, double, int, const std::vector<double>&, SpatialSort*);

template <typename T>
int grid_build(b3d_ctx* ctx, const T* xyz, const Segments& seg, double cell, const std::vector<double>* bounds_in, Grid<T>* out) {
    std::vector<double> bounds_local;
    const std::vector<double>* bounds = bounds_in;
    if (!bounds) {
        B3D_TRY(compute_bounds<T>(ctx, xyz, seg, &bounds_local));
        bounds = &bounds_local;
    }
    B3D_TRY(spatial_sort<T>(ctx, xyz, seg, cell, kLatSearch, *bounds, &out->sort));
    const int64_t n = out->sort.n;
    B3D_TRY(out->pts.alloc(ctx, n));
    if (sizeof(T) == 8) B3D_TRY(out->rec.alloc(ctx, n));
    B3D_LAUNCH(ctx, gather_sorted_kernel<T>, ctx->grid_for(n, 256, 1, 16), 256, 0, xyz, out->sort.order.p, (int32_t)n, out->pts.p, out->sort.keys.p,
               out->sort.lat.p, out->sort.shift, sizeof(T) == 8 ? out->rec.p : (int4*)nullptr);
    // Open addressing with linear probing; a warp's staging round waits for the SLOWEST of its 32 lookups, and most of them ask
    // for empty cells (the box of a surface patch is mostly air), i.e. run to the first empty slot. At a load of 1/2 that is 2.5
    // dependent loads on average and ~8 for the slowest lane; at 1/8 .. 1/16 nearly every lookup is ONE load. Slots are 16 bytes
    // and touched sparsely, so the larger table costs address space, not cache (B3D_HASH_SLOTS_PER_CELL, default 8).
    static const int per_cell = getenv("B3D_HASH_SLOTS_PER_CELL") ? std::max(2, atoi(getenv("B3D_HASH_SLOTS_PER_CELL"))) : 8;
    uint32_t n_slots = 1024;
    while ((int64_t)n_slots < (int64_t)per_cell * out->sort.n_runs && n_slots < (1u << 30)) n_slots <<= 1;
    out->mask = n_slots - 1;
    out->cell = cell;
    B3D_TRY(out->slots.alloc(ctx, n_slots));
    B3D_LAUNCH(ctx, hash_clear_kernel, ctx->grid_for(n_slots, 256, 1, 16), 256, 0, out->slots.p, n_slots);
    B3D_LAUNCH(ctx, hash_insert_kernel, ctx->grid_for(out->sort.n_runs, 256, 1, 16), 256, 0, out->sort.keys.p, out->sort.run_start.p,
               out->sort.n_runs, out->sort.shift, out->slots.p, out->mask);
    return B3D_OK;
}
template int grid_build<float>(b3d_ctx*, const float*, const Segments&, double, const std::vector<double>*, Grid<float>*);
template int grid_build<double>(b3d_ctx*, const double*, const Segments&, double, const std::vector<double>*, Grid<double>*);

}  // namespace b3d
