// b3d_voxel.cu -- K2: voxel down-sampling.
//   legacy : PointCloud.voxel_down_sample (float64; origin min_bound - vs/2)   pointcloud_alignment.py:22-23
//   tensor : t.PointCloud.voxel_down_sample (float32; origin 0)                pointcloud_capture.py:50
// Pipeline: per-cloud lattice -> composite voxel keys -> stable radix sort -> run heads -> ORDERED per-voxel sums.
// The reference sums each voxel's points sequentially in input order; a stable sort keeps that order inside a run,
// and each run is reduced by a strictly sequential chain of additions (one thread per short run, one block per
// long run with the sequential chain fed from shared memory), so the means are bit-identical to the CPU path.
// Output order: ascending (cloud, ix, iy, iz).
#include "b3d_common.cuh"

namespace b3d {
namespace {

constexpr int kLongRun = 64;
constexpr int kLongBlock = 256;

template <typename IndexT>
__device__ __forceinline__ void decode_voxel(const Lattice& L, unsigned long long linear, IndexT* out) {
    const long long nz = L.nz, ny = L.ny;
    const long long cz = (long long)(linear % (unsigned long long)nz);
    const long long t = (long long)(linear / (unsigned long long)nz);
    const long long cy = t % ny, cx = t / ny;
    out[0] = (IndexT)(cx + L.kx0);
    out[1] = (IndexT)(cy + L.ky0);
    out[2] = (IndexT)(cz + L.kz0);
}

// one thread per run; runs longer than kLongRun are queued for the block-per-run kernel
template <typename T, typename IndexT>
__global__ void __launch_bounds__(256) voxel_reduce_short_kernel(const T* __restrict__ xyz, const T* __restrict__ a0, const T* __restrict__ a1,
                                                                 const uint64_t* __restrict__ keys, const uint32_t* __restrict__ order,
                                                                 const int32_t* __restrict__ run_start, int64_t n_runs,
                                                                 const Lattice* __restrict__ lat, int shift, T* __restrict__ o_xyz,
                                                                 T* __restrict__ o_a0, T* __restrict__ o_a1, IndexT* __restrict__ o_index,
                                                                 int32_t* __restrict__ o_count, int32_t* __restrict__ long_list,
                                                                 int32_t* __restrict__ long_count, double* __restrict__ o_xyz64) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = run_start[r], e = run_start[r + 1];
        const uint64_t key = keys[s];
        if (o_index != nullptr) {
            const int cloud = (int)(key >> shift);
            decode_voxel<IndexT>(lat[cloud], key & ((1ull << shift) - 1ull), o_index + 3 * r);
        }
        if (o_count != nullptr) o_count[r] = e - s;
        if (e - s > kLongRun) {
            long_list[atomicAdd(long_count, 1)] = (int32_t)r;
            continue;
        }
        T sp[3] = {0, 0, 0}, s0[3] = {0, 0, 0}, s1[3] = {0, 0, 0};
        for (int32_t j = s; j < e; ++j) {
            const int64_t p = order[j];
            sp[0] += xyz[3 * p]; sp[1] += xyz[3 * p + 1]; sp[2] += xyz[3 * p + 2];
            if (a0 != nullptr) { s0[0] += a0[3 * p]; s0[1] += a0[3 * p + 1]; s0[2] += a0[3 * p + 2]; }
            if (a1 != nullptr) { s1[0] += a1[3 * p]; s1[1] += a1[3 * p + 1]; s1[2] += a1[3 * p + 2]; }
        }
        const T cnt = (T)(e - s);
        // o_xyz64: the means widened to float64 (exact) instead of the flavour's own type -- the registration pipeline's to_legacy()
        if (o_xyz64 != nullptr) {
            o_xyz64[3 * r] = (double)(sp[0] / cnt); o_xyz64[3 * r + 1] = (double)(sp[1] / cnt); o_xyz64[3 * r + 2] = (double)(sp[2] / cnt);
        } else {
            o_xyz[3 * r] = sp[0] / cnt; o_xyz[3 * r + 1] = sp[1] / cnt; o_xyz[3 * r + 2] = sp[2] / cnt;
        }
        if (a0 != nullptr && o_a0 != nullptr) { o_a0[3 * r] = s0[0] / cnt; o_a0[3 * r + 1] = s0[1] / cnt; o_a0[3 * r + 2] = s0[2] / cnt; }
        if (a1 != nullptr && o_a1 != nullptr) { o_a1[3 * r] = s1[0] / cnt; o_a1[3 * r + 1] = s1[1] / cnt; o_a1[3 * r + 2] = s1[2] / cnt; }
    }
}

// one block per long run: 256 points are staged per step (coalesced index loads, gathered coordinates); the nine
// component sums are nine independent sequential chains (threads 0..8), so the order of additions is the input order.
// Chunks whose values are all zero are skipped: x + 0 == x exactly (the sums start at +0), which makes the voxel that
// swallows every zero-depth pixel of an rs.pointcloud() frame (SURVEY.md "hard parts") cost one pass of loads.
template <typename T>
__global__ void __launch_bounds__(kLongBlock) voxel_reduce_long_kernel(const T* __restrict__ xyz, const T* __restrict__ a0, const T* __restrict__ a1,
                                                                       const uint32_t* __restrict__ order, const int32_t* __restrict__ run_start,
                                                                       const int32_t* __restrict__ long_list, const int32_t* __restrict__ long_count,
                                                                       T* __restrict__ o_xyz, T* __restrict__ o_a0, T* __restrict__ o_a1,
                                                                       double* __restrict__ o_xyz64) {
    __shared__ T sm[9][kLongBlock + 1];
    const int n_long = *long_count;
    const int n_comp = 3 + (a0 != nullptr ? 3 : 0) + (a1 != nullptr ? 3 : 0);
    for (int li = blockIdx.x; li < n_long; li += gridDim.x) {
        const int32_t r = long_list[li];
        const int32_t s = run_start[r], e = run_start[r + 1];
        T acc = 0;  // threads 0..8: one component each
        for (int32_t base = s; base < e; base += kLongBlock) {
            const int32_t j = base + threadIdx.x;
            T v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (j < e) {
                const int64_t p = order[j];
                v[0] = xyz[3 * p]; v[1] = xyz[3 * p + 1]; v[2] = xyz[3 * p + 2];
                if (a0 != nullptr) { v[3] = a0[3 * p]; v[4] = a0[3 * p + 1]; v[5] = a0[3 * p + 2]; }
                if (a1 != nullptr) {
                    const int o = a0 != nullptr ? 6 : 3;
                    v[o] = a1[3 * p]; v[o + 1] = a1[3 * p + 1]; v[o + 2] = a1[3 * p + 2];
                }
            }
            int nz = 0;
#pragma unroll
            for (int c = 0; c < 9; ++c) nz |= (v[c] != T(0)) ? 1 : 0;
            if (__syncthreads_or(nz)) {
#pragma unroll
                for (int c = 0; c < 9; ++c) sm[c][threadIdx.x] = v[c];
                __syncthreads();
                if ((int)threadIdx.x < n_comp) {
                    const int cnt = min(kLongBlock, e - base);
                    const T* row = sm[threadIdx.x];
                    for (int k = 0; k < cnt; ++k) acc += row[k];
                }
            }
            __syncthreads();
        }
        if ((int)threadIdx.x < n_comp) {
            const T m = acc / (T)(e - s);
            const int c = threadIdx.x;
            if (c < 3) {
                if (o_xyz64 != nullptr) o_xyz64[3 * (int64_t)r + c] = (double)m;
                else o_xyz[3 * (int64_t)r + c] = m;
            } else if (a0 != nullptr && c < 6) {
                if (o_a0 != nullptr) o_a0[3 * (int64_t)r + (c - 3)] = m;
            } else {
                if (o_a1 != nullptr) o_a1[3 * (int64_t)r + (c - (a0 != nullptr ? 6 : 3))] = m;
            }
        }
    }
}

}  // namespace

// Batched voxel down-sampling. Outputs have capacity seg.total() rows; out_sort (optional) receives the sort products
// (run offsets per cloud = down-sampled cloud offsets).
template <typename T, typename IndexT>
int voxel_downsample_batch(b3d_ctx* ctx, const T* xyz, const T* a0, const T* a1, const Segments& seg, double voxel, int flavour, T* o_xyz,
                           T* o_a0, T* o_a1, IndexT* o_index, int32_t* o_count, SpatialSort* out_sort, const std::vector<double>* bounds_in,
                           double* o_xyz64) {
    std::vector<double> bounds_local;
    const std::vector<double>* bounds = bounds_in;
    if (!bounds) {
        B3D_TRY(compute_bounds<T>(ctx, xyz, seg, &bounds_local));
        bounds = &bounds_local;
    }
    SpatialSort local;
    SpatialSort* ss = out_sort ? out_sort : &local;
    B3D_TRY(spatial_sort<T>(ctx, xyz, seg, voxel, flavour, *bounds, ss));
    DevBuf<int32_t> long_list, long_count;
    const int64_t max_long = ss->n / kLongRun + 1;
    B3D_TRY(long_list.alloc(ctx, (size_t)max_long));
    B3D_TRY(long_count.alloc(ctx, 1));
    B3D_CUDA(cudaMemsetAsync(long_count.p, 0, sizeof(int32_t), ctx->stream));
    B3D_LAUNCH(ctx, (voxel_reduce_short_kernel<T, IndexT>), ctx->grid_for(ss->n_runs, 256, 1, 16), 256, 0, xyz, a0, a1, ss->keys.p, ss->order.p,
               ss->run_start.p, ss->n_runs, ss->lat.p, ss->shift, o_xyz, o_a0, o_a1, o_index, o_count, long_list.p, long_count.p, o_xyz64);
    const int long_grid = (int)std::min<int64_t>(max_long, (int64_t)ctx->sm_count * 8);
    B3D_LAUNCH(ctx, voxel_reduce_long_kernel<T>, long_grid, kLongBlock, 0, xyz, a0, a1, ss->order.p, ss->run_start.p, long_list.p, long_count.p,
               o_xyz, o_a0, o_a1, o_xyz64);
    return B3D_OK;
}

template int voxel_downsample_batch<float, int64_t>(b3d_ctx*, const float*, const float*, const float*, const Segments&, double, int, float*,
                                                    float*, float*, int64_t*, int32_t*, SpatialSort*, const std::vector<double>*, double*);
template int voxel_downsample_batch<double, int32_t>(b3d_ctx*, const double*, const double*, const double*, const Segments&, double, int, double*,
                                                     double*, double*, int32_t*, int32_t*, SpatialSort*, const std::vector<double>*, double*);

}  // namespace b3d

using namespace b3d;

extern "C" {

int b3d_voxel_downsample_legacy(b3d_ctx* ctx, const double* xyz, const double* colors, const double* normals, int64_t n, double voxel_size,
                                double* out_xyz, double* out_colors, double* out_normals, int32_t* out_index, int32_t* out_count, int64_t* m_h) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(m_h != nullptr, "b3d_voxel_downsample_legacy: m_h is NULL");
    *m_h = 0;
    B3D_REQUIRE(voxel_size > 0.0, "voxel_size <= 0.");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && out_xyz, "b3d_voxel_downsample_legacy: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    SpatialSort ss;
    // colours first, normals second; a missing colour array moves the normals into the first attribute slot
    const double* a0 = colors ? colors : normals;
    double* o0 = colors ? out_colors : out_normals;
    const double* a1 = colors ? normals : nullptr;
    double* o1 = colors ? out_normals : nullptr;
    B3D_TRY((voxel_downsample_batch<double, int32_t>(ctx, xyz, a0, a1, seg, voxel_size, kLatLegacyVoxel, out_xyz, o0, o1, out_index, out_count, &ss,
                                                      nullptr, nullptr)));
    *m_h = ss.n_runs;
    return ctx->sync();
}

int b3d_voxel_downsample_tensor(b3d_ctx* ctx, const float* xyz, const float* attr, int64_t n, float voxel_size, float* out_xyz, float* out_attr,
                                int64_t* out_index, int32_t* out_count, int64_t* m_h) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(m_h != nullptr, "b3d_voxel_downsample_tensor: m_h is NULL");
    *m_h = 0;
    B3D_REQUIRE(voxel_size > 0.0f, "voxel_size must be positive.");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && out_xyz, "b3d_voxel_downsample_tensor: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    SpatialSort ss;
    B3D_TRY((voxel_downsample_batch<float, int64_t>(ctx, xyz, attr, nullptr, seg, (double)voxel_size, kLatTensorVoxel, out_xyz, out_attr, nullptr,
                                                     out_index, out_count, &ss, nullptr, nullptr)));
    *m_h = ss.n_runs;
    return ctx->sync();
}

}  // extern "C"
