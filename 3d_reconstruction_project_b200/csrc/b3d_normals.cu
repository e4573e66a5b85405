// b3d_normals.cu -- K3: neighbourhood covariances + closed-form symmetric eigen-solve (normals), GICP covariances,
// and the two outlier filters that share the same neighbour search.
#include "b3d_common.cuh"
#include "b3d_scan.cuh"
#include "b3d_search.cuh"
#include "b3d_stage.cuh"
#include "b3d_stage2.cuh"
#include "b3d_sortnet.cuh"

#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace b3d {

// ---- cell-size policy -------------------------------------------------------------------------------------------
// radius searches: cell slightly above the radius (so that ring 1 always suffices); k-nearest searches: cells holding
// about k/2 points each (surface model: occupancy ~ cell^2). One trial build measures the occupancy; at most two
// rebuilds follow.
template <typename T>
int build_search_grid(b3d_ctx* ctx, const T* xyz, const Segments& seg, int k, double radius, Grid<T>* out, int* rmax_out) {
    std::vector<double> bounds;
    B3D_TRY(compute_bounds<T>(ctx, xyz, seg, &bounds));
    const int64_t n = seg.total();
    double ext[3] = {0, 0, 0};
    for (int b = 0; b < seg.B; ++b) {
        if (seg.off_h[b + 1] == seg.off_h[b]) continue;
        for (int d = 0; d < 3; ++d) {
            double e = bounds[(size_t)b * 6 + 3 + d] - bounds[(size_t)b * 6 + d];
            if (!std::isfinite(e)) return set_error(B3D_E_INVALID, "non-finite coordinates in the cloud");
            ext[d] = std::max(ext[d], e);
        }
    }
    const double target = std::max(2.0, 0.5 * (double)std::max(k, 1));
    double cell;
    if (radius > 0) {
        cell = radius * 1.001;
    } else {
        double area = std::max(std::max(ext[0] * ext[1], ext[1] * ext[2]), ext[0] * ext[2]);
        double per_cloud = std::max<double>(1.0, (double)n / std::max(1, seg.B));
        double spacing = std::sqrt(area / per_cloud);
        cell = spacing * std::sqrt(target);
        double longest = std::max(std::max(ext[0], ext[1]), ext[2]);
        if (!(cell > 0)) cell = longest > 0 ? longest / 64.0 : 1.0;
    }
    for (int attempt = 0; attempt < 3; ++attempt) {
        B3D_TRY(grid_build<T>(ctx, xyz, seg, cell, &bounds, out));
        const double occ = (double)n / (double)std::max<int64_t>(1, out->sort.n_runs);
        if (attempt == 2) break;
        if (radius > 0) {
            // far too many candidates per cell for the k wanted: shrink towards the k-nearest scale
            if (occ > 4.0 * target) {
                double next = cell * std::sqrt(target / occ);
                if (next < cell / 1.5) { cell = next; continue; }
            }
            break;
        }
        if (occ < 0.5 * target || occ > 2.5 * target) {
            double f = std::sqrt(target / occ);
            f = std::min(std::max(f, 0.125), 8.0);
            // a cloud much smaller than k cannot reach the target occupancy: stop once one cell spans it
            double longest = std::max(std::max(ext[0], ext[1]), ext[2]);
            if (f > 1.0 && cell > longest && longest > 0) break;
            cell *= f;
            continue;
        }
        break;
    }
    if (rmax_out) *rmax_out = radius > 0 ? rings_for_radius(radius, cell) : kMaxRing;
    return B3D_OK;
}
template int build_search_grid<float>(b3d_ctx*, const float*, const Segments&, int, double, Grid<float>*, int*);
template int build_search_grid<double>(b3d_ctx*, const double*, const Segments&, int, double, Grid<double>*, int*);

// debug counters of the staged normals (B3D_ICP_STATS=1): [0] chunks, [1] overflowed chunks, [2] lanes needing the k-nearest cut,
// [3] staged candidates, [4] valid lanes
__device__ unsigned long long g_nrm_stats[8];

namespace {

// ---- normals ------------------------------------------------------------------------------------------------------
// One thread per point, walked in sorted (cell) order so that neighbouring threads touch the same cells.
// LEGACY (T = double): raw-moment covariance in neighbour order, /n            
// TENSOR (T = float) : two-pass centred covariance in double, /(n-1), eigen in float
template <typename T, int KMAX, bool TENSOR>
__global__ void __launch_bounds__(128) normals_kernel(GridView<T> g, const uint64_t* __restrict__ keys, const int32_t* __restrict__ off, int k,
                                                      bool use_radius, T r2, int rmax, const T* __restrict__ prior, T* __restrict__ normals,
                                                      const int* __restrict__ todo, const int* __restrict__ todo_count) {
    // todo == NULL: every point of the grid; else only the sorted positions the staged kernel flagged
    const int n_work = todo != nullptr ? *todo_count : g.n;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_work; t += gridDim.x * blockDim.x) {
    const int pos = todo != nullptr ? todo[t] : t;
    const typename PointT<T>::vec4 q = ld_point(g.pts + pos);
    const int cloud = (int)(keys[pos] >> g.shift);
    TopK<T, KMAX> tk;
    knn_hybrid_query<T, KMAX>(g, off, cloud, q.x, q.y, q.z, k, use_radius, r2, rmax, tk);
    const int c = tk.n;
    Sym3<T> C{T(1), T(0), T(0), T(1), T(0), T(1)};
    if (c >= 3) {
        if (!TENSOR) {
            double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < c; ++j) {
                const typename PointT<T>::vec4 p = ld_point(g.pts + tk.pos[j]);
                const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
                cu[0] += x; cu[1] += y; cu[2] += z;
                cu[3] += x * x; cu[4] += x * y; cu[5] += x * z;
                cu[6] += y * y; cu[7] += y * z; cu[8] += z * z;
            }
            const double cn = (double)c;
#pragma unroll
            for (int j = 0; j < 9; ++j) cu[j] /= cn;
            C.a00 = (T)(cu[3] - cu[0] * cu[0]);
            C.a11 = (T)(cu[6] - cu[1] * cu[1]);
            C.a22 = (T)(cu[8] - cu[2] * cu[2]);
            C.a01 = (T)(cu[4] - cu[0] * cu[1]);
            C.a02 = (T)(cu[5] - cu[0] * cu[2]);
            C.a12 = (T)(cu[7] - cu[1] * cu[2]);
        } else {
            double ce[3] = {0, 0, 0};
            for (int j = 0; j < c; ++j) {
                const typename PointT<T>::vec4 p = ld_point(g.pts + tk.pos[j]);
                ce[0] += (double)p.x; ce[1] += (double)p.y; ce[2] += (double)p.z;
            }
            ce[0] /= c; ce[1] /= c; ce[2] /= c;
            double cu[6] = {0, 0, 0, 0, 0, 0};
            for (int j = 0; j < c; ++j) {
                const typename PointT<T>::vec4 p = ld_point(g.pts + tk.pos[j]);
                const double x = (double)p.x - ce[0], y = (double)p.y - ce[1], z = (double)p.z - ce[2];
                cu[0] += x * x; cu[1] += y * y; cu[2] += z * z; cu[3] += x * y; cu[4] += x * z; cu[5] += y * z;
            }
            const double nf = (double)(c - 1);
            C.a00 = (T)(cu[0] / nf); C.a11 = (T)(cu[1] / nf); C.a22 = (T)(cu[2] / nf);
            C.a01 = (T)(cu[3] / nf); C.a02 = (T)(cu[4] / nf); C.a12 = (T)(cu[5] / nf);
        }
    }
    Vec3<T> nrm = sym3_smallest_eigvec<T>(C);
    const int64_t oi = point_index(q);
    if (!TENSOR) {
        const T len = b3d_sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        if (prior != nullptr) {
            const T ox = prior[3 * oi], oy = prior[3 * oi + 1], oz = prior[3 * oi + 2];
            if (len == T(0)) nrm = {ox, oy, oz};
            else if (nrm.x * ox + nrm.y * oy + nrm.z * oz < T(0)) nrm = {-nrm.x, -nrm.y, -nrm.z};
        } else if (len == T(0)) {
            nrm = {T(0), T(0), T(1)};
        }
    } else {
        if (nrm.x == T(0) && nrm.y == T(0) && nrm.z == T(0)) nrm = {T(0), T(0), T(1)};
    }
    normals[3 * oi] = nrm.x;
    normals[3 * oi + 1] = nrm.y;
    normals[3 * oi + 2] = nrm.z;
    }
}

// ---- staged legacy normals (radius searches, k <= kNrmList) --------------------------------------------------------------
// One warp per compact chunk of <= 32 Morton-ordered points (build_query_chunks). The warp stages every point inside the
// chunk's bounding box + radius in shared memory (b3d_stage.cuh); each lane scans that list in float32, confirms the
// borderline candidates in float64 (exact d2 < r2 rule), and keeps the indices of its neighbours in a small shared list.
// Neighbour SETS are exact; the nine raw moments are then summed in staged order (the reference sums in distance order:
// same set, same arithmetic, a different order of float64 additions). Lanes that would need a k-nearest cut (more than k
// neighbours inside the radius) or whose chunk overflows the staging buffers are queued (sorted positions) and redone by the
// per-lane kernel afterwards.
constexpr int kNrmBlock = 128;
constexpr int kNrmList = 32;   // neighbours per lane kept in shared memory (k <= 32 on this path)
#ifndef B3D_NRM_CAP
#define B3D_NRM_CAP 192
#endif
constexpr int kNrmCap = B3D_NRM_CAP;   // staged candidates per warp

struct NrmWarpSmem {
    float4 cand[kNrmCap];
    double xyz[3 * kNrmCap];
    int pos[kNrmCap];
    unsigned int list[32][kNrmList + 1];  // per lane: (float bits of d2 with the low 9 bits replaced by the candidate id), ascending
    StageScratch stage;
};

__global__ void __launch_bounds__(kNrmBlock) normals_staged_kernel(GridView<double> g, const double4* __restrict__ qpts,
                                                                  const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_off,
                                                                  int B, int n_chunks, int k, double radius, double r2,
                                                                  const double* __restrict__ prior, double* __restrict__ normals,
                                                                  int* __restrict__ todo, int* __restrict__ todo_count, int stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    NrmWarpSmem& S = reinterpret_cast<NrmWarpSmem*>(smem_raw)[warp];
    unsigned int* list = S.list[lane];
    for (int c = blockIdx.x * (kNrmBlock / 32) + warp; c < n_chunks; c += gridDim.x * (kNrmBlock / 32)) {
        // cloud of this chunk: last b with chunk_off[b] <= c
        int lo_b = 0, hi_b = B;
        while (hi_b - lo_b > 1) {
            const int m = (lo_b + hi_b) >> 1;
            if (chunk_off[m] <= c) lo_b = m; else hi_b = m;
        }
        const int cloud = lo_b;
        const int32_t i = chunk_start[c] + lane;
        const bool valid = i < chunk_start[c + 1];
        double qx = 0, qy = 0, qz = 0;
        int oi = 0;
        if (valid) {
            const double4 q = ld_point(qpts + i);
            qx = q.x; qy = q.y; qz = q.z;
            oi = point_index(q);
        }
        const float bigf = 3.0e38f;
        float lx = valid ? (float)qx : bigf, ly = valid ? (float)qy : bigf, lz = valid ? (float)qz : bigf;
        float hx = valid ? (float)qx : -bigf, hy = valid ? (float)qy : -bigf, hz = valid ? (float)qz : -bigf;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o));
            ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
            lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
            hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
        }
        const double wid = 2.0e-7;
        const double pad = radius * (1.0 + 1e-9) + 1e-12;
        const double lo[3] = {(double)lx - wid * fabs((double)lx) - pad, (double)ly - wid * fabs((double)ly) - pad, (double)lz - wid * fabs((double)lz) - pad};
        const double hi[3] = {(double)hx + wid * fabs((double)hx) + pad, (double)hy + wid * fabs((double)hy) + pad, (double)hz + wid * fabs((double)hz) + pad};
        const double center[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
        const float H = (float)(0.5 * fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2])) * 1.0001f;
        const int count = warp_stage_box(g, cloud, lo, hi, center, S.cand, S.pos, &S.stage, kNrmCap, S.xyz);
        if (stats) {
            const unsigned int vm = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) {
                atomicAdd(&g_nrm_stats[0], 1ull);
                if (count < 0) atomicAdd(&g_nrm_stats[1], 1ull); else atomicAdd(&g_nrm_stats[3], (unsigned long long)count);
                atomicAdd(&g_nrm_stats[4], (unsigned long long)__popc(vm));
            }
        }
        if (count < 0) {
            if (valid) todo[atomicAdd(todo_count, 1)] = i;
            continue;
        }
        // ---- scan: neighbours with d2 < r2, kept sorted by (float) distance -------------------------------------------
        const float rx = (float)(qx - center[0]), ry = (float)(qy - center[1]), rz = (float)(qz - center[2]);
        const float fx = -2.0f * rx, fy = -2.0f * ry, fz = -2.0f * rz;
        const float qq = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
        const float band = 8.0e-6f * H * H + 4.0e-7f * (float)r2;
        const float r2f = (float)r2;
        int n = 0;
        bool spill = false;
        if (valid) {
            for (int j = 0; j < count; ++j) {
                const float4 cj = S.cand[j];
                const float d2f = fmaf(fx, cj.x, fmaf(fy, cj.y, fmaf(fz, cj.z, cj.w))) + qq;
                bool acc = d2f < r2f - band;
                if (!acc && d2f <= r2f + band)
                    acc = dist2<double>(qx - S.xyz[3 * j], qy - S.xyz[3 * j + 1], qz - S.xyz[3 * j + 2]) < r2;
                if (acc) {
                    if (n < kNrmList) list[n] = (__float_as_uint(fmaxf(d2f, 0.0f)) & ~511u) | (unsigned int)j;
                    else spill = true;
                    ++n;
                }
            }
        }
        __syncwarp();
        if (!valid) continue;
        if (spill || n > k) {
            if (stats) atomicAdd(&g_nrm_stats[2], 1ull);
            todo[atomicAdd(todo_count, 1)] = i;  // needs the k nearest of more than k in-radius points: per-lane kernel
            continue;
        }
        // order by distance (the reference accumulates its moments in (d2, index) order): insertion sort of the short list,
        // all lanes of the warp sorting their own lists at the same time
        for (int a = 1; a < n; ++a) {
            const unsigned int key = list[a];
            int m = a;
            while (m > 0 && list[m - 1] > key) {
                list[m] = list[m - 1];
                --m;
            }
            list[m] = key;
        }
        // ---- raw-moment covariance over the neighbour set, closed-form eigenvector ------------------------------------
        Sym3<double> C{1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
        if (n >= 3) {
            double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < n; ++j) {
                const int cidx = (int)(list[j] & 511u);
                const double px = S.xyz[3 * cidx], py = S.xyz[3 * cidx + 1], pz = S.xyz[3 * cidx + 2];
                cu[0] += px; cu[1] += py; cu[2] += pz;
                cu[3] += px * px; cu[4] += px * py; cu[5] += px * pz;
                cu[6] += py * py; cu[7] += py * pz; cu[8] += pz * pz;
            }
            const double cn = (double)n;
#pragma unroll
            for (int j = 0; j < 9; ++j) cu[j] /= cn;
            C.a00 = cu[3] - cu[0] * cu[0];
            C.a11 = cu[6] - cu[1] * cu[1];
            C.a22 = cu[8] - cu[2] * cu[2];
            C.a01 = cu[4] - cu[0] * cu[1];
            C.a02 = cu[5] - cu[0] * cu[2];
            C.a12 = cu[7] - cu[1] * cu[2];
        }
        Vec3<double> nrm = sym3_smallest_eigvec<double>(C);
        const double len = sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        if (prior != nullptr) {
            const double ox = prior[3 * (int64_t)oi], oy = prior[3 * (int64_t)oi + 1], oz = prior[3 * (int64_t)oi + 2];
            if (len == 0.0) nrm = {ox, oy, oz};
            else if (nrm.x * ox + nrm.y * oy + nrm.z * oz < 0.0) nrm = {-nrm.x, -nrm.y, -nrm.z};
        } else if (len == 0.0) {
            nrm = {0.0, 0.0, 1.0};
        }
        normals[3 * (int64_t)oi] = nrm.x;
        normals[3 * (int64_t)oi + 1] = nrm.y;
        normals[3 * (int64_t)oi + 2] = nrm.z;
    }
}

// ---- staged legacy normals, round 2 ------------------------------------------------------------------------------------
// Two kernels instead of one (profiles/r01n_ncu_normals_staged_p64_digest.txt: 4.7 G warp instructions at 19 active lanes,
// instruction-cache misses, a per-thread insertion sort on 10 of 32 lanes):
//  normals_cov2_kernel: one warp per chunk. Staging through b3d_stage2.cuh (bulk copies of whole cells, float32 filter out of
//    shared memory); every lane collects the candidates inside its radius (exact: borderline ones are confirmed in float64),
//    the short lists are ordered by distance with a SORTING NETWORK HELD IN REGISTERS (uniform control flow, sized by the
//    longest list of the warp), the nine raw moments are summed in that order from the float64 points (re-gathered, L1) and
//    the covariance goes to global memory (48 bytes per point). No float64 copy of the candidates in shared memory.
//  normals_eig2_kernel: one thread per point, closed-form eigenvector + orientation against the prior, every lane busy.
// Points that need the k-nearest cut (more than k in-radius neighbours, or more than 32), or whose box takes more than one
// staging batch, are queued (their covariance slot is marked) for normals_queue2_kernel below: one warp per queued point.
#ifndef B3D_NRM2_CAP
#define B3D_NRM2_CAP 472  // + 8 padding slots = 15 groups of 32: four blocks of four warps still fit one SM's shared memory
#endif
#ifndef B3D_NRM2_MIN_BLOCKS
#define B3D_NRM2_MIN_BLOCKS 4
#endif
constexpr int kNrm2Cap = B3D_NRM2_CAP;
constexpr int kNrm2List = 32;
struct alignas(16) Nrm2Smem {
    StageSmem<kNrm2Cap> stage;
    unsigned int keys[32][kNrm2List + 1];  // per lane: (float bits of d2, low 9 bits = candidate slot) of the in-radius candidates (row stride 33 words)
};

#define B3D_CE(a, b)                          \
    {                                         \
        const unsigned int lo_ = min(k[a], k[b]); \
        k[b] = max(k[a], k[b]);               \
        k[a] = lo_;                           \
    }

// Orders the first n keys of a lane's list ascending: into registers, a sorting network (uniform control flow: every lane runs
// it, shorter lists are padded with +inf), back to shared memory. Out of line: one copy of each unrolled network.
__device__ __noinline__ void nrm2_sort_keys16(unsigned int* __restrict__ list, int n) {
    unsigned int k[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) k[j] = j < n ? list[j] : 0xffffffffu;
    B3D_SORTNET_16(B3D_CE)
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (j < n) list[j] = k[j];
}
__device__ __noinline__ void nrm2_sort_keys32(unsigned int* __restrict__ list, int n) {
    unsigned int k[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) k[j] = j < n ? list[j] : 0xffffffffu;
    B3D_SORTNET_32(B3D_CE)
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if (j < n) list[j] = k[j];
}
#undef B3D_CE

__global__ void __launch_bounds__(kNrmBlock, B3D_NRM2_MIN_BLOCKS) normals_cov2_kernel(GridView<double> g, const int32_t* __restrict__ chunk_start,
                                                                                     const int32_t* __restrict__ chunk_off, int B, int n_chunks, int k_nn,
                                                                                     double radius, double r2, double* __restrict__ cov6,
                                                                                     int* __restrict__ todo, int* __restrict__ todo_count, int stats) {
    extern __shared__ __align__(16) unsigned char nrm2_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Nrm2Smem& W = reinterpret_cast<Nrm2Smem*>(nrm2_smem)[warp];
    auto& S = W.stage;
    stage2_init_barrier(&S.mbar);
    uint32_t parity = 0;
    unsigned int* list = W.keys[lane];
    int cur_cloud = -1;
    UnitFrame F = {};
    float r2u = 0.f;   // radius^2 in units^2
    double ru = 0.0;   // radius in units (+ margins)
    // a block owns a contiguous range of chunks (Morton neighbours: overlapping boxes, so the hash slots and the float64 points one
    // warp gathered are in L1 for the next one); its warps interleave inside the range
    const int c_per = (n_chunks + (int)gridDim.x - 1) / (int)gridDim.x;
    const int c_end = min(n_chunks, ((int)blockIdx.x + 1) * c_per);
    for (int c = (int)blockIdx.x * c_per + warp; c < c_end; c += kNrmBlock / 32) {
        // cloud of this chunk: last b with chunk_off[b] <= c
        int lo_b = 0, hi_b = B;
        while (hi_b - lo_b > 1) {
            const int m = (lo_b + hi_b) >> 1;
            if (chunk_off[m] <= c) lo_b = m; else hi_b = m;
        }
        const int cloud = lo_b;
        if (cloud != cur_cloud) {
            F = unit_frame(g.lat[cloud], g.shift);
            r2u = (float)(r2 * F.per_m * F.per_m);
            ru = radius * F.per_m * (1.0 + 1e-12) + 2.0;
            cur_cloud = cloud;
        }
        const int32_t i = chunk_start[c] + lane;
        const bool valid = i < chunk_start[c + 1];
        double qx = 0, qy = 0, qz = 0;
        int oi = 0;
        if (valid) {
            const double4 q = ld_point(g.pts + i);
            qx = q.x; qy = q.y; qz = q.z;
            oi = point_index(q);
        }
        const double ux = unit_coord_of_query(qx, F.ox, F.per_m), uy = unit_coord_of_query(qy, F.oy, F.per_m), uz = unit_coord_of_query(qz, F.oz, F.per_m);
        // The chunk is searched as one group of 32 lanes; a group whose box does not fit ONE staging batch (dense clouds: the
        // neighbour lists refer to slots of the batch) is split in halves and retried, down to groups of four lanes; what still
        // does not fit is queued (normals_queue2_kernel).
        unsigned int pending = __ballot_sync(0xffffffffu, valid);
        int width = 32;
        while (pending != 0u) {
            const int g0 = (__ffs(pending) - 1) & ~(width - 1);
            const unsigned int gmask = width == 32 ? 0xffffffffu : (((1u << width) - 1u) << g0);
            const bool active = valid && ((gmask >> lane) & 1u);
            // the group's box in fixed-point units: every query's ball, margins for the floor() of the records and the roundings here
            int lox = active ? unit_floor_clamped(ux - ru) : 0x7fffffff, loy = active ? unit_floor_clamped(uy - ru) : 0x7fffffff,
                loz = active ? unit_floor_clamped(uz - ru) : 0x7fffffff;
            int hix = active ? unit_ceil_clamped(ux + ru) : (int)0x80000000, hiy = active ? unit_ceil_clamped(uy + ru) : (int)0x80000000,
                hiz = active ? unit_ceil_clamped(uz + ru) : (int)0x80000000;
            lox = max(__reduce_min_sync(0xffffffffu, lox), 0); loy = max(__reduce_min_sync(0xffffffffu, loy), 0); loz = max(__reduce_min_sync(0xffffffffu, loz), 0);
            hix = __reduce_max_sync(0xffffffffu, hix); hiy = __reduce_max_sync(0xffffffffu, hiy); hiz = __reduce_max_sync(0xffffffffu, hiz);
            const int ccx = (int)(((long long)lox + hix) >> 1), ccy = (int)(((long long)loy + hiy) >> 1), ccz = (int)(((long long)loz + hiz) >> 1);
            const float qox = (float)(ux - (double)ccx), qoy = (float)(uy - (double)ccy), qoz = (float)(uz - (double)ccz);
            const float fx = -2.0f * qox, fy = -2.0f * qoy, fz = -2.0f * qoz;
            const float qq = fmaf(qoz, qoz, fmaf(qoy, qoy, qox * qox));
            const float H = fmaxf(fmaxf((float)(hix - ccx), (float)(hiy - ccy)), (float)(hiz - ccz)) + 1.0f;
            // |d2f - d2| <= rounding of the scanned t (half of stage2_band would do) + of |q|^2 + of the add + of r2u itself
            const float band = stage2_band(H) + 4.0e-7f * r2u;
            int n = 0;
            int kept_all = 0;
            const float2 f2x = make_float2(fx, fx), f2y = make_float2(fy, fy), f2z = make_float2(fz, fz);
            // one candidate: accept when inside the radius (borderline ones confirmed in float64), append its key
            auto take = [&](float t, int j) {
                const float d2f = t + qq;
                bool acc = d2f < r2u - band;
                if (!acc && d2f <= r2u + band) {  // borderline (rare): decide in float64
                    const double4 pj = ld_point(g.pts + S.pos[j]);
                    acc = dist2<double>(qx - pj.x, qy - pj.y, qz - pj.z) < r2;
                }
                // ascending keys = (d2, scan order): the float bits of d2 with the low 9 bits replaced by the slot. Branch-free
                // append: rejected candidates (and overflowing lists) write the spare entry at the end of the row
                list[acc && n < kNrm2List ? n : kNrm2List] = (__float_as_uint(fmaxf(d2f, 0.0f)) & ~511u) | (unsigned int)j;
                n += acc ? 1 : 0;
            };
            auto scan = [&](int kept) {
                kept_all = kept;
                if (!active) return;
                for (int j = 0; j < kept; j += 4) {  // a ragged tail meets padding candidates at +inf
                    const float4 t = quad_t(S.buf, j, f2x, f2y, f2z);
                    take(t.x, j);
                    take(t.y, j + 1);
                    take(t.z, j + 2);
                    take(t.w, j + 3);
                }
            };
            const int nb = stage2_run<kNrm2Cap>(g, F, cloud, lox, loy, loz, hix, hiy, hiz, S, parity, scan);
#ifdef B3D_NRM2_STATS
            if (stats && lane == 0) {
                atomicAdd(&g_nrm_stats[0], 1ull);
                if (nb > 1 || nb < 0) atomicAdd(&g_nrm_stats[1], 1ull); else atomicAdd(&g_nrm_stats[3], (unsigned long long)kept_all);
                atomicAdd(&g_nrm_stats[4], (unsigned long long)__popc(pending & gmask));
            }
#endif
            // one batch holds every candidate (slots fit 9 bits); nb == 0: no point at all in the box (cannot happen: the queries are points)
            const bool group_ok = nb <= 1 && nb >= 0 && kept_all <= 512;
            if (!group_ok && width > 4) {
                width >>= 1;  // retry this part of the chunk in two halves (warp-uniform)
                continue;
            }
            pending &= ~gmask;
            const bool punt = active && (!group_ok || n > k_nn || n > kNrm2List);
            if (punt) {
#ifdef B3D_NRM2_STATS
                if (stats) atomicAdd(&g_nrm_stats[2], 1ull);
#endif
                todo[atomicAdd(todo_count, 1)] = i;
                cov6[6 * (int64_t)oi] = __longlong_as_double(0x7ff8000000000b3dll);  // marked: the eigen kernel skips it
            }
            const bool work = active && !punt;
            if (!work) n = 0;
            const int nmax = __reduce_max_sync(0xffffffffu, n);
            if (nmax > 16) nrm2_sort_keys32(list, n);
            else if (nmax > 2) nrm2_sort_keys16(list, n);
            else if (n == 2 && list[1] < list[0]) { const unsigned int t = list[0]; list[0] = list[1]; list[1] = t; }
            // ---- raw-moment covariance over the neighbour set in distance order (float64 points re-gathered) ----------------
            double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < n; ++j) {
                const double4 pj = ld_point(g.pts + S.pos[list[j] & 511u]);
                const double x = pj.x, y = pj.y, z = pj.z;
                cu[0] += x; cu[1] += y; cu[2] += z;
                cu[3] += x * x; cu[4] += x * y; cu[5] += x * z;
                cu[6] += y * y; cu[7] += y * z; cu[8] += z * z;
            }
            if (work) {
                double C[6] = {1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
                if (n >= 3) {
                    const double cn = (double)n;
#pragma unroll
                    for (int j = 0; j < 9; ++j) cu[j] /= cn;
                    C[0] = cu[3] - cu[0] * cu[0];
                    C[1] = cu[4] - cu[0] * cu[1];
                    C[2] = cu[5] - cu[0] * cu[2];
                    C[3] = cu[6] - cu[1] * cu[1];
                    C[4] = cu[7] - cu[1] * cu[2];
                    C[5] = cu[8] - cu[2] * cu[2];
                }
                double* out = cov6 + 6 * (int64_t)oi;
#pragma unroll
                for (int j = 0; j < 6; ++j) out[j] = C[j];
            }
            __syncwarp();
        }
    }
}

// ---- the queued points of the staged path, one WARP per point ----------------------------------------------------------------
// A point is queued by normals_cov2_kernel when its radius holds more neighbours than k (or than 32), i.e. it needs the k NEAREST of
// them by the library's (d2, index) order, or when its group's box did not fit one staging batch. The per-lane kernel redid such
// points one thread each: on an 8 MP stereo cloud 0.2 % of the points (dense patches) cost 0.38 ms per cloud -- a handful of threads
// each walking ~1000 candidates behind dependent loads. Here the warp stages the point's own ball batch by batch (b3d_stage2.cuh),
// every lane takes candidates 32 apart (float pre-test, exact float64 distance for everything that may lie inside the radius), the
// in-radius ones are compacted into the warp's key rows as {d2, sorted position, index}; whenever the rows fill up, and at the end,
// the entries are RANKED BY COUNTING (rank = number of smaller entries; broadcast reads, no sort) and the k smallest move to the
// front in rank order. Lane 0 then sums the nine raw moments in that order: the arithmetic of the per-lane kernel, bit for bit.
constexpr int kNrmQCap = 232;  // 232 + 32 entries of 16 bytes = the 32 x 33 words of a warp's key rows
struct NrmQEntry {
    double d2;
    int pos, idx;
};
static_assert(sizeof(NrmQEntry) == 16 && (kNrmQCap + 32) * 4 <= 32 * (kNrm2List + 1), "the entries alias the key rows");

__global__ void __launch_bounds__(kNrmBlock) normals_queue2_kernel(GridView<double> g, const uint64_t* __restrict__ keys, const int32_t* __restrict__ off, int k_nn,
                                                                  double radius, double r2, int rmax, const double* __restrict__ prior,
                                                                  double* __restrict__ normals, const int* __restrict__ todo, const int* __restrict__ todo_count) {
    extern __shared__ __align__(16) unsigned char nrm2_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Nrm2Smem& W = reinterpret_cast<Nrm2Smem*>(nrm2_smem)[warp];
    auto& S = W.stage;
    stage2_init_barrier(&S.mbar);
    uint32_t parity = 0;
    NrmQEntry* ent = reinterpret_cast<NrmQEntry*>(&W.keys[0][0]);
    NrmQEntry* best = ent + kNrmQCap;  // the k smallest of a ranking, in rank order
    const int n_work = *todo_count;
    for (int t = (int)blockIdx.x * (kNrmBlock / 32) + warp; t < n_work; t += (int)gridDim.x * (kNrmBlock / 32)) {
        const int i = todo[t];
        const double4 q = ld_point(g.pts + i);
        const int cloud = (int)(keys[i] >> g.shift);
        const UnitFrame F = unit_frame(g.lat[cloud], g.shift);
        const float r2u = (float)(r2 * F.per_m * F.per_m);
        const double ru = radius * F.per_m * (1.0 + 1e-12) + 2.0;
        const double ux = unit_coord_of_query(q.x, F.ox, F.per_m), uy = unit_coord_of_query(q.y, F.oy, F.per_m), uz = unit_coord_of_query(q.z, F.oz, F.per_m);
        const int lox = max(unit_floor_clamped(ux - ru), 0), loy = max(unit_floor_clamped(uy - ru), 0), loz = max(unit_floor_clamped(uz - ru), 0);
        const int hix = unit_ceil_clamped(ux + ru), hiy = unit_ceil_clamped(uy + ru), hiz = unit_ceil_clamped(uz + ru);
        const int ccx = (int)(((long long)lox + hix) >> 1), ccy = (int)(((long long)loy + hiy) >> 1), ccz = (int)(((long long)loz + hiz) >> 1);
        const float qox = (float)(ux - (double)ccx), qoy = (float)(uy - (double)ccy), qoz = (float)(uz - (double)ccz);
        const float fx = -2.0f * qox, fy = -2.0f * qoy, fz = -2.0f * qoz;
        const float qq = fmaf(qoz, qoz, fmaf(qoy, qoy, qox * qox));
        const float H = fmaxf(fmaxf((float)(hix - ccx), (float)(hiy - ccy)), (float)(hiz - ccz)) + 1.0f;
        const float band = stage2_band(H) + 4.0e-7f * r2u;  // as in normals_cov2_kernel
        const float2 f2x = make_float2(fx, fx), f2y = make_float2(fy, fy), f2z = make_float2(fz, fz);
        int n = 0;  // entries held (warp-uniform)
        // ranks the n entries and moves the k smallest to the front, in (d2, index) order
        auto keep_k_smallest = [&]() {
            __syncwarp();
            for (int e = lane; e < n; e += 32) {
                const NrmQEntry a = ent[e];
                int rank = 0;
                for (int j = 0; j < n; ++j) {
                    const double dj = ent[j].d2;
                    rank += (dj < a.d2 || (dj == a.d2 && ent[j].idx < a.idx)) ? 1 : 0;
                }
                if (rank < k_nn) best[rank] = a;
            }
            __syncwarp();
            n = min(n, k_nn);
            if (lane < n) ent[lane] = best[lane];
            __syncwarp();
        };
        auto scan = [&](int kept) {
            for (int j0 = 0; j0 < kept; j0 += 32) {
                if (n + 32 > kNrmQCap) keep_k_smallest();
                const int j = j0 + lane;
                bool acc = false;
                NrmQEntry a{0.0, 0, 0};
                if (j < kept && cand_t(S.buf, j, f2x, f2y, f2z) + qq <= r2u + band) {  // may lie inside the radius: exact distance
                    a.pos = S.pos[j];
                    const double4 pj = ld_point(g.pts + a.pos);
                    a.d2 = dist2<double>(q.x - pj.x, q.y - pj.y, q.z - pj.z);
                    a.idx = point_index(pj);
                    acc = a.d2 < r2;
                }
                const unsigned int m = __ballot_sync(0xffffffffu, acc);
                if (acc) ent[n + __popc(m & ((1u << lane) - 1u))] = a;
                n += __popc(m);
            }
            __syncwarp();
        };
        const int nb = stage2_run<kNrm2Cap>(g, F, cloud, lox, loy, loz, hix, hiy, hiz, S, parity, scan);
        double cu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int c = 0;
        if (nb < 0) {
            // a box the staging cannot take (one cell beyond the buffer, > 1024 cells): the per-lane walk, on lane 0
            if (lane == 0) {
                TopK<double, 32> tk;
                knn_hybrid_query<double, 32>(g, off, cloud, q.x, q.y, q.z, k_nn, true, r2, rmax, tk);
                c = tk.n;
                for (int j = 0; j < c; ++j) best[j].pos = tk.pos[j];
            }
        } else {
            keep_k_smallest();
            c = n;
            if (lane < c) best[lane].pos = ent[lane].pos;
        }
        __syncwarp();
        if (lane == 0) {
            for (int j = 0; j < c; ++j) {
                const double4 pj = ld_point(g.pts + best[j].pos);
                const double x = pj.x, y = pj.y, z = pj.z;
                cu[0] += x; cu[1] += y; cu[2] += z;
                cu[3] += x * x; cu[4] += x * y; cu[5] += x * z;
                cu[6] += y * y; cu[7] += y * z; cu[8] += z * z;
            }
            Sym3<double> C{1.0, 0.0, 0.0, 1.0, 0.0, 1.0};
            if (c >= 3) {
                const double cn = (double)c;
#pragma unroll
                for (int j = 0; j < 9; ++j) cu[j] /= cn;
                C.a00 = cu[3] - cu[0] * cu[0];
                C.a11 = cu[6] - cu[1] * cu[1];
                C.a22 = cu[8] - cu[2] * cu[2];
                C.a01 = cu[4] - cu[0] * cu[1];
                C.a02 = cu[5] - cu[0] * cu[2];
                C.a12 = cu[7] - cu[1] * cu[2];
            }
            Vec3<double> nrm = sym3_smallest_eigvec<double>(C);
            const int64_t oi = point_index(q);
            const double len = sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
            if (prior != nullptr) {
                const double ox = prior[3 * oi], oy = prior[3 * oi + 1], oz = prior[3 * oi + 2];
                if (len == 0.0) nrm = {ox, oy, oz};
                else if (nrm.x * ox + nrm.y * oy + nrm.z * oz < 0.0) nrm = {-nrm.x, -nrm.y, -nrm.z};
            } else if (len == 0.0) {
                nrm = {0.0, 0.0, 1.0};
            }
            normals[3 * oi] = nrm.x;
            normals[3 * oi + 1] = nrm.y;
            normals[3 * oi + 2] = nrm.z;
        }
        __syncwarp();
    }
}

// cov6[i] = {a00 a01 a02 a11 a12 a22} of point i (original index) -> unit normal, oriented against the prior
__global__ void __launch_bounds__(256) normals_eig2_kernel(const double* __restrict__ cov6, int64_t n, const double* __restrict__ prior,
                                                           double* __restrict__ normals) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* c = cov6 + 6 * i;
        const double a00 = c[0];
        if (__double_as_longlong(a00) == 0x7ff8000000000b3dll) continue;  // queued: normals_queue2_kernel writes this normal
        Sym3<double> C{a00, c[1], c[2], c[3], c[4], c[5]};
        Vec3<double> nrm = sym3_smallest_eigvec<double>(C);
        const double len = sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        if (prior != nullptr) {
            const double ox = prior[3 * i], oy = prior[3 * i + 1], oz = prior[3 * i + 2];
            if (len == 0.0) nrm = {ox, oy, oz};
            else if (nrm.x * ox + nrm.y * oy + nrm.z * oz < 0.0) nrm = {-nrm.x, -nrm.y, -nrm.z};
        } else if (len == 0.0) {
            nrm = {0.0, 0.0, 1.0};
        }
        normals[3 * i] = nrm.x;
        normals[3 * i + 1] = nrm.y;
        normals[3 * i + 2] = nrm.z;
    }
}

// ---- k nearest export (b3d_knn_hybrid): arbitrary queries against a grid -------------------------------------------
template <typename T, int KMAX>
__global__ void __launch_bounds__(128) knn_export_kernel(GridView<T> g, const int32_t* __restrict__ off, const T* __restrict__ queries, int64_t nq,
                                                         int k, bool use_radius, T r2, int rmax, int32_t* __restrict__ idx, T* __restrict__ d2,
                                                         int32_t* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    TopK<T, KMAX> tk;
    knn_hybrid_query<T, KMAX>(g, off, 0, queries[3 * i], queries[3 * i + 1], queries[3 * i + 2], k, use_radius, r2, rmax, tk);
    for (int j = 0; j < k; ++j) {
        idx[i * k + j] = j < tk.n ? point_index(ld_point(g.pts + tk.pos[j])) : -1;
        if (d2 != nullptr) d2[i * k + j] = j < tk.n ? tk.d2[j] : T(0);
    }
    if (cnt != nullptr) cnt[i] = tk.n;
}

// ---- statistical outlier: mean distance to the nb nearest (self included) -------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(128) knn_mean_distance_kernel(GridView<double> g, const int32_t* __restrict__ off, int nb, int rmax,
                                                                double* __restrict__ avg) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const double4 q = ld_point(g.pts + pos);
    TopK<double, KMAX> tk;
    knn_hybrid_query<double, KMAX>(g, off, 0, q.x, q.y, q.z, nb, false, 0.0, rmax, tk);
    double mean = -1.0;
    if (tk.n > 0) {
        double s = 0;
        for (int j = 0; j < tk.n; ++j) s += sqrt(tk.d2[j]);
        mean = s / (double)tk.n;
    }
    avg[point_index(q)] = mean;
}

// deterministic two-stage reductions over avg[]: stage 1 -> per-block partial {sum, count}; stage 2 (one block)
constexpr int kRedBlocks = 512;
__global__ void __launch_bounds__(256) stat_sum_kernel(const double* __restrict__ avg, int64_t n, const double* __restrict__ mean_in,
                                                       double* __restrict__ partial) {
    // mean_in == NULL: sum of valid avg and their count; else: sum of squared deviations from *mean_in
    const double mu = mean_in ? mean_in[0] : 0.0;
    double s = 0, c = 0;
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(n, lo + chunk);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double a = avg[i];
        if (a > 0) {
            s += mean_in ? (a - mu) * (a - mu) : a;
            c += 1.0;
        }
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = s;
    sh[1][threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = sh[0][0];
        partial[2 * blockIdx.x + 1] = sh[1][0];
    }
}
// stats: [0] mean, [1] valid count, [2] threshold
__global__ void stat_final_kernel(const double* __restrict__ partial, int nblocks, int stage, double std_ratio, double* __restrict__ stats) {
    if (threadIdx.x != 0) return;
    double s = 0, c = 0;
    for (int b = 0; b < nblocks; ++b) {
        s += partial[2 * b];
        c += partial[2 * b + 1];
    }
    if (stage == 0) {
        stats[0] = c > 0 ? s / c : 0.0;
        stats[1] = c;
    } else {
        const double sd = sqrt(s / (stats[1] - 1.0));
        stats[2] = stats[0] + std_ratio * sd;
    }
}
__global__ void __launch_bounds__(256) stat_keep_kernel(const double* __restrict__ avg, int64_t n, const double* __restrict__ stats,
                                                        uint8_t* __restrict__ keep) {
    const double thr = stats[2];
    const bool any = stats[1] > 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = avg[i];
        keep[i] = (any && a > 0 && a < thr) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128) radius_count_kernel(GridView<double> g, double r2, int rmax, int nb_points, uint8_t* __restrict__ keep) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const double4 q = ld_point(g.pts + pos);
    const int c = count_within_query<double>(g, 0, q.x, q.y, q.z, r2, rmax);
    keep[point_index(q)] = c > nb_points ? 1 : 0;
}

struct KeepPred {
    const uint8_t* keep;
    __device__ __forceinline__ bool operator()(int64_t i) const { return keep[i] != 0; }
};
struct KeepEmit {
    int64_t* out;
    __device__ __forceinline__ void operator()(int64_t i, int64_t slot) const {
        if (out != nullptr) out[slot] = i;
    }
};

__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, int cols,
                                                          double* __restrict__ dst) {
    const int64_t total = n * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / cols;
        const int c = (int)(t - r * cols);
        dst[t] = src[idx[r] * cols + c];
    }
}

// GICP: C = R diag(eps,1,1) R^T with R the rotation taking e1 onto the normal
__global__ void __launch_bounds__(256) cov_from_normals_kernel(const double* __restrict__ normals, int64_t n, double eps, double* __restrict__ cov) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x0 = normals[3 * i], x1 = normals[3 * i + 1], x2 = normals[3 * i + 2];
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        const double c = x0;
        if (!(c < -0.99)) {
            const double v[3] = {0.0, -x2, x1};
            const double sv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
            const double f = 1.0 / (1.0 + c);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    double s2 = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) s2 += sv[3 * r + k] * sv[3 * k + cc];
                    R[3 * r + cc] = (r == cc ? 1.0 : 0.0) + sv[3 * r + cc] + s2 * f;
                }
        }
        const double D[3] = {eps, 1.0, 1.0};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += R[3 * r + k] * D[k] * R[3 * cc + k];
                cov[9 * i + 3 * r + cc] = s;
            }
    }
}

}  // namespace

// normals for a batch of clouds (written by ORIGINAL point index). radius <= 0: pure k nearest.
template <typename T, bool TENSOR>
int estimate_normals_batch(b3d_ctx* ctx, const T* xyz, const Segments& seg, int max_nn, double radius, const T* prior, T* normals,
                           Grid<T>* reuse_grid, int reuse_rmax) {
    B3D_REQUIRE(max_nn >= 1 && max_nn <= 64, "max_nn must be in [1, 64] (got %d)", max_nn);
    Grid<T> local;
    Grid<T>* grid = reuse_grid;
    int rmax = reuse_rmax;
    if (!grid) {
        grid = &local;
        B3D_TRY(build_search_grid<T>(ctx, xyz, seg, max_nn, radius, grid, &rmax));
    }
    const bool use_radius = radius > 0;
    const T r = (T)radius;
    const T r2 = r * r;
    const int n = (int)grid->sort.n;
    const int* todo = nullptr;
    const int* todo_count = nullptr;
    DevBuf<int> todo_buf, todo_count_buf;
    int blocks = (n + 127) / 128;
    if constexpr (std::is_same<T, double>::value && !TENSOR) {
        if (use_radius && max_nn <= kNrmList) {
            // staged fast path; the queued points are redone by normals_queue2_kernel (round-2 kernels) or by the per-lane kernel below (round-1 kernel)
            QueryChunks qc;
            B3D_TRY(chunks_from_grid(ctx, *grid, seg.off, seg.off_h, &qc));
            B3D_TRY(todo_buf.alloc(ctx, (size_t)n));
            B3D_TRY(todo_count_buf.alloc(ctx, 1));
            B3D_CUDA(cudaMemsetAsync(todo_count_buf.p, 0, sizeof(int), ctx->stream));
            static const bool v1 = getenv("B3D_NRM_V1") != nullptr;  // the round-1 kernel, kept for A/B runs
            if (!v1 && grid->rec.p != nullptr) {
                DevBuf<double> cov6;
                B3D_TRY(cov6.alloc(ctx, (size_t)n * 6));
                // contiguous chunk ranges per block: many more blocks than resident slots, so that the ranges' uneven costs even out
                const int sblocks = std::max(1, std::min((qc.n_chunks + kNrmBlock / 32 - 1) / (kNrmBlock / 32), ctx->sm_count * 64));
                const size_t smem = sizeof(Nrm2Smem) * (kNrmBlock / 32);
                B3D_CUDA(cudaFuncSetAttribute(normals_cov2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                B3D_LAUNCH(ctx, normals_cov2_kernel, sblocks, kNrmBlock, smem, grid->view(), qc.chunk_start.p, qc.chunk_off.p, seg.B, qc.n_chunks, max_nn,
                           radius, r2, cov6.p, todo_buf.p, todo_count_buf.p, getenv("B3D_ICP_STATS") ? 1 : 0);
                B3D_LAUNCH(ctx, normals_eig2_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, cov6.p, (int64_t)n, prior, normals);
                // the queued points (k-nearest cuts, boxes beyond one batch): one warp each
                B3D_CUDA(cudaFuncSetAttribute(normals_queue2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                B3D_LAUNCH(ctx, normals_queue2_kernel, std::max(1, std::min(sblocks, ctx->sm_count * 4)), kNrmBlock, smem, grid->view(), grid->sort.keys.p,
                           seg.off, max_nn, radius, r2, rmax, prior, normals, todo_buf.p, todo_count_buf.p);
                return B3D_OK;
            } else {
                const int sblocks = std::max(1, std::min((qc.n_chunks + kNrmBlock / 32 - 1) / (kNrmBlock / 32), ctx->sm_count * 32));
                const size_t smem = sizeof(NrmWarpSmem) * (kNrmBlock / 32);
                B3D_CUDA(cudaFuncSetAttribute(normals_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                B3D_LAUNCH(ctx, normals_staged_kernel, sblocks, kNrmBlock, smem, grid->view(), qc.q, qc.chunk_start.p, qc.chunk_off.p, seg.B, qc.n_chunks,
                           max_nn, radius, r2, prior, normals, todo_buf.p, todo_count_buf.p, getenv("B3D_ICP_STATS") ? 1 : 0);
            }
            todo = todo_buf.p;
            todo_count = todo_count_buf.p;
            blocks = std::min(blocks, ctx->sm_count * 4);  // the queue is short; the kernel strides over it
        }
    }
    if (max_nn <= 32) {
        B3D_LAUNCH(ctx, (normals_kernel<T, 32, TENSOR>), blocks, 128, 0, grid->view(), grid->sort.keys.p, seg.off, max_nn, use_radius, r2, rmax, prior,
                   normals, todo, todo_count);
    } else {
        B3D_LAUNCH(ctx, (normals_kernel<T, 64, TENSOR>), blocks, 128, 0, grid->view(), grid->sort.keys.p, seg.off, max_nn, use_radius, r2, rmax, prior,
                   normals, todo, todo_count);
    }
    return B3D_OK;
}
template int estimate_normals_batch<double, false>(b3d_ctx*, const double*, const Segments&, int, double, const double*, double*, Grid<double>*, int);
template int estimate_normals_batch<float, false>(b3d_ctx*, const float*, const Segments&, int, double, const float*, float*, Grid<float>*, int);
template int estimate_normals_batch<float, true>(b3d_ctx*, const float*, const Segments&, int, double, const float*, float*, Grid<float>*, int);

}  // namespace b3d

using namespace b3d;

struct b3d_grid {
    int is_f64 = 1;
    Grid<double> gd;
    Grid<float> gf;
    DevBuf<int32_t> off;
    Segments seg;
    int rmax_hint = kMaxRing;
    double radius_hint = 0;
};

extern "C" {

int b3d_debug_normals_stats(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_nrm_stats, 8 * sizeof(unsigned long long)) != cudaSuccess) return B3D_E_CUDA;
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(g_nrm_stats, z, sizeof(z)) != cudaSuccess) return B3D_E_CUDA;
    }
    return B3D_OK;
}

int b3d_estimate_normals_legacy(b3d_ctx* ctx, const double* xyz, int64_t n, int max_nn, double radius, const double* prior, double* normals) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && normals, "b3d_estimate_normals_legacy: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    return estimate_normals_batch<double, false>(ctx, xyz, seg, max_nn, radius, prior, normals, nullptr, 0);
}

int b3d_estimate_normals_tensor(b3d_ctx* ctx, const float* xyz, int64_t n, int max_nn, float radius, float* normals) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && normals, "b3d_estimate_normals_tensor: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    return estimate_normals_batch<float, true>(ctx, xyz, seg, max_nn, (double)radius, nullptr, normals, nullptr, 0);
}

int b3d_covariances_from_normals(b3d_ctx* ctx, const double* normals, int64_t n, double eps, double* cov) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(normals && cov, "b3d_covariances_from_normals: NULL buffer");
    B3D_TRY(ctx->bind());
    B3D_LAUNCH(ctx, cov_from_normals_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, normals, n, eps, cov);
    return B3D_OK;
}

static int finish_keep(b3d_ctx* ctx, const uint8_t* keep, int64_t n, int64_t* kept_idx, int64_t* n_kept_h) {
    DevBuf<int64_t> total;
    B3D_TRY(total.alloc(ctx, 1));
    B3D_TRY(compact(ctx, KeepPred{keep}, KeepEmit{kept_idx}, n, total.p));
    return ctx->download(n_kept_h, total.p, sizeof(int64_t));
}

int b3d_statistical_outlier(b3d_ctx* ctx, const double* xyz, int64_t n, int nb_neighbors, double std_ratio, uint8_t* keep, int64_t* kept_idx,
                            int64_t* n_kept_h) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n_kept_h != nullptr, "n_kept_h is NULL");
    *n_kept_h = 0;
    B3D_REQUIRE(nb_neighbors >= 1 && std_ratio > 0.0,
                "Illegal input parameters, the number of neighbors and standard deviation ratio must be positive.");
    B3D_REQUIRE(nb_neighbors <= 64, "nb_neighbors must be <= 64 (got %d)", nb_neighbors);
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && keep, "b3d_statistical_outlier: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    Grid<double> grid;
    int rmax = kMaxRing;
    B3D_TRY(build_search_grid<double>(ctx, xyz, seg, nb_neighbors, 0.0, &grid, &rmax));
    DevBuf<double> avg, partial, stats;
    B3D_TRY(avg.alloc(ctx, n));
    B3D_TRY(partial.alloc(ctx, 2 * kRedBlocks));
    B3D_TRY(stats.alloc(ctx, 4));
    const int blocks = (int)((n + 127) / 128);
    if (nb_neighbors <= 32) {
        B3D_LAUNCH(ctx, knn_mean_distance_kernel<32>, blocks, 128, 0, grid.view(), seg.off, nb_neighbors, rmax, avg.p);
    } else {
        B3D_LAUNCH(ctx, knn_mean_distance_kernel<64>, blocks, 128, 0, grid.view(), seg.off, nb_neighbors, rmax, avg.p);
    }
    B3D_LAUNCH(ctx, stat_sum_kernel, kRedBlocks, 256, 0, avg.p, n, (const double*)nullptr, partial.p);
    B3D_LAUNCH(ctx, stat_final_kernel, 1, 32, 0, partial.p, kRedBlocks, 0, std_ratio, stats.p);
    B3D_LAUNCH(ctx, stat_sum_kernel, kRedBlocks, 256, 0, avg.p, n, (const double*)stats.p, partial.p);
    B3D_LAUNCH(ctx, stat_final_kernel, 1, 32, 0, partial.p, kRedBlocks, 1, std_ratio, stats.p);
    B3D_LAUNCH(ctx, stat_keep_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, avg.p, n, stats.p, keep);
    return finish_keep(ctx, keep, n, kept_idx, n_kept_h);
}

int b3d_radius_outlier(b3d_ctx* ctx, const double* xyz, int64_t n, int nb_points, double radius, uint8_t* keep, int64_t* kept_idx,
                       int64_t* n_kept_h) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n_kept_h != nullptr, "n_kept_h is NULL");
    *n_kept_h = 0;
    B3D_REQUIRE(nb_points >= 1 && radius > 0.0, "Illegal input parameters, number of points and radius must be positive.");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && keep, "b3d_radius_outlier: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    Grid<double> grid;
    const double cell = radius * 1.001;
    B3D_TRY(grid_build<double>(ctx, xyz, seg, cell, nullptr, &grid));
    const int rmax = rings_for_radius(radius, cell);
    B3D_LAUNCH(ctx, radius_count_kernel, (int)((n + 127) / 128), 128, 0, grid.view(), radius * radius, rmax, nb_points, keep);
    return finish_keep(ctx, keep, n, kept_idx, n_kept_h);
}

int b3d_gather_rows_f64(b3d_ctx* ctx, const double* src, const int64_t* kept_idx, int64_t n_kept, int cols, double* dst) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n_kept >= 0 && cols >= 1, "b3d_gather_rows_f64: bad sizes");
    if (n_kept == 0) return B3D_OK;
    B3D_REQUIRE(src && kept_idx && dst, "b3d_gather_rows_f64: NULL buffer");
    B3D_TRY(ctx->bind());
    B3D_LAUNCH(ctx, gather_rows_kernel, ctx->grid_for(n_kept * cols, 256, 1, 8), 256, 0, src, kept_idx, n_kept, cols, dst);
    return B3D_OK;
}

// ---- explicit grid objects (b3d_grid_build / b3d_knn_hybrid) ---------------------------------------------------------
int b3d_grid_build(b3d_ctx* ctx, const void* xyz, int64_t n, int is_f64, double cell_size, int k_hint, double radius_hint, b3d_grid** out) {
    B3D_REQUIRE(ctx != nullptr && out != nullptr, "b3d_grid_build: NULL argument");
    *out = nullptr;
    B3D_REQUIRE(n > 0, "b3d_grid_build: empty cloud");
    B3D_REQUIRE(xyz != nullptr, "b3d_grid_build: NULL buffer");
    B3D_TRY(ctx->bind());
    b3d_grid* g = new b3d_grid();
    g->is_f64 = is_f64 ? 1 : 0;
    g->radius_hint = radius_hint;
    int rc = single_segment(ctx, n, &g->off, &g->seg);
    if (rc == B3D_OK) {
        if (cell_size > 0) {
            rc = is_f64 ? grid_build<double>(ctx, (const double*)xyz, g->seg, cell_size, nullptr, &g->gd)
                        : grid_build<float>(ctx, (const float*)xyz, g->seg, cell_size, nullptr, &g->gf);
        } else {
            rc = is_f64 ? build_search_grid<double>(ctx, (const double*)xyz, g->seg, k_hint, radius_hint, &g->gd, &g->rmax_hint)
                        : build_search_grid<float>(ctx, (const float*)xyz, g->seg, k_hint, radius_hint, &g->gf, &g->rmax_hint);
        }
    }
    if (rc != B3D_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return B3D_OK;
}

int b3d_grid_destroy(b3d_ctx* ctx, b3d_grid* grid) {
    if (!grid) return B3D_OK;
    if (ctx) ctx->bind();
    delete grid;
    return B3D_OK;
}

int b3d_grid_info(b3d_grid* grid, int64_t* n_h, int64_t* n_cells_h, double* cell_size_h) {
    B3D_REQUIRE(grid != nullptr, "grid is NULL");
    const SpatialSort& s = grid->is_f64 ? grid->gd.sort : grid->gf.sort;
    if (n_h) *n_h = s.n;
    if (n_cells_h) *n_cells_h = s.n_runs;
    if (cell_size_h) *cell_size_h = grid->is_f64 ? grid->gd.cell : grid->gf.cell;
    return B3D_OK;
}

int b3d_knn_hybrid(b3d_ctx* ctx, b3d_grid* grid, const void* queries, int64_t nq, int k, double radius, int32_t* idx, void* d2, int32_t* cnt) {
    B3D_REQUIRE(ctx != nullptr && grid != nullptr, "b3d_knn_hybrid: NULL argument");
    B3D_REQUIRE(k >= 1 && k <= 64, "k must be in [1, 64] (got %d)", k);
    B3D_REQUIRE(nq >= 0, "negative query count");
    if (nq == 0) return B3D_OK;
    B3D_REQUIRE(queries && idx, "b3d_knn_hybrid: NULL buffer");
    B3D_TRY(ctx->bind());
    const bool use_radius = radius > 0;
    const int blocks = (int)((nq + 127) / 128);
    if (grid->is_f64) {
        const int rmax = use_radius ? rings_for_radius(radius, grid->gd.cell) : kMaxRing;
        const double r2 = radius * radius;
        if (k <= 32) {
            B3D_LAUNCH(ctx, (knn_export_kernel<double, 32>), blocks, 128, 0, grid->gd.view(), grid->seg.off, (const double*)queries, nq, k, use_radius, r2,
                       rmax, idx, (double*)d2, cnt);
        } else {
            B3D_LAUNCH(ctx, (knn_export_kernel<double, 64>), blocks, 128, 0, grid->gd.view(), grid->seg.off, (const double*)queries, nq, k, use_radius, r2,
                       rmax, idx, (double*)d2, cnt);
        }
    } else {
        const int rmax = use_radius ? rings_for_radius(radius, grid->gf.cell) : kMaxRing;
        const float r = (float)radius;
        const float r2 = r * r;
        if (k <= 32) {
            B3D_LAUNCH(ctx, (knn_export_kernel<float, 32>), blocks, 128, 0, grid->gf.view(), grid->seg.off, (const float*)queries, nq, k, use_radius, r2,
                       rmax, idx, (float*)d2, cnt);
        } else {
            B3D_LAUNCH(ctx, (knn_export_kernel<float, 64>), blocks, 128, 0, grid->gf.view(), grid->seg.off, (const float*)queries, nq, k, use_radius, r2,
                       rmax, idx, (float*)d2, cnt);
        }
    }
    return B3D_OK;
}

}  // extern "C"
