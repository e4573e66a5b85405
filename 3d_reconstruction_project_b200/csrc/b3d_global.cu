// b3d_global.cu -- global registration from feature matches: registration_ransac_based_on_feature_matching(source, target,
// source_fpfh, target_fpfh, mutual_filter, max_correspondence_distance, TransformationEstimationPointToPoint(False), ransac_n,
// [CorrespondenceCheckerBasedOnEdgeLength, CorrespondenceCheckerBasedOnDistance], RANSACConvergenceCriteria(max_iteration,
// confidence)) -- test/mini1.py:269-281, test/check2.py:132-144, test/check3.py:181 (the initial alignment of the reference's
// multiway registration; a "next" row of SURVEY.md 8f).
//
// Two pieces: (1) nearest neighbour of every source feature among the target features (exact brute force in float64, tiles
// in shared memory); (2) the RANSAC loop. The library draws one hypothesis at a time (per OpenMP thread); here a round draws
// thousands at once -- hypothesis `itr` is a pure function of (seed, itr): counter-based picks, Umeyama on the picked pairs,
// the cheap checkers -- the survivors are validated together (nearest target point of every transformed source point
// through the hashed grid, inlier count and quantised sum of squared distances by integer atomics, so the numbers do not
// depend on the order of the additions), and the host replays the library's sequential bookkeeping over the survivors in
// `itr` order (best result, estimated number of iterations still needed for the requested confidence). The outcome is what
// a single-threaded run of the library's loop would give with the same random picks.
#include "b3d_common.cuh"
#include "b3d_rigid.cuh"
#include "b3d_search.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace b3d {
namespace {

constexpr int kRansacMaxN = 8;
constexpr int kFeatTileQ = 128, kFeatTileT = 32, kFeatMaxDim = 64;

// ---- feature matching ------------------------------------------------------------------------------------------------
// squared L2 in the accumulation order of the library's k-d tree metric: four terms at a time, then the tail
__device__ __forceinline__ double feature_dist2(const double* __restrict__ a, const double* __restrict__ b, int dim) {
    double r = 0.0;
    int k = 0;
    for (; k + 3 < dim; k += 4) {
        const double d0 = a[k] - b[k], d1 = a[k + 1] - b[k + 1], d2 = a[k + 2] - b[k + 2], d3 = a[k + 3] - b[k + 3];
        r += ((d0 * d0 + d1 * d1) + d2 * d2) + d3 * d3;
    }
    for (; k < dim; ++k) {
        const double d = a[k] - b[k];
        r += d * d;
    }
    return r;
}

__global__ void __launch_bounds__(kFeatTileQ) feature_nn_kernel(const double* __restrict__ fa, int64_t na, const double* __restrict__ fb, int64_t nb, int dim,
                                                                int32_t* __restrict__ nn) {
    extern __shared__ double smem[];
    double* sq = smem;                               // [kFeatTileQ][dim + 1]
    double* st = smem + kFeatTileQ * (dim + 1);      // [kFeatTileT][dim]
    const int64_t q0 = (int64_t)blockIdx.x * kFeatTileQ;
    for (int e = threadIdx.x; e < kFeatTileQ * dim; e += kFeatTileQ) {
        const int r = e / dim, c = e % dim;
        sq[r * (dim + 1) + c] = q0 + r < na ? fa[(q0 + r) * dim + c] : 0.0;
    }
    __syncthreads();
    const double* mine = sq + threadIdx.x * (dim + 1);
    double best = 1.0e300;
    int32_t bi = -1;
    for (int64_t t0 = 0; t0 < nb; t0 += kFeatTileT) {
        const int nt = (int)min((int64_t)kFeatTileT, nb - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nt * dim; e += kFeatTileQ) st[e] = fb[t0 * dim + e];
        __syncthreads();
        for (int j = 0; j < nt; ++j) {
            const double d = feature_dist2(mine, st + j * dim, dim);
            if (d < best) {  // ascending scan: the first of equals (smallest index) stays
                best = d;
                bi = (int32_t)(t0 + j);
            }
        }
    }
    if (q0 + threadIdx.x < na) nn[q0 + threadIdx.x] = bi;
}

// ---- RANSAC ----------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// j-th pick of hypothesis itr: uniform in [0, nc) (multiply-high of a 64-bit hash)
__device__ __forceinline__ long long ransac_pick(unsigned long long seed, long long itr, int j, long long nc) {
    const unsigned long long x = splitmix64(splitmix64(seed) ^ ((unsigned long long)itr * (unsigned long long)kRansacMaxN + (unsigned long long)j));
    return (long long)__umul64hi(x, (unsigned long long)nc);
}

struct RansacArgs {
    const double* src;
    const double* tgt;
    const int32_t* corres;  // [nc][2]
    long long nc;
    int n;
    double edge_similarity;  // <= 0: checker off
    double check_distance;   // <= 0: checker off
    unsigned long long seed;
};

__global__ void __launch_bounds__(128) ransac_hypothesis_kernel(RansacArgs A, long long itr0, int count, int* __restrict__ n_pass, long long* __restrict__ pass_itr,
                                                                double* __restrict__ pass_T) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= count) return;
    const long long itr = itr0 + tid;
    double s[kRansacMaxN][3], t[kRansacMaxN][3];
    double mu_s[3] = {0, 0, 0}, mu_d[3] = {0, 0, 0};
    for (int j = 0; j < A.n; ++j) {
        const long long c = ransac_pick(A.seed, itr, j, A.nc);
        const int si = A.corres[2 * c], ti = A.corres[2 * c + 1];
        for (int a = 0; a < 3; ++a) {
            s[j][a] = A.src[3 * (int64_t)si + a];
            t[j][a] = A.tgt[3 * (int64_t)ti + a];
            mu_s[a] += s[j][a];
            mu_d[a] += t[j][a];
        }
    }
    // CorrespondenceCheckerBasedOnEdgeLength: every pair of picks keeps its length within the similarity ratio, both ways
    if (A.edge_similarity > 0.0) {
        for (int i = 0; i < A.n; ++i)
            for (int j = i + 1; j < A.n; ++j) {
                const double ds = sqrt(dist2<double>(s[i][0] - s[j][0], s[i][1] - s[j][1], s[i][2] - s[j][2]));
                const double dt = sqrt(dist2<double>(t[i][0] - t[j][0], t[i][1] - t[j][1], t[i][2] - t[j][2]));
                if (ds < dt * A.edge_similarity || dt < ds * A.edge_similarity) return;
            }
    }
    // TransformationEstimationPointToPoint(with_scaling = false): Eigen::umeyama on the picked pairs
    const double inv_n = 1.0 / (double)A.n;
    for (int a = 0; a < 3; ++a) {
        mu_s[a] *= inv_n;
        mu_d[a] *= inv_n;
    }
    double Sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < A.n; ++j)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Sigma[3 * r + c] += (t[j][r] - mu_d[r]) * (s[j][c] - mu_s[c]);
    for (int k = 0; k < 9; ++k) Sigma[k] *= inv_n;
    double T[16];
    umeyama_from_moments(mu_s, mu_d, Sigma, T);
    for (int k = 0; k < 12; ++k)
        if (!isfinite(T[k])) return;
    // CorrespondenceCheckerBasedOnDistance: every picked pair ends up within the threshold
    if (A.check_distance > 0.0) {
        for (int j = 0; j < A.n; ++j) {
            const double x = T[0] * s[j][0] + T[1] * s[j][1] + T[2] * s[j][2] + T[3];
            const double y = T[4] * s[j][0] + T[5] * s[j][1] + T[6] * s[j][2] + T[7];
            const double z = T[8] * s[j][0] + T[9] * s[j][1] + T[10] * s[j][2] + T[11];
            if (sqrt(dist2<double>(x - t[j][0], y - t[j][1], z - t[j][2])) > A.check_distance) return;
        }
    }
    const int slot = atomicAdd(n_pass, 1);
    pass_itr[slot] = itr;
    for (int k = 0; k < 12; ++k) pass_T[12 * (int64_t)slot + k] = T[k];
}

// Validation of the survivors: blockIdx.y = survivor, threads over the source points in the order of their own grid (spatially
// coherent under any rigid motion). Inlier count and sum of squared distances in units of r2 / 2^40 by integer atomics.
__global__ void __launch_bounds__(128) ransac_validate_kernel(GridView<double> src_sorted, GridView<double> tgt, int rmax, double r2, double q_scale,
                                                              const double* __restrict__ pass_T, unsigned int* __restrict__ cnt, unsigned long long* __restrict__ sumq) {
    const int h = blockIdx.y;
    __shared__ double sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = pass_T[12 * (int64_t)h + threadIdx.x];
    __syncthreads();
    unsigned int c = 0;
    unsigned long long sq = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < src_sorted.n; i += gridDim.x * blockDim.x) {
        const double4 p = ld_point(src_sorted.pts + i);
        const double x = sT[0] * p.x + sT[1] * p.y + sT[2] * p.z + sT[3];
        const double y = sT[4] * p.x + sT[5] * p.y + sT[6] * p.z + sT[7];
        const double z = sT[8] * p.x + sT[9] * p.y + sT[10] * p.z + sT[11];
        double d2 = 0.0;
        int idx = 0;
        const int pos = nn_within_query<double>(tgt, 0, x, y, z, r2, rmax, &d2, &idx);
        if (pos >= 0 && d2 < r2) {
            ++c;
            sq += (unsigned long long)(d2 * q_scale);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((threadIdx.x & 31) == 0 && c > 0) {
        atomicAdd(&cnt[h], c);
        atomicAdd(&sumq[h], sq);
    }
}

}  // namespace
}  // namespace b3d

using namespace b3d;

extern "C" int b3d_match_features(b3d_ctx* ctx, const double* feat_a, int64_t na, const double* feat_b, int64_t nb, int dim, int32_t* nn_out) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(na >= 0 && nb >= 0, "negative feature count");
    B3D_REQUIRE(dim >= 1 && dim <= kFeatMaxDim, "feature dimension must be in [1, %d] (got %d)", kFeatMaxDim, dim);
    if (na == 0) return B3D_OK;
    B3D_REQUIRE(feat_a != nullptr && nn_out != nullptr && (feat_b != nullptr || nb == 0), "b3d_match_features: NULL buffer");
    B3D_TRY(ctx->bind());
    const size_t smem = (size_t)(kFeatTileQ * (dim + 1) + kFeatTileT * dim) * sizeof(double);
    B3D_CUDA(cudaFuncSetAttribute(feature_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B3D_LAUNCH(ctx, feature_nn_kernel, (int)((na + kFeatTileQ - 1) / kFeatTileQ), kFeatTileQ, smem, feat_a, na, feat_b, nb, dim, nn_out);
    return B3D_OK;
}

extern "C" int b3d_ransac_correspondence(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const int32_t* corres, int64_t nc,
                                         double max_dist, int ransac_n, double edge_similarity, double checker_distance, int64_t max_iteration,
                                         double confidence, uint64_t seed, b3d_ransac_result* result_h) {
    B3D_REQUIRE(ctx != nullptr && result_h != nullptr, "b3d_ransac_correspondence: NULL argument");
    B3D_REQUIRE(ns >= 0 && nt >= 0 && nc >= 0, "negative count");
    B3D_REQUIRE(ransac_n <= kRansacMaxN, "ransac_n must be <= %d (got %d)", kRansacMaxN, ransac_n);
    B3D_REQUIRE(confidence >= 0.0 && confidence <= 1.0 && max_iteration >= 0, "RANSACConvergenceCriteria out of range");
    for (int i = 0; i < 16; ++i) result_h->transformation[i] = (i % 5 == 0) ? 1.0 : 0.0;
    result_h->fitness = 0.0;
    result_h->inlier_rmse = 0.0;
    result_h->n_correspondences = 0;
    result_h->iterations = 0;
    result_h->validated = 0;
    // the library returns an empty result for these instead of raising
    if (ransac_n < 3 || max_dist <= 0.0 || nc < ransac_n || ns == 0 || nt == 0) return B3D_OK;
    B3D_REQUIRE(src != nullptr && tgt != nullptr && corres != nullptr, "b3d_ransac_correspondence: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> soff, toff;
    Segments sseg, tseg;
    B3D_TRY(single_segment(ctx, ns, &soff, &sseg));
    B3D_TRY(single_segment(ctx, nt, &toff, &tseg));
    Grid<double> sgrid, tgrid;
    int rmax = 1;
    B3D_TRY(build_search_grid<double>(ctx, tgt, tseg, 8, max_dist, &tgrid, &rmax));
    B3D_TRY(build_search_grid<double>(ctx, src, sseg, 8, max_dist, &sgrid, nullptr));
    const int64_t R = std::max<int64_t>(256, std::min<int64_t>(65536, ((int64_t)1 << 26) / std::max<int64_t>(ns, 1)));
    DevBuf<int> n_pass;
    DevBuf<long long> pass_itr;
    DevBuf<double> pass_T;
    DevBuf<unsigned int> cnt;
    DevBuf<unsigned long long> sumq;
    B3D_TRY(n_pass.alloc(ctx, 1));
    B3D_TRY(pass_itr.alloc(ctx, R));
    B3D_TRY(pass_T.alloc(ctx, (size_t)R * 12));
    B3D_TRY(cnt.alloc(ctx, R));
    B3D_TRY(sumq.alloc(ctx, R));
    std::vector<long long> itr_h(R);
    std::vector<double> T_h((size_t)R * 12);
    std::vector<unsigned int> cnt_h(R);
    std::vector<unsigned long long> sumq_h(R);
    std::vector<int> order(R);
    const double r2 = max_dist * max_dist;
    const double q_scale = 1099511627776.0 / r2;  // 2^40 per r2
    RansacArgs A{src, tgt, corres, (long long)nc, ransac_n, edge_similarity, checker_distance, (unsigned long long)seed};
    int64_t est_k = max_iteration;
    bool have = false;
    unsigned int best_cnt = 0;
    unsigned long long best_sumq = 0;
    int64_t validated = 0, r0 = 0;
    const int vblocks = (int)std::min<int64_t>((ns + 127) / 128, 148 * 8);
    for (; r0 < std::min(est_k, max_iteration); r0 += R) {
        const int count = (int)std::min<int64_t>(R, max_iteration - r0);
        B3D_CUDA(cudaMemsetAsync(n_pass.p, 0, sizeof(int), ctx->stream));
        B3D_LAUNCH(ctx, ransac_hypothesis_kernel, (count + 127) / 128, 128, 0, A, (long long)r0, count, n_pass.p, pass_itr.p, pass_T.p);
        int S = 0;
        B3D_TRY(ctx->download(&S, n_pass.p, sizeof(int)));
        if (S == 0) continue;
        B3D_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)S * sizeof(unsigned int), ctx->stream));
        B3D_CUDA(cudaMemsetAsync(sumq.p, 0, (size_t)S * sizeof(unsigned long long), ctx->stream));
        B3D_LAUNCH(ctx, ransac_validate_kernel, dim3(vblocks, S), 128, 0, sgrid.view(), tgrid.view(), rmax, r2, q_scale, pass_T.p, cnt.p, sumq.p);
        B3D_TRY(ctx->download(itr_h.data(), pass_itr.p, (size_t)S * sizeof(long long)));
        B3D_TRY(ctx->download(T_h.data(), pass_T.p, (size_t)S * 12 * sizeof(double)));
        B3D_TRY(ctx->download(cnt_h.data(), cnt.p, (size_t)S * sizeof(unsigned int)));
        B3D_TRY(ctx->download(sumq_h.data(), sumq.p, (size_t)S * sizeof(unsigned long long)));
        std::iota(order.begin(), order.begin() + S, 0);
        std::sort(order.begin(), order.begin() + S, [&](int a, int b) { return itr_h[a] < itr_h[b]; });
        // the library's bookkeeping, one survivor at a time in iteration order
        for (int o = 0; o < S; ++o) {
            const int h = order[o];
            if (itr_h[h] >= est_k) break;
            ++validated;
            const unsigned int c = cnt_h[h];
            // IsBetterRANSACThan: fitness, then inlier_rmse; rmse = sqrt(sum / count) compared as sum_a * count_b < sum_b * count_a
            bool better;
            if (!have) better = c > 0;
            else if (c != best_cnt) better = c > best_cnt;
            else better = c > 0 && sumq_h[h] < best_sumq;
            if (!better) continue;
            have = true;
            best_cnt = c;
            best_sumq = sumq_h[h];
            for (int k = 0; k < 12; ++k) result_h->transformation[k] = T_h[(size_t)h * 12 + k];
            const double ratio = (double)c / (double)nc;
            const double est = std::log(1.0 - confidence) / std::log(1.0 - std::pow(ratio, ransac_n));
            if (est < (double)est_k) est_k = (int64_t)std::ceil(est);
        }
    }
    result_h->iterations = std::min<int64_t>(r0, std::min(est_k, max_iteration));
    result_h->validated = validated;
    if (have) {
        result_h->n_correspondences = best_cnt;
        result_h->fitness = (double)best_cnt / (double)ns;
        result_h->inlier_rmse = std::sqrt((double)best_sumq / q_scale / (double)best_cnt);
    }
    return B3D_OK;
}
