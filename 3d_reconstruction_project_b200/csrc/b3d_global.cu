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
#include "b3d_scan.cuh"
#include "b3d_search.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
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


// Inliers of the CORRESPONDENCE SET under every surviving hypothesis (EvaluateInlierCorrespondenceRatio of the library: the
// exit condition of the RANSAC loop uses this share, not the fitness over the whole source). blockIdx.y = survivor.
__global__ void __launch_bounds__(128) ransac_corres_inliers_kernel(const double* __restrict__ src, const double* __restrict__ tgt,
                                                                    const int32_t* __restrict__ corres, long long nc, double r2,
                                                                    const double* __restrict__ pass_T, unsigned int* __restrict__ inl) {
    const int h = blockIdx.y;
    __shared__ double sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = pass_T[12 * (int64_t)h + threadIdx.x];
    __syncthreads();
    unsigned int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += (long long)gridDim.x * blockDim.x) {
        const double* p = src + 3 * (int64_t)corres[2 * i];
        const double* g = tgt + 3 * (int64_t)corres[2 * i + 1];
        const double e0 = (sT[0] * p[0] + sT[1] * p[1] + sT[2] * p[2] + sT[3]) - g[0];
        const double e1 = (sT[4] * p[0] + sT[5] * p[1] + sT[6] * p[2] + sT[7]) - g[1];
        const double e2 = (sT[8] * p[0] + sT[9] * p[1] + sT[10] * p[2] + sT[11]) - g[2];
        if ((e0 * e0 + e1 * e1) + e2 * e2 < r2) ++c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c > 0) atomicAdd(&inl[h], c);
}

// ---- Fast Global Registration (registration_fgr_based_on_feature_matching -- test/check6.py:236-240, check7.py:245, check8.py:244) ----
// Zhou, Park, Koltun 2016 as the library runs it: both clouds centred and scaled by the largest centred norm, mutual nearest
// features, tuple test (three random matches must keep their edge lengths within tuple_scale on both clouds), then 64
// Gauss-Newton steps on sum_c l(|p_c - T q_c|) with the scaled Geman-McClure weight (mu / (r^2 + mu))^2, mu divided by
// division_factor every fourth step while it exceeds maximum_correspondence_distance.
constexpr int kFgrBlock = 128;

__global__ void __launch_bounds__(kFgrBlock) centroid_partial_kernel(const double* __restrict__ xyz, int64_t n, double* __restrict__ partial) {
    // fixed assignment of points to threads and a fixed tree: the same n always adds in the same order
    __shared__ double sh[3][kFgrBlock];
    double a[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < 3; ++k) a[k] += xyz[3 * i + k];
    for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] = a[k];
    __syncthreads();
    for (int o = kFgrBlock / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 3) partial[3 * blockIdx.x + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void centroid_final_kernel(const double* __restrict__ partial, int blocks, int64_t n, double* __restrict__ mean) {
    if (threadIdx.x < 3) {
        double a = 0.0;
        for (int b = 0; b < blocks; ++b) a += partial[3 * b + threadIdx.x];
        mean[threadIdx.x] = a / (double)n;
    }
}
__global__ void max_norm_kernel(const double* __restrict__ xyz, int64_t n, const double* __restrict__ mean, unsigned long long* __restrict__ best) {
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = xyz[3 * i] - mean[0], y = xyz[3 * i + 1] - mean[1], z = xyz[3 * i + 2] - mean[2];
        m = fmax(m, sqrt(dist2<double>(x, y, z)));
    }
    atomicMax(best, (unsigned long long)__double_as_longlong(m));  // non-negative doubles order like their bit patterns
}
__global__ void normalize_kernel(const double* __restrict__ xyz, int64_t n, const double* __restrict__ mean, const unsigned long long* __restrict__ scale_bits,
                                 int use_absolute_scale, double* __restrict__ out) {
    const double scale_global = use_absolute_scale ? 1.0 : __longlong_as_double((long long)*scale_bits);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (xyz[3 * i + k] - mean[k]) / scale_global;
}

struct MutualPred {
    const int32_t* ij;
    const int32_t* ji;
    __device__ __forceinline__ bool operator()(int64_t i) const {
        const int j = ij[i];
        return j >= 0 && ji[j] == (int32_t)i;
    }
};
struct MutualEmit {
    const int32_t* ij;
    int32_t* corres;
    __device__ __forceinline__ void operator()(int64_t i, int64_t slot) const {
        corres[2 * slot] = (int32_t)i;
        corres[2 * slot + 1] = ij[i];
    }
};

// trial t picks three matches; accepted when all three edges keep their length within tuple_scale on both clouds
struct TuplePred {
    const double* p;  // normalised source
    const double* q;  // normalised target
    const int32_t* corres;
    long long nc;
    double scale;
    unsigned long long seed;
    long long t0;
    __device__ __forceinline__ void picks(int64_t t, int (&a)[3]) const {
        for (int k = 0; k < 3; ++k) a[k] = (int)ransac_pick(seed ^ 0x5851F42D4C957F2Dull, t0 + t, k, nc);
    }
    __device__ __forceinline__ bool operator()(int64_t t) const {
        int a[3];
        picks(t, a);
        double li[3], lj[3];
        for (int k = 0; k < 3; ++k) {
            const int u = a[k], v = a[(k + 1) % 3];
            const int iu = corres[2 * u], iv = corres[2 * v], ju = corres[2 * u + 1], jv = corres[2 * v + 1];
            li[k] = sqrt(dist2<double>(p[3 * iu] - p[3 * iv], p[3 * iu + 1] - p[3 * iv + 1], p[3 * iu + 2] - p[3 * iv + 2]));
            lj[k] = sqrt(dist2<double>(q[3 * ju] - q[3 * jv], q[3 * ju + 1] - q[3 * jv + 1], q[3 * ju + 2] - q[3 * jv + 2]));
        }
        for (int k = 0; k < 3; ++k)
            if (!(li[k] * scale < lj[k] && lj[k] < li[k] / scale)) return false;
        return true;
    }
};
struct TupleEmit {
    TuplePred P;
    int32_t* out;  // [cap][3][2]
    long long base, cap;
    __device__ __forceinline__ void operator()(int64_t t, int64_t slot) const {
        const long long s = base + slot;
        if (s >= cap) return;
        int a[3];
        P.picks(t, a);
        for (int k = 0; k < 3; ++k) {
            out[6 * s + 2 * k] = P.corres[2 * a[k]];
            out[6 * s + 2 * k + 1] = P.corres[2 * a[k] + 1];
        }
    }
};

struct FgrArgs {
    const double* p;   // normalised source (cloud i)
    double* q;         // normalised target (cloud j), transformed in place step by step
    const int32_t* corres;
    int nc;
    double par, max_corr_dist, division_factor;
    int decrease_mu, iterations, use_absolute_scale;
    const double* mean_s;
    const double* mean_t;
    const unsigned long long* scale_bits;
    int64_t nq;
    double* T_out;  // [16] source -> target, original coordinates
};

// The whole optimisation in ONE block: per step every thread adds its matches' 27 normal-equation sums (fixed assignment),
// a fixed tree adds the threads, thread 0 solves the 6x6 system and updates the transform, all threads move the target copy.
__global__ void __launch_bounds__(kFgrBlock) fgr_optimize_kernel(FgrArgs A) {
    __shared__ double sh[27][kFgrBlock + 1];
    __shared__ double s_delta[16], s_trans[16];
    __shared__ double s_par;
    if (threadIdx.x < 16) s_trans[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0 : 0.0;
    if (threadIdx.x == 0) s_par = A.par;
    __syncthreads();
    for (int itr = 0; itr < A.iterations; ++itr) {
        const double par = s_par;
        double a[27];
        for (int k = 0; k < 27; ++k) a[k] = 0.0;
        for (int c = threadIdx.x; c < A.nc; c += kFgrBlock) {
            const int ii = A.corres[2 * c], jj = A.corres[2 * c + 1];
            const double px = A.p[3 * ii], py = A.p[3 * ii + 1], pz = A.p[3 * ii + 2];
            const double qx = A.q[3 * jj], qy = A.q[3 * jj + 1], qz = A.q[3 * jj + 2];
            const double r[3] = {px - qx, py - qy, pz - qz};
            const double temp = par / (dist2<double>(r[0], r[1], r[2]) + par);
            const double w = temp * temp;
            // rows of d(p - q')/d(omega, v) for q' = q + omega x q + v
            const double J[3][6] = {{0.0, -qz, qy, -1.0, 0.0, 0.0}, {qz, 0.0, -qx, 0.0, -1.0, 0.0}, {-qy, qx, 0.0, 0.0, 0.0, -1.0}};
            for (int row = 0; row < 3; ++row) {
                int t = 0;
                for (int u = 0; u < 6; ++u)
                    for (int v = u; v < 6; ++v) a[t++] += J[row][u] * J[row][v] * w;
                for (int u = 0; u < 6; ++u) a[21 + u] += J[row][u] * r[row] * w;
            }
        }
        for (int k = 0; k < 27; ++k) sh[k][threadIdx.x] = a[k];
        __syncthreads();
        for (int o = kFgrBlock / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o)
                for (int k = 0; k < 27; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            double M[36], b[6], x[6];
            int t = 0;
            for (int u = 0; u < 6; ++u)
                for (int v = u; v < 6; ++v) { M[6 * u + v] = sh[t][0]; M[6 * v + u] = sh[t][0]; ++t; }
            for (int u = 0; u < 6; ++u) b[u] = -sh[21 + u][0];
            double D[16];
            mat4_identity(D);
            if (solve6(M, b, x)) vec6_to_mat4(x, D);
            double Tn[16];
            mat4_mul(D, s_trans, Tn);
            for (int k = 0; k < 16; ++k) { s_trans[k] = Tn[k]; s_delta[k] = D[k]; }
            // graduated non-convexity
            if (A.decrease_mu && itr % 4 == 0 && par > A.max_corr_dist) s_par = par / A.division_factor;
        }
        __syncthreads();
        for (int64_t i = threadIdx.x; i < A.nq; i += kFgrBlock) {
            const double x = A.q[3 * i], y = A.q[3 * i + 1], z = A.q[3 * i + 2];
            A.q[3 * i] = s_delta[0] * x + s_delta[1] * y + s_delta[2] * z + s_delta[3];
            A.q[3 * i + 1] = s_delta[4] * x + s_delta[5] * y + s_delta[6] * z + s_delta[7];
            A.q[3 * i + 2] = s_delta[8] * x + s_delta[9] * y + s_delta[10] * z + s_delta[11];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // back to the original coordinates: q_orig -> p_orig is x -> R x + (-R mean_t + t scale + mean_s); the caller wants
        // source -> target, the inverse
        const double sg = A.use_absolute_scale ? 1.0 : __longlong_as_double((long long)*A.scale_bits);
        const double* R = s_trans;
        double t[3];
        for (int r = 0; r < 3; ++r)
            t[r] = -(R[4 * r] * A.mean_t[0] + R[4 * r + 1] * A.mean_t[1] + R[4 * r + 2] * A.mean_t[2]) + R[4 * r + 3] * sg + A.mean_s[r];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) A.T_out[4 * r + c] = R[4 * c + r];
            A.T_out[4 * r + 3] = -(R[r] * t[0] + R[4 + r] * t[1] + R[8 + r] * t[2]);
        }
        A.T_out[12] = 0.0; A.T_out[13] = 0.0; A.T_out[14] = 0.0; A.T_out[15] = 1.0;
    }
}

int launch_feature_nn(b3d_ctx* ctx, const double* fa, int64_t na, const double* fb, int64_t nb, int dim, int32_t* nn) {
    const size_t smem = (size_t)(kFeatTileQ * (dim + 1) + kFeatTileT * dim) * sizeof(double);
    B3D_CUDA(cudaFuncSetAttribute(feature_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B3D_LAUNCH(ctx, feature_nn_kernel, (int)((na + kFeatTileQ - 1) / kFeatTileQ), kFeatTileQ, smem, fa, na, fb, nb, dim, nn);
    return B3D_OK;
}

int normalise_cloud(b3d_ctx* ctx, const double* xyz, int64_t n, double* mean_d, unsigned long long* scale_bits) {
    const int blocks = (int)std::min<int64_t>(1024, (n + kFgrBlock - 1) / kFgrBlock);
    DevBuf<double> partial;
    B3D_TRY(partial.alloc(ctx, (size_t)blocks * 3));
    B3D_LAUNCH(ctx, centroid_partial_kernel, blocks, kFgrBlock, 0, xyz, n, partial.p);
    B3D_LAUNCH(ctx, centroid_final_kernel, 1, 32, 0, partial.p, blocks, n, mean_d);
    B3D_LAUNCH(ctx, max_norm_kernel, blocks, kFgrBlock, 0, xyz, n, mean_d, scale_bits);
    return B3D_OK;
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
    return launch_feature_nn(ctx, feat_a, na, feat_b, nb, dim, nn_out);
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
    // hypotheses per round: bounded by the validation work (survivors x source points) and by gridDim.y of the validation launch
    const int64_t R = std::max<int64_t>(256, std::min<int64_t>(32768, ((int64_t)1 << 26) / std::max<int64_t>(ns, 1)));
    DevBuf<int> n_pass;
    DevBuf<long long> pass_itr;
    DevBuf<double> pass_T;
    DevBuf<unsigned int> cnt, inl;
    DevBuf<unsigned long long> sumq;
    B3D_TRY(n_pass.alloc(ctx, 1));
    B3D_TRY(inl.alloc(ctx, R));
    B3D_TRY(pass_itr.alloc(ctx, R));
    B3D_TRY(pass_T.alloc(ctx, (size_t)R * 12));
    B3D_TRY(cnt.alloc(ctx, R));
    B3D_TRY(sumq.alloc(ctx, R));
    std::vector<long long> itr_h(R);
    std::vector<double> T_h((size_t)R * 12);
    std::vector<unsigned int> cnt_h(R), inl_h(R);
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
        B3D_CUDA(cudaMemsetAsync(inl.p, 0, (size_t)S * sizeof(unsigned int), ctx->stream));
        B3D_LAUNCH(ctx, ransac_corres_inliers_kernel, dim3((unsigned int)std::min<int64_t>((nc + 127) / 128, 148 * 4), S), 128, 0, src, tgt, corres,
                   (long long)nc, r2, pass_T.p, inl.p);
        B3D_TRY(ctx->download(inl_h.data(), inl.p, (size_t)S * sizeof(unsigned int)));
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
            // exit condition: the inlier share of the correspondence set under this hypothesis (Open3D >= 0.13)
            const double ratio = std::min(1.0, (double)inl_h[h] / (double)nc);
            if (ratio > 0.0) {
                const double est = ratio >= 1.0 ? 0.0 : std::log(1.0 - confidence) / std::log(1.0 - std::pow(ratio, ransac_n));
                if (est < (double)est_k) est_k = (int64_t)std::ceil(est);
            }
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

extern "C" int b3d_fgr_feature_matching(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const double* feat_src,
                                        const double* feat_tgt, int dim, const b3d_fgr_option* opt, uint64_t seed, double* T_h, int64_t* n_corres_h) {
    B3D_REQUIRE(ctx != nullptr && opt != nullptr && T_h != nullptr, "b3d_fgr_feature_matching: NULL argument");
    B3D_REQUIRE(ns >= 0 && nt >= 0, "negative point count");
    B3D_REQUIRE(dim >= 1 && dim <= kFeatMaxDim, "feature dimension must be in [1, %d] (got %d)", kFeatMaxDim, dim);
    B3D_REQUIRE(opt->division_factor > 1.0 && opt->tuple_scale > 0.0 && opt->tuple_scale < 1.0 && opt->iteration_number >= 0 &&
                    opt->maximum_tuple_count >= 1 && opt->maximum_correspondence_distance > 0.0,
                "FastGlobalRegistrationOption out of range");
    for (int i = 0; i < 16; ++i) T_h[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (n_corres_h) *n_corres_h = 0;
    B3D_REQUIRE(ns > 0 && nt > 0, "FastGlobalRegistration: the clouds must not be empty");
    B3D_REQUIRE(src && tgt && feat_src && feat_tgt, "b3d_fgr_feature_matching: NULL buffer");
    B3D_REQUIRE(ns < (int64_t)0x7fffffff && nt < (int64_t)0x7fffffff, "too many points");
    B3D_TRY(ctx->bind());
    // 1. centre both clouds, scale by the largest centred norm of either
    DevBuf<double> mean_s, mean_t, p, q, T_d;
    DevBuf<unsigned long long> scale_bits;
    B3D_TRY(mean_s.alloc(ctx, 3));
    B3D_TRY(mean_t.alloc(ctx, 3));
    B3D_TRY(scale_bits.alloc(ctx, 1));
    B3D_TRY(p.alloc(ctx, (size_t)ns * 3));
    B3D_TRY(q.alloc(ctx, (size_t)nt * 3));
    B3D_TRY(T_d.alloc(ctx, 16));
    B3D_CUDA(cudaMemsetAsync(scale_bits.p, 0, sizeof(unsigned long long), ctx->stream));
    B3D_TRY(normalise_cloud(ctx, src, ns, mean_s.p, scale_bits.p));
    B3D_TRY(normalise_cloud(ctx, tgt, nt, mean_t.p, scale_bits.p));
    const int nb_s = (int)std::min<int64_t>(1024, (ns + kFgrBlock - 1) / kFgrBlock), nb_t = (int)std::min<int64_t>(1024, (nt + kFgrBlock - 1) / kFgrBlock);
    B3D_LAUNCH(ctx, normalize_kernel, nb_s, kFgrBlock, 0, src, ns, mean_s.p, scale_bits.p, opt->use_absolute_scale, p.p);
    B3D_LAUNCH(ctx, normalize_kernel, nb_t, kFgrBlock, 0, tgt, nt, mean_t.p, scale_bits.p, opt->use_absolute_scale, q.p);
    // 2. mutual nearest features
    DevBuf<int32_t> ij, ji, corres, tuples;
    DevBuf<int64_t> count_d;
    B3D_TRY(ij.alloc(ctx, ns));
    B3D_TRY(ji.alloc(ctx, nt));
    B3D_TRY(corres.alloc(ctx, (size_t)ns * 2));
    B3D_TRY(count_d.alloc(ctx, 1));
    B3D_TRY(launch_feature_nn(ctx, feat_src, ns, feat_tgt, nt, dim, ij.p));
    B3D_TRY(launch_feature_nn(ctx, feat_tgt, nt, feat_src, ns, dim, ji.p));
    B3D_TRY(compact(ctx, MutualPred{ij.p, ji.p}, MutualEmit{ij.p, corres.p}, ns, count_d.p));
    int64_t nc = 0;
    B3D_TRY(ctx->download(&nc, count_d.p, sizeof(int64_t)));
    const int32_t* use = corres.p;
    int64_t n_use = nc;
    // 3. tuple test: trials in rounds, the first maximum_tuple_count accepted trials (in trial order) contribute their three matches
    if (opt->tuple_test && nc >= 3) {
        const int64_t cap = opt->maximum_tuple_count, trials = nc * 100;
        B3D_TRY(tuples.alloc(ctx, (size_t)cap * 6));
        int64_t got = 0;
        const int64_t round = 1 << 16;
        for (int64_t t0 = 0; t0 < trials && got < cap; t0 += round) {
            const int64_t n_t = std::min(round, trials - t0);
            TuplePred P{p.p, q.p, corres.p, (long long)nc, opt->tuple_scale, (unsigned long long)seed, (long long)t0};
            B3D_TRY(compact(ctx, P, TupleEmit{P, tuples.p, (long long)got, (long long)cap}, n_t, count_d.p));
            int64_t acc = 0;
            B3D_TRY(ctx->download(&acc, count_d.p, sizeof(int64_t)));
            got = std::min(cap, got + acc);
        }
        use = tuples.p;
        n_use = got * 3;
    }
    if (n_corres_h) *n_corres_h = n_use;
    if (n_use < 10) return B3D_OK;  // the library gives the identity for fewer than 10 matches
    // 4. graduated non-convexity optimisation, one block
    double scale_h = 1.0;
    {
        unsigned long long bits = 0;
        B3D_TRY(ctx->download(&bits, scale_bits.p, sizeof(bits)));
        std::memcpy(&scale_h, &bits, sizeof(double));
    }
    FgrArgs A{p.p, q.p, use, (int)n_use, opt->use_absolute_scale ? scale_h : 1.0, opt->maximum_correspondence_distance, opt->division_factor,
              opt->decrease_mu, opt->iteration_number, opt->use_absolute_scale, mean_s.p, mean_t.p, scale_bits.p, nt, T_d.p};
    B3D_LAUNCH(ctx, fgr_optimize_kernel, 1, kFgrBlock, 0, A);
    B3D_TRY(ctx->download(T_h, T_d.p, 16 * sizeof(double)));
    return B3D_OK;
}
