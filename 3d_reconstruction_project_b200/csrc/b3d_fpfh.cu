// b3d_fpfh.cu -- compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn)) -- test/mini1.py:244-250,
// test/check2.py:95-100 (the descriptor behind the reference's RANSAC initialisation; a "next" row of SURVEY.md 8f).
// Three kernels over the hashed grid: (1) hybrid neighbour lists in (d2, index) order, (2) SPFH histograms (3 x 11 bins of the
// pair features of every neighbour), (3) FPFH = distance-weighted sum of the neighbours' SPFH, each 11-bin group normalised
// to 100, plus the point's own SPFH. Sums run in neighbour order, like the reference's loops.
#include "b3d_common.cuh"
#include "b3d_search.cuh"

namespace b3d {
namespace {

constexpr int kFpfhK = 128;  // neighbour list capacity (the reference uses max_nn = 100)

__global__ void __launch_bounds__(64) fpfh_neighbors_kernel(GridView<double> g, const int32_t* __restrict__ off, int k, bool use_radius, double r2, int rmax,
                                                            int32_t* __restrict__ nb, double* __restrict__ nd, int32_t* __restrict__ cnt) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const double4 q = ld_point(g.pts + pos);
    TopK<double, kFpfhK> tk;
    knn_hybrid_query<double, kFpfhK>(g, off, 0, q.x, q.y, q.z, k, use_radius, r2, rmax, tk);
    const int64_t oi = point_index(q);
    for (int j = 0; j < tk.n; ++j) {
        nb[oi * k + j] = point_index(ld_point(g.pts + tk.pos[j]));
        nd[oi * k + j] = tk.d2[j];
    }
    cnt[oi] = tk.n;
}

__device__ __forceinline__ int bin11(double t) {
    const int h = (int)floor(11.0 * t);
    return h < 0 ? 0 : (h > 10 ? 10 : h);
}

__device__ __forceinline__ void pair_features(const double* p1, const double* n1, const double* p2, const double* n2, double* f) {
    f[0] = f[1] = f[2] = f[3] = 0.0;
    double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    const double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (dist == 0.0) return;
    double a[3] = {n1[0], n1[1], n1[2]}, b[3] = {n2[0], n2[1], n2[2]};
    const double angle1 = (a[0] * d[0] + a[1] * d[1] + a[2] * d[2]) / dist;
    const double angle2 = (b[0] * d[0] + b[1] * d[1] + b[2] * d[2]) / dist;
    double f2;
    if (acos(fabs(angle1)) > acos(fabs(angle2))) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = n2[k]; b[k] = n1[k]; d[k] = -d[k]; }
        f2 = -angle2;
    } else {
        f2 = angle1;
    }
    double v[3] = {d[1] * a[2] - d[2] * a[1], d[2] * a[0] - d[0] * a[2], d[0] * a[1] - d[1] * a[0]};
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (vn == 0.0) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] /= vn;
    const double w[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    f[3] = dist;
    f[2] = f2;
    f[1] = v[0] * b[0] + v[1] * b[1] + v[2] * b[2];
    f[0] = atan2(w[0] * b[0] + w[1] * b[1] + w[2] * b[2], a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
}

__global__ void __launch_bounds__(128) spfh_kernel(const double* __restrict__ xyz, const double* __restrict__ nrm, int64_t n, int k,
                                                   const int32_t* __restrict__ nb, const int32_t* __restrict__ cnt, double* __restrict__ spfh) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double h[33];
#pragma unroll
    for (int j = 0; j < 33; ++j) h[j] = 0.0;
    const int c = cnt[i];
    if (c > 1) {
        const double incr = 100.0 / (double)(c - 1);
        const double pi = 3.14159265358979323846;
        const double p1[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, n1[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
        for (int t = 1; t < c; ++t) {
            const int64_t q = nb[i * k + t];
            const double p2[3] = {xyz[3 * q], xyz[3 * q + 1], xyz[3 * q + 2]}, n2[3] = {nrm[3 * q], nrm[3 * q + 1], nrm[3 * q + 2]};
            double f[4];
            pair_features(p1, n1, p2, n2, f);
            h[bin11((f[0] + pi) / (2.0 * pi))] += incr;
            h[11 + bin11((f[1] + 1.0) * 0.5)] += incr;
            h[22 + bin11((f[2] + 1.0) * 0.5)] += incr;
        }
    }
    for (int j = 0; j < 33; ++j) spfh[i * 33 + j] = h[j];
}

__global__ void __launch_bounds__(128) fpfh_kernel(int64_t n, int k, const int32_t* __restrict__ nb, const double* __restrict__ nd,
                                                   const int32_t* __restrict__ cnt, const double* __restrict__ spfh, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[33];
#pragma unroll
    for (int j = 0; j < 33; ++j) o[j] = 0.0;
    const int c = cnt[i];
    if (c > 1) {
        double sum[3] = {0.0, 0.0, 0.0};
        for (int t = 1; t < c; ++t) {
            const double dist = nd[i * k + t];
            if (dist == 0.0) continue;
            const double* sj = spfh + (int64_t)nb[i * k + t] * 33;
#pragma unroll
            for (int j = 0; j < 33; ++j) {
                const double val = sj[j] / dist;
                sum[j / 11] += val;
                o[j] += val;
            }
        }
#pragma unroll
        for (int gq = 0; gq < 3; ++gq)
            if (sum[gq] != 0.0) sum[gq] = 100.0 / sum[gq];
#pragma unroll
        for (int j = 0; j < 33; ++j) {
            o[j] *= sum[j / 11];
            o[j] += spfh[i * 33 + j];
        }
    }
#pragma unroll
    for (int j = 0; j < 33; ++j) out[i * 33 + j] = o[j];
}

}  // namespace
}  // namespace b3d

using namespace b3d;

extern "C" int b3d_compute_fpfh(b3d_ctx* ctx, const double* xyz, const double* normals, int64_t n, int max_nn, double radius, double* out) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n >= 0, "negative point count");
    B3D_REQUIRE(max_nn >= 1 && max_nn <= kFpfhK, "max_nn must be in [1, %d] (got %d)", kFpfhK, max_nn);
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz && out, "b3d_compute_fpfh: NULL buffer");
    B3D_REQUIRE(normals != nullptr, "Failed because input point cloud has no normal.");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    Grid<double> grid;
    int rmax = kMaxRing;
    B3D_TRY(build_search_grid<double>(ctx, xyz, seg, max_nn, radius, &grid, &rmax));
    DevBuf<int32_t> nb, cnt;
    DevBuf<double> nd, spfh;
    B3D_TRY(nb.alloc(ctx, (size_t)n * max_nn));
    B3D_TRY(nd.alloc(ctx, (size_t)n * max_nn));
    B3D_TRY(cnt.alloc(ctx, (size_t)n));
    B3D_TRY(spfh.alloc(ctx, (size_t)n * 33));
    const bool use_radius = radius > 0;
    B3D_LAUNCH(ctx, fpfh_neighbors_kernel, (int)((n + 63) / 64), 64, 0, grid.view(), seg.off, max_nn, use_radius, radius * radius, rmax, nb.p, nd.p, cnt.p);
    B3D_LAUNCH(ctx, spfh_kernel, (int)((n + 127) / 128), 128, 0, xyz, normals, n, max_nn, nb.p, cnt.p, spfh.p);
    B3D_LAUNCH(ctx, fpfh_kernel, (int)((n + 127) / 128), 128, 0, n, max_nn, nb.p, nd.p, cnt.p, spfh.p, out);
    return B3D_OK;
}
