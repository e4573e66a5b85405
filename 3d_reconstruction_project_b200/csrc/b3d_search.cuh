// b3d_search.cuh -- device-side neighbour search over the hashed cell grid (replaces KDTreeFlann queries inside
// estimate_normals / remove_*_outlier / registration_icp of the reference).
//
// Grid: points sorted by composite cell key ((cloud << shift) | linear cell); every occupied cell owns one slot
// {key, start, end} in an open-addressing hash table (linear probing, load factor <= 0.5). A query walks the cells
// around its own cell ring by ring, prunes cells by their box distance, and stops as soon as no unseen cell can hold
// a better candidate. Result order / tie-break everywhere: (d2, original index) ascending, d2 = ((dx*dx+dy*dy)+dz*dz)
// in the flavour's dtype.
#pragma once

#include "b3d_common.cuh"

namespace b3d {

constexpr unsigned long long kEmptyKey = ~0ull;
// beyond this ring a k-nearest query stops walking shells and scans its whole cloud instead (isolated points)
constexpr int kMaxRing = 6;

// 32-bit mix of a slot key (two multiplies and two xor-shifts; the keys are small packed integers)
__host__ __device__ __forceinline__ uint32_t hash_key(unsigned long long k) {
    uint32_t h = (uint32_t)k * 0x9E3779B1u ^ (uint32_t)(k >> 32) * 0x85EBCA6Bu;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

// announces a global address to L1 (no register, no scoreboard: the data is simply there when the load comes)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ld_point(const float4* p) { return __ldg(p); }
__device__ __forceinline__ double4 ld_point(const double4* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ int point_index(const float4& p) { return __float_as_int(p.w); }
__device__ __forceinline__ int point_index(const double4& p) { return (int)__double_as_longlong(p.w); }
__device__ __forceinline__ float index_as_w(float, int i) { return __int_as_float(i); }
__device__ __forceinline__ double index_as_w(double, int i) { return __longlong_as_double((long long)i); }

template <typename T>
__device__ __forceinline__ bool grid_lookup(const GridView<T>& g, unsigned long long key, int& start, int& end) {
    uint32_t s = hash_key(key) & g.mask;
    while (true) {
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(g.slots + s));
        unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
        if (k == key) {
            start = (int)raw.z;
            end = (int)raw.w;
            return true;
        }
        if (k == kEmptyKey) return false;
        s = (s + 1) & g.mask;
    }
}

// two lookups side by side: both first slot loads are issued before either is looked at (want0 / want1: whether to look at all;
// a missing or unwanted cell returns start == end == 0)
template <typename T>
__device__ __forceinline__ void grid_lookup2(const GridView<T>& g, bool want0, unsigned long long key0, bool want1, unsigned long long key1, int& start0,
                                             int& end0, int& start1, int& end1) {
    uint32_t s0 = hash_key(key0) & g.mask, s1 = hash_key(key1) & g.mask;
    uint4 r0 = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u), r1 = r0;  // kEmptyKey: nothing there
    if (want0) r0 = __ldg(reinterpret_cast<const uint4*>(g.slots + s0));
    if (want1) r1 = __ldg(reinterpret_cast<const uint4*>(g.slots + s1));
    start0 = end0 = start1 = end1 = 0;
    while (true) {
        const unsigned long long k = ((unsigned long long)r0.y << 32) | r0.x;
        if (k == key0) { start0 = (int)r0.z; end0 = (int)r0.w; break; }
        if (k == kEmptyKey) break;
        s0 = (s0 + 1) & g.mask;
        r0 = __ldg(reinterpret_cast<const uint4*>(g.slots + s0));
    }
    while (true) {
        const unsigned long long k = ((unsigned long long)r1.y << 32) | r1.x;
        if (k == key1) { start1 = (int)r1.z; end1 = (int)r1.w; break; }
        if (k == kEmptyKey) break;
        s1 = (s1 + 1) & g.mask;
        r1 = __ldg(reinterpret_cast<const uint4*>(g.slots + s1));
    }
}

template <typename T>
struct SearchSlack;
template <>
struct SearchSlack<double> {
    static constexpr double rel = 1e-9;
};
template <>
struct SearchSlack<float> {
    static constexpr double rel = 1e-5;
};

// Walks the cells around q ring by ring (Chebyshev shells 0..rmax). For every candidate point calls
//   visit(sorted_position, point4)            (the visitor computes the distance itself and keeps its own state)
// thr() returns the current squared-distance bound (double): cells whose box is farther are skipped, and the walk
// stops after a ring when every unseen cell is farther. Returns the last ring completed, or -1 when it ran out of
// rings with rmax > kMaxRing budget (the caller then falls back to a scan of the cloud).
template <typename T, typename Visit, typename Thr>
__device__ __forceinline__ int grid_walk(const GridView<T>& g, int cloud, T qx, T qy, T qz, int rmax, Visit visit, Thr thr) {
    const Lattice L = g.lat[cloud];
    const double h = L.cell;
    double ux = ((double)qx - L.ox) / h, uy = ((double)qy - L.oy) / h, uz = ((double)qz - L.oz) / h;
    // clamp so that the integer conversion is defined for far-away queries
    const double lim = 1.0e9;
    ux = fmin(fmax(ux, -lim), lim);
    uy = fmin(fmax(uy, -lim), lim);
    uz = fmin(fmax(uz, -lim), lim);
    const double fx0 = floor(ux), fy0 = floor(uy), fz0 = floor(uz);
    const long long cx = (long long)fx0 - L.kx0, cy = (long long)fy0 - L.ky0, cz = (long long)fz0 - L.kz0;
    const double fx = ux - fx0, fy = uy - fy0, fz = uz - fz0;  // position inside the home cell, [0,1)
    const double slack = 1.0 - SearchSlack<T>::rel;
    const double h2 = h * h * slack;
    double face = fmin(fmin(fmin(fx, 1.0 - fx), fmin(fy, 1.0 - fy)), fmin(fz, 1.0 - fz));
    // rings entirely outside the lattice cannot hold points: the farthest useful ring
    long long far = 0;
    {
        long long a = cx > (L.nx - 1 - cx) ? cx : (L.nx - 1 - cx);
        long long b = cy > (L.ny - 1 - cy) ? cy : (L.ny - 1 - cy);
        long long c = cz > (L.nz - 1 - cz) ? cz : (L.nz - 1 - cz);
        far = a > b ? a : b;
        far = far > c ? far : c;
    }
    int last = -1;
    for (int R = 0; R <= rmax; ++R) {
        if ((long long)R > far) { last = rmax; break; }
        for (int dx = -R; dx <= R; ++dx) {
            const long long x = cx + dx;
            if (x < 0 || x >= L.nx) continue;
            const double gx = dx > 0 ? (double)dx - fx : (dx < 0 ? fx - (double)dx - 1.0 : 0.0);
            const double gx2 = gx * gx;
            for (int dy = -R; dy <= R; ++dy) {
                const long long y = cy + dy;
                if (y < 0 || y >= L.ny) continue;
                const double gy = dy > 0 ? (double)dy - fy : (dy < 0 ? fy - (double)dy - 1.0 : 0.0);
                const double gxy2 = gx2 + gy * gy;
                const bool shell_xy = (dx == -R || dx == R || dy == -R || dy == R);
                const int zstep = shell_xy ? 1 : (R > 0 ? 2 * R : 1);
                for (int dz = -R; dz <= R; dz += zstep) {
                    const long long z = cz + dz;
                    if (z < 0 || z >= L.nz) continue;
                    const double gz = dz > 0 ? (double)dz - fz : (dz < 0 ? fz - (double)dz - 1.0 : 0.0);
                    const double box2 = (gxy2 + gz * gz) * h2;
                    if (box2 > thr()) continue;
                    const unsigned long long key = grid_slot_key(g.shift, cloud, x, y, z);
                    int s, e;
                    if (!grid_lookup(g, key, s, e)) continue;
                    for (int p = s; p < e; ++p) visit(p, ld_point(g.pts + p));
                }
            }
        }
        last = R;
        const double bound = ((double)R + face);
        if (thr() < bound * bound * h2) break;
    }
    return last;
}

// ---- k nearest (hybrid): sorted list of (d2, sorted position), ties by original index -------------------------
template <typename T, int KMAX>
struct TopK {
    T d2[KMAX];
    int pos[KMAX];
    int n = 0;
    int k = 0;
};

template <typename T, int KMAX>
__device__ __forceinline__ void topk_insert(TopK<T, KMAX>& tk, const GridView<T>& g, T d2, int pos, int idx) {
    int j;
    if (tk.n < tk.k) {
        j = tk.n++;
    } else {
        const int l = tk.k - 1;
        const T dl = tk.d2[l];
        if (d2 < dl || (d2 == dl && idx < point_index(ld_point(g.pts + tk.pos[l]))))
            j = l;
        else
            return;
    }
    while (j > 0) {
        const T dp = tk.d2[j - 1];
        if (d2 < dp || (d2 == dp && idx < point_index(ld_point(g.pts + tk.pos[j - 1])))) {
            tk.d2[j] = dp;
            tk.pos[j] = tk.pos[j - 1];
            --j;
        } else {
            break;
        }
    }
    tk.d2[j] = d2;
    tk.pos[j] = pos;
}

// KDTreeFlann::SearchHybrid(q, radius, k) (radius > 0) / SearchKNN(q, k) (radius <= 0): fills tk with the neighbours
// in (d2, index) order. r2 = radius*radius evaluated in T by the caller.
template <typename T, int KMAX>
__device__ __forceinline__ void knn_hybrid_query(const GridView<T>& g, const int32_t* __restrict__ off, int cloud, T qx, T qy, T qz, int k,
                                                 bool use_radius, T r2, int rmax, TopK<T, KMAX>& tk) {
    tk.n = 0;
    tk.k = k;
    auto visit = [&](int p, const typename PointT<T>::vec4& pt) {
        const T dx = qx - pt.x, dy = qy - pt.y, dz = qz - pt.z;
        const T d2 = dist2<T>(dx, dy, dz);
        if (use_radius && !(d2 < r2)) return;
        topk_insert<T, KMAX>(tk, g, d2, p, point_index(pt));
    };
    auto thr = [&]() -> double {
        if (tk.n == tk.k) return (double)tk.d2[tk.k - 1];
        return use_radius ? (double)r2 : 1.0e300;
    };
    int last = grid_walk<T>(g, cloud, qx, qy, qz, rmax, visit, thr);
    if (!use_radius && last >= rmax && rmax >= kMaxRing) {
        // the shell budget ran out before k neighbours were certain: rescan the whole cloud (rare: isolated points)
        const Lattice L = g.lat[cloud];
        bool need = tk.n < tk.k;
        if (!need) {
            // the list is certain only if its worst entry is closer than the unseen region
            const double h = L.cell;
            double ux = ((double)qx - L.ox) / h, uy = ((double)qy - L.oy) / h, uz = ((double)qz - L.oz) / h;
            double fx = ux - floor(ux), fy = uy - floor(uy), fz = uz - floor(uz);
            double face = fmin(fmin(fmin(fx, 1.0 - fx), fmin(fy, 1.0 - fy)), fmin(fz, 1.0 - fz));
            double bound = ((double)last + face) * h;
            need = !((double)tk.d2[tk.k - 1] < bound * bound * (1.0 - SearchSlack<T>::rel));
        }
        if (need) {
            tk.n = 0;
            const int s = off[cloud], e = off[cloud + 1];
            for (int p = s; p < e; ++p) visit(p, ld_point(g.pts + p));
        }
    }
}

// nearest neighbour with d2 < r2 (ICP correspondence): returns sorted position or -1; ties by original index.
// Home cell first; every further ring only enumerates the cells that the ball of the current best distance overlaps
// (for a converged ICP that is 1-4 cells instead of 26), each still pruned by its own box distance.
template <typename T>
__device__ __forceinline__ int nn_within_query(const GridView<T>& g, int cloud, T qx, T qy, T qz, T r2, int rmax, T* d2_out, int* idx_out) {
    T best = r2;
    int best_pos = -1, best_idx = 0x7fffffff;
    const Lattice L = g.lat[cloud];
    const double h = L.cell;
    double ux = ((double)qx - L.ox) / h, uy = ((double)qy - L.oy) / h, uz = ((double)qz - L.oz) / h;
    const double lim = 1.0e9;
    ux = fmin(fmax(ux, -lim), lim);
    uy = fmin(fmax(uy, -lim), lim);
    uz = fmin(fmax(uz, -lim), lim);
    const double fx0 = floor(ux), fy0 = floor(uy), fz0 = floor(uz);
    const long long cx = (long long)fx0 - L.kx0, cy = (long long)fy0 - L.ky0, cz = (long long)fz0 - L.kz0;
    const double fx = ux - fx0, fy = uy - fy0, fz = uz - fz0;
    const double slack = 1.0 - SearchSlack<T>::rel;
    const double h2 = h * h * slack;
    const double face = fmin(fmin(fmin(fx, 1.0 - fx), fmin(fy, 1.0 - fy)), fmin(fz, 1.0 - fz));
    auto scan_cell = [&](long long x, long long y, long long z) {
        const unsigned long long key = grid_slot_key(g.shift, cloud, x, y, z);
        int s, e;
        if (!grid_lookup(g, key, s, e)) return;
        for (int p = s; p < e; ++p) {
            const typename PointT<T>::vec4 pt = ld_point(g.pts + p);
            const T dx = qx - pt.x, dy = qy - pt.y, dz = qz - pt.z;
            const T d2 = dist2<T>(dx, dy, dz);
            if (best_pos < 0) {
                if (d2 < best) { best = d2; best_pos = p; best_idx = point_index(pt); }
            } else if (d2 <= best) {
                const int idx = point_index(pt);
                if (d2 < best || idx < best_idx) { best = d2; best_pos = p; best_idx = idx; }
            }
        }
    };
    if (cx >= 0 && cx < L.nx && cy >= 0 && cy < L.ny && cz >= 0 && cz < L.nz) scan_cell(cx, cy, cz);
    for (int R = 1; R <= rmax; ++R) {
        {
            const double bound = (double)(R - 1) + face;
            if ((double)best < bound * bound * h2) break;  // no unseen cell can hold a better (or tying) point
        }
        // cells overlapped by the ball of radius sqrt(best), in cell units, widened by the rounding slack
        const double reach = sqrt((double)best) / h * (1.0 + 4.0 * SearchSlack<T>::rel) + 1e-12;
        const int lox = max(-R, (int)floor(fx - reach)), hix = min(R, (int)floor(fx + reach));
        const int loy = max(-R, (int)floor(fy - reach)), hiy = min(R, (int)floor(fy + reach));
        const int loz = max(-R, (int)floor(fz - reach)), hiz = min(R, (int)floor(fz + reach));
        for (int dx = lox; dx <= hix; ++dx) {
            const long long x = cx + dx;
            if (x < 0 || x >= L.nx) continue;
            const double gx = dx > 0 ? (double)dx - fx : (dx < 0 ? fx - (double)dx - 1.0 : 0.0);
            for (int dy = loy; dy <= hiy; ++dy) {
                const long long y = cy + dy;
                if (y < 0 || y >= L.ny) continue;
                const double gy = dy > 0 ? (double)dy - fy : (dy < 0 ? fy - (double)dy - 1.0 : 0.0);
                const double gxy2 = gx * gx + gy * gy;
                const bool shell_xy = (dx == -R || dx == R || dy == -R || dy == R);
                for (int dz = loz; dz <= hiz; ++dz) {
                    if (!shell_xy && dz != -R && dz != R) continue;
                    const long long z = cz + dz;
                    if (z < 0 || z >= L.nz) continue;
                    const double gz = dz > 0 ? (double)dz - fz : (dz < 0 ? fz - (double)dz - 1.0 : 0.0);
                    if ((gxy2 + gz * gz) * h2 > (double)best) continue;
                    scan_cell(x, y, z);
                }
            }
        }
    }
    *d2_out = best;
    *idx_out = best_idx;
    return best_pos;
}

// number of points with d2 < r2 (remove_radius_outlier)
template <typename T>
__device__ __forceinline__ int count_within_query(const GridView<T>& g, int cloud, T qx, T qy, T qz, T r2, int rmax) {
    int c = 0;
    auto visit = [&](int, const typename PointT<T>::vec4& pt) {
        const T dx = qx - pt.x, dy = qy - pt.y, dz = qz - pt.z;
        if (dist2<T>(dx, dy, dz) < r2) ++c;
    };
    auto thr = [&]() -> double { return (double)r2; };
    grid_walk<T>(g, cloud, qx, qy, qz, rmax, visit, thr);
    return c;
}

// rings needed so that every point within `radius` of a query is inside the walked cube (with a rounding margin)
inline int rings_for_radius(double radius, double cell) {
    double r = radius / cell * (1.0 + 1e-6);
    int R = (int)ceil(r);
    return R < 1 ? 1 : R;
}

}  // namespace b3d
