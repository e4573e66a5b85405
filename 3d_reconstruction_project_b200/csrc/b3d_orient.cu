// b3d_orient.cu -- PointCloud.orient_normals_consistent_tangent_plane(k) -- normal_estimation.py:21 (k = 100), the one line of
// NormalEstimation.estimate_normals the normals kernels do not cover.
//
// The library routine (Hoppe et al. 1992 as implemented by Open3D): Riemannian graph = Euclidean minimum spanning tree +
// k-nearest-neighbour edges, edge weight 1 - |n_i . n_j|; minimum spanning tree of that graph; walk the tree from the point
// of largest z (its normal turned towards +z), flipping a child whenever it disagrees with its parent.
//
// Here both spanning trees are Boruvka rounds over flat edge lists (every round: per-component minimum crossing edge under
// the strict order (weight, min end, max end) by two atomicMin sweeps, hook, walk to the new root, relabel), and the tree
// walk is replaced by parity tracking: every vertex carries its sign relative to its component's representative, a hook
// over edge (a, b) fixes the relation of the two representatives to sgn[a] sgn[b] sign(n_a . n_b). The product of the
// flips along a tree path does not depend on the order of the walk, so the result equals the queue traversal's.
// The Euclidean tree is searched among the k nearest neighbours first; what the k-NN graph leaves disconnected is joined by
// exact nearest-foreign-point queries issued by every component but the largest: brute force over tiles of the Morton-sorted
// cloud, pruned by tile bounding boxes against a sampled upper bound of the component's distance to the rest.
#include "b3d_common.cuh"
#include "b3d_search.cuh"

#include <algorithm>
#include <vector>

namespace b3d {
namespace {

constexpr int kOrientK = 128;  // neighbour list capacity (the reference uses k = 100)
constexpr unsigned long long kNoEdge = ~0ull;

// The k nearest neighbours as a SET (the graph stages do not care about their order): a binary max-heap on (d2, index)
// replaces the sorted insertion list of the normals kernels -- at k = 100 an insertion moves 50 entries on average, a heap
// replacement seven.
struct NeighborHeap {
    double d2[kOrientK];
    int idx[kOrientK];
    int n = 0, k = 0;
    __device__ __forceinline__ static bool above(double da, int ia, double db, int ib) { return da > db || (da == db && ia > ib); }
    __device__ __forceinline__ void push(double d, int i) {
        if (n < k) {
            int c = n++;
            while (c > 0) {
                const int p = (c - 1) >> 1;
                if (!above(d, i, d2[p], idx[p])) break;
                d2[c] = d2[p];
                idx[c] = idx[p];
                c = p;
            }
            d2[c] = d;
            idx[c] = i;
        } else if (above(d2[0], idx[0], d, i)) {
            int c = 0;
            while (true) {
                int l = 2 * c + 1;
                if (l >= n) break;
                if (l + 1 < n && above(d2[l + 1], idx[l + 1], d2[l], idx[l])) ++l;
                if (!above(d2[l], idx[l], d, i)) break;
                d2[c] = d2[l];
                idx[c] = idx[l];
                c = l;
            }
            d2[c] = d;
            idx[c] = i;
        }
    }
};

__global__ void __launch_bounds__(64) orient_neighbors_kernel(GridView<double> g, const int32_t* __restrict__ off, int k, int rmax,
                                                              int32_t* __restrict__ nb, double* __restrict__ nd) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const double4 q = ld_point(g.pts + pos);
    NeighborHeap h;
    h.k = k;
    auto visit = [&](int, const double4& pt) { h.push(dist2<double>(q.x - pt.x, q.y - pt.y, q.z - pt.z), point_index(pt)); };
    auto thr = [&]() -> double { return h.n == h.k ? h.d2[0] : 1.0e300; };
    const int last = grid_walk<double>(g, 0, q.x, q.y, q.z, rmax, visit, thr);
    if (last >= rmax && rmax >= kMaxRing) {
        // the ring budget ran out: the set is certain only if its worst member is closer than the space the walk has not seen
        const Lattice L = g.lat[0];
        const double hc = L.cell;
        const double ux = (q.x - L.ox) / hc, uy = (q.y - L.oy) / hc, uz = (q.z - L.oz) / hc;
        const double fx = ux - floor(ux), fy = uy - floor(uy), fz = uz - floor(uz);
        const double face = fmin(fmin(fmin(fx, 1.0 - fx), fmin(fy, 1.0 - fy)), fmin(fz, 1.0 - fz));
        const double bound = ((double)last + face) * hc;
        if (h.n < h.k || !(h.d2[0] < bound * bound * (1.0 - 1e-9))) {
            h.n = 0;
            const int s = off[0], e = off[1];
            for (int p = s; p < e; ++p) visit(p, ld_point(g.pts + p));
        }
    }
    const int64_t oi = point_index(q);
    for (int j = 0; j < k; ++j) {
        int u = -1;
        double d = 0.0;
        if (j < h.n) {
            u = h.idx[j];
            d = h.d2[j];
            if (u == (int)oi) u = -1;  // the point itself is not an edge
        }
        nb[oi * k + j] = u;
        nd[oi * k + j] = d;
    }
}

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ double normal_weight(const double* __restrict__ nrm, int a, int b) {
    const double d = (nrm[3 * (int64_t)a] * nrm[3 * (int64_t)b] + nrm[3 * (int64_t)a + 1] * nrm[3 * (int64_t)b + 1]) +
                     nrm[3 * (int64_t)a + 2] * nrm[3 * (int64_t)b + 2];
    return 1.0 - fabs(d);
}

__global__ void normal_weights_kernel(const int32_t* __restrict__ nb, const double* __restrict__ nrm, int64_t n, int k, double* __restrict__ w) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * k) return;
    const int u = nb[e];
    w[e] = u >= 0 ? normal_weight(nrm, (int)(e / k), u) : 0.0;
}

__global__ void edge_weights_kernel(const int32_t* __restrict__ ea, const int32_t* __restrict__ eb, int64_t m, const double* __restrict__ nrm,
                                    double* __restrict__ w) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    w[e] = normal_weight(nrm, ea[e], eb[e]);
}

__device__ __forceinline__ unsigned long long weight_bits(double w) {
    // weights are >= 0 (squared distances, 1 - |cos|); clamp the -0.0 / tiny negative rounding cases so the bit pattern orders
    return (unsigned long long)__double_as_longlong(w > 0.0 ? w : 0.0);
}
__device__ __forceinline__ unsigned long long edge_id(int a, int b) {
    const unsigned int lo = (unsigned int)min(a, b), hi = (unsigned int)max(a, b);
    return ((unsigned long long)lo << 32) | hi;
}

// Sweep 1 / 2 over a list with `deg` entries per row (row r = vertex row_vertex[r] or r itself): per component, the
// smallest crossing weight (sweep 1), then the smallest edge id among the entries of that weight (sweep 2).
// Components only ever merge, so an entry whose two ends share a component is dead for good: sweep 1 overwrites it with -1
// (no component lookup for it in later rounds) and marks rows without a live entry, which later sweeps skip altogether.
template <int SWEEP>
__global__ void boruvka_sweep_kernel(const int32_t* __restrict__ row_vertex, int32_t* __restrict__ nbr, const double* __restrict__ w, int64_t rows,
                                     int deg, const int32_t* __restrict__ comp, unsigned long long* __restrict__ best_w,
                                     unsigned long long* __restrict__ best_e, uint8_t* __restrict__ row_alive) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (row_alive != nullptr && !row_alive[r]) return;
    const int v = row_vertex ? row_vertex[r] : (int)r;
    const int cv = comp[v];
    unsigned long long my_w = kNoEdge, my_e = kNoEdge;
    const unsigned long long cur_v = SWEEP == 2 ? best_w[cv] : 0ull;
    bool any = false;
    for (int j = 0; j < deg; ++j) {
        const int u = nbr[r * deg + j];
        if (u < 0) continue;
        const int cu = comp[u];
        if (cu == cv) {
            if (SWEEP == 1) nbr[r * deg + j] = -1;
            continue;
        }
        any = true;
        const unsigned long long wb = weight_bits(w[r * deg + j]);
        if (SWEEP == 1) {
            my_w = wb < my_w ? wb : my_w;
            if (wb < best_w[cu]) atomicMin(&best_w[cu], wb);
        } else {
            const unsigned long long id = edge_id(v, u);
            if (wb == cur_v && id < my_e) my_e = id;
            if (wb == best_w[cu] && id < best_e[cu]) atomicMin(&best_e[cu], id);
        }
    }
    if (SWEEP == 1) {
        if (row_alive != nullptr && !any) row_alive[r] = 0;
        if (my_w != kNoEdge && my_w < best_w[cv]) atomicMin(&best_w[cv], my_w);
    } else {
        if (my_e != kNoEdge && my_e < best_e[cv]) atomicMin(&best_e[cv], my_e);
    }
}

// Hook: representative c points at the component on the other side of its chosen edge.
__global__ void boruvka_hook_kernel(int64_t n, const int32_t* __restrict__ comp, const unsigned long long* __restrict__ best_e, const double* __restrict__ nrm,
                                    const signed char* __restrict__ sgn, int32_t* __restrict__ parent, signed char* __restrict__ psign) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    parent[c] = (int32_t)c;
    psign[c] = 1;
    if (comp[c] != (int32_t)c) return;
    const unsigned long long e = best_e[c];
    if (e == kNoEdge) return;
    const int a = (int)(e >> 32), b = (int)(e & 0xffffffffu);
    const int ca = comp[a], cb = comp[b];
    parent[c] = ca == (int32_t)c ? cb : ca;
    if (sgn != nullptr) {
        const double d = (nrm[3 * (int64_t)a] * nrm[3 * (int64_t)b] + nrm[3 * (int64_t)a + 1] * nrm[3 * (int64_t)b + 1]) +
                         nrm[3 * (int64_t)a + 2] * nrm[3 * (int64_t)b + 2];
        const signed char r = d < 0.0 ? -1 : 1;  // TestAndOrientNormal: flip only on a strictly negative dot product
        psign[c] = (signed char)(r * sgn[a] * sgn[b]);
    }
}

// Breaks the two-cycles (both components chose the same edge: the smaller representative stays a root), records every tree
// edge once and counts them.
__global__ void boruvka_cut_kernel(int64_t n, const int32_t* __restrict__ comp, const int32_t* __restrict__ parent, const unsigned long long* __restrict__ best_e,
                                   int32_t* __restrict__ parent2, int32_t* __restrict__ tree_a, int32_t* __restrict__ tree_b, int* __restrict__ counters) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    int32_t p = parent[c];
    if (comp[c] == (int32_t)c && p != (int32_t)c) {
        const bool mutual = parent[p] == (int32_t)c;
        if (mutual && (int32_t)c < p) {
            p = (int32_t)c;  // root of the merged component; the partner records the edge
        } else {
            const int slot = atomicAdd(&counters[0], 1);
            atomicAdd(&counters[1], 1);
            if (tree_a != nullptr) {
                const unsigned long long e = best_e[c];
                tree_a[slot] = (int32_t)(e >> 32);
                tree_b[slot] = (int32_t)(e & 0xffffffffu);
            }
        }
    }
    parent2[c] = p;
}

__global__ void boruvka_root_kernel(int64_t n, const int32_t* __restrict__ comp, const int32_t* __restrict__ parent2, const signed char* __restrict__ psign,
                                    int32_t* __restrict__ root, signed char* __restrict__ rsign) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (comp[c] != (int32_t)c) return;
    int32_t x = (int32_t)c;
    signed char s = 1;
    while (true) {
        const int32_t p = parent2[x];
        if (p == x) break;
        s = (signed char)(s * psign[x]);
        x = p;
    }
    root[c] = x;
    rsign[c] = s;
}

__global__ void boruvka_relabel_kernel(int64_t n, int32_t* __restrict__ comp, signed char* __restrict__ sgn, const int32_t* __restrict__ root,
                                       const signed char* __restrict__ rsign, unsigned long long* __restrict__ best_w, unsigned long long* __restrict__ best_e) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int32_t c = comp[v];
    comp[v] = root[c];
    if (sgn != nullptr) sgn[v] = (signed char)(sgn[v] * rsign[c]);
    best_w[v] = kNoEdge;
    best_e[v] = kNoEdge;
}

__global__ void init_forest_kernel(int64_t n, int32_t* __restrict__ comp, signed char* __restrict__ sgn, unsigned long long* __restrict__ best_w,
                                   unsigned long long* __restrict__ best_e) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    comp[v] = (int32_t)v;
    if (sgn != nullptr) sgn[v] = 1;
    best_w[v] = kNoEdge;
    best_e[v] = kNoEdge;
}

// ---- joining what the k-NN graph leaves apart ----------------------------------------------------------------------
__global__ void comp_size_kernel(int64_t n, const int32_t* __restrict__ comp, int32_t* __restrict__ size) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) atomicAdd(&size[comp[v]], 1);
}
__global__ void comp_largest_kernel(int64_t n, const int32_t* __restrict__ comp, const int32_t* __restrict__ size, unsigned long long* __restrict__ largest) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || comp[c] != (int32_t)c) return;
    // largest size, smallest representative among equals
    atomicMax(largest, ((unsigned long long)(unsigned int)size[c] << 32) | (unsigned int)(0x7fffffff - (int)c));
}

// Upper bound of every small component's distance to the rest: each of its points against a strided sample of the cloud.
// The bound is attained by a real pair, so the component's true minimum crossing edge is not longer.
__global__ void __launch_bounds__(128) foreign_bound_kernel(GridView<double> g, int stride, const int32_t* __restrict__ comp,
                                                            const unsigned long long* __restrict__ largest, unsigned long long* __restrict__ ub) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= g.n) return;
    const double4 q = ld_point(g.pts + pos);
    const int cv = comp[point_index(q)];
    if (cv == 0x7fffffff - (int)(*largest & 0xffffffffu)) return;
    double bd = 1.0e300;
    for (int p = 0; p < g.n; p += stride) {
        const double4 pt = ld_point(g.pts + p);
        const double d2 = dist2<double>(q.x - pt.x, q.y - pt.y, q.z - pt.z);
        if (d2 < bd && comp[point_index(pt)] != cv) bd = d2;
    }
    if (bd < 1.0e300) atomicMin(&ub[cv], (unsigned long long)__double_as_longlong(bd));
}

// Tiles of kTile consecutive points of the sorted (Morton) order are spatially compact: their bounding box and, when all
// of them belong to one component, that component.
constexpr int kTile = 128;
__global__ void __launch_bounds__(kTile) tile_info_kernel(GridView<double> g, const int32_t* __restrict__ comp, double* __restrict__ tile_box,
                                                          int32_t* __restrict__ tile_comp) {
    const int t = blockIdx.x;
    const int p = t * kTile + threadIdx.x;
    const bool valid = p < g.n;
    double lo[3] = {1.0e300, 1.0e300, 1.0e300}, hi[3] = {-1.0e300, -1.0e300, -1.0e300};
    int c = -1;
    if (valid) {
        const double4 pt = ld_point(g.pts + p);
        lo[0] = hi[0] = pt.x; lo[1] = hi[1] = pt.y; lo[2] = hi[2] = pt.z;
        c = comp[point_index(pt)];
    }
    __shared__ double s_lo[3][kTile / 32], s_hi[3][kTile / 32];
    __shared__ int s_c[kTile / 32], s_mixed[kTile / 32];
    const int c_first = __shfl_sync(0xffffffffu, c, 0);
    const bool mixed = __any_sync(0xffffffffu, valid && c != c_first);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min_d(lo[a]);
        hi[a] = warp_max_d(hi[a]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        for (int a = 0; a < 3; ++a) { s_lo[a][warp] = lo[a]; s_hi[a][warp] = hi[a]; }
        s_c[warp] = c_first;
        s_mixed[warp] = mixed ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cc = s_c[0];
        bool mx = s_mixed[0] != 0;
        for (int w = 1; w < kTile / 32; ++w) {
            if (s_c[w] < 0) continue;  // a warp past the end of the cloud
            if (s_mixed[w] || s_c[w] != cc) mx = true;
        }
        for (int a = 0; a < 3; ++a) {
            double l = s_lo[a][0], h = s_hi[a][0];
            for (int w = 1; w < kTile / 32; ++w) { l = fmin(l, s_lo[a][w]); h = fmax(h, s_hi[a][w]); }
            tile_box[6 * t + a] = l;
            tile_box[6 * t + 3 + a] = h;
        }
        tile_comp[t] = mx ? -1 : cc;
    }
}

// Nearest point of another component for every point outside the largest component, (d2, edge id) order: brute force over
// the tiles, skipping every tile that lies in the point's own component or farther than the best so far (which starts at
// the component's sampled bound: points that cannot carry the component's minimum edge report nothing).
__global__ void __launch_bounds__(kTile) foreign_nearest_kernel(GridView<double> g, const int32_t* __restrict__ comp, const unsigned long long* __restrict__ largest,
                                                                const unsigned long long* __restrict__ ub, const double* __restrict__ tile_box,
                                                                const int32_t* __restrict__ tile_comp, int n_tiles, int32_t* __restrict__ fnb,
                                                                double* __restrict__ fd) {
    __shared__ double4 s_pt[kTile];
    __shared__ int s_comp[kTile];
    const int pos = blockIdx.x * kTile + threadIdx.x;
    const int big = 0x7fffffff - (int)(*largest & 0xffffffffu);
    double4 q = make_double4(0.0, 0.0, 0.0, 0.0);
    int v = -1, cv = big;
    if (pos < g.n) {
        q = ld_point(g.pts + pos);
        v = point_index(q);
        cv = comp[v];
        fnb[v] = -1;
        fd[v] = 0.0;
    }
    const bool need = cv != big;
    if (!__syncthreads_or(need ? 1 : 0)) return;
    double bd = 1.0e300;
    if (need) {
        const unsigned long long ubits = ub[cv];
        if (ubits != kNoEdge) bd = __longlong_as_double((long long)ubits);
    }
    int bu = -1;
    for (int t = 0; t < n_tiles; ++t) {
        bool skip = !need || tile_comp[t] == cv;
        if (!skip) {
            const double* bx = tile_box + 6 * t;
            const double dx = fmax(fmax(bx[0] - q.x, q.x - bx[3]), 0.0), dy = fmax(fmax(bx[1] - q.y, q.y - bx[4]), 0.0), dz = fmax(fmax(bx[2] - q.z, q.z - bx[5]), 0.0);
            skip = dist2<double>(dx, dy, dz) * (1.0 - 1e-12) > bd;
        }
        if (__syncthreads_and(skip ? 1 : 0)) continue;
        const int p = t * kTile + threadIdx.x;
        if (p < g.n) {
            const double4 pt = ld_point(g.pts + p);
            s_pt[threadIdx.x] = pt;
            s_comp[threadIdx.x] = comp[point_index(pt)];
        } else {
            s_comp[threadIdx.x] = -2;
        }
        __syncthreads();
        if (!skip) {
            for (int j = 0; j < kTile; ++j) {
                const int cu = s_comp[j];
                if (cu == cv || cu == -2) continue;
                const double4 pt = s_pt[j];
                const double d2 = dist2<double>(q.x - pt.x, q.y - pt.y, q.z - pt.z);
                if (d2 > bd) continue;
                const int u = point_index(pt);
                if (d2 < bd || edge_id(v, u) < edge_id(v, bu)) {
                    bd = d2;
                    bu = u;
                }
            }
        }
        __syncthreads();
    }
    if (need && bu >= 0) {
        fnb[v] = bu;
        fd[v] = bd;
    }
}

// first index of the largest z (the library's loop keeps the first maximum): reduce on order-preserving bits of z, then
// take the smallest index that attains it
__device__ __forceinline__ unsigned long long ordered_bits(double z) {
    const long long b = __double_as_longlong(z + 0.0);  // -0.0 -> +0.0
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__global__ void top_z_kernel(const double* __restrict__ xyz, int64_t n, unsigned long long* __restrict__ best) {
    unsigned long long m = 0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long ob = ordered_bits(xyz[3 * i + 2]);
        m = ob > m ? ob : m;
    }
    if (m != 0ull) atomicMax(best, m);
}
__global__ void top_z_index_kernel(const double* __restrict__ xyz, int64_t n, const unsigned long long* __restrict__ best, int* __restrict__ index) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (ordered_bits(xyz[3 * i + 2]) == *best) atomicMin(index, (int)i);
}

__global__ void apply_flips_kernel(int64_t n, const int32_t* __restrict__ comp, const signed char* __restrict__ sgn, const int* __restrict__ top,
                                   const double* __restrict__ nrm, uint8_t* __restrict__ flipped) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int v0 = *top;
    // the start point looks towards +z; everything connected to it follows through the tree
    signed char o = 1;
    if (comp[v] == comp[v0]) {
        const signed char o0 = nrm[3 * (int64_t)v0 + 2] < 0.0 ? -1 : 1;  // (0, 0, 1) . n < 0
        o = (signed char)(sgn[v] * sgn[v0] * o0);
    } else {
        o = sgn[v];
    }
    flipped[v] = o < 0 ? 1 : 0;
}
__global__ void negate_kernel(int64_t n, const uint8_t* __restrict__ flipped, double* __restrict__ nrm) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n || !flipped[v]) return;
    nrm[3 * v] *= -1.0;
    nrm[3 * v + 1] *= -1.0;
    nrm[3 * v + 2] *= -1.0;
}

struct Forest {
    int64_t n = 0;
    DevBuf<int32_t> comp, parent, parent2, root;
    DevBuf<signed char> sgn, psign, rsign;
    DevBuf<unsigned long long> best_w, best_e;
    DevBuf<int> counters;  // [0] tree edges so far, [1] hooks of the current round
    int edges = 0;
};

int forest_init(b3d_ctx* ctx, Forest* f, int64_t n, bool signs) {
    f->n = n;
    f->edges = 0;
    B3D_TRY(f->comp.alloc(ctx, n));
    B3D_TRY(f->parent.alloc(ctx, n));
    B3D_TRY(f->parent2.alloc(ctx, n));
    B3D_TRY(f->root.alloc(ctx, n));
    B3D_TRY(f->psign.alloc(ctx, n));
    B3D_TRY(f->rsign.alloc(ctx, n));
    if (signs) B3D_TRY(f->sgn.alloc(ctx, n));
    B3D_TRY(f->best_w.alloc(ctx, n));
    B3D_TRY(f->best_e.alloc(ctx, n));
    B3D_TRY(f->counters.alloc(ctx, 2));
    B3D_CUDA(cudaMemsetAsync(f->counters.p, 0, 2 * sizeof(int), ctx->stream));
    const int blocks = (int)((n + 255) / 256);
    B3D_LAUNCH(ctx, init_forest_kernel, blocks, 256, 0, n, f->comp.p, signs ? f->sgn.p : (signed char*)nullptr, f->best_w.p, f->best_e.p);
    return B3D_OK;
}

struct EdgeList {  // rows x deg neighbour entries with weights; row r belongs to vertex r (row_vertex == NULL) or row_vertex[r]
    const int32_t* row_vertex;
    int32_t* nbr;  // entries inside one component are overwritten with -1 as the rounds go
    const double* w;
    int64_t rows;
    int deg;
    uint8_t* row_alive;  // optional [rows], 1 = the row still has a live entry
};

// One Boruvka round over the given lists. hooks_out: number of components that merged into another one.
int boruvka_round(b3d_ctx* ctx, Forest* f, const EdgeList* lists, int n_lists, const double* nrm, int32_t* tree_a, int32_t* tree_b, int* hooks_out) {
    const int64_t n = f->n;
    const int vb = (int)((n + 255) / 256);
    for (int l = 0; l < n_lists; ++l) {
        const EdgeList& L = lists[l];
        if (L.rows == 0) continue;
        B3D_LAUNCH(ctx, boruvka_sweep_kernel<1>, (int)((L.rows + 127) / 128), 128, 0, L.row_vertex, L.nbr, L.w, L.rows, L.deg, f->comp.p, f->best_w.p, f->best_e.p, L.row_alive);
    }
    for (int l = 0; l < n_lists; ++l) {
        const EdgeList& L = lists[l];
        if (L.rows == 0) continue;
        B3D_LAUNCH(ctx, boruvka_sweep_kernel<2>, (int)((L.rows + 127) / 128), 128, 0, L.row_vertex, L.nbr, L.w, L.rows, L.deg, f->comp.p, f->best_w.p, f->best_e.p, L.row_alive);
    }
    B3D_LAUNCH(ctx, boruvka_hook_kernel, vb, 256, 0, n, f->comp.p, f->best_e.p, nrm, f->sgn.p, f->parent.p, f->psign.p);
    B3D_CUDA(cudaMemsetAsync(f->counters.p + 1, 0, sizeof(int), ctx->stream));
    B3D_LAUNCH(ctx, boruvka_cut_kernel, vb, 256, 0, n, f->comp.p, f->parent.p, f->best_e.p, f->parent2.p, tree_a, tree_b, f->counters.p);
    B3D_LAUNCH(ctx, boruvka_root_kernel, vb, 256, 0, n, f->comp.p, f->parent2.p, f->psign.p, f->root.p, f->rsign.p);
    B3D_LAUNCH(ctx, boruvka_relabel_kernel, vb, 256, 0, n, f->comp.p, f->sgn.p, f->root.p, f->rsign.p, f->best_w.p, f->best_e.p);
    int c[2];
    B3D_TRY(ctx->download(c, f->counters.p, sizeof(c)));
    f->edges = c[0];
    *hooks_out = c[1];
    return B3D_OK;
}

}  // namespace
}  // namespace b3d

using namespace b3d;

extern "C" int b3d_orient_normals_consistent_tangent_plane(b3d_ctx* ctx, const double* xyz, double* normals, int64_t n, int k, uint8_t* flipped) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n >= 0, "negative point count");
    B3D_REQUIRE(normals != nullptr || n == 0, "No normals in the PointCloud. Call EstimateNormals() first.");
    B3D_REQUIRE(n >= 4, "Not enough points to create a tetrahedral mesh.");  // the library builds a Delaunay mesh first
    B3D_REQUIRE(xyz != nullptr, "b3d_orient_normals_consistent_tangent_plane: NULL buffer");
    B3D_REQUIRE(k >= 1 && k <= kOrientK, "k must be in [1, %d] (got %d)", kOrientK, k);
    B3D_REQUIRE(n < (int64_t)0x7fffffff, "too many points");
    B3D_TRY(ctx->bind());
    DevBuf<int32_t> off;
    Segments seg;
    B3D_TRY(single_segment(ctx, n, &off, &seg));
    Grid<double> grid;
    int rmax = kMaxRing;
    const int kk = (int)std::min<int64_t>(k, n);
    B3D_TRY(build_search_grid<double>(ctx, xyz, seg, kk, 0.0, &grid, &rmax));
    // 1. k nearest neighbours (the point itself dropped), squared distances
    DevBuf<int32_t> nb;
    DevBuf<double> nw;
    B3D_TRY(nb.alloc(ctx, (size_t)n * kk));
    B3D_TRY(nw.alloc(ctx, (size_t)n * kk));
    B3D_LAUNCH(ctx, orient_neighbors_kernel, (int)((n + 63) / 64), 64, 0, grid.view(), seg.off, kk, rmax, nb.p, nw.p);
    // 2. Euclidean minimum spanning tree
    DevBuf<int32_t> tree_a, tree_b;
    B3D_TRY(tree_a.alloc(ctx, (size_t)n));
    B3D_TRY(tree_b.alloc(ctx, (size_t)n));
    {
        Forest f;
        B3D_TRY(forest_init(ctx, &f, n, false));
        // the Euclidean rounds kill entries of a private copy of the lists; the Riemannian stage below needs them all again
        DevBuf<int32_t> nb1;
        DevBuf<uint8_t> alive1;
        B3D_TRY(nb1.alloc(ctx, (size_t)n * kk));
        B3D_TRY(alive1.alloc(ctx, (size_t)n));
        B3D_CUDA(cudaMemcpyAsync(nb1.p, nb.p, (size_t)n * kk * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
        B3D_CUDA(cudaMemsetAsync(alive1.p, 1, (size_t)n, ctx->stream));
        EdgeList knn{nullptr, nb1.p, nw.p, n, kk, alive1.p};
        int hooks = 1;
        while (f.edges < n - 1 && hooks > 0) B3D_TRY(boruvka_round(ctx, &f, &knn, 1, nullptr, tree_a.p, tree_b.p, &hooks));
        if (f.edges < n - 1) {
            // the k-NN graph is not connected: every component but the largest looks for its nearest foreign point
            DevBuf<int32_t> size, fnb;
            DevBuf<double> fd;
            DevBuf<unsigned long long> largest, ub;
            DevBuf<double> tile_box;
            DevBuf<int32_t> tile_comp;
            const int n_tiles = (int)((n + kTile - 1) / kTile);
            B3D_TRY(tile_box.alloc(ctx, (size_t)n_tiles * 6));
            B3D_TRY(tile_comp.alloc(ctx, n_tiles));
            B3D_TRY(ub.alloc(ctx, n));
            B3D_TRY(size.alloc(ctx, n));
            B3D_TRY(fnb.alloc(ctx, n));
            B3D_TRY(fd.alloc(ctx, n));
            B3D_TRY(largest.alloc(ctx, 1));
            const int vb = (int)((n + 255) / 256);
            hooks = 1;
            while (f.edges < n - 1 && hooks > 0) {
                B3D_CUDA(cudaMemsetAsync(size.p, 0, n * sizeof(int32_t), ctx->stream));
                B3D_CUDA(cudaMemsetAsync(largest.p, 0, sizeof(unsigned long long), ctx->stream));
                B3D_LAUNCH(ctx, comp_size_kernel, vb, 256, 0, n, f.comp.p, size.p);
                B3D_LAUNCH(ctx, comp_largest_kernel, vb, 256, 0, n, f.comp.p, size.p, largest.p);
                B3D_CUDA(cudaMemsetAsync(ub.p, 0xff, n * sizeof(unsigned long long), ctx->stream));
                const int stride = (int)std::max<int64_t>(1, n / 4096);
                B3D_LAUNCH(ctx, foreign_bound_kernel, (int)((n + 127) / 128), 128, 0, grid.view(), stride, f.comp.p, largest.p, ub.p);
                B3D_LAUNCH(ctx, tile_info_kernel, n_tiles, kTile, 0, grid.view(), f.comp.p, tile_box.p, tile_comp.p);
                B3D_LAUNCH(ctx, foreign_nearest_kernel, n_tiles, kTile, 0, grid.view(), f.comp.p, largest.p, ub.p, tile_box.p, tile_comp.p, n_tiles, fnb.p, fd.p);
                EdgeList bridge{nullptr, fnb.p, fd.p, n, 1, nullptr};
                B3D_TRY(boruvka_round(ctx, &f, &bridge, 1, nullptr, tree_a.p, tree_b.p, &hooks));
            }
        }
        B3D_REQUIRE(f.edges == n - 1, "internal: the Euclidean spanning tree has %d of %lld edges", f.edges, (long long)(n - 1));
    }
    // 3. Riemannian graph = tree + k-NN edges, weights 1 - |n_i . n_j|; its spanning tree with parity tracking
    DevBuf<double> tw;
    B3D_TRY(tw.alloc(ctx, (size_t)n));
    B3D_LAUNCH(ctx, normal_weights_kernel, (int)(((int64_t)n * kk + 255) / 256), 256, 0, nb.p, normals, n, kk, nw.p);
    B3D_LAUNCH(ctx, edge_weights_kernel, (int)((n - 1 + 255) / 256), 256, 0, tree_a.p, tree_b.p, n - 1, normals, tw.p);
    Forest f;
    B3D_TRY(forest_init(ctx, &f, n, true));
    {
        DevBuf<uint8_t> alive2;
        B3D_TRY(alive2.alloc(ctx, (size_t)n));
        B3D_CUDA(cudaMemsetAsync(alive2.p, 1, (size_t)n, ctx->stream));
        EdgeList lists[2] = {{nullptr, nb.p, nw.p, n, kk, alive2.p}, {tree_a.p, tree_b.p, tw.p, n - 1, 1, nullptr}};
        int hooks = 1;
        while (f.edges < n - 1 && hooks > 0) B3D_TRY(boruvka_round(ctx, &f, lists, 2, normals, nullptr, nullptr, &hooks));
    }
    // 4. start point = first point of largest z, turned towards +z; signs follow the tree
    DevBuf<unsigned long long> topz;
    DevBuf<int> top;
    DevBuf<uint8_t> flips;
    B3D_TRY(topz.alloc(ctx, 1));
    B3D_TRY(top.alloc(ctx, 1));
    B3D_CUDA(cudaMemsetAsync(topz.p, 0, sizeof(unsigned long long), ctx->stream));
    B3D_CUDA(cudaMemsetAsync(top.p, 0x7f, sizeof(int), ctx->stream));
    const int rb = ctx->grid_for(n, 256, 4, 8);
    B3D_LAUNCH(ctx, top_z_kernel, rb, 256, 0, xyz, n, topz.p);
    B3D_LAUNCH(ctx, top_z_index_kernel, rb, 256, 0, xyz, n, topz.p, top.p);
    uint8_t* fl = flipped;
    if (fl == nullptr) {
        B3D_TRY(flips.alloc(ctx, n));
        fl = flips.p;
    }
    const int vb = (int)((n + 255) / 256);
    B3D_LAUNCH(ctx, apply_flips_kernel, vb, 256, 0, n, f.comp.p, f.sgn.p, top.p, normals, fl);
    B3D_LAUNCH(ctx, negate_kernel, vb, 256, 0, n, fl, normals);
    return B3D_OK;
}
