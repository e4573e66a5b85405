// b3d_stage.cuh -- warp-level shared-memory staging of the target points around a compact group of 32 queries.
//
// Why: a per-lane walk of the hashed grid spends its time waiting on dependent loads with 6 of 32 lanes active
// (profiles/r01b_ncu_icp_pass_p8_digest.txt). Here the warp gathers ONCE every target point inside the (dilated)
// bounding box of its 32 queries into shared memory, as float3 offsets from the box centre plus the point's sorted
// position; every lane then scans the same list with uniform control flow and broadcast shared-memory reads.
// The float32 scan is only a pre-filter: candidates within the float rounding band of the best are re-evaluated in
// float64 with the library's exact distance / tie-break rule, so results are bit-identical to the per-lane walk.
#pragma once

#include "b3d_search.cuh"

namespace b3d {

#ifndef B3D_STAGE_CAP
#define B3D_STAGE_CAP 384
#endif
constexpr int kStageCap = B3D_STAGE_CAP;  // staged candidates per warp (float4 each)
constexpr int kStageMaxCells = 256;  // cells one staging call may touch
constexpr int kStagePreCap = 8192;   // points in the touched cells before the box filter (prefix fits 16 bits)

struct StageScratch {  // per warp, shared memory
    int seg_s[kStageMaxCells];                 // first sorted position of the cell's points
    unsigned short seg_off[kStageMaxCells + 2];  // exclusive prefix of the cell sizes
};

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Gathers every point of `cloud` inside the box [lo, hi] (world coordinates, inclusive) into cand[0..count). Two warp-wide phases with all lanes busy and independent loads: (1) one hash
// probe per cell + warp prefix sum of the cell sizes, (2) a flat copy where lane j finds its cell by binary search in the
// prefix. Returns count, or -1 when the box touches more than kStageMaxCells cells / more than kStageCap points (the
// caller falls back to the per-lane walk). All 32 lanes must call it with identical arguments.
__device__ __forceinline__ int warp_stage_box(const GridView<double>& g, int cloud, const double (&lo)[3], const double (&hi)[3],
                                              const double (&center)[3], float4* __restrict__ cand, int* __restrict__ cand_pos, StageScratch* __restrict__ sc,
                                              int cap = kStageCap, double* __restrict__ cand_xyz = nullptr) {
    const Lattice L = g.lat[cloud];
    const int lane = threadIdx.x & 31;
    const double o[3] = {L.ox, L.oy, L.oz};
    const long long k0[3] = {L.kx0, L.ky0, L.kz0};
    const long long nn[3] = {L.nx, L.ny, L.nz};
    const double inv_cell = 1.0 / L.cell;
    long long c0[3], cn[3];
    bool empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // multiply by 1/cell and widen by 1e-6 cells: a superset of the cells the exact division would give
        const double lim = 1.0e9;
        const double u0 = fmin(fmax(floor((lo[a] - o[a]) * inv_cell - 1e-6), -lim), lim);
        const double u1 = fmin(fmax(floor((hi[a] - o[a]) * inv_cell + 1e-6), -lim), lim);
        long long a0 = (long long)u0 - k0[a], a1 = (long long)u1 - k0[a];
        if (a1 < 0 || a0 >= nn[a]) empty = true;
        a0 = a0 < 0 ? 0 : a0;
        a1 = a1 >= nn[a] ? nn[a] - 1 : a1;
        c0[a] = a0;
        cn[a] = a1 - a0 + 1;
    }
    if (empty) return 0;
    if (cn[0] > kStageMaxCells || cn[1] > kStageMaxCells || cn[2] > kStageMaxCells) return -1;
    const int ncell = (int)(cn[0] * cn[1] * cn[2]);
    if (ncell > kStageMaxCells) return -1;
    int total = 0;
    for (int base = 0; base < ncell; base += 32) {
        const int ci = base + lane;
        int s = 0, cnt = 0;
        if (ci < ncell) {
            const int zc = ci % (int)cn[2];
            const int t = ci / (int)cn[2];
            const int yc = t % (int)cn[1], xc = t / (int)cn[1];
            const long long x = c0[0] + xc, y = c0[1] + yc, z = c0[2] + zc;
            const unsigned long long key = grid_slot_key(g.shift, cloud, x, y, z);
            int e;
            if (grid_lookup(g, key, s, e)) cnt = e - s;
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const int off = total + incl - cnt;
        total += __shfl_sync(0xffffffffu, incl, 31);
        if (total > kStagePreCap) return -1;  // warp-uniform
        if (ci < ncell) {
            sc->seg_s[ci] = s;
            sc->seg_off[ci] = (unsigned short)off;
        }
    }
    if (lane == 0) sc->seg_off[ncell] = (unsigned short)total;
    __syncwarp();
    // Copy, keeping only the points inside the box. Lane l owns the contiguous slice [l*m, (l+1)*m) of the flat list of
    // cell points: ONE binary search per lane finds the first cell, after that the lane just walks on to the next
    // non-empty cell. All lanes step through their slices in lock step, so a ballot compacts every step's survivors
    // (deterministic order).
    const int per_lane = (total + 31) >> 5;
    int j = lane * per_lane;
    const int j_end = min(total, j + per_lane);
    int cell = 0, cell_end = 0, src = 0;
    if (j < j_end) {
        int a = 0, b = ncell;  // largest a with seg_off[a] <= j
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if ((int)sc->seg_off[m] <= j) a = m; else b = m;
        }
        cell = a;
        cell_end = (int)sc->seg_off[cell + 1];
        src = sc->seg_s[cell] + (j - (int)sc->seg_off[cell]);
    }
    int kept = 0;
    for (int step = 0; step < per_lane; ++step) {
        bool inside = false;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int pp = 0;
        double ex = 0, ey = 0, ez = 0;
        if (j < j_end) {
            while (j >= cell_end) {  // next non-empty cell
                ++cell;
                cell_end = (int)sc->seg_off[cell + 1];
                src = sc->seg_s[cell];
            }
            const int p = src;
            const double4 pt = ld_point(g.pts + p);
            inside = pt.x >= lo[0] && pt.x <= hi[0] && pt.y >= lo[1] && pt.y <= hi[1] && pt.z >= lo[2] && pt.z <= hi[2];
            const float rx = (float)(pt.x - center[0]), ry = (float)(pt.y - center[1]), rz = (float)(pt.z - center[2]);
            v = make_float4(rx, ry, rz, fmaf(rz, rz, fmaf(ry, ry, rx * rx)));
            pp = p;
            ex = pt.x; ey = pt.y; ez = pt.z;
            ++j;
            ++src;
        }
        const unsigned int m = __ballot_sync(0xffffffffu, inside);
        const int slot = kept + __popc(m & ((1u << lane) - 1u));
        if (inside && slot < cap) {
            cand[slot] = v;
            cand_pos[slot] = pp;
            if (cand_xyz != nullptr) {  // exact coordinates for float64 work on the staged set (normals)
                cand_xyz[3 * slot] = ex;
                cand_xyz[3 * slot + 1] = ey;
                cand_xyz[3 * slot + 2] = ez;
            }
        }
        kept += __popc(m);
    }
    __syncwarp();
    if (kept > cap) return -1;
    total = kept;
    __syncwarp();
    return total;
}

// Nearest staged candidate of q (exact (d2, index) rule), or -1. d2_out / idx_out as nn_within_query; pt_out = the winner's
// point record; others_lb_d2 = a lower bound of the squared distance of every other staged candidate (3e38 when alone).
// Scan in float32 on t = |c|^2 - 2 q.c  (= d2 - |q|^2; offsets from the box centre, |.| <= half_extent): three FMAs per
// candidate. The rounding error of t is below 4e-6 * half_extent^2; every candidate within twice that of the smallest t is
// re-evaluated in float64 with the exact rule, so the result is bit-identical to the per-lane grid walk.
__device__ __forceinline__ int staged_nearest(const GridView<double>& g, const float4* __restrict__ cand, const int* __restrict__ cand_pos,
                                              int count, const double (&center)[3], float half_extent, double qx, double qy, double qz,
                                              double* d2_out, int* idx_out, double4* pt_out = nullptr, double* others_lb_d2 = nullptr) {
    const float fx = -2.0f * (float)(qx - center[0]), fy = -2.0f * (float)(qy - center[1]), fz = -2.0f * (float)(qz - center[2]);
    float best = 3.0e38f, second = 3.0e38f;
    int bi = -1;
#pragma unroll 4
    for (int i = 0; i < count; ++i) {
        const float4 c = cand[i];
        const float t = fmaf(fx, c.x, fmaf(fy, c.y, fmaf(fz, c.z, c.w)));
        const bool nb = t < best;
        second = fminf(second, fmaxf(t, best));
        best = fminf(best, t);
        bi = nb ? i : bi;
    }
    if (bi < 0) return -1;
    const float band = best + 8.0e-6f * half_extent * half_extent;
    int pos = cand_pos[bi];
    double4 pt = ld_point(g.pts + pos);
    double bd = dist2<double>(qx - pt.x, qy - pt.y, qz - pt.z);
    int bidx = point_index(pt);
    const bool ambiguous = second <= band;
    if (ambiguous) {
        // more than one candidate inside the float rounding band of the best: decide in float64
        for (int i = 0; i < count; ++i) {
            const float4 c = cand[i];
            const float t = fmaf(fx, c.x, fmaf(fy, c.y, fmaf(fz, c.z, c.w)));
            if (i != bi && t <= band) {
                const int p2 = cand_pos[i];
                const double4 q2 = ld_point(g.pts + p2);
                const double d2 = dist2<double>(qx - q2.x, qy - q2.y, qz - q2.z);
                const int i2 = point_index(q2);
                if (d2 < bd || (d2 == bd && i2 < bidx)) { bd = d2; bidx = i2; pos = p2; pt = q2; }
            }
        }
    }
    *d2_out = bd;
    *idx_out = bidx;
    if (pt_out != nullptr) *pt_out = pt;
    if (others_lb_d2 != nullptr) {
        // lower bound of the squared distance of every OTHER staged candidate: the second-smallest t of the float scan
        // turned back into a distance (d2 = t + |q - center|^2) minus the rounding band; never below the winner's d2
        double lb = bd;
        if (!ambiguous) {
            const double cx = qx - center[0], cy = qy - center[1], cz = qz - center[2];
            lb = fmax(bd, (cx * cx + cy * cy + cz * cz) + (double)second - 8.0e-6 * (double)half_extent * (double)half_extent);
        }
        *others_lb_d2 = lb;
    }
    return pos;
}

}  // namespace b3d
