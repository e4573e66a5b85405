// b3d_stage2.cuh -- warp-level staging of target points around a compact group of <= 32 queries, round-2 design.
//
// What changed against b3d_stage.cuh (profiles/r01n_ncu_icp_pass_p64_step_digest.txt: 30 % of the stall samples sat on the
// dependent global loads of the staging copy and its float64 box test, at 16 resident warps per SM):
//  * INTEGER GEOMETRY. Every search grid keeps a second, 16-byte copy of its points (Grid<double>::rec, sorted like the
//    float64 points): {ix, iy, iz, sorted position}, the coordinates in fixed point relative to the cloud's lattice origin with
//    2^unit_shift units per cell edge (unit = cell / 2^s, ~6e-9 m on the benchmark grids). A point's cell is ix >> s EXACTLY
//    (the record is floor(u * 2^s) of the same float64 quotient u whose floor is the cell index), so the cells a box overlaps,
//    the box test and the box itself are integer shifts, subtractions and compares; the 32 lanes agree on a box with six
//    single-instruction warp reductions (REDUX) instead of thirty shuffles. Nothing depends on where the cloud lies or how
//    large it is beyond the 31 bits an axis offers.
//  * BULK COPIES. A cell's records are contiguous, so the warp moves whole cells into shared memory with cp.async.bulk (every
//    lane issues the copy of the cell it probed) and waits ONCE on an mbarrier for all of them: no register staging, every copy
//    of a box in flight at the same time.
//  * The filter runs out of shared memory as one flat loop over the copied records (integer box test, ballot compaction IN
//    PLACE) and leaves {float offset from the box centre, |offset|^2} (in units) in groups of 32 slots, component by component
//    (x[32] y[32] z[32] w[32]: the filter's stores and the scan's 16-byte loads are free of bank conflicts), for the
//    dot-product scan: six packed FFMA2 evaluate four candidates.
//  * Boxes that do not fit the buffer are processed in several batches (the scan state lives in registers); only a single cell
//    larger than the buffer, or a box of more than kStage2MaxCells cells, falls back to the per-lane walk.
#pragma once

#include "b3d_search.cuh"

namespace b3d {

template <int CAP>
struct alignas(16) StageSmem {
    static constexpr int kSlots = (CAP + 8 + 31) / 32 * 32;  // whole groups of 32 slots; +8: scan padding
    float4 buf[kSlots];       // raw cell records (int4 bit patterns), then (in place) the filtered candidates in SLOT GROUPS (below)
    int32_t pos[kSlots];      // sorted position of every filtered candidate
    unsigned long long mbar;  // one phase per batch
    unsigned long long pad_;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// one lane of the warp initialises the warp's barrier (count 1) and makes it visible to the async proxy
__device__ __forceinline__ void stage2_init_barrier(unsigned long long* bar) {
    if ((threadIdx.x & 31) == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok = 0;
    uint32_t spins = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a copy that never lands is a bug: fail loudly instead of hanging the device
    }
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); completion counted on bar
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// ---- filtered candidates: slot groups ------------------------------------------------------------------------------------------
// Candidate slot s lives in group s >> 5 (512 bytes: x[32] y[32] z[32] w[32], offsets from the box centre in units, w = |offset|^2)
// at index s & 31. The filter writes consecutive slots to consecutive words (no bank conflicts; the pair blocks of the first
// round-2 version cost 4 wavefronts per store), the scans read four candidates with four 16-byte loads and evaluate
// t = w - 2 q.o for them with six packed FFMA2 (fma.rn.f32x2).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float* cand_slot(float4* buf, int s) { return reinterpret_cast<float*>(buf) + 128 * (s >> 5) + (s & 31); }
__device__ __forceinline__ const float* cand_slot(const float4* buf, int s) { return reinterpret_cast<const float*>(buf) + 128 * (s >> 5) + (s & 31); }
// t of the four candidates gi .. gi + 3 (gi a multiple of 4) for the query factors f = -2 q (each component duplicated into a float2)
// (p: the quad's x entry, i.e. (const float4*)cand_slot(buf, gi))
__device__ __forceinline__ float4 quad_t_at(const float4* __restrict__ p, float2 fx, float2 fy, float2 fz) {
    const float4 X = p[0], Y = p[8], Z = p[16], W = p[24];
    const float2 a = ffma2(fx, make_float2(X.x, X.y), ffma2(fy, make_float2(Y.x, Y.y), ffma2(fz, make_float2(Z.x, Z.y), make_float2(W.x, W.y))));
    const float2 b = ffma2(fx, make_float2(X.z, X.w), ffma2(fy, make_float2(Y.z, Y.w), ffma2(fz, make_float2(Z.z, Z.w), make_float2(W.z, W.w))));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 quad_t(const float4* __restrict__ buf, int gi, float2 fx, float2 fy, float2 fz) {
    return quad_t_at(reinterpret_cast<const float4*>(cand_slot(buf, gi)), fx, fy, fz);
}
// t of the two candidates s, s + 1 (s even)
__device__ __forceinline__ float2 pair_t(const float4* __restrict__ buf, int s, float2 fx, float2 fy, float2 fz) {
    const float2* p = reinterpret_cast<const float2*>(cand_slot(buf, s));
    return ffma2(fx, p[0], ffma2(fy, p[16], ffma2(fz, p[32], p[48])));
}
__device__ __forceinline__ float cand_t(const float4* __restrict__ buf, int s, float2 fx, float2 fy, float2 fz) {
    const float* p = cand_slot(buf, s);
    return fmaf(fx.x, p[0], fmaf(fy.x, p[32], fmaf(fz.x, p[64], p[96])));
}

// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_async_smem() {
#ifndef B3D_STAGE2_NO_FENCE  // measurement only: what the proxy fences cost
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// ---- fixed-point units of a search grid ----------------------------------------------------------------------------------
// unit shift s of a grid whose Morton keys use `shift` bits (shift / 3 per axis): 2^s units per cell, every coordinate below 2^31
__host__ __device__ __forceinline__ int grid_unit_shift(int shift) {
    const int s = 31 - shift / 3;
    return s > 22 ? 22 : (s < 0 ? 0 : s);  // <= 22: box offsets stay below 2^24 units, exactly representable in float32
}
// what a kernel needs of a cloud's lattice to express points in units (warp-uniform)
struct UnitFrame {
    double ox, oy, oz;  // lattice origin
    double per_m;       // units per metre = 2^s / cell
    int nx, ny, nz;     // cells per axis
    int s;              // unit shift
};
__device__ __forceinline__ UnitFrame unit_frame(const Lattice& L, int shift) {
    UnitFrame f;
    f.ox = L.ox; f.oy = L.oy; f.oz = L.oz;
    f.s = grid_unit_shift(shift);
    f.per_m = (double)(1u << f.s) / L.cell;
    f.nx = (int)L.nx; f.ny = (int)L.ny; f.nz = (int)L.nz;  // search grids: at most 2^21 cells per axis
    return f;
}
// The record coordinate of a grid point: floor(u * 2^s) with u = (x - o) / cell, the quotient whose floor is the point's cell
// (lattice_coord, mode 0, k0 = 0 on search grids). The scaling by 2^s is exact, so (record >> s) == cell index, always.
__device__ __forceinline__ int unit_coord_of_point(double x, double o, double cell, int s) {
    const double u = (x - o) / cell;
    return (int)floor(ldexp(u, s));
}
// a query position in units (float64; queries may lie anywhere, also outside the grid)
__device__ __forceinline__ double unit_coord_of_query(double x, double o, double per_m) { return (x - o) * per_m; }
// float64 units -> int32, rounded towards -inf / +inf; the conversion instruction saturates at the int32 range (NaN -> 0)
__device__ __forceinline__ int unit_floor_clamped(double v) { return __double2int_rd(v); }
__device__ __forceinline__ int unit_ceil_clamped(double v) { return __double2int_ru(v); }

constexpr int kStage2MaxCells = 1024;  // cells of one box; beyond that (or beyond 256 on an axis) the caller falls back

// Stages every point of `cloud` with lo <= record <= hi (absolute units, per axis; lo/hi identical on all lanes) batch by batch and
// calls scan(kept) after each batch has been filtered: candidate slots [0, kept) in the slot groups of S.buf (float offsets from the
// box centre c = (lo + hi) >> 1 and their squared length, in units), S.pos[0..kept) = sorted positions, then padding slots at +inf.
// Returns the number of batches scanned (0: the box holds no point), or -1 when the caller has to fall back (box too large or
// one cell with more than CAP points). parity: the warp's mbarrier phase.
template <int CAP, typename Scan>
__device__ __forceinline__ int stage2_run(const GridView<double>& g, const UnitFrame& F, int cloud, int lox, int loy, int loz, int hix, int hiy, int hiz,
                                          StageSmem<CAP>& S, uint32_t& parity, Scan&& scan) {
    const int lane = threadIdx.x & 31;
    // grid points have coordinates in [0, n << s): clamp the box to that range (an empty intersection holds no point)
    lox = max(lox, 0); loy = max(loy, 0); loz = max(loz, 0);
    if (hix < lox || hiy < loy || hiz < loz) return 0;
    const int cx0 = lox >> F.s, cy0 = loy >> F.s, cz0 = loz >> F.s;
    const int cx1 = min(hix >> F.s, F.nx - 1), cy1 = min(hiy >> F.s, F.ny - 1), cz1 = min(hiz >> F.s, F.nz - 1);
    if (cx1 < cx0 || cy1 < cy0 || cz1 < cz0) return 0;
    const int cnx = cx1 - cx0 + 1, cny = cy1 - cy0 + 1, cnz = cz1 - cz0 + 1;
    if (cnx > 256 || cny > 256 || cnz > 256) return -1;
    const int ncell = cnx * cny * cnz;
    if (ncell > kStage2MaxCells) return -1;
    // small exact divisions by float reciprocals (numerators < 1024, divisors <= 256: the error is far below 0.5 / divisor)
    const float rz = 1.0f / (float)cnz, ry = 1.0f / (float)cny;
    const int ccx = (int)(((long long)lox + hix) >> 1), ccy = (int)(((long long)loy + hiy) >> 1), ccz = (int)(((long long)loz + hiz) >> 1);
    const unsigned int ex = (unsigned int)(hix - lox), ey = (unsigned int)(hiy - loy), ez = (unsigned int)(hiz - loz);
    int fill = 0, batches = 0;
    fence_async_smem();  // the caller may have used the buffer through ordinary stores since the last batch
    __syncwarp();

    // One staging loop, ONE place where a batch is consumed (the filter and the caller's scan are inlined there once; four
    // inlined copies made the first version of the round-2 kernels 187 KB large). A round = 32 cells probed by the 32 lanes;
    // `pending` holds a probed round whose copies have not been issued yet (it did not fit the current batch any more);
    // a round that does not fit an EMPTY batch is issued cell by cell (`lane_cursor`).
    int base = 0;
    bool pending = false;
    int lane_cursor = 32;  // < 32: the pending round is being issued one cell at a time, next cell = lane_cursor
    int s = 0, cnt = 0, incl = 0, total = 0;
    int s_nx = 0, cnt_nx = 0;  // the second half of a 64-cell probe, queued as the next round
    bool have_nx = false;
    // cell ci of the box -> its slice of the grid (count 0: empty cell); the first slot load of both halves of a probe are independent
    auto cell_key = [&](int ci) {
        const int t = (int)(((float)ci + 0.5f) * rz);
        const int zc = ci - t * cnz;
        const int xc = (int)(((float)t + 0.5f) * ry);
        const int yc = t - xc * cny;
        return grid_slot_key(g.shift, cloud, cx0 + xc, cy0 + yc, cz0 + zc);
    };
    for (;;) {
        bool cells_left = true;
        for (;;) {
            if (!pending) {
                if (have_nx) {
                    s = s_nx;
                    cnt = cnt_nx;
                    have_nx = false;
                } else {
                    if (base >= ncell) {
                        cells_left = false;
                        break;
                    }
                    // 64 cells per probe, two per lane: the two hash lookups run side by side (one memory latency instead of two
                    // for the typical 3 x 4 x 3-cell box); the second half is queued as the next round
                    const int ci = base + lane, cj = base + 32 + lane;
                    have_nx = base + 32 < ncell;
                    base += 64;
                    s = 0; cnt = 0; s_nx = 0; cnt_nx = 0;
                    int e0 = 0, e1 = 0;
                    grid_lookup2(g, ci < ncell, cell_key(ci), cj < ncell, cell_key(cj), s, e0, s_nx, e1);
                    cnt = e0 - s;
                    cnt_nx = e1 - s_nx;
                }
                if (__any_sync(0xffffffffu, cnt > CAP)) return -1;  // one cell alone overflows the buffer
                incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0) continue;
                pending = true;
                lane_cursor = total > CAP ? 0 : 32;
            }
            if (lane_cursor < 32) {
                // the 32 cells of this round do not fit one batch together: one cell at a time (rare: dense clouds, large cells)
                bool full = false;
                while (lane_cursor < 32) {
                    const int c_l = __shfl_sync(0xffffffffu, cnt, lane_cursor);
                    if (c_l > 0) {
                        if (fill + c_l > CAP) {
                            full = true;
                            break;
                        }
                        if (lane == lane_cursor) {
                            mbar_expect_tx(&S.mbar, (uint32_t)cnt * 16u);
                            bulk_copy_g2s(&S.buf[fill], g.rec + s, (uint32_t)cnt * 16u, &S.mbar);
                        }
                        fill += c_l;
                    }
                    ++lane_cursor;
                }
                __syncwarp();
                if (full) break;
                pending = false;
                lane_cursor = 32;
                continue;
            }
            if (fill + total > CAP) break;  // consume the current batch first (fill > 0 here: total <= CAP)
            if (lane == 0) mbar_expect_tx(&S.mbar, (uint32_t)total * 16u);
            __syncwarp();
            if (cnt > 0) bulk_copy_g2s(&S.buf[fill + incl - cnt], g.rec + s, (uint32_t)cnt * 16u, &S.mbar);
            fill += total;
            pending = false;
            __syncwarp();
        }
        if (fill == 0) break;  // nothing staged and (necessarily) nothing pending
        {
            // ---- consume the batch: wait for its bulk copies, filter + compact in place, hand the candidates to the caller ----
            if (lane == 0) mbar_arrive(&S.mbar);  // the expected bytes were announced copy by copy
            mbar_wait(&S.mbar, parity);
            parity ^= 1u;
            int kept = 0;
            const int4* raw = reinterpret_cast<const int4*>(S.buf);
            for (int j0 = 0; j0 < fill; j0 += 32) {
                const int j = j0 + lane;
                bool inside = false;
                int4 r = make_int4(0, 0, 0, 0);
                if (j < fill) {
                    r = raw[j];
                    inside = (unsigned int)(r.x - lox) <= ex && (unsigned int)(r.y - loy) <= ey && (unsigned int)(r.z - loz) <= ez;
                }
                const unsigned int m = __ballot_sync(0xffffffffu, inside);  // also orders the reads above before the writes below
                if (inside) {
                    const int slot = kept + __popc(m & ((1u << lane) - 1u));
                    const float ox = (float)(r.x - ccx), oy = (float)(r.y - ccy), oz = (float)(r.z - ccz);
                    // in place: the group of slot s overlays the raw records 32 (s >> 5) .. + 31, all of them read already (s <= j0 + 31)
                    float* d = cand_slot(S.buf, slot);
                    d[0] = ox; d[32] = oy; d[64] = oz; d[96] = fmaf(oz, oz, fmaf(oy, oy, ox * ox));
                    S.pos[slot] = r.w;
                }
                kept += __popc(m);
            }
            if (lane < 8) {  // padding candidates at +inf: the scans run in steps of four or eight slots
                float* d = cand_slot(S.buf, kept + lane);
                d[0] = 0.f; d[32] = 0.f; d[64] = 0.f; d[96] = 3.0e38f;
            }
            __syncwarp();
            scan(kept);
            __syncwarp();
            fence_async_smem();  // the next batch's bulk copies overwrite what this batch read and wrote
            __syncwarp();
            fill = 0;
            ++batches;
        }
        if (!cells_left && !pending && !have_nx) break;
    }
    return batches;
}

// Twice the rounding error of a scanned t = |o|^2 - 2 q.o (units^2) against its exact value, for offsets with |component| <= H
// units: the record is floor()ed (1 unit), the float conversions are exact below 2^24 units (6e-8 relative above), the query
// offset is rounded to float32 (6e-8 H), the three FMAs round at 6e-8 of partial sums below 9 H^2:
//   |dt| <= 12 H (1 + 1.2e-7 H) + 2.6e-6 H^2;  two-sided, with margin:
__device__ __forceinline__ float stage2_band(float H) { return 32.0f * H + 1.2e-5f * H * H; }

}  // namespace b3d
