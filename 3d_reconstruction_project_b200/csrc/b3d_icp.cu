// b3d_icp.cu -- K4: registration_icp / registration_generalized_icp.
// One kernel per ICP pass fuses: transform of the source point, nearest-neighbour correspondence search in the hashed
// target grid, residual / Jacobian evaluation, the 29-value normal-equation reduction (warp shuffles -> block -> the
// last block sums the per-block partials in a fixed order, no data-path atomics), the 6x6 LDL^T solve (or the
// Umeyama step for point-to-point), the transform update and the convergence test. The loop state lives on the
// device; the host only enqueues passes. A batch of P independent pairs runs in the same launches (grid.y = pair).
#include "b3d_icp.cuh"
#include "b3d_rigid.cuh"
#include "b3d_search.cuh"
#include "b3d_scan.cuh"
#include "b3d_stage.cuh"
#include "b3d_stage2.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b3d {

// debug counters of the staged search (enabled by the environment variable B3D_ICP_STATS=1; see b3d_debug_icp_stats)
__device__ unsigned long long g_icp_stats[8];

namespace {

#ifndef B3D_ICP_BLOCK
#define B3D_ICP_BLOCK 128
#endif
constexpr int kIcpBlock = B3D_ICP_BLOCK;  // threads per block of the pass kernel (a warp works alone; the block only shares the partial sum)
constexpr int kIcpInformation = 3;  // internal kind: G^T G of get_information_matrix_from_point_clouds (rows from the target point)
// Partial-sum groups (blocks) per pair, a function of the pair's own chunk count only (batch == single-pair results). Fewer,
// longer-lived blocks amortise every warp's cold start (four dependent loads before its first chunk is under way); more
// blocks fill the machine when a single pair runs alone. Measured on config 2, ICP ms per 64-pair step / per single pair:
// 128: 20.3 / 1.70, 256: 20.6 / 1.15, 384: 21.1 / 0.97, 512: 21.5 / 0.89, 1024: 23.4 / 0.89.
// Round 2 (contiguous chunk ranges per block): a block's start (state, lattice, first chunk: ~6 dependent loads) and end (fence,
// ticket) cost about 0.7 chunk-times per warp; 8 / 16 / 32 chunks per warp: 17.75 / 16.98 / 17.06 ms per 64-pair step.
#ifndef B3D_ICP_GROUPS
#define B3D_ICP_GROUPS 256
#endif
#ifndef B3D_ICP_CHUNKS_PER_WARP
#define B3D_ICP_CHUNKS_PER_WARP 16
#endif
constexpr int kIcpMinGroups = B3D_ICP_GROUPS * (128 / kIcpBlock);  // partial-sum groups (blocks) of an ordinary pair
constexpr int kIcpMaxGroups = 4096 * (128 / kIcpBlock);            // ... of a very large one
// Groups of a pair with `chunks` warp chunks: one chunk per warp while the pair is tiny, kIcpMinGroups blocks for ordinary
// frame pairs, about B3D_ICP_CHUNKS_PER_WARP (16) chunks per warp beyond that (a 1e8-point cloud must still fill the machine on its own).
__host__ __device__ inline int icp_groups(int chunks) {
    const int wpb = kIcpBlock / 32;
    const int one_each = (chunks + wpb - 1) / wpb;
    int g = (chunks + B3D_ICP_CHUNKS_PER_WARP * wpb - 1) / (B3D_ICP_CHUNKS_PER_WARP * wpb);
    const int floor_g = one_each < kIcpMinGroups ? one_each : kIcpMinGroups;
    g = g < floor_g ? floor_g : g;
    g = g > kIcpMaxGroups ? kIcpMaxGroups : g;
    return g < 1 ? 1 : g;
}
constexpr int kIcpMaxRunLog2 = 6;  // runs of up to 64 chunk groups (1024 chunks)
constexpr double kIcpReach2 = 1.25;   // search radius of a lane that found nothing last time, in units of d_max (see the pass kernel)

// ---- small dense helpers (device) ----------------------------------------------------------------------------------
// Generalized-ICP weight of one correspondence. Open3D (TransformationEstimationForGeneralizedICP::ComputeTransformation) forms
// W = (M^-1)^(1/2) for M = C_t + R C_s R^T and the three rows J = W [ -[p]x | I ], r = W (p - q); what the Gauss-Newton step consumes
// is only J^T J = G^T W^T W G = G^T M^-1 G and J^T r = G^T M^-1 (p - q). ANY factor F with F^T F = M^-1 gives the same two sums, so the
// kernel takes the cheapest one: the inverse of M's Cholesky factor (M = L L^T, F = L^-1, lower triangular) -- three reciprocal
// square roots and a dozen multiplications (~100 instructions) instead of a symmetric square root (closed-form
// eigenvalues through acos / sincos + a Cayley-Hamilton inverse: ~650 executed instructions and 17 KB of code, which is what the
// pass kernel of an 8 MP stereo pair waited for: profiles/r02e_c3_gicp_stalls.txt, instruction-fetch stalls). The covariances GICP
// builds have eigenvalues in [1e-3, 2]: the factorisation is well conditioned. Results differ from the symmetric-root version
// by rounding only (both are exact factorisations of the same M^-1); the CPU oracle keeps the library's formulation.
// Not positive definite / not finite: the diagonal rule (weights 1 / sqrt(M_ii)), as before.
__device__ __forceinline__ void gicp_factor(const double* M, double* F) {
    const double m00 = M[0], m01 = M[1], m02 = M[2], m11 = M[4], m12 = M[5], m22 = M[8];
    const double f00 = rsqrt(m00);             // 1 / l00 (rsqrt: 1 ulp, no slow path -- F^T F stays M^-1 to rounding)
    const double l10 = m01 * f00, l20 = m02 * f00;
    const double p1 = m11 - l10 * l10;
    const double f11 = rsqrt(p1);              // 1 / l11
    const double l21 = (m12 - l20 * l10) * f11;
    const double p2 = m22 - l20 * l20 - l21 * l21;
    const double f22 = rsqrt(p2);              // 1 / l22
    const double f10 = -l10 * f00 * f11;
    const double f21 = -l21 * f11 * f22;
    const double f20 = -(l20 * f00 + l21 * f10) * f22;
    const bool ok = m00 > 0.0 && p1 > 0.0 && p2 > 0.0 && f00 * f11 * f22 < 1.0e300;  // false for NaN as well
    F[0] = ok ? f00 : rsqrt(m00); F[1] = 0.0; F[2] = 0.0;
    F[3] = ok ? f10 : 0.0; F[4] = ok ? f11 : rsqrt(m11); F[5] = 0.0;
    F[6] = ok ? f20 : 0.0; F[7] = ok ? f21 : 0.0; F[8] = ok ? f22 : rsqrt(m22);
}

// ---- per-warp reduction through shared memory ------------------------------------------------------------------------
// Every lane stores the few values its contributions are products of (row: J0..J5, r, 1, d2 for the plane / generalized
// estimators; p, q, 1, d2 for point-to-point). Output sum j is sum over lanes of row[P(j)] * row[Q(j)]; lane j walks the
// 32 rows in lane order and keeps that one running total -- a thread never holds 29 accumulators across the search.
constexpr int kIcpRow = 9;

__device__ __forceinline__ void icp_sum_operands(int kind, int j, int& p, int& q) {
    // row layouts: P2L / GICP: [J0 J1 J2 J3 J4 J5 r one d2]   P2P: [px py pz qx qy qz - one d2]
    p = 7;
    q = 7;  // default: one * one (j == 27, the correspondence count; also harmless for j >= 29)
    if (j == 28) { p = 8; q = 7; return; }
    if (j >= 27) return;
    if (kind == B3D_ICP_POINT_TO_POINT) {
        if (j < 6) { p = j; q = 7; }                               // sums of p and of q
        else if (j < 15) { p = 3 + (j - 6) / 3; q = (j - 6) % 3; } // q_r * p_c
        else { p = 6; q = 6; }                                      // unused slots: 0 * 0
        return;
    }
    if (j < 21) {
        int u = 0, rem = j;
        while (rem >= 6 - u) { rem -= 6 - u; ++u; }
        p = u;
        q = u + rem;
    } else {
        p = j - 21;
        q = 6;
    }
}

// ---- end of a pass: fitness / rmse / convergence / update (one thread per pair) -------------------------------------
__device__ __noinline__ void icp_finalize_pair(int kind, const double* a, double ns, double rel_fitness, double rel_rmse, int max_iter, IcpPairState* st) {
    const double nc = a[27];
    const double fitness = (nc == 0 || ns == 0) ? 0.0 : nc / ns;
    const double rmse = nc == 0 ? 0.0 : sqrt(a[28] / nc);
    st->fitness = fitness;
    st->rmse = rmse;
    st->n_corr = (long long)nc;
    const int k = st->iter;
    if (k > 0 && fabs(st->prev_fitness - fitness) < rel_fitness && fabs(st->prev_rmse - rmse) < rel_rmse) {
        st->done = 1;
        st->converged = 1;
        return;
    }
    if (k >= max_iter) {
        st->done = 1;
        return;
    }
    double U[16];
    mat4_identity(U);
    if (nc > 0) {
        if (kind == B3D_ICP_POINT_TO_POINT) {
            const double mu_s[3] = {a[0] / nc, a[1] / nc, a[2] / nc}, mu_d[3] = {a[3] / nc, a[4] / nc, a[5] / nc};
            double Sigma[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) Sigma[3 * r + c] = a[6 + 3 * r + c] / nc - mu_d[r] * mu_s[c];
            umeyama_from_moments(mu_s, mu_d, Sigma, U);
        } else {
            double A[36], b[6], x[6];
            int t = 0;
            for (int u = 0; u < 6; ++u)
                for (int v = u; v < 6; ++v) { A[6 * u + v] = a[t]; A[6 * v + u] = a[t]; ++t; }
            for (int u = 0; u < 6; ++u) b[u] = -a[21 + u];
            if (solve6(A, b, x)) vec6_to_mat4(x, U);
        }
    }
    double Tn[16];
    mat4_mul(U, st->T, Tn);
    for (int i = 0; i < 16; ++i) st->T[i] = Tn[i];
    st->prev_fitness = fitness;
    st->prev_rmse = rmse;
    st->iter = k + 1;
}

struct IcpKernelArgs {
    int kind;
    const double4* src_sorted;
    const int32_t* chunk_start;  // [n_chunks + 1] first sorted source position of every warp chunk (<= 32 points, compact)
    const int32_t* chunk_off;    // [P + 1] first chunk of every pair
    int32_t n_chunks;
    const double* src_cov;
    const int32_t* src_off;
    const int64_t* ns_global;
    GridView<double> grid;
    const int32_t* tgt_off;
    const double* tgt_nrm_sorted;
    const double* tgt_cov_sorted;
    double r2;
    int rmax;
    double rel_fitness, rel_rmse;
    int max_iter;
    IcpPairState* state;
    double* partial;
    double* sums;
    int32_t* corr;
    int fused;
    // sticky correspondences: per source point (sorted position) the nearest target found by its last full search, the
    // transformed position it was searched from and a lower bound of the distance of every OTHER target point at that
    // time. While the point has moved less than the slack between the two, the nearest neighbour cannot have changed
    // and the lane skips the search (exact: same partner, distance recomputed in float64).
    float4* keep_ref;   // [ns] x, y, z of the query at search time; w = lower bound of the runner-up distance (0 = none)
    int32_t* keep_pos;  // [ns] sorted target position of the nearest neighbour, -1 = nothing within the searched reach
    // peer exchange (one cloud sharded over several GPUs, P == 1): after the local second pass the last block writes its 29
    // sums into every rank's exchange buffer over NVLink, waits for all ranks' flags of this pass and adds the slots in rank
    // order -- the all-reduce happens inside the pass kernel, no collective launch in between
    int peer_world, peer_rank;
    double* peer_buf[8];
    int stats;  // count chunks / rounds / staged candidates into g_icp_stats
    int run_log2;    // round-2 kernel: a run = 2^run_log2 adjacent chunk groups (the level of the sum tree a launch's blocks work at)
    int run_stride;  // round-2 kernel: nodes reserved per pair and strand in `partial` (>= every pair's run count)
};

// Source points are visited in the order of the TARGET cell they fall into (sorted once per ICP run, icp_prepare), so the
// 32 queries of a warp share their cells: the hash probes and candidate loads hit L1 and the lanes take similar paths.
// Every lane produces its 29 contributions; a transpose-reduce leaves value j's warp total in lane j, which is the only
// accumulator a thread keeps (instead of 29 live doubles across the search loop).
#ifndef B3D_ICP_MIN_BLOCKS
#define B3D_ICP_MIN_BLOCKS (512 / B3D_ICP_BLOCK)
#endif
template <int KIND>
__global__ void __launch_bounds__(kIcpBlock, B3D_ICP_MIN_BLOCKS) icp_pass_kernel(IcpKernelArgs A) {
    const int pair = blockIdx.y;
    IcpPairState* st = A.state + pair;
    __shared__ double sT[16];
    __shared__ double sm[kIcpBlock / 32][32];
    __shared__ double srow[kIcpBlock / 32][32][kIcpRow];
    __shared__ float4 s_cand[kIcpBlock / 32][kStageCap];
    __shared__ int s_cand_pos[kIcpBlock / 32][kStageCap];
    __shared__ StageScratch s_stage[kIcpBlock / 32];
    __shared__ int s_last;
    // one round trip for everything the block needs to start: the loop state (done flag, transform) and the pair's ranges
    // are all requested before the first wait
    const int done = st->done;
    if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
    const int32_t s0 = A.src_off[pair], s1 = A.src_off[pair + 1];
    const int32_t t0 = A.tgt_off[pair];
    const int32_t c0 = A.chunk_off[pair], c1 = A.chunk_off[pair + 1];
    if (done) return;  // uniform over the block
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int op_p, op_q;
    icp_sum_operands(KIND, lane, op_p, op_q);
    double (*rows)[kIcpRow] = srow[warp];
    double acc = 0.0;  // lane j: running total of sum j
    const double dmax = sqrt(A.r2);
    float4* cand = s_cand[warp];
    int* cand_pos = s_cand_pos[warp];
    // affine transforms (last row 0 0 0 1) need no perspective division: x / 1.0 == x exactly
    const bool affine = sT[12] == 0.0 && sT[13] == 0.0 && sT[14] == 0.0 && sT[15] == 1.0;
    // this pair's range of warp chunks (chunks never straddle pairs) and its number of partial-sum groups: a function of
    // the pair's own chunk count only, so a pair's result does not depend on what else is in the batch
    const int groups = icp_groups(c1 - c0);
    if ((int)blockIdx.x >= groups) return;
    // the next chunk's query is fetched while the current one is processed (two dependent loads off the critical path)
    const int32_t c_step = groups * (kIcpBlock / 32);
    int32_t c = c0 + blockIdx.x * (kIcpBlock / 32) + warp;
    double4 nsp = make_double4(0.0, 0.0, 0.0, 0.0);
    float4 nkr = make_float4(0.f, 0.f, 0.f, 0.f);
    int nkp = -1;
    int32_t ni = 0;
    bool nvalid = false;
    if (c < c1) {
        ni = A.chunk_start[c] + lane;
        nvalid = ni < A.chunk_start[c + 1];
        if (nvalid) {
            nsp = ld_point(A.src_sorted + ni);
            if (A.keep_ref != nullptr) {
                nkr = A.keep_ref[ni];
                nkp = A.keep_pos[ni];
            }
        }
    }
    for (; c < c1; c += c_step) {
        const double4 sp = nsp;
        const float4 kr = nkr;
        const int kp = nkp;
        const int32_t si = ni;
        const bool valid = nvalid;
        nvalid = false;
        if (c + c_step < c1) {
            ni = A.chunk_start[c + c_step] + lane;
            nvalid = ni < A.chunk_start[c + c_step + 1];
            if (nvalid) {
                nsp = ld_point(A.src_sorted + ni);
                if (A.keep_ref != nullptr) {
                    nkr = A.keep_ref[ni];
                    nkp = A.keep_pos[ni];
                }
            }
        }
        double e[kIcpRow];
#pragma unroll
        for (int j = 0; j < kIcpRow; ++j) e[j] = 0.0;
        double W[9], gd[3], gp[3];  // generalized ICP only
        bool matched = false;
        double px = 0, py = 0, pz = 0;
        int oi = 0;
        if (valid) {
            oi = point_index(sp);  // original (batch-global) source index
            const double x = sp.x, y = sp.y, z = sp.z;
            // PointCloud::Transform: (T [p,1]).xyz / w
            px = sT[0] * x + sT[1] * y + sT[2] * z + sT[3];
            py = sT[4] * x + sT[5] * y + sT[6] * z + sT[7];
            pz = sT[8] * x + sT[9] * y + sT[10] * z + sT[11];
            if (!affine) {
                const double w = sT[12] * x + sT[13] * y + sT[14] * z + sT[15];
                px /= w; py /= w; pz /= w;
            }
        }
        // ---- correspondence: sticky check, then a staged warp search bounded by the previous partner -----------------
        // (bit-identical to nn_within_query). A lane that had a partner in the last pass knows an upper bound of its
        // nearest-neighbour distance -- the distance u to that partner now -- so its search ball has radius u (+ a slack
        // that buys the sticky test room for the next pass) instead of d_max; one round, no guessing.
        double d2 = 0.0;
        int idx = 0, pos = -1;
        double4 q = make_double4(0.0, 0.0, 0.0, 0.0);  // the partner's point record
        bool need = valid;
        double reach = dmax;  // this lane's search radius
        if (valid && A.keep_ref != nullptr) {
            const double mx = px - (double)kr.x, my = py - (double)kr.y, mz = pz - (double)kr.z;
            // movement since the last search (+ the float rounding of the stored position)
            const double moved = sqrt(mx * mx + my * my + mz * mz) + 2.0e-7 * (fabs(px) + fabs(py) + fabs(pz));
            const double lim = (double)kr.w;  // every other target point was at least this far from the stored position (0: unknown)
            if (kp >= 0) {
                q = ld_point(A.grid.pts + kp);
                const double dk = dist2<double>(px - q.x, py - q.y, pz - q.z);
                const double u = sqrt(dk);
                if (u * (1.0 + 1e-12) + moved < lim) {  // still strictly nearer than anything else can be
                    need = false;
                    pos = kp;
                    d2 = dk;
                    idx = point_index(q);
                } else {
                    const double slack = fmin(fmax(0.5 * moved, 0.01 * dmax), 0.1 * dmax);
                    reach = fmin(u * (1.0 + 1e-9) + slack, dmax);  // nothing beyond d_max counts anyway
                }
            } else if (lim > 0.0) {
                if (dmax * (1.0 + 1e-12) + moved < lim) need = false;  // nothing was within lim, nothing can be within d_max now
                // dilated once the cloud has nearly stopped moving: if the lane finds nothing again it keeps some slack
                else if (moved < 0.5 * (kIcpReach2 - 1.0) * dmax) reach = dmax * kIcpReach2;
            }
        }
        if (A.stats) {
            const unsigned int nm = __ballot_sync(0xffffffffu, need);
            if (lane == 0) {
                if (nm == 0u) atomicAdd(&g_icp_stats[6], 1ull);
                atomicAdd(&g_icp_stats[7], (unsigned long long)__popc(nm));
            }
        }
        if (__any_sync(0xffffffffu, need)) {
            // bounding box of the search balls: reduced in float32 (half the shuffles), widened by the float rounding
            const float bigf = 3.0e38f;
            const float rf = (float)reach * 1.000001f;
            float lx = need ? (float)px - rf : bigf, ly = need ? (float)py - rf : bigf, lz = need ? (float)pz - rf : bigf;
            float hx = need ? (float)px + rf : -bigf, hy = need ? (float)py + rf : -bigf, hz = need ? (float)pz + rf : -bigf;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o));
                ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
                lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
                hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
                hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
                hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
            }
            const double wid = 4.0e-7;  // relative rounding of the float conversions and the float add, with margin
            const double pad = 1e-12;
            const double lo[3] = {(double)lx - wid * fabs((double)lx) - pad, (double)ly - wid * fabs((double)ly) - pad, (double)lz - wid * fabs((double)lz) - pad};
            const double hi[3] = {(double)hx + wid * fabs((double)hx) + pad, (double)hy + wid * fabs((double)hy) + pad, (double)hz + wid * fabs((double)hz) + pad};
            const double center[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
            const float half_extent = (float)(0.5 * fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2])) * 1.0001f;
            double others2 = 3.0e38;  // lower bound of the squared distance of every staged candidate but the winner
            bool bounded = true;      // every target point within `reach` of the query was looked at
            const int count = warp_stage_box(A.grid, pair, lo, hi, center, cand, cand_pos, &s_stage[warp]);
            if (A.stats && lane == 0) {
                atomicAdd(&g_icp_stats[0], 1ull);
                if (count < 0) atomicAdd(&g_icp_stats[2], 1ull);
                else atomicAdd(&g_icp_stats[3], (unsigned long long)count);
                const double ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
                atomicAdd(&g_icp_stats[4], (unsigned long long)(1.0e6 * ex * ey * ez));  // box volume in cm^3
                atomicAdd(&g_icp_stats[5], (unsigned long long)(1.0e4 * fmax(fmax(ex, ey), ez)));  // longest edge in 0.1 mm
            }
            if (count < 0) {
                // the box is too crowded for the staging buffer: per-lane walk of the grid (no runner-up bound)
                if (need) {
                    pos = nn_within_query<double>(A.grid, pair, px, py, pz, A.r2, A.rmax, &d2, &idx);
                    if (pos >= 0) q = ld_point(A.grid.pts + pos);
                }
                bounded = false;
            } else if (need) {
                pos = staged_nearest(A.grid, cand, cand_pos, count, center, half_extent, px, py, pz, &d2, &idx, &q, &others2);
                if (pos < 0) others2 = 3.0e38;
            }
            __syncwarp();
            if (need && A.keep_ref != nullptr) {
                // what this search proved: the nearest point (if any within reach) and that every other point is at least
                // min(runner-up, reach) away; stored rounded down
                float lbf = 0.f;
                if (bounded) lbf = (float)(fmin(sqrt(others2), reach) * (1.0 - 1.0e-6));
                A.keep_ref[si] = make_float4((float)px, (float)py, (float)pz, lbf);
                A.keep_pos[si] = pos;
            }
        }
        if (pos >= 0 && !(d2 < A.r2)) pos = -1;
        if (valid) {
            if (A.corr != nullptr) A.corr[oi] = pos >= 0 ? idx - t0 : -1;
            if (pos >= 0) {
                matched = true;
                e[7] = 1.0;
                e[8] = d2;
                if (KIND == kIcpInformation) {
                    gp[0] = q.x; gp[1] = q.y; gp[2] = q.z;
                } else if (KIND == B3D_ICP_POINT_TO_POINT) {
                    e[0] = px; e[1] = py; e[2] = pz;
                    e[3] = q.x; e[4] = q.y; e[5] = q.z;
                } else if (KIND == B3D_ICP_POINT_TO_PLANE) {
                    const double* nq = A.tgt_nrm_sorted + 3 * (int64_t)pos;
                    const double n0 = __ldg(nq), n1 = __ldg(nq + 1), n2 = __ldg(nq + 2);
                    e[0] = py * n2 - pz * n1; e[1] = pz * n0 - px * n2; e[2] = px * n1 - py * n0;
                    e[3] = n0; e[4] = n1; e[5] = n2;
                    e[6] = (px - q.x) * n0 + (py - q.y) * n1 + (pz - q.z) * n2;
                } else {
                    // generalized ICP: M = C_t + R C_s R^T, W = (M^-1)^(1/2), rows r_k = W_k (p - q), J = W [ -[p]x | I ]
                    const double* Ct = A.tgt_cov_sorted + 9 * (int64_t)pos;
                    const double* Cs = A.src_cov + 9 * (int64_t)oi;
                    double RC[9], M[9];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) RC[3 * r + c] = sT[4 * r] * Cs[c] + sT[4 * r + 1] * Cs[3 + c] + sT[4 * r + 2] * Cs[6 + c];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            M[3 * r + c] = __ldg(Ct + 3 * r + c) + (RC[3 * r] * sT[4 * c] + RC[3 * r + 1] * sT[4 * c + 1] + RC[3 * r + 2] * sT[4 * c + 2]);
                    gicp_factor(M, W);
                    gd[0] = px - q.x; gd[1] = py - q.y; gd[2] = pz - q.z;
                    gp[0] = px; gp[1] = py; gp[2] = pz;
                }
            }
        }
        const int n_rows = (KIND == B3D_ICP_GENERALIZED || KIND == kIcpInformation) ? 3 : 1;
        for (int row = 0; row < n_rows; ++row) {
            if (KIND == kIcpInformation) {
                if (matched) {
                    // G = [ -[t]x | I ] for the target point t (kept in gp)
                    e[0] = row == 0 ? 0.0 : (row == 1 ? -gp[2] : gp[1]);
                    e[1] = row == 0 ? gp[2] : (row == 1 ? 0.0 : -gp[0]);
                    e[2] = row == 0 ? -gp[1] : (row == 1 ? gp[0] : 0.0);
                    e[3] = row == 0 ? 1.0 : 0.0; e[4] = row == 1 ? 1.0 : 0.0; e[5] = row == 2 ? 1.0 : 0.0;
                    e[6] = 0.0;
                    if (row > 0) { e[7] = 0.0; e[8] = 0.0; }
                }
            }
            if (KIND == B3D_ICP_GENERALIZED) {
                if (matched) {
                    const double w0 = W[3 * row], w1 = W[3 * row + 1], w2 = W[3 * row + 2];
                    // J = W_row [ -[p]x | I ],  -[p]x = [0 pz -py; -pz 0 px; py -px 0]
                    e[0] = w1 * (-gp[2]) + w2 * gp[1];
                    e[1] = w0 * gp[2] + w2 * (-gp[0]);
                    e[2] = w0 * (-gp[1]) + w1 * gp[0];
                    e[3] = w0; e[4] = w1; e[5] = w2;
                    e[6] = w0 * gd[0] + w1 * gd[1] + w2 * gd[2];
                    if (row > 0) { e[7] = 0.0; e[8] = 0.0; }
                }
            }
#pragma unroll
            for (int j = 0; j < kIcpRow; ++j) rows[lane][j] = e[j];
            __syncwarp();
            if ((KIND == B3D_ICP_GENERALIZED || KIND == kIcpInformation) && row > 0 && lane >= 27) {
                // count and sum d2 are taken once per correspondence (row 0)
            } else {
#pragma unroll 8
                for (int l = 0; l < 32; ++l) acc += rows[l][op_p] * rows[l][op_q];
            }
            __syncwarp();
        }
    }
    // block: lane j of every warp holds sum j -> fixed-order sum over the warps -> per-block partial
    sm[warp][lane] = acc;
    __syncthreads();
    double* part = A.partial + ((int64_t)pair * gridDim.x + blockIdx.x) * kIcpSums;
    if (threadIdx.x < kIcpSums) {
        double v = sm[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kIcpBlock / 32; ++w) v += sm[w][threadIdx.x];
        part[threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&st->ticket, 1u);
        s_last = (t == (unsigned int)groups - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // last block of this pair: second pass over the per-block partials in a fixed order (warp w takes blocks w, w+4, ...;
    // lane j sums column j), no atomics on the data path
    __threadfence();
    const double* base = A.partial + (int64_t)pair * gridDim.x * kIcpSums;
    double v = 0.0;
    if (lane < kIcpSums)
        for (int b = warp; b < groups; b += kIcpBlock / 32) v += __ldcg(base + (int64_t)b * kIcpSums + lane);
    sm[warp][lane] = v;
    __syncthreads();
    if (threadIdx.x < kIcpSums) {
        double t = sm[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kIcpBlock / 32; ++w) t += sm[w][threadIdx.x];
        sm[0][threadIdx.x] = t;
        A.sums[(int64_t)pair * kIcpSums + threadIdx.x] = t;
    }
    __syncthreads();
    if (A.peer_world > 1) {
        // ---- all-reduce over peer memory. Buffer of a rank: slots [2 parities][world][32] doubles, then flags [2][world]
        // (pass number + 1 as a double). Two parities: a rank can be at most one pass ahead of the slowest reader.
        const int W = A.peer_world;
        const unsigned int pass = st->pass_id;
        const int parity = (int)(pass & 1u);
        const double stamp = (double)(pass + 1u);
        if (threadIdx.x < kIcpSums) {
            const double mine = sm[0][threadIdx.x];
            for (int r = 0; r < W; ++r) {
                volatile double* slot = A.peer_buf[r] + ((int64_t)parity * W + A.peer_rank) * 32;
                slot[threadIdx.x] = mine;
            }
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < W) {
            volatile double* flag = A.peer_buf[threadIdx.x] + (int64_t)2 * W * 32 + (int64_t)parity * W + A.peer_rank;
            *flag = stamp;
        }
        __shared__ int s_timeout;
        if (threadIdx.x == 0) s_timeout = 0;
        __syncthreads();
        if ((int)threadIdx.x < W) {
            volatile double* flag = A.peer_buf[A.peer_rank] + (int64_t)2 * W * 32 + (int64_t)parity * W + threadIdx.x;
            const long long t0c = clock64();
            while (*flag != stamp) {
                if (clock64() - t0c > 4000000000ll) {  // ~2 s: a peer never arrived
                    s_timeout = 1;
                    break;
                }
            }
        }
        __syncthreads();
        __threadfence_system();
        if (threadIdx.x < kIcpSums) {
            double t = 0.0;
            for (int r = 0; r < W; ++r) {
                volatile double* slot = A.peer_buf[A.peer_rank] + ((int64_t)parity * W + r) * 32;
                t += slot[threadIdx.x];
            }
            sm[0][threadIdx.x] = t;
            A.sums[(int64_t)pair * kIcpSums + threadIdx.x] = t;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            st->pass_id = pass + 1u;
            if (s_timeout) {
                st->done = 1;
                st->converged = -1;  // exchange timed out
            }
        }
        if (s_timeout) return;
    }
    if (threadIdx.x == 0) {
        st->ticket = 0;
#ifndef B3D_TEST_NO_FINALIZE
        if (A.fused) {
            double a[kIcpSums];
#pragma unroll
            for (int j = 0; j < kIcpSums; ++j) a[j] = sm[0][j];
            const double ns = A.ns_global ? (double)A.ns_global[pair] : (double)(s1 - s0);
            icp_finalize_pair(A.kind, a, ns, A.rel_fitness, A.rel_rmse, A.max_iter, st);
        }
#endif
    }
}

// ======================================================================================================================
// Round-2 pass kernel. Same algorithm and the same results as icp_pass_kernel above (the float scan is a pre-filter; winners
// and near-ties are decided in float64 with the library's (d2, index) rule), different machinery:
//  * staging through b3d_stage2.cuh: integer fixed-point geometry (boxes, cells and the box filter are shifts, subtractions and
//    compares; the box of a chunk is six REDUX instructions), cp.async.bulk copies of whole grid cells into shared memory, one
//    mbarrier wait per batch, filter and in-place compaction out of shared memory;
//  * the scan takes four candidates per step and keeps (best, runner-up, winning group) with 3 ALU operations per
//    candidate instead of 5; the winner inside the group is resolved once after the loop;
//  * a box that does not fit the buffer is scanned in several batches instead of falling back to the per-lane walk;
//  * warps are independent to the end: every warp hands its 29 sums to the block through shared memory and leaves; the last
//    warp of a block to arrive adds the block's rows, the last block of a pair adds the pair's rows (fixed orders, no atomics
//    on the data path, no block barrier after the start-up);
//  * everything that runs rarely (per-lane walk, float64 tie resolution, the final reduction, the peer exchange, the solve) is
//    out of line: the round-2 kernel first measured at 187 KB of code and spent a third of its issue slots waiting for
//    instructions (profiles/r02a_ncu_icp_pass2_p16_digest.txt).
#ifndef B3D_ICP2_CAP
#define B3D_ICP2_CAP 380
#endif
#ifndef B3D_ICP2_MIN_BLOCKS
#define B3D_ICP2_MIN_BLOCKS 4
#endif
constexpr int kIcp2Cap = B3D_ICP2_CAP;  // raw cell records per batch (+4 scan padding = 384 x 16 bytes)
using Icp2Smem = StageSmem<kIcp2Cap>;
// The reduction's row buffer aliases the candidate buffer, TRANSPOSED: value j of lane l at [j][l], rows 34 doubles apart (272 bytes:
// the eight distinct operand rows a warp reads at once start in eight different 16-byte bank groups). Lane j fetches its two operands
// for TWO source lanes with one 16-byte load each: 32 loads + 32 fused multiply-adds per chunk, one wavefront per load (the [l][j]
// layout took 64 8-byte loads at 1.8 wavefronts, a quarter of the kernel's shared-memory traffic).
constexpr int kIcpRowStride = 34;
static_assert(sizeof(float4) * Icp2Smem::kSlots >= sizeof(double) * kIcpRowStride * kIcpRow, "the row buffer of the reduction aliases the candidate buffer");

// cold: exact per-lane walk of the grid (box too large to stage, or a near-tie in a box that took several batches). The cold helpers
// take and return VALUES: a variable whose address is passed to an out-of-line function lives in local memory for the whole
// kernel (px, py, pz, d2, idx and q did: ~17 local loads / stores per chunk, profiles/r02c_ncu_icp_pass2_p64_digest.txt).
__device__ __noinline__ int icp2_walk(const GridView<double>& g, int pair, double px, double py, double pz, double r2, int rmax) {
    double d2;
    int idx;
    return nn_within_query<double>(g, pair, px, py, pz, r2, rmax, &d2, &idx);
}

// cold: more than one staged candidate inside the rounding band of the best -- decide in float64 with the (d2, index) rule
// (d2, idx: the float winner's). Returns the sorted position of the exact winner.
__device__ __noinline__ int icp2_resolve_ties(const double4* __restrict__ pts, const float4* __restrict__ buf, const int* __restrict__ posb, int kept,
                                              float fx, float fy, float fz, float lim_t, double px, double py, double pz, int wpos, double d2, int idx) {
    int pos = wpos;
    const float2 f2x = make_float2(fx, fx), f2y = make_float2(fy, fy), f2z = make_float2(fz, fz);
    for (int i = 0; i < kept; ++i) {
        if (cand_t(buf, i, f2x, f2y, f2z) <= lim_t) {
            const int p2 = posb[i];
            if (p2 == wpos) continue;
            const double4 q2 = ld_point(pts + p2);
            const double e2 = dist2<double>(px - q2.x, py - q2.y, pz - q2.z);
            const int i2 = point_index(q2);
            if (e2 < d2 || (e2 == d2 && i2 < idx)) { d2 = e2; idx = i2; pos = p2; }
        }
    }
    return pos;
}

// cold: the last block of a pair -- warp w adds strand w's run nodes up to the strand's root (lane j: sum j). Node i of the next level
// = node 2i + node 2i+1 (an odd last node moves up alone), in place; sixteen outputs (32 loads) are in flight per step.
__device__ __noinline__ double icp2_reduce_strand(const IcpKernelArgs& A, int pair, int n_runs, int strand) {
    const int lane = threadIdx.x & 31;
    if (lane >= kIcpSums) return 0.0;
    const int64_t step = 4 * kIcpSums;  // doubles between consecutive runs of one strand
    double* col = A.partial + ((int64_t)pair * A.run_stride * 4 + strand) * kIcpSums + lane;
    int n = n_runs;
    while (n > 1) {
        const int nn = (n + 1) >> 1;
        for (int i0 = 0; i0 < nn; i0 += 16) {
            double t[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int j = 2 * i0 + k;
                t[k] = j < n ? __ldcg(col + (int64_t)j * step) : 0.0;
            }
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                if (i0 + o < nn) {
                    double v = t[2 * o];
                    if (2 * (i0 + o) + 1 < n) v += t[2 * o + 1];
                    __stcg(col + (int64_t)(i0 + o) * step, v);
                }
            }
        }
        n = nn;
    }
    return __ldcg(col);
}

// cold: warp 0 of the last block of a pair with the pair's 29 sums (lane j: sum j) -- the optional all-reduce over peer memory, and
// the solve / update of the loop state
__device__ __noinline__ void icp2_finish_pair(const IcpKernelArgs& A, int pair, double total, int ns_local) {
    const int lane = threadIdx.x & 31;
    IcpPairState* st = A.state + pair;
    if (lane < kIcpSums) A.sums[(int64_t)pair * kIcpSums + lane] = total;
    __syncwarp();
    if (A.peer_world > 1) {
        // ---- all-reduce over peer memory. Buffer of a rank: slots [2 parities][world][32] doubles, then flags [2][world]
        // (pass number + 1 as a double). Two parities: a rank can be at most one pass ahead of the slowest reader.
        // Stores of the sums (relaxed, system scope), then the flag with release semantics; the reader acquires the flag.
        const int W = A.peer_world;
        const unsigned int pass = st->pass_id;
        const int par = (int)(pass & 1u);
        const double stamp = (double)(pass + 1u);
        if (lane < kIcpSums) {
            for (int r = 0; r < W; ++r) {
                double* slot = A.peer_buf[r] + ((int64_t)par * W + A.peer_rank) * 32 + lane;
                asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(slot), "d"(total) : "memory");
            }
        }
        __threadfence_system();
        __syncwarp();
        if (lane < W) {
            double* flag = A.peer_buf[lane] + (int64_t)2 * W * 32 + (int64_t)par * W + A.peer_rank;
            asm volatile("st.release.sys.global.f64 [%0], %1;" ::"l"(flag), "d"(stamp) : "memory");
        }
        int timeout = 0;
        if (lane < W) {
            const double* flag = A.peer_buf[A.peer_rank] + (int64_t)2 * W * 32 + (int64_t)par * W + lane;
            const long long t0c = clock64();
            while (true) {
                double v;
                asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(flag) : "memory");
                if (v == stamp) break;
                if (clock64() - t0c > 20000000000ll) {  // ~10 s (a rank may lag by a module load on a cold box): a peer never arrived
                    timeout = 1;
                    break;
                }
            }
        }
        timeout = __any_sync(0xffffffffu, timeout) ? 1 : 0;
        if (lane < kIcpSums) {
            double t = 0.0;
            for (int r = 0; r < W; ++r) {
                const double* slot = A.peer_buf[A.peer_rank] + ((int64_t)par * W + r) * 32 + lane;
                double v;
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(slot) : "memory");
                t += v;
            }
            total = t;
            A.sums[(int64_t)pair * kIcpSums + lane] = t;
        }
        __syncwarp();
        if (lane == 0) {
            st->pass_id = pass + 1u;
            if (timeout) {
                st->done = 1;
                st->converged = -1;  // exchange timed out
            }
        }
        if (timeout) {
            if (lane == 0) st->ticket = 0;
            return;
        }
    }
    // every lane gathers the 29 totals (lane 0 solves)
    double a[kIcpSums];
#pragma unroll
    for (int j = 0; j < kIcpSums; ++j) a[j] = __shfl_sync(0xffffffffu, total, j);
    if (lane == 0) {
        st->ticket = 0;
#ifndef B3D_TEST_NO_FINALIZE
        if (A.fused) {
            const double ns = A.ns_global ? (double)A.ns_global[pair] : (double)ns_local;
            icp_finalize_pair(A.kind, a, ns, A.rel_fitness, A.rel_rmse, A.max_iter, st);
        }
#endif
    }
}

// cold: transforms with a projective last row (PointCloud::Transform divides by w)
__device__ __noinline__ double icp2_perspective_w(const double* T, double x, double y, double z) {
    return T[12] * x + T[13] * y + T[14] * z + T[15];
}

template <int KIND>
__global__ void __launch_bounds__(kIcpBlock, B3D_ICP2_MIN_BLOCKS) icp_pass2_kernel(const __grid_constant__ IcpKernelArgs A) {
    extern __shared__ __align__(16) unsigned char icp2_smem[];
    const int pair = blockIdx.y;
    IcpPairState* st = A.state + pair;
    __shared__ double sT[16];
    static_assert(kIcpBlock == 128, "the sum tree is built on four warps per block");
    __shared__ double stk[kIcpMaxRunLog2 + 1][kIcpBlock / 32][32];  // [level][warp][sum]: the waiting nodes of the run under way
    __shared__ double s_tot[kIcpBlock / 32][32];                     // the strands' totals (last block of a pair)
    __shared__ unsigned int s_ticket;
    const int done = st->done;
    if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
    const int32_t s0 = A.src_off[pair], s1 = A.src_off[pair + 1];
    const int32_t t0 = A.tgt_off[pair];
    const int32_t c0 = A.chunk_off[pair], c1 = A.chunk_off[pair + 1];
    if (done) return;  // uniform over the block
    // The 29 sums of a pair are added in a FIXED order that no launch parameter can change, so a pair's result is bit-identical alone
    // and inside any batch, on any grid. The chunks are dealt to four STRANDS (chunk i of the pair belongs to strand i mod 4 = the
    // warp that takes it); a GROUP is 16 consecutive chunks; a strand's LEAF is the chain (from zero, in order) over its four chunks
    // of a group; each strand adds its leaves in a binary tree over the groups (node i of a level = node 2i + node 2i+1 of the level
    // below, an odd last node moves up alone); the pair's sum is ((T0 + T1) + T2) + T3 over the strands' roots. A launch chooses freely
    // (from the size of the whole batch: icp_prepare) how many adjacent groups form a RUN (2^run_log2: a warp builds the strand's node
    // of a run on its own, with a binary counter in shared memory, no hand-shake between warps) and how many blocks share a pair's
    // runs (block b takes runs b, b + gridDim.x, ...): a big batch runs long-lived blocks, a single pair is cut to fill the machine
    // exactly. The pair's last block adds the levels above the runs (one strand per warp) and finishes the pass.
    const int32_t n_groups = max(1, (c1 - c0 + 15) >> 4);  // an empty source has one empty group: block 0 still finishes its pass
    const int32_t n_runs = (n_groups + (1 << A.run_log2) - 1) >> A.run_log2;
    const int n_active = min((int)gridDim.x, (int)n_runs);  // blocks of this pair that have a run
    if ((int)blockIdx.x >= n_active) return;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Icp2Smem& S = reinterpret_cast<Icp2Smem*>(icp2_smem)[warp];
    stage2_init_barrier(&S.mbar);
    uint32_t parity = 0;
    const UnitFrame F = unit_frame(A.grid.lat[pair], A.grid.shift);
    const double inv_pm2 = (1.0 / F.per_m) * (1.0 / F.per_m);  // metres^2 per unit^2
    const float inv_pm = (float)(1.0 / F.per_m);                   // metres per unit
    int op_p, op_q;
    icp_sum_operands(KIND, lane, op_p, op_q);
    double* rows_t = reinterpret_cast<double*>(S.buf);
    const double2* row_p = reinterpret_cast<const double2*>(rows_t + op_p * kIcpRowStride);
    const double2* row_q = reinterpret_cast<const double2*>(rows_t + op_q * kIcpRowStride);
    double acc = 0.0;  // lane j: running total of sum j
    const float dmax_up = (float)sqrt(A.r2) * 1.000001f;  // d_max, rounded up
    const bool affine = sT[12] == 0.0 && sT[13] == 0.0 && sT[14] == 0.0 && sT[15] == 1.0;
    // The warps of a block interleave over the chunks of a group (Hilbert neighbours at the same time: the hash slots, partner gathers
    // and normals one warp fetched are in L1 for the other three; a contiguous range per warp measured 15 % slower), a block walks the
    // runs b, b + gridDim.x, ... of its pair (R adjacent groups each: consecutive groups overlap in space).
    const int R = 1 << A.run_log2;
    for (int32_t r = blockIdx.x; r < n_runs; r += (int32_t)gridDim.x) {
        int m = 0;  // leaves of this run pushed so far
        for (int gi = 0; gi < R; ++gi) {
            const int32_t g = r * R + gi;
            if (g >= n_groups) break;
            const int32_t ce = min(c1, c0 + 16 * g + 16);
            for (int32_t c = c0 + 16 * g + warp; c < ce; c += kIcpBlock / 32) {
                const int32_t si = A.chunk_start[c] + lane;
                const bool valid = si < A.chunk_start[c + 1];
                double4 sp = make_double4(0.0, 0.0, 0.0, 0.0);
                float4 kr = make_float4(0.f, 0.f, 0.f, 0.f);
                int kp = -1;
                if (valid) {
                    sp = ld_point(A.src_sorted + si);
                    if (A.keep_ref != nullptr) {
                        kr = A.keep_ref[si];
                        kp = A.keep_pos[si];
                    }
                }
                double px = 0, py = 0, pz = 0;
                int oi = 0;
                if (valid) {
                    oi = point_index(sp);  // original (batch-global) source index
                    const double x = sp.x, y = sp.y, z = sp.z;
                    // PointCloud::Transform: (T [p,1]).xyz / w
                    px = sT[0] * x + sT[1] * y + sT[2] * z + sT[3];
                    py = sT[4] * x + sT[5] * y + sT[6] * z + sT[7];
                    pz = sT[8] * x + sT[9] * y + sT[10] * z + sT[11];
                    if (!affine) {
                        const double w = icp2_perspective_w(sT, x, y, z);
                        px /= w; py /= w; pz /= w;
                    }
                }
                // ---- correspondence: sticky check, then one staged search bounded by the distance to the previous partner ---------
                double d2 = 0.0;
                double nr0 = 0.0, nr1 = 0.0, nr2 = 0.0;  // the partner's normal when the search fetched it together with the point
                bool have_nrm = false;
                int idx = 0, pos = -1;
                double4 q = make_double4(0.0, 0.0, 0.0, 0.0);  // the partner's point record
                bool need = valid;
                float reach = dmax_up;  // this lane's search radius (float32, rounded up where it matters)
                if (valid && A.keep_ref != nullptr) {
                    // Everything here is a conservative float32 bound: movement and distances rounded UP, the stored bound was rounded DOWN.
                    // movement since the last search (+ the float roundings of the stored and of the current position)
                    const float pxf = (float)px, pyf = (float)py, pzf = (float)pz;
                    const float mx = pxf - kr.x, my = pyf - kr.y, mz = pzf - kr.z;
                    const float moved = sqrtf(fmaf(mz, mz, fmaf(my, my, mx * mx))) * 1.000001f + 4.0e-7f * (fabsf(pxf) + fabsf(pyf) + fabsf(pzf));
                    const float lim = kr.w;  // every other target point was at least this far from the stored position (0: unknown)
                    if (kp >= 0) {
                        q = ld_point(A.grid.pts + kp);
                        if (KIND == B3D_ICP_POINT_TO_PLANE) prefetch_l1(A.tgt_nrm_sorted + 3 * (int64_t)kp);  // most lanes keep this partner
                        const double dk = dist2<double>(px - q.x, py - q.y, pz - q.z);
                        const float u = sqrtf((float)dk) * 1.000001f;
#ifdef B3D_ICP2_STATS_MOVES  // [1] lanes with a partner, [4] sum of their movement since the last search, [5] of the room lim - u (micrometres)
                        if (A.stats) {
                            atomicAdd(&g_icp_stats[1], 1ull);
                            atomicAdd(&g_icp_stats[4], (unsigned long long)(1.0e6f * moved));
                            atomicAdd(&g_icp_stats[5], (unsigned long long)(1.0e6f * fmaxf(lim - u, 0.f)));
                        }
#endif
                        if ((u + moved) * 1.000001f < lim) {  // still strictly nearer than anything else can be
                            need = false;
                            pos = kp;
                            d2 = dk;
                            idx = point_index(q);
                        } else {
                            const float slack = fminf(fmaxf(0.5f * moved, 0.01f * dmax_up), 0.1f * dmax_up);
                            reach = fminf(u + slack, dmax_up);  // nothing beyond d_max counts anyway
                        }
                    } else if (lim > 0.0f) {
                        if ((dmax_up + moved) * 1.000001f < lim) need = false;  // nothing was within lim, nothing can be within d_max now
                        else if (moved < 0.5f * ((float)kIcpReach2 - 1.0f) * dmax_up) reach = dmax_up * (float)kIcpReach2;
                    }
                }
                const unsigned int need_mask = __ballot_sync(0xffffffffu, need);
#ifdef B3D_ICP2_STATS
                if (A.stats && lane == 0) {
                    if (need_mask == 0u) atomicAdd(&g_icp_stats[6], 1ull);
                    atomicAdd(&g_icp_stats[7], (unsigned long long)__popc(need_mask));
                }
#endif
                if (need_mask != 0u) {
                    // the chunk's box in fixed-point units of the target grid: every searching lane's ball, one unit of margin for the
                    // floor() of the records and one for the roundings here
                    const double ux = unit_coord_of_query(px, F.ox, F.per_m), uy = unit_coord_of_query(py, F.oy, F.per_m), uz = unit_coord_of_query(pz, F.oz, F.per_m);
                    const double ru = (double)reach * F.per_m + 2.0;
                    int lox = need ? unit_floor_clamped(ux - ru) : 0x7fffffff, loy = need ? unit_floor_clamped(uy - ru) : 0x7fffffff,
                        loz = need ? unit_floor_clamped(uz - ru) : 0x7fffffff;
                    int hix = need ? unit_ceil_clamped(ux + ru) : (int)0x80000000, hiy = need ? unit_ceil_clamped(uy + ru) : (int)0x80000000,
                        hiz = need ? unit_ceil_clamped(uz + ru) : (int)0x80000000;
                    lox = __reduce_min_sync(0xffffffffu, lox); loy = __reduce_min_sync(0xffffffffu, loy); loz = __reduce_min_sync(0xffffffffu, loz);
                    hix = __reduce_max_sync(0xffffffffu, hix); hiy = __reduce_max_sync(0xffffffffu, hiy); hiz = __reduce_max_sync(0xffffffffu, hiz);
                    // how far this lane's query is from the faces of the (unclamped) box: every target point that is NOT staged lies
                    // outside the box, i.e. at least this far away -- usually well beyond the lane's own reach (metres, rounded down)
                    const float d_out = (float)(fmin(fmin(fmin(ux - (double)lox, (double)hix - ux), fmin(uy - (double)loy, (double)hiy - uy)),
                                                     fmin(uz - (double)loz, (double)hiz - uz)) - 2.0) * inv_pm * 0.999999f;
                    lox = max(lox, 0); loy = max(loy, 0); loz = max(loz, 0);
                    // this lane's query as an offset from the box centre (the centre stage2_run uses)
                    const int ccx = (int)(((long long)lox + hix) >> 1), ccy = (int)(((long long)loy + hiy) >> 1), ccz = (int)(((long long)loz + hiz) >> 1);
                    const double qdx = ux - (double)ccx, qdy = uy - (double)ccy, qdz = uz - (double)ccz;
                    const float qfx = (float)qdx, qfy = (float)qdy, qfz = (float)qdz;
                    const float fx = -2.0f * qfx, fy = -2.0f * qfy, fz = -2.0f * qfz;
                    // largest offset component of a candidate or of this query
                    const float H = fmaxf(fmaxf(fmaxf((float)(hix - ccx), (float)(hiy - ccy)), (float)(hiz - ccz)) + 1.0f, fmaxf(fmaxf(fabsf(qfx), fabsf(qfy)), fabsf(qfz)));
                    const float band_w = stage2_band(H);
                    float best = 3.0e38f, second = 3.0e38f;
                    int wpos = -1;      // sorted position of the float winner
                    int last_kept = 0;  // candidates of the last batch (still in shared memory after the call)
                    const float2 f2x = make_float2(fx, fx), f2y = make_float2(fy, fy), f2z = make_float2(fz, fz);
                    auto scan = [&](int kept) {
                        last_kept = kept;
                        if (!need) return;
                        float b = best, s2 = second;
                        int grp = -1;
                        // eight candidates (two quads of one slot group) per step along a running pointer: the group layout costs one
                        // conditional bump per step, no address arithmetic per quad; the padding covers the ragged tail
                        const float4* qp = S.buf;
#pragma unroll 1
                        for (int gi = 0; gi < kept; gi += 8) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float4 t = quad_t_at(qp + h, f2x, f2y, f2z);
                                const float m01 = fminf(t.x, t.y), M01 = fmaxf(t.x, t.y), m23 = fminf(t.z, t.w), M23 = fmaxf(t.z, t.w);
                                const float m = fminf(m01, m23), Mm = fmaxf(m01, m23);
                                const float sg = fminf(fminf(M01, M23), Mm);  // second smallest of the four
                                s2 = fminf(s2, fminf(sg, fmaxf(m, b)));
                                grp = m < b ? gi + 4 * h : grp;
                                b = fminf(b, m);
                            }
                            qp += (gi & 24) == 24 ? 26 : 2;  // next pair of quads; past the group's fourth pair, the next group
                        }
                        if (grp >= 0) {
                            // the new best sits in group grp: the first of the four that reproduces it (ties end up in the float64 path)
                            const float4 t = quad_t(S.buf, grp, f2x, f2y, f2z);
                            int w = grp + 3;
                            if (t.z == b) w = grp + 2;
                            if (t.y == b) w = grp + 1;
                            if (t.x == b) w = grp;
                            wpos = S.pos[w];
                        }
                        best = b;
                        second = s2;
                    };
                    const int nb = stage2_run<kIcp2Cap>(A.grid, F, pair, lox, loy, loz, hix, hiy, hiz, S, parity, scan);
#ifdef B3D_ICP2_STATS
                    if (A.stats && lane == 0) {
                        atomicAdd(&g_icp_stats[0], 1ull);
                        if (nb < 0) atomicAdd(&g_icp_stats[2], 1ull);
                        else atomicAdd(&g_icp_stats[3], (unsigned long long)last_kept);
#ifndef B3D_ICP2_STATS_MOVES
                        if (nb > 1) atomicAdd(&g_icp_stats[1], 1ull);
#endif
                    }
#endif
                    double others2 = 3.0e38;  // lower bound of the squared distance (metres) of every target point but the winner
                    bool bounded = nb >= 0;   // every target point within `reach` of the query was looked at
                    if (need) {
                        const bool ambiguous = nb >= 0 && wpos >= 0 && second <= best + band_w;
                        if (nb < 0 || (ambiguous && nb > 1)) {
                            // box too large to stage, or a near-tie in a box that took several batches (the earlier candidates are gone)
                            pos = icp2_walk(A.grid, pair, px, py, pz, A.r2, A.rmax);
                            if (nb >= 0) bounded = false;
#if defined(B3D_ICP2_STATS) && !defined(B3D_ICP2_STATS_MOVES)
                            if (A.stats) atomicAdd(&g_icp_stats[4], 1ull);
#endif
                        } else {
                            pos = wpos;
                        }
                        if (pos >= 0) {
                            q = ld_point(A.grid.pts + pos);
                            if (KIND == B3D_ICP_POINT_TO_PLANE) {  // the normal's latency runs in parallel with the point's
                                const double* nq = A.tgt_nrm_sorted + 3 * (int64_t)pos;
                                nr0 = __ldg(nq); nr1 = __ldg(nq + 1); nr2 = __ldg(nq + 2);
                                have_nrm = true;
                            }
                            d2 = dist2<double>(px - q.x, py - q.y, pz - q.z);
                            idx = point_index(q);
                        }
                        if (nb >= 0 && wpos >= 0 && !(ambiguous && nb > 1)) {
                            if (ambiguous) {
#if defined(B3D_ICP2_STATS) && !defined(B3D_ICP2_STATS_MOVES)
                                if (A.stats) atomicAdd(&g_icp_stats[5], 1ull);
#endif
                                const int tpos = icp2_resolve_ties(A.grid.pts, S.buf, S.pos, last_kept, fx, fy, fz, best + band_w, px, py, pz, wpos, d2, idx);
                                if (tpos != pos) {
                                    have_nrm = false;
                                    pos = tpos;
                                    q = ld_point(A.grid.pts + pos);
                                    d2 = dist2<double>(px - q.x, py - q.y, pz - q.z);
                                    idx = point_index(q);
                                }
                                others2 = d2;
                            } else {
                                others2 = fmax(d2, ((qdx * qdx + qdy * qdy + qdz * qdz) + (double)second - (double)band_w) * inv_pm2);
                            }
                        }
                    }
                    __syncwarp();
                    if (need && A.keep_ref != nullptr) {
                        // what this search proved: the nearest point (if any within reach) and that every other point is at least
                        // min(runner-up, reach) away; stored rounded down
                        float lbf = 0.f;
                        if (bounded) lbf = fminf(sqrtf((float)fmin(others2, 1.0e30)) * 0.999999f, fmaxf(d_out, reach * 0.999999f));
                        A.keep_ref[si] = make_float4((float)px, (float)py, (float)pz, lbf);
                        A.keep_pos[si] = pos;
                    }
                }
                if (pos >= 0 && !(d2 < A.r2)) pos = -1;
                double e[kIcpRow];
#pragma unroll
                for (int j = 0; j < kIcpRow; ++j) e[j] = 0.0;
                double W[9], gd[3], gp[3];  // generalized ICP / information matrix only
                bool matched = false;
                if (valid) {
                    if (A.corr != nullptr) A.corr[oi] = pos >= 0 ? idx - t0 : -1;
                    if (pos >= 0) {
                        matched = true;
                        e[7] = 1.0;
                        e[8] = d2;
                        if (KIND == kIcpInformation) {
                            gp[0] = q.x; gp[1] = q.y; gp[2] = q.z;
                        } else if (KIND == B3D_ICP_POINT_TO_POINT) {
                            e[0] = px; e[1] = py; e[2] = pz;
                            e[3] = q.x; e[4] = q.y; e[5] = q.z;
                        } else if (KIND == B3D_ICP_POINT_TO_PLANE) {
                            if (!have_nrm) {
                                const double* nq = A.tgt_nrm_sorted + 3 * (int64_t)pos;
                                nr0 = __ldg(nq); nr1 = __ldg(nq + 1); nr2 = __ldg(nq + 2);
                            }
                            const double n0 = nr0, n1 = nr1, n2 = nr2;
                            e[0] = py * n2 - pz * n1; e[1] = pz * n0 - px * n2; e[2] = px * n1 - py * n0;
                            e[3] = n0; e[4] = n1; e[5] = n2;
                            e[6] = (px - q.x) * n0 + (py - q.y) * n1 + (pz - q.z) * n2;
                        } else {
                            // generalized ICP: M = C_t + R C_s R^T, W = (M^-1)^(1/2), rows r_k = W_k (p - q), J = W [ -[p]x | I ]
                            const double* Ct = A.tgt_cov_sorted + 9 * (int64_t)pos;
                            const double* Cs = A.src_cov + 9 * (int64_t)oi;
                            double RC[9], M[9];
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) RC[3 * r + cc] = sT[4 * r] * Cs[cc] + sT[4 * r + 1] * Cs[3 + cc] + sT[4 * r + 2] * Cs[6 + cc];
#pragma unroll
                            for (int r = 0; r < 3; ++r)
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc)
                                    M[3 * r + cc] = __ldg(Ct + 3 * r + cc) + (RC[3 * r] * sT[4 * cc] + RC[3 * r + 1] * sT[4 * cc + 1] + RC[3 * r + 2] * sT[4 * cc + 2]);
                            gicp_factor(M, W);
                            gd[0] = px - q.x; gd[1] = py - q.y; gd[2] = pz - q.z;
                            gp[0] = px; gp[1] = py; gp[2] = pz;
                        }
                    }
                }
                const int n_rows = (KIND == B3D_ICP_GENERALIZED || KIND == kIcpInformation) ? 3 : 1;
                for (int row = 0; row < n_rows; ++row) {
                    if (KIND == kIcpInformation) {
                        if (matched) {
                            // G = [ -[t]x | I ] for the target point t (kept in gp)
                            e[0] = row == 0 ? 0.0 : (row == 1 ? -gp[2] : gp[1]);
                            e[1] = row == 0 ? gp[2] : (row == 1 ? 0.0 : -gp[0]);
                            e[2] = row == 0 ? -gp[1] : (row == 1 ? gp[0] : 0.0);
                            e[3] = row == 0 ? 1.0 : 0.0; e[4] = row == 1 ? 1.0 : 0.0; e[5] = row == 2 ? 1.0 : 0.0;
                            e[6] = 0.0;
                            if (row > 0) { e[7] = 0.0; e[8] = 0.0; }
                        }
                    }
                    if (KIND == B3D_ICP_GENERALIZED) {
                        if (matched) {
                            const double w0 = W[3 * row], w1 = W[3 * row + 1], w2 = W[3 * row + 2];
                            // J = W_row [ -[p]x | I ],  -[p]x = [0 pz -py; -pz 0 px; py -px 0]
                            e[0] = w1 * (-gp[2]) + w2 * gp[1];
                            e[1] = w0 * gp[2] + w2 * (-gp[0]);
                            e[2] = w0 * (-gp[1]) + w1 * gp[0];
                            e[3] = w0; e[4] = w1; e[5] = w2;
                            e[6] = w0 * gd[0] + w1 * gd[1] + w2 * gd[2];
                            if (row > 0) { e[7] = 0.0; e[8] = 0.0; }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kIcpRow; ++j) rows_t[j * kIcpRowStride + lane] = e[j];
                    __syncwarp();
                    if ((KIND == B3D_ICP_GENERALIZED || KIND == kIcpInformation) && row > 0 && lane >= 27) {
                        // count and sum d2 are taken once per correspondence (row 0)
                    } else {
                        // fused multiply-add on purpose (the file is built with -fmad=false): the order of these sums is this kernel's own (lane
                        // order inside a chunk, chunk order inside a warp), nothing compares them bit for bit with the CPU
#pragma unroll 8
                        for (int l2 = 0; l2 < 16; ++l2) {
                            const double2 a = row_p[l2], b = row_q[l2];
                            acc = fma(a.x, b.x, acc);
                            acc = fma(a.y, b.y, acc);
                        }
                    }
                    __syncwarp();
                }
            }
            // the leaf of this strand (this warp's chain over its chunks of the group, 0 if it has none) joins the run's node: a binary
            // counter -- leaf i is added to the waiting node of every level whose bit is set in i, and parks at the first clear one
            {
                double v = acc;
                acc = 0.0;
                int level = 0;
                for (int t = m; t & 1; t >>= 1, ++level) v = stk[level][warp][lane] + v;
                stk[level][warp][lane] = v;
                ++m;
            }
        }
        // the run's node (a ragged last run: the waiting nodes from the lowest level up, the higher one on the left) -> global memory
        {
            double v = 0.0;
            bool have = false;
            for (int level = 0; (m >> level) != 0; ++level)
                if ((m >> level) & 1) {
                    v = have ? stk[level][warp][lane] + v : stk[level][warp][lane];
                    have = true;
                }
            if (lane < kIcpSums) A.partial[(((int64_t)pair * A.run_stride + r) * 4 + warp) * kIcpSums + lane] = v;
        }
    }
    // ---- end of the block: the LAST block of a pair adds the pair's nodes, one strand per warp, then warp 0 finishes the pass ----
    __threadfence();  // this thread's partials are out (device scope) before the block's ticket is drawn
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&st->ticket, 1u);
    __syncthreads();
    if (s_ticket != (unsigned int)n_active - 1u) return;  // uniform over the block
    __threadfence();
    {
        const double t = icp2_reduce_strand(A, pair, n_runs, warp);
        s_tot[warp][lane] = t;
    }
    __syncthreads();
    if (warp != 0) return;
    const double total = ((s_tot[0][lane] + s_tot[1][lane]) + s_tot[2][lane]) + s_tot[3][lane];
    icp2_finish_pair(A, pair, total, s1 - s0);
}

__global__ void icp_finalize_kernel(int kind, const double* __restrict__ sums, const int32_t* __restrict__ src_off, const int64_t* __restrict__ ns_global,
                                    double rel_fitness, double rel_rmse, int max_iter, IcpPairState* state, int P) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= P) return;
    IcpPairState* st = state + pair;
    if (st->done) return;
    double a[kIcpSums];
    for (int j = 0; j < kIcpSums; ++j) a[j] = sums[(int64_t)pair * kIcpSums + j];
    const double ns = ns_global ? (double)ns_global[pair] : (double)(src_off[pair + 1] - src_off[pair]);
    icp_finalize_pair(kind, a, ns, rel_fitness, rel_rmse, max_iter, st);
}

__global__ void icp_init_state_kernel(IcpPairState* state, const double* __restrict__ init, int P) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= P) return;
    IcpPairState s;
    for (int i = 0; i < 16; ++i) s.T[i] = init ? init[16 * pair + i] : ((i % 5 == 0) ? 1.0 : 0.0);
    s.fitness = s.rmse = s.prev_fitness = s.prev_rmse = 0.0;
    s.n_corr = 0;
    s.iter = 0;
    s.done = 0;
    s.converged = 0;
    s.ticket = 0;
    s.pass_id = 0;
    state[pair] = s;
}

// dst[pos] = src[original index of sorted position pos] (cols doubles per row)
__global__ void __launch_bounds__(256) gather_by_sorted_kernel(const double4* __restrict__ pts, int32_t n, const double* __restrict__ src, int cols,
                                                               double* __restrict__ dst) {
    const int64_t total = (int64_t)n * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = t / cols;
        const int c = (int)(t - pos * cols);
        const int64_t oi = (int64_t)__double_as_longlong(pts[pos].w);
        dst[t] = src[oi * cols + c];
    }
}

__global__ void __launch_bounds__(256) transform_kernel(const double* __restrict__ T16, double* __restrict__ xyz, int64_t n, double* __restrict__ nrm,
                                                        double* __restrict__ cov) {
    __shared__ double T[16];
    if (threadIdx.x < 16) T[threadIdx.x] = T16[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const double nx = T[0] * x + T[1] * y + T[2] * z + T[3];
        const double ny = T[4] * x + T[5] * y + T[6] * z + T[7];
        const double nz = T[8] * x + T[9] * y + T[10] * z + T[11];
        const double w = T[12] * x + T[13] * y + T[14] * z + T[15];
        xyz[3 * i] = nx / w; xyz[3 * i + 1] = ny / w; xyz[3 * i + 2] = nz / w;
        if (nrm != nullptr) {
            const double a = nrm[3 * i], b = nrm[3 * i + 1], c = nrm[3 * i + 2];
            nrm[3 * i] = T[0] * a + T[1] * b + T[2] * c;
            nrm[3 * i + 1] = T[4] * a + T[5] * b + T[6] * c;
            nrm[3 * i + 2] = T[8] * a + T[9] * b + T[10] * c;
        }
        if (cov != nullptr) {
            double* C = cov + 9 * i;
            double Cin[9], RC[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) Cin[k] = C[k];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c2 = 0; c2 < 3; ++c2) RC[3 * r + c2] = T[4 * r] * Cin[c2] + T[4 * r + 1] * Cin[3 + c2] + T[4 * r + 2] * Cin[6 + c2];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c2 = 0; c2 < 3; ++c2) C[3 * r + c2] = RC[3 * r] * T[4 * c2] + RC[3 * r + 1] * T[4 * c2 + 1] + RC[3 * r + 2] * T[4 * c2 + 2];
        }
    }
}

}  // namespace

int icp_prepare(b3d_ctx* ctx, const IcpProblem& pb, const double* init_h, IcpWork* w) {
    const int P = pb.P;
    B3D_TRY(w->state.alloc(ctx, P));
    B3D_TRY(w->sums.alloc(ctx, (size_t)P * kIcpSums));
    DevBuf<double> init_d;
    if (init_h) {
        B3D_TRY(init_d.alloc(ctx, (size_t)P * 16));
        B3D_TRY(ctx->upload(init_d.p, init_h, (size_t)P * 16 * sizeof(double)));
    }
    B3D_LAUNCH(ctx, icp_init_state_kernel, (P + 127) / 128, 128, 0, w->state.p, init_h ? init_d.p : (const double*)nullptr, P);
    const Grid<double>& g = *pb.tgt_grid;
    // order the source points along a Morton curve of the target lattice and cut them into compact warp chunks
    B3D_TRY(build_query_chunks(ctx, pb.src, pb.src_off, pb.src_off_h, g.sort, reinterpret_cast<const double*>(w->state.p),
                               (int)(sizeof(IcpPairState) / sizeof(double)), &w->chunks));
    // sticky correspondences (20 bytes per source point); w = 0 marks "never searched"
    w->keep_ref.release();
    w->keep_pos.release();
    {
        static const bool sticky_off = getenv("B3D_ICP_NO_STICKY") != nullptr;
        const size_t ns = (size_t)std::max<int32_t>(pb.src_off_h[P], 1);
        if (!sticky_off) {
            B3D_TRY(w->keep_ref.alloc(ctx, ns));
            B3D_TRY(w->keep_pos.alloc(ctx, ns));
            B3D_CUDA(cudaMemsetAsync(w->keep_ref.p, 0, ns * sizeof(float4), ctx->stream));
            B3D_CUDA(cudaMemsetAsync(w->keep_pos.p, 0xff, ns * sizeof(int32_t), ctx->stream));
        }
    }
    // partial-sum groups per pair depend only on that pair's own chunk count (results do not depend on the batch)
    // How the pass kernel is cut into blocks (results do not depend on it, see the kernel): runs of R adjacent chunk groups (16 chunks
    // each), nb blocks per pair sharing the pair's runs. A block start and hand-over costs ~0.7 chunk-times per warp, consecutive
    // groups overlap in space (L1), so longer runs are better -- as long as the machine stays full:
    //  * many waves (a batch of pairs, a very large cloud): one run of 4 groups per block (16 chunks per warp; measured on the 64-pair
    //    batch at 4 / 16 / 32 chunks per warp: 19.2 / 17.0 / 17.1 ms of ICP), longer runs for one very large cloud;
    //  * about one wave or less (a single pair): R in {4, 2, 1} and nb <= the resident slots, chosen to minimise the number of
    //    group-times of the busiest block (an 8 MP pair, 1218 groups on 592 slots: R = 1, three groups per block at most).
    {
        const int64_t slots = (int64_t)ctx->sm_count * B3D_ICP2_MIN_BLOCKS;
        auto groups_of = [&](int p) { return std::max<int64_t>(1, ((int64_t)w->chunks.chunk_off_h[p + 1] - w->chunks.chunk_off_h[p] + 15) / 16); };
        auto runs_tot = [&](int lg) {
            int64_t t = 0;
            for (int p = 0; p < P; ++p) t += (groups_of(p) + (1 << lg) - 1) >> lg;
            return t;
        };
        int64_t most_groups = 1, tot_groups = 0;
        for (int p = 0; p < P; ++p) {
            most_groups = std::max(most_groups, groups_of(p));
            tot_groups += groups_of(p);
        }
        int lg = 2;
        int64_t nb = 0;  // 0: one run per block
        if (runs_tot(2) >= 4 * slots) {
            // one cloud that dominates the batch (a sharded 1e7 .. 1e8-point cloud): every pass has all its blocks, so the runs
            // grow while ~2.4 waves remain (1e7 points, ms per pass at runs of 4 / 8 / 16 / 32 groups: 1.63 / 1.45 / 1.36 / 1.35).
            // A batch of pairs stays at 4: its late passes have few active pairs left (64 pairs at 4 / 8 / 16: 17.1 / 17.1 / 17.9)
            if (2 * most_groups >= tot_groups)
                while (lg < kIcpMaxRunLog2 && 10 * runs_tot(lg + 1) >= 24 * slots) ++lg;
        } else {
            int64_t best = INT64_MAX;
            for (int cand = 2; cand >= 0; --cand) {
                const int64_t waves = (runs_tot(cand) + slots - 1) / slots;
                const int64_t cost = waves << cand;  // group-times of the busiest block
                if (cost < best) { best = cost; lg = cand; }
            }
            nb = std::max<int64_t>(1, slots / std::max(P, 1));  // the pairs of a small batch share the resident slots
        }
        if (const char* e = getenv("B3D_ICP_RUN_LOG2")) lg = std::min(std::max(atoi(e), 0), kIcpMaxRunLog2);
        const int64_t most_runs = (most_groups + (1 << lg) - 1) >> lg;
        w->run_log2 = lg;
        w->run_stride = (int)most_runs;
        w->blocks2 = (int)(nb > 0 ? std::min(nb, most_runs) : most_runs);
        if (const char* e = getenv("B3D_ICP_BLOCKS_PER_PAIR")) w->blocks2 = (int)std::min<int64_t>(std::max(atoi(e), 1), most_runs);
    }
    w->blocks = icp_groups(w->chunks.most);
    B3D_TRY(w->partial.alloc(ctx, (size_t)P * std::max((size_t)w->blocks, (size_t)w->run_stride * 4) * kIcpSums));
    const int32_t nt = (int32_t)g.sort.n;
    if (pb.kind == B3D_ICP_POINT_TO_PLANE) {
        B3D_TRY(w->tgt_nrm_sorted.alloc(ctx, (size_t)nt * 3));
        B3D_LAUNCH(ctx, gather_by_sorted_kernel, ctx->grid_for((int64_t)nt * 3, 256, 1, 8), 256, 0, g.pts.p, nt, pb.tgt_normals, 3, w->tgt_nrm_sorted.p);
    } else if (pb.kind == B3D_ICP_GENERALIZED) {
        B3D_TRY(w->tgt_cov_sorted.alloc(ctx, (size_t)nt * 9));
        B3D_LAUNCH(ctx, gather_by_sorted_kernel, ctx->grid_for((int64_t)nt * 9, 256, 1, 8), 256, 0, g.pts.p, nt, pb.tgt_cov, 9, w->tgt_cov_sorted.p);
    }
    return B3D_OK;
}

static IcpKernelArgs make_args(const IcpProblem& pb, IcpWork* w, int32_t* corr, bool fused) {
    IcpKernelArgs A{};
    A.kind = pb.kind;
    A.src_sorted = w->chunks.q;
    A.chunk_start = w->chunks.chunk_start.p;
    A.chunk_off = w->chunks.chunk_off.p;
    A.n_chunks = w->chunks.n_chunks;
    A.src_cov = pb.src_cov;
    A.src_off = pb.src_off;
    A.ns_global = pb.ns_global;
    A.grid = pb.tgt_grid->view();
    A.tgt_off = pb.tgt_off;
    A.tgt_nrm_sorted = w->tgt_nrm_sorted.p;
    A.tgt_cov_sorted = w->tgt_cov_sorted.p;
    A.r2 = pb.max_dist * pb.max_dist;
    A.rmax = pb.rmax;
    A.rel_fitness = pb.rel_fitness;
    A.rel_rmse = pb.rel_rmse;
    A.max_iter = pb.max_iter;
    A.state = w->state.p;
    A.partial = w->partial.p;
    A.sums = w->sums.p;
    A.corr = corr;
    A.fused = fused ? 1 : 0;
    A.keep_ref = w->keep_ref.p;
    A.keep_pos = w->keep_pos.p;
    A.run_log2 = w->run_log2;
    A.run_stride = w->run_stride;
    {
        static const int stats_on = getenv("B3D_ICP_STATS") ? 1 : 0;
        A.stats = stats_on;
    }
    if (fused && w->peer_world > 1) {
        A.peer_world = w->peer_world;
        A.peer_rank = w->peer_rank;
        for (int r = 0; r < w->peer_world; ++r) A.peer_buf[r] = w->peer_buf[r];
    }
    return A;
}

template <int KIND>
static int launch_pass2(b3d_ctx* ctx, const IcpKernelArgs& A, dim3 grid) {
    const size_t smem = sizeof(Icp2Smem) * (kIcpBlock / 32);
    B3D_CUDA(cudaFuncSetAttribute(icp_pass2_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per device; cheap
    B3D_LAUNCH(ctx, icp_pass2_kernel<KIND>, grid, kIcpBlock, smem, A);
    return B3D_OK;
}

int icp_pass(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w, int32_t* corr, bool fused) {
    IcpKernelArgs A = make_args(pb, w, corr, fused);
    const dim3 grid(w->blocks, pb.P);
    static const bool v1 = getenv("B3D_ICP_V1") != nullptr;  // the round-1 kernel (kept for A/B runs and as the reference of the equality test)
    if (!v1 && A.grid.rec != nullptr) {
        const dim3 grid2(w->blocks2, pb.P);
        if (pb.kind == B3D_ICP_POINT_TO_POINT) return launch_pass2<B3D_ICP_POINT_TO_POINT>(ctx, A, grid2);
        if (pb.kind == kIcpInformation) return launch_pass2<kIcpInformation>(ctx, A, grid2);
        if (pb.kind == B3D_ICP_POINT_TO_PLANE) return launch_pass2<B3D_ICP_POINT_TO_PLANE>(ctx, A, grid2);
        return launch_pass2<B3D_ICP_GENERALIZED>(ctx, A, grid2);
    }
    if (pb.kind == B3D_ICP_POINT_TO_POINT) {
        B3D_LAUNCH(ctx, icp_pass_kernel<B3D_ICP_POINT_TO_POINT>, grid, kIcpBlock, 0, A);
    } else if (pb.kind == kIcpInformation) {
        B3D_LAUNCH(ctx, icp_pass_kernel<kIcpInformation>, grid, kIcpBlock, 0, A);
    } else if (pb.kind == B3D_ICP_POINT_TO_PLANE) {
        B3D_LAUNCH(ctx, icp_pass_kernel<B3D_ICP_POINT_TO_PLANE>, grid, kIcpBlock, 0, A);
    } else {
        B3D_LAUNCH(ctx, icp_pass_kernel<B3D_ICP_GENERALIZED>, grid, kIcpBlock, 0, A);
    }
    return B3D_OK;
}

int icp_finalize_from_sums(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w) {
    B3D_LAUNCH(ctx, icp_finalize_kernel, (pb.P + 127) / 128, 128, 0, pb.kind, w->sums.p, pb.src_off, pb.ns_global, pb.rel_fitness, pb.rel_rmse,
               pb.max_iter, w->state.p, pb.P);
    return B3D_OK;
}

__global__ void icp_count_active_kernel(const IcpPairState* __restrict__ state, int P, int* __restrict__ active) {
    int n = 0;
    for (int p = threadIdx.x; p < P; p += blockDim.x) n += state[p].done ? 0 : 1;
    n = __reduce_add_sync(0xffffffffu, n);
    if (threadIdx.x == 0) *active = n;
}

int icp_run(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w, int32_t* corr) {
    // max_iter updates need max_iter + 1 correspondence passes; finished pairs return at once. Most runs converge long before
    // max_iter: from the fourth pass on the host looks (every other pass) whether any pair is still active.
    DevBuf<int> active_d;
    B3D_TRY(active_d.alloc(ctx, 1));
    for (int k = 0; k <= pb.max_iter; ++k) {
        B3D_TRY(icp_pass(ctx, pb, w, corr, true));
        if (k >= 3 && (k & 1) == 1 && k < pb.max_iter) {
            B3D_LAUNCH(ctx, icp_count_active_kernel, 1, 32, 0, w->state.p, pb.P, active_d.p);
            int active = 1;
            B3D_TRY(ctx->download(&active, active_d.p, sizeof(int)));
            if (active == 0) break;
        }
    }
    return B3D_OK;
}

int icp_results(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w, b3d_icp_result* results_h) {
    std::vector<IcpPairState> st(pb.P);
    B3D_TRY(ctx->download(st.data(), w->state.p, (size_t)pb.P * sizeof(IcpPairState)));
    for (int p = 0; p < pb.P; ++p) {
        b3d_icp_result& r = results_h[p];
        memcpy(r.transformation, st[p].T, sizeof(r.transformation));
        r.fitness = st[p].fitness;
        r.inlier_rmse = st[p].rmse;
        r.iterations = st[p].iter;
        r.converged = st[p].converged;
        r.n_correspondences = st[p].n_corr;
    }
    return B3D_OK;
}

}  // namespace b3d

using namespace b3d;

struct b3d_icp_state {
    IcpProblem pb;
    IcpWork work;
    Grid<double> grid;
    DevBuf<int32_t> src_off, tgt_off;
    Segments tgt_seg;
    DevBuf<int32_t> corr;
    bool pending_sums = false;
};

static int validate_icp(int kind, const double* src, int64_t ns, const double* src_cov, const double* tgt, int64_t nt, const double* tgt_normals,
                        const double* tgt_cov, double max_dist, int max_iter) {
    B3D_REQUIRE(kind >= 0 && kind <= 2, "unknown ICP kind %d", kind);
    B3D_REQUIRE(max_dist > 0.0, "Invalid max_correspondence_distance.");
    B3D_REQUIRE(ns >= 0 && nt >= 0 && max_iter >= 0, "negative size");
    B3D_REQUIRE(ns == 0 || src != nullptr, "source is NULL");
    B3D_REQUIRE(nt == 0 || tgt != nullptr, "target is NULL");
    if (kind == B3D_ICP_POINT_TO_PLANE)
        B3D_REQUIRE(nt == 0 || tgt_normals != nullptr,
                    "TransformationEstimationPointToPlane and TransformationEstimationColoredICP require pre-computed normal vectors for target "
                    "PointCloud.");
    if (kind == B3D_ICP_GENERALIZED) B3D_REQUIRE((ns == 0 || src_cov) && (nt == 0 || tgt_cov), "GeneralizedICP requires covariances.");
    return B3D_OK;
}

static void empty_result(const double* init_h, b3d_icp_result* r) {
    for (int i = 0; i < 16; ++i) r->transformation[i] = init_h ? init_h[i] : ((i % 5 == 0) ? 1.0 : 0.0);
    r->fitness = 0;
    r->inlier_rmse = 0;
    r->iterations = 0;
    r->converged = 0;
    r->n_correspondences = 0;
}

// Target grid of a stand-alone registration: cells somewhat larger than d_max (fewer hash probes per staged box, see the
// sweep in b3d_pipeline.cu); rmax = rings a per-lane walk needs for d_max on that grid.
static int build_icp_grid(b3d_ctx* ctx, const double* tgt, const Segments& seg, double max_dist, Grid<double>* grid, int* rmax) {
    static const double scale = getenv("B3D_ICP_CELL_SCALE") ? atof(getenv("B3D_ICP_CELL_SCALE")) : 1.3;
    B3D_TRY(build_search_grid<double>(ctx, tgt, seg, 8, max_dist * scale, grid, nullptr));
    *rmax = rings_for_radius(max_dist, grid->cell);
    return B3D_OK;
}

// builds the single-pair problem (target grid included) inside st
static int setup_single(b3d_ctx* ctx, b3d_icp_state* st, int kind, const double* src, int64_t ns_local, const double* src_cov, const double* tgt,
                        int64_t nt, const double* tgt_normals, const double* tgt_cov, double max_dist, const double* init_h, double rel_fitness,
                        double rel_rmse, int max_iter) {
    B3D_TRY(single_segment(ctx, nt, &st->tgt_off, &st->tgt_seg));
    Segments sseg;
    B3D_TRY(single_segment(ctx, ns_local, &st->src_off, &sseg));
    int rmax = 1;
    B3D_TRY(build_icp_grid(ctx, tgt, st->tgt_seg, max_dist, &st->grid, &rmax));
    IcpProblem& pb = st->pb;
    pb.kind = kind;
    pb.P = 1;
    pb.src = src;
    pb.src_cov = src_cov;
    pb.src_off = st->src_off.p;
    pb.src_off_h = sseg.off_h;
    pb.tgt_grid = &st->grid;
    pb.tgt_off = st->tgt_off.p;
    pb.tgt_normals = tgt_normals;
    pb.tgt_cov = tgt_cov;
    pb.max_dist = max_dist;
    pb.rmax = rmax;
    pb.rel_fitness = rel_fitness;
    pb.rel_rmse = rel_rmse;
    pb.max_iter = max_iter;
    return icp_prepare(ctx, pb, init_h, &st->work);
}

extern "C" {

// [0] chunks that staged a search, [1] unused, [2] staging overflows (per-lane fallback), [3] staged candidates, [4] sum of box
// volumes (cm^3), [5] sum of longest box edges (0.1 mm), [6] chunks whose lanes all kept their partner without a search,
// [7] lanes that searched. Debug aid (B3D_ICP_STATS=1), not part of the public header.
int b3d_debug_icp_stats(unsigned long long* out8, int reset) {
    if (out8 && cudaMemcpyFromSymbol(out8, g_icp_stats, 8 * sizeof(unsigned long long)) != cudaSuccess) return B3D_E_CUDA;
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(g_icp_stats, z, sizeof(z)) != cudaSuccess) return B3D_E_CUDA;
    }
    return B3D_OK;
}

int b3d_transform_f64(b3d_ctx* ctx, const double* T_h, double* xyz, int64_t n, double* normals, double* cov) {
    B3D_REQUIRE(ctx != nullptr && T_h != nullptr, "b3d_transform_f64: NULL argument");
    B3D_REQUIRE(n >= 0, "negative point count");
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(xyz != nullptr, "b3d_transform_f64: NULL buffer");
    B3D_TRY(ctx->bind());
    DevBuf<double> T;
    B3D_TRY(T.alloc(ctx, 16));
    B3D_TRY(ctx->upload(T.p, T_h, 16 * sizeof(double)));
    B3D_LAUNCH(ctx, transform_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, T.p, xyz, n, normals, cov);
    return B3D_OK;
}

int b3d_icp_correspondences(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const double* T_h, double max_dist,
                            int32_t* corr, double* stats_h) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_TRY(validate_icp(B3D_ICP_POINT_TO_POINT, src, ns, nullptr, tgt, nt, nullptr, nullptr, max_dist, 0));
    if (stats_h) stats_h[0] = stats_h[1] = 0.0;
    if (ns == 0) return B3D_OK;
    B3D_TRY(ctx->bind());
    if (nt == 0) {
        if (corr) B3D_CUDA(cudaMemsetAsync(corr, 0xff, (size_t)ns * sizeof(int32_t), ctx->stream));
        return B3D_OK;
    }
    b3d_icp_state st;
    B3D_TRY(setup_single(ctx, &st, B3D_ICP_POINT_TO_POINT, src, ns, nullptr, tgt, nt, nullptr, nullptr, max_dist, T_h, 1e-6, 1e-6, 0));
    B3D_TRY(icp_pass(ctx, st.pb, &st.work, corr, false));
    double sums[kIcpSums];
    B3D_TRY(ctx->download(sums, st.work.sums.p, sizeof(sums)));
    if (stats_h) {
        stats_h[0] = sums[27];
        stats_h[1] = sums[28];
    }
    return B3D_OK;
}

int b3d_information_matrix(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const double* T_h, double max_dist,
                           double* info_h) {
    B3D_REQUIRE(ctx != nullptr && info_h != nullptr, "b3d_information_matrix: NULL argument");
    B3D_TRY(validate_icp(B3D_ICP_POINT_TO_POINT, src, ns, nullptr, tgt, nt, nullptr, nullptr, max_dist, 0));
    for (int i = 0; i < 36; ++i) info_h[i] = 0.0;
    if (ns == 0 || nt == 0) return B3D_OK;
    B3D_TRY(ctx->bind());
    b3d_icp_state st;
    B3D_TRY(setup_single(ctx, &st, B3D_ICP_POINT_TO_POINT, src, ns, nullptr, tgt, nt, nullptr, nullptr, max_dist, T_h, 1e-6, 1e-6, 0));
    st.pb.kind = kIcpInformation;
    B3D_TRY(icp_pass(ctx, st.pb, &st.work, nullptr, false));
    double sums[kIcpSums];
    B3D_TRY(ctx->download(sums, st.work.sums.p, sizeof(sums)));
    int t = 0;
    for (int u = 0; u < 6; ++u)
        for (int v = u; v < 6; ++v) {
            info_h[6 * u + v] = sums[t];
            info_h[6 * v + u] = sums[t];
            ++t;
        }
    return B3D_OK;
}

int b3d_icp(b3d_ctx* ctx, int kind, const double* src, int64_t ns, const double* src_cov, const double* tgt, int64_t nt, const double* tgt_normals,
            const double* tgt_cov, double max_dist, const double* init_h, double rel_fitness, double rel_rmse, int max_iter,
            b3d_icp_result* result_h, int32_t* corr) {
    B3D_REQUIRE(ctx != nullptr && result_h != nullptr, "b3d_icp: NULL argument");
    B3D_TRY(validate_icp(kind, src, ns, src_cov, tgt, nt, tgt_normals, tgt_cov, max_dist, max_iter));
    empty_result(init_h, result_h);
    if (ns == 0) return B3D_OK;
    B3D_TRY(ctx->bind());
    if (nt == 0) {
        if (corr) B3D_CUDA(cudaMemsetAsync(corr, 0xff, (size_t)ns * sizeof(int32_t), ctx->stream));
        return ctx->sync();
    }
    b3d_icp_state st;
    B3D_TRY(setup_single(ctx, &st, kind, src, ns, src_cov, tgt, nt, tgt_normals, tgt_cov, max_dist, init_h, rel_fitness, rel_rmse, max_iter));
    B3D_TRY(icp_run(ctx, st.pb, &st.work, corr));
    return icp_results(ctx, st.pb, &st.work, result_h);
}

int b3d_icp_batch(b3d_ctx* ctx, int kind, int n_pairs, const double* src, const int64_t* src_off_h, const double* src_cov, const double* tgt,
                  const int64_t* tgt_off_h, const double* tgt_normals, const double* tgt_cov, double max_dist, const double* init_h,
                  double rel_fitness, double rel_rmse, int max_iter, b3d_icp_result* results_h, int32_t* corr) {
    B3D_REQUIRE(ctx != nullptr && results_h != nullptr && src_off_h != nullptr && tgt_off_h != nullptr, "b3d_icp_batch: NULL argument");
    B3D_REQUIRE(n_pairs >= 1, "b3d_icp_batch: n_pairs must be >= 1");
    const int P = n_pairs;
    B3D_REQUIRE(src_off_h[0] == 0 && tgt_off_h[0] == 0, "b3d_icp_batch: offsets must start at 0");
    for (int p = 0; p < P; ++p)
        B3D_REQUIRE(src_off_h[p + 1] >= src_off_h[p] && tgt_off_h[p + 1] >= tgt_off_h[p], "b3d_icp_batch: offsets must be non-decreasing");
    const int64_t ns = src_off_h[P], nt = tgt_off_h[P];
    B3D_REQUIRE(ns < (int64_t)INT32_MAX && nt < (int64_t)INT32_MAX, "b3d_icp_batch: more than 2^31-1 points");
    B3D_TRY(validate_icp(kind, src, ns, src_cov, tgt, nt, tgt_normals, tgt_cov, max_dist, max_iter));
    for (int p = 0; p < P; ++p) empty_result(init_h ? init_h + 16 * p : nullptr, &results_h[p]);
    if (ns == 0) return B3D_OK;
    B3D_TRY(ctx->bind());
    if (corr) B3D_CUDA(cudaMemsetAsync(corr, 0xff, (size_t)ns * sizeof(int32_t), ctx->stream));
    if (nt == 0) return ctx->sync();
    b3d_icp_state st;
    std::vector<int32_t> so(P + 1), to(P + 1);
    for (int p = 0; p <= P; ++p) {
        so[p] = (int32_t)src_off_h[p];
        to[p] = (int32_t)tgt_off_h[p];
    }
    Segments sseg;
    B3D_TRY(upload_segments(ctx, to, &st.tgt_off, &st.tgt_seg));
    B3D_TRY(upload_segments(ctx, so, &st.src_off, &sseg));
    int rmax = 1;
    B3D_TRY(build_icp_grid(ctx, tgt, st.tgt_seg, max_dist, &st.grid, &rmax));
    IcpProblem& pb = st.pb;
    pb.kind = kind;
    pb.P = P;
    pb.src = src;
    pb.src_cov = src_cov;
    pb.src_off = st.src_off.p;
    pb.src_off_h = so;
    pb.tgt_grid = &st.grid;
    pb.tgt_off = st.tgt_off.p;
    pb.tgt_normals = tgt_normals;
    pb.tgt_cov = tgt_cov;
    pb.max_dist = max_dist;
    pb.rmax = rmax;
    pb.rel_fitness = rel_fitness;
    pb.rel_rmse = rel_rmse;
    pb.max_iter = max_iter;
    B3D_TRY(icp_prepare(ctx, pb, init_h, &st.work));
    B3D_TRY(icp_run(ctx, pb, &st.work, corr));
    return icp_results(ctx, pb, &st.work, results_h);
}

int b3d_icp_begin(b3d_ctx* ctx, int kind, const double* src, int64_t ns_local, int64_t ns_total, const double* src_cov, const double* tgt,
                  int64_t nt, const double* tgt_normals, const double* tgt_cov, double max_dist, const double* init_h, double rel_fitness,
                  double rel_rmse, int max_iter, b3d_icp_state** out) {
    B3D_REQUIRE(ctx != nullptr && out != nullptr, "b3d_icp_begin: NULL argument");
    *out = nullptr;
    B3D_TRY(validate_icp(kind, src, ns_local, src_cov, tgt, nt, tgt_normals, tgt_cov, max_dist, max_iter));
    B3D_REQUIRE(nt > 0, "b3d_icp_begin: empty target");
    B3D_REQUIRE(ns_total >= ns_local, "b3d_icp_begin: ns_total < ns_local");
    B3D_TRY(ctx->bind());
    b3d_icp_state* st = new b3d_icp_state();
    int rc = setup_single(ctx, st, kind, src, ns_local, src_cov, tgt, nt, tgt_normals, tgt_cov, max_dist, init_h, rel_fitness, rel_rmse, max_iter);
    if (rc == B3D_OK) rc = st->work.ns_global.alloc(ctx, 1);
    if (rc == B3D_OK) rc = ctx->upload(st->work.ns_global.p, &ns_total, sizeof(int64_t));
    if (rc == B3D_OK) rc = st->corr.alloc(ctx, (size_t)std::max<int64_t>(ns_local, 1));
    if (rc != B3D_OK) {
        delete st;
        return rc;
    }
    st->pb.ns_global = st->work.ns_global.p;
    *out = st;
    return B3D_OK;
}

int b3d_icp_accumulate(b3d_ctx* ctx, b3d_icp_state* st, double** sums_dev_out) {
    B3D_REQUIRE(ctx != nullptr && st != nullptr && sums_dev_out != nullptr, "b3d_icp_accumulate: NULL argument");
    B3D_TRY(ctx->bind());
    if (st->pb.src_off_h[1] > 0) {
        B3D_TRY(icp_pass(ctx, st->pb, &st->work, st->corr.p, false));
    } else {
        B3D_CUDA(cudaMemsetAsync(st->work.sums.p, 0, kIcpSums * sizeof(double), ctx->stream));
    }
    st->pending_sums = true;
    *sums_dev_out = st->work.sums.p;
    return B3D_OK;
}

int b3d_icp_set_peers(b3d_ctx* ctx, b3d_icp_state* st, int rank, int world, void* const* peer_bufs_h) {
    B3D_REQUIRE(ctx != nullptr && st != nullptr && peer_bufs_h != nullptr, "b3d_icp_set_peers: NULL argument");
    B3D_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "b3d_icp_set_peers: rank %d / world %d out of range (world <= 8)", rank, world);
    st->work.peer_world = world;
    st->work.peer_rank = rank;
    for (int r = 0; r < world; ++r) {
        B3D_REQUIRE(peer_bufs_h[r] != nullptr, "b3d_icp_set_peers: peer buffer %d is NULL", r);
        st->work.peer_buf[r] = static_cast<double*>(peer_bufs_h[r]);
    }
    return B3D_OK;
}

int b3d_icp_pass_peers(b3d_ctx* ctx, b3d_icp_state* st, int* done_h) {
    B3D_REQUIRE(ctx != nullptr && st != nullptr, "b3d_icp_pass_peers: NULL argument");
    if (st->work.peer_world < 1) return set_error(B3D_E_STATE, "b3d_icp_pass_peers called before b3d_icp_set_peers");
    B3D_REQUIRE(st->pb.src_off_h[1] > 0, "b3d_icp_pass_peers: every rank needs a non-empty shard of the source");
    B3D_TRY(ctx->bind());
    B3D_TRY(icp_pass(ctx, st->pb, &st->work, st->corr.p, true));
    if (done_h) {
        IcpPairState s;
        B3D_TRY(ctx->download(&s, st->work.state.p, sizeof(s)));
        if (s.converged < 0) return set_error(B3D_E_STATE, "peer exchange timed out: a rank did not arrive within ~2 s");
        *done_h = s.done;
    }
    return B3D_OK;
}

int b3d_icp_update(b3d_ctx* ctx, b3d_icp_state* st, int* done_h) {
    B3D_REQUIRE(ctx != nullptr && st != nullptr, "b3d_icp_update: NULL argument");
    if (!st->pending_sums) return set_error(B3D_E_STATE, "b3d_icp_update called without a preceding b3d_icp_accumulate");
    B3D_TRY(ctx->bind());
    B3D_TRY(icp_finalize_from_sums(ctx, st->pb, &st->work));
    st->pending_sums = false;
    if (done_h) {
        IcpPairState s;
        B3D_TRY(ctx->download(&s, st->work.state.p, sizeof(s)));
        *done_h = s.done;
    }
    return B3D_OK;
}

int b3d_icp_finish(b3d_ctx* ctx, b3d_icp_state* st, b3d_icp_result* result_h, int32_t* corr) {
    B3D_REQUIRE(ctx != nullptr && st != nullptr, "b3d_icp_finish: NULL argument");
    B3D_TRY(ctx->bind());
    int rc = B3D_OK;
    if (corr && st->pb.src_off_h[1] > 0)
        rc = cudaMemcpyAsync(corr, st->corr.p, (size_t)st->pb.src_off_h[1] * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream) == cudaSuccess
                 ? B3D_OK
                 : set_error(B3D_E_CUDA, "copy of the correspondence set failed");
    if (rc == B3D_OK && result_h) rc = icp_results(ctx, st->pb, &st->work, result_h);
    if (rc == B3D_OK) rc = ctx->sync();
    delete st;
    return rc;
}

}  // extern "C"
