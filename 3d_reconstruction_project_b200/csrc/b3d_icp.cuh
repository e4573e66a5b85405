// b3d_icp.cuh -- internal interface of the batched ICP engine (b3d_icp.cu), shared with the frame-pair pipeline.
#pragma once

#include "b3d_common.cuh"

namespace b3d {

constexpr int kIcpSums = 29;  // 21 JtJ upper + 6 Jtr (or 15 moment sums for point-to-point) + |C| + sum d2

struct IcpPairState {
    double T[16];  // current transform, row-major
    double fitness, rmse;
    double prev_fitness, prev_rmse;
    long long n_corr;
    int iter;       // updates applied so far
    int done;       // the loop has finished
    int converged;  // stopped by the relative criteria (-1: the peer exchange timed out)
    unsigned int ticket;
    unsigned int pass_id;  // passes executed (stamps of the peer exchange)
};

struct IcpProblem {
    int kind = B3D_ICP_POINT_TO_PLANE;
    int P = 1;                          // pairs
    const double* src = nullptr;        // [ns_total, 3]
    const double* src_cov = nullptr;    // [ns_total, 9] (GICP)
    const int32_t* src_off = nullptr;   // device [P+1]
    std::vector<int32_t> src_off_h;
    const int64_t* ns_global = nullptr; // optional device [P]: global source sizes (sharded clouds); else local sizes
    const Grid<double>* tgt_grid = nullptr;
    const int32_t* tgt_off = nullptr;   // device [P+1]
    const double* tgt_normals = nullptr;  // by original target index (P2L)
    const double* tgt_cov = nullptr;      // by original target index (GICP)
    double max_dist = 0;
    int rmax = 1;
    double rel_fitness = 1e-6, rel_rmse = 1e-6;
    int max_iter = 30;
};

struct IcpWork {
    DevBuf<IcpPairState> state;  // [P]
    DevBuf<double> partial;      // [P][blocks][kIcpSums]
    DevBuf<double> sums;         // [P][kIcpSums]
    DevBuf<double> tgt_nrm_sorted, tgt_cov_sorted;
    QueryChunks chunks;  // sources in Morton order of the target lattice, cut into compact warp chunks
    DevBuf<int64_t> ns_global;
    DevBuf<float4> keep_ref;   // sticky correspondences of the pass kernel (per source point, sorted order)
    DevBuf<int32_t> keep_pos;
    int blocks = 1;            // round-1 kernel: partial-sum groups per pair
    int blocks2 = 1;           // round-2 kernel: blocks per pair (grid.x)
    int run_log2 = 0;          // round-2 kernel: a run = 2^run_log2 adjacent chunk groups
    int run_stride = 1;        // round-2 kernel: run nodes reserved per pair and strand in `partial`
    int peer_world = 0, peer_rank = 0;  // exchange over peer memory (b3d_icp_set_peers)
    double* peer_buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

int icp_prepare(b3d_ctx* ctx, const IcpProblem& pb, const double* init_h /*[P][16] or NULL*/, IcpWork* w);
// one pass: correspondences at the current transforms + normal-equation sums; fused = also finalize/solve/update
int icp_pass(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w, int32_t* corr, bool fused);
int icp_finalize_from_sums(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w);
int icp_run(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w, int32_t* corr);
int icp_results(b3d_ctx* ctx, const IcpProblem& pb, IcpWork* w, b3d_icp_result* results_h);

}  // namespace b3d
