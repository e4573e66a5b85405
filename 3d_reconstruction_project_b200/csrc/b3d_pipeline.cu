// b3d_pipeline.cu -- the whole front end + registration path for a BATCH of depth frame pairs, device resident:
//   deproject (rs.pointcloud semantics, pointcloud_capture.py:35-38) -> tensor voxel_down_sample (:50) -> to_legacy (f64, :53)
//   -> legacy hybrid normals on the targets (pointcloud_alignment.py:27-28) [+ sources and GICP covariances for kind 2]
//   -> registration_icp / registration_generalized_icp (pointcloud_alignment.py:35-39, test/mini1.py:293-296, test/GICP1.py:99-102)
// All 2P frames of the batch share every launch: one deprojection, one radix sort (cloud id in the key's top bits),
// one hash grid, one normals launch, and P-wide ICP passes (grid.y = pair). Host synchronisations per batch: a handful
// (bounds -> lattice, run count -> sizes, final results), independent of P.
#include "b3d_common.cuh"
#include <cstdlib>
#include "b3d_icp.cuh"
#include "b3d_search.cuh"

#include <algorithm>

namespace b3d {

int deproject_z16_batch(b3d_ctx* ctx, const uint16_t* depth, const uint8_t* bgr, int w, int h, int frames, float fx, float fy, float ppx,
                        float ppy, float scale, float* xyz, float* rgb);
int reproject_disparity_valid_batch(b3d_ctx* ctx, const int16_t* disp, int w, int h, int frames, const double* Q_h, int min_disp16, float* xyz,
                                    std::vector<int32_t>* off_h);
template <typename T, typename IndexT>
int voxel_downsample_batch(b3d_ctx* ctx, const T* xyz, const T* a0, const T* a1, const Segments& seg, double voxel, int flavour, T* o_xyz,
                           T* o_a0, T* o_a1, IndexT* o_index, int32_t* o_count, SpatialSort* out_sort, const std::vector<double>* bounds_in,
                           double* o_xyz64);
template <typename T, bool TENSOR>
int estimate_normals_batch(b3d_ctx* ctx, const T* xyz, const Segments& seg, int max_nn, double radius, const T* prior, T* normals,
                           Grid<T>* reuse_grid, int reuse_rmax);


}  // namespace b3d

namespace b3d {

struct BackParams {
    float voxel_size;
    int normals_max_nn;
    double normals_radius;
    int icp_kind;
    double icp_max_dist, icp_rel_fitness, icp_rel_rmse;
    int icp_max_iter;
};

// Everything after the voxel down-sampling: vox64 holds the 2P down-sampled clouds back to back (P sources, then P targets), the
// float32 voxel means widened to float64 by the reduction itself (to_legacy(): exact), voff their offsets ([2P + 1]); raw_off the
// offsets of the raw clouds (for the result records).
static int register_voxels_f32(b3d_ctx* ctx, DevBuf<double>& vox64, const std::vector<int32_t>& voff, const std::vector<int32_t>& raw_off, int P,
                               const BackParams* pr, b3d_pair_result* results_h) {
    const int F = 2 * P;
    const int64_t M = voff[F];
    const int32_t Ms = voff[P], Mt = (int32_t)M - voff[P];
    const double* src_pts = vox64.p;
    const double* tgt_pts = vox64.p + 3 * (int64_t)voff[P];
    std::vector<int32_t> soff(P + 1), toff(P + 1);
    for (int p = 0; p <= P; ++p) {
        soff[p] = voff[p];
        toff[p] = voff[P + p] - voff[P];
    }
    DevBuf<int32_t> soff_d, toff_d;
    Segments sseg, tseg;
    B3D_TRY(upload_segments(ctx, soff, &soff_d, &sseg));
    B3D_TRY(upload_segments(ctx, toff, &toff_d, &tseg));

    Grid<double> tgrid, icp_grid_own;
    const Grid<double>* icp_grid = &tgrid;
    int n_rmax = 1, icp_rmax = 1;
    const double r_n = pr->normals_radius, r_i = pr->icp_max_dist;
    static const double cell_scale = getenv("B3D_PIPE_CELL_SCALE") ? atof(getenv("B3D_PIPE_CELL_SCALE")) : 1.3;
    DevBuf<double> nrm_all, cov_all;  // normals (and GICP covariances) of all 2P clouds, sources first like vox64
    const double* tnrm = nullptr;
    const double *scov = nullptr, *tcov = nullptr;
    B3D_TRY(nrm_all.alloc(ctx, (size_t)(3 * M)));
    if (pr->icp_kind == B3D_ICP_GENERALIZED) {
        // 4'. generalized ICP needs normals (-> covariances, eps = 1e-3) on BOTH sides: all 2P clouds go through ONE normals call on a
        // grid of their own (cell from the normals' radius), the targets get a second grid for the correspondence search. One 8 MP
        // pair (1.1 M voxels): 0.99 ms for both clouds together against 0.65 + 0.60 ms one after the other (a single cloud does not
        // fill the machine), and the normals' staged boxes are a third of what they are on the wider ICP cells.
        std::vector<int32_t> aoff(voff.begin(), voff.end());
        DevBuf<int32_t> aoff_d;
        Segments aseg;
        B3D_TRY(upload_segments(ctx, aoff, &aoff_d, &aseg));
        B3D_TRY((estimate_normals_batch<double, false>(ctx, vox64.p, aseg, pr->normals_max_nn, pr->normals_radius, nullptr, nrm_all.p, nullptr, 0)));
        B3D_TRY(cov_all.alloc(ctx, (size_t)(9 * M)));
        B3D_TRY(b3d_covariances_from_normals(ctx, nrm_all.p, M, 1e-3, cov_all.p));
        scov = cov_all.p;
        tcov = cov_all.p + 9 * (int64_t)voff[P];
        tnrm = nrm_all.p + 3 * (int64_t)voff[P];
        const double cell = r_i * 1.001 * cell_scale;
        B3D_TRY(grid_build<double>(ctx, tgt_pts, tseg, cell, nullptr, &tgrid));
        icp_rmax = rings_for_radius(r_i, cell);
    } else {
        // 4. ONE target grid for the normals and for the ICP correspondence search whenever their radii are comparable: cell =
        // 1.3 x the larger radius (the staged searches take any cell size). One sort instead of two. Measured on config 2
        // (ms per 64-pair step at cell = 0.75 / 1 / 1.2 / 1.35 / 1.5 / 2 radii: 59.8 / 50.8 / 48.1 / 48.0 / 50.0 / 58.7): larger
        // cells mean fewer hash probes per staged box, until the boxes of the normals overflow their staging buffer.
        const bool shared = r_n > 0 && std::max(r_n, r_i) <= 3.0 * std::min(r_n, r_i);
        if (shared) {
            const double cell = std::max(r_n, r_i) * 1.001 * cell_scale;
            B3D_TRY(grid_build<double>(ctx, tgt_pts, tseg, cell, nullptr, &tgrid));
            n_rmax = rings_for_radius(r_n, cell);
            icp_rmax = rings_for_radius(r_i, cell);
        } else {
            B3D_TRY(build_search_grid<double>(ctx, tgt_pts, tseg, pr->normals_max_nn, r_n, &tgrid, &n_rmax));
            B3D_TRY(build_search_grid<double>(ctx, tgt_pts, tseg, 8, r_i, &icp_grid_own, &icp_rmax));
            icp_grid = &icp_grid_own;
        }
        double* tn = nrm_all.p + 3 * (int64_t)voff[P];
        B3D_TRY((estimate_normals_batch<double, false>(ctx, tgt_pts, tseg, pr->normals_max_nn, pr->normals_radius, nullptr, tn, &tgrid, n_rmax)));
        tnrm = tn;
    }

    // 6. batched ICP
    IcpProblem pb;
    pb.kind = pr->icp_kind;
    pb.P = P;
    pb.src = src_pts;
    pb.src_cov = scov;
    pb.src_off = soff_d.p;
    pb.src_off_h = soff;
    pb.tgt_grid = icp_grid;
    pb.tgt_off = toff_d.p;
    pb.tgt_normals = tnrm;
    pb.tgt_cov = tcov;
    pb.max_dist = pr->icp_max_dist;
    pb.rmax = icp_rmax;
    pb.rel_fitness = pr->icp_rel_fitness;
    pb.rel_rmse = pr->icp_rel_rmse;
    pb.max_iter = pr->icp_max_iter;
    IcpWork work;
    B3D_TRY(icp_prepare(ctx, pb, nullptr, &work));
    B3D_TRY(icp_run(ctx, pb, &work, nullptr));
    std::vector<b3d_icp_result> res(P);
    B3D_TRY(icp_results(ctx, pb, &work, res.data()));
    for (int p = 0; p < P; ++p) {
        results_h[p].icp = res[p];
        results_h[p].n_raw = (int64_t)(raw_off[p + 1] - raw_off[p]) + (int64_t)(raw_off[P + p + 1] - raw_off[P + p]);
        results_h[p].m_source = soff[p + 1] - soff[p];
        results_h[p].m_target = toff[p + 1] - toff[p];
    }
    return B3D_OK;
}

// tensor voxel down-sampling of the frames [f0, f1) of xyz (raw offsets raw_off) into vox at row *m_done; appends their offsets
static int voxel_group(b3d_ctx* ctx, const float* xyz, const std::vector<int32_t>& raw_off, int f0, int f1, float voxel_size, DevBuf<double>& vox,
                       std::vector<int32_t>* voff, int64_t* m_done) {
    std::vector<int32_t> goff(f1 - f0 + 1);
    for (int f = f0; f <= f1; ++f) goff[f - f0] = raw_off[f] - raw_off[f0];
    DevBuf<int32_t> goff_d;
    Segments gseg;
    B3D_TRY(upload_segments(ctx, goff, &goff_d, &gseg));
    SpatialSort vs;
    B3D_TRY((voxel_downsample_batch<float, int64_t>(ctx, xyz + 3 * (int64_t)raw_off[f0], nullptr, nullptr, gseg, (double)voxel_size, kLatTensorVoxel,
                                                     nullptr, nullptr, nullptr, nullptr, nullptr, &vs, nullptr, vox.p + 3 * *m_done)));
    for (int f = f0; f < f1; ++f) (*voff)[f + 1] = (int32_t)(*m_done + vs.run_off_h[f - f0 + 1]);
    *m_done += vs.n_runs;
    return B3D_OK;
}

// Everything after the deprojection: xyz holds 2P float32 clouds back to back (P sources, then P targets), raw_off their offsets.
static int register_clouds_f32(b3d_ctx* ctx, DevBuf<float>& xyz, const std::vector<int32_t>& raw_off, int P, const BackParams* pr,
                               b3d_pair_result* results_h) {
    const int F = 2 * P;
    DevBuf<double> vox;
    B3D_TRY(vox.alloc(ctx, (size_t)(3 * (int64_t)raw_off[F])));
    std::vector<int32_t> voff(F + 1, 0);
    int64_t m_done = 0;
    B3D_TRY(voxel_group(ctx, xyz.p, raw_off, 0, F, pr->voxel_size, vox, &voff, &m_done));  // all 2P frames in one sort
    xyz.release();
    return register_voxels_f32(ctx, vox, voff, raw_off, P, pr, results_h);
}

}  // namespace b3d

using namespace b3d;

extern "C" {

int b3d_register_depth_pairs(b3d_ctx* ctx, const b3d_pair_params* pr, const uint16_t* depth_src, const uint16_t* depth_tgt, int n_pairs,
                             int device_inputs, b3d_pair_result* results_h) {
    B3D_REQUIRE(ctx != nullptr && pr != nullptr && results_h != nullptr, "b3d_register_depth_pairs: NULL argument");
    B3D_REQUIRE(n_pairs >= 1, "b3d_register_depth_pairs: n_pairs must be >= 1");
    B3D_REQUIRE(pr->w > 0 && pr->h > 0, "b3d_register_depth_pairs: empty image");
    B3D_REQUIRE(depth_src && depth_tgt, "b3d_register_depth_pairs: NULL depth buffer");
    B3D_REQUIRE(pr->voxel_size > 0.0f, "voxel_size must be positive.");
    B3D_REQUIRE(pr->icp_max_dist > 0.0, "Invalid max_correspondence_distance.");
    B3D_REQUIRE(pr->icp_kind >= 0 && pr->icp_kind <= 2, "unknown ICP kind %d", pr->icp_kind);
    B3D_REQUIRE(pr->normals_max_nn >= 1 && pr->normals_max_nn <= 64, "normals_max_nn must be in [1, 64]");
    B3D_REQUIRE(pr->icp_max_iter >= 0, "negative icp_max_iter");
    B3D_REQUIRE(pr->fx != 0.0f && pr->fy != 0.0f, "zero focal length");
    B3D_TRY(ctx->bind());
    const int P = n_pairs, F = 2 * P;
    const int64_t N = (int64_t)pr->w * pr->h;
    B3D_REQUIRE(N * F < (int64_t)INT32_MAX, "batch of %d frames x %lld pixels exceeds 2^31-1 points", F, (long long)N);

    std::vector<int32_t> raw_off(F + 1);
    for (int f = 0; f <= F; ++f) raw_off[f] = (int32_t)(f * N);
    BackParams bp{pr->voxel_size, pr->normals_max_nn, pr->normals_radius, pr->icp_kind, pr->icp_max_dist, pr->icp_rel_fitness, pr->icp_rel_rmse,
                  pr->icp_max_iter};
    DevBuf<float> xyz;
    B3D_TRY(xyz.alloc(ctx, (size_t)(3 * N * F)));
    if (device_inputs) {
        // 1. rasters already resident: two deprojection launches (sources, targets), one sort over all 2P frames
        B3D_TRY(deproject_z16_batch(ctx, depth_src, nullptr, pr->w, pr->h, P, pr->fx, pr->fy, pr->ppx, pr->ppy, pr->depth_scale, xyz.p, nullptr));
        B3D_TRY(deproject_z16_batch(ctx, depth_tgt, nullptr, pr->w, pr->h, P, pr->fx, pr->fy, pr->ppx, pr->ppy, pr->depth_scale, xyz.p + 3 * N * P, nullptr));
        return register_clouds_f32(ctx, xyz, raw_off, P, &bp, results_h);
    }
    // 1'. HOST rasters (the end-to-end leg): the frames go through the front end in groups (by default the sources, then the
    // targets); the host -> device copy of group g + 1 runs on the copy stream while group g is deprojected and voxelised on
    // the compute stream. Per-cloud results do not depend on the grouping (every cloud is sorted on its own).
    if (ctx->copy_stream == nullptr) B3D_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    DevBuf<uint16_t> depth_d;
    B3D_TRY(depth_d.alloc(ctx, (size_t)(N * F)));
    static const int halves_env = getenv("B3D_E2E_HALVES") ? atoi(getenv("B3D_E2E_HALVES")) : 1;  // groups per side; every group costs ~0.6 ms of fixed work (profiles/r02_e2e_groups.txt)
    const int halves = std::max(1, std::min(halves_env, P));
    struct Group { int f0, f1; };
    std::vector<Group> groups;
    for (int side = 0; side < 2; ++side)
        for (int hh = 0; hh < halves; ++hh) groups.push_back({side * P + (P * hh) / halves, side * P + (P * (hh + 1)) / halves});
    std::vector<cudaEvent_t> ev(groups.size() + 1);
    for (auto& e : ev) B3D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // the staging block may be a recycled scratch block: the copies start after everything already queued on the compute stream
    B3D_CUDA(cudaEventRecord(ev.back(), ctx->stream));
    B3D_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ev.back(), 0));
    for (size_t g = 0; g < groups.size(); ++g) {
        const Group& G = groups[g];
        const uint16_t* host = G.f0 < P ? depth_src + N * G.f0 : depth_tgt + N * (G.f0 - P);
        B3D_CUDA(cudaMemcpyAsync(depth_d.p + N * G.f0, host, (size_t)(N * (G.f1 - G.f0)) * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->copy_stream));
        B3D_CUDA(cudaEventRecord(ev[g], ctx->copy_stream));
    }
    DevBuf<double> vox;
    B3D_TRY(vox.alloc(ctx, (size_t)(3 * N * F)));
    std::vector<int32_t> voff(F + 1, 0);
    int64_t m_done = 0;
    int rc = B3D_OK;
    for (size_t g = 0; g < groups.size() && rc == B3D_OK; ++g) {
        const Group& G = groups[g];
        if (cudaStreamWaitEvent(ctx->stream, ev[g], 0) != cudaSuccess) rc = set_error(B3D_E_CUDA, "cudaStreamWaitEvent failed");
        if (rc == B3D_OK)
            rc = deproject_z16_batch(ctx, depth_d.p + N * G.f0, nullptr, pr->w, pr->h, G.f1 - G.f0, pr->fx, pr->fy, pr->ppx, pr->ppy, pr->depth_scale,
                                     xyz.p + 3 * N * G.f0, nullptr);
        if (rc == B3D_OK) rc = voxel_group(ctx, xyz.p, raw_off, G.f0, G.f1, pr->voxel_size, vox, &voff, &m_done);
    }
    if (rc != B3D_OK) cudaStreamSynchronize(ctx->copy_stream);  // nothing may still write into the staging block when it is released
    for (auto& e : ev) cudaEventDestroy(e);
    B3D_TRY(rc);
    depth_d.release();
    xyz.release();
    return register_voxels_f32(ctx, vox, voff, raw_off, P, &bp, results_h);
}

int b3d_register_disparity_pairs(b3d_ctx* ctx, const b3d_disparity_params* pr, const int16_t* disp_src, const int16_t* disp_tgt, int n_pairs,
                                 int device_inputs, b3d_pair_result* results_h) {
    B3D_REQUIRE(ctx != nullptr && pr != nullptr && results_h != nullptr, "b3d_register_disparity_pairs: NULL argument");
    B3D_REQUIRE(n_pairs >= 1, "b3d_register_disparity_pairs: n_pairs must be >= 1");
    B3D_REQUIRE(pr->w > 0 && pr->h > 0, "b3d_register_disparity_pairs: empty image");
    B3D_REQUIRE(disp_src && disp_tgt, "b3d_register_disparity_pairs: NULL disparity buffer");
    B3D_REQUIRE(pr->voxel_size > 0.0f, "voxel_size must be positive.");
    B3D_REQUIRE(pr->icp_max_dist > 0.0, "Invalid max_correspondence_distance.");
    B3D_REQUIRE(pr->icp_kind >= 0 && pr->icp_kind <= 2, "unknown ICP kind %d", pr->icp_kind);
    B3D_REQUIRE(pr->normals_max_nn >= 1 && pr->normals_max_nn <= 64, "normals_max_nn must be in [1, 64]");
    B3D_REQUIRE(pr->icp_max_iter >= 0, "negative icp_max_iter");
    B3D_TRY(ctx->bind());
    const int P = n_pairs, F = 2 * P;
    const int64_t N = (int64_t)pr->w * pr->h;
    B3D_REQUIRE(N * F < (int64_t)INT32_MAX, "batch of %d frames x %lld pixels exceeds 2^31-1 points", F, (long long)N);
    DevBuf<int16_t> disp_d;
    B3D_TRY(disp_d.alloc(ctx, (size_t)(N * F)));  // sources then targets, contiguous for the single compaction pass
    const cudaMemcpyKind kind = device_inputs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    B3D_CUDA(cudaMemcpyAsync(disp_d.p, disp_src, (size_t)(N * P) * sizeof(int16_t), kind, ctx->stream));
    B3D_CUDA(cudaMemcpyAsync(disp_d.p + N * P, disp_tgt, (size_t)(N * P) * sizeof(int16_t), kind, ctx->stream));
    DevBuf<float> xyz;
    B3D_TRY(xyz.alloc(ctx, (size_t)(3 * N * F)));
    std::vector<int32_t> raw_off;
    B3D_TRY(reproject_disparity_valid_batch(ctx, disp_d.p, pr->w, pr->h, F, pr->Q, pr->min_disp16, xyz.p, &raw_off));
    disp_d.release();
    for (int f = 0; f < F; ++f)
        B3D_REQUIRE(raw_off[f + 1] > raw_off[f], "b3d_register_disparity_pairs: frame %d has no valid disparity", f);
    BackParams bp{pr->voxel_size, pr->normals_max_nn, pr->normals_radius, pr->icp_kind, pr->icp_max_dist, pr->icp_rel_fitness, pr->icp_rel_rmse,
                  pr->icp_max_iter};
    return register_clouds_f32(ctx, xyz, raw_off, P, &bp, results_h);
}

int b3d_register_depth_pair(b3d_ctx* ctx, const b3d_pair_params* params, const uint16_t* depth_src, const uint16_t* depth_tgt, int device_inputs,
                            b3d_pair_result* result_h) {
    return b3d_register_depth_pairs(ctx, params, depth_src, depth_tgt, 1, device_inputs, result_h);
}

}  // extern "C"
