/* b200recon.h -- C ABI of libb200recon.so (hand-written sm_100a CUDA, no torch types).
 *
 * The reference (aagsi/3D_Reconstruction_Project) has no FFI layer of its own: its hot path is a set of
 * Python calls into Open3D / librealsense. Each entry point below names the reference
 * call site (file:line under the reference root) whose arithmetic it replaces. The host-side mirror of the
 * reference's Python classes lives in 3d_reconstruction_project_b200/ and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a maintainer would add to the reference).
 *
 * Conventions
 *  - every function returns 0 (B3D_OK) or a negative B3D_E_* code; b3d_last_error() gives the thread-local text
 *  - pointers are DEVICE pointers unless the name ends in _h (host); sizes are element counts
 *  - the caller owns every buffer; outputs of data-dependent length are written into caller buffers of the stated
 *    worst-case capacity and the count is returned through a host pointer (the call synchronises ctx's stream)
 *  - a b3d_ctx binds one device + one stream and owns scratch memory; it is not thread-safe, but different
 *    contexts may be used concurrently from different threads. Every call sets the CUDA device itself, so
 *    calls may come from any thread (the reference calls from a non-main thread, main.py:56-61).
 *  - "legacy" = Open3D o3d.geometry.* semantics (float64); "tensor" = o3d.t.geometry.* semantics (float32)
 *  - stated tie-break everywhere a nearest neighbour is chosen: smallest d2 = ((dx*dx+dy*dy)+dz*dz) evaluated in
 *    the flavour's dtype without FMA, then smallest original index; radius tests are strict (d2 < r*r)
 *  - down-sampled clouds are returned in ascending (ix,iy,iz) voxel order (the reference's order is hash-map
 *    iteration order, i.e. unspecified)
 */
#ifndef B200RECON_H
#define B200RECON_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B3D_OK 0
#define B3D_E_INVALID (-1)   /* bad argument (mirrors the RuntimeError cases of Open3D) */
#define B3D_E_CUDA (-2)      /* CUDA runtime error */
#define B3D_E_RANGE (-3)     /* voxel / cell grid does not fit 63-bit linear keys ("voxel_size is too small") */
#define B3D_E_NOMEM (-4)
#define B3D_E_STATE (-5)     /* call sequence error */

#define B3D_ICP_POINT_TO_POINT 0
#define B3D_ICP_POINT_TO_PLANE 1
#define B3D_ICP_GENERALIZED 2

typedef struct b3d_ctx b3d_ctx;
typedef struct b3d_grid b3d_grid;       /* spatial hash over one cloud (sorted copy + cell table) */
typedef struct b3d_icp_state b3d_icp_state; /* step-wise ICP (sharded clouds) */

int b3d_version(void);
const char* b3d_last_error(void);

/* stream: the cudaStream_t every call of this context runs on (e.g. torch.cuda.current_stream().cuda_stream);
 * NULL = the legacy default stream */
int b3d_ctx_create(int device, void* stream, b3d_ctx** out);
int b3d_ctx_destroy(b3d_ctx* ctx);
int b3d_ctx_synchronize(b3d_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t b3d_ctx_launch_count(b3d_ctx* ctx);
/* Per-kernel device timing for bench.py's roofline: with profiling enabled every launch is bracketed by CUDA events on the
 * context's stream. b3d_ctx_profile_report writes "<kernel>\t<launches>\t<total_ms>\n" lines (descending total time) into
 * buf (a fourth column carries the bytes the launch sites declared, 0 if none) and clears the records; it returns the bytes needed (call with cap 0 to size the buffer). */
int b3d_ctx_profile(b3d_ctx* ctx, int enable);
int64_t b3d_ctx_profile_report(b3d_ctx* ctx, char* buf, int64_t cap);

/* ---- K1 deprojection ------------------------------------------------------------------------------------ */
/* rs.pointcloud().calculate(depth_frame) + get_vertices()  -- pointcloud_capture.py:35,38 (test/GICP1.py:63-66).
 * Every pixel emitted in raster order, zero depth -> (0,0,0). xyz: float32 [h*w,3]. */
int b3d_deproject_z16(b3d_ctx* ctx, const uint16_t* depth, int w, int h, float fx, float fy, float ppx, float ppy,
                      float depth_scale, float* xyz);
/* Same pass also converting the colour raster to float32 colours/255 (pointcloud_capture.py:39), rgb may be NULL */
int b3d_deproject_z16_color(b3d_ctx* ctx, const uint16_t* depth, const uint8_t* bgr, int w, int h, float fx, float fy,
                            float ppx, float ppy, float depth_scale, float* xyz, float* rgb);
/* RGBDImage.create_from_color_and_depth + PointCloud.create_from_rgbd_image (+ flip diag(1,-1,-1,1))
 * -- test/check84.py:155-178, test/mini1.py:148-171. Valid pixels only (0 < z < trunc), raster order.
 * xyz / rgb: float64 [h*w,3] capacity (rgb and color may be NULL). n_valid_h receives the point count. */
int b3d_deproject_rgbd(b3d_ctx* ctx, const uint16_t* depth, const uint8_t* color, int w, int h, double fx, double fy,
                       double cx, double cy, float depth_scale, float depth_trunc, int flip_yz, double* xyz,
                       double* rgb, int64_t* n_valid_h);
/* cv2.reprojectImageTo3D(disp/16, Q) semantics for the Q matrix loaded (and never used) at
 * Calib_depth/depth4.py:98; disparity is SGBM int16 fixed point x16 (Calib_depth/depth1.py:331).
 * Q_h: 16 doubles row-major on the HOST. xyz: float32 [h*w,3]. */
int b3d_reproject_disparity(b3d_ctx* ctx, const int16_t* disp, int w, int h, const double* Q_h, float* xyz);
/* Same arithmetic, valid pixels only (disparity >= min_disp16, fixed point x16; SGBM marks invalid pixels with
 * (minDisparity - 1) * 16), raster order. xyz: float32 [h*w,3] capacity; n_valid_h receives the point count. */
int b3d_reproject_disparity_valid(b3d_ctx* ctx, const int16_t* disp, int w, int h, const double* Q_h, int min_disp16, float* xyz,
                                  int64_t* n_valid_h);

/* ---- K2 voxel down-sampling ----------------------------------------------------------------------------- */
/* Legacy PointCloud.voxel_down_sample -- pointcloud_alignment.py:22-23, test/check84.py:180, test/mini1.py:174.
 * colors / normals (and their outputs) may be NULL. Outputs have capacity n rows.
 * out_index: int32 [n,3] voxel coordinates, out_count: int32 [n] points per voxel (both optional). */
int b3d_voxel_downsample_legacy(b3d_ctx* ctx, const double* xyz, const double* colors, const double* normals, int64_t n,
                                double voxel_size, double* out_xyz, double* out_colors, double* out_normals,
                                int32_t* out_index, int32_t* out_count, int64_t* m_h);
/* Tensor PointCloud.voxel_down_sample -- pointcloud_capture.py:50, pointcloud_processing.py:27,
 * test/gpu-performance.py:18. attr (e.g. colours) optional. out_index: int64 [n,3]. */
int b3d_voxel_downsample_tensor(b3d_ctx* ctx, const float* xyz, const float* attr, int64_t n, float voxel_size,
                                float* out_xyz, float* out_attr, int64_t* out_index, int32_t* out_count, int64_t* m_h);

/* ---- K2 spatial hash + neighbour search ------------------------------------------------------------------ */
/* Builds the radix-sorted cell table over a cloud (replaces KDTreeFlann construction inside estimate_normals /
 * remove_*_outlier / registration_icp). cell_size <= 0 lets the library choose from (k_hint, radius_hint).
 * is_f64: 1 = double points, 0 = float points. */
int b3d_grid_build(b3d_ctx* ctx, const void* xyz, int64_t n, int is_f64, double cell_size, int k_hint, double radius_hint,
                   b3d_grid** out);
int b3d_grid_destroy(b3d_ctx* ctx, b3d_grid* grid);
int b3d_grid_info(b3d_grid* grid, int64_t* n_h, int64_t* n_cells_h, double* cell_size_h);
/* KDTreeFlann.SearchHybrid(query, radius, k) / SearchKNN (radius <= 0) for nq queries: the k nearest, sorted by
 * (d2, index), cut at d2 < radius^2. idx int32 [nq,k] (-1 padded), d2 [nq,k] in the grid dtype (optional),
 * cnt int32 [nq] (optional). k <= 64. */
int b3d_knn_hybrid(b3d_ctx* ctx, b3d_grid* grid, const void* queries, int64_t nq, int k, double radius, int32_t* idx,
                   void* d2, int32_t* cnt);

/* ---- K3 normals / covariances ---------------------------------------------------------------------------- */
/* Legacy estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) -- pointcloud_alignment.py:27-28,
 * test/GICP1.py:77, check84.py:181-182, mini1.py:176-177. radius <= 0: KDTreeSearchParamKNN(max_nn).
 * prior (optional): existing normals used for the sign rule. */
int b3d_estimate_normals_legacy(b3d_ctx* ctx, const double* xyz, int64_t n, int max_nn, double radius, const double* prior,
                                double* normals);
/* Tensor estimate_normals(max_nn, radius) -- normal_estimation.py:20 */
int b3d_estimate_normals_tensor(b3d_ctx* ctx, const float* xyz, int64_t n, int max_nn, float radius, float* normals);
/* GICP covariances from normals, C = R diag(eps,1,1) R^T -- inside registration_generalized_icp, test/GICP1.py:99-102 */
int b3d_covariances_from_normals(b3d_ctx* ctx, const double* normals, int64_t n, double eps, double* cov);

/* PointCloud.orient_normals_consistent_tangent_plane(k) -- normal_estimation.py:21, visualizer.py:68, test/check2.py:253,
 * test/GICP1.py:200 (always k = 100): Riemannian graph (Euclidean
 * minimum spanning tree + k nearest neighbours, weight 1 - |n_i . n_j|), its minimum spanning tree, signs propagated from the
 * first point of largest z (turned towards +z). normals are flipped IN PLACE; flipped (optional, uint8 [n]) gets 1 where a
 * normal was negated (so a float32 copy of the normals can follow). k <= 128; fewer than 4 points is an error like upstream. */
int b3d_orient_normals_consistent_tangent_plane(b3d_ctx* ctx, const double* xyz, double* normals, int64_t n, int k, uint8_t* flipped);

/* compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius, max_nn)) -- test/mini1.py:244-250, test/check2.py:95-100 (the
 * descriptor behind the reference's RANSAC initialisation). normals are required; max_nn <= 128; radius <= 0: KNN.
 * out: float64 [n, 33] row-major (Open3D's Feature.data is its transpose). */
int b3d_compute_fpfh(b3d_ctx* ctx, const double* xyz, const double* normals, int64_t n, int max_nn, double radius, double* out);

/* ---- global registration from feature matches ------------------------------------------------ */
/* Nearest feature of feat_b for every feature of feat_a (squared L2 in float64, ties to the smallest index): the
 * correspondence search inside registration_ransac_based_on_feature_matching -- test/mini1.py:269, test/check2.py:132.
 * Features are row-major [n, dim] (b3d_compute_fpfh's layout), dim <= 64. nn_out: int32 [na] (-1 when nb == 0). */
int b3d_match_features(b3d_ctx* ctx, const double* feat_a, int64_t na, const double* feat_b, int64_t nb, int dim, int32_t* nn_out);
typedef struct b3d_ransac_result {
    double transformation[16]; /* row-major; identity when nothing was found */
    double fitness, inlier_rmse;
    int64_t n_correspondences; /* inliers of the best hypothesis over the whole source cloud */
    int64_t iterations;        /* hypotheses drawn before the confidence criterion (or max_iteration) stopped the loop */
    int64_t validated;         /* hypotheses that passed the checkers and were validated against the target */
} b3d_ransac_result;
/* registration_ransac_based_on_correspondence with TransformationEstimationPointToPoint(False) -- the loop behind
 * registration_ransac_based_on_feature_matching (test/mini1.py:269-281). corres: device int32 [nc, 2] (source index, target
 * index). edge_similarity / checker_distance <= 0 switch the CorrespondenceCheckerBasedOnEdgeLength / ...BasedOnDistance off.
 * Hypothesis i is a pure function of (seed, i); the result is what a single-threaded run of the library's loop gives with
 * those picks. ransac_n in [3, 8]; like upstream, ransac_n < 3 or max_dist <= 0 gives the empty result. */
int b3d_ransac_correspondence(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const int32_t* corres, int64_t nc,
                              double max_dist, int ransac_n, double edge_similarity, double checker_distance, int64_t max_iteration,
                              double confidence, uint64_t seed, b3d_ransac_result* result_h);

/* registration_fgr_based_on_feature_matching(source, target, source_fpfh, target_fpfh, FastGlobalRegistrationOption(...)) --
 * test/check6.py:236-240, check7.py:245-249, check8.py:244-248, check81.py:242-246. Field names and defaults are the library's. */
typedef struct b3d_fgr_option {
    double division_factor;                 /* 1.4 */
    int use_absolute_scale;                 /* 0 */
    int decrease_mu;                        /* 1 */
    double maximum_correspondence_distance; /* 0.025 */
    int iteration_number;                   /* 64 */
    double tuple_scale;                     /* 0.95 */
    int maximum_tuple_count;                /* 1000 */
    int tuple_test;                         /* 1 */
} b3d_fgr_option;
/* Features row-major [n, dim]. T_h: host double[16], source -> target (identity when fewer than 10 matches survive);
 * n_corres_h (optional): matches the optimisation ran on. The tuple test's random triples are a pure function of (seed, trial). */
int b3d_fgr_feature_matching(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const double* feat_src,
                             const double* feat_tgt, int dim, const b3d_fgr_option* opt, uint64_t seed, double* T_h, int64_t* n_corres_h);

/* ---- outlier filters -------------------------------------------------------------------------------------- */
/* remove_statistical_outlier(nb_neighbors, std_ratio) -- pointcloud_processing.py:35-36, test/mini1.py:175.
 * keep: uint8 [n]; kept_idx: int64 [n] ascending indices (optional); n_kept_h: count. */
int b3d_statistical_outlier(b3d_ctx* ctx, const double* xyz, int64_t n, int nb_neighbors, double std_ratio, uint8_t* keep,
                            int64_t* kept_idx, int64_t* n_kept_h);
/* remove_radius_outlier(nb_points, radius) -- pointcloud_processing.py:39 */
int b3d_radius_outlier(b3d_ctx* ctx, const double* xyz, int64_t n, int nb_points, double radius, uint8_t* keep,
                       int64_t* kept_idx, int64_t* n_kept_h);
/* select_by_index on 3-column double arrays: dst[i] = src[kept_idx[i]] */
int b3d_gather_rows_f64(b3d_ctx* ctx, const double* src, const int64_t* kept_idx, int64_t n_kept, int cols, double* dst);

/* ---- K4 registration -------------------------------------------------------------------------------------- */
typedef struct b3d_icp_result {
    double transformation[16]; /* row-major 4x4 */
    double fitness;
    double inlier_rmse;
    int32_t iterations;       /* number of updates applied */
    int32_t converged;        /* 1 if the relative criteria stopped the loop */
    int64_t n_correspondences;
} b3d_icp_result;

/* source.transform(T) -- pointcloud_alignment.py:42. In place. normals / cov optional. T_h: 16 doubles on the host */
int b3d_transform_f64(b3d_ctx* ctx, const double* T_h, double* xyz, int64_t n, double* normals, double* cov);

/* One correspondence search at a fixed transform (GetRegistrationResultAndCorrespondences): corr int32 [ns]
 * (-1 = none). stats_h: {n_corr, sum_d2}. */
int b3d_icp_correspondences(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const double* T_h,
                            double max_dist, int32_t* corr, double* stats_h);

/* get_information_matrix_from_point_clouds(source, target, max_dist, T) -- test/mini1.py:302, test/check2.py:160 (feeds the
 * reference's pose graph): correspondences of T*source in target, G^T G with G = [-[t]x | I] per TARGET point. info_h: 36 doubles. */
int b3d_information_matrix(b3d_ctx* ctx, const double* src, int64_t ns, const double* tgt, int64_t nt, const double* T_h, double max_dist,
                           double* info_h);

/* registration_icp(source, target, max_dist, init, estimation, criteria) / registration_generalized_icp
 *  P2P -- pointcloud_alignment.py:35-39; P2L -- test/mini1.py:293-296, test/check2.py:151-154; GICP -- test/GICP1.py:99-102.
 * tgt_normals required for P2L; src_cov/tgt_cov ([n,9] double) required for GICP. init_h may be NULL (identity).
 * corr (optional) int32 [ns]: final correspondence set. The whole loop runs on the device; one synchronisation at the end. */
int b3d_icp(b3d_ctx* ctx, int kind, const double* src, int64_t ns, const double* src_cov, const double* tgt, int64_t nt,
            const double* tgt_normals, const double* tgt_cov, double max_dist, const double* init_h, double rel_fitness,
            double rel_rmse, int max_iter, b3d_icp_result* result_h, int32_t* corr);

/* A batch of n_pairs independent registrations in the same launches (grid.y = pair): pair p registers
 * src[src_off_h[p] .. src_off_h[p+1]) onto tgt[tgt_off_h[p] .. tgt_off_h[p+1]) (offsets in points, HOST arrays of n_pairs + 1
 * entries starting at 0; normals / covariances are indexed like their clouds). init_h: n_pairs x 16 doubles or NULL.
 * results_h: [n_pairs]; corr (optional): int32 [total source points], target indices local to the pair's target cloud.
 * A pair's result is bit-identical to the single-pair b3d_icp call. An empty target cloud inside a non-empty batch is an error
 * of the grid build only when ALL targets are empty; empty pairs return fitness 0 and their init. */
int b3d_icp_batch(b3d_ctx* ctx, int kind, int n_pairs, const double* src, const int64_t* src_off_h, const double* src_cov, const double* tgt,
                  const int64_t* tgt_off_h, const double* tgt_normals, const double* tgt_cov, double max_dist, const double* init_h,
                  double rel_fitness, double rel_rmse, int max_iter, b3d_icp_result* results_h, int32_t* corr);

/* Step-wise ICP for one cloud sharded by SOURCE points over several GPUs (BASELINE config 5). Each rank holds a
 * contiguous slice of the source and a replica of the target. Per iteration:
 *   b3d_icp_accumulate  -> fills sums (29 doubles on the device: 21 JtJ upper + 6 Jtr + |C| + sum d2)
 *   <all-reduce(sum) of those 29 doubles by the caller, e.g. torch.distributed / NCCL>
 *   b3d_icp_update      -> every rank solves the same 6x6 system and applies the same update
 * ns_total is the global source size (for fitness). done_h receives 1 when the loop has finished. */
int b3d_icp_begin(b3d_ctx* ctx, int kind, const double* src, int64_t ns_local, int64_t ns_total, const double* src_cov,
                  const double* tgt, int64_t nt, const double* tgt_normals, const double* tgt_cov, double max_dist,
                  const double* init_h, double rel_fitness, double rel_rmse, int max_iter, b3d_icp_state** out);
int b3d_icp_accumulate(b3d_ctx* ctx, b3d_icp_state* st, double** sums_dev_out);
int b3d_icp_update(b3d_ctx* ctx, b3d_icp_state* st, int* done_h);
int b3d_icp_finish(b3d_ctx* ctx, b3d_icp_state* st, b3d_icp_result* result_h, int32_t* corr);
/* Fused variant of accumulate -> all-reduce -> update: every rank owns an exchange buffer of B3D_ICP_PEER_DOUBLES(world)
 * doubles, zero-initialised, mapped into every other rank's address space (CUDA IPC / symmetric memory over NVLink);
 * peer_bufs_h[r] is rank r's buffer as seen from THIS process. b3d_icp_pass_peers launches ONE kernel per pass: the last
 * block writes the 29 local sums into every rank's buffer, waits for all ranks' stamps of this pass, adds the slots in rank
 * order (identical result on every rank) and applies the update. All ranks must call it the same number of times
 * (they see the same done flag); world <= 8. done_h may be NULL (no host synchronisation). */
#define B3D_ICP_PEER_DOUBLES(world) (2 * (world) * 32 + 2 * (world))
int b3d_icp_set_peers(b3d_ctx* ctx, b3d_icp_state* st, int rank, int world, void* const* peer_bufs_h);
int b3d_icp_pass_peers(b3d_ctx* ctx, b3d_icp_state* st, int* done_h);

/* ---- whole-path entry points with HOST buffers (what the reference-facing Python classes call; e2e) -------- */
typedef struct b3d_pair_params {
    int w, h;
    float fx, fy, ppx, ppy, depth_scale; /* librealsense intrinsics of the depth stream */
    float voxel_size;                    /* tensor voxel_down_sample (pointcloud_capture.py:50) */
    int normals_max_nn;                  /* legacy hybrid normals (pointcloud_alignment.py:27-28) */
    double normals_radius;
    int icp_kind;
    double icp_max_dist, icp_rel_fitness, icp_rel_rmse;
    int icp_max_iter;
} b3d_pair_params;

typedef struct b3d_pair_result {
    b3d_icp_result icp;
    int64_t n_raw;        /* pixels deprojected, both frames */
    int64_t m_source, m_target; /* down-sampled cloud sizes */
} b3d_pair_result;

/* Frame pair: depth_src_h / depth_tgt_h are HOST uint16 [h,w] rasters (pinned or pageable). Runs
 * deproject -> tensor voxel -> legacy normals (target, and source for GICP) -> ICP(source -> target) entirely on the
 * device and returns the registration result. device_inputs != 0: the two depth pointers are DEVICE pointers
 * (bench "value" leg, inputs resident in HBM). */
int b3d_register_depth_pair(b3d_ctx* ctx, const b3d_pair_params* params, const uint16_t* depth_src, const uint16_t* depth_tgt,
                            int device_inputs, b3d_pair_result* result_h);
/* A batch of n_pairs independent frame pairs (BASELINE config 4: sequence registration): depth_src / depth_tgt are
 * [n_pairs, h, w] uint16 stacks. Every stage runs ONCE for the whole batch (the cloud id rides in the top bits of the
 * spatial keys), the ICP passes are n_pairs wide. results_h: [n_pairs]. Each pair's result equals the single-pair call. */
int b3d_register_depth_pairs(b3d_ctx* ctx, const b3d_pair_params* params, const uint16_t* depth_src, const uint16_t* depth_tgt,
                             int n_pairs, int device_inputs, b3d_pair_result* results_h);

/* Stereo flavour of the same path (BASELINE config 3): SGBM disparity rasters (int16 x16, Calib_depth/depth1.py:331) +
 * the rectification Q matrix (Calib_depth/depth4.py:98) -> valid-pixel clouds -> tensor voxel -> normals -> ICP / GICP.
 * disp_src / disp_tgt: [n_pairs, h, w] int16 stacks (host, or device when device_inputs != 0). */
typedef struct b3d_disparity_params {
    int w, h;
    double Q[16];      /* row-major 4x4; scale its last row to choose the output unit (e.g. x1000: mm -> m) */
    int min_disp16;    /* smallest valid disparity, fixed point x16 */
    float voxel_size;
    int normals_max_nn;
    double normals_radius;
    int icp_kind;
    double icp_max_dist, icp_rel_fitness, icp_rel_rmse;
    int icp_max_iter;
} b3d_disparity_params;

int b3d_register_disparity_pairs(b3d_ctx* ctx, const b3d_disparity_params* params, const int16_t* disp_src, const int16_t* disp_tgt,
                                 int n_pairs, int device_inputs, b3d_pair_result* results_h);

#ifdef __cplusplus
}
#endif
#endif /* B200RECON_H */
