/*
 * mavmap_b200.h — flat C ABI of the B200-native MAVMAP hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI
 * layer: its boundary is the C++ free functions listed below.  Each entry
 * point here is what a same-named C++ shim (mavmap_b200/shim/, see
 * INTEGRATION.md) marshals into; signatures use only plain pointers and
 * sizes.  All buffers named "host" are caller-owned host memory; "_dev"
 * entry points take device pointers (resident in HBM) and a CUDA stream.
 *
 * Reference interface each group replaces (paths relative to the
 * mavmap/mavmap tree):
 *   camera models      src/base3d/camera_models.h:104-423, camera_models.cc:12-52
 *   triangulation      src/base3d/triangulation.cc:12-147, projection.cc:107-149
 *   matching           src/base2d/feature.h:102-110, feature.cc:12-133
 *   bundle adjustment  src/base3d/bundle_adjustment.h:38-230,
 *                      bundle_adjustment.cc:139-225 (pose_refinement),
 *                      :228-613 (bundle_adjustment)
 *
 * There is no CPU fallback behind this ABI: every compute entry point
 * returns MM_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef MAVMAP_B200_H_
#define MAVMAP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_ABI_VERSION 1

/* ---- status codes -------------------------------------------------------- */
enum {
  MM_OK = 0,
  MM_ERR_INVALID_ARG = 1,     /* null pointer, negative size, bad model code  */
  MM_ERR_DATUM = 2,           /* < 7 fixed dof  -> std::invalid_argument
                                 (bundle_adjustment.cc:459-466)               */
  MM_ERR_MIN_TRACK_LEN = 3,   /* min_track_len < 2 -> std::invalid_argument
                                 (bundle_adjustment.cc:468-471)               */
  MM_ERR_NO_DEVICE = 4,       /* no usable CUDA device / kernel image         */
  MM_ERR_CUDA = 5,            /* CUDA runtime error (see mm_last_error)       */
  MM_ERR_ALLOC = 6,
  MM_ERR_NUMERICAL = 7,       /* non-finite cost / linear solve failure       */
  MM_ERR_UNSUPPORTED = 8
};

/* ---- camera models (camera_models.h:104-359) ----------------------------- */
#define MM_MODEL_PINHOLE 1    /* fx fy cx cy                    (h:98-100)    */
#define MM_MODEL_OPENCV  2    /* fx fy cx cy k1 k2 p1 p2        (h:157-159)   */
#define MM_MODEL_CATA    3    /* fx fy cx cy k1 k2 p1 p2 xi     (h:264-266)   */
#define MM_INTR_STRIDE   9    /* doubles reserved per camera in every array   */

int mm_abi_version(void);
const char* mm_last_error(void);
/* number of usable CUDA devices (0 when none); never throws */
int mm_device_count(void);
/* number of kernels this library has launched so far in this process */
uint64_t mm_kernel_launch_count(void);

/* camera_model_name_to_code (camera_models.cc:12-21): "PINHOLE"/"OPENCV"/"CATA" -> 1/2/3, else -1 */
int mm_camera_model_name_to_code(const char* name);
/* number of parameters of a model code (4/8/9), -1 if unknown */
int mm_camera_model_num_params(int model_code);
/* camera_model_image2world_threshold (camera_models.cc:47-52) */
double mm_camera_image2world_threshold(double threshold, int model_code, const double* params);

/* camera_model_world2image (camera_models.h:375-392), batched.
 * xyz: [n*3] camera-frame coordinates, uv: [n*2] pixels.  Host buffers. */
int mm_camera_world2image(int model_code, const double* params, int64_t n,
                          const double* xyz, double* uv);
/* camera_model_image2world (camera_models.h:408-423), batched: uv [n*2] -> xyz [n*3]
 * (10 fixed-point undistortion iterations, camera_models.h:210-217,319-326). */
int mm_camera_image2world(int model_code, const double* params, int64_t n,
                          const double* uv, double* xyz);
/* vector overload camera_models.cc:24-44: uv [n*2] -> normalised (x/z, y/z) [n*2] */
int mm_camera_image2world_normalized(int model_code, const double* params, int64_t n,
                                     const double* uv, double* xy);

/* ---- two-view triangulation + filters ------------------------------------ */
/* triangulate_points (triangulation.cc:53-74) fused with the per-correspondence
 * filters the mapper evaluates next (sequential_mapper.cc:786-801):
 *   X        [n*3]  6x4 DLT null vector, dehomogenised (triangulation.cc:12-50)
 *   reproj1/2[n]    calc_reproj_errors against P1 / P2 (projection.cc:107-130)
 *   depth1/2 [n]    calc_depth against P1 / P2 (projection.cc:133-149)
 *   angle    [n]    calc_tri_angles (triangulation.cc:101-147), NaN -> 0
 * P1, P2: row-major 3x4.  x1, x2: [n*2] normalised image coordinates.
 * Any output pointer except X may be NULL.  Host buffers. */
int mm_triangulate_two_view(const double* P1, const double* P2, int64_t n,
                            const double* x1, const double* x2, double* X,
                            double* reproj1, double* reproj2,
                            double* depth1, double* depth2, double* angle);

/* calc_tri_angles (triangulation.cc:101-147) for GIVEN 3-D points X [n*3]: the angle between the two rays of each point
 * (law of cosines with the camera centres of P1 and P2, NaN -> 0).  angle [n].  Host buffers. */
int mm_tri_angles(const double* P1, const double* P2, int64_t n, const double* X, double* angle);

/* calc_reproj_errors (projection.cc:107-130) and calc_depth (projection.cc:133-149) for given 3-D points
 * (the continued-track filter of sequential_mapper.cc:767-777).  x2d [n*2] normalised coordinates, X [n*3];
 * err and depth are [n]; either output may be NULL.  Host buffers. */
int mm_reproj_errors(const double* P, int64_t n, const double* x2d, const double* X, double* err, double* depth);

/* RANSAC hypothesis scoring (util/estimation.cc:83-126): for every model the number of inliers |r| <= threshold and
 * the sum of |r| over the inliers, with the residual of
 *   kind 0  P3PEstimator::residuals (p3p.cc:172-199):  x = points2D [n*2], y = points3D [n*3], model = [R|t] 3x4 row-major
 *   kind 1  ProjectiveTransformEstimator::residuals (projective_transform.cc:48-74): x = src [n*2], y = dst [n*2], model = H 3x3
 *   kind 2  EssentialMatrixEstimator::residuals (essential_matrix.cc:131-162, signed Sampson distance): x, y [n*2], model = E 3x3
 * `best` = index the reference's rule selects (most inliers, then smallest residual sum); best_residuals [n] and
 * best_mask [n] (optional) are evaluated for that model.  Hypotheses are generated by the caller.  Host buffers. */
#define MM_RANSAC_P3P 0
#define MM_RANSAC_HOMOGRAPHY 1
#define MM_RANSAC_ESSENTIAL 2
int mm_ransac_score(int32_t kind, const double* models, int32_t n_models, int64_t n, const double* x, const double* y,
                    double threshold, int32_t* num_inliers, double* residual_sum, int32_t* best,
                    double* best_residuals, uint8_t* best_mask);

/* ---- brute-force descriptor matching (feature.cc:52-133) ------------------ */
#define MM_MATCH_IMPL_AUTO    0   /* tcgen05 tensor-core path when shapes allow */
#define MM_MATCH_IMPL_SIMT    1   /* exact CUDA-core path (verification mode)   */
#define MM_MATCH_IMPL_TCGEN05 2

typedef struct mm_match_options {
  int32_t ratio_test;      /* feature.h:107  default true                     */
  double  max_ratio;       /* feature.h:108  default 0.6; mapper passes 0.9   */
  double  max_distance;    /* feature.h:109  -1 = no keypoint-distance mask   */
  int32_t impl;            /* MM_MATCH_IMPL_*                                 */
} mm_match_options;

void mm_match_options_default(mm_match_options* o);

/* match_brute_force for one image pair, host buffers.
 *   d1 [n1*k], d2 [n2*k] fp32 row-major (cv::Mat CV_32F), L2 norm.
 *   xy1 [n1*2], xy2 [n2*2] keypoint .pt (only read when max_distance >= 0).
 *   q, t, dist: caller-allocated, capacity >= min(n1,n2); *n_out matches written,
 *   ascending queryIdx, DMatch{queryIdx=q, trainIdx=t, distance=dist}. */
int mm_match_pair(const float* d1, int32_t n1, const float* d2, int32_t n2, int32_t k,
                  const float* xy1, const float* xy2, const mm_match_options* opt,
                  int32_t* q, int32_t* t, float* dist, int32_t* n_out);
/* mm_match_pair keeps the descriptor arrays of its last four images on the device (the mapper matches image i
 * against i-1 and i-2, sequential_mapper.cc process()); a host array is recognised by address, shape and a hash of a
 * sample of its content taken on every call (everything up to 32 KB; above that the first and last row and 64 bytes
 * of every 4 KB page).  MM_MATCH_PAIR_NO_CACHE=1 in the environment uploads both arrays on every call.
 * Counters since process start: calls, descriptor arrays uploaded, bytes uploaded. */
void mm_match_pair_counters(uint64_t* calls, uint64_t* arrays_uploaded, uint64_t* bytes_uploaded);

/* Resident descriptor set for many-pair jobs (configs[2]: all pairs of a sequence).
 * Descriptors are uploaded once; pairs are then matched from HBM. */
typedef struct mm_match_set mm_match_set;
/* desc: concatenated [sum(counts)*k] fp32 (host); xy: concatenated [sum(counts)*2] or NULL */
int mm_match_set_create(const float* desc, const float* xy, const int32_t* counts,
                        int32_t n_images, int32_t k, mm_match_set** out);
/* same, but desc/xy are DEVICE pointers that stay owned by the caller and must outlive the set */
int mm_match_set_create_dev(const float* desc_dev, const float* xy_dev, const int32_t* counts_host,
                            int32_t n_images, int32_t k, mm_match_set** out);
void mm_match_set_destroy(mm_match_set* s);
/* Match n_pairs pairs (img_a[p], img_b[p]).  Outputs (host): match_off [n_pairs+1]
 * exclusive offsets, and q/t/dist with capacity cap (>= sum over pairs of
 * min(n_a,n_b) is always enough).  Returns MM_ERR_INVALID_ARG if cap is too small. */
int mm_match_set_pairs(mm_match_set* s, const int32_t* img_a, const int32_t* img_b,
                       int32_t n_pairs, const mm_match_options* opt,
                       int64_t* match_off, int32_t* q, int32_t* t, float* dist, int64_t cap);
/* Device-resident variant used by the benchmark: enqueue on `stream` (a cudaStream_t),
 * leave per-pair counts in cnt_dev [n_pairs] and matches in q/t/dist_dev with a fixed
 * stride (slot p starts at p*stride; stride >= min(n_a,n_b) for every pair). No host sync. */
int mm_match_set_pairs_dev(mm_match_set* s, const int32_t* img_a_host, const int32_t* img_b_host,
                           int32_t n_pairs, const mm_match_options* opt,
                           int32_t* cnt_dev, int32_t* q_dev, int32_t* t_dev, float* dist_dev,
                           int32_t stride, void* stream);

/* ---- feature cache files of the reference (src/base2d/feature_cache.cc:126-163) ----
 * "<image>.keypoints" / "<image>.descriptors" as the reference mapper writes them, read without OpenCV (only
 * KeyPoint::pt and CV_32F descriptor matrices are used by match_brute_force).  The 8-byte rows/cols fields of the
 * descriptor header carry their value in the low 32 bits (feature_cache.cc:140-141 writes ints with sizeof(size_t)). */
int mm_feature_cache_info(const char* keypoints_path, const char* descriptors_path,
                          int32_t* n_keypoints, int32_t* rows, int32_t* cols, int32_t* cv_type);
int mm_feature_cache_read(const char* keypoints_path, const char* descriptors_path,
                          float* xy /* [2*rows] or NULL */, float* desc /* [rows*cols] or NULL */,
                          int32_t cap_rows, int32_t cols_expected);
/* All images of a sequence from their cache files into one resident match set (upload included). */
int mm_match_set_create_from_cache(const char* const* keypoints_paths, const char* const* descriptors_paths,
                                   int32_t n_images, mm_match_set** out);

/* ---- bundle adjustment (bundle_adjustment.h / .cc) ------------------------ */
#define MM_POSE_FREE    0     /* BA_POSE_FREE    bundle_adjustment.h:33       */
#define MM_POSE_FIXED   1     /* BA_POSE_FIXED   :34                          */
#define MM_POSE_FIXED_X 2     /* BA_POSE_FIXED_X :35                          */

#define MM_LOSS_TRIVIAL 0
#define MM_LOSS_CAUCHY  1     /* ceres::CauchyLoss(loss_scale) bundle_adjustment.cc:477-478 */

#define MM_SOLVER_PCG      0  /* preconditioned CG on the reduced camera system (device)    */
#define MM_SOLVER_CHOLESKY 1  /* direct Cholesky of the reduced system (oracle; = SPARSE_SCHUR) */

/* preconditioner of the device PCG */
#define MM_PRECOND_AUTO          0  /* sparse tile Cholesky where its tiles fit in memory, else two-level   */
#define MM_PRECOND_TWO_LEVEL     1  /* block-Jacobi + similarity-mode aggregates (fixed intrinsics only)    */
#define MM_PRECOND_TILE_CHOLESKY 2  /* exact sparse tile Cholesky of the reduced system (1-2 PCG iterations) */

/* Flat SoA view of the FeatureManager subset that takes part in one BA
 * (feature_management.h:189-230 flattened per SURVEY.md §8a-a9).
 * Observations are given in the reference's residual-block order
 * (bundle_adjustment.cc:511-533); the library re-sorts internally. */
typedef struct mm_ba_problem {
  int32_t n_img, n_cam, n_pt;
  int64_t n_obs;
  double*        poses;       /* [n_img*6] rvec(3) tvec(3), in/out            */
  const uint8_t* pose_const;  /* [n_img*4] 1 = constant: {rvec, tx, ty, tz}
                                 (parameter-block granularity of
                                 bundle_adjustment.h:127, frozen per
                                 bundle_adjustment.cc:361-385)                */
  const int32_t* img_cam;     /* [n_img] camera index of each image           */
  double*        intr;        /* [n_cam*MM_INTR_STRIDE], in/out               */
  const int32_t* cam_model;   /* [n_cam] MM_MODEL_*                           */
  const uint8_t* intr_const;  /* [n_cam] 1 = constant                         */
  double*        pts;         /* [n_pt*3], in/out                             */
  const uint8_t* pt_const;    /* [n_pt] 1 = constant (GCP, .cc:545-549; all
                                 points in pose_refinement, .cc:187)          */
  const double*  obs_xy;      /* [n_obs*2] pixels                             */
  const int32_t* obs_img;     /* [n_obs]                                      */
  const int32_t* obs_pt;      /* [n_obs]                                      */
  double*        pt_err;      /* optional [n_pt]: mean raw residual norm per
                                 point (.cc:575-598); NULL to skip            */
  /* constrain_rotation (bundle_adjustment.cc:390-446, bundle_adjustment.h:191-209): one extra residual
   * weight * ||R(rvec)' - R(rvec0)||_F per constrained image (with the reference's index quirk, .cc:103),
   * no loss function.  NULL / weight 0 = no constraint.  The caller has already rotated the scene into the
   * frame of the constraints (.cc:412-425; the shim does it). */
  const double*  rot_prior;   /* optional [n_img*3] rvec0                     */
  const double*  rot_prior_w; /* optional [n_img] weight (0 = unconstrained)  */
} mm_ba_problem;

typedef struct mm_ba_options {
  /* mirror of BundleAdjustmentOptions (bundle_adjustment.h:38-114) */
  int32_t max_num_iterations;   /* 100  */
  double  function_tolerance;   /* 1e-4 */
  double  gradient_tolerance;   /* 1e-8 */
  int32_t loss_type;            /* MM_LOSS_CAUCHY */
  double  loss_scale;           /* loss_scale_factor = 1 */
  /* Ceres 1.8 trust-region defaults the reference inherits (SURVEY §8a-a3') */
  double  parameter_tolerance;        /* 1e-8  */
  double  initial_trust_region_radius;/* 1e4   */
  double  max_trust_region_radius;    /* 1e16  */
  double  min_trust_region_radius;    /* 1e-32 */
  double  min_relative_decrease;      /* 1e-3  */
  double  min_lm_diagonal;            /* 1e-6  */
  double  max_lm_diagonal;            /* 1e32  */
  int32_t jacobi_scaling;             /* 1     */
  int32_t max_num_consecutive_invalid_steps; /* 10 (.cc:559) */
  /* engine knobs */
  int32_t linear_solver;        /* MM_SOLVER_PCG */
  double  pcg_tolerance;        /* relative residual ||S y - b|| / ||b||, 1e-13 */
  int32_t pcg_max_iterations;   /* 2000 */
  int32_t print_progress;       /* 0 */
  int32_t pcg_preconditioner;   /* MM_PRECOND_AUTO */
  double  tile_cholesky_tolerance; /* 1e-8: with the exact tile factorisation as preconditioner the first PCG iteration IS
                                   the direct solve the reference performs (SPARSE_SCHUR, bundle_adjustment.cc:555); further
                                   iterations refine it only while ||S y - b|| / ||b|| exceeds this (pcg_tolerance applies to
                                   the iterative preconditioners) */
} mm_ba_options;

void mm_ba_options_default(mm_ba_options* o);

#define MM_BA_TRACE_MAX 512
enum { MM_TERM_NO_CONVERGENCE = 0, MM_TERM_FUNCTION_TOLERANCE = 1,
       MM_TERM_GRADIENT_TOLERANCE = 2, MM_TERM_PARAMETER_TOLERANCE = 3,
       MM_TERM_NUMERICAL_FAILURE = 4, MM_TERM_EMPTY = 5 };

typedef struct mm_ba_summary {
  double  initial_cost;        /* 1/2 sum rho(|r|^2) at the start             */
  double  final_cost;
  int64_t num_residuals;       /* scalar residuals = 2 * n_obs (.cc:610)      */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination;         /* MM_TERM_*                                   */
  int32_t num_iterations;      /* entries used in the traces below (incl. iteration 0) */
  double  return_value;        /* sqrt(final_cost / num_residuals) (.cc:610)  */
  double  trace_cost[MM_BA_TRACE_MAX];
  double  trace_radius[MM_BA_TRACE_MAX];
  double  trace_gradient_max_norm[MM_BA_TRACE_MAX];
  int32_t trace_accepted[MM_BA_TRACE_MAX];
  int32_t trace_linear_iterations[MM_BA_TRACE_MAX];
  /* device-time breakdown in ms (CUDA events), accumulated over the solve */
  double  ms_setup, ms_linearize, ms_schur, ms_pcg, ms_update, ms_total;
} mm_ba_summary;

/* One-shot solve on host buffers: upload, structure setup, LM loop on the
 * device, download into problem->poses/intr/pts (and pt_err). */
int mm_ba_solve(mm_ba_problem* problem, const mm_ba_options* opt, mm_ba_summary* summary);

/* Resident-session API (what bench.py times with inputs already in HBM). */
typedef struct mm_ba_session mm_ba_session;
int  mm_ba_session_create(const mm_ba_problem* problem, const mm_ba_options* opt,
                          void* stream, mm_ba_session** out);
/* Multi-GPU bundle adjustment: one process per GPU, every rank creates its session from the SAME full
 * problem plus (rank, world).  The 3-D points - and with them the observations, the Jacobian records (160 B each, the
 * large part of the footprint) and K1/K2/K4 - are partitioned across the ranks; poses, intrinsics and the reduced
 * camera system are replicated.  `allreduce` must sum `count` doubles at device pointer `buf` over all ranks in place,
 * ordered on `stream` (ncclAllReduce on that stream, or torch.distributed.all_reduce on a tensor that wraps the
 * pointer); it is called once per Schur assembly (S | rhs | gradient | LM diagonal | scalars), once per step
 * evaluation (4 doubles), once per solve for x, and at download.  Returns 0 on success. */
typedef int (*mm_allreduce_fn)(void* user, double* buf, int64_t count, void* stream);
int  mm_ba_session_create_sharded(const mm_ba_problem* problem, const mm_ba_options* opt, void* stream,
                                  int32_t rank, int32_t world, mm_allreduce_fn allreduce, void* user,
                                  mm_ba_session** out);
/* Reset parameters to the values given at creation and restart the LM state. */
int  mm_ba_session_reset(mm_ba_session* s);
/* Run up to n LM iterations (each = linearize + Schur + PCG + candidate evaluation +
 * accept/reject); stops early on convergence.  *n_done receives how many ran. */
int  mm_ba_session_iterate(mm_ba_session* s, int32_t n, int32_t* n_done);
/* Copy current parameters to host arrays (any may be NULL). */
int  mm_ba_session_download(mm_ba_session* s, double* poses, double* intr, double* pts,
                            double* pt_err);
int  mm_ba_session_summary(mm_ba_session* s, mm_ba_summary* out);
/* Time single hot kernels on the session's current state (bench/roofline):
 * which: 0 = residual+Jacobian (K1), 1 = Schur assembly (K2), 2 = cost-only evaluation (K4),
 * 3 = one PCG iteration's SpMV, 4 = coarse-level setup (assembly + inverse), 5 = assembly + numeric factorisation of the sparse
 * tile Cholesky, 6 = one application of it (both substitutions).  Runs `reps` launches, returns mean ms via CUDA events. */
int  mm_ba_session_time_kernel(mm_ba_session* s, int32_t which, int32_t reps, double* ms);
/* Sizes of the reduced system: number of 6x6 blocks stored (upper triangle incl. diagonal). */
int64_t mm_ba_session_num_blocks(mm_ba_session* s);
/* What preconditions the PCG of this session: out8 = {kind (0 block-Jacobi / dense Cholesky of a small system, 1 two-level,
 * 2 sparse tile Cholesky), tiles of L, tile products per factorisation, flop per factorisation, tile rows, tiles of the
 * node-block inverses, substitution tasks, MB of tile storage}. */
int  mm_ba_session_solver_info(mm_ba_session* s, double* out8);
/* Unknowns of the coarse level of the two-level PCG preconditioner (7 per aggregate of images;
 * 0 = block-Jacobi only: small systems, refined intrinsics, or MM_PCG_NO_COARSE set). */
int32_t mm_ba_session_coarse_dim(mm_ba_session* s);
/* Test hook for the blocked Gauss-Jordan kernel that inverts the coarse matrix: a (m x m, row-major,
 * symmetric positive definite, host memory) is replaced by its inverse. */
int  mm_debug_spd_inverse(double* a, int32_t m);
/* Test hooks of the sparse tile Cholesky that preconditions the PCG (csrc/tilechol_plan.h, tilechol.cuh).
 * plan_create runs the host-side symbolic analysis (nested dissection, tile structure, task lists) for a block graph given by
 * its off-diagonal blocks (blk_a[e], blk_b[e]); pos = optional camera centres [3 n_img]; n_cam_border = cameras whose 9
 * intrinsics form the dense border.  plan_array copies one of the plan's integer arrays (widened to int64) and returns its
 * length (out may be NULL): 0 scalars {nt_pose, nt, n_l, n_upd, n_nodes, max_height, flops, T, TI, n_w, n_wtask, n_slots,
 * n_stasks, n_tnodes}, 1 img_tile, 2 img_slot, 3 tile_nunk, 4 col_ptr, 5 row_idx, 6 col_idx, 7 has_a, 8 upd_ptr, 9 upd_a,
 * 10 upd_b, 11 rowp_ptr, 12 rowp_tile, 13 rowp_col, 14 unk_of, 15 sc_tile, 16 sc_off, 17 tile_height, 18 tile_node,
 * 19 node_first, 20 node_nt, 21 w_row_ptr, 22 wt_row, 23 wt_col, 24 wt_store, 25 wupd_ptr, 26 wupd_l, 27 wupd_w, 28 wupd_flag,
 * 29 task_order, 30 st_kind, 31 st_out, 32 st_base, 33 st_tile, 34 st_item_ptr, 35 it_mat, 36 it_src (see
 * csrc/tilechol_plan.h for their meaning).  No device needed.
 * tilechol_solve factorises on the device and returns z = M^-1 rhs (needs a device). */
typedef struct mm_tilechol_plan mm_tilechol_plan;
int  mm_debug_tilechol_plan_create(int32_t n_img, int32_t n_off, const int32_t* blk_a, const int32_t* blk_b, const double* pos,
                                   int32_t n_cam_border, mm_tilechol_plan** out);
int64_t mm_debug_tilechol_plan_array(mm_tilechol_plan* plan, int32_t which, int64_t* out, int64_t cap);
void mm_debug_tilechol_plan_destroy(mm_tilechol_plan* plan);
int  mm_debug_tilechol_solve(int32_t n_img, int32_t n_off, const int32_t* blk_a, const int32_t* blk_b, const double* pos, const double* S,
                             int32_t n_cam_border, const double* Bm, const double* Cm, const double* rhs, double* z, int32_t reps,
                             double* ms_factor, double* ms_apply);
void mm_ba_session_destroy(mm_ba_session* s);

/* pose_refinement (bundle_adjustment.cc:139-225): 6-dof refinement of one pose,
 * points and intrinsics constant.  inlier_mask may be NULL (= all). Returns
 * sqrt(final_cost/num_residuals) in *ret. */
int mm_pose_refine(double* rvec, double* tvec, int model_code, const double* params,
                   int64_t n, const double* points2D, const double* points3D,
                   const uint8_t* inlier_mask, const mm_ba_options* opt,
                   mm_ba_summary* summary, double* ret);

/* A batch of independent pose refinements in ONE kernel launch (one CTA per problem; SURVEY.md 8f-1 "batched across candidate
 * pairs": candidate poses of one image, or the images of a re-localisation sweep after a loop closure).  Problem b owns the 2D-3D
 * pairs offsets[b] .. offsets[b+1] of points2D [2 each] / points3D [3 each] (inliers only), the camera (model_codes[b],
 * params + 9 b) and the pose rvecs + 3 b / tvecs + 3 b (in/out).  Every problem is solved exactly as mm_pose_refine solves it.
 * summaries / rets: optional [n_problems]. */
int mm_pose_refine_batch(int32_t n_problems, double* rvecs, double* tvecs, const int32_t* model_codes, const double* params,
                         const int64_t* offsets, const double* points2D, const double* points3D, const mm_ba_options* opt,
                         mm_ba_summary* summaries, double* rets);

#ifdef __cplusplus
}
#endif
#endif /* MAVMAP_B200_H_ */
