/*
 * oracle/orc.h — CPU restatement of the MAVMAP hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call it.
 * The product library (mavmap_b200/libmavmap_b200.so) never links or dlopens this.
 *
 * Parity pinning status:
 *   camera models   pinned: reference camera_models_test.cc vectors + the reference
 *                   header itself compiled into oracle/_ref (oracle/build_ref.sh)
 *   triangulation   pinned: reference triangulation_test.cc vectors (6 pts x 25 poses)
 *   matcher         pinned against the library the reference delegates to
 *                   (cv2.BFMatcher, tests/golden/match_*.npz); the reference itself
 *                   has no matcher test
 *   BA / pose_ref   PARITY UNPINNED: the reference has no BA test and Ceres is not in
 *                   the container; the LM rules below restate Ceres 1.8 from its
 *                   published algorithm (SURVEY.md §8a-a3')
 */
#ifndef ORC_H_
#define ORC_H_

#include "../include/mavmap_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* camera models: camera_models.h:104-359 */
void orc_world2image(int model, const double* params, double x, double y, double z,
                     double* u, double* v);
void orc_image2world(int model, const double* params, double u, double v,
                     double* x, double* y, double* z);
int  orc_camera_world2image(int model, const double* params, int64_t n, const double* xyz, double* uv);
int  orc_camera_image2world(int model, const double* params, int64_t n, const double* uv, double* xyz);
int  orc_camera_image2world_normalized(int model, const double* params, int64_t n,
                                       const double* uv, double* xy);

/* BACostFunction<M>::operator() (bundle_adjustment.h:131-159) evaluated with forward-mode
 * dual numbers exactly as ceres::AutoDiffCostFunction<.,2,3,1,1,1,3,N> does.
 * r[2]; J[2*18] row-major, columns: rvec(3) tx ty tz X(3) intr(9, unused = 0). */
void orc_ba_residual_jet(int model, const double* pose6, const double* X, const double* intr,
                         const double* obs, double* r, double* J);
/* same residual with plain doubles (no Jacobian) */
void orc_ba_residual(int model, const double* pose6, const double* X, const double* intr,
                     const double* obs, double* r);

/* triangulation.cc:12-50, projection.cc:107-149, triangulation.cc:101-147 */
int orc_triangulate_two_view(const double* P1, const double* P2, int64_t n,
                             const double* x1, const double* x2, double* X,
                             double* reproj1, double* reproj2,
                             double* depth1, double* depth2, double* angle);

int orc_tri_angles(const double* P1, const double* P2, int64_t n, const double* X, double* angle);

/* feature.cc:52-133 with exact (double-accumulated, float-rounded) L2 distances */
int orc_match_pair(const float* d1, int32_t n1, const float* d2, int32_t n2, int32_t k,
                   const float* xy1, const float* xy2, const mm_match_options* opt,
                   int32_t* q, int32_t* t, float* dist, int32_t* n_out);
/* knnMatch(k=2) of one direction: idx[n1*2] (-1 = absent), dist[n1*2] */
int orc_knn2(const float* d1, int32_t n1, const float* d2, int32_t n2, int32_t k,
             const float* xy1, const float* xy2, double max_distance,
             int32_t* idx, float* dist);

/* bundle_adjustment.cc:449-613 with Ceres 1.8 LM + SPARSE_SCHUR semantics */
int orc_ba_solve(mm_ba_problem* problem, const mm_ba_options* opt, mm_ba_summary* summary);
int orc_pose_refine(double* rvec, double* tvec, int model_code, const double* params,
                    int64_t n, const double* points2D, const double* points3D,
                    const uint8_t* inlier_mask, const mm_ba_options* opt,
                    mm_ba_summary* summary, double* ret);
/* cost only: 1/2 sum rho(|r|^2) at the problem's current parameters */
int orc_ransac_score(int kind, const double* models, int n_models, int64_t n, const double* x, const double* y, double threshold,
                     int32_t* num_inliers, double* residual_sum, int32_t* best, double* best_residuals, uint8_t* best_mask);
double orc_ba_cost(const mm_ba_problem* problem, const mm_ba_options* opt);
/* BARotationConstraintCostFunction (bundle_adjustment.cc:57-111): residual and d r / d rvec (J may be NULL) */
double orc_rot_prior(const double* rvec, const double* rvec0, double weight, double* J);
void orc_ba_options_default(mm_ba_options* o);
int  orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
