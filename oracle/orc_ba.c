/*
 * oracle/orc_ba.c — CPU restatement of bundle_adjustment / pose_refinement,
 * TEST INFRASTRUCTURE ONLY (see orc.h).  PARITY UNPINNED: the reference has no BA
 * test and its solver (Ceres-Solver 1.7.0/1.8.0, README.md:38) is not in the tree.
 *
 * Follows (mavmap/mavmap):
 *   bundle_adjustment      src/base3d/bundle_adjustment.cc:449-613
 *     which observations / parameter blocks enter:  :228-387 (flattened by the caller
 *     into mm_ba_problem, see mavmap_b200/ba.py and shim/bundle_adjustment.h)
 *     solver options:  SPARSE_SCHUR :555, tolerances :556-558, invalid steps :559
 *     loss:  LossFunctionWrapper(CauchyLoss(scale)) :477-478
 *     return sqrt(final_cost / num_residuals) :610; point errors :575-598
 *   pose_refinement        src/base3d/bundle_adjustment.cc:139-225
 *   BACostFunction         src/base3d/bundle_adjustment.h:117-165 (orc_camera.c)
 *
 * Ceres 1.8 rules restated from its published algorithm (all "from memory", each in a
 * named function so it can be corrected if a Ceres build ever becomes available):
 *   trust_region_minimizer.cc  LM loop: evaluation order, step validity via
 *                              model_cost_change = -(J d).(r + J d / 2), parameter /
 *                              function tolerance tested BEFORE acceptance, gradient
 *                              tolerance relative to the initial max-norm, Jacobi scaling
 *                              1/(1+||col||) estimated once at the start
 *   levenberg_marquardt_strategy.cc  D = sqrt(clamp(diag(J'J))/radius), radius update
 *                              r /= max(1/3, 1-(2q-1)^3), reject: r /= f, f *= 2
 *   corrector.cc               rho'' <= 0 (always for Cauchy): r, J scaled by sqrt(rho')
 *   loss_function.cc           Cauchy: rho = b log(1+s/b), b = a^2
 *   schur_eliminator / SPARSE_SCHUR: eliminate point blocks, factor the reduced camera
 *                              system exactly (here: skyline Cholesky), back-substitute
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc.h"

#define NCOL 15           /* 6 pose + 9 intrinsics columns per observation */
#define JSTR 36           /* per-observation Jacobian doubles: 2 x (6 + 9 + 3) */

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_ba_options_default(mm_ba_options* o) {
  memset(o, 0, sizeof *o);
  o->max_num_iterations = 100;        /* bundle_adjustment.h:40 */
  o->function_tolerance = 1e-4;       /* :41 */
  o->gradient_tolerance = 1e-8;       /* :42 */
  o->loss_type = MM_LOSS_CAUCHY;
  o->loss_scale = 1.0;                /* :45 */
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1;
  o->max_num_consecutive_invalid_steps = 10;
  o->linear_solver = MM_SOLVER_CHOLESKY;
  o->pcg_tolerance = 1e-13;
  o->pcg_max_iterations = 2000;
  o->print_progress = 0;
}

static int model_np(int m) { return m == MM_MODEL_PINHOLE ? 4 : (m == MM_MODEL_OPENCV ? 8 : (m == MM_MODEL_CATA ? 9 : -1)); }

typedef struct {
  const mm_ba_problem* P; const mm_ba_options* O;
  int n_img, n_cam, n_pt; int64_t n_obs; int nc;
  int64_t* pt_start; int64_t* pt_obs;
  uint8_t *act_c, *act_p;
  double *poses, *intr, *pts;       /* current iterate x */
  double *poses2, *intr2, *pts2;    /* x_plus_delta */
  double *J, *r;                    /* per obs: [Jc 2x6 | Ji 2x9 | Jp 2x3] row-major by residual row; r[2] */
  double *scale_c, *scale_p, *diag_c, *diag_p;
  double *g_c, *g_p;                /* gradient (unscaled J) */
  double *step_c, *step_p;
  /* skyline storage of the reduced system */
  int64_t* rowptr; int* first; double* S; double* rhs; int* reach;
  /* rotation constraints (bundle_adjustment.cc:390-446): residual and d r / d rvec per image (scaled with J) */
  double *pr_r, *pr_J; int n_prior;
} ba_t;

/* ---- BARotationConstraintCostFunction (bundle_adjustment.cc:57-111, bundle_adjustment.h:191-209) --------------
 * r = weight * sqrt(sum_i (R_a(i) - R0_b(i))^2) over the nine index pairs the reference uses: ||R' - R0||_F, except that
 * the 8th term compares R[0][2] (not R[2][1]) with R0[1][2] (.cc:103).  Ceres differentiates it with Jets; the same
 * derivative in closed form is  d R / d w_k = [Jl e_k]x R  (Jl = left Jacobian of SO(3)).  Row-major indices. */
static const int PR_A[9] = { 0, 1, 2, 3, 4, 5, 6, 2, 8 }, PR_B[9] = { 0, 3, 6, 1, 4, 7, 2, 5, 8 };
static void so3_R_Jl(const double* w, double* R, double* Jl) {
  const double t2 = w[0]*w[0] + w[1]*w[1] + w[2]*w[2];
  double A = 1.0, Bc = 0.5, Dd = 1.0/6.0;
  if (t2 > 0.0) {
    const double t = sqrt(t2), sh = sin(0.5*t);
    A = sin(t)/t; Bc = 2.0*sh*sh/t2;
    Dd = t < 0.05 ? 1.0/6.0 - t2/120.0 + t2*t2/5040.0 - t2*t2*t2/362880.0 : (t - sin(t))/(t2*t);
  } else { A = 1.0; Bc = 0.0; Dd = 0.0; }
  const double x = w[0], y = w[1], z = w[2];
  R[0] = 1.0 - Bc*(y*y + z*z); R[1] = -A*z + Bc*x*y; R[2] = A*y + Bc*x*z;
  R[3] = A*z + Bc*x*y; R[4] = 1.0 - Bc*(x*x + z*z); R[5] = -A*x + Bc*y*z;
  R[6] = -A*y + Bc*x*z; R[7] = A*x + Bc*y*z; R[8] = 1.0 - Bc*(x*x + y*y);
  Jl[0] = 1.0 - Dd*(y*y + z*z); Jl[1] = -Bc*z + Dd*x*y; Jl[2] = Bc*y + Dd*x*z;
  Jl[3] = Bc*z + Dd*x*y; Jl[4] = 1.0 - Dd*(x*x + z*z); Jl[5] = -Bc*x + Dd*y*z;
  Jl[6] = -Bc*y + Dd*x*z; Jl[7] = Bc*x + Dd*y*z; Jl[8] = 1.0 - Dd*(x*x + y*y);
}
double orc_rot_prior(const double* rvec, const double* rvec0, double weight, double* J /* [3] or NULL */) {
  double R[9], Jl[9], R0[9], J0[9];
  so3_R_Jl(rvec, R, Jl); so3_R_Jl(rvec0, R0, J0);
  double q = 0.0, d[9];
  for (int i = 0; i < 9; ++i) { d[i] = R[PR_A[i]] - R0[PR_B[i]]; q += d[i]*d[i]; }
  const double n = sqrt(q);
  if (J) for (int k = 0; k < 3; ++k) {
    /* dR = [v]x R with v = Jl[:,k] */
    const double v0 = Jl[k], v1 = Jl[3+k], v2 = Jl[6+k];
    double dR[9];
    for (int c = 0; c < 3; ++c) {
      dR[c]     = v1*R[6+c] - v2*R[3+c];
      dR[3 + c] = v2*R[c]   - v0*R[6+c];
      dR[6 + c] = v0*R[3+c] - v1*R[c];
    }
    double s = 0.0;
    for (int i = 0; i < 9; ++i) s += d[i] * dR[PR_A[i]];
    J[k] = n > 0.0 ? weight * s / n : 0.0;
  }
  return weight * n;
}
static int has_prior(const mm_ba_problem* P, int i) { return P->rot_prior && P->rot_prior_w && P->rot_prior_w[i] != 0.0; }

/* ---- loss (loss_function.cc CauchyLoss; corrector.cc) ---------------------- */
static inline void loss_eval(const mm_ba_options* O, double s, double* rho0, double* sqrt_rho1) {
  if (O->loss_type == MM_LOSS_CAUCHY) {
    const double b = O->loss_scale * O->loss_scale, c = 1.0 / b;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    *rho0 = b * log(sum);
    const double rho1 = inv > DBL_MIN ? inv : DBL_MIN;
    *sqrt_rho1 = sqrt(rho1);          /* rho'' < 0 always => Corrector: scale by sqrt(rho') only */
  } else { *rho0 = s; *sqrt_rho1 = 1.0; }
}

/* ---- evaluation ------------------------------------------------------------ */
static double eval_cost(const ba_t* B, const double* poses, const double* intr, const double* pts) {
  const mm_ba_problem* P = B->P; double cost = 0.0;
  #pragma omp parallel for reduction(+:cost) schedule(static)
  for (int64_t o = 0; o < B->n_obs; ++o) {
    const int im = P->obs_img[o], pt = P->obs_pt[o], cam = P->img_cam[im];
    double r[2];
    orc_ba_residual(P->cam_model[cam], poses + 6*(size_t)im, pts + 3*(size_t)pt, intr + MM_INTR_STRIDE*(size_t)cam, P->obs_xy + 2*o, r);
    double rho0, sr; loss_eval(B->O, r[0]*r[0] + r[1]*r[1], &rho0, &sr);
    cost += 0.5 * rho0;
  }
  for (int i = 0; i < P->n_img; ++i) if (has_prior(P, i)) {          /* NULL loss (.cc:442) */
    const double r = orc_rot_prior(poses + 6*(size_t)i, P->rot_prior + 3*(size_t)i, P->rot_prior_w[i], NULL);
    cost += 0.5 * r * r;
  }
  return cost;
}

static double eval_full(ba_t* B) {        /* cost, robustified r and J (unscaled), gradient */
  const mm_ba_problem* P = B->P; double cost = 0.0;
  #pragma omp parallel for reduction(+:cost) schedule(static)
  for (int64_t o = 0; o < B->n_obs; ++o) {
    const int im = P->obs_img[o], pt = P->obs_pt[o], cam = P->img_cam[im];
    double r[2], Jj[36];
    orc_ba_residual_jet(P->cam_model[cam], B->poses + 6*(size_t)im, B->pts + 3*(size_t)pt,
                        B->intr + MM_INTR_STRIDE*(size_t)cam, P->obs_xy + 2*o, r, Jj);
    double rho0, sr; loss_eval(B->O, r[0]*r[0] + r[1]*r[1], &rho0, &sr);
    cost += 0.5 * rho0;
    double* J = B->J + JSTR*(size_t)o;
    for (int row = 0; row < 2; ++row) {
      const double* src = Jj + 18*row; double* dst = J + 18*row;
      for (int k = 0; k < 6; ++k) dst[k]     = B->act_c[6*(size_t)im + k] ? sr * src[k] : 0.0;
      for (int k = 0; k < 9; ++k) dst[6 + k] = B->act_c[6*(size_t)B->n_img + 9*(size_t)cam + k] ? sr * src[9 + k] : 0.0;
      for (int k = 0; k < 3; ++k) dst[15 + k] = B->act_p[3*(size_t)pt + k] ? sr * src[6 + k] : 0.0;
    }
    B->r[2*o] = sr * r[0]; B->r[2*o+1] = sr * r[1];
  }
  for (int i = 0; i < B->n_img && B->pr_r; ++i) {
    B->pr_r[i] = 0.0; B->pr_J[3*(size_t)i] = B->pr_J[3*(size_t)i+1] = B->pr_J[3*(size_t)i+2] = 0.0;
    if (!has_prior(P, i)) continue;
    double J[3];
    B->pr_r[i] = orc_rot_prior(B->poses + 6*(size_t)i, P->rot_prior + 3*(size_t)i, P->rot_prior_w[i], J);
    for (int k = 0; k < 3; ++k) B->pr_J[3*(size_t)i + k] = B->act_c[6*(size_t)i + k] ? J[k] : 0.0;
    cost += 0.5 * B->pr_r[i] * B->pr_r[i];
  }
  return cost;
}

static void column_sums(const ba_t* B, int squared, double* out_c, double* out_p) {
  /* squared = 1: squared column norms of J; 0: gradient J'r */
  const mm_ba_problem* P = B->P;
  memset(out_c, 0, sizeof(double) * (size_t)B->nc);
  memset(out_p, 0, sizeof(double) * 3 * (size_t)B->n_pt);
  for (int64_t o = 0; o < B->n_obs; ++o) {        /* sequential: deterministic sums */
    const int im = P->obs_img[o], pt = P->obs_pt[o], cam = P->img_cam[im];
    const double* J = B->J + JSTR*(size_t)o; const double* r = B->r + 2*o;
    double* oc = out_c + 6*(size_t)im; double* oi = out_c + 6*(size_t)B->n_img + 9*(size_t)cam; double* op = out_p + 3*(size_t)pt;
    for (int row = 0; row < 2; ++row) {
      const double* j = J + 18*row; const double w = squared ? 0.0 : r[row];
      for (int k = 0; k < 6; ++k) oc[k] += squared ? j[k]*j[k] : j[k]*w;
      for (int k = 0; k < 9; ++k) oi[k] += squared ? j[6+k]*j[6+k] : j[6+k]*w;
      for (int k = 0; k < 3; ++k) op[k] += squared ? j[15+k]*j[15+k] : j[15+k]*w;
    }
  }
  for (int i = 0; i < B->n_img && B->pr_r; ++i) for (int k = 0; k < 3; ++k) {
    const double j = B->pr_J[3*(size_t)i + k];
    out_c[6*(size_t)i + k] += squared ? j*j : j*B->pr_r[i];
  }
}

static void scale_jacobian(ba_t* B) {      /* jacobian->ScaleColumns(scale) */
  const mm_ba_problem* P = B->P;
  #pragma omp parallel for schedule(static)
  for (int64_t o = 0; o < B->n_obs; ++o) {
    const int im = P->obs_img[o], pt = P->obs_pt[o], cam = P->img_cam[im];
    double* J = B->J + JSTR*(size_t)o;
    for (int row = 0; row < 2; ++row) {
      double* j = J + 18*row;
      for (int k = 0; k < 6; ++k) j[k] *= B->scale_c[6*(size_t)im + k];
      for (int k = 0; k < 9; ++k) j[6+k] *= B->scale_c[6*(size_t)B->n_img + 9*(size_t)cam + k];
      for (int k = 0; k < 3; ++k) j[15+k] *= B->scale_p[3*(size_t)pt + k];
    }
  }
  for (int i = 0; i < B->n_img && B->pr_r; ++i) for (int k = 0; k < 3; ++k) B->pr_J[3*(size_t)i + k] *= B->scale_c[6*(size_t)i + k];
}

/* ---- skyline reduced system -------------------------------------------------- */
static inline double* S_at(const ba_t* B, int row, int col) { return B->S + B->rowptr[row] + (col - B->first[row]); }

static int build_structure(ba_t* B) {
  const mm_ba_problem* P = B->P;
  const int n_img = B->n_img, nc = B->nc;
  int* minco = malloc(sizeof(int) * (size_t)(n_img > 0 ? n_img : 1));
  for (int i = 0; i < n_img; ++i) minco[i] = i;
  for (int p = 0; p < B->n_pt; ++p) {
    int m = n_img;
    for (int64_t k = B->pt_start[p]; k < B->pt_start[p+1]; ++k) { const int im = P->obs_img[B->pt_obs[k]]; if (im < m) m = im; }
    for (int64_t k = B->pt_start[p]; k < B->pt_start[p+1]; ++k) { const int im = P->obs_img[B->pt_obs[k]]; if (m < minco[im]) minco[im] = m; }
  }
  B->first = malloc(sizeof(int) * (size_t)nc);
  B->rowptr = malloc(sizeof(int64_t) * ((size_t)nc + 1));
  for (int i = 0; i < n_img; ++i) for (int k = 0; k < 6; ++k) B->first[6*i + k] = 6 * minco[i];
  for (int i = 6*n_img; i < nc; ++i) B->first[i] = 0;
  B->rowptr[0] = 0;
  for (int i = 0; i < nc; ++i) B->rowptr[i+1] = B->rowptr[i] + (i - B->first[i] + 1);
  B->S = malloc(sizeof(double) * (size_t)B->rowptr[nc]);
  B->rhs = malloc(sizeof(double) * (size_t)nc);
  /* reach[j] = largest pose row whose profile contains column j */
  B->reach = malloc(sizeof(int) * (size_t)nc);
  for (int j = 0; j < nc; ++j) B->reach[j] = j;
  for (int i = 0; i < 6*n_img; ++i) if (B->reach[B->first[i]] < i) B->reach[B->first[i]] = i;
  for (int j = 1; j < nc; ++j) if (B->reach[j] < B->reach[j-1]) B->reach[j] = B->reach[j-1] > j ? B->reach[j-1] : j;
  free(minco);
  return (B->S && B->rhs) ? MM_OK : MM_ERR_ALLOC;
}

static int invert3(const double* V, double* Vi) {
  const double a = V[0], b = V[1], c = V[2], d = V[4], e = V[5], f = V[8];
  const double A = d*f - e*e, Bc = -(b*f - c*e), C = b*e - c*d;
  const double det = a*A + b*Bc + c*C;
  if (!(det > 0.0) || !isfinite(det)) return 0;
  const double id = 1.0 / det;
  Vi[0] = A*id; Vi[1] = Bc*id; Vi[2] = C*id;
  Vi[3] = Vi[1]; Vi[4] = (a*f - c*c)*id; Vi[5] = -(a*e - b*c)*id;
  Vi[6] = Vi[2]; Vi[7] = Vi[5]; Vi[8] = (a*d - b*b)*id;
  return 1;
}

/* per-point blocks: V (+D), g_p, and for each obs W (15x3).  Returns track length. */
static int point_blocks(const ba_t* B, int p, double* V, double* gp, double* W /* [k][15][3] */, int maxk) {
  memset(V, 0, 9*sizeof(double)); gp[0] = gp[1] = gp[2] = 0.0;
  int k = 0;
  for (int64_t q = B->pt_start[p]; q < B->pt_start[p+1]; ++q, ++k) {
    const int64_t o = B->pt_obs[q];
    const double* J = B->J + JSTR*(size_t)o; const double* r = B->r + 2*o;
    for (int row = 0; row < 2; ++row) {
      const double* j = J + 18*row;
      for (int a = 0; a < 3; ++a) { gp[a] += j[15+a] * r[row]; for (int b = 0; b < 3; ++b) V[3*a+b] += j[15+a]*j[15+b]; }
    }
    if (W && k < maxk) {
      double* w = W + (size_t)k*NCOL*3;
      for (int a = 0; a < NCOL; ++a) for (int b = 0; b < 3; ++b) w[3*a+b] = J[a]*J[15+b] + J[18+a]*J[18+15+b];
    }
  }
  for (int a = 0; a < 3; ++a) V[4*a] += B->diag_p[3*(size_t)p + a];    /* + D_p^2 */
  return k;
}

static int assemble_and_factor(ba_t* B) {
  const mm_ba_problem* P = B->P; const int nc = B->nc, n_img = B->n_img;
  memset(B->S, 0, sizeof(double) * (size_t)B->rowptr[nc]);
  memset(B->rhs, 0, sizeof(double) * (size_t)nc);
  int fail = 0;
  #pragma omp parallel
  {
    int maxk = 64; double* W = malloc(sizeof(double) * (size_t)maxk*NCOL*3); double* Y = malloc(sizeof(double) * (size_t)maxk*NCOL*3);
    int* cols = malloc(sizeof(int) * (size_t)maxk*NCOL);
    #pragma omp for schedule(dynamic, 64)
    for (int p = 0; p < B->n_pt; ++p) {
      const int k = (int)(B->pt_start[p+1] - B->pt_start[p]);
      if (k == 0) continue;
      if (k > maxk) { maxk = 2*k; W = realloc(W, sizeof(double)*(size_t)maxk*NCOL*3); Y = realloc(Y, sizeof(double)*(size_t)maxk*NCOL*3); cols = realloc(cols, sizeof(int)*(size_t)maxk*NCOL); }
      double V[9], Vi[9], gp[3];
      point_blocks(B, p, V, gp, W, maxk);
      if (!invert3(V, Vi)) { fail = 1; continue; }
      for (int i = 0; i < k; ++i) {
        const int64_t o = B->pt_obs[B->pt_start[p] + i];
        const int im = P->obs_img[o], cam = P->img_cam[im];
        int* c = cols + i*NCOL;
        for (int a = 0; a < 6; ++a) c[a] = 6*im + a;
        for (int a = 0; a < 9; ++a) c[6+a] = 6*n_img + 9*cam + a;
        const double* w = W + (size_t)i*NCOL*3; double* y = Y + (size_t)i*NCOL*3;
        for (int a = 0; a < NCOL; ++a) for (int b = 0; b < 3; ++b) y[3*a+b] = w[3*a]*Vi[b] + w[3*a+1]*Vi[3+b] + w[3*a+2]*Vi[6+b];
        /* U and J'r of this observation */
        const double* J = B->J + JSTR*(size_t)o; const double* r = B->r + 2*o;
        for (int a = 0; a < NCOL; ++a) {
          if (J[a] == 0.0 && J[18+a] == 0.0) continue;
          const double ga = J[a]*r[0] + J[18+a]*r[1] - (y[3*a]*gp[0] + y[3*a+1]*gp[1] + y[3*a+2]*gp[2]);
          #pragma omp atomic
          B->rhs[c[a]] += ga;
          for (int b = 0; b < NCOL; ++b) {
            if (c[a] < c[b]) continue;
            const double v = J[a]*J[b] + J[18+a]*J[18+b];
            if (v == 0.0) continue;
            double* dst = S_at(B, c[a], c[b]);
            #pragma omp atomic
            *dst += v;
          }
        }
      }
      for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) {
        const double* y = Y + (size_t)i*NCOL*3; const double* w = W + (size_t)j*NCOL*3;
        const int* ci = cols + i*NCOL; const int* cj = cols + j*NCOL;
        for (int a = 0; a < NCOL; ++a) {
          if (y[3*a] == 0.0 && y[3*a+1] == 0.0 && y[3*a+2] == 0.0) continue;
          for (int b = 0; b < NCOL; ++b) {
            if (ci[a] < cj[b]) continue;
            const double v = y[3*a]*w[3*b] + y[3*a+1]*w[3*b+1] + y[3*a+2]*w[3*b+2];
            if (v == 0.0) continue;
            double* dst = S_at(B, ci[a], cj[b]);
            #pragma omp atomic
            *dst -= v;
          }
        }
      }
    }
    free(W); free(Y); free(cols);
  }
  if (fail) return 0;
  for (int i = 0; i < n_img && B->pr_r; ++i) {      /* rotation-constraint rows touch one rvec block only */
    const double* j = B->pr_J + 3*(size_t)i;
    for (int a = 0; a < 3; ++a) { B->rhs[6*i + a] += j[a] * B->pr_r[i]; for (int b = 0; b <= a; ++b) *S_at(B, 6*i + a, 6*i + b) += j[a] * j[b]; }
  }
  for (int i = 0; i < nc; ++i) *S_at(B, i, i) += B->diag_c[i];
  { const char* dump = getenv("ORC_DUMP_S");     /* debugging aid: reduced system in skyline form */
    static int dump_count = 0;
    if (dump && ++dump_count == (getenv("ORC_DUMP_AT") ? atoi(getenv("ORC_DUMP_AT")) : 3)) {
      FILE* f = fopen(dump, "wb");
      if (f) { fwrite(&nc, sizeof(int), 1, f); fwrite(B->first, sizeof(int), (size_t)nc, f); fwrite(B->rowptr, sizeof(int64_t), (size_t)nc + 1, f);
               fwrite(B->S, sizeof(double), (size_t)B->rowptr[nc], f); fwrite(B->rhs, sizeof(double), (size_t)nc, f);
               fwrite(B->scale_c, sizeof(double), (size_t)nc, f); fwrite(B->poses, sizeof(double), 6*(size_t)n_img, f); fclose(f); } } }
  /* left-looking skyline Cholesky, rows stored contiguously */
  for (int j = 0; j < nc; ++j) {
    double* Lj = B->S + B->rowptr[j]; const int fj = B->first[j];
    double d = Lj[j - fj];
    for (int k = fj; k < j; ++k) d -= Lj[k - fj] * Lj[k - fj];
    if (!(d > 0.0) || !isfinite(d)) return 0;
    const double ljj = sqrt(d); Lj[j - fj] = ljj;
    const int hi = B->reach[j];
    #pragma omp parallel for schedule(static) if (hi - j > 256)
    for (int i = j + 1; i <= hi; ++i) {
      const int fi = B->first[i]; if (fi > j) continue;
      double* Li = B->S + B->rowptr[i];
      const int k0 = fi > fj ? fi : fj;
      double s = Li[j - fi];
      for (int k = k0; k < j; ++k) s -= Li[k - fi] * Lj[k - fj];
      Li[j - fi] = s / ljj;
    }
    for (int i = (hi + 1 > 6*n_img ? hi + 1 : 6*n_img); i < nc; ++i) {   /* intrinsics rows (first = 0) */
      if (i <= j) continue;
      double* Li = B->S + B->rowptr[i];
      double s = Li[j];
      for (int k = fj; k < j; ++k) s -= Li[k] * Lj[k - fj];
      Li[j] = s / ljj;
    }
  }
  return 1;
}

static void chol_solve(const ba_t* B, double* y) {      /* y <- (L L')^-1 y */
  const int nc = B->nc;
  for (int i = 0; i < nc; ++i) {
    const double* Li = B->S + B->rowptr[i]; const int fi = B->first[i];
    double s = y[i];
    for (int k = fi; k < i; ++k) s -= Li[k - fi] * y[k];
    y[i] = s / Li[i - fi];
  }
  for (int i = nc - 1; i >= 0; --i) {
    const double* Li = B->S + B->rowptr[i]; const int fi = B->first[i];
    y[i] /= Li[i - fi];
    for (int k = fi; k < i; ++k) y[k] -= Li[k - fi] * y[i];
  }
}

/* LevenbergMarquardtStrategy::ComputeStep.  Returns 1 on success; fills step_c/step_p
 * (already negated) and *model_cost_change. */
static int compute_step(ba_t* B, double radius, int reuse_diagonal, double* model_cost_change) {
  const mm_ba_problem* P = B->P; const mm_ba_options* O = B->O;
  const int nc = B->nc, n_img = B->n_img;
  if (!reuse_diagonal) {
    column_sums(B, 1, B->diag_c, B->diag_p);                         /* SquaredColumnNorm of scaled J */
    for (int i = 0; i < nc; ++i) B->diag_c[i] = fmin(fmax(B->diag_c[i], O->min_lm_diagonal), O->max_lm_diagonal);
    for (size_t i = 0; i < 3*(size_t)B->n_pt; ++i) B->diag_p[i] = fmin(fmax(B->diag_p[i], O->min_lm_diagonal), O->max_lm_diagonal);
  }
  /* diag_* hold clamp(diag); the system uses D^2 = clamp(diag)/radius */
  double* dc = malloc(sizeof(double)*(size_t)nc); double* dp = malloc(sizeof(double)*3*(size_t)B->n_pt);
  memcpy(dc, B->diag_c, sizeof(double)*(size_t)nc); memcpy(dp, B->diag_p, sizeof(double)*3*(size_t)B->n_pt);
  for (int i = 0; i < nc; ++i) B->diag_c[i] = dc[i] / radius;
  for (size_t i = 0; i < 3*(size_t)B->n_pt; ++i) B->diag_p[i] = dp[i] / radius;
  int ok = assemble_and_factor(B);
  if (ok) {
    memcpy(B->step_c, B->rhs, sizeof(double)*(size_t)nc);
    chol_solve(B, B->step_c);
    /* back-substitution: y_p = V^-1 (g_p - sum_i W_i' y_c) */
    #pragma omp parallel
    {
      int maxk = 64; double* W = malloc(sizeof(double)*(size_t)maxk*NCOL*3);
      #pragma omp for schedule(dynamic, 64)
      for (int p = 0; p < B->n_pt; ++p) {
        const int k = (int)(B->pt_start[p+1] - B->pt_start[p]);
        double* yp = B->step_p + 3*(size_t)p;
        if (k == 0) { yp[0] = yp[1] = yp[2] = 0.0; continue; }
        if (k > maxk) { maxk = 2*k; W = realloc(W, sizeof(double)*(size_t)maxk*NCOL*3); }
        double V[9], Vi[9], gp[3];
        point_blocks(B, p, V, gp, W, maxk);
        invert3(V, Vi);
        double t[3] = { gp[0], gp[1], gp[2] };
        for (int i = 0; i < k; ++i) {
          const int64_t o = B->pt_obs[B->pt_start[p] + i];
          const int im = P->obs_img[o], cam = P->img_cam[im];
          const double* w = W + (size_t)i*NCOL*3;
          for (int a = 0; a < NCOL; ++a) {
            const double yc = B->step_c[a < 6 ? 6*im + a : 6*n_img + 9*cam + (a - 6)];
            t[0] -= w[3*a]*yc; t[1] -= w[3*a+1]*yc; t[2] -= w[3*a+2]*yc;
          }
        }
        for (int a = 0; a < 3; ++a) yp[a] = Vi[3*a]*t[0] + Vi[3*a+1]*t[1] + Vi[3*a+2]*t[2];
      }
      free(W);
    }
    for (int i = 0; i < nc; ++i) { if (!isfinite(B->step_c[i])) ok = 0; B->step_c[i] = -B->step_c[i]; }
    for (size_t i = 0; i < 3*(size_t)B->n_pt; ++i) { if (!isfinite(B->step_p[i])) ok = 0; B->step_p[i] = -B->step_p[i]; }
  }
  memcpy(B->diag_c, dc, sizeof(double)*(size_t)nc); memcpy(B->diag_p, dp, sizeof(double)*3*(size_t)B->n_pt);
  free(dc); free(dp);
  if (!ok) return 0;
  /* model_cost_change = -(J step).(r + J step / 2)   (trust_region_minimizer.cc) */
  double mcc = 0.0;
  #pragma omp parallel for reduction(+:mcc) schedule(static)
  for (int64_t o = 0; o < B->n_obs; ++o) {
    const int im = P->obs_img[o], pt = P->obs_pt[o], cam = P->img_cam[im];
    const double* J = B->J + JSTR*(size_t)o;
    for (int row = 0; row < 2; ++row) {
      const double* j = J + 18*row; double m = 0.0;
      for (int k = 0; k < 6; ++k) m += j[k] * B->step_c[6*(size_t)im + k];
      for (int k = 0; k < 9; ++k) m += j[6+k] * B->step_c[6*(size_t)n_img + 9*(size_t)cam + k];
      for (int k = 0; k < 3; ++k) m += j[15+k] * B->step_p[3*(size_t)pt + k];
      mcc -= m * (B->r[2*o + row] + m / 2.0);
    }
  }
  for (int i = 0; i < n_img && B->pr_r; ++i) {
    double m = 0.0;
    for (int k = 0; k < 3; ++k) m += B->pr_J[3*(size_t)i + k] * B->step_c[6*(size_t)i + k];
    mcc -= m * (B->pr_r[i] + m / 2.0);
  }
  *model_cost_change = mcc;
  return 1;
}

static double active_norm(const ba_t* B, const double* poses, const double* intr, const double* pts) {
  double s = 0.0;
  for (int i = 0; i < B->n_img; ++i) for (int k = 0; k < 6; ++k) if (B->act_c[6*(size_t)i + k]) s += poses[6*(size_t)i+k]*poses[6*(size_t)i+k];
  for (int c = 0; c < B->n_cam; ++c) for (int k = 0; k < 9; ++k) if (B->act_c[6*(size_t)B->n_img + 9*(size_t)c + k]) s += intr[9*(size_t)c+k]*intr[9*(size_t)c+k];
  for (size_t i = 0; i < 3*(size_t)B->n_pt; ++i) if (B->act_p[i]) s += pts[i]*pts[i];
  return sqrt(s);
}

static void trace_push(mm_ba_summary* S, double cost, double radius, double gmax, int accepted, int lin) {
  const int i = S->num_iterations;
  if (i < MM_BA_TRACE_MAX) {
    S->trace_cost[i] = cost; S->trace_radius[i] = radius; S->trace_gradient_max_norm[i] = gmax;
    S->trace_accepted[i] = accepted; S->trace_linear_iterations[i] = lin;
  }
  S->num_iterations = i + 1;
  if (cost < S->final_cost) S->final_cost = cost;      /* SetSummaryFinalCost: min over iterations */
}

static double gradient_max_norm(ba_t* B) {
  column_sums(B, 0, B->g_c, B->g_p);
  double m = 0.0;
  for (int i = 0; i < B->nc; ++i) if (fabs(B->g_c[i]) > m) m = fabs(B->g_c[i]);
  for (size_t i = 0; i < 3*(size_t)B->n_pt; ++i) if (fabs(B->g_p[i]) > m) m = fabs(B->g_p[i]);
  return m;
}

static int validate(const mm_ba_problem* P) {
  if (!P || P->n_img < 0 || P->n_cam < 0 || P->n_pt < 0 || P->n_obs < 0) return MM_ERR_INVALID_ARG;
  if (P->n_obs > 0 && (!P->poses || !P->pose_const || !P->img_cam || !P->intr || !P->cam_model || !P->intr_const || !P->pts || !P->pt_const || !P->obs_xy || !P->obs_img || !P->obs_pt)) return MM_ERR_INVALID_ARG;
  for (int c = 0; c < P->n_cam; ++c) if (model_np(P->cam_model[c]) < 0) return MM_ERR_INVALID_ARG;
  for (int i = 0; i < P->n_img; ++i) if (P->img_cam[i] < 0 || P->img_cam[i] >= P->n_cam) return MM_ERR_INVALID_ARG;
  for (int64_t o = 0; o < P->n_obs; ++o) if (P->obs_img[o] < 0 || P->obs_img[o] >= P->n_img || P->obs_pt[o] < 0 || P->obs_pt[o] >= P->n_pt) return MM_ERR_INVALID_ARG;
  return MM_OK;
}

/* ORC_DETERMINISTIC=1 (what tests/conftest.py sets): every loop of the solve runs on one thread, so that all sums - the cost, and the
 * entries of the reduced system, which the threads otherwise accumulate with `omp atomic` in scheduling order - have ONE fixed
 * order and the oracle reproduces itself bit for bit on any machine.  The timed CPU arms of bench.py leave it unset. */
static void orc_apply_threading(void) {
#ifdef _OPENMP
  static int default_threads = 0;
  if (!default_threads) default_threads = omp_get_max_threads();
  const char* e = getenv("ORC_DETERMINISTIC");
  omp_set_num_threads(e && e[0] == '1' ? 1 : default_threads);
#endif
}

int orc_ba_solve(mm_ba_problem* P, const mm_ba_options* O, mm_ba_summary* S) {
  orc_apply_threading();
  mm_ba_summary local; if (!S) S = &local;
  memset(S, 0, sizeof *S);
  int rc = validate(P); if (rc != MM_OK) return rc;
  if (!O) return MM_ERR_INVALID_ARG;
  ba_t Bs; ba_t* B = &Bs; memset(B, 0, sizeof *B);
  B->P = P; B->O = O; B->n_img = P->n_img; B->n_cam = P->n_cam; B->n_pt = P->n_pt; B->n_obs = P->n_obs;
  B->nc = 6*P->n_img + 9*P->n_cam;
  const int nc = B->nc; const size_t np3 = 3*(size_t)P->n_pt;
  S->num_residuals = 2 * P->n_obs;
  S->final_cost = INFINITY;
  if (P->n_obs == 0) {            /* bundle_adjustment.cc:571-573: warning only; 0/0 -> NaN */
    S->termination = MM_TERM_EMPTY; S->initial_cost = S->final_cost = 0.0; S->return_value = NAN;
    return MM_OK;
  }
  /* point -> observation CSR (counting sort keeps the caller's order inside a point) */
  B->pt_start = calloc((size_t)P->n_pt + 2, sizeof(int64_t)); B->pt_obs = malloc(sizeof(int64_t)*(size_t)P->n_obs);
  for (int64_t o = 0; o < P->n_obs; ++o) B->pt_start[P->obs_pt[o] + 2]++;
  for (int p = 0; p < P->n_pt; ++p) B->pt_start[p+2] += B->pt_start[p+1];
  for (int64_t o = 0; o < P->n_obs; ++o) B->pt_obs[B->pt_start[P->obs_pt[o] + 1]++] = o;
  /* active columns */
  B->act_c = calloc((size_t)nc + 1, 1); B->act_p = calloc(np3 + 1, 1);
  { int* img_n = calloc((size_t)P->n_img + 1, sizeof(int)); int* cam_n = calloc((size_t)P->n_cam + 1, sizeof(int)); int* pt_n = calloc((size_t)P->n_pt + 1, sizeof(int));
    for (int64_t o = 0; o < P->n_obs; ++o) { img_n[P->obs_img[o]]++; cam_n[P->img_cam[P->obs_img[o]]]++; pt_n[P->obs_pt[o]]++; }
    for (int i = 0; i < P->n_img; ++i) if (img_n[i]) {
      for (int k = 0; k < 3; ++k) B->act_c[6*(size_t)i + k] = !P->pose_const[4*(size_t)i];
      for (int k = 0; k < 3; ++k) B->act_c[6*(size_t)i + 3 + k] = !P->pose_const[4*(size_t)i + 1 + k];
    }
    for (int c = 0; c < P->n_cam; ++c) if (cam_n[c] && !P->intr_const[c]) for (int k = 0; k < model_np(P->cam_model[c]); ++k) B->act_c[6*(size_t)P->n_img + 9*(size_t)c + k] = 1;
    for (int p = 0; p < P->n_pt; ++p) if (pt_n[p] && !P->pt_const[p]) B->act_p[3*(size_t)p] = B->act_p[3*(size_t)p+1] = B->act_p[3*(size_t)p+2] = 1;
    free(img_n); free(cam_n); free(pt_n); }
  B->poses = malloc(sizeof(double)*6*(size_t)P->n_img); B->poses2 = malloc(sizeof(double)*6*(size_t)P->n_img);
  B->intr = malloc(sizeof(double)*9*(size_t)P->n_cam);  B->intr2 = malloc(sizeof(double)*9*(size_t)P->n_cam);
  B->pts = malloc(sizeof(double)*np3 + 8); B->pts2 = malloc(sizeof(double)*np3 + 8);
  memcpy(B->poses, P->poses, sizeof(double)*6*(size_t)P->n_img);
  memcpy(B->intr, P->intr, sizeof(double)*9*(size_t)P->n_cam);
  memcpy(B->pts, P->pts, sizeof(double)*np3);
  B->J = malloc(sizeof(double)*JSTR*(size_t)P->n_obs); B->r = malloc(sizeof(double)*2*(size_t)P->n_obs);
  B->scale_c = malloc(sizeof(double)*(size_t)nc); B->scale_p = malloc(sizeof(double)*np3 + 8);
  B->diag_c = malloc(sizeof(double)*(size_t)nc);  B->diag_p = malloc(sizeof(double)*np3 + 8);
  B->g_c = malloc(sizeof(double)*(size_t)nc);     B->g_p = malloc(sizeof(double)*np3 + 8);
  B->step_c = malloc(sizeof(double)*(size_t)nc);  B->step_p = malloc(sizeof(double)*np3 + 8);
  for (int i = 0; i < P->n_img; ++i) if (has_prior(P, i)) B->n_prior++;
  if (B->n_prior) { B->pr_r = calloc((size_t)P->n_img, sizeof(double)); B->pr_J = calloc(3*(size_t)P->n_img, sizeof(double)); S->num_residuals += B->n_prior; }
  rc = build_structure(B);
  if (rc != MM_OK || !B->J) { rc = MM_ERR_ALLOC; goto done; }

  /* best iterate (x_min) lives in P's arrays; x in B->poses/intr/pts */
  double cost = eval_full(B);
  double gmax = gradient_max_norm(B);
  double radius = O->initial_trust_region_radius, decrease_factor = 2.0; int reuse_diagonal = 0;
  S->initial_cost = cost;
  trace_push(S, cost, radius, gmax, 1, 0);
  const double kEpsilon = 1e-12;
  const double abs_gtol = O->gradient_tolerance * fmax(gmax, kEpsilon);
  S->termination = MM_TERM_NO_CONVERGENCE;
  if (!isfinite(cost)) { S->termination = MM_TERM_NUMERICAL_FAILURE; rc = MM_ERR_NUMERICAL; goto finish; }
  if (gmax <= abs_gtol) { S->termination = MM_TERM_GRADIENT_TOLERANCE; goto finish; }
  if (O->jacobi_scaling) {
    column_sums(B, 1, B->scale_c, B->scale_p);
    for (int i = 0; i < nc; ++i) B->scale_c[i] = 1.0 / (1.0 + sqrt(B->scale_c[i]));
    for (size_t i = 0; i < np3; ++i) B->scale_p[i] = 1.0 / (1.0 + sqrt(B->scale_p[i]));
    scale_jacobian(B);
  } else {
    for (int i = 0; i < nc; ++i) B->scale_c[i] = 1.0;
    for (size_t i = 0; i < np3; ++i) B->scale_p[i] = 1.0;
  }
  double x_norm = active_norm(B, B->poses, B->intr, B->pts);
  int n_invalid = 0, iter = 0;
  for (;;) {
    if (iter >= O->max_num_iterations) { S->termination = MM_TERM_NO_CONVERGENCE; break; }
    ++iter;
    double mcc = 0.0; int valid = compute_step(B, radius, reuse_diagonal, &mcc);
    reuse_diagonal = 1;
    if (valid && mcc < 0.0) valid = 0;
    int successful = 0; double rel_dec = 0.0, new_cost = cost;
    if (!valid) {
      if (++n_invalid >= O->max_num_consecutive_invalid_steps) { S->termination = MM_TERM_NUMERICAL_FAILURE; break; }
    } else {
      n_invalid = 0;
      double sn = 0.0;
      for (int i = 0; i < P->n_img; ++i) for (int k = 0; k < 6; ++k) { const double d = B->step_c[6*(size_t)i+k] * B->scale_c[6*(size_t)i+k]; B->poses2[6*(size_t)i+k] = B->poses[6*(size_t)i+k] + d; sn += d*d; }
      for (int c = 0; c < P->n_cam; ++c) for (int k = 0; k < 9; ++k) { const size_t ci = 6*(size_t)P->n_img + 9*(size_t)c + k; const double d = B->step_c[ci] * B->scale_c[ci]; B->intr2[9*(size_t)c+k] = B->intr[9*(size_t)c+k] + d; sn += d*d; }
      for (size_t i = 0; i < np3; ++i) { const double d = B->step_p[i] * B->scale_p[i]; B->pts2[i] = B->pts[i] + d; sn += d*d; }
      new_cost = eval_cost(B, B->poses2, B->intr2, B->pts2);
      const double step_norm = sqrt(sn);
      if (step_norm <= O->parameter_tolerance * (x_norm + O->parameter_tolerance)) { S->termination = MM_TERM_PARAMETER_TOLERANCE; break; }
      const double cost_change = cost - new_cost;
      if (fabs(cost_change) < O->function_tolerance * cost) { S->termination = MM_TERM_FUNCTION_TOLERANCE; break; }
      rel_dec = cost_change / mcc;
      successful = rel_dec > O->min_relative_decrease;
    }
    if (successful) {
      S->num_successful_steps++;
      radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel_dec - 1.0, 3));    /* StepAccepted */
      radius = fmin(O->max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = 0;
      double* t;
      t = B->poses; B->poses = B->poses2; B->poses2 = t;
      t = B->intr; B->intr = B->intr2; B->intr2 = t;
      t = B->pts; B->pts = B->pts2; B->pts2 = t;
      x_norm = active_norm(B, B->poses, B->intr, B->pts);
      cost = eval_full(B);
      gmax = gradient_max_norm(B);
      if (gmax <= abs_gtol) { S->termination = MM_TERM_GRADIENT_TOLERANCE; break; }   /* before x_min = x (Ceres 1.8) */
      if (O->jacobi_scaling) scale_jacobian(B);
      memcpy(P->poses, B->poses, sizeof(double)*6*(size_t)P->n_img);      /* x_min = x (cost < minimum_cost) */
      memcpy(P->intr, B->intr, sizeof(double)*9*(size_t)P->n_cam);
      memcpy(P->pts, B->pts, sizeof(double)*np3);
    } else {
      S->num_unsuccessful_steps++;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;   /* StepRejected / StepIsInvalid */
    }
    if (radius < O->min_trust_region_radius) { S->termination = MM_TERM_PARAMETER_TOLERANCE; break; }
    trace_push(S, cost, radius, gmax, successful, 0);
    if (O->print_progress) printf("% 4d: f:% 8e new:% 8e g:% 3.2e rho:% 3.2e mu:% 3.2e ok:%d\n", iter, cost, new_cost, gmax, rel_dec, radius, successful);
  }
finish:
  S->return_value = sqrt(S->final_cost / (double)S->num_residuals);
  if (P->pt_err) {          /* bundle_adjustment.cc:575-598 at the returned parameters, raw residuals */
    for (int p = 0; p < P->n_pt; ++p) if (B->pt_start[p+1] > B->pt_start[p]) P->pt_err[p] = 0.0;
    for (int64_t o = 0; o < P->n_obs; ++o) {
      const int im = P->obs_img[o], pt = P->obs_pt[o], cam = P->img_cam[im]; double r[2];
      orc_ba_residual(P->cam_model[cam], P->poses + 6*(size_t)im, P->pts + 3*(size_t)pt, P->intr + 9*(size_t)cam, P->obs_xy + 2*o, r);
      P->pt_err[pt] += sqrt(r[0]*r[0] + r[1]*r[1]) / (double)(B->pt_start[pt+1] - B->pt_start[pt]);
    }
  }
done:
  free(B->pt_start); free(B->pt_obs); free(B->act_c); free(B->act_p);
  free(B->poses); free(B->poses2); free(B->intr); free(B->intr2); free(B->pts); free(B->pts2);
  free(B->J); free(B->r); free(B->scale_c); free(B->scale_p); free(B->diag_c); free(B->diag_p);
  free(B->g_c); free(B->g_p); free(B->step_c); free(B->step_p);
  free(B->rowptr); free(B->first); free(B->S); free(B->rhs); free(B->reach); free(B->pr_r); free(B->pr_J);
  return rc;
}

double orc_ba_cost(const mm_ba_problem* P, const mm_ba_options* O) {
  if (validate(P) != MM_OK || !O) return NAN;
  ba_t B; memset(&B, 0, sizeof B); B.P = P; B.O = O; B.n_obs = P->n_obs;
  return eval_cost(&B, P->poses, P->intr, P->pts);
}

int orc_pose_refine(double* rvec, double* tvec, int model_code, const double* params,
                    int64_t n, const double* points2D, const double* points3D,
                    const uint8_t* inlier_mask, const mm_ba_options* opt,
                    mm_ba_summary* summary, double* ret) {
  /* bundle_adjustment.cc:139-225: one pose free, points (:187) and intrinsics (:193) constant */
  if (!rvec || !tvec || !params || n < 0 || model_np(model_code) < 0 || !opt) return MM_ERR_INVALID_ARG;
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) if (!inlier_mask || inlier_mask[i]) ++m;
  double poses[6] = { rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2] };
  uint8_t pose_const[4] = { 0, 0, 0, 0 }; int32_t img_cam[1] = { 0 };
  double intr[9] = { 0 }; memcpy(intr, params, sizeof(double)*(size_t)model_np(model_code));
  int32_t cam_model[1] = { model_code }; uint8_t intr_const[1] = { 1 };
  double* pts = malloc(sizeof(double)*3*(size_t)(m + 1)); uint8_t* pt_const = malloc((size_t)m + 1);
  double* obs = malloc(sizeof(double)*2*(size_t)(m + 1)); int32_t* oi = calloc((size_t)m + 1, sizeof(int32_t)); int32_t* op = malloc(sizeof(int32_t)*(size_t)(m + 1));
  int64_t j = 0;
  for (int64_t i = 0; i < n; ++i) if (!inlier_mask || inlier_mask[i]) {
    memcpy(pts + 3*j, points3D + 3*i, 3*sizeof(double)); memcpy(obs + 2*j, points2D + 2*i, 2*sizeof(double));
    pt_const[j] = 1; op[j] = (int32_t)j; ++j;
  }
  mm_ba_problem P; memset(&P, 0, sizeof P);
  P.n_img = 1; P.n_cam = 1; P.n_pt = (int32_t)m; P.n_obs = m;
  P.poses = poses; P.pose_const = pose_const; P.img_cam = img_cam; P.intr = intr; P.cam_model = cam_model; P.intr_const = intr_const;
  P.pts = pts; P.pt_const = pt_const; P.obs_xy = obs; P.obs_img = oi; P.obs_pt = op;
  mm_ba_summary S; int rc = orc_ba_solve(&P, opt, &S);
  if (rc == MM_OK) { for (int k = 0; k < 3; ++k) { rvec[k] = poses[k]; tvec[k] = poses[3+k]; } if (ret) *ret = S.return_value; if (summary) *summary = S; }
  free(pts); free(pt_const); free(obs); free(oi); free(op);
  return rc;
}
