/*
 * oracle/orc_camera.c — CPU restatement of the camera models and of the BA residual
 * functor, TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Follows (mavmap/mavmap):
 *   PinholeCameraModel::world2image / image2world   src/base3d/camera_models.h:111-145
 *   OpenCVCameraModel::world2image / image2world    src/base3d/camera_models.h:170-223
 *   OpenCVCameraModel::distortion                   src/base3d/camera_models.h:225-242
 *   CataCameraModel::world2image / image2world      src/base3d/camera_models.h:277-338
 *   camera_model_image2world (vector overload)      src/base3d/camera_models.cc:24-44
 *   BACostFunction<M>::operator()                   src/base3d/bundle_adjustment.h:131-159
 * and, for the part that lives in the absent dependency Ceres-Solver 1.8.0
 * (ceres/rotation.h AngleAxisRotatePoint, ceres/jet.h), its published algorithm:
 * Rodrigues' formula when theta^2 > 0, first-order pt + w x pt otherwise; derivatives by
 * forward-mode dual numbers ("Jets").
 */
#include <math.h>
#include <string.h>
#include "orc.h"

/* ------------------------------------------------------------------ jets */
#define JN 18
typedef struct { double a; double v[JN]; } jet;

static inline jet j_c(double a) { jet r; r.a = a; memset(r.v, 0, sizeof r.v); return r; }
static inline jet j_var(double a, int k) { jet r = j_c(a); r.v[k] = 1.0; return r; }
static inline jet j_add(jet x, jet y) { jet r; r.a = x.a + y.a; for (int i = 0; i < JN; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
static inline jet j_sub(jet x, jet y) { jet r; r.a = x.a - y.a; for (int i = 0; i < JN; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
static inline jet j_mul(jet x, jet y) { jet r; r.a = x.a * y.a; for (int i = 0; i < JN; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
static inline jet j_div(jet f, jet g) {
  /* ceres/jet.h: h = 1/g.a; (f.a*h, (f.v - f.a*h*g.v)*h) */
  jet r; const double h = 1.0 / g.a; const double fh = f.a * h; r.a = fh;
  for (int i = 0; i < JN; ++i) r.v[i] = (f.v[i] - fh * g.v[i]) * h; return r;
}
static inline jet j_sqrt(jet f) { jet r; const double t = sqrt(f.a); const double s = 1.0 / (2.0 * t); r.a = t; for (int i = 0; i < JN; ++i) r.v[i] = f.v[i] * s; return r; }
static inline jet j_sin(jet f) { jet r; const double c = cos(f.a); r.a = sin(f.a); for (int i = 0; i < JN; ++i) r.v[i] = c * f.v[i]; return r; }
static inline jet j_cos(jet f) { jet r; const double s = -sin(f.a); r.a = cos(f.a); for (int i = 0; i < JN; ++i) r.v[i] = s * f.v[i]; return r; }

/* --------------------------------------------------- scalar camera models */
static void distortion_d(const double* p, double u, double v, double* du, double* dv) {
  /* camera_models.h:225-242 */
  const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
  const double u2 = u * u, uy = u * v, y2 = v * v, r2 = u2 + y2;
  const double radial = k1 * r2 + k2 * r2 * r2;
  *du = u * radial + 2.0 * p1 * uy + p2 * (r2 + 2.0 * u2);
  *dv = v * radial + 2.0 * p2 * uy + p1 * (r2 + 2.0 * y2);
}

void orc_world2image(int model, const double* p, double x, double y, double z, double* u, double* v) {
  double uu, vv;
  if (model == MM_MODEL_PINHOLE) {            /* camera_models.h:111-130 */
    uu = x / z; vv = y / z;
  } else if (model == MM_MODEL_OPENCV) {      /* :170-193 */
    uu = x / z; vv = y / z;
    double du, dv; distortion_d(p, uu, vv, &du, &dv); uu += du; vv += dv;
  } else {                                    /* CATA :277-302 */
    const double zz = z + p[8] * sqrt(x * x + y * y + z * z);
    uu = x / zz; vv = y / zz;
    double du, dv; distortion_d(p, uu, vv, &du, &dv); uu += du; vv += dv;
  }
  *u = p[0] * uu + p[2];
  *v = p[1] * vv + p[3];
}

void orc_image2world(int model, const double* p, double u, double v, double* x, double* y, double* z) {
  double xx0 = (u - p[2]) / p[0];
  double yy0 = (v - p[3]) / p[1];
  if (model == MM_MODEL_PINHOLE) { *x = xx0; *y = yy0; *z = 1.0; return; }   /* :132-145 */
  /* recursive inverse distortion, 10 iterations (:210-217, :319-326) */
  double xx = xx0, yy = yy0, dx, dy;
  for (int i = 0; i < 10; ++i) { distortion_d(p, xx, yy, &dx, &dy); xx = xx0 - dx; yy = yy0 - dy; }
  *x = xx; *y = yy;
  if (model == MM_MODEL_OPENCV) { *z = 1.0; return; }
  const double xi = p[8];
  if (xi == 1.0) {                                                            /* :332 */
    *z = (1.0 - xx * xx - yy * yy) / 2.0;
  } else {
    const double r2 = xx * xx + yy * yy;
    *z = 1.0 - xi * (r2 + 1.0) / (xi + sqrt(1.0 + (1.0 - xi * xi) * r2));
  }
}

int orc_camera_world2image(int model, const double* p, int64_t n, const double* xyz, double* uv) {
  if (model < 1 || model > 3) return MM_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n; ++i) orc_world2image(model, p, xyz[3*i], xyz[3*i+1], xyz[3*i+2], &uv[2*i], &uv[2*i+1]);
  return MM_OK;
}
int orc_camera_image2world(int model, const double* p, int64_t n, const double* uv, double* xyz) {
  if (model < 1 || model > 3) return MM_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n; ++i) orc_image2world(model, p, uv[2*i], uv[2*i+1], &xyz[3*i], &xyz[3*i+1], &xyz[3*i+2]);
  return MM_OK;
}
int orc_camera_image2world_normalized(int model, const double* p, int64_t n, const double* uv, double* xy) {
  /* camera_models.cc:24-44: divide by z */
  if (model < 1 || model > 3) return MM_ERR_INVALID_ARG;
  for (int64_t i = 0; i < n; ++i) {
    double x, y, z; orc_image2world(model, p, uv[2*i], uv[2*i+1], &x, &y, &z);
    xy[2*i] = x / z; xy[2*i+1] = y / z;
  }
  return MM_OK;
}

/* ------------------------------------------- BA residual: doubles and jets */
static void rotate_point_d(const double* w, const double* pt, double* out) {
  /* ceres::AngleAxisRotatePoint (Ceres 1.8 rotation.h) */
  const double theta2 = w[0]*w[0] + w[1]*w[1] + w[2]*w[2];
  if (theta2 > 0.0) {
    const double theta = sqrt(theta2);
    const double k[3] = { w[0] / theta, w[1] / theta, w[2] / theta };
    const double c = cos(theta), s = sin(theta);
    const double kxp[3] = { k[1]*pt[2] - k[2]*pt[1], k[2]*pt[0] - k[0]*pt[2], k[0]*pt[1] - k[1]*pt[0] };
    const double kdp = k[0]*pt[0] + k[1]*pt[1] + k[2]*pt[2];
    for (int i = 0; i < 3; ++i) out[i] = pt[i] * c + kxp[i] * s + k[i] * (1.0 - c) * kdp;
  } else {
    const double wxp[3] = { w[1]*pt[2] - w[2]*pt[1], w[2]*pt[0] - w[0]*pt[2], w[0]*pt[1] - w[1]*pt[0] };
    for (int i = 0; i < 3; ++i) out[i] = pt[i] + wxp[i];
  }
}

void orc_ba_residual(int model, const double* pose, const double* X, const double* intr,
                     const double* obs, double* r) {
  double pc[3];
  rotate_point_d(pose, X, pc);                 /* bundle_adjustment.h:142 */
  pc[0] += pose[3]; pc[1] += pose[4]; pc[2] += pose[5];   /* :145-147 */
  double u, v;
  orc_world2image(model, intr, pc[0], pc[1], pc[2], &u, &v); /* :150 */
  r[0] = u - obs[0]; r[1] = v - obs[1];        /* :154-155 */
}

static void distortion_j(const jet* p, jet u, jet v, jet* du, jet* dv) {
  const jet k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
  const jet two = j_c(2.0);
  jet u2 = j_mul(u, u), uy = j_mul(u, v), y2 = j_mul(v, v), r2 = j_add(u2, y2);
  jet radial = j_add(j_mul(k1, r2), j_mul(j_mul(k2, r2), r2));
  *du = j_add(j_add(j_mul(u, radial), j_mul(j_mul(two, p1), uy)), j_mul(p2, j_add(r2, j_mul(two, u2))));
  *dv = j_add(j_add(j_mul(v, radial), j_mul(j_mul(two, p2), uy)), j_mul(p1, j_add(r2, j_mul(two, y2))));
}

void orc_ba_residual_jet(int model, const double* pose, const double* X, const double* intr,
                         const double* obs, double* r, double* J) {
  const int np = model == MM_MODEL_PINHOLE ? 4 : (model == MM_MODEL_OPENCV ? 8 : 9);
  jet w[3], t[3], pt[3], p[9];
  for (int i = 0; i < 3; ++i) { w[i] = j_var(pose[i], i); t[i] = j_var(pose[3+i], 3+i); pt[i] = j_var(X[i], 6+i); }
  for (int i = 0; i < 9; ++i) p[i] = i < np ? j_var(intr[i], 9+i) : j_c(0.0);

  jet pc[3];
  jet theta2 = j_add(j_add(j_mul(w[0], w[0]), j_mul(w[1], w[1])), j_mul(w[2], w[2]));
  if (theta2.a > 0.0) {
    jet theta = j_sqrt(theta2);
    jet k[3] = { j_div(w[0], theta), j_div(w[1], theta), j_div(w[2], theta) };
    jet c = j_cos(theta), s = j_sin(theta);
    jet kxp[3] = { j_sub(j_mul(k[1], pt[2]), j_mul(k[2], pt[1])),
                   j_sub(j_mul(k[2], pt[0]), j_mul(k[0], pt[2])),
                   j_sub(j_mul(k[0], pt[1]), j_mul(k[1], pt[0])) };
    jet kdp = j_add(j_add(j_mul(k[0], pt[0]), j_mul(k[1], pt[1])), j_mul(k[2], pt[2]));
    jet omc = j_sub(j_c(1.0), c);
    for (int i = 0; i < 3; ++i)
      pc[i] = j_add(j_add(j_mul(pt[i], c), j_mul(kxp[i], s)), j_mul(j_mul(k[i], omc), kdp));
  } else {
    jet wxp[3] = { j_sub(j_mul(w[1], pt[2]), j_mul(w[2], pt[1])),
                   j_sub(j_mul(w[2], pt[0]), j_mul(w[0], pt[2])),
                   j_sub(j_mul(w[0], pt[1]), j_mul(w[1], pt[0])) };
    for (int i = 0; i < 3; ++i) pc[i] = j_add(pt[i], wxp[i]);
  }
  for (int i = 0; i < 3; ++i) pc[i] = j_add(pc[i], t[i]);

  jet u, v;
  if (model == MM_MODEL_PINHOLE) {
    u = j_div(pc[0], pc[2]); v = j_div(pc[1], pc[2]);
  } else if (model == MM_MODEL_OPENCV) {
    u = j_div(pc[0], pc[2]); v = j_div(pc[1], pc[2]);
    jet du, dv; distortion_j(p, u, v, &du, &dv); u = j_add(u, du); v = j_add(v, dv);
  } else {
    jet n2 = j_add(j_add(j_mul(pc[0], pc[0]), j_mul(pc[1], pc[1])), j_mul(pc[2], pc[2]));
    jet zz = j_add(pc[2], j_mul(p[8], j_sqrt(n2)));
    u = j_div(pc[0], zz); v = j_div(pc[1], zz);
    jet du, dv; distortion_j(p, u, v, &du, &dv); u = j_add(u, du); v = j_add(v, dv);
  }
  u = j_add(j_mul(p[0], u), p[2]);
  v = j_add(j_mul(p[1], v), p[3]);
  r[0] = u.a - obs[0]; r[1] = v.a - obs[1];
  for (int i = 0; i < JN; ++i) { J[i] = u.v[i]; J[JN + i] = v.v[i]; }
}
