// camera.cuh — device (and host) camera-model functions with analytic Jacobians.
//
// Replaces the templated models of the reference (mavmap/mavmap):
//   PinholeCameraModel  src/base3d/camera_models.h:104-147   code 1, fx fy cx cy
//   OpenCVCameraModel   src/base3d/camera_models.h:163-244   code 2, + k1 k2 p1 p2
//   CataCameraModel     src/base3d/camera_models.h:270-359   code 3, + xi
// The reference differentiates these with Ceres Jets (bundle_adjustment.h:124-129); here the
// derivatives are written out by hand so the kernel carries no dual-number state.
#pragma once
#include <math.h>
#include "../../include/mavmap_b200.h"

#ifdef __CUDACC__
#define MM_HD __host__ __device__ __forceinline__
#else
#define MM_HD inline
#endif

namespace mm {

MM_HD int model_num_params(int model) {
  return model == MM_MODEL_PINHOLE ? 4 : (model == MM_MODEL_OPENCV ? 8 : (model == MM_MODEL_CATA ? 9 : -1));
}

// distortion (camera_models.h:225-242 / :340-357)
MM_HD void distortion(const double* p, double u, double v, double& du, double& dv) {
  const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
  const double u2 = u * u, uy = u * v, y2 = v * v, r2 = u2 + y2;
  const double radial = k1 * r2 + k2 * r2 * r2;
  du = u * radial + 2.0 * p1 * uy + p2 * (r2 + 2.0 * u2);
  dv = v * radial + 2.0 * p2 * uy + p1 * (r2 + 2.0 * y2);
}

// world2image. If WITH_J: dX[2][3] = d(u,v)/d(x,y,z), dP[2][9] = d(u,v)/d(params).
template <bool WITH_J>
MM_HD void world2image(int model, const double* p, double x, double y, double z,
                       double& u, double& v, double (*dX)[3], double (*dP)[9]) {
  const double fx = p[0], fy = p[1], cx = p[2], cy = p[3];
  double zz = z, rho = 0.0;
  if (model == MM_MODEL_CATA) { rho = sqrt(x * x + y * y + z * z); zz = z + p[8] * rho; }
  const double iz = 1.0 / zz;
  const double un = x / zz, vn = y / zz;          // same operation as the reference (x / z)
  double ud = un, vd = vn;
  // d(ud,vd)/d(un,vn)
  double a00 = 1.0, a01 = 0.0, a10 = 0.0, a11 = 1.0;
  double r2 = 0.0;
  if (model != MM_MODEL_PINHOLE) {
    const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
    const double u2 = un * un, uy = un * vn, y2 = vn * vn;
    r2 = u2 + y2;
    const double radial = k1 * r2 + k2 * r2 * r2;
    const double du = un * radial + 2.0 * p1 * uy + p2 * (r2 + 2.0 * u2);
    const double dv = vn * radial + 2.0 * p2 * uy + p1 * (r2 + 2.0 * y2);
    ud = un + du; vd = vn + dv;
    if (WITH_J) {
      const double drad = k1 + 2.0 * k2 * r2;
      a00 = 1.0 + radial + 2.0 * u2 * drad + 2.0 * p1 * vn + 6.0 * p2 * un;
      a01 = 2.0 * uy * drad + 2.0 * p1 * un + 2.0 * p2 * vn;
      a10 = 2.0 * uy * drad + 2.0 * p2 * vn + 2.0 * p1 * un;
      a11 = 1.0 + radial + 2.0 * y2 * drad + 2.0 * p2 * un + 6.0 * p1 * vn;
    }
  }
  u = fx * ud + cx;
  v = fy * vd + cy;
  if (WITH_J) {
    // d(un,vn)/d(x,y,z)
    double n00, n01, n02, n10, n11, n12;
    if (model == MM_MODEL_CATA) {
      const double xi = p[8], ir = 1.0 / rho;
      const double gx = xi * x * ir, gy = xi * y * ir, gz = 1.0 + xi * z * ir;   // d zz / d(x,y,z)
      n00 = iz - un * iz * gx; n01 = -un * iz * gy; n02 = -un * iz * gz;
      n10 = -vn * iz * gx; n11 = iz - vn * iz * gy; n12 = -vn * iz * gz;
    } else {
      n00 = iz; n01 = 0.0; n02 = -un * iz;
      n10 = 0.0; n11 = iz; n12 = -vn * iz;
    }
    dX[0][0] = fx * (a00 * n00 + a01 * n10); dX[0][1] = fx * (a00 * n01 + a01 * n11); dX[0][2] = fx * (a00 * n02 + a01 * n12);
    dX[1][0] = fy * (a10 * n00 + a11 * n10); dX[1][1] = fy * (a10 * n01 + a11 * n11); dX[1][2] = fy * (a10 * n02 + a11 * n12);
    if (dP) {
#pragma unroll
      for (int k = 0; k < 9; ++k) { dP[0][k] = 0.0; dP[1][k] = 0.0; }
      dP[0][0] = ud; dP[1][1] = vd; dP[0][2] = 1.0; dP[1][3] = 1.0;
      if (model != MM_MODEL_PINHOLE) {
        const double uy = un * vn;
        dP[0][4] = fx * un * r2;        dP[1][4] = fy * vn * r2;
        dP[0][5] = fx * un * r2 * r2;   dP[1][5] = fy * vn * r2 * r2;
        dP[0][6] = fx * 2.0 * uy;       dP[1][6] = fy * (r2 + 2.0 * vn * vn);
        dP[0][7] = fx * (r2 + 2.0 * un * un); dP[1][7] = fy * 2.0 * uy;
        if (model == MM_MODEL_CATA) {
          const double dun = -un * iz * rho, dvn = -vn * iz * rho;     // d(un,vn)/d xi
          dP[0][8] = fx * (a00 * dun + a01 * dvn);
          dP[1][8] = fy * (a10 * dun + a11 * dvn);
        }
      }
    }
  }
}

// image2world (camera_models.h:132-145, :195-223, :304-338)
MM_HD void image2world(int model, const double* p, double u, double v, double& x, double& y, double& z) {
  const double x0 = (u - p[2]) / p[0], y0 = (v - p[3]) / p[1];
  if (model == MM_MODEL_PINHOLE) { x = x0; y = y0; z = 1.0; return; }
  double xx = x0, yy = y0, dx, dy;
#pragma unroll 1
  for (int i = 0; i < 10; ++i) { distortion(p, xx, yy, dx, dy); xx = x0 - dx; yy = y0 - dy; }
  x = xx; y = yy;
  if (model == MM_MODEL_OPENCV) { z = 1.0; return; }
  const double xi = p[8];
  if (xi == 1.0) { z = (1.0 - xx * xx - yy * yy) / 2.0; }
  else { const double r2 = xx * xx + yy * yy; z = 1.0 - xi * (r2 + 1.0) / (xi + sqrt(1.0 + (1.0 - xi * xi) * r2)); }
}

// Rotation of a point by an angle-axis vector exactly as ceres::AngleAxisRotatePoint does
// (Rodrigues for theta^2 > 0, first-order otherwise); used for residual-only evaluation.
MM_HD void angle_axis_rotate(const double* w, const double* pt, double* out) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > 0.0) {
    const double theta = sqrt(theta2);
    const double k0 = w[0] / theta, k1 = w[1] / theta, k2 = w[2] / theta;
    const double c = cos(theta), s = sin(theta);
    const double c0 = k1 * pt[2] - k2 * pt[1], c1 = k2 * pt[0] - k0 * pt[2], c2 = k0 * pt[1] - k1 * pt[0];
    const double kdp = k0 * pt[0] + k1 * pt[1] + k2 * pt[2];
    out[0] = pt[0] * c + c0 * s + k0 * (1.0 - c) * kdp;
    out[1] = pt[1] * c + c1 * s + k1 * (1.0 - c) * kdp;
    out[2] = pt[2] * c + c2 * s + k2 * (1.0 - c) * kdp;
  } else {
    out[0] = pt[0] + (w[1] * pt[2] - w[2] * pt[1]);
    out[1] = pt[1] + (w[2] * pt[0] - w[0] * pt[2]);
    out[2] = pt[2] + (w[0] * pt[1] - w[1] * pt[0]);
  }
}

// Per-image rotation data: R (row-major 9) and the left Jacobian Jl (9) of SO(3), so that
// d(R X)/d w = -[R X]x Jl  (equal to what autodiff of AngleAxisRotatePoint yields).
MM_HD void rotation_and_left_jacobian(const double* w, double* R, double* Jl) {
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double A, B, Cc, Dd;     // sin t / t, (1-cos t)/t^2, (t - sin t)/t^3 ; A for R, Cc/Dd for Jl
  if (t2 > 0.0) {
    const double t = sqrt(t2);
    const double sh = sin(0.5 * t);
    A = sin(t) / t;
    B = 2.0 * sh * sh / t2;                       // (1 - cos t)/t^2 without cancellation
    if (t < 0.05) Dd = 1.0 / 6.0 - t2 / 120.0 + t2 * t2 / 5040.0 - t2 * t2 * t2 / 362880.0;
    else Dd = (t - sin(t)) / (t2 * t);
    Cc = B;
    // R = I + A [w]x + B [w]x^2
    const double wx = w[0], wy = w[1], wz = w[2];
    R[0] = 1.0 - B * (wy * wy + wz * wz); R[1] = -A * wz + B * wx * wy;       R[2] = A * wy + B * wx * wz;
    R[3] = A * wz + B * wx * wy;          R[4] = 1.0 - B * (wx * wx + wz * wz); R[5] = -A * wx + B * wy * wz;
    R[6] = -A * wy + B * wx * wz;         R[7] = A * wx + B * wy * wz;        R[8] = 1.0 - B * (wx * wx + wy * wy);
    // Jl = I + Cc [w]x + Dd [w]x^2
    Jl[0] = 1.0 - Dd * (wy * wy + wz * wz); Jl[1] = -Cc * wz + Dd * wx * wy;       Jl[2] = Cc * wy + Dd * wx * wz;
    Jl[3] = Cc * wz + Dd * wx * wy;         Jl[4] = 1.0 - Dd * (wx * wx + wz * wz); Jl[5] = -Cc * wx + Dd * wy * wz;
    Jl[6] = -Cc * wy + Dd * wx * wz;        Jl[7] = Cc * wx + Dd * wy * wz;        Jl[8] = 1.0 - Dd * (wx * wx + wy * wy);
  } else {
    // ceres first-order branch: R x = x + w x x (w == 0 here), derivative -[x]x
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    Jl[0] = 1; Jl[1] = 0; Jl[2] = 0; Jl[3] = 0; Jl[4] = 1; Jl[5] = 0; Jl[6] = 0; Jl[7] = 0; Jl[8] = 1;
  }
}

}  // namespace mm
