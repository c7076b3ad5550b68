/*
 * oracle/orc_match.c — CPU restatement of match_brute_force, TEST INFRASTRUCTURE ONLY.
 *
 * Follows (mavmap/mavmap):
 *   ratio_test_          src/base2d/feature.cc:12-20
 *   max_distance_mask_   src/base2d/feature.cc:23-49
 *   match_brute_force    src/base2d/feature.cc:52-133
 * The k-NN search itself lives in the absent dependency OpenCV 2.4.7
 * (cv::BFMatcher::knnMatch / match, called at feature.cc:71-77,107-122).  Its published
 * algorithm (modules/core batchDistance): dist(i,j) = sqrt(sum_k (a_ik - b_jk)^2) in fp32,
 * K smallest kept by strict '<' insertion in ascending train index, so equal distances
 * keep the lower index first; rows start at FLT_MAX / index -1.
 *
 * Distance semantics pinned here: the sum of squares is accumulated exactly (double) and
 * rounded once to fp32 before sqrtf.  OpenCV's own SIMD summation order differs by CPU
 * dispatch (a few ulp), so this is the canonical value every order approximates; the
 * golden tests check that cv2.BFMatcher yields identical index lists on the seeded sets.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

static inline float l2_dist(const float* a, const float* b, int k) {
  double s = 0.0;
  for (int i = 0; i < k; ++i) { const double t = (double)a[i] - (double)b[i]; s += t * t; }
  return sqrtf((float)s);
}

static inline int mask_ok(const float* p, const float* q, double max_distance2) {
  /* feature.cc:33-40: double arithmetic on float coordinates, strict '<' */
  const double dx = (double)p[0] - (double)q[0], dy = (double)p[1] - (double)q[1];
  return dx * dx + dy * dy < max_distance2;
}

int orc_knn2(const float* d1, int32_t n1, const float* d2, int32_t n2, int32_t k,
             const float* xy1, const float* xy2, double max_distance,
             int32_t* idx, float* dist) {
  const int use_mask = max_distance != -1.0;
  const double md2 = max_distance * max_distance;
  if (use_mask && (!xy1 || !xy2)) return MM_ERR_INVALID_ARG;
  #pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < n1; ++i) {
    float b0 = FLT_MAX, b1 = FLT_MAX; int32_t i0 = -1, i1 = -1;
    const float* a = d1 + (size_t)i * k;
    for (int32_t j = 0; j < n2; ++j) {
      if (use_mask && !mask_ok(xy1 + 2*(size_t)i, xy2 + 2*(size_t)j, md2)) continue;
      const float d = l2_dist(a, d2 + (size_t)j * k, k);
      if (d < b1) {                      /* batchDistance insertion, K = 2 */
        if (b0 > d) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
        else        { b1 = d; i1 = j; }
      }
    }
    idx[2*(size_t)i] = i0; idx[2*(size_t)i+1] = i1;
    dist[2*(size_t)i] = b0; dist[2*(size_t)i+1] = b1;
  }
  return MM_OK;
}

int orc_match_pair(const float* d1, int32_t n1, const float* d2, int32_t n2, int32_t k,
                   const float* xy1, const float* xy2, const mm_match_options* opt,
                   int32_t* q, int32_t* t, float* dist, int32_t* n_out) {
  if (!opt || !n_out || n1 < 0 || n2 < 0 || k <= 0) return MM_ERR_INVALID_ARG;
  *n_out = 0;
  if (n1 == 0 || n2 == 0) return MM_OK;
  int32_t* i12 = malloc(sizeof(int32_t) * 2 * (size_t)n1);
  int32_t* i21 = malloc(sizeof(int32_t) * 2 * (size_t)n2);
  float* f12 = malloc(sizeof(float) * 2 * (size_t)n1);
  float* f21 = malloc(sizeof(float) * 2 * (size_t)n2);
  int rc = orc_knn2(d1, n1, d2, n2, k, xy1, xy2, opt->max_distance, i12, f12);
  if (rc == MM_OK) rc = orc_knn2(d2, n2, d1, n1, k, xy2, xy1, opt->max_distance, i21, f21);
  if (rc != MM_OK) { free(i12); free(i21); free(f12); free(f21); return rc; }
  int32_t m = 0;
  if (opt->ratio_test) {
    /* feature.cc:81-101. size() of a knn row = number of valid entries. */
    for (int32_t i = 0; i < n1; ++i) {
      int sz = (i12[2*i] >= 0) + (i12[2*i+1] >= 0);
      if (sz > 1 && (double)(f12[2*i] / f12[2*i+1]) > opt->max_ratio) sz = 0;   /* :15-17 */
      if (sz < 2) continue;                                                       /* :87-89 */
      const int32_t j = i12[2*i];
      int szj = (i21[2*j] >= 0) + (i21[2*j+1] >= 0);
      if (szj > 1 && (double)(f21[2*j] / f21[2*j+1]) > opt->max_ratio) szj = 0;
      if (szj < 2) continue;                                                      /* :92-94 */
      if (i21[2*j] == i) { q[m] = i; t[m] = j; dist[m] = f12[2*i]; ++m; }        /* :95-99 */
    }
  } else if (opt->max_distance == -1.0) {
    /* feature.cc:105-108: BFMatcher(norm, crossCheck=true).match == mutual nearest neighbour */
    for (int32_t i = 0; i < n1; ++i) {
      const int32_t j = i12[2*i];
      if (j >= 0 && i21[2*j] == i) { q[m] = i; t[m] = j; dist[m] = f12[2*i]; ++m; }
    }
  } else {
    /* feature.cc:110-131: match() drops rows without an allowed candidate, then the manual
     * cross-check indexes the COMPACTED matches21 by trainIdx (reference behaviour kept;
     * an out-of-range index is skipped instead of read). */
    int32_t* c21 = malloc(sizeof(int32_t) * (size_t)n2); int32_t nc21 = 0;
    for (int32_t j = 0; j < n2; ++j) if (i21[2*j] >= 0) c21[nc21++] = i21[2*j];
    for (int32_t i = 0; i < n1; ++i) {
      const int32_t j = i12[2*i];
      if (j < 0) continue;
      if (j < nc21 && c21[j] == i) { q[m] = i; t[m] = j; dist[m] = f12[2*i]; ++m; }
    }
    free(c21);
  }
  *n_out = m;
  free(i12); free(i21); free(f12); free(f21);
  return MM_OK;
}
