// match.cu — brute-force descriptor matching (K5) and its C ABI.
//
// Replaces match_brute_force() (mavmap/mavmap src/base2d/feature.cc:52-133): two
// cv::BFMatcher::knnMatch(k=2) passes (:71-77), Lowe ratio test (:12-20, :81-82), symmetric
// cross-check (:85-101), optional keypoint-distance mask (:23-49), and the non-ratio
// paths (:103-132).
//
// Distance semantics (identical to oracle/orc_match.c): sum_k (a_k - b_k)^2 accumulated in
// fp64 in ascending k with separate multiply and add, rounded once to fp32, then sqrtf; the
// two smallest per row are kept by strict '<' insertion in ascending index (OpenCV
// batchDistance), so ties resolve to the lower index.
//
// Two implementations produce the per-row top-2 lists:
//   MM_MATCH_IMPL_SIMT     exact CUDA-core kernel below (verification mode, also the fallback)
//   MM_MATCH_IMPL_TCGEN05  match_tc.cu: TF32 tcgen05 distance GEMM selects candidates, which are
//                          then re-ranked with the exact arithmetic above
#include <float.h>
#include <chrono>
#include <vector>
#include <algorithm>
#include <mutex>
#include "common.cuh"
#include "match.cuh"

namespace mm {

// ---------------------------------------------------------------- exact SIMT k-NN (k = 2)
// block: 128 query rows; grid.y: chunk of train rows.  Query tile and train tiles are staged in
// shared memory as fp64 (the fp32 -> fp64 conversion is exact), k-major for the queries.
constexpr int QT = 128;   // queries per block (= threads)
constexpr int TT = 32;    // train rows per staged tile

__device__ __forceinline__ void top2_insert(float d, int j, float& b0, int& i0, float& b1, int& i1) {
  if (d < b1) {
    if (b0 > d) { b1 = b0; i1 = i0; b0 = d; i0 = j; }
    else { b1 = d; i1 = j; }
  }
}

__global__ void __launch_bounds__(QT) k_knn2_simt(
    const float* __restrict__ Q, int nq, const float* __restrict__ T, int nt, int K,
    const float* __restrict__ xyq, const float* __restrict__ xyt, double max_dist2, int use_mask,
    int chunk, Knn2* __restrict__ part /* [gridDim.y][nq] */) {
  extern __shared__ double smem[];
  double* As = smem;                 // [K][QT]
  double* Bs = smem + (size_t)K * QT;  // [TT][K]
  const int q0 = blockIdx.x * QT, tid = threadIdx.x, q = q0 + tid;
  for (int e = tid; e < QT * K; e += QT) {
    const int r = e / K, k = e % K;
    As[(size_t)k * QT + r] = (q0 + r < nq) ? (double)Q[(size_t)(q0 + r) * K + k] : 0.0;
  }
  float b0 = FLT_MAX, b1 = FLT_MAX; int i0 = -1, i1 = -1;
  double qx = 0, qy = 0;
  if (use_mask && q < nq) { qx = (double)xyq[2 * (size_t)q]; qy = (double)xyq[2 * (size_t)q + 1]; }
  const int t_begin = blockIdx.y * chunk, t_end = min(nt, t_begin + chunk);
  for (int t0 = t_begin; t0 < t_end; t0 += TT) {
    __syncthreads();
    const int nt_tile = min(TT, t_end - t0);
    for (int e = tid; e < nt_tile * K; e += QT) Bs[e] = (double)T[(size_t)t0 * K + e];
    __syncthreads();
    if (q < nq) {
      for (int j = 0; j < nt_tile; ++j) {
        if (use_mask) {
          const double dx = qx - (double)xyt[2 * (size_t)(t0 + j)], dy = qy - (double)xyt[2 * (size_t)(t0 + j) + 1];
          if (!(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < max_dist2)) continue;     // feature.cc:40
        }
        const double* b = Bs + (size_t)j * K;
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
          const double t = As[(size_t)k * QT + tid] - b[k];
          s = __dadd_rn(s, __dmul_rn(t, t));
        }
        top2_insert(__fsqrt_rn((float)s), t0 + j, b0, i0, b1, i1);
      }
    }
  }
  if (q < nq) { Knn2 r; r.d0 = b0; r.d1 = b1; r.i0 = i0; r.i1 = i1; part[(size_t)blockIdx.y * nq + q] = r; }
}

__global__ void k_knn2_merge(int nq, int n_chunks, const Knn2* __restrict__ part, Knn2* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float b0 = FLT_MAX, b1 = FLT_MAX; int i0 = -1, i1 = -1;
  for (int c = 0; c < n_chunks; ++c) {
    const Knn2 r = part[(size_t)c * nq + q];
    if (r.i0 >= 0) top2_insert(r.d0, r.i0, b0, i0, b1, i1);
    if (r.i1 >= 0) top2_insert(r.d1, r.i1, b0, i0, b1, i1);
  }
  Knn2 r; r.d0 = b0; r.d1 = b1; r.i0 = i0; r.i1 = i1; out[q] = r;
}

// ---------------------------------------------------------------- ratio test + cross-check + ordered compaction
// one block per pair; rows are visited in ascending query index so the output order matches the
// reference's push_back order (feature.cc:85-101).
__device__ __forceinline__ int row_size(const Knn2& r, int ratio_test, double max_ratio) {
  int sz = (r.i0 >= 0) + (r.i1 >= 0);
  if (ratio_test && sz > 1 && (double)(r.d0 / r.d1) > max_ratio) sz = 0;     // feature.cc:15-17 (float division, double compare)
  return sz;
}

// One CTA of 1024 threads per pair: the passes below are chains of dependent gathers (a row of direction 1->2, then the row
// of 2->1 it points at) between block-wide barriers, so the time of a pair is the number of sweeps over its rows (256 threads:
// 36 us for 5000 rows, most of what a single-pair call waited for after the distance kernels).
__global__ void __launch_bounds__(1024) k_match_finalize(
    const PairJob* __restrict__ jobs, const Knn2* __restrict__ knn12_all, const Knn2* __restrict__ knn21_all,
    int ratio_test, double max_ratio, int mode /*0 ratio, 1 mutual-NN, 2 masked no-ratio*/,
    int32_t* __restrict__ q_out, int32_t* __restrict__ t_out, float* __restrict__ d_out, int32_t* __restrict__ cnt_out,
    int32_t* __restrict__ scratch /* mode 2: per-pair compacted matches21 train indices, stride = out stride */) {
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int n_warps = blockDim.x >> 5;
  const PairJob job = jobs[blockIdx.x];
  const Knn2* k12 = knn12_all + job.knn12_off;
  const Knn2* k21 = knn21_all + job.knn21_off;
  int32_t* qo = q_out + job.out_off; int32_t* to = t_out + job.out_off; float* dd = d_out + job.out_off;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int nc21 = 0;
  if (mode == 2) {
    // compacted matches21 (feature.cc:119-122): rows of direction 2->1 that have a candidate, in order
    int32_t* c21 = scratch + job.scr_off;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int j0 = 0; j0 < job.n2; j0 += blockDim.x) {
      const int j = j0 + threadIdx.x;
      const bool keep = j < job.n2 && k21[j].i0 >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) warp_cnt[w] = __popc(m);
      __syncthreads();
      int off = base;
      for (int x = 0; x < w; ++x) off += warp_cnt[x];
      // c21 can hold at most min(n1,n2) entries in the output slot; guard the rest
      if (keep) { const int pos = off + __popc(m & ((1u << lane) - 1)); if (pos < job.scr_cap) c21[pos] = k21[j].i0; }
      __syncthreads();
      if (threadIdx.x == 0) { int tot = 0; for (int x = 0; x < n_warps; ++x) tot += warp_cnt[x]; base += tot; }
      __syncthreads();
    }
    nc21 = base;
    __syncthreads();
  }
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < job.n1; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool keep = false; int tj = -1; float dist = 0.f;
    if (i < job.n1) {
      const Knn2 r = k12[i];
      if (mode == 0) {
        if (row_size(r, ratio_test, max_ratio) >= 2) {                 // feature.cc:87-89
          const Knn2 r2 = k21[r.i0];
          if (row_size(r2, ratio_test, max_ratio) >= 2 && r2.i0 == i) { keep = true; tj = r.i0; dist = r.d0; }   // :92-99
        }
      } else if (mode == 1) {
        if (r.i0 >= 0 && k21[r.i0].i0 == i) { keep = true; tj = r.i0; dist = r.d0; }   // crossCheck=true (:105-108)
      } else {
        if (r.i0 >= 0 && r.i0 < nc21 && r.i0 < job.scr_cap && (scratch + job.scr_off)[r.i0] == i) { keep = true; tj = r.i0; dist = r.d0; }   // :124-131
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[w] = __popc(m);
    __syncthreads();
    int off = base;
    for (int x = 0; x < w; ++x) off += warp_cnt[x];
    if (keep) { const int pos = off + __popc(m & ((1u << lane) - 1)); qo[pos] = i; to[pos] = tj; dd[pos] = dist; }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int x = 0; x < n_warps; ++x) tot += warp_cnt[x]; base += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) cnt_out[blockIdx.x] = base;
}

int knn2_simt(cudaStream_t st, const float* Q, int nq, const float* T, int nt, int K,
              const float* xyq, const float* xyt, double max_distance, Knn2* out, Knn2* part_scratch, int max_chunks) {
  if (nq <= 0) return MM_OK;
  const int use_mask = max_distance != -1.0;
  const int row_blocks = (nq + QT - 1) / QT;
  int n_chunks = std::max(1, std::min(max_chunks, (4 * num_sms() + row_blocks - 1) / row_blocks));
  int chunk = std::max(TT, ((nt + n_chunks - 1) / n_chunks + TT - 1) / TT * TT);
  n_chunks = std::max(1, (nt + chunk - 1) / chunk);
  const size_t smem = sizeof(double) * ((size_t)K * QT + (size_t)TT * K);
  MM_CUDA(cudaFuncSetAttribute(k_knn2_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      // per device / context: set on every call
  k_knn2_simt<<<dim3(row_blocks, n_chunks), QT, smem, st>>>(Q, nq, T, nt, K, xyq, xyt, max_distance * max_distance, use_mask, chunk, part_scratch);
  MM_LAUNCH_CHECK();
  k_knn2_merge<<<(nq + 127) / 128, 128, 0, st>>>(nq, n_chunks, part_scratch, out);
  MM_LAUNCH_CHECK();
  return MM_OK;
}

}  // namespace mm

using namespace mm;

struct mm_match_set {
  int n_images = 0, k = 0;
  std::vector<int> counts; std::vector<int64_t> offs;   // row offsets
  const float* desc = nullptr; const float* xy = nullptr;   // device
  DevBuf<float> own_desc, own_xy;
  int max_count = 0;
  // scratch reused across calls
  DevBuf<Knn2> knn12, knn21, part; DevBuf<PairJob> jobs; DevBuf<int32_t> scratch;
  size_t knn_cap = 0, jobs_cap = 0, part_cap = 0, scratch_cap = 0;
};

namespace {

constexpr int MAX_CHUNKS = 64;

int set_init(mm_match_set* s, const int32_t* counts, int n_images, int k) {
  s->n_images = n_images; s->k = k;
  s->counts.assign(counts, counts + n_images); s->offs.resize((size_t)n_images + 1); s->offs[0] = 0;
  for (int i = 0; i < n_images; ++i) {
    if (counts[i] < 0) { set_error("negative descriptor count"); return MM_ERR_INVALID_ARG; }
    s->offs[i + 1] = s->offs[i] + counts[i]; s->max_count = std::max(s->max_count, counts[i]);
  }
  return MM_OK;
}

// Match a list of pairs; results stay on the device (cnt/q/t/dist with a fixed per-pair stride).
int match_pairs_device(mm_match_set* s, const int32_t* ia, const int32_t* ib, int n_pairs, const mm_match_options* opt,
                       int32_t* cnt_dev, int32_t* q_dev, int32_t* t_dev, float* dist_dev, int stride, cudaStream_t st) {
  if (n_pairs == 0) return MM_OK;
  const bool use_mask = opt->max_distance != -1.0;
  if (use_mask && !s->xy) { set_error("max_distance >= 0 needs keypoint coordinates"); return MM_ERR_INVALID_ARG; }
  const int mode = opt->ratio_test ? 0 : (use_mask ? 2 : 1);
  std::vector<PairJob> jobs((size_t)n_pairs);
  size_t o12 = 0, o21 = 0;
  for (int p = 0; p < n_pairs; ++p) {
    if (ia[p] < 0 || ia[p] >= s->n_images || ib[p] < 0 || ib[p] >= s->n_images) { set_error("pair index out of range"); return MM_ERR_INVALID_ARG; }
    PairJob& j = jobs[p];
    j.n1 = s->counts[ia[p]]; j.n2 = s->counts[ib[p]];
    if (std::min(j.n1, j.n2) > stride) { set_error("output stride %d too small for pair %d (needs %d)", stride, p, std::min(j.n1, j.n2)); return MM_ERR_INVALID_ARG; }
    j.knn12_off = (int64_t)o12; j.knn21_off = (int64_t)o21; j.out_off = (int64_t)p * stride; j.cap = stride;
    j.scr_off = (int64_t)p * std::max(s->max_count, 1); j.scr_cap = std::max(s->max_count, 1);
    o12 += (size_t)j.n1; o21 += (size_t)j.n2;
  }
  if (o12 > s->knn_cap || o21 > s->knn_cap) { const size_t c = std::max(o12, o21); MM_CUDA(s->knn12.alloc(c)); MM_CUDA(s->knn21.alloc(c)); s->knn_cap = c; }
  if ((size_t)n_pairs > s->jobs_cap) { MM_CUDA(s->jobs.alloc((size_t)n_pairs)); s->jobs_cap = (size_t)n_pairs; }
  const size_t part_need = (size_t)MAX_CHUNKS * (size_t)std::max(s->max_count, 1);
  if (part_need > s->part_cap) { MM_CUDA(s->part.alloc(part_need)); s->part_cap = part_need; }
  if (mode == 2) { const size_t need = (size_t)n_pairs * (size_t)std::max(s->max_count, 1); if (need > s->scratch_cap) { MM_CUDA(s->scratch.alloc(need)); s->scratch_cap = need; } }
  MM_CUDA(cudaMemcpyAsync(s->jobs.p, jobs.data(), sizeof(PairJob) * (size_t)n_pairs, cudaMemcpyHostToDevice, st));
  const int impl = opt->impl;
  bool done_tc = false;
  if (impl == MM_MATCH_IMPL_TCGEN05 || impl == MM_MATCH_IMPL_AUTO) {
    int rc = match_tc_pairs(s->desc, s->xy, s->k, s->offs.data(), s->offs[s->n_images], ia, ib, jobs.data(), n_pairs, opt->max_distance,
                            s->knn12.p, s->knn21.p, st, impl == MM_MATCH_IMPL_TCGEN05);
    if (rc == MM_OK) done_tc = true;
    else if (rc != MM_ERR_UNSUPPORTED || impl == MM_MATCH_IMPL_TCGEN05) return rc;
  }
  if (!done_tc) {
    for (int p = 0; p < n_pairs; ++p) {
      const PairJob& j = jobs[p];
      const float* A = s->desc + s->offs[ia[p]] * s->k; const float* B = s->desc + s->offs[ib[p]] * s->k;
      const float* xa = s->xy ? s->xy + 2 * s->offs[ia[p]] : nullptr; const float* xb = s->xy ? s->xy + 2 * s->offs[ib[p]] : nullptr;
      int rc = knn2_simt(st, A, j.n1, B, j.n2, s->k, xa, xb, opt->max_distance, s->knn12.p + j.knn12_off, s->part.p, MAX_CHUNKS); if (rc) return rc;
      rc = knn2_simt(st, B, j.n2, A, j.n1, s->k, xb, xa, opt->max_distance, s->knn21.p + j.knn21_off, s->part.p, MAX_CHUNKS); if (rc) return rc;
    }
  }
  k_match_finalize<<<n_pairs, 1024, 0, st>>>(s->jobs.p, s->knn12.p, s->knn21.p, opt->ratio_test, opt->max_ratio, mode,
                                            q_dev, t_dev, dist_dev, cnt_dev, s->scratch.p);
  MM_LAUNCH_CHECK();
  return MM_OK;
}

}  // namespace

extern "C" {

void mm_match_options_default(mm_match_options* o) {
  if (!o) return;
  o->ratio_test = 1; o->max_ratio = 0.6; o->max_distance = -1.0; o->impl = MM_MATCH_IMPL_AUTO;    // feature.h:107-109
}

void mm_match_set_destroy(mm_match_set* s) { if (s) { match_tc_release(s->desc); delete s; } }

int mm_match_set_create_dev(const float* desc_dev, const float* xy_dev, const int32_t* counts, int32_t n_images, int32_t k, mm_match_set** out) {
  if (!out || !counts || n_images < 0 || k <= 0) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  *out = nullptr;
  int rc = ensure_device(); if (rc) return rc;
  mm_match_set* s = new mm_match_set();
  rc = set_init(s, counts, n_images, k); if (rc) { delete s; return rc; }
  s->desc = desc_dev; s->xy = xy_dev;
  *out = s;
  return MM_OK;
}

int mm_match_set_create(const float* desc, const float* xy, const int32_t* counts, int32_t n_images, int32_t k, mm_match_set** out) {
  if (!out || !counts || n_images < 0 || k <= 0) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  *out = nullptr;
  int rc = ensure_device(); if (rc) return rc;
  mm_match_set* s = new mm_match_set();
  rc = set_init(s, counts, n_images, k); if (rc) { delete s; return rc; }
  const size_t rows = (size_t)s->offs[n_images];
  if (rows > 0 && !desc) { delete s; set_error("null descriptors"); return MM_ERR_INVALID_ARG; }
  // +64 rows of zero padding so tiled loads past the last row stay in bounds
  if (s->own_desc.alloc((rows + 256) * (size_t)k) != cudaSuccess) { delete s; set_error("cudaMalloc failed"); cudaGetLastError(); return MM_ERR_ALLOC; }
  cudaMemset(s->own_desc.p, 0, sizeof(float) * (rows + 256) * (size_t)k);
  if (rows && cudaMemcpy(s->own_desc.p, desc, sizeof(float) * rows * (size_t)k, cudaMemcpyHostToDevice) != cudaSuccess) { delete s; set_error("descriptor upload failed"); return MM_ERR_CUDA; }
  s->desc = s->own_desc.p;
  if (xy) {
    if (s->own_xy.alloc(2 * rows + 2) != cudaSuccess) { delete s; set_error("cudaMalloc failed"); cudaGetLastError(); return MM_ERR_ALLOC; }
    if (rows && cudaMemcpy(s->own_xy.p, xy, sizeof(float) * 2 * rows, cudaMemcpyHostToDevice) != cudaSuccess) { delete s; set_error("keypoint upload failed"); return MM_ERR_CUDA; }
    s->xy = s->own_xy.p;
  }
  *out = s;
  return MM_OK;
}

int mm_match_set_pairs_dev(mm_match_set* s, const int32_t* ia, const int32_t* ib, int32_t n_pairs, const mm_match_options* opt,
                           int32_t* cnt_dev, int32_t* q_dev, int32_t* t_dev, float* dist_dev, int32_t stride, void* stream) {
  if (!s || !opt || n_pairs < 0 || (n_pairs > 0 && (!ia || !ib || !cnt_dev || !q_dev || !t_dev || !dist_dev))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  return match_pairs_device(s, ia, ib, n_pairs, opt, cnt_dev, q_dev, t_dev, dist_dev, stride, (cudaStream_t)stream);
}

int mm_match_set_pairs(mm_match_set* s, const int32_t* ia, const int32_t* ib, int32_t n_pairs, const mm_match_options* opt,
                       int64_t* match_off, int32_t* q, int32_t* t, float* dist, int64_t cap) {
  if (!s || !opt || n_pairs < 0 || !match_off || (n_pairs > 0 && (!ia || !ib))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  match_off[0] = 0;
  if (n_pairs == 0) return MM_OK;
  int stride = 1;
  for (int p = 0; p < n_pairs; ++p) {
    if (ia[p] < 0 || ia[p] >= s->n_images || ib[p] < 0 || ib[p] >= s->n_images) { set_error("pair index out of range"); return MM_ERR_INVALID_ARG; }
    stride = std::max(stride, std::min(s->counts[ia[p]], s->counts[ib[p]]));
  }
  // process in chunks bounded by a scratch budget
  const int64_t budget_rows = (int64_t)64 << 20;     // result slots per chunk
  const int chunk_pairs = (int)std::max<int64_t>(1, std::min<int64_t>(n_pairs, budget_rows / stride));
  DevBuf<int32_t> dq, dt, dc; DevBuf<float> dd;
  MM_CUDA(dq.alloc((size_t)chunk_pairs * stride)); MM_CUDA(dt.alloc((size_t)chunk_pairs * stride)); MM_CUDA(dd.alloc((size_t)chunk_pairs * stride)); MM_CUDA(dc.alloc((size_t)chunk_pairs));
  std::vector<int32_t> hc((size_t)chunk_pairs), hq((size_t)chunk_pairs * stride), ht((size_t)chunk_pairs * stride); std::vector<float> hd((size_t)chunk_pairs * stride);
  int64_t total = 0;
  for (int p0 = 0; p0 < n_pairs; p0 += chunk_pairs) {
    const int np = std::min(chunk_pairs, n_pairs - p0);
    int rc = match_pairs_device(s, ia + p0, ib + p0, np, opt, dc.p, dq.p, dt.p, dd.p, stride, nullptr); if (rc) return rc;
    MM_CUDA(cudaMemcpy(hc.data(), dc.p, sizeof(int32_t) * (size_t)np, cudaMemcpyDeviceToHost));
    // copy only the used prefix of each slot when few pairs, else the whole chunk
    MM_CUDA(cudaMemcpy(hq.data(), dq.p, sizeof(int32_t) * (size_t)np * stride, cudaMemcpyDeviceToHost));
    MM_CUDA(cudaMemcpy(ht.data(), dt.p, sizeof(int32_t) * (size_t)np * stride, cudaMemcpyDeviceToHost));
    MM_CUDA(cudaMemcpy(hd.data(), dd.p, sizeof(float) * (size_t)np * stride, cudaMemcpyDeviceToHost));
    for (int p = 0; p < np; ++p) {
      const int c = hc[p];
      if (total + c > cap) { set_error("output capacity too small"); return MM_ERR_INVALID_ARG; }
      if (c > 0 && (!q || !t || !dist)) { set_error("null output"); return MM_ERR_INVALID_ARG; }
      memcpy(q + total, hq.data() + (size_t)p * stride, sizeof(int32_t) * (size_t)c);
      memcpy(t + total, ht.data() + (size_t)p * stride, sizeof(int32_t) * (size_t)c);
      memcpy(dist + total, hd.data() + (size_t)p * stride, sizeof(float) * (size_t)c);
      total += c; match_off[p0 + p + 1] = total;
    }
  }
  return MM_OK;
}

// One-pair entry point.  A process-wide workspace keeps every device buffer alive between calls, and the descriptor
// arrays of the last PAIR_SLOTS images stay resident in fixed-size slots of one device array: the mapper matches image i
// against i-1 and i-2 (sequential_mapper.cc process()), so one of the two arrays of a call has normally been uploaded and
// prepared (TF32 operand copies, norms) by an earlier call.  A host array is recognised by (address, rows, k) and a hash
// of its content taken on every call - all of it up to 32 KB, above that the first and last row and 64 bytes of every 4 KB
// page; an array rewritten in place without touching any sampled byte would go unnoticed, MM_MATCH_PAIR_NO_CACHE=1 uploads
// always.
// The result comes back as one copy (count | q | t | dist) into pinned memory.
namespace {
constexpr int PAIR_SLOTS = 4;
struct PairSlot { const float* host = nullptr; int n = 0; uint64_t hash = 0, stamp = 0; };
struct PairWorkspace {
  mm_match_set set; DevBuf<float> desc, xy; DevBuf<int32_t> out; int32_t* host_out = nullptr;
  size_t slot_rows = 0, out_cap = 0; int k = 0; bool xy_alloc = false; uint64_t clock = 0; PairSlot slot[PAIR_SLOTS];
};
PairWorkspace* g_pair = nullptr;
std::mutex g_pair_mu;
uint64_t g_pair_calls = 0, g_pair_uploads = 0, g_pair_bytes = 0;       // under g_pair_mu

uint64_t sample_hash(const float* d, size_t n, size_t k) {
  const size_t row = k * sizeof(float), bytes = n * row;
  const unsigned char* b = reinterpret_cast<const unsigned char*>(d);
  uint64_t h = 0x9E3779B97F4A7C15ull ^ bytes;
  auto mix = [&h](const unsigned char* p, size_t len) {
    for (size_t i = 0; i < len; i += 8) { uint64_t w = 0; memcpy(&w, p + i, std::min<size_t>(8, len - i)); h = (h ^ w) * 0x100000001B3ull; h ^= h >> 29; }
  };
  if (bytes <= ((size_t)32 << 10)) { mix(b, bytes); return h; }        // small arrays: all of it
  mix(b, row); mix(b + bytes - row, row);
  // one 64-byte line of every 4 KB page, eight independent accumulators and a prefetch a dozen pages ahead: the sample is a
  // stream of cache misses, 11 us per 1.28 MB array this way against 35 us (64 bytes per KB: 90 us) as one dependent chain
  uint64_t acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  for (size_t o = 2048; o + 64 <= bytes; o += 4096) {
    if (o + 12 * 4096 < bytes) __builtin_prefetch(b + o + 12 * 4096);
    uint64_t wv[8]; memcpy(wv, b + o, 64);
    for (int j = 0; j < 8; ++j) acc[j] = (acc[j] ^ wv[j]) * 0x100000001B3ull + (acc[j] >> 31);
  }
  for (int j = 0; j < 8; ++j) { h = (h ^ acc[j]) * 0x100000001B3ull; h ^= h >> 29; }
  return h;
}

int pair_find(const PairWorkspace* w, const float* d, int n, uint64_t h) {
  for (int i = 0; i < PAIR_SLOTS; ++i) if (w->slot[i].host == d && w->slot[i].n == n && w->slot[i].hash == h) return i;
  return -1;
}
// copies `d` into the least recently used slot other than `keep`
int pair_upload(PairWorkspace* w, const float* d, int n, int k, uint64_t h, int keep, int* slot_out) {
  int victim = -1;
  for (int i = 0; i < PAIR_SLOTS; ++i) if (i != keep && (victim < 0 || w->slot[i].stamp < w->slot[victim].stamp)) victim = i;
  w->slot[victim] = PairSlot();                // not valid unless the copy has been issued
  MM_CUDA(cudaMemcpyAsync(w->desc.p + (size_t)victim * w->slot_rows * k, d, sizeof(float) * (size_t)n * k, cudaMemcpyHostToDevice, nullptr));
  w->slot[victim].host = d; w->slot[victim].n = n; w->slot[victim].hash = h;
  ++g_pair_uploads; g_pair_bytes += sizeof(float) * (size_t)n * k;
  *slot_out = victim;
  return MM_OK;
}
}  // namespace

void mm_match_pair_counters(uint64_t* calls, uint64_t* arrays_uploaded, uint64_t* bytes_uploaded) {
  std::lock_guard<std::mutex> lk(g_pair_mu);
  if (calls) *calls = g_pair_calls;
  if (arrays_uploaded) *arrays_uploaded = g_pair_uploads;
  if (bytes_uploaded) *bytes_uploaded = g_pair_bytes;
}

int mm_match_pair(const float* d1, int32_t n1, const float* d2, int32_t n2, int32_t k, const float* xy1, const float* xy2,
                  const mm_match_options* opt, int32_t* q, int32_t* t, float* dist, int32_t* n_out) {
  if (!opt || !n_out || n1 < 0 || n2 < 0 || k <= 0) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  *n_out = 0;
  int rc = ensure_device(); if (rc) return rc;
  if (n1 == 0 || n2 == 0) return MM_OK;          // matches.clear() (feature.cc:62) and nothing to add
  if (!d1 || !d2 || !q || !t || !dist) { set_error("null buffer"); return MM_ERR_INVALID_ARG; }
  if (opt->max_distance != -1.0 && (!xy1 || !xy2)) { set_error("max_distance >= 0 needs keypoints"); return MM_ERR_INVALID_ARG; }
  std::lock_guard<std::mutex> lk(g_pair_mu);
  if (!g_pair) g_pair = new PairWorkspace();
  PairWorkspace* w = g_pair;
  ++g_pair_calls;
  static const bool no_cache = getenv("MM_MATCH_PAIR_NO_CACHE") != nullptr;
  static const bool timing = getenv("MM_MATCH_PAIR_TIMING") != nullptr;      // host microseconds per phase of a call, to stderr
  double t_ph[6] = {0, 0, 0, 0, 0, 0}; int n_ph = 0;
  auto t_last = std::chrono::steady_clock::now();
  auto mark = [&](bool sync) {
    if (!timing) return;
    if (sync) cudaStreamSynchronize(nullptr);
    const auto now = std::chrono::steady_clock::now();
    t_ph[n_ph++] = std::chrono::duration<double, std::micro>(now - t_last).count(); t_last = now;
  };
  const size_t big = (size_t)std::max(n1, n2);
  if (big > w->slot_rows || k != w->k) {          // (re)size the slots: everything resident is dropped
    match_tc_release(w->desc.p);
    for (auto& sl : w->slot) sl = PairSlot();
    const size_t rows = std::max(w->slot_rows, (big + big / 4 + 255) / 256 * 256), total = (size_t)PAIR_SLOTS * rows + 256;
    w->slot_rows = 0; w->k = 0; w->xy_alloc = false;
    MM_CUDA(w->desc.alloc(total * (size_t)k));
    MM_CUDA(cudaMemsetAsync(w->desc.p, 0, sizeof(float) * total * (size_t)k, nullptr));
    w->slot_rows = rows; w->k = k;
  }
  const bool use_xy = xy1 && xy2;
  if (use_xy && !w->xy_alloc) { MM_CUDA(w->xy.alloc(2 * ((size_t)PAIR_SLOTS * w->slot_rows + 1))); w->xy_alloc = true; }
  const size_t cap = (size_t)std::min(n1, n2);
  if (cap > w->out_cap || !w->host_out) {
    const size_t c = 2 * cap;
    if (w->host_out) { cudaFreeHost(w->host_out); w->host_out = nullptr; w->out_cap = 0; }
    MM_CUDA(w->out.alloc(4 + 3 * c));
    MM_CUDA(cudaHostAlloc((void**)&w->host_out, sizeof(int32_t) * (4 + 3 * c), cudaHostAllocDefault));
    w->out_cap = c;
  }
  int sa = -1, sb = -1; bool up_a = false, up_b = false;
  {
    const bool same = d2 == d1 && n2 == n1;
    const bool share = same && !no_cache && !(use_xy && xy1 != xy2);      // one slot serves both sides (its keypoints too)
    const uint64_t h1 = sample_hash(d1, (size_t)n1, (size_t)k), h2 = same ? h1 : sample_hash(d2, (size_t)n2, (size_t)k);
    if (!no_cache) { sa = pair_find(w, d1, n1, h1); sb = same ? (share ? sa : -1) : pair_find(w, d2, n2, h2); }
    if (sa < 0) { rc = pair_upload(w, d1, n1, k, h1, sb, &sa); if (rc) return rc; up_a = true; if (share) sb = sa; }
    if (sb < 0) { rc = pair_upload(w, d2, n2, k, h2, sa, &sb); if (rc) return rc; up_b = true; }
    w->slot[sa].stamp = ++w->clock; w->slot[sb].stamp = ++w->clock;
  }
  mark(true);             // [0] workspace, content hashes, descriptor uploads
  if (use_xy) {           // the keypoints are small: copied on every call, not part of what identifies a slot
    MM_CUDA(cudaMemcpyAsync(w->xy.p + 2 * (size_t)sa * w->slot_rows, xy1, sizeof(float) * 2 * (size_t)n1, cudaMemcpyHostToDevice, nullptr));
    MM_CUDA(cudaMemcpyAsync(w->xy.p + 2 * (size_t)sb * w->slot_rows, xy2, sizeof(float) * 2 * (size_t)n2, cudaMemcpyHostToDevice, nullptr));
  }
  {                       // rows of the TF32 operand copies that changed (no-op until the tensor-core path has prepared the array)
    int64_t r0[2], nr[2]; int cnt = 0;
    if (up_a) { r0[cnt] = (int64_t)sa * (int64_t)w->slot_rows; nr[cnt] = n1; ++cnt; }
    if (up_b) { r0[cnt] = (int64_t)sb * (int64_t)w->slot_rows; nr[cnt] = n2; ++cnt; }
    rc = match_tc_refresh_rows(w->desc.p, k, r0, nr, cnt, nullptr); if (rc) return rc;
  }
  mark(true);             // [1] keypoints, TF32 operand copies and norms of the uploaded rows
  mm_match_set* s = &w->set;
  s->n_images = PAIR_SLOTS; s->k = k; s->desc = w->desc.p; s->xy = use_xy ? w->xy.p : nullptr;
  s->counts.assign(PAIR_SLOTS, 0); s->offs.resize(PAIR_SLOTS + 1); s->max_count = 0;
  for (int i = 0; i <= PAIR_SLOTS; ++i) s->offs[i] = (int64_t)i * (int64_t)w->slot_rows;
  for (int i = 0; i < PAIR_SLOTS; ++i) { s->counts[i] = w->slot[i].host ? w->slot[i].n : 0; s->max_count = std::max(s->max_count, s->counts[i]); }
  int32_t ia = sa, ib = sb;
  int32_t* o = w->out.p;
  rc = match_pairs_device(s, &ia, &ib, 1, opt, o, o + 4, o + 4 + cap, reinterpret_cast<float*>(o + 4 + 2 * cap), (int)cap, nullptr);
  if (rc) { for (auto& sl : w->slot) sl = PairSlot(); match_tc_invalidate(w->desc.p); return rc; }
  mark(false);            // [2] host side of the kernel launches
  mark(true);             // [3] the kernels
  MM_CUDA(cudaMemcpyAsync(w->host_out, o, sizeof(int32_t) * (4 + 3 * cap), cudaMemcpyDeviceToHost, nullptr));
  MM_CUDA(cudaStreamSynchronize(nullptr));
  const int32_t c = w->host_out[0];
  if (c < 0 || (size_t)c > cap) { set_error("internal: match count %d out of range", c); return MM_ERR_CUDA; }
  memcpy(q, w->host_out + 4, sizeof(int32_t) * (size_t)c);
  memcpy(t, w->host_out + 4 + cap, sizeof(int32_t) * (size_t)c);
  memcpy(dist, w->host_out + 4 + 2 * cap, sizeof(float) * (size_t)c);
  *n_out = c;
  mark(false);            // [4] result copy
  if (timing) fprintf(stderr, "[mm_match_pair] us: hash+upload %.1f (%d arrays)  prepare %.1f  launch %.1f  kernels %.1f  result %.1f\n",
                      t_ph[0], (int)up_a + (int)up_b, t_ph[1], t_ph[2], t_ph[3], t_ph[4]);
  return MM_OK;
}

}  // extern "C"
