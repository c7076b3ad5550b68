// match.cuh — shared declarations of the matcher translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mm {

// knnMatch(k=2) result of one query row: indices -1 = absent (fewer than two candidates)
struct Knn2 { float d0, d1; int i0, i1; };

struct PairJob {
  int n1, n2;
  int64_t knn12_off, knn21_off;   // row offsets of this pair's top-2 lists
  int64_t out_off; int cap;       // output slot
  int64_t scr_off; int scr_cap;   // scratch slot (masked non-ratio path)
};

int knn2_simt(cudaStream_t st, const float* Q, int nq, const float* T, int nt, int K,
              const float* xyq, const float* xyt, double max_distance, Knn2* out, Knn2* part_scratch, int max_chunks);

// tensor-core candidate selection + exact re-rank for a list of pairs (match_tc.cu).
// Returns MM_ERR_UNSUPPORTED when the shapes/options are outside what the tcgen05 path handles.
int match_tc_pairs(const float* desc, const float* xy, int K, const int64_t* offs, int64_t total_rows, const int32_t* ia, const int32_t* ib,
                   const PairJob* jobs_host, int n_pairs, double max_distance, Knn2* knn12, Knn2* knn21,
                   cudaStream_t st, bool required);

// drop the TF32 operand copies cached for a descriptor array (call before the array is freed)
void match_tc_release(const float* desc);
// the descriptor array at `desc` was overwritten: refresh the cached copies on next use (buffers are kept)
void match_tc_invalidate(const float* desc);
// rows [r0[i], r0[i] + n[i]) of a prepared array were overwritten: redo their TF32 copies and norms (no-op when the array
// has not been prepared yet - the first tensor-core call prepares all of it)
int match_tc_refresh_rows(const float* desc, int K, const int64_t* r0, const int64_t* n, int count, cudaStream_t st);
void match_tc_stats(uint64_t* rows, uint64_t* flagged);

}  // namespace mm
