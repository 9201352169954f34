// spair_score.cuh -- SPair keypoint scoring shared by mv_k3_spair_errors (one pair) and mv_spair_match_batch.
//
// evaluate_spair_correspondence.py:83-98, :115-121: pred = arg-max pixel of each key point's heat map ->
// (col, row) / w; errors (K, K) = ||pred_k - kps_j[l, :2] / image_size|| / thresh_scale, 1e3 where
// kps_i[k, 2] * kps_j[l, 2] != 1; error_same = diagonal, (error_nn, index_nn) = row minimum, both only for the
// key points present in both images; PCK counts and the confusion matrix as integer counters.
#pragma once
#include <stdint.h>

struct SpairScoreShared {
  float err[64][65];
  unsigned int cnt[2];
};

// One CTA, any block size; K <= 64.  pred_flat may point to shared or global memory.  Ends with a barrier.
// NT == 0: every thread of the CTA takes part (__syncthreads); NT > 0: the first NT threads only, synchronised with named
// barrier 1 (the streaming SPair kernel keeps a producer warp out of it).
template <int NT = 0>
__device__ __forceinline__ void spair_score_sync() {
  if (NT == 0) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
}
// the error matrix as a view: err[k * pitch + l] (pitch > K, odd pitches are conflict-free), cnt[2]
struct SpairScoreView {
  float* err;
  int pitch;
  unsigned int* cnt;
  __device__ __forceinline__ float& at(int k, int l) const { return err[k * pitch + l]; }
};
template <int NT = 0>
__device__ __forceinline__ void spair_score_block(SpairScoreView sh, const int32_t* pred_flat, int K, int w,
                                                  const float* __restrict__ kps_i, const float* __restrict__ kps_j,
                                                  int stride, float image_size, float thresh_scale, float pck,
                                                  float* __restrict__ errors, float* __restrict__ error_same,
                                                  float* __restrict__ error_nn, int32_t* __restrict__ index_nn,
                                                  unsigned long long* __restrict__ hits,
                                                  unsigned long long* __restrict__ confusion, int conf_dim) {
  const int nthreads = NT == 0 ? (int)blockDim.x : NT;
  if (threadIdx.x < 2) sh.cnt[threadIdx.x] = 0;
  for (int t = threadIdx.x; t < K * K; t += nthreads) {
    const int k = t / K, l = t - k * K;
    const int flat = pred_flat[k];
    // argmax_2d -> (col, row); both divided by feats.shape[-1]  (spair:83)
    const float px = __fdiv_rn((float)(flat % w), (float)w), py = __fdiv_rn((float)(flat / w), (float)w);
    const float jx = __fdiv_rn(kps_j[(size_t)l * stride], image_size), jy = __fdiv_rn(kps_j[(size_t)l * stride + 1], image_size);
    const float dx = px - jx, dy = py - jy;
    float e = __fdiv_rn(sqrtf(fmaf(dy, dy, dx * dx)), thresh_scale);
    const bool valid = (kps_i[(size_t)k * stride + 2] * kps_j[(size_t)l * stride + 2]) == 1.f;
    if (!valid) e = 1e3f;
    sh.at(k, l) = e;
    if (errors) errors[t] = e;
  }
  spair_score_sync<NT>();
  for (int k = threadIdx.x; k < K; k += nthreads) {
    const bool in_both = (kps_i[(size_t)k * stride + 2] * kps_j[(size_t)k * stride + 2]) == 1.f;
    float es = -1.f, en = -1.f;
    int in = -1;
    if (in_both) {
      es = sh.at(k, k);
      en = sh.at(k, 0);
      in = 0;
      for (int l = 1; l < K; ++l)
        if (sh.at(k, l) < en) { en = sh.at(k, l); in = l; }
      atomicAdd(&sh.cnt[0], 1u);
      if (es < pck) atomicAdd(&sh.cnt[1], 1u);
      if (confusion) atomicAdd(&confusion[(size_t)k * conf_dim + in], 1ull);
    }
    if (error_same) error_same[k] = es;
    if (error_nn) error_nn[k] = en;
    if (index_nn) index_nn[k] = in;
  }
  spair_score_sync<NT>();
  if (hits && threadIdx.x < 2 && sh.cnt[threadIdx.x]) atomicAdd(&hits[threadIdx.x], (unsigned long long)sh.cnt[threadIdx.x]);
  spair_score_sync<NT>();
}

template <int NT = 0>
__device__ __forceinline__ void spair_score_block(SpairScoreShared& sh, const int32_t* pred_flat, int K, int w,
                                                  const float* __restrict__ kps_i, const float* __restrict__ kps_j,
                                                  int stride, float image_size, float thresh_scale, float pck,
                                                  float* __restrict__ errors, float* __restrict__ error_same,
                                                  float* __restrict__ error_nn, int32_t* __restrict__ index_nn,
                                                  unsigned long long* __restrict__ hits,
                                                  unsigned long long* __restrict__ confusion, int conf_dim) {
  spair_score_block<NT>(SpairScoreView{&sh.err[0][0], 65, sh.cnt}, pred_flat, K, w, kps_i, kps_j, stride, image_size,
                        thresh_scale, pck, errors, error_same, error_nn, index_nn, hits, confusion, conf_dim);
}
