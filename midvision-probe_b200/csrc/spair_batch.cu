// spair_batch.cu -- SPair-71k keypoint transfer for a BATCH of image pairs in one launch.
//
//   mv_spair_match_batch   per pair: bilinear key-point features of the normalised map, K x (h*w) heat map with
//                          the per-pixel normalisation folded in, arg-max, key-point error matrix, PCK counts
//                          -- one CTA per pair, fp32 throughout
//
// Reference behaviour being reproduced (file:line in /root/reference):
//   evaluate_spair_correspondence.py:59     feats = F.normalize(feats, p=2, dim=1)
//   evaluate_spair_correspondence.py:71-79  key points / image size -> NDC -> grid_sample(bilinear, align_corners=True)
//   evaluate_spair_correspondence.py:82-83  heatmaps = einsum("k f, f h w -> k h w"); argmax_2d(...) / w
//   evaluate_spair_correspondence.py:86-98  error matrix, validity, error_same / error_nn        (spair_score.cuh)
//   evaluate_spair_correspondence.py:108, :115-121  loop over pairs, confusion matrix, recall
//
// Why a separate kernel: one pair is 2 * K * h*w * C = 6 MFLOP (K = 20, 14 x 14, C = 768) -- a 128-row
// tensor-core tile would be 84 % padding and the ten launches of the per-pair path (kernels 1-3) are pure
// launch latency.  Here each feature map is read once (1.2 MB per pair, the HBM floor) straight from the
// backbone's (B, 2, C, h, w) output, the heat map lives in registers and nothing but K integers per pair is
// written, so the path is bounded by HBM, not by launches.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "spair_score.cuh"

namespace {

constexpr int SPB_THREADS = 256;
constexpr int SPB_MAX_HW = 1 << 20;  // pixels per feature map (2500 at the reference's 800 x 800 input)
constexpr float SPB_NORM_EPS = 1e-12f;  // F.normalize default eps

struct SpairBatchParams {
  const float* feats;  // (B, 2, C, h, w)
  const float* kps_i;  // (B, K, stride)
  const float* kps_j;
  const float* thresh_scale;  // (B)
  int B, C, h, w, K, stride;
  float image_size, pck;
  int32_t* pred;  // (B, K)
  float* error_same;
  float* error_nn;
  int32_t* index_nn;
  unsigned long long* hits;
  unsigned long long* confusion;
  int conf_dim;
};

// dynamic shared memory: q[C][KT], the key-point features of the current tile
// blockDim.x = min(SPB_THREADS, h*w rounded up to a warp): thread = pixel in the heat-map pass, so a 14 x 14 map
// runs 7 warps with 87 % of the lanes busy instead of 8 warps with 77 %.
template <int KT>
__global__ void __launch_bounds__(SPB_THREADS) spair_batch_kernel(SpairBatchParams p) {
  extern __shared__ float4 spb_dyn[];
  __shared__ SpairScoreShared score;
  __shared__ int s_pred[64];
  __shared__ float s_wt[64][4];   // bilinear weight / max(||f_i[tap]||, eps); 0 for a tap outside the map (zero padding)
  __shared__ int s_tap[64][4];    // pixel index of the tap (clamped into the map)
  __shared__ float s_ss[64][4];   // sum of squares of f_i at the tap
  __shared__ float s_bv[SPB_THREADS / 32][KT];
  __shared__ int s_bi[SPB_THREADS / 32][KT];
  const int C = p.C, hw = p.h * p.w, K = p.K;
  float* q = reinterpret_cast<float*>(spb_dyn);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nthr = blockDim.x, nwarp = nthr >> 5;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* fi = p.feats + (size_t)b * 2 * C * hw;
    const float* fj = fi + (size_t)C * hw;
    const float* ki = p.kps_i + (size_t)b * K * p.stride;
    const float* kj = p.kps_j + (size_t)b * K * p.stride;

    // ---- key-point taps: kp / size * 2 - 1 -> ((g + 1) / 2) * (size - 1)  (align_corners=True) ----
    if (tid < K) {
      const float kx = __fdiv_rn(ki[(size_t)tid * p.stride + 0], p.image_size);
      const float ky = __fdiv_rn(ki[(size_t)tid * p.stride + 1], p.image_size);
      const float gx = __fsub_rn(__fmul_rn(kx, 2.f), 1.f), gy = __fsub_rn(__fmul_rn(ky, 2.f), 1.f);
      const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(p.w - 1));
      const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(p.h - 1));
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float ww = ix - fx, we = 1.f - ww, wn = iy - fy, ws = 1.f - wn;
      const float wt[4] = {ws * we, ws * ww, wn * we, wn * ww};  // nw, ne, sw, se
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        const bool in = xx >= 0 && xx < p.w && yy >= 0 && yy < p.h;
        s_wt[tid][t] = in ? wt[t] : 0.f;
        s_tap[tid][t] = min(max(yy, 0), p.h - 1) * p.w + min(max(xx, 0), p.w - 1);
        s_ss[tid][t] = 0.f;
      }
    }
    __syncthreads();
    // ---- ||f_i|| at the 4K tap pixels only: (tap, channel slice) per thread, 16 independent loads in flight ----
    {
      const int ntap = 4 * K;
      const int slices = max(1, nthr / ntap);
      for (int item = tid; item < ntap * slices; item += nthr) {
        const int tp = item % ntap, sl = item / ntap;
        const float* src = fi + s_tap[tp >> 2][tp & 3];
        float ss = 0.f;
#pragma unroll 16
        for (int c = sl; c < C; c += slices) {
          const float v = __ldg(src + (size_t)c * hw);
          ss = fmaf(v, v, ss);
        }
        atomicAdd(&s_ss[tp >> 2][tp & 3], ss);
      }
    }
    __syncthreads();
    if (tid < 4 * K) {  // fold 1 / max(||f_i[tap]||, eps) into the blend weight
      const int k = tid >> 2, t = tid & 3;
      s_wt[k][t] = __fdiv_rn(s_wt[k][t], fmaxf(sqrtf(s_ss[k][t]), SPB_NORM_EPS));
    }
    __syncthreads();

    for (int k0 = 0; k0 < K; k0 += KT) {
      const int kt = min(KT, K - k0);
      // ---- q[c][k] = sum_t wt * f_i[c][tap] / max(||f_i[tap]||, eps): the grid_sample of the normalised map ----
#pragma unroll 4
      for (int idx = tid; idx < C * KT; idx += nthr) {
        const int c = idx / KT, k = idx - c * KT;
        const int kk = k0 + min(k, kt - 1);
        const float* src = fi + (size_t)c * hw;
        float acc = __ldg(src + s_tap[kk][0]) * s_wt[kk][0];
        acc = fmaf(__ldg(src + s_tap[kk][1]), s_wt[kk][1], acc);
        acc = fmaf(__ldg(src + s_tap[kk][2]), s_wt[kk][2], acc);
        acc = fmaf(__ldg(src + s_tap[kk][3]), s_wt[kk][3], acc);
        q[idx] = (k < kt) ? acc : 0.f;
      }
      __syncthreads();

      // ---- heat[k][px] = (sum_c q[c][k] * f_j[c][px]) / max(||f_j[px]||, eps), f_j read ONCE: the norm is
      //      accumulated in the same pass; running arg-max per thread ----
      float bestv[KT];
      int besti[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        bestv[k] = -CUDART_INF_F;
        besti[k] = 0x7fffffff;
      }
      for (int px = tid; px < hw; px += nthr) {
        float acc[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] = 0.f;
        float ss = 0.f;
        const float* col = fj + px;
        int c0 = 0;
        for (; c0 + 16 <= C; c0 += 16) {
          float v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = __ldg(col + (size_t)(c0 + u) * hw);  // 16 loads in flight
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            ss = fmaf(v[u], v[u], ss);
            const float4* qc = reinterpret_cast<const float4*>(q + (size_t)(c0 + u) * KT);
#pragma unroll
            for (int k4 = 0; k4 < KT / 4; ++k4) {
              const float4 qq = qc[k4];  // same address in every lane: broadcast
              acc[4 * k4 + 0] = fmaf(qq.x, v[u], acc[4 * k4 + 0]);
              acc[4 * k4 + 1] = fmaf(qq.y, v[u], acc[4 * k4 + 1]);
              acc[4 * k4 + 2] = fmaf(qq.z, v[u], acc[4 * k4 + 2]);
              acc[4 * k4 + 3] = fmaf(qq.w, v[u], acc[4 * k4 + 3]);
            }
          }
        }
        for (; c0 < C; ++c0) {
          const float v = __ldg(col + (size_t)c0 * hw);
          ss = fmaf(v, v, ss);
#pragma unroll
          for (int k = 0; k < KT; ++k) acc[k] = fmaf(q[(size_t)c0 * KT + k], v, acc[k]);
        }
        const float nrm = fmaxf(sqrtf(ss), SPB_NORM_EPS);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const float hv = __fdiv_rn(acc[k], nrm);
          if (hv > bestv[k]) {  // pixels are visited in ascending order: the first maximum wins
            bestv[k] = hv;
            besti[k] = px;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        float v = bestv[k];
        int i = besti[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, i, o);
          if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
          }
        }
        if (lane == 0) {
          s_bv[wid][k] = v;
          s_bi[wid][k] = i;
        }
      }
      __syncthreads();
      if (tid < kt) {
        float v = s_bv[0][tid];
        int i = s_bi[0][tid];
        for (int wq = 1; wq < nwarp; ++wq) {
          const float ov = s_bv[wq][tid];
          const int oi = s_bi[wq][tid];
          if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
          }
        }
        i = (i == 0x7fffffff) ? 0 : i;  // an all-NaN heat map: torch.argmax would return a NaN position; we return 0
        s_pred[k0 + tid] = i;
        if (p.pred) p.pred[(size_t)b * K + k0 + tid] = i;
      }
      __syncthreads();  // q, s_bv are rewritten by the next tile
    }

    spair_score_block(score, s_pred, K, p.w, ki, kj, p.stride, p.image_size, __ldg(p.thresh_scale + b), p.pck, nullptr,
                      p.error_same ? p.error_same + (size_t)b * K : nullptr,
                      p.error_nn ? p.error_nn + (size_t)b * K : nullptr,
                      p.index_nn ? p.index_nn + (size_t)b * K : nullptr, p.hits, p.confusion, p.conf_dim);
  }
}


// ------------------------------------------------------------------------------------------
// the streaming form (the default where it applies: h*w <= 256, C % 8 == 0, 16-byte aligned maps)
// ------------------------------------------------------------------------------------------
// A pair's two maps are ONE contiguous 2 * C * h*w * 4 byte block, read three times in channel order: image i for the
// pixel norms, image i again (from L2) for the key-point features q, image j for the heat map.  A producer warp streams
// that sequence through a ring of shared-memory stages with 1-D bulk copies (cp.async.bulk, 8 channels = 6.3 KB per
// stage for a 14 x 14 map, completion on an mbarrier) and runs up to STAGES chunks -- across the passes and across pairs --
// ahead of the 8 consumer warps, which only ever read shared memory.  In the first form every consumer thread issued its
// own 4-byte global loads, 16 in flight, and sat on their DRAM latency (ncu: 45 % of the samples on long-scoreboard
// stalls at 2.2 TB/s); here the bytes in flight do not depend on registers or occupancy.  The arithmetic (order of every
// sum) is that of the first form, so the results are bit-identical.
// consumer threads per CTA.  Measured: 512 (16 consumer warps per CTA, one pixel tile per warp in the tensor-core pass, 54
// registers, still two CTAs per SM) runs 2.16 M pairs/s against 2.56 M for 256 -- the per-stage waits / releases grow with the warp
// count while the light passes (norms, key-point vectors) do not get shorter
#ifndef SPS_CONS_N
#define SPS_CONS_N 256
#endif
constexpr int SPS_CONS = SPS_CONS_N;        // consumer threads: thread = pixel in the norm / heat passes
constexpr int SPS_WARPS = SPS_CONS / 32;
constexpr int SPS_MT = (16 + SPS_WARPS - 1) / SPS_WARPS;  // m16 pixel tiles per warp in the tensor-core pass (h*w <= 256: 16 tiles)
constexpr int SPS_THREADS = SPS_CONS + 32;  // + one producer warp
constexpr int SPS_CC = 8;                   // channels per stage
// ring depth.  Measured (same box): 4 stages x 2 CTAs per SM 2.56 M pairs/s; 12 or 20 stages with ONE CTA per SM (the ring then
// takes the shared memory of the second CTA) 1.75 M -- the kernel is bound by its 16 consumer warps per SM, not by bytes in flight
#ifndef SPS_STAGES_N
#define SPS_STAGES_N 4
#endif
constexpr int SPS_STAGES = SPS_STAGES_N;

// m16n8k8 tf32 tensor-core step (legacy warp-level MMA: the heat map of a pair is 196 x 20 x 768, far below a tcgen05 tile)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi exactly representable in tf32 (the tensor core drops the low 13 mantissa bits of its operands);
// hi*hi + hi*lo + lo*hi carries ~21 mantissa bits ("3xTF32")
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// MMA: the heat-map pass on the tensor cores -- pixels on M (16 per tile, 2 tiles per warp), key points on N (8 per tile),
// one 8-channel stage = one K step, 3xTF32 -- instead of 21 FFMA + 6 LDS per channel and pixel.
template <int KT, bool MMA>
__global__ void __launch_bounds__(SPS_THREADS, SPS_STAGES_N > 5 ? 1 : 2) spair_stream_kernel(SpairBatchParams p) {
  static_assert(SPS_MT == 1 || SPS_MT == 2, "tile loop");
  using namespace sm100;
  extern __shared__ __align__(16) float4 sps_dyn[];
  __shared__ SpairScoreShared score;
  __shared__ int s_pred[64];
  __shared__ float s_wt[64][4];
  __shared__ int s_tap[64][4];
  __shared__ float s_bv[SPS_CONS / 32][KT];
  __shared__ int s_bi[SPS_CONS / 32][KT];
  __shared__ __align__(8) unsigned long long full[SPS_STAGES], empty[SPS_STAGES];
  const int C = p.C, hw = p.h * p.w, K = p.K;
  const int chunk_floats = SPS_CC * hw;
  const uint32_t chunk_bytes = (uint32_t)chunk_floats * 4u;
  float* ring = reinterpret_cast<float*>(sps_dyn);                  // [STAGES][CC * hw]
  float* q = ring + (size_t)SPS_STAGES * chunk_floats;              // [C][KT]
  float* pix_ss = q + (size_t)C * KT;                               // [hw]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nchunk = C / SPS_CC;
  const int nkt = (K + KT - 1) / KT;  // key-point tiles: passes Q and H repeat per tile

  if (tid == 0) {
    for (int s = 0; s < SPS_STAGES; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), SPS_CONS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (wid == SPS_CONS / 32) {
    // ===================================== producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* fi = p.feats + (size_t)b * 2 * C * hw;
        const float* fj = fi + (size_t)C * hw;
        for (int pass = 0; pass < 1 + 2 * nkt; ++pass) {
          const float* base = (pass == 0 || (pass & 1)) ? fi : fj;  // N, then (Q, H) per key-point tile
          for (int ck = 0; ck < nchunk; ++ck) {
            mbar_wait(smem_u32(&empty[stage]), phase ^ 1u);
            mbar_arrive_expect_tx(smem_u32(&full[stage]), chunk_bytes);
            bulk_load_1d(smem_u32(ring + (size_t)stage * chunk_floats), base + (size_t)ck * chunk_floats, chunk_bytes,
                         smem_u32(&full[stage]));
            if (++stage == SPS_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    return;
  }

  // ===================================== consumers =====================================
  int stage = 0;
  uint32_t phase = 0;
  auto cons_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(SPS_CONS) : "memory"); };
  auto release = [&] {  // this warp is done with the current stage
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty[stage]));
    if (++stage == SPS_STAGES) { stage = 0; phase ^= 1u; }
  };
  const int nwarp = SPS_CONS / 32;
  const int px = tid;
  const bool has_px = px < hw;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* ki = p.kps_i + (size_t)b * K * p.stride;
    const float* kj = p.kps_j + (size_t)b * K * p.stride;
    // ---- key-point taps: kp / size * 2 - 1 -> ((g + 1) / 2) * (size - 1)  (align_corners=True) ----
    if (tid < K) {
      const float kx = __fdiv_rn(ki[(size_t)tid * p.stride + 0], p.image_size);
      const float ky = __fdiv_rn(ki[(size_t)tid * p.stride + 1], p.image_size);
      const float gx = __fsub_rn(__fmul_rn(kx, 2.f), 1.f), gy = __fsub_rn(__fmul_rn(ky, 2.f), 1.f);
      const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(p.w - 1));
      const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(p.h - 1));
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float ww = ix - fx, we = 1.f - ww, wn = iy - fy, ws = 1.f - wn;
      const float wt[4] = {ws * we, ws * ww, wn * we, wn * ww};  // nw, ne, sw, se
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        const bool in = xx >= 0 && xx < p.w && yy >= 0 && yy < p.h;
        s_wt[tid][t] = in ? wt[t] : 0.f;
        s_tap[tid][t] = min(max(yy, 0), p.h - 1) * p.w + min(max(xx, 0), p.w - 1);
      }
    }
    // ---- pass N: ||f_i[px]||^2 of every pixel, channels in ascending order ----
    {
      float ss = 0.f;
      for (int ck = 0; ck < nchunk; ++ck) {
        mbar_wait(smem_u32(&full[stage]), phase);
        const float* st = ring + (size_t)stage * chunk_floats + px;
        if (has_px) {
#pragma unroll
          for (int c = 0; c < SPS_CC; ++c) {
            const float v = st[c * hw];
            ss = fmaf(v, v, ss);
          }
        }
        release();
      }
      if (has_px) pix_ss[px] = ss;
    }
    cons_sync();
    if (tid < 4 * K) {  // fold 1 / max(||f_i[tap]||, eps) into the blend weight
      const int k = tid >> 2, t = tid & 3;
      s_wt[k][t] = __fdiv_rn(s_wt[k][t], fmaxf(sqrtf(pix_ss[s_tap[k][t]]), SPB_NORM_EPS));
    }
    cons_sync();

    for (int k0 = 0; k0 < K; k0 += KT) {
      const int kt = min(KT, K - k0);
      // ---- pass Q: q[c][k] = sum_t wt * f_i[c][tap] / max(||f_i[tap]||, eps): the grid_sample of the normalised map ----
      for (int ck = 0; ck < nchunk; ++ck) {
        mbar_wait(smem_u32(&full[stage]), phase);
        const float* st = ring + (size_t)stage * chunk_floats;
        for (int idx = tid; idx < SPS_CC * KT; idx += SPS_CONS) {
          const int c = idx / KT, k = idx - c * KT;
          const int kk = k0 + min(k, kt - 1);
          const float* src = st + c * hw;
          float acc = src[s_tap[kk][0]] * s_wt[kk][0];
          acc = fmaf(src[s_tap[kk][1]], s_wt[kk][1], acc);
          acc = fmaf(src[s_tap[kk][2]], s_wt[kk][2], acc);
          acc = fmaf(src[s_tap[kk][3]], s_wt[kk][3], acc);
          q[(size_t)(ck * SPS_CC + c) * KT + k] = (k < kt) ? acc : 0.f;
        }
        release();
      }
      cons_sync();
      // ---- pass H: heat[k][px] = (sum_c q[c][k] * f_j[c][px]) / max(||f_j[px]||, eps), norm accumulated alongside ----
      if constexpr (MMA) {
        static_assert(SPS_CC == 8, "one stage = one m16n8k8 K step");
        constexpr int NT8 = (KT + 7) / 8;
        const int g = lane >> 2, t = lane & 3;
        float d[SPS_MT][NT8][4];
        float ssr[SPS_MT][2];
#pragma unroll
        for (int i = 0; i < SPS_MT; ++i) {
          ssr[i][0] = ssr[i][1] = 0.f;
#pragma unroll
          for (int j = 0; j < NT8; ++j) d[i][j][0] = d[i][j][1] = d[i][j][2] = d[i][j][3] = 0.f;
        }
        const int ntile = (hw + 15) >> 4;  // this warp owns tiles wid (and wid + SPS_WARPS)
        for (int ck = 0; ck < nchunk; ++ck) {
          mbar_wait(smem_u32(&full[stage]), phase);
          const float* st = ring + (size_t)stage * chunk_floats;
          uint32_t bh[NT8][2], bl[NT8][2];
#pragma unroll
          for (int j = 0; j < NT8; ++j) {
            const int kcol = 8 * j + g;
            const float b0 = kcol < KT ? q[(size_t)(ck * SPS_CC + t) * KT + kcol] : 0.f;
            const float b1 = kcol < KT ? q[(size_t)(ck * SPS_CC + t + 4) * KT + kcol] : 0.f;
            split_tf32(b0, bh[j][0], bl[j][0]);
            split_tf32(b1, bh[j][1], bl[j][1]);
          }
#pragma unroll
          for (int i = 0; i < SPS_MT; ++i) {
            const int mt = wid + SPS_WARPS * i;
            if (mt < ntile) {  // warp-uniform
              const int p0 = mt * 16 + g;
              const float a0 = st[t * hw + p0], a1 = st[t * hw + p0 + 8];            // rows beyond h*w read the following
              const float a2 = st[(t + 4) * hw + p0], a3 = st[(t + 4) * hw + p0 + 8];  // floats of the ring: masked below
              ssr[i][0] = fmaf(a0, a0, ssr[i][0]);
              ssr[i][0] = fmaf(a2, a2, ssr[i][0]);
              ssr[i][1] = fmaf(a1, a1, ssr[i][1]);
              ssr[i][1] = fmaf(a3, a3, ssr[i][1]);
              uint32_t ah[4], al[4];
              split_tf32(a0, ah[0], al[0]);
              split_tf32(a1, ah[1], al[1]);
              split_tf32(a2, ah[2], al[2]);
              split_tf32(a3, ah[3], al[3]);
#pragma unroll
              for (int j = 0; j < NT8; ++j) {
                mma_tf32(d[i][j], al, bh[j][0], bh[j][1]);
                mma_tf32(d[i][j], ah, bl[j][0], bl[j][1]);
                mma_tf32(d[i][j], ah, bh[j][0], bh[j][1]);
              }
            }
          }
          release();
        }
        // per key point this lane owns (columns 2t, 2t + 1 of every N tile): best of its rows, then across the 8 row groups
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            float v = -CUDART_INF_F;
            int bi = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < SPS_MT; ++i) {
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const int pxl = (wid + SPS_WARPS * i) * 16 + g + 8 * r;
                float ss = ssr[i][r];
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                const float hv = __fdiv_rn(d[i][j][2 * r + e2], fmaxf(sqrtf(ss), SPB_NORM_EPS));
                if (wid + SPS_WARPS * i < ntile && pxl < hw && (hv > v || (hv == v && pxl < bi))) {
                  v = hv;
                  bi = pxl;
                }
              }
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, v, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > v || (ov == v && oi < bi)) {
                v = ov;
                bi = oi;
              }
            }
            const int kcol = 8 * j + 2 * t + e2;
            if (g == 0 && kcol < KT) {
              s_bv[wid][kcol] = v;
              s_bi[wid][kcol] = bi;
            }
          }
        }
      } else {
      float acc[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) acc[k] = 0.f;
      float ss = 0.f;
      for (int ck = 0; ck < nchunk; ++ck) {
        mbar_wait(smem_u32(&full[stage]), phase);
        const float* st = ring + (size_t)stage * chunk_floats + px;
        if (has_px) {
#pragma unroll
          for (int c = 0; c < SPS_CC; ++c) {
            const float v = st[c * hw];
            ss = fmaf(v, v, ss);
            const float4* qc = reinterpret_cast<const float4*>(q + (size_t)(ck * SPS_CC + c) * KT);
#pragma unroll
            for (int k4 = 0; k4 < KT / 4; ++k4) {
              const float4 qq = qc[k4];  // same address in every lane: broadcast
              acc[4 * k4 + 0] = fmaf(qq.x, v, acc[4 * k4 + 0]);
              acc[4 * k4 + 1] = fmaf(qq.y, v, acc[4 * k4 + 1]);
              acc[4 * k4 + 2] = fmaf(qq.z, v, acc[4 * k4 + 2]);
              acc[4 * k4 + 3] = fmaf(qq.w, v, acc[4 * k4 + 3]);
            }
          }
        }
        release();
      }
      const float nrm = fmaxf(sqrtf(ss), SPB_NORM_EPS);
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        float v = has_px ? __fdiv_rn(acc[k], nrm) : -CUDART_INF_F;
        int i = has_px ? px : 0x7fffffff;
        if (!(v > -CUDART_INF_F)) {  // NaN / -inf never win (the first form's `hv > bestv` test against -inf)
          v = -CUDART_INF_F;
          i = 0x7fffffff;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, i, o);
          if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
          }
        }
        if (lane == 0) {
          s_bv[wid][k] = v;
          s_bi[wid][k] = i;
        }
      }
      }
      cons_sync();
      if (tid < kt) {
        float v = s_bv[0][tid];
        int i = s_bi[0][tid];
        for (int wq = 1; wq < nwarp; ++wq) {
          const float ov = s_bv[wq][tid];
          const int oi = s_bi[wq][tid];
          if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
          }
        }
        i = (i == 0x7fffffff) ? 0 : i;  // an all-NaN heat map: torch.argmax would return a NaN position; we return 0
        s_pred[k0 + tid] = i;
        if (p.pred) p.pred[(size_t)b * K + k0 + tid] = i;
      }
      cons_sync();  // q, s_bv are rewritten by the next tile
    }

    spair_score_block<SPS_CONS>(score, s_pred, K, p.w, ki, kj, p.stride, p.image_size, __ldg(p.thresh_scale + b), p.pck, nullptr,
                                p.error_same ? p.error_same + (size_t)b * K : nullptr,
                                p.error_nn ? p.error_nn + (size_t)b * K : nullptr,
                                p.index_nn ? p.index_nn + (size_t)b * K : nullptr, p.hits, p.confusion, p.conf_dim);
  }
}

// the streaming form needs one pixel per consumer thread and 16-byte-granular stages
bool spair_stream_ok(const SpairBatchParams& p) {
  static const int off = getenv("MVMATCH_SPAIR_STREAM") && getenv("MVMATCH_SPAIR_STREAM")[0] == '0';
  const int hw = p.h * p.w;
  return !off && hw <= 256 && p.C % SPS_CC == 0 && (hw * SPS_CC) % 4 == 0 && ((uintptr_t)p.feats & 15) == 0 &&
         ((size_t)p.C * hw) % 4 == 0;
}

template <int KT>
int launch_spair_stream(const SpairBatchParams& p, cudaStream_t st) {
  const int hw = p.h * p.w;
  const size_t smem = ((size_t)SPS_STAGES * SPS_CC * hw + (size_t)p.C * KT + (size_t)((hw + 3) / 4 * 4)) * sizeof(float);
  // MVMATCH_SPAIR_MMA=0: the heat-map pass on the CUDA cores (bit-identical to the first form)
  static const bool mma = !(getenv("MVMATCH_SPAIR_MMA") && getenv("MVMATCH_SPAIR_MMA")[0] == '0');
  auto kern = mma ? spair_stream_kernel<KT, true> : spair_stream_kernel<KT, false>;
  static size_t opted[MV_MAX_DEVICES][2];
  size_t& opted_in = opted[mv_device_slot()][mma ? 1 : 0];
  if (smem > opted_in) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      mv_set_error("mv_spair_match_batch: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    opted_in = smem;
  }
  int per_sm = (int)((227u << 10) / (smem + (21u << 10)));  // static shared memory of the kernel is ~20 KB
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int grid = mv_sm_count() * per_sm;
  if (grid > p.B) grid = p.B;
  kern<<<grid, SPS_THREADS, smem, st>>>(p);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

template <int KT>
int launch_spair_batch(const SpairBatchParams& p, cudaStream_t st) {
  if (spair_stream_ok(p) && ((size_t)SPS_STAGES * SPS_CC * p.h * p.w + (size_t)p.C * KT + p.h * p.w + 4) * sizeof(float) <= (200u << 10))
    return launch_spair_stream<KT>(p, st);
  const size_t smem = (size_t)p.C * KT * sizeof(float);
  auto kern = spair_batch_kernel<KT>;
  static size_t opted[MV_MAX_DEVICES];  // per device; static + dynamic shared memory above 48 KB needs the opt-in
  size_t& opted_in = opted[mv_device_slot()];
  if (opted_in < (24u << 10)) opted_in = 24 << 10;
  if (smem > opted_in) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      mv_set_error("mv_spair_match_batch: cannot opt in to %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
    opted_in = smem;
  }
  int per_sm = (int)((200u << 10) / (smem + (24u << 10)));  // static shared memory of the kernel is ~21 KB
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = mv_sm_count() * per_sm;
  if (grid > p.B) grid = p.B;
  int threads = ((p.h * p.w + 31) / 32) * 32;
  if (threads > SPB_THREADS) threads = SPB_THREADS;
  if (threads < 4 * p.K) threads = ((4 * p.K + 31) / 32) * 32;  // the tap set-up uses one thread per tap (K <= 64)
  kern<<<grid, threads, smem, st>>>(p);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

}  // namespace

extern "C" {

int mv_spair_match_batch(const float* feats, int B, int C, int h, int w, const float* kps_i, const float* kps_j, int K,
                         int kp_stride, const float* thresh_scale, float image_size, float pck_thresh,
                         int32_t* pred_flat, float* error_same, float* error_nn, int32_t* index_nn,
                         unsigned long long* hits, unsigned long long* confusion, int conf_dim, mv_stream_t stream) {
  MV_REQUIRE(B >= 0 && C > 0 && h > 0 && w > 0, MV_E_ARG, "mv_spair_match_batch: bad sizes");
  MV_REQUIRE((B == 0 || K == 0) || (feats && kps_i && kps_j && thresh_scale), MV_E_ARG,
             "mv_spair_match_batch: null pointer");
  MV_REQUIRE(h * w <= SPB_MAX_HW, MV_E_RANGE, "mv_spair_match_batch: h*w=%d exceeds %d pixels", h * w, SPB_MAX_HW);
  MV_REQUIRE(K >= 0 && K <= 64, MV_E_RANGE, "mv_spair_match_batch: K=%d must be in [0, 64]", K);
  MV_REQUIRE(kp_stride >= 3 && image_size > 0.f, MV_E_ARG, "mv_spair_match_batch: bad key-point layout");
  MV_REQUIRE(!confusion || conf_dim >= K, MV_E_ARG, "mv_spair_match_batch: conf_dim=%d must be >= K=%d", conf_dim, K);
  if (B == 0 || K == 0) return MV_OK;
  SpairBatchParams p;
  p.feats = feats;
  p.kps_i = kps_i;
  p.kps_j = kps_j;
  p.thresh_scale = thresh_scale;
  p.B = B;
  p.C = C;
  p.h = h;
  p.w = w;
  p.K = K;
  p.stride = kp_stride;
  p.image_size = image_size;
  p.pck = pck_thresh;
  p.pred = pred_flat;
  p.error_same = error_same;
  p.error_nn = error_nn;
  p.index_nn = index_nn;
  p.hits = hits;
  p.confusion = confusion;
  p.conf_dim = conf_dim;
  cudaStream_t st = mv_cuda_stream(stream);
  // key-point tile = accumulators per thread; the tile's features (C * KT floats) must fit shared memory
  const size_t budget = 160u << 10;
  const size_t per_k = (size_t)C * sizeof(float);
  MV_REQUIRE(8 * per_k <= budget, MV_E_RANGE, "mv_spair_match_batch: C=%d too large for the key-point tile", C);
  if (K > 24 && 32 * per_k <= budget) return launch_spair_batch<32>(p, st);
  if (K > 20 && 24 * per_k <= budget) return launch_spair_batch<24>(p, st);
  if (K > 16 && 20 * per_k <= budget) return launch_spair_batch<20>(p, st);  // SPair-71k: up to 20 key points for most classes
  if (K > 12 && 16 * per_k <= budget) return launch_spair_batch<16>(p, st);
  if (K > 8 && 12 * per_k <= budget) return launch_spair_batch<12>(p, st);
  return launch_spair_batch<8>(p, st);
}

}  // extern "C"
