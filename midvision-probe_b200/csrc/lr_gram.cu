// lr_gram.cu -- kernel 2's operands in the basis of the SOURCE pixels ("low-rank proposal", kernel 2g).
//
// Both dense helpers interpolate their rows linearly from an (h*w, C) source map
// (correspondence.py:164-176 bilinear grid_sample, :240-241 bicubic interpolate):
//     x_i = sum_s W0[i,s] src0[s]          y_j = sum_t W1[j,t] src1[t]
// so the cosine similarity kernel 2 ranks (correspondence.py:14-23, :47-48) factors through the source pixels:
//     cos(x_i, y_j) = sum_t A[i,t] * B[j,t]
//     A[i,t] = (x_i / |x_i|) . (src1[t] / |src1[t]|) = sum_s Wn0[i,s] * G[s,t]        G = U0 U1^T, U = unit source rows
//     B[j,t] = W1[j,t] |src1[t]| / |y_j|            (4 or 16 non-zeros per row)         Wn = W |src| / |x|
// i.e. a product over K = h*w instead of K = C (ScanNet-shaped pairs: 300 instead of 2048).  The N x M product itself still
// runs on kernel 2 (tcgen05, fused top-2 / column arg-max); this file builds its fp16 operands:
//     mv_lr_unit_rows      U = fp16(src / |src|), |src|                       (the Gram matrix G = U U^T of both images
//                          stacked comes from kernel 2 itself with the similarity written out, mv_k2_affinity)
//     mv_lr_build_query    A' = fp16(A - c_i), c_i = fp16(max_t A[i,t]), augmentation columns (c_i, c_i)
//     mv_lr_build_target   fp16(B), augmentation columns = two fp16 pieces of beta_j = sum_t B[j,t] (fp32, unrounded)
// so that  sum_k A'[i,k] B[j,k] = cos(x_i, y_j) + (fp16 rounding of A' and B only).  Centring a row on its own MAXIMUM
// makes A' smallest exactly where the competitive columns have their taps: the rounding error of the candidates that
// decide a row's top-2 is ~1e-6 even on nearly collinear CNN features (tools/gram_sim.py: 996-1000 of 1000 matches in
// common with the reference on ResNet-50 features, 1000 of 1000 on ViT features).
// The product only PROPOSES the two candidates of a row; kernel 3 recomputes their fp32 distances from the exact rows
// as before.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr float LR_NORM_EPS = 1e-12f;  // F.normalize's eps (correspondence.py:47-48)

// ---- unit source rows: one CTA per source pixel
__global__ void __launch_bounds__(128) lr_unit_rows_kernel(const float* __restrict__ src, int C, __half* __restrict__ U,
                                                           float* __restrict__ snorm) {
  __shared__ float part[4];
  const int p = blockIdx.x;
  const float4* row = reinterpret_cast<const float4*>(src + (size_t)p * C);
  const int C4 = C >> 2;
  float ss = 0.f;
  for (int c = threadIdx.x; c < C4; c += 128) {
    const float4 v = __ldg(row + c);
    ss = fmaf(v.x, v.x, ss);
    ss = fmaf(v.y, v.y, ss);
    ss = fmaf(v.z, v.z, ss);
    ss = fmaf(v.w, v.w, ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float tot = (part[0] + part[1]) + (part[2] + part[3]);
  const float nrm = sqrtf(tot);
  const float inv = nrm > 0.f ? 1.f / nrm : 0.f;
  uint2* out = reinterpret_cast<uint2*>(U + (size_t)p * C);
  for (int c = threadIdx.x; c < C4; c += 128) {
    const float4 v = __ldg(row + c);
    const __half2 a = __floats2half2_rn(v.x * inv, v.y * inv), b = __floats2half2_rn(v.z * inv, v.w * inv);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&a);
    pk.y = *reinterpret_cast<const uint32_t*>(&b);
    out[c] = pk;
  }
  if (threadIdx.x == 0) snorm[p] = nrm;
}

// Keys cubic convolution, A = -0.75 (ATen UpSample.h), as in kernel 1
__device__ __forceinline__ void lr_cubic(float t, float c[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__device__ __forceinline__ float pick4(const float c[4], int k) { return k == 0 ? c[0] : (k == 1 ? c[1] : (k == 2 ? c[2] : c[3])); }

// lane k < T holds tap k of the point at continuous source coordinates (x, y): source pixel (or -1) and blend weight.
// MV_SAMPLE_BILINEAR_ZEROS: 4 taps, out-of-range ones dropped (ATen grid_sampler, zeros padding);
// MV_SAMPLE_BICUBIC_CLAMP: 16 taps, indices clamped to the map (ATen upsample_bicubic2d) -- the same taps as kernel 1.
template <int MODE>
__device__ __forceinline__ void lane_tap(float x, float y, int h, int w, int lane, int& idx, float& wt) {
  const float fx = floorf(x), fy = floorf(y);
  const int x0 = (int)fx, y0 = (int)fy;
  idx = -1;
  wt = 0.f;
  if (MODE == MV_SAMPLE_BICUBIC_CLAMP) {
    if (lane < 16) {
      float cx[4], cy[4];
      lr_cubic(x - fx, cx);
      lr_cubic(y - fy, cy);
      const int ax = lane & 3, ay = lane >> 2;
      const int xx = min(max(x0 - 1 + ax, 0), w - 1), yy = min(max(y0 - 1 + ay, 0), h - 1);
      idx = yy * w + xx;
      wt = pick4(cy, ay) * pick4(cx, ax);
    }
  } else {
    if (lane < 4) {
      const float ww = x - fx, we = 1.f - ww, wn = y - fy, ws = 1.f - wn;
      const int xx = x0 + (lane & 1), yy = y0 + (lane >> 1);
      if (xx >= 0 && xx < w && yy >= 0 && yy < h) {
        idx = yy * w + xx;
        wt = ((lane >> 1) ? wn : ws) * ((lane & 1) ? ww : we);
      }
    }
  }
}

// |x|^2 = sum_{a,b} V_a V_b G_own[idx_a, idx_b] for the T taps held by lanes 0..T-1 (V = blend weight * |src|); returns
// 1 / max(|x|, eps) on every lane
template <int T>
__device__ __forceinline__ float inv_norm_from_gram(int idx, float V, const float* __restrict__ G, int ld, int off, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < (T * T + 31) / 32; ++k) {
    const int q = lane + 32 * k;
    const int a = (q / T) % T, b = q % T;  // (q / T) % T keeps the shuffle source valid for lanes beyond T*T
    const float Va = __shfl_sync(0xffffffffu, V, a), Vb = __shfl_sync(0xffffffffu, V, b);
    const int ia = __shfl_sync(0xffffffffu, idx, a), ib = __shfl_sync(0xffffffffu, idx, b);
    if (q < T * T && ia >= 0 && ib >= 0) acc = fmaf(Va * Vb, __ldg(G + (size_t)(off + ia) * ld + off + ib), acc);
  }
  acc = warp_sum(acc);
  return 1.f / fmaxf(sqrtf(fmaxf(acc, 0.f)), LR_NORM_EPS);
}

struct LrParams {
  const float* coords;     // (n, 2) continuous source coordinates (x, y) of every point: what kernel 1 samples at
  const int32_t* n_dev;
  int n_max, h, w, hw, hwp;
  const float* snorm;      // (hw) |src[p]| of THIS image
  const float* G;          // stacked cosine Gram of the unit source rows of both images, fp32, row pitch ld
  int ld, off_own, off_tgt;  // row / column offset of this image's and of the target image's source pixels in G
  __half* out;             // (n, pitch) fp16 operand rows
  int pitch;
};

// ---- target rows: fp16(B) scattered into a zero row + the two fp16 pieces of beta.  One warp per point.
template <int MODE>
__global__ void __launch_bounds__(256) lr_build_target_kernel(LrParams p) {
  constexpr int T = MODE == MV_SAMPLE_BICUBIC_CLAMP ? 16 : 4;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = p.n_dev ? min(*p.n_dev, p.n_max) : p.n_max;
  if (j >= n) return;
  const float2 xy = __ldg(reinterpret_cast<const float2*>(p.coords) + j);
  int idx;
  float wt;
  lane_tap<MODE>(xy.x, xy.y, p.h, p.w, lane, idx, wt);
  const float V = idx >= 0 ? wt * __ldg(p.snorm + idx) : 0.f;
  const float inv = inv_norm_from_gram<T>(idx, V, p.G, p.ld, p.off_own, lane);
  const float Vn = V * inv;
  __half* row = p.out + (size_t)j * p.pitch;
  for (int c = lane * 8; c < p.pitch; c += 256) *reinterpret_cast<uint4*>(row + c) = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  // clamped border taps can name the same source pixel more than once: the lowest lane writes the sum (fixed order)
  float tot = 0.f;
  bool first = true;
#pragma unroll
  for (int b = 0; b < T; ++b) {
    const int ib = __shfl_sync(0xffffffffu, idx, b);
    const float vb = __shfl_sync(0xffffffffu, Vn, b);
    if (ib == idx) {
      tot += vb;
      if (b < lane) first = false;
    }
  }
  if (lane < T && idx >= 0 && first) row[idx] = __float2half_rn(tot);
  const float beta = warp_sum(Vn);  // the UNROUNDED row sum: the rounding of B must not leak into the constant term
  if (lane == 0) {
    const __half q0 = __float2half_rn(beta);
    const __half q1 = __float2half_rn(beta - __half2float(q0));
    row[p.hwp] = q0;
    row[p.hwp + 1] = q1;
  }
}

// ---- query rows: A[i,t] = sum_a Vn_a G[idx_a, t], centred on the row maximum.  One warp per point; lane l owns the
// column pairs (2 l + 64 k, 2 l + 64 k + 1).
template <int MODE, int KT>
__global__ void __launch_bounds__(256) lr_build_query_kernel(LrParams p) {
  constexpr int T = MODE == MV_SAMPLE_BICUBIC_CLAMP ? 16 : 4;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = p.n_dev ? min(*p.n_dev, p.n_max) : p.n_max;
  if (i >= n) return;
  const float2 xy = __ldg(reinterpret_cast<const float2*>(p.coords) + i);
  int idx;
  float wt;
  lane_tap<MODE>(xy.x, xy.y, p.h, p.w, lane, idx, wt);
  const float V = idx >= 0 ? wt * __ldg(p.snorm + idx) : 0.f;
  const float inv = inv_norm_from_gram<T>(idx, V, p.G, p.ld, p.off_own, lane);
  const float Vn = V * inv;
  float2 acc[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int a = 0; a < T; ++a) {
    const int ia = __shfl_sync(0xffffffffu, idx, a);
    const float va = __shfl_sync(0xffffffffu, Vn, a);
    if (ia < 0) continue;  // warp-uniform
    const float2* g = reinterpret_cast<const float2*>(p.G + (size_t)(p.off_own + ia) * p.ld + p.off_tgt);
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int c2 = lane + 32 * k;
      if (2 * c2 < p.hwp) {
        const float2 v = __ldg(g + c2);
        acc[k].x = fmaf(va, v.x, acc[k].x);
        acc[k].y = fmaf(va, v.y, acc[k].y);
      }
    }
  }
  float mx = -3.0e38f;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int c = 2 * (lane + 32 * k);
    if (c < p.hw) mx = fmaxf(mx, acc[k].x);
    if (c + 1 < p.hw) mx = fmaxf(mx, acc[k].y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const __half ch = __float2half_rn(mx);
  const float c0 = __half2float(ch);
  __half* row = p.out + (size_t)i * p.pitch;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int c = 2 * (lane + 32 * k);
    if (c < p.hwp) {
      const float a = c < p.hw ? acc[k].x - c0 : 0.f, b = c + 1 < p.hw ? acc[k].y - c0 : 0.f;
      *reinterpret_cast<__half2*>(row + c) = __floats2half2_rn(a, b);
    }
  }
  if (lane == 0) {
    __half aug[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) aug[k] = __float2half_rn(0.f);
    aug[0] = ch;  // against the two pieces of beta_j in the target rows
    aug[1] = ch;
    *reinterpret_cast<uint4*>(row + p.hwp) = *reinterpret_cast<const uint4*>(aug);
  }
}

template <int MODE>
int launch_query(const LrParams& p, cudaStream_t st) {
  const int grid = (p.n_max + 7) / 8;
  const int kt = (p.hwp + 63) / 64;
  if (kt <= 5) lr_build_query_kernel<MODE, 5><<<grid, 256, 0, st>>>(p);
  else if (kt <= 8) lr_build_query_kernel<MODE, 8><<<grid, 256, 0, st>>>(p);
  else if (kt <= 13) lr_build_query_kernel<MODE, 13><<<grid, 256, 0, st>>>(p);
  else lr_build_query_kernel<MODE, 16><<<grid, 256, 0, st>>>(p);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int check_params(const char* who, int mode, const float* coords, int n_max, int h, int w, const float* snorm, const float* G,
                 int ld_g, int off_own, int off_tgt, const void* out, int pitch, int hwp) {
  MV_REQUIRE(mode == MV_SAMPLE_BILINEAR_ZEROS || mode == MV_SAMPLE_BICUBIC_CLAMP, MV_E_ARG, "%s: mode must be bilinear-zeros or bicubic-clamp", who);
  MV_REQUIRE(coords && snorm && G && out, MV_E_ARG, "%s: null pointer", who);
  MV_REQUIRE(n_max > 0 && h > 0 && w > 0, MV_E_ARG, "%s: sizes must be positive", who);
  MV_REQUIRE(hwp >= h * w && hwp % 8 == 0 && hwp <= MV_LR_MAX_SOURCE_PIXELS, MV_E_RANGE,
             "%s: hwp=%d must be a multiple of 8 in [h*w, %d]", who, hwp, MV_LR_MAX_SOURCE_PIXELS);
  MV_REQUIRE(pitch >= hwp + 8 && pitch % 8 == 0 && ((uintptr_t)out & 15) == 0, MV_E_ALIGN,
             "%s: operand rows need a 16-byte aligned base and a pitch >= hwp + 8 that is a multiple of 8", who);
  MV_REQUIRE(off_own >= 0 && off_tgt >= 0 && off_own % 2 == 0 && off_tgt % 2 == 0 && ld_g % 2 == 0 && ld_g >= off_own + hwp &&
                 ld_g >= off_tgt + hwp && ((uintptr_t)G & 7) == 0,
             MV_E_ALIGN, "%s: the Gram matrix needs even offsets / pitch covering both images and an 8-byte aligned base", who);
  return MV_OK;
}

}  // namespace

extern "C" {

int mv_lr_unit_rows(const float* src_hwc, int C, int hw, void* U_f16, float* snorm, mv_stream_t stream) {
  MV_REQUIRE(src_hwc && U_f16 && snorm, MV_E_ARG, "mv_lr_unit_rows: null pointer");
  MV_REQUIRE(C > 0 && C % 4 == 0 && hw > 0, MV_E_ARG, "mv_lr_unit_rows: C must be a positive multiple of 4, hw positive");
  MV_REQUIRE(((uintptr_t)src_hwc & 15) == 0 && ((uintptr_t)U_f16 & 7) == 0, MV_E_ALIGN, "mv_lr_unit_rows: misaligned buffers");
  lr_unit_rows_kernel<<<hw, 128, 0, mv_cuda_stream(stream)>>>(src_hwc, C, reinterpret_cast<__half*>(U_f16), snorm);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_lr_build_query(int mode, const float* coords, const int32_t* n_dev, int n_max, int h, int w, const float* snorm,
                      const float* G, int ld_g, int off_own, int off_tgt, void* A_f16, int pitch, int hwp, mv_stream_t stream) {
  int rc = check_params("mv_lr_build_query", mode, coords, n_max, h, w, snorm, G, ld_g, off_own, off_tgt, A_f16, pitch, hwp);
  if (rc) return rc;
  LrParams p{coords, n_dev, n_max, h, w, h * w, hwp, snorm, G, ld_g, off_own, off_tgt, reinterpret_cast<__half*>(A_f16), pitch};
  return mode == MV_SAMPLE_BICUBIC_CLAMP ? launch_query<MV_SAMPLE_BICUBIC_CLAMP>(p, mv_cuda_stream(stream))
                                         : launch_query<MV_SAMPLE_BILINEAR_ZEROS>(p, mv_cuda_stream(stream));
}

int mv_lr_build_target(int mode, const float* coords, const int32_t* n_dev, int n_max, int h, int w, const float* snorm,
                       const float* G, int ld_g, int off_own, void* B_f16, int pitch, int hwp, mv_stream_t stream) {
  int rc = check_params("mv_lr_build_target", mode, coords, n_max, h, w, snorm, G, ld_g, off_own, off_own, B_f16, pitch, hwp);
  if (rc) return rc;
  LrParams p{coords, n_dev, n_max, h, w, h * w, hwp, snorm, G, ld_g, off_own, off_own, reinterpret_cast<__half*>(B_f16), pitch};
  const int grid = (n_max + 7) / 8;
  if (mode == MV_SAMPLE_BICUBIC_CLAMP) lr_build_target_kernel<MV_SAMPLE_BICUBIC_CLAMP><<<grid, 256, 0, mv_cuda_stream(stream)>>>(p);
  else lr_build_target_kernel<MV_SAMPLE_BILINEAR_ZEROS><<<grid, 256, 0, mv_cuda_stream(stream)>>>(p);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

}  // extern "C"
