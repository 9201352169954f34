// k1_sample.cu -- kernel 1 of the matching path and its small producers.
//
//   mv_chw_to_hwc            (C,h,w) backbone map -> channel-last (h*w, C), optional per-pixel L2 norm
//   mv_compact_valid         stable compaction of the live (z > 0) pixels
//   mv_geom_*                per-point source coordinates for the three reference call sites
//   mv_k1_sample_normalize   gather/upsample + L2-normalise + bf16/fp32 row writer   (HBM-bound)
//
// Reference behaviour being reproduced (file:line in /root/reference):
//   evals/utils/correspondence.py:132-176  get_grid / grid_to_pointcloud / sample_pointcloud_features
//   evals/utils/correspondence.py:240-252  bicubic upsample + valid-pixel gather
//   evals/utils/correspondence.py:47-48    F.normalize of the sampled rows
//   evaluate_spair_correspondence.py:59-79 per-pixel normalise + keypoint grid_sample(align_corners=True)
//
// Data layout: the source map is channel-last so that every tap is one contiguous C-vector read with
// 128-bit loads; consecutive live points are handled by the same CTA, which keeps the taps they share
// in registers (an 8x upsample re-uses each tap ~8 times along x), so the kernel's traffic is the
// row writes: C*h*w*4 bytes read + n*C*(2 [+4]) bytes written.
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int K1_THREADS_MAX = 512;  // 128 registers per thread: the tap window + a batch of accumulators fit
constexpr int K1_BATCH_MAX = 4;  // points per block-level reduction (template parameter K1_BATCH <= this)
constexpr int K1_MAX_RUN = 64;  // most points one CTA owns
constexpr float K1_NORM_EPS = 1e-12f;  // F.normalize default eps

// ------------------------------------------------------------------------------------------
// (C, hw) -> (hw, C) transpose, optional scale by 1/max(norm, eps)
// ------------------------------------------------------------------------------------------
__global__ void pixel_norm_kernel(const float* __restrict__ src, float* __restrict__ norm, int C, int hw) {
  // block = (32 pixels, 8 channel slices)
  __shared__ float part[8][33];
  const int px = blockIdx.x * 32 + threadIdx.x;
  float ss = 0.f;
  if (px < hw) {
    for (int c = threadIdx.y; c < C; c += 8) {
      float v = __ldg(src + (size_t)c * hw + px);
      ss = fmaf(v, v, ss);
    }
  }
  part[threadIdx.y][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.y == 0 && px < hw) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    norm[px] = sqrtf(t);
  }
}

__global__ void chw_to_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int hw,
                                  const float* __restrict__ norm) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  // read: x = pixel (contiguous in src), y = channel
  for (int dy = threadIdx.y; dy < 32; dy += 8) {
    int c = c0 + dy, p = p0 + threadIdx.x;
    float v = 0.f;
    if (c < C && p < hw) v = __ldg(src + (size_t)c * hw + p);
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  // write: x = channel (contiguous in dst), y = pixel
  for (int dy = threadIdx.y; dy < 32; dy += 8) {
    int p = p0 + dy, c = c0 + threadIdx.x;
    if (c < C && p < hw) {
      float v = tile[threadIdx.x][dy];
      if (norm) v = __fdiv_rn(v, fmaxf(__ldg(norm + p), K1_NORM_EPS));
      dst[(size_t)p * C + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// stable compaction (single CTA; n <= 2^20)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) compact_valid_kernel(const float* __restrict__ z, int z_stride, int n,
                                                             int32_t* __restrict__ valid_idx,
                                                             int32_t* __restrict__ n_valid) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (n + 1023) / 1024;
  const int beg = min(tid * per, n), end = min(beg + per, n);
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += (__ldg(z + (size_t)i * z_stride) > 0.f) ? 1 : 0;
  // inclusive warp scan
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int t = warp_tot[lane];
    int s = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    warp_tot[lane] = s - t;  // exclusive prefix of warp totals
    if (lane == 31) *n_valid = s;
  }
  __syncthreads();
  int pos = warp_tot[wid] + incl - cnt;
  for (int i = beg; i < end; ++i)
    if (__ldg(z + (size_t)i * z_stride) > 0.f) valid_idx[pos++] = i;
}

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
struct Mat3 {
  float m[9];
};

// correspondence.py:132-161: points = depth * (x+.5, y+.5, 1); xyz = Kinv @ points
__global__ void backproject_kernel(const float* __restrict__ depth, int H, int W, Mat3 Ki,
                                   float* __restrict__ xyz_all) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  int y = p / W, x = p - y * W;
  float d = __ldg(depth + p);
  float px = __fmul_rn(d, (float)x + 0.5f), py = __fmul_rn(d, (float)y + 0.5f), pz = d;  // depth * grid
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float a = __fmul_rn(Ki.m[3 * r + 0], px);
    a = fmaf(Ki.m[3 * r + 1], py, a);
    a = fmaf(Ki.m[3 * r + 2], pz, a);
    xyz_all[(size_t)p * 3 + r] = a;
  }
}

// correspondence.py:164-170 + ATen CPU grid_sampler (align_corners=False): ix = (g+1)*(size/2) - 0.5
__global__ void project_coords_kernel(const float* __restrict__ xyz_all, const int32_t* __restrict__ valid_idx,
                                      const int32_t* __restrict__ n_dev, int n_max, Mat3 K, int H, int W, int h,
                                      int w, float* __restrict__ xyz, float* __restrict__ coords) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = n_dev ? min(*n_dev, n_max) : n_max;
  if (i >= n) return;
  int p = valid_idx ? valid_idx[i] : i;
  float X = xyz_all[(size_t)p * 3 + 0], Y = xyz_all[(size_t)p * 3 + 1], Z = xyz_all[(size_t)p * 3 + 2];
  xyz[(size_t)i * 3 + 0] = X;
  xyz[(size_t)i * 3 + 1] = Y;
  xyz[(size_t)i * 3 + 2] = Z;
  float uvd[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // uvd = pc @ K^T
    float a = __fmul_rn(X, K.m[3 * r + 0]);
    a = fmaf(Y, K.m[3 * r + 1], a);
    a = fmaf(Z, K.m[3 * r + 2], a);
    uvd[r] = a;
  }
  float den = fmaxf(uvd[2], 1e-9f);
  float u = __fdiv_rn(uvd[0], den), v = __fdiv_rn(uvd[1], den);
  float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, u), (float)W), 1.f);
  float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, v), (float)H), 1.f);
  coords[(size_t)i * 2 + 0] = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w * 0.5f), 0.5f);
  coords[(size_t)i * 2 + 1] = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h * 0.5f), 0.5f);
}

// correspondence.py:240-252: bicubic source index scale*(dst+0.5)-0.5, uv = pixel centre, xyz gather
__global__ void grid_coords_kernel(const float* __restrict__ xyz_grid, const int32_t* __restrict__ valid_idx,
                                   const int32_t* __restrict__ n_dev, int n_max, int H, int W, int h, int w,
                                   float* __restrict__ xyz, float* __restrict__ uv, float* __restrict__ coords) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = n_dev ? min(*n_dev, n_max) : n_max;
  if (i >= n) return;
  int p = valid_idx ? valid_idx[i] : i;
  int y = p / W, x = p - y * W;
  const size_t HW = (size_t)H * W;
  if (xyz) {
    xyz[(size_t)i * 3 + 0] = __ldg(xyz_grid + p);
    xyz[(size_t)i * 3 + 1] = __ldg(xyz_grid + HW + p);
    xyz[(size_t)i * 3 + 2] = __ldg(xyz_grid + 2 * HW + p);
  }
  if (uv) {
    uv[(size_t)i * 2 + 0] = (float)x + 0.5f;
    uv[(size_t)i * 2 + 1] = (float)y + 0.5f;
  }
  float sx = (float)w / (float)W, sy = (float)h / (float)H;
  coords[(size_t)i * 2 + 0] = __fsub_rn(__fmul_rn(sx, (float)x + 0.5f), 0.5f);
  coords[(size_t)i * 2 + 1] = __fsub_rn(__fmul_rn(sy, (float)y + 0.5f), 0.5f);
}

// evaluate_spair_correspondence.py:71-78 + ATen CUDA grid_sampler (align_corners=True):
//   g = kp/size*2-1 ; ix = ((g+1)/2)*(w-1)
__global__ void keypoint_coords_kernel(const float* __restrict__ kps, int stride, int n, float image_size, int h,
                                       int w, float* __restrict__ coords) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float kx = __fdiv_rn(kps[(size_t)i * stride + 0], image_size);
  float ky = __fdiv_rn(kps[(size_t)i * stride + 1], image_size);
  float gx = __fsub_rn(__fmul_rn(kx, 2.f), 1.f), gy = __fsub_rn(__fmul_rn(ky, 2.f), 1.f);
  coords[(size_t)i * 2 + 0] = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(w - 1));
  coords[(size_t)i * 2 + 1] = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(h - 1));
}

// ------------------------------------------------------------------------------------------
// kernel 1
// ------------------------------------------------------------------------------------------
struct K1Params {
  const float* src;
  const float* coords;
  const int32_t* n_dev;
  int n_max, C, h, w, normalize;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  int32_t* taps;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x);
  a.y = fmaf(w, v.y, a.y);
  a.z = fmaf(w, v.z, a.z);
  a.w = fmaf(w, v.w, a.w);
}

// Keys cubic convolution, A = -0.75 (ATen UpSample.h cubic_convolution1/2)
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// One CTA owns a run of consecutive points and all C channels (thread = 4 * NV channels), so the taps that
// neighbouring points share stay in REGISTERS instead of being re-read through L1/L2:
//   bilinear: the raw 2 x 2 tap window is kept; an 8x upsample re-uses it for ~8 consecutive points and a
//             step to the next source cell loads 2 new taps instead of 4 (same arithmetic as a cold point);
//   bicubic : the 4 columns of the 4 x 4 window are pre-blended along y (consecutive points of an output row
//             have the bit-identical y coordinate), so a step in x costs 4 tap loads instead of 16.
template <int MODE, int NV, int K1_BATCH, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k1_sample_normalize_kernel(K1Params p) {
  __shared__ float red[K1_BATCH_MAX][K1_THREADS_MAX / 32];
  __shared__ float bcast[K1_BATCH_MAX];
  __shared__ float2 s_coords[K1_MAX_RUN];
  const int n = p.n_dev ? min(*p.n_dev, p.n_max) : p.n_max;
  const int C4 = p.C >> 2;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
  // run length from the LIVE point count: every CTA of the (host-sized) grid gets an equal share, so a
  // device-resident count far below n_max does not leave most SMs idle
  int ppc = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  ppc = min(max((ppc + K1_BATCH - 1) / K1_BATCH * K1_BATCH, K1_BATCH), K1_MAX_RUN);
  const int pt_beg = blockIdx.x * ppc;
  const int pt_end = min(pt_beg + ppc, n);

  // tap cache (see above)
  constexpr int NWIN = (MODE == MV_SAMPLE_BILINEAR_ZEROS) ? 4 : (MODE == MV_SAMPLE_BICUBIC_CLAMP ? 4 : 1);
  float4 win[NWIN][NV];
  int wx = INT_MIN, wy = INT_MIN;
  float wiy = 0.f;

  auto load_tap = [&](int xx, int yy, float4 (&dst)[NV]) {  // zero outside the map (bilinear zero padding)
    const bool in = xx >= 0 && xx < p.w && yy >= 0 && yy < p.h;
    const float* row = p.src + ((size_t)(in ? yy : 0) * p.w + (in ? xx : 0)) * p.C;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = tid + v * blockDim.x;
      dst[v] = (in && c4 < C4) ? ld4(row + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto load_cubic_col = [&](int xx, int y0, const float (&cy)[4], float4 (&dst)[NV]) {  // sum_i cy[i] * src[clamp(y0-1+i)][clamp(xx)]
    const int xc = min(max(xx, 0), p.w - 1);
#pragma unroll
    for (int v = 0; v < NV; ++v) dst[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int yc = min(max(y0 - 1 + i, 0), p.h - 1);
      const float* row = p.src + ((size_t)yc * p.w + xc) * p.C;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c4 = tid + v * blockDim.x;
        if (c4 < C4) fma4(dst[v], cy[i], ld4(row + 4 * c4));
      }
    }
  };

  // source coordinates of the whole run, once, coalesced (kills the coords -> address latency per point)
  if (MODE != MV_SAMPLE_ROWS) {
    for (int q = tid; q < pt_end - pt_beg; q += blockDim.x)
      s_coords[q] = __ldg(reinterpret_cast<const float2*>(p.coords) + pt_beg + q);
    __syncthreads();
  }

  // points are processed K1_BATCH at a time: one pair of block barriers (the L2-norm reduction) per batch
  for (int pt0 = pt_beg; pt0 < pt_end; pt0 += K1_BATCH) {
    float4 acc[K1_BATCH][NV];
#pragma unroll
    for (int b = 0; b < K1_BATCH; ++b) {
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[b][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int pt = pt0 + b;
      if (pt >= pt_end) continue;

      if (MODE == MV_SAMPLE_ROWS) {
        const float* row = p.src + (size_t)pt * p.C;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          int c4 = tid + v * blockDim.x;
          if (c4 < C4) acc[b][v] = ld4(row + 4 * c4);
        }
      } else if (MODE == MV_SAMPLE_BILINEAR_ZEROS) {
        // ATen GridSamplerKernel.cpp (bilinear, zeros): w = ix - floor(ix), e = 1 - w, n = iy - floor(iy), s = 1 - n
        const float ix = s_coords[pt - pt_beg].x, iy = s_coords[pt - pt_beg].y;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float ww = ix - fx, we = 1.f - ww, wn = iy - fy, ws = 1.f - wn;
        if (p.taps && tid == 0) {
          p.taps[2 * (size_t)pt] = x0;
          p.taps[2 * (size_t)pt + 1] = y0;
        }
        // window layout: win[0] = (x0, y0) nw, win[1] = (x0+1, y0) ne, win[2] = (x0, y0+1) sw, win[3] = (x0+1, y0+1) se
        if (y0 == wy && x0 == wx) {
          // same source cell: every tap is already in registers
        } else if (y0 == wy && x0 == wx + 1) {
#pragma unroll
          for (int v = 0; v < NV; ++v) { win[0][v] = win[1][v]; win[2][v] = win[3][v]; }
          load_tap(x0 + 1, y0, win[1]);
          load_tap(x0 + 1, y0 + 1, win[3]);
        } else {
          load_tap(x0, y0, win[0]);
          load_tap(x0 + 1, y0, win[1]);
          load_tap(x0, y0 + 1, win[2]);
          load_tap(x0 + 1, y0 + 1, win[3]);
        }
        wx = x0;
        wy = y0;
        const float wt[4] = {ws * we, ws * ww, wn * we, wn * ww};  // nw, ne, sw, se
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int v = 0; v < NV; ++v) fma4(acc[b][v], wt[k], win[k][v]);
        }
      } else {  // MV_SAMPLE_BICUBIC_CLAMP
        const float ix = s_coords[pt - pt_beg].x, iy = s_coords[pt - pt_beg].y;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy;
        if (p.taps && tid == 0) {
          p.taps[2 * (size_t)pt] = x0;
          p.taps[2 * (size_t)pt + 1] = y0;
        }
        float cx[4], cy[4];
        cubic_coeffs(ix - fx, cx);
        cubic_coeffs(iy - fy, cy);
        // win[j] = y-blended column x0 - 1 + j; valid while iy is bit-identical
        const bool same_row = (wy == y0) && (__float_as_uint(wiy) == __float_as_uint(iy));
        if (same_row && x0 == wx) {
        } else if (same_row && x0 == wx + 1) {
#pragma unroll
          for (int v = 0; v < NV; ++v) { win[0][v] = win[1][v]; win[1][v] = win[2][v]; win[2][v] = win[3][v]; }
          load_cubic_col(x0 + 2, y0, cy, win[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) load_cubic_col(x0 - 1 + j, y0, cy, win[j]);
        }
        wx = x0;
        wy = y0;
        wiy = iy;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int v = 0; v < NV; ++v) fma4(acc[b][v], cx[j], win[j][v]);
        }
      }
    }

    float denom[K1_BATCH];
#pragma unroll
    for (int b = 0; b < K1_BATCH; ++b) denom[b] = 1.f;
    if (p.normalize) {
#pragma unroll
      for (int b = 0; b < K1_BATCH; ++b) {
        float ss = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          ss = fmaf(acc[b][v].x, acc[b][v].x, ss);
          ss = fmaf(acc[b][v].y, acc[b][v].y, ss);
          ss = fmaf(acc[b][v].z, acc[b][v].z, ss);
          ss = fmaf(acc[b][v].w, acc[b][v].w, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) red[b][wid] = ss;
      }
      __syncthreads();
      for (int b = wid; b < K1_BATCH; b += nwarp) {  // warp b finishes point b
        float t = (lane < nwarp) ? red[b][lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) bcast[b] = __frcp_rn(fmaxf(sqrtf(t), K1_NORM_EPS));
      }
      __syncthreads();
#pragma unroll
      for (int b = 0; b < K1_BATCH; ++b) denom[b] = bcast[b];
    }

#pragma unroll
    for (int b = 0; b < K1_BATCH; ++b) {
      const int pt = pt0 + b;
      if (pt >= pt_end) continue;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        int c4 = tid + v * blockDim.x;
        if (c4 < C4) {
          float4 o = acc[b][v];
          if (p.normalize) {  // x * (1 / max(||x||, eps)): within 1 ulp of F.normalize's division
            o.x *= denom[b];
            o.y *= denom[b];
            o.z *= denom[b];
            o.w *= denom[b];
          }
          if (p.out_f32) __stcs(reinterpret_cast<float4*>(p.out_f32 + (size_t)pt * p.C + 4 * c4), o);
          if (p.out_bf16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(p.out_bf16 + (size_t)pt * p.C + 4 * c4) = pk;
          }
        }
      }
    }
    // `red`/`bcast` are rewritten next batch only after the first __syncthreads there, and every thread
    // has read `bcast` before it can pass that barrier: no extra barrier needed.
  }
}

template <int MODE>
int launch_k1(const K1Params& p, int threads, int nv, int grid, cudaStream_t st) {
  // thread = 4 * NV channels.  NV = 4 amortises the per-point scalar work (coordinates, cubic weights, window
  // bookkeeping, reduction) over 16 outputs; its batch is 2 points to stay inside 128 registers.
  switch (nv) {
    case 1: k1_sample_normalize_kernel<MODE, 1, 4, 128, 1><<<grid, threads, 0, st>>>(p); break;
    case 2: k1_sample_normalize_kernel<MODE, 2, 4, 128, 1><<<grid, threads, 0, st>>>(p); break;
    case 4:
      // C <= 4096: at most 256 threads, registers to spare for the loads in flight; wider rows still run (512 threads)
      if (threads <= 256) k1_sample_normalize_kernel<MODE, 4, 2, 256, 2><<<grid, threads, 0, st>>>(p);
      else k1_sample_normalize_kernel<MODE, 4, 2, 512, 1><<<grid, threads, 0, st>>>(p);
      break;
    default:
      mv_set_error("mv_k1_sample_normalize: unsupported channel split");
      return MV_E_RANGE;
  }
  MV_LAUNCH_CHECK();
  return MV_OK;
}

Mat3 load_mat3(const float* host) {
  Mat3 m;
  for (int i = 0; i < 9; ++i) m.m[i] = host[i];
  return m;
}

}  // namespace

extern "C" {

int mv_chw_to_hwc(const float* src_chw, float* dst_hwc, int C, int hw, int prenorm, float* norm_scratch,
                  mv_stream_t stream) {
  MV_REQUIRE(src_chw && dst_hwc, MV_E_ARG, "mv_chw_to_hwc: null pointer");
  MV_REQUIRE(C > 0 && hw > 0, MV_E_ARG, "mv_chw_to_hwc: C and hw must be positive");
  MV_REQUIRE(!prenorm || norm_scratch, MV_E_ARG, "mv_chw_to_hwc: prenorm needs norm_scratch (hw floats)");
  cudaStream_t st = mv_cuda_stream(stream);
  if (prenorm) {
    pixel_norm_kernel<<<(hw + 31) / 32, dim3(32, 8), 0, st>>>(src_chw, norm_scratch, C, hw);
    MV_LAUNCH_CHECK();
  }
  dim3 grid((hw + 31) / 32, (C + 31) / 32);
  chw_to_hwc_kernel<<<grid, dim3(32, 8), 0, st>>>(src_chw, dst_hwc, C, hw, prenorm ? norm_scratch : nullptr);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_compact_valid(const float* z, int z_stride, int n, int32_t* valid_idx, int32_t* n_valid,
                     mv_stream_t stream) {
  MV_REQUIRE(z && valid_idx && n_valid, MV_E_ARG, "mv_compact_valid: null pointer");
  MV_REQUIRE(n >= 0 && n <= (1 << 20) && z_stride >= 1, MV_E_RANGE, "mv_compact_valid: n must be in [0, 2^20]");
  compact_valid_kernel<<<1, 1024, 0, mv_cuda_stream(stream)>>>(z, z_stride, n, valid_idx, n_valid);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_geom_backproject(const float* depth, int H, int W, const float* Kinv_host, float* xyz_all,
                        mv_stream_t stream) {
  MV_REQUIRE(depth && Kinv_host && xyz_all, MV_E_ARG, "mv_geom_backproject: null pointer");
  MV_REQUIRE(H > 0 && W > 0, MV_E_ARG, "mv_geom_backproject: H and W must be positive");
  int n = H * W;
  backproject_kernel<<<(n + 255) / 256, 256, 0, mv_cuda_stream(stream)>>>(depth, H, W, load_mat3(Kinv_host),
                                                                         xyz_all);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_geom_project_coords(const float* xyz_all, const int32_t* valid_idx, const int32_t* n_dev, int n_max,
                           const float* K_host, int H, int W, int h, int w, float* xyz, float* coords,
                           mv_stream_t stream) {
  MV_REQUIRE(xyz_all && K_host && xyz && coords, MV_E_ARG, "mv_geom_project_coords: null pointer");
  MV_REQUIRE(n_max >= 0 && H > 0 && W > 0 && h > 0 && w > 0, MV_E_ARG, "mv_geom_project_coords: bad sizes");
  if (n_max == 0) return MV_OK;
  project_coords_kernel<<<(n_max + 255) / 256, 256, 0, mv_cuda_stream(stream)>>>(
      xyz_all, valid_idx, n_dev, n_max, load_mat3(K_host), H, W, h, w, xyz, coords);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_geom_grid_coords(const float* xyz_grid, const int32_t* valid_idx, const int32_t* n_dev, int n_max, int H,
                        int W, int h, int w, float* xyz, float* uv, float* coords, mv_stream_t stream) {
  MV_REQUIRE(coords && (xyz_grid || !xyz), MV_E_ARG, "mv_geom_grid_coords: null pointer");
  MV_REQUIRE(n_max >= 0 && H > 0 && W > 0 && h > 0 && w > 0, MV_E_ARG, "mv_geom_grid_coords: bad sizes");
  if (n_max == 0) return MV_OK;
  grid_coords_kernel<<<(n_max + 255) / 256, 256, 0, mv_cuda_stream(stream)>>>(xyz_grid, valid_idx, n_dev, n_max, H,
                                                                             W, h, w, xyz, uv, coords);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_geom_keypoint_coords(const float* kps, int kp_stride, int n, float image_size, int h, int w,
                            float* coords, mv_stream_t stream) {
  MV_REQUIRE(kps && coords, MV_E_ARG, "mv_geom_keypoint_coords: null pointer");
  MV_REQUIRE(n >= 0 && kp_stride >= 2 && h > 0 && w > 0 && image_size > 0.f, MV_E_ARG,
             "mv_geom_keypoint_coords: bad sizes");
  if (n == 0) return MV_OK;
  keypoint_coords_kernel<<<(n + 127) / 128, 128, 0, mv_cuda_stream(stream)>>>(kps, kp_stride, n, image_size, h, w,
                                                                            coords);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_k1_sample_normalize(int mode, const float* src, int C, int h, int w, const float* coords,
                           const int32_t* n_dev, int n_max, int normalize, uint16_t* out_bf16, float* out_f32,
                           int32_t* taps, mv_stream_t stream) {
  MV_REQUIRE(src && (out_bf16 || out_f32), MV_E_ARG, "mv_k1_sample_normalize: null src or no output");
  MV_REQUIRE(mode == MV_SAMPLE_BILINEAR_ZEROS || mode == MV_SAMPLE_BICUBIC_CLAMP || mode == MV_SAMPLE_ROWS,
             MV_E_ARG, "mv_k1_sample_normalize: unknown mode %d", mode);
  MV_REQUIRE(mode == MV_SAMPLE_ROWS || (coords && h > 0 && w > 0), MV_E_ARG,
             "mv_k1_sample_normalize: sampling modes need coords and a map size");
  MV_REQUIRE(C > 0 && C % 8 == 0 && C <= 8192, MV_E_RANGE, "mv_k1_sample_normalize: C=%d must be a multiple of 8, <= 8192", C);
  MV_REQUIRE(n_max >= 0, MV_E_ARG, "mv_k1_sample_normalize: negative n_max");
  MV_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0, MV_E_ALIGN, "mv_k1_sample_normalize: src must be 16-byte aligned");
  MV_REQUIRE(!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0, MV_E_ALIGN,
             "mv_k1_sample_normalize: out_f32 must be 16-byte aligned");
  MV_REQUIRE(!out_bf16 || (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0, MV_E_ALIGN,
             "mv_k1_sample_normalize: out_bf16 must be 8-byte aligned");
  if (n_max == 0) return MV_OK;

  K1Params p;
  p.src = src;
  p.coords = coords;
  p.n_dev = n_dev;
  p.n_max = n_max;
  p.C = C;
  p.h = h;
  p.w = w;
  p.normalize = normalize;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.out_f32 = out_f32;
  p.taps = taps;

  // thread = 4 * nv channels; one CTA = all channels of a run of consecutive points
  const int C4 = C / 4;
  const int nv = C4 >= 256 ? 4 : (C4 >= 128 ? 2 : 1);
  const int threads = (((C4 + nv - 1) / nv + 31) / 32) * 32;  // <= 512 for C <= 8192
  // long runs amortise the tap loads, a few CTAs per SM overlap their barriers; the kernel re-derives the run
  // length from the live count, the grid only has to cover n_max at the longest run (K1_MAX_RUN)
  int per_sm = 512 / threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = mv_sm_count() * per_sm;
  const int need = (n_max + K1_MAX_RUN - 1) / K1_MAX_RUN;
  if (grid < need) grid = need;
  const int most = (n_max + 1) / 2;
  if (grid > most) grid = most;
  if (grid < 1) grid = 1;
  cudaStream_t st = mv_cuda_stream(stream);
  if (mode == MV_SAMPLE_BILINEAR_ZEROS) return launch_k1<MV_SAMPLE_BILINEAR_ZEROS>(p, threads, nv, grid, st);
  if (mode == MV_SAMPLE_BICUBIC_CLAMP) return launch_k1<MV_SAMPLE_BICUBIC_CLAMP>(p, threads, nv, grid, st);
  return launch_k1<MV_SAMPLE_ROWS>(p, threads, nv, grid, st);
}

}  // extern "C"
