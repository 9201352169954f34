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
// The product only PROPOSES the two candidates of a row; kernel 3 recomputes their fp32 distances.
//
// The exact route (mv_lr_gram_exact + mv_k3_ratio_mutual_lr): with the Gram matrix of the RAW source rows in fp32 (CUDA
// cores, blocked accumulation: 16 slices of C / 16 channels per entry, fixed-order tree -- ~1 ulp, like kernel 3's own dot
// products) the fp32 distances of the two candidates are
//     1 - (sum_{a,b} W0[i,a] W1[j,b] G01[s_a,t_b]) / (|x_i| |y_j|),   |x_i|^2 = sum_{a,a'} W0[i,a] W0[i,a'] G00[s_a,s_a']
// i.e. 16 (bilinear) / 256 (bicubic) terms per candidate instead of three C-long rows: the interpolated rows are never
// materialised (no kernel 1, no row planes) -- SURVEY.md 8(f).2 taken to its end for the shapes where h*w << C.
// (The tensor-core Gram cannot serve here: tcgen05 accumulates with truncation, measured -2.4e-6 .. -6.9e-6 of bias over
// K = 2048-3072, tools/gram_probe.py.)  The builders then read the same raw Gram (tap scale 1, column scale 1 / |src1[t]|).
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
  const float* tapscale;   // (hw) or NULL (= 1): factor of a tap's blend weight in the basis G is written in -- |src[p]| of THIS
                           // image for the cosine Gram of unit rows, NULL for the Gram of the raw rows
  const float* colscale;   // query rows: (hw) or NULL (= 1) factor of column t -- 1 / |src1[t]| for the raw Gram.
                           // target rows: (hw) |src1[p]|, the factor of the values written (B[j,t] = w |src1[t]| / |y_j|)
  const float* G;          // stacked Gram matrix of the source rows of both images, fp32, row pitch ld
  int ld, off_own, off_tgt;  // row / column offset of this image's and of the target image's source pixels in G
  __half* out;             // (n, pitch) fp16 operand rows
  int pitch;
  float* inv_out;          // (n) or NULL: 1 / max(|x_i|, eps) of every point (for mv_k3_ratio_mutual_lr)
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
  const float V = idx >= 0 ? (p.tapscale ? wt * __ldg(p.tapscale + idx) : wt) : 0.f;
  const float inv = inv_norm_from_gram<T>(idx, V, p.G, p.ld, p.off_own, lane);
  if (p.inv_out && lane == 0) p.inv_out[j] = inv;
  const float Vn = idx >= 0 ? wt * __ldg(p.colscale + idx) * inv : 0.f;  // w |src1[t]| / |y_j|
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
  const float V = idx >= 0 ? (p.tapscale ? wt * __ldg(p.tapscale + idx) : wt) : 0.f;
  const float inv = inv_norm_from_gram<T>(idx, V, p.G, p.ld, p.off_own, lane);
  if (p.inv_out && lane == 0) p.inv_out[i] = inv;
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
  if (p.colscale) {
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int c2 = lane + 32 * k;
      if (2 * c2 < p.hwp) {
        const float2 cs = __ldg(reinterpret_cast<const float2*>(p.colscale) + c2);
        acc[k].x *= cs.x;
        acc[k].y *= cs.y;
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

// ---- exact fp32 Gram of the raw source rows of both images stacked (image 1 at row / column offset hwp) ----------
// One CTA per 32 x 32 tile of the upper block triangle (mirrored on the way out), 16 warps: warp w accumulates the
// channels [w C/16, (w+1) C/16) of all 1024 entries (lane = 4 rows x 8 columns), the 16 partial sums meet in shared memory
// and are added in a fixed tree.  Every entry is 16 chains of C/16 fused multiply-adds + 4 tree levels: ~1 ulp.
// Each warp stages its slice in 32-channel chunks through its own 8 KB of shared memory (coalesced 128-byte row reads,
// 16-byte granules XOR-swizzled by the lane row / column group so that the 12 LDS.128 of a 4-channel step are conflict
// free): 12 shared loads per 128 FMAs.  (First form, loads straight from global memory: 8 different cache lines per warp
// load, L1-bound -- 95 us for the ScanNet shape under ncu against ~25 us now.)
constexpr int GX_TILE = 32, GX_WARPS = 16, GX_THREADS = GX_WARPS * 32, GX_KC = 32;
constexpr int GX_STAGE_FLOATS = 2 * GX_TILE * GX_KC;                 // A rows + B rows of one chunk, per warp
constexpr int GX_SMEM_BYTES = GX_WARPS * GX_STAGE_FLOATS * 4;        // 128 KB; the 64 KB reduction buffer aliases it afterwards

// Tiles computed: every tile of the cross block (image 0 rows x image 1 columns) and, of the two same-image blocks, the
// diagonal band |ti - tj| <= band only -- a point's taps lie within `reach` source pixels of each other, so the norms never
// read further from the diagonal (ScanNet shape: 100 + 2 x 19 = 138 tiles = one wave of CTAs instead of 190).
__global__ void __launch_bounds__(GX_THREADS, 1) lr_gram_exact_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                                                      int C, int hw, int off1, float* __restrict__ G, int ld, int nti,
                                                                      int band) {
  extern __shared__ __align__(16) float gx_smem[];
  __shared__ const float* rowp[2 * GX_TILE];
  // block index -> tile pair (ti <= tj) in units of 32 stacked rows; image 1 starts at tile nti
  int ti, tj;
  {
    int b = blockIdx.x;
    if (b < nti * nti) {  // cross block
      ti = b / nti;
      tj = nti + b % nti;
    } else {
      b -= nti * nti;
      const int per_img = (band + 1) * nti - band * (band + 1) / 2;  // sum_{d <= band} (nti - d)
      const int img = b / per_img;
      b -= img * per_img;
      int d = 0;
      while (b >= nti - d) {
        b -= nti - d;
        ++d;
      }
      ti = img * nti + b;
      tj = ti + d;
    }
  }
  const int hwp = off1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ly = lane >> 2, lx = lane & 3;
  if (threadIdx.x < 2 * GX_TILE) {  // stacked row r -> its C floats, or NULL for a pad row
    const int r = (threadIdx.x < GX_TILE ? ti * GX_TILE : tj * GX_TILE - GX_TILE) + threadIdx.x;
    const float* q = nullptr;
    if (r < hw) q = src0 + (size_t)r * C;
    else if (r >= hwp && r < hwp + hw) q = src1 + (size_t)(r - hwp) * C;
    rowp[threadIdx.x] = q;
  }
  __syncthreads();
  const int slice = C / GX_WARPS, k_beg = wid * slice;
  float* stA = gx_smem + (size_t)wid * GX_STAGE_FLOATS;
  float* stB = stA + GX_TILE * GX_KC;
  float acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = k_beg; k0 < k_beg + slice; k0 += GX_KC) {
    const int kc4 = min(GX_KC, k_beg + slice - k0) >> 2;  // granules that carry data (a slice need not be a multiple of 32)
    // stage 64 rows x 32 channels: 512 granules of 16 bytes, 16 per lane; 8 consecutive lanes read one row's 128 bytes
    float4 stage[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {  // all 16 loads in flight before the first store
      const int idx = t * 32 + lane, row = idx >> 3, g = idx & 7;
      const float* q = rowp[row];
      stage[t] = (q && g < kc4) ? __ldg(reinterpret_cast<const float4*>(q + k0) + g) : z4;
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int idx = t * 32 + lane, row = idx >> 3, g = idx & 7;
      const int rr = row & 31;
      const int swz = row < GX_TILE ? (rr >> 2) : (rr >> 3);  // A rows by their lane row group, B rows by their column group
      float* dst = (row < GX_TILE ? stA : stB) + rr * GX_KC + 4 * (g ^ swz);
      *reinterpret_cast<float4*>(dst) = stage[t];
    }
    __syncwarp();
#pragma unroll 2
    for (int g = 0; g < GX_KC / 4; ++g) {
      float4 a[4], b[8];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(stA + (4 * ly + r) * GX_KC + 4 * (g ^ ly));
#pragma unroll
      for (int c = 0; c < 8; ++c) b[c] = *reinterpret_cast<const float4*>(stB + (8 * lx + c) * GX_KC + 4 * (g ^ lx));
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[r][c] = fmaf(a[r].x, b[c].x, acc[r][c]);
          acc[r][c] = fmaf(a[r].y, b[c].y, acc[r][c]);
          acc[r][c] = fmaf(a[r].z, b[c].z, acc[r][c]);
          acc[r][c] = fmaf(a[r].w, b[c].w, acc[r][c]);
        }
    }
    __syncwarp();
  }
  __syncthreads();  // every warp is done with its staging area: the reduction buffer takes its place
  float* gx_red = gx_smem;  // [warp][row][col]
  float* mine = gx_red + (size_t)wid * GX_TILE * GX_TILE;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float4* q = reinterpret_cast<float4*>(mine + (4 * ly + r) * GX_TILE + 8 * lx);
    q[0] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    q[1] = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
  __syncthreads();
  const int rows = 2 * nti * GX_TILE;
  for (int o = threadIdx.x; o < GX_TILE * GX_TILE; o += GX_THREADS) {
    float v[GX_WARPS];
#pragma unroll
    for (int q = 0; q < GX_WARPS; ++q) v[q] = gx_red[(size_t)q * GX_TILE * GX_TILE + o];
#pragma unroll
    for (int st = 1; st < GX_WARPS; st <<= 1)  // fixed tree: (0+1), (2+3), ... then pairs of pairs
#pragma unroll
      for (int q = 0; q < GX_WARPS; q += 2 * st) v[q] += v[q + st];
    const int gi = ti * GX_TILE + o / GX_TILE, gj = tj * GX_TILE + o % GX_TILE;
    if (gi < rows && gj < rows) {
      G[(size_t)gi * ld + gj] = v[0];
      G[(size_t)gj * ld + gi] = v[0];
    }
  }
}

// |src[p]| and 1 / |src[p]| of every stacked source row from the diagonal of the raw Gram (0 for pad / zero rows)
__global__ void lr_diag_norms_kernel(const float* __restrict__ G, int ld, int rows, float* __restrict__ snorm, float* __restrict__ rsnorm) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float nrm = sqrtf(fmaxf(G[(size_t)r * ld + r], 0.f));
  snorm[r] = nrm;
  rsnorm[r] = nrm > 0.f ? 1.f / nrm : 0.f;
}

// ---- kernel 3 on the Gram matrix: the fp32 cosine distances of a query's two candidates from 16 / 256 entries of G01,
// then exactly kernel 3's tail (k3_score.cu): fp32 order wins, ratio weight with both clamps, mutual flag.  One warp per query.
constexpr float LR_RATIO_CLAMP = 1e-9f;   // calculate_ratio_test clamps (correspondence.py:105-121)
constexpr float LR_MISSING_DIST = 2.0f;   // a candidate that does not exist (m < 2), as in k3_score.cu

struct LrK3Params {
  const float* coords_q;
  const float* coords_t;
  const float* inv_q;  // 1 / max(|x_i|, eps) from the builders
  const float* inv_t;
  const float* G;
  int ld, off_q, off_t, h, w;
  const int32_t* n_dev;
  int n_max;
  int32_t* row_idx;
  const unsigned long long* col_best;
  int ratio_test;
  float* dists;
  float* weight;
  uint8_t* mutual;
};

template <int MODE>
__global__ void __launch_bounds__(256) lr_k3_ratio_mutual_kernel(LrK3Params p) {
  constexpr int T = MODE == MV_SAMPLE_BICUBIC_CLAMP ? 16 : 4;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = p.n_dev ? min(*p.n_dev, p.n_max) : p.n_max;
  if (i >= n) return;
  const float2 xy = __ldg(reinterpret_cast<const float2*>(p.coords_q) + i);
  int iq;
  float wq;
  lane_tap<MODE>(xy.x, xy.y, p.h, p.w, lane, iq, wq);
  const float inv_i = __ldg(p.inv_q + i);
  int js[2] = {p.row_idx[2 * (size_t)i], p.row_idx[2 * (size_t)i + 1]};
  float d[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int j = js[c];
    if (j < 0) {  // warp-uniform
      d[c] = LR_MISSING_DIST;
      continue;
    }
    const float2 uv = __ldg(reinterpret_cast<const float2*>(p.coords_t) + j);
    int it;
    float wt;
    lane_tap<MODE>(uv.x, uv.y, p.h, p.w, lane, it, wt);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < (T * T + 31) / 32; ++k) {
      const int q = lane + 32 * k;
      const int a = (q / T) % T, b = q % T;
      const float wa = __shfl_sync(0xffffffffu, wq, a), wb = __shfl_sync(0xffffffffu, wt, b);
      const int ia = __shfl_sync(0xffffffffu, iq, a), ib = __shfl_sync(0xffffffffu, it, b);
      if (q < T * T && ia >= 0 && ib >= 0) acc = fmaf(wa * wb, __ldg(p.G + (size_t)(p.off_q + ia) * p.ld + p.off_t + ib), acc);
    }
    acc = warp_sum(acc);
    d[c] = 1.f - (acc * inv_i) * __ldg(p.inv_t + j);
  }
  if (lane != 0) return;
  float d0 = d[0], d1 = d[1];
  int j0 = js[0], j1 = js[1];
  if (j1 >= 0 && (d1 < d0 || (d1 == d0 && j1 < j0))) {  // fp32 order wins over the tensor-core order
    const float td = d0; d0 = d1; d1 = td;
    const int tj = j0; j0 = j1; j1 = tj;
    p.row_idx[2 * (size_t)i] = j0;
    p.row_idx[2 * (size_t)i + 1] = j1;
  }
  if (p.dists) {
    p.dists[2 * (size_t)i] = d0;
    p.dists[2 * (size_t)i + 1] = d1;
  }
  if (p.weight) {
    float wv = d0;
    if (p.ratio_test) wv = 1.f - __fdiv_rn(fmaxf(d0, LR_RATIO_CLAMP), fmaxf(fmaxf(d1, LR_RATIO_CLAMP), LR_RATIO_CLAMP));
    p.weight[i] = wv;
  }
  if (p.mutual) {
    uint8_t f = 0;
    if (p.col_best && j0 >= 0) {
      const unsigned long long pk = p.col_best[j0];
      f = (pk != 0ull && (0xffffffffu - (uint32_t)(pk & 0xffffffffull)) == (uint32_t)i) ? 1 : 0;
    }
    p.mutual[i] = f;
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

int check_params(const char* who, int mode, const float* coords, int n_max, int h, int w, const float* G,
                 int ld_g, int off_own, int off_tgt, const void* out, int pitch, int hwp) {
  MV_REQUIRE(mode == MV_SAMPLE_BILINEAR_ZEROS || mode == MV_SAMPLE_BICUBIC_CLAMP, MV_E_ARG, "%s: mode must be bilinear-zeros or bicubic-clamp", who);
  MV_REQUIRE(coords && G && out, MV_E_ARG, "%s: null pointer", who);
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

int mv_lr_build_query(int mode, const float* coords, const int32_t* n_dev, int n_max, int h, int w, const float* tapscale,
                      const float* colscale, const float* G, int ld_g, int off_own, int off_tgt, void* A_f16, int pitch, int hwp,
                      float* inv_norm_out, mv_stream_t stream) {
  int rc = check_params("mv_lr_build_query", mode, coords, n_max, h, w, G, ld_g, off_own, off_tgt, A_f16, pitch, hwp);
  if (rc) return rc;
  MV_REQUIRE(!colscale || ((uintptr_t)colscale & 7) == 0, MV_E_ALIGN, "mv_lr_build_query: colscale must be 8-byte aligned");
  LrParams p{coords, n_dev, n_max, h, w, h * w, hwp, tapscale, colscale, G, ld_g, off_own, off_tgt, reinterpret_cast<__half*>(A_f16),
             pitch, inv_norm_out};
  return mode == MV_SAMPLE_BICUBIC_CLAMP ? launch_query<MV_SAMPLE_BICUBIC_CLAMP>(p, mv_cuda_stream(stream))
                                         : launch_query<MV_SAMPLE_BILINEAR_ZEROS>(p, mv_cuda_stream(stream));
}

int mv_lr_build_target(int mode, const float* coords, const int32_t* n_dev, int n_max, int h, int w, const float* tapscale,
                       const float* snorm, const float* G, int ld_g, int off_own, void* B_f16, int pitch, int hwp,
                       float* inv_norm_out, mv_stream_t stream) {
  int rc = check_params("mv_lr_build_target", mode, coords, n_max, h, w, G, ld_g, off_own, off_own, B_f16, pitch, hwp);
  if (rc) return rc;
  MV_REQUIRE(snorm, MV_E_ARG, "mv_lr_build_target: null snorm");
  LrParams p{coords, n_dev, n_max, h, w, h * w, hwp, tapscale, snorm, G, ld_g, off_own, off_own, reinterpret_cast<__half*>(B_f16),
             pitch, inv_norm_out};
  const int grid = (n_max + 7) / 8;
  if (mode == MV_SAMPLE_BICUBIC_CLAMP) lr_build_target_kernel<MV_SAMPLE_BICUBIC_CLAMP><<<grid, 256, 0, mv_cuda_stream(stream)>>>(p);
  else lr_build_target_kernel<MV_SAMPLE_BILINEAR_ZEROS><<<grid, 256, 0, mv_cuda_stream(stream)>>>(p);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_lr_gram_exact(const float* src0_hwc, const float* src1_hwc, int C, int hw, int off1, int reach, float* G, int ld_g,
                     float* snorm, float* rsnorm, mv_stream_t stream) {
  MV_REQUIRE(src0_hwc && src1_hwc && G && snorm && rsnorm, MV_E_ARG, "mv_lr_gram_exact: null pointer");
  MV_REQUIRE(C > 0 && C % 64 == 0, MV_E_ALIGN, "mv_lr_gram_exact: C=%d must be a positive multiple of 64", C);
  MV_REQUIRE(hw > 0 && off1 >= hw && off1 % GX_TILE == 0 && off1 <= MV_LR_MAX_SOURCE_PIXELS && ld_g >= 2 * off1 && reach >= 0, MV_E_RANGE,
             "mv_lr_gram_exact: need hw <= off1 <= %d, off1 a multiple of %d, ld_g >= 2 off1, reach >= 0", MV_LR_MAX_SOURCE_PIXELS, GX_TILE);
  MV_REQUIRE((((uintptr_t)src0_hwc | (uintptr_t)src1_hwc) & 15) == 0, MV_E_ALIGN, "mv_lr_gram_exact: maps must be 16-byte aligned");
  static bool done[MV_MAX_DEVICES];
  bool& attr_done = done[mv_device_slot()];
  if (!attr_done) {
    MV_CUDA(cudaFuncSetAttribute(lr_gram_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GX_SMEM_BYTES));
    attr_done = true;
  }
  const int nti = off1 / GX_TILE;
  int band = reach / GX_TILE + 1;  // indices within `reach` of each other lie at most this many tiles apart
  if (band > nti - 1) band = nti - 1;
  const int per_img = (band + 1) * nti - band * (band + 1) / 2;
  cudaStream_t st = mv_cuda_stream(stream);
  lr_gram_exact_kernel<<<nti * nti + 2 * per_img, GX_THREADS, GX_SMEM_BYTES, st>>>(src0_hwc, src1_hwc, C, hw, off1, G, ld_g, nti, band);
  MV_LAUNCH_CHECK();
  lr_diag_norms_kernel<<<(2 * off1 + 255) / 256, 256, 0, st>>>(G, ld_g, 2 * off1, snorm, rsnorm);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_k3_ratio_mutual_lr(int mode, const float* coords_q, const float* coords_t, const float* inv_q, const float* inv_t,
                          const float* G, int ld_g, int off_q, int off_t, int h, int w, const int32_t* n_dev, int n_max,
                          int32_t* row_idx, const unsigned long long* col_best, int ratio_test, float* dists, float* weight,
                          uint8_t* mutual, mv_stream_t stream) {
  MV_REQUIRE(mode == MV_SAMPLE_BILINEAR_ZEROS || mode == MV_SAMPLE_BICUBIC_CLAMP, MV_E_ARG, "mv_k3_ratio_mutual_lr: bad mode");
  MV_REQUIRE(coords_q && coords_t && inv_q && inv_t && G && row_idx, MV_E_ARG, "mv_k3_ratio_mutual_lr: null pointer");
  MV_REQUIRE(n_max >= 0 && h > 0 && w > 0 && ld_g > 0 && off_q >= 0 && off_t >= 0, MV_E_ARG, "mv_k3_ratio_mutual_lr: bad sizes");
  if (n_max == 0) return MV_OK;
  LrK3Params p{coords_q, coords_t, inv_q, inv_t, G, ld_g, off_q, off_t, h, w, n_dev, n_max, row_idx, col_best, ratio_test, dists, weight, mutual};
  const int grid = (n_max + 7) / 8;
  if (mode == MV_SAMPLE_BICUBIC_CLAMP) lr_k3_ratio_mutual_kernel<MV_SAMPLE_BICUBIC_CLAMP><<<grid, 256, 0, mv_cuda_stream(stream)>>>(p);
  else lr_k3_ratio_mutual_kernel<MV_SAMPLE_BILINEAR_ZEROS><<<grid, 256, 0, mv_cuda_stream(stream)>>>(p);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

}  // extern "C"
