// k1_grid.cu -- kernel 1 for the NAVI-style side: bicubic upsampling of a (h, w, C) map by an integer factor onto the
// live pixels of an (H, W) grid, L2-normalise, write f16c rows (evals/utils/correspondence.py:240-252 + :47-48).
//
// Why a second kernel-1 form.  The point-run kernel of k1_sample.cu stages, per CTA, the source columns a run of ~17
// consecutive live pixels of ONE output row needs; at C = 3072 a column is 12 KB, so a window is ~9 columns, every
// window re-blends 4 source rows along y, and the kernel reads 128 MB from L2 to write 62 MB of rows -- its two halves
// (L2-bound window fill, HBM-bound row writes) run one after the other, 16 us each.  A 4x upsample has far more
// reuse than that: the fy output rows of one source-row step share 5 source rows, the 4 pixels of a quad share 5
// source columns.  This kernel tiles in 2-D and splits the CHANNELS over a thread-block cluster:
//
//   tile     = fy output rows x 8 quads of fx pixels (4 x 32 pixels at 4x), window = 5 x 12 source pixels
//   cluster  = C / CS CTAs (CS = 256 / 384 / 512 channels each; 8 CTAs at C = 3072), every CTA holds its channel
//              slice of the window in shared memory (92 KB at CS = 384: two CTAs per SM) -- 40 MB of L2 reads in all
//   unit     = 4 consecutive pixels of one output row: a warp blends the 5 x 4 window values along y once
//              (20 x LDS.128 per 4 channels), then along x for the 4 pixels; the rows stay in registers
//   norm     = the sum of squares (and the row . dotvec of the f16c query role) of a pixel needs all channels:
//              every CTA puts its partial sums in shared memory, barrier.cluster, and reads its peers' through
//              distributed shared memory (ld.shared::cluster) in a fixed order -> identical bits in every CTA
//   store    = scale, subtract the centre, split into fp16 hi / lo (mv_k1_sample_f16c's format), 256-byte segments
//
// The arithmetic per element is the point-run kernel's (y blend first, then x, same FMA order), so the two agree to
// the last bit except for the order in which the sum of squares is accumulated (<= 1 ulp of the norm).
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace {

constexpr float G_NORM_EPS = 1e-12f;
constexpr float G_LO_SCALE = 2048.f;
constexpr int G_THREADS = 256, G_WARPS = 8;
constexpr int G_NR = 5;       // source rows of a tile
constexpr int G_NC = 12;      // source columns of a tile: 8 quads + 4
constexpr int G_MAXCL = 8;    // channel slices = CTAs per cluster

struct GridParams {
  const float* src;       // (h*w, C)
  const int32_t* rank;    // (H*W): row index of a live pixel, -1 otherwise
  int C, h, w, H, W, fx, fy, tiles_x;
  int role, pitch;
  const float* center;
  const float* dotvec;
  __half* out_hi;
  __half* out_lo;
};

__device__ __forceinline__ void cubic4(float t, float c[4]) {  // Keys, A = -0.75 (ATen UpSample.h)
  const float A = -0.75f;
  float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x);
  a.y = fmaf(w, v.y, a.y);
  a.z = fmaf(w, v.z, a.z);
  a.w = fmaf(w, v.w, a.w);
}
__device__ __forceinline__ float dotacc4(const float4& v, const float4& d, float acc) {
  acc = fmaf(v.x, d.x, acc);
  acc = fmaf(v.y, d.y, acc);
  acc = fmaf(v.z, d.z, acc);
  return fmaf(v.w, d.w, acc);
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_dsmem(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}

struct GridShared {
  int pt[8 * 64];               // rank of the tile's pixels, [row][x] with a row pitch of 8 * fx <= 64
  float part[2][G_WARPS][4][2]; // [round parity][warp = unit of the round][pixel][ss, dot]
  float wx[G_WARPS][4][5];      // per warp: the 4 pixels' x weights over the quad's 5 columns (one of them is 0)
  float wy[G_WARPS][4];
};

// NITS = CS / 128: float4 per lane and pixel
template <int NITS>
__global__ void __launch_bounds__(G_THREADS, 2) k1_grid_kernel(GridParams p, int ncl) {
  extern __shared__ float4 gdyn[];
  float* win = reinterpret_cast<float*>(gdyn);  // [G_NR][G_NC][CS]
  __shared__ GridShared sh;
  constexpr int CS = NITS * 128;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slice = blockIdx.x % ncl, tile = blockIdx.x / ncl;
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int fx = p.fx, fy = p.fy;
  const int TW = 8 * fx;                       // tile width in pixels
  const int x_tile = tx * TW, y_tile = ty * fy;
  const int q0 = tx * 8;                       // first quad = first source column step
  const int cbase = slice * CS;

  // ---- which pixels of the tile are live (the same answer in every CTA of the cluster)
  int any = 0;
  for (int i = tid; i < fy * TW; i += G_THREADS) {
    const int r = i / TW, x = x_tile + (i - r * TW);
    const int rk = (x < p.W) ? __ldg(p.rank + (size_t)(y_tile + r) * p.W + x) : -1;
    sh.pt[r * 64 + (i - r * TW)] = rk;
    any |= (rk >= 0);
  }
  if (!__syncthreads_or(any)) return;

  // ---- window fill: source rows ty - 2 .. ty + 2, columns q0 - 2 .. q0 + 9 (border clamp), this CTA's channels
  {
    constexpr int C4 = CS / 4;
    const int items = G_NR * G_NC * C4;
#pragma unroll 4
    for (int idx = tid; idx < items; idx += G_THREADS) {
      const int cell = idx / C4, c4 = idx - cell * C4;
      const int r = cell / G_NC, s = cell - r * G_NC;
      const int yy = min(max(ty - 2 + r, 0), p.h - 1), xx = min(max(q0 - 2 + s, 0), p.w - 1);
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.src + ((size_t)yy * p.w + xx) * p.C + cbase) + c4);
      *reinterpret_cast<float4*>(win + (size_t)idx * 4) = v;
    }
  }
  __syncthreads();

  const int upr = TW / 4;            // units per output row
  const int units = fy * upr;        // a multiple of 8 (fx is 4 or 8)
  const float sx = (float)p.w / (float)p.W, sy = (float)p.h / (float)p.H;
  const bool hasdot = p.dotvec != nullptr;
  const uint32_t part_addr = sm100::smem_u32(&sh.part[0][0][0][0]);

  for (int round = 0, parity = 0; round * G_WARPS < units; ++round, parity ^= 1) {
    const int u = round * G_WARPS + wid;
    const int j = u / upr, xq = u - j * upr;        // output row inside the tile, unit inside the row
    const int xl = 4 * xq;                          // first pixel of the unit, tile-local
    int rk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) rk[e] = sh.pt[j * 64 + xl + e];
    const bool live = (rk[0] >= 0) | (rk[1] >= 0) | (rk[2] >= 0) | (rk[3] >= 0);
    float4 acc[4][NITS];
    float ss[4] = {0.f, 0.f, 0.f, 0.f}, dd[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {  // warp-uniform
      // per-pixel scalars, the formulas (and roundings) of mv_geom_grid_coords + the point-run kernel
      const int y = y_tile + j;
      const float iy = __fsub_rn(__fmul_rn(sy, (float)y + 0.5f), 0.5f);
      const float fyf = floorf(iy);
      const int r0 = (int)fyf - 1 - (ty - 2);       // window row of the first y tap: 0 or 1
      const int quad = (x_tile + xl) / fx;          // source column step of the unit
      const int s0 = quad - q0;                     // window column of the quad's first (of 5) columns
      if (lane < 4) {
        const int x = x_tile + xl + lane;
        const float ix = __fsub_rn(__fmul_rn(sx, (float)x + 0.5f), 0.5f);
        const float fxf = floorf(ix);
        float c[4];
        cubic4(ix - fxf, c);
        const int off = (int)fxf - 1 - (quad - 2);  // 0 or 1: where this pixel's 4 taps start inside the 5 columns
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const int k = s - off;
          sh.wx[wid][lane][s] = (k >= 0 && k < 4) ? c[k] : 0.f;
        }
      }
      if (lane == 4) {
        float c[4];
        cubic4(iy - fyf, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) sh.wy[wid][k] = c[k];
      }
      __syncwarp();
      const float cy0 = sh.wy[wid][0], cy1 = sh.wy[wid][1], cy2 = sh.wy[wid][2], cy3 = sh.wy[wid][3];
      const float* wrow = win + ((size_t)r0 * G_NC + s0) * CS;
#pragma unroll
      for (int it = 0; it < NITS; ++it) {
        const int c = (it * 32 + lane) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e][it] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const float* q = wrow + (size_t)s * CS + c;
          const float4 a = *reinterpret_cast<const float4*>(q), b = *reinterpret_cast<const float4*>(q + G_NC * CS),
                       cc = *reinterpret_cast<const float4*>(q + 2 * G_NC * CS), d = *reinterpret_cast<const float4*>(q + 3 * G_NC * CS);
          float4 yb = make_float4(0.f, 0.f, 0.f, 0.f);  // the four source rows blended along y, once per column
          fma4(yb, cy0, a);
          fma4(yb, cy1, b);
          fma4(yb, cy2, cc);
          fma4(yb, cy3, d);
#pragma unroll
          for (int e = 0; e < 4; ++e) fma4(acc[e][it], sh.wx[wid][e][s], yb);  // a zero weight leaves the sum untouched
        }
        const float4 dv = hasdot ? __ldg(reinterpret_cast<const float4*>(p.dotvec + cbase + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ss[e] = dotacc4(acc[e][it], acc[e][it], ss[e]);
          if (hasdot) dd[e] = dotacc4(acc[e][it], dv, dd[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ss[e] = warp_sum(ss[e]);
        if (hasdot) dd[e] = warp_sum(dd[e]);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sh.part[parity][wid][e][0] = ss[e];
        sh.part[parity][wid][e][1] = dd[e];
      }
    }
    // every CTA of the cluster has written its partial sums of this round
    cluster_arrive();
    cluster_wait();
    if (live) {
      // lane l reads partial (cta = l >> 2, pixel = l & 3) of every peer; totals in CTA order: the same bits everywhere
      float vs = 0.f, vd = 0.f;
      if ((lane >> 2) < ncl) {
        const uint32_t a = sm100::map_to_cta(part_addr + (uint32_t)(((parity * G_WARPS + wid) * 4 + (lane & 3)) * 2) * 4u, (uint32_t)(lane >> 2));
        vs = ld_dsmem(a);
        vd = ld_dsmem(a + 4u);
      }
      float inv[4], rr[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float ts = 0.f, td = 0.f;
        for (int cta = 0; cta < ncl; ++cta) {
          ts += __shfl_sync(0xffffffffu, vs, cta * 4 + e);
          td += __shfl_sync(0xffffffffu, vd, cta * 4 + e);
        }
        inv[e] = __frcp_rn(fmaxf(sqrtf(ts), G_NORM_EPS));
        rr[e] = td * inv[e];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (rk[e] < 0) continue;
        const size_t row = (size_t)rk[e];
#pragma unroll
        for (int it = 0; it < NITS; ++it) {
          const int c = cbase + (it * 32 + lane) * 4;
          float4 o = acc[e][it];
          o.x *= inv[e];
          o.y *= inv[e];
          o.z *= inv[e];
          o.w *= inv[e];
          if (p.center) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(p.center + c));
            o.x -= m.x;
            o.y -= m.y;
            o.z -= m.z;
            o.w -= m.w;
          }
          const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&h0);
          pk.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(p.out_hi + row * p.pitch + c) = pk;
          if (p.out_lo) {
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn((o.x - f0.x) * G_LO_SCALE, (o.y - f0.y) * G_LO_SCALE);
            const __half2 l1 = __floats2half2_rn((o.z - f1.x) * G_LO_SCALE, (o.w - f1.y) * G_LO_SCALE);
            pk.x = *reinterpret_cast<const uint32_t*>(&l0);
            pk.y = *reinterpret_cast<const uint32_t*>(&l1);
            *reinterpret_cast<uint2*>(p.out_lo + row * p.C + c) = pk;
          }
        }
        if (slice == 0 && lane == 0) {  // the 8 augmentation columns of the f16c row
          __half a[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] = __float2half_rn(0.f);
          if (p.role == MV_ROLE_TARGET) {
            a[0] = __float2half_rn(1.f);
            a[1] = __float2half_rn(1.f);
            a[2] = __float2half_rn(1.f / G_LO_SCALE);
          } else {
            a[0] = __float2half_rn(rr[e]);
            const float r1 = rr[e] - __half2float(a[0]);
            a[1] = __float2half_rn(r1);
            a[2] = __float2half_rn((r1 - __half2float(a[1])) * G_LO_SCALE);
          }
          *reinterpret_cast<uint4*>(p.out_hi + row * p.pitch + p.C) = *reinterpret_cast<const uint4*>(a);
        }
      }
    }
  }
  // a CTA must not exit while a peer may still read its partial sums
  cluster_arrive();
  cluster_wait();
}

__global__ void rank_of_valid_kernel(const int32_t* __restrict__ valid_idx, const int32_t* __restrict__ n_dev, int n_max,
                                     int32_t* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  if (i < n) rank[valid_idx[i]] = i;
}

int pick_slices(int C, int* cs) {
  // channel slice per CTA: a multiple of 128 (one float4 per lane), at most 8 slices
  const int cand[3] = {384, 512, 256};
  for (int k = 0; k < 3; ++k)
    if (C % cand[k] == 0 && C / cand[k] <= G_MAXCL) {
      *cs = cand[k];
      return C / cand[k];
    }
  return 0;
}

template <int NITS>
int launch_grid(const GridParams& p, int ncl, int tiles, cudaStream_t st) {
  auto kern = k1_grid_kernel<NITS>;
  const size_t smem = (size_t)G_NR * G_NC * NITS * 128 * sizeof(float);
  static bool done[MV_MAX_DEVICES];
  bool& attr_done = done[mv_device_slot()];
  if (!attr_done) {
    MV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * ncl));
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ncl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MV_CUDA(cudaLaunchKernelEx(&cfg, kern, p, ncl));
  return MV_OK;
}

}  // namespace

extern "C" {

int mv_k1_grid_supported(int C, int h, int w, int H, int W) {
  int cs = 0;
  if (h <= 0 || w <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
  if (W % w != 0 || H % h != 0) return 0;
  const int fx = W / w, fy = H / h;
  if (!(fx == 4 || fx == 8) || fy < 1 || fy > 8) return 0;
  return pick_slices(C, &cs) > 0 ? 1 : 0;
}

int mv_rank_of_valid(const int32_t* valid_idx, const int32_t* n_dev, int n_max, int32_t* rank, int n_pixels, mv_stream_t stream) {
  MV_REQUIRE(valid_idx && rank, MV_E_ARG, "mv_rank_of_valid: null pointer");
  MV_REQUIRE(n_max >= 0 && n_pixels >= n_max, MV_E_ARG, "mv_rank_of_valid: n_pixels must be >= n_max >= 0");
  cudaStream_t st = mv_cuda_stream(stream);
  MV_CUDA(cudaMemsetAsync(rank, 0xff, (size_t)n_pixels * sizeof(int32_t), st));
  if (n_max == 0) return MV_OK;
  rank_of_valid_kernel<<<(n_max + 255) / 256, 256, 0, st>>>(valid_idx, n_dev, n_max, rank);
  MV_LAUNCH_CHECK();
  return MV_OK;
}

int mv_k1_grid_f16c(const float* src, int C, int h, int w, int H, int W, const int32_t* rank, int role, const float* center,
                    const float* dotvec, uint16_t* out_f16, int pitch, uint16_t* out_f16_lo, mv_stream_t stream) {
  MV_REQUIRE(src && rank && out_f16, MV_E_ARG, "mv_k1_grid_f16c: null pointer");
  MV_REQUIRE(mv_k1_grid_supported(C, h, w, H, W), MV_E_RANGE,
             "mv_k1_grid_f16c: needs an integer upsampling factor of 4 or 8 along x, 1..8 along y, and C divisible into <= 8 slices of 256 / 384 / 512 channels");
  MV_REQUIRE(role == MV_ROLE_QUERY || role == MV_ROLE_TARGET, MV_E_ARG, "mv_k1_grid_f16c: unknown role %d", role);
  MV_REQUIRE(pitch >= C + 8 && pitch % 8 == 0, MV_E_ALIGN, "mv_k1_grid_f16c: pitch %d must be >= C + 8 and a multiple of 8", pitch);
  MV_REQUIRE((((uintptr_t)src | (uintptr_t)out_f16 | (uintptr_t)center | (uintptr_t)dotvec) & 15) == 0 && ((uintptr_t)out_f16_lo & 7) == 0,
             MV_E_ALIGN, "mv_k1_grid_f16c: src, out_f16, center, dotvec must be 16-byte aligned (out_f16_lo: 8)");
  GridParams p;
  p.src = src;
  p.rank = rank;
  p.C = C;
  p.h = h;
  p.w = w;
  p.H = H;
  p.W = W;
  p.fx = W / w;
  p.fy = H / h;
  p.tiles_x = (W + 8 * p.fx - 1) / (8 * p.fx);
  p.role = role;
  p.pitch = pitch;
  p.center = center;
  p.dotvec = dotvec;
  p.out_hi = reinterpret_cast<__half*>(out_f16);
  p.out_lo = reinterpret_cast<__half*>(out_f16_lo);
  int cs = 0;
  const int ncl = pick_slices(C, &cs);
  const int tiles = p.tiles_x * h;  // one tile row per source row step
  cudaStream_t st = mv_cuda_stream(stream);
  if (cs == 256) return launch_grid<2>(p, ncl, tiles, st);
  if (cs == 384) return launch_grid<3>(p, ncl, tiles, st);
  return launch_grid<4>(p, ncl, tiles, st);
}

}  // extern "C"
