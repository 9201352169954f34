// Issue interval of the legacy warp-level MMAs on sm_100a: HMMA.1688.F32.TF32 (mma.sync m16n8k8 tf32) and
// HMMA.16816.F32.BF16 (m16n8k16 bf16), with NACC independent accumulators per warp and W warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_tput hmma_tput.cu && ./hmma_tput
// (what bounds step H of the SPair streaming kernel: 18 HMMA.1688.TF32 per K step and warp, 6 accumulators)
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NACC>
__global__ void bench(float* out, long long* cyc, int iters) {
  float d[NACC][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(threadIdx.x * 0.001f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(threadIdx.x * 0.002f + i);
  for (int n = 0; n < NACC; ++n)
    for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[n][0]), "+f"(d[n][1]), "+f"(d[n][2]), "+f"(d[n][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[n][0]), "+f"(d[n][1]), "+f"(d[n][2]), "+f"(d[n][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0.f;
  for (int n = 0; n < NACC; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int NACC>
void run(const char* name, int warps_per_cta, float* out, long long* cyc) {
  const int iters = 2000, grid = 148;
  bench<KIND, NACC><<<grid, 32 * warps_per_cta>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  bench<KIND, NACC><<<grid, 32 * warps_per_cta>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < grid; ++i) c += h[i];
  c /= grid;
  const double per_smsp = (double)iters * NACC * warps_per_cta / 4.0;  // HMMAs one sub-partition issued
  printf("%-22s acc/warp %d  warps/SM %2d : %7.2f cycles per HMMA and sub-partition (%.1f per warp)\n", name, NACC, warps_per_cta,
         c / per_smsp, c / ((double)iters * NACC));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0, 1>("HMMA.1688.F32.TF32", 4, out, cyc);
  run<0, 3>("HMMA.1688.F32.TF32", 4, out, cyc);
  run<0, 6>("HMMA.1688.F32.TF32", 4, out, cyc);
  run<0, 6>("HMMA.1688.F32.TF32", 8, out, cyc);
  run<0, 6>("HMMA.1688.F32.TF32", 16, out, cyc);
  run<1, 1>("HMMA.16816.F32.BF16", 4, out, cyc);
  run<1, 6>("HMMA.16816.F32.BF16", 4, out, cyc);
  run<1, 6>("HMMA.16816.F32.BF16", 16, out, cyc);
  return 0;
}
