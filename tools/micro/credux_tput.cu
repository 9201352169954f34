// Throughput of warp-reduce primitives on sm_100a: CREDUX.MAX.F32 (redux.sync.max.f32), SHFL, VOTE.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o credux_tput credux_tput.cu && ./credux_tput
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void bench(float* out, long long* cyc, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 0.37f + i;
  float acc = 0.f;
  unsigned bacc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) {
        float m;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v[i]));
        acc += m;
      } else if (OP == 1) {
        acc += __shfl_xor_sync(0xffffffffu, v[i], 16);
      } else if (OP == 2) {
        bacc += __ballot_sync(0xffffffffu, v[i] > acc);
      } else {
        unsigned m;
        asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(__float_as_uint(v[i])));
        bacc += m;
      }
      v[i] += 1.0f;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + bacc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 2000;
  const char* names[4] = {"CREDUX.MAX.F32", "SHFL.BFLY", "VOTE.ballot", "REDUX.MAX.U32"};
  for (int warps : {1, 4, 8, 16, 32}) {
    for (int op = 0; op < 4; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) bench<0><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 1) bench<1><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 2) bench<2><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 3) bench<3><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = (double)h[0];
      printf("warps/SM %2d  %-16s %7.2f cycles per warp-instruction per SM (%.2f per warp)\n", warps, names[op],
             c / (iters * 8.0 * warps), c / (iters * 8.0));
    }
  }
  return 0;
}
