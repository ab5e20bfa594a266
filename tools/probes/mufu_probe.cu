// mufu_probe.cu — ex2.approx.ftz.f32 throughput per SM vs number of resident warps (B200).
#include <cstdio>
__global__ void probe(int iters, float seed, float* out, long long* clk) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + threadIdx.x * 1e-3f + i * 0.01f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    probe<<<148, warps * 32>>>(iters, -3.f, out, clk);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double per_sm_per_clk = (double)warps * 32 * 16 * iters / h;
    printf("warps/SM %2d: %6.2f ex2 per clk per SM  (%.1f clk per warp instruction per scheduler-warp)\n", warps, per_sm_per_clk,
           (double)h / (16.0 * iters));
  }
  return 0;
}
