// mufu_probe.cu — ex2 throughput per SM vs resident warps (B200): fp32 and packed bf16x2 / f16x2 forms.
#include <cstdio>
template <int MODE>  // 0: ex2.approx.ftz.f32   1: ex2.approx.ftz.bf16x2   2: ex2.approx.f16x2
__global__ void probe(int iters, float seed, float* out, long long* clk) {
  unsigned x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = MODE == 0 ? __float_as_uint(seed + threadIdx.x * 1e-3f + i * 0.01f) : 0xBF80BF00u + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* clk) {
  const int iters = 2000;
  for (int warps : {4, 8}) {
    probe<MODE><<<148, warps * 32>>>(iters, -3.f, out, clk);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    const int per = MODE == 0 ? 1 : 2;
    printf("%-22s warps/SM %2d: %6.2f results per clk per SM (%5.1f clk per warp instruction)\n", name, warps,
           (double)warps * 32 * 16 * iters * per / h, (double)h / (16.0 * iters) / (warps / 4.0));
  }
}
int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  run<0>("ex2.approx.ftz.f32", out, clk);
  run<1>("ex2.approx.ftz.bf16x2", out, clk);
  run<2>("ex2.approx.f16x2", out, clk);
  return 0;
}
