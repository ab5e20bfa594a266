// mma_probe.cu — micro-benchmark: cycles per tcgen05.mma as a function of HOW the instruction stream
// that issues it is written (descriptor arithmetic per MMA, divergent single-lane branch vs elect in a
// converged warp) and of the operand flavour (SS / TS, N = 64/128/256).  Operand values are irrelevant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../video-generation-for-human-avatars_b200/csrc \
//        mma_probe.cu -o mma_probe -lcuda && ./mma_probe
#include <cstdio>
#include "common.cuh"
using namespace b200;

// VARIANT 0: descriptors rebuilt from the address for every MMA, issuing lane selected by `lane == 0`
// VARIANT 1: base descriptors built once, + (bytes >> 4) per MMA, `lane == 0` branch
// VARIANT 2: like 1, whole warp converged, elect_one() around each group of 4 MMAs
template <int VARIANT, int TS, int N, int A_MN = 0, int B_MN = 0>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sA = sbase, sB = sbase + 32768, bar = sbase + 32768 + 65536, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  const uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
  if (VARIANT < 2) {
    if (warp == 1 && lane == 0) {
      const uint64_t da0 = make_smem_desc(sA, 16, 1024), db0 = make_smem_desc(sB, 16, 1024);
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint64_t da, db;
          if (VARIANT == 0) { da = make_smem_desc(sA + k * 32, 16, 1024); db = make_smem_desc(sB + k * 32, 16, 1024); }
          else { da = da0 + 2 * k; db = db0 + 2 * k; }
          if (TS) umma_ts(tmem_base, tmem_base + 256 + k * 8, db, idesc, 1u);
          else umma_ss(tmem_base, da, db, idesc, 1u);
        }
      }
      long long t1 = clock64();
      umma_commit(bar);
      mbar_wait(bar, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else {
    if (warp == 1) {
      const uint64_t da0 = A_MN ? make_smem_desc(sA, 16384, 1024) : make_smem_desc(sA, 16, 1024);
      const uint64_t db0 = B_MN ? make_smem_desc(sB, 8192, 1024) : make_smem_desc(sB, 16, 1024);
      constexpr int SA = A_MN ? 128 : 2, SB = B_MN ? 128 : 2;  // descriptor step per K=16 ((bytes) >> 4)
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (TS) umma_ts(tmem_base + (k & 1) * 64, tmem_base + 256 + k * 16, db0 + SB * k, idesc, 1u);
            else umma_ss(tmem_base, da0 + SA * k, db0 + SB * k, idesc, 1u);
          }
        }
        __syncwarp();
      }
      long long t1 = clock64();
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int V, int TS, int N, int A_MN = 0, int B_MN = 0>
void run(const char* name, long long* d) {
  const int smem = 32768 + 65536 + 1024, reps = 256;
  cudaFuncSetAttribute(probe<V, TS, N, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<V, TS, N, A_MN, B_MN><<<148, 128, smem>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("variant %d  %-10s issue %6.1f clk/mma   complete %6.1f clk/mma   (ideal %d)  %s\n", V, name,
         (double)h[0] / (reps * 4), (double)h[1] / (reps * 4), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<2, 1, 64, 0, 1>("TS64 B=MN", d); run<2, 0, 64, 0, 1>("SS64 K/MN", d); run<2, 0, 64, 1, 1>("SS64 MN/MN", d);
  run<2, 0, 128, 1, 1>("SS128MN/MN", d); run<2, 1, 128, 0, 1>("TS128 B=MN", d);
  run<0, 0, 64>("SS N=64", d);  run<1, 0, 64>("SS N=64", d);  run<2, 0, 64>("SS N=64", d);
  run<0, 0, 128>("SS N=128", d); run<1, 0, 128>("SS N=128", d); run<2, 0, 128>("SS N=128", d);
  run<0, 0, 256>("SS N=256", d); run<1, 0, 256>("SS N=256", d); run<2, 0, 256>("SS N=256", d);
  run<0, 1, 64>("TS N=64", d);  run<1, 1, 64>("TS N=64", d);  run<2, 1, 64>("TS N=64", d);
  run<0, 1, 128>("TS N=128", d); run<1, 1, 128>("TS N=128", d); run<2, 1, 128>("TS N=128", d);
  return 0;
}
