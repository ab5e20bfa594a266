// tmem_probe.cu — tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps, alone and
// with a concurrent tcgen05.mma stream (does reading accumulators out of TMEM slow the tensor pipe?).
#include <cstdio>
#include "common.cuh"
using namespace b200;

template <int MODE>  // 0: ld32 ; 1: ld16 x2 ; 2: st32 ; 3: ld32 with 2 loads in flight
__global__ void __launch_bounds__(640, 1) probe(int nwarps, int reps, int with_mma, int mma_n, long long* out) {
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
  __shared__ long long t_end[20];
  long long t0 = clock64();
  if (warp == 1) {
    if (with_mma) {
      const uint32_t idesc = make_idesc_bf16(128, mma_n, 0, 0);
      const uint64_t da0 = make_smem_desc(sA, 16, 1024), db0 = make_smem_desc(sB, 16, 1024);
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (with_mma == 2) umma_ts(tmem_base + 256, tmem_base + 384 + k * 8, db0 + 2 * k, idesc, 1u);
            else umma_ss(tmem_base + 256, da0 + 2 * k, db0 + 2 * k, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, 0);
      if (lane == 0) t_end[17] = clock64() - t0;
    }
  } else if (warp >= 2 && warp < 2 + nwarps) {
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = ((warp - 2) >> 2) * 32;  // different columns per warp of a quadrant
    uint32_t acc = 0;
    for (int r = 0; r < reps; ++r) {
      if (MODE == 0) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_bits + col, v);
        tmem_ld_wait();
        acc += v[0] ^ v[31];
      } else if (MODE == 1) {
        uint32_t a[16], b[16];
        tmem_ld16(tmem_base + lane_bits + col, a);
        tmem_ld16(tmem_base + lane_bits + col + 16, b);
        tmem_ld_wait();
        acc += a[0] ^ b[15];
      } else if (MODE == 2) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc + i;
        tmem_st32(tmem_base + lane_bits + col, v);
        tmem_st_wait();
        acc++;
      } else if (MODE == 4) {
        uint32_t a[16], b[16], c8[8];
        tmem_ld16(tmem_base + lane_bits + (col & 127), a);
        tmem_ld16(tmem_base + lane_bits + 128 + (col & 127), b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) c8[i] = a[i] ^ b[i] ^ acc;
        tmem_st8(tmem_base + lane_bits + (col & 127), c8);
        tmem_st8(tmem_base + lane_bits + 128 + (col & 127), c8);
        tmem_st_wait();
        acc += c8[0];
      } else {
        uint32_t a[32], b[32];
        tmem_ld32(tmem_base + lane_bits + col, a);
        tmem_ld32(tmem_base + lane_bits + (col ^ 32), b);
        tmem_ld_wait();
        acc += a[0] ^ b[31];
      }
    }
    if (lane == 0) t_end[warp - 2] = clock64() - t0 + (acc == 0x12345 ? 1 : 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long mx = 0;
    for (int w = 0; w < nwarps; ++w) mx = t_end[w] > mx ? t_end[w] : mx;
    if (nwarps == 0) mx = 1;
    out[0] = mx;
    out[1] = with_mma ? t_end[17] : 0;
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int MODE>
void run(const char* name, int nwarps, int with_mma, int mma_n, long long* d) {
  const int smem = 32768 + 65536 + 1024, reps = 512;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<MODE><<<148, 640, smem>>>(nwarps, reps, with_mma, mma_n, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double bytes = (double)nwarps * reps * 4096 * (MODE == 3 ? 2 : 1);
  printf("%-22s warps %2d  mma %3d : %7.1f B/clk/SM (%6.1f clk per 4 KB warp access)", name, nwarps, with_mma ? mma_n : 0,
         bytes / h[0], (double)h[0] / reps);
  if (with_mma) printf("   mma %6.1f clk each (ideal %d)", (double)h[1] / (reps * 4), mma_n / 2);
  printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int nw : {4, 8, 16}) run<0>("ld 32x32b.x32", nw, 0, 0, d);
  for (int nw : {4, 8, 16}) run<1>("ld 32x32b.x16 x2", nw, 0, 0, d);
  for (int nw : {4, 8, 16}) run<3>("ld x32, 2 in flight", nw, 0, 0, d);
  for (int nw : {4, 8, 16}) run<2>("st 32x32b.x32", nw, 0, 0, d);
  for (int nw : {4, 16}) for (int n : {64, 128, 256}) run<0>("ld x32 + SS mma", nw, 1, n, d);
  for (int nw : {4, 16}) run<2>("st x32 + SS mma", nw, 1, 128, d);
  run<0>("ld x32 + TS mma N=64", 16, 2, 64, d);
  run<4>("ld16x2+st8x2 + TS mma", 16, 2, 64, d);
  run<4>("ld16x2+st8x2 + SS mma", 16, 1, 64, d);
  run<4>("ld16x2+st8x2 alone", 16, 0, 0, d);
  run<0>("(none) TS mma N=64", 0, 2, 64, d);
  return 0;
}
