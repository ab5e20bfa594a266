// adamw.cu — one launch of AdamW over every trainable tensor of the LoRA fine-tune.
// The reference builds `torch.optim.AdamW(params, lr)` (training.py:271) and steps it once per micro-step
// (training.py:206): decoupled weight decay 1e-2, betas (0.9, 0.999), eps 1e-8, optimizer state in the parameter's
// dtype (fp32 LoRA adapters, bf16 caption projection).  Here all 116 tensors (27.3 M elements) go through ONE kernel:
// a device table of (param, grad, exp_avg, exp_avg_sq, numel, dtype) entries and a block -> (entry, chunk) map built
// once by the caller; the step counter and the hyper-parameters live in device memory, so the launch can be captured
// in a CUDA graph and replayed.  HBM bound: 28 B per fp32 element, 14 B per bf16 element.
#include "api_internal.h"
#include "common.cuh"

namespace b200 {

struct AdamWEntry {
  void* p;
  const void* g;
  void* m;
  void* v;
  int64_t numel;
  int32_t is_bf16;
  int32_t pad;
};
static_assert(sizeof(AdamWEntry) == 48, "table layout is part of the C ABI (see include/b200ltx.h)");

constexpr int ADAMW_CHUNK = 8192;  // elements per block

struct AdamWCoef {
  float lr, beta1, beta2, eps, wd, step_size, inv_bc2_sqrt;
};

__device__ __forceinline__ void adamw_math(float& p, float g, float& m, float& v, const AdamWCoef& c) {
  p -= c.lr * c.wd * p;                      // decoupled weight decay
  m = m + (1.f - c.beta1) * (g - m);         // exp_avg
  v = c.beta2 * v + (1.f - c.beta2) * g * g; // exp_avg_sq
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
  p -= c.step_size * (m / denom);
}

// hp (device): {lr, beta1, beta2, eps, weight_decay}; step (device): completed steps, incremented by the last block
__global__ void __launch_bounds__(256) adamw_kernel(const AdamWEntry* __restrict__ table,
                                                    const int32_t* __restrict__ block_map,
                                                    float* __restrict__ step, const float* __restrict__ hp,
                                                    int32_t* __restrict__ done) {
  pdl_launch();
  pdl_wait();
  __shared__ AdamWCoef sc;
  if (threadIdx.x == 0) {
    const double t = (double)step[0] + 1.0;
    sc.lr = hp[0]; sc.beta1 = hp[1]; sc.beta2 = hp[2]; sc.eps = hp[3]; sc.wd = hp[4];
    const double bc1 = 1.0 - pow((double)hp[1], t), bc2 = 1.0 - pow((double)hp[2], t);
    sc.step_size = (float)((double)hp[0] / bc1);
    sc.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const AdamWCoef c = sc;
  const AdamWEntry e = table[block_map[2 * blockIdx.x]];
  const int64_t begin = (int64_t)block_map[2 * blockIdx.x + 1] * ADAMW_CHUNK;
  const int64_t end = begin + ADAMW_CHUNK < e.numel ? begin + ADAMW_CHUNK : e.numel;
  if (e.is_bf16) {
    bf16* p = (bf16*)e.p;
    const bf16* g = (const bf16*)e.g;
    bf16 *m = (bf16*)e.m, *v = (bf16*)e.v;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec)
      for (int64_t i = begin + (int64_t)threadIdx.x * 8; i + 8 <= end; i += 256 * 8) {
        uint4 up = *reinterpret_cast<const uint4*>(p + i), ug = *reinterpret_cast<const uint4*>(g + i);
        uint4 um = *reinterpret_cast<const uint4*>(m + i), uv = *reinterpret_cast<const uint4*>(v + i);
        uint32_t* wp = &up.x; const uint32_t* wg = &ug.x; uint32_t* wm = &um.x; uint32_t* wv = &uv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float p0 = bf16_lo(wp[j]), p1 = bf16_hi(wp[j]), m0 = bf16_lo(wm[j]), m1 = bf16_hi(wm[j]);
          float v0 = bf16_lo(wv[j]), v1 = bf16_hi(wv[j]);
          adamw_math(p0, bf16_lo(wg[j]), m0, v0, c);
          adamw_math(p1, bf16_hi(wg[j]), m1, v1, c);
          wp[j] = pack_bf16x2(p0, p1); wm[j] = pack_bf16x2(m0, m1); wv[j] = pack_bf16x2(v0, v1);
        }
        *reinterpret_cast<uint4*>(p + i) = up;
        *reinterpret_cast<uint4*>(m + i) = um;
        *reinterpret_cast<uint4*>(v + i) = uv;
      }
    // tail (or the whole chunk when a tensor is not 16-byte aligned)
    const int64_t tail0 = vec ? begin + ((end - begin) / 8) * 8 : begin;
    for (int64_t k = tail0 + threadIdx.x; k < end; k += 256) {
      float pp = __bfloat162float(p[k]), mm = __bfloat162float(m[k]), vv = __bfloat162float(v[k]);
      adamw_math(pp, __bfloat162float(g[k]), mm, vv, c);
      p[k] = __float2bfloat16(pp); m[k] = __float2bfloat16(mm); v[k] = __float2bfloat16(vv);
    }
  } else {
    float* p = (float*)e.p;
    const float* g = (const float*)e.g;
    float *m = (float*)e.m, *v = (float*)e.v;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec)
      for (int64_t i = begin + (int64_t)threadIdx.x * 4; i + 4 <= end; i += 256 * 4) {
        float4 fp = *reinterpret_cast<const float4*>(p + i), fg = *reinterpret_cast<const float4*>(g + i);
        float4 fm = *reinterpret_cast<const float4*>(m + i), fv = *reinterpret_cast<const float4*>(v + i);
        adamw_math(fp.x, fg.x, fm.x, fv.x, c);
        adamw_math(fp.y, fg.y, fm.y, fv.y, c);
        adamw_math(fp.z, fg.z, fm.z, fv.z, c);
        adamw_math(fp.w, fg.w, fm.w, fv.w, c);
        *reinterpret_cast<float4*>(p + i) = fp;
        *reinterpret_cast<float4*>(m + i) = fm;
        *reinterpret_cast<float4*>(v + i) = fv;
      }
    const int64_t tail0 = vec ? begin + ((end - begin) / 4) * 4 : begin;
    for (int64_t k = tail0 + threadIdx.x; k < end; k += 256) {
      float pp = p[k], mm = m[k], vv = v[k];
      adamw_math(pp, g[k], mm, vv, c);
      p[k] = pp; m[k] = mm; v[k] = vv;
    }
  }
  // the last block to finish advances the step counter (every block has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      step[0] = step[0] + 1.f;
      *done = 0;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_adamw_chunk_elems(void) { return ADAMW_CHUNK; }

extern "C" int b200_adamw_step(const void* table, const int32_t* block_map, int n_blocks, float* step,
                               const float* hyper, int32_t* done_counter, void* stream) {
  if (n_blocks < 0) return arg_error("adamw_step: bad block count");
  if (n_blocks == 0) return 0;
  if (!(table && block_map && step && hyper && done_counter)) return arg_error("adamw_step: null pointer");
  if (reinterpret_cast<uintptr_t>(table) & 7) return arg_error("adamw_step: table must be 8-byte aligned");
  B200_LAUNCH(adamw_kernel, n_blocks, 256, 0, (cudaStream_t)stream, (const AdamWEntry*)table, block_map, step, hyper,
                                                           done_counter);
  return launch_status("adamw_step");
}
