// guidance.cu — the element-wise tail of one sampling step, fused:
//   classifier-free guidance (optionally CFG* projection rescale), spatio-temporal guidance, std rescale
//       reference: pipelines/pipeline_ltx_video.py:1217-1260
//   Euler update x <- x - dt * v on the fp32 running latents   (schedulers/rf.py:305-374; the reference's latents are
//       fp32 from the first step on because dt is an fp32 tensor with dimensions)
//   conditioning-mask select: only tokens with t - 1e-6 < 1 - conditioning_mask move   (pipeline :1346-1379)
//   and the bf16 model input of the next step, replicated once per condition   (pipeline :1136-1138, :1203)
// The model output holds the conditions back to back, [conds * B, N, C] in the order (uncond,) text (, perturbed).
// Per-sample reductions (the CFG* dot products, the two standard deviations) are accumulated in double through
// atomics into a caller-owned workspace, so the whole tail is at most three small launches after one memset; the
// per-step scalars are read from device memory so that a captured step can be replayed with new values.
#include "api_internal.h"
#include "common.cuh"

namespace b200 {

struct GuidanceParams {
  const bf16* v;             // [conds * B, per]
  float* x;                  // [B, per] running latents
  bf16* x_next;              // [n_next * B, per] or null
  const float* dt;           // [1] or [N]: step per token (shared by the batch, as the reference's timestep[:1])
  const float* noise_level;  // [B, N] = 1 - conditioning_mask, or null
  const float* scalars;      // device: guidance_scale, stg_scale, rescaling_scale, t
  double* ws;                // [B][6]: <text,uncond>, |uncond|^2, sum text, sum text^2, sum pred, sum pred^2
  int64_t per;               // N * C
  int B, C, n_next;
  int has_cfg, has_stg, cfg_star, rescale, dt_per_token;
};

__device__ __forceinline__ void ld8(const bf16* p, float* f) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

template <int NACC>
__device__ __forceinline__ void block_accumulate(double (&acc)[NACC], double* dst) {
  __shared__ double red[NACC][8];
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    double s = acc[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[a][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    atomicAdd(dst + threadIdx.x, s);
  }
}

// guided prediction of 8 consecutive elements, before the std rescale
__device__ __forceinline__ void guided8(const GuidanceParams& p, int b, int64_t e, float alpha, float gs, float stg,
                                        float* pred, float* text) {
  const int i_text = p.has_cfg ? 1 : 0;
  ld8(p.v + ((int64_t)(i_text * p.B + b)) * p.per + e, text);
#pragma unroll
  for (int j = 0; j < 8; ++j) pred[j] = text[j];
  if (p.has_cfg) {
    float un[8];
    ld8(p.v + (int64_t)b * p.per + e, un);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = alpha * un[j];
      pred[j] = u + gs * (text[j] - u);
    }
  }
  if (p.has_stg) {
    float pt[8];
    ld8(p.v + ((int64_t)((i_text + 1) * p.B + b)) * p.per + e, pt);
#pragma unroll
    for (int j = 0; j < 8; ++j) pred[j] += stg * (text[j] - pt[j]);
  }
}

__device__ __forceinline__ float cfg_star_alpha(const GuidanceParams& p, int b) {
  if (!(p.has_cfg && p.cfg_star)) return 1.f;
  return (float)(p.ws[b * 6 + 0] / (p.ws[b * 6 + 1] + 1e-8));
}

// phase 1: <text, uncond> and |uncond|^2 per sample
__global__ void __launch_bounds__(256) guidance_dot_kernel(GuidanceParams p) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.y;
  double acc[2] = {0.0, 0.0};
  const int64_t n8 = p.per / 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
    float un[8], tx[8];
    ld8(p.v + (int64_t)b * p.per + i * 8, un);
    ld8(p.v + (int64_t)(p.B + b) * p.per + i * 8, tx);
    float d = 0.f, s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d += tx[j] * un[j];
      s += un[j] * un[j];
    }
    acc[0] += d;
    acc[1] += s;
  }
  block_accumulate<2>(acc, p.ws + b * 6);
}

// phase 2: sums and sums of squares of the text prediction and of the guided prediction
__global__ void __launch_bounds__(256) guidance_std_kernel(GuidanceParams p) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.y;
  const float gs = p.scalars[0], stg = p.scalars[1];
  const float alpha = cfg_star_alpha(p, b);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t n8 = p.per / 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
    float pred[8], text[8];
    guided8(p, b, i * 8, alpha, gs, stg, pred, text);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0 += text[j];
      a1 += text[j] * text[j];
      a2 += pred[j];
      a3 += pred[j] * pred[j];
    }
    acc[0] += a0; acc[1] += a1; acc[2] += a2; acc[3] += a3;
  }
  block_accumulate<4>(acc, p.ws + b * 6 + 2);
}

// phase 3: guided prediction (x std-rescale factor), Euler step, conditioning select, next model input
__global__ void __launch_bounds__(256) guidance_apply_kernel(GuidanceParams p) {
  pdl_launch();
  pdl_wait();
  const int b = blockIdx.y;
  const float gs = p.scalars[0], stg = p.scalars[1], rs = p.scalars[2], t = p.scalars[3];
  const float alpha = cfg_star_alpha(p, b);
  float factor = 1.f;
  if (p.rescale) {
    // unbiased standard deviations (torch .std default), pipeline :1247-1257
    const double n = (double)p.per;
    const double* w = p.ws + b * 6 + 2;
    const double var_t = (w[1] - w[0] * w[0] / n) / (n - 1.0);
    const double var_p = (w[3] - w[2] * w[2] / n) / (n - 1.0);
    const float f = (float)(sqrt(var_t > 0.0 ? var_t : 0.0) / sqrt(var_p > 0.0 ? var_p : 0.0));
    factor = rs * f + (1.f - rs);
  }
  const int64_t n8 = p.per / 8;
  const int c8 = p.C / 8;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
    float pred[8], text[8];
    guided8(p, b, i * 8, alpha, gs, stg, pred, text);
    const int64_t tok = i / c8;
    const float dt = p.dt[p.dt_per_token ? tok : 0];
    const bool move = p.noise_level == nullptr || (t - 1e-6f < p.noise_level[(int64_t)b * (p.per / p.C) + tok]);
    float* xp = p.x + (int64_t)b * p.per + i * 8;
    float4 x0 = *reinterpret_cast<const float4*>(xp), x1 = *reinterpret_cast<const float4*>(xp + 4);
    float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    if (move) {
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = xv[j] - dt * (pred[j] * factor);
      *reinterpret_cast<float4*>(xp) = make_float4(xv[0], xv[1], xv[2], xv[3]);
      *reinterpret_cast<float4*>(xp + 4) = make_float4(xv[4], xv[5], xv[6], xv[7]);
    }
    if (p.x_next) {
      uint4 u;
      u.x = pack_bf16x2(xv[0], xv[1]);
      u.y = pack_bf16x2(xv[2], xv[3]);
      u.z = pack_bf16x2(xv[4], xv[5]);
      u.w = pack_bf16x2(xv[6], xv[7]);
      for (int r = 0; r < p.n_next; ++r)
        *reinterpret_cast<uint4*>(p.x_next + ((int64_t)(r * p.B + b)) * p.per + i * 8) = u;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_guidance_step_workspace_bytes(int B) { return (int64_t)(B > 0 ? B : 0) * 6 * sizeof(double); }

extern "C" int b200_guidance_step(const void* v, float* x, void* x_next, int n_next, const float* dt,
                                  int dt_per_token, const float* noise_level, const float* scalars, int B,
                                  int64_t N, int C, int has_cfg, int has_stg, int cfg_star, int rescale,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  if (B < 0 || N < 0 || C <= 0 || n_next < 0) return arg_error("guidance_step: bad shape");
  if (B == 0 || N == 0) return 0;
  if (!(v && x && dt && scalars)) return arg_error("guidance_step: null pointer");
  if (n_next > 0 && !x_next) return arg_error("guidance_step: x_next is null but n_next > 0");
  if (C % 8 || (reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(x_next) & 15))
    return arg_error("guidance_step: C must be a multiple of 8 and the tensors 16-byte aligned");
  const bool need_ws = (has_cfg && cfg_star) || rescale;
  if (need_ws && (!workspace || workspace_bytes < b200_guidance_step_workspace_bytes(B) ||
                  (reinterpret_cast<uintptr_t>(workspace) & 7)))
    return arg_error("guidance_step: workspace too small (see b200_guidance_step_workspace_bytes)");
  GuidanceParams p;
  p.v = (const bf16*)v;
  p.x = x;
  p.x_next = n_next > 0 ? (bf16*)x_next : nullptr;
  p.dt = dt;
  p.noise_level = noise_level;
  p.scalars = scalars;
  p.ws = (double*)workspace;
  p.per = N * C;
  p.B = B;
  p.C = C;
  p.n_next = n_next;
  p.has_cfg = has_cfg != 0;
  p.has_stg = has_stg != 0;
  p.cfg_star = cfg_star != 0;
  p.rescale = rescale != 0;
  p.dt_per_token = dt_per_token != 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n8 = p.per / 8;
  int64_t gx = (n8 + 255) / 256;
  const int64_t cap = (148 * 8 + B - 1) / B;  // ~8 CTAs per SM over the whole batch
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)B);
  if (need_ws) {
    cudaError_t e = cudaMemsetAsync(workspace, 0, b200_guidance_step_workspace_bytes(B), s);
    if (e != cudaSuccess) return arg_error("guidance_step: cudaMemsetAsync failed", (int)e);
  }
  if (p.has_cfg && p.cfg_star) B200_LAUNCH(guidance_dot_kernel, grid, 256, 0, s, p);
  if (p.rescale) B200_LAUNCH(guidance_std_kernel, grid, 256, 0, s, p);
  B200_LAUNCH(guidance_apply_kernel, grid, 256, 0, s, p);
  return launch_status("guidance_step");
}
