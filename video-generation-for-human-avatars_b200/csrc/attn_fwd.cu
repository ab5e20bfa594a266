// attn_fwd.cu — tcgen05 / TMEM flash-attention forward, head_dim 64, non-causal (sm_100a).
//
// Replaces F.scaled_dot_product_attention on the LTXV block path (reference attention.py:1057-1064)
// for attn1 (self-attention over N latent tokens, q/k already qk-normed + RoPE'd) and for attn2
// (cross-attention to L caption tokens with the additive key bias the reference builds from
// encoder_attention_mask, attention.py:981-989 / transformer3d.py:440-445).
//
// q, k, v, o are token-major [B * N, ld] bf16 with head h in columns [64h, 64h+64): the layout the
// projection GEMMs write, so no head split/merge copies exist.  TMA reads 128x64 tiles straight
// out of that layout with a 3-D tensor map (col, token, batch).
//
// One CTA = one 128-query tile of one head; two CTAs are resident per SM so one CTA's softmax
// overlaps the other's MMAs.  Warp roles:
//   warp 0 lane 0 : TMA producer (Q once; K/V tiles double-buffered)
//   warp 1        : TMEM allocator; lane 0 issues tcgen05.mma  S = Q K^T (128x128x64) and
//                   O += P V (128x64x128; V is the MN-major B operand)
//   warps 2..5    : softmax, one query row per thread: tcgen05.ld S, running max with lazy
//                   rescale (only when the max grows by > 2^8), exp2, P -> shared memory (bf16,
//                   128B-swizzled K-major A operand), O rescale through tcgen05.ld/st.
#include "api_internal.h"
#include "common.cuh"
#include "tmap.h"

namespace b200 {

struct FaFwdParams {
  int B, H, Nq, Nk, kv_tiles;
  bf16* O;
  int64_t ldo;
  float* lse;             // [B, H, Nq], natural log
  const float* key_bias;  // [B, Nk] additive (natural units) or null
  float scale_log2;       // softmax scale * log2(e)
};

constexpr int FA_SMEM_TILES = 16384 /*Q*/ + 2 * 32768 /*K,V x2*/ + 32768 /*P*/;
constexpr int FA_FWD_SMEM = FA_SMEM_TILES + 128;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// ---- fast path (no key bias, full tile): the per-element code is FMNMX / FFMA+MUFU+FADD only.  The
// next 32-column TMEM chunk is in flight while the current one is processed.
__device__ __forceinline__ float tile_max_fast(uint32_t t_row, float sl2) {
  float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
  uint32_t ra[32], rb[32];
  tmem_ld32(t_row, ra);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t(&cur)[32] = (c & 1) ? rb : ra;
    uint32_t(&nxt)[32] = (c & 1) ? ra : rb;
    if (c < 3) tmem_ld32(t_row + (c + 1) * 32, nxt);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      a0 = fmaxf(a0, __uint_as_float(cur[i]));
      a1 = fmaxf(a1, __uint_as_float(cur[i + 1]));
      a2 = fmaxf(a2, __uint_as_float(cur[i + 2]));
      a3 = fmaxf(a3, __uint_as_float(cur[i + 3]));
    }
    if (c < 3) tmem_ld_wait();
  }
  return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * sl2;
}

__device__ __forceinline__ void store_p8(uint32_t addr, const float (&pv)[8]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(pv[0], pv[1])),
               "r"(pack_bf16x2(pv[2], pv[3])), "r"(pack_bf16x2(pv[4], pv[5])),
               "r"(pack_bf16x2(pv[6], pv[7]))
               : "memory");
}

__device__ __forceinline__ void tile_exp_fast(uint32_t t_row, float sl2, float neg_m, float& l0, float& l1,
                                              float& l2, float& l3) {
  uint32_t ra[32], rb[32];
  tmem_ld32(t_row, ra);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t(&cur)[32] = (c & 1) ? rb : ra;
    uint32_t(&nxt)[32] = (c & 1) ? ra : rb;
    if (c < 3) tmem_ld32(t_row + (c + 1) * 32, nxt);
    // P chunk c (32 keys -> 16 packed bf16x2 columns) overwrites S columns [16c, 16c+16), which belong
    // to S chunk c/2 <= c and have already been read.
    uint32_t pk[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(cur[g * 8 + i]), sl2, neg_m));
      l0 += pv[0] + pv[4];
      l1 += pv[1] + pv[5];
      l2 += pv[2] + pv[6];
      l3 += pv[3] + pv[7];
      pk[g * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
      pk[g * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
      pk[g * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
      pk[g * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
    }
    if (c < 3) tmem_ld_wait();  // chunk c+1 is in registers before any column it covers could be reused
    tmem_st16(t_row + c * 16, pk);
  }
}

// ---- general path (additive key bias and/or ragged last tile): kept out of line so that it does not
// weigh on the fast path's registers.
__device__ __noinline__ float tile_max_general(uint32_t t_row, float sl2, const float* kb, int key0, int Nk) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld32(t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int key = key0 + c * 32 + i;
      float sc = __uint_as_float(r[i]) * sl2;
      if (kb != nullptr && key < Nk) sc += __ldg(kb + key) * kLog2e;
      if (key >= Nk) sc = -INFINITY;
      mx = fmaxf(mx, sc);
    }
  }
  return mx;
}

__device__ __noinline__ float tile_exp_general(uint32_t t_row, float sl2, float neg_m, const float* kb,
                                               int key0, int Nk) {
  float l = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld32(t_row + c * 32, r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = key0 + c * 32 + g * 8 + i;
        float x = __uint_as_float(r[g * 8 + i]) * sl2;
        if (kb != nullptr && key < Nk) x += __ldg(kb + key) * kLog2e;
        if (key >= Nk) x = -INFINITY;
        pv[i] = ex2_approx(x + neg_m);
        l += pv[i];
      }
      pk[g * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
      pk[g * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
      pk[g * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
      pk[g * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
    }
    tmem_st16(t_row + c * 16, pk);
  }
  return l;
}

__global__ void __launch_bounds__(192, 2)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FaFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sQ = sbase, sKV = sbase + 16384, sP = sbase + 16384 + 65536;
  const uint32_t bar = sP + 32768;
  const uint32_t q_full = bar, kv_full0 = bar + 8, kv_empty0 = bar + 24, s_full = bar + 40,
                 p_full = bar + 48, pv_done = bar + 56, tmem_slot = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.kv_tiles;

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      printf("b200 fa_fwd: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(q_full, 16384);
    tma_load_3d(sQ, &tmQ, q_full, h * 64, qt * 128, b);
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      mbar_wait(kv_empty0 + 8 * s, ((j >> 1) & 1) ^ 1);
      mbar_expect_tx(kv_full0 + 8 * s, 32768);
      tma_load_3d(sKV + s * 32768, &tmK, kv_full0 + 8 * s, h * 64, j * 128, b);
      tma_load_3d(sKV + s * 32768 + 16384, &tmV, kv_full0 + 8 * s, h * 64, j * 128, b);
    }
  } else if (warp == 1) {
    // MMA warp: every lane follows the barrier waits, one elected lane issues; descriptors are constant
    // bases plus small offsets so that the instruction stream per MMA is minimal (the issuing thread
    // shares its scheduler with the softmax warps of both resident CTAs).
    const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t dq0 = make_smem_desc(sQ, 16, 1024), dk0 = make_smem_desc(sKV, 16, 1024),
                   dv0 = make_smem_desc(sKV + 16384, 8192, 1024);
    mbar_wait(q_full, 0);
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      mbar_wait(kv_full0 + 8 * s, (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = desc_adv(dk0, s * 32768);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tS, desc_adv(dq0, k * 32), desc_adv(dk, k * 32), idesc_qk, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc_adv(dv0, s * 32768);
        const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // A = P (bf16x2 packed, 8 TMEM columns per K=16 step)
          umma_ts(tO, tS + k * 8, desc_adv(dv, k * 2048), idesc_pv, k > 0 ? 1u : acc0);
        umma_commit(kv_empty0 + 8 * s);
        umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const float* kb = p.key_bias ? p.key_bias + (int64_t)b * p.Nk : nullptr;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int j = 0; j < T; ++j) {
      const int key0 = j * 128;
      const bool general = (kb != nullptr) || (key0 + 128 > p.Nk);  // bias or ragged last tile
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max
      const float mx = general ? tile_max_general(tS + lane_bits, sl2, kb, key0, p.Nk)
                               : tile_max_fast(tS + lane_bits, sl2);
      const float m_new = fmaxf(m_used, mx);
      const bool need = m_new > m_used + 8.f;
      const bool warp_need = __any_sync(0xffffffffu, need);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // PV_{j-1} finished: P buffer free, O consistent
        tc_fence_after();
      }
      if (warp_need) {
        const float alpha = ex2_approx(m_used - m_new);  // 0 on the first tile (m_used = -inf)
        m_used = m_new;
        l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tO + lane_bits + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(tO + lane_bits + c * 32, r);
          }
          tmem_st_wait();
        }
      }
      // pass 2: P = exp2(s * c - m), row sum, bf16 P into the swizzled A-operand tile
      if (general)
        l0 += tile_exp_general(tS + lane_bits, sl2, -m_used, kb, key0, p.Nk);
      else
        tile_exp_fast(tS + lane_bits, sl2, -m_used, l0, l1, l2, l3);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    const float l = (l0 + l1) + (l2 + l3);
    // epilogue: O / l -> bf16, lse
    mbar_wait(pv_done, (T - 1) & 1);
    tc_fence_after();
    const int q = qt * 128 + row;
    const float inv_l = 1.f / l;
    uint32_t r0[32], r1[32];
    tmem_ld32(tO + lane_bits, r0);
    tmem_ld32(tO + lane_bits + 32, r1);
    tmem_ld_wait();
    if (q < p.Nq) {
      bf16* orow = p.O + ((int64_t)b * p.Nq + q) * p.ldo + h * 64;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r0[g * 8 + 0]) * inv_l, __uint_as_float(r0[g * 8 + 1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(r0[g * 8 + 2]) * inv_l, __uint_as_float(r0[g * 8 + 3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(r0[g * 8 + 4]) * inv_l, __uint_as_float(r0[g * 8 + 5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(r0[g * 8 + 6]) * inv_l, __uint_as_float(r0[g * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + g * 8) = u;
        u.x = pack_bf16x2(__uint_as_float(r1[g * 8 + 0]) * inv_l, __uint_as_float(r1[g * 8 + 1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(r1[g * 8 + 2]) * inv_l, __uint_as_float(r1[g * 8 + 3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(r1[g * 8 + 4]) * inv_l, __uint_as_float(r1[g * 8 + 5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(r1[g * 8 + 6]) * inv_l, __uint_as_float(r1[g * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + 32 + g * 8) = u;
      }
      if (p.lse) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = (m_used + log2f(l)) * 0.6931471805599453f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// 3-D map over a token-major [B, N, ld] bf16 tensor restricted to `width` columns: box 64 x 128 x 1.
int make_tmap_tokens(CUtensorMap* out, const void* base, int B, int N, int64_t ld, int width,
                     int box_rows) {
  uint64_t dims[3] = {(uint64_t)width, (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)N * (uint64_t)ld * 2};
  uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tmap(out, base, TM_BF16, 3, dims, str, box, true);
}

}  // namespace b200

using namespace b200;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int b200_fa_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, void* o, int64_t ldo, float* lse, const float* key_bias,
                           int B, int H, int Nq, int Nk, int head_dim, float scale, void* stream) {
  if (!(q && k && v && o)) return arg_error("fa_fwd: null pointer");
  if (head_dim != 64) return arg_error("fa_fwd: only head_dim 64 is built (LTXV-2B: 32 heads x 64)");
  if (B < 0 || H <= 0 || Nq < 0 || Nk < 0) return arg_error("fa_fwd: bad shape");
  if (B == 0 || Nq == 0) return 0;
  if (Nk == 0) return arg_error("fa_fwd: no keys");
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || !al16(q) || !al16(k) || !al16(v) || !al16(o))
    return arg_error("fa_fwd: tensors must be 16-byte aligned with 16-byte-multiple pitches");
  if (ldq < H * 64 || ldk < H * 64 || ldv < H * 64 || ldo < H * 64) return arg_error("fa_fwd: pitch < H*64");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_tokens(&tmQ, q, B, Nq, ldq, H * 64, 128)) ||
      (rc = make_tmap_tokens(&tmK, k, B, Nk, ldk, H * 64, 128)) ||
      (rc = make_tmap_tokens(&tmV, v, B, Nk, ldv, H * 64, 128)))
    return arg_error("fa_fwd: cuTensorMapEncodeTiled failed", rc);
  FaFwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.kv_tiles = (Nk + 127) / 128;
  p.O = (bf16*)o; p.ldo = ldo; p.lse = lse; p.key_bias = key_bias;
  p.scale_log2 = scale * kLog2e;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_FWD_SMEM) != cudaSuccess)
      return launch_status("fa_fwd: cudaFuncSetAttribute");
    cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  dim3 grid((Nq + 127) / 128, H, B);
  fa_fwd_kernel<<<grid, 192, FA_FWD_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  return launch_status("fa_fwd");
}
