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
// One CTA = one 128-query tile of one head, walking the keys 64 at a time; FOUR CTAs are resident per SM.
// At head_dim 64 the exp unit (16 ex2/clk/SM measured = 1024 clk per 128 x 128 scores against 512 clk of MMA) is the
// nominal bound, but what the measurements showed is a latency-bound kernel: with two softmax warps per scheduler
// (2 CTAs x 128-key tiles, 168 registers) the exp unit was ~50 % busy and nothing else saturated -- offloading
// exponentials to the FMA pipe, a speculative single pass, 8 softmax warps per CTA (pair barriers, spills) all
// measured equal or worse (profiles/r02_probes_mma_tmem.md).  Independent CTAs are the cheap source of independent
// warps: a 128 x 64 step needs 128 TMEM columns (S 64 | O 64, P in place), 48 KB of shared memory and, with
// chunk-wise TMEM reads, 80 registers, so four CTAs fit an SM and every scheduler has 4 softmax warps.
//   warp 0 lane 0 : TMA producer (Q once; K/V steps through a 2-stage ring)
//   warp 1        : TMEM allocator + MMA issue (one elected lane): S = Q K^T (128x64x64), O += P V (128x64x64;
//                   V is the MN-major B operand, P the TMEM A operand).  S(j+1) overwrites P(j): it is issued
//                   behind P V(j), and the tensor pipe runs in order.
//   warps 2..5    : softmax, one query row per thread, two passes over the 64 scores in 32-column TMEM chunks:
//                   row max (3-input max), lazy rescale (only when the running max grows by > 2^8), exp2, packed
//                   bf16 P written over the first half of the S columns.  Per-key bias (attn2, ragged steps) is
//                   staged in shared memory once per step.
#include <stdlib.h>

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
  // fa_fwd_db_kernel, last partial wave: work items (q tile, head, batch) with linear id >= n_whole are each handled by
  // `parts` CTAs that walk disjoint ranges of the key steps and leave (O unnormalised fp32, m, l) in `part_ws`
  // [split item][part][128 rows][FA2_PART_LD floats]; fa_fwd_merge_kernel folds the parts.  n_whole = all items: no split.
  int q_tiles, n_whole, parts;
  float* part_ws;
  // STG "attention values" skip (attention.py:1078-1084): batch entries whose flag is 0 get their value rows as the
  // attention output (needs Nq == Nk); their CTAs copy one V tile and leave without touching barriers or TMEM
  const float* batch_keep;   // [B] 1 = attend, 0 = pass `pass_src` through; null = attend everywhere
  const bf16* pass_src;      // [B*Nq, ld_pass]: the value rows ("attention values") or the attention input ("attention skip")
  int64_t ld_pass;
};

// o[rows of this query tile, head h] = pass_src[same rows, head h]   (whole CTA, before any barrier / TMEM set-up)
__device__ __forceinline__ void fa_copy_values(const FaFwdParams& p, int qt, int h, int b, int nthreads) {
  for (int idx = threadIdx.x; idx < 128 * 8; idx += nthreads) {
    const int r = idx >> 3, c8 = (idx & 7) * 8;
    const int q = qt * 128 + r;
    if (q >= p.Nq) continue;
    const uint4 val = *reinterpret_cast<const uint4*>(p.pass_src + ((int64_t)b * p.Nq + q) * p.ld_pass + h * 64 + c8);
    *reinterpret_cast<uint4*>(p.O + ((int64_t)b * p.Nq + q) * p.ldo + h * 64 + c8) = val;
    if (p.lse && c8 == 0) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = 0.f;
  }
}
constexpr int FA2_PART_LD = 66;   // 64 O columns + running max (log2 units) + row sum

// Geometry: 128 queries x 64 keys per step, FOUR CTAs per SM (TMEM 128 columns each: S 64 | O 64, P written in
// place over the first half of S).  Measured: with one 128 x 128 tile per step and two CTAs per SM the exp unit was
// ~50 % busy and nothing else was saturated -- two softmax warps per scheduler cannot cover each other's TMEM round
// trips, dependent-issue stalls and barrier waits.  Halving the key tile halves every per-CTA resource (TMEM, smem,
// registers through chunk-wise TMEM reads), so four independent CTAs = 4 softmax warps per scheduler fit, with no
// cross-warp synchronisation added.
constexpr int FA_BN = 64;                 // keys per step
constexpr int FA_KV_STAGES = 2;
constexpr int FA_KV_STAGE_BYTES = 2 * FA_BN * 128;  // K and V tile of one step
constexpr int FA_SMEM_TILES = 16384 /*Q*/ + FA_KV_STAGES * FA_KV_STAGE_BYTES;
constexpr int FA_FWD_SMEM = FA_SMEM_TILES + 128 + 256;  // + barriers + 64 staged key-bias floats
constexpr int FA_FWD_THREADS = 192;       // TMA warp, MMA warp, 4 softmax warps
constexpr int FA_CTAS_PER_SM = 4;
constexpr int FA_MASK_SCAN_MAX = 1024;   // key counts up to which the bias vector is scanned for trailing masked keys
static_assert(FA_CTAS_PER_SM * (FA_FWD_SMEM + 1024) <= 233472, "fa_fwd CTAs must fit one SM");
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// max over one 32-column chunk held in registers, folded into four running accumulators
__device__ __forceinline__ void chunk_max(const uint32_t (&r)[32], float& a0, float& a1, float& a2, float& a3) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    a0 = max3(a0, __uint_as_float(r[i + 0]), __uint_as_float(r[i + 1]));
    a1 = max3(a1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    a2 = max3(a2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
    a3 = max3(a3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
  }
}

// P = exp2(s * mul + neg_m) for one 32-column chunk -> 16 packed bf16x2 TMEM columns; row sum into l0..l3
__device__ __forceinline__ void chunk_exp(const uint32_t (&r)[32], float mul, float neg_m, uint32_t t_dst,
                                          float& l0, float& l1, float& l2, float& l3) {
  uint32_t pk[16];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(r[g * 8 + i]), mul, neg_m));
    l0 += pv[0] + pv[4];
    l1 += pv[1] + pv[5];
    l2 += pv[2] + pv[6];
    l3 += pv[3] + pv[7];
    pk[g * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
    pk[g * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
    pk[g * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
    pk[g * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
  }
  tmem_st16(t_dst, pk);
}

// bias path: x = s * sl2 + (bias * log2e | -inf past Nk), in place (the staged per-key terms are read
// back from shared memory with broadcast LDS.128)
__device__ __forceinline__ void chunk_add_bias(uint32_t (&r)[32], float sl2, const float* kbs) {
  const float4* kb4 = reinterpret_cast<const float4*>(kbs);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 bb = kb4[i];
    r[4 * i + 0] = __float_as_uint(fmaf(__uint_as_float(r[4 * i + 0]), sl2, bb.x));
    r[4 * i + 1] = __float_as_uint(fmaf(__uint_as_float(r[4 * i + 1]), sl2, bb.y));
    r[4 * i + 2] = __float_as_uint(fmaf(__uint_as_float(r[4 * i + 2]), sl2, bb.z));
    r[4 * i + 3] = __float_as_uint(fmaf(__uint_as_float(r[4 * i + 3]), sl2, bb.w));
  }
}

__global__ void __launch_bounds__(FA_FWD_THREADS, FA_CTAS_PER_SM)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FaFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sQ = sbase, sKV = sbase + 16384;
  const uint32_t bar = sKV + FA_KV_STAGES * FA_KV_STAGE_BYTES;
  const uint32_t q_full = bar, kv_full0 = bar + 8, kv_empty0 = bar + 24, s_full = bar + 40, p_full = bar + 48,
                 pv_done = bar + 56, tmem_slot = bar + 64;
  float* kb_stage = reinterpret_cast<float*>(smem_raw + (bar + 128 - sbase));  // [64] per-key term of the current step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  pdl_launch();
  pdl_wait();   // this kernel reads global memory (key mask, batch_keep) from its first lines on
  if (p.batch_keep != nullptr && p.batch_keep[b] == 0.f) {
    fa_copy_values(p, qt, h, b, FA_FWD_THREADS);
    return;
  }
  // Keys whose additive bias is <= -9000 (the reference masks with -10000, transformer3d.py:440-445) have weight
  // exp(-9000) = 0 exactly in fp32: every key step past the last unmasked key is skipped -- bit-identical, and with
  // the real prompt (~15 valid of 256 caption tokens) three of the four steps of attn2 disappear.
  __shared__ int s_last_key;
  if (threadIdx.x == 0) s_last_key = -1;
  __syncthreads();
  if (p.key_bias != nullptr && p.Nk <= FA_MASK_SCAN_MAX) {
    const float* kbp = p.key_bias + (int64_t)b * p.Nk;
    int last = -1;
    for (int key = threadIdx.x; key < p.Nk; key += FA_FWD_THREADS)
      if (kbp[key] > -9000.f) last = key;
    if (last >= 0) atomicMax(&s_last_key, last);
  }

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      printf("b200 fa_fwd: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tO = tmem_base + 64;
  // number of key steps: all of them, or up to the last unmasked key (none unmasked: keep all, as the reference)
  const int T = s_last_key >= 0 ? min(p.kv_tiles, s_last_key / FA_BN + 1) : p.kv_tiles;

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(q_full, 16384);
    tma_load_3d(sQ, &tmQ, q_full, h * 64, qt * 128, b);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(kv_empty0 + 8 * s, ph ^ 1);
      mbar_expect_tx(kv_full0 + 8 * s, FA_KV_STAGE_BYTES);
      tma_load_3d(sKV + s * FA_KV_STAGE_BYTES, &tmK, kv_full0 + 8 * s, h * 64, j * FA_BN, b);
      tma_load_3d(sKV + s * FA_KV_STAGE_BYTES + FA_BN * 128, &tmV, kv_full0 + 8 * s, h * 64, j * FA_BN, b);
      if (++s == FA_KV_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // MMA warp: every lane follows the barrier waits, one elected lane issues; descriptors are constant bases plus
    // small offsets so that the instruction stream per MMA is minimal.
    const uint32_t idesc_qk = make_idesc_bf16(128, FA_BN, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
    const uint64_t dq0 = make_smem_desc(sQ, 16, 1024), dk0 = make_smem_desc(sKV, 16, 1024),
                   dv0 = make_smem_desc(sKV + FA_BN * 128, 8192, 1024);
    mbar_wait(q_full, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < T; ++j) {
      // S(j) = Q K(j)^T overwrites P(j-1): issued behind P V(j-1), the tensor pipe runs in order
      mbar_wait(kv_full0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = desc_adv(dk0, s * FA_KV_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tS, desc_adv(dq0, k * 32), desc_adv(dk, k * 32), idesc_qk, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc_adv(dv0, s * FA_KV_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)  // A = P (bf16x2 packed, 8 TMEM columns per K=16 step)
          umma_ts(tO, tS + k * 8, desc_adv(dv, k * 2048), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(kv_empty0 + 8 * s);
        umma_commit(pv_done);
      }
      __syncwarp();
      if (++s == FA_KV_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const uint32_t t_row = tS + lane_bits;
    const float* kb = p.key_bias ? p.key_bias + (int64_t)b * p.Nk : nullptr;
    const float sl2 = p.scale_log2;
    float m_used = -INFINITY;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int j = 0; j < T; ++j) {
      const int key0 = j * FA_BN;
      const bool general = (kb != nullptr) || (key0 + FA_BN > p.Nk);  // bias or ragged last step
      if (general) {
        // stage this step's per-key term (first barrier: everyone has finished reading the previous step's)
        named_bar_sync(1, 128);
        if (row < FA_BN) {
          const int key = key0 + row;
          kb_stage[row] = key < p.Nk ? (kb ? kb[key] * kLog2e : 0.f) : -INFINITY;
        }
        named_bar_sync(1, 128);
      }
      const float mul = general ? 1.f : sl2;
      // S(j) complete also means P V(j-1) complete (the tensor pipe runs in order): O is consistent and idle
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max, 32 columns at a time
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < FA_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        if (general) chunk_add_bias(r, sl2, kb_stage + c * 32);
        chunk_max(r, a0, a1, a2, a3);
      }
      const float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * mul;
      const float m_new = fmaxf(m_used, mx);
      const bool need = m_new > m_used + 8.f;
      if (__any_sync(0xffffffffu, need)) {
        // lazy rescale: O and l follow the running max only when it has grown by more than 2^8
        const float alpha = ex2_approx(m_used - m_new);  // 0 on the first step (m_used = -inf)
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
      const float neg_m = -m_used;
      // pass 2: P = exp2(x - m), row sum; P chunk c (32 keys -> 16 packed columns) overwrites S columns
      // [16c, 16c+16), which belong to S chunk c/2 <= c and have already been read in this pass
#pragma unroll
      for (int c = 0; c < FA_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        if (general) chunk_add_bias(r, sl2, kb_stage + c * 32);
        chunk_exp(r, mul, neg_m, t_row + c * 16, l0, l1, l2, l3);
      }
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
    if (q < p.Nq && p.lse) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = (m_used + log2f(l)) * 0.6931471805599453f;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tO + lane_bits + c * 32, r);
      tmem_ld_wait();
      if (q < p.Nq) {
        bf16* orow = p.O + ((int64_t)b * p.Nq + q) * p.ldo + h * 64 + c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv_l, __uint_as_float(r[g * 8 + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv_l, __uint_as_float(r[g * 8 + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv_l, __uint_as_float(r[g * 8 + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv_l, __uint_as_float(r[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + g * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// =====================================================================================================================
// fa_fwd_db_kernel — the attn1 forward (many key steps): S double-buffered in TMEM, Q resident in TMEM.
//
// What the round-1 profile showed for the kernel above (ncu warp-state samples, profiles/r04): the four softmax warps of
// a CTA spend 34 % of their time parked on `s_full` -- with P written in place over S, the tensor pipe cannot start
// Q K(j+1)^T before softmax(j) has finished and P V(j) has consumed it, so every step pays MMA latency + two barrier
// round trips serially -- and the MMA warp waits for P 70 % of its time.  Four CTAs per SM hide part of it.
// Here the dependency is removed instead of hidden:
//   TMEM (256 columns, 2 CTAs / SM):  S0 64 | S1 64 | O 64 | Q 32 (packed bf16, A operand)
//   MMA warp:  Q K(0)^T, Q K(1)^T up front; then per step j:  wait P(j) -> O += P(j) V(j) -> S(j+2) = Q K(j+2)^T
//              into the buffer P(j) just left (in-order tensor pipe).  S(j+1) is complete before softmax(j) ends.
//   softmax:   ONE TMEM read of the 64 scores of a row into registers (max, exp2, row sum, pack, one TMEM store).
//   Q in TMEM: both MMAs are TS-form (32 clk at N = 64 instead of 48 for the shared-memory/shared-memory form,
//              profiles/r02_probes_mma_tmem.md): 256 tensor clocks per 128 x 64 step against 512 of the exp unit.
//   O rescale (lazy, rare): the softmax warp waits for P V(j-1) through `pv_done` only when it actually rescales.
// K/V steps travel through a 4-stage TMA ring.
// =====================================================================================================================
#ifndef FA2_STAGES
#define FA2_STAGES 4
#endif
#ifndef FA2_POLY
#define FA2_POLY 2   // of every 8 exponentials, this many run as a degree-3 polynomial on the FMA pipe (0..4)
#endif
#ifndef FA2_POLY_PAIRS
#define FA2_POLY_PAIRS 1   // > 0: that many PAIRS of every 8 exponentials through the packed polynomial (replaces FA2_POLY)
#endif
#ifndef FA2_F32X2
#define FA2_F32X2 1  // FFMA2 / FADD2 (fp32x2) for the score scaling and the row sum
#endif
#ifndef FA2_TPR
#define FA2_TPR 1      // threads per query row in fa_fwd_db_kernel (1 or 2; 2 measured 8 % slower: 409 vs 377 us at cfg2)
#endif
#ifndef FA2_ROWSUM_MMA
#define FA2_ROWSUM_MMA 0   // row sums of P from the tensor pipe (a constant ones column appended to V) instead of 64 FADDs
#endif
constexpr int FA2_STAGE_BYTES = 2 * FA_BN * 128;                   // K and V tile of one 64-key step
constexpr int FA2_ONES_BYTES = FA_BN * 128;                        // constant B-operand atom: column 0 = 1, the rest 0
constexpr int FA2_SMEM = 16384 + FA2_STAGES * FA2_STAGE_BYTES + FA2_ONES_BYTES + 256 + 256 + 2048;  // + row max / sum exchange
constexpr int FA2_THREADS = 64 + 128 * FA2_TPR;
constexpr int FA2_MIN_KEYS = 512;                                  // below: the 4-CTA kernel above (attn2)
constexpr int FA2_O_COLS = FA2_ROWSUM_MMA ? 80 : 64;               // O | row-sum column (+15 unused) in TMEM

// 2^x for x <= 0 on the FMA pipe: round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 polynomial of 2^f,
// exponent added in the integer domain.  The coefficients are the Taylor ones (ln2, ln2^2/2, ln2^3/6): max rel. error
// 7.9e-4 at |f| = 0.5, always low, against the 2e-3 bf16 rounding of P it feeds (tests/test_device_math_host.py runs
// this function on the host).  The minimax fit with the constant pinned at 1 -- 0.69328293, 0.24221096, 0.05500893 --
// gives 1.0e-4 for the same instructions (-DB200_EX2_MINIMAX); not the default without a GPU run of the parity tests behind it.
#ifdef B200_EX2_MINIMAX   // experiment switch (build.py variant): the minimax constants, 1.0e-4
constexpr float EX2_C1 = 0.69328293f, EX2_C2 = 0.24221096f, EX2_C3 = 0.05500893f;
#else
constexpr float EX2_C1 = 0.6931472f, EX2_C2 = 0.2402265f, EX2_C3 = 0.0555041f;
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.f);
  const float t = x + 12582912.f;                // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);
  float p = fmaf(EX2_C3, f, EX2_C2);
  p = fmaf(p, f, EX2_C1);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// the same for a pair, on packed fp32x2 instructions (FADD2 / FFMA2): ~6 issue slots per exponential instead of 8
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -120.f);
  x.y = fmaxf(x.y, -120.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, nmagic);
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(EX2_C3, EX2_C3), f, make_float2(EX2_C2, EX2_C2));
  p = __ffma2_rn(p, f, make_float2(EX2_C1, EX2_C1));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// GENERAL: a per-key bias and / or a ragged last key step (the per-key term is staged in shared memory per step)
//   TPR (threads per query row) = 2: eight softmax warps per CTA, the two warps of a TMEM lane quadrant each take 32 of
//   the 64 scores of a row (row max exchanged through shared memory behind one 64-thread named barrier per step): four
//   softmax warps per scheduler instead of two.  With one thread per row neither the exp unit (61 %) nor the issue slots
//   (56 %) were saturated -- two warps per scheduler cannot cover each other's dependent-issue latencies.
template <bool GENERAL, int TPR>
__global__ void __launch_bounds__(64 + 128 * TPR, 2)
fa_fwd_db_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FaFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sQ = sbase, sKV = sbase + 16384;
  const uint32_t sOnes = sKV + FA2_STAGES * FA2_STAGE_BYTES;
  const uint32_t bar = sOnes + FA2_ONES_BYTES;
  const uint32_t q_full = bar, q_tmem = bar + 8, pv_done0 = bar + 16 /*2*/, s_full0 = bar + 32 /*2*/, p_full0 = bar + 48 /*2*/,
                 kv_full0 = bar + 64, kv_empty0 = kv_full0 + 8 * FA2_STAGES, tmem_slot = kv_empty0 + 8 * FA2_STAGES;
  // P V(j) completes phase (j >> 1) of pv_done[j & 1].  Two barriers because a softmax warp only looks at them when it
  // rescales O: with one barrier its parity test could not tell "P V(j-1) done" from "P V(j-3) done, two phases behind"
  // (seen on hardware: the epilogue read O before the last P V had landed).  With two, the previous phase of the
  // barrier it waits on is P V(j-3), which s_full(j) already implies (the score MMA of step j is issued only after the
  // MMA warp has observed P V(j-2), see below).
  float* kb_stage = reinterpret_cast<float*>(smem_raw + (bar + 256 - sbase));  // [64] per-key term of the current step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item and key-step range of this CTA
  int item = blockIdx.x, part = -1, j_begin = 0, j_end = p.kv_tiles;
  if (item >= p.n_whole) {
    const int e = item - p.n_whole;
    item = p.n_whole + e / p.parts;
    part = e % p.parts;
    j_begin = (int)((int64_t)part * p.kv_tiles / p.parts);
    j_end = (int)((int64_t)(part + 1) * p.kv_tiles / p.parts);
  }
  const int qt = item % p.q_tiles, h = (item / p.q_tiles) % p.H, b = item / (p.q_tiles * p.H);
  pdl_launch();
  if (p.batch_keep != nullptr) {
    pdl_wait();
    if (p.batch_keep[b] == 0.f) {
      if (part <= 0) fa_copy_values(p, qt, h, b, 64 + 128 * TPR);
      return;
    }
  }

  if (FA2_ROWSUM_MMA) {
    // the ones atom, in the layout TMA gives a V tile (64 key rows of 128 swizzled bytes): element (row, column 0) = 1
    for (int i = threadIdx.x; i < FA2_ONES_BYTES / 16; i += 64 + 128 * TPR) {
      const int r = i >> 3, c = i & 7;
      const uint32_t v0 = (c == (r & 7)) ? 0x00003F80u : 0u;   // bf16 1.0 in the low half: column 0 sits in chunk 0 ^ (r & 7)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(sOnes + i * 16), "r"(v0), "r"(0u) : "memory");
    }
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      printf("b200 fa_fwd_db: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_tmem, 128 * TPR);
    for (int s = 0; s < 2; ++s) {
      mbar_init(pv_done0 + 8 * s, 1);
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(p_full0 + 8 * s, 128 * TPR);
    }
    for (int s = 0; s < FA2_STAGES; ++s) {
      mbar_init(kv_full0 + 8 * s, 1);
      mbar_init(kv_empty0 + 8 * s, 1);
    }
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
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tQ = tmem_base + 128 + FA2_O_COLS;
  const int T = j_end - j_begin;   // key steps of this CTA; step j below is global step j_begin + j
  pdl_wait();   // barrier init / TMEM allocation above overlapped the previous kernel's tail

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(q_full, 16384);
    tma_load_3d(sQ, &tmQ, q_full, h * 64, qt * 128, b);
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(kv_empty0 + 8 * s, ph ^ 1);
      mbar_expect_tx(kv_full0 + 8 * s, FA2_STAGE_BYTES);
      tma_load_3d(sKV + s * FA2_STAGE_BYTES, &tmK, kv_full0 + 8 * s, h * 64, (j_begin + j) * FA_BN, b);
      tma_load_3d(sKV + s * FA2_STAGE_BYTES + FA_BN * 128, &tmV, kv_full0 + 8 * s, h * 64, (j_begin + j) * FA_BN, b);
      if (++s == FA2_STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc_qk = make_idesc_bf16(128, FA_BN, 0, 0);        // A = Q (TMEM), B = K tile, K-major
    const uint32_t idesc_pv = make_idesc_bf16(128, FA2_O_COLS, 0, 1);   // A = P (TMEM), B = [V | ones], MN-major
    const uint64_t dk0 = make_smem_desc(sKV, 16, 1024);
    mbar_wait(q_tmem, 0);
    tc_fence_after();
    int s_qk = 0;            // ring stage of the next Q K^T
    uint32_t ph_qk = 0;
    auto issue_qk = [&](int j) {
      mbar_wait(kv_full0 + 8 * s_qk, ph_qk);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = desc_adv(dk0, s_qk * FA2_STAGE_BYTES);
        const uint32_t d = tS + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(d, tQ + k * 8, desc_adv(dk, k * 32), idesc_qk, k > 0 ? 1u : 0u);
        umma_commit(s_full0 + 8 * (j & 1));
      }
      __syncwarp();
      if (++s_qk == FA2_STAGES) { s_qk = 0; ph_qk ^= 1; }
    };
    issue_qk(0);
    if (T > 1) issue_qk(1);
    int s_pv = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(p_full0 + 8 * (j & 1), (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        // second 64-wide MN atom of the B operand (columns 64..79 of the N = 80 tile) = the constant ones atom: the
        // leading-dimension byte offset points from this stage's V tile to it, the k advance walks both alike
        const uint32_t sV = sKV + s_pv * FA2_STAGE_BYTES + FA_BN * 128;
        const uint64_t dv = make_smem_desc(sV, FA2_ROWSUM_MMA ? (sOnes - sV) : 8192, 1024);
        const uint32_t a = tS + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)  // A = P (bf16x2 packed, 8 TMEM columns per K = 16 step)
          umma_ts(tO, a + k * 8, desc_adv(dv, k * 2048), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(kv_empty0 + 8 * s_pv);
        umma_commit(pv_done0 + 8 * (j & 1));
      }
      __syncwarp();
      if (++s_pv == FA2_STAGES) s_pv = 0;
      if (j + 2 < T) {
        // S(j+2) overwrites P(j): the score MMA is issued once P V(j) has completed (not merely been issued).  The
        // tensor pipe has the slack (256 of ~600 clk per step; measured: no cost), S(j+2) is still ready a whole softmax
        // step early, and every observer of s_full(j+2) thereby knows that P V(j) is done.
        mbar_wait(pv_done0 + 8 * (j & 1), (j >> 1) & 1);
        issue_qk(j + 2);
      }
    }
  } else if (warp >= 2) {
    constexpr int NC = 64 / TPR;        // scores of a row held by this thread
    constexpr int NCH = NC / 32;        // ... in 32-column TMEM chunks
    const int quad = warp & 3;
    const int half = TPR == 2 ? (warp - 2) >> 2 : 0;   // which 32 of the 64 key columns (TPR = 2)
    const int row = quad * 32 + lane;
    const int st = threadIdx.x - 64;    // softmax-thread index
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    float* xch = kb_stage + 64;         // [2 steps][2 halves][128 rows] row-max / row-sum exchange (TPR = 2)
    // Q row -> TMEM as the packed-bf16 A operand: 128 bytes = 8 swizzled 16-byte chunks = 32 columns
    mbar_wait(q_full, 0);
    {
      uint32_t rq[32 / TPR];
#pragma unroll
      for (int c = 0; c < 8 / TPR; ++c)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rq[c * 4]), "=r"(rq[c * 4 + 1]), "=r"(rq[c * 4 + 2]), "=r"(rq[c * 4 + 3])
                     : "r"(sQ + sw128_off(row, half * 4 + c)));
      if (TPR == 1) tmem_st32(tQ + lane_bits, *reinterpret_cast<uint32_t(*)[32]>(rq));
      else tmem_st16(tQ + lane_bits + half * 16, *reinterpret_cast<uint32_t(*)[16]>(rq));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_tmem);
    }
    const float* kb = p.key_bias ? p.key_bias + (int64_t)b * p.Nk : nullptr;
    const float sl2 = p.scale_log2;
    const float mul = GENERAL ? 1.f : sl2;
    float m_used = -INFINITY;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
      if (GENERAL) {
        const int key0 = (j_begin + j) * FA_BN;
        named_bar_sync(1, 128 * TPR);
        if (st < FA_BN) {
          const int key = key0 + st;
          kb_stage[st] = key < p.Nk ? (kb ? kb[key] * kLog2e : 0.f) : -INFINITY;
        }
        named_bar_sync(1, 128 * TPR);
      }
      const uint32_t t_buf = tS + (j & 1) * 64 + lane_bits;
      mbar_wait(s_full0 + 8 * (j & 1), (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(t_buf + half * 32 + c * 32, r[c]);
      tmem_ld_wait();
      if (GENERAL) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) chunk_add_bias(r[c], sl2, kb_stage + half * 32 + c * 32);
      }
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < NCH; ++c) chunk_max(r[c], a0, a1, a2, a3);
      float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * mul;
      if (TPR == 2) {
        // the other half of the row lives in the partner warp of this lane quadrant
        float* slot = xch + (j & 1) * 256;
        slot[half * 128 + row] = mx;
        named_bar_sync(2 + quad, 64);
        mx = fmaxf(mx, slot[(half ^ 1) * 128 + row]);
      }
      const float m_new = fmaxf(m_used, mx);
      const bool need = m_new > m_used + 8.f;
      if (__any_sync(0xffffffffu, need)) {
        // lazy rescale: O and l follow the running max only when it has grown by more than 2^8 (both threads of a
        // row see the same m_used / m_new, so both warps of a quadrant take this branch together)
        const float alpha = ex2_approx(m_used - m_new);  // 0 on the first step (m_used = -inf)
        m_used = m_new;
        l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
        if (j > 0) {
          // O holds P V(0 .. j-1) and the tensor pipe is not writing it
          mbar_wait(pv_done0 + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < FA2_O_COLS / 16 / TPR; ++c) {
            uint32_t ro[16];
            const uint32_t ta = tO + lane_bits + half * (FA2_O_COLS / 2) + c * 16;
            tmem_ld16(ta, ro);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
            tmem_st16(ta, ro);
          }
        }
      }
      const float neg_m = -m_used;
      uint32_t pk[NC / 2];
      // packed fp32x2 arithmetic (FFMA2 / FADD2): x = s * mul - m and the row sum take one instruction per PAIR of
      // scores -- the softmax warps are issue-bound between their exponentials, not FMA-pipe bound
      const float2 mul2 = make_float2(mul, mul), negm2 = make_float2(neg_m, neg_m);
      float2 la = make_float2(l0, l1), lb = make_float2(l2, l3);
#pragma unroll
      for (int hf = 0; hf < NCH; ++hf) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float2 pv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[hf][g * 8 + 2 * i]),
                                                    __uint_as_float(r[hf][g * 8 + 2 * i + 1])), mul2, negm2);
            if (i < FA2_POLY_PAIRS) {
              pv[i] = ex2_poly2(x);
            } else {
              pv[i].x = ex2_approx(x.x);
              pv[i].y = ex2_approx(x.y);
            }
          }
          la = __fadd2_rn(la, __fadd2_rn(pv[0], pv[2]));
          lb = __fadd2_rn(lb, __fadd2_rn(pv[1], pv[3]));
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[hf * 16 + g * 4 + i] = pack_bf16x2(pv[i].x, pv[i].y);
        }
      }
      l0 = la.x; l1 = la.y; l2 = lb.x; l3 = lb.y;
      // P(j): 64 keys as 32 packed columns over the first half of S(j); this thread's keys -> its NC / 2 columns
      if (TPR == 1) tmem_st32(t_buf, *reinterpret_cast<uint32_t(*)[32]>(pk));
      else tmem_st16(t_buf + half * 16, *reinterpret_cast<uint32_t(*)[16]>(pk));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full0 + 8 * (j & 1));
    }
    mbar_wait(pv_done0 + 8 * ((T - 1) & 1), ((T - 1) >> 1) & 1);
    tc_fence_after();
    float l = (l0 + l1) + (l2 + l3);
    if (TPR == 2) {   // row sum = this thread's half + the partner's
      float* slot = xch + (T & 1) * 256;
      slot[half * 128 + row] = l;
      named_bar_sync(2 + quad, 64);
      l += slot[(half ^ 1) * 128 + row];
    }
    const int q = qt * 128 + row;
    constexpr int OC = 64 / TPR;   // O columns this thread writes out
    if (part >= 0) {
      // one of several CTAs on this item: leave the un-normalised accumulator and the softmax state for the merge
      float* dst = p.part_ws + ((((int64_t)(item - p.n_whole) * p.parts + part) * 128) + row) * FA2_PART_LD;
#pragma unroll 1
      for (int c = 0; c < OC / 16; ++c) {
        uint32_t ro[16];
        tmem_ld16(tO + lane_bits + half * OC + c * 16, ro);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          *reinterpret_cast<float2*>(dst + half * OC + c * 16 + i) = make_float2(__uint_as_float(ro[i]), __uint_as_float(ro[i + 1]));
      }
      if (half == 0) *reinterpret_cast<float2*>(dst + 64) = make_float2(m_used, l);
    } else {
      const float inv_l = 1.f / l;
      if (half == 0 && q < p.Nq && p.lse) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = (m_used + log2f(l)) * 0.6931471805599453f;
#pragma unroll 1
      for (int c = 0; c < OC / 32; ++c) {
        uint32_t ro[32];
        tmem_ld32(tO + lane_bits + half * OC + c * 32, ro);
        tmem_ld_wait();
        if (q < p.Nq) {
          bf16* orow = p.O + ((int64_t)b * p.Nq + q) * p.ldo + h * 64 + half * OC + c * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ro[g * 8 + 0]) * inv_l, __uint_as_float(ro[g * 8 + 1]) * inv_l);
            u.y = pack_bf16x2(__uint_as_float(ro[g * 8 + 2]) * inv_l, __uint_as_float(ro[g * 8 + 3]) * inv_l);
            u.z = pack_bf16x2(__uint_as_float(ro[g * 8 + 4]) * inv_l, __uint_as_float(ro[g * 8 + 5]) * inv_l);
            u.w = pack_bf16x2(__uint_as_float(ro[g * 8 + 6]) * inv_l, __uint_as_float(ro[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + g * 8) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Fold the `parts` partial results of every split item: one warp per query row, two output columns per lane.
__global__ void __launch_bounds__(256) fa_fwd_merge_kernel(const FaFwdParams p, int n_split) {
  pdl_launch();
  pdl_wait();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_split * 128) return;
  const int si = gw >> 7, row = gw & 127;
  const int item = p.n_whole + si;
  const int qt = item % p.q_tiles, h = (item / p.q_tiles) % p.H, b = item / (p.q_tiles * p.H);
  const int q = qt * 128 + row;
  if (q >= p.Nq) return;
  if (p.batch_keep != nullptr && p.batch_keep[b] == 0.f) return;   // the values were passed through by the main kernel
  const float* src = p.part_ws + (((int64_t)si * p.parts) * 128 + row) * FA2_PART_LD;
  float M = -INFINITY;
  for (int t = 0; t < p.parts; ++t) M = fmaxf(M, src[(int64_t)t * 128 * FA2_PART_LD + 64]);
  float L = 0.f, o0 = 0.f, o1 = 0.f;
  for (int t = 0; t < p.parts; ++t) {
    const float* s = src + (int64_t)t * 128 * FA2_PART_LD;
    const float w = exp2f(s[64] - M);
    L = fmaf(s[65], w, L);
    const float2 v = *reinterpret_cast<const float2*>(s + 2 * lane);
    o0 = fmaf(v.x, w, o0);
    o1 = fmaf(v.y, w, o1);
  }
  const float inv = 1.f / L;
  *reinterpret_cast<uint32_t*>(p.O + ((int64_t)b * p.Nq + q) * p.ldo + h * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
  if (lane == 0 && p.lse) p.lse[((int64_t)b * p.H + h) * p.Nq + q] = (M + log2f(L)) * 0.6931471805599453f;
}

// How the work items of the attn1 forward are dealt out: `slots` CTAs run concurrently (2 per SM); when the last wave is
// sparsely filled (cfg2: 1536 items on 296 slots = 5.19 waves, 56 CTAs on 148 SMs for a whole CTA lifetime) its items
// are split along the keys into `parts` CTAs each, so that the tail costs 1 / parts of a wave.
static void fa_fwd_plan(int items, int kv_tiles, int* n_whole, int* parts) {
  static int slots = 0;
  if (!slots) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    slots = 2 * sms;
  }
  *n_whole = items;
  *parts = 1;
  const int r = items % slots;
  if (items < slots || r == 0 || r * 4 > slots * 3) return;   // a single wave, or a well-filled last wave
  // r items as r * pr CTAs of 1/pr of the keys each: rounds of `slots` CTAs, each 1/pr of a full CTA lifetime, plus
  // what a part costs (Q load, fp32 partial and the merge); un-split = 1.0
  int max_pr = kv_tiles / 8 < 8 ? kv_tiles / 8 : 8;           // at least 8 key steps per part
  int pr = 1;
  double best = 1.0;
  for (int c = 2; c <= max_pr; ++c) {
    const double cost = (double)((r * c + slots - 1) / slots) / c + 0.015 * c;
    if (cost < best - 1e-9) { best = cost; pr = c; }
  }
  if (pr < 2) return;
  *n_whole = items - r;
  *parts = pr;
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

// kernel-experiment switch (tools/attn_bench.py): B200_FA_FWD_SMALL=1 routes every shape to the 4-CTA kernel
static const bool b200_fa_fwd_force_small = [] {
  const char* e = getenv("B200_FA_FWD_SMALL");
  return e && e[0] == '1';
}();

extern "C" int64_t b200_fa_fwd_workspace_bytes(int B, int H, int Nq, int Nk) {
  if (B <= 0 || H <= 0 || Nq <= 0 || Nk < FA2_MIN_KEYS || b200_fa_fwd_force_small) return 0;
  int n_whole, parts;
  const int items = ((Nq + 127) / 128) * H * B;
  fa_fwd_plan(items, (Nk + FA_BN - 1) / FA_BN, &n_whole, &parts);
  return (int64_t)(items - n_whole) * parts * 128 * FA2_PART_LD * (int64_t)sizeof(float);
}

extern "C" int b200_fa_fwd_ws(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                              int64_t ldo, float* lse, const float* key_bias, const float* batch_keep, const void* pass_src,
                              int64_t ld_pass, int B, int H, int Nq, int Nk, int head_dim, float scale, void* workspace,
                              int64_t workspace_bytes, void* stream);

extern "C" int b200_fa_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, void* o, int64_t ldo, float* lse, const float* key_bias,
                           int B, int H, int Nq, int Nk, int head_dim, float scale, void* stream) {
  return b200_fa_fwd_ws(q, ldq, k, ldk, v, ldv, o, ldo, lse, key_bias, nullptr, nullptr, 0, B, H, Nq, Nk, head_dim, scale,
                        nullptr, 0, stream);
}

extern "C" int b200_fa_fwd_ws(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                              int64_t ldo, float* lse, const float* key_bias, const float* batch_keep, const void* pass_src,
                              int64_t ld_pass, int B, int H, int Nq, int Nk, int head_dim, float scale, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  if (head_dim != 64) return arg_error("fa_fwd: only head_dim 64 is built (LTXV-2B: 32 heads x 64)");
  if (B < 0 || H <= 0 || Nq < 0 || Nk < 0) return arg_error("fa_fwd: bad shape");
  if (B == 0 || Nq == 0) return 0;  // no queries: nothing to do (empty tensors carry null pointers)
  if (!(q && k && v && o)) return arg_error("fa_fwd: null pointer");
  if (Nk == 0) return arg_error("fa_fwd: no keys");
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || !al16(q) || !al16(k) || !al16(v) || !al16(o))
    return arg_error("fa_fwd: tensors must be 16-byte aligned with 16-byte-multiple pitches");
  if (ldq < H * 64 || ldk < H * 64 || ldv < H * 64 || ldo < H * 64) return arg_error("fa_fwd: pitch < H*64");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_tokens(&tmQ, q, B, Nq, ldq, H * 64, 128)) ||
      (rc = make_tmap_tokens(&tmK, k, B, Nk, ldk, H * 64, FA_BN)) ||
      (rc = make_tmap_tokens(&tmV, v, B, Nk, ldv, H * 64, FA_BN)))
    return arg_error("fa_fwd: cuTensorMapEncodeTiled failed", rc);
  FaFwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.kv_tiles = (Nk + FA_BN - 1) / FA_BN;
  p.O = (bf16*)o; p.ldo = ldo; p.lse = lse; p.key_bias = key_bias;
  p.scale_log2 = scale * kLog2e;
  p.q_tiles = (Nq + 127) / 128; p.n_whole = 0; p.parts = 1; p.part_ws = nullptr;
  p.batch_keep = batch_keep;
  p.pass_src = pass_src ? (const bf16*)pass_src : (const bf16*)v;   // default: the value rows (needs Nq == Nk)
  p.ld_pass = pass_src ? ld_pass : ldv;
  if (batch_keep != nullptr && pass_src == nullptr && Nq != Nk)
    return arg_error("fa_fwd: batch_keep with the value rows as pass-through source needs Nq == Nk");
  if (batch_keep != nullptr && (p.ld_pass % 8 || p.ld_pass < H * 64 || !al16(p.pass_src)))
    return arg_error("fa_fwd: pass-through source must be 16-byte aligned with a 16-byte-multiple pitch >= H*64");
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_FWD_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(fa_fwd_db_kernel<false, FA2_TPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA2_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(fa_fwd_db_kernel<true, FA2_TPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA2_SMEM) != cudaSuccess)
      return launch_status("fa_fwd: cudaFuncSetAttribute");
    cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(fa_fwd_db_kernel<false, FA2_TPR>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(fa_fwd_db_kernel<true, FA2_TPR>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  dim3 grid((Nq + 127) / 128, H, B);
  // many key steps (attn1): S double-buffered in TMEM; a few key steps (attn2): four small CTAs per SM
  if (Nk >= FA2_MIN_KEYS && !b200_fa_fwd_force_small) {
    const int items = (int)(grid.x * grid.y * grid.z);
    p.q_tiles = (int)grid.x;
    p.n_whole = items;
    p.parts = 1;
    p.part_ws = nullptr;
    if (workspace != nullptr) {   // without a workspace the last wave simply runs un-split
      fa_fwd_plan(items, p.kv_tiles, &p.n_whole, &p.parts);
      const int64_t need = (int64_t)(items - p.n_whole) * p.parts * 128 * FA2_PART_LD * (int64_t)sizeof(float);
      if (need > workspace_bytes || !al16(workspace)) return arg_error("fa_fwd: workspace too small (see b200_fa_fwd_workspace_bytes)");
      p.part_ws = (float*)workspace;
    }
    const int n_split = items - p.n_whole;
    const unsigned ctas = (unsigned)(p.n_whole + n_split * p.parts);
    if (key_bias != nullptr || Nk % FA_BN != 0)
      { auto kdb = fa_fwd_db_kernel<true, FA2_TPR>; B200_LAUNCH(kdb, ctas, 64 + 128 * FA2_TPR, FA2_SMEM, stream, tmQ, tmK, tmV, p); }
    else
      { auto kdb = fa_fwd_db_kernel<false, FA2_TPR>; B200_LAUNCH(kdb, ctas, 64 + 128 * FA2_TPR, FA2_SMEM, stream, tmQ, tmK, tmV, p); }
    if (n_split > 0)
      B200_LAUNCH(fa_fwd_merge_kernel, (n_split * 128 * 32 + 255) / 256, 256, 0, stream, p, n_split);
  } else
    B200_LAUNCH(fa_fwd_kernel, grid, FA_FWD_THREADS, FA_FWD_SMEM, stream, tmQ, tmK, tmV, p);
  return launch_status("fa_fwd");
}
