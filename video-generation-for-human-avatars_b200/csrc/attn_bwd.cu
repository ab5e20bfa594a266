// attn_bwd.cu — tcgen05 / TMEM flash-attention backward, head_dim 64, non-causal (sm_100a).
//
// Gradient of softmax(q k^T * scale + key_bias) v with respect to q, k, v — what autograd runs
// through F.scaled_dot_product_attention in the reference's LoRA train step
// (training.py:203 -> attention.py:1057).  Same token-major [B*N, ld] layout as attn_fwd.cu.
//
// One CTA owns one 128-key tile of one head and walks over the 128-query tiles.  Everything is
// computed transposed (keys on TMEM lanes) so that each compute thread owns one key row:
//   S^T  = K Q^T              (A = K  resident in TMEM,     B = Q  tile, K-major)
//   dP^T = V dO^T             (A = V  resident in TMEM,     B = dO tile, K-major)
//   P^T  = exp2(S^T*c + bias - lse) ; dS^T = P^T * (dP^T - delta) * scale
//   dV  += P^T  dO            (A = P^T  packed bf16 in TMEM, B = dO tile read MN-major)
//   dK  += dS^T Q             (A = dS^T packed bf16 in TMEM, B = Q  tile read MN-major)
//   dQ   = dS   K             (A = dS^T bytes in shared memory read MN-major, B = K tile read MN-major)
// The tensor pipe is fed from shared memory at well under 128 B/clk, and at head_dim 64 a 128x64x16 MMA
// needs 6 KB of operands per 32 math cycles when both come from shared memory: operand fetch, not math,
// was the limit of the first versions (176 KB of shared-memory operand reads per tile pair).  Hence
// every A operand that can live in TMEM does: K and V are copied there once per CTA, P^T and dS^T
// overwrite the S^T / dP^T columns they were computed from; only dQ reads dS^T from shared memory
// (the transposed view needs the MN-major descriptor).  112 KB of operand reads per tile pair.
// dV/dK accumulate in TMEM over the whole loop; dQ is a per-(q tile, k tile) partial that is
// reduced across key-tile CTAs with TMA reduce-add (cp.reduce.async.bulk.tensor) into a caller-zeroed
// fp32 buffer.
// TMEM: S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ 64 | K 32 | V 32 = 512 columns.
//
// Pipeline at HALF-tile granularity (64 queries): S^T / dP^T are two independent 64-column halves with
// their own barriers.  While the 16 compute warps turn half h into P^T / dS^T, the tensor pipe runs
// dV/dK of half h-1 and the scores of half h+1; the scores of half h+2 go behind dV/dK of half h (they
// overwrite P^T/dS^T(h)).  Q/dO tiles are triple-buffered so the TMA latency never reaches the MMA warp.
// The key-tile CTAs of a head run concurrently and all reduce their dQ partials into the same rows: CTA
// kt therefore walks the query tiles starting at tile kt (mod T), so that at any moment the concurrent
// reduce-adds hit different addresses (the L2 atomic unit serialises per address).
// Four further warps do everything that is not arithmetic: they stream lse/delta into shared memory two
// tiles ahead and drain dQ(i) (TMEM -> per-warp swizzled staging -> TMA reduce-add) without a CTA-wide
// barrier, so the compute warps execute nothing but the element-wise math.
#include <stdlib.h>

#include "api_internal.h"
#include "common.cuh"
#include "tmap.h"

namespace b200 {

int make_tmap_tokens(CUtensorMap* out, const void* base, int B, int N, int64_t ld, int width,
                     int box_rows);

struct FaBwdParams {
  int B, H, Nq, Nk, q_tiles;
  const float* lse;    // [B,H,Nq]
  const float* delta;  // [B,H,Nq]
  const float* key_bias;
  float* dq;  // fp32 [B*Nq, lddq], caller-zeroed
  int64_t lddq;
  bf16* dk;
  int64_t lddk;
  bf16* dv;
  int64_t lddv;
  float scale, scale_log2;
  int q_splits;     // > 1: the query tiles are divided among q_splits CTAs per key tile (few key tiles: attn2)
  float* part_dk;   // [q_splits][B*Nk][H*64] fp32 partial dK / dV, summed by fa_bwd_reduce_kernel
  float* part_dv;
  // many key tiles (attn1): 1-D grid over the (key tile, head, batch) items; the items of a sparsely filled LAST wave
  // (linear id >= n_whole) are each walked by `tail_parts` CTAs over disjoint query-tile ranges whose fp32 dK / dV
  // partials go to tail_ws [(item - n_whole) * tail_parts + part][128 rows][dK 64 | dV 64] (fa_bwd_tail_reduce_kernel)
  int k_tiles, n_whole, tail_parts;
  float* tail_ws;
};

#ifndef FAB_F32X2
#define FAB_F32X2 1
#endif
constexpr int FA_MASK_SCAN_MAX = 1024;  // key counts up to which the bias vector is scanned for masked key tiles
constexpr int FA_BWD_QSTAGES = 3;
constexpr int FA_BWD_SMEM = 16384 * 2 /*K,V*/ + FA_BWD_QSTAGES * 32768 /*Q,dO*/ + 2 * 32768 /*dS^T x2*/ +
                            32768 /*dQ staging*/ + 2 * 2 * 512 /*lse, delta x2*/ + 256;
constexpr float kLog2eB = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifdef B200_TRACE
__device__ unsigned long long g_bwd_trace[8192];
#define TRACE(slot)                                                                       \
  do {                                                                                    \
    if (blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0) \
      g_bwd_trace[(slot)] = clock64();                                                    \
  } while (0)
#else
#define TRACE(slot) do {} while (0)
#endif

// mbarrier arrivals of the compute / drain warps: per thread or one per warp (experiment switch)
#ifdef B200_CW_PER_WARP
#define CW_ARRIVALS 16
#define CW_ARRIVE(bar_) do { __syncwarp(); if (lane == 0) mbar_arrive(bar_); } while (0)
#else
#define CW_ARRIVALS 512
#define CW_ARRIVE(bar_) mbar_arrive(bar_)
#endif
#ifdef B200_DW_PER_WARP
#define DW_ARRIVALS 4
#define DW_ARRIVE(bar_) do { __syncwarp(); if (lane == 0) mbar_arrive(bar_); } while (0)
#else
#define DW_ARRIVALS 128
#define DW_ARRIVE(bar_) mbar_arrive(bar_)
#endif

constexpr int FA_BWD_THREADS = 64 + 512 + 128;  // TMA warp, MMA warp, 16 compute warps, 4 drain/stat warps

__global__ void __launch_bounds__(FA_BWD_THREADS, 1)
fa_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
              const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ FaBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sK = sbase, sV = sbase + 16384, sQdO = sbase + 32768;  // stage s: Q at +s*32768, dO at +16384
  // dS^T for the dQ MMA is double-buffered in shared memory (buffer = tile & 1)
  const uint32_t sDS = sQdO + FA_BWD_QSTAGES * 32768;
  const uint32_t sStage = sDS + 65536;   // fp32 dQ staging: one 8 KB (32 rows x 64 fp32) region per drain warp
  const uint32_t sStat = sStage + 32768;  // [2 stages][lse 128 | delta 128] fp32
  const uint32_t bar = sStat + 2048;
  const uint32_t kv_full = bar, qd_full0 = bar + 8, qd_empty0 = bar + 32, s_full0 = bar + 56,
                 pds_full0 = bar + 72, mma2_done = bar + 88, dq_free = bar + 96, stat_full0 = bar + 104,
                 all_done = bar + 120, kv_tmem = bar + 128, tmem_slot = bar + 136;
  float* stat = reinterpret_cast<float*>(smem_raw + (sStat - sbase));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kt, h, b, split, splits = p.q_splits, tail_slot = -1;
  if (p.k_tiles > 0) {   // 1-D grid (attn1)
    int item = blockIdx.x;
    split = 0;
    if (item >= p.n_whole) {
      const int e = item - p.n_whole;
      item = p.n_whole + e / p.tail_parts;
      split = e % p.tail_parts;
      splits = p.tail_parts;
      tail_slot = e;
    }
    kt = item % p.k_tiles;
    h = (item / p.k_tiles) % p.H;
    b = item / (p.k_tiles * p.H);
  } else {
    kt = blockIdx.x;
    h = blockIdx.y;
    b = blockIdx.z / p.q_splits;
    split = blockIdx.z % p.q_splits;
  }
  // this CTA's query tiles: [t0, t0 + T)
  const int t0 = (int)((int64_t)split * p.q_tiles / splits);
  const int T = (int)((int64_t)(split + 1) * p.q_tiles / splits) - t0;

  // A key tile whose keys all carry a bias <= -9000 (the reference's -10000 mask, transformer3d.py:440-445) while
  // some other key of the batch entry is unmasked has P = 0 exactly (exp of < -900 underflows in fp32): dK = dV = 0
  // and no contribution to dQ.  Such CTAs write their zeros and leave before touching barriers or TMEM.  (With every
  // key masked the softmax is uniform, not zero: nothing is skipped then.)
  pdl_launch();
  if (p.key_bias != nullptr && p.Nk <= FA_MASK_SCAN_MAX) {
    pdl_wait();   // the mask scan reads (and a dead tile writes) global memory
    int live_tile = 0, live_batch = 0;   // block-wide votes: no shared memory to spare next to the dynamic 227 KB
    const float* kbp = p.key_bias + (int64_t)b * p.Nk;
    for (int key_t = threadIdx.x; key_t < p.Nk; key_t += FA_BWD_THREADS)
      if (kbp[key_t] > -9000.f) {
        live_batch = 1;
        if ((key_t >> 7) == kt) live_tile = 1;
      }
    const int s_live_tile = __syncthreads_or(live_tile);
    const int s_live_batch = __syncthreads_or(live_batch);
    const bool s_any_live = s_live_tile || !s_live_batch;
    if (!s_any_live) {
      // 128 rows x 64 head columns of dK and dV (bf16), or of this split's fp32 partials
      for (int idx = threadIdx.x; idx < 128 * 8; idx += FA_BWD_THREADS) {
        const int r = idx >> 3, c8 = (idx & 7) * 8;
        const int key_r = kt * 128 + r;
        if (key_r >= p.Nk) continue;
        if (splits == 1) {
          *reinterpret_cast<uint4*>(p.dk + ((int64_t)b * p.Nk + key_r) * p.lddk + h * 64 + c8) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(p.dv + ((int64_t)b * p.Nk + key_r) * p.lddv + h * 64 + c8) = make_uint4(0, 0, 0, 0);
        } else if (tail_slot >= 0) {
          float4* t4 = reinterpret_cast<float4*>(p.tail_ws + ((int64_t)tail_slot * 128 + r) * 128);
          t4[(c8 >> 2)] = t4[(c8 >> 2) + 1] = t4[16 + (c8 >> 2)] = t4[16 + (c8 >> 2) + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          const int64_t prow = ((int64_t)split * p.B + b) * p.Nk + key_r;
          float4* pk4 = reinterpret_cast<float4*>(p.part_dk + prow * (p.H * 64) + h * 64 + c8);
          float4* pv4 = reinterpret_cast<float4*>(p.part_dv + prow * (p.H * 64) + h * 64 + c8);
          pk4[0] = pk4[1] = pv4[0] = pv4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      return;
    }
  }

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      printf("b200 fa_bwd: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < FA_BWD_QSTAGES; ++s) {
      mbar_init(qd_full0 + 8 * s, 1);
      mbar_init(qd_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(pds_full0 + 8 * s, CW_ARRIVALS);
      mbar_init(stat_full0 + 8 * s, DW_ARRIVALS);
    }
    mbar_init(mma2_done, 1);
    mbar_init(dq_free, DW_ARRIVALS);
    mbar_init(all_done, 1);
    mbar_init(kv_tmem, CW_ARRIVALS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tSt = tmem_base, tdPt = tmem_base + 128, tdV = tmem_base + 256,
                 tdK = tmem_base + 320, tdQ = tmem_base + 384, tK = tmem_base + 448, tV = tmem_base + 480;
  pdl_wait();   // barrier init / TMEM allocation above overlapped the previous kernel's tail

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(kv_full, 32768);
    tma_load_3d(sK, &tmK, kv_full, h * 64, kt * 128, b);
    tma_load_3d(sV, &tmV, kv_full, h * 64, kt * 128, b);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < T; ++i) {
      mbar_wait(qd_empty0 + 8 * s, ph ^ 1);
      mbar_expect_tx(qd_full0 + 8 * s, 32768);
      const int qt = t0 + (i + kt) % T;  // staggered walk, see below
      tma_load_3d(sQdO + s * 32768, &tmQ, qd_full0 + 8 * s, h * 64, qt * 128, b);
      tma_load_3d(sQdO + s * 32768 + 16384, &tmdO, qd_full0 + 8 * s, h * 64, qt * 128, b);
      if (++s == FA_BWD_QSTAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // MMA warp: all 32 lanes follow the control flow (barrier waits), one elected lane issues.  Every
    // descriptor is a constant base plus a small offset, so the instruction stream per MMA is minimal.
    const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);    // A (TMEM) x B K-major, one 64-query half
    const uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);   // A (TMEM), B MN-major
    const uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major
    const uint64_t dQ_k = make_smem_desc(sQdO, 16, 1024), dQ_mn = make_smem_desc(sQdO, 8192, 1024);
    const uint64_t ddS_mn = make_smem_desc(sDS, 16384, 1024), dK_mn = make_smem_desc(sK, 8192, 1024);
    mbar_wait(kv_tmem, 0);  // K and V are resident in TMEM
    tc_fence_after();
    // scores of half-tile g (tile g>>1, queries [64*(g&1), +64)): S^T and dP^T into their half's columns
    auto issue_scores = [&](int g) {
      const int i = g >> 1, hf = g & 1, s = i % FA_BWD_QSTAGES;
      if (hf == 0) {
        mbar_wait(qd_full0 + 8 * s, (i / FA_BWD_QSTAGES) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        const uint64_t bq = desc_adv(dQ_k, s * 32768 + hf * 8192), bo = desc_adv(bq, 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tSt + hf * 64, tK + k * 8, desc_adv(bq, k * 32), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(tdPt + hf * 64, tV + k * 8, desc_adv(bo, k * 32), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full0 + 8 * hf);
      }
      __syncwarp();
    };
    const int G = 2 * T;
    if (G > 0) issue_scores(0);
    if (G > 1) issue_scores(1);
    for (int g = 0; g < G; ++g) {
      const int i = g >> 1, hf = g & 1, s = i % FA_BWD_QSTAGES;
      mbar_wait(pds_full0 + 8 * hf, i & 1);  // P^T(g), dS^T(g) written (TMEM + shared memory)
      tc_fence_after();
      TRACE(g * 4 + 0);
      if (elect_one()) {
        // P^T / dS^T of the 16 queries of k-step k sit in the first 8 of the 16 S^T / dP^T columns they
        // were computed from (each compute warp overwrites only columns it has read itself)
        const uint64_t bq = desc_adv(dQ_mn, s * 32768 + hf * 8192), bo = desc_adv(bq, 16384);
        const uint32_t acc0 = g > 0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dV += P^T dO   (K = 64 queries)
          umma_ts(tdV, tSt + hf * 64 + k * 16, desc_adv(bo, k * 2048), idesc_kv, k > 0 ? 1u : acc0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dK += dS^T Q
          umma_ts(tdK, tdPt + hf * 64 + k * 16, desc_adv(bq, k * 2048), idesc_kv, k > 0 ? 1u : acc0);
        if (hf == 1) umma_commit(qd_empty0 + 8 * s);  // Q/dO(i) are not needed by dQ(i)
      }
      __syncwarp();
      // scores(g+2) overwrite this half's columns: behind dV(g) and dK(g)
      if (g + 2 < G) issue_scores(g + 2);
      TRACE(g * 4 + 1);
      if (hf == 1) {
        if (i > 0) {
          mbar_wait(dq_free, (i - 1) & 1);  // dQ(i-1) has been read out of TMEM
          tc_fence_after();
        }
        if (elect_one()) {
          const uint64_t da = desc_adv(ddS_mn, (i & 1) * 32768);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dQ = dS K      (K = keys; A = dS^T bytes read MN-major)
            umma_ss(tdQ, desc_adv(da, k * 2048), desc_adv(dK_mn, k * 2048), idesc_dq, k > 0 ? 1u : 0u);
          umma_commit(mma2_done);
        }
        __syncwarp();
      }
      TRACE(g * 4 + 2);
    }
    if (elect_one()) umma_commit(all_done);
    __syncwarp();
  } else if (warp >= 2 && warp < 18) {
    // 16 compute warps = 4 per TMEM lane quadrant (four warps per SM sub-partition hide each other's
    // MUFU / LDS / TMEM latencies); warp `part` of a quadrant owns 16 of the 64 query columns of each
    // half of S^T / dP^T and 16 of the 64 columns of the dK / dV accumulators.
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;   // 0..3
    const int row = quad * 32 + lane;   // key row of S^T / dP^T
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const int key = kt * 128 + row;
    const bool key_ok = key < p.Nk;
    // -inf bias => P = 0 for padded keys, without a select in the inner loop
    const float kbias = !key_ok ? -INFINITY
                                : (p.key_bias ? p.key_bias[(int64_t)b * p.Nk + key] * kLog2eB : 0.f);
    const bool has_bias = __any_sync(0xffffffffu, kbias != 0.f);  // warp-uniform fast-path switch
    {
      // K, V rows -> TMEM (A operands of the score MMAs): this warp copies 16 of the 64 head columns,
      // i.e. two 16-byte chunks of the swizzled 128-byte row = 8 packed bf16x2 TMEM columns
      mbar_wait(kv_full, 0);
      uint32_t rk[8], rv[8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t off = sw128_off(row, part * 2 + c);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rk[c * 4]), "=r"(rk[c * 4 + 1]), "=r"(rk[c * 4 + 2]), "=r"(rk[c * 4 + 3])
                     : "r"(sK + off));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(rv[c * 4]), "=r"(rv[c * 4 + 1]), "=r"(rv[c * 4 + 2]), "=r"(rv[c * 4 + 3])
                     : "r"(sV + off));
      }
      tmem_st8(tK + lane_bits + part * 8, rk);
      tmem_st8(tV + lane_bits + part * 8, rv);
      tmem_st_wait();
      tc_fence_before();
      CW_ARRIVE(kv_tmem);
    }
    for (int i = 0; i < T; ++i) {
      const int s = i & 1;
      const float* st = stat + s * 256;
      mbar_wait(stat_full0 + 8 * s, (i >> 1) & 1);  // lse / delta of tile i are in shared memory
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const uint32_t sdSt = sDS + s * 32768 + hf * 16384;
        const int c0 = hf * 64 + part * 16;  // this warp's 16 query columns inside the tile
        if (warp == 2) TRACE(2048 + (2 * i + hf) * 4 + 0);
        mbar_wait(s_full0 + 8 * hf, i & 1);
        tc_fence_after();
        if (warp == 2) TRACE(2048 + (2 * i + hf) * 4 + 1);
        uint32_t rs[16], rd[16];
        tmem_ld16(tSt + lane_bits + c0, rs);
        tmem_ld16(tdPt + lane_bits + c0, rd);
        tmem_ld_wait();
        if (warp == 2) TRACE(2048 + (2 * i + hf) * 4 + 2);
        const float4* lse4 = reinterpret_cast<const float4*>(st + c0);
        const float4* del4 = reinterpret_cast<const float4*>(st + 128 + c0);
        uint32_t pk[8], dk8[8];  // 16 query columns of P^T / dS^T as packed bf16x2 -> 8 TMEM columns each
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float pv[8], ds[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const float4 ls = lse4[g * 2 + h4], dl = del4[g * 2 + h4];  // broadcast LDS.128
            const float lsv[4] = {ls.x, ls.y, ls.z, ls.w}, dlv[4] = {dl.x, dl.y, dl.z, dl.w};
#if FAB_F32X2
            // packed fp32x2: the two FMAs and the product of a PAIR of scores take three instructions instead of six
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int j = g * 8 + h4 * 4 + e;
              // (the staged statistics are NEGATED -- -lse * log2e, -delta * scale -- so they enter as plain addends)
              float2 nl = make_float2(lsv[e], lsv[e + 1]);
              if (has_bias) nl = __fadd2_rn(nl, make_float2(kbias, kbias));
              const float2 arg = __ffma2_rn(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])),
                                            make_float2(p.scale_log2, p.scale_log2), nl);
              const float2 pr = make_float2(ex2_approx_b(arg.x), ex2_approx_b(arg.y));
              const float2 t = __ffma2_rn(make_float2(__uint_as_float(rd[j]), __uint_as_float(rd[j + 1])),
                                          make_float2(p.scale, p.scale), make_float2(dlv[e], dlv[e + 1]));
              const float2 d2 = __fmul2_rn(pr, t);
              pv[h4 * 4 + e] = pr.x; pv[h4 * 4 + e + 1] = pr.y;
              ds[h4 * 4 + e] = d2.x; ds[h4 * 4 + e + 1] = d2.y;
            }
#else
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = g * 8 + h4 * 4 + e;
              const float arg = has_bias ? fmaf(__uint_as_float(rs[j]), p.scale_log2, kbias + lsv[e])
                                         : fmaf(__uint_as_float(rs[j]), p.scale_log2, lsv[e]);
              const float pr = ex2_approx_b(arg);
              pv[h4 * 4 + e] = pr;
              ds[h4 * 4 + e] = pr * fmaf(__uint_as_float(rd[j]), p.scale, dlv[e]);
            }
#endif
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            pk[g * 4 + e] = pack_bf16x2(pv[2 * e], pv[2 * e + 1]);
            dk8[g * 4 + e] = pack_bf16x2(ds[2 * e], ds[2 * e + 1]);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdSt + sw128_off(row, part * 2 + g)),
                       "r"(dk8[g * 4 + 0]), "r"(dk8[g * 4 + 1]), "r"(dk8[g * 4 + 2]), "r"(dk8[g * 4 + 3])
                       : "memory");
        }
        tmem_st8(tSt + lane_bits + c0, pk);    // in place: the first 8 of this warp's own 16 columns
        tmem_st8(tdPt + lane_bits + c0, dk8);
        fence_proxy_async_smem();
        tmem_st_wait();
        tc_fence_before();
        CW_ARRIVE(pds_full0 + 8 * hf);
        if (warp == 2) TRACE(2048 + (2 * i + hf) * 4 + 3);
      }
    }
    // dK, dV of this key tile: each of the 4 warps of a quadrant writes 16 of the 64 head columns
    if (T > 0) {
      mbar_wait(all_done, 0);
      tc_fence_after();
      uint32_t rv[16], rk[16];
      tmem_ld16(tdV + lane_bits + part * 16, rv);
      tmem_ld16(tdK + lane_bits + part * 16, rk);
      tmem_ld_wait();
      if (tail_slot >= 0) {
        // one of several CTAs on this key tile (sparse last wave): fp32 partials, rows past Nk included (never read)
        float* dst = p.tail_ws + ((int64_t)tail_slot * 128 + row) * 128 + part * 16;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<float4*>(dst + g * 4) = make_float4(__uint_as_float(rk[g * 4]), __uint_as_float(rk[g * 4 + 1]),
                                                                __uint_as_float(rk[g * 4 + 2]), __uint_as_float(rk[g * 4 + 3]));
          *reinterpret_cast<float4*>(dst + 64 + g * 4) = make_float4(__uint_as_float(rv[g * 4]), __uint_as_float(rv[g * 4 + 1]),
                                                                     __uint_as_float(rv[g * 4 + 2]), __uint_as_float(rv[g * 4 + 3]));
        }
      } else if (key_ok && splits == 1) {
        bf16* dkr = p.dk + ((int64_t)b * p.Nk + key) * p.lddk + h * 64 + part * 16;
        bf16* dvr = p.dv + ((int64_t)b * p.Nk + key) * p.lddv + h * 64 + part * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(rv[g * 8 + 0]), __uint_as_float(rv[g * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(rv[g * 8 + 2]), __uint_as_float(rv[g * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(rv[g * 8 + 4]), __uint_as_float(rv[g * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(rv[g * 8 + 6]), __uint_as_float(rv[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dvr + g * 8) = u;
          u.x = pack_bf16x2(__uint_as_float(rk[g * 8 + 0]), __uint_as_float(rk[g * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(rk[g * 8 + 2]), __uint_as_float(rk[g * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(rk[g * 8 + 4]), __uint_as_float(rk[g * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(rk[g * 8 + 6]), __uint_as_float(rk[g * 8 + 7]));
          *reinterpret_cast<uint4*>(dkr + g * 8) = u;
        }
      } else if (key_ok) {
        const int64_t prow = ((int64_t)split * p.B + b) * p.Nk + key;
        float4* pk4 = reinterpret_cast<float4*>(p.part_dk + prow * (p.H * 64) + h * 64 + part * 16);
        float4* pv4 = reinterpret_cast<float4*>(p.part_dv + prow * (p.H * 64) + h * 64 + part * 16);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          pk4[g] = make_float4(__uint_as_float(rk[g * 4]), __uint_as_float(rk[g * 4 + 1]),
                               __uint_as_float(rk[g * 4 + 2]), __uint_as_float(rk[g * 4 + 3]));
          pv4[g] = make_float4(__uint_as_float(rv[g * 4]), __uint_as_float(rv[g * 4 + 1]),
                               __uint_as_float(rv[g * 4 + 2]), __uint_as_float(rv[g * 4 + 3]));
        }
      }
    }
  } else if (warp >= 18) {
    // 4 drain / stat warps, one per TMEM lane quadrant (warp % 4).
    const int quad = warp & 3;
    const int dt = (warp - 18) * 32 + lane;  // 0..127: the query (within a tile) whose lse/delta this thread loads
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const float* lse_g = p.lse + ((int64_t)b * p.H + h) * p.Nq;
    const float* del_g = p.delta + ((int64_t)b * p.H + h) * p.Nq;
    auto load_stats = [&](int i, float& l, float& d) {  // -lse in log2 units, -delta pre-scaled; padded queries: P = 0
      const int q = (t0 + (i + kt) % T) * 128 + dt;
      const bool ok = i < T && q < p.Nq;
      l = ok ? -lse_g[q] * kLog2eB : -INFINITY;   // negated: the compute warps add them
      d = ok ? -del_g[q] * p.scale : 0.f;
    };
    float l0, d0, l1, d1;
    load_stats(0, l0, d0);
    load_stats(1, l1, d1);
    stat[dt] = l0; stat[128 + dt] = d0;
    DW_ARRIVE(stat_full0);
    if (T > 1) {
      stat[256 + dt] = l1; stat[384 + dt] = d1;
      DW_ARRIVE(stat_full0 + 8);
    }
    const uint32_t stage = sStage + (warp - 18) * 8192;
    const int row0 = quad * 32;  // first query row (within the tile) of this warp's dQ slab
    for (int i = 0; i < T; ++i) {
      float ln, dn;
      load_stats(i + 2, ln, dn);  // global-load latency hidden behind the wait below
      mbar_wait(mma2_done, i & 1);  // dQ(i) complete; every MMA and compute warp is past tile i
      tc_fence_after();
      if (warp == 18) TRACE(4096 + i * 4 + 0);
      uint32_t r0[32], r1[32];
      tmem_ld32(tdQ + lane_bits, r0);
      tmem_ld32(tdQ + lane_bits + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      DW_ARRIVE(dq_free);  // the dQ columns may be overwritten by dQ(i+1)
      if (warp == 18) TRACE(4096 + i * 4 + 1);
      if (i + 2 < T) {  // stat buffer i&1 was last read for tile i
        stat[(i & 1) * 256 + dt] = ln;
        stat[(i & 1) * 256 + 128 + dt] = dn;
        DW_ARRIVE(stat_full0 + 8 * (i & 1));
      }
      if (lane == 0) tma_store_wait_read<0>();  // this warp's previous reduce-add has read its staging slab
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + sw128_off(lane, g)),
                     "r"(r0[g * 4 + 0]), "r"(r0[g * 4 + 1]), "r"(r0[g * 4 + 2]), "r"(r0[g * 4 + 3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + 4096 + sw128_off(lane, g)),
                     "r"(r1[g * 4 + 0]), "r"(r1[g * 4 + 1]), "r"(r1[g * 4 + 2]), "r"(r1[g * 4 + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int qt = t0 + (i + kt) % T;
        tma_reduce_add_3d(&tmdQ, stage, h * 64, qt * 128 + row0, b);
        tma_reduce_add_3d(&tmdQ, stage + 4096, h * 64 + 32, qt * 128 + row0, b);
        tma_store_commit();
      }
      if (warp == 18) TRACE(4096 + i * 4 + 2);
    }
    if (lane == 0) tma_store_wait_read<0>();  // staging read; the reduce-adds themselves complete with the kernel
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dk/dv (bf16) = sum over the q_splits fp32 partials; 8 columns per thread.
__global__ void __launch_bounds__(256) fa_bwd_reduce_kernel(const float* __restrict__ part_dk,
                                                            const float* __restrict__ part_dv, bf16* dk,
                                                            int64_t lddk, bf16* dv, int64_t lddv, int splits,
                                                            int64_t rows, int D) {
  pdl_launch();
  pdl_wait();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = D / 8;
  if (gid >= rows * per_row) return;
  const int64_t r = gid / per_row;
  const int c = (int)(gid % per_row) * 8;
  float ak[8] = {0, 0, 0, 0, 0, 0, 0, 0}, av[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < splits; ++s) {
    const float4* a = reinterpret_cast<const float4*>(part_dk + ((int64_t)s * rows + r) * D + c);
    const float4* v = reinterpret_cast<const float4*>(part_dv + ((int64_t)s * rows + r) * D + c);
    const float4 a0 = a[0], a1 = a[1], v0 = v[0], v1 = v[1];
    ak[0] += a0.x; ak[1] += a0.y; ak[2] += a0.z; ak[3] += a0.w; ak[4] += a1.x; ak[5] += a1.y; ak[6] += a1.z; ak[7] += a1.w;
    av[0] += v0.x; av[1] += v0.y; av[2] += v0.z; av[3] += v0.w; av[4] += v1.x; av[5] += v1.y; av[6] += v1.z; av[7] += v1.w;
  }
  uint4 u;
  u.x = pack_bf16x2(ak[0], ak[1]); u.y = pack_bf16x2(ak[2], ak[3]); u.z = pack_bf16x2(ak[4], ak[5]); u.w = pack_bf16x2(ak[6], ak[7]);
  *reinterpret_cast<uint4*>(dk + r * lddk + c) = u;
  u.x = pack_bf16x2(av[0], av[1]); u.y = pack_bf16x2(av[2], av[3]); u.z = pack_bf16x2(av[4], av[5]); u.w = pack_bf16x2(av[6], av[7]);
  *reinterpret_cast<uint4*>(dv + r * lddv + c) = u;
}

// dk/dv (bf16) of the split items of the last wave = sum over their `tail_parts` fp32 partials; 8 columns per thread.
__global__ void __launch_bounds__(256) fa_bwd_tail_reduce_kernel(const FaBwdParams p, int n_split) {
  pdl_launch();
  pdl_wait();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (split item, row, 16 groups of 8 columns)
  if (gid >= (int64_t)n_split * 128 * 16) return;
  const int c8 = (int)(gid & 15) * 8, row = (int)((gid >> 4) & 127), si = (int)(gid >> 11);
  const int item = p.n_whole + si;
  const int kt = item % p.k_tiles, h = (item / p.k_tiles) % p.H, b = item / (p.k_tiles * p.H);
  const int key = kt * 128 + row;
  if (key >= p.Nk) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = 0; t < p.tail_parts; ++t) {
    const float* src = p.tail_ws + (((int64_t)si * p.tail_parts + t) * 128 + row) * 128 + c8;
    const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
    acc[0] += a0.x; acc[1] += a0.y; acc[2] += a0.z; acc[3] += a0.w; acc[4] += a1.x; acc[5] += a1.y; acc[6] += a1.z; acc[7] += a1.w;
  }
  uint4 u;
  u.x = pack_bf16x2(acc[0], acc[1]); u.y = pack_bf16x2(acc[2], acc[3]); u.z = pack_bf16x2(acc[4], acc[5]); u.w = pack_bf16x2(acc[6], acc[7]);
  bf16* dst = c8 < 64 ? p.dk + ((int64_t)b * p.Nk + key) * p.lddk + h * 64 + c8
                      : p.dv + ((int64_t)b * p.Nk + key) * p.lddv + h * 64 + (c8 - 64);
  *reinterpret_cast<uint4*>(dst) = u;
}

// attn1: one CTA per SM; 1536 items on 148 SMs are 10.4 waves, i.e. 56 CTAs alone on the machine for a whole CTA
// lifetime.  The items of such a sparse last wave are split over `parts` CTAs along the query walk.
static void fa_bwd_tail_plan(int items, int q_tiles, int* n_whole, int* parts) {
  static int slots = 0;
  if (!slots) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&slots, cudaDevAttrMultiProcessorCount, dev);
    if (slots <= 0) slots = 148;
  }
  *n_whole = items;
  *parts = 1;
  static const bool off = [] { const char* e = getenv("B200_FA_BWD_NO_TAIL"); return e && e[0] == '1'; }();
  if (off) return;
  const int r = items % slots;
  if (items < slots || r == 0 || r * 4 > slots * 3) return;
  // r items as r * pr CTAs of 1/pr of the walk each: rounds of `slots` CTAs, each round 1/pr of a full CTA lifetime,
  // plus what a part costs (K / V load, fp32 partials and their reduction); un-split = 1.0
  int max_pr = q_tiles / 8 < 4 ? q_tiles / 8 : 4;             // at least 8 query tiles per part
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

// How many CTAs share one key tile's query walk: enough to cover the SMs when there are few key tiles.
static int fa_bwd_splits(int B, int H, int Nq, int Nk, bool masked) {
  const int64_t ctas = (int64_t)((Nk + 127) / 128) * H * B;
  const int T = (Nq + 127) / 128;
  if (ctas <= 0 || ctas >= 120 || T < 8) return 1;
  // One wave of CTAs; with a key mask up to two waves' worth: key tiles that turn out to be fully masked leave at
  // once (attn2 with the real prompt: half of them), and with nothing masked two waves of half-length walks cost
  // about the same as one wave of full ones.
  int s = (int)((masked ? 296 : 148) / ctas);
  return s > T / 4 ? (T / 4 > 0 ? T / 4 : 1) : s;
}

}  // namespace b200

using namespace b200;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#ifdef B200_TRACE
extern "C" int b200_debug_bwd_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_bwd_trace, sizeof(unsigned long long) * n);
}
#endif

extern "C" int64_t b200_fa_bwd_workspace_bytes(int B, int H, int Nq, int Nk) {
  if (B <= 0 || H <= 0 || Nq <= 0 || Nk <= 0) return 0;
  const int s = fa_bwd_splits(B, H, Nq, Nk, true);  // upper bound over both split choices
  if (s > 1) return (int64_t)2 * s * B * Nk * H * 64 * (int64_t)sizeof(float);
  int n_whole, parts;
  const int items = ((Nk + 127) / 128) * H * B;
  fa_bwd_tail_plan(items, (Nq + 127) / 128, &n_whole, &parts);
  return (int64_t)(items - n_whole) * parts * 128 * 128 * (int64_t)sizeof(float);
}

extern "C" int b200_fa_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, const void* dout, int64_t lddo, const float* lse,
                           const float* delta, const float* key_bias, float* dq_accum, int64_t lddq,
                           void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Nq, int Nk,
                           int head_dim, float scale, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  if (head_dim != 64) return arg_error("fa_bwd: only head_dim 64 is built (LTXV-2B: 32 heads x 64)");
  if (B < 0 || H <= 0 || Nq < 0 || Nk < 0) return arg_error("fa_bwd: bad shape");
  if (B == 0 || Nk == 0) return 0;
  if (!(q && k && v && dout && lse && delta && dq_accum && dk && dv)) return arg_error("fa_bwd: null pointer");
  if (ldq % 8 || ldk % 8 || ldv % 8 || lddo % 8 || lddk % 8 || lddv % 8 || lddq % 4 || !al16(q) ||
      !al16(k) || !al16(v) || !al16(dout) || !al16(dq_accum) || !al16(dk) || !al16(dv))
    return arg_error("fa_bwd: tensors must be 16-byte aligned with 16-byte-multiple pitches");
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc;
  if ((rc = make_tmap_tokens(&tmQ, q, B, Nq, ldq, H * 64, 128)) ||
      (rc = make_tmap_tokens(&tmK, k, B, Nk, ldk, H * 64, 128)) ||
      (rc = make_tmap_tokens(&tmV, v, B, Nk, ldv, H * 64, 128)) ||
      (rc = make_tmap_tokens(&tmdO, dout, B, Nq, lddo, H * 64, 128)))
    return arg_error("fa_bwd: cuTensorMapEncodeTiled failed", rc);
  CUtensorMap tmdQ;
  {
    uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)Nq, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)lddq * 4, (uint64_t)Nq * (uint64_t)lddq * 4};
    uint32_t box[3] = {32, 32, 1};
    if ((rc = make_tmap(&tmdQ, dq_accum, TM_F32, 3, dims, str, box, true)))
      return arg_error("fa_bwd: cuTensorMapEncodeTiled failed (dq)", rc);
  }
  FaBwdParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.q_tiles = (Nq + 127) / 128;
  p.lse = lse; p.delta = delta; p.key_bias = key_bias;
  p.dq = dq_accum; p.lddq = lddq;
  p.dk = (bf16*)dk; p.lddk = lddk; p.dv = (bf16*)dv; p.lddv = lddv;
  p.scale = scale; p.scale_log2 = scale * kLog2eB;
  p.q_splits = fa_bwd_splits(B, H, Nq, Nk, key_bias != nullptr);
  p.part_dk = p.part_dv = nullptr;
  if (p.q_splits > 1) {
    const int64_t need = (int64_t)2 * p.q_splits * B * Nk * H * 64 * (int64_t)sizeof(float);
    if (!workspace || workspace_bytes < need || !al16(workspace))
      return arg_error("fa_bwd: workspace too small (see b200_fa_bwd_workspace_bytes)");
    p.part_dk = (float*)workspace;
    p.part_dv = p.part_dk + (int64_t)p.q_splits * B * Nk * H * 64;
  }
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_BWD_SMEM) != cudaSuccess)
      return launch_status("fa_bwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  p.k_tiles = 0; p.n_whole = 0; p.tail_parts = 1; p.tail_ws = nullptr;
  dim3 grid((Nk + 127) / 128, H, B * p.q_splits);
  int n_split = 0;
  if (p.q_splits == 1) {   // 1-D grid, the items of a sparse last wave split along the query walk (needs the workspace)
    const int items = (int)(grid.x * grid.y * grid.z);
    p.k_tiles = (int)grid.x;
    p.n_whole = items;
    if (workspace != nullptr && al16(workspace)) {
      fa_bwd_tail_plan(items, p.q_tiles, &p.n_whole, &p.tail_parts);
      n_split = items - p.n_whole;
      if ((int64_t)n_split * p.tail_parts * 128 * 128 * (int64_t)sizeof(float) > workspace_bytes) {
        p.n_whole = items; p.tail_parts = 1; n_split = 0;
      }
      p.tail_ws = (float*)workspace;
    }
    grid = dim3((unsigned)(p.n_whole + n_split * p.tail_parts), 1, 1);
  }
  B200_LAUNCH(fa_bwd_kernel, grid, FA_BWD_THREADS, FA_BWD_SMEM, stream, tmQ, tmK, tmV, tmdO, tmdQ, p);
  if (n_split > 0)
    B200_LAUNCH(fa_bwd_tail_reduce_kernel, (unsigned)(((int64_t)n_split * 128 * 16 + 255) / 256), 256, 0, stream, p, n_split);
  if (p.q_splits > 1) {
    const int64_t rows = (int64_t)B * Nk, threads = rows * (H * 64 / 8);
    B200_LAUNCH(fa_bwd_reduce_kernel, (unsigned)((threads + 255) / 256), 256, 0, stream, p.part_dk, p.part_dv, (bf16*)dk,
                lddk, (bf16*)dv, lddv, p.q_splits, rows, H * 64);
  }
  return launch_status("fa_bwd");
}
