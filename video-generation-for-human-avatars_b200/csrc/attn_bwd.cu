// attn_bwd.cu — tcgen05 / TMEM flash-attention backward, head_dim 64, non-causal (sm_100a).
//
// Gradient of softmax(q k^T * scale + key_bias) v with respect to q, k, v — what autograd runs
// through F.scaled_dot_product_attention in the reference's LoRA train step
// (training.py:203 -> attention.py:1057).  Same token-major [B*N, ld] layout as attn_fwd.cu.
//
// One CTA owns one 128-key tile of one head and walks over the 128-query tiles.  Everything is
// computed transposed (keys on TMEM lanes) so that each compute thread owns one key row:
//   S^T  = K Q^T              (128 x 128 x 64, both operands K-major)
//   dP^T = V dO^T             (128 x 128 x 64)
//   P^T  = exp2(S^T*c + bias - lse) ; dS^T = P^T * (dP^T - delta) * scale    -> shared memory, bf16
//   dV  += P^T  dO            (A = P^T  packed bf16 in TMEM,  B = dO tile read MN-major)
//   dK  += dS^T Q             (A = dS^T K-major,            B = Q  tile read MN-major)
//   dQ   = dS   K             (A = the same dS^T bytes read MN-major, B = K tile read MN-major)
// dV/dK accumulate in TMEM over the whole loop; dQ is a per-(q tile, k tile) partial that is
// reduced across key-tile CTAs with TMA reduce-add (cp.reduce.async.bulk.tensor) into a caller-zeroed
// fp32 buffer (per-lane red.global.add measured ~10 k cycles per tile: the atomics were the bottleneck).
// TMEM: S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ 64 | P^T (bf16x2) 64 = 512 columns.
//
// Pipeline: P^T/dS^T are double-buffered in shared memory, so while the 8 compute warps turn
// S^T/dP^T(i+1) into P^T/dS^T(i+1) the tensor pipe runs dV/dK/dQ(i); dQ(i-1) is drained to global
// memory right after P^T/dS^T(i) are handed over.
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
};

constexpr int FA_BWD_SMEM = 16384 * 2 /*K,V*/ + 2 * 32768 /*Q,dO x2*/ + 2 * 32768 /*dS^T x2*/ +
                            32768 /*dQ staging*/ + 2 * 2 * 512 /*lse, delta x2*/ + 128;
constexpr float kLog2eB = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

constexpr int FA_BWD_THREADS = 64 + 512;  // TMA warp, MMA warp, 16 compute warps

__global__ void __launch_bounds__(FA_BWD_THREADS, 1)
fa_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
              const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ FaBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sK = sbase, sV = sbase + 16384, sQdO = sbase + 32768;  // stage s: Q at +s*32768, dO at +16384
  // dS^T is double-buffered in shared memory (buffer = tile & 1); P^T lives in TMEM (A operand of dV).
  const uint32_t sDS = sQdO + 65536;
  const uint32_t sStage = sDS + 65536;   // fp32 dQ staging tile for the TMA reduce-add
  const uint32_t sStat = sStage + 32768;  // [2 stages][lse 128 | delta 128] fp32
  const uint32_t bar = sStat + 2048;
  const uint32_t kv_full = bar, qd_full0 = bar + 8, qd_empty0 = bar + 24, s_full = bar + 40,
                 pds_full = bar + 48, mma2_done = bar + 56, dq_free = bar + 64, tmem_slot = bar + 72;
  float* stat = reinterpret_cast<float*>(smem_raw + (sStat - sbase));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.q_tiles;

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
    for (int s = 0; s < 2; ++s) {
      mbar_init(qd_full0 + 8 * s, 1);
      mbar_init(qd_empty0 + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pds_full, 512);
    mbar_init(mma2_done, 1);
    mbar_init(dq_free, 512);
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
                 tdK = tmem_base + 320, tdQ = tmem_base + 384, tPt = tmem_base + 448;

  if (warp == 0 && lane == 0) {
    mbar_expect_tx(kv_full, 32768);
    tma_load_3d(sK, &tmK, kv_full, h * 64, kt * 128, b);
    tma_load_3d(sV, &tmV, kv_full, h * 64, kt * 128, b);
    for (int i = 0; i < T; ++i) {
      const int s = i & 1;
      mbar_wait(qd_empty0 + 8 * s, ((i >> 1) & 1) ^ 1);
      mbar_expect_tx(qd_full0 + 8 * s, 32768);
      tma_load_3d(sQdO + s * 32768, &tmQ, qd_full0 + 8 * s, h * 64, i * 128, b);
      tma_load_3d(sQdO + s * 32768 + 16384, &tmdO, qd_full0 + 8 * s, h * 64, i * 128, b);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // K-major x K-major
    const uint32_t idesc_kv = make_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major
    const uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);   // A MN-major, B MN-major
    mbar_wait(kv_full, 0);
    // S^T / dP^T of tile i+1 are issued right behind dV/dK/dQ of tile i, so the tensor pipe keeps
    // running while the compute warps drain dQ_i and the next S^T is ready when they come back.
    auto issue_scores = [&](int i) {
      const int s = i & 1;
      const uint32_t sQ = sQdO + s * 32768, sdO = sQ + 16384;
      mbar_wait(qd_full0 + 8 * s, (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss(tSt, make_smem_desc(sK + k * 32, 16, 1024), make_smem_desc(sQ + k * 32, 16, 1024),
                idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss(tdPt, make_smem_desc(sV + k * 32, 16, 1024), make_smem_desc(sdO + k * 32, 16, 1024),
                idesc_s, k > 0 ? 1u : 0u);
      umma_commit(s_full);
    };
    if (T > 0) issue_scores(0);
    for (int i = 0; i < T; ++i) {
      const int s = i & 1;
      const uint32_t sQ = sQdO + s * 32768, sdO = sQ + 16384;
      const uint32_t sdSt = sDS + s * 32768;
      mbar_wait(pds_full, i & 1);  // P^T(i) is in TMEM, dS^T(i) in shared memory; S^T/dP^T are consumed
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dV += P^T dO   (A = packed bf16 P^T straight from TMEM, K = queries)
        umma_ts(tdV, tPt + k * 8, make_smem_desc(sdO + k * 2048, 8192, 1024), idesc_kv,
                (i > 0 || k > 0) ? 1u : 0u);
      // S^T/dP^T(i+1) go behind dV(i): once they complete, P^T(i) has been consumed and the compute
      // warps may overwrite it while dK/dQ(i) run below
      if (i + 1 < T) issue_scores(i + 1);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dK += dS^T Q
        umma_ss(tdK, make_smem_desc(sdSt + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                make_smem_desc(sQ + k * 2048, 8192, 1024), idesc_kv, (i > 0 || k > 0) ? 1u : 0u);
      if (i > 0) {
        mbar_wait(dq_free, (i - 1) & 1);  // dQ(i-1) has been read out of TMEM
        tc_fence_after();
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dQ = dS K      (K = keys; A = dS^T bytes read MN-major)
        umma_ss(tdQ, make_smem_desc(sdSt + k * 2048, 16384, 1024),
                make_smem_desc(sK + k * 2048, 8192, 1024), idesc_dq, k > 0 ? 1u : 0u);
      umma_commit(qd_empty0 + 8 * s);
      umma_commit(mma2_done);
    }
  } else if (warp >= 2) {
    // 16 compute warps = 4 per TMEM lane quadrant (four warps per SM sub-partition hide each other's
    // MUFU / LDS / TMEM latencies); warp `part` of a quadrant owns one 32-column chunk of the 128 query
    // columns of S^T / dP^T and 16 of the 64 columns of the dQ / dK / dV accumulators.
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;   // 0..3
    const int row = quad * 32 + lane;   // key row of S^T / dP^T, query row of dQ
    const int ctid = threadIdx.x - 64;  // 0..511 among the compute threads
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const int key = kt * 128 + row;
    const bool key_ok = key < p.Nk;
    // -inf bias => P = 0 for padded keys, without a select in the inner loop
    const float kbias = !key_ok ? -INFINITY
                                : (p.key_bias ? p.key_bias[(int64_t)b * p.Nk + key] * kLog2eB : 0.f);
    const bool has_bias = __any_sync(0xffffffffu, kbias != 0.f);  // warp-uniform fast-path switch
    const float* stat_g = (ctid < 128 ? p.lse : p.delta) + ((int64_t)b * p.H + h) * p.Nq;
    const float stat_mul = ctid < 128 ? kLog2eB : p.scale;   // lse -> log2 units, delta -> pre-scaled
    const float stat_pad = ctid < 128 ? INFINITY : 0.f;      // +inf lse => P = 0 for padded queries
    const int sq = ctid & 127;
    const bool stat_thread = ctid < 256;
    // lse / delta of tile i+1 are fetched one tile ahead (global-load latency off the critical path)
    float nxt = (stat_thread && sq < p.Nq) ? stat_g[sq] * stat_mul : stat_pad;
    // dQ(j): TMEM -> swizzled fp32 staging tile (the dead P^T buffer of tile j) -> TMA reduce-add
    // (cp.reduce.async.bulk.tensor .add) into the fp32 dQ accumulator in global memory.
    auto drain_dq = [&](int j) {
      mbar_wait(mma2_done, j & 1);
      tc_fence_after();
      uint32_t r0[16];
      tmem_ld16(tdQ + lane_bits + part * 16, r0);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);  // the dQ columns may be overwritten by dQ(j+1)
      const uint32_t stage = sStage;
      if (ctid == 0) tma_store_wait_read<0>();  // the previous reduce-add has finished reading the staging tile
      named_bar_sync(2, 512);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (part >> 1) * 16384 +
                                                                      sw128_off(row, (part & 1) * 4 + g)),
                     "r"(r0[g * 4 + 0]), "r"(r0[g * 4 + 1]), "r"(r0[g * 4 + 2]), "r"(r0[g * 4 + 3])
                     : "memory");
      fence_proxy_async_smem();
      named_bar_sync(3, 512);
      if (ctid == 0) {
        tma_reduce_add_3d(&tmdQ, stage, h * 64, j * 128, b);
        tma_reduce_add_3d(&tmdQ, stage + 16384, h * 64 + 32, j * 128, b);
        tma_store_commit();
      }
    };
    for (int i = 0; i < T; ++i) {
      const int s = i & 1;
      const int q0 = i * 128;
      float* st = stat + s * 256;
      const uint32_t sdSt = sDS + s * 32768;
      if (stat_thread) {
        st[ctid] = nxt;  // [0,128): lse * log2e ; [128,256): delta * scale
        const int qn = q0 + 128 + sq;
        nxt = qn < p.Nq ? stat_g[qn] * stat_mul : stat_pad;
      }
      // dS^T buffer s was last read by dK/dQ(i-2); mma2_done(i-2) was awaited when dQ(i-2) was drained
      named_bar_sync(1, 512);
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      {
        const int c = part;  // this warp's 32-column chunk of the 128 query columns
        uint32_t rs[32], rd[32];
        tmem_ld32(tSt + lane_bits + c * 32, rs);
        tmem_ld32(tdPt + lane_bits + c * 32, rd);
        tmem_ld_wait();
        const uint32_t off0 = (c >> 1) * 16384;
        const float4* lse4 = reinterpret_cast<const float4*>(st + c * 32);
        const float4* del4 = reinterpret_cast<const float4*>(st + 128 + c * 32);
        uint32_t pk[16];  // this warp's 32 query columns of P^T as packed bf16x2 -> 16 TMEM columns
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pv[8], ds[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const float4 ls = lse4[g * 2 + h4], dl = del4[g * 2 + h4];  // broadcast LDS.128
            const float lsv[4] = {ls.x, ls.y, ls.z, ls.w}, dlv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = g * 8 + h4 * 4 + e;
              const float arg = has_bias ? fmaf(__uint_as_float(rs[j]), p.scale_log2, kbias - lsv[e])
                                         : fmaf(__uint_as_float(rs[j]), p.scale_log2, -lsv[e]);
              const float pr = ex2_approx_b(arg);
              pv[h4 * 4 + e] = pr;
              ds[h4 * 4 + e] = pr * fmaf(__uint_as_float(rd[j]), p.scale, -dlv[e]);
            }
          }
          const uint32_t o = off0 + sw128_off(row, (c & 1) * 4 + g);
          pk[g * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
          pk[g * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
          pk[g * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
          pk[g * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdSt + o),
                       "r"(pack_bf16x2(ds[0], ds[1])), "r"(pack_bf16x2(ds[2], ds[3])),
                       "r"(pack_bf16x2(ds[4], ds[5])), "r"(pack_bf16x2(ds[6], ds[7]))
                       : "memory");
        }
        tmem_st16(tPt + lane_bits + c * 16, pk);
      }
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(pds_full);
      if (i > 0) drain_dq(i - 1);  // overlaps with S^T/dP^T(i+1) and dV/dK/dQ(i) on the tensor pipe
    }
    if (T > 0) drain_dq(T - 1);
    if (ctid == 0) tma_store_wait_all<0>();
    // dK, dV of this key tile: each of the 4 warps of a quadrant writes 16 of the 64 head columns
    if (T > 0) {
      bf16* dkr = p.dk + ((int64_t)b * p.Nk + key) * p.lddk + h * 64 + part * 16;
      bf16* dvr = p.dv + ((int64_t)b * p.Nk + key) * p.lddv + h * 64 + part * 16;
      uint32_t rv[16], rk[16];
      tmem_ld16(tdV + lane_bits + part * 16, rv);
      tmem_ld16(tdK + lane_bits + part * 16, rk);
      tmem_ld_wait();
      if (key_ok) {
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
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200

using namespace b200;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int b200_fa_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                           int64_t ldv, const void* dout, int64_t lddo, const float* lse,
                           const float* delta, const float* key_bias, float* dq_accum, int64_t lddq,
                           void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Nq, int Nk,
                           int head_dim, float scale, void* stream) {
  if (!(q && k && v && dout && lse && delta && dq_accum && dk && dv)) return arg_error("fa_bwd: null pointer");
  if (head_dim != 64) return arg_error("fa_bwd: only head_dim 64 is built (LTXV-2B: 32 heads x 64)");
  if (B < 0 || H <= 0 || Nq < 0 || Nk < 0) return arg_error("fa_bwd: bad shape");
  if (B == 0 || Nk == 0) return 0;
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
    uint32_t box[3] = {32, 128, 1};
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
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_BWD_SMEM) != cudaSuccess)
      return launch_status("fa_bwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  dim3 grid((Nk + 127) / 128, H, B);
  fa_bwd_kernel<<<grid, FA_BWD_THREADS, FA_BWD_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, tmdO, tmdQ, p);
  return launch_status("fa_bwd");
}
