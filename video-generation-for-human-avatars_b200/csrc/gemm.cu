// gemm.cu — tcgen05 / TMEM / TMA bf16 GEMM with fused epilogues (sm_100a).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T  (+ A2[M,K2] * B2[N,K2]^T) )
//
// Replaces, on the LTXV block path, every nn.Linear the reference runs through cuBLASLt plus the
// element-wise kernels around it (reference: attention.py:996-1014, 1089, 1238-1263;
// transformer3d.py:470, 494-499, 561; peft lora.Linear forward, training.py:61-68):
//   * second operand pair  = the LoRA up-projection (s*x*A^T)*B^T accumulated into the same TMEM tile
//   * bias, GELU-tanh (+ pre-activation stash), GELU' (backward), per-sample AdaLN gate, residual
//   * either operand may be stored "MN-major" (reduction dim strided), which is how dgrad
//     (dY * W) and wgrad (dY^T * X) read the very same row-major tensors without a transpose.
//
// Structure: persistent CTAs (one per SM), 128 x BN output tiles, BLOCK_K = 64, warp-specialised:
//   warp 0 lane 0 : TMA producer           (cp.async.bulk.tensor, 128B swizzle, mbarrier tx)
//   warp 1 lane 0 : tcgen05.mma issuer     (M=128, N=BN, K=16 per instruction, fp32 accum in TMEM)
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue (tcgen05.ld 32x32b -> registers -> fused math -> 16-byte global stores);
//                   two warps per TMEM lane quadrant, each draining half of the tile's columns
// Accumulators are double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <cstdlib>

#include "api_internal.h"
#include <type_traits>

#include "common.cuh"
#include "tmap.h"

namespace b200 {

enum { EPI_NONE = 0, EPI_GELU = 1, EPI_GELU_GRAD = 2, EPI_STASH = 3 };  // STASH: aux <- bf16(acc + bias), before gate / residual

struct GemmParams {
  int M, N;
  int m_tiles, n_tiles;
  int group_m;   // tile rasterisation: groups of `group_m` tile rows x all tile columns are walked one after the other
  int kb1, kb2;  // 64-wide k blocks taken from (A,B) and from (A2,B2)
  int splits;    // split-K factor: work item = (tile, split); > 1 only for plain fp32 outputs (atomic accumulate)
  int epi, out_f32;
  void* C;
  int64_t ldc;
  const bf16* bias;
  const bf16* gate;
  int64_t gate_stride, rows_per_gate;
  const bf16* res;
  int64_t ldres;
  bf16* aux;
  int64_t ldaux;
  int tma_store;  // bf16 outputs (C, and the GELU pre-activation) leave through smem staging + TMA stores
  // strided batch: work item (tile, g) reads / writes sub-blocks displaced by g * (rows, cols) inside the
  // operand tensors (groups == 1: plain GEMM).  Offsets are in elements of the tensor as stored.
  int groups;
  int a_gr, a_gc, b_gr, b_gc, a2_gr, a2_gc, b2_gr, b2_gc, c_gr, c_gc, bias_g;
  // stream-K for the last, partial wave (sk_tiles > 0): the k blocks of tiles [0, sk_tiles) are dealt out evenly to
  // all CTAs (pairs) ahead of their whole tiles; a CTA whose share starts inside a tile dumps its fp32 partial into
  // `sk_ws` and raises a flag, the CTA that holds the tile's first k block adds the partials in its epilogue.
  int sk_tiles;
  float* sk_ws;   // [CTA or pair][rank][BN / 4][128][4] fp32 partial accumulators
  int* sk_flags;  // [CTA or pair][rank][8 epilogue warps], zero between launches (the consumer clears them)
};

// Tile id -> (tile row, tile column).  Plain column-major order made every wave of CTAs touch ~3 tile columns and ALL
// tile rows' worth of A only a few at a time: at K = 8192 the 100 MB A operand was re-streamed per tile column
// (ncu: 339 MB of DRAM reads against 134 MB algorithmic).  Grouped order: a wave covers `group_m` tile rows x every
// tile column, so an A tile row is read from DRAM once and re-used out of L2 by all its tile columns.
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& tm, int& tn) {
  const int gsize = p.group_m * p.n_tiles;
  const int gid = tile / gsize, in = tile - gid * gsize;
  const int first = gid * p.group_m;
  const int rows = min(p.m_tiles - first, p.group_m);
  tn = in / rows;
  tm = first + (in - tn * rows);
}

// One unit of work of the persistent walk: k blocks [kb_begin, kb_end) of output tile `tile` (z = split-K slice or
// batch group).  role: 0 = a whole tile, 1 = partial that is dumped for another CTA, 2 = partial that owns the tile's
// epilogue and adds the other CTAs' partials.
struct GemmUnit {
  int tile, z, kb_begin, kb_end, role;
};

struct GemmWalk {
  int w, P, tiles_mn, total, kb_all, kb_per, groups;
  int it, it_end;    // this CTA's remaining share of the stream-K region, in k blocks
  int64_t sk_iters;  // sk_tiles * kb_all
  int work;          // next whole-tile work item

  __device__ GemmWalk(const GemmParams& p, int w_, int P_) : w(w_), P(P_) {
    tiles_mn = p.m_tiles * p.n_tiles;
    total = tiles_mn * p.splits * p.groups;
    kb_all = p.kb1 + p.kb2;
    kb_per = (kb_all + p.splits - 1) / p.splits;
    groups = p.groups;
    sk_iters = (int64_t)p.sk_tiles * kb_all;
    it = (int)share_begin(w);
    it_end = (int)share_begin(w + 1);
    work = p.sk_tiles + w;
  }
  __device__ int64_t share_begin(int cta) const { return sk_iters * cta / P; }
  __device__ bool next(GemmUnit& u) {
    if (it < it_end) {
      u.tile = it / kb_all;
      u.z = 0;
      u.kb_begin = it - u.tile * kb_all;
      u.kb_end = min(kb_all, u.kb_begin + (it_end - it));
      u.role = u.kb_begin > 0 ? 1 : (u.kb_end < kb_all ? 2 : 0);
      it += u.kb_end - u.kb_begin;
      return true;
    }
    if (work >= total) return false;
    u.tile = work % tiles_mn;
    u.z = work / tiles_mn;
    const int split = groups > 1 ? 0 : u.z;
    u.kb_begin = split * kb_per;
    u.kb_end = min(kb_all, u.kb_begin + kb_per);
    u.role = 0;
    work += P;
    return true;
  }
};

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanh_fast(u));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  float x2 = x * x;
  float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
  float t = tanh_fast(u);
  float du = 0.7978845608028654f * (1.f + 3.f * 0.044715f * x2);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * du;
}

// Fused epilogue math on 8 consecutive columns of one output row: + bias -> GELU (on the bf16-rounded
// pre-activation, which is returned in `pre`) / GELU' -> * gate -> + residual.
// `ext` = this thread's 8 values of the external row operand (residual, or the stashed pre-activation for
// GELU'), loaded by the caller ahead of the TMEM reads so that its global-load latency is paid once per slab.
// `bias_u` / `gate_u` = the 8 bias / gate values of these columns, likewise loaded by the caller in one batch per 32
// columns: fetched here, one 16-byte load per 8 columns between the shared-memory stores, they serialised into ~0.3 us
// each and made the epilogue of a tile 6-8 us long (tools/gemm_trace.py) -- the fixed cost of every GEMM launch.
template <int EPI>
__device__ __forceinline__ void epi_math8(float (&v)[8], uint4& pre, const GemmParams& p, bool has_bias,
                                          const uint4& bias_u, bool has_gate, const uint4& gate_u, bool has_res,
                                          const uint4& ext) {
  // packed fp32x2 adds / multiplies (one instruction per PAIR of columns): the slab loop is issue bound
  auto pair = [](uint32_t w) { return make_float2(bf16_lo(w), bf16_hi(w)); };
  auto add2 = [&](int i, uint32_t w) {
    const float2 r = __fadd2_rn(make_float2(v[i], v[i + 1]), pair(w));
    v[i] = r.x; v[i + 1] = r.y;
  };
  auto mul2 = [&](int i, uint32_t w) {
    const float2 r = __fmul2_rn(make_float2(v[i], v[i + 1]), pair(w));
    v[i] = r.x; v[i + 1] = r.y;
  };
  if (has_bias) {
    const uint4 u = bias_u;
    add2(0, u.x); add2(2, u.y); add2(4, u.z); add2(6, u.w);
  }
  if (EPI == EPI_GELU) {
    if (p.aux) {
      pre.x = pack_bf16x2(v[0], v[1]); pre.y = pack_bf16x2(v[2], v[3]);
      pre.z = pack_bf16x2(v[4], v[5]); pre.w = pack_bf16x2(v[6], v[7]);
      v[0] = bf16_lo(pre.x); v[1] = bf16_hi(pre.x); v[2] = bf16_lo(pre.y); v[3] = bf16_hi(pre.y);
      v[4] = bf16_lo(pre.z); v[5] = bf16_hi(pre.z); v[6] = bf16_lo(pre.w); v[7] = bf16_hi(pre.w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = gelu_tanh(v[i]);
  } else if (EPI == EPI_GELU_GRAD) {
    const uint4 u = ext;
    float h[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                  bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= gelu_tanh_grad(h[i]);
  } else if (EPI == EPI_STASH) {
    // the pre-gate output leaves through the second store: a trainable AdaLN gate needs it for d(gate) = sum dy * u
    pre.x = pack_bf16x2(v[0], v[1]); pre.y = pack_bf16x2(v[2], v[3]);
    pre.z = pack_bf16x2(v[4], v[5]); pre.w = pack_bf16x2(v[6], v[7]);
    v[0] = bf16_lo(pre.x); v[1] = bf16_hi(pre.x); v[2] = bf16_lo(pre.y); v[3] = bf16_hi(pre.y);
    v[4] = bf16_lo(pre.z); v[5] = bf16_hi(pre.z); v[6] = bf16_lo(pre.w); v[7] = bf16_hi(pre.w);
  }
  if (has_gate) {
    const uint4 u = gate_u;
    mul2(0, u.x); mul2(2, u.y); mul2(4, u.z); mul2(6, u.w);
  }
  if (has_res) {
    const uint4 u = ext;
    add2(0, u.x); add2(2, u.y); add2(4, u.z); add2(6, u.w);
  }
}

// CTA2: a cluster of two CTAs computes a 256 x BN tile with cta_group::2 MMAs; each CTA stages its 128 rows
// of A and HALF of B's BN rows per k block (32 KB instead of 48 KB per 128 x 256 x 64 of math per SM).
template <int BN, bool CTA2 = false>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int BN_LOCAL = CTA2 ? BN / 2 : BN;  // B rows staged by this CTA
  static constexpr int B_BYTES = BN_LOCAL * BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = CTA2 ? 6 : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int STAGING = 8 * 4096;  // one 32-row x 128-byte swizzled slab per epilogue warp
  static constexpr int SMEM = NSTAGE * STAGE + STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
};

#ifdef B200_TRACE
// debug build only: per-CTA phase timestamps (globaltimer, ns) of the LAST gemm launch, read by tools/gemm_trace.py
__device__ unsigned long long g_gemm_trace[160 * 32];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GTRACE(slot) do { if (blockIdx.x < 160) g_gemm_trace[blockIdx.x * 32 + (slot)] = gtime(); } while (0)
#else
#define GTRACE(slot) do {} while (0)
#endif

template <int BN, bool A_MN, bool B_MN, bool CTA2>
__global__ void __launch_bounds__(384, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
            const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CTA2>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  const int rank = CTA2 ? (int)cluster_ctarank() : 0;  // position in the CTA pair; rank 0 issues the MMAs
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = sbase + NSTAGE * Cfg::STAGE;
  const uint32_t bar_base = stage_base + Cfg::STAGING;
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (NSTAGE + s) * 8; };
  auto tfull_bar = [&](int a) { return bar_base + (2 * NSTAGE + a) * 8; };
  auto tempty_bar = [&](int a) { return bar_base + (2 * NSTAGE + 2 + a) * 8; };
  const uint32_t tmem_slot = bar_base + (2 * NSTAGE + 4) * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) GTRACE(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kb2) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CTA2 ? 16 : 8);  // pair: the epilogue warps of both CTAs arrive on rank 0's barrier
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) {
      tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // everything above overlapped the previous kernel's tail; from here on its outputs are read
  pdl_launch();
  pdl_wait();
  if (threadIdx.x == 0) GTRACE(1);

  // persistent walk: a CTA (or CTA pair) first takes its share of the stream-K region (if any), then every
  // `work_step`-th whole work item (splits > 1, groups > 1 and stream-K exclude each other)
  const int work_first = CTA2 ? blockIdx.x >> 1 : blockIdx.x;
  const int work_step = CTA2 ? gridDim.x >> 1 : gridDim.x;

  if (warp == 0 && lane == 0) {
    // ------------------------------ TMA producer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    GemmWalk walk(p, work_first, work_step);
    GemmUnit u;
    while (walk.next(u)) {
      const int tile = u.tile;
      int tm, nt;
      tile_coords(p, tile, tm, nt);
      const int mt = tm * (CTA2 ? 2 : 1) + rank;
      const int g = p.groups > 1 ? u.z : 0;
      for (int kb = u.kb_begin; kb < u.kb_end; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        // pair: both CTAs' loads are credited to rank 0's barrier, armed by rank 0 for both
        if (!CTA2 || rank == 0) mbar_expect_tx(full_bar(stage), (CTA2 ? 2 : 1) * Cfg::STAGE);
        const bool second = kb >= p.kb1;
        const CUtensorMap* ma = second ? &tmA2 : &tmA;
        const CUtensorMap* mb = second ? &tmB2 : &tmB;
        const int kk = (second ? kb - p.kb1 : kb) * 64;
        // tensor (row, col) displacement of this group's sub-block
        const int ar = g * (second ? p.a2_gr : p.a_gr), ac = g * (second ? p.a2_gc : p.a_gc);
        const int br = g * (second ? p.b2_gr : p.b_gr), bc = g * (second ? p.b2_gc : p.b_gc);
        const uint32_t sa = sbase + stage * Cfg::STAGE, sb = sa + Cfg::A_BYTES;
        const int n0 = nt * BN + rank * Cfg::BN_LOCAL;  // first B row staged by this CTA
        auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
          if (CTA2) tma_load_2d_2sm(dst, m, full_bar(stage), c0, c1);
          else tma_load_2d(dst, m, full_bar(stage), c0, c1);
        };
        if (!A_MN) {
          load(sa, ma, kk + ac, mt * 128 + ar);
        } else {
          load(sa, ma, mt * 128 + ac, kk + ar);
          load(sa + 8192, ma, mt * 128 + 64 + ac, kk + ar);
        }
        if (!B_MN) {
          load(sb, mb, kk + bc, n0 + br);
        } else {
#pragma unroll
          for (int j = 0; j < Cfg::BN_LOCAL / 64; ++j) load(sb + j * 8192, mb, n0 + j * 64 + bc, kk + br);
        }
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------ MMA issuer --------------------------------
    // The whole warp follows the control flow (barrier waits); one elected lane issues.  Descriptors are
    // a per-stage base plus a constant per K step, so the issuing thread's instruction stream per MMA is
    // a single add: it shares its scheduler with two epilogue warps and every extra instruction is
    // tensor-pipe idle time.
    const uint32_t idesc = make_idesc_bf16(CTA2 ? 256 : 128, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const uint64_t da0 = A_MN ? make_smem_desc(sbase, 8192, 1024) : make_smem_desc(sbase, 16, 1024);
    const uint64_t db0 = B_MN ? make_smem_desc(sbase + Cfg::A_BYTES, 8192, 1024)
                              : make_smem_desc(sbase + Cfg::A_BYTES, 16, 1024);
    constexpr uint32_t A_STEP = A_MN ? 2048 : 32, B_STEP = B_MN ? 2048 : 32;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    GemmWalk walk(p, work_first, work_step);
    GemmUnit u;
    while (walk.next(u)) {
      const int kb_begin = u.kb_begin, kb_end = u.kb_end;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (kb == kb_begin && lane == 0) GTRACE(2);   // (last tile's value survives)
        if (elect_one()) {
          const uint64_t da = desc_adv(da0, stage * Cfg::STAGE), db = desc_adv(db0, stage * Cfg::STAGE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (kb > kb_begin || k > 0) ? 1u : 0u;
            if (CTA2) umma_ss_2cta(d_tmem, desc_adv(da, k * A_STEP), desc_adv(db, k * B_STEP), idesc, accum);
            else umma_ss(d_tmem, desc_adv(da, k * A_STEP), desc_adv(db, k * B_STEP), idesc, accum);
          }
          if (CTA2) umma_commit_2cta(empty_bar(stage));  // frees the stage in both CTAs
          else umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) {
        if (CTA2) umma_commit_2cta(tfull_bar(acc));
        else umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (lane == 0) GTRACE(3);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ----------------------------------
    const int ew = (warp - 4) & 3;     // TMEM lane quadrant (== warp % 4)
    const int chalf = (warp - 4) >> 2;  // which half of the tile's columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    GemmWalk walk(p, work_first, work_step);
    GemmUnit u;
    while (walk.next(u)) {
      const int tile = u.tile;
      int tm, nt;
      tile_coords(p, tile, tm, nt);
      const int mt = tm * (CTA2 ? 2 : 1) + rank;
      const int g = p.groups > 1 ? u.z : 0;
      const int crow = g * p.c_gr, ccol = g * p.c_gc;  // output displacement of this group
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (warp == 4 && lane == 0) GTRACE(4);
      // ---- stream-K: partial accumulators travel through global memory ----
      int sk_from[4];  // CTAs (pairs) whose partials of this tile are added here
      int sk_n = 0;
      if (BN >= 128 && u.role == 1) {
        // dump this CTA's fp32 partial, [BN / 4][128][4] so that a warp writes 512 contiguous bytes, and signal
        float* wsb = p.sk_ws + (size_t)(work_first * 2 + rank) * 128 * BN;
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(wsb + ((size_t)(c / 4 + q) * 128 + ew * 32 + lane) * 4) =
                make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                            __uint_as_float(r[q * 4 + 3]));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          int* f = p.sk_flags + (work_first * 2 + rank) * 8 + (warp - 4);
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2) mbar_arrive_even_cta(tempty_bar(acc));
          else mbar_arrive(tempty_bar(acc));
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (BN >= 128 && u.role == 2) {
        // the rest of this tile's k blocks went to the following CTAs (their FIRST unit, so the partials are early)
        const int64_t tile_end = (int64_t)(tile + 1) * walk.kb_all;
        for (int w2 = work_first + 1; w2 < work_step && walk.share_begin(w2) < tile_end; ++w2) {
          if (sk_n == 4) {  // the host bounds the shares so that this cannot happen
            printf("b200 gemm: stream-K tile shared by more than five CTAs\n");
            __trap();
          }
          sk_from[sk_n++] = w2;
        }
        if (lane == 0) {
          for (int ci = 0; ci < sk_n; ++ci) {
            const int* f = p.sk_flags + (sk_from[ci] * 2 + rank) * 8 + (warp - 4);
            int v = 0;
            uint32_t spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              if (!v) {
                __nanosleep(200);
                if (++spins > (1u << 24)) {
                  printf("b200 gemm: stream-K partial of CTA %d never arrived\n", sk_from[ci]);
                  __trap();
                }
              }
            } while (!v);
          }
        }
        __syncwarp();
        __threadfence();
      }
      const int64_t row = (int64_t)mt * 128 + ew * 32 + lane;
      const bool row_ok = row < p.M;
      const bf16* bias_g = p.bias ? p.bias + g * p.bias_g : nullptr;
      const bf16* gate_row = p.gate ? p.gate + (row_ok ? row / p.rows_per_gate : 0) * p.gate_stride : nullptr;
      const bf16* res_row = p.res ? p.res + row * p.ldres : nullptr;
      bf16* aux_row = p.aux ? p.aux + row * p.ldaux : nullptr;
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN;
      if (BN >= 128 && p.tma_store) {
        // Coalesced path: each warp packs a 32-row x 64-column bf16 slab into its swizzled staging
        // buffer and hands it to the TMA (cp.async.bulk.tensor store clips rows/columns past M, N).
        const uint32_t stg = stage_base + (warp - 4) * 4096;
        const int row0 = mt * 128 + ew * 32;
        // The slab loop is compiled three times and chosen per tile: MODE 0 = nothing but the conversion, 1 = bias
        // only, 2 = everything (gate, residual, GELU / GELU' / stash, stream-K partials).  The general code executes
        // ~400 instructions per thread and 32 columns (ncu: 15.7 k warp instructions per 128 x 256 tile, 13 k of them
        // here) whichever operands are present; on 8 warps that is the 5-7 us un-overlapped tail of every launch
        // (tools/gemm_trace.py) and issue slots taken from the MMA warp while it overlaps.  Most launches of a train
        // step (all dgrads, the q/k/v projection) need mode 0 or 1.
        auto run_slabs = [&](auto mode_tag, auto epi_tag, auto bg_tag) {
          constexpr int MODE = decltype(mode_tag)::value;
          constexpr int EPI = decltype(epi_tag)::value;     // general mode: the epilogue function, compile-time
          constexpr bool BG = decltype(bg_tag)::value;      // a bias or a gate vector is present
          const bool stash = MODE == 2 && (EPI == EPI_GELU || EPI == EPI_STASH) && p.aux;  // second store: bf16 pre-activation / pre-gate output
#pragma unroll 1
          for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 64) {
            const int n_slab = nt * BN + c;
            if (n_slab >= p.N) break;
            uint4 prer[8];  // pre-activation of this thread's 64 columns (stash mode), kept for the second store
            // residual / stashed pre-activation of this thread's 64 columns: eight independent row-strided
            // 16-byte loads in flight together, before the TMEM reads (issued one by one behind them they
            // cost ~8 x the L2 latency per slab and made the gate+residual and GELU' epilogues the bottleneck).
            // (A coalesced read -- quarter-warp per row, transposed through the staging buffer -- measured the same:
            // 49.2 vs 49.5 us at N = K = 2048, 164.2 vs 164.2 us for GELU'; the loads are not what these epilogues wait for.)
            uint4 ext[8];
            const bf16* ext_base = nullptr;
            if constexpr (MODE == 2) {
              // Coalesced: a quarter-warp per row, four rows (four full 128-byte lines) per request, eight requests in
              // flight; thread-per-row loads of the same 4 KB were 32 lines per request and 1.2-1.7 us per slab of LSU
              // issue time (tools/gemm_trace.py).  The slab then passes through the staging buffer to the
              // thread-per-row layout of the accumulator.
              ext_base = EPI == EPI_GELU_GRAD ? p.aux : p.res;
              const int64_t ext_ld = EPI == EPI_GELU_GRAD ? p.ldaux : p.ldres;
              if (ext_base != nullptr) {
                const int er = lane >> 3, ec = lane & 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int64_t rr = (int64_t)row0 + j * 4 + er;
                  ext[j] = make_uint4(0, 0, 0, 0);
                  if (rr < p.M && n_slab + ec * 8 < p.N)
                    ext[j] = *reinterpret_cast<const uint4*>(ext_base + rr * ext_ld + n_slab + ec * 8);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) ext[j] = make_uint4(0, 0, 0, 0);
              }
            }
            if (lane == 0) tma_store_wait_read<0>();  // the previous slab has left the staging buffer
            __syncwarp();
            if constexpr (MODE == 2) {
              if (ext_base != nullptr) {
                const int er = lane >> 3, ec = lane & 7;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + sw128_off(j * 4 + er, ec)),
                               "r"(ext[j].x), "r"(ext[j].y), "r"(ext[j].z), "r"(ext[j].w)
                               : "memory");
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(ext[j].x), "=r"(ext[j].y), "=r"(ext[j].z), "=r"(ext[j].w)
                               : "r"(stg + sw128_off(lane, j))
                               : "memory");
                __syncwarp();   // every lane holds its row before any lane overwrites the slab with outputs
              }
            }
            if (warp == 4 && lane == 0) GTRACE(c == 0 ? 8 : 12);
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
              uint32_t r[32];
              tmem_ld32(t_row + c + hc * 32, r);
              // bias and gate of these 32 columns: eight loads in flight under the TMEM read
              uint4 bq[4], gq[4];
              if constexpr (BG) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const int n = n_slab + hc * 32 + g * 8;
                  bq[g] = gq[g] = make_uint4(0, 0, 0, 0);
                  if (n < p.N) {
                    if (bias_g) bq[g] = __ldg(reinterpret_cast<const uint4*>(bias_g + n));
                    if (MODE == 2 && gate_row) gq[g] = __ldg(reinterpret_cast<const uint4*>(gate_row + n));
                  }
                }
              }
              tmem_ld_wait();
              if (warp == 4 && lane == 0 && hc == 0) GTRACE(c == 0 ? 9 : 13);
              if (warp == 4 && lane == 0) GTRACE(16 + (c == 0 ? 0 : 4) + hc * 2);
              if constexpr (MODE == 2) {
                for (int ci = 0; ci < sk_n; ++ci) {
                  const float* pb = p.sk_ws + (size_t)(sk_from[ci] * 2 + rank) * 128 * BN;
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(
                        pb + ((size_t)((c + hc * 32) / 4 + q) * 128 + ew * 32 + lane) * 4));
                    r[q * 4 + 0] = __float_as_uint(__uint_as_float(r[q * 4 + 0]) + a.x);
                    r[q * 4 + 1] = __float_as_uint(__uint_as_float(r[q * 4 + 1]) + a.y);
                    r[q * 4 + 2] = __float_as_uint(__uint_as_float(r[q * 4 + 2]) + a.z);
                    r[q * 4 + 3] = __float_as_uint(__uint_as_float(r[q * 4 + 3]) + a.w);
                  }
                }
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
                uint4 u;
                if constexpr (MODE == 2) {
                  const int n = n_slab + hc * 32 + g * 8;
                  uint4 pre = make_uint4(0, 0, 0, 0);
                  if (n < p.N)
                    epi_math8<EPI>(v, pre, p, BG && bias_g != nullptr, bq[g], BG && gate_row != nullptr, gq[g],
                                   p.res != nullptr, ext[hc * 4 + g]);
                  prer[hc * 4 + g] = pre;
                } else if constexpr (MODE == 1) {   // columns past N get the zero bias loaded above and are clipped by the TMA
                  const uint32_t bw[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {   // packed fp32x2: one add per pair of columns
                    const float2 r2 = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), make_float2(bf16_lo(bw[i]), bf16_hi(bw[i])));
                    v[2 * i] = r2.x; v[2 * i + 1] = r2.y;
                  }
                }
                u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
                u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + sw128_off(lane, hc * 4 + g)),
                             "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                             : "memory");
              }
              if (warp == 4 && lane == 0) GTRACE(17 + (c == 0 ? 0 : 4) + hc * 2);
            }
            if (warp == 4 && lane == 0) GTRACE(c == 0 ? 10 : 14);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, stg, n_slab + ccol, row0 + crow);
              tma_store_commit();
            }
            if (warp == 4 && lane == 0) GTRACE(c == 0 ? 11 : 15);
            if constexpr (MODE == 2) {
              if (stash) {
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + sw128_off(lane, j)), "r"(prer[j].x),
                               "r"(prer[j].y), "r"(prer[j].z), "r"(prer[j].w)
                               : "memory");
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmAux, stg, n_slab + ccol, row0 + crow);
                  tma_store_commit();
                }
              }
            }
          }
        };
        const bool light = p.epi == EPI_NONE && p.gate == nullptr && p.res == nullptr && p.aux == nullptr && sk_n == 0;
        const bool bg = bias_g != nullptr || gate_row != nullptr;
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
#define B200_GENERAL(E)                                                              \
  do {                                                                               \
    if (bg) run_slabs(I2{}, std::integral_constant<int, E>{}, std::true_type{});     \
    else run_slabs(I2{}, std::integral_constant<int, E>{}, std::false_type{});       \
  } while (0)
        if (light && !bg) run_slabs(I0{}, std::integral_constant<int, EPI_NONE>{}, std::false_type{});
        else if (light) run_slabs(I1{}, std::integral_constant<int, EPI_NONE>{}, std::true_type{});
        else if (p.epi == EPI_GELU) B200_GENERAL(EPI_GELU);
        else if (p.epi == EPI_GELU_GRAD) B200_GENERAL(EPI_GELU_GRAD);
        else if (p.epi == EPI_STASH) B200_GENERAL(EPI_STASH);
        else B200_GENERAL(EPI_NONE);
#undef B200_GENERAL
      } else
#pragma unroll 1
      for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
        const int n0 = nt * BN + c;
        if (n0 >= p.N) break;
        uint32_t r[32];
        tmem_ld32(t_row + c, r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int n = n0 + g * 8;
            if (n < p.N) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
              if (bias_g) {
                uint4 u = __ldg(reinterpret_cast<const uint4*>(bias_g + n));
                v[0] += bf16_lo(u.x); v[1] += bf16_hi(u.x); v[2] += bf16_lo(u.y); v[3] += bf16_hi(u.y);
                v[4] += bf16_lo(u.z); v[5] += bf16_hi(u.z); v[6] += bf16_lo(u.w); v[7] += bf16_hi(u.w);
              }
              if (p.epi == EPI_GELU) {
                if (aux_row) {
                  uint4 u;
                  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
                  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
                  *reinterpret_cast<uint4*>(aux_row + n) = u;
                  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
                  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = gelu_tanh(v[i]);
              } else if (p.epi == EPI_GELU_GRAD) {
                uint4 u = *reinterpret_cast<const uint4*>(aux_row + n);
                float h[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                              bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] *= gelu_tanh_grad(h[i]);
              } else if (p.epi == EPI_STASH) {
                uint4 u;
                u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
                u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(aux_row + n) = u;
                v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
                v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
              }
              if (gate_row) {
                uint4 u = __ldg(reinterpret_cast<const uint4*>(gate_row + n));
                v[0] *= bf16_lo(u.x); v[1] *= bf16_hi(u.x); v[2] *= bf16_lo(u.y); v[3] *= bf16_hi(u.y);
                v[4] *= bf16_lo(u.z); v[5] *= bf16_hi(u.z); v[6] *= bf16_lo(u.w); v[7] *= bf16_hi(u.w);
              }
              if (res_row) {
                uint4 u = *reinterpret_cast<const uint4*>(res_row + n);
                v[0] += bf16_lo(u.x); v[1] += bf16_hi(u.x); v[2] += bf16_lo(u.y); v[3] += bf16_hi(u.y);
                v[4] += bf16_lo(u.z); v[5] += bf16_hi(u.z); v[6] += bf16_lo(u.w); v[7] += bf16_hi(u.w);
              }
              if (p.out_f32) {
                float* o = reinterpret_cast<float*>(p.C) + (row + crow) * p.ldc + n + ccol;
                if (p.splits > 1) {  // split-K partial: accumulate into the caller-zeroed output
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
                } else {
                  *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                  *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
              } else {
                uint4 u;
                u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
                u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.C) + (row + crow) * p.ldc + n + ccol) = u;
              }
            }
          }
        }
      }
      if (sk_n > 0) {
        __syncwarp();
        if (lane == 0)
          for (int ci = 0; ci < sk_n; ++ci) p.sk_flags[(sk_from[ci] * 2 + rank) * 8 + (warp - 4)] = 0;  // consumed
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_even_cta(tempty_bar(acc));
        else mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (warp == 4 && lane == 0) GTRACE(5);
    // the staging slab only has to stay valid until the TMA has READ it; the writes complete with the kernel
    if (lane == 0) tma_store_wait_read<0>();
    if (warp == 4 && lane == 0) GTRACE(6);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair's MMAs / signals are in flight
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) GTRACE(7);
}

static int g_num_sms = 0;
// kernel-experiment switches, read once (not per launch)
static bool env_flag_no_pair() {
  static const bool v = getenv("B200_GEMM_NO_PAIR") != nullptr;
  return v;
}
static bool env_flag_no_streamk() {
  static const bool v = getenv("B200_GEMM_NO_STREAMK") != nullptr;
  return v;
}

static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool A_MN, bool B_MN, bool CTA2 = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2,
                       const CUtensorMap& tmB2, const CUtensorMap& tmC, const CUtensorMap& tmAux,
                       const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTA2>;
  static bool attr_set = false;
  auto kern = gemm_kernel<BN, A_MN, B_MN, CTA2>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return launch_status("gemm: cudaFuncSetAttribute");
    attr_set = true;
  }
  int tiles = p.m_tiles * p.n_tiles * p.splits * p.groups;  // CTA2: m_tiles counts 256-row tile pairs
  const int slots = CTA2 ? num_sms() / 2 : num_sms();
  const int ctas = (tiles < slots ? tiles : slots) * (CTA2 ? 2 : 1);
  PdlLaunch L(dim3(ctas, 1, 1), dim3(384, 1, 1), Cfg::SMEM, stream, CTA2 ? 2 : 1);
  cudaError_t e = cudaLaunchKernelEx(&L.cfg, kern, tmA, tmB, tmA2, tmB2, tmC, tmAux, p);
  if (e != cudaSuccess) return launch_status(CTA2 ? "gemm_bf16 (CTA pair launch)" : "gemm_bf16");
  return launch_status("gemm_bf16");
}

}  // namespace b200

using namespace b200;

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 16 KB of flags + one 128 x 256 fp32 partial per CTA and rank slot of the largest grid
extern "C" int64_t b200_gemm_workspace_bytes(void) {
  return 16384 + (int64_t)num_sms() * 2 * 128 * 256 * (int64_t)sizeof(float);
}

// Strided-batch description (b200_gemm_bf16_batched): `groups` problems of identical shape whose operands
// are sub-blocks of the given tensors displaced by g * (rows, cols) elements.
struct GemmGroups {
  int groups;
  int a_gr, a_gc, b_gr, b_gc, a2_gr, a2_gc, b2_gr, b2_gc, c_gr, c_gc, bias_g;
};

static int gemm_impl(const void* A, int64_t lda, int a_kmajor_rows_are_k, const void* B,
                              int64_t ldb, int b_rows_are_k, const void* A2, int64_t lda2,
                              const void* B2, int64_t ldb2, int K2, void* C, int64_t ldc,
                              int out_is_f32, int M, int N, int K, int epilogue, const void* bias,
                              const void* gate, int64_t gate_stride, int64_t rows_per_gate,
                              const void* res, int64_t ldres, void* aux, int64_t ldaux,
                              int block_n, int split_k, void* stream, const GemmGroups& gg,
                              void* workspace = nullptr, int64_t workspace_bytes = 0) {
  const bool a_mn = a_kmajor_rows_are_k != 0, b_mn = b_rows_are_k != 0;
  if (M < 0 || N < 0 || K < 0 || K2 < 0) return arg_error("gemm_bf16: negative dimension");
  if (M == 0 || N == 0) return 0;  // empty output: nothing to do (empty tensors carry null pointers)
  if (!(A && B && C)) return arg_error("gemm_bf16: null operand");
  if (K == 0 && K2 == 0) return arg_error("gemm_bf16: empty reduction");
  // an extent must be a multiple of 8 only where it is the contiguous dimension of some operand: N always
  // (C rows, bias); K / K2 unless both operands are stored [K, *] (the wgrad form: K = token count, any value,
  // TMA zero-fills the rows past it); M for a [K, M] A operand
  const bool k_contig = !a_mn || !b_mn;
  if (N % 8 || (k_contig && (K % 8 || K2 % 8)) || (a_mn && M % 8))
    return arg_error("gemm_bf16: N (and K, K2 for a K-major operand, M for a [K,M] A operand) must be multiples of 8");
  if (lda % 8 || ldb % 8 || ldc % (out_is_f32 ? 4 : 8) || !al16(A) || !al16(B) || !al16(C))
    return arg_error("gemm_bf16: operands must be 16-byte aligned with 16-byte-multiple pitches");
  if (K2 > 0 && (!(A2 && B2) || lda2 % 8 || ldb2 % 8 || !al16(A2) || !al16(B2)))
    return arg_error("gemm_bf16: bad second operand pair");
  if (epilogue < EPI_NONE || epilogue > EPI_STASH) return arg_error("gemm_bf16: unknown epilogue");
  if (epilogue == EPI_GELU_GRAD && !aux) return arg_error("gemm_bf16: GELU_GRAD needs aux (pre-activation)");
  if (epilogue == EPI_STASH && (!aux || out_is_f32)) return arg_error("gemm_bf16: STASH needs aux and a bf16 output");
  if ((aux && (ldaux % 8 || !al16(aux))) || (res && (ldres % 8 || !al16(res))) ||
      (bias && !al16(bias)) || (gate && (gate_stride % 8 || !al16(gate) || rows_per_gate <= 0)))
    return arg_error("gemm_bf16: epilogue operands must be 16-byte aligned");

  int bn = block_n;
  if (bn == 0) {
    const int mt = (M + 127) / 128;
    bn = 64;
    for (int cand : {256, 128}) {
      if (N >= cand && mt * ((N + cand - 1) / cand) >= 96) { bn = cand; break; }
    }
    if (bn == 64 && N > 64 && mt * ((N + 127) / 128) >= 48) bn = 128;
  }
  if (bn != 64 && bn != 128 && bn != 256) return arg_error("gemm_bf16: block_n must be 0, 64, 128 or 256");

  // CTA pairs (256 x 256 tiles, cta_group::2) for the large GEMMs: a third less operand traffic per FLOP
  const bool pair = bn == 256 && M >= 512 && split_k <= 1 && (gg.groups <= 1 || M % 256 == 0) &&
                    !env_flag_no_pair();
  GemmParams p;
  p.M = M; p.N = N;
  p.m_tiles = pair ? (M + 255) / 256 : (M + 127) / 128;
  p.n_tiles = (N + bn - 1) / bn;
  {
    // one wave of concurrent CTAs (pairs) ~ group_m tile rows x all tile columns
    static const int fixed = [] { const char* e = getenv("B200_GEMM_GROUP_M"); return e ? atoi(e) : 0; }();
    const int conc = pair ? num_sms() / 2 : num_sms();
    int gm = fixed > 0 ? fixed : (conc + p.n_tiles / 2) / p.n_tiles;
    // operands that fit the 126 MB L2 together are re-used there whatever the order (measured: the plain order is
    // ~1 % faster at K = 2048); group only when they do not (K = 8192 / 6144: 339 -> 211 MB of DRAM reads, 3-4 % faster)
    if (fixed <= 0 && ((int64_t)M + N) * (int64_t)(K + K2) * 2 <= (int64_t)64 << 20) gm = p.m_tiles;
    if (gm < 1) gm = 1;
    if (gm > p.m_tiles) gm = p.m_tiles;
    p.group_m = gm;
  }
  p.kb1 = (K + 63) / 64;
  p.kb2 = (K2 + 63) / 64;
  p.epi = epilogue; p.out_f32 = out_is_f32;
  {
    const bool plain = out_is_f32 && epilogue == EPI_NONE && !bias && !gate && !res && !aux;
    const int kb_all = p.kb1 + p.kb2;
    int sp = split_k <= 1 ? 1 : split_k;
    if (sp > 1 && !plain) return arg_error("gemm_bf16: split_k > 1 needs a plain fp32 output (no epilogue operands)");
    if (sp > kb_all) sp = kb_all;
    // every split must own at least one k block
    while (sp > 1 && (sp - 1) * ((kb_all + sp - 1) / sp) >= kb_all) --sp;
    p.splits = sp;
  }
  p.groups = gg.groups < 1 ? 1 : gg.groups;
  p.a_gr = gg.a_gr; p.a_gc = gg.a_gc; p.b_gr = gg.b_gr; p.b_gc = gg.b_gc;
  p.a2_gr = gg.a2_gr; p.a2_gc = gg.a2_gc; p.b2_gr = gg.b2_gr; p.b2_gc = gg.b2_gc;
  p.c_gr = gg.c_gr; p.c_gc = gg.c_gc; p.bias_g = gg.bias_g;
  const int64_t G1 = p.groups - 1;
  if (p.groups > 1) {
    // tiles must not straddle sub-blocks: TMA zero-fill / store clipping only exist at the tensor edges
    if (M % 128 || N % bn || K % 64 || K2 % 64)
      return arg_error("gemm_bf16_batched: per-group M, N, K, K2 must be multiples of 128, block_n, 64, 64");
    if (gate || res || aux || epilogue != EPI_NONE)
      return arg_error("gemm_bf16_batched: only the bias epilogue is built for batches");
    if (p.splits > 1) return arg_error("gemm_bf16_batched: split_k and batching exclude each other");
    if (gg.c_gc % 8 || gg.bias_g % 8) return arg_error("gemm_bf16_batched: output / bias displacements must be multiples of 8");
  }
  p.C = C; p.ldc = ldc;
  p.bias = (const bf16*)bias;
  p.gate = (const bf16*)gate; p.gate_stride = gate_stride; p.rows_per_gate = rows_per_gate > 0 ? rows_per_gate : 1;
  p.res = (const bf16*)res; p.ldres = ldres;
  p.aux = (bf16*)aux; p.ldaux = ldaux;

  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  // tensor extents include the displaced sub-blocks of every group
  auto mapA = [&](CUtensorMap* t, const void* ptr, int64_t ld, int kdim, int gr, int gc) {
    return a_mn ? make_tmap_2d_bf16(t, ptr, kdim + G1 * gr, M + G1 * gc, ld, 64, 64)
                : make_tmap_2d_bf16(t, ptr, M + G1 * gr, kdim + G1 * gc, ld, 128, 64);
  };
  auto mapB = [&](CUtensorMap* t, const void* ptr, int64_t ld, int kdim, int gr, int gc) {
    return b_mn ? make_tmap_2d_bf16(t, ptr, kdim + G1 * gr, N + G1 * gc, ld, 64, 64)
                : make_tmap_2d_bf16(t, ptr, N + G1 * gr, kdim + G1 * gc, ld, pair ? bn / 2 : bn, 64);
  };
  if (K > 0) {
    if ((rc = mapA(&tmA, A, lda, K, gg.a_gr, gg.a_gc)) || (rc = mapB(&tmB, B, ldb, K, gg.b_gr, gg.b_gc)))
      return arg_error("gemm_bf16: cuTensorMapEncodeTiled failed", rc);
  }
  if (K2 > 0) {
    if ((rc = mapA(&tmA2, A2, lda2, K2, gg.a2_gr, gg.a2_gc)) || (rc = mapB(&tmB2, B2, ldb2, K2, gg.b2_gr, gg.b2_gc)))
      return arg_error("gemm_bf16: cuTensorMapEncodeTiled failed (pair 2)", rc);
  } else {
    tmA2 = tmA; tmB2 = tmB;
  }
  if (K == 0) { tmA = tmA2; tmB = tmB2; }
  // bf16 outputs of the wide tiles leave through swizzled smem slabs + TMA stores (coalesced 128-byte
  // rows instead of 16-byte-per-lane row-strided stores)
  CUtensorMap tmC = tmA, tmAux = tmA;
  p.tma_store = (!out_is_f32 && bn >= 128) ? 1 : 0;
  if (p.tma_store) {
    if ((rc = make_tmap_2d_bf16(&tmC, C, M + G1 * gg.c_gr, N + G1 * gg.c_gc, ldc, 32, 64)))
      return arg_error("gemm_bf16: cuTensorMapEncodeTiled failed (C)", rc);
    if ((epilogue == EPI_GELU || epilogue == EPI_STASH) && aux && (rc = make_tmap_2d_bf16(&tmAux, aux, M, N, ldaux, 32, 64)))
      return arg_error("gemm_bf16: cuTensorMapEncodeTiled failed (aux)", rc);
  }

  // stream-K over the last, partial wave of tiles (see GemmParams): worth it when that wave would leave a good part
  // of the SMs idle, and bounded so that a tile is never shared by more than four CTAs
  p.sk_tiles = 0;
  p.sk_ws = nullptr;
  p.sk_flags = nullptr;
  if (workspace && p.tma_store && p.splits == 1 && p.groups == 1 && !env_flag_no_streamk()) {
    const int P = pair ? num_sms() / 2 : num_sms();
    const int T = p.m_tiles * p.n_tiles;
    const int R = T % P;
    if (T > P && R * 3 >= P && R * 10 <= 9 * P && p.kb1 + p.kb2 >= 8) {
      if (workspace_bytes < b200_gemm_workspace_bytes() || !al16(workspace))
        return arg_error("gemm_bf16: workspace too small or misaligned (see b200_gemm_workspace_bytes)");
      p.sk_tiles = R;
      p.sk_flags = (int*)workspace;   // first 16 KB: flags (zero between launches)
      p.sk_ws = (float*)((char*)workspace + 16384);
    }
  }

  cudaStream_t st = (cudaStream_t)stream;
#define DISPATCH(BN_)                                                                     \
  if (bn == BN_) {                                                                        \
    if (!a_mn && !b_mn) return launch_gemm<BN_, false, false>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st); \
    if (!a_mn && b_mn) return launch_gemm<BN_, false, true>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);   \
    if (a_mn && !b_mn) return launch_gemm<BN_, true, false>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);   \
    return launch_gemm<BN_, true, true>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);                       \
  }
  if (pair) {
    if (!a_mn && !b_mn) return launch_gemm<256, false, false, true>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);
    if (!a_mn && b_mn) return launch_gemm<256, false, true, true>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);
    if (a_mn && !b_mn) return launch_gemm<256, true, false, true>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);
    return launch_gemm<256, true, true, true>(tmA, tmB, tmA2, tmB2, tmC, tmAux, p, st);
  }
  DISPATCH(256)
  DISPATCH(128)
  DISPATCH(64)
#undef DISPATCH
  return arg_error("gemm_bf16: unreachable");
}

#ifdef B200_TRACE
extern "C" int b200_debug_gemm_trace(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_gemm_trace, sizeof(unsigned long long) * n);
}
#endif

// See include/b200ltx.h for the contracts.
extern "C" int b200_gemm_bf16(const void* A, int64_t lda, int a_kmajor_rows_are_k, const void* B,
                              int64_t ldb, int b_rows_are_k, const void* A2, int64_t lda2,
                              const void* B2, int64_t ldb2, int K2, void* C, int64_t ldc,
                              int out_is_f32, int M, int N, int K, int epilogue, const void* bias,
                              const void* gate, int64_t gate_stride, int64_t rows_per_gate,
                              const void* res, int64_t ldres, void* aux, int64_t ldaux,
                              int block_n, int split_k, void* stream) {
  GemmGroups gg = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  return gemm_impl(A, lda, a_kmajor_rows_are_k, B, ldb, b_rows_are_k, A2, lda2, B2, ldb2, K2, C, ldc, out_is_f32,
                   M, N, K, epilogue, bias, gate, gate_stride, rows_per_gate, res, ldres, aux, ldaux, block_n,
                   split_k, stream, gg);
}

extern "C" int b200_gemm_bf16_ws(const void* A, int64_t lda, int a_kmajor_rows_are_k, const void* B,
                                 int64_t ldb, int b_rows_are_k, const void* A2, int64_t lda2,
                                 const void* B2, int64_t ldb2, int K2, void* C, int64_t ldc,
                                 int out_is_f32, int M, int N, int K, int epilogue, const void* bias,
                                 const void* gate, int64_t gate_stride, int64_t rows_per_gate,
                                 const void* res, int64_t ldres, void* aux, int64_t ldaux,
                                 int block_n, int split_k, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  GemmGroups gg = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  return gemm_impl(A, lda, a_kmajor_rows_are_k, B, ldb, b_rows_are_k, A2, lda2, B2, ldb2, K2, C, ldc, out_is_f32,
                   M, N, K, epilogue, bias, gate, gate_stride, rows_per_gate, res, ldres, aux, ldaux, block_n,
                   split_k, stream, gg, workspace, workspace_bytes);
}

extern "C" int b200_gemm_bf16_batched(const void* A, int64_t lda, int a_rows_are_k, const void* B, int64_t ldb,
                                      int b_rows_are_k, const void* A2, int64_t lda2, const void* B2,
                                      int64_t ldb2, int K2, void* C, int64_t ldc, int out_is_f32, int M, int N,
                                      int K, const void* bias, int block_n, int groups,
                                      const int32_t* group_offsets /* host, 11 ints */, void* stream) {
  if (groups < 1 || !group_offsets) return arg_error("gemm_bf16_batched: bad group description");
  const int32_t* o = group_offsets;
  GemmGroups gg = {groups, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10]};
  if (block_n == 0) block_n = N >= 256 && N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);
  return gemm_impl(A, lda, a_rows_are_k, B, ldb, b_rows_are_k, A2, lda2, B2, ldb2, K2, C, ldc, out_is_f32, M, N, K,
                   EPI_NONE, bias, nullptr, 0, 0, nullptr, 0, nullptr, 0, block_n, 1, stream, gg);
}
