/* b200ltx.h — C ABI of libb200ltx.so: the sm_100a kernels behind the LTX-Video-2B transformer
 * block forward/backward (the hot path of lusinlu/Video-Generation-for-Human-Avatars).
 *
 * The reference is 100 % Python and owns no FFI: its "binding surface" for this path is the
 * nn.Module / attention-processor protocol (ltx_video/models/transformers/attention.py:532-552,
 * 660-718, 935-955).  Each entry point below replaces the library kernels PyTorch launches for the
 * cited reference lines; the ctypes stub a maintainer adds is shown in INTEGRATION.md and lives in
 * video-generation-for-human-avatars_b200/lib.py.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller.  The
 *     library never allocates, frees or retains device memory.
 *   - bf16 tensors are row-major with an explicit row pitch `ld*` in ELEMENTS; pointers must be
 *     16-byte aligned and pitches multiples of 8 elements.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit sync.
 *   - return 0 on success; < 0 = argument / shape / alignment contract violation (nothing was
 *     launched); > 0 = cudaError_t of the launch.  b200_last_error() gives a thread-local message.
 *   - there is NO CPU fallback: on a non-sm_100 device b200_device_check() fails and launches
 *     return cudaErrorNoKernelImageForDevice.
 */
#ifndef B200LTX_H_
#define B200LTX_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int b200_version(void);
const char* b200_last_error(void);
int b200_device_check(void);

/* GEMM epilogue selector for b200_gemm_bf16 */
#define B200_EPI_NONE 0
#define B200_EPI_GELU 1      /* out = gelu_tanh(acc + bias); if aux != NULL the bf16 pre-activation is stored there */
#define B200_EPI_GELU_GRAD 2 /* out = acc * gelu_tanh'(aux) */

/* C[M,N] = epi( A * B^T (+ A2 * B2^T) ), bf16 operands, fp32 accumulation in TMEM (tcgen05).
 *   A : [M,K] row-major (a_rows_are_k = 0) or [K,M] row-major (a_rows_are_k = 1)
 *   B : [N,K] row-major (b_rows_are_k = 0, the nn.Linear weight layout) or [K,N] (b_rows_are_k = 1)
 *   A2/B2/K2 : optional second operand pair with the same layouts, accumulated into the same tile
 *              (the LoRA up-projection (s x A^T) B^T of peft lora.Linear; K2 = 0 to disable)
 *   epilogue order: + bias[N] -> GELU / GELU' -> * gate[row / rows_per_gate, N] -> + res[M,N]
 *   C is bf16 (out_is_f32 = 0) or fp32 (1).  block_n = 0 picks the tile width (64/128/256).
 *   split_k: <= 1 = off; n > 1 = split the reduction over n CTAs per tile (skinny wgrad-type GEMMs);
 *            only for a plain fp32 output, which the CALLER MUST ZERO (partials are accumulated with
 *            red.global.add.f32).
 * Replaces: nn.Linear (cuBLASLt) + bias + F.gelu + AdaLN gate + residual adds of
 *   attention.py:996-1014,1089,265-268,285,305-308,1238-1263; transformer3d.py:470,494-499,561;
 *   and, with the [K,*] layouts, the dgrad / wgrad GEMMs autograd runs for them (training.py:203). */
int b200_gemm_bf16(const void* A, int64_t lda, int a_rows_are_k, const void* B, int64_t ldb,
                   int b_rows_are_k, const void* A2, int64_t lda2, const void* B2, int64_t ldb2,
                   int K2, void* C, int64_t ldc, int out_is_f32, int M, int N, int K, int epilogue,
                   const void* bias, const void* gate, int64_t gate_stride, int64_t rows_per_gate,
                   const void* res, int64_t ldres, void* aux, int64_t ldaux, int block_n,
                   int split_k, void* stream);

/* b200_gemm_bf16 with a caller-owned workspace of b200_gemm_workspace_bytes() bytes that enables stream-K for the
 * last, partial wave of output tiles: the k blocks of those tiles are dealt out evenly to all CTAs, partial fp32
 * accumulators travel through the workspace and the CTA holding a tile's first k block runs its epilogue (same
 * result up to fp32 summation order; deterministic).  The workspace's first 16 KB are flags: the caller zeroes them
 * ONCE, every launch leaves them zero.  One workspace per stream: launches that may overlap must not share it.
 * workspace = NULL behaves exactly like b200_gemm_bf16. */
int64_t b200_gemm_workspace_bytes(void);
int b200_gemm_bf16_ws(const void* A, int64_t lda, int a_rows_are_k, const void* B, int64_t ldb,
                   int b_rows_are_k, const void* A2, int64_t lda2, const void* B2, int64_t ldb2,
                   int K2, void* C, int64_t ldc, int out_is_f32, int M, int N, int K, int epilogue,
                   const void* bias, const void* gate, int64_t gate_stride, int64_t rows_per_gate,
                   const void* res, int64_t ldres, void* aux, int64_t ldaux, int block_n,
                   int split_k, void* workspace, int64_t workspace_bytes,
                     void* stream);

/* Strided batch of `groups` identically shaped GEMMs in ONE launch:  C_g[M,N] = A_g * B_g^T (+ A2_g * B2_g^T)
 * (+ bias_g).  Operand g is the sub-block of the given tensor displaced by g * (rows, cols) elements;
 * `group_offsets` is a HOST array of 11 ints {a_rows, a_cols, b_rows, b_cols, a2_rows, a2_cols, b2_rows,
 * b2_cols, c_rows, c_cols, bias}.  Per-group M, N, K, K2 must be multiples of 128, block_n, 64, 64 (tiles
 * never straddle sub-blocks).  Used for everything on the attn2 key/value side that depends only on the
 * projected caption tokens and can therefore run once for all 28 blocks (reference: the per-block
 * to_k / to_v nn.Linear + peft LoRA calls, attention.py:999-1005). */
int b200_gemm_bf16_batched(const void* A, int64_t lda, int a_rows_are_k, const void* B, int64_t ldb,
                           int b_rows_are_k, const void* A2, int64_t lda2, const void* B2, int64_t ldb2,
                           int K2, void* C, int64_t ldc, int out_is_f32, int M, int N, int K,
                           const void* bias, int block_n, int groups, const int32_t* group_offsets,
                           void* stream);

/* Flash attention forward, head_dim 64, non-causal.  q/k/v/o token-major [B*N, ld], head h in
 * columns [64h, 64h+64).  key_bias: optional fp32 [B,Nk] additive score bias (the -10000 mask bias).
 * lse: optional fp32 [B,H,Nq] log-sum-exp for the backward.
 * Replaces F.scaled_dot_product_attention, attention.py:1057-1064 (+ mask prep :981-989). */
int b200_fa_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                void* o, int64_t ldo, float* lse, const float* key_bias, int B, int H, int Nq, int Nk,
                int head_dim, float scale, void* stream);
/* The same with a caller-owned scratch buffer (16-byte aligned, >= b200_fa_fwd_workspace_bytes(B,H,Nq,Nk); 0 bytes /
 * NULL allowed).  With it, when the (query tile, head, batch) work items fill their last wave of CTAs sparsely, the
 * items of that wave are each split along the keys over several CTAs whose partial (O, max, sum) go through the
 * workspace and are folded by a second small kernel of the same call; results match b200_fa_fwd to rounding.
 * batch_keep: optional fp32 [B]; entries equal to 0 get rows of `pass_src` (bf16 [B*Nq, ld_pass]; NULL = the value
 * rows v, which needs Nq == Nk) as output instead of the attention result -- the spatio-temporal-guidance skips of
 * attention.py:1071-1086 ("attention values": v, "attention skip": the attention input), mask values 0 / 1: their
 * CTAs copy one tile and do no attention work. */
int64_t b200_fa_fwd_workspace_bytes(int B, int H, int Nq, int Nk);
int b200_fa_fwd_ws(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                   void* o, int64_t ldo, float* lse, const float* key_bias, const float* batch_keep,
                   const void* pass_src, int64_t ld_pass, int B, int H, int Nq, int Nk, int head_dim, float scale,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* Online-softmax merge of a partial attention result (o_i bf16, lse_i) over a disjoint key shard into
 * fp32 accumulators (o_acc, lse_acc); first != 0 initialises them; out (bf16, may be NULL) receives the
 * merged rows.  One call per hop of the sequence-sharded ring attn1 (no reference counterpart: the
 * reference runs one F.scaled_dot_product_attention over all tokens, attention.py:1057). */
int b200_attn_merge(float* o_acc, int64_t ldacc, float* lse_acc, const void* o_i, int64_t ldo,
                    const float* lse_i, void* out, int64_t ldout, int B, int H, int N, int first,
                    void* stream);

/* delta[b,h,q] = sum_d o * do  (backward pre-pass). */
int b200_attn_delta(const void* o, int64_t ldo, const void* dout, int64_t lddo, float* delta, int B,
                    int H, int Nq, void* stream);
/* The same pass, also clearing the fp32 dQ accumulator [B*Nq, lddq] (first H*64 columns of every row) that
 * b200_fa_bwd reduces into: one launch instead of the pre-pass plus a fill (the reference's SDPA backward,
 * attention.py:1057, owns both inside the library). dq_zero may be NULL. */
int b200_attn_delta_zero(const void* o, int64_t ldo, const void* dout, int64_t lddo, float* delta,
                         float* dq_zero, int64_t lddq, int B, int H, int Nq, void* stream);

/* Flash attention backward.  dq_accum is fp32 [B*Nq, lddq] and MUST be zeroed by the caller (key-tile
 * CTAs reduce into it with TMA reduce-add); dk/dv are bf16.  With few key tiles (attn2: 256 caption
 * tokens) the query walk of a key tile is split over several CTAs whose fp32 dK/dV partials go through
 * `workspace` (caller-owned, 16-byte aligned, >= b200_fa_bwd_workspace_bytes(B,H,Nq,Nk), may be NULL
 * when that is 0) and are summed by a second small kernel launched by the same call.
 * Replaces the SDPA backward autograd runs under training.py:203. */
int64_t b200_fa_bwd_workspace_bytes(int B, int H, int Nq, int Nk);
int b200_fa_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                const void* dout, int64_t lddo, const float* lse, const float* delta,
                const float* key_bias, float* dq_accum, int64_t lddq, void* dk, int64_t lddk, void* dv,
                int64_t lddv, int B, int H, int Nq, int Nk, int head_dim, float scale, void* workspace,
                int64_t workspace_bytes, void* stream);

/* y = norm(x) * (1 + scale[b]) + shift[b]; RMSNorm (layernorm = 0) or LayerNorm (1), no affine.
 * scale/shift: bf16 rows of D, one per `rows_per_mod` consecutive rows, `mod_stride` elements apart
 * (pointers into the [B,6,D] AdaLN tensor); either may be NULL.
 * Replaces diffusers RMSNorm + AdaLN modulate, attention.py:223-236, 288-290; norm_out + modulate,
 * transformer3d.py:554-559. */
int b200_norm_mod_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, const void* scale,
                      const void* shift, int64_t mod_stride, int64_t rows, int D,
                      int64_t rows_per_mod, float eps, int layernorm, void* stream);
/* dx = dres + d(norm_mod)/dx applied to dy  (dres may be NULL).  prod (bf16 [rows, ldprod], may be NULL) receives
 * dy * xhat: its column sums per modulation group are d(scale), the column sums of dy are d(shift) -- the gradients
 * of the AdaLN tables when they train (training.py:75-91, "full" strategy); see b200_colsum_groups. */
int b200_norm_mod_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const void* scale,
                      int64_t mod_stride, const void* dres, int64_t lddres, void* dx, int64_t lddx,
                      void* prod, int64_t ldprod, int64_t rows, int D, int64_t rows_per_mod, float eps,
                      int layernorm, void* stream);

/* q_norm / k_norm (RMSNorm over the full width with weight) followed by interleaved-pair RoPE.
 * cos/sin: bf16 [rows, D] tables (NULL for attn2: no RoPE).  q and k rows are independent row sets.
 * Replaces attention.py:996-1012 (q_norm/k_norm) and apply_rotary_emb :917-932. */
int b200_qknorm_rope_fwd(const void* xq, int64_t ldq, const void* xk, int64_t ldk, const void* wq,
                         const void* wk, const void* cos_t, const void* sin_t, int64_t ldcs, void* oq,
                         int64_t ldoq, void* ok, int64_t ldok, int64_t rows_q, int64_t rows_k, int D,
                         float eps, void* stream);
/* prod_q / prod_k (bf16, may be NULL): (RoPE^T dq) * xhat_q and the same for k; their column sums are the gradients
 * of q_norm.weight / k_norm.weight when those train. */
int b200_qknorm_rope_bwd(const void* dq, int64_t lddq, int dq_is_f32, const void* dk, int64_t lddk,
                         int dk_is_f32, const void* xq, int64_t ldq, const void* xk, int64_t ldk,
                         const void* wq, const void* wk, const void* cos_t, const void* sin_t,
                         int64_t ldcs, void* oq, int64_t ldoq, void* ok, int64_t ldok, void* prod_q,
                         int64_t ldpq, void* prod_k, int64_t ldpk, int64_t rows_q, int64_t rows_k, int D,
                         float eps, void* stream);

/* Rectified flow: x_t = (1-t) x0 + t eps, v = eps - x0 (either output may be NULL); t fp32 [batch].
 * Replaces RectifiedFlowScheduler.add_noise / build_velocity_target, rf.py:376-386, 400-426. */
int b200_rf_noise(const void* x0, const void* noise, const float* t, void* xt, void* v, int64_t batch,
                  int64_t per_sample, void* stream);
/* loss = mean((out - target)^2); dout = grad_scale * 2 (out - target) / numel (dout may be NULL).
 * Replaces F.mse_loss + its backward, training.py:159-160, 203. */
int64_t b200_rf_loss_workspace_bytes(void);
int b200_rf_loss(const void* out, const void* target, void* dout, float* loss, int64_t numel,
                 float grad_scale, void* workspace, int64_t workspace_bytes, void* stream);

/* In-place conditioning lerp on tokens [B,N,C]: frame 0 <- lerp(tok, ref, w_ref), frames >= 1 <-
 * lerp(tok, pose, w_pose); ref [B,C,1,HW], pose [B,C,F,HW] with F*HW = N_total.  `tokens` may be a
 * contiguous shard [token_offset, token_offset + N) of the clip (sequence-sharded attn1).
 * Replaces transformer3d.py:447-466. */
int b200_lerp_condition(void* tokens, const void* ref, const void* pose, int B, int N, int C, int HW,
                        float w_ref, float w_pose, int token_offset, int N_total, void* stream);

/* Element-wise tail of one sampling step, fused: guidance combine of the model output v [conds*B, N, C] bf16 (order:
 * (uncond,) text (, perturbed)) -- CFG `u + gs (text - u)` with u = uncond or, cfg_star, uncond scaled by
 * <text,uncond>/(|uncond|^2 + 1e-8); STG `+ stg (text - perturbed)`; rescale `* (rs * std(text)/std(pred) + 1 - rs)`
 * with unbiased per-sample stds -- then the Euler update x <- x - dt * pred on the fp32 running latents x [B, N, C]
 * for the tokens with t - 1e-6 < noise_level[b, n] (all tokens when noise_level is NULL; noise_level = 1 -
 * conditioning_mask), and the bf16 model input of the next step written n_next times back to back into x_next
 * [n_next*B, N, C] (n_next may be 0).  dt: fp32 [1] or, dt_per_token, [N] (shared by the batch like the reference's
 * timestep[:1]); scalars: DEVICE fp32 {guidance_scale, stg_scale, rescaling_scale, t}, so a captured step can be
 * replayed with new values.  Workspace: b200_guidance_step_workspace_bytes(B), needed for cfg_star / rescale.
 * Replaces pipelines/pipeline_ltx_video.py:1217-1260 (guidance), :1346-1379 (denoising_step), rf.py:305-374 (Euler). */
int64_t b200_guidance_step_workspace_bytes(int B);
int b200_guidance_step(const void* v, float* x, void* x_next, int n_next, const float* dt, int dt_per_token,
                       const float* noise_level, const float* scalars, int B, int64_t N, int C, int has_cfg,
                       int has_stg, int cfg_star, int rescale, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* AdamW (decoupled weight decay) over a list of tensors in ONE launch.  `table`: DEVICE array of 48-byte entries
 *   { void* param; const void* grad; void* exp_avg; void* exp_avg_sq; int64_t numel; int32_t is_bf16; int32_t pad; }
 * (fp32 or bf16 tensors, state in the parameter's dtype, math in fp32); `block_map`: DEVICE int32 [n_blocks][2] =
 * (entry index, chunk index), one block per b200_adamw_chunk_elems() elements of an entry; `step`: DEVICE fp32 count
 * of completed steps, incremented by the launch; `hyper`: DEVICE fp32 {lr, beta1, beta2, eps, weight_decay};
 * `done_counter`: DEVICE int32, zero between launches.  Everything the update depends on is in device memory, so the
 * launch can be captured and replayed.  Replaces torch.optim.AdamW(params, lr).step(), training.py:206, 271. */
int b200_adamw_chunk_elems(void);
int b200_adamw_step(const void* table, const int32_t* block_map, int n_blocks, float* step, const float* hyper,
                    int32_t* done_counter, void* stream);

/* out[m,:] = x[m,:] * g[m / rows_per_mod,:]  (AdaLN gate applied to an incoming gradient). */
int b200_rowscale(const void* x, int64_t ldx, const void* g, int64_t gstride, void* out, int64_t ldo,
                  int64_t rows, int D, int64_t rows_per_mod, void* stream);
/* out[n] = sum_m x[m,n]  (bias gradients of the trainable caption projection). */
int b200_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int N, void* stream);
/* out[g, n] = sum over the rows of group g (rows_per_group consecutive rows) of a[r, n] * (b ? b[r, n] : 1), fp32
 * (the call zeroes `out` and reduces 16-row chunks into it at the L2; workspace arguments reserved, may be NULL / 0).
 * The gradients of whatever is broadcast over tokens when the reference's "full" strategy trains it
 * (training.py:75-91): AdaLN shift / scale / gate per sample, q_norm / k_norm weights, projection biases. */
int64_t b200_colsum_groups_workspace_bytes(int64_t rows, int N, int64_t rows_per_group);
int b200_colsum_groups(const void* a, int64_t lda, const void* b, int64_t ldb, float* out, int64_t rows, int N,
                       int64_t rows_per_group, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LTX_H_ */
