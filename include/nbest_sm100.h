/*
 * nbest_sm100.h — C ABI of libnbest_sm100.so: the B200 (sm_100a) hot path of N-Best-ASR-Transformer.
 *
 * The reference (/root/reference) is pure Python/PyTorch and has no FFI of its own; its "plugin API" for this
 * path is the Python call surface of models/model.py, models/modules/hierarchical_classifier.py,
 * models/optimization.py, utils/bert_xlnet_inputs.py and the loss code in n_best_asr_bert.py. Each entry point
 * below names the reference lines whose device work it replaces; the Python drop-ins in
 * n-best-asr-transformer_b200/ bind these symbols with ctypes.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.
 *   - every data pointer is a DEVICE pointer owned by the caller; the library never allocates or frees device
 *     memory (one exception: a 536-byte buffer owned by the context — the GEMM scheduler ring, see
 *     nbest_ctx_set_gemm_dynamic, and the per-step state record, see nbest_ctx_set_step_state). `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it, no internal sync.
 *   - "bf16" buffers are raw uint16 bfloat16, row-major; "f32" are float.
 *   - return value: NBEST_OK (0) or a negative nbest_status; nbest_last_error(ctx) gives the message.
 *   - one nbest_ctx per process / GPU rank; calls on one ctx must come from one thread at a time.
 */
#ifndef NBEST_SM100_H_
#define NBEST_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBEST_ABI_VERSION 3

typedef struct nbest_ctx nbest_ctx;

typedef enum {
  NBEST_OK = 0,
  NBEST_EINVAL = -1,       /* bad argument (shape not supported, null pointer, ...) */
  NBEST_ECUDA = -2,        /* a CUDA runtime / driver call failed */
  NBEST_EUNSUPPORTED = -3  /* device is not sm_100 */
} nbest_status;

/* ---- context ------------------------------------------------------------------------------------------- */
int nbest_abi_version(void);
/* Replaces the device pick of n_best_asr_bert.py:116-126 / utils/gpu_selection.py:27-66 for one rank. */
int nbest_ctx_create(nbest_ctx** out, int device);
void nbest_ctx_destroy(nbest_ctx* ctx);
const char* nbest_last_error(nbest_ctx* ctx);
/* Number of kernels launched through ctx so far (bench.py's gpu_launches). */
uint64_t nbest_launch_count(nbest_ctx* ctx);
/* The persistent GEMMs size their grids to (SM count - n_sms): the data-parallel trainer reserves SMs while a gradient
 * all-reduce is in flight so that NCCL's CTAs and the statically scheduled CTA pairs do not wait for each other
 * (replaces nothing in the reference: it has no multi-GPU path, utils/gpu_selection.py:27-66 picks ONE device). */
int nbest_ctx_set_sm_reserve(nbest_ctx* ctx, int n_sms);
/* on != 0 (the data-parallel trainer's setting for world size > 1; NBEST_GEMM_DYNAMIC=0/1 forces it): the persistent
 * GEMM CTAs draw their tiles from a global counter instead of a static round-robin share, so a CTA pair whose SMs are
 * held by a concurrent collective takes fewer tiles rather than delaying its share (measured on B200: +1.7 % step time
 * on one GPU, where nothing competes for SMs — hence off there — and dgrad GEMMs 6 % faster under NCCL overlap). Needs the context's 512-byte device scheduler buffer — the only device memory the
 * library allocates itself (nbest_ctx_create / nbest_ctx_destroy). */
int nbest_ctx_set_gemm_dynamic(nbest_ctx* ctx, int on);
/* Per-step state for CUDA-graph replay of the training step (replaces nothing in the reference, whose step is ~3,300
 * eager launches, n_best_asr_bert.py:254-277). A captured graph bakes every by-value argument in — the dropout seeds and
 * the optimizer's schedule multiplier among them — so what must change from one replay to the next is read from a
 * 24-byte device record owned by the context:
 *   salt         XORed into the (host-mixed) seed of EVERY kernel that draws dropout masks (embedding, GEMM dropout
 *                epilogue, LayerNorm backward, the attention kernels, the classifier head). 0 — the state after
 *                nbest_ctx_create — leaves the by-value seeds as they are.
 *   sched, inv_bc1, inv_sqrt_bc2   BertAdam's / AdamW's schedule multiplier (models/optimization.py:289-293) and
 *                torch.optim.Adam's bias corrections of THIS step; read by nbest_adam_step / nbest_bertadam_step in place
 *                of their by-value arguments while nbest_ctx_set_step_indirect(ctx, 1) is in force.
 * nbest_ctx_set_step_state enqueues a one-thread kernel on `stream` that writes the record: call it ahead of the graph
 * launch (stream order makes it visible to the replayed kernels, which read it after their grid dependency wait). */
int nbest_ctx_set_step_state(nbest_ctx* ctx, uint32_t salt, double sched, float inv_bc1, float inv_sqrt_bc2, void* stream);
int nbest_ctx_set_step_indirect(nbest_ctx* ctx, int on);
/* TMA descriptor cache hits so far (host-overhead diagnostics). */
uint64_t nbest_tmap_cache_hits(nbest_ctx* ctx);

/* ---- A2: packed variable-length batch layout ------------------------------------------------------------- */
/* Replaces the padded tensors of utils/bert_xlnet_inputs.py:91-102 plus `attention_mask = input_ids > 0`
 * (models/model.py:43,45,54,56) by a packed layout built on the GPU.
 *   ids      [B,S] int64, right padded (pad id 0 for BERT, 1 for XLM-R);  seg_ids [B,S] int64 or NULL
 *   pos_mode 0: BERT   pos = index inside the sequence
 *            1: XLM-R  pos = cumsum(ids != 1) * (ids != 1) + 1   (modeling_xlm_roberta.py:147-160)
 * A row's length is (index of its last id > 0) + 1, so BERT rows lose their padding and XLM-R rows keep the
 * reference's <pad>=1 tokens (they are attendable in the reference). key_valid[t] = ids > 0.
 * Outputs: lens[B], cu_seqlens[B+1], and per packed token tokens/pos/seq_of (int32) and seg/key_valid (uint8),
 * each with capacity B*S. total_tokens (device int32) receives T = cu_seqlens[B]. */
int nbest_pack_batch(nbest_ctx* ctx, const int64_t* ids, const int64_t* seg_ids, int B, int S, int pos_mode,
                     int32_t* lens, int32_t* cu_seqlens, int32_t* tokens, uint8_t* seg, int32_t* pos,
                     int32_t* seq_of, uint8_t* key_valid, void* stream);
/* Hypothesis-id map of the packed layout (north_star (1): "segment/hypothesis-id map built on the GPU"). The n-best
 * list is joined as `[CLS] sys [SEP] hyp1 [SEP] hyp2 ... hypN [SEP]` (utils/bert_xlnet_inputs.py:75-85); hyp_id[t] =
 * number of separator tokens (sep_id: 102 for BERT, 2 for XLM-R) at earlier positions of t's sequence, i.e. 0 for
 * [CLS] + system turn + the first [SEP], k for the tokens of hypothesis k and the [SEP] that closes it (saturates at
 * 255). tokens / cu_seqlens are nbest_pack_batch outputs. */
int nbest_pack_hyp_ids(nbest_ctx* ctx, const int32_t* tokens, const int32_t* cu_seqlens, int B, int sep_id,
                       uint8_t* hyp_id, void* stream);
/* Both encoder streams of one step (ASR n-best sequences first, then the transcripts: n_best_asr_bert.py:249-250 builds
 * the two padded batches, models/model.py:43-58 runs the encoder on each) packed into ONE batch in place: lens /
 * cu_seqlens over B_a + B_t sequences, token arrays with capacity B_a*S_a + B_t*S_t. Same per-token contract as
 * nbest_pack_batch; seg_a / seg_t may be NULL. */
int nbest_pack_batch_dual(nbest_ctx* ctx, const int64_t* ids_a, const int64_t* seg_a, int B_a, int S_a, const int64_t* ids_t,
                          const int64_t* seg_t, int B_t, int S_t, int pos_mode, int32_t* lens, int32_t* cu_seqlens,
                          int32_t* tokens, uint8_t* seg, int32_t* pos, int32_t* seq_of, uint8_t* key_valid, void* stream);
/* Row gather / scatter of [*, 768] bf16 activations by packed row index — the [CLS] slice `sequence_output[:, 0, :]` of
 * models/model.py:46-47,57-58 and its autograd scatter. gather: dst[i] = src[row_idx[i]] (fp32 dst if dst_is_f32);
 * scatter: dst [T, 768] = 0 except dst[row_idx[i]] = src[i]. nbest_zero: asynchronous zero-fill (gradient buffers). */
int nbest_rows_gather(nbest_ctx* ctx, const void* src_bf16, const int32_t* row_idx, int n, int hidden, void* dst,
                      int dst_is_f32, void* stream);
int nbest_rows_scatter(nbest_ctx* ctx, const void* src_bf16, const int32_t* row_idx, int n, int T, int hidden,
                       void* dst_bf16, void* stream);
int nbest_zero(nbest_ctx* ctx, void* ptr, int64_t nbytes, void* stream);
/* Row-sparse exchange of the word-embedding gradient in the data-parallel trainer (the reference has no multi-GPU path;
 * its dense embedding gradient comes from autograd, n_best_asr_bert.py:264 — of XLM-R's 250,002 x 768 table at most T
 * rows are non-zero per step). phase 0: flags[n_rows] = 0, then flags[tokens[t]] = 1 for t < T; (the caller MAX-reduces
 * flags over the ranks;) phase 1: rows[0..count) = ascending row indices with a set flag, except skip_row (padding_idx:
 * its gradient row is zero by construction). nbest_rows_move_f32: gather (dst[i] = src[rows[i]]) or scatter
 * (dst[rows[i]] = src[i]) of fp32 [*, 768] rows. */
int nbest_rows_touched(nbest_ctx* ctx, const int32_t* tokens, int T, int n_rows, int skip_row, int32_t* flags, int phase,
                       int32_t* rows, int32_t* count, void* stream);
int nbest_rows_move_f32(nbest_ctx* ctx, const float* src, const int32_t* rows, int n, int hidden, float* dst, int scatter,
                        void* stream);

/* ---- K1: embedding gather + LayerNorm (+dropout) ---------------------------------------------------------- */
/* BertEmbeddings.forward (transformers/models/bert/modeling_bert.py:102-112), called from models/model.py:43-45.
 * y = dropout(LN(word[tok] + posemb[pos] + type[seg])), bf16 out, fp32 statistics kept for backward. */
int nbest_embed_ln_fwd(nbest_ctx* ctx, const int32_t* tokens, const uint8_t* seg, const int32_t* pos, int T,
                       const float* word, const float* posemb, const float* type, const float* gamma,
                       const float* beta, float eps, int hidden, void* y_bf16, float* mean, float* rstd,
                       float p_drop, uint32_t seed, void* stream);
/* Backward of the above (autograd of n_best_asr_bert.py:264): scatter-adds into the fp32 table gradients;
 * rows word_pad_row / pos_pad_row (nn.Embedding padding_idx; -1 = none) receive no gradient. */
int nbest_embed_ln_bwd(nbest_ctx* ctx, const int32_t* tokens, const uint8_t* seg, const int32_t* pos, int T,
                       const float* word, const float* posemb, const float* type, const float* gamma,
                       const float* mean, const float* rstd, int hidden, const void* dy_bf16, float p_drop,
                       uint32_t seed, float* dword, float* dpos, float* dtype, float* dgamma, float* dbeta,
                       int word_pad_row, int pos_pad_row, void* stream);

/* ---- LayerNorm over a packed [T,hidden] bf16 tensor -------------------------------------------------------- */
/* BertSelfOutput / BertOutput LayerNorm (modeling_bert.py:294-298, 352-356). x = residual + dropout(dense), already
 * summed by the GEMM epilogue. */
int nbest_ln_fwd(nbest_ctx* ctx, const void* x_bf16, const float* gamma, const float* beta, float eps, int T,
                 int hidden, void* y_bf16, float* mean, float* rstd, void* stream);
/* The same LayerNorm with the row statistics supplied as partial sums: row_partials [T][n_partials] float2 {sum, sum of
 * squares} over disjoint column groups of the row (what nbest_gemm_bf16 writes to out2 under NBEST_EPI_BIAS_DROP_RES,
 * n_partials = N / 64 <= 16); NULL = compute them here. One pass over the row instead of two. */
int nbest_ln_fwd_stats(nbest_ctx* ctx, const void* x_bf16, const float* gamma, const float* beta, float eps, int T,
                       int hidden, const float* row_partials, int n_partials, void* y_bf16, float* mean, float* rstd,
                       void* stream);
/* dx = LN'(dy); dgamma/dbeta accumulated (+=). If dx_masked != NULL it receives dx * dropout_mask / (1-p) (the
 * gradient entering the preceding dense layer, whose output was dropped with (p_drop, seed)); dbias (+=, may be
 * NULL) gets the column sum of that tensor = gradient of the preceding dense bias. */
int nbest_ln_bwd(nbest_ctx* ctx, const void* dy_bf16, const void* x_bf16, const float* mean, const float* rstd,
                 const float* gamma, int T, int hidden, void* dx_bf16, void* dx_masked_bf16, float p_drop,
                 uint32_t seed, float* dgamma, float* dbeta, float* dbias, void* stream);

/* out[n] += sum_t x[t,n] for a bf16 [T,N] tensor (bias gradients of the QKV and FFN-in projections). */
int nbest_colsum_bf16(nbest_ctx* ctx, const void* x_bf16, int T, int N, float* out, void* stream);
/* dst_bf16[i] = bf16(src[i]) — refresh of the bf16 working copy of fp32 master weights. */
int nbest_cast_f32_bf16(nbest_ctx* ctx, const float* src, void* dst_bf16, int64_t n, void* stream);

/* ---- K2/K4/K5/K6: tcgen05 / TMEM bf16 GEMM, TMA fed ------------------------------------------------------- */
/* The nn.Linear calls of BertSelfAttention (modeling_bert.py:179-181), BertSelfOutput (:294-298),
 * BertIntermediate (:339-342) and BertOutput (:352-356), forward and both backward products.
 *   C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
 *   a_mn_major = 0: A stored [M,K] (k contiguous);  1: A stored [K,M] (m contiguous)
 *   b_mn_major = 0: B stored [N,K] (nn.Linear weight, k contiguous); 1: B stored [K,N] (n contiguous)
 * forward  y = x W^T      : (0,0), A = x [T,in],   B = W [out,in]
 * dgrad    dx = dy W      : (0,1), A = dy [T,out], B = W [out,in]  read as [K=out, N=in]
 * wgrad    dW += dy^T x   : (1,1), A = dy [T,out] read as [K=T, M=out], B = x [T,in]; NBEST_EPI_ACCUM_F32, split over T
 * Shapes: N % 128 == 0 (256-wide tiles are used when N % 256 == 0), K % 64 == 0 for K-major operands (any K for
 * the (1,1) form), any M for the (0,x) forms, M % 128 == 0 for (1,1). lda/ldb/ldc are row pitches in elements. */
typedef enum {
  NBEST_EPI_NONE = 0,        /* C = acc                                                    (bf16 out) */
  NBEST_EPI_BIAS = 1,        /* C = acc + bias[n]                                                     */
  NBEST_EPI_BIAS_GELU = 2,   /* u = acc + bias; C = gelu_erf(u); out2 = gelu_erf'(u) (if out2 != NULL): the
                              * derivative is saved instead of u, so that the backward epilogue is one multiply */
  NBEST_EPI_BIAS_DROP_RES = 3, /* C = dropout(acc + bias[n]; p_drop, seed) + aux[m,n]; if out2 != NULL: out2 (fp32
                              * [M][N/64][2]) = {sum, sum of squares} of C[m, 64u .. 64u+63]: the partial statistics
                              * of the residual LayerNorm that follows (nbest_ln_fwd_stats)            */
  NBEST_EPI_DGELU = 4,       /* C = acc * aux[m,n], aux = gelu_erf'(u) saved by NBEST_EPI_BIAS_GELU; if out2 != NULL:
                              * out2 (fp32 [N]) += sum_m C[m,n]
                              * (the bias gradient of the layer that produced aux, fused)              */
  NBEST_EPI_ADD = 5,         /* C = acc + aux[m,n]                                                    */
  NBEST_EPI_ACCUM_F32 = 6,   /* C (fp32) += acc     (wgrad; atomic accumulation, split-K)             */
  NBEST_EPI_DELTA = 7        /* C = acc; out2 (fp32 [N/64][M]) = per-64-column dot(acc[m,:], aux[m,:]): the
                              * attention backward's delta = rowsum(dO * O) per head, produced by the out-projection
                              * dgrad that computes dO (flash-attention backward preprocess, fused)             */
} nbest_epilogue;

int nbest_gemm_bf16(nbest_ctx* ctx, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                    int b_mn_major, void* C, int64_t ldc, int M, int N, int K, int epilogue, const float* bias,
                    const void* aux_bf16, int64_t ldaux, void* out2_bf16, float p_drop, uint32_t seed,
                    void* stream);

/* ---- K3: fused masked self-attention over packed sequences ------------------------------------------------ */
/* BertSelfAttention scores/softmax/context (modeling_bert.py:115-140,192-205) with the key mask of
 * models/model.py:43. qkv [T, 3*heads*64] bf16 (q | k | v, head-major inside each third), out [T, heads*64] bf16,
 * lse [heads, T] fp32 (natural-log-sum-exp of the scaled scores, kept for backward). head_dim is 64.
 * key_valid may be NULL (all keys valid). Attention-probability dropout uses (p_drop, seed). */
int nbest_attn_varlen_fwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid,
                          int B, int max_len, int heads, int T, void* out_bf16, float* lse, float p_drop,
                          uint32_t seed, void* stream);
/* Backward: dqkv [T_active, 3*heads*64] bf16 from dout [T_active, heads*64]; delta_ws [heads, T] fp32 scratch.
 * T is the forward's token count (stride of lse / delta_ws and part of the dropout index); the B sequences given
 * here cover the first T_active <= T tokens (the gradient-carrying prefix: the ASR stream when the transcript
 * stream is forward-only, n_best_asr_bert.py:166). */
/* out_bf16 may be NULL: delta_ws then already holds delta = rowsum(dO * O) per head with row pitch T_active (written by
 * the out-projection dgrad GEMM with NBEST_EPI_DELTA) and the internal preprocess kernel is skipped. */
int nbest_attn_varlen_bwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid,
                          int B, int max_len, int heads, int T, int T_active, const void* out_bf16, const void* dout_bf16,
                          const float* lse, void* dqkv_bf16, float* delta_ws, float p_drop, uint32_t seed,
                          void* stream);

/* ---- attention on the tcgen05 tensor path for packed SHORT sequences (<= 128 tokens) ----------------------- */
/* Same math and same dropout index map as nbest_attn_varlen_fwd/bwd (modeling_bert.py:115-140,192-205 under the key
 * mask of models/model.py:43,45), organised for Blackwell: whole sequences are packed into 128-row tiles of the
 * packed token axis, one work item = (tile, head) computes S = Q K^T as ONE 128x128x64 tcgen05.mma chain with a
 * block-diagonal (same-sequence) mask, softmax / dropout out of TMEM, O = P V as a 128x64x128 chain; Q/K/V/dO tiles
 * arrive by TMA, persistent CTAs. Sequences longer than 128 tokens are NOT touched: run nbest_attn_varlen_fwd2/bwd2
 * with min_len = 129 for them (counts[2] tells whether there are any).
 *
 * nbest_attn_plan: tiles[2*i] = first packed row, tiles[2*i+1] = rows of tile i (capacity B tiles); counts[0] = number
 * of tiles, counts[1] = number of tiles covering sequences [0, break_at) (no tile straddles break_at: the backward of
 * the ASR prefix uses count_idx = 1), counts[2] = number of sequences longer than 128; row_bounds[2*t], [2*t+1] =
 * first / one-past-last packed row of token t's sequence. cu_seqlens / seq_of are nbest_pack_batch outputs. */
int nbest_attn_plan(nbest_ctx* ctx, const int32_t* cu_seqlens, const int32_t* seq_of, int B, int T, int break_at,
                    int32_t* tiles, int32_t* counts, int32_t* row_bounds, void* stream);
/* max_tiles: host-side upper bound of counts[count_idx] (B is always one); sizes the persistent grid. */
int nbest_attn_tiles_fwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* tiles, const int32_t* counts, int count_idx,
                         int max_tiles, const int32_t* row_bounds, const uint8_t* key_valid, int heads, int T,
                         void* out_bf16, float* lse, float p_drop, uint32_t seed, void* stream);
/* delta[h * delta_pitch + t] = rowsum(dO * O) per head (the out-projection dgrad's NBEST_EPI_DELTA epilogue writes
 * it). dout is [T_active, heads*64]; gradients are written for the tiles' own rows only. */
int nbest_attn_tiles_bwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* tiles, const int32_t* counts, int count_idx,
                         int max_tiles, const int32_t* row_bounds, const uint8_t* key_valid, int heads, int T, int T_active,
                         const void* dout_bf16, const float* lse, const float* delta, int delta_pitch, void* dqkv_bf16,
                         float p_drop, uint32_t seed, void* stream);
/* nbest_attn_varlen_fwd / _bwd restricted to sequences of at least min_len tokens (0 = all). */
int nbest_attn_varlen_fwd2(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid,
                           int B, int max_len, int heads, int T, void* out_bf16, float* lse, float p_drop, uint32_t seed,
                           int min_len, void* stream);
int nbest_attn_varlen_bwd2(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid,
                           int B, int max_len, int heads, int T, int T_active, const void* out_bf16,
                           const void* dout_bf16, const float* lse, void* dqkv_bf16, float* delta_ws, float p_drop,
                           uint32_t seed, int min_len, void* stream);

/* Last encoder layer: only the [CLS] row of each sequence is consumed downstream (models/model.py:46-47,58), so its
 * attention is evaluated for that single query row. out_cls [B, heads*64] bf16, lse_cls [heads, B] fp32. Same
 * arithmetic and the same dropout-mask indices as nbest_attn_varlen_fwd would use for row cu_seqlens[b]. max_len <= 512. */
int nbest_attn_cls_fwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid, int B,
                       int max_len, int heads, int T, void* out_cls_bf16, float* lse_cls, float p_drop, uint32_t seed,
                       void* stream);
/* Backward of the above for the first B sequences: writes ALL rows of dqkv [cu_seqlens[B], 3*heads*64] that belong to
 * them (dQ = 0 except at the CLS rows). lse_cls has row pitch lse_stride (the forward's B). */
int nbest_attn_cls_bwd(nbest_ctx* ctx, const void* qkv_bf16, const int32_t* cu_seqlens, const uint8_t* key_valid, int B,
                       int max_len, int heads, int T, const void* out_cls_bf16, const void* dout_cls_bf16,
                       const float* lse_cls, int lse_stride, void* dqkv_bf16, float p_drop, uint32_t seed, void* stream);

/* ---- K8/K9/K10: CLS gather + hierarchical STC head + losses ----------------------------------------------- */
/* Label hierarchy (memory['top2bottom_dict'], n_best_asr_bert.py:489-496) flattened by the host into int32 tables:
 *   n_top (30), n_bottom (161), n_groups (10 multi-way groups), n_cols = n_top + sum(n_k) (171)
 *   col_group[n_cols]  : 0 for the n_top act-slot columns, g (1..n_groups) for a column of group g
 *   col_bottom[n_cols] : bottom-label id scored by the column; for an act-slot column: its single bottom id, or -1
 *                        when the act-slot owns a multi-way group
 *   grp_off[n_groups+1]: first column of group g is grp_off[g-1]... stored as offsets into [0,n_cols) for g=1..;
 *                        entry 0 is n_top, entry n_groups is n_cols
 *   grp_top[n_groups]  : act-slot (top) index that owns group g */
typedef struct {
  int32_t n_top, n_bottom, n_groups, n_cols;
  const int32_t* col_group;
  const int32_t* col_bottom;
  const int32_t* grp_off;
  const int32_t* grp_top;
} nbest_hierarchy;

/* models/model.py:46-47 (CLS row) + HierarchicalClassifier.forward (hierarchical_classifier.py:35-60).
 * x [T,hidden] bf16 last-layer output, cls rows at cu_seqlens[i]. W [n_cols,hidden], bias [n_cols] fp32 (top rows
 * first, then the groups in ascending top id). Dropout: 11 independent masks of the feature (one per Linear call).
 * Outputs (fp32): cls [B,hidden], logits [B,n_cols], top_scores [B,n_top], bottom_scores [B,n_cols-n_top],
 * final_scores [B,n_bottom]; decode [B,n_bottom] uint8 = pred_one_sample (n_best_asr_bert.py:198-215) as a bitmap
 * (none_col_mask[n_cols] marks the columns whose label ends with NONE). */
int nbest_stc_head_fwd(nbest_ctx* ctx, const void* x_bf16, const int32_t* cu_seqlens, int B, int hidden,
                       const float* W, const float* bias, const nbest_hierarchy* h, const uint8_t* none_col_mask,
                       float p_drop, uint32_t seed, float* cls, float* logits, float* top_scores,
                       float* bottom_scores, float* final_scores, uint8_t* decode, void* stream);
/* cal_total_loss + cal_ce_loss (n_best_asr_bert.py:145-195, utils/STC_util.py:4-51) and their gradient.
 * labels [B,n_bottom] fp32 multi-hot. losses[4] (+=) = {mse, bce_final, bce_top, ce}, all sum-reduced as in the
 * reference (ce carries its 1/n_groups). dlogits [B,n_cols] = d(total)/d(logits). If asr_cls && trans_cls:
 * mse = mean((asr-trans)^2) * mse_scale and d_asr_cls / d_trans_cls [B,hidden] (=) its gradients. */
int nbest_stc_loss_fwd_bwd(nbest_ctx* ctx, const float* logits, const float* labels, int B, const nbest_hierarchy* h,
                           const float* asr_cls, const float* trans_cls, int hidden, float mse_scale, float* losses,
                           float* dlogits, float* d_asr_cls, float* d_trans_cls, void* stream);
/* Generic backward through sigmoid / group softmax / product for the autograd drop-in (when the reference's own
 * cal_total_loss is used): d_top [B,n_top], d_bottom [B,n_cols-n_top], d_final [B,n_bottom] (any may be NULL)
 * -> dlogits [B,n_cols]. */
int nbest_stc_scores_bwd(nbest_ctx* ctx, const float* top_scores, const float* bottom_scores, const float* d_top,
                         const float* d_bottom, const float* d_final, int B, const nbest_hierarchy* h, float* dlogits,
                         void* stream);
/* Backward of the 11 Linear calls: dW [n_cols,hidden] +=, dbias [n_cols] +=, dcls [B,hidden] (+= if accumulate)
 * from dlogits and the (re-generated) dropout masks. */
int nbest_stc_head_bwd(nbest_ctx* ctx, const float* dlogits, const float* cls, const float* W, int B, int hidden,
                       const nbest_hierarchy* h, float p_drop, uint32_t seed, float* dW, float* dbias, float* dcls,
                       int accumulate_dcls, void* stream);
/* dx[T,hidden] bf16 = 0 except row cu_seqlens[i] = dcls[i] * scale (autograd of the [:,0,:] slice, model.py:47). */
int nbest_cls_scatter(nbest_ctx* ctx, const float* dcls, const int32_t* cu_seqlens, int B, int T, int hidden,
                      void* dx_bf16, void* stream);

/* ---- K10b: epoch metrics on the device ------------------------------------------------------------------------
 * Replaces the per-sample host loop of train_epoch / eval_epoch (n_best_asr_bert.py:283-288, 335-350) that calls
 * pred_one_sample (:198-215), filter_informative (:218-229), update_f1 (utils/fscore.py:2-11) and compares label sets.
 * decode [B,n_bottom] uint8 is the head kernel's prediction bitmap, labels [B,n_bottom] fp32 the gold multi-hot of
 * collate_fn (utils/dataset/tod_asr_util.py:118-126; labels missing from label2idx sit in the <unk> column 1 and so
 * count as false negatives, as their strings do in the reference). col_mask [n_bottom] uint8 (nullable) keeps the
 * "informative" columns of the ontology filter (applied to predictions and gold alike, :338-340).
 * counters[4] (int64, accumulated, never reset here) += {TP, FP, FN, utterances whose label sets match exactly};
 * counters[4] is only read by the host once per epoch. */
int nbest_stc_metrics(nbest_ctx* ctx, const uint8_t* decode, const float* labels, const uint8_t* col_mask, int B,
                      int n_bottom, long long* counters, void* stream);

/* ---- K11: fused multi-tensor BertAdam ---------------------------------------------------------------------- */
/* BertAdam.step (models/optimization.py:237-302) with WarmupLinearSchedule (:162-171) and per-tensor
 * clip_grad_norm_ (:270-271) over one flat fp32 parameter buffer. seg[] (device) describes the tensors. */
typedef struct {
  int64_t offset;   /* first element in the flat buffers */
  int64_t numel;
  double lr;        /* the tensor's own param-group lr (n_best_asr_bert.py:548-549) */
  float weight_decay;
  int32_t active;   /* 0: grad is None in the reference (pooler) -> skipped entirely */
} nbest_adam_tensor;
/* chunks[] (device, int32 triples {tensor, start, len}) tile the active tensors; start is relative to the tensor.
 * norms_ws[n_tensors + n_chunks] fp32 scratch (per-tensor sums of squares, then the per-chunk partials they are folded
 * from in a fixed order: replicas fed identical gradients stay bit-identical). sched = schedule.get_lr(step) computed by
 * the host in double.
 * p_bf16 (nullable) receives the refreshed bf16 working copy of p. grad_scale multiplies g before use (1/R etc). */
int nbest_bertadam_step(nbest_ctx* ctx, float* p, const float* g, float* m, float* v, void* p_bf16,
                        const nbest_adam_tensor* tensors, int n_tensors, const int32_t* chunks, int n_chunks,
                        float* norms_ws, double sched, float b1, float b2, float eps, float max_grad_norm,
                        void* stream);

/* The other two optimizers n_best_asr_bert.py:553-569 can select, over the same flat buffers and tensor table:
 *   NBEST_ADAM_HF_ADAMW  transformers(2.3.0).AdamW(correct_bias=False) (:563): p -= lr m/(sqrt(v)+eps); p -= lr wd p
 *   NBEST_ADAM_TORCH     torch.optim.Adam (:554): L2-coupled decay, bias correction with the 1-based `step`
 * global_clip != 0 replaces BertAdam's per-tensor clipping by torch.nn.utils.clip_grad_norm_(all params, max_grad_norm)
 * (n_best_asr_bert.py:268-271): one coefficient from the 2-norm over every active tensor. sched multiplies each
 * tensor's lr (get_linear_schedule_with_warmup's lambda for AdamW, :564-568; 1 for Adam). */
typedef enum { NBEST_ADAM_BERT = 0, NBEST_ADAM_HF_ADAMW = 1, NBEST_ADAM_TORCH = 2 } nbest_adam_mode;
int nbest_adam_step(nbest_ctx* ctx, int mode, float* p, const float* g, float* m, float* v, void* p_bf16,
                    const nbest_adam_tensor* tensors, int n_tensors, const int32_t* chunks, int n_chunks,
                    float* norms_ws, double sched, float b1, float b2, float eps, float max_grad_norm, int global_clip,
                    int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NBEST_SM100_H_ */
