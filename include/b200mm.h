/* b200mm C-ABI — the drop-in boundary below the Python host.
 *
 * Every entry point takes raw DEVICE pointers, int64 sizes, scalar hyper-parameters and a CUDA stream
 * (passed as void* so the header needs no CUDA include) and returns 0 on success or a negative
 * B200MM_ERR_* code; b200mm_last_error() returns the thread-local message of the last failure.
 * Ownership: the caller (torch) owns every buffer; kernels never allocate, free or synchronise.
 * All calls are stateless, re-entrant and stream-ordered, so they are safe from autograd worker threads.
 * There is NO CPU fallback: host pointers are rejected by the CUDA runtime, not silently handled.
 *
 * Each function cites the reference code (relative to the AntMMF tree) whose arithmetic it replaces.
 * bf16 = device buffers of __nv_bfloat16; f32 = float.
 */
#ifndef B200MM_H_
#define B200MM_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MM_OK 0
#define B200MM_ERR_SHAPE (-1)  /* unsupported / inconsistent dimensions */
#define B200MM_ERR_ALIGN (-2)  /* pointer or pitch not aligned as required */
#define B200MM_ERR_ARCH (-3)   /* device is not sm_100 */
#define B200MM_ERR_LAUNCH (-4) /* CUDA launch / driver failure (message holds cudaGetErrorString) */

/* activation selectors (GEMM epilogues, elementwise kernels) */
#define B200MM_ACT_NONE 0
#define B200MM_ACT_QUICKGELU 1 /* x*sigmoid(1.702x): antmmf/modules/vision/backbone/clip/model.py:222-224 */
#define B200MM_ACT_GELU_ERF 2  /* x*0.5*(1+erf(x/sqrt2)): antmmf/modules/vision/backbone/clip/modeling_bert.py:31-37 */

const char* b200mm_last_error(void);
int b200mm_version(void);
/* 0 if the current device is compute capability 10.x, else B200MM_ERR_ARCH */
int b200mm_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 + TMA):  D[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
 *
 * Replaces every nn.Linear / F.linear / torch.matmul on the path:
 *   ViT  in_proj / out_proj / c_fc / c_proj   antmmf/modules/vision/backbone/clip/model.py:231,236-238,251
 *   BERT query/key/value/dense                antmmf/modules/vision/backbone/clip/modeling_bert.py:120-122,135-137,182,221,234
 *   projections  x @ proj                     clip/model.py:332-333, clip/cn_model.py:210
 * and their autograd counterparts (dgrad: A=dY, B=W^T-as-MN-major; wgrad: both operands MN-major).
 *
 * Operand storage (bf16, 16-byte aligned base, pitches multiples of 8 elements):
 *   a_mn == 0: A is row-major [M, K] with row pitch lda      (K-major)
 *   a_mn == 1: A is row-major [K, M] with row pitch lda      (MN-major: the reduction index is the slow one)
 *   b_mn == 0: B is row-major [N, K] with row pitch ldb      (nn.Linear.weight layout)
 *   b_mn == 1: B is row-major [K, N] with row pitch ldb
 * Epilogue, per element, in this order (fp32):
 *   v = alpha*acc; v += bias[n] (bf16, optional); if aux_out (without dact_in): aux_out[m,n] = v (bf16, pre-activation);
 *   v = act(v); if dact_in: v *= act'(dact_in[m,n]) (act is then applied as derivative only, not to v), and with aux_out
 *   also aux_out[m,n] = act(dact_in[m,n]) (bf16: the activation recomputed for the weight gradient that follows);
 *   v += residual[m,n] (bf16, optional; not together with dact_in); D[m,n] = v (bf16 if d_f32 == 0 else f32)
 * splits > 1 partitions K over `splits` CTAs per tile; partial sums go to `workspace`
 * (f32, >= b200mm_gemm_workspace_bytes) and a second kernel reduces them and applies the epilogue.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* A;
  int64_t lda;
  int32_t a_mn;
  const void* B;
  int64_t ldb;
  int32_t b_mn;
  void* D;
  int64_t ldd;
  int32_t d_f32;
  int64_t M, N, K;
  float alpha;
  const void* bias;     /* bf16 [N] or NULL */
  int32_t act;          /* B200MM_ACT_* */
  void* aux_out;        /* bf16 [M,N] pitch ldd: pre-activation copy (act(dact_in) when dact_in is set), or NULL */
  const void* dact_in;  /* bf16 [M,N] pitch ld_dact: multiply by act'(dact_in) instead of applying act */
  int64_t ld_dact;
  const void* residual; /* bf16 [M,N] pitch ldr or NULL */
  int64_t ldr;
  int32_t splits;       /* >= 1 */
  void* workspace;      /* f32, needed iff splits > 1 */
  int64_t workspace_bytes;
  /* fused dropout on act(alpha*acc + bias), BEFORE the residual add: the BertSelfOutput / BertOutput pattern
   * LayerNorm(dropout(dense(x)) + input), clip/modeling_bert.py:176-184,228-236. Element (m, n) is kept iff
   * b200mm_dropout's rule keeps (row m, column n) under the same seed, so backward masks dY with b200mm_dropout. 0 = off. */
  float drop_p;
  uint64_t drop_seed;
} b200mm_gemm_args;

int64_t b200mm_gemm_workspace_bytes(int64_t M, int64_t N, int32_t splits);
int b200mm_gemm_bf16(const b200mm_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm (HBM-bound, fp32 statistics).  Replaces LayerNorm.forward (fp32 compute, eps 1e-5)
 * antmmf/modules/vision/backbone/clip/model.py:213-219 and nn.LayerNorm (eps 1e-12) in
 * clip/modeling_bert.py:83,179,231 plus their autograd.
 *   fwd: s = x[row] + add0[row % add_period] + (row % add_period == 0 ? add1 : 0);  y = LN(s)*w + b
 *        add0/add1 fuse the ViT stem "+ positional_embedding, class_embedding on token 0" (clip/model.py:313-324).
 *        s_out (optional) receives s; mean/rstd [rows] f32 are saved for backward.
 *   bwd: dx = LN'(dy) (+ dadd);  dw[W], db[W] f32 are ACCUMULATED with atomics (caller zero-fills).
 * x, y, s_out, dy, dx, dadd: bf16 [rows, W] contiguous; w, b: bf16 [W]; W % 8 == 0, W <= 2048.
 * ------------------------------------------------------------------------------------------- */
int b200mm_layernorm_fwd(const void* x, const void* add0, const void* add1, int64_t add_period, const void* w, const void* b,
                         void* y, void* s_out, float* mean, float* rstd, int64_t rows, int32_t W, float eps, void* stream);
int b200mm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* w, const void* dadd,
                         void* dx, float* dw, float* db, int64_t rows, int32_t W, void* stream);

/* BertEmbeddings.forward, clip/modeling_bert.py:86-103:  y = LN(word[ids] + pos[row % L] + type[type_ids]).
 * ids/type_ids are int64 device arrays of `rows` entries (bit-exact indexing); s_out gets the bf16 sum. */
int b200mm_embed_layernorm_fwd(const void* word, const int64_t* ids, const void* pos, int64_t L, const void* type,
                               const int64_t* type_ids, const void* w, const void* b, void* y, void* s_out, float* mean,
                               float* rstd, int64_t rows, int32_t W, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head self-attention, forward and backward (S and P never leave the SM).
 *   ViT : nn.MultiheadAttention in ResidualAttentionBlock.attention, clip/model.py:245-251
 *   BERT: BertSelfAttention.forward, clip/modeling_bert.py:134-172 (key_bias = (1-mask)*-10000, f32 [B, L])
 * qkv: bf16 [B, L, ld]; head h of q/k/v at column {q,k,v}_off + h*head_dim.  o: bf16 [B, L, ldo] (head h at h*head_dim).
 * lse: f32 [B, H, L] (natural log).  head_dim: any multiple of 16 up to 128; any L.  scale = 1/sqrt(head_dim).
 * M2-Encoder: MultiheadAttention.forward, prj/M2_Encoder/vlmo/torchscale/component/multihead_attention.py:85-150 (head_dim 64 / 128).
 * bwd writes dq/dk/dv into dqkv (same layout as qkv) and D = rowsum(dO*O) into dsum f32 [B, H, L].
 * ------------------------------------------------------------------------------------------- */
int b200mm_attention_fwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                         const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, void* stream);
int b200mm_attention_bwd(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o,
                         int64_t ldo, const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H,
                         int32_t L, int32_t head_dim, float scale, void* stream);
/* The same with dropout on the attention probabilities (BertSelfAttention: `attention_probs = self.dropout(attention_probs)`,
 * clip/modeling_bert.py:124,158): P is normalised first, then element (b, h, query, key) is zeroed with probability drop_p and the
 * survivors scaled by 1/(1-drop_p). The decision is the counter-based rule of b200mm_dropout with stream = b*H + h and
 * index = query*L + key, regenerated (not stored) in backward from the same drop_seed. lse is that of the un-dropped softmax.
 * drop_p = 0 is exactly b200mm_attention_fwd / _bwd. b200mm_attention_dropout_mask writes the decisions (1 = kept) as bytes [B, H, L, L]
 * (test / debugging aid: the product path never materialises the mask). */
int b200mm_attention_fwd_dropout(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, void* o, int64_t ldo, float* lse,
                                 const float* key_bias, int32_t B, int32_t H, int32_t L, int32_t head_dim, float scale, float drop_p,
                                 uint64_t drop_seed, void* stream);
int b200mm_attention_bwd_dropout(const void* qkv, int64_t ld, int32_t q_off, int32_t k_off, int32_t v_off, const void* o, const void* d_o,
                                 int64_t ldo, const float* lse, const float* key_bias, void* dqkv, float* dsum, int32_t B, int32_t H,
                                 int32_t L, int32_t head_dim, float scale, float drop_p, uint64_t drop_seed, void* stream);
int b200mm_attention_dropout_mask(uint8_t* keep, int32_t B, int32_t H, int32_t L, float drop_p, uint64_t drop_seed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Contrastive similarity + log-softmax (tcgen05 GEMM with reduction epilogues).  z[m,n] = alpha * <a_m, b_n>.
 * Replaces  logit_scale.exp() * I @ T.t()  + cross-entropy (clip/cn_model.py:221-223, dmae_utils.py:528-537),
 * get_l1_simi_matrix + get_mil_nce_loss (prj/base_vtp/roi_univl/univl/model/univl_video_ret.py:146-226).
 * a: bf16 [M, K] (this rank's rows), b: bf16 [N, K] (all gathered rows; b_mn = 1: stored [K, N] like the MoCo queue
 * [dim, K_queue], moco_utils.py:39-52); the positive of row m is column m + diag_off (out of range = no positive column).
 *   lse_partials: per row and 256-column tile the (max, sum exp) pair, and the diagonal logit   (forward, nothing [M,N] written)
 *   lse_merge   : partials of one or two blocks -> lse[m]; loss_sum += sum_m (lse[m] - diag[m]); sub_diag > 0 removes e^diag
 *                 from the sum (MIL-NCE double-counted positive), sub_diag < 0 adds it (MoCo positive logit)
 *   softgrad    : G[m,n] = alpha * coef * (exp(z - row_lse[m]) - diag_sub*[n == m+diag_off]) as bf16 (dL/d<a_m,b_n>),
 *                 dscale += sum dL/dz * z (gradient w.r.t. log-temperature); feed G to b200mm_gemm_bf16 for dA, dB.
 *                 N (a multiple of 8) may include zero padding rows of b: columns >= n_valid get G = 0.
 * ------------------------------------------------------------------------------------------- */
int32_t b200mm_contrast_num_tiles(int64_t N);
int b200mm_contrast_lse_partials(const void* a, int64_t lda, const void* b, int64_t ldb, int32_t b_mn, int64_t M, int64_t N, int64_t K,
                                 float alpha, int64_t diag_off, float* part_max, float* part_sum, float* diag, void* stream);
int b200mm_contrast_lse_merge(const float* maxA, const float* sumA, int32_t tilesA, const float* maxB, const float* sumB,
                              int32_t tilesB, const float* diag, int32_t sub_diag, float* lse, float* loss_sum, int64_t M,
                              void* stream);
int b200mm_contrast_softgrad(const void* a, int64_t lda, const void* b, int64_t ldb, int32_t b_mn, int64_t M, int64_t N, int64_t K,
                             int64_t n_valid,
                             float alpha, int64_t diag_off, const float* row_lse, float coef, float diag_sub, int32_t diag_zero,
                             void* G, int64_t ldg, float* dscale, void* stream);
/* Symmetric losses, both directions per launch (grouped: problem 0 = rows a0 x columns b0, problem 1 = rows a1 x columns b1, same M, N, K;
 * one persistent launch fills the SMs where two half-empty tile waves ran before). alpha_dev / coef_dev (device f32 scalars, may be null)
 * replace alpha / multiply coef, so neither exp(logit_scale) nor the upstream gradient has to be read back by the host.
 *   lse_partials_pair: the lse_partials outputs of both problems.
 *   softgrad_pair    : TWO-SIDED gradient tiles. Logit z[m,n] of problem 0 is also entry (n, m) of problem 1's block on the rank that owns
 *                      row n, where it is normalised by that row's LSE; with the row LSEs of ALL ranks gathered (col_lse [N]) a rank forms
 *                        G[m,n] = alpha*coef*( wr*exp(z - row_lse[m]) + wc*exp(z - col_lse[n]) - diag_sub*[n == m+diag_off] )
 *                      (wr / wc = 0 on the diagonal if row_diag_zero / col_diag_zero: MIL-NCE's excluded own-video logit), and
 *                      d a_m = sum_n G[m,n] b_n is the COMPLETE gradient of the global loss w.r.t. its own rows: no gradient exchange
 *                      (the reduce-scatter of GradientAllGather.backward, antmmf/utils/distributed_utils.py:104-116, is replaced by an
 *                      all-gather of 2*B floats in forward). dscale (problem 0 only) += sum dL/dz * z.
 *                      The stored G leaves the diagonal's -diag_sub term OUT (dscale counts it): the caller adds
 *                      -diag_sub*coef*alpha*b_{m+diag_off} to row m's gradient in fp32, so the one large entry of a row is never rounded to bf16. */
int b200mm_contrast_lse_partials_pair(const void* a0, int64_t lda0, const void* b0, int64_t ldb0, const void* a1, int64_t lda1, const void* b1,
                                      int64_t ldb1, int64_t M, int64_t N, int64_t K, float alpha, const float* alpha_dev, int64_t diag_off,
                                      float* part_max0, float* part_sum0, float* diag0, float* part_max1, float* part_sum1, float* diag1,
                                      void* stream);
int b200mm_contrast_softgrad_pair(const void* a0, int64_t lda0, const void* b0, int64_t ldb0, const void* a1, int64_t lda1, const void* b1,
                                  int64_t ldb1, int64_t M, int64_t N, int64_t K, int64_t n_valid, float alpha, const float* alpha_dev,
                                  int64_t diag_off, const float* row_lse0, const float* col_lse0, const float* row_lse1, const float* col_lse1,
                                  float coef, const float* coef_dev, float diag_sub, int32_t row_diag_zero0, int32_t col_diag_zero0,
                                  int32_t row_diag_zero1, int32_t col_diag_zero1, void* G0, void* G1, int64_t ldg, float* dscale, void* stream);
/* Retrieval rank of the positive pair without materialising the similarity matrix — replaces the sort-based rank of
 * _compute_retrieval_metrics (antmmf/modules/metrics/global_retrieval_recall.py:12-27) and cal_ret_metric
 * (prj/base_vtp/roi_univl/univl/model/univl_video_pretrain.py:294-312):
 *   rank_out[m] += #{ n != pos(m) : alpha*<a_m, b_n> > ref[m] },  pos(m) = gt_col[m] if gt_col else m + diag_off.
 * ref[m] is the positive logit (e.g. `diag` of lse_partials, or b200mm_rowdot for arbitrary gt_col); the caller zero-fills rank_out. */
int b200mm_contrast_rank(const void* a, int64_t lda, const void* b, int64_t ldb, int32_t b_mn, int64_t M, int64_t N, int64_t K,
                         float alpha, int64_t diag_off, const int32_t* gt_col, const float* ref, int32_t* rank_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small HBM-bound helpers.
 * ------------------------------------------------------------------------------------------- */
/* Frame pooling (UnivlVideoBase.forward_img_encoder, prj/base_vtp/roi_univl/univl/model/univl_video_base.py:91-95):
 * y[r,:] = mean over the non-padded positions p of x[r,p,:]; x [R,P,W] bf16, pad [R,P] bytes (!= 0: padded) or null, W % 8 == 0;
 * inv_count[r] = 1/#valid (f32, consumed by the backward: dx[r,p,:] = valid * inv_count[r] * dy[r,:]). */
int b200mm_masked_mean_fwd(const void* x, const uint8_t* pad, void* y, float* inv_count, int64_t R, int32_t P, int32_t W, void* stream);
int b200mm_masked_mean_bwd(const void* dy, const uint8_t* pad, const float* inv_count, void* dx, int64_t R, int32_t P, int32_t W, void* stream);
/* Sub-LayerNorm of the M2-Encoder (BEiT-3 multiway) blocks, fused with the activation in front of it; any row width W % 8 == 0
 * up to 8192 (the FFN sub-LN spans the 4W hidden). Replaces `ffn_layernorm(gelu(fc1 x))`
 * (prj/M2_Encoder/vlmo/torchscale/component/feedforward_network.py:117-128) and, with act = B200MM_ACT_NONE, `inner_attn_ln`
 * (vlmo/torchscale/component/multihead_attention.py:148-149).
 *   fwd: y = LN(act(u))*w + b, mean/rstd [rows] f32 saved;  bwd: du = LN'(dy)*act'(u) with act(u) recomputed, dw/db (f32 [W]) accumulated. */
int b200mm_act_layernorm_fwd(const void* u, int32_t act, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows,
                             int32_t W, float eps, void* stream);
int b200mm_act_layernorm_bwd(const void* dy, const void* u, int32_t act, const float* mean, const float* rstd, const void* w, void* du,
                             float* dw, float* db, int64_t rows, int32_t W, void* stream);
/* y[r,:] = drop[r] ? 0 : x[r,:]  (bf16 [rows, W], drop = bytes): Encoder.forward zeroes the padded token rows,
 * prj/M2_Encoder/vlmo/torchscale/architecture/encoder.py:440; the same call masks the gradient in backward. y may alias x. */
int b200mm_mask_rows(const void* x, const uint8_t* drop, void* y, int64_t rows, int32_t W, void* stream);
/* Stage-2 (cross-modal) retrieval helpers, prj/base_vtp/roi_univl/univl/model/univl_video_ret.py.
 * gather_rows: out[r,:] = src[ids[r],:] (bf16 [n_src, W] -> [rows, W], bit-exact; ids outside [0, n_src) give a zero row). One pass builds
 * the [text ; visual] token rows of many (text, video) pairs, replacing the unsqueeze/repeat/view/cat copies of _cross_similarity (:52-78)
 * and _cross_similarity_hard_mining (:101-131); its backward is b200mm_scatter_add_rows. */
int b200mm_gather_rows(const void* src, const int64_t* ids, void* out, int64_t rows, int64_t n_src, int32_t W, void* stream);
/* ReLU of the similarity head nn.Sequential(Linear, ReLU, Linear) (:24-28): y = max(x, 0); dx = dy * [x > 0]; bf16, n % 8 == 0. */
int b200mm_relu_fwd(const void* x, void* y, int64_t n, void* stream);
int b200mm_relu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream);
/* get_mil_nce_loss (:146-197) on an explicit square f32 score matrix S [B, ld] (n_pair 1) with optional row weights w [B] (forward_stage2
 * :414-433): lse[j] = log(sum_i e^{S[i,j]} + sum_{k!=j} e^{S[j,k]}), *loss_sum += sum_j w_j (lse[j] - S[j,j]) (caller zero-fills and divides by B).
 * bwd: dS [B,B] f32 = (*gout / B) * dLoss/dS with the saved lse. */
int b200mm_mil_nce_matrix_fwd(const float* S, int64_t ld, const float* w, float* lse, float* loss_sum, int32_t B, void* stream);
int b200mm_mil_nce_matrix_bwd(const float* S, int64_t ld, const float* w, const float* lse, const float* gout, float* dS, int32_t B, void* stream);
/* XPOS rotary position embedding (the optional RoPE of the path: prj/M2_Encoder/vlmo/torchscale/component/xpos_relative_position.py:15-62,
 * multihead_attention.py:112-118), in place on the q and k sections of the fused projection output qkv [T, ld] (T = B*L rows, sections of
 * H*hd columns at q_off / k_off): pair (2i, 2i+1) of every head is rotated and scaled with the f32 tables [L, hd/2]
 * q_cos/q_sin = cos/sin(l*inv_freq_i)*scale[l,i], k_cos/k_sin = the same with 1/scale. backward != 0 applies the transposed map. */
int b200mm_xpos_apply(void* qkv, int64_t ld, int32_t q_off, int32_t k_off, const float* q_cos, const float* q_sin, const float* k_cos,
                      const float* k_sin, int64_t T, int32_t L, int32_t H, int32_t hd, int32_t backward, void* stream);
/* Inverted dropout with a counter-based mask (nn.Dropout of the BERT tower: embeddings, self-output, output —
 * clip/modeling_bert.py:84,101,180,232): y[r, c] = keep(seed, r, c) ? x[r, c] / (1 - p) : 0 on bf16 [rows, cols] (cols % 8 == 0, pitches
 * ldx / ldy, y may alias x). keep() is a pure function of (seed, r, c) — hash32 / drop_stream_key / drop_keep in csrc/common.cuh,
 * restated in oracle/restated.py:dropout_keep — so the backward pass applies the SAME call to the incoming gradient, and the fused
 * GEMM epilogue (b200mm_gemm_args.drop_p) produces the identical mask for output element (m, n) = (r, c). */
int b200mm_dropout(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int32_t cols, float p, uint64_t seed, void* stream);
/* y = act(x), bf16, n % 8 == 0 (activation recompute in backward: QuickGELU clip/model.py:222-224, erf-GELU modeling_bert.py:31-37) */
int b200mm_act_fwd(const void* x, void* y, int64_t n, int32_t act, void* stream);
/* out[(row % period), :] += in[row, :]  (f32 atomics, caller zero-fills): period 1 = bias gradient of nn.Linear,
 * period L = gradient of positional_embedding (clip/model.py:323) / position_embeddings (modeling_bert.py:97) */
int b200mm_rowsum_periodic(const void* in, float* out, int64_t rows, int32_t W, int64_t period, void* stream);
/* out[ids[row], :] += in[row, :] (f32 atomics), rows with ids == skip_id dropped: gradient of nn.Embedding with
 * padding_idx (modeling_bert.py:71-73) */
int b200mm_scatter_add_rows(const void* in, const int64_t* ids, float* out, int64_t rows, int32_t W, int64_t skip_id,
                            int64_t n_out_rows, void* stream);
int b200mm_cast_f32_bf16(const float* x, void* y, int64_t n, float scale, void* stream);
/* y = x / max(||x||, eps) per row (cn_model.py:217-218; F.normalize in univl_video_base.py:114,158) and its backward (dy f32) */
int b200mm_rownorm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int32_t W, float eps, void* stream);
int b200mm_rownorm_bwd(const float* dy, const void* x, const float* inv_norm, void* dx, int64_t rows, int32_t W, void* stream);
/* out[r] = scale * <a[r,:], b[r,:]> (f32): MoCo positive logits, einsum("bh,bh->b") in univl_video_ret.py:292-296 */
int b200mm_rowdot(const void* a, const void* b, float* out, int64_t rows, int32_t W, float scale, void* stream);
/* momentum update of a key-encoder parameter, pk (f32) = m*pk + (1-m)*pq (pq bf16 or f32): moco_utils.py:55-69 */
int b200mm_ema_update(float* pk, const void* pq, int32_t pq_is_bf16, int64_t n, float m, void* stream);
/* ViT stem patch extraction for conv1 (kernel == stride == p, no bias; clip/model.py:289-295,310-312):
 * img bf16 [B, C, H, W] -> out bf16 [B*(Np+1), Kp]; row b*(Np+1) (class-token slot) and columns >= C*p*p are zero. */
int b200mm_im2row(const void* img, void* out, int64_t B, int32_t C, int32_t H, int32_t W, int32_t p, int32_t Kp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MM_H_ */
