/* clipk -- C ABI of the B200 (sm_100a) kernels behind EasyNLP's CLIP contrastive path.
 *
 * The reference (alibaba/EasyNLP) is pure Python/PyTorch and has NO FFI on this path:
 * every entry point below replaces a sequence of ATen library calls made by the cited reference lines.
 * The binding a maintainer adds on the reference side is a ctypes stub; the host-side
 * mirror of the reference's plugin interface lives in easynlp_b200/ (Python, like the reference).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless named host_*.
 *   - the caller owns every buffer (workspace included); no hidden allocation, no global state except
 *     a per-process error string, cached device attributes and ONE 512-byte per-device pool of GEMM tile-scheduler
 *     counters (allocated by the first clipk_gemm_bf16 call on a device, which therefore must not be inside a stream capture).
 *   - one host thread per process drives the library (one process per GPU, as torch.distributed.launch does for the
 *     reference); kernels are re-entrant per stream.
 *   - every function is asynchronous on `stream` and returns 0 on success or a negative CLIPK_ERR_*;
 *     clipk_last_error() then describes the failure.
 *   - bf16 = IEEE bfloat16 storage; "ld*" = leading dimension in ELEMENTS.
 */
#ifndef CLIPK_H_
#define CLIPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_API_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define CLIPK_ERR_ARG (-1)
#define CLIPK_ERR_CUDA (-2)
#define CLIPK_ERR_UNSUPPORTED (-3)

#define CLIPK_BF16 0
#define CLIPK_F32 1

/* -------------------------------------------------------------------------------------------- library */
const char* clipk_last_error(void);
int clipk_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t clipk_launch_count(void);

/* -------------------------------------------------------------------------------------------- GEMM
 * Replaces nn.Linear / nn.Conv2d(patch) / MultiheadAttention in/out projections and their autograd
 * (reference: modeling_chineseclip.py:188-204,224-251; modeling_bert.py:145-147,264-268,329-346).   */
#define CLIPK_EPI_LINEAR 0       /* out = alpha*acc + bias + residual                    (out bf16|f32, optional bf16 out2) */
#define CLIPK_EPI_QUICK_GELU 1   /* z = acc + bias; out2 = z*sigmoid(1.702 z) (bf16); out = d/dz of it (bf16, saved for backward)
                                    (modeling_chineseclip.py:179-181)                                                     */
#define CLIPK_EPI_ERF_GELU 2     /* same with gelu_erf (modelzoo/activations.py:45-48): out2 = act(z), out = act'(z)              */
#define CLIPK_EPI_MUL_AUX 3      /* out = (alpha*acc + bias) * aux (+ residual): backward of modes 1/2 with aux = act'(z)          */
#define CLIPK_EPI_RANK_COUNT 4   /* no matrix output: out (int32 [M]) += #{j != label_offset + i : alpha*acc_ij > aux_f32[i]}       */
#define CLIPK_EPI_ATOMIC_ADD 5   /* out(f32) += acc  via red.add -- split-K weight gradients                                */

typedef struct {
  int mode;              /* CLIPK_EPI_* */
  int out_dtype;         /* CLIPK_BF16 | CLIPK_F32 */
  void* out;             /* [M, ldo] */
  int ldo;
  void* out2;            /* optional bf16 [M, ldo2] */
  int ldo2;
  const float* bias;     /* optional [N] */
  const float* residual; /* optional f32 [M, ldr] */
  int ldr;
  const void* aux;       /* bf16 [M, ldaux] multiplier for CLIPK_EPI_MUL_AUX (the saved activation derivative) */
  int ldaux;
  float alpha;           /* 0 is read as 1 */
  float* colsum;         /* optional f32 [N]: += column sums of the stored output (bias gradient of the producing layer) */
  int label_offset;      /* CLIPK_EPI_RANK_COUNT: column label_offset + i is row i's own match and is not counted; aux = f32 [M] thresholds */
} clipk_epilogue_t;

/* D[M,N] = epilogue(op(A) x op(B)), bf16 operands, fp32 accumulation in TMEM (tcgen05).
 *   a_mn_major = 0: A is [M,K] row-major;  1: A is [K,M] row-major (A^T is the logical operand)
 *   b_mn_major = 0: B is [N,K] row-major (nn.Linear weight);  1: B is [K,N] row-major
 *   splits > 1: split-K, requires CLIPK_EPI_ATOMIC_ADD into a pre-zeroed/accumulating fp32 output.
 * lda, ldb, N multiples of 8 (N: any value for CLIPK_EPI_RANK_COUNT); base pointers 16-byte aligned.   */
int clipk_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N, int K,
                    const clipk_epilogue_t* epi, int splits, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- dropout
 * Inverted dropout (keep with prob 1-p, scale 1/(1-p)) as in nn.Dropout of the BERT tower (modeling_bert.py:85,128,238,267,345).
 * Masks are never stored: forward and backward regenerate them from Philox4x32-10 keyed by
 * (seed ^ *dev_offset, site) and the element index.  dev_offset (optional device uint32, e.g. the optimizer's step counter)
 * makes the mask change every step even when the launches are replayed from a CUDA graph.  p == 0 or a NULL pointer = off. */
typedef struct {
  float p;
  unsigned long long seed;
  const unsigned int* dev_offset;
  unsigned int site;         /* distinct per dropout call site (layer, position in the layer) */
} clipk_dropout_t;
/* writes the multipliers (0 or 1/(1-p)) of a [rows, cols] site exactly as the fused kernels apply them (tests / debugging) */
int clipk_dropout_mask(float* out, int rows, int cols, const clipk_dropout_t* drop, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- attention
 * Fused softmax(Q K^T / 8 + key_mask) V for head dim 64 on packed projections qkv[B*L, 3d] (Q|K|V blocks, head h at
 * columns h*64).  Replaces nn.MultiheadAttention's SDPA (modeling_chineseclip.py:188,198-200) and BertSelfAttention
 * (modeling_bert.py:210-244, additive mask from modeling_utils.py:438-439).  Forward: L <= 272 (ViT-L/14 has 257 tokens); backward: L <= 256.
 * key_mask: optional f32 [B, L] additive (0 / -10000).  lse: f32 [B, H, L] saved for backward.                  */
int clipk_attention_fwd(const void* qkv, const float* key_mask, void* ctx, float* lse, int B, int L, int H, int d,
                        const clipk_dropout_t* drop /* optional: dropout on the probabilities, row = (b*H+h)*L+q, col = key */,
                        cudaStream_t stream);
/* dqkv[B*L, 3d] (bf16) from dctx[B*L, d] (bf16); ctx / lse are the forward outputs.
 * dqkv_colsum: optional f32 [3d], += column sums of dqkv taken on the fp32 accumulators = the gradient of the fused
 * query/key/value projection bias (in_proj_bias, modeling_chineseclip.py:188; query/key/value.bias, modeling_bert.py:145-147). */
int clipk_attention_bwd(const void* qkv, const float* key_mask, const void* ctx, const float* lse, const void* dctx,
                        void* dqkv, float* dqkv_colsum, int B, int L, int H, int d,
                        const clipk_dropout_t* drop /* same as forward; L <= 128 */, cudaStream_t stream);

/* Causal self-attention of OPEN_CLIP's text tower (additive -inf above the diagonal, modeling_openclip.py:296-301,346-352); same layout
 * and outputs as above, no key mask / dropout; forward L <= 272, backward L <= 128. */
int clipk_attention_causal_fwd(const void* qkv, void* ctx, float* lse, int B, int L, int H, int d, cudaStream_t stream);
int clipk_attention_causal_bwd(const void* qkv, const void* ctx, const float* lse, const void* dctx, void* dqkv, float* dqkv_colsum, int B, int L,
                               int H, int d, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- LayerNorm
 * y = (x - mean) * rstd * gamma + beta over the last dim d (d % 128 == 0, d <= 1024), fp32 statistics
 * (modeling_chineseclip.py:170-176 eps 1e-5; nn.LayerNorm eps 1e-12 in modeling_bert.py:84,266,344).
 * x: f32 rows with stride ldx; y_bf16 / y_f32 / mean / rstd optional outputs (contiguous).
 * Optional fused residual add: if add_bf16 != NULL (bf16 rows, stride ldadd) the kernel normalises xs = x + add and, if
 * x_out != NULL, stores xs (f32, contiguous) -- `x + attention(...)` / `x + mlp(...)` of modeling_chineseclip.py:203-204 and
 * `dense(...) + input_tensor` of modeling_bert.py:266,344 without a residual read in the GEMM epilogue.               */
int clipk_layernorm_fwd(const float* x, long long ldx, const void* add_bf16, long long ldadd, float* x_out,
                        const float* gamma, const float* beta, float eps, void* y_bf16, float* y_f32, float* mean,
                        float* rstd, int rows, int d, const clipk_dropout_t* drop /* optional */,
                        int drop_mode /* 1: dropout(add) before the sum; 2: dropout on the outputs */, cudaStream_t stream);
/* g = dy (+ dy_add); dx = LN'(g) (+ dx_add) -> dx_f32 (stride lddx) / dx_bf16; dgamma, dbeta, dbias(=colsum dx) are
 * ACCUMULATED with fp32 atomics (optional).                                                                     */
int clipk_layernorm_bwd(const void* dy, int dy_is_f32, const float* dy_add, const float* x, long long ldx,
                        const float* gamma, const float* mean, const float* rstd, const float* dx_add, float* dx_f32,
                        long long lddx, void* dx_bf16, float* dgamma, float* dbeta, float* dbias, int rows, int d,
                        const clipk_dropout_t* drop /* optional */,
                        int drop_mode /* 1: mask dx_bf16 + dbias (fwd mode 1); 2: mask the incoming gradient (fwd mode 2) */,
                        cudaStream_t stream);

/* out[n] += sum over rows of x[rows, n] (bf16 or f32, row stride ldx): bias / positional / token-type gradients */
int clipk_colsum(const void* x, int is_f32, long long ldx, float* out, int rows, int n, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- embeddings
 * ViT patch embed (modeling_chineseclip.py:224,237-241): pixels f32 [B,3,R,R] -> bf16 patches [B*g*g, 3*P*P]
 * (column order of conv1.weight.view(W,-1)); then a GEMM; then token assembly x0 = [cls | patches] + pos.       */
/* ld_out: leading dimension of the patch matrix in elements (0 = 3*P*P); ViT-L-14 pads 588 -> 592 so that rows stay 16-byte multiples
 * for the TMA loads of the patch GEMM (the caller zero-fills the padding columns once) */
int clipk_im2col_patches(const float* pixels, void* patches_bf16, int B, int R, int P, int ld_out, cudaStream_t stream);
int clipk_vit_assemble(const float* patch_f32, const float* cls, const float* pos, float* x0, int B, int L, int W,
                       cudaStream_t stream);
int clipk_vit_assemble_bwd(const float* dx0, void* dpatch_bf16, int B, int L, int W, cudaStream_t stream);
/* BERT embeddings (modeling_bert.py:95-129): e = word[ids] + type[0] + pos[l]; backward scatters into dword
 * (row 0 = padding_idx receives no gradient).  ids: int64 [rows] on device.                                     */
int clipk_bert_embed(const long long* ids, const float* word, const float* pos, const float* type0, float* e,
                     float* key_mask /* optional [rows]: (ids==0) * -10000 */, int rows, int L, int H, int vocab,
                     cudaStream_t stream);
int clipk_bert_embed_bwd(const long long* ids, const float* de, float* dword, int rows, int H, int vocab,
                         cudaStream_t stream);

/* RoBERTa-style embeddings of the huggingface_clip branch's text tower (modelzoo/models/roberta/modeling_roberta.py:65-130):
 * position ids = cumsum(ids != pad) * (ids != pad) + pad (:1497-1510); e = word[ids] + pos[pos_ids] + type[type_ids];
 * key_mask = (1 - attention_mask) * -10000 (the batch's mask, appzoo/clip/model.py:132-134) or from ids != pad when NULL.
 * Backward scatters de into the three tables; rows `pad_id` of the word and position tables (padding_idx) get no gradient. */
int clipk_position_ids(const long long* ids, int* pos_ids, int B, int L, int pad_id, cudaStream_t stream);
int clipk_embed_gather(const long long* ids, const int* pos_ids, const long long* type_ids /* optional */,
                       const long long* attn_mask /* optional */, const float* word, const float* pos, const float* type, float* e,
                       float* key_mask /* optional */, int rows, int H, int vocab, int npos, int ntype, int pad_id, cudaStream_t stream);
int clipk_embed_gather_bwd(const long long* ids, const int* pos_ids, const long long* type_ids, const float* de, float* dword,
                           float* dpos, float* dtype, int rows, int H, int vocab, int npos, int ntype, int pad_id, cudaStream_t stream);
/* EOT pooling of OPEN_CLIP.encode_text (modeling_openclip.py:367-369: x[arange(B), text.argmax(-1)]): idx[b] = first argmax of ids[b, :];
 * out[b, :] = x[b*L + idx[b], :] (bf16 rows, W % 8 == 0); backward: dst[b*L + idx[b], :] = src[b, :] into a zero-filled fp32 [B*L, W] */
int clipk_argmax_rows(const long long* ids, int* idx, int B, int L, cudaStream_t stream);
/* [SEP] pooling of Wukong's TextTransformer (modelzoo/models/wukong/modeling_wukong.py:349,359: x[(ids == 102).nonzero()]): idx[b] = first
 * position of `token` in ids[b, :] (0 when absent), count[b] (optional) = number of occurrences */
int clipk_find_token_rows(const long long* ids, long long token, int* idx, int* count /* optional */, int B, int L, cudaStream_t stream);
int clipk_gather_rows_bf16(const void* x_bf16, const int* idx, void* out_bf16, int B, int L, int W, cudaStream_t stream);
int clipk_scatter_rows_f32(const float* src, const int* idx, float* dst, int B, int L, int W, cudaStream_t stream);
/* masked mean over the T frame embeddings of a video and its backward (Text2VideoRetrieval._mean_pooling_for_similarity_visual,
 * appzoo/text2video_retrieval/model.py:98-104): x f32 [B, T, E], mask int64 [B, T] */
int clipk_frame_pool_fwd(const float* x, const long long* mask, float* out, int B, int T, int E, cudaStream_t stream);
int clipk_frame_pool_bwd(const float* dout, const long long* mask, float* dx, int B, int T, int E, cudaStream_t stream);
/* pooler activation (RobertaPooler: tanh(dense(h[:,0])), modeling_roberta.py:559-575; used as the text feature by the
 * huggingface_clip branch, appzoo/clip/model.py:135): y = tanh(x) (+ bf16 copy); dx = dy * (1 - y^2) */
int clipk_tanh_fwd(const float* x, float* y, void* y_bf16 /* optional */, long long n, cudaStream_t stream);
int clipk_tanh_bwd(const float* dy, const float* y, float* dx /* optional */, void* dx_bf16 /* optional */, long long n, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- head
 * y = x / ||x||_2 per row (modeling_chineseclip.py:360,363) and its backward.                                   */
int clipk_l2norm_fwd(const float* x, float* y, float* norm, int rows, int d, cudaStream_t stream);
int clipk_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx_f32, void* dx_bf16, int rows, int d,
                     cudaStream_t stream);
int clipk_cast_bf16(const float* x, void* y_bf16, long long n, cudaStream_t stream);
/* y += alpha * x (f32, n % 4 == 0): adds reduce-scattered gallery gradients onto the local embedding gradients */
int clipk_axpy(const float* x, float* y, float alpha, long long n, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- contrastive loss
 * One strip of the symmetric InfoNCE (appzoo/clip/model.py:148-164): logits = exp(logit_scale_log) * Q K^T for
 * nq local queries against nk gallery rows, label(i) = label_offset + i; online log-sum-exp, logits optionally
 * written (S_out[i*lds + j], or [j*lds + i] if transpose_out).  loss_rows[i] = lse_i - logit_{i,label}.         */
int clipk_ce_strip_fwd(const float* Q, const float* K, const float* logit_scale_log, int label_offset, float* S_out,
                       long long lds, int transpose_out, float* lse, float* loss_rows, int nq, int nk, int E,
                       cudaStream_t stream);
/* gradient of coef * sum_i CE_i w.r.t. the OWNED rows: own_is_query=1 -> d/dQ (lse indexed by own rows),
 * own_is_query=0 -> d/dK (lse indexed by streamed rows).  out (+)= ...; dscale_log += sum dS*S (query mode).    */
int clipk_ce_strip_bwd(const float* own, const float* streamed, const float* logit_scale_log, const float* lse,
                       int label_offset, float coef, int own_is_query, float* out, int accumulate, float* dscale_log,
                       int n_own, int n_streamed, int E, cudaStream_t stream);
/* Tensor-core form of the same strips (what the training step uses: the strip kernels above occupy nq/32 CTAs and their time
 * grows with the GLOBAL batch).  fp32 rows are split into bf16 hi + lo and concatenated so that ONE clipk_gemm_bf16 with K = 3E
 * gives <q,k> ~= <qh,kh> + <qh,kl> + <ql,kh> (fp32-level logits): pattern 0 = [hi|hi|lo] (query side), 1 = [hi|lo|hi] (gallery). */
int clipk_split_bf16x3(const float* x, void* out_bf16, int rows, int cols, int pattern, long long ld_out, cudaStream_t stream);
/* S[nq, nk] (ld lds): raw dots in, scaled logits exp(logit_scale_log) * S out (in place); lse[i]; loss_rows[i] = lse_i - S[i,label] */
int clipk_ce_rows_fwd(float* S, long long lds, const float* logit_scale_log, int label_offset, float* lse, float* loss_rows,
                      int nq, int nk, cudaStream_t stream);
/* dS[nq, ldds] (bf16, columns >= nk zero-filled) = exp(logit_scale_log) * coef * (exp(S - lse) - onehot(label)) from the scaled
 * logits: the A operand of dOwn = dS * gallery_hi and (MN-major) of dGallery = dS^T * own_hi;  dscale_log += sum coef (p - 1hot) S */
int clipk_ce_rows_bwd(const float* S, long long lds, const float* logit_scale_log, const float* lse, int label_offset, float coef,
                      void* dS_bf16, long long ldds, float* dscale_log, int nq, int nk, cudaStream_t stream);
int clipk_reduce_sum(const float* x, int n, float scale, float* out, int accumulate, cudaStream_t stream);
/* Retrieval: rank_out[i] = #{gallery j : <Q_i,K_j> > <Q_i,K_label(i)>} (int32); hit@K <=> rank < K.  Replaces
 * CLIPEvaluator's N x N matrix + per-row torch.sort (appzoo/clip/evaluator.py:47-61).                          */
int clipk_retrieval_rank(const float* Q, const float* K, int label_offset, int* rank_out, int nq, int nk, int E,
                         cudaStream_t stream);
/* The same ranks from the tensor cores (blocked retrieval, SURVEY 8f.1 / the 1 M x 1 M configuration): bf16 hi/lo splits of both
 * sides (fp32-level dots, K = 3E), the match score thr_i = <Q_i, K_label(i)> per query, then ONE GEMM whose epilogue
 * (CLIPK_EPI_RANK_COUNT) compares every accumulator with thr_i and adds the count to rank_out[i] -- the N x N matrix never exists.
 * workspace: caller-owned device memory of clipk_retrieval_rank_tc_workspace(nq, nk, E) bytes.                                   */
size_t clipk_retrieval_rank_tc_workspace(int nq, int nk, int E);
int clipk_retrieval_rank_tc(const float* Q, const float* K, int label_offset, int* rank_out, int nq, int nk, int E,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- optimizer
 * Global-norm clip (core/trainer.py:325) + the reference AdamW (core/optimizers.py:437-462) over a flat buffer.
 * norm_and_coef[0] = ||g||_2, [1] = min(1, max_norm / (norm + 1e-6)).  workspace: doubles, len >= 4*#SM.        */
int clipk_grad_norm(const float* g, long long n, float max_norm, double* workspace, int workspace_len,
                    float* norm_and_coef, cudaStream_t stream);
/* p, m, v updated in place; w_bf16 (optional) = bf16(p) refreshed for the GEMMs; clip_coef optional device scalar.
 * dev_hyper (optional, device float[2] = {lr, lr*sqrt(1-b2^k)/(1-b1^k)}) overrides lr/step so that a captured CUDA
 * graph can be replayed while the schedule advances on the device (clipk_adam_schedule).                          */
int clipk_adamw_step(float* p, const float* g, float* m, float* v, void* w_bf16, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, const float* clip_coef, const float* dev_hyper,
                     cudaStream_t stream);
/* step_dev[0] += 1 and hyper_dev = {lr_k, step_size_k} for the reference's warmup-linear schedule
 * (core/optimizers.py:191-204); t_total <= 0 -> constant lr.                                                      */
int clipk_adam_schedule(int* step_dev, float* hyper_dev, float base_lr, int warmup_steps, int t_total, float beta1,
                        float beta2, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- peer-memory collectives (multi-GPU head)
 * One process per GPU; every rank allocates ONE peer buffer (galleries + gallery gradients + flags), exports its CUDA-IPC handle, opens
 * the others' (easynlp_b200/distributed.py: PeerGroup).  gallery_ptrs / flag_ptrs / src_ptrs are DEVICE arrays of `world` pointers, entry p
 * = the corresponding region in rank p's buffer (p == rank: the local pointer).  See csrc/peer.cu for the protocol.                     */
int clipk_peer_alloc(void** dev_ptr, size_t bytes);                 /* cudaMalloc'd (IPC-exportable), zero-filled */
int clipk_peer_free(void* dev_ptr);
int clipk_peer_export(const void* dev_ptr, unsigned char* handle64); /* 64-byte cudaIpcMemHandle_t */
int clipk_peer_open(const unsigned char* handle64, void** dev_ptr);
int clipk_peer_close(void* dev_ptr);
/* y = x / ||x|| (modeling_chineseclip.py:360,363) stored locally AND into row block `rank` of every rank's gallery: the fused
 * producer + all-gather of the embedding shards (replaces all_gather_into_tensor of distributed.py) */
int clipk_l2norm_allgather(const float* x, float* y_local, float* norm, float* const* gallery_ptrs, int world, int rank, int rows, int d,
                           cudaStream_t stream);
/* flags[p][channel * world + rank] = epoch on every peer p (after a system-scope fence) / wait until all `world` flags of the channel in
 * MY flag array reached epoch (bounded spin, traps on timeout).  Epochs increase monotonically per channel. */
int clipk_peer_signal(unsigned int* const* flag_ptrs, int world, int rank, int channel, unsigned int epoch, cudaStream_t stream);
int clipk_peer_wait(const unsigned int* my_flags, int world, int channel, unsigned int epoch, cudaStream_t stream);
/* out[r, :] (+)= sum over peers p of src_ptrs[p][(rank * rows + r), :]: reduce-scatter of the gallery gradients by peer loads */
int clipk_peer_reduce_rows(float* const* src_ptrs, int world, int rank, float* out, int rows, int d, int accumulate, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- image preprocessing
 * Decoded 8-bit RGB images -> the normalised fp32 [n, 3, size, size] batch of the image tower: replaces, per batch, the host chain
 * _resize(image, 224, BICUBIC) -> _center_crop(224) -> /255 -> (x - mean) / std of CLIPDataset.convert_single_row_to_example and
 * CLIPPredictor.preprocess (easynlp/appzoo/clip/data.py:29-135,263-272; predictor.py:100-110).  BIT-EXACT with Pillow's 8-bit bicubic
 * resample (fixed-point taps, uint8 intermediate) and numpy's float32 normalisation.
 *   pixels   device blob holding every image as interleaved RGB rows (h * w * 3 bytes at desc[i].src)
 *   desc     device array [n]; `tmp` = byte offset of the image's h * size * 3 byte slot in the scratch part of the workspace
 *   max_h    largest h of the batch (launch bound); kmax >= clipk_preprocess_kmax(w, h, size) of every image
 *   workspace_bytes >= clipk_preprocess_workspace(n, size, kmax, scratch_bytes), scratch_bytes = sum of the slots
 * clipk_preprocess_status reads back (and synchronises on) the overflow flag of the last call: 1 = a kmax too small was passed.      */
typedef struct clipk_image_desc { long long src; int w, h; long long tmp; } clipk_image_desc;
int clipk_preprocess_kmax(int w, int h, int size);                                             /* host helper, no CUDA call */
size_t clipk_preprocess_workspace(int n, int size, int kmax, long long scratch_bytes);        /* host helper, no CUDA call */
int clipk_preprocess_images(const unsigned char* pixels, const clipk_image_desc* desc, int n, int max_h, int size, int kmax,
                            const float* mean3 /* host */, const float* std3 /* host */, float* out, void* workspace,
                            size_t workspace_bytes, long long scratch_bytes, cudaStream_t stream);
int clipk_preprocess_status(const void* workspace, int n, int size, int kmax, cudaStream_t stream);

/* -------------------------------------------------------------------------------------------- native WordPiece (host code)
 * BertTokenizer of the reference (modelzoo/models/bert/tokenization_bert.py:67-504) for the call the CLIP application makes
 * (appzoo/clip/data.py:262-264): [CLS] + wordpieces (truncated to max_length - 2) + [SEP] + [PAD]s, attention mask 1/0.
 * unicode_table_path: the table written by tools/gen_unicode_table.py (Python unicodedata predicates for the BMP).
 * encode returns the number of non-padding positions, CLIPK_WP_FALLBACK when the text holds a code point outside the supported set
 * (the caller tokenizes that text with the Python restatement), or a negative CLIPK_ERR_*.  HOST pointers throughout.          */
#define CLIPK_WP_FALLBACK (-100)
void* clipk_wp_create(const char* vocab_path, const char* unicode_table_path, int do_lower_case);
void clipk_wp_destroy(void* handle);
int clipk_wp_encode(void* handle, const char* utf8_text, int max_length, long long* host_input_ids, long long* host_attention_mask);
int clipk_wp_encode_batch(void* handle, const char* const* utf8_texts, int n, int max_length, long long* host_input_ids,
                          long long* host_attention_mask, int* host_status, int threads);

/* counter_dev[0] += value on the stream: the device-resident dropout stream position (one per training forward pass, so that a
 * replayed CUDA graph and every micro-batch of a gradient-accumulation window draw fresh nn.Dropout masks, modeling_bert.py:128,238) */
int clipk_counter_add(int* counter_dev, int value, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPK_H_ */
