/*
 * act3d_b200.h -- C ABI of libact3d_b200.so (hand-written sm_100a kernels).
 *
 * The reference (zhouxian/act3d-chained-diffuser) has no FFI: its hot path is eager
 * PyTorch.  The drop-in boundary is therefore the nn.Module surface (model.Act3D,
 * model.DiffusionPlanner); these entry points are what those modules call instead of
 * the eager ops.  Each function cites the reference code it replaces (paths relative
 * to /root/reference).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller allocates all inputs, outputs and workspaces (PyTorch owns memory);
 *   - `stream` is a cudaStream_t passed as void*; functions only enqueue work and
 *     never synchronise; one caller thread per device;
 *   - return 0 on success, a negative A3D_E* code otherwise; a3d_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - floats are fp32, indices int32 unless stated; layouts are row-major, last
 *     dimension fastest.
 */
#ifndef ACT3D_B200_H_
#define ACT3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3D_OK 0
#define A3D_EINVAL (-1)   /* bad argument / unsupported shape            */
#define A3D_ECUDA (-2)    /* CUDA runtime error at launch                */
#define A3D_ENOSUPPORT (-3)

#define A3D_TILE_KEYS 64  /* keys per K/V tile image                      */
#define A3D_HEAD_PAD 16   /* head_dim 15 padded to 16 (slot 15: see below) */

const char* a3d_last_error(void);
int a3d_abi_version(void);
/* Process-wide tuning knobs (not part of the numerical contract):
 *   "xattn_core": attention core of a3d_xattn_stack: 0 (default) = per launch: the tcgen05 / TMEM kernel (attention and
 *                 linear layers on tcgen05.mma, part of the exponentials on the FMA pipe; csrc/a3d_xattn6.cu) for launches
 *                 of >= 296 CTAs, the mma.sync kernel for small ones; 2 = mma.sync (HMMA) kernel, 4 = the round-1
 *                 tcgen05 kernel (attention only on tcgen05; kept as A/B reference), 6 = a3d_xattn6.cu.
 *   "xattn6_np":  how many of the 16 score pairs a thread of a3d_xattn6.cu exponentiates per unit go through the FMA-pipe
 *                 polynomial instead of the MUFU unit: 0, 4, 6 (default, fastest measured), 7, 8, 9, 10.
 *   "xattn_poly": the same knob of the round-1 kernels (of every 8 scores): 0 (default), 2, 3 or 4.
 *   "train_attn_core": a3d_attn_fwd / a3d_attn_bwd: 0 (default) = tensor cores (csrc/a3d_train_mma.cu), 1 = fp32 CUDA cores.
 *   "cd_prefetch_tiles": cd_denoise_loop: K/V tiles prefetched into L2 ahead of the ring, 0 (default, fastest measured) .. 64. */
int a3d_set_option(const char* name, int value);
/* Test / diagnostics counters kept on the current device; the call SYNCHRONISES the device (it is not part of the
 * hot path).  "xattn_replays": attention layers that a CTA of a3d_xattn_stack's tcgen05 kernels replayed in safe mode
 * because the unchecked fast pass overflowed (see csrc/a3d_xattn4.cu).  reset != 0 zeroes the counter after reading. */
int a3d_debug_counter(const char* name, int reset, unsigned long long* value_host);

/* ---------------------------------------------------------------------------------
 * Point pyramid.  Replaces F.interpolate(pcd, scale_factor=1/f, mode='bilinear') +
 * rearrange "(bt ncam) c h w -> bt (ncam h w) c"   (act3d.py:379-383, encoder.py:147-158).
 * pcd [BN][3][H][W] -> out [BN*(H/f)*(W/f)][3].  f in {2,4,8}.  All four taps weigh 0.25 and
 * are accumulated in raster order, 0.25*(((p00+p01)+p10)+p11): bit-identical to torch's CPU
 * bilinear kernel in its single-thread form (its multi-thread form differs from itself by 1 ulp).
 */
int a3d_pcd_pyramid(const float* pcd, int bn, int height, int width, int factor, float* out, void* stream);

/* ---------------------------------------------------------------------------------
 * Local context selection.  Replaces
 *   ((pos - pcd)**2).sum(-1).sqrt().topk(k, largest=False).indices   (act3d.py:244-245)
 * center [B][3], pts [B][N][3] -> idx [B][K] (int32, ascending (distance, index):
 * ties resolved toward the lower index), dist [B][K] (may be NULL).
 * Distance is ((dx*dx + dy*dy) + dz*dz) in fp32 without contraction, IEEE sqrt.
 * K <= 8192, N <= 2^24.
 */
int a3d_local_topk(const float* center, const float* pts, int batch, int n, int k,
                   int32_t* idx, float* dist, void* stream);

/* Same selection for ChainedDiffuser's find_traj_nn (model/utils/utils.py:38-48):
 * key = min over L trajectory points of the SQUARED distance (no sqrt).
 * traj [B][L][3]. */
int a3d_traj_topk(const float* traj, int traj_len, const float* pts, int batch, int n, int k,
                  int32_t* idx, float* dist, void* stream);

/* ---------------------------------------------------------------------------------
 * Token gather.  Replaces rearrange "b ncam c h w -> b (ncam h w) c" + per-sample
 * index gather of features and positions (act3d.py:240-254, diffusion_head.py:290-302).
 * feat [B*ncam][E][hw] (channels_last = 0) or [B*ncam][hw][E] (channels_last = 1, the NHWC layout cuDNN
 * leaves the FPN output in: one contiguous row per token), pcd [B][ncam*hw][3], idx [B][K] or NULL
 * (identity, K = ncam*hw).  Writes rows [0,K) of tok [B][tok_rows][E] and pos [B][tok_rows][3].
 * feat_bias [E] or NULL: per-channel bias added to every gathered feature row -- the bias of the FPN
 * output convolution, deferred from the full 128 x 128 map to the rows that are actually read.
 */
int a3d_gather_tokens(const float* feat, const float* pcd, const int32_t* idx, int batch, int ncam,
                      int embed, int hw, int k, float* tok, float* pos, int tok_rows, int channels_last,
                      const float* feat_bias, void* stream);

/* ---------------------------------------------------------------------------------
 * Memory-bound glue of the image trunk; the convolutions themselves stay on cuDNN.  All maps are
 * fp32 NHWC ([images][h][w][channels]); results are bit-identical to the ATen ops they replace.
 *
 * a3d_trunk_normalize: out[n][p][c] = (rgb[n][c][p] - mean[c]) / std[c] for the 3 colour planes
 *   (NCHW in, NHWC out).  Replaces transforms.Normalize (act3d.py:62, diffusion_head.py:40) and the
 *   channels-last conversion.  mean / std are HOST arrays of 3 floats.  out_channels = 3, or 4 to append a zero
 *   channel (rows of 16 bytes: the 7x7 stem convolution then runs on cuDNN's vectorised NHWC kernels with a
 *   weight padded by a zero input channel).
 * a3d_trunk_maxpool: 3x3, stride 2, padding 1 (the ResNet stem pool, utils/resnet.py:44);
 *   out [images][(h-1)/2+1][(w-1)/2+1][channels]; channels % 4 == 0.
 * a3d_trunk_fpn_topdown: out = (lat + bias) + nearest_upsample(top) -- the top-down merge of
 *   torchvision's FeaturePyramidNetwork.forward with the lateral 1x1 convolution's bias folded in
 *   (bias may be NULL); lat/out [images][h][w][channels] (out may alias lat), top
 *   [images][top_h][top_w][channels]; channels % 4 == 0. */
int a3d_trunk_normalize(const float* rgb, const float* mean_host, const float* std_host, int images, int hw,
                        float* out, int out_channels, void* stream);
int a3d_trunk_maxpool(const float* in, int images, int h, int w, int channels, float* out, void* stream);
int a3d_trunk_fpn_topdown(const float* lat, const float* bias, const float* top, int images, int h, int w,
                          int top_h, int top_w, int channels, float* out, void* stream);

/* ---------------------------------------------------------------------------------
 * Context K/V cache.  Replaces, for `nsets` attention layers that share one context,
 *   k, v = F.linear(ctx, W[E:], b[E:]).chunk(2)   (multihead_custom_attention.py:268-275)
 *   k = embed_rotary(k, cos, sin)                  (:352-353, position_encodings.py:31-34,58-97)
 * and the head split (:356-359).  The rotary table is never materialised: angles are
 * evaluated from pos on the fly.
 * tok [B][tok_rows][E], pos [B][tok_rows][3] (first nk rows used), wkv [nsets] fragment-ordered split-fp16
 * weights of the [2*EP] x [EP] matrix (rows 0..E-1 = W_k, rows EP..EP+E-1 = W_v, EP = 16*H; layout of
 * csrc/a3d_mma_gemm.cuh, produced by packing.pack_kv_set; 2*EP*EP*4 bytes per set), bkv [nsets][2*EP],
 * rope_host[nsets] (1: rotate K of that set).  Output: K/V "tile images", fp16:
 *   kv [nsets][B][ntiles][2][H][64][16],  ntiles = ceil(nk/64),
 * slot 15 of every K and V row holds 1.0 for valid keys (V: the PV product then carries the
 * softmax denominator; K: a query whose slot 15 holds -shift gets shifted scores straight out of the
 * QK^T product -- queries that leave it 0 are unaffected), padded keys are all-zero; inside a 32-byte row the two 16-byte
 * halves are swapped when (key>>2)&1 (bank-conflict-free ldmatrix).
 * (embed, heads) in {(60,4), (120,8)}.
 */
size_t a3d_kv_bytes(int nsets, int batch, int nk, int heads);
int a3d_ctx_kv(const float* tok, const float* pos, int batch, int tok_rows, int nk, int embed, int heads,
               const void* wkv, const float* bkv, const int* rope_host, int nsets, void* kv, void* stream);

/* ---------------------------------------------------------------------------------
 * Fused cross-attention stack (the north-star kernel).  Replaces, per layer,
 *   RelativeCrossAttentionLayer.forward + FeedforwardLayer.forward  (layers.py:300-332)
 *   = q-proj, scale, rotary(q), softmax(QK^T)V over all heads, out-proj, +res, LN,
 *     FFN(relu), +res, LN                      (multihead_custom_attention.py:260-462)
 * for `nlayers` layers over one query tile kept on chip, plus the mask-logit einsum
 * (act3d.py:493-494).  No score matrix, attention weight or rotary table reaches HBM.
 *
 * x0: initial query features; x0_stride_b / x0_stride_n in floats (0,0 = one shared row,
 *     i.e. the ghost-point embedding; (E,0) = one row per sample).
 * qpos [B][nq][3] or NULL (no rotary on this stack: K sets must then be unrotated).
 * kv[l] = K/V tile images of layer l for this context (a3d_ctx_kv), passed as kv_base +
 *     l * kv_layer_stride_bytes.
 * w / v: packed layer weights, see act3d_chained_diffuser_b200/packing.py (pack_xattn_layer): per layer four
 *     fragment-ordered fp16 (hi, lo) [64][64] matrices {W_q, W_o, W_1, W_2} for the error-compensated
 *     tensor-core GEMMs (csrc/a3d_mma_gemm.cuh) and eight fp32 vectors {b_q, b_o, g1, be1, b_1, b_2, g2, be2}.
 *     x0 must be 8-byte aligned with even strides.
 * feat_out: NULL or [n_feat_layers][B][feat_rows][E]; rows [0,nq) written; if
 *     feat_all_layers==0 only the last layer is written (n_feat_layers = 1).
 * qvec [nqv][B][E] + logits [nqv][B][nq]: logits[j][b][n] = <qvec[j][b], x_last[b][n]>; NULL to skip.
 * (embed, heads, ffn) in {(60,4,60)}.
 */
int a3d_xattn_stack(const float* x0, long x0_stride_b, long x0_stride_n, const float* qpos,
                    int batch, int nq, int nk, int embed, int heads, int ffn, int nlayers,
                    const void* kv_base, size_t kv_layer_stride_bytes, const void* w, const float* v,
                    float* feat_out, int feat_rows, int feat_all_layers,
                    const float* qvec, int nqv, float* logits, void* stream);
size_t a3d_xattn_layer_words(int embed, int ffn);   /* 32-bit words of fragment-ordered weights per layer (w) */
size_t a3d_xattn_layer_floats(int embed, int ffn);  /* fp32 vector floats per layer (v) */

/* ---------------------------------------------------------------------------------
 * Mask logits as a separate pass: logits[j][b][n] = <qvec[j][b], feat[b][n]>  (act3d.py:493-494).
 * Used when the query-token stack runs concurrently with the ghost-point stack on another stream,
 * so that the ghost kernel cannot fuse the dot product.  feat [B][Ng][E], qvec [nqv][B][E], nqv <= 4.
 */
int a3d_mask_logits(const float* feat, const float* qvec, int batch, int ng, int embed, int nqv,
                    float* logits, void* stream);

/* ---------------------------------------------------------------------------------
 * Top ghost point.  Replaces torch.max(mask, -1).indices + position gather
 * (act3d.py:312-314, 512-513).  Lowest index wins ties.  ghost [B][Ng][3].
 */
int a3d_argmax_pick(const float* logits, const float* ghost, int batch, int ng,
                    int32_t* top_idx, float* pos, void* stream);

/* ---------------------------------------------------------------------------------
 * Ghost-point sampler.  Replaces Act3D._sample_ghost_points + the numpy samplers
 * (act3d.py:394-440, model/utils/utils.py:68-84) with a device-side Philox4x32-10
 * stream (same distributions: uniform box at level 0, uniform ball of `radius` around
 * the anchor intersected with the workspace box at level >= 1).  The anchor stays on the
 * device: no host round trip between levels.  bounds_host = {lo[3], hi[3]}.
 */
int a3d_sample_ghost(const float* anchor, float radius, const float* bounds_host, int batch, int ng,
                     uint64_t seed, uint64_t stream_id, float* out, void* stream);
/* Same sampler with the Philox stream id = stream_id + *stream_base, stream_base a DEVICE counter: a forward captured in a
 * CUDA graph draws fresh ghost points on every replay (a3d_counter_add advances the counter inside the graph). */
int a3d_sample_ghost_ctr(const float* anchor, float radius, const float* bounds_host, int batch, int ng,
                         uint64_t seed, uint64_t stream_id, const uint64_t* stream_base, float* out, void* stream);
int a3d_counter_add(uint64_t* counter, uint64_t inc, void* stream);

/* =================================================================================
 * Training path (fp32, gradients).  The reference trains both models through autograd over the eager
 * attention (multihead_custom_attention.py:157-462), which materialises the (B*H, Nq, Nk) scores, the
 * softmax and their gradients.  These entry points are the attention core and its backward without that
 * tensor; the many-row projections / LayerNorms around them are a3d_linear_fwd / a3d_linear_wgrad / a3d_layernorm_*
 * below, the rest stays torch.nn ops (act3d_chained_diffuser_b200/autograd_ops.py, train_layers.py).
 * head_dim 15 (embed == 15 * heads).
 *
 * a3d_attn_fwd: o = dropout(softmax(q k^T + key_mask)) v per head.  q / o [B][Nq][E], k / v [B][Nk][E]
 *   (head h = columns 15h .. 15h+14, the reference's head split :355-359; q already scaled by 15^-1/2 and
 *   rotated), key_mask [B][Nk] bytes (non-zero = ignore key, :398-404) or NULL, lse [B*H][Nq] =
 *   log sum exp of each score row (saved for backward).  dropout_p in [0,1) on the attention weights
 *   (:413) from a counter-based generator keyed on (seed, b, h, row, key); 0 disables it.
 * a3d_attn_bwd: gradients of the above.  dq [B][Nq][E] and dk / dv [B][Nk][E] are ACCUMULATED into (fp32 atomics
 *   over query / key chunks: zero-fill all three first); dsum is scratch of B*H*(Nq+1) floats, ZERO-FILLED
 *   by the caller ([B*H][Nq] dO . O, then [B*H] max |dO| used to rescale gradients into fp16 range).
 *   The same dropout_p / seed as the forward call regenerate the same mask.
 *   Both run on the tensor cores (error-compensated fp16 pairs on mma.sync m16n8k16, fp32 accumulation:
 *   csrc/a3d_train_mma.cu); a3d_set_option("train_attn_core", 1) selects the fp32 CUDA-core kernels instead.
 * a3d_rope_apply: out = rotary(x; pos) on channel pairs (2i, 2i+1) of the full E vector
 *   (position_encodings.py:31-34 with the 3-D table of :58-97 evaluated on the fly); x / out [rows][E],
 *   pos [rows][3].  transpose != 0 applies the inverse rotation = the backward of the forward call.
 * a3d_gather_tokens_bwd: backward of a3d_gather_tokens w.r.t. the feature map: dfeat (NCHW or channels-last
 *   like the forward's `feat`, zero-filled by the caller) += rows [0, k) of dtok [B][tok_rows][E].
 */
int a3d_attn_fwd(const float* q, const float* k, const float* v, const unsigned char* key_mask, int batch,
                 int heads, int nq, int nk, int embed, float* o, float* lse, float dropout_p, uint64_t seed,
                 void* stream);
int a3d_attn_bwd(const float* q, const float* k, const float* v, const unsigned char* key_mask, const float* o,
                 const float* dout, const float* lse, int batch, int heads, int nq, int nk, int embed, float* dq,
                 float* dk, float* dv, float* dsum, float dropout_p, uint64_t seed, void* stream);
int a3d_rope_apply(const float* x, const float* pos, long rows, int embed, int transpose, float* out,
                   void* stream);
int a3d_gather_tokens_bwd(const float* dtok, const int32_t* idx, int batch, int ncam, int embed, int hw, int k,
                          int tok_rows, int channels_last, float* dfeat, void* stream);

/* Weight / bias gradient of y = x W^T + b over `rows` tokens: dw [O][I] = dy^T x, db [O] = column sums of dy
 * (db may be NULL).  dy [rows][O], x [rows][I] row-major.  Rows are reduced in slices of 128 into `workspace`
 * (a3d_linear_wgrad_workspace bytes) and summed in a fixed order: deterministic, no atomics.
 * Replaces the single-CTA SIMT GEMMs + separate bias reductions autograd issues for the E x E projections of the
 * attention stacks when rows = batch * tokens is in the tens of thousands. */
size_t a3d_linear_wgrad_workspace(long rows, int out_features, int in_features);
int a3d_linear_wgrad(const float* dy, const float* x, long rows, int out_features, int in_features, float* dw,
                     float* db, void* workspace, void* stream);

/* The per-token linear layers and LayerNorms of the attention stacks for training (csrc/a3d_train_rows.cu), replacing
 * the fp32 SIMT GEMMs and ATen LayerNorm kernels autograd issues for them (multihead_custom_attention.py:260-303,452;
 * layers.py:300-310, 328-332, 146-147, 181-182, 205-209) when rows = batch * tokens is in the tens of thousands.
 * a3d_linear_fwd: transpose_w = 0: y [rows][O] = x [rows][I] W^T + bias [ReLU if relu], W [O][I] row-major fp32 as
 *   nn.Linear stores it (bias may be NULL); transpose_w = 1: y [rows][I] = x [rows][O] W, the data gradient of the same
 *   layer.  Tensor cores with error-compensated fp16 pairs (fp32-class accuracy).  `workspace`: a3d_linear_workspace
 *   bytes, 16-byte aligned (the weight is re-packed into MMA fragment order on every call: training weights change).
 *   a3d_linear_supported tells whether a shape is built (even feature counts, contraction length <= 480).
 * a3d_layernorm_fwd: z = x + res (res and z may be NULL: z = x), y = LayerNorm(z; eps) * gamma + beta over the last
 *   dimension (embed 60 or 120), mean / rstd [rows] saved for the backward.  a3d_layernorm_bwd: dz [rows][E] (the
 *   gradient of both x and res), dgamma, dbeta [E] via per-CTA partial sums in `workspace`
 *   (a3d_layernorm_bwd_workspace bytes) added in a fixed order: deterministic. */
int a3d_linear_supported(int out_features, int in_features, int transpose_w);
size_t a3d_linear_workspace(int out_features, int in_features, int transpose_w);
int a3d_linear_fwd(const float* x, const float* w, const float* bias, long rows, int out_features, int in_features,
                   int relu, int transpose_w, float* y, void* workspace, void* stream);
int a3d_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, long rows, int embed,
                      float eps, float* z, float* y, float* mean, float* rstd, void* stream);
size_t a3d_layernorm_bwd_workspace(int embed);
int a3d_layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                      long rows, int embed, float* dz, float* dgamma, float* dbeta, void* workspace, void* stream);

/* Position loss of the keypose trainer in one pass.  Replaces the label construction + F.cross_entropy with
 * probability targets of LossAndMetrics._compute_position_loss (main_keypose.py:387-403) for one pyramid level:
 *   label_n = softmax_n(-||ghost_n - gt|| / spread) * (1 - label_smoothing) + label_smoothing / Ng
 *   loss[b] = -sum_n label_n * log_softmax(logits[b])_n          dlogits[b][n] = softmax(logits[b])_n - label_n
 * logits [B][Ng], ghost [B][Ng][3], gt [B][3]; dlogits may be NULL (evaluation).  The mean over B and the loss
 * coefficients are applied by the caller (act3d_chained_diffuser_b200/losses.py). */
int a3d_soft_ce(const float* logits, const float* ghost, const float* gt, int batch, int ng, float spread,
                float label_smoothing, float* loss, float* dlogits, void* stream);

/* =================================================================================
 * ChainedDiffuser trajectory denoiser (embedding_dim 120, 8 heads, FFN 480, <= 64 waypoints).
 * Linear layers run as error-compensated fp16-split tensor-core GEMMs (csrc/a3d_mma_gemm.cuh):
 * every weight matrix is passed as a fragment-ordered (hi, lo) fp16 buffer ("W" packs, void*) plus
 * an fp32 vector buffer with biases / LayerNorm parameters ("V" packs).  Layouts are produced by
 * act3d_chained_diffuser_b200/packing.py (pack_lang_layer / pack_ada_layer / pack_mlp);
 * cd_pack_floats(0..3) = float count of LangV / AdaV / MlpV / one adaLN table row,
 * cd_pack_floats(4..6) = 32-bit word count of LangW / AdaW / MlpW.
 */
size_t cd_pack_floats(int which);

/* Vision -> language attention, step-invariant.  Replaces the vl_attention ParallelAttention stack
 * (diffusion_head.py:305-314; layers.py:115-218 with cross_attention1 only + FFN, no adaLN/rotary).
 * tok [B][tok_rows][E]: first nctx rows updated in place.  kin / vin [nlayers][B][n_instr][E]:
 * fp32 K and V projections of the instruction tokens.  w / v: nlayers LangW / LangV packs. */
int cd_ctx_lang(float* tok, int batch, int tok_rows, int nctx, int embed, int heads, const float* kin,
                const float* vin, int n_instr, const void* w, const float* v, int nlayers, void* stream);

/* Start of one denoiser evaluation.  Replaces traj_encoder + waypoint sinusoidal embedding +
 * traj_lang_attention (diffusion_head.py:215-216, 326-336) and prepares the fp16 rotary Q of the
 * first adaLN cross-attention layer.  traj [B][L][9] (normalised frame), wp_pe [L][E], t_idx [B]
 * timestep per sample, ada [T][ada_layers][3][2][128] adaLN (scale, shift) table; traj_enc1 = fp32
 * K-major [9][128] weight + [128] bias of the first encoder layer, traj_enc2 / traj_enc2_b = second
 * layer (fragment order / bias); lang_w / lang_v LangW / LangV or NULL (use_instruction=0);
 * lang_k / lang_vv [B][n_instr][E]; x_out [B][64][E]; next_wq / next_bq = W_q (fragment order) and
 * bias of the first cross-attention; q_out [B][H][64][16] fp16. */
int cd_step_begin(const float* traj, int batch, int length, const float* wp_pe, const int* t_idx,
                  const float* ada, int ada_layers, const float* traj_enc1, const void* traj_enc2,
                  const float* traj_enc2_b, const void* lang_w, const float* lang_v, const float* lang_k,
                  const float* lang_vv, int n_instr, float* x_out, const void* next_wq,
                  const float* next_bq, int next_ada_layer, void* q_out, void* stream);

/* Cross-attention of the waypoint tokens over the cached context K/V of one layer (tile images from
 * a3d_ctx_kv).  Replaces the cross_12 attention core of ParallelAttentionLayer (layers.py:135-145;
 * multihead_custom_attention.py:355-451).  q [B][H][64][16] fp16.  The key tiles of every (sample, head)
 * are split over 4 CTAs; `att` receives their unnormalised partial results
 * [B][4][H][17][64] = planes {O (15), denominator, row max} x 64 rows: cd_cross_part_floats(B) floats, merged by cd_post. */
size_t cd_cross_part_floats(int batch);
int cd_cross(const void* q, const void* kv, int batch, int nk, int heads, float* att, void* stream);

/* Rest of one adaLN layer.  Replaces out-proj + norm_12, the adaLN self-attention (rotary q/k,
 * key padding mask) + norm_1, the adaLN FFN + norm_122 (layers.py:146-209, 273-290), optionally a
 * regressor head (diffusion_head.py:179-198, 357-363), the Q of the next layer, and -- on the last
 * layer of a step -- the denoiser output assembly (diffusion_head.py:271-274), inpainting of the
 * conditioned waypoints and the DDPM posterior step of both schedulers (diffusion_model.py:105-117).
 * att = the partial cross-attention results written by cd_cross (merged here).
 * layer_w / layer_v = AdaW / AdaV of this layer, reg_w / reg_v = MlpW / MlpV or NULL.
 * coef_host = {c_x0, c_xt, sigma} for positions then rotations (act3d_chained_diffuser_b200/ddpm.py). */
int cd_post(const float* traj, int batch, int length, const unsigned char* mask, const float* wp_pe,
            const int* t_idx, const float* ada, int ada_layers, int ada_layer, const float* x_in,
            const float* att, const void* layer_w, const float* layer_v, float* x_out, const void* reg_w,
            const float* reg_v, float* reg_out, int reg_dim, const float* next_src, const void* next_wq,
            const float* next_bq, int next_ada_layer, void* q_out, int do_update, int last_step,
            float* traj_out, const float* pos_upd, const float* cond_data, const unsigned char* cond_mask,
            const float* coef_host, const float* noise_pos, const float* noise_rot, void* stream);

/* The whole sampling loop as ONE persistent kernel.  Replaces DiffusionPlanner.conditional_sample's loop
 * (diffusion_model.py:98-117) around DiffusionHead.forward (diffusion_head.py:200-277, 279-363) for the single-offset
 * configuration: `n_steps` denoiser evaluations + DDPM posterior steps (both schedulers, inpainting, clip_sample,
 * fixed_small variance) without leaving the SMs.  One thread-block cluster of 4 CTAs per sample (16 waypoint rows each);
 * trajectory, token tile and all intermediates live in shared memory for the whole loop; the context K/V tile images
 * are fetched once per cluster with multicast bulk copies; the self-attention K/V rows are exchanged through
 * distributed shared memory (csrc/cd_loop.cu).
 * traj [B][L][9]: in = x_T (+ conditioning), out = x_0 (normalised frame).  timesteps [n_steps] (DEVICE int32: row of
 * `ada` / `coef` used by every iteration), coef [T][6] DEVICE fp32 = {c_x0, c_xt, sigma} of the position then the
 * rotation scheduler, noise_pos [n_steps][B][L][3] / noise_rot [n_steps][B][L][6] (the last iteration's are unused).
 * ada_w_host / ada_v_host: HOST arrays of `ada_layers` device pointers (AdaW / AdaV packs: n_traj_layers shared layers,
 * then 2 position and 2 rotation layers); kv: the `ada_layers` K/V sets of a3d_ctx_kv, kv_set_bytes apart.
 * Other arguments as in cd_step_begin / cd_post. */
int cd_denoise_loop(float* traj, int batch, int length, int n_steps, const float* cond, const unsigned char* cond_mask,
                    const unsigned char* mask, const float* wp_pe, const int* timesteps, const float* ada, int ada_layers,
                    int n_traj_layers, const float* coef, const float* noise_pos, const float* noise_rot,
                    const float* traj_enc1, const void* traj_enc2, const float* traj_enc2_b, const void* lang_w,
                    const float* lang_v, const float* lang_k, const float* lang_vv, int n_instr,
                    const void* const* ada_w_host, const float* const* ada_v_host, const void* pos_reg_w,
                    const float* pos_reg_v, const void* rot_reg_w, const float* rot_reg_v, const void* kv,
                    size_t kv_set_bytes, int nk, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACT3D_B200_H_ */
