/* libfocr_sm100.so — C ABI of the B200-native FudanOCR hot path (TBSRN train/eval step).
 *
 * The reference (FudanVI/FudanOCR) has no FFI: its seam is the Python nn.Module API that
 * scene-text-telescope/interfaces/super_resolution.py:69-84 drives.  Each entry point below names the
 * reference code it replaces (paths relative to the reference root, STT = scene-text-telescope).
 *
 * Conventions
 *   - every pointer is a CUDA device pointer unless marked HOST; tensors are dense;
 *     "bf16 (T,C)" = row-major matrix of __nv_bfloat16, NHWC feature maps are (B*H*W, C) matrices
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises or allocates
 *   - return 0 on success, negative on error (focr_last_error() gives the message); there is NO CPU
 *     fallback: without an sm_100 device every compute entry point fails
 *   - workspaces are caller-allocated device memory (torch.empty), sized by the *_workspace_bytes helpers
 */
#ifndef FOCR_H_
#define FOCR_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* focr_last_error(void);
int focr_version(void);
int focr_sync_check(void* stream); /* synchronise + surface asynchronous errors (tests only) */

/* --- conv2d 3x3 / 1x1, stride 1, "same" padding, NHWC bf16, Ci and Co multiples of 64, W in {16,32,64,128}
 * replaces nn.Conv2d at STT/model/tbsrn.py:190,232,237 (64->64) and :264 (64->256, UpsampleBLock).
 * w: fp32 torch layout [Co][Ci][k][k].  flags bit0 relu; bit1 PixelShuffle(2)+mish epilogue (tbsrn.py:266-273):
 * y = pre-activation (B,2H,2W,64), y2 = mish(y).  residual (optional) is added after the activation. */
size_t focr_conv2d_workspace_bytes(int Ci, int Co, int ksize);
int focr_conv2d_fwd(const void* x, const float* w, const float* bias, void* y, void* y2, const void* residual,
                    int B, int H, int W, int Ci, int Co, int ksize, int flags, void* ws, size_t ws_bytes,
                    void* stream);
int focr_conv2d_dgrad(const void* dy, const float* w, void* dx, int B, int H, int W, int Ci, int Co, int ksize,
                      int flags, void* ws, size_t ws_bytes, void* stream);
size_t focr_wgrad_workspace_bytes(void);
int focr_conv2d_wgrad(const void* dy, const void* x, float* dw, int B, int H, int Co, int flags, void* ws,
                      size_t ws_bytes, void* stream);

/* --- nn.Linear on token matrices: STT/model/tbsrn.py:74 (128->64), :103 (4 x 128->128), :158-159 (FFN) -------
 * y[M,N] = x[M,K] w[N,K]^T + bias; flags bit0 relu, bit2 fp32 output; M % 128 == 0; K, N multiples of 64 */
size_t focr_linear_workspace_bytes(int K, int N);
int focr_linear_fwd(const void* x, const float* w, const float* bias, void* y, const void* residual, long M, int K,
                    int N, int flags, void* ws, size_t ws_bytes, void* stream);
int focr_linear_dgrad(const void* dy, const float* w, void* dx, long M, int K, int N, void* ws, size_t ws_bytes,
                      void* stream);
int focr_linear_wgrad(const void* dy, const void* x, float* dw, long M, int K, int N, void* ws, size_t ws_bytes,
                      void* stream);
int focr_bias_grad(const void* dy, float* db, long M, int N, void* ws, size_t ws_bytes, void* stream);
/* dw [N][K] and db [N] (either may be NULL) in ONE pass over dy and x: tcgen05 MN-major operands straight from the stored
 * layouts, bias gradient as an extra MMA against a ones operand (K == 128, N in {64,128,256,384}, M % 64 == 0) */
int focr_linear_wgrad_bias(const void* dy, const void* x, float* dw, float* db, long M, int K, int N, void* ws,
                           size_t ws_bytes, void* stream);

/* --- nn.BatchNorm2d in train mode + activation: STT/model/tbsrn.py:233,238,191, stn_head.py:18-21 ----------
 * act: 0 none, 1 mish (tbsrn.py:277-285), 2 relu.  stats: fp32 [4][C] = mean, invstd, scale, shift. */
size_t focr_bn_workspace_bytes(void);
int focr_bn_train_fwd(const void* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      long long* num_batches_tracked, void* y, float* stats, long T, int C, int act, void* ws,
                      size_t ws_bytes, void* stream);
int focr_bn_bwd(const void* dy, const void* x, const float* stats, void* dx, float* dgamma, float* dbeta, long T, int C,
                int act, void* ws, size_t ws_bytes, void* stream);

/* --- the reference's home-made LayerNorm, features = 128: STT/model/tbsrn.py:23-36 ---------------------------
 * y = a (x - mean) / (std_unbiased + eps) + b */
int focr_layernorm_std_fwd(const void* x, const float* a, const float* b, void* y, long T, float eps, void* stream);
int focr_layernorm_std_bwd(const void* dy, const void* x, const float* a, void* dx, float* da, float* db, long T,
                           float eps, void* ws, size_t ws_bytes, void* stream);

/* --- MultiHeadedAttention core, h = 4, d_k = 32, 1024 tokens: STT/model/tbsrn.py:109-150 ------------------------
 * qkv (B*1024,384) bf16 = [q|k|v]; out (B*1024,128); lse2 fp32 (B*4*1024).  Dropout on P with rate p_drop; the
 * keep mask is a counter-based generator keyed on (seed, stream_id, b, h, q, k).  drop_bits (device, focr_mha_drop_bits_bytes(B)
 * bytes, or NULL): when given, the forward stores its keep decisions there (1 bit per element) and the backward
 * reads them instead of re-hashing (the fast path); with NULL the backward regenerates the mask from the seed.
 * Bit layout: word [b*4+h][k/32][q] (uint32), key k of the 32-key group at bit (k%32)/4 + 8*(k&3).
 * The backward is ONE kernel (S and dP evaluated once per tile: five GEMMs, one exponential per element, dQ of all 1024
 * queries accumulating in tensor memory); ws (focr_mha_bwd_workspace_bytes(B) bytes, device) is only touched by the two-kernel
 * form kept for comparison (focr_attn_set_bwd_two_pass).
 * The rate is held to 2^-15 (p = 0.1 -> 3277/32768 = 0.100006); the 1/(1-p) rescale uses that rate, so E[out] is exact.
 * The forward shifts the softmax by a per-row upper bound of the scores (|q_i| max_j |k_j| / sqrt(d_k)) instead of the
 * row maximum and falls back to the exact two-pass route per (batch, head) when that bound is too loose for fp32. */
size_t focr_mha_drop_bits_bytes(int B);
int focr_mha_flash_fwd(const void* qkv, void* out, float* lse2, int B, float p_drop, unsigned seed, unsigned stream_id,
                       void* drop_bits, void* stream);
size_t focr_mha_bwd_workspace_bytes(int B);
int focr_mha_flash_bwd(const void* qkv, const void* out, const void* d_out, const float* lse2, void* ws, size_t ws_bytes,
                       void* dqkv, int B, float p_drop, unsigned seed, unsigned stream_id, const void* drop_bits,
                       void* stream);

/* tcgen05 operand-descriptor probe (test support for the attention kernels): copies `img` to shared memory, issues
 * nk MMAs D += A.B with the two 64-bit operand descriptors (start address relative to the image, advanced by
 * a_step16 / b_step16 16-byte units per step) and dumps the 128-lane x ncols fp32 TMEM accumulator to `out`. */
int focr_umma_probe(const void* img, int img_bytes, unsigned long long desc_a, unsigned long long desc_b, unsigned idesc,
                    int nk, unsigned a_step16, unsigned b_step16, float* out, int ncols, void* stream);

/* test support: on != 0 forces the exact (row-maximum) route of the attention forward for every (batch, head) */
int focr_attn_set_force_exact(int on);
/* test support: on != 0 selects the two-kernel attention backward (dQ pass + dK/dV pass) instead of the single-pass kernel */
int focr_attn_set_bwd_two_pass(int on);

/* --- step body: STT/interfaces/super_resolution.py:69-84, STT/loss/text_focus_loss.py:86, base.py:194-198 ------ */
int focr_mse_loss_grad(const float* sr, const float* hr, float* d_sr, float* loss, long n, float gscale, void* ws,
                       size_t ws_bytes, void* stream);
int focr_adam_clip_step(const void* chunks, int n_chunks, float gscale, float max_norm, float lr, float beta1,
                        float beta2, float eps, long long* step, float* state, void* ws, size_t ws_bytes, void* stream);

/* --- whole network: STT/model/tbsrn.py:166-226 (TBSRN), stn_head.py, tps_spatial_transformer.py -------------------
 * params / grads: HOST arrays of focr_tbsrn_num_slots() DEVICE pointers; slot i = reference state_dict entry
 * focr_tbsrn_slot_name(srb_nums, i) (fp32 tensors, int64 for num_batches_tracked).  x_lr (B,3,16,64) fp32 NCHW in,
 * sr (B,3,32,128) fp32 NCHW out.  flags bit0 training, bit1 STN present.  backward overwrites every gradient. */
int focr_tbsrn_num_slots(int srb_nums);
const char* focr_tbsrn_slot_name(int srb_nums, int idx);
size_t focr_tbsrn_workspace_bytes(int B, int srb_nums);
int focr_tbsrn_forward(void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags, float p_drop,
                       unsigned seed, void* ws, size_t ws_bytes, void* stream);
/* focr_tbsrn_forward with the dropout seed read on the device (*seed_dev, uint32) when the kernels run, so that a CUDA
 * graph captured around the whole step can be replayed with a new seed each step; masks equal those of the by-value seed.
 * The backward reads the stored keep bits / activations and needs no seed. */
int focr_tbsrn_forward_devseed(void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags,
                               float p_drop, const unsigned* seed_dev, void* ws, size_t ws_bytes, void* stream);
int focr_tbsrn_backward(void* const* params, void* const* grads, const float* x_lr, const float* d_sr, int B,
                        int srb_nums, int flags, float p_drop, unsigned seed, void* ws, size_t ws_bytes, void* stream);
/* TSRN: STT/model/tsrn.py:18-74 (TSRN), :77-98 (RecurrentResidualBlock), :128-145 (GruBlock: conv1x1 + BiGRU(64,32));
 * same conventions as the TBSRN entry points (slot i = state_dict entry focr_tsrn_slot_name(srb_nums, i)). */
int focr_tsrn_num_slots(int srb_nums);
const char* focr_tsrn_slot_name(int srb_nums, int idx);
size_t focr_tsrn_workspace_bytes(int B, int srb_nums);
int focr_tsrn_forward(void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags, void* ws,
                      size_t ws_bytes, void* stream);
int focr_tsrn_backward(void* const* params, void* const* grads, const float* x_lr, const float* d_sr, int B,
                       int srb_nums, int flags, void* ws, size_t ws_bytes, void* stream);
int focr_tsrn_ws_tensor(int B, int srb_nums, const char* name, long long* byte_offset, long long* elems,
                        int* elem_bytes);
int focr_tbsrn_ws_tensor(int B, int srb_nums, const char* name, long long* byte_offset, long long* elems,
                         int* elem_bytes);

/* --- frozen CRNN evaluator + greedy CTC decode (eval pipeline, BASELINE configs[4]) -----------------------------------
 * STT/model/crnn/crnn.py:25-80 (CRNN(32,1,37,256).forward, eval mode), STT/interfaces/base.py:319-325 (parse_crnn_data:
 * bicubic (32,100) + gray), STT/interfaces/super_resolution.py:143-158 (get_crnn_pred) and
 * STT/utils/utils_crnn.py:54-89 (strLabelConverter.decode).  params: HOST array of the 49 state_dict tensors in
 * state_dict order.  Decode output is INT32 and bit-exact: path (B,T) argmax indices (lowest index on ties), out (B,T)
 * collapsed label indices padded with -1, len (B). */
int focr_crnn_num_slots(void);
size_t focr_crnn_workspace_bytes(int B);
int focr_bicubic_gray_32x100(const float* images, float* gray, int B, void* stream);
int focr_crnn_forward(void* const* params, const float* images, int input_is_gray, float* logits, int B, void* ws,
                      size_t ws_bytes, void* stream);
int focr_ctc_greedy_decode(const float* logits, int T, int B, int C, int* path, int* out, int* len, void* stream);

/* --- CTC forward-backward over the CRNN's raw logits (north_star "CTC forward-backward"; the reference itself never calls a
 * CTC loss - SURVEY.md D2 - so the contract is torch.nn.functional.ctc_loss(log_softmax(logits, 2), ...) on the (T, B, C)
 * layout of STT/model/crnn/crnn.py:78-80).  logits fp32 (T,B,C); targets int64 (B,S_max) padded; lengths int64 (B), all on
 * the DEVICE.  reduction: 0 none, 1 mean (nll_b / max(S_b,1), averaged over B), 2 sum.  nll (B) per-sample negative
 * log-likelihood (0 for infeasible samples when zero_infinity); loss: scalar for mean / sum (may be NULL for none);
 * d_logits (T,B,C) = grad_scale * d loss / d logits (for `none`: grad_scale * d nll_b / d logits), or NULL to skip the
 * backward half.  focr_ctc_loss_status copies the status word of the last call (0 ok, 1 length out of range, 2 label outside
 * [0, C)) to the host and synchronises the stream. */
size_t focr_ctc_loss_workspace_bytes(int T, int B, int S_max);
int focr_ctc_loss(const float* logits, int T, int B, int C, const long long* targets, int S_max, const long long* input_lengths,
                  const long long* target_lengths, int blank, int reduction, int zero_infinity, float grad_scale, float* nll,
                  float* loss, float* d_logits, void* ws, size_t ws_bytes, void* stream);
int focr_ctc_loss_status(const void* ws, int T, int B, int S_max, int* status_host, void* stream);

/* --- stroke-/text-focus loss: the frozen recogniser and the attention-map L1 term -------------------------------------
 * text-gestalt/loss/stroke_focus_loss.py:83-122 (StrokeFocusLoss.forward), :12-18 (to_gray_tensor),
 * text-gestalt/loss/transformer_english_decomposition.py:70-168 (ResNet encoder), :276-304 (Decoder), :343-398
 * (Transformer.forward); scene-text-telescope/loss/transformer.py is the same network with 37 classes.
 * params: HOST array of focr_strokenet_num_slots() DEVICE pointers, slot i = state_dict entry
 * focr_strokenet_slot_name(variant, i) (variant 0 text-gestalt key names, 1 scene-text-telescope).  The recogniser is
 * frozen and in eval mode: focr_strokenet_prepare folds BatchNorm into bf16 conv weights once, into a caller-allocated
 * blob of focr_strokenet_prepared_bytes(n_class) bytes that every later call borrows.
 * focr_focus_loss: sr, hr fp32 NCHW (B,3,32,128); text_input int64 (B,T) = the right-shifted label indices the
 * reference's label encoder builds; losses[3] (device) = {mse + lambda * attention, mse, attention};
 * d_sr = gscale * d(loss)/d(sr) (MSE and attention terms; gscale = 100 in the reference step body);
 * map_hr_out / map_sr_out (optional, fp32 (B,16,T,256)) receive the two word-attention maps.  Only the SR branch's
 * input-gradient chain is evaluated: the reference's unused weight gradients and HR-branch backward are not. */
int focr_strokenet_num_slots(void);
const char* focr_strokenet_slot_name(int variant, int idx);
size_t focr_strokenet_prepared_bytes(int n_class);
int focr_strokenet_prepare(void* const* params, int n_class, void* prepared, size_t prepared_bytes, void* stream);
size_t focr_focus_loss_workspace_bytes(int B, int T);
int focr_focus_loss(const void* prepared, size_t prepared_bytes, int n_class, const float* sr, const float* hr,
                    const long long* text_input, int B, int T, float lambda, float gscale, float* d_sr, float* losses,
                    float* map_hr_out, float* map_sr_out, void* ws, size_t ws_bytes, void* stream);
/* TextFocusLoss.forward (scene-text-telescope/loss/text_focus_loss.py:84-99) = mse + lambda_attn * L1(maps) + lambda_ce *
 * weight_cross_entropy(sr logits, text_gt) (scene-text-telescope/loss/weight_ce_loss.py:36-45; weight_table fp32
 * (n_class, n_class) from load_confuse_matrix :10-33).  length (B) and text_gt (sum length) are the int64 tensors of the
 * reference's label_encoder (:62-81).  losses[4] = {total, mse, attention, recognition}; sr_pred_out (optional, fp32
 * (sum length, n_class)) receives the packed SR logits.  The full decoder (value path, FFN, LayerNorms, generator) runs for
 * the SR branch only - the HR branch's logits are never read by the reference either. */
int focr_text_focus_loss(const void* prepared, size_t prepared_bytes, int n_class, const float* sr, const float* hr,
                         const long long* text_input, const long long* length, const long long* text_gt,
                         const float* weight_table, int B, int T, float lambda_attn, float lambda_ce, float gscale,
                         float* d_sr, float* losses, float* map_hr_out, float* map_sr_out, float* sr_pred_out, void* ws,
                         size_t ws_bytes, void* stream);
int focr_focus_loss_ws_tensor(int B, int T, const char* name, long long* byte_offset, long long* elems, int* elem_bytes);

/* Stand-alone forms of the two helpers the reference exports next to its focus losses (code importing them keeps working on
 * CUDA tensors; the fused losses above carry their own copies):
 * weight_cross_entropy(pred, gt) scene-text-telescope/loss/weight_ce_loss.py:36-45: pred fp32 (N, C) logits, gt int64 (N),
 * table fp32 (C, C) = load_confuse_matrix(); loss[0] = -(1/N) sum_i log(w[g_i][g_i] e^{p_i,g_i} / sum_j w[g_i][j] e^{p_ij}) by
 * log-sum-exp; d_pred (optional, fp32 (N, C)) = d loss / d pred; status (optional, int32[1]) is set to 1 when a gt index lies
 * outside [0, C) (torch indexing raises there).
 * to_gray_tensor(t) scene-text-telescope/loss/text_focus_loss.py:16-21 (text-gestalt/loss/stroke_focus_loss.py:12-18):
 * img fp32 NCHW (B, C >= 3, H, W) -> gray (B, 1, H, W) = 0.299 R + 0.587 G + 0.114 B; _bwd writes d_img (channels >= 3: zero). */
size_t focr_weight_cross_entropy_workspace_bytes(long N);
int focr_weight_cross_entropy(const float* pred, const long long* gt, const float* table, float* loss, float* d_pred,
                              int* status, long N, int C, void* ws, size_t ws_bytes, void* stream);
int focr_to_gray(const float* img, float* gray, long B, int C, long HW, void* stream);
int focr_to_gray_bwd(const float* d_gray, float* d_img, long B, int C, long HW, void* stream);

/* --- KV-cached greedy test-time decode of the ResNet + Transformer recognisers ------------------------------
 * replaces the loops stroke-level-decomposition/train.py:110-121 and image-ids-CTR/train.py:118-134, which re-run the whole
 * decoder (SLD/model/transformer.py:303-317) on the growing prefix for each of max_length steps.  Here the cross-attention K / V
 * of the image tokens are projected once, self-attention K / V are cached per emitted position, and every step runs the decoder
 * layer on the ONE new position, arg-max (lowest index on ties) and its softmax probability included; nothing visits the host
 * until the (B, T_max + 1) token matrix is read back.
 * params: HOST array of 29 DEVICE fp32 pointers - 0 embedding table (vocab,512); 1-8 masked self-attention linears q,k,v,out
 * (weight (1024,1024), bias each); 9,10 LayerNorm-1 scale, shift; 11-18 cross-attention linears; 19,20 LayerNorm-2; 21-24 FFN
 * w_1 (2048,1024), b_1, w_2 (1024,2048), b_2; 25,26 LayerNorm-3; 27,28 generator weight (n_out,1024), bias.
 * text_features: NULL (SLD: the generator scores are the class scores) or fp32 (n_feat, n_out) (IDS: the generator output is
 * L2-normalised and matched against them).  feat: bf16, B*n_tok rows zero-padded to a multiple of 128, 1024 columns (the encoder's
 * NHWC map).  pred int64 (B, T_max+1): column 0 = start symbol 0; prob fp32 (B, T_max) = winning softmax probability per step. */
size_t focr_recog_decode_prepared_bytes(int vocab, int n_out, int n_feat);
int focr_recog_decode_prepare(void* const* params, int vocab, int n_out, const float* text_features, int n_feat, void* blob,
                              size_t blob_bytes, void* stream);
size_t focr_recog_decode_workspace_bytes(int B, int n_tok, int T_max, int n_out, int n_feat);
int focr_recog_decode(const void* blob, size_t blob_bytes, int vocab, int n_out, int n_feat, const void* feat, int B, int n_tok,
                      int T_max, long long* pred, float* prob, void* ws, size_t ws_bytes, void* stream);

/* --- evaluation metrics: scene-text-telescope/utils/ssim_psnr.py:9-15 (calculate_psnr), :31-78 (SSIM, window 11, sigma 1.5),
 * as called per validation batch by interfaces/super_resolution.py:191-192.  img1, img2 fp32 NCHW (B, channels >= 3, 32, 128)
 * in [0,1] (first 3 channels are used); window: the 121 fp32 taps of create_window (:24-28).  out[0] = PSNR over the batch
 * (inf when identical), out[1] = mean SSIM; ssim_per_image (optional, B floats) = the size_average=False form. */
size_t focr_psnr_ssim_workspace_bytes(int B);
int focr_psnr_ssim(const float* img1, const float* img2, int B, int channels, const float* window, float* out,
                   float* ssim_per_image, void* ws, size_t ws_bytes, void* stream);

/* --- input pipeline, device side: resizeNormalize = img.resize((W,H), Image.BICUBIC) + ToTensor, applied per crop by
 * alignCollate_real (scene-text-telescope/dataset/dataset.py:136-152, :258-270).  Bit-exact with Pillow's 8-bit resampler
 * (antialiased bicubic a = -0.5, 22-bit fixed-point coefficients, uint8 intermediate).  pixels: packed uint8 RGB crops
 * (h x w x 3 each); meta: DEVICE int64 [B][3] = {byte offset, h, w}; max_h / max_w: HOST upper bounds of the crop sizes;
 * out: fp32 (B,3,out_h,out_w) in [0,1]; status: device int set to 1 if a crop exceeded the bounds. */
int focr_resize_bicubic_normalize(const void* pixels, const long long* meta, int B, int max_h, int max_w, int out_w, int out_h,
                                  float* out, int* status, void* stream);

/* --- building blocks of the TRAINABLE ResNet + Transformer recognisers (SLD = stroke-level-decomposition,
 * IDS = image-ids-CTR).  All matrices bf16 row-major unless stated; every reduction in a fixed order (bit-reproducible). ---------
 * focr_mha_small_*: attention() SLD/model/transformer.py:227-241 inside MultiHeadedAttention.forward :205-223, as used by
 *   Decoder.forward :303-315 (masked self-attention: causal = 1; cross-attention to the image tokens: causal = 0), h heads of
 *   d_k in {64,128,256}.  q (B*Tq, ld_q), k / v (B*Tk, ld_k / ld_v), out (B*Tq, ld_o); head i at columns [i*d_k, (i+1)*d_k).
 *   map fp32 (B,H,Tq,Tk) = dropout(softmax(q k^T / sqrt(d_k))) - the 'map' the reference returns; the backward recomputes the
 *   softmax and reads the keep mask from map's zeros.  Tq*Tk*8 bytes must fit shared memory (200 KB).
 * focr_layernorm_wide_*: LayerNorm SLD/model/transformer.py:244-254 for 512 / 1024 features, y = LN(x [+ res]); sum_out (optional)
 *   receives x + res, which the backward takes as its x.  da / db are overwritten.
 * focr_text_embed_*: Embeddings :277-286 (* sqrt(E)) concatenated with PositionalEncoding :168-186 applied to zeros (:346-348):
 *   out (rows_pad, 2E), rows = B*T; p_drop acts on the positional half only, as in the reference.
 * focr_packed_ce: the packing loop :361-373 + nn.CrossEntropyLoss (SLD/train.py:41,71): mean over the positions t < length[b] of
 *   CE(logits[b,t,:C], gt_packed); logits fp32 (B*T, ld); d_logits bf16 (B*T, ld_d) = gscale * d loss / d logits or NULL.
 * focr_adadelta_step: torch.optim.Adadelta (SLD/train.py:32-36: lr 1, rho 0.9, eps 1e-6; IDS adds weight_decay 1e-4) over a
 *   device table of {param, grad, square_avg, acc_delta, n} records (5 x int64 each); grad is scaled by gscale first.
 * focr_conv3x3_gemm_*: nn.Conv2d(3x3, pad 1) through an explicit im2col: forward for inputs the implicit GEMM does not take
 *   (3-channel NCHW fp32 image, SLD/model/transformer.py:80), and the weight + bias gradient of ANY 3x3 layer (dW = dY^T col).
 *   Exactly one of x_nhwc (bf16) / x_nchw (fp32) is non-NULL; Co % 64 == 0; forward needs B*H*W % 128 == 0.
 * focr_add_relu / focr_relu_bwd: BasicBlock tail :69-71 (out += residual; relu) and its gradient from the stored output.
 * focr_maxpool2x2_*: nn.MaxPool2d((2,2),(2,2)) :84,130 on NHWC maps; focr_dropout: nn.Dropout with a counter-hash mask - calling
 *   it on the gradient with the same (seed, stream_id) is its backward. */
/* Contrastive head of the CCR-CLIP pre-training stage - image-ids-CTR/CCR-CLIP/model.py:209-222 (L2-normalise both feature sets,
 * logit_scale.exp()) and main.py:98-110 (logits_per_image = scale * I T^T, logits_per_text its transpose, loss = (CE(logits_per_image,
 * gt) + CE(logits_per_text, gt)) / 2) - value and gradients in one call.  image / text: fp32 (B, D) UN-normalised tower outputs (on a
 * data-parallel run: the all-gathered global batch), logit_scale: the log-parameter (1 float, device), gt: int64 (B), gt[i] = first
 * sample with sample i's label (main.py:103-105).  d_image / d_text (B, D) + d_logit_scale may be NULL.  *status (optional) is set
 * when a target is outside [0, B).  The towers themselves are not part of this library (out of scope). */
size_t focr_clip_contrastive_workspace_bytes(int B, int D);
int focr_clip_contrastive_loss(const float* image, const float* text, const float* logit_scale, const long long* gt, int B, int D,
                               float* loss, float* d_image, float* d_text, float* d_logit_scale, int* status, void* ws,
                               size_t ws_bytes, void* stream);
/* dropout epoch: a device word XOR-ed into every dropout key of the recogniser kernels (attention map, positional-encoding and
 * FFN dropout).  0 = keys exactly as passed (default).  A step replayed as a CUDA graph freezes its seed arguments; the trainer
 * captures focr_recog_epoch_advance at the head of the step so that every replay draws fresh masks (nn.Dropout's behaviour). */
int focr_recog_epoch_set(unsigned value, void* stream);
int focr_recog_epoch_advance(void* stream);
int focr_mha_small_fwd(const void* q, long ld_q, const void* k, long ld_k, const void* v, long ld_v, void* out, long ld_o,
                       float* map, int B, int H, int d_k, int Tq, int Tk, int causal, float p_drop, unsigned seed,
                       unsigned stream_id, void* stream);
int focr_mha_small_bwd(const void* q, long ld_q, const void* k, long ld_k, const void* v, long ld_v, const void* d_out, long ld_o,
                       const float* map, void* dq, long ld_dq, void* dk, long ld_dk, void* dv, long ld_dv, int B, int H, int d_k,
                       int Tq, int Tk, int causal, float p_drop, void* stream);
/* key counts whose (Tq, Tk) score tile exceeds one CTA's shared memory (2 560 image tokens of 32 x 320 crops): the backward
 * splits into a row pass and a key pass and needs a (B, H, Tq, Tk) fp32 workspace; ..._workspace_bytes is 0 otherwise */
size_t focr_mha_small_bwd_workspace_bytes(int B, int H, int Tq, int Tk);
int focr_mha_small_bwd_ws(const void* q, long ld_q, const void* k, long ld_k, const void* v, long ld_v, const void* d_out, long ld_o,
                          const float* map, void* dq, long ld_dq, void* dk, long ld_dk, void* dv, long ld_dv, int B, int H, int d_k,
                          int Tq, int Tk, int causal, float p_drop, void* ws, size_t ws_bytes, void* stream);
int focr_layernorm_wide_fwd(const void* x, const void* res, const float* a, const float* b, void* sum_out, void* y, long T, int C,
                            float eps, void* stream);
size_t focr_layernorm_wide_workspace_bytes(int C);
int focr_layernorm_wide_bwd(const void* dy, const void* x, const float* a, void* dx, float* da, float* db, long T, int C, float eps,
                            void* ws, size_t ws_bytes, void* stream);
int focr_text_embed_fwd(const long long* idx, const float* lut, int vocab, int E, int B, int T, long rows_pad, void* out,
                        float p_drop, unsigned seed, unsigned stream_id, int* status, void* stream);
int focr_text_embed_bwd(const long long* idx, const void* d_out, int vocab, int E, int B, int T, float* d_lut, void* stream);
size_t focr_packed_ce_workspace_bytes(int B);
int focr_packed_ce(const float* logits, long ld, int B, int T, int C, const long long* length, const long long* gt, float gscale,
                   float* loss, void* d_logits, long ld_d, void* ws, size_t ws_bytes, void* stream);
int focr_dropout(const void* x, void* y, long n, float p_drop, unsigned seed, unsigned stream_id, void* stream);
int focr_add_relu(const void* a, const void* b, void* y, long n, void* stream);
int focr_relu_bwd(const void* dy, const void* y, void* dx, long n, void* stream);
int focr_maxpool2x2_fwd(const void* x, void* y, int B, int H, int W, int C, void* stream);
int focr_maxpool2x2_bwd(const void* x, const void* y, const void* dy, void* dx, int B, int H, int W, int C, void* stream);
int focr_adadelta_step(const void* chunks, int n_chunks, float gscale, float lr, float rho, float eps, float weight_decay,
                       void* stream);
/* weight + bias gradient of a 3x3 convolution on the tcgen05 GEMM: dW = dY^T . col with the pixel dimension as the contraction, both
 * operands transposed to K-major form first (Co % 128 == 0, B*H*W % 128 == 0); same arguments as focr_conv3x3_gemm_wgrad */
size_t focr_conv3x3_wgrad_tc_workspace_bytes(int B, int H, int W, int Ci, int Co);
int focr_conv3x3_wgrad_tc(const void* dy, const void* x_nhwc, const float* x_nchw, float* dw, float* db, int B, int H, int W, int Ci,
                          int Co, void* ws, size_t ws_bytes, void* stream);
/* nn.BatchNorm2d under model.eval() (running statistics) + activation (0 none, 2 relu); stats: fp32 [4][C] scratch */
int focr_bn_eval_fwd(const void* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     void* y, float* stats, long T, int C, int act, void* stream);
/* image-ids-CTR head (IDS/train.py:66-80): row-wise L2 normalisation of the 2048-d predictions (x fp32 (T, ld) -> y bf16 (T, C),
 * inv fp32 (T) = 1 / norm) with its backward dx = inv (dy - y (dy . y)), and nn.MSELoss(pred_n, text_features[text_gt]) over the
 * packed valid rows (value, and gscale * gradient w.r.t. y; zeros at t >= length[b]).  ws as focr_packed_ce. */
int focr_l2norm_rows_fwd(const float* x, long ld, void* y, float* inv, long T, int C, void* stream);
int focr_l2norm_rows_bwd(const void* dy, const void* y, const float* inv, void* dx, long ld_dx, long T, int C, void* stream);
int focr_packed_feat_mse(const void* y, int B, int T, int C, const long long* length, const long long* gt, const float* feats, int V,
                         float gscale, float* loss, void* d_y, void* ws, size_t ws_bytes, void* stream);
size_t focr_conv3x3_gemm_workspace_bytes(int B, int H, int W, int Ci, int Co);
int focr_conv3x3_gemm_fwd(const void* x_nhwc, const float* x_nchw, const float* w, const float* bias, void* y, int B, int H, int W,
                          int Ci, int Co, void* ws, size_t ws_bytes, void* stream);
int focr_conv3x3_gemm_wgrad(const void* dy, const void* x_nhwc, const float* x_nchw, float* dw, float* db, int B, int H, int W,
                            int Ci, int Co, void* ws, size_t ws_bytes, void* stream);

/* --- measurement hooks used by bench.py: CUDA-event scopes on the launching stream + launch counter ----------- */
int focr_prof_enable(int mode /*0 off, 1 all, 2 focus*/, const char* focus_substring);
int focr_prof_collect(char* buf, int cap); /* lines "scope launches total_ms"; synchronises; clears */
long long focr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FOCR_H_ */
