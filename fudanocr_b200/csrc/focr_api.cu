// C-ABI entry points of libfocr_sm100.so (declared in include/focr.h).
// Plain pointers and sizes only; every call enqueues work on the caller's stream and returns
// 0 or a negative error code (message via focr_last_error()).  No CPU fallback exists: on a
// machine without an sm_100 device every compute entry point fails with FOCR_ERR_CUDA.
#include <stdarg.h>
#include <string.h>

#include "kernels.cuh"
#include "tbsrn_engine.cuh"

static thread_local char g_err[512] = "";

void focr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" {

const char* focr_last_error(void) { return g_err; }
int focr_version(void) { return 100; }

int focr_sync_check(void* stream) {
  FOCR_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  FOCR_CHECK_CUDA(cudaGetLastError());
  return FOCR_OK;
}

// ---------------------------------------------------------------------------------------------
// conv2d (3x3 pad 1 or 1x1) on an NHWC bf16 feature map with C_in, C_out multiples of 64 and
// W in {64,128}.  w: fp32 [Co][Ci][k][k] (torch layout).  flags bit0: relu, bit1: PixelShuffle(2)
// epilogue (Co == 256; y = pre-activation (B,2H,2W,64), y2 = mish(y)).
// ---------------------------------------------------------------------------------------------
size_t focr_conv2d_workspace_bytes(int Ci, int Co, int ksize) {
  return align_up((size_t)ksize * ksize * Ci * Co * 2, 256) + align_up((size_t)Co * 4, 256);
}

int focr_conv2d_fwd(const void* x, const float* w, const float* bias, void* y, void* y2,
                    const void* residual, int B, int H, int W, int Ci, int Co, int ksize, int flags,
                    void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(ws_bytes >= focr_conv2d_workspace_bytes(Ci, Co, ksize), "conv2d_fwd: workspace too small");
  const int shuf = (flags >> 1) & 1;
  bf16* wb = (bf16*)ws;
  float* bp = (float*)((char*)ws + align_up((size_t)ksize * ksize * Ci * Co * 2, 256));
  int rc = prep_conv_w_fwd(w, wb, Co, Ci, ksize, shuf, s);
  if (rc) return rc;
  const float* bias_use = bias;
  if (bias && shuf) {
    rc = prep_bias_shuf(bias, bp, Co, s);
    if (rc) return rc;
    bias_use = bp;
  }
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = Co;
  p.kh = p.kw = ksize;
  p.W = W;
  p.H = H;
  p.epi = shuf ? TC_EPI_PIXSHUF : TC_EPI_BF16;
  p.relu = flags & 1;
  p.ldc = Co;
  p.bias = bias_use;
  p.out = y;
  p.out2 = y2;
  p.residual = (const bf16*)residual;
  const bf16* ap[1] = {(const bf16*)x};
  return tc_gemm_launch(ap, 1, Ci, (long)W * Ci, (long)H * W * Ci, Ci, B, wb, Ci, p, s);
}

// Input gradient of the conv above.  Plain: dy (B,H,W,Co) -> dx (B,H,W,Ci).
// flags bit1 (PixelShuffle variant): dy is the gradient w.r.t. the shuffled pre-activation,
// laid out (B,2H,2W,64), Co == 256.
int focr_conv2d_dgrad(const void* dy, const float* w, void* dx, int B, int H, int W, int Ci, int Co,
                      int ksize, int flags, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(ws_bytes >= focr_conv2d_workspace_bytes(Ci, Co, ksize), "conv2d_dgrad: workspace too small");
  const int shuf = (flags >> 1) & 1;
  bf16* wb = (bf16*)ws;
  int rc = prep_conv_w_dgrad(w, wb, Co, Ci, ksize, shuf, s);
  if (rc) return rc;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = Ci;
  p.kh = p.kw = ksize;
  p.W = W;
  p.H = H;
  p.epi = TC_EPI_BF16;
  p.ldc = Ci;
  p.out = dx;
  if (!shuf) {
    const bf16* ap[1] = {(const bf16*)dy};
    return tc_gemm_launch(ap, 1, Co, (long)W * Co, (long)H * W * Co, Co, B, wb, Co, p, s);
  }
  FOCR_REQUIRE(Co == 256, "conv2d_dgrad: PixelShuffle variant needs Co == 256");
  // four strided views of the (B,2H,2W,64) gradient, one per sub-pixel (i,j)
  const bf16* base = (const bf16*)dy;
  const bf16* ap[4];
  for (int sub = 0; sub < 4; ++sub) ap[sub] = base + ((long)(sub >> 1) * 2 * W + (sub & 1)) * 64;
  return tc_gemm_launch(ap, 4, 2 * 64, (long)2 * 2 * W * 64, (long)2 * H * 2 * W * 64, 64, B, wb, Co, p, s);
}

// ---------------------------------------------------------------------------------------------
// nn.Linear on a token matrix: y[M,N] = x[M,K] w[N,K]^T + bias (+relu) (+residual); M % 128 == 0,
// K, N multiples of 64.  dgrad: dx[M,K] = dy[M,N] w[N,K].
// ---------------------------------------------------------------------------------------------
size_t focr_linear_workspace_bytes(int K, int N) { return align_up((size_t)K * N * 2, 256); }

int focr_linear_fwd(const void* x, const float* w, const float* bias, void* y, const void* residual,
                    long M, int K, int N, int flags, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(M % 128 == 0, "linear_fwd: M %% 128 != 0");
  FOCR_REQUIRE(ws_bytes >= focr_linear_workspace_bytes(K, N), "linear_fwd: workspace too small");
  bf16* wb = (bf16*)ws;
  int rc = prep_linear_w(w, wb, nullptr, N, K, 0, 0, s);
  if (rc) return rc;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = N;
  p.kh = p.kw = 1;
  p.W = 64;
  p.H = 2;
  p.epi = (flags & 4) ? TC_EPI_F32 : TC_EPI_BF16;
  p.relu = flags & 1;
  p.ldc = N;
  p.bias = bias;
  p.out = y;
  p.residual = (const bf16*)residual;
  const bf16* ap[1] = {(const bf16*)x};
  return tc_gemm_launch(ap, 1, K, (long)64 * K, (long)128 * K, K, (int)(M / 128), wb, K, p, s);
}

int focr_linear_dgrad(const void* dy, const float* w, void* dx, long M, int K, int N, void* ws,
                      size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(M % 128 == 0, "linear_dgrad: M %% 128 != 0");
  FOCR_REQUIRE(ws_bytes >= focr_linear_workspace_bytes(K, N), "linear_dgrad: workspace too small");
  bf16* wt = (bf16*)ws;
  int rc = prep_linear_w(w, nullptr, wt, N, K, N, 0, s);  // wt: [K][N]
  if (rc) return rc;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = K;
  p.kh = p.kw = 1;
  p.W = 64;
  p.H = 2;
  p.epi = TC_EPI_BF16;
  p.ldc = K;
  p.out = dx;
  const bf16* ap[1] = {(const bf16*)dy};
  return tc_gemm_launch(ap, 1, N, (long)64 * N, (long)128 * N, N, (int)(M / 128), wt, N, p, s);
}


// ---------------------------------------------------------------------------------------------
// weight gradients
// ---------------------------------------------------------------------------------------------
size_t focr_wgrad_workspace_bytes(void) { return (size_t)48 << 20; }

// dW[N][K] (fp32) = dy[M,N]^T x[M,K]; K % 128 == 0, N % 64 == 0
int focr_linear_wgrad(const void* dy, const void* x, float* dw, long M, int K, int N, void* ws, size_t ws_bytes,
                      void* stream) {
  FOCR_REQUIRE(ws_bytes >= focr_wgrad_workspace_bytes(), "linear_wgrad: workspace too small");
  return linear_wgrad((const bf16*)dy, N, (const bf16*)x, K, M, N, K, dw, 1.f, (float*)ws, (cudaStream_t)stream);
}
// weight AND bias gradient in one pass over dY and X (tcgen05 kernel of wgrad_tc.cu when K == 128, N in {64,128,256,384} and
// M % 64 == 0; the streaming mma.sync kernels otherwise)
int focr_linear_wgrad_bias(const void* dy, const void* x, float* dw, float* db, long M, int K, int N, void* ws, size_t ws_bytes,
                           void* stream) {
  FOCR_REQUIRE(ws_bytes >= focr_wgrad_workspace_bytes(), "linear_wgrad_bias: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (linear_wgrad_tc_supported(M, N, K, N, K) && linear_wgrad_tc_partial_bytes(N) <= ws_bytes)
    return linear_wgrad_tc((const bf16*)dy, N, (const bf16*)x, K, M, N, dw, db, (float*)ws, s);
  if (dw) {
    int rc = linear_wgrad((const bf16*)dy, N, (const bf16*)x, K, M, N, K, dw, 1.f, (float*)ws, s);
    if (rc) return rc;
  }
  if (db) return colsum((const bf16*)dy, N, M, N, db, (float*)ws, s);
  return FOCR_OK;
}
// column sums (bias gradient): out[N] = sum_m x[m][n]
int focr_bias_grad(const void* dy, float* db, long M, int N, void* ws, size_t ws_bytes, void* stream) {
  FOCR_REQUIRE(ws_bytes >= focr_wgrad_workspace_bytes(), "bias_grad: workspace too small");
  return colsum((const bf16*)dy, N, M, N, db, (float*)ws, (cudaStream_t)stream);
}
// conv 3x3 (Ci = 64) weight gradient, torch layout [Co][64][3][3]; flags bit1: PixelShuffle variant
int focr_conv2d_wgrad(const void* dy, const void* x, float* dw, int B, int H, int Co, int flags, void* ws,
                      size_t ws_bytes, void* stream) {
  FOCR_REQUIRE(ws_bytes >= focr_wgrad_workspace_bytes(), "conv2d_wgrad: workspace too small");
  return conv3x3_wgrad((const bf16*)dy, (const bf16*)x, B, H, Co, (flags >> 1) & 1, dw, (float*)ws,
                       (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm (train mode) over a (T, C) bf16 matrix;  act: 0 none, 1 mish, 2 relu
// stats: fp32 [4][C] = mean, invstd, scale, shift (produced by fwd, consumed by bwd)
// ---------------------------------------------------------------------------------------------
size_t focr_bn_workspace_bytes(void) { return (size_t)16 << 20; }

int focr_bn_train_fwd(const void* x, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, long long* num_batches_tracked, void* y, float* stats, long T, int C,
                      int act, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(ws_bytes >= focr_bn_workspace_bytes(), "bn_train_fwd: workspace too small");
  int rc = bn_train_stats((const bf16*)x, C, T, C, gamma, beta, running_mean, running_var, num_batches_tracked,
                          1e-5f, 0.1f, (float*)ws, stats, s);
  if (rc) return rc;
  return bn_apply((const bf16*)x, C, stats, (bf16*)y, C, T, C, act, nullptr, 0, nullptr, s);
}
int focr_bn_bwd(const void* dy, const void* x, const float* stats, void* dx, float* dgamma, float* dbeta, long T,
                int C, int act, void* ws, size_t ws_bytes, void* stream) {
  FOCR_REQUIRE(ws_bytes >= focr_bn_workspace_bytes(), "bn_bwd: workspace too small");
  float* partial = (float*)ws;
  float* coef = partial + ((size_t)12 << 20) / 4;
  return bn_backward((const bf16*)dy, C, (const bf16*)x, C, stats, (bf16*)dx, C, T, C, act, dgamma, dbeta, partial,
                     coef, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// the reference's LayerNorm (features = 128): y = a (x-mean)/(std_unbiased + eps) + b
// ---------------------------------------------------------------------------------------------
int focr_layernorm_std_fwd(const void* x, const float* a, const float* b, void* y, long T, float eps, void* stream) {
  return ln_forward((const bf16*)x, a, b, (bf16*)y, T, eps, (cudaStream_t)stream);
}
int focr_layernorm_std_bwd(const void* dy, const void* x, const float* a, void* dx, float* da, float* db, long T,
                           float eps, void* ws, size_t ws_bytes, void* stream) {
  FOCR_REQUIRE(ws_bytes >= focr_bn_workspace_bytes(), "layernorm_bwd: workspace too small");
  return ln_backward((const bf16*)dy, (const bf16*)x, a, (bf16*)dx, da, db, (float*)ws, T, eps,
                     (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// fused 4-head self-attention over 1024 tokens, d_k = 32.  qkv (B*1024, 384) bf16, out (B*1024,128) bf16,
// lse2 fp32 (B*4*1024) (log2-domain log-sum-exp, needed by bwd).  p_drop in [0,1): dropout on P.
// ---------------------------------------------------------------------------------------------
size_t focr_mha_drop_bits_bytes(int B) { return attn_drop_bits_bytes(B); }
int focr_mha_flash_fwd(const void* qkv, void* out, float* lse2, int B, float p_drop, unsigned seed, unsigned stream_id,
                       void* drop_bits, void* stream) {
  const uint32_t th = p_drop > 0.f ? (uint32_t)(p_drop * 65536.0 + 0.5) : 0;
  return attn_forward((const bf16*)qkv, (bf16*)out, lse2, B, drop_key(seed, stream_id), th, (uint32_t*)drop_bits,
                      (cudaStream_t)stream);
}
// ws: focr_mha_bwd_workspace_bytes(B) bytes (D = rowsum(dO o O) per (b, h, q); only the two-kernel form uses it)
size_t focr_mha_bwd_workspace_bytes(int B) { return ((size_t)B * 4 * 1024 * sizeof(float) + 255) & ~(size_t)255; }
int focr_mha_flash_bwd(const void* qkv, const void* out, const void* d_out, const float* lse2, void* ws, size_t ws_bytes,
                       void* dqkv, int B, float p_drop, unsigned seed, unsigned stream_id, const void* drop_bits,
                       void* stream) {
  const uint32_t th = p_drop > 0.f ? (uint32_t)(p_drop * 65536.0 + 0.5) : 0;
  FOCR_REQUIRE(ws != nullptr && ws_bytes >= focr_mha_bwd_workspace_bytes(B), "mha_flash_bwd: workspace of %zu bytes, need %zu",
               ws_bytes, focr_mha_bwd_workspace_bytes(B));
  return attn_backward((const bf16*)qkv, (const bf16*)out, (const bf16*)d_out, lse2, (float*)ws, (bf16*)dqkv, B,
                       drop_key(seed, stream_id), th, (const uint32_t*)drop_bits, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// loss head and optimizer of the reference step body (interfaces/super_resolution.py:69-84)
// ---------------------------------------------------------------------------------------------
// loss = mean((sr-hr)^2) ; d_sr = gscale * 2 (sr-hr) / n      (gscale = 100: "loss_im = loss * 100")
int focr_mse_loss_grad(const float* sr, const float* hr, float* d_sr, float* loss, long n, float gscale, void* ws,
                       size_t ws_bytes, void* stream);
// chunk table: device array of n_chunks records {param*, grad*, exp_avg*, exp_avg_sq*, int64 length};
// state: 4 floats out (grad norm, clip coef * gscale, lr/(1-b1^t), 1/sqrt(1-b2^t)); step: device int64
int focr_adam_clip_step(const void* chunks, int n_chunks, float gscale, float max_norm, float lr, float beta1,
                        float beta2, float eps, long long* step, float* state, void* ws, size_t ws_bytes,
                        void* stream) {
  FOCR_REQUIRE(ws_bytes >= (size_t)n_chunks * 4, "adam_clip_step: workspace too small");
  return adam_clip_step(chunks, n_chunks, gscale, max_norm, lr, beta1, beta2, eps, step, state, (float*)ws,
                        (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// TBSRN engine (model/tbsrn.py TBSRN.forward + its autograd).  params / grads: HOST arrays with one DEVICE
// pointer per slot (slot i is the reference state_dict entry focr_tbsrn_slot_name(i)).
// flags bit0: training (BN batch stats + running update, dropout, STN active), bit1: model has STN.
// ---------------------------------------------------------------------------------------------
int focr_tbsrn_num_slots(int srb_nums) { return tbsrn::Slots(srb_nums).count; }
int focr_tsrn_num_slots(int srb_nums) { return tbsrn::Slots(srb_nums, tbsrn::ARCH_TSRN).count; }

static const char* slot_name_impl(int srb_nums, int arch, int idx) {
  static thread_local std::vector<std::string> names;
  static thread_local int cached = -1;
  const int key = srb_nums * 2 + arch;
  if (cached != key) {
    names = tbsrn::slot_names(srb_nums, arch);
    cached = key;
  }
  if (idx < 0 || idx >= (int)names.size()) return "";
  return names[idx].c_str();
}
const char* focr_tbsrn_slot_name(int srb_nums, int idx) { return slot_name_impl(srb_nums, tbsrn::ARCH_TBSRN, idx); }
const char* focr_tsrn_slot_name(int srb_nums, int idx) { return slot_name_impl(srb_nums, tbsrn::ARCH_TSRN, idx); }

static size_t ws_bytes_impl(int B, int srb_nums, int arch) {
  tbsrn::Ws w;
  tbsrn::layout(w, B, srb_nums, nullptr, arch);
  return w.total_bytes + 256;
}
size_t focr_tbsrn_workspace_bytes(int B, int srb_nums) { return ws_bytes_impl(B, srb_nums, tbsrn::ARCH_TBSRN); }
size_t focr_tsrn_workspace_bytes(int B, int srb_nums) { return ws_bytes_impl(B, srb_nums, tbsrn::ARCH_TSRN); }

static void* align256(void* p) { return (void*)(((uintptr_t)p + 255) & ~(uintptr_t)255); }

static int forward_impl(int arch, void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags,
                        float p_drop, unsigned seed, void* ws, size_t ws_bytes, void* stream,
                        const unsigned* seed_dev = nullptr) {
  FOCR_REQUIRE(B >= 1 && srb_nums >= 0 && srb_nums <= 16, "forward: B=%d srb_nums=%d", B, srb_nums);
  FOCR_REQUIRE(!((flags & 1) && (flags & 2)) || B >= 2, "forward: train mode with STN needs B >= 2 (BatchNorm1d)");
  tbsrn::Ws w;
  tbsrn::layout(w, B, srb_nums, align256(ws), arch);
  FOCR_REQUIRE(ws_bytes >= w.total_bytes + 256, "forward: workspace too small (%zu < %zu)", ws_bytes,
               w.total_bytes + 256);
  tbsrn::Slots sl(srb_nums, arch);
  return tbsrn::forward(sl, params, x_lr, sr, w, flags & 1, (flags >> 1) & 1, p_drop, seed, (cudaStream_t)stream, seed_dev);
}
static int backward_impl(int arch, void* const* params, void* const* grads, const float* x_lr, const float* d_sr, int B,
                         int srb_nums, int flags, float p_drop, unsigned seed, void* ws, size_t ws_bytes, void* stream) {
  FOCR_REQUIRE(flags & 1, "backward: forward must have run in training mode");
  tbsrn::Ws w;
  tbsrn::layout(w, B, srb_nums, align256(ws), arch);
  FOCR_REQUIRE(ws_bytes >= w.total_bytes + 256, "backward: workspace too small");
  tbsrn::Slots sl(srb_nums, arch);
  return tbsrn::backward(sl, params, grads, x_lr, d_sr, w, (flags >> 1) & 1, p_drop, seed, (cudaStream_t)stream);
}

int focr_tbsrn_forward(void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags, float p_drop,
                       unsigned seed, void* ws, size_t ws_bytes, void* stream) {
  return forward_impl(tbsrn::ARCH_TBSRN, params, x_lr, sr, B, srb_nums, flags, p_drop, seed, ws, ws_bytes, stream);
}
// same as focr_tbsrn_forward with the dropout seed read ON THE DEVICE from *seed_dev when the kernels run: a CUDA graph
// captured around the step can be replayed with a fresh seed per step (masks are identical to passing that seed by value)
int focr_tbsrn_forward_devseed(void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags,
                               float p_drop, const unsigned* seed_dev, void* ws, size_t ws_bytes, void* stream) {
  FOCR_REQUIRE(seed_dev != nullptr, "tbsrn_forward_devseed: seed_dev is NULL");
  return forward_impl(tbsrn::ARCH_TBSRN, params, x_lr, sr, B, srb_nums, flags, p_drop, 0, ws, ws_bytes, stream, seed_dev);
}
int focr_tbsrn_backward(void* const* params, void* const* grads, const float* x_lr, const float* d_sr, int B,
                        int srb_nums, int flags, float p_drop, unsigned seed, void* ws, size_t ws_bytes, void* stream) {
  return backward_impl(tbsrn::ARCH_TBSRN, params, grads, x_lr, d_sr, B, srb_nums, flags, p_drop, seed, ws, ws_bytes,
                       stream);
}
// TSRN (model/tsrn.py:18-74): same trunk, SRBs with vertical + horizontal BiGRU instead of the FeatureEnhancer
int focr_tsrn_forward(void* const* params, const float* x_lr, float* sr, int B, int srb_nums, int flags, void* ws,
                      size_t ws_bytes, void* stream) {
  return forward_impl(tbsrn::ARCH_TSRN, params, x_lr, sr, B, srb_nums, flags, 0.f, 0, ws, ws_bytes, stream);
}
int focr_tsrn_backward(void* const* params, void* const* grads, const float* x_lr, const float* d_sr, int B,
                       int srb_nums, int flags, void* ws, size_t ws_bytes, void* stream) {
  return backward_impl(tbsrn::ARCH_TSRN, params, grads, x_lr, d_sr, B, srb_nums, flags, 0.f, 0, ws, ws_bytes, stream);
}

// Byte offset / element count of a named intermediate inside the workspace (parity debugging):
// "x_tps","ctrl","b1","s7","u","opre","srb<i>.<c1|a1|c2|f|qkv|o|y1|y2|out>"
static int ws_tensor_impl(int arch, int B, int srb_nums, const char* name, long long* byte_offset, long long* elems,
                          int* elem_bytes) {
  tbsrn::Ws w;
  char* base = (char*)4096;
  tbsrn::layout(w, B, srb_nums, base, arch);
  const long T = w.T, Thr = w.Thr;
  const void* p = nullptr;
  long n = 0;
  int eb = 2;
  std::string nm(name);
  if (nm == "x_tps") { p = w.x_tps; n = (long)B * 3072; eb = 4; }
  else if (nm == "ctrl") { p = w.ctrl; n = (long)B * 64; eb = 4; }
  else if (nm.rfind("stn.yact", 0) == 0 || nm.rfind("stn.pool", 0) == 0 || nm.rfind("stn.ypre", 0) == 0) {
    const int i = atoi(nm.substr(8).c_str());
    FOCR_REQUIRE(i >= 0 && i < 6, "ws_tensor: bad stn index in %s", name);
    const tbsrn::StnConv& c = tbsrn::kStn[i];
    const long M = (long)B * c.h * c.w;
    if (nm[4] == 'y' && nm[5] == 'a') { p = w.stn_yact[i]; n = M * c.cout; }
    else if (nm[4] == 'y') { p = w.stn_ypre[i]; n = M * c.npad; }
    else { p = w.stn_pool[i]; n = (c.pool_h ? M / (2 * c.pool_h) : M) * c.cout; }
  }
  else if (nm == "stn.f1") { p = w.f1; n = (long)B * 512; }
  else if (nm == "stn.f1pre") { p = w.f1pre; n = (long)B * 512; }
  else if (nm == "b1") { p = w.b1; n = T * 64; }
  else if (nm == "s7") { p = w.s7; n = T * 64; }
  else if (nm == "u") { p = w.u; n = Thr * 64; }
  else if (nm == "opre") { p = w.opre; n = (long)B * 3 * 4096; eb = 4; }
  else if (nm.rfind("srb", 0) == 0) {
    const size_t dot = nm.find('.');
    FOCR_REQUIRE(dot != std::string::npos, "ws_tensor: bad name %s", name);
    const int i = atoi(nm.substr(3, dot - 3).c_str());
    FOCR_REQUIRE(i >= 0 && i < srb_nums, "ws_tensor: bad srb index in %s", name);
    const std::string f = nm.substr(dot + 1);
    const tbsrn::SrbWs& a = w.srb[i];
    if (f == "c1") { p = a.c1; n = T * 64; }
    else if (f == "a1") { p = a.a1; n = T * 64; }
    else if (f == "c2") { p = a.c2; n = T * 64; }
    else if (f == "f") { p = a.f; n = T * 128; }
    else if (f == "qkv") { p = a.qkv; n = T * 384; }
    else if (f == "o") { p = a.o; n = T * 128; }
    else if (f == "y1") { p = a.y1; n = T * 128; }
    else if (f == "y2") { p = a.y2; n = T * 128; }
    else if (f == "out") { p = a.out; n = T * 64; }
    else if (f == "r0" && arch == tbsrn::ARCH_TSRN) { p = a.r0; n = T * 64; }
    else if (f == "o1" && arch == tbsrn::ARCH_TSRN) { p = a.o1; n = T * 64; }
    if (arch == tbsrn::ARCH_TSRN && (f == "f" || f == "qkv" || f == "o" || f == "y1" || f == "y2")) p = nullptr;
  }
  FOCR_REQUIRE(p != nullptr, "ws_tensor: unknown tensor %s", name);
  *byte_offset = (const char*)p - base;
  *elems = n;
  *elem_bytes = eb;
  return FOCR_OK;
}
int focr_tbsrn_ws_tensor(int B, int srb_nums, const char* name, long long* byte_offset, long long* elems,
                         int* elem_bytes) {
  return ws_tensor_impl(tbsrn::ARCH_TBSRN, B, srb_nums, name, byte_offset, elems, elem_bytes);
}
int focr_tsrn_ws_tensor(int B, int srb_nums, const char* name, long long* byte_offset, long long* elems,
                        int* elem_bytes) {
  return ws_tensor_impl(tbsrn::ARCH_TSRN, B, srb_nums, name, byte_offset, elems, elem_bytes);
}

}  // extern "C"

extern "C" int focr_mse_loss_grad(const float* sr, const float* hr, float* d_sr, float* loss, long n, float gscale,
                                  void* ws, size_t ws_bytes, void* stream);
