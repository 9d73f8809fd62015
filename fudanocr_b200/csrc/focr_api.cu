// C-ABI entry points of libfocr_sm100.so (declared in include/focr.h).
// Plain pointers and sizes only; every call enqueues work on the caller's stream and returns
// 0 or a negative error code (message via focr_last_error()).  No CPU fallback exists: on a
// machine without an sm_100 device every compute entry point fails with FOCR_ERR_CUDA.
#include <stdarg.h>
#include <string.h>

#include "kernels.cuh"

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
  p.ksize = ksize;
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
  p.ksize = ksize;
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
  int rc = prep_linear_w(w, wb, nullptr, N, K, s);
  if (rc) return rc;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = N;
  p.ksize = 1;
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
  int rc = prep_linear_w(w, nullptr, wt, N, K, s);  // wt: [K][N]
  if (rc) return rc;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = K;
  p.ksize = 1;
  p.W = 64;
  p.H = 2;
  p.epi = TC_EPI_BF16;
  p.ldc = K;
  p.out = dx;
  const bf16* ap[1] = {(const bf16*)dy};
  return tc_gemm_launch(ap, 1, N, (long)64 * N, (long)128 * N, N, (int)(M / 128), wt, N, p, s);
}

}  // extern "C"
