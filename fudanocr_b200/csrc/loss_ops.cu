// Stand-alone forms of two helpers the reference exports as callables next to its focus losses, so that code importing them
// keeps working on CUDA tensors (the fused losses have their own copies inside focus.cu):
//   weight_cross_entropy(pred, gt)   scene-text-telescope/loss/weight_ce_loss.py:36-45
//   to_gray_tensor(tensor)           scene-text-telescope/loss/text_focus_loss.py:16-21, text-gestalt/loss/stroke_focus_loss.py:12-18
#include "kernels.cuh"

namespace {

// loss = -(1/N) sum_i log( w[g_i][g_i] e^{p_i,g_i} / sum_j w[g_i][j] e^{p_ij} ), evaluated with log-sum-exp (the reference
// exponentiates the raw logits; equal within fp32 rounding wherever the reference does not overflow).
// One warp per row; d_pred (optional) = (softmax_w - onehot) / N; row losses go to `partial` for a deterministic sum.
__global__ void __launch_bounds__(128) wce_rows_kernel(const float* __restrict__ pred, const long long* __restrict__ gt,
                                                       const float* __restrict__ table, float* __restrict__ d_pred,
                                                       float* __restrict__ partial, int* __restrict__ status, int N, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= N) return;
  const long long g64 = gt[row];
  if (g64 < 0 || g64 >= C) {  // torch indexing would raise: report, contribute nothing
    if (lane == 0) {
      atomicExch(status, 1);
      partial[row] = 0.f;
    }
    if (d_pred != nullptr)
      for (int c = lane; c < C; c += 32) d_pred[(long)row * C + c] = 0.f;
    return;
  }
  const int g = (int)g64;
  const float* pr = pred + (long)row * C;
  const float* wr = table + (long)g * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, pr[c] + __logf(wr[c]));
  mx = warp_max(mx);
  float z = 0.f;
  for (int c = lane; c < C; c += 32) z += __expf(pr[c] + __logf(wr[c]) - mx);
  z = warp_sum(z);
  const float zg = pr[g] + __logf(wr[g]);
  if (lane == 0) partial[row] = (mx + __logf(z)) - zg;
  if (d_pred != nullptr) {
    const float inv = 1.f / z, sc = 1.f / (float)N;
    for (int c = lane; c < C; c += 32)
      d_pred[(long)row * C + c] = sc * (__expf(pr[c] + __logf(wr[c]) - mx) * inv - (c == g ? 1.f : 0.f));
  }
}
__global__ void __launch_bounds__(256) wce_sum_kernel(const float* __restrict__ partial, int N, float* __restrict__ loss) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += 256) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    *loss = t / (float)N;
  }
}

__global__ void to_gray_kernel(const float* __restrict__ img, float* __restrict__ gray, int C, long HW, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long b = i / HW, p = i - b * HW;
  const float* s = img + b * C * HW + p;
  gray[i] = 0.299f * s[0] + 0.587f * s[HW] + 0.114f * s[2 * HW];
}
__global__ void to_gray_bwd_kernel(const float* __restrict__ d_gray, float* __restrict__ d_img, int C, long HW, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long b = i / HW, p = i - b * HW;
  float* d = d_img + b * C * HW + p;
  const float g = d_gray[i];
  d[0] = 0.299f * g;
  d[HW] = 0.587f * g;
  d[2 * HW] = 0.114f * g;
  for (int c = 3; c < C; ++c) d[c * HW] = 0.f;
}

}  // namespace

extern "C" size_t focr_weight_cross_entropy_workspace_bytes(long N) { return (size_t)(N + 4) * sizeof(float); }

// pred fp32 (N, C) logits, gt int64 (N), table fp32 (C, C) = load_confuse_matrix(); loss fp32[1]; d_pred fp32 (N, C) or NULL.
// status int32[1] (optional): set to 1 when a gt index is outside [0, C).
extern "C" int focr_weight_cross_entropy(const float* pred, const long long* gt, const float* table, float* loss,
                                         float* d_pred, int* status, long N, int C, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(N >= 1 && C >= 1 && C <= 4096, "weight_cross_entropy: N=%ld C=%d", N, C);
  FOCR_REQUIRE(ws != nullptr && ws_bytes >= focr_weight_cross_entropy_workspace_bytes(N), "weight_cross_entropy: workspace");
  float* partial = (float*)ws;
  int* st = status != nullptr ? status : (int*)(partial + N);
  ProfScope _ps("focus_wce", s);
  wce_rows_kernel<<<focr_cdiv(N, 4), 128, 0, s>>>(pred, gt, table, d_pred, partial, st, (int)N, C);
  FOCR_LAUNCH_CHECK();
  wce_sum_kernel<<<1, 256, 0, s>>>(partial, (int)N, loss);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// img fp32 (B, C >= 3, H, W) NCHW -> gray fp32 (B, 1, H, W) = 0.299 R + 0.587 G + 0.114 B
extern "C" int focr_to_gray(const float* img, float* gray, long B, int C, long HW, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(B >= 1 && C >= 3 && HW >= 1, "to_gray: B=%ld C=%d HW=%ld", B, C, HW);
  const long n = B * HW;
  to_gray_kernel<<<focr_cdiv(n, 256), 256, 0, s>>>(img, gray, C, HW, n);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
extern "C" int focr_to_gray_bwd(const float* d_gray, float* d_img, long B, int C, long HW, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(B >= 1 && C >= 3 && HW >= 1, "to_gray_bwd: B=%ld C=%d HW=%ld", B, C, HW);
  const long n = B * HW;
  to_gray_bwd_kernel<<<focr_cdiv(n, 256), 256, 0, s>>>(d_gray, d_img, C, HW, n);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
