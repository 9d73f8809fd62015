// Contrastive head of the CCR-CLIP pre-training stage (image-ids-CTR/CCR-CLIP/model.py:209-222, main.py:98-110):
//
//   i_n = image_features / |image_features|,  t_n = text_features / |text_features|,  s = exp(logit_scale)
//   logits_per_image = s * i_n t_n^T ;  logits_per_text = logits_per_image^T
//   loss = ( CE(logits_per_image, gt) + CE(logits_per_text, gt) ) / 2         gt[i] = first sample carrying sample i's label
//
// and its gradients w.r.t. the UN-normalised features of both towers and logit_scale, in one call.  The towers (ResNet-50 image
// encoder, 12-layer text transformer) are not part of this repository; the head is what a data-parallel run adds to them: every
// rank all-gathers the (B, D) features (B <= 1024, D = 2048: 8 MB) and evaluates the B x B problem redundantly, so the gradient
// rows of its own shard need no second exchange.  Everything is fp32 SIMT - 2 B^2 D FLOP is 0.07 GFLOP at the reference's B = 128 -
// with fixed-order reductions (bit-reproducible).
#include "kernels.cuh"

namespace {

constexpr int kClipThreads = 256;

// n = x / |x| per row, inv[r] = 1 / |x_r|; blockIdx.y selects the tower (0 image, 1 text)
__global__ void __launch_bounds__(kClipThreads) clip_norm_kernel(const float* __restrict__ img, const float* __restrict__ txt,
                                                                 float* __restrict__ nimg, float* __restrict__ ntxt,
                                                                 float* __restrict__ inv, int B, int D) {
  __shared__ float red[kClipThreads / 32];
  const int r = blockIdx.x;
  const float* x = (blockIdx.y ? txt : img) + (long)r * D;
  float* n = (blockIdx.y ? ntxt : nimg) + (long)r * D;
  float s = 0.f;
  for (int c = threadIdx.x; c < D; c += kClipThreads) s += x[c] * x[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kClipThreads / 32; ++w) tot += red[w];
  const float iv = rsqrtf(tot);
  for (int c = threadIdx.x; c < D; c += kClipThreads) n[c] = x[c] * iv;
  if (threadIdx.x == 0) inv[blockIdx.y * B + r] = iv;
}

// S[i][j] = i_n[i] . t_n[j]   (cosine similarities, unscaled); 32 x 32 tile per CTA, K walked in chunks of 32 through smem
__global__ void __launch_bounds__(kClipThreads) clip_sim_kernel(const float* __restrict__ nimg, const float* __restrict__ ntxt,
                                                                float* __restrict__ S, int B, int D) {
  __shared__ float a[32][33], b[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 row groups of 4 rows
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < D; k0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ty * 4 + q;
      a[r][tx] = (i0 + r < B && k0 + tx < D) ? nimg[(long)(i0 + r) * D + k0 + tx] : 0.f;
      b[r][tx] = (j0 + r < B && k0 + tx < D) ? ntxt[(long)(j0 + r) * D + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float bv = b[tx][k];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += a[ty * 4 + q][k] * bv;
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (i0 + ty * 4 + q < B && j0 + tx < B) S[(long)(i0 + ty * 4 + q) * B + j0 + tx] = acc[q];
}

// one warp per row (blockIdx.y = 0) or column (= 1) of s * S: log-sum-exp, and that line's CE term into terms[y * B + line]
__global__ void __launch_bounds__(128) clip_lse_kernel(const float* __restrict__ S, const float* __restrict__ logit_scale,
                                                       const long long* __restrict__ gt, float* __restrict__ lse,
                                                       float* __restrict__ terms, int* __restrict__ status, int B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int line = blockIdx.x * 4 + warp;
  if (line >= B) return;
  const bool col = blockIdx.y != 0;
  const float sc = __expf(logit_scale[0]);
  const long stride = col ? B : 1;
  const float* p = S + (col ? (long)line : (long)line * B);
  float mx = -INFINITY;
  for (int j = lane; j < B; j += 32) mx = fmaxf(mx, sc * p[j * stride]);
  mx = warp_max(mx);
  float z = 0.f;
  for (int j = lane; j < B; j += 32) z += __expf(sc * p[j * stride] - mx);
  z = warp_sum(z);
  if (lane == 0) {
    const float l = mx + __logf(z);
    lse[blockIdx.y * B + line] = l;
    long long g = gt[line];
    if (g < 0 || g >= B) {   // torch's CrossEntropyLoss would raise: report it, score the line against itself
      atomicExch(status, 1);
      g = line;
    }
    terms[blockIdx.y * B + line] = l - sc * p[g * stride];
  }
}

// G = d loss / d (s S): row-softmax and column-softmax minus their one-hots, each / (2B); rowdot[i] = sum_j G_ij S_ij
__global__ void __launch_bounds__(128) clip_dlogits_kernel(const float* __restrict__ S, const float* __restrict__ logit_scale,
                                                           const long long* __restrict__ gt, const float* __restrict__ lse,
                                                           float* __restrict__ G, float* __restrict__ rowdot, int B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= B) return;
  const float sc = __expf(logit_scale[0]);
  const float w = 0.5f / (float)B;
  long long gi = gt[i];
  if (gi < 0 || gi >= B) gi = i;
  const float lr = lse[i];
  float dot = 0.f;
  for (int j = lane; j < B; j += 32) {
    const float s = S[(long)i * B + j];
    long long gj = gt[j];
    if (gj < 0 || gj >= B) gj = j;
    const float g = w * ((__expf(sc * s - lr) - (j == gi ? 1.f : 0.f)) + (__expf(sc * s - lse[B + j]) - (i == gj ? 1.f : 0.f)));
    G[(long)i * B + j] = g;
    dot += g * s;
  }
  dot = warp_sum(dot);
  if (lane == 0) rowdot[i] = dot;
}

// gradient w.r.t. the un-normalised features of one row: dn = s * sum_j G[i][j] t_n[j]  (image rows; blockIdx.y = 1: text rows,
// dn = s * sum_i G[i][r] i_n[i]), then through the normalisation: dx = inv * (dn - n (n . dn))
__global__ void __launch_bounds__(kClipThreads) clip_grad_kernel(const float* __restrict__ G, const float* __restrict__ nimg,
                                                                 const float* __restrict__ ntxt, const float* __restrict__ inv,
                                                                 const float* __restrict__ logit_scale, float* __restrict__ d_img,
                                                                 float* __restrict__ d_txt, int B, int D) {
  extern __shared__ float sm_clip[];
  float* gl = sm_clip;                    // the G row / column of this feature row
  __shared__ float red[kClipThreads / 32];
  const int r = blockIdx.x;
  const bool text = blockIdx.y != 0;
  const float sc = __expf(logit_scale[0]);
  for (int j = threadIdx.x; j < B; j += kClipThreads) gl[j] = text ? G[(long)j * B + r] : G[(long)r * B + j];
  __syncthreads();
  const float* other = text ? nimg : ntxt;
  const float* self = (text ? ntxt : nimg) + (long)r * D;
  float* out = (text ? d_txt : d_img) + (long)r * D;
  float proj = 0.f;
  for (int c = threadIdx.x; c < D; c += kClipThreads) {
    float acc = 0.f;
    int j = 0;
    for (; j + 4 <= B; j += 4) {   // four rows in flight
      const float v0 = other[(long)j * D + c], v1 = other[(long)(j + 1) * D + c], v2 = other[(long)(j + 2) * D + c],
                  v3 = other[(long)(j + 3) * D + c];
      acc += gl[j] * v0;
      acc += gl[j + 1] * v1;
      acc += gl[j + 2] * v2;
      acc += gl[j + 3] * v3;
    }
    for (; j < B; ++j) acc += gl[j] * other[(long)j * D + c];
    acc *= sc;
    out[c] = acc;                 // dn, finished below
    proj += acc * self[c];
  }
  proj = warp_sum(proj);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = proj;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kClipThreads / 32; ++w) tot += red[w];
  const float iv = inv[(text ? B : 0) + r];
  for (int c = threadIdx.x; c < D; c += kClipThreads) out[c] = iv * (out[c] - self[c] * tot);   // each thread re-reads its own columns
}

// loss = sum(terms) / (2B); d logit_scale = s * sum_i rowdot[i]   (one warp, fixed order)
__global__ void clip_finish_kernel(const float* __restrict__ terms, const float* __restrict__ rowdot,
                                   const float* __restrict__ logit_scale, float* __restrict__ loss,
                                   float* __restrict__ d_logit_scale, int B) {
  const int lane = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < 2 * B; i += 32) a += terms[i];
  for (int i = lane; i < B; i += 32) b += rowdot[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    loss[0] = (float)(a / (2.0 * B));
    if (d_logit_scale != nullptr) d_logit_scale[0] = __expf(logit_scale[0]) * (float)b;
  }
}

size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

size_t focr_clip_contrastive_workspace_bytes(int B, int D) {
  return 2 * up256((size_t)B * D * 4) + 2 * up256((size_t)B * B * 4) + 4 * up256((size_t)2 * B * 4) + 256;
}

// image / text: fp32 (B, D) un-normalised tower outputs; logit_scale: the log-parameter (1 float, device); gt: int64 (B).
// loss (1 float); d_image / d_text (B, D) and d_logit_scale (1 float) may be NULL (value only).  status (optional int, device) is
// set to 1 when a target index is outside [0, B).
int focr_clip_contrastive_loss(const float* image, const float* text, const float* logit_scale, const long long* gt, int B, int D,
                               float* loss, float* d_image, float* d_text, float* d_logit_scale, int* status, void* ws,
                               size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(image && text && logit_scale && gt && loss && ws, "clip_contrastive_loss: null pointer");
  FOCR_REQUIRE(B >= 1 && B <= 8192 && D >= 1, "clip_contrastive_loss: B=%d D=%d", B, D);
  FOCR_REQUIRE((d_image == nullptr) == (d_text == nullptr), "clip_contrastive_loss: pass both feature gradients or neither");
  FOCR_REQUIRE(ws_bytes >= focr_clip_contrastive_workspace_bytes(B, D), "clip_contrastive_loss: workspace too small");
  ProfScope _ps("clip_contrastive", s);
  char* p = (char*)ws;
  float* nimg = (float*)p; p += up256((size_t)B * D * 4);
  float* ntxt = (float*)p; p += up256((size_t)B * D * 4);
  float* S = (float*)p;    p += up256((size_t)B * B * 4);
  float* G = (float*)p;    p += up256((size_t)B * B * 4);
  float* inv = (float*)p;  p += up256((size_t)2 * B * 4);
  float* lse = (float*)p;  p += up256((size_t)2 * B * 4);
  float* terms = (float*)p; p += up256((size_t)2 * B * 4);
  float* rowdot = (float*)p; p += up256((size_t)2 * B * 4);
  int* st = status ? status : (int*)p;
  if (!status) FOCR_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(int), s));
  clip_norm_kernel<<<dim3(B, 2), kClipThreads, 0, s>>>(image, text, nimg, ntxt, inv, B, D);
  FOCR_LAUNCH_CHECK();
  clip_sim_kernel<<<dim3((B + 31) / 32, (B + 31) / 32), kClipThreads, 0, s>>>(nimg, ntxt, S, B, D);
  FOCR_LAUNCH_CHECK();
  clip_lse_kernel<<<dim3((B + 3) / 4, 2), 128, 0, s>>>(S, logit_scale, gt, lse, terms, st, B);
  FOCR_LAUNCH_CHECK();
  const bool want = d_image != nullptr;
  if (want || d_logit_scale) {
    clip_dlogits_kernel<<<(B + 3) / 4, 128, 0, s>>>(S, logit_scale, gt, lse, G, rowdot, B);
    FOCR_LAUNCH_CHECK();
  }
  if (want) {
    clip_grad_kernel<<<dim3(B, 2), kClipThreads, (size_t)B * 4, s>>>(G, nimg, ntxt, inv, logit_scale, d_image, d_text, B, D);
    FOCR_LAUNCH_CHECK();
  }
  clip_finish_kernel<<<1, 32, 0, s>>>(terms, rowdot, logit_scale, loss, d_logit_scale, B);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

}  // extern "C"
