// Evaluation metrics of the SR pipeline on the device: PSNR and SSIM of (sr, hr) batches.
// Reference: scene-text-telescope/utils/ssim_psnr.py:9-15 (calculate_psnr), :18-28 (gaussian window 11, sigma 1.5),
// :31-51 (_ssim: five depth-wise 11x11 "same" convolutions + the SSIM map + mean), :54-78 (SSIM module, first 3 channels);
// called per validation batch by interfaces/super_resolution.py:191-192.  The reference launches ~25 kernels per batch and
// materialises ten (B,3,32,128) temporaries; here one CTA per (image, channel) keeps both zero-padded planes in shared
// memory, evaluates the five windowed moments per pixel in registers and reduces SSIM and the squared error on the fly.
#include "kernels.cuh"

namespace {

constexpr int kH = 32, kW = 128, kR = 5, kWin = 11;
constexpr int kPH = kH + 2 * kR, kPW = kW + 2 * kR;  // 42 x 138

__global__ void __launch_bounds__(256) ssim_psnr_kernel(const float* __restrict__ a, const float* __restrict__ b, int c_total,
                                                        const float* __restrict__ window, float* __restrict__ partial) {
  __shared__ float sa[kPH * kPW];
  __shared__ float sb[kPH * kPW];
  __shared__ float sw[kWin * kWin];
  __shared__ float red[2][8];
  const int img = blockIdx.x / 3, ch = blockIdx.x % 3;
  const float* pa = a + ((long)img * c_total + ch) * kH * kW;
  const float* pb = b + ((long)img * c_total + ch) * kH * kW;
  for (int i = threadIdx.x; i < kPH * kPW; i += 256) {
    const int y = i / kPW - kR, x = i % kPW - kR;
    const bool in = y >= 0 && y < kH && x >= 0 && x < kW;
    sa[i] = in ? pa[y * kW + x] : 0.f;
    sb[i] = in ? pb[y * kW + x] : 0.f;
  }
  if (threadIdx.x < kWin * kWin) sw[threadIdx.x] = window[threadIdx.x];
  __syncthreads();
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  float ssim_sum = 0.f, se_sum = 0.f;
  for (int p = threadIdx.x; p < kH * kW; p += 256) {
    const int y = p / kW, x = p % kW;
    float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll 1
    for (int ky = 0; ky < kWin; ++ky) {
      const float* ra = sa + (y + ky) * kPW + x;
      const float* rb = sb + (y + ky) * kPW + x;
#pragma unroll
      for (int kx = 0; kx < kWin; ++kx) {
        const float w = sw[ky * kWin + kx], u = ra[kx], v = rb[kx];
        mu1 = fmaf(w, u, mu1);
        mu2 = fmaf(w, v, mu2);
        s11 = fmaf(w, u * u, s11);
        s22 = fmaf(w, v * v, s22);
        s12 = fmaf(w, u * v, s12);
      }
    }
    const float m11 = mu1 * mu1, m22 = mu2 * mu2, m12 = mu1 * mu2;
    const float v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
    ssim_sum += ((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
    // img1*255 - img2*255 with each product rounded as in the reference (no FMA contraction: identical images -> exactly 0)
    const float d = __fsub_rn(__fmul_rn(sa[(y + kR) * kPW + x + kR], 255.f), __fmul_rn(sb[(y + kR) * kPW + x + kR], 255.f));
    se_sum = fmaf(d, d, se_sum);
  }
  ssim_sum = warp_sum(ssim_sum);
  se_sum = warp_sum(se_sum);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = ssim_sum;
    red[1][threadIdx.x >> 5] = se_sum;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    partial[(long)blockIdx.x * 2 + threadIdx.x] = s;
  }
}

// out[0] = psnr = 20 log10(255 / sqrt(mse)) (inf when mse == 0), out[1] = mean SSIM, per_image[i] = mean SSIM of image i
__global__ void ssim_psnr_finish_kernel(const float* __restrict__ partial, int B, float* __restrict__ out,
                                        float* __restrict__ per_image) {
  double ss = 0.0, se = 0.0;
  for (int i = threadIdx.x; i < B * 3; i += 32) {
    ss += partial[2 * i];
    se += partial[2 * i + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
  }
  const double n = (double)B * 3 * kH * kW;
  if (threadIdx.x == 0) {
    const double mse = se / n;
    out[0] = mse == 0.0 ? INFINITY : (float)(20.0 * log10(255.0 / sqrt(mse)));
    out[1] = (float)(ss / n);
  }
  if (per_image != nullptr)
    for (int i = threadIdx.x; i < B; i += 32)
      per_image[i] = (partial[6 * i] + partial[6 * i + 2] + partial[6 * i + 4]) / (3.f * kH * kW);
}

}  // namespace

extern "C" size_t focr_psnr_ssim_workspace_bytes(int B) { return (size_t)B * 3 * 2 * sizeof(float); }

extern "C" int focr_psnr_ssim(const float* img1, const float* img2, int B, int channels, const float* window, float* out,
                              float* ssim_per_image, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(B >= 1 && channels >= 3, "psnr_ssim: B=%d channels=%d (the metrics use the first 3 channels)", B, channels);
  FOCR_REQUIRE(ws && ws_bytes >= focr_psnr_ssim_workspace_bytes(B), "psnr_ssim: workspace too small");
  ProfScope _ps("metrics", s);
  ssim_psnr_kernel<<<B * 3, 256, 0, s>>>(img1, img2, channels, window, (float*)ws);
  FOCR_LAUNCH_CHECK();
  ssim_psnr_finish_kernel<<<1, 32, 0, s>>>((const float*)ws, B, out, ssim_per_image);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
