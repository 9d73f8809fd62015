// Weight / input layout preparation kernels (tiny, run once per step because the fp32 master
// weights change every optimizer step).
#include "kernels.cuh"

namespace {

// conv weight fp32 [Co][Ci][k][k]  ->  bf16 forward layout  [tap][Co'][Ci]
// shuf != 0: output channels re-ordered for the PixelShuffle(2) epilogue, co = c*4+sub -> Co' = sub*(Co/4)+c
__global__ void prep_conv_w_fwd_kernel(const float* __restrict__ w, bf16* __restrict__ o, int Co, int Ci,
                                       int ks, int shuf) {
  const int taps = ks * ks;
  const long n = (long)taps * Co * Ci;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int cop = (int)((i / Ci) % Co);
    const int tap = (int)(i / ((long)Ci * Co));
    int co = cop;
    if (shuf) {
      const int q = Co / 4;
      co = (cop % q) * 4 + cop / q;
    }
    o[i] = __float2bfloat16_rn(w[((long)co * Ci + ci) * taps + tap]);
  }
}

// conv weight fp32 [Co][Ci][k][k]  ->  bf16 dgrad layout  [tap'][Ci][Co'] with tap' the flipped tap
// (dX[p] = sum_tap' dY[p + tap'] * W[.., flipped]).
__global__ void prep_conv_w_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ o, int Co, int Ci,
                                         int ks, int shuf) {
  const int taps = ks * ks;
  const long n = (long)taps * Co * Ci;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cop = (int)(i % Co);
    const int ci = (int)((i / Co) % Ci);
    const int tapf = (int)(i / ((long)Ci * Co));
    const int tap = taps - 1 - tapf;
    int co = cop;
    if (shuf) {
      const int q = Co / 4;
      co = (cop % q) * 4 + cop / q;
    }
    o[i] = __float2bfloat16_rn(w[((long)co * Ci + ci) * taps + tap]);
  }
}

__global__ void prep_bias_shuf_kernel(const float* __restrict__ b, float* __restrict__ o, int Co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co) {
    const int q = Co / 4;
    o[i] = b[(i % q) * 4 + i / q];
  }
}

// inverse of prep_bias_shuf: o[c*4 + sub] = b[sub*(Co/4) + c]
__global__ void bias_unshuf_kernel(const float* __restrict__ b, float* __restrict__ o, int Co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Co) {
    const int q = Co / 4;
    o[(i % q) * 4 + i / q] = b[i];
  }
}

// fp32 [R][C] -> bf16 [R][C] and (optionally) bf16 transposed [C][R]
__global__ void prep_linear_w_kernel(const float* __restrict__ w, bf16* __restrict__ o, bf16* __restrict__ ot,
                                     int R, int C, int ot_ld, int ot_off) {
  const long n = (long)R * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), r = (int)(i / C);
    const bf16 v = __float2bfloat16_rn(w[i]);
    if (o) o[i] = v;
    if (ot) ot[(long)c * ot_ld + ot_off + r] = v;
  }
}

}  // namespace

int prep_conv_w_fwd(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s) {
  const long n = (long)ks * ks * Co * Ci;
  prep_conv_w_fwd_kernel<<<focr_cdiv(n, 256), 256, 0, s>>>(w, o, Co, Ci, ks, shuf);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int prep_conv_w_dgrad(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s) {
  const long n = (long)ks * ks * Co * Ci;
  prep_conv_w_dgrad_kernel<<<focr_cdiv(n, 256), 256, 0, s>>>(w, o, Co, Ci, ks, shuf);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int prep_bias_shuf(const float* b, float* o, int Co, cudaStream_t s) {
  prep_bias_shuf_kernel<<<focr_cdiv(Co, 128), 128, 0, s>>>(b, o, Co);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int bias_unshuf(const float* b, float* o, int Co, cudaStream_t s) {
  bias_unshuf_kernel<<<focr_cdiv(Co, 128), 128, 0, s>>>(b, o, Co);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
// ot (optional): transposed copy, element (c, r) written at ot[c*ot_ld + ot_off + r]
int prep_linear_w(const float* w, bf16* o, bf16* ot, int R, int C, int ot_ld, int ot_off, cudaStream_t s) {
  const long n = (long)R * C;
  prep_linear_w_kernel<<<focr_cdiv(n, 256), 256, 0, s>>>(w, o, ot, R, C, ot_ld, ot_off);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
