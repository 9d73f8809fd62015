// The two 9x9 convolutions of TBSRN (tbsrn.py:180 block1 3->64 on the LR image, :196 64->3 on the HR
// feature map) have a 3-channel side, which makes a direct implicit GEMM hopeless (N = 3 or K = 3 per
// tap).  Both are evaluated "dx-unrolled": the 9 horizontal taps of the 3-channel tensor are folded into
// the channel axis (27 of 64 channels), and the 9 vertical taps run through the tcgen05 implicit-GEMM
// engine as a 9x1 (or 1x9) convolution between two 64-channel NHWC maps:
//   block1 fwd  : A1 = im2col_dx(x)           ; Y  = conv9x1(A1; Wv) + bias, PReLU       (tc_gemm)
//   block1 wgrad: T  = conv9x1_wgrad(dY, A1)  ; dW = repack(T)
//   block1 dgrad: Z  = conv1x9(dY; Wh')       ; dX = vertical_gather(Z)                    (STN path only)
//   final fwd   : Z  = conv1x9(U; Wh)         ; O  = vertical_gather(Z) + bias
//   final dgrad : A1 = im2col_dx(dO, flipped) ; dU = conv9x1(A1; Wd)
//   final wgrad : T  = conv9x1_wgrad(A1, U)   ; dW = repack(T)
// This file holds the layout kernels around those GEMMs.
#include "kernels.cuh"

namespace {

// in: fp32 NCHW (B,3,H,W);  out: bf16 NHWC (B,H,W,64), out[.., dx*3+c] = in[b,c,h, w + sgn*(dx-4)] (0 outside),
// channels 27..63 zero.
__global__ void im2col_dx_kernel(const float* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int sgn) {
  const long n = (long)B * H * W * 8;  // one thread = 8 output channels (16 bytes)
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i & 7);
    const long pix = i >> 3;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int b = (int)(pix / ((long)W * H));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      float val = 0.f;
      if (k < 27) {
        const int dx = k / 3, c = k - dx * 3;
        const int ws = w + sgn * (dx - 4);
        if (ws >= 0 && ws < W) val = in[(((long)b * 3 + c) * H + h) * W + ws];
      }
      v[j] = val;
    }
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]);
    u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + pix * 64 + ch * 8) = u;
  }
}

// z: fp32 NHWC (B,H,W,64) with column n = dy*3+c;  out fp32 NCHW (B,3,H,W):
//   out[b,c,y,x] = bias[c] + sum_dy z[b, y + sgn*(dy-4), x][dy*3+c]
__global__ void vgather_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ out,
                               int B, int H, int W, int sgn) {
  const long n = (long)B * H * W;
  for (long pix = blockIdx.x * (long)blockDim.x + threadIdx.x; pix < n; pix += (long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((long)W * H));
    float a0 = bias ? bias[0] : 0.f, a1 = bias ? bias[1] : 0.f, a2 = bias ? bias[2] : 0.f;
#pragma unroll
    for (int dy = 0; dy < 9; ++dy) {
      const int ys = y + sgn * (dy - 4);
      if (ys >= 0 && ys < H) {
        const float* zp = z + (((long)b * H + ys) * W + x) * 64 + dy * 3;
        a0 += zp[0];
        a1 += zp[1];
        a2 += zp[2];
      }
    }
    const long o = ((long)b * 3 * H + y) * W + x;
    out[o] = a0;
    out[o + (long)H * W] = a1;
    out[o + 2L * H * W] = a2;
  }
}

// Weight layouts for the four GEMM forms.  w is the torch tensor [Co][Ci][9][9] (fp32).
//  mode 0 (block1 fwd,  Co=64,Ci=3): o[tap=dy][n=co][k=dx*3+c]      = w[co][c][dy][dx]
//  mode 1 (final dgrad, Co=3,Ci=64): o[tap][n=c][k=dx*3+co]          = w[co][c][8-tap][dx]
//  mode 2 (final fwd,   Co=3,Ci=64): o[tap=dx][n=dy*3+co][k=c]       = w[co][c][dy][dx]
//  mode 3 (block1 dgrad,Co=64,Ci=3): o[tap][n=dy*3+c][k=co]          = w[co][c][dy][8-tap]
// o is bf16 [9][64][64], unused rows/columns zero.
__global__ void prep_w9_kernel(const float* __restrict__ w, bf16* __restrict__ o, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * 64 * 64) return;
  const int k = i & 63, n = (i >> 6) & 63, tap = i >> 12;
  float v = 0.f;
  if (mode == 0) {
    if (k < 27) {
      const int dx = k / 3, c = k % 3;
      v = w[((n * 3 + c) * 9 + tap) * 9 + dx];
    }
  } else if (mode == 1) {
    if (k < 27) {
      const int dx = k / 3, co = k % 3;
      v = w[((co * 64 + n) * 9 + (8 - tap)) * 9 + dx];
    }
  } else if (mode == 2) {
    if (n < 27) {
      const int dy = n / 3, co = n % 3;
      v = w[((co * 64 + k) * 9 + dy) * 9 + tap];
    }
  } else {
    if (n < 27) {
      const int dy = n / 3, c = n % 3;
      v = w[((k * 3 + c) * 9 + dy) * 9 + (8 - tap)];
    }
  }
  o[i] = __float2bfloat16_rn(v);
}

// t: fp32 [64][64][9] from conv9x1_wgrad.
//  mode 0 (block1): t[co][k=dx*3+c][dy] -> dw[co][c][dy][dx]   (dw [64][3][9][9])
//  mode 1 (final) : t[k=dx*3+co][c][dy] -> dw[co][c][dy][dx]   (dw [3][64][9][9])
__global__ void repack_w9_kernel(const float* __restrict__ t, float* __restrict__ dw, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 3 * 81) return;
  const int dx = i % 9, dy = (i / 9) % 9;
  if (mode == 0) {
    const int c = (i / 81) % 3, co = i / 243;
    dw[i] = t[((long)co * 64 + dx * 3 + c) * 9 + dy];
  } else {
    const int c = (i / 81) % 64, co = i / (81 * 64);
    dw[i] = t[((long)(dx * 3 + co) * 64 + c) * 9 + dy];
  }
}

// per-channel sum of an fp32 NCHW (B,3,H,W) tensor -> partial[blk][3]
__global__ void __launch_bounds__(256) nchw3_sum_kernel(const float* __restrict__ x, int B, long hw,
                                                        float* __restrict__ partial) {
  __shared__ float red[3][8];
  float a[3] = {0.f, 0.f, 0.f};
  const long n = (long)B * 3 * hw;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const int c = (int)((i / hw) % 3);
    const float v = x[i];
    a[0] += c == 0 ? v : 0.f;
    a[1] += c == 1 ? v : 0.f;
    a[2] += c == 2 ? v : 0.f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = warp_sum(a[c]);
    if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    partial[blockIdx.x * 3 + threadIdx.x] = s;
  }
}

int grid_for(long n, int per_block) {
  long g = (n + per_block - 1) / per_block;
  if (g > 148L * 8) g = 148L * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int im2col_dx(const float* in, bf16* out, int B, int H, int W, int sgn, cudaStream_t s) {
  im2col_dx_kernel<<<grid_for((long)B * H * W * 8, 256), 256, 0, s>>>(in, out, B, H, W, sgn);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int vgather9(const float* z, const float* bias, float* out, int B, int H, int W, int sgn, cudaStream_t s) {
  vgather_kernel<<<grid_for((long)B * H * W, 256), 256, 0, s>>>(z, bias, out, B, H, W, sgn);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int prep_w9(const float* w, bf16* o, int mode, cudaStream_t s) {
  prep_w9_kernel<<<focr_cdiv(9 * 64 * 64, 256), 256, 0, s>>>(w, o, mode);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int repack_w9(const float* t, float* dw, int mode, cudaStream_t s) {
  repack_w9_kernel<<<focr_cdiv(64 * 3 * 81, 256), 256, 0, s>>>(t, dw, mode);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int nchw3_sum_blocks(int B, long hw) { return grid_for((long)B * 3 * hw, 256 * 8); }
int nchw3_sum(const float* x, int B, long hw, float* out3, float* partial, cudaStream_t s) {
  const int P = nchw3_sum_blocks(B, hw);
  nchw3_sum_kernel<<<P, 256, 0, s>>>(x, B, hw, partial);
  FOCR_LAUNCH_CHECK();
  return reduce_partials(partial, P, 3, 3, out3, 1.f, s);
}
