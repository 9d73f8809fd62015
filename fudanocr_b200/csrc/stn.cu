// STN head + TPS rectification kernels (train-only prologue of TBSRN, tbsrn.py:215-218):
//   stn_head.py:25-99  six conv3x3+BN+ReLU blocks with max-pools, FC512+BN1d+ReLU, FC -> 20 control points
//   tps_spatial_transformer.py:97-112  two small matmuls -> sampling grid -> bilinear grid_sample
// The convolutions run as explicit im2col + the tcgen05 GEMM engine (spatial sizes 16x64 .. 1x2 are too
// small for the tiled-TMA implicit path); this file has the layout kernels (im2col / col2im / max-pool)
// and the fused TPS kernels.  Everything here is a few MFLOP per image.
#include "kernels.cuh"

namespace {

// col[p][tap*C + c] = x[p + tap][c], p = (b,h,w); K padded to Kpad with zeros.  Source either NHWC bf16
// (C % 8 == 0) or the raw fp32 NCHW image (C == 3).
__global__ void im2col3x3_kernel(const bf16* __restrict__ x, const float* __restrict__ x_nchw, bf16* __restrict__ col,
                                 int B, int H, int W, int C, int Kpad) {
  const long n = (long)B * H * W * (Kpad / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % (Kpad / 8));
    const long pix = i / (Kpad / 8);
    const int w = (int)(pix % W), h = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
    uint4 u = make_uint4(0, 0, 0, 0);
    if (x != nullptr) {
      const int k0 = kc * 8;
      if (k0 < 9 * C) {
        const int tap = k0 / C, c0 = k0 - tap * C;
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
          u = *reinterpret_cast<const uint4*>(x + (((long)b * H + hh) * W + ww) * C + c0);
      }
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kc * 8 + j;
        float val = 0.f;
        if (k < 9 * C) {
          const int tap = k / C, c = k - tap * C;
          const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = x_nchw[(((long)b * C + c) * H + hh) * W + ww];
        }
        v[j] = val;
      }
      u.x = pack_bf16x2(v[0], v[1]);
      u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]);
      u.w = pack_bf16x2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(col + pix * Kpad + kc * 8) = u;
  }
}

// dx[p][c] = sum_tap dcol[p - tap_offset][tap*C + c]   (adjoint of im2col3x3)
__global__ void col2im3x3_kernel(const bf16* __restrict__ dcol, bf16* __restrict__ dx, int B, int H, int W, int C,
                                 int Kpad) {
  const long n = (long)B * H * W * (C / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % (C / 8));
    const long pix = i / (C / 8);
    const int w = (int)(pix % W), h = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h - (tap / 3 - 1), ww = w - (tap % 3 - 1);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const uint4 u = *reinterpret_cast<const uint4*>(dcol + (((long)b * H + hh) * W + ww) * Kpad + tap * C + cc * 8);
        float2 f;
        f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dx + pix * C + cc * 8) = o;
  }
}

// NHWC max-pool, window (ph x 2), stride = window (nn.MaxPool2d(2,2) / ((1,2),(1,2)), stn_head.py:33-43)
__global__ void maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C, int ph) {
  const int Ho = H / ph, Wo = W / 2;
  const long n = (long)B * Ho * Wo * (C / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % (C / 8));
    const long pix = i / (C / 8);
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int dy = 0; dy < ph; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((long)b * H + ho * ph + dy) * W + wo * 2 + dx) * C + cc * 8);
        float2 f;
        f = unpack_bf16x2(u.x); m[0] = fmaxf(m[0], f.x); m[1] = fmaxf(m[1], f.y);
        f = unpack_bf16x2(u.y); m[2] = fmaxf(m[2], f.x); m[3] = fmaxf(m[3], f.y);
        f = unpack_bf16x2(u.z); m[4] = fmaxf(m[4], f.x); m[5] = fmaxf(m[5], f.y);
        f = unpack_bf16x2(u.w); m[6] = fmaxf(m[6], f.x); m[7] = fmaxf(m[7], f.y);
      }
    uint4 o;
    o.x = pack_bf16x2(m[0], m[1]);
    o.y = pack_bf16x2(m[2], m[3]);
    o.z = pack_bf16x2(m[4], m[5]);
    o.w = pack_bf16x2(m[6], m[7]);
    *reinterpret_cast<uint4*>(y + pix * C + cc * 8) = o;
  }
}

// gradient goes to the first window element equal to the max (torch's argmax tie-break: scan order)
__global__ void maxpool_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const bf16* __restrict__ dy,
                                   bf16* __restrict__ dx, int B, int H, int W, int C, int ph) {
  const int Ho = H / ph, Wo = W / 2;
  const long n = (long)B * Ho * Wo * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long pix = i / C;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    const float mv = __bfloat162float(y[i]);
    const bf16 g = dy[i];
    bool done = false;
    for (int dyy = 0; dyy < ph; ++dyy)
      for (int dxx = 0; dxx < 2; ++dxx) {
        const long xi = (((long)b * H + ho * ph + dyy) * W + wo * 2 + dxx) * C + c;
        const bool hit = !done && __bfloat162float(x[xi]) == mv;
        dx[xi] = hit ? g : __float2bfloat16_rn(0.f);
        done = done || hit;
      }
  }
}

// ---------------------------------------------------------------------------------------------
// TPS: per image, mapping = inverse_kernel(23x23) * [ctrl(20x2); 0(3x2)];  src = repr(HWx23) * mapping;
// grid = 2*clamp(src,0,1)-1; bilinear grid_sample (zeros padding, align_corners=False).
// ctrl is read from a (B, ld_ctrl) fp32 matrix (first 40 columns).
// ---------------------------------------------------------------------------------------------
constexpr int kTpsH = 16, kTpsW = 64, kTpsHW = kTpsH * kTpsW;

__device__ __forceinline__ void tps_mapping(const float* __restrict__ inv, const float* __restrict__ ctrl, float* smap,
                                            int tid, int nthreads) {
  // smap[r*2+d] = sum_{j<20} inv[r][j] * ctrl[j][d]
  for (int i = tid; i < 46; i += nthreads) {
    const int r = i >> 1, d = i & 1;
    float a = 0.f;
    for (int j = 0; j < 20; ++j) a += inv[r * 23 + j] * ctrl[j * 2 + d];
    smap[i] = a;
  }
}

__global__ void __launch_bounds__(256) tps_fwd_kernel(const float* __restrict__ img, const float* __restrict__ ctrl,
                                                      long ld_ctrl, const float* __restrict__ inv,
                                                      const float* __restrict__ repr, float* __restrict__ out) {
  __shared__ float smap[46];
  const int b = blockIdx.x;
  tps_mapping(inv, ctrl + (long)b * ld_ctrl, smap, threadIdx.x, 256);
  __syncthreads();
  const float* im = img + (long)b * 3 * kTpsHW;
  for (int p = threadIdx.x; p < kTpsHW; p += 256) {
    float sx = 0.f, sy = 0.f;
    for (int r = 0; r < 23; ++r) {
      const float t = repr[p * 23 + r];
      sx += t * smap[r * 2];
      sy += t * smap[r * 2 + 1];
    }
    sx = fminf(fmaxf(sx, 0.f), 1.f);
    sy = fminf(fmaxf(sy, 0.f), 1.f);
    const float gx = 2.f * sx - 1.f, gy = 2.f * sy - 1.f;
    const float ix = ((gx + 1.f) * kTpsW - 1.f) * 0.5f, iy = ((gy + 1.f) * kTpsH - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const bool vx0 = x0 >= 0 && x0 < kTpsW, vx1 = x0 + 1 >= 0 && x0 + 1 < kTpsW;
    const bool vy0 = y0 >= 0 && y0 < kTpsH, vy1 = y0 + 1 >= 0 && y0 + 1 < kTpsH;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* ch = im + c * kTpsHW;
      const float v00 = (vx0 && vy0) ? ch[y0 * kTpsW + x0] : 0.f;
      const float v01 = (vx1 && vy0) ? ch[y0 * kTpsW + x0 + 1] : 0.f;
      const float v10 = (vx0 && vy1) ? ch[(y0 + 1) * kTpsW + x0] : 0.f;
      const float v11 = (vx1 && vy1) ? ch[(y0 + 1) * kTpsW + x0 + 1] : 0.f;
      out[((long)b * 3 + c) * kTpsHW + p] = wy0 * (wx0 * v00 + wx1 * v01) + wy1 * (wx0 * v10 + wx1 * v11);
    }
  }
}

// d ctrl (B,40) from d out (B,3,16,64): grid_sample backward w.r.t. the grid, clamp mask, then the two
// transposed matmuls reduced over the 1024 pixels of the image.
__global__ void __launch_bounds__(256) tps_bwd_kernel(const float* __restrict__ img, const float* __restrict__ ctrl,
                                                      long ld_ctrl, const float* __restrict__ inv,
                                                      const float* __restrict__ repr, const float* __restrict__ dout,
                                                      float* __restrict__ dctrl, long ld_dctrl) {
  __shared__ float smap[46];
  __shared__ float sds[kTpsHW * 2];  // d src per pixel
  __shared__ float sdm[46];          // d mapping
  const int b = blockIdx.x;
  tps_mapping(inv, ctrl + (long)b * ld_ctrl, smap, threadIdx.x, 256);
  __syncthreads();
  const float* im = img + (long)b * 3 * kTpsHW;
  for (int p = threadIdx.x; p < kTpsHW; p += 256) {
    float sx = 0.f, sy = 0.f;
    for (int r = 0; r < 23; ++r) {
      const float t = repr[p * 23 + r];
      sx += t * smap[r * 2];
      sy += t * smap[r * 2 + 1];
    }
    // torch.clamp backward passes the gradient where min <= x <= max
    const bool px = sx >= 0.f && sx <= 1.f, py = sy >= 0.f && sy <= 1.f;
    sx = fminf(fmaxf(sx, 0.f), 1.f);
    sy = fminf(fmaxf(sy, 0.f), 1.f);
    const float gx = 2.f * sx - 1.f, gy = 2.f * sy - 1.f;
    const float ix = ((gx + 1.f) * kTpsW - 1.f) * 0.5f, iy = ((gy + 1.f) * kTpsH - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const bool vx0 = x0 >= 0 && x0 < kTpsW, vx1 = x0 + 1 >= 0 && x0 + 1 < kTpsW;
    const bool vy0 = y0 >= 0 && y0 < kTpsH, vy1 = y0 + 1 >= 0 && y0 + 1 < kTpsH;
    float dix = 0.f, diy = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* ch = im + c * kTpsHW;
      const float v00 = (vx0 && vy0) ? ch[y0 * kTpsW + x0] : 0.f;
      const float v01 = (vx1 && vy0) ? ch[y0 * kTpsW + x0 + 1] : 0.f;
      const float v10 = (vx0 && vy1) ? ch[(y0 + 1) * kTpsW + x0] : 0.f;
      const float v11 = (vx1 && vy1) ? ch[(y0 + 1) * kTpsW + x0 + 1] : 0.f;
      const float g = dout[((long)b * 3 + c) * kTpsHW + p];
      dix += g * (wy0 * (v01 - v00) + wy1 * (v11 - v10));
      diy += g * (wx0 * (v10 - v00) + wx1 * (v11 - v01));
    }
    // ix = (gx+1)*W/2 - 0.5, gx = 2 sx - 1  =>  d ix / d sx = W
    sds[p * 2] = px ? dix * kTpsW : 0.f;
    sds[p * 2 + 1] = py ? diy * kTpsH : 0.f;
  }
  __syncthreads();
  // d mapping[r][d] = sum_p repr[p][r] * dsrc[p][d] : 46 outputs, each reduced by one group of threads
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < 46; o += 8) {
      const int r = o >> 1, d = o & 1;
      float a = 0.f;
      for (int p = lane; p < kTpsHW; p += 32) a += repr[p * 23 + r] * sds[p * 2 + d];
      a = warp_sum(a);
      if (lane == 0) sdm[o] = a;
    }
  }
  __syncthreads();
  // d ctrl[j][d] = sum_r inv[r][j] * dmap[r][d], j < 20
  for (int i = threadIdx.x; i < 40; i += 256) {
    const int j = i >> 1, d = i & 1;
    float a = 0.f;
    for (int r = 0; r < 23; ++r) a += inv[r * 23 + j] * sdm[r * 2 + d];
    dctrl[(long)b * ld_dctrl + i] = a;
  }
}

// bf16 (rows, ld) -> fp32 (rows, n) and back, for the small FC outputs
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ x, long ld, float* __restrict__ y, long rows, int n) {
  const long tot = rows * n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < tot; i += (long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[(i / n) * ld + (i % n)]);
}
// fp32 (rows, n) -> bf16 (rows_pad, ld) zero padded
__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ x, int n, long rows, bf16* __restrict__ y, long ld,
                                       long rows_pad) {
  const long tot = rows_pad * ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < tot; i += (long)gridDim.x * blockDim.x) {
    const long r = i / ld;
    const int c = (int)(i % ld);
    y[i] = __float2bfloat16_rn((r < rows && c < n) ? x[r * n + c] : 0.f);
  }
}

// FC weight preps.  fc1: the reference flattens NCHW (B,256,1,2) -> feature f = c*2 + w, our activations are
// NHWC -> column w*256 + c.   o[n][w*256+c] = w1[n][c*2+w]  (bf16), and the inverse scatter for the gradient.
__global__ void prep_fc1_kernel(const float* __restrict__ w1, bf16* __restrict__ o, bf16* __restrict__ ot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 512 * 512) return;
  const int col = i & 511, n = i >> 9;
  const int w = col >> 8, c = col & 255;
  const bf16 v = __float2bfloat16_rn(w1[n * 512 + c * 2 + w]);
  o[i] = v;                 // [n][col]  (forward B operand)
  ot[col * 512 + n] = v;    // [col][n]  (dgrad B operand)
}
__global__ void unperm_fc1_grad_kernel(const float* __restrict__ g, float* __restrict__ dw1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 512 * 512) return;
  const int col = i & 511, n = i >> 9;
  const int w = col >> 8, c = col & 255;
  dw1[n * 512 + c * 2 + w] = g[i];
}
// fc2: y = fc2(0.1 * feat) (stn_head.py:93): fold the 0.1 into the bf16 weight; rows padded 40 -> 64
__global__ void prep_fc2_kernel(const float* __restrict__ w2, bf16* __restrict__ o, bf16* __restrict__ ot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 512) return;
  const int k = i & 511, n = i >> 9;
  const bf16 v = __float2bfloat16_rn(n < 40 ? 0.1f * w2[n * 512 + k] : 0.f);
  o[i] = v;              // [n 64][k 512]
  ot[k * 64 + n] = v;    // [k 512][n 64]
}
// generic conv weight for the im2col GEMM: o[n][tap*C + c] = w[n][c][tap], K padded to Kpad, N padded to Npad;
// ot = transpose [Kpad][Npad]
__global__ void prep_stn_conv_w_kernel(const float* __restrict__ w, bf16* __restrict__ o, bf16* __restrict__ ot, int Co,
                                       int C, int Npad, int Kpad) {
  const long n_el = (long)Npad * Kpad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_el; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad), n = (int)(i / Kpad);
    float v = 0.f;
    if (n < Co && k < 9 * C) {
      const int tap = k / C, c = k - tap * C;
      v = w[((long)n * C + c) * 9 + tap];
    }
    const bf16 bv = __float2bfloat16_rn(v);
    o[i] = bv;
    if (ot) ot[(long)k * Npad + n] = bv;
  }
}
// dw[n][c][tap] = g[n][tap*C + c]   (g fp32 [Npad][Kpad])
__global__ void unpack_stn_conv_grad_kernel(const float* __restrict__ g, float* __restrict__ dw, int Co, int C, int Kpad) {
  const long n_el = (long)Co * C * 9;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_el; i += (long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9), c = (int)((i / 9) % C), n = (int)(i / (9L * C));
    dw[i] = g[(long)n * Kpad + tap * C + c];
  }
}

int sgrid(long n, int per) {
  long g = (n + per - 1) / per;
  if (g > 148L * 8) g = 148L * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

int im2col3x3(const bf16* x, const float* x_nchw, bf16* col, int B, int H, int W, int C, int Kpad, cudaStream_t s) {
  FOCR_REQUIRE(Kpad % 8 == 0 && (x_nchw != nullptr || C % 8 == 0), "im2col3x3: C=%d Kpad=%d", C, Kpad);
  im2col3x3_kernel<<<sgrid((long)B * H * W * (Kpad / 8), 256), 256, 0, s>>>(x, x_nchw, col, B, H, W, C, Kpad);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int col2im3x3(const bf16* dcol, bf16* dx, int B, int H, int W, int C, int Kpad, cudaStream_t s) {
  col2im3x3_kernel<<<sgrid((long)B * H * W * (C / 8), 256), 256, 0, s>>>(dcol, dx, B, H, W, C, Kpad);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int maxpool_fwd(const bf16* x, bf16* y, int B, int H, int W, int C, int ph, cudaStream_t s) {
  maxpool_fwd_kernel<<<sgrid((long)B * (H / ph) * (W / 2) * (C / 8), 256), 256, 0, s>>>(x, y, B, H, W, C, ph);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int maxpool_bwd(const bf16* x, const bf16* y, const bf16* dy, bf16* dx, int B, int H, int W, int C, int ph,
                cudaStream_t s) {
  maxpool_bwd_kernel<<<sgrid((long)B * (H / ph) * (W / 2) * C, 256), 256, 0, s>>>(x, y, dy, dx, B, H, W, C, ph);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int tps_forward(const float* img, const float* ctrl, long ld_ctrl, const float* inv, const float* repr, float* out,
                int B, cudaStream_t s) {
  tps_fwd_kernel<<<B, 256, 0, s>>>(img, ctrl, ld_ctrl, inv, repr, out);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int tps_backward(const float* img, const float* ctrl, long ld_ctrl, const float* inv, const float* repr,
                 const float* dout, float* dctrl, long ld_dctrl, int B, cudaStream_t s) {
  tps_bwd_kernel<<<B, 256, 0, s>>>(img, ctrl, ld_ctrl, inv, repr, dout, dctrl, ld_dctrl);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int bf16_to_f32(const bf16* x, long ld, float* y, long rows, int n, cudaStream_t s) {
  bf16_to_f32_kernel<<<sgrid(rows * n, 256), 256, 0, s>>>(x, ld, y, rows, n);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int f32_to_bf16_pad(const float* x, int n, long rows, bf16* y, long ld, long rows_pad, cudaStream_t s) {
  f32_to_bf16_pad_kernel<<<sgrid(rows_pad * ld, 256), 256, 0, s>>>(x, n, rows, y, ld, rows_pad);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int prep_fc1(const float* w1, bf16* o, bf16* ot, cudaStream_t s) {
  prep_fc1_kernel<<<focr_cdiv(512 * 512, 256), 256, 0, s>>>(w1, o, ot);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int unperm_fc1_grad(const float* g, float* dw1, cudaStream_t s) {
  unperm_fc1_grad_kernel<<<focr_cdiv(512 * 512, 256), 256, 0, s>>>(g, dw1);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int prep_fc2(const float* w2, bf16* o, bf16* ot, cudaStream_t s) {
  prep_fc2_kernel<<<focr_cdiv(64 * 512, 256), 256, 0, s>>>(w2, o, ot);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int prep_stn_conv_w(const float* w, bf16* o, bf16* ot, int Co, int C, int Npad, int Kpad, cudaStream_t s) {
  prep_stn_conv_w_kernel<<<sgrid((long)Npad * Kpad, 256), 256, 0, s>>>(w, o, ot, Co, C, Npad, Kpad);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int unpack_stn_conv_grad(const float* g, float* dw, int Co, int C, int Kpad, cudaStream_t s) {
  unpack_stn_conv_grad_kernel<<<sgrid((long)Co * C * 9, 256), 256, 0, s>>>(g, dw, Co, C, Kpad);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
