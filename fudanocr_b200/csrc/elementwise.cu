// HBM-bound kernels of the TBSRN hot path: train-mode BatchNorm (stats / apply / backward) with the
// mish / relu activations and the positional-encoding concat folded in, the reference's home-made
// LayerNorm (unbiased std, eps on the std), column sums for bias gradients, PReLU backward,
// tanh + MSE.  All activations are NHWC bf16 = a (rows, C) matrix; every thread moves 16 bytes.
//
// Reference semantics: scene-text-telescope/model/tbsrn.py:23-36 (LayerNorm), :233,238,191
// (BatchNorm2d train), :277-285 (mish), :83-86 (pos-enc concat), :180-182 (PReLU), :225 (tanh),
// loss/text_focus_loss.py:86 (MSE).
#include "kernels.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kLnRows = 4;     // rows per 16-lane group per loop iteration of the LayerNorm forward
constexpr int kLnBwdRows = 1;  // ... of the backward (2 or 4 rows cost 92 / 127 registers and measured 89 us against 66 us for 1)

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float act_fwd(float y, int act) {
  if (act == ACT_MISH) return mish_f(y);
  if (act == ACT_RELU) return fmaxf(y, 0.f);
  return y;
}
__device__ __forceinline__ float act_bwd(float y, int act) {
  if (act == ACT_MISH) return mish_grad_f(y);
  if (act == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  return 1.f;
}

// ---------------------------------------------------------------------------------------------
// per-channel partial sums: partial[blk][0][c] = sum x, partial[blk][1][c] = sum x^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const bf16* __restrict__ x, long ld, long T, int C,
                                                            float* __restrict__ partial) {
  __shared__ float red[2][kThreads][8];
  const int tpr = C >> 3;
  const int rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
#pragma unroll 4
  for (long t = (long)blockIdx.x * rpb + rl; t < T; t += (long)gridDim.x * rpb) {
    float v[8];
    load8(x + t * ld + cg * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += v[j];
      q[j] += v[j] * v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = s[j];
    red[1][threadIdx.x][j] = q[j];
  }
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int r = 1; r < rpb; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += red[0][threadIdx.x + r * tpr][j];
        q[j] += red[1][threadIdx.x + r * tpr][j];
      }
    }
    float* o = partial + (long)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[cg * 8 + j] = s[j];
      o[C + cg * 8 + j] = q[j];
    }
  }
}

// stats[0]=mean [1]=invstd [2]=scale=gamma*invstd [3]=shift=beta-mean*scale ; running stats updated
// as nn.BatchNorm does in train mode (momentum 0.1, unbiased variance).
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// lane's share of sum_p base[p * stride], eight loads in flight (the plain loop waits one L2 round trip per partial: the
// finalize kernels cost as much as the streaming passes they close, ncu launch list of round 1: 9 us vs 10 us)
__device__ __forceinline__ double strided_sum(const float* __restrict__ base, int P, long stride, int lane) {
  double s = 0.0;
  int p = lane;
  for (; p + 32 * 7 < P; p += 32 * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = base[(long)(p + 32 * u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; p < P; p += 32) s += base[(long)p * stride];
  return s;
}

// second-stage reductions: one WARP per output element (lanes stride over the P partials) so the stage is
// a handful of microseconds instead of a serial chain of P dependent L2 round trips
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int P, int C, long T,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float eps, float momentum,
                                   float* __restrict__ stats) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = strided_sum(partial + c, P, 2L * C, lane);
  double q = strided_sum(partial + C + c, P, 2L * C, lane);
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if (lane != 0) return;
  const double mean = s / (double)T;
  double var = q / (double)T - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  stats[c] = (float)mean;
  stats[C + c] = invstd;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = beta[c] - (float)mean * sc;
  if (running_mean != nullptr) {
    const double unb = T > 1 ? var * (double)T / (double)(T - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    if (c == 0 && nbt != nullptr) *nbt += 1;
  }
}

// eval-mode BatchNorm: stats from the running buffers
__global__ void bn_eval_stats_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                     int C, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(rv[c] + eps);
  const float sc = gamma[c] * invstd;
  stats[c] = rm[c];
  stats[C + c] = invstd;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = beta[c] - rm[c] * sc;
}

// out[t][c] = act(x*scale+shift); optional right half out[t][C+c] = pe[t % pe_rows][c]
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const bf16* __restrict__ x, long ld_x,
                                                            const float* __restrict__ stats, bf16* __restrict__ out,
                                                            long ld_out, long T, int C, int act,
                                                            const bf16* __restrict__ pe, int pe_rows,
                                                            const bf16* __restrict__ res) {
  const int tpr = C >> 3;
  const int rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = stats[2 * C + cg * 8 + j];
    sh[j] = stats[3 * C + cg * 8 + j];
  }
#pragma unroll 4
  for (long t = (long)blockIdx.x * rpb + rl; t < T; t += (long)gridDim.x * rpb) {
    float v[8];
    load8(x + t * ld_x + cg * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_fwd(fmaf(v[j], sc[j], sh[j]), act);
    if (res != nullptr) {
      float r[8];
      load8(res + t * ld_out + cg * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    store8(out + t * ld_out + cg * 8, v);
    if (pe != nullptr) {
      const uint4 u = *reinterpret_cast<const uint4*>(pe + (long)(t % pe_rows) * C + cg * 8);
      *reinterpret_cast<uint4*>(out + t * ld_out + C + cg * 8) = u;
    }
  }
}

// backward pass 1: g = dy * act'(bn(x)); partial[blk][0][c] = sum g, [1][c] = sum g * xhat
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, long ld_dy,
                                                                 const bf16* __restrict__ x, long ld_x,
                                                                 const float* __restrict__ stats, long T, int C,
                                                                 int act, float* __restrict__ partial) {
  __shared__ float red[2][kThreads][8];
  const int tpr = C >> 3;
  const int rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float mean[8], istd[8], sc[8], sh[8], s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = stats[cg * 8 + j];
    istd[j] = stats[C + cg * 8 + j];
    sc[j] = stats[2 * C + cg * 8 + j];
    sh[j] = stats[3 * C + cg * 8 + j];
    s[j] = q[j] = 0.f;
  }
#pragma unroll 4
  for (long t = (long)blockIdx.x * rpb + rl; t < T; t += (long)gridDim.x * rpb) {
    float xv[8], gv[8];
    load8(x + t * ld_x + cg * 8, xv);
    load8(dy + t * ld_dy + cg * 8, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = gv[j] * act_bwd(fmaf(xv[j], sc[j], sh[j]), act);
      s[j] += g;
      q[j] += g * (xv[j] - mean[j]) * istd[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = s[j];
    red[1][threadIdx.x][j] = q[j];
  }
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int r = 1; r < rpb; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += red[0][threadIdx.x + r * tpr][j];
        q[j] += red[1][threadIdx.x + r * tpr][j];
      }
    }
    float* o = partial + (long)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[cg * 8 + j] = s[j];
      o[C + cg * 8 + j] = q[j];
    }
  }
}

// coef[0][c] = sum g / T, coef[1][c] = sum g xhat / T; dgamma = sum g xhat, dbeta = sum g
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int P, int C, long T,
                                       float* __restrict__ coef, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = strided_sum(partial + c, P, 2L * C, lane);
  double q = strided_sum(partial + C + c, P, 2L * C, lane);
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  if (lane != 0) return;
  coef[c] = (float)(s / (double)T);
  coef[C + c] = (float)(q / (double)T);
  if (dgamma) dgamma[c] = (float)q;
  if (dbeta) dbeta[c] = (float)s;
}

// backward pass 2: dx = gamma*invstd * (g - mean(g) - xhat * mean(g xhat))
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const bf16* __restrict__ dy, long ld_dy,
                                                                const bf16* __restrict__ x, long ld_x,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ coef, bf16* __restrict__ dx,
                                                                long ld_dx, long T, int C, int act,
                                                                float* __restrict__ colsum_partial) {
  // colsum_partial (optional): per-CTA column sums of the dx values as stored (bf16-rounded) = the gradient of the bias of
  // the convolution in front of this BatchNorm, so that tensor is not read again by a separate column-sum pass
  __shared__ float cred[kThreads][8];
  const int tpr = C >> 3;
  const int rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float mean[8], istd[8], sc[8], sh[8], c1[8], c2[8], cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = stats[cg * 8 + j];
    istd[j] = stats[C + cg * 8 + j];
    sc[j] = stats[2 * C + cg * 8 + j];
    sh[j] = stats[3 * C + cg * 8 + j];
    c1[j] = coef[cg * 8 + j];
    c2[j] = coef[C + cg * 8 + j];
  }
#pragma unroll 4
  for (long t = (long)blockIdx.x * rpb + rl; t < T; t += (long)gridDim.x * rpb) {
    float xv[8], gv[8];
    load8(x + t * ld_x + cg * 8, xv);
    load8(dy + t * ld_dy + cg * 8, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = gv[j] * act_bwd(fmaf(xv[j], sc[j], sh[j]), act);
      const float xh = (xv[j] - mean[j]) * istd[j];
      gv[j] = sc[j] * (g - c1[j] - xh * c2[j]);
      cs[j] += __bfloat162float(__float2bfloat16_rn(gv[j]));
    }
    store8(dx + t * ld_dx + cg * 8, gv);
  }
  if (colsum_partial != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) cred[threadIdx.x][j] = cs[j];
    __syncthreads();
    if (threadIdx.x < tpr) {
      for (int r = 1; r < rpb; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += cred[threadIdx.x + r * tpr][j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) colsum_partial[(long)blockIdx.x * C + cg * 8 + j] = cs[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm of the reference (features = 128): y = a (x - mean) / (std_unbiased + eps) + b
// 16 lanes per row, 8 elements per lane.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kThreads) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ a,
                                                          const float* __restrict__ b, bf16* __restrict__ y, long T,
                                                          float eps) {
  const int l16 = threadIdx.x & 15;
  const int rl = threadIdx.x >> 4;
  float av[8], bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    av[j] = a[l16 * 8 + j];
    bv[j] = b[l16 * 8 + j];
  }
  // kLnRows rows per 16-lane group per iteration: all loads of the iteration are issued before the first shuffle
  // reduction, so each thread keeps 4 x 16 B in flight (one row at a time streamed at ~3.3 TB/s, half of HBM rate)
  constexpr int R = kLnRows;
  const long rows_per_pass = (long)gridDim.x * (kThreads / 16) * R;
  for (long t0 = (long)blockIdx.x * (kThreads / 16) * R; t0 < T; t0 += rows_per_pass) {
    float v[R][8];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long t = t0 + rl + k * (kThreads / 16);
      if (t < T) load8(x + t * 128 + l16 * 8, v[k]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long t = t0 + rl + k * (kThreads / 16);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[k][j];
      const float mean = half_warp_sum(s) * (1.f / 128.f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[k][j] -= mean;
        q += v[k][j] * v[k][j];
      }
      const float sd = sqrtf(half_warp_sum(q) * (1.f / 127.f));
      const float inv = 1.f / (sd + eps);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = fmaf(av[j] * v[k][j], inv, bv[j]);
      if (t < T) store8(y + t * 128 + l16 * 8, v[k]);
    }
  }
}

// dx_i = (h_i - mean(h)) / s  -  d_i * (sum_j h_j d_j) / (s^2 (N-1) sigma),  h = a*g, d = x-mean, s = sigma+eps
// partial[blk][0][i] = sum_rows g_i d_i / s (-> da), partial[blk][1][i] = sum_rows g_i (-> db)
__global__ void __launch_bounds__(kThreads) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                          const float* __restrict__ a, bf16* __restrict__ dx,
                                                          float* __restrict__ partial, long T, float eps) {
  __shared__ float red[2][kThreads / 16][128];
  const int l16 = threadIdx.x & 15;
  const int rl = threadIdx.x >> 4;
  float av[8], da[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    av[j] = a[l16 * 8 + j];
    da[j] = db[j] = 0.f;
  }
  constexpr int R = kLnBwdRows;
  const long rows_per_pass = (long)gridDim.x * (kThreads / 16) * R;
  for (long t0 = (long)blockIdx.x * (kThreads / 16) * R; t0 < T; t0 += rows_per_pass) {
    float v[R][8], g[R][8];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long t = t0 + rl + k * (kThreads / 16);
      if (t < T) {
        load8(x + t * 128 + l16 * 8, v[k]);
        load8(dy + t * 128 + l16 * 8, g[k]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = g[k][j] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long t = t0 + rl + k * (kThreads / 16);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[k][j];
      const float mean = half_warp_sum(s) * (1.f / 128.f);
      float q = 0.f, hs = 0.f, hd = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[k][j] -= mean;
        q += v[k][j] * v[k][j];
        const float h = av[j] * g[k][j];
        hs += h;
        hd += h * v[k][j];
      }
      q = half_warp_sum(q);
      hs = half_warp_sum(hs);
      hd = half_warp_sum(hd);
      const float sigma = sqrtf(q * (1.f / 127.f));
      const float sp = sigma + eps;
      const float inv = 1.f / sp;
      const float hm = hs * (1.f / 128.f);
      const float k2 = sigma > 0.f ? hd * inv * inv / (127.f * sigma) : 0.f;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (av[j] * g[k][j] - hm) * inv - v[k][j] * k2;
        da[j] += g[k][j] * v[k][j] * inv;
        db[j] += g[k][j];
      }
      if (t < T) store8(dx + t * 128 + l16 * 8, o);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][rl][l16 * 8 + j] = da[j];
    red[1][rl][l16 * 8 + j] = db[j];
  }
  __syncthreads();
  {
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < kThreads / 16; ++r) acc += red[which][r][col];
    partial[(long)blockIdx.x * 256 + which * 128 + col] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients): partial[blk][c] = sum_t x[t][c]
// ---------------------------------------------------------------------------------------------
// row t lives at x + (t / inner) * stride_outer + (t % inner) * ld   (inner == 0: plain row stride ld)
__global__ void __launch_bounds__(kThreads) colsum_kernel(const bf16* __restrict__ x, long ld, long T, int C,
                                                          long inner, long stride_outer,
                                                          float* __restrict__ partial) {
  __shared__ float red[kThreads][8];
  const int tpr = C >> 3;
  const int rpb = kThreads / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (rl < rpb) {
  #pragma unroll 4
  for (long t = (long)blockIdx.x * rpb + rl; t < T; t += (long)gridDim.x * rpb) {
      float v[8];
      const long off = inner > 0 ? (t / inner) * stride_outer + (t % inner) * ld : t * ld;
      load8(x + off + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int r = 1; r < rpb; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += red[threadIdx.x + r * tpr][j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) partial[(long)blockIdx.x * C + cg * 8 + j] = s[j];
  }
}

// out[i] = scale * sum_p partial[p*stride + i]   (deterministic second stage of every reduction)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int P, long stride, int n,
                                       float* __restrict__ out, float scale) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  double s = strided_sum(partial + i, P, stride, lane);
  s = warp_sum_d(s);
  if (lane == 0) out[i] = (float)s * scale;
}

// ---------------------------------------------------------------------------------------------
// PReLU backward (single slope): dx = dy * (x > 0 ? 1 : a); da = sum dy * x * (x <= 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) prelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                             const float* __restrict__ a, bf16* __restrict__ dx,
                                                             long n8, float* __restrict__ partial) {
  __shared__ float red[kThreads / 32];
  const float slope = a[0];
  float acc = 0.f;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (long)gridDim.x * kThreads) {
    float xv[8], gv[8];
    load8(x + i * 8, xv);
    load8(dy + i * 8, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (xv[j] <= 0.f) {
        acc += gv[j] * xv[j];
        gv[j] *= slope;
      }
    }
    if (dx != nullptr) store8(dx + i * 8, gv);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// tanh head + MSE:  sr = tanh(o);  loss partial = sum (sr-hr)^2;  do = gscale * 2 (sr-hr)/N * (1-sr^2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) tanh_mse_kernel(const float* __restrict__ o, const float* __restrict__ hr,
                                                            float* __restrict__ sr, float* __restrict__ dout, long n,
                                                            float gscale, float* __restrict__ partial) {
  __shared__ float red[kThreads / 32];
  float acc = 0.f;
  const float k = gscale * 2.f / (float)n;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long)gridDim.x * kThreads) {
    const float s = tanhf(o[i]);
    if (sr) sr[i] = s;
    if (hr) {
      const float d = s - hr[i];
      acc += d * d;
      if (dout) dout[i] = k * d * (1.f - s * s);
    }
  }
  if (partial) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) s += red[w];
      partial[blockIdx.x] = s;
    }
  }
}

// dO = dSR * (1 - sr^2)   (autograd path: the loss lives in the caller)
__global__ void tanh_bwd_kernel(const float* __restrict__ sr, const float* __restrict__ dsr, float* __restrict__ dout,
                                long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float s = sr[i];
    dout[i] = dsr[i] * (1.f - s * s);
  }
}

// out = a + b (bf16, 8 per thread)
__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, long n8) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    load8(a + i * 8, x);
    load8(b + i * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    store8(o + i * 8, x);
  }
}

// dx = dy * mish'(x)  (backward of the activation after PixelShuffle, tbsrn.py:272)
__global__ void mish_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx, long n8) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float g[8], v[8];
    load8(dy + i * 8, g);
    load8(x + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= mish_grad_f(v[j]);
    store8(dx + i * 8, g);
  }
}

// positionalencoding2d(64,16,64) (tbsrn.py:39-61) as a (1024, 64) bf16 token-major table
__global__ void pe_table_kernel(bf16* __restrict__ pe) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1024 * 64) return;
  const int c = i & 63, t = i >> 6;
  const int h = t >> 6, w = t & 63;
  const int half = 32;
  const int cc = c < half ? c : c - half;
  const float pos = c < half ? (float)w : (float)h;
  const float div = expf((float)(cc & ~1) * -(logf(10000.f) / (float)half));
  const float v = (cc & 1) ? cosf(pos * div) : sinf(pos * div);
  pe[i] = __float2bfloat16_rn(v);
}

// grid for reduction passes: 8 CTAs per SM keep enough loads in flight to stream at HBM rate; the second stage is
// warp-parallel, so summing up to 1184 partials per output costs ~40 loads per lane
int ctas_per_sm() {  // tuning knob FOCR_EW_CTAS_PER_SM (measured at T = 262144: 4 is 5-9 % faster than 8, 16+ slower)
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("FOCR_EW_CTAS_PER_SM");
    v = e ? atoi(e) : 4;
    if (v < 1 || v > 32) v = 4;
  }
  return v;
}
int red_grid(long work_items, int per_block) {
  long g = (work_items + per_block - 1) / per_block;
  const long cap = 148L * ctas_per_sm();
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int ew_grid(long work_items, int per_block) {
  long g = (work_items + per_block - 1) / per_block;
  const long cap = 148L * ctas_per_sm();
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int bn_partial_blocks(long T, int C) {
  const int rpb = kThreads / (C >> 3);
  return red_grid(T, rpb * 4);
}

int bn_train_stats(const bf16* x, long ld, long T, int C, const float* gamma, const float* beta, float* rm, float* rv,
                   long long* nbt, float eps, float momentum, float* partial, float* stats, cudaStream_t s) {
  ProfScope _ps("bn_stats", s);
  FOCR_REQUIRE(C % 8 == 0 && kThreads % (C >> 3) == 0 && C <= 2048, "bn: unsupported C=%d", C);
  const int P = bn_partial_blocks(T, C);
  bn_stats_kernel<<<P, kThreads, 0, s>>>(x, ld, T, C, partial);
  FOCR_LAUNCH_CHECK();
  bn_finalize_kernel<<<focr_cdiv(C, 4), 128, 0, s>>>(partial, P, C, T, gamma, beta, rm, rv, nbt, eps, momentum,
                                                        stats);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int bn_eval_stats(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C,
                  float* stats, cudaStream_t s) {
  bn_eval_stats_kernel<<<focr_cdiv(C, 128), 128, 0, s>>>(gamma, beta, rm, rv, eps, C, stats);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int bn_apply(const bf16* x, long ld_x, const float* stats, bf16* out, long ld_out, long T, int C, int act,
             const bf16* pe, int pe_rows, const bf16* res, cudaStream_t s) {
  ProfScope _ps("bn_apply", s);
  FOCR_REQUIRE(C % 8 == 0 && kThreads % (C >> 3) == 0, "bn_apply: unsupported C=%d", C);
  const int rpb = kThreads / (C >> 3);
  bn_apply_kernel<<<ew_grid(T, rpb * 2), kThreads, 0, s>>>(x, ld_x, stats, out, ld_out, T, C, act, pe, pe_rows, res);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int bn_backward(const bf16* dy, long ld_dy, const bf16* x, long ld_x, const float* stats, bf16* dx, long ld_dx, long T,
                int C, int act, float* dgamma, float* dbeta, float* partial, float* coef, cudaStream_t s, float* dx_colsum) {
  ProfScope _ps("bn_bwd", s);
  FOCR_REQUIRE(C % 8 == 0 && kThreads % (C >> 3) == 0, "bn_backward: unsupported C=%d", C);
  const int P = bn_partial_blocks(T, C);
  bn_bwd_reduce_kernel<<<P, kThreads, 0, s>>>(dy, ld_dy, x, ld_x, stats, T, C, act, partial);
  FOCR_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<focr_cdiv(C, 4), 128, 0, s>>>(partial, P, C, T, coef, dgamma, dbeta);
  FOCR_LAUNCH_CHECK();
  const int rpb = kThreads / (C >> 3);
  const int G = ew_grid(T, rpb * 2);
  bn_bwd_apply_kernel<<<G, kThreads, 0, s>>>(dy, ld_dy, x, ld_x, stats, coef, dx, ld_dx, T, C, act,
                                             dx_colsum ? partial : nullptr);  // `partial` is free again after the finalize
  FOCR_LAUNCH_CHECK();
  if (dx_colsum) {
    reduce_partials_kernel<<<focr_cdiv(C, 4), 128, 0, s>>>(partial, G, C, C, dx_colsum, 1.f);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

int ln_partial_blocks(long T) { return red_grid(T, (kThreads / 16) * kLnBwdRows); }

int ln_forward(const bf16* x, const float* a, const float* b, bf16* y, long T, float eps, cudaStream_t s) {
  ProfScope _ps("ln_fwd", s);
  ln_fwd_kernel<<<ew_grid(T, (kThreads / 16) * kLnRows), kThreads, 0, s>>>(x, a, b, y, T, eps);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int ln_backward(const bf16* dy, const bf16* x, const float* a, bf16* dx, float* da, float* db, float* partial, long T,
                float eps, cudaStream_t s) {
  ProfScope _ps("ln_bwd", s);
  const int P = ln_partial_blocks(T);
  ln_bwd_kernel<<<P, kThreads, 0, s>>>(dy, x, a, dx, partial, T, eps);
  FOCR_LAUNCH_CHECK();
  reduce_partials_kernel<<<32, 128, 0, s>>>(partial, P, 256, 128, da, 1.f);
  FOCR_LAUNCH_CHECK();
  reduce_partials_kernel<<<32, 128, 0, s>>>(partial + 128, P, 256, 128, db, 1.f);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int colsum_partial_blocks(long T, int C) { return red_grid(T, (kThreads / (C >> 3)) * 8); }

int colsum(const bf16* x, long ld, long T, int C, float* out, float* partial, cudaStream_t s) {
  ProfScope _ps("colsum", s);
  FOCR_REQUIRE(C % 8 == 0 && (C >> 3) <= kThreads, "colsum: unsupported C=%d", C);
  const int P = colsum_partial_blocks(T, C);
  colsum_kernel<<<P, kThreads, 0, s>>>(x, ld, T, C, 0, 0, partial);
  FOCR_LAUNCH_CHECK();
  reduce_partials_kernel<<<focr_cdiv(C, 4), 128, 0, s>>>(partial, P, C, C, out, 1.f);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int colsum2(const bf16* x, long t_outer, long t_inner, long stride_outer, long stride_inner, int C, float* out,
            float* partial, cudaStream_t s) {
  FOCR_REQUIRE(C % 8 == 0 && (C >> 3) <= kThreads, "colsum2: unsupported C=%d", C);
  const long T = t_outer * t_inner;
  const int P = colsum_partial_blocks(T, C);
  colsum_kernel<<<P, kThreads, 0, s>>>(x, stride_inner, T, C, t_inner, stride_outer, partial);
  FOCR_LAUNCH_CHECK();
  reduce_partials_kernel<<<focr_cdiv(C, 4), 128, 0, s>>>(partial, P, C, C, out, 1.f);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int reduce_partials(const float* partial, int P, long stride, int n, float* out, float scale, cudaStream_t s) {
  reduce_partials_kernel<<<focr_cdiv(n, 4), 128, 0, s>>>(partial, P, stride, n, out, scale);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int prelu_backward(const bf16* dy, const bf16* x, const float* a, bf16* dx, long n, float* da, float* partial,
                   cudaStream_t s) {
  FOCR_REQUIRE(n % 8 == 0, "prelu_backward: n %% 8");
  const int P = ew_grid(n / 8, kThreads * 4);
  prelu_bwd_kernel<<<P, kThreads, 0, s>>>(dy, x, a, dx, n / 8, partial);
  FOCR_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, 32, 0, s>>>(partial, P, 1, 1, da, 1.f);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int tanh_mse(const float* o, const float* hr, float* sr, float* dout, long n, float gscale, float* loss,
             float* partial, cudaStream_t s) {
  const int P = ew_grid(n, kThreads * 4);
  tanh_mse_kernel<<<P, kThreads, 0, s>>>(o, hr, sr, dout, n, gscale, (hr && loss) ? partial : nullptr);
  FOCR_LAUNCH_CHECK();
  if (hr && loss) {
    reduce_partials_kernel<<<1, 32, 0, s>>>(partial, P, 1, 1, loss, 1.f / (float)n);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

int tanh_backward(const float* sr, const float* dsr, float* dout, long n, cudaStream_t s) {
  tanh_bwd_kernel<<<ew_grid(n, 256 * 4), 256, 0, s>>>(sr, dsr, dout, n);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int add_bf16(const bf16* a, const bf16* b, bf16* o, long n, cudaStream_t s) {
  FOCR_REQUIRE(n % 8 == 0, "add_bf16: n %% 8");
  add_bf16_kernel<<<ew_grid(n / 8, 256 * 2), 256, 0, s>>>(a, b, o, n / 8);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int mish_backward(const bf16* dy, const bf16* x, bf16* dx, long n, cudaStream_t s) {
  FOCR_REQUIRE(n % 8 == 0, "mish_backward: n %% 8");
  mish_bwd_kernel<<<ew_grid(n / 8, 256 * 2), 256, 0, s>>>(dy, x, dx, n / 8);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int pe_table(bf16* pe, cudaStream_t s) {
  pe_table_kernel<<<256, 256, 0, s>>>(pe);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// ---------------------------------------------------------------------------------------------
// MSE loss head on the tanh output: loss = mean((sr-hr)^2), d_sr = gscale * 2 (sr-hr) / n
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) mse_loss_grad_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                                            float* __restrict__ d_sr, long n, float k,
                                                            float* __restrict__ partial) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float d = sr[i] - hr[i];
    acc += d * d;
    if (d_sr) d_sr[i] = k * d;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}
}  // namespace

extern "C" int focr_mse_loss_grad(const float* sr, const float* hr, float* d_sr, float* loss, long n, float gscale,
                                  void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int P = ew_grid(n, 256 * 4);
  FOCR_REQUIRE(ws_bytes >= (size_t)P * 4, "mse_loss_grad: workspace too small");
  mse_loss_grad_kernel<<<P, 256, 0, s>>>(sr, hr, d_sr, n, gscale * 2.f / (float)n, (float*)ws);
  FOCR_LAUNCH_CHECK();
  reduce_partials_kernel<<<1, 32, 0, s>>>((const float*)ws, P, 1, 1, loss, 1.f / (float)n);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
