// Weight-gradient kernels: the reduction dimension is the token / pixel axis (T = B*1024 ... B*4096),
// the outputs are tiny (<= 384x128 or 9x64x64), so these are streaming kernels: read dY and X once,
// accumulate in registers with mma.sync (ldmatrix.trans gives the transposed operands for free and
// per-row shared-memory addressing gives the 3x3 tap shifts for free), write one partial per CTA and
// finish with a deterministic second-stage reduction (no atomics -> bitwise run-to-run reproducible).
//
// Reference ops: autograd of nn.Linear (tbsrn.py:74,103,158-159) and nn.Conv2d 3x3 (tbsrn.py:190,232,
// 237,264).
#include "kernels.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// linear wgrad: dW[n][k] = sum_t dY[t][n] * X[t][k]     (K = 128 columns of X per CTA)
// ---------------------------------------------------------------------------------------------
constexpr int kLwChunk = 32;  // token rows per pipeline stage
constexpr int kLwStages = 4;
constexpr int kLwPad = 8;     // bf16 elements of row padding (16 B) -> conflict-free ldmatrix

template <int BN>  // rows of dW per CTA (64 or 128)
struct LwCfg {
  static constexpr int kLdA = BN + kLwPad;
  static constexpr int kLdB = 128 + kLwPad;
  static constexpr int kStageElems = kLwChunk * (kLdA + kLdB);
  static constexpr int kSmemBytes = kLwStages * kStageElems * 2;
  static constexpr int kWarpsN = BN / 64;      // warps along dW rows (64 rows each)
  static constexpr int kWarpsK = 8 / kWarpsN;  // warps along dW cols
  static constexpr int kWN = 128 / kWarpsK;    // dW cols per warp (32 or 16)
  static constexpr int kNT = kWN / 8;          // n-tiles per warp
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
linear_wgrad_kernel(const bf16* __restrict__ dy, long ld_dy, const bf16* __restrict__ x, long ld_x, long T,
                    int n_total, float* __restrict__ partial) {
  using Cfg = LwCfg<BN>;
  extern __shared__ __align__(128) uint8_t sm[];
  bf16* smem = reinterpret_cast<bf16*>(sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_blk = blockIdx.x;       // which BN-row block of dW
  const int split = blockIdx.y, nsplit = gridDim.y;
  x += (long)blockIdx.z * 128;         // which 128-column block of X (K > 128: all blocks in one launch)
  partial += (long)blockIdx.z * nsplit * n_total * 128;
  const long chunks_total = (T + kLwChunk - 1) / kLwChunk;
  const long c_begin = chunks_total * split / nsplit, c_end = chunks_total * (split + 1) / nsplit;
  const int wn = warp % Cfg::kWarpsN, wk = warp / Cfg::kWarpsN;

  float acc[4][Cfg::kNT][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::kNT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  auto issue = [&](long chunk, int stage) {
    bf16* sa = smem + stage * Cfg::kStageElems;
    bf16* sb = sa + kLwChunk * Cfg::kLdA;
    const long t0 = chunk * kLwChunk;
    // A: dY rows, BN columns starting at n_blk*BN
    for (int i = tid; i < kLwChunk * (BN / 8); i += 256) {
      const int r = i / (BN / 8), cc = i % (BN / 8);
      const long t = t0 + r;
      cp_async_16(smem_u32(sa + r * Cfg::kLdA + cc * 8), dy + t * ld_dy + (long)n_blk * BN + cc * 8, t < T);
    }
    for (int i = tid; i < kLwChunk * 16; i += 256) {
      const int r = i >> 4, cc = i & 15;
      const long t = t0 + r;
      cp_async_16(smem_u32(sb + r * Cfg::kLdB + cc * 8), x + t * ld_x + cc * 8, t < T);
    }
  };

  const long nchunks = c_end - c_begin;
  for (int s = 0; s < kLwStages - 1; ++s) {
    if (s < nchunks) issue(c_begin + s, s);
    cp_async_commit();
  }
  for (long ci = 0; ci < nchunks; ++ci) {
    cp_async_wait<kLwStages - 2>();
    __syncthreads();
    {
      const long nxt = ci + kLwStages - 1;
      if (nxt < nchunks) issue(c_begin + nxt, (int)(nxt % kLwStages));
      cp_async_commit();
    }
    const bf16* sa = smem + (ci % kLwStages) * Cfg::kStageElems;
    const bf16* sb = sa + kLwChunk * Cfg::kLdA;
#pragma unroll
    for (int ks = 0; ks < kLwChunk / 16; ++ks) {
      uint32_t a[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        // blocks: lanes 0-7: t 0..7 @ col m0; 8-15: t 0..7 @ m0+8; 16-23: t 8..15 @ m0; 24-31: t 8..15 @ m0+8
        const int trow = ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int col = wn * 64 + mt * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(a[mt], smem_u32(sa + trow * Cfg::kLdA + col));
      }
#pragma unroll
      for (int np = 0; np < Cfg::kNT / 2; ++np) {
        // lanes 0-7: t 0..7 @ n0; 8-15: t 8..15 @ n0; 16-23: t 0..7 @ n0+8; 24-31: t 8..15 @ n0+8
        uint32_t r[4];
        const int trow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = wk * Cfg::kWN + np * 16 + ((lane >> 4) & 1) * 8;
        ldmatrix_x4_trans(r, smem_u32(sb + trow * Cfg::kLdB + col));
        const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          mma_bf16_16816(acc[mt][2 * np], a[mt], b0);
          mma_bf16_16816(acc[mt][2 * np + 1], a[mt], b1);
        }
      }
    }
  }
  cp_async_wait<0>();
  // partial[split][n][k]
  const int g = lane >> 2, c = lane & 3;
  float* out = partial + ((long)split * n_total + (long)n_blk * BN + wn * 64) * 128;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
    for (int nt = 0; nt < Cfg::kNT; ++nt) {
      const int col = wk * Cfg::kWN + nt * 8 + 2 * c;
      float* r0 = out + (long)(mt * 16 + g) * 128 + col;
      *reinterpret_cast<float2*>(r0) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<float2*>(r0 + 8 * 128) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// conv wgrad (KH x KW taps, stride 1, "same" padding), 64 input channels x 64 output channels per CTA:
//   dW[tap][co][ci] = sum_pixels dY[p][co] * X[p + tap][ci]
// Unit of work: 2 output rows x 64 output columns of one image.  The X halo ((2+KH-1) x (64+KW-1) px)
// and the dY tile (128 px) live in shared memory with a 16-byte-chunk XOR swizzle on the pixel index.
// ---------------------------------------------------------------------------------------------
constexpr int kCwRows = 2;
template <int KH, int KW>
struct CwCfg {
  static constexpr int kHaloW = 64 + KW - 1;
  static constexpr int kXPix = (kCwRows + KH - 1) * kHaloW;
  static constexpr int kYPix = kCwRows * 64;
  static constexpr int kStageBytes = (kXPix + kYPix) * 128;
  static constexpr int kSmemBytes = 2 * kStageBytes;
  static constexpr int kTaps = KH * KW;
};

__device__ __forceinline__ uint32_t cw_off(int pix, int chunk) { return (uint32_t)(pix * 128 + ((chunk ^ (pix & 7)) << 4)); }

template <int KH, int KW>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const bf16* __restrict__ dy, long dy_pix, long dy_row, long dy_img,
                  const bf16* __restrict__ x, int B, int H, int W, long sub_stride /*elements between co-groups of dy*/,
                  float* __restrict__ partial) {
  using Cfg = CwCfg<KH, KW>;
  constexpr int kTaps = Cfg::kTaps;
  extern __shared__ __align__(128) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = blockIdx.y;  // output-channel group of 64
  const bf16* dyg = dy + (long)grp * sub_stride;
  const int strips = W / 64;
  const int units_per_img = (H / kCwRows) * strips;
  const int units = B * units_per_img;
  const int wm = warp & 3, wn = warp >> 2;  // 4 co-blocks of 16 x 2 ci-blocks of 32

  float acc[kTaps][4][4];
#pragma unroll
  for (int t = 0; t < kTaps; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j][0] = acc[t][j][1] = acc[t][j][2] = acc[t][j][3] = 0.f;

  auto issue = [&](int unit, int stage) {
    const uint32_t sx = smem_u32(sm) + stage * Cfg::kStageBytes;
    const uint32_t sy = sx + Cfg::kXPix * 128;
    const int b = unit / units_per_img;
    const int u = unit % units_per_img;
    const int h0 = (u / strips) * kCwRows, w0 = (u % strips) * 64;
    for (int i = tid; i < Cfg::kXPix * 8; i += 256) {
      const int pix = i >> 3, ch = i & 7;
      const int r = pix / Cfg::kHaloW, xx = pix % Cfg::kHaloW;
      const int hh = h0 - KH / 2 + r, ww = w0 - KW / 2 + xx;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const bf16* src = x + (((long)b * H + (ok ? hh : 0)) * W + (ok ? ww : 0)) * 64 + ch * 8;
      cp_async_16(sx + cw_off(pix, ch), src, ok);
    }
    for (int i = tid; i < Cfg::kYPix * 8; i += 256) {
      const int pix = i >> 3, ch = i & 7;
      const int r = pix >> 6, ww = w0 + (pix & 63);
      const bf16* src = dyg + (long)b * dy_img + (long)(h0 + r) * dy_row + (long)ww * dy_pix + ch * 8;
      cp_async_16(sy + cw_off(pix, ch), src, true);
    }
  };

  int stage = 0;
  int unit = blockIdx.x;
  if (unit < units) issue(unit, 0);
  cp_async_commit();
  for (; unit < units; unit += gridDim.x) {
    const int nxt = unit + gridDim.x;
    if (nxt < units) issue(nxt, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t sx = smem_u32(sm) + stage * Cfg::kStageBytes;
    const uint32_t sy = sx + Cfg::kXPix * 128;
#pragma unroll 1
    for (int ks = 0; ks < Cfg::kYPix / 16; ++ks) {  // 16 output pixels per step, never crossing an image row
      const int r = ks >> 2, x0 = (ks & 3) * 16;
      uint32_t a[4];
      {
        // A[m=co][k=pixel]: lanes 0-7: pix 0..7 @ co chunk 2*wm; 8-15: pix 0..7 @ chunk 2*wm+1;
        //                   16-23: pix 8..15 @ chunk 2*wm; 24-31: pix 8..15 @ chunk 2*wm+1
        const int pix = r * 64 + x0 + (lane & 7) + ((lane >> 4) & 1) * 8;
        ldmatrix_x4_trans(a, sy + cw_off(pix, 2 * wm + ((lane >> 3) & 1)));
      }
#pragma unroll
      for (int tap = 0; tap < kTaps; ++tap) {
        const int ky = tap / KW, kx = tap % KW;
        // X pixel for output pixel (r, x0+j) and tap (ky,kx): halo (r + ky, x0 + j + kx)
        const int hp = (r + ky) * Cfg::kHaloW + x0 + kx + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          // lanes 0-7: pix 0..7 @ ci chunk n; 8-15: pix 8..15 @ n; 16-23: pix 0..7 @ n+1; 24-31: pix 8..15 @ n+1
          uint32_t rr[4];
          ldmatrix_x4_trans(rr, sx + cw_off(hp, wn * 4 + np * 2 + ((lane >> 4) & 1)));
          const uint32_t b0[2] = {rr[0], rr[1]}, b1[2] = {rr[2], rr[3]};
          mma_bf16_16816(acc[tap][2 * np], a, b0);
          mma_bf16_16816(acc[tap][2 * np + 1], a, b1);
        }
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  cp_async_wait<0>();
  // partial[cta][grp][tap][co 64][ci 64]
  const int g = lane >> 2, c = lane & 3;
  float* out = partial + ((long)blockIdx.x * gridDim.y + grp) * (kTaps * 64 * 64);
#pragma unroll
  for (int tap = 0; tap < kTaps; ++tap) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* r0 = out + ((long)tap * 64 + wm * 16 + g) * 64 + wn * 32 + nt * 8 + 2 * c;
      *reinterpret_cast<float2*>(r0) = make_float2(acc[tap][nt][0], acc[tap][nt][1]);
      *reinterpret_cast<float2*>(r0 + 8 * 64) = make_float2(acc[tap][nt][2], acc[tap][nt][3]);
    }
  }
}

// dW[co_full][ci][tap] = sum_cta partial[cta][grp][tap][co][ci];  co_full = co*co_mul + (grp_base+grp)*grp_mul
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ partial, int P, int groups, int taps, int co_mul,
                                         int grp_mul, int grp_base, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = groups * taps * 64 * 64;
  if (i >= n) return;
  double s = 0.0;
  int p = 0;
  for (; p + 8 <= P; p += 8) {  // eight partials in flight: a dependent L2 round trip per partial made this 19 us
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (long)(p + u) * n + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; p < P; ++p) s += __ldcg(partial + (long)p * n + i);
  const int ci = i & 63, co = (i >> 6) & 63, tap = (i >> 12) % taps, grp = i / (taps * 4096);
  const int co_full = co * co_mul + (grp_base + grp) * grp_mul;
  dw[((long)co_full * 64 + ci) * taps + tap] = (float)s;
}

int wg_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

int linear_wgrad_splits(long T, int N, int K) {
  const int nblk = ((N % 128 == 0) ? N / 128 : N / 64) * (K / 128);
  long chunks = (T + kLwChunk - 1) / kLwChunk;
  long s = wg_sms() / nblk;
  if (s > chunks) s = chunks;
  if (s < 1) s = 1;
  return (int)s;
}
size_t linear_wgrad_partial_bytes(long T, int N, int K) { return (size_t)linear_wgrad_splits(T, N, K) * N * K * 4; }

// out[r*ld_out + c] = scale * sum_p partial[p*stride + r*128 + c], c < 128
__global__ void reduce_partials_2d_kernel(const float* __restrict__ partial, int P, long stride, int rows,
                                          float* __restrict__ out, long ld_out, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 128) return;
  partial += (long)blockIdx.y * P * stride;  // K block
  out += (long)blockIdx.y * 128;
  double s = 0.0;
  int p = 0;
  for (; p + 8 <= P; p += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (long)(p + u) * stride + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; p < P; ++p) s += __ldcg(partial + (long)p * stride + i);
  out[(long)(i >> 7) * ld_out + (i & 127)] = (float)s * scale;
}

// dW fp32 [N][K] (row stride K) = scale * dY^T X ; K a multiple of 128, processed 128 columns at a time
int linear_wgrad(const bf16* dy, long ld_dy, const bf16* x, long ld_x, long T, int N, int K, float* dw, float scale,
                 float* partial, cudaStream_t s) {
  ProfScope _ps("linear_wgrad", s);
  FOCR_REQUIRE(N % 64 == 0 && K % 128 == 0, "linear_wgrad: N=%d K=%d", N, K);
  const int splits = linear_wgrad_splits(T, N, K);
  static bool init = false;
  if (!init) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(linear_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         LwCfg<128>::kSmemBytes));
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(linear_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         LwCfg<64>::kSmemBytes));
    init = true;
  }
  // one launch covers every (row block, token split, 128-column block of X); one reduction finishes all of dW
  const int kblocks = K / 128;
  if (N % 128 == 0)
    linear_wgrad_kernel<128><<<dim3(N / 128, splits, kblocks), 256, LwCfg<128>::kSmemBytes, s>>>(dy, ld_dy, x, ld_x, T, N,
                                                                                                partial);
  else
    linear_wgrad_kernel<64><<<dim3(N / 64, splits, kblocks), 256, LwCfg<64>::kSmemBytes, s>>>(dy, ld_dy, x, ld_x, T, N,
                                                                                              partial);
  FOCR_LAUNCH_CHECK();
  reduce_partials_2d_kernel<<<dim3(focr_cdiv(N * 128, 256), kblocks), 256, 0, s>>>(partial, splits, (long)N * 128, N, dw, K,
                                                                                  scale);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int conv_wgrad_ctas(int B, int H, int W, int groups) {
  int units = B * (H / kCwRows) * (W / 64);
  int g = wg_sms() / groups;
  if (g < 1) g = 1;
  return units < g ? units : g;
}
size_t conv3x3_wgrad_partial_bytes(int B, int H, int groups) {
  // the PixelShuffle variant (groups == 4) runs as two launches of 2 groups: size for the larger need
  const size_t a = (size_t)conv_wgrad_ctas(B, H, 64, groups) * groups * 9 * 64 * 64 * 4;
  const size_t b = groups == 4 ? (size_t)conv_wgrad_ctas(B, H, 64, 2) * 4 * 9 * 64 * 64 * 4 : 0;
  return a > b ? a : b;
}
size_t conv9x1_wgrad_partial_bytes(int B, int H, int W) {
  return (size_t)conv_wgrad_ctas(B, H, W, 1) * 9 * 64 * 64 * 4;
}

template <int KH, int KW>
static int conv_wgrad_set_attr() {
  static bool init = false;
  if (!init) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<KH, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CwCfg<KH, KW>::kSmemBytes));
    init = true;
  }
  return FOCR_OK;
}

// dW fp32 torch layout [Co][64][3][3]; x (B,H,64,64) NHWC; Co = 64*groups.
// shuf = 0: dy (B,H,64,Co) NHWC.  shuf = 1 (PixelShuffle variant, Co = 256): dy (B,2H,2W,64).
int conv3x3_wgrad(const bf16* dy, const bf16* x, int B, int H, int Co, int shuf, float* dw, float* partial,
                  cudaStream_t s) {
  ProfScope _ps("conv3x3_wgrad", s);
  FOCR_REQUIRE(Co % 64 == 0 && H % kCwRows == 0, "conv3x3_wgrad: Co=%d H=%d", Co, H);
  const int groups = Co / 64;
  FOCR_REQUIRE(!shuf || groups == 4, "conv3x3_wgrad: shuffle variant needs Co=256");
  if (conv3x3_wgrad_tc_supported(H, 64)) {  // tcgen05 path (wgrad_tc.cu), one launch per 64-channel output group
    if (!shuf) {
      for (int g = 0; g < groups; ++g) {
        int rc2 = conv3x3_wgrad_tc(dy + g * 64, Co, (long)64 * Co, (long)H * 64 * Co, x, B, H, 1, g * 64, dw, partial, s);
        if (rc2) return rc2;
      }
      return FOCR_OK;
    }
    // PixelShuffle: output channel co * 4 + (2 i + j) lives at HR pixel (2h + i, 2w + j) of the (B, 2H, 128, 64) gradient map
    for (int sub = 0; sub < 4; ++sub) {
      const int i = sub >> 1, j = sub & 1;
      int rc2 = conv3x3_wgrad_tc(dy + (long)i * 128 * 64 + j * 64, 128, (long)2 * 128 * 64, (long)2 * H * 128 * 64, x, B, H, 4, sub, dw,
                                 partial, s);
      if (rc2) return rc2;
    }
    return FOCR_OK;
  }
  int rc = conv_wgrad_set_attr<3, 3>();
  if (rc) return rc;
  constexpr int smem = CwCfg<3, 3>::kSmemBytes;
  const int ctas = conv_wgrad_ctas(B, H, 64, shuf ? 2 : groups);
  if (!shuf) {
    conv_wgrad_kernel<3, 3><<<dim3(ctas, groups), 256, smem, s>>>(dy, Co, (long)64 * Co, (long)H * 64 * Co, x, B, H, 64,
                                                                  64, partial);
    FOCR_LAUNCH_CHECK();
    const int n = groups * 9 * 4096;
    conv_wgrad_reduce_kernel<<<focr_cdiv(n, 256), 256, 0, s>>>(partial, ctas, groups, 9, 1, 64, 0, dw);
    FOCR_LAUNCH_CHECK();
    return FOCR_OK;
  }
  // PixelShuffle: output channel co*4 + sub, sub = 2*i + j lives at HR pixel (2h+i, 2w+j) (tbsrn.py:266).
  // One launch per sub-pixel row i; the two groups j = 0,1 are 64 elements apart.
  for (int i = 0; i < 2; ++i) {
    float* part_i = partial + (long)i * ctas * 2 * 9 * 4096;
    conv_wgrad_kernel<3, 3><<<dim3(ctas, 2), 256, smem, s>>>(dy + (long)i * 2 * 64 * 64, 2 * 64, (long)2 * 2 * 64 * 64,
                                                             (long)2 * H * 2 * 64 * 64, x, B, H, 64, 64, part_i);
    FOCR_LAUNCH_CHECK();
    conv_wgrad_reduce_kernel<<<focr_cdiv(2 * 9 * 4096, 256), 256, 0, s>>>(part_i, ctas, 2, 9, 4, 1, 2 * i, dw);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

// 9x1 (vertical taps) wgrad between two 64-channel NHWC maps of width W in {64,128}:
//   out[co][ci][dy] (fp32, [64][64][9]) = sum_p dY[p][co] * X[p + (dy-4, 0)][ci]
// Used for the dx-unrolled forms of the 9x9 convolutions (conv9x9.cu).
int conv9x1_wgrad(const bf16* dy, const bf16* x, int B, int H, int W, float* out, float* partial, cudaStream_t s) {
  ProfScope _ps("conv9x1_wgrad", s);
  FOCR_REQUIRE((W == 64 || W == 128) && H % kCwRows == 0, "conv9x1_wgrad: H=%d W=%d", H, W);
  int rc = conv_wgrad_set_attr<9, 1>();
  if (rc) return rc;
  const int ctas = conv_wgrad_ctas(B, H, W, 1);
  conv_wgrad_kernel<9, 1><<<dim3(ctas, 1), 256, CwCfg<9, 1>::kSmemBytes, s>>>(dy, 64, (long)W * 64, (long)H * W * 64, x,
                                                                             B, H, W, 64, partial);
  FOCR_LAUNCH_CHECK();
  conv_wgrad_reduce_kernel<<<focr_cdiv(9 * 4096, 256), 256, 0, s>>>(partial, ctas, 1, 9, 1, 64, 0, out);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
