// Bidirectional GRU (hidden 32 per direction) of TSRN's sequence residual blocks
// (scene-text-telescope/model/tsrn.py:128-145 GruBlock: conv1x1 -> nn.GRU(64, 32, bidirectional, batch_first);
// gru1 runs down the columns (T = 16, B*64 sequences), gru2 along the rows (T = 64, B*16 sequences), tsrn.py:96-98).
//
// The input projection W_ih x + b_ih for all time steps is one tcgen05 GEMM (N = 192 = 2 directions x 3 gates);
// these kernels do the latency-bound recurrence: ONE WARP per (sequence, direction), lane j owns hidden unit j, the
// 96x32 recurrent matrix lives in shared memory (row reads are conflict-free, h is broadcast with warp shuffles),
// the state stays in registers in fp32, and the next step's projected input is prefetched while the current step
// computes.  No per-timestep launches.
//   gates (nn.GRU order r, z, n):  r = s(xr + W_hr h + b_hr), z = s(xz + W_hz h + b_hz),
//                                   n = tanh(xn + r * (W_hn h + b_hn)),  h' = (1 - z) n + z h
// Backward is BPTT with gate recomputation from the saved fp32 h_{t-1}; it emits d(xproj) and d(hidden-side
// pre-activations) per token so that every weight gradient is a token-axis reduction GEMM (wgrad.cu).
#include "kernels.cuh"

namespace {

constexpr int kH = 32;
constexpr int kWarps = 8;  // per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

struct SeqGeom {
  long base;
  long tstride;
};
// vertical: sequence (b, w) walks h: token = (b*16 + h)*64 + w ; horizontal: sequence (b, h) walks w
__device__ __forceinline__ SeqGeom seq_geom(long s, int vertical) {
  SeqGeom g;
  if (vertical) {
    const long b = s >> 6, w = s & 63;
    g.base = b * 1024 + w;
    g.tstride = 64;
  } else {
    g.base = s * 64;
    g.tstride = 1;
  }
  return g;
}

template <int T>
__global__ void __launch_bounds__(kWarps * 32)
gru_fwd_kernel(const bf16* __restrict__ xp, const float* __restrict__ whh_f, const float* __restrict__ whh_r,
               const float* __restrict__ bhh_f, const float* __restrict__ bhh_r, bf16* __restrict__ out,
               float* __restrict__ hprev32, bf16* __restrict__ hprev16, long nseq, int vertical) {
  __shared__ float sw[2][96][kH + 1];
  for (int i = threadIdx.x; i < 2 * 96 * kH; i += blockDim.x) {
    const int d = i / (96 * kH), rem = i % (96 * kH);
    sw[d][rem / kH][rem % kH] = (d ? whh_r : whh_f)[rem];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long gw = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (gw >= nseq * 2) return;
  const int dir = (int)(gw & 1);
  const SeqGeom g = seq_geom(gw >> 1, vertical);
  const float* bh = dir ? bhh_r : bhh_f;
  const float b_r = bh[lane], b_z = bh[32 + lane], b_n = bh[64 + lane];
  float h = 0.f;
  long tok = g.base + (dir ? (T - 1) * g.tstride : 0);
  const long dtok = dir ? -g.tstride : g.tstride;
  const bf16* xrow = xp + tok * 192 + dir * 96 + lane;
  float xr = __bfloat162float(xrow[0]), xz = __bfloat162float(xrow[32]), xn = __bfloat162float(xrow[64]);
#pragma unroll 1
  for (int step = 0; step < T; ++step) {
    float nxr = 0.f, nxz = 0.f, nxn = 0.f;
    if (step + 1 < T) {  // prefetch the next step's projected input (independent of h)
      const bf16* nrow = xp + (tok + dtok) * 192 + dir * 96 + lane;
      nxr = __bfloat162float(nrow[0]);
      nxz = __bfloat162float(nrow[32]);
      nxn = __bfloat162float(nrow[64]);
    }
    float ar = b_r, az = b_z, an = b_n;
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const float hk = __shfl_sync(0xffffffffu, h, k);
      ar = fmaf(sw[dir][lane][k], hk, ar);
      az = fmaf(sw[dir][32 + lane][k], hk, az);
      an = fmaf(sw[dir][64 + lane][k], hk, an);
    }
    const float r = sigmoidf_(xr + ar), z = sigmoidf_(xz + az), n = tanhf(xn + r * an);
    hprev32[tok * 64 + dir * 32 + lane] = h;
    hprev16[tok * 64 + dir * 32 + lane] = __float2bfloat16_rn(h);
    h = (1.f - z) * n + z * h;
    out[tok * 64 + dir * 32 + lane] = __float2bfloat16_rn(h);
    tok += dtok;
    xr = nxr;
    xz = nxz;
    xn = nxn;
  }
}

template <int T>
__global__ void __launch_bounds__(kWarps * 32)
gru_bwd_kernel(const bf16* __restrict__ xp, const float* __restrict__ whh_f, const float* __restrict__ whh_r,
               const float* __restrict__ bhh_f, const float* __restrict__ bhh_r, const float* __restrict__ hprev32,
               const bf16* __restrict__ dout, bf16* __restrict__ dxp, bf16* __restrict__ dhid, long nseq,
               int vertical) {
  __shared__ float sw[2][96][kH + 1];
  for (int i = threadIdx.x; i < 2 * 96 * kH; i += blockDim.x) {
    const int d = i / (96 * kH), rem = i % (96 * kH);
    sw[d][rem / kH][rem % kH] = (d ? whh_r : whh_f)[rem];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long gw = (long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (gw >= nseq * 2) return;
  const int dir = (int)(gw & 1);
  const SeqGeom g = seq_geom(gw >> 1, vertical);
  const float* bh = dir ? bhh_r : bhh_f;
  const float b_r = bh[lane], b_z = bh[32 + lane], b_n = bh[64 + lane];
  // walk the forward order backwards: forward direction started at t = 0, reverse direction at t = T-1
  long tok = g.base + (dir ? 0 : (T - 1) * g.tstride);
  const long dtok = dir ? g.tstride : -g.tstride;
  float dh = 0.f;
#pragma unroll 1
  for (int step = 0; step < T; ++step) {
    const long o = tok * 64 + dir * 32 + lane;
    const float hp = hprev32[o];
    const float dht = __bfloat162float(dout[o]) + dh;
    const bf16* xrow = xp + tok * 192 + dir * 96 + lane;
    const float xr = __bfloat162float(xrow[0]), xz = __bfloat162float(xrow[32]), xn = __bfloat162float(xrow[64]);
    float ar = b_r, az = b_z, an = b_n;
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const float hk = __shfl_sync(0xffffffffu, hp, k);
      ar = fmaf(sw[dir][lane][k], hk, ar);
      az = fmaf(sw[dir][32 + lane][k], hk, az);
      an = fmaf(sw[dir][64 + lane][k], hk, an);
    }
    const float r = sigmoidf_(xr + ar), z = sigmoidf_(xz + az), n = tanhf(xn + r * an);
    const float dn = dht * (1.f - z);
    const float dz = dht * (hp - n);
    const float dan = dn * (1.f - n * n);
    const float dar = dan * an * r * (1.f - r);
    const float daz = dz * z * (1.f - z);
    const float dhn = dan * r;
    bf16* dx = dxp + tok * 192 + dir * 96 + lane;
    dx[0] = __float2bfloat16_rn(dar);
    dx[32] = __float2bfloat16_rn(daz);
    dx[64] = __float2bfloat16_rn(dan);
    bf16* dd = dhid + tok * 192 + dir * 96 + lane;
    dd[0] = __float2bfloat16_rn(dar);
    dd[32] = __float2bfloat16_rn(daz);
    dd[64] = __float2bfloat16_rn(dhn);
    // dh_{t-1}[j] = dht*z + sum_k W_hr[k][j] dar_k + W_hz[k][j] daz_k + W_hn[k][j] dhn_k   (column j: conflict-free)
    float acc = dht * z;
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      acc = fmaf(sw[dir][k][lane], __shfl_sync(0xffffffffu, dar, k), acc);
      acc = fmaf(sw[dir][32 + k][lane], __shfl_sync(0xffffffffu, daz, k), acc);
      acc = fmaf(sw[dir][64 + k][lane], __shfl_sync(0xffffffffu, dhn, k), acc);
    }
    dh = acc;
    tok += dtok;
  }
}

// [W_ih ; W_ih_reverse] (2 x [96][64]) -> bf16 [192][64] and its transpose [64][192]; bias cat -> fp32 [192]
__global__ void gru_prep_kernel(const float* __restrict__ wf, const float* __restrict__ wr, const float* __restrict__ bf_,
                                const float* __restrict__ br, bf16* __restrict__ w, bf16* __restrict__ wt,
                                float* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 192 * 64) {
    const int n = i / 64, k = i % 64;
    const bf16 v = __float2bfloat16_rn(n < 96 ? wf[n * 64 + k] : wr[(n - 96) * 64 + k]);
    w[i] = v;
    wt[k * 192 + n] = v;
  }
  if (i < 192) b[i] = i < 96 ? bf_[i] : br[i - 96];
}

// g: fp32 [192][64] = dhid^T hprev (both directions, cross blocks are meaningless):
// dW_hh_f[n][k] = g[n][k], dW_hh_r[n][k] = g[96+n][32+k]
__global__ void gru_unpack_whh_kernel(const float* __restrict__ g, float* __restrict__ df, float* __restrict__ dr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 96 * 32) return;
  const int n = i / 32, k = i % 32;
  if (df) df[i] = g[n * 64 + k];
  if (dr) dr[i] = g[(96 + n) * 64 + 32 + k];
}

// fold of the "two tokens per row" trick: g [2N][128] -> dw[n][k] = g[n][k] + g[N+n][64+k], k < 64
__global__ void fold_k64_kernel(const float* __restrict__ g, float* __restrict__ dw, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 64) return;
  const int n = i / 64, k = i % 64;
  dw[i] = g[(long)n * 128 + k] + g[(long)(N + n) * 128 + 64 + k];
}

}  // namespace

int gru_prep(const float* wih_f, const float* wih_r, const float* bih_f, const float* bih_r, bf16* w, bf16* wt, float* b,
             cudaStream_t s) {
  gru_prep_kernel<<<focr_cdiv(192 * 64, 256), 256, 0, s>>>(wih_f, wih_r, bih_f, bih_r, w, wt, b);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int gru_forward(const bf16* xp, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r, bf16* out,
                float* hprev32, bf16* hprev16, int B, int vertical, cudaStream_t s) {
  ProfScope _ps("gru_fwd", s);
  const long nseq = vertical ? (long)B * 64 : (long)B * 16;
  const int grid = focr_cdiv(nseq * 2, kWarps);
  if (vertical)
    gru_fwd_kernel<16><<<grid, kWarps * 32, 0, s>>>(xp, whh_f, whh_r, bhh_f, bhh_r, out, hprev32, hprev16, nseq, 1);
  else
    gru_fwd_kernel<64><<<grid, kWarps * 32, 0, s>>>(xp, whh_f, whh_r, bhh_f, bhh_r, out, hprev32, hprev16, nseq, 0);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int gru_backward(const bf16* xp, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r,
                 const float* hprev32, const bf16* dout, bf16* dxp, bf16* dhid, int B, int vertical, cudaStream_t s) {
  ProfScope _ps("gru_bwd", s);
  const long nseq = vertical ? (long)B * 64 : (long)B * 16;
  const int grid = focr_cdiv(nseq * 2, kWarps);
  if (vertical)
    gru_bwd_kernel<16><<<grid, kWarps * 32, 0, s>>>(xp, whh_f, whh_r, bhh_f, bhh_r, hprev32, dout, dxp, dhid, nseq, 1);
  else
    gru_bwd_kernel<64><<<grid, kWarps * 32, 0, s>>>(xp, whh_f, whh_r, bhh_f, bhh_r, hprev32, dout, dxp, dhid, nseq, 0);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int gru_unpack_whh(const float* g, float* df, float* dr, cudaStream_t s) {
  gru_unpack_whh_kernel<<<focr_cdiv(96 * 32, 256), 256, 0, s>>>(g, df, dr);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// dW[N][64] (fp32) = dY[T,N]^T X[T,64]: the (T,64) operand is viewed as (T/2,128) and dY as (T/2,2N) so the
// 128-column wgrad kernel applies; the two diagonal blocks of the [2N][128] result are summed.  tmp: >= 2N*128 floats.
int linear_wgrad_k64(const bf16* dy, const bf16* x, long T, int N, float* dw, float* tmp, float* partial, cudaStream_t s) {
  FOCR_REQUIRE(T % 2 == 0 && (2 * N) % 64 == 0, "linear_wgrad_k64: T=%ld N=%d", T, N);
  int rc = linear_wgrad(dy, 2L * N, x, 128, T / 2, 2 * N, 128, tmp, 1.f, partial, s);
  if (rc) return rc;
  fold_k64_kernel<<<focr_cdiv(N * 64, 256), 256, 0, s>>>(tmp, dw, N);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
