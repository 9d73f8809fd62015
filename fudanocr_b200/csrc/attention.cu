// Fused self-attention of the TBSRN FeatureEnhancer: h=4 heads, d_k=32, 1024 tokens
// (scene-text-telescope/model/tbsrn.py:109-150: softmax(QK^T/sqrt(d_k)) -> dropout(0.1) -> PV).
// The reference materialises P = (B,4,1024,1024) fp32 (4 GiB at B=256); here P never leaves
// registers.  Input is the packed projection QKV (T,384) bf16 = [q | k | v], head h at columns
// h*32 of each third; output O (T,128) bf16 in the "concat heads" layout the out-projection reads.
//
// d_k = 32 makes the op exp/issue-bound rather than MMA-bound (128 tensor FLOPs per exp), so the
// kernels are warp-level mma.sync m16n8k16 flash kernels: one CTA per (batch, head) keeps the whole
// K and V (or Q and dO) of that head in shared memory (2 x 64 KB, XOR-swizzled for ldmatrix) so they
// are read from HBM exactly once.
//   forward : warp owns 16 query rows, loops over 16 KV tiles of 64, online softmax
//   backward: two passes without atomics -
//     pass A (dQ)   : warp owns 16 query rows;  dS = P o (dP - D),  dQ = dS K
//     pass B (dK,dV): warp owns 16 key rows;    works on S^T, dP^T; dV = Pd^T dO, dK = dS^T Q
// Dropout uses the counter hash of common.cuh keyed on (b,h,q,k).  The hash is ~60 % of the forward's
// instruction stream, so when the caller provides a keep-bit buffer (1 bit per (b,h,q,k), B x 512 KB) the
// forward stores its keep decisions in its own fragment layout - word [bh][q/16][k/64][lane], bit
// 16*(q%16 >= 8) + 2*((k%64)/8) + (k&1), lane = (q%8)*4 + (k%8)/2 - and both backward passes test bits
// instead of re-hashing (DROP = 2).  Without the buffer the backward regenerates the mask (DROP = 1).
// The 1/(1-p) factor is folded into the output scales: P.V, dV accumulate kept probabilities unscaled,
// dS = P o (keep o dP - D (1-p)) / (1-p).
#include "kernels.cuh"

namespace {

constexpr int kS = 1024;      // tokens
constexpr int kLdQkv = 384;   // row stride of the packed projection
constexpr int kLdO = 128;
constexpr float kScale = 0.17677669529663687f;  // 1/sqrt(32)
constexpr float kScaleLog2 = kScale * 1.4426950408889634f;
constexpr int kTileBytes = kS * 64;  // one [1024][32] bf16 head slice

__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  return (uint32_t)(row * 64 + (((chunk ^ (row >> 1)) & 3) << 4));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ld32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

__device__ __forceinline__ void load_head_tile(uint32_t sbase, const bf16* g, long ld, int tid) {
  for (int i = tid; i < kS * 4; i += (int)blockDim.x) {
    const int row = i >> 2, ch = i & 3;
    cp_async_16(sbase + sw_off(row, ch), g + (long)row * ld + ch * 8, true);
  }
}
// A-operand fragments (16 rows x 32 cols, two k-steps) of a row-major bf16 matrix in global memory
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], const bf16* row0, long ld, int c) {
  const bf16* row1 = row0 + 8 * ld;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    a[ks][0] = ld32(row0 + ks * 16 + 2 * c);
    a[ks][1] = ld32(row1 + ks * 16 + 2 * c);
    a[ks][2] = ld32(row0 + ks * 16 + 8 + 2 * c);
    a[ks][3] = ld32(row1 + ks * 16 + 8 + 2 * c);
  }
}
// acc[8][4] (16 x 64) = A(16x32) * M^T, M = [64 rows of the smem head tile starting at row0][32]
__device__ __forceinline__ void mma_a_mt(float (&acc)[8][4], const uint32_t (&a)[2][4], uint32_t sbase, int row0,
                                         int lane) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    uint32_t r[4];
    ldmatrix_x4(r, sbase + sw_off(row0 + n * 8 + (lane & 7), lane >> 3));
    const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
    mma_bf16_16816(acc[n], a[0], b0);
    mma_bf16_16816(acc[n], a[1], b1);
  }
}
// acc[4][4] (16 x 32) += P(16x64, packed A frags) * M, M = [64 rows starting at row0][32] of the smem tile
__device__ __forceinline__ void mma_p_m(float (&acc)[4][4], const uint32_t (&pa)[4][4], uint32_t sbase, int row0,
                                        int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int row = row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int nd = 0; nd < 4; nd += 2) {
      uint32_t r[4];
      ldmatrix_x4_trans(r, sbase + sw_off(row, nd + (lane >> 4)));
      const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
      mma_bf16_16816(acc[nd], pa[kk], b0);
      mma_bf16_16816(acc[nd + 1], pa[kk], b1);
    }
  }
}
__device__ __forceinline__ void pack_frags(uint32_t (&pa)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    pa[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// ------------------------------------------------------------------------------------------
template <int DROP, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse2, uint32_t key,
                uint32_t thresh16, float inv_keep, uint32_t* __restrict__ drop_bits) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sK = smem_u32(sm), sV = sK + kTileBytes;
  const int bh = blockIdx.x, b = bh >> 2, h = bh & 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const bf16* base = qkv + (long)b * kS * kLdQkv + h * 32;
  load_head_tile(sK, base + 128, kLdQkv, tid);
  load_head_tile(sV, base + 256, kLdQkv, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  for (int u = warp; u < kS / 16; u += NW) {  // 64 units of 16 query rows, round-robin over the warps
    const int q0 = u * 16;
    uint32_t qa[2][4];
    load_a_frags(qa, base + (long)(q0 + g) * kLdQkv, kLdQkv, c);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    const uint32_t rb0 = (uint32_t)(bh * kS + q0 + g) * 512u, rb1 = rb0 + 8u * 512u;

#pragma unroll 1
    for (int kt = 0; kt < kS / 64; ++kt) {
      float s[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      mma_a_mt(s, qa, sK, kt * 64, lane);
      float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      mx0 = quad_max(mx0);
      mx1 = quad_max(mx1);
      const float mn0 = fmaxf(m0, mx0 * kScaleLog2), mn1 = fmaxf(m1, mx1 * kScaleLog2);
      const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);
      m0 = mn0;
      m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        s[n][0] = ex2(fmaf(s[n][0], kScaleLog2, -mn0));
        s[n][1] = ex2(fmaf(s[n][1], kScaleLog2, -mn0));
        s[n][2] = ex2(fmaf(s[n][2], kScaleLog2, -mn1));
        s[n][3] = ex2(fmaf(s[n][3], kScaleLog2, -mn1));
        rs0 += s[n][0] + s[n][1];
        rs1 += s[n][2] + s[n][3];
      }
      l0 = fmaf(l0, al0, rs0);
      l1 = fmaf(l1, al1, rs1);
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        o[nd][0] *= al0;
        o[nd][1] *= al0;
        o[nd][2] *= al1;
        o[nd][3] *= al1;
      }
      if (DROP) {
        const uint32_t cb = (uint32_t)(kt * 32 + c);
        uint32_t bits = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const uint32_t h0 = drop_hash32(key, rb0 + cb + n * 4), h1 = drop_hash32(key, rb1 + cb + n * 4);
          const bool k0 = (h0 & 0xFFFFu) >= thresh16, k1 = (h0 >> 16) >= thresh16;
          const bool k2 = (h1 & 0xFFFFu) >= thresh16, k3 = (h1 >> 16) >= thresh16;
          s[n][0] = k0 ? s[n][0] : 0.f;
          s[n][1] = k1 ? s[n][1] : 0.f;
          s[n][2] = k2 ? s[n][2] : 0.f;
          s[n][3] = k3 ? s[n][3] : 0.f;
          if (DROP == 2) {
            if (k0) bits |= 1u << (2 * n);
            if (k1) bits |= 2u << (2 * n);
            if (k2) bits |= 0x10000u << (2 * n);
            if (k3) bits |= 0x20000u << (2 * n);
          }
        }
        if (DROP == 2) drop_bits[(((size_t)bh * 64 + u) * 16 + kt) * 32 + lane] = bits;
      }
      uint32_t pa[4][4];
      pack_frags(pa, s);
      mma_p_m(o, pa, sV, kt * 64, lane);
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float i0 = inv_keep / l0, i1 = inv_keep / l1;
    bf16* orow0 = out + ((long)b * kS + q0 + g) * kLdO + h * 32;
    bf16* orow1 = orow0 + 8 * kLdO;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      *reinterpret_cast<uint32_t*>(orow0 + nd * 8 + 2 * c) = pack_bf16x2(o[nd][0] * i0, o[nd][1] * i0);
      *reinterpret_cast<uint32_t*>(orow1 + nd * 8 + 2 * c) = pack_bf16x2(o[nd][2] * i1, o[nd][3] * i1);
    }
    if (c == 0) {
      lse2[(long)bh * kS + q0 + g] = m0 + log2f(l0);
      lse2[(long)bh * kS + q0 + g + 8] = m1 + log2f(l1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward kernels.  Both are latency-bound at 2 warps per scheduler (the S -> exp -> dS -> MMA chain is serial
// inside a warp), so they walk the other sequence dimension in 32-wide steps, software-pipelined by hand: the
// QK^T / dO V^T MMAs of step i+1 are issued before the softmax arithmetic of step i (two register buffers A/B).
template <int NT>
__device__ __forceinline__ void mma_a_mt_t(float (&acc)[NT][4], const uint32_t (&a)[2][4], uint32_t sbase, int row0,
                                           int lane) {
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    uint32_t r[4];
    ldmatrix_x4(r, sbase + sw_off(row0 + n * 8 + (lane & 7), lane >> 3));
    const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
    mma_bf16_16816(acc[n], a[0], b0);
    mma_bf16_16816(acc[n], a[1], b1);
  }
}
template <int KK>
__device__ __forceinline__ void mma_p_m_t(float (&acc)[4][4], const uint32_t (&pa)[KK][4], uint32_t sbase, int row0,
                                          int lane) {
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    const int row = row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int nd = 0; nd < 4; nd += 2) {
      uint32_t r[4];
      ldmatrix_x4_trans(r, sbase + sw_off(row, nd + (lane >> 4)));
      const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
      mma_bf16_16816(acc[nd], pa[kk], b0);
      mma_bf16_16816(acc[nd + 1], pa[kk], b1);
    }
  }
}
__device__ __forceinline__ void pack_frags2(uint32_t (&pa)[2][4], const float (&s)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    pa[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ void zero44(float (&x)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i][0] = x[i][1] = x[i][2] = x[i][3] = 0.f;
}

// backward pass A: dQ (and D = rowsum(dO o O), written for pass B)
template <int DROP, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o_in, const bf16* __restrict__ d_o,
                   const float* __restrict__ lse2, float* __restrict__ dsum, bf16* __restrict__ dqkv, uint32_t key,
                   uint32_t thresh16, float inv_keep, const uint32_t* __restrict__ drop_bits) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sK = smem_u32(sm), sV = sK + kTileBytes;
  const int bh = blockIdx.x, b = bh >> 2, h = bh & 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const bf16* base = qkv + (long)b * kS * kLdQkv + h * 32;
  load_head_tile(sK, base + 128, kLdQkv, tid);
  load_head_tile(sV, base + 256, kLdQkv, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  for (int u = warp; u < kS / 16; u += NW) {
    const int q0 = u * 16;
    const long t0 = (long)b * kS + q0 + g;
    uint32_t qa[2][4], da[2][4], oa[2][4];
    load_a_frags(qa, base + (long)(q0 + g) * kLdQkv, kLdQkv, c);
    load_a_frags(da, d_o + t0 * kLdO + h * 32, kLdO, c);
    load_a_frags(oa, o_in + t0 * kLdO + h * 32, kLdO, c);
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(da[ks][j]), y = unpack_bf16x2(oa[ks][j]);
        const float v = x.x * y.x + x.y * y.y;
        if (j & 1) D1 += v; else D0 += v;
      }
    }
    D0 = quad_sum(D0);
    D1 = quad_sum(D1);
    if (c == 0) {
      dsum[(long)bh * kS + q0 + g] = D0;
      dsum[(long)bh * kS + q0 + g + 8] = D1;
    }
    D0 *= 1.f / inv_keep;  // dS = P o (keep o dP - D (1-p)) / (1-p); the 1/(1-p) goes into the final scale
    D1 *= 1.f / inv_keep;
    const float L0 = lse2[(long)bh * kS + q0 + g], L1 = lse2[(long)bh * kS + q0 + g + 8];
    float dq[4][4];
    zero44(dq);
    const uint32_t rb0 = (uint32_t)(bh * kS + q0 + g) * 512u, rb1 = rb0 + 8u * 512u;
    const uint32_t* wb = drop_bits + ((size_t)bh * 64 + u) * 16 * 32 + lane;

    auto compute = [&](float (&s)[4][4], float (&dp)[4][4], int row0) {
      zero44(s);
      zero44(dp);
      mma_a_mt_t<4>(s, qa, sK, row0, lane);
      mma_a_mt_t<4>(dp, da, sV, row0, lane);
    };
    // nb: 0 for the even 32-step of a 64-wide bit word, 4 for the odd one
    auto process = [&](float (&s)[4][4], float (&dp)[4][4], int row0, uint32_t w, const int nb) {
      const uint32_t cb = (uint32_t)(row0 >> 1) + c;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float p0 = ex2(fmaf(s[n][0], kScaleLog2, -L0)), p1 = ex2(fmaf(s[n][1], kScaleLog2, -L0));
        const float p2 = ex2(fmaf(s[n][2], kScaleLog2, -L1)), p3 = ex2(fmaf(s[n][3], kScaleLog2, -L1));
        float e0 = dp[n][0], e1 = dp[n][1], e2 = dp[n][2], e3 = dp[n][3];
        if (DROP == 1) {
          const uint32_t h0 = drop_hash32(key, rb0 + cb + n * 4), h1 = drop_hash32(key, rb1 + cb + n * 4);
          e0 = (h0 & 0xFFFFu) >= thresh16 ? e0 : 0.f;
          e1 = (h0 >> 16) >= thresh16 ? e1 : 0.f;
          e2 = (h1 & 0xFFFFu) >= thresh16 ? e2 : 0.f;
          e3 = (h1 >> 16) >= thresh16 ? e3 : 0.f;
        } else if (DROP == 2) {
          e0 = (w & (1u << (2 * (n + nb)))) ? e0 : 0.f;
          e1 = (w & (2u << (2 * (n + nb)))) ? e1 : 0.f;
          e2 = (w & (0x10000u << (2 * (n + nb)))) ? e2 : 0.f;
          e3 = (w & (0x20000u << (2 * (n + nb)))) ? e3 : 0.f;
        }
        s[n][0] = p0 * (e0 - D0);
        s[n][1] = p1 * (e1 - D0);
        s[n][2] = p2 * (e2 - D1);
        s[n][3] = p3 * (e3 - D1);
      }
      uint32_t pa[2][4];
      pack_frags2(pa, s);
      mma_p_m_t<2>(dq, pa, sK, row0, lane);
    };

    float sA[4][4], dA[4][4], sB[4][4], dB[4][4];
    uint32_t w = 0, w1 = 0, w2 = 0;  // bit words of tiles kt, kt+1, kt+2 (loads run two tiles ahead)
    if (DROP == 2) {
      w1 = __ldg(wb);
      w2 = __ldg(wb + 32);
    }
    compute(sA, dA, 0);
#pragma unroll 1
    for (int kt = 0; kt < kS / 64; ++kt) {
      if (DROP == 2) {
        w = w1;
        w1 = w2;
        if (kt + 2 < kS / 64) w2 = __ldg(wb + (kt + 2) * 32);
      }
      compute(sB, dB, kt * 64 + 32);
      process(sA, dA, kt * 64, w, 0);
      if (kt + 1 < kS / 64) compute(sA, dA, kt * 64 + 64);
      process(sB, dB, kt * 64 + 32, w, 4);
    }
    bf16* r0 = dqkv + t0 * kLdQkv + h * 32;
    bf16* r1 = r0 + 8 * kLdQkv;
    const float osc = kScale * inv_keep;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      *reinterpret_cast<uint32_t*>(r0 + nd * 8 + 2 * c) = pack_bf16x2(dq[nd][0] * osc, dq[nd][1] * osc);
      *reinterpret_cast<uint32_t*>(r1 + nd * 8 + 2 * c) = pack_bf16x2(dq[nd][2] * osc, dq[nd][3] * osc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward pass B: dK, dV (warp owns 16 key rows; everything is the transpose of pass A)
template <int DROP>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse2,
                    const float* __restrict__ dsum, bf16* __restrict__ dqkv, uint32_t key, uint32_t thresh16,
                    float inv_keep, const uint32_t* __restrict__ drop_bits) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sQ = smem_u32(sm), sdO = sQ + kTileBytes;
  float* sL = reinterpret_cast<float*>(sm + 2 * kTileBytes);
  float* sD = sL + kS;
  // DROP == 2: keep-bit words of the current 128-key block, all 64 query units x 2 key tiles x 32 lanes, double-buffered
  uint32_t* sW = reinterpret_cast<uint32_t*>(sD + kS);
  const int bh = blockIdx.x, b = bh >> 2, h = bh & 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const bf16* base = qkv + (long)b * kS * kLdQkv + h * 32;
  auto load_bits = [&](int kb) {
    if (DROP == 2) {
      const uint32_t dst = smem_u32(sW) + (kb & 1) * 16384;
      const uint32_t* src = drop_bits + ((size_t)bh * 64 * 16 + 2 * kb) * 32;
      for (int i = tid; i < 1024; i += 256)
        cp_async_16(dst + (i >> 4) * 256 + (i & 15) * 16, src + (size_t)(i >> 4) * 512 + (i & 15) * 4, true);
    }
  };
  load_head_tile(sQ, base, kLdQkv, tid);
  load_head_tile(sdO, d_o + (long)b * kS * kLdO + h * 32, kLdO, tid);
  load_bits(0);
  cp_async_commit();
  for (int i = tid; i < kS; i += (int)blockDim.x) {
    sL[i] = lse2[(long)bh * kS + i];
    sD[i] = dsum[(long)bh * kS + i] * (1.f / inv_keep);
  }
  cp_async_wait<0>();
  __syncthreads();

  for (int kb = 0; kb < kS / 128; ++kb) {
    if (DROP == 2) {
      cp_async_wait<0>();
      __syncthreads();  // this block's bit words landed, and every warp is done with the buffer refilled next
      if (kb + 1 < kS / 128) load_bits(kb + 1);
      cp_async_commit();
    }
    const int kv0 = kb * 128 + warp * 16;
    uint32_t ka[2][4], va[2][4];
    load_a_frags(ka, base + 128 + (long)(kv0 + g) * kLdQkv, kLdQkv, c);
    load_a_frags(va, base + 256 + (long)(kv0 + g) * kLdQkv, kLdQkv, c);
    float dk[4][4], dv[4][4];
    zero44(dk);
    zero44(dv);
    // element (kv, q): counter = (bh*1024 + q)*512 + kv/2, 16-bit lane = kv & 1
    const uint32_t kvh0 = (uint32_t)((kv0 + g) >> 1), kvh1 = (uint32_t)((kv0 + g + 8) >> 1);
    const int sh = ((kv0 + g) & 1) * 16;  // same parity for row g and g+8
    // keep-bit words of the forward's fragment layout (see the file header): q -> (unit, half, lane group), kv -> bit
    const uint32_t* wb = sW + (kb & 1) * 4096 + (warp >> 2) * 32 + (g >> 1);
    const int wl0 = (2 * c) * 4, wl1 = (2 * c + 1) * 4;
    const int shl = 2 * ((kv0 & 63) >> 3) + (g & 1);

    // step qs covers query rows [qs*32, qs*32+32): two forward units, one pair of bit words each
    auto compute = [&](float (&st)[4][4], float (&dpt)[4][4], uint32_t (&wq)[2][2], int qs) {
      if (DROP == 2) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          wq[m][0] = wb[(qs * 2 + m) * 64 + wl0];
          wq[m][1] = wb[(qs * 2 + m) * 64 + wl1];
        }
      }
      zero44(st);
      zero44(dpt);
      mma_a_mt_t<4>(st, ka, sQ, qs * 32, lane);
      mma_a_mt_t<4>(dpt, va, sdO, qs * 32, lane);
    };
    auto process = [&](float (&st)[4][4], float (&dpt)[4][4], const uint32_t (&wq)[2][2], int qs) {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int q = qs * 32 + n * 8 + 2 * c;
        const float2 Lq = *reinterpret_cast<const float2*>(sL + q);
        const float2 Dq = *reinterpret_cast<const float2*>(sD + q);
        const float p0 = ex2(fmaf(st[n][0], kScaleLog2, -Lq.x)), p1 = ex2(fmaf(st[n][1], kScaleLog2, -Lq.y));
        const float p2 = ex2(fmaf(st[n][2], kScaleLog2, -Lq.x)), p3 = ex2(fmaf(st[n][3], kScaleLog2, -Lq.y));
        float e0 = dpt[n][0], e1 = dpt[n][1], e2 = dpt[n][2], e3 = dpt[n][3];
        float d0 = p0, d1 = p1, d2 = p2, d3 = p3;
        if (DROP) {
          bool k0, k1, k2, k3;
          if (DROP == 1) {
            const uint32_t qb0 = (uint32_t)(bh * kS + q) * 512u, qb1 = qb0 + 512u;
            k0 = ((drop_hash32(key, qb0 + kvh0) >> sh) & 0xFFFFu) >= thresh16;
            k1 = ((drop_hash32(key, qb1 + kvh0) >> sh) & 0xFFFFu) >= thresh16;
            k2 = ((drop_hash32(key, qb0 + kvh1) >> sh) & 0xFFFFu) >= thresh16;
            k3 = ((drop_hash32(key, qb1 + kvh1) >> sh) & 0xFFFFu) >= thresh16;
          } else {
            const uint32_t w0 = wq[n >> 1][0] >> shl, w1 = wq[n >> 1][1] >> shl;
            const uint32_t b0 = (n & 1) ? 0x10000u : 1u, b2 = (n & 1) ? 0x40000u : 4u;
            k0 = w0 & b0;
            k1 = w1 & b0;
            k2 = w0 & b2;
            k3 = w1 & b2;
          }
          e0 = k0 ? e0 : 0.f; d0 = k0 ? p0 : 0.f;
          e1 = k1 ? e1 : 0.f; d1 = k1 ? p1 : 0.f;
          e2 = k2 ? e2 : 0.f; d2 = k2 ? p2 : 0.f;
          e3 = k3 ? e3 : 0.f; d3 = k3 ? p3 : 0.f;
        }
        st[n][0] = p0 * (e0 - Dq.x);
        st[n][1] = p1 * (e1 - Dq.y);
        st[n][2] = p2 * (e2 - Dq.x);
        st[n][3] = p3 * (e3 - Dq.y);
        dpt[n][0] = d0;
        dpt[n][1] = d1;
        dpt[n][2] = d2;
        dpt[n][3] = d3;
      }
      uint32_t pa[2][4];
      pack_frags2(pa, dpt);
      mma_p_m_t<2>(dv, pa, sdO, qs * 32, lane);
      pack_frags2(pa, st);
      mma_p_m_t<2>(dk, pa, sQ, qs * 32, lane);
    };

    float sA[4][4], dA[4][4], sB[4][4], dB[4][4];
    uint32_t wA[2][2], wB[2][2];
    compute(sA, dA, wA, 0);
#pragma unroll 1
    for (int qs = 0; qs < kS / 32; qs += 2) {
      compute(sB, dB, wB, qs + 1);
      process(sA, dA, wA, qs);
      if (qs + 2 < kS / 32) compute(sA, dA, wA, qs + 2);
      process(sB, dB, wB, qs + 1);
    }
    bf16* r0 = dqkv + ((long)b * kS + kv0 + g) * kLdQkv + h * 32;
    bf16* r1 = r0 + 8 * kLdQkv;
    const float ksc = kScale * inv_keep;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      *reinterpret_cast<uint32_t*>(r0 + 128 + nd * 8 + 2 * c) = pack_bf16x2(dk[nd][0] * ksc, dk[nd][1] * ksc);
      *reinterpret_cast<uint32_t*>(r1 + 128 + nd * 8 + 2 * c) = pack_bf16x2(dk[nd][2] * ksc, dk[nd][3] * ksc);
      *reinterpret_cast<uint32_t*>(r0 + 256 + nd * 8 + 2 * c) =
          pack_bf16x2(dv[nd][0] * inv_keep, dv[nd][1] * inv_keep);
      *reinterpret_cast<uint32_t*>(r1 + 256 + nd * 8 + 2 * c) =
          pack_bf16x2(dv[nd][2] * inv_keep, dv[nd][3] * inv_keep);
    }
  }
}

// warps per CTA (one CTA per SM: 128 KB of smem).  More resident warps hide the mma.sync / MUFU / ldmatrix
// latencies of this issue-bound kernel; the register file caps them (65536 / (32 * regs)).
constexpr int kFwdWarps = 16;  // <= 128 registers
constexpr int kDqWarps = 8;   // measured: 8 warps (172 regs, no spills) beats 13/16 warps at the 128-register cap

template <typename K>
int set_smem(K kernel, int bytes) {
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return FOCR_OK;
}

}  // namespace

// p_drop = thresh16 / 65536; thresh16 == 0 disables dropout (eval / parity runs).  drop_bits: optional keep-bit
// buffer of attn_drop_bits_bytes(B) bytes written by the forward and consumed by the backward (may be null).
size_t attn_drop_bits_bytes(int B) { return (size_t)B * 4 * (kS / 16) * (kS / 64) * 32 * sizeof(uint32_t); }

int attn_forward(const bf16* qkv, bf16* out, float* lse2, int B, uint32_t key, uint32_t thresh16, uint32_t* drop_bits,
                 cudaStream_t s) {
  ProfScope _ps("attn_fwd", s);
  FOCR_REQUIRE(B >= 1 && B <= 1024, "attention: B=%d out of range", B);
  const int smem = 2 * kTileBytes;
  const float inv_keep = 65536.f / (65536.f - (float)thresh16);
  static bool init = false;
  if (!init) {
    int rc = set_smem(attn_fwd_kernel<0, kFwdWarps>, smem);
    if (rc) return rc;
    rc = set_smem(attn_fwd_kernel<1, kFwdWarps>, smem);
    if (rc) return rc;
    rc = set_smem(attn_fwd_kernel<2, kFwdWarps>, smem);
    if (rc) return rc;
    init = true;
  }
  const dim3 grid(B * 4), block(kFwdWarps * 32);
  if (!thresh16)
    attn_fwd_kernel<0, kFwdWarps><<<grid, block, smem, s>>>(qkv, out, lse2, key, 0, 1.f, nullptr);
  else if (!drop_bits)
    attn_fwd_kernel<1, kFwdWarps><<<grid, block, smem, s>>>(qkv, out, lse2, key, thresh16, inv_keep, nullptr);
  else
    attn_fwd_kernel<2, kFwdWarps><<<grid, block, smem, s>>>(qkv, out, lse2, key, thresh16, inv_keep, drop_bits);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int attn_backward(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse2, float* dsum, bf16* dqkv, int B,
                  uint32_t key, uint32_t thresh16, const uint32_t* drop_bits, cudaStream_t s) {
  FOCR_REQUIRE(B >= 1 && B <= 1024, "attention: B=%d out of range", B);
  const int smem_a = 2 * kTileBytes, smem_b = 2 * kTileBytes + 2 * kS * 4 + 2 * 16384;
  const float inv_keep = 65536.f / (65536.f - (float)thresh16);
  static bool init = false;
  if (!init) {
    int rc = set_smem(attn_bwd_dq_kernel<0, kDqWarps>, smem_a);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dq_kernel<1, kDqWarps>, smem_a);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dq_kernel<2, kDqWarps>, smem_a);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dkv_kernel<0>, smem_b);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dkv_kernel<1>, smem_b);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dkv_kernel<2>, smem_b);
    if (rc) return rc;
    init = true;
  }
  const dim3 grid(B * 4);
  {
    ProfScope ps("attn_bwd_dq", s);
    if (!thresh16)
      attn_bwd_dq_kernel<0, kDqWarps><<<grid, kDqWarps * 32, smem_a, s>>>(qkv, o, d_o, lse2, dsum, dqkv, key, 0, 1.f,
                                                                          nullptr);
    else if (!drop_bits)
      attn_bwd_dq_kernel<1, kDqWarps><<<grid, kDqWarps * 32, smem_a, s>>>(qkv, o, d_o, lse2, dsum, dqkv, key, thresh16,
                                                                          inv_keep, nullptr);
    else
      attn_bwd_dq_kernel<2, kDqWarps><<<grid, kDqWarps * 32, smem_a, s>>>(qkv, o, d_o, lse2, dsum, dqkv, key, thresh16,
                                                                          inv_keep, drop_bits);
    FOCR_LAUNCH_CHECK();
  }
  {
    ProfScope ps("attn_bwd_dkv", s);
    if (!thresh16)
      attn_bwd_dkv_kernel<0><<<grid, 256, smem_b, s>>>(qkv, d_o, lse2, dsum, dqkv, key, 0, 1.f, nullptr);
    else if (!drop_bits)
      attn_bwd_dkv_kernel<1><<<grid, 256, smem_b, s>>>(qkv, d_o, lse2, dsum, dqkv, key, thresh16, inv_keep, nullptr);
    else
      attn_bwd_dkv_kernel<2><<<grid, 256, smem_b, s>>>(qkv, d_o, lse2, dsum, dqkv, key, thresh16, inv_keep, drop_bits);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}
