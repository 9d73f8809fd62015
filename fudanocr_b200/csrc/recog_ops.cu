// Building blocks of the TRAINABLE ResNet + Transformer recognisers:
//   stroke-level-decomposition/model/transformer.py  (Decoder :289-317, attention :227-241, LayerNorm :244-254,
//   Embeddings :277-286, PositionalEncoding :168-186, Generator :266-274, BasicBlock :43-73) and train.py:63-77
//   (CrossEntropyLoss over the packed valid positions, Adadelta(lr 1, rho 0.9)); image-ids-CTR/model/transformer.py and
//   train.py:63-90 use the same pieces with weight decay 1e-4.
// Everything here is small next to the 38-conv encoder (which runs on the tcgen05 GEMM): the decoder sees <= B*32 text
// positions.  The kernels are plain SIMT with fp32 math, one warp per row / CTA per (sample, head), every reduction in a
// fixed order (bit-reproducible), no atomics.
//   * decoder attention, masked-self or cross, any d_k <= 256 (h = 4, d_k = 256 here): forward keeps the post-dropout map
//     the reference returns as 'map'; backward recomputes the softmax and takes the keep mask from the stored map
//   * the reference's LayerNorm (unbiased std, eps added to the std) for feature widths 512 / 1024 with parameter gradients
//   * embedding * sqrt(d) concatenated with the (dropped-out) sinusoid table, and its gradient
//   * cross entropy over the positions t < length[b], mean over the packed total, with the logits gradient
//   * element-wise dropout, add + ReLU and its backward, 2x2 max-pool wrappers, Adadelta over a chunk table
#include "kernels.cuh"

#include <cmath>
#include <string.h>

namespace {

constexpr float kNegInf = -INFINITY;

// Dropout epoch: a device word XOR-ed into every dropout key of this file.  0 (the default) leaves the keys as the host passed
// them - the masks the oracle twin (oracle/dropout_rng.py) reproduces.  A training step replayed as a CUDA graph has its seed
// arguments frozen at capture time; the trainer then launches recog_epoch_advance_kernel at the head of every replay, so each
// step draws fresh masks, and the backward kernels of the same replay see the same word.
__device__ uint32_t g_recog_epoch = 0u;
__global__ void recog_epoch_advance_kernel() {
  uint32_t x = g_recog_epoch + 0x9E3779B9u;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  g_recog_epoch = x | 1u;
}
__global__ void recog_epoch_set_kernel(uint32_t v) { g_recog_epoch = v; }

__device__ __forceinline__ bool keep16(uint32_t key, unsigned long long e, uint32_t th16) {
  const uint32_t h = drop_hash32(key, (uint32_t)(e >> 1));
  const uint32_t lane = (e & 1) ? (h >> 16) : (h & 0xFFFFu);
  return lane >= th16;
}

__device__ __forceinline__ float ldbf(const bf16* p) { return __bfloat162float(*p); }

// ------------------------------------------------------------------------------------------------------------------
// decoder attention.  q rows (b*Tq + i) with leading dimension ld_q, head h at columns [h*dk, (h+1)*dk); k / v rows
// (b*Tk + j).  One CTA per (b, head), 8 warps.  smem: S[Tq][Tk] fp32 (+ second array in the backward).
// ------------------------------------------------------------------------------------------------------------------
// lane l owns the DKV consecutive elements [l * DKV, (l + 1) * DKV) of a head row: ONE 4- / 8- / 16-byte load per row and lane
// (the first version gave lane l the elements v * 32 + l: eight 2-byte loads per row, and the key loops were bound by their latency)
template <int DKV>  // d_k / 32 elements per lane
__device__ __forceinline__ void load_row(const bf16* p, int lane, float (&r)[DKV]) {
  static_assert(DKV == 2 || DKV == 4 || DKV == 8, "d_k in {64, 128, 256}");
  uint32_t w[DKV / 2];
  if constexpr (DKV == 2) {
    w[0] = *reinterpret_cast<const uint32_t*>(p + lane * 2);
  } else if constexpr (DKV == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(p + lane * 4);
    w[0] = u.x; w[1] = u.y;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(p + lane * 8);
    w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
  }
#pragma unroll
  for (int v = 0; v < DKV / 2; ++v) {
    const float2 f = unpack_bf16x2(w[v]);
    r[2 * v] = f.x;
    r[2 * v + 1] = f.y;
  }
}
template <int DKV>
__device__ __forceinline__ void store_row(bf16* p, int lane, const float (&r)[DKV]) {
  uint32_t w[DKV / 2];
#pragma unroll
  for (int v = 0; v < DKV / 2; ++v) w[v] = pack_bf16x2(r[2 * v], r[2 * v + 1]);
  if constexpr (DKV == 2) {
    *reinterpret_cast<uint32_t*>(p + lane * 2) = w[0];
  } else if constexpr (DKV == 4) {
    *reinterpret_cast<uint2*>(p + lane * 4) = make_uint2(w[0], w[1]);
  } else {
    *reinterpret_cast<uint4*>(p + lane * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// a[u] = x . row(j + u) for four consecutive rows, summed over the warp; the four row loads are issued before any of the math
template <int DKV>
__device__ __forceinline__ void dot_rows4(const bf16* __restrict__ base, long ld, int lane, const float (&x)[DKV], float (&a)[4]) {
  float r[4][DKV];
#pragma unroll
  for (int u = 0; u < 4; ++u) load_row<DKV>(base + (long)u * ld, lane, r[u]);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    a[u] = 0.f;
#pragma unroll
    for (int v = 0; v < DKV; ++v) a[u] += x[v] * r[u][v];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) a[u] = warp_sum(a[u]);
}
// acc += sum_u c[u] * row(j + u) for four consecutive rows
template <int DKV>
__device__ __forceinline__ void axpy_rows4(const bf16* __restrict__ base, long ld, int lane, const float (&c)[4], float (&acc)[DKV]) {
  float r[4][DKV];
#pragma unroll
  for (int u = 0; u < 4; ++u) load_row<DKV>(base + (long)u * ld, lane, r[u]);
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < DKV; ++v) acc[v] += c[u] * r[u][v];
}

// rows [i0, i1) of softmax(q k^T * scale [+ causal mask]) into S (row i at S + (i - i0) * Tk), one warp per row
template <int DKV>
__device__ void attn_probs(const bf16* __restrict__ q, long ld_q, const bf16* __restrict__ k, long ld_k, int i0, int i1, int Tk,
                           int causal, float scale, float* __restrict__ S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = i0 + warp; i < i1; i += nw) {
    float qr[DKV];
    load_row<DKV>(q + (long)i * ld_q, lane, qr);
    float* s = S + (long)(i - i0) * Tk;
    int j = 0;
    for (; j + 4 <= Tk; j += 4) {  // four key rows in flight: the loop is bound by the latency of the row loads otherwise
      float a[4];
      dot_rows4<DKV>(k + (long)j * ld_k, ld_k, lane, qr, a);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s[j + u] = (causal && j + u > i) ? kNegInf : a[u] * scale;
      }
    }
    for (; j < Tk; ++j) {
      float kr[DKV];
      load_row<DKV>(k + (long)j * ld_k, lane, kr);
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < DKV; ++v) a += qr[v] * kr[v];
      a = warp_sum(a);
      if (lane == 0) s[j] = (causal && j > i) ? kNegInf : a * scale;
    }
    __syncwarp();
    float m = kNegInf;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(s[j] - m);
      s[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < Tk; j += 32) s[j] *= inv;
  }
}

template <int DKV>
__global__ void __launch_bounds__(256) mha_small_fwd_kernel(const bf16* __restrict__ q, long ld_q, const bf16* __restrict__ k,
                                                            long ld_k, const bf16* __restrict__ v, long ld_v,
                                                            bf16* __restrict__ out, long ld_o, float* __restrict__ map,
                                                            int H, int Tq, int Tk, int causal, float scale, uint32_t key,
                                                            uint32_t th16, float keep_scale) {
  extern __shared__ float sm_att[];
  float* S = sm_att;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  constexpr int dk = DKV * 32;
  q += (long)b * Tq * ld_q + h * dk;
  k += (long)b * Tk * ld_k + h * dk;
  v += (long)b * Tk * ld_v + h * dk;
  out += (long)b * Tq * ld_o + h * dk;
  attn_probs<DKV>(q, ld_q, k, ld_k, 0, Tq, Tk, causal, scale, S);
  __syncthreads();
  // dropout on P (transformer.py:238-240); the dropped-out map is what the reference returns and multiplies into V
  float* mp = map + (long)blockIdx.x * Tq * Tk;
  for (int e = threadIdx.x; e < Tq * Tk; e += blockDim.x) {
    float p = S[e];
    if (th16) p = keep16(key ^ g_recog_epoch, (unsigned long long)blockIdx.x * Tq * Tk + e, th16) ? p * keep_scale : 0.f;
    S[e] = p;
    mp[e] = p;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < Tq; i += nw) {
    float acc[DKV];
#pragma unroll
    for (int c = 0; c < DKV; ++c) acc[c] = 0.f;
    const float* s = S + (long)i * Tk;
    int j = 0;
    for (; j + 4 <= Tk; j += 4) {
      const float c4[4] = {s[j], s[j + 1], s[j + 2], s[j + 3]};
      axpy_rows4<DKV>(v + (long)j * ld_v, ld_v, lane, c4, acc);
    }
    for (; j < Tk; ++j) {
      const float p = s[j];
      float vr[DKV];
      load_row<DKV>(v + (long)j * ld_v, lane, vr);
#pragma unroll
      for (int c = 0; c < DKV; ++c) acc[c] += p * vr[c];
    }
    store_row<DKV>(out + (long)i * ld_o, lane, acc);
  }
}

template <int DKV>
__global__ void __launch_bounds__(256) mha_small_bwd_kernel(const bf16* __restrict__ q, long ld_q, const bf16* __restrict__ k,
                                                            long ld_k, const bf16* __restrict__ v, long ld_v,
                                                            const bf16* __restrict__ d_out, long ld_o,
                                                            const float* __restrict__ map, bf16* __restrict__ dq, long ld_dq,
                                                            bf16* __restrict__ dk_, long ld_dk, bf16* __restrict__ dv,
                                                            long ld_dv, int H, int Tq, int Tk, int causal, float scale,
                                                            float keep_scale) {
  extern __shared__ float sm_att[];
  float* P = sm_att;                  // [Tq][Tk] softmax probabilities (pre-dropout), recomputed
  float* D = sm_att + (long)Tq * Tk;  // [Tq][Tk] dP, then dS
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  constexpr int dk = DKV * 32;
  q += (long)b * Tq * ld_q + h * dk;
  k += (long)b * Tk * ld_k + h * dk;
  v += (long)b * Tk * ld_v + h * dk;
  d_out += (long)b * Tq * ld_o + h * dk;
  dq += (long)b * Tq * ld_dq + h * dk;
  dk_ += (long)b * Tk * ld_dk + h * dk;
  dv += (long)b * Tk * ld_dv + h * dk;
  const float* mp = map + (long)blockIdx.x * Tq * Tk;
  attn_probs<DKV>(q, ld_q, k, ld_k, 0, Tq, Tk, causal, scale, P);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // dP = keep ? (dO . v_j) * keep_scale : 0 ; then dS = P (dP - sum_j P dP) * scale   (rows are warp-private)
  for (int i = warp; i < Tq; i += nw) {
    float g[DKV];
    load_row<DKV>(d_out + (long)i * ld_o, lane, g);
    float* d = D + (long)i * Tk;
    const float* p = P + (long)i * Tk;
    const float* m = mp + (long)i * Tk;
    int j = 0;
    for (; j + 4 <= Tk; j += 4) {
      float a[4];
      dot_rows4<DKV>(v + (long)j * ld_v, ld_v, lane, g, a);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) d[j + u] = m[j + u] != 0.f ? a[u] * keep_scale : 0.f;
      }
    }
    for (; j < Tk; ++j) {
      float vr[DKV];
      load_row<DKV>(v + (long)j * ld_v, lane, vr);
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < DKV; ++c) a += g[c] * vr[c];
      a = warp_sum(a) * keep_scale;
      if (lane == 0) d[j] = m[j] != 0.f ? a : 0.f;
    }
    __syncwarp();
    float delta = 0.f;
    for (int j = lane; j < Tk; j += 32) delta += p[j] * d[j];
    delta = warp_sum(delta);
    for (int j = lane; j < Tk; j += 32) d[j] = p[j] * (d[j] - delta) * scale;
    __syncwarp();
    // dq_i = sum_j dS_ij k_j
    float acc[DKV];
#pragma unroll
    for (int c = 0; c < DKV; ++c) acc[c] = 0.f;
    for (j = 0; j + 4 <= Tk; j += 4) {
      const float c4[4] = {d[j], d[j + 1], d[j + 2], d[j + 3]};
      axpy_rows4<DKV>(k + (long)j * ld_k, ld_k, lane, c4, acc);
    }
    for (; j < Tk; ++j) {
      const float ds = d[j];
      float kr[DKV];
      load_row<DKV>(k + (long)j * ld_k, lane, kr);
#pragma unroll
      for (int c = 0; c < DKV; ++c) acc[c] += ds * kr[c];
    }
    store_row<DKV>(dq + (long)i * ld_dq, lane, acc);
  }
  __syncthreads();
  // dk_j = sum_i dS_ij q_i ; dv_j = sum_i map_ij dO_i
  for (int j = warp; j < Tk; j += nw) {
    float ak[DKV], av[DKV];
#pragma unroll
    for (int c = 0; c < DKV; ++c) ak[c] = av[c] = 0.f;
    for (int i = 0; i < Tq; ++i) {
      const float ds = D[(long)i * Tk + j];
      const float pm = mp[(long)i * Tk + j];
      if (ds != 0.f) {
        float qr[DKV];
        load_row<DKV>(q + (long)i * ld_q, lane, qr);
#pragma unroll
        for (int c = 0; c < DKV; ++c) ak[c] += ds * qr[c];
      }
      if (pm != 0.f) {
        float g[DKV];
        load_row<DKV>(d_out + (long)i * ld_o, lane, g);
#pragma unroll
        for (int c = 0; c < DKV; ++c) av[c] += pm * g[c];
      }
    }
    store_row<DKV>(dk_ + (long)j * ld_dk, lane, ak);
    store_row<DKV>(dv + (long)j * ld_dv, lane, av);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The same attention for key counts whose (Tq, Tk) score tile does not fit shared memory (the 2 560 image tokens of 32 x 320
// crops): the grid also splits the query rows, kMhaRows (one per warp) per CTA, and the backward runs in two launches - a row
// pass (P, dP, dS, dq; dS goes to a workspace) and a key pass (dk, dv from the dS / map columns staged in shared memory).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMhaRows = 8;
constexpr int kMhaKeyBlk = 32;

template <int DKV>
__global__ void __launch_bounds__(256) mha_rows_fwd_kernel(const bf16* __restrict__ q, long ld_q, const bf16* __restrict__ k,
                                                           long ld_k, const bf16* __restrict__ v, long ld_v,
                                                           bf16* __restrict__ out, long ld_o, float* __restrict__ map,
                                                           int H, int Tq, int Tk, int causal, float scale, uint32_t key,
                                                           uint32_t th16, float keep_scale) {
  extern __shared__ float sm_att[];
  float* S = sm_att;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int i0 = blockIdx.y * kMhaRows, i1 = min(Tq, i0 + kMhaRows);
  constexpr int dk = DKV * 32;
  q += (long)b * Tq * ld_q + h * dk;
  k += (long)b * Tk * ld_k + h * dk;
  v += (long)b * Tk * ld_v + h * dk;
  out += (long)b * Tq * ld_o + h * dk;
  attn_probs<DKV>(q, ld_q, k, ld_k, i0, i1, Tk, causal, scale, S);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i0 + warp;          // rows are warp-private from here on (blockDim.x / 32 == kMhaRows)
  if (i >= i1) return;
  __syncwarp();
  float* s = S + (long)warp * Tk;
  float* mp = map + ((long)blockIdx.x * Tq + i) * Tk;
  const uint32_t epoch = g_recog_epoch;
  for (int j = lane; j < Tk; j += 32) {
    float p = s[j];
    if (th16) p = keep16(key ^ epoch, ((unsigned long long)blockIdx.x * Tq + i) * Tk + j, th16) ? p * keep_scale : 0.f;
    s[j] = p;
    mp[j] = p;
  }
  __syncwarp();
  float acc[DKV];
#pragma unroll
  for (int c = 0; c < DKV; ++c) acc[c] = 0.f;
  int j = 0;
  for (; j + 4 <= Tk; j += 4) {
    const float c4[4] = {s[j], s[j + 1], s[j + 2], s[j + 3]};
    axpy_rows4<DKV>(v + (long)j * ld_v, ld_v, lane, c4, acc);
  }
  for (; j < Tk; ++j) {
    const float p = s[j];
    float vr[DKV];
    load_row<DKV>(v + (long)j * ld_v, lane, vr);
#pragma unroll
    for (int c = 0; c < DKV; ++c) acc[c] += p * vr[c];
  }
  store_row<DKV>(out + (long)i * ld_o, lane, acc);
}

template <int DKV>
__global__ void __launch_bounds__(256) mha_rows_bwd_kernel(const bf16* __restrict__ q, long ld_q, const bf16* __restrict__ k,
                                                           long ld_k, const bf16* __restrict__ v, long ld_v,
                                                           const bf16* __restrict__ d_out, long ld_o,
                                                           const float* __restrict__ map, bf16* __restrict__ dq, long ld_dq,
                                                           float* __restrict__ ds_out, int H, int Tq, int Tk, int causal,
                                                           float scale, float keep_scale) {
  extern __shared__ float sm_att[];
  float* P = sm_att;                          // [kMhaRows][Tk] softmax probabilities (pre-dropout), recomputed
  float* D = sm_att + (long)kMhaRows * Tk;    // [kMhaRows][Tk] dP, then dS
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int i0 = blockIdx.y * kMhaRows, i1 = min(Tq, i0 + kMhaRows);
  constexpr int dk = DKV * 32;
  q += (long)b * Tq * ld_q + h * dk;
  k += (long)b * Tk * ld_k + h * dk;
  v += (long)b * Tk * ld_v + h * dk;
  d_out += (long)b * Tq * ld_o + h * dk;
  dq += (long)b * Tq * ld_dq + h * dk;
  attn_probs<DKV>(q, ld_q, k, ld_k, i0, i1, Tk, causal, scale, P);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = i0 + warp;
  if (i >= i1) return;
  __syncwarp();
  float g[DKV];
  load_row<DKV>(d_out + (long)i * ld_o, lane, g);
  float* d = D + (long)warp * Tk;
  const float* p = P + (long)warp * Tk;
  const float* m = map + ((long)blockIdx.x * Tq + i) * Tk;
  int j = 0;
  for (; j + 4 <= Tk; j += 4) {
    float a[4];
    dot_rows4<DKV>(v + (long)j * ld_v, ld_v, lane, g, a);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) d[j + u] = m[j + u] != 0.f ? a[u] * keep_scale : 0.f;
    }
  }
  for (; j < Tk; ++j) {
    float vr[DKV];
    load_row<DKV>(v + (long)j * ld_v, lane, vr);
    float a = 0.f;
#pragma unroll
    for (int c = 0; c < DKV; ++c) a += g[c] * vr[c];
    a = warp_sum(a) * keep_scale;
    if (lane == 0) d[j] = m[j] != 0.f ? a : 0.f;
  }
  __syncwarp();
  float delta = 0.f;
  for (int j = lane; j < Tk; j += 32) delta += p[j] * d[j];
  delta = warp_sum(delta);
  float* dso = ds_out + ((long)blockIdx.x * Tq + i) * Tk;
  for (int j = lane; j < Tk; j += 32) {
    const float x = p[j] * (d[j] - delta) * scale;
    d[j] = x;
    dso[j] = x;
  }
  __syncwarp();
  float acc[DKV];
#pragma unroll
  for (int c = 0; c < DKV; ++c) acc[c] = 0.f;
  for (j = 0; j + 4 <= Tk; j += 4) {
    const float c4[4] = {d[j], d[j + 1], d[j + 2], d[j + 3]};
    axpy_rows4<DKV>(k + (long)j * ld_k, ld_k, lane, c4, acc);
  }
  for (; j < Tk; ++j) {
    const float ds = d[j];
    float kr[DKV];
    load_row<DKV>(k + (long)j * ld_k, lane, kr);
#pragma unroll
    for (int c = 0; c < DKV; ++c) acc[c] += ds * kr[c];
  }
  store_row<DKV>(dq + (long)i * ld_dq, lane, acc);
}

// dk_j = sum_i dS_ij q_i ; dv_j = sum_i map_ij dO_i for the kMhaKeyBlk keys of blockIdx.y, one warp per key at a time
template <int DKV>
__global__ void __launch_bounds__(256) mha_keys_bwd_kernel(const bf16* __restrict__ q, long ld_q, const bf16* __restrict__ d_out,
                                                           long ld_o, const float* __restrict__ map,
                                                           const float* __restrict__ ds, bf16* __restrict__ dk_, long ld_dk,
                                                           bf16* __restrict__ dv, long ld_dv, int H, int Tq, int Tk) {
  extern __shared__ float sm_att[];
  float* sD = sm_att;                           // [Tq][kMhaKeyBlk]
  float* sM = sm_att + (long)Tq * kMhaKeyBlk;   // [Tq][kMhaKeyBlk]
  bf16* sQ = reinterpret_cast<bf16*>(sM + (long)Tq * kMhaKeyBlk);   // [Tq][dk]: every key of the block walks all query rows
  bf16* sG = sQ + (long)Tq * DKV * 32;                               // [Tq][dk] dO rows
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int j0 = blockIdx.y * kMhaKeyBlk;
  constexpr int dk = DKV * 32;
  q += (long)b * Tq * ld_q + h * dk;
  d_out += (long)b * Tq * ld_o + h * dk;
  dk_ += (long)b * Tk * ld_dk + h * dk;
  dv += (long)b * Tk * ld_dv + h * dk;
  const float* mp = map + (long)blockIdx.x * Tq * Tk;
  const float* dp = ds + (long)blockIdx.x * Tq * Tk;
  for (int e = threadIdx.x; e < Tq * kMhaKeyBlk; e += blockDim.x) {
    const int i = e / kMhaKeyBlk, jj = e % kMhaKeyBlk;
    const bool in = j0 + jj < Tk;
    sD[e] = in ? dp[(long)i * Tk + j0 + jj] : 0.f;
    sM[e] = in ? mp[(long)i * Tk + j0 + jj] : 0.f;
  }
  for (int e = threadIdx.x; e < Tq * dk; e += blockDim.x) {
    const int i = e / dk, c = e % dk;
    sQ[e] = q[(long)i * ld_q + c];
    sG[e] = d_out[(long)i * ld_o + c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int jj = warp; jj < kMhaKeyBlk && j0 + jj < Tk; jj += nw) {
    float ak[DKV], av[DKV];
#pragma unroll
    for (int c = 0; c < DKV; ++c) ak[c] = av[c] = 0.f;
    for (int i = 0; i < Tq; ++i) {
      const float x = sD[i * kMhaKeyBlk + jj];
      const float pm = sM[i * kMhaKeyBlk + jj];
      if (x != 0.f) {
        float qr[DKV];
        load_row<DKV>(sQ + (long)i * dk, lane, qr);
#pragma unroll
        for (int c = 0; c < DKV; ++c) ak[c] += x * qr[c];
      }
      if (pm != 0.f) {
        float g[DKV];
        load_row<DKV>(sG + (long)i * dk, lane, g);
#pragma unroll
        for (int c = 0; c < DKV; ++c) av[c] += pm * g[c];
      }
    }
    const int j = j0 + jj;
    store_row<DKV>(dk_ + (long)j * ld_dk, lane, ak);
    store_row<DKV>(dv + (long)j * ld_dv, lane, av);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm of the reference (transformer.py:244-254): a (x - mean) / (std_unbiased + eps) + b over C = NV*256 features.
// One warp per row; lane l owns the 8-element vectors at columns v*256 + l*8.
// ------------------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void ld_row_vec(const bf16* row, int lane, float (&x)[NV * 8]) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + v * 256 + lane * 8);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    x[v * 8 + 0] = a.x; x[v * 8 + 1] = a.y; x[v * 8 + 2] = b.x; x[v * 8 + 3] = b.y;
    x[v * 8 + 4] = c.x; x[v * 8 + 5] = c.y; x[v * 8 + 6] = d.x; x[v * 8 + 7] = d.y;
  }
}
template <int NV>
__device__ __forceinline__ void st_row_vec(bf16* row, int lane, const float (&x)[NV * 8]) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    uint4 u;
    u.x = pack_bf16x2(x[v * 8 + 0], x[v * 8 + 1]);
    u.y = pack_bf16x2(x[v * 8 + 2], x[v * 8 + 3]);
    u.z = pack_bf16x2(x[v * 8 + 4], x[v * 8 + 5]);
    u.w = pack_bf16x2(x[v * 8 + 6], x[v * 8 + 7]);
    *reinterpret_cast<uint4*>(row + v * 256 + lane * 8) = u;
  }
}

template <int NV>
__global__ void __launch_bounds__(256) lnw_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res,
                                                      const float* __restrict__ a, const float* __restrict__ bb,
                                                      bf16* __restrict__ sum_out, bf16* __restrict__ y, long T, float eps) {
  constexpr int C = NV * 256;
  const int lane = threadIdx.x & 31;
  const long row0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long stride = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = row0; r < T; r += stride) {
    float xv[NV * 8];
    ld_row_vec<NV>(x + r * C, lane, xv);
    if (res) {  // y = LN(x + res); the bf16-rounded sum is what the backward sees
      float rv[NV * 8];
      ld_row_vec<NV>(res + r * C, lane, rv);
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) xv[i] = __bfloat162float(__float2bfloat16(xv[i] + rv[i]));
      if (sum_out) st_row_vec<NV>(sum_out + r * C, lane, xv);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) s += xv[i];
    const float mean = warp_sum(s) * (1.f / C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) {
      xv[i] -= mean;
      ss += xv[i] * xv[i];
    }
    const float sd = sqrtf(warp_sum(ss) * (1.f / (C - 1)));
    const float rr = 1.f / (sd + eps);
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = v * 256 + lane * 8 + e;
        xv[v * 8 + e] = a[c] * xv[v * 8 + e] * rr + bb[c];
      }
    st_row_vec<NV>(y + r * C, lane, xv);
  }
}

// dx and per-CTA partial sums of (da, db): partial[cta][2][C]
template <int NV>
__global__ void __launch_bounds__(256) lnw_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                      const float* __restrict__ a, bf16* __restrict__ dx,
                                                      float* __restrict__ partial, long T, float eps) {
  constexpr int C = NV * 256;
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long row0 = (long)blockIdx.x * 8 + warp;
  const long stride = (long)gridDim.x * 8;
  float da[NV * 8], db[NV * 8];
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) da[i] = db[i] = 0.f;
  for (long r = row0; r < T; r += stride) {
    float xv[NV * 8], g[NV * 8];
    ld_row_vec<NV>(x + r * C, lane, xv);
    ld_row_vec<NV>(dy + r * C, lane, g);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) s += xv[i];
    const float mean = warp_sum(s) * (1.f / C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) {
      xv[i] -= mean;
      ss += xv[i] * xv[i];
    }
    const float sd = sqrtf(warp_sum(ss) * (1.f / (C - 1)));
    const float rr = 1.f / (sd + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = v * 8 + e;
        const float gy = g[i];
        da[i] += gy * xv[i] * rr;
        db[i] += gy;
        g[i] = gy * a[v * 256 + lane * 8 + e];  // d / d xhat
        sg += g[i];
        sgx += g[i] * xv[i];
      }
    sg = warp_sum(sg) * (1.f / C);
    sgx = warp_sum(sgx);
    // dx_i = r (g_i - mean g) - r^2 (sum_j g_j xc_j) / ((C-1) s) xc_i ; s = 0 only for a constant row (then xc = 0)
    const float coef = sd > 0.f ? rr * rr * sgx / ((float)(C - 1) * sd) : 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) g[i] = rr * (g[i] - sg) - coef * xv[i];
    st_row_vec<NV>(dx + r * C, lane, g);
  }
  // cross-warp reduction of the column sums, 64 columns at a time
  float* out = partial + (long)blockIdx.x * 2 * C;
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      for (int half = 0; half < 4; ++half) {  // lanes [8*half, 8*half+8) of this vector: 64 columns
        __syncthreads();
        if ((lane >> 3) == half) {
#pragma unroll
          for (int e = 0; e < 8; ++e) red[warp][(lane & 7) * 8 + e] = which ? db[v * 8 + e] : da[v * 8 + e];
        }
        __syncthreads();
        if (threadIdx.x < 64) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
          out[(long)which * C + v * 256 + half * 64 + threadIdx.x] = t;
        }
      }
    }
  }
}

__global__ void lnw_reduce_kernel(const float* __restrict__ partial, int P, int C, float* __restrict__ da,
                                  float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  double s = 0.0;
  for (int p = 0; p < P; ++p) s += partial[(long)p * 2 * C + i];
  if (i < C) da[i] = (float)s;
  else db[i - C] = (float)s;
}

// ------------------------------------------------------------------------------------------------------------------
// text embedding (transformer.py:277-286, :168-186, :346-348): out[b,t] = [ lut[idx] * sqrt(E) | dropout(pe[t]) ], E = 512
// ------------------------------------------------------------------------------------------------------------------
__global__ void text_embed_pe_kernel(const long long* __restrict__ idx, const float* __restrict__ lut, int vocab, int E,
                                     long rows, int T, long rows_pad, bf16* __restrict__ out, uint32_t key, uint32_t th16,
                                     float keep_scale, int* __restrict__ status) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_pad * 2 * E) return;
  const long r = i / (2 * E);
  const int c = (int)(i - r * 2 * E);
  float val = 0.f;
  if (r < rows) {
    if (c < E) {
      long long t = idx[r];
      if (t < 0 || t >= vocab) {
        atomicExch(status, 2);
        t = 0;
      }
      val = lut[t * E + c] * sqrtf((float)E);
    } else {
      const int pos = (int)(r % T), cc = c - E;
      const float div = expf((float)(cc & ~1) * (-logf(10000.f) / (float)E));
      const float ang = (float)pos * div;
      val = (cc & 1) ? cosf(ang) : sinf(ang);
      if (th16) val = keep16(key ^ g_recog_epoch, (unsigned long long)r * E + cc, th16) ? val * keep_scale : 0.f;
    }
  }
  out[i] = __float2bfloat16(val);
}

// d_lut[v][c] = sqrt(E) * sum over rows with idx == v of d_out[row][c].  One CTA per table entry v: warp 0 compacts the matching
// rows IN ORDER (ballot + popc) into shared memory, chunk by chunk, then every thread sums its columns over that list - fixed
// order, no atomics, and a 4303-entry table (image-ids-CTR) costs 4303 x rows comparisons instead of 4303 x 512 x rows.
constexpr int kEmbChunk = 2048;
__global__ void __launch_bounds__(128) text_embed_bwd_kernel(const long long* __restrict__ idx, const bf16* __restrict__ d_out,
                                                             int vocab, int E, long rows, float* __restrict__ d_lut) {
  __shared__ int list[kEmbChunk];
  __shared__ int count;
  const int v = blockIdx.x, lane = threadIdx.x & 31;
  float acc[8];                                      // columns threadIdx.x + 128 * j, j < E / 128 (E <= 1024)
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long r0 = 0; r0 < rows; r0 += kEmbChunk) {
    const int n = (int)((rows - r0) < kEmbChunk ? (rows - r0) : kEmbChunk);
    if (threadIdx.x < 32) {
      int cnt = 0;
      for (int i = 0; i < n; i += 32) {
        const bool hit = (i + lane < n) && idx[r0 + i + lane] == v;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) list[cnt + __popc(m & ((1u << lane) - 1))] = i + lane;
        cnt += __popc(m);
      }
      if (lane == 0) count = cnt;
    }
    __syncthreads();
    const int cnt = count;
    for (int k = 0; k < cnt; ++k) {
      const bf16* row = d_out + (r0 + list[k]) * 2 * E;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = threadIdx.x + 128 * j;
        if (c < E) acc[j] += ldbf(row + c);
      }
    }
    __syncthreads();
  }
  const float sc = sqrtf((float)E);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = threadIdx.x + 128 * j;
    if (c < E) d_lut[(long)v * E + c] = acc[j] * sc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// cross entropy over the valid positions (transformer.py:361-373 packs rows t < length[b]; train.py:71 takes the mean)
// logits fp32 [(b*T + t)][ld]; gt packed int64 (sum of lengths).  One CTA per sample, one warp per position.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) packed_ce_kernel(const float* __restrict__ logits, long ld, int B, int T, int C,
                                                        const long long* __restrict__ length, const long long* __restrict__ gt,
                                                        float gscale, float* __restrict__ partial, bf16* __restrict__ d_logits,
                                                        long ld_d, int* __restrict__ status) {
  __shared__ float wl[4];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long off = 0, total = 0;
  for (int i = 0; i < B; ++i) {
    long long l = length[i];
    l = l < 0 ? 0 : (l > T ? T : l);
    if (i < b) off += l;
    total += l;
  }
  long long len = length[b];
  if (len < 0 || len > T) {
    if (threadIdx.x == 0) atomicExch(status, 1);
    len = len < 0 ? 0 : T;
  }
  const float inv_total = total > 0 ? 1.f / (float)total : 0.f;
  float loss = 0.f;
  for (int t = warp; t < T; t += 4) {
    const float* row = logits + ((long)b * T + t) * ld;
    bf16* drow = d_logits ? d_logits + ((long)b * T + t) * ld_d : nullptr;
    if (t >= len) {
      if (drow)
        for (int c = lane; c < ld_d; c += 32) drow[c] = __float2bfloat16(0.f);
      continue;
    }
    long long g = gt[off + t];
    if (g < 0 || g >= C) {
      if (lane == 0) atomicExch(status, 2);
      g = 0;
    }
    float m = kNegInf;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    loss += lse - row[g];
    if (drow)
      for (int c = lane; c < ld_d; c += 32) {
        float d = 0.f;
        if (c < C) d = (expf(row[c] - lse) - (c == (int)g ? 1.f : 0.f)) * gscale * inv_total;
        drow[c] = __float2bfloat16(d);
      }
  }
  if (lane == 0) wl[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0) partial[b] = (wl[0] + wl[1] + wl[2] + wl[3]) * inv_total;
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)red[0];
}

// ------------------------------------------------------------------------------------------------------------------
// element-wise pieces
// ------------------------------------------------------------------------------------------------------------------
// y = keep ? x * keep_scale : 0 with the mask regenerated from (key, element index): the same call is its own backward
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long n, uint32_t key, uint32_t th16,
                               float keep_scale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = keep16(key ^ g_recog_epoch, (unsigned long long)i, th16) ? __float2bfloat16(ldbf(x + i) * keep_scale) : __float2bfloat16(0.f);
}
// BasicBlock tail (transformer.py:69-71): y = relu(a + b), 8 elements per thread
__global__ void add_relu_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y, long n8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 ua = a[i], ub = b[i];
  const uint32_t* pa = &ua.x;
  const uint32_t* pb = &ub.x;
  uint4 o;
  uint32_t* po = &o.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fa = unpack_bf16x2(pa[e]), fb = unpack_bf16x2(pb[e]);
    po[e] = pack_bf16x2(fmaxf(fa.x + fb.x, 0.f), fmaxf(fa.y + fb.y, 0.f));
  }
  y[i] = o;
}
// dx = y > 0 ? dy : 0 (gradient of both summands of add_relu, and of a plain ReLU given its output)
__global__ void relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx, long n8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 ug = dy[i], uy = y[i];
  const uint32_t* pg = &ug.x;
  const uint32_t* py = &uy.x;
  uint4 o;
  uint32_t* po = &o.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fg = unpack_bf16x2(pg[e]), fy = unpack_bf16x2(py[e]);
    po[e] = pack_bf16x2(fy.x > 0.f ? fg.x : 0.f, fy.y > 0.f ? fg.y : 0.f);
  }
  dx[i] = o;
}

// Adadelta (torch.optim.Adadelta; stroke-level-decomposition/train.py:32-36 lr 1, rho 0.9; image-ids-CTR adds weight decay)
struct AdaChunk {
  float* p;
  const float* g;
  float* sq;
  float* acc;
  long long n;
};
__global__ void __launch_bounds__(256) adadelta_kernel(const AdaChunk* __restrict__ chunks, float gscale, float lr, float rho,
                                                       float eps, float wd) {
  const AdaChunk c = chunks[blockIdx.x];
  for (long long i = threadIdx.x; i < c.n; i += 256) {
    float w = c.p[i];
    float g = c.g[i] * gscale;
    if (wd != 0.f) g += wd * w;
    const float sq = rho * c.sq[i] + (1.f - rho) * g * g;
    const float acc = c.acc[i];
    const float delta = sqrtf(acc + eps) / sqrtf(sq + eps) * g;
    c.sq[i] = sq;
    c.acc[i] = rho * acc + (1.f - rho) * delta * delta;
    c.p[i] = w - lr * delta;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// image-ids-CTR head (train.py:71-80): text_pred / ||text_pred||, and the distance term -MSE(text_pred_n, text_features[gt])
// ------------------------------------------------------------------------------------------------------------------
// y[r] = x[r] / ||x[r]||_2 (fp32 in, bf16 out), inv[r] = 1 / ||x[r]|| (0 for an all-zero row).  One warp per row.
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, long ld, bf16* __restrict__ y,
                                                         float* __restrict__ inv, long T, int C) {
  const int lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= T) return;
  const float* row = x + r * ld;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) ss += row[c] * row[c];
  ss = warp_sum(ss);
  const float iv = ss > 0.f ? 1.f / sqrtf(ss) : 0.f;
  for (int c = lane; c < C; c += 32) y[r * C + c] = __float2bfloat16(row[c] * iv);
  if (lane == 0) inv[r] = iv;
}
// dx = inv (dy - y (dy . y))
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                         const float* __restrict__ inv, bf16* __restrict__ dx, long ld_dx, long T,
                                                         int C) {
  const int lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= T) return;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot += ldbf(dy + r * C + c) * ldbf(y + r * C + c);
  dot = warp_sum(dot);
  const float iv = inv[r];
  for (int c = lane; c < C; c += 32)
    dx[r * ld_dx + c] = __float2bfloat16(iv * (ldbf(dy + r * C + c) - ldbf(y + r * C + c) * dot));
}
// partial[b] = sum over t < length[b], c of (y[b,t,c] - feats[gt][c])^2 / (total * C); d_y = gscale * 2 (y - feats[gt]) / (total * C)
__global__ void __launch_bounds__(128) packed_feat_mse_kernel(const bf16* __restrict__ y, int B, int T, int C,
                                                              const long long* __restrict__ length,
                                                              const long long* __restrict__ gt, const float* __restrict__ feats,
                                                              int V, float gscale, float* __restrict__ partial,
                                                              bf16* __restrict__ d_y, int* __restrict__ status) {
  __shared__ float wl[4];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long off = 0, total = 0;
  for (int i = 0; i < B; ++i) {
    long long l = length[i];
    l = l < 0 ? 0 : (l > T ? T : l);
    if (i < b) off += l;
    total += l;
  }
  long long len = length[b];
  len = len < 0 ? 0 : (len > T ? T : len);
  const float inv_n = total > 0 ? 1.f / ((float)total * (float)C) : 0.f;
  float acc = 0.f;
  for (int t = warp; t < T; t += 4) {
    const long r = (long)b * T + t;
    if (t >= len) {
      if (d_y)
        for (int c = lane; c < C; c += 32) d_y[r * C + c] = __float2bfloat16(0.f);
      continue;
    }
    long long g = gt[off + t];
    if (g < 0 || g >= V) {
      if (lane == 0) atomicExch(status, 2);
      g = 0;
    }
    const float* f = feats + g * C;
    for (int c = lane; c < C; c += 32) {
      const float d = ldbf(y + r * C + c) - f[c];
      acc += d * d;
      if (d_y) d_y[r * C + c] = __float2bfloat16(gscale * 2.f * d * inv_n);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) wl[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) partial[b] = (wl[0] + wl[1] + wl[2] + wl[3]) * inv_n;
}


// dst[c][r] = src[r][c] for a row-major bf16 matrix (rows x cols, leading dimension ld_src) - 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16* __restrict__ src, long ld_src, long rows, long cols,
                                                             bf16* __restrict__ dst, long ld_dst) {
  __shared__ bf16 tile[32][34];
  const long c0 = (long)blockIdx.x * 32, r0 = (long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < cols && r < rows) dst[c * ld_dst + r] = tile[tx][ty + 8 * i];
  }
}


// colT[k][m] = x[pixel m shifted by tap(k)][channel(k)], k = tap * C + c  (the transpose of the im2col matrix, written directly):
// one CTA turns a [64 pixels][64 channels] window of x (16-byte loads along the channels) into 64 rows x 128 bytes of colT through a
// transposing shared-memory tile.  C % 64 == 0, so a 64-wide block of k lies inside one tap; blocks past 9C are the zero padding.
__global__ void __launch_bounds__(256) im2col3x3_t_kernel(const bf16* __restrict__ x, bf16* __restrict__ colT, int B, int H, int W, int C,
                                                          long M) {
  __shared__ __align__(16) bf16 tile[64][72];
  const long m0 = (long)blockIdx.x * 64;
  const int k0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  if (k0 < 9 * C) {
    const int tap = k0 / C, c0 = k0 - tap * C;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int i = t & 63;
    const long m = m0 + i;
    const int w = (int)(m % W), h = (int)((m / W) % H);
    const long b = m / ((long)W * H);
    const int hh = h + dy, ww = w + dx;
    const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
    const bf16* src = x + ((b * H + (ok ? hh : 0)) * W + (ok ? ww : 0)) * C + c0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int v = (t >> 6) + 4 * pass;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (ok) u = *reinterpret_cast<const uint4*>(src + v * 8);
      const bf16* e = reinterpret_cast<const bf16*>(&u);
#pragma unroll
      for (int q = 0; q < 8; ++q) tile[v * 8 + q][i] = e[q];
    }
  } else {
    for (int q = t; q < 64 * 72; q += 256) (&tile[0][0])[q] = __float2bfloat16(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int idx = t + 256 * pass;
    const int c = idx >> 3, j = idx & 7;
    *reinterpret_cast<uint4*>(colT + (long)(k0 + c) * M + m0 + j * 8) = *reinterpret_cast<const uint4*>(&tile[c][j * 8]);
  }
}

int egrid(long n, int per) { return (int)((n + per - 1) / per); }

template <int DKV>
int launch_mha_fwd(const bf16* q, long ld_q, const bf16* k, long ld_k, const bf16* v, long ld_v, bf16* out, long ld_o, float* map,
                   int B, int H, int Tq, int Tk, int causal, float scale, uint32_t key, uint32_t th16, float ks, size_t smem,
                   cudaStream_t s) {
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(mha_small_fwd_kernel<DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mha_small_fwd_kernel<DKV><<<B * H, 256, smem, s>>>(q, ld_q, k, ld_k, v, ld_v, out, ld_o, map, H, Tq, Tk, causal, scale, key,
                                                      th16, ks);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
template <int DKV>
int launch_mha_bwd(const bf16* q, long ld_q, const bf16* k, long ld_k, const bf16* v, long ld_v, const bf16* d_out, long ld_o,
                   const float* map, bf16* dq, long ld_dq, bf16* dk, long ld_dk, bf16* dv, long ld_dv, int B, int H, int Tq, int Tk,
                   int causal, float scale, float ks, size_t smem, cudaStream_t s) {
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(mha_small_bwd_kernel<DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mha_small_bwd_kernel<DKV><<<B * H, 256, smem, s>>>(q, ld_q, k, ld_k, v, ld_v, d_out, ld_o, map, dq, ld_dq, dk, ld_dk, dv, ld_dv,
                                                      H, Tq, Tk, causal, scale, ks);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

template <int DKV>
int launch_mha_rows_fwd(const bf16* q, long ld_q, const bf16* k, long ld_k, const bf16* v, long ld_v, bf16* out, long ld_o,
                        float* map, int B, int H, int Tq, int Tk, int causal, float scale, uint32_t key, uint32_t th16, float ks,
                        cudaStream_t s) {
  const size_t smem = (size_t)kMhaRows * Tk * 4;
  FOCR_REQUIRE(smem <= 200 * 1024, "mha_small_fwd: %d keys do not fit the row-split kernel's shared memory", Tk);
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(mha_rows_fwd_kernel<DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mha_rows_fwd_kernel<DKV><<<dim3(B * H, (Tq + kMhaRows - 1) / kMhaRows), kMhaRows * 32, smem, s>>>(
      q, ld_q, k, ld_k, v, ld_v, out, ld_o, map, H, Tq, Tk, causal, scale, key, th16, ks);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
template <int DKV>
int launch_mha_rows_bwd(const bf16* q, long ld_q, const bf16* k, long ld_k, const bf16* v, long ld_v, const bf16* d_out, long ld_o,
                        const float* map, bf16* dq, long ld_dq, bf16* dk, long ld_dk, bf16* dv, long ld_dv, float* ds, int B, int H,
                        int Tq, int Tk, int causal, float scale, float ks, cudaStream_t s) {
  const size_t smem = (size_t)kMhaRows * Tk * 8;
  FOCR_REQUIRE(smem <= 200 * 1024, "mha_small_bwd: %d keys do not fit the row-split kernel's shared memory", Tk);
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(mha_rows_bwd_kernel<DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mha_rows_bwd_kernel<DKV><<<dim3(B * H, (Tq + kMhaRows - 1) / kMhaRows), kMhaRows * 32, smem, s>>>(
      q, ld_q, k, ld_k, v, ld_v, d_out, ld_o, map, dq, ld_dq, ds, H, Tq, Tk, causal, scale, ks);
  FOCR_LAUNCH_CHECK();
  const size_t smem2 = (size_t)Tq * kMhaKeyBlk * 8 + (size_t)Tq * DKV * 32 * 2 * 2;
  FOCR_REQUIRE(smem2 <= 200 * 1024, "mha_small_bwd: Tq = %d too long for the key pass", Tq);
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(mha_keys_bwd_kernel<DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mha_keys_bwd_kernel<DKV><<<dim3(B * H, (Tk + kMhaKeyBlk - 1) / kMhaKeyBlk), 256, smem2, s>>>(q, ld_q, d_out, ld_o, map, ds, dk,
                                                                                                ld_dk, dv, ld_dv, H, Tq, Tk);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// the single-CTA kernels keep the whole (Tq, Tk) tile in shared memory; beyond that the row-split pair takes over
bool mha_fits_one_cta(int Tq, int Tk, int bytes_per_score) {
  static int force_rows = -1;
  if (force_rows < 0) force_rows = getenv("FOCR_MHA_ROWS") ? 1 : 0;   // test / tuning knob: always the row-split kernels
  return !force_rows && (size_t)Tq * Tk * bytes_per_score <= 200 * 1024;
}

uint32_t th16_of(float p) {
  if (p <= 0.f) return 0;
  const long t = (long)(p * 65536.0 + 0.5);
  return (uint32_t)(t > 65535 ? 65535 : t);
}

}  // namespace

extern "C" {

// dropout epoch of the recogniser kernels (see g_recog_epoch): set it (0 = keys exactly as passed), or advance it on the device
int focr_recog_epoch_set(unsigned value, void* stream) {
  recog_epoch_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(value);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int focr_recog_epoch_advance(void* stream) {
  recog_epoch_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>();
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// decoder attention core.  q (B*Tq, ld_q) / k, v (B*Tk, ld_k / ld_v) / out (B*Tq, ld_o) bf16, head h at columns h*d_k;
// map fp32 (B, H, Tq, Tk) = dropout(softmax(q k^T / sqrt(d_k) [+ causal mask])) - the tensor the reference returns.
int focr_mha_small_fwd(const void* q, long ld_q, const void* k, long ld_k, const void* v, long ld_v, void* out, long ld_o,
                       float* map, int B, int H, int d_k, int Tq, int Tk, int causal, float p_drop, unsigned seed,
                       unsigned stream_id, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(q && k && v && out && map, "mha_small_fwd: null pointer");
  FOCR_REQUIRE(B >= 1 && H >= 1 && Tq >= 1 && Tk >= 1, "mha_small_fwd: B=%d H=%d Tq=%d Tk=%d", B, H, Tq, Tk);
  FOCR_REQUIRE(d_k == 64 || d_k == 128 || d_k == 256, "mha_small_fwd: d_k %d (64, 128 or 256)", d_k);
  FOCR_REQUIRE(!causal || Tq == Tk, "mha_small_fwd: the causal mask needs Tq == Tk");
  FOCR_REQUIRE((ld_q | ld_k | ld_v | ld_o) % 8 == 0 && (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0,
               "mha_small_fwd: rows must be 16-byte aligned (leading dimensions multiples of 8 elements)");
  FOCR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "mha_small_fwd: p_drop %f", p_drop);
  const size_t smem = (size_t)Tq * Tk * 4;
  ProfScope _ps("mha_small_fwd", s);
  const uint32_t th = th16_of(p_drop);
  const float ks = 65536.f / (65536.f - (float)th);
  const float scale = 1.f / sqrtf((float)d_k);
  const uint32_t key = drop_key(seed, stream_id);
  const bf16 *qq = (const bf16*)q, *kk = (const bf16*)k, *vv = (const bf16*)v;
  if (!mha_fits_one_cta(Tq, Tk, 4)) {
    if (d_k == 64) return launch_mha_rows_fwd<2>(qq, ld_q, kk, ld_k, vv, ld_v, (bf16*)out, ld_o, map, B, H, Tq, Tk, causal, scale, key, th, ks, s);
    if (d_k == 128) return launch_mha_rows_fwd<4>(qq, ld_q, kk, ld_k, vv, ld_v, (bf16*)out, ld_o, map, B, H, Tq, Tk, causal, scale, key, th, ks, s);
    return launch_mha_rows_fwd<8>(qq, ld_q, kk, ld_k, vv, ld_v, (bf16*)out, ld_o, map, B, H, Tq, Tk, causal, scale, key, th, ks, s);
  }
  if (d_k == 64) return launch_mha_fwd<2>(qq, ld_q, kk, ld_k, vv, ld_v, (bf16*)out, ld_o, map, B, H, Tq, Tk, causal, scale, key, th, ks, smem, s);
  if (d_k == 128) return launch_mha_fwd<4>(qq, ld_q, kk, ld_k, vv, ld_v, (bf16*)out, ld_o, map, B, H, Tq, Tk, causal, scale, key, th, ks, smem, s);
  return launch_mha_fwd<8>(qq, ld_q, kk, ld_k, vv, ld_v, (bf16*)out, ld_o, map, B, H, Tq, Tk, causal, scale, key, th, ks, smem, s);
}

// bytes of workspace focr_mha_small_bwd_ws needs: the dS tile (B, H, Tq, Tk) fp32 of the row-split backward, 0 when the whole
// score tile of a (sample, head) fits one CTA's shared memory
size_t focr_mha_small_bwd_workspace_bytes(int B, int H, int Tq, int Tk) {
  if (mha_fits_one_cta(Tq, Tk, 8)) return 0;
  return (size_t)B * H * Tq * Tk * 4;
}

// gradients of the above w.r.t. q, k, v given d_out and the stored map (its zeros are the dropped positions)
int focr_mha_small_bwd_ws(const void* q, long ld_q, const void* k, long ld_k, const void* v, long ld_v, const void* d_out, long ld_o,
                          const float* map, void* dq, long ld_dq, void* dk, long ld_dk, void* dv, long ld_dv, int B, int H, int d_k,
                          int Tq, int Tk, int causal, float p_drop, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(q && k && v && d_out && map && dq && dk && dv, "mha_small_bwd: null pointer");
  FOCR_REQUIRE(B >= 1 && H >= 1 && Tq >= 1 && Tk >= 1, "mha_small_bwd: B=%d H=%d Tq=%d Tk=%d", B, H, Tq, Tk);
  FOCR_REQUIRE(d_k == 64 || d_k == 128 || d_k == 256, "mha_small_bwd: d_k %d (64, 128 or 256)", d_k);
  FOCR_REQUIRE(!causal || Tq == Tk, "mha_small_bwd: the causal mask needs Tq == Tk");
  FOCR_REQUIRE((ld_q | ld_k | ld_v | ld_o | ld_dq | ld_dk | ld_dv) % 8 == 0 &&
                   (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)d_out | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0,
               "mha_small_bwd: rows must be 16-byte aligned (leading dimensions multiples of 8 elements)");
  const size_t smem = (size_t)Tq * Tk * 8;
  ProfScope _ps("mha_small_bwd", s);
  const uint32_t th = th16_of(p_drop);
  const float ks = 65536.f / (65536.f - (float)th);
  const float scale = 1.f / sqrtf((float)d_k);
  const bf16 *qq = (const bf16*)q, *kk = (const bf16*)k, *vv = (const bf16*)v, *gg = (const bf16*)d_out;
  if (!mha_fits_one_cta(Tq, Tk, 8)) {
    FOCR_REQUIRE(ws != nullptr && ws_bytes >= focr_mha_small_bwd_workspace_bytes(B, H, Tq, Tk),
                 "mha_small_bwd: Tq*Tk = %d*%d needs a %zu-byte workspace (focr_mha_small_bwd_workspace_bytes)", Tq, Tk,
                 focr_mha_small_bwd_workspace_bytes(B, H, Tq, Tk));
    float* ds = (float*)ws;
    if (d_k == 64) return launch_mha_rows_bwd<2>(qq, ld_q, kk, ld_k, vv, ld_v, gg, ld_o, map, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, ds, B, H, Tq, Tk, causal, scale, ks, s);
    if (d_k == 128) return launch_mha_rows_bwd<4>(qq, ld_q, kk, ld_k, vv, ld_v, gg, ld_o, map, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, ds, B, H, Tq, Tk, causal, scale, ks, s);
    return launch_mha_rows_bwd<8>(qq, ld_q, kk, ld_k, vv, ld_v, gg, ld_o, map, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, ds, B, H, Tq, Tk, causal, scale, ks, s);
  }
  if (d_k == 64) return launch_mha_bwd<2>(qq, ld_q, kk, ld_k, vv, ld_v, gg, ld_o, map, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, B, H, Tq, Tk, causal, scale, ks, smem, s);
  if (d_k == 128) return launch_mha_bwd<4>(qq, ld_q, kk, ld_k, vv, ld_v, gg, ld_o, map, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, B, H, Tq, Tk, causal, scale, ks, smem, s);
  return launch_mha_bwd<8>(qq, ld_q, kk, ld_k, vv, ld_v, gg, ld_o, map, (bf16*)dq, ld_dq, (bf16*)dk, ld_dk, (bf16*)dv, ld_dv, B, H, Tq, Tk, causal, scale, ks, smem, s);
}

// the same without a workspace: score tiles that fit one CTA only
int focr_mha_small_bwd(const void* q, long ld_q, const void* k, long ld_k, const void* v, long ld_v, const void* d_out, long ld_o,
                       const float* map, void* dq, long ld_dq, void* dk, long ld_dk, void* dv, long ld_dv, int B, int H, int d_k,
                       int Tq, int Tk, int causal, float p_drop, void* stream) {
  return focr_mha_small_bwd_ws(q, ld_q, k, ld_k, v, ld_v, d_out, ld_o, map, dq, ld_dq, dk, ld_dk, dv, ld_dv, B, H, d_k, Tq, Tk, causal,
                               p_drop, nullptr, 0, stream);
}

// y = LN(x [+ res]) over C in {512, 1024} features; sum_out (optional) receives x + res (the tensor the backward needs)
int focr_layernorm_wide_fwd(const void* x, const void* res, const float* a, const float* b, void* sum_out, void* y, long T, int C,
                            float eps, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(x && a && b && y && T >= 1, "layernorm_wide_fwd: null pointer / T=%ld", T);
  FOCR_REQUIRE(C == 512 || C == 1024, "layernorm_wide_fwd: C %d (512 or 1024)", C);
  ProfScope _ps("ln_wide_fwd", s);
  long g = (T + 7) / 8;
  if (g > 148L * 8) g = 148L * 8;
  if (C == 512)
    lnw_fwd_kernel<2><<<(int)g, 256, 0, s>>>((const bf16*)x, (const bf16*)res, a, b, (bf16*)sum_out, (bf16*)y, T, eps);
  else
    lnw_fwd_kernel<4><<<(int)g, 256, 0, s>>>((const bf16*)x, (const bf16*)res, a, b, (bf16*)sum_out, (bf16*)y, T, eps);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

size_t focr_layernorm_wide_workspace_bytes(int C) { return (size_t)148 * 2 * 2 * C * sizeof(float); }

int focr_layernorm_wide_bwd(const void* dy, const void* x, const float* a, void* dx, float* da, float* db, long T, int C, float eps,
                            void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(dy && x && a && dx && da && db && ws && T >= 1, "layernorm_wide_bwd: null pointer / T=%ld", T);
  FOCR_REQUIRE(C == 512 || C == 1024, "layernorm_wide_bwd: C %d (512 or 1024)", C);
  FOCR_REQUIRE(ws_bytes >= focr_layernorm_wide_workspace_bytes(C), "layernorm_wide_bwd: workspace too small");
  ProfScope _ps("ln_wide_bwd", s);
  long g = (T + 7) / 8;
  if (g > 148L * 2) g = 148L * 2;
  float* partial = (float*)ws;
  if (C == 512)
    lnw_bwd_kernel<2><<<(int)g, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, a, (bf16*)dx, partial, T, eps);
  else
    lnw_bwd_kernel<4><<<(int)g, 256, 0, s>>>((const bf16*)dy, (const bf16*)x, a, (bf16*)dx, partial, T, eps);
  FOCR_LAUNCH_CHECK();
  lnw_reduce_kernel<<<egrid(2 * C, 256), 256, 0, s>>>(partial, (int)g, C, da, db);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// out bf16 (rows_pad, 2E): rows = B*T text positions (row-major b, t), padding rows zero.  status: device int (0 ok, 2 index
// outside the table).
int focr_text_embed_fwd(const long long* idx, const float* lut, int vocab, int E, int B, int T, long rows_pad, void* out,
                        float p_drop, unsigned seed, unsigned stream_id, int* status, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(idx && lut && out && status, "text_embed_fwd: null pointer");
  FOCR_REQUIRE(vocab >= 1 && E >= 2 && (E & 1) == 0 && B >= 1 && T >= 1 && rows_pad >= (long)B * T, "text_embed_fwd: shape");
  ProfScope _ps("text_embed", s);
  const uint32_t th = th16_of(p_drop);
  FOCR_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  text_embed_pe_kernel<<<egrid(rows_pad * 2 * E, 256), 256, 0, s>>>(idx, lut, vocab, E, (long)B * T, T, rows_pad, (bf16*)out,
                                                                     drop_key(seed, stream_id), th, 65536.f / (65536.f - (float)th),
                                                                     status);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int focr_text_embed_bwd(const long long* idx, const void* d_out, int vocab, int E, int B, int T, float* d_lut, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(idx && d_out && d_lut, "text_embed_bwd: null pointer");
  ProfScope _ps("text_embed_bwd", s);
  FOCR_REQUIRE(vocab >= 1 && E >= 1 && E <= 1024, "text_embed_bwd: vocab=%d E=%d (E <= 1024)", vocab, E);
  text_embed_bwd_kernel<<<vocab, 128, 0, s>>>(idx, (const bf16*)d_out, vocab, E, (long)B * T, d_lut);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// loss[0] = mean over the packed valid positions of CE(logits[b,t,:C], gt); d_logits bf16 (B*T, ld_d) = gscale * d loss /
// d logits (zeros at t >= length[b] and at the padding columns), or NULL.  ws: (B + 1) floats.
size_t focr_packed_ce_workspace_bytes(int B) { return ((size_t)B + 4) * sizeof(float); }
int focr_packed_ce(const float* logits, long ld, int B, int T, int C, const long long* length, const long long* gt, float gscale,
                   float* loss, void* d_logits, long ld_d, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(logits && length && gt && loss && ws, "packed_ce: null pointer");
  FOCR_REQUIRE(B >= 1 && T >= 1 && C >= 2 && ld >= C && (!d_logits || ld_d >= C), "packed_ce: B=%d T=%d C=%d ld=%ld", B, T, C, ld);
  FOCR_REQUIRE(ws_bytes >= focr_packed_ce_workspace_bytes(B), "packed_ce: workspace too small");
  ProfScope _ps("packed_ce", s);
  float* partial = (float*)ws;
  int* status = (int*)(partial + B);
  FOCR_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  packed_ce_kernel<<<B, 128, 0, s>>>(logits, ld, B, T, C, length, gt, gscale, partial, (bf16*)d_logits, ld_d, status);
  FOCR_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 256, 0, s>>>(partial, B, loss);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int focr_dropout(const void* x, void* y, long n, float p_drop, unsigned seed, unsigned stream_id, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(x && y && n >= 1 && p_drop >= 0.f && p_drop < 1.f, "dropout: bad arguments");
  const uint32_t th = th16_of(p_drop);
  dropout_kernel<<<egrid(n, 256), 256, 0, s>>>((const bf16*)x, (bf16*)y, n, drop_key(seed, stream_id), th,
                                               65536.f / (65536.f - (float)th));
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int focr_add_relu(const void* a, const void* b, void* y, long n, void* stream) {
  FOCR_REQUIRE(a && b && y && n >= 8 && n % 8 == 0, "add_relu: n %ld must be a positive multiple of 8", n);
  add_relu_kernel<<<egrid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, (const uint4*)b, (uint4*)y, n / 8);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int focr_relu_bwd(const void* dy, const void* y, void* dx, long n, void* stream) {
  FOCR_REQUIRE(dy && y && dx && n >= 8 && n % 8 == 0, "relu_bwd: n %ld must be a positive multiple of 8", n);
  relu_bwd_kernel<<<egrid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)y, (uint4*)dx, n / 8);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// nn.MaxPool2d((2,2),(2,2)) on an NHWC bf16 map (transformer.py:84,130) and its backward (gradient to the first maximum)
int focr_maxpool2x2_fwd(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  FOCR_REQUIRE(x && y && (H & 1) == 0 && (W & 1) == 0 && C % 8 == 0, "maxpool2x2_fwd: H=%d W=%d C=%d", H, W, C);
  return maxpool_fwd((const bf16*)x, (bf16*)y, B, H, W, C, 2, (cudaStream_t)stream);
}
int focr_maxpool2x2_bwd(const void* x, const void* y, const void* dy, void* dx, int B, int H, int W, int C, void* stream) {
  FOCR_REQUIRE(x && y && dy && dx && (H & 1) == 0 && (W & 1) == 0 && C % 8 == 0, "maxpool2x2_bwd: H=%d W=%d C=%d", H, W, C);
  return maxpool_bwd((const bf16*)x, (const bf16*)y, (const bf16*)dy, (bf16*)dx, B, H, W, C, 2, (cudaStream_t)stream);
}

// chunks: device array of n_chunks records {param, grad, square_avg, acc_delta, n} (5 x int64), as focr_adam_clip_step
int focr_adadelta_step(const void* chunks, int n_chunks, float gscale, float lr, float rho, float eps, float weight_decay,
                       void* stream) {
  FOCR_REQUIRE(chunks && n_chunks > 0, "adadelta_step: empty chunk table");
  ProfScope _ps("adadelta", (cudaStream_t)stream);
  adadelta_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const AdaChunk*)chunks, gscale, lr, rho, eps, weight_decay);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// 3x3 convolutions of the recogniser encoder through an explicit im2col (transformer.py:80 conv1 with 3 input channels, and
// every weight gradient: dW = dY^T col is a plain token GEMM with K = 9 Ci, compute-bound for Ci >= 128).  Forward / input
// gradient of the 64..1024-channel layers go through the implicit GEMM (focr_conv2d_fwd / _dgrad) instead.
// ------------------------------------------------------------------------------------------------------------------
static inline size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t focr_conv3x3_gemm_workspace_bytes(int B, int H, int W, int Ci, int Co) {
  const size_t M = (size_t)B * H * W, Kp = up((size_t)9 * Ci, 128), Np = up((size_t)Co, 64);
  return up(M * Kp * 2, 256) + up(Np * Kp * 4, 256) + up(Np * 4, 256) + ((size_t)48 << 20);
}

// y bf16 (B*H*W, Co) = conv3x3(x) + bias.  x: NHWC bf16 (x_nhwc) or NCHW fp32 (x_nchw), exactly one non-NULL.
int focr_conv3x3_gemm_fwd(const void* x_nhwc, const float* x_nchw, const float* w, const float* bias, void* y, int B, int H, int W,
                          int Ci, int Co, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE((x_nhwc != nullptr) != (x_nchw != nullptr) && w && y && ws, "conv3x3_gemm_fwd: pointers");
  const long M = (long)B * H * W;
  FOCR_REQUIRE(M % 128 == 0 && Co % 64 == 0 && Ci >= 1, "conv3x3_gemm_fwd: B*H*W=%ld (mult. of 128) Co=%d (mult. of 64)", M, Co);
  FOCR_REQUIRE(ws_bytes >= focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co), "conv3x3_gemm_fwd: workspace too small");
  const int Kp = (int)up((size_t)9 * Ci, 128);
  bf16* col = (bf16*)ws;
  bf16* wb = (bf16*)((char*)ws + up((size_t)M * Kp * 2, 256));
  int rc = im2col3x3((const bf16*)x_nhwc, x_nchw, col, B, H, W, Ci, Kp, s);
  if (rc) return rc;
  rc = prep_stn_conv_w(w, wb, nullptr, Co, Ci, Co, Kp, s);
  if (rc) return rc;
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = Co;
  p.kh = p.kw = 1;
  p.W = 64;
  p.H = 2;
  p.epi = TC_EPI_BF16;
  p.ldc = Co;
  p.bias = bias;
  p.out = y;
  const bf16* ap[1] = {col};
  return tc_gemm_launch(ap, 1, Kp, (long)64 * Kp, (long)128 * Kp, Kp, (int)(M / 128), wb, Kp, p, s);
}

// dw fp32 [Co][Ci][3][3] and db [Co] (either may be NULL) from dy bf16 (B*H*W, Co) and the layer input
int focr_conv3x3_gemm_wgrad(const void* dy, const void* x_nhwc, const float* x_nchw, float* dw, float* db, int B, int H, int W,
                            int Ci, int Co, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE((x_nhwc != nullptr) != (x_nchw != nullptr) && dy && ws, "conv3x3_gemm_wgrad: pointers");
  const long M = (long)B * H * W;
  FOCR_REQUIRE(Co % 64 == 0 && Ci >= 1 && M >= 1, "conv3x3_gemm_wgrad: Co=%d (mult. of 64)", Co);
  FOCR_REQUIRE(ws_bytes >= focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co), "conv3x3_gemm_wgrad: workspace too small");
  const int Kp = (int)up((size_t)9 * Ci, 128);
  bf16* col = (bf16*)ws;
  float* tmpw = (float*)((char*)ws + up((size_t)M * Kp * 2, 256));
  float* tmpb = (float*)((char*)tmpw + up((size_t)Co * Kp * 4, 256));
  float* partial = (float*)((char*)tmpb + up((size_t)Co * 4, 256));
  FOCR_REQUIRE(linear_wgrad_partial_bytes(M, Co, Kp) <= ((size_t)48 << 20), "conv3x3_gemm_wgrad: partial buffer");
  int rc;
  if (dw) {
    rc = im2col3x3((const bf16*)x_nhwc, x_nchw, col, B, H, W, Ci, Kp, s);
    if (rc) return rc;
    rc = linear_wgrad((const bf16*)dy, Co, col, Kp, M, Co, Kp, tmpw, 1.f, partial, s);
    if (rc) return rc;
    rc = unpack_stn_conv_grad(tmpw, dw, Co, Ci, Kp, s);
    if (rc) return rc;
  }
  if (db) return colsum((const bf16*)dy, Co, M, Co, db, partial, s);
  return FOCR_OK;
}


// nn.BatchNorm2d in eval mode (running statistics) + activation (0 none, 2 relu) on a (T, C) bf16 matrix: the recogniser under
// model.eval() (SLD/train.py:88, test-time decode).  stats: fp32 [4][C] scratch.
int focr_bn_eval_fwd(const void* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     void* y, float* stats, long T, int C, int act, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(x && gamma && beta && running_mean && running_var && y && stats, "bn_eval_fwd: null pointer");
  int rc = bn_eval_stats(gamma, beta, running_mean, running_var, 1e-5f, C, stats, s);
  if (rc) return rc;
  return bn_apply((const bf16*)x, C, stats, (bf16*)y, C, T, C, act, nullptr, 0, nullptr, s);
}


// image-ids-CTR/train.py:76 - text_pred / text_pred.norm(dim=1, keepdim=True): x fp32 (T, ld) -> y bf16 (T, C), inv fp32 (T)
int focr_l2norm_rows_fwd(const float* x, long ld, void* y, float* inv, long T, int C, void* stream) {
  FOCR_REQUIRE(x && y && inv && T >= 1 && C >= 1 && ld >= C, "l2norm_rows_fwd: bad arguments");
  l2norm_fwd_kernel<<<egrid(T, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, (bf16*)y, inv, T, C);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
int focr_l2norm_rows_bwd(const void* dy, const void* y, const float* inv, void* dx, long ld_dx, long T, int C, void* stream) {
  FOCR_REQUIRE(dy && y && inv && dx && T >= 1 && C >= 1 && ld_dx >= C, "l2norm_rows_bwd: bad arguments");
  l2norm_bwd_kernel<<<egrid(T, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)y, inv, (bf16*)dx, ld_dx, T, C);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
// image-ids-CTR/train.py:66-69,79: nn.MSELoss(text_pred_n, text_features[text_gt]) over the packed valid rows.
// loss[0] = mean squared difference; d_y bf16 (B*T, C) = gscale * d loss / d y (zeros at t >= length[b]) or NULL.  ws: (B + 4) floats.
int focr_packed_feat_mse(const void* y, int B, int T, int C, const long long* length, const long long* gt, const float* feats, int V,
                         float gscale, float* loss, void* d_y, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(y && length && gt && feats && loss && ws, "packed_feat_mse: null pointer");
  FOCR_REQUIRE(B >= 1 && T >= 1 && C >= 1 && V >= 1, "packed_feat_mse: B=%d T=%d C=%d V=%d", B, T, C, V);
  FOCR_REQUIRE(ws_bytes >= focr_packed_ce_workspace_bytes(B), "packed_feat_mse: workspace too small");
  float* partial = (float*)ws;
  int* status = (int*)(partial + B);
  FOCR_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  packed_feat_mse_kernel<<<B, 128, 0, s>>>((const bf16*)y, B, T, C, length, gt, feats, V, gscale, partial, (bf16*)d_y, status);
  FOCR_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 256, 0, s>>>(partial, B, loss);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}


// Weight + bias gradient of a 3x3 convolution on the tcgen05 GEMM:  dW[Co][9 Ci] = dY^T (Co x M) . col (M x 9 Ci), the pixel
// dimension M = B*H*W being the contraction.  Both operands are brought to K-major form by explicit transposes (dY^T: Co x M,
// col^T: 9Ci x M) so that the same TMA / tcgen05 kernel that runs the forward GEMMs applies, fp32 accumulation in TMEM over all
// of M, fp32 output.  Co % 128 == 0 and M % 128 == 0 (the 64-channel stem keeps the streaming kernel).
size_t focr_conv3x3_wgrad_tc_workspace_bytes(int B, int H, int W, int Ci, int Co) {
  if (conv3x3_wgrad_tc_general_supported(B, H, W, Ci, Co))   // implicit operand: only the per-CTA partials and the bias-gradient scratch
    return up(conv3x3_wgrad_tc_general_partial_bytes(B, H, W, Ci, Co), 256) + up((size_t)Co * 4, 256) + ((size_t)16 << 20);
  const size_t M = (size_t)B * H * W, Kp = up((size_t)9 * Ci, 128);
  return up(M * Kp * 2, 256) + up((size_t)Co * M * 2, 256) + up((size_t)Co * Kp * 4, 256) + up((size_t)Co * 4, 256) + ((size_t)16 << 20);
}
int focr_conv3x3_wgrad_tc(const void* dy, const void* x_nhwc, const float* x_nchw, float* dw, float* db, int B, int H, int W, int Ci,
                          int Co, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(x_nhwc != nullptr && x_nchw == nullptr && dy && dw && ws, "conv3x3_wgrad_tc: pointers (NHWC bf16 input only)");
  const long M = (long)B * H * W;
  FOCR_REQUIRE(Co % 128 == 0 && M % 128 == 0 && Ci % 64 == 0 && Ci >= 64,
               "conv3x3_wgrad_tc: Co=%d (mult. of 128), Ci=%d (mult. of 64), B*H*W=%ld (mult. of 128)", Co, Ci, M);
  FOCR_REQUIRE(ws_bytes >= focr_conv3x3_wgrad_tc_workspace_bytes(B, H, W, Ci, Co), "conv3x3_wgrad_tc: workspace too small");
  if (conv3x3_wgrad_tc_general_supported(B, H, W, Ci, Co)) {
    // implicit form (wgrad_tc.cu): the shifted operand comes straight from x through TMA halo boxes, no col^T, no transposes
    float* part = (float*)ws;
    float* partial_b = (float*)((char*)ws + up(conv3x3_wgrad_tc_general_partial_bytes(B, H, W, Ci, Co), 256) + up((size_t)Co * 4, 256));
    {
      ProfScope _ps("conv_wgrad_tc", s);
      int rc2 = conv3x3_wgrad_tc_general((const bf16*)dy, (const bf16*)x_nhwc, B, H, W, Ci, Co, dw, part, s);
      if (rc2) return rc2;
    }
    if (db) return colsum((const bf16*)dy, Co, M, Co, db, partial_b, s);
    return FOCR_OK;
  }
  const int Kp = (int)up((size_t)9 * Ci, 128);
  char* base = (char*)ws;
  bf16* colT = (bf16*)base;
  base += up((size_t)M * Kp * 2, 256);
  bf16* dyT = (bf16*)base;
  base += up((size_t)Co * M * 2, 256);
  float* tmpw = (float*)base;
  base += up((size_t)Co * Kp * 4, 256);
  float* partial = (float*)(base + up((size_t)Co * 4, 256));
  int rc;
  {
    ProfScope _ps("wgrad_operands", s);
    im2col3x3_t_kernel<<<dim3((unsigned)(M / 64), Kp / 64), 256, 0, s>>>((const bf16*)x_nhwc, colT, B, H, W, Ci, M);
    FOCR_LAUNCH_CHECK();
    transpose_bf16_kernel<<<dim3(Co / 32, (unsigned)(M / 32)), 256, 0, s>>>((const bf16*)dy, Co, M, Co, dyT, M);
    FOCR_LAUNCH_CHECK();
  }
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = Kp;
  p.kh = p.kw = 1;
  p.W = 64;
  p.H = 2;
  p.epi = TC_EPI_F32;
  p.ldc = Kp;
  p.out = tmpw;
  const bf16* ap[1] = {dyT};
  {
    ProfScope _ps("conv_wgrad_tc", s);
    rc = tc_gemm_launch(ap, 1, M, (long)64 * M, (long)128 * M, (int)M, Co / 128, colT, (int)M, p, s);
    if (rc) return rc;
  }
  rc = unpack_stn_conv_grad(tmpw, dw, Co, Ci, Kp, s);
  if (rc) return rc;
  if (db) return colsum((const bf16*)dy, Co, M, Co, db, partial, s);
  return FOCR_OK;
}

}  // extern "C"
