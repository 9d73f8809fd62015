// CTC forward-backward over the CRNN's raw logits (north_star: "the CRNN conv stack and CTC forward-backward").
// The reference has no CTC-loss call site (its CRNN is an eval-only recogniser,
// scene-text-telescope/interfaces/super_resolution.py:143-158); the semantics adopted are those of the framework call the
// north star implies, torch.nn.functional.ctc_loss(log_softmax(logits, 2), targets, input_lengths, target_lengths,
// blank, reduction, zero_infinity), on the (T, B, C) layout CRNN.forward emits (model/crnn/crnn.py:78-80).
//
// One CTA per sample.  log-softmax is fused (only the per-frame log-sum-exp is kept, in shared memory); the alpha lattice
// (T x (2S+1), log domain) goes to a caller-provided workspace, the beta recursion keeps two rows in shared memory and
// emits the gradient with respect to the LOGITS frame by frame:  d nll / d u[t,k] = y[t,k] - (1/p(z|x)) * sum_{s: l'_s = k}
// alpha_t(s) beta_t(s) / y[t,k]   (Graves et al. 2006, eq. 16).  Every reduction runs in a fixed order: the result is
// bit-reproducible run to run.  Latency-bound (T sequential steps of a 2S+1-wide recurrence), not HBM- or tensor-bound.
#include "kernels.cuh"

#include <cmath>

namespace {

constexpr int kThreads = 128;
constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == kNegInf) return kNegInf;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// logits (T,B,C) fp32; targets (B,S_max) int64 padded; alpha_ws (B,T,Lp) fp32 with Lp = 2*S_max+1
__global__ void __launch_bounds__(kThreads) ctc_fwd_bwd_kernel(const float* __restrict__ logits, int T, int B, int C,
                                                               const long long* __restrict__ targets, int S_max,
                                                               const long long* __restrict__ input_lengths,
                                                               const long long* __restrict__ target_lengths, int blank,
                                                               int reduction, int zero_infinity, float grad_scale,
                                                               float* __restrict__ nll_out, float* __restrict__ d_logits,
                                                               float* __restrict__ alpha_ws, int* __restrict__ status) {
  extern __shared__ float sm_f[];
  const int Lp_cap = 2 * S_max + 1;
  float* lse = sm_f;                 // [T]
  float* a0 = lse + T;               // alpha rows (ping-pong), later beta rows
  float* a1 = a0 + Lp_cap;
  float* ab = a1 + Lp_cap;           // alpha_t(s) + beta_t(s)
  int* lab = reinterpret_cast<int*>(ab + Lp_cap);  // extended label l'_s
  int* owner = lab + Lp_cap;         // 1 if s is the first position carrying its class
  __shared__ float s_nll;

  const int b = blockIdx.x, tid = threadIdx.x;
  long long Tb_ll = input_lengths[b], S_ll = target_lengths[b];
  if (Tb_ll < 0 || Tb_ll > T || S_ll < 0 || S_ll > S_max) {  // malformed lengths: flag, emit zeros
    if (tid == 0) {
      atomicExch(status, 1);
      nll_out[b] = 0.f;
    }
    if (d_logits)
      for (long i = tid; i < (long)T * C; i += kThreads) d_logits[((i / C) * B + b) * C + i % C] = 0.f;
    return;
  }
  const int Tb = (int)Tb_ll, S = (int)S_ll, Lp = 2 * S + 1;
  bool bad_label = false;
  for (int s = tid; s < Lp; s += kThreads) {
    int l = blank;
    if (s & 1) {
      const long long v = targets[(long)b * S_max + (s >> 1)];
      if (v < 0 || v >= C) bad_label = true;  // flagged; treated as a blank so that no read goes out of bounds
      else l = (int)v;
    }
    lab[s] = l;
  }
  if (bad_label) atomicExch(status, 2);
  // per-frame log-sum-exp (one warp per frame)
  for (int t = tid >> 5; t < Tb; t += kThreads / 32) {
    const float* row = logits + ((long)t * B + b) * C;
    float m = kNegInf;
    for (int k = tid & 31; k < C; k += 32) m = fmaxf(m, row[k]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int k = tid & 31; k < C; k += 32) sum += expf(row[k] - m);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) lse[t] = m + logf(sum);
  }
  __syncthreads();
  for (int s = tid; s < Lp; s += kThreads) {
    int first = 1;
    const int l = lab[s];
    for (int q = (s & 1); q < s; q += 2)  // same parity: blanks sit on even, labels on odd positions
      if (lab[q] == l) {
        first = 0;
        break;
      }
    // a label equal to `blank` shares the blank's gradient slot: owned by s = 0
    if ((s & 1) && l == blank) first = 0;
    owner[s] = first;
  }
  float* alpha_b = alpha_ws + (long)b * T * Lp_cap;
  auto lp = [&](int t, int k) { return logits[((long)t * B + b) * C + k] - lse[t]; };

  // ---- alpha ----
  float* prev = a0;
  float* cur = a1;
  if (Tb > 0) {
    for (int s = tid; s < Lp; s += kThreads) {
      const float v = s < 2 ? lp(0, lab[s]) : kNegInf;
      prev[s] = v;
      alpha_b[s] = v;
    }
  }
  __syncthreads();
  for (int t = 1; t < Tb; ++t) {
    for (int s = tid; s < Lp; s += kThreads) {
      const int l = lab[s];
      const float x0 = prev[s];
      const float x1 = s >= 1 ? prev[s - 1] : kNegInf;
      const float x2 = (s >= 2 && (s & 1) && lab[s - 2] != l) ? prev[s - 2] : kNegInf;
      const float v = lse3(x0, x1, x2);
      const float r = v == kNegInf ? kNegInf : v + lp(t, l);
      cur[s] = r;
      alpha_b[(long)t * Lp_cap + s] = r;
    }
    __syncthreads();
    float* tmp = prev;
    prev = cur;
    cur = tmp;
  }
  if (tid == 0) {
    float l;
    if (Tb == 0) l = S == 0 ? 0.f : kNegInf;
    else l = lse3(prev[Lp - 1], Lp > 1 ? prev[Lp - 2] : kNegInf, kNegInf);
    s_nll = -l;
  }
  __syncthreads();
  float nll = s_nll;
  const bool infeasible = nll == INFINITY;
  if (tid == 0) nll_out[b] = (infeasible && zero_infinity) ? 0.f : nll;
  if (!d_logits) return;

  float gs = grad_scale;
  if (reduction == 1) gs = grad_scale / ((float)B * (float)(S > 0 ? S : 1));
  if (infeasible) {  // zero_infinity: zero gradient; otherwise the gradient is undefined (torch: nan) - zeros as well
    for (long i = tid; i < (long)T * C; i += kThreads) d_logits[((i / C) * B + b) * C + i % C] = 0.f;
    return;
  }
  // frames past the input length carry no gradient
  for (long i = (long)Tb * C + tid; i < (long)T * C; i += kThreads) d_logits[((i / C) * B + b) * C + i % C] = 0.f;

  // ---- beta + gradient ----
  __syncthreads();  // alpha rows in smem are dead from here: a0 / a1 become the beta rows
  float* bnext = a0;
  float* bcur = a1;
  for (int t = Tb - 1; t >= 0; --t) {
    for (int s = tid; s < Lp; s += kThreads) {
      const int l = lab[s];
      float r;
      if (t == Tb - 1) {
        r = s >= Lp - 2 ? lp(t, l) : kNegInf;
      } else {
        const float x0 = bnext[s];
        const float x1 = s + 1 < Lp ? bnext[s + 1] : kNegInf;
        const float x2 = (s + 2 < Lp && (s & 1) && lab[s + 2] != l) ? bnext[s + 2] : kNegInf;
        const float v = lse3(x0, x1, x2);
        r = v == kNegInf ? kNegInf : v + lp(t, l);
      }
      bcur[s] = r;
      ab[s] = alpha_b[(long)t * Lp_cap + s] + r;
    }
    __syncthreads();
    float* drow = d_logits + ((long)t * B + b) * C;
    const float* row = logits + ((long)t * B + b) * C;
    const float l_t = lse[t];
    for (int k = tid; k < C; k += kThreads) drow[k] = expf(row[k] - l_t) * gs;
    __syncthreads();
    for (int s = tid; s < Lp; s += kThreads) {
      if (!owner[s]) continue;
      const int l = lab[s];
      float m = kNegInf;
      if (s == 0) {  // blank: every even position, plus labels equal to the blank index
        for (int q = 0; q < Lp; ++q)
          if (lab[q] == l) m = fmaxf(m, ab[q]);
      } else {
        for (int q = s; q < Lp; q += 2)
          if (lab[q] == l) m = fmaxf(m, ab[q]);
      }
      float occ = 0.f;
      if (m != kNegInf) {
        float sum = 0.f;
        if (s == 0) {
          for (int q = 0; q < Lp; ++q)
            if (lab[q] == l) sum += expf(ab[q] - m);
        } else {
          for (int q = s; q < Lp; q += 2)
            if (lab[q] == l) sum += expf(ab[q] - m);
        }
        const float lpk = row[l] - l_t;
        occ = expf(m + logf(sum) + nll - lpk);
      }
      drow[l] = (expf(row[l] - l_t) - occ) * gs;
    }
    __syncthreads();
    float* tmp = bnext;
    bnext = bcur;
    bcur = tmp;
  }
}

// loss scalar in a fixed order: none -> untouched, mean -> mean_b(nll_b / max(S_b, 1)), sum -> sum_b nll_b
__global__ void ctc_reduce_kernel(const float* __restrict__ nll, const long long* __restrict__ target_lengths, int B,
                                  int reduction, float* __restrict__ loss) {
  __shared__ float part[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float v = nll[b];
    if (reduction == 1) {
      const long long S = target_lengths[b];
      v /= (float)(S > 0 ? S : 1);
    }
    acc += v;
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    loss[0] = reduction == 1 ? t / (float)B : t;
  }
}

size_t smem_bytes(int T, int S_max) { return ((size_t)T + 5 * (size_t)(2 * S_max + 1)) * 4; }

}  // namespace

extern "C" size_t focr_ctc_loss_workspace_bytes(int T, int B, int S_max) {
  if (T < 1 || B < 1 || S_max < 0) return 0;
  return (size_t)B * T * (2 * (size_t)S_max + 1) * sizeof(float) + 16;
}

extern "C" int focr_ctc_loss(const float* logits, int T, int B, int C, const long long* targets, int S_max,
                             const long long* input_lengths, const long long* target_lengths, int blank, int reduction,
                             int zero_infinity, float grad_scale, float* nll, float* loss, float* d_logits, void* ws,
                             size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(logits && input_lengths && target_lengths && nll && ws, "ctc_loss: null pointer");
  FOCR_REQUIRE(T >= 1 && B >= 1 && C >= 2 && S_max >= 0 && (S_max == 0 || targets), "ctc_loss: T=%d B=%d C=%d S_max=%d", T, B, C,
               S_max);
  FOCR_REQUIRE(blank >= 0 && blank < C, "ctc_loss: blank %d outside [0,%d)", blank, C);
  FOCR_REQUIRE(reduction >= 0 && reduction <= 2, "ctc_loss: reduction %d (0 none, 1 mean, 2 sum)", reduction);
  FOCR_REQUIRE(reduction == 0 || loss, "ctc_loss: loss pointer required for mean / sum");
  if (ws_bytes < focr_ctc_loss_workspace_bytes(T, B, S_max)) {
    focr_set_error("ctc_loss: workspace %zu < %zu bytes", ws_bytes, focr_ctc_loss_workspace_bytes(T, B, S_max));
    return FOCR_ERR_WORKSPACE;
  }
  const size_t smem = smem_bytes(T, S_max);
  FOCR_REQUIRE(smem <= 200 * 1024, "ctc_loss: T=%d S_max=%d needs %zu bytes of shared memory", T, S_max, smem);
  static size_t attr = 48 * 1024;
  if (smem > attr) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(ctc_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  ProfScope _ps("ctc_loss", s);
  float* alpha = reinterpret_cast<float*>(ws);
  int* status = reinterpret_cast<int*>(alpha + (size_t)B * T * (2 * (size_t)S_max + 1));
  FOCR_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  ctc_fwd_bwd_kernel<<<B, kThreads, smem, s>>>(logits, T, B, C, targets, S_max, input_lengths, target_lengths, blank, reduction,
                                               zero_infinity, grad_scale, nll, d_logits, alpha, status);
  FOCR_LAUNCH_CHECK();
  if (reduction != 0) {
    ctc_reduce_kernel<<<1, 256, 0, s>>>(nll, target_lengths, B, reduction, loss);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

// status word of the last focr_ctc_loss call on this workspace: 0 ok, 1 length out of range, 2 label outside [0, C)
extern "C" int focr_ctc_loss_status(const void* ws, int T, int B, int S_max, int* status_host, void* stream) {
  FOCR_REQUIRE(ws && status_host, "ctc_loss_status: null pointer");
  const float* alpha = reinterpret_cast<const float*>(ws);
  const int* status = reinterpret_cast<const int*>(alpha + (size_t)B * T * (2 * (size_t)S_max + 1));
  FOCR_CHECK_CUDA(cudaMemcpyAsync(status_host, status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FOCR_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return FOCR_OK;
}
