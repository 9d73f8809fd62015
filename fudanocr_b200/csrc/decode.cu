// KV-cached greedy test-time decode of the ResNet + Transformer recognisers.
// Reference loops: stroke-level-decomposition/train.py:110-121 and image-ids-CTR/train.py:118-134 - for i in range(max_length)
// the reference re-runs the WHOLE decoder (masked self-attention, cross-attention K / V projections over every image token,
// FFN, generator) on the growing prefix and keeps only the last position.  Here the cross-attention K / V of the image tokens
// are projected once, the self-attention K / V of every emitted position are cached, and each step runs the decoder layer on
// the ONE new position:  embedding | PE -> q, k, v -> attention over the cache -> W_o -> LN -> cross-attention -> W_o -> LN ->
// FFN -> LN -> generator -> arg-max + winning softmax probability, all on the device; the next token never visits the host.
// Decoder.forward: SLD/model/transformer.py:303-317 (h = 4, d_model = 1024, d_ff = 2048, eval mode: no dropout).
// The linears run on the tcgen05 GEMM (tc_gemm.cu) with weights converted to bf16 once per call of focr_recog_decode_prepare.
#include "kernels.cuh"
#include <string.h>

extern "C" int focr_layernorm_wide_fwd(const void* x, const void* res, const float* a, const float* b, void* sum_out, void* y,
                                       long T, int C, float eps, void* stream);

namespace {

constexpr int kD = 1024, kE = 512, kH = 4, kDk = 256, kFF = 2048;

inline size_t al(size_t n) { return (n + 255) & ~(size_t)255; }
inline int pad128(long n) { return (int)((n + 127) / 128 * 128); }
inline int gen_pad(int n) { return n <= 64 ? 64 : pad128(n); }

// prepared blob: bf16 GEMM weights + fp32 biases / LayerNorm parameters / embedding table (+ bf16 text features)
struct Blob {
  size_t wqkv, wo, wq2, wkv2, wo2, w1, w2, wg, feat;  // bf16
  size_t bqkv, bo, bq2, bkv2, bo2, b1, b2, bg;       // fp32
  size_t ln[6];                                       // a1, b1, a2, b2, a3, b3 (1024 each)
  size_t lut;                                         // fp32 (vocab, 512)
  size_t total;
};
Blob blob_layout(int vocab, int n_out, int n_feat) {
  Blob b;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
  const int ng = gen_pad(n_out);
  b.wqkv = take((size_t)3 * kD * kD * 2);
  b.wo = take((size_t)kD * kD * 2);
  b.wq2 = take((size_t)kD * kD * 2);
  b.wkv2 = take((size_t)2 * kD * kD * 2);
  b.wo2 = take((size_t)kD * kD * 2);
  b.w1 = take((size_t)kFF * kD * 2);
  b.w2 = take((size_t)kD * kFF * 2);
  b.wg = take((size_t)ng * kD * 2);
  b.feat = take((size_t)pad128(n_feat) * (size_t)ng * 2);
  b.bqkv = take(3 * kD * 4);
  b.bo = take(kD * 4);
  b.bq2 = take(kD * 4);
  b.bkv2 = take(2 * kD * 4);
  b.bo2 = take(kD * 4);
  b.b1 = take(kFF * 4);
  b.b2 = take(kD * 4);
  b.bg = take((size_t)ng * 4);
  for (int i = 0; i < 6; ++i) b.ln[i] = take(kD * 4);
  b.lut = take((size_t)vocab * kE * 4);
  b.total = o;
  return b;
}

struct Ws {
  size_t x, qkv, att, t0, r1s, r1, q2, a2, r2s, r2, hdn, r3s, r3, logits, gen_n, sim, inv, kc, vc, kv2;
  size_t total;
};
Ws ws_layout(int B, int n_tok, int T_max, int n_out, int n_feat) {
  Ws w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
  const size_t Bp = pad128(B);
  const int ng = gen_pad(n_out);
  w.x = take(Bp * kD * 2);
  w.qkv = take(Bp * 3 * kD * 2);
  w.att = take(Bp * kD * 2);
  w.t0 = take(Bp * kD * 2);
  w.r1s = take(Bp * kD * 2);
  w.r1 = take(Bp * kD * 2);
  w.q2 = take(Bp * kD * 2);
  w.a2 = take(Bp * kD * 2);
  w.r2s = take(Bp * kD * 2);
  w.r2 = take(Bp * kD * 2);
  w.hdn = take(Bp * kFF * 2);
  w.r3s = take(Bp * kD * 2);
  w.r3 = take(Bp * kD * 2);
  w.logits = take(Bp * (size_t)ng * 4);
  w.gen_n = take(Bp * (size_t)ng * 2);
  w.sim = take(Bp * (size_t)pad128(n_feat) * 4);
  w.inv = take(Bp * 4);
  w.kc = take((size_t)B * T_max * kD * 2);
  w.vc = take((size_t)B * T_max * kD * 2);
  w.kv2 = take((size_t)pad128((long)B * n_tok) * 2 * kD * 2);
  w.total = o;
  return w;
}

__device__ __forceinline__ float pe_value(int pos, int cc) {  // PositionalEncoding, SLD/model/transformer.py:168-186
  const float div = expf((float)(cc & ~1) * (-logf(10000.f) / (float)kE));
  const float ang = (float)pos * div;
  return (cc & 1) ? cosf(ang) : sinf(ang);
}
// x[b] = [ lut[tok_b] * sqrt(E) | pe[pos] ]   (Embeddings :277-286, torch.cat with the positional half :346-348)
__device__ __forceinline__ void embed_row(const float* __restrict__ lut, long tok, int pos, bf16* __restrict__ xrow, int tid,
                                          int nthreads) {
  const float sc = sqrtf((float)kE);
  for (int c = tid; c < kD; c += nthreads)
    xrow[c] = __float2bfloat16(c < kE ? lut[tok * kE + c] * sc : pe_value(pos, c - kE));
}
__global__ void __launch_bounds__(256) decode_embed_kernel(const float* __restrict__ lut, long long* __restrict__ pred, int ld_pred,
                                                           bf16* __restrict__ x) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) pred[(long)b * ld_pred] = 0;  // the start symbol
  embed_row(lut, 0, 0, x + (long)b * kD, threadIdx.x, 256);
}

// one new position against the cache: CTA per (sample, head), thread d = one of the 256 head dimensions.
// qkv: this step's [q | k | v] row per sample (bf16, 3072 columns); the k, v thirds are appended to the cache at position t.
__global__ void __launch_bounds__(kDk) decode_self_attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kc, bf16* __restrict__ vc,
                                                               int T_cap, int t, bf16* __restrict__ out) {
  __shared__ float sq[kDk];
  __shared__ float sc[64];
  const int b = blockIdx.x / kH, h = blockIdx.x % kH, d = threadIdx.x, warp = d >> 5, lane = d & 31;
  const bf16* row = qkv + (long)b * 3 * kD + h * kDk;
  sq[d] = __bfloat162float(row[d]);
  bf16* kb = kc + ((long)b * T_cap) * kD + h * kDk;
  bf16* vb = vc + ((long)b * T_cap) * kD + h * kDk;
  kb[(long)t * kD + d] = row[kD + d];
  vb[(long)t * kD + d] = row[2 * kD + d];
  __syncthreads();
  for (int s = warp; s <= t; s += kDk / 32) {
    const bf16* kr = kb + (long)s * kD + lane * 8;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(sq[lane * 8 + j], __bfloat162float(kr[j]), acc);
    acc = warp_sum(acc);
    if (lane == 0) sc[s] = acc * 0.0625f;  // 1 / sqrt(256)
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int s = 0; s <= t; ++s) mx = fmaxf(mx, sc[s]);
  float den = 0.f, acc = 0.f;
  for (int s = 0; s <= t; ++s) {
    const float p = __expf(sc[s] - mx);
    den += p;
    acc = fmaf(p, __bfloat162float(vb[(long)s * kD + d]), acc);
  }
  out[(long)b * kD + h * kDk + d] = __float2bfloat16(acc / den);
}

// cross-attention of the new position to the image tokens: kv2 rows (b * n_tok + k) = [K (1024) | V (1024)] bf16
__global__ void __launch_bounds__(kDk) decode_cross_attn_kernel(const bf16* __restrict__ q2, const bf16* __restrict__ kv2, int n_tok,
                                                                bf16* __restrict__ out) {
  extern __shared__ float sc[];  // n_tok scores
  __shared__ float sq[kDk];
  __shared__ float red[kDk / 32];
  const int b = blockIdx.x / kH, h = blockIdx.x % kH, d = threadIdx.x, warp = d >> 5, lane = d & 31;
  sq[d] = __bfloat162float(q2[(long)b * kD + h * kDk + d]);
  const bf16* base = kv2 + ((long)b * n_tok) * 2 * kD + h * kDk;
  __syncthreads();
  for (int s = warp; s < n_tok; s += kDk / 32) {
    const bf16* kr = base + (long)s * 2 * kD + lane * 8;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(sq[lane * 8 + j], __bfloat162float(kr[j]), acc);
    acc = warp_sum(acc);
    if (lane == 0) sc[s] = acc * 0.0625f;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int s = d; s < n_tok; s += kDk) mx = fmaxf(mx, sc[s]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kDk / 32; ++i) mx = fmaxf(mx, red[i]);
  float den = 0.f, acc = 0.f;
  for (int s = 0; s < n_tok; ++s) {  // every thread walks all tokens: the scores are broadcast reads, V rows are coalesced over d
    const float p = __expf(sc[s] - mx);
    den += p;
    acc = fmaf(p, __bfloat162float(base[(long)s * 2 * kD + kD + d]), acc);
  }
  out[(long)b * kD + h * kDk + d] = __float2bfloat16(acc / den);
}

// rows of fp32 (ld columns, n valid) -> L2-normalised bf16 rows (image-ids-CTR/train.py:128: prediction / prediction.norm)
__global__ void __launch_bounds__(256) decode_l2norm_kernel(const float* __restrict__ x, int ld, int n, bf16* __restrict__ y) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* r = x + (long)b * ld;
  float s = 0.f;
  for (int c = threadIdx.x; c < n; c += 256) s = fmaf(r[c], r[c], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < 8; ++i) t += red[i];
  const float inv = rsqrtf(t);
  for (int c = threadIdx.x; c < ld; c += 256) y[(long)b * ld + c] = __float2bfloat16(c < n ? r[c] * inv : 0.f);
}

// arg-max (lowest index on ties, torch.max semantics) and its softmax probability over the first n_class columns of this sample's
// score row; writes pred[b, pos + 1], prob[b, pos] and the embedding of the chosen token at position pos + 1 for the next step
__global__ void __launch_bounds__(256) decode_pick_kernel(const float* __restrict__ scores, int ld, int n_class,
                                                          const float* __restrict__ lut, int vocab, long long* __restrict__ pred,
                                                          int ld_pred, float* __restrict__ prob, int ld_prob, int pos,
                                                          bf16* __restrict__ x) {
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ float rs[8];
  __shared__ int s_tok;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* r = scores + (long)b * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < n_class; c += 256) {
    const float v = r[c];
    if (v > best) {
      best = v;
      bi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    rv[warp] = best;
    ri[warp] = bi;
  }
  __syncthreads();
  best = rv[0];
  bi = ri[0];
#pragma unroll
  for (int i = 1; i < 8; ++i)
    if (rv[i] > best || (rv[i] == best && ri[i] < bi)) {
      best = rv[i];
      bi = ri[i];
    }
  float den = 0.f;
  for (int c = threadIdx.x; c < n_class; c += 256) den += __expf(r[c] - best);
  den = warp_sum(den);
  if (lane == 0) rs[warp] = den;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += rs[i];
    pred[(long)b * ld_pred + pos + 1] = bi;
    prob[(long)b * ld_prob + pos] = 1.f / t;
    s_tok = bi < vocab ? bi : vocab - 1;  // (a class index beyond the embedding table would raise in nn.Embedding)
  }
  __syncthreads();
  embed_row(lut, s_tok, pos + 1, x + (long)b * kD, threadIdx.x, 256);
}

__global__ void pad_rows_bf16_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst, int rows_pad, int ld) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows_pad * ld) return;
  const int r = (int)(i / ld), c = (int)(i - (long)r * ld);
  dst[i] = __float2bfloat16(r < rows && c < cols ? src[(long)r * cols + c] : 0.f);
}

int lin(const bf16* x, const bf16* w, const float* bias, void* y, int Mp, int K, int N, int relu, int f32, cudaStream_t s) {
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_total = N;
  p.kh = p.kw = 1;
  p.W = 64;
  p.H = 2;
  p.epi = f32 ? TC_EPI_F32 : TC_EPI_BF16;
  p.relu = relu;
  p.ldc = N;
  p.bias = bias;
  p.out = y;
  const bf16* ap[1] = {x};
  return tc_gemm_launch(ap, 1, K, (long)64 * K, (long)128 * K, K, Mp / 128, w, K, p, s);
}

#define TRY(e)             \
  do {                     \
    int _rc = (e);         \
    if (_rc) return _rc;   \
  } while (0)

}  // namespace

// params: HOST array of 29 DEVICE fp32 pointers in this order -
//   0 embedding table (vocab, 512);
//   1-8  masked self-attention linears q, k, v, out: weight (1024,1024), bias (1024) each;   9, 10 LayerNorm 1 scale, shift;
//   11-18 cross-attention linears q, k, v, out;                                              19, 20 LayerNorm 2;
//   21-24 FFN w_1 (2048,1024), b_1, w_2 (1024,2048), b_2;                                    25, 26 LayerNorm 3;
//   27, 28 generator weight (n_out, 1024), bias (n_out).
// text_features: NULL (stroke-level-decomposition: the generator's n_out scores ARE the class scores) or fp32 (n_feat, n_out)
// (image-ids-CTR: the generator output is L2-normalised and matched against these, train.py:127-130).
extern "C" size_t focr_recog_decode_prepared_bytes(int vocab, int n_out, int n_feat) { return blob_layout(vocab, n_out, n_feat).total; }

extern "C" int focr_recog_decode_prepare(void* const* params, int vocab, int n_out, const float* text_features, int n_feat,
                                         void* blob, size_t blob_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(params && blob && vocab >= 1 && n_out >= 1 && n_out <= 4096, "recog_decode_prepare: bad arguments");
  FOCR_REQUIRE((text_features == nullptr) == (n_feat == 0), "recog_decode_prepare: text_features / n_feat mismatch");
  const Blob L = blob_layout(vocab, n_out, n_feat);
  FOCR_REQUIRE(blob_bytes >= L.total, "recog_decode_prepare: blob of %zu bytes, need %zu", blob_bytes, L.total);
  char* base = (char*)blob;
  auto W = [&](size_t off) { return (bf16*)(base + off); };
  auto F = [&](size_t off) { return (float*)(base + off); };
  auto P = [&](int i) { return (const float*)params[i]; };
  auto copyf = [&](size_t off, int i, size_t n) { return cudaMemcpyAsync(base + off, params[i], n * 4, cudaMemcpyDeviceToDevice, s); };
  for (int j = 0; j < 3; ++j) {  // [q ; k ; v] rows of the masked self-attention
    TRY(prep_linear_w(P(1 + 2 * j), W(L.wqkv) + (size_t)j * kD * kD, nullptr, kD, kD, 0, 0, s));
    FOCR_CHECK_CUDA(cudaMemcpyAsync(base + L.bqkv + (size_t)j * kD * 4, params[2 + 2 * j], kD * 4, cudaMemcpyDeviceToDevice, s));
  }
  TRY(prep_linear_w(P(7), W(L.wo), nullptr, kD, kD, 0, 0, s));
  FOCR_CHECK_CUDA(copyf(L.bo, 8, kD));
  TRY(prep_linear_w(P(11), W(L.wq2), nullptr, kD, kD, 0, 0, s));
  FOCR_CHECK_CUDA(copyf(L.bq2, 12, kD));
  for (int j = 0; j < 2; ++j) {  // [k ; v] rows of the cross-attention
    TRY(prep_linear_w(P(13 + 2 * j), W(L.wkv2) + (size_t)j * kD * kD, nullptr, kD, kD, 0, 0, s));
    FOCR_CHECK_CUDA(cudaMemcpyAsync(base + L.bkv2 + (size_t)j * kD * 4, params[14 + 2 * j], kD * 4, cudaMemcpyDeviceToDevice, s));
  }
  TRY(prep_linear_w(P(17), W(L.wo2), nullptr, kD, kD, 0, 0, s));
  FOCR_CHECK_CUDA(copyf(L.bo2, 18, kD));
  TRY(prep_linear_w(P(21), W(L.w1), nullptr, kFF, kD, 0, 0, s));
  FOCR_CHECK_CUDA(copyf(L.b1, 22, kFF));
  TRY(prep_linear_w(P(23), W(L.w2), nullptr, kD, kFF, 0, 0, s));
  FOCR_CHECK_CUDA(copyf(L.b2, 24, kD));
  const int ng = gen_pad(n_out);
  pad_rows_bf16_kernel<<<focr_cdiv((long)ng * kD, 256), 256, 0, s>>>(P(27), n_out, kD, W(L.wg), ng, kD);
  FOCR_LAUNCH_CHECK();
  FOCR_CHECK_CUDA(cudaMemsetAsync(base + L.bg, 0, (size_t)ng * 4, s));
  FOCR_CHECK_CUDA(copyf(L.bg, 28, n_out));
  const int lnp[6] = {9, 10, 19, 20, 25, 26};
  for (int i = 0; i < 6; ++i) FOCR_CHECK_CUDA(copyf(L.ln[i], lnp[i], kD));
  FOCR_CHECK_CUDA(copyf(L.lut, 0, (size_t)vocab * kE));
  if (n_feat) {
    const int nf = pad128(n_feat);
    pad_rows_bf16_kernel<<<focr_cdiv((long)nf * ng, 256), 256, 0, s>>>(text_features, n_feat, n_out, W(L.feat), nf, ng);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

extern "C" size_t focr_recog_decode_workspace_bytes(int B, int n_tok, int T_max, int n_out, int n_feat) {
  return ws_layout(B, n_tok, T_max, n_out, n_feat).total;
}

// feat: encoder output, bf16 (B * n_tok rows, padded with zero rows to a multiple of 128; 1024 columns): the NHWC map the
// recogniser's encoder leaves.  Runs T_max decode steps for every
// sample, as the reference loops do (they cut at the end symbol afterwards, on the host).
//   pred int64 (B, T_max + 1): column 0 = start symbol 0, column i + 1 = arg-max of step i;
//   prob fp32 (B, T_max):      winning softmax probability of each step.
extern "C" int focr_recog_decode(const void* blob, size_t blob_bytes, int vocab, int n_out, int n_feat, const void* feat, int B,
                                 int n_tok, int T_max, long long* pred, float* prob, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(blob && feat && pred && prob && ws, "recog_decode: null pointer");
  FOCR_REQUIRE(B >= 1 && n_tok >= 1 && T_max >= 1 && T_max <= 63, "recog_decode: B=%d n_tok=%d T_max=%d (T_max <= 63)", B, n_tok, T_max);
  FOCR_REQUIRE((size_t)n_tok * 4 <= 160 * 1024, "recog_decode: %d image tokens exceed the score buffer", n_tok);
  const Blob L = blob_layout(vocab, n_out, n_feat);
  const Ws Wl = ws_layout(B, n_tok, T_max, n_out, n_feat);
  FOCR_REQUIRE(blob_bytes >= L.total && ws_bytes >= Wl.total, "recog_decode: blob / workspace too small");
  ProfScope _ps("recog_decode", s);
  const char* bb = (const char*)blob;
  char* wb = (char*)ws;
  auto W = [&](size_t off) { return (const bf16*)(bb + off); };
  auto F = [&](size_t off) { return (const float*)(bb + off); };
  auto A = [&](size_t off) { return (bf16*)(wb + off); };
  const int Bp = pad128(B), ng = gen_pad(n_out), nf = pad128(n_feat);
  static bool init = false;
  if (!init) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(decode_cross_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    init = true;
  }
  // padding rows of the GEMM inputs must stay finite: clear the activations once
  FOCR_CHECK_CUDA(cudaMemsetAsync(wb, 0, Wl.kc, s));
  // cross-attention K, V of every image token, once
  const int rows2 = pad128((long)B * n_tok);  // `feat` holds this many rows (the caller zero-pads the last GEMM tile)
  TRY(lin((const bf16*)feat, W(L.wkv2), F(L.bkv2), A(Wl.kv2), rows2, kD, 2 * kD, 0, 0, s));
  decode_embed_kernel<<<B, 256, 0, s>>>(F(L.lut), pred, T_max + 1, A(Wl.x));
  FOCR_LAUNCH_CHECK();
  for (int i = 0; i < T_max; ++i) {
    TRY(lin(A(Wl.x), W(L.wqkv), F(L.bqkv), A(Wl.qkv), Bp, kD, 3 * kD, 0, 0, s));
    decode_self_attn_kernel<<<B * kH, kDk, 0, s>>>(A(Wl.qkv), A(Wl.kc), A(Wl.vc), T_max, i, A(Wl.att));
    FOCR_LAUNCH_CHECK();
    TRY(lin(A(Wl.att), W(L.wo), F(L.bo), A(Wl.t0), Bp, kD, kD, 0, 0, s));
    TRY(focr_layernorm_wide_fwd(A(Wl.t0), A(Wl.x), F(L.ln[0]), F(L.ln[1]), A(Wl.r1s), A(Wl.r1), Bp, kD, 1e-6f, stream));
    TRY(lin(A(Wl.r1), W(L.wq2), F(L.bq2), A(Wl.q2), Bp, kD, kD, 0, 0, s));
    decode_cross_attn_kernel<<<B * kH, kDk, (size_t)n_tok * 4, s>>>(A(Wl.q2), A(Wl.kv2), n_tok, A(Wl.a2));
    FOCR_LAUNCH_CHECK();
    TRY(lin(A(Wl.a2), W(L.wo2), F(L.bo2), A(Wl.t0), Bp, kD, kD, 0, 0, s));
    TRY(focr_layernorm_wide_fwd(A(Wl.t0), A(Wl.r1), F(L.ln[2]), F(L.ln[3]), A(Wl.r2s), A(Wl.r2), Bp, kD, 1e-6f, stream));
    TRY(lin(A(Wl.r2), W(L.w1), F(L.b1), A(Wl.hdn), Bp, kD, kFF, 1, 0, s));
    TRY(lin(A(Wl.hdn), W(L.w2), F(L.b2), A(Wl.t0), Bp, kFF, kD, 0, 0, s));
    TRY(focr_layernorm_wide_fwd(A(Wl.t0), A(Wl.r2), F(L.ln[4]), F(L.ln[5]), A(Wl.r3s), A(Wl.r3), Bp, kD, 1e-6f, stream));
    TRY(lin(A(Wl.r3), W(L.wg), F(L.bg), wb + Wl.logits, Bp, kD, ng, 0, 1, s));
    const float* scores = (const float*)(wb + Wl.logits);
    int ld = ng, ncls = n_out;
    if (n_feat) {
      decode_l2norm_kernel<<<B, 256, 0, s>>>(scores, ng, n_out, A(Wl.gen_n));
      FOCR_LAUNCH_CHECK();
      TRY(lin(A(Wl.gen_n), W(L.feat), nullptr, wb + Wl.sim, Bp, ng, nf, 0, 1, s));
      scores = (const float*)(wb + Wl.sim);
      ld = nf;
      ncls = n_feat;
    }
    decode_pick_kernel<<<B, 256, 0, s>>>(scores, ld, ncls, F(L.lut), vocab, pred, T_max + 1, prob, T_max, i, A(Wl.x));
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}
