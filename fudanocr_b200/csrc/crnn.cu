// CRNN evaluator + greedy CTC decode (scene-text-telescope/model/crnn/crnn.py:25-80, interfaces/base.py:319-325
// parse_crnn_data, interfaces/super_resolution.py:143-158 get_crnn_pred, utils/utils_crnn.py:54-89 decode).
// Forward only (the reference keeps this recogniser frozen, base.py:309-317).  Convolutions and linears run on
// the tcgen05 GEMM engine through explicit im2col (spatial sizes 32x100 .. 1x26 are not TMA-tile friendly and
// the whole net is 1.4 GFLOP/image); eval-mode BatchNorm is applied by the bn_apply kernel; the two BiLSTMs are
// an input-projection GEMM over all time steps plus, per step, a recurrent GEMM and a fused gate kernel.
#include <string.h>

#include "kernels.cuh"

#define TRY(expr)             \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

namespace {

// ---- bicubic (A = -0.75, align_corners = False) resize W 128 -> 100 (H 32 -> 32 is the identity) + gray ---------
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void bicubic_gray_kernel(const float* __restrict__ img, float* __restrict__ gray, int B, int Hin, int Win,
                                    int Hout, int Wout) {
  const long n = (long)B * Hout * Wout;
  const float A = -0.75f;
  const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wout), y = (int)((i / Wout) % Hout), b = (int)(i / ((long)Wout * Hout));
    const float ry = ((float)y + 0.5f) * sh - 0.5f, rx = ((float)x + 0.5f) * sw - 0.5f;
    const float fy = floorf(ry), fx = floorf(rx);
    const int iy = (int)fy, ix = (int)fx;
    const float ty = ry - fy, tx = rx - fx;
    float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
    float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
    float ch[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = img + ((long)b * 3 + c) * Hin * Win;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yy = min(max(iy - 1 + j, 0), Hin - 1);
        float row = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int xx = min(max(ix - 1 + k, 0), Win - 1);
          row += wx[k] * p[(long)yy * Win + xx];
        }
        acc += wy[j] * row;
      }
      ch[c] = acc;
    }
    gray[i] = 0.299f * ch[0] + 0.587f * ch[1] + 0.114f * ch[2];
  }
}

// ---- generic im2col: col[row][tap*C + c] = x[b, ho + ty - ph, wo + tx - pw, c]; row = (b,ho,wo) or (wo,b) -------
__global__ void im2col_generic_kernel(const bf16* __restrict__ x, const float* __restrict__ x_f32_c1, bf16* __restrict__ col,
                                      int B, int H, int W, int C, int kh, int kw, int ph, int pw, int Ho, int Wo,
                                      int Kpad, int tb_order) {
  const long n = (long)B * Ho * Wo * (Kpad / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % (Kpad / 8));
    const long pix = i / (Kpad / 8);
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    const long row = tb_order ? ((long)wo * B + b) : pix;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (x != nullptr) {
      const int k0 = kc * 8;
      if (k0 < kh * kw * C) {
        const int tap = k0 / C, c0 = k0 - tap * C;
        const int hh = ho + tap / kw - ph, ww = wo + tap % kw - pw;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
          u = *reinterpret_cast<const uint4*>(x + (((long)b * H + hh) * W + ww) * C + c0);
      }
    } else {  // single-channel fp32 image
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kc * 8 + j;
        float val = 0.f;
        if (k < kh * kw) {
          const int hh = ho + k / kw - ph, ww = wo + k % kw - pw;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = x_f32_c1[((long)b * H + hh) * W + ww];
        }
        v[j] = val;
      }
      u.x = pack_bf16x2(v[0], v[1]);
      u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]);
      u.w = pack_bf16x2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(col + row * Kpad + kc * 8) = u;
  }
}

// conv weight fp32 [Co][C][kh][kw] -> bf16 [Co][Kpad] with k = tap*C + c
__global__ void prep_conv_w_generic_kernel(const float* __restrict__ w, bf16* __restrict__ o, int Co, int C, int taps,
                                           int Kpad) {
  const long n = (long)Co * Kpad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad), co = (int)(i / Kpad);
    float v = 0.f;
    if (k < taps * C) {
      const int tap = k / C, c = k - tap * C;
      v = w[((long)co * C + c) * taps + tap];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}

// nn.MaxPool2d((2,2),(2,1),(0,1)) on NHWC: Ho = H/2, Wo = W+1; window rows 2ho..2ho+1, cols wo-1..wo
__global__ void maxpool_s21_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W + 1;
  const long n = (long)B * Ho * Wo * (C / 8);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % (C / 8));
    const long pix = i / (C / 8);
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = -1; dx <= 0; ++dx) {
        const int ww = wo + dx;
        if (ww < 0 || ww >= W) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((long)b * H + 2 * ho + dy) * W + ww) * C + cc * 8);
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

// LSTM cell (gate order i, f, g, o as nn.LSTM): pre = gin[t*B+b][d*4H + ...] + rec[b][...]
__global__ void lstm_gate_kernel(const float* __restrict__ gin, long gin_ld, const float* __restrict__ rec, int B, int H,
                                 int t, int dir, float* __restrict__ c_state, bf16* __restrict__ h_cur,
                                 bf16* __restrict__ seq_out, long seq_ld, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i - b * H;
  const float* g = gin + ((long)t * B + b) * gin_ld + (long)dir * 4 * H;
  const float* r = rec + (long)b * 4 * H;
  float pi = g[j], pf = g[H + j], pg = g[2 * H + j], po = g[3 * H + j];
  if (!first) {
    pi += r[j];
    pf += r[H + j];
    pg += r[2 * H + j];
    po += r[3 * H + j];
  }
  const float ig = 1.f / (1.f + __expf(-pi)), fg = 1.f / (1.f + __expf(-pf)), og = 1.f / (1.f + __expf(-po));
  const float c = (first ? 0.f : fg * c_state[i]) + ig * tanhf(pg);
  c_state[i] = c;
  const bf16 h = __float2bfloat16_rn(og * tanhf(c));
  h_cur[i] = h;
  seq_out[((long)t * B + b) * seq_ld + (long)dir * H + j] = h;
}

__global__ void add_bias2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

// Greedy CTC: argmax over classes (lowest index on ties), drop repeats then blanks (index 0).
// logits (T, B, ld) fp32; path (B,T) int32; out (B,T) int32 padded with -1; len (B)
__global__ void ctc_greedy_kernel(const float* __restrict__ logits, long ld, int T, int B, int C, int* __restrict__ path,
                                  int* __restrict__ out, int* __restrict__ len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int prev = -1, n = 0;
  for (int t = 0; t < T; ++t) {
    const float* p = logits + ((long)t * B + b) * ld;
    int best = 0;
    float bv = p[0];
    for (int c = 1; c < C; ++c) {
      const float v = p[c];
      if (v > bv) {
        bv = v;
        best = c;
      }
    }
    path[b * T + t] = best;
    if (best != 0 && best != prev) out[b * T + n++] = best;
    prev = best;
  }
  len[b] = n;
  for (int t = n; t < T; ++t) out[b * T + t] = -1;
}

int cgrid(long n, int per) {
  long g = (n + per - 1) / per;
  if (g > 148L * 8) g = 148L * 8;
  if (g < 1) g = 1;
  return (int)g;
}
inline long pad128(long m) { return (m + 127) / 128 * 128; }

struct ConvSpec {
  int cin, cout, h, w, k, pad, ho, wo, kpad, bn /*slot base or -1*/, pool /*0 none,1 2x2,2 s21*/;
};
// slot indices follow the reference state_dict order (49 entries)
enum : int {
  C0W = 0, C0B, C1W, C1B, C2W, C2B, BN2W, BN2B, BN2RM, BN2RV, BN2NBT, C3W, C3B, C4W, C4B, BN4W, BN4B, BN4RM, BN4RV,
  BN4NBT, C5W, C5B, C6W, C6B, BN6W, BN6B, BN6RM, BN6RV, BN6NBT,
  R0_WIH, R0_WHH, R0_BIH, R0_BHH, R0_WIH_R, R0_WHH_R, R0_BIH_R, R0_BHH_R, R0_EW, R0_EB,
  R1_WIH, R1_WHH, R1_BIH, R1_BHH, R1_WIH_R, R1_WHH_R, R1_BIH_R, R1_BHH_R, R1_EW, R1_EB, CRNN_SLOTS
};
const ConvSpec kConv[7] = {
    {1, 64, 32, 100, 3, 1, 32, 100, 64, -1, 1},    {64, 128, 16, 50, 3, 1, 16, 50, 576, -1, 1},
    {128, 256, 8, 25, 3, 1, 8, 25, 1152, BN2W, 0}, {256, 256, 8, 25, 3, 1, 8, 25, 2304, -1, 2},
    {256, 512, 4, 26, 3, 1, 4, 26, 2304, BN4W, 0}, {512, 512, 4, 26, 3, 1, 4, 26, 4608, -1, 2},
    {512, 512, 2, 27, 2, 0, 1, 26, 2048, BN6W, 0},
};
const int kConvW[7] = {C0W, C1W, C2W, C3W, C4W, C5W, C6W};

struct Bump {
  char* base;
  size_t off = 0;
  template <typename T>
  T* get(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct CrnnWs {
  float* gray;
  bf16 *col[7], *out[7], *act[7], *pool[7], *wconv[7];
  float* stats;
  // lstm
  bf16 *wih[2], *whh[2][2], *wemb[2];
  float *bsum[2], *embb1;
  float* gin;       // (TBp, 2048) fp32
  float* rec;       // (Bp, 1024) fp32 x 2 directions
  float* cst;       // (B,256) x 2
  bf16* hcur;       // (Bp,256) x 2
  bf16* seq;        // (TBp, 512)
  bf16* emb0;       // (TBp, 256)
  float* logits64;  // (TBp, 64)
  size_t total;
};

void crnn_layout(CrnnWs& w, int B, void* base) {
  Bump b{reinterpret_cast<char*>(base)};
  w.gray = b.get<float>((size_t)B * 3200);
  for (int i = 0; i < 7; ++i) {
    const ConvSpec& c = kConv[i];
    const long Mp = pad128((long)B * c.ho * c.wo);
    w.col[i] = b.get<bf16>(Mp * c.kpad);
    w.out[i] = b.get<bf16>(Mp * c.cout);
    w.act[i] = b.get<bf16>(Mp * c.cout);
    w.pool[i] = b.get<bf16>(Mp * c.cout);
    w.wconv[i] = b.get<bf16>((long)c.cout * c.kpad);
  }
  w.stats = b.get<float>(4 * 512);
  const long TBp = pad128(26L * B), Bp = pad128(B);
  w.wih[0] = b.get<bf16>(2048 * 512);
  w.wih[1] = b.get<bf16>(2048 * 256);
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d) w.whh[l][d] = b.get<bf16>(1024 * 256);
  w.wemb[0] = b.get<bf16>(256 * 512);
  w.wemb[1] = b.get<bf16>(64 * 512);
  w.bsum[0] = b.get<float>(2048);
  w.bsum[1] = b.get<float>(2048);
  w.embb1 = b.get<float>(64);
  w.gin = b.get<float>(TBp * 2048);
  w.rec = b.get<float>(2 * Bp * 1024);
  w.cst = b.get<float>(2L * B * 256);
  w.hcur = b.get<bf16>(2 * Bp * 256);
  w.seq = b.get<bf16>(TBp * 512);
  w.emb0 = b.get<bf16>(TBp * 256);
  w.logits64 = b.get<float>(TBp * 64);
  w.total = (b.off + 255) & ~(size_t)255;
}

TcGemmParams gp() {
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.kh = p.kw = 1;
  p.epi = TC_EPI_BF16;
  return p;
}
int tok_gemm(const bf16* a, int K, long M, const bf16* w, int N, TcGemmParams p, cudaStream_t s) {
  p.n_total = N;
  p.W = 64;
  p.H = 2;
  if (p.ldc == 0) p.ldc = N;
  const bf16* ap[1] = {a};
  return tc_gemm_launch(ap, 1, K, (long)64 * K, (long)128 * K, K, (int)(M / 128), w, K, p, s);
}
template <typename T>
T* P(void* const* prm, int i) {
  return reinterpret_cast<T*>(prm[i]);
}

int lstm_layer(void* const* prm, int base_slot, int layer, const bf16* x, int nin, int B, CrnnWs& w, cudaStream_t s) {
  const int T = 26, H = 256;
  const long TBp = pad128((long)T * B), Bp = pad128(B);
  // weights: [fwd; reverse] input projections stacked to N = 2048, summed biases
  for (int d = 0; d < 2; ++d) {
    TRY(prep_linear_w(P<float>(prm, base_slot + 4 * d), w.wih[layer] + (long)d * 1024 * nin, nullptr, 1024, nin, 0, 0, s));
    TRY(prep_linear_w(P<float>(prm, base_slot + 4 * d + 1), w.whh[layer][d], nullptr, 1024, 256, 0, 0, s));
    add_bias2_kernel<<<8, 128, 0, s>>>(P<float>(prm, base_slot + 4 * d + 2), P<float>(prm, base_slot + 4 * d + 3),
                                       w.bsum[layer] + d * 1024, 1024);
    FOCR_LAUNCH_CHECK();
  }
  TcGemmParams p = gp();
  p.bias = w.bsum[layer];
  p.out = w.gin;
  p.epi = TC_EPI_F32;
  TRY(tok_gemm(x, nin, TBp, w.wih[layer], 2048, p, s));
  for (int step = 0; step < T; ++step) {
    for (int d = 0; d < 2; ++d) {
      const int t = d == 0 ? step : T - 1 - step;
      float* rec = w.rec + (long)d * Bp * 1024;
      bf16* hc = w.hcur + (long)d * Bp * 256;
      if (step > 0) {
        TcGemmParams q = gp();
        q.out = rec;
        q.epi = TC_EPI_F32;
        TRY(tok_gemm(hc, 256, Bp, w.whh[layer][d], 1024, q, s));
      }
      lstm_gate_kernel<<<focr_cdiv((long)B * H, 256), 256, 0, s>>>(w.gin, 2048, rec, B, H, t, d, w.cst + (long)d * B * H,
                                                                   hc, w.seq, 512, step == 0);
      FOCR_LAUNCH_CHECK();
    }
  }
  return FOCR_OK;
}

}  // namespace

extern "C" {

int focr_crnn_num_slots(void) { return CRNN_SLOTS; }

size_t focr_crnn_workspace_bytes(int B) {
  CrnnWs w;
  crnn_layout(w, B, nullptr);
  return w.total + 256;
}

// parse_crnn_data: images (B,3,32,128) fp32 NCHW -> gray (B,1,32,100) fp32
int focr_bicubic_gray_32x100(const float* images, float* gray, int B, void* stream) {
  bicubic_gray_kernel<<<cgrid((long)B * 3200, 256), 256, 0, (cudaStream_t)stream>>>(images, gray, B, 32, 128, 32, 100);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// CRNN(32,1,37,256).forward, eval mode.  params: HOST array of the 49 state_dict tensors (device pointers) in
// state_dict order.  images: (B,3,32,128) fp32 (parse_crnn_data applied inside) or, with input_is_gray, the
// (B,1,32,100) gray tensor CRNN.forward receives in the reference.  logits out (26,B,37) fp32.
int focr_crnn_forward(void* const* params, const float* images, int input_is_gray, float* logits, int B, void* ws_,
                      size_t ws_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  CrnnWs w;
  char* base = (char*)(((uintptr_t)ws_ + 255) & ~(uintptr_t)255);
  crnn_layout(w, B, base);
  FOCR_REQUIRE(ws_bytes >= w.total + 256, "crnn_forward: workspace too small");
  if (input_is_gray) {  // CRNN.forward proper: the caller already ran parse_crnn_data -> (B,1,32,100)
    FOCR_CHECK_CUDA(cudaMemcpyAsync(w.gray, images, (size_t)B * 3200 * 4, cudaMemcpyDeviceToDevice, s));
  } else {
    bicubic_gray_kernel<<<cgrid((long)B * 3200, 256), 256, 0, s>>>(images, w.gray, B, 32, 128, 32, 100);
    FOCR_LAUNCH_CHECK();
  }
  const bf16* x = nullptr;
  for (int i = 0; i < 7; ++i) {
    const ConvSpec& c = kConv[i];
    const long M = (long)B * c.ho * c.wo, Mp = pad128(M);
    const int wslot = kConvW[i];
    prep_conv_w_generic_kernel<<<cgrid((long)c.cout * c.kpad, 256), 256, 0, s>>>(P<float>(params, wslot), w.wconv[i],
                                                                                c.cout, c.cin, c.k * c.k, c.kpad);
    FOCR_LAUNCH_CHECK();
    im2col_generic_kernel<<<cgrid(M * (c.kpad / 8), 256), 256, 0, s>>>(i == 0 ? nullptr : x, i == 0 ? w.gray : nullptr,
                                                                      w.col[i], B, c.h, c.w, c.cin, c.k, c.k, c.pad,
                                                                      c.pad, c.ho, c.wo, c.kpad, i == 6 ? 1 : 0);
    FOCR_LAUNCH_CHECK();
    TcGemmParams p = gp();
    p.bias = P<float>(params, wslot + 1);
    p.relu = c.bn < 0 ? 1 : 0;
    p.out = c.bn < 0 ? w.act[i] : w.out[i];
    TRY(tok_gemm(w.col[i], c.kpad, Mp, w.wconv[i], c.cout, p, s));
    if (c.bn >= 0) {
      TRY(bn_eval_stats(P<float>(params, c.bn), P<float>(params, c.bn + 1), P<float>(params, c.bn + 2),
                        P<float>(params, c.bn + 3), 1e-5f, c.cout, w.stats, s));
      TRY(bn_apply(w.out[i], c.cout, w.stats, w.act[i], c.cout, M, c.cout, ACT_RELU, nullptr, 0, nullptr, s));
    }
    x = w.act[i];
    if (c.pool == 1) {
      TRY(maxpool_fwd(w.act[i], w.pool[i], B, c.ho, c.wo, c.cout, 2, s));
      x = w.pool[i];
    } else if (c.pool == 2) {
      maxpool_s21_kernel<<<cgrid((long)B * (c.ho / 2) * (c.wo + 1) * (c.cout / 8), 256), 256, 0, s>>>(
          w.act[i], w.pool[i], B, c.ho, c.wo, c.cout);
      FOCR_LAUNCH_CHECK();
      x = w.pool[i];
    }
  }
  // x = act[6]: rows (t, b), 512 channels  (crnn.py:74-75: squeeze(2).permute(2,0,1))
  const long TBp = pad128(26L * B);
  TRY(lstm_layer(params, R0_WIH, 0, x, 512, B, w, s));
  TRY(prep_linear_w(P<float>(params, R0_EW), w.wemb[0], nullptr, 256, 512, 0, 0, s));
  {
    TcGemmParams p = gp();
    p.bias = P<float>(params, R0_EB);
    p.out = w.emb0;
    TRY(tok_gemm(w.seq, 512, TBp, w.wemb[0], 256, p, s));
  }
  TRY(lstm_layer(params, R1_WIH, 1, w.emb0, 256, B, w, s));
  FOCR_CHECK_CUDA(cudaMemsetAsync(w.wemb[1], 0, 64 * 512 * 2, s));
  FOCR_CHECK_CUDA(cudaMemsetAsync(w.embb1, 0, 64 * 4, s));
  TRY(prep_linear_w(P<float>(params, R1_EW), w.wemb[1], nullptr, 37, 512, 0, 0, s));
  FOCR_CHECK_CUDA(cudaMemcpyAsync(w.embb1, params[R1_EB], 37 * 4, cudaMemcpyDeviceToDevice, s));
  {
    TcGemmParams p = gp();
    p.bias = w.embb1;
    p.out = w.logits64;
    p.epi = TC_EPI_F32;
    TRY(tok_gemm(w.seq, 512, TBp, w.wemb[1], 64, p, s));
  }
  FOCR_CHECK_CUDA(cudaMemcpy2DAsync(logits, 37 * 4, w.logits64, 64 * 4, 37 * 4, 26L * B, cudaMemcpyDeviceToDevice, s));
  return FOCR_OK;
}

// get_crnn_pred / strLabelConverter.decode on the device: logits (T,B,C) fp32 (class 0 = blank).
// path (B,T) raw argmax indices, out (B,T) collapsed indices padded with -1, len (B).  INT32, bit-exact.
int focr_ctc_greedy_decode(const float* logits, int T, int B, int C, int* path, int* out, int* len, void* stream) {
  ctc_greedy_kernel<<<focr_cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(logits, C, T, B, C, path, out, len);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

}  // extern "C"
