// Stroke-/text-focus loss: the frozen recogniser ("loss net") of text-gestalt and scene-text-telescope and the
// attention-map L1 term built on it, forward for the HR and SR crops and input-gradient-only backward for the SR crop.
//
// Reference (TG = text-gestalt, STT = scene-text-telescope):
//   TG/loss/stroke_focus_loss.py:83-122        StrokeFocusLoss.forward: mse + stroke_lambda * L1(map_hr, map_sr)
//   TG/loss/transformer_english_decomposition.py:70-168   ResNet(1, BasicBlock, [1,2,5,3]) encoder, BN in eval mode
//   ...:276-304 Decoder, :26-66 MultiHeadedAttention(h=16, d_model=1024), :343-398 Transformer.forward
//   STT/loss/transformer.py is the same network with a 37-symbol alphabet (STT/loss/text_focus_loss.py:86-99).
//
// What is NOT computed, on purpose: the reference leaves requires_grad on for the frozen recogniser, so autograd
// also produces every weight gradient of both branches and the whole HR-branch backward; nothing reads them
// (SURVEY.md A13).  Here: HR forward, SR forward, SR input-gradient chain only.  The attention-map term needs the
// decoder only up to the cross-attention softmax; the image-independent text side (embedding, masked
// self-attention, LayerNorm, query projection) is evaluated once and shared by both branches.
//
// Layout: NHWC bf16 feature maps, BatchNorm folded into the conv weights when the recogniser is prepared (it is
// frozen), every 3x3 convolution on the tcgen05 implicit-GEMM engine (tc_gemm.cu) with bias/ReLU/residual fused in
// the epilogue; backward = the same engine on flipped/transposed weights with the ReLU gates (and the skip-connection
// add) fused in the epilogue.
#include <string.h>

#include <string>
#include <vector>

#include "kernels.cuh"

#define TRY(expr)             \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

extern "C" int focr_mse_loss_grad(const float* sr, const float* hr, float* d_sr, float* loss, long n, float gscale,
                                  void* ws, size_t ws_bytes, void* stream);

namespace strokenet {

constexpr int kD = 1024;      // d_model
constexpr int kHeads = 16;    // cross / masked attention heads, d_k = 64
constexpr int kDk = 64;
constexpr int kTok = 256;     // 8 x 32 encoder positions
constexpr int kMaxT = 256;    // longest decoder input handled by the text-side kernels
constexpr int kFF = 2048;     // PositionwiseFeedForward width
constexpr int kGenPad = 64;   // generator rows padded to one GEMM tile (n_class <= 64)

// ---------------------------------------------------------------------------------------------
// topology + slots (slot i = reference state_dict entry i)
// ---------------------------------------------------------------------------------------------
struct ConvDef {
  std::string conv, bn;
  int cin, cout;
};
struct Block {
  int c1, c2, down;
};
struct Stage {
  std::vector<Block> blocks;
  int conv;
};
struct Topo {
  std::vector<ConvDef> convs;  // 0: conv1 (1->64 @32x128), 1: conv2 (64->128 @16x64), rest @8x32
  Stage stages[4];
  Topo() {
    const std::string e = "encoder.cnn.";
    convs.push_back({e + "conv1", e + "bn1", 1, 64});
    convs.push_back({e + "conv2", e + "bn2", 64, 128});
    const int nblk[4] = {1, 2, 5, 3};
    const int cin[4] = {128, 256, 256, 512}, cout[4] = {256, 256, 512, 512};
    for (int li = 0; li < 4; ++li) {
      const std::string L = e + "layer" + std::to_string(li + 1);
      for (int bi = 0; bi < nblk[li]; ++bi) {
        const std::string Bn = L + "." + std::to_string(bi) + ".";
        const int ci = bi == 0 ? cin[li] : cout[li];
        Block b;
        b.c1 = (int)convs.size();
        convs.push_back({Bn + "conv1", Bn + "bn1", ci, cout[li]});
        b.c2 = (int)convs.size();
        convs.push_back({Bn + "conv2", Bn + "bn2", cout[li], cout[li]});
        b.down = -1;
        if (bi == 0 && ci != cout[li]) {
          b.down = (int)convs.size();
          convs.push_back({Bn + "downsample.0", Bn + "downsample.1", ci, cout[li]});
        }
        stages[li].blocks.push_back(b);
      }
      stages[li].conv = (int)convs.size();
      if (li < 3)
        convs.push_back({L + "_conv", L + "_bn", cout[li], cout[li]});
      else
        convs.push_back({L + "_conv2", L + "_conv2_bn", 512, 1024});
    }
  }
};
const Topo& topo() {
  static const Topo t;
  return t;
}

enum Dec : int {
  D_MQ_W, D_MQ_B, D_MK_W, D_MK_B, D_MV_W, D_MV_B, D_MO_W, D_MO_B, D_MC_W, D_MC_B, D_LN1A, D_LN1B,
  D_XQ_W, D_XQ_B, D_XK_W, D_XK_B, D_XV_W, D_XV_B, D_XO_W, D_XO_B, D_XC_W, D_XC_B, D_LN2A, D_LN2B,
  D_W1_W, D_W1_B, D_W2_W, D_W2_B, D_LN3A, D_LN3B, D_GEN_W, D_GEN_B, D_COUNT
};
inline int conv_slot(int ci) { return 2 + 7 * ci; }
inline int dec_slot0() { return 2 + 7 * (int)topo().convs.size(); }
inline int num_slots() { return dec_slot0() + D_COUNT; }

// variant 0: text-gestalt names (embedding_word_with_upperword / generator_word_with_upperword),
// variant 1: scene-text-telescope names (embedding_word / generator_word)
std::vector<std::string> slot_names(int variant) {
  const Topo& t = topo();
  std::vector<std::string> v;
  const std::string suffix = variant == 0 ? "_with_upperword" : "";
  v.push_back("embedding_word" + suffix + ".lut.weight");
  v.push_back("pe.pe");
  for (const ConvDef& c : t.convs) {
    v.push_back(c.conv + ".weight");
    v.push_back(c.conv + ".bias");
    v.push_back(c.bn + ".weight");
    v.push_back(c.bn + ".bias");
    v.push_back(c.bn + ".running_mean");
    v.push_back(c.bn + ".running_var");
    v.push_back(c.bn + ".num_batches_tracked");
  }
  const char* mh[2] = {"decoder.mask_multihead.", "decoder.multihead."};
  const char* ln[3] = {"decoder.mul_layernorm1.", "decoder.mul_layernorm2.", "decoder.mul_layernorm3."};
  for (int m = 0; m < 2; ++m) {
    for (int l = 0; l < 4; ++l) {
      v.push_back(std::string(mh[m]) + "linears." + std::to_string(l) + ".weight");
      v.push_back(std::string(mh[m]) + "linears." + std::to_string(l) + ".bias");
    }
    v.push_back(std::string(mh[m]) + "compress_attention_linear.weight");
    v.push_back(std::string(mh[m]) + "compress_attention_linear.bias");
    v.push_back(std::string(ln[m]) + "a_2");
    v.push_back(std::string(ln[m]) + "b_2");
  }
  v.push_back("decoder.pff.w_1.weight");
  v.push_back("decoder.pff.w_1.bias");
  v.push_back("decoder.pff.w_2.weight");
  v.push_back("decoder.pff.w_2.bias");
  v.push_back(std::string(ln[2]) + "a_2");
  v.push_back(std::string(ln[2]) + "b_2");
  v.push_back("generator_word" + suffix + ".proj.weight");
  v.push_back("generator_word" + suffix + ".proj.bias");
  return v;
}

// ---------------------------------------------------------------------------------------------
// bump allocator shared by the prepared-weights blob and the workspace
// ---------------------------------------------------------------------------------------------
struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base((char*)b) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) / 256 * 256;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct ConvW {
  int cin, cout;
  bf16 *wf, *wd;  // forward [9][cout][cin], dgrad [9][cin][cout] (flipped taps); BN folded
  float* bias;    // folded bias [cout]
};
struct Prep {
  std::vector<ConvW> conv;  // conv[0].wf/wd unused (see w1, b1)
  float *w1, *b1;           // conv1 folded: fp32 [64][9], [64]
  float* lut;               // [n_class][512]
  float* pe;                // [kMaxT][512]
  bf16 *mq, *mk, *mv, *mo;  // masked self-attention linears [1024][1024]
  float *mqb, *mkb, *mvb, *mob;
  float *ln1a, *ln1b;
  bf16 *xq, *xk, *xkT;      // cross-attention query / key projections; xkT = [in][out] for the input gradient
  float *xqb, *xkb;
  // rest of the decoder (recognition term of TextFocusLoss): value / output projections, FFN, generator (padded to 64 rows)
  bf16 *xv, *xvT, *xo, *xoT, *ff1, *ff1T, *ff2, *ff2T, *gen, *genT;
  float *xvb, *xob, *ff1b, *ff2b, *genb, *ln2a, *ln2b, *ln3a, *ln3b;
  float* fold_tmp;          // fp32 scratch for the BN fold (largest conv)
  size_t total;
};
void prep_layout(Prep& p, int n_class, void* base) {
  const Topo& t = topo();
  Bump b(base);
  p.conv.resize(t.convs.size());
  for (size_t i = 0; i < t.convs.size(); ++i) {
    const ConvDef& c = t.convs[i];
    p.conv[i].cin = c.cin;
    p.conv[i].cout = c.cout;
    if (i == 0) {
      p.conv[i].wf = p.conv[i].wd = nullptr;
      p.conv[i].bias = nullptr;
      continue;
    }
    p.conv[i].wf = b.take<bf16>((size_t)9 * c.cin * c.cout);
    p.conv[i].wd = b.take<bf16>((size_t)9 * c.cin * c.cout);
    p.conv[i].bias = b.take<float>(c.cout);
  }
  p.w1 = b.take<float>(64 * 9);
  p.b1 = b.take<float>(64);
  p.lut = b.take<float>((size_t)n_class * 512);
  p.pe = b.take<float>((size_t)kMaxT * 512);
  bf16** lin[7] = {&p.mq, &p.mk, &p.mv, &p.mo, &p.xq, &p.xk, &p.xkT};
  for (auto l : lin) *l = b.take<bf16>((size_t)kD * kD);
  float** vec[8] = {&p.mqb, &p.mkb, &p.mvb, &p.mob, &p.ln1a, &p.ln1b, &p.xqb, &p.xkb};
  for (auto l : vec) *l = b.take<float>(kD);
  bf16** lin2[4] = {&p.xv, &p.xvT, &p.xo, &p.xoT};
  for (auto l : lin2) *l = b.take<bf16>((size_t)kD * kD);
  bf16** ffw[4] = {&p.ff1, &p.ff1T, &p.ff2, &p.ff2T};
  for (auto l : ffw) *l = b.take<bf16>((size_t)kD * kFF);
  p.gen = b.take<bf16>((size_t)kGenPad * kD);
  p.genT = b.take<bf16>((size_t)kD * kGenPad);
  float** vec2[7] = {&p.xvb, &p.xob, &p.ff2b, &p.ln2a, &p.ln2b, &p.ln3a, &p.ln3b};
  for (auto l : vec2) *l = b.take<float>(kD);
  p.ff1b = b.take<float>(kFF);
  p.genb = b.take<float>(kGenPad);
  p.fold_tmp = b.take<float>((size_t)9 * 512 * 1024);
  p.total = (b.off + 255) / 256 * 256;
}

struct Ws {
  long Mt;  // text rows padded to 128
  bf16 *text, *tq, *tk, *tv, *tctx, *x1, *query, *Q;
  bf16 *a1, *p1, *a2, *p2;
  std::vector<bf16*> act;  // per conv index >= 2
  bf16* tmpdown;
  bf16* Kp;
  float *map_hr, *map_sr;
  bf16* g[4];
  bf16 *g_a2, *g_p1, *g_a1;
  float* partial;  // [B*16] L1 partials + mse scratch
  float* scal;     // [4]
  // recognition term (SR branch only)
  bf16 *Vp, *ctx, *x2, *r2, *hff, *x3, *r3, *dlogits, *dr3, *dx3, *dh, *dr2, *dx2, *dctx;
  float *logits, *ce_partial;
  size_t total;
};
void ws_layout(Ws& w, int B, int T, void* base) {
  const Topo& t = topo();
  Bump b(base);
  w.Mt = ((long)B * T + 127) / 128 * 128;
  bf16** txt[8] = {&w.text, &w.tq, &w.tk, &w.tv, &w.tctx, &w.x1, &w.query, &w.Q};
  for (auto p : txt) *p = b.take<bf16>((size_t)w.Mt * kD);
  w.a1 = b.take<bf16>((size_t)B * 4096 * 64);
  w.p1 = b.take<bf16>((size_t)B * 1024 * 64);
  w.a2 = b.take<bf16>((size_t)B * 1024 * 128);
  w.p2 = b.take<bf16>((size_t)B * 256 * 128);
  w.act.assign(t.convs.size(), nullptr);
  for (size_t i = 2; i < t.convs.size(); ++i) w.act[i] = b.take<bf16>((size_t)B * kTok * t.convs[i].cout);
  w.tmpdown = b.take<bf16>((size_t)B * kTok * 512);
  w.Kp = b.take<bf16>((size_t)B * kTok * kD);
  w.map_hr = b.take<float>((size_t)B * kHeads * T * kTok);
  w.map_sr = b.take<float>((size_t)B * kHeads * T * kTok);
  for (int i = 0; i < 4; ++i) w.g[i] = b.take<bf16>((size_t)B * kTok * kD);
  w.g_a2 = b.take<bf16>((size_t)B * 1024 * 128);
  w.g_p1 = b.take<bf16>((size_t)B * 1024 * 64);
  w.g_a1 = b.take<bf16>((size_t)B * 4096 * 64);
  w.partial = b.take<float>((size_t)B * kHeads + 4096);
  w.scal = b.take<float>(8);
  w.Vp = b.take<bf16>((size_t)B * kTok * kD);
  bf16** t1[10] = {&w.ctx, &w.x2, &w.r2, &w.x3, &w.r3, &w.dr3, &w.dx3, &w.dr2, &w.dx2, &w.dctx};
  for (auto p : t1) *p = b.take<bf16>((size_t)w.Mt * kD);
  w.hff = b.take<bf16>((size_t)w.Mt * kFF);
  w.dh = b.take<bf16>((size_t)w.Mt * kFF);
  w.dlogits = b.take<bf16>((size_t)w.Mt * kGenPad);
  w.logits = b.take<float>((size_t)w.Mt * kGenPad);
  w.ce_partial = b.take<float>((size_t)B + 8);
  w.total = (b.off + 255) / 256 * 256;
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
int sgrid(long n, int per) {
  long g = (n + per - 1) / per;
  if (g > 148L * 16) g = 148L * 16;
  if (g < 1) g = 1;
  return (int)g;
}

// eval-mode BatchNorm folded into the preceding conv: w' = w * s[co], b' = (b - mean) * s + beta, s = gamma/sqrt(var+eps)
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ rm, const float* __restrict__ rv,
                               float eps, float* __restrict__ wo, float* __restrict__ bo, int Co, long per) {
  const long n = (long)Co * per;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i / per);
    const float sc = gamma[co] * rsqrtf(rv[co] + eps);
    wo[i] = w[i] * sc;
    if (i % per == 0) bo[co] = (b[co] - rm[co]) * sc + beta[co];
  }
}

__global__ void copy_f32_kernel(const float* __restrict__ a, float* __restrict__ o, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) o[i] = a[i];
}

// gray = 0.299 R + 0.587 G + 0.114 B (stroke_focus_loss.py:12-18) -> conv 1->64 3x3 pad 1 (+folded BN) -> ReLU,
// one CTA per image row, one thread per pixel; out bf16 NHWC (B,32,128,64)
__global__ void __launch_bounds__(128) conv1_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, bf16* __restrict__ out) {
  __shared__ float gs[3][130];
  __shared__ float ws[64 * 9 + 64];
  const int b = blockIdx.x >> 5, y = blockIdx.x & 31, x = threadIdx.x;
  for (int i = x; i < 64 * 9 + 64; i += 128) ws[i] = i < 576 ? w1[i] : b1[i - 576];
  const float* im = img + (long)b * 3 * 4096;
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    float g = 0.f;
    if (yy >= 0 && yy < 32) {
      const long o = (long)yy * 128 + x;
      g = 0.299f * im[o] + 0.587f * im[4096 + o] + 0.114f * im[8192 + o];
    }
    gs[r][x + 1] = g;
    if (x == 0) {
      gs[r][0] = 0.f;
      gs[r][129] = 0.f;
    }
  }
  __syncthreads();
  float t[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) t[r * 3 + c] = gs[r][x + c];
  uint4* op = reinterpret_cast<uint4*>(out + (((long)b * 32 + y) * 128 + x) * 64);
#pragma unroll 1
  for (int c8 = 0; c8 < 8; ++c8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      float a = ws[576 + c];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(ws[c * 9 + k], t[k], a);
      v[j] = fmaxf(a, 0.f);
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    op[c8] = o;
  }
}

// input gradient of conv1 + gray: d_img[b,ch,y,x] += coef[ch] * sum_{tap,c} w1[c][tap] * g[b, y-ky+1, x-kx+1, c]
// (g is already ReLU-gated by the max-pool backward)
__global__ void __launch_bounds__(128) conv1_dgrad_kernel(const bf16* __restrict__ g, const float* __restrict__ w1,
                                                          float* __restrict__ d_img) {
  __shared__ float wt[9][64];
  const int b = blockIdx.x >> 5, y = blockIdx.x & 31, x = threadIdx.x;
  for (int i = x; i < 576; i += 128) wt[i % 9][i / 9] = w1[i];
  __syncthreads();
  float acc = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y - ky + 1;
    if (yy < 0 || yy >= 32) continue;
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x - kx + 1;
      if (xx < 0 || xx >= 128) continue;
      const uint4* gp = reinterpret_cast<const uint4*>(g + (((long)b * 32 + yy) * 128 + xx) * 64);
      const float* wr = wt[ky * 3 + kx];
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 u = gp[c8];
        float2 f;
        f = unpack_bf16x2(u.x); acc = fmaf(f.x, wr[c8 * 8 + 0], acc); acc = fmaf(f.y, wr[c8 * 8 + 1], acc);
        f = unpack_bf16x2(u.y); acc = fmaf(f.x, wr[c8 * 8 + 2], acc); acc = fmaf(f.y, wr[c8 * 8 + 3], acc);
        f = unpack_bf16x2(u.z); acc = fmaf(f.x, wr[c8 * 8 + 4], acc); acc = fmaf(f.y, wr[c8 * 8 + 5], acc);
        f = unpack_bf16x2(u.w); acc = fmaf(f.x, wr[c8 * 8 + 6], acc); acc = fmaf(f.y, wr[c8 * 8 + 7], acc);
      }
    }
  }
  float* d = d_img + (long)b * 3 * 4096 + (long)y * 128 + x;
  d[0] += 0.299f * acc;
  d[4096] += 0.587f * acc;
  d[8192] += 0.114f * acc;
}

// 2x2/2 max-pool, NHWC bf16, 8 channels per thread
__global__ void pool2_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long n = (long)B * Ho * Wo * C8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % C8);
    const long pix = i / C8;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (((long)b * H + ho * 2 + (q >> 1)) * W + wo * 2 + (q & 1)) * C + cc * 8);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uw[j]);
        m[2 * j] = fmaxf(m[2 * j], f.x);
        m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y);
      }
    }
    uint4 o;
    o.x = pack_bf16x2(m[0], m[1]);
    o.y = pack_bf16x2(m[2], m[3]);
    o.z = pack_bf16x2(m[4], m[5]);
    o.w = pack_bf16x2(m[6], m[7]);
    *reinterpret_cast<uint4*>(y + pix * C + cc * 8) = o;
  }
}

// backward of relu -> max-pool: the gradient goes to the first window element equal to the max (torch's scan-order
// tie-break) and only where that maximum is positive (the ReLU in front of the pool)
__global__ void pool2_relu_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                      int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long n = (long)B * Ho * Wo * C8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % C8);
    const long pix = i / C8;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), b = (int)(pix / ((long)Wo * Ho));
    float v[4][8], m[8], g[8];
    long off[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      off[q] = (((long)b * H + ho * 2 + (q >> 1)) * W + wo * 2 + (q & 1)) * C + cc * 8;
      const uint4 u = *reinterpret_cast<const uint4*>(x + off[q]);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uw[j]);
        v[q][2 * j] = f.x;
        v[q][2 * j + 1] = f.y;
        m[2 * j] = fmaxf(m[2 * j], f.x);
        m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y);
      }
    }
    {
      const uint4 u = *reinterpret_cast<const uint4*>(dy + pix * C + cc * 8);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(uw[j]);
        g[2 * j] = f.x;
        g[2 * j + 1] = f.y;
      }
    }
    bool done[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) done[j] = !(m[j] > 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool hit = !done[j] && v[q][j] == m[j];
        o[j] = hit ? g[j] : 0.f;
        done[j] = done[j] || hit;
      }
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]);
      u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + off[q]) = u;
    }
  }
}

// text = lut[text_input] * sqrt(512) || pe[t]   (Transformer.forward :365-369), bf16 rows padded with zeros to Mt
__global__ void text_embed_kernel(const long long* __restrict__ text_input, const float* __restrict__ lut,
                                  const float* __restrict__ pe, bf16* __restrict__ out, int BT, int T, int n_class) {
  const long row = blockIdx.x;
  bf16* o = out + row * kD;
  if (row >= BT) {
    for (int c = threadIdx.x; c < kD; c += blockDim.x) o[c] = __float2bfloat16_rn(0.f);
    return;
  }
  const int t = (int)(row % T);
  long long idx = text_input[row];
  if (idx < 0) idx = 0;
  if (idx >= n_class) idx = n_class - 1;
  const float sc = 22.627416997969522f;  // sqrt(512)
  for (int c = threadIdx.x; c < 512; c += blockDim.x) {
    o[c] = __float2bfloat16_rn(lut[idx * 512 + c] * sc);
    o[512 + c] = __float2bfloat16_rn(pe[(long)t * 512 + c]);
  }
}

// masked (causal) self-attention over the decoder input, one CTA per (sample, head), one warp per query row.
// q,k,v: bf16 (B*T, 1024) with head h in columns [64h, 64h+64)   (MultiHeadedAttention.forward :57-66, attention :26-46)
__global__ void __launch_bounds__(128) text_self_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k,
                                                             const bf16* __restrict__ v, bf16* __restrict__ ctx, int T) {
  extern __shared__ uint32_t sm_u32[];
  uint32_t* Ks = sm_u32;                     // [T][33] bf16 pairs
  uint32_t* Vs = Ks + (size_t)T * 33;        // [T][33]
  float* pw = reinterpret_cast<float*>(Vs + (size_t)T * 33);  // [4][kMaxT]
  float* qs = pw + 4 * kMaxT;                // [4][64]
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < T * 32; i += 128) {
    const int r = i >> 5, c = i & 31;
    const long o = ((long)b * T + r) * kD + h * kDk + 2 * c;
    Ks[r * 33 + c] = *reinterpret_cast<const uint32_t*>(k + o);
    Vs[r * 33 + c] = *reinterpret_cast<const uint32_t*>(v + o);
  }
  __syncthreads();
  for (int t = warp; t < T; t += 4) {
    const long qo = ((long)b * T + t) * kD + h * kDk;
    const float2 qf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(q + qo + 2 * lane));
    qs[warp * 64 + 2 * lane] = qf.x;
    qs[warp * 64 + 2 * lane + 1] = qf.y;
    __syncwarp();
    float sc[kMaxT / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxT / 32; ++j) {
      const int kk = lane + 32 * j;
      float s = -INFINITY;
      if (kk <= t) {
        s = 0.f;
#pragma unroll 8
        for (int d2 = 0; d2 < 32; ++d2) {
          const float2 kf = unpack_bf16x2(Ks[kk * 33 + d2]);
          s = fmaf(qs[warp * 64 + 2 * d2], kf.x, s);
          s = fmaf(qs[warp * 64 + 2 * d2 + 1], kf.y, s);
        }
        s *= 0.125f;
      }
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxT / 32; ++j) {
      const int kk = lane + 32 * j;
      const float e = kk <= t ? __expf(sc[j] - mx) : 0.f;
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < kMaxT / 32; ++j) {
      const int kk = lane + 32 * j;
      if (kk < T) pw[warp * kMaxT + kk] = sc[j] * inv;
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int kk = 0; kk <= t; ++kk) {
      const float p = pw[warp * kMaxT + kk];
      const float2 vf = unpack_bf16x2(Vs[kk * 33 + lane]);
      a0 = fmaf(p, vf.x, a0);
      a1 = fmaf(p, vf.y, a1);
    }
    *reinterpret_cast<uint32_t*>(ctx + qo + 2 * lane) = pack_bf16x2(a0, a1);
    __syncwarp();
  }
}

// the recogniser's LayerNorm over 1024 features: a (x - mean) / (std_unbiased + eps) + b   (:222-234); warp per row
__global__ void ln1024_kernel(const bf16* __restrict__ x, const float* __restrict__ a, const float* __restrict__ bb,
                              bf16* __restrict__ y, long rows, float eps) {
  const int lane = threadIdx.x & 31;
  const long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[32];
  const uint4* xp = reinterpret_cast<const uint4*>(x + row * kD);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = xp[lane + 32 * j];
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = unpack_bf16x2(uw[q]);
      v[j * 8 + 2 * q] = f.x;
      v[j * 8 + 2 * q + 1] = f.y;
      s += f.x + f.y;
    }
  }
  const float mean = warp_sum(s) * (1.f / kD);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = v[j] - mean;
    ss += d * d;
  }
  const float sd = sqrtf(warp_sum(ss) * (1.f / (kD - 1)));
  const float inv = 1.f / (sd + eps);
  uint4* yp = reinterpret_cast<uint4*>(y + row * kD);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = a[c0 + q] * (v[j * 8 + q] - mean) * inv + bb[c0 + q];
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]);
    u.y = pack_bf16x2(o[2], o[3]);
    u.z = pack_bf16x2(o[4], o[5]);
    u.w = pack_bf16x2(o[6], o[7]);
    yp[lane + 32 * j] = u;
  }
}

// cross-attention map P[b,h,t,:] = softmax_k(Q[b,t,h,:] . K[b,k,h,:] / 8): one CTA per (sample, head), K head slice
// (256 x 64) in shared memory, one warp per query row   (Decoder.forward :294-296, attention :26-46, dropout off)
__global__ void __launch_bounds__(256) xattn_map_fwd_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K,
                                                            float* __restrict__ P, int T) {
  __shared__ uint32_t Ks[kTok * 33];
  __shared__ float qs[8][64];
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kTok * 32; i += 256) {
    const int r = i >> 5, c = i & 31;
    Ks[r * 33 + c] = *reinterpret_cast<const uint32_t*>(K + ((long)b * kTok + r) * kD + h * kDk + 2 * c);
  }
  __syncthreads();
  for (int t = warp; t < T; t += 8) {
    const float2 qf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Q + ((long)b * T + t) * kD + h * kDk + 2 * lane));
    qs[warp][2 * lane] = qf.x;
    qs[warp][2 * lane + 1] = qf.y;
    __syncwarp();
    float sc[8];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = lane + 32 * j;
      float s = 0.f;
#pragma unroll 8
      for (int d2 = 0; d2 < 32; ++d2) {
        const float2 kf = unpack_bf16x2(Ks[kk * 33 + d2]);
        s = fmaf(qs[warp][2 * d2], kf.x, s);
        s = fmaf(qs[warp][2 * d2 + 1], kf.y, s);
      }
      sc[j] = s * 0.125f;
      mx = fmaxf(mx, sc[j]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __expf(sc[j] - mx);
      sum += sc[j];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float* pr = P + (((long)b * kHeads + h) * T + t) * kTok;
#pragma unroll
    for (int j = 0; j < 8; ++j) pr[lane + 32 * j] = sc[j] * inv;
    __syncwarp();
  }
}

// L1(map_hr, map_sr) and its gradient pushed through the softmax to the key projection:
//   g = coef * sign(P_sr - P_hr);  dS = P_sr * (g - sum_k g P_sr) / 8;  dK[b,k,h,:] = sum_t dS[t,k] Q[b,t,h,:]
// one CTA per (sample, head); thread k owns dK row k (64 fp32 accumulators); l1_partial[b*16+h] = sum |P_sr - P_hr|
__global__ void __launch_bounds__(256) xattn_map_bwd_kernel(const float* __restrict__ Phr, const float* __restrict__ Psr,
                                                            const bf16* __restrict__ Q, bf16* __restrict__ dK,
                                                            float* __restrict__ l1_partial, int T, float coef) {
  __shared__ float dS[8][kTok];
  __shared__ float qs[8][64];
  __shared__ float red[8];
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
  float l1 = 0.f;
  for (int t0 = 0; t0 < T; t0 += 8) {
    const int t = t0 + warp;
    if (t < T) {
      const long ro = (((long)b * kHeads + h) * T + t) * kTok;
      float ps[8], g[8];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ps[j] = Psr[ro + lane + 32 * j];
        const float df = ps[j] - Phr[ro + lane + 32 * j];
        l1 += fabsf(df);
        g[j] = df > 0.f ? coef : (df < 0.f ? -coef : 0.f);
        dot = fmaf(g[j], ps[j], dot);
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int j = 0; j < 8; ++j) dS[warp][lane + 32 * j] = ps[j] * (g[j] - dot) * 0.125f;
      const float2 qf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Q + ((long)b * T + t) * kD + h * kDk + 2 * lane));
      qs[warp][2 * lane] = qf.x;
      qs[warp][2 * lane + 1] = qf.y;
    }
    __syncthreads();
    const int nt = min(8, T - t0);
    for (int tt = 0; tt < nt; ++tt) {
      const float ds = dS[tt][threadIdx.x];
#pragma unroll
      for (int d = 0; d < 64; ++d) acc[d] = fmaf(ds, qs[tt][d], acc[d]);
    }
    __syncthreads();
  }
  uint4* op = reinterpret_cast<uint4*>(dK + ((long)b * kTok + threadIdx.x) * kD + h * kDk);
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    uint4 u;
    u.x = pack_bf16x2(acc[c8 * 8 + 0], acc[c8 * 8 + 1]);
    u.y = pack_bf16x2(acc[c8 * 8 + 2], acc[c8 * 8 + 3]);
    u.z = pack_bf16x2(acc[c8 * 8 + 4], acc[c8 * 8 + 5]);
    u.w = pack_bf16x2(acc[c8 * 8 + 6], acc[c8 * 8 + 7]);
    op[c8] = u;
  }
  l1 = warp_sum(l1);
  if (lane == 0) red[warp] = l1;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    l1_partial[blockIdx.x] = s;
  }
}

// context rows of the cross-attention: ctx[b,t,h,:] = sum_k P[b,h,t,k] V[b,k,h,:]; one CTA per (sample, head), V head slice in
// shared memory, one warp per query row, lane owns 2 of the 64 channels   (attention() :26-46, MultiHeadedAttention :57-66)
__global__ void __launch_bounds__(256) xattn_ctx_kernel(const float* __restrict__ P, const bf16* __restrict__ V,
                                                        bf16* __restrict__ ctx, int T) {
  __shared__ uint32_t Vs[kTok * 32];
  __shared__ float ps[8][kTok];
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kTok * 32; i += 256) {
    const int r = i >> 5, c = i & 31;
    Vs[i] = *reinterpret_cast<const uint32_t*>(V + ((long)b * kTok + r) * kD + h * kDk + 2 * c);
  }
  __syncthreads();
  for (int t = warp; t < T; t += 8) {
    const float* pr = P + (((long)b * kHeads + h) * T + t) * kTok;
#pragma unroll
    for (int j = 0; j < 8; ++j) ps[warp][lane + 32 * j] = pr[lane + 32 * j];
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < kTok; ++k) {
      const float2 vf = unpack_bf16x2(Vs[k * 32 + lane]);
      a0 = fmaf(ps[warp][k], vf.x, a0);
      a1 = fmaf(ps[warp][k], vf.y, a1);
    }
    *reinterpret_cast<uint32_t*>(ctx + ((long)b * T + t) * kD + h * kDk + 2 * lane) = pack_bf16x2(a0, a1);
    __syncwarp();
  }
}

// input gradient of the recogniser's LayerNorm (1024 features, unbiased std, eps on std); warp per row
//   g = dy * a;  dx = inv (g - mean(g)) - xc inv^2 sum(g xc) / ((n-1) std),  inv = 1/(std+eps), xc = x - mean
__global__ void ln1024_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ a,
                                  bf16* __restrict__ dx, long rows, float eps) {
  const int lane = threadIdx.x & 31;
  const long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[32], g[32];
  const uint4* xp = reinterpret_cast<const uint4*>(x + row * kD);
  const uint4* gp = reinterpret_cast<const uint4*>(dy + row * kD);
  float s = 0.f, sg = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = xp[lane + 32 * j], w = gp[lane + 32 * j];
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, ww[4] = {w.x, w.y, w.z, w.w};
    const int c0 = (lane + 32 * j) * 8;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = unpack_bf16x2(uw[q]), d = unpack_bf16x2(ww[q]);
      v[j * 8 + 2 * q] = f.x;
      v[j * 8 + 2 * q + 1] = f.y;
      g[j * 8 + 2 * q] = d.x * a[c0 + 2 * q];
      g[j * 8 + 2 * q + 1] = d.y * a[c0 + 2 * q + 1];
      s += f.x + f.y;
      sg += g[j * 8 + 2 * q] + g[j * 8 + 2 * q + 1];
    }
  }
  const float mean = warp_sum(s) * (1.f / kD);
  const float gmean = warp_sum(sg) * (1.f / kD);
  float ss = 0.f, sgx = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] -= mean;
    ss = fmaf(v[j], v[j], ss);
    sgx = fmaf(g[j], v[j], sgx);
  }
  const float sd = sqrtf(warp_sum(ss) * (1.f / (kD - 1)));
  sgx = warp_sum(sgx);
  const float inv = 1.f / (sd + eps);
  const float k2 = sd > 0.f ? inv * inv * sgx / ((kD - 1) * sd) : 0.f;
  uint4* op = reinterpret_cast<uint4*>(dx + row * kD);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = inv * (g[j * 8 + q] - gmean) - v[j * 8 + q] * k2;
    uint4 u;
    u.x = pack_bf16x2(o[0], o[1]);
    u.y = pack_bf16x2(o[2], o[3]);
    u.z = pack_bf16x2(o[4], o[5]);
    u.w = pack_bf16x2(o[6], o[7]);
    op[lane + 32 * j] = u;
  }
}

// weight_cross_entropy (STT/loss/weight_ce_loss.py:36-45) on the valid positions (t < length[b]) of the logits (Mt, 64 pad):
//   loss_i = -log( w[gt_i][gt_i] e^{p_gt} / sum_j w[gt_i][j] e^{p_j} ), mean over all valid positions; evaluated with
//   log-sum-exp.  One CTA per sample, one warp per position; writes dlogits = coef * (softmax_w - onehot) (bf16, zero for
//   invalid rows / padded classes), the packed logits (optional) and the per-sample loss sum.
__global__ void __launch_bounds__(128) wce_kernel(const float* __restrict__ logits, const long long* __restrict__ length,
                                                  const long long* __restrict__ gt, const float* __restrict__ table,
                                                  bf16* __restrict__ dlogits, float* __restrict__ packed,
                                                  float* __restrict__ partial, int B, int T, int n_class, float lam_gscale) {
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long start = 0, total = 0;
  for (int i = 0; i < B; ++i) {
    const long long l = length[i];
    if (i < b) start += l;
    total += l;
  }
  const int len = (int)length[b];
  const float coef = lam_gscale / (float)total;
  __shared__ float red[4];
  float lsum = 0.f;
  for (int t = warp; t < T; t += 4) {
    const long row = (long)b * T + t;
    bf16* dr = dlogits + row * kGenPad;
    if (t >= len) {
      dr[lane] = __float2bfloat16_rn(0.f);
      dr[lane + 32] = __float2bfloat16_rn(0.f);
      continue;
    }
    const int g = (int)gt[start + t];
    const float* pr = logits + row * kGenPad;
    float z[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = lane + 32 * j;
      z[j] = c < n_class ? pr[c] + __logf(table[g * n_class + c]) : -INFINITY;
      if (packed != nullptr && c < n_class) packed[(start + t) * n_class + c] = pr[c];
    }
    const float mx = warp_max(fmaxf(z[0], z[1]));
    const float e0 = __expf(z[0] - mx), e1 = __expf(z[1] - mx);
    const float Z = warp_sum(e0 + e1);
    const float zg = __shfl_sync(0xffffffffu, g < 32 ? z[0] : z[1], g & 31);
    lsum += (mx + __logf(Z)) - zg;
    const float inv = 1.f / Z;
    dr[lane] = __float2bfloat16_rn(coef * (e0 * inv - (lane == g ? 1.f : 0.f)));
    dr[lane + 32] = __float2bfloat16_rn(coef * (e1 * inv - (lane + 32 == g ? 1.f : 0.f)));
  }
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    partial[b] = red[0] + red[1] + red[2] + red[3];
    if (b == 0) partial[B] = (float)total;
  }
}

// cross-attention backward with both loss terms: dP = L1-map term + dctx . V^T (recognition term);
//   dS = P (dP - sum_k dP P) / 8;  dK[b,k,h,:] = sum_t dS[t,k] Q[b,t,h,:];  dV[b,k,h,:] = sum_t P[t,k] dctx[b,t,h,:]
// one CTA per (sample, head), thread k owns rows k of dK and dV (2 x 64 fp32 accumulators)
__global__ void __launch_bounds__(256) xattn_full_bwd_kernel(const float* __restrict__ Phr, const float* __restrict__ Psr,
                                                             const bf16* __restrict__ Q, const bf16* __restrict__ V,
                                                             const bf16* __restrict__ dctx, bf16* __restrict__ dK,
                                                             bf16* __restrict__ dV, float* __restrict__ l1_partial, int T,
                                                             float coef) {
  extern __shared__ uint32_t sm_dyn[];
  uint32_t* Vs = sm_dyn;                                         // [256][33] bf16 pairs
  float* dS = reinterpret_cast<float*>(Vs + kTok * 33);          // [8][256]
  float* Pm = dS + 8 * kTok;                                     // [8][256]
  float* qs = Pm + 8 * kTok;                                     // [8][64]
  float* dcs = qs + 8 * 64;                                      // [8][64]
  __shared__ float red[8];
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kTok * 32; i += 256) {
    const int r = i >> 5, c = i & 31;
    Vs[r * 33 + c] = *reinterpret_cast<const uint32_t*>(V + ((long)b * kTok + r) * kD + h * kDk + 2 * c);
  }
  float accK[64], accV[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) accK[d] = accV[d] = 0.f;
  float l1 = 0.f;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += 8) {
    const int t = t0 + warp;
    if (t < T) {
      const long qo = ((long)b * T + t) * kD + h * kDk + 2 * lane;
      const float2 qf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Q + qo));
      const float2 cf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dctx + qo));
      qs[warp * 64 + 2 * lane] = qf.x;
      qs[warp * 64 + 2 * lane + 1] = qf.y;
      dcs[warp * 64 + 2 * lane] = cf.x;
      dcs[warp * 64 + 2 * lane + 1] = cf.y;
      __syncwarp();
      const long ro = (((long)b * kHeads + h) * T + t) * kTok;
      float ps[8], g[8];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = lane + 32 * j;
        ps[j] = Psr[ro + kk];
        const float df = ps[j] - Phr[ro + kk];
        l1 += fabsf(df);
        float gg = df > 0.f ? coef : (df < 0.f ? -coef : 0.f);
#pragma unroll 8
        for (int d2 = 0; d2 < 32; ++d2) {
          const float2 vf = unpack_bf16x2(Vs[kk * 33 + d2]);
          gg = fmaf(dcs[warp * 64 + 2 * d2], vf.x, gg);
          gg = fmaf(dcs[warp * 64 + 2 * d2 + 1], vf.y, gg);
        }
        g[j] = gg;
        dot = fmaf(gg, ps[j], dot);
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dS[warp * kTok + lane + 32 * j] = ps[j] * (g[j] - dot) * 0.125f;
        Pm[warp * kTok + lane + 32 * j] = ps[j];
      }
    }
    __syncthreads();
    const int nt = min(8, T - t0);
    for (int tt = 0; tt < nt; ++tt) {
      const float ds = dS[tt * kTok + threadIdx.x], pv = Pm[tt * kTok + threadIdx.x];
#pragma unroll
      for (int d = 0; d < 64; ++d) {
        accK[d] = fmaf(ds, qs[tt * 64 + d], accK[d]);
        accV[d] = fmaf(pv, dcs[tt * 64 + d], accV[d]);
      }
    }
    __syncthreads();
  }
  uint4* opk = reinterpret_cast<uint4*>(dK + ((long)b * kTok + threadIdx.x) * kD + h * kDk);
  uint4* opv = reinterpret_cast<uint4*>(dV + ((long)b * kTok + threadIdx.x) * kD + h * kDk);
#pragma unroll
  for (int c8 = 0; c8 < 8; ++c8) {
    uint4 u, w;
    u.x = pack_bf16x2(accK[c8 * 8 + 0], accK[c8 * 8 + 1]);
    u.y = pack_bf16x2(accK[c8 * 8 + 2], accK[c8 * 8 + 3]);
    u.z = pack_bf16x2(accK[c8 * 8 + 4], accK[c8 * 8 + 5]);
    u.w = pack_bf16x2(accK[c8 * 8 + 6], accK[c8 * 8 + 7]);
    w.x = pack_bf16x2(accV[c8 * 8 + 0], accV[c8 * 8 + 1]);
    w.y = pack_bf16x2(accV[c8 * 8 + 2], accV[c8 * 8 + 3]);
    w.z = pack_bf16x2(accV[c8 * 8 + 4], accV[c8 * 8 + 5]);
    w.w = pack_bf16x2(accV[c8 * 8 + 6], accV[c8 * 8 + 7]);
    opk[c8] = u;
    opv[c8] = w;
  }
  l1 = warp_sum(l1);
  if (lane == 0) red[warp] = l1;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    l1_partial[blockIdx.x] = s;
  }
}

// losses = {mse + la * attention + lc * recognition, mse, attention, recognition}   (single warp)
__global__ void finish_loss4_kernel(const float* __restrict__ partial, int n, float inv_numel, float la,
                                    const float* __restrict__ ce_partial, int B, float lc, float* __restrict__ losses) {
  float s = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  for (int i = threadIdx.x; i < B; i += 32) c += ce_partial[i];
  s = warp_sum(s);
  c = warp_sum(c);
  if (threadIdx.x == 0) {
    const float att = s * inv_numel;
    const float rec = c / ce_partial[B];
    losses[2] = att;
    losses[3] = rec;
    losses[0] = losses[1] + la * att + lc * rec;
  }
}

// losses[2] = attention loss = sum(partials) / numel; losses[0] = mse + lambda * attention   (single warp)
__global__ void finish_loss_kernel(const float* __restrict__ partial, int n, float inv_numel, float lambda,
                                   float* __restrict__ losses) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const float att = s * inv_numel;
    losses[2] = att;
    losses[0] = losses[1] + lambda * att;
  }
}

// ---------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------
TcGemmParams gp() {
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.kh = p.kw = 1;
  p.epi = TC_EPI_BF16;
  p.gate_scale = 1.f;
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
int conv3(const bf16* x, int B, int H, int W, const ConvW& c, bf16* out, int relu, const bf16* residual, int relu_post,
          cudaStream_t s) {
  TcGemmParams p = gp();
  p.n_total = c.cout;
  p.kh = p.kw = 3;
  p.W = W;
  p.H = H;
  p.relu = relu;
  p.ldc = c.cout;
  p.bias = c.bias;
  p.out = out;
  p.residual = residual;
  p.relu_post = relu_post;
  const bf16* ap[1] = {x};
  return tc_gemm_launch(ap, 1, c.cin, (long)W * c.cin, (long)H * W * c.cin, c.cin, B, c.wf, c.cin, p, s);
}
// dx = conv_transpose(dy) (+ residual) gated by gate > 0
int dgrad3(const bf16* dy, int B, int H, int W, const ConvW& c, bf16* dx, const bf16* residual, const bf16* gate,
           cudaStream_t s) {
  TcGemmParams p = gp();
  p.n_total = c.cin;
  p.kh = p.kw = 3;
  p.W = W;
  p.H = H;
  p.ldc = c.cin;
  p.out = dx;
  p.residual = residual;
  p.gate = gate;
  const bf16* ap[1] = {dy};
  return tc_gemm_launch(ap, 1, c.cout, (long)W * c.cout, (long)H * W * c.cout, c.cout, B, c.wd, c.cout, p, s);
}

int prepare(void* const* prm, int n_class, Prep& pw, cudaStream_t s) {
  const Topo& t = topo();
  for (size_t i = 0; i < t.convs.size(); ++i) {
    const ConvDef& c = t.convs[i];
    const int s0 = conv_slot((int)i);
    const long per = (long)c.cin * 9;
    float* wo = i == 0 ? pw.w1 : pw.fold_tmp;
    float* bo = i == 0 ? pw.b1 : pw.conv[i].bias;
    fold_bn_kernel<<<sgrid((long)c.cout * per, 256), 256, 0, s>>>(
        (const float*)prm[s0], (const float*)prm[s0 + 1], (const float*)prm[s0 + 2], (const float*)prm[s0 + 3],
        (const float*)prm[s0 + 4], (const float*)prm[s0 + 5], 1e-5f, wo, bo, c.cout, per);
    FOCR_LAUNCH_CHECK();
    if (i == 0) continue;
    TRY(prep_conv_w_fwd(pw.fold_tmp, pw.conv[i].wf, c.cout, c.cin, 3, 0, s));
    TRY(prep_conv_w_dgrad(pw.fold_tmp, pw.conv[i].wd, c.cout, c.cin, 3, 0, s));
  }
  auto cp = [&](const void* src, float* dst, long n) {
    copy_f32_kernel<<<sgrid(n, 256), 256, 0, s>>>((const float*)src, dst, n);
    focr_count_launch(1);
  };
  cp(prm[0], pw.lut, (long)n_class * 512);
  cp(prm[1], pw.pe, (long)kMaxT * 512);
  const int d0 = dec_slot0();
  TRY(prep_linear_w((const float*)prm[d0 + D_MQ_W], pw.mq, nullptr, kD, kD, 0, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_MK_W], pw.mk, nullptr, kD, kD, 0, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_MV_W], pw.mv, nullptr, kD, kD, 0, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_MO_W], pw.mo, nullptr, kD, kD, 0, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_XQ_W], pw.xq, nullptr, kD, kD, 0, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_XK_W], pw.xk, pw.xkT, kD, kD, kD, 0, s));
  cp(prm[d0 + D_MQ_B], pw.mqb, kD);
  cp(prm[d0 + D_MK_B], pw.mkb, kD);
  cp(prm[d0 + D_MV_B], pw.mvb, kD);
  cp(prm[d0 + D_MO_B], pw.mob, kD);
  cp(prm[d0 + D_LN1A], pw.ln1a, kD);
  cp(prm[d0 + D_LN1B], pw.ln1b, kD);
  cp(prm[d0 + D_XQ_B], pw.xqb, kD);
  cp(prm[d0 + D_XK_B], pw.xkb, kD);
  // recognition-term half of the decoder
  FOCR_REQUIRE(n_class <= kGenPad, "strokenet_prepare: n_class %d > %d", n_class, kGenPad);
  TRY(prep_linear_w((const float*)prm[d0 + D_XV_W], pw.xv, pw.xvT, kD, kD, kD, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_XO_W], pw.xo, pw.xoT, kD, kD, kD, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_W1_W], pw.ff1, pw.ff1T, kFF, kD, kFF, 0, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_W2_W], pw.ff2, pw.ff2T, kD, kFF, kD, 0, s));
  FOCR_CHECK_CUDA(cudaMemsetAsync(pw.gen, 0, (size_t)kGenPad * kD * 2, s));
  FOCR_CHECK_CUDA(cudaMemsetAsync(pw.genT, 0, (size_t)kGenPad * kD * 2, s));
  FOCR_CHECK_CUDA(cudaMemsetAsync(pw.genb, 0, (size_t)kGenPad * 4, s));
  TRY(prep_linear_w((const float*)prm[d0 + D_GEN_W], pw.gen, pw.genT, n_class, kD, kGenPad, 0, s));
  cp(prm[d0 + D_XV_B], pw.xvb, kD);
  cp(prm[d0 + D_XO_B], pw.xob, kD);
  cp(prm[d0 + D_W1_B], pw.ff1b, kFF);
  cp(prm[d0 + D_W2_B], pw.ff2b, kD);
  cp(prm[d0 + D_GEN_B], pw.genb, n_class);
  cp(prm[d0 + D_LN2A], pw.ln2a, kD);
  cp(prm[d0 + D_LN2B], pw.ln2b, kD);
  cp(prm[d0 + D_LN3A], pw.ln3a, kD);
  cp(prm[d0 + D_LN3B], pw.ln3b, kD);
  FOCR_CHECK_CUDA(cudaGetLastError());
  return FOCR_OK;
}

// image-independent decoder prefix: query = LN1(text + maskedMHA(text)); Q = W_q query + b_q
int text_side(const Prep& pw, const long long* text_input, int B, int T, int n_class, Ws& w, cudaStream_t s) {
  ProfScope _ps("focus_text", s);
  text_embed_kernel<<<(unsigned)w.Mt, 128, 0, s>>>(text_input, pw.lut, pw.pe, w.text, B * T, T, n_class);
  FOCR_LAUNCH_CHECK();
  TcGemmParams p = gp();
  p.bias = pw.mqb;
  p.out = w.tq;
  TRY(tok_gemm(w.text, kD, w.Mt, pw.mq, kD, p, s));
  p.bias = pw.mkb;
  p.out = w.tk;
  TRY(tok_gemm(w.text, kD, w.Mt, pw.mk, kD, p, s));
  p.bias = pw.mvb;
  p.out = w.tv;
  TRY(tok_gemm(w.text, kD, w.Mt, pw.mv, kD, p, s));
  const size_t smem = (size_t)T * 33 * 4 * 2 + (size_t)4 * kMaxT * 4 + 4 * 64 * 4;
  static bool attr = false;
  if (!attr) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(text_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kMaxT * 33 * 4 * 2 + 4 * kMaxT * 4 + 4 * 64 * 4));
    attr = true;
  }
  FOCR_CHECK_CUDA(cudaMemsetAsync(w.tctx, 0, (size_t)w.Mt * kD * 2, s));
  text_self_attn_kernel<<<B * kHeads, 128, smem, s>>>(w.tq, w.tk, w.tv, w.tctx, T);
  FOCR_LAUNCH_CHECK();
  p = gp();
  p.bias = pw.mob;
  p.out = w.x1;
  p.residual = w.text;
  TRY(tok_gemm(w.tctx, kD, w.Mt, pw.mo, kD, p, s));
  ln1024_kernel<<<focr_cdiv(w.Mt, 8), 256, 0, s>>>(w.x1, pw.ln1a, pw.ln1b, w.query, w.Mt, 1e-6f);
  FOCR_LAUNCH_CHECK();
  p = gp();
  p.bias = pw.xqb;
  p.out = w.Q;
  TRY(tok_gemm(w.query, kD, w.Mt, pw.xq, kD, p, s));
  return FOCR_OK;
}

// encoder + key projection + attention map of one crop batch (img: fp32 NCHW (B,3,32,128) in [0,1])
int branch_forward(const Prep& pw, const float* img, int B, int T, Ws& w, float* map_out, cudaStream_t s) {
  const Topo& t = topo();
  {
    ProfScope _ps("focus_stem", s);
    conv1_fwd_kernel<<<B * 32, 128, 0, s>>>(img, pw.w1, pw.b1, w.a1);
    FOCR_LAUNCH_CHECK();
    pool2_fwd_kernel<<<sgrid((long)B * 1024 * 8, 256), 256, 0, s>>>(w.a1, w.p1, B, 32, 128, 64);
    FOCR_LAUNCH_CHECK();
  }
  TRY(conv3(w.p1, B, 16, 64, pw.conv[1], w.a2, 1, nullptr, 0, s));
  {
    ProfScope _ps("focus_stem", s);
    pool2_fwd_kernel<<<sgrid((long)B * 256 * 16, 256), 256, 0, s>>>(w.a2, w.p2, B, 16, 64, 128);
    FOCR_LAUNCH_CHECK();
  }
  const bf16* x = w.p2;
  for (int st = 0; st < 4; ++st) {
    for (const Block& blk : t.stages[st].blocks) {
      TRY(conv3(x, B, 8, 32, pw.conv[blk.c1], w.act[blk.c1], 1, nullptr, 0, s));
      const bf16* res = x;
      if (blk.down >= 0) {
        TRY(conv3(x, B, 8, 32, pw.conv[blk.down], w.tmpdown, 0, nullptr, 0, s));
        res = w.tmpdown;
      }
      TRY(conv3(w.act[blk.c1], B, 8, 32, pw.conv[blk.c2], w.act[blk.c2], 0, res, 1, s));
      x = w.act[blk.c2];
    }
    const int sc = t.stages[st].conv;
    TRY(conv3(x, B, 8, 32, pw.conv[sc], w.act[sc], 1, nullptr, 0, s));
    x = w.act[sc];
  }
  TcGemmParams p = gp();
  p.bias = pw.xkb;
  p.out = w.Kp;
  TRY(tok_gemm(x, kD, (long)B * kTok, pw.xk, kD, p, s));
  {
    ProfScope _ps("focus_xattn", s);
    xattn_map_fwd_kernel<<<B * kHeads, 256, 0, s>>>(w.Q, w.Kp, map_out, T);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

// SR-branch input gradient: dK (in w.g[1]) -> tokens -> encoder -> gray -> d_img (accumulated)
int branch_backward(const Prep& pw, int B, Ws& w, float* d_img, bool with_value_path, cudaStream_t s) {
  const Topo& t = topo();
  const int last = t.stages[3].conv;
  TcGemmParams p = gp();
  if (with_value_path) {  // dV (in w.g[2]) through the value projection first; added to the key path below
    p.out = w.g[3];
    TRY(tok_gemm(w.g[2], kD, (long)B * kTok, pw.xvT, kD, p, s));
    p = gp();
    p.residual = w.g[3];
  }
  p.out = w.g[0];
  p.gate = w.act[last];
  TRY(tok_gemm(w.g[1], kD, (long)B * kTok, pw.xkT, kD, p, s));
  int cur = 0;
  for (int st = 3; st >= 0; --st) {
    const Stage& S = t.stages[st];
    {
      const int nxt = (cur + 1) & 3;
      TRY(dgrad3(w.g[cur], B, 8, 32, pw.conv[S.conv], w.g[nxt], nullptr, w.act[S.blocks.back().c2], s));
      cur = nxt;
    }
    for (int bi = (int)S.blocks.size() - 1; bi >= 0; --bi) {
      const Block& blk = S.blocks[bi];
      const bf16* x_in = bi > 0 ? w.act[S.blocks[bi - 1].c2] : (st > 0 ? w.act[t.stages[st - 1].conv] : w.p2);
      const bf16* gate_x = (bi == 0 && st == 0) ? nullptr : x_in;  // p2 is a pooled map: its ReLU sits before the pool
      const int a = (cur + 1) & 3, bb = (cur + 2) & 3, c = (cur + 3) & 3;
      TRY(dgrad3(w.g[cur], B, 8, 32, pw.conv[blk.c2], w.g[a], nullptr, w.act[blk.c1], s));
      const bf16* res = w.g[cur];
      if (blk.down >= 0) {
        TRY(dgrad3(w.g[cur], B, 8, 32, pw.conv[blk.down], w.g[bb], nullptr, nullptr, s));
        res = w.g[bb];
      }
      TRY(dgrad3(w.g[a], B, 8, 32, pw.conv[blk.c1], w.g[c], res, gate_x, s));
      cur = c;
    }
  }
  {
    ProfScope _ps("focus_stem", s);
    pool2_relu_bwd_kernel<<<sgrid((long)B * 256 * 16, 256), 256, 0, s>>>(w.a2, w.g[cur], w.g_a2, B, 16, 64, 128);
    FOCR_LAUNCH_CHECK();
  }
  TRY(dgrad3(w.g_a2, B, 16, 64, pw.conv[1], w.g_p1, nullptr, nullptr, s));
  {
    ProfScope _ps("focus_stem", s);
    pool2_relu_bwd_kernel<<<sgrid((long)B * 1024 * 8, 256), 256, 0, s>>>(w.a1, w.g_p1, w.g_a1, B, 32, 128, 64);
    FOCR_LAUNCH_CHECK();
    conv1_dgrad_kernel<<<B * 32, 128, 0, s>>>(w.g_a1, pw.w1, d_img);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}

// rest of Decoder.forward for the SR branch (:296-304) + generator + weighted CE, and its backward down to dctx.
// Needs w.Kp / w.map_sr of the SR branch.  Leaves dlogits-driven dctx in w.dctx.
int decoder_tail(const Prep& pw, const long long* length, const long long* text_gt, const float* table, int B, int T,
                 int n_class, float lam_gscale, float* sr_pred_out, Ws& w, cudaStream_t s) {
  const Topo& t = topo();
  const bf16* tokens = w.act[t.stages[3].conv];
  TcGemmParams p = gp();
  p.bias = pw.xvb;
  p.out = w.Vp;
  TRY(tok_gemm(tokens, kD, (long)B * kTok, pw.xv, kD, p, s));
  {
    ProfScope _ps("focus_xattn", s);
    FOCR_CHECK_CUDA(cudaMemsetAsync(w.ctx, 0, (size_t)w.Mt * kD * 2, s));
    xattn_ctx_kernel<<<B * kHeads, 256, 0, s>>>(w.map_sr, w.Vp, w.ctx, T);
    FOCR_LAUNCH_CHECK();
  }
  p = gp();
  p.bias = pw.xob;
  p.out = w.x2;
  p.residual = w.query;
  TRY(tok_gemm(w.ctx, kD, w.Mt, pw.xo, kD, p, s));
  ln1024_kernel<<<focr_cdiv(w.Mt, 8), 256, 0, s>>>(w.x2, pw.ln2a, pw.ln2b, w.r2, w.Mt, 1e-6f);
  FOCR_LAUNCH_CHECK();
  p = gp();
  p.bias = pw.ff1b;
  p.out = w.hff;
  p.relu = 1;
  TRY(tok_gemm(w.r2, kD, w.Mt, pw.ff1, kFF, p, s));
  p = gp();
  p.bias = pw.ff2b;
  p.out = w.x3;
  p.residual = w.r2;
  TRY(tok_gemm(w.hff, kFF, w.Mt, pw.ff2, kD, p, s));
  ln1024_kernel<<<focr_cdiv(w.Mt, 8), 256, 0, s>>>(w.x3, pw.ln3a, pw.ln3b, w.r3, w.Mt, 1e-6f);
  FOCR_LAUNCH_CHECK();
  p = gp();
  p.bias = pw.genb;
  p.out = w.logits;
  p.epi = TC_EPI_F32;
  TRY(tok_gemm(w.r3, kD, w.Mt, pw.gen, kGenPad, p, s));
  {
    ProfScope _ps("focus_wce", s);
    FOCR_CHECK_CUDA(cudaMemsetAsync(w.dlogits, 0, (size_t)w.Mt * kGenPad * 2, s));
    wce_kernel<<<B, 128, 0, s>>>(w.logits, length, text_gt, table, w.dlogits, sr_pred_out, w.ce_partial, B, T, n_class,
                                 lam_gscale);
    FOCR_LAUNCH_CHECK();
  }
  // ---- backward to dctx
  p = gp();
  p.out = w.dr3;
  TRY(tok_gemm(w.dlogits, kGenPad, w.Mt, pw.genT, kD, p, s));
  ln1024_bwd_kernel<<<focr_cdiv(w.Mt, 8), 256, 0, s>>>(w.dr3, w.x3, pw.ln3a, w.dx3, w.Mt, 1e-6f);
  FOCR_LAUNCH_CHECK();
  p = gp();
  p.out = w.dh;
  p.gate = w.hff;
  TRY(tok_gemm(w.dx3, kD, w.Mt, pw.ff2T, kFF, p, s));
  p = gp();
  p.out = w.dr2;
  p.residual = w.dx3;
  TRY(tok_gemm(w.dh, kFF, w.Mt, pw.ff1T, kD, p, s));
  ln1024_bwd_kernel<<<focr_cdiv(w.Mt, 8), 256, 0, s>>>(w.dr2, w.x2, pw.ln2a, w.dx2, w.Mt, 1e-6f);
  FOCR_LAUNCH_CHECK();
  p = gp();
  p.out = w.dctx;
  TRY(tok_gemm(w.dx2, kD, w.Mt, pw.xoT, kD, p, s));
  return FOCR_OK;
}

}  // namespace strokenet

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int focr_strokenet_num_slots(void) { return strokenet::num_slots(); }

const char* focr_strokenet_slot_name(int variant, int idx) {
  static thread_local std::vector<std::string> names[2];
  if (variant < 0 || variant > 1) return nullptr;
  if (names[variant].empty()) names[variant] = strokenet::slot_names(variant);
  if (idx < 0 || idx >= (int)names[variant].size()) return nullptr;
  return names[variant][idx].c_str();
}

size_t focr_strokenet_prepared_bytes(int n_class) {
  strokenet::Prep p;
  strokenet::prep_layout(p, n_class, nullptr);
  return p.total;
}

int focr_strokenet_prepare(void* const* params, int n_class, void* prepared, size_t prepared_bytes, void* stream) {
  FOCR_REQUIRE(n_class >= 1 && n_class <= 4096, "strokenet_prepare: n_class %d", n_class);
  strokenet::Prep p;
  strokenet::prep_layout(p, n_class, prepared);
  FOCR_REQUIRE(prepared && prepared_bytes >= p.total, "strokenet_prepare: buffer too small (%zu < %zu)", prepared_bytes,
               p.total);
  return strokenet::prepare(params, n_class, p, (cudaStream_t)stream);
}

size_t focr_focus_loss_workspace_bytes(int B, int T) {
  strokenet::Ws w;
  strokenet::ws_layout(w, B, T, nullptr);
  return w.total;
}

int focr_focus_loss_ws_tensor(int B, int T, const char* name, long long* byte_offset, long long* elems, int* elem_bytes) {
  using namespace strokenet;
  Ws w;
  char* base = reinterpret_cast<char*>(4096);
  ws_layout(w, B, T, base);
  const Topo& t = topo();
  const void* ptr = nullptr;
  long long n = 0;
  int eb = 2;
  const std::string nm = name;
  if (nm == "query") ptr = w.query, n = w.Mt * kD;
  else if (nm == "Q") ptr = w.Q, n = w.Mt * kD;
  else if (nm == "text") ptr = w.text, n = w.Mt * kD;
  else if (nm == "a1") ptr = w.a1, n = (long long)B * 4096 * 64;
  else if (nm == "a2") ptr = w.a2, n = (long long)B * 1024 * 128;
  else if (nm == "p2") ptr = w.p2, n = (long long)B * 256 * 128;
  else if (nm == "feat") ptr = w.act[t.stages[3].conv], n = (long long)B * kTok * kD;
  else if (nm == "logits") ptr = w.logits, n = w.Mt * kGenPad, eb = 4;
  else if (nm == "ctx") ptr = w.ctx, n = w.Mt * kD;
  else if (nm == "r2") ptr = w.r2, n = w.Mt * kD;
  else if (nm == "r3") ptr = w.r3, n = w.Mt * kD;
  else if (nm == "x2") ptr = w.x2, n = w.Mt * kD;
  else if (nm == "x3") ptr = w.x3, n = w.Mt * kD;
  else if (nm == "hff") ptr = w.hff, n = w.Mt * kFF;
  else if (nm == "V") ptr = w.Vp, n = (long long)B * kTok * kD;
  else if (nm == "K") ptr = w.Kp, n = (long long)B * kTok * kD;
  else if (nm == "map_hr") ptr = w.map_hr, n = (long long)B * kHeads * T * kTok, eb = 4;
  else if (nm == "map_sr") ptr = w.map_sr, n = (long long)B * kHeads * T * kTok, eb = 4;
  else if (nm.rfind("act", 0) == 0) {
    const int i = atoi(name + 3);
    FOCR_REQUIRE(i >= 2 && i < (int)t.convs.size(), "focus_loss_ws_tensor: %s", name);
    ptr = w.act[i], n = (long long)B * kTok * t.convs[i].cout;
  } else {
    focr_set_error("focus_loss_ws_tensor: unknown tensor %s", name);
    return FOCR_ERR_INVALID;
  }
  *byte_offset = (const char*)ptr - base;
  *elems = n;
  *elem_bytes = eb;
  return FOCR_OK;
}

// StrokeFocusLoss.forward (TG/loss/stroke_focus_loss.py:83-122) / the attention-map part of TextFocusLoss.forward
// (STT/loss/text_focus_loss.py:86-99), value and gradient in one call.
int focr_focus_loss(const void* prepared, size_t prepared_bytes, int n_class, const float* sr, const float* hr,
                    const long long* text_input, int B, int T, float lambda, float gscale, float* d_sr, float* losses,
                    float* map_hr_out, float* map_sr_out, void* ws, size_t ws_bytes, void* stream) {
  using namespace strokenet;
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(B >= 1 && T >= 1 && T <= kMaxT, "focus_loss: B %d, T %d (T <= %d)", B, T, kMaxT);
  Prep pw;
  prep_layout(pw, n_class, const_cast<void*>(prepared));
  FOCR_REQUIRE(prepared && prepared_bytes >= pw.total, "focus_loss: prepared blob too small");
  Ws w;
  ws_layout(w, B, T, ws);
  FOCR_REQUIRE(ws && ws_bytes >= w.total, "focus_loss: workspace too small (%zu < %zu)", ws_bytes, w.total);
  const long n = (long)B * 3 * 32 * 128;
  TRY(focr_mse_loss_grad(sr, hr, d_sr, losses + 1, n, gscale, w.partial, ((size_t)B * kHeads + 4096) * 4, stream));
  TRY(text_side(pw, text_input, B, T, n_class, w, s));
  TRY(branch_forward(pw, hr, B, T, w, w.map_hr, s));
  TRY(branch_forward(pw, sr, B, T, w, w.map_sr, s));
  const double numel = (double)B * kHeads * T * kTok;
  {
    ProfScope _ps("focus_xattn", s);
    xattn_map_bwd_kernel<<<B * kHeads, 256, 0, s>>>(w.map_hr, w.map_sr, w.Q, w.g[1], w.partial, T,
                                                    (float)((double)lambda * gscale / numel));
    FOCR_LAUNCH_CHECK();
    finish_loss_kernel<<<1, 32, 0, s>>>(w.partial, B * kHeads, (float)(1.0 / numel), lambda, losses);
    FOCR_LAUNCH_CHECK();
  }
  TRY(branch_backward(pw, B, w, d_sr, false, s));
  if (map_hr_out)
    FOCR_CHECK_CUDA(cudaMemcpyAsync(map_hr_out, w.map_hr, (size_t)numel * 4, cudaMemcpyDeviceToDevice, s));
  if (map_sr_out)
    FOCR_CHECK_CUDA(cudaMemcpyAsync(map_sr_out, w.map_sr, (size_t)numel * 4, cudaMemcpyDeviceToDevice, s));
  return FOCR_OK;
}

// TextFocusLoss.forward with text_focus on (STT/loss/text_focus_loss.py:84-99): mse + lambda_attn * L1(maps) +
// lambda_ce * weight_cross_entropy(sr logits, text_gt)  (STT/loss/weight_ce_loss.py:36-45), value and gradient.
int focr_text_focus_loss(const void* prepared, size_t prepared_bytes, int n_class, const float* sr, const float* hr,
                         const long long* text_input, const long long* length, const long long* text_gt,
                         const float* weight_table, int B, int T, float lambda_attn, float lambda_ce, float gscale,
                         float* d_sr, float* losses, float* map_hr_out, float* map_sr_out, float* sr_pred_out, void* ws,
                         size_t ws_bytes, void* stream) {
  using namespace strokenet;
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(B >= 1 && T >= 1 && T <= kMaxT, "text_focus_loss: B %d, T %d (T <= %d)", B, T, kMaxT);
  FOCR_REQUIRE(n_class >= 2 && n_class <= kGenPad, "text_focus_loss: n_class %d", n_class);
  FOCR_REQUIRE(length && text_gt && weight_table, "text_focus_loss: length / text_gt / weight_table are required");
  Prep pw;
  prep_layout(pw, n_class, const_cast<void*>(prepared));
  FOCR_REQUIRE(prepared && prepared_bytes >= pw.total, "text_focus_loss: prepared blob too small");
  Ws w;
  ws_layout(w, B, T, ws);
  FOCR_REQUIRE(ws && ws_bytes >= w.total, "text_focus_loss: workspace too small (%zu < %zu)", ws_bytes, w.total);
  const long n = (long)B * 3 * 32 * 128;
  TRY(focr_mse_loss_grad(sr, hr, d_sr, losses + 1, n, gscale, w.partial, ((size_t)B * kHeads + 4096) * 4, stream));
  TRY(text_side(pw, text_input, B, T, n_class, w, s));
  TRY(branch_forward(pw, hr, B, T, w, w.map_hr, s));
  TRY(branch_forward(pw, sr, B, T, w, w.map_sr, s));
  TRY(decoder_tail(pw, length, text_gt, weight_table, B, T, n_class, lambda_ce * gscale, sr_pred_out, w, s));
  const double numel = (double)B * kHeads * T * kTok;
  {
    ProfScope _ps("focus_xattn", s);
    const size_t smem = (size_t)kTok * 33 * 4 + 2 * 8 * kTok * 4 + 2 * 8 * 64 * 4;
    static bool attr = false;
    if (!attr) {
      FOCR_CHECK_CUDA(cudaFuncSetAttribute(xattn_full_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    xattn_full_bwd_kernel<<<B * kHeads, 256, smem, s>>>(w.map_hr, w.map_sr, w.Q, w.Vp, w.dctx, w.g[1], w.g[2], w.partial, T,
                                                        (float)((double)lambda_attn * gscale / numel));
    FOCR_LAUNCH_CHECK();
    finish_loss4_kernel<<<1, 32, 0, s>>>(w.partial, B * kHeads, (float)(1.0 / numel), lambda_attn, w.ce_partial, B,
                                         lambda_ce, losses);
    FOCR_LAUNCH_CHECK();
  }
  TRY(branch_backward(pw, B, w, d_sr, true, s));
  if (map_hr_out)
    FOCR_CHECK_CUDA(cudaMemcpyAsync(map_hr_out, w.map_hr, (size_t)numel * 4, cudaMemcpyDeviceToDevice, s));
  if (map_sr_out)
    FOCR_CHECK_CUDA(cudaMemcpyAsync(map_sr_out, w.map_sr, (size_t)numel * 4, cudaMemcpyDeviceToDevice, s));
  return FOCR_OK;
}

}  // extern "C"
