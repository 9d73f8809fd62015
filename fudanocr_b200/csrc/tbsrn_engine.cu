// TBSRN forward / backward orchestration on the hand-written kernels (see tbsrn_engine.cuh).
// Host code only: every launch goes to the caller's stream, no synchronisation, no allocation
// (all buffers come out of the caller-provided workspace), so the whole step is CUDA-graph capturable.
#include "tbsrn_engine.cuh"

#include <string.h>

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

namespace tbsrn {

// ---------------------------------------------------------------------------------------------
// slot names = reference state_dict keys
// ---------------------------------------------------------------------------------------------
std::vector<std::string> slot_names(int n, int arch) {
  Slots sl(n, arch);
  std::vector<std::string> v(sl.count);
  v[sl.b1_w] = "block1.0.weight";
  v[sl.b1_b] = "block1.0.bias";
  v[sl.b1_a] = "block1.1.weight";
  static const char* srb_names[S_COUNT] = {
      "conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var",
      "bn1.num_batches_tracked", "conv2.weight", "conv2.bias", "bn2.weight", "bn2.bias", "bn2.running_mean",
      "bn2.running_var", "bn2.num_batches_tracked",
      "feature_enhancer.multihead.linears.0.weight", "feature_enhancer.multihead.linears.0.bias",
      "feature_enhancer.multihead.linears.1.weight", "feature_enhancer.multihead.linears.1.bias",
      "feature_enhancer.multihead.linears.2.weight", "feature_enhancer.multihead.linears.2.bias",
      "feature_enhancer.multihead.linears.3.weight", "feature_enhancer.multihead.linears.3.bias",
      "feature_enhancer.mul_layernorm1.a_2", "feature_enhancer.mul_layernorm1.b_2",
      "feature_enhancer.pff.w_1.weight", "feature_enhancer.pff.w_1.bias", "feature_enhancer.pff.w_2.weight",
      "feature_enhancer.pff.w_2.bias", "feature_enhancer.mul_layernorm3.a_2", "feature_enhancer.mul_layernorm3.b_2",
      "feature_enhancer.linear.weight", "feature_enhancer.linear.bias"};
  static const char* gru_names[G_COUNT] = {"conv1.weight", "conv1.bias", "gru.weight_ih_l0", "gru.weight_hh_l0",
                                            "gru.bias_ih_l0", "gru.bias_hh_l0", "gru.weight_ih_l0_reverse",
                                            "gru.weight_hh_l0_reverse", "gru.bias_ih_l0_reverse",
                                            "gru.bias_hh_l0_reverse"};
  static const char* tsrn_head[7] = {"conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias", "bn1.running_mean",
                                     "bn1.running_var", "bn1.num_batches_tracked"};
  static const char* tsrn_mid[7] = {"conv2.weight", "conv2.bias", "bn2.weight", "bn2.bias", "bn2.running_mean",
                                    "bn2.running_var", "bn2.num_batches_tracked"};
  for (int b = 0; b < n; ++b) {
    const std::string pre = "block" + std::to_string(b + 2) + ".";
    if (arch == ARCH_TSRN) {
      for (int j = 0; j < 7; ++j) v[sl.srb(b, TS_C1W + j)] = pre + tsrn_head[j];
      for (int j = 0; j < G_COUNT; ++j) v[sl.srb(b, TS_G1 + j)] = pre + "gru1." + gru_names[j];
      for (int j = 0; j < 7; ++j) v[sl.srb(b, TS_C2W + j)] = pre + tsrn_mid[j];
      for (int j = 0; j < G_COUNT; ++j) v[sl.srb(b, TS_G2 + j)] = pre + "gru2." + gru_names[j];
    } else {
      for (int s = 0; s < S_COUNT; ++s) v[sl.srb(b, s)] = pre + srb_names[s];
    }
  }
  static const char* bn_names[5] = {"weight", "bias", "running_mean", "running_var", "num_batches_tracked"};
  const std::string b7 = "block" + std::to_string(n + 2), b8 = "block" + std::to_string(n + 3);
  v[sl.b7_w] = b7 + ".0.weight";
  v[sl.b7_b] = b7 + ".0.bias";
  for (int j = 0; j < 5; ++j) v[sl.b7_bn + j] = b7 + ".1." + bn_names[j];
  v[sl.up_w] = b8 + ".0.conv.weight";
  v[sl.up_b] = b8 + ".0.conv.bias";
  v[sl.fin_w] = b8 + ".1.weight";
  v[sl.fin_b] = b8 + ".1.bias";
  for (int c = 0; c < 6; ++c) {
    const std::string p = "stn_head.stn_convnet." + std::to_string(2 * c);
    v[sl.stn(c, 0)] = p + ".0.weight";
    v[sl.stn(c, 1)] = p + ".0.bias";
    for (int j = 0; j < 5; ++j) v[sl.stn(c, 2 + j)] = p + ".1." + bn_names[j];
  }
  v[sl.fc1_w] = "stn_head.stn_fc1.0.weight";
  v[sl.fc1_b] = "stn_head.stn_fc1.0.bias";
  for (int j = 0; j < 5; ++j) v[sl.bn1d + j] = std::string("stn_head.stn_fc1.1.") + bn_names[j];
  v[sl.fc2_w] = "stn_head.stn_fc2.weight";
  v[sl.fc2_b] = "stn_head.stn_fc2.bias";
  v[sl.tps_inv] = "tps.inverse_kernel";
  v[sl.tps_repr] = "tps.target_coordinate_repr";
  return v;
}

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
namespace {
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
inline long pad128(long m) { return (m + 127) / 128 * 128; }
}  // namespace

void layout(Ws& w, int B, int n, void* base, int arch) {
  Bump b{reinterpret_cast<char*>(base)};
  w.B = B;
  w.srb_nums = n;
  w.arch = arch;
  const long T = (long)B * 1024, Thr = (long)B * 4096;
  w.T = T;
  w.Thr = Thr;
  const long Bp = pad128(B);
  w.pe = b.get<bf16>(1024 * 64);
  for (int i = 0; i < 6; ++i) {
    const StnConv& c = kStn[i];
    const long M = (long)B * c.h * c.w, Mp = pad128(M);
    w.stn_col[i] = b.get<bf16>(Mp * c.kpad);
    w.stn_ypre[i] = b.get<bf16>(Mp * c.npad);
    w.stn_yact[i] = b.get<bf16>(Mp * c.cout > Bp * 512 ? Mp * c.cout : Bp * 512);
    w.stn_pool[i] = b.get<bf16>(Mp * c.cout);
    w.stn_stats[i] = b.get<float>(4 * 256);
    w.stn_bias[i] = b.get<float>(256);
    w.stn_wf[i] = b.get<bf16>((long)c.npad * c.kpad);
    w.stn_wt[i] = b.get<bf16>((long)c.npad * c.kpad);
    w.stn_dypre[i] = b.get<bf16>(Mp * c.npad);
    w.stn_dyact[i] = b.get<bf16>(Mp * c.cout);
    w.stn_dcol[i] = b.get<bf16>(Mp * c.kpad);
    w.stn_dpool[i] = b.get<bf16>(Mp * c.cout);
  }
  w.fc1_wf = b.get<bf16>(512 * 512);
  w.fc1_wt = b.get<bf16>(512 * 512);
  w.fc2_wf = b.get<bf16>(64 * 512);
  w.fc2_wt = b.get<bf16>(64 * 512);
  w.fc2_bias = b.get<float>(64);
  w.f1pre = b.get<bf16>(Bp * 512);
  w.f1 = b.get<bf16>(Bp * 512);
  w.bn1d_stats = b.get<float>(4 * 512);
  w.ctrl = b.get<float>(Bp * 64);
  w.x_tps = b.get<float>((long)B * 3 * 1024);
  w.a1x = b.get<bf16>(T * 64);
  w.b1pre = b.get<bf16>(T * 64);
  w.b1 = b.get<bf16>(T * 64);
  w.w_b1 = b.get<bf16>(9 * 4096);
  w.w_b1d = b.get<bf16>(9 * 4096);
  w.srb.resize(n);
  w.srbw.resize(n);
  w.gruw.resize(arch == ARCH_TSRN ? 2 * n : 0);
  for (int i = 0; i < n; ++i) {
    SrbWs& s = w.srb[i];
    SrbW& q = w.srbw[i];
    s.c1 = b.get<bf16>(T * 64);
    s.a1 = b.get<bf16>(T * 64);
    s.c2 = b.get<bf16>(T * 64);
    s.out = b.get<bf16>(T * 64);
    s.st1 = b.get<float>(4 * 64);
    s.st2 = b.get<float>(4 * 64);
    q.c1f = b.get<bf16>(9 * 4096);
    q.c1d = b.get<bf16>(9 * 4096);
    q.c2f = b.get<bf16>(9 * 4096);
    q.c2d = b.get<bf16>(9 * 4096);
    if (arch == ARCH_TSRN) {
      s.r0 = b.get<bf16>(T * 64);
      s.g1in = b.get<bf16>(T * 64);
      s.xp1 = b.get<bf16>(T * 192);
      s.o1 = b.get<bf16>(T * 64);
      s.hp1b = b.get<bf16>(T * 64);
      s.hp1 = b.get<float>(T * 64);
      s.ssum = b.get<bf16>(T * 64);
      s.g2in = b.get<bf16>(T * 64);
      s.xp2 = b.get<bf16>(T * 192);
      s.hp2b = b.get<bf16>(T * 64);
      s.hp2 = b.get<float>(T * 64);
      for (int gidx = 0; gidx < 2; ++gidx) {
        GruW& gw = w.gruw[2 * i + gidx];
        gw.cw = b.get<bf16>(64 * 64);
        gw.cwT = b.get<bf16>(64 * 64);
        gw.wih = b.get<bf16>(192 * 64);
        gw.wihT = b.get<bf16>(192 * 64);
        gw.bih = b.get<float>(192);
      }
      continue;
    }
    s.f = b.get<bf16>(T * 128);
    s.qkv = b.get<bf16>(T * 384);
    s.o = b.get<bf16>(T * 128);
    s.y1pre = b.get<bf16>(T * 128);
    s.y1 = b.get<bf16>(T * 128);
    s.hd = b.get<bf16>(T * 128);
    s.y2pre = b.get<bf16>(T * 128);
    s.y2 = b.get<bf16>(T * 128);
    s.lse = b.get<float>((long)B * 4 * 1024);
    s.dropbits = b.get<uint32_t>((long)(attn_drop_bits_bytes(B) / 4));
    q.qkv = b.get<bf16>(384 * 128);
    q.qkvT = b.get<bf16>(384 * 128);
    q.wo = b.get<bf16>(128 * 128);
    q.woT = b.get<bf16>(128 * 128);
    q.w1 = b.get<bf16>(128 * 128);
    q.w1T = b.get<bf16>(128 * 128);
    q.w2 = b.get<bf16>(128 * 128);
    q.w2T = b.get<bf16>(128 * 128);
    q.lin = b.get<bf16>(64 * 128);
    q.linT = b.get<bf16>(64 * 128);
    q.bqkv = b.get<float>(384);
  }
  w.c7 = b.get<bf16>(T * 64);
  w.s7 = b.get<bf16>(T * 64);
  w.st7 = b.get<float>(4 * 64);
  w.w7f = b.get<bf16>(9 * 4096);
  w.w7d = b.get<bf16>(9 * 4096);
  w.wupf = b.get<bf16>(9 * 256 * 64);
  w.wupd = b.get<bf16>(9 * 256 * 64);
  w.wfin = b.get<bf16>(9 * 4096);
  w.wfind = b.get<bf16>(9 * 4096);
  w.bup = b.get<float>(256);
  w.upre = b.get<bf16>(Thr * 64);
  w.u = b.get<bf16>(Thr * 64);
  w.z = b.get<float>(Thr * 64);
  w.opre = b.get<float>((long)B * 3 * 4096);
  w.sr = b.get<float>((long)B * 3 * 4096);
  w.d_o = b.get<float>((long)B * 3 * 4096);
  w.a1d = b.get<bf16>(Thr * 64);
  w.du = b.get<bf16>(Thr * 64);
  w.dupre = b.get<bf16>(Thr * 64);
  for (int i = 0; i < 5; ++i) w.g64[i] = b.get<bf16>(T * 64);
  for (int i = 0; i < 4; ++i) w.g128[i] = b.get<bf16>(T * 128);
  w.g384 = b.get<bf16>(T * 384);
  w.gdxp = w.g384;  // TSRN reuses the (T,384) buffer as two (T,192) halves
  w.gdhid = w.g384 + T * 192;
  w.dsum = b.get<float>((long)B * 4 * 1024);
  w.dx_tps = b.get<float>((long)B * 3 * 1024);
  w.dctrl = b.get<float>((long)B * 40);
  w.dctrl_b = b.get<bf16>(Bp * 64);
  w.df1 = b.get<bf16>(Bp * 512);
  w.df1pre = b.get<bf16>(Bp * 512);
  w.dfeat = b.get<bf16>(Bp * 512);
  // scratch for reductions: the largest users are the conv wgrad (148 CTAs x 9*64*64 fp32, x2 for the
  // pixel-shuffle variant) and the linear wgrad (<= 148 x 128 x 128 fp32 per 128-row block)
  w.partial_bytes = (size_t)48 << 20;
  w.partial = b.get<float>(w.partial_bytes / 4);
  w.coef = b.get<float>(2 * 2048);
  w.tmpw = b.get<float>(256 * 2304);
  w.tmpb = b.get<float>(1024);
  w.total_bytes = (b.off + 255) & ~(size_t)255;
}

// ---------------------------------------------------------------------------------------------
// GEMM helpers
// ---------------------------------------------------------------------------------------------
namespace {

TcGemmParams gp() {
  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.kh = p.kw = 1;
  p.epi = TC_EPI_BF16;
  return p;
}
// token-matrix GEMM: out[M,N] = a[M,K] w[N,K]^T (+...)  (M padded to 128 by the caller)
int tok_gemm(const bf16* a, int K, long M, const bf16* w, int N, TcGemmParams p, cudaStream_t s) {
  p.n_total = N;
  p.W = 64;
  p.H = 2;
  if (p.ldc == 0) p.ldc = N;
  const bf16* ap[1] = {a};
  return tc_gemm_launch(ap, 1, K, (long)64 * K, (long)128 * K, K, (int)(M / 128), w, K, p, s);
}
// conv over a 64-channel NHWC map (B,H,W,64)
int map_conv(const bf16* a, int B, int H, int W, int kh, int kw, const bf16* w, int N, TcGemmParams p,
             cudaStream_t s) {
  p.n_total = N;
  p.W = W;
  p.H = H;
  p.kh = kh;
  p.kw = kw;
  if (p.ldc == 0) p.ldc = N;
  const bf16* ap[1] = {a};
  return tc_gemm_launch(ap, 1, 64, (long)W * 64, (long)H * W * 64, 64, B, w, 64, p, s);
}
int d2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  FOCR_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
  return FOCR_OK;
}
template <typename T>
T* P(void* const* prm, int i) {
  return reinterpret_cast<T*>(prm[i]);
}
uint32_t thresh_of(float p) {
  if (p <= 0.f) return 0;
  long t = (long)(p * 65536.0 + 0.5);
  if (t > 65535) t = 65535;
  return (uint32_t)t;
}

int prep_weights(const Slots& sl, void* const* prm, Ws& w, bool stn, cudaStream_t s) {
  TRY(pe_table(w.pe, s));
  TRY(prep_w9(P<float>(prm, sl.b1_w), w.w_b1, 0, s));
  TRY(prep_w9(P<float>(prm, sl.b1_w), w.w_b1d, 3, s));
  for (int i = 0; i < sl.srb_nums; ++i) {
    SrbW& q = w.srbw[i];
    const int c1w = sl.arch == ARCH_TSRN ? (int)TS_C1W : (int)S_C1W, c2w = sl.arch == ARCH_TSRN ? (int)TS_C2W : (int)S_C2W;
    TRY(prep_conv_w_fwd(P<float>(prm, sl.srb(i, c1w)), q.c1f, 64, 64, 3, 0, s));
    TRY(prep_conv_w_dgrad(P<float>(prm, sl.srb(i, c1w)), q.c1d, 64, 64, 3, 0, s));
    TRY(prep_conv_w_fwd(P<float>(prm, sl.srb(i, c2w)), q.c2f, 64, 64, 3, 0, s));
    TRY(prep_conv_w_dgrad(P<float>(prm, sl.srb(i, c2w)), q.c2d, 64, 64, 3, 0, s));
    if (sl.arch == ARCH_TSRN) {
      for (int gidx = 0; gidx < 2; ++gidx) {
        const int g0 = sl.srb(i, gidx ? TS_G2 : TS_G1);
        GruW& gw = w.gruw[2 * i + gidx];
        TRY(prep_linear_w(P<float>(prm, g0 + G_CW), gw.cw, gw.cwT, 64, 64, 64, 0, s));
        TRY(gru_prep(P<float>(prm, g0 + G_WIH), P<float>(prm, g0 + G_WIH_R), P<float>(prm, g0 + G_BIH),
                     P<float>(prm, g0 + G_BIH_R), gw.wih, gw.wihT, gw.bih, s));
      }
      continue;
    }
    for (int j = 0; j < 3; ++j) {
      TRY(prep_linear_w(P<float>(prm, sl.srb(i, S_LQW + 2 * j)), q.qkv + j * 128 * 128, q.qkvT, 128, 128, 384,
                        j * 128, s));
      TRY(d2d(q.bqkv + j * 128, prm[sl.srb(i, S_LQB + 2 * j)], 128 * 4, s));
    }
    TRY(prep_linear_w(P<float>(prm, sl.srb(i, S_LOW)), q.wo, q.woT, 128, 128, 128, 0, s));
    TRY(prep_linear_w(P<float>(prm, sl.srb(i, S_W1W)), q.w1, q.w1T, 128, 128, 128, 0, s));
    TRY(prep_linear_w(P<float>(prm, sl.srb(i, S_W2W)), q.w2, q.w2T, 128, 128, 128, 0, s));
    TRY(prep_linear_w(P<float>(prm, sl.srb(i, S_LINW)), q.lin, q.linT, 64, 128, 64, 0, s));
  }
  TRY(prep_conv_w_fwd(P<float>(prm, sl.b7_w), w.w7f, 64, 64, 3, 0, s));
  TRY(prep_conv_w_dgrad(P<float>(prm, sl.b7_w), w.w7d, 64, 64, 3, 0, s));
  TRY(prep_conv_w_fwd(P<float>(prm, sl.up_w), w.wupf, 256, 64, 3, 1, s));
  TRY(prep_conv_w_dgrad(P<float>(prm, sl.up_w), w.wupd, 256, 64, 3, 1, s));
  TRY(prep_bias_shuf(P<float>(prm, sl.up_b), w.bup, 256, s));
  TRY(prep_w9(P<float>(prm, sl.fin_w), w.wfin, 2, s));
  TRY(prep_w9(P<float>(prm, sl.fin_w), w.wfind, 1, s));
  if (stn) {
    for (int i = 0; i < 6; ++i)
      TRY(prep_stn_conv_w(P<float>(prm, sl.stn(i, 0)), w.stn_wf[i], w.stn_wt[i], kStn[i].cout, kStn[i].cin,
                          kStn[i].npad, kStn[i].kpad, s));
    for (int i = 0; i < 6; ++i) {
      FOCR_CHECK_CUDA(cudaMemsetAsync(w.stn_bias[i], 0, 256 * 4, s));
      TRY(d2d(w.stn_bias[i], prm[sl.stn(i, 1)], kStn[i].cout * 4, s));
    }
    TRY(prep_fc1(P<float>(prm, sl.fc1_w), w.fc1_wf, w.fc1_wt, s));
    TRY(prep_fc2(P<float>(prm, sl.fc2_w), w.fc2_wf, w.fc2_wt, s));
    FOCR_CHECK_CUDA(cudaMemsetAsync(w.fc2_bias, 0, 64 * 4, s));
    TRY(d2d(w.fc2_bias, prm[sl.fc2_b], 40 * 4, s));
  }
  return FOCR_OK;
}

// BatchNorm forward on (T, C) with row stride ld: train -> batch stats (+ running update), eval -> running
int bn_fwd_stats(const bf16* x, long ld, long T, int C, void* const* prm, int bn_slot, bool training, Ws& w,
                 float* stats, cudaStream_t s) {
  if (training)
    return bn_train_stats(x, ld, T, C, P<float>(prm, bn_slot + BN_W), P<float>(prm, bn_slot + BN_B),
                          P<float>(prm, bn_slot + BN_RM), P<float>(prm, bn_slot + BN_RV),
                          P<long long>(prm, bn_slot + BN_NBT), 1e-5f, 0.1f, w.partial, stats, s);
  return bn_eval_stats(P<float>(prm, bn_slot + BN_W), P<float>(prm, bn_slot + BN_B), P<float>(prm, bn_slot + BN_RM),
                       P<float>(prm, bn_slot + BN_RV), 1e-5f, C, stats, s);
}


// ---------------------------------------------------------------------------------------------
// STN head + TPS  (train-only prologue)
// ---------------------------------------------------------------------------------------------
int stn_forward(const Slots& sl, void* const* prm, const float* x_lr, Ws& w, cudaStream_t s) {
  const int B = w.B;
  const long Bp = pad128(B);
  for (int i = 0; i < 6; ++i) {
    const StnConv& c = kStn[i];
    const long M = (long)B * c.h * c.w, Mp = pad128(M);
    if (i == 0)
      TRY(im2col3x3(nullptr, x_lr, w.stn_col[0], B, c.h, c.w, c.cin, c.kpad, s));
    else
      TRY(im2col3x3(w.stn_pool[i - 1], nullptr, w.stn_col[i], B, c.h, c.w, c.cin, c.kpad, s));
    TcGemmParams p = gp();
    p.bias = w.stn_bias[i];
    p.out = w.stn_ypre[i];
    TRY(tok_gemm(w.stn_col[i], c.kpad, Mp, w.stn_wf[i], c.npad, p, s));
    TRY(bn_fwd_stats(w.stn_ypre[i], c.npad, M, c.cout, prm, sl.stn(i, 2), true, w, w.stn_stats[i], s));
    TRY(bn_apply(w.stn_ypre[i], c.npad, w.stn_stats[i], w.stn_yact[i], c.cout, M, c.cout, ACT_RELU, nullptr, 0,
                 nullptr, s));
    if (c.pool_h) TRY(maxpool_fwd(w.stn_yact[i], w.stn_pool[i], B, c.h, c.w, c.cout, c.pool_h, s));
  }
  // flatten: yact[5] is (B,1,2,256) NHWC = (B,512) with column w*256+c (fc1 weight columns permuted to match)
  {
    TcGemmParams p = gp();
    p.bias = P<float>(prm, sl.fc1_b);
    p.out = w.f1pre;
    TRY(tok_gemm(w.stn_yact[5], 512, Bp, w.fc1_wf, 512, p, s));
    TRY(bn_fwd_stats(w.f1pre, 512, B, 512, prm, sl.bn1d, true, w, w.bn1d_stats, s));
    TRY(bn_apply(w.f1pre, 512, w.bn1d_stats, w.f1, 512, B, 512, ACT_RELU, nullptr, 0, nullptr, s));
    TcGemmParams q = gp();
    q.bias = w.fc2_bias;
    q.out = w.ctrl;
    q.epi = TC_EPI_F32;
    TRY(tok_gemm(w.f1, 512, Bp, w.fc2_wf, 64, q, s));
  }
  return tps_forward(x_lr, w.ctrl, 64, P<float>(prm, sl.tps_inv), P<float>(prm, sl.tps_repr), w.x_tps, B, s);
}

int stn_backward(const Slots& sl, void* const* prm, void* const* grd, const float* x_lr, Ws& w, cudaStream_t s) {
  const int B = w.B;
  const long Bp = pad128(B);
  TRY(tps_backward(x_lr, w.ctrl, 64, P<float>(prm, sl.tps_inv), P<float>(prm, sl.tps_repr), w.dx_tps, w.dctrl, 40, B,
                   s));
  TRY(f32_to_bf16_pad(w.dctrl, 40, B, w.dctrl_b, 64, Bp, s));
  // fc2 (input 0.1*f1): dW2 = 0.1 dC^T f1, db2 = colsum(dC), df1 = dC (0.1 W2)
  TRY(linear_wgrad(w.dctrl_b, 64, w.f1, 512, B, 64, 512, w.tmpw, 0.1f, w.partial, s));
  if (grd[sl.fc2_w]) TRY(d2d(grd[sl.fc2_w], w.tmpw, 40 * 512 * 4, s));
  TRY(colsum(w.dctrl_b, 64, B, 64, w.tmpb, w.partial, s));
  if (grd[sl.fc2_b]) TRY(d2d(grd[sl.fc2_b], w.tmpb, 40 * 4, s));
  {
    TcGemmParams p = gp();
    p.out = w.df1;
    TRY(tok_gemm(w.dctrl_b, 64, Bp, w.fc2_wt, 512, p, s));
  }
  TRY(bn_backward(w.df1, 512, w.f1pre, 512, w.bn1d_stats, w.df1pre, 512, B, 512, ACT_RELU,
                  P<float>(grd, sl.bn1d + BN_W), P<float>(grd, sl.bn1d + BN_B), w.partial, w.coef, s));
  TRY(linear_wgrad(w.df1pre, 512, w.stn_yact[5], 512, B, 512, 512, w.tmpw, 1.f, w.partial, s));
  if (grd[sl.fc1_w]) TRY(unperm_fc1_grad(w.tmpw, P<float>(grd, sl.fc1_w), s));
  if (grd[sl.fc1_b]) TRY(colsum(w.df1pre, 512, B, 512, P<float>(grd, sl.fc1_b), w.partial, s));
  {
    TcGemmParams p = gp();
    p.out = w.dfeat;
    TRY(tok_gemm(w.df1pre, 512, Bp, w.fc1_wt, 512, p, s));
  }
  const bf16* dnext = w.dfeat;  // gradient w.r.t. the (pooled) output of conv block i
  for (int i = 5; i >= 0; --i) {
    const StnConv& c = kStn[i];
    const long M = (long)B * c.h * c.w, Mp = pad128(M);
    const bf16* dyact = dnext;
    if (c.pool_h) {
      TRY(maxpool_bwd(w.stn_yact[i], w.stn_pool[i], dnext, w.stn_dyact[i], B, c.h, c.w, c.cout, c.pool_h, s));
      dyact = w.stn_dyact[i];
    }
    if (c.cout < c.npad) FOCR_CHECK_CUDA(cudaMemsetAsync(w.stn_dypre[i], 0, Mp * c.npad * 2, s));
    TRY(bn_backward(dyact, c.cout, w.stn_ypre[i], c.npad, w.stn_stats[i], w.stn_dypre[i], c.npad, M, c.cout, ACT_RELU,
                    P<float>(grd, sl.stn(i, 2)), P<float>(grd, sl.stn(i, 3)), w.partial, w.coef, s));
    TRY(linear_wgrad(w.stn_dypre[i], c.npad, w.stn_col[i], c.kpad, M, c.npad, c.kpad, w.tmpw, 1.f, w.partial, s));
    if (grd[sl.stn(i, 0)]) TRY(unpack_stn_conv_grad(w.tmpw, P<float>(grd, sl.stn(i, 0)), c.cout, c.cin, c.kpad, s));
    TRY(colsum(w.stn_dypre[i], c.npad, M, c.npad, w.tmpb, w.partial, s));
    if (grd[sl.stn(i, 1)]) TRY(d2d(grd[sl.stn(i, 1)], w.tmpb, c.cout * 4, s));
    if (i > 0) {
      TcGemmParams p = gp();
      p.out = w.stn_dcol[i];
      TRY(tok_gemm(w.stn_dypre[i], c.npad, Mp, w.stn_wt[i], c.kpad, p, s));
      TRY(col2im3x3(w.stn_dcol[i], w.stn_dpool[i - 1], B, c.h, c.w, c.cin, c.kpad, s));
      dnext = w.stn_dpool[i - 1];
    }
  }
  return FOCR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// TSRN sequence residual block (tsrn.py:77-98, 128-145)
// ---------------------------------------------------------------------------------------------
namespace {

// GruBlock forward: gin = conv1x1(x); xp = gin W_ih^T + b_ih (both directions); recurrence -> out (T,64)
int gru_block_forward(void* const* prm, int g0, const GruW& gw, const bf16* x, bf16* gin, bf16* xp, bf16* out, float* hp32,
                      bf16* hp16, int vertical, Ws& w, cudaStream_t s) {
  TcGemmParams p = gp();
  p.bias = P<float>(prm, g0 + G_CB);
  p.out = gin;
  TRY(tok_gemm(x, 64, w.T, gw.cw, 64, p, s));
  p = gp();
  p.bias = gw.bih;
  p.out = xp;
  TRY(tok_gemm(gin, 64, w.T, gw.wih, 192, p, s));
  return gru_forward(xp, P<float>(prm, g0 + G_WHH), P<float>(prm, g0 + G_WHH_R), P<float>(prm, g0 + G_BHH),
                     P<float>(prm, g0 + G_BHH_R), out, hp32, hp16, w.B, vertical, s);
}

// GruBlock backward: dout (T,64) -> dx (T,64) (+ residual), all parameter gradients of the block
int gru_block_backward(void* const* prm, void* const* grd, int g0, const GruW& gw, const bf16* x, const bf16* gin,
                       const bf16* xp, const float* hp32, const bf16* hp16, const bf16* dout, bf16* dgin, bf16* dx,
                       const bf16* dx_residual, int vertical, Ws& w, cudaStream_t s) {
  const long T = w.T;
  TRY(gru_backward(xp, P<float>(prm, g0 + G_WHH), P<float>(prm, g0 + G_WHH_R), P<float>(prm, g0 + G_BHH),
                   P<float>(prm, g0 + G_BHH_R), hp32, dout, w.gdxp, w.gdhid, w.B, vertical, s));
  // recurrent weights / biases: dW_hh = dhid^T h_{t-1}
  float* tmp2 = w.tmpw + 192 * 64;  // second half of the scratch: the [2N][128] intermediate of the k64 trick
  TRY(linear_wgrad_k64(w.gdhid, hp16, T, 192, w.tmpw, tmp2, w.partial, s));
  TRY(gru_unpack_whh(w.tmpw, P<float>(grd, g0 + G_WHH), P<float>(grd, g0 + G_WHH_R), s));
  TRY(colsum(w.gdhid, 192, T, 192, w.tmpb, w.partial, s));
  if (grd[g0 + G_BHH]) TRY(d2d(grd[g0 + G_BHH], w.tmpb, 96 * 4, s));
  if (grd[g0 + G_BHH_R]) TRY(d2d(grd[g0 + G_BHH_R], w.tmpb + 96, 96 * 4, s));
  // input projection
  TRY(linear_wgrad_k64(w.gdxp, gin, T, 192, w.tmpw, tmp2, w.partial, s));
  if (grd[g0 + G_WIH]) TRY(d2d(grd[g0 + G_WIH], w.tmpw, 96 * 64 * 4, s));
  if (grd[g0 + G_WIH_R]) TRY(d2d(grd[g0 + G_WIH_R], w.tmpw + 96 * 64, 96 * 64 * 4, s));
  TRY(colsum(w.gdxp, 192, T, 192, w.tmpb, w.partial, s));
  if (grd[g0 + G_BIH]) TRY(d2d(grd[g0 + G_BIH], w.tmpb, 96 * 4, s));
  if (grd[g0 + G_BIH_R]) TRY(d2d(grd[g0 + G_BIH_R], w.tmpb + 96, 96 * 4, s));
  TcGemmParams p = gp();
  p.out = dgin;
  TRY(tok_gemm(w.gdxp, 192, T, gw.wihT, 64, p, s));
  // conv1x1
  if (grd[g0 + G_CW]) TRY(linear_wgrad_k64(dgin, x, T, 64, P<float>(grd, g0 + G_CW), tmp2, w.partial, s));
  if (grd[g0 + G_CB]) TRY(colsum(dgin, 64, T, 64, P<float>(grd, g0 + G_CB), w.partial, s));
  p = gp();
  p.out = dx;
  p.residual = dx_residual;
  return tok_gemm(dgin, 64, T, gw.cwT, 64, p, s);
}

int tsrn_srb_forward(const Slots& sl, void* const* prm, int i, const bf16* x, Ws& w, bool training, cudaStream_t s) {
  SrbWs& a = w.srb[i];
  SrbW& q = w.srbw[i];
  const int B = w.B;
  const long T = w.T;
  TcGemmParams p = gp();
  p.bias = P<float>(prm, sl.srb(i, TS_C1B));
  p.out = a.c1;
  TRY(map_conv(x, B, 16, 64, 3, 3, q.c1f, 64, p, s));
  TRY(bn_fwd_stats(a.c1, 64, T, 64, prm, sl.srb(i, TS_BN1W), training, w, a.st1, s));
  TRY(bn_apply(a.c1, 64, a.st1, a.a1, 64, T, 64, ACT_MISH, nullptr, 0, nullptr, s));
  p = gp();
  p.bias = P<float>(prm, sl.srb(i, TS_C2B));
  p.out = a.c2;
  TRY(map_conv(a.a1, B, 16, 64, 3, 3, q.c2f, 64, p, s));
  TRY(bn_fwd_stats(a.c2, 64, T, 64, prm, sl.srb(i, TS_BN2W), training, w, a.st2, s));
  TRY(bn_apply(a.c2, 64, a.st2, a.r0, 64, T, 64, ACT_NONE, nullptr, 0, nullptr, s));
  // gru1 on the transposed map = sequences down the columns (tsrn.py:96)
  TRY(gru_block_forward(prm, sl.srb(i, TS_G1), w.gruw[2 * i], a.r0, a.g1in, a.xp1, a.o1, a.hp1, a.hp1b, 1, w, s));
  TRY(add_bf16(x, a.o1, a.ssum, T * 64, s));
  // gru2(x + residual): sequences along the rows (tsrn.py:98)
  return gru_block_forward(prm, sl.srb(i, TS_G2), w.gruw[2 * i + 1], a.ssum, a.g2in, a.xp2, a.out, a.hp2, a.hp2b, 0, w, s);
}

// dout: gradient w.r.t. the SRB output; dx_out receives the gradient w.r.t. its input
int tsrn_srb_backward(const Slots& sl, void* const* prm, void* const* grd, int i, const bf16* x_in, const bf16* dout,
                      bf16* dx_out, Ws& w, cudaStream_t s) {
  SrbWs& a = w.srb[i];
  SrbW& q = w.srbw[i];
  const int B = w.B;
  const long T = w.T;
  bf16 *dgin = w.g64[3], *dsum = w.g64[4];
  // gru2: input ssum = x + o1
  TRY(gru_block_backward(prm, grd, sl.srb(i, TS_G2), w.gruw[2 * i + 1], a.ssum, a.g2in, a.xp2, a.hp2, a.hp2b, dout, dgin,
                         dsum, nullptr, 0, w, s));
  // gru1: input r0, its output gradient is dsum (the o1 branch of the sum)
  bf16* dr0 = w.g128[0];  // (T,64) view of a free (T,128) buffer
  TRY(gru_block_backward(prm, grd, sl.srb(i, TS_G1), w.gruw[2 * i], a.r0, a.g1in, a.xp1, a.hp1, a.hp1b, dsum, dgin, dr0,
                         nullptr, 1, w, s));
  bf16 *dc2 = w.g128[1], *da1 = w.g128[2], *dc1 = w.g128[1];
  TRY(bn_backward(dr0, 64, a.c2, 64, a.st2, dc2, 64, T, 64, ACT_NONE, P<float>(grd, sl.srb(i, TS_BN2W)),
                  P<float>(grd, sl.srb(i, TS_BN2B)), w.partial, w.coef, s, P<float>(grd, sl.srb(i, TS_C2B))));
  if (grd[sl.srb(i, TS_C2W)]) TRY(conv3x3_wgrad(dc2, a.a1, B, 16, 64, 0, P<float>(grd, sl.srb(i, TS_C2W)), w.partial, s));
  TcGemmParams p = gp();
  p.out = da1;
  TRY(map_conv(dc2, B, 16, 64, 3, 3, q.c2d, 64, p, s));
  TRY(bn_backward(da1, 64, a.c1, 64, a.st1, dc1, 64, T, 64, ACT_MISH, P<float>(grd, sl.srb(i, TS_BN1W)),
                  P<float>(grd, sl.srb(i, TS_BN1B)), w.partial, w.coef, s, P<float>(grd, sl.srb(i, TS_C1B))));
  if (grd[sl.srb(i, TS_C1W)]) TRY(conv3x3_wgrad(dc1, x_in, B, 16, 64, 0, P<float>(grd, sl.srb(i, TS_C1W)), w.partial, s));
  p = gp();
  p.out = dx_out;
  p.residual = dsum;  // x also feeds gru2 directly through the sum
  return map_conv(dc1, B, 16, 64, 3, 3, q.c1d, 64, p, s);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
int forward(const Slots& sl, void* const* prm, const float* x_lr, float* sr_out, Ws& w, bool training, bool stn,
            float p_drop, uint32_t seed, cudaStream_t s, const uint32_t* seed_dev) {
  if (seed_dev != nullptr) seed = 0;  // the kernels XOR the device word into drop_key(0, stream)
  const int B = w.B, n = sl.srb_nums;
  const long T = w.T;
  const bool use_stn = stn && training;  // tbsrn.py:215
  const uint32_t th = training ? thresh_of(p_drop) : 0;
  const float keep_scale = 65536.f / (65536.f - (float)th);
  TRY(prep_weights(sl, prm, w, use_stn, s));
  const float* x_in = x_lr;
  if (use_stn) {
    TRY(stn_forward(sl, prm, x_lr, w, s));
    x_in = w.x_tps;
  }
  // block1: 9x9 conv 3->64 + PReLU (tbsrn.py:179-183)
  TRY(im2col_dx(x_in, w.a1x, B, 16, 64, +1, s));
  {
    TcGemmParams p = gp();
    p.bias = P<float>(prm, sl.b1_b);
    p.prelu_slope = P<float>(prm, sl.b1_a);
    p.out = w.b1;
    p.out2 = w.b1pre;
    TRY(map_conv(w.a1x, B, 16, 64, 9, 1, w.w_b1, 64, p, s));
  }
  const bf16* x = w.b1;
  for (int i = 0; i < n; ++i) {
    SrbWs& a = w.srb[i];
    SrbW& q = w.srbw[i];
    if (sl.arch == ARCH_TSRN) {
      TRY(tsrn_srb_forward(sl, prm, i, x, w, training, s));
      x = a.out;
      continue;
    }
    TcGemmParams p = gp();
    p.bias = P<float>(prm, sl.srb(i, S_C1B));
    p.out = a.c1;
    TRY(map_conv(x, B, 16, 64, 3, 3, q.c1f, 64, p, s));
    TRY(bn_fwd_stats(a.c1, 64, T, 64, prm, sl.srb(i, S_BN1W), training, w, a.st1, s));
    TRY(bn_apply(a.c1, 64, a.st1, a.a1, 64, T, 64, ACT_MISH, nullptr, 0, nullptr, s));
    p = gp();
    p.bias = P<float>(prm, sl.srb(i, S_C2B));
    p.out = a.c2;
    TRY(map_conv(a.a1, B, 16, 64, 3, 3, q.c2f, 64, p, s));
    TRY(bn_fwd_stats(a.c2, 64, T, 64, prm, sl.srb(i, S_BN2W), training, w, a.st2, s));
    // FeatureEnhancer input: [bn2(c2) | positional encoding] as (T,128) tokens (tbsrn.py:83-86)
    TRY(bn_apply(a.c2, 64, a.st2, a.f, 128, T, 64, ACT_NONE, w.pe, 1024, nullptr, s));
    p = gp();
    p.bias = q.bqkv;
    p.out = a.qkv;
    TRY(tok_gemm(a.f, 128, T, q.qkv, 384, p, s));
    TRY(attn_forward(a.qkv, a.o, a.lse, B, drop_key(seed, 2 * i), th, a.dropbits, s, seed_dev));
    p = gp();
    p.bias = P<float>(prm, sl.srb(i, S_LOB));
    p.out = a.y1pre;
    p.residual = a.f;
    TRY(tok_gemm(a.o, 128, T, q.wo, 128, p, s));
    TRY(ln_forward(a.y1pre, P<float>(prm, sl.srb(i, S_LN1A)), P<float>(prm, sl.srb(i, S_LN1B)), a.y1, T, 1e-6f, s));
    p = gp();
    p.bias = P<float>(prm, sl.srb(i, S_W1B));
    p.out = a.hd;
    p.relu = 1;
    p.drop_thresh16 = th;
    p.drop_scale = keep_scale;
    p.drop_key = drop_key(seed, 2 * i + 1);
    p.drop_seed_dev = seed_dev;
    TRY(tok_gemm(a.y1, 128, T, q.w1, 128, p, s));
    p = gp();
    p.bias = P<float>(prm, sl.srb(i, S_W2B));
    p.out = a.y2pre;
    p.residual = a.y1;
    TRY(tok_gemm(a.hd, 128, T, q.w2, 128, p, s));
    TRY(ln_forward(a.y2pre, P<float>(prm, sl.srb(i, S_LN3A)), P<float>(prm, sl.srb(i, S_LN3B)), a.y2, T, 1e-6f, s));
    p = gp();
    p.bias = P<float>(prm, sl.srb(i, S_LINB));
    p.out = a.out;
    p.residual = x;
    TRY(tok_gemm(a.y2, 128, T, q.lin, 64, p, s));
    x = a.out;
  }
  // block7: conv + BN, then the global skip block1 + block7 (tbsrn.py:188-192, 223-224)
  {
    TcGemmParams p = gp();
    p.bias = P<float>(prm, sl.b7_b);
    p.out = w.c7;
    TRY(map_conv(x, B, 16, 64, 3, 3, w.w7f, 64, p, s));
    TRY(bn_fwd_stats(w.c7, 64, T, 64, prm, sl.b7_bn, training, w, w.st7, s));
    TRY(bn_apply(w.c7, 64, w.st7, w.s7, 64, T, 64, ACT_NONE, nullptr, 0, w.b1, s));
  }
  // block8: conv 64->256, PixelShuffle(2), mish, then 9x9 conv 64->3, tanh (tbsrn.py:195-197, 225, 261-274)
  {
    TcGemmParams p = gp();
    p.bias = w.bup;
    p.out = w.upre;
    p.out2 = w.u;
    p.epi = TC_EPI_PIXSHUF;
    TRY(map_conv(w.s7, B, 16, 64, 3, 3, w.wupf, 256, p, s));
    p = gp();
    p.out = w.z;
    p.epi = TC_EPI_F32;
    TRY(map_conv(w.u, B, 32, 128, 1, 9, w.wfin, 64, p, s));
    TRY(vgather9(w.z, P<float>(prm, sl.fin_b), w.opre, B, 32, 128, +1, s));
    TRY(tanh_mse(w.opre, nullptr, w.sr, nullptr, (long)B * 3 * 4096, 0.f, nullptr, nullptr, s));
    if (sr_out != nullptr && sr_out != w.sr) TRY(d2d(sr_out, w.sr, (size_t)B * 3 * 4096 * 4, s));
  }
  return FOCR_OK;
}

// ---------------------------------------------------------------------------------------------
// backward: d_sr = dL/d(sr) fp32 (B,3,32,128); writes every parameter gradient (overwrite, not accumulate)
// ---------------------------------------------------------------------------------------------
namespace {
int lin_grads(const bf16* dy, long ld_dy, const bf16* x, long T, int N, void* const* grd, int w_slot, int b_slot, Ws& w,
              cudaStream_t s) {
  if (linear_wgrad_tc_supported(T, N, 128, ld_dy, 128) && linear_wgrad_tc_partial_bytes(N) <= w.partial_bytes &&
      (grd[w_slot] || grd[b_slot]))  // weight and bias gradient in one pass over dY and X (tcgen05, wgrad_tc.cu)
    return linear_wgrad_tc(dy, ld_dy, x, 128, T, N, P<float>(grd, w_slot), P<float>(grd, b_slot), w.partial, s);
  if (grd[w_slot]) TRY(linear_wgrad(dy, ld_dy, x, 128, T, N, 128, P<float>(grd, w_slot), 1.f, w.partial, s));
  if (grd[b_slot]) TRY(colsum(dy, ld_dy, T, N, P<float>(grd, b_slot), w.partial, s));
  return FOCR_OK;
}
}  // namespace

int backward(const Slots& sl, void* const* prm, void* const* grd, const float* x_lr, const float* d_sr, Ws& w,
             bool stn, float p_drop, uint32_t seed, cudaStream_t s) {
  const int B = w.B, n = sl.srb_nums;
  const long T = w.T, Thr = w.Thr;
  const uint32_t th = thresh_of(p_drop);
  const float keep_scale = 65536.f / (65536.f - (float)th);
  // tanh, final 9x9 conv
  TRY(tanh_backward(w.sr, d_sr, w.d_o, (long)B * 3 * 4096, s));
  if (grd[sl.fin_b]) TRY(nchw3_sum(w.d_o, B, 4096, P<float>(grd, sl.fin_b), w.partial, s));
  TRY(im2col_dx(w.d_o, w.a1d, B, 32, 128, -1, s));
  if (grd[sl.fin_w]) {
    TRY(conv9x1_wgrad(w.a1d, w.u, B, 32, 128, w.tmpw, w.partial, s));
    TRY(repack_w9(w.tmpw, P<float>(grd, sl.fin_w), 1, s));
  }
  {
    TcGemmParams p = gp();
    p.out = w.du;
    TRY(map_conv(w.a1d, B, 32, 128, 9, 1, w.wfind, 64, p, s));
  }
  TRY(mish_backward(w.du, w.upre, w.dupre, Thr * 64, s));
  // up-conv (PixelShuffle layout)
  if (grd[sl.up_w]) TRY(conv3x3_wgrad(w.dupre, w.s7, B, 16, 256, 1, P<float>(grd, sl.up_w), w.partial, s));
  if (grd[sl.up_b]) {
    for (int i = 0; i < 2; ++i)
      TRY(colsum2(w.dupre + (long)i * 128 * 64, (long)B * 16, 64, 2L * 128 * 64, 128, 128, w.tmpb + i * 128, w.partial,
                  s));
    TRY(bias_unshuf(w.tmpb, P<float>(grd, sl.up_b), 256, s));
  }
  bf16* dS = w.g64[0];
  {
    TcGemmParams p = gp();
    p.n_total = 64;
    p.kh = p.kw = 3;
    p.W = 64;
    p.H = 16;
    p.ldc = 64;
    p.out = dS;
    const bf16* ap[4];
    for (int sub = 0; sub < 4; ++sub) ap[sub] = w.dupre + ((long)(sub >> 1) * 128 + (sub & 1)) * 64;
    TRY(tc_gemm_launch(ap, 4, 128, 2L * 128 * 64, 32L * 128 * 64, 64, B, w.wupd, 256, p, s));
  }
  // block7
  bf16* dc7 = w.g64[1];
  TRY(bn_backward(dS, 64, w.c7, 64, w.st7, dc7, 64, T, 64, ACT_NONE, P<float>(grd, sl.b7_bn + BN_W),
                  P<float>(grd, sl.b7_bn + BN_B), w.partial, w.coef, s, P<float>(grd, sl.b7_b)));
  const bf16* x_last = n > 0 ? w.srb[n - 1].out : w.b1;
  if (grd[sl.b7_w]) TRY(conv3x3_wgrad(dc7, x_last, B, 16, 64, 0, P<float>(grd, sl.b7_w), w.partial, s));
  bf16* dcur = w.g64[2];
  {
    TcGemmParams p = gp();
    p.out = dcur;
    TRY(map_conv(dc7, B, 16, 64, 3, 3, w.w7d, 64, p, s));
  }
  bf16* dfree = w.g64[1];  // free once dcur has been produced
  // SRBs in reverse
  for (int i = n - 1; i >= 0; --i) {
    SrbWs& a = w.srb[i];
    SrbW& q = w.srbw[i];
    const bf16* x_in = i > 0 ? w.srb[i - 1].out : w.b1;
    if (sl.arch == ARCH_TSRN) {
      TRY(tsrn_srb_backward(sl, prm, grd, i, x_in, dcur, dfree, w, s));
      bf16* t = dcur;
      dcur = dfree;
      dfree = t;
      continue;
    }
    bf16 *gA = w.g128[0], *gB = w.g128[1], *gC = w.g128[2];
    TcGemmParams p = gp();
    // out = x + linear(y2)
    p.out = gA;
    TRY(tok_gemm(dcur, 64, T, q.linT, 128, p, s));
    TRY(lin_grads(dcur, 64, a.y2, T, 64, grd, sl.srb(i, S_LINW), sl.srb(i, S_LINB), w, s));
    TRY(ln_backward(gA, a.y2pre, P<float>(prm, sl.srb(i, S_LN3A)), gB, P<float>(grd, sl.srb(i, S_LN3A)),
                    P<float>(grd, sl.srb(i, S_LN3B)), w.partial, T, 1e-6f, s));
    // y2pre = y1 + w2(hd);  hd = dropout(relu(w1 y1))
    p = gp();
    p.out = gC;
    p.gate = a.hd;
    p.gate_scale = keep_scale;
    TRY(tok_gemm(gB, 128, T, q.w2T, 128, p, s));
    TRY(lin_grads(gB, 128, a.hd, T, 128, grd, sl.srb(i, S_W2W), sl.srb(i, S_W2B), w, s));
    p = gp();
    p.out = gA;
    p.residual = gB;
    TRY(tok_gemm(gC, 128, T, q.w1T, 128, p, s));
    TRY(lin_grads(gC, 128, a.y1, T, 128, grd, sl.srb(i, S_W1W), sl.srb(i, S_W1B), w, s));
    TRY(ln_backward(gA, a.y1pre, P<float>(prm, sl.srb(i, S_LN1A)), gB, P<float>(grd, sl.srb(i, S_LN1A)),
                    P<float>(grd, sl.srb(i, S_LN1B)), w.partial, T, 1e-6f, s));
    // y1pre = f + wo(attn(qkv(f)))
    p = gp();
    p.out = gC;
    TRY(tok_gemm(gB, 128, T, q.woT, 128, p, s));
    TRY(lin_grads(gB, 128, a.o, T, 128, grd, sl.srb(i, S_LOW), sl.srb(i, S_LOB), w, s));
    TRY(attn_backward(a.qkv, a.o, gC, a.lse, w.dsum, w.g384, B, drop_key(seed, 2 * i), th, a.dropbits, s));
    p = gp();
    p.out = gA;
    p.residual = gB;
    TRY(tok_gemm(w.g384, 384, T, q.qkvT, 128, p, s));
    if (linear_wgrad_tc_supported(T, 384, 128, 384, 128) && linear_wgrad_tc_partial_bytes(384) <= w.partial_bytes) {
      TRY(linear_wgrad_tc(w.g384, 384, a.f, 128, T, 384, w.tmpw, w.tmpb, w.partial, s));
    } else {
      TRY(linear_wgrad(w.g384, 384, a.f, 128, T, 384, 128, w.tmpw, 1.f, w.partial, s));
      TRY(colsum(w.g384, 384, T, 384, w.tmpb, w.partial, s));
    }
    for (int j = 0; j < 3; ++j) {
      if (grd[sl.srb(i, S_LQW + 2 * j)]) TRY(d2d(grd[sl.srb(i, S_LQW + 2 * j)], w.tmpw + j * 128 * 128, 128 * 128 * 4, s));
      if (grd[sl.srb(i, S_LQB + 2 * j)]) TRY(d2d(grd[sl.srb(i, S_LQB + 2 * j)], w.tmpb + j * 128, 128 * 4, s));
    }
    // f = [bn2(c2) | pe]: only the left 64 columns carry gradient
    bf16 *dc2 = w.g64[3], *da1 = w.g64[4];
    TRY(bn_backward(gA, 128, a.c2, 64, a.st2, dc2, 64, T, 64, ACT_NONE, P<float>(grd, sl.srb(i, S_BN2W)),
                    P<float>(grd, sl.srb(i, S_BN2B)), w.partial, w.coef, s, P<float>(grd, sl.srb(i, S_C2B))));
    if (grd[sl.srb(i, S_C2W)]) TRY(conv3x3_wgrad(dc2, a.a1, B, 16, 64, 0, P<float>(grd, sl.srb(i, S_C2W)), w.partial, s));
    p = gp();
    p.out = da1;
    TRY(map_conv(dc2, B, 16, 64, 3, 3, q.c2d, 64, p, s));
    bf16* dc1 = w.g64[3];
    TRY(bn_backward(da1, 64, a.c1, 64, a.st1, dc1, 64, T, 64, ACT_MISH, P<float>(grd, sl.srb(i, S_BN1W)),
                    P<float>(grd, sl.srb(i, S_BN1B)), w.partial, w.coef, s, P<float>(grd, sl.srb(i, S_C1B))));
    if (grd[sl.srb(i, S_C1W)]) TRY(conv3x3_wgrad(dc1, x_in, B, 16, 64, 0, P<float>(grd, sl.srb(i, S_C1W)), w.partial, s));
    p = gp();
    p.out = dfree;
    p.residual = dcur;  // the SRB's identity branch
    TRY(map_conv(dc1, B, 16, 64, 3, 3, q.c1d, 64, p, s));
    bf16* t = dcur;
    dcur = dfree;
    dfree = t;
  }
  // block1 output feeds SRB 1 and the global skip
  bf16 *db1 = w.g64[3], *db1pre = w.g64[4];
  TRY(add_bf16(dcur, dS, db1, T * 64, s));
  TRY(prelu_backward(db1, w.b1pre, P<float>(prm, sl.b1_a), db1pre, T * 64, P<float>(grd, sl.b1_a), w.partial, s));
  if (grd[sl.b1_w]) {
    TRY(conv9x1_wgrad(db1pre, w.a1x, B, 16, 64, w.tmpw, w.partial, s));
    TRY(repack_w9(w.tmpw, P<float>(grd, sl.b1_w), 0, s));
  }
  if (grd[sl.b1_b]) TRY(colsum(db1pre, 64, T, 64, P<float>(grd, sl.b1_b), w.partial, s));
  if (stn) {
    TcGemmParams p = gp();
    p.out = w.z;
    p.epi = TC_EPI_F32;
    TRY(map_conv(db1pre, B, 16, 64, 1, 9, w.w_b1d, 64, p, s));
    TRY(vgather9(w.z, nullptr, w.dx_tps, B, 16, 64, -1, s));
    TRY(stn_backward(sl, prm, grd, x_lr, w, s));
  }
  return FOCR_OK;
}

}  // namespace tbsrn
