// TBSRN engine: parameter-slot table and workspace layout shared by forward and backward.
// The network is scene-text-telescope/model/tbsrn.py:166-226 (TBSRN), :229-257 (SRB), :63-92
// (FeatureEnhancer), model/stn_head.py, model/tps_spatial_transformer.py.
#pragma once
#include <string>
#include <vector>

#include "kernels.cuh"

namespace tbsrn {

// ---- parameter / buffer slots (the caller passes one device pointer per slot) ------------------
enum Srb : int {
  S_C1W, S_C1B, S_BN1W, S_BN1B, S_BN1RM, S_BN1RV, S_BN1NBT,
  S_C2W, S_C2B, S_BN2W, S_BN2B, S_BN2RM, S_BN2RV, S_BN2NBT,
  S_LQW, S_LQB, S_LKW, S_LKB, S_LVW, S_LVB, S_LOW, S_LOB,
  S_LN1A, S_LN1B, S_W1W, S_W1B, S_W2W, S_W2B, S_LN3A, S_LN3B, S_LINW, S_LINB, S_COUNT
};
enum Bn : int { BN_W, BN_B, BN_RM, BN_RV, BN_NBT };
// TSRN's SRB (scene-text-telescope/model/tsrn.py:77-98): conv/BN/mish/conv/BN + vertical BiGRU + horizontal BiGRU
enum Gru : int { G_CW, G_CB, G_WIH, G_WHH, G_BIH, G_BHH, G_WIH_R, G_WHH_R, G_BIH_R, G_BHH_R, G_COUNT };
enum TsrnSrb : int {
  TS_C1W, TS_C1B, TS_BN1W, TS_BN1B, TS_BN1RM, TS_BN1RV, TS_BN1NBT, TS_G1,
  TS_C2W = TS_G1 + G_COUNT, TS_C2B, TS_BN2W, TS_BN2B, TS_BN2RM, TS_BN2RV, TS_BN2NBT, TS_G2, TS_COUNT = TS_G2 + G_COUNT
};
enum Arch : int { ARCH_TBSRN = 0, ARCH_TSRN = 1 };

struct Slots {
  int srb_nums;
  int arch = ARCH_TBSRN;
  int srb_stride = S_COUNT;
  int b1_w = 0, b1_b = 1, b1_a = 2;
  int srb0 = 3;
  int b7_w, b7_b, b7_bn;          // bn: +0 w, +1 b, +2 rm, +3 rv, +4 nbt
  int up_w, up_b, fin_w, fin_b;
  int stn_conv0;                  // 6 x {w, b, bn w, bn b, rm, rv, nbt}
  int fc1_w, fc1_b, bn1d, fc2_w, fc2_b;
  int tps_inv, tps_repr;
  int count;
  explicit Slots(int n, int arch_ = ARCH_TBSRN) : srb_nums(n), arch(arch_) {
    srb_stride = arch == ARCH_TSRN ? (int)TS_COUNT : (int)S_COUNT;
    int i = srb0 + n * srb_stride;
    b7_w = i++; b7_b = i++; b7_bn = i; i += 5;
    up_w = i++; up_b = i++; fin_w = i++; fin_b = i++;
    stn_conv0 = i; i += 6 * 7;
    fc1_w = i++; fc1_b = i++; bn1d = i; i += 5; fc2_w = i++; fc2_b = i++;
    tps_inv = i++; tps_repr = i++;
    count = i;
  }
  int srb(int blk, int s) const { return srb0 + blk * srb_stride + s; }
  int stn(int conv, int j) const { return stn_conv0 + conv * 7 + j; }
};

std::vector<std::string> slot_names(int srb_nums, int arch = ARCH_TBSRN);

// ---- STN conv geometry ---------------------------------------------------------------------------
struct StnConv {
  int cin, cout, h, w, kpad, npad, pool_h;  // pool_h: 0 none, 2 = (2,2), 1 = (1,2)
};
static const StnConv kStn[6] = {
    {3, 32, 16, 64, 128, 64, 2},   {32, 64, 8, 32, 384, 64, 2},    {64, 128, 4, 16, 640, 128, 2},
    {128, 256, 2, 8, 1152, 256, 2}, {256, 256, 1, 4, 2304, 256, 1}, {256, 256, 1, 2, 2304, 256, 0},
};

// ---- workspace -------------------------------------------------------------------------------------
struct SrbWs {
  bf16 *c1, *a1, *c2, *f, *qkv, *o, *y1pre, *y1, *hd, *y2pre, *y2, *out;
  float *lse, *st1, *st2;  // stats [4][64]
  uint32_t* dropbits;      // attention keep bits, 1 per (b,h,q,k)
  // TSRN: r0 = bn2(c2); per GRU block: conv1x1 output gin, projected input xp (T,192), output, saved h_{t-1}
  bf16 *r0, *g1in, *xp1, *o1, *hp1b, *ssum, *g2in, *xp2, *hp2b;
  float *hp1, *hp2;
};
struct GruW {  // prepared GRU-block weights
  bf16 *cw, *cwT;      // conv1x1 [64][64] and transpose
  bf16 *wih, *wihT;    // [192][64], [64][192]
  float* bih;          // [192]
};
struct SrbW {  // prepared bf16 weights
  bf16 *c1f, *c1d, *c2f, *c2d;             // conv fwd / dgrad layouts [9][64][64]
  bf16 *qkv, *qkvT, *wo, *woT, *w1, *w1T, *w2, *w2T, *lin, *linT;
  float* bqkv;                              // [384]
};
struct Ws {
  int B, srb_nums;
  long T, Thr;
  // constants
  bf16* pe;            // [1024][64]
  // stn
  bf16 *stn_col[6], *stn_ypre[6], *stn_yact[6], *stn_pool[6];
  float* stn_stats[6];
  float* stn_bias[6];  // conv bias padded to npad
  bf16 *stn_wf[6], *stn_wt[6];
  bf16 *fc1_wf, *fc1_wt, *fc2_wf, *fc2_wt;
  float* fc2_bias;     // [64]
  bf16 *f1pre, *f1;    // (Bpad,512)
  float* bn1d_stats;   // [4][512]
  float* ctrl;         // (Bpad,64) fp32
  float* x_tps;        // (B,3,16,64)
  // trunk
  bf16 *a1x, *b1pre, *b1;
  bf16 *w_b1, *w_b1d;  // [9][64][64]
  std::vector<SrbWs> srb;
  std::vector<SrbW> srbw;
  std::vector<GruW> gruw;  // 2 per SRB (TSRN)
  bf16 *gdxp, *gdhid;      // (T,192) backward temporaries (TSRN)
  int arch;
  bf16 *c7, *s7;       // conv7 output, b1 + bn7(c7)
  float* st7;
  bf16 *w7f, *w7d, *wupf, *wupd, *wfin, *wfind;
  float* bup;          // shuffled bias [256]
  bf16 *upre, *u;      // (Thr,64)
  float* z;            // (Thr,64) fp32
  float* opre;         // (B,3,32,128)
  float* sr;           // (B,3,32,128) tanh output (saved for backward)
  // backward temporaries
  float* d_o;          // (B,3,32,128)
  bf16* a1d;           // (Thr,64)
  bf16 *du, *dupre;    // (Thr,64)
  bf16* g64[5];        // (T,64)
  bf16* g128[4];       // (T,128)
  bf16* g384;          // (T,384)
  float* dsum;         // (B*4*1024)
  float* dx_tps;       // (B,3,16,64)
  float* dctrl;        // (B,40)
  bf16* dctrl_b;       // (Bpad,64)
  bf16 *df1, *df1pre, *dfeat;  // (Bpad,512)
  bf16 *stn_dypre[6], *stn_dyact[6], *stn_dcol[6], *stn_dpool[6];
  // scratch
  float* partial;      // reduction partials
  size_t partial_bytes;
  float* coef;         // [2][2048]
  float* tmpw;         // fp32 scratch for weight grads (>= 256*2304)
  float* tmpb;         // fp32 [512]
  size_t total_bytes;
};

// Carves the layout out of `base` (may be nullptr to only measure).  Deterministic in (B, srb_nums).
void layout(Ws& w, int B, int srb_nums, void* base, int arch = ARCH_TBSRN);

int forward(const Slots& sl, void* const* prm, const float* x_lr, float* sr_out, Ws& w, bool training, bool stn,
            float p_drop, uint32_t seed, cudaStream_t s, const uint32_t* seed_dev = nullptr);
int backward(const Slots& sl, void* const* prm, void* const* grd, const float* x_lr, const float* d_sr, Ws& w,
             bool stn, float p_drop, uint32_t seed, cudaStream_t s);

}  // namespace tbsrn
