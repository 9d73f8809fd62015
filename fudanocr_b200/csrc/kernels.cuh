// Internal host-side launchers shared by the C-ABI layer (focr_api.cu) and the TBSRN engine.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

enum ActKind : int { ACT_NONE = 0, ACT_MISH = 1, ACT_RELU = 2 };

// prep.cu
int prep_conv_w_fwd(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s);
int prep_conv_w_dgrad(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s);
int prep_bias_shuf(const float* b, float* o, int Co, cudaStream_t s);
int bias_unshuf(const float* b, float* o, int Co, cudaStream_t s);
int prep_linear_w(const float* w, bf16* o, bf16* ot, int R, int C, int ot_ld, int ot_off, cudaStream_t s);

// elementwise.cu

int bn_partial_blocks(long T, int C);
int bn_train_stats(const bf16* x, long ld, long T, int C, const float* gamma, const float* beta, float* rm, float* rv,
                   long long* nbt, float eps, float momentum, float* partial, float* stats, cudaStream_t s);
int bn_eval_stats(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C,
                  float* stats, cudaStream_t s);
int bn_apply(const bf16* x, long ld_x, const float* stats, bf16* out, long ld_out, long T, int C, int act,
             const bf16* pe, int pe_rows, const bf16* res, cudaStream_t s);
// dx_colsum (optional, fp32 [C]): column sums of dx as stored = bias gradient of the conv feeding this BatchNorm
int bn_backward(const bf16* dy, long ld_dy, const bf16* x, long ld_x, const float* stats, bf16* dx, long ld_dx, long T,
                int C, int act, float* dgamma, float* dbeta, float* partial, float* coef, cudaStream_t s,
                float* dx_colsum = nullptr);
int ln_partial_blocks(long T);
int ln_forward(const bf16* x, const float* a, const float* b, bf16* y, long T, float eps, cudaStream_t s);
int ln_backward(const bf16* dy, const bf16* x, const float* a, bf16* dx, float* da, float* db, float* partial, long T,
                float eps, cudaStream_t s);
int colsum_partial_blocks(long T, int C);
int colsum(const bf16* x, long ld, long T, int C, float* out, float* partial, cudaStream_t s);
int colsum2(const bf16* x, long t_outer, long t_inner, long stride_outer, long stride_inner, int C, float* out,
            float* partial, cudaStream_t s);
int reduce_partials(const float* partial, int P, long stride, int n, float* out, float scale, cudaStream_t s);
int prelu_backward(const bf16* dy, const bf16* x, const float* a, bf16* dx, long n, float* da, float* partial,
                   cudaStream_t s);
int tanh_mse(const float* o, const float* hr, float* sr, float* dout, long n, float gscale, float* loss,
             float* partial, cudaStream_t s);
int tanh_backward(const float* sr, const float* dsr, float* dout, long n, cudaStream_t s);
int add_bf16(const bf16* a, const bf16* b, bf16* o, long n, cudaStream_t s);
int mish_backward(const bf16* dy, const bf16* x, bf16* dx, long n, cudaStream_t s);
int pe_table(bf16* pe, cudaStream_t s);

// attention.cu
size_t attn_drop_bits_bytes(int B);
// seed_dev (optional): device word XOR-ed into `key` inside the kernel, so a captured CUDA graph can be replayed with a
// fresh dropout seed per step (key = drop_key(0, stream) in that case)
int attn_forward(const bf16* qkv, bf16* out, float* lse2, int B, uint32_t key, uint32_t thresh16, uint32_t* drop_bits,
                 cudaStream_t s, const uint32_t* seed_dev = nullptr);
int attn_backward(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse2, float* dsum, bf16* dqkv, int B,
                  uint32_t key, uint32_t thresh16, const uint32_t* drop_bits, cudaStream_t s);

// wgrad.cu
int linear_wgrad_splits(long T, int N, int K = 128);
size_t linear_wgrad_partial_bytes(long T, int N, int K = 128);
int linear_wgrad(const bf16* dy, long ld_dy, const bf16* x, long ld_x, long T, int N, int K, float* dw, float scale,
                 float* partial, cudaStream_t s);
size_t conv9x1_wgrad_partial_bytes(int B, int H, int W);
int conv9x1_wgrad(const bf16* dy, const bf16* x, int B, int H, int W, float* out, float* partial, cudaStream_t s);
size_t conv3x3_wgrad_partial_bytes(int B, int H, int groups);
int conv3x3_wgrad(const bf16* dy, const bf16* x, int B, int H, int Co, int shuf, float* dw, float* partial,
                  cudaStream_t s);

// wgrad_tc.cu: tcgen05 weight + bias gradient of a linear layer with K = 128 inputs
bool linear_wgrad_tc_supported(long T, int N, int K, long ld_dy, long ld_x);
size_t linear_wgrad_tc_partial_bytes(int N);
// wgrad_tc.cu: tcgen05 weight gradient of a 3x3 conv between 64-channel maps of width 64 (two taps stacked along M)
bool conv3x3_wgrad_tc_supported(int H, int W);
size_t conv3x3_wgrad_tc_partial_bytes();
int conv3x3_wgrad_tc(const bf16* dy, long dy_pix, long dy_row, long dy_img, const bf16* x, int B, int H, int co_mul, int co_add,
                     float* dw, float* partial, cudaStream_t s);
// ... and for any channel counts / map widths (the recognisers' encoders): implicit shifted operand instead of a materialised im2col
bool conv3x3_wgrad_tc_general_supported(int B, int H, int W, int Ci, int Co);
size_t conv3x3_wgrad_tc_general_partial_bytes(int B, int H, int W, int Ci, int Co);
int conv3x3_wgrad_tc_general(const bf16* dy, const bf16* x, int B, int H, int W, int Ci, int Co, float* dw, float* partial,
                             cudaStream_t s);
int linear_wgrad_tc(const bf16* dy, long ld_dy, const bf16* x, long ld_x, long T, int N, float* dw, float* db, float* partial,
                    cudaStream_t s);

// conv9x9.cu
int im2col_dx(const float* in, bf16* out, int B, int H, int W, int sgn, cudaStream_t s);
int vgather9(const float* z, const float* bias, float* out, int B, int H, int W, int sgn, cudaStream_t s);
int prep_w9(const float* w, bf16* o, int mode, cudaStream_t s);
int repack_w9(const float* t, float* dw, int mode, cudaStream_t s);
int nchw3_sum_blocks(int B, long hw);
int nchw3_sum(const float* x, int B, long hw, float* out3, float* partial, cudaStream_t s);

// stn.cu
int im2col3x3(const bf16* x, const float* x_nchw, bf16* col, int B, int H, int W, int C, int Kpad, cudaStream_t s);
int col2im3x3(const bf16* dcol, bf16* dx, int B, int H, int W, int C, int Kpad, cudaStream_t s);
int maxpool_fwd(const bf16* x, bf16* y, int B, int H, int W, int C, int ph, cudaStream_t s);
int maxpool_bwd(const bf16* x, const bf16* y, const bf16* dy, bf16* dx, int B, int H, int W, int C, int ph,
                cudaStream_t s);
int tps_forward(const float* img, const float* ctrl, long ld_ctrl, const float* inv, const float* repr, float* out,
                int B, cudaStream_t s);
int tps_backward(const float* img, const float* ctrl, long ld_ctrl, const float* inv, const float* repr,
                 const float* dout, float* dctrl, long ld_dctrl, int B, cudaStream_t s);
int bf16_to_f32(const bf16* x, long ld, float* y, long rows, int n, cudaStream_t s);
int f32_to_bf16_pad(const float* x, int n, long rows, bf16* y, long ld, long rows_pad, cudaStream_t s);
int prep_fc1(const float* w1, bf16* o, bf16* ot, cudaStream_t s);
int unperm_fc1_grad(const float* g, float* dw1, cudaStream_t s);
int prep_fc2(const float* w2, bf16* o, bf16* ot, cudaStream_t s);
int prep_stn_conv_w(const float* w, bf16* o, bf16* ot, int Co, int C, int Npad, int Kpad, cudaStream_t s);
int unpack_stn_conv_grad(const float* g, float* dw, int Co, int C, int Kpad, cudaStream_t s);

// optim.cu
int adam_clip_step(const void* chunks, int n_chunks, float gscale, float max_norm, float lr, float b1, float b2,
                   float eps, long long* step, float* state, float* partial, cudaStream_t s);

// gru.cu
int gru_prep(const float* wih_f, const float* wih_r, const float* bih_f, const float* bih_r, bf16* w, bf16* wt, float* b,
             cudaStream_t s);
int gru_forward(const bf16* xp, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r, bf16* out,
                float* hprev32, bf16* hprev16, int B, int vertical, cudaStream_t s);
int gru_backward(const bf16* xp, const float* whh_f, const float* whh_r, const float* bhh_f, const float* bhh_r,
                 const float* hprev32, const bf16* dout, bf16* dxp, bf16* dhid, int B, int vertical, cudaStream_t s);
int gru_unpack_whh(const float* g, float* df, float* dr, cudaStream_t s);
int linear_wgrad_k64(const bf16* dy, const bf16* x, long T, int N, float* dw, float* tmp, float* partial, cudaStream_t s);
