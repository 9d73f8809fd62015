// Internal host-side launchers shared by the C-ABI layer (focr_api.cu) and the TBSRN engine.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

enum ActKind : int { ACT_NONE = 0, ACT_MISH = 1, ACT_RELU = 2 };

// prep.cu
int prep_conv_w_fwd(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s);
int prep_conv_w_dgrad(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s);
int prep_bias_shuf(const float* b, float* o, int Co, cudaStream_t s);
int prep_linear_w(const float* w, bf16* o, bf16* ot, int R, int C, cudaStream_t s);

// elementwise.cu

int bn_partial_blocks(long T, int C);
int bn_train_stats(const bf16* x, long ld, long T, int C, const float* gamma, const float* beta, float* rm, float* rv,
                   long long* nbt, float eps, float momentum, float* partial, float* stats, cudaStream_t s);
int bn_eval_stats(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C,
                  float* stats, cudaStream_t s);
int bn_apply(const bf16* x, long ld_x, const float* stats, bf16* out, long ld_out, long T, int C, int act,
             const bf16* pe, int pe_rows, cudaStream_t s);
int bn_backward(const bf16* dy, long ld_dy, const bf16* x, long ld_x, const float* stats, bf16* dx, long ld_dx, long T,
                int C, int act, float* dgamma, float* dbeta, float* partial, float* coef, cudaStream_t s);
int ln_partial_blocks(long T);
int ln_forward(const bf16* x, const float* a, const float* b, bf16* y, long T, float eps, cudaStream_t s);
int ln_backward(const bf16* dy, const bf16* x, const float* a, bf16* dx, float* da, float* db, float* partial, long T,
                float eps, cudaStream_t s);
int colsum_partial_blocks(long T, int C);
int colsum(const bf16* x, long ld, long T, int C, float* out, float* partial, cudaStream_t s);
int reduce_partials(const float* partial, int P, long stride, int n, float* out, float scale, cudaStream_t s);
int prelu_backward(const bf16* dy, const bf16* x, const float* a, bf16* dx, long n, float* da, float* partial,
                   cudaStream_t s);
int tanh_mse(const float* o, const float* hr, float* sr, float* dout, long n, float gscale, float* loss,
             float* partial, cudaStream_t s);
int tanh_backward(const float* sr, const float* dsr, float* dout, long n, cudaStream_t s);
int add_bf16(const bf16* a, const bf16* b, bf16* o, long n, cudaStream_t s);
int pe_table(bf16* pe, cudaStream_t s);

// attention.cu
int attn_forward(const bf16* qkv, bf16* out, float* lse2, int B, uint32_t key, uint32_t thresh16, cudaStream_t s);
int attn_backward(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse2, float* dsum, bf16* dqkv, int B,
                  uint32_t key, uint32_t thresh16, cudaStream_t s);
