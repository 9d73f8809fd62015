// Internal host-side launchers shared by the C-ABI layer (focr_api.cu) and the TBSRN engine.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

// prep.cu
int prep_conv_w_fwd(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s);
int prep_conv_w_dgrad(const float* w, bf16* o, int Co, int Ci, int ks, int shuf, cudaStream_t s);
int prep_bias_shuf(const float* b, float* o, int Co, cudaStream_t s);
int prep_linear_w(const float* w, bf16* o, bf16* ot, int R, int C, cudaStream_t s);
