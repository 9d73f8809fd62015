// tcgen05 + TMA implicit-GEMM engine for the NHWC bf16 feature maps of TBSRN/TSRN.
//
// One kernel serves
//   * nn.Linear forward / input-gradient      (a 1x1 "conv" over the (B,16,64,C) token map)
//   * nn.Conv2d 3x3 pad 1 forward / dgrad     (9 taps, TMA zero-fill supplies the padding)
// as  D[M=pixels, N=Cout] = sum_{tap,chunk} A_tap[M, 64] * W_tap[N, 64]^T
// with the fp32 accumulator in TMEM.  Reference call sites being replaced:
//   scene-text-telescope/model/tbsrn.py:232,237 (SRB convs), :190 (block7), :264 (up-conv),
//   :103,119-129 (attention linears), :158-163 (FFN), :74 (Linear 128->64).
#pragma once
#include "common.cuh"

enum TcEpilogue : int {
  TC_EPI_BF16 = 0,     // out bf16 [M, ldc]: acc + bias (+relu) (+residual)
  TC_EPI_F32 = 1,      // out fp32 [M, ldc]: acc + bias
  TC_EPI_PIXSHUF = 2,  // N = 4*64 ordered (sub, c): PixelShuffle(2) scatter to (B,2H,2W,64); out = pre-act,
                       // out2 = mish(pre-act)
};

struct TcGemmParams {
  int m_tiles;         // M / 128
  int n_blocks;        // N_total / BLOCK_N
  int n_total;         // N_total
  int kh, kw;          // filter taps (1x1, 3x3, 1x9, 9x1); zero padding kh/2, kw/2 via TMA OOB fill
  int chunks;          // 64-wide K chunks per tap (Cin / 64)
  int chunks_per_map;  // chunks served by one A tensor map
  int W, H;            // spatial size of the A feature map (W in {64,128}); tile = 128 consecutive pixels
  int epi;
  int relu;
  int ldc;             // leading dimension (elements) of out / residual
  const float* bias;   // [N_total] or nullptr
  void* out;
  void* out2;
  const bf16* residual;  // nullptr or [M, ldc]
  // forward dropout applied after relu (PositionwiseFeedForward, tbsrn.py:162-163); 0 = off
  uint32_t drop_thresh16;
  float drop_scale;      // 1/(1-p)
  uint32_t drop_key;     // drop_key(seed, stream)
  const uint32_t* drop_seed_dev;  // optional device word XOR-ed into drop_key in the kernel (CUDA-graph replay)
  // backward gate: v = gate[m][n] > 0 ? v * gate_scale : 0  (relu'(.) * dropout mask, read from the
  // stored post-dropout activation).  With a residual as well the order is (acc + residual) then gate.
  const bf16* gate;
  float gate_scale;
  // PReLU epilogue (block1, tbsrn.py:180-182): out = x > 0 ? x : slope[0]*x ; out2 (optional) = x
  const float* prelu_slope;
  int relu_post;  // ReLU applied after the residual add (BasicBlock: relu(bn(conv) + skip))
  int col_mode;    // set by tc_gemm_launch: 3x3 taps taken column-wise from (rows + 2)-row tiles (see tc_gemm.cu)
  int b_resident;  // set by tc_gemm_launch: whole weight operand resident in shared memory (see tc_gemm.cu)
  int tma_out;  // set by tc_gemm_launch: epilogue goes TMEM -> smem -> TMA store, aux tile prefetched by TMA
  // set by tc_gemm_launch: maps whose width is not a power of two up to 128 (the 16 x 160 maps of 32 x 320 crops) are cut into
  // boxes of bw x (128 / bw) pixels, wblocks = W / bw of them side by side; wblocks == 1 is the row-contiguous tiling
  int bw, wblocks, tiles_per_img;
};

int tc_gemm_launch(const bf16* const* a_ptrs, int n_amaps, long a_pix_stride /*elements between pixels*/,
                   long a_row_stride /*elements between image rows*/, long a_img_stride, int a_channels,
                   int B, const bf16* w /*[taps][N_total][Cin]*/, int cin, TcGemmParams p,
                   cudaStream_t stream);

// 2-D bf16 tensor map {inner, outer} with row stride `row_bytes`, box {box_inner, box_outer} and a 32/64/128-byte
// swizzle (shared with the attention kernels).
int focr_make_tmap_4d(CUtensorMap* out, const void* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                      const unsigned box[4]);
int focr_make_tmap_2d(CUtensorMap* out, const void* base, unsigned long long inner, unsigned long long outer,
                      unsigned long long row_bytes, unsigned box_inner, unsigned box_outer, int swizzle_bytes);
