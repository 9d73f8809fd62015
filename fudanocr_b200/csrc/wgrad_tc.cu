// Weight + bias gradient of nn.Linear on the 5th-generation tensor cores:
//     dW[n][k] = sum_t dY[t][n] X[t][k]   (K = 128),      db[n] = sum_t dY[t][n]
// (autograd of the FeatureEnhancer linears, scene-text-telescope/model/tbsrn.py:74,103,158-159).
//
// The reduction runs over the token axis (T = B*1024), the outputs are tiny, so this is a streaming kernel that
// must read dY and X exactly once at HBM speed.  Both operands are consumed AS STORED: a [64 tokens][64 columns]
// TMA box with the 128-byte swizzle is, read MN-major, a 64-wide slice of dY^T (A operand, M = output features) or
// of X^T (B operand, N = input features) with the tokens as the MMA K dimension - no transposes, no register
// staging.  M = 128 / N = 128 operands span two such atoms (LBO = atom stride); the descriptor semantics are pinned
// by tests/test_gpu_umma_layouts.py.  The bias gradient rides along as one more MMA per step against a constant
// all-ones B operand (N = 16), so dY is not read a second time.
//
// One persistent CTA per SM owns a contiguous range of 64-token tiles:
//   warp 0: TMA producer (ring of stages, each = N/64 dY atoms + 2 X atoms)      warp 1: MMA issuer
//   warp 2: TMEM allocator (512 columns: up to 3 x [128 x 128] fp32 + 3 x [128 x 16])
//   warps 4-7: epilogue - after the last tile, TMEM -> one fp32 partial per CTA.  Second stage in the same launch: grid
//              barrier, then each CTA sums the <= 148 partials of its slice of dW / db in a fixed order (no atomics on
//              data: run-to-run reproducible)
#include "kernels.cuh"

#define TRY_RC(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

namespace {

constexpr int kTokTile = 64;
constexpr int kAtomBytes = kTokTile * 128;  // [64 tokens][64 bf16], 128-byte swizzled
constexpr int kSmemRing = 192 * 1024;
constexpr int kOnesBytes = 8192;
constexpr int kThreads = 256;
constexpr int kBiasCol0 = 384;  // TMEM column of the first bias accumulator
constexpr int kMaxStages = 8;

__global__ void __launch_bounds__(kThreads, 1)
linear_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mY, const __grid_constant__ CUtensorMap mX, int n_atoms,
                       int tiles_total, int stages, float* __restrict__ partial_w, float* __restrict__ partial_b,
                       float* __restrict__ dw, float* __restrict__ db, unsigned int* __restrict__ sync_ctr, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ones = smem + kSmemRing;
  uint64_t* full = reinterpret_cast<uint64_t*>(ones + kOnesBytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* done = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = n_atoms * 64;
  const int m_blocks = n_atoms >= 2 ? n_atoms / 2 : 1;
  const int M = n_atoms >= 2 ? 128 : 64;
  const int stage_bytes = (n_atoms + 2) * kAtomBytes;
  const int t_begin = (int)((long)tiles_total * blockIdx.x / gridDim.x);
  const int t_end = (int)((long)tiles_total * (blockIdx.x + 1) / gridDim.x);

  for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // bf16 1.0 x2
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mY);
    tma_prefetch_desc(&mX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();  // the ones tile is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait_parked(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
        uint8_t* sy = smem + stage * stage_bytes;
        uint8_t* sx = sy + n_atoms * kAtomBytes;
        for (int a = 0; a < n_atoms; ++a) tma_load_2d(sy + a * kAtomBytes, &mY, &full[stage], a * 64, tile * kTokTile);
        for (int a = 0; a < 2; ++a) tma_load_2d(sx + a * kAtomBytes, &mX, &full[stage], a * 64, tile * kTokTile);
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_w = umma_idesc_bf16_ex(M, 128, 1, 1);
      const uint32_t idesc_b = umma_idesc_bf16_ex(M, 16, 1, 1);
      const uint64_t d_ones = umma_desc(smem_u32(ones), kAtomBytes, 1024, 2);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait_parked(&full[stage], phase);
        tc_fence_after();
        const uint32_t sy = smem_u32(smem + stage * stage_bytes);
        const uint32_t sx = sy + n_atoms * kAtomBytes;
        const uint64_t db = umma_desc(sx, kAtomBytes, 1024, 2);
        for (int mb = 0; mb < m_blocks; ++mb) {
          // M = 128 spans two atoms along M (LBO = atom stride); a single atom uses the pinned LBO = 16
          const uint64_t da = umma_desc(sy + mb * 2 * kAtomBytes, M == 128 ? kAtomBytes : 16, 1024, 2);
#pragma unroll
          for (int ks = 0; ks < kTokTile / 16; ++ks) {
            const uint32_t acc = (tile > t_begin || ks > 0) ? 1u : 0u;
            // 16 tokens = 16 rows of 128 bytes = 128 sixteen-byte units along the MMA K dimension
            if (!(dbg & 1)) tc_mma_bf16(tmem_base + mb * 128, da + 128 * ks, db + 128 * ks, idesc_w, acc);
            if (!(dbg & 2)) tc_mma_bf16(tmem_base + kBiasCol0 + mb * 16, da + 128 * ks, d_ones, idesc_b, acc);
          }
        }
        tc_commit(&empty[stage]);
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      tc_commit(done);
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    mbar_wait_parked(done, 0);
    tc_fence_after();
    const bool has_work = t_end > t_begin;
    // M = 128: TMEM lane = row; M = 64: row r sits on lane (r % 16) + 32 (r / 16)
    const int row_in_block = M == 128 ? ew * 32 + lane : ew * 16 + lane;
    const bool valid = M == 128 || lane < 16;
    for (int mb = 0; mb < m_blocks; ++mb) {
      const int row = mb * 128 + row_in_block;
      float* out = partial_w + ((long)blockIdx.x * N + row) * 128;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + mb * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = has_work ? make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                              __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(out + c0 + 4 * q) = v;
          }
        }
      }
      uint32_t rb[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(ew * 32) << 16) + kBiasCol0 + mb * 16, rb);
      tmem_ld_wait();
      if (valid) partial_b[(long)blockIdx.x * N + row] = has_work ? __uint_as_float(rb[0]) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  // ---- second stage inside the same launch: grid barrier (all CTAs are co-resident: grid <= #SMs, 1 CTA/SM), then every
  // CTA sums the gridDim.x partials of its slice of the outputs in a fixed order (deterministic, no atomics on data) ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&sync_ctr[0], 1u);
    while (*reinterpret_cast<volatile unsigned int*>(&sync_ctr[0]) < gridDim.x) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
  const int P = gridDim.x;
  const long n_w4 = (long)N * 32;  // float4 outputs
  if (dw != nullptr && !(dbg & 8)) {
    // slice of this CTA: float4 outputs [o_begin, o_end); thread = (output o, partial quarter); each thread streams its
    // quarter of the partials with 8 independent 16-byte loads in flight (bandwidth-, not latency-bound)
    const long o_begin = n_w4 * blockIdx.x / gridDim.x, o_end = n_w4 * (blockIdx.x + 1) / gridDim.x;
    const int col = threadIdx.x & 63, part = threadIdx.x >> 6;
    float4* red = reinterpret_cast<float4*>(smem);
    const float4* pw4 = reinterpret_cast<const float4*>(partial_w);
    for (long base = o_begin; base < o_end; base += 64) {
      const long o = base + col;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o < o_end) {
        const int p0 = P * part / 4, p1 = P * (part + 1) / 4;
        int p = p0;
        for (; p + 8 <= p1; p += 8) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __ldcg(pw4 + (long)(p + j) * n_w4 + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc.x += v[j].x;
            acc.y += v[j].y;
            acc.z += v[j].z;
            acc.w += v[j].w;
          }
        }
        for (; p < p1; ++p) {
          const float4 v = __ldcg(pw4 + (long)p * n_w4 + o);
          acc.x += v.x;
          acc.y += v.y;
          acc.z += v.z;
          acc.w += v.w;
        }
      }
      red[part * 64 + col] = acc;
      __syncthreads();
      if (part == 0 && o < o_end) {
        const float4 a0 = red[col], a1 = red[64 + col], a2 = red[128 + col], a3 = red[192 + col];
        float* d = dw + 4 * o;  // the caller's gradient tensor is only guaranteed 4-byte aligned
        d[0] = (a0.x + a1.x) + (a2.x + a3.x);
        d[1] = (a0.y + a1.y) + (a2.y + a3.y);
        d[2] = (a0.z + a1.z) + (a2.z + a3.z);
        d[3] = (a0.w + a1.w) + (a2.w + a3.w);
      }
      __syncthreads();
    }
  }
  if (db != nullptr) {
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < N; i += gridDim.x * kThreads) {
      float acc = 0.f;
#pragma unroll 8
      for (int p = 0; p < P; ++p) acc += __ldcg(partial_b + (long)p * N + i);
      db[i] = acc;
    }
  }
  // self-resetting counters: the last CTA to leave clears both for the next launch
  if (threadIdx.x == 0) {
    if (atomicAdd(&sync_ctr[1], 1u) == gridDim.x - 1) {
      sync_ctr[0] = 0;
      sync_ctr[1] = 0;
      __threadfence();
    }
  }
}

__device__ unsigned int g_wgrad_sync[2] = {0u, 0u};

int wg_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}


// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 convolution between two 64-channel NHWC maps of width 64 on the same tensor cores:
//     dW[tap][ci][co] = sum_pixels X[p + tap][ci] * dY[p][co]            (autograd of nn.Conv2d, tbsrn.py:232,237; tsrn.py:80-86)
// A = X (MN-major: the channel row of a pixel is 128 contiguous bytes, pixels are the MMA K dimension), B = dY (same).
// A tile is two image rows (128 pixels).  X arrives as three TMA boxes of (2 + 2) rows x 64 pixels, one per horizontal shift
// dx (zero-filled halo); the three vertical taps of a shift are 8 KB-aligned windows of the same box (64 pixels x 128 B per
// row).  Two taps are STACKED along M: the descriptor's leading-dimension offset is simply the byte distance between the two
// windows (8 KB for vertically adjacent taps, the box stride for horizontally adjacent ones), so five M = 128 accumulators of
// 64 columns hold the nine 64 x 64 products (the last pair repeats a tap; its first half is ignored).
//   pair 0..2: taps (dy -1, dx) over (dy 0, dx), dx = -1, 0, +1      pair 3: (dy +1, dx -1) over (dy +1, dx 0)
//   pair 4   : (dy 0, dx +1) [ignored] over (dy +1, dx +1)
// One persistent CTA per SM walks a contiguous range of tiles; partial[cta][tap][co][ci] fp32, summed by a second kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kCvXBox = 4 * 64 * 128;            // one dx box: 4 rows x 64 pixels x 128 B
constexpr int kCvYTile = 128 * 128;              // dY tile: 128 pixels x 128 B
constexpr int kCvStage = 3 * kCvXBox + kCvYTile; // 112 KB
constexpr int kCvStages = 2;
constexpr int kCvThreads = 256;

__global__ void __launch_bounds__(kCvThreads, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mY, int tiles_total,
                        int tiles_per_img, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kCvStages * kCvStage);
  uint64_t* empty = full + kCvStages;
  uint64_t* done = empty + kCvStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = (int)((long)tiles_total * blockIdx.x / gridDim.x);
  const int t_end = (int)((long)tiles_total * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mX);
    tma_prefetch_desc(&mY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kCvStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int b = tile / tiles_per_img, h0 = (tile - b * tiles_per_img) * 2;
        mbar_wait_parked(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)kCvStage);
        uint8_t* st = smem + stage * kCvStage;
        for (int dxi = 0; dxi < 3; ++dxi) tma_load_4d(st + dxi * kCvXBox, &mX, &full[stage], 0, dxi - 1, h0 - 1, b);
        tma_load_4d(st + 3 * kCvXBox, &mY, &full[stage], 0, 0, h0, b);
        if (++stage == kCvStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, one elected lane per instruction (no per-lane predication in the loop body)
    constexpr uint32_t idesc = umma_idesc_bf16_ex(128, 64, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait_parked(&full[stage], phase);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + stage * kCvStage);
      const uint64_t db = umma_desc(st + 3 * kCvXBox, 16, 1024, 2);
      // window of tap (dy, dx): box dx + 1, rows (dy + 1) .. (dy + 2) of the 4-row box
      const uint32_t a0[5] = {st + 0 * kCvXBox, st + 1 * kCvXBox, st + 2 * kCvXBox, st + 0 * kCvXBox + 2 * 8192,
                              st + 2 * kCvXBox + 8192};
      const uint32_t lbo[5] = {8192, 8192, 8192, (uint32_t)kCvXBox, 8192};
      const uint32_t acc = tile > t_begin ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
          const uint64_t da = umma_desc(a0[pr], lbo[pr], 1024, 2);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // 16 pixels = 16 rows of 128 bytes = 128 sixteen-byte units
            tc_mma_bf16(tmem_base + pr * 64, da + 128 * ks, db + 128 * ks, idesc, acc | (ks > 0 ? 1u : 0u));
        }
        tc_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == kCvStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) tc_commit(done);
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp & 3;
    mbar_wait_parked(done, 0);
    tc_fence_after();
    const bool has_work = t_end > t_begin;
    const int row = ew * 32 + lane;            // TMEM lane: rows 0-63 = first tap of the pair, 64-127 = second
    const int half = row >> 6, ci = row & 63;
    // tap index (dy + 1) * 3 + (dx + 1) of each pair's two halves; -1 = ignored
    const int tap_of[5][2] = {{0, 3}, {1, 4}, {2, 5}, {6, 7}, {-1, 8}};
    float* out = partial + (long)blockIdx.x * 9 * 4096;
#pragma unroll 1
    for (int pr = 0; pr < 5; ++pr) {
      const int tap = tap_of[pr][half];
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + pr * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        tmem_ld_wait();
        if (tap >= 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)   // [tap][co][ci]: the 32 lanes of a warp write 32 consecutive ci
            out[(long)tap * 4096 + (c0 + j) * 64 + ci] = has_work ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[co][ci][tap] (torch layout [Co][64][3][3], Co = 64 * groups) = sum_cta partial[cta][tap][co][ci]
__global__ void conv3x3_wgrad_tc_reduce_kernel(const float* __restrict__ partial, int P, int co_mul, int co_add,
                                               float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * 4096) return;
  double s = 0.0;
  int p = 0;
  for (; p + 8 <= P; p += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (long)(p + u) * 9 * 4096 + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; p < P; ++p) s += __ldcg(partial + (long)p * 9 * 4096 + i);
  const int ci = i & 63, co = (i >> 6) & 63, tap = i >> 12;
  dw[((long)(co * co_mul + co_add) * 64 + ci) * 9 + tap] = (float)s;
}


// ------------------------------------------------------------------------------------------------------------------
// The same product for ANY channel counts and map widths (the ResNet encoders of the trainable recognisers: 64 ... 1024 channels,
// 16 x 16 ... 16 x 160 maps): a CTA owns one (64 input channels) x (64 output channels) block of dW for all nine taps and a range
// of 128-pixel tiles; the grid is (channel blocks) x (pixel splits), partial[split][block][tap][co][ci] is summed by a second
// kernel straight into the torch layout.  Replaces the materialised transposed im2col matrix (9 Ci x pixels, 4.25 GB per step at
// batch 64) + plain GEMM of round 1: the shifted operand is formed by TMA.  A tile is bw x (128 / bw) pixels (bw = the map width
// when it is 16 / 32 / 64 / 128, else the largest of 128 / 64 / 32 / 16 dividing it - 32 for the 160-wide maps), the halo box has
// two more rows, the vertical taps are windows bw x 128 B apart.
// ------------------------------------------------------------------------------------------------------------------
struct ConvWgGeom {
  int bw, tile_rows, wblocks, tiles_per_img, tiles_total;
  int ci_blocks, co_blocks, ksplit;
  int xbox_bytes, stage_bytes, stages;
};

__global__ void __launch_bounds__(kCvThreads, 1)
conv3x3_wgrad_tc_general_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mY,
                                const ConvWgGeom g, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
  uint64_t* empty = full + 4;
  uint64_t* done = empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = g.ci_blocks * g.co_blocks;
  const int blk = blockIdx.x % n_blocks, split = blockIdx.x / n_blocks;   // neighbours share the pixel range (L2 reuse)
  const int ci0 = (blk % g.ci_blocks) * 64, co0 = (blk / g.ci_blocks) * 64;
  const int t_begin = (int)((long)g.tiles_total * split / g.ksplit);
  const int t_end = (int)((long)g.tiles_total * (split + 1) / g.ksplit);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mX);
    tma_prefetch_desc(&mY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int b = tile / g.tiles_per_img, r = tile - b * g.tiles_per_img;
        const int h0 = (r / g.wblocks) * g.tile_rows, w0 = (r % g.wblocks) * g.bw;
        mbar_wait_parked(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)g.stage_bytes);
        uint8_t* st = smem + stage * g.stage_bytes;
        for (int dxi = 0; dxi < 3; ++dxi) tma_load_4d(st + dxi * g.xbox_bytes, &mX, &full[stage], ci0, w0 + dxi - 1, h0 - 1, b);
        tma_load_4d(st + 3 * g.xbox_bytes, &mY, &full[stage], co0, w0, h0, b);
        if (++stage == g.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_ex(128, 64, 1, 1);
    const uint32_t win = (uint32_t)g.bw * 128;   // bytes between vertically adjacent tap windows
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait_parked(&full[stage], phase);
      tc_fence_after();
      const uint32_t st = smem_u32(smem + stage * g.stage_bytes);
      const uint32_t xb = (uint32_t)g.xbox_bytes;
      const uint64_t db = umma_desc(st + 3 * xb, 16, 1024, 2);
      const uint32_t a0[5] = {st, st + xb, st + 2 * xb, st + 2 * win, st + 2 * xb + win};
      const uint32_t lbo[5] = {win, win, win, xb, win};
      const uint32_t acc = tile > t_begin ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
          const uint64_t da = umma_desc(a0[pr], lbo[pr], 1024, 2);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            tc_mma_bf16(tmem_base + pr * 64, da + 128 * ks, db + 128 * ks, idesc, acc | (ks > 0 ? 1u : 0u));
        }
        tc_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == g.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) tc_commit(done);
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp & 3;
    mbar_wait_parked(done, 0);
    tc_fence_after();
    const bool has_work = t_end > t_begin;
    const int row = ew * 32 + lane;
    const int half = row >> 6, ci = row & 63;
    const int tap_of[5][2] = {{0, 3}, {1, 4}, {2, 5}, {6, 7}, {-1, 8}};
    float* out = partial + ((long)split * n_blocks + blk) * 9 * 4096;
#pragma unroll 1
    for (int pr = 0; pr < 5; ++pr) {
      const int tap = tap_of[pr][half];
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + pr * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        tmem_ld_wait();
        if (tap >= 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) out[(long)tap * 4096 + (c0 + j) * 64 + ci] = has_work ? __uint_as_float(r[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dw[co][ci][tap] (torch (Co, Ci, 3, 3)) = sum_split partial[split][block][tap][co % 64][ci % 64]
__global__ void conv3x3_wgrad_tc_general_reduce_kernel(const float* __restrict__ partial, int ksplit, int ci_blocks, int co_blocks,
                                                       int Ci, float* __restrict__ dw) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_blocks = ci_blocks * co_blocks;
  if (i >= (long)n_blocks * 9 * 4096) return;
  const long stride = (long)n_blocks * 9 * 4096;
  double s = 0.0;
  int p = 0;
  for (; p + 4 <= ksplit; p += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcg(partial + (long)(p + u) * stride + i);
#pragma unroll
    for (int u = 0; u < 4; ++u) s += v[u];
  }
  for (; p < ksplit; ++p) s += __ldcg(partial + (long)p * stride + i);
  const int e = (int)(i % (9 * 4096)), blk = (int)(i / (9 * 4096));
  const int ci = e & 63, co = (e >> 6) & 63, tap = e >> 12;
  const int cig = (blk % ci_blocks) * 64 + ci, cog = (blk / ci_blocks) * 64 + co;
  dw[((long)cog * Ci + cig) * 9 + tap] = (float)s;
}

bool conv_wg_geom(int B, int H, int W, int Ci, int Co, ConvWgGeom* g) {
  if (Ci % 64 || Co % 64 || Ci < 64 || Co < 64) return false;
  int bw = 0;
  if (W == 16 || W == 32 || W == 64 || W == 128) bw = W;
  else
    for (int c = 128; c >= 16; c >>= 1)
      if (W % c == 0 && H % (128 / c) == 0) {
        bw = c;
        break;
      }
  if (!bw) return false;
  const int tile_rows = 128 / bw;
  if (H % tile_rows != 0) return false;   // (maps smaller than a tile - the 2 x 16 maps of image-ids-CTR - keep the im2col path)
  g->bw = bw;
  g->tile_rows = tile_rows;
  g->wblocks = W / bw;
  g->tiles_per_img = g->wblocks * (H / tile_rows);
  g->tiles_total = B * g->tiles_per_img;
  g->ci_blocks = Ci / 64;
  g->co_blocks = Co / 64;
  const int n_blocks = g->ci_blocks * g->co_blocks;
  int ks = (2 * wg_num_sms() + n_blocks - 1) / n_blocks;   // about two CTAs' worth of work per SM
  if (n_blocks >= wg_num_sms()) ks = 1;
  if (ks > g->tiles_total) ks = g->tiles_total;
  if (ks < 1) ks = 1;
  g->ksplit = ks;
  g->xbox_bytes = (tile_rows + 2) * bw * 128;
  g->stage_bytes = 3 * g->xbox_bytes + kCvYTile;
  g->stages = (220 * 1024) / g->stage_bytes;
  if (g->stages > 4) g->stages = 4;
  return g->stages >= 1;
}

}  // namespace

bool linear_wgrad_tc_supported(long T, int N, int K, long ld_dy, long ld_x) {
  return K == 128 && (N == 64 || N == 128 || N == 256 || N == 384) && T % kTokTile == 0 && T >= kTokTile && ld_dy % 8 == 0 &&
         ld_x % 8 == 0;
}
size_t linear_wgrad_tc_partial_bytes(int N) { return (size_t)wg_num_sms() * N * (128 + 1) * 4; }

// dw fp32 [N][128], db fp32 [N] (either may be null); partial: linear_wgrad_tc_partial_bytes(N) bytes
int linear_wgrad_tc(const bf16* dy, long ld_dy, const bf16* x, long ld_x, long T, int N, float* dw, float* db, float* partial,
                    cudaStream_t s) {
  ProfScope _ps("linear_wgrad", s);
  FOCR_REQUIRE(linear_wgrad_tc_supported(T, N, 128, ld_dy, ld_x), "linear_wgrad_tc: T=%ld N=%d", T, N);
  CUtensorMap mY, mX;
  TRY_RC(focr_make_tmap_2d(&mY, dy, (unsigned long long)N, (unsigned long long)T, (unsigned long long)ld_dy * 2, 64, kTokTile, 128));
  TRY_RC(focr_make_tmap_2d(&mX, x, 128ull, (unsigned long long)T, (unsigned long long)ld_x * 2, 64, kTokTile, 128));
  const int n_atoms = N / 64;
  const int tiles = (int)(T / kTokTile);
  int stages = kSmemRing / ((n_atoms + 2) * kAtomBytes);
  if (stages > kMaxStages) stages = kMaxStages;
  const int grid = tiles < wg_num_sms() ? tiles : wg_num_sms();
  const size_t smem = kSmemRing + kOnesBytes + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(linear_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  float* pw = partial;
  float* pb = partial + (size_t)grid * N * 128;
  static unsigned int* ctr = nullptr;  // resolved once (outside any stream capture: the first step runs eagerly)
  if (ctr == nullptr) FOCR_CHECK_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&ctr), g_wgrad_sync));
  static int dbg = -1;
  if (dbg < 0) dbg = getenv("FOCR_WGRAD_DEBUG") ? atoi(getenv("FOCR_WGRAD_DEBUG")) : 0;  // tuning aid: bit0 skip W MMAs, bit1 skip bias MMAs
  if (dbg & 4) stages = 2;
  linear_wgrad_tc_kernel<<<grid, kThreads, smem, s>>>(mY, mX, n_atoms, tiles, stages, pw, pb, dw, db, ctr, dbg);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// ---- 3x3 conv weight gradient (64 -> 64 channels per group, W = 64 maps) -----------------------------------------------------------
bool conv3x3_wgrad_tc_supported(int H, int W) {
  static int off = -1;
  if (off < 0) off = getenv("FOCR_CONV_WGRAD_LEGACY") ? 1 : 0;   // tuning / cross-check knob: keep the mma.sync kernel
  return !off && W == 64 && H % 2 == 0 && H >= 2;
}
size_t conv3x3_wgrad_tc_partial_bytes() { return (size_t)wg_num_sms() * 9 * 4096 * 4; }

// x (B, H, 64, 64) NHWC bf16; dy: 64-channel map addressed through element strides (dy_pix, dy_row, dy_img) - a plain NHWC
// map, or one sub-pixel plane of the PixelShuffle layout; dw fp32 torch layout: element (co * co_mul + co_add, ci, tap)
int conv3x3_wgrad_tc(const bf16* dy, long dy_pix, long dy_row, long dy_img, const bf16* x, int B, int H, int co_mul, int co_add,
                     float* dw, float* partial, cudaStream_t s) {
  FOCR_REQUIRE(conv3x3_wgrad_tc_supported(H, 64), "conv3x3_wgrad_tc: H=%d", H);
  CUtensorMap mX, mY;
  {
    const unsigned long long dims[4] = {64, 64, (unsigned long long)H, (unsigned long long)B};
    const unsigned long long str[3] = {64 * 2, 64 * 64 * 2, (unsigned long long)H * 64 * 64 * 2};
    const unsigned box[4] = {64, 64, 4, 1};
    TRY_RC(focr_make_tmap_4d(&mX, x, dims, str, box));
  }
  {
    const unsigned long long dims[4] = {64, 64, (unsigned long long)H, (unsigned long long)B};
    const unsigned long long str[3] = {(unsigned long long)dy_pix * 2, (unsigned long long)dy_row * 2, (unsigned long long)dy_img * 2};
    const unsigned box[4] = {64, 64, 2, 1};
    TRY_RC(focr_make_tmap_4d(&mY, dy, dims, str, box));
  }
  const int tiles_per_img = H / 2;
  const int tiles = B * tiles_per_img;
  const int grid = tiles < wg_num_sms() ? tiles : wg_num_sms();
  const size_t smem = (size_t)kCvStages * kCvStage + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  conv3x3_wgrad_tc_kernel<<<grid, kCvThreads, smem, s>>>(mX, mY, tiles, tiles_per_img, partial);
  FOCR_LAUNCH_CHECK();
  conv3x3_wgrad_tc_reduce_kernel<<<focr_cdiv(9 * 4096, 256), 256, 0, s>>>(partial, grid, co_mul, co_add, dw);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

// ---- 3x3 conv weight gradient, any channel counts (multiples of 64) and map widths that cut into 128-pixel boxes ------------------
bool conv3x3_wgrad_tc_general_supported(int B, int H, int W, int Ci, int Co) {
  static int off = -1;
  if (off < 0) off = getenv("FOCR_CONV_WGRAD_IM2COL") ? 1 : 0;   // tuning / cross-check knob: keep the im2col + GEMM path
  ConvWgGeom g;
  return !off && conv_wg_geom(B, H, W, Ci, Co, &g);
}
size_t conv3x3_wgrad_tc_general_partial_bytes(int B, int H, int W, int Ci, int Co) {
  ConvWgGeom g;
  if (!conv_wg_geom(B, H, W, Ci, Co, &g)) return 0;
  return (size_t)g.ksplit * g.ci_blocks * g.co_blocks * 9 * 4096 * 4;
}
// x (B, H, W, Ci), dy (B, H, W, Co) NHWC bf16; dw fp32 (Co, Ci, 3, 3)
int conv3x3_wgrad_tc_general(const bf16* dy, const bf16* x, int B, int H, int W, int Ci, int Co, float* dw, float* partial,
                             cudaStream_t s) {
  ConvWgGeom g;
  FOCR_REQUIRE(conv_wg_geom(B, H, W, Ci, Co, &g), "conv3x3_wgrad_tc_general: B=%d H=%d W=%d Ci=%d Co=%d", B, H, W, Ci, Co);
  CUtensorMap mX, mY;
  {
    const unsigned long long dims[4] = {(unsigned long long)Ci, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
    const unsigned long long str[3] = {(unsigned long long)Ci * 2, (unsigned long long)W * Ci * 2, (unsigned long long)H * W * Ci * 2};
    const unsigned box[4] = {64, (unsigned)g.bw, (unsigned)(g.tile_rows + 2), 1};
    TRY_RC(focr_make_tmap_4d(&mX, x, dims, str, box));
  }
  {
    const unsigned long long dims[4] = {(unsigned long long)Co, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
    const unsigned long long str[3] = {(unsigned long long)Co * 2, (unsigned long long)W * Co * 2, (unsigned long long)H * W * Co * 2};
    const unsigned box[4] = {64, (unsigned)g.bw, (unsigned)g.tile_rows, 1};
    TRY_RC(focr_make_tmap_4d(&mY, dy, dims, str, box));
  }
  const size_t smem = (size_t)g.stages * g.stage_bytes + 256 + 1024;
  static size_t attr = 0;
  if (smem > attr) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_blocks = g.ci_blocks * g.co_blocks;
  conv3x3_wgrad_tc_general_kernel<<<n_blocks * g.ksplit, kCvThreads, smem, s>>>(mX, mY, g, partial);
  FOCR_LAUNCH_CHECK();
  const long n = (long)n_blocks * 9 * 4096;
  conv3x3_wgrad_tc_general_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(partial, g.ksplit, g.ci_blocks, g.co_blocks, Ci, dw);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
