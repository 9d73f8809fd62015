// tcgen05 + TMA implicit-GEMM engine (see tc_gemm.cuh for what it replaces in the reference).
//
// Persistent, warp-specialised CTA (256 threads, 1 CTA/SM):
//   warp 0   : TMA producer  (A tile: 4-D box {64ch, W, 128/W rows, 1 image}, zero-filled halo;
//                             B tile: 2-D box {64, BLOCK_N} of the per-tap weight matrix)
//   warp 1   : MMA issuer    (one lane; 4 x tcgen05.mma 128xBLOCK_Nx16 per 64-wide K block)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld 32x32b -> bias/relu/residual/pixel-shuffle -> global)
// Pipelines: smem ring full/empty (TMA<->MMA) and a 2-deep TMEM accumulator ring (MMA<->epilogue).
#include "tc_gemm.cuh"

#include <mutex>
#include <unordered_map>

namespace {

constexpr int kThreads = 384;   // producer, MMA, TMEM-alloc, spare + 2 x 4 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kTileM * 128;
constexpr int kCdBlk = kTileM * 128;  // one 128-row x 64-column bf16 block, 128-byte swizzled (TMA box)

// One persistent CTA per SM.  Shared memory: K-block ring (A+B per stage), a double-buffered output tile
// that the epilogue fills and a TMA store drains, and one auxiliary input tile (residual or gate) that the
// producer prefetches with TMA, so the epilogue warps never touch global memory.
// Resident-weights mode (BLOCK_N = 64, one N block, <= 9 K blocks: every 64->64 conv and the 128->64 linear): the whole
// weight operand (<= 72 KB) is loaded into shared memory ONCE per CTA and the ring carries A tiles only.  The kernel is
// bound by the L2 -> shared-memory fill rate (a 128-pixel tile needs 9 x 16 KB of shifted A views), so not re-fetching
// 9 x 8 KB of weights per tile removes a third of that traffic.
constexpr int kResMaxKb = 9;
template <int BLOCK_N>
struct TcCfg {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 128 ? 3 : 6;
  static constexpr int kResBytes = kResMaxKb * kBBytes;         // resident weights (BLOCK_N == 64 only)
  static constexpr int kResStages = 5;                           // A-only ring behind the resident weights
  static constexpr int kCdBytes = (BLOCK_N / 64) * kCdBlk;  // per buffer
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kResRingBytes = kResBytes + kResStages * kABytes;
  static constexpr int kOffCd = (BLOCK_N == 64 && kResRingBytes > kRingBytes) ? kResRingBytes : kRingBytes;
  static constexpr int kOffAux = kOffCd + 2 * kCdBytes;
  static constexpr int kOffBar = kOffAux + 2 * kCdBytes;  // aux tile double-buffered like the output tile
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 2 accumulator stages
  static constexpr int kBiasFloats = 384;  // bias vector staged in shared memory when N_total fits (every TBSRN layer)
  static constexpr int kSmemBytes = kOffBar + 512 /*barriers*/ + kBiasFloats * 4 + 1024 /*align*/;
};

// byte offset of (row, 16-byte chunk) inside a 128-byte-swizzled 128x64 bf16 block
__device__ __forceinline__ uint32_t cd_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
               const __grid_constant__ CUtensorMap mA2, const __grid_constant__ CUtensorMap mA3,
               const __grid_constant__ CUtensorMap mB, const __grid_constant__ CUtensorMap mC,
               const __grid_constant__ CUtensorMap mR, const TcGemmParams p) {
  using Cfg = TcCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;   // [2] auxiliary (residual / gate) tile landed
  uint64_t* aempty = afull + 2;   // [2] ... and has been consumed by the 4 epilogue warps
  uint64_t* bfull = aempty + 2;   // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);
  const bool resident = BLOCK_N == 64 && p.b_resident != 0;
  // Column mode (3x3, 64 -> 64, resident weights, no auxiliary tile): one TMA box of (tile rows + 2) image rows at
  // horizontal shift dx serves the three taps dy = -1, 0, +1 of that column as 1024-byte-aligned windows of the same
  // shared-memory tile (window offset = dy * W pixels * 128 B), so a tile needs 3 x (128 + 2W) pixel rows of fills
  // instead of 9 x 128.  The ring then also takes over the unused auxiliary-tile region.
  const bool colmode = resident && p.col_mode != 0;
  uint8_t* ring = smem + (resident ? Cfg::kResBytes : 0);
  const int col_bytes = (kTileM + 2 * p.W) * 128;
  const int stage_bytes = colmode ? col_bytes : (resident ? kABytes : Cfg::kStageBytes);
  int n_stages = resident ? Cfg::kResStages : Cfg::kStages;
  if (colmode) {
    n_stages = (Cfg::kOffAux - Cfg::kResBytes) / col_bytes;
    if (n_stages > Cfg::kStages) n_stages = Cfg::kStages;
  }
  uint8_t* cd_base = smem + ((BLOCK_N == 64 && p.b_resident != 0 && p.col_mode != 0) ? Cfg::kOffAux : Cfg::kOffCd);
  uint8_t* aux_base = smem + Cfg::kOffAux;
  const bool use_aux = p.tma_out && (p.residual != nullptr || p.gate != nullptr);
  // the epilogue adds the bias to every tile: reading it from global memory costs a long-scoreboard stall per 32-column
  // chunk (ncu source view: the FADDs behind those loads were the top stall sites of the epilogue warps)
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kOffBar + 512);
  const bool bias_in_smem = p.bias != nullptr && p.n_total <= Cfg::kBiasFloats;
  if (bias_in_smem)
    for (int i = threadIdx.x; i < p.n_total; i += kThreads) sbias[i] = p.bias[i];
  const float* bias_src = bias_in_smem ? sbias : p.bias;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.kh * p.kw;
  const int pad_h = p.kh >> 1, pad_w = p.kw >> 1;
  const int nk = taps * p.chunks;
  const int total_tiles = p.m_tiles * p.n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mA0);
    tma_prefetch_desc(&mB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiThreads / 32);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&afull[a], 1);
      mbar_init(&aempty[a], kEpiThreads / 32);
    }
    mbar_init(bfull, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* amaps[4] = {&mA0, &mA1, &mA2, &mA3};
      int stage = 0;
      uint32_t phase = 0;
      int ait = 0;
      const int hw = p.H * p.W;
      if (resident) {  // the whole weight operand, once
        mbar_arrive_expect_tx(bfull, (uint32_t)(nk * Cfg::kBBytes));
        for (int kb = 0; kb < nk; ++kb)
          tma_load_2d(smem + kb * Cfg::kBBytes, &mB, bfull, (kb % p.chunks) * kChunkK, (kb / p.chunks) * p.n_total);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m = tile / p.n_blocks, nb = tile % p.n_blocks;
        const int p0 = m * kTileM;
        int b = p0 / hw;
        int h0 = (p0 - b * hw) / p.W;
        int w0 = 0;
        if (p.wblocks > 1) {  // bw x (128 / bw)-pixel boxes, wblocks side by side
          b = m / p.tiles_per_img;
          const int r = m - b * p.tiles_per_img;
          h0 = (r / p.wblocks) * (kTileM / p.bw);
          w0 = (r % p.wblocks) * p.bw;
        }
        if (use_aux) {
          const int ab = ait & 1;
          mbar_wait(&aempty[ab], ((ait >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&afull[ab], Cfg::kCdBytes);
          for (int blk = 0; blk < BLOCK_N / 64; ++blk) {
            if (p.wblocks > 1)
              tma_load_4d(aux_base + ab * Cfg::kCdBytes + blk * kCdBlk, &mR, &afull[ab], nb * BLOCK_N + blk * 64, w0, h0, b);
            else
              tma_load_2d(aux_base + ab * Cfg::kCdBytes + blk * kCdBlk, &mR, &afull[ab], nb * BLOCK_N + blk * 64, p0);
          }
          ++ait;
        }
        if (colmode) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
            tma_load_4d(ring + stage * stage_bytes, &mA1, &full[stage], 0, dxi - 1, h0 - 1, b);
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          continue;
        }
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = tap / p.kw - pad_h, dx = tap % p.kw - pad_w;
          for (int ch = 0; ch < p.chunks; ++ch) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
            uint8_t* sa = ring + stage * stage_bytes;
            const int mi = ch / p.chunks_per_map;
            const int c0 = (ch - mi * p.chunks_per_map) * kChunkK;
            tma_load_4d(sa, amaps[mi], &full[stage], c0, w0 + dx, h0 + dy, b);
            if (!resident)
              tma_load_2d(sa + kABytes, &mB, &full[stage], ch * kChunkK, tap * p.n_total + nb * BLOCK_N);
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t aphase = 0;
      if (resident) {
        mbar_wait(bfull, 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        if (colmode) {
          for (int dxi = 0; dxi < 3; ++dxi) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * stage_bytes);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const uint64_t da = umma_desc_k_sw128(sa + dyi * p.W * 128);
              const uint64_t db = umma_desc_k_sw128(smem_u32(smem + (dyi * 3 + dxi) * Cfg::kBBytes));
#pragma unroll
              for (int k = 0; k < kChunkK / 16; ++k)
                tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (dxi | dyi | k) != 0);
            }
            tc_commit(&empty[stage]);
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          tc_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) aphase ^= 1;
          continue;
        }
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * stage_bytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(resident ? smem_u32(smem + kb * Cfg::kBBytes) : sa + kABytes);
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in 16-byte units
            tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty[stage]);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const uint32_t dkey = p.drop_key ^ ((p.drop_thresh16 != 0 && p.drop_seed_dev != nullptr) ? __ldg(p.drop_seed_dev) : 0u);
    const int ew = warp & 3;        // TMEM lane quarter this warp may access (warp % 4)
    const int egrp = (warp - 4) >> 2;  // epilogue group 0 / 1: interleaved 32-column chunks
    const int row = ew * 32 + lane;
    int acc = 0;
    uint32_t aphase = 0;
    int it = 0;
    const int hw = p.H * p.W;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m = tile / p.n_blocks, nb = tile % p.n_blocks;
      long mg = (long)m * kTileM + row;
      int tb = 0, th0 = 0, tw0 = 0;
      if (p.wblocks > 1) {  // true pixel index of this thread's row inside a bw-wide box
        tb = m / p.tiles_per_img;
        const int r = m - tb * p.tiles_per_img;
        th0 = (r / p.wblocks) * (kTileM / p.bw);
        tw0 = (r % p.wblocks) * p.bw;
        mg = ((long)tb * p.H + th0 + row / p.bw) * p.W + tw0 + row % p.bw;
      }
      if (p.tma_out) {
        // ---- TMEM -> registers -> swizzled smem tile -> TMA store.  No global access from these warps. ------------
        uint8_t* cd = cd_base + (it & 1) * Cfg::kCdBytes;
        if (warp == 4 && lane == 0) bulk_wait_read<1>();  // the store issued two tiles ago has finished reading `cd`
        named_bar_sync(1, kEpiThreads);
        mbar_wait(&tfull[acc], aphase);
        tc_fence_after();
        const uint8_t* aux = aux_base + (it & 1) * Cfg::kCdBytes;
        if (use_aux) mbar_wait(&afull[it & 1], (it >> 1) & 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
        for (int c0 = egrp * 32; c0 < BLOCK_N; c0 += 64) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c0, r);
          tmem_ld_wait();
          if (c0 + 64 >= BLOCK_N) {  // this warp's last chunk: its share of the accumulator is drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          const int n0 = nb * BLOCK_N + c0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias_src + n0 + j);
              v[j] += bb.x;
              v[j + 1] += bb.y;
              v[j + 2] += bb.z;
              v[j + 3] += bb.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.drop_thresh16 != 0) {
            const uint32_t c0h = (uint32_t)((mg * p.ldc + n0) >> 1);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const uint32_t hsh = drop_hash32(dkey, c0h + q);
              v[2 * q] = (hsh & 0xFFFFu) >= p.drop_thresh16 ? v[2 * q] * p.drop_scale : 0.f;
              v[2 * q + 1] = (hsh >> 16) >= p.drop_thresh16 ? v[2 * q + 1] * p.drop_scale : 0.f;
            }
          }
          const int blk = c0 >> 6, kc = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t so = (uint32_t)(blk * kCdBlk) + cd_off(row, kc + q);
            if (use_aux) {
              const uint4 av = *reinterpret_cast<const uint4*>(aux + so);
              const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(aw[j]);
                if (p.residual == nullptr) {
                  v[q * 8 + 2 * j] = f.x > 0.f ? v[q * 8 + 2 * j] * p.gate_scale : 0.f;
                  v[q * 8 + 2 * j + 1] = f.y > 0.f ? v[q * 8 + 2 * j + 1] * p.gate_scale : 0.f;
                } else {
                  v[q * 8 + 2 * j] += f.x;
                  v[q * 8 + 2 * j + 1] += f.y;
                }
              }
            }
            if (p.residual != nullptr && p.gate != nullptr) {
              // both a residual (through the TMA-prefetched tile) and a gate: the gate row comes straight from global
              const uint4 gv = *reinterpret_cast<const uint4*>(p.gate + mg * p.ldc + n0 + q * 8);
              const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(gw[j]);
                v[q * 8 + 2 * j] = f.x > 0.f ? v[q * 8 + 2 * j] * p.gate_scale : 0.f;
                v[q * 8 + 2 * j + 1] = f.y > 0.f ? v[q * 8 + 2 * j + 1] * p.gate_scale : 0.f;
              }
            }
            if (p.relu_post) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[q * 8 + j] = fmaxf(v[q * 8 + j], 0.f);
            }
            uint4 o;
            o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            *reinterpret_cast<uint4*>(cd + so) = o;
          }
        }
        if (use_aux) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&aempty[it & 1]);
        }
        fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA (async proxy)
        named_bar_sync(1, kEpiThreads);
        if (warp == 4 && lane == 0) {
          for (int blk = 0; blk < BLOCK_N / 64; ++blk) {
            if (p.wblocks > 1)
              tma_store_4d(&mC, cd + blk * kCdBlk, nb * BLOCK_N + blk * 64, tw0, th0, tb);
            else
              tma_store_2d(&mC, cd + blk * kCdBlk, nb * BLOCK_N + blk * 64, m * kTileM);
          }
          bulk_commit();
        }
        ++it;
        acc ^= 1;
        if (acc == 0) aphase ^= 1;
        continue;
      }
      mbar_wait(&tfull[acc], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c0 = egrp * 32; c0 < BLOCK_N; c0 += 64) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        tmem_ld_wait();
        const int n0 = nb * BLOCK_N + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 bb = *reinterpret_cast<const float4*>(bias_src + n0 + j);
            v[j] += bb.x;
            v[j + 1] += bb.y;
            v[j + 2] += bb.z;
            v[j + 3] += bb.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.drop_thresh16 != 0) {
          const uint32_t c0h = (uint32_t)((mg * p.ldc + n0) >> 1);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const uint32_t hsh = drop_hash32(dkey, c0h + q);
            v[2 * q] = (hsh & 0xFFFFu) >= p.drop_thresh16 ? v[2 * q] * p.drop_scale : 0.f;
            v[2 * q + 1] = (hsh >> 16) >= p.drop_thresh16 ? v[2 * q + 1] * p.drop_scale : 0.f;
          }
        }
        if (p.gate != nullptr) {
          const uint4* gp = reinterpret_cast<const uint4*>(p.gate + mg * p.ldc + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 gv = gp[q];
            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = unpack_bf16x2(gw[j]);
              v[q * 8 + 2 * j] = f.x > 0.f ? v[q * 8 + 2 * j] * p.gate_scale : 0.f;
              v[q * 8 + 2 * j + 1] = f.y > 0.f ? v[q * 8 + 2 * j + 1] * p.gate_scale : 0.f;
            }
          }
        }
        if (p.epi == TC_EPI_BF16) {
          const long off = mg * p.ldc + n0;
          if (p.prelu_slope != nullptr) {
            if (p.out2 != nullptr) {
              uint4* op2 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out2) + off);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 o;
                o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                op2[q] = o;
              }
            }
            const float slope = p.prelu_slope[0];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              // differentiate what was stored: the pre-activation is kept in bf16
              const float xr = __bfloat162float(__float2bfloat16_rn(v[j]));
              v[j] = xr > 0.f ? xr : slope * xr;
            }
          }
          if (p.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 rv = rp[q];
              float2 f;
              f = unpack_bf16x2(rv.x); v[q * 8 + 0] += f.x; v[q * 8 + 1] += f.y;
              f = unpack_bf16x2(rv.y); v[q * 8 + 2] += f.x; v[q * 8 + 3] += f.y;
              f = unpack_bf16x2(rv.z); v[q * 8 + 4] += f.x; v[q * 8 + 5] += f.y;
              f = unpack_bf16x2(rv.w); v[q * 8 + 6] += f.x; v[q * 8 + 7] += f.y;
            }
          }
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            op[q] = o;
          }
        } else if (p.epi == TC_EPI_F32) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + mg * p.ldc + n0);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            op[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {  // TC_EPI_PIXSHUF: n = sub*64 + c, sub = 2*i + j -> HR pixel (2h+i, 2w+j)
          const int sub = n0 >> 6, c = n0 & 63;
          const int b = (int)(mg / hw);
          const int rem = (int)(mg - (long)b * hw);
          const int h = rem / p.W, w = rem - h * p.W;
          const long hp = ((long)b * 2 * p.H + 2 * h + (sub >> 1)) * (2 * p.W) + 2 * w + (sub & 1);
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + hp * 64 + c);
          uint4* op2 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out2) + hp * 64 + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o, o2;
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // round the pre-activation to bf16 first so that backward (which re-reads it)
              // differentiates exactly the function that forward evaluated
              a[j] = __bfloat162float(__float2bfloat16_rn(v[q * 8 + j]));
            }
            o.x = pack_bf16x2(a[0], a[1]);
            o.y = pack_bf16x2(a[2], a[3]);
            o.z = pack_bf16x2(a[4], a[5]);
            o.w = pack_bf16x2(a[6], a[7]);
            o2.x = pack_bf16x2(mish_f(a[0]), mish_f(a[1]));
            o2.y = pack_bf16x2(mish_f(a[2]), mish_f(a[3]));
            o2.z = pack_bf16x2(mish_f(a[4]), mish_f(a[5]));
            o2.w = pack_bf16x2(mish_f(a[6]), mish_f(a[7]));
            op[q] = o;
            op2[q] = o2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) aphase ^= 1;
    }
    if (p.tma_out && warp == 4 && lane == 0) bulk_wait<0>();  // all output tiles are in global memory
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point fetched through the runtime: no libcuda link)
// ------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(f);
  });
  return fn;
}

int make_map(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims,
             const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) {
    focr_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return FOCR_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    focr_set_error("cuTensorMapEncodeTiled failed (%d), rank %d dims %llu %llu", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1]);
    return FOCR_ERR_CUDA;
  }
  return FOCR_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N>
int launch_impl(const CUtensorMap* am, const CUtensorMap& bm, const CUtensorMap& cm, const CUtensorMap& rm,
                const TcGemmParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_blocks;
  const int grid = total < num_sms() ? total : num_sms();
  tc_gemm_kernel<BLOCK_N><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(am[0], am[1], am[2], am[3], bm, cm, rm, p);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

}  // namespace

int focr_make_tmap_2d(CUtensorMap* out, const void* base, unsigned long long inner, unsigned long long outer,
                      unsigned long long row_bytes, unsigned box_inner, unsigned box_outer, int swizzle_bytes) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) {
    focr_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return FOCR_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t str[1] = {row_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    focr_set_error("cuTensorMapEncodeTiled failed (%d) for a %llu x %llu map", (int)r, inner, outer);
    return FOCR_ERR_CUDA;
  }
  return FOCR_OK;
}

// 4-D bf16 NHWC map {channels, W, H, B} with element strides and a 128-byte-swizzled box (shared with wgrad_tc.cu)
int focr_make_tmap_4d(CUtensorMap* out, const void* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                      const unsigned box[4]) {
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  return make_map(out, base, 4, d, st, bx);
}

int tc_gemm_block_n(int n_total) {
  if (n_total % 128 == 0) return 128;
  return 64;
}

int tc_gemm_launch(const bf16* const* a_ptrs, int n_amaps, long a_pix_stride, long a_row_stride,
                   long a_img_stride, int a_channels, int B, const bf16* w, int cin, TcGemmParams p,
                   cudaStream_t stream) {
  FOCR_REQUIRE(n_amaps >= 1 && n_amaps <= 4, "tc_gemm: n_amaps %d", n_amaps);
  // a 128-pixel tile is (128 / W) image rows, or - for maps smaller than one tile (H * W < 128) - several whole images:
  // the TMA box then extends over the batch dimension, the zero-filled halo still being per image.  Widths that are not
  // one of 16 / 32 / 64 / 128 (the 16 x 160 maps of the 32 x 320 recogniser crops) are cut into bw-wide boxes of 128 / bw rows.
  int bw = p.W;
  if (!(p.W == 16 || p.W == 32 || p.W == 64 || p.W == 128)) {
    bw = 0;
    for (int c = 128; c >= 16; c >>= 1)
      if (p.W % c == 0 && p.H % (kTileM / c) == 0) {
        bw = c;
        break;
      }
    FOCR_REQUIRE(bw != 0, "tc_gemm: a %dx%d map does not tile into 128-pixel boxes (W must be a multiple of 16 / 32 / 64 / 128 "
                 "with H a multiple of 8 / 4 / 2 / 1)", p.H, p.W);
  }
  p.bw = bw;
  p.wblocks = p.W / bw;
  const int tile_rows = (p.H * bw >= kTileM) ? kTileM / bw : p.H;
  const int tile_imgs = kTileM / (bw * tile_rows);
  FOCR_REQUIRE(p.H % tile_rows == 0 && tile_imgs * tile_rows * bw == kTileM && B % tile_imgs == 0,
               "tc_gemm: a %dx%d map (batch %d) does not tile into 128-pixel boxes", p.H, p.W, B);
  p.tiles_per_img = p.wblocks * (p.H / tile_rows);
  FOCR_REQUIRE(p.wblocks == 1 || (p.epi == TC_EPI_BF16 && p.prelu_slope == nullptr && tile_imgs == 1),
               "tc_gemm: %d-wide maps need the bf16 TMA epilogue", p.W);
  FOCR_REQUIRE(!p.relu_post || (p.epi == TC_EPI_BF16 && p.prelu_slope == nullptr), "tc_gemm: relu_post needs the bf16 TMA epilogue");
  FOCR_REQUIRE(cin % 64 == 0 && a_channels % 64 == 0, "tc_gemm: channels must be multiples of 64");
  FOCR_REQUIRE(p.kh >= 1 && p.kw >= 1 && (p.kh & 1) && (p.kw & 1) && p.kh <= 9 && p.kw <= 9, "tc_gemm: taps %dx%d",
               p.kh, p.kw);
  FOCR_REQUIRE(p.n_total % 64 == 0, "tc_gemm: N %d", p.n_total);
  const int block_n = tc_gemm_block_n(p.n_total);
  p.n_blocks = p.n_total / block_n;
  p.m_tiles = (int)((long)B * p.H * p.W / kTileM);
  p.chunks = cin / 64;
  p.chunks_per_map = a_channels / 64;
  FOCR_REQUIRE(p.chunks_per_map * n_amaps == p.chunks, "tc_gemm: %d maps x %d ch != cin %d", n_amaps,
               a_channels, cin);
  if (p.epi == TC_EPI_PIXSHUF) FOCR_REQUIRE(p.n_total == 256, "pixel-shuffle epilogue needs N=256");

  CUtensorMap am[4];
  for (int i = 0; i < 4; ++i) {
    const bf16* base = a_ptrs[i < n_amaps ? i : 0];
    cuuint64_t dims[4] = {(cuuint64_t)a_channels, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)a_pix_stride * 2, (cuuint64_t)a_row_stride * 2,
                         (cuuint64_t)a_img_stride * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)tile_rows, (cuuint32_t)tile_imgs};
    int rc = make_map(&am[i], base, 4, dims, str, box);
    if (rc) return rc;
  }
  CUtensorMap bm;
  {
    const int taps = p.kh * p.kw;
    cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)taps * p.n_total};
    cuuint64_t str[1] = {(cuuint64_t)cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    int rc = make_map(&bm, w, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    static int no_res = -1;
    if (no_res < 0) no_res = getenv("FOCR_TC_NO_RESIDENT") ? 1 : 0;  // tuning knob
    p.b_resident = (!no_res && block_n == 64 && p.n_blocks == 1 && p.kh * p.kw * p.chunks <= kResMaxKb &&
                    p.m_tiles >= 2 * num_sms()) ? 1 : 0;
  }
  p.col_mode = 0;
  // output / auxiliary tiles through TMA for the plain bf16 epilogue (every hot GEMM); the rare epilogues
  // (fp32 output, PixelShuffle scatter, PReLU with saved pre-activation) keep direct stores
  p.tma_out = (p.epi == TC_EPI_BF16 && p.prelu_slope == nullptr) ? 1 : 0;
  FOCR_REQUIRE(p.tma_out || !(p.residual && p.gate), "tc_gemm: residual + gate needs the bf16 TMA epilogue");
  CUtensorMap cm = bm, rm = bm;
  if (p.tma_out && p.wblocks > 1) {  // output / auxiliary tiles are bw x tile_rows boxes of the (B, H, W, ldc) map
    cuuint64_t dims[4] = {(cuuint64_t)p.n_total, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)p.ldc * 2, (cuuint64_t)p.W * p.ldc * 2, (cuuint64_t)p.H * p.W * p.ldc * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)tile_rows, 1};
    int rc = make_map(&cm, p.out, 4, dims, str, box);
    if (rc) return rc;
    const void* aux = p.residual ? (const void*)p.residual : (const void*)p.gate;
    if (aux) {
      rc = make_map(&rm, aux, 4, dims, str, box);
      if (rc) return rc;
    }
  } else if (p.tma_out) {
    cuuint64_t dims[2] = {(cuuint64_t)p.n_total, (cuuint64_t)p.m_tiles * kTileM};
    cuuint64_t str[1] = {(cuuint64_t)p.ldc * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)kTileM};
    int rc = make_map(&cm, p.out, 2, dims, str, box);
    if (rc) return rc;
    const void* aux = p.residual ? (const void*)p.residual : (const void*)p.gate;
    if (aux) {
      rc = make_map(&rm, aux, 2, dims, str, box);
      if (rc) return rc;
    }
  }
  {
    static int no_col = -1;
    if (no_col < 0) no_col = getenv("FOCR_TC_NO_COLMODE") ? 1 : 0;  // tuning knob
    if (!no_col && p.b_resident && p.tma_out && p.kh == 3 && p.kw == 3 && p.chunks == 1 && n_amaps == 1 && !p.residual &&
        !p.gate && tile_imgs == 1 && p.wblocks == 1) {
      cuuint64_t dims[4] = {(cuuint64_t)a_channels, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)B};
      cuuint64_t str[3] = {(cuuint64_t)a_pix_stride * 2, (cuuint64_t)a_row_stride * 2, (cuuint64_t)a_img_stride * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)p.W, (cuuint32_t)(kTileM / p.W + 2), 1};
      int rc = make_map(&am[1], a_ptrs[0], 4, dims, str, box);
      if (rc) return rc;
      p.col_mode = 1;
    }
  }
  const char* scope = p.kh * p.kw == 1 ? "tc_linear" : (p.kh == 3 && p.kw == 3 ? "tc_conv3x3" : "tc_conv9tap");
  ProfScope _ps(scope, stream);
  if (block_n == 128) return launch_impl<128>(am, bm, cm, rm, p, stream);
  return launch_impl<64>(am, bm, cm, rm, p, stream);
}
