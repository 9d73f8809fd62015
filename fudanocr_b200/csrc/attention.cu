// Fused self-attention of the TBSRN FeatureEnhancer on the 5th-generation tensor cores: h = 4 heads, d_k = 32,
// 1024 tokens (scene-text-telescope/model/tbsrn.py:109-150: softmax(QK^T/sqrt(d_k)) -> dropout(0.1) -> PV).
// The reference materialises P = (B,4,1024,1024) fp32 (4 GiB at B = 256); here P only ever exists as 128 x 64 tiles
// in tensor memory / shared memory.  Input is the packed projection QKV (T,384) bf16 = [q | k | v], head h at columns
// h*32 of each third; output O (T,128) bf16 in the "concat heads" layout the out-projection reads.
//
// One CTA per (batch, head), 384 threads:
//   warp 0      TMA producer (64-byte-swizzled boxes of the [1024][32] head slices)
//   warp 1, 2   one tcgen05.mma issuing thread per softmax group
//   warp 3      TMEM allocator
//   warps 4-7 / 8-11   softmax groups 0 / 1: thread = one query row of a 128-row tile (TMEM lane = row), so row
//               max / row sum / LSE / D are thread-local scalars and no shuffles are needed
// Tiles are 128 queries x 64 keys.  S = Q K^T and dP = dO V^T are K-major x K-major MMAs (d_k = 32 = two K steps),
// accumulators in TMEM; the softmax threads read them with tcgen05.ld, write bf16 P / dS rows into a 128-byte-swizzled
// shared tile, and the second GEMM of each tile consumes that tile either K-major (P V, dS K) or MN-major (P^T dO,
// dS^T Q - M = 64 keys) with V / K / dO / Q read in place as MN-major B operands.  The descriptor semantics used
// here are pinned by tests/test_gpu_umma_layouts.py.
//   forward   : two passes over the keys per query tile (exact row max first, then exp / sum / PV) - no rescaling of
//               the TMEM accumulator is ever needed; S is double-buffered in TMEM
//   backward A: dQ  (query tile outer, key tiles inner; also writes D = rowsum(dO o O))
//   backward B: dK, dV (64-key block outer, query tiles inner)
// Dropout: counter hash of common.cuh keyed on (b,h,q,k/4); each 32-bit hash carries four 7-bit lanes (the low 7
// bits of its bytes), element k is DROPPED when lane (k & 3) < th7, th7 = round(p * 128) - the rate is quantised to
// 1/128 (p = 0.1 -> 13/128) and the 1/(1-p) rescale uses the quantised rate, so the expectation is exact.  The
// forward can store its keep decisions, one word per (q, 32 keys): word [bh][k/32][q], bit 8 (k&3) + (k%32)/4; the
// backward then reads bits instead of re-hashing.  The 1/(1-p) factor is folded into the output scales: P V and dV accumulate kept probabilities
// unscaled, dS = P o (keep o dP - D (1-p)) / (1-p).
#include "kernels.cuh"

namespace {

constexpr int kS = 1024;      // tokens
constexpr int kLdQkv = 384;   // row stride of the packed projection
constexpr int kLdO = 128;
constexpr float kScale = 0.17677669529663687f;  // 1/sqrt(32)
constexpr float kScaleLog2 = kScale * 1.4426950408889634f;
constexpr int kThreads = 384;
constexpr int kHeadBytes = kS * 64;  // one [1024][32] bf16 head slice
constexpr int kQTile = 128 * 64;     // [128][32] bf16
constexpr int kPTile = 128 * 128;    // [128 q][64 keys] bf16, 128-byte swizzled

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte permute with sign replication: selector nibble 8+i yields 0xFF if the msb of source byte i is set, else 0x00
__device__ __forceinline__ uint32_t prmt(uint32_t x, uint32_t sel) {
  uint32_t y;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(y) : "r"(x), "r"(0u), "r"(sel));
  return y;
}
__device__ __forceinline__ uint32_t p_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// operand descriptors: [rows][32] bf16 tiles written by TMA with the 64-byte swizzle (8-row groups 512 B apart) and
// the [128][64] bf16 P / dS tile with the 128-byte swizzle (8-row groups 1024 B apart).  The same fields serve the
// K-major and the MN-major reading of a tile; the instruction descriptor says which one is meant.
__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr) { return umma_desc(addr, 16, 512, 4); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) { return umma_desc(addr, 16, 1024, 2); }

// S / dP tile: D[128 x 64] = A[128 x 32] B[64 x 32]^T
__device__ __forceinline__ void mma_qk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr) {
  constexpr uint32_t idesc = umma_idesc_bf16_ex(128, 64, 0, 0);
  const uint64_t da = desc_sw64(a_addr), db = desc_sw64(b_addr);
  tc_mma_bf16(tmem_d, da, db, idesc, 0);
  tc_mma_bf16(tmem_d, da + 2, db + 2, idesc, 1);
}
// D[128 x 32] (+)= A[128 q x 64 keys] (K-major, sw128) * B[64 keys x 32] (MN-major, sw64)
__device__ __forceinline__ void mma_pv(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool acc) {
  constexpr uint32_t idesc = umma_idesc_bf16_ex(128, 32, 0, 1);
  const uint64_t da = desc_sw128(a_addr), db = desc_sw64(b_addr);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) tc_mma_bf16(tmem_d, da + 2 * ks, db + 64 * ks, idesc, (acc || ks) ? 1u : 0u);
}
// D[64 keys x 32] (+)= A^T, A = [128 q x 64 keys] tile read MN-major (M = keys), B = [128 q x 32] MN-major
__device__ __forceinline__ void mma_ptdo(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool acc) {
  constexpr uint32_t idesc = umma_idesc_bf16_ex(64, 32, 1, 1);
  const uint64_t da = desc_sw128(a_addr), db = desc_sw64(b_addr);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) tc_mma_bf16(tmem_d, da + 128 * ks, db + 64 * ks, idesc, (acc || ks) ? 1u : 0u);
}

// keep decisions of 4 consecutive keys of one query row: x = (hash & 0x7F7F7F7F) + addc has the msb of byte j set iff
// key 4*ctr + j is kept (addc = (128 - th7) * 0x01010101); 8 such x build the 32-key word.
__device__ __forceinline__ uint32_t keep_x(uint32_t key, uint32_t ctr, uint32_t addc) {
  return (drop_hash32(key, ctr) & 0x7F7F7F7Fu) + addc;
}
__device__ __forceinline__ uint32_t hash_word(uint32_t key, uint32_t ctr0, uint32_t addc) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) w = (w >> 1) | (keep_x(key, ctr0 + i, addc) & 0x80808080u);
  return w;
}

// optional clock trace of one softmax warp (block 0, warp 4, lane 0) for tuning: focr_attn_set_trace()
__device__ long long* g_attn_trace = nullptr;
struct Tracer {
  long long* p;
  __device__ __forceinline__ Tracer(int warp, int lane) {
    long long* t = g_attn_trace;
    p = (t != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0) ? t : nullptr;
  }
  __device__ __forceinline__ void mark() {
    if (p) *p++ = clock64();
  }
};

struct WgBars {      // per softmax group
  uint64_t a_full;   // per-tile operand(s) of this group landed (TMA)
  uint64_t a_empty;  // ... and every MMA reading them has completed
  uint64_t a2_full[2];   // dKV kernel: the K_j / V_j double buffer
  uint64_t a2_empty[2];
  uint64_t s_full[2];    // S (and dP) accumulator written
  uint64_t s_empty[2];   // ... and drained by the 4 softmax warps
  uint64_t p_full;       // P / dS tile written to shared memory (4 warps)
  uint64_t p_empty;      // ... and consumed by the second GEMM
  uint64_t o_full;       // output accumulator (O / dQ / dK,dV) complete
  uint64_t o_empty;      // ... and drained
};
struct Bars {
  uint64_t res_full;  // the resident head slices landed
  WgBars wg[4];
  uint32_t tmem_slot;
};

__device__ __forceinline__ void init_bars(Bars* bars) {
  mbar_init(&bars->res_full, 1);
  for (int g = 0; g < 4; ++g) {
    WgBars& w = bars->wg[g];
    mbar_init(&w.a_full, 1);
    mbar_init(&w.a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w.a2_full[i], 1);
      mbar_init(&w.a2_empty[i], 1);
      mbar_init(&w.s_full[i], 1);
      mbar_init(&w.s_empty[i], 4);
    }
    mbar_init(&w.p_full, 4);
    mbar_init(&w.p_empty, 1);
    mbar_init(&w.o_full, 1);
    mbar_init(&w.o_empty, 4);
  }
  fence_mbar_init();
}
// one arrival per softmax warp, after every lane's TMEM loads have completed
__device__ __forceinline__ void warp_release_tmem(uint64_t* bar, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// one arrival per softmax warp, after every lane's shared-memory stores are visible to the tensor-core proxy
__device__ __forceinline__ void warp_publish_smem(uint64_t* bar, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void store_chunks4(uint8_t* tile, int row, int chunk0, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(tile + p_off(row, chunk0 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// ------------------------------------------------------------------------------------------------------------------
// forward.  NWG softmax groups (2, 3 or 4) work on different 128-query tiles of the same (batch, head) at once; warp g
// (lane 0) issues the MMAs and the Q-tile TMA of group g, warps 4.. are the softmax groups.
// shared: K, V resident (2 x 64 KB); per group a Q tile (8 KB) and a P tile (16 KB).
// TMEM per group: S ring (2 x 64 columns; 1 x 64 with four groups) and O (32 columns).
// ------------------------------------------------------------------------------------------------------------------
template <int NWG>
struct FwdCfg {
  static constexpr int kSB = NWG == 4 ? 1 : 2;       // S buffers per group
  static constexpr int kCols = kSB * 64 + 32;        // TMEM columns per group
  static constexpr int kThreads = 128 + NWG * 128;
  static constexpr int kSmem = 2 * kHeadBytes + NWG * (kQTile + kPTile) + 1024 /*barriers*/ + 1024 /*alignment*/;
};

template <int DROP, int NWG>
__global__ void __launch_bounds__(FwdCfg<NWG>::kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mQkv, bf16* __restrict__ out, float* __restrict__ lse2, uint32_t key,
                uint32_t th7, float inv_keep, uint32_t* __restrict__ drop_bits, const uint32_t* __restrict__ seed_dev) {
  using Cfg = FwdCfg<NWG>;
  // device-resident seed (CUDA-graph replay): drop_key(seed, stream) = seed ^ f(stream), `key` then carries f(stream)
  if (DROP != 0 && seed_dev != nullptr) key ^= __ldg(seed_dev);
  constexpr int SB = Cfg::kSB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kHeadBytes;
  uint8_t* sQ = smem + 2 * kHeadBytes;             // [NWG]
  uint8_t* sP = sQ + NWG * kQTile;                 // [NWG]
  Bars* bars = reinterpret_cast<Bars*>(sP + NWG * kPTile);
  const int bh = blockIdx.x, b = bh >> 2, h = bh & 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mQkv);
    init_bars(bars);
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp < NWG) {
    if (lane == 0) {
      const int g = warp;
      WgBars& w = bars->wg[g];
      if (g == 0) {
        mbar_arrive_expect_tx(&bars->res_full, 2 * kHeadBytes);
        for (int i = 0; i < 8; ++i) {
          tma_load_2d(sK + i * kQTile, &mQkv, &bars->res_full, 128 + h * 32, b * kS + i * 128);
          tma_load_2d(sV + i * kQTile, &mQkv, &bars->res_full, 256 + h * 32, b * kS + i * 128);
        }
      }
      const uint32_t tS = tmem + g * Cfg::kCols, tO = tS + SB * 64;
      const uint32_t aQ = smem_u32(sQ + g * kQTile), aP = smem_u32(sP + g * kPTile);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      uint32_t ns = 0, np = 0;
      for (int it = 0; g + NWG * it < 8; ++it) {
        mbar_wait_parked(&w.a_empty, (it & 1) ^ 1);  // every S MMA of the previous tile has read the Q buffer
        mbar_arrive_expect_tx(&w.a_full, kQTile);
        tma_load_2d(sQ + g * kQTile, &mQkv, &w.a_full, h * 32, b * kS + (g + NWG * it) * 128);
        if (it == 0) mbar_wait_parked(&bars->res_full, 0);
        mbar_wait_parked(&w.a_full, it & 1);
        tc_fence_after();
        for (int j = 0; j < 32; ++j) {  // 16 key tiles for the max pass, 16 for the exp / PV pass
          const int sb = ns % SB;
          mbar_wait_parked(&w.s_empty[sb], ((ns / SB) & 1) ^ 1);
          tc_fence_after();
          mma_qk(tS + sb * 64, aQ, aK + (j & 15) * 4096);
          tc_commit(&w.s_full[sb]);
          ++ns;
          if (j == 31) tc_commit(&w.a_empty);
          if (j >= 17) {  // P V of key tile j - 17
            const int jj = j - 17;
            mbar_wait_parked(&w.p_full, np & 1);
            tc_fence_after();
            if (jj == 0) {
              mbar_wait_parked(&w.o_empty, (it & 1) ^ 1);
              tc_fence_after();
            }
            mma_pv(tO, aP, aV + jj * 4096, jj != 0);
            tc_commit(&w.p_empty);
            ++np;
          }
        }
        mbar_wait_parked(&w.p_full, np & 1);
        tc_fence_after();
        mma_pv(tO, aP, aV + 15 * 4096, true);
        tc_commit(&w.p_empty);
        ++np;
        tc_commit(&w.o_full);
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2, quad = warp & 3, row = quad * 32 + lane;
    WgBars& w = bars->wg[g];
    const uint32_t tS = tmem + ((uint32_t)(quad * 32) << 16) + g * Cfg::kCols, tO = tS + SB * 64;
    uint8_t* myP = sP + g * kPTile;
    const uint32_t addc = (128u - th7) * 0x01010101u;
    uint32_t ns = 0, np = 0;
    Tracer tr(warp, lane);
    for (int it = 0; g + NWG * it < 8; ++it) {
      const int q = (g + NWG * it) * 128 + row;
      // ---- pass 1: exact row maximum ----
      float m = -INFINITY;
      for (int j = 0; j < 16; ++j) {
        const int sb = ns % SB;
        tr.mark();
        mbar_wait(&w.s_full[sb], (ns / SB) & 1);
        tr.mark();
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tS + sb * 64 + hf * 32, r);
          tmem_ld_wait();
          if (hf == 1) warp_release_tmem(&w.s_empty[sb], lane);
          float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]);
#pragma unroll
          for (int e = 2; e < 32; e += 2) {
            m0 = fmaxf(m0, __uint_as_float(r[e]));
            m1 = fmaxf(m1, __uint_as_float(r[e + 1]));
          }
          m = fmaxf(m, fmaxf(m0, m1));
        }
        ++ns;
      }
      const float mneg = m * kScaleLog2;
      // ---- pass 2: P = exp2(S c - m c), row sum, dropout, P -> shared ----
      float l0 = 0.f, l1 = 0.f;
      const uint32_t rowctr = (uint32_t)(bh * kS + q) * 256u;
      for (int j = 0; j < 16; ++j) {
        const int sb = ns % SB;
        tr.mark();
        mbar_wait(&w.s_full[sb], (ns / SB) & 1);
        tr.mark();
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tS + sb * 64 + hf * 32, r);
          tmem_ld_wait();
          if (hf == 1) warp_release_tmem(&w.s_empty[sb], lane);
          tr.mark();
          uint32_t pk[16];
          uint32_t word = 0;
          const uint32_t ctr0 = rowctr + (uint32_t)(j * 16 + hf * 8);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) p[e] = ex2(fmaf(__uint_as_float(r[4 * i + e]), kScaleLog2, -mneg));
            l0 += p[0] + p[2];
            l1 += p[1] + p[3];
            pk[2 * i] = pack_bf16x2(p[0], p[1]);
            pk[2 * i + 1] = pack_bf16x2(p[2], p[3]);
            if (DROP) {
              const uint32_t x = keep_x(key, ctr0 + i, addc);
              pk[2 * i] &= prmt(x, 0x9988u);
              pk[2 * i + 1] &= prmt(x, 0xBBAAu);
              if (DROP == 2) word = (word >> 1) | (x & 0x80808080u);
            }
          }
          tr.mark();
          if (hf == 0) mbar_wait(&w.p_empty, (np & 1) ^ 1);  // the P V of the previous tile has read the buffer
          tr.mark();
          store_chunks4(myP, row, hf * 4, pk);
          if (DROP == 2) drop_bits[((size_t)bh * 32 + j * 2 + hf) * kS + q] = word;
        }
        tr.mark();
        warp_publish_smem(&w.p_full, lane);
        tr.mark();
        ++np;
        ++ns;
      }
      // ---- epilogue: O / l ----
      tr.mark();
      mbar_wait(&w.o_full, it & 1);
      tr.mark();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tO, r);
      tmem_ld_wait();
      warp_release_tmem(&w.o_empty, lane);
      const float l = l0 + l1;
      const float sc = inv_keep / l;
      uint4* orow = reinterpret_cast<uint4*>(out + ((long)b * kS + q) * kLdO + h * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]) * sc, __uint_as_float(r[8 * c + 1]) * sc);
        o.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * sc, __uint_as_float(r[8 * c + 3]) * sc);
        o.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * sc, __uint_as_float(r[8 * c + 5]) * sc);
        o.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * sc, __uint_as_float(r[8 * c + 7]) * sc);
        orow[c] = o;
      }
      lse2[(long)bh * kS + q] = mneg + log2f(l);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward, shared by both passes: one 32-key half of a 128 x 64 tile.
//   P = exp2(S c - L),  dS = P o (keep o dP - D'),  Pd = keep o P   (bf16 pairs)
// ------------------------------------------------------------------------------------------------------------------
template <int DROP, bool WITH_P>
__device__ __forceinline__ void bwd_half(const uint32_t (&rs)[32], const uint32_t (&rd)[32], float L, float Dp,
                                         uint32_t word, uint32_t (&ds)[16], uint32_t (&pd)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float p[4], e[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      p[c] = ex2(fmaf(__uint_as_float(rs[4 * i + c]), kScaleLog2, -L));
      e[c] = __uint_as_float(rd[4 * i + c]);
    }
    uint32_t pp0 = 0, pp1 = 0;
    if (WITH_P) {
      pp0 = pack_bf16x2(p[0], p[1]);
      pp1 = pack_bf16x2(p[2], p[3]);
    }
    if (DROP) {
      const uint32_t x = word << (7 - i);  // bit 8 c + i -> msb of byte c
      e[0] = __uint_as_float(rd[4 * i + 0] & prmt(x, 0x8888u));
      e[1] = __uint_as_float(rd[4 * i + 1] & prmt(x, 0x9999u));
      e[2] = __uint_as_float(rd[4 * i + 2] & prmt(x, 0xAAAAu));
      e[3] = __uint_as_float(rd[4 * i + 3] & prmt(x, 0xBBBBu));
      if (WITH_P) {
        pp0 &= prmt(x, 0x9988u);
        pp1 &= prmt(x, 0xBBAAu);
      }
    }
    ds[2 * i] = pack_bf16x2(p[0] * (e[0] - Dp), p[1] * (e[1] - Dp));
    ds[2 * i + 1] = pack_bf16x2(p[2] * (e[2] - Dp), p[3] * (e[3] - Dp));
    if (WITH_P) {
      pd[2 * i] = pp0;
      pd[2 * i + 1] = pp1;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward pass A: dQ (and D = rowsum(dO o O), written for pass B).
// shared: K, V resident; per group a Q tile, a dO tile and a dS tile.  TMEM per group: S 64, dP 64, dQ 32 columns.
// ------------------------------------------------------------------------------------------------------------------
template <int NWG>
struct DqCfg {
  static constexpr int kThreads = 128 + NWG * 128;
  static constexpr int kSmem = 2 * kHeadBytes + NWG * (2 * kQTile + kPTile) + 1024 + 1024;
};

template <int DROP, int NWG>
__global__ void __launch_bounds__(DqCfg<NWG>::kThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mQkv, const __grid_constant__ CUtensorMap mDo,
                   const bf16* __restrict__ o_in, const bf16* __restrict__ d_o, const float* __restrict__ lse2,
                   float* __restrict__ dsum, bf16* __restrict__ dqkv, uint32_t key, uint32_t th7, float inv_keep,
                   const uint32_t* __restrict__ drop_bits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kHeadBytes;
  uint8_t* sQ = smem + 2 * kHeadBytes;   // [NWG]
  uint8_t* sDo = sQ + NWG * kQTile;      // [NWG]
  uint8_t* sDs = sDo + NWG * kQTile;     // [NWG]
  Bars* bars = reinterpret_cast<Bars*>(sDs + NWG * kPTile);
  const int bh = blockIdx.x, b = bh >> 2, h = bh & 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mQkv);
    tma_prefetch_desc(&mDo);
    init_bars(bars);
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp < NWG) {
    if (lane == 0) {
      const int g = warp;
      WgBars& w = bars->wg[g];
      if (g == 0) {
        mbar_arrive_expect_tx(&bars->res_full, 2 * kHeadBytes);
        for (int i = 0; i < 8; ++i) {
          tma_load_2d(sK + i * kQTile, &mQkv, &bars->res_full, 128 + h * 32, b * kS + i * 128);
          tma_load_2d(sV + i * kQTile, &mQkv, &bars->res_full, 256 + h * 32, b * kS + i * 128);
        }
      }
      const uint32_t tS = tmem + g * 160, tDp = tS + 64, tDq = tS + 128;
      const uint32_t aQ = smem_u32(sQ + g * kQTile), aDo = smem_u32(sDo + g * kQTile), aDs = smem_u32(sDs + g * kPTile);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      uint32_t n = 0;
      for (int it = 0; g + NWG * it < 8; ++it) {
        mbar_wait_parked(&w.a_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(&w.a_full, 2 * kQTile);
        tma_load_2d(sQ + g * kQTile, &mQkv, &w.a_full, h * 32, b * kS + (g + NWG * it) * 128);
        tma_load_2d(sDo + g * kQTile, &mDo, &w.a_full, h * 32, b * kS + (g + NWG * it) * 128);
        if (it == 0) mbar_wait_parked(&bars->res_full, 0);
        mbar_wait_parked(&w.a_full, it & 1);
        tc_fence_after();
        for (int j = 0; j <= 16; ++j) {
          if (j < 16) {
            mbar_wait_parked(&w.s_empty[0], (n & 1) ^ 1);
            tc_fence_after();
            mma_qk(tS, aQ, aK + j * 4096);
            mma_qk(tDp, aDo, aV + j * 4096);
            tc_commit(&w.s_full[0]);
            if (j == 15) tc_commit(&w.a_empty);
            ++n;
          }
          if (j >= 1) {  // dQ += dS K of key tile j - 1  (its tile counter is n - 2 for j < 16, n - 1 for j == 16)
            const int jj = j - 1;
            const uint32_t nn = (j < 16) ? n - 2 : n - 1;
            mbar_wait_parked(&w.p_full, nn & 1);
            tc_fence_after();
            if (jj == 0) {
              mbar_wait_parked(&w.o_empty, (it & 1) ^ 1);
              tc_fence_after();
            }
            mma_pv(tDq, aDs, aK + jj * 4096, jj != 0);
            tc_commit(&w.p_empty);
          }
        }
        tc_commit(&w.o_full);
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2, quad = warp & 3, row = quad * 32 + lane;
    WgBars& w = bars->wg[g];
    const uint32_t tS = tmem + ((uint32_t)(quad * 32) << 16) + g * 160, tDp = tS + 64, tDq = tS + 128;
    uint8_t* myDs = sDs + g * kPTile;
    const uint32_t addc = (128u - th7) * 0x01010101u;
    const float keep_prob = 1.f / inv_keep;
    uint32_t n = 0;
    for (int it = 0; g + NWG * it < 8; ++it) {
      const int q = (g + NWG * it) * 128 + row;
      const long t = (long)b * kS + q;
      const float L = lse2[(long)bh * kS + q];
      float D = 0.f;
      {
        const uint4* po = reinterpret_cast<const uint4*>(o_in + t * kLdO + h * 32);
        const uint4* pd = reinterpret_cast<const uint4*>(d_o + t * kLdO + h * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 a = po[c], d = pd[c];
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
            D = fmaf(x.x, y.x, D);
            D = fmaf(x.y, y.y, D);
          }
        }
      }
      dsum[(long)bh * kS + q] = D;
      const float Dp = D * keep_prob;
      const uint32_t rowctr = (uint32_t)(bh * kS + q) * 256u;
      uint32_t wn[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};  // keep-bit words, loaded one key tile ahead
      if (DROP == 2) {
        wn[0] = __ldg(drop_bits + ((size_t)bh * 32) * kS + q);
        wn[1] = __ldg(drop_bits + ((size_t)bh * 32 + 1) * kS + q);
      }
      for (int j = 0; j < 16; ++j) {
        uint32_t wd[2] = {wn[0], wn[1]};
        if (DROP == 2 && j < 15) {
          wn[0] = __ldg(drop_bits + ((size_t)bh * 32 + j * 2 + 2) * kS + q);
          wn[1] = __ldg(drop_bits + ((size_t)bh * 32 + j * 2 + 3) * kS + q);
        }
        mbar_wait(&w.s_full[0], n & 1);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t rs[32], rd[32];
          tmem_ld_32x32b_x32(tS + hf * 32, rs);
          tmem_ld_32x32b_x32(tDp + hf * 32, rd);
          tmem_ld_wait();
          if (hf == 1) warp_release_tmem(&w.s_empty[0], lane);
          if (DROP == 1) wd[hf] = hash_word(key, rowctr + (uint32_t)(j * 16 + hf * 8), addc);
          uint32_t ds[16], pd[16];
          bwd_half<DROP, false>(rs, rd, L, Dp, wd[hf], ds, pd);
          if (hf == 0) mbar_wait(&w.p_empty, (n & 1) ^ 1);
          store_chunks4(myDs, row, hf * 4, ds);
        }
        warp_publish_smem(&w.p_full, lane);
        ++n;
      }
      mbar_wait(&w.o_full, it & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tDq, r);
      tmem_ld_wait();
      warp_release_tmem(&w.o_empty, lane);
      const float sc = kScale * inv_keep;
      uint4* drow = reinterpret_cast<uint4*>(dqkv + t * kLdQkv + h * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]) * sc, __uint_as_float(r[8 * c + 1]) * sc);
        o.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * sc, __uint_as_float(r[8 * c + 3]) * sc);
        o.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * sc, __uint_as_float(r[8 * c + 5]) * sc);
        o.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * sc, __uint_as_float(r[8 * c + 7]) * sc);
        drow[c] = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward pass B: dK, dV.  shared: Q, dO resident; per group a double-buffered (K_j, V_j) 64-key block, a P tile and
// a dS tile.  TMEM per group: S 64, dP 64, dK 32, dV 32 columns (the two M = 64 accumulators use lanes
// (r % 16) + 32 (r / 16)).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kKvBlk = 64 * 64;  // [64 keys][32] bf16
constexpr int kDkvSmem = 2 * kHeadBytes + 2 * (4 * kKvBlk + 2 * kPTile) + 1024 + 1024;

template <int DROP>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap mQkv, const __grid_constant__ CUtensorMap mQkv64,
                    const __grid_constant__ CUtensorMap mDo, const float* __restrict__ lse2,
                    const float* __restrict__ dsum, bf16* __restrict__ dqkv, uint32_t key, uint32_t th7, float inv_keep,
                    const uint32_t* __restrict__ drop_bits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sDo = smem + kHeadBytes;
  uint8_t* sKv = smem + 2 * kHeadBytes;   // [2 groups][2 buffers][K_j | V_j]
  uint8_t* sP = sKv + 2 * 4 * kKvBlk;     // [2 groups]
  uint8_t* sDs = sP + 2 * kPTile;         // [2 groups]
  Bars* bars = reinterpret_cast<Bars*>(sDs + 2 * kPTile);
  const int bh = blockIdx.x, b = bh >> 2, h = bh & 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mQkv);
    tma_prefetch_desc(&mQkv64);
    tma_prefetch_desc(&mDo);
  }
  if (warp == 1 && lane == 0) init_bars(bars);
  if (warp == 3) {
    tmem_alloc(&bars->tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->res_full, 2 * kHeadBytes);
      for (int i = 0; i < 8; ++i) {
        tma_load_2d(sQ + i * kQTile, &mQkv, &bars->res_full, h * 32, b * kS + i * 128);
        tma_load_2d(sDo + i * kQTile, &mDo, &bars->res_full, h * 32, b * kS + i * 128);
      }
      for (int jt = 0; jt < 8; ++jt)
        for (int g = 0; g < 2; ++g) {
          WgBars& w = bars->wg[g];
          const int kb = jt & 1, j = g + 2 * jt;
          mbar_wait_parked(&w.a2_empty[kb], ((jt >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&w.a2_full[kb], 2 * kKvBlk);
          uint8_t* dst = sKv + (g * 2 + kb) * 2 * kKvBlk;
          tma_load_2d(dst, &mQkv64, &w.a2_full[kb], 128 + h * 32, b * kS + j * 64);
          tma_load_2d(dst + kKvBlk, &mQkv64, &w.a2_full[kb], 256 + h * 32, b * kS + j * 64);
        }
    }
  } else if (warp == 1 || warp == 2) {
    if (lane == 0) {
      const int g = warp - 1;
      WgBars& w = bars->wg[g];
      const uint32_t tS = tmem + g * 192, tDp = tS + 64, tDk = tS + 128, tDv = tS + 160;
      const uint32_t aQ = smem_u32(sQ), aDo = smem_u32(sDo);
      const uint32_t aP = smem_u32(sP + g * kPTile), aDs = smem_u32(sDs + g * kPTile);
      mbar_wait_parked(&bars->res_full, 0);
      uint32_t n = 0;
      for (int jt = 0; jt < 8; ++jt) {
        const int kb = jt & 1;
        const uint32_t aKj = smem_u32(sKv + (g * 2 + kb) * 2 * kKvBlk), aVj = aKj + kKvBlk;
        mbar_wait_parked(&w.a2_full[kb], (jt >> 1) & 1);
        tc_fence_after();
        for (int i = 0; i <= 8; ++i) {
          if (i < 8) {
            mbar_wait_parked(&w.s_empty[0], (n & 1) ^ 1);
            tc_fence_after();
            mma_qk(tS, aQ + i * kQTile, aKj);
            mma_qk(tDp, aDo + i * kQTile, aVj);
            tc_commit(&w.s_full[0]);
            if (i == 7) tc_commit(&w.a2_empty[kb]);
            ++n;
          }
          if (i >= 1) {  // dV += Pd^T dO, dK += dS^T Q of query tile i - 1
            const int ii = i - 1;
            const uint32_t nn = (i < 8) ? n - 2 : n - 1;
            mbar_wait_parked(&w.p_full, nn & 1);
            tc_fence_after();
            if (ii == 0) {
              mbar_wait_parked(&w.o_empty, (jt & 1) ^ 1);
              tc_fence_after();
            }
            mma_ptdo(tDv, aP, aDo + ii * kQTile, ii != 0);
            mma_ptdo(tDk, aDs, aQ + ii * kQTile, ii != 0);
            tc_commit(&w.p_empty);
          }
        }
        tc_commit(&w.o_full);
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2, quad = warp & 3, row = quad * 32 + lane;
    WgBars& w = bars->wg[g];
    const uint32_t tS = tmem + ((uint32_t)(quad * 32) << 16) + g * 192, tDp = tS + 64, tDk = tS + 128, tDv = tS + 160;
    uint8_t* myP = sP + g * kPTile;
    uint8_t* myDs = sDs + g * kPTile;
    const uint32_t addc = (128u - th7) * 0x01010101u;
    const float keep_prob = 1.f / inv_keep;
    uint32_t n = 0;
    // per-tile row scalars and keep-bit words are loaded one tile ahead of their use
    float Ln = lse2[(long)bh * kS + row], Dn = dsum[(long)bh * kS + row];
    uint32_t wn[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    if (DROP == 2) {
      wn[0] = __ldg(drop_bits + ((size_t)bh * 32 + g * 2) * kS + row);
      wn[1] = __ldg(drop_bits + ((size_t)bh * 32 + g * 2 + 1) * kS + row);
    }
    for (int jt = 0; jt < 8; ++jt) {
      const int j = g + 2 * jt;  // 64-key block
      for (int i = 0; i < 8; ++i) {
        const int q = i * 128 + row;
        const float L = Ln;
        const float Dp = Dn * keep_prob;
        uint32_t wd[2] = {wn[0], wn[1]};
        if (i < 7 || jt < 7) {
          const int qn = ((i + 1) & 7) * 128 + row, jn = (i < 7) ? j : j + 2;
          Ln = lse2[(long)bh * kS + qn];
          Dn = dsum[(long)bh * kS + qn];
          if (DROP == 2) {
            wn[0] = __ldg(drop_bits + ((size_t)bh * 32 + jn * 2) * kS + qn);
            wn[1] = __ldg(drop_bits + ((size_t)bh * 32 + jn * 2 + 1) * kS + qn);
          }
        }
        const uint32_t rowctr = (uint32_t)(bh * kS + q) * 256u;
        mbar_wait(&w.s_full[0], n & 1);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t rs[32], rd[32];
          tmem_ld_32x32b_x32(tS + hf * 32, rs);
          tmem_ld_32x32b_x32(tDp + hf * 32, rd);
          tmem_ld_wait();
          if (hf == 1) warp_release_tmem(&w.s_empty[0], lane);
          if (DROP == 1) wd[hf] = hash_word(key, rowctr + (uint32_t)(j * 16 + hf * 8), addc);
          uint32_t ds[16], pd[16];
          bwd_half<DROP, true>(rs, rd, L, Dp, wd[hf], ds, pd);
          if (hf == 0) mbar_wait(&w.p_empty, (n & 1) ^ 1);
          store_chunks4(myDs, row, hf * 4, ds);
          store_chunks4(myP, row, hf * 4, pd);
        }
        warp_publish_smem(&w.p_full, lane);
        ++n;
      }
      mbar_wait(&w.o_full, jt & 1);
      tc_fence_after();
      uint32_t rk[32], rv[32];
      tmem_ld_32x32b_x32(tDk, rk);
      tmem_ld_32x32b_x32(tDv, rv);
      tmem_ld_wait();
      warp_release_tmem(&w.o_empty, lane);
      if (lane < 16) {  // M = 64 accumulator: key row quad*16 + lane lives on TMEM lane quad*32 + lane
        const long t = (long)b * kS + j * 64 + quad * 16 + lane;
        const float ksc = kScale * inv_keep;
        uint4* krow = reinterpret_cast<uint4*>(dqkv + t * kLdQkv + 128 + h * 32);
        uint4* vrow = reinterpret_cast<uint4*>(dqkv + t * kLdQkv + 256 + h * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(rk[8 * c + 0]) * ksc, __uint_as_float(rk[8 * c + 1]) * ksc);
          o.y = pack_bf16x2(__uint_as_float(rk[8 * c + 2]) * ksc, __uint_as_float(rk[8 * c + 3]) * ksc);
          o.z = pack_bf16x2(__uint_as_float(rk[8 * c + 4]) * ksc, __uint_as_float(rk[8 * c + 5]) * ksc);
          o.w = pack_bf16x2(__uint_as_float(rk[8 * c + 6]) * ksc, __uint_as_float(rk[8 * c + 7]) * ksc);
          krow[c] = o;
          o.x = pack_bf16x2(__uint_as_float(rv[8 * c + 0]) * inv_keep, __uint_as_float(rv[8 * c + 1]) * inv_keep);
          o.y = pack_bf16x2(__uint_as_float(rv[8 * c + 2]) * inv_keep, __uint_as_float(rv[8 * c + 3]) * inv_keep);
          o.z = pack_bf16x2(__uint_as_float(rv[8 * c + 4]) * inv_keep, __uint_as_float(rv[8 * c + 5]) * inv_keep);
          o.w = pack_bf16x2(__uint_as_float(rv[8 * c + 6]) * inv_keep, __uint_as_float(rv[8 * c + 7]) * inv_keep);
          vrow[c] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <typename K>
int set_smem(K kernel, int bytes) {
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return FOCR_OK;
}

struct AttnDrop {
  uint32_t th7;
  float inv_keep;
};
AttnDrop drop_params(uint32_t thresh16) {
  AttnDrop d;
  d.th7 = (thresh16 + 256) >> 9;  // p * 128, rounded
  if (thresh16 && d.th7 == 0) d.th7 = 1;
  if (d.th7 > 127) d.th7 = 127;
  d.inv_keep = 128.f / (128.f - (float)d.th7);
  return d;
}

}  // namespace

// p_drop = thresh16 / 65536; thresh16 == 0 disables dropout (eval / parity runs).  drop_bits: optional keep-bit
// buffer of attn_drop_bits_bytes(B) bytes written by the forward and consumed by the backward (may be null: the
// backward then regenerates the mask from the seed).
// tuning aid: device buffer (>= 64 KB of int64) that block 0 / warp 4 fills with clock64() marks, or null to disable
extern "C" int focr_attn_set_trace(void* buf) {
  long long* p = (long long*)buf;
  FOCR_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_trace, &p, sizeof(p)));
  return FOCR_OK;
}

size_t attn_drop_bits_bytes(int B) { return (size_t)B * 4 * (kS / 32) * kS * sizeof(uint32_t); }

template <int NWG>
int launch_fwd(const CUtensorMap& mq, bf16* out, float* lse2, int B, uint32_t key, uint32_t thresh16, uint32_t* drop_bits,
               const uint32_t* seed_dev, cudaStream_t s) {
  using Cfg = FwdCfg<NWG>;
  static bool init = false;
  if (!init) {
    int rc = set_smem(attn_fwd_kernel<0, NWG>, Cfg::kSmem);
    if (rc) return rc;
    rc = set_smem(attn_fwd_kernel<1, NWG>, Cfg::kSmem);
    if (rc) return rc;
    rc = set_smem(attn_fwd_kernel<2, NWG>, Cfg::kSmem);
    if (rc) return rc;
    init = true;
  }
  const AttnDrop d = drop_params(thresh16);
  const dim3 grid(B * 4), block(Cfg::kThreads);
  if (!thresh16)
    attn_fwd_kernel<0, NWG><<<grid, block, Cfg::kSmem, s>>>(mq, out, lse2, key, 0, 1.f, nullptr, nullptr);
  else if (!drop_bits)
    attn_fwd_kernel<1, NWG><<<grid, block, Cfg::kSmem, s>>>(mq, out, lse2, key, d.th7, d.inv_keep, nullptr, seed_dev);
  else
    attn_fwd_kernel<2, NWG><<<grid, block, Cfg::kSmem, s>>>(mq, out, lse2, key, d.th7, d.inv_keep, drop_bits, seed_dev);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int attn_forward(const bf16* qkv, bf16* out, float* lse2, int B, uint32_t key, uint32_t thresh16, uint32_t* drop_bits,
                 cudaStream_t s, const uint32_t* seed_dev) {
  ProfScope _ps("attn_fwd", s);
  FOCR_REQUIRE(B >= 1 && B <= 1024, "attention: B=%d out of range", B);
  static_assert(sizeof(Bars) <= 1024, "barrier block");
  CUtensorMap mq;
  int rc = focr_make_tmap_2d(&mq, qkv, kLdQkv, (unsigned long long)B * kS, kLdQkv * 2, 32, 128, 64);
  if (rc) return rc;
  static int nwg = 0;
  if (!nwg) {
    const char* e = getenv("FOCR_ATTN_FWD_NWG");  // tuning knob: softmax groups per CTA (2, 3 or 4)
    nwg = e ? atoi(e) : 4;
    if (nwg < 2 || nwg > 4) nwg = 4;
  }
  if (nwg == 2) return launch_fwd<2>(mq, out, lse2, B, key, thresh16, drop_bits, seed_dev, s);
  if (nwg == 3) return launch_fwd<3>(mq, out, lse2, B, key, thresh16, drop_bits, seed_dev, s);
  return launch_fwd<4>(mq, out, lse2, B, key, thresh16, drop_bits, seed_dev, s);
}

template <int NWG>
int launch_dq(const CUtensorMap& mq, const CUtensorMap& mdo, const bf16* o, const bf16* d_o, const float* lse2, float* dsum,
              bf16* dqkv, int B, uint32_t key, uint32_t thresh16, const uint32_t* drop_bits, cudaStream_t s) {
  using Cfg = DqCfg<NWG>;
  static bool init = false;
  if (!init) {
    int rc = set_smem(attn_bwd_dq_kernel<0, NWG>, Cfg::kSmem);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dq_kernel<1, NWG>, Cfg::kSmem);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dq_kernel<2, NWG>, Cfg::kSmem);
    if (rc) return rc;
    init = true;
  }
  const AttnDrop d = drop_params(thresh16);
  const dim3 grid(B * 4), block(Cfg::kThreads);
  if (!thresh16)
    attn_bwd_dq_kernel<0, NWG><<<grid, block, Cfg::kSmem, s>>>(mq, mdo, o, d_o, lse2, dsum, dqkv, key, 0, 1.f, nullptr);
  else if (!drop_bits)
    attn_bwd_dq_kernel<1, NWG><<<grid, block, Cfg::kSmem, s>>>(mq, mdo, o, d_o, lse2, dsum, dqkv, key, d.th7, d.inv_keep,
                                                                nullptr);
  else
    attn_bwd_dq_kernel<2, NWG><<<grid, block, Cfg::kSmem, s>>>(mq, mdo, o, d_o, lse2, dsum, dqkv, key, d.th7, d.inv_keep,
                                                                drop_bits);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}

int attn_backward(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse2, float* dsum, bf16* dqkv, int B,
                  uint32_t key, uint32_t thresh16, const uint32_t* drop_bits, cudaStream_t s) {
  FOCR_REQUIRE(B >= 1 && B <= 1024, "attention: B=%d out of range", B);
  static bool init = false;
  if (!init) {
    int rc = set_smem(attn_bwd_dkv_kernel<0>, kDkvSmem);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dkv_kernel<1>, kDkvSmem);
    if (rc) return rc;
    rc = set_smem(attn_bwd_dkv_kernel<2>, kDkvSmem);
    if (rc) return rc;
    init = true;
  }
  CUtensorMap mq, mq64, mdo;
  int rc = focr_make_tmap_2d(&mq, qkv, kLdQkv, (unsigned long long)B * kS, kLdQkv * 2, 32, 128, 64);
  if (rc) return rc;
  rc = focr_make_tmap_2d(&mq64, qkv, kLdQkv, (unsigned long long)B * kS, kLdQkv * 2, 32, 64, 64);
  if (rc) return rc;
  rc = focr_make_tmap_2d(&mdo, d_o, kLdO, (unsigned long long)B * kS, kLdO * 2, 32, 128, 64);
  if (rc) return rc;
  const AttnDrop d = drop_params(thresh16);
  const dim3 grid(B * 4), block(kThreads);
  {
    ProfScope ps("attn_bwd_dq", s);
    static int nwg = 0;
    if (!nwg) {
      const char* e = getenv("FOCR_ATTN_DQ_NWG");  // tuning knob: softmax groups per CTA (2 or 3)
      nwg = e ? atoi(e) : 2;  // measured: a third group does not pay (0.63 vs 0.60 ms at B = 256)
      if (nwg < 2 || nwg > 3) nwg = 2;
    }
    rc = nwg == 2 ? launch_dq<2>(mq, mdo, o, d_o, lse2, dsum, dqkv, B, key, thresh16, drop_bits, s)
                  : launch_dq<3>(mq, mdo, o, d_o, lse2, dsum, dqkv, B, key, thresh16, drop_bits, s);
    if (rc) return rc;
  }
  {
    ProfScope ps("attn_bwd_dkv", s);
    if (!thresh16)
      attn_bwd_dkv_kernel<0><<<grid, block, kDkvSmem, s>>>(mq, mq64, mdo, lse2, dsum, dqkv, key, 0, 1.f, nullptr);
    else if (!drop_bits)
      attn_bwd_dkv_kernel<1><<<grid, block, kDkvSmem, s>>>(mq, mq64, mdo, lse2, dsum, dqkv, key, d.th7, d.inv_keep,
                                                           nullptr);
    else
      attn_bwd_dkv_kernel<2><<<grid, block, kDkvSmem, s>>>(mq, mq64, mdo, lse2, dsum, dqkv, key, d.th7, d.inv_keep,
                                                           drop_bits);
    FOCR_LAUNCH_CHECK();
  }
  return FOCR_OK;
}
