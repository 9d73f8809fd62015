// Layout probe for the tcgen05 operand descriptors the attention kernels rely on (K-major / MN-major operands,
// 64- and 128-byte swizzles, M = 64 accumulators).  The caller supplies the exact shared-memory image and the two
// 64-bit operand descriptors (start-address field relative to the image); the kernel issues `nk` MMAs, advancing each
// descriptor by a fixed number of 16-byte units per step, and dumps the TMEM accumulator lane by lane.
// tests/test_gpu_umma_layouts.py builds the images in numpy and checks the result against A.B^T - this pins the
// descriptor semantics independently of any pipeline code.
#include "kernels.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ img, int img_bytes, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                  int nk, uint32_t a_step16, uint32_t b_step16, float* __restrict__ out, int ncols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid * 16; i < img_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(img + i);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint64_t base16 = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
    for (int k = 0; k < nk; ++k)
      tc_mma_bf16(tmem, desc_a + base16 + (uint64_t)k * a_step16, desc_b + base16 + (uint64_t)k * b_step16, idesc,
                  k != 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(long)(warp * 32 + lane) * ncols + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

extern "C" int focr_umma_probe(const void* img, int img_bytes, unsigned long long desc_a, unsigned long long desc_b,
                               unsigned idesc, int nk, unsigned a_step16, unsigned b_step16, float* out, int ncols,
                               void* stream) {
  FOCR_REQUIRE(img_bytes > 0 && img_bytes % 16 == 0 && img_bytes <= 200 * 1024, "umma_probe: image of %d bytes", img_bytes);
  FOCR_REQUIRE(ncols % 32 == 0 && ncols >= 32 && ncols <= 256, "umma_probe: ncols %d", ncols);
  const int smem = img_bytes + 1024;
  FOCR_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)img, img_bytes, desc_a, desc_b, idesc, nk,
                                                              a_step16, b_step16, out, ncols);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
