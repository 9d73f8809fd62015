// Input pipeline, device side: the bicubic resize + ToTensor the reference's collate functions apply to every crop on
// the CPU through Pillow - `img.resize((W, H), Image.BICUBIC)` and `transforms.ToTensor()` in resizeNormalize
// (scene-text-telescope/dataset/dataset.py:136-152), called per image by alignCollate_real (:258-270) for the HR (128 x 32)
// and LR (64 x 16) batches.  Bit-exact with Pillow's 8-bit resampler (src/libImaging/Resample.c): per axis a window of
// bicubic (a = -0.5) weights with antialiasing support = 2 * max(in/out, 1), normalised in double precision, quantised to
// 22-bit fixed point; horizontal pass into a uint8 intermediate, vertical pass, clip8((2^21 + sum k p) >> 22); then /255.
// One CTA per crop: the ragged uint8 batch arrives as one packed buffer + (offset, h, w) per crop, coefficients and the
// intermediate live in shared memory, the output is the dense fp32 NCHW batch the networks consume.
#include "kernels.cuh"

namespace {

constexpr int kPrec = 32 - 8 - 2;

__device__ __forceinline__ double bicubic_w(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0)  // ((a + 2) x - (a + 3)) x x + 1, evaluated left to right without contraction
    return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0)  // (((x - 5) x + 8) x - 4) a
    return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output position xx of an axis in_size -> out_size (full box)
__device__ void axis_coeffs(int in_size, int out_size, int ksize, int xx, int* __restrict__ kk, int* __restrict__ bounds) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale > 1.0 ? scale : 1.0;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn((double)xx + 0.5, scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x)
    ww = __dadd_rn(ww, bicubic_w(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
  for (int x = 0; x < ksize; ++x) {
    double k = 0.0;
    if (x < xmax) {
      k = bicubic_w(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
      if (ww != 0.0) k = __ddiv_rn(k, ww);
    }
    const double v = __dmul_rn(k, (double)(1 << kPrec));
    kk[x] = k < 0.0 ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
  }
  bounds[0] = xmin;
  bounds[1] = xmax;
}

__device__ __forceinline__ int axis_ksize(int in_size, int out_size) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale > 1.0 ? scale : 1.0;
  return (int)ceil(__dmul_rn(2.0, filterscale)) * 2 + 1;
}
__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrec;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// meta: int64 [B][3] = {byte offset of the crop in `pixels`, h, w}; crop = h x w x 3 uint8, rows contiguous
__global__ void __launch_bounds__(256) resize_normalize_kernel(const unsigned char* __restrict__ pixels,
                                                               const long long* __restrict__ meta, int ow, int oh,
                                                               int ks_cap, int h_cap, float* __restrict__ out,
                                                               int* __restrict__ status) {
  extern __shared__ int sm_i[];
  int* kh = sm_i;                          // [ow][ks_cap]
  int* kv = kh + ow * ks_cap;              // [oh][ks_cap]
  int* bh = kv + oh * ks_cap;              // [ow][2]
  int* bv = bh + 2 * ow;                   // [oh][2]
  unsigned char* tmp = reinterpret_cast<unsigned char*>(bv + 2 * oh);  // [h][ow][3]
  const int b = blockIdx.x;
  const unsigned char* src = pixels + meta[3 * b];
  const int h = (int)meta[3 * b + 1], w = (int)meta[3 * b + 2];
  const int ksh = axis_ksize(w, ow), ksv = axis_ksize(h, oh);
  if (h < 1 || w < 1 || h > h_cap || ksh > ks_cap || ksv > ks_cap) {  // sized by the host from max_h / max_w: cannot happen
    if (threadIdx.x == 0) atomicExch(status, 1);
    return;
  }
  for (int i = threadIdx.x; i < ow + oh; i += 256) {
    if (i < ow) axis_coeffs(w, ow, ksh, i, kh + i * ks_cap, bh + 2 * i);
    else axis_coeffs(h, oh, ksv, i - ow, kv + (i - ow) * ks_cap, bv + 2 * (i - ow));
  }
  __syncthreads();
  // horizontal pass (skipped by Pillow when the width already matches)
  for (int i = threadIdx.x; i < h * ow; i += 256) {
    const int y = i / ow, xx = i % ow;
    unsigned char* t = tmp + (size_t)i * 3;
    if (w == ow) {
      const unsigned char* p = src + ((size_t)y * w + xx) * 3;
      t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
      continue;
    }
    const int x0 = bh[2 * xx], n = bh[2 * xx + 1];
    const int* k = kh + xx * ks_cap;
    const unsigned char* p = src + ((size_t)y * w + x0) * 3;
    int s0 = 1 << (kPrec - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
      s0 += p[3 * x] * k[x];
      s1 += p[3 * x + 1] * k[x];
      s2 += p[3 * x + 2] * k[x];
    }
    t[0] = clip8(s0); t[1] = clip8(s1); t[2] = clip8(s2);
  }
  __syncthreads();
  // vertical pass + ToTensor (uint8 / 255 -> fp32, CHW)
  float* o = out + (size_t)b * 3 * oh * ow;
  for (int i = threadIdx.x; i < oh * ow; i += 256) {
    const int yy = i / ow, xx = i % ow;
    unsigned char r0, r1, r2;
    if (h == oh) {
      const unsigned char* t = tmp + ((size_t)yy * ow + xx) * 3;
      r0 = t[0]; r1 = t[1]; r2 = t[2];
    } else {
      const int y0 = bv[2 * yy], n = bv[2 * yy + 1];
      const int* k = kv + yy * ks_cap;
      int s0 = 1 << (kPrec - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < n; ++y) {
        const unsigned char* t = tmp + ((size_t)(y0 + y) * ow + xx) * 3;
        s0 += t[0] * k[y];
        s1 += t[1] * k[y];
        s2 += t[2] * k[y];
      }
      r0 = clip8(s0); r1 = clip8(s1); r2 = clip8(s2);
    }
    o[i] = __fdiv_rn((float)r0, 255.f);
    o[oh * ow + i] = __fdiv_rn((float)r1, 255.f);
    o[2 * oh * ow + i] = __fdiv_rn((float)r2, 255.f);
  }
}

int host_ksize(int in_size, int out_size) {
  const double scale = (double)in_size / (double)out_size;
  const double fs = scale > 1.0 ? scale : 1.0;
  return (int)ceil(2.0 * fs) * 2 + 1;
}

}  // namespace

// pixels: packed uint8 RGB crops (device); meta: device int64 [B][3] = {byte offset, h, w}; max_h / max_w: upper bounds of
// the crop sizes in this batch (host values, they size the shared memory); out: fp32 (B, 3, out_h, out_w) in [0, 1];
// status: device int, set to 1 if a crop exceeded the bounds (the output of that crop is then undefined).
extern "C" int focr_resize_bicubic_normalize(const void* pixels, const long long* meta, int B, int max_h, int max_w,
                                             int out_w, int out_h, float* out, int* status, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  FOCR_REQUIRE(B >= 1 && max_h >= 1 && max_w >= 1 && out_w >= 1 && out_h >= 1 && out_w <= 1024 && out_h <= 1024,
               "resize_bicubic_normalize: B=%d max %dx%d out %dx%d", B, max_h, max_w, out_h, out_w);
  int ks = host_ksize(max_w, out_w);
  const int ksv = host_ksize(max_h, out_h);
  if (ksv > ks) ks = ksv;
  const size_t smem = ((size_t)(out_w + out_h) * (ks + 2)) * 4 + (size_t)max_h * out_w * 3 + 16;
  FOCR_REQUIRE(smem <= 200 * 1024, "resize_bicubic_normalize: crops up to %dx%d need %zu bytes of shared memory", max_h, max_w,
               smem);
  static size_t attr = 0;
  if (smem > attr) {
    FOCR_CHECK_CUDA(cudaFuncSetAttribute(resize_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  ProfScope _ps("resize", s);
  FOCR_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  resize_normalize_kernel<<<B, 256, smem, s>>>((const unsigned char*)pixels, meta, out_w, out_h, ks, max_h, out, status);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
