// Fused optimizer step of the reference training loop (interfaces/super_resolution.py:83-84,
// interfaces/base.py:194-198): torch.nn.utils.clip_grad_norm_(params, 0.25) followed by
// torch.optim.Adam(lr, betas=(0.5, 0.999), eps=1e-8, no weight decay).
// Multi-tensor: a device-resident table of chunks (param, grad, exp_avg, exp_avg_sq pointers, length)
// so that 230 small tensors cost two launches; the clip coefficient and the bias corrections stay on
// the device (no host synchronisation, CUDA-graph friendly).
#include "kernels.cuh"

namespace {

struct OptChunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};

__global__ void __launch_bounds__(256) gradsq_kernel(const OptChunk* __restrict__ chunks, float* __restrict__ partial) {
  __shared__ float red[8];
  const OptChunk c = chunks[blockIdx.x];
  float a = 0.f;
  long long i0 = 0;
  if ((reinterpret_cast<uintptr_t>(c.g) & 15) == 0) {  // 16-byte aligned chunk: vector loads
    const float4* g4 = reinterpret_cast<const float4*>(c.g);
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 g = g4[i];
      a += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
    i0 = n4 << 2;
  }
  for (long long i = i0 + threadIdx.x; i < c.n; i += 256) {
    const float g = c.g[i];
    a += g * g;
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

// state[0] = grad norm (of gscale*g), [1] = clip coef * gscale, [2] = lr / (1-b1^t), [3] = 1/sqrt(1-b2^t);
// step counter (int64) incremented here.
__global__ void optim_finalize_kernel(const float* __restrict__ partial, int n, float gscale, float max_norm,
                                      float lr, float b1, float b2, long long* __restrict__ step,
                                      float* __restrict__ state) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(red[0]) * gscale;
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
    if (coef > 1.f) coef = 1.f;
    const long long t = *step + 1;
    *step = t;
    state[0] = norm;
    state[1] = coef * gscale;
    state[2] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
    state[3] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
  }
}

__global__ void __launch_bounds__(256) adam_kernel(const OptChunk* __restrict__ chunks, const float* __restrict__ state,
                                                   float b1, float b2, float eps) {
  const OptChunk c = chunks[blockIdx.x];
  const float gs = state[1], step_size = state[2], inv_bc2 = state[3];
  long long i0 = 0;
  if (((reinterpret_cast<uintptr_t>(c.g) | reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.m) |
        reinterpret_cast<uintptr_t>(c.v)) & 15) == 0) {
    const float4* g4 = reinterpret_cast<const float4*>(c.g);
    float4* p4 = reinterpret_cast<float4*>(c.p);
    float4* m4 = reinterpret_cast<float4*>(c.m);
    float4* v4 = reinterpret_cast<float4*>(c.v);
    const long long n4 = c.n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 g = g4[i];
      float4 m = m4[i], v = v4[i], w = p4[i];
      const float gx = g.x * gs, gy = g.y * gs, gz = g.z * gs, gw = g.w * gs;
      m.x = b1 * m.x + (1.f - b1) * gx;
      m.y = b1 * m.y + (1.f - b1) * gy;
      m.z = b1 * m.z + (1.f - b1) * gz;
      m.w = b1 * m.w + (1.f - b1) * gw;
      v.x = b2 * v.x + (1.f - b2) * gx * gx;
      v.y = b2 * v.y + (1.f - b2) * gy * gy;
      v.z = b2 * v.z + (1.f - b2) * gz * gz;
      v.w = b2 * v.w + (1.f - b2) * gw * gw;
      w.x -= step_size * m.x / (sqrtf(v.x) * inv_bc2 + eps);
      w.y -= step_size * m.y / (sqrtf(v.y) * inv_bc2 + eps);
      w.z -= step_size * m.z / (sqrtf(v.z) * inv_bc2 + eps);
      w.w -= step_size * m.w / (sqrtf(v.w) * inv_bc2 + eps);
      m4[i] = m;
      v4[i] = v;
      p4[i] = w;
    }
    i0 = n4 << 2;
  }
  for (long long i = i0 + threadIdx.x; i < c.n; i += 256) {
    const float g = c.g[i] * gs;
    const float m = b1 * c.m[i] + (1.f - b1) * g;
    const float v = b2 * c.v[i] + (1.f - b2) * g * g;
    c.m[i] = m;
    c.v[i] = v;
    c.p[i] -= step_size * m / (sqrtf(v) * inv_bc2 + eps);
  }
}

}  // namespace

// chunks: device array of n_chunks records {p, g, m, v, n} (5 x int64); partial: n_chunks floats;
// state: 4 floats (out); step: device int64 counter.
int adam_clip_step(const void* chunks, int n_chunks, float gscale, float max_norm, float lr, float b1, float b2,
                   float eps, long long* step, float* state, float* partial, cudaStream_t s) {
  ProfScope _ps("adam_clip", s);
  FOCR_REQUIRE(n_chunks > 0, "adam_clip_step: empty chunk table");
  const OptChunk* c = reinterpret_cast<const OptChunk*>(chunks);
  gradsq_kernel<<<n_chunks, 256, 0, s>>>(c, partial);
  FOCR_LAUNCH_CHECK();
  optim_finalize_kernel<<<1, 256, 0, s>>>(partial, n_chunks, gscale, max_norm, lr, b1, b2, step, state);
  FOCR_LAUNCH_CHECK();
  adam_kernel<<<n_chunks, 256, 0, s>>>(c, state, b1, b2, eps);
  FOCR_LAUNCH_CHECK();
  return FOCR_OK;
}
