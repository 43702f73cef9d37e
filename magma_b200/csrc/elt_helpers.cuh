// magma_b200 — 8-element bf16 vector helpers and block reductions of the HBM-bound kernels. FRAGMENT: included by
// elementwise.cu inside `namespace mb200` (and, unchanged, by oracle/kernel_host_exec.cpp for CPU execution of kernel source).
// ---------------------------------------------------------------------------------------------
// small vector helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 h1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]);
  __nv_bfloat162 h3 = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}

template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < kThreads / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
template <int kThreads>
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < kThreads / 32) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) {
    t = warp_max(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

