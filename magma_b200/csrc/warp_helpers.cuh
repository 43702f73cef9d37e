// magma_b200 — warp reductions. FRAGMENT: included by common.cuh inside `namespace mb200`, in the device-code section
// (and, unchanged, by oracle/kernel_host_exec.cpp, which executes kernel source on the CPU with emulated shuffles).
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

