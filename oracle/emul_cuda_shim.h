// TEST INFRASTRUCTURE — force-included (g++ -include) when magma_b200/csrc/engine.cu is compiled as plain C++ for the
// CPU dry run: the handful of CUDA runtime calls its HOST schedule makes are mapped onto host memory operations. The
// device code of engine.cu is behind #ifdef __CUDACC__ and is not compiled here; under nvcc this header is never seen and
// the preprocessed engine.cu is byte-identical to the version without the guards.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

extern "C" int mb200_emul_tracing(void);  // cabi_emul.cpp: launch-plan trace active -> nothing touches memory

static inline cudaError_t emul_memcpy_async(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  if (mb200_emul_tracing()) return cudaSuccess;
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t emul_memset_async(void* d, int v, size_t n, cudaStream_t) {
  if (mb200_emul_tracing()) return cudaSuccess;
  memset(d, v, n);
  return cudaSuccess;
}
#define cudaMemcpyAsync emul_memcpy_async
#define cudaMemsetAsync emul_memset_async
#define cudaGetLastError() cudaSuccess
// two-stream execution (off by default) needs real streams and events: report "unavailable" so it stays off
#define cudaStreamCreateWithFlags(...) cudaErrorNotSupported
#define cudaEventCreateWithFlags(...) cudaErrorNotSupported
#define cudaEventRecord(...) cudaSuccess
#define cudaStreamWaitEvent(...) cudaSuccess
