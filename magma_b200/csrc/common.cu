// magma_b200 — error plumbing and device queries shared by all translation units.
#include "common.cuh"

#include <stdarg.h>

namespace mb200 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at: %s", (int)e, cudaGetErrorString(e), what);
  return MB200_E_CUDA;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int check_arch() {
  static int cached = 1;  // 1 = unknown
  if (cached == 1) {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
      set_error("no usable CUDA device (%s); magma_b200 has no CPU fallback", cudaGetErrorString(e));
      return MB200_E_ARCH;
    }
    if (major != 10) {
      set_error("device compute capability %d.x is not sm_100; magma_b200 kernels are sm_100a only", major);
      return MB200_E_ARCH;
    }
    cached = 0;
  }
  return cached;
}

}  // namespace mb200

namespace mb200 { const char* last_error(); }

extern "C" int mb200_version(void) { return MB200_VERSION; }
extern "C" const char* mb200_last_error(void) { return mb200::last_error(); }
extern "C" int mb200_check_device(void) { return mb200::check_arch(); }
