// TEST INFRASTRUCTURE — executes CUDA KERNEL SOURCE of the product on the CPU.
//
// The kernels added after the last GPU session (magma_b200/csrc/train_kernels.cuh) have not run on a B200 yet. Their
// schedules are dry-run with emulated operators (cabi_emul.cpp), but that says nothing about the kernel code itself.
// This file compiles the kernel source UNCHANGED (the same .cuh fragments elementwise.cu includes) as host C++ and runs
// it with the CUDA execution model emulated: every thread of a block is an OS thread with its own threadIdx, blocks run
// one after another, `__shared__` is block-shared static storage, `__syncthreads()` a block barrier,
// `__shfl_xor_sync` an exchange through a per-warp buffer, `atomicAdd` a locked add. Indexing, vector handling, masks,
// reductions and accumulation logic of the kernels are thereby checked against torch (tests/test_kernel_source_cpu.py).
// What this cannot show: anything the hardware decides — alignment faults, memory ordering, performance.
//
// Only tests build this (oracle/build_emul.py -> oracle/_build/libkernel_host_exec.so).
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector_types.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// ---- the CUDA execution model, emulated -------------------------------------------------------------------------
struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;

class Barrier {  // reusable barrier for n threads
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    const int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_; });
    }
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};

static Barrier* g_block_barrier = nullptr;
static std::vector<Barrier*> g_warp_barrier;
static float g_xch[1024];
static std::mutex g_atomic;

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __ldg(p) (*(p))
#define __expf expf
static inline void __syncthreads() { g_block_barrier->wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int o) {
  const int t = (int)threadIdx.x, w = t >> 5;
  g_xch[t] = v;
  g_warp_barrier[w]->wait();
  const float r = g_xch[t ^ o];
  g_warp_barrier[w]->wait();
  return r;
}
static inline float atomicAdd(float* p, float v) {
  std::lock_guard<std::mutex> lk(g_atomic);
  const float old = *p;
  *p = old + v;
  return old;
}
template <typename T>
static inline T min(T a, T b) {
  return a < b ? a : b;
}

template <typename F>
static void run_grid(Dim3 grid, int threads, F&& body) {
  gridDim = grid;
  blockDim = Dim3{(unsigned)threads, 1, 1};
  Barrier bb(threads);
  g_block_barrier = &bb;
  std::vector<Barrier*> wb;
  for (int w = 0; w < (threads + 31) / 32; ++w) {
    const int lanes = (w + 1) * 32 <= threads ? 32 : threads - w * 32;
    wb.push_back(new Barrier(lanes));
  }
  g_warp_barrier = wb;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      std::vector<std::thread> ts;
      for (int t = 0; t < threads; ++t)
        ts.emplace_back([&, t] {
          threadIdx = Dim3{(unsigned)t, 1, 1};
          blockIdx = Dim3{bx, by, 1};
          body();
        });
      for (auto& th : ts) th.join();
    }
  for (auto* b : wb) delete b;
}

// ---- the product's kernel source, unchanged ---------------------------------------------------------------------
namespace mb200 {
typedef __nv_bfloat16 bf16;
#include "../magma_b200/csrc/warp_helpers.cuh"
#include "../magma_b200/csrc/elt_helpers.cuh"
#include "../magma_b200/csrc/train_kernels.cuh"
}  // namespace mb200

using namespace mb200;

// ---- launchers: same grids / block sizes as the host wrappers in elementwise.cu ---------------------------------------
static int grid_for(long long n, int threads, int cap = 64) {  // any grid size is legal for the grid-stride kernels
  long long g = (n + threads - 1) / threads;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" {

void hx_quick_gelu_bwd(const void* dy, const void* pre, void* dx, long long n) {
  run_grid(Dim3{(unsigned)grid_for(n / 8, 64), 1, 1}, 64,
           [&] { quick_gelu_bwd_kernel((const bf16*)dy, (const bf16*)pre, (bf16*)dx, n / 8); });
}

void hx_scale_add(const void* u, const float* s, const void* r1, const void* r2, void* out, long long n) {
  run_grid(Dim3{(unsigned)grid_for(n / 8, 64), 1, 1}, 64,
           [&] { scale_add_kernel((const bf16*)u, s, (const bf16*)r1, (const bf16*)r2, (bf16*)out, n / 8); });
}

void hx_dot(const void* a, const void* b, long long n, float* out, int accumulate) {
  if (!accumulate) *out = 0.f;
  run_grid(Dim3{(unsigned)grid_for(n / 8, 256, 4), 1, 1}, 256, [&] { dot_kernel((const bf16*)a, (const bf16*)b, n / 8, out); });
}

void hx_layernorm_param_grad_rows(const void* dy, long long lddy, const void* x, long long ldx, const float* mean,
                                  const float* rstd, float* dgamma, float* dbeta, int rows, int d, int accumulate) {
  if (!accumulate) {
    memset(dgamma, 0, (size_t)d * 4);
    memset(dbeta, 0, (size_t)d * 4);
  }
  run_grid(Dim3{(unsigned)((d + 63) / 64), (unsigned)((rows + kLnPgRows - 1) / kLnPgRows), 1}, 256, [&] {
    layernorm_param_grad_rows_kernel((const bf16*)dy, lddy, (const bf16*)x, ldx, mean, rstd, dgamma, dbeta, rows, d);
  });
}

void hx_col_moments(const void* u, long long ldu, const void* v, long long ldv, const void* mask, long long ldm, int rows,
                    int cols, float* out1, float* out2) {
  memset(out1, 0, (size_t)cols * 4);
  memset(out2, 0, (size_t)cols * 4);
  run_grid(Dim3{(unsigned)((cols + 63) / 64), (unsigned)((rows + kMomRows - 1) / kMomRows), 1}, 256, [&] {
    col_moments_kernel((const bf16*)u, ldu, (const bf16*)v, ldv, (const bf16*)mask, ldm, rows, cols, out1, out2);
  });
}

void hx_channel_affine(const void* x1, const float* a1, const void* x2, const float* a2, const float* c0, const void* mask,
                       const void* res, int relu, void* y, long long rows, int C) {
  run_grid(Dim3{(unsigned)grid_for(rows * (C / 8), 64), 1, 1}, 64, [&] {
    channel_affine_kernel((const bf16*)x1, a1, (const bf16*)x2, a2, c0, (const bf16*)mask, (const bf16*)res, relu, (bf16*)y,
                          rows, C);
  });
}

void hx_col2im3x3(const void* dcols, void* dx, int B, int H, int W, int C, int stride) {
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  run_grid(Dim3{(unsigned)grid_for((long long)B * H * W * (C / 8), 64), 1, 1}, 64,
           [&] { col2im3x3_kernel((const bf16*)dcols, (bf16*)dx, B, H, W, C, stride, Ho, Wo); });
}

void hx_avgpool_nhwc_bwd(const void* dy, void* dx, int B, int H, int W, int C, int k) {
  run_grid(Dim3{(unsigned)grid_for((long long)B * H * W * (C / 8), 64), 1, 1}, 64,
           [&] { avgpool_nhwc_bwd_kernel((const bf16*)dy, (bf16*)dx, B, H, W, C, k); });
}

}  // extern "C"
