/*
 * magma_b200 — C ABI of libmagma_b200.so
 *
 * B200-native (sm_100a) re-backing of the MAGMA forward/backward hot path. The reference
 * (Aleph-Alpha/magma) has no FFI layer of its own: its seams are Python factories and nn.Module
 * duck types. Each entry point below names the reference call site whose arithmetic
 * it replaces. All pointers are raw DEVICE pointers into caller-owned (PyTorch-owned) storage; the
 * library never allocates or frees tensor memory. All work is enqueued on the passed CUDA stream
 * (`void* stream` is a cudaStream_t); no entry point synchronises the device unless stated.
 *
 * Return value: 0 on success, negative MB200_E_* on failure; mb200_last_error() returns a
 * thread-local message. Nothing throws across this boundary. There is no CPU fallback: every
 * compute entry point refuses to run (MB200_E_ARCH) unless the current device is sm_100.
 */
#ifndef MAGMA_B200_H_
#define MAGMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_VERSION 100

enum {
  MB200_OK = 0,
  MB200_E_SHAPE = -1,
  MB200_E_DTYPE = -2,
  MB200_E_ALIGN = -3,
  MB200_E_ARCH = -4,
  MB200_E_CUDA = -5,
  MB200_E_NCCL = -6,
  MB200_E_ARG = -7
};

enum { MB200_BF16 = 0, MB200_F32 = 1 };

/* epilogue activation (forward) */
enum {
  MB200_ACT_NONE = 0,
  MB200_ACT_GELU_NEW = 1,   /* GPT-J MLP: transformers/activations.py:59-66 (NewGELUActivation) */
  MB200_ACT_QUICK_GELU = 2, /* CLIP ViT MLP: x * sigmoid(1.702 x) */
  MB200_ACT_RELU = 3        /* magma/adapters.py:11 (Adapter default activation) */
};
/* epilogue activation-derivative multiplier (backward): out = acc * f'(aux_in) */
enum {
  MB200_DACT_NONE = 0,
  MB200_DACT_GELU_NEW = 1, /* aux_in = pre-activation */
  MB200_DACT_RELU = 3      /* aux_in = relu output (or pre-activation): mask aux_in > 0 */
};

int mb200_version(void);
const char* mb200_last_error(void);
/* 0 when the current CUDA device is sm_100 (B200), MB200_E_ARCH otherwise. */
int mb200_check_device(void);

/* -------------------------------------------------------------------------------------------
 * GEMM core (tcgen05.mma + TMEM accumulators + TMA operand staging, persistent, warp-specialised)
 *
 *   C[b][M,N] = epilogue( alpha * A[b][M,K] * B[b][N,K]^T )
 *
 * Replaces every nn.Linear / torch.matmul on the hot path: GPT-J q/k/v/out/fc_in/fc_out/lm_head
 * (site-packages/transformers/models/gptj/modeling_gptj.py:182-184,222,375-377,623), Adapter
 * down/up projections (magma/adapters.py:19-23), ImagePrefix.proj (magma/image_prefix.py:72,93),
 * CLIP-ViT linears, and their dgrad/wgrad in backward (autograd in train_loop.py:18).
 *
 * Operands are bf16. An operand with mn_major == 0 is stored [rows = M or N][K] (K contiguous);
 * with mn_major == 1 it is stored [K][M or N] (M/N contiguous) — this is what lets dgrad (dY * W)
 * and wgrad (dY^T * X) run on the original tensors without transposed copies. `ld` is the element
 * stride between stored rows, bs0/bs1 the element strides of the two batch indices.
 * Batch index z in [0, nb0*nb1) maps to (z % nb0, z / nb0).
 * Alignment: base pointers 16 B, ld and batch strides multiples of 8 elements.
 *
 * Epilogue order per element (fp32): v = alpha*acc; v += bias[n]; aux_out = v; v = act(v);
 * v *= dact'(aux_in); v += res1 + res2; (accumulate: v += C_old, f32 output only); C = v.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  int64_t ld;
  int64_t bs0, bs1;
  int32_t mn_major;
  int32_t _pad;
} mb200_operand;

typedef struct {
  int32_t M, N, K;
  int32_t nb0, nb1;
  int32_t c_dtype; /* MB200_BF16 or MB200_F32 */
  mb200_operand A, B;
  void* C;
  int64_t ldc, c_bs0, c_bs1;
  float alpha;
  int32_t act;
  int32_t dact;
  int32_t accumulate;
  const void* bias;   /* bf16 [N] or NULL */
  void* aux_out;      /* bf16, same ld / batch strides as C, or NULL */
  const void* aux_in; /* bf16, same ld / batch strides as C, or NULL (required when dact != 0) */
  const void* res1;   /* bf16 [M,N], row stride ld_res, batch strides as C, or NULL */
  const void* res2;   /* bf16 [M,N], row stride ld_res, or NULL */
  int64_t ld_res;
  int32_t force_bn;   /* 0 = auto tile width, else 64/128/256 (testing / tuning) */
  int32_t _pad;
} mb200_gemm_args;

int mb200_gemm(const mb200_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGMA_B200_H_ */
