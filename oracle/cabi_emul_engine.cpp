// TEST INFRASTRUCTURE — the internal (C++) symbols magma_b200/csrc/engine.cu links against, for the CPU dry-run build.
// engine.cu's host schedule (GPT-J forward / backward / prefill / decode, frozen ViT forward) is compiled unchanged as
// plain C++ (oracle/build_emul.py, with oracle/emul_cuda_shim.h force-included); what it calls inside the library —
// gemm_impl, the fused attention entry points, device queries, launch accounting — is forwarded here to the emulated
// primitives of cabi_emul.cpp. Nothing under magma_b200/ links this.
#include "../magma_b200/csrc/common.cuh"

namespace mb200 {

int check_cuda(cudaError_t e, const char* what) {
  set_error("CUDA error %d at: %s (CPU dry run)", (int)e, what);
  return MB200_E_CUDA;
}
int num_sms() { return 148; }
int gemm_sms() { return 148; }
int check_arch() { return 0; }
bool pdl_enabled() { return false; }
void count_launch(int) {}

int gemm_impl(const mb200_gemm_args* a, cudaStream_t stream) { return mb200_gemm(a, (void*)stream); }

bool attn_tile_supported(int S, int hd) { return S >= 1 && S <= 128 && hd >= 64 && hd <= 256 && hd % 64 == 0; }
int attn_fwd_tile(const bf16* qkv, long long ld_qkv, bf16* P, long long ldP, bf16* O, long long ldo, int B, int S, int H,
                  int hd, cudaStream_t st) {
  return mb200_attn_fwd_tile(qkv, ld_qkv, P, ldP, O, ldo, B, S, H, hd, (void*)st);
}
int attn_bwd_tile(const bf16* qkv, long long ld_qkv, const bf16* dO, long long ld_do, const bf16* P, long long ldP,
                  bf16* dqkv, long long ld_dqkv, const float* rope_tab, int rot, int B, int S, int H, int hd,
                  cudaStream_t st) {
  return mb200_attn_bwd_tile(qkv, ld_qkv, dO, ld_do, P, ldP, dqkv, ld_dqkv, rope_tab, rot, B, S, H, hd, (void*)st);
}

}  // namespace mb200
