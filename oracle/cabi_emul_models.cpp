// TEST INFRASTRUCTURE — the MODEL-LEVEL entry points of the C ABI, for the CPU emulation library only.
//
// In the product, mb200_gptj_forward / mb200_gptj_backward / mb200_vit_forward live in csrc/engine.cu, which contains
// kernels and runs on a GPU only. So that the DEFAULT Python paths that call them (language_model.py::_run_forward /
// _run_backward with chunked backward, image_encoders.py::B200VisionTransformer.forward, Magma.forward, B200Engine,
// __graft_entry__.smoke) can be replayed on the CPU by tests, this file provides the same entry points by DELEGATING
// to the product's host-only schedules compiled into the same library (csrc/gptj_sched.cu, csrc/vit_train.cu), which
// compute the same arithmetic from the emulated primitives (cabi_emul.cpp), including KV-cache prefill / decode.
//
// What a replay through this file shows: the Python wiring of the default path (pointer tables, workspaces, autograd
// functions, arena, engine) is intact. What it does not show: anything about engine.cu's own schedule or any kernel.
#include <string.h>

#include <vector>

#include "../include/magma_b200.h"

namespace mb200 {
void set_error(const char* fmt, ...);
}

namespace {

struct ExModel {
  mb200_gptj_model_ex m;
  std::vector<mb200_gptj_layer_ex> layers;
};

void to_ex(const mb200_adapter& a, mb200_adapter_ex& e) {
  memset(&e, 0, sizeof(e));
  e.wd = a.wd;
  e.bd = a.bd;
  e.wu = a.wu;
  e.bu = a.bu;
  e.g_wd = a.g_wd;
  e.g_bd = a.g_bd;
  e.g_wu = a.g_wu;
  e.g_bu = a.g_bu;
}

void convert(const mb200_gptj_model* m, ExModel& x) {
  x.layers.resize(m->n_layer);
  for (int l = 0; l < m->n_layer; ++l) {
    const mb200_gptj_layer& s = m->layers[l];
    mb200_gptj_layer_ex& d = x.layers[l];
    d.ln1_g = s.ln1_g;
    d.ln1_b = s.ln1_b;
    d.w_qkv = s.w_qkv;
    d.w_out = s.w_out;
    d.w_fc_in = s.w_fc_in;
    d.b_fc_in = s.b_fc_in;
    d.w_fc_out = s.w_fc_out;
    d.b_fc_out = s.b_fc_out;
    to_ex(s.mlp_ad, d.mlp_ad);
    to_ex(s.attn_ad, d.attn_ad);
  }
  memset(&x.m, 0, sizeof(x.m));
  x.m.n_layer = m->n_layer;
  x.m.d = m->d;
  x.m.n_head = m->n_head;
  x.m.rotary_dim = m->rotary_dim;
  x.m.vocab = m->vocab;
  x.m.d_ff = m->d_ff;
  x.m.mlp_adapter = m->mlp_adapter;
  x.m.mlp_adapter_r = m->mlp_adapter_r;
  x.m.attn_adapter = m->attn_adapter;
  x.m.attn_adapter_r = m->attn_adapter_r;
  x.m.ln_eps = m->ln_eps;
  x.m.layers = x.layers.data();
  x.m.lnf_g = m->lnf_g;
  x.m.lnf_b = m->lnf_b;
  x.m.w_lm = m->w_lm;
  x.m.b_lm = m->b_lm;
}

}  // namespace

extern "C" {

size_t mb200_gptj_workspace_bytes(const mb200_gptj_model* m, int32_t B, int32_t S, int32_t S_kv_max, int32_t) {
  ExModel x;
  convert(m, x);
  const size_t a = mb200_gptj_sched_workspace_bytes(&x.m, B, S);
  const size_t b = mb200_gptj_sched_infer_workspace_bytes(&x.m, B, S, S_kv_max);
  return a > b ? a : b;
}

// labels (a loss, and possibly a backward pass afterwards) -> the training pass of the general schedule; everything
// else (plain logits, KV-cache prefill / decode, last-position logits, hidden output) -> its inference pass
int mb200_gptj_forward(const mb200_gptj_model* m, const void* xin, const int64_t* labels, void* logits, int64_t ldv,
                       int32_t last_only, float* loss, void* hidden, void* kcache, void* vcache, int32_t S_kv_max,
                       int32_t pos0, int32_t B, int32_t S, int32_t training, void* ws, size_t ws_bytes, void* stream) {
  ExModel x;
  convert(m, x);
  if (labels || training) {
    if (kcache || vcache || last_only || hidden || pos0 != 0) {
      mb200::set_error("emulation: a loss together with a KV cache / last-position / hidden output is not a path of the product");
      return MB200_E_ARG;
    }
    return mb200_gptj_sched_forward(&x.m, xin, labels, logits, ldv, loss, B, S, ws, ws_bytes, stream);
  }
  return mb200_gptj_sched_infer(&x.m, xin, logits, ldv, last_only, hidden, kcache, vcache, S_kv_max, pos0, B, S, ws,
                                ws_bytes, stream);
}

int mb200_gptj_backward(const mb200_gptj_model* m, void* dx, float loss_scale, int32_t layer_hi, int32_t layer_lo,
                        int32_t accumulate, int32_t B, int32_t S, void* ws, size_t ws_bytes, void* stream) {
  ExModel x;
  convert(m, x);
  return mb200_gptj_sched_backward_range(&x.m, dx, loss_scale, layer_hi, layer_lo, accumulate, B, S, ws, ws_bytes, stream);
}

size_t mb200_vit_workspace_bytes(const mb200_vit_model* m, int32_t B) { return mb200_vit_train_workspace_bytes(m, B); }

int mb200_vit_forward(const mb200_vit_model* m, const void* images, void* feats, int32_t B, void* ws, size_t ws_bytes,
                      void* stream) {
  return mb200_vit_forward_train(m, images, feats, B, ws, ws_bytes, stream);
}

}  // extern "C"
