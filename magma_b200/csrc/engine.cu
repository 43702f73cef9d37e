// magma_b200 — model-level runtime: GPT-J (+ MAGMA adapters) forward / backward and CLIP-ViT forward, scheduled in
// C++ on one CUDA stream behind the C ABI. Every matmul goes through the tcgen05 GEMM core (gemm.cu), including
// attention (QK^T, PV and their gradients as strided batched GEMMs straight on the fused qkv buffer), dgrad and
// wgrad (MN-major operands — no transposed copies of weights or activations).
//
// Reference semantics restated here (no code shared):
//   GPT-J block   hf:gptj/modeling_gptj.py:400-413 (parallel residual), attention :166-225 / _attn :129-151
//   adapters      magma/adapters.py:38-39,109-116,63-66,85-92 wired as in magma/magma.py:128-169
//   LM head + CE  hf:gptj/modeling_gptj.py:573,623 ; hf:loss/loss_utils.py:28-67
//   CLIP ViT      hf:clip/modeling_clip.py:138-219,282-386,647-694 (== openai/CLIP VisionTransformer)
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

extern "C" {
int mb200_layernorm_fwd(const void*, int64_t, const void*, const void*, void*, int64_t, float*, float*, int32_t,
                        int32_t, float, void*);
int mb200_layernorm_bwd(const void*, int64_t, const void*, int64_t, const void*, const float*, const float*,
                        const void*, int64_t, void*, int64_t, int32_t, int32_t, void*);
int mb200_rope(void*, int64_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, void*);
int mb200_rope_table(float*, int32_t, int32_t, int32_t, void*);
int mb200_softmax_fwd(const float*, int64_t, int64_t, void*, int64_t, int64_t, int32_t, int32_t, int32_t, float,
                      int32_t, int32_t, void*);
int mb200_softmax_bwd(const float*, int64_t, int64_t, const void*, int64_t, int64_t, void*, int64_t, int64_t, int32_t,
                      int32_t, int32_t, float, void*);
int mb200_cross_entropy(const void*, int64_t, const int64_t*, int32_t, int32_t, int32_t, float*, int32_t*, float*,
                        void*, float, void*);
int mb200_colsum(const void*, int64_t, int32_t, int32_t, float*, int32_t, void*);
int mb200_patchify(const void*, void*, int64_t, int32_t, int32_t, int32_t, void*);
int mb200_vit_assemble(void*, const void*, const void*, const void*, int32_t, int32_t, int32_t, void*);
}

namespace mb200 {

int gemm_impl(const mb200_gemm_args* a, cudaStream_t stream);
bool attn_tile_supported(int S, int hd);
int attn_fwd_tile(const bf16* qkv, long long ld_qkv, bf16* P, long long ldP, bf16* O, long long ldo, int B, int S, int H,
                  int hd, cudaStream_t st);
int attn_bwd_tile(const bf16* qkv, long long ld_qkv, const bf16* dO, long long ld_do, const bf16* P, long long ldP,
                  bf16* dqkv, long long ld_dqkv, const float* rope_tab, int rot, int B, int S, int H, int hd,
                  cudaStream_t st);

// Two-stream execution of the parallel-residual block: the attention branch (qkv -> attention -> out_proj) and the MLP
// branch (fc_in -> fc_out) only meet at the block's final sum, so they are issued on two streams (fork after ln_1, join
// before the sum). Kernels of one branch fill the SMs the other leaves idle (64 pair-tiles on 74 SM pairs for the
// N = 4096 GEMMs, 128-thread attention CTAs, kernel tails).
static bool use_two_streams() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MB200_TWO_STREAM");  // measured: no gain at B=8,S=128 (33.34 vs 33.40 ms) -> off by default
    v = e ? atoi(e) : 0;
  }
  return v != 0;
}
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork[64], join[64];
  bool ok = false;
};
static SideStream& side_stream() {
  static SideStream ss;
  if (!ss.ok) {
    if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) == cudaSuccess) {
      ss.ok = true;
      for (int i = 0; i < 64 && ss.ok; ++i)
        ss.ok = cudaEventCreateWithFlags(&ss.fork[i], cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&ss.join[i], cudaEventDisableTiming) == cudaSuccess;
    }
  }
  return ss;
}

static bool use_attn_tile() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MB200_ATTN_TILE");  // 0 = always use the GEMM-based attention path
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

#define MB_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------
struct Carver {
  uint8_t* base;
  size_t off;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct LayerActs {
  bf16* x_in;     // [M,d] residual stream entering the block
  bf16* h;        // [M,d] ln_1 output
  float* mean;    // [M]
  float* rstd;    // [M]
  bf16* qkv;      // [M,3d] after rotary
  bf16* P;        // [B,H,S,ldP]
  bf16* attn_o;   // [M,d] merged heads, before out_proj
  bf16* pre;      // [M,dff] fc_in pre-activation
  bf16* mlp_out;  // [M,d] fc_out output (adapter input)
  bf16* t_mlp;    // [M,r] mlp-adapter hidden (post ReLU)
  bf16* a_out;    // [M,d] out_proj output (attention-adapter input)
  bf16* t_attn;   // [M,ra]
};

struct GptjPlan {
  int M, d, dff, H, hd, S, B, ldP, ldS;
  LayerActs* acts;  // host array, n_layer entries (inference: all alias one set)
  LayerActs acts_store[64];
  // transient
  float* scores;  // [B,H,S,ldS] fp32 (also dP)
  bf16* hact;     // [M,dff]
  bf16* ax;       // [M,d]
  bf16* x_final;  // [M,d] output of the last block
  bf16* xf_ln;    // [M,d] ln_f output
  float* lnf_mean;
  float* lnf_rstd;
  bf16* dlogits;  // [M,ldv]
  float* row_loss;
  int* n_valid;
  float* rope_tab;  // [S][rot/2] (cos, sin) for positions pos0 .. pos0+S-1
  float* splitk_ws;  // fp32 split-K scratch for small-M (decode) GEMMs: [splits][M][N]
  size_t splitk_bytes;
  // backward temporaries
  bf16* g0;
  bf16* g1;
  bf16* dt;
  bf16* dt2;      // second adapter-hidden gradient buffer (attention adapter runs on the side stream)
  bf16* dm;
  bf16* dhact;
  bf16* dh_mlp;
  bf16* dattn_o;
  bf16* dqkv;
  bf16* dS;
  bf16* dh;
  bf16* da;       // attention-adapter: gradient wrt a_out
  long long ldv;
  size_t bytes;
};

static int make_plan(GptjPlan& P, const mb200_gptj_model* m, int B, int S, int S_kv_max, int training, void* ws) {
  MB_REQUIRE(m->n_layer > 0 && m->n_layer <= 64, MB200_E_SHAPE, "gptj: n_layer=%d out of range", m->n_layer);
  MB_REQUIRE(m->d % m->n_head == 0 && m->d % 8 == 0, MB200_E_SHAPE, "gptj: bad d/n_head");
  Carver c(ws);
  P.B = B;
  P.S = S;
  P.M = B * S;
  P.d = m->d;
  P.dff = m->d_ff;
  P.H = m->n_head;
  P.hd = m->d / m->n_head;
  const int Sk = S_kv_max > S ? S_kv_max : S;
  P.ldP = (int)align_up(Sk, 8);
  P.ldS = (int)align_up(Sk, 8);
  P.ldv = (long long)align_up(m->vocab, 64);
  const size_t M = P.M, d = P.d, dff = P.dff;
  const int rm = m->mlp_adapter ? m->mlp_adapter_r : 0;
  const int ra = m->attn_adapter ? m->attn_adapter_r : 0;
  const int nsets = training ? m->n_layer : 1;
  for (int l = 0; l < nsets; ++l) {
    LayerActs& a = P.acts_store[l];
    a.x_in = c.take<bf16>(M * d);
    a.h = c.take<bf16>(M * d);
    a.mean = c.take<float>(M);
    a.rstd = c.take<float>(M);
    a.qkv = c.take<bf16>(M * 3 * d);
    a.P = c.take<bf16>((size_t)B * P.H * S * P.ldP);
    a.attn_o = c.take<bf16>(M * d);
    a.pre = training ? c.take<bf16>(M * dff) : nullptr;
    a.mlp_out = c.take<bf16>(M * d);
    a.t_mlp = rm ? c.take<bf16>(M * rm) : nullptr;
    a.a_out = c.take<bf16>(M * d);
    a.t_attn = ra ? c.take<bf16>(M * ra) : nullptr;
  }
  for (int l = nsets; l < m->n_layer; ++l) P.acts_store[l] = P.acts_store[0];
  P.acts = P.acts_store;
  P.scores = c.take<float>((size_t)B * P.H * S * P.ldS);
  P.hact = c.take<bf16>(M * dff);
  P.ax = c.take<bf16>(M * d);
  P.x_final = c.take<bf16>(M * d);
  P.xf_ln = c.take<bf16>(M * d);
  P.lnf_mean = c.take<float>(M);
  P.lnf_rstd = c.take<float>(M);
  P.row_loss = c.take<float>(M);
  P.n_valid = c.take<int>(4);
  P.rope_tab = c.take<float>((size_t)S * m->rotary_dim);
  P.splitk_bytes = M <= 128 ? (size_t)16 * M * d * sizeof(float) : 0;  // 16 splits of an [M, d] output
  P.splitk_ws = P.splitk_bytes ? c.take<float>(P.splitk_bytes / sizeof(float)) : nullptr;
  if (training) {
    P.dlogits = c.take<bf16>(M * (size_t)P.ldv);
    P.g0 = c.take<bf16>(M * d);
    P.g1 = c.take<bf16>(M * d);
    const int rmax = rm > ra ? rm : ra;
    P.dt = rmax ? c.take<bf16>(M * rmax) : nullptr;
    P.dt2 = ra ? c.take<bf16>(M * ra) : nullptr;
    P.dm = c.take<bf16>(M * d);
    P.dhact = c.take<bf16>(M * dff);
    P.dh_mlp = c.take<bf16>(M * d);
    P.dattn_o = c.take<bf16>(M * d);
    P.dqkv = c.take<bf16>(M * 3 * d);
    P.dS = c.take<bf16>((size_t)B * P.H * S * P.ldP);
    P.dh = c.take<bf16>(M * d);
    P.da = c.take<bf16>(M * d);
  } else {
    P.dlogits = nullptr;
    P.dt2 = nullptr;
    P.g0 = P.g1 = P.dt = P.dm = P.dhact = P.dh_mlp = P.dattn_o = P.dqkv = P.dS = P.dh = P.da = nullptr;
  }
  P.bytes = align_up(c.off, 256);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GEMM call helpers
// ---------------------------------------------------------------------------------------------
struct Mat {
  const void* p;
  long long ld, bs0, bs1;
  int mn;
};
static inline Mat mat(const void* p, long long ld, int mn = 0, long long bs0 = 0, long long bs1 = 0) {
  return Mat{p, ld, bs0, bs1, mn};
}

struct Epi {
  float alpha = 1.f;
  const void* bias = nullptr;
  int act = 0;
  void* aux_out = nullptr;
  const void* aux_in = nullptr;
  int dact = 0;
  const void* res1 = nullptr;
  const void* res2 = nullptr;
  long long ld_res = 0;
  int accumulate = 0;
  const float* rope_tab = nullptr;
  int rope_mode = 0, rope_S = 0, rope_hd = 0, rope_rot = 0, rope_ncols = 0;
};

// split-K workspace of the pass being issued (set by gptj_forward for small-M passes, else null)
static thread_local float* g_splitk_ws = nullptr;
static thread_local size_t g_splitk_bytes = 0;

static int gemm(cudaStream_t st, int M, int N, int K, Mat A, Mat B, void* C, long long ldc, int c_f32,
                const Epi& e = Epi(), int nb0 = 1, int nb1 = 1, long long c_bs0 = 0, long long c_bs1 = 0) {
  mb200_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.M = M;
  g.N = N;
  g.K = K;
  g.nb0 = nb0;
  g.nb1 = nb1;
  g.c_dtype = c_f32 ? MB200_F32 : MB200_BF16;
  g.A.ptr = A.p;
  g.A.ld = A.ld;
  g.A.bs0 = A.bs0;
  g.A.bs1 = A.bs1;
  g.A.mn_major = A.mn;
  g.B.ptr = B.p;
  g.B.ld = B.ld;
  g.B.bs0 = B.bs0;
  g.B.bs1 = B.bs1;
  g.B.mn_major = B.mn;
  g.C = C;
  g.ldc = ldc;
  g.c_bs0 = c_bs0;
  g.c_bs1 = c_bs1;
  g.alpha = e.alpha;
  g.act = e.act;
  g.dact = e.dact;
  g.accumulate = e.accumulate;
  g.bias = e.bias;
  g.aux_out = e.aux_out;
  g.aux_in = e.aux_in;
  g.res1 = e.res1;
  g.res2 = e.res2;
  g.ld_res = e.ld_res;
  g.rope_tab = e.rope_tab;
  g.rope_mode = e.rope_mode;
  g.rope_S = e.rope_S;
  g.rope_hd = e.rope_hd;
  g.rope_rot = e.rope_rot;
  g.rope_ncols = e.rope_ncols;
  g.splitk_ws = g_splitk_ws;
  g.splitk_ws_bytes = (long long)g_splitk_bytes;
  return gemm_impl(&g, st);
}

// ---------------------------------------------------------------------------------------------
// KV-cache append (prefill and decode): cache[b][h][pos0+s][:] = qkv[b*S+s][which][h][:]
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__  // device code; the host schedule below also compiles as plain C++ for the CPU dry run (oracle/)
__global__ void kv_append_kernel(const bf16* __restrict__ qkv, long long ld, bf16* __restrict__ kc,
                                 bf16* __restrict__ vc, int B, int S, int H, int hd, int Smax, int pos0) {
  const int vec = hd >> 3;
  const long long total = (long long)B * S * H * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vec);
    long long t = i / vec;
    const int h = (int)(t % H);
    t /= H;
    const int s = (int)(t % S);
    const int b = (int)(t / S);
    const long long src = ((long long)b * S + s) * ld + (long long)h * hd + c * 8;
    const long long dst = (((long long)b * H + h) * Smax + (pos0 + s)) * hd + c * 8;
    *reinterpret_cast<uint4*>(kc + dst) = *reinterpret_cast<const uint4*>(qkv + src + (long long)H * hd);
    *reinterpret_cast<uint4*>(vc + dst) = *reinterpret_cast<const uint4*>(qkv + src + 2LL * H * hd);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused decode-step attention (Sq = 1) over the KV cache. One CTA per (b, h): scores in shared memory (fp32),
// softmax in fp32, probabilities rounded to bf16 before P*V exactly like the prefill path / the reference
// (`attn_weights.to(value.dtype)`, hf:gptj/modeling_gptj.py:146). HBM-bound: K and V are each read once, with
// 512-byte coalesced rows.
// ---------------------------------------------------------------------------------------------
static constexpr int kDecThreads = 256;
__global__ void __launch_bounds__(kDecThreads)
attn_decode_kernel(const bf16* __restrict__ qkv, long long ld_qkv, bf16* __restrict__ kc, bf16* __restrict__ vc,
                   bf16* __restrict__ out, long long ld_out, int H, int hd, int Smax, int pos) {
  extern __shared__ float sc[];  // [pos+1] scores, then [32] reduction scratch
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = pos + 1;
  float* red = sc + ((nk + 31) & ~31);
  const bf16* qrow = qkv + (long long)b * ld_qkv + (long long)h * hd;
  bf16* kbase = kc + ((long long)b * H + h) * (long long)Smax * hd;
  bf16* vbase = vc + ((long long)b * H + h) * (long long)Smax * hd;
  // append this step's k, v
  for (int c = threadIdx.x; c < hd; c += kDecThreads) {
    kbase[(long long)pos * hd + c] = qrow[(long long)H * hd + c];
    vbase[(long long)pos * hd + c] = qrow[2LL * H * hd + c];
  }
  __syncthreads();
  const float scale = rsqrtf((float)hd);
  // scores: one warp per key, lanes stride the head dim in 8-element vectors
  const int vecs = hd >> 3;
  for (int j = warp; j < nk; j += kDecThreads / 32) {
    float acc = 0.f;
    for (int v = lane; v < vecs; v += 32) {
      const uint4 ku = *reinterpret_cast<const uint4*>(kbase + (long long)j * hd + v * 8);
      const uint4 qu = *reinterpret_cast<const uint4*>(qrow + v * 8);
      const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&ku);
      const __nv_bfloat162* qh = reinterpret_cast<const __nv_bfloat162*>(&qu);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __bfloat1622float2(kh[e]), q2 = __bfloat1622float2(qh[e]);
        acc += a.x * q2.x + a.y * q2.y;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) sc[j] = acc * scale;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = threadIdx.x; j < nk; j += kDecThreads) m = fmaxf(m, sc[j]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < kDecThreads / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < nk; j += kDecThreads) {
    const float e = __expf(sc[j] - m);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < kDecThreads / 32; ++w) sum += red[w];
  const float inv = 1.f / sum;
  // out[c] = sum_j bf16(p_j) * v[j][c]
  for (int c = threadIdx.x; c < hd; c += kDecThreads) {
    float acc = 0.f;
    for (int j = 0; j < nk; ++j) {
      const float pj = __bfloat162float(__float2bfloat16(sc[j] * inv));
      acc += pj * __bfloat162float(vbase[(long long)j * hd + c]);
    }
    out[(long long)b * ld_out + (long long)h * hd + c] = __float2bfloat16(acc);
  }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// adapter bottleneck forward: out = [alpha](relu(X Wd^T + bd) Wu^T + bu) + res1 + res2 ; saves t = relu(...)
// ---------------------------------------------------------------------------------------------
static int adapter_fwd(cudaStream_t st, const mb200_adapter& ad, int M, int d, int r, const bf16* X, bf16* t,
                       bf16* out, const bf16* res1, const bf16* res2) {
  Epi e1;
  e1.bias = ad.bd;
  e1.act = MB200_ACT_RELU;
  MB_TRY(gemm(st, M, r, d, mat(X, d), mat(ad.wd, d), t, r, 0, e1));
  Epi e2;
  e2.bias = ad.bu;
  e2.res1 = res1;
  e2.res2 = res2;
  e2.ld_res = d;
  MB_TRY(gemm(st, M, d, r, mat(t, r), mat(ad.wu, r), out, d, 0, e2));
  return 0;
}

// adapter backward. g = dL/d(adapter output) [M,d]; X = adapter input; t = saved hidden.
// dX_out = (add_g ? g : 0) + (g Wu ∘ relu'(t)) Wd ; accumulates fp32 wgrads when the adapter has grad buffers.
static int adapter_bwd(cudaStream_t st, const mb200_adapter& ad, int M, int d, int r, const bf16* g, const bf16* X,
                       const bf16* t, bf16* dt, bf16* dX_out, const bf16* res, int accumulate) {
  Epi e1;
  e1.dact = MB200_DACT_RELU;
  e1.aux_in = t;
  MB_TRY(gemm(st, M, r, d, mat(g, d), mat(ad.wu, r, 1), dt, r, 0, e1));  // dt = (g Wu) * 1[t>0]
  if (ad.g_wu) {
    Epi ew;
    ew.accumulate = accumulate;
    MB_TRY(gemm(st, d, r, M, mat(g, d, 1), mat(t, r, 1), ad.g_wu, r, 1, ew));   // dWu[d,r] = g^T t
    MB_TRY(mb200_colsum(g, d, M, d, ad.g_bu, accumulate, st));
    MB_TRY(gemm(st, r, d, M, mat(dt, r, 1), mat(X, d, 1), ad.g_wd, d, 1, ew));  // dWd[r,d] = dt^T X
    MB_TRY(mb200_colsum(dt, r, M, r, ad.g_bd, accumulate, st));
  }
  Epi e2;
  e2.res1 = res;
  e2.ld_res = d;
  MB_TRY(gemm(st, M, d, r, mat(dt, r), mat(ad.wd, d, 1), dX_out, d, 0, e2));  // dX = dt Wd (+ res)
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GPT-J forward
// ---------------------------------------------------------------------------------------------
static int gptj_forward(const mb200_gptj_model* m, const bf16* x, const int64_t* labels, bf16* logits, long long ldv,
                        int last_only, float* loss, bf16* hidden, bf16* kcache, bf16* vcache, int Smax, int pos0,
                        int B, int S, int training, void* ws, size_t ws_bytes, cudaStream_t st) {
  GptjPlan P;
  MB_TRY(make_plan(P, m, B, S, kcache ? Smax : S, training, ws));
  MB_REQUIRE(ws != nullptr && ws_bytes >= P.bytes, MB200_E_ARG, "gptj_forward: workspace too small (%zu < %zu)",
             ws_bytes, P.bytes);
  MB_REQUIRE(!(training && kcache), MB200_E_ARG, "gptj_forward: training with a KV cache is not supported");
  MB_REQUIRE(!kcache || pos0 + S <= Smax, MB200_E_SHAPE, "gptj_forward: pos0+S=%d exceeds cache length %d", pos0 + S,
             Smax);
  MB_REQUIRE(!(labels && last_only), MB200_E_ARG, "gptj_forward: labels with last_only");
  const int M = P.M, d = P.d, dff = P.dff, H = P.H, hd = P.hd;
  const int Sk = kcache ? pos0 + S : S;
  const float scale = 1.0f / sqrtf((float)hd);
  const size_t cache_layer = (size_t)B * H * Smax * hd;

  MB_TRY(mb200_rope_table(P.rope_tab, S, m->rotary_dim, pos0, st));
  struct SplitKScope {  // the workspace is only valid while this pass is being issued
    SplitKScope(float* w, size_t b) { g_splitk_ws = w; g_splitk_bytes = b; }
    ~SplitKScope() { g_splitk_ws = nullptr; g_splitk_bytes = 0; }
  } splitk_scope(P.splitk_ws, P.splitk_bytes);
  SideStream& SS = side_stream();
  const bool two = use_two_streams() && SS.ok && M >= 256;  // decode steps stay on one stream
  const bf16* xin = x;
  if (training) {
    MB_CUDA(cudaMemcpyAsync(P.acts[0].x_in, x, (size_t)M * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    xin = P.acts[0].x_in;
  }
  for (int l = 0; l < m->n_layer; ++l) {
    const mb200_gptj_layer& L = m->layers[l];
    LayerActs& a = P.acts[l];
    bf16* xout = training ? (l + 1 < m->n_layer ? P.acts[l + 1].x_in : P.x_final)
                          : ((l & 1) ? P.x_final : P.acts[0].x_in);
    // ln_1 (one LN feeds both branches of the parallel-residual block)
    MB_TRY(mb200_layernorm_fwd(xin, d, L.ln1_g, L.ln1_b, a.h, d, a.mean, a.rstd, M, d, m->ln_eps, st));
    cudaStream_t sa = st;  // stream of the attention branch
    if (two) {
      MB_CUDA(cudaEventRecord(SS.fork[l], st));
      MB_CUDA(cudaStreamWaitEvent(SS.s, SS.fork[l], 0));
      sa = SS.s;
    }
    // fused q/k/v projection with the rotary embedding applied in the GEMM epilogue (q and k column ranges)
    {
      Epi e;
      e.rope_tab = P.rope_tab;
      e.rope_mode = 1;
      e.rope_S = S;
      e.rope_hd = hd;
      e.rope_rot = m->rotary_dim;
      e.rope_ncols = 2 * d;
      MB_TRY(gemm(sa, M, 3 * d, d, mat(a.h, d), mat(L.w_qkv, d), a.qkv, 3 * d, 0, e));
    }
    if (kcache && S == 1) {
      bf16* kc = kcache + (size_t)l * cache_layer;
      bf16* vc = vcache + (size_t)l * cache_layer;
#ifdef __CUDACC__
      const size_t smem = (((size_t)(pos0 + 1) + 31) & ~(size_t)31) * 4 + 32 * 4;
      attn_decode_kernel<<<B * H, kDecThreads, smem, sa>>>(a.qkv, 3 * d, kc, vc, a.attn_o, d, H, hd, Smax, pos0);
      count_launch();
      MB_CUDA(cudaGetLastError());
#else
      MB_TRY(mb200_attn_decode(a.qkv, 3 * d, kc, vc, a.attn_o, d, B, H, hd, Smax, pos0, sa));
#endif
    } else if (use_attn_tile() && attn_tile_supported(S, hd) && pos0 == 0) {
      // whole sequence in one tile: fused QK^T / softmax / PV kernel, one CTA per (batch, head)
      if (kcache) {
        bf16* kc = kcache + (size_t)l * cache_layer;
        bf16* vc = vcache + (size_t)l * cache_layer;
#ifdef __CUDACC__
        const long long tot = (long long)B * S * H * (hd / 8);
        int grid = (int)((tot + 255) / 256);
        if (grid > num_sms() * 8) grid = num_sms() * 8;
        kv_append_kernel<<<grid, 256, 0, sa>>>(a.qkv, 3 * d, kc, vc, B, S, H, hd, Smax, pos0);
        count_launch();
        MB_CUDA(cudaGetLastError());
#else
        MB_TRY(mb200_kv_append(a.qkv, 3 * d, kc, vc, B, S, H, hd, Smax, pos0, sa));
#endif
      }
      MB_TRY(attn_fwd_tile(a.qkv, 3 * d, a.P, P.ldP, a.attn_o, d, B, S, H, hd, sa));
    } else {
      Mat Q = mat(a.qkv, 3 * d, 0, hd, (long long)S * 3 * d);
      Mat Kk, Vv;
      if (kcache) {
        bf16* kc = kcache + (size_t)l * cache_layer;
        bf16* vc = vcache + (size_t)l * cache_layer;
#ifdef __CUDACC__
        const long long tot = (long long)B * S * H * (hd / 8);
        int grid = (int)((tot + 255) / 256);
        if (grid > num_sms() * 8) grid = num_sms() * 8;
        kv_append_kernel<<<grid, 256, 0, sa>>>(a.qkv, 3 * d, kc, vc, B, S, H, hd, Smax, pos0);
        count_launch();
      MB_CUDA(cudaGetLastError());
#else
        MB_TRY(mb200_kv_append(a.qkv, 3 * d, kc, vc, B, S, H, hd, Smax, pos0, sa));
#endif
        Kk = mat(kc, hd, 0, (long long)Smax * hd, (long long)H * Smax * hd);
        Vv = mat(vc, hd, 1, (long long)Smax * hd, (long long)H * Smax * hd);
      } else {
        Kk = mat(a.qkv + d, 3 * d, 0, hd, (long long)S * 3 * d);
        Vv = mat(a.qkv + 2 * d, 3 * d, 1, hd, (long long)S * 3 * d);
      }
      // scores = Q K^T (fp32), P = softmax(scores / sqrt(hd) + causal mask), O = P V
      MB_TRY(gemm(sa, S, Sk, hd, Q, Kk, P.scores, P.ldS, 1, Epi(), H, B, (long long)S * P.ldS,
                  (long long)H * S * P.ldS));
      MB_TRY(mb200_softmax_fwd(P.scores, P.ldS, (long long)S * P.ldS, a.P, P.ldP, (long long)S * P.ldP, B * H, S, Sk,
                               scale, 1, Sk - S, sa));
      MB_TRY(gemm(sa, S, hd, Sk, mat(a.P, P.ldP, 0, (long long)S * P.ldP, (long long)H * S * P.ldP), Vv, a.attn_o, d,
                  0, Epi(), H, B, hd, (long long)S * d));
    }
    // attention output projection; ax = attention branch + residual x
    if (m->attn_adapter == MB200_ADAPTER_NONE) {
      Epi e;
      e.res1 = xin;
      e.ld_res = d;
      MB_TRY(gemm(sa, M, d, d, mat(a.attn_o, d), mat(L.w_out, d), P.ax, d, 0, e));
    } else if (m->attn_adapter == MB200_ADAPTER_NORMAL) {
      // AdapterWrapper: adapter(attn_out) + attn_out   (magma/adapters.py:109-116)
      MB_TRY(gemm(sa, M, d, d, mat(a.attn_o, d), mat(L.w_out, d), a.a_out, d, 0));
      MB_TRY(adapter_fwd(sa, L.attn_ad, M, d, m->attn_adapter_r, a.a_out, a.t_attn, P.ax, a.a_out, xin));
    } else {
      // ParallelAdapterWrapper: attn(h) + adapter(h), h = ln_1 output   (magma/adapters.py:85-92)
      Epi e;
      e.res1 = xin;
      e.ld_res = d;
      MB_TRY(gemm(sa, M, d, d, mat(a.attn_o, d), mat(L.w_out, d), a.a_out, d, 0, e));
      MB_TRY(adapter_fwd(sa, L.attn_ad, M, d, m->attn_adapter_r, a.h, a.t_attn, P.ax, a.a_out, nullptr));
    }
    if (two) MB_CUDA(cudaEventRecord(SS.join[l], sa));
    // MLP: fc_in + bias + gelu_new (pre-activation saved for backward), then fc_out + bias
    {
      Epi e;
      e.bias = L.b_fc_in;
      e.act = MB200_ACT_GELU_NEW;
      e.aux_out = a.pre;
      MB_TRY(gemm(st, M, dff, d, mat(a.h, d), mat(L.w_fc_in, d), P.hact, dff, 0, e));
    }
    if (m->mlp_adapter == MB200_ADAPTER_NONE) {
      if (two) MB_CUDA(cudaStreamWaitEvent(st, SS.join[l], 0));  // ax is consumed by the fc_out epilogue
      Epi e;
      e.bias = L.b_fc_out;
      e.res1 = P.ax;
      e.ld_res = d;
      MB_TRY(gemm(st, M, d, dff, mat(P.hact, dff), mat(L.w_fc_out, dff), xout, d, 0, e));
    } else if (m->mlp_adapter == MB200_ADAPTER_NORMAL) {
      // nn.Sequential(mlp, Adapter): adapter(mlp(h)) + mlp(h)   (magma/magma.py:143-148, adapters.py:38-39)
      Epi e;
      e.bias = L.b_fc_out;
      MB_TRY(gemm(st, M, d, dff, mat(P.hact, dff), mat(L.w_fc_out, dff), a.mlp_out, d, 0, e));
      if (two) MB_CUDA(cudaStreamWaitEvent(st, SS.join[l], 0));  // ax is consumed by the adapter's up-projection
      MB_TRY(adapter_fwd(st, L.mlp_ad, M, d, m->mlp_adapter_r, a.mlp_out, a.t_mlp, xout, a.mlp_out, P.ax));
    } else {
      // ParallelAdapter: mlp(h) + adapter(h)   (magma/adapters.py:63-66)
      if (two) MB_CUDA(cudaStreamWaitEvent(st, SS.join[l], 0));
      Epi e;
      e.bias = L.b_fc_out;
      e.res1 = P.ax;
      e.ld_res = d;
      MB_TRY(gemm(st, M, d, dff, mat(P.hact, dff), mat(L.w_fc_out, dff), a.mlp_out, d, 0, e));
      MB_TRY(adapter_fwd(st, L.mlp_ad, M, d, m->mlp_adapter_r, a.h, a.t_mlp, xout, a.mlp_out, nullptr));
    }
    xin = xout;
  }
  // ln_f + LM head (+ shifted cross-entropy)
  const bool need_head = logits != nullptr || labels != nullptr || hidden != nullptr;
  if (!need_head) return 0;
  if (last_only) {
    MB_TRY(mb200_layernorm_fwd(xin + (size_t)(S - 1) * d, (long long)S * d, m->lnf_g, m->lnf_b, P.xf_ln, d, nullptr,
                               nullptr, B, d, m->ln_eps, st));
    if (hidden) MB_CUDA(cudaMemcpyAsync(hidden, P.xf_ln, (size_t)B * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
    if (logits) {
      Epi e;
      e.bias = m->b_lm;
      MB_TRY(gemm(st, B, m->vocab, d, mat(P.xf_ln, d), mat(m->w_lm, d), logits, ldv, 0, e));
    }
    return 0;
  }
  if (training && xin != P.x_final) {
    set_error("gptj_forward: internal buffer mismatch");
    return MB200_E_ARG;
  }
  MB_TRY(mb200_layernorm_fwd(xin, d, m->lnf_g, m->lnf_b, P.xf_ln, d, P.lnf_mean, P.lnf_rstd, M, d, m->ln_eps, st));
  if (hidden) MB_CUDA(cudaMemcpyAsync(hidden, P.xf_ln, (size_t)M * d * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  bf16* lg = logits;
  long long ldl = ldv;
  if (!lg && labels) {
    MB_REQUIRE(training, MB200_E_ARG, "gptj_forward: labels without a logits buffer needs training workspace");
    lg = P.dlogits;  // logits live in the workspace and are overwritten by their own gradient
    ldl = P.ldv;
  }
  if (lg) {
    MB_REQUIRE(ldl % 8 == 0 && ldl >= m->vocab, MB200_E_ALIGN, "gptj_forward: ldv=%lld must be >= vocab and %%8", ldl);
    Epi e;
    e.bias = m->b_lm;
    MB_TRY(gemm(st, M, m->vocab, d, mat(P.xf_ln, d), mat(m->w_lm, d), lg, ldl, 0, e));
  }
  if (labels) {
    MB_REQUIRE(loss != nullptr, MB200_E_ARG, "gptj_forward: labels given but loss pointer is NULL");
    bf16* dl = nullptr;
    if (training) {
      MB_REQUIRE(ldl == P.ldv, MB200_E_ARG, "gptj_forward: training needs ldv == %lld (vocab rounded up to 64)", P.ldv);
      dl = P.dlogits;
    }
    MB_TRY(mb200_cross_entropy(lg, ldl, labels, B, S, m->vocab, P.row_loss, P.n_valid, loss, dl, 1.0f, st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GPT-J backward (LM frozen: dgrad through every GEMM, wgrad only for adapters)
// ---------------------------------------------------------------------------------------------
static int gptj_backward(const mb200_gptj_model* m, bf16* dx, float loss_scale, int layer_hi, int layer_lo,
                         int accumulate, int B, int S, void* ws, size_t ws_bytes, cudaStream_t st) {
  GptjPlan P;
  MB_TRY(make_plan(P, m, B, S, S, 1, ws));
  MB_REQUIRE(ws != nullptr && ws_bytes >= P.bytes, MB200_E_ARG, "gptj_backward: workspace too small");
  MB_REQUIRE(0 <= layer_lo && layer_lo <= layer_hi && layer_hi <= m->n_layer, MB200_E_ARG,
             "gptj_backward: bad layer range [%d,%d)", layer_lo, layer_hi);
  const int M = P.M, d = P.d, dff = P.dff, H = P.H, hd = P.hd;
  const float scale = 1.0f / sqrtf((float)hd);
  // gradient w.r.t. the residual stream entering layer l lives in g[(l) & 1]
  auto gbuf = [&](int l) { return (l & 1) ? P.g1 : P.g0; };
  SideStream& SS = side_stream();
  // parallel adapters thread the ln_1-output gradient through both chains sequentially: keep those on one stream
  const bool twob = use_two_streams() && SS.ok && M >= 256 && m->mlp_adapter != MB200_ADAPTER_PARALLEL &&
                    m->attn_adapter != MB200_ADAPTER_PARALLEL;

  if (layer_hi == m->n_layer) {
    // dxf = loss_scale * dlogits Wlm ; g = LN_f backward
    Epi e;
    e.alpha = loss_scale;
    MB_TRY(gemm(st, M, d, m->vocab, mat(P.dlogits, P.ldv), mat(m->w_lm, d, 1), P.dh, d, 0, e));
    MB_TRY(mb200_layernorm_bwd(P.dh, d, P.x_final, d, m->lnf_g, P.lnf_mean, P.lnf_rstd, nullptr, 0,
                               gbuf(m->n_layer), d, M, d, st));
  }
  for (int l = layer_hi - 1; l >= layer_lo; --l) {
    const mb200_gptj_layer& L = m->layers[l];
    LayerActs& a = P.acts[l];
    const bf16* g = gbuf(l + 1);
    bf16* gout = (l == 0 && dx) ? dx : gbuf(l);
    const bf16* dh_acc = nullptr;  // running sum of gradients w.r.t. h = ln_1 output
    cudaStream_t sa = st;
    if (twob) {
      // fork BEFORE the MLP chain is enqueued: g (complete on st here) feeds both chains
      MB_CUDA(cudaEventRecord(SS.fork[l], st));
      MB_CUDA(cudaStreamWaitEvent(SS.s, SS.fork[l], 0));
      sa = SS.s;
    }

    // ---- MLP branch ----
    const bf16* dm = g;
    if (m->mlp_adapter == MB200_ADAPTER_NORMAL) {
      MB_TRY(adapter_bwd(st, L.mlp_ad, M, d, m->mlp_adapter_r, g, a.mlp_out, a.t_mlp, P.dt, P.dm, g, accumulate));
      dm = P.dm;
    } else if (m->mlp_adapter == MB200_ADAPTER_PARALLEL) {
      MB_TRY(adapter_bwd(st, L.mlp_ad, M, d, m->mlp_adapter_r, g, a.h, a.t_mlp, P.dt, P.dm, nullptr, accumulate));
      dh_acc = P.dm;
    }
    {
      Epi e;
      e.dact = MB200_DACT_GELU_NEW;
      e.aux_in = a.pre;
      MB_TRY(gemm(st, M, dff, d, mat(dm, d), mat(L.w_fc_out, dff, 1), P.dhact, dff, 0, e));
      Epi e2;
      e2.res1 = dh_acc;
      e2.ld_res = d;
      MB_TRY(gemm(st, M, d, dff, mat(P.dhact, dff), mat(L.w_fc_in, d, 1), P.dh_mlp, d, 0, e2));
      dh_acc = P.dh_mlp;
    }
    // ---- attention branch (side stream when two-stream execution is on) ----
    const bf16* da = g;
    if (m->attn_adapter == MB200_ADAPTER_NORMAL) {
      MB_TRY(adapter_bwd(sa, L.attn_ad, M, d, m->attn_adapter_r, g, a.a_out, a.t_attn, twob ? P.dt2 : P.dt, P.da, g, accumulate));
      da = P.da;
    } else if (m->attn_adapter == MB200_ADAPTER_PARALLEL) {
      MB_TRY(adapter_bwd(sa, L.attn_ad, M, d, m->attn_adapter_r, g, a.h, a.t_attn, P.dt, P.da, dh_acc, accumulate));
      dh_acc = P.da;
    }
    MB_TRY(gemm(sa, M, d, d, mat(da, d), mat(L.w_out, d, 1), P.dattn_o, d, 0));  // d(attn_o) = da Wo
    if (use_attn_tile() && attn_tile_supported(S, hd)) {
      MB_TRY(attn_bwd_tile(a.qkv, 3 * d, P.dattn_o, d, a.P, P.ldP, P.dqkv, 3 * d, P.rope_tab, m->rotary_dim, B, S, H, hd, sa));
    } else {
      const long long qb0 = hd, qb1 = (long long)S * 3 * d;           // fused-qkv batch strides (h, b)
      const long long pb0 = (long long)S * P.ldP, pb1 = (long long)H * S * P.ldP;
      Mat dO = mat(P.dattn_o, d, 0, hd, (long long)S * d);
      Mat dO_mn = mat(P.dattn_o, d, 1, hd, (long long)S * d);
      // dP = dO V^T (fp32)
      MB_TRY(gemm(sa, S, S, hd, dO, mat(a.qkv + 2 * d, 3 * d, 0, qb0, qb1), P.scores, P.ldS, 1, Epi(), H, B,
                  (long long)S * P.ldS, (long long)H * S * P.ldS));
      // dV = P^T dO
      MB_TRY(gemm(sa, S, hd, S, mat(a.P, P.ldP, 1, pb0, pb1), dO_mn, P.dqkv + 2 * d, 3 * d, 0, Epi(), H, B, qb0,
                  qb1));
      // dS = P * (dP - rowsum(dP * P)) / sqrt(hd)
      MB_TRY(mb200_softmax_bwd(P.scores, P.ldS, (long long)S * P.ldS, a.P, P.ldP, pb0, P.dS, P.ldP, pb0, B * H, S, S,
                               scale, sa));
      // dQ = dS K ; dK = dS^T Q (gradients w.r.t. the rotated q, k) with the inverse rotation fused in the epilogue
      Epi er;
      er.rope_tab = P.rope_tab;
      er.rope_mode = -1;
      er.rope_S = S;
      er.rope_hd = hd;
      er.rope_rot = m->rotary_dim;
      er.rope_ncols = hd;
      MB_TRY(gemm(sa, S, hd, S, mat(P.dS, P.ldP, 0, pb0, pb1), mat(a.qkv + d, 3 * d, 1, qb0, qb1), P.dqkv, 3 * d, 0,
                  er, H, B, qb0, qb1));
      MB_TRY(gemm(sa, S, hd, S, mat(P.dS, P.ldP, 1, pb0, pb1), mat(a.qkv, 3 * d, 1, qb0, qb1), P.dqkv + d, 3 * d, 0,
                  er, H, B, qb0, qb1));
    }
    if (twob) {
      MB_CUDA(cudaEventRecord(SS.join[l], sa));
      MB_CUDA(cudaStreamWaitEvent(st, SS.join[l], 0));
    }
    {
      Epi e;
      e.res1 = dh_acc;
      e.ld_res = d;
      MB_TRY(gemm(st, M, d, 3 * d, mat(P.dqkv, 3 * d), mat(L.w_qkv, d, 1), P.dh, d, 0, e));  // dh = dqkv Wqkv + ...
    }
    // residual + ln_1 backward
    MB_TRY(mb200_layernorm_bwd(P.dh, d, a.x_in, d, L.ln1_g, a.mean, a.rstd, g, d, gout, d, M, d, st));
  }
  if (layer_lo == 0 && dx && m->n_layer == 0) return 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// CLIP ViT forward (inference; the image encoder is frozen on the measured path)
// ---------------------------------------------------------------------------------------------
struct VitPlan {
  int T, M, ldS, ldpatch;
  bf16 *x, *h, *qkv, *P, *attn_o, *hact, *patches, *pooled;
  float* scores;
  size_t bytes;
};
static int make_vit_plan(VitPlan& P, const mb200_vit_model* m, int B, void* ws) {
  MB_REQUIRE(m->image % m->patch == 0 && m->width % m->n_head == 0, MB200_E_SHAPE, "vit: bad geometry");
  const int g = m->image / m->patch;
  Carver c(ws);
  P.T = g * g + 1;
  P.M = B * P.T;
  P.ldS = (int)align_up(P.T, 8);
  P.ldpatch = (int)align_up(3 * m->patch * m->patch, 8);
  const size_t M = P.M, w = m->width;
  P.x = c.take<bf16>(M * w);
  P.h = c.take<bf16>(M * w);
  P.qkv = c.take<bf16>(M * 3 * w);
  P.scores = c.take<float>((size_t)B * m->n_head * P.T * P.ldS);
  P.P = c.take<bf16>((size_t)B * m->n_head * P.T * P.ldS);
  P.attn_o = c.take<bf16>(M * w);
  P.hact = c.take<bf16>(M * (size_t)m->mlp);
  P.patches = c.take<bf16>((size_t)B * g * g * P.ldpatch);
  P.pooled = c.take<bf16>((size_t)B * w);
  P.bytes = align_up(c.off, 256);
  return 0;
}

static int vit_forward(const mb200_vit_model* m, const bf16* images, bf16* feats, int B, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  VitPlan P;
  MB_TRY(make_vit_plan(P, m, B, ws));
  MB_REQUIRE(ws != nullptr && ws_bytes >= P.bytes, MB200_E_ARG, "vit_forward: workspace too small (%zu < %zu)",
             ws_bytes, P.bytes);
  const int w = m->width, H = m->n_head, hd = w / H, T = P.T, M = P.M, g = m->image / m->patch;
  const int Kp = 3 * m->patch * m->patch;
  const float scale = 1.0f / sqrtf((float)hd);
  // conv1 as im2col + GEMM (patch embeddings staged in h), then [cls; patches] + positional embedding
  MB_CUDA(cudaMemsetAsync(P.patches, 0, (size_t)B * g * g * P.ldpatch * sizeof(bf16), st));
  MB_TRY(mb200_patchify(images, P.patches, P.ldpatch, B, m->image, m->patch, st));
  MB_TRY(gemm(st, B * g * g, w, Kp, mat(P.patches, P.ldpatch), mat(m->w_conv, m->ld_conv), P.h, w, 0));
  MB_TRY(mb200_vit_assemble(P.x, P.h, m->cls, m->pos, B, T, w, st));
  // ln_pre (in place: each row is cached in registers before it is rewritten)
  MB_TRY(mb200_layernorm_fwd(P.x, w, m->ln_pre_g, m->ln_pre_b, P.x, w, nullptr, nullptr, M, w, 1e-5f, st));
  const long long qb0 = hd, qb1 = (long long)T * 3 * w;
  const long long pb0 = (long long)T * P.ldS, pb1 = (long long)H * T * P.ldS;
  for (int l = 0; l < m->n_layer; ++l) {
    const mb200_vit_layer& L = m->layers[l];
    MB_TRY(mb200_layernorm_fwd(P.x, w, L.ln1_g, L.ln1_b, P.h, w, nullptr, nullptr, M, w, 1e-5f, st));
    {
      Epi e;
      e.bias = L.b_qkv;
      MB_TRY(gemm(st, M, 3 * w, w, mat(P.h, w), mat(L.w_qkv, w), P.qkv, 3 * w, 0, e));
    }
    MB_TRY(gemm(st, T, T, hd, mat(P.qkv, 3 * w, 0, qb0, qb1), mat(P.qkv + w, 3 * w, 0, qb0, qb1), P.scores, P.ldS, 1,
                Epi(), H, B, pb0, pb1));
    MB_TRY(mb200_softmax_fwd(P.scores, P.ldS, pb0, P.P, P.ldS, pb0, B * H, T, T, scale, 0, 0, st));
    MB_TRY(gemm(st, T, hd, T, mat(P.P, P.ldS, 0, pb0, pb1), mat(P.qkv + 2 * w, 3 * w, 1, qb0, qb1), P.attn_o, w, 0,
                Epi(), H, B, hd, (long long)T * w));
    {
      Epi e;
      e.bias = L.b_out;
      e.res1 = P.x;
      e.ld_res = w;
      MB_TRY(gemm(st, M, w, w, mat(P.attn_o, w), mat(L.w_out, w), P.x, w, 0, e));  // x += out_proj(attn)
    }
    MB_TRY(mb200_layernorm_fwd(P.x, w, L.ln2_g, L.ln2_b, P.h, w, nullptr, nullptr, M, w, 1e-5f, st));
    {
      Epi e;
      e.bias = L.b_fc;
      e.act = MB200_ACT_QUICK_GELU;
      MB_TRY(gemm(st, M, m->mlp, w, mat(P.h, w), mat(L.w_fc, w), P.hact, m->mlp, 0, e));
      Epi e2;
      e2.bias = L.b_proj;
      e2.res1 = P.x;
      e2.ld_res = w;
      MB_TRY(gemm(st, M, w, m->mlp, mat(P.hact, m->mlp), mat(L.w_proj, m->mlp), P.x, w, 0, e2));  // x += mlp
    }
  }
  // ln_post on the class token, then the visual projection
  MB_TRY(mb200_layernorm_fwd(P.x, (long long)T * w, m->ln_post_g, m->ln_post_b, P.pooled, w, nullptr, nullptr, B, w,
                             1e-5f, st));
  MB_TRY(gemm(st, B, m->out_dim, w, mat(P.pooled, w), mat(m->proj_t, w), feats, m->out_dim, 0));
  return 0;
}

}  // namespace mb200

using namespace mb200;

extern "C" size_t mb200_gptj_workspace_bytes(const mb200_gptj_model* m, int32_t B, int32_t S, int32_t S_kv_max,
                                             int32_t training) {
  GptjPlan P;
  if (make_plan(P, m, B, S, S_kv_max, training, nullptr)) return 0;
  return P.bytes;
}

extern "C" int mb200_gptj_forward(const mb200_gptj_model* m, const void* x, const int64_t* labels, void* logits,
                                  int64_t ldv, int32_t last_only, float* loss, void* hidden, void* kcache,
                                  void* vcache, int32_t S_kv_max, int32_t pos0, int32_t B, int32_t S,
                                  int32_t training, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  return gptj_forward(m, (const bf16*)x, labels, (bf16*)logits, ldv, last_only, loss, (bf16*)hidden, (bf16*)kcache,
                      (bf16*)vcache, S_kv_max, pos0, B, S, training, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int mb200_gptj_backward(const mb200_gptj_model* m, void* dx, float loss_scale, int32_t layer_hi,
                                   int32_t layer_lo, int32_t accumulate, int32_t B, int32_t S, void* ws,
                                   size_t ws_bytes, void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  return gptj_backward(m, (bf16*)dx, loss_scale, layer_hi, layer_lo, accumulate, B, S, ws, ws_bytes,
                       (cudaStream_t)stream);
}

extern "C" size_t mb200_vit_workspace_bytes(const mb200_vit_model* m, int32_t B) {
  VitPlan P;
  if (make_vit_plan(P, m, B, nullptr)) return 0;
  return P.bytes;
}

extern "C" int mb200_vit_forward(const mb200_vit_model* m, const void* images, void* feats, int32_t B, void* ws,
                                 size_t ws_bytes, void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  return vit_forward(m, (const bf16*)images, (bf16*)feats, B, ws, ws_bytes, (cudaStream_t)stream);
}

#ifdef __CUDACC__  // kernel wrappers (the CPU dry-run build takes these two from oracle/cabi_emul.cpp)
extern "C" int mb200_attn_decode(const void* qkv, int64_t ld_qkv, void* kcache, void* vcache, void* out,
                                 int64_t ld_out, int32_t B, int32_t H, int32_t hd, int32_t S_kv_max, int32_t pos,
                                 void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  MB_REQUIRE(hd % 8 == 0 && pos >= 0 && pos < S_kv_max, MB200_E_SHAPE, "attn_decode: bad hd=%d pos=%d Smax=%d", hd, pos,
             S_kv_max);
  const size_t smem = (((size_t)(pos + 1) + 31) & ~(size_t)31) * 4 + 32 * 4;
  attn_decode_kernel<<<B * H, kDecThreads, smem, (cudaStream_t)stream>>>((const bf16*)qkv, ld_qkv, (bf16*)kcache,
                                                                         (bf16*)vcache, (bf16*)out, ld_out, H, hd,
                                                                         S_kv_max, pos);
  count_launch();
  MB_CUDA(cudaGetLastError());
  return 0;
}

// K/V rows of a prefill (or any S > 1 continuation) into the static cache — the launch gptj_forward issues, exposed for
// the host-only general schedule (csrc/gptj_sched.cu).
extern "C" int mb200_kv_append(const void* qkv, int64_t ld_qkv, void* kcache, void* vcache, int32_t B, int32_t S,
                               int32_t H, int32_t hd, int32_t S_kv_max, int32_t pos0, void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  MB_REQUIRE(hd % 8 == 0 && B > 0 && S > 0 && pos0 >= 0 && pos0 + S <= S_kv_max, MB200_E_SHAPE,
             "kv_append: bad hd=%d B=%d S=%d pos0=%d Smax=%d", hd, B, S, pos0, S_kv_max);
  const long long tot = (long long)B * S * H * (hd / 8);
  int grid = (int)((tot + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  kv_append_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)qkv, ld_qkv, (bf16*)kcache, (bf16*)vcache, B, S, H,
                                                           hd, S_kv_max, pos0);
  count_launch();
  MB_CUDA(cudaGetLastError());
  return 0;
}
#endif  // __CUDACC__
