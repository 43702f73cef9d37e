// magma_b200 — HBM-bound kernels of the hot path (normalisation, rotary, softmax, cross-entropy, gathers,
// label building, reductions). All use 128-bit loads/stores where the layout allows, fp32 math, bf16 storage.
#include "common.cuh"

namespace mb200 {

#include "elt_helpers.cuh"

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta ; one CTA per row, row cached in registers.
// (torch.nn.LayerNorm in GPT-J ln_1/ln_f, CLIP ln_*, magma/image_prefix.py:106-107)
// ---------------------------------------------------------------------------------------------
static constexpr int kLnThreads = 256;
static constexpr int kLnMaxVec = 4;  // d <= 256*4*8 = 8192

__global__ void __launch_bounds__(kLnThreads)
layernorm_fwd_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ gamma,
                     const bf16* __restrict__ beta, bf16* __restrict__ y, long long ldy, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int d, float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const int nvec = d >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  float v[kLnMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
      unpack8(xr[c], v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = block_sum<kLnThreads>(s, red) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = v[i][e] - mean;
        q += t * t;
      }
    }
  }
  const float var = block_sum<kLnThreads>(q, red) / (float)d;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
  const uint4* gr = reinterpret_cast<const uint4*>(gamma);
  const uint4* br = reinterpret_cast<const uint4*>(beta);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
      float g[8], b[8], o[8];
      unpack8(__ldg(gr + c), g);
      unpack8(__ldg(br + c), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
      yr[c] = pack8(o);
    }
  }
}

// LayerNorm backward (input gradient): dx = res + rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma
__global__ void __launch_bounds__(kLnThreads)
layernorm_bwd_kernel(const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ x, long long ldx,
                     const bf16* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const bf16* __restrict__ res, long long ldres, bf16* __restrict__ dx, long long lddx, int d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const int nvec = d >> 3;
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + row * lddy);
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  const uint4* gr = reinterpret_cast<const uint4*>(gamma);
  const float mu = mean[row], rs = rstd[row];
  float g[kLnMaxVec][8], xh[kLnMaxVec][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
      float a[8], b[8], gm[8];
      unpack8(dyr[c], a);
      unpack8(xr[c], b);
      unpack8(__ldg(gr + c), gm);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g[i][e] = a[e] * gm[e];
        xh[i][e] = (b[e] - mu) * rs;
        s1 += g[i][e];
        s2 += g[i][e] * xh[i][e];
      }
    }
  }
  const float m1 = block_sum<kLnThreads>(s1, red) / (float)d;
  const float m2 = block_sum<kLnThreads>(s2, red) / (float)d;
  uint4* dxr = reinterpret_cast<uint4*>(dx + row * lddx);
  const uint4* rr = res ? reinterpret_cast<const uint4*>(res + row * ldres) : nullptr;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int c = threadIdx.x + i * kLnThreads;
    if (c < nvec) {
      float o[8], r[8];
      if (rr) unpack8(rr[c], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o[e] = rs * (g[i][e] - m1 - xh[i][e] * m2);
        if (rr) o[e] += r[e];
      }
      dxr[c] = pack8(o);
    }
  }
}

// LayerNorm parameter gradients for a (small) number of rows: dgamma[c] (+)= sum_r dy*xhat, dbeta[c] (+)= sum_r dy
__global__ void layernorm_param_grad_kernel(const bf16* __restrict__ dy, long long lddy, const bf16* __restrict__ x,
                                            long long ldx, const float* __restrict__ mean,
                                            const float* __restrict__ rstd, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int rows, int d, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float sg = 0.f, sb = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float g = __bfloat162float(dy[(long long)r * lddy + c]);
    const float xh = (__bfloat162float(x[(long long)r * ldx + c]) - mean[r]) * rstd[r];
    sg += g * xh;
    sb += g;
  }
  dgamma[c] = accumulate ? dgamma[c] + sg : sg;
  dbeta[c] = accumulate ? dbeta[c] + sb : sb;
}

// ---------------------------------------------------------------------------------------------
// Rotary embedding, in place on the fused qkv buffer [rows = B*S][3][H][hd] (q and k, first rot dims of each
// head; interleaved pairs — rotate_every_two, modeling_gptj.py:57-67). inverse != 0 applies the transpose
// rotation (backward). position of row r = pos0 + (r % S).
// ---------------------------------------------------------------------------------------------
__global__ void rope_kernel(bf16* __restrict__ qkv, long long ld, int rows, int S, int H, int hd, int rot, int pos0,
                            int inverse) {
  const int half = rot >> 1;
  const long long total = (long long)rows * 2 * H * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % half);
    long long t = i / half;
    const int h = (int)(t % H);
    t /= H;
    const int which = (int)(t % 2);
    const long long r = t / 2;
    const int pos = pos0 + (int)(r % S);
    // inv_freq = 10000^(-2p/rot); fp32 like create_sinusoidal_positions (modeling_gptj.py:47-50)
    const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * p) / (float)rot);
    float sn, cs;
    sincosf((float)pos * inv_freq, &sn, &cs);
    if (inverse) sn = -sn;
    __nv_bfloat162* ptr =
        reinterpret_cast<__nv_bfloat162*>(qkv + r * ld + (long long)which * H * hd + (long long)h * hd + 2 * p);
    const float2 v = __bfloat1622float2(*ptr);
    *ptr = __floats2bfloat162_rn(v.x * cs - v.y * sn, v.y * cs + v.x * sn);
  }
}

__global__ void rope_table_kernel(float2* __restrict__ tab, int S, int half, int rot, int pos0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * half) return;
  const int p = i % half, s = i / half;
  const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * p) / (float)rot);
  float sn, cs;
  sincosf((float)(pos0 + s) * inv_freq, &sn, &cs);
  tab[i] = make_float2(cs, sn);
}

// ---------------------------------------------------------------------------------------------
// Softmax over fp32 scores -> bf16 probabilities. One warp per row. causal: key j visible iff j <= i + koff.
// (GPTJAttention._attn, modeling_gptj.py:136-147: fp32 scores / sqrt(hd), mask, softmax, cast to value dtype)
// rows are indexed (z, i): z = batch*head, i in [0, Sq).
// ---------------------------------------------------------------------------------------------
__global__ void softmax_fwd_kernel(const float* __restrict__ s, long long lds, long long s_bs, bf16* __restrict__ p,
                                   long long ldp, long long p_bs, int nz, int Sq, int Sk, float scale, int causal,
                                   int koff) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)nz * Sq) return;
  const int z = (int)(wid / Sq), i = (int)(wid % Sq);
  const float* sr = s + (long long)z * s_bs + (long long)i * lds;
  bf16* pr = p + (long long)z * p_bs + (long long)i * ldp;
  const int lim = causal ? min(Sk, i + koff + 1) : Sk;
  float m = -INFINITY;
  for (int j = lane; j < lim; j += 32) m = fmaxf(m, sr[j] * scale);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < lim; j += 32) sum += __expf(sr[j] * scale - m);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < Sk; j += 32) pr[j] = __float2bfloat16(j < lim ? __expf(sr[j] * scale - m) * inv : 0.f);
}

// dS = P * (dP - sum_j dP*P) * scale  (bf16 out)
__global__ void softmax_bwd_kernel(const float* __restrict__ dp, long long lddp, long long dp_bs,
                                   const bf16* __restrict__ p, long long ldp, long long p_bs, bf16* __restrict__ ds,
                                   long long ldds, long long ds_bs, int nz, int Sq, int Sk, float scale) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)nz * Sq) return;
  const int z = (int)(wid / Sq), i = (int)(wid % Sq);
  const float* dpr = dp + (long long)z * dp_bs + (long long)i * lddp;
  const bf16* pr = p + (long long)z * p_bs + (long long)i * ldp;
  bf16* dsr = ds + (long long)z * ds_bs + (long long)i * ldds;
  float acc = 0.f;
  for (int j = lane; j < Sk; j += 32) acc += dpr[j] * __bfloat162float(pr[j]);
  acc = warp_sum(acc);
  for (int j = lane; j < Sk; j += 32) dsr[j] = __float2bfloat16(__bfloat162float(pr[j]) * (dpr[j] - acc) * scale);
}

// ---------------------------------------------------------------------------------------------
// build_labels (magma/utils.py:334-364), integer kernel, bit-exact:
//   labels[b, s] = -100                      for s < L
//                = captions[b, s - L]        for s >= L           (captions[:, :-L])
//   then every position AFTER the first eos in the row -> -100 (the first eos itself is kept).
// One warp per row; ballot scan for the first eos.
// ---------------------------------------------------------------------------------------------
__global__ void build_labels_kernel(const long long* __restrict__ captions, long long ldc,
                                    long long* __restrict__ labels, int B, int S, int L, long long eos) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  int first = S;  // index in label space of the first eos
  for (int base = L; base < S && first == S; base += 32) {
    const int s = base + lane;
    const bool hit = s < S && captions[(long long)b * ldc + (s - L)] == eos;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) first = base + __ffs(m) - 1;
  }
  for (int s = lane; s < S; s += 32) {
    long long v;
    if (s < L)
      v = -100;
    else
      v = captions[(long long)b * ldc + (s - L)];
    if (s > first) v = -100;
    labels[(long long)b * S + s] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Input assembly (magma/magma.py:258-267): x[b, 0:L] = prefix[b]; x[b, L+s] = wte[captions[b, s]], s < S-L.
// Gathers straight into the fused [B,S,d] buffer (the reference embeds the full caption then slices).
// ---------------------------------------------------------------------------------------------
__global__ void embed_assemble_kernel(const long long* __restrict__ captions, long long ldc,
                                      const bf16* __restrict__ wte, const bf16* __restrict__ prefix, int L,
                                      bf16* __restrict__ x, int B, int S, int d, int vocab) {
  const int row = blockIdx.x;  // b*S + s
  const int b = row / S, s = row % S;
  const uint4* src;
  if (s < L) {
    src = reinterpret_cast<const uint4*>(prefix + ((long long)b * L + s) * d);
  } else {
    long long tok = captions[(long long)b * ldc + (s - L)];
    if (tok < 0 || tok >= vocab) tok = 0;  // defensive: never read out of bounds
    src = reinterpret_cast<const uint4*>(wte + tok * (long long)d);
  }
  uint4* dst = reinterpret_cast<uint4*>(x + (long long)row * d);
  for (int c = threadIdx.x; c < (d >> 3); c += blockDim.x) dst[c] = __ldg(src + c);
}

// plain row gather: out[r] = wte[ids[r]]  (Magma.embed / decode-step input_ids path)
__global__ void embed_gather_kernel(const long long* __restrict__ ids, const bf16* __restrict__ wte,
                                    bf16* __restrict__ out, int d, int vocab) {
  long long tok = ids[blockIdx.x];
  if (tok < 0 || tok >= vocab) tok = 0;
  const uint4* src = reinterpret_cast<const uint4*>(wte + tok * (long long)d);
  uint4* dst = reinterpret_cast<uint4*>(out + (long long)blockIdx.x * d);
  for (int c = threadIdx.x; c < (d >> 3); c += blockDim.x) dst[c] = __ldg(src + c);
}

// ---------------------------------------------------------------------------------------------
// Shifted cross-entropy over bf16 logits [M = B*S][ldv] (ForCausalLMLoss, loss_utils.py:28-67):
// target of row (b, s) is labels[b, s+1]; the last position and label -100 are ignored; mean over valid.
// Pass 1 (ce_count): n_valid. Pass 2 (ce_row): per-row loss + in-place dlogits = (softmax - onehot)/n_valid.
// Pass 3 (ce_reduce): deterministic tree sum of row losses -> mean.
// ---------------------------------------------------------------------------------------------
__global__ void ce_count_kernel(const long long* __restrict__ labels, int B, int S, int* __restrict__ n_valid) {
  __shared__ float red[32];
  float c = 0.f;
  for (int i = threadIdx.x; i < B * S; i += blockDim.x) {
    const int s = i % S;
    if (s + 1 < S && labels[i + 1] != -100) c += 1.f;
  }
  const float t = block_sum<1024>(c, red);
  if (threadIdx.x == 0) *n_valid = (int)(t + 0.5f);
}

static constexpr int kCeThreads = 512;
__global__ void __launch_bounds__(kCeThreads)
ce_row_kernel(const bf16* logits, long long ldv, const long long* __restrict__ labels, int S, int V,
              const int* __restrict__ n_valid, float* __restrict__ row_loss, bf16* dlogits, float grad_scale) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int s = row % S;
  long long tgt = (s + 1 < S) ? labels[row + 1] : -100;
  const bf16* lr = logits + (long long)row * ldv;
  bf16* dr = dlogits ? dlogits + (long long)row * ldv : nullptr;
  const int nvec = V >> 3;
  if (tgt == -100 || tgt < 0 || tgt >= V) {
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    if (dr) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int c = threadIdx.x; c < nvec; c += kCeThreads) reinterpret_cast<uint4*>(dr)[c] = z;
      for (int j = nvec * 8 + threadIdx.x; j < V; j += kCeThreads) dr[j] = __float2bfloat16(0.f);
    }
    return;
  }
  float m = -INFINITY;
  for (int c = threadIdx.x; c < nvec; c += kCeThreads) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(lr)[c], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, f[e]);
  }
  for (int j = nvec * 8 + threadIdx.x; j < V; j += kCeThreads) m = fmaxf(m, __bfloat162float(lr[j]));
  m = block_max<kCeThreads>(m, red);
  float sum = 0.f;
  for (int c = threadIdx.x; c < nvec; c += kCeThreads) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(lr)[c], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += __expf(f[e] - m);
  }
  for (int j = nvec * 8 + threadIdx.x; j < V; j += kCeThreads) sum += __expf(__bfloat162float(lr[j]) - m);
  sum = block_sum<kCeThreads>(sum, red);
  const float lse = m + logf(sum);
  const float tl = __bfloat162float(lr[tgt]);
  __syncthreads();  // every thread has read lr[tgt] before anyone overwrites it (dlogits may alias logits)
  if (threadIdx.x == 0) row_loss[row] = lse - tl;
  if (dr) {
    const float gs = grad_scale / (float)max(*n_valid, 1);
    const float inv = 1.f / sum;
    for (int c = threadIdx.x; c < nvec; c += kCeThreads) {
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(lr)[c], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float pj = __expf(f[e] - m) * inv;
        if (c * 8 + e == tgt) pj -= 1.f;
        f[e] = pj * gs;
      }
      reinterpret_cast<uint4*>(dr)[c] = pack8(f);
    }
    for (int j = nvec * 8 + threadIdx.x; j < V; j += kCeThreads) {
      float pj = __expf(__bfloat162float(lr[j]) - m) * inv;
      if (j == tgt) pj -= 1.f;
      dr[j] = __float2bfloat16(pj * gs);
    }
  }
}

__global__ void ce_reduce_kernel(const float* __restrict__ row_loss, int M, const int* __restrict__ n_valid,
                                 float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) s += row_loss[i];
  const float t = block_sum<1024>(s, red);
  if (threadIdx.x == 0) *loss = t / (float)max(*n_valid, 1);
}

// ---------------------------------------------------------------------------------------------
// column sum: out[c] (+)= sum_r x[r, c]   (bias gradients). grid.x covers 64-column strips.
// ---------------------------------------------------------------------------------------------
static constexpr int kColsumRows = 64;  // rows per CTA: grid = (col strips of 64) x (row chunks) for parallelism
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, long long ldx, int rows, int cols, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[8][64];
  const int cl = threadIdx.x & 31;        // column pair within the strip
  const int rg = threadIdx.x >> 5;        // row group 0..7
  const int c0 = blockIdx.x * 64 + cl * 2;
  const int r0 = blockIdx.y * kColsumRows;
  const int r1 = min(rows, r0 + kColsumRows);
  float a0 = 0.f, a1 = 0.f;
  if (c0 < cols) {
    for (int r = r0 + rg; r < r1; r += 8) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (long long)r * ldx + c0));
      a0 += v.x;
      a1 += v.y;
    }
  }
  part[rg][cl * 2] = a0;
  part[rg][cl * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < cols) atomicAdd(out + c, s);  // one atomic per (column, row chunk); out is zeroed first unless accumulating
  }
}

// ---------------------------------------------------------------------------------------------
// dropout with a counter-based hash RNG (nn.Dropout in magma/image_prefix.py:104): y = x * mask / (1-p),
// mask saved as bytes for the backward pass. Also used as the backward (same mask, same scale).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (uint32_t)k;
}
__global__ void dropout_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ mask,
                                   long long n, float p, unsigned long long seed) {
  const float scale = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = (float)(hash32(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i) >> 8) * (1.0f / 16777216.0f);
    const uint8_t keep = u >= p ? 1 : 0;
    mask[i] = keep;
    y[i] = __float2bfloat16(keep ? __bfloat162float(x[i]) * scale : 0.f);
  }
}
__global__ void dropout_apply_kernel(const bf16* __restrict__ x, const uint8_t* __restrict__ mask,
                                     bf16* __restrict__ y, long long n, float p) {
  const float scale = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(mask[i] ? __bfloat162float(x[i]) * scale : 0.f);
}

// ---------------------------------------------------------------------------------------------
// ViT front end: im2col of non-overlapping patches (CLIP conv1, stride = kernel = P, no bias) and the
// [cls; patches] + positional embedding assembly.
// images [B,3,R,R] -> patches [B*g*g][ldp], column order (c, py, px) = conv weight.view(width, 3*P*P) order.
// ---------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ img, bf16* __restrict__ patches, long long ldp, int B, int R,
                                int P) {
  const int g = R / P;
  const int K = 3 * P * P;
  const long long total = (long long)B * g * g * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long pr = i / K;  // patch row index b*g*g + gy*g + gx
    const int gx = (int)(pr % g);
    const int gy = (int)((pr / g) % g);
    const int b = (int)(pr / ((long long)g * g));
    const int px = k % P, py = (k / P) % P, c = k / (P * P);
    patches[pr * ldp + k] = img[(((long long)b * 3 + c) * R + (gy * P + py)) * R + (gx * P + px)];
  }
}
// x[b, 0, :] = cls + pos[0]; x[b, 1+p, :] = pe[b, p, :] + pos[1+p]   (pe = patch embeddings [B, T-1, w])
__global__ void vit_assemble_kernel(bf16* __restrict__ x, const bf16* __restrict__ pe, const bf16* __restrict__ cls,
                                    const bf16* __restrict__ pos, int B, int T, int w) {
  const long long total = (long long)B * T * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % w);
    const int t = (int)((i / w) % T);
    const long long b = i / ((long long)w * T);
    const float base = t == 0 ? __bfloat162float(cls[c]) : __bfloat162float(pe[(b * (T - 1) + (t - 1)) * w + c]);
    x[i] = __float2bfloat16(base + __bfloat162float(pos[(long long)t * w + c]));
  }
}

// ---------------------------------------------------------------------------------------------
// Conv-trunk support (CLIP ModifiedResNet, image_encoders.py:65-74). Activations are NHWC bf16, so a 1x1 convolution is
// a plain GEMM over [B*H*W, C]; a 3x3 convolution is im2col (column order (kh, kw, c), matching weights packed as
// [Cout][3][3][Cin]) followed by the same GEMM with the folded BatchNorm as bias. All three kernels move 16-byte
// vectors of 8 channels and are HBM-bound.
// ---------------------------------------------------------------------------------------------
// images [B, C<=8, H, W] bf16 -> [B, H, W, 8] bf16, channels C..7 zero (so that the stem's K = 9*8 is TMA-aligned)
__global__ void nchw_to_nhwc8_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int C, int H, int W) {
  const long long hw = (long long)H * W, total = (long long)B * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, px = i - b * hw;
    alignas(16) bf16 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < C ? src[(b * C + c) * hw + px] : __float2bfloat16(0.f);
    *reinterpret_cast<uint4*>(dst + i * 8) = *reinterpret_cast<const uint4*>(v);
  }
}
// src [B,H,W,C] -> dst [B*Ho*Wo][9*C], 3x3 window, padding 1, stride s (Ho = (H-1)/s + 1); out-of-image taps are zero
template <typename Idx>  // 32-bit index arithmetic whenever the vector count allows it (the divisions dominate otherwise)
__global__ void im2col3x3_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int H, int W, int C,
                                 int stride, int Ho, int Wo) {
  const Idx cv = (Idx)(C >> 3);
  const Idx total = (Idx)B * Ho * Wo * 9 * cv;
  for (Idx i = blockIdx.x * (Idx)blockDim.x + threadIdx.x; i < total; i += (Idx)gridDim.x * blockDim.x) {
    const Idx t = i / cv;
    const int c8 = (int)(i - t * cv);
    const Idx row = t / 9;
    const int tap = (int)(t - row * 9);
    const Idx r2 = row / Wo;
    const int wo = (int)(row - r2 * Wo);
    const Idx b = r2 / Ho;
    const int ho = (int)(r2 - b * Ho);
    const int hi = ho * stride - 1 + tap / 3, wi = wo * stride - 1 + tap % 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      v = __ldg(reinterpret_cast<const uint4*>(src + (((long long)b * H + hi) * W + wi) * C) + c8);
    reinterpret_cast<uint4*>(dst)[i] = v;  // (row, tap, c8) is exactly the linear index
  }
}
// nn.AvgPool2d(k) on NHWC: dst [B, H/k, W/k, C], fp32 accumulation
__global__ void avgpool_nhwc_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int H, int W, int C,
                                    int k) {
  const int cv = C >> 3, Ho = H / k, Wo = W / k;
  const long long total = (long long)B * Ho * Wo * cv;
  const float inv = 1.f / (float)(k * k);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const long long px = i / cv;
    const int wo = (int)(px % Wo);
    const int ho = (int)((px / Wo) % Ho);
    const long long b = px / ((long long)Wo * Ho);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + ((b * H + ho * k + dy) * W + wo * k + dx) * C) + c8);
        const bf16* e = reinterpret_cast<const bf16*>(&u);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += __bfloat162float(e[c]);
      }
    alignas(16) bf16 o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = __float2bfloat16(acc[c] * inv);
    reinterpret_cast<uint4*>(dst)[i] = *reinterpret_cast<const uint4*>(o);
  }
}

// ---------------------------------------------------------------------------------------------
// argmax over the last dim of bf16 rows, compared in fp32 like sampling.py:92,97 (logits.float(); argmax).
// Ties resolve to the LOWEST index (torch.argmax on CPU/CUDA returns the first maximal element).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
argmax_kernel(const bf16* __restrict__ x, long long ldx, int V, long long* __restrict__ out) {
  __shared__ float sv[16];
  __shared__ int si[16];
  const bf16* r = x + (long long)blockIdx.x * ldx;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float v = __bfloat162float(r[j]);
    if (v > best || (v == best && j < bi)) {
      best = v;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sv[w] = best;
    si[w] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      if (sv[k] > best || (sv[k] == best && si[k] < bi)) {
        best = sv[k];
        bi = si[k];
      }
    out[blockIdx.x] = bi;
  }
}

// y = a + b (+ c) elementwise, bf16 (gradient joins)
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ c,
                           bf16* __restrict__ y, long long nvec) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float fa[8], fb[8], fc[8];
    unpack8(reinterpret_cast<const uint4*>(a)[i], fa);
    unpack8(reinterpret_cast<const uint4*>(b)[i], fb);
    if (c) unpack8(reinterpret_cast<const uint4*>(c)[i], fc);
#pragma unroll
    for (int e = 0; e < 8; ++e) fa[e] += fb[e] + (c ? fc[e] : 0.f);
    reinterpret_cast<uint4*>(y)[i] = pack8(fa);
  }
}

// ---------------------------------------------------------------------------------------------
// Optimizer step for the (small) trainable set: fused AdamW over a flat fp32 arena (torch.optim.AdamW semantics,
// train.py:96-101 betas=(0.9,0.95)), with global-norm gradient clipping (config.py:126 gradient_clipping) folded in
// through a device-side squared-norm, and the bf16 compute copy of the weights refreshed in the same pass.
// HBM-bound: 16 B read + 14 B written per parameter.
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    s += v * v;
  }
  const float t = block_sum<256>(s, red);
  if (threadIdx.x == 0) atomicAdd(out, t);
}

__global__ void adamw_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m1,
                             float* __restrict__ m2, bf16* __restrict__ shadow, long long n4, float lr, float b1,
                             float b2, float eps, float wd, float grad_scale, const float* __restrict__ gnorm_sq,
                             float max_norm, float bc1, float bc2, int zero_grad) {
  float coef = grad_scale;
  if (gnorm_sq != nullptr && max_norm > 0.f) {
    const float nrm = sqrtf(*gnorm_sq) * grad_scale;
    coef *= fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  const float inv_sqrt_bc2 = rsqrtf(bc2), step = lr / bc1, decay = 1.f - lr * wd;
  float4* w4 = reinterpret_cast<float4*>(w);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* a4 = reinterpret_cast<float4*>(m1);
  float4* v4 = reinterpret_cast<float4*>(m2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 wv = w4[i], gv = g4[i], av = a4[i], vv = v4[i];
    float* wp = &wv.x;
    float* gp = &gv.x;
    float* ap = &av.x;
    float* vp = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gi = gp[e] * coef;
      float wi = wp[e] * decay;
      const float a = b1 * ap[e] + (1.f - b1) * gi;
      const float v = b2 * vp[e] + (1.f - b2) * gi * gi;
      ap[e] = a;
      vp[e] = v;
      wi -= step * (a / (sqrtf(v) * inv_sqrt_bc2 + eps));
      wp[e] = wi;
    }
    w4[i] = wv;
    a4[i] = av;
    v4[i] = vv;
    if (shadow) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(wv.x, wv.y), h1 = __floats2bfloat162_rn(wv.z, wv.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
    if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}

#include "train_kernels.cuh"

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mb200

using namespace mb200;

// =============================================================================================
// C ABI
// =============================================================================================
#define MB_ENTER()                 \
  do {                             \
    int _rc = mb200::check_arch(); \
    if (_rc) return _rc;           \
  } while (0)
#define MB_LAUNCH_CHECK()           \
  do {                              \
    mb200::count_launch();          \
    MB_CUDA(cudaGetLastError());    \
  } while (0)
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mb200_layernorm_fwd(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y,
                                   int64_t ldy, float* mean, float* rstd, int32_t rows, int32_t d, float eps,
                                   void* stream) {
  MB_ENTER();
  MB_REQUIRE(rows > 0 && d > 0 && d % 8 == 0 && d <= kLnThreads * kLnMaxVec * 8, MB200_E_SHAPE,
             "layernorm: d=%d must be a multiple of 8 and <= %d", d, kLnThreads * kLnMaxVec * 8);
  MB_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0, MB200_E_ALIGN, "layernorm: row strides must be multiples of 8");
  MB_CUDA(launch_pdl(layernorm_fwd_kernel, dim3(rows), dim3(kLnThreads), 0, ST(stream), (const bf16*)x, (long long)ldx,
                     (const bf16*)gamma, (const bf16*)beta, (bf16*)y, (long long)ldy, mean, rstd, (int)d, eps));
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const void* gamma,
                                   const float* mean, const float* rstd, const void* res, int64_t ldres, void* dx,
                                   int64_t lddx, int32_t rows, int32_t d, void* stream) {
  MB_ENTER();
  MB_REQUIRE(rows > 0 && d > 0 && d % 8 == 0 && d <= kLnThreads * kLnMaxVec * 8, MB200_E_SHAPE,
             "layernorm_bwd: bad d=%d", d);
  MB_REQUIRE(lddy % 8 == 0 && ldx % 8 == 0 && lddx % 8 == 0 && (!res || ldres % 8 == 0), MB200_E_ALIGN,
             "layernorm_bwd: row strides must be multiples of 8");
  MB_CUDA(launch_pdl(layernorm_bwd_kernel, dim3(rows), dim3(kLnThreads), 0, ST(stream), (const bf16*)dy,
                     (long long)lddy, (const bf16*)x, (long long)ldx, (const bf16*)gamma, mean, rstd, (const bf16*)res,
                     (long long)ldres, (bf16*)dx, (long long)lddx, (int)d));
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_layernorm_param_grad(const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                          const float* mean, const float* rstd, float* dgamma, float* dbeta,
                                          int32_t rows, int32_t d, int32_t accumulate, void* stream) {
  MB_ENTER();
  layernorm_param_grad_kernel<<<(d + 127) / 128, 128, 0, ST(stream)>>>((const bf16*)dy, lddy, (const bf16*)x, ldx,
                                                                       mean, rstd, dgamma, dbeta, rows, d, accumulate);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_layernorm_param_grad_rows(const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                               const float* mean, const float* rstd, float* dgamma, float* dbeta,
                                               int32_t rows, int32_t d, int32_t accumulate, void* stream) {
  MB_ENTER();
  MB_REQUIRE(rows > 0 && d > 0 && d % 2 == 0 && lddy % 2 == 0 && ldx % 2 == 0, MB200_E_ALIGN,
             "layernorm_param_grad_rows: d and row strides must be even");
  {
    // many rows: 2-D decomposition with one atomic per (column, row chunk)
    if (!accumulate) {
      MB_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)d * sizeof(float), ST(stream)));
      MB_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)d * sizeof(float), ST(stream)));
    }
    dim3 grid((d + 63) / 64, (rows + kLnPgRows - 1) / kLnPgRows);
    layernorm_param_grad_rows_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)dy, lddy, (const bf16*)x, ldx, mean, rstd,
                                                                   dgamma, dbeta, rows, d);
    MB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int mb200_col_moments(const void* u, int64_t ldu, const void* v, int64_t ldv, const void* mask, int64_t ldm,
                                 int32_t rows, int32_t cols, float* out1, float* out2, void* stream) {
  MB_ENTER();
  MB_REQUIRE(rows > 0 && cols > 0 && cols % 2 == 0 && ldu % 2 == 0 && ldv % 2 == 0 && (!mask || ldm % 2 == 0),
             MB200_E_ALIGN, "col_moments: cols and row strides must be even");
  MB_CUDA(cudaMemsetAsync(out1, 0, (size_t)cols * sizeof(float), ST(stream)));
  MB_CUDA(cudaMemsetAsync(out2, 0, (size_t)cols * sizeof(float), ST(stream)));
  dim3 grid((cols + 63) / 64, (rows + kMomRows - 1) / kMomRows);
  col_moments_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)u, ldu, (const bf16*)v, ldv, (const bf16*)mask, ldm, rows,
                                                   cols, out1, out2);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_channel_affine(const void* x1, const float* a1, const void* x2, const float* a2, const float* c0,
                                    const void* mask, const void* res, int32_t relu, void* y, int64_t rows, int32_t C,
                                    void* stream) {
  MB_ENTER();
  MB_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && x1 && a1 && y && (!x2 || a2), MB200_E_ARG,
             "channel_affine: C must be a multiple of 8, x1 / a1 / y non-null, a2 given with x2");
  MB_REQUIRE(((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(mask) |
               reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, MB200_E_ALIGN,
             "channel_affine: pointers must be 16-byte aligned");
  channel_affine_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, ST(stream)>>>(
      (const bf16*)x1, a1, (const bf16*)x2, a2, c0, (const bf16*)mask, (const bf16*)res, relu, (bf16*)y, rows, C);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_col2im3x3(const void* dcols, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride,
                               void* stream) {
  MB_ENTER();
  MB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2), MB200_E_SHAPE,
             "col2im3x3: B=%d H=%d W=%d C=%d (multiple of 8) stride=%d (1 or 2)", B, H, W, C, stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  col2im3x3_kernel<<<grid_for((long long)B * H * W * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)dcols, (bf16*)dx, B,
                                                                                          H, W, C, stride, Ho, Wo);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_avgpool_nhwc_bwd(const void* dy, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k,
                                      void* stream) {
  MB_ENTER();
  MB_REQUIRE(B > 0 && k >= 1 && H >= k && W >= k && C > 0 && C % 8 == 0, MB200_E_SHAPE,
             "avgpool_nhwc_bwd: B=%d H=%d W=%d C=%d k=%d", B, H, W, C, k);
  avgpool_nhwc_bwd_kernel<<<grid_for((long long)B * H * W * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)dy, (bf16*)dx,
                                                                                                B, H, W, C, k);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_scale_add(const void* u, const float* s, const void* r1, const void* r2, void* out, int64_t n,
                               void* stream) {
  MB_ENTER();
  MB_REQUIRE(n > 0 && n % 8 == 0, MB200_E_SHAPE, "scale_add: n=%lld must be a positive multiple of 8", (long long)n);
  MB_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(r1) | reinterpret_cast<uintptr_t>(r2) |
               reinterpret_cast<uintptr_t>(out)) & 15) == 0, MB200_E_ALIGN, "scale_add: pointers must be 16-byte aligned");
  scale_add_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)u, s, (const bf16*)r1, (const bf16*)r2,
                                                                 (bf16*)out, n / 8);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_dot(const void* a, const void* b, int64_t n, float* out, int32_t accumulate, void* stream) {
  MB_ENTER();
  MB_REQUIRE(n > 0 && n % 8 == 0, MB200_E_SHAPE, "dot: n=%lld must be a positive multiple of 8", (long long)n);
  MB_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0, MB200_E_ALIGN,
             "dot: pointers must be 16-byte aligned");
  if (!accumulate) MB_CUDA(cudaMemsetAsync(out, 0, sizeof(float), ST(stream)));
  const int grid = grid_for(n / 8, 256) > 1024 ? 1024 : grid_for(n / 8, 256);
  dot_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, n / 8, out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_quick_gelu_bwd(const void* dy, const void* pre, void* dx, int64_t n, void* stream) {
  MB_ENTER();
  MB_REQUIRE(n > 0 && n % 8 == 0, MB200_E_SHAPE, "quick_gelu_bwd: n=%lld must be a positive multiple of 8", (long long)n);
  MB_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
             MB200_E_ALIGN, "quick_gelu_bwd: pointers must be 16-byte aligned");
  quick_gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)dy, (const bf16*)pre, (bf16*)dx, n / 8);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_rope(void* qkv, int64_t ld, int32_t rows, int32_t S, int32_t H, int32_t hd, int32_t rot,
                          int32_t pos0, int32_t inverse, void* stream) {
  MB_ENTER();
  MB_REQUIRE(rot % 2 == 0 && rot <= hd && rows > 0, MB200_E_SHAPE, "rope: bad rot=%d hd=%d", rot, hd);
  const long long total = (long long)rows * 2 * H * (rot / 2);
  rope_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>((bf16*)qkv, ld, rows, S, H, hd, rot, pos0, inverse);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_rope_table(float* tab, int32_t S, int32_t rot, int32_t pos0, void* stream) {
  MB_ENTER();
  MB_REQUIRE(S > 0 && rot > 0 && rot % 2 == 0, MB200_E_SHAPE, "rope_table: bad S=%d rot=%d", S, rot);
  const int n = S * (rot / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, ST(stream)>>>(reinterpret_cast<float2*>(tab), S, rot / 2, rot, pos0);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_softmax_fwd(const float* s, int64_t lds, int64_t s_bs, void* p, int64_t ldp, int64_t p_bs,
                                 int32_t nz, int32_t Sq, int32_t Sk, float scale, int32_t causal, int32_t koff,
                                 void* stream) {
  MB_ENTER();
  const long long warps = (long long)nz * Sq;
  MB_CUDA(launch_pdl(softmax_fwd_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, ST(stream), s, (long long)lds,
                     (long long)s_bs, (bf16*)p, (long long)ldp, (long long)p_bs, (int)nz, (int)Sq, (int)Sk, scale,
                     (int)causal, (int)koff));
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_softmax_bwd(const float* dp, int64_t lddp, int64_t dp_bs, const void* p, int64_t ldp,
                                 int64_t p_bs, void* ds, int64_t ldds, int64_t ds_bs, int32_t nz, int32_t Sq,
                                 int32_t Sk, float scale, void* stream) {
  MB_ENTER();
  const long long warps = (long long)nz * Sq;
  MB_CUDA(launch_pdl(softmax_bwd_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, ST(stream), dp,
                     (long long)lddp, (long long)dp_bs, (const bf16*)p, (long long)ldp, (long long)p_bs, (bf16*)ds,
                     (long long)ldds, (long long)ds_bs, (int)nz, (int)Sq, (int)Sk, scale));
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_build_labels(const int64_t* captions, int64_t ldc, int64_t* labels, int32_t B, int32_t S,
                                  int32_t L, int64_t eos, void* stream) {
  MB_ENTER();
  MB_REQUIRE(B > 0 && S > 0 && L >= 0 && L <= S, MB200_E_SHAPE, "build_labels: need 0 <= L=%d <= S=%d", L, S);
  build_labels_kernel<<<(B * 32 + 127) / 128, 128, 0, ST(stream)>>>((const long long*)captions, ldc,
                                                                    (long long*)labels, B, S, L, eos);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_embed_assemble(const int64_t* captions, int64_t ldc, const void* wte, const void* prefix,
                                    int32_t L, void* x, int32_t B, int32_t S, int32_t d, int32_t vocab,
                                    void* stream) {
  MB_ENTER();
  MB_REQUIRE(d % 8 == 0 && L >= 0 && L <= S, MB200_E_SHAPE, "embed_assemble: bad d=%d L=%d S=%d", d, L, S);
  embed_assemble_kernel<<<B * S, 256, 0, ST(stream)>>>((const long long*)captions, ldc, (const bf16*)wte,
                                                       (const bf16*)prefix, L, (bf16*)x, B, S, d, vocab);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_embed_gather(const int64_t* ids, const void* wte, void* out, int32_t n, int32_t d,
                                  int32_t vocab, void* stream) {
  MB_ENTER();
  MB_REQUIRE(d % 8 == 0 && n > 0, MB200_E_SHAPE, "embed_gather: bad n=%d d=%d", n, d);
  embed_gather_kernel<<<n, 256, 0, ST(stream)>>>((const long long*)ids, (const bf16*)wte, (bf16*)out, d, vocab);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_cross_entropy(const void* logits, int64_t ldv, const int64_t* labels, int32_t B, int32_t S,
                                   int32_t V, float* row_loss, int32_t* n_valid, float* loss, void* dlogits,
                                   float grad_scale, void* stream) {
  MB_ENTER();
  MB_REQUIRE(ldv % 8 == 0 && V <= ldv, MB200_E_ALIGN, "cross_entropy: ldv=%lld must be a multiple of 8 and >= V",
             (long long)ldv);
  ce_count_kernel<<<1, 1024, 0, ST(stream)>>>((const long long*)labels, B, S, n_valid);
  ce_row_kernel<<<B * S, kCeThreads, 0, ST(stream)>>>((const bf16*)logits, ldv, (const long long*)labels, S, V,
                                                      n_valid, row_loss, (bf16*)dlogits, grad_scale);
  ce_reduce_kernel<<<1, 1024, 0, ST(stream)>>>(row_loss, B * S, n_valid, loss);
  mb200::count_launch(2);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_colsum(const void* x, int64_t ldx, int32_t rows, int32_t cols, float* out, int32_t accumulate,
                            void* stream) {
  MB_ENTER();
  MB_REQUIRE(cols % 2 == 0 && ldx % 2 == 0, MB200_E_ALIGN, "colsum: cols and ldx must be even");
  if (!accumulate) MB_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), ST(stream)));
  dim3 grid((cols + 63) / 64, (rows + kColsumRows - 1) / kColsumRows);
  MB_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, ST(stream), (const bf16*)x, (long long)ldx, (int)rows, (int)cols,
                     out));
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_dropout_fwd(const void* x, void* y, uint8_t* mask, int64_t n, float p, uint64_t seed,
                                 void* stream) {
  MB_ENTER();
  MB_REQUIRE(p >= 0.f && p < 1.f, MB200_E_ARG, "dropout: p=%f out of range", p);
  dropout_fwd_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y, mask, n, p, seed);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_dropout_apply(const void* x, const uint8_t* mask, void* y, int64_t n, float p, void* stream) {
  MB_ENTER();
  dropout_apply_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>((const bf16*)x, mask, (bf16*)y, n, p);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_patchify(const void* img, void* patches, int64_t ldp, int32_t B, int32_t R, int32_t P,
                              void* stream) {
  MB_ENTER();
  MB_REQUIRE(R % P == 0 && ldp >= 3 * P * P, MB200_E_SHAPE, "patchify: R=%d P=%d ldp=%lld", R, P, (long long)ldp);
  const long long total = (long long)B * (R / P) * (R / P) * 3 * P * P;
  patchify_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)img, (bf16*)patches, ldp, B, R, P);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_vit_assemble(void* x, const void* pe, const void* cls, const void* pos, int32_t B, int32_t T,
                                  int32_t w, void* stream) {
  MB_ENTER();
  vit_assemble_kernel<<<grid_for((long long)B * T * w, 256), 256, 0, ST(stream)>>>(
      (bf16*)x, (const bf16*)pe, (const bf16*)cls, (const bf16*)pos, B, T, w);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_nchw_to_nhwc8(const void* src, void* dst, int32_t B, int32_t C, int32_t H, int32_t W,
                                   void* stream) {
  MB_ENTER();
  MB_REQUIRE(C >= 1 && C <= 8 && B > 0 && H > 0 && W > 0, MB200_E_SHAPE, "nchw_to_nhwc8: B=%d C=%d H=%d W=%d", B, C, H, W);
  nchw_to_nhwc8_kernel<<<grid_for((long long)B * H * W, 256), 256, 0, ST(stream)>>>((const bf16*)src, (bf16*)dst, B, C,
                                                                                   H, W);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_im2col3x3(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride,
                                void* stream) {
  MB_ENTER();
  MB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2), MB200_E_SHAPE,
             "im2col3x3: B=%d H=%d W=%d C=%d (multiple of 8) stride=%d (1 or 2)", B, H, W, C, stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long nvec = (long long)B * Ho * Wo * 9 * (C / 8);
  if (nvec < (1LL << 31) - (1LL << 24))  // headroom for the grid-stride increment
    im2col3x3_kernel<unsigned int><<<grid_for(nvec, 256), 256, 0, ST(stream)>>>((const bf16*)src, (bf16*)dst, B, H, W, C,
                                                                                stride, Ho, Wo);
  else
    im2col3x3_kernel<long long><<<grid_for(nvec, 256), 256, 0, ST(stream)>>>((const bf16*)src, (bf16*)dst, B, H, W, C,
                                                                             stride, Ho, Wo);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_avgpool_nhwc(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, int32_t k,
                                   void* stream) {
  MB_ENTER();
  MB_REQUIRE(B > 0 && k >= 1 && H >= k && W >= k && C > 0 && C % 8 == 0, MB200_E_SHAPE,
             "avgpool_nhwc: B=%d H=%d W=%d C=%d (multiple of 8) k=%d", B, H, W, C, k);
  avgpool_nhwc_kernel<<<grid_for((long long)B * (H / k) * (W / k) * (C / 8), 256), 256, 0, ST(stream)>>>(
      (const bf16*)src, (bf16*)dst, B, H, W, C, k);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_argmax(const void* x, int64_t ldx, int32_t rows, int32_t V, int64_t* out, void* stream) {
  MB_ENTER();
  argmax_kernel<<<rows, 512, 0, ST(stream)>>>((const bf16*)x, ldx, V, (long long*)out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_add(const void* a, const void* b, const void* c, void* y, int64_t n, void* stream) {
  MB_ENTER();
  MB_REQUIRE(n % 8 == 0, MB200_E_ALIGN, "add: n must be a multiple of 8");
  add_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, (const bf16*)c, (bf16*)y,
                                                           n / 8);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_sumsq(const float* x, int64_t n, float* out, void* stream) {
  MB_ENTER();
  sumsq_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, n, out);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_adamw_step(float* master, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                                int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                float grad_scale, const float* gnorm_sq, float max_norm, int32_t step,
                                int32_t zero_grad, void* stream) {
  MB_ENTER();
  MB_REQUIRE(step >= 1, MB200_E_ARG, "adamw: step must be >= 1");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  MB_REQUIRE(n % 4 == 0, MB200_E_ALIGN, "adamw: n must be a multiple of 4 (the arena pads every tensor to 64)");
  adamw_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(master, grad, exp_avg, exp_avg_sq, (bf16*)shadow_bf16, n / 4, lr,
                                                         beta1, beta2, eps, weight_decay, grad_scale, gnorm_sq,
                                                         max_norm, bc1, bc2, zero_grad);
  MB_LAUNCH_CHECK();
  return 0;
}

extern "C" int mb200_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  MB_ENTER();
  cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(src, (bf16*)dst, n);
  MB_LAUNCH_CHECK();
  return 0;
}
extern "C" int mb200_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  MB_ENTER();
  cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>((const bf16*)src, dst, n);
  MB_LAUNCH_CHECK();
  return 0;
}
