// magma_b200 — pieces shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 GEMM kernels: tile configuration,
// kernel parameter block, UMMA shared-memory descriptor, and the fused epilogue.
#pragma once
#include "common.cuh"

namespace mb200 {

static constexpr int BM = 128;       // UMMA M (cta_group::1)
static constexpr int BK = 64;        // 64 bf16 = 128 bytes = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;    // fixed for 16-bit inputs
static constexpr int kThreads = 256; // 8 warps
static constexpr int kSmemBudget = 192 * 1024;  // operand ring; + 34 KB epilogue staging + barriers < 227 KB

// AROWS = rows of A actually staged per k-block. 128 normally. 32 for small-M (decode) problems: only a 32-row TMA box
// is loaded per stage and the A slots are packed 4 KB apart; the UMMA descriptor still spans 128 rows (16 KB), so
// rows 32..127 of the product are computed from whatever follows in shared memory — those TMEM lanes are never stored.
// The freed shared memory goes to deeper rings: what bounds a weight-streaming GEMM is bytes of B in flight per SM.
//
// In that mode a ring stage also carries kKS consecutive k-blocks (4 at BN = 64, 2 at BN = 128): the single MMA-issuing
// thread spends ~350 cycles per barrier round trip (try_wait + fence + 4 tcgen05.mma + commit), which hides under the
// 512 tensor-pipe cycles of a 128x256 k-block but not under the 128 cycles of a 128x64 one.
template <int BN, int AROWS = BM>
struct Cfg {
  static constexpr int kKS = AROWS == BM ? 1 : (BN == 64 ? 4 : (BN == 128 ? 2 : 1));  // k-blocks per ring stage
  static constexpr int kASub = AROWS * BK * 2;    // bytes of one k-block of A (sub-slot stride)
  static constexpr int kBSub = BN * BK * 2;
  static constexpr int kABytes = kASub * kKS;     // slot stride of the A ring
  static constexpr int kBBytes = kBSub * kKS;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kAWindow = BM * BK * 2;    // bytes an MMA reads starting at an A slot
  static constexpr int kMaxStages = AROWS == BM ? 8 : 16;
  static constexpr int kStagesRaw = (kSmemBudget - (kAWindow - kASub)) / kStageBytes;
  static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers; 128/256/512 — powers of two
  static constexpr int kEpiPitch = 64;      // floats per staged row; 16-byte chunks XOR-swizzled by (row & 15)
  static constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;  // per-warp [32 rows][64 cols] fp32 staging, 4 warps
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct GemmKernelParams {
  int M, N, K;
  int nb0;
  int tiles_m, tiles_n, total_tiles;
  void* C;
  long long ldc, c_bs0, c_bs1;
  float alpha;
  int act, dact, accumulate;
  const bf16* bias;
  bf16* aux_out;
  const bf16* aux_in;
  const bf16* res1;
  const bf16* res2;
  long long ld_res;
  // fused rotary embedding (rotate_every_two) on column pairs: applied when rope_mode != 0
  const float2* rope_tab;  // [rope_S][rope_rot/2] (cos, sin) of the position of row (row % rope_S)
  int rope_mode;           // +1 forward, -1 inverse (transpose rotation)
  int rope_S, rope_hd, rope_rot, rope_ncols;
  int epi_kind;  // EK_*: which specialised epilogue handles full float4 column groups (0 = generic only)
  // split-K (small-M / weight-streaming GEMMs, e.g. decode): work item = (tile, k-range); partial tiles are reduced
  // written to splitk_ws [split][M][ld_ws] (fp32) and summed in fixed order, with the fused epilogue, by
  // splitk_finalize_kernel
  int split_k, kb_per_split;
  float* splitk_ws;
  long long ld_ws;
};

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, SWIZZLE_128B, version 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}


// ---------------------------------------------------------------------------------------------
// Epilogue. Accumulator rows live one-per-thread in TMEM; outputs (and residual / aux inputs) want row-contiguous
// global access. Per 64-column group: phase 1 moves TMEM -> registers -> a per-warp XOR-swizzled smem block, phase 2
// re-reads it so that 16 lanes x 4 columns cover one row segment (full 128-byte lines per row).
// The four epilogue warps run ONE warp per scheduler, so phase 2 is latency-bound unless it is straight-line with
// high ILP: epi_tile<> is specialised at compile time on what the epilogue does, fully unrolled over the 16 row pairs,
// uses pointer bumps instead of per-row address arithmetic, and issues ALL global loads of a group (residuals, aux)
// before anything is stored (a store may alias a later load as far as the compiler knows, which would serialise one
// DRAM round trip per row pair). The kernel dispatches once per tile on p.epi_kind (chosen on the host).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void bf16x4_to_f32(const uint2& u, float (&f)[4]) {
  const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y;
}
__device__ __forceinline__ uint2 f32x4_to_bf16(const float (&v)[4]) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  return u;
}
__device__ __forceinline__ uint2 ldg64(const bf16* ptr) {
  uint2 u;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "l"(ptr));
  return u;
}

enum { EK_GENERIC = 0, EK_PLAIN, EK_ROPE, EK_GELU, EK_GELU_AUX, EK_QGELU, EK_RELU, EK_DGELU, EK_DRELU, EK_RES1,
       EK_RES2, EK_ACCUM, EK_SPLITK };

struct EpiCtx {
  uint32_t tmem_acc;  // TMEM address of this warp's lanes, column 0 of the accumulator buffer
  uint32_t stg_s;     // smem address of this warp's staging block
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
  int empty_remote;   // 0: arrive on the local tmem_empty barrier; 1: arrive on the leader CTA's (2-CTA kernel, peer CTA)
  uint32_t full_phase;
  long long boff;     // batch offset in C (elements)
  int row0, nrows, n_blk, lane;
  int ks;             // split-K index of this work item (0 when K is not split)
};

// slow per-lane path: any combination of epilogue options, any number (1..4) of valid columns. Used for float4
// groups cut by the N edge and for option combinations without a specialised instantiation.
template <typename OutT>
__device__ __noinline__ void epi_lane_generic(const GemmKernelParams& p, const EpiCtx& c, int col, int nvalid, int rsub,
                                              int c4) {
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < nvalid) bv[e] = __bfloat162float(p.bias[col + e]);
  }
  int rope_p = -1;
  if (p.rope_mode != 0 && col < p.rope_ncols) {
    const int dim = col % p.rope_hd;
    if (dim < p.rope_rot) rope_p = dim >> 1;
  }
  for (int it = 0; it < 16; ++it) {
    const int rl = it * 2 + rsub;
    if (rl >= c.nrows) continue;
    const int row = c.row0 + rl;
    const float4 sv = lds128(c.stg_s + (uint32_t)(rl * 256 + ((((c4 >> 2)) ^ (rl & 15)) << 4)));
    float v[4] = {sv.x + bv[0], sv.y + bv[1], sv.z + bv[2], sv.w + bv[3]};
    const long long coff = c.boff + (long long)row * p.ldc + col;
    if (rope_p >= 0) {  // pairs never straddle the float4 (col % 4 == 0, rot % 4 == 0)
      const float2* tp = p.rope_tab + (long long)(row % p.rope_S) * (p.rope_rot >> 1) + rope_p;
      const float2 cs0 = __ldg(tp), cs1 = __ldg(tp + 1);
      const float sg = p.rope_mode > 0 ? 1.f : -1.f;
      const float a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
      v[0] = a0 * cs0.x - a1 * cs0.y * sg;
      v[1] = a1 * cs0.x + a0 * cs0.y * sg;
      v[2] = a2 * cs1.x - a3 * cs1.y * sg;
      v[3] = a3 * cs1.x + a2 * cs1.y * sg;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (e >= nvalid) continue;
      float x = v[e];
      if (p.aux_out) p.aux_out[coff + e] = __float2bfloat16(x);
      if (p.act == MB200_ACT_GELU_NEW) x = gelu_new_f(x);
      else if (p.act == MB200_ACT_QUICK_GELU) x = quick_gelu_f(x);
      else if (p.act == MB200_ACT_RELU) x = fmaxf(x, 0.f);
      if (p.dact) {
        const float a = __bfloat162float(p.aux_in[coff + e]);
        x = p.dact == MB200_DACT_GELU_NEW ? x * gelu_new_grad_f(a) : (a > 0.f ? x : 0.f);
      }
      const long long roff = c.boff + (long long)row * p.ld_res + col + e;
      if (p.res1) x += __bfloat162float(p.res1[roff]);
      if (p.res2) x += __bfloat162float(p.res2[roff]);
      if (p.act == MB200_ACT_RELU_POST) x = fmaxf(x, 0.f);
      if constexpr (sizeof(OutT) == 4) {
        float* dst = reinterpret_cast<float*>(p.C) + coff + e;
        *dst = p.accumulate ? *dst + x : x;
      } else {
        reinterpret_cast<bf16*>(p.C)[coff + e] = __float2bfloat16(x);
      }
    }
  }
}

template <int BN, int DACT, int NRES, bool ACCUM, bool GENERIC, int NLD>
__device__ __forceinline__ void epi_preload(const GemmKernelParams& p, const EpiCtx& c, int g, int c4, int rsub,
                                            long long cstep, long long rstep, uint2 (&pre)[NLD > 0 ? NLD : 1][16]) {
  if constexpr (NLD > 0 && !GENERIC) {
    const int col = c.n_blk * BN + g * 64 + c4;
    if (col + 4 > p.N) return;
    const long long aoff = c.boff + (long long)(c.row0 + rsub) * p.ldc + col;
    const long long roff = c.boff + (long long)(c.row0 + rsub) * p.ld_res + col;
    constexpr int kD = 0, kR1 = (DACT != 0 ? 1 : 0), kR2 = kR1 + 1, kA = kR1 + NRES;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      if (it * 2 + rsub < c.nrows) {
        if constexpr (DACT != 0) pre[kD][it] = ldg64(p.aux_in + aoff + it * cstep);
        if constexpr (NRES >= 1) pre[kR1][it] = ldg64(p.res1 + roff + it * rstep);
        if constexpr (NRES >= 2) pre[kR2][it] = ldg64(p.res2 + roff + it * rstep);
        if constexpr (ACCUM) {
          const float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.C) + aoff + it * cstep);
          pre[kA][it] = make_uint2(__float_as_uint(o.x), __float_as_uint(o.y));
          pre[kA + 1][it] = make_uint2(__float_as_uint(o.z), __float_as_uint(o.w));
        }
      }
    }
  }
}

template <int BN, int ACT, int DACT, int NRES, bool AUX, bool ROPE, bool ACCUM, bool GENERIC, typename OutT>
__device__ __forceinline__ void epi_tile(const GemmKernelParams& p, const EpiCtx& c) {
  constexpr int NLD = (DACT != 0 ? 1 : 0) + NRES + (ACCUM ? 2 : 0);  // uint2 loads per row pair
  const int lane = c.lane;
  const int c4 = (lane & 15) * 4;
  const int rsub = lane >> 4;
  const int n_tile_end = min(p.N, (c.n_blk + 1) * BN);
  const long long cstep = 2 * p.ldc, rstep = 2 * p.ld_res;
  uint2 pre[NLD > 0 ? NLD : 1][16];

  epi_preload<BN, DACT, NRES, ACCUM, GENERIC, NLD>(p, c, 0, c4, rsub, cstep, rstep, pre);  // overlaps the accumulator wait
  mbar_wait(c.tmem_full, c.full_phase);
  tc_fence_after();

#pragma unroll 1
  for (int g = 0; g < BN / 64; ++g) {
    const int n0 = c.n_blk * BN + g * 64;
    if (n0 >= p.N) break;  // warp-uniform
    // ---- phase 1: TMEM -> registers -> per-warp smem staging (thread i owns accumulator row i) ----
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t rr[32];
      tmem_ld_32x32(c.tmem_acc + (uint32_t)(g * 64 + h * 32), rr);
      tmem_ld_wait();
      const uint32_t drow = c.stg_s + (uint32_t)(lane * 256);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        sts128(drow + (uint32_t)(((((h * 32 + j) >> 2) ^ (lane & 15)) << 4)), __uint_as_float(rr[j]) * p.alpha,
               __uint_as_float(rr[j + 1]) * p.alpha, __uint_as_float(rr[j + 2]) * p.alpha,
               __uint_as_float(rr[j + 3]) * p.alpha);
    }
    if (n0 + 64 >= n_tile_end) {
      // last column group: the accumulator has been fully read -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (c.empty_remote) mbar_arrive_remote(c.tmem_empty, 0);
        else mbar_arrive(c.tmem_empty);
      }
    }
    __syncwarp();
    // ---- phase 2 ----
    const int col = n0 + c4;
    const int nvalid = p.N - col;
    if (p.epi_kind == EK_SPLITK) {
      // partial tile of a split-K work item: plain fp32 stores into this split's slice of the workspace (deterministic:
      // the finalize kernel sums the slices in a fixed order and applies the fused epilogue)
      if (nvalid > 0) {
        float* wrow = p.splitk_ws + ((long long)c.ks * p.M + (c.row0 + rsub)) * p.ld_ws + col;
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
          const int rl = it * 2 + rsub;
          if (rl < c.nrows) {
            const float4 sv = lds128(c.stg_s + (uint32_t)(rl * 256 + ((((c4 >> 2)) ^ (rl & 15)) << 4)));
            *reinterpret_cast<float4*>(wrow + (long long)it * 2 * p.ld_ws) = sv;  // ld_ws % 4 == 0: padded columns exist
          }
        }
      }
    } else if (nvalid >= 4 && !GENERIC) {
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.bias) bf16x4_to_f32(ldg64(p.bias + col), bv);
      int rope_p = -1;
      if constexpr (ROPE) {
        if (col < p.rope_ncols) {
          const int dim = col % p.rope_hd;
          if (dim < p.rope_rot) rope_p = dim >> 1;
        }
      }
      const long long aoff = c.boff + (long long)(c.row0 + rsub) * p.ldc + col;
      OutT* cptr = reinterpret_cast<OutT*>(p.C) + aoff;
      bf16* auxo = AUX ? p.aux_out + aoff : nullptr;
      constexpr int kFence = (ACT == MB200_ACT_GELU_NEW || ACT == MB200_ACT_QUICK_GELU || DACT == MB200_DACT_GELU_NEW) ? 1 : 4;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        // compiler scheduling fence: keeps the fully unrolled body from hoisting every smem read / conversion of all 16
        // row pairs at once (register pressure); the global loads were all issued by epi_preload already
        if (it % kFence == 0) asm volatile("" ::: "memory");
        const int rl = it * 2 + rsub;
        if (rl < c.nrows) {
          const float4 sv = lds128(c.stg_s + (uint32_t)(rl * 256 + ((((c4 >> 2)) ^ (rl & 15)) << 4)));
          float v[4] = {sv.x + bv[0], sv.y + bv[1], sv.z + bv[2], sv.w + bv[3]};
          if constexpr (ROPE) {
            if (rope_p >= 0) {
              const float2* tp = p.rope_tab + (long long)((c.row0 + rl) % p.rope_S) * (p.rope_rot >> 1) + rope_p;
              const float2 cs0 = __ldg(tp), cs1 = __ldg(tp + 1);
              const float sg = p.rope_mode > 0 ? 1.f : -1.f;
              const float a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
              v[0] = a0 * cs0.x - a1 * cs0.y * sg;
              v[1] = a1 * cs0.x + a0 * cs0.y * sg;
              v[2] = a2 * cs1.x - a3 * cs1.y * sg;
              v[3] = a3 * cs1.x + a2 * cs1.y * sg;
            }
          }
          if constexpr (AUX) *reinterpret_cast<uint2*>(auxo + it * cstep) = f32x4_to_bf16(v);
          if constexpr (ACT == MB200_ACT_GELU_NEW) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = gelu_new_f(v[e]);
          } else if constexpr (ACT == MB200_ACT_QUICK_GELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = quick_gelu_f(v[e]);
          } else if constexpr (ACT == MB200_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          constexpr int kD = 0, kR1 = (DACT != 0 ? 1 : 0), kR2 = kR1 + 1, kA = kR1 + NRES;
          if constexpr (DACT != 0) {
            float a[4];
            bf16x4_to_f32(pre[kD][it], a);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (DACT == MB200_DACT_GELU_NEW) v[e] *= gelu_new_grad_f(a[e]);
              else v[e] = a[e] > 0.f ? v[e] : 0.f;
            }
          }
          if constexpr (NRES >= 1) {
            float a[4];
            bf16x4_to_f32(pre[kR1][it], a);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += a[e];
          }
          if constexpr (NRES >= 2) {
            float a[4];
            bf16x4_to_f32(pre[kR2][it], a);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += a[e];
          }
          if constexpr (NRES >= 1) {
            if (p.act == MB200_ACT_RELU_POST) {  // warp-uniform
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
          }
          if constexpr (sizeof(OutT) == 4) {
            float4 o = make_float4(v[0], v[1], v[2], v[3]);
            if constexpr (ACCUM) {
              o.x += __uint_as_float(pre[kA][it].x);
              o.y += __uint_as_float(pre[kA][it].y);
              o.z += __uint_as_float(pre[kA + 1][it].x);
              o.w += __uint_as_float(pre[kA + 1][it].y);
            }
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(cptr) + it * cstep) = o;
          } else {
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(cptr) + it * cstep) = f32x4_to_bf16(v);
          }
        }
      }
    } else if (GENERIC && nvalid > 0) {
      epi_lane_generic<OutT>(p, c, col, nvalid < 4 ? nvalid : 4, rsub, c4);
    }
    __syncwarp();  // staging block is reused by the next column group
    if (g + 1 < BN / 64)  // latency overlaps phase 1 of the next group
      epi_preload<BN, DACT, NRES, ACCUM, GENERIC, NLD>(p, c, g + 1, c4, rsub, cstep, rstep, pre);
  }
  if constexpr (!GENERIC) {
    if (p.epi_kind == EK_SPLITK) return;
    // a float4 column group cut by the N edge (N % 4 != 0, e.g. the 50258-wide LM head) can only be in the LAST group
    // of the tile, whose staging block is still intact: finish those lanes on the slow path, outside the hot loop.
    if (p.N & 3) {
      const int lastg = (n_tile_end - 1 - c.n_blk * BN) >> 6;
      const int col = c.n_blk * BN + lastg * 64 + c4;
      const int nvalid = p.N - col;
      if (nvalid > 0 && nvalid < 4) epi_lane_generic<OutT>(p, c, col, nvalid, rsub, c4);
    }
  }
}


// host helpers shared by gemm.cu / gemm2.cu
int make_operand_map(CUtensorMap* out, const mb200_operand& op, int rows, int K, int nb0, int nb1, int box_rows);
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKernelParams& kp, bool a_mn, bool b_mn,
                 bool f32, cudaStream_t stream);

}  // namespace mb200
