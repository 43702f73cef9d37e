// magma_b200 — bf16 GEMM core for sm_100a.
//
//   C[b][M,N] = epilogue(alpha * A[b][M,K] * B[b][N,K]^T)
//
// Design (B200-first, not a translation of anything in the reference, which only calls cuBLAS through
// torch.nn.Linear — e.g. magma/adapters.py:19-23, magma/image_prefix.py:72):
//   * persistent grid (<= one CTA per SM), static round-robin tile scheduler, m-fastest tile order so
//     CTAs that run concurrently share the same weight (B) tile through the 126 MB L2;
//   * warp-specialised: warp 0 = TMA producer (one lane), warp 1 = tcgen05.mma issuer (one lane),
//     warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> fused epilogue -> HBM);
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage smem ring,
//     completion tracked with mbarriers; smem slots are released by tcgen05.commit;
//   * fp32 accumulators live in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1;
//   * both operand majors (K-major and MN-major) are supported through the UMMA shared-memory
//     descriptors, so dgrad (dY*W) and wgrad (dY^T*X) read the original tensors — no transposes.
#include "common.cuh"

#include <mutex>

namespace mb200 {

static constexpr int BM = 128;       // UMMA M (cta_group::1)
static constexpr int BK = 64;        // 64 bf16 = 128 bytes = one SWIZZLE_128B row
static constexpr int UMMA_K = 16;    // fixed for 16-bit inputs
static constexpr int kThreads = 256; // 8 warps
static constexpr int kSmemBudget = 200 * 1024;

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers; 128/256/512 — powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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

template <int BN, bool A_MN, bool B_MN, typename OutT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmKernelParams p) {
  using C_ = Cfg<BN>;
  constexpr int kStages = C_::kStages;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * C_::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C_::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc<C_::kTmemCols>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int tpb = p.tiles_m * p.tiles_n;
        const int z = t / tpb;
        const int r = t - z * tpb;
        const int m_blk = r % p.tiles_m;
        const int n_blk = r / p.tiles_m;
        const int z0 = z % p.nb0, z1 = z / p.nb0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], C_::kStageBytes);
          uint8_t* sa = smem_a + stage * C_::kABytes;
          uint8_t* sb = smem_b + stage * C_::kBBytes;
          if constexpr (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_4d(sa + i * 8192, &tmA, &full_bar[stage], m_blk * BM + i * 64, kb * BK, z0, z1);
          } else {
            tma_load_4d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM, z0, z1);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_4d(sb + i * 8192, &tmB, &full_bar[stage], n_blk * BN + i * 64, kb * BK, z0, z1);
          } else {
            tma_load_4d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN, z0, z1);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // K-major SW128: 8-row atoms of 128 B rows -> SBO = 1024, LBO unused; advance K by 16 elems = 32 B.
      // MN-major SW128: atom = 64 MN-elements (128 B) x 8 k-rows; SBO = 1024 between k-groups of 8,
      //                 LBO = 8192 between 64-wide MN chunks (one TMA box each); advance K by 16 rows = 2048 B.
      constexpr uint32_t a_lbo = A_MN ? 8192u : 0u, b_lbo = B_MN ? 8192u : 0u;
      constexpr uint32_t a_kadv = A_MN ? 2048u : 32u, b_kadv = B_MN ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * C_::kABytes);
          const uint32_t sb = smem_u32(smem_b + stage * C_::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_kadv, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + k * b_kadv, b_lbo, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int tpb = p.tiles_m * p.tiles_n;
      const int z = t / tpb;
      const int r = t - z * tpb;
      const int m_blk = r % p.tiles_m;
      const int n_blk = r / p.tiles_m;
      const int z0 = z % p.nb0, z1 = z / p.nb0;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();

      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const long long boff = (long long)z0 * p.c_bs0 + (long long)z1 * p.c_bs1;
      const long long coff = boff + (long long)row * p.ldc;
      const long long roff = boff + (long long)row * p.ld_res;

#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), rr);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) * p.alpha;
        const bool full = (n0 + 32 <= p.N);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || n0 + j < p.N) v[j] += __bfloat162float(__ldg(p.bias + n0 + j));
        }
        if (row_ok) {
          if (p.aux_out) {
            bf16* dst = p.aux_out + coff + n0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h0);
                u.y = *reinterpret_cast<uint32_t*>(&h1);
                u.z = *reinterpret_cast<uint32_t*>(&h2);
                u.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(dst + j) = u;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) dst[j] = __float2bfloat16(v[j]);
            }
          }
          if (p.act == MB200_ACT_GELU_NEW) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_new_f(v[j]);
          } else if (p.act == MB200_ACT_QUICK_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = quick_gelu_f(v[j]);
          } else if (p.act == MB200_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.dact) {
            const bf16* src = p.aux_in + coff + n0;
            float a[32];
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u = *reinterpret_cast<const uint4*>(src + j);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __bfloat1622float2(h[e]);
                  a[j + 2 * e] = f.x;
                  a[j + 2 * e + 1] = f.y;
                }
              }
            } else {
              for (int j = 0; j < 32; ++j) a[j] = (n0 + j < p.N) ? __bfloat162float(src[j]) : 0.f;
            }
            if (p.dact == MB200_DACT_GELU_NEW) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= gelu_new_grad_f(a[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = a[j] > 0.f ? v[j] : 0.f;
            }
          }
#pragma unroll 1
          for (int ri = 0; ri < 2; ++ri) {
            const bf16* rp = ri == 0 ? p.res1 : p.res2;
            if (!rp) continue;
            const bf16* src = rp + roff + n0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u = *reinterpret_cast<const uint4*>(src + j);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __bfloat1622float2(h[e]);
                  v[j + 2 * e] += f.x;
                  v[j + 2 * e + 1] += f.y;
                }
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) v[j] += __bfloat162float(src[j]);
            }
          }
          if constexpr (sizeof(OutT) == 4) {
            float* dst = reinterpret_cast<float*>(p.C) + coff + n0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (p.accumulate) {
                  float4 old = *reinterpret_cast<const float4*>(dst + j);
                  o.x += old.x;
                  o.y += old.y;
                  o.z += old.z;
                  o.w += old.w;
                }
                *reinterpret_cast<float4*>(dst + j) = o;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) dst[j] = p.accumulate ? dst[j] + v[j] : v[j];
            }
          } else {
            bf16* dst = reinterpret_cast<bf16*>(p.C) + coff + n0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t*>(&h0);
                u.y = *reinterpret_cast<uint32_t*>(&h1);
                u.z = *reinterpret_cast<uint32_t*>(&h2);
                u.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(dst + j) = u;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (n0 + j < p.N) dst[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above) -> release it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C_::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// rank-4 bf16 tensor map over an operand. K-major: dims (K, rows, nb0, nb1), box (64, box_rows).
// MN-major: dims (rows, K, nb0, nb1), box (64, 64).
static int make_operand_map(CUtensorMap* out, const mb200_operand& op, int rows, int K, int nb0, int nb1,
                            int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  MB_REQUIRE(enc != nullptr, MB200_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  MB_REQUIRE((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, MB200_E_ALIGN, "gemm operand pointer not 16B aligned");
  MB_REQUIRE(op.ld % 8 == 0, MB200_E_ALIGN, "gemm operand ld (%lld) must be a multiple of 8 elements",
             (long long)op.ld);
  MB_REQUIRE((nb0 == 1 || op.bs0 % 8 == 0) && (nb1 == 1 || op.bs1 % 8 == 0), MB200_E_ALIGN,
             "gemm operand batch strides must be multiples of 8 elements");
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (!op.mn_major) {
    dims[0] = (cuuint64_t)K;
    dims[1] = (cuuint64_t)rows;
    box[0] = BK;
    box[1] = (cuuint32_t)box_rows;
  } else {
    dims[0] = (cuuint64_t)rows;
    dims[1] = (cuuint64_t)K;
    box[0] = 64;
    box[1] = BK;
  }
  dims[2] = (cuuint64_t)nb0;
  dims[3] = (cuuint64_t)nb1;
  box[2] = 1;
  box[3] = 1;
  // strides of dims 1..3 in bytes; a size-1 batch dim still needs a legal (16B-multiple) stride
  strides[0] = (cuuint64_t)op.ld * 2;
  strides[1] = (cuuint64_t)(nb0 > 1 ? op.bs0 : op.ld) * 2;
  strides[2] = (cuuint64_t)(nb1 > 1 ? op.bs1 : op.ld) * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MB_REQUIRE(r == CUDA_SUCCESS, MB200_E_CUDA,
             "cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u)",
             (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
             (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
             (unsigned long long)strides[2], box[0], box[1]);
  return 0;
}

template <int BN, bool A_MN, bool B_MN, typename OutT>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKernelParams& kp,
                       cudaStream_t stream) {
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, OutT>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes));
    attr_set = true;
  }
  int grid = kp.total_tiles < num_sms() ? kp.total_tiles : num_sms();
  {
    const double nb = (double)kp.total_tiles / ((double)kp.tiles_m * kp.tiles_n);
    const double flops = 2.0 * kp.M * (double)kp.N * kp.K * nb;
    const double bytes = nb * (2.0 * ((double)kp.M * kp.K + (double)kp.N * kp.K) + (double)sizeof(OutT) * kp.M * kp.N);
    GemmProfScope prof(stream, flops, bytes);
    kern<<<grid, kThreads, Cfg<BN>::kSmemBytes, stream>>>(tmA, tmB, kp);
  }
  count_launch();
  MB_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, typename OutT>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                          const GemmKernelParams& kp, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false, OutT>(tmA, tmB, kp, s);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true, OutT>(tmA, tmB, kp, s);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false, OutT>(tmA, tmB, kp, s);
  return launch_gemm<BN, true, true, OutT>(tmA, tmB, kp, s);
}

static int pick_bn(int M, int N, int batches) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // prefer 256-wide tiles (halves A re-reads, 96 B/clk smem demand) unless they leave the grid short
  const int tiles_m = (M + BM - 1) / BM;
  const long long t256 = (long long)tiles_m * ((N + 255) / 256) * batches;
  const long long t128 = (long long)tiles_m * ((N + 127) / 128) * batches;
  const int sms = num_sms();
  if (t256 >= sms) return 256;
  // few tiles: compare wave efficiency
  auto eff = [&](long long t) { return (double)t / (double)(((t + sms - 1) / sms) * sms); };
  return eff(t256) + 1e-9 >= eff(t128) ? 256 : 128;
}

int gemm_impl(const mb200_gemm_args* a, cudaStream_t stream) {
  MB_REQUIRE(a != nullptr, MB200_E_ARG, "null gemm args");
  MB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->nb0 > 0 && a->nb1 > 0, MB200_E_SHAPE,
             "gemm: bad shape M=%d N=%d K=%d nb=(%d,%d)", a->M, a->N, a->K, a->nb0, a->nb1);
  MB_REQUIRE(a->c_dtype == MB200_BF16 || a->c_dtype == MB200_F32, MB200_E_DTYPE, "gemm: bad c_dtype %d", a->c_dtype);
  MB_REQUIRE(!(a->accumulate && a->c_dtype != MB200_F32), MB200_E_DTYPE, "gemm: accumulate needs f32 output");
  MB_REQUIRE(!(a->dact && !a->aux_in), MB200_E_ARG, "gemm: dact needs aux_in");
  const int celt = a->c_dtype == MB200_F32 ? 4 : 8;
  MB_REQUIRE((reinterpret_cast<uintptr_t>(a->C) & 15) == 0 && a->ldc % celt == 0, MB200_E_ALIGN,
             "gemm: C must be 16B aligned with ldc multiple of %d", celt);
  MB_REQUIRE((a->nb0 == 1 || a->c_bs0 % celt == 0) && (a->nb1 == 1 || a->c_bs1 % celt == 0), MB200_E_ALIGN,
             "gemm: C batch strides must be multiples of %d elements", celt);
  if (a->res1 || a->res2) MB_REQUIRE(a->ld_res % 8 == 0, MB200_E_ALIGN, "gemm: ld_res must be a multiple of 8");
  if (a->aux_in || a->aux_out)
    MB_REQUIRE(a->ldc % 8 == 0, MB200_E_ALIGN, "gemm: aux tensors share ldc, which must be a multiple of 8");
  int rc = check_arch();
  if (rc) return rc;

  int bn = a->force_bn ? a->force_bn : pick_bn(a->M, a->N, a->nb0 * a->nb1);
  MB_REQUIRE(bn == 64 || bn == 128 || bn == 256, MB200_E_ARG, "gemm: force_bn must be 64/128/256");

  CUtensorMap tmA, tmB;
  rc = make_operand_map(&tmA, a->A, a->M, a->K, a->nb0, a->nb1, BM);
  if (rc) return rc;
  rc = make_operand_map(&tmB, a->B, a->N, a->K, a->nb0, a->nb1, bn);
  if (rc) return rc;

  GemmKernelParams kp;
  kp.M = a->M;
  kp.N = a->N;
  kp.K = a->K;
  kp.nb0 = a->nb0;
  kp.tiles_m = (a->M + BM - 1) / BM;
  kp.tiles_n = (a->N + bn - 1) / bn;
  kp.total_tiles = kp.tiles_m * kp.tiles_n * a->nb0 * a->nb1;
  kp.C = a->C;
  kp.ldc = a->ldc;
  kp.c_bs0 = a->c_bs0;
  kp.c_bs1 = a->c_bs1;
  kp.alpha = a->alpha;
  kp.act = a->act;
  kp.dact = a->dact;
  kp.accumulate = a->accumulate;
  kp.bias = reinterpret_cast<const bf16*>(a->bias);
  kp.aux_out = reinterpret_cast<bf16*>(a->aux_out);
  kp.aux_in = reinterpret_cast<const bf16*>(a->aux_in);
  kp.res1 = reinterpret_cast<const bf16*>(a->res1);
  kp.res2 = reinterpret_cast<const bf16*>(a->res2);
  kp.ld_res = a->ld_res;

  const bool amn = a->A.mn_major != 0, bmn = a->B.mn_major != 0;
  const bool f32 = a->c_dtype == MB200_F32;
  switch (bn) {
    case 64:
      return f32 ? dispatch_major<64, float>(amn, bmn, tmA, tmB, kp, stream)
                 : dispatch_major<64, bf16>(amn, bmn, tmA, tmB, kp, stream);
    case 128:
      return f32 ? dispatch_major<128, float>(amn, bmn, tmA, tmB, kp, stream)
                 : dispatch_major<128, bf16>(amn, bmn, tmA, tmB, kp, stream);
    default:
      return f32 ? dispatch_major<256, float>(amn, bmn, tmA, tmB, kp, stream)
                 : dispatch_major<256, bf16>(amn, bmn, tmA, tmB, kp, stream);
  }
}

}  // namespace mb200

extern "C" int mb200_gemm(const mb200_gemm_args* args, void* stream) {
  return mb200::gemm_impl(args, reinterpret_cast<cudaStream_t>(stream));
}
