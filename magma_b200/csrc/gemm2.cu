// magma_b200 — 2-CTA (cta_group::2) variant of the bf16 GEMM core.
//
// Why: with one CTA per 128x256 tile every SM pulls (128 + 256) x 64 x 2 B = 48 KB per 512 tensor-pipe cycles from L2
// (94 B/clk/SM), ~84 % of the per-SM crossbar peak; ncu shows the tensor pipe active only 73-82 % of the time on the
// GPT-J shapes (profiles/r01_gemm_ncu_full_summary_v1.txt). Pairing the two SMs of a TPC on one 256x256 tile
// (tcgen05.mma.cta_group::2, UMMA M = 256) lets each SM stage its own 128 rows of A and only HALF of the B tile:
// 32 KB per 512 cycles (64 B/clk/SM), and the smaller stages give a 6-deep ring instead of 4.
//
// Protocol (rank 0 = leader):
//   * both CTAs run a TMA producer; all TMA completions are credited to the LEADER's full barrier
//     (cp.async.bulk.tensor .cta_group::2 with the peer bit of the barrier address cleared); the leader posts
//     expect_tx for the bytes of both CTAs;
//   * only the leader issues MMAs; tcgen05.commit .multicast::cluster releases the smem slot in BOTH CTAs and
//     publishes the accumulator to BOTH CTAs' epilogue warps;
//   * each CTA's epilogue drains its own 128 TMEM lanes — epilogue v3 (gemm_common.cuh): row-per-thread fused
//     arithmetic, SWIZZLE_128B staging, output by TMA store / reduce-add — and arrives on the leader's tmem_empty
//     barrier (the peer through mapa / mbarrier.arrive.shared::cluster).
#include "gemm_common.cuh"

namespace mb200 {

static constexpr int BN2 = 256;   // tile N (UMMA N)
static constexpr int HB = 128;    // B rows staged per CTA
static constexpr int kA2 = BM * BK * 2;
static constexpr int kB2 = HB * BK * 2;
static constexpr int kStage2 = kA2 + kB2;               // 32 KB per CTA per stage
static constexpr int kStages2 = kSmemBudget / kStage2;  // 6
static constexpr int kEpi2 = 4 * 32 * 64 * 4;
static constexpr int kSmem2 = kStages2 * kStage2 + kEpi2 + 1024 + 256;

// SK = true is the stream-K variant (opt-in, gemm.cu): kept in its own instantiation because its owner / helper epilogue
// paths cost the default kernel registers (ptxas: 408 bytes of spills in the epilogue warps when both lived in one kernel,
// +5-13 % on every short GEMM of the step).
// Registers: the epilogue warps need ~230 registers (a 32-column fp32 row chunk + prefetched epilogue inputs), the
// producer / MMA / TMEM-allocator warps < 72 (40 spills a little). Launched at kRegsLaunch per thread and redistributed with setmaxnreg
// (warpgroup 0 shrinks to kRegsLean, warpgroup 1 grows to kRegsEpi: 128 * 72 + 128 * 232 = 256 * 152), the CTA holds
// 38 912 of the SM's 65 536 registers instead of all of them, so blocks of OTHER kernels — the optimizer on its side
// stream, the next launch's PDL prologue of a shared-memory-light kernel — can be resident beside a persistent GEMM CTA.
static constexpr int kRegsLaunch = 152, kRegsLean = 72, kRegsEpi = 232;
static_assert(128 * kRegsLean + 128 * kRegsEpi == kThreads * kRegsLaunch, "setmaxnreg budget");

template <bool A_MN, bool B_MN, typename OutT, bool SK>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(kRegsLaunch)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                     const __grid_constant__ CUtensorMap tmWs, const __grid_constant__ GemmKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages2 * kA2;
  float* epi_stage = reinterpret_cast<float*>(smem + kStages2 * kStage2);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * kStage2 + kEpi2);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full = empty_bar + kStages2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_kb = (p.K + BK - 1) / BK;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmC);
    if (p.aux_out) tma_prefetch_desc(&tmAux);
    if (SK && p.sk_on) tma_prefetch_desc(&tmWs);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's: one arrive.expect_tx by the leader's producer (+ tx bytes of both CTAs)
      mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // leader's: 4 epilogue warps x 2 CTAs
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // PDL: every thread that touches global memory written by earlier kernels waits for them (griddepcontrol.wait) —
  // the producer before its first A load, the epilogue warps before their first residual load / store. The MMA
  // issuer only reads shared memory and TMEM.

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLean));
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // Frozen weights are not produced by any kernel of the stream: the B halves of the first ring stages are requested
      // NOW, while this launch still waits for its predecessor, so the first HBM round trip (~1.5 us of weights the L2
      // has never seen) is not added to the exposed start-up of every GEMM.
      int npre = 0;
      WorkItem w0;
      const bool any = next_item<SK>(p, cid, ncl, num_kb, 0, w0);
      if (p.b_static && any) {
        const int tpb = p.tiles_m * p.tiles_n;
        const int z = w0.tile / tpb;
        const int r = w0.tile - z * tpb;
        const int n_row = (r / p.tiles_m) * BN2 + (int)rank * HB;
        const int z0 = z % p.nb0, z1 = z / p.nb0;
        const int nk0 = w0.kb1 - w0.kb0;
        npre = nk0 < kStages2 ? nk0 : kStages2;
        for (int i = 0; i < npre; ++i) {
          const int kb = w0.kb0 + i;
          if (rank == 0) mbar_expect_tx(&full_bar[i], 2 * kStage2);
          uint8_t* sb = smem_b + i * kB2;
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < HB / 64; ++c)
              tma_load_4d_2sm(sb + c * 8192, &tmB, &full_bar[i], n_row + c * 64, kb * BK, z0, z1);
          } else {
            tma_load_4d_2sm(sb, &tmB, &full_bar[i], kb * BK, n_row, z0, z1);
          }
        }
      }
      pdl_wait();
      WorkItem w;
      for (int it = 0; next_item<SK>(p, cid, ncl, num_kb, it, w); ++it) {
        const int tpb = p.tiles_m * p.tiles_n;
        const int z = w.tile / tpb;
        const int r = w.tile - z * tpb;
        const int m_blk = (r % p.tiles_m) * 2 + (int)rank;  // in 128-row units
        const int n_row = (r / p.tiles_m) * BN2 + (int)rank * HB;
        const int z0 = z % p.nb0, z1 = z / p.nb0;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          const bool preloaded = it == 0 && kb - w.kb0 < npre;  // slot known free, expect_tx posted, B already in flight
          if (!preloaded) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStage2);
          }
          uint8_t* sa = smem_a + stage * kA2;
          uint8_t* sb = smem_b + stage * kB2;
          if constexpr (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_4d_2sm(sa + i * 8192, &tmA, &full_bar[stage], m_blk * BM + i * 64, kb * BK, z0, z1);
          } else {
            tma_load_4d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM, z0, z1);
          }
          if (!preloaded) {
            if constexpr (B_MN) {
#pragma unroll
              for (int i = 0; i < HB / 64; ++i)
                tma_load_4d_2sm(sb + i * 8192, &tmB, &full_bar[stage], n_row + i * 64, kb * BK, z0, z1);
            } else {
              tma_load_4d_2sm(sb, &tmB, &full_bar[stage], kb * BK, n_row, z0, z1);
            }
          }
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, single thread) =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      constexpr uint32_t a_lbo = A_MN ? 8192u : 0u, b_lbo = B_MN ? 8192u : 0u;
      constexpr uint32_t a_kadv = A_MN ? 2048u : 32u, b_kadv = B_MN ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      WorkItem w;
      for (int it = 0; next_item<SK>(p, cid, ncl, num_kb, it, w); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN2);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * kA2);
          const uint32_t sb = smem_u32(smem_b + stage * kB2);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_kadv, a_lbo, 1024);
            const uint64_t db = make_smem_desc(sb + k * b_kadv, b_lbo, 1024);
            umma_bf16_2sm(d_tmem, da, db, idesc, ((kb - w.kb0) | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], 3);  // frees the slot in both CTAs
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full[acc], 3);  // accumulator complete -> both epilogues
      }
    }
  }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    pdl_wait();
    const int q = warp & 3;
    uint32_t box = 0;  // boxes published so far (staging slot = box & 1)
    WorkItem w;
    for (int it = 0; next_item<SK>(p, cid, ncl, num_kb, it, w); ++it) {
      const int t = w.tile;
      const int tpb = p.tiles_m * p.tiles_n;
      const int z = t / tpb;
      const int r = t - z * tpb;
      const int m_blk = (r % p.tiles_m) * 2 + (int)rank;
      const int n_blk = r / p.tiles_m;
      const int z0 = z % p.nb0, z1 = z / p.nb0;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if (p.epi_kind == EK_GENERIC) {
        // option combinations without a specialised instantiation: the v2 per-lane path (fp32 staging in the same pool)
        EpiCtx c;
        c.tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2);
        c.stg_s = smem_u32(epi_stage + q * (32 * 64));
        c.tmem_full = &tmem_full[acc];
        c.tmem_empty = &tmem_empty[acc];
        c.empty_remote = rank != 0;
        c.ks = 0;
        c.full_phase = acc_phase;
        c.boff = (long long)z0 * p.c_bs0 + (long long)z1 * p.c_bs1;
        c.row0 = m_blk * BM + q * 32;
        c.nrows = max(0, min(32, p.M - c.row0));
        c.n_blk = n_blk;
        c.lane = lane;
        epi_tile<BN2, 0, 0, 0, false, false, false, true, OutT>(p, c);
        continue;
      }
      Epi3Ctx c;
      c.tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2);
      c.pool_s = smem_u32(epi_stage);
      c.tmem_full = &tmem_full[acc];
      c.tmem_empty = &tmem_empty[acc];
      c.empty_remote = rank != 0;
      c.full_phase = acc_phase;
      c.boff = (long long)z0 * p.c_bs0 + (long long)z1 * p.c_bs1;
      c.row_cta0 = m_blk * BM;
      c.q = q;
      c.lane = lane;
      c.n_blk = n_blk;
      c.z0 = z0;
      c.z1 = z1;
      c.tmC = &tmC;
      c.tmAux = &tmAux;
      c.npart = 0;
      if constexpr (SK) {
      if (w.kind == 2) {
        // ---- stream-K helper: this piece's fp32 partial tile -> workspace slot (plain TMA stores), then the flag ----
        GemmKernelParams pw = p;
        pw.N = BN2;
        pw.M = 0x7fffffff;
        pw.alpha = 1.f;
        pw.bias = nullptr;
        c.boff = 0;
        c.row_cta0 = w.slot * 256 + (int)rank * BM;
        c.n_blk = 0;
        c.z0 = c.z1 = 0;
        c.tmC = &tmWs;
        epi_tile_v3<BN2, 0, 0, 0, false, false, false, float>(pw, c, box);
        if (q == 0 && lane == 0) {
          tma_store_wait_all();  // the bulk stores have been performed ...
          asm volatile("fence.proxy.async;" ::: "memory");  // ... (async proxy -> generic proxy)
          __threadfence();       // ... and are ordered before the flag for any observer on the device
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + w.slot * 2 + (int)rank), "r"(p.sk_epoch)
                       : "memory");
        }
        continue;
      }
      if (w.kind == 1) {
        // ---- stream-K owner: the helpers' pieces of this tile's k-blocks [h, num_kb) ----
        const int ti = w.tile - p.sk_W * ncl;
        const long long x0 = (long long)ti * p.sk_tail, x1 = x0 + p.sk_tail;
        int j = (int)(x0 * p.sk_nh / p.sk_U);
        if (j > 0) --j;
        while (j + 1 < p.sk_nh && (long long)(j + 1) * p.sk_U / p.sk_nh <= x0) ++j;  // last helper starting at or before x0
        for (; j < p.sk_nh && c.npart < 6; ++j) {
          const long long a = (long long)j * p.sk_U / p.sk_nh, b = (long long)(j + 1) * p.sk_U / p.sk_nh;
          if (a >= x1) break;
          if (b <= x0 || b <= a) continue;
          const int slot = j * p.sk_pmax + (int)(ti - a / p.sk_tail);
          if (lane == 0) {  // wait until that helper's half for this CTA has landed
            const unsigned int* f = p.sk_flags + slot * 2 + (int)rank;
            unsigned int v;
            unsigned int spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              if (v != p.sk_epoch && ++spins > MB200_SPIN_LIMIT) __trap();
            } while (v != p.sk_epoch);
          }
          c.part[c.npart++] = p.sk_ws + ((long long)slot * 256 + (long long)rank * BM) * 256;
        }
        __syncwarp();
      }
      }  // SK
#define MB_EPI(ACT, DACT, NRES, AUX, ROPE, ACCUM)                                                   \
  do {                                                                                              \
    if constexpr (SK) {                                                                             \
      if (w.kind == 1) epi_tile_v3<BN2, ACT, DACT, NRES, AUX, ROPE, ACCUM, OutT, true>(p, c, box);  \
      else epi_tile_v3<BN2, ACT, DACT, NRES, AUX, ROPE, ACCUM, OutT, false>(p, c, box);             \
    } else {                                                                                        \
      epi_tile_v3<BN2, ACT, DACT, NRES, AUX, ROPE, ACCUM, OutT, false>(p, c, box);                  \
    }                                                                                               \
  } while (0)
      if constexpr (sizeof(OutT) == 4) {
        switch (p.epi_kind) {
          case EK_ACCUM: MB_EPI(0, 0, 0, false, false, true); break;
          default: MB_EPI(0, 0, 0, false, false, false); break;  // EK_PLAIN
        }
      } else {
        switch (p.epi_kind) {
          case EK_ROPE: MB_EPI(0, 0, 0, false, true, false); break;
          case EK_GELU: MB_EPI(MB200_ACT_GELU_NEW, 0, 0, false, false, false); break;
          case EK_GELU_AUX: MB_EPI(MB200_ACT_GELU_NEW, 0, 0, true, false, false); break;
          case EK_QGELU: MB_EPI(MB200_ACT_QUICK_GELU, 0, 0, false, false, false); break;
          case EK_RELU: MB_EPI(MB200_ACT_RELU, 0, 0, false, false, false); break;
          case EK_DGELU: MB_EPI(0, MB200_DACT_GELU_NEW, 0, false, false, false); break;
          case EK_DRELU: MB_EPI(0, MB200_DACT_RELU, 0, false, false, false); break;
          case EK_RES1: MB_EPI(0, 0, 1, false, false, false); break;
          case EK_RES2: MB_EPI(0, 0, 2, false, false, false); break;
          default: MB_EPI(0, 0, 0, false, false, false); break;  // EK_PLAIN
        }
      }
#undef MB_EPI
    }
    // the staging slots must outlive the bulk stores that read them (the writes themselves are complete, like every
    // memory operation of the grid, before a dependent grid's griddepcontrol.wait returns)
    if (q == 0 && lane == 0) tma_store_wait_read0();
  }

  // neither CTA may exit (or free TMEM) while the pair still reads its shared memory / TMEM
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <bool A_MN, bool B_MN, typename OutT, bool SK>
static int launch2_sk(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux,
                   const CUtensorMap& tmWs, const GemmKernelParams& kp, cudaStream_t stream) {
  auto kern = gemm2_tcgen05_kernel<A_MN, B_MN, OutT, SK>;
  static bool attr_set = false;
  if (!attr_set) {
    MB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
    attr_set = true;
  }
  const int max_clusters = gemm_sms() / 2;
  const int clusters = kp.sk_on ? kp.sk_R + kp.sk_nh : (kp.total_tiles < max_clusters ? kp.total_tiles : max_clusters);
  {
    const double nb = (double)kp.total_tiles / ((double)kp.tiles_m * kp.tiles_n);
    const double flops = 2.0 * kp.M * (double)kp.N * kp.K * nb;
    const double bytes = nb * (2.0 * ((double)kp.M * kp.K + (double)kp.N * kp.K) + (double)sizeof(OutT) * kp.M * kp.N);
    GemmProfScope prof(stream, flops, bytes);
    MB_CUDA(launch_pdl(kern, dim3(clusters * 2), dim3(kThreads), kSmem2, stream, tmA, tmB, tmC, tmAux, tmWs, kp));
  }
  count_launch();
  MB_CUDA(cudaGetLastError());
  return 0;
}

template <bool A_MN, bool B_MN, typename OutT>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux,
                   const CUtensorMap& tmWs, const GemmKernelParams& kp, cudaStream_t stream) {
  // stream-K is planned for K-major frozen-weight GEMMs with bf16 output only (gemm.cu); everything else takes the
  // default instantiation
  if constexpr (!A_MN && sizeof(OutT) == 2) {
    if (kp.sk_on) return launch2_sk<A_MN, B_MN, OutT, true>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
  }
  if (kp.sk_on) {
    set_error("gemm: stream-K was planned for an operand combination it is not built for");
    return MB200_E_ARG;
  }
  return launch2_sk<A_MN, B_MN, OutT, false>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
}

int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux,
                 const CUtensorMap& tmWs, const GemmKernelParams& kp, bool a_mn, bool b_mn, bool f32, cudaStream_t stream) {
  if (f32) {
    if (!a_mn && !b_mn) return launch2<false, false, float>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
    if (!a_mn && b_mn) return launch2<false, true, float>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
    if (a_mn && !b_mn) return launch2<true, false, float>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
    return launch2<true, true, float>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
  }
  if (!a_mn && !b_mn) return launch2<false, false, bf16>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
  if (!a_mn && b_mn) return launch2<false, true, bf16>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
  if (a_mn && !b_mn) return launch2<true, false, bf16>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
  return launch2<true, true, bf16>(tmA, tmB, tmC, tmAux, tmWs, kp, stream);
}

}  // namespace mb200
