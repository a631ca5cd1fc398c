// D1 (2-SM variant): the same persistent warp-specialised GEMM as gemm_tc.cu, but every work item is
// computed by a CTA PAIR with `tcgen05.mma.cta_group::2`: one 256 x BN x 16 UMMA spans both SMs
// (each CTA holds its 128 rows of A, half of the B tile and its 128 accumulator lanes in TMEM), which
// halves the shared-memory operand bandwidth per SM -- the limiter of the 1-SM kernel at CTR shapes
// (UMMA operand reads + TMA writes exceed 128 B/clk/SM, see profiles/gemm_microbench.md).
//   * both CTAs issue TMA loads (`.cta_group::2`) for their halves; all bytes are accounted on the
//     LEADER's full barrier (peer bit of the shared-window address cleared)
//   * only the leader's elected thread issues MMAs; `tcgen05.commit...multicast::cluster` releases the
//     smem slot / publishes the accumulator in BOTH CTAs
//   * both CTAs run the fused epilogue on their own 128 TMEM lanes and release the accumulator stage
//     with a remote mbarrier arrive on the leader
#include <cstdio>
#include <cstdlib>

#include "gemm_epilogue.cuh"

namespace hctr {

constexpr int T2_BLOCK_M = 128;  // rows per CTA (256 per pair)
constexpr int T2_BLOCK_K = 64;
constexpr int T2_UMMA_K = 16;
constexpr int T2_THREADS = 256;
constexpr int T2_EPI_WARP0 = 4;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

HCTR_DEVICE void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0,
                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
HCTR_DEVICE void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
HCTR_DEVICE void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
HCTR_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask)
               : "memory");
}
HCTR_DEVICE void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
HCTR_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

constexpr int kSlabBytes = 32 * 128;          // [32 rows x 64 bf16], 128-byte swizzled rows
constexpr int kSmemLimit = 232448;            // 227 KB opt-in limit of sm_100

template <int BN, int KIND>
struct Smem2 {
  static constexpr int kABytes = T2_BLOCK_M * T2_BLOCK_K * 2;      // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * T2_BLOCK_K * 2;        // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue slabs per warp: 2 output buffers (+2 aux-output for CROSS) + 2 x (operand slabs)
  static constexpr int kLdTensors = KIND == EK_CROSS ? 2 : (KIND == EK_GENERIC ? 0 : 1);
  static constexpr int kStSlabs = KIND == EK_CROSS ? 4 : 2;
  static constexpr int kWarpEpiBytes = (kStSlabs + 2 * kLdTensors) * kSlabBytes;
  static constexpr int kEpiBytes = 4 * kWarpEpiBytes;
  static constexpr int kMaxStages = (BN == 256) ? 6 : 8;
  static constexpr int kFit = (kSmemLimit - 2048 - kEpiBytes) / kStageBytes;
  static constexpr int kStages = kFit < kMaxStages ? kFit : kMaxStages;
  static constexpr int kTotal = kStages * kStageBytes + 1024 + 1024 + kEpiBytes;
};

template <int BN, bool A_MN, bool B_MN, int KIND>
__global__ void __launch_bounds__(T2_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                    const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmXl,
                    const GemmParams p, const int tma_epi) {
  using L = Smem2<BN, KIND>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * L::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
  uint64_t* full_bar = bars;                    // waited by the leader's MMA warp
  uint64_t* empty_bar = bars + kStages;         // per CTA: slot drained (leader commit, multicast)
  uint64_t* tmem_full = bars + 2 * kStages;     // per CTA: accumulator ready (leader commit, multicast)
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // leader only: 8 epilogue warps of the pair
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* epi_bar = bars + 2 * kStages + 6;   // 2 operand-slab barriers per epilogue warp
  uint8_t* epi_smem = smem + kStages * L::kStageBytes + 1024;

  const int warp_idx = threadIdx.x >> 5;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const bool leader = cta_rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int total_items = m_pairs * p.n_tiles * p.splits;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_idx == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&epi_bar[i], 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc_2sm(tmem_ptr, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < total_items; t += num_pairs) {
        const int n_blk = t % p.n_tiles;
        const int m_blk = ((t / p.n_tiles) % m_pairs) * 2 + cta_rank;
        const int split = t / (p.n_tiles * m_pairs);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        const int m0 = m_blk * T2_BLOCK_M;
        const int n0 = n_blk * BN + cta_rank * (BN / 2);   // this CTA's half of the N range
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          uint8_t* sa = smem_a + stage * L::kABytes;
          uint8_t* sb = smem_b + stage * L::kBBytes;
          const int k0 = kb * T2_BLOCK_K;
          if constexpr (A_MN) {
#pragma unroll
            for (int c = 0; c < T2_BLOCK_M / 64; ++c)
              tma_load_2d_2sm(sa + c * (T2_BLOCK_K * 128), &tmA, &full_bar[stage], m0 + c * 64, k0);
          } else {
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 2 / 64; ++c)
              tma_load_2d_2sm(sb + c * (T2_BLOCK_K * 128), &tmB, &full_bar[stage], n0 + c * 64, k0);
          } else {
            tma_load_2d_2sm(sb, &tmB, &full_bar[stage], k0, n0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1 && leader) {
    // ===================== MMA issuer (leader CTA, one elected thread) =====================
    constexpr uint32_t idesc =
        make_idesc(kFmtBF16, kFmtBF16, A_MN ? 1u : 0u, B_MN ? 1u : 0u, 2 * T2_BLOCK_M, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = pair_id; t < total_items; t += num_pairs) {
      const int split = t / (p.n_tiles * m_pairs);
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem_a + stage * L::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * L::kBBytes);
          const uint64_t a_desc = A_MN ? make_smem_desc_sw128(a_addr, T2_BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = B_MN ? make_smem_desc_sw128(b_addr, T2_BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(b_addr, 16, 1024);
          constexpr uint32_t a_step = A_MN ? (T2_UMMA_K * 128) >> 4 : (T2_UMMA_K * 2) >> 4;
          constexpr uint32_t b_step = B_MN ? (T2_UMMA_K * 128) >> 4 : (T2_UMMA_K * 2) >> 4;
#pragma unroll
          for (int k = 0; k < T2_BLOCK_K / T2_UMMA_K; ++k) {
            umma_f16_2sm(d_tmem, a_desc + static_cast<uint64_t>(k * a_step),
                         b_desc + static_cast<uint64_t>(k * b_step), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm_mc(&empty_bar[stage], 0b11);                  // slot free in both CTAs
          if (kb == kb1 - 1) umma_commit_2sm_mc(&tmem_full[as], 0b11);  // accumulator ready in both
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp_idx >= T2_EPI_WARP0) {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes) =====================
    const int ew = warp_idx - T2_EPI_WARP0;
    const int lane = threadIdx.x & 31;
    int as = 0;
    uint32_t aphase = 0;
    const int flags = p.flags;
    if (!tma_epi) {
      // ---- direct epilogue (fp32 / atomic outputs, unaligned operands): one thread per row
      for (int t = pair_id; t < total_items; t += num_pairs) {
        const int n_blk = t % p.n_tiles;
        const int m_blk = ((t / p.n_tiles) % m_pairs) * 2 + cta_rank;
        const int m = m_blk * T2_BLOCK_M + ew * 32 + lane;
        const bool row_ok = m < p.M;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int n0 = n_blk * BN + c * 32;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + as * BN + c * 32 + (static_cast<uint32_t>(ew * 32) << 16), r);
          tmem_ld_wait();
          epilogue_chunk(p, flags, m, row_ok, n0, r);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);   // 8 arrivals (4 warps x 2 CTAs)
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    } else {
      // ---- TMA-staged epilogue: per-warp [32 x 64] slabs, bulk tensor stores, operand slabs
      // requested two 64-column groups (>= one 128-column tile) ahead
      constexpr int kG = BN / 64;
      constexpr int kLd = L::kLdTensors;
      const uint32_t wbase = smem_u32(epi_smem + ew * L::kWarpEpiBytes);
      const uint32_t ld_base = wbase + L::kStSlabs * kSlabBytes;
      uint64_t* lbar = epi_bar + ew * 2;
      const bool has_aux = (KIND == EK_CROSS) && p.aux != nullptr;
      // issue cursor of the operand loads (runs two groups ahead of the consume cursor)
      int it = pair_id, ig = 0;
      uint32_t qi = 0;
      auto issue_next = [&]() {
        if constexpr (kLd > 0) {
          if (it < total_items) {
            if (lane == 0) {
              const int n0 = (it % p.n_tiles) * BN + ig * 64;
              const int m0 = (((it / p.n_tiles) % m_pairs) * 2 + cta_rank) * T2_BLOCK_M + ew * 32;
              const uint32_t b = qi & 1u;
              uint8_t* dst = epi_smem + ew * L::kWarpEpiBytes + L::kStSlabs * kSlabBytes +
                             b * (kLd * kSlabBytes);
              mbar_arrive_expect_tx(&lbar[b], kLd * kSlabBytes);
              tma_load_2d(dst, &tmX0, &lbar[b], n0, m0);
              if constexpr (kLd == 2) tma_load_2d(dst + kSlabBytes, &tmXl, &lbar[b], n0, m0);
            }
            ++qi;
            if (++ig == kG) {
              ig = 0;
              it += num_pairs;
            }
          }
        }
      };
      issue_next();
      issue_next();
      uint32_t q = 0;
      for (int t = pair_id; t < total_items; t += num_pairs) {
        const int n_blk = t % p.n_tiles;
        const int m0 = (((t / p.n_tiles) % m_pairs) * 2 + cta_rank) * T2_BLOCK_M + ew * 32;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int g = 0; g < kG; ++g, ++q) {
          const uint32_t b = q & 1u;
          const uint32_t st_out = wbase + b * kSlabBytes;
          const uint32_t st_aux = has_aux ? wbase + (2 + b) * kSlabBytes : 0u;
          const uint32_t ld_a = ld_base + b * (kLd * kSlabBytes);
          const uint32_t ld_b = (KIND == EK_CROSS) ? ld_a + kSlabBytes : ld_a;
          // the slab pair `b` was handed to the TMA store engine two groups ago
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          if constexpr (kLd > 0) mbar_wait(&lbar[b], (q >> 1) & 1u);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c = g * 2 + half;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + as * BN + c * 32 + (static_cast<uint32_t>(ew * 32) << 16), r);
            tmem_ld_wait();
            epilogue_chunk_tma<KIND>(p, flags, lane, n_blk * BN + c * 32, half, r, st_out, st_aux, ld_a,
                                     ld_b);
          }
          fence_proxy_async();
          __syncwarp();
          if (p.colsum != nullptr) {
            // fused bias gradient of the consumer layer: column sums of this warp's finished slab
            // (lane owns 2 columns; a slab row is read conflict-free: all lanes hit the same 128 bytes)
            const int rows_valid = min(32, p.M - m0);
            float cs0 = 0.f, cs1 = 0.f;
            for (int r = 0; r < rows_valid; ++r) {
              uint32_t w;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w)
                           : "r"(st_out + slab_off(r, lane >> 2) + ((lane & 3) << 2)));
              cs0 += bf16_lo(w);
              cs1 += bf16_hi(w);
            }
            const int col = n_blk * BN + g * 64 + 2 * lane;
            if (rows_valid > 0) {
              if (col < p.N) atomicAdd(p.colsum + col, cs0);
              if (col + 1 < p.N) atomicAdd(p.colsum + col + 1, cs1);
            }
          }
          if (lane == 0) {
            const int n0 = n_blk * BN + g * 64;
            if (n0 < p.N) {
              tma_store_2d(&tmOut, epi_smem + ew * L::kWarpEpiBytes + b * kSlabBytes, n0, m0);
              if (has_aux)
                tma_store_2d(&tmAux, epi_smem + ew * L::kWarpEpiBytes + (2 + b) * kSlabBytes, n0, m0);
            }
            tma_store_commit();
          }
          issue_next();   // operand slab `b` has been fully consumed (ordered by the __syncwarp above)
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
      if (lane == 0) tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                     const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled2 get_encode2() {
  static PFN_encodeTiled2 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) !=
            cudaSuccess || ptr == nullptr)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled2>(ptr);
  }
  return fn;
}
static int make_tmap2(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled2 enc = get_encode2();
  if (enc == nullptr) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

static int g2_num_sms = 0;

struct EpiMaps {
  CUtensorMap out, aux, x0, xl;
  int enabled;
};

template <int BN, bool A_MN, bool B_MN, int KIND>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const EpiMaps& em,
                   const GemmParams& p, cudaStream_t stream) {
  using L = Smem2<BN, KIND>;
  static bool attr_set = false;
  auto kern = gemm_tc2_kernel<BN, A_MN, B_MN, KIND>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess)
      return -3;
    attr_set = true;
  }
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int total = m_pairs * p.n_tiles * p.splits;
  int pairs = g2_num_sms / 2;
  if (total < pairs) pairs = total;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pairs * 2);
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, tb, em.out, em.aux, em.x0, em.xl, p, em.enabled) != cudaSuccess)
    return -4;
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace hctr

using namespace hctr;

// same contract as hctr_gemm_bf16 (gemm_tc.cu); block_n in {128, 256}
extern "C" int hctr_gemm_bf16_2sm(const void* A, const void* B, void* out, int M, int N, int K,
                                  long long lda, long long ldb, long long ldo, int a_mn, int b_mn,
                                  const float* bias, const void* mask, long long ldmask,
                                  const void* x0, const void* xl, long long ldx, void* aux,
                                  long long ldaux, float alpha, int flags, int splits, int block_n,
                                  const float* addf, long long ldaddf, float* colsum, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (g2_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g2_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int BN = block_n == 256 ? 256 : 128;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + T2_BLOCK_M - 1) / T2_BLOCK_M;
  p.n_tiles = (N + BN - 1) / BN;
  p.k_blocks = (K + T2_BLOCK_K - 1) / T2_BLOCK_K;
  if (splits < 1) splits = 1;
  if (splits > p.k_blocks) splits = p.k_blocks;
  p.kb_per_split = (p.k_blocks + splits - 1) / splits;
  p.splits = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits > 1 && !(flags & EPI_ATOMIC)) return -10;
  p.out = out; p.ldo = ldo; p.aux = aux; p.ldaux = ldaux; p.bias = bias;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask); p.ldmask = ldmask;
  p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0);
  p.xl = reinterpret_cast<const __nv_bfloat16*>(xl); p.ldx = ldx;
  p.alpha = alpha; p.flags = flags; p.addf = addf; p.ldaddf = ldaddf;
  p.colsum = nullptr;
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_tmap2(&ta, A, M, K, lda, 64, T2_BLOCK_K);
  else      rc = make_tmap2(&ta, A, K, M, lda, T2_BLOCK_K, T2_BLOCK_M);
  if (rc) return rc;
  if (b_mn) rc = make_tmap2(&tb, B, N, K, ldb, 64, T2_BLOCK_K);
  else      rc = make_tmap2(&tb, B, K, N, ldb, T2_BLOCK_K, BN / 2);
  if (rc) return rc - 10;
  // TMA epilogue: bf16 output, 16-byte aligned operands with 16-byte multiple row pitches
  int kind = epi_kind_of(flags, addf);
  EpiMaps em;
  em.enabled = 0;
  auto tma_ok = [](const void* ptr, long long ld) {
    return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld % 8) == 0;
  };
  static const bool tma_epi_off = getenv("HCTR_GEMM_DIRECT_EPI") != nullptr;
  if (kind >= 0 && !tma_epi_off && tma_ok(out, ldo)) {
    bool ok = make_tmap2(&em.out, out, N, M, ldo, 64, 32) == 0;
    em.aux = em.x0 = em.xl = em.out;
    if (ok && kind == EK_CROSS) {
      ok = tma_ok(x0, ldx) && tma_ok(xl, ldx) && (aux == nullptr || tma_ok(aux, ldaux)) &&
           make_tmap2(&em.x0, x0, N, M, ldx, 64, 32) == 0 && make_tmap2(&em.xl, xl, N, M, ldx, 64, 32) == 0 &&
           (aux == nullptr || make_tmap2(&em.aux, aux, N, M, ldaux, 64, 32) == 0);
    } else if (ok && kind == EK_ADD) {
      ok = tma_ok(xl, ldx) && make_tmap2(&em.x0, xl, N, M, ldx, 64, 32) == 0;
    } else if (ok && kind == EK_MASK) {
      ok = tma_ok(mask, ldmask) && make_tmap2(&em.x0, mask, N, M, ldmask, 64, 32) == 0;
    }
    em.enabled = ok ? 1 : 0;
  }
  // the fused column sums need the staged slabs; rows >= M must be zero in the slab, which holds for
  // the bias-free mask / plain epilogues (zero-filled operands)
  bool colsum_fused = false;
  if (em.enabled && colsum != nullptr && bias == nullptr && (kind == EK_MASK || (kind == EK_GENERIC && !(flags & EPI_SIGMOID)))) {
    p.colsum = colsum;
    colsum_fused = true;
  }
  if (!em.enabled) {
    kind = EK_GENERIC;
    em.out = em.aux = em.x0 = em.xl = ta;   // unused
  }
#define HCTR_RET2(X) { const int rc_ = (X); return (rc_ == 0 && colsum_fused) ? 100 : rc_; }
#define HCTR_DISPATCH2K(BNV, KV)                                             \
  if (a_mn) {                                                                \
    if (b_mn) HCTR_RET2((launch2<BNV, true, true, KV>(ta, tb, em, p, stream)))        \
    HCTR_RET2((launch2<BNV, true, false, KV>(ta, tb, em, p, stream)))                 \
  } else {                                                                   \
    if (b_mn) HCTR_RET2((launch2<BNV, false, true, KV>(ta, tb, em, p, stream)))       \
    HCTR_RET2((launch2<BNV, false, false, KV>(ta, tb, em, p, stream)))                \
  }
#define HCTR_DISPATCH2(BNV)                                                  \
  switch (kind) {                                                            \
    case EK_CROSS: HCTR_DISPATCH2K(BNV, EK_CROSS)                            \
    case EK_ADD: HCTR_DISPATCH2K(BNV, EK_ADD)                                \
    case EK_MASK: HCTR_DISPATCH2K(BNV, EK_MASK)                              \
    default: HCTR_DISPATCH2K(BNV, EK_GENERIC)                                \
  }
  if (BN == 256) { HCTR_DISPATCH2(256) }
  HCTR_DISPATCH2(128)
#undef HCTR_DISPATCH2K
#undef HCTR_DISPATCH2
}
