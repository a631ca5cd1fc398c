// D1 / D3: persistent warp-specialised bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, TMEM
// accumulators, TMA operand staging, mbarrier pipelines) with fused epilogues:
//   bias + ReLU            (MLP fprop;       replaces cuBLASLt RELU_AUX_BIAS epilogue,
//                           reference HugeCTR/src/layers/functors/fused_gemm_functors.cu:21-289)
//   dReLU mask multiply    (MLP dgrad;       reference DRELU_BGRAD epilogue)
//   fp32 (atomic) accum    (MLP/Cross wgrad; reference beta=1 wgrad, fused_fc_layer_functors.cu:161)
//   x0 * (acc + b) + xl    (DCNv2 cross;     reference HugeCTR/src/layers/multi_cross_layer.cu:625-672)
// Operands may be K-major or MN-major, so fprop (X*W, W=[in,out]), dgrad (dY*W^T) and wgrad
// (X^T*dY) all run from the same row-major tensors without any transposed copies.
#include <cstdio>

#include "gemm_epilogue.cuh"

namespace hctr {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kEpiWarp0 = 4;

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kBarBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024 align slack
};

// MC == 2: thread-block cluster of two CTAs stacked along M that share the B tile -- each CTA loads
// half of it and TMA-multicasts it into both shared memories (halves the L2->SM operand traffic of
// B; the kernel is L2-bandwidth bound at CTR shapes, see profiles/).
template <int BN, bool A_MN, bool B_MN, int MC>
__global__ void __launch_bounds__(kNumThreads, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * L::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp_idx = threadIdx.x >> 5;
  // work items are (m-group of MC tiles, n tile, k split); the CTAs of a cluster take the MC
  // consecutive m tiles of one item
  const int m_groups = (p.m_tiles + MC - 1) / MC;
  const int total_tiles = m_groups * p.n_tiles * p.splits;
  const int cta_rank = (MC > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = (MC > 1) ? (blockIdx.x / MC) : blockIdx.x;
  const int num_clusters = (MC > 1) ? (gridDim.x / MC) : gridDim.x;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_idx == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], MC);   // every CTA of the cluster must have drained the slot
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const int n_blk = t % p.n_tiles;
        const int m_blk = ((t / p.n_tiles) % m_groups) * MC + cta_rank;
        const int split = t / (p.n_tiles * m_groups);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        const int m0 = m_blk * BLOCK_M, n0 = n_blk * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          uint8_t* sa = smem_a + stage * L::kABytes;
          uint8_t* sb = smem_b + stage * L::kBBytes;
          const int k0 = kb * BLOCK_K;
          if constexpr (A_MN) {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * (BLOCK_K * 128), &tmA, &full_bar[stage], m0 + c * 64, k0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          }
          if constexpr (MC == 1) {
            if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d(sb + c * (BLOCK_K * 128), &tmB, &full_bar[stage], n0 + c * 64, k0);
            } else {
              tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
            }
          } else {
            // this CTA fetches its 1/MC share of the B tile and multicasts it to the whole cluster
            constexpr uint16_t kMask = (1u << MC) - 1;
            if constexpr (B_MN) {
              constexpr int kChunks = BN / 64 / MC;
#pragma unroll
              for (int c = 0; c < kChunks; ++c) {
                const int cc = cta_rank * kChunks + c;
                tma_load_2d_mc(sb + cc * (BLOCK_K * 128), &tmB, &full_bar[stage], n0 + cc * 64, k0, kMask);
              }
            } else {
              constexpr int kRowsPer = BN / MC;
              tma_load_2d_mc(sb + cta_rank * (kRowsPer * 128), &tmB, &full_bar[stage], k0,
                             n0 + cta_rank * kRowsPer, kMask);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single elected thread) =====================
    constexpr uint32_t idesc =
        make_idesc(kFmtBF16, kFmtBF16, A_MN ? 1u : 0u, B_MN ? 1u : 0u, BLOCK_M, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int split = t / (p.n_tiles * m_groups);
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
          // K-major SW128 : 8-row atoms of 1024 B ; MN-major SW128: 64-wide chunks BLOCK_K*128 B
          // apart (LBO), 8-k-row groups 1024 B apart (SBO)
          const uint64_t a_desc = A_MN ? make_smem_desc_sw128(a_addr, BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = B_MN ? make_smem_desc_sw128(b_addr, BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(b_addr, 16, 1024);
          constexpr uint32_t a_step = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
          constexpr uint32_t b_step = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            umma_f16(d_tmem, a_desc + static_cast<uint64_t>(k * a_step),
                     b_desc + static_cast<uint64_t>(k * b_step), idesc,
                     (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if constexpr (MC == 1) umma_commit(&empty_bar[stage]);   // frees the smem slot
          else umma_commit_mc(&empty_bar[stage], (1u << MC) - 1);  // ... in every CTA of the cluster
          if (kb == kb1 - 1) umma_commit(&tmem_full[as]);   // accumulator ready
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
  } else if (warp_idx >= kEpiWarp0) {
    // ===================== epilogue: TMEM -> regs -> fused math -> global =====================
    const int ew = warp_idx - kEpiWarp0;  // == warp_idx % 4 : TMEM lane quarter
    const int lane = threadIdx.x & 31;
    int as = 0;
    uint32_t aphase = 0;
    const int flags = p.flags;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      const int n_blk = t % p.n_tiles;
      const int m_blk = ((t / p.n_tiles) % m_groups) * MC + cta_rank;
      const int m = m_blk * BLOCK_M + ew * 32 + lane;
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
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();   // no CTA may exit while a peer can still multicast into it
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        ptr == nullptr)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D bf16 row-major tensor [outer, inner] with leading dimension ld (elements)
static int make_tmap_bf16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer,
                          uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) return -1;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

static int g_num_sms = 0;

template <int BN, bool A_MN, bool B_MN, int MC>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                  cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MC>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) !=
        cudaSuccess)
      return -3;
    attr_set = true;
  }
  const int m_groups = (p.m_tiles + MC - 1) / MC;
  const int total = m_groups * p.n_tiles * p.splits;       // cluster work items
  int clusters = g_num_sms / MC;
  if (total < clusters) clusters = total;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * MC);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, tb, p) != cudaSuccess) return -4;
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace hctr

using namespace hctr;

// A: a_mn ? [K, M] row-major (lda) : [M, K] row-major (lda)
// B: b_mn ? [K, N] row-major (ldb) : [N, K] row-major (ldb)
// out: [M, N] row-major (ldo) bf16 (default) or fp32 (EPI_OUT_F32 | EPI_ATOMIC | EPI_ACCUM)
extern "C" int hctr_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K,
                              long long lda, long long ldb, long long ldo, int a_mn, int b_mn,
                              const float* bias, const void* mask, long long ldmask,
                              const void* x0, const void* xl, long long ldx, void* aux,
                              long long ldaux, float alpha, int flags, int splits, int block_n,
                              const float* addf, long long ldaddf, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int mc = (block_n >= 1000) ? 2 : 1;   // block_n = 1000 + BN selects the 2-CTA multicast kernel
  if (block_n >= 1000) block_n -= 1000;
  const int BN = (block_n == 256 || block_n == 64) ? block_n : 128;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = (N + BN - 1) / BN;
  p.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
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
  if (a_mn) rc = make_tmap_bf16(&ta, A, M, K, lda, 64, BLOCK_K);
  else      rc = make_tmap_bf16(&ta, A, K, M, lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_bf16(&tb, B, N, K, ldb, 64, BLOCK_K);
  else      rc = make_tmap_bf16(&tb, B, K, N, ldb, BLOCK_K, BN / mc);
  if (rc) return rc - 10;

#define HCTR_DISPATCH(BNV, MCV)                                                   \
  if (a_mn) {                                                                     \
    if (b_mn) return launch<BNV, true, true, MCV>(ta, tb, p, stream);             \
    return launch<BNV, true, false, MCV>(ta, tb, p, stream);                      \
  } else {                                                                        \
    if (b_mn) return launch<BNV, false, true, MCV>(ta, tb, p, stream);            \
    return launch<BNV, false, false, MCV>(ta, tb, p, stream);                     \
  }
  if (mc == 2) {
    if (BN == 256) { HCTR_DISPATCH(256, 2) }
    HCTR_DISPATCH(128, 2)
  }
  if (BN == 256) { HCTR_DISPATCH(256, 1) }
  if (BN == 64) { HCTR_DISPATCH(64, 1) }
  HCTR_DISPATCH(128, 1)
#undef HCTR_DISPATCH
}
