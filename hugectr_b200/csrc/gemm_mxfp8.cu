// D1-fp8: block-scaled (MX) fp8 GEMM on the 5th-generation tensor cores.
//   out[M, N] (bf16 / fp32) = epilogue( sum_k  (A_q[m, k] * 2^(SFA[m, k/32] - 127)) * (B_q[n, k] * 2^(SFB[n, k/32] - 127)) )
// A_q [M, K], B_q [N, K]: e4m3, K-major.  SFA / SFB: one UE8M0 scale per 32 consecutive K elements (OCP MX
// format), applied by the tensor core itself: `tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale` reads the
// scales from TENSOR MEMORY next to the accumulator.  Pipeline per CTA (one 128 x 128 output tile):
//   warp 0   TMA producer: A / B tiles (128 rows x 128 bytes, SWIZZLE_128B) + the 512-byte scale blocks of the
//            k-block (1-D bulk copies) into a 4-stage ring, completion on the stage's mbarrier
//   warp 1   one elected thread: `tcgen05.cp.32x128b.warpx4` moves the two scale blocks smem -> TMEM, then four
//            K = 32 MMAs with scale-factor id 0..3 (the byte of the 32-bit TMEM column that holds the scale of this
//            32-wide slice); tcgen05.commit releases the stage / publishes the accumulator
//   warps 4-7 epilogue: tcgen05.ld 32 x 32 columns -> bias / ReLU / ... -> global
// Scale layout in global memory = the layout the copy needs, so a k-block's scales are ONE contiguous 512-byte
// block per 128 rows:  block (row / 128, k / 128), byte offset (row % 32) * 16 + ((row % 128) / 32) * 4 + (k % 128) / 32
// (row r of the tile lands in TMEM lane r % 32 -- replicated to the 4 sub-partitions --, 32-bit column (r % 128) / 32).
// The quantiser (mx_quantize_kernel) writes exactly this; it is the reference's cublasLt path replaced
// (HugeCTR/src/layers/functors/fused_gemm_functors.cu:21-289 has no fp8 at all).
#include <cuda_fp8.h>

#include "gemm_epilogue.cuh"

namespace hctr {

constexpr int MX_BM = 128, MX_BN = 128, MX_BK = 128, MX_STAGES = 4, MX_THREADS = 256;
constexpr int MX_TILE_BYTES = MX_BM * MX_BK;          // 16 KB (fp8)
constexpr int MX_SF_BYTES = 512;
constexpr int MX_SMEM = MX_STAGES * (2 * MX_TILE_BYTES + 2 * MX_SF_BYTES) + 1024 + 1024;

HCTR_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor without swizzle: 8-row x 16-byte core matrices, rows 16 bytes apart
HCTR_DEVICE uint64_t make_smem_desc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
HCTR_DEVICE void utccp_32x128b_warpx4(uint32_t tmem_dst, uint64_t smem_desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}
HCTR_DEVICE void umma_mxf8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate,
                           uint32_t sfa_tmem, uint32_t sfb_tmem) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
      : "memory");
}
// block-scaled instruction descriptor (kind::mxf8f6f4): [4,6) b_sf_id  [7,10) a_format  [10,13) b_format
// [15] a_major  [16] b_major  [17,23) N >> 3  [23] scale format (1 = UE8M0)  [24,29) M >> 4  [29,31) a_sf_id
__host__ __device__ constexpr uint32_t make_idesc_mx(uint32_t M, uint32_t N) {
  return (kFmtE4M3 << 7) | (kFmtE4M3 << 10) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(MX_THREADS, 1)
    gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + MX_STAGES * MX_TILE_BYTES;
  uint8_t* smem_sfa = smem + 2 * MX_STAGES * MX_TILE_BYTES;
  uint8_t* smem_sfb = smem_sfa + MX_STAGES * MX_SF_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_sfb + MX_STAGES * MX_SF_BYTES);
  uint64_t* empty_bar = full_bar + MX_STAGES;
  uint64_t* tmem_full = empty_bar + MX_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int k_blocks = p.k_blocks;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp_idx == 1 && elect_one()) {
    for (int i = 0; i < MX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr, 256);       // 128 accumulator columns + 2 x 4 scale columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_sfa = tmem_base + 128, tmem_sfb = tmem_base + 132;

  if (warp_idx == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], 2 * MX_TILE_BYTES + 2 * MX_SF_BYTES);
        tma_load_2d(smem_a + stage * MX_TILE_BYTES, &tmA, &full_bar[stage], kb * MX_BK, m_blk * MX_BM);
        tma_load_2d(smem_b + stage * MX_TILE_BYTES, &tmB, &full_bar[stage], kb * MX_BK, n_blk * MX_BN);
        bulk_load_1d(smem_sfa + stage * MX_SF_BYTES,
                     sfa + (static_cast<long long>(m_blk) * k_blocks + kb) * MX_SF_BYTES, MX_SF_BYTES, &full_bar[stage]);
        bulk_load_1d(smem_sfb + stage * MX_SF_BYTES,
                     sfb + (static_cast<long long>(n_blk) * k_blocks + kb) * MX_SF_BYTES, MX_SF_BYTES, &full_bar[stage]);
        if (++stage == MX_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_idx == 1) {
    constexpr uint32_t idesc0 = make_idesc_mx(MX_BM, MX_BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        utccp_32x128b_warpx4(tmem_sfa, make_smem_desc_none(smem_u32(smem_sfa + stage * MX_SF_BYTES), 16, 128));
        utccp_32x128b_warpx4(tmem_sfb, make_smem_desc_none(smem_u32(smem_sfb + stage * MX_SF_BYTES), 16, 128));
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * MX_TILE_BYTES), 16, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * MX_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < MX_BK / 32; ++k) {
          const uint32_t idesc = idesc0 | (static_cast<uint32_t>(k) << 4) | (static_cast<uint32_t>(k) << 29);
          umma_mxf8(tmem_base, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u, tmem_sfa, tmem_sfb);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == k_blocks - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == MX_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp_idx >= 4) {
    const int ew = warp_idx - 4;
    const int lane = threadIdx.x & 31;
    const int m = m_blk * MX_BM + ew * 32 + lane;
    const bool row_ok = m < p.M;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < MX_BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + c * 32 + (static_cast<uint32_t>(ew * 32) << 16), r);
      tmem_ld_wait();
      epilogue_chunk(p, p.flags, m, row_ok, n_blk * MX_BN + c * 32, r);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------ MX quantiser
// x [rows, cols] (bf16 / fp32, row-major, cols % 128 == 0) -> q [rows_pad, cols] e4m3 + scales in the block layout
// above (rows_pad = rows rounded up to 128; padding rows are zero with scale 2^-127).  One thread per 32-wide block.
template <typename T>
__global__ void __launch_bounds__(256)
    mx_quantize_kernel(const T* __restrict__ x, long long ldx, uint8_t* __restrict__ q, uint8_t* __restrict__ sf,
                       int rows, int rows_pad, int cols, int transposed) {
  const int blocks_per_row = cols / 32;
  const long long total = static_cast<long long>(rows_pad) * blocks_per_row;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / blocks_per_row);
    const int cb = static_cast<int>(i - static_cast<long long>(r) * blocks_per_row);
    float v[32];
    float amax = 0.f;
    if (r < rows) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = static_cast<float>(transposed ? x[static_cast<long long>(cb * 32 + j) * ldx + r]
                                             : x[static_cast<long long>(r) * ldx + cb * 32 + j]);
        amax = fmaxf(amax, fabsf(v[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    // power-of-two scale: smallest 2^e with amax / 2^e <= 448 (e4m3 max)
    // (frexp instead of log2: exact and independent of --use_fast_math; v = m * 2^ex, m in [0.5, 1))
    int e = -127;
    if (amax > 0.f) {
      int ex;
      const float m = frexpf(amax * (1.f / 448.f), &ex);
      e = (m > 0.5f) ? ex : ex - 1;
      e = max(-127, min(127, e));
    }
    const float inv = ldexpf(1.f, -e);
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j] * inv, v[4 * j + 1] * inv),
                                                                  __NV_SATFINITE, __NV_E4M3);
      const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j + 2] * inv, v[4 * j + 3] * inv),
                                                                  __NV_SATFINITE, __NV_E4M3);
      w[j] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(q + static_cast<long long>(r) * cols + cb * 32);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    const int kb = cb >> 2, k = cb & 3;
    const long long blk = static_cast<long long>(r >> 7) * (cols / 128) + kb;
    sf[blk * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + k] = static_cast<uint8_t>(e + 127);
  }
}

typedef CUresult (*PFN_encodeTiledMx)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tmap_u8(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_bytes) {
  static PFN_encodeTiledMx enc = nullptr;
  if (enc == nullptr) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || !fp) return -1;
    enc = reinterpret_cast<PFN_encodeTiledMx>(fp);
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {128, 128};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace hctr

using namespace hctr;

// x: bf16 (is_bf16) or fp32 [rows, cols], row pitch ldx elements (transposed: stored [cols, rows], pitch ldx);
// q: [rows_pad, cols] bytes; sf: (rows_pad/128)*(cols/128)*512 bytes
extern "C" int hctr_mx_quantize(const void* x, long long ldx, void* q, void* sf, int rows, int cols, int is_bf16,
                                int transposed, void* stream) {
  if (cols % 128) return -2;
  const int rows_pad = (rows + 127) / 128 * 128;
  const long long total = static_cast<long long>(rows_pad) * (cols / 32);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (is_bf16)
    mx_quantize_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<uint8_t*>(q), reinterpret_cast<uint8_t*>(sf),
        rows, rows_pad, cols, transposed);
  else
    mx_quantize_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        reinterpret_cast<const float*>(x), ldx, reinterpret_cast<uint8_t*>(q), reinterpret_cast<uint8_t*>(sf), rows,
        rows_pad, cols, transposed);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// A_q [M_pad, K], B_q [N_pad, K] e4m3 (row pitch K bytes), scales in block layout; out [M, N] bf16 (or fp32 with
// EPI_OUT_F32), bias fp32 [N] optional, flags: EPI_RELU / EPI_SIGMOID / EPI_OUT_F32
extern "C" int hctr_gemm_mxfp8(const void* Aq, const void* sfa, const void* Bq, const void* sfb, void* out, int M,
                               int N, int K, long long ldo, const float* bias, float alpha, int flags,
                               void* stream_) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % 128) return -2;
  const int m_pad = (M + 127) / 128 * 128, n_pad = (N + 127) / 128 * 128;
  CUtensorMap ta, tb;
  if (make_tmap_u8(&ta, Aq, K, m_pad, K)) return -3;
  if (make_tmap_u8(&tb, Bq, K, n_pad, K)) return -4;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = m_pad / 128; p.n_tiles = n_pad / 128; p.k_blocks = K / 128; p.splits = 1; p.kb_per_split = p.k_blocks;
  p.out = out; p.ldo = ldo; p.bias = bias; p.alpha = alpha; p.flags = flags;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM) != cudaSuccess)
      return -5;
    attr = true;
  }
  gemm_mxfp8_kernel<<<dim3(p.n_tiles, p.m_tiles), MX_THREADS, MX_SMEM, reinterpret_cast<cudaStream_t>(stream_)>>>(
      ta, tb, reinterpret_cast<const uint8_t*>(sfa), reinterpret_cast<const uint8_t*>(sfb), p);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
