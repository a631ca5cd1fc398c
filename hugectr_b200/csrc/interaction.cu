// D2: DLRM dot-product interaction on tcgen05 (fwd + bwd), bf16, n <= 32 rows, D <= 128, D % 16 == 0.
//   fwd : out[s] = [ mlp[s] | strict lower triangle of X X^T (row-major) | 0 ],  X = [mlp[s]; emb[s]]
//   bwd : dX = (G + G^T) X  with G the lower-triangular gradient matrix;  dmlp += dout[:, :D]
// Reference: HugeCTR/src/layers/interaction_layer.cu:47-596 (Volta wmma 16x16x16, one warp per
// sample).  Here FOUR samples share one UMMA tile: their (zero padded) 32-row blocks are stacked
// into a 128-row operand that sits in shared memory in the canonical SWIZZLE_128B layout; one
// 128x128xD tcgen05.mma produces all four 32x32 Gram blocks on the TMEM diagonal, the epilogue
// warps read their own block with tcgen05.ld and store the triangle.  The backward builds a
// block-diagonal (G+G^T) operand and multiplies it with the same X tile used as an MN-major B.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hctr {

using bf16 = __nv_bfloat16;
constexpr int kSPT = 4;          // samples per tile
[[maybe_unused]] constexpr int kRows = 32;  // padded rows per sample (layout constant, documents kSPT * kRows == 128)
constexpr int kHalfBytes = 128 * 128;  // one 64-column half of a 128-row tile

// byte offset of 16-byte chunk `c` (0..15 over D=128) of tile row `row` in the SW128 layout
HCTR_DEVICE uint32_t tile_off(int row, int c) {
  return static_cast<uint32_t>((c >> 3) * kHalfBytes + row * 128 + (((c & 7) ^ (row & 7)) << 4));
}

// cooperative fill of the stacked X tile for samples [s0, s0+4)
HCTR_DEVICE void fill_x_tile(uint8_t* smem_x, const bf16* __restrict__ mlp,
                             const bf16* __restrict__ emb, int s0, int B, int n, int D) {
  const int chunks = D >> 3;  // 16-byte chunks per row
  for (int t = threadIdx.x; t < 128 * 16; t += blockDim.x) {
    const int row = t >> 4, c = t & 15;
    const int i = row >> 5, j = row & 31;
    const int s = s0 + i;
    int4 v = make_int4(0, 0, 0, 0);
    if (c < chunks && s < B && j < n) {
      const bf16* src = (j == 0) ? mlp + static_cast<long long>(s) * D
                                 : emb + (static_cast<long long>(s) * (n - 1) + (j - 1)) * D;
      v = ld_nc_v4(reinterpret_cast<const int4*>(src) + c);
    }
    *reinterpret_cast<int4*>(smem_x + tile_off(row, c)) = v;
  }
}

__global__ void __launch_bounds__(256, 2)
    interaction_fwd_kernel(const bf16* __restrict__ mlp, const bf16* __restrict__ emb,
                           bf16* __restrict__ out, int B, int n, int D, int out_w) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_x = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 128);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, 0, 0, 128, 128);
  const int ntiles = (B + kSPT - 1) / kSPT;
  const int ntri = n * (n - 1) / 2;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int s0 = tile * kSPT;
    fill_x_tile(smem_x, mlp, emb, s0, B, n, D);
    fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    __syncthreads();
    if (warp == 1) {
      if (elect_one()) {
        const uint32_t base = smem_u32(smem_x);
        for (int k = 0; k < (D >> 4); ++k) {
          const uint32_t a = base + (k >> 2) * kHalfBytes + (k & 3) * 32;
          const uint64_t d = make_smem_desc_sw128(a, 16, 1024);
          umma_f16(tmem, d, d, idesc, k > 0 ? 1u : 0u);   // Z = X X^T : A and B are the same tile
        }
        umma_commit(&mma_bar);
      }
      __syncwarp();
    }
    // warps 4..7: passthrough of the bottom-MLP output + zero pad column (overlaps the MMA)
    if (warp >= 4) {
      const int i = warp - 4, s = s0 + i;
      if (s < B) {
        const int4* src = reinterpret_cast<const int4*>(mlp + static_cast<long long>(s) * D);
        bf16* o = out + static_cast<long long>(s) * out_w;
        if ((out_w & 7) == 0) {
          for (int c = lane; c < (D >> 3); c += 32) reinterpret_cast<int4*>(o)[c] = ld_nc_v4(src + c);
        } else {
          const bf16* sm = mlp + static_cast<long long>(s) * D;
          for (int c = lane; c < D; c += 32) o[c] = sm[c];
        }
        if (lane == 0) o[D + ntri] = __float2bfloat16(0.f);
      }
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (warp < 4) {
      // TMEM lanes 32*warp.. hold sample `warp`; its Gram block is at columns 32*warp..
      uint32_t r[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + warp * 32, r);
      tmem_ld_wait();
      const int s = s0 + warp, j = lane;
      if (s < B && j < n && j > 0) {
        bf16* o = out + static_cast<long long>(s) * out_w + D + j * (j - 1) / 2;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c < j) o[c] = __float2bfloat16(__uint_as_float(r[c]));
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM reads + smem tile consumed before the next tile overwrites them
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(256, 2)
    interaction_bwd_kernel(const bf16* __restrict__ mlp, const bf16* __restrict__ emb,
                           const bf16* __restrict__ dout, bf16* __restrict__ dmlp,
                           bf16* __restrict__ demb, int B, int n, int D, int out_w) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_x = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
  uint8_t* smem_g = smem_x + 2 * kHalfBytes;   // block-diagonal (G + G^T), K-major, 2 halves
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 128);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // D[128, N=D] = A[128, K=128] (K-major) * B (MN-major: rows = K, 64-wide N chunks)
  const uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, 0, 1, 128, static_cast<uint32_t>(D));
  const int ntiles = (B + kSPT - 1) / kSPT;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int s0 = tile * kSPT;
    fill_x_tile(smem_x, mlp, emb, s0, B, n, D);
    // zero the G tile, then write the four 32x32 diagonal blocks
    for (int t = threadIdx.x; t < 2 * kHalfBytes / 16; t += blockDim.x)
      reinterpret_cast<int4*>(smem_g)[t] = make_int4(0, 0, 0, 0);
    __syncthreads();
    for (int t = threadIdx.x; t < 128 * 4; t += blockDim.x) {
      const int row = t >> 2, cc = t & 3;       // cc: 8-column chunk inside the sample's block
      const int i = row >> 5, r = row & 31, s = s0 + i;
      if (s >= B || r >= n) continue;
      const bf16* g = dout + static_cast<long long>(s) * out_w + D;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = cc * 8 + q * 2 + e;
          float x = 0.f;
          if (c < n && c != r) {
            const int hi = max(r, c), lo = min(r, c);
            x = __bfloat162float(g[hi * (hi - 1) / 2 + lo]);
          }
          v[e] = x;
        }
        w[q] = pack_bf16x2(v[0], v[1]);
      }
      const int col_chunk = i * 4 + cc;          // 16-byte chunk index along K (0..15)
      *reinterpret_cast<int4*>(smem_g + tile_off(row, col_chunk)) =
          make_int4(static_cast<int>(w[0]), static_cast<int>(w[1]), static_cast<int>(w[2]),
                    static_cast<int>(w[3]));
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 1) {
      if (elect_one()) {
        const uint32_t ga = smem_u32(smem_g), xb = smem_u32(smem_x);
        for (int k = 0; k < 8; ++k) {            // K = 128 stacked rows
          const uint64_t a = make_smem_desc_sw128(ga + (k >> 2) * kHalfBytes + (k & 3) * 32, 16, 1024);
          const uint64_t b = make_smem_desc_sw128(xb + k * 2048, kHalfBytes, 1024);
          umma_f16(tmem, a, b, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (warp < 4) {
      const int s = s0 + warp, j = lane;
      const bool ok = s < B && j < n;
      bf16* dst = nullptr;
      const bf16* pass = nullptr;
      if (ok) {
        if (j == 0) {
          dst = dmlp + static_cast<long long>(s) * D;
          pass = dout + static_cast<long long>(s) * out_w;
        } else {
          dst = demb + (static_cast<long long>(s) * (n - 1) + (j - 1)) * D;
        }
      }
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = __uint_as_float(r[q * 8 + e]);
              if (pass) v[e] += __bfloat162float(pass[c0 + q * 8 + e]);
            }
            if (c0 + q * 8 < D)
              *reinterpret_cast<uint4*>(dst + c0 + q * 8) =
                  make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                             pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace hctr

using namespace hctr;

extern "C" int hctr_interaction_fwd(const void* mlp, const void* emb, void* out, int B, int n, int D,
                                    void* stream) {
  if (n > 32 || D > 128 || (D & 15)) return -2;
  static bool attr = false;
  const int smem = 2 * kHalfBytes + 1024;
  if (!attr) {
    cudaFuncSetAttribute(interaction_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int out_w = D + n * (n - 1) / 2 + 1;
  const int tiles = (B + kSPT - 1) / kSPT;
  const int grid = tiles < 296 ? tiles : 296;
  interaction_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const bf16*)mlp, (const bf16*)emb, (bf16*)out, B, n, D, out_w);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_interaction_bwd(const void* mlp, const void* emb, const void* dout, void* dmlp,
                                    void* demb, int B, int n, int D, void* stream) {
  if (n > 32 || D > 128 || (D & 15)) return -2;
  static bool attr = false;
  const int smem = 4 * kHalfBytes + 1024;
  if (!attr) {
    cudaFuncSetAttribute(interaction_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int out_w = D + n * (n - 1) / 2 + 1;
  const int tiles = (B + kSPT - 1) / kSPT;
  const int grid = tiles < 296 ? tiles : 296;
  interaction_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const bf16*)mlp, (const bf16*)emb, (const bf16*)dout, (bf16*)dmlp, (bf16*)demb, B, n, D, out_w);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
