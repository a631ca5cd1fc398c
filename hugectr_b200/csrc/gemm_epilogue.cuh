// Shared declarations + the fused epilogue of the tcgen05 GEMM kernels (gemm_tc.cu, gemm_tc2.cu).
#pragma once
#include "ptx.cuh"

namespace hctr {

enum EpiFlags : int {
  EPI_RELU = 1,
  EPI_OUT_F32 = 2,
  EPI_ATOMIC = 4,   // fp32 red.add into out (split-K / beta=1)
  EPI_ACCUM = 8,    // out(fp32) += acc, non atomic
  EPI_CROSS = 16,   // out = x0 * (acc + bias) + xl ; aux (optional) = acc + bias
  EPI_MASK = 32,    // out = acc * (mask > 0)
  EPI_SIGMOID = 64, // out = sigmoid(acc + bias)
  EPI_ADD = 128,    // out = acc + xl (bf16) [+ addf (fp32)]   (residual / gradient accumulation)
};

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split, k_blocks;
  void* out;
  long long ldo;
  void* aux;
  long long ldaux;
  const float* bias;
  const __nv_bfloat16* mask;
  long long ldmask;
  const __nv_bfloat16* x0;
  const __nv_bfloat16* xl;
  long long ldx;
  float alpha;
  int flags;
  const float* addf;
  long long ldaddf;
};


// One thread owns row `m`, columns [n0, n0+32) of the tile: r[] are the fp32 accumulators read from
// TMEM with tcgen05.ld.32x32b.x32.
HCTR_DEVICE void epilogue_chunk(const GemmParams& p, const int flags, const int m, const bool row_ok,
                                const int n0, const uint32_t (&r)[32]) {
        if (n0 >= p.N) return;  // warp-uniform
  const int nvalid = min(32, p.N - n0);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  if (p.bias != nullptr) {
    if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __ldg(p.bias + n0 + j);
    }
  }
  if (!row_ok) return;
  if (flags & EPI_CROSS) {
    // aux = acc + bias ; out = x0 * aux + xl
    const __nv_bfloat16* px0 = p.x0 + static_cast<long long>(m) * p.ldx + n0;
    const __nv_bfloat16* pxl = p.xl + static_cast<long long>(m) * p.ldx + n0;
    if (p.aux != nullptr && nvalid == 32) {
      uint4* pa = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.aux) +
                                           static_cast<long long>(m) * p.ldaux + n0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        pa[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]),
                           pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                           pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
    } else if (p.aux != nullptr) {
      __nv_bfloat16* pa = reinterpret_cast<__nv_bfloat16*>(p.aux) +
                          static_cast<long long>(m) * p.ldaux + n0;
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) pa[j] = __float2bfloat16(v[j]);
    }
    if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(px0) + j);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(pxl) + j);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[8 * j + 2 * q] = bf16_lo(aw[q]) * v[8 * j + 2 * q] + bf16_lo(bw[q]);
          v[8 * j + 2 * q + 1] = bf16_hi(aw[q]) * v[8 * j + 2 * q + 1] + bf16_hi(bw[q]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] = __bfloat162float(px0[j]) * v[j] + __bfloat162float(pxl[j]);
    }
  }
  if (flags & EPI_ADD) {
    const __nv_bfloat16* pxl = p.xl + static_cast<long long>(m) * p.ldx + n0;
    if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(pxl) + j);
        const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          v[8 * j + 2 * q] += bf16_lo(bw[q]);
          v[8 * j + 2 * q + 1] += bf16_hi(bw[q]);
        }
      }
      if (p.addf != nullptr) {
        const float* pf = p.addf + static_cast<long long>(m) * p.ldaddf + n0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(pf + j));
          v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) {
          v[j] += __bfloat162float(pxl[j]);
          if (p.addf != nullptr) v[j] += p.addf[static_cast<long long>(m) * p.ldaddf + n0 + j];
        }
    }
  }
  if (flags & EPI_MASK) {
    const __nv_bfloat16* pm = p.mask + static_cast<long long>(m) * p.ldmask + n0;
    if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(pm) + j);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!(bf16_lo(aw[q]) > 0.f)) v[8 * j + 2 * q] = 0.f;
          if (!(bf16_hi(aw[q]) > 0.f)) v[8 * j + 2 * q + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid && !(__bfloat162float(pm[j]) > 0.f)) v[j] = 0.f;
    }
  }
  if (flags & EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (flags & EPI_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
  }
  if (flags & (EPI_OUT_F32 | EPI_ATOMIC | EPI_ACCUM)) {
    float* po = reinterpret_cast<float*>(p.out) + static_cast<long long>(m) * p.ldo + n0;
    if (flags & EPI_ATOMIC) {
      if (nvalid == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(po + j),
                       "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
      } else {
        _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) atomicAdd(po + j, v[j]);
      }
    } else if (flags & EPI_ACCUM) {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) po[j] += v[j];
    } else if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(po + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) po[j] = v[j];
    }
  } else {
    __nv_bfloat16* po =
        reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldo + n0;
    if (nvalid == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reinterpret_cast<uint4*>(po)[j] =
            make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]),
                       pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                       pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (j < nvalid) po[j] = __float2bfloat16(v[j]);
    }
  }
}

}  // namespace hctr
