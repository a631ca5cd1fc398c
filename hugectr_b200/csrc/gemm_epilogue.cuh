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
  float* colsum;   // optional fp32 [N]: += column sums of the (bf16-rounded) output -- the bias gradient
                   // of the layer that consumes this GEMM's output as its dY (TMA epilogue only)
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

// ---------------------------------------------------------------------------------------------
// TMA-staged epilogue (gemm_tc2.cu).  The direct epilogue above is one thread per row: every 16-byte
// global access of a warp touches 32 different cache lines, so operand-reading / large-N epilogues
// are bound by L1TEX wavefronts, not by DRAM.  Here every epilogue warp owns a [32 rows x 64 cols]
// bf16 slab in shared memory (128-byte rows, SWIZZLE_128B): outputs are written to the slab and
// leave through cp.async.bulk.tensor stores, the x0 / xl / mask operands arrive in such slabs through
// TMA loads issued one tile ahead.  Out-of-range rows / columns are clipped (stores) or zero filled
// (loads) by the TMA unit, so there is no ragged-edge code at all.
//   EK_GENERIC: out = act(alpha * acc + bias)                       act in {id, relu, sigmoid}
//   EK_CROSS:   aux = alpha * acc + bias ; out = x0 * aux + xl      (aux store optional)
//   EK_ADD:     out = alpha * acc + bias + xl
//   EK_MASK:    out = (alpha * acc + bias) * (mask > 0)
enum EpiKind : int { EK_GENERIC = 0, EK_CROSS = 1, EK_ADD = 2, EK_MASK = 3 };

// host-side classification of a flag word; -1: not expressible by the TMA epilogue
inline int epi_kind_of(int flags, const void* addf) {
  if (flags & (EPI_OUT_F32 | EPI_ATOMIC | EPI_ACCUM)) return -1;
  if (addf != nullptr) return -1;
  const int op = flags & (EPI_CROSS | EPI_ADD | EPI_MASK);
  const int act = flags & (EPI_RELU | EPI_SIGMOID);
  if (op == 0) return EK_GENERIC;
  if (act) return -1;
  if (op == EPI_CROSS) return EK_CROSS;
  if (op == EPI_ADD) return EK_ADD;
  if (op == EPI_MASK) return EK_MASK;
  return -1;
}

// 16-byte chunk k (0..7) of row `row` inside a 128B-swizzled slab
HCTR_DEVICE uint32_t slab_off(int row, int k) { return row * 128 + ((k ^ (row & 7)) << 4); }

HCTR_DEVICE uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
HCTR_DEVICE void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One 32-column chunk: accumulators r[] -> slab `st_out` (and `st_aux` for the cross pre-activation),
// operands from slabs `ld_a` / `ld_b`.  `half` selects the 64-byte half of the 128-byte slab row.
template <int KIND>
HCTR_DEVICE void epilogue_chunk_tma(const GemmParams& p, const int flags, const int lane, const int n0,
                                    const int half, const uint32_t (&r)[32], const uint32_t st_out,
                                    const uint32_t st_aux, const uint32_t ld_a, const uint32_t ld_b) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  if (p.bias != nullptr) {
    if (n0 + 32 <= p.N) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
    }
  }
  if constexpr (KIND == EK_CROSS) {
    if (st_aux != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts_v4(st_aux + slab_off(lane, half * 4 + j),
               make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
    }
  }
  if constexpr (KIND != EK_GENERIC) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
      if constexpr (KIND == EK_CROSS || KIND == EK_MASK) a = lds_v4(ld_a + slab_off(lane, half * 4 + j));
      if constexpr (KIND == EK_CROSS || KIND == EK_ADD) b = lds_v4(ld_b + slab_off(lane, half * 4 + j));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float& lo = v[8 * j + 2 * q];
        float& hi = v[8 * j + 2 * q + 1];
        if constexpr (KIND == EK_CROSS) {
          lo = bf16_lo(aw[q]) * lo + bf16_lo(bw[q]);
          hi = bf16_hi(aw[q]) * hi + bf16_hi(bw[q]);
        } else if constexpr (KIND == EK_ADD) {
          lo += bf16_lo(bw[q]);
          hi += bf16_hi(bw[q]);
        } else {
          if (!(bf16_lo(aw[q]) > 0.f)) lo = 0.f;
          if (!(bf16_hi(aw[q]) > 0.f)) hi = 0.f;
        }
      }
    }
  } else {
    if (flags & EPI_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (flags & EPI_SIGMOID) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts_v4(st_out + slab_off(lane, half * 4 + j),
           make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                      pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
}

}  // namespace hctr
