// Sparse hot path: model-parallel embedding lookup / pooling and fused backward + optimizer.
//
// Owner-side design (one process per GPU, tables sharded table-wise and/or row-wise):
//   forward  : ONE kernel on the owner pulls the requesters' keys straight out of their (peer
//              mapped) key buffers, gathers + pools the rows from local HBM and stores the pooled
//              vectors straight into the requester's activation tensor over NVLink.  This replaces
//              the reference chain  key all-to-all (X11/X12, with a host sync)  ->  lookup  ->
//              ModelForward -> NCCL all-to-all (X14) -> NetworkForward
//              (HugeCTR/embedding/model_parallel_embedding.cpp:213-231,
//               HugeCTR/embedding/operators/generic_lookup.cuh:319-716).
//   backward : gradient rows are read from the requesters' (peer) top-grad tensors, de-duplicated
//              with a transient open-addressing hash (no sort, no host sync) and accumulated in
//              fp32; a fused per-unique-row optimizer kernel then updates weights + state.
//              Replaces NetworkBackward + all-to-all + index calculation (CUB sorts) + LocalReduce
//              + update (HugeCTR/embedding/operators/index_calculation.cu:102-887,
//              HugeCTR/embedding_storage/ragged_static_embedding.cu:93-345).
// With num_ranks == 1 all "peer" pointers are local and the same kernels serve a single GPU.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hctr {

constexpr int kMaxRanks = 16;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned int kInvalidVal = 0xFFFFFFFFu;

struct EmbLookup {
  long long table_row_off;  // first row of this table shard inside the local arena
  long long key_off;        // element offset of the lookup's keys in a rank's feature-major buffer
  long long nnz_off;        // element offset of per-bucket nnz (variable hotness) or -1
  long long out_off;        // element offset of the lookup's output block inside the output slab
  long long grad_off;       // element offset of the lookup's top-grad block inside the grad slab
  int hotness;              // fixed / max hotness (keys read per bucket)
  int key_stride;           // keys per sample in the key block (>= hotness; concat combiner: H)
  int num_shards;           // row sharding: owned iff key % num_shards == shard_idx
  int shard_idx;
  int out_stride;           // elements per sample row of the output block
  int grad_stride;          // elements per sample row of the grad block
  int combiner;             // 0 = sum, 1 = mean
  int ev_size;
  int rows;                 // rows of the local shard (bounds guard)
  int pad_;
};

struct EmbParams {
  int num_ranks, my_rank, batch, num_lookups;
  const void* keys[kMaxRanks];   // per source rank: feature-major keys
  const int* nnz[kMaxRanks];     // per source rank: per-bucket nnz (optional)
  void* out[kMaxRanks];          // per source rank: output activation base
  const void* grad[kMaxRanks];   // per source rank: top-grad base
  const EmbLookup* lookups;      // device array [num_lookups]
  float* table;                  // [rows, ev] fp32 arena
  int ev_size;                   // row pitch of the arena (all tables of a group share it)
};

struct UniqueTable {
  unsigned long long* keys;  // [capacity] hash keys (arena row index)
  unsigned int* vals;        // [capacity] compact unique id
  unsigned int* counter;     // [1] number of unique rows
  unsigned long long* rows;  // [max_unique] uid -> arena row
  unsigned int* slots;       // [max_unique] uid -> hash slot (for O(n) clearing)
  unsigned int mask;         // capacity - 1
  unsigned int max_unique;
};

HCTR_DEVICE unsigned int hash64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<unsigned int>(k);
}

template <typename T>
HCTR_DEVICE void store_vec4(T* p, float a, float b, float c, float d);
template <>
HCTR_DEVICE void store_vec4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
HCTR_DEVICE void store_vec4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
template <typename T>
HCTR_DEVICE float4 load_vec4(const T* p);
template <>
HCTR_DEVICE float4 load_vec4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
HCTR_DEVICE float4 load_vec4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}
template <typename T>
HCTR_DEVICE float to_f(T v);
template <>
HCTR_DEVICE float to_f<float>(float v) { return v; }
template <>
HCTR_DEVICE float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
HCTR_DEVICE T from_f(float v);
template <>
HCTR_DEVICE float from_f<float>(float v) { return v; }
template <>
HCTR_DEVICE __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// ------------------------------------------------------------------ forward
// One group of G lanes per (src rank, lookup, sample) bucket; each lane owns VEC consecutive
// floats per 32*VEC-wide column chunk.  G*VEC >= ev for ev <= 128 (VEC=4), wider rows loop.
template <typename KeyT, typename OutT, int VEC>
__global__ void __launch_bounds__(256)
    emb_fwd_kernel(const EmbParams p, const int G, const long long total_items) {
  const int lane = threadIdx.x & 31;
  const int groups_per_warp = 32 / G;
  const int gl = lane % G;  // lane inside group
  const int gi = lane / G;  // group inside warp
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long item = warp_global * groups_per_warp + gi;
  const bool active = item < total_items;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gi * G));

  int s = 0, l = 0, src = 0;
  if (active) {
    s = static_cast<int>(item % p.batch);
    const long long t = item / p.batch;
    l = static_cast<int>(t % p.num_lookups);
    src = static_cast<int>(t / p.num_lookups);
  }
  const EmbLookup lk = p.lookups[l];
  const int ev = lk.ev_size;
  int nnz = lk.hotness;
  const KeyT* kbase = reinterpret_cast<const KeyT*>(p.keys[src]) + lk.key_off +
                      static_cast<long long>(s) * lk.key_stride;
  if (active && lk.nnz_off >= 0) nnz = min(lk.hotness, p.nnz[src][lk.nnz_off + s]);
  if (!active) nnz = 0;

  constexpr int MAXC = 8;  // up to ev = 8 * 32 * VEC (1024 for VEC=4) columns per lane
  float acc[MAXC][VEC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
  const int nchunk = (ev + G * VEC - 1) / (G * VEC);
  int owned = 0;

  for (int h0 = 0; h0 < nnz; h0 += G) {
    // lanes of the group fetch up to G keys at once (coalesced peer/local read)
    long long mykey = -1;
    if (h0 + gl < nnz) mykey = static_cast<long long>(kbase[h0 + gl]);
    const int cnt = min(G, nnz - h0);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const long long key = __shfl_sync(gmask, mykey, gi * G + j);
      if (key < 0) continue;
      if (lk.num_shards > 1 && (key % lk.num_shards) != lk.shard_idx) continue;
      long long r = key / lk.num_shards;
      if (r >= lk.rows) continue;  // out-of-vocabulary guard
      ++owned;
      const float* rowp = p.table + (lk.table_row_off + r) * static_cast<long long>(p.ev_size);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < nchunk) {
          const int col = (c * G + gl) * VEC;
          if (col < ev) {
            if constexpr (VEC == 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(rowp + col));
              acc[c][0] += v.x; acc[c][1] += v.y; acc[c][2] += v.z; acc[c][3] += v.w;
            } else {
              acc[c][0] += __ldg(rowp + col);
            }
          }
        }
      }
    }
  }
  if (!active) return;
  float scale = 1.f;
  if (lk.combiner == 1 && nnz > 0) scale = 1.f / static_cast<float>(nnz);
  OutT* o = reinterpret_cast<OutT*>(p.out[src]) + lk.out_off +
            static_cast<long long>(s) * lk.out_stride;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < nchunk) {
      const int col = (c * G + gl) * VEC;
      if (col < ev) {
        if constexpr (VEC == 4) {
          store_vec4<OutT>(o + col, acc[c][0] * scale, acc[c][1] * scale, acc[c][2] * scale,
                           acc[c][3] * scale);
        } else {
          o[col] = from_f<OutT>(acc[c][0] * scale);
        }
      }
    }
  }
  (void)owned;
}

// ------------------------------------------------------------------ unique (transient hash)
HCTR_DEVICE unsigned int unique_get_insert(const UniqueTable& t, unsigned long long row) {
  unsigned int h = hash64(row) & t.mask;
  while (true) {
    const unsigned long long prev = atomicCAS(&t.keys[h], kEmptyKey, row);
    if (prev == kEmptyKey) {
      const unsigned int uid = atomicAdd(t.counter, 1u);
      if (uid < t.max_unique) {
        t.rows[uid] = row;
        t.slots[uid] = h;
      }
      __threadfence();
      atomicExch(&t.vals[h], uid);
      return uid;
    }
    if (prev == row) {
      unsigned int v;
      do {
        v = *reinterpret_cast<volatile unsigned int*>(&t.vals[h]);
      } while (v == kInvalidVal);
      return v;
    }
    h = (h + 1) & t.mask;
  }
}

// ------------------------------------------------------------------ backward accumulate
// Same bucket decomposition as forward.  Each group loads its bucket's gradient row once and
// red.adds it (fp32) into wgrad_unique[uid] for every owned key of the bucket.
template <typename KeyT, typename GradT, int VEC>
__global__ void __launch_bounds__(256)
    emb_bwd_accum_kernel(const EmbParams p, const UniqueTable ut, float* __restrict__ wgrad_unique,
                         const float grad_scale, const int G, const long long total_items) {
  const int lane = threadIdx.x & 31;
  const int groups_per_warp = 32 / G;
  const int gl = lane % G;
  const int gi = lane / G;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long item = warp_global * groups_per_warp + gi;
  const bool active = item < total_items;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gi * G));
  int s = 0, l = 0, src = 0;
  if (active) {
    s = static_cast<int>(item % p.batch);
    const long long t = item / p.batch;
    l = static_cast<int>(t % p.num_lookups);
    src = static_cast<int>(t / p.num_lookups);
  }
  const EmbLookup lk = p.lookups[l];
  const int ev = lk.ev_size;
  int nnz = lk.hotness;
  const KeyT* kbase = reinterpret_cast<const KeyT*>(p.keys[src]) + lk.key_off +
                      static_cast<long long>(s) * lk.key_stride;
  if (active && lk.nnz_off >= 0) nnz = min(lk.hotness, p.nnz[src][lk.nnz_off + s]);
  if (!active) nnz = 0;

  constexpr int MAXC = 8;
  float g[MAXC][VEC];
  const int nchunk = (ev + G * VEC - 1) / (G * VEC);
  float scale = grad_scale;
  if (lk.combiner == 1 && nnz > 0) scale /= static_cast<float>(nnz);
  if (active) {
    const GradT* gp = reinterpret_cast<const GradT*>(p.grad[src]) + lk.grad_off +
                      static_cast<long long>(s) * lk.grad_stride;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < nchunk) {
        const int col = (c * G + gl) * VEC;
        if (col < ev) {
          if constexpr (VEC == 4) {
            const float4 v = load_vec4<GradT>(gp + col);
            g[c][0] = v.x * scale; g[c][1] = v.y * scale; g[c][2] = v.z * scale; g[c][3] = v.w * scale;
          } else {
            g[c][0] = to_f<GradT>(gp[col]) * scale;
          }
        }
      }
    }
  }
  for (int h0 = 0; h0 < nnz; h0 += G) {
    long long mykey = -1;
    if (h0 + gl < nnz) mykey = static_cast<long long>(kbase[h0 + gl]);
    // every lane resolves the unique id of ITS key (parallel hash probes), then broadcast
    unsigned int myuid = kInvalidVal;
    if (mykey >= 0 && (lk.num_shards == 1 || (mykey % lk.num_shards) == lk.shard_idx)) {
      const long long r = mykey / lk.num_shards;
      if (r < lk.rows) {
        const unsigned long long arow = static_cast<unsigned long long>(lk.table_row_off + r);
        // ut.keys == nullptr : dense wgrad [rows, ev] (data-parallel tables), uid == arena row
        myuid = ut.keys ? unique_get_insert(ut, arow) : static_cast<unsigned int>(arow);
      }
    }
    const int cnt = min(G, nnz - h0);
    for (int j = 0; j < cnt; ++j) {
      const unsigned int uid = __shfl_sync(gmask, myuid, gi * G + j);
      if (uid == kInvalidVal || uid >= ut.max_unique) continue;
      float* dst = wgrad_unique + static_cast<long long>(uid) * p.ev_size;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < nchunk) {
          const int col = (c * G + gl) * VEC;
          if (col < ev) {
            if constexpr (VEC == 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + col),
                           "f"(g[c][0]), "f"(g[c][1]), "f"(g[c][2]), "f"(g[c][3])
                           : "memory");
            } else {
              atomicAdd(dst + col, g[c][0]);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ fused sparse optimizers
enum SparseOpt : int { OPT_SGD = 0, OPT_ADAGRAD, OPT_ADAM, OPT_FTRL, OPT_MOMENTUM, OPT_NESTEROV,
                       OPT_RMSPROP };

struct OptHyper {
  const float* lr_ptr;    // device learning rate (graph friendly)
  float lr_scale;         // extra multiplier (1)
  float scaler;           // loss scaler: g /= scaler
  float beta1, beta2, epsilon;
  float lambda1, lambda2, ftrl_beta;
  float momentum;
  float initial_accu;
  const unsigned int* step_ptr;  // device step counter (Adam bias correction), 1-based
};

template <int OPT, typename StateT>
HCTR_DEVICE void apply_opt(float& w, float g, StateT* s0, StateT* s1, long long idx,
                           const OptHyper& hp, float lr, float bc1, float bc2) {
  if constexpr (OPT == OPT_SGD) {
    w -= lr * g;
  } else if constexpr (OPT == OPT_ADAGRAD) {
    float a = to_f<StateT>(s0[idx]) + g * g;
    s0[idx] = from_f<StateT>(a);
    w -= lr * g / (sqrtf(a) + hp.epsilon);
  } else if constexpr (OPT == OPT_ADAM) {
    float m = hp.beta1 * to_f<StateT>(s0[idx]) + (1.f - hp.beta1) * g;
    float v = hp.beta2 * to_f<StateT>(s1[idx]) + (1.f - hp.beta2) * g * g;
    s0[idx] = from_f<StateT>(m);
    s1[idx] = from_f<StateT>(v);
    const float alpha = lr * sqrtf(bc2) / bc1;
    w -= alpha * m / (sqrtf(v) + hp.epsilon);
  } else if constexpr (OPT == OPT_FTRL) {
    // state0 = z, state1 = n (reference ftrl_optimizer.cu:28-42)
    const float n = to_f<StateT>(s1[idx]);
    const float n_new = n + g * g;
    const float ef = hp.ftrl_beta;
    float z = to_f<StateT>(s0[idx]) + g + (sqrtf(n + ef) - sqrtf(n_new + ef)) * w / lr;
    s0[idx] = from_f<StateT>(z);
    s1[idx] = from_f<StateT>(n_new);
    const float p = (z > 0.f ? 1.f : -1.f) * hp.lambda1 - z;
    const float q = sqrtf(n_new + ef) / lr + hp.lambda2;
    w = (fabsf(z) > hp.lambda1) ? p / q : 0.f;
  } else if constexpr (OPT == OPT_MOMENTUM) {
    float m = hp.momentum * to_f<StateT>(s0[idx]) - lr * g;
    s0[idx] = from_f<StateT>(m);
    w += m;
  } else if constexpr (OPT == OPT_NESTEROV) {
    const float a = to_f<StateT>(s0[idx]);
    const float a_new = hp.momentum * a - lr * g;
    s0[idx] = from_f<StateT>(a_new);
    w += -hp.momentum * a + (1.f + hp.momentum) * a_new;
  } else if constexpr (OPT == OPT_RMSPROP) {
    float v = hp.beta2 * to_f<StateT>(s0[idx]) + (1.f - hp.beta2) * g * g;
    s0[idx] = from_f<StateT>(v);
    w -= lr * g / (sqrtf(v) + hp.epsilon);
  }
}

// One warp per unique row: reads the accumulated fp32 gradient (and zeroes it for the next step),
// updates weight + optimizer state in place and releases the row's hash slot.
template <int OPT, typename StateT>
__global__ void __launch_bounds__(256)
    emb_update_kernel(float* __restrict__ table, StateT* __restrict__ s0, StateT* __restrict__ s1,
                      float* __restrict__ wgrad_unique, const UniqueTable ut, const int ev,
                      const OptHyper hp) {
  const unsigned int n = min(*ut.counter, ut.max_unique);
  const float lr = (hp.lr_ptr ? *hp.lr_ptr : 1.f) * hp.lr_scale;
  float bc1 = 1.f, bc2 = 1.f;
  if constexpr (OPT == OPT_ADAM) {
    const float t = static_cast<float>(hp.step_ptr ? *hp.step_ptr : 1u);
    bc1 = 1.f - powf(hp.beta1, t);
    bc2 = 1.f - powf(hp.beta2, t);
  }
  const float inv_scaler = 1.f / hp.scaler;
  const int lane = threadIdx.x & 31;
  const unsigned int warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned int uid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; uid < n; uid += warps) {
    const unsigned long long row = ut.rows[uid];
    float* gsrc = wgrad_unique + static_cast<long long>(uid) * ev;
    const long long base = static_cast<long long>(row) * ev;
    if ((ev & 3) == 0) {
      for (int col = lane * 4; col < ev; col += 128) {
        float4 g = *reinterpret_cast<float4*>(gsrc + col);
        *reinterpret_cast<float4*>(gsrc + col) = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 w = *reinterpret_cast<float4*>(table + base + col);
        apply_opt<OPT, StateT>(w.x, g.x * inv_scaler, s0, s1, base + col, hp, lr, bc1, bc2);
        apply_opt<OPT, StateT>(w.y, g.y * inv_scaler, s0, s1, base + col + 1, hp, lr, bc1, bc2);
        apply_opt<OPT, StateT>(w.z, g.z * inv_scaler, s0, s1, base + col + 2, hp, lr, bc1, bc2);
        apply_opt<OPT, StateT>(w.w, g.w * inv_scaler, s0, s1, base + col + 3, hp, lr, bc1, bc2);
        *reinterpret_cast<float4*>(table + base + col) = w;
      }
    } else {
      for (int col = lane; col < ev; col += 32) {
        const float g = gsrc[col];
        gsrc[col] = 0.f;
        float w = table[base + col];
        apply_opt<OPT, StateT>(w, g * inv_scaler, s0, s1, base + col, hp, lr, bc1, bc2);
        table[base + col] = w;
      }
    }
    if (lane == 0) {
      const unsigned int slot = ut.slots[uid];
      ut.keys[slot] = kEmptyKey;
      ut.vals[slot] = kInvalidVal;
    }
  }
}

__global__ void emb_reset_counter_kernel(unsigned int* counter, unsigned int* overflow_flag,
                                         unsigned int max_unique) {
  if (*counter > max_unique) atomicMax(overflow_flag, *counter);
  *counter = 0;
}

// plain gather (no pooling): out[i] = table[rows[i]]  (concat combiner / dense lookups / SOK)
template <typename OutT>
__global__ void emb_gather_rows_kernel(const float* __restrict__ table, const long long* rows,
                                       OutT* out, long long n, int ev) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long r = rows[w];
  for (int c = lane; c < ev; c += 32)
    out[w * ev + c] = from_f<OutT>(r < 0 ? 0.f : table[r * ev + c]);
}

}  // namespace hctr

using namespace hctr;

static inline int pick_group(int ev, int vec) {
  int g = (ev + vec - 1) / vec;
  int p = 1;
  while (p < g && p < 32) p <<= 1;
  return p;
}

// key_bytes: 4 or 8 ; out_bf16 / grad_bf16 : element type of activations
extern "C" int hctr_emb_forward(const EmbParams* p, int max_ev, int key_bytes, int out_bf16,
                                void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int vec = (max_ev % 4 == 0) ? 4 : 1;
  const int G = pick_group(max_ev, vec);
  const long long items = static_cast<long long>(p->num_ranks) * p->num_lookups * p->batch;
  if (items == 0) return 0;
  const long long warps = (items + (32 / G) - 1) / (32 / G);
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
#define LAUNCH(K, O, V) emb_fwd_kernel<K, O, V><<<(unsigned)blocks, threads, 0, st>>>(*p, G, items)
  if (vec == 4) {
    if (key_bytes == 8) { if (out_bf16) LAUNCH(long long, __nv_bfloat16, 4); else LAUNCH(long long, float, 4); }
    else { if (out_bf16) LAUNCH(int, __nv_bfloat16, 4); else LAUNCH(int, float, 4); }
  } else {
    if (key_bytes == 8) { if (out_bf16) LAUNCH(long long, __nv_bfloat16, 1); else LAUNCH(long long, float, 1); }
    else { if (out_bf16) LAUNCH(int, __nv_bfloat16, 1); else LAUNCH(int, float, 1); }
  }
#undef LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_emb_backward_accum(const EmbParams* p, const UniqueTable* ut,
                                       float* wgrad_unique, float grad_scale, int max_ev,
                                       int key_bytes, int grad_bf16, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int vec = (max_ev % 4 == 0) ? 4 : 1;
  const int G = pick_group(max_ev, vec);
  const long long items = static_cast<long long>(p->num_ranks) * p->num_lookups * p->batch;
  if (items == 0) return 0;
  const long long warps = (items + (32 / G) - 1) / (32 / G);
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
#define LAUNCH(K, O, V) \
  emb_bwd_accum_kernel<K, O, V><<<(unsigned)blocks, threads, 0, st>>>(*p, *ut, wgrad_unique, grad_scale, G, items)
  if (vec == 4) {
    if (key_bytes == 8) { if (grad_bf16) LAUNCH(long long, __nv_bfloat16, 4); else LAUNCH(long long, float, 4); }
    else { if (grad_bf16) LAUNCH(int, __nv_bfloat16, 4); else LAUNCH(int, float, 4); }
  } else {
    if (key_bytes == 8) { if (grad_bf16) LAUNCH(long long, __nv_bfloat16, 1); else LAUNCH(long long, float, 1); }
    else { if (grad_bf16) LAUNCH(int, __nv_bfloat16, 1); else LAUNCH(int, float, 1); }
  }
#undef LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_emb_update(float* table, void* s0, void* s1, float* wgrad_unique,
                               const UniqueTable* ut, int ev, int opt, int state_bf16,
                               const OptHyper* hp, unsigned int* overflow_flag, int num_sms,
                               void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int threads = 256;
  const int blocks = num_sms * 8;
#define LAUNCH(O)                                                                             \
  if (state_bf16)                                                                             \
    emb_update_kernel<O, __nv_bfloat16><<<blocks, threads, 0, st>>>(                         \
        table, (__nv_bfloat16*)s0, (__nv_bfloat16*)s1, wgrad_unique, *ut, ev, *hp);          \
  else                                                                                        \
    emb_update_kernel<O, float><<<blocks, threads, 0, st>>>(table, (float*)s0, (float*)s1,   \
                                                            wgrad_unique, *ut, ev, *hp);
  switch (opt) {
    case OPT_SGD: LAUNCH(OPT_SGD); break;
    case OPT_ADAGRAD: LAUNCH(OPT_ADAGRAD); break;
    case OPT_ADAM: LAUNCH(OPT_ADAM); break;
    case OPT_FTRL: LAUNCH(OPT_FTRL); break;
    case OPT_MOMENTUM: LAUNCH(OPT_MOMENTUM); break;
    case OPT_NESTEROV: LAUNCH(OPT_NESTEROV); break;
    case OPT_RMSPROP: LAUNCH(OPT_RMSPROP); break;
    default: return -2;
  }
#undef LAUNCH
  emb_reset_counter_kernel<<<1, 1, 0, st>>>(ut->counter, overflow_flag, ut->max_unique);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_emb_gather_rows(const float* table, const long long* rows, void* out,
                                    long long n, int ev, int out_bf16, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (n == 0) return 0;
  const long long blocks = (n * 32 + 255) / 256;
  if (out_bf16)
    emb_gather_rows_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(table, rows, (__nv_bfloat16*)out, n, ev);
  else
    emb_gather_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(table, rows, (float*)out, n, ev);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
