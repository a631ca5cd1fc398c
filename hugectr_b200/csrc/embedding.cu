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

#include "embedding.cuh"

namespace hctr {

// ------------------------------------------------------------------ forward
// One group of G lanes per (src rank, lookup, sample) bucket; each lane owns VEC consecutive
// floats per 32*VEC-wide column chunk.  G*VEC >= ev for ev <= 128 (VEC=4), wider rows loop.
template <typename KeyT, typename OutT, int VEC, int U>
__global__ void __launch_bounds__(256, 3)
    emb_fwd_kernel(const EmbParams p, const int G, const long long total_items) {
  // grid.y = (src rank, lookup): uniform per block, no per-thread integer division
  const int lane = threadIdx.x & 31;
  const int groups_per_warp = 32 / G;
  const int gl = lane % G;  // lane inside group
  const int gi = lane / G;  // group inside warp
  const int l = blockIdx.y % p.num_lookups;
  const int src = blockIdx.y / p.num_lookups;
  const int s = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * groups_per_warp + gi;
  const bool active = s < p.batch;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gi * G));
  (void)total_items;
  const EmbLookup lk = p.lookups[l];
  const int ev = lk.ev_size;
  int nnz = lk.hotness;
  const KeyT* kbase = reinterpret_cast<const KeyT*>(p.keys[src]) + lk.key_off +
                      static_cast<long long>(s) * lk.key_stride;
  if (active && lk.nnz_off >= 0) nnz = min(lk.hotness, p.nnz[src][lk.nnz_off + s]);
  if (!active) nnz = 0;

  constexpr int MAXC = (U >= 8) ? 8 : 4;  // chunks per lane: ev <= MAXC * G * VEC
  float acc[MAXC][VEC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[c][v] = 0.f;
  const int nchunk = (ev + G * VEC - 1) / (G * VEC);

  // gathers U rows (rp[u] == nullptr: skip) into the accumulators
  auto gather = [&](const float* const (&rp)[U]) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < nchunk) {
        const int col = (c * G + gl) * VEC;
        if (col < ev) {
          if constexpr (VEC == 4) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
              v[u] = rp[u] ? __ldg(reinterpret_cast<const float4*>(rp[u] + col))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              acc[c][0] += v[u].x; acc[c][1] += v[u].y; acc[c][2] += v[u].z; acc[c][3] += v[u].w;
            }
          } else {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = rp[u] ? __ldg(rp[u] + col) : 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) acc[c][0] += v[u];
          }
        }
      }
    }
  };

  if (lk.num_shards > 1) {
    // Row-sharded table (block-uniform branch): this shard owns ~1/num_shards of the bag, so the
    // owned keys are compacted first -- every lane fetches up to 4 keys, a group ballot marks the
    // owned ones, and rows are gathered U at a time from the compacted set.  (Gathering in key order
    // would issue one row per round for an 8-way shard: a chain of dependent DRAM latencies.)
    const int KPL = (G >= 32) ? 1 : ((G >= 16) ? 2 : 4);   // keys per lane and round; KPL * G <= 32
    const unsigned int ns = static_cast<unsigned int>(lk.num_shards);
    for (int h0 = 0; h0 < nnz; h0 += KPL * G) {
      unsigned int rowv[4] = {0u, 0u, 0u, 0u};
      unsigned int m = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < KPL) {
          const int h = h0 + i * G + gl;
          bool ok = false;
          if (h < nnz) {
            const long long key = static_cast<long long>(kbase[h]);
            if (key >= 0) {
              const unsigned long long uk = static_cast<unsigned long long>(key);
              unsigned long long q;
              unsigned int rem;
              if (uk <= 0xFFFFFFFFull) {        // 32-bit fast path (no 64-bit division)
                const unsigned int k32 = static_cast<unsigned int>(uk);
                q = k32 / ns;
                rem = k32 - static_cast<unsigned int>(q) * ns;
              } else {
                q = uk / ns;
                rem = static_cast<unsigned int>(uk - q * ns);
              }
              ok = rem == static_cast<unsigned int>(lk.shard_idx) && q < static_cast<unsigned long long>(lk.rows);
              rowv[i] = static_cast<unsigned int>(q);
            }
          }
          const unsigned int bal = (__ballot_sync(gmask, ok) >> (gi * G)) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
          m |= bal << (i * G);
        }
      }
      while (m) {               // group-uniform
        const float* rp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          rp[u] = nullptr;
          const bool take = m != 0u;
          const int b = take ? (__ffs(m) - 1) : 0;
          if (take) m &= m - 1u;
          const int i = b / G, sl = b - i * G;
          const unsigned int cand = i == 0 ? rowv[0] : (i == 1 ? rowv[1] : (i == 2 ? rowv[2] : rowv[3]));
          const unsigned int r = __shfl_sync(gmask, cand, gi * G + sl);
          if (take) rp[u] = p.table + (lk.table_row_off + r) * static_cast<long long>(p.ev_size);
        }
        gather(rp);
      }
    }
  } else {
    for (int h0 = 0; h0 < nnz; h0 += G) {
      // lanes of the group fetch up to G keys at once (coalesced peer/local read)
      long long mykey = -1;
      if (h0 + gl < nnz) mykey = static_cast<long long>(kbase[h0 + gl]);
      const int cnt = min(G, nnz - h0);
      for (int j0 = 0; j0 < cnt; j0 += U) {
        const float* rp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u;
          const long long key = __shfl_sync(gmask, mykey, gi * G + (j < cnt ? j : 0));
          const bool ok = (j < cnt) && key >= 0 && key < lk.rows;
          rp[u] = ok ? p.table + (lk.table_row_off + key) * static_cast<long long>(p.ev_size) : nullptr;
        }
        gather(rp);
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

// Data-parallel tables (small, replicated): dense wgrad [rows, pitch] += scattered bag gradients of
// the LOCAL batch.  Hot rows of tiny tables would serialise global reductions on a handful of L2
// lines, so a block (one lookup x one chunk of samples) first accumulates into a shared-memory copy
// of the table when it fits and flushes it with one vector reduction per non-zero 16-byte group.
template <typename KeyT, typename GradT>
__global__ void __launch_bounds__(256)
    emb_dp_wgrad_kernel(const EmbParams p, float* __restrict__ wgrad, const float grad_scale,
                        const int smem_floats) {
  extern __shared__ float dp_acc[];
  const EmbLookup lk = p.lookups[blockIdx.y];
  const int ev = lk.ev_size, pitch = p.ev_size;
  const long long tbl_floats = static_cast<long long>(lk.rows) * ev;
  const bool priv = tbl_floats <= smem_floats;
  const int per = (p.batch + gridDim.x - 1) / gridDim.x;
  const int s0 = blockIdx.x * per, s1 = min(p.batch, s0 + per);
  if (priv) {
    for (int i = threadIdx.x; i < tbl_floats; i += blockDim.x) dp_acc[i] = 0.f;
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const KeyT* keys = reinterpret_cast<const KeyT*>(p.keys[0]) + lk.key_off;
  const GradT* grads = reinterpret_cast<const GradT*>(p.grad[0]) + lk.grad_off;
  for (int s = s0 + warp; s < s1; s += 8) {
    int nnz = lk.hotness;
    if (lk.nnz_off >= 0) nnz = min(lk.hotness, p.nnz[0][lk.nnz_off + s]);
    if (nnz <= 0) continue;
    float scale = grad_scale;
    if (lk.combiner == 1) scale /= static_cast<float>(nnz);
    const KeyT* kb = keys + static_cast<long long>(s) * lk.key_stride;
    const GradT* gp = grads + static_cast<long long>(s) * lk.grad_stride;
    for (int col = lane * 4; col < ev; col += 128) {
      float4 v = load_vec4<GradT>(gp + col);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      for (int h = 0; h < nnz; ++h) {
        const long long key = static_cast<long long>(kb[h]);
        if (key < 0 || key >= lk.rows) continue;
        if (priv) {
          float* d = dp_acc + key * ev + col;
          atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
        } else {
          float* d = wgrad + (lk.table_row_off + key) * pitch + col;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v.x), "f"(v.y),
                       "f"(v.z), "f"(v.w) : "memory");
        }
      }
    }
  }
  if (priv) {
    __syncthreads();
    for (long long i = threadIdx.x * 4ll; i < tbl_floats; i += blockDim.x * 4ll) {
      const float4 v = *reinterpret_cast<const float4*>(dp_acc + i);
      if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
      const long long r = i / ev;
      const int c = static_cast<int>(i - r * ev);
      float* d = wgrad + (lk.table_row_off + r) * pitch + c;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w) : "memory");
    }
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
        apply_opt4<OPT, StateT>(w, make_float4(g.x * inv_scaler, g.y * inv_scaler, g.z * inv_scaler, g.w * inv_scaler), s0, s1, base + col, hp, lr, bc1, bc2);
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
  int G = pick_group(max_ev, vec);
  // wide rows: 8 lanes per bucket, every lane owns up to 4 float4 chunks of the row -> 4 buckets
  // per warp and 16 independent 16-byte loads in flight per lane (latency-bound -> bandwidth-bound)
  const bool narrow = (vec == 4 && max_ev >= 64 && max_ev <= 128);
  if (narrow) G = 8;
  const long long items = static_cast<long long>(p->num_ranks) * p->num_lookups * p->batch;
  if (items == 0) return 0;
  const int threads = 256;
  const int buckets_per_block = (threads / 32) * (32 / G);
  const dim3 blocks((p->batch + buckets_per_block - 1) / buckets_per_block,
                    p->num_ranks * p->num_lookups);
#define LAUNCH(K, O, V, UU) \
  emb_fwd_kernel<K, O, V, UU><<<blocks, threads, 0, st>>>(*p, G, items)
#define LAUNCH_U(K, O, V)            \
  if (narrow) { LAUNCH(K, O, V, 4); } \
  else { LAUNCH(K, O, V, 8); }
  if (vec == 4) {
    if (key_bytes == 8) { if (out_bf16) { LAUNCH_U(long long, __nv_bfloat16, 4) } else { LAUNCH_U(long long, float, 4) } }
    else { if (out_bf16) { LAUNCH_U(int, __nv_bfloat16, 4) } else { LAUNCH_U(int, float, 4) } }
  } else {
    if (key_bytes == 8) { if (out_bf16) { LAUNCH(long long, __nv_bfloat16, 1, 8); } else { LAUNCH(long long, float, 1, 8); } }
    else { if (out_bf16) { LAUNCH(int, __nv_bfloat16, 1, 8); } else { LAUNCH(int, float, 1, 8); } }
  }
#undef LAUNCH_U
#undef LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Requester-side split of a row-sharded bag: one warp per sample compacts, for every shard j, the
// local row indices (key / k) of the keys with key % k == j (original order), pads with -1 and
// writes the list length.  out: [k, batch, hotness], nnz: [k, batch].
template <typename KeyT>
__global__ void __launch_bounds__(256)
    emb_shard_split_kernel(const KeyT* __restrict__ keys, KeyT* __restrict__ out, int* __restrict__ nnz,
                           const int batch, const int hotness, const int k) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < batch; s += warps) {
    const KeyT* kb = keys + static_cast<long long>(s) * hotness;
    for (int j = 0; j < k; ++j) {
      KeyT* ob = out + (static_cast<long long>(j) * batch + s) * hotness;
      int pos = 0;
      for (int h0 = 0; h0 < hotness; h0 += 32) {
        const int h = h0 + lane;
        long long key = -1;
        if (h < hotness) key = static_cast<long long>(kb[h]);
        const bool match = key >= 0 && (key % k) == j;
        const unsigned bal = __ballot_sync(0xffffffffu, match);
        if (match) ob[pos + __popc(bal & ((1u << lane) - 1u))] = static_cast<KeyT>(key / k);
        pos += __popc(bal);
      }
      for (int h = pos + lane; h < hotness; h += 32) ob[h] = static_cast<KeyT>(-1);
      if (lane == 0) nnz[j * batch + s] = pos;
    }
  }
}

extern "C" int hctr_emb_shard_split(const void* keys, void* out, int* nnz, int batch, int hotness,
                                    int k, int key_bytes, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (batch <= 0 || hotness <= 0 || k <= 1) return 0;
  const int blocks = (batch * 32 + 255) / 256;
  if (key_bytes == 8)
    emb_shard_split_kernel<long long><<<blocks, 256, 0, st>>>(
        reinterpret_cast<const long long*>(keys), reinterpret_cast<long long*>(out), nnz, batch, hotness, k);
  else
    emb_shard_split_kernel<int><<<blocks, 256, 0, st>>>(reinterpret_cast<const int*>(keys),
                                                        reinterpret_cast<int*>(out), nnz, batch, hotness, k);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_emb_backward_accum(const EmbParams* p, const UniqueTable* ut,
                                       float* wgrad_unique, float grad_scale, int max_ev,
                                       int key_bytes, int grad_bf16, int dp_ev4, void* stream_) {
  // dp_ev4: every lookup's ev_size is a multiple of 4 and its gradient rows are 16-byte aligned
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int vec = (max_ev % 4 == 0) ? 4 : 1;
  const int G = pick_group(max_ev, vec);
  const long long items = static_cast<long long>(p->num_ranks) * p->num_lookups * p->batch;
  if (items == 0) return 0;
  if (ut->keys == nullptr && p->num_ranks == 1 && dp_ev4 && p->ev_size % 4 == 0) {
    // dense wgrad of replicated tables: privatised block-level accumulation
    constexpr int kSmem = 96 * 1024;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(emb_dp_wgrad_kernel<int, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      cudaFuncSetAttribute(emb_dp_wgrad_kernel<int, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      cudaFuncSetAttribute(emb_dp_wgrad_kernel<long long, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      cudaFuncSetAttribute(emb_dp_wgrad_kernel<long long, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      attr_set = true;
    }
    int chunks = (2 * 148 + p->num_lookups - 1) / p->num_lookups;
    if (chunks > 64) chunks = 64;
    if (chunks * 8 > p->batch) chunks = (p->batch + 7) / 8;
    if (chunks < 1) chunks = 1;
    const dim3 grid(chunks, p->num_lookups);
#define LAUNCH_DP(K, O) \
  emb_dp_wgrad_kernel<K, O><<<grid, 256, kSmem, st>>>(*p, wgrad_unique, grad_scale, kSmem / 4)
    if (key_bytes == 8) { if (grad_bf16) LAUNCH_DP(long long, __nv_bfloat16); else LAUNCH_DP(long long, float); }
    else { if (grad_bf16) LAUNCH_DP(int, __nv_bfloat16); else LAUNCH_DP(int, float); }
#undef LAUNCH_DP
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
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
