// Requester-side halves of the two embedding exchanges, issued as posted peer STORES over NVLink:
//
//   forward   emb_dispatch_kernel (keys)   every rank scatters the key blocks of its batch into the
//             inbox of the rank(s) that own the table: whole blocks for table-wise / column-wise
//             placement; for a table that is row-sharded k ways the bag is split on the fly -- each
//             owner receives only ITS compacted list of local row indices (key / k for key % k == s)
//             plus the list lengths.  This is the "sparse-key all-to-all dispatch" of the reference
//             (data_distributor: label keys by owner, count, sort by GPU, NCCL all2all of counts and
//             keys with a host sync -- HugeCTR/embedding/data_distributor/sparse_data_distribution_op_impl.cu:213-395,
//             key_filtering_operators.cu:37-710) collapsed into one kernel without counts exchange or
//             host sync: the inbox has a fixed slot per (source rank, lookup), sized for the worst case.
//   backward  the same kernel (generic 2-D block routes) pushes the top-gradient rows of every lookup to
//             the owner's gradient inbox -- the backward all-to-all (NetworkBackward + NCCL all2all,
//             HugeCTR/embedding/model_parallel_embedding.cpp:240-304) -- straight out of the buffer the
//             dense backward wrote.
//
// After either kernel one device barrier makes the inboxes visible; the owner-side gather/pool and
// index/reduce/update kernels (embedding.cu, embedding_bwd.cu) then touch only LOCAL memory, apart from
// the pooled-vector stores into the requesters' activation tensors.  No NVLink load latency is exposed
// anywhere on the path (remote traffic = posted writes).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "embedding.cuh"

namespace hctr {

struct DispatchRoute {
  long long src_off;   // element offset of the block in the local source buffer
  long long dst_off;   // element offset inside the destination rank's inbox (source slot included)
  long long nnz_off;   // split routes: element offset of the list lengths in the dest nnz inbox
  int rows;            // rows of the block (copy: 1 for a contiguous run)
  int row_elems;       // elements per row (split: hotness)
  int src_stride;      // elements between rows in the source
  int dst_stride;      // elements between rows in the destination
  int dst_rank;
  int kind;            // 0 = copy, 1 = split (keep key % k == shard, write key / k, pad -1, + nnz)
  int k;
  int shard;
};

struct DispatchDst {
  void* data[kMaxRanks];   // inbox base per destination rank (peer mapped)
  int* nnz[kMaxRanks];     // nnz inbox base per destination rank
};

template <typename T>
__global__ void __launch_bounds__(256)
    emb_dispatch_kernel(const T* __restrict__ src, const DispatchRoute* __restrict__ routes,
                        const DispatchDst dst) {
  const DispatchRoute r = routes[blockIdx.y];
  T* out = reinterpret_cast<T*>(dst.data[r.dst_rank]) + r.dst_off;
  const T* in = src + r.src_off;
  if (r.kind == 0) {
    constexpr int EPV = 16 / static_cast<int>(sizeof(T));     // elements per 16-byte vector
    const bool vec = (r.row_elems % EPV == 0) && (r.src_stride % EPV == 0) && (r.dst_stride % EPV == 0) &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) {
      const int vpr = r.row_elems / EPV;
      const long long total = static_cast<long long>(r.rows) * vpr;
      for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
           i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / vpr;
        const int c = static_cast<int>(i - row * vpr);
        const int4 v = ld_nc_v4(reinterpret_cast<const int4*>(in + row * r.src_stride) + c);
        reinterpret_cast<int4*>(out + row * r.dst_stride)[c] = v;
      }
    } else {
      const long long total = static_cast<long long>(r.rows) * r.row_elems;
      for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
           i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / r.row_elems;
        const int c = static_cast<int>(i - row * r.row_elems);
        out[row * r.dst_stride + c] = in[row * r.src_stride + c];
      }
    }
    return;
  }
  // split: one warp per sample
  if constexpr (sizeof(T) == 4 || sizeof(T) == 8) {
    using KeyT = typename std::conditional<sizeof(T) == 8, long long, int>::type;
    const KeyT* kin = reinterpret_cast<const KeyT*>(in);
    KeyT* kout = reinterpret_cast<KeyT*>(out);
    int* nnz = dst.nnz[r.dst_rank] + r.nnz_off;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const unsigned int k = static_cast<unsigned int>(r.k);
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < r.rows; s += warps) {
      const KeyT* kb = kin + static_cast<long long>(s) * r.src_stride;
      KeyT* ob = kout + static_cast<long long>(s) * r.dst_stride;
      int pos = 0;
      for (int h0 = 0; h0 < r.row_elems; h0 += 32) {
        const int h = h0 + lane;
        long long key = -1;
        if (h < r.row_elems) key = static_cast<long long>(kb[h]);
        bool match = false;
        long long q = 0;
        if (key >= 0) {
          const unsigned long long uk = static_cast<unsigned long long>(key);
          if (uk <= 0xFFFFFFFFull) {
            const unsigned int k32 = static_cast<unsigned int>(uk);
            const unsigned int qq = k32 / k;
            match = (k32 - qq * k) == static_cast<unsigned int>(r.shard);
            q = qq;
          } else {
            const unsigned long long qq = uk / k;
            match = (uk - qq * k) == static_cast<unsigned long long>(r.shard);
            q = static_cast<long long>(qq);
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, match);
        if (match) ob[pos + __popc(bal & ((1u << lane) - 1u))] = static_cast<KeyT>(q);
        pos += __popc(bal);
      }
      for (int h = pos + lane; h < r.row_elems; h += 32) ob[h] = static_cast<KeyT>(-1);
      if (lane == 0) nnz[s] = pos;
    }
  }
}

}  // namespace hctr

using namespace hctr;

extern "C" int hctr_abi_size_dispatch_route() { return static_cast<int>(sizeof(DispatchRoute)); }

// elem_bytes: 2 (bf16 gradients), 4 (fp32 gradients / int32 keys), 8 (int64 keys)
extern "C" int hctr_emb_dispatch(const void* src, const void* routes_dev, int num_routes,
                                 void* const* dst_data, void* const* dst_nnz, int num_ranks,
                                 int elem_bytes, int blocks_x, void* stream_) {
  if (num_routes <= 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  DispatchDst d;
  for (int i = 0; i < kMaxRanks; ++i) {
    d.data[i] = i < num_ranks ? dst_data[i] : nullptr;
    d.nnz[i] = (i < num_ranks && dst_nnz) ? reinterpret_cast<int*>(dst_nnz[i]) : nullptr;
  }
  const dim3 grid(blocks_x > 0 ? blocks_x : 8, num_routes);
  const DispatchRoute* r = reinterpret_cast<const DispatchRoute*>(routes_dev);
  if (elem_bytes == 2)
    emb_dispatch_kernel<unsigned short><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(src), r, d);
  else if (elem_bytes == 4)
    emb_dispatch_kernel<int><<<grid, 256, 0, st>>>(reinterpret_cast<const int*>(src), r, d);
  else if (elem_bytes == 8)
    emb_dispatch_kernel<long long><<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(src), r, d);
  else
    return -2;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
