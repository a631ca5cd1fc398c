// Peer-memory runtime for the fused compute+collective kernels (NVLink 5 / NVSwitch):
//   * symmetric heap plumbing: cudaMalloc + CUDA IPC handles (one process per GPU)
//   * device-side all-GPU barrier through peer-mapped flag words (no host sync, CUDA-graph safe;
//     the role of HugeCTR/embedding/gpu_barrier/gpu_barrier.cu:23-70)
//   * in-place dense-gradient all-reduce: two-shot reduce-scatter + all-gather issued as P2P loads
//     and stores from ONE kernel, and a one-shot variant for small buffers
//     (reference custom kernel: HugeCTR/src/collectives/all_reduce_comm.cu:161-284)
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace hctr {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  void* p[kMaxPeers];
};

// Each rank owns flags[kMaxPeers] (uint32) in peer-mapped memory and a private epoch counter.
// Arrive: write the new epoch into slot [my_rank] of every peer's flag array (release.sys).
// Wait:   spin until every slot of MY array has reached the epoch (acquire.sys).
__global__ void barrier_kernel(PeerPtrs flags, uint32_t* epoch_ctr, int my_rank, int n) {
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = *epoch_ctr + 1;
    *epoch_ctr = epoch_s;
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  const int t = threadIdx.x;
  if (t < n) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(flags.p[t]) + my_rank, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[my_rank]) + t;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
}

// In-kernel barrier usable by a whole (co-resident) grid: block 0 does the cross-GPU handshake,
// other blocks wait on a local flag.  Used inside the fused all-reduce.
__device__ __forceinline__ void grid_peer_barrier(const PeerPtrs& flags, uint32_t* epoch_ctr,
                                                  uint32_t* local_gate, int my_rank, int n,
                                                  uint32_t& epoch) {
  // all blocks arrive on local_gate[0]; last one performs the peer barrier and opens local_gate[1]
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t prev = atomicAdd(&local_gate[0], 1u);
    if (prev == gridDim.x - 1) {
      local_gate[0] = 0;   // re-armed before anyone is released (grid sizes may differ per launch)
      const uint32_t e = *epoch_ctr + 1;
      *epoch_ctr = e;
      for (int t = 0; t < n; ++t)
        st_release_sys(reinterpret_cast<uint32_t*>(flags.p[t]) + my_rank, e);
      for (int t = 0; t < n; ++t) {
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[my_rank]) + t;
        while (static_cast<int32_t>(ld_acquire_sys(mine) - e) < 0) {
        }
      }
      __threadfence_system();
      atomicExch(&local_gate[1], epoch + 1);
    } else {
      while (static_cast<int32_t>(ld_acquire_sys(&local_gate[1]) - (epoch + 1)) < 0) {
      }
    }
  }
  __syncthreads();
  ++epoch;
}

// Two-shot in-place all-reduce (sum) of fp32 buf[n] present on every rank at peers.p[r].
//   phase 1: rank r reduces slice r by loading 16-byte lines from all peers
//   phase 2: rank r stores the reduced slice into every peer (all-gather by stores)
// Barriers before (all inputs final) / between are peer barriers; a trailing one makes the result
// visible before the optimizer reads it.  Grid must be co-resident (<= #SM blocks).
__global__ void __launch_bounds__(512)
    allreduce_twoshot_kernel(PeerPtrs peers, PeerPtrs flags, uint32_t* epoch_ctr,
                             uint32_t* local_gate, uint32_t* gate_epoch, long long n, int my_rank,
                             int nranks) {
  uint32_t epoch = *gate_epoch;
  grid_peer_barrier(flags, epoch_ctr, local_gate, my_rank, nranks, epoch);
  const long long n4 = n / 4;  // n is padded to a multiple of 4 * nranks by the host
  const long long per = n4 / nranks;
  const long long lo = per * my_rank;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < per;
       i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int r = 0; r < nranks; ++r) {
      const int src = (my_rank + r) % nranks;  // stagger peers to spread link load
      const int4 v = ld_nc_v4(reinterpret_cast<const float4*>(peers.p[src]) + lo + i);
      acc.x += __int_as_float(v.x); acc.y += __int_as_float(v.y);
      acc.z += __int_as_float(v.z); acc.w += __int_as_float(v.w);
    }
    const int4 o = make_int4(__float_as_int(acc.x), __float_as_int(acc.y), __float_as_int(acc.z),
                             __float_as_int(acc.w));
#pragma unroll 8
    for (int r = 0; r < nranks; ++r) {
      const int dst = (my_rank + r) % nranks;
      st_na_v4(reinterpret_cast<float4*>(peers.p[dst]) + lo + i, o);
    }
  }
  grid_peer_barrier(flags, epoch_ctr, local_gate, my_rank, nranks, epoch);
  if (blockIdx.x == 0 && threadIdx.x == 0) *gate_epoch = epoch;
}

// bulk pull: dst(local)[i] = src(peer)[i]  -- 16-byte vectorised copy of `n16` lines per peer region
struct PullDesc {
  const void* src[kMaxPeers];
  void* dst[kMaxPeers];
  long long n16[kMaxPeers];
};
__global__ void __launch_bounds__(512) peer_pull_kernel(PullDesc d, int nranks) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int r = 0; r < nranks; ++r) {
    const int4* s = reinterpret_cast<const int4*>(d.src[r]);
    int4* o = reinterpret_cast<int4*>(d.dst[r]);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < d.n16[r];
         i += stride)
      o[i] = ld_nc_v4(s + i);
  }
}

}  // namespace hctr

using namespace hctr;

extern "C" void* hctr_ipc_alloc(long long bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, static_cast<size_t>(bytes)) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, static_cast<size_t>(bytes));
  return p;
}
extern "C" int hctr_ipc_free(void* p) { return cudaFree(p) == cudaSuccess ? 0 : -1; }
extern "C" int hctr_ipc_get_handle(void* p, void* out64) {
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) return -1;
  memcpy(out64, &h, sizeof(h));
  return 0;
}
extern "C" void* hctr_ipc_open(const void* handle64) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return nullptr;
  return p;
}
extern "C" int hctr_ipc_close(void* p) { return cudaIpcCloseMemHandle(p) == cudaSuccess ? 0 : -1; }
extern "C" const char* hctr_last_cuda_error() { return cudaGetErrorString(cudaGetLastError()); }

extern "C" int hctr_peer_barrier(void* const* flags, void* epoch_ctr, int my_rank, int n,
                                 void* stream) {
  PeerPtrs f;
  for (int i = 0; i < n; ++i) f.p[i] = flags[i];
  barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      f, reinterpret_cast<uint32_t*>(epoch_ctr), my_rank, n);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_allreduce_twoshot(void* const* bufs, void* const* flags, void* epoch_ctr,
                                      void* local_gate, void* gate_epoch, long long n, int my_rank,
                                      int nranks, int blocks, void* stream) {
  PeerPtrs b, f;
  for (int i = 0; i < nranks; ++i) {
    b.p[i] = bufs[i];
    f.p[i] = flags[i];
  }
  if (n % (4ll * nranks)) return -2;
  allreduce_twoshot_kernel<<<blocks, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      b, f, reinterpret_cast<uint32_t*>(epoch_ctr), reinterpret_cast<uint32_t*>(local_gate),
      reinterpret_cast<uint32_t*>(gate_epoch), n, my_rank, nranks);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int hctr_peer_pull(void* const* src, void* const* dst, const long long* n16, int nranks,
                              int blocks, void* stream) {
  PullDesc d;
  for (int i = 0; i < nranks; ++i) {
    d.src[i] = src[i];
    d.dst[i] = dst[i];
    d.n16[i] = n16[i];
  }
  peer_pull_kernel<<<blocks, 512, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d, nranks);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
