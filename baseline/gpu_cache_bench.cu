// Same-box kernel baseline: the UNMODIFIED reference gpu_cache library (baseline/_ref/libgpu_cache.so, built
// from /root/reference/gpu_cache/src by baseline/build_reference.py) against our cache kernels
// (hugectr_b200/lib/libhctr_cuda.so: hctr_cache_query / hctr_cache_replace) on one workload:
// capacity C rows of 128 floats, batches of n keys drawn from a power-law, Query then Replace of the misses.
// Prints one JSON object.  Build + run: python baseline/gpu_cache_bench.py   (GPU box)
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <vector>

#include <nv_gpu_cache.hpp>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

using RefCache = gpu_cache::gpu_cache<long long, uint64_t, std::numeric_limits<long long>::max(), SET_ASSOCIATIVITY, SLAB_SIZE>;

typedef int (*q_fn)(void*, void*, void*, void*, void*, int, int, int, const void*, long long, void*, void*, void*, void*, void*);
typedef int (*r_fn)(void*, void*, void*, void*, void*, int, int, int, const void*, const void*, long long, void*, void*, void*, void*);

int main(int argc, char** argv) {
  const int ev = 128;
  const size_t cap_rows = 1 << 20;            // 1 Mi rows = 512 MB of vectors
  const size_t n = 1 << 18;                   // keys per batch
  const int iters = 20;
  const long long vocab = 40000000;
  std::mt19937_64 rng(1);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  const double a = 1.0 - 1.1, hi = std::pow(double(vocab + 1), a);
  std::vector<std::vector<long long>> batches(iters + 5, std::vector<long long>(n));
  for (auto& b : batches) {
    for (auto& k : b) k = std::min<long long>(vocab - 1, (long long)std::floor(std::pow((hi - 1.0) * U(rng) + 1.0, 1.0 / a)) - 1);
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());       // the caches are queried with unique keys
  }
  long long* d_keys; float* d_vals; uint64_t* d_mi; long long* d_mk; size_t* d_ml; float* d_src;
  CK(cudaMalloc(&d_keys, n * 8)); CK(cudaMalloc(&d_vals, n * ev * 4)); CK(cudaMalloc(&d_mi, n * 8));
  CK(cudaMalloc(&d_mk, n * 8)); CK(cudaMalloc(&d_ml, 8)); CK(cudaMalloc(&d_src, n * ev * 4));
  CK(cudaMemset(d_src, 0, n * ev * 4));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);

  // ---------------- reference
  double ref_q = 0, ref_r = 0; size_t ref_hits = 0, ref_tot = 0;
  {
    RefCache cache(cap_rows / (SET_ASSOCIATIVITY * SLAB_SIZE), ev);
    for (size_t it = 0; it < batches.size(); ++it) {
      const size_t len = batches[it].size();
      CK(cudaMemcpyAsync(d_keys, batches[it].data(), len * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemsetAsync(d_ml, 0, 8, st));
      cudaEventRecord(e0, st);
      cache.Query(d_keys, len, d_vals, d_mi, d_mk, d_ml, st);
      cudaEventRecord(e1, st);
      size_t ml = 0; CK(cudaMemcpyAsync(&ml, d_ml, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 5) { ref_q += ms; ref_hits += len - ml; ref_tot += len; }
      cudaEventRecord(e0, st);
      if (ml) cache.Replace(d_mk, ml, d_src, st);
      cudaEventRecord(e1, st);
      CK(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 5) ref_r += ms;
    }
  }
  // ---------------- ours (same keys, same order)
  double our_q = 0, our_r = 0; size_t our_hits = 0, our_tot = 0;
  {
    void* h = dlopen(argc > 1 ? argv[1] : "hugectr_b200/lib/libhctr_cuda.so", RTLD_NOW);
    if (!h) { printf("{\"error\": \"dlopen: %s\"}\n", dlerror()); return 1; }
    q_fn query = (q_fn)dlsym(h, "hctr_cache_query");
    r_fn replace = (r_fn)dlsym(h, "hctr_cache_replace");
    const int ways = 64, sets = cap_rows / ways;
    long long *keys, *stamps, *clock, *cnt; float* vals; int* locks;
    CK(cudaMalloc(&keys, cap_rows * 8)); CK(cudaMemset(keys, 0xFF, cap_rows * 8));
    CK(cudaMalloc(&stamps, cap_rows * 8)); CK(cudaMemset(stamps, 0, cap_rows * 8));
    CK(cudaMalloc(&vals, cap_rows * ev * 4)); CK(cudaMalloc(&locks, sets * 4)); CK(cudaMemset(locks, 0, sets * 4));
    CK(cudaMalloc(&clock, 8)); CK(cudaMemset(clock, 0, 8)); CK(cudaMalloc(&cnt, 8));
    long long* d_mi2; CK(cudaMalloc(&d_mi2, n * 8));
    for (size_t it = 0; it < batches.size(); ++it) {
      const size_t len = batches[it].size();
      CK(cudaMemcpyAsync(d_keys, batches[it].data(), len * 8, cudaMemcpyHostToDevice, st));
      CK(cudaMemsetAsync(cnt, 0, 8, st));
      cudaEventRecord(e0, st);
      query(keys, stamps, vals, locks, clock, sets, ways, ev, d_keys, (long long)len, d_vals, d_mi2, d_mk, cnt, st);
      cudaEventRecord(e1, st);
      long long ml = 0; CK(cudaMemcpyAsync(&ml, cnt, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 5) { our_q += ms; our_hits += len - ml; our_tot += len; }
      cudaEventRecord(e0, st);
      if (ml) replace(keys, stamps, vals, locks, clock, sets, ways, ev, d_mk, d_src, ml, nullptr, nullptr, nullptr, st);
      cudaEventRecord(e1, st);
      CK(cudaStreamSynchronize(st));
      cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 5) our_r += ms;
    }
  }
  printf("{\"workload\": \"capacity %zu rows x %d fp32, unique power-law(1.1) keys of a %zu-key batch over %lld ids, %d timed batches\", "
         "\"reference_gpu_cache\": {\"query_ms\": %.4f, \"replace_ms\": %.4f, \"hit_rate\": %.4f}, "
         "\"hugectr_b200\": {\"query_ms\": %.4f, \"replace_ms\": %.4f, \"hit_rate\": %.4f}}\n",
         cap_rows, ev, n, vocab, iters, ref_q / iters, ref_r / iters, double(ref_hits) / ref_tot, our_q / iters,
         our_r / iters, double(our_hits) / our_tot);
  return 0;
}
