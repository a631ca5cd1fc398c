// Native host-side parameter-server index and row movers (host part of C35 / HPS).
// The Python HostParameterServer keeps the weight / optimizer-state rows in (pinned) host tensors;
// this file owns the key -> row index (sharded open hash maps, lock per shard) and the bulk row
// gather / scatter between the host table and staging buffers, both OpenMP-parallel, so that a
// pull / push of a training batch's unique keys does not go through a Python dict row by row.
// (reference: HugeCTR/include/embedding_training_cache/parameter_server.hpp,
//  hmem_cache/hmem_cache.hpp -- headers only in the reference tree.)
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

namespace {

constexpr int kShards = 64;

struct PsIndex {
  std::unordered_map<long long, long long> map[kShards];
  std::mutex mu[kShards];
  long long next_row{0};
  static int shard_of(long long k) {
    unsigned long long x = static_cast<unsigned long long>(k) * 0x9E3779B97F4A7C15ull;
    return static_cast<int>(x >> 58);  // top 6 bits -> 64 shards
  }
};

}  // namespace

extern "C" {

void* hctr_ps_create() { return new PsIndex(); }
void hctr_ps_destroy(void* h) { delete static_cast<PsIndex*>(h); }
long long hctr_ps_size(void* h) { return static_cast<PsIndex*>(h)->next_row; }

// rows_out[i] = row of keys[i] (-1 when absent and !create).  is_new[i] = 1 for the FIRST occurrence
// of a key that was inserted by this call (its row must be initialised by the caller).  Row ids are
// assigned in key order of first appearance, i.e. deterministically.  Returns the number of new rows,
// or -1 when `capacity` would be exceeded (nothing is inserted in that case).
long long hctr_ps_lookup(void* h, const long long* keys, long long n, long long* rows_out,
                         unsigned char* is_new, int create, long long capacity) {
  PsIndex* ps = static_cast<PsIndex*>(h);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    const int s = PsIndex::shard_of(keys[i]);
    // read phase: the maps are not modified concurrently with this loop
    auto it = ps->map[s].find(keys[i]);
    rows_out[i] = it == ps->map[s].end() ? -1 : it->second;
    if (is_new) is_new[i] = 0;
  }
  if (!create) return 0;
  long long missing = 0;
  for (long long i = 0; i < n; ++i)
    if (rows_out[i] < 0) ++missing;   // upper bound (duplicates counted more than once)
  if (missing == 0) return 0;
  long long created = 0;
  const long long first = ps->next_row;
  for (long long i = 0; i < n; ++i) {
    if (rows_out[i] >= 0) continue;
    const int s = PsIndex::shard_of(keys[i]);
    auto ins = ps->map[s].emplace(keys[i], ps->next_row);
    if (ins.second) {
      if (capacity > 0 && ps->next_row >= capacity) {
        // roll back everything inserted by this call
        ps->map[s].erase(ins.first);
        for (long long j = 0; j < i; ++j)
          if (is_new && is_new[j]) ps->map[PsIndex::shard_of(keys[j])].erase(keys[j]);
        ps->next_row = first;
        return -1;
      }
      ++ps->next_row;
      ++created;
      if (is_new) is_new[i] = 1;
    }
    rows_out[i] = ins.first->second;
  }
  return created;
}

// every (key, row) pair, in unspecified order; returns the count
long long hctr_ps_dump(void* h, long long* keys_out, long long* rows_out) {
  PsIndex* ps = static_cast<PsIndex*>(h);
  long long c = 0;
  for (int s = 0; s < kShards; ++s)
    for (auto& kv : ps->map[s]) {
      keys_out[c] = kv.first;
      rows_out[c] = kv.second;
      ++c;
    }
  return c;
}

// out[i, :] = table[rows[i], :]   (row_bytes each; rows[i] < 0 -> zeros)
void hctr_ps_gather(const char* table, const long long* rows, long long n, long long row_bytes,
                    char* out) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    if (rows[i] >= 0) memcpy(out + i * row_bytes, table + rows[i] * row_bytes, row_bytes);
    else memset(out + i * row_bytes, 0, row_bytes);
  }
}

// table[rows[i], :] = in[i, :]   (rows must be distinct for a defined result)
void hctr_ps_scatter(char* table, const long long* rows, long long n, long long row_bytes,
                     const char* in) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i)
    if (rows[i] >= 0) memcpy(table + rows[i] * row_bytes, in + i * row_bytes, row_bytes);
}

// table[rows[i], :] = U(-bound, bound), counter-based: the value of a cell depends on (seed, key, column) only --
// independent of the thread count and of the order in which rows are created.  (The embedding initializer
// of rows seen for the first time; a single-threaded generator made first-epoch steps creation-bound.)
void hctr_ps_init_rows(float* table, const long long* rows, const long long* keys, long long n, int ev, float bound,
                       unsigned long long seed) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    if (rows[i] < 0) continue;
    float* dst = table + rows[i] * static_cast<long long>(ev);
    // keyed by the KEY (not the row it happened to get): a key's first-sight vector does not depend on arrival order,
    // sharding or a resume in between
    unsigned long long x = seed ^ (static_cast<unsigned long long>(keys ? keys[i] : rows[i]) * 0x9E3779B97F4A7C15ull);
    for (int c = 0; c < ev; ++c) {
      x += 0x9E3779B97F4A7C15ull;                       // splitmix64 stream per row
      unsigned long long z = x;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      const float u = static_cast<float>(z >> 40) * (1.0f / 16777216.0f);     // 24 random bits -> [0, 1)
      dst[c] = (2.0f * u - 1.0f) * bound;
    }
  }
}

}  // extern "C"
