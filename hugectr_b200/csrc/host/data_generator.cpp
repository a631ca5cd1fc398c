// Native synthetic data set writer (host side of C14) for the Norm and Raw formats.
// Every record draws from its own counter-based stream (splitmix64 keyed by seed, file id and record
// index), so the bytes written do not depend on the number of threads.  Keys are uniform or follow the
// reference's inverse-CDF power law over [1, vocab + 1) with pdf ~ x^-alpha
// (HugeCTR/include/data_generator.hpp:109-131).  Formats: Norm = DataSetHeader + variable-length
// records with an optional per-record checksum (data_generator.hpp:137-330), Raw = one file of fixed
// records [label][dense][keys] (data_generator.hpp:978-1070).
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed, uint64_t a, uint64_t b) {
    s = seed;
    s = next() ^ (a * 0x9E3779B97F4A7C15ull);
    s = next() ^ (b * 0xD1B54A32D192ED03ull);
    next();
  }
  inline uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  inline double uniform() { return static_cast<double>(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// per-slot sampler with the vocabulary-dependent constants hoisted out of the per-key path
struct KeyDist {
  bool power_law{false};
  bool log_case{false};
  long long vocab{1};
  double span{0.0};    // (V+1)^(1-alpha) - 1, or log(V+1) when alpha == 1
  double inv_a{1.0};   // 1 / (1 - alpha)
  KeyDist() = default;
  KeyDist(bool pl, double alpha, long long v) : power_law(pl), vocab(v) {
    if (!pl || v <= 1) return;
    if (fabs(alpha - 1.0) < 1e-9) {
      log_case = true;
      span = log(static_cast<double>(v) + 1.0);
    } else {
      const double a = 1.0 - alpha;
      span = pow(static_cast<double>(v) + 1.0, a) - 1.0;
      inv_a = 1.0 / a;
    }
  }
  inline long long draw(Rng& r) const {
    if (vocab <= 1) return 0;
    if (!power_law) return static_cast<long long>(r.next() % static_cast<uint64_t>(vocab));
    const double u = r.uniform();
    const double x = log_case ? exp(u * span) : exp(log(span * u + 1.0) * inv_a);
    long long k = static_cast<long long>(x) - 1;
    if (k < 0) k = 0;
    if (k > vocab - 1) k = vocab - 1;
    return k;
  }
};

std::vector<KeyDist> make_dists(int num_slot, const long long* sizes, int power_law, double alpha) {
  std::vector<KeyDist> d;
  for (int s = 0; s < num_slot; ++s) d.emplace_back(power_law != 0, alpha, sizes[s]);
  return d;
}

inline void put(std::vector<char>& b, const void* p, size_t n) {
  const char* c = static_cast<const char*>(p);
  b.insert(b.end(), c, c + n);
}

bool write_all(int fd, const char* p, size_t n, long long off) {
  while (n) {
    ssize_t w = pwrite(fd, p, n, off);
    if (w <= 0) return false;
    p += w; n -= static_cast<size_t>(w); off += w;
  }
  return true;
}

}  // namespace

// One Norm file.  nnz[s] == 1 -> exactly one key, else a uniform count in [1, nnz[s]].  Returns 0 / -1.
extern "C" int hctr_gen_norm_file(const char* path, unsigned long long seed, long long file_id,
                                  long long n, int label_dim, int dense_dim, int num_slot,
                                  const long long* slot_sizes, const int* nnz, int key_bytes, int check,
                                  int power_law, double alpha) {
  int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return -1;
  const std::vector<KeyDist> kd = make_dists(num_slot, slot_sizes, power_law, alpha);
  std::vector<char> out;
  out.reserve(1 << 22);
  long long hdr[8] = {check ? 1 : 0, n, label_dim, dense_dim, num_slot, 0, 0, 0};
  put(out, hdr, 64);
  long long off = 0;
  std::vector<char> rec;
  bool ok = true;
  for (long long i = 0; i < n && ok; ++i) {
    Rng r(seed, static_cast<uint64_t>(file_id), static_cast<uint64_t>(i));
    rec.clear();
    for (int j = 0; j < label_dim; ++j) { float v = r.uniform() < 0.5 ? 0.f : 1.f; put(rec, &v, 4); }
    for (int j = 0; j < dense_dim; ++j) { float v = static_cast<float>(r.uniform()); put(rec, &v, 4); }
    for (int s = 0; s < num_slot; ++s) {
      const int m = nnz ? nnz[s] : 1;
      int32_t c = m <= 1 ? 1 : 1 + static_cast<int32_t>(r.next() % static_cast<uint64_t>(m));
      put(rec, &c, 4);
      for (int h = 0; h < c; ++h) {
        const long long k = kd[s].draw(r);
        if (key_bytes == 4) { uint32_t t = static_cast<uint32_t>(k); put(rec, &t, 4); }
        else put(rec, &k, 8);
      }
    }
    if (check) {
      int32_t nb = static_cast<int32_t>(rec.size());
      int8_t sum = 0;
      for (char ch : rec) sum = static_cast<int8_t>(sum + static_cast<int8_t>(ch));
      put(out, &nb, 4);
      put(out, rec.data(), rec.size());
      put(out, &sum, 1);
    } else {
      put(out, rec.data(), rec.size());
    }
    if (out.size() >= (1u << 22)) {
      ok = write_all(fd, out.data(), out.size(), off);
      off += static_cast<long long>(out.size());
      out.clear();
    }
  }
  if (ok && !out.empty()) ok = write_all(fd, out.data(), out.size(), off);
  close(fd);
  return ok ? 0 : -1;
}

// One Raw file of num_samples fixed records, written by num_threads workers in 8 Ki-record chunks.
// float_label_dense: labels/dense are f32; otherwise labels are i32 {0,1} and dense i32 in [0, 100).
extern "C" int hctr_gen_raw_file(const char* path, unsigned long long seed, long long num_samples,
                                 int label_dim, int dense_dim, int num_slot,
                                 const long long* slot_sizes, const int* nnz, int key_bytes,
                                 int float_label_dense, int power_law, double alpha, int num_threads) {
  int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return -1;
  long long hot = 0;
  for (int s = 0; s < num_slot; ++s) hot += nnz ? nnz[s] : 1;
  const long long rec_bytes = 4ll * (label_dim + dense_dim) + hot * key_bytes;
  if (ftruncate(fd, rec_bytes * num_samples) != 0) { close(fd); return -1; }
  const std::vector<KeyDist> kd = make_dists(num_slot, slot_sizes, power_law, alpha);
  const long long chunk = 1 << 13;
  const long long nchunks = (num_samples + chunk - 1) / chunk;
  std::atomic<long long> next{0};
  std::atomic<bool> ok{true};
  auto work = [&] {
    std::vector<char> buf(static_cast<size_t>(chunk * rec_bytes));
    for (;;) {
      const long long c = next.fetch_add(1);
      if (c >= nchunks || !ok.load()) return;
      const long long lo = c * chunk, hi = lo + chunk < num_samples ? lo + chunk : num_samples;
      char* p = buf.data();
      for (long long i = lo; i < hi; ++i) {
        Rng r(seed, 0x5241575full, static_cast<uint64_t>(i));
        for (int j = 0; j < label_dim; ++j) {
          const bool one = r.uniform() >= 0.5;
          if (float_label_dense) { float v = one ? 1.f : 0.f; memcpy(p, &v, 4); }
          else { int32_t v = one ? 1 : 0; memcpy(p, &v, 4); }
          p += 4;
        }
        for (int j = 0; j < dense_dim; ++j) {
          const double u = r.uniform();
          if (float_label_dense) { float v = static_cast<float>(u); memcpy(p, &v, 4); }
          else { int32_t v = static_cast<int32_t>(u * 100.0); memcpy(p, &v, 4); }
          p += 4;
        }
        for (int s = 0; s < num_slot; ++s) {
          const int m = nnz ? nnz[s] : 1;
          for (int h = 0; h < m; ++h) {
            const long long k = kd[s].draw(r);
            if (key_bytes == 4) { uint32_t t = static_cast<uint32_t>(k); memcpy(p, &t, 4); p += 4; }
            else { memcpy(p, &k, 8); p += 8; }
          }
        }
      }
      if (!write_all(fd, buf.data(), static_cast<size_t>((hi - lo) * rec_bytes), lo * rec_bytes))
        ok.store(false);
    }
  };
  const int nt = num_threads < 1 ? 1 : (num_threads > 64 ? 64 : num_threads);
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  close(fd);
  return ok.load() ? 0 : -1;
}
