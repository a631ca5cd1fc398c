// Native Criteo TSV -> Raw binary preprocessing (the role of the reference's tools/raw_script/criteo2raw.cpp
// and tools/dlrm_script/dlrm_raw.cu, which does the categorification with cuDF on the GPU).
// Line: label \t I1..I13 (decimal, may be empty / negative) \t C1..C26 (hex, may be empty)
// fit():       multi-threaded frequency count per categorical column over any number of files
// finalize():  per column, values with count >= min_freq get ids 1.. by (count desc, value asc),
//              optionally truncated to max_size - 1 ids; 0 = rare / unseen
// transform(): label i32 | dense u32 = max(0, x) | ids u32 (optionally % max_ind_range), fixed
//              160-byte records for the default 13 + 26 layout (the RawAsync reader applies log(x+1))
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

using Counts = std::unordered_map<uint64_t, uint64_t>;

struct Pre {
  int num_dense, num_cat;
  std::vector<Counts> counts;                                  // fit state
  std::vector<std::unordered_map<uint64_t, uint32_t>> vocab;   // value -> id (>= 1)
};

struct Map {
  const char* p{nullptr};
  size_t n{0};
  int fd{-1};
  bool open_(const char* path) {
    fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    fstat(fd, &st);
    n = static_cast<size_t>(st.st_size);
    if (n == 0) { p = nullptr; return true; }
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); fd = -1; return false; }
    p = static_cast<const char*>(m);
    return true;
  }
  ~Map() {
    if (p) munmap(const_cast<char*>(p), n);
    if (fd >= 0) close(fd);
  }
};

// [begin, end) byte ranges aligned to line starts
std::vector<std::pair<size_t, size_t>> split_lines(const Map& m, int parts) {
  std::vector<std::pair<size_t, size_t>> out;
  size_t start = 0;
  for (int t = 0; t < parts && start < m.n; ++t) {
    size_t end = (t == parts - 1) ? m.n : m.n / parts * (t + 1);
    if (end <= start) continue;
    if (end < m.n) {   // extend to just past the next newline
      const void* nl = memchr(m.p + end - 1, '\n', m.n - (end - 1));
      end = nl ? static_cast<size_t>(static_cast<const char*>(nl) - m.p) + 1 : m.n;
    }
    out.emplace_back(start, end);
    start = end;
  }
  return out;
}

inline int hexval(char c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}

// Parses one line starting at q (ends at '\n' or e).  Returns the position after the line.
// label/dense may be null (fit pass only needs the categoricals).
inline const char* parse_line(const char* q, const char* e, int nd, int nc, int32_t* label,
                              uint32_t* dense, uint64_t* cat) {
  int field = 0;
  const int nf = 1 + nd + nc;
  while (field < nf) {
    const char* s = q;
    while (q < e && *q != '\t' && *q != '\n') ++q;
    if (field <= nd) {
      long long v = 0;
      bool neg = false;
      const char* c = s;
      if (c < q && (*c == '-' || *c == '+')) { neg = *c == '-'; ++c; }
      for (; c < q && *c >= '0' && *c <= '9'; ++c) v = v * 10 + (*c - '0');
      if (neg) v = -v;
      if (field == 0) { if (label) *label = static_cast<int32_t>(v); }
      else if (dense) dense[field - 1] = v > 0 ? static_cast<uint32_t>(v) : 0u;
    } else {
      uint64_t v = 0;
      for (const char* c = s; c < q; ++c) {
        const int h = hexval(*c);
        if (h < 0) break;
        v = (v << 4) | static_cast<uint64_t>(h);
      }
      cat[field - 1 - nd] = v;
    }
    ++field;
    if (q < e && *q == '\t') { ++q; continue; }
    break;   // newline or end of buffer: remaining fields are empty
  }
  for (; field < nf; ++field) {
    if (field == 0) { if (label) *label = 0; }
    else if (field <= nd) { if (dense) dense[field - 1] = 0; }
    else cat[field - 1 - nd] = 0;
  }
  while (q < e && *q != '\n') ++q;
  return q < e ? q + 1 : q;
}

template <typename F>
void run_parts(int parts, F&& f) {
  std::vector<std::thread> th;
  for (int t = 1; t < parts; ++t) th.emplace_back([&f, t] { f(t); });
  if (parts > 0) f(0);
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" void* hctr_criteo_open(int num_dense, int num_cat) {
  Pre* p = new Pre();
  p->num_dense = num_dense; p->num_cat = num_cat;
  p->counts.resize(num_cat);
  return p;
}

extern "C" void hctr_criteo_close(void* h) { delete static_cast<Pre*>(h); }

// pass 1 over one file; returns the number of lines, -1 on IO error
extern "C" long long hctr_criteo_fit(void* h, const char* path, int num_threads) {
  Pre* P = static_cast<Pre*>(h);
  Map m;
  if (!m.open_(path)) return -1;
  const int nt = std::max(1, std::min(64, num_threads));
  auto parts = split_lines(m, nt);
  const int np = static_cast<int>(parts.size());
  std::vector<std::vector<Counts>> local(np, std::vector<Counts>(P->num_cat));
  std::vector<long long> lines(np, 0);
  run_parts(np, [&](int t) {
    std::vector<uint64_t> cat(P->num_cat);
    const char* q = m.p + parts[t].first;
    const char* e = m.p + parts[t].second;
    while (q < e) {
      q = parse_line(q, e, P->num_dense, P->num_cat, nullptr, nullptr, cat.data());
      for (int j = 0; j < P->num_cat; ++j) ++local[t][j][cat[j]];
      ++lines[t];
    }
  });
  // merge: one thread per column
  std::atomic<int> col{0};
  run_parts(std::min(nt, P->num_cat), [&](int) {
    for (;;) {
      const int j = col.fetch_add(1);
      if (j >= P->num_cat) return;
      for (int t = 0; t < np; ++t)
        for (auto& kv : local[t][j]) P->counts[j][kv.first] += kv.second;
    }
  });
  long long tot = 0;
  for (long long l : lines) tot += l;
  return tot;
}

extern "C" int hctr_criteo_finalize(void* h, long long min_freq, long long max_size) {
  Pre* P = static_cast<Pre*>(h);
  P->vocab.assign(P->num_cat, {});
  for (int j = 0; j < P->num_cat; ++j) {
    std::vector<std::pair<uint64_t, uint64_t>> kept;   // (value, count)
    for (auto& kv : P->counts[j])
      if (static_cast<long long>(kv.second) >= min_freq) kept.push_back(kv);
    std::sort(kept.begin(), kept.end(), [](const auto& a, const auto& b) {
      return a.second != b.second ? a.second > b.second : a.first < b.first;
    });
    if (max_size > 0 && static_cast<long long>(kept.size()) > max_size - 1) kept.resize(max_size - 1);
    auto& v = P->vocab[j];
    v.reserve(kept.size() * 2);
    for (size_t i = 0; i < kept.size(); ++i) v[kept[i].first] = static_cast<uint32_t>(i + 1);
  }
  return 0;
}

extern "C" long long hctr_criteo_vocab_size(void* h, int col) {
  Pre* P = static_cast<Pre*>(h);
  return static_cast<long long>(P->vocab[col].size()) + 1;
}

// dump of one column's vocabulary in id order (ids 1..size-1): values[i] has id i + 1
extern "C" void hctr_criteo_vocab_dump(void* h, int col, unsigned long long* values) {
  Pre* P = static_cast<Pre*>(h);
  for (auto& kv : P->vocab[col]) values[kv.second - 1] = kv.first;
}

// load a vocabulary (ids 1..n in the given order), e.g. one fitted earlier
extern "C" void hctr_criteo_vocab_load(void* h, int col, const unsigned long long* values, long long n) {
  Pre* P = static_cast<Pre*>(h);
  if (static_cast<int>(P->vocab.size()) != P->num_cat) P->vocab.assign(P->num_cat, {});
  auto& v = P->vocab[col];
  v.clear();
  v.reserve(static_cast<size_t>(n) * 2);
  for (long long i = 0; i < n; ++i) v[values[i]] = static_cast<uint32_t>(i + 1);
}

// pass 2: returns the number of records written (appended at record offset `first_record`), -1 on error
extern "C" long long hctr_criteo_transform(void* h, const char* tsv, const char* out_path,
                                           long long first_record, long long max_ind_range,
                                           int num_threads) {
  Pre* P = static_cast<Pre*>(h);
  if (static_cast<int>(P->vocab.size()) != P->num_cat) return -1;
  Map m;
  if (!m.open_(tsv)) return -1;
  int fd = open(out_path, O_WRONLY | O_CREAT, 0644);
  if (fd < 0) return -1;
  const int nt = std::max(1, std::min(64, num_threads));
  auto parts = split_lines(m, nt);
  const int np = static_cast<int>(parts.size());
  const int nd = P->num_dense, nc = P->num_cat;
  const size_t rec = 4ull * (1 + nd + nc);
  // line counts per part -> record offsets
  std::vector<long long> lines(np, 0);
  run_parts(np, [&](int t) {
    const char* q = m.p + parts[t].first;
    const char* e = m.p + parts[t].second;
    long long c = 0;
    while (q < e) {
      const void* nl = memchr(q, '\n', static_cast<size_t>(e - q));
      ++c;
      if (!nl) break;
      q = static_cast<const char*>(nl) + 1;
    }
    lines[t] = c;
  });
  std::vector<long long> first(np + 1, first_record);
  for (int t = 0; t < np; ++t) first[t + 1] = first[t] + lines[t];
  std::atomic<bool> ok{true};
  run_parts(np, [&](int t) {
    std::vector<uint64_t> cat(nc);
    std::vector<char> buf;
    buf.reserve(1 << 20);
    const char* q = m.p + parts[t].first;
    const char* e = m.p + parts[t].second;
    long long off = first[t] * static_cast<long long>(rec);
    std::vector<uint32_t> r(1 + nd + nc);
    auto flush = [&] {
      const char* p = buf.data();
      size_t n = buf.size();
      while (n) {
        ssize_t w = pwrite(fd, p, n, off);
        if (w <= 0) { ok.store(false); return; }
        p += w; n -= static_cast<size_t>(w); off += w;
      }
      buf.clear();
    };
    while (q < e && ok.load()) {
      int32_t label;
      q = parse_line(q, e, nd, nc, &label, r.data() + 1, cat.data());
      memcpy(r.data(), &label, 4);
      for (int j = 0; j < nc; ++j) {
        auto it = P->vocab[j].find(cat[j]);
        uint32_t id = it == P->vocab[j].end() ? 0u : it->second;
        if (max_ind_range > 0) id = static_cast<uint32_t>(id % static_cast<uint64_t>(max_ind_range));
        r[1 + nd + j] = id;
      }
      const char* c = reinterpret_cast<const char*>(r.data());
      buf.insert(buf.end(), c, c + rec);
      if (buf.size() >= (1u << 20)) flush();
    }
    if (!buf.empty() && ok.load()) flush();
  });
  close(fd);
  return ok.load() ? first[np] - first_record : -1;
}
