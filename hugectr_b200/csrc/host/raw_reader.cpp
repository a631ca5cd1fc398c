// Native asynchronous multi-hot raw reader (host side of C13).
// Worker threads pread() fixed-size records of THIS rank's slice of every global batch, split them
// into label / dense (optionally log1p of integer features) / feature-major keys directly inside
// pinned staging slots owned by the Python side, and keep `depth` batches in flight.
// Record: [label_dim x 4B][dense_dim x 4B][sum(hotness) x key_bytes]
// (reference: HugeCTR/src/data_readers/multi_hot/detail/{batch_file_reader,data_reader_impl}.cpp,
//  split kernel HugeCTR/src/data_readers/multi_hot/split_batch.cu:43-88 -- done on the CPU workers
//  here so the H2D copy lands in its final layout and needs no extra device kernel).
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Slot {
  float* label;
  float* dense;
  void* keys;
  std::atomic<int> state{0};  // 0 free, 1 filling, 2 ready
  long long seq{-1};
  int valid{0};
};

struct RawReader {
  int fd{-1};
  long long file_bytes{0}, num_samples{0};
  int label_dim, dense_dim, key_bytes_in, key_bytes_out, dense_is_float;
  std::vector<int> hot;
  long long rec_bytes;
  int batch_global, batch_local, rank;
  bool repeat;
  std::vector<Slot> slots;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<bool> stop{false};
  std::atomic<long long> next_seq{0};   // next batch to be produced
  long long consume_seq{0};
  long long batches_per_epoch;

  void fill(Slot& s, long long seq) {
    const long long it = repeat ? (seq % batches_per_epoch) : seq;
    const long long first = it * batch_global + static_cast<long long>(rank) * batch_local;
    int n = batch_local;
    if (first + n > num_samples) n = static_cast<int>(std::max(0ll, num_samples - first));
    s.valid = n;
    std::vector<char> buf(static_cast<size_t>(std::max(n, 1)) * rec_bytes);
    if (n > 0) {
      size_t off = 0, total = static_cast<size_t>(n) * rec_bytes;
      while (off < total) {
        ssize_t r = pread(fd, buf.data() + off, total - off, first * rec_bytes + off);
        if (r <= 0) break;
        off += r;
      }
    }
    int tot_hot = 0;
    for (int h : hot) tot_hot += h;
    for (int i = 0; i < batch_local; ++i) {
      const char* rec = buf.data() + static_cast<size_t>(i < n ? i : 0) * rec_bytes;
      const bool ok = i < n;
      const uint32_t* ld = reinterpret_cast<const uint32_t*>(rec);
      for (int j = 0; j < label_dim; ++j) {
        float v = 0.f;
        if (ok) {
          if (dense_is_float & 2) memcpy(&v, ld + j, 4);      // bit 1: labels stored as floats
          else v = static_cast<float>(static_cast<int32_t>(ld[j]));
        }
        s.label[static_cast<size_t>(i) * label_dim + j] = v;
      }
      for (int j = 0; j < dense_dim; ++j) {
        float v = 0.f;
        if (ok) {
          if (dense_is_float & 1) memcpy(&v, ld + label_dim + j, 4);   // bit 0: dense stored as floats
          else v = logf(static_cast<float>(ld[label_dim + j]) + 1.f);   // log(x+1) of uint features
        }
        s.dense[static_cast<size_t>(i) * dense_dim + j] = v;
      }
    }
    // feature-major keys: block f = [batch_local, hot[f]]
    const size_t key_base = static_cast<size_t>(label_dim + dense_dim) * 4;
    size_t out_off = 0;   // elements
    int in_off = 0;       // keys inside the record
    for (size_t f = 0; f < hot.size(); ++f) {
      const int H = hot[f];
      for (int i = 0; i < batch_local; ++i) {
        const bool ok = i < n;
        const char* kp = buf.data() + static_cast<size_t>(ok ? i : 0) * rec_bytes + key_base +
                         static_cast<size_t>(in_off) * key_bytes_in;
        for (int h = 0; h < H; ++h) {
          long long k = -1;
          if (ok) {
            if (key_bytes_in == 4) { uint32_t t; memcpy(&t, kp + 4 * h, 4); k = t; }
            else { memcpy(&k, kp + 8 * h, 8); }
          }
          const size_t o = out_off + static_cast<size_t>(i) * H + h;
          if (key_bytes_out == 4) reinterpret_cast<int32_t*>(s.keys)[o] = static_cast<int32_t>(k);
          else reinterpret_cast<int64_t*>(s.keys)[o] = k;
        }
      }
      out_off += static_cast<size_t>(batch_local) * H;
      in_off += H;
    }
  }

  void worker() {
    while (!stop.load()) {
      long long seq = next_seq.fetch_add(1);
      if (!repeat && seq >= batches_per_epoch) {
        // publish an end-of-data marker in order
        Slot& s = slots[seq % slots.size()];
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop.load() || (s.state.load() == 0 && consume_seq + (long long)slots.size() > seq); });
        if (stop.load()) return;
        s.seq = seq; s.valid = -1; s.state.store(2);
        cv.notify_all();
        return;
      }
      Slot& s = slots[seq % slots.size()];
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop.load() || (s.state.load() == 0 && consume_seq + (long long)slots.size() > seq); });
        if (stop.load()) return;
        s.state.store(1);
      }
      fill(s, seq);
      {
        std::lock_guard<std::mutex> lk(mu);
        s.seq = seq;
        s.state.store(2);
      }
      cv.notify_all();
    }
  }
};

}  // namespace

extern "C" void* hctr_raw_open(const char* path, int label_dim, int dense_dim, const int* hot,
                               int num_feat, int key_bytes_in, int key_bytes_out, int dense_is_float,
                               int batch_global, int batch_local, int rank, int repeat,
                               long long num_samples_hint) {
  RawReader* r = new RawReader();
  r->fd = open(path, O_RDONLY);
  if (r->fd < 0) { delete r; return nullptr; }
  struct stat st;
  fstat(r->fd, &st);
  r->file_bytes = st.st_size;
  r->label_dim = label_dim; r->dense_dim = dense_dim;
  r->hot.assign(hot, hot + num_feat);
  r->key_bytes_in = key_bytes_in; r->key_bytes_out = key_bytes_out; r->dense_is_float = dense_is_float;
  long long th = 0; for (int h : r->hot) th += h;
  r->rec_bytes = (label_dim + dense_dim) * 4ll + th * key_bytes_in;
  r->num_samples = r->file_bytes / r->rec_bytes;
  if (num_samples_hint > 0 && num_samples_hint < r->num_samples) r->num_samples = num_samples_hint;
  r->batch_global = batch_global; r->batch_local = batch_local; r->rank = rank; r->repeat = repeat != 0;
  r->batches_per_epoch = (r->num_samples + batch_global - 1) / batch_global;
  if (r->batches_per_epoch < 1) r->batches_per_epoch = 1;
  return r;
}

extern "C" long long hctr_raw_num_samples(void* h) { return static_cast<RawReader*>(h)->num_samples; }
extern "C" long long hctr_raw_batches_per_epoch(void* h) { return static_cast<RawReader*>(h)->batches_per_epoch; }

extern "C" int hctr_raw_start(void* h, int num_threads, int depth, float** labels, float** denses,
                              void** keys) {
  RawReader* r = static_cast<RawReader*>(h);
  r->slots = std::vector<Slot>(depth);
  for (int i = 0; i < depth; ++i) {
    r->slots[i].label = labels[i]; r->slots[i].dense = denses[i]; r->slots[i].keys = keys[i];
  }
  r->stop.store(false);
  r->next_seq.store(0);
  r->consume_seq = 0;
  for (int t = 0; t < num_threads; ++t) r->workers.emplace_back([r] { r->worker(); });
  return 0;
}

// blocks until the next batch (in order) is ready; returns the slot index, *valid = samples (or -1 at
// end of data). The previously returned slot is released by this call.
extern "C" int hctr_raw_next(void* h, int* valid) {
  RawReader* r = static_cast<RawReader*>(h);
  std::unique_lock<std::mutex> lk(r->mu);
  if (r->consume_seq > 0) {
    Slot& prev = r->slots[(r->consume_seq - 1) % r->slots.size()];
    prev.state.store(0);
    r->cv.notify_all();
  }
  const long long seq = r->consume_seq;
  Slot& s = r->slots[seq % r->slots.size()];
  r->cv.wait(lk, [&] { return s.state.load() == 2 && s.seq == seq; });
  *valid = s.valid;
  r->consume_seq = seq + 1;
  return static_cast<int>(seq % r->slots.size());
}

extern "C" void hctr_raw_close(void* h) {
  RawReader* r = static_cast<RawReader*>(h);
  r->stop.store(true);
  r->cv.notify_all();
  for (auto& t : r->workers) if (t.joinable()) t.join();
  if (r->fd >= 0) close(r->fd);
  delete r;
}

// ------------------------------------------------------------------------------------------------
// Device-split mode: the workers only move BYTES -- this rank's slice of a batch is read with O_DIRECT
// (page-cache bypass, aligned offset / length / buffer like the reference's libaio reader,
// HugeCTR/src/data_readers/multi_hot/detail/{aio_context,batch_file_reader}.cpp) straight into a pinned,
// page-aligned staging slot; the consumer copies the slot to the device in one H2D transfer and a device
// kernel (csrc/reader_split.cu) splits it into label / dense / feature-major keys.  `depth` reads are in
// flight (one positional read per worker thread at a time).
namespace {

struct RawSlotD {
  char* buf;
  std::atomic<int> state{0};
  long long seq{-1};
  int valid{0};
  int skew{0};          // byte offset of the first record inside buf
};

struct RawReaderD {
  int fd{-1};
  bool direct{false};
  long long num_samples{0}, rec_bytes{0}, slot_bytes{0};
  int batch_global, batch_local, rank;
  bool repeat;
  long long batches_per_epoch;
  std::vector<RawSlotD> slots;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<bool> stop{false};
  std::atomic<long long> next_seq{0};
  long long consume_seq{0};

  void fill(RawSlotD& s, long long seq) {
    const long long it = repeat ? (seq % batches_per_epoch) : seq;
    const long long first = it * batch_global + static_cast<long long>(rank) * batch_local;
    int n = batch_local;
    if (first + n > num_samples) n = static_cast<int>(std::max(0ll, num_samples - first));
    s.valid = n;
    s.skew = 0;
    if (n <= 0) return;
    const long long beg = first * rec_bytes, end = beg + static_cast<long long>(n) * rec_bytes;
    const long long A = 4096;
    const long long abeg = direct ? (beg / A) * A : beg;
    const long long aend = direct ? ((end + A - 1) / A) * A : end;
    s.skew = static_cast<int>(beg - abeg);
    long long off = 0;
    const long long total = aend - abeg;
    while (off < total) {
      ssize_t r = pread(fd, s.buf + off, static_cast<size_t>(total - off), abeg + off);
      if (r <= 0) break;          // (the tail of the file may be shorter than the aligned length)
      off += r;
    }
  }

  void worker() {
    while (!stop.load()) {
      const long long seq = next_seq.fetch_add(1);
      RawSlotD& s = slots[seq % slots.size()];
      const bool eof = !repeat && seq >= batches_per_epoch;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop.load() || (s.state.load() == 0 && consume_seq + (long long)slots.size() > seq); });
        if (stop.load()) return;
        s.state.store(1);
      }
      if (eof) s.valid = -1; else fill(s, seq);
      {
        std::lock_guard<std::mutex> lk(mu);
        s.seq = seq;
        s.state.store(2);
      }
      cv.notify_all();
      if (eof) return;
    }
  }
};

}  // namespace

extern "C" void* hctr_rawd_open(const char* path, long long rec_bytes, int batch_global, int batch_local, int rank,
                                int repeat, long long num_samples_hint, int use_direct) {
  RawReaderD* r = new RawReaderD();
  r->fd = -1;
  if (use_direct) {
    r->fd = open(path, O_RDONLY | O_DIRECT);
    r->direct = r->fd >= 0;
  }
  if (r->fd < 0) r->fd = open(path, O_RDONLY);
  if (r->fd < 0) { delete r; return nullptr; }
  struct stat st;
  fstat(r->fd, &st);
  r->rec_bytes = rec_bytes;
  r->num_samples = st.st_size / rec_bytes;
  if (num_samples_hint > 0 && num_samples_hint < r->num_samples) r->num_samples = num_samples_hint;
  r->batch_global = batch_global; r->batch_local = batch_local; r->rank = rank; r->repeat = repeat != 0;
  r->batches_per_epoch = std::max(1ll, (r->num_samples + batch_global - 1) / batch_global);
  r->slot_bytes = ((static_cast<long long>(batch_local) * rec_bytes + 4095) / 4096 + 2) * 4096;
  return r;
}
extern "C" long long hctr_rawd_slot_bytes(void* h) { return static_cast<RawReaderD*>(h)->slot_bytes; }
extern "C" long long hctr_rawd_num_samples(void* h) { return static_cast<RawReaderD*>(h)->num_samples; }
extern "C" int hctr_rawd_is_direct(void* h) { return static_cast<RawReaderD*>(h)->direct ? 1 : 0; }
extern "C" int hctr_rawd_start(void* h, int num_threads, int depth, void** bufs) {
  RawReaderD* r = static_cast<RawReaderD*>(h);
  r->slots = std::vector<RawSlotD>(depth);
  for (int i = 0; i < depth; ++i) r->slots[i].buf = static_cast<char*>(bufs[i]);
  r->stop.store(false);
  r->next_seq.store(0);
  r->consume_seq = 0;
  for (int t = 0; t < num_threads; ++t) r->workers.emplace_back([r] { r->worker(); });
  return 0;
}
// returns the slot index; *valid = samples in the slot (-1: end of data), *skew = byte offset of record 0
extern "C" int hctr_rawd_next(void* h, int* valid, int* skew) {
  RawReaderD* r = static_cast<RawReaderD*>(h);
  std::unique_lock<std::mutex> lk(r->mu);
  if (r->consume_seq > 0) {
    r->slots[(r->consume_seq - 1) % r->slots.size()].state.store(0);
    r->cv.notify_all();
  }
  const long long seq = r->consume_seq;
  RawSlotD& s = r->slots[seq % r->slots.size()];
  r->cv.wait(lk, [&] { return s.state.load() == 2 && s.seq == seq; });
  *valid = s.valid;
  *skew = s.skew;
  r->consume_seq = seq + 1;
  return static_cast<int>(seq % r->slots.size());
}
extern "C" void hctr_rawd_close(void* h) {
  RawReaderD* r = static_cast<RawReaderD*>(h);
  r->stop.store(true);
  r->cv.notify_all();
  for (auto& t : r->workers) if (t.joinable()) t.join();
  if (r->fd >= 0) close(r->fd);
  delete r;
}
