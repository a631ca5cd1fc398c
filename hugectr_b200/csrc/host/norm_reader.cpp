// Native Norm-format reader (host side of C12/C46 for the legacy "Norm" data format).
// A producer thread scans the (variable-length) records of the file list in order, verifies the
// per-record checksum when the files carry one, and decodes THIS rank's slice of every global batch
// straight into pinned staging slots in the layout the model consumes:
//   label [b, L] f32 | dense [b, D] f32 | keys: per sparse param a block [b, S, H] (-1 padded, slot
//   offsets added) | nnz: per sparse param a block [S, b] i32
// File: DataSetHeader = 8 x int64 {error_check, number_of_records, label_dim, dense_dim, slot_num,
// reserved[3]} (reference: HugeCTR/include/common.hpp:184-191), then records
// [label f32 x L][dense f32 x D]{int nnz, keys[nnz]} x slot_num; with error_check == 1 every record is
// wrapped as {int nbytes, payload[nbytes], char sum} (HugeCTR/include/data_readers/check_sum.hpp,
// writer side HugeCTR/include/data_generator.hpp:137-188).
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct NSlot {
  float* label;
  float* dense;
  void* keys;
  int32_t* nnz;
  int state{0};          // 0 free, 2 ready
  long long seq{-1};
  int valid_global{0};   // records of the whole global batch; -1 end of data; -2 error
};

struct Mapped {
  const char* p{nullptr};
  size_t n{0};
  int fd{-1};
  void close_() {
    if (p) munmap(const_cast<char*>(p), n);
    if (fd >= 0) close(fd);
    p = nullptr; fd = -1; n = 0;
  }
};

struct NormReader {
  std::vector<std::string> files;
  int label_dim, dense_dim, total_slots, key_in, key_out, check;
  int batch_global, batch_local, rank;
  bool repeat;
  std::vector<int> blkS, blkH;
  std::vector<long long> slot_off;   // empty = none
  std::vector<NSlot> slots;
  std::thread producer;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<bool> stop{false};
  long long consume_seq{0};
  std::string error;

  // scan state (producer thread only)
  size_t file_i{0};
  Mapped cur;
  size_t pos{0};
  long long left{0};
  int file_check{0};
  long long records_seen{0};

  bool fail(const std::string& m) { error = m; return false; }

  bool open_next() {   // false: no more data (or error -> error non-empty)
    for (;;) {
      cur.close_();
      if (file_i >= files.size()) {
        if (!repeat || records_seen == 0) return false;
        file_i = 0;
      }
      const std::string& fp = files[file_i++];
      cur.fd = open(fp.c_str(), O_RDONLY);
      if (cur.fd < 0) return fail(fp + ": cannot open");
      struct stat st;
      fstat(cur.fd, &st);
      cur.n = static_cast<size_t>(st.st_size);
      if (cur.n < 64) return fail(fp + ": shorter than the data set header");
      void* m = mmap(nullptr, cur.n, PROT_READ, MAP_PRIVATE, cur.fd, 0);
      if (m == MAP_FAILED) { cur.p = nullptr; return fail(fp + ": mmap failed"); }
      cur.p = static_cast<const char*>(m);
      madvise(m, cur.n, MADV_SEQUENTIAL);
      long long h[8];
      memcpy(h, cur.p, 64);
      if (h[2] != label_dim || h[3] != dense_dim || h[4] != total_slots)
        return fail(fp + ": header (label " + std::to_string(h[2]) + ", dense " + std::to_string(h[3]) +
                    ", slots " + std::to_string(h[4]) + ") does not match the model");
      if (h[0] == 1 && !check) return fail(fp + " was written with check_sum, use Check_t.Sum");
      file_check = static_cast<int>(h[0]);
      left = h[1];
      pos = 64;
      if (left > 0) return true;
    }
  }

  // next record: payload pointer + length bound. false at end of data / on error.
  bool next_record(const char** rec, size_t* len, bool verify) {
    while (left == 0) {
      if (!open_next()) return false;
    }
    const std::string& fp = files[file_i - 1];
    if (file_check == 1) {
      if (pos + 4 > cur.n) return fail(fp + ": truncated record");
      int32_t nb;
      memcpy(&nb, cur.p + pos, 4);
      if (nb < 0 || pos + 5 + static_cast<size_t>(nb) > cur.n) return fail(fp + ": truncated record");
      const char* pl = cur.p + pos + 4;
      if (verify) {   // every record is verified by the rank that consumes it
        int8_t s = 0;
        for (int i = 0; i < nb; ++i) s = static_cast<int8_t>(s + static_cast<int8_t>(pl[i]));
        if (s != static_cast<int8_t>(pl[nb])) return fail(fp + ": checksum mismatch");
      }
      *rec = pl; *len = static_cast<size_t>(nb);
      pos += 5 + static_cast<size_t>(nb);
    } else {
      size_t q = pos + 4ull * (label_dim + dense_dim);
      for (int s = 0; s < total_slots; ++s) {
        if (q + 4 > cur.n) return fail(fp + ": truncated record");
        int32_t c;
        memcpy(&c, cur.p + q, 4);
        if (c < 0) return fail(fp + ": negative nnz");
        q += 4 + static_cast<size_t>(c) * key_in;
      }
      if (q > cur.n) return fail(fp + ": truncated record");
      *rec = cur.p + pos; *len = q - pos;
      pos = q;
    }
    --left;
    ++records_seen;
    return true;
  }

  inline void put_key(void* base, size_t o, long long k) const {
    if (key_out == 4) static_cast<int32_t*>(base)[o] = static_cast<int32_t>(k);
    else static_cast<int64_t*>(base)[o] = k;
  }

  void clear_slot(NSlot& s) {
    const size_t b = batch_local;
    memset(s.label, 0, b * label_dim * 4);
    memset(s.dense, 0, b * dense_dim * 4);
    size_t ko = 0, no = 0;
    for (size_t k = 0; k < blkS.size(); ++k) {
      const size_t n = b * blkS[k] * blkH[k];
      if (key_out == 4) for (size_t i = 0; i < n; ++i) static_cast<int32_t*>(s.keys)[ko + i] = -1;
      else for (size_t i = 0; i < n; ++i) static_cast<int64_t*>(s.keys)[ko + i] = -1;
      ko += n;
      memset(s.nnz + no, 0, b * blkS[k] * 4);
      no += b * blkS[k];
    }
  }

  bool decode(NSlot& s, int i, const char* rec, size_t len) {
    const size_t b = batch_local;
    size_t q = 0;
    if (len < 4ull * (label_dim + dense_dim)) return fail("record shorter than label + dense");
    memcpy(s.label + static_cast<size_t>(i) * label_dim, rec, 4ull * label_dim);
    q += 4ull * label_dim;
    memcpy(s.dense + static_cast<size_t>(i) * dense_dim, rec + q, 4ull * dense_dim);
    q += 4ull * dense_dim;
    size_t ko = 0, no = 0;
    int si = 0;
    for (size_t k = 0; k < blkS.size(); ++k) {
      const int S = blkS[k], H = blkH[k];
      for (int sl = 0; sl < S; ++sl, ++si) {
        if (q + 4 > len) return fail("record shorter than its slot list");
        int32_t c;
        memcpy(&c, rec + q, 4);
        q += 4;
        if (c < 0 || q + static_cast<size_t>(c) * key_in > len) return fail("record shorter than its keys");
        const long long off = slot_off.empty() ? 0 : slot_off[si];
        const int take = c < H ? c : H;
        const size_t o = ko + (static_cast<size_t>(i) * S + sl) * H;
        for (int h = 0; h < take; ++h) {
          long long key;
          if (key_in == 4) { uint32_t t; memcpy(&t, rec + q + 4ull * h, 4); key = t; }
          else memcpy(&key, rec + q + 8ull * h, 8);
          put_key(s.keys, o + h, key + off);
        }
        s.nnz[no + static_cast<size_t>(sl) * b + i] = take;
        q += static_cast<size_t>(c) * key_in;
      }
      ko += b * S * H;
      no += b * S;
    }
    return true;
  }

  void publish(NSlot& s, long long seq, int valid) {
    {
      std::lock_guard<std::mutex> lk(mu);
      s.seq = seq; s.valid_global = valid; s.state = 2;
    }
    cv.notify_all();
  }

  void run() {
    for (long long seq = 0; !stop.load(); ++seq) {
      NSlot& s = slots[seq % slots.size()];
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop.load() || (s.state == 0 && consume_seq + (long long)slots.size() > seq); });
        if (stop.load()) return;
      }
      clear_slot(s);
      int got = 0;
      const int lo = rank * batch_local, hi = lo + batch_local;
      for (int g = 0; g < batch_global; ++g) {
        const char* rec; size_t len;
        const bool mine = g >= lo && g < hi;
        if (!next_record(&rec, &len, mine)) break;
        if (mine && !decode(s, g - lo, rec, len)) break;
        ++got;
      }
      if (!error.empty()) { publish(s, seq, -2); return; }
      if (got == 0) { publish(s, seq, -1); return; }
      publish(s, seq, got);
    }
  }
};

}  // namespace

extern "C" void* hctr_norm_open(const char** files, int nfiles, int label_dim, int dense_dim,
                                const int* blk_s, const int* blk_h, int nblocks,
                                const long long* slot_off, int key_in, int key_out, int check,
                                int batch_global, int batch_local, int rank, int repeat) {
  NormReader* r = new NormReader();
  for (int i = 0; i < nfiles; ++i) r->files.emplace_back(files[i]);
  r->label_dim = label_dim; r->dense_dim = dense_dim;
  r->blkS.assign(blk_s, blk_s + nblocks);
  r->blkH.assign(blk_h, blk_h + nblocks);
  r->total_slots = 0;
  for (int s : r->blkS) r->total_slots += s;
  if (slot_off) r->slot_off.assign(slot_off, slot_off + r->total_slots);
  r->key_in = key_in; r->key_out = key_out; r->check = check;
  r->batch_global = batch_global; r->batch_local = batch_local; r->rank = rank; r->repeat = repeat != 0;
  return r;
}

extern "C" int hctr_norm_start(void* h, int depth, float** labels, float** denses, void** keys,
                               int32_t** nnz) {
  NormReader* r = static_cast<NormReader*>(h);
  r->slots = std::vector<NSlot>(depth);
  for (int i = 0; i < depth; ++i) {
    r->slots[i].label = labels[i]; r->slots[i].dense = denses[i];
    r->slots[i].keys = keys[i]; r->slots[i].nnz = nnz[i];
  }
  r->producer = std::thread([r] { r->run(); });
  return 0;
}

// Blocks until the next global batch is decoded.  Returns the slot index; *valid_global = number of
// records of the global batch (-1: end of data, -2: error, see hctr_norm_error).  The slot handed
// out by the previous call is recycled.
extern "C" int hctr_norm_next(void* h, int* valid_global) {
  NormReader* r = static_cast<NormReader*>(h);
  std::unique_lock<std::mutex> lk(r->mu);
  if (r->consume_seq > 0) {
    r->slots[(r->consume_seq - 1) % r->slots.size()].state = 0;
    r->cv.notify_all();
  }
  const long long seq = r->consume_seq;
  NSlot& s = r->slots[seq % r->slots.size()];
  r->cv.wait(lk, [&] { return s.state == 2 && s.seq == seq; });
  *valid_global = s.valid_global;
  if (s.valid_global >= 0) r->consume_seq = seq + 1;   // end / error markers stay readable
  return static_cast<int>(seq % r->slots.size());
}

extern "C" const char* hctr_norm_error(void* h) { return static_cast<NormReader*>(h)->error.c_str(); }

extern "C" void hctr_norm_close(void* h) {
  NormReader* r = static_cast<NormReader*>(h);
  r->stop.store(true);
  { std::lock_guard<std::mutex> lk(r->mu); }
  r->cv.notify_all();
  if (r->producer.joinable()) r->producer.join();
  r->cur.close_();
  delete r;
}
