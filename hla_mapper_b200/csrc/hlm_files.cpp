// hlm_files.cpp — file-level front and back end of the hot path:
//
//   FASTQ / FASTQ.gz ingest          replaces the getline loops of src/preselect.cpp:652-749,
//                                    :757-851 (.gz), :886-1043 (single end)
//   selected*.fastq writers          src/preselect.cpp:861-877, :1051-1203
//   addressing_table.txt writer      src/map_dna.cpp:1036-1051 (rows: :932-1016)
//   per-gene FASTQ emit              src/map_dna.cpp:1056-1351
//
// One reader thread per input file (zlib gzread handles plain and .gz alike) cuts the text into
// batches of pairs laid out as the SoA hlm_batch wants; the calling thread drives hlm_run() on
// batch k while the readers parse batch k+1.  Everything that the reference keeps per *selected*
// read (id, original record, trim window, scores, targets) is kept in memory, as the reference
// does; the unselected bulk of the input is never stored.
//
// Byte formats are the reference's (its std::map iteration orders included); only the orders
// that the reference itself leaves to thread timing (selected.R*.fastq, table rows) are fixed
// here to input order / id order.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "hlm_bgzf.h"
#include "hlm_pgzip.h"
#include "hlm_db.h"
#include "hlm_files.h"

extern "C" const char* hlm_last_error(const hlm_ctx* ctx);
void hlm_ctx_set_error(hlm_ctx* ctx, const std::string& msg);
const hlm_db* hlm_ctx_db(const hlm_ctx* ctx);
const hlm_params* hlm_ctx_params(const hlm_ctx* ctx);

namespace {

// f(t) for t in [0, n) on n host threads; an exception in a worker becomes one in the caller
template <class F>
void parallel_for(int n, F f) {
  if (n <= 1) {
    f(0);
    return;
  }
  std::vector<std::thread> th;
  std::atomic<int> failed(0);
  for (int t = 0; t < n; t++)
    th.emplace_back([&, t] {
      try {
        f(t);
      } catch (const std::exception&) {
        failed = 1;
      }
    });
  for (auto& x : th) x.join();
  if (failed) throw std::bad_alloc();
}

// Batch sizes of a run.  The GPU cannot start before its first batch is parsed, and the last
// batch is collected and written after the GPU has gone idle, so the run starts (and, where the
// length of the input is known, ends) with small batches: B/8, B/8, B/4, B/2, then B pairs.
// Runs with batches below 512 Ki pairs are cut uniformly.
int64_t ramp_min() {  // HLM_RAMP_MIN_PAIRS: lets the tests run the ramp on their small inputs
  const char* e = getenv("HLM_RAMP_MIN_PAIRS");
  long long v = e ? atoll(e) : 0;
  return v >= 8 ? (int64_t)v : (int64_t)8 << 16;
}
int64_t batch_size(int64_t k, int64_t B) {
  if (B < ramp_min()) return B;
  return k < 2 ? B / 8 : k == 2 ? B / 4 : k == 3 ? B / 2 : B;
}
// record index at which every batch of an input of N records starts (+ N at the end)
std::vector<int64_t> batch_plan(int64_t N, int64_t B) {
  std::vector<int64_t> rec(1, 0);
  for (int64_t k = 0; rec.back() < N; k++) rec.push_back(std::min(N, rec.back() + batch_size(k, B)));
  if (B >= ramp_min() && rec.size() >= 2) {
    // the last batch gives way to  rest | B/4 | B/8
    int64_t a = rec[rec.size() - 2], R = N - a, t1 = B / 4, t2 = B / 8;
    if (R > t1 + t2) {
      rec.pop_back();
      rec.push_back(N - t1 - t2);
      rec.push_back(N - t2);
      rec.push_back(N);
    }
  }
  return rec;
}

// reads 4-line records straight out of a large refill buffer (one memchr per line, one copy per
// field into the SoA vectors)
class FastqReader {
 public:
  explicit FastqReader(const char* path) {
    // plain text goes through read(2) directly; bgzip-compressed files (BGZF: independent
    // deflate blocks) and plain gzip files (one deflate stream: hlm_pgzip.h) are inflated on
    // several host threads, ahead of this reader's parsing
    fd_ = open(path, O_RDONLY);
    if (fd_ >= 0) {
      unsigned char magic[64];
      memset(magic, 0, sizeof magic);
      ssize_t r = pread(fd_, magic, sizeof magic, 0);
      if (r >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        close(fd_);
        fd_ = -1;
        int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / 2));
        if (hlm_is_bgzf(magic, (size_t)r) && map_.open(path)) {
          bg_.reset(new HlmBgzf(map_.p, map_.n, nt));
          bga_.reset(new HlmBgzfAhead(bg_.get()));  // inflated ahead of the parser
        } else if (!getenv("HLM_GZIP_ZLIB") && map_.open(path)) {
          // plain gzip: one deflate stream, inflated by several threads (hlm_pgzip.h)
          pg_.reset(new HlmPgzip(map_.p, map_.n, nt));
        } else {  // HLM_GZIP_ZLIB=1 (or a file that cannot be mapped): zlib's single-threaded reader
          f_ = gzopen(path, "rb");
          if (f_) gzbuffer(f_, 1 << 22);
        }
      } else {
        posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
      }
    }
    buf_.resize((size_t)32 << 20);
  }
  ~FastqReader() {
    if (f_) gzclose(f_);
    if (fd_ >= 0) close(fd_);
  }
  bool ok() const { return f_ != nullptr || fd_ >= 0 || bg_ != nullptr || pg_ != nullptr; }
  // appends up to max_pairs records to b; returns number appended
  int64_t read(MateBatch& b, int64_t max_pairs) {
    int64_t got = 0;
    while (got < max_pairs) {
      // find the four line ends of the next record inside the buffer; refill when incomplete
      const char* base = buf_.data();
      size_t e[4], p = pos_;
      int k = 0;
      for (; k < 4; k++) {
        const char* nl = (const char*)memchr(base + p, '\n', have_ - p);
        if (!nl) break;
        e[k] = (size_t)(nl - base);
        p = e[k] + 1;
      }
      if (k < 4) {
        if (!eof_) {
          refill();
          continue;
        }
        // end of file: a last line without newline completes the record, anything else is dropped
        if (k == 3 && p < have_) {
          e[3] = have_;
          p = have_;
        } else {
          if (pos_ < have_ && (k > 0 || have_ - pos_ > 0)) b.bad = b.bad || k > 0;
          break;
        }
      }
      size_t h0 = pos_, s0 = e[0] + 1, q0 = e[2] + 1;
      size_t hl = e[0] - h0, sl = e[1] - s0, ql = e[3] - q0;
      if (got == 0 && b.n == 0) {
        // room for the whole batch, sized from its first record: vectors that grow by doubling
        // were copied (and their fresh pages faulted in) about twice over - most of the parsing
        // time.  Untouched reserve costs nothing; longer reads later on still grow the vectors.
        size_t m = (size_t)max_pairs;
        b.hdr.reserve(m * (hl + 4));
        b.bases.reserve(m * (sl + 1));
        b.quals.reserve(m * (sl + 1));
        b.offs.reserve(m + 1);
        b.hoff.reserve(m + 1);
      }
      b.hdr.insert(b.hdr.end(), base + h0, base + h0 + hl);
      b.hoff.push_back(b.hdr.size());
      b.bases.insert(b.bases.end(), (const uint8_t*)base + s0, (const uint8_t*)base + s0 + sl);
      // the quality line is cut / padded to the sequence length only for the device arrays;
      // hla-mapper itself never checks that they agree
      size_t q = std::min(sl, ql);
      b.quals.insert(b.quals.end(), (const uint8_t*)base + q0, (const uint8_t*)base + q0 + q);
      if (q < sl) b.quals.insert(b.quals.end(), sl - q, (uint8_t)'!');
      b.offs.push_back(b.bases.size());
      pos_ = p;
      got++;
    }
    b.n += got;
    if (corrupt_) b.io_error = true;
    return got;
  }

 private:
  void refill() {  // move the unread tail to the front and read on
    size_t tail = have_ - pos_;
    memmove(buf_.data(), buf_.data() + pos_, tail);
    pos_ = 0;
    have_ = tail;
    if (buf_.size() - have_ < (1u << 16)) buf_.resize(buf_.size() * 2);  // room for one BGZF block
    size_t want = std::min<size_t>(buf_.size() - have_, 1u << 30);
    long r;
    if (bga_) {
      r = bga_->fill((uint8_t*)buf_.data() + have_, want);
      if (r < 0) corrupt_ = true;
    } else if (pg_) {
      r = pg_->fill((uint8_t*)buf_.data() + have_, want);
      if (r < 0) corrupt_ = true;
    } else if (f_) {
      r = (long)gzread(f_, buf_.data() + have_, (unsigned)want);
      int en = Z_OK;
      if (r <= 0) gzerror(f_, &en);
      if (r < 0 || (en != Z_OK && en != Z_STREAM_END)) corrupt_ = true;  // truncated / damaged gzip stream
    } else {
      r = (long)::read(fd_, buf_.data() + have_, want);
    }
    if (r <= 0) eof_ = true;
    else have_ += (size_t)r;
  }
  gzFile f_ = nullptr;
  int fd_ = -1;
  HlmMapped map_;
  std::unique_ptr<HlmBgzf> bg_;
  std::unique_ptr<HlmBgzfAhead> bga_;  // declared after bg_: destroyed (its thread joined) first
  std::unique_ptr<HlmPgzip> pg_;
  bool corrupt_ = false;
  std::vector<char> buf_;
  size_t pos_ = 0, have_ = 0;
  bool eof_ = false;
};

// bounded producer queue of batches
class BatchStream {
 public:
  BatchStream(const char* path, int64_t batch_pairs) : rd_(path), batch_(batch_pairs) {
    if (rd_.ok()) th_ = std::thread([this] { run(); });
  }
  ~BatchStream() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  bool ok() const { return rd_.ok(); }
  // nullptr at end of file
  MateBatch* next() {
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return !q_.empty() || done_; });
    if (q_.empty()) return nullptr;
    MateBatch* b = q_.front();
    q_.pop();
    cv_.notify_all();
    return b;
  }
  void recycle(MateBatch* b) {
    std::lock_guard<std::mutex> g(m_);
    free_.push_back(b);
    cv_.notify_all();
  }

 private:
  void run() {
    try {
      run_unguarded();
    } catch (const std::exception&) {  // out of memory on the reader thread: the run ends with an
      emergency_.clear();              // input error instead of std::terminate or a short input
      emergency_.io_error = true;
      std::lock_guard<std::mutex> g(m_);
      q_.push(&emergency_);
      done_ = true;
      cv_.notify_all();
    }
  }
  void run_unguarded() {
    for (;;) {
      MateBatch* b = nullptr;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || q_.size() < 2; });
        if (stop_) break;
        if (!free_.empty()) {
          b = free_.back();
          free_.pop_back();
        }
      }
      if (!b) {
        pool_.emplace_back(new MateBatch());
        b = pool_.back().get();
      }
      b->clear();
      int64_t want = batch_size(k_++, batch_);
      int64_t got = rd_.read(*b, want);
      std::lock_guard<std::mutex> g(m_);
      if (got > 0 || b->io_error) q_.push(b);
      else free_.push_back(b);
      if (got < want) {
        done_ = true;
        cv_.notify_all();
        break;
      }
      cv_.notify_all();
    }
  }
  FastqReader rd_;
  int64_t batch_, k_ = 0;
  std::thread th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::queue<MateBatch*> q_;
  std::vector<MateBatch*> free_;
  std::vector<std::unique_ptr<MateBatch>> pool_;
  MateBatch emergency_;
  bool stop_ = false, done_ = false;
};

// r1 (+ r2) FASTQ files, one reader thread each
class FileSource : public PairSource {
 public:
  FileSource(const char* r1, const char* r2, int64_t batch_pairs) : in1_(r1, batch_pairs) {
    if (r2) in2_.reset(new BatchStream(r2, batch_pairs));
  }
  bool ok1() const { return in1_.ok(); }
  bool ok2() const { return !in2_ || in2_->ok(); }
  bool next(MateBatch*& b1, MateBatch*& b2, int& rc, std::string& err) override {
    b1 = in1_.next();
    b2 = in2_ ? in2_->next() : nullptr;
    if ((b1 && b1->io_error) || (b2 && b2->io_error)) {
      err = std::string(b1 && b1->io_error ? "r1" : "r2") + " is corrupt or truncated (gzip / BGZF stream)";
      rc = HLM_E_IO;
      return false;
    }
    bool end = !b1 || (in2_ && !b2);
    if (in2_ && ((end && (b1 != nullptr) != (b2 != nullptr)) || (!end && b1->n != b2->n))) {
      err = "r1 and r2 hold different numbers of reads";
      rc = HLM_E_IO;
      return false;
    }
    return !end;
  }
  void recycle(MateBatch* b1, MateBatch* b2) override {
    in1_.recycle(b1);
    if (in2_) in2_->recycle(b2);
  }

 private:
  BatchStream in1_;
  std::unique_ptr<BatchStream> in2_;
};

// Plain (uncompressed) FASTQ files: mapped and parsed by several host threads.  A newline count
// per region (parallel) places every batch boundary - record b * B starts after newline 4 b B - so
// any thread can parse any batch of either file on its own; batches are delivered in order.
// Same record rules as FastqReader: four lines per record, split at '\n' only; a last line
// without newline completes its record; an incomplete record at the end is dropped.
class MappedFastq {
 public:
  bool open(const char* path, int64_t batch, int nt) {
    if (!map_.open(path)) return false;
    if (map_.n >= 2 && map_.p[0] == 0x1f && map_.p[1] == 0x8b) return false;  // gzip / BGZF: not here
    const char* p = (const char*)map_.p;
    size_t n = map_.n;
    nt = std::max(1, nt);
    size_t R = (size_t)nt * 4;  // regions
    std::vector<size_t> cut(R + 1), cnt(R + 1, 0);
    for (size_t r = 0; r <= R; r++) cut[r] = n * r / R;
    parallel_for(nt, [&](int t) {
      for (size_t r = (size_t)t; r < R; r += (size_t)nt) {
        size_t c = 0;
        const char* q = p + cut[r];
        const char* e = p + cut[r + 1];
        while (q < e) {
          const char* nl = (const char*)memchr(q, '\n', (size_t)(e - q));
          if (!nl) break;
          c++;
          q = nl + 1;
        }
        cnt[r + 1] = c;
      }
    });
    for (size_t r = 0; r < R; r++) cnt[r + 1] += cnt[r];
    size_t lines = cnt[R] + ((n > 0 && p[n - 1] != '\n') ? 1 : 0);
    n_records_ = (int64_t)(lines / 4);
    rec_ = batch_plan(n_records_, batch);
    int64_t nb = (int64_t)rec_.size() - 1;
    start_.assign((size_t)nb + 1, n);
    // byte offset of record k = one past newline number 4 k (1-based), 0 for k = 0
    auto offset_of_record = [&](int64_t k) -> size_t {
      if (k == 0) return 0;
      size_t want = (size_t)k * 4;  // this many newlines precede the record
      if (want > cnt[R]) return n;
      size_t r = (size_t)(std::upper_bound(cnt.begin(), cnt.end(), want - 1) - cnt.begin()) - 1;  // cnt[r] <= want-1 < cnt[r+1]
      size_t need = want - cnt[r];
      const char* q = p + cut[r];
      const char* e = p + cut[r + 1];
      while (q < e) {
        const char* nl = (const char*)memchr(q, '\n', (size_t)(e - q));
        if (!nl) break;
        q = nl + 1;
        if (--need == 0) return (size_t)(q - p);
      }
      return n;
    };
    parallel_for(nt, [&](int t) {
      for (int64_t b = t; b <= nb; b += nt) start_[(size_t)b] = offset_of_record(rec_[(size_t)b]);
    });
    return true;
  }
  int64_t n_records() const { return n_records_; }
  int64_t n_batches() const { return (int64_t)start_.size() - 1; }
  // parse batch b into out (cleared first)
  void parse(int64_t b, MateBatch& out) const {
    out.clear();
    const char* base = (const char*)map_.p;
    size_t p = start_[(size_t)b], end = map_.n;
    int64_t want = rec_[(size_t)b + 1] - rec_[(size_t)b];
    out.offs.reserve((size_t)want + 1);
    out.hoff.reserve((size_t)want + 1);
    size_t span = start_[(size_t)b + 1] - p;
    out.bases.reserve(span / 2);
    out.quals.reserve(span / 2);
    for (int64_t k = 0; k < want; k++) {
      size_t e[4], q = p;
      for (int x = 0; x < 4; x++) {
        const char* nl = (const char*)memchr(base + q, '\n', end - q);
        e[x] = nl ? (size_t)(nl - base) : end;
        q = nl ? e[x] + 1 : end;
      }
      size_t s0 = e[0] + 1, q0 = e[2] + 1;
      size_t hl = e[0] - p, sl = e[1] - s0, ql = e[3] - q0;
      if (k == 0) out.hdr.reserve((size_t)want * (hl + 4));  // sized from the first header line
      out.hdr.insert(out.hdr.end(), base + p, base + p + hl);
      out.hoff.push_back(out.hdr.size());
      out.bases.insert(out.bases.end(), (const uint8_t*)base + s0, (const uint8_t*)base + s0 + sl);
      size_t qq = std::min(sl, ql);
      out.quals.insert(out.quals.end(), (const uint8_t*)base + q0, (const uint8_t*)base + q0 + qq);
      if (qq < sl) out.quals.insert(out.quals.end(), sl - qq, (uint8_t)'!');
      out.offs.push_back(out.bases.size());
      p = q;
    }
    out.n = want;
  }

 private:
  HlmMapped map_;
  std::vector<size_t> start_;
  std::vector<int64_t> rec_;  // first record of every batch (batch_plan)
  int64_t n_records_ = 0;
};

// r1 (+ r2) plain FASTQ files parsed by a pool of host threads, batches delivered in order
class MappedSource : public PairSource {
 public:
  MappedSource(int64_t batch, int nt) : batch_(batch), nt_(std::max(1, nt)) {}
  ~MappedSource() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // false: not plain FASTQ (compressed, empty, unreadable) - the caller falls back to FileSource
  bool open(const char* r1, const char* r2) {
    if (!f1_.open(r1, batch_, nt_)) return false;
    paired_ = r2 != nullptr;
    if (paired_ && !f2_.open(r2, batch_, nt_)) return false;
    return true;
  }
  bool counts_agree() const { return !paired_ || f1_.n_records() == f2_.n_records(); }
  void start() {
    nb_ = f1_.n_batches();
    int workers = std::max(1, std::min(nt_, 8));
    for (int w = 0; w < workers; w++) th_.emplace_back([this] { work(); });
  }
  bool next(MateBatch*& b1, MateBatch*& b2, int& rc, std::string& err) override {
    std::unique_lock<std::mutex> g(m_);
    if (deliver_ >= nb_) return false;
    const int M = paired_ ? 2 : 1;
    cv_.wait(g, [&] {
      auto it = ready_.find(deliver_);
      return failed_ || (it != ready_.end() && it->second.done == M);
    });
    if (failed_) {  // a parser thread ran out of memory
      rc = HLM_E_NOMEM;
      err = "out of memory while parsing the FASTQ files";
      return false;
    }
    auto it = ready_.find(deliver_);
    b1 = it->second.m[0];
    b2 = it->second.m[1];
    ready_.erase(it);
    deliver_++;
    cv_.notify_all();
    return true;
  }
  void recycle(MateBatch* b1, MateBatch* b2) override {
    std::lock_guard<std::mutex> g(m_);
    if (b1) free_.push_back(b1);
    if (b2) free_.push_back(b2);
  }

 private:
  MateBatch* fresh() {  // m_ held
    if (!free_.empty()) {
      MateBatch* b = free_.back();
      free_.pop_back();
      return b;
    }
    pool_.emplace_back(new MateBatch());
    return pool_.back().get();
  }
  // a work item is one mate file of one batch, so that the two files of a batch are parsed at the
  // same time (the consumer waits for the slower of the two, not for their sum)
  void work() {
    const int M = paired_ ? 2 : 1;
    for (;;) {
      int64_t b;
      int mate;
      MateBatch* mb;
      {
        std::unique_lock<std::mutex> g(m_);
        // at most kAhead batches parsed ahead of the consumer
        cv_.wait(g, [&] { return stop_ || claim_ >= nb_ * M || claim_ / M < deliver_ + kAhead; });
        if (stop_ || claim_ >= nb_ * M) return;
        b = claim_ / M;
        mate = (int)(claim_ % M);
        claim_++;
        mb = fresh();
        ready_[b].m[mate] = mb;
      }
      try {
        (mate == 0 ? f1_ : f2_).parse(b, *mb);
        std::lock_guard<std::mutex> g(m_);
        ready_[b].done++;
      } catch (const std::exception&) {  // no exception leaves a library thread
        std::lock_guard<std::mutex> g(m_);
        failed_ = true;
        stop_ = true;
      }
      cv_.notify_all();
    }
  }
  struct Parsed {
    MateBatch* m[2] = {nullptr, nullptr};
    int done = 0;
  };
  static constexpr int64_t kAhead = 6;
  MappedFastq f1_, f2_;
  bool paired_ = false;
  int64_t batch_, nb_ = 0, claim_ = 0, deliver_ = 0;
  int nt_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::map<int64_t, Parsed> ready_;
  std::vector<MateBatch*> free_;
  std::vector<std::unique_ptr<MateBatch>> pool_;
  bool stop_ = false, failed_ = false;
};

void erase_all(std::string& s, const char* pat) {  // boost::replace_all(s, pat, "")
  size_t n = strlen(pat), p = 0;
  while ((p = s.find(pat, p)) != std::string::npos) s.erase(p, n);
}

// stable bump allocator for the per-read records that are kept until the end of the run
class Arena {
 public:
  char* put(const char* src, size_t n) {
    char* d = alloc(n);
    memcpy(d, src, n);
    return d;
  }
  char* alloc(size_t n) {
    if (blocks_.empty() || used_ + n > cap_) {
      cap_ = (std::max<size_t>(n, (size_t)64 << 20) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
      void* m = aligned_alloc((size_t)2 << 20, cap_);
      if (!m) throw std::bad_alloc();
      madvise(m, cap_, MADV_HUGEPAGE);  // first-touch faults dominate the collection otherwise
      blocks_.emplace_back((char*)m);
      used_ = 0;
    }
    char* d = blocks_.back().get() + used_;
    used_ += n;
    return d;
  }

 private:
  struct Free {
    void operator()(char* p) const { free(p); }
  };
  std::vector<std::unique_ptr<char, Free>> blocks_;
  size_t cap_ = 0, used_ = 0;
};

struct Kept {
  const char *idgen, *h1, *h2, *s1, *q1, *s2, *q2;  // arena pointers
  const uint16_t *sc1, *sc2;                        // per target
  uint32_t idl, tokl, h1l, h2l, l1, l2;             // tok = h1[0 .. tokl)
  uint16_t lo1, n1, lo2, n2;
  uint32_t locus_mask, target, visible;
  uint16_t os1, os2;
};

// the kept reads in id order: a plain array whose pages are first touched by the threads that fill
// it (a std::vector would zero 100+ MB on one thread first)
struct KeptVec {
  std::unique_ptr<Kept[]> p;
  size_t n = 0;
  void allocate(size_t count) {
    p.reset(new Kept[count]);
    n = count;
  }
  size_t size() const { return n; }
  Kept& operator[](size_t i) { return p[i]; }
  const Kept& operator[](size_t i) const { return p[i]; }
  Kept* begin() { return p.get(); }
  Kept* end() { return p.get() + n; }
  const Kept* begin() const { return p.get(); }
  const Kept* end() const { return p.get() + n; }
};

struct Out {
  FILE* f = nullptr;
  std::string buf;
  bool bad = false;  // open, a short write or the close failed (ENOSPC, permissions ...)
  bool used = false;
  bool open(const std::string& p) {
    f = fopen(p.c_str(), "wb");
    buf.reserve(1 << 22);
    used = true;
    if (!f) bad = true;
    return f != nullptr;
  }
  bool failed() const { return used && bad; }
  void flush() {
    if (f && !buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) bad = true;
    buf.clear();
  }
  void add(const std::string& s) { add(s.data(), s.size()); }
  void add(const char* s, size_t n) {
    buf.append(s, n);
    if (buf.size() > (1u << 22)) flush();
  }
  void fill(char c, size_t n) {
    buf.append(n, c);
    if (buf.size() > (1u << 22)) flush();
  }
  void close() {
    flush();
    if (f && fclose(f) != 0) bad = true;
    f = nullptr;
  }
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// result arrays of one batch
// Zero-filled arrays that are never touched before the results land in them: calloc() of a large
// block hands out untouched zero pages, so growing to the size of the largest batch costs nothing
// up front (a std::vector would write 100 MB of zeros on the driver thread, between two batches).
template <class Tp>
struct ZeroArray {
  Tp* p = nullptr;
  size_t cap = 0;
  ZeroArray() {}
  ZeroArray(const ZeroArray&) = delete;
  ZeroArray& operator=(const ZeroArray&) = delete;
  ~ZeroArray() { free(p); }
  void resize(size_t n) {  // contents are not kept: every batch rewrites them
    if (n <= cap) return;
    free(p);
    cap = std::max(n, cap * 2);
    p = (Tp*)calloc(cap, sizeof(Tp));
    if (!p) {
      cap = 0;
      throw std::bad_alloc();
    }
  }
  Tp* data() { return p; }
  const Tp* data() const { return p; }
  Tp* begin() { return p; }
  Tp& operator[](size_t i) { return p[i]; }
  const Tp& operator[](size_t i) const { return p[i]; }
};

struct Results {
  ZeroArray<uint8_t> flags;
  ZeroArray<uint16_t> lo1, len1, lo2, len2, s1, s2, minsum, os1, os2;
  ZeroArray<uint32_t> lmask, target, visible;
  std::vector<int32_t> so;
  hlm_batch hb;
  hlm_screen_out scr;
  hlm_score_out sc;
  hlm_assign_out as;
  MateBatch *b1 = nullptr, *b2 = nullptr;
  void bind(MateBatch* m1, MateBatch* m2, int T, bool paired) {
    b1 = m1;
    b2 = m2;
    int64_t n = m1->n;
    flags.resize(n); lo1.resize(n); len1.resize(n); lo2.resize(n); len2.resize(n);
    lmask.resize(n); target.resize(n); visible.resize(n); minsum.resize(n); os1.resize(n); os2.resize(n);
    s1.resize((size_t)n * T); s2.resize((size_t)n * T);
    memset(&hb, 0, sizeof hb);
    hb.n_pairs = n;
    hb.bases1 = m1->bases.data(); hb.quals1 = m1->quals.data(); hb.offs1 = m1->offs.data();
    if (paired) {
      hb.bases2 = m2->bases.data(); hb.quals2 = m2->quals.data(); hb.offs2 = m2->offs.data();
    } else {
      // single end: sequence_size_r1 = length of the header line (src/preselect.cpp:1199)
      so.resize(n);
      for (int64_t i = 0; i < n; i++) so[i] = (int32_t)(m1->hoff[i + 1] - m1->hoff[i]);
      hb.size_override1 = so.data();
    }
    scr = {flags.data(), lo1.data(), len1.data(), lo2.data(), len2.data(), lmask.data()};
    memset(&sc, 0, sizeof sc);
    sc.s1 = s1.data();
    sc.s2 = s2.data();
    as = {target.data(), visible.data(), minsum.data(), os1.data(), os2.data()};
  }
};

}  // namespace

extern "C" int hlm_screen_trim(hlm_ctx*, const hlm_batch*, hlm_screen_out*);
extern "C" int hlm_run_async(hlm_ctx*, const hlm_batch*, hlm_screen_out*, hlm_score_out*, hlm_assign_out*);
extern "C" int hlm_wait(hlm_ctx*);
extern "C" int hlm_get_profile(const hlm_ctx*, hlm_profile*);

extern "C" int hlm_score(hlm_ctx*, const hlm_batch*, const hlm_screen_out*, hlm_score_out*);
extern "C" int hlm_assign(hlm_ctx*, const hlm_batch*, const hlm_screen_out*, const hlm_score_out*,
                          hlm_assign_out*);

namespace {

// ---- rna: STAR between the screen and the scorer (src/map_rna.cpp:648-801) -------------------
// STAR's per-gene options, src/map_rna.cpp:676-722
struct StarOpt {
  const char *mnmax, *readLmax, *mismatchN, *seed;
};
StarOpt star_options(const std::string& g) {
  StarOpt o{"70", "1", "999", "60"};
  if (g == "HLA-G" || g == "HLA-E" || g == "HLA-F") o = {"30", "0.1", "999", "50"};
  if (g == "HLA-DMA" || g == "HLA-DMB" || g == "HLA-DOA") o = {"30", "0.1", "10", "50"};
  if (g == "HLA-DOB" || g == "MICA" || g == "MICB") o = {"30", "0.1", "10", "50"};
  if (g == "TAP1" || g == "TAP2" || g == "HLA-H") o = {"50", "0.1", "999", "50"};
  if (g == "HLA-DRA" || g == "HLA-DPA1") o = {"50", "0.1", "999", "50"};
  return o;
}

// The (start, size) ranges of SEQ that src/map_rna.cpp:748-787 writes to sorted_spliced_<gene>.fastq
// for one SAM record: the CIGAR is cut after every N; a piece's size is the sum of its operation
// lengths except N and D (S, H, I, M, P, X count); piece k starts at the SIZE of piece k-1 (`start =
// read_size`, :784 - not cumulative), piece 0 at 0.  A CIGAR without N: the whole SEQ.
void star_record_pieces(const char* cig, size_t cl, size_t seq_len, std::vector<std::pair<int, int>>& out) {
  out.clear();
  if (!memchr(cig, 'N', cl)) {
    out.emplace_back(0, (int)seq_len);
    return;
  }
  int start = 0;
  size_t p = 0;
  while (p < cl) {
    const char* e = (const char*)memchr(cig + p, 'N', cl - p);
    size_t end = e ? (size_t)(e - cig) + 1 : cl;
    int read_size = 0;
    size_t q = p;
    while (q < end) {  // splitcigar(): a token ends with one of S H M D I N P X
      size_t t = q;
      while (t < end && !strchr("SHMDINPX", cig[t])) t++;
      if (t >= end) break;  // trailing characters without an operation: an unsplit token, no type
      int v = 0;            // stoi(): the leading digits of the token
      for (size_t x = q; x < t && cig[x] >= '0' && cig[x] <= '9'; x++) v = v * 10 + (cig[x] - '0');
      if (cig[t] != 'N' && cig[t] != 'D') read_size += v;
      q = t + 1;
    }
    out.emplace_back(start, read_size);
    start = read_size;
    p = end;
  }
}

struct RnaBlocks {
  std::vector<std::vector<hlm_rna_block>> of;  // per kept read
};

// <pre><gene>_splice_Aligned.out.sam -> blocks of the kept reads + sorted_spliced_<gene>.fastq
int read_star_sam(const std::string& sam_path, const std::string& spliced_path, int gene,
                  const KeptVec& kept, RnaBlocks& rb, std::mutex& mu, std::string& err) {
  HlmMapped map;
  Out sp;
  if (!sp.open(spliced_path)) {
    err = "cannot write " + spliced_path;
    return HLM_E_IO;
  }
  if (!map.open(sam_path.c_str())) {  // no SAM (STAR wrote nothing): no records for this gene
    sp.close();
    return sp.failed() ? HLM_E_IO : 0;
  }
  const char* base = (const char*)map.p;
  size_t n = map.n, p = 0;
  std::vector<std::pair<int, int>> pieces;
  std::vector<std::pair<int64_t, hlm_rna_block>> found;
  char num[32];
  while (p < n) {
    const char* nl = (const char*)memchr(base + p, '\n', n - p);
    size_t e = nl ? (size_t)(nl - base) : n;
    size_t ls = p;
    p = e + 1;
    if (e == ls || base[ls] == '@') continue;
    // the first eleven tab-separated fields
    const char* f[12];
    size_t fl[12];
    int nf = 0;
    size_t a = ls;
    while (nf < 11 && a <= e) {
      const char* t = (const char*)memchr(base + a, '\t', e - a);
      size_t b = t ? (size_t)(t - base) : e;
      f[nf] = base + a;
      fl[nf] = b - a;
      nf++;
      a = b + 1;
    }
    if (nf < 11) continue;
    int flag = atoi(std::string(f[1], fl[1]).c_str());
    size_t L = fl[9];
    star_record_pieces(f[5], fl[5], L, pieces);
    // which kept read: QNAME is the id (first header token, /1 /2 removed)
    int64_t lo = 0, hi = (int64_t)kept.size();
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      const Kept& k = kept[mid];
      int c = memcmp(k.idgen, f[0], std::min<size_t>(k.idl, fl[0]));
      if (c < 0 || (c == 0 && k.idl < fl[0])) lo = mid + 1;
      else hi = mid;
    }
    bool known = lo < (int64_t)kept.size() && kept[lo].idl == fl[0] && memcmp(kept[lo].idgen, f[0], fl[0]) == 0;
    // Which mate, and in which orientation, SEQ is: by content (SEQ is the trimmed mate as STAR
    // read it from sorted_<gene>_R*.fastq, or its reverse complement); the FLAG bits 0x80 / 0x10
    // only break ties and stand in when the content matches neither (clipped / edited SEQ)
    int mate = (flag & 0x80) ? 1 : 0;
    bool rev = (flag & 0x10) != 0;
    if (known) {
      const Kept& k = kept[lo];
      auto same = [&](const char* m, size_t ml, bool rc) {
        if (ml != L) return false;
        if (!rc) return memcmp(m, f[9], L) == 0;
        for (size_t x = 0; x < L; x++) {
          char c = m[L - 1 - x], w;
          switch (c) {
            case 'A': w = 'T'; break; case 'C': w = 'G'; break; case 'G': w = 'C'; break; case 'T': w = 'A'; break;
            case 'a': w = 't'; break; case 'c': w = 'g'; break; case 'g': w = 'c'; break; case 't': w = 'a'; break;
            default: w = c;
          }
          if (w != f[9][x]) return false;
        }
        return true;
      };
      const char* ms[2] = {k.s1 + k.lo1, k.s2 ? k.s2 + k.lo2 : nullptr};
      size_t ml[2] = {k.n1, k.n2};
      bool done = false;
      for (int pass = 0; pass < 4 && !done; pass++) {
        int m = (pass & 1) ? 1 - mate : mate;
        bool r = (pass & 2) ? !rev : rev;
        if (ms[m] && same(ms[m], ml[m], r)) {
          mate = m;
          rev = r;
          done = true;
        }
      }
    }
    int blk = 1;
    for (auto& pc : pieces) {
      size_t st = std::min<size_t>((size_t)pc.first, L), en = std::min<size_t>((size_t)pc.first + (size_t)pc.second, L);
      sp.add("@", 1); sp.add(f[0], fl[0]); sp.add("$", 1); sp.add(f[1], fl[1]);
      int m = snprintf(num, sizeof num, "$%d\n", blk++);
      sp.add(num, m);
      sp.add(f[9] + st, en - st); sp.add("\n+\n", 3);
      size_t qs = std::min<size_t>(st, fl[10]), qe = std::min<size_t>(en, fl[10]);
      sp.add(f[10] + qs, qe - qs); sp.add("\n", 1);
      if (!known) continue;
      hlm_rna_block b;
      b.gene = (uint8_t)gene;
      b.mate = (uint8_t)mate;
      // a record STAR printed reverse-complemented covers [L - en, L - st) of the mate itself
      b.start = (uint16_t)(rev ? L - en : st);
      b.len = (uint16_t)(en - st);
      b.reserved = 0;
      found.emplace_back(lo, b);
    }
  }
  sp.close();
  if (sp.failed()) {
    err = "write error on " + spliced_path;
    return HLM_E_IO;
  }
  std::lock_guard<std::mutex> g(mu);
  for (auto& fb : found) rb.of[fb.first].push_back(fb.second);
  return 0;
}

std::string shq(const std::string& s) {  // single-quoted for sh
  std::string o = "'";
  for (char c : s) {
    if (c == '\'') o += "'\\''";
    else o += c;
  }
  return o + "'";
}

// the second half of `hla-mapper rna` for the kept reads (already sorted by id): per-gene sorted
// FASTQs -> STAR (when a command is given) -> block lists -> hlm_score + hlm_assign
int rna_score_kept(hlm_ctx* ctx, const hlm_db* db, const hlm_file_opts& o, const std::string& pre, bool paired,
                   KeptVec& kept) {
  const char* tag = paired ? "R1" : "R0";
  int G = db->n_loci, T = G + 1;
  int nt = o.threads > 0 ? o.threads : (int)std::max(1u, std::thread::hardware_concurrency());
  nt = std::max(1, std::min(nt, G));
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  RnaBlocks rb;
  rb.of.resize(kept.size());
  std::mutex mu;
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int w = 0; w < nt; w++)
    th.emplace_back([&] {
      for (int g; (g = next.fetch_add(1)) < G;) {
        try {
          const std::string& gene = db->names[g];
          std::string in1 = pre + "sorted_" + gene + "_" + tag + ".fastq", in2 = pre + "sorted_" + gene + "_R2.fastq";
          // ---- sorted_<gene>_R*.fastq (src/map_rna.cpp:544-636): the trimmed records of every
          // kept read that holds a 20-mer of the gene
          Out s1, s2;
          if (!s1.open(in1) || (paired && !s2.open(in2))) {
            rcs[g] = HLM_E_IO;
            errs[g] = "cannot write " + in1;
            continue;
          }
          for (const Kept& k : kept) {
            if (!((k.locus_mask >> g) & 1u)) continue;
            s1.add(k.h1, k.h1l); s1.add("\n", 1); s1.add(k.s1 + k.lo1, k.n1); s1.add("\n+\n", 3);
            s1.fill('A', k.n1); s1.add("\n", 1);
            if (paired) {
              s2.add(k.h2, k.h2l); s2.add("\n", 1); s2.add(k.s2 + k.lo2, k.n2); s2.add("\n+\n", 3);
              s2.fill('A', k.n2); s2.add("\n", 1);
            }
          }
          s1.close();
          s2.close();
          if (s1.failed() || s2.failed()) {
            rcs[g] = HLM_E_IO;
            errs[g] = "write error on " + in1;
            continue;
          }
          // ---- STAR, one run per gene with the reference's arguments (src/map_rna.cpp:724-733)
          if (o.star_cmd && o.star_cmd[0]) {
            StarOpt so = star_options(gene);
            std::string sub(pre.begin() + (pre.rfind('/') + 1), pre.end() - 1);  // the sample name
            std::string prefix = pre + gene + (paired ? "_splice_" : "_");
            std::string cmd = std::string(o.star_cmd) + " --runThreadN 1 --genomeDir " +
                              shq(db->dir + "/reference/rna/loci/" + gene + "/") + " --readFilesIn " + shq(in1) +
                              (paired ? " " + shq(in2) : std::string()) + " --outFileNamePrefix " + shq(prefix) +
                              " --seedPerWindowNmax " + so.seed + " --outFilterMismatchNmax " + so.mismatchN +
                              " --outFilterMultimapNmax " + so.mnmax + " --outFilterMismatchNoverReadLmax " +
                              so.readLmax +
                              " --outSAMunmapped Within --scoreDelOpen -1 --scoreDelBase -1 --scoreInsOpen -1"
                              " --scoreInsBase -1 --outSAMattrRGline ID:" + sub + " LB:" + sub + " SM:" + sub +
                              " > " + shq(pre + gene + ".splice.log") + " 2>&1";
            int sr = system(cmd.c_str());
            if (sr != 0) {
              rcs[g] = HLM_E_IO;
              errs[g] = "STAR failed for " + gene + " (see " + pre + gene + ".splice.log)";
              continue;
            }
          }
          // ---- its SAM -> sorted_spliced_<gene>.fastq + the block list (src/map_rna.cpp:735-793;
          // the reference opens <gene>_splice_Aligned.out.sam for single-end runs too)
          rcs[g] = read_star_sam(pre + gene + "_splice_Aligned.out.sam", pre + "sorted_spliced_" + gene + ".fastq", g,
                                 kept, rb, mu, errs[g]);
        } catch (const std::exception& e) {
          rcs[g] = HLM_E_NOMEM;
          errs[g] = e.what();
        }
      }
    });
  for (auto& t : th) t.join();
  for (int g = 0; g < G; g++)
    if (rcs[g]) {
      hlm_ctx_set_error(ctx, errs[g]);
      return rcs[g];
    }
  // ---- score + assign from memory, in chunks of kept reads
  const int64_t CH = 1 << 18;
  std::vector<uint8_t> b1, b2, flags;
  std::vector<uint64_t> o1, o2, boff;
  std::vector<uint16_t> lo1, n1, lo2, n2, s1, ms, os1, os2;
  std::vector<uint32_t> lm, tg, vis;
  std::vector<hlm_rna_block> blocks;
  for (int64_t a = 0; a < (int64_t)kept.size(); a += CH) {
    int64_t b = std::min<int64_t>((int64_t)kept.size(), a + CH), n = b - a;
    b1.clear(); b2.clear(); blocks.clear();
    o1.assign(1, 0); o2.assign(1, 0); boff.assign(1, 0);
    flags.assign(n, HLM_F_SCREENED | HLM_F_KEPT);
    lo1.resize(n); n1.resize(n); lo2.resize(n); n2.resize(n); lm.resize(n);
    for (int64_t i = a; i < b; i++) {
      const Kept& k = kept[i];
      b1.insert(b1.end(), (const uint8_t*)k.s1, (const uint8_t*)k.s1 + k.l1);
      o1.push_back(b1.size());
      if (paired) {
        b2.insert(b2.end(), (const uint8_t*)k.s2, (const uint8_t*)k.s2 + k.l2);
        o2.push_back(b2.size());
      }
      lo1[i - a] = k.lo1; n1[i - a] = k.n1; lo2[i - a] = k.lo2; n2[i - a] = k.n2; lm[i - a] = k.locus_mask;
      blocks.insert(blocks.end(), rb.of[i].begin(), rb.of[i].end());
      boff.push_back(blocks.size());
    }
    if (blocks.empty()) blocks.resize(1);  // a valid pointer for an empty list
    hlm_batch hb;
    memset(&hb, 0, sizeof hb);
    hb.n_pairs = n;
    hb.bases1 = b1.data(); hb.offs1 = o1.data();
    if (paired) { hb.bases2 = b2.data(); hb.offs2 = o2.data(); }
    hb.rna_block_off = boff.data();
    hb.rna_blocks = blocks.data();
    hlm_screen_out scr = {flags.data(), lo1.data(), n1.data(), lo2.data(), n2.data(), lm.data()};
    s1.assign((size_t)n * T, 0);
    tg.resize(n); vis.resize(n); ms.resize(n); os1.resize(n); os2.resize(n);
    hlm_score_out sc;
    memset(&sc, 0, sizeof sc);
    sc.s1 = s1.data();
    hlm_assign_out as = {tg.data(), vis.data(), ms.data(), os1.data(), os2.data()};
    int rc = hlm_score(ctx, &hb, &scr, &sc);
    if (rc == 0) rc = hlm_assign(ctx, &hb, &scr, &sc, &as);
    if (rc) return rc;
    for (int64_t i = a; i < b; i++) {
      Kept& k = kept[i];
      memcpy(const_cast<uint16_t*>(k.sc1), s1.data() + (size_t)(i - a) * T, 2 * (size_t)T);
      k.target = tg[i - a];
      k.visible = vis[i - a];
      k.os1 = os1[i - a];
      k.os2 = os2[i - a];
    }
  }
  return 0;
}

}  // namespace

extern "C" int hlm_map_fastq(hlm_ctx* ctx, const char* r1_path, const char* r2_path, const char* sample,
                             const char* out_dir, const hlm_file_opts* opts_in, hlm_file_stats* stats) {
  if (!ctx || !r1_path || !sample || !out_dir) return HLM_E_ARG;
  const hlm_params* P = hlm_ctx_params(ctx);
  bool paired = P->paired != 0;
  if (paired != (r2_path != nullptr)) {
    hlm_ctx_set_error(ctx, "paired context needs r1 and r2; single-end context needs r2 == NULL");
    return HLM_E_ARG;
  }
  int64_t batch = opts_in && opts_in->struct_size >= 12 && opts_in->batch_pairs > 0 ? opts_in->batch_pairs : 1 << 19;
  try {  // no exception crosses the C ABI
    {  // plain FASTQ: mapped and parsed by several threads; anything else: one reader thread per file
      MappedSource ms(batch, (int)std::max(1u, std::thread::hardware_concurrency()));
      if (ms.open(r1_path, r2_path)) {
        if (!ms.counts_agree()) {
          hlm_ctx_set_error(ctx, "r1 and r2 hold different numbers of reads");
          return HLM_E_IO;
        }
        ms.start();
        return hlm_map_source(ctx, ms, paired, sample, out_dir, opts_in, stats);
      }
    }
    FileSource src(r1_path, r2_path, batch);
    if (!src.ok1() || !src.ok2()) {
      hlm_ctx_set_error(ctx, std::string("cannot open ") + (src.ok1() ? r2_path : r1_path));
      return HLM_E_IO;
    }
    return hlm_map_source(ctx, src, paired, sample, out_dir, opts_in, stats);
  } catch (const std::exception& e) {
    hlm_ctx_set_error(ctx, std::string("hlm_map_fastq: ") + e.what());
    return HLM_E_NOMEM;
  }
}

namespace {

// A file whose records are produced by several threads at once: every thread formats a contiguous
// range of the records into a local buffer and writes it with pwrite() at the offset the ranges
// before it end at (sizes from a first pass over the same records).
class ParallelFile {
 public:
  bool open(const std::string& p) {
    fd_ = ::open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    return fd_ >= 0;
  }
  ~ParallelFile() {
    if (fd_ >= 0) ::close(fd_);
  }
  bool write_at(const char* d, size_t n, uint64_t off) {
    while (n) {
      ssize_t w = pwrite(fd_, d, n, (off_t)off);
      if (w <= 0) {
        bad_ = true;
        return false;
      }
      d += w;
      n -= (size_t)w;
      off += (uint64_t)w;
    }
    return true;
  }
  bool close() {
    if (fd_ >= 0 && ::close(fd_) != 0) bad_ = true;
    fd_ = -1;
    return !bad_;
  }
  bool bad() const { return bad_; }

 private:
  int fd_ = -1;
  std::atomic<bool> bad_{false};
};

// buffered writer of one thread's range of a ParallelFile
struct RangeOut {
  // bytes formatted between two pwrite() calls (256 KB and 1 MB were measured: no difference)
  static constexpr size_t kBuf = (size_t)1 << 22;
  ParallelFile* f = nullptr;
  uint64_t off = 0;
  std::string buf;
  void begin(ParallelFile* file, uint64_t at) {
    f = file;
    off = at;
    buf.clear();
    buf.reserve(kBuf + 4096);
  }
  void add(const char* s, size_t n) {
    buf.append(s, n);
    if (buf.size() > kBuf) flush();
  }
  void fill(char c, size_t n) {
    buf.append(n, c);
    if (buf.size() > kBuf) flush();
  }
  void raw(const char* s, size_t n) {  // a block that is already text: no copy through the buffer
    flush();
    if (f && n) f->write_at(s, n, off);
    off += n;
  }
  void flush() {
    if (f && !buf.empty()) f->write_at(buf.data(), buf.size(), off);
    off += buf.size();
    buf.clear();
  }
};

// The kept reads, collected slice by slice, into one vector in the reference's std::map order
// (byte-wise on the id; stable: equal ids keep their input order).  Only 16-byte keys are sorted -
// runs on nt host threads, then pairwise merges - and the records are gathered once at the end.
struct IdKey {
  const char* s;
  uint64_t k;  // the eight id bytes that follow the prefix ALL ids share, big-endian, zero-padded
  uint32_t l, part;
  uint32_t idx, pad;
};
// ids agree on their first `w0` bytes, so (k, then the whole id) orders like the whole id; most
// comparisons are decided by k without touching the id bytes (Illumina ids share a long prefix)
bool key_less(const IdKey& a, const IdKey& b) {
  if (a.k != b.k) return a.k < b.k;
  int c = memcmp(a.s, b.s, std::min(a.l, b.l));
  return c < 0 || (c == 0 && a.l < b.l);
}
uint64_t id_window(const char* s, uint32_t l, uint32_t w0) {
  uint64_t k = 0;
  for (uint32_t x = 0; x < 8; x++) k = (k << 8) | (w0 + x < l ? (uint8_t)s[w0 + x] : 0u);
  return k;
}
// an array whose elements are not value-initialised (a std::vector would zero tens of MB on one
// thread before the parallel code touches them)
template <class Tp>
struct RawArray {
  std::unique_ptr<Tp[]> p;
  size_t n = 0;
  explicit RawArray(size_t count = 0) : p(count ? new Tp[count] : nullptr), n(count) {}
  Tp* begin() { return p.get(); }
  Tp* end() { return p.get() + n; }
  Tp& operator[](size_t i) { return p[i]; }
  size_t size() const { return n; }
  void swap(RawArray& o) {
    p.swap(o.p);
    std::swap(n, o.n);
  }
};

// keys into id order (stable) on nt host threads
void sort_keys(RawArray<IdKey>& keys, int nt, size_t min_parallel = (size_t)1 << 15) {
  // sample sort: splitters from a regular sample, every key classified once, a stable scatter
  // into buckets (thread-major inside a bucket, so equal ids keep their input order), every
  // bucket sorted on its own - no merge passes over the whole array
  const size_t n = keys.size();
  if (n < min_parallel || nt <= 1) {
    std::stable_sort(keys.begin(), keys.end(), key_less);
  } else {
    const size_t NB = (size_t)std::min(256, 4 * nt), over = 32, S = NB * over;
    std::vector<IdKey> sample(S);
    for (size_t i = 0; i < S; i++) sample[i] = keys[(i * n + n / 2) / S];
    std::sort(sample.begin(), sample.end(), key_less);
    std::vector<IdKey> split(NB - 1);
    for (size_t j = 1; j < NB; j++) split[j - 1] = sample[j * over];
    RawArray<uint8_t> bucket(n);
    std::vector<size_t> count((size_t)nt * NB, 0);
    parallel_for(nt, [&](int t) {
      size_t* c = count.data() + (size_t)t * NB;
      for (size_t j = n * (size_t)t / nt; j < n * (size_t)(t + 1) / nt; j++) {
        // number of splitters <= key: equal keys always land in the same bucket
        size_t bk = (size_t)(std::upper_bound(split.begin(), split.end(), keys[j], key_less) - split.begin());
        bucket[j] = (uint8_t)bk;
        c[bk]++;
      }
    });
    std::vector<size_t> bstart(NB + 1, 0), at((size_t)nt * NB);
    for (size_t bk = 0; bk < NB; bk++) {
      size_t o = bstart[bk];
      for (int t = 0; t < nt; t++) {
        at[(size_t)t * NB + bk] = o;
        o += count[(size_t)t * NB + bk];
      }
      bstart[bk + 1] = o;
    }
    RawArray<IdKey> tmp(n);
    parallel_for(nt, [&](int t) {
      size_t* o = at.data() + (size_t)t * NB;
      for (size_t j = n * (size_t)t / nt; j < n * (size_t)(t + 1) / nt; j++) tmp[o[bucket[j]]++] = keys[j];
    });
    std::atomic<size_t> next(0);
    parallel_for(nt, [&](int) {
      for (size_t bk; (bk = next.fetch_add(1)) < NB;)
        std::stable_sort(tmp.begin() + bstart[bk], tmp.begin() + bstart[bk + 1], key_less);
    });
    keys.swap(tmp);
  }
}

void sort_kept(std::vector<std::vector<Kept>>& parts, KeptVec& out, int nt) {
  size_t n = 0;
  std::vector<size_t> first(parts.size() + 1, 0);
  for (size_t p = 0; p < parts.size(); p++) first[p + 1] = first[p] + parts[p].size();
  n = first[parts.size()];
  RawArray<IdKey> keys(n);
  nt = std::max(1, nt);
  // the prefix every id shares with the first one (exact: a minimum over all ids)
  const Kept* k0 = nullptr;
  for (auto& v : parts)
    if (!v.empty()) {
      k0 = &v[0];
      break;
    }
  std::vector<uint32_t> lcp(nt, k0 ? k0->idl : 0);
  parallel_for(nt, [&](int t) {
    uint32_t m = lcp[t];
    for (size_t p = (size_t)t; p < parts.size(); p += (size_t)nt)
      for (const Kept& k : parts[p]) {
        uint32_t x = 0, lim = std::min(m, k.idl);
        while (x < lim && k.idgen[x] == k0->idgen[x]) x++;
        m = x;
      }
    lcp[t] = m;
  });
  uint32_t w0 = k0 ? *std::min_element(lcp.begin(), lcp.end()) : 0;
  parallel_for(nt, [&](int t) {
    for (size_t p = (size_t)t; p < parts.size(); p += (size_t)nt)
      for (size_t i = 0; i < parts[p].size(); i++)
        keys[first[p] + i] = IdKey{parts[p][i].idgen, id_window(parts[p][i].idgen, parts[p][i].idl, w0),
                                   parts[p][i].idl, (uint32_t)p, (uint32_t)i, 0};
  });
  sort_keys(keys, nt);
  out.allocate(n);
  parallel_for(nt, [&](int t) {
    for (size_t j = n * (size_t)t / nt; j < n * (size_t)(t + 1) / nt; j++) out[j] = parts[keys[j].part][keys[j].idx];
  });
  // the slices are released by the threads as well (one thread would spend milliseconds per GB)
  parallel_for(nt, [&](int t) {
    for (size_t p = (size_t)t; p < parts.size(); p += (size_t)nt) std::vector<Kept>().swap(parts[p]);
  });
  std::vector<std::vector<Kept>>().swap(parts);
}

}  // namespace

// Self test of the FASTQ sources (tests/test_host_sources.py; not part of the interface): the
// mapped, multi-threaded source (batch boundaries from the newline count, batch sizes ramped up at
// the start and down at the end) and the sequential reader must deliver the same records in the
// same order - headers, bases, qualities - for the same files, the two mates of a batch always
// side by side; the sequential reader may be given the same reads compressed (gzip: inflated by
// several threads, hlm_pgzip.h; bgzip: hlm_bgzf.h).  0 when they do, else the failed check.
extern "C" int hlm_selftest_fastq_sources(const char* r1, const char* r2, int64_t batch, int threads,
                                          int64_t* n_batches, int64_t* n_records, const char* seq1,
                                          const char* seq2) {
  try {
    struct All {
      std::vector<char> hdr[2];
      std::vector<uint8_t> bases[2], quals[2];
      std::vector<uint64_t> hl[2], sl[2];
      int64_t batches = 0, records = 0;
      bool ok = true;
      void add(const MateBatch* m, int k) {
        if (!m) return;
        hdr[k].insert(hdr[k].end(), m->hdr.begin(), m->hdr.end());
        bases[k].insert(bases[k].end(), m->bases.begin(), m->bases.end());
        quals[k].insert(quals[k].end(), m->quals.begin(), m->quals.end());
        if ((int64_t)m->offs.size() != m->n + 1 || (int64_t)m->hoff.size() != m->n + 1) ok = false;
        for (int64_t i = 0; i < m->n && ok; i++) {
          hl[k].push_back(m->hoff[i + 1] - m->hoff[i]);
          sl[k].push_back(m->offs[i + 1] - m->offs[i]);
        }
      }
    } A, B;
    auto drain = [&](PairSource& src, All& out) -> int {
      for (;;) {
        MateBatch *m1 = nullptr, *m2 = nullptr;
        int rc = 0;
        std::string err;
        if (!src.next(m1, m2, rc, err)) return rc ? 4 : 0;
        if (m2 && m1->n != m2->n) return 5;
        out.add(m1, 0);
        out.add(m2, 1);
        out.batches++;
        out.records += m1->n;
        src.recycle(m1, m2);
      }
    };
    {
      MappedSource ms(batch, threads);
      if (!ms.open(r1, r2)) return 1;
      if (!ms.counts_agree()) return 2;
      ms.start();
      if (int rc = drain(ms, A)) return rc;
    }
    {
      // seq1 / seq2: the same reads in other files (gzip, bgzip) for the sequential reader
      FileSource fs(seq1 ? seq1 : r1, seq1 ? seq2 : r2, batch);
      if (!fs.ok1() || !fs.ok2()) return 3;
      if (int rc = drain(fs, B)) return 10 + rc;
    }
    if (!A.ok || !B.ok) return 6;
    if (A.records != B.records) return 7;
    for (int k = 0; k < 2; k++)
      if (A.hdr[k] != B.hdr[k] || A.bases[k] != B.bases[k] || A.quals[k] != B.quals[k] || A.hl[k] != B.hl[k] ||
          A.sl[k] != B.sl[k])
        return 8 + k;
    if (n_batches) *n_batches = A.batches;
    if (n_records) *n_records = A.records;
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// Self test of the parallel gzip reader (tests/test_pgzip.py; not part of the interface): the
// file inflated by HlmPgzip on `threads` threads with chunks of `chunk_kb` KiB of compressed data,
// and by zlib's gzread.  0: same text (its length and CRC-32 are returned, with the number of
// chunks and of chunks that had to be decoded a second time); 1: the texts differ; 2: HlmPgzip
// reports an error (`zlib_ok` tells whether zlib read the file without one); -1: out of memory.
extern "C" int hlm_selftest_pgzip(const char* path, int threads, int chunk_kb, int64_t* length, uint32_t* crc,
                                  int64_t* n_chunks, int64_t* n_slow, int* zlib_ok, double* seconds) {
  try {
    std::vector<uint8_t> want;
    bool zok = true;
    {
      gzFile f = gzopen(path, "rb");
      if (!f) return 3;
      gzbuffer(f, 1 << 20);
      std::vector<uint8_t> buf((size_t)1 << 22);
      for (;;) {
        int r = gzread(f, buf.data(), (unsigned)buf.size());
        if (r < 0) zok = false;
        if (r <= 0) break;
        want.insert(want.end(), buf.begin(), buf.begin() + r);
      }
      int en = Z_OK;
      gzerror(f, &en);
      if (en != Z_OK && en != Z_STREAM_END) zok = false;
      gzclose(f);
    }
    if (zlib_ok) *zlib_ok = zok ? 1 : 0;
    HlmMapped m;
    if (!m.open(path)) return 3;
    char kb[32];
    snprintf(kb, sizeof kb, "%d", chunk_kb);
    if (chunk_kb > 0) setenv("HLM_PGZIP_CHUNK_KB", kb, 1);
    double t0 = now_s();
    HlmPgzip pg(m.p, m.n, threads);
    if (chunk_kb > 0) unsetenv("HLM_PGZIP_CHUNK_KB");
    std::vector<uint8_t> got, buf((size_t)3 << 20);
    for (;;) {
      long r = pg.fill(buf.data(), buf.size());
      if (r < 0) return 2;
      if (r == 0) break;
      got.insert(got.end(), buf.begin(), buf.begin() + r);
    }
    if (seconds) *seconds = now_s() - t0;
    if (n_chunks) *n_chunks = pg.n_chunks;
    if (n_slow) *n_slow = pg.n_slow;
    if (length) *length = (int64_t)got.size();
    if (crc) *crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), got.data(), (uInt)got.size());
    return got == want ? 0 : 1;
  } catch (const std::exception&) {
    return -1;
  }
}

// Reader throughput (tools/reader_bench.py; not part of the interface): the sequential-reader
// source drained over r1 (+ r2), whatever their compression; seconds and records out.
extern "C" int hlm_selftest_reader_speed(const char* r1, const char* r2, int64_t batch, double* seconds,
                                         int64_t* n_records, int64_t* n_bases) {
  try {
    double t0 = now_s();
    FileSource fs(r1, r2, batch);
    if (!fs.ok1() || !fs.ok2()) return 3;
    int64_t nr = 0, nbs = 0;
    for (;;) {
      MateBatch *m1 = nullptr, *m2 = nullptr;
      int rc = 0;
      std::string err;
      if (!fs.next(m1, m2, rc, err)) {
        if (rc) return 4;
        break;
      }
      nr += m1->n;
      nbs += (int64_t)m1->bases.size() + (m2 ? (int64_t)m2->bases.size() : 0);
      fs.recycle(m1, m2);
    }
    if (seconds) *seconds = now_s() - t0;
    if (n_records) *n_records = nr;
    if (n_bases) *n_bases = nbs;
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// Self test of hlm_crc32 (csrc/hlm_crc32.h; tests/test_pgzip.py): random buffers, offsets, lengths
// and start values against zlib's crc32().  Returns the number of disagreements; *fast tells
// whether the carry-less-multiplication path is in use on this CPU.
extern "C" int hlm_selftest_crc32(int iterations, uint32_t seed, int* fast) {
  std::vector<uint8_t> d((size_t)1 << 20);
  uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&] {
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return x;
  };
  for (auto& b : d) b = (uint8_t)rnd();
  int bad = 0;
  for (int it = 0; it < iterations; it++) {
    size_t off = rnd() % 100, n = rnd() % (it % 8 ? 3000 : d.size() - 100);
    uint32_t start = it % 3 ? (uint32_t)rnd() : 0;
    if ((uint32_t)crc32(start, d.data() + off, (uInt)n) != hlm_crc32(start, d.data() + off, n)) bad++;
  }
#ifdef HLM_CRC32_CLMUL
  if (fast) *fast = hlm_crc32_clmul_ok();
#else
  if (fast) *fast = 0;
#endif
  return bad;
}

// Self test of the parallel id sort (tests/test_host_sort.py; not part of the interface): n random
// ids with a shared prefix, duplicates and different lengths, sorted by sort_keys() on `threads`
// threads and by one std::stable_sort; 0 when both orders are the same.
extern "C" int hlm_selftest_sort_keys(int64_t n, uint32_t seed, int threads) {
  try {
    std::vector<std::string> ids((size_t)n);
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&] {
      x ^= x << 13;
      x ^= x >> 7;
      x ^= x << 17;
      return x;
    };
    for (auto& id : ids) {
      id = "A00296:39:HFH3TDSXX:";
      uint64_t r = rnd();
      int extra = 1 + (int)(r % 14);
      uint64_t v = (r >> 8) % (uint64_t)std::max<int64_t>(1, n / 2 + 1);  // about every second id twice
      for (int k = 0; k < extra; k++) {
        id += (char)('0' + v % 10);
        v /= 10;
      }
    }
    RawArray<IdKey> keys((size_t)n);
    const uint32_t w0 = 20;  // the shared prefix
    for (size_t i = 0; i < (size_t)n; i++)
      keys[i] = IdKey{ids[i].data(), id_window(ids[i].data(), (uint32_t)ids[i].size(), w0), (uint32_t)ids[i].size(), 0,
                      (uint32_t)i, 0};
    std::vector<IdKey> want(keys.begin(), keys.end());
    std::stable_sort(want.begin(), want.end(), key_less);
    sort_keys(keys, threads, 64);
    for (size_t i = 0; i < (size_t)n; i++)
      if (keys[i].idx != want[i].idx) return 1;
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

std::vector<int64_t> hlm_batch_plan(int64_t N, int64_t B) { return batch_plan(N, B); }

int hlm_map_source(hlm_ctx* ctx, PairSource& src, bool normalise_pair_headers, const char* sample,
                   const char* out_dir, const hlm_file_opts* opts_in, hlm_file_stats* stats) {
  return hlm_map_source_multi(&ctx, 1, src, normalise_pair_headers, sample, out_dir, opts_in, stats);
}

// One driver thread per context pulls batches from the shared source (round robin by arrival),
// runs them on its GPU (batch k+1 in flight while batch k is collected) and hands the collected
// batch to an in-order commit, so that whatever the number of GPUs the files are those of a
// single-GPU run.  Errors are reported on ctxs[0].
int hlm_map_source_multi(hlm_ctx* const* ctxs, int n_ctx, PairSource& src, bool normalise_pair_headers,
                         const char* sample, const char* out_dir, const hlm_file_opts* opts_in,
                         hlm_file_stats* stats) {
  hlm_ctx* ctx = ctxs[0];
  const hlm_db* db = hlm_ctx_db(ctx);
  const hlm_params* P = hlm_ctx_params(ctx);
  bool paired = P->paired != 0;
  // rna (src/map_rna.cpp): the scoring step consumes what STAR made of the per-gene sorted reads,
  // so the reads are screened + trimmed first, STAR runs per gene between the two halves, and the
  // kept reads are scored from memory afterwards (rna_score_kept)
  const bool rna = P->rna != 0;
  hlm_file_opts o;
  memset(&o, 0, sizeof o);
  o.struct_size = (int32_t)sizeof o;
  o.downsampling = 30;  // src/main.cpp:86
  o.batch_pairs = 1 << 19;
  o.write_selected = 1;
  if (opts_in) memcpy(&o, opts_in, std::min((size_t)opts_in->struct_size, sizeof o));
  if (o.batch_pairs <= 0) o.batch_pairs = 1 << 19;
  if (o.downsampling < 5) o.downsampling = 5;  // src/main.cpp:434
  const bool screen_only = o.select_only || rna;
  int T = db->n_loci + 1;
  std::string out = std::string(out_dir);
  if (out.empty() || out.back() != '/') out += "/";  // src/main.cpp:438-448
  std::string pre = out + sample + "_";               // v_sample = sample + "_" (main.cpp:450-455)
  const char* tag = paired ? "R1" : "R0";
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());

  double t_start = now_s();
  // HLM_FILE_TRACE: the moments of the run on stderr (seconds from its start), a measurement aid
  const bool ftrace = getenv("HLM_FILE_TRACE") != nullptr;
  std::mutex ftrace_mu;
  auto mark = [&](const char* what, int64_t a = -1) {
    if (!ftrace) return;
    std::lock_guard<std::mutex> g(ftrace_mu);
    fprintf(stderr, "[hlm_files] %8.3f s  %s", now_s() - t_start, what);
    if (a >= 0) fprintf(stderr, " %lld", (long long)a);
    fputc('\n', stderr);
  };
  ParallelFile sel1, sel2;
  if (o.write_selected) {
    if (!sel1.open(pre + "selected." + tag + ".fastq") || (paired && !sel2.open(pre + "selected.R2.fastq"))) {
      hlm_ctx_set_error(ctx, "cannot write into " + out);
      return HLM_E_IO;
    }
  }
  // host threads that collect one finished batch (per driver)
  const int n_workers = std::max(1, std::min(16, hw / (2 * n_ctx) + 1));

  // what the reference keeps per selected read, taken from a finished batch: the batch is cut
  // into one slice per host thread; slices are committed in order, so the output is in input order
  struct Part {
    std::vector<Kept> kept;
    std::string sel1, sel2;
    int64_t n_selected = 0, sel_bases = 0;
  };
  struct BatchOut {
    int64_t seq = 0, n = 0;
    std::vector<Part> parts;
  };
  // A batch most of whose reads are selected is kept alive instead of being copied read by read
  // (on-target input: the copy would be the whole batch anyway); the source then hands out a
  // fresh one.  Sparse batches (whole-genome input) are copied into the arena and recycled.
  auto collect_slice = [&](Results& R, int64_t lo, int64_t hi, Arena& arena, Part& P, bool retained) {
    MateBatch *b1 = R.b1, *b2 = R.b2;
    std::string h1, h2, idg;
    if (o.write_selected && retained && hi > lo) {  // most records of the slice will be written out
      P.sel1.reserve((size_t)(b1->hoff[hi] - b1->hoff[lo]) + 2 * (size_t)(b1->offs[hi] - b1->offs[lo]) + 6 * (size_t)(hi - lo));
      if (paired)
        P.sel2.reserve((size_t)(b2->hoff[hi] - b2->hoff[lo]) + 2 * (size_t)(b2->offs[hi] - b2->offs[lo]) + 6 * (size_t)(hi - lo));
    }
    for (int64_t i = lo; i < hi; i++) {
      if (!(R.flags[i] & HLM_F_SCREENED)) continue;
      P.n_selected++;
      h1.assign(b1->hdr.data() + b1->hoff[i], b1->hoff[i + 1] - b1->hoff[i]);
      if (paired) h2.assign(b2->hdr.data() + b2->hoff[i], b2->hoff[i + 1] - b2->hoff[i]);
      if (normalise_pair_headers) {  // src/preselect.cpp:664-665, :675-676
        if (h1.find('/') != std::string::npos) { erase_all(h1, "/1"); erase_all(h1, "/2"); }
        if (h2.find('/') != std::string::npos) { erase_all(h2, "/1"); erase_all(h2, "/2"); }
      }
      const char* s1p = (const char*)b1->bases.data() + b1->offs[i];
      const char* q1p = (const char*)b1->quals.data() + b1->offs[i];
      size_t L1 = b1->offs[i + 1] - b1->offs[i], L2 = 0;
      const char *s2p = nullptr, *q2p = nullptr;
      if (paired) {
        s2p = (const char*)b2->bases.data() + b2->offs[i];
        q2p = (const char*)b2->quals.data() + b2->offs[i];
        L2 = b2->offs[i + 1] - b2->offs[i];
      }
      P.sel_bases += (int64_t)L1;
      if (o.write_selected) {
        P.sel1.append(h1); P.sel1.append("\n", 1); P.sel1.append(s1p, L1); P.sel1.append("\n+\n", 3);
        P.sel1.append(q1p, L1); P.sel1.append("\n", 1);
        if (paired) {
          P.sel2.append(h2); P.sel2.append("\n", 1); P.sel2.append(s2p, L2); P.sel2.append("\n+\n", 3);
          P.sel2.append(q2p, L2); P.sel2.append("\n", 1);
        }
      }
      if (!(R.flags[i] & HLM_F_KEPT)) continue;
      P.kept.emplace_back();
      Kept& k = P.kept.back();
      size_t sp = h1.find(' ');
      k.tokl = (uint32_t)(sp == std::string::npos ? h1.size() : sp);
      idg.assign(h1, k.tokl ? 1 : 0, k.tokl ? k.tokl - 1 : 0);
      if (idg.find('/') != std::string::npos) {  // src/preselect.cpp:1076-1078, :1170-1172
        erase_all(idg, "/1");
        erase_all(idg, "/2");
      }
      // one arena record: h1 | h2 | idgen | scores [| s1 | q1 | s2 | q2 when the batch is not retained]
      size_t bytes = h1.size() + h2.size() + idg.size() + (retained ? 0 : 2 * L1 + 2 * L2);
      char* p = arena.alloc(bytes + 1 + 4 * (size_t)T);
      k.h1 = p; k.h1l = (uint32_t)h1.size(); memcpy(p, h1.data(), h1.size()); p += h1.size();
      k.h2 = p; k.h2l = (uint32_t)h2.size(); memcpy(p, h2.data(), h2.size()); p += h2.size();
      k.l1 = (uint32_t)L1;
      k.l2 = (uint32_t)L2;
      if (retained) {
        k.s1 = s1p; k.q1 = q1p; k.s2 = s2p ? s2p : p; k.q2 = q2p ? q2p : p;
      } else {
        k.s1 = p; memcpy(p, s1p, L1); p += L1;
        k.q1 = p; memcpy(p, q1p, L1); p += L1;
        k.s2 = p; if (L2) memcpy(p, s2p, L2); p += L2;
        k.q2 = p; if (L2) memcpy(p, q2p, L2); p += L2;
      }
      k.idgen = p; k.idl = (uint32_t)idg.size(); memcpy(p, idg.data(), idg.size()); p += idg.size();
      p += ((uintptr_t)p & 1);
      uint16_t* sc = (uint16_t*)p;
      memcpy(sc, R.s1.data() + (size_t)i * T, 2 * (size_t)T);
      memcpy(sc + T, R.s2.data() + (size_t)i * T, 2 * (size_t)T);
      k.sc1 = sc; k.sc2 = sc + T;
      k.lo1 = R.lo1[i]; k.n1 = R.len1[i]; k.lo2 = R.lo2[i]; k.n2 = R.len2[i];
      k.locus_mask = R.lmask[i]; k.target = R.target[i]; k.visible = R.visible[i];
      k.os1 = R.os1[i]; k.os2 = R.os2[i];
    }
  };

  // ---- shared state of the drivers
  std::mutex src_mu, commit_mu, err_mu;
  bool src_end = false;
  int64_t next_seq = 0;
  int rc_all = 0;
  std::string err_all;
  auto fail = [&](int code, const std::string& m) {
    std::lock_guard<std::mutex> g(err_mu);
    if (!rc_all) {
      rc_all = code;
      err_all = m;
    }
  };
  auto failed = [&] {
    std::lock_guard<std::mutex> g(err_mu);
    return rc_all != 0;
  };
  KeptVec kept;
  std::map<int64_t, std::unique_ptr<BatchOut>> pending;
  int64_t commit_seq = 0, n_pairs = 0, n_selected = 0, sel_bases = 0;
  double t_gpu = 0, t_post = 0, t_wait = 0;
  std::vector<std::vector<Arena>> arenas(n_ctx);
  for (auto& a : arenas) a = std::vector<Arena>(n_workers);

  // in-order commit: the selected records and the kept list follow the input order.  The selected
  // text goes to two writer threads (one per file, plain sequential writes) so that no driver
  // waits for the page cache; the kept records stay in their per-slice vectors until the sort.
  std::vector<std::vector<Kept>> kept_parts;
  std::mutex wq_mu;
  std::condition_variable wq_cv;
  std::deque<std::shared_ptr<BatchOut>> wq[2];
  bool wq_done = false;
  auto sel_writer = [&](int which) {
    ParallelFile& f = which == 0 ? sel1 : sel2;
    uint64_t off = 0;
    for (;;) {
      std::shared_ptr<BatchOut> b;
      {
        std::unique_lock<std::mutex> g(wq_mu);
        wq_cv.wait(g, [&] { return wq_done || !wq[which].empty(); });
        if (wq[which].empty()) return;
        b = std::move(wq[which].front());
        wq[which].pop_front();
      }
      for (Part& P : b->parts) {
        std::string& t = which == 0 ? P.sel1 : P.sel2;
        if (!t.empty()) f.write_at(t.data(), t.size(), off);
        off += t.size();
        std::string().swap(t);
      }
    }
  };
  std::vector<std::thread> sel_threads;
  struct Joiner {  // the writer threads never outlive this frame, whatever leaves it
    std::function<void()> f;
    ~Joiner() { f(); }
  } sel_joiner{[&] {
    {
      std::lock_guard<std::mutex> q(wq_mu);
      wq_done = true;
      wq_cv.notify_all();
    }
    for (auto& t : sel_threads)
      if (t.joinable()) t.join();
  }};
  if (o.write_selected) {
    sel_threads.emplace_back(sel_writer, 0);
    if (paired) sel_threads.emplace_back(sel_writer, 1);
  }
  auto commit = [&](std::unique_ptr<BatchOut> bo) {
    std::lock_guard<std::mutex> g(commit_mu);
    pending[bo->seq] = std::move(bo);
    while (!pending.empty() && pending.begin()->first == commit_seq) {
      std::shared_ptr<BatchOut> b(std::move(pending.begin()->second));
      pending.erase(pending.begin());
      commit_seq++;
      for (Part& P : b->parts) {
        n_selected += P.n_selected;
        sel_bases += P.sel_bases;
        kept_parts.emplace_back(std::move(P.kept));
      }
      n_pairs += b->n;
      if (o.write_selected) {
        std::lock_guard<std::mutex> q(wq_mu);
        wq[0].push_back(b);
        if (paired) wq[1].push_back(b);
        wq_cv.notify_all();
      }
    }
  };

  auto drive = [&](int di) {
    hlm_ctx* c = ctxs[di];
    Results res[2];
    int64_t seq_of[2] = {0, 0};
    int cur = 0;
    bool inflight = false;
    auto collect = [&](Results& R, int64_t seq) {
      int64_t n = R.b1->n, nsel = 0;
      for (int64_t i = 0; i < n; i++) nsel += (R.flags[i] & HLM_F_SCREENED) ? 1 : 0;
      bool retained = nsel * 2 > n;
      std::unique_ptr<BatchOut> bo(new BatchOut());
      bo->seq = seq;
      bo->n = n;
      bo->parts.resize(n_workers);
      parallel_for(n_workers, [&](int w) {
        collect_slice(R, n * w / n_workers, n * (w + 1) / n_workers, arenas[di][w], bo->parts[w], retained);
      });
      if (!retained) {
        std::lock_guard<std::mutex> g(src_mu);
        src.recycle(R.b1, R.b2);
      }
      commit(std::move(bo));
    };
    for (;;) {
      MateBatch *b1 = nullptr, *b2 = nullptr;
      bool end = true;
      int64_t seq = 0;
      double t0 = now_s();
      {
        std::lock_guard<std::mutex> g(src_mu);
        if (!src_end && !failed()) {
          int rc = 0;
          std::string e;
          end = !src.next(b1, b2, rc, e);
          if (rc) fail(rc, e);
          if (end) src_end = true;
          else seq = next_seq++;
        }
      }
      double dt_wait = now_s() - t0;
      mark(end ? "input ended" : "batch read, seq", end ? -1 : seq);
      if (inflight) {  // finish the previous batch on the GPU
        int r2 = hlm_wait(c);
        hlm_profile pr;
        if (hlm_get_profile(c, &pr) == 0) {
          std::lock_guard<std::mutex> g(err_mu);
          t_gpu += pr.ms_total * 1e-3;  // device time of the batch
        }
        inflight = false;
        if (r2) fail(r2, hlm_last_error(c));
        mark("gpu done, seq", seq_of[cur ^ 1]);
      }
      int prev = cur ^ 1;  // results of the batch that just finished
      bool have_prev = res[prev].b1 != nullptr;
      if (!end && !failed()) {
        // reads this build cannot hold (HLM_RAW_MAX raw bases) end the run here, with the read named,
        // before the batch reaches the GPU
        for (MateBatch* mb : {b1, b2}) {
          if (!mb) continue;
          for (int64_t i = 0; i < mb->n; i++)
            if (mb->offs[i + 1] - mb->offs[i] > HLM_RAW_MAX) {
              fail(HLM_E_TOOLONG, "read " + std::string(mb->hdr.data() + mb->hoff[i], mb->hoff[i + 1] - mb->hoff[i]) +
                                      " is " + std::to_string(mb->offs[i + 1] - mb->offs[i]) +
                                      " bases long: this build holds reads of up to " + std::to_string(HLM_RAW_MAX) +
                                      " bases (trimmed mates of up to " + std::to_string(HLM_MAX_READ) + ")");
              break;
            }
        }
      }
      if (!end && !failed()) {
        res[cur].bind(b1, b2, T, paired);
        seq_of[cur] = seq;
        int rc;
        if (screen_only) {  // `hla-mapper select`: main_preselect() only; rna: scored later
          rc = hlm_screen_trim(c, &res[cur].hb, &res[cur].scr);
          hlm_profile pr;
          if (rc == 0 && hlm_get_profile(c, &pr) == 0) {
            std::lock_guard<std::mutex> g(err_mu);
            t_gpu += pr.ms_total * 1e-3;
          }
          std::fill(res[cur].target.begin(), res[cur].target.begin() + b1->n, 0u);
          std::fill(res[cur].visible.begin(), res[cur].visible.begin() + b1->n, 0u);
        } else {
          rc = hlm_run_async(c, &res[cur].hb, &res[cur].scr, &res[cur].sc, &res[cur].as);
          inflight = rc == 0;
        }
        if (rc) fail(rc, hlm_last_error(c));
      }
      double t1 = now_s();
      if (have_prev && !failed()) {
        collect(res[prev], seq_of[prev]);
        mark("collected, seq", seq_of[prev]);
      }
      {
        std::lock_guard<std::mutex> g(err_mu);
        t_wait += dt_wait;
        t_post += now_s() - t1;
      }
      res[prev].b1 = nullptr;
      if (end || failed()) break;
      cur ^= 1;
    }
    if (inflight) hlm_wait(c);
  };
  {
    std::vector<std::thread> drivers;
    std::atomic<int> oom(0);
    for (int d = 0; d < n_ctx; d++)
      drivers.emplace_back([&, d] {
        try {
          drive(d);
        } catch (const std::exception& e) {
          fail(HLM_E_NOMEM, std::string("file pipeline: ") + e.what());
          hlm_wait(ctxs[d]);
        }
      });
    for (auto& t : drivers) t.join();
  }
  mark("drivers done");
  // the two writers of selected.R*.fastq may still be behind: they finish while the kept reads are
  // sorted and the other files prepared, and are waited for before the run returns
  bool sel_ok = true;
  auto finish_selected = [&] {
    sel_joiner.f();
    mark("selected files written");
    sel_ok = sel1.close();
    sel_ok = sel2.close() && sel_ok;
  };
  if (rc_all || rna) finish_selected();
  if (rc_all) {  // nothing half-written stays behind a failed run
    if (o.write_selected) {
      unlink((pre + "selected." + tag + ".fastq").c_str());
      if (paired) unlink((pre + "selected.R2.fastq").c_str());
    }
    hlm_ctx_set_error(ctx, err_all);
    return rc_all;
  }
  if (n_ctx > 1) {  // per-driver times were summed over the drivers
    t_post /= n_ctx;
    t_wait /= n_ctx;
  }

  double t0 = now_s();
  const int nt = std::max(1, std::min(32, hw));
  // The output files are created by a background thread while the kept reads are sorted and the
  // files prepared: creating a hundred files in a cold directory takes tens of milliseconds.
  // `paths` lists them in the order the jobs below are defined; a run that fails before the jobs
  // start removes what was created.  (rna: STAR runs in between, the files are created afterwards.)
  struct Opener {
    std::vector<std::string> paths;
    std::vector<std::unique_ptr<ParallelFile>> files;
    std::thread th;
    bool taken = false;
    void start() {
      files.resize(paths.size());
      th = std::thread([this] {
        try {
          for (size_t i = 0; i < paths.size(); i++) {
            std::unique_ptr<ParallelFile> f(new ParallelFile());
            if (f->open(paths[i])) files[i] = std::move(f);
          }
        } catch (const std::exception&) {  // the files left out are reported by the jobs
        }
      });
    }
    void join() {
      if (th.joinable()) th.join();
    }
    ~Opener() {
      join();
      if (taken) return;
      for (size_t i = 0; i < files.size(); i++)
        if (files[i]) {
          files[i]->close();
          unlink(paths[i].c_str());
        }
    }
  } opener;
  if (!rna) {
    for (int m = 0; m < (paired ? 2 : 1); m++) opener.paths.push_back(pre + "selected.trim." + (m == 0 ? tag : "R2") + ".fastq");
    if (!o.select_only) {
      opener.paths.push_back(pre + "addressing_table.txt");
      for (int g = 0; g < db->n_loci; g++)
        for (int m = 0; m < (paired ? 2 : 1); m++) {
          opener.paths.push_back(pre + db->names[g] + "_" + (m == 0 ? tag : "R2") + ".fastq");
          opener.paths.push_back(pre + db->names[g] + ".trim." + (m == 0 ? tag : "R2") + ".fastq");
        }
    }
    opener.start();
  }
  sort_kept(kept_parts, kept, nt);
  double t_sort = now_s() - t0;
  mark("kept reads sorted,", (int64_t)kept.size());
  if (rna && !o.select_only) {
    double tg = now_s();
    int rc = rna_score_kept(ctx, db, o, pre, paired, kept);
    t_gpu += now_s() - tg;
    if (rc) return rc;
  }
  // ---- the output files.  Every file is written by all host threads at once: a thread formats a
  // contiguous range of the (id-sorted) kept reads and writes it at the offset the ranges before
  // it end at.  A failed open / write / close (ENOSPC, permissions) ends the run with HLM_E_IO.
  const size_t NK = kept.size();
  std::vector<size_t> cut(nt + 1);
  for (int t = 0; t <= nt; t++) cut[t] = NK * (size_t)t / nt;
  int werr = 0;
  std::string werr_msg;
  auto wfail = [&](const std::string& m) {
    std::lock_guard<std::mutex> g(err_mu);
    if (!werr) {
      werr = HLM_E_IO;
      werr_msg = m;
    }
  };
  // Every file is a job: `sizes(t)` returns the bytes range t of the kept reads will produce,
  // `body(t, RangeOut&)` produces them at the offset the ranges before it end at.  All (file,
  // range) tasks go through one pool of host threads, ordered so that concurrent tasks belong to
  // different files (writes into one file serialise on its inode lock).
  struct Job {
    std::string path;
    std::function<uint64_t(int)> sizes;
    std::function<void(int, RangeOut&)> body;
    std::vector<uint64_t> sz;
    std::unique_ptr<ParallelFile> f;
  };
  std::vector<Job> jobs;
  auto write_file = [&](const std::string& path, std::function<uint64_t(int)> sizes,
                        std::function<void(int, RangeOut&)> body) {
    jobs.push_back(Job{path, std::move(sizes), std::move(body), {}, nullptr});
  };
  auto run_jobs = [&] {
    const size_t J = jobs.size();
    opener.join();
    opener.taken = true;  // from here on the files stay, as they did before they were pre-created
    for (size_t k = 0; k < J; k++) {
      Job& j = jobs[k];
      j.sz.assign(nt + 1, 0);
      if (k < opener.paths.size() && opener.paths[k] == j.path) {
        j.f = std::move(opener.files[k]);  // created in the background (null: could not be)
        if (!j.f) wfail("cannot write " + j.path);
      } else {
        j.f.reset(new ParallelFile());
        if (!j.f->open(j.path)) wfail("cannot write " + j.path);
      }
    }
    if (werr) return;
    std::atomic<size_t> next(0);
    parallel_for(nt, [&](int) {
      for (size_t k; (k = next.fetch_add(1)) < J * (size_t)nt;) jobs[k % J].sz[k / J + 1] = jobs[k % J].sizes((int)(k / J));
    });
    for (Job& j : jobs)
      for (int t = 0; t < nt; t++) j.sz[t + 1] += j.sz[t];
    mark("output sized, files:", (int64_t)J);
    next = 0;
    parallel_for(nt, [&](int) {
      RangeOut ro;
      for (size_t k; (k = next.fetch_add(1)) < J * (size_t)nt;) {
        Job& j = jobs[k % J];
        int t = (int)(k / J);
        if (j.sz[t + 1] == j.sz[t]) continue;
        ro.begin(j.f.get(), j.sz[t]);
        j.body(t, ro);
        ro.flush();
      }
    });
    for (Job& j : jobs)
      if (!j.f->close()) wfail("write error on " + j.path);
  };
  int64_t n_assigned = 0;
  int64_t read_mean_size = n_selected > 0 ? sel_bases / n_selected : 0;  // :1125-1133 (int division)
  std::vector<int64_t> assigned(nt, 0);
  std::vector<int64_t> downS(db->n_loci, 0);
  std::string header = "Read";
  // One pass over the kept reads per range (all ranges at once) prepares every file: the sizes of
  // the two trim files, the rows of the addressing table as text (formatted once, written later),
  // and per gene the list of its members with the bytes they take in the gene's files.  The
  // (file, range) jobs below only copy.
  struct RangePrep {
    uint64_t trim_bytes[2] = {0, 0};
    std::string table;
    std::vector<std::vector<uint32_t>> members;     // per gene: indices into kept, ascending
    std::vector<uint64_t> gene_bytes[2];            // per gene: bytes in <gene>_R1 / _R2.fastq
    std::vector<uint64_t> gene_trim_bytes[2];       // per gene: bytes in <gene>.trim.R1 / R2 (after the ranks)
    std::vector<int64_t> rank0;                     // per gene: members in the ranges before this one
  };
  std::vector<RangePrep> prep(nt);
  try {
    const int G = db->n_loci;
    for (int g = 0; g < T; g++) header += "\t" + db->names[g];
    header += "\tTarget\n";
    for (int g = 0; g < G; g++) {
      int64_t size = (int64_t)db->gene_opt_end[g] - db->gene_opt_start[g];
      downS[g] = read_mean_size > 0 ? ((int64_t)o.downsampling * size) / read_mean_size : 0;  // :1198-1200
    }
    auto put_u = [](char* p, unsigned v) {  // decimal, returns one past the last digit
      char tmp[12];
      int n = 0;
      do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
      } while (v);
      while (n) *p++ = tmp[--n];
      return p;
    };
    // row of <sample>_addressing_table.txt (src/map_dna.cpp:932-1051, src/map_rna.cpp:945-1049)
    size_t names_len = 0;
    for (int g = 0; g < T; g++) names_len += db->names[g].size() + 1;
    auto row = [&](const Kept& k, std::string& out, std::string& scratch, bool& any) {
      size_t need = k.idl + (size_t)T * 14 + names_len + 4;
      if (scratch.size() < need) scratch.resize(need);
      char* p0 = &scratch[0];
      char* p = p0;
      memcpy(p, k.idgen, k.idl);
      p += k.idl;
      for (int g = 0; g < T; g++) {
        *p++ = '\t';
        if (!((k.visible >> g) & 1u)) {
          *p++ = '-';
          continue;
        }
        unsigned a = g == T - 1 ? k.os1 : k.sc1[g], b = g == T - 1 ? k.os2 : k.sc2[g];
        if (rna) a = k.sc1[g];  // one summed score per target (src/map_rna.cpp:945-961)
        p = put_u(p, a);
        if (paired && !rna) {
          *p++ = ',';
          p = put_u(p, b);
        }
      }
      *p++ = '\t';
      any = false;
      for (int g = 0; g < T; g++)
        if ((k.target >> g) & 1u) {
          if (any) *p++ = ',';
          memcpy(p, db->names[g].data(), db->names[g].size());
          p += db->names[g].size();
          any = true;
        }
      if (!any) *p++ = '-';
      *p++ = '\n';
      out.append(p0, (size_t)(p - p0));
    };
    auto member = [&](int g, const Kept& k) {
      if (!((k.locus_mask >> g) & 1u)) return false;  // was in sorted_<gene>_R1.fastq
      if (!((k.target >> g) & 1u)) return false;       // sequence_address[(gene, id)]
      // the lookup key is the header token minus '@' (:1225): it only equals the id when the
      // token carries no /1 /2 (always true for paired input, whose headers were normalised)
      return k.tokl >= 1 && k.tokl - 1 == k.idl && memcmp(k.h1 + 1, k.idgen, k.idl) == 0;
    };
    parallel_for(nt, [&](int t) {
      RangePrep& R = prep[t];
      R.members.assign(G, {});
      for (int m = 0; m < 2; m++) {
        R.gene_bytes[m].assign(G, 0);
        R.gene_trim_bytes[m].assign(G, 0);
      }
      R.rank0.assign(G, 0);
      std::string scratch;
      if (!o.select_only) {
        if (t == 0) R.table = header;
        R.table.reserve(R.table.size() + (cut[t + 1] - cut[t]) * (size_t)(48 + 2 * T));
      }
      for (size_t i = cut[t]; i < cut[t + 1]; i++) {
        const Kept& k = kept[i];
        R.trim_bytes[0] += k.h1l + 2 * (uint64_t)k.n1 + 5;
        R.trim_bytes[1] += k.h2l + 2 * (uint64_t)k.n2 + 5;
        if (o.select_only) continue;
        bool any;
        row(k, R.table, scratch, any);
        assigned[t] += any ? 1 : 0;
        for (uint32_t m = k.locus_mask & k.target; m; m &= m - 1) {
          int g = __builtin_ctz(m);
          if (g >= G || !member(g, k)) continue;
          R.members[g].push_back((uint32_t)(i - cut[t]));
          R.gene_bytes[0][g] += k.h1l + 2 * (uint64_t)k.l1 + 5;
          R.gene_bytes[1][g] += k.h2l + 2 * (uint64_t)k.l2 + 5;
        }
      }
    });
    // the first downS members of a gene (in id order) also go to its .trim files
    if (!o.select_only) {
      for (int g = 0; g < G; g++) {
        int64_t r = 0;
        for (int t = 0; t < nt; t++) {
          prep[t].rank0[g] = r;
          r += (int64_t)prep[t].members[g].size();
        }
      }
      parallel_for(nt, [&](int t) {
        RangePrep& R = prep[t];
        for (int g = 0; g < G; g++) {
          int64_t r = R.rank0[g];
          for (uint32_t x : R.members[g]) {
            if (r++ >= downS[g]) break;
            const Kept& k = kept[cut[t] + x];
            R.gene_trim_bytes[0][g] += k.tokl + 8 + 2 * (uint64_t)k.n1 + 4;
            R.gene_trim_bytes[1][g] += k.tokl + 8 + 2 * (uint64_t)k.n2 + 4;
          }
        }
      });
    }
    mark("output prepared");
    // ---- selected.trim.R*.fastq (src/preselect.cpp:1121-1149, :1182-1203)
    for (int m = 0; m < (paired ? 2 : 1); m++)
      write_file(pre + "selected.trim." + (m == 0 ? tag : "R2") + ".fastq",
                 [&, m](int t) { return prep[t].trim_bytes[m]; },
                 [&, m](int t, RangeOut& w) {
                   for (size_t i = cut[t]; i < cut[t + 1]; i++) {
                     const Kept& k = kept[i];
                     if (m == 0) {
                       w.add(k.h1, k.h1l); w.add("\n", 1); w.add(k.s1 + k.lo1, k.n1); w.add("\n+\n", 3);
                       w.fill('A', k.n1); w.add("\n", 1);
                     } else {
                       w.add(k.h2, k.h2l); w.add("\n", 1); w.add(k.s2 + k.lo2, k.n2); w.add("\n+\n", 3);
                       w.fill('A', k.n2); w.add("\n", 1);
                     }
                   }
                 });
    // ---- <sample>_addressing_table.txt: the rows are already text
    if (!o.select_only)
      write_file(pre + "addressing_table.txt", [&](int t) { return (uint64_t)prep[t].table.size(); },
                 [&](int t, RangeOut& w) {
                   w.raw(prep[t].table.data(), prep[t].table.size());
                   std::string().swap(prep[t].table);
                 });
    // ---- per-gene FASTQ files (src/map_dna.cpp:1056-1351, src/map_rna.cpp:1072-1344)
    for (int g = 0; g < G && !o.select_only; g++) {
      const std::string& gene = db->names[g];
      for (int m = 0; m < (paired ? 2 : 1); m++) {
        write_file(pre + gene + "_" + (m == 0 ? tag : "R2") + ".fastq",
                   [&, g, m](int t) { return prep[t].gene_bytes[m][g]; },
                   [&, g, m](int t, RangeOut& w) {
                     for (uint32_t x : prep[t].members[g]) {
                       const Kept& k = kept[cut[t] + x];
                       if (m == 0) {
                         w.add(k.h1, k.h1l); w.add("\n", 1); w.add(k.s1, k.l1); w.add("\n+\n", 3); w.add(k.q1, k.l1);
                         w.add("\n", 1);
                       } else {
                         w.add(k.h2, k.h2l); w.add("\n", 1); w.add(k.s2, k.l2); w.add("\n+\n", 3); w.add(k.q2, k.l2);
                         w.add("\n", 1);
                       }
                     }
                   });
        write_file(pre + gene + ".trim." + (m == 0 ? tag : "R2") + ".fastq",
                   [&, g, m](int t) { return prep[t].gene_trim_bytes[m][g]; },
                   [&, g, m](int t, RangeOut& w) {
                     int64_t r = prep[t].rank0[g];
                     for (uint32_t x : prep[t].members[g]) {
                       if (r++ >= downS[g]) break;
                       const Kept& k = kept[cut[t] + x];
                       w.add(k.h1, k.tokl);
                       if (m == 0) {
                         w.add(" 1:trim\n", 8); w.add(k.s1 + k.lo1, k.n1); w.add("\n+\n", 3); w.fill('A', k.n1);
                       } else {
                         w.add(" 2:trim\n", 8); w.add(k.s2 + k.lo2, k.n2); w.add("\n+\n", 3); w.fill('A', k.n2);
                       }
                       w.add("\n", 1);
                     }
                   });
      }
    }
    run_jobs();
    mark("output written");
    if (!rna) finish_selected();
    for (int64_t a : assigned) n_assigned += a;
    if (werr) {
      hlm_ctx_set_error(ctx, werr_msg);
      return werr;
    }
    if (o.write_selected && !sel_ok) {
      hlm_ctx_set_error(ctx, "cannot write " + pre + "selected.*.fastq");
      return HLM_E_IO;
    }
    double t_write = now_s() - t0;
    if (stats) {
      memset(stats, 0, sizeof *stats);
      stats->n_pairs = n_pairs;
      stats->n_selected = n_selected;
      stats->n_kept = (int64_t)kept.size();
      stats->n_assigned = n_assigned;
      stats->read_mean_size = read_mean_size;
      stats->s_total = now_s() - t_start;
      stats->s_gpu = t_gpu;
      stats->s_wait_input = t_wait;
      stats->s_collect = t_post;
      stats->s_write = t_write;
      stats->s_sort = t_sort;
    }
  } catch (const std::exception& e) {
    hlm_ctx_set_error(ctx, std::string("output writer: ") + e.what());
    return HLM_E_NOMEM;
  }
  return HLM_OK;
}

extern "C" int hlm_multi_map_fastq(hlm_multi* mm, const char* r1_path, const char* r2_path, const char* sample,
                                   const char* out_dir, const hlm_file_opts* opts_in, hlm_file_stats* stats) {
  if (!mm || !r1_path || !sample || !out_dir) return HLM_E_ARG;
  int n = 0;
  hlm_ctx* const* cs = hlm_multi_contexts(mm, &n);
  if (n <= 0) return HLM_E_ARG;
  const hlm_params* P = hlm_ctx_params(cs[0]);
  bool paired = P->paired != 0;
  if (paired != (r2_path != nullptr)) {
    hlm_multi_set_error(mm, "paired contexts need r1 and r2; single-end contexts need r2 == NULL");
    return HLM_E_ARG;
  }
  int64_t batch = opts_in && opts_in->struct_size >= 12 && opts_in->batch_pairs > 0 ? opts_in->batch_pairs : 1 << 19;
  int rc;
  try {  // no exception crosses the C ABI
    MappedSource ms(batch, (int)std::max(1u, std::thread::hardware_concurrency()));
    if (ms.open(r1_path, r2_path)) {
      if (!ms.counts_agree()) {
        hlm_multi_set_error(mm, "r1 and r2 hold different numbers of reads");
        return HLM_E_IO;
      }
      ms.start();
      rc = hlm_map_source_multi(cs, n, ms, paired, sample, out_dir, opts_in, stats);
    } else {
      FileSource src(r1_path, r2_path, batch);
      if (!src.ok1() || !src.ok2()) {
        hlm_multi_set_error(mm, std::string("cannot open ") + (src.ok1() ? r2_path : r1_path));
        return HLM_E_IO;
      }
      rc = hlm_map_source_multi(cs, n, src, paired, sample, out_dir, opts_in, stats);
    }
  } catch (const std::exception& e) {
    hlm_multi_set_error(mm, std::string("hlm_multi_map_fastq: ") + e.what());
    return HLM_E_NOMEM;
  }
  if (rc) hlm_multi_set_error(mm, hlm_last_error(cs[0]));
  return rc;
}

// ---- public STAR SAM reader (include/hlamap.h): FLAG-based mate / orientation
extern "C" int64_t hlm_star_sam_read(const char* sam_path, int gene, hlm_star_rec** out) {
  if (!sam_path || !out || gene < 0 || gene > 255) return HLM_E_ARG;
  *out = nullptr;
  try {
    HlmMapped map;
    if (!map.open(sam_path)) return 0;
    const char* base = (const char*)map.p;
    size_t n = map.n, p = 0, pool = 0;
    struct Tmp {
      size_t q0;
      int32_t ql, flag, L, piece;
      hlm_rna_block b;
    };
    std::vector<Tmp> recs;
    std::vector<std::pair<int, int>> pieces;
    while (p < n) {
      const char* nl = (const char*)memchr(base + p, '\n', n - p);
      size_t e = nl ? (size_t)(nl - base) : n, ls = p;
      p = e + 1;
      if (e == ls || base[ls] == '@') continue;
      const char* f[11];
      size_t fl[11];
      int nf = 0;
      size_t a = ls;
      while (nf < 11 && a <= e) {
        const char* t = (const char*)memchr(base + a, '\t', e - a);
        size_t b = t ? (size_t)(t - base) : e;
        f[nf] = base + a;
        fl[nf] = b - a;
        nf++;
        a = b + 1;
      }
      if (nf < 11) continue;
      int flag = atoi(std::string(f[1], fl[1]).c_str());
      size_t L = fl[9];
      star_record_pieces(f[5], fl[5], L, pieces);
      int k = 1;
      for (auto& pc : pieces) {
        size_t st = std::min<size_t>((size_t)pc.first, L), en = std::min<size_t>((size_t)pc.first + (size_t)pc.second, L);
        Tmp t;
        t.q0 = (size_t)(f[0] - base);
        t.ql = (int32_t)fl[0];
        t.flag = flag;
        t.L = (int32_t)L;
        t.piece = k++;
        t.b.gene = (uint8_t)gene;
        t.b.mate = (uint8_t)((flag & 0x80) ? 1 : 0);
        t.b.start = (uint16_t)((flag & 0x10) ? L - en : st);
        t.b.len = (uint16_t)(en - st);
        t.b.reserved = 0;
        recs.push_back(t);
        pool += fl[0];
      }
    }
    if (recs.empty()) return 0;
    char* mem = (char*)malloc(recs.size() * sizeof(hlm_star_rec) + pool + 1);
    if (!mem) return HLM_E_NOMEM;
    hlm_star_rec* r = (hlm_star_rec*)mem;
    char* sp = mem + recs.size() * sizeof(hlm_star_rec);
    for (size_t i = 0; i < recs.size(); i++) {
      memcpy(sp, base + recs[i].q0, (size_t)recs[i].ql);
      r[i].qname = sp;
      r[i].qname_len = recs[i].ql;
      r[i].flag = recs[i].flag;
      r[i].seq_len = recs[i].L;
      r[i].piece = recs[i].piece;
      r[i].block = recs[i].b;
      sp += recs[i].ql;
    }
    *out = r;
    return (int64_t)recs.size();
  } catch (const std::exception&) {
    return HLM_E_NOMEM;
  }
}
extern "C" void hlm_star_free(hlm_star_rec* recs) { free(recs); }
