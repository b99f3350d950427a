// hlm_bgzf.h — BGZF (blocked gzip: RFC 1952 members of at most 64 KiB carrying their own size in
// a "BC" extra subfield; SAM/BAM specification section 4.1) read as independent deflate streams:
// the block boundaries are walked on the calling thread, the blocks inflated on all host
// threads straight into one contiguous buffer.  Used by the BAM front end (hlm_bam.cpp) and by
// the FASTQ reader for bgzip-compressed FASTQ files (hlm_files.cpp).
#ifndef HLM_BGZF_H
#define HLM_BGZF_H

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "hlm_pgzip.h"

struct HlmMapped {
  const uint8_t* p = nullptr;
  size_t n = 0;
  int fd = -1;
  bool open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) return false;
    n = (size_t)st.st_size;
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) return false;
    madvise(m, n, MADV_SEQUENTIAL);
    p = (const uint8_t*)m;
    return true;
  }
  HlmMapped() = default;
  HlmMapped(const HlmMapped&) = delete;
  HlmMapped& operator=(const HlmMapped&) = delete;
  ~HlmMapped() {
    if (p) munmap((void*)p, n);
    if (fd >= 0) close(fd);
  }
};

// does the file start with a BGZF block?
inline bool hlm_is_bgzf(const uint8_t* d, size_t n) {
  if (n < 18 || d[0] != 0x1f || d[1] != 0x8b || d[2] != 8 || !(d[3] & 4)) return false;
  uint32_t xlen = d[10] | (d[11] << 8);
  for (size_t x = 12; x + 4 <= 12 + (size_t)xlen && x + 6 <= n;) {
    uint32_t slen = d[x + 2] | (d[x + 3] << 8);
    if (d[x] == 'B' && d[x + 1] == 'C' && slen == 2) return true;
    x += 4 + slen;
  }
  return false;
}

class HlmBgzf {
 public:
  HlmBgzf(const uint8_t* data, size_t size, int threads) : d_(data), n_(size), threads_(threads < 1 ? 1 : threads) {}
  bool verify_crc = true;
  std::string err;
  bool at_end() const { return pos_ >= n_; }
  // Inflates the next blocks - at most max_blocks, and only whole blocks that fit `cap` bytes -
  // to dst.  Returns the number of bytes written (0 at the end of the file, or when the next
  // block does not fit) or -1 with `err` set.
  long fill(uint8_t* dst, size_t cap, int max_blocks) {
    struct Blk { size_t in, in_len, out; uint32_t isize, crc; };
    std::vector<Blk> blks;
    size_t out = 0;
    while (pos_ < n_ && (int)blks.size() < max_blocks) {
      size_t pos = pos_;
      if (n_ - pos < 18 || d_[pos] != 0x1f || d_[pos + 1] != 0x8b || d_[pos + 2] != 8 || !(d_[pos + 3] & 4))
        return fail("not a BGZF block at byte " + std::to_string(pos));
      uint32_t xlen = le16(d_ + pos + 10), bsize = 0;
      bool found = false;
      for (size_t x = pos + 12; x + 4 <= pos + 12 + xlen && x + 4 <= n_;) {
        uint32_t slen = le16(d_ + x + 2);
        if (d_[x] == 'B' && d_[x + 1] == 'C' && slen == 2 && x + 6 <= n_) {
          bsize = le16(d_ + x + 4);
          found = true;
          break;
        }
        x += 4 + slen;
      }
      size_t total = (size_t)bsize + 1;
      if (!found || total < 12 + xlen + 8 || pos + total > n_)
        return fail("truncated BGZF block at byte " + std::to_string(pos));
      Blk b;
      b.in = pos + 12 + xlen;
      b.in_len = total - 12 - xlen - 8;
      b.crc = le32(d_ + pos + total - 8);
      b.isize = le32(d_ + pos + total - 4);
      if (b.isize > 65536) return fail("BGZF block larger than 64 KiB at byte " + std::to_string(pos));
      if (out + b.isize > cap) break;
      b.out = out;
      out += b.isize;
      blks.push_back(b);
      pos_ += total;
    }
    if (blks.empty()) return 0;
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    // the blocks are inflated by the library's own decoder (hlm_pgzip.h: about twice zlib's speed on
    // sequencing data); HLM_BGZF_ZLIB=1 selects zlib's inflate
    static const bool use_zlib = getenv("HLM_BGZF_ZLIB") != nullptr;
    auto work_own = [&] {
      std::vector<uint8_t> scratch(65536 + 2048);
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= blks.size()) break;
        const Blk& b = blks[i];
        if (b.isize == 0) continue;
        if (!hlm_pgz::inflate_raw(d_ + b.in, b.in_len, scratch.data(), b.isize) ||
            (verify_crc && hlm_crc32(0, scratch.data(), b.isize) != b.crc)) {
          bad = 1;
          break;
        }
        memcpy(dst + b.out, scratch.data(), b.isize);
      }
    };
    auto work_zlib = [&] {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (inflateInit2(&zs, -15) != Z_OK) {
        bad = 1;
        return;
      }
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= blks.size()) break;
        const Blk& b = blks[i];
        if (b.isize == 0) continue;
        inflateReset(&zs);
        zs.next_in = const_cast<Bytef*>(d_ + b.in);
        zs.avail_in = (uInt)b.in_len;
        zs.next_out = dst + b.out;
        zs.avail_out = b.isize;
        int r = inflate(&zs, Z_FINISH);
        if (r != Z_STREAM_END || zs.avail_out != 0 ||
            (verify_crc && (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst + b.out, b.isize) != b.crc)) {
          bad = 1;
          break;
        }
      }
      inflateEnd(&zs);
    };
    auto work = [&] {
      if (use_zlib) work_zlib();
      else work_own();
    };
    int nt = (int)std::min<size_t>((size_t)threads_, blks.size());
    std::vector<std::thread> th;
    for (int w = 1; w < nt; w++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    if (bad) return fail("corrupt BGZF block (inflate / CRC)");
    return (long)out;
  }

 private:
  static uint32_t le32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
  static uint16_t le16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
  long fail(const std::string& m) {
    err = m;
    return -1;
  }
  const uint8_t* d_;
  size_t n_, pos_ = 0;
  int threads_;
};

// BGZF text inflated AHEAD of its consumer: one producer thread keeps two blocks of text (24 MB
// each, every one inflated on all of the reader's threads) ready while the caller parses the
// block before them.  fill(): up to `cap` bytes; 0 at the end of the file, -1 with the reader's
// `err` set.
class HlmBgzfAhead {
 public:
  explicit HlmBgzfAhead(HlmBgzf* bg) : bg_(bg) {
    th_ = std::thread([this] { run(); });
  }
  ~HlmBgzfAhead() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
  }
  HlmBgzfAhead(const HlmBgzfAhead&) = delete;
  HlmBgzfAhead& operator=(const HlmBgzfAhead&) = delete;
  long fill(uint8_t* dst, size_t cap) {
    size_t got = 0;
    while (got < cap) {
      if (at_ >= cur_.size()) {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return !q_.empty() || finished_; });
        if (q_.empty()) {
          if (failed_ && !got) return -1;
          break;
        }
        cur_.swap(q_.front());
        q_.pop_front();
        at_ = 0;
        cv_.notify_all();
        continue;
      }
      size_t k = std::min(cap - got, cur_.size() - at_);
      memcpy(dst + got, cur_.data() + at_, k);
      got += k;
      at_ += k;
    }
    return (long)got;
  }

 private:
  void run() {
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return stop_ || q_.size() < 2; });
        if (stop_) return;
      }
      std::vector<uint8_t> blk;
      long r = -1;
      try {
        blk.resize((size_t)24 << 20);
        r = bg_->fill(blk.data(), blk.size(), 384);
        if (r > 0) blk.resize((size_t)r);
      } catch (const std::exception&) {
        bg_->err = "out of memory while inflating";
        r = -1;
      }
      std::lock_guard<std::mutex> g(mu_);
      if (r > 0) q_.push_back(std::move(blk));
      if (r <= 0) {
        failed_ = r < 0;
        finished_ = true;
      }
      cv_.notify_all();
      if (r <= 0) return;
    }
  }
  HlmBgzf* bg_;
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::vector<uint8_t>> q_;
  std::vector<uint8_t> cur_;
  size_t at_ = 0;
  bool stop_ = false, finished_ = false, failed_ = false;
};

#endif
