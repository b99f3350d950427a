// hlm_bam.cpp — BAM front end of the hot path (second half).
//
// Replaces the four `samtools` processes and the SAM / FASTQ temp files of the reference's bam=
// input (src/preselect.cpp:315-447) with one in-process reader:
//
//   samtools view -ML <db>/bed/select_dna.bed bam     records overlapping a BED interval     :354
//   samtools view -f4 bam                             records with the unmapped flag         :383
//   (names of both sets -> read_names.txt)                                                  :365-403
//   samtools view -h -N read_names.txt bam            every record carrying one of the names :411
//   samtools fastq -0 r0 -1 r1 -2 r2                  primary records as FASTQ, split on the
//                                                     READ1/READ2 flags, "/1" "/2" appended  :417
//   v_map_type = size(r0) > size(r1) ? single : paired                                      :420-425
//   header edits of the screen loops ("/1" resp. "/2": first occurrence removed)            :469-472, :549-552
//   R1 joined to R2 by id, std::map order                                                   :556-603
//
// The BAM is scanned twice (names, then records); BGZF blocks are independent deflate streams and
// are inflated on all host threads, a chunk of blocks at a time; the inflated chunks of the first
// scan are kept for the second when they fit a memory budget.  samtools itself is not in the
// sandbox: the record selection restates its documented behaviour (overlap = [pos, end) against
// the half-open BED interval, end from the CIGAR's reference-consuming operations, pos + 1 when
// there are none; `fastq` drops secondary / supplementary records, reverse-complements records
// flagged reverse, writes quality 1 ('"') when the BAM holds none).
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "hlm_bgzf.h"
#include "hlm_db.h"
#include "hlm_files.h"

void hlm_ctx_set_error(hlm_ctx* ctx, const std::string& msg);
const hlm_params* hlm_ctx_params(const hlm_ctx* ctx);

struct hlm_reads {
  MateBatch m1, m2;  // paired: R1 / R2 joined by id, id order; single: m1 = R0
  int paired = 1;
};

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline uint32_t le32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t lei32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t le16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct Interval { int32_t beg, end; };

// BED intervals per chromosome name, merged and sorted
bool read_bed(const std::string& path, std::unordered_map<std::string, std::vector<Interval>>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char line[4096];
  while (fgets(line, sizeof line, f)) {
    if (line[0] == '#' || !strncmp(line, "track", 5) || !strncmp(line, "browser", 7)) continue;
    char chrom[1024];
    long long b, e;
    int k = sscanf(line, "%1023s %lld %lld", chrom, &b, &e);
    if (k == 2) { e = b; b = b - 1; }  // "chrom pos" position list, 1-based
    else if (k != 3) continue;
    if (e <= b) continue;
    out[chrom].push_back({(int32_t)std::max(0ll, b), (int32_t)std::min(e, 0x7fffffffll)});
  }
  fclose(f);
  for (auto& kv : out) {
    auto& v = kv.second;
    std::sort(v.begin(), v.end(), [](const Interval& a, const Interval& b) { return a.beg < b.beg; });
    size_t w = 0;
    for (size_t i = 1; i < v.size(); i++) {
      if (v[i].beg <= v[w].end) v[w].end = std::max(v[w].end, v[i].end);
      else v[++w] = v[i];
    }
    if (!v.empty()) v.resize(w + 1);
  }
  return true;
}

struct BamHeader {
  std::vector<std::string> ref_names;
  std::vector<std::vector<Interval>> regions;  // per refID
  bool parsed = false;
};

// BAM record field offsets after the 4-byte block_size
enum { R_REFID = 0, R_POS = 4, R_LNAME = 8, R_NCIG = 12, R_FLAG = 14, R_LSEQ = 16, R_FIXED = 32 };

struct Rec {
  const uint8_t* p;  // first byte after block_size
  uint32_t size;
  int32_t ref() const { return lei32(p + R_REFID); }
  int32_t pos() const { return lei32(p + R_POS); }
  uint32_t l_name() const { return p[R_LNAME]; }
  uint32_t n_cigar() const { return le16(p + R_NCIG); }
  uint32_t flag() const { return le16(p + R_FLAG); }
  uint32_t l_seq() const { return le32(p + R_LSEQ); }
  const char* name() const { return (const char*)p + R_FIXED; }
  size_t name_len() const { uint32_t l = l_name(); return l ? l - 1 : 0; }  // stored with its NUL
  const uint8_t* cigar() const { return p + R_FIXED + l_name(); }
  const uint8_t* seq() const { return cigar() + 4 * (size_t)n_cigar(); }
  const uint8_t* qual() const { return seq() + (l_seq() + 1) / 2; }
  bool sane() const {
    return size >= R_FIXED && (size_t)R_FIXED + l_name() + 4 * (size_t)n_cigar() + (l_seq() + 1) / 2 + l_seq() <= size;
  }
  int32_t end() const {  // bam_endpos(): a record flagged unmapped covers one base whatever its CIGAR
    if (flag() & 4u) return pos() + 1;
    int32_t len = 0;
    const uint8_t* c = cigar();
    for (uint32_t i = 0; i < n_cigar(); i++) {
      uint32_t v = le32(c + 4 * i), op = v & 15;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) len += (int32_t)(v >> 4);  // M D N = X
    }
    return pos() + (len ? len : 1);
  }
};

bool overlaps(const std::vector<Interval>& iv, int32_t beg, int32_t end) {
  // first interval with iv.end > beg
  size_t lo = 0, hi = iv.size();
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (iv[mid].end > beg) hi = mid;
    else lo = mid + 1;
  }
  return lo < iv.size() && iv[lo].beg < end;
}

uint64_t hash_name(const char* s, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull ^ (n * 0x9E3779B97F4A7C15ull);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, s + i, 8);
    h = (h ^ w) * 0x9FB21C651E98DF25ull;
    h ^= h >> 29;
  }
  uint64_t w = 0;
  if (i < n) memcpy(&w, s + i, n - i);
  h = (h ^ w) * 0x9FB21C651E98DF25ull;
  h ^= h >> 32;
  return h | 1;  // 0 marks an empty slot
}

// the BAM scanner: inflates kChunkBlocks BGZF blocks at a time on all threads (hlm_bgzf.h), walks
// the records and hands (records of the chunk) to `fn(worker, first, last)` on all threads
class BamScanner {
 public:
  BamScanner(const uint8_t* data, size_t size, int threads) : d_(data), n_(size), threads_(threads) {}
  std::string err;
  int64_t n_records = 0, bytes_inflated = 0;
  bool verify_crc = true;  // the second scan of the same mapping skips the check

  // Inflated chunks kept from the first scan so that the second does not inflate again; given
  // up (and emptied) as soon as it would exceed its budget.
  struct Cache {
    struct Chunk {
      std::unique_ptr<uint8_t[]> buf;
      size_t begin, end;  // whole records in buf[begin, end)
    };
    std::vector<Chunk> chunks;
    size_t bytes = 0, budget = 0;
    bool valid = false;
  };

  template <class Fn>
  bool scan(BamHeader& hdr, Fn&& fn, Cache* keep = nullptr) {
    HlmBgzf bg(d_, n_, threads_);
    bg.verify_crc = verify_crc;
    const size_t chunk_cap = (size_t)kChunkBlocks * 65536;
    size_t cap = chunk_cap + (1u << 20);
    std::unique_ptr<uint8_t[]> buf(new uint8_t[cap]);
    bool caching = keep != nullptr && keep->budget > 0;
    if (keep) keep->valid = false;
    size_t carry = 0;
    bool first = true;
    std::vector<Rec> recs;
    while (!bg.at_end()) {
      // ---- inflate the next chunk of blocks behind the bytes carried over from the previous one
      if (cap < carry + chunk_cap) {
        size_t ncap = carry + chunk_cap;
        std::unique_ptr<uint8_t[]> nb(new uint8_t[ncap]);
        memcpy(nb.get(), buf.get(), carry);
        buf = std::move(nb);
        cap = ncap;
      }
      long got = bg.fill(buf.get() + carry, cap - carry, kChunkBlocks);
      if (got < 0) return fail(bg.err);
      size_t out = carry + (size_t)got;
      bytes_inflated += got;
      // ---- header (first chunk), then the record boundaries
      size_t p = 0;
      if (first) {
        first = false;
        if (!parse_header(buf.get(), out, p, hdr)) return false;
        on_header();
      }
      size_t begin = p;
      if (!walk(buf.get(), p, out, recs)) return false;
      process(recs, fn);
      carry = out - p;
      if (caching && keep->bytes + cap > keep->budget) {  // does not fit: scan twice after all
        keep->chunks.clear();
        keep->bytes = 0;
        caching = false;
      }
      if (caching) {
        // the chunk stays where it is; the unfinished record moves to the front of a new buffer
        std::unique_ptr<uint8_t[]> nb(new uint8_t[cap]);
        memcpy(nb.get(), buf.get() + p, carry);
        keep->chunks.push_back({std::move(buf), begin, p});
        keep->bytes += cap;
        buf = std::move(nb);
      } else {
        memmove(buf.get(), buf.get() + p, carry);
      }
    }
    if (first) return fail("empty BAM file");
    if (carry) return fail("truncated BAM record at end of file");
    if (keep) keep->valid = caching;
    return true;
  }
  // the second scan over chunks kept by the first
  template <class Fn>
  bool scan(const Cache& kept, Fn&& fn) {
    std::vector<Rec> recs;
    for (const Cache::Chunk& c : kept.chunks) {
      size_t p = c.begin;
      if (!walk(c.buf.get(), p, c.end, recs) || p != c.end) return fail("malformed BAM record");
      process(recs, fn);
    }
    return true;
  }
  // both run on the scanning thread: after the header has been parsed; after every chunk (merge
  // of the per-worker results)
  std::function<void()> on_header = [] {};
  std::function<void()> on_chunk = [] {};
  int threads() const { return threads_; }

 private:
  enum { kChunkBlocks = 2048 };  // <= 128 MiB inflated per chunk
  bool fail(const std::string& m) {
    err = m;
    return false;
  }
  // record boundaries of buf[p, out): p ends behind the last whole record
  bool walk(const uint8_t* buf, size_t& p, size_t out, std::vector<Rec>& recs) {
    recs.clear();
    while (p + 4 <= out) {
      uint32_t bs = le32(buf + p);
      if (bs < R_FIXED) return fail("malformed BAM record");
      if (p + 4 + (size_t)bs > out) break;
      Rec r{buf + p + 4, bs};
      if (!r.sane()) return fail("malformed BAM record");
      recs.push_back(r);
      p += 4 + (size_t)bs;
    }
    return true;
  }
  template <class Fn>
  void process(const std::vector<Rec>& recs, Fn& fn) {
    n_records += (int64_t)recs.size();
    size_t nr = recs.size();
    run_parallel([&](int w) { fn(w, recs.data() + nr * w / threads_, recs.data() + nr * (w + 1) / threads_); });
    on_chunk();
  }
  template <class F>
  void run_parallel(F&& f) {
    std::vector<std::thread> th;
    for (int w = 1; w < threads_; w++) th.emplace_back([&f, w] { f(w); });
    f(0);
    for (auto& t : th) t.join();
  }
  bool parse_header(const uint8_t* b, size_t n, size_t& p, BamHeader& hdr) {
    if (n < 12 || memcmp(b, "BAM\1", 4) != 0) return fail("not a BAM file (magic)");
    uint32_t l_text = le32(b + 4);
    p = 8 + (size_t)l_text;
    if (p + 4 > n) return fail("BAM header does not fit the first chunk");
    uint32_t n_ref = le32(b + p);
    p += 4;
    std::vector<std::string> names;
    for (uint32_t i = 0; i < n_ref; i++) {
      if (p + 4 > n) return fail("BAM header does not fit the first chunk");
      uint32_t l = le32(b + p);
      if (p + 4 + (size_t)l + 4 > n) return fail("BAM header does not fit the first chunk");
      names.emplace_back((const char*)b + p + 4, l ? l - 1 : 0);
      p += 4 + (size_t)l + 4;
    }
    if (!hdr.parsed) {
      hdr.ref_names = names;
      hdr.parsed = true;
    }
    return true;
  }
  const uint8_t* d_;
  size_t n_;
  int threads_;
};

const char kSeqCode[] = "=ACMGRSVTWYHKDBN";
// complement of a 4-bit base code = its bits reversed (A1<->T8, C2<->G4, M3<->K12, ...)
const uint8_t kCompCode[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};

void erase_first(std::string& s, const char* pat) {  // string::find + erase, preselect.cpp:469-472
  size_t i = s.find(pat);
  if (i != std::string::npos) s.erase(i, strlen(pat));
}

// appends one record as `samtools fastq` would print it, with the header edit of the reference's
// screen loops applied; returns the bytes of the FASTQ record samtools would have written
size_t emit_fastq(const Rec& r, int which, MateBatch& dst, std::string& h, std::vector<uint8_t>& s,
                  std::vector<uint8_t>& q) {
  uint32_t l = r.l_seq();
  bool rev = (r.flag() & 0x10) != 0;
  const uint8_t* sq = r.seq();
  const uint8_t* ql = r.qual();
  s.resize(l);
  q.resize(l);
  for (uint32_t i = 0; i < l; i++) {
    uint32_t j = rev ? l - 1 - i : i;
    uint8_t code = (sq[j >> 1] >> ((~j & 1) << 2)) & 15;
    s[i] = (uint8_t)kSeqCode[rev ? kCompCode[code] : code];
  }
  if (l && ql[0] == 0xff) std::fill(q.begin(), q.end(), (uint8_t)(33 + 1));  // fastq -v 1
  else
    for (uint32_t i = 0; i < l; i++) q[i] = (uint8_t)(33 + ql[rev ? l - 1 - i : i]);
  h.assign("@");
  h.append(r.name(), r.name_len());
  if (which == 1) h += "/1";
  if (which == 2) h += "/2";
  size_t bytes = h.size() + 1 + 2 * ((size_t)l + 1) + 2;
  erase_first(h, which == 2 ? "/2" : "/1");
  dst.push(h.data(), h.size(), s.data(), q.data(), l);
  return bytes;
}

void set_err(char* err, size_t errlen, const std::string& m) {
  if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
}

// reads held in memory, served in slices
class MemSource : public PairSource {
 public:
  MemSource(const hlm_reads* r, int64_t batch) : r_(r), plan_(hlm_batch_plan(r->m1.n, batch)) {}
  bool next(MateBatch*& b1, MateBatch*& b2, int&, std::string&) override {
    if (k_ + 1 >= plan_.size()) return false;
    int64_t lo = plan_[k_], hi = plan_[k_ + 1];
    k_++;
    b1 = slice(r_->m1, lo, hi);
    b2 = r_->paired ? slice(r_->m2, lo, hi) : nullptr;
    return true;
  }
  void recycle(MateBatch* b1, MateBatch* b2) override {
    if (b1) free_.push_back(b1);
    if (b2) free_.push_back(b2);
  }

 private:
  MateBatch* slice(const MateBatch& m, int64_t lo, int64_t hi) {
    MateBatch* b;
    if (!free_.empty()) {
      b = free_.back();
      free_.pop_back();
    } else {
      pool_.emplace_back(new MateBatch());
      b = pool_.back().get();
    }
    b->clear();
    b->bases.assign(m.bases.begin() + m.offs[lo], m.bases.begin() + m.offs[hi]);
    b->quals.assign(m.quals.begin() + m.offs[lo], m.quals.begin() + m.offs[hi]);
    b->hdr.assign(m.hdr.begin() + m.hoff[lo], m.hdr.begin() + m.hoff[hi]);
    for (int64_t i = lo; i < hi; i++) {
      b->offs.push_back(m.offs[i + 1] - m.offs[lo]);
      b->hoff.push_back(m.hoff[i + 1] - m.hoff[lo]);
    }
    b->n = hi - lo;
    return b;
  }
  const hlm_reads* r_;
  std::vector<int64_t> plan_;  // first record of every batch
  size_t k_ = 0;
  std::vector<MateBatch*> free_;
  std::vector<std::unique_ptr<MateBatch>> pool_;
};

struct IdRef {
  const char* s;
  uint32_t n;
  int64_t idx;
};
// idgen of a FASTQ header line: first space-separated token minus '@' (preselect.cpp:498-500)
IdRef idgen_of(const MateBatch& m, int64_t i) {
  const char* h = m.hdr.data() + m.hoff[i];
  size_t hl = m.hoff[i + 1] - m.hoff[i];
  const char* sp = (const char*)memchr(h, ' ', hl);
  size_t tok = sp ? (size_t)(sp - h) : hl;
  return {h + (tok ? 1 : 0), (uint32_t)(tok ? tok - 1 : 0), i};
}
bool id_less(const IdRef& a, const IdRef& b) {
  int c = memcmp(a.s, b.s, std::min(a.n, b.n));
  return c < 0 || (c == 0 && a.n < b.n);
}
bool id_eq(const IdRef& a, const IdRef& b) { return a.n == b.n && memcmp(a.s, b.s, a.n) == 0; }

// sorted by id, one entry per id: the last record in file order (std::map assignment semantics)
std::vector<IdRef> sorted_unique_ids(const MateBatch& m) {
  std::vector<IdRef> v((size_t)m.n);
  for (int64_t i = 0; i < m.n; i++) v[(size_t)i] = idgen_of(m, i);
  std::stable_sort(v.begin(), v.end(), id_less);
  size_t w = 0;
  for (size_t i = 0; i < v.size(); i++) {
    if (i + 1 < v.size() && id_eq(v[i], v[i + 1])) continue;
    v[w++] = v[i];
  }
  v.resize(w);
  return v;
}

void append_read(MateBatch& dst, const MateBatch& src, int64_t i) {
  dst.push(src.hdr.data() + src.hoff[i], src.hoff[i + 1] - src.hoff[i], src.bases.data() + src.offs[i],
           src.quals.data() + src.offs[i], src.offs[i + 1] - src.offs[i]);
}

}  // namespace

static hlm_reads* bam_extract(const hlm_db* db, const char* bam_path, const char* bed_path,
                              int skip_unmapped, int threads, hlm_bam_stats* stats, char* err,
                              size_t errlen) {
  if (!db || !bam_path) {
    set_err(err, errlen, "hlm_bam_extract: db and bam_path are required");
    return nullptr;
  }
  double t_start = now_s();
  if (threads <= 0) threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::string bed = bed_path ? bed_path : db->dir + (db->rna ? "/bed/select_rna.bed" : "/bed/select_dna.bed");
  std::unordered_map<std::string, std::vector<Interval>> bed_iv;
  if (!read_bed(bed, bed_iv)) {
    set_err(err, errlen, "cannot read " + bed);
    return nullptr;
  }
  HlmMapped file;
  if (!file.open(bam_path)) {
    set_err(err, errlen, std::string("cannot open ") + bam_path);
    return nullptr;
  }

  // the inflated BAM is kept between the two scans when it fits a quarter of the machine's memory,
  // at most 64 GB (HLM_BAM_CACHE_GB overrides; 0 = always scan twice)
  BamScanner::Cache cache;
  {
    double gb = std::min(64.0, 0.25 * (double)sysconf(_SC_PHYS_PAGES) * (double)sysconf(_SC_PAGE_SIZE) / 1e9);
    if (const char* e = getenv("HLM_BAM_CACHE_GB")) gb = atof(e);
    cache.budget = gb > 0 ? (size_t)(gb * 1e9) : 0;
  }
  // ---- pass 1: names of the records in the BED regions and of the unmapped records
  BamHeader hdr;
  std::unordered_set<std::string> names;
  int64_t n_region = 0, n_unmapped = 0;
  double t1 = now_s();
  {
    BamScanner sc(file.p, file.n, threads);
    struct W {
      std::vector<std::string> names;
      int64_t n_region = 0, n_unmapped = 0;
    };
    std::vector<W> ws((size_t)threads);
    sc.on_header = [&] {  // -L matches the BED chromosome column against @SQ names literally
      hdr.regions.assign(hdr.ref_names.size(), {});
      for (size_t i = 0; i < hdr.ref_names.size(); i++) {
        auto it = bed_iv.find(hdr.ref_names[i]);
        if (it != bed_iv.end()) hdr.regions[i] = it->second;
      }
    };
    sc.on_chunk = [&] {
      for (W& w : ws) {
        for (std::string& s : w.names) names.insert(std::move(s));
        w.names.clear();
      }
    };
    bool ok = sc.scan(hdr, [&](int w, const Rec* a, const Rec* b) {
      W& o = ws[(size_t)w];
      for (const Rec* r = a; r < b; r++) {
        bool hit = false;
        int32_t ref = r->ref();
        if (ref >= 0 && (size_t)ref < hdr.regions.size() && !hdr.regions[(size_t)ref].empty()) {
          const auto& iv = hdr.regions[(size_t)ref];
          int32_t pos = r->pos();
          if (pos < iv.back().end && overlaps(iv, pos, r->end())) {
            hit = true;
            o.n_region++;
          }
        }
        if (!skip_unmapped && (r->flag() & 4)) {
          hit = true;
          o.n_unmapped++;
        }
        if (hit) o.names.emplace_back(r->name(), r->name_len());
      }
    }, &cache);
    if (!ok) {
      set_err(err, errlen, std::string(bam_path) + ": " + sc.err);
      return nullptr;
    }
    for (W& w : ws) {
      n_region += w.n_region;
      n_unmapped += w.n_unmapped;
    }
    if (stats) {
      memset(stats, 0, sizeof *stats);
      stats->n_records = sc.n_records;
      stats->bytes_compressed = (int64_t)file.n;
      stats->bytes_inflated = sc.bytes_inflated;
    }
  }
  double t2 = now_s();
  if (n_region == 0 || names.empty()) {  // empty hg38.region.sam / read_names.txt (preselect.cpp:360-363, :405-408)
    set_err(err, errlen, "Extraction failed!");
    return nullptr;
  }

  // ---- pass 2: every primary record carrying one of the names, as FASTQ, split on READ1 / READ2
  size_t cap = 16;
  while (cap < names.size() * 2) cap <<= 1;
  std::vector<uint64_t> table(cap, 0);
  for (const std::string& s : names) {
    uint64_t h = hash_name(s.data(), s.size());
    size_t i = (size_t)(h >> 1) & (cap - 1);
    while (table[i] && table[i] != h) i = (i + 1) & (cap - 1);
    table[i] = h;
  }
  MateBatch sets[3];
  for (MateBatch& m : sets) m.clear();
  uint64_t fq_bytes[3] = {0, 0, 0};
  {
    BamScanner sc(file.p, file.n, threads);
    sc.verify_crc = false;
    struct W {
      MateBatch out[3];
      uint64_t bytes[3] = {0, 0, 0};
      std::string h, key;
      std::vector<uint8_t> s, q;
    };
    std::vector<W> ws((size_t)threads);
    for (W& w : ws)
      for (MateBatch& m : w.out) m.clear();
    sc.on_chunk = [&] {  // worker order = file order
      for (W& w : ws)
        for (int k = 0; k < 3; k++) {
          MateBatch& src = w.out[k];
          for (int64_t i = 0; i < src.n; i++) append_read(sets[k], src, i);
          fq_bytes[k] += w.bytes[k];
          w.bytes[k] = 0;
          src.clear();
        }
    };
    BamHeader again;
    auto emit = [&](int wi, const Rec* a, const Rec* b) {
      W& w = ws[(size_t)wi];
      for (const Rec* r = a; r < b; r++) {
        uint32_t flag = r->flag();
        if (flag & 0x900) continue;  // samtools fastq -F 0x900
        uint64_t h = hash_name(r->name(), r->name_len());
        size_t i = (size_t)(h >> 1) & (cap - 1);
        while (table[i] && table[i] != h) i = (i + 1) & (cap - 1);
        if (!table[i]) continue;
        w.key.assign(r->name(), r->name_len());
        if (!names.count(w.key)) continue;
        bool r1 = (flag & 0x40) != 0, r2 = (flag & 0x80) != 0;
        int which = (r1 && !r2) ? 1 : (r2 && !r1) ? 2 : 0;
        w.bytes[which] += emit_fastq(*r, which, w.out[which], w.h, w.s, w.q);
      }
    };
    bool ok = cache.valid ? sc.scan(cache, emit) : sc.scan(again, emit);
    if (!ok) {
      set_err(err, errlen, std::string(bam_path) + ": " + sc.err);
      return nullptr;
    }
  }
  double t3 = now_s();
  if (fq_bytes[0] == 0 && fq_bytes[1] == 0) {  // preselect.cpp:431-440
    set_err(err, errlen, "There is not enough data to be processed. No mapped reads retrieved from BAM.");
    return nullptr;
  }

  // ---- v_map_type and the R1 / R2 join (preselect.cpp:420-425, :556-603)
  std::unique_ptr<hlm_reads> out(new hlm_reads());
  out->m1.clear();
  out->m2.clear();
  out->paired = fq_bytes[0] > fq_bytes[1] ? 0 : 1;
  if (out->paired) {
    std::vector<IdRef> a = sorted_unique_ids(sets[1]), b = sorted_unique_ids(sets[2]);
    size_t j = 0;
    for (const IdRef& x : a) {
      while (j < b.size() && id_less(b[j], x)) j++;
      if (j < b.size() && id_eq(b[j], x)) {
        append_read(out->m1, sets[1], x.idx);
        append_read(out->m2, sets[2], b[j].idx);
      }
    }
  } else {
    for (const IdRef& x : sorted_unique_ids(sets[0])) append_read(out->m1, sets[0], x.idx);
  }
  if (stats) {
    stats->n_region = n_region;
    stats->n_unmapped = n_unmapped;
    stats->n_names = (int64_t)names.size();
    stats->n_r0 = sets[0].n;
    stats->n_r1 = sets[1].n;
    stats->n_r2 = sets[2].n;
    stats->n_out = out->m1.n;
    stats->paired = out->paired;
    stats->cached = cache.valid ? 1 : 0;
    stats->s_pass1 = t2 - t1;
    stats->s_pass2 = t3 - t2;
    stats->s_total = now_s() - t_start;
  }
  return out.release();
}

extern "C" hlm_reads* hlm_bam_extract(const hlm_db* db, const char* bam_path, const char* bed_path,
                                      int skip_unmapped, int threads, hlm_bam_stats* stats, char* err,
                                      size_t errlen) {
  try {  // no exception crosses the C ABI
    return bam_extract(db, bam_path, bed_path, skip_unmapped, threads, stats, err, errlen);
  } catch (const std::exception& e) {
    set_err(err, errlen, std::string("hlm_bam_extract: ") + e.what());
    return nullptr;
  }
}

extern "C" int hlm_reads_paired(const hlm_reads* r) { return r ? r->paired : 0; }
extern "C" int64_t hlm_reads_count(const hlm_reads* r) { return r ? r->m1.n : 0; }
extern "C" void hlm_reads_free(hlm_reads* r) { delete r; }

extern "C" int hlm_reads_write_fastq(const hlm_reads* r, const char* r1_path, const char* r2_path) {
  if (!r || !r1_path || (r->paired && !r2_path)) return HLM_E_ARG;
  const MateBatch* ms[2] = {&r->m1, &r->m2};
  const char* paths[2] = {r1_path, r2_path};
  for (int k = 0; k < (r->paired ? 2 : 1); k++) {
    FILE* f = fopen(paths[k], "wb");
    if (!f) return HLM_E_IO;
    const MateBatch& m = *ms[k];
    for (int64_t i = 0; i < m.n; i++) {
      fwrite(m.hdr.data() + m.hoff[i], 1, m.hoff[i + 1] - m.hoff[i], f);
      fputc('\n', f);
      fwrite(m.bases.data() + m.offs[i], 1, m.offs[i + 1] - m.offs[i], f);
      fputs("\n+\n", f);
      fwrite(m.quals.data() + m.offs[i], 1, m.offs[i + 1] - m.offs[i], f);
      fputc('\n', f);
    }
    if (fclose(f) != 0) return HLM_E_IO;
  }
  return HLM_OK;
}

extern "C" int hlm_map_reads(hlm_ctx* ctx, const hlm_reads* reads, const char* sample, const char* out_dir,
                             const hlm_file_opts* opts, hlm_file_stats* stats) {
  if (!ctx || !reads || !sample || !out_dir) return HLM_E_ARG;
  const hlm_params* P = hlm_ctx_params(ctx);
  if ((P->paired != 0) != (reads->paired != 0)) {
    hlm_ctx_set_error(ctx, reads->paired ? "the BAM holds paired reads: create the context with paired = 1"
                                         : "the BAM holds single-end reads: create the context with paired = 0");
    return HLM_E_ARG;
  }
  if (P->screen_variant != HLM_SCREEN_BAM) {
    hlm_ctx_set_error(ctx, "reads extracted from a BAM are screened by the BAM loop (preselect.cpp:484-506): "
                           "create the context with screen_variant = HLM_SCREEN_BAM");
    return HLM_E_ARG;
  }
  int64_t batch = opts && opts->struct_size >= 12 && opts->batch_pairs > 0 ? opts->batch_pairs : 1 << 19;
  try {
    MemSource src(reads, batch);
    return hlm_map_source(ctx, src, /*normalise_pair_headers=*/false, sample, out_dir, opts, stats);
  } catch (const std::exception& e) {
    hlm_ctx_set_error(ctx, std::string("hlm_map_reads: ") + e.what());
    return HLM_E_NOMEM;
  }
}

extern "C" const char* hlm_last_error(const hlm_ctx* ctx);
extern "C" int hlm_multi_map_reads(hlm_multi* mm, const hlm_reads* reads, const char* sample, const char* out_dir,
                                   const hlm_file_opts* opts, hlm_file_stats* stats) {
  if (!mm || !reads || !sample || !out_dir) return HLM_E_ARG;
  int n = 0;
  hlm_ctx* const* cs = hlm_multi_contexts(mm, &n);
  if (n <= 0) return HLM_E_ARG;
  const hlm_params* P = hlm_ctx_params(cs[0]);
  if ((P->paired != 0) != (reads->paired != 0) || P->screen_variant != HLM_SCREEN_BAM) {
    hlm_multi_set_error(mm, "reads extracted from a BAM need contexts with the BAM's `paired` and "
                            "screen_variant = HLM_SCREEN_BAM");
    return HLM_E_ARG;
  }
  int64_t batch = opts && opts->struct_size >= 12 && opts->batch_pairs > 0 ? opts->batch_pairs : 1 << 19;
  int rc;
  try {
    MemSource src(reads, batch);
    rc = hlm_map_source_multi(cs, n, src, /*normalise_pair_headers=*/false, sample, out_dir, opts, stats);
  } catch (const std::exception& e) {
    hlm_multi_set_error(mm, std::string("hlm_multi_map_reads: ") + e.what());
    return HLM_E_NOMEM;
  }
  if (rc) hlm_multi_set_error(mm, hlm_last_error(cs[0]));
  return rc;
}
