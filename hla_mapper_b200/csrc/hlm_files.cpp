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
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "hlm_db.h"

extern "C" const char* hlm_last_error(const hlm_ctx* ctx);
void hlm_ctx_set_error(hlm_ctx* ctx, const std::string& msg);
const hlm_db* hlm_ctx_db(const hlm_ctx* ctx);
const hlm_params* hlm_ctx_params(const hlm_ctx* ctx);

namespace {

struct MateBatch {
  std::vector<uint8_t> bases, quals;
  std::vector<uint64_t> offs;       // n + 1
  std::vector<char> hdr;            // header lines, no newline, concatenated
  std::vector<uint64_t> hoff;       // n + 1
  int64_t n = 0;
  bool bad = false;                 // truncated / malformed input
  void clear() {
    bases.clear(); quals.clear(); hdr.clear();
    offs.assign(1, 0); hoff.assign(1, 0);
    n = 0;
  }
};

// reads 4-line records; returns false at EOF
class FastqReader {
 public:
  explicit FastqReader(const char* path) {
    f_ = gzopen(path, "rb");
    if (f_) gzbuffer(f_, 1 << 22);
    buf_.resize(1 << 24);
  }
  ~FastqReader() {
    if (f_) gzclose(f_);
  }
  bool ok() const { return f_ != nullptr; }
  // appends up to max_pairs records to b; returns number appended
  int64_t read(MateBatch& b, int64_t max_pairs) {
    int64_t got = 0;
    while (got < max_pairs) {
      const char* l[4];
      size_t n[4];
      bool eof = false;
      for (int k = 0; k < 4; k++)
        if (!line(l[k], n[k])) {
          if (k != 0) b.bad = true;  // the reference would read garbage here; we stop
          eof = true;
          break;
        }
      if (eof) break;
      b.hdr.insert(b.hdr.end(), l[0], l[0] + n[0]);
      b.hoff.push_back(b.hdr.size());
      b.bases.insert(b.bases.end(), l[1], l[1] + n[1]);
      // the quality line is cut / padded to the sequence length only for the device arrays;
      // hla-mapper itself never checks that they agree
      size_t q = std::min(n[1], n[3]);
      b.quals.insert(b.quals.end(), l[3], l[3] + q);
      b.quals.insert(b.quals.end(), n[1] - q, (uint8_t)'!');
      b.offs.push_back(b.bases.size());
      got++;
    }
    b.n += got;
    return got;
  }

 private:
  // next line (without '\n'); pointers stay valid until the 4th next call at least because a
  // refill moves the unread tail to the front only when fewer than 4 lines can be pending:
  // we simply copy the four lines of a record into a small side buffer
  bool line(const char*& p, size_t& n) {
    for (;;) {
      char* nl = (char*)memchr(buf_.data() + pos_, '\n', have_ - pos_);
      if (nl) {
        size_t len = nl - (buf_.data() + pos_);
        keep(buf_.data() + pos_, len, p);
        n = len;
        pos_ += len + 1;
        return true;
      }
      if (eof_) {
        if (pos_ < have_) {  // last line without newline
          size_t len = have_ - pos_;
          keep(buf_.data() + pos_, len, p);
          n = len;
          pos_ = have_;
          return true;
        }
        return false;
      }
      // refill: move the tail to the front
      size_t tail = have_ - pos_;
      memmove(buf_.data(), buf_.data() + pos_, tail);
      pos_ = 0;
      have_ = tail;
      if (have_ == buf_.size()) buf_.resize(buf_.size() * 2);
      int r = gzread(f_, buf_.data() + have_, (unsigned)std::min<size_t>(buf_.size() - have_, 1u << 30));
      if (r <= 0) eof_ = true;
      else have_ += (size_t)r;
    }
  }
  void keep(const char* s, size_t len, const char*& out) {
    std::vector<char>& v = side_[side_i_++ & 3];
    v.assign(s, s + len);
    out = v.data();
  }
  gzFile f_ = nullptr;
  std::vector<char> buf_;
  size_t pos_ = 0, have_ = 0;
  bool eof_ = false;
  std::vector<char> side_[4];
  unsigned side_i_ = 0;
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
      int64_t got = rd_.read(*b, batch_);
      std::lock_guard<std::mutex> g(m_);
      if (got > 0) q_.push(b);
      else free_.push_back(b);
      if (got < batch_) {
        done_ = true;
        cv_.notify_all();
        break;
      }
      cv_.notify_all();
    }
  }
  FastqReader rd_;
  int64_t batch_;
  std::thread th_;
  std::mutex m_;
  std::condition_variable cv_;
  std::queue<MateBatch*> q_;
  std::vector<MateBatch*> free_;
  std::vector<std::unique_ptr<MateBatch>> pool_;
  bool stop_ = false, done_ = false;
};

void erase_all(std::string& s, const char* pat) {  // boost::replace_all(s, pat, "")
  size_t n = strlen(pat), p = 0;
  while ((p = s.find(pat, p)) != std::string::npos) s.erase(p, n);
}

struct Kept {
  std::string idgen;      // sequence_size_r1 key
  std::string tok;        // first blank-separated token of the (normalised) header, with '@'
  std::string h1, h2;     // header lines as written to selected.R*.fastq
  std::string s1, q1, s2, q2;  // original (untrimmed) records
  uint16_t lo1, n1, lo2, n2;
  uint32_t locus_mask, target, visible;
  uint16_t os1, os2;
  std::vector<uint16_t> sc1, sc2;  // per target
};

struct Out {
  FILE* f = nullptr;
  std::string buf;
  bool open(const std::string& p) {
    f = fopen(p.c_str(), "wb");
    buf.reserve(1 << 22);
    return f != nullptr;
  }
  void flush() {
    if (f && !buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    buf.clear();
  }
  void add(const std::string& s) {
    buf += s;
    if (buf.size() > (1u << 22)) flush();
  }
  void add(const char* s, size_t n) {
    buf.append(s, n);
    if (buf.size() > (1u << 22)) flush();
  }
  void close() {
    flush();
    if (f) fclose(f);
    f = nullptr;
  }
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int hlm_map_fastq(hlm_ctx* ctx, const char* r1_path, const char* r2_path, const char* sample,
                             const char* out_dir, const hlm_file_opts* opts_in, hlm_file_stats* stats) {
  if (!ctx || !r1_path || !sample || !out_dir) return HLM_E_ARG;
  const hlm_db* db = hlm_ctx_db(ctx);
  const hlm_params* P = hlm_ctx_params(ctx);
  if (P->rna) {
    hlm_ctx_set_error(ctx, "hlm_map_fastq drives the dna path; rna needs the STAR block split as input");
    return HLM_E_ARG;
  }
  bool paired = P->paired != 0;
  if (paired != (r2_path != nullptr)) {
    hlm_ctx_set_error(ctx, "paired context needs r1 and r2; single-end context needs r2 == NULL");
    return HLM_E_ARG;
  }
  hlm_file_opts o;
  memset(&o, 0, sizeof o);
  o.struct_size = (int32_t)sizeof o;
  o.downsampling = 30;  // src/main.cpp:86
  o.batch_pairs = 1 << 20;
  o.write_selected = 1;
  if (opts_in) memcpy(&o, opts_in, std::min((size_t)opts_in->struct_size, sizeof o));
  if (o.batch_pairs <= 0) o.batch_pairs = 1 << 20;
  if (o.downsampling < 5) o.downsampling = 5;  // src/main.cpp:434
  int T = db->n_loci + 1;
  std::string out = std::string(out_dir);
  if (out.empty() || out.back() != '/') out += "/";  // src/main.cpp:438-448
  std::string pre = out + sample + "_";               // v_sample = sample + "_" (main.cpp:450-455)
  const char* tag = paired ? "R1" : "R0";

  double t_start = now_s(), t_gpu = 0, t_post = 0, t_wait = 0;
  BatchStream in1(r1_path, o.batch_pairs);
  std::unique_ptr<BatchStream> in2;
  if (paired) in2.reset(new BatchStream(r2_path, o.batch_pairs));
  if (!in1.ok() || (paired && !in2->ok())) {
    hlm_ctx_set_error(ctx, std::string("cannot open ") + (in1.ok() ? r2_path : r1_path));
    return HLM_E_IO;
  }
  Out sel1, sel2;
  if (o.write_selected) {
    if (!sel1.open(pre + "selected." + tag + ".fastq") || (paired && !sel2.open(pre + "selected.R2.fastq"))) {
      hlm_ctx_set_error(ctx, "cannot write into " + out);
      return HLM_E_IO;
    }
  }
  std::vector<Kept> kept;
  int64_t n_pairs = 0, n_selected = 0, sel_bases = 0;
  std::vector<uint8_t> flags;
  std::vector<uint16_t> lo1, len1, lo2, len2, s1, s2, minsum, os1, os2;
  std::vector<uint32_t> lmask, target, visible;
  std::vector<int32_t> so;
  int rc = 0;
  for (;;) {
    double t0 = now_s();
    MateBatch* b1 = in1.next();
    MateBatch* b2 = paired ? in2->next() : nullptr;
    t_wait += now_s() - t0;
    if (!b1 || (paired && !b2)) {
      if (paired && ((b1 != nullptr) != (b2 != nullptr))) {
        hlm_ctx_set_error(ctx, "r1 and r2 hold different numbers of reads");
        rc = HLM_E_IO;
      }
      break;
    }
    if (paired && b1->n != b2->n) {
      hlm_ctx_set_error(ctx, "r1 and r2 hold different numbers of reads");
      rc = HLM_E_IO;
      break;
    }
    int64_t n = b1->n;
    flags.resize(n); lo1.resize(n); len1.resize(n); lo2.resize(n); len2.resize(n);
    lmask.resize(n); target.resize(n); visible.resize(n); minsum.resize(n); os1.resize(n); os2.resize(n);
    s1.resize((size_t)n * T); s2.resize((size_t)n * T);
    hlm_batch hb;
    memset(&hb, 0, sizeof hb);
    hb.n_pairs = n;
    hb.bases1 = b1->bases.data(); hb.quals1 = b1->quals.data(); hb.offs1 = b1->offs.data();
    if (paired) {
      hb.bases2 = b2->bases.data(); hb.quals2 = b2->quals.data(); hb.offs2 = b2->offs.data();
    } else {
      // single end: sequence_size_r1 = length of the header line (src/preselect.cpp:1199)
      so.resize(n);
      for (int64_t i = 0; i < n; i++) so[i] = (int32_t)(b1->hoff[i + 1] - b1->hoff[i]);
      hb.size_override1 = so.data();
    }
    hlm_screen_out scr = {flags.data(), lo1.data(), len1.data(), lo2.data(), len2.data(), lmask.data()};
    hlm_score_out sc;
    memset(&sc, 0, sizeof sc);
    sc.s1 = s1.data();
    sc.s2 = s2.data();
    hlm_assign_out as = {target.data(), visible.data(), minsum.data(), os1.data(), os2.data()};
    t0 = now_s();
    rc = hlm_run(ctx, &hb, &scr, &sc, &as);
    t_gpu += now_s() - t0;
    if (rc) break;
    t0 = now_s();
    for (int64_t i = 0; i < n; i++) {
      if (!(flags[i] & HLM_F_SCREENED)) continue;
      n_selected++;
      std::string h1(b1->hdr.data() + b1->hoff[i], b1->hoff[i + 1] - b1->hoff[i]);
      std::string h2;
      if (paired) {  // src/preselect.cpp:664-665, :675-676
        h2.assign(b2->hdr.data() + b2->hoff[i], b2->hoff[i + 1] - b2->hoff[i]);
        erase_all(h1, "/1"); erase_all(h1, "/2");
        erase_all(h2, "/1"); erase_all(h2, "/2");
      }
      const char* q1raw = (const char*)b1->quals.data() + b1->offs[i];
      size_t L1 = b1->offs[i + 1] - b1->offs[i];
      sel_bases += (int64_t)L1;
      if (o.write_selected) {
        sel1.add(h1); sel1.add("\n", 1);
        sel1.add((const char*)b1->bases.data() + b1->offs[i], L1); sel1.add("\n+\n", 3);
        sel1.add(q1raw, L1); sel1.add("\n", 1);
        if (paired) {
          size_t L2 = b2->offs[i + 1] - b2->offs[i];
          sel2.add(h2); sel2.add("\n", 1);
          sel2.add((const char*)b2->bases.data() + b2->offs[i], L2); sel2.add("\n+\n", 3);
          sel2.add((const char*)b2->quals.data() + b2->offs[i], L2); sel2.add("\n", 1);
        }
      }
      if (!(flags[i] & HLM_F_KEPT)) continue;
      kept.emplace_back();
      Kept& k = kept.back();
      k.h1 = h1;
      k.h2 = h2;
      size_t sp = h1.find(' ');
      k.tok = h1.substr(0, sp);
      k.idgen = k.tok.size() ? k.tok.substr(1) : std::string();
      erase_all(k.idgen, "/1");  // src/preselect.cpp:1076-1078, :1170-1172
      erase_all(k.idgen, "/2");
      k.s1.assign((const char*)b1->bases.data() + b1->offs[i], L1);
      k.q1.assign(q1raw, L1);
      if (paired) {
        size_t L2 = b2->offs[i + 1] - b2->offs[i];
        k.s2.assign((const char*)b2->bases.data() + b2->offs[i], L2);
        k.q2.assign((const char*)b2->quals.data() + b2->offs[i], L2);
      }
      k.lo1 = lo1[i]; k.n1 = len1[i]; k.lo2 = lo2[i]; k.n2 = len2[i];
      k.locus_mask = lmask[i]; k.target = target[i]; k.visible = visible[i];
      k.os1 = os1[i]; k.os2 = os2[i];
      k.sc1.assign(s1.begin() + (size_t)i * T, s1.begin() + (size_t)(i + 1) * T);
      k.sc2.assign(s2.begin() + (size_t)i * T, s2.begin() + (size_t)(i + 1) * T);
    }
    n_pairs += n;
    t_post += now_s() - t0;
    in1.recycle(b1);
    if (paired) in2->recycle(b2);
  }
  sel1.close();
  sel2.close();
  if (rc) return rc;

  double t0 = now_s();
  // std::map<string, ...> order of the reference (byte-wise on the id)
  std::stable_sort(kept.begin(), kept.end(), [](const Kept& a, const Kept& b) { return a.idgen < b.idgen; });
  // the four groups of output files are independent: write them on separate host threads
  std::vector<std::thread> writers;
  // ---- selected.trim.R*.fastq (src/preselect.cpp:1121-1149, :1182-1203)
  writers.emplace_back([&] {
    Out t1, t2;
    t1.open(pre + "selected.trim." + tag + ".fastq");
    if (paired) t2.open(pre + "selected.trim.R2.fastq");
    std::string qa;
    for (const Kept& k : kept) {
      t1.add(k.h1); t1.add("\n", 1);
      t1.add(k.s1.data() + k.lo1, k.n1); t1.add("\n+\n", 3);
      qa.assign(k.n1, 'A'); t1.add(qa); t1.add("\n", 1);
      if (paired) {
        t2.add(k.h2); t2.add("\n", 1);
        t2.add(k.s2.data() + k.lo2, k.n2); t2.add("\n+\n", 3);
        qa.assign(k.n2, 'A'); t2.add(qa); t2.add("\n", 1);
      }
    }
    t1.close();
    t2.close();
  });
  // ---- <sample>_addressing_table.txt (src/map_dna.cpp:932-1051)
  int64_t n_assigned = 0;
  writers.emplace_back([&] {
    Out tb;
    tb.open(pre + "addressing_table.txt");
    std::string line = "Read";
    for (int g = 0; g < T; g++) line += "\t" + db->names[g];
    line += "\tTarget\n";
    tb.add(line);
    for (const Kept& k : kept) {
      line = k.idgen;
      for (int g = 0; g < T; g++) {
        line += "\t";
        if (!((k.visible >> g) & 1u)) {
          line += "-";
          continue;
        }
        int a = g == T - 1 ? k.os1 : k.sc1[g], b = g == T - 1 ? k.os2 : k.sc2[g];
        line += std::to_string(a);
        if (paired) line += "," + std::to_string(b);
      }
      std::string tg;
      for (int g = 0; g < T; g++)
        if ((k.target >> g) & 1u) tg += (tg.empty() ? "" : ",") + db->names[g];
      if (tg.empty()) tg = "-";
      else n_assigned++;
      line += "\t" + tg + "\n";
      tb.add(line);
    }
    tb.close();
  });
  // ---- per-gene FASTQ files (src/map_dna.cpp:1056-1351)
  int64_t read_mean_size = n_selected > 0 ? sel_bases / n_selected : 0;  // :1125-1133 (int division)
  int n_gene_threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
  for (int w = 0; w < n_gene_threads; w++)
    writers.emplace_back([&, w] {
  for (int g = w; g < db->n_loci; g += n_gene_threads) {
    const std::string& gene = db->names[g];
    Out r1, r2, c1, c2;
    r1.open(pre + gene + "_" + tag + ".fastq");
    c1.open(pre + gene + ".trim." + tag + ".fastq");
    if (paired) {
      r2.open(pre + gene + "_R2.fastq");
      c2.open(pre + gene + ".trim.R2.fastq");
    }
    int64_t size = (int64_t)db->gene_opt_end[g] - db->gene_opt_start[g];
    int64_t downS = read_mean_size > 0 ? ((int64_t)o.downsampling * size) / read_mean_size : 0;  // :1198-1200
    int64_t counter = 0;
    std::string qa;
    for (const Kept& k : kept) {
      if (!((k.locus_mask >> g) & 1u)) continue;   // was in sorted_<gene>_R1.fastq
      if (!((k.target >> g) & 1u)) continue;        // sequence_address[(gene, id)]
      // the lookup key is the header token minus '@' (:1225): it only equals the id when the
      // token carries no /1 /2 (always true for paired input, whose headers were normalised)
      if (k.tok.size() < 1 || k.tok.compare(1, std::string::npos, k.idgen) != 0) continue;
      r1.add(k.h1); r1.add("\n", 1); r1.add(k.s1); r1.add("\n+\n", 3); r1.add(k.q1); r1.add("\n", 1);
      if (paired) {
        r2.add(k.h2); r2.add("\n", 1); r2.add(k.s2); r2.add("\n+\n", 3); r2.add(k.q2); r2.add("\n", 1);
      }
      if (counter < downS) {
        c1.add(k.tok); c1.add(" 1:trim\n", 8);
        c1.add(k.s1.data() + k.lo1, k.n1); c1.add("\n+\n", 3);
        qa.assign(k.n1, 'A'); c1.add(qa); c1.add("\n", 1);
        if (paired) {
          c2.add(k.tok); c2.add(" 2:trim\n", 8);
          c2.add(k.s2.data() + k.lo2, k.n2); c2.add("\n+\n", 3);
          qa.assign(k.n2, 'A'); c2.add(qa); c2.add("\n", 1);
        }
      }
      counter++;
    }
    r1.close(); r2.close(); c1.close(); c2.close();
  }
    });
  for (auto& th : writers) th.join();
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
  }
  return HLM_OK;
}
