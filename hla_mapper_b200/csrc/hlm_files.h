// hlm_files.h — shared pieces of the file-level front ends (hlm_files.cpp: FASTQ, hlm_bam.cpp: BAM)
#ifndef HLM_FILES_H
#define HLM_FILES_H

#include <cstdint>
#include <string>
#include <vector>

#include "hlm_common.h"

// one mate (or the single-end file) of a batch of reads, SoA as hlm_batch wants it
struct MateBatch {
  std::vector<uint8_t> bases, quals;
  std::vector<uint64_t> offs;       // n + 1
  std::vector<char> hdr;            // header lines, no newline, concatenated
  std::vector<uint64_t> hoff;       // n + 1
  int64_t n = 0;
  bool bad = false;                 // a partial record at the end of the text (dropped, like getline loops do)
  bool io_error = false;            // the compressed stream is corrupt / truncated
  void clear() {
    bases.clear(); quals.clear(); hdr.clear();
    offs.assign(1, 0); hoff.assign(1, 0);
    n = 0;
    bad = false;
    io_error = false;
  }
  void push(const char* h, size_t hl, const uint8_t* s, const uint8_t* q, size_t l) {
    hdr.insert(hdr.end(), h, h + hl);
    hoff.push_back(hdr.size());
    bases.insert(bases.end(), s, s + l);
    quals.insert(quals.end(), q, q + l);
    offs.push_back(bases.size());
    n++;
  }
};

// record index at which every batch of an input of N records starts (+ N at the end): batches of
// B pairs, small ones at the start and at the end of a large run (hlm_files.cpp)
std::vector<int64_t> hlm_batch_plan(int64_t N, int64_t B);

// where the batches of the mapping loop come from
class PairSource {
 public:
  virtual ~PairSource() {}
  // false at end of input; rc != 0 with err set on malformed input
  virtual bool next(MateBatch*& b1, MateBatch*& b2, int& rc, std::string& err) = 0;
  virtual void recycle(MateBatch* b1, MateBatch* b2) = 0;
};

// the loop shared by hlm_map_fastq and hlm_map_reads.  normalise_pair_headers: the FASTQ path's
// removal of /1 and /2 from both header lines (src/preselect.cpp:664-665, :675-676); the BAM
// path's headers arrive already edited (src/preselect.cpp:469-472, :549-552).
int hlm_map_source(hlm_ctx* ctx, PairSource& src, bool normalise_pair_headers, const char* sample,
                   const char* out_dir, const hlm_file_opts* opts, hlm_file_stats* stats);
// the same over several contexts (one per GPU): batches are dealt to the devices as they arrive,
// results committed in input order - the files do not depend on the number of devices
int hlm_map_source_multi(hlm_ctx* const* ctxs, int n_ctx, PairSource& src, bool normalise_pair_headers,
                         const char* sample, const char* out_dir, const hlm_file_opts* opts,
                         hlm_file_stats* stats);
// the per-device contexts of an hlm_multi (hlm_api.cu)
hlm_ctx* const* hlm_multi_contexts(const hlm_multi* m, int* n);
void hlm_multi_set_error(hlm_multi* m, const std::string& msg);

#endif
