// hlm_db.h — host-side database: reference layout parser + packed index builder.
#ifndef HLM_DB_H
#define HLM_DB_H

#include <mutex>
#include <string>
#include <vector>

#include "hlm_common.h"

struct hlm_db {
  int rna = 0;
  std::string dir;  // database directory as given (bed/select_*.bed is read from it on demand)
  int n_loci = 0;
  int v_size = 50;  // src/main.cpp:79; overridden by SIZE: (map_dna.cpp:351-354)
  std::vector<std::string> names;  // n_loci + "others"
  std::vector<int> gene_opt_start, gene_opt_end;  // GENE: fields 5, 6 (map_dna.cpp:331-332)

  // k-mer screen index
  std::vector<uint64_t> kt;
  std::vector<uint32_t> mask_tab;
  std::vector<uint64_t> kb;  // Bloom front of kt (hlm_kb_probe)
  int kt_bits = 0;
  int64_t n_kmers = 0;
  int64_t n_kmers_skipped = 0;  // 20-char motif lines holding a non-ACGT byte (kept in odd_*)
  std::vector<uint8_t> odd_keys;    // those lines, 20 raw bytes each, sorted
  std::vector<uint32_t> odd_masks;  // their membership masks

  // tiles + seed index
  std::vector<uint32_t> tiles;
  std::vector<uint32_t> tile_start;  // n_loci + 2
  std::vector<uint32_t> seed_off;
  std::vector<uint32_t> seed_ent;
  int64_t n_alleles = 0, n_bases = 0, n_tiles_raw = 0;

  // tile hierarchy: a tile within HLM_KMAX differing columns of an earlier tile of the same
  // locus and window index is stored as a "child" of that representative ("rep") tile
  std::vector<uint32_t> tile_parent;  // rep: own id
  std::vector<uint32_t> tile_diff;    // ndiff << 24 | col2 << 16 | col1 << 8 | col0 (0 for reps)
  std::vector<uint32_t> child_off;    // CSR over tiles: children of rep t = child_ent[child_off[t] .. child_off[t+1])
  std::vector<uint32_t> child_ent;    // child tile ids
  std::vector<uint32_t> child_diff;   // tile_diff of that child (parallel to child_ent)
  int64_t n_reps = 0;

  // per-allele view (hlm_score_alleles): record names of every target set in file order, and the
  // alleles that hold each tile (CSR over tiles; allele index inside its locus)
  std::vector<std::vector<std::string>> allele_names;  // n_loci + 1
  std::vector<uint32_t> tile_al_off, tile_al_ent;

  // device copies, one per CUDA device that has a context on this DB
  struct Dev {
    int device;
    HlmDevDB view;
    void* ptrs[14];
    int refs;
  };
  std::vector<Dev> devs;
  std::mutex devs_mutex;  // contexts may be created from several host threads
};

hlm_db* hlm_db_build(const char* dir, int rna, std::string& err);

#endif
