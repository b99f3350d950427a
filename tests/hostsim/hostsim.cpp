// hostsim.cpp — compiles the product's scoring arithmetic (hla_mapper_b200/csrc/hlm_core.cuh)
// and its host-side index builder (hlm_db.cpp) as plain C++ so the CPU test-suite can check
// them against the oracle without a GPU.  TEST-ONLY: never part of the product path.
#include <cstring>
#include <string>

#include "../../hla_mapper_b200/csrc/hlm_core.cuh"
#include "../../hla_mapper_b200/csrc/hlm_db.h"

static void pack_read(const uint8_t* codes, int n, uint32_t* rd) {
  memset(rd, 0, 18 * 4);
  for (int w = 0; w < 6; w++) rd[12 + w] = 0xFFFFFFFFu;
  for (int i = 0; i < n; i++) {
    int c = codes[i];
    if (c > 3) continue;
    rd[12 + (i >> 5)] &= ~(1u << (i & 31));
    if (c & 1) rd[i >> 5] |= 1u << (i & 31);
    if (c & 2) rd[6 + (i >> 5)] |= 1u << (i & 31);
  }
}
static void pack_tile(const uint8_t* cols, uint32_t* t) {
  memset(t, 0, 24 * 4);
  for (int c = 0; c < 256; c++) {
    int v = cols[c];
    if (v > 3) {
      t[16 + (c >> 5)] |= 1u << (c & 31);
      continue;
    }
    if (v & 1) t[c >> 5] |= 1u << (c & 31);
    if (v & 2) t[8 + (c >> 5)] |= 1u << (c & 31);
  }
}

extern "C" {
int hs_myers(const uint8_t* read, int n, const uint8_t* tile_cols, int off) {
  uint32_t rd[18], t[24];
  pack_read(read, n, rd);
  pack_tile(tile_cols, t);
  return hlm_myers_lb(rd, n, t, off);
}
void hs_dp(const uint8_t* read, int n, const uint8_t* tile_cols, int off, int* out4) {
  uint32_t rd[18], t[24];
  pack_read(read, n, rd);
  pack_tile(tile_cols, t);
  HlmAln a = hlm_dp_align(rd, n, t, off);
  out4[0] = a.S; out4[1] = a.cost; out4[2] = a.lead; out4[3] = a.trail;
}
int hs_seed(const uint8_t* read, int n, const uint8_t* tile_cols, int off, int q, uint32_t* code) {
  uint32_t rd[18], t[24];
  pack_read(read, n, rd);
  pack_tile(tile_cols, t);
  int ok = hlm_seed_code(rd, q * HLM_SEED, code) ? 1 : 0;
  return ok | (hlm_seed_match(rd, t, off, q * HLM_SEED) ? 2 : 0);
}
// index builder
void* hs_db_open(const char* dir, int rna, char* err, size_t errlen) {
  std::string e;
  hlm_db* db = hlm_db_build(dir, rna, e);
  if (!db && err) snprintf(err, errlen, "%s", e.c_str());
  return db;
}
void hs_db_close(void* p) { delete (hlm_db*)p; }
int64_t hs_db_stat(void* p, int what) {
  hlm_db* db = (hlm_db*)p;
  switch (what) {
    case 0: return db->n_loci;
    case 1: return db->n_kmers;
    case 2: return (int64_t)(db->tiles.size() / HLM_TILE_WORDS);
    case 3: return db->n_tiles_raw;
    case 4: return (int64_t)db->seed_ent.size();
    case 5: return db->n_alleles;
    case 6: return db->v_size;
    case 7: return db->n_kmers_skipped;
    case 8: return db->n_reps;
    case 9: return (int64_t)db->child_ent.size();
    default: return -1;
  }
}
uint32_t hs_db_kmer(void* p, const char* kmer20) {
  hlm_db* db = (hlm_db*)p;
  uint64_t key = 0;
  for (int x = 0; x < 20; x++) {
    int c = kmer20[x] == 'A' ? 0 : kmer20[x] == 'C' ? 1 : kmer20[x] == 'G' ? 2 : kmer20[x] == 'T' ? 3 : 4;
    if (c > 3) return 0;
    key |= (uint64_t)c << (2 * x);
  }
  uint64_t mask = (1ull << db->kt_bits) - 1, h = hlm_kt_hash(key, db->kt_bits);
  for (;;) {
    uint64_t s = db->kt[h];
    if (!(s & HLM_KT_OCC)) return 0;
    if ((s & HLM_KT_KEYMASK) == key) return db->mask_tab[(s >> 40) & 0xFFFF];
    h = (h + 1) & mask;
  }
}
// does the seed at read offset rp of a read placed at tile column off cover one of the child's
// differing columns?
static bool seed_covers(uint32_t diff, int off, int rp) {
  for (int k = 0; k < HLM_DIFF_N(diff); k++) {
    int x = HLM_DIFF_COL(diff, k), o = off + rp;
    if (x >= o && x < o + HLM_SEED) return true;
  }
  return false;
}
// all (tile, off) candidates of one locus for a read strand: seed-index hits (first matching
// seed owns the candidate) plus the children of every rep hit (same ownership rule; children
// whose differing columns all lie outside the DP footprint are identical to the rep: skipped)
int hs_db_candidates(void* p, const uint8_t* read, int n, int locus, uint32_t* out, int cap) {
  hlm_db* db = (hlm_db*)p;
  uint32_t rd[18];
  pack_read(read, n, rd);
  int cnt = 0, Q = hlm_seed_count(n, db->rna), st = hlm_seed_stride(n, db->rna);
  int n_index = 0;
  for (int q = 0; q < Q; q++) {
    uint32_t code;
    if (!hlm_seed_code(rd, q * st, &code)) continue;
    for (uint32_t i = db->seed_off[code]; i < db->seed_off[code + 1]; i++) {
      uint32_t e = db->seed_ent[i], tile = e & 0xFFFFFFu;
      int off = (int)(e >> 24) - q * st;
      if (off < HLM_OFF_MIN || off >= HLM_OFF_MIN + HLM_TILE_STRIDE) continue;
      if (tile < db->tile_start[locus] || tile >= db->tile_start[locus + 1]) continue;
      // one candidate per (tile, off) hit through the index: the first index entry that hits
      // it (children only list the seeds covering one of their differing columns)
      uint32_t diff = db->tile_diff[tile];
      bool first = true;
      for (int qq = 0; qq < q && first; qq++)
        if (hlm_seed_match(rd, &db->tiles[(size_t)tile * HLM_TILE_WORDS], off, qq * st) &&
            (HLM_DIFF_N(diff) == 0 || seed_covers(diff, off, qq * st)))
          first = false;
      if (!first) continue;
      if (cnt < cap) out[cnt] = (tile << 8) | (uint32_t)off;
      cnt++;
    }
  }
  n_index = cnt;
  for (int c = 0; c < n_index && c < cap; c++) {
    uint32_t tile = out[c] >> 8;
    int off = out[c] & 0xFF;
    for (uint32_t k = db->child_off[tile]; k < db->child_off[tile + 1]; k++) {
      uint32_t ch = db->child_ent[k], diff = db->child_diff[k];
      bool inside = false;
      for (int x = 0; x < HLM_DIFF_N(diff); x++) {
        int col = HLM_DIFF_COL(diff, x);
        if (col >= off - HLM_BAND && col < off + n + HLM_BAND) inside = true;
      }
      if (!inside) continue;
      // the expansion owns the child iff some seed matches and no matching seed covers a
      // differing column (such a seed is in the index, which then produced the candidate)
      bool any = false, indexed = false;
      for (int q = 0; q < Q; q++)
        if (hlm_seed_match(rd, &db->tiles[(size_t)ch * HLM_TILE_WORDS], off, q * st)) {
          any = true;
          if (seed_covers(diff, off, q * st)) indexed = true;
        }
      if (!any || indexed) continue;
      if (cnt < cap) out[cnt] = (ch << 8) | (uint32_t)off;
      cnt++;
    }
  }
  return cnt;
}
// every (tile, off) with at least one matching seed, by brute force over all tiles of the locus
int hs_db_candidates_brute(void* p, const uint8_t* read, int n, int locus, uint32_t* out, int cap) {
  hlm_db* db = (hlm_db*)p;
  uint32_t rd[18];
  pack_read(read, n, rd);
  int cnt = 0, Q = hlm_seed_count(n, db->rna), st = hlm_seed_stride(n, db->rna);
  for (uint32_t tile = db->tile_start[locus]; tile < db->tile_start[locus + 1]; tile++)
    for (int off = HLM_OFF_MIN; off < HLM_OFF_MIN + HLM_TILE_STRIDE; off++) {
      bool any = false;
      for (int q = 0; q < Q && !any; q++)
        any = hlm_seed_match(rd, &db->tiles[(size_t)tile * HLM_TILE_WORDS], off, q * st);
      if (!any) continue;
      if (cnt < cap) out[cnt] = (tile << 8) | (uint32_t)off;
      cnt++;
    }
  return cnt;
}
uint32_t hs_db_tile_info(void* p, uint32_t tile, int what) {
  hlm_db* db = (hlm_db*)p;
  return what == 0 ? db->tile_parent[tile] : db->tile_diff[tile];
}
// 64-bit hash of the DP footprint (tile columns off - BAND .. off + n + BAND) of a candidate and its
// Myers bound: statistics of footprint duplicates among the candidates (analysis only)
uint64_t hs_db_footprint(void* p, uint32_t cand, int n) {
  hlm_db* db = (hlm_db*)p;
  const uint32_t* T = &db->tiles[(size_t)(cand >> 8) * HLM_TILE_WORDS];
  int off = cand & 0xFF;
  uint64_t h = 1469598103934665603ULL;
  for (int c = off - HLM_BAND; c < off + n + HLM_BAND; c++) {
    int v = 4;
    if (c >= 0 && c < HLM_TILE_COLS && !((T[16 + (c >> 5)] >> (c & 31)) & 1u))
      v = (int)((T[c >> 5] >> (c & 31)) & 1u) | (int)(((T[8 + (c >> 5)] >> (c & 31)) & 1u) << 1);
    h = (h ^ (uint64_t)v) * 1099511628211ULL;
  }
  return h;
}
// band minimum after every 8 rows (analysis of early termination): out[k] = bound after 8 (k + 1) rows
int hs_db_lb_profile(void* p, const uint8_t* read, int n, uint32_t cand, int* out) {
  hlm_db* db = (hlm_db*)p;
  uint32_t rd[18];
  pack_read(read, n, rd);
  int k = 0;
  for (int r = 8; r <= n; r += 8) {
    int plb = 0;
    hlm_myers_lb(rd, n, &db->tiles[(size_t)(cand >> 8) * HLM_TILE_WORDS], cand & 0xFF, r, &plb);
    out[k++] = plb;
  }
  return k;
}
int hs_db_lb(void* p, const uint8_t* read, int n, uint32_t cand) {
  hlm_db* db = (hlm_db*)p;
  uint32_t rd[18];
  pack_read(read, n, rd);
  return hlm_myers_lb(rd, n, &db->tiles[(size_t)(cand >> 8) * HLM_TILE_WORDS], cand & 0xFF);
}
int hs_db_eval(void* p, const uint8_t* read, int n, uint32_t cand, int* out5) {
  hlm_db* db = (hlm_db*)p;
  uint32_t rd[18];
  pack_read(read, n, rd);
  const uint32_t* T = &db->tiles[(size_t)(cand >> 8) * HLM_TILE_WORDS];
  int off = cand & 0xFF;
  out5[0] = hlm_myers_lb(rd, n, T, off);
  HlmAln a = hlm_dp_align(rd, n, T, off);
  out5[1] = a.S; out5[2] = a.cost; out5[3] = a.lead; out5[4] = a.trail;
  return 0;
}
}
