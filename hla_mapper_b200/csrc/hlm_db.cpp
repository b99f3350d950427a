// hlm_db.cpp — parses the hla-mapper database layout and builds the packed GPU indexes.
//
// Reads exactly the files the reference reads for the hot path:
//   db_{dna,rna}.info                        src/map_dna.cpp:299-365, src/map_rna.cpp:281-335
//   motif/<kind>/all/motif_all.txt           src/preselect.cpp:297-309
//   motif/<kind>/loci/<g>.txt                src/map_dna.cpp:565-581, src/map_rna.cpp:498-514
//   mapper/<kind>/<g>/<g>.fas                src/map_dna.cpp:735,  src/map_rna.cpp:833
//   others/<kind>/hla_gen_others.fasta       src/map_dna.cpp:824,  src/map_rna.cpp:858
#include "hlm_db.h"

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <unordered_map>
#include <thread>

namespace {

bool read_file(const std::string& path, std::vector<char>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  out.resize(got);
  return true;
}

inline int base_code(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

// motif lines -> (key, mask bit).  A line can only ever equal seq.substr(c,20) when it is
// exactly 20 bytes.  A 20-byte line holding a non-ACGT byte (N, lower case ...) is not
// representable in 2 bits: it goes to the small "odd" list of raw strings, which the screen
// consults only for read windows that themselves hold a non-ACGT byte (the reference compares
// raw strings, so such a line matches a read with the same bytes in place).
void load_motifs(const std::string& path, uint32_t bit,
                 std::vector<std::pair<uint64_t, uint32_t>>& out, int64_t& skipped,
                 std::vector<std::pair<std::string, uint32_t>>& odd) {
  std::vector<char> buf;
  if (!read_file(path, buf)) return;
  size_t n = buf.size();
  unsigned nt = n > (8u << 20) ? std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1;
  std::vector<std::vector<std::pair<uint64_t, uint32_t>>> part(nt);
  std::vector<int64_t> skip(nt, 0);
  std::vector<std::vector<std::pair<std::string, uint32_t>>> oddp(nt);
  auto work = [&](unsigned k) {
    size_t i = n * k / nt, end = n * (k + 1) / nt;
    if (k > 0) {  // start at the first line that begins inside this chunk
      while (i < n && buf[i - 1] != '\n') i++;
    }
    auto& o = part[k];
    while (i < end) {
      size_t e = i;
      while (e < n && buf[e] != '\n') e++;
      if (e - i == HLM_MOTIF) {
        uint64_t key = 0;
        bool ok = true;
        for (int x = 0; x < HLM_MOTIF; x++) {
          unsigned char c = (unsigned char)buf[i + x];
          int code = (c == 'A') ? 0 : (c == 'C') ? 1 : (c == 'G') ? 2 : (c == 'T') ? 3 : 4;
          if (code > 3) {
            ok = false;
            break;
          }
          key |= (uint64_t)code << (2 * x);
        }
        if (ok) o.emplace_back(key, bit);
        else {
          skip[k]++;
          oddp[k].emplace_back(std::string(&buf[i], HLM_MOTIF), bit);
        }
      }
      i = e + 1;
    }
  };
  std::vector<std::thread> th;
  for (unsigned k = 0; k < nt; k++) th.emplace_back(work, k);
  for (auto& t : th) t.join();
  for (unsigned k = 0; k < nt; k++) {
    out.insert(out.end(), part[k].begin(), part[k].end());
    skipped += skip[k];
    odd.insert(odd.end(), oddp[k].begin(), oddp[k].end());
  }
}

struct Planes {
  std::vector<uint32_t> p0, p1, nm;
  int len = 0;
};

// FASTA -> bit planes of each N-padded allele
void load_fasta(const std::string& path, std::vector<Planes>& out, int64_t& n_bases,
                std::vector<std::string>* names = nullptr) {
  std::vector<char> buf;
  if (!read_file(path, buf)) return;
  std::vector<uint8_t> codes;
  auto flush = [&](bool have) {
    if (!have) return;
    Planes P;
    P.len = (int)codes.size();
    size_t nw = (size_t)(HLM_PAD + P.len + HLM_PAD + 31) / 32 + 8;
    P.p0.assign(nw, 0);
    P.p1.assign(nw, 0);
    P.nm.assign(nw, 0xFFFFFFFFu);
    for (int x = 0; x < P.len; x++) {
      int c = codes[x];
      if (c > 3) continue;
      size_t col = (size_t)HLM_PAD + x;
      uint32_t b = 1u << (col & 31);
      P.nm[col >> 5] &= ~b;
      if (c & 1) P.p0[col >> 5] |= b;
      if (c & 2) P.p1[col >> 5] |= b;
    }
    n_bases += P.len;
    out.push_back(std::move(P));
    codes.clear();
  };
  size_t i = 0, n = buf.size();
  bool have = false;
  while (i < n) {
    size_t e = i;
    while (e < n && buf[e] != '\n') e++;
    size_t le = e;
    while (le > i && (buf[le - 1] == '\r')) le--;
    if (le > i) {
      if (buf[i] == '>') {
        flush(have);
        have = true;
        if (names) {  // the record's name: the header up to the first blank (what SAM calls RNAME)
          size_t x = i + 1;
          while (x < le && buf[x] != ' ' && buf[x] != '\t') x++;
          names->emplace_back(&buf[i + 1], x - (i + 1));
        }
      } else if (have) {
        for (size_t x = i; x < le; x++) codes.push_back((uint8_t)base_code((unsigned char)buf[x]));
      }
    }
    i = e + 1;
  }
  flush(have);
}

inline uint64_t mix64(uint64_t h, uint64_t v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  return h ^ (h >> 32);
}

}  // namespace

static double db_now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define DB_T(label)                                                                   \
  if (getenv("HLM_DB_TIMING")) {                                                      \
    fprintf(stderr, "[hlm_db] %-28s %.2fs\n", label, db_now() - t_stage);             \
    t_stage = db_now();                                                               \
  }

hlm_db* hlm_db_build(const char* dir, int rna, std::string& err) {
  double t_stage = db_now();
  std::string base(dir);
  const char* kind = rna ? "rna" : "dna";
  std::vector<char> info;
  std::string info_path = base + "/db_" + kind + ".info";
  if (!read_file(info_path, info)) {
    err = "Error accessing database " + info_path;
    return nullptr;
  }
  hlm_db* db = new hlm_db();
  db->rna = rna;
  db->dir = dir;
  bool version_ok = false;
  {
    size_t i = 0, n = info.size();
    while (i < n) {
      size_t e = i;
      while (e < n && info[e] != '\n') e++;
      std::string line(info.data() + i, e - i);
      if (line == "hla-mapper:4.3+") version_ok = true;
      else if (line.compare(0, 5, "GENE:") == 0) {
        size_t q = line.find(':', 5);
        std::string g = line.substr(5, q == std::string::npos ? std::string::npos : q - 5);
        g.erase(std::remove(g.begin(), g.end(), ' '), g.end());
        if (!g.empty() && db->n_loci >= HLM_MAX_LOCI) {
          // one uint32 membership / target mask per pair: a 32nd locus cannot be represented,
          // and dropping it silently would score its reads against the wrong targets
          err = "db_" + std::string(kind) + ".info lists more than " + std::to_string(HLM_MAX_LOCI) +
                " GENE: lines (limit of this build: loci + others must fit a 32-bit mask)";
          delete db;
          return nullptr;
        }
        if (!g.empty()) {
          db->names.push_back(g);
          db->n_loci++;
          // GENE:name:chr:pos:chrsize:optstart:optend:type
          std::vector<std::string> f;
          size_t a = 0;
          while (a <= line.size()) {
            size_t b = line.find(':', a);
            if (b == std::string::npos) b = line.size();
            f.push_back(line.substr(a, b - a));
            a = b + 1;
          }
          db->gene_opt_start.push_back(f.size() > 5 ? atoi(f[5].c_str()) : 0);
          db->gene_opt_end.push_back(f.size() > 6 ? atoi(f[6].c_str()) : 0);
        }
      } else if (line.compare(0, 5, "SIZE:") == 0) {
        db->v_size = atoi(line.c_str() + 5);
      }
      i = e + 1;
    }
  }
  if (!version_ok) {
    err = "The database is not compatible with this version of hla-mapper.";
    delete db;
    return nullptr;
  }
  db->names.push_back("others");

  // ---- k-mer screen index
  {
    std::vector<std::pair<uint64_t, uint32_t>> kv;
    std::vector<std::pair<std::string, uint32_t>> odd;
    load_motifs(base + "/motif/" + kind + "/all/motif_all.txt", 1u << 31, kv, db->n_kmers_skipped, odd);
    for (int g = 0; g < db->n_loci; g++)
      load_motifs(base + "/motif/" + kind + "/loci/" + db->names[g] + ".txt", 1u << g, kv,
                  db->n_kmers_skipped, odd);
    std::sort(odd.begin(), odd.end());
    for (size_t i = 0; i < odd.size();) {  // sorted raw strings, masks merged
      uint32_t m = 0;
      size_t j = i;
      while (j < odd.size() && odd[j].first == odd[i].first) m |= odd[j++].second;
      db->odd_keys.insert(db->odd_keys.end(), odd[i].first.begin(), odd[i].first.end());
      db->odd_masks.push_back(m);
      i = j;
    }
    std::sort(kv.begin(), kv.end());
    std::vector<std::pair<uint64_t, uint32_t>> uniq;
    for (size_t i = 0; i < kv.size();) {
      uint64_t k = kv[i].first;
      uint32_t m = 0;
      while (i < kv.size() && kv[i].first == k) m |= kv[i++].second;
      uniq.emplace_back(k, m);
    }
    db->n_kmers = (int64_t)uniq.size();
    int bits = 10;
    while ((1ull << bits) < 2 * uniq.size() + 2) bits++;
    db->kt_bits = bits;
    db->kt.assign(1ull << bits, 0);
    std::map<uint32_t, uint32_t> mask_id;
    uint64_t cap_mask = (1ull << bits) - 1;
    for (auto& e : uniq) {
      auto it = mask_id.find(e.second);
      if (it == mask_id.end()) {
        it = mask_id.emplace(e.second, (uint32_t)db->mask_tab.size()).first;
        db->mask_tab.push_back(e.second);
      }
      uint64_t h = hlm_kt_hash(e.first, bits);
      while (db->kt[h] & HLM_KT_OCC) h = (h + 1) & cap_mask;
      db->kt[h] = HLM_KT_OCC | ((uint64_t)it->second << 40) | e.first;
    }
    if (db->mask_tab.size() > 65535) {
      err = "too many distinct motif membership masks";
      delete db;
      return nullptr;
    }
    if (db->mask_tab.empty()) db->mask_tab.push_back(0);
    // Bloom front of the table (hlm_common.h): 4 bits per key in one 64-bit word, ~2 keys per
    // word, so that it stays L2-resident (4 bytes per key) while the exact table does not
    db->kb.assign(std::max<size_t>(64, (uniq.size() + 1) / 2), 0);
    for (auto& e : uniq) {
      uint32_t w;
      uint64_t m = hlm_kb_probe(e.first, (uint32_t)db->kb.size(), &w);
      db->kb[w] |= m;
    }
  }

  DB_T("k-mer table")
  // ---- tiles: per locus, distinct 256-column windows at stride 32 of the padded alleles.
  // A new window that differs from an earlier window of the same locus and window index in
  // <= HLM_KMAX columns becomes a "child" of that representative; everything else is a "rep".
  // loci are independent: build their tile sets on host threads, then concatenate
  struct LocusTiles {
    std::vector<uint32_t> tiles, parent, diff;
    std::vector<uint32_t> al_tiles;  // tile (local id) of every window of every allele, allele by allele
    std::vector<std::string> al_names;
    int64_t n_alleles = 0, n_bases = 0, n_raw = 0, n_reps = 0;
  };
  std::vector<LocusTiles> per(db->n_loci + 1);
  auto build_locus = [&](int g) {
    LocusTiles& L = per[g];
    std::vector<Planes> al;
    std::string path = g < db->n_loci
                           ? base + "/mapper/" + kind + "/" + db->names[g] + "/" + db->names[g] + ".fas"
                           : base + "/others/" + kind + "/hla_gen_others.fasta";
    load_fasta(path, al, L.n_bases, &L.al_names);
    L.n_alleles += (int64_t)al.size();
    size_t raw = 0;
    for (auto& P : al) raw += (size_t)(HLM_PAD + P.len) / 32 + 1;
    L.n_raw += (int64_t)raw;
    size_t cap = 1024;
    while (cap < raw * 2) cap <<= 1;
    std::vector<uint32_t> slot(cap, 0xFFFFFFFFu);  // tile id per hash slot
    // reps indexed by (window index, word-column x, content of that word-column): two windows
    // within KMAX differing columns share >= 8 - KMAX of their 8 word-columns
    std::unordered_map<uint64_t, std::vector<uint32_t>> rep_lsh;
    std::vector<uint32_t> seen;
    uint32_t w[HLM_TILE_WORDS];
    uint32_t al_id = 0;
    for (auto& P : al) {
      int jmax = (HLM_PAD + P.len) / 32;
      uint32_t this_al = al_id++;
      for (int j = 0; j <= jmax; j++) {
        uint64_t h = 0x1234567ull;
        for (int x = 0; x < 8; x++) {
          w[x] = P.p0[j + x];
          w[8 + x] = P.p1[j + x];
          w[16 + x] = P.nm[j + x];
        }
        for (int x = 0; x < HLM_TILE_WORDS; x += 2) h = mix64(h, (uint64_t)w[x] | ((uint64_t)w[x + 1] << 32));
        size_t s = h & (cap - 1);
        bool found = false;
        while (slot[s] != 0xFFFFFFFFu) {
          if (memcmp(&L.tiles[(size_t)slot[s] * HLM_TILE_WORDS], w, sizeof w) == 0) {
            found = true;
            break;
          }
          s = (s + 1) & (cap - 1);
        }
        L.al_tiles.push_back(this_al);
        L.al_tiles.push_back(found ? slot[s] : (uint32_t)(L.tiles.size() / HLM_TILE_WORDS));
        if (found) continue;
        uint32_t id = (uint32_t)(L.tiles.size() / HLM_TILE_WORDS);
        slot[s] = id;
        L.tiles.insert(L.tiles.end(), w, w + HLM_TILE_WORDS);
        // nearest representative at this window index
        uint32_t best_rep = id, best_diff = 0;
        int best_n = HLM_KMAX + 1;
        uint64_t keys[8];
        seen.clear();
        for (int x = 0; x < 8; x++) {
          uint64_t kx = mix64(mix64((uint64_t)j * 8 + x, (uint64_t)w[x] | ((uint64_t)w[8 + x] << 32)), w[16 + x]);
          keys[x] = kx;
          auto it = rep_lsh.find(kx);
          if (it != rep_lsh.end())
            for (uint32_t r : it->second) seen.push_back(r);
        }
        std::sort(seen.begin(), seen.end());
        seen.erase(std::unique(seen.begin(), seen.end()), seen.end());
        for (uint32_t r : seen) {
          const uint32_t* R = &L.tiles[(size_t)r * HLM_TILE_WORDS];
          int n = 0;
          uint32_t dmask[8];
          for (int x = 0; x < 8 && n < best_n; x++) {
            dmask[x] = (w[x] ^ R[x]) | (w[8 + x] ^ R[8 + x]) | (w[16 + x] ^ R[16 + x]);
            n += __builtin_popcount(dmask[x]);
          }
          if (n >= best_n) continue;
          uint32_t d = (uint32_t)n << 24;
          int k = 0;
          for (int x = 0; x < 8; x++)
            for (uint32_t m = dmask[x]; m; m &= m - 1) d |= (uint32_t)(32 * x + __builtin_ctz(m)) << (8 * k++);
          best_n = n;
          best_rep = r;
          best_diff = d;
          if (n == 1) break;
        }
        if (best_rep == id) {
          for (int x = 0; x < 8; x++) rep_lsh[keys[x]].push_back(id);
          L.n_reps++;
        }
        L.parent.push_back(best_rep);
        L.diff.push_back(best_diff);
      }
    }
  };
  {
    std::atomic<int> next(0);
    unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), db->n_loci + 1));
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; k++)
      th.emplace_back([&] {
        for (int g; (g = next.fetch_add(1)) <= db->n_loci;) build_locus(g);
      });
    for (auto& t : th) t.join();
  }
  db->tile_start.assign(db->n_loci + 2, 0);
  db->allele_names.resize(db->n_loci + 1);
  std::vector<std::pair<uint32_t, uint32_t>> tile_al;  // (global tile, allele index inside its locus)
  for (int g = 0; g <= db->n_loci; g++) {
    LocusTiles& L = per[g];
    uint32_t base_id = (uint32_t)(db->tiles.size() / HLM_TILE_WORDS);
    db->tile_start[g] = base_id;
    db->allele_names[g] = std::move(L.al_names);
    for (size_t k = 0; k + 1 < L.al_tiles.size(); k += 2) tile_al.emplace_back(L.al_tiles[k + 1] + base_id, L.al_tiles[k]);
    db->tiles.insert(db->tiles.end(), L.tiles.begin(), L.tiles.end());
    for (uint32_t pr : L.parent) db->tile_parent.push_back(pr + base_id);
    db->tile_diff.insert(db->tile_diff.end(), L.diff.begin(), L.diff.end());
    db->n_alleles += L.n_alleles;
    db->n_bases += L.n_bases;
    db->n_tiles_raw += L.n_raw;
    db->n_reps += L.n_reps;
    L = LocusTiles();
  }
  size_t n_tiles = db->tiles.size() / HLM_TILE_WORDS;
  db->tile_start[db->n_loci + 1] = (uint32_t)n_tiles;
  // tile -> alleles that hold it (CSR; an allele is listed once per tile), for hlm_score_alleles
  std::sort(tile_al.begin(), tile_al.end());
  tile_al.erase(std::unique(tile_al.begin(), tile_al.end()), tile_al.end());
  db->tile_al_off.assign(n_tiles + 1, 0);
  for (auto& e : tile_al) db->tile_al_off[e.first + 1]++;
  for (size_t t = 0; t < n_tiles; t++) db->tile_al_off[t + 1] += db->tile_al_off[t];
  db->tile_al_ent.resize(tile_al.size());
  for (size_t k = 0; k < tile_al.size(); k++) db->tile_al_ent[k] = tile_al[k].second;
  tile_al = {};
  if (n_tiles >= (1u << 24)) {
    err = "too many distinct tiles for the 24-bit tile id";
    delete db;
    return nullptr;
  }
  DB_T("tiles + hierarchy")
  // children CSR
  db->child_off.assign(n_tiles + 1, 0);
  for (size_t t = 0; t < n_tiles; t++)
    if (db->tile_parent[t] != t) db->child_off[db->tile_parent[t] + 1]++;
  for (size_t t = 0; t < n_tiles; t++) db->child_off[t + 1] += db->child_off[t];
  db->child_ent.resize(db->child_off[n_tiles]);
  db->child_diff.resize(db->child_off[n_tiles]);
  {
    std::vector<uint32_t> cur(db->child_off.begin(), db->child_off.end() - 1);
    for (size_t t = 0; t < n_tiles; t++)
      if (db->tile_parent[t] != t) {
        uint32_t at = cur[db->tile_parent[t]]++;
        db->child_ent[at] = (uint32_t)t;
        db->child_diff[at] = db->tile_diff[t];
      }
  }

  DB_T("children CSR")
  // ---- seed index: CSR over 12-mers.  Reps: every seed start column in [OFF_MIN, 256-12];
  // children: only the seeds that cover one of their differing columns (every other seed of a
  // child is also a seed of its rep; such candidates are produced by expanding the rep).
  db->seed_off.assign((size_t)HLM_SEED_CODES + 1, 0);
  // both passes run over disjoint tile ranges on host threads; slots are claimed with atomic adds
  // (the order inside one 12-mer does not matter: every range is sorted afterwards)
  auto scan_range = [&](bool fill, std::vector<uint32_t>& cursor, size_t t_lo, size_t t_hi) {
    for (size_t t = t_lo; t < t_hi; t++) {
      const uint32_t* T = &db->tiles[t * HLM_TILE_WORDS];
      uint32_t dd = db->tile_diff[t];
      int nd = HLM_DIFF_N(dd);
      uint32_t code = 0;
      int run = 0;
      for (int c = HLM_OFF_MIN; c < HLM_TILE_COLS; c++) {
        int isn = (T[16 + (c >> 5)] >> (c & 31)) & 1;
        if (isn) {
          run = 0;
          code = 0;
          continue;
        }
        uint32_t b = ((T[c >> 5] >> (c & 31)) & 1) | (((T[8 + (c >> 5)] >> (c & 31)) & 1) << 1);
        code = (code >> 2) | (b << (2 * (HLM_SEED - 1)));
        if (++run >= HLM_SEED) {
          int o = c - HLM_SEED + 1;
          if (nd) {
            bool covers = false;
            for (int k = 0; k < nd; k++) {
              int x = HLM_DIFF_COL(dd, k);
              if (x >= o && x < o + HLM_SEED) covers = true;
            }
            if (!covers) continue;
          }
          if (!fill) __atomic_fetch_add(&db->seed_off[code + 1], 1u, __ATOMIC_RELAXED);
          else db->seed_ent[__atomic_fetch_add(&cursor[code], 1u, __ATOMIC_RELAXED)] = ((uint32_t)o << 24) | (uint32_t)t;
        }
      }
    }
  };
  auto scan = [&](bool fill, std::vector<uint32_t>& cursor) {
    unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; k++)
      th.emplace_back([&, k] { scan_range(fill, cursor, n_tiles * k / nt, n_tiles * (k + 1) / nt); });
    for (auto& t : th) t.join();
  };
  std::vector<uint32_t> cursor;
  scan(false, cursor);
  for (size_t c = 0; c < HLM_SEED_CODES; c++) db->seed_off[c + 1] += db->seed_off[c];
  db->seed_ent.resize(db->seed_off[HLM_SEED_CODES]);
  cursor.assign(db->seed_off.begin(), db->seed_off.end() - 1);
  scan(true, cursor);
  DB_T("seed index scan x2")
  // entries of one 12-mer sorted by (column, tile): a read seed at offset 12q can only use columns
  // [OFF_MIN + 12q, OFF_MIN + 32 + 12q), which the kernel finds by binary search
  {
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; k++)
      th.emplace_back([&, k]() {
        size_t lo = (size_t)HLM_SEED_CODES * k / nt, hi = (size_t)HLM_SEED_CODES * (k + 1) / nt;
        for (size_t c = lo; c < hi; c++) {
          uint32_t a = db->seed_off[c], b = db->seed_off[c + 1];
          if (b - a > 1) std::sort(db->seed_ent.begin() + a, db->seed_ent.begin() + b);
        }
      });
    for (auto& t : th) t.join();
  }
  DB_T("seed index sort")
  return db;
}
