/*
 * hlm_oracle.c — CPU oracle for the hla-mapper hot path (see hlm_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product path.
 *
 * Every function names the reference lines (relative to /root/reference) it restates.
 * The aligner section restates no reference code: the reference delegates alignment to
 * the external `bwa aln -o 1 -n 20` + `bwa samse` (src/map_dna.cpp:749-770), which is not
 * available; it implements the scoring definition of DESIGN.md §3 ("parity unpinned").
 */
#define _GNU_SOURCE
#include "hlm_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define K20 HLM_MOTIF
#define KS HLM_SEED
#define BAND HLM_BAND
#define NB (2 * BAND + 1)

/* composite DP value: V = S*SC - cost*CU - lead  (lexicographic max S, min cost, min lead) */
#define SC (1 << 20)
#define CU (1 << 9)
#define V_MATCH (SC)
#define V_MISM (-4 * SC - CU)
#define V_GAPO (-(7 * SC + CU)) /* open 6 + first extension 1 */
#define V_GAPE (-(SC + CU))
#define V_CLIP (5 * SC)
#define V_NEG (-(1 << 30))

/* ------------------------------------------------------------------ k-mer set */
/* exact-string hash set of 20-byte motifs -> 32-bit membership mask
 * (bit 31: motif_all.txt, bit g: loci/<g>.txt).  String semantics on purpose:
 * the reference keys unordered_map<string,int> by the raw line
 * (src/preselect.cpp:297-309, src/map_dna.cpp:575-581). */
typedef struct {
  char k[K20];
  uint32_t mask;
  uint8_t used;
} kent;
typedef struct {
  kent* e;
  size_t cap, n;
} kset;

static uint64_t fnv(const char* s, int n) {
  uint64_t h = 1469598103934665603ULL;
  for (int i = 0; i < n; i++) {
    h ^= (uint8_t)s[i];
    h *= 1099511628211ULL;
  }
  return h ^ (h >> 29);
}
static void kset_init(kset* s, size_t cap) {
  size_t c = 1024;
  while (c < cap * 2) c <<= 1;
  s->e = (kent*)calloc(c, sizeof(kent));
  s->cap = c;
  s->n = 0;
}
static void kset_grow(kset* s);
static void kset_add(kset* s, const char* k, uint32_t mask) {
  if ((s->n + 1) * 2 > s->cap) kset_grow(s);
  size_t i = fnv(k, K20) & (s->cap - 1);
  while (s->e[i].used) {
    if (memcmp(s->e[i].k, k, K20) == 0) {
      s->e[i].mask |= mask;
      return;
    }
    i = (i + 1) & (s->cap - 1);
  }
  memcpy(s->e[i].k, k, K20);
  s->e[i].mask = mask;
  s->e[i].used = 1;
  s->n++;
}
static void kset_grow(kset* s) {
  kset o = *s;
  s->cap = o.cap * 2;
  s->e = (kent*)calloc(s->cap, sizeof(kent));
  s->n = 0;
  for (size_t i = 0; i < o.cap; i++)
    if (o.e[i].used) kset_add(s, o.e[i].k, o.e[i].mask);
  free(o.e);
}
static uint32_t kset_get(const kset* s, const char* k) {
  size_t i = fnv(k, K20) & (s->cap - 1);
  while (s->e[i].used) {
    if (memcmp(s->e[i].k, k, K20) == 0) return s->e[i].mask;
    i = (i + 1) & (s->cap - 1);
  }
  return 0;
}

/* ------------------------------------------------------------------ database */
typedef struct {
  uint8_t* codes; /* 0..3 = ACGT, 4 = anything else */
  int len;
} oseq;
typedef struct {
  uint64_t key; /* seed code << 40 | allele << 16 | pos */
} oseed;
typedef struct {
  char name[64];
  oseq* al;
  int64_t n_al;
  oseed* seeds; /* sorted: every 12-mer of every allele */
  int64_t n_seeds;
  int rna; /* target set of an `rna` database: short units use overlapping seeds (collect_cands) */
} olocus;
struct orc_db {
  int rna, n_loci, v_size;
  olocus loc[HLM_MAX_LOCI + 1]; /* [n_loci] = others */
  kset motifs;
};

static int code_of(int c) {
  switch (toupper(c)) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
  }
}

static int load_motifs(kset* s, const char* path, uint32_t mask) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  char* line = NULL;
  size_t cap = 0;
  ssize_t n;
  while ((n = getline(&line, &cap, f)) >= 0) {
    if (n > 0 && line[n - 1] == '\n') n--;
    /* a line of any other length can never equal seq.substr(c,20) */
    if (n == K20) kset_add(s, line, mask);
  }
  free(line);
  fclose(f);
  return 0;
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

static int load_fasta(olocus* L, const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  char* line = NULL;
  size_t cap = 0;
  ssize_t n;
  int64_t acap = 0;
  int64_t scap = 0;
  oseq* cur = NULL;
  while ((n = getline(&line, &cap, f)) >= 0) {
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) n--;
    if (n == 0) continue;
    if (line[0] == '>') {
      if (L->n_al == acap) {
        acap = acap ? acap * 2 : 64;
        L->al = (oseq*)realloc(L->al, acap * sizeof(oseq));
      }
      cur = &L->al[L->n_al++];
      cur->codes = NULL;
      cur->len = 0;
      scap = 0;
      continue;
    }
    if (!cur) continue;
    if (cur->len + n > scap) {
      scap = (cur->len + n) * 2 + 64;
      cur->codes = (uint8_t*)realloc(cur->codes, scap);
    }
    for (ssize_t i = 0; i < n; i++) cur->codes[cur->len++] = (uint8_t)code_of(line[i]);
  }
  free(line);
  fclose(f);
  /* seed list: every N-free 12-mer of every allele */
  int64_t tot = 0;
  for (int64_t a = 0; a < L->n_al; a++) tot += L->al[a].len;
  L->seeds = (oseed*)malloc((tot + 1) * sizeof(oseed));
  L->n_seeds = 0;
  for (int64_t a = 0; a < L->n_al; a++) {
    const oseq* s = &L->al[a];
    uint32_t code = 0;
    int run = 0;
    for (int p = 0; p < s->len; p++) {
      if (s->codes[p] > 3) {
        run = 0;
        code = 0;
        continue;
      }
      code = ((code << 2) | s->codes[p]) & 0xFFFFFFu;
      if (++run >= KS)
        L->seeds[L->n_seeds++].key =
            ((uint64_t)code << 40) | ((uint64_t)a << 16) | (uint64_t)(p - KS + 1);
    }
  }
  qsort(L->seeds, L->n_seeds, sizeof(oseed), cmp_u64);
  return 0;
}

/* db_{dna,rna}.info: src/map_dna.cpp:299-365, src/map_rna.cpp:281-335 */
orc_db* orc_db_open(const char* dir, int rna, char* err, size_t errlen) {
  char path[4096];
  const char* kind = rna ? "rna" : "dna";
  snprintf(path, sizeof path, "%s/db_%s.info", dir, kind);
  FILE* f = fopen(path, "rb");
  if (!f) {
    if (err) snprintf(err, errlen, "Error accessing database %s", path);
    return NULL;
  }
  orc_db* db = (orc_db*)calloc(1, sizeof(orc_db));
  db->rna = rna;
  db->v_size = 50; /* src/main.cpp:79 */
  int ok = 0;
  char* line = NULL;
  size_t cap = 0;
  ssize_t n;
  while ((n = getline(&line, &cap, f)) >= 0) {
    if (n > 0 && line[n - 1] == '\n') line[--n] = 0;
    if (strcmp(line, "hla-mapper:4.3+") == 0) {
      ok = 1;
      continue;
    }
    if (strncmp(line, "GENE:", 5) == 0) {
      char* p = line + 5;
      char* q = strchr(p, ':');
      if (q) *q = 0;
      if (db->n_loci < HLM_MAX_LOCI) {
        /* the reference erases spaces from the gene list (map_dna.cpp:515) */
        char* d = db->loc[db->n_loci].name;
        for (const char* c = p; *c && d - db->loc[db->n_loci].name < 63; c++)
          if (*c != ' ') *d++ = *c;
        *d = 0;
        if (db->loc[db->n_loci].name[0]) db->n_loci++;
      }
    } else if (strncmp(line, "SIZE:", 5) == 0) {
      db->v_size = atoi(line + 5);
    }
  }
  free(line);
  fclose(f);
  if (!ok) {
    if (err)
      snprintf(err, errlen, "The database is not compatible with this version of hla-mapper.");
    free(db);
    return NULL;
  }
  strcpy(db->loc[db->n_loci].name, "others");
  kset_init(&db->motifs, 1 << 16);
  snprintf(path, sizeof path, "%s/motif/%s/all/motif_all.txt", dir, kind);
  load_motifs(&db->motifs, path, 1u << 31);
  for (int g = 0; g < db->n_loci; g++) {
    snprintf(path, sizeof path, "%s/motif/%s/loci/%s.txt", dir, kind, db->loc[g].name);
    load_motifs(&db->motifs, path, 1u << g);
    snprintf(path, sizeof path, "%s/mapper/%s/%s/%s.fas", dir, kind, db->loc[g].name,
             db->loc[g].name);
    load_fasta(&db->loc[g], path);
    db->loc[g].rna = rna;
  }
  snprintf(path, sizeof path, "%s/others/%s/hla_gen_others.fasta", dir, kind);
  load_fasta(&db->loc[db->n_loci], path);
  db->loc[db->n_loci].rna = rna;
  return db;
}
void orc_db_close(orc_db* db) {
  if (!db) return;
  for (int g = 0; g <= db->n_loci; g++) {
    for (int64_t a = 0; a < db->loc[g].n_al; a++) free(db->loc[g].al[a].codes);
    free(db->loc[g].al);
    free(db->loc[g].seeds);
  }
  free(db->motifs.e);
  free(db);
}
int orc_db_n_loci(const orc_db* db) { return db->n_loci; }
const char* orc_db_locus_name(const orc_db* db, int i) {
  return (i >= 0 && i <= db->n_loci) ? db->loc[i].name : NULL;
}
int orc_db_size(const orc_db* db) { return db->v_size; }
int64_t orc_db_n_alleles(const orc_db* db, int g) { return db->loc[g].n_al; }

/* ------------------------------------------------------------------ mtrim */
/* src/functions.cpp:89-145.  Returns 1 and (lo,len) when the trimmed read is kept
 * (len >= v_size), 0 when mtrim() returns "".  qlen = qual.length(), slen = seq.size(). */
int orc_mtrim(const uint8_t* qual, int qlen, int slen, double err, int v_size, int* lo_out,
              int* len_out) {
  int start = -1, end = -1, low = 0, high = 0;
  for (int a = 0; a < qlen; a++) {
    double chr_phred = (double)((int)(char)qual[a] - 33); /* phred(), functions.cpp:77-79 */
    double P = pow((double)10, (chr_phred / (double)(-10)));
    if ((P <= err) && (start == -1)) {
      start = a;
      end = a;
      continue;
    }
    if ((P > err) && (start >= 0)) {
      end = a - 1;
      if ((end - start) > (high - low)) {
        low = start;
        high = end;
      }
      start = -1;
      end = -1;
      continue;
    }
  }
  if (start >= 0) {
    end = qlen; /* one past: functions.cpp:121 */
    if ((end - start) > (high - low)) {
      low = start;
      high = end;
    }
  }
  int len = 0;
  if (low < high) {
    /* seq.substr(low, high-low+1) clamps to the end of seq */
    len = high - low + 1;
    if (low > slen) len = 0; /* substr would throw; unreachable when qlen == slen */
    else if (low + len > slen) len = slen - low;
  }
  *lo_out = low;
  *len_out = len;
  return len >= v_size && len > 0;
}

/* ------------------------------------------------------------------ screens */
/* src/preselect.cpp:707-731 (FASTQ: `if hit {hit++;c++} ... c++` inside `for(;;c++)`)
 * src/preselect.cpp:484-506 (BAM-derived: `if hit {hit++; c=c+1}` only).
 * Length pre-checks are the caller's (they differ between the two loops). */
int orc_screen_read(const orc_db* db, const uint8_t* seq, int len, int variant, int minhit) {
  int hitcount = 0;
  for (int c = 0; c < len - K20; c++) {
    if (kset_get(&db->motifs, (const char*)seq + c) & (1u << 31)) {
      hitcount++;
      c++;
    }
    if (hitcount == minhit) return 1;
    if (variant == HLM_SCREEN_FASTQ) c++;
  }
  return 0;
}

/* src/map_dna.cpp:611-627 / src/map_rna.cpp:544-560: stride 1 over the trimmed R1,
 * `c < size - 20` so the last window is never tested. */
uint32_t orc_locus_mask(const orc_db* db, const uint8_t* seq, int len) {
  uint32_t m = 0;
  for (int c = 0; c < len - K20; c++) m |= kset_get(&db->motifs, (const char*)seq + c);
  return m & ~(1u << 31);
}

int orc_screen_trim(const orc_db* db, const hlm_params* p, const hlm_batch* in, hlm_screen_out* out,
                    int nthreads) {
  int v_size = p->v_size > 0 ? p->v_size : db->v_size;
  int64_t n = in->n_pairs;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; i++) {
    const uint8_t* s1 = in->bases1 + in->offs1[i];
    int l1 = (int)(in->offs1[i + 1] - in->offs1[i]);
    int l2 = 0;
    if (p->paired) l2 = (int)(in->offs2[i + 1] - in->offs2[i]);
    uint8_t fl = 0;
    int lo1 = 0, n1 = 0, lo2 = 0, n2 = 0;
    uint32_t mask = 0;
    int ok = 1;
    /* preselect.cpp:707-710 (paired FASTQ), :484-485 (BAM: R1 only), :918-919 (single) */
    if (l1 < K20 || l1 < v_size) ok = 0;
    if (p->paired && p->screen_variant == HLM_SCREEN_FASTQ && (l2 < K20 || l2 < v_size)) ok = 0;
    if (ok && (p->skip_screen || orc_screen_read(db, s1, l1, p->screen_variant, p->minhit)))
      fl |= HLM_F_SCREENED;
    if (fl & HLM_F_SCREENED) {
      /* preselect.cpp:1060-1149: both mates must survive mtrim */
      int k1 = orc_mtrim(in->quals1 + in->offs1[i], l1, l1, p->mtrim_error, v_size, &lo1, &n1);
      int k2 = 1;
      if (p->paired)
        k2 = orc_mtrim(in->quals2 + in->offs2[i], l2, l2, p->mtrim_error, v_size, &lo2, &n2);
      if (!k1) n1 = 0;
      if (!k2) n2 = 0;
      if (k1 && k2) {
        fl |= HLM_F_KEPT;
        /* map_dna.cpp:611-613: both trimmed mates need >= 20 bases */
        if (n1 >= K20 && (!p->paired || n2 >= K20)) mask = orc_locus_mask(db, s1 + lo1, n1);
      }
    }
    out->flags[i] = fl;
    out->trim_lo1[i] = (uint16_t)lo1;
    out->trim_len1[i] = (uint16_t)n1;
    if (out->trim_lo2) out->trim_lo2[i] = (uint16_t)lo2;
    if (out->trim_len2) out->trim_len2[i] = (uint16_t)n2;
    out->locus_mask[i] = mask;
  }
  return 0;
}

/* ------------------------------------------------------------------ aligner */
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int refc(const uint8_t* t, int tlen, int x) { return (x < 0 || x >= tlen) ? 4 : t[x]; }

/* Banded affine-gap alignment of the whole read r[0..n) against t on diagonal d
 * (cell (i,j): i read bases and j reference bases consumed, |j - i - d| <= BAND).
 * match +1, mismatch -4, gap of length k -(6+k), clipping an end -5; the reference
 * is an infinite string of N outside [0,tlen); N never matches.
 * Among co-optimal paths: max S, then min (NM + clipped bases), then min leading clip,
 * then the end row with the shortest trailing clip. */
/* HLM_CLIP_NONE (hlm_params.clip_mode = 2): the alignment never clips - no row starts a new
 * alignment, only the last row ends one.  Set by the whole-stage entry points from the parameters
 * (and by the stand-in bwa from HLM_ORACLE_NOCLIP); read-only while a stage runs. */
static int g_noclip = 0;
void orc_set_noclip(int on) { g_noclip = on ? 1 : 0; }

void orc_align(const uint8_t* r, int n, const uint8_t* t, int tlen, int d, int* S_out,
               int* cost_out, int* lead_out, int* trail_out) {
  int Hp[NB], Ep[NB], Hc[NB], Ec[NB];
  for (int k = 0; k < NB; k++) {
    Hp[k] = 0;
    Ep[k] = V_NEG;
  }
  int best = V_NEG, best_i = 0;
  for (int i = 1; i <= n; i++) {
    int start = -(V_CLIP + i * CU + i); /* clip the first i bases and start here */
    if (g_noclip) start = V_NEG;
    int rb = r[i - 1];
    int f = V_NEG, hleft = V_NEG;
    int rowmax = V_NEG;
    for (int k = 0; k < NB; k++) {
      int j = i + d - BAND + k;
      int tb = refc(t, tlen, j - 1);
      int sub = (rb < 4 && rb == tb) ? V_MATCH : V_MISM;
      int h = Hp[k] + sub;
      int e = (k + 1 < NB) ? imax(Ep[k + 1] + V_GAPE, Hp[k + 1] + V_GAPO) : V_NEG;
      f = (k >= 1) ? imax(f + V_GAPE, hleft + V_GAPO) : V_NEG;
      h = imax(imax(h, start), imax(e, f));
      Hc[k] = h;
      Ec[k] = e;
      hleft = h;
      rowmax = imax(rowmax, h);
    }
    int fin = rowmax - (i < n ? V_CLIP + (n - i) * CU : 0);
    if (g_noclip && i < n) fin = V_NEG - 1;
    if (fin >= best) {
      best = fin;
      best_i = i;
    }
    memcpy(Hp, Hc, sizeof Hp);
    memcpy(Ep, Ec, sizeof Ep);
  }
  int S = (best + SC - 1) >> 20; /* ceil(best / SC) */
  int rem = S * SC - best;
  *S_out = S;
  *cost_out = rem >> 9;
  *lead_out = rem & 511;
  *trail_out = n - best_i;
}

/* unit-cost banded semi-global distance over the same band: a lower bound of orc_align's cost */
int orc_banded_ed(const uint8_t* r, int n, const uint8_t* t, int tlen, int d) {
  int Dp[NB], Dc[NB];
  const int INF = 1 << 20;
  for (int k = 0; k < NB; k++) Dp[k] = 0;
  for (int i = 1; i <= n; i++) {
    int rb = r[i - 1];
    for (int k = 0; k < NB; k++) {
      int j = i + d - BAND + k;
      int tb = refc(t, tlen, j - 1);
      int v = Dp[k] + ((rb < 4 && rb == tb) ? 0 : 1);
      if (k + 1 < NB) v = imin(v, Dp[k + 1] + 1);
      if (k >= 1) v = imin(v, Dc[k - 1] + 1);
      Dc[k] = imin(v, INF);
    }
    memcpy(Dp, Dc, sizeof Dp);
  }
  int m = INF;
  for (int k = 0; k < NB; k++) m = imin(m, Dp[k]);
  return m;
}

typedef struct {
  int cost, S, lead, trail;
  int valid;
} ares;

static inline int ares_better(const ares* a, const ares* b) { /* a strictly better than b */
  if (!b->valid) return a->valid;
  if (!a->valid) return 0;
  if (a->cost != b->cost) return a->cost < b->cost;
  if (a->S != b->S) return a->S > b->S;
  if (a->lead != b->lead) return a->lead < b->lead;
  return a->trail < b->trail;
}

typedef struct {
  int32_t allele;
  int32_t d;
} cand;
static int cmp_cand(const void* a, const void* b) {
  const cand* x = (const cand*)a;
  const cand* y = (const cand*)b;
  if (x->allele != y->allele) return x->allele < y->allele ? -1 : 1;
  return x->d < y->d ? -1 : x->d > y->d;
}

/* all (allele, diagonal) pairs where some seed matches the allele exactly.  Seeds (DESIGN.md
 * section 3): 12-mers at read offsets 0, st, 2 st, ... <= n - 12; st = 12 (non-overlapping), except
 * for the units of an `rna` run shorter than HLM_SEED_SHORT (96) bases - the blocks STAR cuts a
 * mate into - which use st = 6 (overlapping) */
static int64_t collect_cands(const olocus* L, const uint8_t* r, int n, cand** buf, int64_t* cap) {
  int64_t nc = 0;
  const int st = (L->rna && n < HLM_SEED_SHORT) ? HLM_SEED_STRIDE_SHORT : KS;
  for (int rp = 0; rp + KS <= n; rp += st) {
    uint32_t code = 0;
    int bad = 0;
    for (int x = 0; x < KS; x++) {
      int c = r[rp + x];
      if (c > 3) bad = 1;
      code = (code << 2) | (c & 3);
    }
    if (bad) continue;
    uint64_t lo_key = (uint64_t)code << 40, hi_key = ((uint64_t)code + 1) << 40;
    int64_t lo = 0, hi = L->n_seeds;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (L->seeds[mid].key < lo_key) lo = mid + 1;
      else hi = mid;
    }
    for (int64_t s = lo; s < L->n_seeds && L->seeds[s].key < hi_key; s++) {
      if (nc == *cap) {
        *cap = *cap ? *cap * 2 : 1024;
        *buf = (cand*)realloc(*buf, *cap * sizeof(cand));
      }
      (*buf)[nc].allele = (int32_t)((L->seeds[s].key >> 16) & 0xFFFFFF);
      (*buf)[nc].d = (int32_t)(L->seeds[s].key & 0xFFFF) - rp;
      nc++;
    }
  }
  if (nc > 1) {
    qsort(*buf, nc, sizeof(cand), cmp_cand);
    int64_t w = 1;
    for (int64_t i = 1; i < nc; i++)
      if (cmp_cand(&(*buf)[i], &(*buf)[w - 1]) != 0) (*buf)[w++] = (*buf)[i];
    nc = w;
  }
  return nc;
}

static void revcomp(const uint8_t* r, int n, uint8_t* out) {
  for (int i = 0; i < n; i++) {
    int c = r[n - 1 - i];
    out[i] = (uint8_t)(c > 3 ? 4 : 3 - c);
  }
}

typedef struct {
  cand* buf;
  int64_t cap;
  /* fast mode scratch */
  uint64_t* fh;
  int32_t* fidx;
  int32_t* lb;
  int64_t fcap;
} scratch;

/* exhaustive definition: DP every (allele, strand, diagonal) candidate */
static ares score_exhaustive(const olocus* L, const uint8_t* r, int n, scratch* sc,
                             int64_t* cells) {
  ares best = {0, 0, 0, 0, 0};
  uint8_t rc[HLM_MAX_READ + 8];
  for (int strand = 0; strand < 2; strand++) {
    const uint8_t* q = r;
    if (strand) {
      revcomp(r, n, rc);
      q = rc;
    }
    int64_t nc = collect_cands(L, q, n, &sc->buf, &sc->cap);
    for (int64_t c = 0; c < nc; c++) {
      const oseq* a = &L->al[sc->buf[c].allele];
      ares x;
      orc_align(q, n, a->codes, a->len, sc->buf[c].d, &x.S, &x.cost, &x.lead, &x.trail);
      x.valid = 1;
      *cells += (int64_t)n * NB;
      if (ares_better(&x, &best)) best = x;
    }
  }
  return best;
}

/* fast CPU port: identical candidates, but (1) candidates whose DP footprint
 * t[d-BAND .. d+n+BAND) is byte-identical are evaluated once, (2) the unit-cost banded
 * distance orders them and only candidates that can still tie or beat the best cost get
 * the affine DP.  Must return exactly what score_exhaustive returns. */
static ares score_fast(const olocus* L, const uint8_t* r, int n, scratch* sc, int mm_max,
                       int64_t* cells) {
  ares best = {0, 0, 0, 0, 0};
  uint8_t rc[HLM_MAX_READ + 8];
  uint8_t fp[HLM_MAX_READ + 2 * BAND + 8], fp2[HLM_MAX_READ + 2 * BAND + 8];
  int flen = n + 2 * BAND;
  for (int strand = 0; strand < 2; strand++) {
    const uint8_t* q = r;
    if (strand) {
      revcomp(r, n, rc);
      q = rc;
    }
    int64_t nc = collect_cands(L, q, n, &sc->buf, &sc->cap);
    if (nc > sc->fcap) {
      sc->fcap = nc * 2;
      sc->fh = (uint64_t*)realloc(sc->fh, sc->fcap * sizeof(uint64_t));
      sc->fidx = (int32_t*)realloc(sc->fidx, sc->fcap * sizeof(int32_t));
      sc->lb = (int32_t*)realloc(sc->lb, sc->fcap * sizeof(int32_t));
    }
    /* footprint hash per candidate, then keep the first of each identical footprint */
    int64_t nu = 0;
    for (int64_t c = 0; c < nc; c++) {
      const oseq* a = &L->al[sc->buf[c].allele];
      int d = sc->buf[c].d;
      uint64_t h = 1469598103934665603ULL;
      for (int x = 0; x < flen; x++) {
        h ^= (uint64_t)refc(a->codes, a->len, d - BAND + x);
        h *= 1099511628211ULL;
      }
      int dup = 0;
      for (int64_t u = 0; u < nu && !dup; u++) {
        if (sc->fh[u] != h) continue;
        const oseq* b = &L->al[sc->buf[sc->fidx[u]].allele];
        int db_ = sc->buf[sc->fidx[u]].d;
        for (int x = 0; x < flen; x++) {
          fp[x] = (uint8_t)refc(a->codes, a->len, d - BAND + x);
          fp2[x] = (uint8_t)refc(b->codes, b->len, db_ - BAND + x);
        }
        if (memcmp(fp, fp2, flen) == 0) dup = 1;
      }
      if (!dup) {
        sc->fh[nu] = h;
        sc->fidx[nu] = (int32_t)c;
        nu++;
      }
    }
    int emin = 1 << 20;
    for (int64_t u = 0; u < nu; u++) {
      const cand* c = &sc->buf[sc->fidx[u]];
      const oseq* a = &L->al[c->allele];
      sc->lb[u] = orc_banded_ed(q, n, a->codes, a->len, c->d);
      if (sc->lb[u] < emin) emin = sc->lb[u];
    }
    /* round 1: lb == emin; round 2: emin < lb <= best cost so far */
    for (int round = 0; round < 2; round++) {
      for (int64_t u = 0; u < nu; u++) {
        int lb = sc->lb[u];
        if (lb > mm_max) continue;
        if (round == 0 ? (lb != emin) : !(lb > emin && best.valid && lb <= best.cost)) continue;
        const cand* c = &sc->buf[sc->fidx[u]];
        const oseq* a = &L->al[c->allele];
        ares x;
        orc_align(q, n, a->codes, a->len, c->d, &x.S, &x.cost, &x.lead, &x.trail);
        x.valid = 1;
        *cells += (int64_t)n * NB;
        if (ares_better(&x, &best)) best = x;
      }
    }
  }
  return best;
}

/* Seedless check of the aligner definition (mode 2): EVERY diagonal of every allele, both
 * strands, whose band can hold an alignment of cost <= mm_max - no seed index involved.  Sellers'
 * unbanded semi-global edit distance finds every reference end position e (reference bases
 * consumed, N beyond the allele ends) with distance <= mm_max; an alignment of cost <= mm_max
 * that ends on diagonal e - n lies in the band of every candidate diagonal within BAND of it, so
 * all of those get the affine DP.  The result is the same lexicographic minimum as
 * score_exhaustive, taken over this superset of its candidates.  Test infrastructure for the
 * completeness statement of DESIGN.md section 3 (small databases only: O(n * allele length) per
 * mate, allele and strand). */
static ares score_seedless(const olocus* L, const uint8_t* r, int n, int mm_max, int64_t* cells) {
  ares best = {0, 0, 0, 0, 0};
  uint8_t rc[HLM_MAX_READ + 8];
  int D[HLM_MAX_READ + 2];
  const int P = BAND + mm_max + 1;
  for (int strand = 0; strand < 2; strand++) {
    const uint8_t* q = r;
    if (strand) {
      revcomp(r, n, rc);
      q = rc;
    }
    for (int al = 0; al < L->n_al; al++) {
      const oseq* a = &L->al[al];
      int tlen = a->len;
      int dlo = -n - P - BAND, nd = tlen + 2 * P + n + 2 * BAND + 4;
      uint8_t* mark = (uint8_t*)calloc((size_t)nd, 1);
      for (int i = 0; i <= n; i++) D[i] = i;
      for (int x = -P; x < tlen + P; x++) {
        int tb = refc(a->codes, tlen, x);
        int diag = D[0]; /* D_old[i-1] */
        D[0] = 0;
        for (int i = 1; i <= n; i++) {
          int up = D[i];
          int v = diag + ((q[i - 1] < 4 && q[i - 1] == tb) ? 0 : 1);
          if (up + 1 < v) v = up + 1;
          if (D[i - 1] + 1 < v) v = D[i - 1] + 1;
          diag = up;
          D[i] = v;
        }
        if (D[n] <= mm_max) {
          int dc = (x + 1) - n;
          for (int d = dc - BAND; d <= dc + BAND; d++) mark[d - dlo] = 1;
        }
      }
      for (int k = 0; k < nd; k++) {
        if (!mark[k]) continue;
        ares xr;
        orc_align(q, n, a->codes, tlen, k + dlo, &xr.S, &xr.cost, &xr.lead, &xr.trail);
        xr.valid = 1;
        *cells += (int64_t)n * NB;
        if (ares_better(&xr, &best)) best = xr;
      }
      free(mark);
    }
  }
  return best;
}

static ares score_task(int mode, const olocus* L, const uint8_t* r, int n, scratch* sc, int mm_max,
                       int64_t* cells) {
  if (mode == 2) return score_seedless(L, r, n, mm_max, cells);
  return mode ? score_fast(L, r, n, sc, mm_max, cells) : score_exhaustive(L, r, n, sc, cells);
}

static void to_codes(const uint8_t* s, int n, uint8_t* out) {
  for (int i = 0; i < n; i++) {
    /* reads are matched as written: only upper-case ACGT are bases */
    switch (s[i]) {
      case 'A': out[i] = 0; break;
      case 'C': out[i] = 1; break;
      case 'G': out[i] = 2; break;
      case 'T': out[i] = 3; break;
      default: out[i] = 4;
    }
  }
}

/* hla-mapper score of one reported alignment, as read_sam_dna would compute it from the
 * SAM record (src/map_dna.cpp:56-78): NM + leading clip (+ trailing clip in CLIP_BOTH) */
static inline int hla_dna_score(const ares* a, int clip_mode) {
  int nm = a->cost - a->lead - a->trail;
  return nm + a->lead + (clip_mode == HLM_CLIP_BOTH ? a->trail : 0) + 1;
}
/* read_sam_rna (src/map_rna.cpp:67-96): clips are only added while mm != len(SEQ) */
static inline int hla_rna_score(const ares* a, int len, int clip_mode) {
  int mm = a->cost - a->lead - a->trail;
  if (a->lead > 0 && mm != len) mm += a->lead;
  if (clip_mode == HLM_CLIP_BOTH && a->trail > 0 && mm != len) mm += a->trail;
  return mm;
}

int orc_score(const orc_db* db, const hlm_params* p, const hlm_batch* in, const hlm_screen_out* scr,
              hlm_score_out* out, int mode, int nthreads, int64_t* n_dp_cells) {
  int64_t n = in->n_pairs;
  g_noclip = p->clip_mode == HLM_CLIP_NONE;
  int T = db->n_loci + 1;
  int64_t total_cells = 0;
  int toolong = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : total_cells)
  {
    scratch sc;
    memset(&sc, 0, sizeof sc);
    uint8_t codes[HLM_MAX_READ + 8];
#pragma omp for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; i++) {
      for (int g = 0; g < T; g++) {
        int64_t o = i * T + g;
        out->s1[o] = p->rna ? 0xFFFF : 0;
        if (out->s2) out->s2[o] = 0;
        if (out->aln1) out->aln1[o] = 0;
        if (out->aln2) out->aln2[o] = 0;
        if (out->nm1) out->nm1[o] = 0;
        if (out->nm2) out->nm2[o] = 0;
        if (out->lead1) out->lead1[o] = 0;
        if (out->lead2) out->lead2[o] = 0;
        if (out->trail1) out->trail1[o] = 0;
        if (out->trail2) out->trail2[o] = 0;
      }
      if (!(scr->flags[i] & HLM_F_KEPT)) continue;
      int nm_ = p->paired ? 2 : 1;
      for (int m = 0; m < nm_; m++) {
        const uint8_t* s = (m == 0 ? in->bases1 + in->offs1[i] + scr->trim_lo1[i]
                                   : in->bases2 + in->offs2[i] + scr->trim_lo2[i]);
        int len = m == 0 ? scr->trim_len1[i] : scr->trim_len2[i];
        if (len > HLM_MAX_READ) {
          toolong = 1;
          continue;
        }
        to_codes(s, len, codes);
        int split = 0;
        if (p->rna) {
          const uint16_t* sp = m == 0 ? in->rna_split1 : in->rna_split2;
          if (sp && sp[i] > 0 && sp[i] < len) split = sp[i];
        }
        if (p->rna && in->rna_block_off) {
          /* the STAR pre-split in full (map_rna.cpp:740-790): `others` sees the whole mate
           * (map_rna.cpp:858-892), every gene exactly the records listed for it; each record is
           * one line of read_sam_rna (map_rna.cpp:39-101) */
          {
            int64_t o = i * T + db->n_loci;
            int sum = out->s1[o] == 0xFFFF ? 0 : out->s1[o];
            ares a = score_task(mode, &db->loc[db->n_loci], codes, len, &sc, p->mm_max, &total_cells);
            if (a.valid && a.cost <= p->mm_max) sum += hla_rna_score(&a, len, p->clip_mode);
            else sum += len;
            out->s1[o] = (uint16_t)sum;
          }
          for (uint64_t b = in->rna_block_off[i]; b < in->rna_block_off[i + 1]; b++) {
            hlm_rna_block blk = in->rna_blocks[b];
            if (blk.mate != m || blk.gene >= db->n_loci) continue;
            int st = blk.start < len ? blk.start : len;
            int bl = blk.len < len - st ? blk.len : len - st;
            int64_t o = i * T + blk.gene;
            int sum = out->s1[o] == 0xFFFF ? 0 : out->s1[o];
            if (bl > 0) {
              ares a = score_task(mode, &db->loc[blk.gene], codes + st, bl, &sc, p->mm_max, &total_cells);
              if (a.valid && a.cost <= p->mm_max) sum += hla_rna_score(&a, bl, p->clip_mode);
              else sum += bl;
            }
            out->s1[o] = (uint16_t)sum;
          }
          continue;
        }
        for (int g = 0; g < T; g++) {
          /* loci: only pairs sorted into the locus (map_dna.cpp:611-627);
           * others: every trimmed pair (map_dna.cpp:824-860) */
          if (g < db->n_loci && !((scr->locus_mask[i] >> g) & 1)) continue;
          int64_t o = i * T + g;
          if (!p->rna) {
            ares a = score_task(mode, &db->loc[g], codes, len, &sc, p->mm_max, &total_cells);
            if (a.valid && a.cost <= p->mm_max) {
              uint16_t sv = (uint16_t)hla_dna_score(&a, p->clip_mode);
              if (m == 0) {
                out->s1[o] = sv;
                if (out->aln1) out->aln1[o] = (int16_t)a.S;
                if (out->nm1) out->nm1[o] = (uint8_t)(a.cost - a.lead - a.trail);
                if (out->lead1) out->lead1[o] = (uint8_t)a.lead;
                if (out->trail1) out->trail1[o] = (uint8_t)a.trail;
              } else {
                out->s2[o] = sv;
                if (out->aln2) out->aln2[o] = (int16_t)a.S;
                if (out->nm2) out->nm2[o] = (uint8_t)(a.cost - a.lead - a.trail);
                if (out->lead2) out->lead2[o] = (uint8_t)a.lead;
                if (out->trail2) out->trail2[o] = (uint8_t)a.trail;
              }
            }
          } else {
            /* rna: blocks of a mate are scored separately and summed over blocks and
             * mates (map_rna.cpp:39-101); "others" sees whole mates (map_rna.cpp:858-892) */
            int nb = (split && g < db->n_loci) ? 2 : 1;
            int sum = out->s1[o] == 0xFFFF ? 0 : out->s1[o];
            for (int b = 0; b < nb; b++) {
              int b0 = (nb == 2 && b == 1) ? split : 0;
              int bl = (nb == 2) ? (b == 0 ? split : len - split) : len;
              ares a = score_task(mode, &db->loc[g], codes + b0, bl, &sc, p->mm_max, &total_cells);
              if (a.valid && a.cost <= p->mm_max) sum += hla_rna_score(&a, bl, p->clip_mode);
              else sum += bl; /* unmapped record adds len(SEQ), map_rna.cpp:55-65 */
            }
            out->s1[o] = (uint16_t)sum;
          }
        }
      }
    }
    free(sc.buf);
    free(sc.fh);
    free(sc.fidx);
    free(sc.lb);
  }
  if (n_dp_cells) *n_dp_cells = total_cells;
  return toolong ? HLM_E_TOOLONG : 0;
}

/* Per-allele scoring (what `bwa mem -a` of a gene's reads against
 * every allele feeds src/functions.cpp:347-520).  The aligner definition of DESIGN.md section 3
 * restricted to ONE allele at a time: for every kept pair, mate and allele of the target set, the
 * lexicographic minimum (cost, -S, lead, trail) over that allele's seed candidates, both strands.
 * nm = 255: no candidate, or cost > mm_max.  Arrays [n_pairs * n_alleles], pair-major. */
int orc_score_alleles(const orc_db* db, const hlm_params* p, const hlm_batch* in,
                      const hlm_screen_out* scr, int locus, uint8_t* nm1, uint8_t* clip1, uint8_t* nm2,
                      uint8_t* clip2, int nthreads) {
  int64_t n = in->n_pairs;
  g_noclip = p->clip_mode == HLM_CLIP_NONE;
  const olocus* L = &db->loc[locus];
  int64_t nA = L->n_al;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    scratch sc;
    memset(&sc, 0, sizeof sc);
    uint8_t codes[HLM_MAX_READ + 8], rc[HLM_MAX_READ + 8];
    ares* best = (ares*)malloc((size_t)nA * sizeof(ares));
#pragma omp for schedule(dynamic, 4)
    for (int64_t i = 0; i < n; i++) {
      for (int m = 0; m < (p->paired ? 2 : 1); m++) {
        uint8_t* nm = m == 0 ? nm1 : nm2;
        uint8_t* clip = m == 0 ? clip1 : clip2;
        for (int64_t a = 0; a < nA; a++) {
          nm[i * nA + a] = 255;
          if (clip) clip[i * nA + a] = 0;
        }
        if (!(scr->flags[i] & HLM_F_KEPT)) continue;
        const uint8_t* s = (m == 0 ? in->bases1 + in->offs1[i] + scr->trim_lo1[i]
                                   : in->bases2 + in->offs2[i] + scr->trim_lo2[i]);
        int len = m == 0 ? scr->trim_len1[i] : scr->trim_len2[i];
        if (len > HLM_MAX_READ || len <= 0) continue;
        to_codes(s, len, codes);
        memset(best, 0, (size_t)nA * sizeof(ares));
        for (int strand = 0; strand < 2; strand++) {
          const uint8_t* q = codes;
          if (strand) {
            revcomp(codes, len, rc);
            q = rc;
          }
          int64_t nc = collect_cands(L, q, len, &sc.buf, &sc.cap);
          for (int64_t c = 0; c < nc; c++) {
            const oseq* al = &L->al[sc.buf[c].allele];
            ares x;
            orc_align(q, len, al->codes, al->len, sc.buf[c].d, &x.S, &x.cost, &x.lead, &x.trail);
            x.valid = 1;
            if (ares_better(&x, &best[sc.buf[c].allele])) best[sc.buf[c].allele] = x;
          }
        }
        for (int64_t a = 0; a < nA; a++)
          if (best[a].valid && best[a].cost <= p->mm_max) {
            nm[i * nA + a] = (uint8_t)(best[a].cost - best[a].lead - best[a].trail);
            if (clip) clip[i * nA + a] = (uint8_t)(best[a].lead + best[a].trail);
          }
      }
    }
    free(best);
    free(sc.buf);
  }
  return 0;
}

/* ------------------------------------------------------------------ compare / assign */
/* dna: src/map_dna.cpp:903-1029; rna: src/map_rna.cpp:925-1027 */
int orc_assign(const orc_db* db, const hlm_params* p, const hlm_batch* in, const hlm_screen_out* scr,
               const hlm_score_out* sc, hlm_assign_out* out) {
  int64_t n = in->n_pairs;
  int G = db->n_loci, T = G + 1;
  for (int64_t i = 0; i < n; i++) {
    uint32_t tmask = 0, vmask = 0;
    int minsum = 0;
    int o1 = 0, o2 = 0;
    if (scr->flags[i] & HLM_F_KEPT) {
      int len1 = scr->trim_len1[i];
      int len2 = p->paired ? scr->trim_len2[i] : 0;
      if (!p->paired && in->size_override1) len1 = in->size_override1[i]; /* preselect.cpp:1199 */
      if (!p->rna) {
        int nm_sum[HLM_MAX_LOCI + 1];
        o1 = sc->s1[i * T + G];
        o2 = p->paired ? sc->s2[i * T + G] : 0;
        if (p->paired) { /* map_dna.cpp:903-915 */
          if (o1 != 0 && o2 == 0) o2 = o1;
          else if (o2 != 0 && o1 == 0) o1 = o2;
        }
        int min_nm_sum = 100000, count = 0;
        for (int g = 0; g < T; g++) {
          nm_sum[g] = 10000;
          int s1 = g == G ? o1 : sc->s1[i * T + g];
          int s2 = p->paired ? (g == G ? o2 : sc->s2[i * T + g]) : 0;
          if (p->paired) {
            if (s1 != 0 && s2 != 0 && s1 != 10000 && s2 != 10000) {
              int v_max_r1 = (int)(len1 * p->tolerance);
              int v_max_r2 = (int)(len2 * p->tolerance);
              if ((s1 - 1) <= v_max_r1 && (s2 - 1) <= v_max_r2) {
                vmask |= 1u << g;
                if (min_nm_sum > s1 + s2) {
                  min_nm_sum = s1 + s2;
                  count++;
                }
                nm_sum[g] = s1 + s2;
              }
            }
          } else {
            if (s1 != 0 && s1 != 10000) {
              int v_max_r1 = (int)(len1 * p->tolerance);
              if ((s1 - 1) <= v_max_r1) {
                vmask |= 1u << g;
                if (min_nm_sum > s1) {
                  min_nm_sum = s1;
                  count++;
                }
                nm_sum[g] = s1;
              }
            }
          }
        }
        if (count >= 1) {
          for (int g = 0; g < T; g++)
            if (nm_sum[g] == min_nm_sum) tmask |= 1u << g;
          minsum = min_nm_sum;
        }
      } else {
        int nm_sum[HLM_MAX_LOCI + 1];
        int invalid_read = 1000, count = 0, min_score = invalid_read;
        for (int g = 0; g < T; g++) {
          nm_sum[g] = invalid_read;
          if (sc->s1[i * T + g] != 0xFFFF) nm_sum[g] = sc->s1[i * T + g];
          if (nm_sum[g] != invalid_read) {
            vmask |= 1u << g;
            count++;
          }
          if (nm_sum[g] < min_score) min_score = nm_sum[g];
        }
        if (count >= 1) {
          for (int g = 0; g < T; g++) {
            if (g != G && nm_sum[g] == min_score) {
              float max = (len1 * 2) * p->tolerance; /* map_rna.cpp:975 */
              if (min_score <= max) tmask |= 1u << g;
            }
            if (g == G && nm_sum[g] == min_score) {
              if (min_score < 30) tmask |= 1u << g; /* min_score_others, map_rna.cpp:939 */
            }
          }
          minsum = min_score;
        }
      }
    }
    out->target_mask[i] = tmask;
    if (out->visible_mask) out->visible_mask[i] = vmask;
    if (out->min_sum) out->min_sum[i] = (uint16_t)minsum;
    if (out->others_s1) out->others_s1[i] = (uint16_t)o1;
    if (out->others_s2) out->others_s2[i] = (uint16_t)o2;
  }
  return 0;
}

/* ------------------------------------------------------------------ SAM text reducers */
/* Literal restatements of read_sam_dna (src/map_dna.cpp:44-118) and read_sam_rna
 * (src/map_rna.cpp:39-101) for ONE record, used to pin the CIGAR / NM quirks against the
 * compiled reference.  Return the value that would be written to sequence_list_r*
 * (dna: mm+1; rna: the addend), or -1 when the record is skipped. */
static int split_tabs(const char* line, const char** f, int* fl, int maxf) {
  int nf = 0;
  const char* s = line;
  while (nf < maxf) {
    const char* e = strchr(s, '\t');
    f[nf] = s;
    fl[nf] = e ? (int)(e - s) : (int)strlen(s);
    nf++;
    if (!e) break;
    s = e + 1;
  }
  return nf;
}
static void cigar_clips(const char* cg, int len, int* lead, int* trail) {
  /* lead = length of the first op when it is S/H (what cigardata[0] yields);
   * trail = length of the last op when it is S/H and is not also the first op. */
  *lead = 0;
  *trail = 0;
  int i = 0, nops = 0, lastnum = 0;
  char lastop = 0;
  while (i < len) {
    int num = 0;
    while (i < len && isdigit((unsigned char)cg[i])) num = num * 10 + (cg[i++] - '0');
    if (i >= len) break;
    char op = cg[i++];
    if (nops == 0 && (op == 'S' || op == 'H')) *lead = num;
    lastnum = num;
    lastop = op;
    nops++;
  }
  if (nops >= 2 && (lastop == 'S' || lastop == 'H')) *trail = lastnum;
}
int orc_sam_score_dna(const char* line, int clip_mode, char* qname, size_t qlen) {
  const char* f[16];
  int fl[16];
  if (line[0] == '@') return -1;
  int nf = split_tabs(line, f, fl, 16);
  if (nf < 13) return -1;
  if ((fl[2] == 1 && f[2][0] == '*') || (fl[5] == 1 && f[5][0] == '*')) return -1;
  /* sam_line[12] split on ':' -> [2] */
  const char* c1 = memchr(f[12], ':', fl[12]);
  const char* c2 = c1 ? memchr(c1 + 1, ':', fl[12] - (int)(c1 + 1 - f[12])) : NULL;
  if (!c2) return -1;
  int mm = atoi(c2 + 1);
  int lead = 0, trail = 0;
  /* splitcigar() appends ',' after every op and boost::split keeps the empty token after
   * the final ',': cigardata.back() is "" for every well-formed CIGAR, so the reference
   * only ever adds the FIRST op (map_dna.cpp:60-78).  CLIP_BOTH adds the last one too. */
  cigar_clips(f[5], fl[5], &lead, &trail);
  if (mm != 10000) mm += lead;
  if (clip_mode == HLM_CLIP_BOTH && mm != 10000) mm += trail;
  int ql = fl[0];
  if (ql >= 2 && f[0][ql - 2] == '/' && (f[0][ql - 1] == '1' || f[0][ql - 1] == '2')) ql -= 2;
  if (qname && qlen) {
    int c = ql < (int)qlen - 1 ? ql : (int)qlen - 1;
    memcpy(qname, f[0], c);
    qname[c] = 0;
  }
  return mm + 1;
}
int orc_sam_score_rna(const char* line, int clip_mode, char* qname, size_t qlen) {
  const char* f[16];
  int fl[16];
  if (line[0] == '@') return -1;
  int nf = split_tabs(line, f, fl, 16);
  if (nf < 10) return -1;
  int ql = 0;
  while (ql < fl[0] && f[0][ql] != '$') ql++;
  if (qname && qlen) {
    int c = ql < (int)qlen - 1 ? ql : (int)qlen - 1;
    memcpy(qname, f[0], c);
    qname[c] = 0;
  }
  int seqlen = fl[9];
  if (fl[2] == 1 && f[2][0] == '*') return seqlen;
  if (nf < 13) return -1;
  const char* c1 = memchr(f[12], ':', fl[12]);
  const char* c2 = c1 ? memchr(c1 + 1, ':', fl[12] - (int)(c1 + 1 - f[12])) : NULL;
  if (!c2) return -1;
  int mm = atoi(c2 + 1);
  int lead = 0, trail = 0;
  cigar_clips(f[5], fl[5], &lead, &trail);
  if (lead > 0 && mm != seqlen) mm += lead;
  if (clip_mode == HLM_CLIP_BOTH && trail > 0 && mm != seqlen) mm += trail;
  return mm;
}

/* ------------------------------------------------------------------ single-FASTA aligner */
/* What the stand-in `bwa samse` (oracle/standin_bwa.c) calls so that the UNMODIFIED
 * reference binary can be driven end to end: one multi-FASTA = one target set. */
struct orc_ref {
  olocus L;
};
/* index cache next to the FASTA (what `bwa index` files are to the real bwa: the reference's
 * database ships prebuilt aligner indexes, so rebuilding ours per process would not be a fair
 * stand-in).  Layout: magic, n_al, n_seeds, lens[n_al], codes (concatenated), seeds[n_seeds]. */
#include <sys/stat.h>
#include <unistd.h>
static int cache_load(olocus* L, const char* fasta) {
  char path[4200];
  snprintf(path, sizeof path, "%s.orcidx", fasta);
  struct stat sf, sc;
  if (stat(fasta, &sf) != 0 || stat(path, &sc) != 0 || sc.st_mtime < sf.st_mtime) return -1;
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  int64_t hdr[3];
  if (fread(hdr, 8, 3, f) != 3 || hdr[0] != 0x4F52434944583031LL) {
    fclose(f);
    return -1;
  }
  L->n_al = hdr[1];
  L->n_seeds = hdr[2];
  int32_t* lens = (int32_t*)malloc((size_t)(L->n_al + 1) * 4);
  int ok = fread(lens, 4, (size_t)L->n_al, f) == (size_t)L->n_al;
  int64_t tot = 0;
  for (int64_t a = 0; ok && a < L->n_al; a++) tot += lens[a];
  uint8_t* codes = (uint8_t*)malloc((size_t)tot + 8);
  ok = ok && fread(codes, 1, (size_t)tot, f) == (size_t)tot;
  L->seeds = (oseed*)malloc((size_t)(L->n_seeds + 1) * sizeof(oseed));
  ok = ok && fread(L->seeds, sizeof(oseed), (size_t)L->n_seeds, f) == (size_t)L->n_seeds;
  fclose(f);
  if (!ok) {
    free(lens);
    free(codes);
    free(L->seeds);
    memset(L, 0, sizeof *L);
    return -1;
  }
  L->al = (oseq*)malloc((size_t)(L->n_al + 1) * sizeof(oseq));
  int64_t o = 0;
  for (int64_t a = 0; a < L->n_al; a++) {
    L->al[a].len = lens[a];
    L->al[a].codes = (uint8_t*)malloc((size_t)lens[a] + 8); /* orc_ref_close frees per allele */
    memcpy(L->al[a].codes, codes + o, (size_t)lens[a]);
    o += lens[a];
  }
  free(lens);
  free(codes);
  return 0;
}
static void cache_store(const olocus* L, const char* fasta) {
  char path[4200], tmp[4300];
  snprintf(path, sizeof path, "%s.orcidx", fasta);
  snprintf(tmp, sizeof tmp, "%s.tmp%d", path, (int)getpid());
  FILE* f = fopen(tmp, "wb");
  if (!f) return; /* read-only database: just no cache */
  int64_t hdr[3] = {0x4F52434944583031LL, L->n_al, L->n_seeds};
  int ok = fwrite(hdr, 8, 3, f) == 3;
  for (int64_t a = 0; ok && a < L->n_al; a++) {
    int32_t l = L->al[a].len;
    ok = fwrite(&l, 4, 1, f) == 1;
  }
  for (int64_t a = 0; ok && a < L->n_al; a++)
    ok = fwrite(L->al[a].codes, 1, (size_t)L->al[a].len, f) == (size_t)L->al[a].len;
  ok = ok && fwrite(L->seeds, sizeof(oseed), (size_t)L->n_seeds, f) == (size_t)L->n_seeds;
  fclose(f);
  if (ok) rename(tmp, path);
  else remove(tmp);
}
orc_ref* orc_ref_open(const char* fasta) {
  orc_ref* r = (orc_ref*)calloc(1, sizeof(orc_ref));
  /* the stand-in bwa only sees a FASTA path: the reference keeps the rna target sets under
   * mapper/rna/ and others/rna/ (map_rna.cpp:833, :858) */
  int rna = strstr(fasta, "/mapper/rna/") != NULL || strstr(fasta, "/others/rna/") != NULL;
  if (cache_load(&r->L, fasta) != 0) {
    if (load_fasta(&r->L, fasta) != 0) {
      free(r);
      return NULL;
    }
    cache_store(&r->L, fasta);
  }
  r->L.rna = rna;
  return r;
}
void orc_ref_close(orc_ref* r) {
  if (!r) return;
  for (int64_t a = 0; a < r->L.n_al; a++) free(r->L.al[a].codes);
  free(r->L.al);
  free(r->L.seeds);
  free(r);
}
int orc_ref_align(const orc_ref* r, const uint8_t* ascii, int len, int mode, int mm_max, int* S,
                  int* cost, int* lead, int* trail) {
  if (len > HLM_MAX_READ) return 0;
  uint8_t codes[HLM_MAX_READ + 8];
  scratch sc;
  memset(&sc, 0, sizeof sc);
  int64_t cells = 0;
  to_codes(ascii, len, codes);
  ares a = mode ? score_fast(&r->L, codes, len, &sc, mm_max, &cells)
                : score_exhaustive(&r->L, codes, len, &sc, &cells);
  free(sc.buf);
  free(sc.fh);
  free(sc.fidx);
  free(sc.lb);
  if (!a.valid || a.cost > mm_max) return 0;
  *S = a.S;
  *cost = a.cost;
  *lead = a.lead;
  *trail = a.trail;
  return 1;
}

/* many reads at once on nthreads host threads (what `bwa aln -t T` does); out[4*i..] = S, cost,
 * lead, trail and ok[i] = 1 when an alignment with cost <= mm_max exists */
void orc_ref_align_batch(const orc_ref* r, const uint8_t* const* reads, const int* lens, int64_t n,
                         int mode, int mm_max, int nthreads, int32_t* out, uint8_t* ok) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    scratch sc;
    memset(&sc, 0, sizeof sc);
    uint8_t codes[HLM_MAX_READ + 8];
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; i++) {
      ok[i] = 0;
      if (lens[i] > HLM_MAX_READ) continue;
      int64_t cells = 0;
      to_codes(reads[i], lens[i], codes);
      ares a = mode ? score_fast(&r->L, codes, lens[i], &sc, mm_max, &cells)
                    : score_exhaustive(&r->L, codes, lens[i], &sc, &cells);
      if (!a.valid || a.cost > mm_max) continue;
      ok[i] = 1;
      out[4 * i] = a.S;
      out[4 * i + 1] = a.cost;
      out[4 * i + 2] = a.lead;
      out[4 * i + 3] = a.trail;
    }
    free(sc.buf);
    free(sc.fh);
    free(sc.fidx);
    free(sc.lb);
  }
}
