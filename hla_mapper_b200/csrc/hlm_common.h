// hlm_common.h — formats shared by the host index builder and the sm_100a kernels.
//
// Base codes: A=0 C=1 G=2 T=3 (complement = 3-c); everything else is N and never matches.
//
// k-mer screen index   Bloom front (kb: 4 bytes per key, L2-resident) + open-addressing table of u64 slots:
//                        bit 63 = occupied, bits 40..55 = id into mask_tab[], bits 0..39 = 20-mer
//                        (base i of the k-mer at bits 2i..2i+1).  mask_tab[id] = membership
//                        mask: bit g = motif/<kind>/loci/<g>.txt, bit 31 = motif_all.txt.
// tiles                the allele set of a locus is cut into 256-column windows at stride 32
//                        over the N-padded allele (HLM_PAD N's on both sides); windows with
//                        identical content are stored once per locus.  A tile is 24 words:
//                        p0[8] (low code bit), p1[8] (high code bit), nm[8] (1 = N), column c at
//                        bit (c & 31) of word (c >> 5).
// tile hierarchy       a tile that differs from an earlier tile of the same locus and window
//                        index in <= HLM_KMAX columns is a "child" of that "rep": tile_diff =
//                        ndiff << 24 | differing columns; children of rep t are
//                        child_ent[child_off[t] .. child_off[t+1]).
// seed index           CSR over all 4^12 12-mers: seed_off[code] .. seed_off[code+1] index
//                        seed_ent[], entries (column << 24 | tile_id) sorted by (column, tile).
//                        Reps list every seed; children only the seeds covering a differing
//                        column (their other candidates come from expanding the rep).
// packed reads         a "unit" (trimmed mate, or an RNA block of it) is 18 words per strand: r0[6],
//                        r1[6], rn[6], base i at bit (i & 31) of word (i >> 5); bits >= len are 0
//                        in r0/r1 and 1 in rn; plus 24 words per strand of 4-bit codes (code bit 0
//                        / 1 = r0 / r1, bit 2 = N), eight bases per word, for the Myers kernel.
#ifndef HLM_COMMON_H
#define HLM_COMMON_H

#include <stdint.h>

#include "../../include/hlamap.h"

#define HLM_TILE_COLS 256
#define HLM_TILE_STRIDE 32
#define HLM_TILE_WORDS 24
#define HLM_PAD 256                 /* N padding on both sides of every allele            */
#define HLM_OFF_MIN HLM_BAND        /* candidate column of read base 0: [OFF_MIN, OFF_MIN+32) */
#define HLM_NB (2 * HLM_BAND + 1)   /* band cells per DP row                               */
#define HLM_UNIT_WORDS 18
#define HLM_CODE_WORDS 24            /* 4-bit read codes, 8 per word, per strand               */
#define HLM_UNIT_STRIDE (2 * HLM_UNIT_WORDS + 2 * HLM_CODE_WORDS) /* planes fwd, planes rc, codes fwd, codes rc */
#define HLM_READ_WORDS 6
#define HLM_RAW_MAX 256             /* longest raw read line the screen kernel accepts     */
#define HLM_SEED_CODES (1u << (2 * HLM_SEED))
#define HLM_KMAX 3                  /* a child tile differs from its rep in <= KMAX columns  */
#define HLM_DIFF_N(d) ((int)((d) >> 24))
#define HLM_DIFF_COL(d, i) ((int)(((d) >> (8 * (i))) & 0xFFu))

#define HLM_KT_OCC (1ull << 63)
#define HLM_KT_KEYMASK ((1ull << 40) - 1)

/* composite DP value V = S*SC - cost*CU - lead (lexicographic: max S, min cost, min lead) */
#define HLM_SC (1 << 20)
#define HLM_CU (1 << 9)
#define HLM_V_MATCH (HLM_SC)
#define HLM_V_MISM (-4 * HLM_SC - HLM_CU)
#define HLM_V_GAPO (-(7 * HLM_SC + HLM_CU))
#define HLM_V_GAPE (-(HLM_SC + HLM_CU))
#define HLM_V_CLIP (5 * HLM_SC)
#define HLM_V_NEG (-(1 << 30))

#ifdef __CUDACC__
#define HLM_COMMON_HD __host__ __device__ __forceinline__
#else
#define HLM_COMMON_HD static inline
#endif
/* Read seeds: 12-mers at read offsets 0, st, 2 st, ... <= n - 12.  st = 12
 * (non-overlapping), except for the units of an `rna` run shorter than HLM_SEED_SHORT bases - the
 * blocks STAR cuts a mate into - which use overlapping seeds, st = 6, so that an error-free 12-mer
 * is found wherever it lies to within 6 bases (at most 15 seeds per strand either way). */
HLM_COMMON_HD int hlm_seed_stride(int n, int rna) {
  return (rna && n < HLM_SEED_SHORT) ? HLM_SEED_STRIDE_SHORT : HLM_SEED;
}
HLM_COMMON_HD int hlm_seed_count(int n, int rna) {
  return n < HLM_SEED ? 0 : (n - HLM_SEED) / hlm_seed_stride(n, rna) + 1;
}

HLM_COMMON_HD uint64_t hlm_kt_hash(uint64_t key, int bits) {
  return (key * 0x9E3779B97F4A7C15ull) >> (64 - bits);
}

/* Bloom front of the k-mer table: one 64-bit word per probe, four bits per key.  The word comes
 * from the high half of the multiplicative hash, the bit positions from its low half.  A key
 * whose four bits are not all set is in no motif file; the others go on to the exact table. */
HLM_COMMON_HD uint64_t hlm_kb_probe(uint64_t key, uint32_t n_words, uint32_t* word) {
  uint64_t h = key * 0x9E3779B97F4A7C15ull;
  uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h;
  *word = (uint32_t)(((uint64_t)hi * n_words) >> 32);
  return (1ull << (lo & 63)) | (1ull << ((lo >> 6) & 63)) | (1ull << ((lo >> 12) & 63)) |
         (1ull << ((lo >> 18) & 63));
}

/* device view of the database (all pointers are device pointers) */
struct HlmDevDB {
  const uint64_t* kt;        /* k-mer table                                  */
  const uint64_t* kb;        /* its Bloom front: kb_words 64-bit words       */
  uint32_t kb_words;
  const uint32_t* mask_tab;  /* mask id -> membership mask                   */
  const uint8_t* odd_keys;   /* motif lines with a non-ACGT byte: 20 raw bytes each, sorted */
  const uint32_t* odd_masks;
  int n_odd;
  int kt_bits;
  const uint32_t* tiles;     /* n_tiles * 24 words                           */
  const uint32_t* seed_off;  /* 4^12 + 1                                     */
  const uint32_t* seed_ent;
  const uint32_t* tile_diff; /* per tile: ndiff << 24 | diff columns (0 = rep)        */
  const uint32_t* child_off; /* n_tiles + 1                                           */
  const uint32_t* child_ent; /* child tile ids                                        */
  const uint32_t* child_diff;
  const uint32_t* tile_al_off; /* n_tiles + 1: alleles (index inside the locus) that hold tile t = */
  const uint32_t* tile_al_ent; /*   tile_al_ent[tile_al_off[t] .. tile_al_off[t+1])  (hlm_score_alleles) */
  uint32_t tile_start[HLM_MAX_LOCI + 3]; /* first tile id of locus g; [n_loci+1] = n_tiles */
  int n_loci;
};

/* candidate record (u64): unit (26) | tile (24) | off-OFF_MIN (5) | strand (1) | lb (8) */
#define HLM_CAND(unit, tile, offm, strand, lb)                                              \
  (((uint64_t)(unit) << 38) | ((uint64_t)(tile) << 14) | ((uint64_t)(offm) << 9) |           \
   ((uint64_t)(strand) << 8) | (uint64_t)(lb))
#define HLM_CAND_UNIT(c) ((uint32_t)((c) >> 38))
#define HLM_CAND_TILE(c) ((uint32_t)(((c) >> 14) & 0xFFFFFF))
#define HLM_CAND_OFFM(c) ((uint32_t)(((c) >> 9) & 31))
#define HLM_CAND_STRAND(c) ((uint32_t)(((c) >> 8) & 1))
#define HLM_CAND_LB(c) ((uint32_t)((c) & 0xFF))

/* per-task best alignment, packed so that a smaller u64 is a better alignment:
 * cost (16) | 4096 - S (16) | lead (16) | trail (16) */
#define HLM_BEST_NONE 0xFFFFFFFFFFFFFFFFull
#define HLM_BEST_PACK(cost, S, lead, trail)                                                  \
  (((uint64_t)(cost) << 48) | ((uint64_t)(4096 - (S)) << 32) | ((uint64_t)(lead) << 16) |    \
   (uint64_t)(trail))
#define HLM_BEST_COST(b) ((int)((b) >> 48))
#define HLM_BEST_S(b) (4096 - (int)(((b) >> 32) & 0xFFFF))
#define HLM_BEST_LEAD(b) ((int)(((b) >> 16) & 0xFFFF))
#define HLM_BEST_TRAIL(b) ((int)((b) & 0xFFFF))

#endif
