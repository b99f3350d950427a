/* synth_reads.c - fast, counter-based generator of synthetic read pairs (bench inputs only).
 *
 * Same read model as hla_mapper_b200/synth.py:_block: fragments N(mu, sd)
 * drawn from a diploid sample of the database alleles (+ flanks), R2 reverse-complemented, the
 * orientation of the pair flipped at random, substitution errors ramping 0.2 % -> 1 % along
 * the read, indel errors (the read stays read_len long, as a sequencer's does: a deletion
 * consumes one more template base, an insertion one fewer), a few N calls, four-bin qualities
 * with low-quality tails / heads, and off-target decoys (uniform random, 10 % low complexity).
 *
 * Every pair has its own random stream keyed by (seed, global pair index), so any shard of a
 * 100 M-pair run is reproduced independently of the others (config 4) and the output does not
 * depend on the thread count.  Not part of the product path and not an oracle: it only makes
 * inputs.  Built by __graft_entry__.build() into tools/libsynthreads.so, bound in synth.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  uint64_t s;
} rng_t;
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t next(rng_t* r) { return mix64(r->s += 0x9E3779B97F4A7C15ull); }
static inline double unif(rng_t* r) { return (double)(next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t below(rng_t* r, uint32_t n) { return (uint32_t)(((next(r) >> 32) * (uint64_t)n) >> 32); }

static const char ACGT[4] = {'A', 'C', 'G', 'T'};
static const char QBIN[4] = {'?', '5', '+', '\''};

/* one mate: template bases tpl[0], tpl[step], ... (step = +1 forward, -1 for the reverse
 * strand, complemented), errors applied while copying */
static void emit_mate(rng_t* r, const uint8_t* flat, int64_t flat_len, int64_t pos, int step, int L,
                      double indel_rate, double lowq_tail, uint8_t* b, uint8_t* q) {
  uint32_t indel_thr = (uint32_t)(indel_rate * 4294967296.0);
  for (int i = 0; i < L; i++) {
    uint64_t x = next(r);
    uint32_t e24 = (uint32_t)(x & 0xFFFFFF), n20 = (uint32_t)((x >> 24) & 0xFFFFF), q16 = (uint32_t)(x >> 48);
    uint32_t id = (uint32_t)(next(r) >> 32);
    int code;
    if (id < indel_thr) {
      if (id & 1) {  /* insertion: a random base, the template does not advance */
        code = (int)((id >> 1) & 3);
        goto have;
      }
      pos += step;   /* deletion: skip one template base */
    }
    {
      int64_t p = pos < 0 ? 0 : (pos >= flat_len ? flat_len - 1 : pos);
      code = flat[p];
      if (step < 0) code = 3 - code;
      pos += step;
    }
  have:;
    /* substitution error: 0.002 -> 0.01 along the read */
    double perr = 0.002 + (0.01 - 0.002) * (L > 1 ? (double)i / (double)(L - 1) : 0.0);
    if (e24 < (uint32_t)(perr * 16777216.0)) code = (code + 1 + (int)(e24 % 3)) & 3;
    b[i] = (uint8_t)ACGT[code];
    if (n20 < 524) b[i] = 'N'; /* 0.0005 */
    /* qualities: 95.5 / 4 / 0.4 / 0.1 % */
    q[i] = (uint8_t)QBIN[q16 < 62587 ? 0 : (q16 < 65208 ? 1 : (q16 < 65470 ? 2 : 3))];
  }
  /* low-quality calls cluster in tails / heads (this is what mtrim cuts) */
  if (unif(r) < lowq_tail && L >= 4) {
    int tl = 1 + (int)below(r, (uint32_t)(L / 2 - 1 > 0 ? L / 2 - 1 : 1));
    memset(q + L - tl, '\'', (size_t)tl);
  }
  if (unif(r) < lowq_tail / 4 && L >= 8) {
    int hl = 1 + (int)below(r, (uint32_t)(L / 4 - 1 > 0 ? L / 4 - 1 : 1));
    memset(q, '+', (size_t)hl);
  }
}

static void emit_decoy(rng_t* r, int L, int lowc, double lowq_tail, uint8_t* b, uint8_t* q) {
  int unit[3] = {(int)below(r, 4), (int)below(r, 4), (int)below(r, 4)};
  for (int i = 0; i < L; i++) {
    uint64_t x = next(r);
    uint32_t q16 = (uint32_t)(x >> 48);
    int code = lowc ? unit[i % 3] : (int)(x & 3);
    b[i] = (uint8_t)ACGT[code];
    if (((x >> 24) & 0xFFFFF) < 524) b[i] = 'N';
    q[i] = (uint8_t)QBIN[q16 < 62587 ? 0 : (q16 < 65208 ? 1 : (q16 < 65470 ? 2 : 3))];
  }
  if (unif(r) < lowq_tail && L >= 4) {
    int tl = 1 + (int)below(r, (uint32_t)(L / 2 - 1 > 0 ? L / 2 - 1 : 1));
    memset(q + L - tl, '\'', (size_t)tl);
  }
}

/* flat / starts / lens / locus: the diploid sample (hla_mapper_b200/synth.py:_Source).
 * Outputs are uniform-length: pair i occupies [i * L, (i + 1) * L) of every array. */
int synth_reads(const uint8_t* flat, int64_t flat_len, const int64_t* starts, const int64_t* lens,
                const int32_t* locus, int n_src, uint64_t seed, int64_t first_pair, int64_t n_pairs,
                int L, double on_target, double frag_mu, double frag_sd, double lowq_tail,
                double indel_rate, uint8_t* b1, uint8_t* q1, uint8_t* b2, uint8_t* q2, int32_t* truth,
                int threads) {
  if (n_src <= 0 || L <= 0) return -1;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_pairs; i++) {
    rng_t r;
    r.s = mix64(seed * 0xD6E8FEB86659FD93ull + 0x1234567ull) ^ mix64((uint64_t)(first_pair + i) + 0x51ED27ull);
    uint8_t *pb1 = b1 + i * L, *pq1 = q1 + i * L, *pb2 = b2 + i * L, *pq2 = q2 + i * L;
    if (unif(&r) >= on_target) {
      truth[i] = -1;
      emit_decoy(&r, L, unif(&r) < 0.10, lowq_tail, pb1, pq1);
      emit_decoy(&r, L, 0, lowq_tail, pb2, pq2);
      continue;
    }
    int h = (int)below(&r, (uint32_t)n_src);
    /* Box-Muller */
    double u1 = unif(&r), u2 = unif(&r);
    if (u1 < 1e-300) u1 = 1e-300;
    double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    int64_t frag = (int64_t)floor(frag_mu + frag_sd * z + 0.5);
    if (frag < L) frag = L;
    if (frag > lens[h]) frag = lens[h];
    int64_t st = (int64_t)(unif(&r) * (double)(lens[h] - frag + 1));
    int flip = unif(&r) < 0.5;
    int64_t p_fwd = starts[h] + st, p_rev = starts[h] + st + frag - 1;
    truth[i] = locus[h];
    /* R1 reads the fragment forwards, R2 its reverse complement from the far end; a flipped pair
     * swaps the roles */
    emit_mate(&r, flat, flat_len, flip ? p_rev : p_fwd, flip ? -1 : 1, L, indel_rate, lowq_tail, pb1, pq1);
    emit_mate(&r, flat, flat_len, flip ? p_fwd : p_rev, flip ? 1 : -1, L, indel_rate, lowq_tail, pb2, pq2);
  }
  return 0;
}
