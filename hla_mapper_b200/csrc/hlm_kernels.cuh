// hlm_kernels.cuh — launch interface of the sm_100a kernels (hlm_kernels.cu).
#ifndef HLM_KERNELS_CUH
#define HLM_KERNELS_CUH

#include <cuda_runtime.h>

#include "hlm_common.h"

// per-chunk device state
struct HlmChunk {
  // inputs (device)
  int64_t n_pairs;
  const uint8_t *bases1, *quals1, *bases2, *quals2;
  const uint64_t *offs1, *offs2;
  const int32_t* size_override1;
  const uint16_t *rna_split1, *rna_split2;
  // rna block list (optional): offsets are absolute, `blk` is biased so that they index it
  // directly; block b is unit n_pairs * units_per_pair + (b - blk_first)
  const uint64_t* blk_off;
  const hlm_rna_block* blk;
  int64_t blk_first, n_blocks;
  // screen outputs
  uint8_t* flags;
  uint16_t *trim_lo1, *trim_len1, *trim_lo2, *trim_len2;
  uint32_t* locus_mask;
  // units: units_per_pair * n_pairs (+ n_blocks); HLM_UNIT_STRIDE words each
  int units_per_pair;
  int64_t n_units;
  uint32_t* unit_planes;
  uint16_t* unit_len;    // 0 = inactive
  uint32_t* unit_mask;   // targets this unit is scored against
  uint16_t* unit_cap;    // capped scoring: int(len * tolerance) of the unit's mate (map_dna.cpp:948-958)
  // scoring work
  uint64_t* cand;        // candidate records
  int64_t cand_cap;
  uint32_t* dp_list;     // indices into cand, in selection order
  uint32_t* dp_sorted;   // ... counting-sorted by DP rows (descending)
  uint8_t* dp_len;       // DP rows of dp_list[i]
  uint32_t* dp_hist;     // [256] rows histogram, then ...
  uint32_t* dp_cursor;   // [256] ... scatter cursors
  int64_t dp_cap;
  uint32_t* pre_list;            // k_prescreen: pairs the Bloom-only prefilter could not reject
  unsigned long long* dp_tab;    // footprint de-duplication of the DP list: 2^dp_tab_bits u64 slots
  int dp_tab_bits;
  unsigned long long* counters;  // [0] n_cand, [1] n_dp (current round), [2] overflow flag,
                                 // [3] n_kept, [4] dp cells, [5] myers rows, [6] n_dp total,
                                 // [7] read too long, [8] n_cand after the seed index,
                                 // [9] / [10] after expansion round 0 / 1 (slots incl. padding),
                                 // [11] real candidates, [12] real seed-index candidates,
                                 // [13] DP list entries after de-duplication, [14] / [15] work counters
  uint32_t* task_emin;   // units * n_targets: minimum lower bound over evaluated candidates
  uint32_t* task_emin0;  // ... over the seed-index candidates only (expansion gate)
  uint32_t* task_pmin;   // capped scoring: minimum prefix bound (rows the trailing clip cannot reach)
  unsigned long long* task_best;
  // per-allele scoring (hlm_score_alleles; null / 0 otherwise)
  unsigned long long* cand_res;     // per candidate: packed alignment (HLM_BEST_PACK) or HLM_BEST_NONE
  unsigned long long* allele_best;  // units * n_al: best packed alignment of the unit vs each allele
  int n_al;
  // outputs
  uint16_t *s1, *s2;
  int16_t *aln1, *aln2;
  uint8_t *nm1, *nm2, *lead1, *lead2, *trail1, *trail2;
  uint32_t *target_mask, *visible_mask;
  uint16_t *min_sum, *others_s1, *others_s2;
};

#define HLM_N_COUNTERS 16

struct HlmKParams {
  int rna, paired, screen_variant, v_size, minhit, mm_max, clip_mode, skip_screen, score_cap;
  int allele_locus;  // >= 0: per-allele scoring of this target set (hlm_score_alleles); -1 otherwise
  double tolerance;
  uint32_t good_q[8];  // bit q: quality byte q passes mtrim's P <= error test
  int q_thr;           // >= 0: the good bytes are exactly q_thr .. 127 (byte-parallel test); -1: use good_q
  int n_loci;
  int one;  // == 1; opaque multiplier that keeps the DP's adds on the FMA pipe (hlm_add_fma)
};

void hlm_launch_screen(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int sms,
                       cudaStream_t st);
void hlm_launch_pack(const HlmKParams& kp, const HlmChunk& ck, int sms, cudaStream_t st);
void hlm_launch_clear_tasks(const HlmChunk& ck, int n_targets, int sms, cudaStream_t st);
void hlm_launch_seed(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int sms,
                     cudaStream_t st);
void hlm_launch_myers(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                      int phase, int sms, cudaStream_t st);
void hlm_launch_snapshot(const HlmChunk& ck, int which, cudaStream_t st);
void hlm_launch_expand(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                       int round, int sms, cudaStream_t st);
void hlm_launch_select(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                       int round, int sms, cudaStream_t st);
void hlm_launch_dp(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                   int sms, cudaStream_t st);
void hlm_launch_allele_gather(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, uint8_t* nm1,
                              uint8_t* clip1, uint8_t* nm2, uint8_t* clip2, int sms, cudaStream_t st);
void hlm_launch_scores(const HlmKParams& kp, const HlmChunk& ck, int n_targets, int sms,
                       cudaStream_t st);
void hlm_launch_assign(const HlmKParams& kp, const HlmChunk& ck, int n_targets, int sms,
                       cudaStream_t st);
void hlm_kernels_init();
double hlm_int_peak_run(int which, int sms, cudaStream_t st);
void hlm_launch_dpab(int which, const uint32_t* reads, const uint32_t* tiles, const int* offn, int m,
                     int4* out, int sms, cudaStream_t st);

#endif
