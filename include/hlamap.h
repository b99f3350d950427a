/*
 * hlamap.h — C-ABI of libhlamap, the B200-native hot path of hla-mapper.
 *
 * The reference (erickcastelli/hla-mapper v4.5.1) has no library / FFI surface:
 * its hot path is a chain of file-coupled stages inside main_preselect() and
 * main_dna_map() / main_rna_map().  Each entry point below replaces one of those
 * stages and cites the reference lines it stands in for.  Everything is plain C:
 * pointers + sizes, caller-owned buffers, int return codes (0 = ok, <0 = error,
 * text via hlm_last_error()).  No exceptions, no global mutable state.
 *
 *   stage                      reference                                  entry point
 *   -------------------------  -----------------------------------------  -------------------
 *   DB layout reader           src/map_dna.cpp:299-365, map_rna.cpp:281-335,
 *                              preselect.cpp:297-309, map_dna.cpp:565-581  hlm_db_open
 *   global k-mer screen        src/preselect.cpp:707-731 (FASTQ),
 *                              :484-506 (BAM-derived)                      hlm_screen_trim
 *   quality trim (mtrim)       src/functions.cpp:89-145,
 *                              src/preselect.cpp:1060-1149                 hlm_screen_trim
 *   per-locus k-mer sort       src/map_dna.cpp:611-627, map_rna.cpp:544-560 hlm_screen_trim
 *   aligner + SAM reducer      src/map_dna.cpp:749-776 (bwa aln/samse),
 *                              :44-118 (read_sam_dna), map_rna.cpp:39-101  hlm_score
 *   others back-fill, compare, src/map_dna.cpp:903-1029,
 *   argmin, tie set            src/map_rna.cpp:925-1027                    hlm_assign
 *   ThreadPool dispatch        src/ThreadPool.hpp:18-100                   hlm_run (stream pipeline)
 *   FASTQ(.gz) loops, writers  src/preselect.cpp:652-1203,
 *                              src/map_dna.cpp:1036-1351                   hlm_map_fastq
 *   bam= input (4 x samtools)  src/preselect.cpp:315-631                   hlm_bam_extract, hlm_map_reads
 *   rna: STAR SAM -> blocks    src/map_rna.cpp:740-790                     hlm_batch.rna_blocks (input)
 */
#ifndef HLAMAP_H
#define HLAMAP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HLM_ABI_VERSION 5

/* limits of this build */
#define HLM_MAX_LOCI 31   /* loci + "others" must fit a uint32 mask            */
#define HLM_MAX_READ 184  /* longest trimmed mate the 256-column tiles can hold */
#define HLM_MOTIF 20      /* motif_size, src/main.cpp:73                        */
#define HLM_SEED 12       /* aligner seed length (new)           */
#define HLM_SEED_SHORT 96 /* rna: units shorter than this use overlapping seeds */
#define HLM_SEED_STRIDE_SHORT 6
#define HLM_BAND 20       /* DP half band = v_mm_max, src/main.cpp:77           */

/* error codes */
#define HLM_OK 0
#define HLM_E_ARG -1
#define HLM_E_IO -2
#define HLM_E_DB -3
#define HLM_E_CUDA -4
#define HLM_E_NOMEM -5
#define HLM_E_TOOLONG -6
#define HLM_E_NODEVICE -7

/* screen variants: the reference has two differently-strided loops */
#define HLM_SCREEN_FASTQ 0 /* preselect.cpp:707-731: stride 2, 3 after a hit   */
#define HLM_SCREEN_BAM 1   /* preselect.cpp:484-506: stride 1, 2 after a hit   */

/* how clipped bases enter the hla-mapper score (read_sam_dna) */
#define HLM_CLIP_LEAD_ONLY 0 /* literal reference behaviour: the trailing token of
                                splitcigar() is always "", so only the first CIGAR
                                op is ever added (map_dna.cpp:60-78)            */
#define HLM_CLIP_BOTH 1      /* what the code visibly intends: first + last op  */
#define HLM_CLIP_NONE 2      /* the aligner itself never clips: the whole mate is aligned end to
                                end, as `bwa aln` reports reads that lie inside a sequence (it
                                has no soft clipping; trailing mismatches count in NM).  Read
                                bases past an allele end align to padding and count as
                                mismatches.  DESIGN.md section 3b                */

/* hlm_screen_out.flags bits */
#define HLM_F_SCREENED 1 /* passed the >=minhit k-mer screen                   */
#define HLM_F_KEPT 2     /* ... and both mates survived mtrim (pair is scored)  */

typedef struct hlm_db hlm_db;
typedef struct hlm_ctx hlm_ctx;
typedef struct hlm_devbatch hlm_devbatch;

typedef struct hlm_params {
  int32_t struct_size;    /* sizeof(hlm_params), for forward compatibility      */
  int32_t rna;            /* 0 = dna, 1 = rna (map_rna.cpp reducer + compare)    */
  int32_t paired;         /* 1 = R1+R2, 0 = single-end (r0=)                     */
  int32_t screen_variant; /* HLM_SCREEN_FASTQ | HLM_SCREEN_BAM                   */
  int32_t v_size;         /* min read length; <=0: take SIZE: from db_*.info     */
  int32_t minhit;         /* preselect.cpp:26 (3)                                */
  int32_t mm_max;         /* bwa aln -n (20), main.cpp:77                        */
  int32_t clip_mode;      /* HLM_CLIP_*                                          */
  double tolerance;       /* main.cpp:80 (0.05); rna forces 0.08 map_rna.cpp:116 */
  double mtrim_error;     /* main.cpp:81 ((double)0.08f)                         */
  int32_t chunk_pairs;    /* pairs per pipeline chunk (0 = default)              */
  int32_t skip_screen;    /* 1: treat every pair as screened (scoring-only runs) */
  int64_t cand_capacity;  /* device candidate slots per chunk (0 = default)      */
  int32_t single_slot;    /* 1: one pipeline slot, chunks strictly serialised (per-kernel
                             timing runs; the default three slots overlap copies and kernels) */
  int32_t score_cap;      /* 0 (default): every score up to mm_max is exact, as the reference holds
                             them in memory.  1: tolerance-capped scoring (dna only; rna ignores
                             it): scores the compare step cannot use are not computed - a
                             (pair, target) whose alignment fails (s - 1) <= int(len * tolerance)
                             for either mate (map_dna.cpp:948-958) may be reported as 0 instead of
                             its value.  target_mask, visible_mask, min_sum and every score of a
                             visible target - i.e. the addressing table and the per-gene files -
                             are identical to the exact mode            */
  int32_t no_dp_dedup;    /* 1: switch off the footprint de-duplication of the DP list (A/B and
                             test switch; results are identical either way)                   */
  int32_t reserved1;
} hlm_params;

/* One record of the reference's sorted_spliced_<gene>.fastq (map_rna.cpp:740-790): a block of a
 * trimmed mate that STAR, run on the reads sorted into `gene`, reported between two splice
 * junctions (or the whole mate when the alignment has none).  start / len are in the trimmed
 * mate's own orientation; a record STAR printed reverse-complemented covers
 * [n - start' - len, n - start') of it.  Blocks are clipped to the mate. */
typedef struct hlm_rna_block {
  uint8_t gene;      /* DB locus index (< n_loci; others is never spliced)               */
  uint8_t mate;      /* 0 = R1 / R0, 1 = R2                                              */
  uint16_t start;
  uint16_t len;
  uint16_t reserved; /* 0                                                                */
} hlm_rna_block;

/* SoA batch of read pairs.  bases/quals are the FASTQ lines as read (ASCII),
 * concatenated; offs has n_pairs+1 entries.  For single-end leave mate 2 NULL.
 * size_override1 (optional, single-end only) replaces len(trimmed R1) in the
 * tolerance test, reproducing preselect.cpp:1199 (length of the header line). */
typedef struct hlm_batch {
  int64_t n_pairs;
  const uint8_t* bases1;
  const uint8_t* quals1;
  const uint64_t* offs1;
  const uint8_t* bases2;
  const uint8_t* quals2;
  const uint64_t* offs2;
  const int32_t* size_override1;
  /* RNA only, optional: split point inside the trimmed mate (0 = one block).  Stands in
   * for the STAR splice pre-split of map_rna.cpp:648-801: a mate with a split is scored
   * as two blocks that read_sam_rna sums (map_rna.cpp:39-101). */
  const uint16_t* rna_split1;
  const uint16_t* rna_split2;
  /* RNA only, optional, replaces rna_split*: the STAR pre-split in full, one hlm_rna_block per
   * record the reference would write to sorted_spliced_<gene>.fastq - per gene, any number of
   * blocks per mate, one set per reported alignment (multi-mappers included).  Blocks of pair i:
   * rna_blocks[rna_block_off[i] .. rna_block_off[i+1]).  The loci columns of hlm_score_out.s1
   * are then exactly read_sam_rna's sums over these records (map_rna.cpp:39-101; a gene without
   * a record is 0xFFFF); `others` still sees the whole trimmed mates (map_rna.cpp:858-892). */
  const uint64_t* rna_block_off;
  const hlm_rna_block* rna_blocks;
} hlm_batch;

/* one entry per pair */
typedef struct hlm_screen_out {
  uint8_t* flags;       /* HLM_F_*                                              */
  uint16_t* trim_lo1;   /* mtrim substring start / length per mate               */
  uint16_t* trim_len1;
  uint16_t* trim_lo2;
  uint16_t* trim_len2;
  uint32_t* locus_mask; /* bit g: pair is a candidate of DB locus g              */
} hlm_screen_out;

/* n_pairs x n_targets (n_targets = n_loci + 1, last column = "others"),
 * row-major.  s = hla-mapper score (NM + clips + 1; 0 = no alignment);
 * aln = affine alignment score of the reported alignment (1/-4/-6/-1/clip -5);
 * nm / lead / trail = its edit count and clip lengths (reference orientation).
 * RNA mode: s1 holds the per-(pair,target) sum of map_rna.cpp:39-101 and
 * 0xFFFF marks "no record"; s2 is unused. */
typedef struct hlm_score_out {
  uint16_t* s1;
  uint16_t* s2;
  int16_t* aln1;
  int16_t* aln2;
  uint8_t* nm1;
  uint8_t* nm2;
  uint8_t* lead1;
  uint8_t* lead2;
  uint8_t* trail1;
  uint8_t* trail2;
} hlm_score_out;

/* one entry per pair */
typedef struct hlm_assign_out {
  uint32_t* target_mask;  /* bit g = addressed to locus g, bit n_loci = others   */
  uint32_t* visible_mask; /* targets whose score column is printed (not "-")     */
  uint16_t* min_sum;      /* min_nm_sum (dna) / min_score (rna); 0 = none        */
  uint16_t* others_s1;    /* "others" scores after the one-sided mate back-fill   */
  uint16_t* others_s2;    /* (map_dna.cpp:903-915); dna paired only               */
} hlm_assign_out;

/* per-stage device time of the last hlm_run*, CUDA events on the library's own
 * streams (ms), plus work counters used for GCUPS / roofline accounting */
typedef struct hlm_profile {
  double ms_h2d, ms_screen, ms_seed, ms_myers, ms_expand, ms_select, ms_dp, ms_assign, ms_d2h,
      ms_total;
  int64_t n_pairs, n_kept, n_seed_candidates, n_candidates, n_dp, n_dp_cells, n_myers_cols;
  int64_t launches;
  int64_t h2d_bytes, d2h_bytes;
  int64_t bases_bytes; /* bases+quals streamed by the screen kernel               */
  int64_t n_overflow_splits; /* chunks re-run in halves because their candidate list overflowed */
} hlm_profile;

void hlm_params_default(hlm_params* p, int rna);
int hlm_abi_version(void);

/* DB: parses the reference on-disk layout (db_{dna,rna}.info, motif/.., mapper/.., others/..) */
hlm_db* hlm_db_open(const char* db_dir, int rna, char* err, size_t errlen);
void hlm_db_close(hlm_db* db);
int hlm_db_n_loci(const hlm_db* db);
const char* hlm_db_locus_name(const hlm_db* db, int i); /* i == n_loci -> "others" */
int hlm_db_size(const hlm_db* db);                      /* SIZE: line, or 50      */
int64_t hlm_db_stat(const hlm_db* db, const char* key); /* "alleles","bases","tiles","reps","children","kmers","seed_entries" */

/* one context per GPU; not thread-safe; distinct contexts may run concurrently */
hlm_ctx* hlm_ctx_create(const hlm_db* db, int cuda_device, const hlm_params* params, char* err,
                        size_t errlen);
void hlm_ctx_destroy(hlm_ctx* ctx);
const char* hlm_last_error(const hlm_ctx* ctx);

/* stage-wise entry points (host buffers in, host buffers out) */
int hlm_screen_trim(hlm_ctx* ctx, const hlm_batch* in, hlm_screen_out* out);
int hlm_score(hlm_ctx* ctx, const hlm_batch* in, const hlm_screen_out* scr, hlm_score_out* out);
int hlm_assign(hlm_ctx* ctx, const hlm_batch* in, const hlm_screen_out* scr,
               const hlm_score_out* sc, hlm_assign_out* out);

/* Per-allele scoring.  The reference's pseudo-typing step runs
 * `bwa mem -a` of a gene's reads against a FASTA of all its alleles and reads one NM per
 * (read, allele) from the SAM (src/functions.cpp:347-520: counterhits / error per allele at
 * :394-404, nmdata[(allele, read)] at :544-545; rna twin :589-780).  This entry point returns
 * that matrix from the aligner of DESIGN.md section 3 restricted to one allele at a time: for
 * every kept pair, each mate and each record i of the locus' FASTA
 * (mapper/<kind>/<g>/<g>.fas order; locus == n_loci: others), the best alignment over the seed
 * candidates of that record, both strands - NM (mismatches + gap bases) and the clipped bases
 * (lead + trail).  nm = 255: no alignment with cost <= mm_max.  Arrays are [n_pairs * n_alleles],
 * pair-major; rows of pairs that were not kept are 255 / 0.  clip1 / clip2 may be NULL. */
int hlm_db_n_alleles(const hlm_db* db, int locus);
const char* hlm_db_allele_name(const hlm_db* db, int locus, int allele); /* FASTA header, first token */
typedef struct hlm_allele_out {
  int32_t n_alleles;      /* in: hlm_db_n_alleles(db, locus) - the row length of the arrays */
  int32_t reserved0;
  uint8_t *nm1, *clip1;
  uint8_t *nm2, *clip2;   /* paired only */
} hlm_allele_out;
int hlm_score_alleles(hlm_ctx* ctx, const hlm_batch* in, const hlm_screen_out* scr, int locus,
                      hlm_allele_out* out);

/* fused screen -> score -> assign over the context's pinned, multi-buffered
 * stream pipeline.  Any of scr / sc / asg may be NULL (result not copied back). */
int hlm_run(hlm_ctx* ctx, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc,
            hlm_assign_out* asg);

/* asynchronous form of hlm_run: the pipeline runs on a library thread; hlm_wait returns its code.
 * One run in flight per context; the arrays must stay valid until hlm_wait returns. */
int hlm_run_async(hlm_ctx* ctx, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc,
                  hlm_assign_out* asg);
int hlm_wait(hlm_ctx* ctx);

/* device-resident variant: upload once, run many times (kernel-only timing) */
hlm_devbatch* hlm_batch_upload(hlm_ctx* ctx, const hlm_batch* in);
void hlm_batch_free(hlm_ctx* ctx, hlm_devbatch* b);
int hlm_run_device(hlm_ctx* ctx, hlm_devbatch* b);
int hlm_fetch(hlm_ctx* ctx, hlm_devbatch* b, hlm_screen_out* scr, hlm_score_out* sc,
              hlm_assign_out* asg);

/* all GPUs of one box: one context + host thread per device, the batch cut into one contiguous
 * range per device, results written straight into the caller's arrays (the pairs
 * are independent, src/map_dna.cpp:917-1029, so there is no collective).  devices == NULL: all.
 * A device may be listed more than once: it then holds that many contexts.  For the file-level
 * entry points (hlm_multi_map_fastq / hlm_multi_map_reads) two contexts on one GPU keep it busy
 * across batch boundaries - one context stages and copies its next batch while the other one's
 * kernels run - which is how the command line drives a single GPU. */
typedef struct hlm_multi hlm_multi;
hlm_multi* hlm_multi_create(const hlm_db* db, const int* devices, int n_devices,
                            const hlm_params* params, char* err, size_t errlen);
void hlm_multi_destroy(hlm_multi* m);
int hlm_multi_n_devices(const hlm_multi* m);
const char* hlm_multi_last_error(const hlm_multi* m);
int hlm_multi_run(hlm_multi* m, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc,
                  hlm_assign_out* asg);

int hlm_get_profile(const hlm_ctx* ctx, hlm_profile* out);

/* ---- file-level front / back end of the dna path ----
 * Reads r1 (+ r2; NULL for single end) as FASTQ or FASTQ.gz, runs hlm_run batch by batch and
 * writes what the reference writes around the hot path, byte for byte:
 *   <out>/<sample>_selected.R{1,2}.fastq, _selected.trim.R{1,2}.fastq  (preselect.cpp:861-877,
 *                                                                      :1121-1149; R0 single end)
 *   <out>/<sample>_addressing_table.txt                                (map_dna.cpp:932-1051)
 *   <out>/<sample>_<gene>_R{1,2}.fastq, _<gene>.trim.R{1,2}.fastq      (map_dna.cpp:1056-1351)
 * Orders the reference leaves to thread timing (selected.R*.fastq records, table rows) are input
 * order / id order here. */
typedef struct hlm_file_opts {
  int32_t struct_size;
  int32_t downsampling;   /* main.cpp:86 (30); values < 5 become 5 (main.cpp:434)          */
  int32_t batch_pairs;    /* pairs per hlm_run call (0 = 1 Mi)                              */
  int32_t write_selected; /* 0: skip <sample>_selected.R*.fastq                             */
  int32_t select_only;    /* 1: `hla-mapper select`: screen + trim, only selected*.fastq      */
  int32_t threads;        /* rna: concurrent STAR runs / SAM readers (0 = one per host core)  */
  const char* star_cmd;   /* rna: the STAR executable.  It is run once per gene on
                             <sample>_sorted_<gene>_R{1,2}.fastq with the reference's arguments
                             (map_rna.cpp:676-733).  NULL: STAR is not run - the files
                             <out>/<sample>_<gene>_splice_Aligned.out.sam must already be there
                             (a missing one means "no record for this gene")                 */
} hlm_file_opts;
typedef struct hlm_file_stats {
  int64_t n_pairs, n_selected, n_kept, n_assigned, read_mean_size;
  double s_total, s_gpu, s_wait_input, s_collect, s_write;
  double s_sort;          /* id sort of the kept reads (part of s_write)                      */
} hlm_file_stats;
int hlm_map_fastq(hlm_ctx* ctx, const char* r1_path, const char* r2_path, const char* sample,
                  const char* out_dir, const hlm_file_opts* opts, hlm_file_stats* stats);
/* the same on every GPU of an hlm_multi: the reader threads' batches are dealt to the devices as
 * they arrive, results are committed in input order - the files do not depend on the number of
 * devices */
int hlm_multi_map_fastq(hlm_multi* m, const char* r1_path, const char* r2_path, const char* sample,
                        const char* out_dir, const hlm_file_opts* opts, hlm_file_stats* stats);

/* ---- rna: STAR's SAM -> block records ----
 * For callers that drive STAR themselves and score through hlm_batch.rna_blocks: every record of
 * <sample>_<gene>_splice_Aligned.out.sam cut the way src/map_rna.cpp:740-790 cuts it into
 * sorted_spliced_<gene>.fastq - the CIGAR split after every N, a piece's size = the sum of its
 * operation lengths except N and D, piece k starting at the SIZE of piece k-1 (:784), a CIGAR
 * without N = the whole SEQ; pieces clipped to SEQ.  One hlm_star_rec per piece, in file order.
 * block.start / len are in the MATE's own orientation (FLAG 0x10: SEQ is printed reverse-
 * complemented, the piece [s, e) of SEQ covers [L - e, L - s) of the mate), block.mate from FLAG
 * 0x80, block.gene = `gene`.  The caller maps qname (the read id: first header token without
 * /1 /2) to its pair index and concatenates the blocks per pair.  (hlm_map_fastq does all of this
 * itself, and decides mate / orientation by comparing SEQ with the trimmed mates.)
 * Returns the number of records (0 for a missing or empty file), < 0 on error; *out is one
 * allocation released with hlm_star_free. */
typedef struct hlm_star_rec {
  const char* qname;   /* not NUL-terminated: qname_len bytes */
  int32_t qname_len;
  int32_t flag;        /* SAM FLAG of the record */
  int32_t seq_len;     /* length of SEQ */
  int32_t piece;       /* 1-based number of the piece inside its record ($block of the reference) */
  hlm_rna_block block;
} hlm_star_rec;
int64_t hlm_star_sam_read(const char* sam_path, int gene, hlm_star_rec** out);
void hlm_star_free(hlm_star_rec* recs);

/* ---- BAM front end of the dna path ----
 * hlm_bam_extract replaces the four samtools processes of the reference's bam= input
 * (preselect.cpp:315-418: view -ML <db>/bed/select_{dna,rna}.bed, view -f4, view -N, fastq) with
 * an in-process reader (BGZF blocks inflated on `threads` host threads, 0 = all): reads that
 * overlap a BED interval or are unmapped, plus every primary record sharing their names, as the
 * R0 / R1 / R2 sets `samtools fastq` would write.  The result says whether the reference would
 * treat the BAM as paired (v_map_type, preselect.cpp:420-425); create the context with that
 * `paired` and screen_variant = HLM_SCREEN_BAM, then hlm_map_reads runs the same pipeline and
 * writes the same files as hlm_map_fastq.  bed_path == NULL: the database's own BED.
 * Errors carry the reference's messages ("Extraction failed!", "There is not enough data ..."). */
typedef struct hlm_reads hlm_reads;
typedef struct hlm_bam_stats {
  int64_t n_records, n_region, n_unmapped, n_names, n_r0, n_r1, n_r2, n_out;
  int64_t bytes_compressed, bytes_inflated;
  int32_t paired;
  int32_t cached; /* 1: the inflated BAM was kept in memory between the two scans */
  double s_pass1, s_pass2, s_total;
} hlm_bam_stats;
hlm_reads* hlm_bam_extract(const hlm_db* db, const char* bam_path, const char* bed_path,
                           int skip_unmapped, int threads, hlm_bam_stats* stats, char* err,
                           size_t errlen);
int hlm_reads_paired(const hlm_reads* r);
int64_t hlm_reads_count(const hlm_reads* r); /* pairs (paired) or reads (single end) to be mapped */
void hlm_reads_free(hlm_reads* r);
/* the reads as FASTQ, in the order and with the header lines they enter the screen with (paired:
 * both files; single end: r2_path ignored) */
int hlm_reads_write_fastq(const hlm_reads* r, const char* r1_path, const char* r2_path);
int hlm_map_reads(hlm_ctx* ctx, const hlm_reads* reads, const char* sample, const char* out_dir,
                  const hlm_file_opts* opts, hlm_file_stats* stats);
int hlm_multi_map_reads(hlm_multi* m, const hlm_reads* reads, const char* sample, const char* out_dir,
                        const hlm_file_opts* opts, hlm_file_stats* stats);

/* INT32 micro-benchmark (IADD3 / IMAD / VIADDMNMX chains) used as the DP roofline
 * denominator; returns giga warp-lane ops per second */
int hlm_int32_peak(int cuda_device, double* giga_iadd, double* giga_imad, double* giga_mix,
                   double* giga_viaddmnmx);

/* the same plus the instruction classes k_myers is made of: out[0..6] = IADD3, IMAD, IADD3+IMAD mix,
 * VIADDMNMX, LOP3, SHF, 3 LOP3 : 1 SHF mix */
int hlm_int32_peaks(int cuda_device, double* out, int n);

/* A/B of the two work layouts of the banded DP on synthetic candidates: thread per candidate with
 * the band row in registers (the product kernel) versus warp per candidate with the band across
 * lanes, shuffles and a max-plus prefix scan for the horizontal gaps (the "wavefront with warp
 * shuffles" layout).  Returns the kernel times, the number of candidates whose results differ
 * (must be 0) and the algorithmic cells per launch. */
int hlm_dp_layout_ab(int cuda_device, int n_cands, int read_len, double* ms_thread, double* ms_warp,
                     int64_t* n_mismatch, int64_t* n_cells);

#ifdef __cplusplus
}
#endif
#endif /* HLAMAP_H */
