/*
 * hlm_oracle.h — CPU oracle for the hla-mapper hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * executed from the product (libhlamap / hla_mapper_b200).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / reported baseline.
 *
 * It restates, in plain C, the reference's algorithm for every stage of the path
 * (each function cites the reference file:line it follows) plus the aligner
 * definition that replaces the external `bwa aln/samse` call.
 *
 * Parity status: screen / mtrim / pair filter / locus sort / SAM reducer / compare are
 * pinned against the unmodified reference sources compiled into oracle/_ref
 * (tests/test_oracle_vs_reference.py, tests/golden/).  The DP scoring (orc_align) is
 * pinned to the real `bwa mem` records of the reference's test BAM
 * (tests/test_bwa_mem_records.py: clips, CIGAR score, NM).  The aligner as a whole
 * against `bwa aln` is "parity unpinned": the reference shells out to bwa 0.7.17
 * (hla-mapper.yml:12), which is neither vendored nor installed, and the reference
 * ships no `aln` output.
 */
#ifndef HLM_ORACLE_H
#define HLM_ORACLE_H

#include "../include/hlamap.h" /* plain-C struct layouts only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_db orc_db;

orc_db* orc_db_open(const char* db_dir, int rna, char* err, size_t errlen);
void orc_db_close(orc_db* db);
int orc_db_n_loci(const orc_db* db);
const char* orc_db_locus_name(const orc_db* db, int i);
int orc_db_size(const orc_db* db);
int64_t orc_db_n_alleles(const orc_db* db, int locus);

/* unit-level pieces */
int orc_mtrim(const uint8_t* qual, int qlen, int slen, double err, int v_size, int* lo, int* len);
int orc_screen_read(const orc_db* db, const uint8_t* seq, int len, int variant, int minhit);
uint32_t orc_locus_mask(const orc_db* db, const uint8_t* seq, int len);
/* banded affine alignment of read codes (0..3, 4 = N) against ref codes on diagonal d */
void orc_align(const uint8_t* r, int n, const uint8_t* t, int tlen, int d, int* S, int* cost,
               int* lead, int* trail);
void orc_set_noclip(int on); /* unit-level calls: orc_align in HLM_CLIP_NONE mode (no clipping at all) */
int orc_banded_ed(const uint8_t* r, int n, const uint8_t* t, int tlen, int d);
/* SAM reducers on real SAM text (used to pin read_sam_dna / read_sam_rna) */
int orc_sam_score_dna(const char* sam_line, int clip_mode, char* qname, size_t qlen);
int orc_sam_score_rna(const char* sam_line, int clip_mode, char* qname, size_t qlen);

/* whole stages, same array layouts as hlamap.h.  mode: 0 = exhaustive definition
 * (every allele, every seed diagonal), 1 = fast CPU port (footprint de-duplication +
 * edit-distance ordering; must equal mode 0).  nthreads <= 0: all cores. */
int orc_screen_trim(const orc_db* db, const hlm_params* p, const hlm_batch* in, hlm_screen_out* out,
                    int nthreads);
int orc_score(const orc_db* db, const hlm_params* p, const hlm_batch* in, const hlm_screen_out* scr,
              hlm_score_out* out, int mode, int nthreads, int64_t* n_dp_cells);
/* per-allele matrix of one target set (hlm_score_alleles): exhaustive definition */
int orc_score_alleles(const orc_db* db, const hlm_params* p, const hlm_batch* in,
                      const hlm_screen_out* scr, int locus, uint8_t* nm1, uint8_t* clip1, uint8_t* nm2,
                      uint8_t* clip2, int nthreads);
int orc_assign(const orc_db* db, const hlm_params* p, const hlm_batch* in, const hlm_screen_out* scr,
               const hlm_score_out* sc, hlm_assign_out* out);

/* one multi-FASTA as a target set (stand-in bwa) */
typedef struct orc_ref orc_ref;
orc_ref* orc_ref_open(const char* fasta);
void orc_ref_close(orc_ref* r);
int orc_ref_align(const orc_ref* r, const uint8_t* ascii, int len, int mode, int mm_max, int* S,
                  int* cost, int* lead, int* trail);
void orc_ref_align_batch(const orc_ref* r, const uint8_t* const* reads, const int* lens, int64_t n,
                         int mode, int mm_max, int nthreads, int32_t* out, uint8_t* ok);

#ifdef __cplusplus
}
#endif
#endif
