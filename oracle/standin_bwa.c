/*
 * standin_bwa.c — a stand-in for the external `bwa` binary, for driving the UNMODIFIED
 * reference (oracle/_ref/hla-mapper-ref) through its scoring stage in a sandbox where bwa
 * 0.7.17 is not installed.  TEST INFRASTRUCTURE ONLY.
 *
 * It honours exactly the two invocations the reference makes for the hot path
 * (/root/reference/src/map_dna.cpp:749-770, 837-853):
 *     bwa aln -o 1 -t T -n 20 '<ref>.fas' <reads.fastq>  > x.sai
 *     bwa samse '<ref>.fas' x.sai <reads.fastq>           > x.sam
 * `aln` writes an empty .sai; `samse` aligns every read with the oracle aligner
 * (oracle/hlm_oracle.c) and prints one bwa-samse-shaped SAM record per read:
 *     QNAME FLAG RNAME POS MAPQ CIGAR * 0 0 SEQ QUAL XT:A:U NM:i:k
 * (NM is the 13th column, which is the one read_sam_dna parses, map_dna.cpp:57.)
 * `mem`, `index` and anything else exit 0 without output: they are outside the hot path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hlm_oracle.h"

static void strip_quotes(char* s) {
  size_t n = strlen(s);
  if (n >= 2 && s[0] == '\'' && s[n - 1] == '\'') {
    memmove(s, s + 1, n - 2);
    s[n - 2] = 0;
  }
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "\nProgram: bwa (stand-in for hla-mapper oracle tests)\nVersion: 0.7.17-standin\n");
    return 1;
  }
  if (strcmp(argv[1], "aln") == 0) return 0; /* empty .sai on stdout */
  if (strcmp(argv[1], "samse") != 0) return 0;
  if (argc < 5) return 2;
  int mode = getenv("HLM_ORACLE_EXHAUSTIVE") ? 0 : 1;
  int mm_max = getenv("HLM_MM_MAX") ? atoi(getenv("HLM_MM_MAX")) : 20;
  char* ref = strdup(argv[2]);
  strip_quotes(ref);
  orc_ref* R = orc_ref_open(ref);
  if (!R) {
    fprintf(stderr, "standin bwa: cannot open %s\n", ref);
    return 2;
  }
  printf("@SQ\tSN:standin\tLN:1\n");
  FILE* f = fopen(argv[4], "rb");
  if (!f) return 2;
  char *id = NULL, *seq = NULL, *plus = NULL, *qual = NULL;
  size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  ssize_t n1, n2, n4;
  while ((n1 = getline(&id, &c1, f)) >= 0) {
    n2 = getline(&seq, &c2, f);
    if (getline(&plus, &c3, f) < 0) break;
    n4 = getline(&qual, &c4, f);
    if (n2 < 0 || n4 < 0) break;
    while (n1 > 0 && (id[n1 - 1] == '\n' || id[n1 - 1] == '\r')) id[--n1] = 0;
    while (n2 > 0 && (seq[n2 - 1] == '\n' || seq[n2 - 1] == '\r')) seq[--n2] = 0;
    while (n4 > 0 && (qual[n4 - 1] == '\n' || qual[n4 - 1] == '\r')) qual[--n4] = 0;
    char* q = id + 1; /* QNAME = header up to the first blank, as bwa does */
    char* sp = strpbrk(q, " \t");
    if (sp) *sp = 0;
    int S, cost, lead, trail;
    if (orc_ref_align(R, (const uint8_t*)seq, (int)n2, mode, mm_max, &S, &cost, &lead, &trail)) {
      char cigar[64];
      int m = (int)n2 - lead - trail, o = 0;
      if (lead) o += sprintf(cigar + o, "%dS", lead);
      o += sprintf(cigar + o, "%dM", m);
      if (trail) o += sprintf(cigar + o, "%dS", trail);
      printf("%s\t0\tstandin\t1\t37\t%s\t*\t0\t0\t%s\t%s\tXT:A:U\tNM:i:%d\n", q, cigar, seq, qual,
             cost - lead - trail);
    } else {
      printf("%s\t4\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", q, seq, qual);
    }
  }
  fclose(f);
  orc_ref_close(R);
  return 0;
}
