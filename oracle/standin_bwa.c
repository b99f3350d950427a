/*
 * standin_bwa.c — a stand-in for the external `bwa` binary, for driving the UNMODIFIED
 * reference (oracle/_ref/hla-mapper-ref) through its scoring stage in a sandbox where bwa
 * 0.7.17 is not installed.  TEST INFRASTRUCTURE ONLY.
 *
 * It honours exactly the two invocations the reference makes for the hot path
 * (/root/reference/src/map_dna.cpp:749-770, 837-853):
 *     bwa aln -o 1 -t T -n 20 '<ref>.fas' <reads.fastq>  > x.sai
 *     bwa samse '<ref>.fas' x.sai <reads.fastq>           > x.sam
 * Like the real bwa, `aln` does the search on T threads and writes its result per read to the
 * .sai stream (here: S, cost, lead, trail of the oracle aligner, oracle/hlm_oracle.c, DESIGN.md
 * section 3); `samse` turns .sai + FASTQ into one bwa-samse-shaped SAM record per read:
 *     QNAME FLAG RNAME POS MAPQ CIGAR * 0 0 SEQ QUAL XT:A:U NM:i:k
 * (NM is the 13th column, which is the one read_sam_dna parses, map_dna.cpp:57.)
 * The index of a FASTA is cached next to it (<fasta>.orcidx), as `bwa index` files are.
 * `mem`, `index` and anything else exit 0 without output: they are outside the hot path.
 */
#define _GNU_SOURCE
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

typedef struct {
  char** id;
  char** seq;
  char** qual;
  int* len;
  int64_t n, cap;
} fq;

static int read_fastq(const char* path, fq* F) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  memset(F, 0, sizeof *F);
  char *l1 = NULL, *l2 = NULL, *l3 = NULL, *l4 = NULL;
  size_t c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  ssize_t n1, n2, n4;
  while ((n1 = getline(&l1, &c1, f)) >= 0) {
    n2 = getline(&l2, &c2, f);
    if (getline(&l3, &c3, f) < 0) break;
    n4 = getline(&l4, &c4, f);
    if (n2 < 0 || n4 < 0) break;
    while (n1 > 0 && (l1[n1 - 1] == '\n' || l1[n1 - 1] == '\r')) l1[--n1] = 0;
    while (n2 > 0 && (l2[n2 - 1] == '\n' || l2[n2 - 1] == '\r')) l2[--n2] = 0;
    while (n4 > 0 && (l4[n4 - 1] == '\n' || l4[n4 - 1] == '\r')) l4[--n4] = 0;
    if (F->n == F->cap) {
      F->cap = F->cap ? F->cap * 2 : 4096;
      F->id = (char**)realloc(F->id, F->cap * sizeof(char*));
      F->seq = (char**)realloc(F->seq, F->cap * sizeof(char*));
      F->qual = (char**)realloc(F->qual, F->cap * sizeof(char*));
      F->len = (int*)realloc(F->len, F->cap * sizeof(int));
    }
    char* q = l1 + 1; /* QNAME = header up to the first blank, as bwa does */
    char* sp = strpbrk(q, " \t");
    if (sp) *sp = 0;
    F->id[F->n] = strdup(q);
    F->seq[F->n] = strdup(l2);
    F->qual[F->n] = strdup(l4);
    F->len[F->n] = (int)n2;
    F->n++;
  }
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "\nProgram: bwa (stand-in for hla-mapper oracle tests)\nVersion: 0.7.17-standin\n");
    return 1;
  }
  int mode = getenv("HLM_ORACLE_EXHAUSTIVE") ? 0 : 1;
  if (strcmp(argv[1], "aln") == 0) {
    /* bwa aln -o 1 -t T -n 20 ref reads */
    int threads = 1, mm_max = 20, i = 2;
    for (; i + 1 < argc && argv[i][0] == '-'; i += 2) {
      if (strcmp(argv[i], "-t") == 0) threads = atoi(argv[i + 1]);
      if (strcmp(argv[i], "-n") == 0) mm_max = atoi(argv[i + 1]);
    }
    if (i + 1 >= argc) return 2;
    char* ref = strdup(argv[i]);
    strip_quotes(ref);
    orc_ref* R = orc_ref_open(ref);
    fq F;
    if (!R || read_fastq(argv[i + 1], &F) != 0) {
      fprintf(stderr, "standin bwa aln: cannot open %s / %s\n", ref, argv[i + 1]);
      return 2;
    }
    int32_t* out = (int32_t*)calloc((size_t)F.n * 4 + 4, 4);
    uint8_t* ok = (uint8_t*)calloc((size_t)F.n + 1, 1);
    orc_ref_align_batch(R, (const uint8_t* const*)F.seq, F.len, F.n, mode, mm_max, threads, out, ok);
    int64_t hdr[2] = {0x5341495354443031LL, F.n};
    fwrite(hdr, 8, 2, stdout);
    fwrite(ok, 1, (size_t)F.n, stdout);
    fwrite(out, 4, (size_t)F.n * 4, stdout);
    return 0;
  }
  if (strcmp(argv[1], "samse") != 0) return 0;
  if (argc < 5) return 2;
  /* bwa samse ref x.sai reads */
  fq F;
  if (read_fastq(argv[4], &F) != 0) return 2;
  FILE* s = fopen(argv[3], "rb");
  int64_t hdr[2] = {0, 0};
  if (!s || fread(hdr, 8, 2, s) != 2 || hdr[0] != 0x5341495354443031LL || hdr[1] != F.n) {
    fprintf(stderr, "standin bwa samse: bad .sai %s\n", argv[3]);
    return 2;
  }
  uint8_t* ok = (uint8_t*)malloc((size_t)F.n + 1);
  int32_t* out = (int32_t*)malloc(((size_t)F.n * 4 + 4) * 4);
  if (fread(ok, 1, (size_t)F.n, s) != (size_t)F.n || fread(out, 4, (size_t)F.n * 4, s) != (size_t)F.n * 4)
    return 2;
  fclose(s);
  printf("@SQ\tSN:standin\tLN:1\n");
  for (int64_t i = 0; i < F.n; i++) {
    if (ok[i]) {
      int cost = out[4 * i + 1], lead = out[4 * i + 2], trail = out[4 * i + 3];
      char cigar[64];
      int m = F.len[i] - lead - trail, o = 0;
      if (lead) o += sprintf(cigar + o, "%dS", lead);
      o += sprintf(cigar + o, "%dM", m);
      if (trail) o += sprintf(cigar + o, "%dS", trail);
      printf("%s\t0\tstandin\t1\t37\t%s\t*\t0\t0\t%s\t%s\tXT:A:U\tNM:i:%d\n", F.id[i], cigar, F.seq[i],
             F.qual[i], cost - lead - trail);
    } else {
      printf("%s\t4\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", F.id[i], F.seq[i], F.qual[i]);
    }
  }
  return 0;
}
