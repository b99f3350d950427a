// hlm_cli.cpp — thin command line over libhlamap that keeps the reference's `hla-mapper dna` /
// `select` key=value interface (src/main.cpp:304-671) for the part of the run the hot path
// covers.  It is not a re-implementation of hla-mapper's control plane: no setup wizard, no
// typing, no realignment - it ends where src/map_dna.cpp:1367 begins, leaving the files the
// downstream steps read (selected*.fastq, addressing_table.txt, per-gene FASTQs).
//
//   hla-mapper-b200 dna    r1=R1.fastq[.gz] r2=R2.fastq[.gz] sample=NAME [db=DIR] [output=DIR] ...
//   hla-mapper-b200 dna    r0=R0.fastq[.gz] sample=NAME ...
//   hla-mapper-b200 dna    bam=ALIGNED.bam sample=NAME [--skip-unmapped] ...   (no samtools needed)
//   hla-mapper-b200 select ...           (screen + trim only: selected*.fastq)
//   hla-mapper-b200 rna    r1=... r2=... sample=NAME [star=STAR] ...   (src/map_rna.cpp: screen, trim,
//                          per-gene sort, STAR per gene, block scoring, compare; star defaults to the
//                          `star=` line of ~/.hla-mapper; star=none: the *_splice_Aligned.out.sam
//                          files are already in the output directory)
//
// Options honoured: db= output= sample= size= tolerance= error= downsample= gpu= gpus= threads=
// --quiet --skip-unmapped (threads= sizes the BAM reader only; buffer= is accepted and ignored).
// gpus=0,1,2,3 feeds several GPUs (batches dealt to the devices, files as from one); a device may
// be listed twice.  Without gpus=, dna / select runs hold two contexts on the GPU gpu= names, so
// that it stays busy across batch boundaries.
// db defaults to the `db=` line of ~/.hla-mapper (home from getpwuid, src/main.cpp:266-273).
#include <pwd.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/hlamap.h"

static bool starts(const std::string& s, const char* k) { return s.compare(0, strlen(k), k) == 0; }

static std::string config_value(const char* key) {
  struct passwd* pw = getpwuid(getuid());
  if (!pw) return "";
  std::ifstream f(std::string(pw->pw_dir) + "/.hla-mapper");
  for (std::string line; std::getline(f, line);)
    if (starts(line, key)) return line.substr(strlen(key));
  return "";
}
static std::string config_db() { return config_value("db="); }

int main(int argc, char** argv) {
  if (argc < 2 || (strcmp(argv[1], "dna") != 0 && strcmp(argv[1], "select") != 0 && strcmp(argv[1], "rna") != 0)) {
    fprintf(stderr,
            "Usage:     hla-mapper-b200 dna r1=R1.gz r2=R2.gz sample=your_sample_name <options>\n"
            "           hla-mapper-b200 dna r0=R0.gz sample=your_sample_name <options>\n"
            "           hla-mapper-b200 dna bam=sample.bam sample=your_sample_name <options>\n"
            "           hla-mapper-b200 select r1=... r2=... sample=... <options>\n"
            "           hla-mapper-b200 rna r1=R1.gz r2=R2.gz sample=your_sample_name [star=STAR] <options>\n"
            "Options:   db= output= size= tolerance= error= downsample= gpu= gpus= threads= --quiet --skip-unmapped\n"
            "           --exact-scores\n");
    return 1;
  }
  bool select_only = strcmp(argv[1], "select") == 0, quiet = false;
  const int rna = strcmp(argv[1], "rna") == 0 ? 1 : 0;
  std::string r0, r1, r2, bam, sample, db, output, star, gpus;
  double tolerance = -1, error = -1;
  int size = 0, downsample = 30, gpu = 0, threads = 0, skip_unmapped = 0, exact_scores = 0;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    if (starts(a, "r0=")) r0 = a.substr(3);
    else if (starts(a, "r1=")) r1 = a.substr(3);
    else if (starts(a, "r2=")) r2 = a.substr(3);
    else if (starts(a, "sample=")) sample = a.substr(7);
    else if (starts(a, "db=")) db = a.substr(3);
    else if (starts(a, "output=")) output = a.substr(7);
    else if (starts(a, "size=")) size = atoi(a.c_str() + 5);
    else if (starts(a, "tolerance=")) tolerance = atof(a.c_str() + 10);
    else if (starts(a, "error=")) error = std::stod(a.substr(6));  // stod, src/main.cpp:658-662
    else if (starts(a, "downsample=")) downsample = atoi(a.c_str() + 11);
    else if (starts(a, "gpus=")) gpus = a.substr(5);
    else if (starts(a, "gpu=")) gpu = atoi(a.c_str() + 4);
    else if (a == "--quiet") quiet = true;
    else if (a == "--skip-unmapped") skip_unmapped = 1;  // src/main.cpp:626-630
    else if (a == "--exact-scores") exact_scores = 1;
    else if (starts(a, "threads=")) threads = atoi(a.c_str() + 8);
    else if (starts(a, "bam=")) bam = a.substr(4);
    else if (starts(a, "star=")) star = a.substr(5);
    else if (starts(a, "buffer=") || starts(a, "--")) continue;
  }
  bool paired = r0.empty();
  if (sample.empty() || (bam.empty() && paired && (r1.empty() || r2.empty()))) {
    fprintf(stderr, "You must indicate r1 and r2, or r0, or bam, and a sample name.\n");
    return 2;
  }
  if (db.empty()) db = config_db();
  if (db.empty()) {
    fprintf(stderr, "No database: pass db= or set it in ~/.hla-mapper\n");
    return 2;
  }
  const std::string& first = !bam.empty() ? bam : paired ? r1 : r0;
  if (output.empty()) {  // src/map_dna.cpp:404-409
    size_t sl = first.find_last_of('/');
    output = (sl == std::string::npos ? std::string(".") : first.substr(0, sl)) + "/hla-mapper/";
  }
  mkdir(output.c_str(), 0777);
  char err[512];
  if (rna && !bam.empty()) {
    fprintf(stderr, "rna: bam= input is not supported by this front end; pass r1= r2= (or r0=)\n");
    return 2;
  }
  if (rna && star.empty()) star = config_value("star=");
  if (star == "none") star.clear();
  hlm_db* D = hlm_db_open(db.c_str(), rna, err, sizeof err);
  if (!D) {
    fprintf(stderr, "%s\n", err);
    return 3;
  }
  // bam=: the reads come out of the BAM first; whether they are paired is a property of the file
  // (v_map_type, src/preselect.cpp:420-425)
  hlm_reads* R = nullptr;
  hlm_bam_stats bs;
  if (!bam.empty()) {
    R = hlm_bam_extract(D, bam.c_str(), nullptr, skip_unmapped, threads, &bs, err, sizeof err);
    if (!R) {
      fprintf(stderr, "%s\n", err);
      hlm_db_close(D);
      return 3;
    }
    paired = hlm_reads_paired(R) != 0;
    if (!quiet)
      fprintf(stderr, "%lld BAM records (%.2f GB inflated) in %.2fs: %lld in the regions, %lld unmapped, %lld names -> %lld %s\n",
              (long long)bs.n_records, bs.bytes_inflated / 1e9, bs.s_total, (long long)bs.n_region,
              (long long)bs.n_unmapped, (long long)bs.n_names, (long long)bs.n_out, paired ? "pairs" : "reads");
  }
  hlm_params p;
  hlm_params_default(&p, rna);
  p.paired = paired ? 1 : 0;
  if (R) p.screen_variant = HLM_SCREEN_BAM;
  // the command line only produces files: scores the compare step cannot use are not computed
  // (tolerance-capped scoring: the files are byte-identical); --exact-scores
  // switches it off
  p.score_cap = exact_scores ? 0 : 1;
  if (size > 0) p.v_size = size;
  if (tolerance >= 0) p.tolerance = tolerance;
  if (error >= 0) p.mtrim_error = error;
  // rna keeps one context (STAR and the block scoring run on it after the screen); dna / select
  // go through hlm_multi: the listed devices, or two contexts on one GPU
  std::vector<int> devs;
  for (size_t at = 0; at < gpus.size();) {
    size_t c = gpus.find(',', at);
    if (c == std::string::npos) c = gpus.size();
    if (c > at) devs.push_back(atoi(gpus.substr(at, c - at).c_str()));
    at = c + 1;
  }
  if (devs.empty() && !rna) devs = {gpu, gpu};
  hlm_ctx* C = nullptr;
  hlm_multi* M = nullptr;
  if (rna) C = hlm_ctx_create(D, devs.empty() ? gpu : devs[0], &p, err, sizeof err);
  else M = hlm_multi_create(D, devs.data(), (int)devs.size(), &p, err, sizeof err);
  if (!C && !M) {
    fprintf(stderr, "%s\n", err);
    hlm_reads_free(R);
    hlm_db_close(D);
    return 3;
  }
  hlm_file_opts fo;
  memset(&fo, 0, sizeof fo);
  fo.struct_size = (int32_t)sizeof fo;
  fo.downsampling = downsample;
  fo.write_selected = 1;
  fo.select_only = select_only ? 1 : 0;
  fo.threads = threads;
  fo.star_cmd = star.empty() ? nullptr : star.c_str();
  hlm_file_stats st;
  int rc;
  if (M)
    rc = R ? hlm_multi_map_reads(M, R, sample.c_str(), output.c_str(), &fo, &st)
           : hlm_multi_map_fastq(M, first.c_str(), paired ? r2.c_str() : nullptr, sample.c_str(), output.c_str(), &fo, &st);
  else
    rc = hlm_map_fastq(C, first.c_str(), paired ? r2.c_str() : nullptr, sample.c_str(), output.c_str(), &fo, &st);
  hlm_reads_free(R);
  if (rc != HLM_OK) fprintf(stderr, "hla-mapper-b200: %s\n", M ? hlm_multi_last_error(M) : hlm_last_error(C));
  else if (!quiet)
    fprintf(stderr,
            "%lld pairs, %lld selected, %lld kept, %lld addressed; %.2fs (gpu wait %.2fs, input wait %.2fs, "
            "collect %.2fs, write %.2fs) -> %s\n",
            (long long)st.n_pairs, (long long)st.n_selected, (long long)st.n_kept, (long long)st.n_assigned,
            st.s_total, st.s_gpu, st.s_wait_input, st.s_collect, st.s_write, output.c_str());
  if (M) hlm_multi_destroy(M);
  if (C) hlm_ctx_destroy(C);
  hlm_db_close(D);
  return rc == HLM_OK ? 0 : 4;
}
