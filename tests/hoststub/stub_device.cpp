// stub_device.cpp — TEST INFRASTRUCTURE: a stand-in for the device side of libhlamap so that the
// host file pipeline (csrc/hlm_files.cpp: readers, batch plan, driver threads, collection, id
// sort, output preparation, writers) runs on a machine without a GPU, and under the address /
// undefined-behaviour sanitizers.  hlm_files.cpp is compiled unchanged and linked against this
// file instead of hlm_api.cu; "results" are a deterministic function of the CRC-32 of a pair's
// first mate, which tests/test_host_pipeline.py restates in Python to predict every output file
// byte for byte (formats: src/preselect.cpp:863-864, :1121-1149; src/map_dna.cpp:1036-1051,
// :1168-1266).  Nothing of the product links or loads this file.
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hlamap.h"
#include "../../hla_mapper_b200/csrc/hlm_db.h"

struct hlm_ctx {
  hlm_db* db;
  hlm_params p;
  std::string err;
  int delay_us_per_kpair = 0;  // pretend device time
};
struct hlm_multi {
  std::vector<hlm_ctx*> ctx;
  std::string err;
};

const hlm_db* hlm_ctx_db(const hlm_ctx* c) { return c->db; }
const hlm_params* hlm_ctx_params(const hlm_ctx* c) { return &c->p; }
void hlm_ctx_set_error(hlm_ctx* c, const std::string& m) { c->err = m; }
hlm_ctx* const* hlm_multi_contexts(const hlm_multi* m, int* n) {
  *n = (int)m->ctx.size();
  return m->ctx.data();
}
void hlm_multi_set_error(hlm_multi* m, const std::string& s) { m->err = s; }

static void fake_results(hlm_ctx* c, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc, hlm_assign_out* as) {
  const int G = c->db->n_loci, T = G + 1;
  const bool paired = c->p.paired != 0;
  for (int64_t i = 0; i < in->n_pairs; i++) {
    uint64_t a = in->offs1[i], b = in->offs1[i + 1];
    uint32_t L1 = (uint32_t)(b - a), L2 = paired ? (uint32_t)(in->offs2[i + 1] - in->offs2[i]) : 0;
    uint32_t h = (uint32_t)crc32(crc32(0L, Z_NULL, 0), in->bases1 + a, L1);
    bool screened = (h % 10) != 0, kept = screened && ((h >> 4) % 7) != 0;
    scr->flags[i] = (uint8_t)((screened ? HLM_F_SCREENED : 0) | (kept ? HLM_F_KEPT : 0));
    uint32_t lo1 = L1 ? (h % 5) % L1 : 0, lo2 = L2 ? ((h >> 3) % 4) % L2 : 0;
    int n1 = (int)L1 - (int)lo1 - (int)((h >> 8) % 3), n2 = (int)L2 - (int)lo2 - (int)((h >> 10) % 2);
    scr->trim_lo1[i] = (uint16_t)lo1;
    scr->trim_len1[i] = (uint16_t)(n1 < 1 ? (L1 ? 1 : 0) : n1);
    scr->trim_lo2[i] = (uint16_t)lo2;
    scr->trim_len2[i] = (uint16_t)(n2 < 1 ? (L2 ? 1 : 0) : n2);
    uint32_t lmask = (h >> 12) & ((1u << G) - 1);
    scr->locus_mask[i] = lmask;
    if (!sc) continue;
    for (int g = 0; g < T; g++) {
      bool on = g == T - 1 || ((lmask >> g) & 1u);
      sc->s1[i * T + g] = (uint16_t)(on ? 1 + ((h >> (g + 1)) % 9) : 0);
      sc->s2[i * T + g] = (uint16_t)(on ? 1 + ((h >> (g + 2)) % 11) : 0);
    }
    uint32_t visible = lmask | (((h >> 22) & 1u) << G), target = 0;
    for (int g = 0; g < T; g++)
      if (((visible >> g) & 1u) && ((h >> (g + 5)) & 1u)) target |= 1u << g;
    as->target_mask[i] = kept ? target : 0;
    as->visible_mask[i] = kept ? visible : 0;
    as->min_sum[i] = (uint16_t)(h % 40);
    as->others_s1[i] = (uint16_t)(1 + (h >> 24) % 7);
    as->others_s2[i] = (uint16_t)(1 + (h >> 27) % 5);
  }
  if (c->delay_us_per_kpair)
    std::this_thread::sleep_for(std::chrono::microseconds((int64_t)c->delay_us_per_kpair * in->n_pairs / 1000));
}

extern "C" {
const char* hlm_last_error(const hlm_ctx* c) { return c->err.c_str(); }
int hlm_screen_trim(hlm_ctx* c, const hlm_batch* in, hlm_screen_out* out) {
  fake_results(c, in, out, nullptr, nullptr);
  return 0;
}
// the real call returns at once and hlm_wait() joins the work; here the work is done in place
int hlm_run_async(hlm_ctx* c, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc, hlm_assign_out* as) {
  fake_results(c, in, scr, sc, as);
  return 0;
}
int hlm_wait(hlm_ctx*) { return 0; }
int hlm_get_profile(const hlm_ctx*, hlm_profile* p) {
  memset(p, 0, sizeof *p);
  return 0;
}
int hlm_score(hlm_ctx*, const hlm_batch*, const hlm_screen_out*, hlm_score_out*) { return HLM_E_ARG; }
int hlm_assign(hlm_ctx*, const hlm_batch*, const hlm_screen_out*, const hlm_score_out*, hlm_assign_out*) {
  return HLM_E_ARG;
}
}

// hoststub_run r1 r2|- sample out_dir batch_pairs n_ctx n_loci downsampling [select_only]
int main(int argc, char** argv) {
  if (argc < 9) {
    fprintf(stderr, "usage: hoststub_run r1 r2|- sample out_dir batch_pairs n_ctx n_loci downsampling [select_only]\n");
    return 2;
  }
  const char* r1 = argv[1];
  const char* r2 = strcmp(argv[2], "-") ? argv[2] : nullptr;
  int batch = atoi(argv[5]), n_ctx = atoi(argv[6]), G = atoi(argv[7]), down = atoi(argv[8]);
  int select_only = argc > 9 ? atoi(argv[9]) : 0;
  hlm_db db;
  db.n_loci = G;
  for (int g = 0; g < G; g++) {
    db.names.push_back("HLA-T" + std::to_string(g));
    db.gene_opt_start.push_back(1000 * g);
    db.gene_opt_end.push_back(1000 * g + 400 + 350 * g);
  }
  db.names.push_back("others");
  std::vector<hlm_ctx> ctxs((size_t)n_ctx);
  hlm_multi mm;
  for (auto& c : ctxs) {
    c.db = &db;
    memset(&c.p, 0, sizeof c.p);
    c.p.struct_size = (int32_t)sizeof c.p;
    c.p.paired = r2 ? 1 : 0;
    const char* d = getenv("STUB_DELAY_US_PER_KPAIR");
    c.delay_us_per_kpair = d ? atoi(d) : 0;
    mm.ctx.push_back(&c);
  }
  hlm_file_opts fo;
  memset(&fo, 0, sizeof fo);
  fo.struct_size = (int32_t)sizeof fo;
  fo.downsampling = down;
  fo.batch_pairs = batch;
  fo.write_selected = 1;
  fo.select_only = select_only;
  hlm_file_stats st;
  memset(&st, 0, sizeof st);
  int rc = n_ctx > 1 ? hlm_multi_map_fastq(&mm, r1, r2, argv[3], argv[4], &fo, &st)
                     : hlm_map_fastq(&ctxs[0], r1, r2, argv[3], argv[4], &fo, &st);
  if (rc) {
    fprintf(stderr, "rc %d: %s\n", rc, n_ctx > 1 ? mm.err.c_str() : ctxs[0].err.c_str());
    return 1;
  }
  printf("{\"n_pairs\": %lld, \"n_selected\": %lld, \"n_kept\": %lld, \"n_assigned\": %lld, \"read_mean_size\": %lld, "
         "\"s_total\": %.4f, \"s_collect\": %.4f, \"s_sort\": %.4f, \"s_write\": %.4f}\n",
         (long long)st.n_pairs, (long long)st.n_selected, (long long)st.n_kept, (long long)st.n_assigned,
         (long long)st.read_mean_size, st.s_total, st.s_collect, st.s_sort, st.s_write);
  return 0;
}
