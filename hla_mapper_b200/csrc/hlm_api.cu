// hlm_api.cu — the C-ABI of libhlamap and the host runtime behind it.
//
// Replaces the reference's per-read ThreadPool dispatch (src/ThreadPool.hpp:18-100, one
// packaged_task + two global-mutex acquisitions per read, preselect.cpp:696-745) with a
// chunked, multi-buffered CUDA-stream pipeline: three slots, each with its own stream, pinned
// staging buffers and device work buffers; while slot A's kernels run, slot B's inputs are
// staged and copied, and A's results drain back.  No host<->device synchronisation happens
// inside a chunk; all kernel grids are persistent (multiples of the SM count) and read their
// work counts from device memory.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "hlm_db.h"
#include "hlm_kernels.cuh"

#define HLM_STAGE_SCREEN 1
#define HLM_STAGE_SCORE 2
#define HLM_STAGE_ASSIGN 4
#define HLM_STAGE_ALLELES 8  /* hlm_score_alleles: needs provided screen results, excludes the others */
#define HLM_MAX_SLOTS 4
#define HLM_E_OVERFLOW -100 /* internal: a chunk produced more candidates than cand_capacity */

namespace {

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

enum { EV_START, EV_H2D, EV_SCREEN, EV_SEED, EV_MYA, EV_EX0, EV_MYB0, EV_SEL0, EV_DP0, EV_EX1,
       EV_MYB1, EV_SEL1, EV_DP1, EV_ASSIGN, EV_D2H, EV_COUNT };

struct Slot {
  cudaStream_t st = nullptr;
  cudaEvent_t ev[EV_COUNT];
  Buf h_in, h_out, d_in, d_out, d_units, d_unit_meta, d_cand, d_dp, d_tasks, d_counters;
  Buf d_cres, d_albest, d_alout;  // hlm_score_alleles only
  Buf d_dptab;                    // footprint de-duplication table of the DP list
  Buf d_prelist;                  // k_prescreen's list of pairs for k_screen
  HlmChunk ck;
  bool busy = false;
  int64_t first = 0, n = 0;
  size_t out_off[32];
  size_t out_bytes = 0;
};

}  // namespace

struct hlm_devbatch {
  int64_t n_pairs = 0;
  Buf in, out;
  size_t in_off[16];
  size_t out_off[32];
  int64_t max_len = 0;
  bool paired = true;
  bool has_override = false, has_split1 = false, has_split2 = false;
  std::vector<uint64_t> blk_off;  // host copy of rna_block_off (chunk slicing); empty = no block list
  int64_t read_bytes = 0;
};

struct hlm_ctx {
  const hlm_db* db = nullptr;
  hlm_params p;
  HlmKParams kp;
  HlmDevDB ddb;
  int device = 0, sms = 148;
  int n_targets = 0, units_per_pair = 2, n_slots = 2;
  int64_t chunk_pairs = 0, cand_cap = 0;
  Slot slot[HLM_MAX_SLOTS];
  std::string err;
  hlm_profile prof;
  cudaEvent_t run_start = nullptr;
  std::future<int> pending;  // hlm_run_async
  // measurement aid (HLM_TRACE=<file>): the time of every stage event of every chunk relative
  // to the start of the run, one line per chunk, appended when a run ends
  std::vector<float> trace;
  bool tracing = false;
  // consecutive chunks run out of phase: a chunk's kernels start when the previous chunk (on
  // another slot) has passed stage event `stagger_ev`, so that one chunk's list kernels run
  // beside another chunk's ALU-bound kernels instead of beside the same kernels of the other
  // slots (-1: the slots start together and stay in lockstep)
  int stagger_ev = -1;
  Slot* last_submitted = nullptr;
};

namespace {

bool cuda_ok(hlm_ctx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  c->err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}
#define CU(call)                                   \
  do {                                             \
    if (!cuda_ok(ctx, (call), #call)) return HLM_E_CUDA; \
  } while (0)

size_t align16(size_t x) { return (x + 255) & ~(size_t)255; }

int ensure_dev(hlm_ctx* ctx, Buf& b, size_t need) {
  if (need <= b.cap) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t cap = need + need / 8 + 4096;
  if (!cuda_ok(ctx, cudaMalloc(&b.p, cap), "cudaMalloc")) return HLM_E_NOMEM;
  b.cap = cap;
  return 0;
}
int ensure_pinned(hlm_ctx* ctx, Buf& b, size_t need) {
  if (need <= b.cap) return 0;
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
  size_t cap = need + need / 8 + 4096;
  if (!cuda_ok(ctx, cudaMallocHost(&b.p, cap), "cudaMallocHost")) return HLM_E_NOMEM;
  b.cap = cap;
  return 0;
}

// output block layout for n pairs and T targets; returns total bytes
enum { O_FLAGS, O_LO1, O_LEN1, O_LO2, O_LEN2, O_MASK, O_S1, O_S2, O_ALN1, O_ALN2, O_NM1, O_NM2,
       O_LEAD1, O_LEAD2, O_TRAIL1, O_TRAIL2, O_TARGET, O_VISIBLE, O_MINSUM, O_OS1, O_OS2, O_COUNTERS,
       O_COUNT };
size_t out_layout(int64_t n, int T, size_t* off) {
  size_t z = (size_t)n, zt = (size_t)n * (size_t)T;
  size_t sz[O_COUNT] = {z, 2 * z, 2 * z, 2 * z, 2 * z, 4 * z, 2 * zt, 2 * zt, 2 * zt, 2 * zt,
                        zt, zt, zt, zt, zt, zt, 4 * z, 4 * z, 2 * z, 2 * z, 2 * z,
                        (size_t)8 * HLM_N_COUNTERS};
  size_t o = 0;
  for (int i = 0; i < O_COUNT; i++) {
    off[i] = o;
    o += align16(sz[i]);
  }
  return o;
}

void bind_outputs(HlmChunk& ck, uint8_t* base, const size_t* off, int64_t first, int T) {
  ck.flags = base + off[O_FLAGS] + first;
  ck.trim_lo1 = (uint16_t*)(base + off[O_LO1]) + first;
  ck.trim_len1 = (uint16_t*)(base + off[O_LEN1]) + first;
  ck.trim_lo2 = (uint16_t*)(base + off[O_LO2]) + first;
  ck.trim_len2 = (uint16_t*)(base + off[O_LEN2]) + first;
  ck.locus_mask = (uint32_t*)(base + off[O_MASK]) + first;
  ck.s1 = (uint16_t*)(base + off[O_S1]) + first * T;
  ck.s2 = (uint16_t*)(base + off[O_S2]) + first * T;
  ck.aln1 = (int16_t*)(base + off[O_ALN1]) + first * T;
  ck.aln2 = (int16_t*)(base + off[O_ALN2]) + first * T;
  ck.nm1 = base + off[O_NM1] + first * T;
  ck.nm2 = base + off[O_NM2] + first * T;
  ck.lead1 = base + off[O_LEAD1] + first * T;
  ck.lead2 = base + off[O_LEAD2] + first * T;
  ck.trail1 = base + off[O_TRAIL1] + first * T;
  ck.trail2 = base + off[O_TRAIL2] + first * T;
  ck.target_mask = (uint32_t*)(base + off[O_TARGET]) + first;
  ck.visible_mask = (uint32_t*)(base + off[O_VISIBLE]) + first;
  ck.min_sum = (uint16_t*)(base + off[O_MINSUM]) + first;
  ck.others_s1 = (uint16_t*)(base + off[O_OS1]) + first;
  ck.others_s2 = (uint16_t*)(base + off[O_OS2]) + first;
}

// n pairs plus n_blocks rna block units
int ensure_work(hlm_ctx* ctx, Slot& s, int64_t n, int64_t n_blocks = 0, int n_al = 0) {
  int U = ctx->units_per_pair, T = ctx->n_targets;
  size_t NU = (size_t)n * U + (size_t)n_blocks;
  if (NU >= (1ull << 26)) {
    ctx->err = "more than 2^26 units (mates + rna blocks) in one chunk: lower hlm_params.chunk_pairs";
    return HLM_E_ARG;
  }
  int rc;
  if ((rc = ensure_dev(ctx, s.d_units, NU * HLM_UNIT_STRIDE * 4 + 256))) return rc;
  if ((rc = ensure_dev(ctx, s.d_unit_meta, 2 * align16(NU * 2) + NU * 4 + 256))) return rc;
  if ((rc = ensure_dev(ctx, s.d_cand, (size_t)ctx->cand_cap * 8))) return rc;
  int64_t dp_cap = ctx->cand_cap / 2;  // DPs are a fraction of the candidates; overflow splits the chunk
  if ((rc = ensure_dev(ctx, s.d_dp, (size_t)dp_cap * 9 + 4096))) return rc;
  if ((rc = ensure_dev(ctx, s.d_tasks, 3 * align16(NU * T * 4) + NU * T * 8 + 256))) return rc;
  if ((rc = ensure_dev(ctx, s.d_counters, 256))) return rc;
  // de-duplication table: 4 slots per DP a chunk typically selects (about 50 per pair per round);
  // a fuller table only de-duplicates less
  int tab_bits = 16;
  while (tab_bits < 26 && ((int64_t)1 << tab_bits) < std::min<int64_t>(dp_cap * 2, (int64_t)n * 256)) tab_bits++;
  if ((rc = ensure_dev(ctx, s.d_dptab, (size_t)8 << tab_bits))) return rc;
  if ((rc = ensure_dev(ctx, s.d_prelist, (size_t)n * 4 + 256))) return rc;
  HlmChunk& ck = s.ck;
  ck.units_per_pair = U;
  ck.n_units = (int64_t)NU;
  ck.unit_planes = (uint32_t*)s.d_units.p;
  ck.unit_len = (uint16_t*)s.d_unit_meta.p;
  ck.unit_cap = (uint16_t*)((uint8_t*)s.d_unit_meta.p + align16(NU * 2));
  ck.unit_mask = (uint32_t*)((uint8_t*)s.d_unit_meta.p + 2 * align16(NU * 2));
  ck.cand = (uint64_t*)s.d_cand.p;
  ck.cand_cap = ctx->cand_cap;
  ck.dp_list = (uint32_t*)s.d_dp.p;
  ck.dp_sorted = ck.dp_list + dp_cap;
  ck.dp_hist = ck.dp_sorted + dp_cap;
  ck.dp_cursor = ck.dp_hist + 256;
  ck.dp_len = (uint8_t*)(ck.dp_cursor + 256);
  ck.dp_cap = dp_cap;
  ck.pre_list = (uint32_t*)s.d_prelist.p;
  ck.dp_tab = ctx->p.no_dp_dedup ? nullptr : (unsigned long long*)s.d_dptab.p;
  ck.dp_tab_bits = tab_bits;
  ck.counters = (unsigned long long*)s.d_counters.p;
  ck.task_emin = (uint32_t*)s.d_tasks.p;
  ck.task_emin0 = (uint32_t*)((uint8_t*)s.d_tasks.p + align16(NU * T * 4));
  ck.task_pmin = (uint32_t*)((uint8_t*)s.d_tasks.p + 2 * align16(NU * T * 4));
  ck.task_best = (unsigned long long*)((uint8_t*)s.d_tasks.p + 3 * align16(NU * T * 4));
  ck.cand_res = nullptr;
  ck.allele_best = nullptr;
  ck.n_al = n_al;
  if (n_al > 0) {
    if ((rc = ensure_dev(ctx, s.d_cres, (size_t)ctx->cand_cap * 8))) return rc;
    if ((rc = ensure_dev(ctx, s.d_albest, NU * (size_t)n_al * 8 + 256))) return rc;
    if ((rc = ensure_dev(ctx, s.d_alout, 4 * (size_t)n * n_al + 256))) return rc;
    ck.cand_res = (unsigned long long*)s.d_cres.p;
    ck.allele_best = (unsigned long long*)s.d_albest.p;
  }
  return 0;
}

void launch_stages(hlm_ctx* ctx, Slot& s, int stages, int allele_locus = -1) {
  const HlmChunk& ck = s.ck;
  int T = ctx->n_targets, sms = ctx->sms;
  cudaStream_t st = s.st;
  hlm_launch_clear_tasks(ck, T, sms, st);
  if (stages & HLM_STAGE_ALLELES) {
    // per-allele scoring of one target set: every seed candidate, every child that can hold a
    // reportable alignment, every live candidate aligned; results gathered per allele
    HlmKParams kp = ctx->kp;
    kp.allele_locus = allele_locus;
    kp.score_cap = 0;
    size_t nal = (size_t)ck.n_pairs * ck.n_al;
    uint8_t* o = (uint8_t*)s.d_alout.p;
    cudaMemsetAsync(ck.cand_res, 0xFF, (size_t)ck.cand_cap * 8, st);
    cudaMemsetAsync(ck.allele_best, 0xFF, (size_t)ck.n_units * ck.n_al * 8, st);
    cudaEventRecord(s.ev[EV_SCREEN], st);
    hlm_launch_pack(kp, ck, sms, st);
    hlm_launch_seed(ctx->ddb, kp, ck, sms, st);
    hlm_launch_snapshot(ck, 0, st);
    cudaEventRecord(s.ev[EV_SEED], st);
    hlm_launch_myers(ctx->ddb, kp, ck, T, 0, sms, st);
    cudaEventRecord(s.ev[EV_MYA], st);
    hlm_launch_expand(ctx->ddb, kp, ck, T, 2, sms, st);
    cudaEventRecord(s.ev[EV_EX0], st);
    hlm_launch_myers(ctx->ddb, kp, ck, T, 1, sms, st);
    cudaEventRecord(s.ev[EV_MYB0], st);
    hlm_launch_select(ctx->ddb, kp, ck, T, 2, sms, st);
    cudaEventRecord(s.ev[EV_SEL0], st);
    hlm_launch_dp(ctx->ddb, kp, ck, T, sms, st);
    for (int e = EV_DP0; e <= EV_DP1; e++) cudaEventRecord(s.ev[e], st);
    hlm_launch_allele_gather(ctx->ddb, kp, ck, o, o + nal, o + 2 * nal, o + 3 * nal, sms, st);
    cudaEventRecord(s.ev[EV_ASSIGN], st);
    ctx->prof.launches += 16;
    return;
  }
  if (stages & HLM_STAGE_SCREEN) hlm_launch_screen(ctx->ddb, ctx->kp, ck, sms, st);
  cudaEventRecord(s.ev[EV_SCREEN], st);
  if (stages & HLM_STAGE_SCORE) {
    size_t n_tasks = (size_t)ck.n_units * T;
    hlm_launch_pack(ctx->kp, ck, sms, st);
    hlm_launch_seed(ctx->ddb, ctx->kp, ck, sms, st);
    hlm_launch_snapshot(ck, 0, st);
    cudaEventRecord(s.ev[EV_SEED], st);
    hlm_launch_myers(ctx->ddb, ctx->kp, ck, T, 0, sms, st);
    cudaMemcpyAsync(ck.task_emin0, ck.task_emin, n_tasks * 4, cudaMemcpyDeviceToDevice, st);
    cudaEventRecord(s.ev[EV_MYA], st);
    hlm_launch_expand(ctx->ddb, ctx->kp, ck, T, 0, sms, st);
    cudaEventRecord(s.ev[EV_EX0], st);
    hlm_launch_myers(ctx->ddb, ctx->kp, ck, T, 1, sms, st);
    cudaEventRecord(s.ev[EV_MYB0], st);
    hlm_launch_select(ctx->ddb, ctx->kp, ck, T, 0, sms, st);
    cudaEventRecord(s.ev[EV_SEL0], st);
    hlm_launch_dp(ctx->ddb, ctx->kp, ck, T, sms, st);
    cudaEventRecord(s.ev[EV_DP0], st);
    hlm_launch_expand(ctx->ddb, ctx->kp, ck, T, 1, sms, st);
    cudaEventRecord(s.ev[EV_EX1], st);
    hlm_launch_myers(ctx->ddb, ctx->kp, ck, T, 2, sms, st);
    cudaEventRecord(s.ev[EV_MYB1], st);
    hlm_launch_select(ctx->ddb, ctx->kp, ck, T, 1, sms, st);
    cudaEventRecord(s.ev[EV_SEL1], st);
    hlm_launch_dp(ctx->ddb, ctx->kp, ck, T, sms, st);
    cudaEventRecord(s.ev[EV_DP1], st);
    hlm_launch_scores(ctx->kp, ck, T, sms, st);
    ctx->prof.launches += 23;
  } else {
    for (int e = EV_SEED; e <= EV_DP1; e++) cudaEventRecord(s.ev[e], st);
  }
  if (stages & HLM_STAGE_ASSIGN) {
    hlm_launch_assign(ctx->kp, ck, T, sms, st);
    ctx->prof.launches += 1;
  }
  cudaEventRecord(s.ev[EV_ASSIGN], st);
  ctx->prof.launches += 1 + ((stages & HLM_STAGE_SCREEN) ? 1 : 0);
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// wait for a slot's chunk, fold its timings / counters into the profile
int finish_slot(hlm_ctx* ctx, Slot& s, unsigned long long* counters_host) {
  if (!s.busy) return 0;
  CU(cudaEventSynchronize(s.ev[EV_D2H]));
  s.busy = false;
  hlm_profile& pr = ctx->prof;
  pr.ms_h2d += ev_ms(s.ev[EV_START], s.ev[EV_H2D]);
  pr.ms_screen += ev_ms(s.ev[EV_H2D], s.ev[EV_SCREEN]);
  pr.ms_seed += ev_ms(s.ev[EV_SCREEN], s.ev[EV_SEED]);
  pr.ms_myers += ev_ms(s.ev[EV_SEED], s.ev[EV_MYA]) + ev_ms(s.ev[EV_EX0], s.ev[EV_MYB0]) +
                 ev_ms(s.ev[EV_EX1], s.ev[EV_MYB1]);
  pr.ms_expand += ev_ms(s.ev[EV_MYA], s.ev[EV_EX0]) + ev_ms(s.ev[EV_DP0], s.ev[EV_EX1]);
  pr.ms_select += ev_ms(s.ev[EV_MYB0], s.ev[EV_SEL0]) + ev_ms(s.ev[EV_MYB1], s.ev[EV_SEL1]);
  pr.ms_dp += ev_ms(s.ev[EV_SEL0], s.ev[EV_DP0]) + ev_ms(s.ev[EV_SEL1], s.ev[EV_DP1]);
  pr.ms_assign += ev_ms(s.ev[EV_DP1], s.ev[EV_ASSIGN]);
  pr.ms_d2h += ev_ms(s.ev[EV_ASSIGN], s.ev[EV_D2H]);
  float t = ev_ms(ctx->run_start, s.ev[EV_D2H]);
  if (t > pr.ms_total) pr.ms_total = t;
  if (ctx->tracing) {
    ctx->trace.push_back((float)(&s - ctx->slot));
    for (int e = 0; e < EV_COUNT; e++) ctx->trace.push_back(ev_ms(ctx->run_start, s.ev[e]));
  }
  if (counters_host) {
    pr.n_candidates += (int64_t)counters_host[11];
    pr.n_seed_candidates += (int64_t)counters_host[12];
    pr.n_kept += (int64_t)counters_host[3];
    pr.n_dp_cells += (int64_t)counters_host[4];
    pr.n_myers_cols += (int64_t)counters_host[5];
    pr.n_dp += (int64_t)counters_host[6];
    if (counters_host[2]) {
      ctx->err = "candidate capacity exceeded for a single pair: raise hlm_params.cand_capacity";
      return HLM_E_OVERFLOW;  // the caller re-runs the chunk in two halves
    }
    if (counters_host[7]) {
      ctx->err = "a trimmed read is longer than HLM_MAX_READ (184)";
      return HLM_E_TOOLONG;
    }
  }
  return 0;
}

struct HostIO {
  const hlm_batch* in = nullptr;
  const hlm_screen_out* scr_in = nullptr;  // provided screen results (score / assign stages)
  const hlm_score_out* sc_in = nullptr;    // provided scores (assign stage)
  hlm_screen_out* scr = nullptr;
  hlm_score_out* sc = nullptr;
  hlm_assign_out* asg = nullptr;
  hlm_allele_out* al = nullptr;  // hlm_score_alleles
  int al_locus = -1;
};

template <class Tp>
void copy_out(Tp* dst, const uint8_t* base, size_t off, int64_t first, int64_t n, int per) {
  if (dst) memcpy(dst + first * per, base + off, sizeof(Tp) * (size_t)n * per);
}

void scatter_results(hlm_ctx* ctx, Slot& s, const HostIO& io) {
  const uint8_t* b = (const uint8_t*)s.h_out.p;
  const size_t* o = s.out_off;
  int T = ctx->n_targets;
  int64_t f = s.first, n = s.n;
  if (io.scr) {
    copy_out(io.scr->flags, b, o[O_FLAGS], f, n, 1);
    copy_out(io.scr->trim_lo1, b, o[O_LO1], f, n, 1);
    copy_out(io.scr->trim_len1, b, o[O_LEN1], f, n, 1);
    copy_out(io.scr->trim_lo2, b, o[O_LO2], f, n, 1);
    copy_out(io.scr->trim_len2, b, o[O_LEN2], f, n, 1);
    copy_out(io.scr->locus_mask, b, o[O_MASK], f, n, 1);
  }
  if (io.sc) {
    // the two score tables are most of the result bytes: copied side by side on large chunks
    if (io.sc->s2 && (size_t)n * T >= ((size_t)1 << 19)) {
      std::thread t2([&] { copy_out(io.sc->s2, b, o[O_S2], f, n, T); });
      copy_out(io.sc->s1, b, o[O_S1], f, n, T);
      t2.join();
    } else {
      copy_out(io.sc->s1, b, o[O_S1], f, n, T);
      copy_out(io.sc->s2, b, o[O_S2], f, n, T);
    }
    copy_out(io.sc->aln1, b, o[O_ALN1], f, n, T);
    copy_out(io.sc->aln2, b, o[O_ALN2], f, n, T);
    copy_out(io.sc->nm1, b, o[O_NM1], f, n, T);
    copy_out(io.sc->nm2, b, o[O_NM2], f, n, T);
    copy_out(io.sc->lead1, b, o[O_LEAD1], f, n, T);
    copy_out(io.sc->lead2, b, o[O_LEAD2], f, n, T);
    copy_out(io.sc->trail1, b, o[O_TRAIL1], f, n, T);
    copy_out(io.sc->trail2, b, o[O_TRAIL2], f, n, T);
  }
  if (io.asg) {
    copy_out(io.asg->target_mask, b, o[O_TARGET], f, n, 1);
    copy_out(io.asg->visible_mask, b, o[O_VISIBLE], f, n, 1);
    copy_out(io.asg->min_sum, b, o[O_MINSUM], f, n, 1);
    copy_out(io.asg->others_s1, b, o[O_OS1], f, n, 1);
    copy_out(io.asg->others_s2, b, o[O_OS2], f, n, 1);
  }
}

// the rna block list: CSR offsets ascending, both arrays present
int check_blocks(hlm_ctx* ctx, const hlm_batch* in) {
  if (!in->rna_block_off) return 0;
  for (int64_t i = 0; i < in->n_pairs; i++)
    if (in->rna_block_off[i + 1] < in->rna_block_off[i]) {
      ctx->err = "rna_block_off is not ascending";
      return HLM_E_ARG;
    }
  if (in->n_pairs > 0 && in->rna_block_off[in->n_pairs] > in->rna_block_off[0] && !in->rna_blocks) {
    ctx->err = "rna_block_off without rna_blocks";
    return HLM_E_ARG;
  }
  return 0;
}

int check_lengths(hlm_ctx* ctx, const hlm_batch* in) {
  uint64_t longest = 0;
  for (int64_t i = 0; i < in->n_pairs; i++) {
    longest = std::max(longest, in->offs1[i + 1] - in->offs1[i]);
    if (ctx->p.paired && in->offs2) longest = std::max(longest, in->offs2[i + 1] - in->offs2[i]);
    if (in->offs1[i + 1] < in->offs1[i] || in->offs1[i + 1] - in->offs1[i] > HLM_RAW_MAX) {
      ctx->err = "read 1 longer than 256 bases (or offsets not ascending)";
      return HLM_E_TOOLONG;
    }
    if (ctx->p.paired && (in->offs2[i + 1] < in->offs2[i] || in->offs2[i + 1] - in->offs2[i] > HLM_RAW_MAX)) {
      ctx->err = "read 2 longer than 256 bases (or offsets not ascending)";
      return HLM_E_TOOLONG;
    }
  }
  (void)longest;
  return 0;
}

// host-buffer pipeline over chunks
int run_host_unguarded(hlm_ctx* ctx, const HostIO& io, int stages);
int run_host(hlm_ctx* ctx, const HostIO& io, int stages) {
  try {  // no exception crosses the C ABI
    return run_host_unguarded(ctx, io, stages);
  } catch (const std::exception& x) {
    ctx->err = std::string("libhlamap: ") + x.what();
    return HLM_E_NOMEM;
  }
}
int run_host_unguarded(hlm_ctx* ctx, const HostIO& io, int stages) {
  const hlm_batch* in = io.in;
  if (!in || in->n_pairs < 0) {
    ctx->err = "bad batch";
    return HLM_E_ARG;
  }
  bool need_reads = stages & (HLM_STAGE_SCREEN | HLM_STAGE_SCORE | HLM_STAGE_ALLELES);
  if (need_reads) {
    if (!in->bases1 || !in->offs1 || ((stages & HLM_STAGE_SCREEN) && !in->quals1) ||
        (ctx->p.paired && (!in->bases2 || !in->offs2 || ((stages & HLM_STAGE_SCREEN) && !in->quals2)))) {
      ctx->err = "batch is missing read arrays";
      return HLM_E_ARG;
    }
    int rc = check_lengths(ctx, in);
    if (rc) return rc;
    if (ctx->p.rna && (rc = check_blocks(ctx, in))) return rc;
  }
  CU(cudaSetDevice(ctx->device));
  memset(&ctx->prof, 0, sizeof ctx->prof);
  int T = ctx->n_targets;
  bool paired = ctx->p.paired;
  int64_t N = in->n_pairs, C = ctx->chunk_pairs;
  int n_al = 0;
  if (stages & HLM_STAGE_ALLELES) {
    // units x alleles x 8 bytes of per-allele minima per chunk: keep it near 1 GB
    n_al = io.al->n_alleles;
    C = std::max<int64_t>(32, std::min<int64_t>(C, ((int64_t)1 << 26) / std::max(1, n_al)));
  }
  CU(cudaEventRecord(ctx->run_start, ctx->slot[0].st));
  for (int i = 1; i < HLM_MAX_SLOTS; i++) CU(cudaStreamWaitEvent(ctx->slot[i].st, ctx->run_start, 0));
  // stage + copy + launch one chunk on a slot (asynchronous)
  auto submit = [&](Slot& s, int64_t first, int64_t n) -> int {
    int rc = 0;
    // ---- stage inputs into pinned memory
    size_t b1 = 0, b2 = 0;
    uint64_t f1 = 0, f2 = 0;
    if (need_reads) {
      f1 = in->offs1[first];
      b1 = (size_t)(in->offs1[first + n] - f1);
      if (paired) {
        f2 = in->offs2[first];
        b2 = (size_t)(in->offs2[first + n] - f2);
      }
    }
    bool q = stages & HLM_STAGE_SCREEN;
    size_t io_off[16], o = 0;
    auto sec = [&](int i, size_t bytes) { io_off[i] = o; o += align16(bytes + 32); };
    sec(0, b1);                       // bases1 (16 bytes of lead padding inside)
    sec(1, q ? b1 : 0);               // quals1
    sec(2, b2);                       // bases2
    sec(3, q ? b2 : 0);               // quals2
    sec(4, (n + 1) * 8);              // offs1
    sec(5, paired ? (n + 1) * 8 : 0); // offs2
    sec(6, in->size_override1 ? n * 4 : 0);
    sec(7, in->rna_split1 ? n * 2 : 0);
    sec(8, in->rna_split2 ? n * 2 : 0);
    bool blocks = ctx->p.rna && in->rna_block_off && (stages & HLM_STAGE_SCORE);
    uint64_t bf = blocks ? in->rna_block_off[first] : 0;
    int64_t nb = blocks ? (int64_t)(in->rna_block_off[first + n] - bf) : 0;
    sec(9, blocks ? (n + 1) * 8 : 0);
    sec(10, (size_t)nb * sizeof(hlm_rna_block));
    size_t in_bytes = o;
    if ((rc = ensure_pinned(ctx, s.h_in, in_bytes))) return rc;
    if ((rc = ensure_dev(ctx, s.d_in, in_bytes))) return rc;
    size_t out_bytes = out_layout(n, T, s.out_off);
    s.out_bytes = out_bytes;
    if ((rc = ensure_pinned(ctx, s.h_out, out_bytes))) return rc;
    if ((rc = ensure_dev(ctx, s.d_out, out_bytes))) return rc;
    if ((rc = ensure_work(ctx, s, n, nb, n_al))) return rc;
    uint8_t* h = (uint8_t*)s.h_in.p;
    uint8_t* d = (uint8_t*)s.d_in.p;
    if (blocks) {
      memcpy(h + io_off[9], in->rna_block_off + first, (n + 1) * 8);
      if (nb) memcpy(h + io_off[10], in->rna_blocks + bf, (size_t)nb * sizeof(hlm_rna_block));
    }
    if (need_reads) {
      // the four read arrays of a large chunk are staged by four host threads: on screen-heavy
      // input (config 3: the GPU needs 2.6 ms per 65 536-pair chunk) one thread's memcpy into the
      // pinned buffer was the end-to-end limiter
      auto c0 = [&] { memcpy(h + io_off[0] + 16, in->bases1 + f1, b1); };
      auto c1 = [&] { if (q) memcpy(h + io_off[1] + 16, in->quals1 + f1, b1); };
      auto c2 = [&] { if (paired) memcpy(h + io_off[2] + 16, in->bases2 + f2, b2); };
      auto c3 = [&] { if (paired && q) memcpy(h + io_off[3] + 16, in->quals2 + f2, b2); };
      if (b1 + b2 >= ((size_t)4 << 20)) {
        std::thread t1(c1), t2(c2), t3(c3);
        c0();
        t1.join();
        t2.join();
        t3.join();
      } else {
        c0(); c1(); c2(); c3();
      }
      memcpy(h + io_off[4], in->offs1 + first, (n + 1) * 8);
      if (paired) memcpy(h + io_off[5], in->offs2 + first, (n + 1) * 8);
    }
    if (in->size_override1) memcpy(h + io_off[6], in->size_override1 + first, n * 4);
    if (in->rna_split1) memcpy(h + io_off[7], in->rna_split1 + first, n * 2);
    if (in->rna_split2) memcpy(h + io_off[8], in->rna_split2 + first, n * 2);
    HlmChunk& ck = s.ck;
    ck.n_pairs = n;
    // offsets stay absolute: bias the base pointers instead
    ck.bases1 = d + io_off[0] + 16 - f1;
    ck.quals1 = d + io_off[1] + 16 - f1;
    ck.bases2 = d + io_off[2] + 16 - f2;
    ck.quals2 = d + io_off[3] + 16 - f2;
    ck.offs1 = (const uint64_t*)(d + io_off[4]);
    ck.offs2 = (const uint64_t*)(d + io_off[5]);
    ck.size_override1 = in->size_override1 ? (const int32_t*)(d + io_off[6]) : nullptr;
    ck.rna_split1 = in->rna_split1 ? (const uint16_t*)(d + io_off[7]) : nullptr;
    ck.rna_split2 = in->rna_split2 ? (const uint16_t*)(d + io_off[8]) : nullptr;
    ck.blk_off = blocks ? (const uint64_t*)(d + io_off[9]) : nullptr;
    ck.blk = blocks ? (const hlm_rna_block*)(d + io_off[10]) - bf : nullptr;
    ck.blk_first = (int64_t)bf;
    ck.n_blocks = nb;
    bind_outputs(ck, (uint8_t*)s.d_out.p, s.out_off, 0, T);
    s.first = first;
    s.n = n;
    cudaStream_t st = s.st;
    CU(cudaEventRecord(s.ev[EV_START], st));
    CU(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
    ctx->prof.h2d_bytes += (int64_t)in_bytes;
    // provided stage results (stage-wise entry points)
    uint8_t* ho = (uint8_t*)s.h_out.p;
    uint8_t* dout = (uint8_t*)s.d_out.p;
    if (!(stages & HLM_STAGE_SCREEN) && io.scr_in) {
      const hlm_screen_out* si = io.scr_in;
      memcpy(ho + s.out_off[O_FLAGS], si->flags + first, n);
      memcpy(ho + s.out_off[O_LO1], si->trim_lo1 + first, 2 * n);
      memcpy(ho + s.out_off[O_LEN1], si->trim_len1 + first, 2 * n);
      if (si->trim_lo2) memcpy(ho + s.out_off[O_LO2], si->trim_lo2 + first, 2 * n);
      else memset(ho + s.out_off[O_LO2], 0, 2 * n);
      if (si->trim_len2) memcpy(ho + s.out_off[O_LEN2], si->trim_len2 + first, 2 * n);
      else memset(ho + s.out_off[O_LEN2], 0, 2 * n);
      memcpy(ho + s.out_off[O_MASK], si->locus_mask + first, 4 * n);
      size_t bytes = s.out_off[O_MASK] + 4 * n;
      CU(cudaMemcpyAsync(dout, ho, bytes, cudaMemcpyHostToDevice, st));
      ctx->prof.h2d_bytes += (int64_t)bytes;
    }
    if (!(stages & HLM_STAGE_SCORE) && (stages & HLM_STAGE_ASSIGN) && io.sc_in) {
      memcpy(ho + s.out_off[O_S1], io.sc_in->s1 + first * T, 2 * n * T);
      if (io.sc_in->s2) memcpy(ho + s.out_off[O_S2], io.sc_in->s2 + first * T, 2 * n * T);
      else memset(ho + s.out_off[O_S2], 0, 2 * n * T);
      size_t a = s.out_off[O_S1], e = s.out_off[O_S2] + 2 * n * T;
      CU(cudaMemcpyAsync(dout + a, ho + a, e - a, cudaMemcpyHostToDevice, st));
      ctx->prof.h2d_bytes += (int64_t)(e - a);
    }
    CU(cudaEventRecord(s.ev[EV_H2D], st));
    if (ctx->stagger_ev >= 0 && ctx->last_submitted && ctx->last_submitted != &s)
      CU(cudaStreamWaitEvent(st, ctx->last_submitted->ev[ctx->stagger_ev], 0));
    ctx->last_submitted = &s;
    launch_stages(ctx, s, stages, io.al_locus);
    CU(cudaMemcpyAsync((uint8_t*)s.d_out.p + s.out_off[O_COUNTERS], ck.counters, 8 * HLM_N_COUNTERS,
                       cudaMemcpyDeviceToDevice, st));
    {  // copy back only the result groups the caller asked for (+ the counters)
      auto d2h = [&](int a, int b) {  // sections [a, b)
        size_t lo = s.out_off[a], hi = (b < O_COUNT) ? s.out_off[b] : out_bytes;
        cudaMemcpyAsync((uint8_t*)s.h_out.p + lo, (uint8_t*)s.d_out.p + lo, hi - lo,
                        cudaMemcpyDeviceToHost, st);
        ctx->prof.d2h_bytes += (int64_t)(hi - lo);
      };
      if (io.scr) d2h(O_FLAGS, O_S1);
      if (io.sc) {
        d2h(O_S1, O_ALN1);
        if (io.sc->aln1 || io.sc->aln2 || io.sc->nm1 || io.sc->nm2 || io.sc->lead1 || io.sc->lead2 ||
            io.sc->trail1 || io.sc->trail2)
          d2h(O_ALN1, O_TARGET);
      }
      if (io.asg) d2h(O_TARGET, O_COUNTERS);
      d2h(O_COUNTERS, O_COUNT);
    }
    CU(cudaEventRecord(s.ev[EV_D2H], st));
    s.busy = true;
    ctx->prof.bases_bytes += (int64_t)(q ? 2 : 1) * (int64_t)(b1 + b2);
    return 0;
  };
  auto counters_of = [](Slot& s) { return (unsigned long long*)((uint8_t*)s.h_out.p + s.out_off[O_COUNTERS]); };
  // wait for a slot; a chunk whose candidates overflowed the device lists is re-run in halves
  std::function<int(Slot&)> complete = [&](Slot& s) -> int {
    if (!s.busy) return 0;
    int rc = finish_slot(ctx, s, counters_of(s));
    if (rc == HLM_E_OVERFLOW && s.n > 1) {
      int64_t first = s.first, n = s.n, h = n / 2;
      ctx->prof.n_overflow_splits++;
      if ((rc = submit(s, first, h))) return rc;
      if ((rc = complete(s))) return rc;
      if ((rc = submit(s, first + h, n - h))) return rc;
      return complete(s);
    }
    if (rc == HLM_E_OVERFLOW) rc = HLM_E_NOMEM;
    if (rc == 0) scatter_results(ctx, s, io);
    if (rc == 0 && io.al) {  // per-allele bytes straight into the caller's rows (synchronous copies)
      size_t nal = (size_t)s.n * n_al, at = (size_t)s.first * n_al;
      const uint8_t* o = (const uint8_t*)s.d_alout.p;
      uint8_t* dst[4] = {io.al->nm1, io.al->clip1, io.al->nm2, io.al->clip2};
      for (int k = 0; k < 4; k++)
        if (dst[k]) {
          CU(cudaMemcpy(dst[k] + at, o + k * nal, nal, cudaMemcpyDeviceToHost));
          ctx->prof.d2h_bytes += (int64_t)nal;
        }
    }
    return rc;
  };
  int rc = 0;
  int64_t k = 0;
  for (int64_t first = 0; first < N && rc == 0; first += C, k++) {
    Slot& s = ctx->slot[k % ctx->n_slots];
    int64_t n = std::min(C, N - first);
    if ((rc = complete(s))) break;  // drain the slot's previous chunk
    rc = submit(s, first, n);
  }
  for (int i = 0; i < ctx->n_slots; i++) {
    int r2 = complete(ctx->slot[(k + i) % ctx->n_slots]);
    if (rc == 0) rc = r2;
  }
  ctx->prof.n_pairs = N;
  cudaError_t e = cudaGetLastError();
  if (rc == 0 && e != cudaSuccess) {
    ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e);
    rc = HLM_E_CUDA;
  }
  return rc;
}

}  // namespace

// accessors for the host file pipeline (hlm_files.cpp)
void hlm_ctx_set_error(hlm_ctx* ctx, const std::string& msg) { ctx->err = msg; }
const hlm_db* hlm_ctx_db(const hlm_ctx* ctx) { return ctx->db; }
const hlm_params* hlm_ctx_params(const hlm_ctx* ctx) { return &ctx->p; }

// =====================================================================================
extern "C" {

int hlm_abi_version(void) { return HLM_ABI_VERSION; }

void hlm_params_default(hlm_params* p, int rna) {
  memset(p, 0, sizeof *p);
  p->struct_size = (int32_t)sizeof *p;
  p->rna = rna ? 1 : 0;
  p->paired = 1;
  p->screen_variant = HLM_SCREEN_FASTQ;
  p->v_size = 0;
  p->minhit = 3;                        // preselect.cpp:26
  p->mm_max = 20;                       // main.cpp:77
  p->clip_mode = HLM_CLIP_LEAD_ONLY;
  p->tolerance = rna ? 0.08 : 0.05;     // main.cpp:80, map_rna.cpp:116
  p->mtrim_error = (double)0.08f;       // main.cpp:81: `double v_mtrim_error = 0.08f;`
  p->chunk_pairs = 0;
  p->skip_screen = 0;
  p->cand_capacity = 0;
  p->single_slot = 0;
  p->score_cap = 0;
  p->no_dp_dedup = 0;
  p->reserved1 = 0;
}

hlm_db* hlm_db_open(const char* db_dir, int rna, char* err, size_t errlen) {
  std::string e;
  hlm_db* db = nullptr;
  try {  // no exception crosses the C ABI
    db = db_dir ? hlm_db_build(db_dir, rna, e) : nullptr;
  } catch (const std::exception& x) {
    e = std::string("hlm_db_open: ") + x.what();
  }
  if (!db && err && errlen) snprintf(err, errlen, "%s", e.empty() ? "bad db_dir" : e.c_str());
  return db;
}

void hlm_db_close(hlm_db* db) {
  if (!db) return;
  for (auto& d : db->devs) {
    cudaSetDevice(d.device);
    for (void* p : d.ptrs)
      if (p) cudaFree(p);
  }
  delete db;
}
int hlm_db_n_loci(const hlm_db* db) { return db ? db->n_loci : 0; }
const char* hlm_db_locus_name(const hlm_db* db, int i) {
  return (db && i >= 0 && i <= db->n_loci) ? db->names[i].c_str() : nullptr;
}
int hlm_db_size(const hlm_db* db) { return db ? db->v_size : 0; }
int hlm_db_n_alleles(const hlm_db* db, int locus) {
  return (db && locus >= 0 && locus <= db->n_loci) ? (int)db->allele_names[locus].size() : 0;
}
const char* hlm_db_allele_name(const hlm_db* db, int locus, int allele) {
  if (!db || locus < 0 || locus > db->n_loci) return nullptr;
  const auto& v = db->allele_names[locus];
  return (allele >= 0 && (size_t)allele < v.size()) ? v[allele].c_str() : nullptr;
}
int64_t hlm_db_stat(const hlm_db* db, const char* key) {
  if (!db || !key) return -1;
  std::string k(key);
  if (k == "alleles") return db->n_alleles;
  if (k == "bases") return db->n_bases;
  if (k == "tiles") return (int64_t)(db->tiles.size() / HLM_TILE_WORDS);
  if (k == "tiles_raw") return db->n_tiles_raw;
  if (k == "kmers") return db->n_kmers;
  if (k == "kmers_skipped") return db->n_kmers_skipped;
  if (k == "kmers_odd") return (int64_t)db->odd_masks.size();
  if (k == "kt_slots") return (int64_t)db->kt.size();
  if (k == "bloom_words") return (int64_t)db->kb.size();
  if (k == "seed_entries") return (int64_t)db->seed_ent.size();
  if (k == "reps") return db->n_reps;
  if (k == "children") return (int64_t)db->child_ent.size();
  return -1;
}

static int upload_db(hlm_ctx* ctx, hlm_db* db) {
  std::lock_guard<std::mutex> guard(db->devs_mutex);
  for (auto& d : db->devs)
    if (d.device == ctx->device) {
      ctx->ddb = d.view;
      d.refs++;
      return 0;
    }
  hlm_db::Dev dv;
  dv.device = ctx->device;
  dv.refs = 1;
  for (auto& p : dv.ptrs) p = nullptr;
  auto up = [&](int slot, const void* src, size_t bytes, const void** out) -> int {
    void* p = nullptr;
    if (!cuda_ok(ctx, cudaMalloc(&p, bytes + 256), "cudaMalloc(db)")) return HLM_E_NOMEM;
    if (!cuda_ok(ctx, cudaMemset(p, 0, bytes + 256), "cudaMemset(db)")) return HLM_E_CUDA;
    if (bytes && !cuda_ok(ctx, cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy(db)"))
      return HLM_E_CUDA;
    dv.ptrs[slot] = p;
    *out = p;
    return 0;
  };
  HlmDevDB v;
  memset(&v, 0, sizeof v);
  int rc = 0;
  if (!rc) rc = up(0, db->kt.data(), db->kt.size() * 8, (const void**)&v.kt);
  if (!rc) rc = up(1, db->mask_tab.data(), db->mask_tab.size() * 4, (const void**)&v.mask_tab);
  if (!rc) rc = up(2, db->tiles.data(), db->tiles.size() * 4, (const void**)&v.tiles);
  if (!rc) rc = up(3, db->seed_off.data(), db->seed_off.size() * 4, (const void**)&v.seed_off);
  if (!rc) rc = up(4, db->seed_ent.data(), db->seed_ent.size() * 4, (const void**)&v.seed_ent);
  if (!rc) rc = up(5, db->tile_diff.data(), db->tile_diff.size() * 4, (const void**)&v.tile_diff);
  if (!rc) rc = up(6, db->child_off.data(), db->child_off.size() * 4, (const void**)&v.child_off);
  if (!rc) rc = up(7, db->child_ent.data(), db->child_ent.size() * 4, (const void**)&v.child_ent);
  if (!rc) rc = up(8, db->child_diff.data(), db->child_diff.size() * 4, (const void**)&v.child_diff);
  if (!rc) rc = up(9, db->odd_keys.data(), db->odd_keys.size(), (const void**)&v.odd_keys);
  if (!rc) rc = up(10, db->odd_masks.data(), db->odd_masks.size() * 4, (const void**)&v.odd_masks);
  if (!rc) rc = up(11, db->kb.data(), db->kb.size() * 8, (const void**)&v.kb);
  if (!rc) rc = up(12, db->tile_al_off.data(), db->tile_al_off.size() * 4, (const void**)&v.tile_al_off);
  if (!rc) rc = up(13, db->tile_al_ent.data(), db->tile_al_ent.size() * 4, (const void**)&v.tile_al_ent);
  v.kb_words = (uint32_t)db->kb.size();
  v.n_odd = (int)db->odd_masks.size();
  if (rc) {  // nothing half-uploaded stays behind
    for (void* q : dv.ptrs)
      if (q) cudaFree(q);
    return rc;
  }
  v.kt_bits = db->kt_bits;
  v.n_loci = db->n_loci;
  for (int g = 0; g < db->n_loci + 2; g++) v.tile_start[g] = db->tile_start[g];
  for (int g = db->n_loci + 2; g < HLM_MAX_LOCI + 3; g++) v.tile_start[g] = 0xFFFFFFFFu;
  dv.view = v;
  db->devs.push_back(dv);
  ctx->ddb = v;
  return 0;
}

hlm_ctx* hlm_ctx_create(const hlm_db* db, int cuda_device, const hlm_params* params, char* err,
                        size_t errlen) {
  auto fail = [&](const std::string& m) -> hlm_ctx* {
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
    return nullptr;
  };
  if (!db) return fail("null db");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("libhlamap needs a CUDA device: there is no CPU path (no device found)");
  if (cuda_device < 0 || cuda_device >= ndev) return fail("cuda_device out of range");
  hlm_ctx* ctx = new hlm_ctx();
  ctx->db = db;
  ctx->device = cuda_device;
  if (params) {
    memcpy(&ctx->p, params, std::min((size_t)params->struct_size, sizeof(hlm_params)));
    if ((size_t)params->struct_size < sizeof(hlm_params)) {
      hlm_params d;
      hlm_params_default(&d, params->rna);
      memcpy((char*)&ctx->p + params->struct_size, (char*)&d + params->struct_size,
             sizeof(hlm_params) - params->struct_size);
    }
  } else {
    hlm_params_default(&ctx->p, db->rna);
  }
  hlm_params& p = ctx->p;
  if (p.rna != db->rna) {
    delete ctx;
    return fail("params.rna does not match the database kind");
  }
  if (p.clip_mode < HLM_CLIP_LEAD_ONLY || p.clip_mode > HLM_CLIP_NONE) {
    delete ctx;
    return fail("params.clip_mode must be HLM_CLIP_LEAD_ONLY, HLM_CLIP_BOTH or HLM_CLIP_NONE");
  }
  if (p.mm_max < 0 || p.mm_max > 100 || p.minhit < 0 || p.tolerance < 0) {
    delete ctx;
    return fail("params out of range (0 <= mm_max <= 100; the alignment band is fixed at +-20)");
  }
  if (cudaSetDevice(cuda_device) != cudaSuccess) {
    delete ctx;
    return fail("cudaSetDevice failed");
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cuda_device);
  ctx->sms = prop.multiProcessorCount;
  hlm_kernels_init();
  ctx->n_targets = db->n_loci + 1;
  ctx->units_per_pair = (p.paired ? 2 : 1) * (p.rna ? 3 : 1);
  ctx->n_slots = p.single_slot ? 1 : 3;
  ctx->tracing = getenv("HLM_TRACE") != nullptr;
  if (const char* e = getenv("HLM_STAGGER")) {  // measurement knob: stage event index, -1 = off
    int v = atoi(e);
    if (v >= -1 && v < EV_COUNT) ctx->stagger_ev = v;
  }
  if (const char* e = getenv("HLM_SLOTS")) {  // measurement knob (tools/grid_sweep.py)
    int v = atoi(e);
    if (!p.single_slot && v >= 1 && v <= HLM_MAX_SLOTS) ctx->n_slots = v;
  }
  ctx->chunk_pairs = p.chunk_pairs > 0 ? p.chunk_pairs : 65536;
  ctx->cand_cap = p.cand_capacity > 0 ? p.cand_capacity : ctx->chunk_pairs * 2048;
  if (ctx->cand_cap >= (1ll << 32)) ctx->cand_cap = (1ll << 32) - 1;
  if (ctx->chunk_pairs * ctx->units_per_pair >= (1ll << 26)) {
    delete ctx;
    return fail("chunk_pairs too large for the 26-bit unit id");
  }
  HlmKParams& kp = ctx->kp;
  memset(&kp, 0, sizeof kp);
  kp.rna = p.rna;
  kp.paired = p.paired;
  kp.screen_variant = p.screen_variant;
  kp.v_size = p.v_size > 0 ? p.v_size : db->v_size;
  kp.minhit = p.minhit;
  kp.mm_max = p.mm_max;
  kp.clip_mode = p.clip_mode;
  kp.skip_screen = p.skip_screen;
  kp.score_cap = (p.score_cap && !p.rna) ? 1 : 0;
  kp.allele_locus = -1;
  kp.tolerance = p.tolerance;
  kp.n_loci = db->n_loci;
  kp.one = 1;
  // mtrim's test, evaluated once per quality byte with the reference's own expression
  // (functions.cpp:99-103): P = pow(10, phred / -10) <= v_mtrim_error, phred = char - 33
  for (int c = 0; c < 256; c++) {
    double chr_phred = (double)((int)(signed char)c - 33);
    double P = pow((double)10, (chr_phred / (double)(-10)));
    if (P <= p.mtrim_error) kp.good_q[c >> 5] |= 1u << (c & 31);
  }
  // P falls with the byte value, so the good bytes are normally one run thr .. 127 (bytes >= 128
  // are negative chars in the reference: never good): the kernel then tests eight bytes at once
  kp.q_thr = -1;
  {
    int thr = 0;
    while (thr < 128 && !((kp.good_q[thr >> 5] >> (thr & 31)) & 1u)) thr++;
    bool run = thr < 128;
    for (int c = 0; c < 256 && run; c++)
      if ((((kp.good_q[c >> 5] >> (c & 31)) & 1u) != 0) != (c >= thr && c < 128)) run = false;
    if (run) kp.q_thr = thr;
  }
  if (upload_db(ctx, const_cast<hlm_db*>(db)) != 0) {
    std::string m = ctx->err;
    delete ctx;
    return fail(m);
  }
  for (int i = 0; i < HLM_MAX_SLOTS; i++) {
    cudaStreamCreateWithFlags(&ctx->slot[i].st, cudaStreamNonBlocking);
    for (int e = 0; e < EV_COUNT; e++) cudaEventCreate(&ctx->slot[i].ev[e]);
    memset(&ctx->slot[i].ck, 0, sizeof(HlmChunk));
  }
  cudaEventCreate(&ctx->run_start);
  memset(&ctx->prof, 0, sizeof ctx->prof);
  return ctx;
}

void hlm_ctx_destroy(hlm_ctx* ctx) {
  if (!ctx) return;
  if (ctx->pending.valid()) ctx->pending.get();
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < HLM_MAX_SLOTS; i++) {
    Slot& s = ctx->slot[i];
    for (Buf* b : {&s.d_in, &s.d_out, &s.d_units, &s.d_unit_meta, &s.d_cand, &s.d_dp, &s.d_tasks, &s.d_counters,
                   &s.d_cres, &s.d_albest, &s.d_alout, &s.d_dptab, &s.d_prelist})
      if (b->p) cudaFree(b->p);
    for (Buf* b : {&s.h_in, &s.h_out})
      if (b->p) cudaFreeHost(b->p);
    for (int e = 0; e < EV_COUNT; e++) cudaEventDestroy(s.ev[e]);
    if (s.st) cudaStreamDestroy(s.st);
  }
  cudaEventDestroy(ctx->run_start);
  hlm_db* db = const_cast<hlm_db*>(ctx->db);
  {
    std::lock_guard<std::mutex> guard(db->devs_mutex);
    for (auto& d : db->devs)
      if (d.device == ctx->device) d.refs--;
  }
  delete ctx;
}

const char* hlm_last_error(const hlm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int hlm_screen_trim(hlm_ctx* ctx, const hlm_batch* in, hlm_screen_out* out) {
  if (!ctx || !in || !out) return HLM_E_ARG;
  HostIO io;
  io.in = in;
  io.scr = out;
  return run_host(ctx, io, HLM_STAGE_SCREEN);
}

int hlm_score(hlm_ctx* ctx, const hlm_batch* in, const hlm_screen_out* scr, hlm_score_out* out) {
  if (!ctx || !in || !scr || !out) return HLM_E_ARG;
  HostIO io;
  io.in = in;
  io.scr_in = scr;
  io.sc = out;
  return run_host(ctx, io, HLM_STAGE_SCORE);
}

int hlm_score_alleles(hlm_ctx* ctx, const hlm_batch* in, const hlm_screen_out* scr, int locus,
                      hlm_allele_out* out) {
  if (!ctx || !in || !scr || !out) return HLM_E_ARG;
  if (locus < 0 || locus > ctx->db->n_loci) {
    ctx->err = "hlm_score_alleles: locus out of range (0 .. n_loci, n_loci = others)";
    return HLM_E_ARG;
  }
  if (out->n_alleles != (int32_t)ctx->db->allele_names[locus].size() || out->n_alleles <= 0) {
    ctx->err = "hlm_score_alleles: out->n_alleles must equal hlm_db_n_alleles(db, locus)";
    return HLM_E_ARG;
  }
  if (!out->nm1 || (ctx->p.paired && !out->nm2)) {
    ctx->err = "hlm_score_alleles: nm arrays missing";
    return HLM_E_ARG;
  }
  HostIO io;
  io.in = in;
  io.scr_in = scr;
  io.al = out;
  io.al_locus = locus;
  return run_host(ctx, io, HLM_STAGE_ALLELES);
}

int hlm_assign(hlm_ctx* ctx, const hlm_batch* in, const hlm_screen_out* scr, const hlm_score_out* sc,
               hlm_assign_out* out) {
  if (!ctx || !in || !scr || !sc || !out) return HLM_E_ARG;
  HostIO io;
  io.in = in;
  io.scr_in = scr;
  io.sc_in = sc;
  io.asg = out;
  return run_host(ctx, io, HLM_STAGE_ASSIGN);
}

int hlm_run(hlm_ctx* ctx, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc,
            hlm_assign_out* asg) {
  if (!ctx || !in) return HLM_E_ARG;
  HostIO io;
  io.in = in;
  io.scr = scr;
  io.sc = sc;
  io.asg = asg;
  return run_host(ctx, io, HLM_STAGE_SCREEN | HLM_STAGE_SCORE | HLM_STAGE_ASSIGN);
}

// asynchronous form: the pipeline runs on a worker thread, the caller overlaps its own host work
// (parsing the next batch, writing the previous one) and collects the return code with hlm_wait.
// The arrays named by the arguments must stay valid and untouched until hlm_wait returns; the
// small argument structs themselves are copied.
int hlm_run_async(hlm_ctx* ctx, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc,
                  hlm_assign_out* asg) {
  if (!ctx || !in) return HLM_E_ARG;
  if (ctx->pending.valid()) {
    ctx->err = "hlm_run_async: a run is already in flight (call hlm_wait first)";
    return HLM_E_ARG;
  }
  hlm_batch b = *in;
  bool hs = scr != nullptr, hc = sc != nullptr, ha = asg != nullptr;
  hlm_screen_out s1{};
  hlm_score_out s2{};
  hlm_assign_out s3{};
  if (hs) s1 = *scr;
  if (hc) s2 = *sc;
  if (ha) s3 = *asg;
  ctx->pending = std::async(std::launch::async, [ctx, b, s1, s2, s3, hs, hc, ha]() mutable {
    return hlm_run(ctx, &b, hs ? &s1 : nullptr, hc ? &s2 : nullptr, ha ? &s3 : nullptr);
  });
  return HLM_OK;
}

int hlm_wait(hlm_ctx* ctx) {
  if (!ctx) return HLM_E_ARG;
  if (!ctx->pending.valid()) return HLM_OK;
  return ctx->pending.get();
}

// ---------------------------------------------------------------- device-resident batches
hlm_devbatch* hlm_batch_upload(hlm_ctx* ctx, const hlm_batch* in) {
  if (!ctx || !in) return nullptr;
  if (check_lengths(ctx, in)) return nullptr;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  hlm_devbatch* b = new hlm_devbatch();
  int64_t n = in->n_pairs;
  bool paired = ctx->p.paired;
  b->n_pairs = n;
  b->paired = paired;
  size_t b1 = (size_t)in->offs1[n], b2 = paired ? (size_t)in->offs2[n] : 0;
  size_t o = 0;
  auto sec = [&](int i, size_t bytes) { b->in_off[i] = o; o += align16(bytes + 32); };
  sec(0, b1); sec(1, b1); sec(2, b2); sec(3, b2);
  sec(4, (n + 1) * 8); sec(5, paired ? (n + 1) * 8 : 0);
  sec(6, in->size_override1 ? n * 4 : 0);
  sec(7, in->rna_split1 ? n * 2 : 0);
  sec(8, in->rna_split2 ? n * 2 : 0);
  bool blocks = ctx->p.rna && in->rna_block_off;
  if (blocks) {
    if (check_blocks(ctx, in)) { delete b; return nullptr; }
    b->blk_off.assign(in->rna_block_off, in->rna_block_off + n + 1);
  }
  sec(9, blocks ? (n + 1) * 8 : 0);
  sec(10, blocks ? (size_t)(b->blk_off[n] - b->blk_off[0]) * sizeof(hlm_rna_block) : 0);
  if (ensure_dev(ctx, b->in, o)) { delete b; return nullptr; }
  uint8_t* d = (uint8_t*)b->in.p;
  auto cp = [&](int i, const void* src, size_t bytes, size_t lead) {
    if (src && bytes) cudaMemcpy(d + b->in_off[i] + lead, src, bytes, cudaMemcpyHostToDevice);
  };
  cp(0, in->bases1, b1, 16); cp(1, in->quals1, b1, 16);
  cp(2, in->bases2, b2, 16); cp(3, in->quals2, b2, 16);
  cp(4, in->offs1, (n + 1) * 8, 0);
  if (paired) cp(5, in->offs2, (n + 1) * 8, 0);
  cp(6, in->size_override1, n * 4, 0);
  cp(7, in->rna_split1, n * 2, 0);
  cp(8, in->rna_split2, n * 2, 0);
  if (blocks) {
    cp(9, in->rna_block_off, (n + 1) * 8, 0);
    cp(10, in->rna_blocks + b->blk_off[0], (size_t)(b->blk_off[n] - b->blk_off[0]) * sizeof(hlm_rna_block), 0);
  }
  b->read_bytes = 2 * (int64_t)(b1 + b2);
  b->has_override = in->size_override1 != nullptr;
  b->has_split1 = in->rna_split1 != nullptr;
  b->has_split2 = in->rna_split2 != nullptr;
  size_t ob = out_layout(n, ctx->n_targets, b->out_off);
  if (ensure_dev(ctx, b->out, ob)) { cudaFree(b->in.p); delete b; return nullptr; }
  if (cudaDeviceSynchronize() != cudaSuccess) { ctx->err = "upload failed"; }
  return b;
}

void hlm_batch_free(hlm_ctx* ctx, hlm_devbatch* b) {
  if (!b) return;
  if (ctx) cudaSetDevice(ctx->device);
  if (b->in.p) cudaFree(b->in.p);
  if (b->out.p) cudaFree(b->out.p);
  delete b;
}

int hlm_run_device(hlm_ctx* ctx, hlm_devbatch* b) {
  if (!ctx || !b) return HLM_E_ARG;
  CU(cudaSetDevice(ctx->device));
  memset(&ctx->prof, 0, sizeof ctx->prof);
  int T = ctx->n_targets;
  int64_t N = b->n_pairs, C = ctx->chunk_pairs;
  uint8_t* d = (uint8_t*)b->in.p;
  CU(cudaEventRecord(ctx->run_start, ctx->slot[0].st));
  for (int i = 1; i < HLM_MAX_SLOTS; i++) CU(cudaStreamWaitEvent(ctx->slot[i].st, ctx->run_start, 0));
  auto submit = [&](Slot& s, int64_t first, int64_t n) -> int {
    int rc = 0;
    if ((rc = ensure_pinned(ctx, s.h_out, 4096))) return rc;
    bool blocks = !b->blk_off.empty();
    int64_t nb = blocks ? (int64_t)(b->blk_off[first + n] - b->blk_off[first]) : 0;
    if ((rc = ensure_work(ctx, s, n, nb))) return rc;
    HlmChunk& ck = s.ck;
    ck.n_pairs = n;
    ck.blk_off = blocks ? (const uint64_t*)(d + b->in_off[9]) + first : nullptr;
    ck.blk = blocks ? (const hlm_rna_block*)(d + b->in_off[10]) - b->blk_off[0] : nullptr;
    ck.blk_first = blocks ? (int64_t)b->blk_off[first] : 0;
    ck.n_blocks = nb;
    ck.bases1 = d + b->in_off[0] + 16;
    ck.quals1 = d + b->in_off[1] + 16;
    ck.bases2 = d + b->in_off[2] + 16;
    ck.quals2 = d + b->in_off[3] + 16;
    ck.offs1 = (const uint64_t*)(d + b->in_off[4]) + first;
    ck.offs2 = (const uint64_t*)(d + b->in_off[5]) + first;
    ck.size_override1 = b->has_override ? (const int32_t*)(d + b->in_off[6]) + first : nullptr;
    ck.rna_split1 = b->has_split1 ? (const uint16_t*)(d + b->in_off[7]) + first : nullptr;
    ck.rna_split2 = b->has_split2 ? (const uint16_t*)(d + b->in_off[8]) + first : nullptr;
    bind_outputs(ck, (uint8_t*)b->out.p, b->out_off, first, T);
    s.first = first;
    s.n = n;
    cudaStream_t st = s.st;
    CU(cudaEventRecord(s.ev[EV_START], st));
    CU(cudaEventRecord(s.ev[EV_H2D], st));
    if (ctx->stagger_ev >= 0 && ctx->last_submitted && ctx->last_submitted != &s)
      CU(cudaStreamWaitEvent(st, ctx->last_submitted->ev[ctx->stagger_ev], 0));
    ctx->last_submitted = &s;
    launch_stages(ctx, s, HLM_STAGE_SCREEN | HLM_STAGE_SCORE | HLM_STAGE_ASSIGN);
    CU(cudaMemcpyAsync(s.h_out.p, ck.counters, 8 * HLM_N_COUNTERS, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(s.ev[EV_D2H], st));
    s.busy = true;
    return 0;
  };
  std::function<int(Slot&)> complete = [&](Slot& s) -> int {
    if (!s.busy) return 0;
    int rc = finish_slot(ctx, s, (unsigned long long*)s.h_out.p);
    if (rc == HLM_E_OVERFLOW && s.n > 1) {
      int64_t first = s.first, n = s.n, h = n / 2;
      ctx->prof.n_overflow_splits++;
      if ((rc = submit(s, first, h))) return rc;
      if ((rc = complete(s))) return rc;
      if ((rc = submit(s, first + h, n - h))) return rc;
      return complete(s);
    }
    return rc == HLM_E_OVERFLOW ? HLM_E_NOMEM : rc;
  };
  int rc = 0;
  int64_t k = 0;
  for (int64_t first = 0; first < N && rc == 0; first += C, k++) {
    Slot& s = ctx->slot[k % ctx->n_slots];
    if ((rc = complete(s))) break;
    rc = submit(s, first, std::min(C, N - first));
  }
  for (int i = 0; i < ctx->n_slots; i++) {
    int r2 = complete(ctx->slot[(k + i) % ctx->n_slots]);
    if (rc == 0) rc = r2;
  }
  ctx->prof.n_pairs = N;
  ctx->prof.bases_bytes = b->read_bytes;
  if (ctx->tracing) {
    if (FILE* f = fopen(getenv("HLM_TRACE"), "w")) {
      for (size_t i = 0; i < ctx->trace.size(); i++)
        fprintf(f, "%.3f%c", ctx->trace[i], (i + 1) % (EV_COUNT + 1) ? ' ' : '\n');
      fclose(f);
    }
    ctx->trace.clear();
  }
  cudaError_t e = cudaGetLastError();
  if (rc == 0 && e != cudaSuccess) {
    ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e);
    rc = HLM_E_CUDA;
  }
  return rc;
}

int hlm_fetch(hlm_ctx* ctx, hlm_devbatch* b, hlm_screen_out* scr, hlm_score_out* sc,
              hlm_assign_out* asg) {
  if (!ctx || !b) return HLM_E_ARG;
  CU(cudaSetDevice(ctx->device));
  int T = ctx->n_targets;
  int64_t n = b->n_pairs;
  const uint8_t* d = (const uint8_t*)b->out.p;
  const size_t* o = b->out_off;
  auto get = [&](void* dst, int sec, size_t bytes) -> int {
    if (!dst) return 0;
    return cudaMemcpy(dst, d + o[sec], bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : HLM_E_CUDA;
  };
  int rc = 0;
  if (scr) {
    rc |= get(scr->flags, O_FLAGS, n); rc |= get(scr->trim_lo1, O_LO1, 2 * n);
    rc |= get(scr->trim_len1, O_LEN1, 2 * n); rc |= get(scr->trim_lo2, O_LO2, 2 * n);
    rc |= get(scr->trim_len2, O_LEN2, 2 * n); rc |= get(scr->locus_mask, O_MASK, 4 * n);
  }
  if (sc) {
    rc |= get(sc->s1, O_S1, 2 * n * T); rc |= get(sc->s2, O_S2, 2 * n * T);
    rc |= get(sc->aln1, O_ALN1, 2 * n * T); rc |= get(sc->aln2, O_ALN2, 2 * n * T);
    rc |= get(sc->nm1, O_NM1, n * T); rc |= get(sc->nm2, O_NM2, n * T);
    rc |= get(sc->lead1, O_LEAD1, n * T); rc |= get(sc->lead2, O_LEAD2, n * T);
    rc |= get(sc->trail1, O_TRAIL1, n * T); rc |= get(sc->trail2, O_TRAIL2, n * T);
  }
  if (asg) {
    rc |= get(asg->target_mask, O_TARGET, 4 * n); rc |= get(asg->visible_mask, O_VISIBLE, 4 * n);
    rc |= get(asg->min_sum, O_MINSUM, 2 * n); rc |= get(asg->others_s1, O_OS1, 2 * n);
    rc |= get(asg->others_s2, O_OS2, 2 * n);
  }
  if (rc) ctx->err = "cudaMemcpy (fetch) failed";
  return rc ? HLM_E_CUDA : 0;
}

// ---------------------------------------------------------------- all GPUs of the box
// Pairs are independent (src/map_dna.cpp:917-1029 only touches keys of its own read id): the batch
// is cut into one contiguous range per GPU, each range runs through its own context on its own
// host thread, and because outputs are caller-owned arrays the gather is pointer arithmetic.
// No collective, no peer traffic; the database is replicated.
struct hlm_multi {
  std::vector<hlm_ctx*> ctx;
  int n_targets = 0;
  std::string err;
};

}  // extern "C"
hlm_ctx* const* hlm_multi_contexts(const hlm_multi* m, int* n) {
  *n = (int)m->ctx.size();
  return m->ctx.data();
}
void hlm_multi_set_error(hlm_multi* m, const std::string& msg) { m->err = msg; }
extern "C" {

hlm_multi* hlm_multi_create(const hlm_db* db, const int* devices, int n_devices, const hlm_params* params,
                            char* err, size_t errlen) {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  std::vector<int> devs;
  if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
  else
    for (int d = 0; d < ndev; d++) devs.push_back(d);
  if (devs.empty()) {
    if (err && errlen) snprintf(err, errlen, "libhlamap needs a CUDA device: there is no CPU path (no device found)");
    return nullptr;
  }
  hlm_multi* m = new hlm_multi();
  for (int d : devs) {  // sequential: the per-device copy of the database is made here
    hlm_ctx* c = hlm_ctx_create(db, d, params, err, errlen);
    if (!c) {
      for (hlm_ctx* x : m->ctx) hlm_ctx_destroy(x);
      delete m;
      return nullptr;
    }
    m->ctx.push_back(c);
  }
  m->n_targets = db->n_loci + 1;
  return m;
}

void hlm_multi_destroy(hlm_multi* m) {
  if (!m) return;
  for (hlm_ctx* c : m->ctx) hlm_ctx_destroy(c);
  delete m;
}

int hlm_multi_n_devices(const hlm_multi* m) { return m ? (int)m->ctx.size() : 0; }
const char* hlm_multi_last_error(const hlm_multi* m) { return m ? m->err.c_str() : "null"; }

int hlm_multi_run(hlm_multi* m, const hlm_batch* in, hlm_screen_out* scr, hlm_score_out* sc,
                  hlm_assign_out* asg) {
  if (!m || !in) return HLM_E_ARG;
  int G = (int)m->ctx.size(), T = m->n_targets;
  int64_t N = in->n_pairs, per = (N + G - 1) / G;
  std::vector<int> rcs(G, 0);
  std::vector<std::thread> th;
  for (int g = 0; g < G; g++) {
    int64_t lo = std::min(N, (int64_t)g * per), hi = std::min(N, lo + per);
    if (lo >= hi) continue;
    th.emplace_back([=, &rcs]() {
      hlm_batch b = *in;  // offsets stay absolute: only the offset arrays and per-pair arrays move
      b.n_pairs = hi - lo;
      b.offs1 = in->offs1 + lo;
      if (in->offs2) b.offs2 = in->offs2 + lo;
      if (in->size_override1) b.size_override1 = in->size_override1 + lo;
      if (in->rna_split1) b.rna_split1 = in->rna_split1 + lo;
      if (in->rna_split2) b.rna_split2 = in->rna_split2 + lo;
      if (in->rna_block_off) b.rna_block_off = in->rna_block_off + lo;  // absolute into rna_blocks
      hlm_screen_out s1{};
      hlm_score_out s2{};
      hlm_assign_out s3{};
      auto at = [&](auto* p, int64_t per_pair) { return p ? p + lo * per_pair : p; };
      if (scr) s1 = {at(scr->flags, 1), at(scr->trim_lo1, 1), at(scr->trim_len1, 1), at(scr->trim_lo2, 1),
                     at(scr->trim_len2, 1), at(scr->locus_mask, 1)};
      if (sc) s2 = {at(sc->s1, T), at(sc->s2, T), at(sc->aln1, T), at(sc->aln2, T), at(sc->nm1, T),
                    at(sc->nm2, T), at(sc->lead1, T), at(sc->lead2, T), at(sc->trail1, T), at(sc->trail2, T)};
      if (asg) s3 = {at(asg->target_mask, 1), at(asg->visible_mask, 1), at(asg->min_sum, 1),
                     at(asg->others_s1, 1), at(asg->others_s2, 1)};
      rcs[g] = hlm_run(m->ctx[g], &b, scr ? &s1 : nullptr, sc ? &s2 : nullptr, asg ? &s3 : nullptr);
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < G; g++)
    if (rcs[g]) {
      m->err = hlm_last_error(m->ctx[g]);
      return rcs[g];
    }
  return HLM_OK;
}

// A/B of the two DP work layouts on synthetic candidates (see k_dpab_* in hlm_kernels.cu)
int hlm_dp_layout_ab(int cuda_device, int n_cands, int read_len, double* ms_thread, double* ms_warp,
                     int64_t* n_mismatch, int64_t* n_cells) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cuda_device < 0 || cuda_device >= ndev) return HLM_E_NODEVICE;
  if (n_cands <= 0 || read_len < 20 || read_len > HLM_MAX_READ) return HLM_E_ARG;
  cudaSetDevice(cuda_device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cuda_device);
  int m = n_cands, n = read_len;
  std::vector<uint32_t> reads((size_t)m * 18), tiles((size_t)m * HLM_TILE_WORDS);
  std::vector<int> offn((size_t)m * 2);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 32); };
  std::vector<uint8_t> t(256), r(HLM_MAX_READ + 8);
  for (int c = 0; c < m; c++) {
    for (int x = 0; x < 256; x++) t[x] = (uint8_t)(rnd() & 3);
    if (rnd() % 8 == 0) t[rnd() % 256] = 4;
    int off = HLM_OFF_MIN + (int)(rnd() % HLM_TILE_STRIDE), len = n;
    int shift = (int)(rnd() % 7) - 3, k = 0;
    for (int x = 0; x < len; x++) r[k++] = t[std::min(255, std::max(0, off + shift + x))];
    int nerr = rnd() % 16;
    for (int e = 0; e < nerr; e++) r[rnd() % len] = (uint8_t)(rnd() % 5);
    if (rnd() % 3 == 0) {  // a small indel
      int p = 5 + (int)(rnd() % (len - 10)), g = 1 + (int)(rnd() % 3);
      if (rnd() & 1) memmove(&r[p], &r[p + g], len - p - g);
      else { memmove(&r[p + g], &r[p], len - p - g); for (int x = 0; x < g; x++) r[p + x] = (uint8_t)(rnd() & 3); }
    }
    if (rnd() % 4 == 0) for (int x = 0, j = 1 + (int)(rnd() % 12); x < j; x++) r[x] = (uint8_t)(rnd() & 3);
    if (rnd() % 4 == 0) for (int x = 0, j = 1 + (int)(rnd() % 12); x < j; x++) r[len - 1 - x] = (uint8_t)(rnd() & 3);
    uint32_t* R = &reads[(size_t)c * 18];
    for (int x = 0; x < 18; x++) R[x] = x >= 12 ? 0xFFFFFFFFu : 0u;
    for (int x = 0; x < len; x++) {
      if (r[x] > 3) continue;
      R[12 + (x >> 5)] &= ~(1u << (x & 31));
      if (r[x] & 1) R[x >> 5] |= 1u << (x & 31);
      if (r[x] & 2) R[6 + (x >> 5)] |= 1u << (x & 31);
    }
    uint32_t* Tt = &tiles[(size_t)c * HLM_TILE_WORDS];
    for (int x = 0; x < HLM_TILE_WORDS; x++) Tt[x] = 0;
    for (int x = 0; x < 256; x++) {
      if (t[x] > 3) { Tt[16 + (x >> 5)] |= 1u << (x & 31); continue; }
      if (t[x] & 1) Tt[x >> 5] |= 1u << (x & 31);
      if (t[x] & 2) Tt[8 + (x >> 5)] |= 1u << (x & 31);
    }
    offn[2 * c] = off;
    offn[2 * c + 1] = len;
  }
  uint32_t *d_r = nullptr, *d_t = nullptr;
  int* d_o = nullptr;
  int4 *d_a = nullptr, *d_b = nullptr;
  if (cudaMalloc(&d_r, reads.size() * 4) || cudaMalloc(&d_t, tiles.size() * 4) || cudaMalloc(&d_o, offn.size() * 4) ||
      cudaMalloc(&d_a, (size_t)m * 16) || cudaMalloc(&d_b, (size_t)m * 16))
    return HLM_E_NOMEM;
  cudaMemcpy(d_r, reads.data(), reads.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_t, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_o, offn.data(), offn.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double ms[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    int4* out = which ? d_b : d_a;
    hlm_launch_dpab(which, d_r, d_t, d_o, m, out, prop.multiProcessorCount, 0);
    cudaEventRecord(e0, 0);
    for (int rep = 0; rep < 3; rep++) hlm_launch_dpab(which, d_r, d_t, d_o, m, out, prop.multiProcessorCount, 0);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float f = 0;
    cudaEventElapsedTime(&f, e0, e1);
    ms[which] = f / 3.0;
  }
  std::vector<int> a((size_t)m * 4), b((size_t)m * 4);
  cudaMemcpy(a.data(), d_a, (size_t)m * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), d_b, (size_t)m * 16, cudaMemcpyDeviceToHost);
  int64_t bad = 0;
  for (size_t i = 0; i < a.size(); i += 4)
    if (memcmp(&a[i], &b[i], 16) != 0) bad++;
  cudaError_t ce = cudaGetLastError();
  cudaFree(d_r); cudaFree(d_t); cudaFree(d_o); cudaFree(d_a); cudaFree(d_b);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ms_thread) *ms_thread = ms[0];
  if (ms_warp) *ms_warp = ms[1];
  if (n_mismatch) *n_mismatch = bad;
  if (n_cells) *n_cells = (int64_t)m * n * HLM_NB;
  return ce == cudaSuccess ? HLM_OK : HLM_E_CUDA;
}

int hlm_get_profile(const hlm_ctx* ctx, hlm_profile* out) {
  if (!ctx || !out) return HLM_E_ARG;
  *out = ctx->prof;
  return 0;
}

int hlm_int32_peak(int cuda_device, double* giga_iadd, double* giga_imad, double* giga_mix,
                   double* giga_viaddmnmx) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cuda_device < 0 || cuda_device >= ndev)
    return HLM_E_NODEVICE;
  cudaSetDevice(cuda_device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cuda_device);
  int sms = prop.multiProcessorCount;
  double* outs[4] = {giga_iadd, giga_imad, giga_mix, giga_viaddmnmx};
  for (int w = 0; w < 4; w++)
    if (outs[w]) *outs[w] = hlm_int_peak_run(w, sms, 0);
  return 0;
}

// the same, with the logic / shift chains k_myers is made of: out[0..6] = IADD3, IMAD, IADD3+IMAD mix,
// VIADDMNMX, LOP3, SHF, 3 LOP3 : 1 SHF mix (giga warp-lane ops per second); n = entries wanted
int hlm_int32_peaks(int cuda_device, double* out, int n) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cuda_device < 0 || cuda_device >= ndev)
    return HLM_E_NODEVICE;
  if (!out || n < 0) return HLM_E_ARG;
  cudaSetDevice(cuda_device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cuda_device);
  for (int w = 0; w < n && w < 7; w++) out[w] = hlm_int_peak_run(w, prop.multiProcessorCount, 0);
  return 0;
}

}  // extern "C"
