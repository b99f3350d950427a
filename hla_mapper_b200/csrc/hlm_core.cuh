// hlm_core.cuh — the arithmetic of the scoring kernels as inline functions.
//
// Compiled for sm_100a inside hlm_kernels.cu; the same source also compiles as plain C++ (see
// tools/hostsim.cpp) so the CPU test-suite can check the bit-vector prefilter and the register
// DP against the oracle without a GPU.  That host build is test-only: the product never
// computes on the CPU.
//
// Geometry of one candidate (tile, off): read base i sits on tile column off + i on the main
// diagonal; DP row i (1..n) owns band cells k = 0..40 <-> tile column off - BAND + (i-1) + k for
// the base compared on the diagonal move into the cell.  off is in [BAND, BAND+32).
#ifndef HLM_CORE_CUH
#define HLM_CORE_CUH

#include "hlm_common.h"

#ifdef __CUDACC__
#define HLM_HD __host__ __device__ __forceinline__
#else
#define HLM_HD inline
#endif

HLM_HD uint32_t hlm_funnel(uint32_t lo, uint32_t hi, int s) {  // bits [s, s+32) of hi:lo, s in 0..31
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
HLM_HD int hlm_max(int a, int b) { return a > b ? a : b; }
HLM_HD int hlm_addmax(int a, int b, int c) {  // max(a + b, c)
#if defined(__CUDA_ARCH__)
  return __viaddmax_s32(a, b, c);
#else
  return hlm_max(a + b, c);
#endif
}
HLM_HD int hlm_max3(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __vimax3_s32(a, b, c);
#else
  return hlm_max(hlm_max(a, b), c);
#endif
}

// 64 consecutive plane bits starting at column pos (only words 0..7 exist; beyond -> word 7,
// which only feeds band cells that are masked off)
struct HlmWin {
  uint32_t a, b, c;  // three consecutive words of one tile plane
};
HLM_HD uint32_t hlm_tw(const uint32_t* plane, int w) { return plane[w < 8 ? w : 7]; }

// ------------------------------------------------------------------ Myers / Hyyro lower bound
// Bit-parallel unit-cost distance over the same 41-diagonal band the DP uses.  Row-wise
// formulation: the 64-bit vectors hold HORIZONTAL deltas of one DP row, bit k = band cell k; going
// to the next row shifts the band one reference column to the right, so the previous row's
// vectors are shifted right by one and the entering top cell assumes delta +1, the cell left of
// the band assumes vertical delta +1.  Those assumptions keep every computed value between the
// unbanded and the banded (+inf outside) distance, so the result never exceeds the cost field of
// any alignment the DP can produce: it is a valid lower bound (tested against the
// oracle in tests/test_core_hostsim.py).
HLM_HD int hlm_myers_lb(const uint32_t* __restrict__ rd, int n, const uint32_t* __restrict__ tile,
                        int off) {
  const uint64_t MASK = (1ull << HLM_NB) - 1, TOP = 1ull << (HLM_NB - 1);
  const uint32_t* T0 = tile;
  const uint32_t* T1 = tile + 8;
  const uint32_t* TN = tile + 16;
  uint64_t HP = 0, HN = 0;
  int score = 0;  // D(i, k=0)
  int pos = off - HLM_BAND;
  int w = pos >> 5, s = pos & 31;
  HlmWin w0 = {hlm_tw(T0, w), hlm_tw(T0, w + 1), hlm_tw(T0, w + 2)};
  HlmWin w1 = {hlm_tw(T1, w), hlm_tw(T1, w + 1), hlm_tw(T1, w + 2)};
  HlmWin wn = {hlm_tw(TN, w), hlm_tw(TN, w + 1), hlm_tw(TN, w + 2)};
  uint32_t r0 = 0, r1 = 0, rn = 0;
  for (int i = 0; i < n; i++) {
    if ((i & 31) == 0) {
      r0 = rd[i >> 5];
      r1 = rd[6 + (i >> 5)];
      rn = rd[12 + (i >> 5)];
    }
    uint64_t t0 = (uint64_t)hlm_funnel(w0.a, w0.b, s) | ((uint64_t)hlm_funnel(w0.b, w0.c, s) << 32);
    uint64_t t1 = (uint64_t)hlm_funnel(w1.a, w1.b, s) | ((uint64_t)hlm_funnel(w1.b, w1.c, s) << 32);
    uint64_t tn = (uint64_t)hlm_funnel(wn.a, wn.b, s) | ((uint64_t)hlm_funnel(wn.b, wn.c, s) << 32);
    uint64_t m0 = (uint64_t)0 - (uint64_t)(r0 & 1), m1 = (uint64_t)0 - (uint64_t)(r1 & 1),
             mn = (uint64_t)0 - (uint64_t)(rn & 1);
    uint64_t Eq = ~(t0 ^ m0) & ~(t1 ^ m1) & ~(tn | mn) & MASK;
    r0 >>= 1;
    r1 >>= 1;
    rn >>= 1;
    uint64_t HPa = (HP >> 1) | TOP, HNa = HN >> 1;
    uint64_t D0 = (((Eq & HPa) + HPa) ^ HPa) | Eq | HNa;
    uint64_t VP = HNa | ~(D0 | HPa), VN = D0 & HPa;
    uint64_t X = (VP << 1) | 1ull, Y = VN << 1;
    HP = (Y | ~(D0 | X)) & MASK;
    HN = (X & D0) & MASK;
    score += 1 - (int)(D0 & 1ull);
    if (++s == 32) {
      s = 0;
      w++;
      w0.a = w0.b; w0.b = w0.c; w0.c = hlm_tw(T0, w + 2);
      w1.a = w1.b; w1.b = w1.c; w1.c = hlm_tw(T1, w + 2);
      wn.a = wn.b; wn.b = wn.c; wn.c = hlm_tw(TN, w + 2);
    }
  }
  int best = score, cur = score;
  for (int k = 1; k < HLM_NB; k++) {
    cur += (int)((HP >> k) & 1ull) - (int)((HN >> k) & 1ull);
    best = cur < best ? cur : best;
  }
  return best;
}

// ------------------------------------------------------------------ banded affine-gap DP
// One candidate per thread; the 41 H and 41 E band cells live in registers (diagonal-indexed,
// so no shifting between rows).  Same recurrence, constants and tie rules as the aligner
// definition of DESIGN.md §3 (the test-suite checks bit-identity against the CPU oracle).
struct HlmAln {
  int S, cost, lead, trail;
};

HLM_HD HlmAln hlm_dp_align(const uint32_t* __restrict__ rd, int n, const uint32_t* __restrict__ tile,
                           int off) {
  const uint32_t* T0 = tile;
  const uint32_t* T1 = tile + 8;
  const uint32_t* TN = tile + 16;
  int H[HLM_NB], E[HLM_NB];
#pragma unroll
  for (int k = 0; k < HLM_NB; k++) {
    H[k] = 0;
    E[k] = HLM_V_NEG;
  }
  int best = HLM_V_NEG, best_i = 0;
  int pos = off - HLM_BAND;
  int w = pos >> 5, s = pos & 31;
  HlmWin w0 = {hlm_tw(T0, w), hlm_tw(T0, w + 1), hlm_tw(T0, w + 2)};
  HlmWin w1 = {hlm_tw(T1, w), hlm_tw(T1, w + 1), hlm_tw(T1, w + 2)};
  HlmWin wn = {hlm_tw(TN, w), hlm_tw(TN, w + 1), hlm_tw(TN, w + 2)};
  uint32_t r0 = 0, r1 = 0, rn = 0;
  int start = 0;                                  // -(CLIP + i*CU + i), built incrementally
  int trailpen = HLM_V_CLIP + n * HLM_CU;         // CLIP + (n-i)*CU
  for (int i = 1; i <= n; i++) {
    if (((i - 1) & 31) == 0) {
      r0 = rd[(i - 1) >> 5];
      r1 = rd[6 + ((i - 1) >> 5)];
      rn = rd[12 + ((i - 1) >> 5)];
    }
    uint32_t m0 = 0u - (r0 & 1u), m1 = 0u - (r1 & 1u), mn = 0u - (rn & 1u);
    r0 >>= 1;
    r1 >>= 1;
    rn >>= 1;
    uint32_t eq_lo = ~(hlm_funnel(w0.a, w0.b, s) ^ m0) & ~(hlm_funnel(w1.a, w1.b, s) ^ m1) &
                     ~(hlm_funnel(wn.a, wn.b, s) | mn);
    uint32_t eq_hi = ~(hlm_funnel(w0.b, w0.c, s) ^ m0) & ~(hlm_funnel(w1.b, w1.c, s) ^ m1) &
                     ~(hlm_funnel(wn.b, wn.c, s) | mn);
    start = (i == 1) ? -(HLM_V_CLIP + HLM_CU + 1) : start - (HLM_CU + 1);
    trailpen -= HLM_CU;
    int f = HLM_V_NEG, hl = HLM_V_NEG, rowmax = HLM_V_NEG;
#pragma unroll
    for (int k = 0; k < HLM_NB; k++) {
      uint32_t bit = (k < 32) ? ((eq_lo >> k) & 1u) : ((eq_hi >> (k - 32)) & 1u);
      int sub = bit ? HLM_V_MATCH : HLM_V_MISM;
      int e = HLM_V_NEG;
      if (k + 1 < HLM_NB) e = hlm_addmax(E[k + 1], HLM_V_GAPE, H[k + 1] + HLM_V_GAPO);
      if (k >= 1) f = hlm_addmax(f, HLM_V_GAPE, hl + HLM_V_GAPO);
      int h = hlm_addmax(H[k], sub, start);
      h = hlm_max3(h, e, f);
      H[k] = h;
      E[k] = e;
      hl = h;
      rowmax = hlm_max(rowmax, h);
    }
    int fin = rowmax - (i < n ? trailpen : 0);
    if (fin >= best) {
      best = fin;
      best_i = i;
    }
    if (++s == 32) {
      s = 0;
      w++;
      w0.a = w0.b; w0.b = w0.c; w0.c = hlm_tw(T0, w + 2);
      w1.a = w1.b; w1.b = w1.c; w1.c = hlm_tw(T1, w + 2);
      wn.a = wn.b; wn.b = wn.c; wn.c = hlm_tw(TN, w + 2);
    }
  }
  HlmAln r;
  r.S = (best + HLM_SC - 1) >> 20;
  int rem = r.S * HLM_SC - best;
  r.cost = rem >> 9;
  r.lead = rem & 511;
  r.trail = n - best_i;
  return r;
}

// does read seed q (12 bases at read offset 12q) match the tile exactly at column off + 12q ?
HLM_HD bool hlm_seed_match(const uint32_t* __restrict__ rd, const uint32_t* __restrict__ tile,
                           int off, int q) {
  int rp = q * HLM_SEED, tc = off + rp;
  int rw = rp >> 5, rs = rp & 31, tw = tc >> 5, ts = tc & 31;
  uint32_t a0 = hlm_funnel(rd[rw], rw + 1 < 6 ? rd[rw + 1] : 0u, rs);
  uint32_t a1 = hlm_funnel(rd[6 + rw], rw + 1 < 6 ? rd[6 + rw + 1] : 0u, rs);
  uint32_t an = hlm_funnel(rd[12 + rw], rw + 1 < 6 ? rd[12 + rw + 1] : 0xFFFFFFFFu, rs);
  uint32_t b0 = hlm_funnel(tile[tw], hlm_tw(tile, tw + 1), ts);
  uint32_t b1 = hlm_funnel(tile[8 + tw], hlm_tw(tile + 8, tw + 1), ts);
  uint32_t bn = hlm_funnel(tile[16 + tw], hlm_tw(tile + 16, tw + 1), ts);
  return (((a0 ^ b0) | (a1 ^ b1) | an | bn) & 0xFFFu) == 0;
}

// 24-bit seed code (base x at bits 2x) of read seed q; returns false if the seed holds an N
HLM_HD bool hlm_seed_code(const uint32_t* __restrict__ rd, int q, uint32_t* code) {
  int rp = q * HLM_SEED, rw = rp >> 5, rs = rp & 31;
  uint32_t a0 = hlm_funnel(rd[rw], rw + 1 < 6 ? rd[rw + 1] : 0u, rs) & 0xFFFu;
  uint32_t a1 = hlm_funnel(rd[6 + rw], rw + 1 < 6 ? rd[6 + rw + 1] : 0u, rs) & 0xFFFu;
  uint32_t an = hlm_funnel(rd[12 + rw], rw + 1 < 6 ? rd[12 + rw + 1] : 0xFFFFFFFFu, rs) & 0xFFFu;
  // interleave: bit x of a0 -> bit 2x, bit x of a1 -> bit 2x+1
  uint32_t c = 0;
#pragma unroll
  for (int x = 0; x < HLM_SEED; x++) c |= (((a0 >> x) & 1u) | (((a1 >> x) & 1u) << 1)) << (2 * x);
  *code = c;
  return an == 0;
}

#endif
