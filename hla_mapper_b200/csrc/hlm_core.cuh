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
// a + b issued as a multiply-add (a * one + b, one == 1 at run time but opaque to the compiler):
// IMAD runs on the FMA pipe, which the DP otherwise leaves idle while the ALU pipe (all min/max
// and DPX instructions) is the bottleneck
HLM_HD int hlm_add_fma(int a, int b, int one) {
#if defined(__CUDA_ARCH__)
  int d;
  asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
  return d;
#else
  (void)one;
  return a + b;
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
// Bit-parallel unit-cost distance over the same 41-diagonal band the DP uses (Hyyro's
// diagonal-band formulation, transposed to rows).  Band cell k of read row i is reference
// column pos + i + k, c_k its distance value.  Between rows the state is the pair of 64-bit
// vectors (SP, SN) = the horizontal deltas c_{k+1} - c_k = +1 / -1 of the current row: going to
// the next row the band moves one column to the right, so exactly these deltas separate the
// diagonal predecessor (same k) from the vertical predecessor (k + 1) of every new cell and
// only the diagonal-delta vector D0 has to be shifted:
//     D0  = (((Eq & SP) + SP) ^ SP) | Eq | SN          c'_k == c_k
//     VP  = SN | ~(D0 | SP),   VN = D0 & SP            c'_k - c_{k+1}
//     SP' = VN | ~((D0 >> 1) | VP),   SN' = (D0 >> 1) & VP
// The cell right of the band (k = 41) is assumed one above its left neighbour (bit 40 of SP is
// always set) and the cell left of the band is never read.  Those assumptions keep every value
// between the unbanded and the banded (+inf outside) distance, so the result never exceeds the
// cost field of any alignment the DP can produce: a valid lower bound (DESIGN.md section 5;
// tests/test_core_hostsim.py checks it against the oracle).
//
// Eq comes from per-candidate match vectors PEQ[b] (bit c = tile column c holds base b; row 4 is
// all zero for N read bases), kept in shared memory on the device; the read is consumed as
// 4-bit codes, eight per word (k_pack_units spreads them from the bit planes with three
// shift-or-mask steps and stores them next to the planes).
#define HLM_PEQ_ROWS 41 /* 5 vectors x 8 words + 1 pad word (read when the window ends at word 7) */

HLM_HD uint32_t hlm_spread8(uint32_t x) {  // bit j of x (j < 8) -> bit 4j
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// fills PEQ rows from the tile planes; peq(r) is an lvalue for row r
template <class PEQ>
HLM_HD void hlm_peq_build(const uint32_t* __restrict__ tile, PEQ peq) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int w = 0; w < 8; w++) {
    uint32_t p0 = tile[w], p1 = tile[8 + w], nn = ~tile[16 + w];
    peq(w) = ~p0 & ~p1 & nn;
    peq(8 + w) = p0 & ~p1 & nn;
    peq(16 + w) = ~p0 & p1 & nn;
    peq(24 + w) = p0 & p1 & nn;
    peq(32 + w) = 0u;
  }
  peq(40) = 0u;
}

// codes: the read as 4-bit codes, 8 per word (code bit 0 / 1 = the base, bit 2 = N -> PEQ row 4)
// rstar / plb: see hlm_myers_core_dev (the band minimum after rstar rows, any 0 <= rstar <= n)
template <class PEQ>
HLM_HD int hlm_myers_core(const uint32_t* __restrict__ codes, int n, int off, PEQ peq, int rstar = -1,
                          int* plb = nullptr) {
  const uint64_t MASK = (1ull << HLM_NB) - 1, TOP = 1ull << (HLM_NB - 1);
  uint64_t SP = TOP, SN = 0;
  int score = 0;  // c_0
  int p = off - HLM_BAND;
  uint32_t cw = 0;
#define HLM_MYERS_REFILL(i) cw = codes[(i) >> 3];
#define HLM_MYERS_ROW()                                                                            \
  {                                                                                               \
    int row = (int)(cw & 7u) * 8 + (p >> 5); /* codes 5..7 cannot occur (N clears the code bits) */ \
    cw >>= 4;                                                                                     \
    uint32_t a = peq(row), b = peq(row + 1), c = peq(row + 2);                                    \
    /* bits above the band (>= 41) of Eq, SP and VP are garbage: the addition only carries */    \
    /* upwards and D0 is masked, so they never reach a band bit */                                \
    uint64_t Eq = (uint64_t)hlm_funnel(a, b, p & 31) | ((uint64_t)hlm_funnel(b, c, p & 31) << 32); \
    p++;                                                                                          \
    uint64_t D0 = ((((Eq & SP) + SP) ^ SP) | Eq | SN) & MASK;                                     \
    uint64_t VP = SN | ~(D0 | SP), VN = (D0 & SP) | TOP; /* TOP: the cell right of the band */    \
    uint64_t D0s = D0 >> 1;                                                                       \
    SP = VN | ~(D0s | VP);                                                                        \
    SN = D0s & VP;                                                                                \
    score += 1 - (int)(D0 & 1ull);                                                                \
  }
#define HLM_MYERS_BANDMIN(out)                                        \
  {                                                                   \
    int best_ = score, cur_ = score;                                  \
    for (int k = 0; k < HLM_NB - 1; k++) {                            \
      cur_ += (int)((SP >> k) & 1ull) - (int)((SN >> k) & 1ull);     \
      best_ = cur_ < best_ ? cur_ : best_;                            \
    }                                                                 \
    out = best_;                                                      \
  }
  int i = 0;
  if (plb && rstar == 0) *plb = 0;
  for (; i < n; i++) {
    if ((i & 7) == 0) HLM_MYERS_REFILL(i)
    HLM_MYERS_ROW()
    if (plb && i + 1 == rstar) HLM_MYERS_BANDMIN(*plb)
  }
#undef HLM_MYERS_REFILL
#undef HLM_MYERS_ROW
  int best;
  HLM_MYERS_BANDMIN(best)
#undef HLM_MYERS_BANDMIN
  return best;
}

#ifdef __CUDACC__
// ---- the same recurrence as hlm_myers_core, arranged for the sm_100a pipes (k_myers is bound by
// the INT32 ALU pipe: profiles/r1g_k_myers_regions.txt).  What can run as a plain 32-bit
// multiply-add goes to the FMA pipe, which is otherwise idle: the column counter, the byte
// address of the PEQ row, the times-eight of the read codes.  (High-half multiplies - a way to
// express right shifts on that pipe - were measured and are too slow to pay.)  The "| TOP" of the
// generic form is dropped: with D0 cut to the band, bit 40 of SP regenerates itself
// (VP_40 = 0 -> SP'_40 = ~D0_41 = 1) and nothing above the band is ever set.  The score is not
// updated per row: bit 0 of D0 is shifted into an accumulator and counted once per eight rows.
// `one` == 1 at run time, opaque to the compiler.
__device__ __forceinline__ uint32_t hlm_fma_mad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// minimum over a band row: c0 + prefix sums of the horizontal deltas (SP / SN), four columns per
// look-up in a 256-entry table (hlm_band_min_entry) instead of forty single steps
__device__ __forceinline__ int hlm_band_min(uint64_t SP, uint64_t SN, int c0, const uint16_t* __restrict__ mintab) {
  int best = c0, cur8 = best - 8;
  uint32_t spl = (uint32_t)SP, sph = (uint32_t)(SP >> 32), snl = (uint32_t)SN, snh = (uint32_t)(SN >> 32);
  // bytes of ev / od: (SP nibble << 4 | SN nibble) of the even / odd nibbles of a word
  uint32_t ev = ((spl & 0x0F0F0F0Fu) << 4) | (snl & 0x0F0F0F0Fu);
  uint32_t od = (spl & 0xF0F0F0F0u) | ((snl >> 4) & 0x0F0F0F0Fu);
  uint32_t hi = ((sph & 0x0Fu) << 4) | (snh & 0x0Fu) | ((sph & 0xF0u) << 8) | ((snh & 0xF0u) << 4);
#define HLM_BAND_MIN_STEP(IDX)                                                \
  {                                                                           \
    uint32_t e = mintab[IDX];                                                 \
    best = __viaddmin_s32(cur8, (int)(e >> 8), best);                         \
    cur8 += (int)(e & 0xFFu) - 8;                                             \
  }
  HLM_BAND_MIN_STEP(__byte_perm(ev, 0u, 0x4440))
  HLM_BAND_MIN_STEP(__byte_perm(od, 0u, 0x4440))
  HLM_BAND_MIN_STEP(__byte_perm(ev, 0u, 0x4441))
  HLM_BAND_MIN_STEP(__byte_perm(od, 0u, 0x4441))
  HLM_BAND_MIN_STEP(__byte_perm(ev, 0u, 0x4442))
  HLM_BAND_MIN_STEP(__byte_perm(od, 0u, 0x4442))
  HLM_BAND_MIN_STEP(__byte_perm(ev, 0u, 0x4443))
  HLM_BAND_MIN_STEP(__byte_perm(od, 0u, 0x4443))
  HLM_BAND_MIN_STEP(hi & 0xFFu)   // columns 32..35
  HLM_BAND_MIN_STEP(hi >> 8)      // columns 36..39
#undef HLM_BAND_MIN_STEP
  return best;
}
// peq_addr: shared-memory byte address of this thread's PEQ row 0; rows are ROW_BYTES apart
template <int ROW_BYTES>
// rstar / plb (capped scoring): the band minimum after rstar rows (a multiple of 8, or < 0 for
// "not wanted") is returned through *plb - a lower bound of cost - trailing clip of every alignment
// the DP can report (the row minimum never decreases from one row to the next, and a reported
// alignment clips at most mm_max trailing bases)
__device__ __forceinline__ int hlm_myers_core_dev(const uint32_t* __restrict__ codes, int n, int off,
                                                  uint32_t peq_addr, uint32_t one,
                                                  const uint16_t* __restrict__ mintab, int rstar, int* plb) {
  const uint64_t MASK = (1ull << HLM_NB) - 1;
  uint64_t SP = 1ull << (HLM_NB - 1), SN = 0;
  uint32_t p = (uint32_t)(off - HLM_BAND);
  uint32_t acc = 0;
  int zeros = 0;
  const uint32_t krow = one * ROW_BYTES;
  asm volatile("" ::: "memory");  // the PEQ rows were written with plain stores
  // C8: eight times the read code of the row; S: its first band column (the funnel shift wraps)
#define HLM_MYERS_DEV_ROW(C8, S)                                                                    \
  {                                                                                               \
    uint32_t ad = hlm_fma_mad((C8) + ((S) >> 5), krow, peq_addr);                                  \
    uint32_t a, b, c;                                                                             \
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(ad));                                   \
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(b) : "r"(ad), "n"(ROW_BYTES));                \
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(c) : "r"(ad), "n"(2 * ROW_BYTES));            \
    uint64_t Eq = (uint64_t)__funnelshift_r(a, b, S) | ((uint64_t)__funnelshift_r(b, c, S) << 32); \
    uint64_t D0 = ((((Eq & SP) + SP) ^ SP) | Eq | SN) & MASK;                                     \
    uint64_t VP = SN | ~(D0 | SP), VN = D0 & SP;                                                  \
    uint64_t D0s = D0 >> 1;                                                                       \
    SP = VN | ~(D0s | VP);                                                                        \
    SN = D0s & VP;                                                                                \
    acc = __funnelshift_r(acc, (uint32_t)D0, 1); /* (acc >> 1) | (bit 0 of D0) << 31 */           \
  }
  int i = 0;
  // two passes over ONE copy of the row loop: rows [0, rstar), the prefix bound, rows [rstar, n)
  int stop = rstar >= 0 ? rstar : n;
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
  for (; i + 8 <= stop; i += 8) {
    uint32_t cw = codes[i >> 3];
    // codes of the even / odd rows, one per byte, times eight
    uint32_t ev = hlm_fma_mad(cw & 0x07070707u, 8u * one, 0u), od = hlm_fma_mad((cw >> 4) & 0x07070707u, 8u * one, 0u);
    HLM_MYERS_DEV_ROW(__byte_perm(ev, 0u, 0x4440), p)
    HLM_MYERS_DEV_ROW(__byte_perm(od, 0u, 0x4440), hlm_fma_mad(one, 1u, p))
    HLM_MYERS_DEV_ROW(__byte_perm(ev, 0u, 0x4441), hlm_fma_mad(one, 2u, p))
    HLM_MYERS_DEV_ROW(__byte_perm(od, 0u, 0x4441), hlm_fma_mad(one, 3u, p))
    HLM_MYERS_DEV_ROW(__byte_perm(ev, 0u, 0x4442), hlm_fma_mad(one, 4u, p))
    HLM_MYERS_DEV_ROW(__byte_perm(od, 0u, 0x4442), hlm_fma_mad(one, 5u, p))
    HLM_MYERS_DEV_ROW(__byte_perm(ev, 0u, 0x4443), hlm_fma_mad(one, 6u, p))
    HLM_MYERS_DEV_ROW(__byte_perm(od, 0u, 0x4443), hlm_fma_mad(one, 7u, p))
    zeros += __popc(acc >> 24);
    p = hlm_fma_mad(one, 8u, p);
  }
  if (pass == 0 && rstar >= 0) *plb = hlm_band_min(SP, SN, rstar - zeros, mintab);
  if (stop == n) break;
  stop = n;
  }
  if (i < n) {
    uint32_t cw = codes[i >> 3];
    int rem = n - i;
    for (int r = 0; r < rem; r++) {
      HLM_MYERS_DEV_ROW((cw << 3) & 0x38u, p)
      cw >>= 4;
      p++;
    }
    zeros += __popc(acc >> (32 - rem));
  }
#undef HLM_MYERS_DEV_ROW
  asm volatile("" ::: "memory");  // the next candidate overwrites the PEQ rows
  return hlm_band_min(SP, SN, n - zeros, mintab);  // minimum over the last band row
}
// entry t = (SP nibble << 4 | SN nibble): low byte = sum of the four deltas + 8, high byte = the
// smallest of the four prefix sums + 8
__host__ __device__ inline uint16_t hlm_band_min_entry(int t) {
  int sp = t >> 4, sn = t & 15, cur = 0, mn = 8;
  for (int k = 0; k < 4; k++) {
    cur += ((sp >> k) & 1) - ((sn >> k) & 1);
    mn = cur < mn ? cur : mn;
  }
  return (uint16_t)((cur + 8) | ((mn + 8) << 8));
}
#endif

// host / generic entry point: PEQ in a local array
struct HlmPeqLocal {
  uint32_t* v;
  HLM_HD uint32_t& operator()(int r) const { return v[r]; }
};
HLM_HD int hlm_myers_lb(const uint32_t* __restrict__ rd, int n, const uint32_t* __restrict__ tile,
                        int off, int rstar = -1, int* plb = nullptr) {
  uint32_t buf[HLM_PEQ_ROWS], codes[24];
  HlmPeqLocal peq{buf};
  hlm_peq_build(tile, peq);
  for (int j = 0; j < 24; j++) {
    int w = j >> 2, sh = 8 * (j & 3);
    codes[j] = hlm_spread8((rd[w] >> sh) & 0xFFu) | (hlm_spread8((rd[6 + w] >> sh) & 0xFFu) << 1) |
               (hlm_spread8((rd[12 + w] >> sh) & 0xFFu) << 2);
  }
  return hlm_myers_core(codes, n, off, peq, rstar, plb);
}

// ------------------------------------------------------------------ banded affine-gap DP
// One candidate per thread; the 41 H and 41 E band cells live in registers (diagonal-indexed,
// so no shifting between rows).  Same recurrence, constants and tie rules as the aligner
// definition of DESIGN.md §3 (the test-suite checks bit-identity against the CPU oracle).
struct HlmAln {
  int S, cost, lead, trail;
};

// noclip (HLM_CLIP_NONE): no row may start a new alignment (no leading clip) and only the last
// row may end one (no trailing clip); two row-level selects, nothing in the cell loop.
HLM_HD HlmAln hlm_dp_align(const uint32_t* __restrict__ rd, int n, const uint32_t* __restrict__ tile,
                           int off, int one = 1, int noclip = 0) {
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
    if (noclip) start = HLM_V_NEG;
    trailpen -= HLM_CU;
    // Cell update, arranged so that the ALU pipe (min/max, DPX) sees 5.5 instructions per cell:
    //   e  = max(E[k+1] + GAPE, H[k+1] + GAPO)                  VIADDMNMX (+ add)
    //   f  = max(f + GAPE, h'[k-1] + GAPO)                       VIADDMNMX (+ add)
    //   h' = max(H[k] + sub, e, start)                           VIMNMX3   (+ predicated add)
    //   h  = max(h', f)                                          VIMNMX
    // f may use h' (the cell value without its own f) instead of h: h = max(h', f_prev) and
    // f_prev + GAPO < f_prev + GAPE, so the dropped term never wins.
    int f = HLM_V_NEG, hp = HLM_V_NEG, rowmax = HLM_V_NEG, hprev = HLM_V_NEG;
#pragma unroll
    for (int k = 0; k < HLM_NB; k++) {
      uint32_t bit = (k < 32) ? (eq_lo & (1u << k)) : (eq_hi & (1u << (k - 32)));
      int e = HLM_V_NEG;
      if (k + 1 < HLM_NB) e = hlm_addmax(E[k + 1], HLM_V_GAPE, hlm_add_fma(H[k + 1], HLM_V_GAPO, one));
      if (k >= 1) f = hlm_addmax(f, HLM_V_GAPE, hlm_add_fma(hp, HLM_V_GAPO, one));
      int hm = hlm_add_fma(H[k], HLM_V_MISM, one);
      if (bit) hm = hlm_add_fma(H[k], HLM_V_MATCH, one);
      hp = hlm_max3(hm, e, start);
      int h = hlm_max(hp, f);
      H[k] = h;
      E[k] = e;
      if (k & 1) rowmax = hlm_max3(rowmax, h, hprev);
      hprev = h;
    }
    rowmax = hlm_max(rowmax, hprev);  // NB is odd: the last cell
    int fin = rowmax - (i < n ? trailpen : 0);
    if (noclip && i < n) fin = HLM_V_NEG - 1;
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

// does the read seed at read offset rp (12 bases) match the tile exactly at column off + rp ?
HLM_HD bool hlm_seed_match(const uint32_t* __restrict__ rd, const uint32_t* __restrict__ tile,
                           int off, int rp) {
  int tc = off + rp;
  int rw = rp >> 5, rs = rp & 31, tw = tc >> 5, ts = tc & 31;
  uint32_t a0 = hlm_funnel(rd[rw], rw + 1 < 6 ? rd[rw + 1] : 0u, rs);
  uint32_t a1 = hlm_funnel(rd[6 + rw], rw + 1 < 6 ? rd[6 + rw + 1] : 0u, rs);
  uint32_t an = hlm_funnel(rd[12 + rw], rw + 1 < 6 ? rd[12 + rw + 1] : 0xFFFFFFFFu, rs);
  uint32_t b0 = hlm_funnel(tile[tw], hlm_tw(tile, tw + 1), ts);
  uint32_t b1 = hlm_funnel(tile[8 + tw], hlm_tw(tile + 8, tw + 1), ts);
  uint32_t bn = hlm_funnel(tile[16 + tw], hlm_tw(tile + 16, tw + 1), ts);
  return (((a0 ^ b0) | (a1 ^ b1) | an | bn) & 0xFFFu) == 0;
}

// 24-bit seed code (base x at bits 2x) of the read seed at read offset rp; returns false if the
// seed holds an N
HLM_HD bool hlm_seed_code(const uint32_t* __restrict__ rd, int rp, uint32_t* code) {
  int rw = rp >> 5, rs = rp & 31;
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
