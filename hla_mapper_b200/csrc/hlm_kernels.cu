// hlm_kernels.cu — hand-written sm_100a kernels of the hla-mapper hot path.
//
//   k_screen       warp per pair: 20-mer screen (preselect.cpp:707-731 / :484-506), mtrim
//                  (functions.cpp:89-145), pair filter (preselect.cpp:1132) and per-locus
//                  candidate mask (map_dna.cpp:611-627) in one streaming pass over the reads.
//   k_pack_units   warp per pair: bit-plane packing of the trimmed mates (both strands).
//   k_seed         warp per unit: 12-mer seed lookups in the column-sorted seed index, hits
//                  flattened over the lanes, de-duplicated in a shared-memory hash set.
//   k_myers        thread per candidate: Hyyro diagonal-band bit-parallel edit distance with the
//                  tile's match vectors in shared memory (lossless lower bound of the DP cost).
//   k_expand       children of the rep candidates that the one-column Lipschitz bound cannot
//                  exclude (two rounds).
//   k_select       thread per candidate: which candidates still need the affine DP.
//   k_dp           thread per candidate: banded affine-gap DP in registers (DPX min/max on the
//                  ALU pipe, adds on the FMA pipe).
//   k_scores       warp per pair, lane per target: SAM-reducer semantics (map_dna.cpp:44-118,
//                  map_rna.cpp:39-101) on the per-task best alignment.
//   k_assign       warp per pair, lane per target: others back-fill, tolerance filter, argmin,
//                  tie set (map_dna.cpp:903-1029, map_rna.cpp:925-1027).
//   k_dpab_*       the DP in two work layouts (evidence for the layout choice, hlm_dp_layout_ab)
//   k_int_peak     INT32 issue-rate micro-kernels (roofline denominators)
//
// All of it is integer / byte work: HBM-, L2- or INT32-bound, no tensor-core shape anywhere.
#include "hlm_kernels.cuh"
#include <cstdlib>
#include <cstring>

#include "hlm_core.cuh"

#define FULL 0xFFFFFFFFu
#define WARPS_PER_BLOCK 8

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
  // streaming read: read bases / qualities are touched once
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// pull the 96-byte tile of a candidate towards L2 while the current candidate is being computed
__device__ __forceinline__ void prefetch_tile(const HlmDevDB& db, uint64_t c) {
  if (c == 0xFFFFFFFFFFFFFFFFull) return;
  const uint8_t* t = (const uint8_t*)(db.tiles + (size_t)HLM_CAND_TILE(c) * HLM_TILE_WORDS);
  prefetch_l2(t);
  prefetch_l2(t + 95);
}

// bytes [8*lane, 8*lane+8) of the byte string p[0..len); zero beyond len.  len <= 256.
__device__ __forceinline__ uint64_t warp_load8(const uint8_t* p, int len, int lane) {
  uintptr_t a = (uintptr_t)p;
  int shift = (int)(a & 7);
  const uint64_t* q = (const uint64_t*)(a - shift);
  int nwords = (shift + len + 7) >> 3;
  uint64_t w = (lane < nwords) ? ld_stream_u64(q + lane) : 0ull;
  uint64_t wn = __shfl_down_sync(FULL, w, 1);
  uint64_t w32 = (nwords > 32 && lane == 0) ? ld_stream_u64(q + 32) : 0ull;
  w32 = __shfl_sync(FULL, w32, 0);
  if (lane == 31) wn = w32;
  uint64_t v = shift ? (w >> (8 * shift)) | (wn << (64 - 8 * shift)) : w;
  int rem = len - 8 * lane;
  if (rem <= 0) v = 0;
  else if (rem < 8) v &= (1ull << (8 * rem)) - 1;
  return v;
}

// 8 ASCII bases -> 2-bit stream (16 bits), plane bits (8 each).  Only upper-case ACGT are
// bases (the reference compares raw strings); everything else, and padding, is N.
// Byte-parallel on the two 32-bit halves: code = ((c >> 1) ^ (c >> 2)) & 3 maps A C G T to
// 0 1 2 3; a byte is a base iff it equals the letter its code stands for
// ('A' ^ (b0 << 1) ^ (b1 * 6) ^ (b0 & b1) * 0x11); the bits of the four bytes of a half are
// gathered with one multiply each.
__device__ __forceinline__ void encode4(uint32_t v, uint32_t& pk, uint32_t& b0, uint32_t& b1, uint32_t& bn) {
  uint32_t code = ((v >> 1) ^ (v >> 2)) & 0x03030303u;
  uint32_t c0 = code & 0x01010101u, c1 = (code >> 1) & 0x01010101u, c01 = c0 & c1;
  uint32_t expect = 0x41414141u ^ (c0 << 1) ^ (c1 * 6u) ^ (c01 * 0x11u);
  uint32_t t = v ^ expect;  // zero byte <=> a base
  // 0x80 in every non-zero byte of t (exact, no carries across bytes)
  uint32_t nz = (((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
  uint32_t keep = 0x01010101u ^ (nz >> 7);  // 1 in every base byte
  c0 &= keep;
  c1 &= keep;
  // gather bit 0 of bytes 0..3 into bits 0..3: (x * 0x01020408) >> 24
  b0 = (c0 * 0x01020408u) >> 24;
  b1 = (c1 * 0x01020408u) >> 24;
  bn = ((nz >> 7) * 0x01020408u) >> 24;
  uint32_t two = c0 | (c1 << 1);  // 2-bit code per byte
  // bytes 0..3 -> bit pairs 0..3: x | x >> 6 | x >> 12 | x >> 18, low byte
  pk = (two | (two >> 6) | (two >> 12) | (two >> 18)) & 0xFFu;
}
__device__ __forceinline__ void encode8(uint64_t v, uint32_t& pk, uint32_t& b0, uint32_t& b1,
                                        uint32_t& bn) {
  uint32_t pl, p0l, p1l, pnl, ph, p0h, p1h, pnh;
  encode4((uint32_t)v, pl, p0l, p1l, pnl);
  encode4((uint32_t)(v >> 32), ph, p0h, p1h, pnh);
  pk = pl | (ph << 8);
  b0 = p0l | (p0h << 4);
  b1 = p1l | (p1h << 4);
  bn = pnl | (pnh << 4);
}

// Warp-level list appender: a warp reserves HLM_APPEND_BLOCK slots with ONE atomic and fills them
// from its ballots; only the unused slots of a warp's LAST block are padded with an invalid record
// that every consumer skips.
// (One atomic per ballot on a single counter serialises in L2 at ~10^7 appends per chunk.)
#define HLM_APPEND_BLOCK 128
#define HLM_CAND_INVALID 0xFFFFFFFFFFFFFFFFull
#define HLM_DP_INVALID 0xFFFFFFFFu
template <class T>
struct WarpAppender {
  unsigned long long pos, end, real;
  T* list;
  unsigned long long* counter;
  unsigned long long* overflow;
  int64_t cap;
  T invalid;
  __device__ __forceinline__ WarpAppender(T* l, unsigned long long* c, unsigned long long* ov,
                                          int64_t cp, T inv)
      : pos(0), end(0), real(0), list(l), counter(c), overflow(ov), cap(cp), invalid(inv) {}
  __device__ __forceinline__ void pad(int lane) {
    for (unsigned long long p = pos + lane; p < end; p += 32)
      if ((int64_t)p < cap) list[p] = invalid;
    pos = end;
  }
  // called by all 32 lanes (converged)
  __device__ __forceinline__ void push(bool take, T v, int lane) {
    uint32_t bal = __ballot_sync(FULL, take);
    if (!bal) return;
    int cnt = __popc(bal);
    unsigned long long rank = __popc(bal & ((1u << lane) - 1u)), at = pos + rank;
    if (pos + cnt > end) {
      // the block is finished with the first records of this push and the rest opens a new one:
      // no padding inside the list (padding slots idle a k_myers lane each)
      unsigned long long room = end - pos, np = 0;
      if (lane == 0) np = atomicAdd(counter, (unsigned long long)HLM_APPEND_BLOCK);
      np = __shfl_sync(FULL, np, 0);
      if (rank >= room) at = np + (rank - room);
      pos = np - room;  // so that pos + cnt below is the next free slot of the new block
      end = np + HLM_APPEND_BLOCK;
    }
    if (take) {
      if ((int64_t)at < cap) list[at] = v;
      else *overflow = 1;
    }
    pos += cnt;
    real += cnt;
  }
  __device__ __forceinline__ void finish(int lane, unsigned long long* real_counter) {
    pad(lane);
    if (lane == 0 && real && real_counter) atomicAdd(real_counter, real);
  }
};

struct WarpScratch {
  uint32_t sq[18];    // 2-bit stream of R1 (16 bases / word) + zero padding
  uint32_t p0[9];     // planes of the mate being processed (+1 pad word)
  uint32_t p1[9];
  uint32_t pn[9];
  uint32_t good[9];   // mtrim "good quality" bits
  uint32_t hits[8];   // global-screen hit bits: [0..3] even R1 positions (bit j <-> 2j), [4..7] odd
  uint32_t rev[3][8]; // scratch for reverse complement
};

__device__ __forceinline__ uint32_t probe_kmer(const HlmDevDB& db, uint64_t key) {
  uint64_t mask = (1ull << db.kt_bits) - 1;
  uint64_t h = hlm_kt_hash(key, db.kt_bits);
  for (;;) {
    uint64_t s = __ldg(db.kt + h);
    if (!(s & HLM_KT_OCC)) return 0u;
    if ((s & HLM_KT_KEYMASK) == key) return __ldg(db.mask_tab + ((s >> 40) & 0xFFFFu));
    h = (h + 1) & mask;
  }
}

// motif lines holding a non-ACGT byte: raw 20-byte compare against the read as written
__device__ __forceinline__ uint32_t probe_odd(const HlmDevDB& db, const uint8_t* w) {
  int lo = 0, hi = db.n_odd;
  while (lo < hi) {
    int mid = (lo + hi) >> 1, c = 0;
    const uint8_t* k = db.odd_keys + (size_t)mid * HLM_MOTIF;
    for (int x = 0; x < HLM_MOTIF && c == 0; x++) c = (int)__ldg(k + x) - (int)w[x];
    if (c == 0) return __ldg(db.odd_masks + mid);
    if (c < 0) lo = mid + 1; else hi = mid;
  }
  return 0u;
}

// first set bit at or after pos (< limit), or -1
__device__ __forceinline__ int next_one(const uint32_t* m, int pos, int limit) {
  while (pos < limit) {
    uint32_t w = m[pos >> 5] >> (pos & 31);
    if (w) {
      int p = pos + __ffs(w) - 1;
      return p < limit ? p : -1;
    }
    pos = (pos | 31) + 1;
  }
  return -1;
}
// first clear bit at or after pos, or limit
__device__ __forceinline__ int next_zero(const uint32_t* m, int pos, int limit) {
  while (pos < limit) {
    uint32_t w = (~m[pos >> 5]) >> (pos & 31);
    if (w) {
      int p = pos + __ffs(w) - 1;
      return p < limit ? p : limit;
    }
    pos = (pos | 31) + 1;
  }
  return limit;
}

// mtrim (functions.cpp:89-145) on the good-quality bit string.  Interior runs score
// end-start = length-1, the run touching the end scores length (end = qual.length()); the
// first strictly larger score wins, starting from 0.
__device__ __forceinline__ bool mtrim_bits(const uint32_t* good, int qlen, int slen, int v_size,
                                           int& lo, int& len) {
  int low = 0, high = 0, pos = 0;
  while (pos < qlen) {
    int s = next_one(good, pos, qlen);
    if (s < 0) break;
    int e = next_zero(good, s, qlen);
    if (e < qlen) {
      if ((e - 1 - s) > (high - low)) { low = s; high = e - 1; }
    } else {
      if ((qlen - s) > (high - low)) { low = s; high = qlen; }
    }
    pos = e;
  }
  len = 0;
  if (low < high) {
    len = high - low + 1;
    if (low > slen) len = 0;
    else if (low + len > slen) len = slen - low;
  }
  lo = low;
  return len >= v_size && len > 0;
}

// load one mate's bases into the warp scratch planes (and the 2-bit stream when sq != null)
__device__ __forceinline__ void stage_bases(const uint8_t* p, int len, int lane, WarpScratch& ws,
                                            bool want_stream) {
  uint64_t v = warp_load8(p, len, lane);
  uint32_t pk, b0, b1, bn;
  encode8(v, pk, b0, b1, bn);
  // assemble words with shuffles: word w of a plane = bytes of lanes 4w..4w+3
  uint32_t x0 = b0 << (8 * (lane & 3)), x1 = b1 << (8 * (lane & 3)), xn = bn << (8 * (lane & 3));
  x0 |= __shfl_xor_sync(FULL, x0, 1); x0 |= __shfl_xor_sync(FULL, x0, 2);
  x1 |= __shfl_xor_sync(FULL, x1, 1); x1 |= __shfl_xor_sync(FULL, x1, 2);
  xn |= __shfl_xor_sync(FULL, xn, 1); xn |= __shfl_xor_sync(FULL, xn, 2);
  uint32_t s2 = pk << (16 * (lane & 1));
  s2 |= __shfl_xor_sync(FULL, s2, 1);
  if ((lane & 3) == 0) {
    ws.p0[lane >> 2] = x0;
    ws.p1[lane >> 2] = x1;
    ws.pn[lane >> 2] = xn;
  }
  if (want_stream && (lane & 1) == 0) ws.sq[lane >> 1] = s2;
  if (lane == 0) {
    ws.p0[8] = 0; ws.p1[8] = 0; ws.pn[8] = 0xFFFFFFFFu;
    if (want_stream) { ws.sq[16] = 0; ws.sq[17] = 0; }
  }
  __syncwarp();
}

__device__ __forceinline__ void stage_quals(const uint8_t* p, int len, int lane, WarpScratch& ws,
                                            const HlmKParams& kp) {
  uint64_t v = warp_load8(p, len, lane);
  uint32_t g = 0;
  if (kp.q_thr >= 0) {
    // good <=> q_thr <= byte < 128, four bytes per 32-bit half, no carries across bytes
    uint32_t thr4 = (uint32_t)kp.q_thr * 0x01010101u;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t w = (uint32_t)(v >> (32 * h));
      uint32_t ge = (((w | 0x80808080u) - thr4) & ~w) & 0x80808080u;
      g |= (((ge >> 7) * 0x01020408u) >> 24) << (4 * h);
    }
    int rem = len - 8 * lane;  // padding bytes are zero: below any threshold unless q_thr == 0
    if (rem < 8) g &= rem <= 0 ? 0u : ((1u << rem) - 1u);
  } else {
#pragma unroll
    for (int x = 0; x < 8; x++) {
      uint32_t c = (uint32_t)(v >> (8 * x)) & 0xFFu;
      uint32_t ok = (kp.good_q[c >> 5] >> (c & 31)) & 1u;
      if (8 * lane + x >= len) ok = 0;
      g |= ok << x;
    }
  }
  uint32_t x0 = g << (8 * (lane & 3));
  x0 |= __shfl_xor_sync(FULL, x0, 1); x0 |= __shfl_xor_sync(FULL, x0, 2);
  if ((lane & 3) == 0) ws.good[lane >> 2] = x0;
  if (lane == 0) ws.good[8] = 0;
  __syncwarp();
}

// ------------------------------------------------------------------ k_prescreen
// The FASTQ loop (preselect.cpp:707-731: stride 2, one extra step after a hit) only leaves the
// even positions of R1 through a hit, so a pair whose even positions hold no 20-mer of
// motif_all.txt is rejected whatever else it contains - the common case on whole-genome / exome
// input.  One THREAD per pair streams R1 as aligned 8-byte words, keeps the last 32 bases as a
// 2-bit window and asks the Bloom front about the four new even positions of every word (four
// independent L2 loads in flight); no shared memory, no shuffles.  A pair with a Bloom-positive
// window (a real hit or a false positive), or anything the shortcut does not cover, goes on a list
// for k_screen, which does the exact work.  Applies to: FASTQ stride rule, minhit >= 1, the screen
// not skipped, no motif lines with non-ACGT bytes.
__device__ __forceinline__ bool prescreen_applies(const HlmDevDB& db, const HlmKParams& kp) {
  return kp.screen_variant == HLM_SCREEN_FASTQ && !kp.skip_screen && kp.minhit >= 1 && db.n_odd == 0;
}
__global__ void __launch_bounds__(128)
k_prescreen(HlmDevDB db, HlmKParams kp, HlmChunk ck) {
  int lane = threadIdx.x & 31;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t nloop = (ck.n_pairs + stride - 1) / stride;
  for (int64_t it = 0; it < nloop; it++) {
    int64_t pair = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool maybe = false, have = pair < ck.n_pairs;
    if (have) {
      uint64_t o1 = ck.offs1[pair];
      int l1 = (int)(ck.offs1[pair + 1] - o1);
      bool ok = !(l1 < HLM_MOTIF || l1 < kp.v_size);
      if (kp.paired) {
        int l2 = (int)(ck.offs2[pair + 1] - ck.offs2[pair]);
        if (l2 < HLM_MOTIF || l2 < kp.v_size) ok = false;
      }
      if (ok) {
        uintptr_t a = (uintptr_t)(ck.bases1 + o1);
        int shift = (int)(a & 7);
        const uint64_t* q = (const uint64_t*)(a - shift);
        int nwords = (l1 + 7) >> 3, npos = l1 - HLM_MOTIF;
        int last_aligned = (shift + l1 - 1) >> 3;  // aligned word that holds the read's last byte
        uint64_t win = 0;
        uint32_t nwin = 0xFFFFFFFFu;
        // four read words per round: their five aligned words are loaded together, the sixteen
        // Bloom words they need are loaded together - two memory round trips per 32 bases
        for (int k0 = 0; k0 < nwords && !maybe; k0 += 4) {
          uint64_t aw[5];
#pragma unroll
          for (int x = 0; x < 5; x++) aw[x] = (k0 + x <= last_aligned) ? ld_stream_u64(q + k0 + x) : 0ull;
          uint64_t bm[16];
          uint32_t bw[16], act = 0;
#pragma unroll
          for (int x = 0; x < 4; x++) {
            int k = k0 + x;
            uint64_t v = shift ? (aw[x] >> (8 * shift)) | (aw[x + 1] << (64 - 8 * shift)) : aw[x];
            int rem = l1 - 8 * k;
            if (rem <= 0) v = 0;
            else if (rem < 8) v &= (1ull << (8 * rem)) - 1;
            uint32_t pk, b0, b1, bn;
            encode8(v, pk, b0, b1, bn);
            // window = bases [8k - 24, 8k + 8): base j at bits 2 (j - 8k + 24)
            win = (win >> 16) | ((uint64_t)pk << 48);
            nwin = (nwin >> 8) | (bn << 24);
            // the even positions that became complete with this word: 8k - 18, -16, -14, -12
#pragma unroll
            for (int y = 0; y < 4; y++) {
              int off = 6 + 2 * y, c = 8 * k - 24 + off;
              if (c >= 0 && c < npos && ((nwin >> off) & 0xFFFFFu) == 0) act |= 1u << (4 * x + y);
              uint64_t key = (win >> (2 * off)) & HLM_KT_KEYMASK;
              bm[4 * x + y] = hlm_kb_probe(key, db.kb_words, &bw[4 * x + y]);
            }
          }
          uint64_t word[16];
#pragma unroll
          for (int z = 0; z < 16; z++) word[z] = ((act >> z) & 1u) ? __ldg(db.kb + bw[z]) : 0ull;
#pragma unroll
          for (int z = 0; z < 16; z++) maybe = maybe || (((act >> z) & 1u) && (word[z] & bm[z]) == bm[z]);
        }
        // (the last probed position, l1 - 21, is complete with the last word: 8 (nwords - 1) - 12 >= l1 - 21)
      }
      if (!maybe) {  // rejected: what k_screen writes for a pair that fails the screen
        ck.flags[pair] = 0;
        ck.trim_lo1[pair] = 0;
        ck.trim_len1[pair] = 0;
        ck.trim_lo2[pair] = 0;
        ck.trim_len2[pair] = 0;
        ck.locus_mask[pair] = 0;
      }
    }
    // the others, in pair order inside the warp, one atomic per warp
    uint32_t bal = __ballot_sync(FULL, maybe);
    if (bal) {
      unsigned long long base = 0;
      if (lane == __ffs(bal) - 1) base = atomicAdd(&ck.counters[HLM_N_COUNTERS], (unsigned long long)__popc(bal));
      base = __shfl_sync(FULL, base, __ffs(bal) - 1);
      if (maybe) ck.pre_list[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)pair;
    }
  }
}

// ------------------------------------------------------------------ k_screen
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_screen(HlmDevDB db, HlmKParams kp, HlmChunk ck) {
  __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpScratch& ws = scratch[wib];
  int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  unsigned long long kept_local = 0;
  // with the prefilter (k_prescreen) only the pairs it could not reject arrive here, as a list
  const bool listed = prescreen_applies(db, kp);
  int64_t n_items = listed ? (int64_t)ck.counters[HLM_N_COUNTERS] : ck.n_pairs;
  for (int64_t item = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib; item < n_items; item += nwarps) {
    int64_t pair = listed ? (int64_t)ck.pre_list[item] : item;
    uint64_t o1 = ck.offs1[pair];
    int l1 = (int)(ck.offs1[pair + 1] - o1);
    uint64_t o2 = 0;
    int l2 = 0;
    if (kp.paired) {
      o2 = ck.offs2[pair];
      l2 = (int)(ck.offs2[pair + 1] - o2);
    }
    // length pre-checks: preselect.cpp:707-710 (paired FASTQ), :484-485 (BAM, R1 only), :918-919
    bool ok = !(l1 < HLM_MOTIF || l1 < kp.v_size);
    if (kp.paired && kp.screen_variant == HLM_SCREEN_FASTQ && (l2 < HLM_MOTIF || l2 < kp.v_size))
      ok = false;
    // membership masks of the R1 windows at positions 2 * (32 r + lane) (even) and + 1 (odd)
    uint32_t val_e[4], val_o[4];
#pragma unroll
    for (int r = 0; r < 4; r++) val_e[r] = val_o[r] = 0;
    bool screened = false;
    if (ok) {
      stage_bases(ck.bases1 + o1, l1, lane, ws, true);
      int npos = l1 - HLM_MOTIF;  // positions c < npos are probed (the last window never is)
      int n_even = (npos + 1) >> 1, n_oddpos = npos >> 1;
      auto probe_at = [&](int c) -> uint32_t {
        int w = c >> 4, sh = 2 * (c & 15);
        uint64_t lo = (uint64_t)ws.sq[w] | ((uint64_t)ws.sq[w + 1] << 32);
        uint64_t key = lo >> sh;
        if (sh) key |= (uint64_t)ws.sq[w + 2] << (64 - sh);
        key &= HLM_KT_KEYMASK;
        uint32_t nn = hlm_funnel(ws.pn[c >> 5], ws.pn[(c >> 5) + 1], c & 31) & 0xFFFFFu;
        if (nn == 0) {
          // Bloom front first: one 8-byte load that hits L2 (4 bytes per key); only windows
          // it cannot exclude go on to the exact table, which does not fit L2
          uint32_t bw;
          uint64_t bm = hlm_kb_probe(key, db.kb_words, &bw);
          return ((__ldg(db.kb + bw) & bm) == bm) ? probe_kmer(db, key) : 0u;
        }
        return db.n_odd ? probe_odd(db, ck.bases1 + o1 + c) : 0u;
      };
      // Phase 1: the even positions.  The FASTQ loop (stride 2, one extra step after a hit) only
      // ever leaves them through a hit, so a read without an even hit is rejected after half
      // the probes - the common case for off-target input.  Phase 2: the odd positions, for
      // everything else (the BAM loop has stride 1; a kept pair needs every position for its
      // locus mask).
      uint32_t any_e = 0;
      int nhits = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        uint32_t hb = 0;
        if (r * 32 < n_even) {
          int j = r * 32 + lane;
          uint32_t v = j < n_even ? probe_at(2 * j) : 0u;
          val_e[r] = v;
          hb = __ballot_sync(FULL, v >> 31);
        }
        if (lane == 0) ws.hits[r] = hb;
        any_e |= hb;
        nhits += __popc(hb);
      }
      bool need_odd = kp.skip_screen || kp.screen_variant != HLM_SCREEN_FASTQ || any_e || kp.minhit <= 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        uint32_t hb = 0;
        if (need_odd && r * 32 < n_oddpos) {
          int j = r * 32 + lane;
          uint32_t v = j < n_oddpos ? probe_at(2 * j + 1) : 0u;
          val_o[r] = v;
          hb = __ballot_sync(FULL, v >> 31);
        }
        if (lane == 0) ws.hits[4 + r] = hb;
        nhits += __popc(hb);
      }
      __syncwarp();
      if (kp.skip_screen) {
        screened = true;
      } else if (nhits >= kp.minhit) {
        // the reference's strided scan: FASTQ loop `if hit {hit++;c++}; ...; c++` inside
        // for(;;c++); BAM loop without the trailing c++.  Uniform across the warp.
        int hitcount = 0;
        int step = kp.screen_variant == HLM_SCREEN_FASTQ ? 2 : 1;
        for (int c = 0; c < npos; c += step) {
          if ((ws.hits[(c & 1) * 4 + (c >> 6)] >> ((c >> 1) & 31)) & 1u) {
            hitcount++;
            c++;
          }
          if (hitcount == kp.minhit) {
            screened = true;
            break;
          }
        }
      }
    }
    int lo1 = 0, n1 = 0, lo2 = 0, n2 = 0;
    uint32_t fl = 0, lmask = 0;
    if (screened) {
      fl |= HLM_F_SCREENED;
      stage_quals(ck.quals1 + o1, l1, lane, ws, kp);
      bool k1 = mtrim_bits(ws.good, l1, l1, kp.v_size, lo1, n1);
      bool k2 = true;
      __syncwarp();
      if (kp.paired) {
        stage_quals(ck.quals2 + o2, l2, lane, ws, kp);
        k2 = mtrim_bits(ws.good, l2, l2, kp.v_size, lo2, n2);
        __syncwarp();
      }
      if (!k1) n1 = 0;
      if (!k2) n2 = 0;
      if (k1 && k2) {
        fl |= HLM_F_KEPT;
        // map_dna.cpp:611-627: any 20-mer of the trimmed R1 at c < len-20, both mates >= 20
        if (n1 >= HLM_MOTIF && (!kp.paired || n2 >= HLM_MOTIF)) {
          uint32_t m = 0;
#pragma unroll
          for (int r = 0; r < 4; r++) {
            int c = 2 * (r * 32 + lane);
            if (c >= lo1 && c < lo1 + n1 - HLM_MOTIF) m |= val_e[r];
            if (c + 1 >= lo1 && c + 1 < lo1 + n1 - HLM_MOTIF) m |= val_o[r];
          }
          m = __reduce_or_sync(FULL, m);
          lmask = m & 0x7FFFFFFFu;
        }
      }
    }
    if (lane == 0) {
      ck.flags[pair] = (uint8_t)fl;
      ck.trim_lo1[pair] = (uint16_t)lo1;
      ck.trim_len1[pair] = (uint16_t)n1;
      ck.trim_lo2[pair] = (uint16_t)lo2;
      ck.trim_len2[pair] = (uint16_t)n2;
      ck.locus_mask[pair] = lmask;
      if (fl & HLM_F_KEPT) kept_local++;
    }
    __syncwarp();
  }
  if (lane == 0 && kept_local) atomicAdd(&ck.counters[3], kept_local);
}

// ------------------------------------------------------------------ k_pack_units
// unit layout per pair: dna: [mate]; rna: [mate*3 + {whole, block A, block B}]
__device__ __forceinline__ void write_unit(const HlmChunk& ck, int64_t unit, WarpScratch& ws, int lo,
                                           int n, uint32_t tmask, int lane, int cap = 0) {
  uint32_t* out = ck.unit_planes + unit * HLM_UNIT_STRIDE;
  // forward planes: bits [lo, lo+n) of the staged planes
  uint32_t f0 = 0, f1 = 0, fn = 0xFFFFFFFFu;
  if (lane < 6) {
    int b = lo + 32 * lane;
    int w = b >> 5, s = b & 31;
    uint32_t a0 = w < 8 ? ws.p0[w] : 0u, a1 = w + 1 < 8 ? ws.p0[w + 1] : 0u;
    uint32_t c0 = w < 8 ? ws.p1[w] : 0u, c1 = w + 1 < 8 ? ws.p1[w + 1] : 0u;
    uint32_t e0 = w < 8 ? ws.pn[w] : 0xFFFFFFFFu, e1 = w + 1 < 8 ? ws.pn[w + 1] : 0xFFFFFFFFu;
    f0 = hlm_funnel(a0, a1, s);
    f1 = hlm_funnel(c0, c1, s);
    fn = hlm_funnel(e0, e1, s);
    int rem = n - 32 * lane;
    uint32_t valid = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    fn |= ~valid;
    f0 &= ~fn;
    f1 &= ~fn;
    out[lane] = f0;
    out[6 + lane] = f1;
    out[12 + lane] = fn;
    // bit-reversed words for the reverse complement (word 5-lane of the 192-bit reversal)
    ws.rev[0][5 - lane] = __brev(f0);
    ws.rev[1][5 - lane] = __brev(f1);
    ws.rev[2][5 - lane] = __brev(fn);
  }
  if (lane >= 6 && lane < 8) {
    ws.rev[0][lane] = 0;
    ws.rev[1][lane] = 0;
    ws.rev[2][lane] = 0xFFFFFFFFu;
  }
  __syncwarp();
  if (lane < 6) {
    int sh = 192 - n;  // rc bit i = reversed bit i + sh
    int w = lane + (sh >> 5), s = sh & 31;
    uint32_t a0 = w < 6 ? ws.rev[0][w] : 0u, a1 = w + 1 < 6 ? ws.rev[0][w + 1] : 0u;
    uint32_t c0 = w < 6 ? ws.rev[1][w] : 0u, c1 = w + 1 < 6 ? ws.rev[1][w + 1] : 0u;
    uint32_t e0 = w < 6 ? ws.rev[2][w] : 0xFFFFFFFFu, e1 = w + 1 < 6 ? ws.rev[2][w + 1] : 0xFFFFFFFFu;
    uint32_t rn = hlm_funnel(e0, e1, s);
    int rem = n - 32 * lane;
    uint32_t valid = rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
    rn |= ~valid;
    uint32_t r0 = ~hlm_funnel(a0, a1, s) & ~rn;  // complement = 3 - code
    uint32_t r1 = ~hlm_funnel(c0, c1, s) & ~rn;
    out[18 + lane] = r0;
    out[24 + lane] = r1;
    out[30 + lane] = rn;
  }
  if (lane == 0) {
    ck.unit_len[unit] = (uint16_t)n;
    ck.unit_mask[unit] = tmask;
    ck.unit_cap[unit] = (uint16_t)cap;
  }
  __syncwarp();
  // 4-bit codes of both strands for k_myers: word j = bases 8j .. 8j+7 (lanes 0-23 one strand each
  // pass); the planes were just written by this warp
  for (int strand = 0; strand < 2; strand++) {
    if (lane < HLM_CODE_WORDS) {
      const uint32_t* P = out + 18 * strand;
      int w = lane >> 2, sh = 8 * (lane & 3);
      uint32_t x0 = (P[w] >> sh) & 0xFFu, x1 = (P[6 + w] >> sh) & 0xFFu, xn = (P[12 + w] >> sh) & 0xFFu;
      out[2 * HLM_UNIT_WORDS + HLM_CODE_WORDS * strand + lane] =
          hlm_spread8(x0) | (hlm_spread8(x1) << 1) | (hlm_spread8(xn) << 2);
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_pack_units(HlmKParams kp, HlmChunk ck) {
  __shared__ WarpScratch scratch[WARPS_PER_BLOCK];
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpScratch& ws = scratch[wib];
  int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  int U = ck.units_per_pair;
  int nm = kp.paired ? 2 : 1;
  for (int64_t pair = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib; pair < ck.n_pairs; pair += nwarps) {
    if (lane < U) ck.unit_len[pair * U + lane] = 0;
    int64_t b0 = 0, b1 = 0;
    if (kp.rna && ck.blk_off) {
      b0 = (int64_t)ck.blk_off[pair];
      b1 = (int64_t)ck.blk_off[pair + 1];
      for (int64_t b = b0 + lane; b < b1; b += 32) ck.unit_len[ck.n_pairs * U + (b - ck.blk_first)] = 0;
    }
    __syncwarp();
    if (!(ck.flags[pair] & HLM_F_KEPT)) continue;
    uint32_t lmask = ck.locus_mask[pair];
    uint32_t others = 1u << kp.n_loci;
    for (int m = 0; m < nm; m++) {
      const uint8_t* p = m == 0 ? ck.bases1 + ck.offs1[pair] : ck.bases2 + ck.offs2[pair];
      int l = m == 0 ? (int)(ck.offs1[pair + 1] - ck.offs1[pair]) : (int)(ck.offs2[pair + 1] - ck.offs2[pair]);
      int lo = m == 0 ? ck.trim_lo1[pair] : ck.trim_lo2[pair];
      int n = m == 0 ? ck.trim_len1[pair] : ck.trim_len2[pair];
      if (n > HLM_MAX_READ) {  // does not fit the 256-column tiles: reported by the host
        if (lane == 0) ck.counters[7] = 1;
        continue;
      }
      stage_bases(p, l, lane, ws, false);
      if (kp.allele_locus >= 0) {
        // hlm_score_alleles: the whole mate against one target set, whatever its locus mask
        write_unit(ck, pair * U + (kp.rna ? m * 3 : m), ws, lo, n, 1u << kp.allele_locus, lane);
      } else if (!kp.rna) {
        // the compare step's tolerance for this mate, map_dna.cpp:948-958 (single end: the length
        // the reference holds for the read, preselect.cpp:1199)
        int len = (!kp.paired && ck.size_override1) ? ck.size_override1[pair] : n;
        int cap = (int)(len * kp.tolerance);
        cap = cap < 0 ? 0 : (cap > 60000 ? 60000 : cap);
        write_unit(ck, pair * U + m, ws, lo, n, lmask | others, lane, cap);
      } else if (ck.blk_off) {
        // the STAR pre-split in full: whole mate vs others only; every block vs its own gene
        write_unit(ck, pair * U + m * 3, ws, lo, n, others, lane);
        for (int64_t b = b0; b < b1; b++) {
          hlm_rna_block blk = ck.blk[b];
          if (blk.mate != m || blk.gene >= kp.n_loci) continue;
          int st = blk.start < n ? blk.start : n;
          int bl = blk.len < n - st ? blk.len : n - st;
          if (bl > 0) write_unit(ck, ck.n_pairs * U + (b - ck.blk_first), ws, lo + st, bl, 1u << blk.gene, lane);
        }
      } else {
        const uint16_t* sp = m == 0 ? ck.rna_split1 : ck.rna_split2;
        int split = sp ? sp[pair] : 0;
        if (!(split > 0 && split < n)) split = 0;
        // whole mate: always vs others (map_rna.cpp:858-892); vs loci only when unsplit
        write_unit(ck, pair * U + m * 3, ws, lo, n, split ? others : (lmask | others), lane);
        if (split) {
          write_unit(ck, pair * U + m * 3 + 1, ws, lo, split, lmask, lane);
          write_unit(ck, pair * U + m * 3 + 2, ws, lo + split, n - split, lmask, lane);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ k_clear_tasks
__global__ void k_clear_tasks(HlmChunk ck, int64_t n_tasks) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (i <= HLM_N_COUNTERS) ck.counters[i] = 0;  // [HLM_N_COUNTERS]: length of the k_prescreen list
  if (i < 256) ck.dp_hist[i] = 0;
  for (; i < n_tasks; i += stride) {
    ck.task_emin[i] = 0xFFFFFFFFu;
    ck.task_pmin[i] = 0xFFFFFFFFu;
    ck.task_best[i] = HLM_BEST_NONE;
  }
}

// ------------------------------------------------------------------ k_seed
__device__ __forceinline__ int tile_locus(const HlmDevDB& db, uint32_t tile) {
  int g = 0;
  // tile_start is ascending; n_loci + 1 target sets
  int lo = 0, hi = db.n_loci + 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (tile >= db.tile_start[mid]) lo = mid; else hi = mid;
  }
  g = lo;
  return g;
}

// A (tile, off, strand) hit by several read seeds must become ONE candidate.  Each warp keeps the
// keys it has emitted for the current unit in a shared-memory hash set (exact: full key
// compare), so duplicates are dropped without touching the tile data; units with more index hits
// than the set can hold fall back to testing the earlier seeds against the tile.
#define SEED_WARPS 4
#define SEED_HT 1024
#define SEED_HT_MAX 700
__global__ void __launch_bounds__(SEED_WARPS * 32)
k_seed(HlmDevDB db, HlmKParams kp, HlmChunk ck) {
  __shared__ uint32_t s_rd[SEED_WARPS][36];
  __shared__ uint32_t s_pre[SEED_WARPS][32], s_beg[SEED_WARPS][32];
  __shared__ uint32_t s_ht[SEED_WARPS][SEED_HT];
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* rd = s_rd[wib];
  uint32_t* pre = s_pre[wib];
  uint32_t* beg = s_beg[wib];
  uint32_t* ht = s_ht[wib];
  int64_t nwarps = (int64_t)gridDim.x * SEED_WARPS;
  int64_t n_units = ck.n_units;
  WarpAppender<uint64_t> app(ck.cand, &ck.counters[0], &ck.counters[2], ck.cand_cap, HLM_CAND_INVALID);
  for (int64_t unit = (int64_t)blockIdx.x * SEED_WARPS + wib; unit < n_units; unit += nwarps) {
    int n = ck.unit_len[unit];
    if (n == 0) continue;
    uint32_t umask = ck.unit_mask[unit];
    __syncwarp();
    rd[lane] = ck.unit_planes[unit * HLM_UNIT_STRIDE + lane];
    if (lane < 4) rd[32 + lane] = ck.unit_planes[unit * HLM_UNIT_STRIDE + 32 + lane];
    __syncwarp();
    int Q = hlm_seed_count(n, kp.rna), st = hlm_seed_stride(n, kp.rna);  // Q <= 15
    // lane (strand*16 + q) owns seed q of that strand (read offset q * st): CSR range of its
    // 12-mer narrowed to the entries whose column lies in [OFF_MIN + q st, OFF_MIN + 32 + q st)
    // (entries are sorted by column), i.e. to the windows in which read base 0 falls on a column
    // in [OFF_MIN, OFF_MIN+32)
    int my_q = lane & 15, my_strand = lane >> 4;
    uint32_t my_a = 0, my_b = 0;
    if (my_q < Q) {
      uint32_t code;
      if (hlm_seed_code(rd + 18 * my_strand, my_q * st, &code)) {
        uint32_t a = __ldg(db.seed_off + code), b = __ldg(db.seed_off + code + 1);
        uint32_t klo = (uint32_t)(HLM_OFF_MIN + my_q * st) << 24;
        uint32_t khi = (uint32_t)(HLM_OFF_MIN + HLM_TILE_STRIDE + my_q * st) << 24;
        // both bounds at once: two independent chains of dependent loads instead of one after
        // the other (this part of the kernel is pure memory latency)
        uint32_t lo1 = a, hi1 = b, lo2 = a, hi2 = b;
        if (b - a <= 16) {
          // a small bucket (the common case: 5 entries on average) is counted through instead of
          // searched: its loads do not depend on one another, a binary search's do
          uint32_t below_lo = 0, below_hi = 0;
#pragma unroll 4
          for (uint32_t x = a; x < b; x++) {
            uint32_t e = __ldg(db.seed_ent + x);
            below_lo += e < klo ? 1u : 0u;
            below_hi += e < khi ? 1u : 0u;
          }
          lo1 = hi1 = a + below_lo;
          lo2 = hi2 = a + below_hi;
        }
        while (lo1 < hi1 || lo2 < hi2) {
          uint32_t m1 = (lo1 + hi1) >> 1, m2 = (lo2 + hi2) >> 1;
          uint32_t e1 = lo1 < hi1 ? __ldg(db.seed_ent + m1) : 0u;
          uint32_t e2 = lo2 < hi2 ? __ldg(db.seed_ent + m2) : 0u;
          if (lo1 < hi1) {
            if (e1 < klo) lo1 = m1 + 1; else hi1 = m1;
          }
          if (lo2 < hi2) {
            if (e2 < khi) lo2 = m2 + 1; else hi2 = m2;
          }
        }
        my_a = lo1;
        my_b = lo2;
      }
    }
    // flatten the 2 x Q entry ranges over the lanes (load-balanced expansion): entry t of the
    // concatenation belongs to the lane j with pre[j] <= t < pre[j + 1]
    uint32_t cnt = my_b - my_a, inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += v;
    }
    uint32_t total = __shfl_sync(FULL, inc, 31);
    pre[lane] = inc - cnt;  // exclusive prefix
    beg[lane] = my_a;
    // The set holds one key per candidate emitted so far (distinct keys, not index hits: a unit
    // with thousands of hits has a few hundred candidates).  While it has room a hit is a
    // candidate iff its key is new.  Once SEED_HT_MAX keys are in, new keys are no longer stored:
    // a hit whose key is not in the set is then tested the slow way - it is the candidate's first
    // hit iff no earlier read seed matches the tile (hits arrive in seed order, so both rules
    // name the same owner).
    for (int x = lane; x < SEED_HT; x += 32) ht[x] = 0xFFFFFFFFu;
    int n_keys = 0;  // warp-uniform
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32) {
      uint32_t t = base + lane;
      bool own = false;
      uint32_t tile = 0;
      int off = 0, strand = 0;
      bool insert = n_keys + 32 <= SEED_HT_MAX;
      if (t < total) {
        int j = 0;  // last lane with pre[j] <= t
#pragma unroll
        for (int o = 16; o; o >>= 1)
          if (j + o < 32 && pre[j + o] <= t) j += o;
        int q = j & 15;
        strand = j >> 4;
        uint32_t e = __ldg(db.seed_ent + beg[j] + (t - pre[j]));
        tile = e & 0xFFFFFFu;
        off = (int)(e >> 24) - q * st;
        int g = tile_locus(db, tile);
        if ((umask >> g) & 1u) {
          uint32_t key = (tile << 6) | ((uint32_t)(off - HLM_OFF_MIN) << 1) | (uint32_t)strand;
          uint32_t h = (key * 2654435761u) >> 22;
          bool found = false;
          for (;;) {
            uint32_t old = insert ? atomicCAS(&ht[h], 0xFFFFFFFFu, key) : ht[h];
            if (old == 0xFFFFFFFFu) break;  // not in the set (and now stored, when inserting)
            if (old == key) {
              found = true;
              break;
            }
            h = (h + 1) & (SEED_HT - 1);
          }
          if (!found && insert) {
            own = true;
          } else if (!found) {
            // the set is full: the candidate belongs to the first read seed whose index entry
            // hits it (children only list the seeds that cover one of their differing columns)
            const uint32_t* R = rd + 18 * strand;
            const uint32_t* T = db.tiles + (size_t)tile * HLM_TILE_WORDS;
            uint32_t diff = __ldg(db.tile_diff + tile);
            own = true;
            for (int qq = 0; qq < q && own; qq++) {
              if (!hlm_seed_match(R, T, off, qq * st)) continue;
              if (HLM_DIFF_N(diff) == 0) own = false;
              int o = off + qq * st;
              for (int k = 0; k < HLM_DIFF_N(diff); k++) {
                int x = HLM_DIFF_COL(diff, k);
                if (x >= o && x < o + HLM_SEED) own = false;
              }
            }
          }
        }
      }
      if (insert) n_keys += __popc(__ballot_sync(FULL, own));
      __syncwarp();
      app.push(own, HLM_CAND(unit, tile, off - HLM_OFF_MIN, strand, 0xFF), lane);
    }
  }
  app.finish(lane, &ck.counters[11]);
}

// counters[8 + which] = number of candidates appended so far (phase boundaries)
__global__ void k_snapshot(HlmChunk ck, int which) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long n = ck.counters[0];
    if ((int64_t)n > ck.cand_cap) n = (unsigned long long)ck.cand_cap;
    ck.counters[8 + which] = n;
    ck.counters[15] = 0;  // work counter of the k_myers launch that follows
    if (which == 0) ck.counters[12] = ck.counters[11];
  }
}

// phase 0: seed-index candidates; 1 / 2: children appended by expansion round 0 / 1
__device__ __forceinline__ void cand_range(const HlmChunk& ck, int phase, int64_t& a, int64_t& b) {
  a = phase == 0 ? 0 : (int64_t)ck.counters[8 + phase - 1];
  b = (int64_t)ck.counters[8 + phase];
}

// ------------------------------------------------------------------ k_myers
// per-thread PEQ rows interleaved across the block: row r of thread t at s_peq[r * 256 + t]
// (bank = lane: conflict-free whatever row each lane reads)
struct HlmPeqShared {
  uint32_t* base;
  __device__ __forceinline__ uint32_t& operator()(int r) const { return base[r * 256]; }
};

__global__ void __launch_bounds__(256)
k_myers(HlmDevDB db, HlmKParams kp, HlmChunk ck, int n_targets, int phase) {
  __shared__ uint32_t s_peq[HLM_PEQ_ROWS * 256];
  __shared__ uint16_t s_mintab[256];
  s_mintab[threadIdx.x] = hlm_band_min_entry((int)threadIdx.x);  // blockDim.x == 256
  __syncthreads();
  HlmPeqShared peq{s_peq + threadIdx.x};
  int64_t first, ncand;
  cand_range(ck, phase, first, ncand);
  unsigned long long rows = 0;
  int lane_ = threadIdx.x & 31;
  // warps pull groups of 32 candidates from a device counter (dead / short candidates make a
  // static split uneven); the next group's tiles are prefetched towards L2 meanwhile
  for (;;) {
    unsigned long long base = 0;
    if (lane_ == 0) base = atomicAdd(&ck.counters[15], 32ull);
    base = __shfl_sync(FULL, base, 0);
    int64_t i = first + (int64_t)base + lane_;
    if (first + (int64_t)base >= ncand) break;
    if (i >= ncand) continue;
    uint64_t c = ck.cand[i];
    if (i + 32 * (int64_t)gridDim.x * (blockDim.x >> 5) < ncand)
      prefetch_tile(db, ck.cand[i + 32 * (int64_t)gridDim.x * (blockDim.x >> 5)]);
    if (c == HLM_CAND_INVALID) continue;
    uint32_t unit = HLM_CAND_UNIT(c), tile = HLM_CAND_TILE(c);
    int off = (int)HLM_CAND_OFFM(c) + HLM_OFF_MIN, strand = (int)HLM_CAND_STRAND(c);
    int n = ck.unit_len[unit];
    const uint32_t* rd = ck.unit_planes + (size_t)unit * HLM_UNIT_STRIDE + 18 * strand;
    const uint32_t* T = db.tiles + (size_t)tile * HLM_TILE_WORDS;
    int lb, plb = 0;
    bool dead = false;
    if (phase == 1 || phase == 2) {
      // a child reached by expanding its rep: it is this path's candidate only when some seed
      // matches and none of its matching seeds covers a differing column (a covering seed is in
      // the seed index, which then emitted the candidate)
      uint32_t diff = __ldg(db.tile_diff + tile);
      int Q = hlm_seed_count(n, kp.rna), st = hlm_seed_stride(n, kp.rna);
      // the seeds that cover a differing column (at most KMAX with non-overlapping seeds, two per
      // column with the overlapping seeds of short units) decide ownership; after them the first
      // matching seed settles "some seed matches": a few tests, not Q
      uint32_t cover = 0;
      for (int k = 0; k < HLM_DIFF_N(diff); k++) {
        int r = HLM_DIFF_COL(diff, k) - off;  // read position on the diagonal
        if (r < 0) continue;
        // seeds q with q st <= r < q st + 12
        int q1 = r / st, q0 = (r - HLM_SEED + st) / st;  // q0 = ceil((r - 11) / st), >= 0 below
        if (r < HLM_SEED) q0 = 0;
        if (q1 > Q - 1) q1 = Q - 1;
        for (int q = q0; q <= q1; q++) cover |= 1u << q;
      }
      for (uint32_t m = cover; m && !dead; m &= m - 1)
        dead = hlm_seed_match(rd, T, off, (__ffs(m) - 1) * st);  // the seed index lists this one
      bool any = dead;
      for (int q = 0; q < Q && !any; q++)
        if (!((cover >> q) & 1u)) any = hlm_seed_match(rd, T, off, q * st);
      if (!any) dead = true;
    }
    if (dead) {
      lb = 254;
    } else {
      hlm_peq_build(T, peq);
      // capped scoring: the bound after the rows a trailing clip (<= mm_max bases of a reported
      // alignment) cannot reach; with both clips in the score the bound of the whole read
      int rstar = -1;
      if (kp.score_cap && kp.clip_mode == HLM_CLIP_LEAD_ONLY) rstar = n > kp.mm_max ? ((n - kp.mm_max) & ~7) : 0;
      lb = hlm_myers_core_dev<256 * 4>(ck.unit_planes + (size_t)unit * HLM_UNIT_STRIDE + 2 * HLM_UNIT_WORDS +
                                           HLM_CODE_WORDS * strand,
                                       n, off, (uint32_t)__cvta_generic_to_shared(s_peq + threadIdx.x), (uint32_t)kp.one,
                                       s_mintab, rstar, &plb);
      if (rstar < 0) plb = lb;
      if (lb > 253) lb = 253;
      rows += n;
    }
    ck.cand[i] = (c & ~0xFFull) | (uint64_t)lb;
    if (lb <= kp.mm_max) {
      int g = tile_locus(db, tile);
      atomicMin(&ck.task_emin[(int64_t)unit * n_targets + g], (uint32_t)lb);
    }
    // (a rep up to KMAX above mm_max can still hand a reportable child to the expansion)
    if (kp.score_cap && lb <= kp.mm_max + HLM_KMAX)
      atomicMin(&ck.task_pmin[(int64_t)unit * n_targets + tile_locus(db, tile)], (uint32_t)plb);
  }
  // one atomic per warp
  for (int o = 16; o; o >>= 1) rows += __shfl_xor_sync(FULL, rows, o);
  if ((threadIdx.x & 31) == 0 && rows) atomicAdd(&ck.counters[5], rows);
}

// ------------------------------------------------------------------ capped scoring: task gate
// hlm_params.score_cap (dna): the compare step (map_dna.cpp:948-958) can only use the score of a
// (mate, target) when s - 1 = cost - trailing clip <= cap = int(len * tolerance) - for BOTH mates
// of a pair.  task_pmin = minimum over the task's candidates of a lower bound of cost - trailing
// clip; a child that is not materialised yet differs from a materialised rep in <= KMAX columns,
// so  pmin > cap + KMAX  proves that no alignment of the task can pass, at every stage of the
// flow.  Such a task gets no child expansion, no selection and no DP; its score is reported as 0
// ("none"), which the compare step treats exactly like the failing score it stands for.
// `others` is different: the mere presence of a score feeds the mate back-fill
// (map_dna.cpp:903-915), so it stays exact whenever EITHER mate could pass.
__device__ __forceinline__ bool task_live(const HlmKParams& kp, const HlmChunk& ck, uint32_t unit, int g,
                                          int n_targets) {
  if (!kp.score_cap) return true;
  bool own = ck.task_pmin[(int64_t)unit * n_targets + g] <= (uint32_t)ck.unit_cap[unit] + HLM_KMAX;
  if (!kp.paired) return own;
  uint32_t mate = unit ^ 1u;  // dna units: pair * 2 + mate
  bool other = ck.task_pmin[(int64_t)mate * n_targets + g] <= (uint32_t)ck.unit_cap[mate] + HLM_KMAX;
  return g < kp.n_loci ? (own && other) : (own || other);
}

// ------------------------------------------------------------------ k_expand
// Children of the rep candidates.  A child differs from its rep in nd <= KMAX columns, so its
// lower bound is >= lbd = max(0, lb_rep - nd) (one changed reference column changes the banded
// unit-cost distance by at most one).  Round 0 appends the children with lbd <= min(emin0, mm_max)
// (emin0 = task minimum over the seed-index candidates): every child left out has lb > emin0.
// Round 1 appends those with emin0 < lbd <= min(cost(best), mm_max).  Children whose differing
// columns all lie outside the DP footprint [off - BAND, off + n + BAND) align exactly like the rep
// and are never materialised.
__global__ void __launch_bounds__(256)
k_expand(HlmDevDB db, HlmKParams kp, HlmChunk ck, int n_targets, int round) {
  int64_t nA = (int64_t)ck.counters[8];
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int lane = threadIdx.x & 31;
  int64_t nloop = (nA + stride - 1) / stride;
  WarpAppender<uint64_t> app(ck.cand, &ck.counters[0], &ck.counters[2], ck.cand_cap, HLM_CAND_INVALID);
  for (int64_t it = 0; it < nloop; it++) {
    int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    uint64_t c = 0;
    int lb = 0, thr_lo = -1, thr_hi = -1;
    if (i < nA) {
      c = ck.cand[i];
      uint32_t tile = HLM_CAND_TILE(c);
      lb = (int)HLM_CAND_LB(c);
      if (lb < 254 && __ldg(db.child_off + tile + 1) > __ldg(db.child_off + tile)) {
        int g = tile_locus(db, tile);
        int64_t task = (int64_t)HLM_CAND_UNIT(c) * n_targets + g;
        uint32_t e0 = ck.task_emin0[task];
        int e0c = e0 > (uint32_t)kp.mm_max ? kp.mm_max : (int)e0;
        if (round == 2) {
          // per-allele scoring: every child that can hold a reportable alignment
          thr_hi = kp.mm_max;
          pass = lb - HLM_KMAX <= thr_hi;
        } else if (!task_live(kp, ck, HLM_CAND_UNIT(c), g, n_targets)) {
          // capped scoring: nothing of this task can pass the tolerance
        } else if (round == 0) {
          thr_hi = e0c;
          pass = lb - HLM_KMAX <= thr_hi;
        } else {
          unsigned long long b = ck.task_best[task];
          if (b != HLM_BEST_NONE && e0 <= (uint32_t)kp.mm_max) {
            thr_lo = e0c;
            thr_hi = HLM_BEST_COST(b) < kp.mm_max ? HLM_BEST_COST(b) : kp.mm_max;
            pass = lb > thr_lo && lb - HLM_KMAX <= thr_hi;
          }
        }
      }
    }
    uint32_t bal = __ballot_sync(FULL, pass);
    while (bal) {
      int src = __ffs(bal) - 1;
      bal &= bal - 1;
      uint64_t cc = __shfl_sync(FULL, c, src);
      int lbr = __shfl_sync(FULL, lb, src), lo = __shfl_sync(FULL, thr_lo, src),
          hi = __shfl_sync(FULL, thr_hi, src);
      uint32_t unit = HLM_CAND_UNIT(cc), tile = HLM_CAND_TILE(cc);
      int off = (int)HLM_CAND_OFFM(cc) + HLM_OFF_MIN;
      int n = ck.unit_len[unit];
      uint32_t a = __ldg(db.child_off + tile), b = __ldg(db.child_off + tile + 1);
      for (uint32_t base = a; base < b; base += 32) {
        uint32_t k = base + lane;
        bool take = false;
        uint32_t child = 0;
        if (k < b) {
          uint32_t diff = __ldg(db.child_diff + k);
          int nd = HLM_DIFF_N(diff);
          int lbd = lbr - nd < 0 ? 0 : lbr - nd;
          if (lbd > lo && lbd <= hi) {
            for (int x = 0; x < nd; x++) {
              int col = HLM_DIFF_COL(diff, x);
              if (col >= off - HLM_BAND && col < off + n + HLM_BAND) take = true;
            }
            if (take) child = __ldg(db.child_ent + k);
          }
        }
        app.push(take, HLM_CAND(unit, child, HLM_CAND_OFFM(cc), HLM_CAND_STRAND(cc), 0xFF), lane);
      }
    }
  }
  app.finish(lane, &ck.counters[11]);
}

// ------------------------------------------------------------------ k_select
// round 0: candidates whose lower bound equals the task minimum.
// round 1: candidates with emin < lb <= cost of the best alignment found in round 0.
// After round 1 no unevaluated candidate can tie or beat the best (cost >= lb).
__global__ void __launch_bounds__(256)
k_select(HlmDevDB db, HlmKParams kp, HlmChunk ck, int n_targets, int round) {
  __shared__ int s_wcnt[8];
  __shared__ unsigned long long s_base;
  int64_t ncand = (int64_t)ck.counters[0];
  if (ncand > ck.cand_cap) ncand = ck.cand_cap;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t nloop = (ncand + stride - 1) / stride;
  unsigned long long n_take = 0;
  // the kernel is a chain of dependent loads per candidate (record -> locus -> task minimum ->
  // best): two candidates per thread are in flight at a time (four measured: slower)
  constexpr int UN = 2;
  for (int64_t it = 0; it < nloop; it += UN) {
    int64_t idx[UN];
    uint64_t c[UN];
    int lb[UN], g[UN];
    bool live[UN], take[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      idx[u] = (it + u) * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      c[u] = idx[u] < ncand ? ck.cand[idx[u]] : HLM_CAND_INVALID;
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      lb[u] = (int)HLM_CAND_LB(c[u]);
      live[u] = lb[u] <= kp.mm_max;  // padding and dead candidates carry lb >= 254
      g[u] = live[u] ? tile_locus(db, HLM_CAND_TILE(c[u])) : 0;
      if (live[u] && kp.score_cap) live[u] = task_live(kp, ck, HLM_CAND_UNIT(c[u]), g[u], n_targets);
    }
    uint32_t emin[UN];
    unsigned long long best[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      int64_t task = (int64_t)HLM_CAND_UNIT(c[u]) * n_targets + g[u];
      emin[u] = live[u] ? ck.task_emin[task] : 0u;
      best[u] = (live[u] && (round != 0 || lb[u] == 0)) ? ck.task_best[task] : HLM_BEST_NONE;
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      take[u] = false;
      if (live[u]) {
        int64_t task = (int64_t)HLM_CAND_UNIT(c[u]) * n_targets + g[u];
        if (round == 2) {
          // per-allele scoring: every live candidate is aligned; its result stays with the candidate
          take[u] = lb[u] > 0;
          if (lb[u] == 0)
            ck.cand_res[idx[u]] = HLM_BEST_PACK(0, (int)ck.unit_len[HLM_CAND_UNIT(c[u])], 0, 0);
        } else if (round == 0) {
          take[u] = (lb[u] == (int)emin[u]);
          if (take[u] && lb[u] == 0) {
            // a zero bound is an exact match of the whole read on a band diagonal: the DP result
            // is known (S = n, no edits, no clips) and nothing can beat or tie-break it
            take[u] = false;
            unsigned long long exact = HLM_BEST_PACK(0, (int)ck.unit_len[HLM_CAND_UNIT(c[u])], 0, 0);
            if (best[u] > exact) atomicMin(&ck.task_best[task], exact);  // most tasks hold it already
          }
        } else {
          take[u] = (lb[u] > (int)emin[u]) && best[u] != HLM_BEST_NONE && lb[u] <= HLM_BEST_COST(best[u]);
        }
      }
    }
    // the DP list stays dense (no padding).  List slots are claimed once per BLOCK per iteration
    // (warp ballots -> shared-memory prefix -> one global atomic): a per-warp atomic on the one
    // list counter was what bounded this kernel.  The number of DP rows (= trimmed length) goes
    // with every entry: after k_dp_dedup the list is counting-sorted by length so that the 32
    // candidates of a DP warp are alike.
    uint32_t bal[UN];
    int wcnt = 0;
#pragma unroll
    for (int u = 0; u < UN; u++) {
      bal[u] = __ballot_sync(FULL, take[u]);
      wcnt += __popc(bal[u]);
    }
    if (lane == 0) s_wcnt[wib] = wcnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; w++) {
        int cw = s_wcnt[w];
        s_wcnt[w] = tot;
        tot += cw;
      }
      s_base = tot ? atomicAdd(&ck.counters[1], (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    unsigned long long pos = s_base + (unsigned long long)s_wcnt[wib];
    __syncthreads();  // s_wcnt / s_base are rewritten by the next iteration
#pragma unroll
    for (int u = 0; u < UN; u++) {
      if (take[u]) {
        int nrows = (int)ck.unit_len[HLM_CAND_UNIT(c[u])];
        unsigned long long at = pos + __popc(bal[u] & ((1u << lane) - 1u));
        if ((int64_t)at < ck.dp_cap) {
          ck.dp_list[at] = (uint32_t)idx[u];
          ck.dp_len[at] = (uint8_t)nrows;
        } else {
          ck.counters[2] = 1;
        }
      }
      pos += __popc(bal[u]);
    }
    n_take += wcnt;
  }
  (void)n_take;  // the DP count (counters[6]) and the length histogram come from k_dp_dedup
}

// ------------------------------------------------------------------ DP list: footprint de-duplication
// Two selected candidates of the same unit, strand and locus whose tiles agree on the DP footprint
// (columns off - BAND .. off + n + BAND) are the same alignment problem: the windows differ only
// in columns the band cannot reach (typically two children of one rep that share the differing
// column inside the footprint).  The DP depends on nothing else, so one of them is enough - the
// result goes to the task minimum either way.  SURVEY.md section 8d counts algorithmic cells per
// distinct footprint, as the oracle's fast port evaluates them; without this pass the DP did 1.35x
// that.  Exact: a 32-bit tag of the footprint hash finds the earlier entry, the footprints are
// then compared word for word before anything is dropped.  Table: open addressing, u64 slots
// (tag << 32 | list index + 1), cleared per round; a full neighbourhood just keeps the entry.
struct Footprint {
  uint32_t w[21];  // 3 planes x 7 words: bits [0, n + 2 BAND) of the planes shifted to column off - BAND
};
__device__ __forceinline__ void footprint_of(const uint32_t* __restrict__ T, int off, int n, Footprint& f) {
  int sh = off - HLM_BAND, width = n + 2 * HLM_BAND;  // sh in [0, 32), width <= 224
#pragma unroll
  for (int p = 0; p < 3; p++) {
#pragma unroll
    for (int k = 0; k < 7; k++) {
      uint32_t lo = __ldg(T + 8 * p + k), hi = __ldg(T + 8 * p + k + 1);  // k + 1 <= 7
      uint32_t x = __funnelshift_r(lo, hi, sh);
      int bits = width - 32 * k;
      if (bits <= 0) x = 0;
      else if (bits < 32) x &= (1u << bits) - 1u;
      f.w[7 * p + k] = x;
    }
  }
}
#define DP_TAB_PROBES 32
__global__ void __launch_bounds__(256)
k_dp_dedup(HlmDevDB db, HlmChunk ck, int dedup) {
  __shared__ uint32_t s_hist[256];
  s_hist[threadIdx.x] = 0;  // blockDim.x == 256
  __syncthreads();
  int64_t nd = (int64_t)ck.counters[1];
  if (nd > ck.dp_cap) nd = ck.dp_cap;
  if (ck.counters[2]) nd = 0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long n_keep = 0;
  const uint64_t mask = ((uint64_t)1 << ck.dp_tab_bits) - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
    int n = ck.dp_len[i];
    bool keep = true;
    if (dedup) {
      uint64_t c = ck.cand[ck.dp_list[i]];
      uint32_t unit = HLM_CAND_UNIT(c);
      int off = (int)HLM_CAND_OFFM(c) + HLM_OFF_MIN, strand = (int)HLM_CAND_STRAND(c);
      Footprint f;
      footprint_of(db.tiles + (size_t)HLM_CAND_TILE(c) * HLM_TILE_WORDS, off, n, f);
      // (the same window content under two loci - close paralogs - is two tasks: the locus is part
      // of the key)
      int g = tile_locus(db, HLM_CAND_TILE(c));
      uint64_t h = (((uint64_t)unit * 2 + strand) * 64 + (uint64_t)g) * 0x9E3779B97F4A7C15ull + (uint64_t)n;
#pragma unroll
      for (int k = 0; k < 21; k++) h = (h ^ f.w[k]) * 0x100000001B3ull + (h >> 29);
      h ^= h >> 32;
      h *= 0xD6E8FEB86659FD93ull;
      h ^= h >> 32;
      uint64_t slot = h & mask;
      unsigned long long mine = (h & 0xFFFFFFFF00000000ull) | (unsigned long long)(i + 1);
      for (int probe = 0; probe < DP_TAB_PROBES; probe++) {
        unsigned long long old = atomicCAS(&ck.dp_tab[slot], 0ull, mine);
        if (old == 0ull) break;  // first of its footprint
        if ((old >> 32) == (mine >> 32)) {
          uint64_t c2 = ck.cand[ck.dp_list[(old & 0xFFFFFFFFull) - 1]];
          if (HLM_CAND_UNIT(c2) == unit && (int)HLM_CAND_STRAND(c2) == strand &&
              tile_locus(db, HLM_CAND_TILE(c2)) == g) {
            Footprint g;
            footprint_of(db.tiles + (size_t)HLM_CAND_TILE(c2) * HLM_TILE_WORDS, (int)HLM_CAND_OFFM(c2) + HLM_OFF_MIN,
                         n, g);
            bool same = true;
#pragma unroll
            for (int k = 0; k < 21; k++) same = same && (f.w[k] == g.w[k]);
            if (same) {
              keep = false;
              break;
            }
          }
        }
        slot = (slot + 1) & mask;
      }
    }
    if (keep) {
      atomicAdd(&s_hist[n], 1u);
      n_keep++;
    } else {
      ck.dp_len[i] = 0;  // dropped: k_dp_scatter skips it
    }
  }
  __syncthreads();
  if (s_hist[threadIdx.x]) atomicAdd(&ck.dp_hist[threadIdx.x], s_hist[threadIdx.x]);
  for (int o = 16; o; o >>= 1) n_keep += __shfl_xor_sync(FULL, n_keep, o);
  if ((threadIdx.x & 31) == 0 && n_keep) {
    atomicAdd(&ck.counters[6], n_keep);
    atomicAdd(&ck.counters[13], n_keep);
  }
}

// ------------------------------------------------------------------ DP list: counting sort by length
// dp_hist[n] -> dp_cursor[n] = first slot of length n in descending order of n (long DPs first:
// the tail of the launch is filled by the short ones); the histogram is cleared for the next round
__global__ void k_dp_scan(HlmChunk ck) {
  __shared__ uint32_t s[256];
  int t = threadIdx.x;  // 256 threads; thread t owns length 255 - t
  uint32_t v = ck.dp_hist[255 - t];
  s[t] = v;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    uint32_t x = t >= o ? s[t - o] : 0u;
    __syncthreads();
    s[t] += x;
    __syncthreads();
  }
  ck.dp_cursor[255 - t] = s[t] - v;
  ck.dp_hist[255 - t] = 0;
}

__global__ void __launch_bounds__(256)
k_dp_scatter(HlmChunk ck) {
  if (ck.counters[2]) return;  // a list overflowed: the host re-runs the chunk in halves
  int64_t nd = (int64_t)ck.counters[1];
  if (nd > ck.dp_cap) nd = ck.dp_cap;
  int lane = threadIdx.x & 31;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t nloop = (nd + stride - 1) / stride;
  for (int64_t it = 0; it < nloop; it++) {
    int64_t i = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool act = i < nd && ck.dp_len[i] != 0;  // 0: dropped by k_dp_dedup
    int n = act ? (int)ck.dp_len[i] : -1 - lane;  // inactive lanes never match anyone
    uint32_t peers = __match_any_sync(FULL, n);
    if (act) {
      int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&ck.dp_cursor[n], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      ck.dp_sorted[base + __popc(peers & ((1u << lane) - 1u))] = ck.dp_list[i];
    }
  }
}

// ------------------------------------------------------------------ k_dp
__global__ void __launch_bounds__(128)
k_dp(HlmDevDB db, HlmKParams kp, HlmChunk ck, int n_targets) {
  if (ck.counters[2]) return;  // a list overflowed: the host re-runs the chunk in halves
  int64_t nd = (int64_t)ck.counters[13];  // entries of the sorted list (after de-duplication)
  if (nd > ck.dp_cap) nd = ck.dp_cap;
  int lane = threadIdx.x & 31;
  unsigned long long cells = 0;
  // warps pull groups of 32 entries of the length-sorted list from a device counter
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&ck.counters[14], 32ull);
    base = __shfl_sync(FULL, base, 0);
    if ((int64_t)base >= nd) break;
    int64_t i = (int64_t)base + lane;
    if (i >= nd) continue;
    uint32_t di = ck.dp_sorted[i];
    uint64_t c = ck.cand[di];
    uint32_t unit = HLM_CAND_UNIT(c), tile = HLM_CAND_TILE(c);
    int off = (int)HLM_CAND_OFFM(c) + HLM_OFF_MIN, strand = (int)HLM_CAND_STRAND(c);
    int n = ck.unit_len[unit];
    const uint32_t* rd = ck.unit_planes + (size_t)unit * HLM_UNIT_STRIDE + 18 * strand;
    const uint32_t* T = db.tiles + (size_t)tile * HLM_TILE_WORDS;
    HlmAln a = hlm_dp_align(rd, n, T, off, kp.one, kp.clip_mode == HLM_CLIP_NONE);
    int g = tile_locus(db, tile);
    atomicMin(&ck.task_best[(int64_t)unit * n_targets + g],
              (unsigned long long)HLM_BEST_PACK(a.cost, a.S, a.lead, a.trail));
    if (ck.cand_res) ck.cand_res[di] = (unsigned long long)HLM_BEST_PACK(a.cost, a.S, a.lead, a.trail);
    cells += (unsigned long long)n * HLM_NB;
  }
  for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(FULL, cells, o);
  if (lane == 0 && cells) atomicAdd(&ck.counters[4], cells);
}

__global__ void k_dp_round_end(HlmChunk ck) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    ck.counters[1] = 0;
    ck.counters[13] = 0;
    ck.counters[14] = 0;
  }
}

// ------------------------------------------------------------------ per-allele scoring
// hlm_score_alleles (what `bwa mem -a` against every allele of a gene
// feeds functions.cpp:347-520).  Every candidate with a result hands it to the alleles that hold
// its tile (tile_al CSR).  A rep also speaks for those of its children that were never
// materialised: the ones without a differing column inside the DP footprint align exactly like it.
// Warp per candidate, lanes over the allele lists; one atomicMin per (unit, allele) incidence.
__global__ void __launch_bounds__(256)
k_allele_scatter(HlmDevDB db, HlmKParams kp, HlmChunk ck) {
  int64_t ncand = (int64_t)ck.counters[0];
  if (ncand > ck.cand_cap) ncand = ck.cand_cap;
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < ncand; i += nwarps) {
    unsigned long long r = ck.cand_res[i];
    if (r == HLM_BEST_NONE || HLM_BEST_COST(r) > kp.mm_max) continue;
    uint64_t c = ck.cand[i];
    uint32_t unit = HLM_CAND_UNIT(c), tile = HLM_CAND_TILE(c);
    int off = (int)HLM_CAND_OFFM(c) + HLM_OFF_MIN;
    int n = ck.unit_len[unit];
    unsigned long long* row = ck.allele_best + (size_t)unit * ck.n_al;
    uint32_t a = __ldg(db.tile_al_off + tile), b = __ldg(db.tile_al_off + tile + 1);
    for (uint32_t k = a + lane; k < b; k += 32) atomicMin(&row[__ldg(db.tile_al_ent + k)], r);
    uint32_t ca = __ldg(db.child_off + tile), cb = __ldg(db.child_off + tile + 1);
    for (uint32_t k = ca; k < cb; k++) {  // (children lists are short next to the allele lists)
      uint32_t diff = __ldg(db.child_diff + k);
      bool inside = false;
      for (int x = 0; x < HLM_DIFF_N(diff); x++) {
        int col = HLM_DIFF_COL(diff, x);
        if (col >= off - HLM_BAND && col < off + n + HLM_BAND) inside = true;
      }
      if (inside) continue;  // materialised on its own (seed index or expansion) when it is a candidate
      uint32_t child = __ldg(db.child_ent + k);
      uint32_t xa = __ldg(db.tile_al_off + child), xb = __ldg(db.tile_al_off + child + 1);
      for (uint32_t q = xa + lane; q < xb; q += 32) atomicMin(&row[__ldg(db.tile_al_ent + q)], r);
    }
  }
}

// allele_best -> nm / clip bytes per (pair, allele); unit of mate m of a pair: pair * U + (rna ? 3 m : m)
__global__ void __launch_bounds__(256)
k_allele_out(HlmKParams kp, HlmChunk ck, uint8_t* nm1, uint8_t* clip1, uint8_t* nm2, uint8_t* clip2) {
  int64_t total = ck.n_pairs * (int64_t)ck.n_al;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int U = ck.units_per_pair;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t pair = i / ck.n_al;
    int a = (int)(i - pair * ck.n_al);
    for (int m = 0; m < (kp.paired ? 2 : 1); m++) {
      int64_t unit = pair * U + (kp.rna ? 3 * m : m);
      unsigned long long r = ck.allele_best[(size_t)unit * ck.n_al + a];
      uint8_t nm = 255, clip = 0;
      if (r != HLM_BEST_NONE && HLM_BEST_COST(r) <= kp.mm_max) {
        clip = (uint8_t)(HLM_BEST_LEAD(r) + HLM_BEST_TRAIL(r));
        nm = (uint8_t)(HLM_BEST_COST(r) - (int)clip);
      }
      (m == 0 ? nm1 : nm2)[i] = nm;
      (m == 0 ? clip1 : clip2)[i] = clip;
    }
  }
}

// ------------------------------------------------------------------ k_scores
// lane g of the warp handles target g of one pair.  dna: read_sam_dna (map_dna.cpp:44-118):
// score = NM + leading clip (+ trailing clip in CLIP_BOTH) + 1, 0 when nothing mapped.
// rna: read_sam_rna (map_rna.cpp:39-101): sum over blocks and mates of NM + clips for mapped
// records (clips only while mm != len(SEQ)) and len(SEQ) for unmapped ones.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_scores(HlmKParams kp, HlmChunk ck, int n_targets) {
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  int U = ck.units_per_pair;
  for (int64_t pair = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib; pair < ck.n_pairs; pair += nwarps) {
    if (lane >= n_targets) continue;
    int g = lane;
    int64_t o = pair * n_targets + g;
    bool kept = ck.flags[pair] & HLM_F_KEPT;
    if (!kp.rna) {
      for (int m = 0; m < 2; m++) {
        uint16_t s = 0;
        int16_t aln = 0;
        uint8_t nm = 0, lead = 0, trail = 0;
        if (kept && (m == 0 || kp.paired)) {
          int64_t unit = pair * U + m;
          if (ck.unit_len[unit] && ((ck.unit_mask[unit] >> g) & 1u)) {
            unsigned long long b = ck.task_best[unit * n_targets + g];
            if (b != HLM_BEST_NONE && HLM_BEST_COST(b) <= kp.mm_max) {
              int cost = HLM_BEST_COST(b), ld = HLM_BEST_LEAD(b), tr = HLM_BEST_TRAIL(b);
              int nmv = cost - ld - tr;
              s = (uint16_t)(nmv + ld + (kp.clip_mode == HLM_CLIP_BOTH ? tr : 0) + 1);
              aln = (int16_t)HLM_BEST_S(b);
              nm = (uint8_t)nmv;
              lead = (uint8_t)ld;
              trail = (uint8_t)tr;
            }
          }
        }
        if (m == 0) {
          ck.s1[o] = s; ck.aln1[o] = aln; ck.nm1[o] = nm; ck.lead1[o] = lead; ck.trail1[o] = trail;
        } else {
          ck.s2[o] = s; ck.aln2[o] = aln; ck.nm2[o] = nm; ck.lead2[o] = lead; ck.trail2[o] = trail;
        }
      }
    } else {
      int sum = 0;
      bool any = false;
      // one record of read_sam_rna: a mapped one adds NM + clips, an unmapped one its length
      auto record = [&](int64_t unit, int len) {
        any = true;
        unsigned long long b = ck.task_best[unit * n_targets + g];
        if (len > 0 && b != HLM_BEST_NONE && HLM_BEST_COST(b) <= kp.mm_max) {
          int ld = HLM_BEST_LEAD(b), tr = HLM_BEST_TRAIL(b);
          int mm = HLM_BEST_COST(b) - ld - tr;
          if (ld > 0 && mm != len) mm += ld;
          if (kp.clip_mode == HLM_CLIP_BOTH && tr > 0 && mm != len) mm += tr;
          sum += mm;
        } else {
          sum += len;
        }
      };
      if (kept) {
        for (int u = 0; u < U; u++) {
          int64_t unit = pair * U + u;
          int len = ck.unit_len[unit];
          if (!len || !((ck.unit_mask[unit] >> g) & 1u)) continue;
          record(unit, len);
        }
        if (ck.blk_off && g < kp.n_loci) {
          int nm_ = kp.paired ? 2 : 1;
          for (int64_t b = (int64_t)ck.blk_off[pair]; b < (int64_t)ck.blk_off[pair + 1]; b++) {
            hlm_rna_block blk = ck.blk[b];
            if (blk.gene != g || blk.mate >= nm_) continue;
            int64_t unit = ck.n_pairs * U + (b - ck.blk_first);
            record(unit, ck.unit_len[unit]);  // an empty block is still a record (adds 0)
          }
        }
      }
      ck.s1[o] = any ? (uint16_t)sum : (uint16_t)0xFFFF;
      ck.s2[o] = 0; ck.aln1[o] = 0; ck.aln2[o] = 0; ck.nm1[o] = 0; ck.nm2[o] = 0;
      ck.lead1[o] = 0; ck.lead2[o] = 0; ck.trail1[o] = 0; ck.trail2[o] = 0;
    }
  }
}

// ------------------------------------------------------------------ k_assign
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_assign(HlmKParams kp, HlmChunk ck, int n_targets) {
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t nwarps = (int64_t)gridDim.x * WARPS_PER_BLOCK;
  int G = n_targets - 1;
  for (int64_t pair = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib; pair < ck.n_pairs; pair += nwarps) {
    bool kept = ck.flags[pair] & HLM_F_KEPT;
    int g = lane;
    bool act = g < n_targets;
    int64_t o = pair * n_targets + (act ? g : 0);
    uint32_t tmask = 0, vmask = 0;
    int minsum = 0, o1 = 0, o2 = 0;
    if (kept) {
      int len1 = ck.trim_len1[pair];
      int len2 = kp.paired ? ck.trim_len2[pair] : 0;
      if (!kp.paired && ck.size_override1) len1 = ck.size_override1[pair];  // preselect.cpp:1199
      if (!kp.rna) {
        int s1 = act ? ck.s1[o] : 0, s2 = (act && kp.paired) ? ck.s2[o] : 0;
        // others back-fill (map_dna.cpp:903-915)
        o1 = __shfl_sync(FULL, s1, G);
        o2 = __shfl_sync(FULL, s2, G);
        if (kp.paired) {
          if (o1 != 0 && o2 == 0) o2 = o1;
          else if (o2 != 0 && o1 == 0) o1 = o2;
          if (g == G) { s1 = o1; s2 = o2; }
        }
        int nm_sum = 10000;
        bool valid = false;
        if (act) {
          if (kp.paired) {
            if (s1 != 0 && s2 != 0 && s1 != 10000 && s2 != 10000) {
              int v1 = (int)(len1 * kp.tolerance), v2 = (int)(len2 * kp.tolerance);
              if ((s1 - 1) <= v1 && (s2 - 1) <= v2) { valid = true; nm_sum = s1 + s2; }
            }
          } else if (s1 != 0 && s1 != 10000) {
            int v1 = (int)(len1 * kp.tolerance);
            if ((s1 - 1) <= v1) { valid = true; nm_sum = s1; }
          }
        }
        vmask = __ballot_sync(FULL, valid);
        int mn = __reduce_min_sync(FULL, valid ? nm_sum : 100000);
        if (vmask) {
          // Target = every g whose nm_sum equals the minimum (invalid ones hold 10000)
          tmask = __ballot_sync(FULL, act && nm_sum == mn);
          minsum = mn;
        }
      } else {
        int v = act ? ck.s1[o] : 0xFFFF;
        int nm_sum = (v == 0xFFFF) ? 1000 : v;
        bool vis = act && nm_sum != 1000;
        vmask = __ballot_sync(FULL, vis);
        int mn = __reduce_min_sync(FULL, act ? nm_sum : 1000);
        if (mn > 1000) mn = 1000;
        if (vmask) {
          bool t = false;
          if (act && nm_sum == mn) {
            if (g != G) {
              float mx = (float)((len1 * 2) * kp.tolerance);  // map_rna.cpp:975
              t = ((float)mn <= mx);
            } else {
              t = mn < 30;  // min_score_others, map_rna.cpp:939
            }
          }
          tmask = __ballot_sync(FULL, t);
          minsum = mn;
        }
      }
    }
    if (lane == 0) {
      ck.target_mask[pair] = tmask;
      ck.visible_mask[pair] = vmask;
      ck.min_sum[pair] = (uint16_t)minsum;
      ck.others_s1[pair] = (uint16_t)o1;
      ck.others_s2[pair] = (uint16_t)o2;
    }
  }
}

// ------------------------------------------------------------------ DP layout A/B (evidence)
// The same banded affine DP in the two work layouts, on flat arrays of (read planes, tile, off):
//   k_dpab_thread  thread per candidate, band row in registers (what k_dp does)
//   k_dpab_warp    warp per candidate, band cells across lanes (lane k owns cells k and k + 32),
//                  neighbours fetched with shuffles, the horizontal-gap recurrence as a warp
//                  max-plus prefix scan - the "wavefront with warp shuffles" layout
// Results are bit-identical (tests/test_gpu_parity.py); bench.py reports both cell rates.
__global__ void __launch_bounds__(128)
k_dpab_thread(const uint32_t* __restrict__ reads, const uint32_t* __restrict__ tiles,
              const int* __restrict__ offn, int m, int one, int4* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    HlmAln a = hlm_dp_align(reads + (size_t)i * 18, offn[2 * i + 1], tiles + (size_t)i * HLM_TILE_WORDS,
                            offn[2 * i], one);
    out[i] = make_int4(a.S, a.cost, a.lead, a.trail);
  }
}

__global__ void __launch_bounds__(256)
k_dpab_warp(const uint32_t* __restrict__ reads, const uint32_t* __restrict__ tiles,
            const int* __restrict__ offn, int m, int4* __restrict__ out) {
  int lane = threadIdx.x & 31;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int GE = -HLM_V_GAPE, GO = -HLM_V_GAPO;
  bool has1 = lane < HLM_NB - 32;  // lanes 0..8 also own cell lane + 32
  for (int c = warp; c < m; c += nwarps) {
    const uint32_t* rd = reads + (size_t)c * 18;
    const uint32_t* T = tiles + (size_t)c * HLM_TILE_WORDS;
    int off = offn[2 * c], n = offn[2 * c + 1];
    int H0 = 0, E0 = HLM_V_NEG, H1 = 0, E1 = HLM_V_NEG;
    long long best = -(1ll << 60);  // (value << 8 | row) of this lane's cells
    int start = 0, trailpen = HLM_V_CLIP + n * HLM_CU;
    for (int i = 1; i <= n; i++) {
      int pos = off - HLM_BAND + (i - 1), w = pos >> 5, sh = pos & 31;
      uint32_t r0 = (rd[(i - 1) >> 5] >> ((i - 1) & 31)) & 1u, r1 = (rd[6 + ((i - 1) >> 5)] >> ((i - 1) & 31)) & 1u,
               rn = (rd[12 + ((i - 1) >> 5)] >> ((i - 1) & 31)) & 1u;
      uint32_t m0 = 0u - r0, m1 = 0u - r1, mn = 0u - rn;
      uint32_t eq_lo = ~(hlm_funnel(hlm_tw(T, w), hlm_tw(T, w + 1), sh) ^ m0) &
                       ~(hlm_funnel(hlm_tw(T + 8, w), hlm_tw(T + 8, w + 1), sh) ^ m1) &
                       ~(hlm_funnel(hlm_tw(T + 16, w), hlm_tw(T + 16, w + 1), sh) | mn);
      uint32_t eq_hi = ~(hlm_funnel(hlm_tw(T, w + 1), hlm_tw(T, w + 2), sh) ^ m0) &
                       ~(hlm_funnel(hlm_tw(T + 8, w + 1), hlm_tw(T + 8, w + 2), sh) ^ m1) &
                       ~(hlm_funnel(hlm_tw(T + 16, w + 1), hlm_tw(T + 16, w + 2), sh) | mn);
      start = (i == 1) ? -(HLM_V_CLIP + HLM_CU + 1) : start - (HLM_CU + 1);
      trailpen -= HLM_CU;
      // vertical predecessors: cell k + 1 of the previous row
      int Hu0 = __shfl_down_sync(FULL, H0, 1), Eu0 = __shfl_down_sync(FULL, E0, 1);
      int Hx = __shfl_sync(FULL, H1, 0), Ex = __shfl_sync(FULL, E1, 0);
      if (lane == 31) { Hu0 = Hx; Eu0 = Ex; }
      int Hu1 = __shfl_down_sync(FULL, H1, 1), Eu1 = __shfl_down_sync(FULL, E1, 1);
      int e0 = hlm_max(Eu0 + HLM_V_GAPE, Hu0 + HLM_V_GAPO);
      int e1 = (lane < HLM_NB - 33) ? hlm_max(Eu1 + HLM_V_GAPE, Hu1 + HLM_V_GAPO) : HLM_V_NEG;
      int hp0 = hlm_max3(H0 + (((eq_lo >> lane) & 1u) ? HLM_V_MATCH : HLM_V_MISM), e0, start);
      int hp1 = has1 ? hlm_max3(H1 + (((eq_hi >> lane) & 1u) ? HLM_V_MATCH : HLM_V_MISM), e1, start) : HLM_V_NEG;
      // horizontal gaps: f_k = max_{j<k} (hp_j + j*GE) - GO - (k-1)*GE   (max-plus prefix scan)
      int u0 = hp0 + lane * GE, s0 = u0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(FULL, s0, o);
        if (lane >= o) s0 = hlm_max(s0, v);
      }
      int ex0 = __shfl_up_sync(FULL, s0, 1);  // exclusive
      int tot0 = __shfl_sync(FULL, s0, 31);
      int f0 = lane == 0 ? HLM_V_NEG : ex0 - GO - (lane - 1) * GE;
      int u1 = has1 ? hp1 + (lane + 32) * GE : HLM_V_NEG, s1 = u1;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        int v = __shfl_up_sync(FULL, s1, o);
        if (lane >= o) s1 = hlm_max(s1, v);
      }
      int ex1 = __shfl_up_sync(FULL, s1, 1);
      ex1 = lane == 0 ? tot0 : hlm_max(ex1, tot0);
      int f1 = ex1 - GO - (lane + 31) * GE;
      int h0 = hlm_max(hp0, f0), h1 = has1 ? hlm_max(hp1, f1) : HLM_V_NEG;
      H0 = h0; E0 = e0; H1 = h1; E1 = e1;
      int fin = hlm_max(h0, h1) - (i < n ? trailpen : 0);
      long long key = ((long long)fin << 8) | (long long)i;
      if (key > best) best = key;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      long long v = __shfl_xor_sync(FULL, best, o);
      if (v > best) best = v;
    }
    if (lane == 0) {
      int bv = (int)(best >> 8), bi = (int)(best & 0xFF);
      int S = (bv + HLM_SC - 1) >> 20, rem = S * HLM_SC - bv;
      out[c] = make_int4(S, rem >> 9, rem & 511, n - bi);
    }
  }
}

void hlm_launch_dpab(int which, const uint32_t* reads, const uint32_t* tiles, const int* offn, int m,
                     int4* out, int sms, cudaStream_t st) {
  if (which == 0) k_dpab_thread<<<sms * 4, 128, 0, st>>>(reads, tiles, offn, m, 1, out);
  else k_dpab_warp<<<sms * 8, 256, 0, st>>>(reads, tiles, offn, m, out);
}

// ------------------------------------------------------------------ INT32 peak micro-kernels
template <int WHICH>
__global__ void __launch_bounds__(256) k_int_peak(int* out, int iters, int seed) {
  int a = threadIdx.x + seed, b = blockIdx.x ^ seed, c = a * 3 + 1, d = b * 5 + 2;
  int e = a + 7, f = b + 11, g = c + 13, h = d + 17;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (WHICH == 0) {  // IADD3 chains (alu pipe)
        a = a + b + seed; c = c + d + seed; e = e + f + seed; g = g + h + seed;
        b = b + a + seed; d = d + c + seed; f = f + e + seed; h = h + g + seed;
      } else if (WHICH == 1) {  // IMAD chains (fma pipe)
        a = a * seed + b; c = c * seed + d; e = e * seed + f; g = g * seed + h;
        b = b * seed + a; d = d * seed + c; f = f * seed + e; h = h * seed + g;
      } else if (WHICH == 2) {  // mixed add + mad
        a = a + b + seed; c = c * seed + d; e = e + f + seed; g = g * seed + h;
        b = b + a + seed; d = d * seed + c; f = f + e + seed; h = h * seed + g;
      } else if (WHICH == 3) {  // VIADDMNMX chains (what the DP issues)
        a = __viaddmax_s32(a, seed, b); c = __viaddmax_s32(c, seed, d);
        e = __viaddmax_s32(e, seed, f); g = __viaddmax_s32(g, seed, h);
        b = __viaddmax_s32(b, seed, a); d = __viaddmax_s32(d, seed, c);
        f = __viaddmax_s32(f, seed, e); h = __viaddmax_s32(h, seed, g);
      } else if (WHICH == 4) {  // LOP3 chains (three-input logic: 119 of the 246 k_myers loop instructions)
#define HLM_LOP3(x, y) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(seed));
        HLM_LOP3(a, b) HLM_LOP3(c, d) HLM_LOP3(e, f) HLM_LOP3(g, h)
        HLM_LOP3(b, a) HLM_LOP3(d, c) HLM_LOP3(f, e) HLM_LOP3(h, g)
      } else if (WHICH == 5) {  // SHF chains (funnel shifts: 42 of the 246)
#define HLM_SHF(x, y) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(seed));
        HLM_SHF(a, b) HLM_SHF(c, d) HLM_SHF(e, f) HLM_SHF(g, h)
        HLM_SHF(b, a) HLM_SHF(d, c) HLM_SHF(f, e) HLM_SHF(h, g)
      } else {  // the k_myers row mix: 3 LOP3 : 1 SHF
        HLM_LOP3(a, b) HLM_LOP3(c, d) HLM_LOP3(e, f) HLM_SHF(g, h)
        HLM_LOP3(b, a) HLM_LOP3(d, c) HLM_LOP3(f, e) HLM_SHF(h, g)
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

double hlm_int_peak_run(int which, int sms, cudaStream_t st) {
  int blocks = sms * 8, threads = 256, iters = 4096;
  int* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(int)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, st);
    switch (which) {
      case 0: k_int_peak<0><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
      case 1: k_int_peak<1><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
      case 2: k_int_peak<2><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
      case 3: k_int_peak<3><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
      case 4: k_int_peak<4><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
      case 5: k_int_peak<5><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
      default: k_int_peak<6><<<blocks, threads, 0, st>>>(out, iters, 3 + rep); break;
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0;
    double g = ops / (ms * 1e-3) / 1e9;
    if (rep > 0 && g > best) best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

// ------------------------------------------------------------------ launchers
static inline int grid_for(int sms, int per_sm) { return sms * per_sm; }

// blocks per SM of the persistent grids whose size decides what can run beside them on another
// pipeline slot's stream (all of them take work from device counters, so any size is correct).
// HLM_TUNE="myers=3,dp=3,seed=10,expand=8,select=8" overrides them when a context is created:
// a measurement knob (tools/grid_sweep.py), not part of the interface.
enum { TUNE_MYERS, TUNE_DP, TUNE_SEED, TUNE_EXPAND, TUNE_SELECT, TUNE_N };
static int g_tune[TUNE_N] = {4, 4, 10, 8, 8};
void hlm_tune_reload() {
  static const char* names[TUNE_N] = {"myers", "dp", "seed", "expand", "select"};
  static const int defaults[TUNE_N] = {4, 4, 10, 8, 8};
  for (int i = 0; i < TUNE_N; i++) g_tune[i] = defaults[i];
  const char* e = getenv("HLM_TUNE");
  if (!e) return;
  for (int i = 0; i < TUNE_N; i++) {
    const char* q = strstr(e, names[i]);
    if (!q || q[strlen(names[i])] != '=') continue;
    int v = atoi(q + strlen(names[i]) + 1);
    if (v >= 1 && v <= 32) g_tune[i] = v;
  }
}

void hlm_launch_screen(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int sms,
                       cudaStream_t st) {
  if (kp.screen_variant == HLM_SCREEN_FASTQ && !kp.skip_screen && kp.minhit >= 1 && db.n_odd == 0)
    k_prescreen<<<grid_for(sms, 8), 128, 0, st>>>(db, kp, ck);
  k_screen<<<grid_for(sms, 8), WARPS_PER_BLOCK * 32, 0, st>>>(db, kp, ck);
}
void hlm_launch_pack(const HlmKParams& kp, const HlmChunk& ck, int sms, cudaStream_t st) {
  k_pack_units<<<grid_for(sms, 8), WARPS_PER_BLOCK * 32, 0, st>>>(kp, ck);
}
void hlm_launch_clear_tasks(const HlmChunk& ck, int n_targets, int sms, cudaStream_t st) {
  int64_t n_tasks = ck.n_units * n_targets;
  k_clear_tasks<<<grid_for(sms, 8), 256, 0, st>>>(ck, n_tasks);
}
void hlm_launch_seed(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int sms,
                     cudaStream_t st) {
  k_seed<<<grid_for(sms, g_tune[TUNE_SEED]), SEED_WARPS * 32, 0, st>>>(db, kp, ck);
}
void hlm_launch_myers(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                      int phase, int sms, cudaStream_t st) {
  k_myers<<<grid_for(sms, g_tune[TUNE_MYERS]), 256, 0, st>>>(db, kp, ck, n_targets, phase);
}
// per-device kernel attributes (call once per context, after cudaSetDevice)
void hlm_kernels_init() {
  hlm_tune_reload();
  // k_myers: 5 blocks x 42 KB of PEQ rows per SM; k_seed: 10 blocks x 18 KB of hash sets (register-bound)
  cudaFuncSetAttribute(k_myers, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(k_seed, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}
void hlm_launch_snapshot(const HlmChunk& ck, int which, cudaStream_t st) {
  k_snapshot<<<1, 32, 0, st>>>(ck, which);
}
void hlm_launch_expand(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                       int round, int sms, cudaStream_t st) {
  k_expand<<<grid_for(sms, g_tune[TUNE_EXPAND]), 256, 0, st>>>(db, kp, ck, n_targets, round);
  k_snapshot<<<1, 32, 0, st>>>(ck, round == 2 ? 1 : 1 + round);  // round 2 (per-allele): one expansion
}
void hlm_launch_select(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                       int round, int sms, cudaStream_t st) {
  k_select<<<grid_for(sms, g_tune[TUNE_SELECT]), 256, 0, st>>>(db, kp, ck, n_targets, round);
}
void hlm_launch_dp(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, int n_targets,
                   int sms, cudaStream_t st) {
  // per-allele scoring keeps one result per candidate: no de-duplication there
  int dedup = (ck.cand_res == nullptr && ck.dp_tab != nullptr) ? 1 : 0;
  if (dedup) cudaMemsetAsync(ck.dp_tab, 0, (size_t)8 << ck.dp_tab_bits, st);
  k_dp_dedup<<<grid_for(sms, 8), 256, 0, st>>>(db, ck, dedup);
  k_dp_scan<<<1, 256, 0, st>>>(ck);
  k_dp_scatter<<<grid_for(sms, 4), 256, 0, st>>>(ck);
  k_dp<<<grid_for(sms, g_tune[TUNE_DP]), 128, 0, st>>>(db, kp, ck, n_targets);
  k_dp_round_end<<<1, 32, 0, st>>>(ck);
}
void hlm_launch_allele_gather(const HlmDevDB& db, const HlmKParams& kp, const HlmChunk& ck, uint8_t* nm1,
                              uint8_t* clip1, uint8_t* nm2, uint8_t* clip2, int sms, cudaStream_t st) {
  k_allele_scatter<<<grid_for(sms, 8), 256, 0, st>>>(db, kp, ck);
  k_allele_out<<<grid_for(sms, 8), 256, 0, st>>>(kp, ck, nm1, clip1, nm2, clip2);
}
void hlm_launch_scores(const HlmKParams& kp, const HlmChunk& ck, int n_targets, int sms,
                       cudaStream_t st) {
  k_scores<<<grid_for(sms, 8), WARPS_PER_BLOCK * 32, 0, st>>>(kp, ck, n_targets);
}
void hlm_launch_assign(const HlmKParams& kp, const HlmChunk& ck, int n_targets, int sms,
                       cudaStream_t st) {
  k_assign<<<grid_for(sms, 8), WARPS_PER_BLOCK * 32, 0, st>>>(kp, ck, n_targets);
}
