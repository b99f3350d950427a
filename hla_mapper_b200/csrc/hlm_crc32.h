// hlm_crc32.h — CRC-32 (the gzip polynomial) with carry-less multiplication where the CPU has it.
// zlib's crc32() runs at 1.5 - 2 GB/s per thread, which made the CRC check a quarter of the CPU work
// of the multi-threaded gzip reader (hlm_pgzip.h); folding 64 bytes per step with PCLMULQDQ
// (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel
// 2009) is several times faster.  The fast path is used only after it has reproduced zlib's
// result on a test vector in this process; otherwise - and for short or unaligned tails -
// zlib's crc32() does the work.  hlm_crc32(crc, p, n) has crc32()'s meaning.
#ifndef HLM_CRC32_H
#define HLM_CRC32_H

#include <zlib.h>

#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) && (defined(__GNUC__) || defined(__clang__))
#include <immintrin.h>
#define HLM_CRC32_CLMUL 1

// `len` a multiple of 16, at least 64; crc is the raw register (complemented by the caller)
__attribute__((target("pclmul,sse4.1"))) static inline uint32_t hlm_crc32_clmul_raw(const uint8_t* buf, size_t len,
                                                                                  uint32_t crc) {
  static const uint64_t __attribute__((aligned(16))) k1k2[] = {0x0154442bd4ull, 0x01c6e41596ull};
  static const uint64_t __attribute__((aligned(16))) k3k4[] = {0x01751997d0ull, 0x00ccaa009eull};
  static const uint64_t __attribute__((aligned(16))) k5k0[] = {0x0163cd6124ull, 0x0000000000ull};
  static const uint64_t __attribute__((aligned(16))) poly[] = {0x01db710641ull, 0x01f7011641ull};
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = _mm_load_si128((const __m128i*)k1k2);
  buf += 64;
  len -= 64;
  while (len >= 64) {  // four 128-bit lanes folded over 64 bytes
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
    x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
    x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64;
    len -= 64;
  }
  x0 = _mm_load_si128((const __m128i*)k3k4);  // the four lanes into one
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16;
    len -= 16;
  }
  // 128 -> 64 bits
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_loadl_epi64((const __m128i*)k5k0);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  // Barrett reduction to 32 bits
  x0 = _mm_load_si128((const __m128i*)poly);
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}

// 1 when the CPU has PCLMULQDQ + SSE4.1 AND the routine above reproduces zlib on a test vector
static inline int hlm_crc32_clmul_ok() {
  static int state = -1;  // benign race: every thread computes the same answer
  if (state >= 0) return state;
  int ok = 0;
  __builtin_cpu_init();
  if (__builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1")) {
    uint8_t v[64 * 5 + 48];
    for (size_t i = 0; i < sizeof v; i++) v[i] = (uint8_t)(i * 131u + (i >> 3) * 17u + 7u);
    uint32_t seed = 0x1234abcdu;
    uint32_t want = (uint32_t)crc32(seed, v, (uInt)sizeof v);
    uint32_t got = ~hlm_crc32_clmul_raw(v, sizeof v, ~seed);
    ok = want == got;
  }
  state = ok;
  return ok;
}
#endif

static inline uint32_t hlm_crc32(uint32_t crc, const uint8_t* p, size_t n) {
#ifdef HLM_CRC32_CLMUL
  if (n >= 256 && hlm_crc32_clmul_ok()) {
    size_t body = n & ~(size_t)15;
    crc = ~hlm_crc32_clmul_raw(p, body, ~crc);
    p += body;
    n -= body;
  }
#endif
  while (n) {  // zlib takes uInt lengths
    size_t k = n > ((size_t)1 << 30) ? ((size_t)1 << 30) : n;
    crc = (uint32_t)crc32(crc, p, (uInt)k);
    p += k;
    n -= k;
  }
  return crc;
}

#endif
