// Fuzz harness of the parallel gzip reader (csrc/hlm_pgzip.h) against zlib, to be built with the
// sanitizers:  g++ -O1 -g -fsanitize=address,undefined -std=c++17 -o fuzz tools/pgzip_fuzz.cpp -lz -lpthread
//              ./fuzz file.gz <iterations> <seed>
// Every iteration truncates the file, flips bits, overwrites a stretch with noise or leaves it
// intact, picks a chunk size (1 - 24 KiB) and a thread count, and checks that the reader reports an
// error exactly when zlib does and the same text otherwise.
#include "../hla_mapper_b200/csrc/hlm_pgzip.h"
#include <cstdio>
#include <random>
int main(int argc, char** argv) {
  // argv[1]: gz file ; runs random corruptions
  FILE* f = fopen(argv[1], "rb"); std::vector<uint8_t> z; { uint8_t b[65536]; size_t r; while ((r = fread(b,1,sizeof b,f))>0) z.insert(z.end(), b, b+r);} fclose(f);
  int iters = atoi(argv[2]);
  std::mt19937_64 rng(atoi(argv[3]));
  int n_ok=0,n_err=0,n_mismatch=0;
  for (int it = 0; it < iters; it++) {
    std::vector<uint8_t> d = z;
    int kind = rng() % 4;
    if (kind == 0) d.resize(20 + rng() % (d.size() - 20));
    else if (kind == 1) { for (int k = 0; k < 1 + (int)(rng()%3); k++) d[10 + rng() % (d.size()-10)] ^= 1 << (rng()%8); }
    else if (kind == 2) { size_t a = 10 + rng() % (d.size()-10), l = rng() % 2000; for (size_t k = a; k < a + l && k < d.size(); k++) d[k] = (uint8_t)rng(); }
    else { /* intact */ }
    char kb[16]; snprintf(kb, sizeof kb, "%d", (int)(1 + rng() % 24)); setenv("HLM_PGZIP_CHUNK_KB", kb, 1);
    // zlib reference
    std::vector<uint8_t> want; bool zok = true;
    { z_stream zs; memset(&zs,0,sizeof zs); inflateInit2(&zs, 31); zs.next_in = d.data(); zs.avail_in = d.size(); std::vector<uint8_t> buf(1<<20);
      for(;;){ zs.next_out = buf.data(); zs.avail_out = buf.size(); int r = inflate(&zs, Z_NO_FLUSH); want.insert(want.end(), buf.data(), buf.data() + (buf.size()-zs.avail_out));
        if (r == Z_STREAM_END) { // further members?
          size_t rem = zs.avail_in; const uint8_t* nx = zs.next_in; while (rem && *nx==0) {nx++; rem--;}
          if (!rem) break; zs.next_in=(Bytef*)nx; zs.avail_in=rem; if (inflateReset(&zs)!=Z_OK){zok=false;break;} continue; }
        if (r != Z_OK) { zok = false; break; } if (zs.avail_in == 0 && zs.avail_out != 0) { zok = false; break; } }
      inflateEnd(&zs); }
    std::vector<uint8_t> got; bool pok = true;
    { HlmPgzip pg(d.data(), d.size(), 1 + rng()%6); std::vector<uint8_t> buf(777777);
      for(;;){ long r = pg.fill(buf.data(), buf.size()); if (r < 0) { pok = false; break; } if (r == 0) break; got.insert(got.end(), buf.data(), buf.data()+r); } }
    if (pok && zok) { n_ok++; if (got != want) { n_mismatch++; printf("MISMATCH it %d kind %d\n", it, kind); } }
    else if (!pok && !zok) n_err++;
    else { printf("DISAGREE it %d kind %d pok %d zok %d got %zu want %zu\n", it, kind, pok, zok, got.size(), want.size()); n_mismatch++; }
  }
  // the byte-mode decoder of single raw deflate streams (BGZF blocks): slices of the text deflated
  // at random levels, intact or damaged, against zlib
  {
    std::vector<uint8_t> text;
    { z_stream zs; memset(&zs,0,sizeof zs); inflateInit2(&zs, 31); zs.next_in = z.data(); zs.avail_in = z.size(); std::vector<uint8_t> buf(1<<20);
      for(;;){ zs.next_out = buf.data(); zs.avail_out = buf.size(); int r = inflate(&zs, Z_NO_FLUSH); text.insert(text.end(), buf.data(), buf.data()+(buf.size()-zs.avail_out)); if (r != Z_OK) break; }
      inflateEnd(&zs); }
    int raw_ok = 0, raw_err = 0;
    for (int it = 0; it < iters * 4 && text.size() > 70000; it++) {
      size_t len = 1 + rng() % 65536, at = rng() % (text.size() - len);
      if (it % 7 == 0) len = rng() % 40;
      std::vector<uint8_t> comp(len + len / 8 + 256);
      z_stream zs; memset(&zs,0,sizeof zs); deflateInit2(&zs, (int)(rng() % 10), Z_DEFLATED, -15, 8, it % 5 == 0 ? Z_FIXED : Z_DEFAULT_STRATEGY);
      zs.next_in = text.data() + at; zs.avail_in = len; zs.next_out = comp.data(); zs.avail_out = comp.size(); deflate(&zs, Z_FINISH); comp.resize(zs.total_out); deflateEnd(&zs);
      int kind = rng() % 3;
      if (kind == 1 && comp.size() > 2) comp[rng() % comp.size()] ^= 1 << (rng() % 8);
      if (kind == 2 && comp.size() > 2) comp.resize(1 + rng() % (comp.size() - 1));
      std::vector<uint8_t> want(len + 1); bool zok;
      { z_stream is; memset(&is,0,sizeof is); inflateInit2(&is, -15); is.next_in = comp.data(); is.avail_in = comp.size(); is.next_out = want.data(); is.avail_out = want.size();
        int r = inflate(&is, Z_FINISH); zok = r == Z_STREAM_END && is.total_out == len; inflateEnd(&is); }
      std::vector<uint8_t> scratch(len + 2048);
      bool pok = hlm_pgz::inflate_raw(comp.data(), comp.size(), scratch.data(), len);
      if (pok != zok || (pok && memcmp(scratch.data(), want.data(), len) != 0)) { n_mismatch++; printf("RAW MISMATCH it %d kind %d len %zu pok %d zok %d\n", it, kind, len, pok, zok); }
      else if (pok) raw_ok++; else raw_err++;
    }
    printf("raw streams: ok %d err %d\n", raw_ok, raw_err);
  }
  printf("ok %d err %d mismatch %d\n", n_ok, n_err, n_mismatch);
  return n_mismatch != 0;
}
