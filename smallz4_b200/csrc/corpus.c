/*
 * corpus.c -- deterministic synthetic corpora for the benchmarks and parity tests.
 *
 * Every byte is a pure function of (kind, seed, absolute offset): the generator works in
 * 64 KiB cells, so any rank can produce any byte range (its shard plus the 64 KiB halo in
 * front of it) without generating what precedes it.  BASELINE.json asks for: text-like
 * Markov, mixed binary, zeros/runs, incompressible random, and a "mixed" corpus of those.
 *
 * Built into libsz4corpus.so (not part of the compressor).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { SZ4_TEXT = 0, SZ4_BINARY = 1, SZ4_RUNS = 2, SZ4_ZEROS = 3, SZ4_RANDOM = 4, SZ4_MIXED = 5 };
enum { CELL = 65536, VOCAB = 4096, FOLLOW = 4 };

static uint64_t mix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
typedef struct { uint64_t s; } rng;
static uint64_t next64(rng* r) { r->s += 0x9E3779B97F4A7C15ull; return mix64(r->s); }
static uint32_t below(rng* r, uint32_t n) { return (uint32_t)((next64(r) >> 32) * (uint64_t)n >> 32); }

/* ---- text: Zipf-distributed words with first-order (word -> word) Markov preferences ---- */
typedef struct
{
  uint64_t seed; int ready;
  uint8_t  word[VOCAB][14]; uint8_t wlen[VOCAB];
  uint16_t follow[VOCAB][FOLLOW];
  uint16_t zipf[65536];                    /* inverse-CDF table */
} vocab_t;
static vocab_t g_vocab;

static void build_vocab(uint64_t seed)
{
  if (g_vocab.ready && g_vocab.seed == seed) return;
  rng r = { mix64(seed ^ 0x7465787421ull) };
  static const char letters[] = "etaoinshrdlucmfwypvbgkqjxz";
  for (int w = 0; w < VOCAB; w++)
  {
    int len = 1 + (int)below(&r, 3) + (int)below(&r, 4) + (int)below(&r, 5);
    if (w < 64 && len > 4) len = 2 + w % 3;
    g_vocab.wlen[w] = (uint8_t)len;
    for (int k = 0; k < len; k++)
    {
      uint32_t a = below(&r, 26), b = below(&r, 26);
      g_vocab.word[w][k] = (uint8_t)letters[a < b ? a : b];      /* skew towards frequent letters */
    }
    if (below(&r, 16) == 0) g_vocab.word[w][0] = (uint8_t)(g_vocab.word[w][0] - 32);
  }
  /* zipf(s=1) inverse CDF sampled at 65536 points */
  double tot = 0, acc = 0;
  static double wgt[VOCAB];
  for (int w = 0; w < VOCAB; w++) { wgt[w] = 1.0 / (1.0 + w); tot += wgt[w]; }
  int w = 0; acc = wgt[0] / tot;
  for (int k = 0; k < 65536; k++)
  {
    double u = (k + 0.5) / 65536.0;
    while (u > acc && w < VOCAB - 1) { w++; acc += wgt[w] / tot; }
    g_vocab.zipf[k] = (uint16_t)w;
  }
  for (int v = 0; v < VOCAB; v++)
    for (int k = 0; k < FOLLOW; k++)
      g_vocab.follow[v][k] = g_vocab.zipf[below(&r, 65536)];
  g_vocab.seed = seed; g_vocab.ready = 1;
}

static void cell_text(uint8_t* out, uint64_t seed, uint64_t cell)
{
  build_vocab(seed);
  rng r = { mix64(seed * 31 + cell * 0x100000001B3ull + 1) };
  uint32_t o = 0, prev = g_vocab.zipf[below(&r, 65536)], since_nl = 0;
  while (o < CELL)
  {
    uint32_t pick = below(&r, 100);
    uint32_t w = pick < 55 ? g_vocab.follow[prev][below(&r, FOLLOW)] : g_vocab.zipf[below(&r, 65536)];
    for (uint32_t k = 0; k < g_vocab.wlen[w] && o < CELL; k++) out[o++] = g_vocab.word[w][k];
    since_nl += g_vocab.wlen[w] + 1;
    uint32_t p = below(&r, 64);
    if (o < CELL && p == 0) out[o++] = ',';
    if (o < CELL && p == 1) out[o++] = '.';
    if (o < CELL) { if (since_nl > 60 + below(&r, 30)) { out[o++] = '\n'; since_nl = 0; } else out[o++] = ' '; }
    prev = w;
  }
}

/* ---- binary: fixed-layout records with counters, enums, noise and padding ---- */
static void cell_binary(uint8_t* out, uint64_t seed, uint64_t cell)
{
  rng r = { mix64(seed * 131 + cell * 0x9E3779B1ull + 2) };
  uint32_t rec = 24 + 8 * (uint32_t)(mix64(seed ^ (cell >> 3)) % 6);       /* 24..64 bytes */
  uint32_t counter = (uint32_t)(cell * (CELL / 16));
  uint32_t o = 0;
  while (o < CELL)
  {
    uint8_t tmp[64];
    memset(tmp, 0, sizeof(tmp));
    memcpy(tmp, &counter, 4); counter++;
    tmp[4] = (uint8_t)below(&r, 4); tmp[5] = (uint8_t)(0x40 + below(&r, 3)); tmp[6] = 0; tmp[7] = (uint8_t)0x80;
    uint64_t noise = next64(&r);
    memcpy(tmp + 8, &noise, (below(&r, 4) == 0) ? 8 : 3);
    uint32_t tag = below(&r, 12);
    memcpy(tmp + 16, "TAG", 3); tmp[19] = (uint8_t)('A' + tag);
    for (uint32_t k = 20; k < rec; k++) tmp[k] = (k & 7) == 0 ? (uint8_t)below(&r, 256) : (uint8_t)(k * 7 + tag);
    uint32_t take = rec < CELL - o ? rec : CELL - o;
    memcpy(out + o, tmp, take); o += take;
  }
}

/* ---- runs: short runs, plus groups of cells that are one byte (runs far above 65 299) ---- */
static void cell_runs(uint8_t* out, uint64_t seed, uint64_t cell)
{
  uint64_t g = mix64(seed * 17 + (cell >> 2) * 0x51ED27ull + 3);
  if ((g & 3) == 0) { memset(out, (int)((g >> 8) & 0xFF), CELL); return; }
  rng r = { mix64(seed * 19 + cell * 0xC2B2AE35ull + 4) };
  uint32_t o = 0;
  while (o < CELL)
  {
    uint32_t len = 1 + below(&r, 8);
    uint32_t k = below(&r, 100);
    if (k < 20) len = 4 + below(&r, 300);
    if (k == 0) len = 1000 + below(&r, 20000);
    uint8_t b = (uint8_t)(below(&r, 100) < 70 ? below(&r, 4) * 0x55 : below(&r, 256));
    for (uint32_t j = 0; j < len && o < CELL; j++) out[o++] = b;
  }
}

static void cell_random(uint8_t* out, uint64_t seed, uint64_t cell)
{
  rng r = { mix64(seed * 23 + cell * 0x27D4EB2Full + 5) };
  for (uint32_t o = 0; o < CELL; o += 8) { uint64_t v = next64(&r); memcpy(out + o, &v, 8); }
}

static int mixed_kind(uint64_t seed, uint64_t cell)
{
  uint32_t k = (uint32_t)(mix64(seed * 29 + (cell >> 2) * 0x165667B1ull + 6) % 100);   /* 256 KiB regions */
  if (k < 50) return SZ4_TEXT;
  if (k < 75) return SZ4_BINARY;
  if (k < 85) return SZ4_RUNS;
  if (k < 95) return SZ4_RANDOM;
  return SZ4_ZEROS;
}

static void make_cell(uint8_t* out, int kind, uint64_t seed, uint64_t cell)
{
  if (kind == SZ4_MIXED) kind = mixed_kind(seed, cell);
  switch (kind)
  {
    case SZ4_TEXT:   cell_text(out, seed, cell);   break;
    case SZ4_BINARY: cell_binary(out, seed, cell); break;
    case SZ4_RUNS:   cell_runs(out, seed, cell);   break;
    case SZ4_ZEROS:  memset(out, 0, CELL);         break;
    default:         cell_random(out, seed, cell); break;
  }
}

/* fill dst[0..n) with corpus bytes [offset, offset+n) */
void sz4_corpus_fill(uint8_t* dst, uint64_t offset, uint64_t n, int kind, uint64_t seed)
{
  uint8_t* tmp = (uint8_t*)malloc(CELL);
  uint64_t done = 0;
  while (done < n)
  {
    uint64_t at = offset + done, cell = at / CELL, in = at % CELL;
    uint64_t take = CELL - in; if (take > n - done) take = n - done;
    if (in == 0 && take == CELL) make_cell(dst + done, kind, seed, cell);
    else { make_cell(tmp, kind, seed, cell); memcpy(dst + done, tmp + in, take); }
    done += take;
  }
  free(tmp);
}
