/*
 * smallz4_oracle.c -- TEST INFRASTRUCTURE ONLY (see smallz4_oracle.h).
 *
 * Sequential CPU restatement of the reference's per-block loop.  It is organised
 * differently from the reference (whole input in memory, absolute positions, one
 * function per phase) but every decision follows the reference line cited next to it.
 * Anything that looks odd here (ring slots addressed two different ways, a position
 * inserted twice, costs that ignore literals in front of the cursor) is reference
 * behaviour that the CUDA path has to reproduce bit for bit.
 *
 * Reference: /root/reference/smallz4.h (smalLZ4 1.5).
 */
#include "smallz4_oracle.h"

#include <stdlib.h>
#include <string.h>

enum
{
  kMinMatch      = 4,               /* smallz4.h:95  MinMatch          */
  kEndNoMatch    = 12,              /* smallz4.h:99  BlockEndNoMatch   */
  kEndLiterals   = 5,               /* smallz4.h:101 BlockEndLiterals  */
  kHashBits      = 20,              /* smallz4.h:104                   */
  kWindow        = 65535,           /* smallz4.h:111 MaxDistance       */
  kSameLetter    = 19 + 255 * 256,  /* smallz4.h:118 MaxSameLetter     */
  kGreedyMax     = 3,               /* smallz4.h:77  ShortChainsGreedy */
  kLazyMax       = 6,               /* smallz4.h:79  ShortChainsLazy   */
  kBlockModern   = 4 * 1024 * 1024, /* smallz4.h:124                   */
  kBlockLegacy   = 8 * 1024 * 1024  /* smallz4.h:127                   */
};

#define NO_POS UINT64_MAX           /* smallz4.h:514 NoLastHash */

typedef struct cand { uint32_t len; uint16_t dist; } cand;

typedef struct octx
{
  const uint8_t* buf;        /* dictionary prefix (if any) followed by the input          */
  uint64_t       total;      /* bytes in buf                                              */
  uint64_t*      last_hash;  /* smallz4.h:515 lastHash                                    */
  uint16_t*      ring_hash;  /* smallz4.h:518 previousHash,  slot = index & 0xFFFF        */
  uint16_t*      ring_exact; /* smallz4.h:519 previousExact                               */
  uint32_t       max_chain;
  sz4o_stats*    st;
  const sz4o_trace* tr;
} octx;

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* smallz4.h:164 getHash32 */
static uint32_t hash20(uint32_t four) { return ((four * 48271u) >> (32 - kHashBits)) & ((1u << kHashBits) - 1); }

/* smallz4.h:173 findLongestMatch.  pos/stop are absolute offsets into buf.  The ring is read
   with the ABSOLUTE position (smallz4.h:190,200) although it was written with the
   block-relative one (smallz4.h:656): identical unless a dictionary shifts the blocks. */
static cand longest_match(const octx* c, uint64_t pos, uint64_t stop)
{
  cand best = { 1, 0 };
  uint32_t budget = c->max_chain;
  const uint8_t* cur = c->buf + pos;
  const uint8_t* lim = c->buf + stop;

  uint16_t hop  = c->ring_exact[pos & 0xFFFF];
  int64_t  back = 0;
  while (hop != 0)
  {
    back += hop;
    if (back > kWindow)
      break;
    hop = c->ring_exact[(pos - (uint64_t)back) & 0xFFFF];
    if (c->st) c->st->chain_hops++;

    const uint8_t* need = cur + best.len + 1;            /* first byte a longer match must cover */
    if (need > lim)
      break;

    const uint8_t* probe = need - 4;                     /* backwards, 4 bytes at a time */
    while (probe > cur && rd32(probe) == rd32(probe - back))
      probe -= 4;
    if (probe > cur)
      continue;

    const uint8_t* fwd = need;                           /* forwards */
    while (fwd + 4 <= lim && rd32(fwd) == rd32(fwd - back))
      fwd += 4;
    while (fwd < lim && *fwd == *(fwd - back))
      fwd++;

    best.dist = (uint16_t)back;
    best.len  = (uint32_t)(fwd - cur);
    if (--budget == 0)
      break;
  }
  return best;
}

/* smallz4.h:603-747: hash chains + match finder for one block.
   blk/end: absolute block borders; floor_pos: dataZero; lookback >= 0 positions before blk. */
static void find_matches(octx* c, uint64_t blk, uint64_t end, uint64_t floor_pos,
                         int64_t lookback, cand* m)
{
  const int64_t  n      = (int64_t)(end - blk);
  const int      greedy = c->max_chain <= kGreedyMax;
  const int      lazy   = !greedy && c->max_chain <= kLazyMax;
  uint64_t skip = 0;         /* skipMatches   */
  int      peek = 0;         /* lazyEvaluation */
  int64_t  i;

  for (i = -lookback; i + kEndNoMatch <= n; i++)
  {
    if ((int64_t)blk + i < 0)
      continue;              /* 65536-byte dictionary: the reference's first iteration stores
                                NoLastHash over NoLastHash and zeroes two zero slots: a no-op */
    const uint64_t pos = (uint64_t)((int64_t)blk + i);
    const uint8_t* cur = c->buf + pos;

    /* smallz4.h:632 long runs of one byte: copy the predecessor's match, insert nothing */
    if (i > 0 && cur[0] == cur[-1])
    {
      cand prev = m[i - 1];
      if (prev.dist == 1 && prev.len > kSameLetter)
      {
        m[i].dist = 1;
        m[i].len  = prev.len - 1;
        if (c->st) c->st->selfmatch_skips++;
        if (c->tr && c->tr->skipped) c->tr->skipped[pos] = 1;
        continue;
      }
    }

    const uint32_t four = rd32(cur);
    const uint32_t h    = hash20(four);
    const uint64_t seen = c->last_hash[h];
    c->last_hash[h] = pos;

    const uint32_t slot = (uint32_t)i & 0xFFFF;          /* smallz4.h:656 block-relative slot */
    if (c->tr && c->tr->prev_exact) c->tr->prev_exact[pos] = 0;

    if (seen == NO_POS) { c->ring_hash[slot] = 0; c->ring_exact[slot] = 0; continue; }

    uint64_t gap = pos - seen;
    if (gap > kWindow)  { c->ring_hash[slot] = 0; c->ring_exact[slot] = 0; continue; }
    c->ring_hash[slot] = (uint16_t)gap;

    /* smallz4.h:681 follow the hash chain until the four bytes are really equal */
    uint64_t at = seen;
    uint32_t there;
    if (at < floor_pos)
    {
      /* UB-1: the reference reads up to 11 bytes in front of its buffer here (only reachable
         for lookback positions).  We define the outcome as "no exact predecessor". */
      if (c->st) c->st->oob_first_reads++;
      there = ~four;
    }
    else
      for (;;)
      {
        there = rd32(c->buf + at);
        if (there == four)
          break;
        if (hash20(there) != h)
          break;
        uint16_t step = c->ring_hash[at & 0xFFFF];       /* absolute slot, smallz4.h:694 */
        if (step == 0)
          break;
        gap += step;
        if (gap > kWindow)
          break;
        at -= step;
        if (at < floor_pos)
          break;
      }

    if (there != four) { c->ring_exact[slot] = 0; continue; }
    c->ring_exact[slot] = (uint16_t)gap;
    if (c->tr && c->tr->prev_exact) c->tr->prev_exact[pos] = (uint16_t)gap;

    if (i < 0)
      continue;

    /* smallz4.h:727 greedy / lazy levels search only some positions */
    if (skip > 0)
    {
      skip--;
      if (!peek)
        continue;
      peek = 0;
    }

    if (c->st) c->st->searches++;
    { uint64_t h0 = c->st ? c->st->chain_hops : 0;
    m[i] = longest_match(c, pos, end - kEndLiterals);
    if (c->st && c->tr && c->tr->hops) c->tr->hops[pos] = (uint32_t)(c->st->chain_hops - h0); }

    if ((lazy || greedy) && m[i].len != 1)
    {
      peek = (skip == 0);
      skip = m[i].len;
    }
  }
  /* smallz4.h:746 (for blocks shorter than 11 bytes the reference writes in front of its
     vector here; the values it would have written are never used) */
  if (i < 0) i = 0;
  for (; i < n; i++)
    m[i].len = 1;
}

/* smallz4.h:376 estimateCosts */
static void price_block(cand* m, int64_t n, uint32_t* cost)
{
  uint64_t run = kEndLiterals;
  memset(cost, 0, (size_t)n * sizeof(uint32_t));
  for (int64_t i = n - (1 + kEndLiterals); i >= 0; i--)
  {
    run++;
    uint32_t keep   = 1;
    uint32_t lowest = cost[i + 1] + 1;
    if (run >= 15 && (run == 15 || (run >= 15 + 255 && (run - 15) % 255 == 0)))
      lowest++;

    const cand x = m[i];
    if (x.len >= kSameLetter && x.dist == 1)
    {
      keep   = x.len;
      lowest = cost[i + x.len] + 1 + 2 + 1 + (x.len - 19) / 255;
    }
    else
    {
      uint32_t overhead = 1 + 2;
      uint32_t bump_at  = 18;
      for (uint32_t len = kMinMatch; len <= x.len; len++)
      {
        uint32_t here = cost[i + len] + overhead;
        if (here <= lowest) { lowest = here; keep = len; }
        if (len == bump_at) { overhead++; bump_at += 255; }
      }
    }
    cost[i]  = lowest;
    m[i].len = keep;
    if (keep != 1)
      run = 0;
  }
}

/* smallz4.h:259 selectBestMatches; returns bytes written to out */
static size_t emit_block(const cand* m, size_t n, const uint8_t* src, uint8_t* out)
{
  size_t o = 0, at = 0, lit_from = 0, lit_n = 0;
  while (at < n)
  {
    const cand x = m[at];
    int last = 0;
    if (x.len <= 1)
    {
      if (lit_n == 0) lit_from = at;
      lit_n++;
      at++;
      if (at < n) continue;
      last = 1;
    }
    else
      at += x.len;

    int ml = (int)x.len - kMinMatch;
    if (last) ml = 0;
    uint8_t tok = (ml < 15) ? (uint8_t)ml : 15;
    if (lit_n < 15)
      out[o++] = (uint8_t)(tok | (lit_n << 4));
    else
    {
      out[o++] = tok | 0xF0;
      int rest = (int)lit_n - 15;
      while (rest >= 255) { out[o++] = 255; rest -= 255; }
      out[o++] = (uint8_t)rest;
    }
    if (lit_n > 0)
    {
      memcpy(out + o, src + lit_from, lit_n);
      o += lit_n;
      if (last) break;
      lit_n = 0;
    }
    out[o++] = (uint8_t)(x.dist & 0xFF);
    out[o++] = (uint8_t)(x.dist >> 8);
    if (ml >= 15)
    {
      ml -= 15;
      while (ml >= 255) { out[o++] = 255; ml -= 255; }
      out[o++] = (uint8_t)ml;
    }
  }
  return o;
}

static uint32_t block_bytes(const sz4o_opts* o)
{
  if (o->block_size) return o->block_size;
  return o->legacy ? kBlockLegacy : kBlockModern;
}

size_t sz4o_bound(size_t n, const sz4o_opts* opts)
{
  size_t bs = block_bytes(opts);
  size_t blocks = n / bs + 1;
  /* greedy levels with a dictionary can emit 3 bytes per 2 input bytes (DESIGN.md "Q-dict") */
  return 2 * n + blocks * 8 + 64;
}

int64_t sz4o_compress(const uint8_t* src, size_t n, const sz4o_opts* opts,
                      uint8_t* dst, size_t cap, sz4o_stats* stats, const sz4o_trace* trace)
{
  const uint32_t bs = block_bytes(opts);
  if (bs % 65536 != 0 || bs < 131072) return -1;
  if (opts->max_chain > 65535) return -1;
  if (stats) memset(stats, 0, sizeof(*stats));

  /* smallz4.h:554-570: a dictionary is right-aligned in a 65535-byte prefix; the CLI
     (smallz4.cpp:291-302) keeps at most its last 65536 bytes */
  const uint8_t* dict = opts->dict;
  size_t dict_len = dict ? opts->dict_len : 0;
  if (dict_len > 65536) { dict += dict_len - 65536; dict_len = 65536; }
  const int with_dict = dict_len > 0;
  const size_t prefix = with_dict ? kWindow : 0;

  uint8_t* buf = (uint8_t*)calloc(prefix + n + 16, 1);
  if (!buf) return -1;
  if (with_dict)
  {
    size_t keep = dict_len < (size_t)kWindow ? dict_len : (size_t)kWindow;
    memcpy(buf + prefix - keep, dict + dict_len - keep, keep);
  }
  if (n) memcpy(buf + prefix, src, n);

  octx c;
  c.buf = buf; c.total = prefix + n; c.max_chain = opts->max_chain; c.st = stats; c.tr = trace;
  c.last_hash  = (uint64_t*)malloc(sizeof(uint64_t) << kHashBits);
  c.ring_hash  = (uint16_t*)calloc(65536, sizeof(uint16_t));
  c.ring_exact = (uint16_t*)calloc(65536, sizeof(uint16_t));
  cand*     m    = (cand*)malloc(sizeof(cand) * (size_t)bs);
  uint32_t* cost = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)bs);
  uint8_t*  tmp  = (uint8_t*)malloc(2 * (size_t)bs + 64);
  int64_t result = -1;
  size_t o = 0;
  if (!c.last_hash || !c.ring_hash || !c.ring_exact || !m || !cost || !tmp) goto done;
  for (size_t k = 0; k < ((size_t)1 << kHashBits); k++) c.last_hash[k] = NO_POS;

#define PUT(ptr, len) do { if (o + (len) > cap) goto done; memcpy(dst + o, (ptr), (len)); o += (len); } while (0)

  /* smallz4.h:479-496 */
  if (opts->legacy) { static const uint8_t h[4] = { 0x02, 0x21, 0x4C, 0x18 }; PUT(h, 4); }
  else              { static const uint8_t h[7] = { 0x04, 0x22, 0x4D, 0x18, 1 << 6, 7 << 4, 0xDF }; PUT(h, 7); }

  const int raw_only = (opts->max_chain == 0);
  uint64_t floor_pos = 0;                 /* dataZero  */
  uint64_t next = prefix;                 /* nextBlock */
  int dict_pending = with_dict;

  while (next < c.total)
  {
    const uint64_t blk = next;
    next = (c.total - blk > bs) ? blk + bs : c.total;
    const size_t len = (size_t)(next - blk);

    /* smallz4.h:615-624 */
    int64_t lookback = (int64_t)floor_pos;
    if (lookback > kEndNoMatch && !dict_pending) lookback = kEndNoMatch;
    if (dict_pending) lookback = (int64_t)dict_len;
    if (opts->legacy || raw_only) lookback = 0;

    size_t packed = 0;
    if (!raw_only)
    {
      memset(m, 0, sizeof(cand) * len);
      find_matches(&c, blk, next, floor_pos, lookback, m);
      if (trace && trace->len_found)  for (size_t k = 0; k < len; k++) trace->len_found[blk + k]  = m[k].len;
      if (trace && trace->dist_found) for (size_t k = 0; k < len; k++) trace->dist_found[blk + k] = m[k].dist;
      /* smallz4.h:755 */
      if (len > kEndNoMatch && opts->max_chain > kGreedyMax)
      {
        price_block(m, (int64_t)len, cost);
        if (trace && trace->cost) memcpy(trace->cost + blk, cost, len * sizeof(uint32_t));
      }
      if (trace && trace->len_final) for (size_t k = 0; k < len; k++) trace->len_final[blk + k] = m[k].len;
      packed = emit_block(m, len, buf + blk, tmp);
    }
    dict_pending = 0;

    /* smallz4.h:765-780 */
    int use_packed = (packed < len) && !raw_only;
    if (opts->legacy) use_packed = 1;
    uint32_t nbytes = (uint32_t)(use_packed ? packed : len);
    uint32_t tagged = nbytes | (use_packed ? 0u : 0x80000000u);
    uint8_t hdr[4] = { (uint8_t)tagged, (uint8_t)(tagged >> 8), (uint8_t)(tagged >> 16), (uint8_t)(tagged >> 24) };
    PUT(hdr, 4);
    if (use_packed) PUT(tmp, nbytes); else PUT(buf + blk, nbytes);
    if (stats) { stats->blocks++; if (!use_packed) stats->raw_blocks++; }

    /* smallz4.h:783-805 */
    if (opts->legacy)
    {
      floor_pos = next;
      memset(c.ring_hash, 0, 65536 * sizeof(uint16_t));
      memset(c.ring_exact, 0, 65536 * sizeof(uint16_t));
      for (size_t k = 0; k < ((size_t)1 << kHashBits); k++) c.last_hash[k] = NO_POS;
    }
    else if (next - floor_pos > (uint64_t)kWindow)
      floor_pos = next - kWindow;
  }

  /* smallz4.h:809 */
  if (!opts->legacy) { static const uint8_t z[4] = { 0, 0, 0, 0 }; PUT(z, 4); }
#undef PUT
  result = (int64_t)o;

done:
  free(buf); free(c.last_hash); free(c.ring_hash); free(c.ring_exact); free(m); free(cost); free(tmp);
  return result;
}
