/*
 * lz4cat_oracle.c -- TEST INFRASTRUCTURE ONLY (see smallz4_oracle.h).
 *
 * In-memory restatement of the reference decoder /root/reference/smallz4cat.c
 * (unlz4_userPtr, smallz4cat.c:112-360).  The reference streams through a 64 KiB
 * ring; here the whole output is one flat buffer with the dictionary placed in
 * front of it, which gives the same bytes for every well-formed frame.
 */
#include "smallz4_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct rd { const uint8_t* p; size_t n, at; int bad; } rd;

static uint8_t take(rd* r)
{
  if (r->at >= r->n) { r->bad = 1; return 0; }   /* smallz4cat.c:90 "out of data" */
  return r->p[r->at++];
}

int64_t sz4o_decompress(const uint8_t* src, size_t n, const uint8_t* dict, size_t dict_len,
                        uint8_t* dst, size_t cap)
{
  rd r = { src, n, 0, 0 };

  /* smallz4cat.c:115-123 */
  uint32_t sig = take(&r); sig |= (uint32_t)take(&r) << 8; sig |= (uint32_t)take(&r) << 16; sig |= (uint32_t)take(&r) << 24;
  const int modern = (sig == 0x184D2204u), legacy = (sig == 0x184C2102u);
  if (r.bad || (!modern && !legacy)) return -1;

  int block_sum = 0, content_sum = 0;
  if (modern)
  {
    /* smallz4cat.c:129-158 */
    uint8_t flags = take(&r);
    block_sum = flags & 16; content_sum = flags & 4;
    if ((flags >> 6) != 1) return -1;
    int ignore = 1 + ((flags & 8) ? 8 : 0) + ((flags & 1) ? 4 : 0) + 1;
    while (ignore--) take(&r);
    if (r.bad) return -1;
  }

  /* smallz4cat.c:169-187: only the last 64 KiB of a dictionary are reachable */
  if (dict && dict_len > 65536) { dict += dict_len - 65536; dict_len = 65536; }
  if (!dict) dict_len = 0;
  uint8_t* hist = (uint8_t*)malloc(65536 + cap + 8);
  if (!hist) return -2;
  memset(hist, 0, 65536);
  if (dict_len) memcpy(hist + 65536 - dict_len, dict, dict_len);
  uint8_t* out = hist + 65536;
  size_t o = 0;
  int64_t rc = -1;

  for (;;)
  {
    /* smallz4cat.c:193-205 */
    uint32_t bsz = take(&r); bsz |= (uint32_t)take(&r) << 8; bsz |= (uint32_t)take(&r) << 16; bsz |= (uint32_t)take(&r) << 24;
    if (r.bad)
    {
      if (legacy && r.at >= r.n) break;      /* legacy frames simply end with the data */
      goto fail;
    }
    int packed = legacy || (bsz & 0x80000000u) == 0;
    if (modern) bsz &= 0x7FFFFFFFu;
    if (bsz == 0) break;

    if (packed)
    {
      /* smallz4cat.c:207-323 */
      size_t used = 0, wrote = 0;
      while (used < bsz)
      {
        uint8_t tok = take(&r); used++;
        size_t lits = tok >> 4;
        if (lits == 15) { uint8_t b; do { b = take(&r); lits += b; used++; } while (b == 255 && !r.bad); }
        used += lits;
        if (r.bad || r.at + lits > r.n) goto fail;
        if (o + lits > cap) { rc = -2; goto fail; }
        memcpy(out + o, r.p + r.at, lits); r.at += lits; o += lits; wrote += lits;
        if (used == bsz) break;

        uint32_t delta = take(&r); delta |= (uint32_t)take(&r) << 8; used += 2;
        if (delta == 0 || r.bad) goto fail;                       /* smallz4cat.c:266 */
        size_t mlen = 4 + (tok & 15);
        if (mlen == 19) { uint8_t b; do { b = take(&r); mlen += b; used++; } while (b == 255 && !r.bad); }
        if (r.bad) goto fail;
        if (o + mlen > cap) { rc = -2; goto fail; }
        if (delta > o + 65536) goto fail;
        const uint8_t* from = out + o - delta;
        for (size_t k = 0; k < mlen; k++) out[o + k] = from[k];   /* overlap-safe */
        o += mlen; wrote += mlen;
      }
      if (legacy && wrote < 8u * 1024 * 1024) break;              /* smallz4cat.c:326 */
    }
    else
    {
      /* smallz4cat.c:331-342 */
      if (r.at + bsz > r.n) goto fail;
      if (o + bsz > cap) { rc = -2; goto fail; }
      memcpy(out + o, r.p + r.at, bsz); r.at += bsz; o += bsz;
    }
    if (block_sum) { take(&r); take(&r); take(&r); take(&r); }
  }
  if (content_sum) { take(&r); take(&r); take(&r); take(&r); }

  memcpy(dst, out, o);
  rc = (int64_t)o;
fail:
  free(hist);
  return rc;
}
