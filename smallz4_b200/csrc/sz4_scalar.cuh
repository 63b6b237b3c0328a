// sz4_scalar.cuh -- exact single-thread match finder for the one configuration the parallel
// finder does not cover: a dictionary stream (ring reads shifted by one slot, DESIGN.md Q-dict)
// that also contains runs long enough for the long-run shortcut (smallz4.h:632), where the
// reference leaves stale ring slots behind.  One device thread replays smallz4.h:603-747 with the
// reference's own data structures (lastHash table, 65536-slot rings).  Slow by construction and
// only reachable with -D plus a run of more than 65 000 equal bytes; phases 3a/3b still run in
// their parallel kernels afterwards.
#pragma once
#include "sz4_device.cuh"

namespace sz4
{
struct ScalarRings
{
  unsigned long long* last_hash;   // 1 << 20 entries, all ones = NoLastHash (smallz4.h:514)
  uint16_t* ring_hash;             // previousHash  (65536 slots)
  uint16_t* ring_exact;            // previousExact (65536 slots)
};

__device__ inline void scalar_longest(const uint8_t* buf, const uint16_t* ring_exact, uint32_t pos, uint32_t stop,
                                      uint32_t max_chain, uint32_t& out_len, uint32_t& out_dist)
{
  uint32_t len = 1, dist = 0, budget = max_chain;
  uint32_t hop = ring_exact[pos & 0xFFFF];
  uint32_t back = 0;
  while (hop != 0)
  {
    back += hop;
    if (back > kWindow) break;
    hop = ring_exact[(pos - back) & 0xFFFF];
    const uint32_t need = len + 1;
    if (pos + need > stop) break;
    bool ok = true;
    for (int32_t off = (int32_t)need - 4; off > 0; off -= 4)
      if (ld32u(buf + pos + off) != ld32u(buf + pos + off - back)) { ok = false; break; }
    if (!ok) continue;
    uint32_t f = need;
    while (pos + f + 4 <= stop && ld32u(buf + pos + f) == ld32u(buf + pos + f - back)) f += 4;
    while (pos + f < stop && buf[pos + f] == buf[pos + f - back]) f++;
    dist = back; len = f;
    if (--budget == 0) break;
  }
  out_len = len; out_dist = dist;
}

// state: [ last_hash (8 MiB) | ring_hash (128 KiB) | ring_exact (128 KiB) ], last_hash preset to all ones
__global__ void k_scalar_find(const uint8_t* buf, unsigned long long* state, uint16_t* ph_unused, uint16_t* pe_out,
                              uint32_t* mlen, uint16_t* mdist, Geom g)
{
  (void)ph_unused;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long* last_hash = state;
  uint16_t* ring_hash = (uint16_t*)(state + (1u << kHashBits));
  uint16_t* ring_exact = ring_hash + 65536;
  for (uint32_t k = 0; k < 65536; k++) { ring_hash[k] = 0; ring_exact[k] = 0; }

  const bool greedy = g.max_chain <= kGreedyMax;
  const bool lazy = !greedy && g.max_chain <= kLazyMax;
  for (uint32_t j = 0; j < g.n_blocks; j++)
  {
    const uint32_t blk = block_begin(g, j), end = block_end(g, j);
    const int64_t n = (int64_t)(end - blk);
    const uint32_t floor_pos = (j == 0 && g.stream_first) ? 0 : (blk > kWindow ? blk - kWindow : 0);
    int64_t lookback = (j == 0 && g.stream_first) ? (int64_t)(g.halo - g.first_ins) : (int64_t)kEndNoMatch;
    uint64_t skip = 0;
    bool peek = false;
    uint32_t prev_len = 0, prev_dist = 0;
    for (int64_t i = -lookback; i + kEndNoMatch <= n; i++)
    {
      const uint32_t pos = (uint32_t)((int64_t)blk + i);
      if (i > 0 && buf[pos] == buf[pos - 1] && prev_dist == 1 && prev_len > kSameLetter)
      {
        prev_len -= 1;
        mlen[pos] = prev_len; mdist[pos] = 1;
        continue;
      }
      prev_len = 0; prev_dist = 0;
      const uint32_t four = ld32u(buf + pos);
      const uint32_t h = hash20(four);
      const unsigned long long seen = last_hash[h];
      last_hash[h] = pos;
      const uint32_t slot = (uint32_t)i & 0xFFFF;
      if (seen == ~0ull) { ring_hash[slot] = 0; ring_exact[slot] = 0; continue; }
      uint32_t gap = pos - (uint32_t)seen;
      if (gap > kWindow) { ring_hash[slot] = 0; ring_exact[slot] = 0; continue; }
      ring_hash[slot] = (uint16_t)gap;
      uint32_t at = (uint32_t)seen;
      uint32_t there = ~four;
      while (at >= floor_pos)
      {
        there = ld32u(buf + at);
        if (there == four) break;
        if (hash20(there) != h) break;
        uint32_t step = ring_hash[at & 0xFFFF];
        if (step == 0) break;
        gap += step;
        if (gap > kWindow) break;
        at -= step;
      }
      if (there != four) { ring_exact[slot] = 0; continue; }
      ring_exact[slot] = (uint16_t)gap;
      if (i >= 0) pe_out[pos] = (uint16_t)gap;
      if (i < 0) continue;
      if (skip > 0)
      {
        skip--;
        if (!peek) continue;
        peek = false;
      }
      uint32_t len, dist;
      scalar_longest(buf, ring_exact, pos, end - kEndLiterals, g.max_chain, len, dist);
      mlen[pos] = len; mdist[pos] = (uint16_t)dist;
      prev_len = len; prev_dist = dist;
      if ((lazy || greedy) && len != 1)
      {
        peek = (skip == 0);
        skip = len;
      }
    }
  }
}

}  // namespace sz4
