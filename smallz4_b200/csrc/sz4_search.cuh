// sz4_search.cuh -- phase 2: longest match for every position (smallz4.h:173 findLongestMatch).
//
// Without a dictionary (the normal case) the search is split three ways, all exact (DESIGN.md "The match finder"):
//   k_start   one position per thread: the candidates the reference keeps while the match is shorter than 8 bytes are
//             the nearest positions with the same 4, 5, .. 8 bytes -- read straight off the tables of sz4_lsd.cuh;
//   k_search  one CTA per tile of kTile positions, for the searches that go on: the CTA stages its slice of the input
//             plus the 64 KiB of history in front of it, and the 8-byte chain entries (pe8) of the same range, into
//             shared memory (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier); every lane walks the 8-byte
//             chain of one position at a time and pulls the next one from the tile's queue when it is done
//             (persistent lanes), so that short and long chains mix inside a warp;
//   k_long    one warp per walk, for the few walks that go on for thousands of candidates: the chain is read 128
//             members at a time from the sorted arrays.
// With a dictionary k_search alone walks previousExact from the first candidate on, a literal restatement of the
// reference loop including what it does NOT check: the first candidate is accepted without comparing bytes 0..1,
// later ones without byte 0 (harmless without a dictionary because chain members share their first bytes; with a
// dictionary the shifted ring makes it visible, DESIGN.md Q-dict).
// Result either way: longest match, nearest candidate on ties, at most max_chain improvements.
#pragma once
#include "sz4_device.cuh"

namespace sz4
{
#ifndef SZ4_TILE
#define SZ4_TILE 11264
#endif
#ifndef SZ4_LOOK
#define SZ4_LOOK 1024
#endif
enum : uint32_t
{
  kTile          = SZ4_TILE, // positions per CTA
  kLook          = SZ4_LOOK, // bytes staged behind the tile for match extension
  kHist          = 65536,    // history staged in front of the tile
#ifndef SZ4_SEARCH_THREADS
#define SZ4_SEARCH_THREADS 1024
#endif
  kSearchThreads = SZ4_SEARCH_THREADS,
  kDataBytes     = kHist + kTile + kLook + 16,
  kChainElems    = kHist + 32 + kTile,
  kSearchSmem    = kDataBytes + 2 * kChainElems
};

#ifndef SZ4_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SZ4_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SZ4_DONE_%=;\n"
      "bra SZ4_WAIT_%=;\n"
      "SZ4_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// Copy two global ranges into shared memory.  Sizes and addresses are multiples of 16.
__device__ __forceinline__ void stage_two(unsigned char* dst_a, const unsigned char* src_a, uint32_t bytes_a,
                                          unsigned char* dst_b, const unsigned char* src_b, uint32_t bytes_b,
                                          uint64_t* bar, int use_bulk)
{
#ifndef SZ4_EMU
  if (use_bulk)
  {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0)
    {
      mbar_expect_tx(bar, bytes_a + bytes_b);
      const uint32_t kChunk = 16384;
      for (uint32_t o = 0; o < bytes_a; o += kChunk) bulk_g2s(dst_a + o, src_a + o, min(kChunk, bytes_a - o), bar);
      for (uint32_t o = 0; o < bytes_b; o += kChunk) bulk_g2s(dst_b + o, src_b + o, min(kChunk, bytes_b - o), bar);
    }
    mbar_wait(bar, 0);
    return;
  }
#endif
  (void)bar; (void)use_bulk;
  const uint4* sa = (const uint4*)src_a; uint4* da = (uint4*)dst_a;
  for (uint32_t k = threadIdx.x; k < bytes_a / 16; k += blockDim.x) da[k] = sa[k];
  const uint4* sb = (const uint4*)src_b; uint4* db = (uint4*)dst_b;
  for (uint32_t k = threadIdx.x; k < bytes_b / 16; k += blockDim.x) db[k] = sb[k];
  __syncthreads();
}

// Shared-memory loads by 32-bit shared address (CUDA build) so that the hot loop does not pay for
// generic-address arithmetic; plain pointers in the emulated build.
#ifdef SZ4_EMU
typedef const unsigned char* smem_addr;
__device__ __forceinline__ smem_addr smem_base(const void* p) { return (const unsigned char*)p; }
__device__ __forceinline__ uint32_t lds_u8(smem_addr a) { return *a; }
__device__ __forceinline__ uint32_t lds_u16(smem_addr a) { return *(const uint16_t*)a; }
__device__ __forceinline__ uint32_t lds_u32(smem_addr a) { return *(const uint32_t*)a; }
#else
typedef uint32_t smem_addr;
__device__ __forceinline__ smem_addr smem_base(const void* p) { return smem_u32(p); }
__device__ __forceinline__ uint32_t lds_u8(smem_addr a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(smem_addr a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(smem_addr a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#endif

struct SearchView
{
  smem_addr      s_data;         // staged bytes [dlo, dhi)
  smem_addr      s_pe;           // staged previousExact [clo, ...)
  const uint8_t* g_data;         // the whole batch in HBM (for extensions that leave the staged range)
  uint32_t dlo, dhi, clo, shift;

  __device__ __forceinline__ uint32_t chain(uint32_t r) const { return lds_u16(s_pe + 2 * (r - shift - clo)); }
  __device__ __forceinline__ uint32_t word_at(uint32_t pos) const
  {
    if (pos + 4 <= dhi)
    {
      const uint32_t a = pos - dlo;
      const smem_addr w = s_data + (a & ~3u);
      return __funnelshift_r(lds_u32(w), lds_u32(w + 4), (a & 3) * 8);
    }
    return ld32u(g_data + pos);
  }
};

// One candidate of the reference loop (smallz4.h:202-247): true if it is longer than the best so far,
// in which case len and tail are updated.  tail = the four bytes p+len-3 .. p+len, i.e. the first group
// the reference's backward scan compares (smallz4.h:224-225).  The caller has checked that a longer
// match still fits (smallz4.h:205).
// runs (may be null) = run_fwd of sz4_runs.cuh: long stretches of one byte are stepped over by their lengths
// instead of being compared word by word (same result; without it one position at the head of a run of a few
// hundred KiB keeps its CTA busy for milliseconds).  Only valid where candidates share p's first bytes (no dictionary).
// sure: the caller knows that the first len + 1 bytes are equal (a member of p's (len+1)-byte prefix class), phase 1 is skipped.
template <class View>
__device__ __forceinline__ bool try_candidate(const View& v, uint32_t p, uint32_t q, uint32_t stop, uint32_t& len, uint32_t& tail,
                                              const uint32_t* runs, bool sure = false)
{
  const uint32_t need = len + 1;
  if (len >= 4 && !sure)
  {
    // phase 1, smallz4.h:224-233: bytes (0, need) in 4-byte groups from the top
    if (tail != v.word_at(q + len - 3)) return false;                  // first group from the top
    int32_t off = (int32_t)need - 8;
    int32_t known = 0;                                                 // bytes [0, known) are the same byte on both sides
    if (runs != nullptr && off > 64) known = (int32_t)min(min(runs[p], runs[q]), 0x7fffffffu);
    for (; off > 0 && off + 4 > known; off -= 4)
      if (v.word_at(p + off) != v.word_at(q + off)) return false;
  }
  // phase 2, smallz4.h:236-243: four bytes at a time, then the last one to three -- here: the first byte that differs,
  // not beyond `stop`
  uint32_t f = need, uniform = 0;
  for (;;)
  {
    const uint32_t left = stop - (p + f);                              // (need <= stop - p: the caller's check)
    if (left == 0) break;
    const uint32_t w = v.word_at(p + f);
    const uint32_t x = w ^ v.word_at(q + f);
    if (x != 0 || left < 4)
    {
      const uint32_t same = x != 0 ? ((uint32_t)__ffs((int)x) - 1) >> 3 : 4u;
      f += min(same, left);
      break;
    }
    f += 4;
    if (runs != nullptr && w == __funnelshift_r(w, w, 8) && ++uniform >= 16)
    {
      // both sides are inside runs of the byte just compared: equal for as long as the shorter one lasts
      uint32_t s = min(runs[p + f - 1], runs[q + f - 1]) - 1;
      s = min(s, stop - (p + f));
      f += s;
      uniform = 0;
    }
  }
  len = f;
  tail = v.word_at(p + f - 3);
  return true;
}

// Stretch shortcut.  p lies in a run of one byte b (R = run_fwd[p] >= 4), so every member of its chain
// starts four b's.  `top` = p - total is a candidate whose chain entry is 1: the candidates
// q_k = top - k, k = 0..s (s = ones_back[top]) lie in ONE run of b's and have R0 + k bytes of b
// in front of them (R0 = run_fwd[top]).  Their match lengths with p follow from the run lengths:
//     R0+k <  R :  min(C, R0+k)            (q's run ends first)
//     R0+k >  R :  min(C, R)               (p's run ends first)
//     R0+k == R :  needs a real comparison (both runs end together)
// with C = stop - p.  The reference visits them in order k = 0, 1, ... and takes every strictly
// longer one (smallz4.h:232,246-251); that sequence is replayed here in closed form.
// Returns true when the walk is over; otherwise total/hop are left at the last member of the stretch.
__device__ __forceinline__ bool walk_stretch(const SearchView& v, const uint32_t* run_fwd, const uint16_t* ones_back,
                                             uint32_t p, uint32_t stop, uint32_t R, uint32_t& total, uint32_t& hop,
                                             uint32_t& len, uint32_t& dist, uint32_t& budget, uint32_t& tail, uint32_t limit)
{
  const uint32_t len_in = len;
  const uint32_t C = stop - p;
  const uint32_t top = p - total;
  const uint32_t s = ones_back[top];
  const uint32_t kmax = min(s, limit - total);
  const uint32_t R0 = run_fwd[top];
  if (R0 > R)
  {
    // every member is deeper inside its run than p: only the first one can be an improvement
    const uint32_t Rc = min(R, C);
    if (Rc > len) { len = Rc; dist = total; if (--budget == 0) return true; }
  }
  else
  {
    const uint32_t kstar = R - R0;                       // the member whose run ends together with p's
    if (kstar > 0)
    {
      const uint32_t kend = min(kmax, kstar - 1);
      const uint32_t k1 = len + 1 > R0 ? len + 1 - R0 : 0;   // first member with more than len bytes
      if (k1 <= kend)
      {
        if (R0 + k1 >= C) { len = C; dist = total + k1; budget--; return true; }
        const uint32_t klast = min(kend, C - R0);        // lengths R0+k1 .. R0+klast, each one longer
        const uint32_t cnt = klast - k1 + 1;
        if (budget <= cnt) { const uint32_t k = k1 + budget - 1; len = R0 + k; dist = total + k; budget = 0; return true; }
        budget -= cnt; len = R0 + klast; dist = total + klast;
        if (len == C) return true;
      }
    }
    if (kstar <= kmax)
    {
      if (p + len + 1 > stop) return true;
      if (len != len_in) tail = v.word_at(p + len - 3);
      if (try_candidate(v, p, top - kstar, stop, len, tail, run_fwd))
      {
        dist = total + kstar;
        if (--budget == 0) return true;
      }
    }
  }
  if (kmax < s) return true;                             // the next hop of 1 would exceed 65535 (smallz4.h:196) or the chain's end
  if (len != len_in) tail = v.word_at(p + len - 3);
  total += s;
  hop = v.chain(p - total);
  return p + len + 1 > stop;
}

// lane states of the walk
// kWalk: in the fast loop (the bytes a longer match needs first are staged); kSlowWalk: hops one candidate per round in the slow part
enum : uint32_t { kIdle = 0, kWalk = 1, kCheck = 2, kStretch = 3, kFinish = 4, kSlowWalk = 5 };
#ifndef SZ4_FAST_BLOCK
#define SZ4_FAST_BLOCK 8
#endif
#ifndef SZ4_STRETCHES
#define SZ4_STRETCHES 8
#endif
enum : uint32_t { kFastHops = SZ4_FAST_BLOCK, kStretchesPerVisit = SZ4_STRETCHES };   // runs taken per visit of the slow part   // candidates per lane between two looks at how many lanes still walk

// The rejecting path of the walk (smallz4.h:192-233) for up to `hops` candidates per lane; see k_search.
#ifdef SZ4_SEARCH_STATS
__device__ unsigned long long g_stats[8];   // 0 hop slots (lanes x hops offered), 1 useful hops, 2 rounds, 3 slow-part lane events, 4 refills, 5 slow-part warp passes
#endif
__device__ __forceinline__ void fast_hops_loop(const SearchView& v, uint32_t hops, uint32_t& state, uint32_t& total, uint32_t& hop,
                                               uint32_t run, uint32_t tail, smem_addr cbase, smem_addr dl, uint32_t min_lanes, uint32_t limit)
{
  (void)v;
  uint32_t w0 = 0, w1 = 0;                                         // only looked at by lanes that loaded them
  // eight hops at a time, for as long as enough lanes are still walking: the parked ones wait for the slow part
#pragma unroll 1
  for (uint32_t it = 0; it < hops; it += kFastHops)
  {
    if (it != 0 && (uint32_t)__popc(__ballot_sync(0xffffffffu, state == kWalk)) < min_lanes) break;
    const uint32_t total_in = total;
#pragma unroll
    for (uint32_t un = 0; un < kFastHops; un++)
    {
      const uint32_t tot2 = total + hop;
      const bool ends = hop == 0 || tot2 > limit;                  // smallz4.h:192,196 (limit: 65535, or where the chain ends)
      const bool walking = state == kWalk;
      const bool go = walking && !ends;
      // the candidate q = p - tot2: its chain entry (at 65535 its value ends the walk either way) and its bytes
      // q+len-3 .. q+len, the group the reference compares first with p's (an unaligned 32-bit read from shared
      // memory: two aligned words and a funnel shift, which only looks at the low five bits of the shift)
      const smem_addr a = dl - tot2;
#ifdef SZ4_EMU
      const uint32_t lowbits = (uint32_t)(a - v.s_data);           // the staged bytes start at a multiple of 16
#else
      const uint32_t lowbits = a;
#endif
      uint32_t hop2 = hop;
      if (go)
      {
        const smem_addr w = a - (lowbits & 3u);
        hop2 = lds_u16(cbase - 2 * tot2);
        w0 = lds_u32(w); w1 = lds_u32(w + 4);
      }
      const bool same = __funnelshift_r(w0, w1, lowbits * 8) == tail;
      uint32_t next_state = same ? kCheck : kWalk;                 // kCheck: worth a closer look
      if (ends) next_state = kFinish;
      if (go) total = tot2;
#ifdef SZ4_SEARCH_STATS
      { const uint32_t m = __ballot_sync(0xffffffffu, go); if ((threadIdx.x & 31) == 0) { atomicAdd(&g_stats[0], 32ull); atomicAdd(&g_stats[1], (unsigned long long)__popc(m)); } }
#endif
      hop = hop2;
      if (walking) state = next_state;
    }
    // eight hops of one position each inside a byte run, and the next one is another: a stretch, which
    // walk_stretch does in closed form (shorter ones are cheaper to walk through the filter above)
    if (run != 0 && state == kWalk && hop == 1 && total - total_in == kFastHops) state = kStretch;
  }
}

// ---------------------------------------------------------------------------------------------
// Order of the tiles.  A tile ends with its longest walk, and tiles full of byte runs (every other run of the
// same byte in the window costs a round of the slow part) take up to 40 times the average: started last, one
// of them keeps a single SM busy for milliseconds after all others are done.  So the tiles are started in the
// order of an estimate of that cost: run positions (previousExact == 1) times run heads.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tile_cost(const uint16_t* pe, uint32_t tiles_per_block, Geom g, uint32_t* cost)
{
  __shared__ uint32_t ws[32], tot;
  const uint32_t j = blockIdx.x / tiles_per_block, t = blockIdx.x % tiles_per_block;
  const uint32_t t0 = block_begin(g, j) + t * kTile, s_end = search_end(g, j);
  // run positions (previousExact == 1) and run heads (the first of a stretch of them): a position inside a run
  // pays one round of the slow part for every other run of its byte in the window
  uint32_t mine = 0, heads = 0;
  if (t0 < s_end)
  {
    const uint32_t t1 = min(t0 + kTile, s_end);
    for (uint32_t p = t0 + threadIdx.x; p < t1; p += blockDim.x)
    {
      const bool one = pe[p] == 1;
      mine += one ? 1u : 0u;
      heads += (one && pe[p - 1] != 1) ? 1u : 0u;              // (p >= 1: the arrays are padded in front)
    }
  }
  (void)block_excl_scan(mine, ws, &tot);
  const uint32_t run_positions = tot;
  __syncthreads();
  (void)block_excl_scan(heads, ws, &tot);
  if (threadIdx.x == 0)
  {
    const unsigned long long c = (unsigned long long)run_positions * (tot + 1);
    cost[blockIdx.x] = c > 0xffffffffull ? 0xffffffffu : (uint32_t)c;
  }
}

// one CTA: stable partition of the tile numbers into eight classes of estimated cost, the dearest first
__device__ __forceinline__ uint32_t tile_class(uint32_t c)
{
  // eight classes: more than 128, 64, 32, 16, 8, 4 tiles' worth, more than one (any run structure at all), the rest
  const uint32_t x = c / kTile;
  if (x >= 128) return 0u;
  if (x >= 4) return 7u - (uint32_t)(31 - __clz((int)x));       // 64..127 -> 1, 32..63 -> 2, ..., 4..7 -> 5
  return c > kTile ? 6u : 7u;
}

__global__ void __launch_bounds__(256)
k_tile_order(const uint32_t* cost, uint32_t n_tiles, uint32_t* order)
{
  __shared__ uint32_t ws[32], tot;
  const uint32_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(threadIdx.x * per, n_tiles), hi = min(lo + per, n_tiles);
  uint32_t base = 0;
  for (uint32_t cls = 0; cls < 8; cls++)
  {
    uint32_t mine = 0;
    for (uint32_t i = lo; i < hi; i++) mine += tile_class(cost[i]) == cls ? 1u : 0u;
    uint32_t at = base + block_excl_scan(mine, ws, &tot);
    for (uint32_t i = lo; i < hi; i++) if (tile_class(cost[i]) == cls) order[at++] = i;
    base += tot;
    __syncthreads();
  }
}

#ifndef SZ4_LONG_WIDE
#define SZ4_LONG_WIDE 4
#endif
enum : uint32_t { kLongWide = SZ4_LONG_WIDE };
struct LongWalk { uint32_t p, len, dist, total, budget; };     // a walk k_search hands to k_long: state behind its last candidate

// How far back the reference's chain of position p reaches (no dictionary).  The tables of sz4_lsd.cuh are "pure"
// (previous position with the same prefix, at most 65535 back); the reference's chain is shorter in two cases:
//  * legacy frames clear the tables at every block (smallz4.h:783-795): nothing in front of the block;
//  * Q-twice: the position T twelve bytes in front of a block border is inserted again by the next block's
//    lookback (smallz4.h:615-629), finds itself at distance 0 and zeroes its own entries.  Every chain that meets T
//    ends there: the chains of the positions behind T whose four bytes hash like T's (smallz4.h:164,694-697).
// word_p = the four bytes at p.
__device__ __forceinline__ uint32_t chain_limit(const Geom& g, const uint8_t* data, uint32_t p, uint32_t word_p)
{
  if (g.legacy) return p >= g.halo ? min((uint32_t)kWindow, p - block_begin(g, (p - g.halo) / g.block_size)) : (uint32_t)kWindow;
  if (p + kEndNoMatch - 1 < g.halo) return kWindow;
  const uint32_t border = g.halo + (p + kEndNoMatch - 1 - g.halo) / g.block_size * g.block_size;   // the last one with T < p
  if (border < kEndNoMatch) return kWindow;
  const uint32_t T = border - kEndNoMatch;
  if (p - T > kWindow || !is_twice_inserted(g, T)) return kWindow;
  return hash20(ld32u(data + T)) == hash20(word_p) ? p - T : (uint32_t)kWindow;
}

// test hook (sz4_debug_fetch "pe"): previousExact as the reference's ring holds it, from the pure pe4 table
__global__ void __launch_bounds__(256)
k_debug_own4(const uint8_t* data, const uint16_t* jump16, uint16_t* out, Geom g)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.n_total) return;
  const uint32_t d4 = jump16[(size_t)(p + 4) * 4];
  const uint32_t lim = chain_limit(g, data, p, ld32u(data + p));
  out[p] = (d4 != 0 && d4 <= lim && !is_twice_inserted(g, p)) ? (uint16_t)d4 : (uint16_t)0;
}

// ---------------------------------------------------------------------------------------------
// k_start: the first part of every search, one position per thread (no dictionary).  It takes the candidates the
// reference keeps while the match is shorter than 8 bytes -- the nearest position with the same 4 bytes, then the
// nearest with the same len+1 bytes (sz4_lsd.cuh) -- straight from the tables.  Most searches end here.  The ones that
// reach 8 bytes and could still grow go into their tile's queue for k_search, which walks the 8-byte chain; their
// state travels in mlen / mdist: length | improvements << 24 | 1 << 31, distance of the last candidate.
// ---------------------------------------------------------------------------------------------
enum : uint32_t { kStartLenMask = 0x00ffffffu, kStartCountShift = 24, kStartPending = 0x80000000u };

struct GlobalView
{
  const uint8_t* g_data;
  __device__ __forceinline__ uint32_t word_at(uint32_t pos) const
  {
    const uint32_t* w = (const uint32_t*)(g_data + (pos & ~3u));      // (the batch starts on an aligned address)
    return __funnelshift_r(w[0], w[1], (pos & 3) * 8);
  }
};

__global__ void __launch_bounds__(256)
k_start(const uint8_t* data, const uint16_t* jump16, const uint16_t* pe8, const uint32_t* run_fwd, uint32_t* mlen, uint16_t* mdist,
        uint32_t* tile_count, uint32_t* tile_queue, uint32_t tiles_per_block, Geom g)
{
  const uint32_t p = g.halo + blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  GlobalView v; v.g_data = data;
  uint32_t tile = 0xffffffffu;                                       // the tile whose queue p goes into, if any
  if (p < g.n_total)
  {
    const uint32_t j = (p - g.halo) / g.block_size;
    const uint32_t b = block_begin(g, j);
    const uint32_t stop = block_end(g, j) - kEndLiterals;             // smallz4.h:736 end - BlockEndLiterals
    const uint32_t d4 = p < search_end(g, j) ? jump16[(size_t)(p + 4) * 4] : 0u;
    // the reference searches p only if it has an exact predecessor its chain can reach (smallz4.h:712-717)
    const uint32_t lim = d4 != 0 ? chain_limit(g, data, p, v.word_at(p)) : 0u;
    if (d4 != 0 && d4 <= lim)
    {
      uint32_t budget = g.max_chain, len = 1, tail = 0, total = d4, kept = 1;
      uint32_t run = run_fwd[p]; if (run < 8) run = 0;
      // the first candidate is accepted unseen (smallz4.h:224-233 compare nothing while the best length is 1)
      if (d4 == 1 && run != 0) len = min(run, stop - p);              // inside a run: the rest of the run
      else (void)try_candidate(v, p, p - total, stop, len, tail, run_fwd);
      bool finish = --budget == 0 || p + len + 1 > stop;
      while (!finish && len < 8)
      {
        const uint32_t d = len == 7 ? pe8[p] : jump16[(size_t)(p + len + 1) * 4 + (len - 3)];   // pe5..pe7, pe8
        if (d == 0 || d > lim) { finish = true; break; }
        total = d;
        (void)try_candidate(v, p, p - total, stop, len, tail, run_fwd, true);
        kept++;
        finish = --budget == 0 || p + len + 1 > stop;
      }
      mdist[p] = (uint16_t)total;
      if (finish) mlen[p] = len;
      else
      {
        mlen[p] = len | (kept << kStartCountShift) | kStartPending;
        tile = j * tiles_per_block + (p - b) / kTile;
      }
    }
  }
  // append to the tile's queue (a warp's positions nearly always belong to one tile: one atomic per warp)
  const uint32_t peers = __match_any_sync(0xffffffffu, tile);
  const int leader = __ffs((int)peers) - 1;
  uint32_t at = 0;
  if (tile != 0xffffffffu && lane == (uint32_t)leader) at = atomicAdd(&tile_count[tile], (uint32_t)__popc(peers));
  at = __shfl_sync(0xffffffffu, at, leader);
  if (tile != 0xffffffffu) tile_queue[(size_t)tile * kTile + at + (uint32_t)__popc(peers & ((1u << lane) - 1))] = p;
}

#ifdef SZ4_TILE_STATS
// debugging aid (tools/tile_stats.py): duration and start of every tile in microseconds
__device__ uint32_t g_tile_us[1 << 16], g_tile_t0[1 << 16];
__device__ __forceinline__ unsigned long long tile_clock() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

__global__ void __launch_bounds__(kSearchThreads, 1)
k_search(const uint8_t* data, const uint16_t* pe, const uint32_t* saved_pe, const uint32_t* run_fwd,
         const uint16_t* ones_back, uint32_t* mlen, uint16_t* mdist, uint32_t tiles_per_block, Geom g, int use_bulk,
         uint32_t fast_hops, uint32_t fast_lanes, uint32_t dense_a, uint32_t dense_b, const uint32_t* tile_order,
         const uint32_t* tile_count, const uint32_t* tile_queue, LongWalk* long_list, uint32_t* long_count, uint32_t long_cap,
         uint32_t long_age, uint32_t tail_lanes)
{
  SZ4_DYN_SMEM(smem);
  __shared__ uint64_t bar;
  __shared__ uint32_t next_pos;

  const uint32_t tile = tile_order != nullptr ? tile_order[blockIdx.x] : blockIdx.x;   // CTAs start in blockIdx order
#ifdef SZ4_TILE_STATS
  const unsigned long long tile_t0 = tile_clock();
#endif
  const uint32_t j = tile / tiles_per_block;
  const uint32_t t = tile % tiles_per_block;
  const uint32_t t0 = block_begin(g, j) + t * kTile;
  const uint32_t s_end = search_end(g, j);
  if (t0 >= s_end) return;
  if (tile_queue != nullptr && tile_count[tile] == 0) return;       // nothing left to walk in this tile
  const uint32_t t1 = min(t0 + kTile, s_end);
  const uint32_t stop = block_end(g, j) - kEndLiterals;               // smallz4.h:736 end - BlockEndLiterals

  SearchView v;
  v.g_data = data;
  v.shift = g.shift;
  v.dlo = (t0 > kHist ? t0 - kHist : 0) & ~15u;
  v.dhi = min(v.dlo + kDataBytes - 16, (g.n_total + 31) & ~15u);     // stays inside the zero padding
  v.clo = (t0 > kHist + 8 ? t0 - kHist - 8 : 0) & ~7u;
  const uint32_t chi = (t1 + 7) & ~7u;
  unsigned char* s_data = smem;
  uint16_t* s_pe = (uint16_t*)(smem + kDataBytes);
  v.s_data = smem_base(s_data);
  v.s_pe = smem_base(s_pe);

  // the one position of the tile (if any) whose chain entry was zeroed by the next block's lookback (Q-twice)
  __shared__ uint32_t tw_pos_s;
  if (threadIdx.x == 0)
  {
    next_pos = 0;
    uint32_t tw = 0xffffffffu;
    const uint32_t cand = block_end(g, j) - kEndNoMatch;
    if (block_len(g, j) >= kEndNoMatch && cand >= t0 && cand < t1 && is_twice_inserted(g, cand)) tw = cand;
    tw_pos_s = tw;
  }
  stage_two(s_data, data + v.dlo, v.dhi - v.dlo, (unsigned char*)s_pe, (const unsigned char*)(pe + v.clo),
            (chi - v.clo) * 2, &bar, use_bulk);
  __syncthreads();
  const uint32_t tw_pos = tw_pos_s;
  const uint32_t tw_own = tw_pos != 0xffffffffu ? saved_pe[(tw_pos + kEndNoMatch - g.halo) / g.block_size] : 0;

  const uint32_t lane = threadIdx.x & 31;
  // tile_queue != nullptr: k_start has done every search up to a match of 8 bytes; `pe` is pe8 (sz4_lsd.cuh) and the
  // tile's queue lists the positions whose walk goes on along the 8-byte chain.  nullptr: the classic walk along
  // previousExact from the first candidate on (dictionary streams).
  const bool jump = tile_queue != nullptr;
  const uint32_t tlen = jump ? tile_count[tile] : t1 - t0, n_pass = (dense_a != 0 && !jump) ? 3u : 1u;
  const uint32_t* queue = jump ? tile_queue + (size_t)tile * kTile : nullptr;
  const uint32_t* runs = g.shift == 0 ? run_fwd : nullptr;       // filled in only without a dictionary
  uint32_t state = kIdle;
  bool exhausted = false;
  uint32_t p = 0, len = 1, dist = 0, total = 0, hop = 0, budget = 0;
  uint32_t limit = kWindow;            // the chain of p ends this far back (chain_limit)
  uint32_t age = 0;                    // rounds this walk has been going on
  uint32_t run = 0;                    // bytes equal to data[p] from p on, when the stretch shortcut applies
  uint32_t tail = 0;                   // bytes p+len-3 .. p+len: the group a longer match has to reproduce first
  smem_addr cbase = v.s_pe, dl = v.s_data;   // &chain(p), &data[p + len - 3] in shared memory
  bool fast = false;                   // the next candidate's bytes up to q + len are inside the staged range

  // One ballot per iteration keeps `idle` (lanes that want a new position) current for all lanes.
  uint32_t idle = 0xffffffffu;
  for (;;)
  {
    // ---- refill idle lanes with the next positions of the tile (a few tries: positions without an exact
    // predecessor are not searched, smallz4.h:712-717, and must not cost their lane a whole round)
#pragma unroll 1
    for (uint32_t tries = 0; tries < 4 && idle; tries++)
    {
      uint32_t base = 0;
      const int leader = __ffs((int)idle) - 1;
      if (lane == (uint32_t)leader) base = atomicAdd(&next_pos, (uint32_t)__popc(idle));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (state == kIdle && !exhausted && jump)
      {
        const uint32_t idx = base + (uint32_t)__popc(idle & ((1u << lane) - 1));
        if (idx >= tlen) exhausted = true;
        else
        {
          // a search k_start left with a match of at least eight bytes: its last candidate is on p's 8-byte chain,
          // which is walked from there
          p = queue[idx];
          const uint32_t packed = mlen[p];
          len = packed & kStartLenMask;
          budget = g.max_chain - ((packed >> kStartCountShift) & 15u);
          total = mdist[p]; dist = total;
          limit = chain_limit(g, data, p, v.word_at(p));
          run = run_fwd[p]; if (run < 8) run = 0;                      // the 8-byte chain runs through p's run
          cbase = v.s_pe + 2 * (p - v.clo);
          tail = v.word_at(p + len - 3);
          hop = v.chain(p - total);
          dl = v.s_data + (p + len - 3 - v.dlo);
          fast = p + len + 1 <= v.dhi || total + hop + v.dhi >= p + len + 1;
          state = fast ? kWalk : kSlowWalk;
          if (run != 0 && hop == 1) state = kStretch;                  // inside a run: the closed form, right away
          age = 0;
        }
      }
      else if (state == kIdle && !exhausted)
      {
        // The tile is gone through three times: positions whose first two hops are short (a dense class: a long
        // chain) first, the bulk last -- so that the longest walks do not start when the tile is nearly done.
        const uint32_t idx = base + (uint32_t)__popc(idle & ((1u << lane) - 1));
        const uint32_t pass = idx >= 2 * tlen ? 2u : (idx >= tlen ? 1u : 0u);
        p = t0 + idx - pass * tlen;
        if (idx >= n_pass * tlen) exhausted = true;
        else
        {
          const uint32_t own = p == tw_pos ? tw_own : lds_u16(v.s_pe + 2 * (p - v.clo));
          uint32_t first = own != 0 ? v.chain(p) : 0;                // smallz4.h:190 (absolute slot)
          if (g.shift == 0) first = own;
          const uint32_t nh = first != 0 ? v.chain(p - first) : 0;
          const uint32_t two = nh != 0 ? first + nh : 0xffffffffu;
          const uint32_t cls = n_pass == 1 ? 0u : (two < dense_a ? 0u : (two < dense_b ? 1u : 2u));
          if (own != 0 && cls == pass)
          {
            state = kWalk; len = 1; dist = 0; total = 0; budget = g.max_chain; tail = 0; limit = kWindow;
            hop = first;
            run = 0;
            if (g.shift == 0) { run = run_fwd[p]; if (run < kMinMatch) run = 0; }
            cbase = v.s_pe + 2 * (p - v.shift - v.clo);
            dl = v.s_data + (p + len - v.dlo);
            fast = false;
            // The first candidate is accepted unseen (smallz4.h:224-233 compare nothing while the best
            // length is 1), so it is handled right here instead of costing the lane a round -- unless
            // it opens a stretch of a run, which walk_stretch does without touching the bytes.
            if (hop != 0)
            {
              if (!(run != 0 && nh == 1))
              {
                total = hop; hop = nh;
                (void)try_candidate(v, p, p - total, stop, len, tail, runs);
                dist = total;
                dl = v.s_data + (p + len - 3 - v.dlo);
                fast = len >= 4 && (p + len + 1 <= v.dhi || total + hop + v.dhi >= p + len + 1);
                if (--budget == 0 || p + len + 1 > stop)
                {
                  mlen[p] = len;
                  mdist[p] = (uint16_t)dist;
                  state = kIdle;
                }
              }
            }
            if (state == kWalk && !fast) state = kSlowWalk;
          }
        }
      }
      idle = __ballot_sync(0xffffffffu, state == kIdle && !exhausted);
    }
    if (!__any_sync(0xffffffffu, state != kIdle))
    {
      // nobody is walking a chain: done when the tile has no positions left for this warp
      if (idle == 0) break;
      continue;
    }

    // ---- fast hops (smallz4.h:192-233, the rejecting path): follow the chain while the bytes a longer match
    // would need first differ.  A lane that meets anything else parks in a state for the slow part.
    // Written without branches around the loads: every lane executes the same 17 instructions per candidate.
    fast_hops_loop(v, fast_hops, state, total, hop, run, tail, cbase, dl, fast_lanes, limit);
#ifdef SZ4_SEARCH_STATS
    { const uint32_t m = __ballot_sync(0xffffffffu, state >= kCheck); if (lane == 0) { atomicAdd(&g_stats[2], 1ull); atomicAdd(&g_stats[3], (unsigned long long)__popc(m)); if (m) atomicAdd(&g_stats[5], 1ull); } }
#endif

    // ---- slow part: candidates that passed the first byte, stretches, finished walks
    if (state >= kCheck)
    {
      bool rejected = false;
      if (state == kSlowWalk)
      {
        // one hop of a lane whose match reaches beyond the staged bytes: every candidate gets the closer look,
        // and any 1-hop inside a run goes to the closed form
        const uint32_t tot2 = total + hop;
        if (hop == 0 || tot2 > limit) state = kFinish;               // smallz4.h:192,196
        else
        {
          total = tot2;
          hop = lds_u16(cbase - 2 * tot2);
          state = (run != 0 && hop == 1) ? kStretch : kCheck;
          // p heads `run` equal bytes and already has a match at least that long: a candidate can only be longer
          // if exactly as many of that byte follow it (fewer: its run ends first; more: p's does) -- one read
          // instead of a comparison
          rejected = state == kCheck && run != 0 && len >= run && run_fwd[p - total] != run;
        }
      }
      bool finish = state == kFinish;
      if (!finish)
      {
        const uint32_t len_in = len;
        // (a candidate that is worth a closer look and opens a stretch of its run goes to the closed form as well:
        // inside p's own run every candidate passes the filter)
        if (rejected) { }                                             // cannot be longer: on to the next candidate
        else if (state == kStretch || (run != 0 && hop == 1))
        {
          finish = walk_stretch(v, run_fwd, ones_back, p, stop, run, total, hop, len, dist, budget, tail, limit);
          // The chain goes on in the previous run of that byte (eight or more of it).  If it lands inside one -- the
          // landing place's own entry is 1 again -- that run is the next stretch: taken right here instead of after
          // another eight hops of the fast loop.  (A landing place with any other entry is a plain candidate.)
#pragma unroll 1
          for (uint32_t more = 0; more < kStretchesPerVisit && !finish; more++)
          {
            const uint32_t tot2 = total + hop;
            if (hop == 0 || tot2 > limit) break;                       // the walk ends: the fast loop sees it
            const uint32_t hop2 = lds_u16(cbase - 2 * tot2);
            if (hop2 != 1) break;
            total = tot2; hop = hop2;
            finish = walk_stretch(v, run_fwd, ones_back, p, stop, run, total, hop, len, dist, budget, tail, limit);
          }
        }
        else if (try_candidate(v, p, p - total, stop, len, tail, runs))
        {
          dist = total;
          if (--budget == 0) finish = true;
        }
        // p heads a run longer than the window and has a match as long as the run: no candidate can be longer (one
        // in another run of that byte has fewer of them, or the two runs would be one; one in p's own run has more,
        // so p's run ends first).  Without this, every position of a zero page behind binary data walks thousands
        // of four-zero candidates through this slow part.
        if (run > kWindow && len >= run) finish = true;
        if (len != len_in)
        {
          if (p + len + 1 > stop) finish = true;                     // smallz4.h:205: nothing longer fits
          dl = v.s_data + (p + len - 3 - v.dlo);
        }
        // the filter of the fast loop reads the candidate's bytes q+len-3 .. q+len from the staged range: fine when
        // p's are staged, and otherwise for every candidate far enough back (the next one is total + hop back)
        fast = len >= 4 && (p + len + 1 <= v.dhi || total + hop + v.dhi >= p + len + 1);
        state = fast ? kWalk : kSlowWalk;
      }
      if (finish)
      {
        mlen[p] = len;
        mdist[p] = (uint16_t)dist;
        state = kIdle;
      }
    }
    // A walk that has been going on for many rounds is one of the few very long ones (an 8-byte class with thousands of
    // members in the window): the tile would end with it, most lanes idle.  It is handed to k_long, which takes the
    // class 32 members at a time from the sorted array.  (Walks inside byte runs stay: they advance by whole runs.)
    // The same goes for the walks that are left when the tile's queue is empty and the warp cannot fill a third of its
    // lanes any more: every further round would cost the warp as much as a full one.
    bool hand_over = false;
    if (long_list != nullptr)
    {
      const bool movable = (state == kWalk || state == kSlowWalk) && run == 0;
      const bool drained = __any_sync(0xffffffffu, exhausted);
      const uint32_t walking = __ballot_sync(0xffffffffu, state != kIdle);
      hand_over = movable && (++age > long_age || (drained && (uint32_t)__popc(walking) < tail_lanes));
    }
    if (hand_over)
    {
      const uint32_t slot = atomicAdd(long_count, 1u);
      if (slot < long_cap)
      {
        LongWalk w; w.p = p; w.len = len; w.dist = dist; w.total = total; w.budget = budget;
        long_list[slot] = w;
        state = kIdle;
      }
      else age = 0;                                                  // the list is full: keep walking here
    }
    idle = __ballot_sync(0xffffffffu, state == kIdle && !exhausted);
  }
#ifdef SZ4_TILE_STATS
  __syncthreads();
  if (threadIdx.x == 0 && tile < (1u << 16)) { g_tile_us[tile] = (uint32_t)((tile_clock() - tile_t0) / 1000); g_tile_t0[tile] = (uint32_t)(tile_t0 / 1000); }
#endif
}

// ---------------------------------------------------------------------------------------------
// k_long: the rest of the few very long walks, one warp each.  The members of p's 8-byte class are the elements in
// front of p in the sorted arrays of sz4_lsd.cuh (same key, positions ascending), so the chain p would follow hop by
// hop is read 32 members at a time, coalesced; the filter (the four bytes a longer match has to reproduce first,
// smallz4.h:224-225) runs on all 32 at once and only a candidate that passes it gets the closer look.  Same candidates
// in the same order as the walk: same result.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 4)
k_long(const uint8_t* data, const uint64_t* skey, const uint32_t* spos, const uint32_t* rank, const LongWalk* list,
       const uint32_t* count, uint32_t cap, const uint32_t* run_fwd, uint32_t* mlen, uint16_t* mdist, uint32_t region_elems, Geom g)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
  const uint32_t n = min(*count, cap);
  GlobalView v; v.g_data = data;
  const uint32_t a_min = g.first_ins + 8;                            // the first anchor whose position is in the batch
  // the next walk's head (its entry, its rank, its key: three dependent trips to L2) is fetched while this one is walked
  uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  LongWalk nw; nw.p = 0; nw.len = 0; nw.dist = 0; nw.total = 0; nw.budget = 0;
  uint32_t nr = 0; uint64_t nkey = 0;
  if (e < n) { nw = list[e]; nr = rank[nw.p]; nkey = skey[nr]; }
  for (; e < n; e += nwarps)
  {
    const LongWalk w = nw;
    const uint32_t r = nr;
    const uint64_t key_p = nkey;
    if (e + nwarps < n) { nw = list[e + nwarps]; nr = rank[nw.p]; nkey = skey[nr]; }
    const uint32_t p = w.p, a = p + 8;
    uint32_t len = w.len, dist = w.dist, budget = w.budget;
    const uint32_t stop = block_end(g, (p - g.halo) / g.block_size) - kEndLiterals;
    const uint32_t limit = chain_limit(g, data, p, v.word_at(p));
    uint32_t tail = v.word_at(p + len - 3);
    const uint32_t lo = r / region_elems * region_elems;             // the chunk's first element
    // The candidates are at r-1, r-2, ...; the ones the lanes' walk has seen (up to w.total back) are passed over.
    // kLongWide x 32 candidates per step: their loads are in flight together (a step is two dependent trips to L2)
    uint32_t k = 0;
    for (bool done = false; !done; )
    {
      uint32_t a2[kLongWide]; bool valid[kLongWide], pass[kLongWide];
#pragma unroll
      for (uint32_t u = 0; u < kLongWide; u++)
      {
        const uint32_t off = k + u * 32 + lane;
        const bool in = r >= lo + 1 + off;
        const uint32_t idx = in ? r - 1 - off : r;
        a2[u] = spos[idx];
        // (the sorted arrays start with a few anchors whose position lies in front of the batch -- sz4_lsd.cuh leaves them
        // out of the tables the same way: the chain ends there)
        valid[u] = in && skey[idx] == key_p && a2[u] >= a_min;
      }
#pragma unroll
      for (uint32_t u = 0; u < kLongWide; u++)
      {
        valid[u] = valid[u] && a - a2[u] <= limit;                   // a member of the chain (smallz4.h:192-197)
        pass[u] = valid[u] && a - a2[u] > w.total && v.word_at(a2[u] - 8 + len - 3) == tail;
      }
      uint32_t advance = kLongWide * 32;
#pragma unroll
      for (uint32_t u = 0; u < kLongWide; u++)
      {
        if (advance != kLongWide * 32 || done) continue;             // (uniform: an earlier group has decided)
        const uint32_t inv = __ballot_sync(0xffffffffu, !valid[u]);
        uint32_t m = __ballot_sync(0xffffffffu, pass[u]);
        if (inv != 0) m &= (1u << (__ffs((int)inv) - 1)) - 1;        // nothing behind the chain's end
        if (m == 0) { if (inv != 0) done = true; continue; }
        const int l = __ffs((int)m) - 1;
        const uint32_t qa = __shfl_sync(0xffffffffu, a2[u], l);
        if (try_candidate(v, p, qa - 8, stop, len, tail, run_fwd))
        {
          dist = a - qa;
          if (--budget == 0 || p + len + 1 > stop) done = true;      // smallz4.h:250, 205
        }
        advance = u * 32 + (uint32_t)l + 1;                          // go on behind that candidate (with the new length)
      }
      k += advance;
    }
    if (lane == 0) { mlen[p] = len; mdist[p] = (uint16_t)dist; }
  }
}

// ---------------------------------------------------------------------------------------------
// Long runs (smallz4.h:632-643).  Once a position has a match {distance 1, length > 65299} the
// reference copies {length-1, 1} to the following positions without inserting them into the
// chains, until the length drops to 65299.  DESIGN.md "Q-run" shows that (without a dictionary)
// the searches of all other positions are unaffected, except the first position after the skipped
// stretch, whose predecessor becomes the seed.  Levels >= 7 fix this up in parallel here;
// levels 1..6 do it inside the sequential greedy/lazy filter below.
// ---------------------------------------------------------------------------------------------
struct Seed { uint32_t pos, len; };

__device__ __forceinline__ bool seed_like(const uint32_t* mlen, const uint16_t* mdist, uint32_t p)
{
  return mdist[p] == 1 && mlen[p] > kSameLetter;
}

__global__ void __launch_bounds__(256)
k_seed_detect(const uint8_t* data, const uint32_t* mlen, const uint16_t* mdist, Seed* seeds, uint32_t* n_seeds,
              uint32_t max_seeds, Geom g)
{
  uint32_t p = g.halo + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.n_total) return;
  if (!seed_like(mlen, mdist, p)) return;
  const uint32_t j = (p - g.halo) / g.block_size;
  const uint32_t b = block_begin(g, j);
  // a true seed is the first seed-like position of its run inside its block
  const uint8_t c = data[p];
  for (uint32_t q = p; q > b; )
  {
    q--;
    if (data[q] != c) break;
    if (seed_like(mlen, mdist, q)) return;
  }
  uint32_t k = atomicAdd(n_seeds, 1u);
  if (k < max_seeds) { seeds[k].pos = p; seeds[k].len = mlen[p]; }
}

// one CTA per seed: rewrite the skipped stretch and the position behind it
__global__ void __launch_bounds__(256)
k_seed_fix(uint32_t* mlen, uint16_t* mdist, const Seed* seeds, const uint32_t* n_seeds, Geom g)
{
  if (blockIdx.x >= *n_seeds) return;
  const uint32_t s = seeds[blockIdx.x].pos, l = seeds[blockIdx.x].len;
  const uint32_t count = l - kSameLetter;                       // skipped positions s+1 .. s+count
  for (uint32_t k = 1 + threadIdx.x; k <= count; k += blockDim.x)
  {
    mlen[s + k] = l - k;
    mdist[s + k] = 1;
  }
  if (threadIdx.x == 0)
  {
    const uint32_t e1 = s + count + 1;                          // first position that is inserted again
    const uint32_t gap = e1 - s;
    if (gap > kWindow) { mlen[e1] = 0; mdist[e1] = 0; }         // smallz4.h:668: predecessor too far
    else if (mdist[e1] == 1) mdist[e1] = (uint16_t)gap;         // nearest candidate is the seed, not e1-1
  }
}

// ---------------------------------------------------------------------------------------------
// Levels 1..6 (smallz4.h:606-612, 727-743): greedy / lazy levels search only some positions.
// k_search has produced the match of every eligible position; the passes below replay the reference's
// skipMatches / lazyEvaluation state machine and the long-run shortcut, and clear the matches the
// reference would not have looked for -- in segments, see k_greedy_spec / k_greedy_join / k_greedy_apply.
// ---------------------------------------------------------------------------------------------
#ifndef SZ4_GREEDY_SEG
#define SZ4_GREEDY_SEG 32768
#endif
#ifndef SZ4_GREEDY_WARM
#define SZ4_GREEDY_WARM 2048
#endif
enum : uint32_t
{
  kGreedyChunk = 8,                 // windows of 32 positions fetched ahead together
  kGreedySeg   = SZ4_GREEDY_SEG,    // positions per segment (a multiple of 256)
  kGreedyWarm  = SZ4_GREEDY_WARM    // positions a speculative walk starts in front of its segment
};

struct GreedyBlock                  // what the walk needs to know about its block
{
  uint32_t b, s_end, tw_pos, tw_own; bool tw;
};
__device__ __forceinline__ GreedyBlock greedy_block(const uint32_t* saved_pe, const Geom& g, uint32_t j)
{
  GreedyBlock k;
  k.b = block_begin(g, j); k.s_end = search_end(g, j);
  k.tw_pos = block_end(g, j) - kEndNoMatch;                      // the only position whose own entry is in saved_pe
  k.tw = g.shift != 0 && block_len(g, j) >= kEndNoMatch && is_twice_inserted(g, k.tw_pos);   // (pure tables are never zeroed)
  k.tw_own = k.tw ? saved_pe[(k.tw_pos + kEndNoMatch - g.halo) / g.block_size] : 0;
  return k;
}

// The reference's skipMatches / lazyEvaluation state machine (smallz4.h:727-743) and its long-run shortcut
// (smallz4.h:632-643) from position `at` on, entered with nothing left to skip.  It runs until it is again in
// that state at or behind `until` (or the block's positions are used up) and returns where.  kWrite = false
// only follows the decisions; kWrite = true also clears the matches the reference would not have looked for.
// The reference walks position by position; here a window of 32 positions is looked at at once and the
// state machine advances from event to event (a searched position, or a batch of skipped ones).  The windows
// come in chunks of eight: while one chunk is worked on out of shared memory, the loads of the next one are in
// flight (a window takes far less time than a DRAM round trip; nothing the loop writes is read again, except
// behind a jump over a long run, where the chunks are fetched anew).
template <bool kWrite>
__device__ __forceinline__ uint32_t greedy_walk(const GreedyBlock& k, const uint16_t* pe, uint32_t* mlen, uint16_t* mdist, const Geom& g,
                                                uint32_t at, uint32_t until, uint32_t* q_own, uint32_t* q_len, uint32_t* q_dist)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t b = k.b, s_end = k.s_end;
  uint32_t skip = 0;            // skipMatches
  bool peek = false;            // lazyEvaluation
  uint32_t seed = 0xffffffffu;  // last position in front of a stretch the long-run shortcut skipped
  uint32_t n_own[kGreedyChunk], n_len[kGreedyChunk], n_dist[kGreedyChunk];
  auto fetch = [&](uint32_t cb)                                  // chunk that starts at position cb -> registers
  {
#pragma unroll
    for (uint32_t i = 0; i < kGreedyChunk; i++)
    {
      const uint32_t q = cb + i * 32 + lane;
      n_own[i] = 0; n_len[i] = 0; n_dist[i] = 0;
      if (q < s_end)
      {
        n_len[i] = mlen[q]; n_dist[i] = mdist[q];
        // pe == nullptr (jump tables): searched iff the position has an exact predecessor iff k_search wrote a length
        n_own[i] = pe == nullptr ? n_len[i] : ((k.tw && q == k.tw_pos) ? k.tw_own : pe[q]);
      }
    }
  };
  auto commit = [&]()
  {
#pragma unroll
    for (uint32_t i = 0; i < kGreedyChunk; i++) { q_own[i * 32 + lane] = n_own[i]; q_len[i * 32 + lane] = n_len[i]; q_dist[i * 32 + lane] = n_dist[i]; }
  };
  if (at >= s_end) return s_end;
  uint32_t chunk = b + (at - b) / (kGreedyChunk * 32) * (kGreedyChunk * 32);   // first position of the chunk in shared memory
  fetch(chunk); commit(); fetch(chunk + kGreedyChunk * 32);
  while (at < s_end)
  {
    if (skip == 0 && at >= until) return at;
    const uint32_t w = b + ((at - b) & ~31u);
    const uint32_t p = w + lane;
    if (w - chunk >= kGreedyChunk * 32)
    {
      const uint32_t cb = b + (w - b) / (kGreedyChunk * 32) * (kGreedyChunk * 32);
      if (cb != chunk + kGreedyChunk * 32) fetch(cb);            // a jump: what is in flight is not what comes next
      chunk = cb;
      commit();
      fetch(chunk + kGreedyChunk * 32);
    }
    const uint32_t qi = (w - chunk) + lane;
    uint32_t own = q_own[qi], fl = q_len[qi], fd = q_dist[qi];
    if (seed != 0xffffffffu && p == at && g.shift == 0 && p < s_end)
    {
      // first position behind a skipped stretch: its predecessor in every chain is the seed
      const uint32_t gap = p - seed;
      if (gap > kWindow) { own = 0; if (kWrite) { mlen[p] = 0; mdist[p] = 0; } }
      else if (fd == 1) { fd = gap; if (kWrite) mdist[p] = (uint16_t)gap; }
    }
    seed = 0xffffffffu;
    const uint32_t eligible = __ballot_sync(0xffffffffu, own != 0);
    uint32_t o = at - w;
    bool jumped = false;
    while (o < 32)
    {
      if (skip == 0 && w + o >= until) return w + o;
      const uint32_t rem = eligible & (0xffffffffu << o);
      if (rem == 0) break;
      if (skip > 0 && !peek)
      {
        // smallz4.h:727-731: the next `skip` eligible positions are not searched
        const uint32_t c = (uint32_t)__popc(rem), t = min(c, skip);
        const uint32_t rank = (uint32_t)__popc(rem & ((2u << lane) - 1));      // 1-based among the remaining eligible lanes
        const bool mine = ((rem >> lane) & 1u) && rank <= t;
        if (kWrite && mine) { mlen[p] = 0; mdist[p] = 0; }
        skip -= t;
        const uint32_t last = __ballot_sync(0xffffffffu, mine && rank == t);
        o = (uint32_t)__ffs((int)last);                                        // lane of the t-th one, plus one
        if (t == c) { if (skip != 0) break; continue; }                        // (nothing left to skip: may be where to stop)
        continue;
      }
      // a searched position: the first remaining eligible one
      const int l = __ffs((int)rem) - 1;
      if (skip > 0) { skip--; peek = false; }                                   // the lazy peek
      const uint32_t L = __shfl_sync(0xffffffffu, fl, l);
      const uint32_t D = __shfl_sync(0xffffffffu, fd, l);
      if (L != 1) { peek = (skip == 0); skip = L; }                            // smallz4.h:739-743
      o = (uint32_t)l + 1;
      if (D == 1 && L > kSameLetter)
      {
        // smallz4.h:632-643: the following positions copy {length-1, 1} and are not inserted
        const uint32_t s = w + (uint32_t)l, count = L - kSameLetter;
        if (kWrite) for (uint32_t i = 1 + lane; i <= count; i += 32) { mlen[s + i] = L - i; mdist[s + i] = 1; }
        seed = s;
        at = s + count + 1;
        jumped = true;
        break;
      }
    }
    if (!jumped) at = w + 32;
    __syncwarp();
  }
  return s_end;
}

// Levels 1..6 in three steps.  The state of the machine is "eligible positions still to skip", and it is zero again
// behind every skipped stretch; two walks that are in that state in front of the same position stay together.  So:
// k_greedy_spec: every segment of kGreedySeg positions is walked (decisions only) from kGreedyWarm positions in front
//                of it; `entry` = where the walk first has nothing to skip inside the segment, `leave` = the same for
//                the next segment.
// k_greedy_join: one warp per block checks entry[k] == leave[k-1] from the first segment on (the first one starts in
//                the true state) and walks again, from the true entry, wherever that fails.
// k_greedy_apply: every segment is walked once more from its true entry, this time clearing the matches.
struct GreedySeg { uint32_t entry, leave; };

__global__ void __launch_bounds__(128)
k_greedy_spec(const uint16_t* pe, const uint32_t* saved_pe, uint32_t* mlen, uint16_t* mdist, GreedySeg* segs, uint32_t segs_per_block, Geom g)
{
  __shared__ uint32_t q_all[4][3][kGreedyChunk * 32];
  const uint32_t wi = threadIdx.x >> 5;
  const uint32_t id = blockIdx.x * 4 + wi;
  const uint32_t j = id / segs_per_block, sk = id % segs_per_block;
  if (j >= g.n_blocks) return;
  const GreedyBlock k = greedy_block(saved_pe, g, j);
  const uint32_t s0 = k.b + sk * kGreedySeg;
  GreedySeg r; r.entry = k.s_end; r.leave = k.s_end;
  if (s0 < k.s_end)
  {
    const uint32_t from = sk == 0 ? k.b : s0 - kGreedyWarm;      // (segments are longer than the warm-up)
    r.entry = sk == 0 ? k.b : greedy_walk<false>(k, pe, mlen, mdist, g, from, s0, q_all[wi][0], q_all[wi][1], q_all[wi][2]);
    r.leave = greedy_walk<false>(k, pe, mlen, mdist, g, r.entry, s0 + kGreedySeg, q_all[wi][0], q_all[wi][1], q_all[wi][2]);
  }
  if ((threadIdx.x & 31) == 0) segs[id] = r;
}

__global__ void __launch_bounds__(128)
k_greedy_join(const uint16_t* pe, const uint32_t* saved_pe, uint32_t* mlen, uint16_t* mdist, GreedySeg* segs, uint32_t segs_per_block,
              uint32_t* redo_count, Geom g)
{
  __shared__ uint32_t q_all[4][3][kGreedyChunk * 32];
  const uint32_t wi = threadIdx.x >> 5;
  const uint32_t j = blockIdx.x * 4 + wi;
  if (j >= g.n_blocks) return;
  const GreedyBlock k = greedy_block(saved_pe, g, j);
  uint32_t at = k.b;                                             // true entry of the segment at hand
  for (uint32_t sk = 0; sk < segs_per_block; sk++)
  {
    const uint32_t id = j * segs_per_block + sk;
    const uint32_t s0 = k.b + sk * kGreedySeg;
    if (s0 >= k.s_end) break;
    GreedySeg r = segs[id];
    uint32_t leave;
    if (at >= s0 + kGreedySeg) leave = at;                       // the previous walk went right across this segment
    else if (at == r.entry) leave = r.leave;
    else
    {
      leave = greedy_walk<false>(k, pe, mlen, mdist, g, at, s0 + kGreedySeg, q_all[wi][0], q_all[wi][1], q_all[wi][2]);
      if ((threadIdx.x & 31) == 0) atomicAdd(redo_count, 1u);
    }
    if ((threadIdx.x & 31) == 0) { r.entry = at; r.leave = leave; segs[id] = r; }
    at = leave;
  }
}

__global__ void __launch_bounds__(128)
k_greedy_apply(const uint16_t* pe, const uint32_t* saved_pe, uint32_t* mlen, uint16_t* mdist, const GreedySeg* segs, uint32_t segs_per_block, Geom g)
{
  __shared__ uint32_t q_all[4][3][kGreedyChunk * 32];
  const uint32_t wi = threadIdx.x >> 5;
  const uint32_t id = blockIdx.x * 4 + wi;
  const uint32_t j = id / segs_per_block, sk = id % segs_per_block;
  if (j >= g.n_blocks) return;
  const GreedyBlock k = greedy_block(saved_pe, g, j);
  const uint32_t s0 = k.b + sk * kGreedySeg;
  if (s0 >= k.s_end) return;
  const GreedySeg r = segs[id];
  if (r.entry >= s0 + kGreedySeg || r.entry >= k.s_end) return;  // nothing of this segment is looked at
  (void)greedy_walk<true>(k, pe, mlen, mdist, g, r.entry, s0 + kGreedySeg, q_all[wi][0], q_all[wi][1], q_all[wi][2]);
}

}  // namespace sz4
