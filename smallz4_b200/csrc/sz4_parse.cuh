// sz4_parse.cuh -- phase 3a: backward cost DP (smallz4.h:376 estimateCosts) and the walk along the
// chosen parse (the control flow of smallz4.h:259 selectBestMatches), one warp per LZ4 block.
//
// estimateCosts, for i from n-6 down to 0:
//     literal:  cost[i+1] + 1 (+1 when the run of literal decisions behind i reaches 15, 270, 525, ...)
//     match L in 4..len[i]:  cost[i+L] + 3 + e(L),  e(L) = 0 (L<=18) else 1 + (L-19)/255
//     smallest cost wins; among equal costs the longest match, and a match beats the literal.
//     len >= 65299 at distance 1 is taken unpriced (smallz4.h:410-416).
// The recurrence is sequential in i, but a candidate L >= 4 only needs costs that are at least four
// positions old.  The warp therefore works on groups of 32 positions:
//   far phase   (lane = position): candidates that end two or more groups ahead.  They are
//               grouped by e(L) into classes of 255 consecutive lengths; the minimum of a class is a
//               range-minimum query answered from three sparse-table levels (windows of 32/64/128
//               costs, ties resolved towards the larger index) that the warp maintains as it goes.
//               A match of length 60 000 costs ~470 table reads instead of 60 000 additions.
//               Candidates that end in the next group up are answered from five small table levels
//               (windows of 1..16 costs) kept in shared memory.
//   near phase  (32 sequential steps): each lane keeps the best candidate of ITS position in two
//               registers and updates it whenever a cost above it is decided (4 instructions, all
//               lanes at once); the decision for position l then only needs a shuffle from lane l
//               and the scalar recurrence cost[l] = min(cost[l+1] + literal, best candidate).
// The most recent kDpRing (512) costs and table entries stay in shared-memory rings -- small enough for eleven
// single-warp CTAs per SM, which is what bounds the throughput of k_dp_spec; older ones come from L2/HBM.
#pragma once
#include "sz4_device.cuh"
#include "sz4_sort.cuh"

namespace sz4
{
struct DpScratch
{
  uint32_t* cost;   // cost[i]                                   (indexed by batch position)
  uint32_t* st5;    // (min cost over [i, i+31])  << 8 | (255 - offset of the LAST minimum)
  uint32_t* st6;    // same over [i, i+63]
  uint32_t* st7;    // same over [i, i+127]
};

__device__ __forceinline__ uint32_t match_extra(uint32_t len) { return len < 19 ? 0 : 1 + (len - 19) / 255; }

// candidate = (cost << 32) | length ; better = smaller cost, then larger length
__device__ __forceinline__ void take_better(uint32_t& best_cost, uint32_t& best_len, uint32_t c, uint32_t l)
{
  if (c < best_cost || (c == best_cost && l > best_len)) { best_cost = c; best_len = l; }
}

// Tunables of the segment-parallel DP (overridable for the emulated tests, which use small blocks)
#ifndef SZ4_DP_RING
#define SZ4_DP_RING 512
#endif
#ifndef SZ4_DP_SEG
#define SZ4_DP_SEG 32768
#endif
#ifndef SZ4_DP_WARM
#define SZ4_DP_WARM 2048
#endif
#ifndef SZ4_DP_SLACK
#define SZ4_DP_SLACK 512
#endif
enum : uint32_t
{
  kDpRing  = SZ4_DP_RING,    // positions whose costs / tables stay in shared memory
  kDpSeg   = SZ4_DP_SEG,     // nominal segment length
  kDpWarm  = SZ4_DP_WARM,    // positions a segment starts to the right of its own range
  kDpSlack = SZ4_DP_SLACK,   // a boundary is only placed where no match from the left reaches further than this
  kDpOvl   = kDpWarm + 64,   // entries of a segment's warm-up overlay
  kDpURing = 512,
  kDpChunk = 8,              // groups of 32 positions fetched ahead together
  kDpSmem  = 4 * kDpRing * 4 + 5 * 48 * 4 + kDpURing * 16 + 32 * 8 + 4 * kDpChunk * 32 * 4,
  kDpSmemSpec = kDpSmem - 2 * kDpChunk * 32 * 4   // a first pass has nothing to compare with: no queue for the earlier results
};

// One segment of a block: own range [lo, hi) (block relative, multiples of 32; hi == block length for the
// top segment).  reach = largest i + len[i] over i < hi: positions left of hi never look beyond it.
// run_end != 0: hi lies inside a long run whose positions all take their match unpriced (smallz4.h:410-416) and all
// end at run_end; such matches are left out of reach (their cost is cost[run_end] plus a function of the length,
// so the segment prices itself relative to cost[run_end] := 0, see dp_segment).
// ub_start != 0: the warm-up starts in a stretch without any match candidate; ub_start is the literal counter
// (until_bump) there if the stretch runs to the end of the block (exact) or if the first candidate behind it
// is taken (a guess, checked like everything else by k_dp_verify).
struct DpTask
{
  uint32_t lo, hi, reach, aux;                                   // aux = run_end | ub_start << 24
  __device__ __forceinline__ uint32_t run_end() const { return aux & 0xffffffu; }
  __device__ __forceinline__ uint32_t ub_start() const { return aux >> 24; }
};
// literal counter after k literal decisions that follow a state with `first` decisions left until the next length byte
__device__ __forceinline__ uint32_t until_bump_after(uint32_t first, uint32_t k)
{
  return k < first ? first - k : 255u - ((k - first) % 255u);
}
// until_bump after the positions >= hi / >= lo were done; cum = offset of this segment's costs to the true ones
struct DpState { uint32_t ub_hi, ub_lo, redone, cum; };
struct DpOverlay { uint32_t* cost; uint32_t* st5; uint32_t* st6; uint32_t* st7; };   // warm-up results, kDpOvl each

// The warp's view of already priced positions: the most recent kDpRing of them live in shared-memory
// rings (29-cycle reads instead of an L2 round trip), everything older is read from HBM/L2 -- from the
// segment's private overlay for positions at or beyond `split` (its warm-up zone), else from the block arrays.
struct DpView
{
  const uint32_t* r_cost; const uint32_t* r_st5; const uint32_t* r_st6; const uint32_t* r_st7;   // rings
  DpScratch s;
  DpOverlay o;       // o.cost == nullptr: no overlay (everything is in the block arrays)
  uint32_t b;        // block start (batch position)
  uint32_t ring_hi;  // block-relative positions below this are in the rings
  uint32_t split;

  __device__ __forceinline__ uint32_t cost(uint32_t i) const
  {
    if (i < ring_hi) return r_cost[i & (kDpRing - 1)];
    if (o.cost != nullptr && i >= split) return __ldcg(o.cost + (i - split));
    return __ldcg(s.cost + b + i);
  }
  __device__ __forceinline__ uint32_t tab(uint32_t w, uint32_t i) const
  {
    if (i < ring_hi)
    {
      const uint32_t* r = w == 128 ? r_st7 : (w == 64 ? r_st6 : r_st5);
      return r[i & (kDpRing - 1)];
    }
    if (o.cost != nullptr && i >= split)
    {
      const uint32_t* t = w == 128 ? o.st7 : (w == 64 ? o.st6 : o.st5);
      return __ldcg(t + (i - split));
    }
    const uint32_t* t = w == 128 ? s.st7 : (w == 64 ? s.st6 : s.st5);
    return __ldcg(t + b + i);
  }
};

// minimum of cost[a..b] (block relative, a <= b, b-a < 255), ties -> largest index.
__device__ __forceinline__ void range_min(const DpView& v, uint32_t a, uint32_t b, uint32_t& out_cost, uint32_t& out_idx)
{
  const uint32_t span = b - a + 1;
  if (span < 32)
  {
    uint32_t bc = 0xffffffffu, bi = a;
    for (uint32_t j = a; j <= b; j++)
    {
      uint32_t c = v.cost(j);
      if (c <= bc) { bc = c; bi = j; }
    }
    out_cost = bc; out_idx = bi;
    return;
  }
  const uint32_t w = span >= 128 ? 128 : (span >= 64 ? 64 : 32);
  const uint32_t a2 = b + 1 - w;
  const uint32_t v1 = v.tab(w, a), v2 = v.tab(w, a2);
  const uint32_t c1 = v1 >> 8, i1 = a + (255 - (v1 & 255));
  const uint32_t c2 = v2 >> 8, i2 = a2 + (255 - (v2 & 255));
  if (c2 < c1 || (c2 == c1 && i2 > i1)) { out_cost = c2; out_idx = i2; }
  else { out_cost = c1; out_idx = i1; }
}

enum : uint32_t { kDpSmall = 5 * 48 * 4 };   // five small sparse-table levels over the 32 costs of the previous group

// minimum over the previous group's costs [xa..xb] (lane indices, xa <= xb), ties -> larger index.
// lvl[k][x] = min over x..x+2^k-1 of (cost << 5 | 31 - index), padded with 0xffffffff.
__device__ __forceinline__ void small_min(const uint32_t* lvl, uint32_t xa, uint32_t xb, uint32_t& out_cost, uint32_t& out_x)
{
  const uint32_t span = xb - xa + 1;
  const uint32_t k = min(4u, 31u - (uint32_t)__clz((int)span));   // span <= 32: two windows of 16 cover it
  const uint32_t* t = lvl + k * 48;
  const uint32_t r = min(t[xa], t[xb + 1 - (1u << k)]);
  out_cost = r >> 5; out_x = 31u - (r & 31u);
}

// The sequential part of one group: positions i0+31 .. i0, high to low (smallz4.h:389-471).
// Every lane keeps the best candidate of ITS position in (bc, bl_f) and updates it whenever the cost
// of a position above it is decided (no reduction).  Lane l's candidate is final once the cost of
// position l+4 is known, so its broadcast is issued four steps before it is used.
//   next_cost   cost[i+1]
//   until_bump  literal decisions left until the run of literals needs another length byte:
//               numLiterals reaches 15, 270, 525, ... (smallz4.h:398-404)
//   out[l]      = {cost, chosen length} of position i0 + l
template <bool kTop>
__device__ __forceinline__ void dp_steps(uint32_t lane, uint32_t i0, uint32_t last_priced, uint32_t wlen, uint32_t bc,
                                         uint32_t bl_f, uint32_t& next_cost, uint32_t& until_bump, uint2* out)
{
  uint32_t cq[4], lq[4];
#pragma unroll
  for (int32_t l = 31; l >= 28; l--)
  {
    cq[l & 3] = __shfl_sync(0xffffffffu, bc, l);
    lq[l & 3] = __shfl_sync(0xffffffffu, bl_f, l);
  }
  const bool writer = lane == 0;
#pragma unroll
  for (int32_t l = 31; l >= 0; l--)
  {
    const uint32_t cb = cq[l & 3];                     // best candidate of position l: cost (all ones = none)
    const uint32_t lb = lq[l & 3];                     // its length, bit 31 = taken unpriced (smallz4.h:410)
    uint32_t lowest = 0, choice = 1;                   // the last five positions cost nothing (smallz4.h:383-389)
    if (!kTop || i0 + (uint32_t)l <= last_priced)      // uniform
    {
      const bool bump = until_bump == 1;
      uint32_t lit = next_cost + 1;
      if (bump) lit++;
      if ((int32_t)lb < 0) lit = 0xffffffffu;          // unpriced long run: the literal is not considered
      const bool take = cb <= lit;                     // a match wins ties (smallz4.h:431)
      lowest = min(cb, lit);
      choice = take ? (lb & 0x7fffffffu) : 1u;
      until_bump = take ? 15u : (bump ? 255u : until_bump - 1);
      next_cost = lowest;
    }
    if (writer) out[l] = make_uint2(lowest, choice);
    {
      // this cost is a candidate for the positions below: length l - lane
      const uint32_t L = (uint32_t)l - lane;
      const uint32_t c = lowest + (L >= 19 ? 4u : 3u);
      if (L - 4 < wlen && c < bc) { bc = c; bl_f = L; }
    }
    if (l >= 4)
    {
      cq[l & 3] = __shfl_sync(0xffffffffu, bc, l - 4);
      lq[l & 3] = __shfl_sync(0xffffffffu, bl_f, l - 4);
    }
  }
}

// The DP over the positions [lo, start) of one block, high to low, where start = hi (resume: the state at
// hi is loaded from the block arrays, which hold the true values of the segment to the right) or
// start = min(hi + kDpWarm, n) (cold start: as if the block ended at `start`; the results for positions
// >= hi go to the overlay and only serve to let the costs settle before the own range begins).
// Own-range results go to the block arrays (s) and the final lengths to mfin.
__device__ __forceinline__ void dp_segment(const uint32_t* mlen, const uint16_t* mdist, uint32_t* mfin, DpScratch s, DpOverlay ovl,
                                           uint32_t b, uint32_t n, uint32_t lo, uint32_t hi, bool resume, uint32_t ub_resume,
                                           unsigned char* smem, uint32_t& ub_hi, uint32_t& ub_lo, const uint32_t* reach_before = nullptr,
                                           bool* stopped = nullptr, uint32_t run_end = 0, uint32_t ub_start = 0)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t start = resume ? hi : min(hi + (uint32_t)kDpWarm, n);
  const bool true_end = (start == n) && !resume;
  // a cold start prices the positions as if the block ended at `start` (smallz4.h:389: the last five are free,
  // matches end five bytes before the end); at the real end this is the reference's own rule
  // (a cold start inside a long run, run_end != 0, is exact instead: every position of the warm-up zone takes
  // its match to run_end unpriced, so its cost is a function of the distance to run_end alone)
  // (a cold start in a stretch without match candidates, ub_start != 0, takes over the literal counter the plan worked out)
  const uint32_t run_e = resume ? 0 : run_end;
  const uint32_t ub_cold = (resume || true_end || run_e != 0) ? 0 : ub_start;
  const uint32_t last_priced = (resume || run_e != 0 || ub_cold != 0) ? 0xffffffffu : start - (1 + kEndLiterals);
  const uint32_t cap_end = (resume || true_end) ? 0xffffffffu : start - kEndLiterals;
  bool bad = false;                                              // the run did not look as the plan assumed
  const uint32_t top_group = (start - 1) / 32;

  uint32_t* r_cost = (uint32_t*)smem;
  uint32_t* r_st5 = r_cost + kDpRing;
  uint32_t* r_st6 = r_st5 + kDpRing;
  uint32_t* r_st7 = r_st6 + kDpRing;
  uint32_t* lvl = r_st7 + kDpRing;                               // [5][48]
  uint4* u_ring = (uint4*)(lvl + 5 * 48);                        // per position: {position, match end, best of classes >= 2, its end}
  for (uint32_t k = lane; k < kDpURing; k += 32) u_ring[k] = make_uint4(0xffffffffu, 0, 0xffffffffu, 0);
  uint2* s_out = (uint2*)(u_ring + kDpURing);                    // {cost, chosen length} of the group's positions
  DpView v;
  v.r_cost = r_cost; v.r_st5 = r_st5; v.r_st6 = r_st6; v.r_st7 = r_st7; v.s = s; v.b = b;
  v.o = ovl; v.split = hi;
  if (resume) v.o.cost = nullptr;
  for (uint32_t k = lane; k < 5 * 48; k += 32) lvl[k] = 0xffffffffu;
  __syncwarp();

  uint32_t until_bump = ub_cold != 0 ? ub_cold : 15 - kEndLiterals;   // numLiterals starts at 5 (smallz4.h:387)
  uint32_t next_cost = 0;                                        // cost[i+1], uniform
  uint32_t prv = 0;                                              // cost of (group+1)*32 + lane
  uint32_t p5 = 0xffffffffu, p6a = 0xffffffffu, p6b = 0xffffffffu;   // st5 of group+1, st6 of group+1 / group+2
  bool have_prev = false;                                        // a group above this one exists
  uint32_t far_end = run_e != 0 ? run_e : 0xffffffffu, far_cost = 0;   // last end position of an unpriced long run and its cost
  uint32_t settled = 0;                                          // resume: positions in a row that reproduce the earlier result
  bool settled_match = false, early_stop = false;
  uint32_t settled_diff = 0;
  ub_hi = until_bump;
  if (resume)
  {
    // take over the state at position hi from the block arrays
    until_bump = ub_resume;
    const uint32_t i = hi + lane;
    prv = i < n ? __ldcg(s.cost + b + i) : 0;
    next_cost = __shfl_sync(0xffffffffu, prv, 0);
    if (i < n) { p5 = __ldcg(s.st5 + b + i); p6a = __ldcg(s.st6 + b + i); }
    if (i + 32 < n) p6b = __ldcg(s.st6 + b + i + 32);
    have_prev = true;
    uint32_t small = (prv << 5) | (31u - lane);
    lvl[lane] = small;
#pragma unroll
    for (uint32_t k = 1; k < 5; k++)
    {
      uint32_t tt = __shfl_down_sync(0xffffffffu, small, 1u << (k - 1));
      if (lane + (1u << (k - 1)) < 32) small = min(small, tt);
      lvl[k * 48 + lane] = small;
    }
    for (uint32_t y = hi + lane; y < hi + kDpRing && y < n; y += 32)
    {
      const uint32_t slot = y & (kDpRing - 1);
      r_cost[slot] = __ldcg(s.cost + b + y); r_st5[slot] = __ldcg(s.st5 + b + y);
      r_st6[slot] = __ldcg(s.st6 + b + y);  r_st7[slot] = __ldcg(s.st7 + b + y);
    }
    __syncwarp();
  }

  // Matches (and, for a redo, what the earlier pass left behind) are fetched a chunk of kDpChunk groups at a
  // time: while the warp works on one chunk out of shared memory, the loads of the chunk below it are in
  // flight into registers.  A cheap group (all literals, all long runs) takes far less time than a DRAM
  // round trip, so anything closer than that would make the memory latency the time per group.
  const bool compare = resume && reach_before != nullptr;
  uint32_t* q_m = (uint32_t*)(s_out + 32);                       // [kDpChunk][32] each, lane-private columns
  uint32_t* q_d = q_m + kDpChunk * 32;
  uint32_t* q_oc = q_d + kDpChunk * 32;
  uint32_t* q_ok = q_oc + kDpChunk * 32;
  uint32_t nM[kDpChunk], nD[kDpChunk], nOC[kDpChunk], nOK[kDpChunk];
  auto fetch_chunk = [&](int32_t c)
  {
#pragma unroll
    for (uint32_t k = 0; k < kDpChunk; k++)
    {
      nM[k] = 0; nD[k] = 0; nOC[k] = 0; nOK[k] = 0;
      const uint32_t i = ((uint32_t)c * kDpChunk + k) * 32 + lane;
      if (c >= 0 && i < start)
      {
        nM[k] = mlen[b + i]; nD[k] = mdist[b + i];
        if (compare) { nOC[k] = __ldcg(s.cost + b + i); nOK[k] = mfin[b + i]; }
      }
    }
  };
  auto commit_chunk = [&]()
  {
#pragma unroll
    for (uint32_t k = 0; k < kDpChunk; k++)
    {
      q_m[k * 32 + lane] = nM[k]; q_d[k * 32 + lane] = nD[k];
      if (compare) { q_oc[k * 32 + lane] = nOC[k]; q_ok[k * 32 + lane] = nOK[k]; }
    }
  };
  fetch_chunk((int32_t)(top_group / kDpChunk));
  commit_chunk();
  fetch_chunk((int32_t)(top_group / kDpChunk) - 1);

  for (int32_t grp = (int32_t)top_group; grp >= (int32_t)(lo / 32); grp--)
  {
    const uint32_t i0 = (uint32_t)grp * 32;
    const uint32_t i = i0 + lane;                                // this lane's position (block relative)
    const bool exists = i < start;
    const bool priced = exists && i <= last_priced;
    if (((uint32_t)grp & (kDpChunk - 1)) == kDpChunk - 1 && (uint32_t)grp != top_group)
    {
      commit_chunk();                                            // the chunk that starts here was fetched a chunk ago
      fetch_chunk(grp / (int32_t)kDpChunk - 1);
    }
    const uint32_t qi = ((uint32_t)grp & (kDpChunk - 1)) * 32 + lane;
    uint32_t M = priced ? q_m[qi] : 0;
    const uint32_t D = q_d[qi];
    const bool to_run_end = run_e != 0 && M >= kSameLetter && D == 1 && i + M == run_e;
    if (M != 0 && i + M > cap_end && !to_run_end)
    {
      if (run_e != 0 && M >= kSameLetter && D == 1) bad = true;  // an unpriced match to somewhere else: not covered by `reach`
      M = cap_end > i ? cap_end - i : 0;
      if (M < kMinMatch) M = 0;
    }
    if (run_e != 0 && i0 >= hi && !__all_sync(0xffffffffu, to_run_end)) bad = true;
    uint32_t old_cost = 0, old_keep = 0;
    if (compare) { old_cost = q_oc[qi]; old_keep = q_ok[qi]; }
    v.ring_hi = min(i0 + 32 + (uint32_t)kDpRing, resume ? 0xffffffffu : start);

    uint32_t cur, keep;
    if (i0 + 31 <= last_priced && __all_sync(0xffffffffu, M == 0))
    {
      // no position of the group has a match: 32 literal steps in closed form.  The next length byte is
      // due at step until_bump (smallz4.h:398-404), at most once within 32 steps.
      const uint32_t t = 32 - lane;                              // steps from the top of the group down to this lane
      cur = next_cost + t + (t >= until_bump ? 1u : 0u);
      keep = 1;
      next_cost += 32 + (until_bump <= 32 ? 1u : 0u);
      until_bump = until_bump <= 32 ? 255 - (32 - until_bump) : until_bump - 32;
    }
    else
    {
      // ------------------------------ parallel part (lane = position): best candidate among everything that is
      // already priced, i.e. lengths that end in the next group or beyond.  bc/bl = its cost / length.
      uint32_t bc = 0xffffffffu, bl = 0;
      bool forced = false;
      if (M >= kSameLetter && D == 1)
      {
        forced = true;                                             // smallz4.h:410-416
        // inside one long run every position's match ends at the same place: remember that cost
        const uint32_t end = i + M;
        if (end != far_end) { far_end = end; far_cost = to_run_end ? 0u : v.cost(end); }
        bc = far_cost + 1 + 2 + 1 + (M - 19) / 255;
        bl = M;
      }
      else if (M >= kMinMatch)
      {
        if (i + M >= i0 + 64)
        {
          // Lengths that end two groups ahead or further, by class of extra length bytes:
          //   class 1 = lengths 19..273 (3+1 bytes), class c = 19+255(c-1) .. 18+255c (3+c bytes).
          // class 1: only its part beyond the next group
          {
            const uint32_t lo = i0 + 64 - i, hi = min(M, 273u);   // lo is 33..64
            if (lo <= hi)
            {
              uint32_t c, at;
              range_min(v, i + lo, i + hi, c, at);
              take_better(bc, bl, c + 4, at - i);
            }
          }
          // classes >= 2.  The classes >= 3 of position i are the classes >= 2 of position i+255 whenever
          // both matches end at the same position (inside one long match or run they do), each one
          // extra byte more expensive -- so the best of "classes >= 2" is kept per position in a small
          // ring and long matches cost two table lookups instead of one per 255 bytes of length.
          uint32_t uc = 0xffffffffu, ul = 0;
          if (M >= 274)
          {
            uint32_t c, at;
            range_min(v, i + 274, i + min(M, 528u), c, at);
            take_better(uc, ul, c + 5, at - i);
            if (M >= 529)
            {
              const uint4 u = u_ring[(i + 255) & (kDpURing - 1)];
              if (u.x == i + 255 && u.y == i + M)
              {
                if (u.z != 0xffffffffu) take_better(uc, ul, u.z + 1, u.w - i);
              }
              else
              {
                uint32_t lo = 529;
                while (lo <= M)
                {
                  const uint32_t e = match_extra(lo);
                  uint32_t hi = 18 + 255 * e;
                  if (hi > M) hi = M;
                  range_min(v, i + lo, i + hi, c, at);
                  take_better(uc, ul, c + 3 + e, at - i);
                  lo = hi + 1;
                }
              }
            }
            take_better(bc, bl, uc, ul);
          }
          u_ring[i & (kDpURing - 1)] = make_uint4(i, i + M, uc, i + ul);
        }
        if (have_prev && lane + M >= 32)
        {
          // lengths 32+x-lane that end at lane x of the next group: x in [max(0,lane-28), min(31,lane+M-32)]
          const uint32_t x_hi = min(31u, lane + M - 32);
          const uint32_t x_lo = lane >= 28 ? lane - 28 : 0;
          if (x_lo <= x_hi)
          {
            uint32_t c, x;
            const uint32_t split = lane >= 13 ? lane - 13 : 0;     // first x whose length is >= 19
            if (split <= x_hi)
            {
              small_min(lvl, max(x_lo, split), x_hi, c, x);        // one extra length byte
              take_better(bc, bl, c + 4, 32 + x - lane);
            }
            if (split > x_lo)
            {
              small_min(lvl, x_lo, min(x_hi, split - 1), c, x);
              take_better(bc, bl, c + 3, 32 + x - lane);
            }
          }
        }
      }
      // lengths that end inside this group are priced on the fly below: window of valid lengths 4..M
      const uint32_t wlen = (!forced && M >= kMinMatch) ? M - 3 : 0;
      uint32_t bl_f = bl | (forced ? 0x80000000u : 0u);

      if (i0 + 31 <= last_priced && __all_sync(0xffffffffu, forced))
      {
        // every position of the group takes its long run unpriced (smallz4.h:410-416): the costs only depend
        // on costs >= 65299 positions ahead, nothing is sequential, and each match resets the literal counter
        cur = bc; keep = bl;
        next_cost = __shfl_sync(0xffffffffu, bc, 0);
        until_bump = 15;
      }
      else
      {
        // ------------------------------ sequential part: 32 positions, high to low (dp_steps above)
        if (i0 + 31 <= last_priced)
          dp_steps<false>(lane, i0, last_priced, wlen, bc, bl_f, next_cost, until_bump, s_out);
        else
          dp_steps<true>(lane, i0, last_priced, wlen, bc, bl_f, next_cost, until_bump, s_out);
        __syncwarp();
        const uint2 mine = s_out[lane];
        cur = mine.x; keep = mine.y;
      }

    }

    // ------------------------------ publish the group: costs, final lengths, sparse-table levels
    // st5: min over cost[i .. i+31] = suffix of this group from `lane` + prefix of the next group below `lane`
    uint32_t kc = (cur << 6) | (63u - lane);                     // index lane      (this group)
    uint32_t kp = (prv << 6) | (31u - lane);                     // index 32 + lane (next group)
    uint32_t small = (cur << 5) | (31u - lane);
    lvl[lane] = small;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      uint32_t tt = __shfl_down_sync(0xffffffffu, kc, d);
      if (lane + d < 32) kc = min(kc, tt);
      uint32_t u = __shfl_up_sync(0xffffffffu, kp, d);
      if (lane >= d) kp = min(kp, u);
    }
#pragma unroll
    for (uint32_t k = 1; k < 5; k++)
    {
      uint32_t tt = __shfl_down_sync(0xffffffffu, small, 1u << (k - 1));
      if (lane + (1u << (k - 1)) < 32) small = min(small, tt);
      lvl[k * 48 + lane] = small;
    }
    uint32_t kp_excl = __shfl_up_sync(0xffffffffu, kp, 1);
    if (lane == 0) kp_excl = 0xffffffffu;
    const uint32_t k5 = min(kc, kp_excl);
    const uint32_t off5 = (63u - (k5 & 63u)) - lane;             // 0..31
    const uint32_t v5 = ((k5 >> 6) << 8) | (255u - off5);
    // st6 = st5[i] (+) st5[i+32] ; st7 = st6[i] (+) st6[i+64]; the later window wins ties
    uint32_t v6 = v5;
    if (p5 != 0xffffffffu && (p5 >> 8) <= (v5 >> 8)) v6 = ((p5 >> 8) << 8) | ((p5 & 255u) - 32u);
    uint32_t v7 = v6;
    if (p6b != 0xffffffffu && (p6b >> 8) <= (v6 >> 8)) v7 = ((p6b >> 8) << 8) | ((p6b & 255u) - 64u);
    if (resume && reach_before != nullptr && i0 + 32 <= hi)
    {
      // A redo may stop once it provably reproduces the earlier (speculative) pass: over a span longer than
      // anything further left can look at, the costs differ from the earlier ones by one constant, the chosen
      // lengths are identical, and a match was chosen in the span (which also pins the literal counter).
      const uint32_t diff = cur - old_cost;
      const uint32_t d0 = __shfl_sync(0xffffffffu, diff, 0);
      // lanes (from the lowest position up) that reproduce the earlier pass with the offset of lane 0
      const uint32_t ok_lanes = __ballot_sync(0xffffffffu, diff == d0 && keep == old_keep);
      const uint32_t low = ok_lanes == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~ok_lanes) - 1;
      const uint32_t low_mask = low == 32 ? 0xffffffffu : ((1u << low) - 1);
      const bool match_here = (__ballot_sync(0xffffffffu, keep != 1) & low_mask) != 0;
      if (low == 32 && settled != 0 && d0 == settled_diff)
      {
        settled += 32;
        settled_match = settled_match || match_here;
      }
      else
      {
        // a new span starts in this group: its top is the first lane that does not fit
        settled = low; settled_diff = d0; settled_match = match_here;
      }
      // the span must cover everything positions further left can look at (reach_before = largest position + length left of this group)
      if (settled_match && settled >= 256 && __ldcg(reach_before + i0 / 32) <= i0 + settled && i0 >= lo + kDpRing + 256)
      {
        // everything from lo up to here stays as the earlier pass left it (its costs in its own frame)
        early_stop = true;
        break;
      }
    }
    if (exists)
    {
      if (i0 >= hi)
      {
        // warm-up zone: private overlay
        const uint32_t k = i - hi;
        ovl.cost[k] = cur; ovl.st5[k] = v5; ovl.st6[k] = v6; ovl.st7[k] = v7;
      }
      else
      {
        s.cost[b + i] = cur; s.st5[b + i] = v5; s.st6[b + i] = v6; s.st7[b + i] = v7;
        if (priced) mfin[b + i] = keep;
      }
    }
    if (i0 == hi) ub_hi = until_bump;                            // everything at or beyond hi is done
    const uint32_t slot = i & (kDpRing - 1);
    r_cost[slot] = cur; r_st5[slot] = v5; r_st6[slot] = v6; r_st7[slot] = v7;
    p6b = p6a; p6a = v6; p5 = v5;
    prv = cur;
    have_prev = true;
    __syncwarp();
  }
  if (stopped) *stopped = early_stop;
  if (!early_stop) ub_lo = until_bump;
  if (__any_sync(0xffffffffu, bad)) ub_hi = 0xfffffffeu;          // never equals a literal counter: the segment is redone
}

// ---------------------------------------------------------------------------------------------
// Segment-parallel DP.  The recurrence is sequential, but its state (cost differences over the
// span later positions can reach, and the literal-run counter) forgets a wrong start within a few
// dozen positions.  So a block is cut into segments where no match crosses (k_dp_plan), every
// segment is priced by its own warp starting kDpWarm positions to its right from a cold state
// (k_dp_spec), and k_dp_verify then walks the boundaries from the block end to its start: if the
// costs a segment assumed over the span [hi, reach] differ from its right neighbour's true ones
// only by a constant, and the literal counter agrees, its decisions are the reference's (the
// recurrence only uses differences); otherwise that segment is priced again from the true state.
// Exactness never depends on the guess -- only the speed does.
// ---------------------------------------------------------------------------------------------
// Per group of 32 positions: reach = max over its positions of (position + found match length), 0 if none;
// reach_nf = the same without the unpriced long-run matches (length >= 65299 at distance 1, smallz4.h:410);
// run_end = where those matches end if all 32 positions have one and they all end at the same place, else 0.
// One warp per 1024 positions; lane g ends up with the values of group g and they are stored coalesced.
__global__ void __launch_bounds__(256)
k_dp_group_reach(const uint32_t* mlen, const uint16_t* mdist, uint32_t* group_reach, uint32_t* group_reach_nf, uint32_t* group_run_end,
                 uint32_t* group_first, uint32_t* chunk_reach, uint32_t* chunk_reach_nf, uint32_t* chunk_first,
                 uint32_t groups_per_block, Geom g)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t chunks_per_block = (groups_per_block + 31) / 32;
  const uint32_t j = warp / chunks_per_block, c = warp % chunks_per_block;
  if (j >= g.n_blocks) return;
  const uint32_t b = block_begin(g, j), n = block_len(g, j);
  uint32_t mine = 0, mine_nf = 0, mine_run = 0, mine_first = 0xffffffffu;   // first = first position with a match candidate
#pragma unroll 4
  for (uint32_t k = 0; k < 32; k++)
  {
    const uint32_t i = (c * 32 + k) * 32 + lane;
    uint32_t r = 0, rn = 0, e = 0, f = 0xffffffffu;
    if (i < n)
    {
      const uint32_t M = mlen[b + i];
      if (M > 1) { r = i + M; if (M >= kSameLetter && mdist[b + i] == 1) e = r; else rn = r; }
      if (M >= kMinMatch) f = i;
    }
    f = __reduce_min_sync(0xffffffffu, f);
    r = __reduce_max_sync(0xffffffffu, r);
    rn = __reduce_max_sync(0xffffffffu, rn);
    const uint32_t e_hi = __reduce_max_sync(0xffffffffu, e), e_lo = __reduce_min_sync(0xffffffffu, e);
    if (lane == k) { mine = r; mine_nf = rn; mine_run = e_hi == e_lo ? e_hi : 0; mine_first = f; }
  }
  const uint32_t grp = c * 32 + lane;
  if (grp < groups_per_block)
  {
    const size_t at = (size_t)j * groups_per_block + grp;
    group_reach[at] = mine; group_reach_nf[at] = mine_nf; group_run_end[at] = mine_run; group_first[at] = mine_first;
  }
  // the same over the whole chunk of 32 groups, for the two-level scans of k_dp_chunk_scan
  const uint32_t cr = __reduce_max_sync(0xffffffffu, mine), crn = __reduce_max_sync(0xffffffffu, mine_nf);
  const uint32_t cf = __reduce_min_sync(0xffffffffu, mine_first);
  if (lane == 0)
  {
    const size_t at = (size_t)j * chunks_per_block + c;
    chunk_reach[at] = cr; chunk_reach_nf[at] = crn; chunk_first[at] = cf;
  }
}

// One warp per block, over its chunks of 32 groups (1024 positions): chunk_reach / chunk_reach_nf become the reach of
// everything LEFT of the chunk (exclusive prefix maximum), chunk_first the first match candidate BEHIND it (exclusive
// suffix minimum).  128 steps of 32 chunks for a 4 MiB block.
// (kernels that give one warp a sequential job pack four such warps into a CTA: a warp's scheduler is
// its index in the CTA modulo 4, so single-warp CTAs would all queue on the first of the SM's four schedulers)
__global__ void __launch_bounds__(128)
k_dp_chunk_scan(uint32_t* chunk_reach, uint32_t* chunk_reach_nf, uint32_t* chunk_first, uint32_t chunks_per_block, Geom g)
{
  const uint32_t j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= g.n_blocks) return;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t* cr = chunk_reach + (size_t)j * chunks_per_block;
  uint32_t* cn = chunk_reach_nf + (size_t)j * chunks_per_block;
  uint32_t* cf = chunk_first + (size_t)j * chunks_per_block;
  uint32_t carry = 0, carry_nf = 0;
#pragma unroll 2
  for (uint32_t c0 = 0; c0 < chunks_per_block; c0 += 32)
  {
    const bool in = c0 + lane < chunks_per_block;
    uint32_t incl = in ? cr[c0 + lane] : 0, incl_nf = in ? cn[c0 + lane] : 0;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d), tn = __shfl_up_sync(0xffffffffu, incl_nf, d);
      if (lane >= d) { incl = max(incl, t); incl_nf = max(incl_nf, tn); }
    }
    uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1), before_nf = __shfl_up_sync(0xffffffffu, incl_nf, 1);
    before = lane == 0 ? carry : max(before, carry);
    before_nf = lane == 0 ? carry_nf : max(before_nf, carry_nf);
    if (in) { cr[c0 + lane] = before; cn[c0 + lane] = before_nf; }
    carry = max(carry, __shfl_sync(0xffffffffu, incl, 31));
    carry_nf = max(carry_nf, __shfl_sync(0xffffffffu, incl_nf, 31));
  }
  uint32_t behind = 0xffffffffu;
#pragma unroll 2
  for (int32_t c0 = (int32_t)((chunks_per_block - 1) & ~31u); c0 >= 0; c0 -= 32)
  {
    const bool in = (uint32_t)c0 + lane < chunks_per_block;
    uint32_t f = in ? cf[(uint32_t)c0 + lane] : 0xffffffffu;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_down_sync(0xffffffffu, f, d);
      if (lane + d < 32) f = min(f, t);
    }
    uint32_t after = __shfl_down_sync(0xffffffffu, f, 1);
    after = lane == 31 ? behind : min(after, behind);
    if (in) cf[(uint32_t)c0 + lane] = after;
    behind = min(behind, __shfl_sync(0xffffffffu, f, 0));
  }
}

// One warp per chunk: reach_before[g] = largest position + length over everything left of group g (the redo of
// k_dp_verify stops early with it).
__global__ void __launch_bounds__(256)
k_dp_reach_before(const uint32_t* group_reach, const uint32_t* chunk_reach, uint32_t* reach_before, uint32_t groups_per_block, Geom g)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t chunks_per_block = groups_per_block / 32;
  if (warp >= g.n_blocks * chunks_per_block) return;
  const size_t at = (size_t)warp * 32 + lane;                  // = block * groups_per_block + chunk * 32 + lane
  uint32_t incl = group_reach[at];
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1)
  {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = max(incl, t);
  }
  uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
  const uint32_t carry = chunk_reach[warp];
  reach_before[at] = lane == 0 ? carry : max(before, carry);
}

// One warp per block picks the segment boundaries: the first group at least kDpSeg behind the previous boundary
// where nothing from the left reaches more than kDpSlack beyond it -- or that lies inside a long run of unpriced
// matches, see DpTask.  Only the chunks it looks at are evaluated (a boundary is usually found in the first one).
struct DpPlanIn
{
  const uint32_t* group_reach; const uint32_t* group_reach_nf; const uint32_t* group_run_end; const uint32_t* group_first;
  const uint32_t* chunk_reach; const uint32_t* chunk_reach_nf; const uint32_t* chunk_first;      // after k_dp_chunk_scan
};

__global__ void __launch_bounds__(128)
k_dp_plan(DpPlanIn in, uint32_t groups_per_block, DpTask* tasks, uint32_t* task_count, uint32_t max_seg, Geom g)
{
  const uint32_t j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= g.n_blocks) return;
  const uint32_t n = block_len(g, j);
  const uint32_t lane = threadIdx.x & 31;
  DpTask* out = tasks + (size_t)j * max_seg;
  if (n <= kEndNoMatch) { if (lane == 0) task_count[j] = 0; return; }     // smallz4.h:755
  const uint32_t chunks_per_block = groups_per_block / 32;
  const uint32_t* gr = in.group_reach + (size_t)j * groups_per_block;
  const uint32_t* gn = in.group_reach_nf + (size_t)j * groups_per_block;
  const uint32_t* ge = in.group_run_end + (size_t)j * groups_per_block;
  const uint32_t* gf = in.group_first + (size_t)j * groups_per_block;
  const uint32_t* cr = in.chunk_reach + (size_t)j * chunks_per_block;
  const uint32_t* cn = in.chunk_reach_nf + (size_t)j * chunks_per_block;
  const uint32_t* cf = in.chunk_first + (size_t)j * chunks_per_block;
  const uint32_t groups = (n + 31) / 32;
  uint32_t count = 0, last = 0;
  uint32_t gq = kDpSeg / 32;                                     // first group that may carry the next boundary
  while (gq < groups && count + 2 < max_seg)
  {
    const uint32_t c = gq / 32, gi = c * 32 + lane;
    // reach of everything left of group gi: the chunk's carry and the groups of the chunk in front of it
    uint32_t incl = gr[gi], incl_nf = gn[gi];
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d), tn = __shfl_up_sync(0xffffffffu, incl_nf, d);
      if (lane >= d) { incl = max(incl, t); incl_nf = max(incl_nf, tn); }
    }
    uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1), before_nf = __shfl_up_sync(0xffffffffu, incl_nf, 1);
    const uint32_t carry = cr[c], carry_nf = cn[c];
    before = lane == 0 ? carry : max(before, carry);
    before_nf = lane == 0 ? carry_nf : max(before_nf, carry_nf);
    const uint32_t x0 = gi * 32;
    const bool inside = gi >= gq && gi < groups && x0 + 64 <= n;
    const bool fits = inside && before <= x0 + kDpSlack;
    // ... or the boundary lies inside a long run of unpriced matches that all end at the same place, its whole
    // warm-up zone does too, and nothing else reaches across
    uint32_t run_end = 0;
    if (inside && !fits && before_nf <= x0 + kDpSlack && gi + kDpWarm / 32 + 1 < groups)
    {
      const uint32_t e = ge[gi];
      if (e != 0 && ge[gi + kDpWarm / 32 + 1] == e) run_end = e;
    }
    const uint32_t cand = __ballot_sync(0xffffffffu, fits || run_end != 0);
    if (cand == 0) { gq = c * 32 + 32; continue; }
    const int l = __ffs((int)cand) - 1;
    const uint32_t xb = (c * 32 + (uint32_t)l) * 32;
    const uint32_t e = __shfl_sync(0xffffffffu, run_end, l);
    const uint32_t reach = __shfl_sync(0xffffffffu, e != 0 ? before_nf : before, l);
    // literal counter at the top of the warm-up zone when no match candidate is near (see DpTask)
    uint32_t ub = 0;
    const uint32_t top = xb + kDpWarm;
    if (e == 0 && top + 64 <= n)
    {
      // first candidate at or behind `top`: the rest of its chunk, then everything behind the chunk
      const uint32_t tg = top / 32, tc = tg / 32;
      uint32_t f = tc * 32 + lane >= tg ? gf[tc * 32 + lane] : 0xffffffffu;
      f = min(__reduce_min_sync(0xffffffffu, f), cf[tc]);
      if (f == 0xffffffffu) ub = until_bump_after(15 - kEndLiterals, n - kEndLiterals - top);
      else if (f - top >= 64) ub = until_bump_after(15, f - top);
    }
    if (lane == 0) { DpTask t; t.lo = last; t.hi = xb; t.reach = reach; t.aux = e | (ub << 24); out[count] = t; }
    count++;
    last = xb;
    gq = (xb + kDpSeg) / 32;
  }
  if (lane == 0) { DpTask t; t.lo = last; t.hi = n; t.reach = n; t.aux = 0; out[count] = t; task_count[j] = count + 1; }
}

__device__ __forceinline__ DpOverlay overlay_of(uint32_t* base, uint32_t task_index)
{
  DpOverlay o;
  o.cost = base + (size_t)task_index * 4 * kDpOvl;
  o.st5 = o.cost + kDpOvl; o.st6 = o.st5 + kDpOvl; o.st7 = o.st6 + kDpOvl;
  return o;
}

// Starting order of the tasks of k_dp_spec: the long ones first (a segment that could not be cut -- a run of
// matches shorter than 65 299 each, the unforced tail of a long run -- takes several times the nominal segment,
// and started last it would finish alone).  One CTA, stable partition into four classes by length.
__global__ void __launch_bounds__(256)
k_dp_task_order(const DpTask* tasks, const uint32_t* task_count, uint32_t max_seg, uint32_t n_tasks, uint32_t* order)
{
  __shared__ uint32_t ws[32], tot;
  const uint32_t per = (n_tasks + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(threadIdx.x * per, n_tasks), hi = min(lo + per, n_tasks);
  auto cls_of = [&](uint32_t i) -> uint32_t
  {
    if (i % max_seg >= task_count[i / max_seg]) return 3u;       // not a task
    const uint32_t len = tasks[i].hi - tasks[i].lo;
    return len >= 4 * kDpSeg ? 0u : (len >= 2 * kDpSeg ? 1u : 2u);
  };
  uint32_t base = 0;
  for (uint32_t cls = 0; cls < 4; cls++)
  {
    uint32_t mine = 0;
    for (uint32_t i = lo; i < hi; i++) mine += cls_of(i) == cls ? 1u : 0u;
    uint32_t at = base + block_excl_scan(mine, ws, &tot);
    for (uint32_t i = lo; i < hi; i++) if (cls_of(i) == cls) order[at++] = i;
    base += tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(32)
k_dp_spec(const uint32_t* mlen, const uint16_t* mdist, uint32_t* mfin, DpScratch s, const DpTask* tasks,
          const uint32_t* task_count, DpState* states, uint32_t* overlays, uint32_t max_seg, uint32_t* stats, const uint32_t* order, Geom g)
{
  SZ4_DYN_SMEM(smem);
  const uint32_t warp = 0;
  const uint32_t ti = order[blockIdx.x];                                // one task per (single-warp) CTA: tasks differ a lot in length
  const uint32_t j = ti / max_seg, k = ti % max_seg;
  if (j >= g.n_blocks || k >= task_count[j]) return;
  const DpTask t = tasks[ti];
  uint32_t ub_hi, ub_lo;
#ifndef SZ4_EMU
  const long long t0 = clock64();
#endif
  dp_segment(mlen, mdist, mfin, s, overlay_of(overlays, ti), block_begin(g, j), block_len(g, j), t.lo, t.hi,
             false, 0, smem + warp * kDpSmem, ub_hi, ub_lo, nullptr, nullptr, t.run_end(), t.ub_start());
  if ((threadIdx.x & 31) == 0)
  {
    DpState st; st.ub_hi = ub_hi; st.ub_lo = ub_lo; st.redone = 0; st.cum = 0; states[ti] = st;
#ifndef SZ4_EMU
    if (stats) { const unsigned dt = (unsigned)((clock64() - t0) >> 10); atomicMax(stats + 4, dt); atomicAdd(stats + 5, dt); if (dt > 4000) atomicMax(stats + 6, t.hi - t.lo); }
#endif
  }
}

__global__ void __launch_bounds__(128, 1)
k_dp_verify(const uint32_t* mlen, const uint16_t* mdist, uint32_t* mfin, DpScratch s, const DpTask* tasks,
            const uint32_t* task_count, DpState* states, uint32_t* overlays, uint32_t max_seg, uint32_t* redo_count,
            bool allow_early_stop, const uint32_t* reach_before, uint32_t groups_per_block, Geom g)
{
  SZ4_DYN_SMEM(smem_all);
  unsigned char* smem = smem_all + (threadIdx.x >> 5) * kDpSmem;
  const uint32_t j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= g.n_blocks) return;
  const uint32_t cnt = task_count[j];
  if (cnt < 2) return;
  const uint32_t b = block_begin(g, j), n = block_len(g, j);
  const uint32_t lane = threadIdx.x & 31;
  uint32_t cum = 0;                                              // costs of the segment to the right minus true costs: 0 at the top
#ifndef SZ4_EMU
  const long long tv0 = clock64();
  long long tredo = 0;
#endif
  for (int32_t k = (int32_t)cnt - 2; k >= 0; k--)
  {
    const uint32_t idx = j * max_seg + (uint32_t)k;
    const DpTask t = tasks[idx];
    DpState st = states[idx];
    const DpState right = states[idx + 1];
    const DpOverlay o = overlay_of(overlays, idx);
    // the span positions left of hi can look at: [hi, max(hi, reach)]
    const uint32_t span = (t.reach > t.hi ? t.reach - t.hi : 0) + 1;
    const uint32_t delta = __ldcg(s.cost + b + t.hi) - __ldcg(o.cost);
    bool same = st.ub_hi == right.ub_lo && span <= kDpWarm;
    for (uint32_t y = lane; y < span && y < kDpOvl; y += 32)
      if (__ldcg(s.cost + b + t.hi + y) - __ldcg(o.cost + y) != delta) same = false;
    same = __all_sync(0xffffffffu, same);
    if (same)
    {
      // costs of this segment = costs in the right neighbour's frame - delta
      cum = cum - delta;
      st.cum = cum;
    }
    else
    {
      uint32_t ub_hi, ub_lo;
#ifndef SZ4_EMU
      const long long tr0 = clock64();
#endif
      bool stopped = false;
      ub_lo = st.ub_lo;                                          // an early stop keeps the earlier pass's state at lo
      dp_segment(mlen, mdist, mfin, s, o, b, n, t.lo, t.hi, true, right.ub_lo, smem, ub_hi, ub_lo, allow_early_stop ? reach_before + (size_t)j * groups_per_block : nullptr, &stopped);
#ifndef SZ4_EMU
      tredo += clock64() - tr0;
#endif
      st.ub_lo = ub_lo; st.redone = 1; st.cum = cum;
      if (lane == 0) atomicAdd(redo_count, 1u);
    }
    if (lane == 0) states[idx] = st;
    __syncwarp();
  }
#ifndef SZ4_EMU
  if (lane == 0) { atomicMax(redo_count + 8, (unsigned)((clock64() - tv0) >> 10)); atomicMax(redo_count + 9, (unsigned)(tredo >> 10)); }
#endif
}

// debug only: bring every segment's costs into the true frame (they differ by a constant per segment)
__global__ void __launch_bounds__(256)
k_dp_cost_fix(DpScratch s, const DpTask* tasks, const uint32_t* task_count, const DpState* states, uint32_t max_seg, Geom g)
{
  const uint32_t j = blockIdx.x / max_seg, k = blockIdx.x % max_seg;
  if (j >= g.n_blocks || k >= task_count[j]) return;
  const DpTask t = tasks[blockIdx.x];
  const uint32_t cum = states[blockIdx.x].cum;
  if (cum == 0) return;
  const uint32_t b = block_begin(g, j);
  for (uint32_t y = t.lo + threadIdx.x; y < t.hi; y += blockDim.x) s.cost[b + y] -= cum;
}


// ---------------------------------------------------------------------------------------------
// Walk the chosen parse from the start of the block and list its sequences.
// record = { position of the match, match length (0 = final literals), distance, offset of the
// sequence inside the compressed block }.  The literals of a sequence start where the previous
// record's match ended.
// ---------------------------------------------------------------------------------------------
struct SeqRec { uint32_t pos, len, dist, out; };

__device__ __forceinline__ uint32_t seq_bytes(uint32_t lits, uint32_t len, bool last)
{
  uint32_t sz = 1 + len_ext_bytes(lits) + lits;
  if (!last)
  {
    sz += 2;
    if (len >= kMinMatch + 15) sz += 1 + (len - kMinMatch - 15) / 255;
  }
  return sz;
}

__device__ __forceinline__ void path_load(const uint32_t* mlen, const uint16_t* mdist, uint32_t b, uint32_t n, uint32_t sw,
                                          uint32_t lane, uint32_t (&L)[4], uint32_t (&D)[4])
{
#pragma unroll
  for (uint32_t k = 0; k < 4; k++)
  {
    const uint32_t i = sw + k * 32 + lane;
    L[k] = 0; D[k] = 0;
    if (i < n) { L[k] = mlen[b + i]; D[k] = mdist[b + i]; }
  }
}

#ifndef SZ4_PATH_SEG
#define SZ4_PATH_SEG 32768
#endif
#ifndef SZ4_PATH_WARM
#define SZ4_PATH_WARM 2048
#endif
enum : uint32_t { kPathSeg = SZ4_PATH_SEG, kPathWarm = SZ4_PATH_WARM };

// Walk the parse from position `from` up to (not including) `hi` and list the matches that start in [lo, hi).
//   cover   = end of the last match that starts before lo  (so the walk enters [lo, hi) at max(lo, cover))
//   leave   = end of the last match that starts before hi  (so it leaves at max(hi, leave))
__device__ __forceinline__ void path_walk(const uint32_t* mlen, const uint16_t* mdist, uint32_t b, uint32_t n, uint32_t from,
                                          uint32_t lo, uint32_t hi, SeqRec* out, uint32_t& count, uint32_t& cover, uint32_t& leave)
{
  const uint32_t lane = threadIdx.x & 31;
  uint32_t at = from;         // current position on the path (block relative)
  count = 0; cover = 0; leave = 0;
  // the walk reads 128 positions at a time and keeps the following 256 in flight (two sets of registers,
  // enough to cover a DRAM round trip); it only lists the matches on the path -- sizes and output
  // offsets are computed in parallel afterwards (k_seq_scan)
  uint32_t cur_sw = 0xffffffffu, n1_sw = 0xffffffffu, n2_sw = 0xffffffffu;
  uint32_t Lc[4], Dc[4], L1[4], D1[4], L2[4], D2[4];
  while (at < hi)
  {
    const uint32_t sw = at & ~127u;
    if (sw != cur_sw)
    {
      if (sw == n1_sw)
      {
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) { Lc[k] = L1[k]; Dc[k] = D1[k]; L1[k] = L2[k]; D1[k] = D2[k]; }
        n1_sw = n2_sw;
      }
      else if (sw == n2_sw)
      {
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) { Lc[k] = L2[k]; Dc[k] = D2[k]; }
        n1_sw = sw + 128;
        if (n1_sw < n) path_load(mlen, mdist, b, n, n1_sw, lane, L1, D1);
      }
      else
      {
        path_load(mlen, mdist, b, n, sw, lane, Lc, Dc);
        n1_sw = sw + 128;
        if (n1_sw < n) path_load(mlen, mdist, b, n, n1_sw, lane, L1, D1);
      }
      cur_sw = sw;
      n2_sw = sw + 256;
      if (n2_sw < n) path_load(mlen, mdist, b, n, n2_sw, lane, L2, D2);
    }
    const uint32_t k = (at - sw) >> 5;
    const uint32_t w = sw + k * 32;
    const uint32_t L = k == 0 ? Lc[0] : (k == 1 ? Lc[1] : (k == 2 ? Lc[2] : Lc[3]));
    const uint32_t Dd = k == 0 ? Dc[0] : (k == 1 ? Dc[1] : (k == 2 ? Dc[2] : Dc[3]));
    const uint32_t is_match = __ballot_sync(0xffffffffu, L > 1);
    uint32_t o = at - w;
    while (o < 32)
    {
      const uint32_t m = is_match & (0xffffffffu << o);
      if (m == 0) { at = w + 32; break; }
      const int ml = __ffs((int)m) - 1;
      const uint32_t pos = w + (uint32_t)ml;
      if (pos >= hi) { at = hi; break; }
      const uint32_t len = __shfl_sync(0xffffffffu, L, ml);
      const uint32_t dist = __shfl_sync(0xffffffffu, Dd, ml);
      if (pos >= lo)
      {
        if (lane == 0) { SeqRec r; r.pos = pos; r.len = len; r.dist = dist; r.out = 0; out[count] = r; }
        count++;
      }
      else cover = pos + len;
      leave = pos + len;
      at = pos + len;
      o = at - w;                      // may be >= 32: leaves the window
    }
  }
}

// Per block, the parse is walked in fixed segments of kPathSeg positions.  A walk started at an arbitrary
// position joins the true path as soon as both land on the same position, so every segment starts
// kPathWarm positions early (k_path_spec); k_path_join then goes through the segments in order: a
// segment whose assumed entry point equals the true one (where the segment before it left) is kept,
// otherwise it is walked again from the true entry.  k_path_compact makes the record list contiguous.
struct PathSeg { uint32_t entry, leave, count, base; };

__device__ __forceinline__ SeqRec* path_seg_records(SeqRec* tmp, uint32_t seq_stride, uint32_t j, uint32_t k, uint32_t min_len)
{
  return tmp + (size_t)j * seq_stride + (size_t)k * (kPathSeg / min_len + 1);
}

__global__ void __launch_bounds__(128)
k_path_spec(const uint32_t* mlen, const uint16_t* mdist, SeqRec* tmp, uint32_t seq_stride, PathSeg* segs, uint32_t max_seg,
            uint32_t min_len, Geom g)
{
  const uint32_t si = blockIdx.x * 4 + (threadIdx.x >> 5);       // one segment per warp
  const uint32_t j = si / max_seg, k = si % max_seg;
  if (j >= g.n_blocks) return;
  const uint32_t b = block_begin(g, j), n = block_len(g, j);
  const uint32_t lo = k * kPathSeg;
  if (lo >= n && !(k == 0 && n == 0)) return;
  const uint32_t hi = min(lo + (uint32_t)kPathSeg, n);
  const uint32_t from = lo > kPathWarm ? lo - kPathWarm : 0;
  uint32_t count, cover, leave;
  path_walk(mlen, mdist, b, n, from, lo, hi, path_seg_records(tmp, seq_stride, j, k, min_len), count, cover, leave);
  if ((threadIdx.x & 31) == 0)
  {
    PathSeg ps; ps.entry = max(lo, cover); ps.leave = max(hi, leave); ps.count = count; ps.base = 0;
    segs[si] = ps;
  }
}

__global__ void __launch_bounds__(128)
k_path_join(const uint32_t* mlen, const uint16_t* mdist, SeqRec* tmp, uint32_t seq_stride, PathSeg* segs, uint32_t max_seg,
            uint32_t min_len, uint32_t* seq_count, uint32_t* redo_count, Geom g)
{
  const uint32_t j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= g.n_blocks) return;
  const uint32_t b = block_begin(g, j), n = block_len(g, j);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nseg = (n + kPathSeg - 1) / kPathSeg;
  uint32_t entry = 0, base = 0;                                  // true entry into the next segment; records so far
  for (uint32_t k = 0; k < nseg; k++)
  {
    const uint32_t idx = j * max_seg + k;
    PathSeg ps = segs[idx];
    const uint32_t lo = k * kPathSeg, hi = min(lo + (uint32_t)kPathSeg, n);
    if (ps.entry != entry)
    {
      if (entry >= hi) { ps.count = 0; ps.leave = entry; }       // a match jumps over the whole segment
      else
      {
        uint32_t count, cover, leave;
        path_walk(mlen, mdist, b, n, entry, lo, hi, path_seg_records(tmp, seq_stride, j, k, min_len), count, cover, leave);
        ps.count = count; ps.leave = max(hi, leave);
      }
      ps.entry = entry;
      if (lane == 0) atomicAdd(redo_count + 1, 1u);
    }
    ps.base = base;
    if (lane == 0) segs[idx] = ps;
    base += ps.count;
    entry = ps.leave;
    __syncwarp();
  }
  if (lane == 0) seq_count[j] = base + 1;                        // + the final literals
}

__global__ void __launch_bounds__(128)
k_path_compact(const SeqRec* tmp, SeqRec* seqs, uint32_t seq_stride, const PathSeg* segs, uint32_t max_seg, uint32_t min_len,
               const uint32_t* seq_count, Geom g)
{
  const uint32_t j = blockIdx.x / max_seg, k = blockIdx.x % max_seg;
  if (j >= g.n_blocks) return;
  const uint32_t n = block_len(g, j);
  SeqRec* out = seqs + (size_t)j * seq_stride;
  if (k == 0 && threadIdx.x == 0)
  {
    // final literals (smallz4.h:292-308: the last token has no match)
    SeqRec r; r.pos = n; r.len = 0; r.dist = 0; r.out = 0;
    out[seq_count[j] - 1] = r;
  }
  if (k * kPathSeg >= n) return;
  const PathSeg ps = segs[blockIdx.x];
  const SeqRec* src = tmp + (size_t)j * seq_stride + (size_t)k * (kPathSeg / min_len + 1);
  for (uint32_t t = threadIdx.x; t < ps.count; t += blockDim.x) out[ps.base + t] = src[t];
}

// Offsets of the sequences inside the compressed block = exclusive prefix sum of their sizes
// (smallz4.h:303-367 gives the size of one sequence).  One CTA per block.
__global__ void __launch_bounds__(1024)
k_seq_scan(SeqRec* seqs, uint32_t seq_stride, const uint32_t* seq_count, uint32_t* packed_size, Geom g)
{
  __shared__ uint32_t ws[32];
  __shared__ uint32_t tot;
  const uint32_t j = blockIdx.x;
  if (j >= g.n_blocks) return;
  SeqRec* sq = seqs + (size_t)j * seq_stride;
  const uint32_t cnt = seq_count[j];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < cnt; base += 4 * blockDim.x)
  {
    const uint32_t k0 = base + threadIdx.x * 4;
    uint32_t sz[4] = { 0, 0, 0, 0 };
    uint32_t prev_end = 0;
    if (k0 > 0 && k0 < cnt) { const SeqRec q = sq[k0 - 1]; prev_end = q.pos + q.len; }
#pragma unroll
    for (uint32_t t = 0; t < 4; t++)
      if (k0 + t < cnt)
      {
        const SeqRec r = sq[k0 + t];
        sz[t] = seq_bytes(r.pos - prev_end, r.len, r.len == 0);
        prev_end = r.pos + r.len;
      }
    uint32_t run = carry + block_excl_scan(sz[0] + sz[1] + sz[2] + sz[3], ws, &tot);
#pragma unroll
    for (uint32_t t = 0; t < 4; t++)
      if (k0 + t < cnt) { sq[k0 + t].out = run; run += sz[t]; }
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) packed_size[j] = carry;
}

}  // namespace sz4
