// sz4_chain.cuh -- phase 1: previousHash / previousExact for every position of a batch, in parallel.
//
// Reference: smallz4.h:645-720.  For position p the reference stores
//   previousHash[p]  = distance to the most recent earlier inserted position with the same hash20
//                      (0 if none or farther than 65535)                           smallz4.h:659-676
//   previousExact[p] = distance to the first position on that hash chain whose four bytes equal
//                      p's (0 if the walk ends first)                              smallz4.h:681-720
// Both live in 65536-slot rings there; here they are flat arrays indexed by position ("ph", "pe").
// A reader at position r sees  ring[r & 0xFFFF]  which is the entry of position r - shift
// (shift = 0, or 1 when a dictionary offsets the blocks by 65535: DESIGN.md Q-dict), so readers
// use  ph[r - shift]  /  pe[r - shift].
//
// Reference quirk Q-twice: the position 12 bytes before the end of a block is inserted again by the
// next block's lookback (smallz4.h:615-629) and finds itself at distance 0, so both of its ring
// entries become 0 for every later reader.  k_twice_save keeps the first value (its own search in
// the earlier block used it) in a side array and zeroes the flat entry.
#pragma once
#include "sz4_device.cuh"

namespace sz4
{
// sorted[r] = (four bytes << 32) | position, ascending hash20 of the bytes, then position.
// previousHash entry of position p whose predecessor in its hash class is q (smallz4.h:659-676, 783-795)
__device__ __forceinline__ uint32_t hash_link(const Geom& g, uint32_t p, uint32_t q)
{
  uint32_t d = p - q;
  if (d > kWindow) d = 0;                                                     // smallz4.h:668
  else if (g.legacy && (q - g.halo) / g.block_size != (p - g.halo) / g.block_size) d = 0;
  return d;
}

// previousHash only (the dictionary path, whose shifted ring k_exact_walk below replays from the flat array)
__global__ void __launch_bounds__(256)
k_link(const uint64_t* sorted, uint32_t n, uint16_t* ph, Geom g)
{
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  uint64_t e = sorted[r];
  uint32_t p = (uint32_t)e;
  uint32_t d = 0;
  if (r > 0)
  {
    uint64_t e0 = sorted[r - 1];
    if (hash20((uint32_t)(e0 >> 32)) == hash20((uint32_t)(e >> 32))) d = hash_link(g, p, (uint32_t)e0);
  }
  ph[p] = (uint16_t)d;
}

// previousExact straight from the sorted array (no dictionary): the elements in front of r with the same hash ARE
// the hash chain of position p (smallz4.h:681-720 follows previousHash from p), nearest first, and they carry
// their four bytes.  Same stops as k_exact_walk: a link longer than 65535 or across a legacy block, the sum of
// the links beyond 65535, the walk leaving what the reference could look at (floor_of), and a chain member whose
// ring entry the next block's lookback has zeroed (Q-twice).  One coalesced read per chain member instead of a
// 2-byte and a 4-byte random read; previousHash itself is never stored.
__global__ void __launch_bounds__(256)
k_chain(const uint64_t* sorted, uint32_t n, uint16_t* pe, Geom g)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = r < n;                                    // (no early exit: the warp works together further down)
  // the element and its two predecessors are fetched together: nearly every walk ends at one of them
  const uint64_t e = valid ? sorted[r] : 0, e1 = valid && r >= 1 ? sorted[r - 1] : 0, e2 = valid && r >= 2 ? sorted[r - 2] : 0;
  const uint32_t p = (uint32_t)e, four = (uint32_t)(e >> 32), h = hash20(four);
  const uint32_t floor_pos = floor_of(g, p);
  // The only chain member with a zeroed ring entry a walk from p can meet is the twice-inserted position of the
  // nearest block border behind p (blocks are longer than the window).  A link across a legacy block needs no
  // test of its own: its target lies below floor_pos, which ends the walk with the same "no match".
  uint32_t zeroed = 0xffffffffu;
  if (!g.legacy && p + kEndNoMatch >= g.halo)
  {
    const uint32_t x = g.halo + (p + kEndNoMatch - g.halo) / g.block_size * g.block_size;
    if (x >= kEndNoMatch && x - kEndNoMatch < p && is_twice_inserted(g, x - kEndNoMatch)) zeroed = x - kEndNoMatch;
  }
  uint32_t result = 0, total = 0, at = p;
  // one chain member: true when the walk is over
  auto visit = [&](uint64_t e0) -> bool
  {
    const uint32_t w = (uint32_t)(e0 >> 32), q = (uint32_t)e0;
    if (hash20(w) != h) return true;                           // front of the hash class
    if (at == zeroed) return true;                             // ring entry zeroed by the next block's lookback (Q-twice)
    const uint32_t step = at - q;
    if (step > kWindow) return true;                           // smallz4.h:668
    total += step;
    if (total > kWindow) return true;
    if (q < floor_pos) return true;                            // in front of the reference's buffer (UB-1): "no match"
    if (w == four) { result = total; return true; }
    if (total == kWindow) return true;
    at = q;
    return false;
  };
  // Nearly every walk ends at its first or second member.  The few that go on (another frequent word in the same hash
  // class, e.g. a colliding word behind a long run: up to 65535 members) are taken over by the whole warp, 32 members
  // per step: nothing in a step depends on the step before it (the sum of the links so far is p - q).
  uint32_t k = r;
  bool done = !valid || k == 0;
  if (!done) { k--; done = visit(e1) || k == 0; }
  if (!done) { k--; done = visit(e2) || k == 0; }
  const uint32_t lane = threadIdx.x & 31;
  uint32_t pending = __ballot_sync(0xffffffffu, !done);
  while (pending)
  {
    const int src = __ffs((int)pending) - 1;
    pending &= pending - 1;
    uint32_t bk = __shfl_sync(0xffffffffu, k, src), bat = __shfl_sync(0xffffffffu, at, src);
    const uint32_t bp = __shfl_sync(0xffffffffu, p, src), bfour = __shfl_sync(0xffffffffu, four, src);
    const uint32_t bfloor = __shfl_sync(0xffffffffu, floor_pos, src), bzero = __shfl_sync(0xffffffffu, zeroed, src);
    const uint32_t bh = hash20(bfour);
    uint32_t found = 0;
    for (;;)
    {
      const bool have = bk > lane;                             // member bk - 1 - lane of the sorted array
      const uint64_t e0 = have ? sorted[bk - 1 - lane] : 0;
      const uint32_t w = (uint32_t)(e0 >> 32), q = (uint32_t)e0;
      uint32_t prev = __shfl_up_sync(0xffffffffu, q, 1);       // the member in front of it on the chain
      if (lane == 0) prev = bat;
      const uint32_t total_here = bp - q;
      // the same tests in the same order as `visit`
      const bool dead = !have || hash20(w) != bh || prev == bzero || prev - q > kWindow || total_here > kWindow || q < bfloor;
      const bool hit = !dead && w == bfour;
      const bool stop = dead || hit || total_here == kWindow;
      const uint32_t stops = __ballot_sync(0xffffffffu, stop);
      if (stops != 0)
      {
        const int first = __ffs((int)stops) - 1;
        found = __shfl_sync(0xffffffffu, hit ? total_here : 0u, first);
        break;
      }
      bat = __shfl_sync(0xffffffffu, q, 31);
      bk -= 32;
      if (bk == 0) break;
    }
    if (lane == (uint32_t)src) result = found;
  }
  if (valid) pe[p] = (uint16_t)result;
}

// one thread per block border k: position halo + k*block_size - 12
__global__ void k_twice_save(uint16_t* arr, uint32_t* saved, Geom g)
{
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= g.n_blocks) return;
  uint32_t border = g.halo + k * g.block_size;
  uint32_t v = 0xffffffffu;
  if (border >= kEndNoMatch)
  {
    uint32_t p = border - kEndNoMatch;
    if (is_twice_inserted(g, p)) { v = arr[p]; arr[p] = 0; }
  }
  saved[k] = v;
}

__device__ __forceinline__ uint32_t own_entry(const uint16_t* arr, const uint32_t* saved, const Geom& g, uint32_t p)
{
  if (is_twice_inserted(g, p)) return saved[(p + kEndNoMatch - g.halo) / g.block_size];
  return arr[p];
}

// previousExact: follow the hash chain until the four bytes match (smallz4.h:681-720)
__global__ void __launch_bounds__(256)
k_exact_walk(const uint8_t* data, const uint16_t* ph, const uint32_t* saved_ph, uint16_t* pe, uint32_t first, uint32_t n, Geom g)
{
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const uint32_t p = first + idx;
  uint32_t total = own_entry(ph, saved_ph, g, p);
  uint32_t result = 0;
  if (total != 0)
  {
    const uint32_t four = ld32u(data + p);
    const uint32_t h = hash20(four);
    const uint32_t floor_pos = floor_of(g, p);
    uint32_t at = p - total;
    // at < floor_pos on entry: the reference reads in front of its buffer (UB-1) -> defined as "no match"
    while (at >= floor_pos)
    {
      uint32_t there = ld32u(data + at);
      if (there == four) { result = total; break; }
      if (hash20(there) != h) break;                       // smallz4.h:690 (only reachable with shift = 1)
      if (total == kWindow) break;                         // whatever the next entry holds, the sum exceeds 65535
      uint32_t step = at >= g.shift ? ph[at - g.shift] : 0;   // smallz4.h:694
      if (step == 0) break;
      total += step;
      if (total > kWindow) break;
      at -= step;
    }
  }
  pe[p] = (uint16_t)result;
}

}  // namespace sz4
