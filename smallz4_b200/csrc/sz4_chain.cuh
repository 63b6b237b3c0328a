// sz4_chain.cuh -- phase 1 of DICTIONARY streams: previousHash / previousExact for every position of a batch, in parallel.
// (Streams without a dictionary get prefix-class tables from sz4_lsd.cuh instead and never come here.)
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
