// sz4_runs.cuh -- two per-position helpers for byte runs, both "distance to the nearest flagged
// position" scans (three small kernels each: per-chunk reduce, chunk carry, per-chunk apply):
//
//   run_fwd[p]  = number of bytes equal to data[p] starting at p          (flag: data[x] != data[x+1])
//   ones_back[p] = number of consecutive positions p, p-1, ... whose previousExact entry is 1
//                  (flag: pe[q] != 1), saturated at 65535
//
// The match finder uses them to step over whole runs of chain candidates that the reference
// visits one by one and provably rejects (sz4_search.cuh, "stretch").
#pragma once
#include "sz4_device.cuh"

namespace sz4
{
enum : uint32_t { kFlagThreads = 256, kFlagItems = 16, kFlagChunk = kFlagThreads * kFlagItems };
#define SZ4_NOFLAG 0xffffffffu

// Direction-agnostic index mapping: scan index k -> array position.
struct FlagFwdRuns      // scanning right-to-left over data: k = 0 is the last position
{
  const uint8_t* data; uint32_t n;
  __device__ __forceinline__ uint32_t pos(uint32_t k) const { return n - 1 - k; }
  __device__ __forceinline__ bool flag(uint32_t k) const { uint32_t x = n - 1 - k; return k == 0 || data[x] != data[x + 1]; }
};
struct FlagOnesBack     // scanning left-to-right over pe
{
  const uint16_t* pe; uint32_t n;
  __device__ __forceinline__ uint32_t pos(uint32_t k) const { return k; }
  __device__ __forceinline__ bool flag(uint32_t k) const { return pe[k] != 1; }
};

// last flagged scan index inside each chunk (or SZ4_NOFLAG)
template <typename F>
__global__ void __launch_bounds__(kFlagThreads) k_flag_reduce(F f, uint32_t* chunk_last)
{
  __shared__ uint32_t best;
  if (threadIdx.x == 0) best = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kFlagChunk + threadIdx.x * kFlagItems;
  uint32_t mine = 0;                                   // stored as index + 1, 0 = none
  for (uint32_t k = 0; k < kFlagItems; k++)
  {
    uint32_t i = base + k;
    if (i < f.n && f.flag(i)) mine = i + 1;
  }
  if (mine) atomicMax(&best, mine);
  __syncthreads();
  if (threadIdx.x == 0) chunk_last[blockIdx.x] = best ? best - 1 : SZ4_NOFLAG;
}

// one warp: chunk_carry[c] = last flagged index before chunk c (exclusive max-scan, 32 chunks per step)
__global__ void __launch_bounds__(32) k_flag_carry(const uint32_t* chunk_last, uint32_t* chunk_carry, uint32_t chunks)
{
  if (blockIdx.x != 0) return;
  const uint32_t lane = threadIdx.x;
  uint32_t run = 0;                                    // index + 1 of the last flag so far, 0 = none
  // four steps of 32 chunks per round, their loads issued together (the scan itself is a dependent chain)
  for (uint32_t c0 = 0; c0 < chunks; c0 += 128)
  {
    uint32_t vv[4];
#pragma unroll
    for (uint32_t u = 0; u < 4; u++) { const uint32_t c = c0 + 32 * u + lane; vv[u] = c < chunks ? chunk_last[c] : SZ4_NOFLAG; }
#pragma unroll
    for (uint32_t u = 0; u < 4; u++)
    {
      const uint32_t c = c0 + 32 * u + lane;
      const uint32_t v = vv[u];
      uint32_t incl = v == SZ4_NOFLAG ? 0 : v + 1;     // flags only move right: a later one is larger
#pragma unroll
      for (uint32_t d = 1; d < 32; d <<= 1)
      {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = max(incl, t);
      }
      uint32_t before = __shfl_up_sync(0xffffffffu, incl, 1);
      before = lane == 0 ? run : max(before, run);
      if (c < chunks) chunk_carry[c] = before ? before - 1 : SZ4_NOFLAG;
      run = max(run, __shfl_sync(0xffffffffu, incl, 31));
    }
  }
}

// out(pos(i)) = i - (last flagged index <= i)  [+1 for runs: counts the element itself], saturated
template <typename F, typename Out, bool kCountSelf>
__global__ void __launch_bounds__(kFlagThreads) k_flag_apply(F f, const uint32_t* chunk_carry, Out* out, uint32_t cap)
{
  __shared__ uint32_t tlast[kFlagThreads];
  const uint32_t base = blockIdx.x * kFlagChunk + threadIdx.x * kFlagItems;
  uint32_t mine = 0;
  for (uint32_t k = 0; k < kFlagItems; k++)
  {
    uint32_t i = base + k;
    if (i < f.n && f.flag(i)) mine = i + 1;
  }
  tlast[threadIdx.x] = mine;
  __syncthreads();
  // last flag in earlier threads of the chunk (simple backward search; flags are dense in practice)
  uint32_t prev = 0;
  for (int t = (int)threadIdx.x - 1; t >= 0; t--)
    if (tlast[t]) { prev = tlast[t]; break; }
  if (!prev)
  {
    uint32_t c = chunk_carry[blockIdx.x];
    prev = c == SZ4_NOFLAG ? 0 : c + 1;
  }
  for (uint32_t k = 0; k < kFlagItems; k++)
  {
    uint32_t i = base + k;
    if (i >= f.n) break;
    if (f.flag(i)) prev = i + 1;
    // prev - 1 = last flagged index <= i (prev == 0: none)
    uint32_t d = prev ? i - (prev - 1) : i + 1;
    if (kCountSelf) d += 1;
    out[f.pos(i)] = (Out)(d > cap ? cap : d);
  }
}

// maximum of an array (grid-stride; one atomic per CTA)
__global__ void __launch_bounds__(256) k_max_u32(const uint32_t* in, uint32_t n, uint32_t* out)
{
  __shared__ uint32_t best;
  if (threadIdx.x == 0) best = 0;
  __syncthreads();
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, in[i]);
  if (m) atomicMax(&best, m);
  __syncthreads();
  if (threadIdx.x == 0 && best) atomicMax(out, best);
}

}  // namespace sz4
