// sz4_runs.cuh -- two per-position helpers for byte runs (three small kernels: per-chunk reduce, chunk carry,
// per-chunk apply):
//
//   run_fwd[p]   = number of bytes equal to data[p] starting at p
//   ones_back[p] = number of consecutive positions p, p-1, ... whose 8-byte chain entry (pe8) is 1, saturated at 65535
//
// The match finder uses them to step over whole runs of chain candidates that the reference
// visits one by one and provably rejects (sz4_search.cuh, "stretch").
#pragma once
#include "sz4_device.cuh"

namespace sz4
{
#define SZ4_NOFLAG 0xffffffffu

// ---------------------------------------------------------------------------------------------
// Both helpers from the run structure of the data alone, in three launches (no dictionary).  A position x is a
// boundary if a run starts there (x == 0 or data[x] != data[x-1]).  With s = the last boundary <= p and e = the first
// boundary > p (or n):   run_fwd[p] = e - p;   pe8[p] == 1 (sz4_lsd.cuh) iff the nine bytes from p-1 on are equal, i.e.
// p > s and e - p >= 8, and then all of s+1 .. p have it too:   ones_back[p] = p - s  (saturated), else 0.
// ---------------------------------------------------------------------------------------------
enum : uint32_t { kRunThreads = 256, kRunChunk = 4096, kRunWords = kRunChunk / 32 };

__device__ __forceinline__ bool run_boundary(const uint8_t* data, uint32_t x) { return x == 0 || data[x] != data[x - 1]; }

// first and last boundary of every chunk (SZ4_NOFLAG: none)
__global__ void __launch_bounds__(kRunThreads) k_run_reduce(const uint8_t* data, uint32_t n, uint32_t* chunk_first, uint32_t* chunk_last)
{
  __shared__ uint32_t lo, hi;
  if (threadIdx.x == 0) { lo = SZ4_NOFLAG; hi = 0; }
  __syncthreads();
  const uint32_t base = blockIdx.x * kRunChunk;
  uint32_t mn = SZ4_NOFLAG, mx = 0;                                  // mx: position + 1, 0 = none
  for (uint32_t k = 0; k < kRunChunk / kRunThreads; k++)
  {
    const uint32_t x = base + k * kRunThreads + threadIdx.x;
    if (x < n && run_boundary(data, x)) { mn = min(mn, x); mx = x + 1; }
  }
  if (mx) { atomicMin(&lo, mn); atomicMax(&hi, mx); }
  __syncthreads();
  if (threadIdx.x == 0) { chunk_first[blockIdx.x] = lo; chunk_last[blockIdx.x] = hi ? hi - 1 : SZ4_NOFLAG; }
}

// one CTA: prev_b[c] = last boundary in the chunks in front of c, next_b[c] = first boundary behind c (or n)
__global__ void __launch_bounds__(256) k_run_carry(const uint32_t* chunk_first, const uint32_t* chunk_last, uint32_t chunks, uint32_t n,
                                                   uint32_t* prev_b, uint32_t* next_b)
{
  __shared__ uint32_t ws[8];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // forward: running maximum of (last boundary + 1)
  uint32_t run = 0;
  for (uint32_t c0 = 0; c0 < chunks; c0 += 256)
  {
    const uint32_t c = c0 + threadIdx.x;
    const uint32_t v = c < chunks ? chunk_last[c] : SZ4_NOFLAG;
    uint32_t incl = v == SZ4_NOFLAG ? 0 : v + 1;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = max(incl, t); }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t before = run;
    for (uint32_t w = 0; w < warp; w++) before = max(before, ws[w]);
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, 1);
    const uint32_t excl = max(before, lane == 0 ? 0u : up);
    if (c < chunks) prev_b[c] = excl ? excl - 1 : 0;                 // (position 0 is a boundary: chunk 0 never asks)
    for (uint32_t w = 0; w < 8; w++) run = max(run, ws[w]);
    __syncthreads();
  }
  // backward: running minimum of the first boundary
  uint32_t runmin = n;
  for (uint32_t c0 = 0; c0 < chunks; c0 += 256)
  {
    const uint32_t k = c0 + threadIdx.x;                             // k-th chunk from the end
    const uint32_t c = chunks - 1 - k;
    const uint32_t v = k < chunks ? chunk_first[c] : SZ4_NOFLAG;
    uint32_t incl = v == SZ4_NOFLAG ? n : v;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = min(incl, t); }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t before = runmin;
    for (uint32_t w = 0; w < warp; w++) before = min(before, ws[w]);
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, 1);
    const uint32_t excl = min(before, lane == 0 ? n : up);
    if (k < chunks) next_b[c] = excl;
    for (uint32_t w = 0; w < 8; w++) runmin = min(runmin, ws[w]);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kRunThreads) k_run_apply(const uint8_t* data, uint32_t n, const uint32_t* prev_b, const uint32_t* next_b,
                                                           uint32_t* run_fwd, uint16_t* ones_back)
{
  __shared__ uint32_t words[kRunWords];                              // boundary bits of the chunk
  __shared__ uint32_t prev_nz[kRunWords], next_nz[kRunWords];       // nearest non-empty word at or in front of / at or behind (index + 1, 0 = none)
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = blockIdx.x * kRunChunk;
  for (uint32_t k = 0; k < kRunChunk / kRunThreads; k++)
  {
    const uint32_t x = base + k * kRunThreads + threadIdx.x;
    const uint32_t m = __ballot_sync(0xffffffffu, x < n && run_boundary(data, x));
    if (lane == 0) words[k * (kRunThreads / 32) + warp] = m;
  }
  __syncthreads();
  if (threadIdx.x < kRunWords)
  {
    // 128 words: each thread looks for its nearest non-empty word (the boundaries are dense in anything but long runs)
    uint32_t a = 0, b = 0;
    for (int i = (int)threadIdx.x; i >= 0; i--) if (words[i]) { a = (uint32_t)i + 1; break; }
    for (uint32_t i = threadIdx.x; i < kRunWords; i++) if (words[i]) { b = i + 1; break; }
    prev_nz[threadIdx.x] = a; next_nz[threadIdx.x] = b;
  }
  __syncthreads();
  const uint32_t carry_s = prev_b[blockIdx.x], carry_e = next_b[blockIdx.x];
  for (uint32_t k = 0; k < kRunChunk / kRunThreads; k++)
  {
    const uint32_t j = k * kRunThreads + threadIdx.x, p = base + j;
    if (p >= n) break;
    const uint32_t w = j >> 5, bit = j & 31;
    const uint32_t le = words[w] & (0xffffffffu >> (31 - bit));      // boundaries at or in front of p in its word
    uint32_t s;
    if (le) s = base + w * 32 + (31 - (uint32_t)__clz((int)le));
    else
    {
      const uint32_t wp = w > 0 ? prev_nz[w - 1] : 0;
      s = wp ? base + (wp - 1) * 32 + (31 - (uint32_t)__clz((int)words[wp - 1])) : carry_s;
    }
    const uint32_t gt = bit == 31 ? 0u : words[w] & (0xffffffffu << (bit + 1));   // boundaries behind p in its word
    uint32_t e;
    if (gt) e = base + w * 32 + ((uint32_t)__ffs((int)gt) - 1);
    else
    {
      const uint32_t wn = w + 1 < kRunWords ? next_nz[w + 1] : 0;
      e = wn ? base + (wn - 1) * 32 + ((uint32_t)__ffs((int)words[wn - 1]) - 1) : carry_e;
    }
    run_fwd[p] = e - p;
    ones_back[p] = (p > s && e - p >= 8) ? (uint16_t)min(p - s, 65535u) : (uint16_t)0;
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
