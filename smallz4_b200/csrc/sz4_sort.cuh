// sz4_sort.cuh -- stable LSD radix sort of (four bytes, position) pairs by hash20 and a generic exclusive scan.
//
// Phase 1 of the per-block loop (smallz4.h:645-676: lastHash / previousHash) asks, for every
// position, for the most recent earlier position with the same 20-bit hash.  The reference gets
// it from a sequentially updated table; here all positions of a batch are sorted by hash with
// their order preserved, so the predecessor is simply the left neighbour in the sorted array
// (k_link in sz4_chain.cuh).  HBM-bound integer work: 3 passes x (histogram + scatter) over
// 8-byte elements, coalesced reads, digit-grouped writes.
#pragma once
#include "sz4_device.cuh"

namespace sz4
{
enum : uint32_t
{
  kSortThreads = 256,
  kSortItems   = 16,
  kSortTile    = kSortThreads * kSortItems,   // 4096 elements per CTA
  kSortBits    = 7,
  kSortBins    = 1u << kSortBits,             // 128
  kScanThreads = 256,
  kScanItems   = 16,
  kScanChunk   = kScanThreads * kScanItems    // 4096
};

// ------------------------------------------------------------------ exclusive scan (uint32 sum)
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1)
  {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (blockDim.x == kScanThreads); returns the exclusive
// prefix and leaves the block total in *total (shared)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums /* >= 32 */, uint32_t* total)
{
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0)
  {
    uint32_t w = lane < nwarps ? warp_sums[lane] : 0;
    uint32_t wi = warp_incl_scan(w, lane);
    warp_sums[lane] = wi - w;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  uint32_t r = incl - v + warp_sums[warp];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t* in, uint32_t n, uint32_t* partial)
{
  __shared__ uint32_t ws[32];
  __shared__ uint32_t tot;
  const uint32_t base = blockIdx.x * kScanChunk;
  uint32_t s = 0;
#pragma unroll
  for (uint32_t k = 0; k < kScanItems; k++)
  {
    uint32_t i = base + k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  (void)block_excl_scan(s, ws, &tot);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// single CTA: in-place exclusive scan of m partial sums
__global__ void __launch_bounds__(kScanThreads) k_scan_partials(uint32_t* partial, uint32_t m)
{
  __shared__ uint32_t ws[32];
  __shared__ uint32_t tot;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < m; base += kScanThreads)
  {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < m ? partial[i] : 0;
    uint32_t e = block_excl_scan(v, ws, &tot);
    if (i < m) partial[i] = carry + e;
    carry += tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* partial)
{
  __shared__ uint32_t ws[32];
  __shared__ uint32_t tot;
  // thread t owns kScanItems consecutive elements so that the scan order is the array order
  const uint32_t base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (uint32_t k = 0; k < kScanItems; k++)
  {
    uint32_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  uint32_t run = partial[blockIdx.x] + block_excl_scan(s, ws, &tot);
#pragma unroll
  for (uint32_t k = 0; k < kScanItems; k++)
  {
    uint32_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
}

// ------------------------------------------------------------------ radix sort passes
// element = (the four bytes at the position << 32) | position, sorted by hash20 of the four bytes (smallz4.h:164):
// the words ride along so that the exact chains (sz4_chain.cuh) never have to read the data.  Pass 0 builds the
// elements from the data.
template <bool kFromData>
__device__ __forceinline__ uint64_t sort_load(const uint64_t* in, const uint8_t* data, uint32_t first, uint32_t i)
{
  if (kFromData)
  {
    uint32_t p = first + i;
    return ((uint64_t)ld32u(data + p) << 32) | p;
  }
  return in[i];
}

template <bool kFromData>
__global__ void __launch_bounds__(kSortThreads)
k_sort_hist(const uint64_t* in, const uint8_t* data, uint32_t first, uint32_t n, uint32_t shift, uint32_t mask,
            uint32_t* hist, uint32_t num_tiles)
{
  __shared__ uint32_t h[kSortBins];
  if (threadIdx.x < kSortBins) h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
  for (uint32_t k = 0; k < kSortItems; k++)
  {
    uint32_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n)
    {
      uint64_t e = sort_load<kFromData>(in, data, first, i);
      atomicAdd(&h[(hash20((uint32_t)(e >> 32)) >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < kSortBins) hist[threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter.  Warp w of the CTA owns elements [w*512, (w+1)*512) of the tile, visited as 16
// rows of 32 in order; __match_any_sync ranks equal digits inside a row, per-warp counters rank
// rows, a scan over the warps ranks warps, the scanned histogram ranks tiles.
template <bool kFromData>
__global__ void __launch_bounds__(kSortThreads)
k_sort_scatter(const uint64_t* in, uint64_t* out, const uint8_t* data, uint32_t first, uint32_t n, uint32_t shift,
               uint32_t mask, const uint32_t* hist_scanned, uint32_t num_tiles)
{
  __shared__ uint32_t cnt[kSortThreads / 32][kSortBins];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t k = threadIdx.x; k < (kSortThreads / 32) * kSortBins; k += kSortThreads) (&cnt[0][0])[k] = 0;
  __syncthreads();

  const uint32_t base = blockIdx.x * kSortTile + warp * (32 * kSortItems);
  uint64_t elem[kSortItems];
  uint32_t rank[kSortItems];
#pragma unroll
  for (uint32_t r = 0; r < kSortItems; r++)
  {
    uint32_t i = base + r * 32 + lane;
    bool valid = i < n;
    elem[r] = valid ? sort_load<kFromData>(in, data, first, i) : 0;
    uint32_t digit = valid ? ((hash20((uint32_t)(elem[r] >> 32)) >> shift) & mask) : 0xffffffffu;
    uint32_t peers = __match_any_sync(0xffffffffu, digit);
    uint32_t leader = (uint32_t)__ffs((int)peers) - 1;
    uint32_t before = (uint32_t)__popc(peers & ((1u << lane) - 1));
    uint32_t start = 0;
    if (valid && lane == leader)
    {
      start = cnt[warp][digit];
      cnt[warp][digit] = start + (uint32_t)__popc(peers);
    }
    start = __shfl_sync(0xffffffffu, start, (int)leader);
    rank[r] = start + before;
    __syncwarp();
  }
  __syncthreads();
  // Per digit: where its elements start inside the tile (lstart) and in the output (gbase), and where each warp's
  // share starts inside the digit.  The tile is first put in digit order in shared memory and then written out
  // with consecutive threads on consecutive addresses: a warp's store covers a few long runs instead of 32
  // scattered 8-byte pieces.
  __shared__ uint32_t lstart[kSortBins], gbase[kSortBins];
  __shared__ uint64_t stage[kSortTile];
  if (threadIdx.x < kSortBins)
  {
    uint32_t run = 0;
#pragma unroll
    for (uint32_t w = 0; w < kSortThreads / 32; w++)
    {
      uint32_t c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = run;
      run += c;
    }
    lstart[threadIdx.x] = run;                                  // for now: the number of elements of this digit in the tile
    gbase[threadIdx.x] = hist_scanned[threadIdx.x * num_tiles + blockIdx.x];
  }
  __syncthreads();
  if (warp == 0)
  {
    // exclusive scan of the 128 counts: four per lane
    uint32_t c[kSortBins / 32], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < kSortBins / 32; k++) { c[k] = lstart[lane * (kSortBins / 32) + k]; sum += c[k]; }
    uint32_t incl = sum;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    uint32_t run = incl - sum;
#pragma unroll
    for (uint32_t k = 0; k < kSortBins / 32; k++) { lstart[lane * (kSortBins / 32) + k] = run; run += c[k]; }
  }
  __syncthreads();
#pragma unroll
  for (uint32_t r = 0; r < kSortItems; r++)
  {
    uint32_t i = base + r * 32 + lane;
    if (i < n)
    {
      uint32_t digit = (hash20((uint32_t)(elem[r] >> 32)) >> shift) & mask;
      stage[lstart[digit] + cnt[warp][digit] + rank[r]] = elem[r];
    }
  }
  __syncthreads();
  const uint32_t tile_base = blockIdx.x * kSortTile;
  const uint32_t tile_n = min((uint32_t)kSortTile, n - tile_base);
  for (uint32_t k = threadIdx.x; k < tile_n; k += kSortThreads)
  {
    const uint64_t e = stage[k];
    const uint32_t digit = (hash20((uint32_t)(e >> 32)) >> shift) & mask;
    out[gbase[digit] + (k - lstart[digit])] = e;
  }
}

}  // namespace sz4
