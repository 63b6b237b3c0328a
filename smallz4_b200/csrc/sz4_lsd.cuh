// sz4_lsd.cuh -- phase 1 without a dictionary: prefix-class tables for every position, from ONE byte-wise LSD radix sort.
//
// What the reference builds (smallz4.h:645-720) is, per position, the distance to the previous position that starts
// with the same four bytes (previousExact); findLongestMatch (smallz4.h:173-255) then walks that chain and keeps every
// strictly longer candidate.  A candidate can only be longer than a match of L bytes if it shares the first L+1 bytes,
// so the candidates the walk actually keeps are: the nearest position with the same 4 bytes, then the nearest with the
// same (L1+1) bytes, and so on.  With the distance to the previous position of the same K-byte prefix for K = 4..8
// ("pe4" .. "pe8") the match finder jumps from keeper to keeper and only walks a chain -- the 8-byte one -- once the
// match is at least 8 bytes long (sz4_search.cuh).  On the mixed corpus this visits 25 candidates per position instead
// of 170.
//
// All five tables fall out of one sort.  Element "anchor a" is sorted by the bytes in FRONT of it, nearest first:
// pass j (1..8) is a stable counting sort by data[a - j].  After pass j the elements are ordered by the string
// data[a-j .. a-1], ties by a -- that is, positions p = a - j are grouped by their j-byte prefix, each group in
// position order, so the previous position with the same j-byte prefix is simply the left neighbour.  Passes 5..8 read
// the order of passes 4..7 anyway and pick up pe4..pe7 on the way (carried along in `car`); k_lsd_extract reads the
// final order for pe8 and writes the tables out by position.
//
// A pass is one kernel (k_lsd_pass2): global digit offsets come from a byte histogram of the input (the digits of pass
// j are the input bytes shifted by j), the rank of a tile inside its digits from a decoupled look-back over the tiles in
// front of it (each tile publishes its digit counts, then the running sums), and the tile leaves in digit order, so that
// a warp's store covers a few long pieces.  HBM-bound by nature: 24-40 B per element and pass.
#pragma once
#include "sz4_device.cuh"
#include "sz4_sort.cuh"

namespace sz4
{
enum : uint32_t
{
  kLsdTile    = 4096,                      // elements per tile (the regions are multiples of it)
  kLsdBins    = 256,
  kLsdPasses  = 8,
#ifndef SZ4_LSD_CHUNK
#define SZ4_LSD_CHUNK (1u << 20)
#endif
  kLsdChunk   = SZ4_LSD_CHUNK,             // anchors a chunk owns (a multiple of the tile)
  kLsdHalo    = 65536,                     // anchors in front of them that it sorts along: their possible predecessors
  kLsdRegion  = kLsdChunk + kLsdHalo,      // elements per chunk in the buffers (a multiple of the tile)
  kLsdSpin    = 1u << 24                   // polls of one look-back slot before the kernel gives up (never in practice)
};

// The sort is done chunk by chunk: chunk c owns the anchors [c*kLsdChunk, (c+1)*kLsdChunk) and sorts them together with
// the kLsdHalo anchors in front (a predecessor is at most 65535 back), in its own region of the buffers.  The order
// inside a chunk is all the tables need, and everything a chunk's extraction writes lies in a window of kLsdChunk
// positions: the 2- and 8-byte scatters meet in L2 instead of each costing a DRAM sector.
struct LsdGeom
{
  int32_t a0, a1;                          // anchors of the batch: [a0, a1)
  uint32_t chunks;
  __host__ __device__ __forceinline__ int32_t own_lo(uint32_t c) const { const int32_t x = (int32_t)(c * kLsdChunk); return x > a0 ? x : a0; }
  __host__ __device__ __forceinline__ int32_t lo(uint32_t c) const { const int32_t x = (int32_t)(c * kLsdChunk) - (int32_t)kLsdHalo; return x > a0 ? x : a0; }
  __host__ __device__ __forceinline__ int32_t hi(uint32_t c) const { const int64_t x = (int64_t)(c + 1) * kLsdChunk; return x < a1 ? (int32_t)x : a1; }
  __host__ __device__ __forceinline__ uint32_t count(uint32_t c) const { return (uint32_t)(hi(c) - lo(c)); }
};

struct LsdBuf
{
  uint64_t* key;    // data[a-8 .. a-1], data[a-1] in the low byte
  uint32_t* pos;    // a
  uint64_t* car;    // pe4[a-4] | pe5[a-5] << 16 | pe6[a-6] << 32 | pe7[a-7] << 48, as far as known
};

__device__ __forceinline__ uint64_t lsd_key(const uint8_t* data, uint32_t a)
{
  // the eight bytes in front of a, first byte in memory most significant
  const uintptr_t at = (uintptr_t)(data + a) - 8;
  const uint64_t* w = (const uint64_t*)(at & ~(uintptr_t)7);
  const uint32_t sh = (uint32_t)(at & 7) * 8;
  const uint64_t lo = w[0];
  const uint64_t v = sh ? (lo >> sh) | (w[1] << (64 - sh)) : lo;
  const uint32_t a0 = __byte_perm((uint32_t)v, 0, 0x0123), a1 = __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
  return ((uint64_t)a0 << 32) | a1;
}

// ---- digit histograms per chunk: hist[c][v] = number of x in [lo(c) - 1, hi(c) - 8) with data[x] == v (the part all
// passes have in common).  Indices are signed: the first anchors look at the zero padding in front of the batch.
// grid = (number of chunks from chunk0 on) * kLsdHistSplit.
enum : uint32_t { kLsdHistSplit = 8 };
__global__ void __launch_bounds__(256)
k_lsd_hist(const uint8_t* data, LsdGeom lg, uint32_t chunk0, uint32_t* hist)
{
  __shared__ uint32_t h[kLsdBins];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t c = chunk0 + blockIdx.x / kLsdHistSplit, part = blockIdx.x % kLsdHistSplit;
  const int32_t lo = lg.lo(c) - 1, hi = max(lg.hi(c) - 8, lo);
  // 16 aligned bytes per thread and step; equal neighbours are counted in a register first (runs of one byte would
  // otherwise serialise on one counter)
  const int32_t lo16 = lo & ~15;
  const uint32_t steps = hi > lo16 ? (uint32_t)(hi - lo16 + 15) / 16 : 0;
  for (uint32_t s = part * blockDim.x + threadIdx.x; s < steps; s += kLsdHistSplit * blockDim.x)
  {
    const int32_t x0 = lo16 + (int32_t)(s * 16);
    const uint4 q = *(const uint4*)(data + x0);
    const uint32_t w[4] = { q.x, q.y, q.z, q.w };
    uint32_t last = 0, cnt = 0;
#pragma unroll
    for (int32_t k = 0; k < 16; k++)
    {
      const int32_t x = x0 + k;
      if (x < lo || x >= hi) continue;
      const uint32_t b = (w[k >> 2] >> ((k & 3) * 8)) & 255u;
      if (cnt != 0 && b == last) cnt++;
      else { if (cnt) atomicAdd(&h[last], cnt); last = b; cnt = 1; }
    }
    if (cnt) atomicAdd(&h[last], cnt);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[c * kLsdBins + threadIdx.x], h[threadIdx.x]);
}

// one CTA of 256 threads per chunk: bases[c][j-1][v] = number of the chunk's anchors whose digit of pass j (data[a-j])
// is below v.  Pass j looks at data[lo-j, hi-j) = the common part plus j-1 bytes in front and 8-j behind.
__global__ void __launch_bounds__(256)
k_lsd_bases(const uint8_t* data, LsdGeom lg, uint32_t chunk0, const uint32_t* common, uint32_t* bases)
{
  __shared__ uint32_t ws[32], tot;
  const uint32_t c = chunk0 + blockIdx.x;
  const int32_t a0 = lg.lo(c), a1 = lg.hi(c);
  const int32_t clo = a0 - 1, chi = max(a1 - 8, clo);
  for (int32_t j = 1; j <= (int32_t)kLsdPasses; j++)
  {
    uint32_t n = common[c * kLsdBins + threadIdx.x];
    for (int32_t x = a0 - j; x < min(clo, a1 - j); x++) if (data[x] == threadIdx.x) n++;
    for (int32_t x = max(chi, a0 - j); x < a1 - j; x++) if (data[x] == threadIdx.x) n++;
    const uint32_t e = block_excl_scan(n, ws, &tot);
    bases[(c * kLsdPasses + (j - 1)) * kLsdBins + threadIdx.x] = e;
  }
}

// ---- decoupled look-back (one 64-bit word per tile and digit: status | pass tag | value)
__device__ __forceinline__ uint64_t lsd_word(uint32_t status, uint32_t tag, uint32_t value)
{
  return ((uint64_t)status << 62) | ((uint64_t)tag << 56) | value;
}
__device__ __forceinline__ uint64_t lsd_peek(const uint64_t* p)
{
#ifdef SZ4_EMU
  return *p;
#else
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void lsd_post(uint64_t* p, uint64_t v)
{
#ifdef SZ4_EMU
  *p = v;
#else
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// One pass.  A tile of 4096 elements arrives in shared memory by bulk-async copies (cp.async.bulk -> UBLKCP, completion
// on an mbarrier) and is read where the copy put it: ranking gives every element its slot, the inverse (slot ->
// element) is written to shared memory, and the output loop gathers from the staged input and stores with consecutive
// threads on consecutive addresses.  The look-back reads four predecessors per trip.
// ---------------------------------------------------------------------------------------------
enum : uint32_t
{
#ifndef SZ4_LSD2_THREADS
#define SZ4_LSD2_THREADS 512
#endif
#ifndef SZ4_LSD2_CTAS
#define SZ4_LSD2_CTAS 2
#endif
  kLsd2Threads = SZ4_LSD2_THREADS,
  kLsd2Items   = 8,
  kLsd2Tile    = kLsd2Threads * kLsd2Items,                            // elements per CTA
  kLsd2TilesPerChunk = kLsdRegion / kLsd2Tile,
  kLsd2Warps   = kLsd2Threads / 32,
  kLsd2CntBytes = kLsd2Warps * kLsdBins * 4
};
// dynamic shared memory: kHalves input tiles (key, car, pos), slot -> element, per-warp digit counters, digit starts
template <uint32_t kHalves, bool kWithCar, bool kWithPos> struct Lsd2Layout
{
  static constexpr uint32_t key = 0, car = key + kHalves * kLsd2Tile * 8, pos = car + (kWithCar ? kHalves * kLsd2Tile * 8 : 0),
                            src = pos + (kWithPos ? kHalves * kLsd2Tile * 4 : 0), cnt = src + kLsd2Tile * 2, misc = cnt + kLsd2CntBytes,
                            bytes = misc + 2 * kLsdBins * 4 + 64;
};

#ifndef SZ4_EMU
__device__ __forceinline__ uint32_t lsd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lsd_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  for (uint32_t o = 0; o < bytes; o += 16384)
  {
    const uint32_t piece = min(16384u, bytes - o);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(lsd_smem_u32((const unsigned char*)dst + o)), "l"((const unsigned char*)src + o), "r"(piece), "r"(lsd_smem_u32(bar)) : "memory");
  }
}
#endif

// kMode: 0 = pass 1 (elements are made from the data), 1 = passes 2..4 (key, pos), 2 = pass 5 (key, pos in; pe4 starts
// `car`), 3 = passes 6..8 (key, pos, car).
// kPersist: one CTA per SM runs through the tiles, two halves of shared memory, the next tile's copies in flight while
// this one is processed.  Otherwise: one tile per CTA, one half, two CTAs per SM (registers: 64 x 512 x 2).
template <uint32_t kMode, bool kPersist>
__global__ void __launch_bounds__(kLsd2Threads, kPersist ? 1 : SZ4_LSD2_CTAS)
k_lsd_pass2(LsdBuf in, LsdBuf out, const uint8_t* data, LsdGeom lg, uint32_t pass, uint32_t first, const uint32_t* bases,
            uint64_t* tile_state, uint32_t* tile_counter, uint32_t tile0, uint32_t total_tiles, uint32_t* err)
{
  constexpr bool kFirst = kMode == 0, kCar = kMode == 3, kCarOut = kMode >= 2;
  const uint32_t level = kCarOut ? pass - 1 : 0;                     // the table read off the input order (passes 5..8: pe4..pe7)
  SZ4_DYN_SMEM(smem);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t s_next;
  typedef Lsd2Layout<kPersist ? 2 : 1, kCarOut, !kFirst> L;
  uint64_t* s_key = (uint64_t*)(smem + L::key);
  uint64_t* s_car = (uint64_t*)(smem + L::car);
  uint32_t* s_pos = (uint32_t*)(smem + L::pos);
  uint16_t* s_src = (uint16_t*)(smem + L::src);
  uint32_t (*cnt)[kLsdBins] = (uint32_t (*)[kLsdBins])(smem + L::cnt);
  uint32_t* lstart = (uint32_t*)(smem + L::misc);
  uint32_t* gbase = lstart + kLsdBins;

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t shift = (pass - 1) * 8;

  // input of tile `t` -> half `b`.  Sizes are rounded up to 16 bytes (the region has room: it is a multiple of the tile).
  auto fetch = [&](uint32_t t, uint32_t b)
  {
    const uint32_t chunk = t / kLsd2TilesPerChunk, ltile = t % kLsd2TilesPerChunk;
    const uint32_t n = lg.count(chunk), tile_base = ltile * kLsd2Tile;
    if (tile_base >= n) return;                                        // an empty tile: nothing to wait for either
    if (kFirst) return;                                                // (pass 1 makes its elements itself)
    const uint32_t tile_n = (min((uint32_t)kLsd2Tile, n - tile_base) + 3) & ~3u;
    const size_t at = (size_t)chunk * kLsdRegion + tile_base;
#ifndef SZ4_EMU
    const uint32_t bytes = tile_n * (kCar ? 20u : 12u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the half was read and written by plain accesses before
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lsd_smem_u32(&bar[b])), "r"(bytes) : "memory");
    lsd_bulk(s_key + b * kLsd2Tile, in.key + at, tile_n * 8, &bar[b]);
    lsd_bulk(s_pos + b * kLsd2Tile, in.pos + at, tile_n * 4, &bar[b]);
    if (kCar) lsd_bulk(s_car + b * kLsd2Tile, in.car + at, tile_n * 8, &bar[b]);
#else
    memcpy(s_key + b * kLsd2Tile, in.key + at, tile_n * 8);
    memcpy(s_pos + b * kLsd2Tile, in.pos + at, tile_n * 4);
    if (kCar) memcpy(s_car + b * kLsd2Tile, in.car + at, tile_n * 8);
#endif
  };

  if (threadIdx.x == 0)
  {
#ifndef SZ4_EMU
    for (uint32_t b = 0; b < 2; b++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lsd_smem_u32(&bar[b])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    s_next = tile0 + atomicAdd(tile_counter, 1u);                     // tiles [tile0, total_tiles) are handed out in starting order
  }
  __syncthreads();
  uint32_t tile = s_next;
  if (threadIdx.x == 0 && tile < total_tiles) fetch(tile, 0);
  uint32_t phase0 = 0, phase1 = 0;
  uint32_t b = 0;
  while (tile < total_tiles)
  {
    __syncthreads();                                                   // everybody has read s_next; the other half is free
    if (threadIdx.x == 0)
    {
      const uint32_t nx = kPersist ? tile0 + atomicAdd(tile_counter, 1u) : 0xffffffffu;
      s_next = nx;
      if (nx < total_tiles) fetch(nx, b ^ 1);
    }
    for (uint32_t k = threadIdx.x; k < kLsd2Warps * kLsdBins; k += kLsd2Threads) (&cnt[0][0])[k] = 0;
    const uint32_t chunk = tile / kLsd2TilesPerChunk, ltile = tile % kLsd2TilesPerChunk;
    const uint32_t n = lg.count(chunk), tile_base = ltile * kLsd2Tile;
    const bool empty = tile_base >= n;
    const uint32_t tile_n = empty ? 0 : min((uint32_t)kLsd2Tile, n - tile_base);
    const size_t region = (size_t)chunk * kLsdRegion;
    // the element in front of the tile (for the table of the input order): asked for before the wait
    uint64_t fk = 0; uint32_t fp = 0;
    if (kCarOut && threadIdx.x == 0 && !empty && tile_base > 0) { fk = in.key[region + tile_base - 1]; fp = in.pos[region + tile_base - 1]; }
    if (!empty && !kFirst)
    {
#ifndef SZ4_EMU
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "LSD_WAIT_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra LSD_DONE_%=;\n"
          "bra LSD_WAIT_%=;\n"
          "LSD_DONE_%=:\n"
          "}\n" ::"r"(lsd_smem_u32(&bar[b])), "r"(b ? phase1 : phase0) : "memory");
      if (b) phase1 ^= 1; else phase0 ^= 1;
#endif
    }
    __syncthreads();                                                   // counters zeroed (and, emulated, the tile copied)
    const uint64_t* t_key = s_key + b * kLsd2Tile;
    uint64_t* t_car = s_car + b * kLsd2Tile;
    const uint32_t* t_pos = s_pos + b * kLsd2Tile;
    if (!empty)
    {
      // ---- warp w owns elements [w*256, (w+1)*256) of the tile, as 8 rows of 32 in order
      const uint32_t wbase = warp * (32 * kLsd2Items);
      uint64_t key[kLsd2Items];
      uint32_t rank[kLsd2Items];
#pragma unroll
      for (uint32_t r = 0; r < kLsd2Items; r++)
      {
        const uint32_t e = wbase + r * 32 + lane;
        if (kFirst)
        {
          key[r] = e < tile_n ? lsd_key(data, (uint32_t)lg.lo(chunk) + tile_base + e) : 0;
          (s_key + b * kLsd2Tile)[e] = key[r];
        }
        else key[r] = e < tile_n ? t_key[e] : 0;
      }
      // ---- table of the input order: the left neighbour is the previous position with the same `level`-byte prefix
      if (kCarOut)
      {
        const uint32_t keep = 64 - 8 * level;
#pragma unroll
        for (uint32_t r = 0; r < kLsd2Items; r++)
        {
          const uint32_t e = wbase + r * 32 + lane;
          if (e < tile_n)
          {
            uint64_t lk = fk; uint32_t lp = fp;
            if (e > 0) { lk = t_key[e - 1]; lp = t_pos[e - 1]; }
            const uint32_t d = t_pos[e] - lp;
            // both positions (anchor - level) must be ones the reference inserts (>= first)
            const bool hit = (e > 0 || tile_base > 0) && ((key[r] ^ lk) << keep) == 0 && d <= kWindow && lp >= first + level;
            const uint64_t old = kCar ? t_car[e] : 0;
            t_car[e] = old | ((uint64_t)(hit ? d : 0u) << (16 * (level - 4)));
          }
        }
      }
      // ---- rank inside the warp: __match_any_sync ranks equal digits inside a row, per-warp counters rank rows
#pragma unroll
      for (uint32_t r = 0; r < kLsd2Items; r++)
      {
        const uint32_t e = wbase + r * 32 + lane;
        const bool valid = e < tile_n;
        const uint32_t digit = valid ? (uint32_t)(key[r] >> shift) & 255u : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t leader = (uint32_t)__ffs((int)peers) - 1;
        const uint32_t before = (uint32_t)__popc(peers & ((1u << lane) - 1));
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
      // ---- per digit (one thread each): rank of the warps, the tile's count, the look-back over the chunk's tiles in front
      if (threadIdx.x < kLsdBins)
      {
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (uint32_t w = 0; w < kLsd2Warps; w++)
        {
          const uint32_t c = cnt[w][d];
          cnt[w][d] = run;
          run += c;
        }
        lstart[d] = run;                                              // for now: this digit's count in the tile
        uint64_t* mine = tile_state + (size_t)tile * kLsdBins + d;
        uint32_t excl = 0;
        if (ltile == 0) lsd_post(mine, lsd_word(2, pass, run));
        else
        {
          lsd_post(mine, lsd_word(1, pass, run));
          const uint32_t first_tile = tile - ltile;
          bool done = false;
          for (uint32_t t = tile; !done && t > first_tile; )
          {
            // four tiles in front per trip (their words are independent reads)
            uint64_t s[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; u++) s[u] = t >= first_tile + u + 1 ? lsd_peek(tile_state + (size_t)(t - 1 - u) * kLsdBins + d) : 0;
#pragma unroll
            for (uint32_t u = 0; u < 4; u++)
            {
              if (done || t < first_tile + u + 1) continue;
              const uint64_t* there = tile_state + (size_t)(t - 1 - u) * kLsdBins + d;
              uint32_t spins = 0;
              while ((uint32_t)((s[u] >> 56) & 63u) != pass || (s[u] >> 62) == 0)
              {
                if (++spins > kLsdSpin) { *err = 1; done = true; break; }   // a tile in front never published: do not hang the GPU
                s[u] = lsd_peek(there);
              }
              excl += (uint32_t)s[u];
              if ((s[u] >> 62) == 2) done = true;
            }
            t = t >= 4 ? t - 4 : 0;
          }
          lsd_post(mine, lsd_word(2, pass, excl + run));
        }
        gbase[d] = bases[(chunk * kLsdPasses + (pass - 1)) * kLsdBins + d] + excl;
      }
      __syncthreads();
      if (warp == 0)
      {
        // exclusive scan of the 256 counts: eight per lane
        uint32_t c[kLsdBins / 32], sum = 0;
#pragma unroll
        for (uint32_t k = 0; k < kLsdBins / 32; k++) { c[k] = lstart[lane * (kLsdBins / 32) + k]; sum += c[k]; }
        const uint32_t incl = warp_incl_scan(sum, lane);
        uint32_t run = incl - sum;
#pragma unroll
        for (uint32_t k = 0; k < kLsdBins / 32; k++) { lstart[lane * (kLsdBins / 32) + k] = run; run += c[k]; }
      }
      __syncthreads();
      // ---- slot of every element; its inverse goes to shared memory
#pragma unroll
      for (uint32_t r = 0; r < kLsd2Items; r++)
      {
        const uint32_t e = wbase + r * 32 + lane;
        if (e < tile_n)
        {
          const uint32_t digit = (uint32_t)(key[r] >> shift) & 255u;
          s_src[lstart[digit] + cnt[warp][digit] + rank[r]] = (uint16_t)e;
        }
      }
      if (threadIdx.x < kLsdBins) gbase[threadIdx.x] -= lstart[threadIdx.x];   // (unsigned wrap is fine: only the sum with k is used)
      __syncthreads();
      // ---- out, in digit order: consecutive threads on consecutive addresses
      uint64_t* o_key = out.key + region;
      uint32_t* o_pos = out.pos + region;
      uint64_t* o_car = out.car + region;
#pragma unroll
      for (uint32_t m = 0; m < kLsd2Items; m++)
      {
        const uint32_t k = m * kLsd2Threads + threadIdx.x;
        if (k < tile_n)
        {
          const uint32_t e = s_src[k];
          const uint64_t kk = t_key[e];
          const uint32_t digit = (uint32_t)(kk >> shift) & 255u;
          const uint32_t dst = gbase[digit] + k;                      // (gbase: where the digit starts in the output minus where it starts in the tile)
          o_key[dst] = kk;
          o_pos[dst] = kFirst ? (uint32_t)lg.lo(chunk) + tile_base + e : t_pos[e];
          if (kCarOut) o_car[dst] = t_car[e];
        }
      }
    }
    tile = s_next;
    if (kPersist) b ^= 1;
  }
}

// The final order (all eight bytes): pe8 from the left neighbour, and the tables out by position:
//   jump[a] = { pe4[a-4], pe5[a-5], pe6[a-6], pe7[a-7] }  (8 bytes per anchor),  pe8[a-8]  (its own array: the match
//   finder stages a 64 KiB window of it in shared memory), and rank[a-8] = the element's index in the sorted arrays.
// A chunk writes the anchors it owns.
// grid = chunks * kLsdRegion / 256.
__global__ void __launch_bounds__(256)
k_lsd_extract(LsdBuf in, LsdGeom lg, uint32_t first, uint64_t* jump, uint16_t* pe8, uint32_t* rank)
{
  const uint32_t chunk = blockIdx.x / (kLsdRegion / 256);
  const uint32_t i = (blockIdx.x % (kLsdRegion / 256)) * 256 + threadIdx.x;
  const uint32_t n = lg.count(chunk);
  const size_t region = (size_t)chunk * kLsdRegion;
  const uint32_t lane = threadIdx.x & 31;
  const bool valid = i < n;
  const uint64_t key = valid ? in.key[region + i] : 0;
  const uint32_t pos = valid ? in.pos[region + i] : 0;
  uint64_t lk = __shfl_up_sync(0xffffffffu, key, 1);
  uint32_t lp = __shfl_up_sync(0xffffffffu, pos, 1);
  if (lane == 0 && valid && i > 0) { lk = in.key[region + i - 1]; lp = in.pos[region + i - 1]; }
  if (!valid || (int32_t)pos < lg.own_lo(chunk)) return;
  const uint32_t d = pos - lp;
  const bool hit = i > 0 && key == lk && d <= kWindow && lp >= first + 8;
  jump[pos] = in.car[region + i];
  if (pos >= first + 8)
  {
    pe8[pos - 8] = hit ? (uint16_t)d : (uint16_t)0;
    rank[pos - 8] = (uint32_t)(region + i);                         // where k_long finds the position's class
  }
}

}  // namespace sz4
