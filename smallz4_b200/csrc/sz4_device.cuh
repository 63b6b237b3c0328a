// sz4_device.cuh -- geometry of one device batch and small device helpers.
//
// A batch is a contiguous piece of the input stream resident in HBM:
//
//   buf:  [ halo bytes | block 0 | block 1 | ... | block n_blocks-1 ]   (+ zero padding both sides)
//          ^0           ^halo
//
// Positions are 32-bit indices into buf.  The halo is the read-only history in front of the
// first block (>= 65535+12 bytes when the stream has that much; a dictionary prefix for the
// first block of a -D stream, smallz4.h:554-570).  Blocks are LZ4 blocks (smallz4.h:124,127).
#pragma once
#include "sz4_platform.h"

namespace sz4
{
enum : uint32_t
{
  kMinMatch    = 4,       // smallz4.h:95
  kEndNoMatch  = 12,      // smallz4.h:99
  kEndLiterals = 5,       // smallz4.h:101
  kHashBits    = 20,      // smallz4.h:104
  kWindow      = 65535,   // smallz4.h:111 MaxDistance
  kSameLetter  = 19 + 255 * 256,   // smallz4.h:118 MaxSameLetter
  kGreedyMax   = 3,       // smallz4.h:77
  kLazyMax     = 6,       // smallz4.h:79
  kPad         = 256      // zeroed bytes in front of and behind every device array
};

struct Geom
{
  uint32_t n_total;       // halo + all block bytes
  uint32_t halo;          // block 0 starts here
  uint32_t block_size;    // nominal block size (multiple of 65536)
  uint32_t n_blocks;
  uint32_t first_ins;     // first position that the reference inserts into its hash chains
  uint32_t max_chain;     // smallz4::maxChainLength
  uint32_t shift;         // 1 when a dictionary shifts ring reads by one slot (DESIGN.md Q-dict), else 0
  uint8_t  legacy;        // legacy frame: blocks are independent (smallz4.h:783-795)
  uint8_t  stream_first;  // block 0 is the first block of the stream
  uint8_t  stream_last;   // the last block is the last block of the stream
  uint8_t  pad_;
};

__host__ __device__ __forceinline__ uint32_t block_begin(const Geom& g, uint32_t j) { return g.halo + j * g.block_size; }
__host__ __device__ __forceinline__ uint32_t block_end(const Geom& g, uint32_t j)
{
  uint32_t e = g.halo + (j + 1) * g.block_size;
  return e < g.n_total ? e : g.n_total;
}
__host__ __device__ __forceinline__ uint32_t block_len(const Geom& g, uint32_t j) { return block_end(g, j) - block_begin(g, j); }
// one past the last position the match finder visits in block j (i + 12 <= n, smallz4.h:629)
__host__ __device__ __forceinline__ uint32_t search_end(const Geom& g, uint32_t j)
{
  uint32_t n = block_len(g, j);
  return n >= kEndNoMatch ? block_begin(g, j) + n - (kEndNoMatch - 1) : block_begin(g, j);
}

// Position inserted a second time by the next block's lookback (smallz4.h:615-629): block end - 12
// of every block that has a successor.  Its chain entries end up zero for every later reader.
__host__ __device__ __forceinline__ bool is_twice_inserted(const Geom& g, uint32_t p)
{
  if (g.legacy) return false;
  uint32_t q = p + kEndNoMatch;
  if (q < g.halo) return false;
  uint32_t k = (q - g.halo) / g.block_size;
  if (q != g.halo + k * g.block_size) return false;
  if (k == 0) return !g.stream_first && g.halo >= kEndNoMatch;
  // the border behind the last block of the batch belongs to the next batch (its k == 0 case)
  return k < g.n_blocks;
}

// dataZero (smallz4.h:506,801) of the block whose loop inserts position p.
__host__ __device__ __forceinline__ uint32_t floor_of(const Geom& g, uint32_t p)
{
  if (p < g.halo)
  {
    // the last 11 positions in front of block 0 are inserted by block 0's lookback
    if (!g.legacy && !g.stream_first && p + kEndNoMatch > g.halo)
      return g.halo > kWindow ? g.halo - kWindow : 0;
    return 0;
  }
  uint32_t j = (p - g.halo) / g.block_size;
  uint32_t b = block_begin(g, j);
  if (g.legacy) return b;
  uint32_t n = block_len(g, j);
  if (p - b + kEndNoMatch > n && j + 1 < g.n_blocks)      // inserted by block j+1's lookback
  {
    uint32_t nb = block_begin(g, j + 1);
    return nb > kWindow ? nb - kWindow : 0;
  }
  if (j == 0 && g.stream_first) return 0;
  return b > kWindow ? b - kWindow : 0;
}

__host__ __device__ __forceinline__ uint32_t hash20(uint32_t four)   // smallz4.h:164
{
  return ((four * 48271u) >> (32 - kHashBits)) & ((1u << kHashBits) - 1);
}

__device__ __forceinline__ uint32_t ld32u(const uint8_t* p)          // unaligned little-endian load
{
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// extra bytes needed to code a length beyond the 4-bit token field (LZ4 block format)
__host__ __device__ __forceinline__ uint32_t len_ext_bytes(uint32_t v) { return v < 15 ? 0 : 1 + (v - 15) / 255; }

}  // namespace sz4
