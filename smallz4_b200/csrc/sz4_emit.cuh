// sz4_emit.cuh -- phase 3b: token / literal emission (the byte layout of smallz4.h:303-367).
//
// k_path has listed the sequences of each block with the offset of each inside the compressed
// block (a running sum, i.e. the prefix scan of the sequence sizes).  Emission is then
// output-parallel: every thread produces 16 consecutive bytes of the compressed block, finds the
// sequence they belong to by binary search over the offsets, and stores them with one 16-byte
// store -- fully coalesced writes, gathered literal reads.
#pragma once
#include "sz4_parse.cuh"

namespace sz4
{
struct SeqView
{
  uint32_t lit_from, lits, len, dist, out;   // literals, match, where the sequence starts
  uint32_t lit_ext, size;
  bool last;
};

__device__ __forceinline__ SeqView load_seq(const SeqRec* seqs, uint32_t k)
{
  SeqRec r = seqs[k];
  SeqView v;
  v.lit_from = 0;
  if (k > 0) { SeqRec q = seqs[k - 1]; v.lit_from = q.pos + q.len; }
  v.lits = r.pos - v.lit_from;
  v.len = r.len; v.dist = r.dist; v.out = r.out;
  v.last = (r.len == 0);
  v.lit_ext = len_ext_bytes(v.lits);
  v.size = seq_bytes(v.lits, v.len, v.last);
  return v;
}

// byte number r of the length-extension bytes that encode value v (v >= 15): 255, 255, ..., rest
__device__ __forceinline__ uint32_t ext_byte(uint32_t v, uint32_t r)
{
  uint32_t rest = v - 15;
  uint32_t full = rest / 255;
  return r < full ? 255u : rest - full * 255u;
}

__device__ __forceinline__ uint32_t seq_byte(const SeqView& s, uint32_t r, const uint8_t* block_data)
{
  const int ml = s.last ? 0 : (int)s.len - (int)kMinMatch;                 // smallz4.h:304-308
  if (r == 0)
  {
    uint32_t tok = ml < 15 ? ((uint32_t)ml & 0xffu) : 15u;                // smallz4.h:311
    return (s.lits < 15 ? (tok | (s.lits << 4)) : (tok | 0xF0u)) & 0xffu;  // smallz4.h:314-323
  }
  r -= 1;
  if (r < s.lit_ext) return ext_byte(s.lits, r);
  r -= s.lit_ext;
  if (r < s.lits) return block_data[s.lit_from + r];
  r -= s.lits;
  if (r == 0) return s.dist & 0xffu;                                       // smallz4.h:351-352
  if (r == 1) return (s.dist >> 8) & 0xffu;
  return ext_byte((uint32_t)ml, r - 2);                                    // smallz4.h:355-366
}

// where block j's record goes inside the batch's output segment
struct BlockOut { uint32_t off; uint32_t nbytes; uint32_t packed; uint32_t pad_; };

// smallz4.h:765-775: compressed unless that is not smaller (legacy: always compressed); writes the
// 4-byte block sizes and computes where every payload starts.  One thread: n_blocks is small.
__global__ void k_block_offsets(const uint32_t* packed_size, BlockOut* bo, unsigned long long* seg_total, uint8_t* seg, Geom g)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long run = 0;
  for (uint32_t j = 0; j < g.n_blocks; j++)
  {
    const uint32_t len = block_len(g, j), pk = packed_size[j];
    const bool use_packed = g.legacy || pk < len;
    const uint32_t nbytes = use_packed ? pk : len;
    const uint32_t tagged = nbytes | (use_packed ? 0u : 0x80000000u);
    seg[run + 0] = (uint8_t)tagged; seg[run + 1] = (uint8_t)(tagged >> 8);
    seg[run + 2] = (uint8_t)(tagged >> 16); seg[run + 3] = (uint8_t)(tagged >> 24);
    BlockOut o; o.off = (uint32_t)(run + 4); o.nbytes = nbytes; o.packed = use_packed ? 1u : 0u; o.pad_ = 0;
    bo[j] = o;
    run += 4ull + nbytes;
  }
  *seg_total = run;
}

// thread c of block j produces the 16-byte aligned piece number c of the block's payload in the segment
__global__ void __launch_bounds__(256)
k_emit(const uint8_t* data, const SeqRec* seqs, uint32_t seq_stride, const uint32_t* seq_count,
       const BlockOut* bo, uint8_t* seg, uint32_t chunks_per_block, Geom g)
{
  const uint32_t j = blockIdx.x / chunks_per_block;
  const uint32_t c = (blockIdx.x % chunks_per_block) * blockDim.x + threadIdx.x;
  if (j >= g.n_blocks) return;
  const BlockOut b = bo[j];
  const uint32_t first_piece = b.off >> 4;
  const uint64_t piece_lo = ((uint64_t)first_piece + c) << 4;            // segment offsets of this piece
  const uint64_t seg_lo = piece_lo > b.off ? piece_lo : b.off;
  const uint64_t seg_hi = min(piece_lo + 16, (uint64_t)b.off + b.nbytes);
  if (seg_lo >= seg_hi) return;
  uint8_t* dst = seg + seg_lo;
  const uint32_t o0 = (uint32_t)(seg_lo - b.off), end = (uint32_t)(seg_hi - b.off);
  if (!b.packed)
  {
    // stored block (smallz4.h:779): plain copy of the input
    const uint8_t* src = data + block_begin(g, j);
    for (uint32_t o = o0; o < end; o++) dst[o - o0] = src[o];
    return;
  }
  const SeqRec* sq = seqs + (size_t)j * seq_stride;
  const uint32_t cnt = seq_count[j];
  const uint8_t* block_data = data + block_begin(g, j);

  // last sequence whose offset is <= o0
  uint32_t lo = 0, hi = cnt - 1;
  while (lo < hi)
  {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (sq[mid].out <= o0) lo = mid; else hi = mid - 1;
  }
  uint32_t k = lo;
  SeqView s = load_seq(sq, k);
  uint32_t w[4] = { 0, 0, 0, 0 };
  for (uint32_t o = o0; o < end; o++)
  {
    while (o - s.out >= s.size) { k++; s = load_seq(sq, k); }
    uint32_t v = seq_byte(s, o - s.out, block_data);
    w[(o - o0) >> 2] |= v << (8 * ((o - o0) & 3));
  }
  if (end - o0 == 16) *(uint4*)dst = make_uint4(w[0], w[1], w[2], w[3]);
  else for (uint32_t o = o0; o < end; o++) dst[o - o0] = (uint8_t)(w[(o - o0) >> 2] >> (8 * ((o - o0) & 3)));
}

}  // namespace sz4
