// sz4_pipeline.cu -- host side of libsmallz4_b200.so: device buffers, the per-batch kernel sequence,
// frame assembly and the C ABI (include/smallz4_b200.h).
//
// Reference control flow being replaced: smallz4::compress(), smallz4.h:476-814.  The reference
// runs three phases per 4 MiB block, sequentially; here a batch of blocks runs each phase as one
// or a few grid-wide launches (DESIGN.md has the kernel table).  There is no CPU implementation of
// any phase in this library: without a CUDA device every entry point returns SZ4_ERR_CUDA.
#include "sz4_platform.h"
#include "sz4_device.cuh"
#include "sz4_sort.cuh"
#include "sz4_lsd.cuh"
#include "sz4_chain.cuh"
#include "sz4_runs.cuh"
#include "sz4_search.cuh"
#include "sz4_parse.cuh"
#include "sz4_emit.cuh"
#include "sz4_scalar.cuh"

#include "../../include/smallz4_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

using namespace sz4;

namespace
{
const uint32_t kBlockModern = 4u << 20;     // smallz4.h:124
const uint32_t kBlockLegacy = 8u << 20;     // smallz4.h:127
const uint32_t kHaloBytes   = 131072;       // >= 65535 + 12, keeps block starts 16-byte aligned
const uint32_t kDictRunLimit = 60000;       // -D: runs from here on can fire the reference's long-run shortcut (65 299) inside the window

struct DevBuf
{
  void*  p = nullptr;
  size_t bytes = 0;
};
}  // namespace

struct sz4_ctx
{
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
  cudaStream_t copy_stream = nullptr;                 // input of the next batch / output of the previous one, next to the kernels
  cudaEvent_t  ev_in = nullptr, ev_out[2] = { nullptr, nullptr }, ev_user = nullptr;
  // the input of a batch may arrive in up to eight pieces (compress_blocks): piece i ends at byte piece_end[i] of the batch
  // buffer and ev_piece[i] fires when it is there; 0 pieces: the caller has made the stream wait for the whole input
  enum { kMaxPieces = 8 };
  cudaEvent_t  ev_piece[kMaxPieces] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  uint32_t     piece_end[kMaxPieces] = { 0, 0, 0, 0, 0, 0, 0, 0 }, n_pieces = 0;
  cudaEvent_t  pev[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  double       phase_ms[7] = { 0, 0, 0, 0, 0, 0, 0 };   // sort, chain, search, fixup, dp, path, emit
  int          profile = 0;
  std::string  err;
  // options
  uint32_t batch_blocks = 64;
  uint32_t block_size_override = 0;
  int      stage_bulk = 1;
  int      debug_keep = 0;
  int      force_scalar = 0;
  int      allow_scalar_dict = 0;   // 1: a -D stream with a run > 60 000 bytes is replayed by one device thread (sz4_scalar.cuh)
  int      lsd_persist = 0;    // 1 = passes 2..8 by persistent CTAs with two tiles in flight (slower: the look-back is exposed)
  uint32_t sm_count = 148;
  int      debug_stop = 0;     // tests: 1 = stop behind phase 1 (the tables stay for sz4_debug_fetch)
  uint32_t fast_hops = 64;     // k_search: at most this many candidates per lane and round in the fast loop ...
  uint32_t dense_a = 256, dense_b = 2048;   // k_search: first-two-hops distance below which a position goes in the first / second pass (0 = one pass)
  uint32_t tail_lanes = 0;    // k_search: with the tile's queue empty, a warp with fewer walking lanes than this hands its walks to k_long (0: never; measured: no gain)
  uint32_t long_age = 8;       // k_search: rounds after which a walk is handed to k_long (tests lower it)
  uint32_t fast_lanes = 8;     // ... which goes on in steps of eight while at least this many lanes are still walking
  // device memory (grow-only)
  DevBuf rank, long_list;                                // sz4_lsd.cuh / k_long: index of every position in the sorted arrays, handed-over walks
  DevBuf tile_queue;                                     // k_start -> k_search: per tile, the positions whose walk goes on
  DevBuf jump, lsd_state, lsd_misc, dbg_pe;              // sz4_lsd.cuh: pe4..pe7 by anchor, look-back words, histograms / counters
  DevBuf greedy_segs;                                  // k_greedy_*: entry / leave of every segment, and the number of second walks
  DevBuf dp_order;                                     // k_dp_spec: its tasks in starting order
  DevBuf tile_order;                                   // k_search: run positions per tile, and the tiles in starting order
  DevBuf data2, seg2;                                  // the other halves of the double-buffered input and output
  DevBuf data, ph, pe, mlen, mdist, scratch, hist, hist_scanned, partials, seqs, seq_count, packed,
         saved_ph, saved_pe, seeds, nseeds, seg, block_out, seg_total, dbg_len, dbg_dist, scalar_state,
         run_fwd, ones_back, flag_last, flag_carry, mfin, dp_tasks, dp_count, dp_states, dp_overlays, dp_redo, dp_reach, seqs_tmp, path_segs;
  uint8_t* h_in[2] = { nullptr, nullptr };       // sz4_lz4: pinned halves for the stream coming in ...
  uint8_t* h_out[2] = { nullptr, nullptr };      // ... and the records going out
  size_t   h_in_bytes = 0, h_out_bytes = 0;
  uint32_t stream_blocks = 32; // sz4_lz4: blocks per batch (bounds the pinned host memory: 2 x 2 x 128 MiB)
  unsigned long long* h_seg_total = nullptr;    // pinned: [0] segment bytes, [1..6] DP counters, [7] sort error flag, [8] longest byte run (-D)
  unsigned long long dp_redos = 0, path_redos = 0;
  bool               dp_ran = false;
  // stats
  double             kernel_ms = 0;
  unsigned long long launches = 0;
  Geom               last_geom;
  bool               last_scalar = false;
  bool               attr_set = false, dp_attr_set = false, lsd_attr_set = false;

  int fail(const char* what, cudaError_t e)
  {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return SZ4_ERR_CUDA;
  }
};

#define CK(call)                                               \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return ctx->fail(#call, e_);        \
  } while (0)

#define LAUNCH(ctx, kernel, grid, block, smem, ...)                        \
  do {                                                                     \
    SZ4_LAUNCH(kernel, grid, block, smem, (ctx)->stream, __VA_ARGS__);     \
    (ctx)->launches++;                                                     \
  } while (0)

static int reserve(sz4_ctx* ctx, DevBuf& b, size_t bytes)
{
  if (b.bytes >= bytes) return SZ4_OK;
  if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
  size_t want = bytes + bytes / 8 + 4096;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) { b.p = nullptr; ctx->err = "cudaMalloc failed"; return SZ4_ERR_NOMEM; }
  b.bytes = want;
  return SZ4_OK;
}
#define PHASE(k) do { if (ctx->profile) CK(cudaEventRecord(ctx->pev[k], ctx->stream)); } while (0)
#define RSV(buf, bytes) do { int r_ = reserve(ctx, ctx->buf, (bytes)); if (r_ != SZ4_OK) return r_; } while (0)

static uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// exclusive scan of n uint32 on the device (in -> out)
static int device_scan(sz4_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n)
{
  uint32_t chunks = div_up(n, kScanChunk);
  RSV(partials, (size_t)chunks * 4 + 64);
  uint32_t* part = (uint32_t*)ctx->partials.p;
  LAUNCH(ctx, k_scan_reduce, chunks, kScanThreads, 0, in, n, part);
  LAUNCH(ctx, k_scan_partials, 1, kScanThreads, 0, part, chunks);
  LAUNCH(ctx, k_scan_apply, chunks, kScanThreads, 0, in, out, n, (const uint32_t*)part);
  return SZ4_OK;
}

// ---------------------------------------------------------------------------------------------
// One batch: ctx->data holds kPad zero bytes, g.n_total input bytes, zero padding.  Produces the
// concatenated block records in ctx->seg and their total size in *ctx->h_seg_total (after a sync).
// ---------------------------------------------------------------------------------------------
// batch_submit enqueues everything on ctx->stream and returns; batch_finish waits for it and reads the sizes.
static int batch_submit(sz4_ctx* ctx, const Geom& g, bool scalar_finder)
{
  const uint32_t N = g.n_total;
  uint8_t* data = (uint8_t*)ctx->data.p + kPad;

  RSV(ph, ((size_t)N + 2 * kPad) * 2);
  RSV(pe, ((size_t)N + 2 * kPad) * 2);
  RSV(mlen, ((size_t)N + kPad) * 4);
  RSV(mdist, ((size_t)N + kPad) * 2);
  // the cost DP's four arrays; before that the ping-pong buffers of the sort (8-byte elements with a dictionary, else
  // key + position + carried tables = 20 bytes per anchor, twice)
  LsdGeom lg;
  lg.a0 = (int32_t)g.first_ins; lg.a1 = (int32_t)N + 5;             // anchors first .. N+4
  lg.chunks = div_up((uint64_t)lg.a1, kLsdChunk);
  const size_t lsd_stride = (size_t)lg.chunks * kLsdRegion;
  RSV(scratch, ((size_t)N + kPad) * 16 + 256 > 2 * 20 * lsd_stride ? ((size_t)N + kPad) * 16 + 256 : 2 * 20 * lsd_stride);
  RSV(saved_ph, (size_t)g.n_blocks * 4 + 64);
  RSV(saved_pe, (size_t)g.n_blocks * 4 + 64);
  RSV(seq_count, (size_t)g.n_blocks * 4 + 64);
  RSV(packed, (size_t)g.n_blocks * 4 + 64);
  RSV(block_out, (size_t)g.n_blocks * sizeof(BlockOut) + 64);
  RSV(seg_total, 64);
  RSV(nseeds, 64);
  const uint32_t max_seeds = N / 65000 + g.n_blocks + 16;
  RSV(seeds, (size_t)max_seeds * sizeof(Seed));
  // a sequence ends with a match of >= 4 bytes (>= 2 with a dictionary, DESIGN.md Q-dict)
  const uint32_t seq_stride = (g.block_size / kPathSeg + 1) * (kPathSeg / (g.shift ? 2 : 4) + 1) + 8;
  RSV(seqs, (size_t)g.n_blocks * seq_stride * sizeof(SeqRec));
  const uint64_t per_block_out = g.legacy ? (uint64_t)g.block_size + g.block_size / 255 + 64 : (uint64_t)g.block_size;
  const uint64_t seg_cap = (uint64_t)g.n_blocks * (per_block_out + 4) + 64;
  RSV(seg, seg_cap);

  uint16_t* ph = (uint16_t*)ctx->ph.p + kPad;
  uint16_t* pe = (uint16_t*)ctx->pe.p + kPad;
  uint32_t* mlen = (uint32_t*)ctx->mlen.p;
  uint16_t* mdist = (uint16_t*)ctx->mdist.p;
  uint32_t* saved_ph = (uint32_t*)ctx->saved_ph.p;
  uint32_t* saved_pe = (uint32_t*)ctx->saved_pe.p;

  ctx->h_seg_total[7] = 0;                                         // error flag of the sort's look-back
  ctx->h_seg_total[8] = 0;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(cudaMemsetAsync(ctx->ph.p, 0, ((size_t)N + 2 * kPad) * 2, ctx->stream));
  CK(cudaMemsetAsync(ctx->pe.p, 0, ((size_t)N + 2 * kPad) * 2, ctx->stream));
  CK(cudaMemsetAsync(mlen, 0, (size_t)N * 4, ctx->stream));
  CK(cudaMemsetAsync(mdist, 0, (size_t)N * 2, ctx->stream));
  PHASE(0);
  // The input may still be on its way, piece by piece (the memsets above did not need it).  wait_pieces(k): the stream
  // goes on when the first k pieces are there.
  uint32_t pieces_waited = 0;
  auto wait_pieces = [&](uint32_t upto) -> int
  {
    for (; pieces_waited < upto && pieces_waited < ctx->n_pieces; pieces_waited++)
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_piece[pieces_waited], 0));
    return SZ4_OK;
  };

  if (scalar_finder)
  {
    { int r = wait_pieces(ctx->n_pieces); if (r != SZ4_OK) return r; }
    // dictionary stream with runs long enough for the reference's long-run shortcut: the ring
    // semantics are replayed literally by one thread (sz4_scalar.cuh; DESIGN.md Q-dict)
    RSV(scalar_state, (sizeof(uint64_t) << kHashBits) + 2 * 65536 * sizeof(uint16_t));
    CK(cudaMemsetAsync(ctx->scalar_state.p, 0xff, sizeof(uint64_t) << kHashBits, ctx->stream));
    LAUNCH(ctx, k_scalar_find, 1, 32, 0, (const uint8_t*)data, (unsigned long long*)ctx->scalar_state.p, ph, pe, mlen, mdist, g);
    PHASE(1); PHASE(2); PHASE(3); PHASE(4);
  }
  else
  {
    // ---- phase 1: the chains.  Without a dictionary: prefix-class tables pe4..pe8 from one byte-wise LSD sort
    // (sz4_lsd.cuh).  With one: previousHash by a sort on the hash, then the reference's walk over the shifted ring.
    const uint32_t first = g.first_ins;
    const uint32_t count = N >= first + 4 ? N - 3 - first : 0;
    const bool jump_tables = g.shift == 0;
    LsdBuf lsd_sorted; lsd_sorted.key = nullptr; lsd_sorted.pos = nullptr; lsd_sorted.car = nullptr;
    if (count > 0)
    {
      if (jump_tables)
      {
        const uint32_t tiles = lg.chunks * kLsd2TilesPerChunk;
        const size_t misc_words = (size_t)lg.chunks * kLsdBins + (size_t)lg.chunks * kLsdPasses * kLsdBins + 64;
        RSV(lsd_state, (size_t)tiles * kLsdBins * 8 + 64);
        RSV(lsd_misc, misc_words * 4);
        RSV(jump, ((size_t)N + 64) * 8);
        RSV(rank, ((size_t)N + 64) * 4);
        uint32_t* common = (uint32_t*)ctx->lsd_misc.p;
        uint32_t* bases = common + (size_t)lg.chunks * kLsdBins;
        uint32_t* counters = bases + (size_t)lg.chunks * kLsdPasses * kLsdBins;   // [0..7] tile counters of the passes, [8] error flag, [9..16] pass 1 by piece
        LsdBuf A, B;
        A.key = (uint64_t*)ctx->scratch.p; A.car = A.key + lsd_stride; A.pos = (uint32_t*)(A.car + lsd_stride);
        B.key = (uint64_t*)(A.pos + lsd_stride); B.car = B.key + lsd_stride; B.pos = (uint32_t*)(B.car + lsd_stride);
        CK(cudaMemsetAsync(ctx->lsd_misc.p, 0, misc_words * 4, ctx->stream));
        CK(cudaMemsetAsync(ctx->lsd_state.p, 0, (size_t)tiles * kLsdBins * 8, ctx->stream));
        if (!ctx->lsd_attr_set)
        {
          CK(cudaFuncSetAttribute(k_lsd_pass2<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<2, false, true>::bytes));
          CK(cudaFuncSetAttribute(k_lsd_pass2<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<2, true, true>::bytes));
          CK(cudaFuncSetAttribute(k_lsd_pass2<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<2, true, true>::bytes));
          CK(cudaFuncSetAttribute(k_lsd_pass2<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<1, false, false>::bytes));
          CK(cudaFuncSetAttribute(k_lsd_pass2<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<1, false, true>::bytes));
          CK(cudaFuncSetAttribute(k_lsd_pass2<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<1, true, true>::bytes));
          CK(cudaFuncSetAttribute(k_lsd_pass2<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lsd2Layout<1, true, true>::bytes));
          ctx->lsd_attr_set = true;
        }
        LsdBuf src = A, dst = A;
        for (uint32_t pass = 1; pass <= kLsdPasses; pass++)
        {
          const uint32_t mode = pass == 1 ? 0u : (pass <= 4 ? 1u : (pass == 5 ? 2u : 3u));
          uint64_t* state = (uint64_t*)ctx->lsd_state.p;
          uint32_t* cnt = counters + (pass - 1);
#define SZ4_PASS2(M, P, GRID, SMEM, T0, T1) LAUNCH(ctx, (k_lsd_pass2<M, P>), GRID, kLsd2Threads, SMEM, src, dst, (const uint8_t*)data, lg, pass, first, (const uint32_t*)bases, state, cnt, T0, T1, counters + 8)
          if (pass == 1)
          {
            // Histogram, digit offsets and pass 1 only look at the data: they run chunk range by chunk range behind the
            // pieces of the input as they arrive, so that most of a host-to-device copy hides behind them.
            const uint32_t np = ctx->n_pieces ? ctx->n_pieces : 1;
            uint32_t c0 = 0;
            for (uint32_t i = 0; i < np; i++)
            {
              { int r = wait_pieces(i + 1); if (r != SZ4_OK) return r; }
              uint32_t c1 = lg.chunks;                                    // the last piece: everything that is left
              if (i + 1 < np)
                for (c1 = c0; c1 < lg.chunks && (uint32_t)lg.hi(c1) + 8 <= ctx->piece_end[i]; ) c1++;   // chunks whose bytes are all there
              if (c1 == c0) continue;
              LAUNCH(ctx, k_lsd_hist, (c1 - c0) * kLsdHistSplit, 256, 0, (const uint8_t*)data, lg, c0, common);
              LAUNCH(ctx, k_lsd_bases, c1 - c0, 256, 0, (const uint8_t*)data, lg, c0, (const uint32_t*)common, bases);
              cnt = counters + 9 + i;
              SZ4_PASS2(0, false, (c1 - c0) * kLsd2TilesPerChunk, (Lsd2Layout<1, false, false>::bytes), c0 * kLsd2TilesPerChunk, c1 * kLsd2TilesPerChunk);
              c0 = c1;
            }
          }
          else if (ctx->lsd_persist)
          {
            const uint32_t grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
            if (mode == 1) SZ4_PASS2(1, true, grid, (Lsd2Layout<2, false, true>::bytes), 0u, tiles);
            else if (mode == 2) SZ4_PASS2(2, true, grid, (Lsd2Layout<2, true, true>::bytes), 0u, tiles);
            else SZ4_PASS2(3, true, grid, (Lsd2Layout<2, true, true>::bytes), 0u, tiles);
          }
          else if (mode == 1) SZ4_PASS2(1, false, tiles, (Lsd2Layout<1, false, true>::bytes), 0u, tiles);
          else if (mode == 2) SZ4_PASS2(2, false, tiles, (Lsd2Layout<1, true, true>::bytes), 0u, tiles);
          else SZ4_PASS2(3, false, tiles, (Lsd2Layout<1, true, true>::bytes), 0u, tiles);
#undef SZ4_PASS2
          src = dst;
          dst = (dst.key == A.key) ? B : A;
        }
        lsd_sorted = src;
        PHASE(1);
        LAUNCH(ctx, k_lsd_extract, lg.chunks * (kLsdRegion / 256), 256, 0, src, lg, first, (uint64_t*)ctx->jump.p, pe, (uint32_t*)ctx->rank.p);
        CK(cudaMemcpyAsync(ctx->h_seg_total + 7, counters + 8, 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->debug_stop)
        {
          CK(cudaStreamSynchronize(ctx->stream));
          ctx->last_geom = g;
          ctx->err = "stopped behind phase 1 (debug_stop)";
          return SZ4_ERR_ARG;
        }
      }
      else
      {
        { int r = wait_pieces(ctx->n_pieces); if (r != SZ4_OK) return r; }
        uint64_t* bufA = (uint64_t*)ctx->scratch.p;
        uint64_t* bufB = bufA + (((size_t)N + kPad + 1) & ~(size_t)1);
        const uint32_t tiles = div_up(count, kSortTile);
        const uint32_t hist_n = tiles * kSortBins;
        RSV(hist, (size_t)hist_n * 4 + 64);
        RSV(hist_scanned, (size_t)hist_n * 4 + 64);
        uint32_t* hist = (uint32_t*)ctx->hist.p;
        uint32_t* hscan = (uint32_t*)ctx->hist_scanned.p;
        const uint32_t shifts[3] = { 0, 7, 14 }, masks[3] = { 127, 127, 63 };
        uint64_t* src = nullptr;
        uint64_t* dst = bufA;
        for (int pass = 0; pass < 3; pass++)
        {
          if (pass == 0)
            LAUNCH(ctx, k_sort_hist<true>, tiles, kSortThreads, 0, (const uint64_t*)nullptr, (const uint8_t*)data, first, count, shifts[0], masks[0], hist, tiles);
          else
            LAUNCH(ctx, k_sort_hist<false>, tiles, kSortThreads, 0, (const uint64_t*)src, (const uint8_t*)data, first, count, shifts[pass], masks[pass], hist, tiles);
          int r = device_scan(ctx, hist, hscan, hist_n);
          if (r != SZ4_OK) return r;
          if (pass == 0)
            LAUNCH(ctx, k_sort_scatter<true>, tiles, kSortThreads, 0, (const uint64_t*)nullptr, dst, (const uint8_t*)data, first, count, shifts[0], masks[0], (const uint32_t*)hscan, tiles);
          else
            LAUNCH(ctx, k_sort_scatter<false>, tiles, kSortThreads, 0, (const uint64_t*)src, dst, (const uint8_t*)data, first, count, shifts[pass], masks[pass], (const uint32_t*)hscan, tiles);
          src = dst;
          dst = (dst == bufA) ? bufB : bufA;
        }
        PHASE(1);
        // the ring is read one slot off (DESIGN.md Q-dict): previousHash as a flat array, then the reference's walk
        LAUNCH(ctx, k_link, div_up(count, 256), 256, 0, (const uint64_t*)src, count, ph, g);
        LAUNCH(ctx, k_twice_save, div_up(g.n_blocks, 64), 64, 0, ph, saved_ph, g);
        LAUNCH(ctx, k_exact_walk, div_up(count, 256), 256, 0, (const uint8_t*)data, (const uint16_t*)ph, (const uint32_t*)saved_ph, pe, first, count, g);
        LAUNCH(ctx, k_twice_save, div_up(g.n_blocks, 64), 64, 0, pe, saved_pe, g);
      }

      // ---- helpers for runs of one byte (sz4_runs.cuh): run_fwd and ones_back from the run structure of the data.
      // The search uses them only with the undisturbed ring (no dictionary); a dictionary stream asks for its longest
      // run among the bytes the reference inserts (see compress_blocks).
      RSV(run_fwd, ((size_t)N + 64) * 4);
      RSV(ones_back, ((size_t)N + 64) * 2);
      {
        const uint32_t rchunks = div_up(N, kRunChunk);
        RSV(flag_last, (size_t)rchunks * 8 + 64);
        RSV(flag_carry, (size_t)rchunks * 8 + 64);
        uint32_t* cf = (uint32_t*)ctx->flag_last.p;
        uint32_t* cl = cf + rchunks;
        uint32_t* pb = (uint32_t*)ctx->flag_carry.p;
        uint32_t* nb = pb + rchunks;
        LAUNCH(ctx, k_run_reduce, rchunks, kRunThreads, 0, (const uint8_t*)data, N, cf, cl);
        LAUNCH(ctx, k_run_carry, 1, 256, 0, (const uint32_t*)cf, (const uint32_t*)cl, rchunks, N, pb, nb);
        LAUNCH(ctx, k_run_apply, rchunks, kRunThreads, 0, (const uint8_t*)data, N, (const uint32_t*)pb, (const uint32_t*)nb,
               (uint32_t*)ctx->run_fwd.p, (uint16_t*)ctx->ones_back.p);
      }
      if (g.shift != 0)
      {
        uint32_t* longest = (uint32_t*)ctx->seg_total.p + 4;
        CK(cudaMemsetAsync(longest, 0, 4, ctx->stream));
        // (only what the reference inserts into its chains counts: the zero padding in front of a short dictionary,
        // smallz4.h:557-563, is never looked at)
        LAUNCH(ctx, k_max_u32, 148 * 4, 256, 0, (const uint32_t*)ctx->run_fwd.p + g.first_ins, N - g.first_ins, longest);
        CK(cudaMemcpyAsync(ctx->h_seg_total + 8, longest, 4, cudaMemcpyDeviceToHost, ctx->stream));
      }
      // ---- phase 2: longest match per position
      const uint32_t tiles_per_block = div_up(g.block_size, kTile);
      if (!ctx->attr_set)
      {
        CK(cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSearchSmem));
        ctx->attr_set = true;
      }
      const uint32_t n_tiles = g.n_blocks * tiles_per_block;
      RSV(tile_order, (size_t)n_tiles * 8 + 64);
      uint32_t* tile_cost = (uint32_t*)ctx->tile_order.p;
      uint32_t* tile_order = tile_cost + n_tiles;
      LAUNCH(ctx, k_tile_cost, n_tiles, 256, 0, (const uint16_t*)pe, tiles_per_block, g, tile_cost);
      LAUNCH(ctx, k_tile_order, 1, 256, 0, (const uint32_t*)tile_cost, n_tiles, tile_order);
      uint32_t* tile_count = nullptr;
      uint32_t* tile_queue = nullptr;
      if (jump_tables)
      {
        // the searches up to a match of eight bytes, one position per thread; the rest is queued per tile
        RSV(tile_queue, ((size_t)n_tiles * kTile + n_tiles) * 4 + 64);
        tile_queue = (uint32_t*)ctx->tile_queue.p;
        tile_count = tile_queue + (size_t)n_tiles * kTile;
        CK(cudaMemsetAsync(tile_count, 0, (size_t)n_tiles * 4, ctx->stream));
        LAUNCH(ctx, k_start, div_up(N - g.halo, 256), 256, 0, (const uint8_t*)data, (const uint16_t*)ctx->jump.p, (const uint16_t*)pe,
               (const uint32_t*)ctx->run_fwd.p, mlen, mdist, tile_count, tile_queue, tiles_per_block, g);
      }
      PHASE(2);
      // the few walks that go on for thousands of candidates leave k_search for k_long (one warp each, over the sorted
      // arrays, which are still in `scratch`)
      const uint32_t long_cap = jump_tables ? N / 16 + 1024 : 0;
      LongWalk* long_list = nullptr;
      uint32_t* long_count = nullptr;
      if (jump_tables)
      {
        RSV(long_list, (size_t)long_cap * sizeof(LongWalk) + 128);
        long_list = (LongWalk*)((uint8_t*)ctx->long_list.p + 64);
        long_count = (uint32_t*)ctx->long_list.p;
        CK(cudaMemsetAsync(long_count, 0, 4, ctx->stream));
      }
      LAUNCH(ctx, k_search, n_tiles, kSearchThreads, kSearchSmem, (const uint8_t*)data, (const uint16_t*)pe,
             (const uint32_t*)saved_pe, (const uint32_t*)ctx->run_fwd.p, (const uint16_t*)ctx->ones_back.p, mlen, mdist,
             tiles_per_block, g, ctx->stage_bulk, ctx->fast_hops, ctx->fast_lanes, ctx->dense_a, ctx->dense_b, (const uint32_t*)tile_order,
             (const uint32_t*)tile_count, (const uint32_t*)tile_queue, long_list, long_count, long_cap, ctx->long_age, ctx->tail_lanes);
      if (jump_tables)
        LAUNCH(ctx, k_long, 148 * 8, 256, 0, (const uint8_t*)data, (const uint64_t*)lsd_sorted.key, (const uint32_t*)lsd_sorted.pos,
               (const uint32_t*)ctx->rank.p, (const LongWalk*)long_list, (const uint32_t*)long_count, long_cap,
               (const uint32_t*)ctx->run_fwd.p, mlen, mdist, (uint32_t)kLsdRegion, g);
      PHASE(3);
      if (g.max_chain <= kLazyMax)
      {
        // greedy / lazy levels: which positions the reference would have searched at all (sz4_search.cuh, k_greedy_*)
        const uint32_t segs_per_block = div_up(g.block_size, kGreedySeg);
        const uint32_t n_segs = g.n_blocks * segs_per_block;
        RSV(greedy_segs, (size_t)n_segs * sizeof(GreedySeg) + 128);
        GreedySeg* gsegs = (GreedySeg*)ctx->greedy_segs.p;
        uint32_t* gredo = (uint32_t*)(gsegs + n_segs);
        CK(cudaMemsetAsync(gredo, 0, 4, ctx->stream));
        // (with the jump tables there is no array of the reference's own previousExact values: a position was searched
        // iff it has an exact predecessor, iff k_search wrote a length)
        const uint16_t* own_pe = jump_tables ? (const uint16_t*)nullptr : (const uint16_t*)pe;
        LAUNCH(ctx, k_greedy_spec, div_up(n_segs, 4), 128, 0, own_pe, (const uint32_t*)saved_pe, mlen, mdist, gsegs, segs_per_block, g);
        LAUNCH(ctx, k_greedy_join, div_up(g.n_blocks, 4), 128, 0, own_pe, (const uint32_t*)saved_pe, mlen, mdist, gsegs, segs_per_block, gredo, g);
        LAUNCH(ctx, k_greedy_apply, div_up(n_segs, 4), 128, 0, own_pe, (const uint32_t*)saved_pe, mlen, mdist, (const GreedySeg*)gsegs, segs_per_block, g);
      }
      else
      {
        CK(cudaMemsetAsync(ctx->nseeds.p, 0, 4, ctx->stream));
        LAUNCH(ctx, k_seed_detect, div_up(N - g.halo, 256), 256, 0, (const uint8_t*)data, (const uint32_t*)mlen, (const uint16_t*)mdist,
               (Seed*)ctx->seeds.p, (uint32_t*)ctx->nseeds.p, max_seeds, g);
        LAUNCH(ctx, k_seed_fix, max_seeds, 256, 0, mlen, mdist, (const Seed*)ctx->seeds.p, (const uint32_t*)ctx->nseeds.p, g);
      }
      PHASE(4);
    }
    else { PHASE(1); PHASE(2); PHASE(3); PHASE(4); }
  }

  { int r = wait_pieces(ctx->n_pieces); if (r != SZ4_OK) return r; }        // (inputs too small for phase 1 come by here)
  if (ctx->debug_keep)
  {
    RSV(dbg_pe, (size_t)N * 2 + 64);
    if (g.shift == 0 && N >= g.first_ins + 4)
      LAUNCH(ctx, k_debug_own4, div_up(N, 256), 256, 0, (const uint8_t*)data, (const uint16_t*)ctx->jump.p, (uint16_t*)ctx->dbg_pe.p, g);
    else
      CK(cudaMemcpyAsync(ctx->dbg_pe.p, pe, (size_t)N * 2, cudaMemcpyDeviceToDevice, ctx->stream));
    RSV(dbg_len, (size_t)N * 4 + 64);
    RSV(dbg_dist, (size_t)N * 2 + 64);
    CK(cudaMemcpyAsync(ctx->dbg_len.p, mlen, (size_t)N * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dbg_dist.p, mdist, (size_t)N * 2, cudaMemcpyDeviceToDevice, ctx->stream));
  }

  // ---- phase 3: cost DP (levels > 3), parse walk, emission
  DpScratch dp;
  dp.cost = (uint32_t*)ctx->scratch.p;
  dp.st5 = dp.cost + N + 64;
  dp.st6 = dp.st5 + N + 64;
  dp.st7 = dp.st6 + N + 64;
  const uint32_t* final_len = mlen;
  ctx->dp_ran = false;
  RSV(dp_redo, 64);
  CK(cudaMemsetAsync(ctx->dp_redo.p, 0, 64, ctx->stream));
  if (g.max_chain > kGreedyMax)                                   // smallz4.h:755
  {
    if (!ctx->dp_attr_set)
    {
      CK(cudaFuncSetAttribute(k_dp_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDpSmemSpec));
      CK(cudaFuncSetAttribute(k_dp_verify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * kDpSmem)));
      ctx->dp_attr_set = true;
    }
    const uint32_t max_seg = g.block_size / kDpSeg + 2;
    const uint32_t n_tasks = g.n_blocks * max_seg;
    RSV(mfin, ((size_t)N + kPad) * 4);
    RSV(dp_tasks, (size_t)n_tasks * sizeof(DpTask) + 64);
    RSV(dp_count, (size_t)g.n_blocks * 4 + 64);
    RSV(dp_states, (size_t)n_tasks * sizeof(DpState) + 64);
    RSV(dp_overlays, (size_t)n_tasks * 4 * kDpOvl * 4 + 64);
    uint32_t* mfin = (uint32_t*)ctx->mfin.p;
    // positions the DP does not price (the last five of a block, blocks of <= 12 bytes) keep what was found
    CK(cudaMemcpyAsync(mfin, mlen, (size_t)N * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    const uint32_t groups_per_block = g.block_size / 32;
    const size_t n_groups = (size_t)g.n_blocks * groups_per_block, n_chunks = n_groups / 32;
    RSV(dp_reach, 5 * (n_groups * 4 + 64) + 3 * (n_chunks * 4 + 64));
    uint32_t* reach_all = (uint32_t*)ctx->dp_reach.p;
    uint32_t* reach_nf = reach_all + n_groups + 16;
    uint32_t* run_ends = reach_nf + n_groups + 16;
    uint32_t* first_cand = run_ends + n_groups + 16;
    uint32_t* reach_before = first_cand + n_groups + 16;
    uint32_t* c_reach = reach_before + n_groups + 16;
    uint32_t* c_reach_nf = c_reach + n_chunks + 16;
    uint32_t* c_first = c_reach_nf + n_chunks + 16;
    LAUNCH(ctx, k_dp_group_reach, div_up((uint64_t)n_chunks * 32, 256), 256, 0, (const uint32_t*)mlen, (const uint16_t*)mdist,
           reach_all, reach_nf, run_ends, first_cand, c_reach, c_reach_nf, c_first, groups_per_block, g);
    LAUNCH(ctx, k_dp_chunk_scan, div_up(g.n_blocks, 4), 128, 0, c_reach, c_reach_nf, c_first, groups_per_block / 32, g);
    LAUNCH(ctx, k_dp_reach_before, div_up((uint64_t)n_chunks * 32, 256), 256, 0, (const uint32_t*)reach_all, (const uint32_t*)c_reach,
           reach_before, groups_per_block, g);
    DpPlanIn plan_in;
    plan_in.group_reach = reach_all; plan_in.group_reach_nf = reach_nf; plan_in.group_run_end = run_ends; plan_in.group_first = first_cand;
    plan_in.chunk_reach = c_reach; plan_in.chunk_reach_nf = c_reach_nf; plan_in.chunk_first = c_first;
    LAUNCH(ctx, k_dp_plan, div_up(g.n_blocks, 4), 128, 0, plan_in, groups_per_block, (DpTask*)ctx->dp_tasks.p, (uint32_t*)ctx->dp_count.p,
           max_seg, g);
    RSV(dp_order, (size_t)n_tasks * 4 + 64);
    LAUNCH(ctx, k_dp_task_order, 1, 256, 0, (const DpTask*)ctx->dp_tasks.p, (const uint32_t*)ctx->dp_count.p, max_seg, n_tasks,
           (uint32_t*)ctx->dp_order.p);
    LAUNCH(ctx, k_dp_spec, n_tasks, 32, kDpSmemSpec, (const uint32_t*)mlen, (const uint16_t*)mdist, mfin, dp, (const DpTask*)ctx->dp_tasks.p,
           (const uint32_t*)ctx->dp_count.p, (DpState*)ctx->dp_states.p, (uint32_t*)ctx->dp_overlays.p, max_seg, (uint32_t*)ctx->dp_redo.p, (const uint32_t*)ctx->dp_order.p, g);
    LAUNCH(ctx, k_dp_verify, div_up(g.n_blocks, 4), 128, 4 * kDpSmem, (const uint32_t*)mlen, (const uint16_t*)mdist, mfin, dp, (const DpTask*)ctx->dp_tasks.p,
           (const uint32_t*)ctx->dp_count.p, (DpState*)ctx->dp_states.p, (uint32_t*)ctx->dp_overlays.p, max_seg,
           (uint32_t*)ctx->dp_redo.p, ctx->debug_keep == 0, (const uint32_t*)reach_before, groups_per_block, g);
    if (ctx->debug_keep)
      LAUNCH(ctx, k_dp_cost_fix, n_tasks, 256, 0, dp, (const DpTask*)ctx->dp_tasks.p, (const uint32_t*)ctx->dp_count.p,
             (const DpState*)ctx->dp_states.p, max_seg, g);
    final_len = mfin;
    ctx->dp_ran = true;
  }
  PHASE(5);
  {
    const uint32_t min_len = g.shift ? 2 : 4;
    const uint32_t max_pseg = g.block_size / kPathSeg + 1;
    RSV(seqs_tmp, (size_t)g.n_blocks * seq_stride * sizeof(SeqRec));
    RSV(path_segs, (size_t)g.n_blocks * max_pseg * sizeof(PathSeg) + 64);
    LAUNCH(ctx, k_path_spec, div_up(g.n_blocks * max_pseg, 4), 128, 0, final_len, (const uint16_t*)mdist, (SeqRec*)ctx->seqs_tmp.p, seq_stride,
           (PathSeg*)ctx->path_segs.p, max_pseg, min_len, g);
    LAUNCH(ctx, k_path_join, div_up(g.n_blocks, 4), 128, 0, final_len, (const uint16_t*)mdist, (SeqRec*)ctx->seqs_tmp.p, seq_stride,
           (PathSeg*)ctx->path_segs.p, max_pseg, min_len, (uint32_t*)ctx->seq_count.p, (uint32_t*)ctx->dp_redo.p, g);
    LAUNCH(ctx, k_path_compact, g.n_blocks * max_pseg, 128, 0, (const SeqRec*)ctx->seqs_tmp.p, (SeqRec*)ctx->seqs.p, seq_stride,
           (const PathSeg*)ctx->path_segs.p, max_pseg, min_len, (const uint32_t*)ctx->seq_count.p, g);
  }
  LAUNCH(ctx, k_seq_scan, g.n_blocks, 1024, 0, (SeqRec*)ctx->seqs.p, seq_stride, (const uint32_t*)ctx->seq_count.p,
         (uint32_t*)ctx->packed.p, g);
  PHASE(6);
  LAUNCH(ctx, k_block_offsets, 1, 32, 0, (const uint32_t*)ctx->packed.p, (BlockOut*)ctx->block_out.p,
         (unsigned long long*)ctx->seg_total.p, (uint8_t*)ctx->seg.p, g);
  const uint32_t chunks = div_up(per_block_out / 16 + 2, 256);
  LAUNCH(ctx, k_emit, g.n_blocks * chunks, 256, 0, (const uint8_t*)data, (const SeqRec*)ctx->seqs.p, seq_stride,
         (const uint32_t*)ctx->seq_count.p, (const BlockOut*)ctx->block_out.p, (uint8_t*)ctx->seg.p, chunks, g);
  PHASE(7);
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_seg_total, ctx->seg_total.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_seg_total + 1, ctx->dp_redo.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->last_geom = g;
  ctx->last_scalar = scalar_finder;
  return SZ4_OK;
}

static int batch_finish(sz4_ctx* ctx)
{
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  if ((uint32_t)ctx->h_seg_total[7] != 0) { ctx->err = "k_lsd_pass: a tile waited for its predecessors in vain"; return SZ4_ERR_CUDA; }
  if (ctx->last_geom.shift != 0 && !ctx->last_scalar && (uint32_t)ctx->h_seg_total[8] >= kDictRunLimit)
  {
    // smallz4.h:632-643 meets the ring a dictionary shifts by one slot (DESIGN.md Q-dict): the parallel match finder is
    // not exact there, and the reference's own output for such a stream does not decode.  Opt in for the exact replay.
    ctx->err = "dictionary stream with a run of 60 000 or more equal bytes: set option allow_scalar_dict=1 for the "
               "single-thread replay of the reference's ring (slow), or compress without -D";
    return SZ4_ERR_ARG;
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->kernel_ms += ms;
  ctx->dp_redos += (uint32_t)ctx->h_seg_total[1];
  ctx->path_redos += (uint32_t)(ctx->h_seg_total[1] >> 32);
  if (ctx->profile)
    for (int k = 0; k < 7; k++)
    {
      float pm = 0;
      CK(cudaEventElapsedTime(&pm, ctx->pev[k], ctx->pev[k + 1]));
      ctx->phase_ms[k] += pm;
    }
  return SZ4_OK;
}

// ---------------------------------------------------------------------------------------------
// Stream driver: cuts [prefix | input] into batches of whole blocks with their halo
// ---------------------------------------------------------------------------------------------
struct StreamJob
{
  const uint8_t* src = nullptr;     // halo_in_src bytes of history, then n bytes of blocks
  bool   src_on_device = false;
  size_t halo_in_src = 0;
  size_t n = 0;
  bool   first = true, last = true;
  const uint8_t* dict = nullptr;    // only with first
  size_t dict_len = 0;
  uint8_t* dst = nullptr;
  bool   dst_on_device = false;
  size_t cap = 0;
  uint32_t max_chain = 65535;
  bool   legacy = false;
};

static bool has_long_run(const uint8_t* p, size_t n, size_t limit)
{
  size_t run = 1;
  for (size_t i = 1; i < n; i++)
  {
    if (p[i] == p[i - 1]) { if (++run >= limit) return true; }
    else run = 1;
  }
  return false;
}

static int compress_blocks(sz4_ctx* ctx, const StreamJob& job, size_t* out_len)
{
  const uint32_t bs = ctx->block_size_override ? ctx->block_size_override : (job.legacy ? kBlockLegacy : kBlockModern);
  if (bs % 65536 != 0 || bs < 131072) { ctx->err = "block_size must be a multiple of 65536 and >= 131072"; return SZ4_ERR_ARG; }
  const cudaMemcpyKind in_kind = job.src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out_kind = job.dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;

  // dictionary prefix (smallz4.h:554-570; the CLI keeps the last 65536 bytes, smallz4.cpp:291-302)
  std::vector<uint8_t> prefix;
  uint32_t first_ins = 0;
  const uint8_t* dict = job.dict;
  size_t dict_len = job.first && dict ? job.dict_len : 0;
  if (dict_len > 65536) { dict += dict_len - 65536; dict_len = 65536; }
  const bool with_dict = dict_len > 0;
  if (with_dict)
  {
    if (job.legacy) { ctx->err = "legacy format does not support dictionaries"; return SZ4_ERR_ARG; }   // smallz4.cpp:275
    if (job.src_on_device) { ctx->err = "dictionaries are supported through the host entry points only"; return SZ4_ERR_ARG; }
    prefix.assign(kWindow, 0);
    size_t keep = dict_len < (size_t)kWindow ? dict_len : (size_t)kWindow;
    memcpy(prefix.data() + kWindow - keep, dict + dict_len - keep, keep);
    first_ins = dict_len >= 65536 ? 0 : (uint32_t)(kWindow - keep);
  }
  // With a dictionary the parallel finder is exact only while the long-run shortcut (smallz4.h:632) cannot fire.  By
  // default such a stream is refused (batch_finish looks at the longest run the device found); with the option
  // allow_scalar_dict the host decides up front and one device thread replays the reference's ring literally.
  bool scalar = false;
  if (with_dict && (ctx->allow_scalar_dict || ctx->force_scalar))
    scalar = ctx->force_scalar || has_long_run(job.src, job.n, kDictRunLimit) || has_long_run(prefix.data(), prefix.size(), kDictRunLimit);

  const uint64_t blocks_total = (job.n + bs - 1) / bs;
  uint32_t per_batch = ctx->batch_blocks ? ctx->batch_blocks : 64;
  const uint64_t max_batch_bytes = 1ull << 30;
  if ((uint64_t)per_batch * bs > max_batch_bytes) per_batch = (uint32_t)(max_batch_bytes / bs);
  if (scalar) per_batch = (uint32_t)blocks_total;                 // ring state lives in one launch
  if (scalar && job.n + kWindow > max_batch_bytes) { ctx->err = "dictionary stream with long runs is limited to 1 GiB"; return SZ4_ERR_ARG; }

  // Batches are double-buffered: while the kernels of batch k run on ctx->stream, ctx->copy_stream brings in the
  // input of batch k+1 and takes out the records of batch k-1.  batch_submit always works on ctx->data / ctx->seg;
  // the halves are swapped around it.  The input itself comes in pieces, and the first kernels follow the pieces.
  auto geometry = [&](uint64_t kb, Geom& g, size_t& pay_lo, size_t& pay_hi, bool& batch_first)
  {
    const uint64_t ke = kb + per_batch < blocks_total ? kb + per_batch : blocks_total;
    pay_lo = (size_t)(kb * bs);
    pay_hi = (size_t)(ke * bs) < job.n ? (size_t)(ke * bs) : job.n;
    batch_first = job.first && kb == 0;
    size_t halo;
    if (batch_first) halo = with_dict ? kWindow : 0;
    else if (job.legacy) halo = 0;
    else
    {
      size_t avail = pay_lo + job.halo_in_src;                     // history present in src
      halo = avail < kHaloBytes ? avail : kHaloBytes;
    }
    memset(&g, 0, sizeof(g));
    g.halo = (uint32_t)halo;
    g.n_total = (uint32_t)(halo + (pay_hi - pay_lo));
    g.block_size = bs;
    g.n_blocks = (uint32_t)(ke - kb);
    g.first_ins = batch_first ? first_ins : 0;
    g.max_chain = job.max_chain;
    g.shift = with_dict ? 1 : 0;
    g.legacy = job.legacy ? 1 : 0;
    g.stream_first = batch_first ? 1 : 0;
    g.stream_last = (job.last && ke == blocks_total) ? 1 : 0;
  };
  // input of the batch that starts at block kb -> ctx->data2, on the copy stream
  auto fetch = [&](uint64_t kb) -> int
  {
    Geom g; size_t pay_lo, pay_hi; bool batch_first;
    geometry(kb, g, pay_lo, pay_hi, batch_first);
    RSV(data2, (size_t)g.n_total + 2 * kPad + 64);
    uint8_t* d = (uint8_t*)ctx->data2.p;
    CK(cudaMemsetAsync(d, 0, kPad, ctx->copy_stream));
    CK(cudaMemsetAsync(d + kPad + g.n_total, 0, kPad + 64, ctx->copy_stream));
    if (batch_first && with_dict)
    {
      CK(cudaMemcpyAsync(d + kPad, prefix.data(), kWindow, cudaMemcpyHostToDevice, ctx->copy_stream));
      CK(cudaMemcpyAsync(d + kPad + kWindow, job.src + job.halo_in_src + pay_lo, pay_hi - pay_lo, in_kind, ctx->copy_stream));
    }
    else
    {
      // in up to eight pieces that end on chunk borders of the sort: its first kernels follow the pieces (batch_submit)
      const size_t all = g.halo + (pay_hi - pay_lo);
      const uint8_t* from = job.src + job.halo_in_src + pay_lo - g.halo;
      // (a device-to-device copy is over in no time: one piece)
      const uint32_t np = (!job.src_on_device && all >= 16 * (size_t)kLsdChunk) ? 4u : 1u;
      size_t at = 0;
      for (uint32_t i = 0; i < np; i++)
      {
        const size_t end = i + 1 == np ? all : (all / np * (i + 1)) / kLsdChunk * kLsdChunk;
        CK(cudaMemcpyAsync(d + kPad + at, from + at, end - at, in_kind, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_piece[i], ctx->copy_stream));
        ctx->piece_end[i] = (uint32_t)end;
        at = end;
      }
      ctx->n_pieces = np;
      return SZ4_OK;
    }
    CK(cudaEventRecord(ctx->ev_piece[0], ctx->copy_stream));
    ctx->piece_end[0] = g.n_total;
    ctx->n_pieces = 1;
    return SZ4_OK;
  };

  size_t o = 0;
  int slot = 0;
  bool out_pending[2] = { false, false };
  { int r = fetch(0); if (r != SZ4_OK) return r; }
  for (uint64_t kb = 0; kb < blocks_total; kb += per_batch)
  {
    Geom g; size_t pay_lo, pay_hi; bool batch_first;
    geometry(kb, g, pay_lo, pay_hi, batch_first);
    std::swap(ctx->data, ctx->data2);                              // ctx->data: this batch's input, on its way
    // ctx->seg is about to be written again: its last copy-out (two batches ago) must be over
    if (out_pending[slot]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[slot], 0));

    int r = batch_submit(ctx, g, scalar);                          // (waits for the pieces of its input as it needs them)
    if (r != SZ4_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return r; }
    if (kb + per_batch < blocks_total)
    {
      // ctx->data2 held the previous batch's input, whose kernels have finished
      r = fetch(kb + per_batch);
      if (r != SZ4_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return r; }
    }
    r = batch_finish(ctx);
    if (r != SZ4_OK) { cudaStreamSynchronize(ctx->copy_stream); return r; }
    const size_t seg_len = (size_t)*ctx->h_seg_total;
    if (o + seg_len > job.cap) { cudaStreamSynchronize(ctx->copy_stream); ctx->err = "destination too small"; return SZ4_ERR_DST_SMALL; }
    CK(cudaMemcpyAsync(job.dst + o, ctx->seg.p, seg_len, out_kind, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_out[slot], ctx->copy_stream));
    out_pending[slot] = true;
    std::swap(ctx->seg, ctx->seg2);
    slot ^= 1;
    o += seg_len;
  }
  CK(cudaStreamSynchronize(ctx->copy_stream));
  *out_len = o;
  return SZ4_OK;
}

// ---------------------------------------------------------------------------------------------
// sz4_lz4 without a dictionary: the stream is pulled and pushed batch by batch with bounded memory, like the reference
// (smallz4.h:574-585 pulls 64 KiB requests, :770-780 pushes every block as it is finished, :798-804 keeps 64 KiB of
// history).  Two pinned halves each way: while the kernels of batch k run, the host pushes the records of batch k-1
// through send_bytes and pulls batch k+1 through get_bytes; the copies ride on ctx->copy_stream.
// ---------------------------------------------------------------------------------------------
static int compress_stream(sz4_ctx* ctx, sz4_get_bytes get_bytes, sz4_send_bytes send_bytes, void* user, uint32_t max_chain, bool legacy)
{
  const uint32_t bs = ctx->block_size_override ? ctx->block_size_override : (legacy ? kBlockLegacy : kBlockModern);
  if (bs % 65536 != 0 || bs < 131072) { ctx->err = "block_size must be a multiple of 65536 and >= 131072"; return SZ4_ERR_ARG; }
  uint32_t per_batch = ctx->stream_blocks < ctx->batch_blocks ? ctx->stream_blocks : ctx->batch_blocks;
  if (per_batch < 1) per_batch = 1;
  const size_t B = (size_t)per_batch * bs;
  const size_t in_bytes = kHaloBytes + B, out_bytes = B + B / 255 + (size_t)per_batch * 68 + 64;
  if (ctx->h_in_bytes < in_bytes || ctx->h_out_bytes < out_bytes)
  {
    for (int k = 0; k < 2; k++)
    {
      if (ctx->h_in[k]) cudaFreeHost(ctx->h_in[k]);
      if (ctx->h_out[k]) cudaFreeHost(ctx->h_out[k]);
      ctx->h_in[k] = ctx->h_out[k] = nullptr;
    }
    ctx->h_in_bytes = ctx->h_out_bytes = 0;
    for (int k = 0; k < 2; k++)
      if (cudaMallocHost((void**)&ctx->h_in[k], in_bytes) != cudaSuccess || cudaMallocHost((void**)&ctx->h_out[k], out_bytes) != cudaSuccess)
      { ctx->err = "cudaMallocHost failed"; return SZ4_ERR_NOMEM; }
    ctx->h_in_bytes = in_bytes; ctx->h_out_bytes = out_bytes;
  }
  // pull one batch into the half `slot` (behind the room for its history); 64 KiB requests, 0 bytes = end of input
  auto pull = [&](int slot) -> size_t
  {
    uint8_t* at = ctx->h_in[slot] + kHaloBytes;
    size_t have = 0;
    while (have < B)
    {
      const size_t want = B - have < 65536 ? B - have : 65536;
      const size_t got = get_bytes(at + have, want, user);
      if (got == 0) break;
      have += got;
    }
    return have;
  };

  int slot = 0;
  size_t n = pull(0), prev_n = 0;
  bool out_pending = false;
  int out_slot = 0;
  size_t out_len = 0;
  uint64_t batch = 0;
  while (n > 0)
  {
    const size_t halo = (batch == 0 || legacy) ? 0 : kHaloBytes;          // (every batch but the last is B >= 128 KiB long)
    Geom g;
    memset(&g, 0, sizeof(g));
    g.halo = (uint32_t)halo; g.n_total = (uint32_t)(halo + n); g.block_size = bs; g.n_blocks = (uint32_t)((n + bs - 1) / bs);
    g.max_chain = max_chain; g.legacy = legacy ? 1 : 0; g.stream_first = batch == 0 ? 1 : 0;
    // input of this batch -> ctx->data2 (its last reader, batch k-2, is long done)
    RSV(data2, (size_t)g.n_total + 2 * kPad + 64);
    uint8_t* d = (uint8_t*)ctx->data2.p;
    CK(cudaMemsetAsync(d, 0, kPad, ctx->copy_stream));
    CK(cudaMemsetAsync(d + kPad + g.n_total, 0, kPad + 64, ctx->copy_stream));
    CK(cudaMemcpyAsync(d + kPad, ctx->h_in[slot] + kHaloBytes - halo, halo + n, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_in, ctx->copy_stream));
    // the batch in front: wait for its kernels, start its records on their way out
    if (batch > 0)
    {
      int r = batch_finish(ctx);
      if (r != SZ4_OK) { cudaStreamSynchronize(ctx->copy_stream); return r; }
      out_len = (size_t)*ctx->h_seg_total;
      if (out_len > ctx->h_out_bytes) { cudaStreamSynchronize(ctx->copy_stream); ctx->err = "internal: records larger than their bound"; return SZ4_ERR_DST_SMALL; }
      out_slot = (int)((batch - 1) & 1);
      CK(cudaMemcpyAsync(ctx->h_out[out_slot], ctx->seg.p, out_len, cudaMemcpyDeviceToHost, ctx->copy_stream));
      CK(cudaEventRecord(ctx->ev_out[0], ctx->copy_stream));
      out_pending = true;
      std::swap(ctx->seg, ctx->seg2);                                 // this batch writes the other half
    }
    std::swap(ctx->data, ctx->data2);
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in, 0));
    ctx->n_pieces = 0;
    int r = batch_submit(ctx, g, false);
    if (r != SZ4_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return r; }
    // while the kernels run: push the previous records, pull the next batch
    if (out_pending)
    {
      CK(cudaEventSynchronize(ctx->ev_out[0]));
      send_bytes(ctx->h_out[out_slot], out_len, user);
      out_pending = false;
    }
    prev_n = n;
    const int next = slot ^ 1;
    size_t next_n = 0;
    if (n == B)
    {
      // history for the next batch: the last 128 KiB of this one, in front of where its bytes go
      // (the H2D copy of this batch out of h_in[slot] only reads)
      if (!legacy) memcpy(ctx->h_in[next], ctx->h_in[slot] + kHaloBytes + B - kHaloBytes, kHaloBytes);
      next_n = pull(next);
    }
    slot = next; n = next_n; batch++;
  }
  (void)prev_n;
  if (batch > 0)
  {
    int r = batch_finish(ctx);
    if (r != SZ4_OK) return r;
    out_len = (size_t)*ctx->h_seg_total;
    if (out_len > ctx->h_out_bytes) { ctx->err = "internal: records larger than their bound"; return SZ4_ERR_DST_SMALL; }
    CK(cudaMemcpyAsync(ctx->h_out[0], ctx->seg.p, out_len, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    send_bytes(ctx->h_out[0], out_len, user);
  }
  return SZ4_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* sz4_version(void) { return "1.5-b200.1"; }

int sz4_create(sz4_ctx** out, int device)
{
  if (!out) return SZ4_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
  {
    fprintf(stderr, "smallz4_b200: no usable CUDA device (%s); this library has no CPU path\n",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return SZ4_ERR_CUDA;
  }
  sz4_ctx* ctx = new sz4_ctx();
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->pev[0]) != cudaSuccess || cudaEventCreate(&ctx->pev[1]) != cudaSuccess ||
      cudaEventCreate(&ctx->pev[2]) != cudaSuccess || cudaEventCreate(&ctx->pev[3]) != cudaSuccess ||
      cudaEventCreate(&ctx->pev[4]) != cudaSuccess || cudaEventCreate(&ctx->pev[5]) != cudaSuccess ||
      cudaEventCreate(&ctx->pev[6]) != cudaSuccess || cudaEventCreate(&ctx->pev[7]) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_out[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_out[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_user, cudaEventDisableTiming) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_seg_total, 128) != cudaSuccess)
  {
    fprintf(stderr, "smallz4_b200: cannot initialise CUDA device %d\n", device);
    delete ctx;
    return SZ4_ERR_CUDA;
  }
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0) ctx->sm_count = (uint32_t)prop.multiProcessorCount;
  }
  for (int k = 0; k < sz4_ctx::kMaxPieces; k++)
    if (cudaEventCreateWithFlags(&ctx->ev_piece[k], cudaEventDisableTiming) != cudaSuccess) { sz4_destroy(ctx); return SZ4_ERR_CUDA; }
  const char* env = getenv("SZ4_STAGE_BULK");          // debugging aid: 0 = stage with plain loads
  if (env && env[0] == '0') ctx->stage_bulk = 0;
  *out = ctx;
  return SZ4_OK;
}

void sz4_destroy(sz4_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  DevBuf* all[] = { &ctx->rank, &ctx->long_list, &ctx->tile_queue, &ctx->jump, &ctx->lsd_state, &ctx->lsd_misc, &ctx->dbg_pe, &ctx->greedy_segs, &ctx->dp_order, &ctx->tile_order, &ctx->data2, &ctx->seg2, &ctx->data, &ctx->ph, &ctx->pe, &ctx->mlen, &ctx->mdist, &ctx->scratch, &ctx->hist, &ctx->hist_scanned,
                    &ctx->partials, &ctx->seqs, &ctx->seq_count, &ctx->packed, &ctx->saved_ph, &ctx->saved_pe, &ctx->seeds,
                    &ctx->nseeds, &ctx->seg, &ctx->block_out, &ctx->seg_total, &ctx->dbg_len, &ctx->dbg_dist, &ctx->scalar_state,
                    &ctx->run_fwd, &ctx->ones_back, &ctx->flag_last, &ctx->flag_carry, &ctx->mfin, &ctx->dp_tasks,
                    &ctx->dp_count, &ctx->dp_states, &ctx->dp_overlays, &ctx->dp_redo, &ctx->dp_reach, &ctx->seqs_tmp,
                    &ctx->path_segs };
  for (DevBuf* b : all) if (b->p) cudaFree(b->p);
  if (ctx->h_seg_total) cudaFreeHost(ctx->h_seg_total);
  for (int k = 0; k < 2; k++) { if (ctx->h_in[k]) cudaFreeHost(ctx->h_in[k]); if (ctx->h_out[k]) cudaFreeHost(ctx->h_out[k]); }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int k = 0; k < 8; k++) if (ctx->pev[k]) cudaEventDestroy(ctx->pev[k]);
  if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
  for (int k = 0; k < sz4_ctx::kMaxPieces; k++) if (ctx->ev_piece[k]) cudaEventDestroy(ctx->ev_piece[k]);
  if (ctx->ev_user) cudaEventDestroy(ctx->ev_user);
  for (int k = 0; k < 2; k++) if (ctx->ev_out[k]) cudaEventDestroy(ctx->ev_out[k]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* sz4_last_error(const sz4_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int sz4_set_option(sz4_ctx* ctx, const char* name, long long value)
{
  if (!ctx || !name) return SZ4_ERR_ARG;
  if (!strcmp(name, "batch_blocks")) { if (value < 1) return SZ4_ERR_ARG; ctx->batch_blocks = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "stream_blocks")) { if (value < 1 || value > 256) return SZ4_ERR_ARG; ctx->stream_blocks = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "block_size")) { ctx->block_size_override = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "stage_bulk")) { ctx->stage_bulk = value != 0; return SZ4_OK; }
  if (!strcmp(name, "debug_keep")) { ctx->debug_keep = value != 0; return SZ4_OK; }
  if (!strcmp(name, "fast_hops")) { if (value < 1 || value > 1024) return SZ4_ERR_ARG; ctx->fast_hops = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "dense_a")) { if (value < 0 || value > 65536) return SZ4_ERR_ARG; ctx->dense_a = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "dense_b")) { if (value < 0 || value > 65536) return SZ4_ERR_ARG; ctx->dense_b = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "fast_lanes")) { if (value < 0 || value > 32) return SZ4_ERR_ARG; ctx->fast_lanes = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "profile")) { ctx->profile = value != 0; return SZ4_OK; }
  if (!strcmp(name, "force_scalar")) { ctx->force_scalar = value != 0; return SZ4_OK; }
  if (!strcmp(name, "allow_scalar_dict")) { ctx->allow_scalar_dict = value != 0; return SZ4_OK; }
  if (!strcmp(name, "tail_lanes")) { if (value < 0 || value > 33) return SZ4_ERR_ARG; ctx->tail_lanes = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "long_age")) { if (value < 0 || value > 1000000) return SZ4_ERR_ARG; ctx->long_age = (uint32_t)value; return SZ4_OK; }
  if (!strcmp(name, "lsd_persist")) { ctx->lsd_persist = value != 0; return SZ4_OK; }
  if (!strcmp(name, "debug_stop")) { ctx->debug_stop = value != 0; return SZ4_OK; }
  ctx->err = "unknown option";
  return SZ4_ERR_ARG;
}

size_t sz4_compress_bound(size_t n, int use_legacy_format)
{
  // modern: a block is stored when packing does not shrink it; legacy blocks are always packed
  size_t bs = use_legacy_format ? kBlockLegacy : kBlockModern;
  size_t blocks = n / bs + 1;
  size_t worst = use_legacy_format ? n + n / 255 + 64 * blocks : n;
  return worst + 4 * blocks + 16;
}

size_t sz4_frame_header(unsigned char* dst, int use_legacy_format)
{
  if (use_legacy_format) { const unsigned char h[4] = { 0x02, 0x21, 0x4C, 0x18 }; memcpy(dst, h, 4); return 4; }   // smallz4.h:482
  const unsigned char h[7] = { 0x04, 0x22, 0x4D, 0x18, 1 << 6, 7 << 4, 0xDF };                                     // smallz4.h:488-494
  memcpy(dst, h, 7);
  return 7;
}

size_t sz4_frame_end(unsigned char* dst, int use_legacy_format)
{
  if (use_legacy_format) return 0;
  memset(dst, 0, 4);                                                                                               // smallz4.h:809-813
  return 4;
}

int sz4_compress_host(sz4_ctx* ctx, const void* src, size_t n, void* dst, size_t cap, size_t* frame_len,
                      unsigned short max_chain, const unsigned char* dict, size_t dict_len, int legacy)
{
  if (!ctx || (!src && n) || !dst || !frame_len) return SZ4_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->kernel_ms = 0; ctx->launches = 0; ctx->dp_redos = 0; ctx->path_redos = 0;
  for (int k = 0; k < 7; k++) ctx->phase_ms[k] = 0;
  uint8_t* out = (uint8_t*)dst;
  if (cap < 16) { ctx->err = "destination too small"; return SZ4_ERR_DST_SMALL; }
  size_t o = sz4_frame_header(out, legacy);

  if (max_chain == 0)
  {
    // level -0 (smallz4.h:511,765-780): blocks are stored; nothing to compute.  Legacy frames
    // cannot mark stored blocks, the reference then writes empty blocks -- so do we.
    const size_t bs = ctx->block_size_override ? ctx->block_size_override : (legacy ? kBlockLegacy : kBlockModern);
    for (size_t at = 0; at < n; at += bs)
    {
      size_t len = n - at < bs ? n - at : bs;
      uint32_t tagged = legacy ? 0u : ((uint32_t)len | 0x80000000u);
      size_t body = legacy ? 0 : len;
      if (o + 4 + body + 4 > cap) { ctx->err = "destination too small"; return SZ4_ERR_DST_SMALL; }
      out[o++] = (uint8_t)tagged; out[o++] = (uint8_t)(tagged >> 8); out[o++] = (uint8_t)(tagged >> 16); out[o++] = (uint8_t)(tagged >> 24);
      memcpy(out + o, (const uint8_t*)src + at, body);
      o += body;
    }
  }
  else
  {
    StreamJob job;
    job.src = (const uint8_t*)src; job.n = n; job.dict = dict; job.dict_len = dict_len;
    job.dst = out + o; job.cap = cap - o - 4; job.max_chain = max_chain; job.legacy = legacy != 0;
    size_t seg = 0;
    int r = compress_blocks(ctx, job, &seg);
    if (r != SZ4_OK) return r;
    o += seg;
  }
  if (o + 4 > cap) { ctx->err = "destination too small"; return SZ4_ERR_DST_SMALL; }
  o += sz4_frame_end(out + o, legacy);
  *frame_len = o;
  return SZ4_OK;
}

int sz4_compress_device(sz4_ctx* ctx, const void* d_src, size_t halo, size_t n, int is_first, int is_last,
                        void* d_dst, size_t cap, size_t* segment_len, unsigned short max_chain, int legacy, void* cuda_stream)
{
  if (!ctx || !d_src || !d_dst || !segment_len || max_chain == 0) return SZ4_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  // the caller's stream may still be producing d_src (or reading d_dst): the context's streams start behind it.
  // The call returns after its own work has finished, so nothing has to be handed back to that stream.
  CK(cudaEventRecord(ctx->ev_user, (cudaStream_t)cuda_stream));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_user, 0));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_user, 0));
  ctx->kernel_ms = 0; ctx->launches = 0; ctx->dp_redos = 0; ctx->path_redos = 0;
  for (int k = 0; k < 7; k++) ctx->phase_ms[k] = 0;
  StreamJob job;
  job.src = (const uint8_t*)d_src; job.src_on_device = true; job.halo_in_src = halo; job.n = n;
  job.first = is_first != 0; job.last = is_last != 0;
  job.dst = (uint8_t*)d_dst; job.dst_on_device = true; job.cap = cap; job.max_chain = max_chain; job.legacy = legacy != 0;
  if (job.first && halo != 0) { ctx->err = "the first block of a stream has no halo"; return SZ4_ERR_ARG; }
  return compress_blocks(ctx, job, segment_len);
}

int sz4_compress_host_range(sz4_ctx* ctx, const void* src, size_t halo, size_t n, int is_first, int is_last,
                            void* dst, size_t cap, size_t* segment_len, unsigned short max_chain, int legacy)
{
  if (!ctx || !src || !dst || !segment_len || max_chain == 0) return SZ4_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->kernel_ms = 0; ctx->launches = 0; ctx->dp_redos = 0; ctx->path_redos = 0;
  for (int k = 0; k < 7; k++) ctx->phase_ms[k] = 0;
  StreamJob job;
  job.src = (const uint8_t*)src; job.halo_in_src = halo; job.n = n;
  job.first = is_first != 0; job.last = is_last != 0;
  job.dst = (uint8_t*)dst; job.cap = cap; job.max_chain = max_chain; job.legacy = legacy != 0;
  if (job.first && halo != 0) { ctx->err = "the first block of a stream has no halo"; return SZ4_ERR_ARG; }
  return compress_blocks(ctx, job, segment_len);
}

int sz4_lz4(sz4_ctx* ctx, sz4_get_bytes get_bytes, sz4_send_bytes send_bytes, unsigned short max_chain,
            const unsigned char* dict, size_t dict_len, int legacy, void* user)
{
  if (!ctx || !get_bytes || !send_bytes) return SZ4_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (dict == nullptr || dict_len == 0)
  {
    ctx->kernel_ms = 0; ctx->launches = 0; ctx->dp_redos = 0; ctx->path_redos = 0;
    for (int k = 0; k < 7; k++) ctx->phase_ms[k] = 0;
    unsigned char mark[8];
    send_bytes(mark, sz4_frame_header(mark, legacy), user);                                  // smallz4.h:479-496
    if (max_chain == 0)
    {
      // level -0 (smallz4.h:511,765-780): blocks are stored, nothing to compute; legacy frames cannot mark stored
      // blocks and the reference then writes empty ones -- so do we
      const size_t bs = ctx->block_size_override ? ctx->block_size_override : (legacy ? kBlockLegacy : kBlockModern);
      std::vector<uint8_t> block(bs);
      for (;;)
      {
        size_t have = 0;
        while (have < bs)
        {
          const size_t got = get_bytes(block.data() + have, bs - have < 65536 ? bs - have : 65536, user);
          if (got == 0) break;
          have += got;
        }
        if (have == 0) break;
        const uint32_t tagged = legacy ? 0u : ((uint32_t)have | 0x80000000u);
        const unsigned char word[4] = { (unsigned char)tagged, (unsigned char)(tagged >> 8), (unsigned char)(tagged >> 16), (unsigned char)(tagged >> 24) };
        send_bytes(word, 4, user);
        if (!legacy) send_bytes(block.data(), have, user);
        if (have < bs) break;
      }
    }
    else
    {
      int r = compress_stream(ctx, get_bytes, send_bytes, user, max_chain, legacy != 0);
      if (r != SZ4_OK) return r;
    }
    const size_t e = sz4_frame_end(mark, legacy);                                           // smallz4.h:809-813
    if (e) send_bytes(mark, e, user);
    return SZ4_OK;
  }
  // With a dictionary the whole stream is taken in first: whether the parallel match finder applies depends on
  // every byte of it (DESIGN.md Q-dict), and the reference's own -D frames do not decode anyway.
  std::vector<uint8_t> in;
  const size_t kChunk = 64 * 1024;
  for (;;)
  {
    size_t at = in.size();
    in.resize(at + kChunk);
    size_t got = get_bytes(in.data() + at, kChunk, user);
    in.resize(at + got);
    if (got == 0) break;
  }
  size_t cap = sz4_compress_bound(in.size(), legacy) + in.size();
  std::vector<uint8_t> out(cap);
  size_t len = 0;
  int r = sz4_compress_host(ctx, in.data(), in.size(), out.data(), cap, &len, max_chain, dict, dict_len, legacy);
  if (r != SZ4_OK) return r;
  send_bytes(out.data(), len, user);
  return SZ4_OK;
}

int sz4_last_stats(const sz4_ctx* ctx, double* kernel_ms, unsigned long long* launches)
{
  if (!ctx) return SZ4_ERR_ARG;
  if (kernel_ms) *kernel_ms = ctx->kernel_ms;
  if (launches) *launches = ctx->launches;
  return SZ4_OK;
}

long long sz4_last_dp_redos(const sz4_ctx* ctx) { return ctx ? (long long)ctx->dp_redos : -1; }
/* debug: raw counters of the last batch (k-cycles): [2] longest spec task, total spec, its length; [4] verify per block max, redo max */
#ifdef SZ4_TILE_STATS
// debugging aid, only in builds with -DSZ4_TILE_STATS (tools/tile_stats.py)
int sz4_debug_tile_stats(unsigned* us, unsigned* t0, unsigned n)
{
  if (cudaMemcpyFromSymbol(us, sz4::g_tile_us, n * 4) != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(t0, sz4::g_tile_t0, n * 4) != cudaSuccess ? -1 : 0;
}
#endif

#ifdef SZ4_SEARCH_STATS
int sz4_debug_search_stats(unsigned long long* out, int reset)
{
  if (cudaMemcpyFromSymbol(out, sz4::g_stats, 64) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[8] = { 0 }; cudaMemcpyToSymbol(sz4::g_stats, z, 64); }
  return 0;
}
#endif

const unsigned* sz4_debug_counters(const sz4_ctx* ctx) { return ctx ? (const unsigned*)(ctx->h_seg_total + 1) : nullptr; }
long long sz4_last_path_redos(const sz4_ctx* ctx) { return ctx ? (long long)ctx->path_redos : -1; }

int sz4_last_phase_ms(const sz4_ctx* ctx, double* out7)
{
  if (!ctx || !out7) return SZ4_ERR_ARG;
  for (int k = 0; k < 7; k++) out7[k] = ctx->phase_ms[k];
  return SZ4_OK;
}

int sz4_debug_fetch(sz4_ctx* ctx, const char* what, void* dst, size_t count)
{
  if (!ctx || !what || !dst) return SZ4_ERR_ARG;
  const Geom& g = ctx->last_geom;
  if (count > g.n_total) count = g.n_total;
  const void* src = nullptr; size_t elem = 0;
  if (!strcmp(what, "pe")) { src = ctx->dbg_pe.p; elem = 2; }
  else if (!strcmp(what, "ph")) { src = (uint16_t*)ctx->ph.p + kPad; elem = 2; }
  else if (!strcmp(what, "pe8")) { src = (uint16_t*)ctx->pe.p + kPad; elem = 2; }
  else if (!strcmp(what, "jump")) { src = ctx->jump.p; elem = 8; }
  else if (!strcmp(what, "len_found")) { src = ctx->dbg_len.p; elem = 4; }
  else if (!strcmp(what, "dist_found")) { src = ctx->dbg_dist.p; elem = 2; }
  else if (!strcmp(what, "len_final")) { src = ctx->dp_ran ? ctx->mfin.p : ctx->mlen.p; elem = 4; }
  else if (!strcmp(what, "dist_final")) { src = ctx->mdist.p; elem = 2; }
  else if (!strcmp(what, "cost")) { src = ctx->scratch.p; elem = 4; }
  if (!src) { ctx->err = "unknown array"; return SZ4_ERR_ARG; }
  CK(cudaMemcpy(dst, src, count * elem, cudaMemcpyDeviceToHost));
  return SZ4_OK;
}

}  // extern "C"
