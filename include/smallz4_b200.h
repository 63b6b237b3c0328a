/*
 * smallz4_b200.h -- C ABI of libsmallz4_b200.so, the B200-native replacement for the hot path of
 * smalLZ4 (optimal-parse LZ4 compression).  Plain pointers and sizes only; bind it from C, C++
 * (include/smallz4.h is the drop-in class), cgo, JNI or ctypes.
 *
 * What each entry point replaces in the reference (/root/reference):
 *   sz4_lz4()                smallz4::lz4(getBytes, sendBytes, maxChainLength, dictionary,
 *                            useLegacyFormat, userPtr)                       smallz4.h:47-64
 *                            -> smallz4::compress()                          smallz4.h:476-814
 *   sz4_compress_host()      the same call with in-memory callbacks (what smallz4.cpp:66-117
 *                            getBytesFromIn / sendBytesToOut do with FILE*)
 *   sz4_compress_device()    the per-block loop of compress() (smallz4.h:547-806) for input that
 *                            is already resident in HBM; used for multi-GPU shards
 *   sz4_version()            smallz4::getVersion()                           smallz4.h:67
 *   SZ4_LEVEL_*              ShortChainsGreedy / ShortChainsLazy             smallz4.h:74-80
 *
 * Output is byte-identical to the reference's for the same input, maxChainLength, dictionary and
 * format.  All compute runs in hand-written sm_100a CUDA kernels; there is no CPU fallback:
 * every call fails with SZ4_ERR_CUDA when no CUDA device is usable.
 */
#ifndef SMALLZ4_B200_H
#define SMALLZ4_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SZ4_OK             0
#define SZ4_ERR_CUDA      -1   /* no device / CUDA runtime error (sz4_last_error has the text) */
#define SZ4_ERR_ARG       -2   /* invalid argument                                             */
#define SZ4_ERR_NOMEM     -3   /* host or device allocation failed                             */
#define SZ4_ERR_DST_SMALL -4   /* destination buffer too small                                 */

#define SZ4_LEVEL_GREEDY_MAX 3      /* smallz4::ShortChainsGreedy */
#define SZ4_LEVEL_LAZY_MAX   6      /* smallz4::ShortChainsLazy   */
#define SZ4_MAX_CHAIN_DEFAULT 65535 /* smallz4::MaxChainLength: level -9 */

/* same signatures as smallz4::GET_BYTES / smallz4::SEND_BYTES (smallz4.h:42-44) */
typedef size_t (*sz4_get_bytes)(void* data, size_t numBytes, void* userPtr);
typedef void   (*sz4_send_bytes)(const void* data, size_t numBytes, void* userPtr);

typedef struct sz4_ctx sz4_ctx;

/* device < 0: current device.  The context owns its stream and device buffers; one per thread. */
int         sz4_create(sz4_ctx** out, int device);
void        sz4_destroy(sz4_ctx* ctx);
const char* sz4_last_error(const sz4_ctx* ctx);
const char* sz4_version(void);

/* tuning / test knobs: "batch_blocks" (blocks per device batch), "block_size" (tests only: a multiple
   of 65536, >= 131072; 0 = format default), "stage_bulk" (1 = cp.async.bulk staging, 0 = plain loads),
   "debug_keep" (keep intermediates of the last batch for sz4_debug_fetch), "profile" (per-phase CUDA-event
   timing for sz4_last_phase_ms), "force_scalar" (tests: route a dictionary stream through the scalar finder), "allow_scalar_dict" (a -D stream
   that contains 60 000 or more equal bytes in a row is refused with SZ4_ERR_ARG unless this is 1: then one device
   thread replays the reference's ring, exact but slow), "stream_blocks" (sz4_lz4: blocks per batch, bounds its pinned
   host memory), "long_age" (match finder: rounds after which a walk moves to the warp-per-walk kernel), "tail_lanes" (0 = off:
   hand a warp's last walks over when the tile's queue is empty), "lsd_persist" (sort passes by persistent CTAs);
   match-finder scheduling, results never depend on them: "fast_hops" (candidates per lane and round, 1..1024),
   "fast_lanes" (lanes that must still be walking for a round to go on, 0..32), "dense_a" / "dense_b" (positions
   whose first two chain hops add up to less than this go first / second; dense_a = 0: one pass) */
int sz4_set_option(sz4_ctx* ctx, const char* name, long long value);

/* worst-case size of the frame produced for n input bytes */
size_t sz4_compress_bound(size_t n, int use_legacy_format);

/* drop-in for smallz4::lz4(): pulls the whole stream through get_bytes, pushes the frame through
   send_bytes.  max_chain_length 0..65535 (levels -0..-8 map to 0..8, -9 to 65535, smallz4.cpp:175,233) */
int sz4_lz4(sz4_ctx* ctx, sz4_get_bytes get_bytes, sz4_send_bytes send_bytes,
            unsigned short max_chain_length, const unsigned char* dictionary, size_t dictionary_len,
            int use_legacy_format, void* user_ptr);

/* host buffer -> host buffer (pinned buffers make the copies asynchronous) */
int sz4_compress_host(sz4_ctx* ctx, const void* src, size_t n, void* dst, size_t dst_capacity, size_t* frame_len,
                      unsigned short max_chain_length, const unsigned char* dictionary, size_t dictionary_len,
                      int use_legacy_format);

/* Device-resident whole blocks.  d_src points at `halo` bytes of history followed by `n` bytes that
   start on a block border of the stream; is_stream_first / is_stream_last say whether the range
   contains the first / last block of the stream.  Writes the concatenated [size][payload] block
   records (no frame header, no end mark) to d_dst (device) and their total length to *segment_len.
   This is the unit one GPU of a sharded job produces; blocks depend only on their 64 KiB halo.
   cuda_stream (a cudaStream_t, NULL = the default stream): work already enqueued there (e.g. the producer of
   d_src) is waited for on the device; the call itself returns when the records are complete. */
int sz4_compress_device(sz4_ctx* ctx, const void* d_src, size_t halo, size_t n, int is_stream_first, int is_stream_last,
                        void* d_dst, size_t dst_capacity, size_t* segment_len,
                        unsigned short max_chain_length, int use_legacy_format, void* cuda_stream);

/* The same unit from and to HOST memory (pinned buffers make the copies asynchronous): src points at `halo` bytes of
   history followed by `n` bytes of whole blocks; the block records go to dst (host).  What one rank of a sharded job
   calls when its range of the input lives in host memory; the copies overlap the kernels batch by batch. */
int sz4_compress_host_range(sz4_ctx* ctx, const void* src, size_t halo, size_t n, int is_stream_first, int is_stream_last,
                            void* dst, size_t dst_capacity, size_t* segment_len,
                            unsigned short max_chain_length, int use_legacy_format);

/* frame header / end mark (smallz4.h:479-496, 809-813) for callers that assemble shards themselves */
size_t sz4_frame_header(unsigned char* dst, int use_legacy_format);
size_t sz4_frame_end(unsigned char* dst, int use_legacy_format);

/* milliseconds the device spent in the kernels of the last call and number of kernel launches */
int sz4_last_stats(const sz4_ctx* ctx, double* kernel_ms, unsigned long long* launches);

/* with option "profile"=1: device milliseconds of the last call per phase, out7 =
   { sort, chain (link + exact walk), search, fix-up / greedy filter, cost DP, parse walk, emission } */
int sz4_last_phase_ms(const sz4_ctx* ctx, double* out7);

/* segments of the cost DP that failed verification in the last call and were priced again */
long long sz4_last_dp_redos(const sz4_ctx* ctx);
long long sz4_last_path_redos(const sz4_ctx* ctx);   /* same for the segments of the parse walk */

/* test hook: copy an intermediate array of the last batch to the host (needs option debug_keep=1).
   what: "pe" u16 (previousExact as the reference's ring holds it), "pe8" u16, "jump" u64 (pe4..pe7 by anchor),
   "len_found" u32, "dist_found" u16, "len_final" u32, "cost" u32; count = elements */
int sz4_debug_fetch(sz4_ctx* ctx, const char* what, void* dst, size_t count);

#ifdef __cplusplus
}
#endif
#endif
