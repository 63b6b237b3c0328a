// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A small SIMT emulator: compiles the repository's .cu/.cuh sources with g++ (-DSZ4_EMU) and
// runs every CUDA thread as a cooperative fiber, one CTA at a time, so the kernel LOGIC
// (indexing, barriers, warp collectives, atomics) can be checked against the oracle in the
// CPU test-suite.  It is never part of libsmallz4_b200.so and nothing in the product path
// can reach it: the product library is built by nvcc without SZ4_EMU and fails loudly when
// no CUDA device is present.
//
// Supported: __global__/__device__ functions, threadIdx/blockIdx/blockDim/gridDim (.x only used),
// static and dynamic __shared__, __syncthreads, full-mask warp collectives (__syncwarp, __ballot_sync,
// __shfl_*_sync, __match_any_sync, __reduce_{min,max,add}_sync, __any/__all_sync), atomics,
// bit intrinsics, and the handful of runtime calls the host pipeline makes.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint3_emu { unsigned x, y, z; };
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int4 { int x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{ a, b, c, d }; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{ a, b }; }

namespace emu
{
struct WarpSlot { uint64_t val[32]; uint32_t arrived; };
struct Warp { WarpSlot slot[2]; uint32_t alive; };
struct Fiber
{
  void* sp = nullptr;
  void* stack = nullptr;
  uint3_emu tid{ 0, 0, 0 };
  int lane = 0, warp = 0;
  uint64_t coll_seq = 0;
  unsigned or_seq = 0;
  bool done = false;
};
struct Cta
{
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  unsigned alive = 0, bar_arrived = 0;
  uint64_t bar_gen = 0;
  uint64_t progress = 0;
  int or_val[2] = { 0, 0 };
};

extern Fiber* cur;
extern Cta* cta;
extern void* sched_sp;
extern uint3_emu g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* dyn_smem;
extern const std::function<void()>* body;

void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn);
[[noreturn]] void die(const char* msg);

uint64_t collective(uint64_t v, uint64_t* out32 /* array of 32 gathered values */);
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

// ---------------------------------------------------------------- barriers and warp collectives
static inline void __syncthreads()
{
  emu::Cta* c = emu::cta;
  uint64_t gen = c->bar_gen;
  c->bar_arrived++;
  for (;;)
  {
    if (c->bar_gen != gen) return;
    if (c->bar_arrived >= c->alive) { c->bar_arrived = 0; c->bar_gen++; c->progress++; return; }
    emu::yield();
  }
}
// barrier + OR of the predicate over the CTA (two slots in turn: a slot is cleared behind its second barrier, while
// the next call already uses the other one)
static inline int __syncthreads_or(int pred)
{
  emu::Cta* c = emu::cta;
  const unsigned slot = emu::cur->or_seq++ & 1;
  if (pred) c->or_val[slot] = 1;
  __syncthreads();
  const int r = c->or_val[slot];
  __syncthreads();
  c->or_val[slot] = 0;
  return r;
}
static inline void emu_check_mask(unsigned mask)
{
  if (mask != 0xffffffffu) emu::die("emulator supports full-mask warp collectives only");
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu_check_mask(mask); uint64_t g[32]; emu::collective(0, g); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
  emu_check_mask(mask); uint64_t g[32]; emu::collective(pred ? 1 : 0, g);
  unsigned r = 0; for (int i = 0; i < 32; i++) if (g[i]) r |= 1u << i; return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
  emu_check_mask(mask); uint64_t g[32]; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); emu::collective(raw, g);
  int lane = emu::cur->lane; int base = lane & ~(width - 1);
  uint64_t r = g[base + (src & (width - 1))]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
  emu_check_mask(mask); uint64_t g[32]; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); emu::collective(raw, g);
  int lane = emu::cur->lane; int base = lane & ~(width - 1); int src = lane - (int)delta;
  uint64_t r = (src < base) ? g[lane] : g[src]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
  emu_check_mask(mask); uint64_t g[32]; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); emu::collective(raw, g);
  int lane = emu::cur->lane; int base = lane & ~(width - 1); int src = lane + (int)delta;
  uint64_t r = (src >= base + width) ? g[lane] : g[src]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
  emu_check_mask(mask); uint64_t g[32]; uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); emu::collective(raw, g);
  int lane = emu::cur->lane; uint64_t r = g[(lane ^ lanemask) & 31]; T out; memcpy(&out, &r, sizeof(T)); return out;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v)
{
  emu_check_mask(mask); uint64_t g[32]; emu::collective(v, g);
  unsigned r = 0; for (int i = 0; i < 32; i++) if ((unsigned)g[i] == v) r |= 1u << i; return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
  emu_check_mask(mask); uint64_t g[32]; emu::collective(v, g);
  unsigned r = 0xffffffffu; for (int i = 0; i < 32; i++) r = std::min(r, (unsigned)g[i]); return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
  emu_check_mask(mask); uint64_t g[32]; emu::collective(v, g);
  unsigned r = 0; for (int i = 0; i < 32; i++) r = std::max(r, (unsigned)g[i]); return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v)
{
  emu_check_mask(mask); uint64_t g[32]; emu::collective(v, g);
  unsigned r = 0; for (int i = 0; i < 32; i++) r += (unsigned)g[i]; return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }

// ---------------------------------------------------------------- atomics and bit intrinsics
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
  uint64_t x = ((uint64_t)hi << 32) | lo; return (unsigned)(x >> (sh & 31));
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
  uint64_t x = ((uint64_t)b << 32) | a; unsigned r = 0;
  for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((x >> (8 * sel)) & 0xff) << (8 * i); }
  return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
using std::max;
using std::min;

// ---------------------------------------------------------------- runtime API subset
typedef int cudaError_t;
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; int major, minor; size_t totalGlobalMem; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
  memset(p, 0, sizeof(*p)); p->multiProcessorCount = 4; strcpy(p->name, "SIMT emulator"); p->major = 10; p->totalGlobalMem = (size_t)8 << 30; return 0;
}
// device memory is NOT zeroed by the real cudaMalloc: poison it so that reads of uninitialised memory show up
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xA5, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event{ 0 }; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

// SZ4_EMU_TRACE=1 in the environment prints every launch (which kernel a crash belongs to)
#define SZ4_LAUNCH(kernel, grid, block, smem, stream, ...)                                                        \
  do {                                                                                                            \
    if (getenv("SZ4_EMU_TRACE")) fprintf(stderr, "emu launch %s grid %u block %u\n", #kernel, (unsigned)dim3(grid).x, (unsigned)dim3(block).x); \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); });                         \
  } while (0)
