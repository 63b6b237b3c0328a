// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler behind cuda_emu.h.
#include "cuda_emu.h"

#include <sys/mman.h>

namespace emu
{
Fiber* cur = nullptr;
Cta* cta = nullptr;
void* sched_sp = nullptr;
uint3_emu g_blockIdx{ 0, 0, 0 };
dim3 g_blockDim, g_gridDim;
unsigned char* dyn_smem = nullptr;
const std::function<void()>* body = nullptr;

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

[[noreturn]] void die(const char* msg)
{
  fprintf(stderr, "SIMT emulator: %s\n", msg);
  abort();
}

void yield() { emu_switch(&cur->sp, sched_sp); }

static void fiber_main()
{
  (*body)();
  Fiber* f = cur;
  f->done = true;
  cta->alive--;
  cta->warps[f->warp].alive &= ~(1u << f->lane);
  cta->progress++;
  emu_switch(&f->sp, sched_sp);
  die("resumed a finished fiber");
}

static const size_t kStack = 256 * 1024;
static std::vector<void*> stack_pool;

static void* get_stack()
{
  if (!stack_pool.empty()) { void* s = stack_pool.back(); stack_pool.pop_back(); return s; }
  void* s = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (s == MAP_FAILED) die("cannot map fiber stack");
  return s;
}

static void prepare(Fiber& f)
{
  f.stack = get_stack();
  uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
  uint64_t* sp = (uint64_t*)top;
  *--sp = 0;                          // fake return address slot keeps (rsp+8) % 16 == 0 at entry
  *--sp = (uint64_t)(uintptr_t)&fiber_main;
  for (int i = 0; i < 6; i++) *--sp = 0;   // r15 r14 r13 r12 rbx rbp
  f.sp = sp;
}

// Gather one 64-bit value from every lane of the calling warp (all 32 lanes must call).
// Collective k of a warp uses slot k&1.  The lane that completes collective k clears the other
// slot: every lane has left collective k-1 by then, and none can deposit for k+1 before k completes.
uint64_t collective(uint64_t v, uint64_t* out32)
{
  Fiber* f = cur;
  Warp& w = cta->warps[f->warp];
  if (w.alive != 0xffffffffu) die("warp collective after some lanes of the warp exited");
  const uint64_t seq = f->coll_seq++;
  WarpSlot& s = w.slot[seq & 1];
  WarpSlot& other = w.slot[(seq + 1) & 1];
  const uint32_t bit = 1u << f->lane;
  if (s.arrived & bit) die("warp collective re-entered (lanes out of step)");
  s.val[f->lane] = v;
  s.arrived |= bit;
  if (s.arrived == 0xffffffffu) { other.arrived = 0; cta->progress++; }
  while (s.arrived != 0xffffffffu)
  {
    if (w.alive != 0xffffffffu) die("a lane exited while its warp waits in a collective");
    yield();
  }
  memcpy(out32, s.val, sizeof(s.val));
  return 0;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn)
{
  if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) die("emulator supports 1-D launches only");
  if (block.x % 32 != 0) die("blockDim.x must be a multiple of 32");
  g_blockDim = block; g_gridDim = grid;
  std::vector<unsigned char> dyn(smem + 64);
  dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  body = &fn;
  Cta c;
  cta = &c;
  for (unsigned b = 0; b < grid.x; b++)
  {
    g_blockIdx = uint3_emu{ b, 0, 0 };
    c.fibers.assign(block.x, Fiber());
    c.warps.assign(block.x / 32, Warp());
    for (auto& w : c.warps) { w.slot[0].arrived = w.slot[1].arrived = 0; w.alive = 0xffffffffu; }
    c.alive = block.x; c.bar_arrived = 0; c.bar_gen = 0; c.progress = 0; c.or_val[0] = c.or_val[1] = 0;
    for (unsigned t = 0; t < block.x; t++)
    {
      Fiber& f = c.fibers[t];
      f.tid = uint3_emu{ t, 0, 0 }; f.lane = t & 31; f.warp = t >> 5; f.coll_seq = 0; f.or_seq = 0; f.done = false;
      prepare(f);
    }
    while (c.alive)
    {
      uint64_t before = c.progress;
      for (unsigned t = 0; t < block.x; t++)
      {
        Fiber& f = c.fibers[t];
        if (f.done) continue;
        cur = &f;
        emu_switch(&sched_sp, f.sp);
      }
      if (c.alive && c.progress == before) die("deadlock: no thread of the CTA can make progress");
    }
    for (auto& f : c.fibers) stack_pool.push_back(f.stack);
  }
  cur = nullptr; cta = nullptr; body = nullptr; dyn_smem = nullptr;
}
}  // namespace emu
