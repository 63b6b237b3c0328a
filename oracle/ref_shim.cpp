// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// In-memory entry point around the UNMODIFIED reference header.  It is compiled with
// -I/root/reference by oracle/Makefile into oracle/_ref/libsmallz4ref.so; no reference
// source is copied into this repository.  Used (a) to pin the oracle and (b) as the
// "reference" CPU arm of bench.py.
#include "smallz4.h"   // /root/reference/smallz4.h

#include <cstring>
#include <vector>

namespace
{
struct Io
{
  const unsigned char* src; size_t n, at;
  unsigned char* dst; size_t cap, out; bool overflow;
};
size_t pull(void* data, size_t want, void* user)
{
  Io* io = static_cast<Io*>(user);
  size_t left = io->n - io->at;
  size_t got = want < left ? want : left;
  if (got) std::memcpy(data, io->src + io->at, got);
  io->at += got;
  return got;
}
void push(const void* data, size_t len, void* user)
{
  Io* io = static_cast<Io*>(user);
  if (io->out + len > io->cap) { io->overflow = true; return; }
  std::memcpy(io->dst + io->out, data, len);
  io->out += len;
}
}

extern "C" long long ref_smallz4_compress(const unsigned char* src, size_t n,
                                          const unsigned char* dict, size_t dict_len,
                                          unsigned max_chain, int legacy,
                                          unsigned char* dst, size_t cap)
{
  Io io = { src, n, 0, dst, cap, 0, false };
  std::vector<unsigned char> d;
  if (dict && dict_len) d.assign(dict, dict + dict_len);
  smallz4::lz4(pull, push, (unsigned short)max_chain, d, legacy != 0, &io);
  return io.overflow ? -1 : (long long)io.out;
}

extern "C" const char* ref_smallz4_version() { return smallz4::getVersion(); }
