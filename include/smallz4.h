// smallz4.h -- C++ drop-in for the reference's public class (reference: smallz4.h:38-80).
//
// Same class name, same static entry points, same callback types, same defaults, so code written
// against smalLZ4 compiles unchanged:
//
//     #include "smallz4.h"
//     smallz4::lz4(getBytes, sendBytes);                                       // level 9
//     smallz4::lz4(getBytes, sendBytes, maxChainLength, dictionary, useLegacyFormat, userPtr);
//
// The work is done by libsmallz4_b200.so (include/smallz4_b200.h) on a B200; there is no CPU path.
// A failure (no CUDA device, out of memory) throws std::runtime_error -- the reference has no error
// channel in this API, and writing nothing silently would be worse.
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "smallz4_b200.h"

class smallz4
{
public:
  // smallz4.h:42-44
  typedef size_t (*GET_BYTES)(void* data, size_t numBytes, void* userPtr);
  typedef void (*SEND_BYTES)(const void* data, size_t numBytes, void* userPtr);

  // smallz4.h:74-80
  enum
  {
    ShortChainsGreedy = SZ4_LEVEL_GREEDY_MAX,
    ShortChainsLazy = SZ4_LEVEL_LAZY_MAX
  };

  // smallz4.h:47-53
  static void lz4(GET_BYTES getBytes, SEND_BYTES sendBytes, unsigned short maxChainLength = SZ4_MAX_CHAIN_DEFAULT,
                  bool useLegacyFormat = false, void* userPtr = NULL)
  {
    lz4(getBytes, sendBytes, maxChainLength, std::vector<unsigned char>(), useLegacyFormat, userPtr);
  }

  // smallz4.h:56-64
  static void lz4(GET_BYTES getBytes, SEND_BYTES sendBytes, unsigned short maxChainLength,
                  const std::vector<unsigned char>& dictionary, bool useLegacyFormat = false, void* userPtr = NULL)
  {
    sz4_ctx* ctx = NULL;
    if (sz4_create(&ctx, -1) != SZ4_OK)
      throw std::runtime_error("smallz4_b200: no usable CUDA device");
    int rc = sz4_lz4(ctx, getBytes, sendBytes, maxChainLength, dictionary.empty() ? NULL : &dictionary[0],
                     dictionary.size(), useLegacyFormat ? 1 : 0, userPtr);
    std::string msg = rc == SZ4_OK ? std::string() : std::string(sz4_last_error(ctx));
    sz4_destroy(ctx);
    if (rc != SZ4_OK)
      throw std::runtime_error("smallz4_b200: " + msg);
  }

  // smallz4.h:67
  static const char* getVersion() { return sz4_version(); }
};
