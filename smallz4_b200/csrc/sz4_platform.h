// sz4_platform.h -- CUDA runtime (product build, nvcc, sm_100a) or the SIMT emulator (tests only).
#pragma once

#ifdef SZ4_EMU
  #include "cuda_emu.h"          // tests/emu/cuda_emu.h; never defined in the product build
  #define SZ4_DYN_SMEM(name) unsigned char* name = emu::dyn_smem
#else
  #include <cuda_runtime.h>
  #define SZ4_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
  #define SZ4_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

#include <stdint.h>
#include <stddef.h>
