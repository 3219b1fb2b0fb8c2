// TEST-ONLY translation unit: the product sources (C ABI + host glue + kernels), compiled with g++
// against the fiber emulator in cuda_emu.h.  Built into tests/emu/_build/ (git- and gpurun-ignored)
// by tests/emu/build_emu.py; never part of the package.
#define SFB_EMU 1
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../simfire_b200/csrc/sfb.cu"

extern "C" int sfb_emu_marker(void) { return 1; }
extern "C" long long sfb_emu_launches(void) { return emu::st().launches; }
