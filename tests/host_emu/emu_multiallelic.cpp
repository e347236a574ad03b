// tests/host_emu/emu_multiallelic.cpp -- TEST INFRASTRUCTURE.  The --treat_multiallelics kernels and their C ABI
// (csrc/multiallelic.cu) compiled for the host over the emulated runtime: every kernel is a grid-stride loop, so one
// emulated thread runs it whole.
#define UGVC_HOST_EMU 1
#include <cuda_runtime.h>

#include "../../variantcalling_b200/csrc/multiallelic.cu"
