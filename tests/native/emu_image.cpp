// TEST INFRASTRUCTURE — csrc/arah_image.cu (kernels AND C-ABI entry points, unchanged source) on the CPU execution model of cuda_emu.h.
#include "cuda_emu.h"
#include "../../arah_release_b200/csrc/arah_image.cu"
