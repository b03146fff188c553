// TEST INFRASTRUCTURE: storage for the CUDA-emulation globals (see cuda_emu.h).
#define DKTB_EMU_IMPL
#include "cuda_emu.h"
