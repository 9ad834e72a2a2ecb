// Host launchers of the tcgen05 fused attention kernels (csrc/attn/fmha_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
