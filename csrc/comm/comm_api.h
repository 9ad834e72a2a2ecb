// Host launchers of the peer-memory collective kernels (csrc/comm/*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
