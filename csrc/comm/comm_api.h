// Host launchers of the peer-memory collective kernels (csrc/comm/*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

constexpr int kMaxPeers = 8;         // one NVSwitch domain (HGX B200: 8 GPUs)
constexpr int kMaxCommBlocks = 64;   // flag slots: [kMaxCommBlocks][kMaxPeers] uint32 per rank

enum CommAlgo : int { kAlgoAuto = 0, kAlgoOneShot = 1, kAlgoTwoShot = 2, kAlgoNvls = 3 };

// Symmetric-memory view handed to the kernels by value.
struct CommPeers {
  void* buf[kMaxPeers];     // peer-mapped base address of every rank's data buffer
  void* flags[kMaxPeers];   // peer-mapped base address of every rank's flag buffer (zero-initialised)
  void* multicast;          // NVLS multicast alias of the data buffers (nullptr when unsupported)
  int rank, world;
};

int pick_allreduce_algo(long long bytes, int world, bool has_multicast);

// In-place sum over ranks of bytes [byte_offset, byte_offset + bytes) of the symmetric buffer
// (16-byte aligned range), fp32 accumulation in rank order, result scaled by `scale`.
// sq_acc (nullable, local device memory): this rank atomically adds the sum of squares of the REDUCED, scaled
// values of its 1/world slice of the range - summed over ranks that is the squared L2 norm of the result, so the
// gradient-norm pass over the arena (a separate 2-byte-per-parameter read) is not needed.
void launch_allreduce(const CommPeers& peers, long long byte_offset, long long bytes, int dtype, float scale, int algo,
                      int blocks, float* sq_acc, cudaStream_t stream);

}  // namespace ub
