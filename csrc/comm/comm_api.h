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
// scatter_only (experimental, sharded optimizer): stop after the reduce-scatter half - rank r's 1/world slice of the
// range holds the result in ITS buffer only (slice = vectors [begin + r*per, begin + (r+1)*per), per = ceil(n/world)).
void launch_allreduce(const CommPeers& peers, long long byte_offset, long long bytes, int dtype, float scale, int algo,
                      int blocks, float* sq_acc, cudaStream_t stream, bool scatter_only = false);

constexpr int kMaxShardRanges = 48;  // one contiguous shard, or one slice per gradient bucket

// ---- optimizer step fused with the parameter all-gather (experimental: UNICORE_B200_SHARD_OPTIMIZER=1) --------
// Rank r owns a set of element ranges of a flat parameter group (one contiguous 1/N shard after a full all-reduce, or
// its slice of every bucket after reduce-scatter-only buckets): it runs Adam on that shard of the fp32 master / moment
// arrays (reading the already reduced 16-bit gradients of its local arena) and stores the new 16-bit parameters
// into EVERY rank's parameter arena - one multimem.st per 16-byte vector through the NVLS alias when there is one,
// peer stores otherwise.  `params` describes the symmetric PARAMETER arena.  A handshake ahead of the kernel keeps
// a fast rank from overwriting parameters a slower rank's backward still reads; a flag barrier at the end makes
// every shard visible everywhere before the next forward.
struct ShardAdam {
  float* master;            // fp32 master weights of the group (full length, indexed by absolute element)
  float* exp_avg;
  float* exp_avg_sq;
  const void* grad;         // local, reduced 16-bit gradient arena of the group
  int nranges;              // this rank's shard = union of element ranges [lo, hi), every lo % 8 == 0
  long long range_lo[kMaxShardRanges], range_hi[kMaxShardRanges];
  float beta1, beta2, eps, step_size, decay_mul;
  float inv_scale;          // gradients are multiplied by inv_scale / (*scale_dev if given)
  const float* scale_dev;   // non-finite or zero => the update is skipped on every rank (overflow)
  int stochastic_rounding;  // bf16 parameters only
  unsigned long long seed, offset, elem_base;
};
void launch_sharded_adam(const CommPeers& params, const ShardAdam& a, int dtype, int blocks, cudaStream_t stream);

}  // namespace ub
