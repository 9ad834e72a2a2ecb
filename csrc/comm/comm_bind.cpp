// Bindings of the peer-memory collective kernels.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "../api.h"
#include "comm_api.h"

namespace {

// buffer_ptrs / flag_ptrs: integer device addresses of every rank's (peer-mapped) buffers as returned
// by the symmetric-memory rendezvous; multicast_ptr: NVLS alias or 0.
void symm_allreduce(const std::vector<int64_t>& buffer_ptrs, const std::vector<int64_t>& flag_ptrs,
                    int64_t multicast_ptr, int64_t rank, int64_t byte_offset, int64_t bytes, int64_t dtype,
                    double scale, int64_t algo, int64_t blocks, int64_t sq_acc_ptr) {
  const int world = (int)buffer_ptrs.size();
  TORCH_CHECK(world >= 1 && world <= ub::kMaxPeers, "world size must be in [1, ", ub::kMaxPeers, "]");
  TORCH_CHECK((int)flag_ptrs.size() == world && rank >= 0 && rank < world);
  TORCH_CHECK(byte_offset % 16 == 0 && bytes % 16 == 0, "range must be 16-byte aligned");
  ub::CommPeers peers{};
  for (int i = 0; i < world; ++i) {
    peers.buf[i] = reinterpret_cast<void*>(buffer_ptrs[i]);
    peers.flags[i] = reinterpret_cast<void*>(flag_ptrs[i]);
  }
  peers.multicast = reinterpret_cast<void*>(multicast_ptr);
  peers.rank = (int)rank;
  peers.world = world;
  ub::launch_allreduce(peers, byte_offset, bytes, (int)dtype, (float)scale, (int)algo, (int)blocks,
                       reinterpret_cast<float*>(sq_acc_ptr), at::cuda::getCurrentCUDAStream().stream());
  cudaError_t err = cudaGetLastError();
  TORCH_CHECK(err == cudaSuccess, "symm_allreduce launch failed: ", cudaGetErrorString(err));
}

int64_t symm_pick_algo(int64_t bytes, int64_t world, bool has_multicast) {
  return ub::pick_allreduce_algo(bytes, (int)world, has_multicast);
}

}  // namespace

void register_comm(pybind11::module_& m) {
  m.def("symm_allreduce", &symm_allreduce);
  m.def("symm_pick_algo", &symm_pick_algo);
  m.attr("SYMM_MAX_BLOCKS") = (int64_t)ub::kMaxCommBlocks;
  m.attr("SYMM_MAX_PEERS") = (int64_t)ub::kMaxPeers;
}
