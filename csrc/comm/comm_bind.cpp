// Bindings of the peer-memory collective kernels.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cmath>
#include <optional>
#include <vector>

#include "../api.h"
#include "comm_api.h"

namespace {

// buffer_ptrs / flag_ptrs: integer device addresses of every rank's (peer-mapped) buffers as returned
// by the symmetric-memory rendezvous; multicast_ptr: NVLS alias or 0.
void symm_reduce_impl(const std::vector<int64_t>& buffer_ptrs, const std::vector<int64_t>& flag_ptrs,
                      int64_t multicast_ptr, int64_t rank, int64_t byte_offset, int64_t bytes, int64_t dtype,
                      double scale, int64_t algo, int64_t blocks, int64_t sq_acc_ptr, bool scatter_only) {
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
                       reinterpret_cast<float*>(sq_acc_ptr), at::cuda::getCurrentCUDAStream().stream(), scatter_only);
  cudaError_t err = cudaGetLastError();
  TORCH_CHECK(err == cudaSuccess, "symm_allreduce launch failed: ", cudaGetErrorString(err));
}

void symm_allreduce(const std::vector<int64_t>& buffer_ptrs, const std::vector<int64_t>& flag_ptrs,
                    int64_t multicast_ptr, int64_t rank, int64_t byte_offset, int64_t bytes, int64_t dtype,
                    double scale, int64_t algo, int64_t blocks, int64_t sq_acc_ptr) {
  symm_reduce_impl(buffer_ptrs, flag_ptrs, multicast_ptr, rank, byte_offset, bytes, dtype, scale, algo, blocks,
                   sq_acc_ptr, false);
}

// Reduce-scatter half only: afterwards rank r's buffer holds the reduced values of ITS slice of the range
// (16-byte vectors [begin + r*per, begin + (r+1)*per), per = ceil(n_vectors / world)); the rest is stale.
void symm_reduce_scatter(const std::vector<int64_t>& buffer_ptrs, const std::vector<int64_t>& flag_ptrs,
                         int64_t multicast_ptr, int64_t rank, int64_t byte_offset, int64_t bytes, int64_t dtype,
                         double scale, int64_t blocks, int64_t sq_acc_ptr) {
  symm_reduce_impl(buffer_ptrs, flag_ptrs, multicast_ptr, rank, byte_offset, bytes, dtype, scale, ub::kAlgoAuto, blocks,
                   sq_acc_ptr, true);
}

// One flat parameter group: Adam on this rank's shard (a list of element ranges) + all-gather of the new 16-bit parameters by the
// kernel's own stores (see comm_api.h).  param_ptrs / multicast_ptr describe the symmetric PARAMETER arena of the
// group; `grad` is this rank's (already reduced) gradient arena; master / exp_avg / exp_avg_sq are full-length fp32.
void symm_sharded_adam(const std::vector<int64_t>& param_ptrs, const std::vector<int64_t>& flag_ptrs,
                       int64_t multicast_ptr, int64_t rank, const at::Tensor& grad, at::Tensor master,
                       at::Tensor exp_avg, at::Tensor exp_avg_sq, const std::vector<int64_t>& range_lo,
                       const std::vector<int64_t>& range_hi, double lr, double beta1,
                       double beta2, double eps, int64_t step, bool bias_correction, double weight_decay,
                       double grad_scale, const std::optional<at::Tensor>& scale_dev, bool stochastic_rounding,
                       int64_t seed, int64_t offset, int64_t blocks) {
  const int world = (int)param_ptrs.size();
  TORCH_CHECK(world >= 1 && world <= ub::kMaxPeers && (int)flag_ptrs.size() == world && rank >= 0 && rank < world);
  TORCH_CHECK(grad.is_cuda() && grad.is_contiguous() &&
              (grad.scalar_type() == at::kHalf || grad.scalar_type() == at::kBFloat16), "16-bit gradients expected");
  for (const at::Tensor* t : {&master, &exp_avg, &exp_avg_sq}) {
    TORCH_CHECK(t->is_cuda() && t->is_contiguous() && t->scalar_type() == at::kFloat, "fp32 optimizer state expected");
    TORCH_CHECK((reinterpret_cast<uintptr_t>(t->data_ptr()) & 15) == 0, "optimizer state must be 16-byte aligned");
  }
  const int64_t n = master.numel();
  TORCH_CHECK(exp_avg.numel() == n && exp_avg_sq.numel() == n && grad.numel() >= n);
  TORCH_CHECK(range_lo.size() == range_hi.size() && (int)range_lo.size() <= ub::kMaxShardRanges, "too many shard ranges");
  for (size_t i = 0; i < range_lo.size(); ++i)
    TORCH_CHECK(0 <= range_lo[i] && range_lo[i] <= range_hi[i] && range_hi[i] <= n &&
                    (range_lo[i] % 8 == 0 || range_lo[i] == range_hi[i]),
                "shard ranges must start on 16-byte boundaries inside the group");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(grad.data_ptr()) & 15) == 0, "gradient arena must be 16-byte aligned");
  const c10::cuda::CUDAGuard guard(master.device());
  ub::CommPeers peers{};
  for (int i = 0; i < world; ++i) {
    TORCH_CHECK((param_ptrs[i] & 15) == 0, "parameter arenas must be 16-byte aligned");
    peers.buf[i] = reinterpret_cast<void*>(param_ptrs[i]);
    peers.flags[i] = reinterpret_cast<void*>(flag_ptrs[i]);
  }
  peers.multicast = reinterpret_cast<void*>(multicast_ptr);
  peers.rank = (int)rank;
  peers.world = world;
  ub::ShardAdam a{};
  a.master = master.data_ptr<float>();
  a.exp_avg = exp_avg.data_ptr<float>();
  a.exp_avg_sq = exp_avg_sq.data_ptr<float>();
  a.grad = grad.data_ptr();
  a.nranges = (int)range_lo.size();
  for (int i = 0; i < a.nranges; ++i) {
    a.range_lo[i] = range_lo[i];
    a.range_hi[i] = range_hi[i];
  }
  double step_size = lr;
  if (bias_correction) {
    const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
    step_size = lr * std::sqrt(bc2) / bc1;
  }
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.eps = (float)eps;
  a.step_size = (float)step_size;
  a.decay_mul = (float)(1.0 - step_size * weight_decay);
  a.inv_scale = (float)(1.0 / grad_scale);
  a.scale_dev = nullptr;
  if (scale_dev.has_value() && scale_dev->defined()) {
    TORCH_CHECK(scale_dev->is_cuda() && scale_dev->scalar_type() == at::kFloat && scale_dev->numel() == 1);
    a.scale_dev = scale_dev->data_ptr<float>();
  }
  a.stochastic_rounding = stochastic_rounding && grad.scalar_type() == at::kBFloat16 ? 1 : 0;
  a.seed = (unsigned long long)seed;
  a.offset = (unsigned long long)offset;
  a.elem_base = 0;
  ub::launch_sharded_adam(peers, a, grad.scalar_type() == at::kHalf ? ub::kF16 : ub::kBF16, (int)blocks,
                          at::cuda::getCurrentCUDAStream().stream());
  cudaError_t err = cudaGetLastError();
  TORCH_CHECK(err == cudaSuccess, "symm_sharded_adam launch failed: ", cudaGetErrorString(err));
}

int64_t symm_pick_algo(int64_t bytes, int64_t world, bool has_multicast) {
  return ub::pick_allreduce_algo(bytes, (int)world, has_multicast);
}

}  // namespace

void register_comm(pybind11::module_& m) {
  m.def("symm_allreduce", &symm_allreduce);
  m.def("symm_reduce_scatter", &symm_reduce_scatter);
  m.def("symm_pick_algo", &symm_pick_algo);
  m.def("symm_sharded_adam", &symm_sharded_adam);
  m.attr("SYMM_MAX_BLOCKS") = (int64_t)ub::kMaxCommBlocks;
  m.attr("SYMM_MAX_PEERS") = (int64_t)ub::kMaxPeers;
  m.attr("SYMM_MAX_SHARD_RANGES") = (int64_t)ub::kMaxShardRanges;
}
