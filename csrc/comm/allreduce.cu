// Gradient all-reduce over NVLink 5 / NVSwitch peer memory (sm_100a), replacing the NCCL all-reduce
// issued by DDP in the reference (unicore/models/distributed_unicore_model.py:37-46) on the
// gradient path.  The flat 16-bit gradient arena lives in symmetric memory (every rank maps every
// peer's buffer + an NVLS multicast alias), so the kernels below read/write peers with plain
// ld/st.global or multimem.* and synchronise with flags in a symmetric signal buffer:
//
//   one-shot : every rank reads the range from all peers and reduces locally (latency optimal,
//              used for small ranges / statistics vectors)
//   two-shot : rank r reduces its 1/N slice from all peers (reduce-scatter by peer loads) and
//              writes the result into every peer (all-gather by peer stores)
//   nvls     : rank r issues multimem.ld_reduce on its slice (reduction inside the NVSwitch, fp32
//              accumulate) and multimem.st to broadcast it - each byte crosses each link once
//
// Reduction order is the fixed rank order 0..N-1 with fp32 accumulation, so every replica ends up
// with bit-identical gradients.  An optional scale (1/world) is applied before the 16-bit store.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../common.cuh"
#include "comm_api.h"

namespace ub {

constexpr int kCommThreads = 512;

// ---- cross-GPU flag barrier (CAS put / CAS take: self-resetting, safe for back-to-back use) -----------------
UB_DEVICE void flag_put(uint32_t* addr) {
  while (atomicCAS_system(addr, 0u, 1u) != 0u) __nanosleep(32);
}
UB_DEVICE void flag_take(uint32_t* addr) {
  const long long t0 = clock64();
  while (atomicCAS_system(addr, 1u, 0u) != 1u) {
    __nanosleep(32);  // keep the polling traffic off the links the data kernels of other buckets use
    if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a peer died; do not hang the GPU forever
  }
}

// All ranks' CTAs that use flag slot `slot` meet. Slot layout in every rank's flag buffer: [slot][sender].
UB_DEVICE void slot_barrier(const CommPeers& peers, int slot, bool release_first) {
  if (release_first) __threadfence_system();
  __syncthreads();
  if (threadIdx.x < (unsigned)peers.world) {
    const int t = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(peers.flags[t]) + slot * peers.world + peers.rank;
    uint32_t* mine = reinterpret_cast<uint32_t*>(peers.flags[peers.rank]) + slot * peers.world + t;
    flag_put(remote);
    flag_take(mine);
  }
  __syncthreads();
  __threadfence_system();
}
UB_DEVICE void block_barrier(const CommPeers& peers, bool release_first) {
  slot_barrier(peers, blockIdx.x, release_first);
}

// Rendezvous of the ranks before a data kernel: ONE warp per GPU waits until every peer's stream has
// reached the same point (i.e. its producer kernels have finished).  Ranks are skewed by up to a few
// hundred microseconds inside a backward pass; parking that wait in a 32-thread kernel instead of in
// the data kernel's CTAs leaves the SMs to the compute kernels the reduction overlaps with.
constexpr int kHandshakeSlot = kMaxCommBlocks - 1;
__global__ void __launch_bounds__(32) symm_handshake_kernel(CommPeers peers) {
  slot_barrier(peers, kHandshakeSlot, /*release_first=*/false);
}

template <typename T>
UB_DEVICE void acc_add(float (&acc)[16 / sizeof(T)], const Vec16& v) {
  float t[16 / sizeof(T)];
  unpack<T>(v, t);
#pragma unroll
  for (int e = 0; e < (int)(16 / sizeof(T)); ++e) acc[e] += t[e];
}

// sum of squares over the CTA -> one atomicAdd into the local accumulator
UB_DEVICE void publish_sq(float sq, float* sq_acc) {
  if (sq_acc == nullptr) return;
  __shared__ float red[kCommThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kCommThreads / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && v != 0.f) atomicAdd(sq_acc, v);
  }
}

// ---- one-shot / two-shot ------------------------------------------------------------------------------------------
// Range = [begin_vec, end_vec) in 16-byte vectors relative to each buffer base.
constexpr int kCommUnroll = 4;  // vectors per thread in flight per peer (NVLink latency ~2-3 us)

// One-shot, in place: every rank reads the whole range from every peer, so nobody may store its result
// before all peers have finished reading - the grid covers the range with ONE vector per thread, the
// reduced value waits in registers across a second barrier.  (Host guarantees grid * 512 >= vectors.)
template <typename T>
__global__ void __launch_bounds__(kCommThreads) allreduce_oneshot_kernel(CommPeers peers, long long begin_vec,
                                                                          long long end_vec, float scale, float* sq_acc) {
  constexpr int EPV = 16 / sizeof(T);
  // (the handshake kernel ahead of us in the stream established that every rank's producers finished)
  const long long v = begin_vec + (long long)blockIdx.x * kCommThreads + threadIdx.x;
  const bool active = v < end_vec;
  float acc[EPV];
#pragma unroll
  for (int e = 0; e < EPV; ++e) acc[e] = 0.f;
  if (active) {
    Vec16 in[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < peers.world) in[p] = ld_global_v4(reinterpret_cast<const uint8_t*>(peers.buf[p]) + v * 16);
    }
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < peers.world) acc_add<T>(acc, in[p]);
    }
  }
  block_barrier(peers, /*release_first=*/false);  // all peers hold their sums in registers
  float sq = 0.f;
  if (active) {
#pragma unroll
    for (int e = 0; e < EPV; ++e) acc[e] *= scale;
    st_global_v4(reinterpret_cast<uint8_t*>(peers.buf[peers.rank]) + v * 16, pack<T>(acc));
    // every rank holds the whole result here; each counts only "its" 1/world slice so that the ranks' sums add up
    const long long n = end_vec - begin_vec, per = (n + peers.world - 1) / peers.world;
    const long long lo = begin_vec + per * peers.rank;
    if (v >= lo && v < lo + per) {
#pragma unroll
      for (int e = 0; e < EPV; ++e) sq += acc[e] * acc[e];
    }
  }
  publish_sq(sq, sq_acc);
}

// Two-shot: rank r owns slice r: reduce-scatter by peer loads, all-gather by peer stores.
// W = compile-time bound on the world size (2 / 4 / 8); W * kU = 16 peer vectors in flight per thread.
template <typename T, int W, bool kScatterOnly = false>
__global__ void __launch_bounds__(kCommThreads) allreduce_twoshot_kernel(CommPeers peers, long long begin_vec,
                                                                          long long end_vec, float scale, float* sq_acc) {
  constexpr int EPV = 16 / sizeof(T);
  constexpr int kU = 16 / W;
  const long long n = end_vec - begin_vec;
  const long long per = (n + peers.world - 1) / peers.world;
  long long lo = begin_vec + per * peers.rank;
  const long long hi = lo + per < end_vec ? lo + per : end_vec;
  if (lo > end_vec) lo = end_vec;
  const long long stride = (long long)gridDim.x * kCommThreads;
  float sq = 0.f;
  for (long long v0 = lo + (long long)blockIdx.x * kCommThreads + threadIdx.x; v0 < hi; v0 += stride * kU) {
    Vec16 in[kU][W];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long v = v0 + u * stride;
#pragma unroll
      for (int p = 0; p < W; ++p) {
        if (p < peers.world && v < hi) in[u][p] = ld_global_v4(reinterpret_cast<const uint8_t*>(peers.buf[p]) + v * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) {
        float acc[EPV];
#pragma unroll
        for (int e = 0; e < EPV; ++e) acc[e] = 0.f;
#pragma unroll
        for (int p = 0; p < W; ++p) {
          if (p < peers.world) acc_add<T>(acc, in[u][p]);
        }
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
          acc[e] *= scale;
          sq += acc[e] * acc[e];
        }
        const Vec16 out = pack<T>(acc);
        if (kScatterOnly) {
          st_global_v4(reinterpret_cast<uint8_t*>(peers.buf[peers.rank]) + v * 16, out);
        } else {
#pragma unroll
          for (int p = 0; p < W; ++p) {
            if (p < peers.world) st_global_v4(reinterpret_cast<uint8_t*>(peers.buf[p]) + v * 16, out);
          }
        }
      }
    }
  }
  publish_sq(sq, sq_acc);
  block_barrier(peers, /*release_first=*/true);  // my stores are visible at the peers before anyone proceeds
}

// ---- NVLS (multimem) --------------------------------------------------------------------------------------------------
template <typename T>
UB_DEVICE Vec16 multimem_ld_reduce(const void* mc_addr);
template <>
UB_DEVICE Vec16 multimem_ld_reduce<__half>(const void* a) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(a)
               : "memory");
  return v;
}
template <>
UB_DEVICE Vec16 multimem_ld_reduce<__nv_bfloat16>(const void* a) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(a)
               : "memory");
  return v;
}
template <>
UB_DEVICE Vec16 multimem_ld_reduce<float>(const void* a) {
  Vec16 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(a)
               : "memory");
  return v;
}
UB_DEVICE void multimem_st(void* mc_addr, const Vec16& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "r"(v.w[0]), "r"(v.w[1]),
               "r"(v.w[2]), "r"(v.w[3])
               : "memory");
}

template <typename T, bool kScatterOnly = false>
__global__ void __launch_bounds__(kCommThreads) allreduce_nvls_kernel(CommPeers peers, long long begin_vec,
                                                                        long long end_vec, float scale, float* sq_acc) {
  constexpr int EPV = 16 / sizeof(T);
  const long long n = end_vec - begin_vec;
  const long long per = (n + peers.world - 1) / peers.world;
  long long lo = begin_vec + per * peers.rank;
  long long hi = lo + per < end_vec ? lo + per : end_vec;
  if (lo > end_vec) lo = end_vec;
  uint8_t* mc = reinterpret_cast<uint8_t*>(peers.multicast);
  const long long stride = (long long)gridDim.x * kCommThreads;
  float sq = 0.f;
  for (long long v0 = lo + (long long)blockIdx.x * kCommThreads + threadIdx.x; v0 < hi; v0 += stride * kCommUnroll) {
    Vec16 r[kCommUnroll];
#pragma unroll
    for (int u = 0; u < kCommUnroll; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) r[u] = multimem_ld_reduce<T>(mc + v * 16);
    }
#pragma unroll
    for (int u = 0; u < kCommUnroll; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) {
        if (scale != 1.f || sq_acc != nullptr) {
          float acc[EPV];
          unpack<T>(r[u], acc);
#pragma unroll
          for (int e = 0; e < EPV; ++e) {
            acc[e] *= scale;
            sq += acc[e] * acc[e];
          }
          r[u] = pack<T>(acc);
        }
        if (kScatterOnly) st_global_v4(reinterpret_cast<uint8_t*>(peers.buf[peers.rank]) + v * 16, r[u]);
        else multimem_st(mc + v * 16, r[u]);
      }
    }
  }
  publish_sq(sq, sq_acc);
  block_barrier(peers, true);
}

// ---- sharded Adam + parameter all-gather ----------------------------------------------------------------------------
UB_DEVICE void shard_adam_math(float& p, float& m, float& v, float g, const ShardAdam& a) {
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  p = p * a.decay_mul - a.step_size * (m / (sqrtf(v) + a.eps));
}
UB_DEVICE uint32_t shard_bf16_sr(float x, uint32_t rnd16) {
  uint32_t bits = __float_as_uint(x);
  if ((bits & 0x7f800000u) != 0x7f800000u) bits += rnd16;  // inf / nan stay as they are
  return bits >> 16;
}

template <typename T>
UB_DEVICE void shard_adam_range(const CommPeers& params, const ShardAdam& a, long long lo, long long hi, float gmul) {
  const T* G = reinterpret_cast<const T*>(a.grad);
  const bool sr = a.stochastic_rounding != 0 && sizeof(T) == 2;
  const long long vend = lo + ((hi - lo) & ~7ll);
  const long long stride = (long long)gridDim.x * kCommThreads * 8;
  for (long long i = lo + ((long long)blockIdx.x * kCommThreads + threadIdx.x) * 8; i < vend; i += stride) {
    float g[8], p[8], m[8], v[8];
    unpack<T>(ld_global_nc_v4(G + i), g);
    const Vec16 p0 = ld_global_v4(a.master + i), p1 = ld_global_v4(a.master + i + 4);
    const Vec16 m0 = ld_global_v4(a.exp_avg + i), m1 = ld_global_v4(a.exp_avg + i + 4);
    const Vec16 v0 = ld_global_v4(a.exp_avg_sq + i), v1 = ld_global_v4(a.exp_avg_sq + i + 4);
    unpack<float>(p0, p);
    unpack<float>(p1, p + 4);
    unpack<float>(m0, m);
    unpack<float>(m1, m + 4);
    unpack<float>(v0, v);
    unpack<float>(v1, v + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) shard_adam_math(p[k], m[k], v[k], g[k] * gmul, a);
    st_global_v4(a.master + i, pack<float>(p));
    st_global_v4(a.master + i + 4, pack<float>(p + 4));
    st_global_v4(a.exp_avg + i, pack<float>(m));
    st_global_v4(a.exp_avg + i + 4, pack<float>(m + 4));
    st_global_v4(a.exp_avg_sq + i, pack<float>(v));
    st_global_v4(a.exp_avg_sq + i + 4, pack<float>(v + 4));
    Vec16 o;
    if (sr) {
      const Philox4 r = philox4x32_10(a.seed, a.offset, (a.elem_base + (unsigned long long)i) >> 3);
      const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o.w[k] = shard_bf16_sr(p[2 * k], rw[k] & 0xffffu) | (shard_bf16_sr(p[2 * k + 1], rw[k] >> 16) << 16);
    } else {
      o = pack<T>(p);
    }
    // the all-gather IS this store: one instruction through the switch, or one store per peer
    if (params.multicast != nullptr) {
      multimem_st(reinterpret_cast<T*>(params.multicast) + i, o);
    } else {
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r) {
        if (r < params.world) st_global_v4(reinterpret_cast<T*>(params.buf[r]) + i, o);
      }
    }
  }
  // scalar tail (group length not a multiple of 8; at most one range has one): plain peer stores
  if (blockIdx.x == 0) {
    for (long long i = vend + threadIdx.x; i < hi; i += kCommThreads) {
      float p = a.master[i], m = a.exp_avg[i], v = a.exp_avg_sq[i];
      shard_adam_math(p, m, v, to_f32<T>(G[i]) * gmul, a);
      a.master[i] = p;
      a.exp_avg[i] = m;
      a.exp_avg_sq[i] = v;
      const T out = from_f32<T>(p);
      for (int r = 0; r < params.world; ++r) reinterpret_cast<T*>(params.buf[r])[i] = out;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kCommThreads) sharded_adam_kernel(CommPeers params, ShardAdam a) {
  const float sdev = a.scale_dev ? __ldg(a.scale_dev) : 1.f;
  const bool skip = a.scale_dev != nullptr && !(isfinite(sdev) && sdev != 0.f);
  if (!skip) {
    const float gmul = a.inv_scale / sdev;
    for (int r = 0; r < a.nranges; ++r) shard_adam_range<T>(params, a, a.range_lo[r], a.range_hi[r], gmul);
  }
  block_barrier(params, /*release_first=*/true);  // every shard has landed everywhere before anyone proceeds
}

void launch_sharded_adam(const CommPeers& params, const ShardAdam& a, int dtype, int blocks, cudaStream_t stream) {
  if (blocks <= 0) blocks = 48;
  if (blocks > kMaxCommBlocks - 1) blocks = kMaxCommBlocks - 1;
  symm_handshake_kernel<<<1, 32, 0, stream>>>(params);  // nobody still reads the old parameters
  if (dtype == kF16) sharded_adam_kernel<__half><<<blocks, kCommThreads, 0, stream>>>(params, a);
  else if (dtype == kBF16) sharded_adam_kernel<__nv_bfloat16><<<blocks, kCommThreads, 0, stream>>>(params, a);
}

// ---- host --------------------------------------------------------------------------------------------------------------
template <typename T>
static void run_allreduce(const CommPeers& peers, long long begin_vec, long long end_vec, float scale, int algo,
                          int blocks, float* sq_acc, cudaStream_t stream, bool scatter_only) {
  symm_handshake_kernel<<<1, 32, 0, stream>>>(peers);
  if (scatter_only) {  // reduce-scatter half only (sharded optimizer); one-shot has no such form
    if (algo == kAlgoNvls)
      allreduce_nvls_kernel<T, true><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
    else if (peers.world <= 2)
      allreduce_twoshot_kernel<T, 2, true><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
    else if (peers.world <= 4)
      allreduce_twoshot_kernel<T, 4, true><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
    else
      allreduce_twoshot_kernel<T, 8, true><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
    return;
  }
  if (algo == kAlgoNvls) {
    allreduce_nvls_kernel<T><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
  } else if (algo == kAlgoTwoShot) {
    if (peers.world <= 2)
      allreduce_twoshot_kernel<T, 2><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
    else if (peers.world <= 4)
      allreduce_twoshot_kernel<T, 4><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
    else
      allreduce_twoshot_kernel<T, 8><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
  } else {
    allreduce_oneshot_kernel<T><<<blocks, kCommThreads, 0, stream>>>(peers, begin_vec, end_vec, scale, sq_acc);
  }
}

int pick_allreduce_algo(long long bytes, int world, bool has_multicast) {
  if (bytes <= 256 * 1024) return kAlgoOneShot;
  if (has_multicast && world > 2) return kAlgoNvls;
  return kAlgoTwoShot;
}

void launch_allreduce(const CommPeers& peers, long long byte_offset, long long bytes, int dtype, float scale, int algo,
                      int blocks, float* sq_acc, cudaStream_t stream, bool scatter_only) {
  const long long begin_vec = byte_offset / 16, end_vec = (byte_offset + bytes) / 16;
  if (end_vec <= begin_vec) return;
  if (algo == kAlgoAuto) algo = pick_allreduce_algo(bytes, peers.world, peers.multicast != nullptr);
  if (scatter_only && algo == kAlgoOneShot) algo = peers.multicast != nullptr && peers.world > 2 ? kAlgoNvls : kAlgoTwoShot;
  if (algo == kAlgoNvls && peers.multicast == nullptr) algo = kAlgoTwoShot;
  const long long one_shot_blocks = (end_vec - begin_vec + kCommThreads - 1) / kCommThreads;
  if (algo == kAlgoOneShot && one_shot_blocks > kMaxCommBlocks - 1) algo = kAlgoTwoShot;  // range too large
  if (algo == kAlgoOneShot) {
    blocks = (int)one_shot_blocks;  // exactly one vector per thread (see kernel)
  } else {
    if (blocks <= 0) blocks = 24;
    if (blocks > kMaxCommBlocks - 1) blocks = kMaxCommBlocks - 1;  // the last slot belongs to the handshake
  }
  if (dtype == kF32) run_allreduce<float>(peers, begin_vec, end_vec, scale, algo, blocks, sq_acc, stream, scatter_only);
  else if (dtype == kF16) run_allreduce<__half>(peers, begin_vec, end_vec, scale, algo, blocks, sq_acc, stream, scatter_only);
  else run_allreduce<__nv_bfloat16>(peers, begin_vec, end_vec, scale, algo, blocks, sq_acc, stream, scatter_only);
}

}  // namespace ub
