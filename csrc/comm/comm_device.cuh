// Device-side building blocks shared by the peer-memory collective kernels (allreduce.cu, fused_step.cu):
// cross-GPU flag barriers over a symmetric signal buffer, NVLS multimem wrappers, fp32 accumulation helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../common.cuh"
#include "comm_api.h"

namespace ub {

constexpr int kCommThreads = 512;

// ---- host-visible error word -------------------------------------------------------------------------------------
// A peer that never arrives (crashed process, mismatched collective order) must not hang or trap the GPU: the waiting
// thread gives up after ~10 s, records WHY in a word of pinned host memory (`peers.err`, mapped into the device
// address space), and the kernel finishes without touching data.  Every later collective kernel sees the word and
// returns immediately; the host raises from `SymmAllReduce.check_health()` at its next synchronisation point.
enum CommError : uint32_t { kCommOk = 0, kCommTimeout = 1, kCommTagMismatch = 2 };

UB_DEVICE bool comm_failed(const CommPeers& peers) {
  return peers.err != nullptr && *reinterpret_cast<volatile uint32_t*>(peers.err) != kCommOk;
}
UB_DEVICE void comm_fail(const CommPeers& peers, uint32_t code, uint32_t detail) {
  if (peers.err == nullptr) __trap();  // no error channel configured: the old behaviour
  volatile uint32_t* e = reinterpret_cast<volatile uint32_t*>(peers.err);
  if (e[0] == kCommOk) {
    e[1] = detail;
    e[2] = (uint32_t)peers.rank;
    __threadfence_system();
    e[0] = code;
  }
  __threadfence_system();
}

// ---- cross-GPU flag barrier (CAS put / CAS take: self-resetting, safe for back-to-back use) -----------------
// A flag carries the caller's TAG (bucket index / kernel kind, never zero) instead of a bare 1: two ranks that pair up
// on different collectives - e.g. gradient buckets launched in a different order - see a foreign tag and fail loudly
// instead of reducing unrelated data.
// The error word lives in HOST memory (a read costs a PCIe round trip): it is looked at only by a thread that has been
// spinning for a while, never on the fast path - a healthy barrier costs exactly one flag round trip over NVLink.
constexpr int kSpinsPerHealthCheck = 2048;

UB_DEVICE bool flag_put(const CommPeers& peers, uint32_t* addr, uint32_t tag) {
  const long long t0 = clock64();
  int spins = 0;
  while (atomicCAS_system(addr, 0u, tag) != 0u) {
    __nanosleep(32);
    if (++spins % kSpinsPerHealthCheck == 0) {
      if (comm_failed(peers)) return false;  // somebody (this GPU, earlier) already gave up: do not wait again
      if (clock64() - t0 > 20000000000LL) {
        comm_fail(peers, kCommTimeout, tag);
        return false;
      }
    }
  }
  return true;
}
UB_DEVICE bool flag_take(const CommPeers& peers, uint32_t* addr, uint32_t tag) {
  const long long t0 = clock64();
  int spins = 0;
  for (;;) {
    const uint32_t seen = atomicCAS_system(addr, tag, 0u);
    if (seen == tag) return true;
    if (seen != 0u) {  // a peer is inside a DIFFERENT collective
      comm_fail(peers, kCommTagMismatch, (tag << 16) | (seen & 0xffffu));
      return false;
    }
    __nanosleep(32);  // keep the polling traffic off the links the data kernels of other buckets use
    if (++spins % kSpinsPerHealthCheck == 0) {
      if (comm_failed(peers)) return false;
      if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; do not hang the GPU forever
        comm_fail(peers, kCommTimeout, tag);
        return false;
      }
    }
  }
}

// Reserved flag slots (layout in every rank's flag buffer: [slot][sender] uint32).  Data kernels use slot == blockIdx.x.
constexpr int kHandshakeSlot = kMaxCommBlocks - 1;
constexpr int kTailSlot = kMaxCommBlocks - 2;    // fused optimizer tail (one CTA talks to the peers)
constexpr int kStatsSlot = kMaxCommBlocks - 3;   // stand-alone statistics reduction (may run on another stream)
constexpr int kMaxDataBlocks = kMaxCommBlocks - 3;

// All ranks' CTAs that use flag slot `slot` meet.  Returns false when the communicator is (or just became) broken.
UB_DEVICE bool slot_barrier(const CommPeers& peers, int slot, bool release_first) {
  __shared__ int ok_s;
  if (release_first) __threadfence_system();
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();
  if (threadIdx.x < (unsigned)peers.world) {
    const int t = threadIdx.x;
    uint32_t* remote = reinterpret_cast<uint32_t*>(peers.flags[t]) + slot * peers.world + peers.rank;
    uint32_t* mine = reinterpret_cast<uint32_t*>(peers.flags[peers.rank]) + slot * peers.world + t;
    const uint32_t tag = peers.tag == 0u ? 1u : peers.tag;
    if (!flag_put(peers, remote, tag) || !flag_take(peers, mine, tag)) ok_s = 0;
  }
  __syncthreads();
  __threadfence_system();
  return ok_s != 0;
}
UB_DEVICE bool block_barrier(const CommPeers& peers, bool release_first) {
  return slot_barrier(peers, blockIdx.x, release_first);
}

template <typename T>
UB_DEVICE void acc_add(float (&acc)[16 / sizeof(T)], const Vec16& v) {
  float t[16 / sizeof(T)];
  unpack<T>(v, t);
#pragma unroll
  for (int e = 0; e < (int)(16 / sizeof(T)); ++e) acc[e] += t[e];
}

// CTA-wide sum (kCommThreads threads); result valid in thread 0
UB_DEVICE float comm_block_sum(float v) {
  __shared__ float red[kCommThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < kCommThreads / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;
}

// ---- NVLS (multimem) -------------------------------------------------------------------------------------------------
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

}  // namespace ub
