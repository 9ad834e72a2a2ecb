// Host launchers of the tcgen05 fused attention kernels (csrc/attn/fmha_*_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

struct FmhaFwdParams {
  const void* q;   // [B, Lq, H, 64] via strides (elements)
  const void* k;   // [B, Lk, H, 64]
  const void* v;
  void* out;       // [B, Lq, H, 64] contiguous
  float* lse;      // [B, H, Lq]
  const void* bias;        // [bias_batch (1 or B), H, Lq, Lk] contiguous, fp32 or same 16-bit type; may be null
  const uint8_t* kpm;      // [B, Lk] bool, nonzero = masked; may be null
  long long q_sb, q_sl, q_sh, k_sb, k_sl, k_sh, v_sb, v_sl, v_sh;
  int B, H, Lq, Lk;
  int bias_batch, bias_is_f32, is_bf16;
  float scale, p_drop;
  unsigned long long seed, offset;
};
void launch_fmha_fwd(const FmhaFwdParams& p, cudaStream_t stream);

struct FmhaBwdParams {
  FmhaFwdParams f;   // same inputs as forward (out/lse hold the forward results)
  const void* dout;  // [B, Lq, H, 64] contiguous
  float* delta;      // [B, H, Lq] scratch: rowsum(dO * O)
  float* dq_acc;     // [B, Lq, H, 64] fp32, zero-initialised (atomically accumulated)
  void* dq;          // [B, Lq, H, 64] 16-bit result
  void* dk;          // [B, Lk, H, 64]
  void* dv;
  float* dbias;      // [bias_batch, H, Lq, Lk] fp32, zero-initialised; null when not needed
};
void launch_fmha_bwd(const FmhaBwdParams& p, cudaStream_t stream);

}  // namespace ub
