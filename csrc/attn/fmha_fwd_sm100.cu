// Fused multi-head attention forward for sm_100a (head_dim 64, fp16/bf16):
//   O = dropout(softmax(scale * Q K^T + bias + key_padding)) V        LSE = logsumexp of the logits
//
// Tensor-core path: tcgen05.mma (cta_group::1, M=128) with fp32 accumulators in tensor memory.
//   S (128 x 128 fp32) lives in TMEM columns [0,128), O (128 x 64 fp32) in columns [128,192).
//   One CTA = one 128-row query tile of one (batch, head); it walks the key tiles (128 keys each):
//     1. K_j, V_j -> shared memory (8x8 core-matrix layout, see tcgen05.cuh)
//     2. S  = Q K_j^T          4 x tcgen05.mma (K=16 each), commit -> mbarrier
//     3. softmax: thread t owns query row t (TMEM lane t): tcgen05.ld the row, add bias / masks,
//        online max/sum, Philox dropout, write P (16-bit) to shared memory, rescale O in TMEM
//     4. O += P V_j            8 x tcgen05.mma (V presented MN-major from the same row-major bytes)
//   q/k/v are read through strides straight out of the packed in_proj output; O is written as
//   [B, Lq, H, 64] so that out_proj consumes it without a transpose.
// Two CTAs are resident per SM (80 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps
// the other's loads and MMAs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include <type_traits>

#include "../common.cuh"
#include "fmha_api.h"
#include "tcgen05.cuh"

namespace ub {

using namespace tc;

constexpr int kBlockM = 128;   // query rows per CTA
constexpr int kBlockN = 128;   // keys per tile
constexpr int kHeadDim = 64;
constexpr int kFwdThreads = 128;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemColS = 0, kTmemColO = 128;

constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = 16384;
constexpr uint32_t kSmemV = 32768;
constexpr uint32_t kSmemP = 49152;            // 128 x 128 x 2 = 32768 bytes
constexpr uint32_t kSmemBar = 81920;          // 2 mbarriers + tmem base
constexpr uint32_t kFwdSmemBytes = 81920 + 64;

// copy a [128 rows x 64] 16-bit tile (row stride `row_stride` elements) into core-matrix layout.
// Rows >= valid_rows are zero filled.
template <typename T>
UB_DEVICE void load_tile64(uint8_t* smem_tile, const T* gbase, long long row_stride, int valid_rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r_in8 = lane & 7, c_lo = lane >> 3;
  Vec16 regs[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int u = it * 4 + warp;
    const int row = (u >> 1) * 8 + r_in8;
    const int c = (u & 1) * 4 + c_lo;
    if (row < valid_rows) {
      regs[it] = ld_global_nc_v4(gbase + (long long)row * row_stride + c * 8);
    } else {
      regs[it].w[0] = regs[it].w[1] = regs[it].w[2] = regs[it].w[3] = 0u;
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int u = it * 4 + warp;
    const int row = (u >> 1) * 8 + r_in8;
    const int c = (u & 1) * 4 + c_lo;
    *reinterpret_cast<Vec16*>(smem_tile + tile64_off(row, c)) = regs[it];
  }
}

template <typename T>
UB_DEVICE uint32_t pack2(float a, float b);
template <>
UB_DEVICE uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
UB_DEVICE uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// logits of 32 consecutive keys of this thread's row: s = acc * scale + bias, -inf where masked
template <typename T, bool kBiasF32>
UB_DEVICE void logits32(const uint32_t (&acc)[32], float (&s)[32], float scale, const void* bias_row, int key0,
                        const uint8_t* kpm_row, int Lk, bool row_valid) {
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(acc[i]) * scale;
  if (bias_row != nullptr && row_valid) {
    if (kBiasF32) {
      const float* bp = reinterpret_cast<const float*>(bias_row) + key0;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        if (key0 + v * 4 < Lk) {
          const Vec16 b = ld_global_v4(bp + v * 4);
#pragma unroll
          for (int e = 0; e < 4; ++e) s[v * 4 + e] += __uint_as_float(b.w[e]);
        }
      }
    } else {
      const T* bp = reinterpret_cast<const T*>(bias_row) + key0;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (key0 + v * 8 < Lk) {
          float t[8];
          unpack<T>(ld_global_v4(bp + v * 8), t);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[v * 8 + e] += t[e];
        }
      }
    }
  }
  if (kpm_row != nullptr) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (key0 + v * 8 < Lk) {
        const uint2 m = *reinterpret_cast<const uint2*>(kpm_row + key0 + v * 8);  // 8 bool bytes
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if ((m.x >> (8 * e)) & 0xffu) s[v * 8 + e] = -CUDART_INF_F;
          if ((m.y >> (8 * e)) & 0xffu) s[v * 8 + 4 + e] = -CUDART_INF_F;
        }
      }
    }
  }
  if (key0 + 32 > Lk) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (key0 + i >= Lk) s[i] = -CUDART_INF_F;
  }
}

template <typename T, bool kBiasF32>
__global__ void __launch_bounds__(kFwdThreads, 2) fmha_fwd_kernel(FmhaFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * kBlockM, h = blockIdx.y, b = blockIdx.z;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_s = smem_base + kSmemBar, bar_o = smem_base + kSmemBar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 16);

  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbarrier_init();
  }
  const T* qg = reinterpret_cast<const T*>(p.q) + (long long)b * p.q_sb + (long long)h * p.q_sh + (long long)q0 * p.q_sl;
  const T* kg = reinterpret_cast<const T*>(p.k) + (long long)b * p.k_sb + (long long)h * p.k_sh;
  const T* vg = reinterpret_cast<const T*>(p.v) + (long long)b * p.v_sb + (long long)h * p.v_sh;
  const int q_valid = min(kBlockM, p.Lq - q0);
  load_tile64<T>(smem + kSmemQ, qg, p.q_sl, q_valid);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);  // this warp's TMEM lanes
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t idesc_qk = make_idesc_f16(kBlockM, kBlockN, kFmt, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_f16(kBlockM, kHeadDim, kFmt, 0, 1);

  const int row = q0 + tid;
  const bool row_valid = row < p.Lq;
  const void* bias_row = nullptr;
  if (p.bias != nullptr) {
    const long long boff = ((long long)(p.bias_batch > 1 ? b : 0) * p.H + h) * p.Lq + (row_valid ? row : 0);
    bias_row = kBiasF32 ? (const void*)(reinterpret_cast<const float*>(p.bias) + boff * p.Lk)
                        : (const void*)(reinterpret_cast<const T*>(p.bias) + boff * p.Lk);
  }
  const uint8_t* kpm_row = p.kpm != nullptr ? p.kpm + (long long)b * p.Lk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint32_t thresh = dropout_thresh16(p.p_drop);
  const float keep_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  const unsigned long long drop_row_base = (((unsigned long long)b * p.H + h) * p.Lq + (row_valid ? row : 0)) * p.Lk;

  constexpr float kLog2e = 1.4426950408889634f;
  float m_run = -CUDART_INF_F, l_run = 0.f;
  uint32_t phase_s = 0, phase_o = 0;
  const int n_tiles = (p.Lk + kBlockN - 1) / kBlockN;

  for (int j = 0; j < n_tiles; ++j) {
    const int key_tile0 = j * kBlockN;
    if (j > 0) {  // previous P V must be done before K/V/P shared memory is overwritten
      mbar_wait(bar_o, phase_o);
      phase_o ^= 1;
      fence_after_thread_sync();
    }
    const int k_valid = min(kBlockN, p.Lk - key_tile0);
    load_tile64<T>(smem + kSmemK, kg + (long long)key_tile0 * p.k_sl, p.k_sl, k_valid);
    load_tile64<T>(smem + kSmemV, vg + (long long)key_tile0 * p.v_sl, p.v_sl, k_valid);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_thread_sync();
#pragma unroll
      for (int kk = 0; kk < kHeadDim / 16; ++kk) {
        const uint64_t da = make_smem_desc(smem_base + kSmemQ + kk * 256, 128, 1024);
        const uint64_t db = make_smem_desc(smem_base + kSmemK + kk * 256, 128, 1024);
        umma_f16_ss(tmem_base + kTmemColS, da, db, idesc_qk, kk > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, phase_s);
    phase_s ^= 1;
    fence_after_thread_sync();

    // ---- pass 1: row maximum -------------------------------------------------------------------
    float m_tile = -CUDART_INF_F;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t acc[32];
      float s[32];
      tmem_ld32(lane_base + kTmemColS + c * 32, acc);
      tmem_wait_ld();
      logits32<T, kBiasF32>(acc, s, p.scale, bias_row, key_tile0 + c * 32, kpm_row, p.Lk, row_valid);
#pragma unroll
      for (int i = 0; i < 32; ++i) m_tile = fmaxf(m_tile, s[i]);
    }
    const float m_new = fmaxf(m_run, m_tile);
    const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
    const float alpha = exp2f((m_run - m_use) * kLog2e);  // m_run = -inf -> 0
    l_run *= alpha;
    m_run = m_new;

    // ---- pass 2: probabilities, dropout, P -> shared memory ---------------------------------------
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t acc[32];
      float s[32];
      tmem_ld32(lane_base + kTmemColS + c * 32, acc);
      tmem_wait_ld();
      logits32<T, kBiasF32>(acc, s, p.scale, bias_row, key_tile0 + c * 32, kpm_row, p.Lk, row_valid);
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s[i] = exp2f((s[i] - m_use) * kLog2e);
        psum += s[i];
      }
      l_run += psum;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (drop) {
          const unsigned long long idx = drop_row_base + (unsigned long long)(key_tile0 + c * 32 + v * 8);
          const uint32_t keep = dropout_keep8(p.seed, p.offset, idx >> 3, thresh);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[v * 8 + e] = ((keep >> e) & 1u) ? s[v * 8 + e] * keep_scale : 0.f;
        }
        Vec16 o;
#pragma unroll
        for (int e = 0; e < 4; ++e) o.w[e] = pack2<T>(s[v * 8 + 2 * e], s[v * 8 + 2 * e + 1]);
        *reinterpret_cast<Vec16*>(smem + kSmemP + tile128_off(tid, c * 4 + v)) = o;
      }
    }

    // ---- rescale the running output accumulator ------------------------------------------------------
    if (j > 0) {
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t acc[32];
        tmem_ld32(lane_base + kTmemColO + c * 32, acc);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) * alpha);
        tmem_st32(lane_base + kTmemColO + c * 32, acc);
      }
      tmem_wait_st();
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_thread_sync();
#pragma unroll
      for (int kk = 0; kk < kBlockN / 16; ++kk) {
        // A = P [128 q x 128 keys] K-major: 16 keys per step = 2 core matrices of 128 B
        const uint64_t da = make_smem_desc(smem_base + kSmemP + kk * 256, 128, 2048);
        // B = V [64 d x 128 keys] MN-major view of the row-major [key][d] tile: 16 keys = 2 x 1024 B
        const uint64_t db = make_smem_desc(smem_base + kSmemV + kk * 2048, 1024, 128);
        umma_f16_ss(tmem_base + kTmemColO, da, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------
  mbar_wait(bar_o, phase_o);
  fence_after_thread_sync();
  const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
  T* og = reinterpret_cast<T*>(p.out) + (((long long)b * p.Lq + row) * p.H + h) * kHeadDim;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t acc[32];
    tmem_ld32(lane_base + kTmemColO + c * 32, acc);
    tmem_wait_ld();
    if (row_valid) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        Vec16 o;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o.w[e] = pack2<T>(__uint_as_float(acc[v * 8 + 2 * e]) * inv_l, __uint_as_float(acc[v * 8 + 2 * e + 1]) * inv_l);
        st_global_v4(og + c * 32 + v * 8, o);
      }
    }
  }
  if (row_valid) {
    p.lse[((long long)b * p.H + h) * p.Lq + row] = (l_run > 0.f) ? m_run + __logf(l_run) : -CUDART_INF_F;
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

void launch_fmha_fwd(const FmhaFwdParams& p, cudaStream_t stream) {
  dim3 grid((p.Lq + kBlockM - 1) / kBlockM, p.H, p.B);
#define UB_FMHA_FWD_LAUNCH(T, BF32)                                                                            \
  do {                                                                                                         \
    auto kern = fmha_fwd_kernel<T, BF32>;                                                                      \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes);               \
    kern<<<grid, kFwdThreads, kFwdSmemBytes, stream>>>(p);                                                     \
  } while (0)
  if (p.is_bf16) {
    if (p.bias_is_f32) UB_FMHA_FWD_LAUNCH(__nv_bfloat16, true);
    else UB_FMHA_FWD_LAUNCH(__nv_bfloat16, false);
  } else {
    if (p.bias_is_f32) UB_FMHA_FWD_LAUNCH(__half, true);
    else UB_FMHA_FWD_LAUNCH(__half, false);
  }
#undef UB_FMHA_FWD_LAUNCH
}

}  // namespace ub
