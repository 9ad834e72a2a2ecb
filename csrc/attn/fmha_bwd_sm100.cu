// Fused multi-head attention backward for sm_100a (head_dim 64, fp16/bf16), flash-attention style:
// probabilities are recomputed from the saved log-sum-exp, the dropout mask from the Philox counters.
//
//   grid = (key tiles, H, B); one CTA owns K_j, V_j (128 keys) and walks the query tiles i:
//     S   = Q_i K_j^T                      tcgen05.mma  -> TMEM [  0,128)
//     dP  = dO_i V_j^T                     tcgen05.mma  -> TMEM [128,256)
//     P   = exp(scale*S + bias - LSE_i) ;  dS = P o (dropout(dP) - delta_i)        (256 threads:
//           thread = (query row, column half); no cross-thread reductions are needed)
//     dV_j += dropout(P)^T dO_i            tcgen05.mma  -> TMEM [256,320)   (A and B MN-major views)
//     dK_j += scale * dS^T Q_i             tcgen05.mma  -> TMEM [320,384)
//     dQ_i  = scale * dS K_j               tcgen05.mma  -> TMEM [384,448) -> red.global.add.v4.f32
//     dBias += dS                          red.global.add.v4.f32 (bias broadcast over the batch)
//   Every shared-memory tile is written once in the 8x8 core-matrix layout and presented to the
//   tensor core as K-major or MN-major by swapping descriptor strides (no transposes).
// Pre-pass:  delta = rowsum(dO o O).   Post-pass: dq = cast(dq_acc).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include <type_traits>

#include "../common.cuh"
#include "fmha_api.h"
#include "tcgen05.cuh"

namespace ub {
namespace {

using namespace tc;

constexpr int kBM = 128, kBN = 128, kD = 64;
constexpr int kBwdThreads = 256;
constexpr uint32_t kBwdTmemCols = 512;
constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;

constexpr uint32_t kOffQ = 0, kOffDO = 16384, kOffK = 32768, kOffV = 49152, kOffP = 65536, kOffDS = 98304;
constexpr uint32_t kOffBar = 131072;
constexpr uint32_t kBwdSmemBytes = 131072 + 64;

template <typename T>
UB_DEVICE void bwd_load_tile64(uint8_t* smem_tile, const T* gbase, long long row_stride, int valid_rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
  const int r_in8 = lane & 7, c_lo = lane >> 3;
  Vec16 regs[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int u = it * 8 + warp;
    const int row = (u >> 1) * 8 + r_in8;
    const int c = (u & 1) * 4 + c_lo;
    if (row < valid_rows) {
      regs[it] = ld_global_nc_v4(gbase + (long long)row * row_stride + c * 8);
    } else {
      regs[it].w[0] = regs[it].w[1] = regs[it].w[2] = regs[it].w[3] = 0u;
    }
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int u = it * 8 + warp;
    const int row = (u >> 1) * 8 + r_in8;
    const int c = (u & 1) * 4 + c_lo;
    *reinterpret_cast<Vec16*>(smem_tile + tile64_off(row, c)) = regs[it];
  }
}

template <typename T>
UB_DEVICE uint32_t bwd_pack2(float a, float b);
template <>
UB_DEVICE uint32_t bwd_pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <>
UB_DEVICE uint32_t bwd_pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

UB_DEVICE void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// s = acc*scale + bias, -inf for masked / out-of-range keys (same rule as the forward kernel)
template <typename T, bool kBiasF32>
UB_DEVICE void bwd_logits32(const uint32_t (&acc)[32], float (&s)[32], float scale, const void* bias_row, int key0,
                            const uint8_t* kpm_row, int Lk) {
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(acc[i]) * scale;
  if (bias_row != nullptr) {
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
        const uint2 m = *reinterpret_cast<const uint2*>(kpm_row + key0 + v * 8);
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

// ---- pre-pass: delta[b,h,q] = sum_d dO * O -----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fmha_delta_kernel(const T* __restrict__ dout, const T* __restrict__ out,
                                                           float* __restrict__ delta, int B, int H, int Lq) {
  // 8 threads per (b, q, h) row of 64 elements
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rowi = gid >> 3;
  const int part = (int)(gid & 7);
  const long long nrows = (long long)B * Lq * H;
  float acc = 0.f;
  if (rowi < nrows) {
    float a[8], o[8];
    unpack<T>(ld_global_nc_v4(dout + rowi * 64 + part * 8), a);
    unpack<T>(ld_global_nc_v4(out + rowi * 64 + part * 8), o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += a[e] * o[e];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (rowi < nrows && part == 0) {
    const long long hq = rowi % ((long long)Lq * H);
    const long long b = rowi / ((long long)Lq * H);
    const int q = (int)(hq / H), h = (int)(hq % H);
    delta[(b * H + h) * Lq + q] = acc;
  }
}

// ---- post-pass: fp32 accumulator -> 16-bit ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fmha_cast_kernel(const float* __restrict__ in, T* __restrict__ out, long long nvec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float x[8];
    unpack<float>(ld_global_nc_v4(in + v * 8), x);
    unpack<float>(ld_global_nc_v4(in + v * 8 + 4), x + 4);
    st_global_v4(out + v * 8, pack<T>(x));
  }
}

// ---- main kernel ----------------------------------------------------------------------------------------------
template <typename T, bool kBiasF32>
__global__ void __launch_bounds__(kBwdThreads, 1) fmha_bwd_kernel(FmhaBwdParams bp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const FmhaFwdParams& p = bp.f;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;  // thread = (tile row, 64-column half)
  const int key_tile0 = blockIdx.x * kBN, h = blockIdx.y, b = blockIdx.z;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_a = smem_base + kOffBar, bar_b = smem_base + kOffBar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 16);

  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), kBwdTmemCols);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_b, 1);
    fence_mbarrier_init();
  }
  const T* qg = reinterpret_cast<const T*>(p.q) + (long long)b * p.q_sb + (long long)h * p.q_sh;
  const T* kg = reinterpret_cast<const T*>(p.k) + (long long)b * p.k_sb + (long long)h * p.k_sh;
  const T* vg = reinterpret_cast<const T*>(p.v) + (long long)b * p.v_sb + (long long)h * p.v_sh;
  const long long o_sl = (long long)p.H * kD;  // contiguous [B, L, H, 64] tensors
  const T* dog = reinterpret_cast<const T*>(bp.dout) + ((long long)b * p.Lq * p.H + h) * kD;
  const int k_valid = min(kBN, p.Lk - key_tile0);
  bwd_load_tile64<T>(smem + kOffK, kg + (long long)key_tile0 * p.k_sl, p.k_sl, k_valid);
  bwd_load_tile64<T>(smem + kOffV, vg + (long long)key_tile0 * p.v_sl, p.v_sl, k_valid);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t idesc_s = make_idesc_f16(kBM, kBN, kFmt, 0, 0);    // S, dP: both operands K-major
  constexpr uint32_t idesc_t = make_idesc_f16(kBN, kD, kFmt, 1, 1);     // dV, dK: both operands MN-major
  constexpr uint32_t idesc_q = make_idesc_f16(kBM, kD, kFmt, 0, 1);     // dQ: A K-major, B MN-major

  const uint8_t* kpm_row = p.kpm != nullptr ? p.kpm + (long long)b * p.Lk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint32_t thresh = dropout_thresh16(p.p_drop);
  const float keep_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  constexpr float kLog2e = 1.4426950408889634f;
  const int bb = p.bias_batch > 1 ? b : 0;
  uint32_t phase_a = 0, phase_b = 0;
  const int n_qtiles = (p.Lq + kBM - 1) / kBM;

  for (int i = 0; i < n_qtiles; ++i) {
    const int q0 = i * kBM;
    const int q_valid = min(kBM, p.Lq - q0);
    // Q_i, dO_i -> shared memory (previous iteration's MMAs were drained via bar_b below)
    bwd_load_tile64<T>(smem + kOffQ, qg + (long long)q0 * p.q_sl, p.q_sl, q_valid);
    bwd_load_tile64<T>(smem + kOffDO, dog + (long long)q0 * o_sl, o_sl, q_valid);
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_thread_sync();
#pragma unroll
      for (int kk = 0; kk < kD / 16; ++kk) {
        const uint64_t dq_ = make_smem_desc(smem_base + kOffQ + kk * 256, 128, 1024);
        const uint64_t dk_ = make_smem_desc(smem_base + kOffK + kk * 256, 128, 1024);
        umma_f16_ss(tmem_base + kColS, dq_, dk_, idesc_s, kk > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < kD / 16; ++kk) {
        const uint64_t ddo = make_smem_desc(smem_base + kOffDO + kk * 256, 128, 1024);
        const uint64_t dv_ = make_smem_desc(smem_base + kOffV + kk * 256, 128, 1024);
        umma_f16_ss(tmem_base + kColDP, ddo, dv_, idesc_s, kk > 0 ? 1u : 0u);
      }
      umma_commit(bar_a);
    }
    const int row = q0 + r;
    const bool row_valid = row < p.Lq;
    const long long stat_idx = ((long long)b * p.H + h) * p.Lq + (row_valid ? row : 0);
    const float lse = row_valid ? p.lse[stat_idx] : CUDART_INF_F;
    const float delta = row_valid ? bp.delta[stat_idx] : 0.f;
    const float lse_use = (lse == -CUDART_INF_F) ? CUDART_INF_F : lse;  // fully masked row -> p = 0
    const void* bias_row = nullptr;
    if (p.bias != nullptr && row_valid) {
      const long long boff = (((long long)bb * p.H + h) * p.Lq + row) * p.Lk;
      bias_row = kBiasF32 ? (const void*)(reinterpret_cast<const float*>(p.bias) + boff)
                          : (const void*)(reinterpret_cast<const T*>(p.bias) + boff);
    }
    float* dbias_row = (bp.dbias != nullptr && row_valid)
                           ? bp.dbias + (((long long)bb * p.H + h) * p.Lq + row) * p.Lk
                           : nullptr;
    const unsigned long long drop_row_base = (((unsigned long long)b * p.H + h) * p.Lq + (row_valid ? row : 0)) * p.Lk;

    mbar_wait(bar_a, phase_a);
    phase_a ^= 1;
    fence_after_thread_sync();

#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = half * 64 + c * 32;      // column inside the key tile
      const int key0 = key_tile0 + col0;
      uint32_t acc[32], dpr[32];
      float s[32];
      tmem_ld32(lane_base + kColS + col0, acc);
      tmem_ld32(lane_base + kColDP + col0, dpr);
      tmem_wait_ld();
      bwd_logits32<T, kBiasF32>(acc, s, p.scale, bias_row, key0, kpm_row, p.Lk);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t keep = 0xffu;
        if (drop) {
          const unsigned long long idx = drop_row_base + (unsigned long long)(key0 + v * 8);
          keep = dropout_keep8(p.seed, p.offset, idx >> 3, thresh);
        }
        float pd[8], ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float pr = row_valid ? exp2f((s[v * 8 + e] - lse_use) * kLog2e) : 0.f;
          const float km = ((keep >> e) & 1u) ? keep_scale : 0.f;
          pd[e] = pr * km;
          ds[e] = pr * (__uint_as_float(dpr[v * 8 + e]) * km - delta);
        }
        if (dbias_row != nullptr && key0 + v * 8 < p.Lk) {
          red_add_v4(dbias_row + key0 + v * 8, ds[0], ds[1], ds[2], ds[3]);
          red_add_v4(dbias_row + key0 + v * 8 + 4, ds[4], ds[5], ds[6], ds[7]);
        }
        Vec16 op, od;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          op.w[e] = bwd_pack2<T>(pd[2 * e], pd[2 * e + 1]);
          od.w[e] = bwd_pack2<T>(ds[2 * e] * p.scale, ds[2 * e + 1] * p.scale);
        }
        const uint32_t off = tile128_off(r, (col0 >> 3) + v);
        *reinterpret_cast<Vec16*>(smem + kOffP + off) = op;
        *reinterpret_cast<Vec16*>(smem + kOffDS + off) = od;
      }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_thread_sync();
#pragma unroll
      for (int kk = 0; kk < kBM / 16; ++kk) {  // reduction over the 128 query rows, 16 per step
        const uint64_t a_p = make_smem_desc(smem_base + kOffP + kk * 4096, 2048, 128);    // P^T  (MN-major)
        const uint64_t b_do = make_smem_desc(smem_base + kOffDO + kk * 2048, 1024, 128);  // dO   (MN-major)
        umma_f16_ss(tmem_base + kColDV, a_p, b_do, idesc_t, (i > 0 || kk > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < kBM / 16; ++kk) {
        const uint64_t a_ds = make_smem_desc(smem_base + kOffDS + kk * 4096, 2048, 128);  // dS^T (MN-major)
        const uint64_t b_q = make_smem_desc(smem_base + kOffQ + kk * 2048, 1024, 128);    // Q    (MN-major)
        umma_f16_ss(tmem_base + kColDK, a_ds, b_q, idesc_t, (i > 0 || kk > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < kBN / 16; ++kk) {  // reduction over the 128 keys
        const uint64_t a_ds = make_smem_desc(smem_base + kOffDS + kk * 256, 128, 2048);   // dS   (K-major)
        const uint64_t b_k = make_smem_desc(smem_base + kOffK + kk * 2048, 1024, 128);    // K    (MN-major)
        umma_f16_ss(tmem_base + kColDQ, a_ds, b_k, idesc_q, kk > 0 ? 1u : 0u);
      }
      umma_commit(bar_b);
    }
    mbar_wait(bar_b, phase_b);
    phase_b ^= 1;
    fence_after_thread_sync();
    {  // dQ_i partial -> global fp32 accumulator
      uint32_t acc[32];
      tmem_ld32(lane_base + kColDQ + half * 32, acc);
      tmem_wait_ld();
      if (row_valid) {
        float* dst = bp.dq_acc + (((long long)b * p.Lq + row) * p.H + h) * kD + half * 32;
#pragma unroll
        for (int v = 0; v < 8; ++v)
          red_add_v4(dst + v * 4, __uint_as_float(acc[v * 4]), __uint_as_float(acc[v * 4 + 1]),
                     __uint_as_float(acc[v * 4 + 2]), __uint_as_float(acc[v * 4 + 3]));
      }
    }
    fence_before_thread_sync();  // TMEM reads done before the next iteration's MMAs overwrite S/dP/dQ
  }

  // ---- epilogue: dK_j, dV_j -------------------------------------------------------------------------------------
  {
    const int key = key_tile0 + r;
    uint32_t accv[32], acck[32];
    tmem_ld32(lane_base + kColDV + half * 32, accv);
    tmem_ld32(lane_base + kColDK + half * 32, acck);
    tmem_wait_ld();
    if (key < p.Lk) {
      T* dvg = reinterpret_cast<T*>(bp.dv) + (((long long)b * p.Lk + key) * p.H + h) * kD + half * 32;
      T* dkg = reinterpret_cast<T*>(bp.dk) + (((long long)b * p.Lk + key) * p.H + h) * kD + half * 32;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        Vec16 ov, ok;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ov.w[e] = bwd_pack2<T>(__uint_as_float(accv[v * 8 + 2 * e]), __uint_as_float(accv[v * 8 + 2 * e + 1]));
          ok.w[e] = bwd_pack2<T>(__uint_as_float(acck[v * 8 + 2 * e]), __uint_as_float(acck[v * 8 + 2 * e + 1]));
        }
        st_global_v4(dvg + v * 8, ov);
        st_global_v4(dkg + v * 8, ok);
      }
    }
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kBwdTmemCols);
}

template <typename T>
void run_bwd(const FmhaBwdParams& bp, cudaStream_t stream) {
  const FmhaFwdParams& p = bp.f;
  const long long nrows = (long long)p.B * p.Lq * p.H;
  fmha_delta_kernel<T><<<(unsigned)((nrows * 8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const T*>(bp.dout), reinterpret_cast<const T*>(p.out), bp.delta, p.B, p.H, p.Lq);
  dim3 grid((p.Lk + kBN - 1) / kBN, p.H, p.B);
  if (p.bias_is_f32) {
    auto kern = fmha_bwd_kernel<T, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes);
    kern<<<grid, kBwdThreads, kBwdSmemBytes, stream>>>(bp);
  } else {
    auto kern = fmha_bwd_kernel<T, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes);
    kern<<<grid, kBwdThreads, kBwdSmemBytes, stream>>>(bp);
  }
  const long long nvec = nrows * 64 / 8;
  long long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fmha_cast_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(bp.dq_acc, reinterpret_cast<T*>(bp.dq), nvec);
}

}  // namespace

void launch_fmha_bwd(const FmhaBwdParams& bp, cudaStream_t stream) {
  if (bp.f.is_bf16) run_bwd<__nv_bfloat16>(bp, stream);
  else run_bwd<__half>(bp, stream);
}

}  // namespace ub
