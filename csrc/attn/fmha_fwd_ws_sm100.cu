// Warp-specialised fused multi-head attention forward for sm_100a (head_dim 64, fp16/bf16), same contract as
// fmha_fwd_sm100.cu:   O = dropout(softmax(scale * Q K^T + bias + key_padding)) V,  LSE = logsumexp of the logits
// (reference formulation: unicore/modules/multihead_attention.py:47-113 + csrc/softmax_dropout/softmax_fast.h:207-434).
//
// Why a second kernel: the one-role kernel (fmha_fwd_sm100.cu) runs  QK^T -> softmax -> PV  of a tile back to back
// with CTA-wide barriers in between; ncu showed 35 % issue-active / 9 % tensor pipe, i.e. everybody waits for
// everybody.  Here the roles are separate warps that only meet through mbarriers:
//
//   warp 0   (1 lane)   TMA producer: Q_A, Q_B once; per key tile K_j, the two bias tiles, V_j
//   warp 1   (1 lane)   tcgen05.mma issuer: S_g = Q_g K_j^T and O_g += P_g V_j for the two query tiles g = A, B
//   warps 2-9           softmax group A (256 threads: thread = (query row, 64-key half); TMEM lane = row)
//   warps 10-17         softmax group B
//
// One CTA per SM owns TWO 128-row query tiles of one (batch, head).  S_g(j+1) is issued as soon as group g has read
// S_g(j) out of tensor memory, so the tensor core computes the next logits (and the other group's P V) while a group is
// in its exp / dropout / pack phase; a group never waits for an MMA it has just requested.
//   TMEM: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)   (fp32, 512 columns allocated)
//   smem: Q_A Q_B K V (16 KB each)  P_A P_B bias_A bias_B (32 KB each)  exchange / key mask / barriers  = 211 KB
// Softmax of a tile is two passes over TMEM (tcgen05.ld is cheap: 32 columns per instruction): pass 1 finds the row
// maximum, pass 2 recomputes the logits, exponentiates, applies dropout and stores P - 32 live logits instead of 64
// keep the 18 warps under 112 registers.  The running maximum is lazy (see kRescaleSlack) so O is almost never
// rescaled.  Philox indexing and the keep-bit layout are exactly those of fmha_fwd_sm100.cu (the backward reads them).
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

constexpr int kBlockM = 128;   // query rows per group
constexpr int kBlockN = 128;   // keys per tile
constexpr int kHeadDim = 64;
constexpr int kGroupThreads = 256;
constexpr int kWsThreads = 64 + 2 * kGroupThreads;   // 576
constexpr int kWsMaxKeys = 4096;                     // key-mask table in shared memory
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemS = 0, kTmemO = 256;   // + g * 128 / + g * 64

constexpr uint32_t kTile = kBlockM * kHeadDim * 2;    // 16 KB
constexpr uint32_t kTile2 = kBlockM * kBlockN * 2;    // 32 KB
constexpr uint32_t kSmQ = 0;                          // + g * kTile
constexpr uint32_t kSmK = 2 * kTile;
constexpr uint32_t kSmV = 3 * kTile;
constexpr uint32_t kSmP = 4 * kTile;                  // + g * kTile2
constexpr uint32_t kSmBias = kSmP + 2 * kTile2;       // + g * kTile2
constexpr uint32_t kSmXchg = kSmBias + 2 * kTile2;    // float [group][parity][half][128]
constexpr uint32_t kSmKAdd = kSmXchg + 2 * 2 * 256 * 4;   // float [kWsMaxKeys]: 0 or -inf per key
constexpr uint32_t kSmFlag = kSmKAdd + kWsMaxKeys * 4;    // int [32]: tile has a masked key
constexpr uint32_t kSmBar = kSmFlag + 128;
constexpr uint32_t kWsSmemBytes = kSmBar + 256;

// mbarrier slots
enum : int { kBarQ = 0, kBarQB, kBarKFull, kBarKEmpty, kBarVFull, kBarVEmpty, kBarGroup0 };
enum : int { kGSFull = 0, kGSFree, kGPFull, kGPvDone, kGBiasFull, kGBiasEmpty, kGCount };
constexpr int kNumBars = kBarGroup0 + 2 * kGCount;
static_assert(kNumBars * 8 + 8 <= 256, "barrier block");

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

// whole warp: every lane has finished what the barrier protects -> one arrival
UB_DEVICE void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

#define UB_WTRACE(cond, slot)                                                               \
  do {                                                                                      \
    if (trace != nullptr && (cond) && j < 64) trace[j * 12 + (slot)] = clock64();           \
  } while (0)

template <typename T>
__global__ void __launch_bounds__(kWsThreads, 1) fmha_fwd_ws_kernel(const __grid_constant__ FmhaFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * (2 * kBlockM), h = blockIdx.y, b = blockIdx.z;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + kSmBar;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  auto gbar = [&](int g, int i) { return bars + 8u * (uint32_t)(kBarGroup0 + g * kGCount + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 8 * kNumBars);
  float* kadd = reinterpret_cast<float*>(smem + kSmKAdd);
  int* tile_flag = reinterpret_cast<int*>(smem + kSmFlag);

  // profiling only: clock64() stamps of CTA (0,1,1): slots 0-6 lane 0 of the first softmax warp, 7-11 the MMA issuer
  long long* trace = (p.trace != nullptr && lane == 0 && blockIdx.x == 0 && blockIdx.y == 1 && blockIdx.z == 1) ? p.trace : nullptr;
  const bool has_bias = p.bias != nullptr;
  const int n_tiles = (p.Lk + kBlockN - 1) / kBlockN;
  const int nq = (q0 + kBlockM < p.Lq) ? 2 : 1;   // the second query tile may not exist
  const int bias_nb = (p.bias_batch > 1 ? b : 0) * p.H + h;

  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(bar(kBarQ), 1);
    mbar_init(bar(kBarQB), 1);
    mbar_init(bar(kBarKFull), 1);
    mbar_init(bar(kBarKEmpty), 1);
    mbar_init(bar(kBarVFull), 1);
    mbar_init(bar(kBarVEmpty), 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(gbar(g, kGSFull), 1);
      mbar_init(gbar(g, kGSFree), kGroupThreads / 32);
      mbar_init(gbar(g, kGPFull), kGroupThreads / 32);
      mbar_init(gbar(g, kGPvDone), 1);
      mbar_init(gbar(g, kGBiasFull), 1);
      mbar_init(gbar(g, kGBiasEmpty), kGroupThreads / 32);
    }
    fence_mbarrier_init();
  }
  if (tid < 32) tile_flag[tid] = 0;
  __syncthreads();
  {  // additive key mask of every key tile (key padding + keys past Lk), once per CTA
    const uint8_t* kpm_row = p.kpm != nullptr ? p.kpm + (long long)b * p.Lk : nullptr;
    for (int key = tid; key < n_tiles * kBlockN; key += kWsThreads) {
      const bool masked = key >= p.Lk || (kpm_row != nullptr && kpm_row[key] != 0);
      kadd[key] = masked ? -CUDART_INF_F : 0.f;
      if (masked) tile_flag[key >> 7] = 1;
    }
  }
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      // first what the first MMA needs (Q_A, K_0), then Q_B: group A starts a third of the copy time earlier, and the
      // two groups stay out of phase (their exp phases then do not compete for the SFU)
      mbar_expect_tx(bar(kBarQ), kTile);
      tma_load_4d(smem_base + kSmQ, &p.sw_q, 0, h, q0, b, bar(kBarQ));
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t free_parity = (uint32_t)((j & 1) ^ 1);   // passes on a fresh barrier, then follows the consumer
        mbar_wait(bar(kBarKEmpty), free_parity);
        mbar_expect_tx(bar(kBarKFull), kTile);
        tma_load_4d(smem_base + kSmK, &p.sw_k, 0, h, j * kBlockN, b, bar(kBarKFull));
        if (j == 0 && nq > 1) {
          mbar_expect_tx(bar(kBarQB), kTile);
          tma_load_4d(smem_base + kSmQ + kTile, &p.sw_q, 0, h, q0 + kBlockM, b, bar(kBarQB));
        }
        if (has_bias) {
          for (int g = 0; g < nq; ++g) {
            mbar_wait(gbar(g, kGBiasEmpty), free_parity);
            mbar_expect_tx(gbar(g, kGBiasFull), kTile2);
            for (int hb = 0; hb < 2; ++hb)   // 128 x 128 tile = two 64-column boxes, one per thread half
              tma_load_3d(smem_base + kSmBias + g * kTile2 + hb * kTile, &p.sw_bias, j * kBlockN + hb * 64,
                          q0 + g * kBlockM, bias_nb, gbar(g, kGBiasFull));
          }
        }
        mbar_wait(bar(kBarVEmpty), free_parity);
        mbar_expect_tx(bar(kBarVFull), kTile);
        tma_load_4d(smem_base + kSmV, &p.sw_v, 0, h, j * kBlockN, b, bar(kBarVFull));
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(kBlockM, kBlockN, kFmt, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(kBlockM, kHeadDim, kFmt, 0, 1);
      auto issue_qk = [&](int g) {   // S_g = Q_g K^T: 4 x (K = 16)
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(smem_base + kSmQ + g * kTile + kk * 32);
          const uint64_t db = make_smem_desc_sw128(smem_base + kSmK + kk * 32);
          umma_f16_ss(tmem_base + kTmemS + g * 128, da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(gbar(g, kGSFull));
      };
      auto issue_pv = [&](int g, int j) {   // O_g += P_g V: 8 x (K = 16), V presented MN-major
#pragma unroll
        for (int kk = 0; kk < kBlockN / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(smem_base + kSmP + g * kTile2 + (kk >> 2) * kTile + (kk & 3) * 32);
          const uint64_t db = make_smem_desc_sw128(smem_base + kSmV + kk * 2048);
          umma_f16_ss(tmem_base + kTmemO + g * 64, da, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(gbar(g, kGPvDone));
      };
      mbar_wait(bar(kBarQ), 0);
      mbar_wait(bar(kBarKFull), 0);
      fence_after_thread_sync();
      issue_qk(0);
      if (nq > 1) {
        mbar_wait(bar(kBarQB), 0);
        fence_after_thread_sync();
        issue_qk(1);
      }
      umma_commit(bar(kBarKEmpty));
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t par = (uint32_t)(j & 1);
        const bool more = j + 1 < n_tiles;
        UB_WTRACE(true, 7);
        if (more) mbar_wait(bar(kBarKFull), par ^ 1);
        for (int g = 0; g < nq; ++g) {
          if (more) {   // group g has read S_g(j) out of tensor memory: the next logits may overwrite it
            mbar_wait(gbar(g, kGSFree), par);
            fence_after_thread_sync();
            issue_qk(g);
            if (g == nq - 1) umma_commit(bar(kBarKEmpty));   // K_{j+1} has been read by every S MMA: K_{j+2} may land
          }
          UB_WTRACE(true, 8 + 2 * g);
          mbar_wait(gbar(g, kGPFull), par);   // P_g(j) is in shared memory, O_g has been rescaled
          if (g == 0) mbar_wait(bar(kBarVFull), par);
          fence_after_thread_sync();
          UB_WTRACE(true, 9 + 2 * g);
          issue_pv(g, j);
        }
        umma_commit(bar(kBarVEmpty));
      }
    }
  } else {
    // ================================ softmax groups ==============================
    const int g = (warp - 2) >> 3;
    if (g < nq) {
      const int widx = (warp - 2) & 7;
      const int half = widx >> 2;
      const int r = (warp & 3) * 32 + lane;   // TMEM lanes of a warp are fixed by warp % 4
      const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const uint32_t tS = lane_base + kTmemS + g * 128 + half * 64;
      const uint32_t tO = lane_base + kTmemO + g * 64 + half * 32;
      const uint8_t* sBias = smem + kSmBias + g * kTile2 + half * kTile;   // my 64-column box
      uint8_t* sP = smem + kSmP + g * kTile2;
      float* xchg = reinterpret_cast<float*>(smem + kSmXchg) + g * 512;   // [parity][half][128]
      const int bar_id = 1 + g;

      const int row = q0 + g * kBlockM + r;
      const bool row_valid = row < p.Lq;
      const bool drop = p.p_drop > 0.f;
      const uint32_t t14 = dropout_thresh14(p.p_drop);
      const uint32_t t14x2 = t14 | (t14 << 16);
      const float keep_scale = drop ? dropout_keep_scale14(t14) : 1.f;   // applied once, to the normalised row
      const unsigned long long row_lin = ((unsigned long long)b * p.H + h) * p.Lq + (row_valid ? row : 0);
      const unsigned long long drop_row_base = row_lin * p.Lk;
      uint32_t* bits_row = (drop && p.drop_bits != nullptr) ? p.drop_bits + row_lin * ((p.Lk + 31) / 32) : nullptr;

      constexpr float kLog2e = 1.4426950408889634f;
      // Lazy rescaling (same rule as fmha_fwd_sm100.cu): the reference maximum only moves when the true one has grown
      // by more than kRescaleSlack; exp(8) = 2981 keeps P and the fp32 sums far from overflow.
      constexpr float kRescaleSlack = 8.f;
      float m_run = -CUDART_INF_F, l_run = 0.f;
      const F2 scale_2 = f2(p.scale);

      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t par = (uint32_t)(j & 1);
        const int key_tile0 = j * kBlockN;
        const bool tile_masked = tile_flag[j] != 0;
        const bool tw = warp == 2;
        UB_WTRACE(tw, 0);
        mbar_wait(gbar(g, kGSFull), par);
        fence_after_thread_sync();
        if (has_bias) mbar_wait(gbar(g, kGBiasFull), par);
        UB_WTRACE(tw, 1);

        // logits of 8 keys: x = acc * scale + bias (+ key mask), packed fp32x2
        auto logits8 = [&](const uint32_t (&acc)[32], int col0, int v, F2 (&x)[4]) {
          float bf[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) bf[e] = 0.f;
          if (has_bias) unpack<T>(*reinterpret_cast<const Vec16*>(sBias + sw128_off(r, ((col0 >> 3) & 7) + v)), bf);
          if (tile_masked) {
            const float4 ka = *reinterpret_cast<const float4*>(kadd + key_tile0 + col0 + v * 8);
            const float4 kb = *reinterpret_cast<const float4*>(kadd + key_tile0 + col0 + v * 8 + 4);
            bf[0] += ka.x; bf[1] += ka.y; bf[2] += ka.z; bf[3] += ka.w;
            bf[4] += kb.x; bf[5] += kb.y; bf[6] += kb.z; bf[7] += kb.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const F2 a2 = F2{__uint_as_float(acc[v * 8 + 2 * e]), __uint_as_float(acc[v * 8 + 2 * e + 1])};
            x[e] = fma2(a2, scale_2, F2{bf[2 * e], bf[2 * e + 1]});
          }
        };

        // ---- pass 1: row maximum of my 64 columns ---------------------------------------------------------
        float m_part = -CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t acc[32];
          tmem_ld32(tS + c * 32, acc);
          tmem_wait_ld();
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            F2 x[4];
            logits8(acc, half * 64 + c * 32, v, x);
#pragma unroll
            for (int e = 0; e < 4; ++e) m_part = fmaxf(m_part, fmaxf(x[e].x, x[e].y));
          }
        }
        UB_WTRACE(tw, 2);
        xchg[par * 256 + half * 128 + r] = m_part;
        named_bar_sync(bar_id, kGroupThreads);
        UB_WTRACE(tw, 3);
        const float m_tile = fmaxf(m_part, xchg[par * 256 + (half ^ 1) * 128 + r]);
        const float m_true = fmaxf(m_run, m_tile);
        const bool first = m_run == -CUDART_INF_F;
        const bool jump = first || (m_true - m_run > kRescaleSlack);
        const float m_new = jump ? m_true : m_run;
        const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
        const float alpha = jump ? exp2f((m_run - m_use) * kLog2e) : 1.f;   // m_run = -inf -> 0
        const bool warp_rescales = __any_sync(0xffffffffu, jump && j > 0);
        l_run *= alpha;
        m_run = m_new;

        // ---- P_g(j-1) V has been consumed: O_g may be rescaled, the P buffer rewritten -------------------
        if (j > 0) {
          mbar_wait(gbar(g, kGPvDone), par ^ 1);
          fence_after_thread_sync();
        }
        if (warp_rescales) {
          uint32_t acc[32];
          tmem_ld32(tO, acc);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) * alpha);
          tmem_st32(tO, acc);
          tmem_wait_st();
        }

        UB_WTRACE(tw, 4);
        // ---- pass 2: probabilities, dropout, P -> shared memory -------------------------------------------
        F2 psum2 = f2(0.f);
        const F2 log2e_2 = f2(kLog2e), nm_2 = f2(-m_use * kLog2e);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = half * 64 + c * 32;
          uint32_t acc[32];
          tmem_ld32(tS + c * 32, acc);
          tmem_wait_ld();
          if (c == 1) {   // S_g(j) is in registers: the tensor core may start S_g(j+1)
            fence_before_thread_sync();
            warp_arrive(gbar(g, kGSFree), lane);
          }
          uint32_t keep_word = 0u;   // same bit layout as fmha_fwd_sm100.cu
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t km[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
            if (drop) {
              const unsigned long long idx = drop_row_base + (unsigned long long)(key_tile0 + col0 + v * 8);
              const Philox4 rnd = philox4x32<7>(p.seed, p.offset, idx >> 3);
              km[0] = keep_mask2(rnd.x, t14x2);
              km[1] = keep_mask2(rnd.y, t14x2);
              km[2] = keep_mask2(rnd.z, t14x2);
              km[3] = keep_mask2(rnd.w, t14x2);
#pragma unroll
              for (int e = 0; e < 4; ++e) keep_word |= km[e] & (0x00010001u << (v * 4 + e));
            }
            F2 x[4];
            logits8(acc, col0, v, x);
            Vec16 o;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const F2 arg = fma2(x[e], log2e_2, nm_2);
              F2 pr;
              pr.x = ex2_approx(arg.x);
              pr.y = ex2_approx(arg.y);
              psum2 = add2(psum2, pr);
              o.w[e] = pack2<T>(pr.x, pr.y) & km[e];
            }
            *reinterpret_cast<Vec16*>(sP + half * kTile + sw128_off(r, c * 4 + v)) = o;   // two 64-key swizzled boxes
          }
          if (drop && bits_row != nullptr && row_valid && key_tile0 + col0 < p.Lk)
            bits_row[(key_tile0 + col0) >> 5] = keep_word;
        }
        if (has_bias) warp_arrive(gbar(g, kGBiasEmpty), lane);   // the next bias tile may land
        l_run += psum2.x + psum2.y;
        UB_WTRACE(tw, 5);
        fence_proxy_async_smem();   // my P stores (generic proxy) before the tensor core (async proxy) reads them
        fence_before_thread_sync();
        warp_arrive(gbar(g, kGPFull), lane);
        UB_WTRACE(tw, 6);
      }

      // ---- epilogue: O_g / row sum -> global, LSE ---------------------------------------------------------
      const uint32_t epar = (uint32_t)(n_tiles & 1);
      xchg[epar * 256 + half * 128 + r] = l_run;
      mbar_wait(gbar(g, kGPvDone), (uint32_t)((n_tiles - 1) & 1));
      fence_after_thread_sync();
      named_bar_sync(bar_id, kGroupThreads);
      const float l_tot = l_run + xchg[epar * 256 + (half ^ 1) * 128 + r];
      const float inv_l = l_tot > 0.f ? keep_scale / l_tot : 0.f;
      uint32_t acc[32];
      tmem_ld32(tO, acc);
      tmem_wait_ld();
      if (row_valid) {
        T* og = reinterpret_cast<T*>(p.out) + (((long long)b * p.Lq + row) * p.H + h) * kHeadDim + half * 32;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          Vec16 o;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o.w[e] = pack2<T>(__uint_as_float(acc[v * 8 + 2 * e]) * inv_l, __uint_as_float(acc[v * 8 + 2 * e + 1]) * inv_l);
          st_global_v4(og + v * 8, o);
        }
        if (half == 0) p.lse[row_lin] = (l_tot > 0.f) ? m_run + log2f(l_tot) * 0.6931471805599453f : -CUDART_INF_F;
      }
    }
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

bool fmha_fwd_ws_supported(const FmhaFwdParams& p) { return p.Lk <= kWsMaxKeys; }

void launch_fmha_fwd_ws(const FmhaFwdParams& p, cudaStream_t stream) {
  dim3 grid((p.Lq + 2 * kBlockM - 1) / (2 * kBlockM), p.H, p.B);
  if (p.is_bf16) {
    auto kern = fmha_fwd_ws_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes);
    kern<<<grid, kWsThreads, kWsSmemBytes, stream>>>(p);
  } else {
    auto kern = fmha_fwd_ws_kernel<__half>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes);
    kern<<<grid, kWsThreads, kWsSmemBytes, stream>>>(p);
  }
}

}  // namespace ub
