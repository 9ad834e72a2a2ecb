// Fused multi-head attention backward for sm_100a (head_dim 64, fp16/bf16), flash-attention style
// (reference: autograd through unicore/modules/multihead_attention.py:47-113 + softmax_fast.h:513-666):
// probabilities are recomputed from the saved log-sum-exp; the dropout keep bits come from the
// forward pass (1 bit / element).
//
//   grid = (key tiles, H, B); one CTA owns K_j, V_j (128 keys) and walks the query tiles i.  16 math warps
//   (thread = (query row, 32-column quarter): no cross-thread reductions) + 1 warp whose elected lane issues every TMA
//   copy and every tcgen05.mma.  The roles meet only through mbarriers (no CTA-wide barrier inside the loop):
//     S   = Q_i K_j^T                      tcgen05.mma  -> TMEM [  0,128)
//     dP  = dO_i V_j^T                     tcgen05.mma  -> TMEM [128,256)
//     P   = exp2((S*scale + bias + kmask - LSE_i) * log2e) ;  dS = P o (drop(dP) - delta_i)   (packed fp32x2)
//     dQ_i  = dS K_j                       tcgen05.mma  -> TMEM [384,448) / [448,512) alternating -> 16-bit partial
//     dV_j += drop(P)^T dO_i               tcgen05.mma  -> TMEM [256,320)   (A and B MN-major views)
//     dK_j += dS^T Q_i                     tcgen05.mma  -> TMEM [320,384)
//     dS tile (16-bit, already in shared memory for the MMAs) -> TMA store; a follow-up kernel sums it over the batch
//     into dBias.  (The first versions used red.global.add.v4.f32 for dQ and dBias: 6144 per tile, atomics-bound.)
//   Pipeline of one query tile (issuing warp | math warps):
//     S_{i+1}, dP_{i+1} are issued as soon as the math warps hold S_i, dP_i in registers (kBarSFree), so they - and
//     dQ_i / dV_i / dK_i, issued when P_i / dS_i are in shared memory (kBarPds) - run under the math of tiles i, i+1;
//     dQ_{i-1} is read out of its own tensor-memory buffer after the math warps have handed over tile i, off the
//     critical path.  Q / dO are fetched two tiles ahead (3 stages), the bias tile one ahead into the buffer that later
//     receives dS (same swizzled layout => in-place overwrite, chunk by chunk).
//   All TMA boxes are 128-byte-swizzled (csrc/attn/tma_map.h; the 16-byte-row boxes of the first version cost ~4000
//   cycles per tile); every operand tile is written once and presented to the tensor core as K-major or MN-major by
//   the descriptor (no transposes).  Measured (profiles/): the tile is now bound by the shared-memory port - ~400 KB
//   of operand reads + tile writes per query tile - rather than by copies or by waiting.
// Pre-pass:  delta = rowsum(dO o O).   Post-pass: dq = sum of the per-key-tile partials (and dbias = batch sum of dS).
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
constexpr int kBwdMathThreads = 512;            // 16 warps: thread = (query row, 32-column quarter)
constexpr int kBwdThreads = kBwdMathThreads + 32;  // + one warp that only issues TMA copies and tcgen05.mma
constexpr uint32_t kBwdTmemCols = 512;
constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;   // dQ: two 64-column buffers

// shared memory map (bytes)
constexpr int kQStages = 3;            // Q / dO are fetched two query tiles ahead
constexpr uint32_t kOffQ = 0;          // 3 x 16 KB (128-byte-swizzled rows)
constexpr uint32_t kOffDO = 49152;     // 3 x 16 KB
constexpr uint32_t kOffK = 98304;
constexpr uint32_t kOffV = 114688;
constexpr uint32_t kOffP = 131072;     // 32 KB, core-matrix layout (written by the math warps; as the MN-major A operand
                                       // of dV this measured 2 % faster than the two-box swizzled layout dS uses)
constexpr uint32_t kOffDS = 163840;    // 2 x 32 KB: bias tile, then dS (in place); two 64-column swizzled boxes each
constexpr uint32_t kOffKAdd = 229376;  // float[128]
constexpr uint32_t kOffBar = kOffKAdd + 512;
// mbarriers
enum : int { kBarKV = 0, kBarIn0, kBarBias0 = kBarIn0 + kQStages, kBarSdp = kBarBias0 + 2, kBarSFree, kBarPds, kBarDq, kBarDvk,
             kNumBwdBars };
constexpr uint32_t kBwdSmemBytes = kOffBar + 8 * kNumBwdBars + 16;
static_assert(kBwdSmemBytes <= 232448, "shared memory budget");

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

// 256-bit store (sm_100: STG.E.256): one full 32-byte sector per thread and instruction
UB_DEVICE void st_global_v8(void* addr, const Vec16& lo, const Vec16& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(addr), "r"(lo.w[0]), "r"(lo.w[1]), "r"(lo.w[2]),
               "r"(lo.w[3]), "r"(hi.w[0]), "r"(hi.w[1]), "r"(hi.w[2]), "r"(hi.w[3])
               : "memory");
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

// ---- post-pass: dq = sum over key tiles of the 16-bit partials (fp32 accumulate) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fmha_dq_sum_kernel(const T* __restrict__ part, T* __restrict__ out, long long nvec,
                                                            int n_parts, int H, int Lq, int n_qt, long long sb,
                                                            long long sl, long long sh) {
  // part: [n_parts] x [B][H][n_qt][4][128][16] (see the main kernel); out: 16-bit with (batch, seq, head) strides
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < n_parts; j0 += 4) {
      Vec16 in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < n_parts) in[u] = ld_global_nc_v4(part + ((long long)(j0 + u) * nvec + v) * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < n_parts) {
          float x[8];
          unpack<T>(in[u], x);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += x[e];
        }
      }
    }
    // v enumerates the partial layout [b][h][q tile][quarter][row][2 x 8 columns]
    const int c8 = (int)(v & 1);
    const long long t1 = v >> 1;
    const int rr = (int)(t1 % 128);
    const long long t2 = t1 / 128;
    const int qt = (int)(t2 & 3);
    const long long t3 = t2 >> 2;
    const int tile = (int)(t3 % n_qt);
    const long long t4 = t3 / n_qt;
    const int hh = (int)(t4 % H);
    const long long bb = t4 / H;
    const long long qq = (long long)tile * 128 + rr;
    if (qq < Lq) st_global_v4(out + bb * sb + qq * sl + (long long)hh * sh + qt * 16 + c8 * 8, pack<T>(acc));
  }
}

// ---- main kernel ----------------------------------------------------------------------------------------------
#define UB_BTRACE(slot)                                                                     \
  do {                                                                                      \
    if (trace != nullptr && i < 64) trace[i * 12 + (slot)] = clock64();                     \
  } while (0)

// whole warp: every lane has finished what the barrier protects -> one arrival
UB_DEVICE void bwd_warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

template <typename T>
__global__ void __launch_bounds__(kBwdThreads, 1) fmha_bwd_kernel(const __grid_constant__ FmhaBwdParams bp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const FmhaFwdParams& p = bp.f;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = tid & 127, quarter = (tid >> 7) & 3;  // math thread = (tile row, 32-column quarter)
  const int col0 = quarter * 32;                       // my columns inside the key tile
  const int key_tile0 = blockIdx.x * kBN, h = blockIdx.y, b = blockIdx.z;
  long long* trace = (bp.trace != nullptr && tid == 0 && blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 1) ? bp.trace : nullptr;
  const uint32_t smem_base = smem_u32(smem);
  auto bar = [&](int k) { return smem_base + kOffBar + 8u * (uint32_t)k; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * kNumBwdBars);
  float* kadd = reinterpret_cast<float*>(smem + kOffKAdd);

  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), kBwdTmemCols);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(bar(kBarKV), 1);
    for (int s = 0; s < kQStages; ++s) mbar_init(bar(kBarIn0 + s), 1);
    mbar_init(bar(kBarBias0), 1);
    mbar_init(bar(kBarBias0 + 1), 1);
    mbar_init(bar(kBarSdp), 1);
    mbar_init(bar(kBarSFree), kBwdMathThreads / 32);
    mbar_init(bar(kBarPds), kBwdMathThreads / 32);
    mbar_init(bar(kBarDq), 1);
    mbar_init(bar(kBarDvk), 1);
    fence_mbarrier_init();
  }
  const int bb = p.bias_batch > 1 ? b : 0;
  const bool has_bias = p.bias != nullptr;
  const int n_qtiles = (p.Lq + kBM - 1) / kBM;
  constexpr uint32_t kTileBytes = kBM * kD * 2, kBiasBytes = kBM * kBN * 2;
  const int bias_nb = bb * p.H + h;

  bool key_masked = false;
  if (tid < kBN) {
    const int key = key_tile0 + tid;
    key_masked = key >= p.Lk || (p.kpm != nullptr && p.kpm[(long long)b * p.Lk + key] != 0);
    kadd[tid] = key_masked ? -CUDART_INF_F : 0.f;
  }
  fence_before_thread_sync();
  const bool tile_masked = __syncthreads_or(key_masked) != 0;  // usually no key of the tile is masked
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

  if (warp == kBwdMathThreads / 32) {
    // ============================ issuing warp: every TMA copy and every tcgen05.mma ============================
    if (lane == 0) {
      constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
      constexpr uint32_t idesc_s = make_idesc_f16(kBM, kBN, kFmt, 0, 0);    // S, dP: both operands K-major
      constexpr uint32_t idesc_t = make_idesc_f16(kBN, kD, kFmt, 1, 1);     // dV, dK: both operands MN-major
      constexpr uint32_t idesc_q = make_idesc_f16(kBM, kD, kFmt, 0, 1);     // dQ: A K-major, B MN-major
      auto load_q_do = [&](int i) {   // Q_i, dO_i -> stage i % 3
        const uint32_t st = (uint32_t)(i % kQStages);
        mbar_expect_tx(bar(kBarIn0 + st), 2 * kTileBytes);
        tma_load_4d(smem_base + kOffQ + st * 16384, &p.sw_q, 0, h, i * kBM, b, bar(kBarIn0 + st));
        tma_load_4d(smem_base + kOffDO + st * 16384, &bp.sw_do, 0, h, i * kBM, b, bar(kBarIn0 + st));
      };
      auto load_bias = [&](int i) {   // bias tile (i, j) -> the buffer that later receives dS_i; two 64-column boxes
        const uint32_t buf = (uint32_t)(i & 1);
        mbar_expect_tx(bar(kBarBias0 + buf), kBiasBytes);
        tma_load_3d(smem_base + kOffDS + buf * 32768, &p.sw_bias, key_tile0, i * kBM, bias_nb, bar(kBarBias0 + buf));
        tma_load_3d(smem_base + kOffDS + buf * 32768 + 16384, &p.sw_bias, key_tile0 + 64, i * kBM, bias_nb,
                    bar(kBarBias0 + buf));
      };
      // S_i = Q_i K^T and dP_i = dO_i V^T
      auto issue_s_dp = [&](int i) {
        const uint32_t st = (uint32_t)(i % kQStages);
        const uint32_t sQ = smem_base + kOffQ + st * 16384, sDO = smem_base + kOffDO + st * 16384;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss(tmem_base + kColS, make_smem_desc_sw128(sQ + kk * 32), make_smem_desc_sw128(smem_base + kOffK + kk * 32),
                      idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss(tmem_base + kColDP, make_smem_desc_sw128(sDO + kk * 32), make_smem_desc_sw128(smem_base + kOffV + kk * 32),
                      idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(bar(kBarSdp));
      };
      mbar_expect_tx(bar(kBarKV), 2 * kTileBytes);
      tma_load_4d(smem_base + kOffK, &p.sw_k, 0, h, key_tile0, b, bar(kBarKV));
      tma_load_4d(smem_base + kOffV, &p.sw_v, 0, h, key_tile0, b, bar(kBarKV));
      load_q_do(0);
      if (has_bias) load_bias(0);
      if (n_qtiles > 1) load_q_do(1);
      mbar_wait(bar(kBarKV), 0);
      mbar_wait(bar(kBarIn0), 0);
      issue_s_dp(0);
      for (int i = 0; i < n_qtiles; ++i) {
        const uint32_t par = (uint32_t)(i & 1);
        // (a) the math warps hold S_i / dP_i in registers: the next pair may overwrite tensor memory
        if (i + 1 < n_qtiles) {
          mbar_wait(bar(kBarIn0 + (i + 1) % kQStages), (uint32_t)(((i + 1) / kQStages) & 1));
          mbar_wait(bar(kBarSFree), par);
          fence_after_thread_sync();
          issue_s_dp(i + 1);
        }
        // (b) dV / dK of tile i-1 are complete: its Q / dO stage and its dS buffer may be refilled
        if (i >= 1) mbar_wait(bar(kBarDvk), par ^ 1);
        tma_store_wait_read();
        if (i + 1 < n_qtiles && has_bias) load_bias(i + 1);
        if (i + 2 < n_qtiles) load_q_do(i + 2);
        // (c) P_i and dS_i are in shared memory
        mbar_wait(bar(kBarPds), par);
        fence_after_thread_sync();
        const uint32_t st = (uint32_t)(i % kQStages);
        const uint32_t sQ = smem_base + kOffQ + st * 16384, sDO = smem_base + kOffDO + st * 16384;
        const uint32_t sDS = smem_base + kOffDS + par * 32768;
        if (bp.ds_buf != nullptr) {  // bias gradient: the dS tile leaves through two TMA stores (64 columns each)
          tma_store_3d(&bp.sw_ds, key_tile0, i * kBM, b * p.H + h, sDS);
          tma_store_3d(&bp.sw_ds, key_tile0 + 64, i * kBM, b * p.H + h, sDS + 16384);
          tma_store_commit();
        }
#pragma unroll
        for (int kk = 0; kk < kBN / 16; ++kk) {  // dQ_i = dS K: reduction over the 128 keys
          const uint64_t a_ds = make_smem_desc_sw128(sDS + (kk >> 2) * 16384 + (kk & 3) * 32);   // dS (K-major, 2 x 64 keys)
          const uint64_t b_k = make_smem_desc_sw128(smem_base + kOffK + kk * 2048);              // K  (MN-major)
          umma_f16_ss(tmem_base + kColDQ + par * 64, a_ds, b_k, idesc_q, kk > 0 ? 1u : 0u);   // two dQ buffers
        }
        umma_commit(bar(kBarDq));
#pragma unroll
        for (int kk = 0; kk < kBM / 16; ++kk) {  // dV += P^T dO: reduction over the 128 query rows
          const uint64_t a_p = make_smem_desc(smem_base + kOffP + kk * 4096, 2048, 128);    // P^T  (MN-major, core matrices)
          const uint64_t b_do = make_smem_desc_sw128(sDO + kk * 2048);                      // dO   (MN-major)
          umma_f16_ss(tmem_base + kColDV, a_p, b_do, idesc_t, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < kBM / 16; ++kk) {  // dK += dS^T Q
          const uint64_t a_ds = make_smem_desc_sw128(sDS + kk * 2048, 16384);               // dS^T (MN-major, M = 2 x 64 keys)
          const uint64_t b_q = make_smem_desc_sw128(sQ + kk * 2048);                        // Q    (MN-major)
          umma_f16_ss(tmem_base + kColDK, a_ds, b_q, idesc_t, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(bar(kBarDvk));
      }
      tma_store_wait_read();  // shared memory must outlive the last dS store's reads
    }
  } else {
    // ============================ math warps ==================================================================
    const bool drop = p.p_drop > 0.f && p.drop_bits != nullptr;
    // same 14-bit threshold arithmetic as the forward kernel (common.cuh)
    const float keep_scale_m = p.p_drop > 0.f ? dropout_keep_scale14(dropout_thresh14(p.p_drop)) : 1.f;
    constexpr float kLog2e = 1.4426950408889634f;
    const F2 scale_2 = f2(p.scale), log2e_2 = f2(kLog2e);
    const int words_per_row = (p.Lk + 31) / 32;
    auto load_row_stats = [&](int i, float& lse, float& delta, uint32_t& keep) {
      const int row = i * kBM + r;
      const bool ok = row < p.Lq;
      const long long stat_idx = ((long long)b * p.H + h) * p.Lq + (ok ? row : 0);
      lse = ok ? p.lse[stat_idx] : CUDART_INF_F;
      delta = ok ? bp.delta[stat_idx] : 0.f;
      keep = 0xffffffffu;
      if (drop && ok && key_tile0 + col0 < p.Lk) keep = p.drop_bits[stat_idx * words_per_row + ((key_tile0 + col0) >> 5)];
    };
    // dQ_i partial of this key tile -> 16-bit partial buffer (16 of the 64 columns per thread)
    auto read_out_dq = [&](int i) {   // (the caller has waited for kBarDq phase i)
      uint32_t acc[16];
      tmem_ld16(lane_base + kColDQ + (uint32_t)(i & 1) * 64 + quarter * 16, acc);
      tmem_wait_ld();
      if (i * kBM + r < p.Lq && !(bp.debug_flags & 2)) {
        // partial layout [key tile][b][h][q tile][quarter][128 rows][16]: a warp (32 rows, one quarter) writes 1 KB
        // of consecutive bytes
        T* dst = reinterpret_cast<T*>(bp.dq_part) +
                 ((((((long long)blockIdx.x * p.B + b) * p.H + h) * n_qtiles + i) * 4 + quarter) * kBM + r) * 16;
        Vec16 o[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[v].w[e] = bwd_pack2<T>(__uint_as_float(acc[v * 8 + 2 * e]) * p.scale, __uint_as_float(acc[v * 8 + 2 * e + 1]) * p.scale);
        }
        st_global_v8(dst, o[0], o[1]);
      }
    };
    float nx_lse, nx_delta;
    uint32_t nx_keep;
    load_row_stats(0, nx_lse, nx_delta, nx_keep);

    for (int i = 0; i < n_qtiles; ++i) {
      const uint32_t par = (uint32_t)(i & 1);
      const uint32_t offDS = kOffDS + par * 32768;
      const bool row_valid = i * kBM + r < p.Lq;
      UB_BTRACE(0);
      if (has_bias) mbar_wait(bar(kBarBias0 + par), (uint32_t)((i >> 1) & 1));  // bias tile i is visible to ordinary loads
      UB_BTRACE(1);
      // per-row statistics and keep bits of this tile were requested during the previous tile's math
      const float lse = nx_lse, delta = nx_delta;
      const uint32_t keep_word = nx_keep;
      if (i + 1 < n_qtiles) load_row_stats(i + 1, nx_lse, nx_delta, nx_keep);
      // fully masked row (lse = -inf) or padding row -> p = 0
      const float lse2 = (lse == -CUDART_INF_F || !row_valid) ? CUDART_INF_F : lse * kLog2e;
      const F2 nlse_2 = f2(-lse2), ndelta_2 = f2(-delta);
      mbar_wait(bar(kBarSdp), par);
      fence_after_thread_sync();
      UB_BTRACE(2);
      Vec16 op[4], od[4];
      {
        uint32_t acc[32], dpr[32];
        tmem_ld32(lane_base + kColS + col0, acc);
        tmem_ld32(lane_base + kColDP + col0, dpr);
        tmem_wait_ld();
        fence_before_thread_sync();
        bwd_warp_arrive(bar(kBarSFree), lane);   // S_{i+1} / dP_{i+1} may be issued: they run under this tile's math
        UB_BTRACE(3);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t off_ds = (uint32_t)(quarter >> 1) * 16384u + sw128_off(r, (quarter & 1) * 4 + v);  // bias / dS
          float bf[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) bf[e] = 0.f;
          if (has_bias) unpack<T>(*reinterpret_cast<const Vec16*>(smem + offDS + off_ds), bf);
          if (tile_masked) {
            const float4 ka = *reinterpret_cast<const float4*>(kadd + col0 + v * 8);
            const float4 kb = *reinterpret_cast<const float4*>(kadd + col0 + v * 8 + 4);
            bf[0] += ka.x; bf[1] += ka.y; bf[2] += ka.z; bf[3] += ka.w;
            bf[4] += kb.x; bf[5] += kb.y; bf[6] += kb.z; bf[7] += kb.w;
          }
          float pd[8], ds[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // keep bits of the pair (keys 2i, 2i+1 of my 32): bit i and bit 16 + i (see fmha_fwd)
            const int pi = v * 4 + e;
            const bool k0 = (keep_word >> pi) & 1u, k1 = (keep_word >> (16 + pi)) & 1u;
            const F2 a2 = F2{__uint_as_float(acc[v * 8 + 2 * e]), __uint_as_float(acc[v * 8 + 2 * e + 1])};
            const F2 x2 = fma2(a2, scale_2, F2{bf[2 * e], bf[2 * e + 1]});
            const F2 arg = fma2(x2, log2e_2, nlse_2);
            F2 pr;
            pr.x = ex2_approx(arg.x);
            pr.y = ex2_approx(arg.y);
            const F2 km = F2{k0 ? keep_scale_m : 0.f, k1 ? keep_scale_m : 0.f};
            const F2 dp2 = F2{__uint_as_float(dpr[v * 8 + 2 * e]), __uint_as_float(dpr[v * 8 + 2 * e + 1])};
            const F2 d2 = mul2(pr, fma2(dp2, km, ndelta_2));
            pd[2 * e] = k0 ? pr.x : 0.f;       // keep_scale is applied to dV in the epilogue
            pd[2 * e + 1] = k1 ? pr.y : 0.f;
            ds[2 * e] = d2.x;                   // the softmax scale is applied to dQ / dK at read-out
            ds[2 * e + 1] = d2.y;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            op[v].w[e] = bwd_pack2<T>(pd[2 * e], pd[2 * e + 1]);
            od[v].w[e] = bwd_pack2<T>(ds[2 * e], ds[2 * e + 1]);
          }
        }
      }
      UB_BTRACE(4);
      if (i > 0) {   // dV / dK of tile i-1 have consumed P (and the dS buffer of tile i-2 was released even earlier)
        mbar_wait(bar(kBarDvk), par ^ 1);
        fence_after_thread_sync();
      }
      UB_BTRACE(5);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        *reinterpret_cast<Vec16*>(smem + kOffP + tile128_off(r, (col0 >> 3) + v)) = op[v];   // P: core-matrix layout
        const uint32_t off_ds = (uint32_t)(quarter >> 1) * 16384u + sw128_off(r, (quarter & 1) * 4 + v);
        *reinterpret_cast<Vec16*>(smem + offDS + off_ds) = od[v];   // in place over the consumed bias chunk
      }
      // dQ_{i-1} finished long ago, but its barrier phase must be OBSERVED before this thread lets dQ_i be issued (the
      // arrival below): a parity wait that falls two phases behind would wait for ever.  (compute-sanitizer's slow-down
      // exposed exactly that when this wait sat after the arrival.)
      if (i > 0) {
        mbar_wait(bar(kBarDq), par ^ 1);
        fence_after_thread_sync();
      }
      fence_proxy_async_smem();   // my P / dS stores (generic proxy) before the tensor core (async proxy) reads them
      fence_before_thread_sync();
      bwd_warp_arrive(bar(kBarPds), lane);
      UB_BTRACE(6);
      // The previous tile's dQ (in the other tensor-memory buffer) has long been finished: its read-out runs here, while
      // the tensor core works on this tile's dQ / dV / dK - off the critical path.
      if (i > 0) read_out_dq(i - 1);
      UB_BTRACE(7);
    }
    mbar_wait(bar(kBarDq), (uint32_t)((n_qtiles - 1) & 1));
    fence_after_thread_sync();
    read_out_dq(n_qtiles - 1);

    // ---- epilogue: dK_j, dV_j (16 of the 64 columns per thread) ------------------------------------------------
    mbar_wait(bar(kBarDvk), (uint32_t)((n_qtiles - 1) & 1));   // the last dV / dK accumulation has completed
    fence_after_thread_sync();
    const int key = key_tile0 + r;
    uint32_t accv[16], acck[16];
    tmem_ld16(lane_base + kColDV + quarter * 16, accv);
    tmem_ld16(lane_base + kColDK + quarter * 16, acck);
    tmem_wait_ld();
    if (key < p.Lk) {
      T* dvg = reinterpret_cast<T*>(bp.dv) + (long long)b * bp.dv_sb + (long long)key * bp.dv_sl + (long long)h * bp.dv_sh +
               quarter * 16;
      T* dkg = reinterpret_cast<T*>(bp.dk) + (long long)b * bp.dk_sb + (long long)key * bp.dk_sl + (long long)h * bp.dk_sh +
               quarter * 16;
      Vec16 ov[2], ok[2];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ov[v].w[e] = bwd_pack2<T>(__uint_as_float(accv[v * 8 + 2 * e]) * keep_scale_m,
                                    __uint_as_float(accv[v * 8 + 2 * e + 1]) * keep_scale_m);
          ok[v].w[e] = bwd_pack2<T>(__uint_as_float(acck[v * 8 + 2 * e]) * p.scale,
                                    __uint_as_float(acck[v * 8 + 2 * e + 1]) * p.scale);
        }
      }
      st_global_v8(dvg, ov[0], ov[1]);   // 32-byte aligned: tensor bases are, and head / quarter offsets are multiples of 32 B
      st_global_v8(dkg, ok[0], ok[1]);
    }
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kBwdTmemCols);
}

// dbias[n] = sum_b ds[b, n] (fp32 accumulate), n over H * Lq * Lk in 16-byte vectors
template <typename T>
__global__ void __launch_bounds__(256) fmha_dbias_reduce_kernel(const T* __restrict__ ds, T* __restrict__ dbias, int B,
                                                                  long long nvec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b0 = 0; b0 < B; b0 += 4) {
      Vec16 in[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (b0 + u < B) in[u] = ld_global_nc_v4(ds + ((long long)(b0 + u) * nvec + v) * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b0 + u < B) {
          float x[8];
          unpack<T>(in[u], x);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += x[e];
        }
      }
    }
    st_global_v4(dbias + v * 8, pack<T>(acc));
  }
}

template <typename T>
void run_bwd(const FmhaBwdParams& bp, cudaStream_t stream) {
  const FmhaFwdParams& p = bp.f;
  const long long nrows = (long long)p.B * p.Lq * p.H;
  fmha_delta_kernel<T><<<(unsigned)((nrows * 8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const T*>(bp.dout), reinterpret_cast<const T*>(p.out), bp.delta, p.B, p.H, p.Lq);
  dim3 grid((p.Lk + kBN - 1) / kBN, p.H, p.B);
  auto kern = fmha_bwd_kernel<T>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmemBytes);
  kern<<<grid, kBwdThreads, kBwdSmemBytes, stream>>>(bp);
  if (bp.ds_buf != nullptr && bp.dbias != bp.ds_buf) {
    const long long nv = (long long)p.H * p.Lq * p.Lk / 8;
    long long rb = (nv + 255) / 256;
    if (rb > 148 * 16) rb = 148 * 16;
    fmha_dbias_reduce_kernel<T><<<(unsigned)rb, 256, 0, stream>>>(reinterpret_cast<const T*>(bp.ds_buf),
                                                                 reinterpret_cast<T*>(bp.dbias), p.B, nv);
  }
  const int n_qt = (p.Lq + kBM - 1) / kBM;
  const long long nvec = (long long)p.B * p.H * n_qt * kBM * 64 / 8;   // partial layout is padded to whole q tiles
  long long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fmha_dq_sum_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const T*>(bp.dq_part),
                                                              reinterpret_cast<T*>(bp.dq), nvec, (p.Lk + kBN - 1) / kBN, p.H,
                                                              p.Lq, n_qt, bp.dq_sb, bp.dq_sl, bp.dq_sh);
}

}  // namespace

void launch_fmha_bwd(const FmhaBwdParams& bp, cudaStream_t stream) {
  if (bp.f.is_bf16) run_bwd<__nv_bfloat16>(bp, stream);
  else run_bwd<__half>(bp, stream);
}

}  // namespace ub
