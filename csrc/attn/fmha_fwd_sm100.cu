// Fused multi-head attention forward for sm_100a (head_dim 64, fp16/bf16):
//   O = dropout(softmax(scale * Q K^T + bias + key_padding)) V        LSE = logsumexp of the logits
// One kernel for what the reference runs as bmm -> masked_fill -> softmax_dropout -> bmm plus four transposes
// (unicore/modules/multihead_attention.py:47-113, csrc/softmax_dropout/softmax_fast.h:207-434).
//
// Tensor-core path: tcgen05.mma (cta_group::1, M=128) with fp32 accumulators in tensor memory.
//   S (128 x 128 fp32) lives in TMEM columns [0,128), O (128 x 64 fp32) in columns [128,192).
//   One CTA (256 threads) = one 128-row query tile of one (batch, head); it walks the key tiles:
//     1. TMA (128-byte-swizzled boxes, csrc/attn/tma_map.h: one request per 128-byte row): V_j at the top of the
//        tile; K_{j+1} as soon as S_j is complete; the bias tile of tile j+1 as soon as tile j's has been consumed
//        (own 32 KB buffer) - every copy is one instruction from one thread, completion on an mbarrier
//     2. S  = Q K_j^T          4 x tcgen05.mma (K=16 each), commit -> mbarrier
//     3. softmax in ONE pass over registers: thread = (query row, 64-key half); TMEM lane = row.
//        tcgen05.ld, scale/bias FMA and exp2 argument on packed fp32x2 (FFMA2), row max exchanged
//        between the two halves through shared memory, Philox dropout decided two keys per HSET2
//        (the 0xffff/0 masks AND the packed P; keep bits are stored for backward), P -> tensor memory
//        columns [192,256) with tcgen05.st (two 16-bit probabilities per column), O rescaled in TMEM only when the
//        row maximum has moved by more than the slack (lazy rescaling)
//     4. O += P V_j            8 x tcgen05.mma with the A operand (P) read from TENSOR MEMORY and V presented MN-major
//        from the same row-major bytes; V_j is only waited for here, so its latency hides behind the softmax.
//        (P through a shared-memory tile - UNICORE_B200_FMHA_P=smem - costs a 32 KB store and a 32 KB operand read per
//        tile: 0.120 instead of 0.118 ms with the tile 128-byte-swizzled, 0.137 ms with unswizzled core matrices.)
//   q/k/v are read through strides straight out of the packed in_proj output; O is written as
//   [B, Lq, H, 64] so that out_proj consumes it without a transpose.
// Two CTAs are resident per SM (113 KB smem, 256 TMEM columns, <= 128 registers/thread) so one CTA's
// softmax overlaps the other's copies and MMAs; 16 warps per SM hide the ALU/MUFU latencies.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdlib>
#include <type_traits>

#include "../common.cuh"
#include "fmha_api.h"
#include "tcgen05.cuh"

namespace ub {
namespace {

using namespace tc;

constexpr int kBlockM = 128;   // query rows per CTA
constexpr int kBlockN = 128;   // keys per tile
constexpr int kHeadDim = 64;
constexpr int kFwdThreads = 256;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemColS = 0, kTmemColO = 128, kTmemColP = 192;   // P (16-bit, 64 columns) when it feeds PV from TMEM

constexpr uint32_t kSmemQ = 0;
constexpr uint32_t kSmemK = 16384;
constexpr uint32_t kSmemV = 32768;
constexpr uint32_t kSmemP = 49152;            // 128 x 128 x 2 = 32768 bytes: P
constexpr uint32_t kSmemBias = 81920;         // 32768 bytes: bias tile of the NEXT key tile (TMA, one tile ahead)
constexpr uint32_t kSmemXchg = kSmemBias;     // float[256] row max / row sum exchange: aliases the bias tile, which
                                              // is dead between "logits done" and the next bias copy
constexpr uint32_t kSmemKAdd = 81920 + 32768; // float[128]: 0 or -inf per key of the tile
constexpr uint32_t kSmemBar = kSmemKAdd + 512;
constexpr uint32_t kFwdSmemBytes = kSmemBar + 64;   // 115,264 B: two CTAs (+1 KB reserve each) fit in 228 KB

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

#define UB_TRACE(slot)                                                                      \
  do {                                                                                      \
    if (trace != nullptr) trace[j * 12 + (slot)] = clock64();                               \
  } while (0)

// kPTmem: the probabilities are handed to the second MMA through tensor memory (tcgen05.st + A operand from TMEM)
// instead of a 32 KB shared-memory tile: the P V product then reads only V from shared memory (16 KB instead of 48 KB of
// operand traffic per tile - tcgen05.mma sustains about half of the nominal 128 B/clk of operand reads next to the
// softmax threads' own shared-memory traffic, so the M128 N64 product is operand-bound).
template <typename T, bool kPTmem>
__global__ void __launch_bounds__(kFwdThreads, 2) fmha_fwd_kernel(const __grid_constant__ FmhaFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;   // thread = (tile row, 64-key half)
  const int q0 = blockIdx.x * kBlockM, h = blockIdx.y, b = blockIdx.z;
  long long* trace = (p.trace != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 1 && blockIdx.z == 1) ? p.trace : nullptr;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_s = smem_base + kSmemBar, bar_o = smem_base + kSmemBar + 8;
  // TMA completion barriers: K tiles (the first phase also covers Q), V tiles, bias tiles
  const uint32_t bar_k = smem_base + kSmemBar + 24, bar_v = smem_base + kSmemBar + 32, bar_b = smem_base + kSmemBar + 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemBar + 16);
  float* kadd = reinterpret_cast<float*>(smem + kSmemKAdd);
  // exchange slots of the two threads that share a query row: inside the bias tile, each in the first chunk
  // that only its owner ever reads (so no other thread's bias loads can race with the write)
  float* xchg_mine = reinterpret_cast<float*>(smem + kSmemXchg + half * 16384 + sw128_off(r, 0));
  float* xchg_peer = reinterpret_cast<float*>(smem + kSmemXchg + (half ^ 1) * 16384 + sw128_off(r, 0));

  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_b, 1);
    fence_mbarrier_init();
  }
  const bool has_bias = p.bias != nullptr;
  // All CTAs walk the key tiles in the same order on purpose: CTAs of different batch entries then hit the
  // same bias lines in L2 at about the same time (a rotated order measured 4 % slower).
  const int n_tiles = (p.Lk + kBlockN - 1) / kBlockN;
  constexpr int rot = 0;
  constexpr uint32_t kTileBytes = kBlockM * kHeadDim * 2, kBiasBytes = kBlockM * kBlockN * 2;
  const int bias_nb = (p.bias_batch > 1 ? b : 0) * p.H + h;
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  if (tid == 0) {  // Q tile and K_0: two TMA boxes, one barrier phase
    mbar_expect_tx(bar_k, 2 * kTileBytes);
    tma_load_4d(smem_base + kSmemQ, &p.sw_q, 0, h, q0, b, bar_k);
    tma_load_4d(smem_base + kSmemK, &p.sw_k, 0, h, rot * kBlockN, b, bar_k);
    if (p.bias != nullptr) {   // a 128 x 128 bias tile = two 64-column boxes (one per thread half)
      mbar_expect_tx(bar_b, kBiasBytes);
      tma_load_3d(smem_base + kSmemBias, &p.sw_bias, rot * kBlockN, q0, bias_nb, bar_b);
      tma_load_3d(smem_base + kSmemBias + 16384, &p.sw_bias, rot * kBlockN + 64, q0, bias_nb, bar_b);
    }
  }
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's TMEM lanes
  constexpr int kFmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t idesc_qk = make_idesc_f16(kBlockM, kBlockN, kFmt, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_f16(kBlockM, kHeadDim, kFmt, 0, 1);

  const int row = q0 + r;
  const bool row_valid = row < p.Lq;
  const uint8_t* kpm_row = p.kpm != nullptr ? p.kpm + (long long)b * p.Lk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint32_t t14 = dropout_thresh14(p.p_drop);
  const uint32_t t14x2 = t14 | (t14 << 16);
  // the 1/(1-p) of the kept probabilities is applied once, to the normalised output row
  const float keep_scale = drop ? dropout_keep_scale14(t14) : 1.f;
  const unsigned long long row_lin = ((unsigned long long)b * p.H + h) * p.Lq + (row_valid ? row : 0);
  const unsigned long long drop_row_base = row_lin * p.Lk;
  uint32_t* bits_row = (drop && p.drop_bits != nullptr) ? p.drop_bits + row_lin * ((p.Lk + 31) / 32) : nullptr;

  constexpr float kLog2e = 1.4426950408889634f;
  float m_run = -CUDART_INF_F, l_run = 0.f;   // running max of the logits (common to both halves), partial sum
  uint32_t phase_s = 0, phase_o = 0;

  for (int j = 0; j < n_tiles; ++j) {
    const int jt = (j + rot) % n_tiles, jt_next = (j + 1 + rot) % n_tiles;
    const int key_tile0 = jt * kBlockN;
    UB_TRACE(0);
    if (j > 0) {  // previous P V must be done before V / P shared memory is overwritten
      mbar_wait(bar_o, phase_o);
      phase_o ^= 1;
      fence_after_thread_sync();
    }
    UB_TRACE(1);
    if (tid == 0) {  // V_j into the V buffer (PV_{j-1} has released it); bias_j was requested one tile ago
      mbar_expect_tx(bar_v, kTileBytes);
      tma_load_4d(smem_base + kSmemV, &p.sw_v, 0, h, key_tile0, b, bar_v);
    }
    bool masked = false;
    if (tid < kBlockN) {  // additive key mask of this tile
      const int key = key_tile0 + tid;
      masked = key >= p.Lk || (kpm_row != nullptr && kpm_row[key] != 0);
      kadd[tid] = masked ? -CUDART_INF_F : 0.f;
    }
    mbar_wait(bar_k, (uint32_t)(j & 1));   // K_j (and Q on the first tile) have landed
    const bool tile_masked = __syncthreads_or(masked) != 0;  // most tiles have no masked key: skip the adds
    UB_TRACE(2);
    if (tid == 0) {
      fence_after_thread_sync();
#pragma unroll
      for (int kk = 0; kk < kHeadDim / 16; ++kk) {
        const uint64_t da = make_smem_desc_sw128(smem_base + kSmemQ + kk * 32);
        const uint64_t db = make_smem_desc_sw128(smem_base + kSmemK + kk * 32);
        umma_f16_ss(tmem_base + kTmemColS, da, db, idesc_qk, kk > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    UB_TRACE(3);
    if (has_bias) mbar_wait(bar_b, (uint32_t)(j & 1));   // the bias tile is visible to ordinary loads now
    UB_TRACE(4);
    mbar_wait(bar_s, phase_s);
    phase_s ^= 1;
    fence_after_thread_sync();
    UB_TRACE(5);
    if (tid == 0 && j + 1 < n_tiles) {  // S is complete, so the K buffer is free: prefetch K_{j+1} under the softmax
      mbar_expect_tx(bar_k, kTileBytes);
      tma_load_4d(smem_base + kSmemK, &p.sw_k, 0, h, jt_next * kBlockN, b, bar_k);
    }
    UB_TRACE(6);

    // ---- logits of my 64 columns in registers: x = acc*scale + bias (+ key mask), packed fp32x2 math ------
    F2 x[32];
    float m_part = -CUDART_INF_F;
    const F2 scale_2 = f2(p.scale);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = half * 64 + c * 32;
      uint32_t acc[32];
      tmem_ld32(lane_base + kTmemColS + col0, acc);
      tmem_wait_ld();
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float bf[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bf[e] = 0.f;
        if (has_bias) unpack<T>(*reinterpret_cast<const Vec16*>(smem + kSmemBias + half * 16384 + sw128_off(r, c * 4 + v)), bf);
        if (tile_masked) {
          const float4 ka = *reinterpret_cast<const float4*>(kadd + col0 + v * 8);
          const float4 kb = *reinterpret_cast<const float4*>(kadd + col0 + v * 8 + 4);
          bf[0] += ka.x; bf[1] += ka.y; bf[2] += ka.z; bf[3] += ka.w;
          bf[4] += kb.x; bf[5] += kb.y; bf[6] += kb.z; bf[7] += kb.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const F2 a2 = F2{__uint_as_float(acc[v * 8 + 2 * e]), __uint_as_float(acc[v * 8 + 2 * e + 1])};
          const F2 xi = fma2(a2, scale_2, F2{bf[2 * e], bf[2 * e + 1]});
          x[c * 16 + v * 4 + e] = xi;
          m_part = fmaxf(m_part, fmaxf(xi.x, xi.y));
        }
      }
    }
    UB_TRACE(7);
    *xchg_mine = m_part;           // my own (already consumed) bias chunk doubles as exchange space
    __syncthreads();
    UB_TRACE(8);
    const float m_tile = fmaxf(m_part, *xchg_peer);
    __syncthreads();               // exchange read by everybody: the bias tile of the next key tile may land now
    if (tid == 0 && has_bias && j + 1 < n_tiles) {
      mbar_expect_tx(bar_b, kBiasBytes);
      tma_load_3d(smem_base + kSmemBias, &p.sw_bias, jt_next * kBlockN, q0, bias_nb, bar_b);
      tma_load_3d(smem_base + kSmemBias + 16384, &p.sw_bias, jt_next * kBlockN + 64, q0, bias_nb, bar_b);
    }
    // Lazy rescaling: the running reference m_run only follows the true row maximum when that has grown by more than
    // kRescaleSlack (natural-log units; exp(8) = 2981 keeps P, the row sum and the fp32 accumulator far from overflow:
    // P <= 2981 < 65504).  exp(x - m_run) with a slightly stale m_run is exact arithmetic - the final division by the
    // row sum and LSE = m_run + log(l) absorb it - and the TMEM round trip that rescales O (tcgen05.ld, 32 multiplies,
    // tcgen05.st, two waits: the longest dependent chain of a tile) disappears from almost every tile.
    constexpr float kRescaleSlack = 8.f;
    const float m_true = fmaxf(m_run, m_tile);
    const bool first = m_run == -CUDART_INF_F;
    const bool jump = first || (m_true - m_run > kRescaleSlack);
    const float m_new = jump ? m_true : m_run;
    const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
    const float alpha = jump ? exp2f((m_run - m_use) * kLog2e) : 1.f;  // m_run = -inf -> 0
    // both threads of a row (different warps) take the same decision; a warp skips the O round trip when none of its
    // rows moved its reference
    const bool warp_rescales = __any_sync(0xffffffffu, jump && j > 0);
    l_run *= alpha;
    m_run = m_new;

    // ---- probabilities, dropout, P -> shared memory (in place over the bias chunks) ---------------------
    F2 psum2 = f2(0.f);
    const F2 log2e_2 = f2(kLog2e), nm_2 = f2(-m_use * kLog2e);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int col0 = half * 64 + c * 32;
      uint32_t keep_word = 0u;   // bit (lane * 16 + i) = pair i (keys col0 + 2i, col0 + 2i + 1), lane = key parity
      uint32_t pw[16];           // kPTmem: the 32 probabilities of this chunk, packed
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
        Vec16 o;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const F2 arg = fma2(x[c * 16 + v * 4 + e], log2e_2, nm_2);
          F2 pr;
          pr.x = ex2_approx(arg.x);
          pr.y = ex2_approx(arg.y);
          psum2 = add2(psum2, pr);
          o.w[e] = pack2<T>(pr.x, pr.y) & km[e];
        }
        if (kPTmem) {
#pragma unroll
          for (int e = 0; e < 4; ++e) pw[v * 4 + e] = o.w[e];
        } else {
          *reinterpret_cast<Vec16*>(smem + kSmemP + half * 16384 + sw128_off(r, c * 4 + v)) = o;   // two 64-key swizzled boxes
        }
      }
      if (kPTmem) tmem_st16(lane_base + kTmemColP + half * 32 + c * 16, pw);   // keys col0 .. col0+31 of my row
      if (drop && bits_row != nullptr && row_valid && key_tile0 + col0 < p.Lk) bits_row[(key_tile0 + col0) >> 5] = keep_word;
    }
    l_run += psum2.x + psum2.y;
    UB_TRACE(9);

    // ---- rescale the running output accumulator (32 of the 64 columns per thread) ------------------------
    if (warp_rescales) {
      uint32_t acc[32];
      tmem_ld32(lane_base + kTmemColO + half * 32, acc);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) * alpha);
      tmem_st32(lane_base + kTmemColO + half * 32, acc);
      tmem_wait_st();
    }
    UB_TRACE(10);
    mbar_wait(bar_v, (uint32_t)(j & 1));   // V_j has landed
    if (kPTmem) tmem_wait_st();    // my P columns are in tensor memory
    else fence_proxy_async_smem(); // my P stores (generic proxy) before the tensor core (async proxy) reads them
    fence_before_thread_sync();
    __syncthreads();
    UB_TRACE(11);
    if (tid == 0) {
      fence_after_thread_sync();
#pragma unroll
      for (int kk = 0; kk < kBlockN / 16; ++kk) {
        // A = P [128 q x 128 keys] K-major in two 64-key swizzled boxes: a 16-key step is +32 B inside a box
        const uint64_t da = make_smem_desc_sw128(smem_base + kSmemP + (kk >> 2) * 16384 + (kk & 3) * 32);
        const uint32_t ta = tmem_base + kTmemColP + kk * 8;   // kPTmem: 16 keys = 8 columns of packed pairs
        // B = V [64 d x 128 keys] MN-major view of the row-major [key][d] tile: 16 keys = 2 x 1024 B (swizzled)
        const uint64_t db = make_smem_desc_sw128(smem_base + kSmemV + kk * 2048);
        if (kPTmem) umma_f16_ts(tmem_base + kTmemColO, ta, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        else umma_f16_ss(tmem_base + kTmemColO, da, db, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  }

  // ---- epilogue ---------------------------------------------------------------------------------------
  *xchg_mine = l_run;
  mbar_wait(bar_o, phase_o);
  fence_after_thread_sync();
  __syncthreads();
  const float l_tot = l_run + *xchg_peer;
  const float inv_l = l_tot > 0.f ? keep_scale / l_tot : 0.f;
  {
    uint32_t acc[32];
    tmem_ld32(lane_base + kTmemColO + half * 32, acc);
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
    }
  }
  if (row_valid && half == 0) {
    // natural-log LSE of the logits
    p.lse[row_lin] = (l_tot > 0.f) ? m_run + log2f(l_tot) * 0.6931471805599453f : -CUDART_INF_F;
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

void launch_fmha_fwd(const FmhaFwdParams& p, cudaStream_t stream) {
  // Measured (B200, B=32 H=12 L=512, bias + mask + dropout): this kernel 0.137 ms, the warp-specialised one 0.144 ms:
  // this one makes one pass over the logits instead of two (~40 % fewer instructions; Philox dropout is 60 % of them)
  // and its second CTA per SM hides prologue and epilogue.  UNICORE_B200_FMHA_FWD=ws selects the warp-specialised kernel.
  static const bool use_ws = [] {
    const char* e = getenv("UNICORE_B200_FMHA_FWD");
    return e != nullptr && e[0] == 'w' && e[1] == 's';
  }();
  if (use_ws && fmha_fwd_ws_supported(p)) {
    launch_fmha_fwd_ws(p, stream);
    return;
  }
  dim3 grid((p.Lq + kBlockM - 1) / kBlockM, p.H, p.B);
  static const bool p_in_smem = [] {
    const char* e = getenv("UNICORE_B200_FMHA_P");
    return e != nullptr && e[0] == 's';   // "smem": the earlier hand-over of P through shared memory
  }();
  auto run = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes);
    kern<<<grid, kFwdThreads, kFwdSmemBytes, stream>>>(p);
  };
  if (p.is_bf16) {
    if (p_in_smem) run(fmha_fwd_kernel<__nv_bfloat16, false>);
    else run(fmha_fwd_kernel<__nv_bfloat16, true>);
  } else {
    if (p_in_smem) run(fmha_fwd_kernel<__half, false>);
    else run(fmha_fwd_kernel<__half, true>);
  }
}

}  // namespace ub
