// Fused element-wise / row-reduction kernels around the transformer GEMMs (sm_100a):
//   bias + exact GELU (fwd/bwd), and softmax cross-entropy over the vocabulary (fwd/bwd) that never
//   materialises fp32 log-probabilities.  All are pure HBM streams: 128-bit (or the widest legal)
//   accesses, fp32 math, grid sized to fill 148 SMs x 8 CTAs.
#include <math_constants.h>

#include "../api.h"
#include "../common.cuh"

namespace ub {

static int sm_count3() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// ================================================================================================
// bias + GELU
// ================================================================================================
UB_DEVICE float gelu_fwd(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
UB_DEVICE float gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

template <typename T, bool kBwd>
__global__ void __launch_bounds__(256) bias_gelu_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                          const T* __restrict__ bias, T* __restrict__ out,
                                                          long long nvec, int cols_vec) {
  constexpr int EPV = 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float xs[EPV], bs[EPV], o[EPV];
    unpack<T>(ld_global_nc_v4(x + v * EPV), xs);
    if (bias != nullptr) {
      unpack<T>(ld_global_v4(bias + (v % cols_vec) * EPV), bs);
#pragma unroll
      for (int e = 0; e < EPV; ++e) xs[e] += bs[e];
    }
    if (kBwd) {
      float g[EPV];
      unpack<T>(ld_global_nc_v4(dy + v * EPV), g);
#pragma unroll
      for (int e = 0; e < EPV; ++e) o[e] = g[e] * gelu_grad(xs[e]);
    } else {
#pragma unroll
      for (int e = 0; e < EPV; ++e) o[e] = gelu_fwd(xs[e]);
    }
    st_global_v4(out + v * EPV, pack<T>(o));
  }
}

template <bool kBwd>
static void run_bias_gelu(const void* dy, const void* x, const void* bias, void* out, long long rows, int cols,
                          int dtype, cudaStream_t stream) {
  const long long nvec = rows * cols / 8;
  if (nvec <= 0) return;
  const long long need = (nvec + 255) / 256;
  const long long cap = (long long)sm_count3() * 8;
  const int grid = (int)(need < cap ? need : cap);
  if (dtype == kF16)
    bias_gelu_kernel<__half, kBwd><<<grid, 256, 0, stream>>>((const __half*)dy, (const __half*)x, (const __half*)bias,
                                                             (__half*)out, nvec, cols / 8);
  else
    bias_gelu_kernel<__nv_bfloat16, kBwd><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                                    (const __nv_bfloat16*)bias, (__nv_bfloat16*)out,
                                                                    nvec, cols / 8);
}

void launch_bias_gelu_fwd(const void* x, const void* bias, void* y, long long rows, int cols, int dtype,
                          cudaStream_t stream) {
  run_bias_gelu<false>(nullptr, x, bias, y, rows, cols, dtype, stream);
}
void launch_bias_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, long long rows, int cols,
                          int dtype, cudaStream_t stream) {
  run_bias_gelu<true>(dy, x, bias, dx, rows, cols, dtype, stream);
}

// ================================================================================================
// softmax cross entropy: one CTA per row, online (max, sum) in a single pass over the logits
// ================================================================================================
constexpr int kXentThreads = 512;

UB_DEVICE void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -CUDART_INF_F) {
    s = 0.f;
  } else {
    s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  }
  m = mn;
}

template <typename T>
UB_DEVICE void row_max_sum(const T* row, int cols, float& m, float& s) {
  // per-thread online softmax statistics, widest aligned access that the row start allows
  m = -CUDART_INF_F;
  s = 0.f;
  constexpr int EPV = VecTraits<T>::kElems;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(row);
  int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (head > cols) head = cols;
  const int nvec = (cols - head) / EPV;
  const int tail_begin = head + nvec * EPV;
  for (int c = threadIdx.x; c < head; c += blockDim.x) {
    const float v = to_f32<T>(row[c]);
    const float mn = fmaxf(m, v);
    s = s * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    float xs[EPV];
    unpack<T>(ld_global_nc_v4(row + head + vi * EPV), xs);
    float lm = xs[0];
#pragma unroll
    for (int e = 1; e < EPV; ++e) lm = fmaxf(lm, xs[e]);
    const float mn = fmaxf(m, lm);
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < EPV; ++e) acc += __expf(xs[e] - mn);
    s = s * __expf(m - mn) + acc;
    m = mn;
  }
  for (int c = tail_begin + threadIdx.x; c < cols; c += blockDim.x) {
    const float v = to_f32<T>(row[c]);
    const float mn = fmaxf(m, v);
    s = s * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
}

template <typename T>
__global__ void __launch_bounds__(kXentThreads) xent_fwd_kernel(const T* __restrict__ logits,
                                                                  const long long* __restrict__ target,
                                                                  float* __restrict__ loss_rows,
                                                                  float* __restrict__ lse_out, int cols,
                                                                  long long ignore_index) {
  __shared__ float sm_m[32], sm_s[32];
  const int rowi = blockIdx.x;
  const T* row = logits + (size_t)rowi * cols;
  float m, s;
  row_max_sum<T>(row, cols, m, s);
  // warp then block merge
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) {
    sm_m[wid] = m;
    sm_s[wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
    m = lane < nw ? sm_m[lane] : -CUDART_INF_F;
    s = lane < nw ? sm_s[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      online_merge(m, s, m2, s2);
    }
    if (lane == 0) {
      const float lse = m + __logf(s);
      lse_out[rowi] = lse;
      const long long t = target[rowi];
      loss_rows[rowi] = (t == ignore_index || t < 0 || t >= cols) ? 0.f : lse - to_f32<T>(row[t]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kXentThreads) xent_bwd_kernel(const T* __restrict__ logits,
                                                                  const long long* __restrict__ target,
                                                                  const float* __restrict__ lse,
                                                                  const float* __restrict__ dloss,
                                                                  T* __restrict__ dlogits, int cols,
                                                                  long long ignore_index) {
  constexpr int EPV = VecTraits<T>::kElems;
  const int rowi = blockIdx.x;
  const T* row = logits + (size_t)rowi * cols;
  T* drow = dlogits + (size_t)rowi * cols;
  const long long t = target[rowi];
  const bool ignored = (t == ignore_index || t < 0 || t >= cols);
  const float scale = ignored ? 0.f : __ldg(dloss);
  const float l = lse[rowi];
  const uintptr_t addr = reinterpret_cast<uintptr_t>(row);
  int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (head > cols) head = cols;
  const bool same_align = ((reinterpret_cast<uintptr_t>(drow) & 15) == (addr & 15));
  const int nvec = same_align ? (cols - head) / EPV : 0;
  if (!same_align) head = 0;
  const int tail_begin = head + nvec * EPV;
  for (int c = threadIdx.x; c < head; c += blockDim.x) {
    const float pr = __expf(to_f32<T>(row[c]) - l);
    drow[c] = from_f32<T>((pr - (c == t ? 1.f : 0.f)) * scale);
  }
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const int c0 = head + vi * EPV;
    float xs[EPV], o[EPV];
    unpack<T>(ld_global_nc_v4(row + c0), xs);
#pragma unroll
    for (int e = 0; e < EPV; ++e) o[e] = (__expf(xs[e] - l) - ((c0 + e) == t ? 1.f : 0.f)) * scale;
    st_global_v4(drow + c0, pack<T>(o));
  }
  for (int c = tail_begin + threadIdx.x; c < cols; c += blockDim.x) {
    const float pr = __expf(to_f32<T>(row[c]) - l);
    drow[c] = from_f32<T>((pr - (c == t ? 1.f : 0.f)) * scale);
  }
}

void launch_softmax_xent_fwd(const void* logits, const long long* target, float* loss_rows, float* lse, int rows,
                             int cols, long long ignore_index, int dtype, cudaStream_t stream) {
  if (rows <= 0) return;
  if (dtype == kF32)
    xent_fwd_kernel<float><<<rows, kXentThreads, 0, stream>>>((const float*)logits, target, loss_rows, lse, cols,
                                                              ignore_index);
  else if (dtype == kF16)
    xent_fwd_kernel<__half><<<rows, kXentThreads, 0, stream>>>((const __half*)logits, target, loss_rows, lse, cols,
                                                               ignore_index);
  else
    xent_fwd_kernel<__nv_bfloat16><<<rows, kXentThreads, 0, stream>>>((const __nv_bfloat16*)logits, target, loss_rows,
                                                                      lse, cols, ignore_index);
}

void launch_softmax_xent_bwd(const void* logits, const long long* target, const float* lse, const float* dloss,
                             void* dlogits, int rows, int cols, long long ignore_index, int dtype,
                             cudaStream_t stream) {
  if (rows <= 0) return;
  if (dtype == kF32)
    xent_bwd_kernel<float><<<rows, kXentThreads, 0, stream>>>((const float*)logits, target, lse, dloss,
                                                              (float*)dlogits, cols, ignore_index);
  else if (dtype == kF16)
    xent_bwd_kernel<__half><<<rows, kXentThreads, 0, stream>>>((const __half*)logits, target, lse, dloss,
                                                               (__half*)dlogits, cols, ignore_index);
  else
    xent_bwd_kernel<__nv_bfloat16><<<rows, kXentThreads, 0, stream>>>((const __nv_bfloat16*)logits, target, lse,
                                                                      dloss, (__nv_bfloat16*)dlogits, cols,
                                                                      ignore_index);
}

}  // namespace ub
