// LayerNorm / RMSNorm forward + single-pass backward, and the fused
// "bias + dropout + residual + LayerNorm" block epilogue, for sm_100a.
//
// Replaces reference csrc/layernorm/{layernorm.cu,layernorm_backward.cu} and csrc/rmsnorm/*:
//   * any hidden size that is a multiple of the 16-byte vector width (reference: 16 fixed sizes);
//   * a row is owned by a group of TPR threads (TPR = 1..256, power of two, picked from the hidden
//     size) that keeps the whole row in registers: VPT 16-byte vectors per thread; groups of <= 32
//     threads reduce with shuffles only;
//   * backward reads x and dy ONCE and emits dx plus per-CTA partial dgamma/dbeta (fp32), which a
//     small second kernel reduces deterministically (reference: x, dy read twice + 2 extra kernels);
//   * fp32 statistics, biased variance, rstd = rsqrt(var + eps).
#include "../api.h"
#include "../common.cuh"

namespace ub {

constexpr int kNormThreads = 256;

template <typename T>
UB_DEVICE Vec16 ldv(const T* p) { return ld_global_v4(p); }

// Reduce `v` over the TPR threads that own a row. TPR <= 32: shuffles; else via shared memory.
// `scratch` holds (kNormThreads/32) floats per reduced quantity slot.
template <int NVAL>
UB_DEVICE void group_sum(float (&v)[NVAL], int tpr, float* scratch) {
  const int lim = tpr < 32 ? tpr : 32;
#pragma unroll
  for (int k = 0; k < NVAL; ++k) {
    for (int o = lim >> 1; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (tpr > 32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpg = tpr >> 5;                 // warps per group
    const int gfirst = (warp / wpg) * wpg;    // first warp of my group
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NVAL; ++k) scratch[k * 8 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NVAL; ++k) {
      float s = 0.f;
      for (int w = 0; w < wpg; ++w) s += scratch[k * 8 + gfirst + w];
      v[k] = s;
    }
  }
}

struct NormGeom {
  int rows, cols, tpr, nvec;  // nvec = cols / elems-per-vector
};

// ------------------------------------------------------------------------------------------------
// forward.  kRMS: RMSNorm (no mean / beta).  kFused: h = residual + dropout(x + bias) first.
// ------------------------------------------------------------------------------------------------
template <typename T, int VPT, bool kRMS, bool kFused>
__global__ void __launch_bounds__(kNormThreads) norm_fwd_kernel(
    const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y,
    float* __restrict__ mean_out, float* __restrict__ rstd_out, NormGeom g, float eps,
    // fused-only arguments
    const T* __restrict__ bias, const T* __restrict__ residual, T* __restrict__ summed, float p, float keep_scale,
    unsigned long long seed, unsigned long long offset) {
  constexpr int EPV = VecTraits<T>::kElems;
  __shared__ float scratch[2 * 8];
  const int tpr = g.tpr;
  const int rows_per_cta = kNormThreads / tpr;
  const int grp = threadIdx.x / tpr, j = threadIdx.x % tpr;
  const float inv_cols = 1.f / (float)g.cols;
  const uint32_t thresh = kFused ? dropout_thresh16(p) : 0u;

  // per-thread column slice of gamma / beta / bias (fixed for all rows)
  float gam[VPT][EPV], bet[VPT][EPV], bia[VPT][EPV];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = j + k * tpr;
    if (vi < g.nvec) {
      unpack<T>(ldv(gamma + vi * EPV), gam[k]);
      if (!kRMS) unpack<T>(ldv(beta + vi * EPV), bet[k]);
      if (kFused && bias != nullptr) unpack<T>(ldv(bias + vi * EPV), bia[k]);
    }
    if (!(kFused && bias != nullptr)) {
#pragma unroll
      for (int e = 0; e < EPV; ++e) bia[k][e] = 0.f;
    }
  }

  const int nrow_iters = (g.rows + gridDim.x * rows_per_cta - 1) / (gridDim.x * rows_per_cta);
  for (int it = 0; it < nrow_iters; ++it) {
    const int row = (it * gridDim.x + blockIdx.x) * rows_per_cta + grp;
    const bool active = row < g.rows;
    float xs[VPT][EPV];
    float acc[1] = {0.f};
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int vi = j + k * tpr;
      if (active && vi < g.nvec) {
        const size_t off = (size_t)row * g.cols + (size_t)vi * EPV;
        unpack<T>(ld_global_nc_v4(x + off), xs[k]);
        if (kFused) {
          float res[EPV];
          unpack<T>(ld_global_nc_v4(residual + off), res);
          uint32_t keep = 0xffu;
          if (p > 0.f) keep = dropout_keep8(seed, offset, off / 8, thresh);
#pragma unroll
          for (int e = 0; e < EPV; ++e) {
            const float h = (xs[k][e] + bia[k][e]) * (((keep >> e) & 1u) ? keep_scale : 0.f);
            // round the sum to T first: backward and the pre-LN consumer see exactly this value
            xs[k][e] = to_f32<T>(from_f32<T>(res[e] + h));
          }
          st_global_v4(summed + off, pack<T>(xs[k]));
        }
#pragma unroll
        for (int e = 0; e < EPV; ++e) acc[0] += kRMS ? xs[k][e] * xs[k][e] : xs[k][e];
      } else {
#pragma unroll
        for (int e = 0; e < EPV; ++e) xs[k][e] = 0.f;
      }
    }
    group_sum<1>(acc, tpr, scratch);
    float mu = 0.f, rs;
    if (kRMS) {
      rs = rsqrtf(acc[0] * inv_cols + eps);
    } else {
      mu = acc[0] * inv_cols;
      float var[1] = {0.f};
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        if (j + k * tpr < g.nvec) {
#pragma unroll
          for (int e = 0; e < EPV; ++e) {
            const float d = xs[k][e] - mu;
            var[0] += d * d;
          }
        }
      }
      group_sum<1>(var, tpr, scratch);
      rs = rsqrtf(var[0] * inv_cols + eps);
    }
    if (active) {
      if (j == 0) {
        if (!kRMS) mean_out[row] = mu;
        rstd_out[row] = rs;
      }
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int vi = j + k * tpr;
        if (vi < g.nvec) {
          float o[EPV];
#pragma unroll
          for (int e = 0; e < EPV; ++e) {
            const float xh = (xs[k][e] - mu) * rs;
            o[e] = kRMS ? xh * gam[k][e] : xh * gam[k][e] + bet[k][e];
          }
          st_global_v4(y + (size_t)row * g.cols + (size_t)vi * EPV, pack<T>(o));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: dx (+ optional dropout-masked copy for the fused op) and per-CTA dgamma/dbeta partials
// ------------------------------------------------------------------------------------------------
template <typename T, int VPT, bool kRMS, bool kFused>
__global__ void __launch_bounds__(kNormThreads) norm_bwd_kernel(
    const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const T* __restrict__ gamma, T* __restrict__ dx, float* __restrict__ dgamma_part,
    float* __restrict__ dbeta_part, NormGeom g,
    // fused-only: dx_drop = keep ? dx * keep_scale : 0
    T* __restrict__ dx_drop, float p, float keep_scale, unsigned long long seed, unsigned long long offset) {
  constexpr int EPV = VecTraits<T>::kElems;
  extern __shared__ float sm_acc[];  // [2][cols] cross-group accumulators (dgamma, dbeta)
  __shared__ float scratch[2 * 8];
  const int tpr = g.tpr;
  const int rows_per_cta = kNormThreads / tpr;
  const int grp = threadIdx.x / tpr, j = threadIdx.x % tpr;
  const float inv_cols = 1.f / (float)g.cols;
  const uint32_t thresh = kFused ? dropout_thresh16(p) : 0u;

  for (int c = threadIdx.x; c < (kRMS ? 1 : 2) * g.cols; c += kNormThreads) sm_acc[c] = 0.f;

  float gam[VPT][EPV], dg[VPT][EPV], db[VPT][EPV];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = j + k * tpr;
    if (vi < g.nvec) unpack<T>(ldv(gamma + vi * EPV), gam[k]);
#pragma unroll
    for (int e = 0; e < EPV; ++e) dg[k][e] = db[k][e] = 0.f;
  }

  const int nrow_iters = (g.rows + gridDim.x * rows_per_cta - 1) / (gridDim.x * rows_per_cta);
  for (int it = 0; it < nrow_iters; ++it) {
    const int row = (it * gridDim.x + blockIdx.x) * rows_per_cta + grp;
    const bool active = row < g.rows;
    const float mu = (active && !kRMS) ? mean[row] : 0.f;
    const float rs = active ? rstd[row] : 0.f;
    float xh[VPT][EPV], gy[VPT][EPV];
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int vi = j + k * tpr;
      if (active && vi < g.nvec) {
        const size_t off = (size_t)row * g.cols + (size_t)vi * EPV;
        float xv[EPV], dv[EPV];
        unpack<T>(ld_global_nc_v4(x + off), xv);
        unpack<T>(ld_global_nc_v4(dy + off), dv);
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
          xh[k][e] = (xv[e] - mu) * rs;
          gy[k][e] = dv[e] * gam[k][e];
          s[0] += gy[k][e];
          s[1] += gy[k][e] * xh[k][e];
          dg[k][e] += dv[e] * xh[k][e];
          if (!kRMS) db[k][e] += dv[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPV; ++e) xh[k][e] = gy[k][e] = 0.f;
      }
    }
    group_sum<2>(s, tpr, scratch);
    const float m1 = kRMS ? 0.f : s[0] * inv_cols, m2 = s[1] * inv_cols;
    if (active) {
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int vi = j + k * tpr;
        if (vi < g.nvec) {
          const size_t off = (size_t)row * g.cols + (size_t)vi * EPV;
          float o[EPV];
#pragma unroll
          for (int e = 0; e < EPV; ++e) o[e] = rs * (gy[k][e] - m1 - xh[k][e] * m2);
          st_global_v4(dx + off, pack<T>(o));
          if (kFused) {
            uint32_t keep = 0xffu;
            if (p > 0.f) keep = dropout_keep8(seed, offset, off / 8, thresh);
#pragma unroll
            for (int e = 0; e < EPV; ++e) o[e] = ((keep >> e) & 1u) ? o[e] * keep_scale : 0.f;
            st_global_v4(dx_drop + off, pack<T>(o));
          }
        }
      }
    }
  }

  // combine the row groups of this CTA, then publish one fp32 partial row per CTA
  __syncthreads();
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int vi = j + k * tpr;
    if (vi < g.nvec) {
#pragma unroll
      for (int e = 0; e < EPV; ++e) {
        atomicAdd(&sm_acc[vi * EPV + e], dg[k][e]);
        if (!kRMS) atomicAdd(&sm_acc[g.cols + vi * EPV + e], db[k][e]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < g.cols; c += kNormThreads) {
    dgamma_part[(size_t)blockIdx.x * g.cols + c] = sm_acc[c];
    if (!kRMS) dbeta_part[(size_t)blockIdx.x * g.cols + c] = sm_acc[g.cols + c];
  }
}

// partial rows -> dgamma / dbeta. block = (32 columns, 8 row slices)
template <typename T>
__global__ void __launch_bounds__(256) norm_param_grad_kernel(const float* __restrict__ dg_part,
                                                                const float* __restrict__ db_part, int parts,
                                                                int cols, T* __restrict__ dgamma,
                                                                T* __restrict__ dbeta) {
  __shared__ float red[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float a = 0.f, b = 0.f;
  if (col < cols) {
    for (int r = ry; r < parts; r += 8) {
      a += dg_part[(size_t)r * cols + col];
      if (db_part != nullptr) b += db_part[(size_t)r * cols + col];
    }
  }
  red[0][ry][cx] = a;
  red[1][ry][cx] = b;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      sa += red[0][r][cx];
      sb += red[1][r][cx];
    }
    dgamma[col] = from_f32<T>(sa);
    if (db_part != nullptr) dbeta[col] = from_f32<T>(sb);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static NormGeom make_geom(int rows, int cols, int epv, int& vpt) {
  NormGeom g;
  g.rows = rows;
  g.cols = cols;
  g.nvec = cols / epv;
  if (g.nvec <= 128) {
    g.tpr = pow2ceil(g.nvec) < 32 ? pow2ceil(g.nvec) : 32;
  } else {
    g.tpr = pow2ceil((g.nvec + 3) / 4);
  }
  if (g.tpr > 256) g.tpr = 256;
  vpt = (g.nvec + g.tpr - 1) / g.tpr;
  return g;
}

static int fwd_grid(const NormGeom& g) {
  const int rows_per_cta = kNormThreads / g.tpr;
  const long long need = ((long long)g.rows + rows_per_cta - 1) / rows_per_cta;
  const long long cap = (long long)sm_count() * 8;
  return (int)(need < cap ? need : cap);
}

int norm_bwd_parts(int rows, int cols) {
  // one partial row per CTA; 2 CTAs per SM keeps the finalize pass short
  (void)cols;
  const long long cap = (long long)sm_count() * 2;
  const long long need = ((long long)rows + 7) / 8;
  long long n = need < cap ? need : cap;
  return (int)(n < 1 ? 1 : n);
}

#define UB_DISPATCH_VPT(VPT_VALUE, ...)                                   \
  switch (VPT_VALUE) {                                                    \
    case 1: { constexpr int VPT = 1; __VA_ARGS__; break; }                \
    case 2: { constexpr int VPT = 2; __VA_ARGS__; break; }                \
    case 3: { constexpr int VPT = 3; __VA_ARGS__; break; }                \
    case 4: { constexpr int VPT = 4; __VA_ARGS__; break; }                \
    default: break;                                                       \
  }

#define UB_DISPATCH_DTYPE(DTYPE_VALUE, ...)                               \
  switch (DTYPE_VALUE) {                                                  \
    case kF32: { using T = float; __VA_ARGS__; break; }                   \
    case kF16: { using T = __half; __VA_ARGS__; break; }                  \
    case kBF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }          \
    default: break;                                                       \
  }

template <typename T, bool kRMS, bool kFused>
static void run_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd, int rows,
                    int cols, float eps, const void* bias, const void* residual, void* summed, float p,
                    unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
  int vpt;
  NormGeom g = make_geom(rows, cols, VecTraits<T>::kElems, vpt);
  const int grid = fwd_grid(g);
  const float keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  UB_DISPATCH_VPT(vpt, (norm_fwd_kernel<T, VPT, kRMS, kFused><<<grid, kNormThreads, 0, stream>>>(
                           (const T*)x, (const T*)gamma, (const T*)beta, (T*)y, mean, rstd, g, eps, (const T*)bias,
                           (const T*)residual, (T*)summed, p, keep_scale, seed, offset)));
}

template <typename T, bool kRMS, bool kFused>
static void run_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma, void* dx,
                    void* dgamma, void* dbeta, float* dg_part, float* db_part, int rows, int cols, void* dx_drop,
                    float p, unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
  int vpt;
  NormGeom g = make_geom(rows, cols, VecTraits<T>::kElems, vpt);
  const int parts = norm_bwd_parts(rows, cols);
  const size_t smem = (size_t)(kRMS ? 1 : 2) * cols * sizeof(float);
  const float keep_scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  UB_DISPATCH_VPT(vpt, {
    auto kern = norm_bwd_kernel<T, VPT, kRMS, kFused>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<parts, kNormThreads, smem, stream>>>((const T*)dy, (const T*)x, mean, rstd, (const T*)gamma, (T*)dx,
                                                dg_part, db_part, g, (T*)dx_drop, p, keep_scale, seed, offset);
  });
  norm_param_grad_kernel<T><<<(cols + 31) / 32, 256, 0, stream>>>(dg_part, kRMS ? nullptr : db_part, parts, cols,
                                                                  (T*)dgamma, (T*)dbeta);
}

void launch_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                          int rows, int cols, float eps, int dtype, cudaStream_t stream) {
  UB_DISPATCH_DTYPE(dtype, (run_fwd<T, false, false>(x, gamma, beta, y, mean, rstd, rows, cols, eps, nullptr, nullptr,
                                                     nullptr, 0.f, 0, 0, stream)));
}

void launch_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const void* gamma,
                          void* dx, void* dgamma, void* dbeta, float* dgamma_part, float* dbeta_part,
                          unsigned* counter, int rows, int cols, int dtype, cudaStream_t stream) {
  (void)counter;
  UB_DISPATCH_DTYPE(dtype, (run_bwd<T, false, false>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, dgamma_part,
                                                     dbeta_part, rows, cols, nullptr, 0.f, 0, 0, stream)));
}

void launch_rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, int rows, int cols, float eps,
                        int dtype, cudaStream_t stream) {
  UB_DISPATCH_DTYPE(dtype, (run_fwd<T, true, false>(x, gamma, nullptr, y, nullptr, rstd, rows, cols, eps, nullptr,
                                                    nullptr, nullptr, 0.f, 0, 0, stream)));
}

void launch_rmsnorm_bwd(const void* dy, const void* x, const float* rstd, const void* gamma, void* dx, void* dgamma,
                        float* dgamma_part, unsigned* counter, int rows, int cols, int dtype, cudaStream_t stream) {
  (void)counter;
  UB_DISPATCH_DTYPE(dtype, (run_bwd<T, true, false>(dy, x, nullptr, rstd, gamma, dx, dgamma, nullptr, dgamma_part,
                                                    nullptr, rows, cols, nullptr, 0.f, 0, 0, stream)));
}

void launch_bias_dropout_add_ln_fwd(const void* x, const void* bias, const void* residual, const void* gamma,
                                    const void* beta, void* y, void* summed, float* mean, float* rstd, int rows,
                                    int cols, float p, float eps, unsigned long long seed, unsigned long long offset,
                                    int dtype, cudaStream_t stream) {
  if (dtype == kF16) {
    run_fwd<__half, false, true>(x, gamma, beta, y, mean, rstd, rows, cols, eps, bias, residual, summed, p, seed,
                                 offset, stream);
  } else if (dtype == kBF16) {
    run_fwd<__nv_bfloat16, false, true>(x, gamma, beta, y, mean, rstd, rows, cols, eps, bias, residual, summed, p,
                                        seed, offset, stream);
  }
}

void launch_bias_dropout_add_ln_bwd(const void* dy, const void* summed, const float* mean, const float* rstd,
                                    const void* gamma, void* dsum, void* dx, void* dgamma, void* dbeta,
                                    float* dgamma_part, float* dbeta_part, unsigned* counter, int rows, int cols,
                                    float p, unsigned long long seed, unsigned long long offset, int dtype,
                                    cudaStream_t stream) {
  (void)counter;
  if (dtype == kF16) {
    run_bwd<__half, false, true>(dy, summed, mean, rstd, gamma, dsum, dgamma, dbeta, dgamma_part, dbeta_part, rows,
                                 cols, dx, p, seed, offset, stream);
  } else if (dtype == kBF16) {
    run_bwd<__nv_bfloat16, false, true>(dy, summed, mean, rstd, gamma, dsum, dgamma, dbeta, dgamma_part, dbeta_part,
                                        rows, cols, dx, p, seed, offset, stream);
  }
}

}  // namespace ub
