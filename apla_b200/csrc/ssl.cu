// HBM-bound row kernels of the DINOv2 self-supervised objective (BASELINE config C4): everything
// that touches the [rows, K = 65 536] prototype scores, plus the small vector kernels around them.
//   * softmax_center     teacher targets softmax((t - centre) / temp)      dino_clstoken_loss.py:28-31, ibot_patch_loss.py:39-51
//   * colsum / centre EMA                                                   dino_clstoken_loss.py:76-98, ibot_patch_loss.py:123-145
//   * soft_ce fwd / bwd  -sum_k q_k log_softmax(s / T)_k with q the sum of one or two teacher rows (DINO crop pairs,
//                        iBOT masked patches with per-row weights)           dino_clstoken_loss.py:62-74, ibot_patch_loss.py:102-121
//   * l2norm fwd / bwd   F.normalize in DINOHead.forward                    layers/dino_head.py:36-41
//   * weightnorm fwd/bwd weight_norm(Linear(bottleneck, K)), W = g v/||v||  layers/dino_head.py:27-31
//   * koleo              nearest neighbour, distance, loss and gradient     loss/koleo_loss.py:17-45
//   * ema                teacher <- m teacher + (1 - m) student             models.py:437-447
// (paths relative to src/self_supervised/dinov2/ of the reference).  All arithmetic is fp32; one CTA per K-wide row with
// an online max / sum-of-exponentials pass and 128-bit loads, one warp per row for the narrow (bottleneck-wide) rows.
// Reductions are fixed-order (no atomics): the losses are reproducible run to run.
// STATUS: compiled for sm_100a; round 1's GPU budget was spent before these existed, so their first HARDWARE run is round
// 2's first job (tools/round2_first_gpu.sh).  Until then this source is executed, unchanged, by a CPU SIMT emulator
// (tests/emu/, one OS thread per CUDA thread, real barriers and warp shuffles) against the same parity tests as on the GPU
// (tests/test_ssl_emu.py): indexing, reductions and barrier placement are pinned; speed and fast-math rounding are not.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace apla {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions, result broadcast to every thread.  `red` = 33 floats of shared memory; the trailing barrier
// makes back-to-back calls on the same buffer safe.  blockDim.x is a multiple of 32, at most 1024.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = l < nw ? red[l] : 0.f;
    t = warp_sum(t);
    if (l == 0) red[32] = t;
  }
  __syncthreads();
  const float r = red[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = l < nw ? red[l] : -INFINITY;
    t = warp_max(t);
    if (l == 0) red[32] = t;
  }
  __syncthreads();
  const float r = red[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ float max4(const float4& a) { return fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)); }

// element i of a row that is fp32 or bf16
__device__ __forceinline__ float ld_elem(const void* p, int64_t i, bool f32) {
  return f32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

constexpr int kRowThreads = 512;

// ------------------------------------------------------------------------------------------------
// teacher targets: out[row] = softmax((t[row] - centre) * inv_temp), K % 4 == 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
softmax_center_kernel(const float* __restrict__ t, int64_t ldt, const float* __restrict__ center, float inv_temp, int K,
                      float* __restrict__ out, int64_t ldo) {
  __shared__ float red[33];
  const int64_t row = blockIdx.x;
  const float4* tr = reinterpret_cast<const float4*>(t + row * ldt);
  const float4* cr = reinterpret_cast<const float4*>(center);
  const int nv = K >> 2;
  float m = -INFINITY, l = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 a = tr[i];
    const float4 c = __ldg(cr + i);
    a.x = (a.x - c.x) * inv_temp; a.y = (a.y - c.y) * inv_temp; a.z = (a.z - c.z) * inv_temp; a.w = (a.w - c.w) * inv_temp;
    const float cm = max4(a);
    if (cm > m) { l *= __expf(m - cm); m = cm; }
    l += (__expf(a.x - m) + __expf(a.y - m)) + (__expf(a.z - m) + __expf(a.w - m));
  }
  const float M = block_max(m, red);
  const float L = block_sum(l > 0.f ? l * __expf(m - M) : 0.f, red);
  const float inv = 1.f / L;
  float4* orow = reinterpret_cast<float4*>(out + row * ldo);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 a = tr[i];
    const float4 c = __ldg(cr + i);
    a.x = __expf((a.x - c.x) * inv_temp - M) * inv; a.y = __expf((a.y - c.y) * inv_temp - M) * inv;
    a.z = __expf((a.z - c.z) * inv_temp - M) * inv; a.w = __expf((a.w - c.w) * inv_temp - M) * inv;
    orow[i] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a [rows, K] fp32 matrix in two fixed-order stages (partials[splits, K], then out[K])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_f32_partial_kernel(const float* __restrict__ a, int64_t ld, int rows, int K, int rows_per_split,
                          float* __restrict__ partial) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= K) return;
  const int r0 = blockIdx.y * rows_per_split;
  const int r1 = min(rows, r0 + rows_per_split);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    acc0 += a[int64_t(r) * ld + col];
    acc1 += a[int64_t(r + 1) * ld + col];
    acc2 += a[int64_t(r + 2) * ld + col];
    acc3 += a[int64_t(r + 3) * ld + col];
  }
  for (; r < r1; ++r) acc0 += a[int64_t(r) * ld + col];
  partial[int64_t(blockIdx.y) * K + col] = (acc0 + acc1) + (acc2 + acc3);
}
__global__ void __launch_bounds__(256)
colsum_f32_final_kernel(const float* __restrict__ partial, int splits, int K, float scale, float* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= K) return;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[int64_t(s) * K + col];
  out[col] = acc * scale;
}
__global__ void __launch_bounds__(256)
center_ema_kernel(float* __restrict__ center, const float* __restrict__ batch_sum, int K, float inv_count, float momentum) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < K) center[col] = center[col] * momentum + batch_sum[col] * inv_count * (1.f - momentum);
}

// ------------------------------------------------------------------------------------------------
// soft-target cross-entropy of one student row against the sum q of one or two teacher rows (row % t_rows of t0 / t1):
//   z = s * inv_temp,  lse = logsumexp(z),  mass = sum_k q_k,  row_loss = -w (sum_k q_k z_k - mass lse)
//   dz/ds:             ds_k = -w inv_temp g (q_k - mass exp(z_k - lse))
// with w = w_uniform * (w_row ? w_row[row] : 1) and g = *gscale (the upstream gradient, read on the device).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
soft_ce_fwd_kernel(const float* __restrict__ s, int64_t lds, int K, const float* __restrict__ t0,
                   const float* __restrict__ t1, int64_t ldt, int t_rows, const float* __restrict__ w_row,
                   float w_uniform, float inv_temp, float* __restrict__ row_loss, float* __restrict__ lse_out,
                   float* __restrict__ mass_out) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const int trow = row % t_rows;
  const float4* sr = reinterpret_cast<const float4*>(s + int64_t(row) * lds);
  const float4* q0 = reinterpret_cast<const float4*>(t0 + int64_t(trow) * ldt);
  const float4* q1 = t1 ? reinterpret_cast<const float4*>(t1 + int64_t(trow) * ldt) : nullptr;
  const int nv = K >> 2;
  float m = -INFINITY, l = 0.f, dot = 0.f, mass = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 a = sr[i];
    float4 q = q0[i];
    if (q1) {
      const float4 b = q1[i];
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    a.x *= inv_temp; a.y *= inv_temp; a.z *= inv_temp; a.w *= inv_temp;
    const float cm = max4(a);
    if (cm > m) { l *= __expf(m - cm); m = cm; }
    l += (__expf(a.x - m) + __expf(a.y - m)) + (__expf(a.z - m) + __expf(a.w - m));
    dot += (q.x * a.x + q.y * a.y) + (q.z * a.z + q.w * a.w);
    mass += (q.x + q.y) + (q.z + q.w);
  }
  const float M = block_max(m, red);
  const float L = block_sum(l > 0.f ? l * __expf(m - M) : 0.f, red);
  const float DOT = block_sum(dot, red);
  const float MASS = block_sum(mass, red);
  if (threadIdx.x == 0) {
    const float lse = M + logf(L);
    const float w = w_uniform * (w_row ? w_row[row] : 1.f);
    row_loss[row] = -w * (DOT - MASS * lse);
    lse_out[row] = lse;
    mass_out[row] = MASS;
  }
}

template <typename OutT>
__global__ void __launch_bounds__(kRowThreads)
soft_ce_bwd_kernel(const float* __restrict__ s, int64_t lds, int K, const float* __restrict__ t0,
                   const float* __restrict__ t1, int64_t ldt, int t_rows, const float* __restrict__ w_row,
                   float w_uniform, float inv_temp, const float* __restrict__ lse_in, const float* __restrict__ mass_in,
                   const float* __restrict__ gscale, OutT* __restrict__ ds, int64_t ldd) {
  const int row = blockIdx.x;
  const int trow = row % t_rows;
  const float4* sr = reinterpret_cast<const float4*>(s + int64_t(row) * lds);
  const float4* q0 = reinterpret_cast<const float4*>(t0 + int64_t(trow) * ldt);
  const float4* q1 = t1 ? reinterpret_cast<const float4*>(t1 + int64_t(trow) * ldt) : nullptr;
  const float lse = lse_in[row], mass = mass_in[row];
  const float c = -w_uniform * (w_row ? w_row[row] : 1.f) * inv_temp * (gscale ? gscale[0] : 1.f);
  OutT* drow = ds + int64_t(row) * ldd;
  const int nv = K >> 2;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const float4 a = sr[i];
    float4 q = q0[i];
    if (q1) {
      const float4 b = q1[i];
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float4 g;
    g.x = c * (q.x - mass * __expf(a.x * inv_temp - lse));
    g.y = c * (q.y - mass * __expf(a.y * inv_temp - lse));
    g.z = c * (q.z - mass * __expf(a.z * inv_temp - lse));
    g.w = c * (q.w - mass * __expf(a.w * inv_temp - lse));
    if constexpr (sizeof(OutT) == 4) {
      reinterpret_cast<float4*>(drow)[i] = g;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(g.x, g.y), hi = __floats2bfloat162_rn(g.z, g.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(drow)[i] = pk;
    }
  }
}

// forward and backward of one row in ONE launch (used when the upstream gradient is known when the loss is taken, i.e. by
// ssl_objective): pass 1 = soft_ce_fwd_kernel, pass 2 = soft_ce_bwd_kernel walking the row BACKWARDS so that it starts
// on the float4s this thread touched last -- the second read of the student / teacher rows is meant to come from L2
// (0.5-0.75 MB per resident CTA), which makes the HBM traffic of the pair read-once + one write of ds.
// w_fwd scales the reported row loss, w_bwd the gradient (loss_dict scale vs. scale x loss weight).
template <typename OutT>
__global__ void __launch_bounds__(kRowThreads)
soft_ce_fused_kernel(const float* __restrict__ s, int64_t lds, int K, const float* __restrict__ t0,
                     const float* __restrict__ t1, int64_t ldt, int t_rows, const float* __restrict__ w_row, float w_fwd,
                     float w_bwd, float inv_temp, const float* __restrict__ gscale, float* __restrict__ row_loss,
                     OutT* __restrict__ ds, int64_t ldd) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const int trow = row % t_rows;
  const float4* sr = reinterpret_cast<const float4*>(s + int64_t(row) * lds);
  const float4* q0 = reinterpret_cast<const float4*>(t0 + int64_t(trow) * ldt);
  const float4* q1 = t1 ? reinterpret_cast<const float4*>(t1 + int64_t(trow) * ldt) : nullptr;
  const int nv = K >> 2;
  float m = -INFINITY, l = 0.f, dot = 0.f, mass = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 a = sr[i];
    float4 q = q0[i];
    if (q1) {
      const float4 b = q1[i];
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    a.x *= inv_temp; a.y *= inv_temp; a.z *= inv_temp; a.w *= inv_temp;
    const float cm = max4(a);
    if (cm > m) { l *= __expf(m - cm); m = cm; }
    l += (__expf(a.x - m) + __expf(a.y - m)) + (__expf(a.z - m) + __expf(a.w - m));
    dot += (q.x * a.x + q.y * a.y) + (q.z * a.z + q.w * a.w);
    mass += (q.x + q.y) + (q.z + q.w);
  }
  const float M = block_max(m, red);
  const float L = block_sum(l > 0.f ? l * __expf(m - M) : 0.f, red);
  const float DOT = block_sum(dot, red);
  const float MASS = block_sum(mass, red);
  const float lse = M + logf(L);
  const float wr = w_row ? w_row[row] : 1.f;
  if (threadIdx.x == 0) row_loss[row] = -w_fwd * wr * (DOT - MASS * lse);
  const float c = -w_bwd * wr * inv_temp * (gscale ? gscale[0] : 1.f);
  OutT* drow = ds + int64_t(row) * ldd;
  const int tid = threadIdx.x, bd = blockDim.x;
  const int last = tid < nv ? tid + ((nv - 1 - tid) / bd) * bd : -1;   // the last i of this thread's first pass
  for (int i = last; i >= 0; i -= bd) {
    const float4 a = sr[i];
    float4 q = q0[i];
    if (q1) {
      const float4 b = q1[i];
      q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
    }
    float4 g;
    g.x = c * (q.x - MASS * __expf(a.x * inv_temp - lse));
    g.y = c * (q.y - MASS * __expf(a.y * inv_temp - lse));
    g.z = c * (q.z - MASS * __expf(a.z * inv_temp - lse));
    g.w = c * (q.w - MASS * __expf(a.w * inv_temp - lse));
    if constexpr (sizeof(OutT) == 4) {
      reinterpret_cast<float4*>(drow)[i] = g;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(g.x, g.y), hi = __floats2bfloat162_rn(g.z, g.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(drow)[i] = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sinkhorn-Knopp teacher targets (dino_clstoken_loss.py:33-60, ibot_patch_loss.py:53-83) in the [samples, K] layout:
//   sk_exp        P = exp(t * inv_temp)
//   sk_normalize  one iteration on a row: p_k <- p_k * col_scale / colsum[k]  (each prototype's mass -> 1/K), then
//                 p_k <- p_k * row_scale / sum_k p_k  (each sample's mass -> row_scale); the column sums come from
//                 colsum_f32 (+ the data-parallel all-reduce).  The reference's initial division by the total mass
//                 cancels in the first column normalisation and is not performed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowThreads)
sk_exp_kernel(const float* __restrict__ t, int64_t ldt, float inv_temp, int K, float* __restrict__ out, int64_t ldo) {
  const int64_t row = blockIdx.x;
  const float4* tr = reinterpret_cast<const float4*>(t + row * ldt);
  float4* orow = reinterpret_cast<float4*>(out + row * ldo);
  const int nv = K >> 2;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 a = tr[i];
    a.x = __expf(a.x * inv_temp); a.y = __expf(a.y * inv_temp); a.z = __expf(a.z * inv_temp); a.w = __expf(a.w * inv_temp);
    orow[i] = a;
  }
}

__global__ void __launch_bounds__(kRowThreads)
sk_normalize_kernel(float* __restrict__ p, int64_t ld, int K, const float* __restrict__ colsum, float col_scale,
                    float row_scale) {
  __shared__ float red[33];
  const int64_t row = blockIdx.x;
  float4* pr = reinterpret_cast<float4*>(p + row * ld);
  const float4* cs = reinterpret_cast<const float4*>(colsum);
  const int nv = K >> 2;
  float acc = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const float4 a = pr[i];
    const float4 c = __ldg(cs + i);
    acc += (a.x * (col_scale / c.x) + a.y * (col_scale / c.y)) + (a.z * (col_scale / c.z) + a.w * (col_scale / c.w));
  }
  const float inv = row_scale / block_sum(acc, red);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float4 a = pr[i];
    const float4 c = __ldg(cs + i);
    a.x = a.x * (col_scale / c.x) * inv; a.y = a.y * (col_scale / c.y) * inv;
    a.z = a.z * (col_scale / c.z) * inv; a.w = a.w * (col_scale / c.w) * inv;
    pr[i] = a;
  }
}

// out[0] = scale * sum_i a[i] in a fixed order (single CTA)
__global__ void __launch_bounds__(1024) sum_f32_kernel(const float* __restrict__ a, int n, float scale, float* __restrict__ out) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = acc * scale;
}

// ------------------------------------------------------------------------------------------------
// narrow rows (one warp per row): L2 normalisation and weight normalisation
// ------------------------------------------------------------------------------------------------
// y = x / max(||x||, eps)
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const void* __restrict__ x, int64_t ldx, bool x_f32, int rows, int d, float eps,
                  __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32, int64_t ldy) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = ld_elem(x, int64_t(row) * ldx + j, x_f32);
    ss += v * v;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), eps);
  for (int j = lane; j < d; j += 32) {
    const float v = ld_elem(x, int64_t(row) * ldx + j, x_f32) * inv;
    if (y_bf16) y_bf16[int64_t(row) * ldy + j] = __float2bfloat16_rn(v);
    if (y_f32) y_f32[int64_t(row) * ldy + j] = v;
  }
}
// dx = (dy - y (y . dy)) / ||x||  when ||x|| > eps, dy / eps otherwise (the clamp passes no gradient to the norm)
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const void* __restrict__ x, int64_t ldx, bool x_f32, const void* __restrict__ dy, int64_t ld_dy,
                  bool g_f32, int rows, int d, float eps, void* __restrict__ dx, int64_t ld_dx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float ss = 0.f, xd = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v = ld_elem(x, int64_t(row) * ldx + j, x_f32);
    const float g = ld_elem(dy, int64_t(row) * ld_dy + j, g_f32);
    ss += v * v;
    xd += v * g;
  }
  ss = warp_sum(ss);
  xd = warp_sum(xd);
  const float nrm = sqrtf(ss);
  const bool clamped = !(nrm > eps);
  const float inv = 1.f / fmaxf(nrm, eps);
  const float coef = clamped ? 0.f : xd * inv * inv * inv;  // x (x . dy) / ||x||^3
  for (int j = lane; j < d; j += 32) {
    const float v = ld_elem(x, int64_t(row) * ldx + j, x_f32);
    const float g = ld_elem(dy, int64_t(row) * ld_dy + j, g_f32);
    const float r = g * inv - v * coef;
    if (g_f32) reinterpret_cast<float*>(dx)[int64_t(row) * ld_dx + j] = r;
    else reinterpret_cast<__nv_bfloat16*>(dx)[int64_t(row) * ld_dx + j] = __float2bfloat16_rn(r);
  }
}

// W[k, :] = g[k] v[k, :] / ||v[k, :]||
__global__ void __launch_bounds__(256)
weightnorm_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v, int K, int d,
                      __nv_bfloat16* __restrict__ w_bf16, float* __restrict__ w_f32) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= K) return;
  const float* vr = v + int64_t(row) * d;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) ss += vr[j] * vr[j];
  const float sc = g[row] / sqrtf(warp_sum(ss));
  for (int j = lane; j < d; j += 32) {
    const float r = vr[j] * sc;
    if (w_bf16) w_bf16[int64_t(row) * d + j] = __float2bfloat16_rn(r);
    if (w_f32) w_f32[int64_t(row) * d + j] = r;
  }
}
// dg[k] = (dW[k] . v[k]) / ||v[k]|| ;  dv[k] = g[k] / ||v[k]|| (dW[k] - v^[k] (v^[k] . dW[k]))
__global__ void __launch_bounds__(256)
weightnorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v, const float* __restrict__ dW, int64_t ld_dw,
                      int K, int d, float* __restrict__ dg, float* __restrict__ dv) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= K) return;
  const float* vr = v + int64_t(row) * d;
  const float* dr = dW + int64_t(row) * ld_dw;
  float ss = 0.f, vd = 0.f;
  for (int j = lane; j < d; j += 32) {
    ss += vr[j] * vr[j];
    vd += vr[j] * dr[j];
  }
  ss = warp_sum(ss);
  vd = warp_sum(vd);
  const float inv = 1.f / sqrtf(ss);
  if (lane == 0 && dg) dg[row] = vd * inv;
  if (dv) {
    const float sc = g[row] * inv;
    const float proj = vd * inv * inv;  // (v . dW) / ||v||^2
    for (int j = lane; j < d; j += 32) dv[int64_t(row) * d + j] = sc * (dr[j] - vr[j] * proj);
  }
}

// ------------------------------------------------------------------------------------------------
// KoLeo on L2-normalised rows xn[groups * n, D] (each group of n rows is one independent call of the reference)
// ------------------------------------------------------------------------------------------------
// nearest neighbour by inner product (diagonal excluded, first maximum wins), distance ||xn_i - xn_nn + 1e-8||,
// row_loss[i] = -log(dist + eps) * w / n
__global__ void __launch_bounds__(256)
koleo_nn_kernel(const float* __restrict__ xn, int n, int D, float eps, float w, int* __restrict__ nn,
                float* __restrict__ dist, float* __restrict__ row_loss) {
  extern __shared__ float sm[];  // [D] the row, then 33 floats, then per-warp best value / index
  float* xi = sm;
  float* red = sm + D;
  float* bval = red + 33;
  int* bidx = reinterpret_cast<int*>(bval + 32);
  const int i = blockIdx.x, grp = blockIdx.y;
  const float* base = xn + int64_t(grp) * n * D;
  for (int k = threadIdx.x; k < D; k += blockDim.x) xi[k] = base[int64_t(i) * D + k];
  __syncthreads();
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float best = -INFINITY;
  int besti = -1;
  for (int j = wid; j < n; j += nw) {
    float acc = 0.f;
    const float* xj = base + int64_t(j) * D;
    for (int k = lane; k < D; k += 32) acc += xi[k] * xj[k];
    acc = warp_sum(acc);
    if (j == i) acc = -1.f;
    if (acc > best) { best = acc; besti = j; }
  }
  if (lane == 0) { bval[wid] = best; bidx[wid] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = -INFINITY;
    int bi = -1;
    for (int k = 0; k < nw; ++k)
      if (bidx[k] >= 0 && (bval[k] > b || (bval[k] == b && bidx[k] < bi))) { b = bval[k]; bi = bidx[k]; }
    bidx[0] = bi;
  }
  __syncthreads();
  // a row holding NaN wins no comparison: fall back to a valid neighbour so that the NaN propagates through the
  // distance (as in the reference) instead of indexing with -1
  const int j = bidx[0] >= 0 ? bidx[0] : (i + 1) % n;
  const float* xj = base + int64_t(j) * D;
  float acc = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const float u = xi[k] - xj[k] + 1e-8f;
    acc += u * u;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float dd = sqrtf(acc);
    nn[grp * n + i] = j;
    dist[grp * n + i] = dd;
    row_loss[grp * n + i] = -logf(dd + eps) * w / n;
  }
}
// gradient of sum_i row_loss[i] with respect to the UN-normalised rows x (xn = x * inv_norm): own term plus the terms
// of every row that chose i as its neighbour, then the backward of the normalisation
__global__ void __launch_bounds__(256)
koleo_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xn, int n, int D, float eps, float norm_eps,
                 float w, const int* __restrict__ nn, const float* __restrict__ dist, const float* __restrict__ gscale,
                 float* __restrict__ dx) {
  extern __shared__ float sm[];  // [D] gradient w.r.t. xn_i, then 33 floats
  float* gi = sm;
  float* red = sm + D;
  const int i = blockIdx.x, grp = blockIdx.y;
  const float* xb = xn + int64_t(grp) * n * D;
  const int* nnb = nn + grp * n;
  const float* db = dist + grp * n;
  const float up = -(w / n) * (gscale ? gscale[0] : 1.f);
  const float* xi = xb + int64_t(i) * D;
  {
    const float di = db[i];
    const float ci = up / ((di + eps) * di);
    const float* xj = xb + int64_t(nnb[i]) * D;
    // a row that is its own neighbour (a group of one row): its two terms cancel exactly -- written out, not left to
    // c * u - c * u, which a fused multiply-add turns into the rounding error of the first product
    const bool self = nnb[i] == i;
    for (int k = threadIdx.x; k < D; k += blockDim.x) gi[k] = self ? 0.f : ci * (xi[k] - xj[k] + 1e-8f);
  }
  for (int j = 0; j < n; ++j) {
    if (nnb[j] != i || j == i) continue;  // uniform over the block
    const float dj = db[j];
    const float cj = up / ((dj + eps) * dj);
    const float* xj = xb + int64_t(j) * D;
    for (int k = threadIdx.x; k < D; k += blockDim.x) gi[k] -= cj * (xj[k] - xi[k] + 1e-8f);
  }
  // each thread only ever touches its own k: no barrier needed until the dot product
  const float* xr = x + (int64_t(grp) * n + i) * D;
  float ss = 0.f, xg = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    ss += xr[k] * xr[k];
    xg += xr[k] * gi[k];
  }
  ss = block_sum(ss, red);
  xg = block_sum(xg, red);
  const float nrm = sqrtf(ss);
  const float inv = 1.f / fmaxf(nrm, norm_eps);
  const float coef = nrm > norm_eps ? xg * inv * inv * inv : 0.f;
  float* dr = dx + (int64_t(grp) * n + i) * D;
  for (int k = threadIdx.x; k < D; k += blockDim.x) dr[k] = gi[k] * inv - xr[k] * coef;
}

// teacher <- m teacher + (1 - m) student
__global__ void __launch_bounds__(256)
ema_kernel(float* __restrict__ t, const float* __restrict__ s, int64_t n, float m) {
  const float om = 1.f - m;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    t[i] = t[i] * m + s[i] * om;
}
__global__ void __launch_bounds__(256)
ema_kernel_v4(float4* __restrict__ t, const float4* __restrict__ s, int64_t n4, float m) {
  const float om = 1.f - m;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 a = t[i];
    const float4 b = s[i];
    a.x = a.x * m + b.x * om; a.y = a.y * m + b.y * om; a.z = a.z * m + b.z * om; a.w = a.w * m + b.w * om;
    t[i] = a;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// ds rows are written as float4 (f32) or as uint2 = four bf16
inline bool aligned_ds(const void* p, int is_bf16) { return (reinterpret_cast<uintptr_t>(p) & (is_bf16 ? 7 : 15)) == 0; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int ssl_softmax_center(const float* t, int64_t ldt, const float* center, float inv_temp, int rows, int K, float* out,
                       int64_t ldo, cudaStream_t s) {
  APLA_CHECK(rows >= 0 && K > 0 && K % 4 == 0, "softmax_center: K=%d must be a positive multiple of 4", K);
  APLA_CHECK(ldt % 4 == 0 && ldo % 4 == 0 && aligned16(t) && aligned16(out) && aligned16(center),
             "softmax_center: rows must be 16-byte aligned");
  if (rows == 0) return 0;
  softmax_center_kernel<<<rows, kRowThreads, 0, s>>>(t, ldt, center, inv_temp, K, out, ldo);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_colsum_f32(const float* a, int64_t ld, int rows, int K, float* ws, int splits, float scale, float* out,
                   cudaStream_t s) {
  APLA_CHECK(rows >= 0 && K > 0 && splits >= 1, "colsum_f32: bad sizes rows=%d K=%d splits=%d", rows, K, splits);
  const int rps = rows > 0 ? cdiv(rows, splits) : 1;
  colsum_f32_partial_kernel<<<dim3(cdiv(K, 256), splits), 256, 0, s>>>(a, ld, rows, K, rps, ws);
  colsum_f32_final_kernel<<<cdiv(K, 256), 256, 0, s>>>(ws, splits, K, scale, out);
  count_launch(2);
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_center_ema(float* center, const float* batch_sum, int K, float inv_count, float momentum, cudaStream_t s) {
  APLA_CHECK(K > 0, "center_ema: K=%d", K);
  center_ema_kernel<<<cdiv(K, 256), 256, 0, s>>>(center, batch_sum, K, inv_count, momentum);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

static int soft_ce_check(const float* sp, int64_t lds, int rows, int K, const float* t0, const float* t1, int64_t ldt,
                         int t_rows) {
  APLA_CHECK(rows >= 0 && K > 0 && K % 4 == 0, "soft_ce: K=%d must be a positive multiple of 4", K);
  APLA_CHECK(t_rows > 0, "soft_ce: t_rows=%d", t_rows);
  APLA_CHECK(lds % 4 == 0 && ldt % 4 == 0 && aligned16(sp) && aligned16(t0) && aligned16(t1),
             "soft_ce: rows must be 16-byte aligned");
  return 0;
}

int ssl_soft_ce_fwd(const float* sp, int64_t lds, int rows, int K, const float* t0, const float* t1, int64_t ldt,
                    int t_rows, const float* w_row, float w_uniform, float inv_temp, float* row_loss, float* lse,
                    float* mass, cudaStream_t s) {
  if (int rc = soft_ce_check(sp, lds, rows, K, t0, t1, ldt, t_rows)) return rc;
  if (rows == 0) return 0;
  soft_ce_fwd_kernel<<<rows, kRowThreads, 0, s>>>(sp, lds, K, t0, t1, ldt, t_rows, w_row, w_uniform, inv_temp, row_loss,
                                                  lse, mass);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_soft_ce_bwd(const float* sp, int64_t lds, int rows, int K, const float* t0, const float* t1, int64_t ldt,
                    int t_rows, const float* w_row, float w_uniform, float inv_temp, const float* lse, const float* mass,
                    const float* gscale, void* ds, int64_t ldd, int ds_is_bf16, cudaStream_t s) {
  if (int rc = soft_ce_check(sp, lds, rows, K, t0, t1, ldt, t_rows)) return rc;
  APLA_CHECK(ldd % 4 == 0 && aligned_ds(ds, ds_is_bf16), "soft_ce_bwd: ds rows must be 16-byte (f32) / 8-byte (bf16) aligned");
  if (rows == 0) return 0;
  if (ds_is_bf16)
    soft_ce_bwd_kernel<__nv_bfloat16><<<rows, kRowThreads, 0, s>>>(sp, lds, K, t0, t1, ldt, t_rows, w_row, w_uniform,
                                                                  inv_temp, lse, mass, gscale,
                                                                  reinterpret_cast<__nv_bfloat16*>(ds), ldd);
  else
    soft_ce_bwd_kernel<float><<<rows, kRowThreads, 0, s>>>(sp, lds, K, t0, t1, ldt, t_rows, w_row, w_uniform, inv_temp,
                                                          lse, mass, gscale, reinterpret_cast<float*>(ds), ldd);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_soft_ce_fwd_bwd(const float* sp, int64_t lds, int rows, int K, const float* t0, const float* t1, int64_t ldt,
                        int t_rows, const float* w_row, float w_fwd, float w_bwd, float inv_temp, const float* gscale,
                        float* row_loss, void* ds, int64_t ldd, int ds_is_bf16, cudaStream_t s) {
  if (int rc = soft_ce_check(sp, lds, rows, K, t0, t1, ldt, t_rows)) return rc;
  APLA_CHECK(ds != nullptr && ldd % 4 == 0 && aligned_ds(ds, ds_is_bf16),
             "soft_ce_fwd_bwd: ds rows must be 16-byte (f32) / 8-byte (bf16) aligned");
  if (rows == 0) return 0;
  if (ds_is_bf16)
    soft_ce_fused_kernel<__nv_bfloat16><<<rows, kRowThreads, 0, s>>>(sp, lds, K, t0, t1, ldt, t_rows, w_row, w_fwd, w_bwd,
                                                                    inv_temp, gscale, row_loss,
                                                                    reinterpret_cast<__nv_bfloat16*>(ds), ldd);
  else
    soft_ce_fused_kernel<float><<<rows, kRowThreads, 0, s>>>(sp, lds, K, t0, t1, ldt, t_rows, w_row, w_fwd, w_bwd,
                                                            inv_temp, gscale, row_loss, reinterpret_cast<float*>(ds), ldd);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_sk_exp(const float* t, int64_t ldt, float inv_temp, int rows, int K, float* out, int64_t ldo, cudaStream_t s) {
  APLA_CHECK(rows >= 0 && K > 0 && K % 4 == 0, "sk_exp: K=%d must be a positive multiple of 4", K);
  APLA_CHECK(ldt % 4 == 0 && ldo % 4 == 0 && aligned16(t) && aligned16(out), "sk_exp: rows must be 16-byte aligned");
  if (rows == 0) return 0;
  sk_exp_kernel<<<rows, kRowThreads, 0, s>>>(t, ldt, inv_temp, K, out, ldo);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_sk_normalize(float* p, int64_t ld, int rows, int K, const float* colsum, float col_scale, float row_scale,
                     cudaStream_t s) {
  APLA_CHECK(rows >= 0 && K > 0 && K % 4 == 0, "sk_normalize: K=%d must be a positive multiple of 4", K);
  APLA_CHECK(ld % 4 == 0 && aligned16(p) && aligned16(colsum), "sk_normalize: rows must be 16-byte aligned");
  if (rows == 0) return 0;
  sk_normalize_kernel<<<rows, kRowThreads, 0, s>>>(p, ld, K, colsum, col_scale, row_scale);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_sum_f32(const float* a, int n, float scale, float* out, cudaStream_t s) {
  APLA_CHECK(n >= 0, "sum_f32: n=%d", n);
  sum_f32_kernel<<<1, 1024, 0, s>>>(a, n, scale, out);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_l2norm_fwd(const void* x, int64_t ldx, int x_is_f32, int rows, int d, float eps, void* y_bf16, float* y_f32,
                   int64_t ldy, cudaStream_t s) {
  APLA_CHECK(rows >= 0 && d > 0, "l2norm_fwd: rows=%d d=%d", rows, d);
  APLA_CHECK(y_bf16 || y_f32, "l2norm_fwd: no output");
  if (rows == 0) return 0;
  l2norm_fwd_kernel<<<cdiv(rows, 8), 256, 0, s>>>(x, ldx, x_is_f32 != 0, rows, d, eps,
                                                  reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, ldy);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_l2norm_bwd(const void* x, int64_t ldx, int x_is_f32, const void* dy, int64_t ld_dy, int grads_are_f32, int rows,
                   int d, float eps, void* dx, int64_t ld_dx, cudaStream_t s) {
  APLA_CHECK(rows >= 0 && d > 0, "l2norm_bwd: rows=%d d=%d", rows, d);
  if (rows == 0) return 0;
  l2norm_bwd_kernel<<<cdiv(rows, 8), 256, 0, s>>>(x, ldx, x_is_f32 != 0, dy, ld_dy, grads_are_f32 != 0, rows, d, eps, dx,
                                                  ld_dx);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_weightnorm_fwd(const float* g, const float* v, int K, int d, void* w_bf16, float* w_f32, cudaStream_t s) {
  APLA_CHECK(K > 0 && d > 0, "weightnorm_fwd: K=%d d=%d", K, d);
  APLA_CHECK(w_bf16 || w_f32, "weightnorm_fwd: no output");
  weightnorm_fwd_kernel<<<cdiv(K, 8), 256, 0, s>>>(g, v, K, d, reinterpret_cast<__nv_bfloat16*>(w_bf16), w_f32);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_weightnorm_bwd(const float* g, const float* v, const float* dW, int64_t ld_dw, int K, int d, float* dg, float* dv,
                       cudaStream_t s) {
  APLA_CHECK(K > 0 && d > 0, "weightnorm_bwd: K=%d d=%d", K, d);
  weightnorm_bwd_kernel<<<cdiv(K, 8), 256, 0, s>>>(g, v, dW, ld_dw, K, d, dg, dv);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_koleo_fwd(const float* xn, int groups, int n, int D, float eps, float w, int* nn, float* dist, float* row_loss,
                  cudaStream_t s) {
  // (n == 1, one image per process: the only candidate is the row itself, as in the reference -- argmax over a diagonal filled
  //  with -1 -- and the loss is the constant -log(1e-8 sqrt(D) + eps) with a zero gradient)
  APLA_CHECK(groups >= 1 && n >= 1 && D > 0, "koleo_fwd: groups=%d n=%d D=%d", groups, n, D);
  const size_t smem = size_t(D + 33 + 64) * sizeof(float);
  APLA_CHECK(smem <= 48 * 1024, "koleo_fwd: D=%d too wide for the row buffer", D);
  koleo_nn_kernel<<<dim3(n, groups), 256, smem, s>>>(xn, n, D, eps, w, nn, dist, row_loss);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_koleo_bwd(const float* x, const float* xn, int groups, int n, int D, float eps, float norm_eps, float w,
                  const int* nn, const float* dist, const float* gscale, float* dx, cudaStream_t s) {
  APLA_CHECK(groups >= 1 && n >= 1 && D > 0, "koleo_bwd: groups=%d n=%d D=%d", groups, n, D);
  const size_t smem = size_t(D + 33) * sizeof(float);
  APLA_CHECK(smem <= 48 * 1024, "koleo_bwd: D=%d too wide for the row buffer", D);
  koleo_bwd_kernel<<<dim3(n, groups), 256, smem, s>>>(x, xn, n, D, eps, norm_eps, w, nn, dist, gscale, dx);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

int ssl_ema(float* t, const float* sp, int64_t n, float m, cudaStream_t s) {
  APLA_CHECK(n >= 0, "ema: n=%lld", (long long)n);
  if (n == 0) return 0;
  const int grid = sm_count() * 8;
  if (n % 4 == 0 && aligned16(t) && aligned16(sp))
    ema_kernel_v4<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(t), reinterpret_cast<const float4*>(sp), n / 4, m);
  else
    ema_kernel<<<grid, 256, 0, s>>>(t, sp, n, m);
  count_launch();
  APLA_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// The objective of one self-supervised step on given head outputs, as ONE native launch sequence (no host code between
// the launches): teacher targets for the CLS rows and the masked-patch rows, the two centre statistics, and forward +
// backward of the three cross-entropy terms with the reference's scales (models.py:227-234, 374-433; two global crops).
//   s_scores [n_local*B + 2B + n_masked, K]  student head output: local CLS rows (crop-major), global CLS rows, masked rows
//   t_scores [2B + n_masked, K]              teacher head output: global CLS rows ALREADY swapped (models.py:244), masked rows
//   t_probs  same shape, workspace           softmax((t - centre) / teacher_temp), kept for the backward
//   row_ws   3 * student rows floats         per-row loss / log-sum-exp / target mass
//   losses[3]                                dino_local_crops_loss, dino_global_crops_loss, 2 * ibot_loss  (loss_dict scale,
//                                            i.e. before dino_weight / ibot_weight)
//   ds                                       d(dino_weight (l + g) + ibot_weight i) / d s_scores, times *gscale
//   dino_batch_sum[K], ibot_batch_mean[K]    inputs of apla_center_ema after the data-parallel all-reduce
// ------------------------------------------------------------------------------------------------
int ssl_objective(const float* s_scores, int64_t lds, const float* t_scores, int64_t ldt, float* t_probs, int64_t ldp,
                  const float* dino_center, const float* ibot_center, const float* masks_weight, int B, int n_local,
                  int n_masked, int K, float teacher_temp, float student_temp, float dino_weight, float ibot_weight,
                  float* row_ws, float* col_ws, int splits, void* ds, int64_t ldd, int ds_is_bf16, const float* gscale,
                  float* losses, float* dino_batch_sum, float* ibot_batch_mean, cudaStream_t s) {
  APLA_CHECK(B > 0 && n_local >= 0 && n_masked >= 0 && K > 0, "ssl_objective: bad sizes B=%d n_local=%d n_masked=%d K=%d",
             B, n_local, n_masked, K);
  APLA_CHECK(teacher_temp > 0.f && student_temp > 0.f, "ssl_objective: temperatures must be positive");
  const int n_g = 2 * B, n_l = n_local * B, rows = n_l + n_g + n_masked;
  const float inv_tt = 1.f / teacher_temp, inv_st = 1.f / student_temp;
  const int terms = 2 + (n_local * 2 > 1 ? n_local * 2 : 1);              // n_global_terms + n_local_terms
  const size_t esz = ds_is_bf16 ? 2 : 4;
  float *row_loss = row_ws, *lse = row_ws + rows, *mass = row_ws + 2 * size_t(rows);
  // teacher: targets and centre statistics
  if (int rc = ssl_softmax_center(t_scores, ldt, dino_center, inv_tt, n_g, K, t_probs, ldp, s)) return rc;
  if (int rc = ssl_softmax_center(t_scores + n_g * ldt, ldt, ibot_center, inv_tt, n_masked, K, t_probs + n_g * ldp, ldp, s))
    return rc;
  if (int rc = ssl_colsum_f32(t_scores, ldt, n_g, K, col_ws, splits, 1.f, dino_batch_sum, s)) return rc;
  if (int rc = ssl_colsum_f32(t_scores + n_g * ldt, ldt, n_masked, K, col_ws, splits,
                              n_masked > 0 ? 1.f / n_masked : 0.f, ibot_batch_mean, s))
    return rc;
  // the three terms: {first student row, rows, t0, t1, t_rows, per-row weights, loss_dict scale, gradient weight}
  struct Term { int r0, n; const float *t0, *t1; int t_rows; const float* w; float scale, weight; };
  const Term term[3] = {
      {0, n_l, t_probs, t_probs + int64_t(B) * ldp, B, nullptr, 1.f / (float(B) * terms), dino_weight},
      {n_l, n_g, t_probs, nullptr, n_g, nullptr, 2.f / (float(n_g) * terms), dino_weight},
      {n_l + n_g, n_masked, t_probs + int64_t(n_g) * ldp, nullptr, n_masked > 0 ? n_masked : 1, masks_weight,
       1.f / float(n_g), ibot_weight}};                                    // 2 (loss_scales) * 1/2 (ibot_loss_scale) / 2B
  for (int i = 0; i < 3; ++i) {
    const Term& t = term[i];
    const float* sp = s_scores + int64_t(t.r0) * lds;
    // one launch per term when the gradient is wanted (second read of the rows from L2); APLA_SSL_SPLIT_CE=1 keeps the
    // two-kernel form for A/B measurements
    static const bool split = [] { const char* e = getenv("APLA_SSL_SPLIT_CE"); return e && atoi(e) != 0; }();
    if (ds != nullptr && !split) {
      void* dp = reinterpret_cast<char*>(ds) + size_t(t.r0) * size_t(ldd) * esz;
      if (int rc = ssl_soft_ce_fwd_bwd(sp, lds, t.n, K, t.t0, t.t1, ldp, t.t_rows, t.w, t.scale, t.scale * t.weight, inv_st,
                                       gscale, row_loss + t.r0, dp, ldd, ds_is_bf16, s))
        return rc;
    } else {
      if (int rc = ssl_soft_ce_fwd(sp, lds, t.n, K, t.t0, t.t1, ldp, t.t_rows, t.w, t.scale, inv_st, row_loss + t.r0,
                                   lse + t.r0, mass + t.r0, s))
        return rc;
      if (ds != nullptr) {
        void* dp = reinterpret_cast<char*>(ds) + size_t(t.r0) * size_t(ldd) * esz;
        if (int rc = ssl_soft_ce_bwd(sp, lds, t.n, K, t.t0, t.t1, ldp, t.t_rows, t.w, t.scale * t.weight, inv_st,
                                     lse + t.r0, mass + t.r0, gscale, dp, ldd, ds_is_bf16, s))
          return rc;
      }
    }
    if (int rc = ssl_sum_f32(row_loss + t.r0, t.n, 1.f, losses + i, s)) return rc;
  }
  return 0;
}

}  // namespace apla
