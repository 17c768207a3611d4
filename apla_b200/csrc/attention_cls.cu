// Attention of the LAST block, for the CLS query only.
//
// The classifier reads norm(x)[:, 0] (src/utils/transformers/vit.py:417-419, src/defaults/models.py:87): of the last
// block's attention output only the CLS row of every image is ever used, and on the way back only that row carries a
// gradient.  Forward is then one query row against all keys per (image, head); backward produces dK / dV of every key
// from that single row, dQ of the CLS row, and exact zeros for the dQ of all other rows -- what the dense kernels
// (appla_attn.py:56-62 semantics) would compute for the same dO, at 1/N of the work.  33 k FMA per (image, head): CUDA
// cores, fp32, one block per (image, head); the kernels are bound by reading K / V once and writing dqkv once.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace {

constexpr int kThreads = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  // red: [kThreads / 32] floats of shared memory; every thread gets the result
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// dot of a 64-element bf16 row (128 contiguous bytes in global memory) with 64 floats in shared memory
__device__ __forceinline__ float dot64(const __nv_bfloat16* __restrict__ row, const float* __restrict__ v) {
  const uint4* r4 = reinterpret_cast<const uint4*>(row);
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 x = __ldg(r4 + u);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      acc = fmaf(bf16_lo(w[t]), v[8 * u + 2 * t], acc);
      acc = fmaf(bf16_hi(w[t]), v[8 * u + 2 * t + 1], acc);
    }
  }
  return acc;
}

// out[b*N, h*64 ..] = softmax(scale * q_cls K^T) V ; lse[b*N, h] = log sum exp(scale * s)
__global__ void __launch_bounds__(kThreads)
attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N,
                    int pn, int H, float scale) {
  extern __shared__ float sm[];          // p[pn] (pn = max(N, 512): reused for the [8][64] partial outputs) | q[64] | red[8]
  float* p = sm;
  float* q = p + pn;
  float* red = q + 64;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int D = H * 64;
  const size_t row0 = size_t(b) * N;
  const __nv_bfloat16* base = qkv + row0 * (3 * D) + h * 64;
  if (threadIdx.x < 64) q[threadIdx.x] = __bfloat162float(base[threadIdx.x]) * (scale * LOG2E);
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += kThreads) {
    const float s = dot64(base + size_t(j) * (3 * D) + D, q);      // log2-domain score
    p[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int j = threadIdx.x; j < N; j += kThreads) {
    const float e = exp2f(p[j] - mx);
    p[j] = e;
    sum += e;
  }
  const float l = block_reduce(sum, red, false);                    // (also orders the p[] writes before the reads below)
  // O[2 dp .. 2 dp + 1] over the keys part, part + 8, ...: 32 column pairs x 8 key slices
  const int dp = threadIdx.x & 31, slice = threadIdx.x >> 5;
  float o0 = 0.f, o1 = 0.f;
  for (int j = slice; j < N; j += kThreads / 32) {
    const uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(base + size_t(j) * (3 * D) + 2 * D) + dp);
    o0 = fmaf(p[j], bf16_lo(vv), o0);
    o1 = fmaf(p[j], bf16_hi(vv), o1);
  }
  __syncthreads();
  float* acc = p;                         // reuse: [8][64]
  __syncthreads();
  acc[slice * 64 + 2 * dp] = o0;
  acc[slice * 64 + 2 * dp + 1] = o1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float o = 0.f;
#pragma unroll
    for (int s8 = 0; s8 < kThreads / 32; ++s8) o += acc[s8 * 64 + threadIdx.x];
    out[row0 * D + h * 64 + threadIdx.x] = __float2bfloat16_rn(o / l);
  }
  if (threadIdx.x == 0) lse[row0 * H + h] = (mx + log2f(l)) * (1.0f / LOG2E);
}

// dqkv[b*N + j] for every key j of (b, h): dK_j = scale * dS_j q_cls, dV_j = P_j dO_cls, dQ_j = 0 (j > 0),
// dQ_cls = scale * sum_j dS_j K_j, with P_j = exp(scale s_j - lse), dS_j = P_j (dO_cls . V_j - dO_cls . O_cls)
__global__ void __launch_bounds__(kThreads)
attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    __nv_bfloat16* __restrict__ dqkv, int N, int H, float scale) {
  __shared__ float q[64], dO[64], dq[64], red[kThreads / 32];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int D = H * 64;
  const size_t row0 = size_t(b) * N;
  const __nv_bfloat16* base = qkv + row0 * (3 * D) + h * 64;
  float dl = 0.f;
  if (threadIdx.x < 64) {
    q[threadIdx.x] = __bfloat162float(base[threadIdx.x]);
    const float g = __bfloat162float(dout[row0 * D + h * 64 + threadIdx.x]);
    dO[threadIdx.x] = g;
    dq[threadIdx.x] = 0.f;
    dl = g * __bfloat162float(out[row0 * D + h * 64 + threadIdx.x]);
  }
  const float delta = block_reduce(dl, red, false);
  const float l2 = lse[row0 * H + h] * LOG2E;
  float dqa[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) dqa[d] = 0.f;
  for (int j = threadIdx.x; j < N; j += kThreads) {
    const __nv_bfloat16* kr = base + size_t(j) * (3 * D) + D;
    const float s = dot64(kr, q);
    const float pj = exp2f(s * (scale * LOG2E) - l2);
    const float dp = dot64(kr + D, dO);
    const float ds = pj * (dp - delta) * scale;
    __nv_bfloat16* o = dqkv + (row0 + j) * (3 * D) + h * 64;
    const uint4* k4 = reinterpret_cast<const uint4*>(kr);
    uint4* dq4 = reinterpret_cast<uint4*>(o);
    uint4* dk4 = reinterpret_cast<uint4*>(o + D);
    uint4* dv4 = reinterpret_cast<uint4*>(o + 2 * D);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint4 kk = __ldg(k4 + u);
      const uint32_t kw[4] = {kk.x, kk.y, kk.z, kk.w};
      uint32_t wk[4], wv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int d = 8 * u + 2 * t;
        dqa[d] = fmaf(ds, bf16_lo(kw[t]), dqa[d]);
        dqa[d + 1] = fmaf(ds, bf16_hi(kw[t]), dqa[d + 1]);
        wk[t] = pack_bf16(ds * q[d], ds * q[d + 1]);
        wv[t] = pack_bf16(pj * dO[d], pj * dO[d + 1]);
      }
      dk4[u] = make_uint4(wk[0], wk[1], wk[2], wk[3]);
      dv4[u] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      if (j > 0) dq4[u] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // dQ of the CLS row: warp-reduce every component, one shared atomic per warp and component
#pragma unroll
  for (int d = 0; d < 64; ++d) {
    float v = dqa[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&dq[d], v);
  }
  __syncthreads();
  if (threadIdx.x < 64) dqkv[row0 * (3 * D) + h * 64 + threadIdx.x] = __float2bfloat16_rn(dq[threadIdx.x]);
}

}  // namespace

int attn_cls_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, cudaStream_t stream) {
  APLA_CHECK(B > 0 && N > 0 && H > 0 && N <= 8192, "attn_cls_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  const int pn = N < 512 ? 512 : N;
  const size_t smem = (size_t(pn) + 64 + 8) * sizeof(float);
  attn_cls_fwd_kernel<<<B * H, kThreads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<__nv_bfloat16*>(out), lse, N, pn, H, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int attn_cls_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int N, int H,
                 float scale, cudaStream_t stream) {
  APLA_CHECK(B > 0 && N > 0 && H > 0, "attn_cls_bwd: bad shape B=%d N=%d H=%d", B, N, H);
  attn_cls_bwd_kernel<<<B * H, kThreads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                      reinterpret_cast<const __nv_bfloat16*>(out),
                                                      reinterpret_cast<const __nv_bfloat16*>(dout), lse,
                                                      reinterpret_cast<__nv_bfloat16*>(dqkv), N, H, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
