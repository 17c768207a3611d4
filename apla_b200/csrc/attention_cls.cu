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

// Thread layout of both kernels: 8 consecutive lanes share one key -- lane `sub` owns the 16-byte piece (8 head dims) of
// its K / V row, so every 128-bit load / store instruction of a warp covers 4 whole 128-byte rows -- and the 32 groups of
// 8 lanes walk the keys j = group, group + 32, ...  Dot products are 8 FMAs + 3 shuffles.
__device__ __forceinline__ void unpack8(const uint4 x, float (&f)[8]) {
  f[0] = bf16_lo(x.x); f[1] = bf16_hi(x.x); f[2] = bf16_lo(x.y); f[3] = bf16_hi(x.y);
  f[4] = bf16_lo(x.z); f[5] = bf16_hi(x.z); f[6] = bf16_lo(x.w); f[7] = bf16_hi(x.w);
}
__device__ __forceinline__ float dot8_group(const float (&a)[8], const float (&b)[8]) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  return acc;
}

// out[b*N, h*64 ..] = softmax(scale * q_cls K^T) V ; lse[b*N, h] = log sum exp(scale * s)
__global__ void __launch_bounds__(kThreads)
attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N,
                    int H, float scale) {
  extern __shared__ float sm[];          // p[N] | o[8 warps][64] | red[8]
  float* p = sm;
  float* osum = p + N;
  float* red = osum + (kThreads / 32) * 64;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int D = H * 64;
  const size_t row0 = size_t(b) * N;
  const __nv_bfloat16* base = qkv + row0 * (3 * D) + h * 64;
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
  float q[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base) + sub), q);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] *= scale * LOG2E;                 // scores in the log2 domain
  float mx = -INFINITY;
  const int n_pass = (N + kThreads / 8 - 1) / (kThreads / 8);   // warp-uniform trip count: the dot products shuffle
  for (int it = 0; it < n_pass; ++it) {
    const int j = grp + it * (kThreads / 8);
    const bool ok = j < N;
    float k[8];
    unpack8(ok ? __ldg(reinterpret_cast<const uint4*>(base + size_t(j) * (3 * D) + D) + sub) : make_uint4(0u, 0u, 0u, 0u), k);
    const float s = dot8_group(q, k);
    if (ok) {
      if (sub == 0) p[j] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = block_reduce(mx, red, true);                                  // (its barriers also publish p[])
  float sum = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
  for (int j = grp; j < N; j += kThreads / 8) {                 // (no shuffles in this loop)
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + size_t(j) * (3 * D) + 2 * D) + sub), v);
    const float e = exp2f(p[j] - mx);
    if (sub == 0) sum += e;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(e, v[i], o[i]);
  }
  const float l = block_reduce(sum, red, false);
  // the 32 key groups that share a piece: the 4 groups of a warp by shuffles, the 8 warps in a fixed order (no atomics:
  // this row feeds the whole backward pass, run-to-run rounding noise here would reach every gradient)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = o[i];
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    if ((threadIdx.x & 31) < 8) osum[(threadIdx.x >> 5) * 64 + 8 * sub + i] = t;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += osum[w * 64 + threadIdx.x];
    out[row0 * D + h * 64 + threadIdx.x] = __float2bfloat16_rn(t / l);
  }
  if (threadIdx.x == 0) lse[row0 * H + h] = (mx + log2f(l)) * (1.0f / LOG2E);
}

// dqkv[b*N + j] for every key j of (b, h): dK_j = scale * dS_j q_cls, dV_j = P_j dO_cls, dQ_j = 0 (j > 0),
// dQ_cls = scale * sum_j dS_j K_j, with P_j = exp(scale s_j - lse), dS_j = P_j (dO_cls . V_j - dO_cls . O_cls)
__global__ void __launch_bounds__(kThreads)
attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    __nv_bfloat16* __restrict__ dqkv, int N, int H, float scale) {
  __shared__ float dqs[kThreads / 32][64];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int D = H * 64;
  const size_t row0 = size_t(b) * N;
  const __nv_bfloat16* base = qkv + row0 * (3 * D) + h * 64;
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
  float q[8], g[8], oc[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base) + sub), q);
  unpack8(__ldg(reinterpret_cast<const uint4*>(dout + row0 * D + h * 64) + sub), g);
  unpack8(__ldg(reinterpret_cast<const uint4*>(out + row0 * D + h * 64) + sub), oc);
  const float delta = dot8_group(g, oc);                              // dO_cls . O_cls
  const float l2 = __ldg(lse + row0 * H + h) * LOG2E;
  float dq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dq[i] = 0.f;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const int n_pass = (N + kThreads / 8 - 1) / (kThreads / 8);   // warp-uniform trip count: the dot products shuffle
  for (int it = 0; it < n_pass; ++it) {
    const int j = grp + it * (kThreads / 8);
    const bool ok = j < N;
    const __nv_bfloat16* kr = base + size_t(ok ? j : 0) * (3 * D) + D;
    float k[8], v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(kr) + sub), k);
    unpack8(__ldg(reinterpret_cast<const uint4*>(kr + D) + sub), v);
    const float s = dot8_group(q, k);
    const float dp = dot8_group(g, v);
    if (!ok) continue;
    const float pj = exp2f(s * (scale * LOG2E) - l2);
    const float ds = pj * (dp - delta) * scale;
    __nv_bfloat16* o = dqkv + (row0 + j) * (3 * D) + h * 64;
    reinterpret_cast<uint4*>(o + D)[sub] = make_uint4(pack_bf16(ds * q[0], ds * q[1]), pack_bf16(ds * q[2], ds * q[3]),
                                                      pack_bf16(ds * q[4], ds * q[5]), pack_bf16(ds * q[6], ds * q[7]));
    reinterpret_cast<uint4*>(o + 2 * D)[sub] = make_uint4(pack_bf16(pj * g[0], pj * g[1]), pack_bf16(pj * g[2], pj * g[3]),
                                                          pack_bf16(pj * g[4], pj * g[5]), pack_bf16(pj * g[6], pj * g[7]));
    if (j > 0) reinterpret_cast<uint4*>(o)[sub] = zero;
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i] = fmaf(ds, k[i], dq[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = dq[i];
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    if ((threadIdx.x & 31) < 8) dqs[threadIdx.x >> 5][8 * sub + i] = t;   // fixed-order sum below: deterministic
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += dqs[w][threadIdx.x];
    dqkv[row0 * (3 * D) + h * 64 + threadIdx.x] = __float2bfloat16_rn(t);
  }
}

}  // namespace

int attn_cls_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, cudaStream_t stream) {
  APLA_CHECK(B > 0 && N > 0 && H > 0 && N <= 8192, "attn_cls_fwd: bad shape B=%d N=%d H=%d", B, N, H);
  const size_t smem = (size_t(N) + (kThreads / 32) * 64 + 8) * sizeof(float);
  attn_cls_fwd_kernel<<<B * H, kThreads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                         reinterpret_cast<__nv_bfloat16*>(out), lse, N, H, scale);
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
