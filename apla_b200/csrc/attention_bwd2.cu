// Second-generation fused tcgen05 / TMEM attention backward for short sequences (N <= 257 tokens, head_dim 64): the
// ViT-at-224-px case (257 / 197 / 50 tokens).  One launch produces dQ, dK and dV; the scores are recomputed once.
// Replaces the backward autograd derives for src/apla/appla_attn.py:56-60 (softmax(q k^T scale) v).
//
// What changed against attention_fused.cu (kept as the fallback for 258..272 tokens):
//   * The ODD TOKEN is off the tensor path.  N = 257 = 4 * 64 + 1: as a 5th query chunk and a 3rd key tile the one extra
//     token cost 7 of 15 (key tile, query chunk) pairs, five of them at almost full MMA price (M is always 128).  When
//     n % 64 == 1 the last token o is handled by a helper warpgroup on CUDA cores (fp32):
//         as a key   : p_qo, dS_qo for every query q  ->  dK[o], dV[o] (sums over q) and a rank-1 update of dQ
//         as a query : p_ok, dS_ok for every key k    ->  dQ[o] (sum over k)       and rank-1 updates of dK, dV
//     The rank-1 updates are added by the epilogue warps when they take the accumulators out of TMEM.  The tensor path
//     then sees 256 tokens: 2 key tiles x 4 query chunks, every pair full.
//   * P^T is written over the (already consumed) dP^T columns instead of over S^T, so the S^T slot is free as soon as
//     the warpgroup has the scores in registers: S^T of chunk c+2 is issued while chunk c is still being processed and
//     the exponentials of a chunk no longer wait for dV of the chunk two before it.  dP^T is double-buffered (the dQ
//     accumulator needs 128 columns instead of 192 now), and S^T / dP^T complete separate barriers so the exponentials
//     start while dP^T is still in the tensor pipe.
//   * The MMA issuer is event-driven (polls "scores consumed" and "P / dS written") instead of walking a fixed order.
//
// One persistent CTA per SM walks over (sequence, head) groups.  A group's Q and dO rows stay in shared memory (64-row
// blocks, released one by one during the group's last key tile); 128-key tiles of K and V pass through a two-slot ring.
// Orientation: TMEM lanes = keys.  Per (key tile, 64-query chunk):
//     S    : S^T = K_tile . Q_chunk^T                                           (smem x smem -> TMEM, fp32)
//     dP   : dP^T = V_tile . dO_chunk^T
//     WG   : P^T = exp2(S^T*scale*log2e - lse_q),  dS^T = P^T o (dP^T - delta_q)
//            P^T  -> TMEM (bf16, over the first half of dP^T)  : A operand of the dV MMA
//            dS^T -> shared-memory panel [128 keys x 64 q]     : A operand of the dK MMA (K-major view) AND of the
//                                                                dQ MMA (the same bytes viewed MN-major = transposed)
//     ACC  : dV += P^T . dO_chunk,   dK += dS^T . Q_chunk,   and once per 128 queries  dQ_t += dS_t . K_tile
// No accumulator ever leaves the CTA: no atomics, deterministic.
//
// Warp roles (20 warps = 5 warpgroups, registers re-balanced with setmaxnreg):
//   WG0: 0 = TMA producer, 1 = MMA issuer (one thread), 2 = per-query lse/delta staging, 3 idle
//   WG1 / WG2 (warps 4-7 / 8-11): the two compute warpgroups taking alternate chunks
//   WG3 (12-15): epilogue (TMEM accumulators + rank-1 terms -> bf16 -> global)
//   WG4 (16-19): odd-token helper
// TMEM (512 columns): dQ 2 x 64 | dK 64 | dV 64 | S^T ring 2 x 64 | dP^T / P^T ring 2 x 64.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace afb {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int NT_MAX = 256;                  // tokens on the tensor path
constexpr int CW = 64;                       // queries per chunk
constexpr int MAX_CHUNKS = NT_MAX / CW;      // 4
constexpr int kThreads = 640;
constexpr uint32_t RES_BYTES = NT_MAX * 128;             // one resident operand (Q or dO)
constexpr uint32_t TILE_BYTES = 128 * 128;               // one 128-row K or V tile / one dS^T panel
constexpr uint32_t OFF_Q = 0, OFF_DO = RES_BYTES, OFF_KV = 2 * RES_BYTES;   // KV: [2 slots][K, V]
constexpr uint32_t OFF_DS = OFF_KV + 4 * TILE_BYTES;                          // [2 pairs][2 panels]
constexpr uint32_t OFF_STAT = OFF_DS + 4 * TILE_BYTES;                        // lse2[256], delta[256]
constexpr uint32_t OFF_VEC = OFF_STAT + 2 * NT_MAX * 4;   // [2 group parities][Q, K, V, dO of the odd token][64] floats
constexpr uint32_t OFF_COL = OFF_VEC + 2 * 4 * 64 * 4;    // [2 group parities][p_qo, dS_qo][256] floats
constexpr uint32_t OFF_ROW = OFF_COL + 2 * 2 * NT_MAX * 4;  // [2 kv slots][p_ok, dS_ok][128] floats
constexpr uint32_t OFF_PART = OFF_ROW + 2 * 2 * 128 * 4;  // [4 parts][dQ_o, dK_o, dV_o][64] floats
constexpr uint32_t OFF_BAR = OFF_PART + 4 * 3 * 64 * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr uint32_t TM_DQ = 0, TM_DK = 128, TM_DV = 192, TM_S = 256, TM_DP = 384, TM_COLS = 512;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// Register budget per warpgroup.  setmaxnreg moves registers inside the pool the CTA was LAUNCHED with -- 640 threads x 96
// registers (the most a 640-thread CTA can be compiled for), not the SM's 64 K -- so the five budgets must sum to 5 x 96.
// (The first version summed to 512: the second compute warpgroup's setmaxnreg.inc waited forever.)
constexpr int REG_LAUNCH = 96;
constexpr int REG_WG0 = 56, REG_COMPUTE = 136, REG_EPI = 80, REG_HELP = 72;
static_assert(REG_WG0 + 2 * REG_COMPUTE + REG_EPI + REG_HELP <= 5 * REG_LAUNCH, "setmaxnreg budget exceeds the CTA's pool");

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Shared-memory matrix descriptors (128B swizzle, 8-row groups 1024 B apart) split into 32-bit halves so that the
// per-MMA address arithmetic is a single 32-bit add:  lo = (addr >> 4) | LBO>>4 << 16,  hi = SBO>>4 | version | layout.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_LO_K = (16u >> 4) << 16;        // K-major operand: +2 per 16-element (32 B) k-step
constexpr uint32_t DESC_LO_MN = (16384u >> 4) << 16;    // MN-major operand: +128 per 16-row (2048 B) k-step; 64-wide
                                                        // MN boxes 16 KB apart (the two dS^T panels of a query tile)

__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}

// -DAPLA_AFB_DEBUG: every wait is tagged and a wait that never completes prints where it hung before trapping
#ifdef APLA_AFB_DEBUG
__device__ __forceinline__ void wait_tag(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 16)) {
      if ((threadIdx.x & 31) == 0)
        printf("afb HANG cta %d warp %d tag %d parity %u\n", blockIdx.x, threadIdx.x >> 5, tag, parity);
      __trap();
    }
  }
}
#define WAIT(bar, parity, tag) wait_tag(bar, parity, tag)
#elif defined(APLA_AFB_PROF)
// -DAPLA_AFB_PROF: CTA 0 prints, per warp, the cycles spent blocked at each tagged wait (development aid)
#define WAIT(bar, parity, tag)                          \
  do {                                                  \
    const long long _t0 = clock64();                    \
    mbar_wait(bar, parity);                             \
    prof[tag] += uint32_t(clock64() - _t0);             \
  } while (0)
#else
#define WAIT(bar, parity, tag) mbar_wait(bar, parity)
#endif

struct Problem {
  const int* cu;
  int n_fixed, H, G;
  bool no_odd;     // ablation: drop the odd token's terms
};

// Deterministic walk over this CTA's (group, key tile, query chunk) sequence; every warp role runs its own copy.
//   gi = groups done, ts = key tiles done, c = chunks done, tqb = query tiles (chunk pairs) done before this key tile
//   n = tokens of the sequence, nt = tokens on the tensor path, odd = 1 when token nt is the helper's
struct Walk {
  int g, gi, ts, c, tqb, jt, j;
  int row_start, n, nt, odd, h, ntiles, nchunks;
  uint32_t qpar;   // bit j: parity of the number of groups that have used Q/dO block j (groups differ in length)
  __device__ __forceinline__ void load(const Problem& p) {
    while (g < p.G) {
      const int b = g / p.H;
      h = g - b * p.H;
      if (p.cu) {
        row_start = p.cu[b];
        n = p.cu[b + 1] - row_start;
      } else {
        row_start = b * p.n_fixed;
        n = p.n_fixed;
      }
      if (n > 0) break;
      g += gridDim.x;
    }
    row_start = __shfl_sync(0xffffffffu, row_start, 0);
    n = __shfl_sync(0xffffffffu, n, 0);
    odd = (n > CW && (n & (CW - 1)) == 1) ? 1 : 0;
    nt = n - odd;
    if (p.no_odd) odd = 0;
    ntiles = (nt + 127) >> 7;
    nchunks = (nt + CW - 1) / CW;
  }
  __device__ __forceinline__ void init(const Problem& p) {
    g = blockIdx.x;
    gi = ts = c = tqb = jt = j = 0;
    row_start = n = nt = odd = h = 0;
    ntiles = nchunks = 1;
    qpar = 0;
    load(p);
  }
  __device__ __forceinline__ bool done(const Problem& p) const { return g >= p.G; }
  __device__ __forceinline__ void next_group(const Problem& p) {
    jt = 0;
    ++gi;
    qpar ^= (1u << nchunks) - 1u;
    g += gridDim.x;
    load(p);
  }
  __device__ __forceinline__ void next_tile(const Problem& p) {
    j = 0;
    ++ts;
    tqb += (nchunks + 1) >> 1;
    if (++jt == ntiles) next_group(p);
  }
  __device__ __forceinline__ void next_chunk(const Problem& p) {
    ++c;
    if (++j == nchunks) next_tile(p);
  }
  __device__ __forceinline__ int valid_cols() const { return min(CW, nt - j * CW); }       // queries in this chunk
  __device__ __forceinline__ int valid_keys() const { return min(128, nt - jt * 128); }    // keys in this tile
  __device__ __forceinline__ int tq() const { return tqb + (j >> 1); }                      // query-tile sequence number
  __device__ __forceinline__ bool last_chunk() const { return j == nchunks - 1; }
  __device__ __forceinline__ bool last_tile() const { return jt == ntiles - 1; }
};

struct Maps {
  CUtensorMap qkv64, qkv16, do64, do16;   // 64- and 16-row boxes (64 bf16 columns) of the packed qkv matrix / dO
};

// Two dot products at once -- rows ra . va and rb . vb, 128-byte (64 bf16) rows of 128B-swizzled tiles against 64-float
// vectors (broadcast reads) -- so that four independent FMA chains and six loads are in flight per step: the helper is
// latency-bound (one warp per scheduler), a single dot product per loop ran at ~3 000 cycles.
__device__ __forceinline__ void dot2_row64(uint32_t ra, uint32_t rb, int row_in_tile, uint32_t va, uint32_t vb, float& da,
                                           float& db) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  const uint32_t sw = uint32_t(row_in_tile & 7);
#pragma unroll 4
  for (int u = 0; u < 8; ++u) {
    const uint32_t off = (uint32_t(u) ^ sw) << 4;
    const uint4 xa = lds_u4(ra + off);
    const uint4 xb = lds_u4(rb + off);
    const float4 p0 = lds_f4(va + 32 * u);
    const float4 p1 = lds_f4(va + 32 * u + 16);
    const float4 q0 = lds_f4(vb + 32 * u);
    const float4 q1 = lds_f4(vb + 32 * u + 16);
    a0 = fmaf(bf16_lo(xa.x), p0.x, a0);
    a1 = fmaf(bf16_hi(xa.x), p0.y, a1);
    b0 = fmaf(bf16_lo(xb.x), q0.x, b0);
    b1 = fmaf(bf16_hi(xb.x), q0.y, b1);
    a0 = fmaf(bf16_lo(xa.y), p0.z, a0);
    a1 = fmaf(bf16_hi(xa.y), p0.w, a1);
    b0 = fmaf(bf16_lo(xb.y), q0.z, b0);
    b1 = fmaf(bf16_hi(xb.y), q0.w, b1);
    a0 = fmaf(bf16_lo(xa.z), p1.x, a0);
    a1 = fmaf(bf16_hi(xa.z), p1.y, a1);
    b0 = fmaf(bf16_lo(xb.z), q1.x, b0);
    b1 = fmaf(bf16_hi(xb.z), q1.y, b1);
    a0 = fmaf(bf16_lo(xa.w), p1.z, a0);
    a1 = fmaf(bf16_hi(xa.w), p1.w, a1);
    b0 = fmaf(bf16_lo(xb.w), q1.z, b0);
    b1 = fmaf(bf16_hi(xb.w), q1.w, b1);
  }
  da = a0 + a1;
  db = b0 + b1;
}
__device__ __forceinline__ float dot_vec64(uint32_t a, uint32_t b) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
  for (int u = 0; u < 16; ++u) {
    const float4 x = lds_f4(a + 16 * u);
    const float4 y = lds_f4(b + 16 * u);
    s0 = fmaf(x.x, y.x, s0);
    s1 = fmaf(x.y, y.y, s1);
    s0 = fmaf(x.z, y.z, s0);
    s1 = fmaf(x.w, y.w, s1);
  }
  return s0 + s1;
}

// one 64-column fp32 accumulator row of this thread -> (acc + coef * vec[c]) * mul -> bf16 -> 128 contiguous bytes
__device__ __forceinline__ void drain_row64(uint32_t taddr, float mul, bool rank1, float coef, uint32_t vec,
                                            __nv_bfloat16* dst, bool store) {
  uint32_t pk[32];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t ov[32];
    tmem_ld_32x32(taddr + c * 32, ov);
    tmem_ld_wait();
    if (rank1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = lds_f4(vec + (c * 32 + 4 * i) * 4);
        ov[4 * i + 0] = __float_as_uint(fmaf(coef, v.x, __uint_as_float(ov[4 * i + 0])));
        ov[4 * i + 1] = __float_as_uint(fmaf(coef, v.y, __uint_as_float(ov[4 * i + 1])));
        ov[4 * i + 2] = __float_as_uint(fmaf(coef, v.z, __uint_as_float(ov[4 * i + 2])));
        ov[4 * i + 3] = __float_as_uint(fmaf(coef, v.w, __uint_as_float(ov[4 * i + 3])));
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      pk[c * 16 + i] = pack_bf16(__uint_as_float(ov[2 * i]) * mul, __uint_as_float(ov[2 * i + 1]) * mul);
  }
  if (store) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd2_kernel(const __grid_constant__ Maps maps, const __nv_bfloat16* __restrict__ qkv,
                 const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                 __nv_bfloat16* __restrict__ dqkv, const int* __restrict__ cu_seqlens, int n_fixed, int H, int G,
                 float scale, int ablate) {
#ifndef APLA_AFB_ABLATE
  ablate = 0;     // (timing experiments only: -DAPLA_AFB_ABLATE + APLA_AFB_ABLATE=<mask>; results are wrong with any bit set)
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* qdo_full = bars;          // [4]  Q/dO block + its statistics landed            (TMA tx + stats warp)
  uint64_t* qdo_empty = bars + 4;     // [4]  block no longer read                           (commit + helper)
  uint64_t* kv_full = bars + 8;       // [2]
  uint64_t* kv_empty = bars + 10;     // [2]                                                 (commit + helper)
  uint64_t* s_full = bars + 12;       // [2]  S^T of a chunk is in TMEM                      (commit)
  uint64_t* s_free = bars + 14;       // [2]  S^T has been read into registers               (4 warps)
  uint64_t* dp_full = bars + 16;      // [2]  dP^T of a chunk is in TMEM                     (commit)
  uint64_t* pd_full = bars + 18;      // [2]  P^T in TMEM and dS^T panel in smem are written (4 warps)
  uint64_t* ds_empty = bars + 20;     // [2]  panel pair no longer read by any MMA           (commit)
  uint64_t* acc_full = bars + 22;     //      dK/dV of a key tile complete                   (commit)
  uint64_t* acc_empty = bars + 23;    //      ... and taken out of TMEM                      (4 warps)
  uint64_t* dq_full = bars + 24;      //      dQ of a group complete                         (commit)
  uint64_t* dq_empty = bars + 25;     //      ... and taken out of TMEM                      (4 warps)
  uint64_t* hrow_full = bars + 26;    // [2]  odd-token row coefficients of a key tile       (helper)
  uint64_t* hrow_empty = bars + 28;   // [2]  ... consumed                                   (4 epilogue warps)
  uint64_t* hcol_full = bars + 30;    // [2]  odd-token vectors + column coefficients        (helper)
  uint64_t* hcol_empty = bars + 32;   // [2]  ... consumed                                   (4 epilogue warps)
  uint64_t* pk_full = bars + 34;      // [4]  same event as pd_full for the dK / dQ issuer, which may trail the warpgroups
                                      //      by up to three chunks (a two-slot parity barrier would alias)  (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const Problem prob{cu_seqlens, n_fixed, H, G, (ablate & 2) != 0};
  const int D = H * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.qkv64);
    tma_prefetch_desc(&maps.qkv16);
    tma_prefetch_desc(&maps.do64);
    tma_prefetch_desc(&maps.do16);
    for (int i = 0; i < MAX_CHUNKS; ++i) {
      mbar_init(&qdo_full[i], 2);
      mbar_init(&qdo_empty[i], 4);   // three MMA issuers + the helper
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 4);    // three MMA issuers + the helper
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&dp_full[i], 1);
      mbar_init(&pd_full[i], 4);
      mbar_init(&ds_empty[i], 1);
      mbar_init(&hrow_full[i], 1);
      mbar_init(&hrow_empty[i], 4);
      mbar_init(&hcol_full[i], 1);
      mbar_init(&hcol_empty[i], 4);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&pk_full[i], 4);
    mbar_init(acc_full, 2);          // dV issuer + dK issuer
    mbar_init(acc_empty, 4);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, TM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
#ifdef APLA_AFB_PROF
  uint32_t prof[20];
#pragma unroll
  for (int i = 0; i < 20; ++i) prof[i] = 0;
  const long long prof_t0 = clock64();
#endif

#ifdef APLA_AFB_TRACE
  __shared__ long long tr_buf[6][200];
  __shared__ int tr_n[6];
  __shared__ long long tr_t0s;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) tr_n[i] = 0;
    tr_t0s = clock64();
  }
  __syncthreads();
  const long long tr_t0 = tr_t0s;
#define TR(role, ev, arg)                                                                              \
  do {                                                                                                 \
    if (blockIdx.x == 0 && lane == 0 && tr_n[role] < 200)                                              \
      tr_buf[role][tr_n[role]++] = ((clock64() - tr_t0) << 16) | ((long long)(ev) << 8) | ((arg) & 255); \
  } while (0)
#else
#define TR(role, ev, arg)
#endif
  // 32-bit shared addresses (every shared-memory access below is an explicit ld.shared / st.shared, see ptx.cuh)
  const uint32_t sb = smem_u32(smem);
  const uint32_t a_sl = sb + OFF_STAT, a_sd = a_sl + NT_MAX * 4;   // lse * log2 e, delta of the tensor-path queries
  const uint32_t a_vec = sb + OFF_VEC;                             // [par][4][64] floats: 0 Q_o, 1 K_o, 2 V_o, 3 dO_o
  const uint32_t a_col = sb + OFF_COL;                             // [par][2][256] floats: 0 p_qo, 1 dS_qo
  const uint32_t a_row = sb + OFF_ROW;                             // [slot][2][128] floats: 0 p_ok, 1 dS_ok

  if (warp < 4) {
    reg_dealloc<REG_WG0>();
    const bool leader = elect_one();
    const uint32_t q_lo = smem_u32(smem + OFF_Q) >> 4, do_lo = smem_u32(smem + OFF_DO) >> 4;
    const uint32_t kv_lo = smem_u32(smem + OFF_KV) >> 4, ds_lo = smem_u32(smem + OFF_DS) >> 4;
    if (warp == 0) {
      // ---------------------------------------------------------------------------------- producer + statistics
      // Load order per group: key tile 0, the Q/dO blocks, the remaining key tiles -- the order in which the consumers
      // free the buffers, so no wait here can depend on a load that is issued later.  The same warp stages lse (times
      // log2 e) and delta of every tensor-path query: gathered into registers one group ahead, dropped into a block's
      // slot as soon as the block is released; +inf / 0 past the end masks the padded columns (P = dS = 0).
      auto load_rows = [&](uint8_t* dst, const CUtensorMap* m64, const CUtensorMap* m16, uint64_t* bar, int col, int row,
                           int rows) {
        int r = 0;
        for (; r + 64 <= rows; r += 64) tma_load_2d(dst + r * 128, m64, bar, col, row + r);
        for (; r < rows; r += 16) tma_load_2d(dst + r * 128, m16, bar, col, row + r);
      };
      Walk k;
      k.init(prob);
      float vl[NT_MAX / 32], vd[NT_MAX / 32];
      auto fetch_stats = [&]() {
#pragma unroll
        for (int u = 0; u < NT_MAX / 32; ++u) {
          const int i = u * 32 + lane;
          const bool ok = i < k.nt;
          const size_t idx = size_t(k.row_start + (ok ? i : 0)) * H + k.h;
          vl[u] = ok ? __ldg(lse + idx) * LOG2E : INFINITY;
          vd[u] = ok ? __ldg(delta + idx) : 0.f;
        }
      };
      if (!k.done(prob)) fetch_stats();
      int ts = 0;   // key tiles loaded so far
      while (!k.done(prob)) {
        const int np = (k.nt + 15) & ~15;
        auto load_kv = [&](int jt) {
          const int slot = ts & 1;
          if (lane == 0) {
            const int rows = min(128, np - jt * 128);
            WAIT(&kv_empty[slot], ((ts >> 1) & 1) ^ 1, 1);
            TR(4, 1, ts);
            mbar_arrive_expect_tx(&kv_full[slot], 2u * rows * 128u);
            uint8_t* dk = smem + OFF_KV + slot * 2 * TILE_BYTES;
            load_rows(dk, &maps.qkv64, &maps.qkv16, &kv_full[slot], D + k.h * 64, k.row_start + jt * 128, rows);
            load_rows(dk + TILE_BYTES, &maps.qkv64, &maps.qkv16, &kv_full[slot], 2 * D + k.h * 64,
                      k.row_start + jt * 128, rows);
          }
          ++ts;
        };
        load_kv(0);
        for (int j = 0; j < k.nchunks; ++j) {
          const int rows = min(CW, np - j * CW);
          WAIT(&qdo_empty[j], ((k.qpar >> j) & 1) ^ 1, 2);
          if (lane == 0) {
            TR(4, 2, k.gi * 4 + j);
            mbar_arrive_expect_tx(&qdo_full[j], 2u * rows * 128u);
            load_rows(smem + OFF_Q + j * CW * 128, &maps.qkv64, &maps.qkv16, &qdo_full[j], k.h * 64,
                      k.row_start + j * CW, rows);
            load_rows(smem + OFF_DO + j * CW * 128, &maps.do64, &maps.do16, &qdo_full[j], k.h * 64, k.row_start + j * CW,
                      rows);
          }
#pragma unroll
          for (int u = 0; u < NT_MAX / 32; ++u) {
            if (u >> 1 == j) {
              sts_f32(a_sl + (u * 32 + lane) * 4, vl[u]);
              sts_f32(a_sd + (u * 32 + lane) * 4, vd[u]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&qdo_full[j]);
        }
        for (int jt = 1; jt < k.ntiles; ++jt) load_kv(jt);
        k.next_group(prob);
        if (!k.done(prob)) fetch_stats();
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------------------------------------- S^T issuer
      // Three single-purpose MMA issuers (one elected thread each) instead of one event loop: the bookkeeping of one warp
      // walking three cursors cost ~2 400 cycles per chunk -- more than the chunk's arithmetic.  Ordering between
      // MMAs of different issuers is always through a completed barrier (s_free, pd_full, acc_empty, ...).
      // S^T(c) needs its slot read out (s_free of chunk c - 2) and the operands landed.
      const uint32_t idesc_dummy = 0;
      (void)idesc_dummy;
      Walk w;
      w.init(prob);
      while (!w.done(prob)) {
        const int slot = w.ts & 1;
        if (w.c >= 2) WAIT(&s_free[w.c & 1], ((w.c - 2) >> 1) & 1, 3);
        if (w.jt == 0) WAIT(&qdo_full[w.j], (w.qpar >> w.j) & 1, 4);
        if (w.j == 0) WAIT(&kv_full[slot], (w.ts >> 1) & 1, 5);
        tc_fence_after();
        const int n_mma = (w.valid_cols() + 15) & ~15;
        const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
        const uint32_t a_k = DESC_LO_K + kv_lo + slot * (2 * TILE_BYTES >> 4);
        const uint32_t b_q = DESC_LO_K + q_lo + w.j * (CW * 128 >> 4);
        const uint32_t t_s = tmem + TM_S + (w.c & 1) * CW;
        TR(0, 1, w.c);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(t_s, a_k + 2 * kk, b_q + 2 * kk, idesc_s, kk > 0);
          umma_commit(&s_full[w.c & 1]);
          if (w.last_tile()) umma_commit(&qdo_empty[w.j]);     // this issuer's last read of Q block j
          if (w.last_chunk()) umma_commit(&kv_empty[slot]);    // ... and of the K tile
        }
        __syncwarp();
        w.next_chunk(prob);
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------------------------------- dP^T / dV issuer
      // dP^T(c) goes into the slot that held P^T(c - 2): it is issued right behind dV(c - 2) (same thread: in order).
      const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);   // dV: A = P^T in TMEM, B = dO chunk MN-major
      auto wait_operands = [&](const Walk& w) {
        if (w.jt == 0) WAIT(&qdo_full[w.j], (w.qpar >> w.j) & 1, 6);
        if (w.j == 0) WAIT(&kv_full[w.ts & 1], (w.ts >> 1) & 1, 7);
        tc_fence_after();
      };
      auto issue_dp = [&](const Walk& w) {
        const int slot = w.ts & 1;
        const int n_mma = (w.valid_cols() + 15) & ~15;
        const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
        const uint32_t a_v = DESC_LO_K + kv_lo + slot * (2 * TILE_BYTES >> 4) + (TILE_BYTES >> 4);
        const uint32_t b_do = DESC_LO_K + do_lo + w.j * (CW * 128 >> 4);
        const uint32_t t_dp = tmem + TM_DP + (w.c & 1) * CW;
        TR(0, 2, w.c);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(t_dp, a_v + 2 * kk, b_do + 2 * kk, idesc_s, kk > 0);
          umma_commit(&dp_full[w.c & 1]);
        }
        __syncwarp();
      };
      // dP^T may run up to two chunks ahead of dV, but a chunk AHEAD may only be issued when its operands have already
      // landed: a Q/dO block of a later group is released by (among others) this thread's own dV of an earlier chunk, so
      // blocking for it here could wait for itself (groups of a single chunk: 50-token crops).  The chunk dV is about to
      // need is the oldest one in flight -- everything its operands wait for has been issued -- and may block.
      auto operands_ready = [&](const Walk& w) {
        bool ok = true;
        if (w.jt == 0) ok = mbar_test(&qdo_full[w.j], (w.qpar >> w.j) & 1);
        if (ok && w.j == 0) ok = mbar_test(&kv_full[w.ts & 1], (w.ts >> 1) & 1);
        return __all_sync(0xffffffffu, ok) != 0;
      };
      Walk pc, dc;
      pc.init(prob);
      dc = pc;
      while (!dc.done(prob)) {
        while (!pc.done(prob) && pc.c <= dc.c) {
          wait_operands(pc);
          issue_dp(pc);
          pc.next_chunk(prob);
        }
        if (!pc.done(prob) && pc.c - dc.c < 2 && operands_ready(pc)) {
          tc_fence_after();
          issue_dp(pc);
          pc.next_chunk(prob);
        }
        const int slot = dc.ts & 1;
        WAIT(&pd_full[dc.c & 1], (dc.c >> 1) & 1, 8);
        if (dc.j == 0) WAIT(acc_empty, (dc.ts & 1) ^ 1, 9);
        tc_fence_after();
        const int n_k = (dc.valid_cols() + 15) >> 4;
        const uint32_t b_do = DESC_LO_MN + do_lo + dc.j * (CW * 128 >> 4);
        const uint32_t t_p = tmem + TM_DP + (dc.c & 1) * CW;
        TR(0, 4, dc.c);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < CW / 16; ++kk)
            if (kk < n_k) umma_ts(tmem + TM_DV, t_p + kk * 8, b_do + kk * 128, idesc_acc, (dc.j > 0 || kk > 0) ? 1u : 0u);
          if (dc.last_tile()) umma_commit(&qdo_empty[dc.j]);    // dO block j: read by dP^T and dV of this thread
          if (dc.last_chunk()) {
            umma_commit(acc_full);
            umma_commit(&kv_empty[slot]);                       // the V tile: read by this thread's dP^T MMAs
          }
        }
        __syncwarp();
        dc.next_chunk(prob);
        // dP^T of the chunk two ahead goes right behind this dV (same TMEM slot) if its operands are there already
        while (!pc.done(prob) && pc.c - dc.c < 2 && operands_ready(pc)) {
          tc_fence_after();
          issue_dp(pc);
          pc.next_chunk(prob);
        }
      }
    } else {
      // ------------------------------------------------------------------------------------------ dK / dQ issuer
      // dK += dS^T . Q_chunk, and once per query tile dQ_t += dS_t . K_tile; releases what it was the last to read.
      const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);   // dK: A = K-major panel, B = Q chunk MN-major
      const uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);    // dQ: A = the panel pair read MN-major (transposed)
      Walk a;
      a.init(prob);
      while (!a.done(prob)) {
        const int slot = a.ts & 1;
        WAIT(&pk_full[a.c & 3], (a.c >> 2) & 1, 10);
        if (a.j == 0) WAIT(acc_empty, (a.ts & 1) ^ 1, 11);
        tc_fence_after();
        const int n_k = (a.valid_cols() + 15) >> 4;
        const int tq = a.tq();
        const uint32_t a_ds = DESC_LO_K + ds_lo + ((tq & 1) * 2 + (a.j & 1)) * (TILE_BYTES >> 4);
        const uint32_t b_q = DESC_LO_MN + q_lo + a.j * (CW * 128 >> 4);
        TR(0, 5, a.c);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < CW / 16; ++kk)
            if (kk < n_k && !(ablate & 16)) umma_ss(tmem + TM_DK, a_ds + 2 * kk, b_q + kk * 128, idesc_acc, (a.j > 0 || kk > 0) ? 1u : 0u);
          if (a.last_tile()) umma_commit(&qdo_empty[a.j]);
        }
        __syncwarp();
        if ((a.j & 1) || a.last_chunk()) {
          // contraction over the keys of this tile (16 per k-step)
          const int t = a.j >> 1;
          if (a.jt == 0 && t == 0) {
            WAIT(dq_empty, (a.gi & 1) ^ 1, 12);
            tc_fence_after();
          }
          const int n_kk = (a.valid_keys() + 15) >> 4;
          const uint32_t a_pair = DESC_LO_MN + ds_lo + (tq & 1) * (2 * TILE_BYTES >> 4);
          const uint32_t b_k = DESC_LO_MN + kv_lo + slot * (2 * TILE_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < n_kk && !(ablate & 16))
                umma_ss(tmem + TM_DQ + t * 64, a_pair + kk * 128, b_k + kk * 128, idesc_dq, (a.jt > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&ds_empty[tq & 1]);
          }
          __syncwarp();
        }
        if (a.last_chunk()) {
          if (leader) {
            umma_commit(acc_full);
            umma_commit(&kv_empty[slot]);
            if (a.last_tile()) umma_commit(dq_full);
          }
          __syncwarp();
        }
        a.next_chunk(prob);
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------------------------------------ compute
    reg_alloc<REG_COMPUTE>();
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                  // key row of the tile == TMEM lane
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    const float sl2 = scale * LOG2E;
    const uint32_t sw = uint32_t(row & 7);
    Walk k;
    k.init(prob);
    while (!k.done(prob)) {
      if ((k.c & 1) != wg) {
        k.next_chunk(prob);
        continue;
      }
      const int valid = k.valid_cols();
      const int n_mma = (valid + 15) & ~15;
      const bool two = n_mma > 32;                       // the chunk has a second 32-query half
      const int kvalid = k.valid_keys();
      const int tq = k.tq();
      const bool active = quad * 32 < ((kvalid + 15) & ~15);   // warps past the last 16-key step have nothing to do
      const bool key_ok = row < kvalid;
      const uint32_t t_s = lane_addr + TM_S + (k.c & 1) * CW, t_dp = lane_addr + TM_DP + (k.c & 1) * CW;
      const uint32_t panel_row = sb + OFF_DS + ((tq & 1) * 2 + (k.j & 1)) * TILE_BYTES + row * 128;
      const uint32_t sL = a_sl + k.j * CW * 4, sD = a_sd + k.j * CW * 4;
      WAIT(&qdo_full[k.j], (k.qpar >> k.j) & 1, 6);                // this block's statistics are visible
      if (quad == 0) TR(1 + wg, 1, k.c);
      WAIT(&s_full[k.c & 1], (k.c >> 1) & 1, 7);
      if (quad == 0) TR(1 + wg, 2, k.c);
      tc_fence_after();
#ifdef APLA_AFB_PROF
      long long cp_t = clock64();
#define CP(i) do { const long long _n = clock64(); prof[i] += uint32_t(_n - cp_t); cp_t = _n; } while (0)
#else
#define CP(i)
#endif
      uint32_t p0[32], p1[32];
      if (active && !(ablate & 32)) {
        tmem_ld_32x32(t_s, p0);
        if (two) tmem_ld_32x32(t_s + 32, p1);
        tmem_ld_wait();
      }
      // the scores are in registers: the issuer may overwrite the slot with the scores of chunk c + 2
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[k.c & 1]);
      CP(10);
      if (active && !(ablate & 1)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 l = lds_f4(sL + 16 * i);
          p0[4 * i + 0] = __float_as_uint(exp2f(fmaf(__uint_as_float(p0[4 * i + 0]), sl2, -l.x)));
          p0[4 * i + 1] = __float_as_uint(exp2f(fmaf(__uint_as_float(p0[4 * i + 1]), sl2, -l.y)));
          p0[4 * i + 2] = __float_as_uint(exp2f(fmaf(__uint_as_float(p0[4 * i + 2]), sl2, -l.z)));
          p0[4 * i + 3] = __float_as_uint(exp2f(fmaf(__uint_as_float(p0[4 * i + 3]), sl2, -l.w)));
        }
        if (two) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 l = lds_f4(sL + 128 + 16 * i);
            p1[4 * i + 0] = __float_as_uint(exp2f(fmaf(__uint_as_float(p1[4 * i + 0]), sl2, -l.x)));
            p1[4 * i + 1] = __float_as_uint(exp2f(fmaf(__uint_as_float(p1[4 * i + 1]), sl2, -l.y)));
            p1[4 * i + 2] = __float_as_uint(exp2f(fmaf(__uint_as_float(p1[4 * i + 2]), sl2, -l.z)));
            p1[4 * i + 3] = __float_as_uint(exp2f(fmaf(__uint_as_float(p1[4 * i + 3]), sl2, -l.w)));
          }
        }
      }
      CP(11);
      if (quad == 0) TR(1 + wg, 3, k.c);
      WAIT(&dp_full[k.c & 1], (k.c >> 1) & 1, 8);
      if (quad == 0) TR(1 + wg, 4, k.c);
      tc_fence_after();
      CP(19);
      if (active && !(ablate & 8)) {
        // first half: dS^T = P^T o (dP^T - delta) -> panel; P^T -> columns [0, 16) of the dP^T slot (dP^T columns
        // [0, 32) are in registers by then)
        {
          uint32_t dv[32], pk[16];
          tmem_ld_32x32(t_dp, dv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 d = lds_f4(sD + 16 * i);
            pk[2 * i] = pack_bf16(__uint_as_float(p0[4 * i + 0]) * (__uint_as_float(dv[4 * i + 0]) - d.x),
                                  __uint_as_float(p0[4 * i + 1]) * (__uint_as_float(dv[4 * i + 1]) - d.y));
            pk[2 * i + 1] = pack_bf16(__uint_as_float(p0[4 * i + 2]) * (__uint_as_float(dv[4 * i + 2]) - d.z),
                                      __uint_as_float(p0[4 * i + 3]) * (__uint_as_float(dv[4 * i + 3]) - d.w));
          }
          CP(12);
          WAIT(&ds_empty[tq & 1], ((tq >> 1) & 1) ^ 1, 9);          // panel pair free of older MMAs
          CP(19);
          // dS^T row of this key: 32 queries = 64 B = four 16-byte units of the 128B-swizzled panel row; rows of keys
          // past the end of the sequence must contribute nothing to dQ
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 val = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            if (!key_ok) val = make_uint4(0u, 0u, 0u, 0u);
            sts_u4(panel_row + ((uint32_t(u) ^ sw) << 4), val);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(__uint_as_float(p0[2 * i]), __uint_as_float(p0[2 * i + 1]));
          tmem_st_32x16(t_dp, pk);
        }
        if (two) {
          uint32_t dv[32], pk[16];
          tmem_ld_32x32(t_dp + 32, dv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 d = lds_f4(sD + 128 + 16 * i);
            pk[2 * i] = pack_bf16(__uint_as_float(p1[4 * i + 0]) * (__uint_as_float(dv[4 * i + 0]) - d.x),
                                  __uint_as_float(p1[4 * i + 1]) * (__uint_as_float(dv[4 * i + 1]) - d.y));
            pk[2 * i + 1] = pack_bf16(__uint_as_float(p1[4 * i + 2]) * (__uint_as_float(dv[4 * i + 2]) - d.z),
                                      __uint_as_float(p1[4 * i + 3]) * (__uint_as_float(dv[4 * i + 3]) - d.w));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 val = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            if (!key_ok) val = make_uint4(0u, 0u, 0u, 0u);
            sts_u4(panel_row + ((uint32_t(4 + u) ^ sw) << 4), val);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(__uint_as_float(p1[2 * i]), __uint_as_float(p1[2 * i + 1]));
          tmem_st_32x16(t_dp + 16, pk);
        }
        CP(13);
        tmem_st_wait();
        fence_proxy_async_smem();
        CP(14);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&pd_full[k.c & 1]);
        mbar_arrive(&pk_full[k.c & 3]);
      }
      if (quad == 0) TR(1 + wg, 5, k.c);
      k.next_chunk(prob);
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------------------------------------------ epilogue
    reg_dealloc<REG_EPI>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    Walk k;
    k.init(prob);
    while (!k.done(prob)) {
      const int par = k.gi & 1, slot = k.ts & 1;
      const bool odd = k.odd != 0;
      const uint32_t vecs = a_vec + par * 1024;
      if (quad == 0) TR(3, 1, k.ts);
      WAIT(acc_full, k.ts & 1, 10);
      if (quad == 0) TR(3, 2, k.ts);
      tc_fence_after();
      WAIT(&hrow_full[slot], (k.ts >> 1) & 1, 11);
      if (quad == 0) TR(3, 3, k.ts);
      {
        const int r = k.jt * 128 + row;
        const bool ok = r < k.nt;
        __nv_bfloat16* base = dqkv + size_t(k.row_start + (ok ? r : 0)) * (3 * D) + k.h * 64;
        float cp = 0.f, cds = 0.f;
        if (odd) {
          cp = lds_f32(a_row + (slot * 256 + row) * 4);
          cds = lds_f32(a_row + (slot * 256 + 128 + row) * 4);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&hrow_empty[slot]);
        if (!(ablate & 4))
        drain_row64(lane_addr + TM_DK, scale, odd, cds, vecs, base + D, ok);        // dK = scale (dS^T Q + dS_ok Q_o)
        // dV = P^T dO + p_ok dO_o; the accumulators go back to the issuer before the (fire-and-forget) stores
        {
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            if (!(ablate & 4)) {
              tmem_ld_32x32(lane_addr + TM_DV + c * 32, ov);
              tmem_ld_wait();
            }
            if (c == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty);
            }
            if (odd) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 v = lds_f4(vecs + (192 + c * 32 + 4 * i) * 4);
                ov[4 * i + 0] = __float_as_uint(fmaf(cp, v.x, __uint_as_float(ov[4 * i + 0])));
                ov[4 * i + 1] = __float_as_uint(fmaf(cp, v.y, __uint_as_float(ov[4 * i + 1])));
                ov[4 * i + 2] = __float_as_uint(fmaf(cp, v.z, __uint_as_float(ov[4 * i + 2])));
                ov[4 * i + 3] = __float_as_uint(fmaf(cp, v.w, __uint_as_float(ov[4 * i + 3])));
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
              pk[c * 16 + i] = pack_bf16(__uint_as_float(ov[2 * i]), __uint_as_float(ov[2 * i + 1]));
          }
          if (ok && !(ablate & 4)) {
            uint4* d4 = reinterpret_cast<uint4*>(base + 2 * D);
#pragma unroll
            for (int j = 0; j < 8; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
        }
      }
      if (k.last_tile()) {
        WAIT(dq_full, k.gi & 1, 12);
        tc_fence_after();
        WAIT(&hcol_full[par], (k.gi >> 1) & 1, 13);
        for (int t = 0; t < k.ntiles; ++t) {
          const int r = t * 128 + row;
          const bool ok = r < k.nt;
          __nv_bfloat16* base = dqkv + size_t(k.row_start + (ok ? r : 0)) * (3 * D) + k.h * 64;
          const float cq = odd ? lds_f32(a_col + (par * 512 + 256 + r) * 4) : 0.f;
          // dQ = scale (dS K + dS_qo K_o); after the last tile's loads the accumulator goes back to the issuer (the
          // stores inside drain_row64 are fire-and-forget, so the hand-over only trails them by their issue)
          if (!(ablate & 4))
          drain_row64(lane_addr + TM_DQ + t * 64, scale, odd, cq, vecs + 64 * 4, base, ok);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dq_empty);
          mbar_arrive(&hcol_empty[par]);
        }
      }
      if (quad == 0) TR(3, 4, k.ts);
      // advance to the next key tile
      const int ts0 = k.ts;
      while (!k.done(prob) && k.ts == ts0) k.next_chunk(prob);
    }
  } else {
    // ------------------------------------------------------------------------------------------------ odd-token helper
    reg_dealloc<REG_HELP>();
    const int e = threadIdx.x - 512;                   // 0 .. 127
    const int dpair = e & 31, part = e >> 5;           // reductions: dims 2 dpair, 2 dpair + 1; rows part, part + 4, ...
    const float sl2 = scale * LOG2E;
    const uint32_t a_part = sb + OFF_PART;                 // [4 parts][dQ_o, dK_o, dV_o][64] floats
    const uint32_t dp_off = (uint32_t(dpair & 3) << 2);   // byte offset of the dim pair inside its 16-byte unit
#ifdef APLA_AFB_PROF
    long long hp_t = clock64();
#define HP(i) do { const long long _n = clock64(); prof[i] += uint32_t(_n - hp_t); hp_t = _n; } while (0)
#else
#define HP(i)
#endif
    Walk k;
    k.init(prob);
    int ts = 0;
    while (!k.done(prob)) {
      HP(0);
      const int par = k.gi & 1;
      const bool odd = k.odd != 0;
      const int o = k.nt;                               // the odd token's index within its sequence
      const uint32_t vecs = a_vec + par * 1024, cols = a_col + par * 2048;
      // this parity's vectors / column coefficients were last read by the epilogue of group gi - 2
      WAIT(&hcol_empty[par], ((k.gi >> 1) & 1) ^ 1, 14);
      float lse2_o = 0.f, delta_o = 0.f;
      if (odd) {
        const int v = e >> 5;                           // 0 Q_o, 1 K_o, 2 V_o, 3 dO_o
        const size_t tok = size_t(k.row_start + o);
        const __nv_bfloat16* src = v < 3 ? qkv + tok * (3 * D) + v * D + k.h * 64 : dout + tok * D + k.h * 64;
        const uint32_t w2 = __ldg(reinterpret_cast<const uint32_t*>(src) + lane);
        sts_f32(vecs + (v * 64 + 2 * lane) * 4, bf16_lo(w2));
        sts_f32(vecs + (v * 64 + 2 * lane + 1) * 4, bf16_hi(w2));
        lse2_o = __ldg(lse + tok * H + k.h) * LOG2E;
        delta_o = __ldg(delta + tok * H + k.h);
      }
      HP(1);
      // every Q/dO block of the group (and its statistics) must have landed before the column pass
      for (int j = 0; j < k.nchunks; ++j) WAIT(&qdo_full[j], (k.qpar >> j) & 1, 15);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      HP(0);
      float p_oo = 0.f, ds_oo = 0.f;
      float ko0 = 0.f, ko1 = 0.f, vo0 = 0.f, vo1 = 0.f;   // dK_o / dV_o partial sums over the queries of this part
      if (odd) {
        // the odd token against itself
        const float s_oo = dot_vec64(vecs, vecs + 64 * 4), dp_oo = dot_vec64(vecs + 192 * 4, vecs + 128 * 4);
        p_oo = exp2f(fmaf(s_oo, sl2, -lse2_o));
        ds_oo = p_oo * (dp_oo - delta_o);
        // as a key: p_qo = exp2(q . K_o * sl2 - lse_q), dS_qo = p_qo (dO_q . V_o - delta_q) for the queries e, e + 128
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          const int q = e + 128 * qq;
          float p = 0.f, ds = 0.f;
          if (q < k.nt) {
            float s, dp;
            dot2_row64(sb + OFF_Q + q * 128, sb + OFF_DO + q * 128, q, vecs + 64 * 4, vecs + 128 * 4, s, dp);
            p = exp2f(fmaf(s, sl2, -lds_f32(a_sl + q * 4)));
            ds = p * (dp - lds_f32(a_sd + q * 4));
          }
          sts_f32(cols + q * 4, p);
          sts_f32(cols + (256 + q) * 4, ds);
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      HP(2);
      if (odd) {
        // dK_o = sum_q dS_qo Q_q, dV_o = sum_q p_qo dO_q
#pragma unroll 4
        for (int q = part; q < k.nt; q += 4) {
          const float cds = lds_f32(cols + (256 + q) * 4), cp = lds_f32(cols + q * 4);
          const uint32_t off = uint32_t(q) * 128 + (((uint32_t(dpair) >> 2) ^ uint32_t(q & 7)) << 4) + dp_off;
          const uint32_t qw = lds_u32(sb + OFF_Q + off);
          const uint32_t dw = lds_u32(sb + OFF_DO + off);
          ko0 = fmaf(cds, bf16_lo(qw), ko0);
          ko1 = fmaf(cds, bf16_hi(qw), ko1);
          vo0 = fmaf(cp, bf16_lo(dw), vo0);
          vo1 = fmaf(cp, bf16_hi(dw), vo1);
        }
      }
      // done with the group's Q/dO blocks and statistics; the vectors and column coefficients are complete
      asm volatile("bar.sync 2, 128;" ::: "memory");
      HP(3);
      if (e == 0) {
        for (int j = 0; j < k.nchunks; ++j) mbar_arrive(&qdo_empty[j]);
        mbar_arrive(&hcol_full[par]);
      }
      // as a query: p_ok, dS_ok for the keys of every tile; dQ_o = sum_k dS_ok K_k
      float dq0 = 0.f, dq1 = 0.f;
      for (int jt = 0; jt < k.ntiles; ++jt, ++ts) {
        const int slot = ts & 1;
        const uint32_t rows = a_row + slot * 1024;
        const uint32_t sK = sb + OFF_KV + slot * 2 * TILE_BYTES;
        WAIT(&kv_full[slot], (ts >> 1) & 1, 16);
        WAIT(&hrow_empty[slot], ((ts >> 1) & 1) ^ 1, 17);
        const int kmax = min(128, k.nt - jt * 128);
        HP(0);
        if (odd) {
          float p = 0.f, ds = 0.f;
          if (e < kmax) {
            float s, dp;
            dot2_row64(sK + e * 128, sK + TILE_BYTES + e * 128, e, vecs, vecs + 192 * 4, s, dp);
            p = exp2f(fmaf(s, sl2, -lse2_o));
            ds = p * (dp - delta_o);
          }
          sts_f32(rows + e * 4, p);
          sts_f32(rows + (128 + e) * 4, ds);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        HP(4);
        if (e == 0) mbar_arrive(&hrow_full[slot]);
        if (odd) {
#pragma unroll 4
          for (int kk = part; kk < kmax; kk += 4) {
            const float c = lds_f32(rows + (128 + kk) * 4);
            const uint32_t kw = lds_u32(sK + uint32_t(kk) * 128 + (((uint32_t(dpair) >> 2) ^ uint32_t(kk & 7)) << 4) + dp_off);
            dq0 = fmaf(c, bf16_lo(kw), dq0);
            dq1 = fmaf(c, bf16_hi(kw), dq1);
          }
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");   // every helper thread is done with this tile's K / V
        HP(5);
        if (e == 0) mbar_arrive(&kv_empty[slot]);
      }
      if (odd) {
        // the three rows of the odd token: reduce the four parts, add the self term, write
        sts_f32(a_part + ((part * 3 + 0) * 64 + 2 * dpair) * 4, dq0);
        sts_f32(a_part + ((part * 3 + 0) * 64 + 2 * dpair + 1) * 4, dq1);
        sts_f32(a_part + ((part * 3 + 1) * 64 + 2 * dpair) * 4, ko0);
        sts_f32(a_part + ((part * 3 + 1) * 64 + 2 * dpair + 1) * 4, ko1);
        sts_f32(a_part + ((part * 3 + 2) * 64 + 2 * dpair) * 4, vo0);
        sts_f32(a_part + ((part * 3 + 2) * 64 + 2 * dpair + 1) * 4, vo1);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (e < 96) {
          const int v = e >> 5;                         // 0 dQ_o, 1 dK_o, 2 dV_o
          float x0 = 0.f, x1 = 0.f;
#pragma unroll
          for (int pt = 0; pt < 4; ++pt) {
            x0 += lds_f32(a_part + ((pt * 3 + v) * 64 + 2 * dpair) * 4);
            x1 += lds_f32(a_part + ((pt * 3 + v) * 64 + 2 * dpair + 1) * 4);
          }
          // self term: dQ_o += dS_oo K_o, dK_o += dS_oo Q_o, dV_o += p_oo dO_o
          const uint32_t sv = vecs + (v == 0 ? 64 : (v == 1 ? 0 : 192)) * 4;
          const float cs = v == 2 ? p_oo : ds_oo;
          const float mul = v == 2 ? 1.0f : scale;
          x0 = fmaf(cs, lds_f32(sv + 2 * dpair * 4), x0) * mul;
          x1 = fmaf(cs, lds_f32(sv + (2 * dpair + 1) * 4), x1) * mul;
          __nv_bfloat16* dst = dqkv + size_t(k.row_start + o) * (3 * D) + v * D + k.h * 64;
          reinterpret_cast<uint32_t*>(dst)[dpair] = pack_bf16(x0, x1);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");   // hpart is reused by the next group
      }
      HP(6);
      k.next_group(prob);
    }
  }
#ifdef APLA_AFB_PROF
  if (blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 8 || warp == 5))
    printf("afb compute warp %d phases: S-load %u exps %u dP-load+dS-half0 %u stores+half1 %u st-wait+fence %u\n", warp, prof[10], prof[11],
           prof[12], prof[13], prof[14]);
  if (blockIdx.x == 0 && lane == 0 && warp == 16)
    printf("afb helper phases: other/waits %u vec-load %u col %u dk-loop %u tile-dots %u dq-loop %u finalize %u\n", prof[0], prof[1],
           prof[2], prof[3], prof[4], prof[5], prof[6]);
  if (blockIdx.x == 0 && lane == 0 && (warp <= 2 || (warp & 3) == 0) && warp != 16)
    printf("afb prof warp %2d total %lld | kv_empty %u qdo_empty %u acc_empty %u dq_empty %u qdo_empty(st) %u | qdo_full %u s_full %u "
           "dp_full %u ds_empty %u | acc_full %u hrow_full %u dq_full %u hcol_full %u | hcol_empty %u qdo_full(h) %u kv_full(h) %u "
           "hrow_empty %u | idle-spin %u issued %u\n",
           warp, clock64() - prof_t0, prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7], prof[8], prof[9], prof[10],
           prof[11], prof[12], prof[13], prof[14], prof[15], prof[16], prof[17], prof[18], prof[19]);
#endif
  __syncwarp();
  tc_fence_before();
  __syncthreads();
#ifdef APLA_AFB_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const char* names[5] = {"issuer", "wg0", "wg1", "epi", "producer"};
    for (int r = 0; r < 5; ++r) {
      printf("afbtrace %s:", names[r]);
      for (int i = 0; i < tr_n[r]; ++i)
        printf(" %lld.%lld@%lld", (tr_buf[r][i] >> 8) & 255, tr_buf[r][i] & 255, tr_buf[r][i] >> 16);
      printf("\n");
    }
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, TM_COLS);
  }
}

}  // namespace afb

bool attn_bwd2_supported(int max_seqlen) { return max_seqlen > 0 && max_seqlen <= afb::NT_MAX + 1; }

int attn_bwd2(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
              const int* cu_seqlens, int num_seqs, int max_seqlen, int total_tokens, int H, float scale,
              cudaStream_t stream) {
  using namespace afb;
  APLA_CHECK(attn_bwd2_supported(max_seqlen), "attn_bwd2: max_seqlen %d exceeds the resident limit", max_seqlen);
  const int D = H * 64;
  const uint64_t T = total_tokens;
  Maps m;
  if (int rc = make_tmap_2d(&m.qkv64, qkv, 2, T, 3 * D, 3 * D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.qkv16, qkv, 2, T, 3 * D, 3 * D, 16, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.do64, dout, 2, T, D, D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.do16, dout, 2, T, D, D, 16, 64, true)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    APLA_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int G = num_seqs * H;
  const int grid = G < sm_count() ? G : sm_count();
  int ablate = 0;
#ifdef APLA_AFB_ABLATE
  if (const char* e = getenv("APLA_AFB_ABLATE")) ablate = atoi(e);
#endif
  attn_bwd2_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(
      m, reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(dout), lse, delta,
      reinterpret_cast<__nv_bfloat16*>(dqkv), cu_seqlens, max_seqlen, H, G, scale, ablate);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
