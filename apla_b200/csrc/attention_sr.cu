// "Sequence-resident" tcgen05 / TMEM attention backward for short sequences (N <= 272 tokens, head_dim 64): the
// ViT-at-224-px case (257 / 197 / 50 tokens) where one whole (sequence, head) pair fits in shared memory.
//
// One persistent CTA per SM walks over (sequence, head) groups.  For a group, the "resident" operand pair (all N rows)
// stays in shared memory while 128-row "stationary" tiles of the other pair stream past it; each (tile, 32-column
// chunk) is one pipeline unit:
//     SP   : T0 = A0 . B0_chunk^T,  T1 = A1 . B1_chunk^T       (smem x smem -> TMEM stage, fp32)
//     WG   : P = exp2(T0*scale*log2e - lse),  dS = P o (T1 - delta)   (one thread per TMEM lane, bf16 written in place)
//     ACC  : acc_dS += dS . B0_chunk  [, acc_P += P . B1_chunk]   (A from TMEM, B = the resident chunk read MN-major)
//   mode DKDV : stationary K, V tile (lanes = keys), resident Q, dO  ->  dK (acc_dS), dV (acc_P)
//   mode DQ   : stationary Q, dO tile (lanes = queries), resident K, V  ->  dQ (acc_dS)
// Both modes recompute the scores so that no accumulator crosses CTAs (no atomics, deterministic).
//
// Warp roles (14 warps): 0 = TMA producer (+ per-column lse/delta staging), 1 = MMA issuer (one thread), 2-5 / 6-9 = two
// compute warpgroups that take alternate chunks, 10-13 = epilogue (TMEM accumulators -> bf16 -> global).
// TMEM (512 columns): two accumulator buffers of 128 columns (epilogue of tile i overlaps the MMAs of tile i+1) and
// NS = 4 stages of (T0, T1) = 64 columns; the MMA thread runs LAG = 3 chunks ahead of the accumulate MMAs, so the
// tensor pipe always has queued work while the warpgroups do the exponentials.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace asr {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int NP_MAX = 272;                 // resident rows (257 padded to the MMA-N granularity of 16)
constexpr int STAT_N = 320;                 // per-column statistics, padded to a whole chunk
constexpr int kThreads = 448;
constexpr uint32_t RES_BYTES = NP_MAX * 128, STA_BYTES = 128 * 128;
constexpr uint32_t OFF_STA = 4 * RES_BYTES, OFF_STAT = OFF_STA + 4 * STA_BYTES, OFF_BAR = OFF_STAT + 4 * STAT_N * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr uint32_t TM_STAGE0 = 256, TM_COLS = 512;
// (chunk width, stages, MMA look-ahead): stages * 2 * CW = 256 TMEM columns
#ifndef APLA_SR_CW
#define APLA_SR_CW 64
#endif
constexpr int CW = APLA_SR_CW, NS = 128 / CW, LAG = NS - 1;
enum { DKDV = 0, DQ = 1 };

__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptors (128B swizzle, 8-row groups 1024 B apart) split into 32-bit halves so that the
// per-MMA address arithmetic is a single 32-bit add:  lo = (addr >> 4) | LBO>>4 << 16,  hi = SBO>>4 | version | layout.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_LO_K = (16u >> 4) << 16;        // K-major operand: +2 per 16-element (32 B) k-step
constexpr uint32_t DESC_LO_MN = (16384u >> 4) << 16;    // MN-major operand: +128 per 16-row (2048 B) k-step

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

struct Problem {
  const int* cu;
  int n_fixed, H, G;
};

// Deterministic walk over this CTA's (group, tile, chunk) sequence; every warp role runs its own copy.
struct Cursor {
  int g, gi, t, ii, j, c;
  int row_start, n, h, ntiles, nchunks;
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
    // every lane of a warp walks the same schedule: say so, so that the schedule lives in uniform registers
    row_start = __shfl_sync(0xffffffffu, row_start, 0);
    n = __shfl_sync(0xffffffffu, n, 0);
    ntiles = (n + 127) >> 7;
    nchunks = (n + CW - 1) / CW;
  }
  __device__ __forceinline__ void init(const Problem& p) {
    g = blockIdx.x;
    gi = t = ii = j = c = 0;
    row_start = n = h = 0;
    load(p);
  }
  __device__ __forceinline__ bool done(const Problem& p) const { return g >= p.G; }
  __device__ __forceinline__ void next_group(const Problem& p) {
    t = 0;
    ++gi;
    g += gridDim.x;
    load(p);
  }
  __device__ __forceinline__ void next_item(const Problem& p) {
    j = 0;
    ++ii;
    if (++t == ntiles) next_group(p);
  }
  __device__ __forceinline__ void next_chunk(const Problem& p) {
    ++c;
    if (++j == nchunks) next_item(p);
  }
  __device__ __forceinline__ int valid_cols() const { return min(CW, n - j * CW); }
};

// 64 fp32 accumulator columns of this thread's row -> bf16 -> 128 contiguous bytes in global memory
__device__ __forceinline__ void store_row64(uint32_t taddr, __nv_bfloat16* dst, float mul, bool store) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t ov[32];
    tmem_ld_32x32(taddr + c * 32, ov);
    tmem_ld_wait();
    if (store) {
      uint4* d4 = reinterpret_cast<uint4*>(dst) + c * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(ov[8 * j + i]) * mul;
        d4[j] = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
    }
  }
}

struct Maps {
  CUtensorMap sta0, sta1;      // 128-row boxes of the stationary pair
  CUtensorMap res0_64, res0_16, res1_64, res1_16;  // 64- and 16-row boxes of the resident pair
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_sr_kernel(const __grid_constant__ Maps maps, const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dqkv, const int* __restrict__ cu_seqlens, int n_fixed, int H, int G,
                   float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* res_full = bars;        // [2]
  uint64_t* res_empty = bars + 2;   // [2]
  uint64_t* sta_full = bars + 4;    // [2]
  uint64_t* sta_empty = bars + 6;   // [2]
  uint64_t* sp_full = bars + 8;     // [NS]
  uint64_t* pd_full = bars + 12;    // [NS]
  uint64_t* acc_full = bars + 16;   // [2]
  uint64_t* acc_empty = bars + 18;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const Problem prob{cu_seqlens, n_fixed, H, G};
  const int D = H * 64;
  // column offsets of the two operand pairs inside the packed qkv matrix / dO
  // DKDV: stationary (K, V), resident (Q, dO);  DQ: stationary (Q, dO), resident (K, V)
  const int sta0_col = MODE == DKDV ? D : 0, sta1_col = MODE == DKDV ? 2 * D : 0;
  const int res0_col = MODE == DKDV ? 0 : D, res1_col = MODE == DKDV ? 0 : 2 * D;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.sta0);
    tma_prefetch_desc(&maps.sta1);
    tma_prefetch_desc(&maps.res0_64);
    tma_prefetch_desc(&maps.res0_16);
    tma_prefetch_desc(&maps.res1_64);
    tma_prefetch_desc(&maps.res1_16);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], MODE == DKDV ? 2 : 1);
      mbar_init(&res_empty[i], 1);
      mbar_init(&sta_full[i], 1);
      mbar_init(&sta_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&sp_full[i], 1);
      mbar_init(&pd_full[i], 4);
    }
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

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ producer
    Cursor k;
    k.init(prob);
    while (!k.done(prob)) {
      const int gb = k.gi & 1;
      if (lane == 0) {
        mbar_wait(&res_empty[gb], ((k.gi >> 1) & 1) ^ 1);
        const int np = (k.n + 15) & ~15;
        mbar_arrive_expect_tx(&res_full[gb], 2u * np * 128u);
        uint8_t* r0 = smem + (gb * 2 + 0) * RES_BYTES;
        uint8_t* r1 = smem + (gb * 2 + 1) * RES_BYTES;
        int r = 0;
        for (; r + 64 <= np; r += 64) {
          tma_load_2d(r0 + r * 128, &maps.res0_64, &res_full[gb], res0_col + k.h * 64, k.row_start + r);
          tma_load_2d(r1 + r * 128, &maps.res1_64, &res_full[gb], res1_col + k.h * 64, k.row_start + r);
        }
        for (; r < np; r += 16) {
          tma_load_2d(r0 + r * 128, &maps.res0_16, &res_full[gb], res0_col + k.h * 64, k.row_start + r);
          tma_load_2d(r1 + r * 128, &maps.res1_16, &res_full[gb], res1_col + k.h * 64, k.row_start + r);
        }
      }
      __syncwarp();
      if (MODE == DKDV) {
        // per-column (= per-query) statistics of the group; +inf / 0 past the end masks the padded columns
        float* sL = reinterpret_cast<float*>(smem + OFF_STAT) + (gb * 2 + 0) * STAT_N;
        float* sD = reinterpret_cast<float*>(smem + OFF_STAT) + (gb * 2 + 1) * STAT_N;
        for (int i = lane; i < STAT_N; i += 32) {
          const bool ok = i < k.n;
          const size_t idx = size_t(k.row_start + (ok ? i : 0)) * H + k.h;
          sL[i] = ok ? __ldg(lse + idx) * LOG2E : INFINITY;
          sD[i] = ok ? __ldg(delta + idx) : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&res_full[gb]);
      }
      const int gi0 = k.gi;
      do {   // all tiles of this group
        const int sb = k.ii & 1;
        if (lane == 0) {
          mbar_wait(&sta_empty[sb], ((k.ii >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&sta_full[sb], 2 * STA_BYTES);
          uint8_t* s0 = smem + OFF_STA + (sb * 2 + 0) * STA_BYTES;
          uint8_t* s1 = smem + OFF_STA + (sb * 2 + 1) * STA_BYTES;
          tma_load_2d(s0, &maps.sta0, &sta_full[sb], sta0_col + k.h * 64, k.row_start + k.t * 128);
          tma_load_2d(s1, &maps.sta1, &sta_full[sb], sta1_col + k.h * 64, k.row_start + k.t * 128);
        }
        __syncwarp();
        k.next_item(prob);
      } while (k.gi == gi0);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    // The whole warp walks the (uniform) schedule so that the address arithmetic stays on the uniform datapath; one
    // thread issues every tcgen05.mma / commit.
    const bool leader = elect_one();
    const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t res_lo = smem_u32(smem) >> 4, sta_lo = smem_u32(smem + OFF_STA) >> 4;
    auto do_acc = [&](const Cursor& a) {
      const int stage = a.c % NS;
      const int n_k = (a.valid_cols() + 15) >> 4;
      const int ab = a.ii & 1, gb = a.gi & 1;
      mbar_wait(&pd_full[stage], (a.c / NS) & 1);
      if (a.j == 0) mbar_wait(&acc_empty[ab], ((a.ii >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t b0 = DESC_LO_MN + res_lo + (gb * 2) * (RES_BYTES >> 4) + a.j * (CW * 128 >> 4);
      const uint32_t b1 = b0 + (RES_BYTES >> 4);
      const uint32_t t0 = tmem + TM_STAGE0 + stage * (2 * CW), t1 = t0 + CW;
      const uint32_t acc = tmem + ab * 128;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < CW / 16; ++kk) {
          if (kk < n_k) {
            const uint32_t accum = (a.j > 0 || kk > 0) ? 1u : 0u;
            umma_ts(acc, t1 + kk * 8, b0 + kk * 128, idesc_acc, accum);
            if (MODE == DKDV) umma_ts(acc + 64, t0 + kk * 8, b1 + kk * 128, idesc_acc, accum);
          }
        }
        if (a.j == a.nchunks - 1) {
          umma_commit(&acc_full[ab]);
          if (a.t == a.ntiles - 1) umma_commit(&res_empty[gb]);
        }
      }
      __syncwarp();
    };
    Cursor sp, ac;
    sp.init(prob);
    ac = sp;
    int ahead = 0;
    while (!sp.done(prob)) {
      const int gb = sp.gi & 1, sb = sp.ii & 1;
      if (sp.j == 0) {
        if (sp.t == 0) mbar_wait(&res_full[gb], (sp.gi >> 1) & 1);
        mbar_wait(&sta_full[sb], (sp.ii >> 1) & 1);
        tc_fence_after();
      }
      const int stage = sp.c % NS;
      const int n_mma = (sp.valid_cols() + 15) & ~15;
      const uint32_t a0 = DESC_LO_K + sta_lo + (sb * 2) * (STA_BYTES >> 4), a1 = a0 + (STA_BYTES >> 4);
      const uint32_t b0 = DESC_LO_K + res_lo + (gb * 2) * (RES_BYTES >> 4) + sp.j * (CW * 128 >> 4);
      const uint32_t b1 = b0 + (RES_BYTES >> 4);
      const uint32_t t0 = tmem + TM_STAGE0 + stage * (2 * CW), t1 = t0 + CW;
      const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(t0, a0 + 2 * kk, b0 + 2 * kk, idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(t1, a1 + 2 * kk, b1 + 2 * kk, idesc_s, kk > 0);
        umma_commit(&sp_full[stage]);
        if (sp.j == sp.nchunks - 1) umma_commit(&sta_empty[sb]);
      }
      __syncwarp();
      sp.next_chunk(prob);
      if (++ahead > LAG) {
        do_acc(ac);
        ac.next_chunk(prob);
      }
    }
    while (!ac.done(prob)) {
      do_acc(ac);
      ac.next_chunk(prob);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------------------------------------ compute
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                  // row of the stationary tile == TMEM lane
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    const float sl2 = scale * LOG2E;
    Cursor k;
    k.init(prob);
    int cur_gi = -1, cur_ii = -1;
    float lse2 = 0.f, dl = 0.f;          // DQ: per-row statistics of the current tile
    float nx_lse2 = 0.f, nx_dl = 0.f;    // DQ: prefetched for the next tile
    int nx_ii = -1;
    while (!k.done(prob)) {
      if ((k.c & 1) != wg) {
        k.next_chunk(prob);
        continue;
      }
      const int gb = k.gi & 1;
      if (MODE == DKDV) {
        if (k.gi != cur_gi) {
          mbar_wait(&res_full[gb], (k.gi >> 1) & 1);   // the staged statistics of this group are visible
          cur_gi = k.gi;
        }
      } else if (k.ii != cur_ii) {
        if (nx_ii == k.ii) {
          lse2 = nx_lse2;
          dl = nx_dl;
        } else {
          const int r = k.t * 128 + row;
          const bool ok = r < k.n;
          const size_t idx = size_t(k.row_start + (ok ? r : 0)) * H + k.h;
          lse2 = ok ? __ldg(lse + idx) * LOG2E : 0.f;
          dl = ok ? __ldg(delta + idx) : 0.f;
        }
        cur_ii = k.ii;
        Cursor nx = k;
        nx.next_item(prob);
        if (!nx.done(prob)) {
          const int r = nx.t * 128 + row;
          const bool ok = r < nx.n;
          const size_t idx = size_t(nx.row_start + (ok ? r : 0)) * H + nx.h;
          nx_lse2 = ok ? __ldg(lse + idx) * LOG2E : 0.f;
          nx_dl = ok ? __ldg(delta + idx) : 0.f;
          nx_ii = nx.ii;
        }
      }
      const int stage = k.c % NS;
      const int valid = k.valid_cols();
      const uint32_t t0 = lane_addr + TM_STAGE0 + stage * (2 * CW), t1 = t0 + CW;
      mbar_wait(&sp_full[stage], (k.c / NS) & 1);
      tc_fence_after();
      // warps whose 32 rows all lie past the end of the sequence have nothing to compute (their TMEM rows hold
      // garbage that only reaches accumulator rows which are never stored)
      if (k.t * 128 + quad * 32 < k.n) {
        const int n_mma = (valid + 15) & ~15;
#pragma unroll
        for (int hf = 0; hf < CW / 32; ++hf) {
          if (hf * 32 < n_mma) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(t0 + hf * 32, sv);
            tmem_ld_32x32(t1 + hf * 32, dv);
            tmem_ld_wait();
            uint32_t pp[16], pd[16];
            if (MODE == DKDV) {
              const float* sL =
                  reinterpret_cast<const float*>(smem + OFF_STAT) + (gb * 2 + 0) * STAT_N + k.j * CW + hf * 32;
              const float* sD =
                  reinterpret_cast<const float*>(smem + OFF_STAT) + (gb * 2 + 1) * STAT_N + k.j * CW + hf * 32;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 l = reinterpret_cast<const float4*>(sL)[i];
                const float4 d = reinterpret_cast<const float4*>(sD)[i];
                const float p0 = exp2f(fmaf(__uint_as_float(sv[4 * i + 0]), sl2, -l.x));
                const float p1 = exp2f(fmaf(__uint_as_float(sv[4 * i + 1]), sl2, -l.y));
                const float p2 = exp2f(fmaf(__uint_as_float(sv[4 * i + 2]), sl2, -l.z));
                const float p3 = exp2f(fmaf(__uint_as_float(sv[4 * i + 3]), sl2, -l.w));
                pp[2 * i] = pack_bf16(p0, p1);
                pp[2 * i + 1] = pack_bf16(p2, p3);
                pd[2 * i] =
                    pack_bf16(p0 * (__uint_as_float(dv[4 * i + 0]) - d.x), p1 * (__uint_as_float(dv[4 * i + 1]) - d.y));
                pd[2 * i + 1] =
                    pack_bf16(p2 * (__uint_as_float(dv[4 * i + 2]) - d.z), p3 * (__uint_as_float(dv[4 * i + 3]) - d.w));
              }
              tmem_st_32x16(t0 + hf * 16, pp);
              tmem_st_32x16(t1 + hf * 16, pd);
            } else {
              if (valid == CW) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float p0 = exp2f(fmaf(__uint_as_float(sv[2 * i]), sl2, -lse2));
                  const float p1 = exp2f(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -lse2));
                  pd[i] = pack_bf16(p0 * (__uint_as_float(dv[2 * i]) - dl), p1 * (__uint_as_float(dv[2 * i + 1]) - dl));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int col = hf * 32 + 2 * i;
                  const float p0 = exp2f(fmaf(__uint_as_float(sv[2 * i]), sl2, -lse2));
                  const float p1 = exp2f(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -lse2));
                  const float e0 = col < valid ? p0 * (__uint_as_float(dv[2 * i]) - dl) : 0.f;
                  const float e1 = col + 1 < valid ? p1 * (__uint_as_float(dv[2 * i + 1]) - dl) : 0.f;
                  pd[i] = pack_bf16(e0, e1);
                }
              }
              tmem_st_32x16(t1 + hf * 16, pd);
            }
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full[stage]);
      k.next_chunk(prob);
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    Cursor k;
    k.init(prob);
    while (!k.done(prob)) {
      const int ab = k.ii & 1;
      mbar_wait(&acc_full[ab], (k.ii >> 1) & 1);
      tc_fence_after();
      const int r = k.t * 128 + row;
      const bool ok = r < k.n;
      __nv_bfloat16* base = dqkv + size_t(k.row_start + (ok ? r : 0)) * (3 * D) + k.h * 64;
      if (MODE == DKDV) {
        store_row64(lane_addr + ab * 128, base + D, scale, ok);           // dK = scale * dS^T Q
        store_row64(lane_addr + ab * 128 + 64, base + 2 * D, 1.0f, ok);   // dV = P^T dO
      } else {
        store_row64(lane_addr + ab * 128, base, scale, ok);               // dQ = scale * dS K
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      k.next_item(prob);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, TM_COLS);
  }
}

}  // namespace asr

bool attn_sr_supported(int max_seqlen) { return max_seqlen > 0 && max_seqlen <= asr::NP_MAX; }

int attn_bwd_sr(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                const int* cu_seqlens, int num_seqs, int max_seqlen, int total_tokens, int H, float scale,
                cudaStream_t stream) {
  using namespace asr;
  APLA_CHECK(attn_sr_supported(max_seqlen), "attn_bwd_sr: max_seqlen %d exceeds the resident limit", max_seqlen);
  const int D = H * 64;
  const uint64_t T = total_tokens;
  Maps mkv, mq;   // mkv: DKDV mode (stationary K,V / resident Q,dO);  mq: DQ mode (stationary Q,dO / resident K,V)
  if (int rc = make_tmap_2d(&mkv.sta0, qkv, 2, T, 3 * D, 3 * D, 128, 64, true)) return rc;
  mkv.sta1 = mkv.sta0;
  if (int rc = make_tmap_2d(&mkv.res0_64, qkv, 2, T, 3 * D, 3 * D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&mkv.res0_16, qkv, 2, T, 3 * D, 3 * D, 16, 64, true)) return rc;
  if (int rc = make_tmap_2d(&mkv.res1_64, dout, 2, T, D, D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&mkv.res1_16, dout, 2, T, D, D, 16, 64, true)) return rc;
  mq.sta0 = mkv.sta0;
  if (int rc = make_tmap_2d(&mq.sta1, dout, 2, T, D, D, 128, 64, true)) return rc;
  mq.res0_64 = mkv.res0_64;
  mq.res0_16 = mkv.res0_16;
  mq.res1_64 = mkv.res0_64;
  mq.res1_16 = mkv.res0_16;
  static bool attr_set = false;
  if (!attr_set) {
    APLA_CUDA(cudaFuncSetAttribute(attn_bwd_sr_kernel<DKDV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    APLA_CUDA(cudaFuncSetAttribute(attn_bwd_sr_kernel<DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int G = num_seqs * H;
  const int grid = G < sm_count() ? G : sm_count();
  attn_bwd_sr_kernel<DKDV><<<grid, kThreads, SMEM_BYTES, stream>>>(mkv, lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv),
                                                                  cu_seqlens, max_seqlen, H, G, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  attn_bwd_sr_kernel<DQ><<<grid, kThreads, SMEM_BYTES, stream>>>(mq, lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv),
                                                                cu_seqlens, max_seqlen, H, G, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
