// Fused single-pass tcgen05 / TMEM attention backward for short sequences (N <= 272 tokens, head_dim 64) -- the
// ViT-at-224-px case (257 / 197 / 50 tokens).  One launch produces dQ, dK and dV; the scores are recomputed ONCE.
//
// One persistent CTA per SM walks over (sequence, head) groups.  A group's Q and dO rows stay in shared memory
// (in 64-row blocks that are released one by one during the group's last key tile, so the next group's rows stream
// in underneath); 128-key tiles of K and V pass through a two-slot ring.  Orientation: TMEM lanes = keys.
// Per (key tile, 64-query chunk):
//     SP   : S^T = K_tile . Q_chunk^T,   dP^T = V_tile . dO_chunk^T           (smem x smem -> TMEM, fp32)
//     WG   : P^T = exp2(S^T*scale*log2e - lse_q),  dS^T = P^T o (dP^T - delta_q)
//            P^T  -> TMEM (bf16, in place of S^T)            : A operand of the dV MMA
//            dS^T -> shared-memory panel [128 keys x 64 q]   : A operand of the dK MMA (K-major view) AND of the
//                                                              dQ MMA (the same bytes viewed MN-major = transposed)
//     ACC  : dV += P^T . dO_chunk,   dK += dS^T . Q_chunk,   and once per 128 queries  dQ_t += dS_t . K_tile
// dK/dV accumulate across the chunks of a key tile, dQ (all query tiles of the group) across its key tiles, so no
// accumulator ever leaves the CTA: no atomics, deterministic.
//
// Warp roles (15 warps): 0 = TMA producer, 1 = MMA issuer (one thread), 2-5 / 6-9 = two compute warpgroups taking
// alternate chunks, 10-13 = epilogue (TMEM accumulators -> bf16 -> global), 14 = per-query lse/delta staging.
// TMEM (512 columns): dQ 3 x 64 | dK 64 | dV 64 | S^T/P^T ring 2 x 64 | dP^T 64.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace afu {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int NP_MAX = 272;                  // resident rows (257 padded to the MMA granularity of 16)
constexpr int CW = 64;                       // queries per chunk
constexpr int MAX_CHUNKS = 5;                // ceil(272 / 64)
constexpr int kThreads = 480;
constexpr uint32_t RES_BYTES = NP_MAX * 128;             // one resident operand (Q or dO)
constexpr uint32_t TILE_BYTES = 128 * 128;               // one 128-row K or V tile / one dS^T panel
constexpr uint32_t OFF_Q = 0, OFF_DO = RES_BYTES, OFF_KV = 2 * RES_BYTES;   // KV: [2 slots][K, V]
constexpr uint32_t OFF_DS = OFF_KV + 4 * TILE_BYTES;                          // [2 pairs][2 panels]
constexpr uint32_t OFF_STAT = OFF_DS + 4 * TILE_BYTES;                        // lse2[320], delta[320]
constexpr int STAT_N = MAX_CHUNKS * CW;
constexpr uint32_t OFF_BAR = OFF_STAT + 2 * STAT_N * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr uint32_t TM_DQ = 0, TM_DK = 192, TM_DV = 256, TM_S = 320, TM_DP = 448, TM_COLS = 512;

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

// -DAPLA_AFU_PROF: CTA 0 prints the cycles each role spent blocked at each kind of wait (development aid)
#ifdef APLA_AFU_PROF
#define PW(id, ...)                      \
  do {                                   \
    const long long _t0 = clock64();     \
    __VA_ARGS__;                         \
    prof[id] += clock64() - _t0;         \
  } while (0)
#define PROF_DECL long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64()
#define PROF_DUMP(role)                                                                                          \
  if (blockIdx.x == 0 && lane == 0)                                                                              \
  printf("afu %s: total %lld  w0 %lld w1 %lld w2 %lld w3 %lld w4 %lld w5 %lld w6 %lld w7 %lld\n", role,             \
         clock64() - prof_t0, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7])
#else
#define PW(id, ...) \
  do {              \
    __VA_ARGS__;    \
  } while (0)
#define PROF_DECL
#define PROF_DUMP(role)
#endif

struct Problem {
  const int* cu;
  int n_fixed, H, G;
};

// Deterministic walk over this CTA's (group, key tile, query chunk) sequence; every warp role runs its own copy.
//   gi = groups done, ts = key tiles done, c = chunks done, tqb = query tiles (chunk pairs) done before this key tile
struct Walk {
  int g, gi, ts, c, tqb, jt, j;
  int row_start, n, h, ntiles, nchunks;
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
    ntiles = (n + 127) >> 7;
    nchunks = (n + CW - 1) / CW;
  }
  __device__ __forceinline__ void init(const Problem& p) {
    g = blockIdx.x;
    gi = ts = c = tqb = jt = j = 0;
    row_start = n = h = 0;
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
  __device__ __forceinline__ int valid_cols() const { return min(CW, n - j * CW); }       // queries in this chunk
  __device__ __forceinline__ int valid_keys() const { return min(128, n - jt * 128); }    // keys in this tile
  __device__ __forceinline__ int tq() const { return tqb + (j >> 1); }                      // query-tile sequence number
  __device__ __forceinline__ bool last_chunk() const { return j == nchunks - 1; }
  __device__ __forceinline__ bool last_tile() const { return jt == ntiles - 1; }
};

// 64 fp32 accumulator columns of this thread's row -> 32 packed bf16 pairs in registers
__device__ __forceinline__ void load_row64(uint32_t taddr, float mul, uint32_t (&pk)[32]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t ov[32];
    tmem_ld_32x32(taddr + c * 32, ov);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      pk[c * 16 + i] = pack_bf16(__uint_as_float(ov[2 * i]) * mul, __uint_as_float(ov[2 * i + 1]) * mul);
  }
}
// ... -> 128 contiguous bytes in global memory
__device__ __forceinline__ void store_row64(const uint32_t (&pk)[32], __nv_bfloat16* dst, bool store) {
  if (store) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}

struct Maps {
  CUtensorMap qkv64, qkv16, do64, do16;   // 64- and 16-row boxes (64 bf16 columns) of the packed qkv matrix / dO
};

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_fused_kernel(const __grid_constant__ Maps maps, const float* __restrict__ lse, const float* __restrict__ delta,
                      __nv_bfloat16* __restrict__ dqkv, const int* __restrict__ cu_seqlens, int n_fixed, int H, int G,
                      float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* qdo_full = bars;          // [5]  Q/dO block + its statistics landed            (TMA tx + producer warp)
  uint64_t* qdo_empty = bars + 5;     // [5]  block no longer read by any MMA                 (commit)
  uint64_t* kv_full = bars + 10;      // [2]
  uint64_t* kv_empty = bars + 12;     // [2]
  uint64_t* sp_full = bars + 14;      // [2]  S^T and dP^T of a chunk are in TMEM             (commit)
  uint64_t* pd_full = bars + 16;      // [2]  P^T in TMEM and dS^T panel in smem are written  (4 warps)
  uint64_t* dp_free = bars + 18;      //      dP^T columns have been read into registers      (4 warps)
  uint64_t* ds_empty = bars + 19;     // [2]  panel pair no longer read by any MMA            (commit)
  uint64_t* acc_full = bars + 21;     //      dK/dV of a key tile complete                    (commit)
  uint64_t* acc_empty = bars + 22;    //      ... and stored                                  (4 warps)
  uint64_t* dq_full = bars + 23;      //      dQ of a group complete                          (commit)
  uint64_t* dq_empty = bars + 24;     //      ... and stored                                  (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const Problem prob{cu_seqlens, n_fixed, H, G};
  const int D = H * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.qkv64);
    tma_prefetch_desc(&maps.qkv16);
    tma_prefetch_desc(&maps.do64);
    tma_prefetch_desc(&maps.do16);
    for (int i = 0; i < MAX_CHUNKS; ++i) {
      mbar_init(&qdo_full[i], 2);
      mbar_init(&qdo_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&sp_full[i], 1);
      mbar_init(&pd_full[i], 4);
      mbar_init(&ds_empty[i], 1);
    }
    mbar_init(dp_free, 4);
    mbar_init(acc_full, 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ producer
    // Load order per group: key tile 0, the Q/dO blocks, the remaining key tiles -- the order in which the consumers
    // free the buffers, so no wait here can depend on a load that is issued later.
    auto load_rows = [&](uint8_t* dst, const CUtensorMap* m64, const CUtensorMap* m16, uint64_t* bar, int col, int row,
                         int rows) {
      int r = 0;
      for (; r + 64 <= rows; r += 64) tma_load_2d(dst + r * 128, m64, bar, col, row + r);
      for (; r < rows; r += 16) tma_load_2d(dst + r * 128, m16, bar, col, row + r);
    };
    Walk k;
    k.init(prob);
    PROF_DECL;
    int ts = 0;   // key tiles loaded so far
    while (!k.done(prob)) {
      const int np = (k.n + 15) & ~15;
      auto load_kv = [&](int jt) {
        const int slot = ts & 1;
        if (lane == 0) {
          const int rows = min(128, np - jt * 128);
          PW(0, mbar_wait(&kv_empty[slot], ((ts >> 1) & 1) ^ 1));
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
        if (lane == 0) {
          PW(1, mbar_wait(&qdo_empty[j], ((k.qpar >> j) & 1) ^ 1));
          mbar_arrive_expect_tx(&qdo_full[j], 2u * rows * 128u);
          load_rows(smem + OFF_Q + j * CW * 128, &maps.qkv64, &maps.qkv16, &qdo_full[j], k.h * 64,
                    k.row_start + j * CW, rows);
          load_rows(smem + OFF_DO + j * CW * 128, &maps.do64, &maps.do16, &qdo_full[j], k.h * 64, k.row_start + j * CW,
                    rows);
        }
      }
      for (int jt = 1; jt < k.ntiles; ++jt) load_kv(jt);
      k.next_group(prob);
    }
    PROF_DUMP("producer(kv_empty,qdo_empty)");
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    // The whole warp walks the (uniform) schedule; one thread issues every tcgen05.mma / commit.
    const bool leader = elect_one();
    PROF_DECL;
    const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);   // dV (A: TMEM), dK (A: K-major panel); B MN-major
    const uint32_t idesc_dq = make_idesc_bf16(128, 64, 1, 1);    // dQ: A = the panel pair read MN-major (transposed)
    const uint32_t q_lo = smem_u32(smem + OFF_Q) >> 4, do_lo = smem_u32(smem + OFF_DO) >> 4;
    const uint32_t kv_lo = smem_u32(smem + OFF_KV) >> 4, ds_lo = smem_u32(smem + OFF_DS) >> 4;

    auto issue_sp = [&](const Walk& w) {
      const int slot = w.ts & 1;
      if (w.jt == 0) PW(0, mbar_wait(&qdo_full[w.j], (w.qpar >> w.j) & 1));
      if (w.j == 0) PW(1, mbar_wait(&kv_full[slot], (w.ts >> 1) & 1));
      tc_fence_after();
      const int n_mma = (w.valid_cols() + 15) & ~15;
      const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
      const uint32_t a_k = DESC_LO_K + kv_lo + slot * (2 * TILE_BYTES >> 4), a_v = a_k + (TILE_BYTES >> 4);
      const uint32_t b_q = DESC_LO_K + q_lo + w.j * (CW * 128 >> 4), b_do = DESC_LO_K + do_lo + w.j * (CW * 128 >> 4);
      const uint32_t t_s = tmem + TM_S + (w.c & 1) * CW, t_dp = tmem + TM_DP;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(t_s, a_k + 2 * kk, b_q + 2 * kk, idesc_s, kk > 0);
      }
      __syncwarp();
      if (w.c > 0) {   // the single dP^T buffer: the previous chunk's warps must have it in registers
        PW(2, mbar_wait(dp_free, (w.c - 1) & 1));
        tc_fence_after();
      }
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(t_dp, a_v + 2 * kk, b_do + 2 * kk, idesc_s, kk > 0);
        umma_commit(&sp_full[w.c & 1]);
      }
      __syncwarp();
    };
    // dV += P^T . dO_chunk first: it is what keeps the S^T/P^T slot of this chunk busy
    auto issue_dv = [&](const Walk& a) {
      const int n_k = (a.valid_cols() + 15) >> 4;
      PW(3, mbar_wait(&pd_full[a.c & 1], (a.c >> 1) & 1));
      if (a.j == 0) PW(4, mbar_wait(acc_empty, (a.ts & 1) ^ 1));
      tc_fence_after();
      const uint32_t b_do = DESC_LO_MN + do_lo + a.j * (CW * 128 >> 4);
      const uint32_t t_p = tmem + TM_S + (a.c & 1) * CW;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < CW / 16; ++kk)
          if (kk < n_k) umma_ts(tmem + TM_DV, t_p + kk * 8, b_do + kk * 128, idesc_acc, (a.j > 0 || kk > 0) ? 1u : 0u);
      }
      __syncwarp();
    };
    // dK += dS^T . Q_chunk, and once per query tile dQ_t += dS_t . K_tile; releases what this chunk was the last to read
    auto issue_dk_dq = [&](const Walk& a) {
      const int slot = a.ts & 1;
      const int n_k = (a.valid_cols() + 15) >> 4;
      const int tq = a.tq();
      const uint32_t a_ds = DESC_LO_K + ds_lo + ((tq & 1) * 2 + (a.j & 1)) * (TILE_BYTES >> 4);
      const uint32_t b_q = DESC_LO_MN + q_lo + a.j * (CW * 128 >> 4);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < CW / 16; ++kk)
          if (kk < n_k) umma_ss(tmem + TM_DK, a_ds + 2 * kk, b_q + kk * 128, idesc_acc, (a.j > 0 || kk > 0) ? 1u : 0u);
        if (a.last_tile()) umma_commit(&qdo_empty[a.j]);
      }
      __syncwarp();
      if ((a.j & 1) || a.last_chunk()) {
        // contraction over the keys of this tile (16 per k-step)
        const int t = a.j >> 1;
        if (a.jt == 0 && t == 0) {
          PW(5, mbar_wait(dq_empty, (a.gi & 1) ^ 1));
          tc_fence_after();
        }
        const int n_kk = (a.valid_keys() + 15) >> 4;
        const uint32_t a_pair = DESC_LO_MN + ds_lo + (tq & 1) * (2 * TILE_BYTES >> 4);
        const uint32_t b_k = DESC_LO_MN + kv_lo + slot * (2 * TILE_BYTES >> 4);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            if (kk < n_kk)
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
    };

    // The score MMAs run up to two chunks ahead of the accumulate MMAs (two S^T slots): S^T/dP^T of chunk c+2 are
    // issued right behind dV of chunk c, ahead of its dK/dQ, so that the warpgroup that just finished chunk c finds
    // its next chunk ready.  Looking ahead into the next group is only possible for the Q/dO blocks that the
    // accumulate MMAs issued so far have released (chunks before `ac` of the group's last key tile).
    Walk sp, ac;
    sp.init(prob);
    ac = sp;
    auto may_issue = [&](int max_ahead) {
      if (sp.done(prob) || sp.c - ac.c > max_ahead) return false;
      if (sp.gi == ac.gi) return true;
      return sp.gi == ac.gi + 1 && ((ac.last_tile() && sp.j < ac.j) || sp.j >= ac.nchunks);
    };
    while (may_issue(1)) {
      issue_sp(sp);
      sp.next_chunk(prob);
    }
    while (!ac.done(prob)) {
      issue_dv(ac);
      if (may_issue(2)) {   // dV(ac) has been issued: its S^T/P^T slot may be overwritten by chunk ac + 2
        issue_sp(sp);
        sp.next_chunk(prob);
      }
      issue_dk_dq(ac);
      ac.next_chunk(prob);
      while (may_issue(1)) {
        issue_sp(sp);
        sp.next_chunk(prob);
      }
    }
    PROF_DUMP("issuer(qdo_full,kv_full,dp_free,pd_full,acc_empty,dq_empty)");
  } else if (warp < 10) {
    // ------------------------------------------------------------------------------------------------ compute
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                  // key row of the tile == TMEM lane
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    const float sl2 = scale * LOG2E;
    const float* sLall = reinterpret_cast<const float*>(smem + OFF_STAT);
    const float* sDall = sLall + STAT_N;
    const uint32_t sw = uint32_t(row & 7);
    Walk k;
    k.init(prob);
    PROF_DECL;
    while (!k.done(prob)) {
      if ((k.c & 1) != wg) {
        k.next_chunk(prob);
        continue;
      }
      const int valid = k.valid_cols();
      const int n_mma = (valid + 15) & ~15;
      const int kvalid = k.valid_keys();
      const int tq = k.tq();
      const bool active = quad * 32 < ((kvalid + 15) & ~15);   // warps past the last 16-key step have nothing to do
      const bool key_ok = row < kvalid;
      const uint32_t t_s = lane_addr + TM_S + (k.c & 1) * CW, t_dp = lane_addr + TM_DP;
      uint8_t* panel_row = smem + OFF_DS + ((tq & 1) * 2 + (k.j & 1)) * TILE_BYTES + row * 128;
      PW(0, mbar_wait(&qdo_full[k.j], (k.qpar >> k.j) & 1));                // this block's statistics are visible
      PW(2, mbar_wait(&sp_full[k.c & 1], (k.c >> 1) & 1));
#ifdef APLA_AFU_PROF
      const long long tc0 = clock64();
#endif
      tc_fence_after();
      if (active) {
        const int n_half = (n_mma + 31) >> 5;
#pragma unroll
        for (int hf = 0; hf < CW / 32; ++hf) {
          if (hf < n_half) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(t_s + hf * 32, sv);
            tmem_ld_32x32(t_dp + hf * 32, dv);
            tmem_ld_wait();
            if (hf == n_half - 1) {   // dP^T is in registers: the issuer may overwrite it with the next chunk's
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(dp_free);
            }
            const float* sL = sLall + k.j * CW + hf * 32;
            const float* sD = sDall + k.j * CW + hf * 32;
            uint32_t pp[16], pd[16];
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
            tmem_st_32x16(t_s + hf * 16, pp);
            if (hf == 0) PW(1, mbar_wait(&ds_empty[tq & 1], ((tq >> 1) & 1) ^ 1));   // panel pair free of older MMAs
            // dS^T row of this key: 32 queries = 64 B = four 16-byte units of the 128B-swizzled panel row; rows of
            // keys past the end of the sequence must contribute nothing to dQ
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 val = make_uint4(pd[4 * u], pd[4 * u + 1], pd[4 * u + 2], pd[4 * u + 3]);
              if (!key_ok) val = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(panel_row + ((uint32_t(hf * 4 + u) ^ sw) << 4)) = val;
            }
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(dp_free);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full[k.c & 1]);
#ifdef APLA_AFU_PROF
      prof[3] += clock64() - tc0;
      prof[4] += 1;
#endif
      k.next_chunk(prob);
    }
    if (warp == 2 || warp == 6) PROF_DUMP("compute(qdo_full,ds_empty,sp_full,busy,chunks)");
  } else if (warp == 14) {
    // ------------------------------------------------------------------------------------------------ statistics
    // lse (pre-multiplied by log2 e) and delta of every query of the group, gathered into registers ahead of time and
    // dropped into a block's slot as soon as the block is released; +inf / 0 past the end of the sequence masks the
    // padded columns (P = dS = 0).
    float* sL = reinterpret_cast<float*>(smem + OFF_STAT);
    float* sD = sL + STAT_N;
    Walk k;
    k.init(prob);
    while (!k.done(prob)) {
      float vl[STAT_N / 32], vd[STAT_N / 32];
#pragma unroll
      for (int u = 0; u < STAT_N / 32; ++u) {
        const int i = u * 32 + lane;
        const bool ok = i < k.n;
        const size_t idx = size_t(k.row_start + (ok ? i : 0)) * H + k.h;
        vl[u] = ok ? __ldg(lse + idx) * LOG2E : INFINITY;
        vd[u] = ok ? __ldg(delta + idx) : 0.f;
      }
      for (int j = 0; j < k.nchunks; ++j) {
        mbar_wait(&qdo_empty[j], ((k.qpar >> j) & 1) ^ 1);
#pragma unroll
        for (int u = 0; u < STAT_N / 32; ++u) {
          if (u >> 1 == j) {
            sL[u * 32 + lane] = vl[u];
            sD[u * 32 + lane] = vd[u];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&qdo_full[j]);
      }
      k.next_group(prob);
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16);
    Walk k;
    k.init(prob);
    PROF_DECL;
    while (!k.done(prob)) {
      PW(0, mbar_wait(acc_full, k.ts & 1));
      tc_fence_after();
      {
        // accumulators -> registers, release them to the issuer, then write
        const int r = k.jt * 128 + row;
        const bool ok = r < k.n;
        __nv_bfloat16* base = dqkv + size_t(k.row_start + (ok ? r : 0)) * (3 * D) + k.h * 64;
        uint32_t pk[32], pv[32];
        load_row64(lane_addr + TM_DK, scale, pk);     // dK = scale * dS^T Q
        load_row64(lane_addr + TM_DV, 1.0f, pv);      // dV = P^T dO
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        store_row64(pk, base + D, ok);
        store_row64(pv, base + 2 * D, ok);
      }
      if (k.last_tile()) {
        PW(1, mbar_wait(dq_full, k.gi & 1));
        tc_fence_after();
        for (int t = 0; t < k.ntiles; ++t) {
          const int r = t * 128 + row;
          const bool ok = r < k.n;
          __nv_bfloat16* base = dqkv + size_t(k.row_start + (ok ? r : 0)) * (3 * D) + k.h * 64;
          uint32_t pq[32];
          load_row64(lane_addr + TM_DQ + t * 64, scale, pq);   // dQ = scale * dS K
          if (t == k.ntiles - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_empty);
          }
          store_row64(pq, base, ok);
        }
      }
      // advance to the next key tile
      const int ts0 = k.ts;
      while (!k.done(prob) && k.ts == ts0) k.next_chunk(prob);
    }
    if (warp == 10) PROF_DUMP("epilogue(acc_full,dq_full)");
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, TM_COLS);
  }
}

}  // namespace afu

bool attn_fused_supported(int max_seqlen) { return max_seqlen > 0 && max_seqlen <= afu::NP_MAX; }

int attn_bwd_fused(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                   const int* cu_seqlens, int num_seqs, int max_seqlen, int total_tokens, int H, float scale,
                   cudaStream_t stream) {
  using namespace afu;
  APLA_CHECK(attn_fused_supported(max_seqlen), "attn_bwd_fused: max_seqlen %d exceeds the resident limit", max_seqlen);
  const int D = H * 64;
  const uint64_t T = total_tokens;
  Maps m;
  if (int rc = make_tmap_2d(&m.qkv64, qkv, 2, T, 3 * D, 3 * D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.qkv16, qkv, 2, T, 3 * D, 3 * D, 16, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.do64, dout, 2, T, D, D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.do16, dout, 2, T, D, D, 16, 64, true)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    APLA_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int G = num_seqs * H;
  const int grid = G < sm_count() ? G : sm_count();
  attn_bwd_fused_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(m, lse, delta, reinterpret_cast<__nv_bfloat16*>(dqkv),
                                                               cu_seqlens, max_seqlen, H, G, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
