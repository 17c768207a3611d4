// Sequence-resident persistent tcgen05 / TMEM attention forward for short sequences (N <= 272 tokens, head_dim 64).
//
// One persistent CTA per SM walks over (sequence, head) groups.  A group's K and V rows stay in shared memory
// (double-buffered by group); its 128-row Q tiles are the work items.  Items alternate between two independent
// streams, each owned by one compute warpgroup with its own TMEM: two 96-column score slots and one 64-column
// output accumulator.  Per (item, 96-key chunk):
//     S   : S = Q_tile . K_chunk^T                          (smem x smem -> TMEM slot, fp32)
//     WG  : row max -> P = exp2(S*scale*log2e - m) -> bf16, written over S in place; running row sum in registers.
//           The maximum is lazy: it is only raised (and O / l rescaled in TMEM) when a row's maximum grows by more
//           than 2^8, so after the first chunk the rescale practically never runs.
//     PV  : O += P . V_chunk                                (A from TMEM, B = the resident V rows read MN-major)
// The issuer keeps every stream two chunks ahead with the score MMAs (S of chunk k+2 is issued right behind PV of
// chunk k), so a warpgroup that finishes a chunk finds the next scores waiting and the exponentials of one stream
// overlap the MMAs of both.
//
// Warp roles (16 warps): 0 = TMA producer, 1 / 2 = MMA issuer of stream 0 / 1 (one thread each), 3 idle, 4-7 / 8-11 =
// the two compute warpgroups (softmax + epilogue: O / l -> bf16 -> global, lse), 12-15 = helpers: the 257th (and
// 258th) query row of a sequence is computed on CUDA cores instead of occupying a 128-row tile of its own.
// TMEM (512 columns): stream w: S slots at w*256 + {0, 96}, O at w*256 + 192.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace afs {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int NP_MAX = 272;
constexpr int CW = 96;                        // keys per chunk
constexpr int kThreads = 512;
constexpr int SIMT_ROWS_MAX = 2;               // query rows past 256 handled by the helper warps instead of a whole tile
constexpr uint32_t RES_BYTES = NP_MAX * 128;  // resident K or V of one group
constexpr uint32_t QT_BYTES = 128 * 128;      // one Q tile
constexpr uint32_t OFF_KV = 0;                // [2 groups][K, V]
constexpr uint32_t OFF_Q = 4 * RES_BYTES;     // [4 slots]
constexpr uint32_t OFF_BAR = OFF_Q + 4 * QT_BYTES;
constexpr uint32_t OFF_HELP = OFF_BAR + 256;    // helper scratch: p[384], partial O[4][64], reductions[8] (floats)
constexpr uint32_t SMEM_BYTES = OFF_HELP + (384 + 256 + 8) * 4 + 1024;
constexpr uint32_t TM_STREAM = 256, TM_O = 192, TM_COLS = 512;

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

constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_LO_K = (16u >> 4) << 16;
constexpr uint32_t DESC_LO_MN = (16384u >> 4) << 16;

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

// -DAPLA_AFS_PROF: CTA 0 prints the cycles each role spent blocked at each kind of wait (development aid)
#ifdef APLA_AFS_PROF
#define PW(id, ...)                      \
  do {                                   \
    const long long _t0 = clock64();     \
    __VA_ARGS__;                         \
    prof[id] += clock64() - _t0;         \
  } while (0)
#define PROF_DECL long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = clock64()
#define PROF_DUMP(role)                                                                                          \
  if (blockIdx.x == 0 && lane == 0)                                                                              \
  printf("afs %s: total %lld  w0 %lld w1 %lld w2 %lld w3 %lld w4 %lld w5 %lld w6 %lld w7 %lld\n", role,             \
         clock64() - prof_t0, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7])
// event trace of CTA 0 kept in shared memory (a global-memory trace costs a DRAM round trip per event)
#define TR(role, ev)                                                                   \
  do {                                                                                 \
    if (blockIdx.x == 0 && lane == 0 && g_trace_n[role] < 160)                         \
      g_trace[role][g_trace_n[role]++] = ((clock64() - tr_t0) << 8) | (long long)(ev); \
  } while (0)
#else
#define TR(role, ev)
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

// Walk over this CTA's (group, q tile) items; ii = items done, gi = groups done.
struct Walk {
  int g, gi, ii, t;
  int row_start, n, h, ntiles, nchunks, simt_rows;
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
    // a query tile that would hold only one or two rows is not worth 128 lanes: those rows go to the helper warps
    simt_rows = (n > 256 && n - 256 <= SIMT_ROWS_MAX) ? n - 256 : 0;
    ntiles = (n - simt_rows + 127) >> 7;
    nchunks = (n + CW - 1) / CW;
  }
  __device__ __forceinline__ void init(const Problem& p) {
    g = blockIdx.x;
    gi = ii = t = 0;
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
    ++ii;
    if (++t == ntiles) next_group(p);
  }
  __device__ __forceinline__ int valid_cols(int j) const { return min(CW, n - j * CW); }
};

// One chunk of one row.  The compute warpgroups run with 184 registers (setmaxnreg; the producer / issuer / helper
// warpgroups give theirs up), so all CW = 96 scores of the row are loaded from TMEM ONCE, by three back-to-back
// tcgen05.ld behind a single wait, and stay in registers for the row maximum and for P = exp2(S*sl2 - ms), which is
// packed to bf16 and written over S in place.  (The first version made two passes over TMEM with two 32-column
// pieces in flight; its four exposed TMEM round trips per chunk were more than half of the softmax time.)
// FULL: all CW columns are valid keys (no masking).
template <bool FULL>
__device__ __forceinline__ void chunk_load(uint32_t t_s, int n_mma, uint32_t (&a)[32], uint32_t (&b)[32],
                                           uint32_t (&c)[32]) {
  tmem_ld_32x32(t_s, a);
  if (FULL || n_mma > 32) tmem_ld_32x32(t_s + 32, b);
  if (FULL || n_mma > 64) tmem_ld_32x32(t_s + 64, c);
  tmem_ld_wait();
}
template <bool FULL>
__device__ __forceinline__ float piece_max(const uint32_t (&v)[32], int valid, int base) {
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    if (FULL) {
      m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
      m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
    } else {
      if (base + i < valid) m0 = fmaxf(m0, __uint_as_float(v[i]));
      if (base + i + 1 < valid) m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
      if (base + i + 2 < valid) m0 = fmaxf(m0, __uint_as_float(v[i + 2]));
      if (base + i + 3 < valid) m1 = fmaxf(m1, __uint_as_float(v[i + 3]));
    }
  }
  return fmaxf(m0, m1);
}
template <bool FULL>
__device__ __forceinline__ float chunk_row_max(const uint32_t (&a)[32], const uint32_t (&b)[32], const uint32_t (&c)[32],
                                               int valid, int n_mma) {
  // (a piece whose 32 columns are all valid takes the unmasked code even in a partial chunk)
  float m = (FULL || valid >= 32) ? piece_max<true>(a, valid, 0) : piece_max<false>(a, valid, 0);
  if (FULL || n_mma > 32) m = fmaxf(m, (FULL || valid >= 64) ? piece_max<true>(b, valid, 32) : piece_max<false>(b, valid, 32));
  if (FULL || n_mma > 64) m = fmaxf(m, FULL ? piece_max<true>(c, valid, 64) : piece_max<false>(c, valid, 64));
  return m;
}
template <bool FULL>
__device__ __forceinline__ float piece_exp(uint32_t t_p, const uint32_t (&v)[32], int valid, int base, float sl2, float ms) {
  uint32_t pk[16];
  float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float p0 = exp2f(fmaf(__uint_as_float(v[2 * i]), sl2, -ms));
    float p1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -ms));
    if (!FULL) {
      p0 = base + 2 * i < valid ? p0 : 0.f;
      p1 = base + 2 * i + 1 < valid ? p1 : 0.f;
    }
    rs0 += p0;
    rs1 += p1;
    pk[i] = pack_bf16(p0, p1);
  }
  tmem_st_32x16(t_p, pk);
  return rs0 + rs1;
}
// (piece pc of P lands in columns [16 pc, 16 pc + 16) of the slot; every score is already in registers)
template <bool FULL>
__device__ __forceinline__ float chunk_exp(uint32_t t_s, const uint32_t (&a)[32], const uint32_t (&b)[32],
                                           const uint32_t (&c)[32], int valid, int n_mma, float sl2, float ms) {
  float rs = (FULL || valid >= 32) ? piece_exp<true>(t_s, a, valid, 0, sl2, ms) : piece_exp<false>(t_s, a, valid, 0, sl2, ms);
  if (FULL || n_mma > 32)
    rs += (FULL || valid >= 64) ? piece_exp<true>(t_s + 16, b, valid, 32, sl2, ms) : piece_exp<false>(t_s + 16, b, valid, 32, sl2, ms);
  if (FULL || n_mma > 64) rs += FULL ? piece_exp<true>(t_s + 32, c, valid, 64, sl2, ms) : piece_exp<false>(t_s + 32, c, valid, 64, sl2, ms);
  return rs;
}

struct Maps {
  CUtensorMap q128, kv64, kv16;   // 128-row boxes for Q tiles, 64- and 16-row boxes for the resident K / V rows
  CUtensorMap o32;                // 32-row x 64-column boxes of the output (one warp's rows of one head)
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_sr_kernel(const __grid_constant__ Maps maps, const __nv_bfloat16* __restrict__ qkv,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                   const int* __restrict__ cu_seqlens, int n_fixed, int H, int G, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;         // [2]
  uint64_t* kv_empty = bars + 2;    // [2]
  uint64_t* q_full = bars + 4;      // [4]
  uint64_t* q_empty = bars + 8;     // [4]
  uint64_t* s_full = bars + 12;     // [2 streams][2 slots]
  uint64_t* p_full = bars + 16;     // [2 streams][2 slots]
  uint64_t* o_full = bars + 20;     // [2 streams][2]: the PV of the stream's chunk k completes o_full[w][k & 1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const Problem prob{cu_seqlens, n_fixed, H, G};
  const int D = H * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q128);
    tma_prefetch_desc(&maps.kv64);
    tma_prefetch_desc(&maps.kv16);
    tma_prefetch_desc(&maps.o32);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 3);   // both issuers and the helper warps
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 5);   // the issuer's last score MMA + the four warps of the item's compute warpgroup
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
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
#ifdef APLA_AFS_PROF
  __shared__ long long tr_t0_s;
  __shared__ long long g_trace[4][160];
  __shared__ int g_trace_n[4];
  if (threadIdx.x == 0) {
    tr_t0_s = clock64();
    for (int i = 0; i < 4; ++i) g_trace_n[i] = 0;
  }
  __syncthreads();
  const long long tr_t0 = tr_t0_s;
#endif

  // Register budget per warpgroup (65536 = 128 x 64 + 256 x 184 + 128 x 80; setmaxnreg at the head of every role
  // branch): the softmax keeps a whole 96-score row live.
  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ producer
    reg_dealloc<64>();
    // (the whole warp walks the schedule -- Walk uses warp shuffles -- and lane 0 issues the copies)
    Walk k;
    k.init(prob);
    while (!k.done(prob)) {
      const int gb = k.gi & 1;
      const int np = (k.n + 15) & ~15;
      if (lane == 0) {
        mbar_wait(&kv_empty[gb], ((k.gi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[gb], 2u * np * 128u);
        uint8_t* dk = smem + OFF_KV + gb * 2 * RES_BYTES;
        uint8_t* dv = dk + RES_BYTES;
        int r = 0;
        for (; r + 64 <= np; r += 64) {
          tma_load_2d(dk + r * 128, &maps.kv64, &kv_full[gb], D + k.h * 64, k.row_start + r);
          tma_load_2d(dv + r * 128, &maps.kv64, &kv_full[gb], 2 * D + k.h * 64, k.row_start + r);
        }
        for (; r < np; r += 16) {
          tma_load_2d(dk + r * 128, &maps.kv16, &kv_full[gb], D + k.h * 64, k.row_start + r);
          tma_load_2d(dv + r * 128, &maps.kv16, &kv_full[gb], 2 * D + k.h * 64, k.row_start + r);
        }
      }
      __syncwarp();
      const int gi0 = k.gi;
      do {   // the Q tiles of this group
        const int qs = k.ii & 3;
        if (lane == 0) {
          mbar_wait(&q_empty[qs], ((k.ii >> 2) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qs], QT_BYTES);
          tma_load_2d(smem + OFF_Q + qs * QT_BYTES, &maps.q128, &q_full[qs], k.h * 64, k.row_start + k.t * 128);
        }
        __syncwarp();
        k.next_item(prob);
      } while (k.gi == gi0);
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------------------------------------ MMA issuers
    // One issuer warp per stream (items of even / odd index); chunk k of the stream's chunk sequence lives in S slot
    // k & 1.  Order: S(0), S(1), then for every k: [P(k) ready] PV(k), S(k+2).  A score MMA whose operands have not
    // landed yet is not waited for (the rows may be held up by a buffer that only this stream's next PV releases).
    reg_dealloc<64>();
    const int w = warp - 1;
    const bool leader = elect_one();
    PROF_DECL;
    const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t kv_lo = smem_u32(smem + OFF_KV) >> 4, q_lo = smem_u32(smem + OFF_Q) >> 4;
    const uint32_t t_base = tmem + w * TM_STREAM;
    Walk sc, ac;          // cursors of the score MMAs and of the PV MMAs (all items are walked, others' are skipped)
    int sj = 0, aj = 0;   // chunk within the item
    int sk = 0, ak = 0;   // chunks issued so far in this stream
    int released = 0;     // groups whose K/V buffer this stream has released
    auto skip_to_parity = [&](Walk& c) {
      while (!c.done(prob) && (c.ii & 1) != w) c.next_item(prob);
    };
    // Both streams release every group (two arrivals per phase of kv_empty), also the groups in which they own no
    // item.  An arrival must not leak into the previous phase of the same buffer, so that phase is awaited first.
    auto release_groups = [&](int upto) {
      for (; released < upto; ++released) {
        const int gb = released & 1, ph = released >> 1;
        if (ph > 0) mbar_wait(&kv_empty[gb], (ph - 1) & 1);
        if (leader) umma_commit(&kv_empty[gb]);
        __syncwarp();
      }
    };
    auto issue_s = [&]() {
      const int gb = sc.gi & 1, qs = sc.ii & 3;
      const int n_mma = (sc.valid_cols(sj) + 15) & ~15;
      const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
      const uint32_t a_q = DESC_LO_K + q_lo + qs * (QT_BYTES >> 4);
      const uint32_t b_k = DESC_LO_K + kv_lo + gb * (2 * RES_BYTES >> 4) + sj * (CW * 128 >> 4);
      const uint32_t t_s = t_base + (sk & 1) * CW;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(t_s, a_q + 2 * kk, b_k + 2 * kk, idesc_s, kk > 0);
        umma_commit(&s_full[w * 2 + (sk & 1)]);
        if (sj == sc.nchunks - 1) umma_commit(&q_empty[qs]);
      }
      __syncwarp();
      ++sk;
      if (++sj == sc.nchunks) {
        sj = 0;
        sc.next_item(prob);
        skip_to_parity(sc);
      }
    };
    auto issue_pv = [&]() {
      const int gb = ac.gi & 1;
      const int n_k = (ac.valid_cols(aj) + 15) >> 4;
      const uint32_t b_v = DESC_LO_MN + kv_lo + gb * (2 * RES_BYTES >> 4) + (RES_BYTES >> 4) + aj * (CW * 128 >> 4);
      const uint32_t t_p = t_base + (ak & 1) * CW, t_o = t_base + TM_O;
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < CW / 16; ++kk)
          if (kk < n_k) umma_ts(t_o, t_p + kk * 8, b_v + kk * 128, idesc_o, (aj > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&o_full[w * 2 + (ak & 1)]);
      }
      __syncwarp();
      ++ak;
      if (++aj == ac.nchunks) {
        aj = 0;
        ac.next_item(prob);
        skip_to_parity(ac);
        release_groups(ac.gi);   // every group this stream has left behind
      }
    };
    sc.init(prob);
    skip_to_parity(sc);
    ac = sc;
    release_groups(ac.gi);
    uint32_t spins = 0;
    while (!ac.done(prob)) {
      bool progressed = false;
      if (!sc.done(prob) && sk - ak < 2) {
        bool ready = true;
        if (sj == 0) {
          ready = mbar_test(&q_full[sc.ii & 3], (sc.ii >> 2) & 1) && mbar_test(&kv_full[sc.gi & 1], (sc.gi >> 1) & 1);
          ready = __all_sync(0xffffffffu, ready) != 0;
          if (ready) tc_fence_after();
        }
        if (ready) {
          TR(w, 0x10 + (sk & 15));
          issue_s();
          TR(w, 0x20);
          progressed = true;
        }
      }
      if (ak < sk) {
        // try_wait may suspend the thread until the phase completes: only use it when no score MMA is pending
        const bool may_block = sc.done(prob) || sk - ak >= 2;
        bool ready;
        if (may_block) PW(0, ready = mbar_try_wait(&p_full[w * 2 + (ak & 1)], (ak >> 1) & 1));
        else ready = mbar_test(&p_full[w * 2 + (ak & 1)], (ak >> 1) & 1);
        if (__all_sync(0xffffffffu, ready)) {
          tc_fence_after();
          TR(w, 0x30 + (ak & 15));
          issue_pv();
          TR(w, 0x40);
          progressed = true;
        }
      }
      if (progressed) spins = 0;
      else if (++spins > (1u << 24)) __trap();
    }
    PROF_DUMP("issuer(p_full)");
  } else if (warp == 3) {
    reg_dealloc<64>();   // idle warp of the first warpgroup: every warp of a warpgroup executes the same setmaxnreg
  } else if (warp >= 12) {
    // ------------------------------------------------------------------------------------------------ helpers
    reg_dealloc<80>();
    // Query rows 256.. of a 257/258-token sequence on CUDA cores (4 warps, fp32): scores against the resident K,
    // softmax, P.V against the resident V.  The helpers also take part in releasing every group's K/V buffer.
    // (explicit ld.shared / st.shared by 32-bit address: the pointer arithmetic on the aligned dynamic buffer otherwise
    //  compiles to generic LD / ST, which take the slow local/global path -- see ptx.cuh)
    const int e = threadIdx.x - 384, hw = warp - 12;
    const uint32_t sb = smem_u32(smem);
    const uint32_t hp = sb + OFF_HELP, hpart = hp + 384 * 4, hred = hpart + 256 * 4;
    const float sl2 = scale * LOG2E;
    Walk k;
    k.init(prob);
    while (!k.done(prob)) {
      const int gb = k.gi & 1;
      mbar_wait(&kv_full[gb], (k.gi >> 1) & 1);
      const uint32_t sK = sb + OFF_KV + gb * 2 * RES_BYTES;
      const uint32_t sV = sK + RES_BYTES;
      for (int r = 0; r < k.simt_rows; ++r) {
        const int T = 256 + r;
        const uint4* qg = reinterpret_cast<const uint4*>(qkv + size_t(k.row_start + T) * (3 * D) + k.h * 64);
        uint4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = __ldg(qg + u);
        // scores of keys e, e + 128, e + 256
        float sc[3], mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int key = e + 128 * i;
          sc[i] = -INFINITY;
          if (key < k.n) {
            const uint32_t kr = sK + key * 128;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint4 kv = lds_u4(kr + ((u ^ (key & 7)) << 4));
              const uint32_t kw[4] = {kv.x, kv.y, kv.z, kv.w}, qw[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                a0 = fmaf(bf16_lo(qw[x]), bf16_lo(kw[x]), a0);
                a1 = fmaf(bf16_hi(qw[x]), bf16_hi(kw[x]), a1);
              }
            }
            sc[i] = a0 + a1;
            mx = fmaxf(mx, sc[i]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) sts_f32(hred + hw * 4, mx);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {
          const float4 m4 = lds_f4(hred);
          mx = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int key = e + 128 * i;
          const float pv = key < k.n ? exp2f((sc[i] - mx) * sl2) : 0.f;
          if (key < 384) sts_f32(hp + key * 4, pv);
          sum += pv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) sts_f32(hred + (4 + hw) * 4, sum);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const float4 l4 = lds_f4(hred + 16);
        const float l = (l4.x + l4.y) + (l4.z + l4.w);
        // O[T, 2 dp .. 2 dp + 1] over the keys part, part + 4, ...
        const int dp = e & 31, part = e >> 5;
        float o0 = 0.f, o1 = 0.f;
#pragma unroll 4
        for (int key = part; key < k.n; key += 4) {
          const float pv = lds_f32(hp + key * 4);
          const uint32_t vv = lds_u32(sV + key * 128 + (((dp >> 2) ^ (key & 7)) << 4) + (dp & 3) * 4);
          o0 = fmaf(pv, bf16_lo(vv), o0);
          o1 = fmaf(pv, bf16_hi(vv), o1);
        }
        sts_f32(hpart + (part * 64 + 2 * dp) * 4, o0);
        sts_f32(hpart + (part * 64 + 2 * dp + 1) * 4, o1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (e < 64) {
          const float o = (lds_f32(hpart + e * 4) + lds_f32(hpart + (64 + e) * 4)) +
                          (lds_f32(hpart + (128 + e) * 4) + lds_f32(hpart + (192 + e) * 4));
          out[size_t(k.row_start + T) * D + k.h * 64 + e] = __float2bfloat16_rn(o / l);
        }
        if (e == 0) lse[size_t(k.row_start + T) * H + k.h] = mx * scale + logf(l);
        asm volatile("bar.sync 1, 128;" ::: "memory");   // scratch is reused by the next row / group
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");     // every helper thread is done with this group's K / V
      if (e == 0) mbar_arrive(&kv_empty[gb]);
      k.next_group(prob);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------------------ compute
    reg_alloc<184>();
    const int w = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                  // row of the q tile == TMEM lane
    const uint32_t lane_addr = tmem + (uint32_t(quad * 32) << 16) + w * TM_STREAM;
    const float sl2 = scale * LOG2E;
    Walk k;
    k.init(prob);
    PROF_DECL;
    int kc = 0;      // chunks of this stream processed so far
    int o_seen = 0;  // PVs of this stream whose completion this warp has consumed, in order.  PV(i) completes barrier
                     // i & 1; a parity wait cannot tell "two phases behind" from "done", so PV(i) is consumed before
                     // chunk i + 2 is handed to the issuer -- by then it is two chunks old and the wait is free.
    auto consume_pv = [&](int upto) {
      for (; o_seen < upto; ++o_seen) mbar_wait(&o_full[w * 2 + (o_seen & 1)], (o_seen >> 1) & 1);
    };
    while (!k.done(prob)) {
      if ((k.ii & 1) != w) {
        k.next_item(prob);
        continue;
      }
      float m_used = -INFINITY, l_run = 0.f;
      const bool tile_active = k.t * 128 + quad * 32 < k.n;   // warps whose 32 rows all lie past the end only sync
      for (int j = 0; j < k.nchunks; ++j, ++kc) {
        const int valid = k.valid_cols(j);
        const int n_mma = (valid + 15) & ~15;
        const uint32_t t_s = lane_addr + (kc & 1) * CW;
        if (quad == 0) TR(2 + w, 0x50 + (kc & 15));
        PW(0, mbar_wait(&s_full[w * 2 + (kc & 1)], (kc >> 1) & 1));
        if (quad == 0) TR(2 + w, 0x60);
#ifdef APLA_AFS_PROF
        const long long tc0 = clock64();
#endif
        tc_fence_after();
        if (tile_active) {
          // pass 1: row maximum of the chunk
          const bool full = valid == CW;
          uint32_t va[32], vb[32], vc[32];
          if (full) chunk_load<true>(t_s, n_mma, va, vb, vc);
          else chunk_load<false>(t_s, n_mma, va, vb, vc);
          const float mx = full ? chunk_row_max<true>(va, vb, vc, valid, n_mma) : chunk_row_max<false>(va, vb, vc, valid, n_mma);
          if (j == 0) {
            m_used = mx;
          } else {
            const bool need = (mx - m_used) * sl2 > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              // rare: raise the reference maximum of the rows that need it and rescale their O / l in place
              const float f = need ? exp2f((m_used - mx) * sl2) : 1.0f;
              if (need) m_used = mx;
              l_run *= f;
              consume_pv(kc);
              tc_fence_after();
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint32_t ov[16];
                tmem_ld_32x16(lane_addr + TM_O + c * 16, ov);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
                tmem_st_32x16(lane_addr + TM_O + c * 16, ov);
              }
            }
          }
          // pass 2: the exponentials
          const float ms = m_used * sl2;
          const float rs = full ? chunk_exp<true>(t_s, va, vb, vc, valid, n_mma, sl2, ms)
                                : chunk_exp<false>(t_s, va, vb, vc, valid, n_mma, sl2, ms);
          l_run += rs;
          tmem_st_wait();
        }
#ifdef APLA_AFS_PROF
        prof[3] += clock64() - tc0;
        prof[4] += 1;
#endif
        PW(1, consume_pv(kc - 1));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[w * 2 + (kc & 1)]);
        if (quad == 0) TR(2 + w, 0x70);
      }
      // epilogue: O / l -> bf16 -> global, lse
      PW(2, consume_pv(kc));
      if (quad == 0) TR(2 + w, 0x81);
      tc_fence_after();
      // O / l -> bf16.  A warp whose 32 rows all exist stages them (128B-swizzled) in ITS rows of the item's Q tile slot
      // -- every MMA that read the slot has completed -- and writes them with one TMA store; per-thread 16-byte stores
      // to 32 different rows cost ~2500 cycles per item here.  The slot goes back to the producer once the store has
      // read it.  Warps with a partial or empty row range keep the per-row stores.
      if (tile_active) {
        const int r = k.t * 128 + row;
        const bool store = r < k.n;
        const bool whole = k.t * 128 + quad * 32 + 32 <= k.n;   // warp-uniform
        const float inv = 1.f / l_run;
        uint4* dst = reinterpret_cast<uint4*>(out + size_t(k.row_start + (store ? r : 0)) * D + k.h * 64);
        const uint32_t stage = smem_u32(smem + OFF_Q + (k.ii & 3) * QT_BYTES + quad * 4096);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32(lane_addr + TM_O + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(ov[8 * i + e]) * inv;
            const uint4 pk =
                make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            if (whole) {
              const uint32_t a = stage + lane * 128 + (((c * 4 + i) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w)
                           : "memory");
            } else if (store) {
              dst[c * 4 + i] = pk;
            }
          }
        }
        if (quad == 0) TR(2 + w, 0x82);
        if (store) lse[size_t(k.row_start + r) * H + k.h] = m_used * scale + logf(l_run);
        if (whole) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&maps.o32, smem + OFF_Q + (k.ii & 3) * QT_BYTES + quad * 4096, k.h * 64,
                         k.row_start + k.t * 128 + quad * 32);
            tma_store_commit();
            tma_store_wait_read<0>();
          }
        }
      }
      __syncwarp();
      if (quad == 0) TR(2 + w, 0x83);
      if (lane == 0) mbar_arrive(&q_empty[k.ii & 3]);
      tc_fence_before();
      if (quad == 0) TR(2 + w, 0x80);
      k.next_item(prob);
    }
    if (lane == 0) tma_store_wait<0>();
    if (quad == 2) PROF_DUMP("compute(s_full, o_full in-chunk, o_full epilogue, busy, chunks)");
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, TM_COLS);
  }
#ifdef APLA_AFS_PROF
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int r = 0; r < 4; ++r) {
      printf("afstrace role %d:", r);
      for (int i = 0; i < g_trace_n[r] && i < 160; ++i) printf(" %llx@%lld", g_trace[r][i] & 255, g_trace[r][i] >> 8);
      printf("\n");
    }
  }
#endif
}

}  // namespace afs

int attn_fwd_sr(const void* qkv, void* out, float* lse, const int* cu_seqlens, int num_seqs, int max_seqlen,
                int total_tokens, int H, float scale, cudaStream_t stream) {
  using namespace afs;
  APLA_CHECK(max_seqlen > 0 && max_seqlen <= NP_MAX, "attn_fwd_sr: max_seqlen %d exceeds the resident limit", max_seqlen);
  const int D = H * 64;
  const uint64_t T = total_tokens;
  Maps m;
  if (int rc = make_tmap_2d(&m.q128, qkv, 2, T, 3 * D, 3 * D, 128, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.kv64, qkv, 2, T, 3 * D, 3 * D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.kv16, qkv, 2, T, 3 * D, 3 * D, 16, 64, true)) return rc;
  if (int rc = make_tmap_2d(&m.o32, out, 2, T, D, D, 32, 64, true)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    APLA_CUDA(cudaFuncSetAttribute(attn_fwd_sr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int G = num_seqs * H;
  const int grid = G < sm_count() ? G : sm_count();
  attn_fwd_sr_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(m, reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                            reinterpret_cast<__nv_bfloat16*>(out), lse, cu_seqlens,
                                                            max_seqlen, H, G, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
