// Causal GQA flash attention for sm_100a, written against the packed qkv activation [B, S, (Hq + 2 Hkv) * 128].
//
// Forward (attn_fwd_kernel): one CTA = one 128-row query block of one query head.
//   * TMA (128B swizzle) stages Q once and K / V tiles of 128 keys in two-deep rings, straight out of the packed activation
//     (one tensor map, column coordinate = head * 128) — no q/k/v split or transpose copies;
//   * S = Q K^T and O_tile = P V run on the 5th-gen tensor cores: tcgen05.mma issued by one thread, S and O_tile accumulate in
//     TMEM (2 x 128 + 2 x 128 columns), P is written back to TMEM by the softmax warps as packed bf16 over the S columns and fed
//     to the second MMA as its A operand *from TMEM* (no shared-memory round trip), V is the MN-major B operand;
//   * 4 softmax warps, one thread per query row (tcgen05.ld 32x32b: a thread owns a row, so row max / row sum need no shuffles):
//     two passes over the S tile in TMEM (max, then exp2 + pack), online softmax with the running output in registers
//     (O = O * alpha + O_tile, the rescale never touches TMEM); QK^T of tile j+1 is issued before softmax(j) so the tensor core
//     works under the softmax, and the accumulation of O_tile(j-1) is deferred until after P(j) is handed to the MMA warp;
//   * only the diagonal tile is masked; heavy (late) query blocks are scheduled first; the Hq/Hkv query heads that share a KV head
//     are adjacent in launch order so their K/V tiles hit in L2;
//   * epilogue: O / l -> bf16 -> swizzled smem -> TMA store into [B, S, Hq * 128]; log-sum-exp (natural log) saved for backward.
//
// Parity: the reference calls library attention (aten SDPA / flash_attn: legacy/vescale/dtensor/ops/matrix_ops.py:278-470); this is
// the hand-written Blackwell replacement.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "gemm_sm100.cuh"

using namespace vb;

namespace {

constexpr int kAQ = 128;             // query rows per CTA (UMMA M)
constexpr int kAKV = 128;            // keys per tile
constexpr int kAD = 128;             // head dim
constexpr int kAHalf = 128 * 64 * 2;  // one [128 x 64] bf16 TMA box (128-byte rows, 128B swizzle)
constexpr int kATile = 2 * kAHalf;    // a [128 x 128] operand tile = two boxes side by side
constexpr int kAThreads = 256;        // w0 TMA, w1 MMA, w2 TMEM alloc, w3 spare, w4-7 softmax / epilogue
constexpr uint32_t kColS = 0, kColOT = 256;  // TMEM columns: S0 S1 | OT0 OT1 (128 each); P_b aliases the first 64 columns of S_b

VB_DEVICE void tmem_st_32x32_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
      "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
VB_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 64 columns in ONE request (one wait instead of two: the softmax warps are bound by tcgen05.ld round trips)
VB_DEVICE void tmem_ld_32x64(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
        "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
        "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
        "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
}
VB_DEVICE void tmem_st_32x64(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,"
      "%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
      "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]),
      "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]), "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]),
      "r"(r[53]), "r"(r[54]), "r"(r[55]), "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63])
      : "memory");
}
VB_DEVICE void tmem_st_32x32_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
      "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (P, packed bf16, one query row per lane) never leaves tensor memory
VB_DEVICE void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major operand made of [rows x 64] boxes `lbo_bytes` apart (see make_sw128_desc_mn: there the boxes have 64 rows = 8 KB)
VB_DEVICE uint64_t make_sw128_desc_mn_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AttnBars {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full[2], p_full[2], s_free[2], ot_full[2], ot_free[2];
  uint32_t tmem_holder;
  uint32_t pad;
};

__global__ void __launch_bounds__(kAThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o, float* __restrict__ lse, int B, int S, int Hq, int Hkv,
                float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kATile;
  uint8_t* sV = smem + 3 * kATile;
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + 5 * kATile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = Hq / Hkv, nq = S / kAQ;
  int idx = blockIdx.x;
  const int g = idx % G;
  idx /= G;
  const int qblk = nq - 1 - idx % nq;  // heavy query blocks first
  idx /= nq;
  const int kvh = idx % Hkv, b = idx / Hkv;
  const int h = kvh * G + g;
  const int n_tiles = qblk + 1;  // causal: keys [0, (qblk + 1) * 128)
  const int row0 = b * S + qblk * kAQ;
  const int col_q = h * kAD, col_k = (Hq + kvh) * kAD, col_v = (Hq + Hkv + kvh) * kAD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_o);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->k_full[s], 1);
      mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1);
      mbar_init(&bars->v_empty[s], 1);
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], 4);
      mbar_init(&bars->s_free[s], 1);
      mbar_init(&bars->ot_full[s], 1);
      mbar_init(&bars->ot_free[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(&bars->q_full, kATile);
      tma_load_2d(sQ, &tm_qkv, &bars->q_full, col_q, row0);
      tma_load_2d(sQ + kAHalf, &tm_qkv, &bars->q_full, col_q + 64, row0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int krow = b * S + j * kAKV;
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->k_full[s], kATile);
        tma_load_2d(sK + s * kATile, &tm_qkv, &bars->k_full[s], col_k, krow);
        tma_load_2d(sK + s * kATile + kAHalf, &tm_qkv, &bars->k_full[s], col_k + 64, krow);
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->v_full[s], kATile);
        tma_load_2d(sV + s * kATile, &tm_qkv, &bars->v_full[s], col_v, krow);
        tma_load_2d(sV + s * kATile + kAHalf, &tm_qkv, &bars->v_full[s], col_v + 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16_major(kAQ, kAKV, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(kAQ, kAD, false, true);
      mbar_wait(&bars->q_full, 0);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bars->k_full[s], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&bars->s_free[s], ((j >> 1) - 1) & 1);  // P(j-2) has been consumed by its PV MMA
        tc_fence_after();
        const uint32_t d = tmem + kColS + s * 128;
#pragma unroll
        for (int k = 0; k < kAD / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sQ + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          const uint64_t bb = make_sw128_desc(smem_u32(sK + s * kATile + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(d, a, bb, idesc_qk, k ? 1u : 0u);
        }
        umma_commit(&bars->k_empty[s]);
        umma_commit(&bars->s_full[s]);
      };
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);  // the tensor core computes S(j+1) while the softmax warps work on S(j)
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&bars->p_full[s], ph);
        mbar_wait(&bars->v_full[s], ph);
        if (j >= 2) mbar_wait(&bars->ot_free[s], ((j >> 1) - 1) & 1);  // O_tile(j-2) has been folded into the running output
        tc_fence_after();
        const uint32_t d = tmem + kColOT + s * 128;
        const uint32_t p = tmem + kColS + s * 128;
        const uint64_t vdesc = make_sw128_desc_mn_lbo(smem_u32(sV + s * kATile), kAHalf);
#pragma unroll
        for (int k = 0; k < kAKV / 16; ++k) umma_bf16_ts(d, p + k * 8, vdesc + (uint64_t)(k * 128), idesc_pv, k ? 1u : 0u);
        umma_commit(&bars->v_empty[s]);
        umma_commit(&bars->s_free[s]);
        umma_commit(&bars->ot_full[s]);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + running output: one thread per query row =====================
    const int qd = warp - 4;
    const int row = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float m = -INFINITY, l = 0.f, alpha_pending = 0.f;
    float o[kAD];
#pragma unroll
    for (int i = 0; i < kAD; ++i) o[i] = 0.f;

    auto accumulate = [&](int t, float a) {
      const int bsel = t & 1;
      mbar_wait(&bars->ot_full[bsel], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + lane_base + kColOT + bsel * 128 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * a + __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ot_free[bsel]);
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const bool diag = j == qblk;
      mbar_wait(&bars->s_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t sb = tmem + lane_base + kColS + s * 128;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(sb + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(r[i]);
          if (!diag || c * 32 + i <= row) mx = fmaxf(mx, v);
        }
      }
      const float m_new = fmaxf(m, mx * scale_log2);
      const float alpha = exp2f(m - m_new);  // first tile: exp2(-inf) = 0
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(sb + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = exp2f(__uint_as_float(r[2 * i]) * scale_log2 - m_new);
          float p1 = exp2f(__uint_as_float(r[2 * i + 1]) * scale_log2 - m_new);
          if (diag) {
            if (c * 32 + 2 * i > row) p0 = 0.f;
            if (c * 32 + 2 * i + 1 > row) p1 = 0.f;
          }
          rs += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        // P(j) overwrites columns [16c, 16c+16) of the S buffer: this thread has already read S columns [0, 32c+32) of its row
        tmem_st_32x32_x16(sb + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[s]);
      l = l * alpha + rs;
      m = m_new;
      if (j >= 1) accumulate(j - 1, alpha_pending);  // O_tile(j-1) finished while this tile's softmax ran
      alpha_pending = alpha;
    }
    accumulate(n_tiles - 1, alpha_pending);

    // ---- epilogue: normalise, bf16, swizzled smem (the Q tile is dead: every QK^T has completed), TMA store
    const float inv_l = 1.f / l;
    if (lse != nullptr) lse[((size_t)b * Hq + h) * S + qblk * kAQ + row] = (m + log2f(l)) * 0.6931471805599453f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = o[c * 64 + i] * inv_l;
      uint8_t* buf = sQ + (qd * 2 + c) * 4096;
      epi_write_row_swizzled(buf, lane, v);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_o, buf, h * kAD + c * 64, row0 + qd * 32);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward, variant 2: EIGHT softmax warps.  Warps w and w+4 share a TMEM lane quadrant and split every row's columns in halves
// (S columns / head-dim columns [64 hf, 64 hf + 64)), so the per-row exp2 / convert / accumulate work — the forward's bottleneck
// on Blackwell (16 ex2 per clock per SM against 8192 tensor FLOP per clock) — is spread over 256 threads and each thread keeps
// only 64 output accumulators.  The row maximum is agreed on through shared memory (one 64-thread named barrier per quadrant
// per tile).  P gets its own TMEM columns (no aliasing with S, which the partner thread may still be reading):
//     S0 [0,128) S1 [128,256) | P0 [256,320) P1 [320,384) | O_tile [384,512)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kA2Threads = 384;  // w0 TMA, w1 MMA, w2 TMEM alloc, w3 spare, w4-11 softmax (quadrant = w % 4, half = (w - 4) / 4)
constexpr uint32_t k2ColP = 256, k2ColOT = 384;

struct Attn2Bars {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t s_full[2], p_full[2], s_free[2], ot_full, ot_free;
  uint32_t tmem_holder;
  uint32_t pad;
  float xmax[2][2][128];  // [tile parity][half][row]
  float xsum[2][128];
};

VB_DEVICE void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <bool X64>
__global__ void __launch_bounds__(kA2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o, float* __restrict__ lse, int B, int S, int Hq, int Hkv,
                 float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kATile;
  uint8_t* sV = smem + 3 * kATile;
  Attn2Bars* bars = reinterpret_cast<Attn2Bars*>(smem + 5 * kATile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = Hq / Hkv, nq = S / kAQ;
  int idx = blockIdx.x;
  const int g = idx % G;
  idx /= G;
  const int qblk = nq - 1 - idx % nq;
  idx /= nq;
  const int kvh = idx % Hkv, b = idx / Hkv;
  const int h = kvh * G + g;
  const int n_tiles = qblk + 1;
  const int row0 = b * S + qblk * kAQ;
  const int col_q = h * kAD, col_k = (Hq + kvh) * kAD, col_v = (Hq + Hkv + kvh) * kAD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_o);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->k_full[s], 1);
      mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1);
      mbar_init(&bars->v_empty[s], 1);
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], 8);
      mbar_init(&bars->s_free[s], 1);
    }
    mbar_init(&bars->ot_full, 1);
    mbar_init(&bars->ot_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars->q_full, kATile);
      tma_load_2d(sQ, &tm_qkv, &bars->q_full, col_q, row0);
      tma_load_2d(sQ + kAHalf, &tm_qkv, &bars->q_full, col_q + 64, row0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int krow = b * S + j * kAKV;
        mbar_wait(&bars->k_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->k_full[s], kATile);
        tma_load_2d(sK + s * kATile, &tm_qkv, &bars->k_full[s], col_k, krow);
        tma_load_2d(sK + s * kATile + kAHalf, &tm_qkv, &bars->k_full[s], col_k + 64, krow);
        mbar_wait(&bars->v_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->v_full[s], kATile);
        tma_load_2d(sV + s * kATile, &tm_qkv, &bars->v_full[s], col_v, krow);
        tma_load_2d(sV + s * kATile + kAHalf, &tm_qkv, &bars->v_full[s], col_v + 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_major(kAQ, kAKV, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(kAQ, kAD, false, true);
      mbar_wait(&bars->q_full, 0);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bars->k_full[s], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&bars->s_free[s], ((j >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem + kColS + s * 128;
#pragma unroll
        for (int k = 0; k < kAD / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sQ + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          const uint64_t bb = make_sw128_desc(smem_u32(sK + s * kATile + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(d, a, bb, idesc_qk, k ? 1u : 0u);
        }
        umma_commit(&bars->k_empty[s]);
        umma_commit(&bars->s_full[s]);
      };
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&bars->p_full[s], ph);
        mbar_wait(&bars->v_full[s], ph);
        if (!X64 && j >= 1) mbar_wait(&bars->ot_free, (j - 1) & 1);  // O_tile(j-1) folded into the running output
        tc_fence_after();
        const uint32_t p = tmem + k2ColP + s * 64;
        const uint64_t vdesc = make_sw128_desc_mn_lbo(smem_u32(sV + s * kATile), kAHalf);
        // X64 (lazy rescaling): the output accumulates IN TMEM across all key tiles; the softmax warps rescale it in place only
        // when a row maximum grew by more than 2^8 (p_full(j) is signalled after any such rescale, so it is ordered before PV(j))
#pragma unroll
        for (int k = 0; k < kAKV / 16; ++k) umma_bf16_ts(tmem + k2ColOT, p + k * 8, vdesc + (uint64_t)(k * 128), idesc_pv, (X64 ? (j | k) : k) ? 1u : 0u);
        umma_commit(&bars->v_empty[s]);
        umma_commit(&bars->s_free[s]);
        umma_commit(&bars->ot_full);
      }
    }
  } else if (warp >= 4) {
    const int qd = warp & 3, hf = (warp - 4) >> 2;
    const int row = qd * 32 + lane;
    const int c0 = hf * 64;  // first S / head-dim column of this thread
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float m = -INFINITY, l = 0.f, alpha_pending = 0.f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;

    auto accumulate = [&](int t, float a) {
      mbar_wait(&bars->ot_full, t & 1);
      tc_fence_after();
      if (X64) {
        uint32_t r[64];
        tmem_ld_32x64(tmem + lane_base + k2ColOT + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = o[i] * a + __uint_as_float(r[i]);
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem + lane_base + k2ColOT + c0 + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * a + __uint_as_float(r[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ot_free);
    };

    auto tile = [&](int j, auto diag_tag) {
      constexpr bool diag = decltype(diag_tag)::value;  // compare-and-select code only in the diagonal-tile instantiation
      const int s = j & 1;
      mbar_wait(&bars->s_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t sb = tmem + lane_base + kColS + s * 128 + c0;
      float mx = -INFINITY, m_new, alpha, rs = 0.f;
      if (X64) {
        // one TMEM round trip per tile: the 64 S values of this thread stay in registers between the max and the exp2 pass
        uint32_t r[64];
        tmem_ld_32x64(sb, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (!diag || c0 + i <= row) mx = fmaxf(mx, __uint_as_float(r[i]));
        bars->xmax[j & 1][hf][row] = mx;
        named_bar_sync(1 + qd, 64);
        mx = fmaxf(mx, bars->xmax[j & 1][hf ^ 1][row]);
        // Lazy rescaling: `m` is the REFERENCE maximum of the row, not necessarily the running one.  It is raised — and the
        // output accumulator in TMEM rescaled — only when the true maximum exceeds it by more than 8 (a factor 256, far inside
        // bf16 / fp32 range for P and O); until then P = exp2(s - m) may exceed 1.  The decision is warp-uniform (TMEM loads and
        // stores are warp-collective) and identical in the partner warp, which sees the same maxima.
        m_new = fmaxf(m, mx * scale_log2);
        const bool grow = m_new > m + 8.f;
        alpha = 1.f;
        if (j == 0) {
          m = m_new;
        } else if (__any_sync(0xffffffffu, grow)) {
          if (grow) {
            alpha = exp2f(m - m_new);
            m = m_new;
          }
          mbar_wait(&bars->ot_full, (j - 1) & 1);  // PV(j-1) has landed in the accumulator
          tc_fence_after();
          uint32_t oo[64];
          tmem_ld_32x64(tmem + lane_base + k2ColOT + c0, oo);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 64; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
          tmem_st_32x64(tmem + lane_base + k2ColOT + c0, oo);
        }
        m_new = m;
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float p0 = exp2f(__uint_as_float(r[2 * i]) * scale_log2 - m_new);
          float p1 = exp2f(__uint_as_float(r[2 * i + 1]) * scale_log2 - m_new);
          if (diag) {
            if (c0 + 2 * i > row) p0 = 0.f;
            if (c0 + 2 * i + 1 > row) p1 = 0.f;
          }
          rs += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32_x32(tmem + lane_base + k2ColP + s * 64 + hf * 32, pk);
      } else {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(sb + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!diag || c0 + c * 32 + i <= row) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      // agree on the row maximum with the thread that owns the other half of this row
      bars->xmax[j & 1][hf][row] = mx;
      named_bar_sync(1 + qd, 64);
      mx = fmaxf(mx, bars->xmax[j & 1][hf ^ 1][row]);
      m_new = fmaxf(m, mx * scale_log2);
      alpha = exp2f(m - m_new);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(sb + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = exp2f(__uint_as_float(r[2 * i]) * scale_log2 - m_new);
          float p1 = exp2f(__uint_as_float(r[2 * i + 1]) * scale_log2 - m_new);
          if (diag) {
            if (c0 + c * 32 + 2 * i > row) p0 = 0.f;
            if (c0 + c * 32 + 2 * i + 1 > row) p1 = 0.f;
          }
          rs += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32_x16(tmem + lane_base + k2ColP + s * 64 + hf * 32 + c * 16, pk);
      }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[s]);
      l = l * alpha + rs;
      m = m_new;
      if (!X64) {
        if (j >= 1) accumulate(j - 1, alpha_pending);
        alpha_pending = alpha;
      }
    };
    for (int j = 0; j < n_tiles; ++j) {
      if (j == qblk) tile(j, std::true_type{});
      else tile(j, std::false_type{});
    }
    if (X64) accumulate(n_tiles - 1, 0.f);  // o = 0 * 0 + accumulator: the finished output comes out of TMEM once
    else accumulate(n_tiles - 1, alpha_pending);

    // row sum: add the partner's half
    bars->xsum[hf][row] = l;
    named_bar_sync(1 + qd, 64);
    l += bars->xsum[hf ^ 1][row];
    const float inv_l = 1.f / l;
    if (hf == 0 && lse != nullptr) lse[((size_t)b * Hq + h) * S + qblk * kAQ + row] = (m + log2f(l)) * 0.6931471805599453f;
    {
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) v[i] = o[i] * inv_l;
      uint8_t* buf = sQ + (qd * 2 + hf) * 4096;
      epi_write_row_swizzled(buf, lane, v);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_o, buf, h * kAD + c0, row0 + qd * 32);
        tma_store_commit();
        tma_store_wait<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

constexpr int kAttnFwd2Smem = 5 * kATile + (int)sizeof(Attn2Bars) + 1024;
static_assert(kAttnFwd2Smem <= 232448, "attn_fwd2_kernel: shared memory over the 227 KB CTA limit");

constexpr int kAttnFwdSmem = 5 * kATile + (int)sizeof(AttnBars) + 1024;
static_assert(kAttnFwdSmem <= 232448, "attn_fwd_kernel: shared memory over the 227 KB CTA limit");

}  // namespace

// qkv [B, S, (Hq + 2 Hkv) * 128] bf16 contiguous (RoPE already applied) -> out [B, S, Hq * 128] bf16, lse [B, Hq, S] fp32
void attn_fwd(const at::Tensor& qkv, at::Tensor out, at::Tensor lse, int64_t n_q, int64_t n_kv, double softmax_scale, int64_t variant) {
  TORCH_CHECK(qkv.is_cuda() && qkv.scalar_type() == at::kBFloat16 && qkv.dim() == 3 && qkv.is_contiguous(), "attn_fwd: qkv must be a contiguous bf16 [B, S, C] CUDA tensor");
  const int64_t B = qkv.size(0), S = qkv.size(1), C = qkv.size(2);
  TORCH_CHECK(C == (n_q + 2 * n_kv) * kAD, "attn_fwd: head dim must be 128 and C == (Hq + 2 Hkv) * 128");
  TORCH_CHECK(S % kAQ == 0 && n_q % n_kv == 0, "attn_fwd: S must be a multiple of 128 and Hq a multiple of Hkv");
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kBFloat16 && out.is_contiguous() && out.numel() == B * S * n_q * kAD);
  TORCH_CHECK(lse.is_cuda() && lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == B * n_q * S);
  c10::cuda::CUDAGuard guard(qkv.device());
  const CUtensorMap tq = make_tmap_2d(qkv.data_ptr(), B * S, C, C * 2, 128, 64, 2, true);
  const CUtensorMap to = make_tmap_2d(out.data_ptr(), B * S, n_q * kAD, n_q * kAD * 2, 32, 64, 2, true);
  static bool attr = false;
  if (!attr) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnFwdSmem));
    attr = true;
  }
  const int grid = (int)(B * n_kv * (S / kAQ) * (n_q / n_kv));
  const float scale_log2 = (float)(softmax_scale * 1.4426950408889634);
  static const int env_variant = [] {
    const char* e = getenv("VESCALE_B200_ATTN_FWD");
    return e ? atoi(e) : 3;
  }();
  if ((variant > 0 ? (int)variant : env_variant) == 1) {
    attn_fwd_kernel<<<grid, kAThreads, kAttnFwdSmem, at::cuda::getCurrentCUDAStream()>>>(tq, to, lse.data_ptr<float>(), (int)B, (int)S, (int)n_q, (int)n_kv, scale_log2);
  } else {
    static bool attr2 = false;
    if (!attr2) {
      C10_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnFwd2Smem));
      C10_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnFwd2Smem));
      attr2 = true;
    }
    if ((variant > 0 ? (int)variant : env_variant) == 2)
      attn_fwd2_kernel<false><<<grid, kA2Threads, kAttnFwd2Smem, at::cuda::getCurrentCUDAStream()>>>(tq, to, lse.data_ptr<float>(), (int)B, (int)S, (int)n_q, (int)n_kv, scale_log2);
    else  // 3: single TMEM round trip per tile (x64 loads)
      attn_fwd2_kernel<true><<<grid, kA2Threads, kAttnFwd2Smem, at::cuda::getCurrentCUDAStream()>>>(tq, to, lse.data_ptr<float>(), (int)B, (int)S, (int)n_q, (int)n_kv, scale_log2);
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// =====================================================================================================================
// Backward.  One CTA = one 128-key tile j of one KV head; it loops over the G query heads sharing that KV head and over the
// query blocks i >= j (causal), accumulating dK_j and dV_j in TMEM and adding each dQ_i contribution to an fp32 accumulator in
// global memory.  Everything is computed TRANSPOSED (keys on the TMEM lanes) so that the probabilities never leave tensor
// memory on their way into the dV / dK products:
//     S^T  = K_j Q_i^T                     (lanes = keys, columns = queries)      -> TMEM R0
//     dP^T = V_j dO_i^T                                                            -> TMEM R1
//     softmax threads (one per key row):  P^T = exp2(S^T * c - lse_i),  dS^T = P^T o (dP^T - D_i)
//         P^T  (bf16) -> TMEM R0[0,64)    : A operand (from TMEM) of  dV_j += P^T dO_i
//         dS^T (bf16) -> TMEM R1[64,128)  : A operand (from TMEM) of  dK_j += dS^T Q_i
//         dS^T (bf16) -> shared memory     : MN-major A operand of     dQ_i  = dS K_j      -> TMEM columns [64,192) (dead halves of R0/R1)
//     dQ_i tile: tcgen05.ld -> red.global.add.v4.f32 into dq_acc[B, S, Hq, 128]
// Q_i and dO_i tiles are TMA-loaded into a two-deep ring (the same bytes serve as K-major A/B of the first two products and as
// MN-major B of dV / dK), lse_i and D_i = rowsum(dO o O) ride along as 512-byte bulk copies.
// Pre-pass: attn_bwd_prep_kernel (D, zero dq_acc).  Post-pass: attn_bwd_dq_kernel (dq_acc * scale -> bf16 into the packed dqkv).
// =====================================================================================================================
namespace {

struct BwdBars {
  uint64_t kv_full;
  uint64_t q_full[2], q_empty[2];   // Q_i + lse_i + D_i (one stage) ; released when S^T and dK have consumed it
  uint64_t do_full[2], do_empty[2];
  uint64_t s_full;                  // S^T ready                             (MMA -> softmax)
  uint64_t dp_full;                 // dP^T ready                            (MMA -> softmax)
  uint64_t p_full;                  // P^T written to TMEM                   (softmax -> MMA), count 8
  uint64_t ds_full;                 // dS^T written to shared memory         (softmax -> MMA), count 8
  uint64_t dq_full;                 // dQ tile ready                         (MMA -> softmax)
  uint64_t dq_free;                 // dQ tile drained, R1 reusable          (softmax -> MMA), count 8
  uint64_t acc_full;                // dK / dV complete                     (MMA -> epilogue)
  uint64_t stage_free;              // attn_bwd2: dQ reduces have read the staging slices (drain -> softmax), count 4
  uint32_t tmem_holder;
  uint32_t pad;
};

constexpr uint32_t kR0 = 0, kR1 = 128, kRdV = 256, kRdK = 384, kRdQ = 128;  // dQ reuses R1 once dP^T has been read
constexpr int kBwdSmemTiles = 7;  // K, V, Q x2, dO x2, dS^T
constexpr int kBwdVecBytes = 2 * 2 * 512;  // (lse, D) x 2 stages x 128 floats
// no slack for manual alignment here (7 tiles + vectors + barriers = 226.1 KB): the dynamic shared memory window is declared
// __align__(1024) and the kernel traps if the base is not 1024-byte aligned (128B-swizzled TMA boxes need it)
constexpr int kAttnBwdSmem = kBwdSmemTiles * kATile + kBwdVecBytes + (int)sizeof(BwdBars);
static_assert(kAttnBwdSmem <= 232448, "attn_bwd_kernel: shared memory over the 227 KB CTA limit");

VB_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
VB_DEVICE void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// one 128-element bf16 row (row r of a [128 x 128] tile made of two [128 x 64] 128B-swizzled boxes), from 64 packed bf16x2 words
VB_DEVICE void write_row_sw128(uint8_t* tile, int r, const uint32_t* pk /*64*/) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t base = smem_u32(tile + half * kAHalf) + r * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_shared_v4(base + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4), pk[half * 32 + 4 * j], pk[half * 32 + 4 * j + 1], pk[half * 32 + 4 * j + 2], pk[half * 32 + 4 * j + 3]);
  }
}

__global__ void __launch_bounds__(kA2Threads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dqkv,
                const __grid_constant__ CUtensorMap tm_dq, const float* __restrict__ lse, const float* __restrict__ dvec, float* __restrict__ dq_acc, int B, int S, int Hq, int Hkv, float scale_log2,
                float scale) {
  extern __shared__ __align__(1024) uint8_t smem_bwd[];
  uint8_t* smem = smem_bwd;
  if (smem_u32(smem) & 1023u) {
    if (threadIdx.x == 0) printf("[vescale_b200] attn_bwd_kernel: dynamic shared memory base %u is not 1024-byte aligned\n", smem_u32(smem));
    __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = smem + kATile;
  uint8_t* sQ = smem + 2 * kATile;   // 2 stages
  uint8_t* sDO = smem + 4 * kATile;  // 2 stages
  uint8_t* sDS = smem + 6 * kATile;
  float* sVec = reinterpret_cast<float*>(smem + 7 * kATile);  // [stage][lse 128 | D 128]
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + 7 * kATile + kBwdVecBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = Hq / Hkv, nq = S / kAQ;
  int idx = blockIdx.x;
  const int j = idx % nq;  // early key tiles first: they see the most query blocks
  idx /= nq;
  const int kvh = idx % Hkv, b = idx / Hkv;
  const int n_i = nq - j;          // query blocks j .. nq-1
  const int n_iter = G * n_i;      // iteration t -> (g = t / n_i, i = j + t % n_i)
  const int krow = b * S + j * kAKV;
  const int col_k = (Hq + kvh) * kAD, col_v = (Hq + Hkv + kvh) * kAD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_dqkv);
    prefetch_tmap(&tm_dq);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->q_full[s], 1);
      mbar_init(&bars->q_empty[s], 1);
      mbar_init(&bars->do_full[s], 1);
      mbar_init(&bars->do_empty[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->dp_full, 1);
    mbar_init(&bars->p_full, 8);
    mbar_init(&bars->ds_full, 8);
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dq_free, 8);
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(&bars->kv_full, 2 * kATile);
      tma_load_2d(sK, &tm_qkv, &bars->kv_full, col_k, krow);
      tma_load_2d(sK + kAHalf, &tm_qkv, &bars->kv_full, col_k + 64, krow);
      tma_load_2d(sV, &tm_qkv, &bars->kv_full, col_v, krow);
      tma_load_2d(sV + kAHalf, &tm_qkv, &bars->kv_full, col_v + 64, krow);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int g = t / n_i, i = j + t % n_i;
        const int h = kvh * G + g;
        const int qrow = b * S + i * kAQ;
        const size_t vec_off = ((size_t)b * Hq + h) * S + (size_t)i * kAQ;
        mbar_wait(&bars->q_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->q_full[s], kATile + 1024);
        tma_load_2d(sQ + s * kATile, &tm_qkv, &bars->q_full[s], h * kAD, qrow);
        tma_load_2d(sQ + s * kATile + kAHalf, &tm_qkv, &bars->q_full[s], h * kAD + 64, qrow);
        bulk_load_1d(sVec + s * 256, lse + vec_off, 512, &bars->q_full[s]);  // `lse` here = lse * log2(e), written by the pre-pass
        bulk_load_1d(sVec + s * 256 + 128, dvec + vec_off, 512, &bars->q_full[s]);
        mbar_wait(&bars->do_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->do_full[s], kATile);
        tma_load_2d(sDO + s * kATile, &tm_do, &bars->do_full[s], h * kAD, qrow);
        tma_load_2d(sDO + s * kATile + kAHalf, &tm_do, &bars->do_full[s], h * kAD + 64, qrow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_kk = make_idesc_bf16_major(128, 128, false, false);  // both K-major
      constexpr uint32_t idesc_tb = make_idesc_bf16_major(128, 128, false, true);   // A from TMEM (K-major), B MN-major
      constexpr uint32_t idesc_mm = make_idesc_bf16_major(128, 128, true, true);    // A and B MN-major
      constexpr uint32_t idesc_kb = make_idesc_bf16_major(128, 128, false, true);   // A K-major (smem), B MN-major
      mbar_wait(&bars->kv_full, 0);
      const uint64_t k_mn = make_sw128_desc_mn_lbo(smem_u32(sK), kAHalf);
      const uint64_t ds_mn = make_sw128_desc_mn_lbo(smem_u32(sDS), kAHalf);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        uint8_t* q = sQ + s * kATile;
        uint8_t* dO = sDO + s * kATile;
        // ---- S^T = K_j Q_i^T -> R0.  tcgen05.mma executes in issue order, so this may follow dV(t-1) (the reader of P^T(t-1) in
        // R0) directly: the tensor core starts the next tile's scores while the softmax warps still drain dQ(t-1)
        mbar_wait(&bars->q_full[s], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAD / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sK + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          const uint64_t bb = make_sw128_desc(smem_u32(q + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + kR0, a, bb, idesc_kk, k ? 1u : 0u);
        }
        umma_commit(&bars->s_full);
        // ---- dP^T = V_j dO_i^T -> R1 (holds dQ(t-1) until the softmax warps have drained it)
        mbar_wait(&bars->do_full[s], ph);
        if (t >= 1) mbar_wait(&bars->dq_free, (t - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAD / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sV + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          const uint64_t bb = make_sw128_desc(smem_u32(dO + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + kR1, a, bb, idesc_kk, k ? 1u : 0u);
        }
        umma_commit(&bars->dp_full);
        // ---- dV_j += P^T dO_i : A = P^T (TMEM R0[0,64)), B = dO_i as MN-major [q rows x d]; runs while the softmax warps compute dS
        mbar_wait(&bars->p_full, t & 1);
        tc_fence_after();
        const uint64_t do_mn = make_sw128_desc_mn_lbo(smem_u32(dO), kAHalf);
        const uint64_t q_mn = make_sw128_desc_mn_lbo(smem_u32(q), kAHalf);
#pragma unroll
        for (int k = 0; k < kAQ / 16; ++k) umma_bf16_ts(tmem + kRdV, tmem + kR0 + k * 8, do_mn + (uint64_t)(k * 128), idesc_tb, (t | k) ? 1u : 0u);
        umma_commit(&bars->do_empty[s]);
        // ---- dK_j += dS^T Q_i : A = dS^T from shared memory, K-major ([key rows x query cols]: rows = M, queries = K); B = Q_i MN-major
        mbar_wait(&bars->ds_full, t & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAQ / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sDS + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + kRdK, a, q_mn + (uint64_t)(k * 128), idesc_kb, (t | k) ? 1u : 0u);
        }
        umma_commit(&bars->q_empty[s]);
        // ---- dQ_i = dS K_j -> R1 : A = the same dS^T tile read MN-major (A[q][key]), B = K_j MN-major
#pragma unroll
        for (int k = 0; k < kAKV / 16; ++k) umma_bf16(tmem + kRdQ, ds_mn + (uint64_t)(k * 128), k_mn + (uint64_t)(k * 128), idesc_mm, k ? 1u : 0u);
        umma_commit(&bars->dq_full);
      }
      umma_commit(&bars->acc_full);
    }
  } else if (warp >= 4) {
    // ===================== softmax / dS / dQ drain: EIGHT warps, two threads per key row (query / head-dim columns split in halves) ====
    // Instruction issue of these warps is what bounds the kernel (ncu: ALU pipe highest, tensor pipe < 40 %), so per element the
    // inner loops are kept to FFMA + EX2 (+ 1/2 convert) for P and FADD + FMUL (+ 1/2 convert) for dS; lse arrives pre-multiplied by
    // log2(e), the lse / D vectors are read as float4, and the causal compare-and-select code exists only in the diagonal-tile
    // instantiation.
    const int qd = warp & 3, hf = (warp - 4) >> 2;
    const int row = qd * 32 + lane;  // key index inside the tile (S^T / dP^T lanes); query index for the dQ drain
    const int c0 = hf * 64;          // first query column (S^T, dP^T) / head-dim column (dQ, dK, dV) of this thread
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    auto tile = [&](int t, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      const int s = t & 1;
      const int g = t / n_i, i = j + t % n_i;
      const int h = kvh * G + g;
      const float4* vl4 = reinterpret_cast<const float4*>(sVec + s * 256 + c0);        // lse * log2(e)
      const float4* vd4 = reinterpret_cast<const float4*>(sVec + s * 256 + 128 + c0);  // D
      mbar_wait(&bars->s_full, t & 1);
      tc_fence_after();
      float p[64];
      uint32_t pk[32];
      {
        uint32_t r[64];
        tmem_ld_32x64(tmem + lane_base + kR0 + c0, r);  // one TMEM round trip for the thread's 64 scores
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float4 l4 = vl4[w];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int q = 4 * w + e;
            float v = exp2f(__uint_as_float(r[q]) * scale_log2 - lv[e]);
            if (DIAG && row > c0 + q) v = 0.f;  // key `row` is visible to query q iff row <= q
            p[q] = v;
          }
        }
      }
#pragma unroll
      for (int w = 0; w < 32; ++w) pk[w] = pack_bf16x2(p[2 * w], p[2 * w + 1]);
      // P^T (bf16) overwrites S^T columns [0,64) of the row: the partner thread must have finished READING its half first
      named_bar_sync(1 + qd, 64);
      tmem_st_32x32_x32(tmem + lane_base + kR0 + hf * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);  // the dV product starts now, under the dS computation below
      mbar_wait(&bars->dp_full, t & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + lane_base + kR1 + c0 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float4 d4 = vd4[c * 8 + w];
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
          float ds4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) ds4[e] = p[c * 32 + 4 * w + e] * (__uint_as_float(r[4 * w + e]) - dv[e]);
          pk[c * 16 + 2 * w] = pack_bf16x2(ds4[0], ds4[1]);
          pk[c * 16 + 2 * w + 1] = pack_bf16x2(ds4[2], ds4[3]);
        }
      }
      if (lane == 0) tma_store_wait_read<0>();  // the previous tile's dQ reduce no longer reads this warp's slice of the tile
      __syncwarp();
      {  // dS^T half-row into shared memory (box `hf` of the [128 keys x 128 queries] tile, 128B swizzle): operand of dK and dQ
        const uint32_t base = smem_u32(sDS + hf * kAHalf) + row * 128;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) st_shared_v4(base + (((uint32_t)jj ^ ((uint32_t)row & 7u)) << 4), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
      }
      fence_proxy_async();
      tc_fence_before();  // the dP^T reads above are ordered before the dQ product that overwrites R1
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ds_full);
      // ---- drain dQ_i (lanes = queries of block i, columns = head dim: this thread's half) into the fp32 accumulator.
      // 16K scalar reductions per tile from the SMs saturated the L2 atomic units (the first version spent most of its time
      // here); instead each warp stages its [32 queries x 32 dims] fp32 block in shared memory — the 4 KB slice of the dS^T
      // tile it owns, dead once the dQ product has completed — and ONE bulk tensor reduce-add (TMA) adds it to dq_acc.
      mbar_wait(&bars->dq_full, t & 1);
      tc_fence_after();
      uint8_t* stage = sDS + hf * kAHalf + qd * 4096;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + lane_base + kRdQ + c0 + c * 32, r);
        tmem_ld_wait();
        if (c == 1) {  // R0 / R1 are free for the next tile's S^T / dP^T as soon as the last TMEM read has landed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->dq_free);
        }
        if (lane == 0) tma_store_wait_read<0>();  // the previous reduce has finished reading the staging block
        __syncwarp();
        const uint32_t base = smem_u32(stage) + lane * 128;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) st_shared_v4(base + (((uint32_t)jj ^ ((uint32_t)lane & 7u)) << 4), r[4 * jj], r[4 * jj + 1], r[4 * jj + 2], r[4 * jj + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tm_dq, stage, h * kAD + c0 + c * 32, b * S + i * kAQ + qd * 32);
          tma_store_commit();
        }
      }
    };
    for (int t = 0; t < n_iter; ++t) {
      if (t % n_i == 0) tile(t, std::true_type{});  // i == j: the diagonal block
      else tile(t, std::false_type{});
    }
    // ---- epilogue: dK_j * scale and dV_j -> bf16 -> packed dqkv (the Q / dO rings are dead: reuse sQ as staging)
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t reg = which == 0 ? kRdK : kRdV;
      const float mul = which == 0 ? scale : 1.f;
      const int col = which == 0 ? col_k : col_v;
      float v[64];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + lane_base + reg + c0 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) v[cc * 32 + x] = __uint_as_float(r[x]) * mul;
      }
      uint8_t* buf = sQ + ((warp - 4) * 2 + which) * 4096;
      epi_write_row_swizzled(buf, lane, v);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_dqkv, buf, col + c0, krow + qd * 32);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// attn_bwd2_kernel: the same algorithm and memory layout as attn_bwd_kernel with the critical chain of one (key tile, query
// block) pair shortened.  ncu on the first kernel: tensor pipe < 40 %, the eight softmax warps do P, dS AND the dQ drain one
// after the other and then sit idle while dK / dQ / the next S^T run.  Changes:
//   * a fourth warpgroup (w12-15) owns the dQ drain (TMEM -> shared staging -> TMA reduce-add), so the softmax warps go straight
//     from dS(t) to P(t+1);
//   * the MMA issue order is software-pipelined: S^T(t+1) is issued right after dV(t), BEFORE waiting for dS(t), so the scores of
//     the next pair are ready when the softmax warps come back;
//   * P^T is stored in the thread's OWN column range of R0 (half 0 -> cols [0,32), half 1 -> cols [64,96)): no partner thread
//     reads those columns, which removes the named barrier between the S^T read and the P^T store;
//   * 512 threads, registers re-balanced with setmaxnreg (TMA/MMA group 80, drain group 96, softmax groups 168).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBwd2Threads = 512;
template <int N>
VB_DEVICE void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
VB_DEVICE void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

__global__ void __launch_bounds__(kBwd2Threads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dqkv,
                 const __grid_constant__ CUtensorMap tm_dq, const float* __restrict__ lse, const float* __restrict__ dvec, int B, int S, int Hq, int Hkv, float scale_log2, float scale) {
  extern __shared__ __align__(1024) uint8_t smem_bwd[];
  uint8_t* smem = smem_bwd;
  if (smem_u32(smem) & 1023u) {
    if (threadIdx.x == 0) printf("[vescale_b200] attn_bwd2_kernel: dynamic shared memory base %u is not 1024-byte aligned\n", smem_u32(smem));
    __trap();
  }
  uint8_t* sK = smem;
  uint8_t* sV = smem + kATile;
  uint8_t* sQ = smem + 2 * kATile;   // 2 stages
  uint8_t* sDO = smem + 4 * kATile;  // 2 stages
  uint8_t* sDS = smem + 6 * kATile;
  float* sVec = reinterpret_cast<float*>(smem + 7 * kATile);  // [stage][lse 128 | D 128]
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + 7 * kATile + kBwdVecBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = Hq / Hkv, nq = S / kAQ;
  int idx = blockIdx.x;
  const int j = idx % nq;
  idx /= nq;
  const int kvh = idx % Hkv, b = idx / Hkv;
  const int n_i = nq - j;
  const int n_iter = G * n_i;
  const int krow = b * S + j * kAKV;
  const int col_k = (Hq + kvh) * kAD, col_v = (Hq + Hkv + kvh) * kAD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_dqkv);
    prefetch_tmap(&tm_dq);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->q_full[s], 1);
      mbar_init(&bars->q_empty[s], 1);
      mbar_init(&bars->do_full[s], 1);
      mbar_init(&bars->do_empty[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->dp_full, 1);
    mbar_init(&bars->p_full, 8);
    mbar_init(&bars->ds_full, 8);
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dq_free, 4);
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->stage_free, 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_holder;

  if (warp < 4) {
    reg_dec<80>();
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(&bars->kv_full, 2 * kATile);
      tma_load_2d(sK, &tm_qkv, &bars->kv_full, col_k, krow);
      tma_load_2d(sK + kAHalf, &tm_qkv, &bars->kv_full, col_k + 64, krow);
      tma_load_2d(sV, &tm_qkv, &bars->kv_full, col_v, krow);
      tma_load_2d(sV + kAHalf, &tm_qkv, &bars->kv_full, col_v + 64, krow);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int g = t / n_i, i = j + t % n_i;
        const int h = kvh * G + g;
        const int qrow = b * S + i * kAQ;
        const size_t vec_off = ((size_t)b * Hq + h) * S + (size_t)i * kAQ;
        mbar_wait(&bars->q_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->q_full[s], kATile + 1024);
        tma_load_2d(sQ + s * kATile, &tm_qkv, &bars->q_full[s], h * kAD, qrow);
        tma_load_2d(sQ + s * kATile + kAHalf, &tm_qkv, &bars->q_full[s], h * kAD + 64, qrow);
        bulk_load_1d(sVec + s * 256, lse + vec_off, 512, &bars->q_full[s]);
        bulk_load_1d(sVec + s * 256 + 128, dvec + vec_off, 512, &bars->q_full[s]);
        mbar_wait(&bars->do_empty[s], ph ^ 1);
        mbar_expect_tx(&bars->do_full[s], kATile);
        tma_load_2d(sDO + s * kATile, &tm_do, &bars->do_full[s], h * kAD, qrow);
        tma_load_2d(sDO + s * kATile + kAHalf, &tm_do, &bars->do_full[s], h * kAD + 64, qrow);
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer (software-pipelined: S^T(t+1) goes out before dS(t) is waited for) =====================
      constexpr uint32_t idesc_kk = make_idesc_bf16_major(128, 128, false, false);
      constexpr uint32_t idesc_tb = make_idesc_bf16_major(128, 128, false, true);
      constexpr uint32_t idesc_mm = make_idesc_bf16_major(128, 128, true, true);
      constexpr uint32_t idesc_kb = make_idesc_bf16_major(128, 128, false, true);
      mbar_wait(&bars->kv_full, 0);
      const uint64_t k_mn = make_sw128_desc_mn_lbo(smem_u32(sK), kAHalf);
      const uint64_t ds_mn = make_sw128_desc_mn_lbo(smem_u32(sDS), kAHalf);
      auto issue_s = [&](int t) {  // S^T(t) = K_j Q_i^T -> R0
        const int s = t & 1;
        uint8_t* q = sQ + s * kATile;
        mbar_wait(&bars->q_full[s], (t >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAD / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sK + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          const uint64_t bb = make_sw128_desc(smem_u32(q + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + kR0, a, bb, idesc_kk, k ? 1u : 0u);
        }
        umma_commit(&bars->s_full);
      };
      auto issue_dp = [&](int t) {  // dP^T(t) = V_j dO_i^T -> R1 (holds dQ(t-1) until the drain warps have read it)
        const int s = t & 1;
        uint8_t* dO = sDO + s * kATile;
        mbar_wait(&bars->do_full[s], (t >> 1) & 1);
        if (t >= 1) mbar_wait(&bars->dq_free, (t - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAD / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sV + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          const uint64_t bb = make_sw128_desc(smem_u32(dO + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + kR1, a, bb, idesc_kk, k ? 1u : 0u);
        }
        umma_commit(&bars->dp_full);
      };
      issue_s(0);
      issue_dp(0);
      for (int t = 0; t < n_iter; ++t) {
        const int s = t & 1;
        uint8_t* q = sQ + s * kATile;
        uint8_t* dO = sDO + s * kATile;
        const uint64_t do_mn = make_sw128_desc_mn_lbo(smem_u32(dO), kAHalf);
        const uint64_t q_mn = make_sw128_desc_mn_lbo(smem_u32(q), kAHalf);
        // ---- dV_j += P^T dO_i : A = P^T from TMEM (queries 16k..16k+15 -> 8 packed columns at (k / 4) * 64 + (k % 4) * 8)
        mbar_wait(&bars->p_full, t & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAQ / 16; ++k)
          umma_bf16_ts(tmem + kRdV, tmem + kR0 + (uint32_t)((k >> 2) * 64 + (k & 3) * 8), do_mn + (uint64_t)(k * 128), idesc_tb, (t | k) ? 1u : 0u);
        umma_commit(&bars->do_empty[s]);
        // ---- next pair's scores: R0 is free (every S^T(t) read precedes p_full; P^T(t) is consumed by the dV product above, in order)
        if (t + 1 < n_iter) issue_s(t + 1);
        // ---- dK_j += dS^T Q_i
        mbar_wait(&bars->ds_full, t & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kAQ / 16; ++k) {
          const uint64_t a = make_sw128_desc(smem_u32(sDS + (k >> 2) * kAHalf)) + (uint64_t)((k & 3) * 2);
          umma_bf16(tmem + kRdK, a, q_mn + (uint64_t)(k * 128), idesc_kb, (t | k) ? 1u : 0u);
        }
        umma_commit(&bars->q_empty[s]);
        // ---- dQ_i = dS K_j -> R1
#pragma unroll
        for (int k = 0; k < kAKV / 16; ++k) umma_bf16(tmem + kRdQ, ds_mn + (uint64_t)(k * 128), k_mn + (uint64_t)(k * 128), idesc_mm, k ? 1u : 0u);
        umma_commit(&bars->dq_full);
        if (t + 1 < n_iter) issue_dp(t + 1);
      }
      umma_commit(&bars->acc_full);
    }
  } else if (warp >= 12) {
    // ===================== dQ drain: four warps, one thread per query row, all 128 head-dim columns in four [32 x 32] fp32 blocks.
    // Each warp stages through its two 4 KB slices of the dS^T tile (dead once the dQ product has completed) and adds every block
    // to dq_acc with ONE bulk tensor reduce (TMA); stage_free tells the softmax warps when the slices may hold dS^T again.
    reg_dec<96>();
    const int qd = warp & 3;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    uint8_t* stage = sDS + qd * 8192;
    for (int t = 0; t < n_iter; ++t) {
      const int g = t / n_i, i = j + t % n_i;
      const int h = kvh * G + g;
      mbar_wait(&bars->dq_full, t & 1);
      tc_fence_after();
#pragma unroll
      for (int cp = 0; cp < 2; ++cp) {
        uint32_t r[64];
        tmem_ld_32x64(tmem + lane_base + kRdQ + cp * 64, r);
        tmem_ld_wait();
        if (cp == 1) {  // R1 is free for the next pair's dP^T as soon as the last TMEM read has landed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bars->dq_free);
            tma_store_wait_read<0>();  // the two reduces of the first half have finished reading the staging slices
          }
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t base = smem_u32(stage + c * 4096) + lane * 128;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            st_shared_v4(base + (((uint32_t)jj ^ ((uint32_t)lane & 7u)) << 4), r[c * 32 + 4 * jj], r[c * 32 + 4 * jj + 1], r[c * 32 + 4 * jj + 2], r[c * 32 + 4 * jj + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tm_dq, stage, h * kAD + cp * 64, b * S + i * kAQ + qd * 32);
          tma_reduce_add_2d(&tm_dq, stage + 4096, h * kAD + cp * 64 + 32, b * S + i * kAQ + qd * 32);
          tma_store_commit();
        }
      }
      if (lane == 0) {
        tma_store_wait_read<0>();
        mbar_arrive(&bars->stage_free);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait<0>();
  } else {
    // ===================== softmax / dS: EIGHT warps, two threads per key row (query columns split in halves) =====================
    reg_inc<168>();
    const int qd = warp & 3, hf = (warp - 4) >> 2;
    const int row = qd * 32 + lane;
    const int c0 = hf * 64;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    auto tile = [&](int t, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      const int s = t & 1;
      const float4* vl4 = reinterpret_cast<const float4*>(sVec + s * 256 + c0);        // lse * log2(e)
      const float4* vd4 = reinterpret_cast<const float4*>(sVec + s * 256 + 128 + c0);  // D
      mbar_wait(&bars->s_full, t & 1);
      tc_fence_after();
      float p[64];
      uint32_t pk[32];
      {
        uint32_t r[64];
        tmem_ld_32x64(tmem + lane_base + kR0 + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float4 l4 = vl4[w];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int q = 4 * w + e;
            float v = exp2f(__uint_as_float(r[q]) * scale_log2 - lv[e]);
            if (DIAG && row > c0 + q) v = 0.f;
            p[q] = v;
          }
        }
      }
#pragma unroll
      for (int w = 0; w < 32; ++w) pk[w] = pack_bf16x2(p[2 * w], p[2 * w + 1]);
      tmem_st_32x32_x32(tmem + lane_base + kR0 + c0, pk);  // own columns: [c0, c0 + 32)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
      mbar_wait(&bars->dp_full, t & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + lane_base + kR1 + c0 + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float4 d4 = vd4[c * 8 + w];
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
          float ds4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) ds4[e] = p[c * 32 + 4 * w + e] * (__uint_as_float(r[4 * w + e]) - dv[e]);
          pk[c * 16 + 2 * w] = pack_bf16x2(ds4[0], ds4[1]);
          pk[c * 16 + 2 * w + 1] = pack_bf16x2(ds4[2], ds4[3]);
        }
      }
      if (t >= 1) mbar_wait(&bars->stage_free, (t - 1) & 1);  // the previous pair's dQ reduces no longer read the dS^T tile
      {
        const uint32_t base = smem_u32(sDS + hf * kAHalf) + row * 128;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) st_shared_v4(base + (((uint32_t)jj ^ ((uint32_t)row & 7u)) << 4), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
      }
      fence_proxy_async();
      tc_fence_before();  // the dP^T reads above are ordered before the dQ product that overwrites R1
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ds_full);
    };
    for (int t = 0; t < n_iter; ++t) {
      if (t % n_i == 0) tile(t, std::true_type{});
      else tile(t, std::false_type{});
    }
    // ---- epilogue: dK_j * scale and dV_j -> bf16 -> packed dqkv (the Q / dO rings are dead: reuse sQ as staging)
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t reg = which == 0 ? kRdK : kRdV;
      const float mul = which == 0 ? scale : 1.f;
      const int col = which == 0 ? col_k : col_v;
      float v[64];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + lane_base + reg + c0 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) v[cc * 32 + x] = __uint_as_float(r[x]) * mul;
      }
      uint8_t* buf = sQ + ((warp - 4) * 2 + which) * 4096;
      epi_write_row_swizzled(buf, lane, v);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm_dqkv, buf, col + c0, krow + qd * 32);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// D[b, h, s] = sum_d dO[b, s, h, d] * O[b, s, h, d] (fp32) ; dq_acc zeroed.  One warp per (b, s, h) row of 128 elements.
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO, const float* __restrict__ lse,
                                                            float* __restrict__ dvec, float* __restrict__ lse2, float* __restrict__ dq_acc, int B, int S, int Hq) {
  const int64_t nrow = (int64_t)B * S * Hq;
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < nrow; r += (int64_t)gridDim.x * 8) {
    const uint2 a = *reinterpret_cast<const uint2*>(o + r * kAD + lane * 4);
    const uint2 c = *reinterpret_cast<const uint2*>(dO + r * kAD + lane * 4);
    const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&c);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float2 x = __bfloat1622float2(ah[k]), y = __bfloat1622float2(ch[k]);
      sum += x.x * y.x + x.y * y.y;
    }
    sum = warp_sum(sum);
    const int64_t bs = r / Hq;
    const int h = (int)(r - bs * Hq);
    const int64_t bb = bs / S, ss = bs - bb * S;
    if (lane == 0) {
      dvec[(bb * Hq + h) * S + ss] = sum;
      lse2[(bb * Hq + h) * S + ss] = lse[(bb * Hq + h) * S + ss] * 1.4426950408889634f;
    }
    *reinterpret_cast<float4*>(dq_acc + r * kAD + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// dq_acc [B*S, Hq*128] fp32 * scale -> bf16 into the first Hq*128 columns of dqkv [B*S, C]
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, int64_t rows, int qcols, int C, float scale) {
  const int64_t n8 = rows * (qcols / 8);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / (qcols / 8);
    const int c8 = (int)(e - r * (qcols / 8));
    const float4 a = *reinterpret_cast<const float4*>(dq_acc + r * qcols + c8 * 8);
    const float4 b = *reinterpret_cast<const float4*>(dq_acc + r * qcols + c8 * 8 + 4);
    float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
    st8(dqkv + r * C + c8 * 8, pack8(f));
  }
}

}  // namespace

// qkv, out, dout as in attn_fwd; lse [B, Hq, S]; dqkv [B, S, C] bf16 (fully written); scratch: dvec [2, B, Hq, S] fp32 (D | lse*log2e), dq_acc [B, S, Hq*128] fp32
static int g_attn_bwd_variant = [] {
  const char* e = std::getenv("VESCALE_B200_ATTN_BWD");
  return e ? std::atoi(e) : 2;
}();
void attn_set_bwd_variant(int64_t v) { g_attn_bwd_variant = (int)v; }
int64_t attn_get_bwd_variant() { return g_attn_bwd_variant; }

void attn_bwd(const at::Tensor& qkv, const at::Tensor& out, const at::Tensor& dout, const at::Tensor& lse, at::Tensor dqkv, at::Tensor dvec, at::Tensor dq_acc,
              int64_t n_q, int64_t n_kv, double softmax_scale) {
  TORCH_CHECK(qkv.is_cuda() && qkv.scalar_type() == at::kBFloat16 && qkv.dim() == 3 && qkv.is_contiguous());
  const int64_t B = qkv.size(0), S = qkv.size(1), C = qkv.size(2);
  TORCH_CHECK(C == (n_q + 2 * n_kv) * kAD && S % kAQ == 0 && n_q % n_kv == 0);
  TORCH_CHECK(out.is_contiguous() && dout.is_contiguous() && out.scalar_type() == at::kBFloat16 && dout.scalar_type() == at::kBFloat16 && out.numel() == B * S * n_q * kAD &&
              dout.numel() == out.numel());
  TORCH_CHECK(dqkv.is_contiguous() && dqkv.scalar_type() == at::kBFloat16 && dqkv.numel() == qkv.numel());
  TORCH_CHECK(lse.scalar_type() == at::kFloat && lse.is_contiguous() && lse.numel() == B * n_q * S && dvec.scalar_type() == at::kFloat && dvec.numel() == 2 * lse.numel() &&
              dq_acc.scalar_type() == at::kFloat && dq_acc.is_contiguous() && dq_acc.numel() == B * S * n_q * kAD);
  c10::cuda::CUDAGuard guard(qkv.device());
  auto stream = at::cuda::getCurrentCUDAStream();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  float* lse2 = dvec.data_ptr<float>() + lse.numel();  // second half of the scratch vector: lse * log2(e)
  attn_bwd_prep_kernel<<<sms * 8, 256, 0, stream>>>((const __nv_bfloat16*)out.data_ptr(), (const __nv_bfloat16*)dout.data_ptr(), lse.data_ptr<float>(), dvec.data_ptr<float>(), lse2,
                                                    dq_acc.data_ptr<float>(), (int)B, (int)S, (int)n_q);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  const CUtensorMap tq = make_tmap_2d(qkv.data_ptr(), B * S, C, C * 2, 128, 64, 2, true);
  const CUtensorMap tdo = make_tmap_2d(dout.data_ptr(), B * S, n_q * kAD, n_q * kAD * 2, 128, 64, 2, true);
  const CUtensorMap tdqkv = make_tmap_2d(dqkv.data_ptr(), B * S, C, C * 2, 32, 64, 2, true);
  const CUtensorMap tdq = make_tmap_2d(dq_acc.data_ptr(), B * S, n_q * kAD, n_q * kAD * 4, 32, 32, 4, true);  // fp32 reduce-add target, [32 x 32] boxes
  static bool attr = false;
  if (!attr) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnBwdSmem));
    C10_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnBwdSmem));
    attr = true;
  }
  const int grid = (int)(B * n_kv * (S / kAQ));
  const int bwd_variant = g_attn_bwd_variant;  // 2 = attn_bwd2_kernel (default), 1 = attn_bwd_kernel
  if (bwd_variant == 1) {
    attn_bwd_kernel<<<grid, kA2Threads, kAttnBwdSmem, stream>>>(tq, tdo, tdqkv, tdq, lse2, dvec.data_ptr<float>(), dq_acc.data_ptr<float>(), (int)B, (int)S,
                                                              (int)n_q, (int)n_kv, (float)(softmax_scale * 1.4426950408889634), (float)softmax_scale);
  } else {
    attn_bwd2_kernel<<<grid, kBwd2Threads, kAttnBwdSmem, stream>>>(tq, tdo, tdqkv, tdq, lse2, dvec.data_ptr<float>(), (int)B, (int)S, (int)n_q, (int)n_kv,
                                                                 (float)(softmax_scale * 1.4426950408889634), (float)softmax_scale);
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  attn_bwd_dq_kernel<<<sms * 4, 256, 0, stream>>>(dq_acc.data_ptr<float>(), (__nv_bfloat16*)dqkv.data_ptr(), B * S, (int)(n_q * kAD), (int)C, (float)softmax_scale);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
