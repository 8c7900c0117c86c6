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

constexpr int kAttnFwdSmem = 5 * kATile + (int)sizeof(AttnBars) + 1024;
static_assert(kAttnFwdSmem <= 232448, "attn_fwd_kernel: shared memory over the 227 KB CTA limit");

}  // namespace

// qkv [B, S, (Hq + 2 Hkv) * 128] bf16 contiguous (RoPE already applied) -> out [B, S, Hq * 128] bf16, lse [B, Hq, S] fp32
void attn_fwd(const at::Tensor& qkv, at::Tensor out, at::Tensor lse, int64_t n_q, int64_t n_kv, double softmax_scale) {
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
  attn_fwd_kernel<<<grid, kAThreads, kAttnFwdSmem, at::cuda::getCurrentCUDAStream()>>>(tq, to, lse.data_ptr<float>(), (int)B, (int)S, (int)n_q, (int)n_kv, scale_log2);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
