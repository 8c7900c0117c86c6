// bf16 GEMM for sm_100a on the 5th-gen tensor cores:  C[M,N] (+)= A[M,K] * B[N,K]^T   (both K-major)
//
//   * operands staged global -> shared by TMA (cp.async.bulk.tensor, 128B swizzle), 4-stage mbarrier ring
//   * tcgen05.mma (UMMA 128x256x16, kind::f16) issued by ONE thread, fp32 accumulators in TMEM
//   * two TMEM accumulator stages (2 x 256 columns = all 512): the epilogue of tile i overlaps the MMAs of tile i+1
//   * persistent grid (one CTA per SM), M-fastest tile order so concurrently running CTAs share B (weight) tiles in L2
//   * warp roles: w0 = TMA producer, w1 = MMA issuer, w2 = TMEM allocator, w4..7 = epilogue (tcgen05.ld -> bf16 -> global)
//
// The same mainloop is reused by the fused communication kernels (ag_gemm: B tiles fetched from peer GPUs'
// symmetric memory through per-peer tensor maps; gemm_rs: epilogue reduces into the owner rank's buffer).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <string>
#include <cudaTypedefs.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_sm100.cuh"

using namespace vb;

namespace vb {

// ------------------------------------------------------------------------------------------------- tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    auto err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    TORCH_CHECK(err == cudaSuccess && qres == cudaDriverEntryPointSuccess && p != nullptr, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// 2-D row-major [rows, cols] tensor of `elem_bytes`-wide elements, box = [box_rows, box_cols], 128B swizzle.
CUtensorMap make_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols,
                         int elem_bytes, bool swizzle128) {
  CUtensorMap m;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {row_pitch_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : (elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = get_encode_fn()(&m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", (int)r, " (ptr=", ptr, " rows=", rows, " cols=", cols, ")");
  return m;
}

}  // namespace vb

namespace {

// ------------------------------------------------------------------------------------------------- the kernel
template <int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, __nv_bfloat16* __restrict__ C, int M, int N,
               int K, int ldc, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (kABytes + kBBytes));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + kBM - 1) / kBM, num_n = (N + kBN - 1) / kBN;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpilogueThreads);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % num_m) * kBM, n0 = (tile / num_m) * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], kABytes + kBBytes);
          tma_load_2d(smem_a + s * kABytes, &tma_a, &full_bar[s], kb * kBK, m0);
          tma_load_2d(smem_b + s * kBBytes, &tma_b, &full_bar[s], kb * kBK, n0);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer (single thread) =====================
      constexpr uint32_t idesc = make_idesc_bf16(kBM, kBN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aph ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_u32(smem_a + s * kABytes));
          const uint64_t b_desc = make_sw128_desc(smem_u32(smem_b + s * kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes (= 16 bf16) along K inside the 128B swizzle atom: start-address field is in 16 B units
            umma_bf16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> bf16 -> global =====================
    const int ew = warp - 4;  // TMEM lane quadrant = warp_id % 4
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * kBM, n0 = (tile / num_m) * kBN;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + as * kBN + c * 32, r);
        tmem_ld_wait();
        if (row < M) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = n0 + c * 32 + q * 8;
            if (col < N) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[q * 8 + i]);
              if (accumulate) {
                float o[8];
                unpack8(ld8(crow + c * 32 + q * 8), o);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += o[i];
              }
              st8(crow + c * 32 + q * 8, pack8(f));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------- 2-CTA kernel
// CTA pair (cluster 2x1, cta_group::2): one UMMA 256x256x16 spans both SMs.  Each CTA stages its own 128 rows
// of A and HALF of the B tile (128 of 256 rows), so per-SM shared-memory traffic (TMA writes + tensor-core
// reads) halves for B — the 1-CTA kernel is smem-bandwidth bound at ~60% tensor-pipe utilisation (ncu,
// profiles/gemm_v1_ncu.md).  The leader CTA's single MMA thread issues for the pair; completion is multicast
// to both CTAs' barriers; both CTAs' epilogues drain their own 128 TMEM lanes.
constexpr int kB2Bytes = (kBN / 2) * kBK * 2;  // half B tile per CTA
constexpr int kStage2 = kABytes + kB2Bytes;    // 32 KB
constexpr int kEpiBytes = 4 * 2 * 4096;         // epilogue staging: 4 warps x 2 buffers x (32 rows x 128 B)

// CLC = true: tile scheduling by cluster launch control.  The grid has one cluster per tile; a cluster that finishes its tile
// cancels a not-yet-launched cluster and runs that tile itself, so tiles flow to whichever SM pairs are making progress.  The
// static `tile += num_pairs` schedule makes every GEMM as slow as its slowest CTA pair, which hurts exactly when a communication
// kernel shares some of the SMs.  Roles: warp 3 of the leader CTA is the scheduler; the TMA producer, the
// MMA issuer and the four epilogue warps of both CTAs consume every response, in order, through a 2-slot ring.
template <int STAGES, bool A_MN, bool B_MN, bool CLC = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_nt_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_c,
                    __nv_bfloat16* __restrict__ C, int M, int N, int K, int ldc, int accumulate, int group_m, const int* __restrict__ tile_expert,
                    int expert_n) {
  // grouped mode (MoE): `tile_expert[m_blk]` names the expert whose weights (rows [e*expert_n, (e+1)*expert_n) of B)
  // multiply the 256-row block m_blk of A; -1 skips the block.  All roles evaluate it identically.
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytes;
  uint8_t* smem_epi = smem + STAGES * kStage2;  // 4 warps x 2 buffers x 4 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  constexpr int kClcSlots = 2;
  uint64_t* clc_full = tempty_bar + 4;   // [kClcSlots]  (tmem_holder occupies the 16 bytes in between)
  uint64_t* clc_empty = clc_full + kClcSlots;
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(clc_empty + kClcSlots);  // kClcSlots x 16 B, 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  constexpr int BM2 = 2 * kBM;  // 256 rows per pair tile
  const int num_m = (M + BM2 - 1) / BM2, num_n = (N + kBN - 1) / kBN;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kBK;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  // grouped rasterisation: `group_m` M-tiles x all N-tiles form a group, M fastest inside -> A and B slabs stay in L2
  auto tile_coord = [&](int t, int& m_blk, int& n_blk) {
    const int per_group = group_m * num_n;
    const int g = t / per_group;
    const int first_m = g * group_m;
    const int gsize = min(group_m, num_m - first_m);
    const int r = t - g * per_group;
    m_blk = first_m + r % gsize;
    n_blk = r / gsize;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpilogueThreads);  // both CTAs' epilogue threads release the pair's accumulator
    }
    if (CLC) {
      for (int s = 0; s < kClcSlots; ++s) {
        mbar_init(&clc_full[s], 1);
        // consumers of one response (arrive on the leader's barrier): leader = producer + MMA + 4 epilogue warps + scheduler,
        // peer = producer + 4 epilogue warps
        mbar_init(&clc_empty[s], 12);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_holder, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  // next tile of this cluster.  Static: stride by the number of resident pairs.  CLC: the next response of the ring.
  int sslot = 0;
  uint32_t sph = 0;
  auto next_tile = [&](int cur, bool arrive) -> int {
    if (!CLC) return cur + num_pairs;
    mbar_wait(&clc_full[sslot], sph);
    const int x = clc_read(smem_u32(clc_resp + sslot * 16));
    fence_proxy_async();  // this generic-proxy read is ordered before the async-proxy write of the next response into the slot
    if (arrive) mbar_arrive_cluster(mapa(smem_u32(&clc_empty[sslot]), 0));
    if (++sslot == kClcSlots) {
      sslot = 0;
      sph ^= 1;
    }
    return x >= 0 ? (x >> 1) : num_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs; bytes land on the leader's full barrier) =====================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile = next_tile(tile, true)) {
        int m_blk, n_blk;
        tile_coord(tile, m_blk, n_blk);
        const int ex = tile_expert ? tile_expert[m_blk] : 0;
        if (ex < 0) continue;
        const int m0 = m_blk * BM2 + (int)cta * kBM;
        const int n0 = ex * expert_n + n_blk * kBN + (int)cta * (kBN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t full_leader = mapa(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * kStage2);
          if (!A_MN) {
            tma_load_2d_2sm(smem_a + s * kABytes, &tma_a, full_leader, kb * kBK, m0);
          } else {  // A stored [K, M]: two [64 k x 64 m] boxes side by side
            tma_load_2d_2sm(smem_a + s * kABytes, &tma_a, full_leader, m0, kb * kBK);
            tma_load_2d_2sm(smem_a + s * kABytes + 8192, &tma_a, full_leader, m0 + 64, kb * kBK);
          }
          if (!B_MN) {
            tma_load_2d_2sm(smem_b + s * kB2Bytes, &tma_b, full_leader, kb * kBK, n0);
          } else {  // B stored [K, N]
            tma_load_2d_2sm(smem_b + s * kB2Bytes, &tma_b, full_leader, n0, kb * kBK);
            tma_load_2d_2sm(smem_b + s * kB2Bytes + 8192, &tma_b, full_leader, n0 + 64, kb * kBK);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer: one thread of the leader CTA drives both tensor cores =====================
      constexpr uint32_t idesc = make_idesc_bf16_major(BM2, kBN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = pair; tile < num_tiles; tile = next_tile(tile, true)) {
        if (tile_expert) {
          int m_blk, n_blk;
          tile_coord(tile, m_blk, n_blk);
          if (tile_expert[m_blk] < 0) continue;
        }
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t a_desc = A_MN ? make_sw128_desc_mn(smem_u32(smem_a + s * kABytes)) : make_sw128_desc(smem_u32(smem_a + s * kABytes));
          const uint64_t b_desc = B_MN ? make_sw128_desc_mn(smem_u32(smem_b + s * kB2Bytes)) : make_sw128_desc(smem_u32(smem_b + s * kB2Bytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // K-major: +32 B inside the swizzle atom; MN-major: 16 k-rows = two 1024-byte atoms further
            umma_bf16_2cta(d_tmem, a_desc + (uint64_t)(A_MN ? k * 128 : k * 2), b_desc + (uint64_t)(B_MN ? k * 128 : k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit_2cta_mc(&empty_bar[s], 0b11);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit_2cta_mc(&tfull_bar[as], 0b11);
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (CLC && warp == 3) {
    if (lane == 0 && leader) {
      // ===================== tile scheduler (cluster launch control) =====================
      int slot = 0;
      uint32_t ph = 0;
      while (true) {
        mbar_wait(&clc_empty[slot], ph ^ 1);  // every consumer of both CTAs has read the previous response in this slot
        mbar_expect_tx(&clc_full[slot], 16);
        mbar_expect_tx_cluster(mapa(smem_u32(&clc_full[slot]), 1), 16);
        clc_try_cancel_mc(smem_u32(clc_resp + slot * 16), smem_u32(&clc_full[slot]));
        mbar_wait(&clc_full[slot], ph);
        const int x = clc_read(smem_u32(clc_resp + slot * 16));
        fence_proxy_async();
        mbar_arrive(&clc_empty[slot]);
        if (x < 0) break;  // nothing left to cancel: the consumers see the same response and stop too
        if (++slot == kClcSlots) {
          slot = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (each CTA drains its own 128 accumulator rows) =====================
    const int ew = warp - 4;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = pair; tile < num_tiles; tile = (__syncwarp(), next_tile(tile, lane == 0))) {
      int m_blk, n_blk;
      tile_coord(tile, m_blk, n_blk);
      if (tile_expert && tile_expert[m_blk] < 0) continue;
      const int m0 = m_blk * BM2 + (int)cta * kBM, n0 = n_blk * kBN;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      if (accumulate) {
        // C += A B^T : read-modify-write per lane (rare path)
        __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + as * kBN + c * 32, r);
          tmem_ld_wait();
          if (row < M) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = n0 + c * 32 + q * 8;
              if (col < N) {
                float f[8], o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[q * 8 + i]);
                unpack8(ld8(crow + c * 32 + q * 8), o);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += o[i];
                st8(crow + c * 32 + q * 8, pack8(f));
              }
            }
          }
        }
      } else {
        // coalesced path: TMEM -> registers -> swizzled smem -> one TMA store per 32x64 block (double buffered)
#pragma unroll 1
        for (int c = 0; c < kBN / 64; ++c) {
          float v[64];
          {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + as * kBN + c * 64, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + as * kBN + c * 64 + 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r[i]);
          }
          uint8_t* buf = smem_epi + (ew * 2 + (c & 1)) * 4096;
          if (lane == 0) tma_store_wait_read<1>();  // the store issued two chunks ago has finished reading this buffer
          __syncwarp();
          epi_write_row_swizzled(buf, lane, v);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && n0 + c * 64 < N) {
            tma_store_2d(&tma_c, buf, n0 + c * 64, m0 + ew * 32);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa(smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
  }
  if (warp >= 4 && lane == 0) tma_store_wait<0>();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// =================================================================================================
// Variant 3 (experimental, opt-in: VESCALE_B200_GEMM_VARIANT=3 or variant=3): 2x2 cluster.  Two CTA pairs share one
// cluster and compute vertically adjacent 256x256 tiles of the same n-block, so the B tile is identical for both pairs:
// each of the 4 CTAs loads one quarter of it (64 rows x 64 k) and TMA-multicasts it to the CTA of the same parity in the
// other pair.  L2->SM traffic per 512x256x64 step: 64 KB of A + 32 KB of B instead of 128 KB (-25 %), which is what limits
// *sustained* throughput under the power cap (profiles/gemm_2cta_ncu_r1.md).  Not yet validated on hardware: the default
// path and the autotune never select it.
//   ranks: pair p = rank >> 1 (m-block 2*super_m + p), c = rank & 1 (upper / lower 128 rows of the pair tile, B half c)
//   slot reuse: a CTA's B slot is written by both pairs' producers, so `empty` needs the MMA commits of *both* pairs (count 2,
//   commit multicast to all four CTAs)
// =================================================================================================
template <int STAGES>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_nt_4cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_bq, const __grid_constant__ CUtensorMap tma_c,
                    int M, int N, int K, int group_m) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytes;
  uint8_t* smem_epi = smem + STAGES * kStage2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pr = rank >> 1, c = rank & 1;  // pair in cluster, CTA in pair
  const uint32_t lead = rank & ~1u;             // cluster rank of my pair's leader
  const bool leader = c == 0;
  constexpr int BM2 = 2 * kBM;
  const int num_m = (M + BM2 - 1) / BM2, num_n = (N + kBN - 1) / kBN;
  const int num_sm = (num_m + 1) / 2;  // super tiles along M (512 rows)
  const int num_tiles = num_sm * num_n;
  const int num_kb = K / kBK;
  const int cl = blockIdx.x >> 2, num_cl = gridDim.x >> 2;

  auto tile_coord = [&](int t, int& sm_blk, int& n_blk) {
    const int per_group = group_m * num_n;
    const int g = t / per_group;
    const int first = g * group_m;
    const int gsize = min(group_m, num_sm - first);
    const int r = t - g * per_group;
    sm_blk = first + r % gsize;
    n_blk = r / gsize;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_bq);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);  // both pairs must have consumed the slot before anyone overwrites it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpilogueThreads);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_holder, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint16_t bmask = (uint16_t)((1u << c) | (1u << (c + 2)));  // same-parity CTA of both pairs
      for (int tile = cl; tile < num_tiles; tile += num_cl) {
        int sm_blk, n_blk;
        tile_coord(tile, sm_blk, n_blk);
        const int m0 = (sm_blk * 2 + (int)pr) * BM2 + (int)c * kBM;           // rows past M are zero-filled by TMA
        const int nq = n_blk * kBN + (int)c * (kBN / 2) + (int)pr * (kBN / 4);  // my quarter of the shared B tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t full_lead = mapa(smem_u32(&full_bar[s]), lead);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * kStage2);
          tma_load_2d_2sm(smem_a + s * kABytes, &tma_a, full_lead, kb * kBK, m0);
          tma_load_2d_2sm_mc(smem_b + s * kB2Bytes + pr * (kB2Bytes / 2), &tma_bq, full_lead, kb * kBK, nq, bmask);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16_major(BM2, kBN, false, false);
      const uint16_t pair_mask = (uint16_t)(0b11u << lead);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = cl; tile < num_tiles; tile += num_cl) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_u32(smem_a + s * kABytes));
          const uint64_t b_desc = make_sw128_desc(smem_u32(smem_b + s * kB2Bytes));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16_2cta(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          umma_commit_2cta_mc(&empty_bar[s], 0b1111);  // every CTA of the cluster learns that this pair is done with slot s
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit_2cta_mc(&tfull_bar[as], pair_mask);
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = cl; tile < num_tiles; tile += num_cl) {
      int sm_blk, n_blk;
      tile_coord(tile, sm_blk, n_blk);
      const int m0 = (sm_blk * 2 + (int)pr) * BM2 + (int)c * kBM, n0 = n_blk * kBN;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < kBN / 64; ++cc) {
        float v[64];
        {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + as * kBN + cc * 64, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + as * kBN + cc * 64 + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r[i]);
        }
        uint8_t* buf = smem_epi + (ew * 2 + (cc & 1)) * 4096;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        epi_write_row_swizzled(buf, lane, v);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && n0 + cc * 64 < N && m0 + ew * 32 < M) {
          tma_store_2d(&tma_c, buf, n0 + cc * 64, m0 + ew * 32);  // rows / columns past the edge are clipped by the tensor map
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa(smem_u32(&tempty_bar[as]), lead));
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
  }
  if (warp >= 4 && lane == 0) tma_store_wait<0>();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

struct MapKey {
  const void* p;
  int64_t r, c, pitch;
  int br;
  bool operator==(const MapKey& o) const { return p == o.p && r == o.r && c == o.c && pitch == o.pitch && br == o.br; }
};
struct MapHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.p) ^ (std::hash<int64_t>()(k.r) * 31) ^ (std::hash<int64_t>()(k.c) * 131) ^ (std::hash<int64_t>()(k.pitch) * 7) ^ k.br;
  }
};

}  // namespace

namespace vb {

const CUtensorMap& cached_tmap_bf16(const void* p, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows) {
  static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  MapKey k{p, rows, cols, pitch_elems, box_rows};
  auto it = cache.find(k);
  if (it == cache.end()) {
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(k, make_tmap_2d(p, rows, cols, pitch_elems * 2, box_rows, kBK, 2, true)).first;
  }
  return it->second;
}

// store map: box = 32 rows x 64 columns (128 B inner, 128B swizzle) — one epilogue warp's chunk
const CUtensorMap& cached_tmap_store_bf16(const void* p, int64_t rows, int64_t cols, int64_t pitch_elems) {
  static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  MapKey k{p, rows, cols, pitch_elems, 32};
  auto it = cache.find(k);
  if (it == cache.end()) {
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(k, make_tmap_2d(p, rows, cols, pitch_elems * 2, 32, 64, 2, true)).first;
  }
  return it->second;
}

// tile scheduling of the 2-CTA kernels: 0 = static persistent grid, 1 = cluster launch control (one cluster per tile, running
// clusters pull the remaining tiles).  VESCALE_B200_GEMM_SCHED=clc|static, or gemm_set_sched() at run time.
int g_gemm_sched = [] {
  const char* e = getenv("VESCALE_B200_GEMM_SCHED");
  return (e && std::string(e) == "clc") ? 1 : 0;
}();

int gemm_smem_bytes(int stages) { return stages * (kABytes + kBBytes) + (2 * stages + 4) * 8 + 16 + 1024; }

}  // namespace vb

void gemm_set_sched(int64_t mode) { vb::g_gemm_sched = (int)mode; }
int64_t gemm_get_sched() { return vb::g_gemm_sched; }

void gemm_nt(const at::Tensor& a, const at::Tensor& b, at::Tensor c, bool accumulate, int64_t variant_arg) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && c.is_cuda(), "gemm_nt: CUDA tensors required");
  TORCH_CHECK(a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && c.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1 && c.stride(1) == 1);
  const int64_t M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(b.size(1) == K && c.size(0) == M && c.size(1) == N, "gemm_nt: shape mismatch");
  TORCH_CHECK(K % kBK == 0 && N % 8 == 0, "gemm_nt: K must be a multiple of 64 and N of 8");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) % 16 == 0) && (reinterpret_cast<uintptr_t>(b.data_ptr()) % 16 == 0) &&
              (reinterpret_cast<uintptr_t>(c.data_ptr()) % 16 == 0) && a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0 && c.stride(0) % 8 == 0);
  if (M == 0 || N == 0) return;
  c10::cuda::CUDAGuard guard(a.device());
  static const int env_variant = [] {
    const char* e = getenv("VESCALE_B200_GEMM_VARIANT");
    return e ? atoi(e) : 2;
  }();
  const int variant = variant_arg > 0 ? (int)variant_arg : env_variant;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  if (variant == 3 && !accumulate) {
    constexpr int STAGES3 = 6;
    const CUtensorMap& ta3 = cached_tmap_bf16(a.data_ptr(), M, K, a.stride(0), kBM);
    const CUtensorMap& tb3 = cached_tmap_bf16(b.data_ptr(), N, K, b.stride(0), kBN / 4);  // quarter-tile boxes (64 rows)
    const CUtensorMap& tc3 = cached_tmap_store_bf16(c.data_ptr(), M, N, c.stride(0));
    const int smem3 = STAGES3 * kStage2 + kEpiBytes + (2 * STAGES3 + 4) * 8 + 16 + 1024;
    static int max_clusters = 0;
    if (!max_clusters) {
      C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_4cta_kernel<STAGES3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(sms / 4 * 4);
      cfg.blockDim = dim3(kGemmThreads);
      cfg.dynamicSmemBytes = smem3;
      cudaLaunchAttribute at_[1];
      at_[0].id = cudaLaunchAttributeClusterDimension;
      at_[0].val.clusterDim.x = 4;
      at_[0].val.clusterDim.y = 1;
      at_[0].val.clusterDim.z = 1;
      cfg.attrs = at_;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_nt_4cta_kernel<STAGES3>, &cfg) != cudaSuccess || n <= 0) n = sms / 4;
      max_clusters = n;
    }
    const int tiles3 = (((M + 2 * kBM - 1) / (2 * kBM) + 1) / 2) * ((N + kBN - 1) / kBN);
    const int clusters = std::max(1, std::min(max_clusters, tiles3));
    static const int group_m3 = [] {
      const char* e = getenv("VESCALE_B200_GEMM_GROUP_M");
      return e ? std::max(1, atoi(e) / 2) : 4;
    }();
    gemm_nt_4cta_kernel<STAGES3><<<clusters * 4, kGemmThreads, smem3, at::cuda::getCurrentCUDAStream()>>>(ta3, tb3, tc3, (int)M, (int)N, (int)K, group_m3);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return;
  }
  if (variant == 2 || variant == 3) {
    constexpr int STAGES2 = 6;
    const CUtensorMap& ta2 = cached_tmap_bf16(a.data_ptr(), M, K, a.stride(0), kBM);
    const CUtensorMap& tb2 = cached_tmap_bf16(b.data_ptr(), N, K, b.stride(0), kBN / 2);
    const int smem2 = STAGES2 * kStage2 + kEpiBytes + (2 * STAGES2 + 4) * 8 + 16 + 96 + 1024;
    const CUtensorMap& tc2 = cached_tmap_store_bf16(c.data_ptr(), M, N, c.stride(0));
    static bool attr2 = false;
    if (!attr2) {
      C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_2cta_kernel<STAGES2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr2 = true;
    }
    const int tiles2 = ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + kBN - 1) / kBN);
    const int pairs = std::max(1, std::min(sms / 2, tiles2));
    static const int group_m = [] {
      const char* e = getenv("VESCALE_B200_GEMM_GROUP_M");
      return e ? atoi(e) : 8;
    }();
    if (vb::g_gemm_sched == 1) {
      static bool attr2c = false;
      if (!attr2c) {
        C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_2cta_kernel<STAGES2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        attr2c = true;
      }
      gemm_nt_2cta_kernel<STAGES2, false, false, true><<<tiles2 * 2, kGemmThreads, smem2, at::cuda::getCurrentCUDAStream()>>>(
          ta2, tb2, tc2, (__nv_bfloat16*)c.data_ptr(), (int)M, (int)N, (int)K, (int)c.stride(0), accumulate ? 1 : 0, group_m, nullptr, 0);
    } else {
      gemm_nt_2cta_kernel<STAGES2, false, false><<<pairs * 2, kGemmThreads, smem2, at::cuda::getCurrentCUDAStream()>>>(
          ta2, tb2, tc2, (__nv_bfloat16*)c.data_ptr(), (int)M, (int)N, (int)K, (int)c.stride(0), accumulate ? 1 : 0, group_m, nullptr, 0);
    }
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return;
  }
  constexpr int STAGES = 4;
  const CUtensorMap& ta = cached_tmap_bf16(a.data_ptr(), M, K, a.stride(0), kBM);
  const CUtensorMap& tb = cached_tmap_bf16(b.data_ptr(), N, K, b.stride(0), kBN);
  const int smem = gemm_smem_bytes(STAGES);
  static bool attr_set = false;
  if (!attr_set) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int tiles = ((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN);
  const int grid = std::min(sms, tiles);
  gemm_nt_kernel<STAGES><<<grid, kGemmThreads, smem, at::cuda::getCurrentCUDAStream()>>>(ta, tb, (__nv_bfloat16*)c.data_ptr(), (int)M, (int)N, (int)K,
                                                                                         (int)c.stride(0), accumulate ? 1 : 0);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}


// ------------------------------------------------------------------------------------------------- transposed-operand variants
namespace {
template <bool A_MN, bool B_MN>
void launch_2cta_major(const void* a, const void* b, at::Tensor& c, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, bool accumulate) {
  constexpr int STAGES2 = 6;
  // K-major operand: map [rows=MN, cols=K], box [128, 64]; MN-major operand: map [rows=K, cols=MN], box [64, 64]
  const CUtensorMap ta = A_MN ? make_tmap_2d(a, K, M, lda * 2, 64, 64, 2, true) : make_tmap_2d(a, M, K, lda * 2, kBM, kBK, 2, true);
  const CUtensorMap tb = B_MN ? make_tmap_2d(b, K, N, ldb * 2, 64, 64, 2, true) : make_tmap_2d(b, N, K, ldb * 2, kBN / 2, kBK, 2, true);
  const CUtensorMap& tc = cached_tmap_store_bf16(c.data_ptr(), M, N, c.stride(0));
  const int smem2 = STAGES2 * kStage2 + kEpiBytes + (2 * STAGES2 + 4) * 8 + 16 + 96 + 1024;
  static bool attr = false;
  if (!attr) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_2cta_kernel<STAGES2, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    attr = true;
  }
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int tiles2 = ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + kBN - 1) / kBN);
  const int pairs = std::max(1, std::min(sms / 2, tiles2));
  if (vb::g_gemm_sched == 1) {
    static bool attrc = false;
    if (!attrc) {
      C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_2cta_kernel<STAGES2, A_MN, B_MN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attrc = true;
    }
    gemm_nt_2cta_kernel<STAGES2, A_MN, B_MN, true><<<tiles2 * 2, kGemmThreads, smem2, at::cuda::getCurrentCUDAStream()>>>(
        ta, tb, tc, (__nv_bfloat16*)c.data_ptr(), (int)M, (int)N, (int)K, (int)c.stride(0), accumulate ? 1 : 0, 8, nullptr, 0);
  } else {
    gemm_nt_2cta_kernel<STAGES2, A_MN, B_MN><<<pairs * 2, kGemmThreads, smem2, at::cuda::getCurrentCUDAStream()>>>(
        ta, tb, tc, (__nv_bfloat16*)c.data_ptr(), (int)M, (int)N, (int)K, (int)c.stride(0), accumulate ? 1 : 0, 8, nullptr, 0);
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
}  // namespace

// c[M,N] = a[M,K] @ b[K,N]      (dgrad shape: B is MN-major)
void gemm_nn(const at::Tensor& a, const at::Tensor& b, at::Tensor c) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1 && c.stride(1) == 1);
  const int64_t M = a.size(0), K = a.size(1), N = b.size(1);
  TORCH_CHECK(b.size(0) == K && c.size(0) == M && c.size(1) == N && K % kBK == 0 && N % 8 == 0 && a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0);
  if (M == 0 || N == 0) return;
  c10::cuda::CUDAGuard guard(a.device());
  launch_2cta_major<false, true>(a.data_ptr(), b.data_ptr(), c, M, N, K, a.stride(0), b.stride(0), false);
}

// c[M,N] (+)= a[K,M]^T @ b[K,N]  (wgrad shape: both operands MN-major)
void gemm_tn(const at::Tensor& a, const at::Tensor& b, at::Tensor c, bool accumulate) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1 && c.stride(1) == 1);
  const int64_t K = a.size(0), M = a.size(1), N = b.size(1);
  TORCH_CHECK(b.size(0) == K && c.size(0) == M && c.size(1) == N && K % kBK == 0 && M % 8 == 0 && N % 8 == 0 && a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0);
  if (M == 0 || N == 0) return;
  c10::cuda::CUDAGuard guard(a.device());
  launch_2cta_major<true, true>(a.data_ptr(), b.data_ptr(), c, M, N, K, a.stride(0), b.stride(0), accumulate);
}


// Grouped GEMM for MoE experts: c[r, :] = a[r, :] @ b[e(r)]^T where e(r) is constant per 256-row block
// (`tile_expert`, device-resident, -1 = skip): no host knowledge of the token counts is needed.
// a [C, K], b [E_local * N, K] (stacked expert weights), c [C, N].
void grouped_gemm_nt(const at::Tensor& a, const at::Tensor& b, at::Tensor c, const at::Tensor& tile_expert, int64_t expert_n) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16 && c.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.stride(1) == 1 && b.stride(1) == 1 && c.stride(1) == 1 && tile_expert.scalar_type() == at::kInt);
  const int64_t M = a.size(0), K = a.size(1), N = expert_n;
  TORCH_CHECK(M % (2 * kBM) == 0 && tile_expert.numel() >= M / (2 * kBM) && b.size(0) % N == 0 && c.size(0) == M && c.size(1) == N && K % kBK == 0 && N % 8 == 0);
  if (M == 0) return;
  c10::cuda::CUDAGuard guard(a.device());
  constexpr int STAGES2 = 6;
  const CUtensorMap ta = make_tmap_2d(a.data_ptr(), M, K, a.stride(0) * 2, kBM, kBK, 2, true);
  const CUtensorMap tb = make_tmap_2d(b.data_ptr(), b.size(0), K, b.stride(0) * 2, kBN / 2, kBK, 2, true);
  const CUtensorMap tcm = make_tmap_2d(c.data_ptr(), M, N, c.stride(0) * 2, 32, 64, 2, true);
  const int smem2 = STAGES2 * kStage2 + kEpiBytes + (2 * STAGES2 + 4) * 8 + 16 + 96 + 1024;
  C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_2cta_kernel<STAGES2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int tiles2 = (M / (2 * kBM)) * ((N + kBN - 1) / kBN);
  const int pairs = std::max(1, std::min(sms / 2, tiles2));
  gemm_nt_2cta_kernel<STAGES2, false, false><<<pairs * 2, kGemmThreads, smem2, at::cuda::getCurrentCUDAStream()>>>(
      ta, tb, tcm, (__nv_bfloat16*)c.data_ptr(), (int)M, (int)N, (int)K, (int)c.stride(0), 0, 8, tile_expert.data_ptr<int>(), (int)N);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
