// Fused tensor-parallel GEMM ⊕ collective kernels over NVLink symmetric memory (sm_100a, CTA-pair tcgen05).
//
//   ag_gemm   :  Y[M, Nr]  = allgather_rows(X)[M, K] * Wr[Nr, K]^T           (SP -> column-parallel linear, C9)
//   gemm_rs   :  Y[M/W, N] = reduce_scatter_rows( Xr[M, Kr] * Wr[N, Kr]^T )  (row-parallel linear -> SP, C10)
//   wag_gemm  :  Y[M, N]   = X[M, K] * allgather_rows(Wshards)[N, K]^T        (FSDP unit all-gather ⊕ first GEMM, C1/C8)
//
// Both reuse the plain GEMM's mainloop (TMA -> 6-stage smem ring -> tcgen05.mma.cta_group::2 -> TMEM -> epilogue);
// the collective is done by the same kernel, tile by tile, so transfer and math overlap:
//
// ag_gemm: tiles whose A rows are local run first.  Meanwhile warp 3 of every CTA is a *TMA copy engine*: it
//   pulls 32x256-element boxes of the peers' X shards over NVLink into shared memory and stores them into the
//   local gathered buffer (a few MB in flight per GPU from one thread per CTA), then bumps a per-row-block
//   arrival counter.  The producer of a tile that needs remote rows acquires that counter before issuing its TMA
//   loads (reads then hit local HBM/L2: every remote byte crosses NVLink exactly once).
//
// wag_gemm: the B operand is an FSDP-sharded weight: rank p owns the contiguous row range [rb[p], rb[p+1]) of W inside its
//   symmetric parameter shard (RaggedShard: uneven, possibly empty).  The copy engine gathers all W row ranges (own rows
//   first) into the unit's gathered buffer, n-blocks are visited in arrival order, and the producer of a tile acquires its
//   n-block's arrival counter before loading B.  The first GEMM of the unit therefore starts after ~one n-block (2 MB) has
//   arrived instead of after the whole unit's all-gather.
//
// gemm_rs: tiles whose output rows belong to a peer run first; their epilogue stores the bf16 partial tile
//   straight into that peer's staging slot (st.global to peer memory) and the last CTA to finish a peer's rows
//   releases a system-scope flag.  Tiles of the rows this rank owns run last; their epilogue adds the peers'
//   partials (fp32 accumulate) and writes the reduced output.  No all-reduce / reduce-scatter kernel runs.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>

#include "common.cuh"
#include "gemm_sm100.cuh"

using namespace vb;

namespace {

constexpr int kMaxW = 8;
constexpr int kB2 = (kBN / 2) * kBK * 2;   // half B tile per CTA (16 KB)
constexpr int kStage = kABytes + kB2;      // 32 KB
constexpr int STAGES = 5;                  // 160 KB ring + 32 KB copy staging
constexpr int kCopyRows = 32, kCopyCols = 256;
constexpr int kCopyBytes = kCopyRows * kCopyCols * 2;  // 16 KB box
constexpr int kCopySlots = 2;
constexpr int kEpiBytes = 4 * 2 * 4096;

struct TmapArray {
  CUtensorMap m[kMaxW];
};
struct PtrArray {
  void* p[kMaxW];
};

VB_DEVICE void spin_ge_sys(const uint32_t* f, uint32_t want) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
    if (clock64() - t0 > 20000000000LL) {
      printf("[vescale_b200] fused-TP flag timeout want %u have %u\n", want, ld_relaxed_sys(f));
      __trap();
    }
  }
}
VB_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
VB_DEVICE void spin_ge_gpu(const uint32_t* f, uint32_t want) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_gpu(f) - want) < 0) {
    if (clock64() - t0 > 20000000000LL) {
      printf("[vescale_b200] ag_gemm arrival timeout want %u have %u\n", want, ld_acquire_gpu(f));
      __trap();
    }
  }
}
VB_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct FusedArgs {
  int M, N, K, ldc;
  int world, rank;
  uint32_t epoch;
  // ag_gemm
  uint32_t* arrive;      // [M/256] arrival counters (local), monotonically increasing
  int copy_units_per_blk;  // copy boxes per 256-row block = (256/32) * (K/256)
  // gemm_rs
  uint32_t* done;        // [W] finished-CTA counters (local), monotonically increasing
  PtrArray staging;      // peer p's staging buffer base: [W slots][M/W rows][N] bf16
  PtrArray flags;        // peer p's flag array: [W] uint32 "slot src complete" ; entry [W + src] = "entered"
  // wag_gemm
  int rb[kMaxW + 1];     // weight-row ownership boundaries: rank p owns rows [rb[p], rb[p+1])
  int wait_peers;        // 0: the caller knows every peer's shard is already final (a handshake happened earlier this step)
};

// MODE 1 = ag_gemm, MODE 2 = gemm_rs, MODE 3 = wag_gemm
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
fused_tp_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_a_local, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ TmapArray peer_x, const __grid_constant__ CUtensorMap tma_full_store, const __grid_constant__ CUtensorMap tma_c,
                __nv_bfloat16* __restrict__ C, const FusedArgs fa) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * kABytes;
  uint8_t* smem_copy = smem + STAGES * kStage;  // kCopySlots x 16 KB
  uint8_t* smem_epi = smem_copy + kCopySlots * kCopyBytes;  // 4 warps x 2 buffers x 4 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* copy_bar = tempty_bar + 2;  // kCopySlots
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(copy_bar + kCopySlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int M = fa.M, N = fa.N, K = fa.K, W = fa.world, rank = fa.rank;
  constexpr int BM2 = 2 * kBM;
  const int num_m = M / BM2, num_n = (N + kBN - 1) / kBN;
  const int mpo = num_m / W;  // row blocks per owner
  const int tiles_per_owner = mpo * num_n;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kBK;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  // owner order: ag_gemm consumes its own rows first; gemm_rs produces its own rows last
  const int first_nblk = MODE == 3 ? fa.rb[rank] / kBN : 0;
  auto tile_coord = [&](int t, int& m_blk, int& n_blk, int& owner) {
    if (MODE == 3) {  // n-blocks in the order their rows arrive (own rows first), all row blocks of X per n-block
      owner = rank;
      m_blk = t % num_m;
      n_blk = (t / num_m + first_nblk) % num_n;
      return;
    }
    const int oi = t / tiles_per_owner;
    owner = MODE == 1 ? (rank + oi) % W : (rank + 1 + oi) % W;
    const int within = t - oi * tiles_per_owner;
    m_blk = owner * mpo + within % mpo;
    n_blk = within / mpo;
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
      mbar_init(&tempty_bar[s], 2 * kEpilogueThreads);
    }
    for (int s = 0; s < kCopySlots; ++s) mbar_init(&copy_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_holder, 512);
    tmem_relinquish_2cta();
  }
  // entry handshake: "I have entered call `epoch`" => my inputs are final and my staging from the last call is consumed
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(fa.flags.p[rank]);
  if (blockIdx.x == 0 && warp == 3 && lane < W) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(fa.flags.p[lane]) + W + rank, fa.epoch);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk, owner;
        tile_coord(tile, m_blk, n_blk, owner);
        const CUtensorMap* amap = &tma_a;
        int m0 = m_blk * BM2 + (int)cta * kBM;
        if (MODE == 1) {
          if (owner == rank) {
            amap = &tma_a_local;
            m0 -= rank * mpo * BM2;
          } else {
            spin_ge_gpu(fa.arrive + m_blk, fa.epoch * (uint32_t)fa.copy_units_per_blk);
            fence_proxy_async_all();
          }
        }
        if (MODE == 3) {
          spin_ge_gpu(fa.arrive + n_blk, fa.epoch * (uint32_t)fa.copy_units_per_blk);
          fence_proxy_async_all();
        }
        const int n0 = n_blk * kBN + (int)cta * (kBN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t full_leader = mapa(smem_u32(&full_bar[s]), 0);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * kStage);
          tma_load_2d_2sm(smem_a + s * kABytes, amap, full_leader, kb * kBK, m0);
          tma_load_2d_2sm(smem_b + s * kB2, &tma_b, full_leader, kb * kBK, n0);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(BM2, kBN);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_u32(smem_a + s * kABytes));
          const uint64_t b_desc = make_sw128_desc(smem_u32(smem_b + s * kB2));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16_2cta(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
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
  } else if (warp == 3) {
    if (MODE == 3 && lane == 0) {
      // ===================== TMA copy engine: every rank's weight rows -> local gathered weight =====================
      if (fa.wait_peers)
        for (int p = 0; p < W; ++p) spin_ge_sys(my_flags + W + p, fa.epoch);
      const int boxes_per_row = K / kCopyCols;
      const int row_boxes = N / kCopyRows;
      const int rot = fa.rb[rank] / kCopyRows;
      const int total = row_boxes * boxes_per_row;
      int slot = 0;
      uint32_t cph = 0;
      for (int u = blockIdx.x; u < total; u += gridDim.x) {
        const int rbx = (u / boxes_per_row + rot) % row_boxes, cb = u % boxes_per_row;
        const int r = rbx * kCopyRows;
        int p = 0;
        while (p + 1 < W && r >= fa.rb[p + 1]) ++p;
        mbar_expect_tx(&copy_bar[slot], kCopyBytes);
        tma_load_2d(smem_copy + slot * kCopyBytes, &peer_x.m[p], &copy_bar[slot], cb * kCopyCols, r - fa.rb[p]);
        mbar_wait(&copy_bar[slot], cph);
        tma_store_2d(&tma_full_store, smem_copy + slot * kCopyBytes, cb * kCopyCols, r);
        tma_store_commit();
        tma_store_wait<0>();
        __threadfence();
        atomicAdd(fa.arrive + r / kBN, 1u);
        if (++slot == kCopySlots) {
          slot = 0;
          cph ^= 1;
        }
      }
    }
    if (MODE == 1 && lane == 0) {
      // ===================== TMA copy engine: peers' X shards -> local gathered buffer =====================
      // wait until every peer has entered this call (its X shard is final)
      for (int p = 0; p < W; ++p) spin_ge_sys(my_flags + W + p, fa.epoch);
      const int boxes_per_row = K / kCopyCols;                   // boxes along K
      const int rows_per_rank = M / W;
      const int boxes_per_rank = (rows_per_rank / kCopyRows) * boxes_per_row;
      int slot = 0;
      uint32_t cph = 0;
      for (int pi = 1; pi < W; ++pi) {
        const int p = (rank + pi) % W;
        for (int u = blockIdx.x; u < boxes_per_rank; u += gridDim.x) {
          const int rb = u / boxes_per_row, cb = u % boxes_per_row;
          mbar_expect_tx(&copy_bar[slot], kCopyBytes);
          tma_load_2d(smem_copy + slot * kCopyBytes, &peer_x.m[p], &copy_bar[slot], cb * kCopyCols, rb * kCopyRows);
          mbar_wait(&copy_bar[slot], cph);
          const int grow = p * rows_per_rank + rb * kCopyRows;
          tma_store_2d(&tma_full_store, smem_copy + slot * kCopyBytes, cb * kCopyCols, grow);
          tma_store_commit();
          // the box must be globally visible before its arrival is counted
          tma_store_wait<0>();
          __threadfence();
          atomicAdd(fa.arrive + grow / BM2, 1u);
          if (++slot == kCopySlots) {
            slot = 0;
            cph ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int et = threadIdx.x - 4 * 32;  // 0..127
    int as = 0;
    uint32_t aph = 0;
    bool peers_entered = false, peers_done = false;
    const int rows_per_rank = M / W;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk, owner;
      tile_coord(tile, m_blk, n_blk, owner);
      const int m0 = m_blk * BM2 + (int)cta * kBM, n0 = n_blk * kBN;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      if (MODE == 2) {
        if (owner != rank && !peers_entered) {
          if (et < W) spin_ge_sys(my_flags + W + et, fa.epoch);  // peer has consumed its staging of the previous call
          asm volatile("bar.sync 1, 128;" ::: "memory");
          peers_entered = true;
        }
        if (owner == rank && !peers_done) {
          if (et < W && et != rank) spin_ge_sys(my_flags + et, fa.epoch);  // peer `et` has delivered all partials for my rows
          asm volatile("bar.sync 1, 128;" ::: "memory");
          peers_done = true;
        }
      }
      // destination of this warp's 32 rows: my output, or the owner's staging slot [rank] (peer memory, TMA store)
      const CUtensorMap* omap = &tma_c;
      int orow = m0 + ew * 32;
      if (MODE == 2) {
        const int lrow0 = orow - owner * rows_per_rank;
        if (owner == rank) {
          orow = lrow0;
        } else {
          omap = &peer_x.m[owner];
          orow = rank * rows_per_rank + lrow0;
        }
      }
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
        if (MODE == 2 && owner == rank) {
          // add the partial sums the peers pushed into my staging slots (fp32 accumulate)
          const int lrow = row - rank * rows_per_rank;
          const __nv_bfloat16* st = reinterpret_cast<const __nv_bfloat16*>(fa.staging.p[rank]);
          for (int p = 0; p < W; ++p) {
            if (p == rank) continue;
            const __nv_bfloat16* src = st + ((size_t)p * rows_per_rank + lrow) * N + n0 + c * 64;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (n0 + c * 64 + q * 8 < N) {
                float o[8];
                unpack8(ld8(src + q * 8), o);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[q * 8 + i] += o[i];
              }
            }
          }
        }
        uint8_t* buf = smem_epi + (ew * 2 + (c & 1)) * 4096;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        epi_write_row_swizzled(buf, lane, v);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && n0 + c * 64 < N) {
          tma_store_2d(omap, buf, n0 + c * 64, orow);
          tma_store_commit();
        }
      }
      if (MODE == 2 && owner != rank) {
        // the partial rows must have landed in the peer's memory before they are counted
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa(smem_u32(&tempty_bar[as]), 0));
      if (MODE == 2 && owner != rank) {
        // publish: my partial rows of this tile are in the peer's staging slot
        __threadfence_system();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const uint32_t old = atomicAdd(fa.done + owner, 1u);
          if (old + 1 == fa.epoch * (uint32_t)(tiles_per_owner * 2)) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<uint32_t*>(fa.flags.p[owner]) + rank, fa.epoch);
          }
        }
      }
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

int fused_smem_bytes() { return STAGES * kStage + kCopySlots * kCopyBytes + kEpiBytes + (2 * STAGES + 4 + kCopySlots) * 8 + 16 + 1024; }

PtrArray to_ptr_array(const std::vector<int64_t>& v) {
  PtrArray a{};
  for (size_t i = 0; i < v.size() && i < (size_t)kMaxW; ++i) a.p[i] = reinterpret_cast<void*>(v[i]);
  return a;
}

}  // namespace

// x_local [M/W, K] (symmetric; peer addresses in x_ptrs), w [Nr, K], x_full [M, K] scratch, y [M, Nr]
void ag_gemm(const at::Tensor& x_local, std::vector<int64_t> x_ptrs, const at::Tensor& w, at::Tensor x_full, at::Tensor y, at::Tensor arrive,
             std::vector<int64_t> flag_ptrs, int64_t rank, int64_t epoch) {
  const int W = x_ptrs.size();
  TORCH_CHECK(W >= 1 && W <= kMaxW);
  const int64_t Ml = x_local.size(0), K = x_local.size(1), M = Ml * W, Nr = w.size(0);
  TORCH_CHECK(x_local.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && x_local.is_contiguous() && w.is_contiguous());
  TORCH_CHECK(x_full.is_contiguous() && x_full.size(0) == M && x_full.size(1) == K && y.size(0) == M && y.size(1) == Nr && y.stride(1) == 1);
  TORCH_CHECK(Ml % (2 * kBM) == 0 && K % kCopyCols == 0 && Nr % 8 == 0, "ag_gemm: M/W must be a multiple of 256 and K of 256");
  TORCH_CHECK(arrive.scalar_type() == at::kInt && arrive.numel() >= M / (2 * kBM));
  c10::cuda::CUDAGuard guard(x_local.device());
  const CUtensorMap ta = make_tmap_2d(x_full.data_ptr(), M, K, K * 2, kBM, kBK, 2, true);
  const CUtensorMap tal = make_tmap_2d(x_local.data_ptr(), Ml, K, K * 2, kBM, kBK, 2, true);
  const CUtensorMap tb = make_tmap_2d(w.data_ptr(), Nr, K, w.stride(0) * 2, kBN / 2, kBK, 2, true);
  TmapArray px{};
  for (int p = 0; p < W; ++p) px.m[p] = make_tmap_2d(reinterpret_cast<void*>(x_ptrs[p]), Ml, K, K * 2, kCopyRows, kCopyCols, 2, false);
  const CUtensorMap tst = make_tmap_2d(x_full.data_ptr(), M, K, K * 2, kCopyRows, kCopyCols, 2, false);
  FusedArgs fa{};
  fa.M = M, fa.N = Nr, fa.K = K, fa.ldc = y.stride(0), fa.world = W, fa.rank = rank, fa.epoch = (uint32_t)epoch;
  fa.arrive = reinterpret_cast<uint32_t*>(arrive.data_ptr<int>());
  fa.copy_units_per_blk = (2 * kBM / kCopyRows) * (K / kCopyCols);
  fa.flags = to_ptr_array(flag_ptrs);
  const int smem = fused_smem_bytes();
  static bool attr = false;
  if (!attr) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(fused_tp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int tiles = (M / (2 * kBM)) * ((Nr + kBN - 1) / kBN);
  const int pairs = std::max(1, std::min(sms / 2, tiles));
  const CUtensorMap tcy = make_tmap_2d(y.data_ptr(), M, Nr, y.stride(0) * 2, 32, 64, 2, true);
  fused_tp_kernel<1><<<pairs * 2, kGemmThreads, smem, at::cuda::getCurrentCUDAStream()>>>(ta, tal, tb, px, tst, tcy, (__nv_bfloat16*)y.data_ptr(), fa);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// x [M, Kr], w [N, Kr]  ->  y [M/W, N] = sum over ranks, rows scattered.  staging: [W, M/W, N] bf16 symmetric.
void gemm_rs(const at::Tensor& x, const at::Tensor& w, at::Tensor y, std::vector<int64_t> staging_ptrs, at::Tensor done, std::vector<int64_t> flag_ptrs,
             int64_t rank, int64_t epoch) {
  const int W = staging_ptrs.size();
  TORCH_CHECK(W >= 1 && W <= kMaxW);
  const int64_t M = x.size(0), K = x.size(1), N = w.size(0);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && x.is_contiguous() && w.is_contiguous() && w.size(1) == K);
  TORCH_CHECK(M % (2 * kBM * W) == 0 && K % kBK == 0 && N % 8 == 0, "gemm_rs: M must be a multiple of 256*W");
  TORCH_CHECK(y.size(0) == M / W && y.size(1) == N && y.stride(1) == 1 && done.scalar_type() == at::kInt && done.numel() >= W);
  c10::cuda::CUDAGuard guard(x.device());
  const CUtensorMap ta = make_tmap_2d(x.data_ptr(), M, K, x.stride(0) * 2, kBM, kBK, 2, true);
  const CUtensorMap tb = make_tmap_2d(w.data_ptr(), N, K, w.stride(0) * 2, kBN / 2, kBK, 2, true);
  TmapArray px{};  // store maps over every peer's staging buffer viewed as [W * M/W, N]
  for (int p = 0; p < W; ++p) px.m[p] = make_tmap_2d(reinterpret_cast<void*>(staging_ptrs[p]), M, N, N * 2, 32, 64, 2, true);
  FusedArgs fa{};
  fa.M = M, fa.N = N, fa.K = K, fa.ldc = y.stride(0), fa.world = W, fa.rank = rank, fa.epoch = (uint32_t)epoch;
  fa.done = reinterpret_cast<uint32_t*>(done.data_ptr<int>());
  fa.staging = to_ptr_array(staging_ptrs);
  fa.flags = to_ptr_array(flag_ptrs);
  const int smem = fused_smem_bytes();
  static bool attr = false;
  if (!attr) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(fused_tp_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int tiles = (M / (2 * kBM)) * ((N + kBN - 1) / kBN);
  const int pairs = std::max(1, std::min(sms / 2, tiles));
  const CUtensorMap tcy = make_tmap_2d(y.data_ptr(), M / W, N, y.stride(0) * 2, 32, 64, 2, true);
  fused_tp_kernel<2><<<pairs * 2, kGemmThreads, smem, at::cuda::getCurrentCUDAStream()>>>(ta, ta, tb, px, ta, tcy, (__nv_bfloat16*)y.data_ptr(), fa);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// x [M, K]; w_full [N, K] = the weight's slice of the unit's gathered buffer (filled by this kernel); shard_ptrs[p] = address
// of the first row rank p owns (inside its symmetric parameter shard); row_bounds[W+1] = ownership boundaries.
void wag_gemm(const at::Tensor& x, at::Tensor w_full, std::vector<int64_t> shard_ptrs, std::vector<int64_t> row_bounds, at::Tensor y, at::Tensor arrive,
              std::vector<int64_t> flag_ptrs, int64_t rank, int64_t epoch, bool wait_peers) {
  const int W = shard_ptrs.size();
  TORCH_CHECK(W >= 1 && W <= kMaxW && (int)row_bounds.size() == W + 1);
  const int64_t M = x.size(0), K = x.size(1), N = w_full.size(0);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && w_full.scalar_type() == at::kBFloat16 && x.is_contiguous() && w_full.is_contiguous() && w_full.size(1) == K);
  TORCH_CHECK(M % (2 * kBM) == 0 && K % kCopyCols == 0 && N % kBN == 0, "wag_gemm: M and N must be multiples of 256, K of 256");
  TORCH_CHECK(y.size(0) == M && y.size(1) == N && y.stride(1) == 1 && arrive.scalar_type() == at::kInt && arrive.numel() >= N / kBN);
  TORCH_CHECK(row_bounds[0] == 0 && row_bounds[W] == N);
  c10::cuda::CUDAGuard guard(x.device());
  FusedArgs fa{};
  for (int p = 0; p <= W; ++p) {
    TORCH_CHECK(row_bounds[p] % kCopyRows == 0 && (p == 0 || row_bounds[p] >= row_bounds[p - 1]), "wag_gemm: ownership boundaries must be multiples of 32 rows");
    fa.rb[p] = (int)row_bounds[p];
  }
  const CUtensorMap ta = make_tmap_2d(x.data_ptr(), M, K, x.stride(0) * 2, kBM, kBK, 2, true);
  const CUtensorMap tb = make_tmap_2d(w_full.data_ptr(), N, K, K * 2, kBN / 2, kBK, 2, true);
  TmapArray px{};
  for (int p = 0; p < W; ++p) {
    const int64_t rows = row_bounds[p + 1] - row_bounds[p];
    px.m[p] = rows > 0 ? make_tmap_2d(reinterpret_cast<void*>(shard_ptrs[p]), rows, K, K * 2, kCopyRows, kCopyCols, 2, false) : ta;
  }
  const CUtensorMap tst = make_tmap_2d(w_full.data_ptr(), N, K, K * 2, kCopyRows, kCopyCols, 2, false);
  fa.M = M, fa.N = N, fa.K = K, fa.ldc = y.stride(0), fa.world = W, fa.rank = rank, fa.epoch = (uint32_t)epoch;
  fa.arrive = reinterpret_cast<uint32_t*>(arrive.data_ptr<int>());
  fa.copy_units_per_blk = (kBN / kCopyRows) * (K / kCopyCols);
  fa.flags = to_ptr_array(flag_ptrs);
  fa.wait_peers = wait_peers ? 1 : 0;
  const int smem = fused_smem_bytes();
  static bool attr = false;
  if (!attr) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(fused_tp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int tiles = (M / (2 * kBM)) * (N / kBN);
  const int pairs = std::max(1, std::min(sms / 2, tiles));
  const CUtensorMap tcy = make_tmap_2d(y.data_ptr(), M, N, y.stride(0) * 2, 32, 64, 2, true);
  fused_tp_kernel<3><<<pairs * 2, kGemmThreads, smem, at::cuda::getCurrentCUDAStream()>>>(ta, ta, tb, px, tst, tcy, (__nv_bfloat16*)y.data_ptr(), fa);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
