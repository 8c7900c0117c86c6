// MXFP8 GEMM for sm_100a:  C[M,N] (bf16) = A[M,K] * B[N,K]^T with both operands in OCP microscaling format
// (e4m3 elements, one E8M0 power-of-two scale per 32 consecutive K elements of a row).
//
//   * tcgen05.mma.kind::mxf8f6f4.block_scale (UMMA 128x256x32): the tensor core applies the block scales itself, so the
//     mainloop is as dense as the bf16 one -- no per-block promotion of partial sums on the CUDA cores.
//   * operands: TMA, 128-byte swizzle, K block = 128 e4m3 = one swizzle row, 4 MMAs per stage.
//   * scale factors: stored in global memory in the tensor core's native order (512-byte atoms = 128 rows x 4 K-blocks,
//     byte (r % 32) * 16 + (r / 32) * 4 + kblk; atoms of one row block are consecutive along K), one bulk copy per atom into
//     shared memory, then `tcgen05.cp.32x128b.warpx4` into 4 TMEM columns per 128 rows.  tcgen05.cp and tcgen05.mma execute
//     in issue order, so the same TMEM columns are reused by every stage.  The scale of K-sub-block k of a stage is byte k of
//     the column, selected by the a_sf_id / b_sf_id fields of the instruction descriptor.
//   * TMEM budget: 256 accumulator columns + 4 (SFA) + 8 (SFB).  One accumulator stage; the epilogue drains it into registers
//     (packed bf16) and hands it back to the MMA warp BEFORE it stores, so the next tile's mainloop overlaps the stores.
//   * persistent grid, same warp roles and barriers as gemm_sm100.cu (w0 TMA, w1 MMA, w2 TMEM alloc, w4-7 epilogue).
//
// Numerics specification: vescale_b200/ops/fp8.py (quantize_mx / dequantize_mx / mxfp8_gemm_nt emulation).
// This is the "block-scaled fp8" path of BASELINE.json's config 5; the reference has no fp8 path at all.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "gemm_sm100.cuh"

using namespace vb;

namespace {

constexpr int kFBK = 128;                   // K elements per stage (= bytes: one 128 B swizzle row)
constexpr int kFA = kBM * kFBK;             // 16 KB
constexpr int kFB = kBN * kFBK;             // 32 KB
constexpr int kSFAtom = 512;                // 128 rows x 4 scales
constexpr int kSFStage = 3 * kSFAtom;       // SFA atom + 2 SFB atoms
constexpr int kFStages = 4;
constexpr int kFEpi = 4 * 4096;             // 4 warps x one staging buffer of 32 rows x 128 B (two would push the CTA past 227 KB)
constexpr uint32_t kSfaCol = 256, kSfbCol = 260;

VB_DEVICE void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
// shared-memory descriptor of one scale-factor atom for tcgen05.cp: no swizzle, 8-row x 16-byte core matrices 128 B apart
VB_DEVICE uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address >> 4
  d |= (uint64_t)(128 >> 4) << 32;              // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  return d;                                     // leading byte offset 0 (one core matrix along K), layout type 0 = no swizzle
}
VB_DEVICE void tmem_cp_sf(uint32_t taddr, uint64_t sdesc) { asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory"); }
// block-scaled instruction descriptor: e4m3 x e4m3, E8M0 scales, fp32 accumulate, both operands K-major
VB_DEVICE constexpr uint32_t make_idesc_mxf8(int M, int N) { return ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24); }
VB_DEVICE void umma_mxf8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
      : "memory");
}
VB_DEVICE void epi_write_row_swizzled_packed(uint8_t* buf, int lane, const uint32_t* p /*32 packed bf16x2 = 64 columns*/) {
  const uint32_t base = smem_u32(buf) + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) st_shared_v4(base + (((uint32_t)j ^ ((uint32_t)lane & 7u)) << 4), p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_c,
                  const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kFStages * kFA;
  uint8_t* smem_sf = smem + kFStages * (kFA + kFB);
  uint8_t* smem_epi = smem_sf + kFStages * kSFStage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + kFEpi);
  uint64_t* empty_bar = full_bar + kFStages;
  uint64_t* tfull_bar = empty_bar + kFStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + kBM - 1) / kBM, num_n = (N + kBN - 1) / kBN;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kFBK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    prefetch_tmap(&tma_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, kEpilogueThreads);
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
      // ===================== producer: operand tiles by TMA, scale-factor atoms by bulk copy =====================
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % num_m, n_blk = tile / num_m;
        const uint8_t* sfa_row = sfa + (size_t)m_blk * num_kb * kSFAtom;
        const uint8_t* sfb_row0 = sfb + (size_t)(2 * n_blk) * num_kb * kSFAtom;
        const uint8_t* sfb_row1 = sfb_row0 + (size_t)num_kb * kSFAtom;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], kFA + kFB + kSFStage);
          tma_load_2d(smem_a + s * kFA, &tma_a, &full_bar[s], kb * kFBK, m_blk * kBM);
          tma_load_2d(smem_b + s * kFB, &tma_b, &full_bar[s], kb * kFBK, n_blk * kBN);
          uint8_t* sf = smem_sf + s * kSFStage;
          bulk_load(sf, sfa_row + (size_t)kb * kSFAtom, kSFAtom, &full_bar[s]);
          bulk_load(sf + kSFAtom, sfb_row0 + (size_t)kb * kSFAtom, kSFAtom, &full_bar[s]);
          bulk_load(sf + 2 * kSFAtom, sfb_row1 + (size_t)kb * kSFAtom, kSFAtom, &full_bar[s]);
          if (++s == kFStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_mxf8(kBM, kBN);
      int s = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar, aph ^ 1);  // the epilogue has copied the previous tile's accumulators out of TMEM
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sf = smem_u32(smem_sf + s * kSFStage);
          tmem_cp_sf(tmem_base + kSfaCol, make_sf_desc(sf));
          tmem_cp_sf(tmem_base + kSfbCol, make_sf_desc(sf + kSFAtom));
          tmem_cp_sf(tmem_base + kSfbCol + 4, make_sf_desc(sf + 2 * kSFAtom));
          const uint64_t a_desc = make_sw128_desc(smem_u32(smem_a + s * kFA));
          const uint64_t b_desc = make_sw128_desc(smem_u32(smem_b + s * kFB));
#pragma unroll
          for (int k = 0; k < kFBK / 32; ++k) {
            // +32 bytes (= 32 e4m3) along K inside the swizzle row; scale byte k of the SF columns
            umma_mxf8(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc | ((uint32_t)k << 29) | ((uint32_t)k << 4), (kb | k) ? 1u : 0u,
                      tmem_base + kSfaCol, tmem_base + kSfbCol);
          }
          umma_commit(&empty_bar[s]);  // operand and scale-factor smem of this stage reusable once these MMAs (and copies) are done
          if (++s == kFStages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(tfull_bar);
        aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> packed bf16 registers -> (release TMEM) -> swizzled smem -> TMA store =====================
    const int ew = warp - 4;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * kBM, n0 = (tile / num_m) * kBN;
      mbar_wait(tfull_bar, aph);
      tc_fence_after();
      uint32_t packed[kBN / 2];
#pragma unroll
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[c * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);  // the MMA warp may start the next tile now
      aph ^= 1;
#pragma unroll
      for (int c = 0; c < kBN / 64; ++c) {
        uint8_t* buf = smem_epi + ew * 4096;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        epi_write_row_swizzled_packed(buf, lane, packed + c * 32);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && n0 + c * 64 < N) {
          tma_store_2d(&tma_c, buf, n0 + c * 64, m0 + ew * 32);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------- bf16 -> MXFP8 quantiser
// One thread per 1x32 block: amax -> E8M0 exponent floor(log2(amax)) - 8 (the block maximum lands in e4m3's top binade) ->
// e4m3 elements (round to nearest even, saturating) + the scale byte written straight into the GEMM's atom order.  Same
// arithmetic as ops/fp8.py::quantize_mx, bit for bit.
__global__ void __launch_bounds__(256) mx_quantize_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf, int64_t rows,
                                                          int kblocks /* K / 32 */) {
  const int64_t nblk = rows * (int64_t)kblocks;
  const int katoms = kblocks / 4;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nblk; b += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = b / kblocks;
    const int kb = (int)(b - r * kblocks);
    const uint4* src = reinterpret_cast<const uint4*>(x + b * 32);
    float v[32];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = ld_stream(src + i);
      unpack8(*reinterpret_cast<const bf16x8*>(&u), v + 8 * i);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(v[i]));
    // floor(log2(amax)) for a normal float is its exponent field; anything below 2^-119 (incl. 0 and subnormals) clamps to -127
    int e = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127 - 8;
    if (amax < 1.5046328e-36f /* 2^-119 */) e = -127;
    const float inv = __uint_as_float((uint32_t)(127 - e) << 23);  // 2^-e, exact (e in [-127, 119] -> exponent field in [8, 254])
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t packed = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // mul.rn.f32 WITHOUT .ftz: the build uses --use_fast_math (ftz), which would flush subnormal inputs (blocks below 2^-126)
        // to zero here while the specification (ops/fp8.py::quantize_mx) scales them up into range
        float y;
        asm("mul.rn.f32 %0, %1, %2;" : "=f"(y) : "f"(v[4 * i + j]), "f"(inv));
        y = fminf(fmaxf(y, -448.f), 448.f);
        packed |= (uint32_t)__nv_cvt_float_to_fp8(y, __NV_SATFINITE, __NV_E4M3) << (8 * j);
      }
      w[i] = packed;
    }
    uint4* dst = reinterpret_cast<uint4*>(q + b * 32);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    sf[((r >> 7) * katoms + (kb >> 2)) * 512 + (r & 31) * 16 + ((r & 127) >> 5) * 4 + (kb & 3)] = (uint8_t)(e + 127);
  }
}

constexpr int kMxSmemBytes = kFStages * (kFA + kFB + kSFStage) + kFEpi + (2 * kFStages + 2) * 8 + 16 + 1024;
static_assert(kMxSmemBytes <= 232448, "gemm_mxfp8_kernel: dynamic shared memory exceeds the 227 KB per-CTA limit of sm_100");
int mxfp8_smem_bytes() { return kMxSmemBytes; }

}  // namespace

// a_q [M,K], b_q [N,K]: e4m3 bytes, K-major, K % 128 == 0.  sfa / sfb: E8M0 bytes in atom order (see the header), rows padded to
// 128 (A) / 256 (B).  c [M,N] bf16 contiguous.
void mxfp8_gemm_nt(const at::Tensor& a_q, const at::Tensor& sfa, const at::Tensor& b_q, const at::Tensor& sfb, at::Tensor c) {
  TORCH_CHECK(a_q.is_cuda() && b_q.is_cuda() && c.is_cuda() && sfa.is_cuda() && sfb.is_cuda());
  TORCH_CHECK(a_q.dim() == 2 && b_q.dim() == 2 && a_q.element_size() == 1 && b_q.element_size() == 1 && a_q.is_contiguous() && b_q.is_contiguous());
  TORCH_CHECK(c.scalar_type() == at::kBFloat16 && c.is_contiguous() && sfa.element_size() == 1 && sfb.element_size() == 1 && sfa.is_contiguous() && sfb.is_contiguous());
  const int64_t M = a_q.size(0), K = a_q.size(1), N = b_q.size(0);
  TORCH_CHECK(b_q.size(1) == K && c.size(0) == M && c.size(1) == N, "mxfp8_gemm_nt: shape mismatch");
  TORCH_CHECK(K % kFBK == 0 && K > 0, "mxfp8_gemm_nt: K must be a multiple of 128");
  TORCH_CHECK(N % 8 == 0, "mxfp8_gemm_nt: N must be a multiple of 8 (16-byte rows of C)");
  const int64_t num_kb = K / kFBK, m_atoms = (M + 127) / 128, n_atoms = (N + kBN - 1) / kBN * 2;
  TORCH_CHECK(sfa.numel() == m_atoms * num_kb * kSFAtom, "mxfp8_gemm_nt: sfa must hold ", m_atoms * num_kb, " atoms of 512 bytes");
  TORCH_CHECK(sfb.numel() == n_atoms * num_kb * kSFAtom, "mxfp8_gemm_nt: sfb must hold ", n_atoms * num_kb, " atoms of 512 bytes (rows padded to 256)");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(sfa.data_ptr()) % 16 == 0 && reinterpret_cast<uintptr_t>(sfb.data_ptr()) % 16 == 0);
  if (M == 0 || N == 0) return;
  c10::cuda::CUDAGuard guard(c.device());
  const CUtensorMap ta = make_tmap_2d(a_q.data_ptr(), M, K, K, kBM, kFBK, 1, true);
  const CUtensorMap tb = make_tmap_2d(b_q.data_ptr(), N, K, K, kBN, kFBK, 1, true);
  const CUtensorMap& tc = cached_tmap_store_bf16(c.data_ptr(), M, N, N);
  static bool attr_set = false;
  const int smem = mxfp8_smem_bytes();
  if (!attr_set) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int tiles = (int)(((M + kBM - 1) / kBM) * ((N + kBN - 1) / kBN));
  gemm_mxfp8_kernel<<<std::min(tiles, sms), kGemmThreads, smem, at::cuda::getCurrentCUDAStream()>>>(
      ta, tb, tc, reinterpret_cast<const uint8_t*>(sfa.data_ptr()), reinterpret_cast<const uint8_t*>(sfb.data_ptr()), (int)M, (int)N, (int)K);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// x [R, K] bf16 (K % 128 == 0) -> q [R, K] e4m3 bytes, sf: E8M0 bytes in atom order; the caller pre-fills `sf` with 127 (the
// padding rows keep that neutral scale).
void mx_quantize(const at::Tensor& x, at::Tensor q, at::Tensor sf) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous(), "mx_quantize: x must be a contiguous 2-D bf16 CUDA tensor");
  const int64_t R = x.size(0), K = x.size(1);
  TORCH_CHECK(K % 128 == 0, "mx_quantize: K must be a multiple of 128");
  TORCH_CHECK(q.is_cuda() && q.element_size() == 1 && q.is_contiguous() && q.numel() == R * K && sf.is_cuda() && sf.element_size() == 1 && sf.is_contiguous());
  TORCH_CHECK(sf.numel() % 512 == 0 && sf.numel() >= (R + 127) / 128 * (K / 128) * 512, "mx_quantize: sf too small");
  if (R == 0) return;
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t nblk = R * (K / 32);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<int64_t>((nblk + 255) / 256, (int64_t)sms * 16);
  mx_quantize_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>(reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), reinterpret_cast<uint8_t*>(q.data_ptr()),
                                                                        reinterpret_cast<uint8_t*>(sf.data_ptr()), R, (int)(K / 32));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
