// Tile constants and host helpers shared by the GEMM family (plain, AG-fused, RS-fused, grouped).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vb {
constexpr int kBM = 128;          // UMMA_M (cta_group::1)
constexpr int kBN = 256;          // UMMA_N
constexpr int kBK = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int kABytes = kBM * kBK * 2;
constexpr int kBBytes = kBN * kBK * 2;
constexpr int kGemmThreads = 256;  // 8 warps: TMA, MMA, TMEM-alloc, spare, 4 x epilogue
constexpr int kEpilogueThreads = 128;

CUtensorMap make_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols,
                         int elem_bytes, bool swizzle128);
const CUtensorMap& cached_tmap_bf16(const void* p, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows);
const CUtensorMap& cached_tmap_store_bf16(const void* p, int64_t rows, int64_t cols, int64_t pitch_elems);
int gemm_smem_bytes(int stages);
}  // namespace vb
