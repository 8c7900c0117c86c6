// Expert-parallel token dispatch / combine over NVLink symmetric memory (sm_100a) — no host synchronisation.
//
//   moe_exchange_counts : every rank publishes its per-expert token counts to all peers (C13: the reference does
//                         an all_to_all_single of int64 splits followed by .tolist() — a host sync)
//   moe_plan            : from the [W, E] count matrix each rank derives, on the device: its receive layout
//                         (per local expert, 256-row aligned segments ordered by source rank), its send offsets into
//                         every destination, and the m-tile -> expert map that drives the grouped GEMM
//   moe_dispatch_put    : token rows (sorted by expert) are stored straight into the destination expert's segment
//                         of the peer's receive buffer (one warp per row, 16-byte vectors, 8 KB contiguous per row)
//   moe_combine_get     : each token pulls its k expert outputs back from the peers, scales them by the gate
//                         weights and accumulates in fp32 (fuses all-to-all + `*= weight` + index_add_, C15)
//
// The same two data movers implement the backward pass (combine^T = put of scaled gradients, dispatch^T = get).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "common.cuh"

using namespace vb;

namespace {

constexpr int kMaxW = 8;
struct Ptrs {
  void* p[kMaxW];
};
inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }
Ptrs to_ptrs(const std::vector<int64_t>& v) {
  Ptrs a{};
  for (size_t i = 0; i < v.size() && i < (size_t)kMaxW; ++i) a.p[i] = reinterpret_cast<void*>(v[i]);
  return a;
}

VB_DEVICE void spin_ge(const uint32_t* f, uint32_t want) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
    if (clock64() - t0 > 20000000000LL) {
      printf("[vescale_b200] moe flag timeout want %u have %u\n", want, ld_relaxed_sys(f));
      __trap();
    }
  }
}

// counts_all on every rank: [W, E] int32 ; my row is written into all peers, then a flag.
__global__ void exchange_counts_kernel(const int* __restrict__ my_counts, Ptrs counts_all, Ptrs flags, int W, int E, int rank, uint32_t epoch) {
  for (int p = 0; p < W; ++p) {
    int* dst = reinterpret_cast<int*>(counts_all.p[p]) + rank * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) dst[e] = my_counts[e];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < W) st_release_sys(reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]) + rank, epoch);
  // wait for everyone's row
  if (threadIdx.x < W) spin_ge(reinterpret_cast<uint32_t*>(flags.p[rank]) + threadIdx.x, epoch);
}

// One block.  Outputs (all int32, device):
//   recv_seg_start[El*W]  start row of segment (local expert e, source s) in MY receive buffer
//   send_off[E]           start row, inside the destination's buffer, of MY rows for global expert e
//   tile_expert[max_tiles] local expert of every 256-row tile of my receive buffer (-1 = empty)
//   expert_rows[El], total_rows[1]
__global__ void plan_kernel(const int* __restrict__ counts_all, int* recv_seg_start, int* send_off, int* tile_expert, int* expert_rows, int* total_rows,
                            int W, int E, int rank, int max_tiles, int tile_rows) {
  const int El = E / W;
  if (threadIdx.x == 0) {
    // receive layout of every destination d (needed for my send offsets); mine also fills recv_seg_start / tile map
    for (int d = 0; d < W; ++d) {
      int row = 0;
      for (int le = 0; le < El; ++le) {
        const int e = d * El + le;
        const int seg0 = row;
        for (int s = 0; s < W; ++s) {
          if (d == rank) recv_seg_start[le * W + s] = row;
          if (s == rank) send_off[e] = row;
          row += counts_all[s * E + e];
        }
        if (d == rank) {
          expert_rows[le] = row - seg0;
          const int t0 = seg0 / tile_rows, t1 = (row + tile_rows - 1) / tile_rows;
          for (int t = t0; t < t1 && t < max_tiles; ++t) tile_expert[t] = le;
        }
        row = (row + tile_rows - 1) / tile_rows * tile_rows;  // next expert starts on a tile boundary
      }
      if (d == rank) {
        total_rows[0] = row;
        for (int t = row / tile_rows; t < max_tiles; ++t) tile_expert[t] = -1;
      }
    }
  }
}

// rows: my token copies sorted by global expert.  row_expert[r] = global expert, row_pos[r] = index of r among my rows of that expert.
__global__ void __launch_bounds__(256) dispatch_put_kernel(const __nv_bfloat16* __restrict__ rows, const int* __restrict__ row_expert,
                                                           const int* __restrict__ row_pos, const int* __restrict__ send_off, Ptrs recv, Ptrs flags,
                                                           int n_rows, int H, int W, int E, int rank, uint32_t epoch, uint32_t* done_counter) {
  const int El = E / W;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int hv = H / 8;
  for (int r = w; r < n_rows; r += warps) {
    const int e = row_expert[r];
    const int d = e / El;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(recv.p[d]) + (size_t)(send_off[e] + row_pos[r]) * H;
    const __nv_bfloat16* src = rows + (size_t)r * H;
    for (int v = lane; v < hv; v += 32) st_stream(dst + v * 8, ld_stream(src + v * 8));
  }
  // completion: last CTA to finish publishes "rank's rows delivered" to every peer
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(done_counter, 1u);
    if (old + 1 == epoch * gridDim.x) {
      __threadfence_system();
      for (int p = 0; p < W; ++p) st_release_sys(reinterpret_cast<uint32_t*>(flags.p[p]) + W + rank, epoch);
    }
  }
}

__global__ void wait_flags_kernel(const uint32_t* my_flags, int W, int base, uint32_t epoch) {
  if ((int)threadIdx.x < W) spin_ge(my_flags + base + threadIdx.x, epoch);
}
__global__ void signal_flags_kernel(Ptrs flags, int W, int base, int rank, uint32_t epoch) {
  if ((int)threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]) + base + rank, epoch);
  }
}

// out[t] = sum_j gate[t, j] * peer_rows(copy (t, j))   — copy c = t*k + j lives at (dest rank, row) given by slot_rank/slot_row
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) combine_get_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ gate, const int* __restrict__ slot_rank,
                                                          const int* __restrict__ slot_row, Ptrs src, int T, int k, int H) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int hv = H / 8;
  for (int t = w; t < T; t += warps) {
    for (int v = lane; v < hv; v += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < k; ++j) {
        const int c = t * k + j;
        const __nv_bfloat16* row = reinterpret_cast<const __nv_bfloat16*>(src.p[slot_rank[c]]) + (size_t)slot_row[c] * H;
        const uint4 raw = ld_stream(row + v * 8);
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(&raw), f);
        const float g = WEIGHTED ? gate[c] : 1.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += g * f[i];
      }
      st8(out + (size_t)t * H + v * 8, pack8(acc));
    }
  }
}

}  // namespace

void moe_exchange_counts(const at::Tensor& my_counts, std::vector<int64_t> counts_all_ptrs, std::vector<int64_t> flag_ptrs, int64_t rank, int64_t epoch) {
  TORCH_CHECK(my_counts.is_cuda() && my_counts.scalar_type() == at::kInt && my_counts.is_contiguous());
  const int W = counts_all_ptrs.size(), E = my_counts.numel();
  c10::cuda::CUDAGuard guard(my_counts.device());
  exchange_counts_kernel<<<1, 256, 0, cur_stream()>>>(my_counts.data_ptr<int>(), to_ptrs(counts_all_ptrs), to_ptrs(flag_ptrs), W, E, (int)rank, (uint32_t)epoch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void moe_plan(const at::Tensor& counts_all, at::Tensor recv_seg_start, at::Tensor send_off, at::Tensor tile_expert, at::Tensor expert_rows,
              at::Tensor total_rows, int64_t rank, int64_t tile_rows) {
  const int W = counts_all.size(0), E = counts_all.size(1);
  c10::cuda::CUDAGuard guard(counts_all.device());
  plan_kernel<<<1, 32, 0, cur_stream()>>>(counts_all.data_ptr<int>(), recv_seg_start.data_ptr<int>(), send_off.data_ptr<int>(), tile_expert.data_ptr<int>(),
                                          expert_rows.data_ptr<int>(), total_rows.data_ptr<int>(), W, E, (int)rank, (int)tile_expert.numel(), (int)tile_rows);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void moe_dispatch_put(const at::Tensor& rows, const at::Tensor& row_expert, const at::Tensor& row_pos, const at::Tensor& send_off,
                      std::vector<int64_t> recv_ptrs, std::vector<int64_t> flag_ptrs, at::Tensor done_counter, int64_t E, int64_t rank, int64_t epoch) {
  TORCH_CHECK(rows.is_cuda() && rows.scalar_type() == at::kBFloat16 && rows.is_contiguous() && rows.size(1) % 8 == 0);
  const int W = recv_ptrs.size();
  c10::cuda::CUDAGuard guard(rows.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  dispatch_put_kernel<<<sms, 256, 0, cur_stream()>>>((const __nv_bfloat16*)rows.data_ptr(), row_expert.data_ptr<int>(), row_pos.data_ptr<int>(),
                                                    send_off.data_ptr<int>(), to_ptrs(recv_ptrs), to_ptrs(flag_ptrs), (int)rows.size(0), (int)rows.size(1), W,
                                                    (int)E, (int)rank, (uint32_t)epoch, reinterpret_cast<uint32_t*>(done_counter.data_ptr<int>()));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void moe_wait(int64_t my_flags, int64_t W, int64_t base, int64_t epoch) {
  wait_flags_kernel<<<1, 32, 0, cur_stream()>>>(reinterpret_cast<const uint32_t*>(my_flags), (int)W, (int)base, (uint32_t)epoch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void moe_signal(std::vector<int64_t> flag_ptrs, int64_t base, int64_t rank, int64_t epoch) {
  signal_flags_kernel<<<1, 32, 0, cur_stream()>>>(to_ptrs(flag_ptrs), (int)flag_ptrs.size(), (int)base, (int)rank, (uint32_t)epoch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void moe_combine_get(at::Tensor out, c10::optional<at::Tensor> gate, const at::Tensor& slot_rank, const at::Tensor& slot_row, std::vector<int64_t> src_ptrs,
                     int64_t k) {
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kBFloat16 && out.is_contiguous() && out.size(1) % 8 == 0);
  c10::cuda::CUDAGuard guard(out.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int T = out.size(0), H = out.size(1);
  if (T == 0) return;
  if (gate.has_value())
    combine_get_kernel<true><<<sms * 2, 256, 0, cur_stream()>>>((__nv_bfloat16*)out.data_ptr(), gate->data_ptr<float>(), slot_rank.data_ptr<int>(),
                                                               slot_row.data_ptr<int>(), to_ptrs(src_ptrs), T, (int)k, H);
  else
    combine_get_kernel<false><<<sms * 2, 256, 0, cur_stream()>>>((__nv_bfloat16*)out.data_ptr(), nullptr, slot_rank.data_ptr<int>(), slot_row.data_ptr<int>(),
                                                                to_ptrs(src_ptrs), T, (int)k, H);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
