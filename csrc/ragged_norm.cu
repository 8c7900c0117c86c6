// Segmented p-norm partials of a RaggedShard's local rows (sm_100a), without the zeros(global) + copy + vector_norm sequence of the
// reference (vescale/dtensor/_dispatch.py:146-151, a torch.compile'd helper that allocates a global-shape temporary) and without
// the fp32 upcast / power temporaries of an eager formulation: one pass over the shard in its storage dtype.
//
//   x: the local shard viewed as [rows, C] (rows = whole rows of the flattened leading dims owned by this rank, C = trailing size)
//   mode 0 ("row"): out[r] = sum_c |x[r, c]|^p          (the trailing dims are the reduced ones)   -- one warp per row
//   mode 1 ("col"): out[c] (+)= sum_r |x[r, c]|^p       (the leading dims are the reduced ones)    -- one thread per column, rows split
//   p in {1, 2, inf}; the caller takes the root / max across ranks (Partial("norm p")).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "common.cuh"

using namespace vb;

namespace {

template <int P>
VB_DEVICE float powp(float v) {
  const float a = fabsf(v);
  return P == 2 ? a * a : a;
}
template <int P>
VB_DEVICE float comb(float a, float b) { return P == 0 ? fmaxf(a, b) : a + b; }

template <typename T>
VB_DEVICE float ldf(const T* p);
template <>
VB_DEVICE float ldf<float>(const float* p) { return *p; }
template <>
VB_DEVICE float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// P: 0 = inf (max |x|), 1, 2
template <typename T, int P>
__global__ void __launch_bounds__(256) ragged_norm_rows_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows, int64_t C) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    const T* row = x + r * C;
    float acc = 0.f;
    for (int64_t c = lane; c < C; c += 32) acc = comb<P>(acc, powp<P>(ldf<T>(row + c)));
    acc = P == 0 ? warp_max(acc) : warp_sum(acc);
    if (lane == 0) out[r] = acc;
  }
}

template <typename T, int P>
__global__ void __launch_bounds__(256) ragged_norm_cols_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows, int64_t C, int64_t rows_per_split) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split, r1 = min(rows, r0 + rows_per_split);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc = comb<P>(acc, powp<P>(ldf<T>(x + r * C + c)));  // consecutive threads read consecutive columns
  if (P == 0) atomicMax(reinterpret_cast<int*>(out + c), __float_as_int(acc));              // non-negative floats order like ints
  else atomicAdd(out + c, acc);
}

template <typename T>
void launch(const at::Tensor& x, at::Tensor& out, int64_t rows, int64_t C, int mode, int p) {
  auto stream = at::cuda::getCurrentCUDAStream();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const T* xp = reinterpret_cast<const T*>(x.data_ptr());
  float* op = out.data_ptr<float>();
  if (mode == 0) {
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)sms * 8);
    if (p == 2) ragged_norm_rows_kernel<T, 2><<<grid, 256, 0, stream>>>(xp, op, rows, C);
    else if (p == 1) ragged_norm_rows_kernel<T, 1><<<grid, 256, 0, stream>>>(xp, op, rows, C);
    else ragged_norm_rows_kernel<T, 0><<<grid, 256, 0, stream>>>(xp, op, rows, C);
  } else {
    const int gx = (int)((C + 255) / 256);
    const int splits = (int)std::max<int64_t>(1, std::min<int64_t>(rows, std::max<int64_t>(1, (int64_t)sms * 4 / gx)));
    const int64_t rps = (rows + splits - 1) / splits;
    dim3 grid(gx, (unsigned)((rows + rps - 1) / rps));
    if (p == 2) ragged_norm_cols_kernel<T, 2><<<grid, 256, 0, stream>>>(xp, op, rows, C, rps);
    else if (p == 1) ragged_norm_cols_kernel<T, 1><<<grid, 256, 0, stream>>>(xp, op, rows, C, rps);
    else ragged_norm_cols_kernel<T, 0><<<grid, 256, 0, stream>>>(xp, op, rows, C, rps);
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace

// x [rows * C] (contiguous local shard), mode 0 -> out [rows], mode 1 -> out [C] (must be zero-filled by the caller); p: 0 = inf, 1, 2
void ragged_norm_partial(const at::Tensor& x, at::Tensor out, int64_t rows, int64_t C, int64_t mode, int64_t p) {
  TORCH_CHECK(x.is_cuda() && x.is_contiguous() && out.is_cuda() && out.scalar_type() == at::kFloat && out.is_contiguous());
  TORCH_CHECK(x.numel() == rows * C && (mode == 0 ? out.numel() == rows : out.numel() == C) && (p == 0 || p == 1 || p == 2));
  if (rows == 0 || C == 0) return;
  c10::cuda::CUDAGuard guard(x.device());
  if (x.scalar_type() == at::kBFloat16) launch<__nv_bfloat16>(x, out, rows, C, (int)mode, (int)p);
  else if (x.scalar_type() == at::kFloat) launch<float>(x, out, rows, C, (int)mode, (int)p);
  else TORCH_CHECK(false, "ragged_norm_partial: bf16 or fp32 input expected");
}
