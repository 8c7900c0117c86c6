// Counter-based sharded random fill: every element of a shard gets exactly the value a single device would have
// produced for its global position.  Element with global linear index g uses Philox4x32-10 counter (g/4 + offset)
// and output lane g%4 — computed directly, no curand_init skip-ahead per element (the reference's patched aten
// kernels call curand_init(seed, virtual_thread, offset) for every element,
// legacy/patches/patched_pytorch_v2.2.1_rc3.patch:299-448).  Specification: vescale_b200/dtensor/random.py.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <cuda_fp16.h>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kMaxDim = 8;
struct BoxDesc {
  int ndim;
  long long size[kMaxDim];      // box extent per dim
  long long goff[kMaxDim];      // global offset per dim
  long long gstride[kMaxDim];   // global (row-major) strides in elements
  long long lstride[kMaxDim];   // local tensor strides in elements
  long long lbase;              // local element offset of the box origin
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

template <typename T>
__global__ void philox_fill_kernel(T* __restrict__ data, BoxDesc box, long long numel, uint64_t seed, uint64_t offset, int normal, float a, float b) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, g = 0, l = box.lbase;
#pragma unroll 1
    for (int d = box.ndim - 1; d >= 0; --d) {
      const long long idx = rem % box.size[d];
      rem /= box.size[d];
      g += (box.goff[d] + idx) * box.gstride[d];
      l += idx * box.lstride[d];
    }
    const uint64_t ctr = (uint64_t)(g / 4) + offset;
    const int lane = (int)(g % 4);
    uint32_t r[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    float v;
    if (!normal) {
      v = (float)(r[lane] >> 8) * (1.0f / 16777216.0f) * (b - a) + a;
    } else {
      uint32_t r2[4];
      philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r2);
      const float u1 = ((float)(r[lane] >> 8) + 1.0f) * (1.0f / 16777216.0f);
      const float u2 = (float)(r2[lane] >> 8) * (1.0f / 16777216.0f);
      const float mag = (float)sqrt(-2.0 * log((double)u1));
      v = mag * (float)cos(6.283185307179586 * (double)u2) * b + a;  // a = mean, b = std
    }
    if constexpr (sizeof(T) == 4) {
      data[l] = v;
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      data[l] = __float2bfloat16(v);
    } else {
      data[l] = __float2half(v);
    }
  }
}

// Sharded fused dropout (the reference's patched ``fused_dropout_kernel``, legacy/patches/patched_pytorch_v2.2.1_rc3.patch:449-720):
// one pass reads x, draws the uniform of the element's *global* position (same counter / lane rule as the fill above, so the
// mask is what a single device would have drawn), writes x * keep / (1 - p) and the boolean mask.
template <typename T>
__global__ void philox_dropout_kernel(const T* __restrict__ x, T* __restrict__ out, bool* __restrict__ mask, BoxDesc box, long long numel, uint64_t seed,
                                      uint64_t offset, float p, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, g = 0, l = box.lbase;
#pragma unroll 1
    for (int d = box.ndim - 1; d >= 0; --d) {
      const long long idx = rem % box.size[d];
      rem /= box.size[d];
      g += (box.goff[d] + idx) * box.gstride[d];
      l += idx * box.lstride[d];
    }
    const uint64_t ctr = (uint64_t)(g / 4) + offset;
    uint32_t r[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const float u = (float)(r[(int)(g % 4)] >> 8) * (1.0f / 16777216.0f);
    const bool keep = u >= p;
    mask[l] = keep;
    float v;
    if constexpr (sizeof(T) == 4) {
      v = x[l];
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      v = __bfloat162float(x[l]);
    } else {
      v = __half2float(x[l]);
    }
    v = keep ? v * scale : 0.f;
    if constexpr (sizeof(T) == 4) {
      out[l] = v;
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      out[l] = __float2bfloat16(v);
    } else {
      out[l] = __float2half(v);
    }
  }
}

}  // namespace

// x, out, mask: same shape and (contiguous) layout; the box describes which global positions the local elements hold
void philox_dropout_box(const at::Tensor& x, at::Tensor out, at::Tensor mask, std::vector<int64_t> size, std::vector<int64_t> goff, std::vector<int64_t> gstride,
                        std::vector<int64_t> lstride, int64_t lbase, int64_t seed, int64_t offset, double p) {
  TORCH_CHECK(x.is_cuda() && out.is_cuda() && mask.is_cuda() && mask.scalar_type() == at::kBool && x.scalar_type() == out.scalar_type());
  TORCH_CHECK(x.is_contiguous() && out.is_contiguous() && mask.is_contiguous() && x.numel() == out.numel() && x.numel() == mask.numel());
  TORCH_CHECK(size.size() <= (size_t)kMaxDim && size.size() == goff.size() && size.size() == gstride.size() && size.size() == lstride.size());
  TORCH_CHECK(p >= 0.0 && p < 1.0, "dropout probability must be in [0, 1)");
  BoxDesc bd{};
  bd.ndim = size.size();
  long long numel = 1;
  for (int d = 0; d < bd.ndim; ++d) {
    bd.size[d] = size[d], bd.goff[d] = goff[d], bd.gstride[d] = gstride[d], bd.lstride[d] = lstride[d];
    numel *= size[d];
  }
  bd.lbase = lbase;
  if (numel == 0) return;
  c10::cuda::CUDAGuard guard(x.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<long long>((numel + 255) / 256, (long long)sms * 16);
  auto stream = at::cuda::getCurrentCUDAStream();
  const float scale = (float)(1.0 / (1.0 - p));
  if (x.scalar_type() == at::kFloat)
    philox_dropout_kernel<float><<<grid, 256, 0, stream>>>(x.data_ptr<float>(), out.data_ptr<float>(), mask.data_ptr<bool>(), bd, numel, (uint64_t)seed, (uint64_t)offset, (float)p, scale);
  else if (x.scalar_type() == at::kBFloat16)
    philox_dropout_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x.data_ptr(), (__nv_bfloat16*)out.data_ptr(), mask.data_ptr<bool>(), bd, numel, (uint64_t)seed, (uint64_t)offset, (float)p, scale);
  else if (x.scalar_type() == at::kHalf)
    philox_dropout_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)x.data_ptr(), (__half*)out.data_ptr(), mask.data_ptr<bool>(), bd, numel, (uint64_t)seed, (uint64_t)offset, (float)p, scale);
  else
    TORCH_CHECK(false, "philox_dropout: unsupported dtype");
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void philox_fill_box(at::Tensor local, std::vector<int64_t> size, std::vector<int64_t> goff, std::vector<int64_t> gstride, std::vector<int64_t> lstride,
                     int64_t lbase, int64_t seed, int64_t offset, bool normal, double a, double b) {
  TORCH_CHECK(local.is_cuda() && size.size() <= (size_t)kMaxDim && size.size() == goff.size() && size.size() == gstride.size() && size.size() == lstride.size());
  BoxDesc bd{};
  bd.ndim = size.size();
  long long numel = 1;
  for (int d = 0; d < bd.ndim; ++d) {
    bd.size[d] = size[d], bd.goff[d] = goff[d], bd.gstride[d] = gstride[d], bd.lstride[d] = lstride[d];
    numel *= size[d];
  }
  bd.lbase = lbase;
  if (numel == 0) return;
  c10::cuda::CUDAGuard guard(local.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<long long>((numel + 255) / 256, (long long)sms * 16);
  auto stream = at::cuda::getCurrentCUDAStream();
  if (local.scalar_type() == at::kFloat)
    philox_fill_kernel<float><<<grid, 256, 0, stream>>>(local.data_ptr<float>(), bd, numel, (uint64_t)seed, (uint64_t)offset, normal ? 1 : 0, (float)a, (float)b);
  else if (local.scalar_type() == at::kBFloat16)
    philox_fill_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)local.data_ptr(), bd, numel, (uint64_t)seed, (uint64_t)offset, normal ? 1 : 0, (float)a, (float)b);
  else if (local.scalar_type() == at::kHalf)
    philox_fill_kernel<__half><<<grid, 256, 0, stream>>>((__half*)local.data_ptr(), bd, numel, (uint64_t)seed, (uint64_t)offset, normal ? 1 : 0, (float)a, (float)b);
  else
    TORCH_CHECK(false, "philox_fill: unsupported dtype");
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
